"""sfft_b200 -- B200 (sm_100a) native SFFT subtraction core behind the reference's call signatures
(sfft/__init__.py:16-24: Customized_Packet, PureCupy_Customized_Packet; sfft/sfftcore/__init__.py:7-8)."""
from .CustomizedPacket import Customized_Packet
from .PureCupyCustomizedPacket import PureCupy_Customized_Packet
from .sfftcore import (SingleSFFTConfigure, ElementalSFFTSubtract, GeneralSFFTSubtract,
                       GeneralSFFTSubtract_PureCupy)

__version__ = '0.1.0'
