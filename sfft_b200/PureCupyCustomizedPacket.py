"""
PureCupy_Customized_Packet.PCCP -- mirror of sfft/PureCupyCustomizedPacket.py:41-187: device arrays in,
device arrays out, no FITS, no host round trip.  "Cupy" in the name is the reference's; arrays here are torch
CUDA tensors or anything exposing __cuda_array_interface__.
"""
from .sfftcore.SFFTConfigure import SingleSFFTConfigure
from .sfftcore.SFFTSubtract import GeneralSFFTSubtract_PureCupy

__all__ = ['PureCupy_Customized_Packet']


class PureCupy_Customized_Packet:
    @staticmethod
    def PCCP(PixA_REF_GPU, PixA_SCI_GPU, PixA_mREF_GPU, PixA_mSCI_GPU, ForceConv, GKerHW,
             KerPolyOrder=2, BGPolyOrder=2, ConstPhotRatio=True, CUDA_DEVICE_4SUBTRACT='0', VERBOSE_LEVEL=2,
             STORAGE='fp64', SFFTConfig=None):
        import torch
        arrs = (PixA_REF_GPU, PixA_SCI_GPU, PixA_mREF_GPU, PixA_mSCI_GPU)
        for a in arrs:                                                                  # :105-116
            assert len(a.shape) == 2
            cai = a.__cuda_array_interface__
            assert cai['typestr'] in ('<f8', '<f4')
            assert cai.get('strides') is None
        for m in (PixA_mREF_GPU, PixA_mSCI_GPU):
            assert not bool(torch.isnan(torch.as_tensor(m)).any())
        assert ForceConv in ['REF', 'SCI']
        ConvdSide, KerHW = ForceConv, GKerHW
        if SFFTConfig is None:
            SFFTConfig = SingleSFFTConfigure.SSC(NX=PixA_REF_GPU.shape[0], NY=PixA_REF_GPU.shape[1], KerHW=KerHW,
                                                 KerPolyOrder=KerPolyOrder, BGPolyOrder=BGPolyOrder,
                                                 ConstPhotRatio=ConstPhotRatio, BACKEND_4SUBTRACT='B200',
                                                 VERBOSE_LEVEL=VERBOSE_LEVEL, CUDA_DEVICE=int(CUDA_DEVICE_4SUBTRACT),
                                                 STORAGE=STORAGE)
        if ConvdSide == 'REF':
            I, J, mI, mJ = PixA_REF_GPU, PixA_SCI_GPU, PixA_mREF_GPU, PixA_mSCI_GPU
        else:
            I, J, mI, mJ = PixA_SCI_GPU, PixA_REF_GPU, PixA_mSCI_GPU, PixA_mREF_GPU
        Solution_GPU, PixA_DIFF_GPU, _ = GeneralSFFTSubtract_PureCupy.GSS(
            PixA_I_GPU=I, PixA_J_GPU=J, PixA_mI_GPU=mI, PixA_mJ_GPU=mJ, SFFTConfig=SFFTConfig,
            ContamMask_I_GPU=None, VERBOSE_LEVEL=VERBOSE_LEVEL)
        if ConvdSide == 'SCI':
            PixA_DIFF_GPU *= -1.0                                                        # :183-185
        return Solution_GPU, PixA_DIFF_GPU
