"""Solution consumers (sfft/utils/SFFTSolutionReader.py)."""
from .SFFTSolutionReader import (Read_SFFTSolution, SVKDict_ST2SFFT, SVKDict_SFFT2ST, Realize_MatchingKernel,
                                 Realize_FluxScaling)
