"""DeCorrelation_Calculator.DCC (sfft/utils/DeCorrelationCalculator.py:11-103) on the B200 core: same signature, the kernel
spectra and the truncated inverse transform are evaluated by sfftb_decorr (csrc/tu_decorr.cu)."""
import math
import numpy as np

from ._decorr import decorr, check_modes

__all__ = ['DeCorrelation_Calculator']


def _sizes(MK_Queue, KERatio):
    sh0 = [MK.shape[0] for MK in MK_Queue if MK is not None]
    sh1 = [MK.shape[1] for MK in MK_Queue if MK is not None]
    L0 = int(round(KERatio * np.max(sh0)))
    L1 = int(round(KERatio * np.max(sh1)))
    if L0 % 2 == 0: L0 += 1
    if L1 % 2 == 0: L1 += 1
    # trivial image size, just typically larger than the kernel size (:64-66)
    N0 = 2 ** (math.ceil(np.log2(np.max(sh0))) + 1)
    N1 = 2 ** (math.ceil(np.log2(np.max(sh1))) + 1)
    return L0, L1, N0, N1


class DeCorrelation_Calculator:
    @staticmethod
    def DCC(MK_JLst, SkySig_JLst, MK_ILst=[], SkySig_ILst=[], MK_Fin=None, KERatio=2.0, VERBOSE_LEVEL=2, CUDA_DEVICE='0',
            _CLIP_RATIO=0.0):
        sub = check_modes(MK_JLst, MK_ILst, MK_Fin)
        MK_Queue = list(MK_JLst)
        if sub: MK_Queue += [MK_Fin] + list(MK_ILst)
        L0, L1, N0, N1 = _sizes(MK_Queue, KERatio)
        if VERBOSE_LEVEL in [1, 2]:
            print('MeLOn CheckPoint: DeCorrelation Kernel with size [%d, %d]' % (L0, L1))
        KDeCo, lost = decorr(N0, N1, list(MK_JLst), list(SkySig_JLst), list(MK_ILst) if sub else [], list(SkySig_ILst) if sub else [],
                             MK_Fin if sub else None, clip_ratio=_CLIP_RATIO, real_size=(L0, L1), normalize=True, device=int(CUDA_DEVICE))
        if VERBOSE_LEVEL in [1, 2]:
            print('MeLOn CheckPoint: Tail-Truncation Lost-Weight [%.4f %s] (Absolute Percentage Error)' % (lost * 100, '%'))
        return KDeCo
