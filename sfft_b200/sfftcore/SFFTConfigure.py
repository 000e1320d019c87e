"""
SingleSFFTConfigure.SSC -- host-side mirror of sfft/sfftcore/SFFTConfigure.py:1371-1395.

Same call signature and the same SFFTParam_dict keys (SFFTConfigure.py:35-75); the second element
of the returned SFFTConfig is opaque to callers in the reference (a dict of JIT-compiled kernels)
and here carries the native plan.  Backend strings: 'B200' (this package) and 'Cupy' (accepted as
an alias so that unmodified callers such as Easy*Packet, which default to 'Cupy', land on the
B200 path).  'Numpy' is refused: the product has no CPU path.
"""
import numpy as np

from ..plan import Plan

__all__ = ['SingleSFFTConfigure', 'SingleSFFTConfigure_B200']

_BACKENDS = ('B200', 'Cupy')


def _current_device():
    try:
        import sys
        torch = sys.modules.get('torch')
        if torch is not None and torch.cuda.is_available() and torch.cuda.is_initialized():
            return int(torch.cuda.current_device())
    except Exception:
        pass
    return 0


class SingleSFFTConfigure_B200:
    @staticmethod
    def SSCB(NX, NY, KerHW, KerPolyOrder=2, BGPolyOrder=2, ConstPhotRatio=True, VERBOSE_LEVEL=2,
             CUDA_DEVICE=None, STORAGE='fp64', FOLD=0):
        N0, N1 = int(NX), int(NY)
        w0, w1 = int(KerHW), int(KerHW)
        DK, DB = int(KerPolyOrder), int(BGPolyOrder)
        MaxThreadPerB = 8      # kept for dict compatibility (SFFTConfigure.py:15); not a launch parameter here
        if DK not in [0, 1, 2, 3]:
            raise Exception('MeLOn ERROR: Input KerPolyOrder should be 0/1/2/3!')          # :19-21
        if DB not in [0, 1, 2, 3]:
            raise Exception('MeLOn ERROR: Input BGPolyOrder should be 0/1/2/3!')           # :23-25
        if (N0 < MaxThreadPerB) or (N1 < MaxThreadPerB):
            raise Exception('MeLOn ERROR: Input Image has dramatically small size!')       # :27-28
        if VERBOSE_LEVEL in [1, 2]:
            print('\n --//--//--//--//-- TRIGGER SFFT COMPILATION --//--//--//--//-- ')
            print('\n ---//--- KerPolyOrder %d | BGPolyOrder %d | KerHW [%d] ---//--- ' % (DK, DB, w0))

        L0, L1 = 2 * w0 + 1, 2 * w1 + 1
        Fab = L0 * L1
        Fij = (DK + 1) * (DK + 2) // 2
        Fpq = (DB + 1) * (DB + 2) // 2
        SCALE = np.float64(1 / (N0 * N1))
        SCALE_L = np.float64(1 / SCALE)
        Fijab = Fij * Fab
        NEQ = Fijab + Fpq
        SFFTParam_dict = dict(
            N0=N0, N1=N1, w0=w0, w1=w1, DK=DK, DB=DB, ConstPhotRatio=ConstPhotRatio, MaxThreadPerB=MaxThreadPerB,
            L0=L0, L1=L1, Fab=Fab, Fij=Fij, Fpq=Fpq, SCALE=SCALE, SCALE_L=SCALE_L, NEQ=NEQ, Fijab=Fijab,
            NEQ_FSfree=NEQ - (Fij - 1), FOMG=Fij ** 2, FGAM=Fij * Fpq, FTHE=Fij, FPSI=Fpq * Fij, FPHI=Fpq ** 2,
            FDEL=Fpq)
        device = _current_device() if CUDA_DEVICE is None else int(CUDA_DEVICE)
        plan = Plan(N0, N1, w0, w1, DK, DB, ConstPhotRatio, device=device, storage=STORAGE, fold=FOLD)
        SFFTModule_dict = {'BACKEND': 'B200', 'plan': plan, 'device': device, 'storage': STORAGE}
        if VERBOSE_LEVEL in [1, 2]:
            print('\n --//--//--//--//-- EXIT SFFT COMPILATION --//--//--//--//-- ')
        return (SFFTParam_dict, SFFTModule_dict)


class SingleSFFTConfigure:
    @staticmethod
    def SSC(NX, NY, KerHW, KerPolyOrder=2, BGPolyOrder=2, ConstPhotRatio=True,
            BACKEND_4SUBTRACT='B200', NUM_CPU_THREADS_4SUBTRACT=8, NUMBA_CACHE=True, VERBOSE_LEVEL=2,
            CUDA_DEVICE=None, STORAGE='fp64', FOLD=0):
        """Arguments as in the reference (SFFTConfigure.py:1374-1385); NUM_CPU_THREADS_4SUBTRACT and
        NUMBA_CACHE are accepted and ignored.  Extra keyword-only-by-convention arguments:
        CUDA_DEVICE (default: the current torch device, else 0), STORAGE ('fp64' | 'fp32': precision of
        the spectra kept in HBM; arithmetic is fp64 in both), FOLD (column-pass fold factor, 0 = auto)."""
        if BACKEND_4SUBTRACT not in _BACKENDS:
            raise Exception("MeLOn ERROR: BACKEND_4SUBTRACT=%r is not available in sfft_b200 (use 'B200'); "
                            "there is no CPU backend" % (BACKEND_4SUBTRACT,))
        return SingleSFFTConfigure_B200.SSCB(NX=NX, NY=NY, KerHW=KerHW, KerPolyOrder=KerPolyOrder,
                                             BGPolyOrder=BGPolyOrder, ConstPhotRatio=ConstPhotRatio,
                                             VERBOSE_LEVEL=VERBOSE_LEVEL, CUDA_DEVICE=CUDA_DEVICE, STORAGE=STORAGE,
                                             FOLD=FOLD)
