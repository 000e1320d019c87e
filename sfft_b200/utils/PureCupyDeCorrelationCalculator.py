"""PureCupy_DeCorrelation_Calculator.PCDC (sfft/utils/PureCupyDeCorrelationCalculator.py:46-125) on the B200 core.  Same
signature; kernels may be torch CUDA tensors (what the PureCupy-style pipelines hold) or host arrays, the result is a torch
CUDA tensor like the reference's CuPy array.  The image-sized fft2 of every padded kernel is replaced by the closed-form
spectrum of its taps (csrc/tu_decorr.cu), so the only image-sized array is the output itself."""
import numpy as np

from ._decorr import decorr

__all__ = ['PureCupy_DeCorrelation_Calculator']


class PureCupy_DeCorrelation_Calculator:
    @staticmethod
    def PCDC(NX_IMG, NY_IMG, KERNEL_GPU_JQueue, BKGSIG_JQueue, KERNEL_GPU_IQueue=[], BKGSIG_IQueue=[], MATCH_KERNEL_GPU=None,
             REAL_OUTPUT=False, REAL_OUTPUT_SIZE=None, NORMALIZE_OUTPUT=True, VERBOSE_LEVEL=2, CUDA_DEVICE='0'):
        """Decorrelation Kernel Calculation in Pure Cupy"""
        import torch
        NUM_I, NUM_J = len(KERNEL_GPU_IQueue), len(KERNEL_GPU_JQueue)
        assert NUM_J > 0
        if NUM_I == 0:
            if NUM_J < 2:
                raise Exception('MeLOn ERROR: %s' % 'IMAGE-STACKING MODE Requires at least 2 J-IMAGE!')
            if np.sum([K is not None for K in KERNEL_GPU_JQueue]) == 0:
                raise Exception('MeLOn ERROR: %s' % 'IMAGE-STACKING MODE Requires at least 1 non-None J-KERNEL!')
        if NUM_I >= 1:
            _Q = list(KERNEL_GPU_JQueue) + list(KERNEL_GPU_IQueue) + [MATCH_KERNEL_GPU]
            if np.sum([K is not None for K in _Q]) == 0:
                raise Exception('MeLOn ERROR: %s' % 'IMAGE-SUBTRACTION MODE Requires at least 1 non-None J/I/MATCH-KERNEL!')
        dev = int(CUDA_DEVICE)
        for K in list(KERNEL_GPU_JQueue) + list(KERNEL_GPU_IQueue) + [MATCH_KERNEL_GPU]:
            if K is not None and hasattr(K, 'is_cuda') and K.is_cuda:
                dev = K.device.index or 0
        tdev = torch.device('cuda', dev)
        stream = torch.cuda.current_stream(tdev).cuda_stream or 0x1
        if not REAL_OUTPUT:
            FKDECO_GPU = torch.empty((NX_IMG, NY_IMG), dtype=torch.float64, device=tdev)
            decorr(NX_IMG, NY_IMG, list(KERNEL_GPU_JQueue), list(BKGSIG_JQueue), list(KERNEL_GPU_IQueue), list(BKGSIG_IQueue),
                   MATCH_KERNEL_GPU, real_size=None, normalize=NORMALIZE_OUTPUT, device=dev, stream=stream, out_device_ptr=FKDECO_GPU.data_ptr())
            return FKDECO_GPU
        if NORMALIZE_OUTPUT:
            assert REAL_OUTPUT_SIZE is not None
        K, lost = decorr(NX_IMG, NY_IMG, list(KERNEL_GPU_JQueue), list(BKGSIG_JQueue), list(KERNEL_GPU_IQueue), list(BKGSIG_IQueue),
                         MATCH_KERNEL_GPU, real_size=REAL_OUTPUT_SIZE, normalize=NORMALIZE_OUTPUT, device=dev, stream=stream)
        if VERBOSE_LEVEL in [1, 2] and lost == lost:
            print("MeLOn CheckPoint: Kernel Truncation Loses APE = [%.4f %s] " % (lost * 100, '%'))
        return torch.from_numpy(K).to(tdev)
