"""PureCupy_FFTKits (sfft/utils/PureCupyFFTKits.py:37-105) for device-resident arrays (torch CUDA tensors stand where the
reference has CuPy arrays): KERNEL_CSZ / KERNEL_CSZ_INV are index bookkeeping, FFT_CONVOLVE runs sfftb_convolve -- the
convolution with the zero (constant) padding and NaN fill of the reference, evaluated directly in real space.  One
difference by construction: with NAN_FILL_VALUE=None a NaN sample spoils only the kernel footprint around it, where the
reference's FFT product returns an all-NaN image."""
import numpy as np

from .. import _lib as B

__all__ = ['PureCupy_FFTKits']


class PureCupy_FFTKits:
    @staticmethod
    def KERNEL_CSZ(KERNEL_GPU, NX_IMG, NY_IMG, NORMALIZE_KERNEL=False):
        """ Circular Shift the kernel and extend to the target size """
        import torch
        L0, L1 = KERNEL_GPU.shape
        W0, W1 = (L0 - 1) // 2, (L1 - 1) // 2
        assert L0 % 2 == 1 and L1 % 2 == 1
        K = KERNEL_GPU / KERNEL_GPU.sum() if NORMALIZE_KERNEL else KERNEL_GPU
        out = torch.zeros((NX_IMG, NY_IMG), dtype=K.dtype, device=K.device)
        out[:L0, :L1] = K
        return torch.roll(out, shifts=(-W0, -W1), dims=(0, 1))

    @staticmethod
    def KERNEL_CSZ_INV(KIMG_GPU, NX_KERN, NY_KERN, VERBOSE_LEVEL=2):
        """ Inverse Circular Shift the kernel and truncate to the target size """
        import torch
        L0, L1 = NX_KERN, NY_KERN
        W0, W1 = (L0 - 1) // 2, (L1 - 1) // 2
        assert L0 % 2 == 1 and L1 % 2 == 1
        K = torch.roll(KIMG_GPU, shifts=(W0, W1), dims=(0, 1))
        KERNEL_GPU = K[:L0, :L1]
        if VERBOSE_LEVEL in [1, 2]:
            LOSE_RATIO = 1. - float(KERNEL_GPU.abs().sum() / K.abs().sum())
            print("MeLOn CheckPoint: Kernel Truncation Loses APE = [%.4f %s] " % (LOSE_RATIO * 100, '%'))
        return KERNEL_GPU

    @staticmethod
    def FFT_CONVOLVE(PixA_Inp_GPU, KERNEL_GPU, PAD_FILL_VALUE=0., NAN_FILL_VALUE=0., NORMALIZE_KERNEL=False,
                     FORCE_OUTPUT_C_CONTIGUOUS=False, FFT_BACKEND="Cupy"):
        """ FFT Convolition """
        import torch
        x = PixA_Inp_GPU
        if not x.is_cuda:
            raise Exception('MeLOn ERROR: FFT_CONVOLVE expects a CUDA tensor')
        if x.dtype not in (torch.float32, torch.float64):
            x = x.to(torch.float64)
        x = x.contiguous()
        k = KERNEL_GPU.detach().cpu().numpy() if hasattr(KERNEL_GPU, 'detach') else np.asarray(KERNEL_GPU)
        k = np.ascontiguousarray(k, dtype=np.float64)
        L0, L1 = k.shape
        assert L0 % 2 == 1 and L1 % 2 == 1
        out = torch.empty_like(x)
        stream = torch.cuda.current_stream(x.device).cuda_stream or 0x1
        B.check(B.lib().sfftb_convolve(x.device.index or 0, stream, x.data_ptr(), B.F64 if x.dtype == torch.float64 else B.F32, x.shape[0], x.shape[1],
                                       k.ctypes.data, L0, L1, float(PAD_FILL_VALUE), 0.0 if NAN_FILL_VALUE is None else float(NAN_FILL_VALUE),
                                       0 if NAN_FILL_VALUE is None else 1, int(bool(NORMALIZE_KERNEL)), out.data_ptr(), B.MEM_DEVICE))
        return out
