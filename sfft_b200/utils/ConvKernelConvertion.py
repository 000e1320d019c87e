"""ConvKernel_Convertion (sfft/utils/ConvKernelConvertion.py:15-31): circular shift + tail zero padding of a convolution
kernel to the image size, and its inverse with tail truncation.  Host bookkeeping (a few hundred taps), NumPy like the
reference; the decorrelation routines of this package never materialise the padded kernel (csrc/tu_decorr.cu)."""
import numpy as np

__all__ = ['ConvKernel_Convertion']


class ConvKernel_Convertion:
    def CSZ(ConvKernel, N0, N1):
        L0, L1 = ConvKernel.shape
        w0, w1 = (L0 - 1) // 2, (L1 - 1) // 2
        TailZP = np.pad(ConvKernel, ((0, N0 - L0), (0, N1 - L1)), 'constant', constant_values=(0, 0))
        return np.roll(np.roll(TailZP, -w0, axis=0), -w1, axis=1)

    def iCSZ(KIMG, L0, L1):
        w0, w1 = (L0 - 1) // 2, (L1 - 1) // 2
        KIMG_iCSZ = np.roll(np.roll(KIMG, w1, axis=1), w0, axis=0)
        ConvKernel = KIMG_iCSZ[:L0, :L1]
        lost_weight = 1.0 - np.sum(np.abs(ConvKernel)) / np.sum(np.abs(KIMG_iCSZ))
        print('MeLOn CheckPoint: Tail-Truncation Lost-Weight [%.4f %s] (Absolute Percentage Error) ' % (lost_weight * 100, '%'))
        return ConvKernel
