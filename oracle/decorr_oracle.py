"""
TEST INFRASTRUCTURE -- CPU restatement (NumPy) of the reference's noise-decorrelation path, used only by tests/, smoke()
and bench.py's CPU legs as the checker of the CUDA routines sfftb_decorr / sfftb_convolve.  Never imported by the product.

Restates, line by line:
  * ConvKernel_Convertion.CSZ / iCSZ            sfft/utils/ConvKernelConvertion.py:15-31 (also sfft/BSplineSFFT.py:4739-4753)
  * DeCorrelation_Calculator.DCC                sfft/utils/DeCorrelationCalculator.py:11-103
  * BSpline_DeCorrelation.BDC                   sfft/BSplineSFFT.py:4757-4868 (DCC + denominator clipping)
  * PureCupy_DeCorrelation_Calculator.PCDC      sfft/utils/PureCupyDeCorrelationCalculator.py:46-125
  * PureCupy_FFTKits.KERNEL_CSZ / KERNEL_CSZ_INV / FFT_CONVOLVE   sfft/utils/PureCupyFFTKits.py:37-105

Pinned: dcc() against tests/golden/decorr_cases.npz -- the UNMODIFIED reference DCC run on the reference's own match kernels
(test/difference_noise_decorrelation/input_data) with the sky sigmas of its SkyLevel_Estimator, and against the reference's
known answer 4check/DeCorrKernel.fits (tests/golden/make_golden_decorr.py).  PCDC / FFT_CONVOLVE need CuPy in the reference and
are pinned through their shared arithmetic with DCC (same denominator) and by structure tests -- parity unpinned for the
PCDC-only output modes.
"""
import math
import numpy as np

UMK = np.array([[0, 0, 0], [0, 1, 0], [0, 0, 0]], dtype=float)


def csz(ConvKernel, N0, N1):
    """Circular shift + tail zero padding (ConvKernelConvertion.py:15-21)."""
    L0, L1 = ConvKernel.shape
    w0, w1 = (L0 - 1) // 2, (L1 - 1) // 2
    tz = np.pad(ConvKernel, ((0, N0 - L0), (0, N1 - L1)), 'constant', constant_values=(0, 0))
    return np.roll(np.roll(tz, -w0, axis=0), -w1, axis=1)


def icsz(KIMG, L0, L1):
    """Inverse circular shift + tail truncation; also the lost weight (ConvKernelConvertion.py:23-31)."""
    w0, w1 = (L0 - 1) // 2, (L1 - 1) // 2
    k = np.roll(np.roll(KIMG, w1, axis=1), w0, axis=0)
    ck = k[:L0, :L1]
    return ck, 1.0 - np.sum(np.abs(ck)) / np.sum(np.abs(k))


def _check_modes(MK_JLst, MK_ILst, MK_Fin):
    NumI, NumJ = len(MK_ILst), len(MK_JLst)
    if NumI == 0:
        if NumJ < 2:
            raise Exception('MeLOn ERROR: Image-Stacking Mode requires at least 2 J-images!')
        if sum(m is not None for m in MK_JLst) == 0:
            raise Exception('MeLOn ERROR: Image-Stacking Mode requires at least 1 not-None J-kernel!')
        return 'Image-Stacking'
    if NumJ == 0:
        raise Exception('MeLOn ERROR: Image-Subtraction Mode requires at least 1 I-image & 1 J-image!')
    if sum(m is not None for m in list(MK_JLst) + list(MK_ILst) + [MK_Fin]) == 0:
        raise Exception('MeLOn ERROR: Image-Subtraction Mode requires at least 1 not-None J/I/Fin-kernel!')
    return 'Image-Subtraction'


def denominator(N0, N1, MK_JLst, SkySig_JLst, MK_ILst, SkySig_ILst, MK_Fin):
    """DeNo map on the (N0, N1) Fourier grid (DeCorrelationCalculator.py:68-96; PCDC :75-101)."""
    NumI, NumJ = len(MK_ILst), len(MK_JLst)

    def k2(MK):
        kft = np.fft.fft2(csz(UMK if MK is None else np.asarray(MK, float), N0, N1))
        return (np.conj(kft) * kft).real
    fin2 = k2(MK_Fin)
    deno = 0.0
    for MKj, s in zip(MK_JLst, SkySig_JLst):
        deno = deno + (s ** 2 * k2(MKj)) / NumJ ** 2
    for MKi, s in zip(MK_ILst, SkySig_ILst):
        deno = deno + (s ** 2 * k2(MKi) * fin2) / NumI ** 2
    return deno


def dcc_sizes(MK_Queue, KERatio):
    """(L0_KDeCo, L1_KDeCo, N0, N1) of DCC / BDC (DeCorrelationCalculator.py:55-66)."""
    sh0 = [m.shape[0] for m in MK_Queue if m is not None]
    sh1 = [m.shape[1] for m in MK_Queue if m is not None]
    L0 = int(round(KERatio * np.max(sh0)))
    L1 = int(round(KERatio * np.max(sh1)))
    L0 += 1 - L0 % 2
    L1 += 1 - L1 % 2
    N0 = 2 ** (math.ceil(np.log2(np.max(sh0))) + 1)
    N1 = 2 ** (math.ceil(np.log2(np.max(sh1))) + 1)
    return L0, L1, N0, N1


def dcc(MK_JLst, SkySig_JLst, MK_ILst=(), SkySig_ILst=(), MK_Fin=None, KERatio=2.0, DENO_CLIP_RATIO=None):
    """DCC (DENO_CLIP_RATIO=None) or BDC (DENO_CLIP_RATIO=1e5 by default there): the real-space decorrelation kernel with unit
    sum; returns (KDeCo, lost_weight)."""
    MK_JLst, MK_ILst = list(MK_JLst), list(MK_ILst)
    mode = _check_modes(MK_JLst, MK_ILst, MK_Fin)
    queue = MK_JLst + ([MK_Fin] + MK_ILst if mode == 'Image-Subtraction' else [])
    L0, L1, N0, N1 = dcc_sizes(queue, KERatio)
    deno = denominator(N0, N1, MK_JLst, SkySig_JLst, MK_ILst if mode == 'Image-Subtraction' else [], SkySig_ILst, MK_Fin)
    if DENO_CLIP_RATIO is not None:                       # BSplineSFFT.py:4838-4841
        thr = np.max(deno) / DENO_CLIP_RATIO
        deno = np.where(deno < thr, thr, deno)
    fdeco = np.sqrt(1.0 / deno)
    deco = np.fft.ifft2(fdeco).real
    k, lost = icsz(deco, L0, L1)
    return k / np.sum(k), lost


def pcdc(NX_IMG, NY_IMG, KERNEL_JQueue, BKGSIG_JQueue, KERNEL_IQueue=(), BKGSIG_IQueue=(), MATCH_KERNEL=None, REAL_OUTPUT=False,
         REAL_OUTPUT_SIZE=None, NORMALIZE_OUTPUT=True):
    """PCDC (PureCupyDeCorrelationCalculator.py:46-125) on the host."""
    KJ, KI = list(KERNEL_JQueue), list(KERNEL_IQueue)
    if len(KI) == 0:
        if len(KJ) < 2:
            raise Exception('MeLOn ERROR: IMAGE-STACKING MODE Requires at least 2 J-IMAGE!')
        if sum(k is not None for k in KJ) == 0:
            raise Exception('MeLOn ERROR: IMAGE-STACKING MODE Requires at least 1 non-None J-KERNEL!')
    elif sum(k is not None for k in KJ + KI + [MATCH_KERNEL]) == 0:
        raise Exception('MeLOn ERROR: IMAGE-SUBTRACTION MODE Requires at least 1 non-None J/I/MATCH-KERNEL!')
    deno = denominator(NX_IMG, NY_IMG, KJ, BKGSIG_JQueue, KI, BKGSIG_IQueue, MATCH_KERNEL)
    fk = 1.0 / np.sqrt(deno)
    if not REAL_OUTPUT:
        return fk * (1.0 / fk[0, 0]) if NORMALIZE_OUTPUT else fk
    k = icsz(np.fft.ifft2(fk).real, REAL_OUTPUT_SIZE[0], REAL_OUTPUT_SIZE[1])[0]
    return k * (1.0 / np.sum(k)) if NORMALIZE_OUTPUT else k


def fft_convolve(PixA_Inp, KERNEL, PAD_FILL_VALUE=0.0, NAN_FILL_VALUE=0.0, NORMALIZE_KERNEL=False):
    """FFT_CONVOLVE (PureCupyFFTKits.py:71-105)."""
    N0, N1 = PixA_Inp.shape
    L0, L1 = KERNEL.shape
    assert L0 % 2 == 1 and L1 % 2 == 1
    W0, W1 = (L0 - 1) // 2, (L1 - 1) // 2
    e = np.pad(np.asarray(PixA_Inp, float), ((W0, W0), (W1, W1)), mode='constant', constant_values=PAD_FILL_VALUE)
    if NAN_FILL_VALUE is not None:
        e[np.isnan(e)] = NAN_FILL_VALUE
    k = np.asarray(KERNEL, float)
    if NORMALIZE_KERNEL:
        k = k / np.sum(k)
    kimg = csz(k, N0 + 2 * W0, N1 + 2 * W1)
    out = np.fft.ifft2(np.fft.fft2(e) * np.fft.fft2(kimg)).real
    return out[W0: W0 + N0, W1: W1 + N1]


def gsvc(PixA_obj, AllocatedL, KerStack, nan_fill_value=0.0, normalize_kernel=True):
    """BSpline_GridConvolve.GSVC_GPU with use_fft=False (sfft/BSplineSFFT.py:4951-5008): per cell, the mini image (cell extended by
    w + 1, clipped to the image) convolved with scipy's convolve2d(mode='same', boundary='fill', fillvalue=0) -- cupyx's convolve2d
    is its mirror -- and the cell pasted back.  Parity unpinned (the reference needs CuPy / astropy); restated line by line."""
    from scipy.signal import convolve2d
    PixA_in = np.array(PixA_obj, dtype=float)
    PixA_in[np.isnan(PixA_in)] = nan_fill_value
    N0, N1 = PixA_in.shape
    Nseg, L0, L1 = KerStack.shape
    w0, w1 = int((L0 - 1) / 2), int((L1 - 1) / 2)
    IBx, IBy = w0 + 1, w1 + 1
    K = np.asarray(KerStack, float)
    if normalize_kernel:
        K = K / np.sum(K, axis=(1, 2))[:, np.newaxis, np.newaxis]
    out = np.zeros((N0, N1))
    for idx in range(Nseg):
        lX, lY = np.where(AllocatedL == idx)
        if lX.size == 0:
            continue
        xs, xe, ys, ye = lX.min(), lX.max(), lY.min(), lY.max()
        xEs, xEe = max([0, xs - IBx]), min([N0 - 1, xe + IBx])
        yEs, yEe = max([0, ys - IBy]), min([N1 - 1, ye + IBy])
        c = convolve2d(PixA_in[xEs: xEe + 1, yEs: yEe + 1], K[idx], mode='same', boundary='fill', fillvalue=0.0)
        out[xs: xe + 1, ys: ye + 1] = c[xs - xEs: (xs - xEs) + (xe + 1 - xs), ys - yEs: (ys - yEs) + (ye + 1 - ys)]
    return out
