"""Noise decorrelation: oracle pinned on the reference (CPU), CUDA routines against the oracle (GPU)."""
import os
import numpy as np
import pytest

from oracle import decorr_oracle as do

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, 'golden', 'decorr_cases.npz'))
S = [None] + [G['mk04' + t] for t in 'bcde']
R = [None] + [G['mk18' + t] for t in 'bcde']


def test_oracle_matches_reference_dcc():
    """oracle dcc() == the UNMODIFIED reference DCC on its own test inputs (tests/golden/make_golden_decorr.py), and within 2e-5
    of the decorrelation kernel the reference ships (4check/DeCorrKernel.fits, written by an older release)."""
    k, lost = do.dcc(S, G['sig04'], R, G['sig18'], G['mkfin'])
    assert np.max(np.abs(k - G['dcc_sub'])) <= 1e-15
    assert np.max(np.abs(k - G['shipped'])) <= 2e-5
    assert 0.0 < lost < 0.01
    k2, _ = do.dcc(S, G['sig04'])
    assert np.max(np.abs(k2 - G['dcc_stack'])) <= 1e-15
    assert abs(k.sum() - 1.0) < 1e-12


def test_oracle_pcdc_shares_the_denominator():
    """PCDC on the DCC grid reproduces DCC (same DeNo map, same inverse); the Fourier output is normalised at [0, 0]."""
    L0, L1, N0, N1 = do.dcc_sizes(S + [G['mkfin']] + R, 2.0)
    k = do.pcdc(N0, N1, S, G['sig04'], R, G['sig18'], G['mkfin'], REAL_OUTPUT=True, REAL_OUTPUT_SIZE=(L0, L1))
    assert np.max(np.abs(k - G['dcc_sub'])) <= 1e-14
    f = do.pcdc(96, 80, S, G['sig04'], R, G['sig18'], G['mkfin'])
    assert abs(f[0, 0] - 1.0) < 1e-15 and f.shape == (96, 80) and np.all(f > 0)


def test_oracle_fft_convolve_is_a_zero_padded_convolution():
    rng = np.random.default_rng(3)
    img = rng.normal(size=(40, 37))
    img[5, 7] = np.nan
    ker = rng.normal(size=(5, 7))
    out = do.fft_convolve(img, ker, PAD_FILL_VALUE=0.5, NAN_FILL_VALUE=2.0)
    e = np.pad(np.where(np.isnan(img), 2.0, img), ((2, 2), (3, 3)), constant_values=0.5)
    ref = np.zeros_like(img)
    for a in range(5):
        for b in range(7):
            ref += ker[a, b] * e[4 - a: 4 - a + 40, 6 - b: 6 - b + 37]
    # (the wrap-around of the circular product only touches the cropped margin)
    assert np.max(np.abs(out - ref)) < 1e-12


def test_mode_errors_like_reference():
    from sfft_b200.utils.DeCorrelationCalculator import DeCorrelation_Calculator
    with pytest.raises(Exception, match='Image-Stacking Mode requires at least 2 J-images'):
        DeCorrelation_Calculator.DCC([G['mkfin']], [1.0])
    with pytest.raises(Exception, match='at least 1 not-None J-kernel'):
        DeCorrelation_Calculator.DCC([None, None], [1.0, 1.0])
    with pytest.raises(Exception, match='at least 1 not-None J/I/Fin-kernel'):
        DeCorrelation_Calculator.DCC([None], [1.0], [None], [1.0], None)


@pytest.mark.gpu
def test_dcc_bdc_on_device_match_oracle():
    from sfft_b200.utils.DeCorrelationCalculator import DeCorrelation_Calculator
    from sfft_b200.BSplineSFFT import BSpline_DeCorrelation
    k = DeCorrelation_Calculator.DCC(S, list(G['sig04']), R, list(G['sig18']), G['mkfin'], VERBOSE_LEVEL=0)
    assert k.shape == G['dcc_sub'].shape
    assert np.max(np.abs(k - G['dcc_sub'])) <= 1e-12 * np.max(np.abs(G['dcc_sub']))
    k2 = DeCorrelation_Calculator.DCC(S, list(G['sig04']), VERBOSE_LEVEL=0)
    assert np.max(np.abs(k2 - G['dcc_stack'])) <= 1e-12 * np.max(np.abs(G['dcc_stack']))
    for clip in (1e5, 30.0):
        kb = BSpline_DeCorrelation.BDC(S, list(G['sig04']), R, list(G['sig18']), G['mkfin'], DENO_CLIP_RATIO=clip, VERBOSE_LEVEL=0)
        ob, _ = do.dcc(S, G['sig04'], R, G['sig18'], G['mkfin'], DENO_CLIP_RATIO=clip)
        assert np.max(np.abs(kb - ob)) <= 1e-12 * np.max(np.abs(ob))


@pytest.mark.gpu
def test_pcdc_and_fft_convolve_on_device_match_oracle():
    import torch
    from sfft_b200.utils.PureCupyDeCorrelationCalculator import PureCupy_DeCorrelation_Calculator as P
    from sfft_b200.utils.PureCupyFFTKits import PureCupy_FFTKits as K
    dev = torch.device('cuda', 0)
    tS = [None if m is None else torch.from_numpy(m).to(dev) for m in S]
    tR = [None if m is None else torch.from_numpy(m).to(dev) for m in R]
    fin = torch.from_numpy(G['mkfin']).to(dev)
    N0, N1 = 300, 256
    f = P.PCDC(N0, N1, tS, list(G['sig04']), tR, list(G['sig18']), MATCH_KERNEL_GPU=fin, VERBOSE_LEVEL=0)
    of = do.pcdc(N0, N1, S, G['sig04'], R, G['sig18'], G['mkfin'])
    assert f.is_cuda and tuple(f.shape) == (N0, N1)
    assert np.max(np.abs(f.cpu().numpy() - of)) <= 1e-12 * np.max(np.abs(of))
    fr = P.PCDC(N0, N1, tS, list(G['sig04']), tR, list(G['sig18']), MATCH_KERNEL_GPU=fin, NORMALIZE_OUTPUT=False, VERBOSE_LEVEL=0)
    ofr = do.pcdc(N0, N1, S, G['sig04'], R, G['sig18'], G['mkfin'], NORMALIZE_OUTPUT=False)
    assert np.max(np.abs(fr.cpu().numpy() - ofr)) <= 1e-12 * np.max(np.abs(ofr))
    k = P.PCDC(N0, N1, tS, list(G['sig04']), tR, list(G['sig18']), MATCH_KERNEL_GPU=fin, REAL_OUTPUT=True, REAL_OUTPUT_SIZE=(31, 27), VERBOSE_LEVEL=0)
    ok = do.pcdc(N0, N1, S, G['sig04'], R, G['sig18'], G['mkfin'], REAL_OUTPUT=True, REAL_OUTPUT_SIZE=(31, 27))
    assert np.max(np.abs(k.cpu().numpy() - ok)) <= 1e-11 * np.max(np.abs(ok))
    # stacking mode, no match kernel
    fs = P.PCDC(64, 96, tS, list(G['sig04']), VERBOSE_LEVEL=0)
    assert np.max(np.abs(fs.cpu().numpy() - do.pcdc(64, 96, S, G['sig04']))) <= 1e-12
    # the convolution that applies the kernel: zero padding, NaN fill, both dtypes
    rng = np.random.default_rng(5)
    img = rng.normal(size=(257, 190)) * 10 + 100
    img[3, 4] = np.nan
    img[200:203, 100] = np.nan
    ker = ok / ok.sum()
    for dt, tol in ((torch.float64, 1e-12), (torch.float32, 2e-6)):
        x = torch.from_numpy(img).to(dev).to(dt)
        out = K.FFT_CONVOLVE(x, torch.from_numpy(ker).to(dev), PAD_FILL_VALUE=0.0, NAN_FILL_VALUE=0.0)
        ref = do.fft_convolve(x.cpu().numpy().astype(float), ker, 0.0, 0.0)
        assert out.dtype == dt
        assert np.max(np.abs(out.cpu().numpy() - ref)) <= tol * np.max(np.abs(ref))
    # NAN_FILL_VALUE=None: the reference's FFT product turns the WHOLE output into NaN as soon as one sample is NaN; the direct
    # evaluation keeps the damage local (the kernel footprint around the sample), which is what the test pins
    clean = np.where(np.isnan(img), 100.0, img)
    out = K.FFT_CONVOLVE(torch.from_numpy(clean).to(dev), torch.from_numpy(G['mkfin']).to(dev), PAD_FILL_VALUE=7.0, NAN_FILL_VALUE=None,
                         NORMALIZE_KERNEL=True).cpu().numpy()
    ref = do.fft_convolve(clean, G['mkfin'], 7.0, None, True)
    assert np.max(np.abs(out - ref)) <= 1e-12 * np.max(np.abs(ref))
    out = K.FFT_CONVOLVE(torch.from_numpy(img).to(dev), torch.from_numpy(G['mkfin']).to(dev), NAN_FILL_VALUE=None).cpu().numpy()
    bad = np.isnan(out)
    assert bad[3, 4] and bad[201, 100] and not bad[3 + 11, 4] and not bad[100, 50] and bad.sum() <= 4 * 21 * 21
    kc = K.KERNEL_CSZ(fin, 64, 48)
    assert np.array_equal(kc.cpu().numpy(), do.csz(G['mkfin'], 64, 48))
    assert np.array_equal(K.KERNEL_CSZ_INV(kc, 21, 21, VERBOSE_LEVEL=0).cpu().numpy(), G['mkfin'])


def _grid_case(N0=150, N1=131, TiHW=12, L=7, seed=11):
    rng = np.random.default_rng(seed)
    img = rng.normal(size=(N0, N1)) * 5 + 50
    img[10, 20] = np.nan
    TiN = 2 * TiHW + 1
    lab, AllocatedL = 0, np.zeros((N0, N1), dtype=int)
    for xs in np.arange(0, N0, TiN):
        for ys in np.arange(0, N1, TiN):
            AllocatedL[xs: min(xs + TiN, N0), ys: min(ys + TiN, N1)] = lab       # the tiling of :4876-4893
            lab += 1
    K = rng.normal(size=(lab, L, L)) + 2.0
    return img, AllocatedL, K


def test_oracle_gsvc_is_a_per_pixel_label_convolution():
    img, AL, K = _grid_case()
    out = do.gsvc(img, AL, K, nan_fill_value=1.5, normalize_kernel=True)
    src = np.where(np.isnan(img), 1.5, img)
    e = np.pad(src, 3)
    Kn = K / K.sum(axis=(1, 2))[:, None, None]
    for (r, c) in [(0, 0), (24, 25), (25, 24), (77, 130), (149, 0), (60, 60)]:
        k = Kn[AL[r, c]]
        want = sum(k[a, b] * e[r + 3 + 3 - a, c + 3 + 3 - b] for a in range(7) for b in range(7))
        assert abs(out[r, c] - want) < 1e-12


@pytest.mark.gpu
def test_grid_convolve_on_device_matches_oracle():
    from sfft_b200.BSplineSFFT import BSpline_GridConvolve
    for norm in (True, False):
        img, AL, K = _grid_case(seed=12 + norm)
        got = BSpline_GridConvolve(img, AL, K, nan_fill_value=0.25, normalize_kernel=norm).GSVC_GPU()
        want = do.gsvc(img, AL, K, nan_fill_value=0.25, normalize_kernel=norm)
        assert np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want))
