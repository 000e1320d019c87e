"""Solution consumers against golden vectors produced by the unmodified
sfft/utils/SFFTSolutionReader.py (tests/golden/make_golden_reader.py): host NumPy routines on the CPU, the CUDA
kernel behind sfftb_realize on the GPU.  Tolerance: 1e-13 relative to max|.| (pure fp64 sums of <= 10 terms)."""
import os
import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
Z = np.load(os.path.join(HERE, 'golden', 'reader_cases.npz'))
NC = int(Z['ncases'])


def _case(n):
    N0, N1, w0, w1, DK, DB = [int(v) for v in Z['c%d_params' % n]]
    kw = dict(Solution=Z['c%d_sol' % n], N0=N0, N1=N1, L0=2 * w0 + 1, L1=2 * w1 + 1, DK=DK,
              Fpq=(DB + 1) * (DB + 2) // 2)
    return (N0, N1, w0, w1, DK, DB), kw


def _close(a, b):
    return np.max(np.abs(np.asarray(a) - b)) <= 1e-13 * max(1e-300, np.max(np.abs(b)))


@pytest.mark.parametrize('n', range(NC))
def test_reader_host_matches_reference(n):
    from sfft_b200.utils import (Read_SFFTSolution, SVKDict_SFFT2ST, SVKDict_ST2SFFT, Realize_MatchingKernel,
                                 Realize_FluxScaling)
    (N0, N1, w0, w1, DK, DB), kw = _case(n)
    XY = Z['c%d_xy' % n]
    XY0 = XY.copy()
    assert _close(Realize_MatchingKernel(XY).FromArray(**kw), Z['c%d_ker' % n])
    assert _close(Realize_FluxScaling(XY).FromArray(**kw), Z['c%d_fs' % n])
    assert np.array_equal(XY, XY0)                       # the request is not modified
    sd = Read_SFFTSolution().FromArray(**kw)
    keys = sorted(sd)
    assert keys == sorted((i, j) for i in range(DK + 1) for j in range(DK + 1 - i))
    assert _close(np.stack([sd[k] for k in keys]), Z['c%d_sfft' % n])
    st = SVKDict_SFFT2ST.convert(DK, DK, sd)
    assert _close(np.stack([st[k] for k in keys]), Z['c%d_std' % n])
    back = SVKDict_ST2SFFT.convert(DK, DK, st)
    assert _close(np.stack([back[k] for k in keys]), Z['c%d_back' % n])
    # the kernel sum equals the flux scaling at every coordinate (:170-172)
    ker = Realize_MatchingKernel(XY).FromArray(**kw)
    fs = Realize_FluxScaling(XY).FromArray(**kw)
    assert np.max(np.abs(ker.sum(axis=(1, 2)) - fs)) <= 1e-10 * np.max(np.abs(ker)) * ker[0].size


def test_reader_fits_roundtrip(tmp_path):
    from sfft_b200 import fitsio
    from sfft_b200.utils import Realize_MatchingKernel, Realize_FluxScaling, Read_SFFTSolution
    (N0, N1, w0, w1, DK, DB), kw = _case(1)
    Fij, Fab = (DK + 1) * (DK + 2) // 2, (2 * w0 + 1) * (2 * w1 + 1)
    ups = [(k, v, 'MeLOn: SFFT') for k, v in (('N0', N0), ('N1', N1), ('DK', DK), ('DB', DB), ('L0', kw['L0']),
           ('L1', kw['L1']), ('FIJ', Fij), ('FAB', Fab), ('FPQ', kw['Fpq']), ('FIJAB', Fij * Fab))]
    path = str(tmp_path / 'solution.fits')
    fitsio.writeto(path, kw['Solution'].reshape((-1, 1)).T, base_cards=None, updates=ups)   # what CP writes (:205-221)
    XY = Z['c1_xy']
    assert _close(Realize_MatchingKernel(XY).FromFITS(path), Z['c1_ker'])
    assert _close(Realize_FluxScaling(XY).FromFITS(path), Z['c1_fs'])
    sd = Read_SFFTSolution().FromFITS(path)
    assert _close(np.stack([sd[k] for k in sorted(sd)]), Z['c1_sfft'])


@pytest.mark.gpu
@pytest.mark.parametrize('n', range(NC))
def test_reader_device_matches_reference(n):
    import torch
    from sfft_b200 import SingleSFFTConfigure
    from sfft_b200.plan import Plan
    from sfft_b200.utils import Realize_MatchingKernel, Realize_FluxScaling
    (N0, N1, w0, w1, DK, DB), kw = _case(n)
    if w0 == w1:
        cfg = SingleSFFTConfigure.SSC(NX=N0, NY=N1, KerHW=w0, KerPolyOrder=DK, BGPolyOrder=DB, ConstPhotRatio=True,
                                      VERBOSE_LEVEL=0)
    else:       # sfftcore's SSC takes one KerHW (SFFTConfigure.py:1371); the native plan takes both half widths
        cfg = ({}, {'plan': Plan(N0, N1, w0, w1, DK, DB, True)})
    sol = torch.from_numpy(kw['Solution']).cuda()
    XY = Z['c%d_xy' % n]
    ker = Realize_MatchingKernel(XY).FromDevice(sol, cfg)
    fs = Realize_FluxScaling(XY).FromDevice(sol, cfg)
    assert ker.is_cuda and fs.is_cuda
    assert _close(ker.cpu().numpy(), Z['c%d_ker' % n])
    assert _close(fs.cpu().numpy(), Z['c%d_fs' % n])
    # host pointers through the same C entry point
    from sfft_b200 import _lib as B
    plan = cfg[1]['plan']
    kh = np.empty(Z['c%d_ker' % n].shape)
    fh = np.empty(len(XY))
    xy = np.ascontiguousarray(XY, np.float64)
    s = np.ascontiguousarray(kw['Solution'])
    B.check(B.lib().sfftb_realize(plan._h, s.ctypes.data, B.MEM_HOST, xy.ctypes.data, B.MEM_HOST, len(XY),
                                  kh.ctypes.data, fh.ctypes.data, B.MEM_HOST))
    assert _close(kh, Z['c%d_ker' % n]) and _close(fh, Z['c%d_fs' % n])
