"""sfft_b200/BSplineSFFT.py: the BSplineSFFT call signatures on the native plan (polynomial bases, ENTANGLED /
SEPARATE-CONSTANT scaling, kernel regularisation).  CPU part: argument behaviour and the regulariser factors against
the oracle's dense REGMAT; GPU part: GSS against oracle/bspline_oracle.py (design-matrix restatement of
sfft/BSplineSFFT.py; the regulariser has no executable reference in the build container -- parity unpinned, both
sides restate :3570-3700 independently).  Tolerances: LHMAT 1e-10 of max|.|, DIFF 1e-8 relative RMS (fp64 storage)."""
import numpy as np
import pytest

from oracle import bspline_oracle as bo
from util import relrms


def test_argument_checks_mirror_the_reference():
    """Argument validation that happens before any plan is built (BSplineSFFT.py:33-38, 79-81, 190)."""
    from sfft_b200.BSplineSFFT import SingleSFFTConfigure
    with pytest.raises(AssertionError):          # ScaFij <= Fij
        SingleSFFTConfigure.SSC(64, 64, KerHW=2, KerSpDegree=1, SEPARATE_SCALING=True, ScaSpDegree=2, VERBOSE_LEVEL=0)
    with pytest.raises(Exception, match='not available in sfft_b200'):
        SingleSFFTConfigure.SSC(64, 64, KerHW=2, BACKEND_4SUBTRACT='Numpy', VERBOSE_LEVEL=0)
    with pytest.raises(AssertionError):
        SingleSFFTConfigure.SSC(64, 64, KerHW=2, KerSpType='Fourier', VERBOSE_LEVEL=0)
    with pytest.raises(AssertionError):          # a degree-0 B-spline has no internal knots
        SingleSFFTConfigure.SSC(64, 64, KerHW=2, KerSpType='B-Spline', KerSpDegree=0, KerIntKnotX=[30.0], VERBOSE_LEVEL=0)
    with pytest.raises(AssertionError):
        SingleSFFTConfigure.SSC(64, 64, KerHW=2, BkgSpType='B-Spline', BkgSpDegree=0, BkgIntKnotY=[30.0], VERBOSE_LEVEL=0)


def test_basis_tables_match_oracle_basis_planes():
    """The 1-D tables handed to the CUDA library reproduce the oracle's 2-D basis planes (Create_BSplineBasis :2624-2634,
    REF_ij / REF_pq order :2764-2772), B-spline partition of unity included."""
    from sfft_b200.BSplineSFFT import _basis_tables
    N0, N1 = 37, 52
    for SpType, deg, KX, KY in (('B-Spline', 2, [12.0, 25.0], [30.0]), ('B-Spline', 3, [], [20.0, 21.5]), ('B-Spline', 0, [], []),
                                ('Polynomial', 3, [], []), ('Polynomial', 5, [], [])):
        U, V, fu, fv = _basis_tables(SpType, deg, KX, KY, N0, N1)
        planes = np.array([np.outer(U[i], V[j]) for i, j in zip(fu, fv)])
        ref = bo._basis_planes(SpType, deg, KX, KY, N0, N1)
        assert planes.shape == ref.shape and np.array_equal(planes, ref)
        if SpType == 'B-Spline':
            assert np.max(np.abs(planes.sum(axis=0) - 1.0)) < 1e-14
            assert len(fu) == U.shape[0] * V.shape[0]


@pytest.mark.parametrize('mode', ['ENTANGLED', 'SEPARATE-CONSTANT', 'SEPARATE-VARYING'])
def test_gram_factors_match_oracle_regmat_bspline(mode):
    """_gram_factors (the Kronecker factors handed to sfftb_set_regularizer[_varying]) against the oracle's dense REGMAT for a
    B-spline kernel and, in SEPARATE-VARYING, a B-spline scaling basis (fill_regmat :2091-2166)."""
    from sfft_b200.BSplineSFFT import _gram_factors
    rng = np.random.default_rng(8)
    N0, N1, w = 40, 36, 1
    XY = np.stack([rng.uniform(0.5, N0 + 0.5, 9), rng.uniform(0.5, N1 + 0.5, 9)], axis=1)
    W = rng.uniform(0.5, 2.0, 9)
    kw = dict(KerSpType='B-Spline', KerSpDegree=2, KerIntKnotX=[15.0], KerIntKnotY=[], BkgSpType='Polynomial', BkgSpDegree=1)
    if mode == 'ENTANGLED':
        kw.update(SEPARATE_SCALING=False)
    elif mode == 'SEPARATE-CONSTANT':
        kw.update(SEPARATE_SCALING=True, ScaSpDegree=0)
    else:
        kw.update(SEPARATE_SCALING=True, ScaSpType='B-Spline', ScaSpDegree=1, ScaIntKnotX=[], ScaIntKnotY=[18.0])
    P = bo.ssc_params(N0, N1, w, REGULARIZE_KERNEL=True, XY_REGULARIZE=XY, WEIGHT_REGULARIZE=W, IGNORE_LAPLACIAN_KERCENT=True, **kw)
    assert P['SCALING_MODE'] == mode
    R = bo.regularizer(P)
    fac = _gram_factors(P, N0, N1, w, w, XY, W, True)
    SST, iREG = fac[0], fac[1]
    Fab, Fij, c0 = P['Fab'], P['Fij'], w * P['L1'] + w
    K = np.kron(SST, iREG)
    if mode == 'SEPARATE-VARYING':
        CSST, DSST = fac[2], fac[3]
        cen = np.zeros(Fab, bool)
        cen[c0] = True
        one = np.ones((Fij, Fij))
        K = np.where(np.kron(one, np.outer(~cen, cen)).astype(bool), np.kron(CSST, iREG), K)
        K = np.where(np.kron(one, np.outer(cen, ~cen)).astype(bool), np.kron(CSST.T, iREG), K)
        K = np.where(np.kron(one, np.outer(cen, cen)).astype(bool), np.kron(DSST, iREG), K)
    full = np.zeros_like(R)
    full[:P['Fijab'], :P['Fijab']] = P['SCALE'] ** 2 * K
    assert np.max(np.abs(full - R)) <= 1e-13 * np.max(np.abs(R))


@pytest.mark.parametrize('ignore', [True, False])
@pytest.mark.parametrize('weighted', [False, True])
def test_regularizer_factors_match_oracle_regmat(ignore, weighted):
    from sfft_b200.BSplineSFFT import regularizer_factors
    rng = np.random.default_rng(3)
    N0, N1, w, DK = 40, 36, 2, 2
    XY = np.stack([rng.uniform(0.5, N0 + 0.5, 11), rng.uniform(0.5, N1 + 0.5, 11)], axis=1)
    W = rng.uniform(0.5, 2.0, 11) if weighted else None
    P = bo.ssc_params(N0, N1, w, KerSpType='Polynomial', KerSpDegree=DK, SEPARATE_SCALING=False, BkgSpType='Polynomial',
                      BkgSpDegree=1, REGULARIZE_KERNEL=True, XY_REGULARIZE=XY, WEIGHT_REGULARIZE=W,
                      IGNORE_LAPLACIAN_KERCENT=ignore)
    R = bo.regularizer(P)
    SST, iREG = regularizer_factors(N0, N1, w, w, DK, XY, W, ignore)
    K = np.zeros_like(R)
    K[:P['Fijab'], :P['Fijab']] = P['SCALE'] ** 2 * np.kron(SST, iREG)
    assert np.max(np.abs(K - R)) <= 1e-14 * np.max(np.abs(R))


@pytest.mark.gpu
@pytest.mark.parametrize('separate', [False, True])
@pytest.mark.parametrize('regularize', [False, True])
def test_bsplinesfft_polynomial_modes_match_oracle(separate, regularize):
    from sfft_b200.BSplineSFFT import SingleSFFTConfigure, GeneralSFFTSubtract
    from sfft_b200.synth import make_pair
    N0, N1, w, DK, DB = 96, 80, 2, 2, 1
    d = make_pair(N0, N1, seed=77, density=8e-3)
    rng = np.random.default_rng(9)
    XY = np.stack([rng.uniform(0.5, N0 + 0.5, 25), rng.uniform(0.5, N1 + 0.5, 25)], axis=1)
    kw = dict(KerSpType='Polynomial', KerSpDegree=DK, SEPARATE_SCALING=separate, ScaSpDegree=0, BkgSpType='Polynomial',
              BkgSpDegree=DB)
    lam = 1e-6
    if regularize:
        # a penalty that matters: a tenth of the largest kernel-block entry of LHMAT
        P0 = bo.ssc_params(N0, N1, w, REGULARIZE_KERNEL=True, XY_REGULARIZE=XY, LAMBDA_REGULARIZE=1.0, **kw)
        ex0 = {}
        bo.ess(d['mREF'], d['mSCI'], dict(P0, REGULARIZE_KERNEL=False), None, False, export=ex0)
        lam = 0.1 * np.max(np.abs(ex0['LHMAT'][:P0['Fijab'], :P0['Fijab']])) / np.max(np.abs(bo.regularizer(P0)))
    rkw = dict(REGULARIZE_KERNEL=regularize, XY_REGULARIZE=XY if regularize else None, LAMBDA_REGULARIZE=lam)
    cfg = SingleSFFTConfigure.SSC(N0, N1, KerHW=w, VERBOSE_LEVEL=0, **kw, **rkw)
    P = bo.ssc_params(N0, N1, w, **kw, **rkw)
    for k in ('N0', 'N1', 'w0', 'w1', 'DK', 'DB', 'L0', 'L1', 'Fab', 'Fi', 'Fj', 'Fij', 'Fp', 'Fq', 'Fpq', 'Fijab', 'NEQ', 'NEQt'):
        assert cfg[0][k] == P[k], k
    sol, diff, _ = GeneralSFFTSubtract.GSS(d['REF'], d['SCI'], d['mREF'], d['mSCI'], cfg, VERBOSE_LEVEL=0)
    ex = {}
    osol, _ = bo.ess(d['mREF'], d['mSCI'], P, None, False, export=ex)
    _, odiff = bo.ess(d['REF'], d['SCI'], P, osol, True)
    L, b = cfg[1]['plan'].export_normal_eq()
    assert np.max(np.abs(L - ex['LHMAT'])) <= 1e-10 * np.max(np.abs(ex['LHMAT']))
    assert np.max(np.abs(b - ex['RHb'])) <= 1e-10 * np.max(np.abs(ex['RHb']))
    assert relrms(diff, odiff) < 1e-8
    if regularize:      # the penalty changed the answer (the test would otherwise not see a missing regulariser)
        _, d_plain = bo.gss(d['REF'], d['SCI'], d['mREF'], d['mSCI'], dict(P, REGULARIZE_KERNEL=False))
        assert relrms(diff, d_plain) > 1e-6
        cfg[1]['plan'].set_regularizer(None)
        _, diff_off, _ = GeneralSFFTSubtract.GSS(d['REF'], d['SCI'], d['mREF'], d['mSCI'], cfg, VERBOSE_LEVEL=0)
        assert relrms(diff_off, d_plain) < 1e-8


@pytest.mark.gpu
def test_bsp_packet_fits_roundtrip(tmp_path):
    from sfft_b200 import fitsio
    from sfft_b200.BSplineSFFT import BSpline_Packet
    from sfft_b200.synth import make_pair
    N0, N1 = 64, 72
    d = make_pair(N0, N1, seed=5, density=8e-3)
    d['SCI'][3, 5] = np.nan
    d['mSCI'][3, 5] = 0.0
    paths = {}
    for k in ('REF', 'SCI', 'mREF', 'mSCI'):
        paths[k] = str(tmp_path / (k + '.fits'))
        fitsio.writeto(paths[k], d[k].T)
    out, solp = str(tmp_path / 'diff.fits'), str(tmp_path / 'sol.fits')
    sol, diff = BSpline_Packet.BSP(paths['REF'], paths['SCI'], paths['mREF'], paths['mSCI'], FITS_DIFF=out, FITS_Solution=solp,
                                   ForceConv='SCI', GKerHW=2, KerSpDegree=1, BkgSpDegree=1, VERBOSE_LEVEL=0)
    assert np.isnan(diff[3, 5]) and np.isfinite(np.delete(diff.ravel(), 3 * N1 + 5)).all()
    P = bo.ssc_params(N0, N1, 2, KerSpDegree=1, BkgSpDegree=1, SEPARATE_SCALING=True, ScaSpDegree=0)
    I, J = d['SCI'].copy(), d['REF'].copy()              # ForceConv='SCI': the science image is convolved
    I[3, 5], J[3, 5] = d['mSCI'][3, 5], d['mREF'][3, 5]
    odiff = -bo.gss(I, J, d['mSCI'], d['mREF'], P)[1]
    ok = np.isfinite(diff)
    assert relrms(diff[ok], odiff[ok]) < 1e-8
    assert relrms(fitsio.getdata(out).T[ok], odiff[ok]) < 1e-8
    assert np.array_equal(np.asarray(fitsio.getdata(solp), np.float64)[0], sol)


@pytest.mark.gpu
@pytest.mark.parametrize('DK,DS,DB,regularize', [(2, 1, 1, False), (2, 2, 2, False), (3, 1, 0, False), (1, 1, 1, False),
                                                  (2, 1, 1, True), (2, 2, 0, True)])
def test_separate_varying_polynomial_scaling_matches_oracle(DK, DS, DB, regularize):
    """SEPARATE-VARYING with polynomial bases (BSplineSFFT.py:2487-2495, 3733-3747; no executable reference here -- the
    oracle restates it in design-matrix form, parity unpinned): normal equations, solution layout and DIFF."""
    from sfft_b200.BSplineSFFT import SingleSFFTConfigure, GeneralSFFTSubtract
    from sfft_b200.synth import make_pair
    N0, N1, w = 88, 96, 2
    d = make_pair(N0, N1, seed=31 + DK + DS, density=8e-3)
    kw = dict(KerSpType='Polynomial', KerSpDegree=DK, SEPARATE_SCALING=True, ScaSpType='Polynomial', ScaSpDegree=DS,
              BkgSpType='Polynomial', BkgSpDegree=DB)
    if regularize:
        rng = np.random.default_rng(DK + 7 * DS)
        XY = np.stack([rng.uniform(0.5, N0 + 0.5, 30), rng.uniform(0.5, N1 + 0.5, 30)], axis=1)
        P0 = bo.ssc_params(N0, N1, w, REGULARIZE_KERNEL=True, XY_REGULARIZE=XY, LAMBDA_REGULARIZE=1.0, **kw)
        ex0 = {}
        bo.ess(d['mREF'], d['mSCI'], dict(P0, REGULARIZE_KERNEL=False), None, False, export=ex0)
        lam = 0.1 * np.max(np.abs(ex0['LHMAT'][:P0['Fijab'], :P0['Fijab']])) / np.max(np.abs(bo.regularizer(P0)))
        kw.update(REGULARIZE_KERNEL=True, XY_REGULARIZE=XY, WEIGHT_REGULARIZE=rng.uniform(0.5, 2.0, 30), LAMBDA_REGULARIZE=lam)
    cfg = SingleSFFTConfigure.SSC(N0, N1, KerHW=w, VERBOSE_LEVEL=0, **kw)
    P = bo.ssc_params(N0, N1, w, **kw)
    assert P['SCALING_MODE'] == 'SEPARATE-VARYING'
    for k in ('Fij', 'Fpq', 'Fijab', 'NEQ', 'NEQt', 'ScaFij', 'DS'):
        assert cfg[0][k] == P[k], k
    assert cfg[1]['plan'].dims['NEQ_FSfree'] == P['NEQt']
    sol, diff, _ = GeneralSFFTSubtract.GSS(d['REF'], d['SCI'], d['mREF'], d['mSCI'], cfg, VERBOSE_LEVEL=0)
    ex = {}
    osol, _ = bo.ess(d['mREF'], d['mSCI'], P, None, False, export=ex)
    _, odiff = bo.ess(d['REF'], d['SCI'], P, osol, True)
    L, b = cfg[1]['plan'].export_normal_eq()
    assert np.max(np.abs(L - ex['LHMAT'])) <= 1e-10 * np.max(np.abs(ex['LHMAT']))
    assert np.max(np.abs(b - ex['RHb'])) <= 1e-10 * np.max(np.abs(ex['RHb']))
    ij00 = np.arange(w * P['L1'] + w, P['Fijab'], P['Fab'])
    assert np.all(sol[ij00[P['ScaFij']:]] == 0.0)
    assert relrms(diff, odiff) < 1e-8
    # apply with the oracle's own solution isolates the apply step (solution remap + FIR)
    d2 = cfg[1]['plan'].apply(d['REF'], d['SCI'], osol)
    assert relrms(d2, odiff) < 1e-10
