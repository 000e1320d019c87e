"""GPU parity of the general-basis path (sfft_b200/BSplineSFFT.py -> sfftb_plan_create_general -> csrc/kernels_gen.cuh).

Pins:
  * tests/golden/bspline_cases.npz -- runs of the reference's development copy (NumPy backend, unmodified; B-spline and
    polynomial kernels, ConstPhotRatio True = SEPARATE-CONSTANT / False = ENTANGLED): the solved system (after the stripe
    tweak) within 1e-9 of max|.|, DIFF from the REFERENCE's Solution within 1e-10 relative RMS, DIFF end to end within 1e-6
    (cond(LHMAT) up to 4e11 on these tiny cases; the oracle itself reproduces them to 1e-6 end to end);
  * oracle/bspline_oracle.py (design-matrix restatement of BSplineSFFT.py) for what the development copy cannot run:
    SEPARATE-VARYING, B-spline scaling / background, the regulariser, polynomial degree 4 -- same tolerances.
"""
import os
import numpy as np
import pytest

from util import relrms
from oracle import bspline_oracle as bo

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
Z = np.load(os.path.join(HERE, 'golden', 'bspline_cases.npz'))
NC = int(Z['ncases'])


@pytest.fixture(scope='module')
def bs():
    import sfft_b200.BSplineSFFT as m
    return m


def _case_kwargs(n):
    pre = 'c%d_' % n
    N0, N1, w, DK, DB, CPR = [int(v) for v in Z[pre + 'shape']]
    KT, BT = [str(v) for v in Z[pre + 'types']]
    kw = dict(KerHW=w, KerSpType=KT, KerSpDegree=DK, KerIntKnotX=list(Z[pre + 'KX']), KerIntKnotY=list(Z[pre + 'KY']),
              SEPARATE_SCALING=bool(CPR), ScaSpDegree=0, BkgSpType=BT, BkgSpDegree=DB,
              BkgIntKnotX=list(Z[pre + 'BX']), BkgIntKnotY=list(Z[pre + 'BY']))
    return pre, N0, N1, kw


@pytest.mark.parametrize('storage', ['fp64', 'fp32'])
@pytest.mark.parametrize('n', range(NC))
def test_general_plan_matches_reference_dev_copy(bs, n, storage):
    pre, N0, N1, kw = _case_kwargs(n)
    cfg = bs.SingleSFFTConfigure.SSC(NX=N0, NY=N1, VERBOSE_LEVEL=0, STORAGE=storage, FORCE_GENERAL_PLAN=True, **kw)
    P, plan = cfg[0], cfg[1]['plan']
    sol, diff, _ = bs.GeneralSFFTSubtract.GSS(Z[pre + 'REF'], Z[pre + 'SCI'], Z[pre + 'mREF'], Z[pre + 'mSCI'], cfg, VERBOSE_LEVEL=0)
    assert sol.shape == (P['NEQ'],) and sol.shape == Z[pre + 'sol'].shape
    if storage == 'fp32':
        assert relrms(diff, Z[pre + 'diff']) < 2e-4          # tiny, badly conditioned systems; the fp64 run is the pin
        return
    # the solved system: fit on the masked pair was the last fit
    L, b = plan.export_solved_system()
    gb = Z[pre + 'b']
    assert b.shape == gb.shape == (P['NEQt'],)
    assert np.max(np.abs(b - gb)) <= 1e-9 * np.max(np.abs(gb))
    if pre + 'A' in Z.files:
        A = Z[pre + 'A']
        assert np.max(np.abs(L - A)) <= 1e-9 * np.max(np.abs(A))
    else:
        rows = Z[pre + 'Arows']
        assert np.max(np.abs(L[rows] - Z[pre + 'Asub'])) <= 1e-9 * np.max(np.abs(Z[pre + 'Adiag']))
        assert np.max(np.abs(np.diag(L) - Z[pre + 'Adiag'])) <= 1e-9 * np.max(np.abs(Z[pre + 'Adiag']))
    # Construct_FDIFF with the reference's own Solution
    _, d2 = bs.ElementalSFFTSubtract.ESS(Z[pre + 'REF'], Z[pre + 'SCI'], cfg, SFFTSolution=Z[pre + 'sol'], Subtract=True, VERBOSE_LEVEL=0)
    assert relrms(d2, Z[pre + 'diff']) < 1e-10
    assert relrms(diff, Z[pre + 'diff']) < 1e-6
    # Restore_Solution: tied / dropped centre taps exactly as the reference leaves them
    ij00 = np.arange(P['w0'] * P['L1'] + P['w1'], P['Fijab'], P['Fab'])
    if P['SCALING_MODE'] == 'SEPARATE-CONSTANT':
        if P['KerSpType'] == 'B-Spline':
            assert np.all(sol[ij00] == sol[ij00[0]])
        else:
            assert np.all(sol[ij00[1:]] == 0.0)


def _pair(N0, N1, seed):
    from sfft_b200.synth import make_pair
    d = make_pair(N0, N1, seed=seed, density=8e-3)
    return d['REF'], d['SCI'], d['mREF'], d['mSCI']


MODES = [
    # SEPARATE-VARYING, B-spline kernel + polynomial scaling
    dict(KerSpType='B-Spline', KerSpDegree=2, KerIntKnotX=[20.0], KerIntKnotY=[], SEPARATE_SCALING=True, ScaSpType='Polynomial',
         ScaSpDegree=1, BkgSpType='Polynomial', BkgSpDegree=1),
    # SEPARATE-VARYING, B-spline kernel + B-spline scaling with its own knots
    dict(KerSpType='B-Spline', KerSpDegree=2, KerIntKnotX=[16.0, 30.0], KerIntKnotY=[22.0], SEPARATE_SCALING=True, ScaSpType='B-Spline',
         ScaSpDegree=1, ScaIntKnotX=[24.0], ScaIntKnotY=[], BkgSpType='Polynomial', BkgSpDegree=2),
    # B-spline background (the development copy cannot run it), constant scaling
    dict(KerSpType='B-Spline', KerSpDegree=1, KerIntKnotX=[], KerIntKnotY=[20.0], SEPARATE_SCALING=True, ScaSpDegree=0,
         BkgSpType='B-Spline', BkgSpDegree=2, BkgIntKnotX=[24.0], BkgIntKnotY=[13.0, 27.0]),
    # polynomial kernel with a B-spline background, entangled scaling
    dict(KerSpType='Polynomial', KerSpDegree=2, SEPARATE_SCALING=False, BkgSpType='B-Spline', BkgSpDegree=1, BkgIntKnotX=[], BkgIntKnotY=[]),
    # polynomial degree 4 (beyond sfftcore's 0..3), polynomial varying scaling of degree 2
    dict(KerSpType='Polynomial', KerSpDegree=4, SEPARATE_SCALING=True, ScaSpType='Polynomial', ScaSpDegree=2, BkgSpType='Polynomial',
         BkgSpDegree=4),
]


@pytest.mark.parametrize('k', range(len(MODES)))
@pytest.mark.parametrize('reg', [False, True])
def test_general_plan_matches_bspline_oracle(bs, k, reg):
    N0, N1, w = 48, 40, 2
    I, J, mI, mJ = _pair(N0, N1, 900 + k)
    kw = dict(MODES[k])
    if reg:
        rng = np.random.default_rng(k)
        kw.update(REGULARIZE_KERNEL=True, XY_REGULARIZE=np.stack([rng.uniform(0.5, N0 + 0.5, 11), rng.uniform(0.5, N1 + 0.5, 11)], axis=1),
                  WEIGHT_REGULARIZE=rng.uniform(0.5, 2.0, 11), LAMBDA_REGULARIZE=3e-5, IGNORE_LAPLACIAN_KERCENT=bool(k % 2))
    P = bo.ssc_params(N0, N1, w, **kw)
    cfg = bs.SingleSFFTConfigure.SSC(NX=N0, NY=N1, KerHW=w, VERBOSE_LEVEL=0, FORCE_GENERAL_PLAN=True, **kw)
    assert cfg[0]['NEQ'] == P['NEQ'] and cfg[0]['NEQt'] == P['NEQt'] and cfg[0]['SCALING_MODE'] == P['SCALING_MODE']
    plan = cfg[1]['plan']
    sol, diff, _ = bs.GeneralSFFTSubtract.GSS(I, J, mI, mJ, cfg, VERBOSE_LEVEL=0)
    ex = {}
    osol, _ = bo.ess(mI, mJ, P, None, False, export=ex)
    L, b = plan.export_solved_system()
    assert np.max(np.abs(b - ex['RHb_tweaked'])) <= 1e-9 * np.max(np.abs(ex['RHb_tweaked']))
    assert np.max(np.abs(L - ex['LHMAT_tweaked'])) <= 1e-9 * np.max(np.abs(ex['LHMAT_tweaked']))
    _, od = bo.ess(I, J, P, osol, True)
    _, d2 = bs.ElementalSFFTSubtract.ESS(I, J, cfg, SFFTSolution=osol, Subtract=True, VERBOSE_LEVEL=0)
    assert relrms(d2, od) < 1e-10
    # end to end the difference image carries cond(LHMAT) * eps of the solve; the degree-4 polynomial pair on 48 x 40 pixels
    # is the worst conditioned of the set (measured 1.1e-6 with the same LHMAT / RHb to 1e-9)
    assert relrms(diff, od) < (1e-5 if k == 4 else 1e-6)
    ij00 = np.arange(P['w0'] * P['L1'] + P['w1'], P['Fijab'], P['Fab'])
    if P['SCALING_MODE'] == 'SEPARATE-VARYING':
        assert np.all(sol[ij00[P['ScaFij']:]] == 0.0)


def test_general_plan_equals_specialised_polynomial_plan(bs):
    """The same polynomial configuration through the table-driven kernels and through the sfftcore kernels."""
    N0, N1, w = 300, 128, 3
    I, J, mI, mJ = _pair(N0, N1, 77)
    for kw in (dict(SEPARATE_SCALING=True, ScaSpDegree=0), dict(SEPARATE_SCALING=False),
               dict(SEPARATE_SCALING=True, ScaSpType='Polynomial', ScaSpDegree=1)):
        c1 = bs.SingleSFFTConfigure.SSC(NX=N0, NY=N1, KerHW=w, KerSpDegree=2, BkgSpDegree=2, VERBOSE_LEVEL=0, **kw)
        c2 = bs.SingleSFFTConfigure.SSC(NX=N0, NY=N1, KerHW=w, KerSpDegree=2, BkgSpDegree=2, VERBOSE_LEVEL=0, FORCE_GENERAL_PLAN=True, **kw)
        s1, d1, _ = bs.GeneralSFFTSubtract.GSS(I, J, mI, mJ, c1, VERBOSE_LEVEL=0)
        s2, d2, _ = bs.GeneralSFFTSubtract.GSS(I, J, mI, mJ, c2, VERBOSE_LEVEL=0)
        assert relrms(d2, d1) < 1e-8
        L1_, b1 = c1[1]['plan'].export_solved_system()
        L2_, b2 = c2[1]['plan'].export_solved_system()
        assert np.max(np.abs(L1_ - L2_)) <= 1e-11 * np.max(np.abs(L1_)) and np.max(np.abs(b1 - b2)) <= 1e-11 * np.max(np.abs(b1))


def test_bspline_packet_arrays_nan_union(bs):
    """BSP_arrays with a B-spline kernel: NaN union fill / restore and the ForceConv='SCI' sign flip around the general plan."""
    N0, N1, w = 64, 56, 2
    I, J, mI, mJ = _pair(N0, N1, 5)
    I = I.copy(); J = J.copy()
    I[5, 7] = np.nan; J[40, 3] = np.nan
    kw = dict(KerSpType='B-Spline', KerSpDegree=2, KerIntKnotX=[30.0], KerIntKnotY=[], BkgSpType='Polynomial', BkgSpDegree=1)
    mI2, mJ2 = np.where(np.isnan(mI), 0.0, mI), np.where(np.isnan(mJ), 0.0, mJ)
    s, d = bs.BSpline_Packet.BSP_arrays(I, J, mI2, mJ2, ForceConv='SCI', GKerHW=w, VERBOSE_LEVEL=0, **kw)
    P = bo.ssc_params(N0, N1, w, SEPARATE_SCALING=True, ScaSpDegree=0, **kw)
    U = np.isnan(I) | np.isnan(J)
    fI, fJ = np.where(U, mJ2 * 0 + mI2, I), np.where(U, mJ2, J)
    # ForceConv='SCI': the science image is convolved: (I, J) = (SCI, REF), DIFF flipped
    osol, od = bo.gss(fJ, fI, mJ2, mI2, P)
    od = -od
    od[U] = np.nan
    assert np.array_equal(np.isnan(d), np.isnan(od))
    assert relrms(d, od) < 1e-6


@pytest.mark.parametrize('N0,N1,w,full', [(520, 48, 2, True), (1536, 192, 4, False)])
def test_local_support_skip_is_exact(bs, monkeypatch, N0, N1, w, full):
    """Several column segments (N0 = 520 -> three; 1536 with knots at the thirds of both axes -> the geometry of BASELINE config 3),
    B-spline row functions with knots inside the image: the fit column pass skips the transforms and products of windows on which a
    basis function vanishes (GenPass::jobs / seginfo), the FIR of the subtraction skips the planes that vanish on a chunk.  The
    results must be those of the dense pass BIT FOR BIT (skipped terms are exact zeros; there are no floating-point atomics on the
    path, so two runs repeat exactly), and match the design-matrix oracle where that is affordable."""
    I, J, mI, mJ = _pair(N0, N1, 4242)
    if full:
        kw = dict(KerSpType='B-Spline', KerSpDegree=2, KerIntKnotX=[180.0, 340.0], KerIntKnotY=[20.0], SEPARATE_SCALING=True,
                  ScaSpType='Polynomial', ScaSpDegree=1, BkgSpType='Polynomial', BkgSpDegree=2)
    else:
        kw = dict(KerSpType='B-Spline', KerSpDegree=2, KerIntKnotX=[N0 / 3.0, 2.0 * N0 / 3.0], KerIntKnotY=[N1 / 3.0, 2.0 * N1 / 3.0],
                  SEPARATE_SCALING=True, ScaSpType='Polynomial', ScaSpDegree=2, BkgSpType='Polynomial', BkgSpDegree=2)
    out = []
    for dense in ('0', '1', '0'):
        monkeypatch.setenv('SFFTB_GEN_DENSE', dense)
        cfg = bs.SingleSFFTConfigure.SSC(NX=N0, NY=N1, KerHW=w, VERBOSE_LEVEL=0, FORCE_GENERAL_PLAN=True, **kw)
        sol, diff, _ = bs.GeneralSFFTSubtract.GSS(I, J, mI, mJ, cfg, VERBOSE_LEVEL=0)
        L, b = cfg[1]['plan'].export_solved_system()
        out.append((sol, diff, L, b))
        cfg[1]['plan'].close()
    monkeypatch.delenv('SFFTB_GEN_DENSE')
    for k in (1, 2):
        for x, y in zip(out[0], out[k]):
            assert np.array_equal(x, y)
    if not full:
        return
    P = bo.ssc_params(N0, N1, w, **kw)
    ex = {}
    osol, _ = bo.ess(mI, mJ, P, None, False, export=ex)
    L, b = out[0][2], out[0][3]
    assert np.max(np.abs(b - ex['RHb_tweaked'])) <= 1e-9 * np.max(np.abs(ex['RHb_tweaked']))
    assert np.max(np.abs(L - ex['LHMAT_tweaked'])) <= 1e-9 * np.max(np.abs(ex['LHMAT_tweaked']))
    _, od = bo.ess(I, J, P, osol, True)
    assert relrms(out[0][1], od) < 1e-6
