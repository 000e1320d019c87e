"""oracle/bspline_oracle.py (design-matrix restatement of sfft/BSplineSFFT.py) against the golden vectors produced by
the reference's development copy with a NumPy backend (tests/golden/make_golden_bspline.py), plus the internal
consistency of the modes that have no executable reference here (SEPARATE-VARYING, regulariser: parity UNPINNED).
Tolerances: LHMAT / RHb 1e-12 of max|.|; DIFF with the reference's own Solution 1e-11 relative RMS; DIFF end to end
1e-6 (cond(LHMAT) up to 4e11 on these cases)."""
import os
import numpy as np
import pytest

from oracle import bspline_oracle as bo
from oracle import sfft_oracle as orc
from util import relrms

HERE = os.path.dirname(os.path.abspath(__file__))
Z = np.load(os.path.join(HERE, 'golden', 'bspline_cases.npz'))
NC = int(Z['ncases'])


def _params(n):
    pre = 'c%d_' % n
    N0, N1, w, DK, DB, CPR = [int(v) for v in Z[pre + 'shape']]
    KT, BT = [str(v) for v in Z[pre + 'types']]
    P = bo.ssc_params(N0, N1, w, KerSpType=KT, KerSpDegree=DK, KerIntKnotX=list(Z[pre + 'KX']), KerIntKnotY=list(Z[pre + 'KY']),
                      SEPARATE_SCALING=bool(CPR), ScaSpDegree=0, BkgSpType=BT, BkgSpDegree=DB,
                      BkgIntKnotX=list(Z[pre + 'BX']), BkgIntKnotY=list(Z[pre + 'BY']))
    return pre, P


@pytest.mark.parametrize('n', range(NC))
def test_bspline_oracle_matches_reference_numpy_backend(n):
    pre, P = _params(n)
    ex = {}
    sol, _ = bo.ess(Z[pre + 'mREF'], Z[pre + 'mSCI'], P, None, False, export=ex)
    b = Z[pre + 'b']
    assert ex['RHb_tweaked'].shape == b.shape == (P['NEQt'],)
    assert np.max(np.abs(ex['RHb_tweaked'] - b)) <= 1e-12 * np.max(np.abs(b))
    L = ex['LHMAT_tweaked']
    if pre + 'A' in Z.files:
        A = Z[pre + 'A']
        assert np.max(np.abs(L - A)) <= 1e-12 * np.max(np.abs(A))
    else:
        rows = Z[pre + 'Arows']
        assert np.max(np.abs(L[rows] - Z[pre + 'Asub'])) <= 1e-12 * np.max(np.abs(Z[pre + 'Adiag']))
        assert np.max(np.abs(np.diag(L) - Z[pre + 'Adiag'])) <= 1e-12 * np.max(np.abs(Z[pre + 'Adiag']))
    _, d_ref_sol = bo.ess(Z[pre + 'REF'], Z[pre + 'SCI'], P, Z[pre + 'sol'], True)
    assert relrms(d_ref_sol, Z[pre + 'diff']) < 1e-11
    _, d = bo.ess(Z[pre + 'REF'], Z[pre + 'SCI'], P, sol, True)
    assert relrms(d, Z[pre + 'diff']) < 1e-6
    # the restored Solution ties the stripes exactly as Restore_Solution does (BSplineSFFT.py:3764-3771)
    ij00 = np.arange(P['w0'] * P['L1'] + P['w1'], P['Fijab'], P['Fab'])
    if P['SCALING_MODE'] == 'SEPARATE-CONSTANT':
        if P['KerSpType'] == 'B-Spline':
            assert np.all(sol[ij00] == sol[ij00[0]]) and np.all(Z[pre + 'sol'][ij00] == Z[pre + 'sol'][ij00[0]])
        else:
            assert np.all(sol[ij00[1:]] == 0.0) and np.all(Z[pre + 'sol'][ij00[1:]] == 0.0)


def test_polynomial_case_equals_sfftcore_oracle():
    """Polynomial kernel + background, SEPARATE-CONSTANT == sfftcore with ConstPhotRatio=True (BSplineSFFT is a superset)."""
    rng = np.random.default_rng(11)
    N0, N1 = 20, 24
    I, J = rng.normal(10, 3, (N0, N1)), rng.normal(12, 3, (N0, N1))
    for CPR in (True, False):
        P = bo.ssc_params(N0, N1, 2, KerSpType='Polynomial', KerSpDegree=2, SEPARATE_SCALING=CPR, ScaSpDegree=0,
                          BkgSpType='Polynomial', BkgSpDegree=1)
        Pc = orc.ssc_params(N0, N1, 2, 2, 1, CPR)
        s1, d1 = bo.gss(I, J, I, J, P)
        s2, d2 = orc.gss(I, J, I, J, Pc)[:2]
        assert relrms(d1, d2) < 1e-9
        assert P['NEQ'] == Pc['NEQ'] and P['NEQt'] == (Pc['NEQ_FSfree'] if CPR else Pc['NEQ'])


def test_separate_varying_and_regularizer_consistency():
    """No executable reference for these modes in this container (parity unpinned); the structure is checked:
    SEPARATE-VARYING with a degree-0 polynomial scaling basis spans the same model as SEPARATE-CONSTANT (B-spline
    partition of unity), the regulariser is symmetric positive semi-definite, vanishes on a delta kernel when the
    centre rows are ignored, and lambda -> 0 recovers the unregularised solution."""
    rng = np.random.default_rng(5)
    N0, N1, w = 24, 20, 1
    I, J = rng.normal(10, 3, (N0, N1)), rng.normal(12, 3, (N0, N1))
    kw = dict(KerSpType='B-Spline', KerSpDegree=1, KerIntKnotX=[12.0], KerIntKnotY=[], BkgSpType='Polynomial', BkgSpDegree=1)
    Pc = bo.ssc_params(N0, N1, w, SEPARATE_SCALING=True, ScaSpDegree=0, **kw)
    dc = bo.gss(I, J, I, J, Pc)[1]
    Pv = bo.ssc_params(N0, N1, w, SEPARATE_SCALING=True, ScaSpType='Polynomial', ScaSpDegree=1, **kw)
    assert Pv['SCALING_MODE'] == 'SEPARATE-VARYING' and Pv['ScaFij'] == 3 and Pv['NEQt'] == Pv['NEQ'] - Pv['Fij'] + 3
    sv, dv = bo.gss(I, J, I, J, Pv)
    ij00 = np.arange(w * Pv['L1'] + w, Pv['Fijab'], Pv['Fab'])
    assert np.all(sv[ij00[3:]] == 0.0)
    assert np.sum(dv ** 2) <= np.sum(dc ** 2) * (1 + 1e-9)          # richer scaling model fits at least as well
    XY = np.stack([rng.uniform(0.5, N0 + 0.5, 9), rng.uniform(0.5, N1 + 0.5, 9)], axis=1)
    for mode_kw in (dict(SEPARATE_SCALING=False), dict(SEPARATE_SCALING=True, ScaSpDegree=0),
                    dict(SEPARATE_SCALING=True, ScaSpType='Polynomial', ScaSpDegree=1)):
        Pr = bo.ssc_params(N0, N1, w, REGULARIZE_KERNEL=True, XY_REGULARIZE=XY, LAMBDA_REGULARIZE=1e-20, **mode_kw, **kw)
        R = bo.regularizer(Pr)
        assert np.max(np.abs(R - R.T)) <= 1e-12 * np.max(np.abs(R))
        assert np.min(np.linalg.eigvalsh(R)) >= -1e-9 * np.max(np.abs(R))
        assert np.all(R[Pr['Fijab']:, :] == 0.0)
        P0 = dict(Pr, REGULARIZE_KERNEL=False)
        assert relrms(bo.gss(I, J, I, J, Pr)[1], bo.gss(I, J, I, J, P0)[1]) < 1e-8
        if mode_kw == dict(SEPARATE_SCALING=False):
            delta = np.zeros(Pr['NEQ'])
            delta[ij00] = 1.0                                          # a pure delta kernel at every control point
            assert abs(delta @ R @ delta) <= 1e-12 * np.max(np.abs(R))
