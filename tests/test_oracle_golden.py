"""Pin the CPU oracle (oracle/sfft_oracle.py) against the reference's own golden vector and
against outputs of the unmodified reference NumPy backend (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from goldenio import load_case
from util import relrms, golden_diff
from oracle import sfft_oracle as orc

SYNTH = ['ref_a_64', 'ref_b_96x128', 'ref_c_128x96', 'ref_d_80x60', 'ref_e_256', 'ref_c1_512']


def _run(case, export=None):
    w, DK, DB, CPR = [int(v) for v in case['params']]
    FC = str(case['ForceConv']) if 'ForceConv' in case else 'REF'
    return orc.cp_arrays(case['REF'], case['SCI'], case['mREF'], case['mSCI'], FC, w, DK, DB, bool(CPR),
                         export=export)


def test_oracle_reproduces_reference_golden_ztf():
    """Known-answer test: test/subtract_test_customized/4check/sfft_diff4check.fits
    (KerHW=4, DK=2, DB=2, ConstPhotRatio=True, ForceConv='REF'; subtract4customized.py:18-30)."""
    case = load_case('ztf1024')
    ex = {}
    sol, diff = _run(case, ex)
    gold32 = case['GOLD4CHECK_f32'].astype(np.float64)
    assert np.array_equal(np.isnan(diff), np.isnan(gold32))
    assert int(np.isnan(diff).sum()) == 1801
    assert relrms(diff, gold32) < 1e-7                      # float32 storage floor of the fixture
    assert relrms(diff[::8, ::8], case['GOLD4CHECK_sub8']) < 2e-8   # float64 subsample; reference-run itself: 4e-9
    assert relrms(diff[::8, ::8], case['REFRUN_DIFF_sub8']) < 2e-8
    # known answers recorded in SURVEY.md section 8c
    N = 1024 * 1024
    assert sol.shape == (492,)
    assert np.all(sol[[121, 202, 283, 364, 445]] == 0.0)
    assert abs(sol[40] / N - 0.99509441646) < 1e-6
    np.testing.assert_allclose(sol[-6:], [263.0866851, -5.236421345, 5.540468481, -4.422884005,
                                          1.260148795, 1.222664856], rtol=2e-4)
    # the matrix the reference handed to its solver
    idx = np.setdiff1d(np.arange(492), [121, 202, 283, 364, 445])
    L = ex['LHMAT'][np.ix_(idx, idx)]
    assert np.max(np.abs(L - case['LHMAT_solved'])) <= 1e-9 * np.max(np.abs(case['LHMAT_solved']))
    assert np.max(np.abs(ex['RHb'][idx] - case['RHb_solved'])) <= 1e-9 * np.max(np.abs(case['RHb_solved']))


@pytest.mark.parametrize('name', SYNTH)
def test_oracle_matches_reference_run(name):
    case = load_case(name)
    ex = {}
    sol, diff = _run(case, ex)
    full, sub8 = golden_diff(case)
    assert np.array_equal(np.isnan(diff), np.isnan(full))
    if sub8 is None:
        assert relrms(diff, full) < 1e-8
    else:
        assert relrms(diff, full) < 1e-7
        assert relrms(diff[::8, ::8], sub8) < 1e-8
    w, DK, DB, CPR = [int(v) for v in case['params']]
    P = orc.ssc_params(case['REF'].shape[0], case['REF'].shape[1], w, DK, DB, bool(CPR))
    L, b = ex['LHMAT'], ex['RHb']
    if CPR:
        idx = orc._index_tables(P)[4]
        L, b = L[np.ix_(idx, idx)], b[idx]
    assert np.max(np.abs(L - case['LHMAT_solved'])) <= 1e-9 * np.max(np.abs(case['LHMAT_solved']))
    assert np.max(np.abs(b - case['RHb_solved'])) <= 1e-9 * np.max(np.abs(case['RHb_solved']))


def test_structural_identity_tiny():
    """LHMAT = D^T D / N, RHb = D^T J / N, DIFF = J - D @ Solution."""
    rng = np.random.default_rng(5)
    N0, N1 = 12, 10
    I = rng.normal(10, 3, (N0, N1))
    J = rng.normal(12, 3, (N0, N1))
    for CPR in (True, False):
        P = orc.ssc_params(N0, N1, 1, 1, 1, CPR)
        ex = {}
        sol, _ = orc.ess(I, J, P, None, False, export=ex)
        _, diff = orc.ess(I, J, P, sol, True)
        D = orc.design_matrix(I, P)
        N = N0 * N1
        np.testing.assert_allclose(ex['LHMAT'], D.T @ D / N, rtol=0, atol=1e-12 * np.abs(ex['LHMAT']).max())
        np.testing.assert_allclose(ex['RHb'], D.T @ J.ravel() / N, rtol=0, atol=1e-12 * np.abs(ex['RHb']).max())
        np.testing.assert_allclose(diff.ravel(), J.ravel() - D @ sol, rtol=0, atol=1e-9)
