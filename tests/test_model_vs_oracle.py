"""Pin the NumPy model of the B200 algorithm (tests/model_b200.py) against the CPU oracle."""
import numpy as np
import pytest

import model_b200 as mdl
from oracle import sfft_oracle as orc
from util import relrms


@pytest.mark.parametrize('N0,N1,w,DK,DB,V', [(16, 12, 1, 1, 1, 2), (24, 20, 2, 2, 2, 4), (20, 18, 2, 3, 3, 1),
                                              (32, 16, 3, 0, 0, 8), (16, 16, 2, 2, 1, 4)])
def test_model_normal_equations_and_apply(N0, N1, w, DK, DB, V):
    rng = np.random.default_rng(N0 * 100 + N1)
    I = rng.normal(10, 3, (N0, N1))
    J = rng.normal(12, 3, (N0, N1))
    P = orc.ssc_params(N0, N1, w, DK, DB, True)
    ex = {}
    sol, _ = orc.ess(I, J, P, None, False, export=ex)
    L, b = mdl.fit_normal_eq(I, J, DK, DB, w, w, V)
    assert np.max(np.abs(L - ex['LHMAT'])) <= 1e-11 * np.max(np.abs(ex['LHMAT']))
    assert np.max(np.abs(b - ex['RHb'])) <= 1e-11 * np.max(np.abs(ex['RHb']))
    _, diff = orc.ess(I, J, P, sol, True)
    d2 = mdl.apply_solution(I, J, sol, DK, DB, w, w, V)
    assert relrms(d2, diff) < 1e-10
