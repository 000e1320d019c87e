"""
Golden vectors for the B-spline variant: run the reference's own development copy with a NumPy backend
(misc/beta4spline/new_version_sfftcore, UNMODIFIED, loaded by path; pyfftw shimmed with scipy.fft as in
make_golden.py) on small seeded pairs.  It covers B-spline / polynomial kernels and backgrounds with
ConstPhotRatio=True (== SEPARATE-CONSTANT of sfft/BSplineSFFT.py) and False (== ENTANGLED).  The released
BSplineSFFT.py has no CPU backend (:2605-2607), so this is the only executable pin in the build container.

    python tests/golden/make_golden_bspline.py   ->  tests/golden/bspline_cases.npz
"""
import os
import sys
import importlib.util
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
REFROOT = os.environ.get('SFFT_REFERENCE', '/root/reference')

from make_golden import load_reference            # noqa: E402  (installs the pyfftw shim)
from sfft_b200.synth import make_pair              # noqa: E402

CASES = [
    # N0, N1, KerHW, KerSpType, DK, KnotX, KnotY, BkgSpType, DB, BKnotX, BKnotY, ConstPhotRatio
    # (B-spline backgrounds crash in the development copy's NumPy backend -- 'BkgSpatial' is never bound,
    #  new_version_sfftcore/SFFTConfigure.py:1266 -- so only polynomial backgrounds can be pinned this way)
    (48, 40, 2, 'B-Spline', 2, [], [], 'Polynomial', 1, [], [], True),
    (48, 40, 2, 'B-Spline', 2, [20.0], [], 'Polynomial', 1, [], [], False),
    (40, 48, 1, 'B-Spline', 1, [14.0, 27.0], [25.0], 'Polynomial', 2, [], [], True),
    (36, 36, 2, 'Polynomial', 2, [], [], 'Polynomial', 2, [], [], True),
    (36, 36, 2, 'Polynomial', 1, [], [], 'Polynomial', 1, [], [], False),
    (64, 56, 3, 'B-Spline', 3, [30.0], [28.0], 'Polynomial', 0, [], [], True),
]


def load_beta():
    load_reference()       # pyfftw shim into sys.modules

    def load(name, path):
        s = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(s)
        s.loader.exec_module(m)
        return m
    base = os.path.join(REFROOT, 'misc/beta4spline/new_version_sfftcore')
    return load('beta_cfg', os.path.join(base, 'SFFTConfigure.py')), load('beta_sub', os.path.join(base, 'SFFTSubtract.py'))


if __name__ == '__main__':
    C, S = load_beta()
    out = {'ncases': np.array(len(CASES))}
    for n, (N0, N1, w, KT, DK, KX, KY, BT, DB, BX, BY, CPR) in enumerate(CASES):
        d = make_pair(N0, N1, seed=700 + n, density=8e-3)
        cfg = C.SingleSFFTConfigure.SSC(NX=N0, NY=N1, KerHW=w, KerSpType=KT, KerSpDegree=DK, KerIntKnotX=KX,
                                        KerIntKnotY=KY, BkgSpType=BT, BkgSpDegree=DB, BkgIntKnotX=BX, BkgIntKnotY=BY,
                                        ConstPhotRatio=CPR, BACKEND_4SUBTRACT='Numpy', NUM_CPU_THREADS_4SUBTRACT=4,
                                        VERBOSE_LEVEL=0)
        cap = {}
        _solve = np.linalg.solve

        def spy(A, b):
            cap['A'], cap['b'] = A.copy(), b.copy()
            return _solve(A, b)
        np.linalg.solve = spy
        try:
            sol, diff, _ = S.GeneralSFFTSubtract.GSS(PixA_I=d['REF'], PixA_J=d['SCI'], PixA_mI=d['mREF'], PixA_mJ=d['mSCI'],
                                                     SFFTConfig=cfg, ContamMask_I=None, BACKEND_4SUBTRACT='Numpy',
                                                     NUM_CPU_THREADS_4SUBTRACT=4, VERBOSE_LEVEL=0)
        finally:
            np.linalg.solve = _solve
        pre = 'c%d_' % n
        out.update({pre + 'shape': np.array([N0, N1, w, DK, DB, int(CPR)]), pre + 'types': np.array([KT, BT]),
                    pre + 'KX': np.array(KX, float), pre + 'KY': np.array(KY, float), pre + 'BX': np.array(BX, float),
                    pre + 'BY': np.array(BY, float), pre + 'REF': d['REF'], pre + 'SCI': d['SCI'], pre + 'mREF': d['mREF'],
                    pre + 'mSCI': d['mSCI'], pre + 'sol': sol, pre + 'diff': diff, pre + 'b': cap['b']})
        if cap['A'].shape[0] <= 320:
            out[pre + 'A'] = cap['A']
        else:       # large system: keep the diagonal and a seeded sample of rows (fixture size)
            rows = np.sort(np.random.default_rng(n).choice(cap['A'].shape[0], 48, replace=False))
            out.update({pre + 'Arows': rows, pre + 'Asub': cap['A'][rows], pre + 'Adiag': np.diag(cap['A']).copy()})
        print(n, KT, BT, CPR, 'NEQ', cfg[0]['NEQ'], 'solved', cap['A'].shape, 'cond %.2e' % np.linalg.cond(cap['A']))
    np.savez_compressed(os.path.join(HERE, 'bspline_cases.npz'), **out)
    print('wrote bspline_cases.npz')
