"""
Golden vectors for the Solution consumers: run the UNMODIFIED sfft/utils/SFFTSolutionReader.py (loaded by path,
astropy.io.fits shimmed -- only FromArray is exercised) on seeded random Solutions.

    python tests/golden/make_golden_reader.py      ->  tests/golden/reader_cases.npz
"""
import os
import sys
import types
import importlib.util
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFROOT = os.environ.get('SFFT_REFERENCE', '/root/reference')


def load_reader():
    ap, io, fits = types.ModuleType('astropy'), types.ModuleType('astropy.io'), types.ModuleType('astropy.io.fits')
    ap.io, io.fits = io, fits
    sys.modules.update({'astropy': ap, 'astropy.io': io, 'astropy.io.fits': fits})
    s = importlib.util.spec_from_file_location('ref_reader', os.path.join(REFROOT, 'sfft/utils/SFFTSolutionReader.py'))
    m = importlib.util.module_from_spec(s)
    s.loader.exec_module(m)
    return m


CASES = [(64, 80, 2, 3, 0, 1), (512, 512, 4, 4, 2, 2), (4096, 4096, 8, 8, 2, 2), (300, 200, 5, 3, 3, 3), (128, 96, 1, 1, 1, 0)]

if __name__ == '__main__':
    R = load_reader()
    out = {}
    for n, (N0, N1, w0, w1, DK, DB) in enumerate(CASES):
        rng = np.random.default_rng(100 + n)
        L0, L1 = 2 * w0 + 1, 2 * w1 + 1
        Fij, Fpq = (DK + 1) * (DK + 2) // 2, (DB + 1) * (DB + 2) // 2
        sol = rng.normal(0, 1, Fij * L0 * L1 + Fpq) * N0 * N1 * 1e-2
        XY = np.stack([rng.uniform(0.5, N0 + 0.5, 7), rng.uniform(0.5, N1 + 0.5, 7)], axis=1)
        kw = dict(Solution=sol, N0=N0, N1=N1, L0=L0, L1=L1, DK=DK, Fpq=Fpq)
        ker = R.Realize_MatchingKernel(XY.copy()).FromArray(**kw)
        fs = R.Realize_FluxScaling(XY.copy()).FromArray(**kw)
        sd = R.Read_SFFTSolution().FromArray(**kw)
        st = R.SVKDict_SFFT2ST.convert(DK, DK, sd)
        back = R.SVKDict_ST2SFFT.convert(DK, DK, st)
        keys = sorted(sd)
        out.update({'c%d_params' % n: np.array([N0, N1, w0, w1, DK, DB]), 'c%d_sol' % n: sol, 'c%d_xy' % n: XY,
                    'c%d_ker' % n: ker, 'c%d_fs' % n: fs, 'c%d_sfft' % n: np.stack([sd[k] for k in keys]),
                    'c%d_std' % n: np.stack([st[k] for k in keys]), 'c%d_back' % n: np.stack([back[k] for k in keys])})
    out['ncases'] = np.array(len(CASES))
    np.savez_compressed(os.path.join(HERE, 'reader_cases.npz'), **out)
    print('wrote reader_cases.npz with', len(CASES), 'cases')
