"""
Golden fixtures of the BSplineSFFT Solution consumers: Read_SFFTSolution.FromArray and BSpline_MatchingKernel.FromArray
(sfft/BSplineSFFT.py:4358-4662) of the UNMODIFIED reference.  The module itself imports astropy (absent here), so the two class
definitions are executed straight from the reference's source text (nothing is copied into this repository) with NumPy and
scipy.interpolate.BSpline in scope.

    python tests/golden/make_golden_bsreaders.py   ->  tests/golden/bsreader_cases.npz
"""
import os
import numpy as np
from scipy.interpolate import BSpline

HERE = os.path.dirname(os.path.abspath(__file__))
REFROOT = os.environ.get('SFFT_REFERENCE', '/root/reference')


def main():
    src = open(os.path.join(REFROOT, 'sfft/BSplineSFFT.py')).read()
    a, b = src.index('class Read_SFFTSolution:'), src.index('class ConvKernel_Convertion:')
    ns = {'np': np, 'BSpline': BSpline, 'fits': None}
    exec(compile(src[a:b], 'ref:BSplineSFFT.py', 'exec'), ns)
    rng = np.random.default_rng(20261017)
    N0, N1, w = 300, 240, 2
    L0 = L1 = 2 * w + 1
    kx, ky = [100.0, 200.0], [120.0]
    cases = {
        'bs_ent': dict(KerSpType='B-Spline', DK=2, KerIntKnotX=kx, KerIntKnotY=ky, SEPARATE_SCALING=False, ScaSpType=None, DS=None, ScaIntKnotX=None, ScaIntKnotY=None),
        'bs_const': dict(KerSpType='B-Spline', DK=2, KerIntKnotX=kx, KerIntKnotY=ky, SEPARATE_SCALING=True, ScaSpType='Polynomial', DS=0, ScaIntKnotX=[], ScaIntKnotY=[]),
        'bs_varpoly': dict(KerSpType='B-Spline', DK=2, KerIntKnotX=kx, KerIntKnotY=ky, SEPARATE_SCALING=True, ScaSpType='Polynomial', DS=1, ScaIntKnotX=[], ScaIntKnotY=[]),
        'bs_varbs': dict(KerSpType='B-Spline', DK=2, KerIntKnotX=kx, KerIntKnotY=ky, SEPARATE_SCALING=True, ScaSpType='B-Spline', DS=1, ScaIntKnotX=[150.0], ScaIntKnotY=[]),
        'poly_ent': dict(KerSpType='Polynomial', DK=2, KerIntKnotX=[], KerIntKnotY=[], SEPARATE_SCALING=False, ScaSpType=None, DS=None, ScaIntKnotX=None, ScaIntKnotY=None),
        'poly_varpoly': dict(KerSpType='Polynomial', DK=3, KerIntKnotX=[], KerIntKnotY=[], SEPARATE_SCALING=True, ScaSpType='Polynomial', DS=2, ScaIntKnotX=[], ScaIntKnotY=[]),
    }
    XY = np.array([[1.0, 1.0], [150.5, 120.25], [300.0, 240.0], [37.0, 201.5], [299.2, 3.3]])
    out = {'XY': XY, 'dims': np.array([N0, N1, L0, L1])}
    for name, c in cases.items():
        DK = c['DK']
        if c['KerSpType'] == 'B-Spline':
            Fi, Fj = len(c['KerIntKnotX']) + DK + 1, len(c['KerIntKnotY']) + DK + 1
            Fij = Fi * Fj
        else:
            Fi = Fj = -1
            Fij = (DK + 1) * (DK + 2) // 2
        ScaFi = ScaFj = None
        if c['SEPARATE_SCALING'] and c['DS'] > 0:
            ScaFi = ScaFj = -1                     # placeholder of SingleSFFTConfigure for a polynomial scaling (:175)
            if c['ScaSpType'] == 'B-Spline':
                ScaFi, ScaFj = len(c['ScaIntKnotX']) + c['DS'] + 1, len(c['ScaIntKnotY']) + c['DS'] + 1
        Fpq = 6
        sol = rng.normal(size=Fij * L0 * L1 + Fpq) * N0 * N1
        kd, sd = ns['Read_SFFTSolution']().FromArray(Solution=sol, KerSpType=c['KerSpType'], N0=N0, N1=N1, DK=DK, L0=L0, L1=L1, Fi=Fi, Fj=Fj,
                                                     Fpq=Fpq, SEPARATE_SCALING=c['SEPARATE_SCALING'], ScaSpType=c['ScaSpType'], DS=c['DS'],
                                                     ScaFi=ScaFi, ScaFj=ScaFj)
        ks = ns['BSpline_MatchingKernel'](XY_q=XY, VERBOSE_LEVEL=0).FromArray(
            Solution=sol, KerSpType=c['KerSpType'], KerIntKnotX=c['KerIntKnotX'], KerIntKnotY=c['KerIntKnotY'], N0=N0, N1=N1, DK=DK, L0=L0,
            L1=L1, Fi=Fi, Fj=Fj, Fpq=Fpq, SEPARATE_SCALING=c['SEPARATE_SCALING'], ScaSpType=c['ScaSpType'], ScaIntKnotX=c['ScaIntKnotX'],
            ScaIntKnotY=c['ScaIntKnotY'], DS=c['DS'], ScaFi=ScaFi, ScaFj=ScaFj)
        out[name + '_sol'] = sol
        out[name + '_kerstack'] = ks
        keys = sorted(kd)
        out[name + '_kerkeys'] = np.array(keys)
        out[name + '_kerdict'] = np.array([kd[k] for k in keys])
        if sd is not None:
            skeys = sorted(sd)
            out[name + '_scakeys'] = np.array(skeys)
            out[name + '_scadict'] = np.array([sd[k] for k in skeys])
        out[name + '_F'] = np.array([Fi, Fj, Fpq, -9 if ScaFi is None else ScaFi, -9 if ScaFj is None else ScaFj])
    np.savez_compressed(os.path.join(HERE, 'bsreader_cases.npz'), **out)
    print('wrote bsreader_cases.npz:', sorted(k for k in out if k.endswith('_kerstack')))


if __name__ == '__main__':
    main()
