"""BSplineSFFT Solution consumers (Read_SFFTSolution, BSpline_MatchingKernel; sfft/BSplineSFFT.py:4358-4723) against the
unmodified reference (tests/golden/bsreader_cases.npz, tests/golden/make_golden_bsreaders.py), and the FITS round trip through
the Solution header BSpline_Packet.BSP writes."""
import os
import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, 'golden', 'bsreader_cases.npz'))
KX, KY = [100.0, 200.0], [120.0]
CASES = {
    'bs_ent': dict(KerSpType='B-Spline', DK=2, KerIntKnotX=KX, KerIntKnotY=KY, SEPARATE_SCALING=False, ScaSpType=None, DS=None, ScaIntKnotX=None, ScaIntKnotY=None),
    'bs_const': dict(KerSpType='B-Spline', DK=2, KerIntKnotX=KX, KerIntKnotY=KY, SEPARATE_SCALING=True, ScaSpType='Polynomial', DS=0, ScaIntKnotX=[], ScaIntKnotY=[]),
    'bs_varpoly': dict(KerSpType='B-Spline', DK=2, KerIntKnotX=KX, KerIntKnotY=KY, SEPARATE_SCALING=True, ScaSpType='Polynomial', DS=1, ScaIntKnotX=[], ScaIntKnotY=[]),
    'bs_varbs': dict(KerSpType='B-Spline', DK=2, KerIntKnotX=KX, KerIntKnotY=KY, SEPARATE_SCALING=True, ScaSpType='B-Spline', DS=1, ScaIntKnotX=[150.0], ScaIntKnotY=[]),
    'poly_ent': dict(KerSpType='Polynomial', DK=2, KerIntKnotX=[], KerIntKnotY=[], SEPARATE_SCALING=False, ScaSpType=None, DS=None, ScaIntKnotX=None, ScaIntKnotY=None),
    'poly_varpoly': dict(KerSpType='Polynomial', DK=3, KerIntKnotX=[], KerIntKnotY=[], SEPARATE_SCALING=True, ScaSpType='Polynomial', DS=2, ScaIntKnotX=[], ScaIntKnotY=[]),
}


def _dims(name):
    N0, N1, L0, L1 = [int(v) for v in G['dims']]
    Fi, Fj, Fpq, ScaFi, ScaFj = [int(v) for v in G[name + '_F']]
    return N0, N1, L0, L1, Fi, Fj, Fpq, (None if ScaFi == -9 else ScaFi), (None if ScaFj == -9 else ScaFj)


@pytest.mark.parametrize('name', sorted(CASES))
def test_readers_match_reference(name):
    from sfft_b200.BSplineSFFT import Read_SFFTSolution, BSpline_MatchingKernel
    c = CASES[name]
    N0, N1, L0, L1, Fi, Fj, Fpq, ScaFi, ScaFj = _dims(name)
    sol = G[name + '_sol']
    kd, sd = Read_SFFTSolution().FromArray(Solution=sol, KerSpType=c['KerSpType'], N0=N0, N1=N1, DK=c['DK'], L0=L0, L1=L1, Fi=Fi, Fj=Fj,
                                           Fpq=Fpq, SEPARATE_SCALING=c['SEPARATE_SCALING'], ScaSpType=c['ScaSpType'], DS=c['DS'],
                                           ScaFi=ScaFi, ScaFj=ScaFj)
    for key, ref in zip(G[name + '_kerkeys'], G[name + '_kerdict']):
        np.testing.assert_array_equal(kd[tuple(int(v) for v in key)], ref)           # NaN centre taps included
    if name + '_scakeys' in G.files:
        ref = {tuple(int(v) for v in k): float(x) for k, x in zip(G[name + '_scakeys'], G[name + '_scadict'])}
        for k, v in sd.items():
            assert ref[k] == v
    else:
        assert sd is None
    ks = BSpline_MatchingKernel(XY_q=G['XY'], VERBOSE_LEVEL=0).FromArray(
        Solution=sol, KerSpType=c['KerSpType'], KerIntKnotX=c['KerIntKnotX'], KerIntKnotY=c['KerIntKnotY'], N0=N0, N1=N1, DK=c['DK'], L0=L0,
        L1=L1, Fi=Fi, Fj=Fj, Fpq=Fpq, SEPARATE_SCALING=c['SEPARATE_SCALING'], ScaSpType=c['ScaSpType'], ScaIntKnotX=c['ScaIntKnotX'],
        ScaIntKnotY=c['ScaIntKnotY'], DS=c['DS'], ScaFi=ScaFi, ScaFj=ScaFj)
    ref = G[name + '_kerstack']
    assert np.array_equal(np.isnan(ks), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert np.max(np.abs(ks[ok] - ref[ok])) <= 1e-13 * np.max(np.abs(ref[ok]))


def test_solution_fits_header_round_trip(tmp_path):
    """The header keys BSP writes next to the Solution are the ones FromFITS reads (:4277-4354, :4525-4553, :4664-4723)."""
    from sfft_b200 import fitsio
    from sfft_b200.BSplineSFFT import BSpline_MatchingKernel, Read_SFFTSolution
    name = 'bs_varbs'
    N0, N1, L0, L1, Fi, Fj, Fpq, ScaFi, ScaFj = _dims(name)
    sol = G[name + '_sol']
    C = 'SFFT'
    ups = [('KERHW', (L0 - 1) // 2, C), ('KSPTYPE', 'B-Spline', C), ('NKIKX', 2, C), ('KIKX0', 100.0, C), ('KIKX1', 200.0, C), ('NKIKY', 1, C),
           ('KIKY0', 120.0, C), ('SEPSCA', 'True', C), ('SSPTYPE', 'B-Spline', C), ('SSPDEG', 1, C), ('NSIKX', 1, C), ('SIKX0', 150.0, C),
           ('NSIKY', 0, C), ('N0', N0, C), ('N1', N1, C), ('DK', 2, C), ('L0', L0, C), ('L1', L1, C), ('FI', Fi, C), ('FJ', Fj, C), ('FPQ', Fpq, C),
           ('SCAFI', ScaFi, C), ('SCAFJ', ScaFj, C)]
    p = str(tmp_path / 'sol.fits')
    fitsio.writeto(p, sol.reshape((-1, 1)).T, base_cards=None, updates=ups)
    ks = BSpline_MatchingKernel(XY_q=G['XY'], VERBOSE_LEVEL=0).FromFITS(p)
    ref = G[name + '_kerstack']
    ok = ~np.isnan(ref)
    assert np.max(np.abs(ks[ok] - ref[ok])) <= 1e-13 * np.max(np.abs(ref[ok]))
    kd, sd = Read_SFFTSolution().FromFITS(p)
    assert sd is not None and len(kd) == Fi * Fj
