import os, sys
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from util import relrms
from oracle import bspline_oracle as bo
import sfft_b200.BSplineSFFT as bs
from sfft_b200.synth import make_pair

def run(N0, N1, w, kw, tag):
    d = make_pair(N0, N1, seed=900, density=8e-3)
    I, J = d['REF'], d['SCI']
    P = bo.ssc_params(N0, N1, w, **kw)
    cfg = bs.SingleSFFTConfigure.SSC(NX=N0, NY=N1, KerHW=w, VERBOSE_LEVEL=0, FORCE_GENERAL_PLAN=True, **kw)
    plan = cfg[1]['plan']
    sol, diff = bs.ElementalSFFTSubtract.ESS(I, J, cfg, Subtract=True, VERBOSE_LEVEL=0)
    ex = {}
    osol, od = bo.ess(I, J, P, None, True, export=ex)
    L, b = plan.export_solved_system()
    Lo, bo_ = ex['LHMAT_tweaked'], ex['RHb_tweaked']
    nk = P['NEQt'] - P['Fpq']
    def e(a, b2): return np.max(np.abs(a - b2)) / max(np.max(np.abs(b2)), 1e-300)
    print(tag, 'n', P['NEQt'], 'b_ker %.2e b_bkg %.2e | L_kk %.2e L_kb %.2e L_bb %.2e | diff(osol) %.2e diff %.2e' % (
        e(b[:nk], bo_[:nk]), e(b[nk:], bo_[nk:]), e(L[:nk, :nk], Lo[:nk, :nk]), e(L[:nk, nk:], Lo[:nk, nk:]), e(L[nk:, nk:], Lo[nk:, nk:]),
        relrms(bs.ElementalSFFTSubtract.ESS(I, J, cfg, SFFTSolution=osol, Subtract=True, VERBOSE_LEVEL=0)[1], od), relrms(diff, od)), flush=True)
    if e(b[:nk], bo_[:nk]) > 1e-6:
        r = (b[:nk] / bo_[:nk])
        print('   ratio b_ker first 12:', np.round(r[:12], 4))
        Fab = P['Fab']
        print('   per-plane max err:', [float('%.2e' % e(b[A*Fab:(A+1)*Fab if (A+1)*Fab <= nk else nk], bo_[A*Fab:(A+1)*Fab if (A+1)*Fab <= nk else nk])) for A in range(min(P['Fij'], nk // Fab))])

poly1 = dict(KerSpType='Polynomial', KerSpDegree=1, BkgSpType='Polynomial', BkgSpDegree=1)
run(48, 40, 2, dict(SEPARATE_SCALING=False, **dict(poly1, KerSpDegree=0)), 'poly0 ent ')
run(48, 40, 2, dict(SEPARATE_SCALING=False, **poly1), 'poly1 ent ')
run(48, 40, 2, dict(SEPARATE_SCALING=True, ScaSpDegree=0, **poly1), 'poly1 drop')
run(300, 64, 2, dict(SEPARATE_SCALING=False, **poly1), 'poly1 ent 300x64')
run(600, 64, 3, dict(SEPARATE_SCALING=False, **poly1), 'poly1 ent 600x64')
bsp = dict(KerSpType='B-Spline', KerSpDegree=2, KerIntKnotX=[], KerIntKnotY=[], BkgSpType='Polynomial', BkgSpDegree=1)
run(48, 40, 2, dict(SEPARATE_SCALING=False, **bsp), 'bsp2 ent  ')
run(48, 40, 2, dict(SEPARATE_SCALING=True, ScaSpDegree=0, **bsp), 'bsp2 sum  ')
run(48, 40, 2, dict(SEPARATE_SCALING=True, ScaSpType='Polynomial', ScaSpDegree=1, **bsp), 'bsp2 vary ')
