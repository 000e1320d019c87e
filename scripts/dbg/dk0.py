import os, sys, itertools
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from util import relrms
from oracle import sfft_oracle as orc
from sfft_b200.plan import Plan
from sfft_b200.synth import make_pair

def run(N0, N1, w, DK, DB, storage, mode, env):
    for k, v in env.items():
        os.environ[k] = v
    d = make_pair(N0, N1, seed=N0 + N1 + w)
    I, J = d['REF'].astype(np.float32), d['SCI'].astype(np.float32)
    I64, J64 = I.astype(np.float64), J.astype(np.float64)
    P = orc.ssc_params(N0, N1, w, DK, DB, True)
    o = orc.gss(I64, J64, I64, J64, P)[1]
    plan = Plan(N0, N1, w, w, DK, DB, True, storage=storage)
    if mode == 'gss32':
        s, dd = plan.gss(I, J, I, J, out_dtype=np.float32)
    elif mode == 'gss64':
        s, dd = plan.gss(I64, J64, I64, J64)
    else:
        s = plan.fit(I64, J64); dd = plan.apply(I64, J64, s)
    print(N0, N1, w, DK, DB, storage, mode, env, 'relrms %.3e' % relrms(dd, o), 'solver', plan.last_solver, flush=True)
    plan.close()
    for k in env:
        os.environ.pop(k)

for shp in [(512, 1024, 2, 0, 1), (2048, 4096, 2, 0, 1)]:
    for storage in ('fp64', 'fp32'):
        for mode in ('gss32', 'gss64', 'sep'):
            run(*shp, storage, mode, {})
    run(*shp, 'fp32', 'gss32', {'SFFTB_NO_HOSTPIPE': '1'})
    run(*shp, 'fp32', 'gss32', {'SFFTB_FIT_NOSEG': '1'})
    run(*shp, 'fp32', 'gss32', {'SFFTB_ROW_NOV8': '1'})
