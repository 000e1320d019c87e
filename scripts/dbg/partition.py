# Device-resident 4096^2 pairs through `depth` plans, each on its own stream, with / without the SM partition.
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, '.')
from sfft_b200 import _lib as B
from sfft_b200.batch import PairPipeline
from sfft_b200.synth import make_pair
N = int(os.environ.get('PN', 4096))
d = make_pair(N, N, 20261019)
dev = torch.device('cuda', 0)
devt = {k: torch.from_numpy(np.ascontiguousarray(v.astype(np.float32))).to(dev) for k, v in d.items()}
ref = {}
def run(depth, sms, K=24, env=None, timing=False):
    for k, v in (env or {}).items(): os.environ[k] = v
    pipe = PairPipeline(N, N, 8, 2, 2, True, device=0, storage='fp32', depth=depth, solver_sms=sms)
    for pl in pipe.plans: pl.set_timing(timing)
    diffs = [torch.empty((N, N), dtype=torch.float32, device=dev) for _ in range(depth)]
    sols = [torch.empty(pipe.plans[0].NEQ, dtype=torch.float64, device=dev) for _ in range(depth)]
    def step(k):
        pipe.submit_device(devt['REF'].data_ptr(), devt['SCI'].data_ptr(), devt['mREF'].data_ptr(), devt['mSCI'].data_ptr(), B.F32,
                           sols[k % depth].data_ptr(), diffs[k % depth].data_ptr(), B.F32)
    for k in range(depth + 2): step(k)
    pipe.drain(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(K): step(k)
    pipe.drain(); torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / K * 1e3
    s = sols[0].cpu().numpy(); df = diffs[0].cpu().numpy()
    if 's' not in ref: ref['s'], ref['d'] = s, df
    print('depth %d solver_sms %3d %s ms/pair %.3f  (%.0f Mpix/s)  sol maxdiff %.2e diff maxdiff %.2e solver %s' % (
        depth, sms, env or '', dt, N * N / 1e3 / dt, np.abs(s - ref['s']).max(), np.abs(df - ref['d']).max(), pipe.plans[0].last_solver), flush=True)
    if timing: print('   stage ms', {k: round(v, 3) for k, v in pipe.plans[0].timings().items()})
    pipe.close()
    for k in (env or {}): os.environ.pop(k)
import sys as _s
if len(_s.argv) > 1 and _s.argv[1] == 'sweep2':
    for sms in (11, 18, 19, 20, 21, 22, 24, 28, 32):
        run(3, sms)
    run(4, 20); run(2, 20)
elif len(_s.argv) > 1 and _s.argv[1] == 'timing':
    run(2, 0); run(3, 16); run(3, 16, timing=True); run(2, 16, timing=True); run(1, 0, timing=True)
    print(pipe_t if 0 else '')
else:
    run(2, 0)
    for depth in (2, 3):
        for sms in (8, 12, 16, 24, 32, 48):
            run(depth, sms)
