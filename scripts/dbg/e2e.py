import os, sys, time
import numpy as np
import torch
sys.path.insert(0, '.')
from sfft_b200.batch import PairPipeline, sparse_delta
from sfft_b200.synth import make_pair
N = 4096
d = make_pair(N, N, 20261019)
host = {k: torch.from_numpy(np.ascontiguousarray(v.astype(np.float32))).pin_memory() for k, v in d.items()}
dI = tuple(torch.from_numpy(a).pin_memory().numpy() for a in sparse_delta(host['REF'].numpy(), host['mREF'].numpy()))
dJ = tuple(torch.from_numpy(a).pin_memory().numpy() for a in sparse_delta(host['SCI'].numpy(), host['mSCI'].numpy()))
def run(depth, mode, K=12, env=None):
    for k, v in (env or {}).items(): os.environ[k] = v
    pipe = PairPipeline(N, N, 8, 2, 2, True, device=0, storage='fp32', depth=depth)
    diffs = [torch.empty((N, N), dtype=torch.float32).pin_memory() for _ in range(depth)]
    sols = [torch.empty(pipe.plans[0].NEQ, dtype=torch.float64).pin_memory() for _ in range(depth)]
    def step(k):
        if mode == 'delta': pipe.submit_delta(host['REF'], host['SCI'], dI, dJ, out_dtype=np.float32, Solution_out=sols[k % depth], DIFF_out=diffs[k % depth])
        else: pipe.submit(host['REF'], host['SCI'], host['mREF'], host['mSCI'], out_dtype=np.float32, Solution_out=sols[k % depth], DIFF_out=diffs[k % depth])
    for k in range(depth + 1): step(k)
    pipe.drain(); torch.cuda.synchronize()
    t0 = time.perf_counter(); ts = []
    for k in range(K):
        step(k); ts.append(time.perf_counter() - t0)
    pipe.drain(); torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / K * 1e3
    print('depth', depth, mode, env, 'ms/pair %.3f' % dt, 'submit-return times (ms):', ' '.join('%.1f' % (t * 1e3) for t in ts[:8]), flush=True)
    pipe.close()
    for k in (env or {}): os.environ.pop(k)
run(2, 'full'); run(2, 'delta'); run(3, 'delta'); run(4, 'delta'); run(3, 'full')
run(2, 'delta', env={'SFFTB_OVERLAP': '0'}); run(3, 'delta', env={'SFFTB_OVERLAP': '0'})
