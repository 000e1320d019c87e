# one C2 fit (device-resident) -- used with a library built with -DFS4_DEBUG to print the in-kernel cycle counters
import sys, numpy as np, torch
sys.path.insert(0, '.')
from sfft_b200 import _lib as B
from sfft_b200.plan import Plan
from sfft_b200.synth import make_pair
N = 4096
d = make_pair(N, N, 20261019)
dev = torch.device('cuda', 0)
devt = {k: torch.from_numpy(np.ascontiguousarray(v.astype(np.float32))).to(dev) for k, v in d.items()}
plan = Plan(N, N, 8, 8, 2, 2, True, device=0, storage='fp32')
plan.set_timing(True)
sol = torch.empty(plan.NEQ, dtype=torch.float64, device=dev); diff = torch.empty((N, N), dtype=torch.float32, device=dev)
for _ in range(2):
    plan.gss_device(devt['REF'].data_ptr(), devt['SCI'].data_ptr(), devt['mREF'].data_ptr(), devt['mSCI'].data_ptr(), B.F32, sol.data_ptr(), diff.data_ptr(), B.F32)
    torch.cuda.synchronize()
    print(plan.timings(), flush=True)
