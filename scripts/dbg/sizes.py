# stage timings of one GSS at instrument-like image sizes (row lengths that are not powers of two)
import sys, numpy as np, torch
sys.path.insert(0, '.')
from sfft_b200 import _lib as B
from sfft_b200.plan import Plan
from sfft_b200.synth import make_pair
dev = torch.device('cuda', 0)
sizes = [(4096, 4096), (3080, 3072), (3072, 3072), (6144, 6144), (2560, 2560), (1536, 1536)]
if len(sys.argv) > 1:
    sizes = [tuple(int(x) for x in a.split('x')) for a in sys.argv[1:]]
for (N0, N1) in sizes:
    d = make_pair(N0, N1, 7)
    devt = {k: torch.from_numpy(np.ascontiguousarray(v.astype(np.float32))).to(dev) for k, v in d.items()}
    plan = Plan(N0, N1, 8, 8, 2, 2, True, device=0, storage='fp32')
    plan.set_timing(True)
    sol = torch.empty(plan.NEQ, dtype=torch.float64, device=dev); diff = torch.empty((N0, N1), dtype=torch.float32, device=dev)
    for _ in range(3):
        plan.gss_device(devt['REF'].data_ptr(), devt['SCI'].data_ptr(), devt['mREF'].data_ptr(), devt['mSCI'].data_ptr(), B.F32, sol.data_ptr(), diff.data_ptr(), B.F32)
        torch.cuda.synchronize()
    t = plan.timings()
    print((N0, N1), 'total %.2f ms' % sum(v for k, v in t.items() if k != 'fit_cols_kernel'), {k: round(v, 2) for k, v in t.items()}, flush=True)
    plan.close()
