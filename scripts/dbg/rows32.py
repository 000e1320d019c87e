# one fit on a 1184 x 16384 fp64 pair (8 rows per SM): profile target for the 32 x 256 row kernel
import sys, numpy as np, torch
sys.path.insert(0, '.')
from sfft_b200.plan import Plan
N0, N1 = 1184, 16384
rng = np.random.default_rng(1)
I = rng.normal(10, 3, (N0, N1)); J = rng.normal(12, 3, (N0, N1))
plan = Plan(N0, N1, 4, 4, 3, 2, True, device=0, storage='fp64')
plan.set_timing(True)
for _ in range(2):
    sol = plan.fit(I, J)
print({k: round(v, 3) for k, v in plan.timings().items()})
