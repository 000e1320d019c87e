import sys, numpy as np
sys.path.insert(0, '.')
from sfft_b200.plan import Plan
from sfft_b200.synth import make_pair
N0, N1, w, DK, DB, st = [int(v) if v.isdigit() else v for v in sys.argv[1:7]]
d = make_pair(N0, N1, seed=N0 + N1 + w)
plan = Plan(N0, N1, w, w, DK, DB, True, storage=st)
for _ in range(2):
    sol = plan.fit(d['REF'], d['SCI'])
print('ok', float(np.abs(sol).max()))
