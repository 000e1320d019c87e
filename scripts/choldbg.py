import os, sys, numpy as np
sys.path.insert(0, '/root/repo')
from sfft_b200.plan import Plan
from sfft_b200.synth import make_pair
d = make_pair(1024, 1024, seed=5)
plan = Plan(1024, 1024, 8, 8, 2, 2, True, storage='fp64')
plan.fit(d['REF'], d['SCI'])
