"""Multi-process (gloo, world_size 2, CPU) test of the batch driver: round-robin sharding, the single template-state
broadcast and the result gather.  The numeric backend is a stand-in built on the CPU oracle (tests may use oracle/);
the product backend is sfft_b200.plan.Plan, exercised on the GPU in test_gpu_parity.py."""
import os
import sys
import socket
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_indices_partition():
    from sfft_b200.batch import shard_indices
    for n in (0, 1, 7, 64):
        for world in (1, 2, 3, 8):
            seen = sorted(k for r in range(world) for k in shard_indices(n, r, world))
            assert seen == list(range(n))
    assert shard_indices(10, 1, 4) == [1, 5, 9]
    with pytest.raises(ValueError):
        shard_indices(4, 4, 4)


class _OracleBackend:
    """Stand-in with the Plan template interface; the 'state' is the pair of template images as a byte tensor."""

    def __init__(self, N, w, DK, DB):
        import torch
        from oracle import sfft_oracle as orc
        self.orc, self.P = orc, orc.ssc_params(N, N, w, DK, DB, True)
        self.N = N
        self.state = torch.zeros(2 * N * N * 8, dtype=torch.uint8)
        self.ready = False
        self.prepared = 0

    def template_prepare(self, I, mI):
        import torch
        buf = np.concatenate([np.asarray(I, np.float64).ravel(), np.asarray(mI, np.float64).ravel()])
        self.state.copy_(torch.from_numpy(buf.view(np.uint8)))
        self.ready, self.prepared = True, self.prepared + 1

    def template_state_tensor(self):
        return self.state

    def template_mark_ready(self):
        self.ready = True

    def gss_template(self, J, mJ):
        assert self.ready
        a = self.state.numpy().view(np.float64)
        I, mI = a[:self.N * self.N].reshape(self.N, self.N), a[self.N * self.N:].reshape(self.N, self.N)
        sol, diff, _ = self.orc.gss(I, J, mI, mJ, self.P)
        return sol, diff


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    import torch.distributed as dist
    from sfft_b200.batch import TemplateBatch, gather_results
    from sfft_b200.synth import make_pair
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        N, w, DK, DB, ntiles = 48, 2, 1, 1, 5
        be = _OracleBackend(N, w, DK, DB)
        tb = TemplateBatch(be, rank, world)
        base = make_pair(N, N, seed=11)
        # only the source rank has the template
        if rank == 0:
            tb.set_template(base['REF'], base['mREF'])
        else:
            tb.set_template()
        rng = np.random.default_rng(5)
        tiles = [(base['SCI'] + rng.normal(0, 0.5, (N, N)),) * 2 for _ in range(ntiles)]
        res = tb.run(tiles)
        mine = {k: np.array([v[1].sum(), v[0][0]]) for k, v in res.items()}
        allr = gather_results(mine, ntiles, rank, world)
        q.put((rank, sorted(res.keys()), be.prepared, {k: v.tolist() for k, v in allr.items()}))
    finally:
        dist.destroy_process_group()


def test_template_batch_world2_gloo():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    out.sort()
    assert out[0][1] == [0, 2, 4] and out[1][1] == [1, 3]
    assert out[0][2] == 1 and out[1][2] == 0            # the template was transformed on the source rank only
    assert out[0][3] == out[1][3] and len(out[0][3]) == 5
    # single-process reference of the same batch
    from sfft_b200.synth import make_pair
    from oracle import sfft_oracle as orc
    N, w, DK, DB = 48, 2, 1, 1
    base = make_pair(N, N, seed=11)
    rng = np.random.default_rng(5)
    P = orc.ssc_params(N, N, w, DK, DB, True)
    for k in range(5):
        J = base['SCI'] + rng.normal(0, 0.5, (N, N))
        sol, diff, _ = orc.gss(base['REF'], J, base['mREF'], J, P)
        np.testing.assert_allclose(out[0][3][k], [diff.sum(), sol[0]], rtol=1e-9, atol=1e-9)


class _FakePlan:
    """Records what a TemplatePipeline asks of its plans (no GPU): a tile's 'result' is (plan id, tile tag)."""
    count = 0

    def __init__(self, *a, **k):
        self.id = _FakePlan.count
        _FakePlan.count += 1
        self.log, self.pending, self.factor, self.cache, self.tiles = [], None, False, False, 0

    def template_prepare(self, I, mI):
        self.log.append('prepare'); self.factor = self.cache = False; self.tiles = 0

    def template_clone(self, src):
        self.log.append(('clone', src.id, src.factor, src.cache)); self.factor, self.cache = src.factor, src.cache

    def gss_template_submit(self, J, mJ, out_dtype=None, Solution_out=None, DIFF_out=None):
        assert self.pending is None
        self.pending = (self.id, J, 'cached' if self.factor else 'factorised')
        if self.factor:
            self.cache = True                # first cached-factor tile builds the spectra cache
        self.factor = True
        self.tiles += 1

    def gss_finish(self):
        r, self.pending = self.pending, None
        assert r is not None
        return r

    def close(self):
        pass


def test_template_pipeline_start_up_and_order(monkeypatch):
    """Host logic of batch.TemplatePipeline without a GPU: the first two tiles of a template go through the first plan (factorisation,
    spectra cache), the third submission copies that state to the other plans (one clone each, of a plan that holds factor AND cache),
    every later tile runs from a cached factor round-robin, results come back in submission order, a new template starts over."""
    import sfft_b200.plan as planmod
    from sfft_b200.batch import TemplatePipeline
    monkeypatch.setattr(planmod, 'Plan', _FakePlan)
    _FakePlan.count = 0
    pipe = TemplatePipeline(64, 64, 2, depth=3)
    assert [p.id for p in pipe.plans] == [0, 1, 2]
    for rnd in range(2):
        pipe.set_template('I%d' % rnd, 'mI%d' % rnd)
        for p in pipe.plans[1:]:
            assert p.log[-1] == ('clone', 0, False, False)           # spectra only so far
        got = pipe.run([(t, t) for t in range(8)])
        assert [g[1] for g in got] == list(range(8))                 # submission order
        assert [g[0] for g in got] == [0, 0, 0, 1, 2, 0, 1, 2]        # two start-up tiles on plan 0, then round-robin
        assert [g[2] for g in got] == ['factorised'] + ['cached'] * 7
        for p in pipe.plans[1:]:
            assert p.log[-1] == ('clone', 0, True, True)             # shared once, after factor and cache existed
            assert sum(1 for e in p.log if isinstance(e, tuple)) == 2 * (rnd + 1)
    # a short batch never needs the other plans
    pipe.set_template('I', 'mI')
    assert [g[0] for g in pipe.run([(0, 0), (1, 1)])] == [0, 0]
    pipe.close()
