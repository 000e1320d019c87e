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
