"""
Batch / multi-GPU driver: independent image pairs sharded over the ranks of one node.

The reference's only multi-GPU strategy is task parallelism over pairs, one Python thread per GPU pulling tasks from a
status dict (sfft/MultiEasySparsePacket.py:391-420, 510-552, 931-943); there is no collective anywhere in it.  Here it
is one process per GPU (torch.distributed) with a static round-robin shard of the pair list, no data-path collective,
and -- when every science tile shares one template that is the convolved image (ForceConv='REF') -- ONE broadcast of
the template's row-spectrum state (NCCL over NVLink on GPUs; gloo in the CPU tests), after which a tile costs only its
own row transforms (BASELINE config 4).
"""
import numpy as np

__all__ = ['shard_indices', 'TemplateBatch', 'gather_results']


def shard_indices(n_items, rank, world):
    """Static round-robin partition (item k -> rank k % world), like the task polling order of MESP_Cupy."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError('bad rank/world')
    return list(range(rank, n_items, world))


def gather_results(local, n_items, rank, world, group=None):
    """Gather {index: small numpy vector} dicts onto rank 0 (flux scalings, checksums ... not images)."""
    if world == 1:
        return dict(local)
    import torch.distributed as dist
    out = [None] * world
    dist.all_gather_object(out, {int(k): np.asarray(v) for k, v in local.items()}, group=group)
    merged = {}
    for d in out:
        merged.update(d)
    if len(merged) != n_items:
        raise RuntimeError('batch incomplete: %d of %d results' % (len(merged), n_items))
    return merged


class TemplateBatch:
    """Shared-template batch on one rank.

    backend: object with the Plan template interface (template_prepare, template_state_tensor, template_mark_ready,
    gss_template); the product passes sfft_b200.plan.Plan, the CPU tests pass a stand-in so that the sharding and the
    broadcast protocol are exercised without a GPU."""

    def __init__(self, backend, rank=0, world=1, group=None, src=0):
        self.backend, self.rank, self.world, self.group, self.src = backend, int(rank), int(world), group, int(src)
        self.ready = False

    def set_template(self, PixA_REF=None, PixA_mREF=None):
        """Rank `src` transforms the template; everyone else receives the state with one broadcast."""
        if self.rank == self.src:
            if PixA_REF is None or PixA_mREF is None:
                raise ValueError('the source rank needs the template images')
            self.backend.template_prepare(PixA_REF, PixA_mREF)
        if self.world > 1:
            import torch.distributed as dist
            state = self.backend.template_state_tensor()
            dist.broadcast(state, src=self.src, group=self.group)
            if self.rank != self.src:
                self.backend.template_mark_ready()
        self.ready = True

    def run(self, tiles, n_items=None):
        """tiles: callable k -> (PixA_SCI, PixA_mSCI) or a sequence; processes this rank's shard.
        Returns {k: (Solution, PixA_DIFF)}."""
        if not self.ready:
            raise RuntimeError('set_template() first')
        n = len(tiles) if n_items is None else n_items
        get = tiles if callable(tiles) else (lambda k: tiles[k])
        out = {}
        for k in shard_indices(n, self.rank, self.world):
            J, mJ = get(k)
            out[k] = self.backend.gss_template(J, mJ)
        return out
