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

__all__ = ['shard_indices', 'TemplateBatch', 'gather_results', 'PairPipeline', 'TemplatePipeline', 'run_pairs_threaded']


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
            if state.is_cuda:
                # NCCL broadcasts are ordered against torch's current stream only and do not block the host, while the
                # plan works on its own stream: wait until the state has landed before any tile may read it
                import torch
                torch.cuda.synchronize(state.device)
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


def sparse_delta(unmasked, masked):
    """(idx, val): flat C-order indices where `masked` differs from `unmasked` and the masked values there -- the form
    sfftb_gss_submit_delta takes.  NaN == NaN counts as equal."""
    u, m = np.asarray(unmasked), np.asarray(masked)
    diff = ~((m == u) | (np.isnan(m) & np.isnan(u)))
    idx = np.flatnonzero(diff.ravel()).astype(np.int64)
    return idx, np.ascontiguousarray(m.ravel()[idx])


class PairPipeline:
    """A queue of independent image pairs in host memory through ONE GPU (what a worker thread of
    sfft/MultiEasySparsePacket.py:568-649 does pair after pair): two plans share one compute stream and are driven
    alternately through sfftb_gss_submit / sfftb_gss_finish, so the host-to-device copies of pair k + 1 run while the
    kernels and the device-to-host copy of pair k are in flight.  Results come back in submission order."""

    def __init__(self, N0, N1, KerHW, KerPolyOrder=2, BGPolyOrder=2, ConstPhotRatio=True, device=0, storage='fp64',
                 stream_ptr=None, depth=2, solver_sms=0):
        from .plan import Plan
        self.plans = [Plan(N0, N1, KerHW, KerHW, KerPolyOrder, BGPolyOrder, ConstPhotRatio, device=device, storage=storage)
                      for _ in range(depth)]
        if solver_sms:
            # SM partition: every plan on its OWN stream; the Cholesky of pair k occupies `solver_sms` SMs while the row
            # and column passes of pair k + 1 run on the others (sfftb_plan_set_partition)
            for pl in self.plans:
                pl.set_stream(0)
                pl.set_partition(solver_sms)
        elif not stream_ptr:
            # one dedicated compute stream for all plans: their kernels serialise (two full-grid cooperative Cholesky
            # kernels must never wait for each other's SMs), only the copies on the plans' copy streams overlap
            import torch
            self._stream = torch.cuda.Stream(device=torch.device('cuda', device))
            stream_ptr = self._stream.cuda_stream
        if not solver_sms:
            for pl in self.plans:
                pl.set_stream(stream_ptr)
        self._busy = [False] * depth
        self._k = 0

    def submit(self, PixA_I, PixA_J, PixA_mI, PixA_mJ, out_dtype=np.float64, Solution_out=None, DIFF_out=None):
        """Returns the result of the pair that had to leave its slot to make room (or None)."""
        slot = self._k % len(self.plans)
        done = None
        if self._busy[slot]:
            done = self.plans[slot].gss_finish()
        self.plans[slot].gss_submit(PixA_I, PixA_J, PixA_mI, PixA_mJ, out_dtype, Solution_out, DIFF_out)
        self._busy[slot] = True
        self._k += 1
        return done

    def submit_delta(self, PixA_I, PixA_J, delta_I, delta_J, out_dtype=np.float64, Solution_out=None, DIFF_out=None):
        """submit() with the masked pair as sparse deltas (see sparse_delta): half the host-to-device bytes."""
        slot = self._k % len(self.plans)
        done = None
        if self._busy[slot]:
            done = self.plans[slot].gss_finish()
        self.plans[slot].gss_submit_delta(PixA_I, PixA_J, delta_I, delta_J, out_dtype, Solution_out, DIFF_out)
        self._busy[slot] = True
        self._k += 1
        return done

    def submit_device(self, pI, pJ, pmI, pmJ, img_dtype, psol, pdiff, diff_dtype):
        """Device-resident pair (raw device pointers): queued without a host synchronisation."""
        slot = self._k % len(self.plans)
        if self._busy[slot]:
            self.plans[slot].gss_finish()
        self.plans[slot].gss_submit_device(pI, pJ, pmI, pmJ, img_dtype, psol, pdiff, diff_dtype)
        self._busy[slot] = True
        self._k += 1

    def drain(self):
        """Finish everything in flight, oldest first."""
        out = []
        n = len(self.plans)
        for d in range(n):
            slot = (self._k + d) % n
            if self._busy[slot]:
                out.append(self.plans[slot].gss_finish())
                self._busy[slot] = False
        return out

    def run(self, pairs, out_dtype=np.float64):
        """pairs: iterable of (I, J, mI, mJ) host arrays -> list of (Solution, DIFF) in order."""
        res = []
        for (I, J, mI, mJ) in pairs:
            r = self.submit(I, J, mI, mJ, out_dtype)
            if r is not None:
                res.append(r)
        res.extend(self.drain())
        return res

    def close(self):
        for pl in self.plans:
            pl.close()


class TemplatePipeline:
    """Science tiles against ONE shared template on one GPU (BASELINE config 4): `depth` plans hold the same template state;
    tiles go through sfftb_gss_template_submit / sfftb_gss_finish alternately, so the copies (host tiles) and kernels of
    tile k + 1 run under the kernels and the device-to-host copy of tile k.  Under torch.distributed the first plan is
    the one that received TemplateBatch's single broadcast (`first_plan`)."""

    def __init__(self, N0, N1, KerHW, KerPolyOrder=2, BGPolyOrder=2, ConstPhotRatio=True, device=0, storage='fp64',
                 stream_ptr=None, first_plan=None, depth=2):
        import torch
        from .plan import Plan
        self.device = device
        depth = max(2, int(depth))
        mk = lambda: Plan(N0, N1, KerHW, KerHW, KerPolyOrder, BGPolyOrder, ConstPhotRatio, device=device, storage=storage)
        self.plans = [first_plan if first_plan is not None else mk()] + [mk() for _ in range(depth - 1)]
        self._own = [first_plan is None] + [True] * (depth - 1)
        # stream_ptr given: both plans queue on it (tiles strictly one after the other).  None / 0: every plan keeps its OWN
        # stream, so the kernels of tile k + 1 also overlap the latency-bound substitutions of tile k; that is safe because
        # the only cooperative kernel of a cached-factor tile is the substitution kernel, whose grid is capped at half of the
        # SMs, and the factorising first tile of each plan completes synchronously inside its submit call
        if stream_ptr:
            for pl in self.plans:
                pl.set_stream(stream_ptr)
        self._busy = [False] * depth
        self._k = 0
        self._shared = False
        self._warm = 0

    def set_template(self, PixA_I=None, PixA_mI=None):
        """Prepare the template on the first plan (unless it already holds one, e.g. received by broadcast).  The other plans
        receive its state by share_state(): submit() does that after the first plan has processed two tiles, so that the
        Cholesky factor of the template's normal equations and the cached segment spectra are computed ONCE per template
        and copied (device-to-device), not once per plan."""
        if PixA_I is not None:
            self.plans[0].template_prepare(PixA_I, PixA_mI)
        self._shared = False
        self._warm = 0
        for pl in self.plans[1:]:
            pl.template_clone(self.plans[0])               # spectra only at this point: every plan is usable right away

    def share_state(self):
        """Copy the first plan's template state (spectra + factor + lag rows + cached segment spectra, whatever it holds) into
        the other plans.  Call with no tile in flight."""
        for pl in self.plans[1:]:
            pl.template_clone(self.plans[0])
        self._shared = True

    def submit(self, PixA_J, PixA_mJ, out_dtype=np.float64, Solution_out=None, DIFF_out=None):
        if not self._shared:
            # start-up: the first two tiles go through the first plan one after the other (factorisation, then the tile that
            # builds the spectra cache); the third submission copies that state to the other plans
            done = self.plans[0].gss_finish() if self._busy[0] else None
            self._busy[0] = False
            if self._warm == 2:
                self.share_state()
            else:
                self.plans[0].gss_template_submit(PixA_J, PixA_mJ, out_dtype, Solution_out, DIFF_out)
                self._busy[0] = True
                self._warm += 1
                return done
            # (falls through: this submission is the first of the round-robin phase; `done` is the second tile's result)
            slot = self._k % len(self.plans)
            self.plans[slot].gss_template_submit(PixA_J, PixA_mJ, out_dtype, Solution_out, DIFF_out)
            self._busy[slot] = True
            self._k += 1
            return done
        slot = self._k % len(self.plans)
        done = None
        if self._busy[slot]:
            done = self.plans[slot].gss_finish()
        self.plans[slot].gss_template_submit(PixA_J, PixA_mJ, out_dtype, Solution_out, DIFF_out)
        self._busy[slot] = True
        self._k += 1
        return done

    def drain(self):
        out = []
        n = len(self.plans)
        for d in range(n):
            slot = (self._k + d) % n
            if self._busy[slot]:
                out.append(self.plans[slot].gss_finish())
                self._busy[slot] = False
        return out

    def run(self, tiles, out_dtype=np.float64):
        res = []
        for (J, mJ) in tiles:
            r = self.submit(J, mJ, out_dtype)
            if r is not None:
                res.append(r)
        res.extend(self.drain())
        return res

    def close(self):
        for pl, own in zip(self.plans, self._own):
            if own:
                pl.close()


def run_pairs_threaded(pairs, devices, KerHW, KerPolyOrder=2, BGPolyOrder=2, ConstPhotRatio=True, storage='fp64',
                       out_dtype=np.float64):
    """The reference's own multi-GPU model (MultiEasy_SparsePacket.MESP_Cupy, sfft/MultiEasySparsePacket.py:391-420,
    510-552, 931-943): ONE process, one host thread per GPU, every thread pulls the next pair from a shared queue and
    runs it on its device -- here through a PairPipeline per thread, so that the copies of the next pair overlap the
    kernels of the current one.  `pairs`: list of (I, J, mI, mJ) host arrays of one shape; `devices`: CUDA ordinals (an
    ordinal may appear twice).  Returns the list of (Solution, DIFF) in the order of `pairs`."""
    import threading
    import torch
    if not pairs:
        return []
    N0, N1 = pairs[0][0].shape
    results = [None] * len(pairs)
    lock = threading.Lock()
    state = {'next': 0}
    errors = []

    def worker(dev):
        try:
            with torch.cuda.device(dev):
                pipe = PairPipeline(N0, N1, KerHW, KerPolyOrder, BGPolyOrder, ConstPhotRatio, device=dev, storage=storage)
                order = []
                while True:
                    with lock:                       # the task-status dict of MESP_Cupy (:510-552) as a counter
                        k = state['next']
                        state['next'] += 1
                    if k >= len(pairs):
                        break
                    done = pipe.submit(*pairs[k], out_dtype=out_dtype)
                    order.append(k)
                    if done is not None:
                        results[order.pop(0)] = done
                for done in pipe.drain():
                    results[order.pop(0)] = done
                pipe.close()
        except BaseException as e:                    # noqa: BLE001 -- reported to the caller below
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(int(d),)) for d in devices]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return results
