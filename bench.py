#!/usr/bin/env python
"""
bench.py -- Mpix/s per 4K x 4K SFFT subtraction (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps 50 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference algorithm's CPU port on the host cores

A "step" is one GeneralSFFTSubtract.GSS (fit on the masked pair + Fourier-space apply on the unmasked pair) on the
BASELINE config 2 workload: 4096 x 4096 DECam-like synthetic pair, KerHW=8, KerPolyOrder=2, BGPolyOrder=2, fp32
images and fp32 spectra in HBM (all arithmetic fp64).  Every rank owns an independent pair (weak scaling; no
data-path collective).  `value` is device-timed with inputs resident in HBM; `e2e` goes through the C-ABI call
with pinned HOST buffers (H2D of the four images and D2H of the difference image inside the timed region).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (N0, N1, KerHW, DK, DB, storage, seed-id)
    'c2_4096_w8_dk2_db2_fp32': (4096, 4096, 8, 2, 2, 'fp32', 2),
    'c2_4096_w8_dk2_db2_fp64': (4096, 4096, 8, 2, 2, 'fp64', 2),
    'c1_512_w4_dk0_db0_fp64': (512, 512, 4, 0, 0, 'fp64', 1),
    'c4_2048_w8_dk2_db2_fp32': (2048, 2048, 8, 2, 2, 'fp32', 4),
    'c5_16384_w12_dk3_db2_fp64': (16384, 16384, 12, 3, 2, 'fp64', 5),
    'c5half_8192_w12_dk3_db2_fp64': (8192, 8192, 12, 3, 2, 'fp64', 5),
    'dev_1024_w4_dk2_db2_fp32': (1024, 1024, 4, 2, 2, 'fp32', 2),
    # BASELINE config 4: science tiles against ONE shared template; the template row spectra are computed on rank 0
    # and broadcast once (NCCL), a step is one tile through sfftb_gss_template
    'c4_template_2048_w8_dk2_db2_fp32': (2048, 2048, 8, 2, 2, 'fp32', 4),
    # BASELINE config 3: B-spline spatially varying kernel through sfft_b200.BSplineSFFT (general-basis plan); see C3_SPEC
    'c3_6144_bspline_fp32': (6144, 6144, 8, 2, 2, 'fp32', 3),
    'c3dev_1536_bspline_fp32': (1536, 1536, 8, 2, 2, 'fp32', 3),
}
# B-spline recipe of config 3 (SURVEY.md 8d assumed Fij = 25, ScaFij = 6): quadratic B-spline kernel with two internal knots
# per axis (Fi = Fj = 5), SEPARATE-VARYING photometric scaling as a quadratic polynomial, quadratic polynomial background
def c3_spec(N0, N1):
    return dict(KerSpType='B-Spline', KerSpDegree=2, KerIntKnotX=[N0 / 3.0, 2.0 * N0 / 3.0], KerIntKnotY=[N1 / 3.0, 2.0 * N1 / 3.0],
                SEPARATE_SCALING=True, ScaSpType='Polynomial', ScaSpDegree=2, BkgSpType='Polynomial', BkgSpDegree=2)
DEFAULT_WORKLOAD = 'c2_4096_w8_dk2_db2_fp32'
# device leg: pairs in flight (one plan + stream each) and the SMs set aside for the Cholesky (sfftb_plan_set_partition)
PIPE_DEPTH = int(os.environ.get('SFFTB_BENCH_DEPTH', '4'))
SOLVER_SMS = int(os.environ.get('SFFTB_BENCH_SOLVER_SMS', '20'))
HBM_FALLBACK_GBS = 6650.0     # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def hbm_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            for k in ('hbm_gbs', 'hbm_gb_s', 'hbm'):
                if k in d:
                    return float(d[k]), 'measured'
        except Exception:
            pass
    return HBM_FALLBACK_GBS, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix='.csv')
            os.close(fd)
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(',')]
                if len(f) < 6:
                    continue
                try:
                    sm.append(float(f[0]))
                    mx.append(float(f[1]))
                except ValueError:
                    continue
                for n, v in zip(names, f[2:6]):
                    if v.lower().startswith('active'):
                        reasons.add(n)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out['sm_mhz'] = statistics.median(sm)
            out['sm_max_mhz'] = max(mx)
            out['samples'] = len(sm)
        out['reasons'] = sorted(reasons)
        return out


def pin_to_gpu_numa_node(local):
    """Bind this process to the CPUs of the NUMA node its GPU hangs off (sysfs; no numactl in the image), so that the
    pinned host buffers, which are first touched by this process, are allocated on that node.  Returns the node or None."""
    try:
        out = subprocess.run(['nvidia-smi', '-i', str(local), '--query-gpu=pci.bus_id', '--format=csv,noheader'],
                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=20).stdout.strip()
        bus = out.lower()
        if bus.startswith('00000000:'):
            bus = bus[4:]
        node = int(open('/sys/bus/pci/devices/%s/numa_node' % bus).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open('/sys/devices/system/node/node%d/cpulist' % node).read().strip().split(','):
            a, _, b = part.partition('-')
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


def link_probe(torch, dist, dev, world, nbytes, reps=4):
    """Bare pinned-host <-> device copies of one image-sized buffer in both directions at once, on all ranks at the same
    time: the ceiling the end-to-end leg can reach on this box at this N (PCIe link per GPU, shared root complexes and
    host memory when N > 1).  GB/s per GPU = the slowest rank's."""
    try:
        h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        d_in = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        d_out = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        res = {}
        for mode in ('h2d', 'd2h', 'duplex'):
            for timed in (False, True):
                torch.cuda.synchronize(dev)
                if world > 1:
                    dist.barrier()
                t0 = time.perf_counter()
                for _ in range(reps):
                    if mode in ('h2d', 'duplex'):
                        with torch.cuda.stream(s1):
                            d_in.copy_(h_in, non_blocking=True)
                    if mode in ('d2h', 'duplex'):
                        with torch.cuda.stream(s2):
                            h_out.copy_(d_out, non_blocking=True)
                torch.cuda.synchronize(dev)
                dt = time.perf_counter() - t0
                if timed:
                    t = torch.tensor([dt], dtype=torch.float64, device=dev)
                    if world > 1:
                        dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    res[mode] = reps * nbytes / float(t[0]) / 1e9
        return {'h2d_gbs_per_gpu': res['h2d'], 'd2h_gbs_per_gpu': res['d2h'],
                'duplex_each_way_gbs_per_gpu': res['duplex'], 'aggregate_h2d_gbs': world * res['h2d'],
                'buffer_bytes': nbytes, 'ranks': world,
                'how': 'pinned cudaMemcpyAsync of one image-sized buffer, %d repeats, all ranks at once, slowest rank' % reps}
    except Exception as e:                      # the probe must never take the bench line down
        return None


def run_config4(torch, dist, B, world, rank, local, dev, ntiles=64, side=2048, w=8, DK=2, DB=2, storage='fp32'):
    """BASELINE config 4 inside the default run: 64 science tiles of 2048^2 against ONE shared template, sharded over the
    ranks; rank 0 transforms the template once, its state goes to the other GPUs with ONE NCCL broadcast (timed after a
    warm-up collective, so communicator set-up is excluded), every rank then runs its tiles device-resident through
    TemplatePipeline (two tiles in flight).  Returns the sub-object of the JSON line (rank 0) or None."""
    from sfft_b200.plan import Plan
    from sfft_b200.batch import TemplateBatch, TemplatePipeline, shard_indices
    from sfft_b200.synth import make_pair, CONFIG_SEEDS
    tdt = torch.float32 if storage == 'fp32' else torch.float64
    code = B.F32 if storage == 'fp32' else B.F64
    d = make_pair(side, side, CONFIG_SEEDS[4])                       # the same template on every rank (only rank 0 uses it)
    rng = np.random.default_rng(4000 + rank)
    npdt = np.float32 if storage == 'fp32' else np.float64
    dev_of = lambda a: torch.from_numpy(np.ascontiguousarray(a.astype(npdt))).to(dev)
    REF, mREF = dev_of(d['REF']), dev_of(d['mREF'])
    masked = d['mSCI'] != d['SCI']
    tiles = []
    for k in range(2):                                               # two distinct science tiles per rank, cycled
        sci = d['SCI'] + rng.normal(0.0, 0.5, d['SCI'].shape)
        msci = np.where(masked, 0.0, sci)
        tiles.append((dev_of(sci), dev_of(msci)))
    plan = Plan(side, side, w, w, DK, DB, True, device=local, storage=storage)
    tb = TemplateBatch(plan, rank, world)
    if world > 1:
        warm = torch.zeros(1024, device=dev)
        dist.broadcast(warm, src=0)                                  # communicator warm-up, not part of the number
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    if rank == 0:
        plan.template_prepare(REF, mREF)
    t_prep = time.perf_counter() - t0
    torch.cuda.synchronize(dev)
    t1 = time.perf_counter()
    tb.ready = False
    if world > 1:
        state = plan.template_state_tensor()
        dist.broadcast(state, src=0)
        torch.cuda.synchronize(dev)
        if rank != 0:
            plan.template_mark_ready()
    t_bcast = time.perf_counter() - t1
    TD = int(os.environ.get('SFFTB_BENCH_TILE_DEPTH', '6'))          # tiles in flight (one plan + stream each)
    tp = TemplatePipeline(side, side, w, DK, DB, True, device=local, storage=storage, stream_ptr=None, first_plan=plan, depth=TD)
    tp.set_template()
    mine = list(shard_indices(ntiles, rank, world))
    diffs = [torch.empty((side, side), dtype=tdt, device=dev) for _ in range(TD)]
    sols = [torch.empty(plan.NEQ, dtype=torch.float64, device=dev) for _ in range(TD)]
    busy = [False] * TD

    def run_tiles(idxs):
        for n_, k in enumerate(idxs):
            slot = n_ % TD
            if busy[slot]:
                tp.plans[slot].gss_finish()
            J, mJ = tiles[k % 2]
            tp.plans[slot].gss_template_submit_device(J.data_ptr(), mJ.data_ptr(), code, sols[slot].data_ptr(), diffs[slot].data_ptr(), code)
            busy[slot] = True
        for slot in range(TD):
            if busy[slot]:
                tp.plans[slot].gss_finish()
                busy[slot] = False
    # ONE factorisation and ONE spectra cache per template and GPU: two tiles on the first plan, whose state (factor, lag rows, cached
    # segment spectra) is then copied into the other plans (sfftb_template_clone)
    for k in range(2):
        J, mJ = tiles[k % 2]
        tp.plans[0].gss_template_submit_device(J.data_ptr(), mJ.data_ptr(), code, sols[0].data_ptr(), diffs[0].data_ptr(), code)
        tp.plans[0].gss_finish()
    tp.share_state()
    run_tiles(range(2 * TD))                                         # warm-up of every plan (all served from the shared state)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t2 = time.perf_counter()
    run_tiles(mine)
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t2
    t = torch.tensor([dt, t_bcast, t_prep], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    state_bytes = plan.template_state()[1]
    tp.close()
    plan.close()
    if rank != 0:
        return None
    dt, t_bcast = float(t[0]), float(t[1])
    return {'workload': 'c4: %d science tiles of %dx%d against one shared template, KerHW=%d DK=%d DB=%d, %s storage' % (ntiles, side, side, w, DK, DB, storage),
            'value': ntiles * side * side / 1e6 / dt, 'unit': 'Mpix/s', 'n_gpus': world, 'tiles': ntiles,
            'tiles_per_gpu': len(mine), 'ms_per_tile_per_gpu': dt * 1e3 / max(1, len(mine)), 'tiles_in_flight': TD,
            'template_prepare_ms': float(t[2]) * 1e3, 'template_broadcast_ms': t_bcast * 1e3 if world > 1 else None,
            'template_state_bytes': state_bytes,
            'collective': 'one torch.distributed.broadcast (NCCL) of the template row spectra, timed after a warm-up collective' if world > 1 else None,
            'timing': 'host clock around the rank\'s tiles after a device synchronise, max over ranks'}


def c3_make(name, rank):
    from sfft_b200.synth import make_pair, CONFIG_SEEDS
    N0, N1, w, DK, DB, storage, sid = WORKLOADS[name]
    d = make_pair(N0, N1, CONFIG_SEEDS[sid] + 1000 * rank)
    return (N0, N1, w, storage), d


def run_c3(args, world, rank, local, numa):
    """BASELINE config 3: 6144^2 pair, B-spline spatially varying kernel through sfft_b200.BSplineSFFT (general-basis
    plan).  Same JSON contract as the default workload; every rank owns a pair (weak scaling, no collective)."""
    stage_overlapped, serial_ms = None, None           # (general-basis plans run one pair at a time)
    import torch
    import torch.distributed as dist
    from sfft_b200 import _lib as B
    import sfft_b200.BSplineSFFT as bs
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    W, K = max(3, args.warmup), max(1, args.steps)
    (N0, N1, w, storage), d = c3_make(args.workload, rank)
    spec = c3_spec(N0, N1)
    npdt = np.float32 if storage == 'fp32' else np.float64
    tdt = torch.float32 if storage == 'fp32' else torch.float64
    code = B.F32 if storage == 'fp32' else B.F64
    esz = 4 if storage == 'fp32' else 8
    host = {k: torch.from_numpy(np.ascontiguousarray(v.astype(npdt))).pin_memory() for k, v in d.items()}
    devt = {k: v.to(dev) for k, v in host.items()}
    cfg = bs.SingleSFFTConfigure.SSC(NX=N0, NY=N1, KerHW=w, VERBOSE_LEVEL=0, CUDA_DEVICE=local, STORAGE=storage, **spec)
    P, plan = cfg[0], cfg[1]['plan']
    stream = torch.cuda.current_stream(dev)
    plan.bind_torch_stream(stream)
    plan.set_timing(True)
    diff_d = torch.empty((N0, N1), dtype=tdt, device=dev)
    diff_h = torch.empty((N0, N1), dtype=tdt).pin_memory()
    sol_d = torch.empty(plan.NEQ, dtype=torch.float64, device=dev)
    sol_h = np.empty(plan.NEQ, np.float64)
    L = B.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_device():
        plan.gss_device(devt['REF'].data_ptr(), devt['SCI'].data_ptr(), devt['mREF'].data_ptr(), devt['mSCI'].data_ptr(),
                        code, sol_d.data_ptr(), diff_d.data_ptr(), code)

    def step_host():
        B.check(L.sfftb_gss(plan._h, host['REF'].data_ptr(), host['SCI'].data_ptr(), host['mREF'].data_ptr(),
                            host['mSCI'].data_ptr(), B.MEM_HOST, code, sol_h.ctypes.data, B.MEM_HOST,
                            diff_h.data_ptr(), B.MEM_HOST, code))
    for _ in range(W):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = plan.launch_count
    stage = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(K):
        step_device()
        for k, v in plan.timings().items():
            stage[k] = stage.get(k, 0.0) + v
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1) / K
    launches = plan.launch_count - l0
    stage = {k: v / K for k, v in stage.items()}
    KE = args.e2e_steps or min(K, 5)
    step_host()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e2.record(stream)
    for _ in range(KE):
        step_host()
    e3.record(stream)
    barrier()
    ms_e2e = max((time.perf_counter() - t0) * 1e3 / KE, e2.elapsed_time(e3) / KE)
    clocks = sampler.stop()
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    if rank == 0:
        mpix = N0 * N1 / 1e6
        info = plan.gen_info()
        NH = N1 // 2 + 1
        csz = 2 * esz
        peak, which = hbm_peak()
        # dominant kernel: the block passes of fit_gen_kernel.  Per launch a pass must read the stored planes it stages once
        # and write its lag rows; summed over the passes of one fit: staged planes total x NH x N0 complex + all lag rows
        alg_bytes = info['staged_planes_total'] * NH * N0 * csz + info['lag_rows'] * NH * 16
        t_kernel = stage.get('fit_cols', 0.0) / 1e3                  # all passes of the kernel (bytes and time both summed over the passes)
        achieved = alg_bytes / t_kernel / 1e9 if t_kernel > 0 else None
        n_pl = P['Fij'] + P.get('ScaFij', 0) + 1
        step_bytes = (4 * n_pl + 7) * N0 * N1 * esz                  # SURVEY.md 8d, SEPARATE-VARYING: n_pl = Fij + ScaFij + 1
        out = {
            'metric': 'Mpix/s per 6Kx6K B-spline SFFT subtraction (GSS: fit + apply)', 'value': world * mpix / (ms / 1e3), 'unit': 'Mpix/s',
            'n_gpus': world, 'steps': K, 'warmup': W, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': args.workload, 'image': [N0, N1], 'KerHW': w, 'storage': storage, 'arithmetic': 'fp64',
                       'spec': {k: (v if not isinstance(v, list) else [float(x) for x in v]) for k, v in spec.items()},
                       'Fij': P['Fij'], 'ScaFij': P.get('ScaFij'), 'Fpq': P['Fpq'], 'NEQ': P['NEQ'], 'NEQt': P['NEQt'],
                       'SCALING_MODE': P['SCALING_MODE'], 'plan': info, 'pairs_per_step_per_gpu': 1,
                       'l2_policy': 'working set per step (inputs 4x%.0f MB, %d stored planes of %.0f MB) exceeds the 126 MB L2' % (
                           N0 * N1 * esz / 1e6, info['stored_planes'], NH * N0 * csz / 1e6)},
            'stage_ms': stage,
            'stage_ms_note': ('one pair at a time on the whole GPU (5 blocking calls after the timed region): kernel durations; '
                              'the timed region itself keeps %d pairs in flight' % PIPE_DEPTH) if stage_overlapped is not None else 'timed region',
            'stage_ms_pairs_in_flight': stage_overlapped,
            'one_pair_at_a_time_ms': serial_ms,
            'assembly_solve_ms': stage.get('fit_cols', 0) + stage.get('fit_reduce_fill', 0) + stage.get('fit_solve', 0),
            'solver': plan.last_solver, 'clocks': clocks,
            'e2e': {'value': world * mpix / (ms_e2e / 1e3), 'unit': 'Mpix/s', 'ms_per_step': ms_e2e, 'steps': KE,
                    'mode': 'one blocking sfftb_gss call per step (host buffers)', 'h2d_bytes_per_step': 4 * N0 * N1 * esz,
                    'd2h_bytes_per_step': N0 * N1 * esz + plan.NEQ * 8, 'host_numa_node': numa},
            'gpu_launches': launches,
            'roofline': {'bound': 'hbm', 'kernel': 'fit_gen4_kernel', 'achieved': achieved, 'peak': peak, 'peak_source': which, 'unit': 'GB/s',
                         'frac': (achieved / peak) if achieved else None, 'traffic': None,
                         'algorithmic_bytes_per_launch': alg_bytes / max(1, info['passes']), 'launches_per_step_of_kernel': info['passes'],
                         'kernel_ms': t_kernel * 1e3, 'step_algorithmic_bytes': step_bytes,
                         'step_frac': step_bytes / (ms / 1e3) / 1e9 / peak,
                         'note': 'fp64-issue / shared-memory bound like the polynomial fit kernel (DESIGN.md section 4); %d passes of up to '
                                 '5 x 5 pair accumulators, n = %d Cholesky' % (info['passes'], info['unknowns'])},
            'cpu_baseline': None,
        }
        if world == 1 and not args.no_cpu_baseline:
            out['cpu_baseline'] = cpu_port_c3(args.workload, min(args.cpu_sample, 256))
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def cpu_port_c3(name, side):
    """The B-spline oracle (design-matrix restatement, dense D: small crops only) on the host cores."""
    from oracle import bspline_oracle as bo
    (N0, N1, w, storage), d = c3_make(name, 0)
    s = min(side, N0, N1)
    w2 = min(w, 2)                               # dense D = s^2 x NEQ doubles: the kernel half width is reduced with the crop
    crop = {k: np.ascontiguousarray(v[:s, :s]) for k, v in d.items()}
    P = bo.ssc_params(s, s, w2, **c3_spec(s, s))
    t0 = time.time()
    bo.gss(crop['REF'], crop['SCI'], crop['mREF'], crop['mSCI'], P)
    dt = time.time() - t0
    return {'value': (s * s / 1e6) / dt, 'unit': 'Mpix/s', 'cores': os.cpu_count(), 'kind': 'port', 'seconds': dt,
            'sample': '%dx%d crop, KerHW=%d (reduced from %d: the oracle builds the dense design matrix), same B-spline recipe; '
                      'oracle/bspline_oracle.py, BLAS threads of the host' % (s, s, w2, w)}


def run_reference_c3(args, ncores):
    W, K = max(0, args.warmup), max(1, args.steps)
    side = min(args.cpu_sample, 256)
    first = cpu_port_c3(args.workload, side)
    while first['seconds'] * (K + W) > 240.0 and side > 64:
        side //= 2
        first = cpu_port_c3(args.workload, side)
    for _ in range(W):
        cpu_port_c3(args.workload, side)
    t0 = time.time()
    for _ in range(K):
        last = cpu_port_c3(args.workload, side)
    dt = (time.time() - t0) / K
    v = side * side / 1e6 / dt
    print(json.dumps({
        'impl': 'reference', 'metric': 'Mpix/s per 6Kx6K B-spline SFFT subtraction (GSS: fit + apply)', 'value': v, 'unit': 'Mpix/s',
        'n_gpus': args.gpus, 'steps': K, 'warmup': W, 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': args.workload, 'sample': last['sample'], 'same_config': False},
        'cpu_baseline': {'value': v, 'unit': 'Mpix/s', 'cores': ncores, 'kind': 'port', 'sample': last['sample']},
        'e2e': {'value': v, 'unit': 'Mpix/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}))


def make_workload(name, rank):
    from sfft_b200.synth import make_pair, CONFIG_SEEDS
    N0, N1, w, DK, DB, storage, sid = WORKLOADS[name]
    d = make_pair(N0, N1, CONFIG_SEEDS[sid] + 1000 * rank, varying_psf=(sid != 1))
    return (N0, N1, w, DK, DB, storage), d


def cpu_port_mpix(name, sample_side, repeats=1):
    """The oracle (NumPy restatement of the reference NumPy backend, validated against the reference on the golden
    fixtures) timed on the host cores on a bounded crop of the same workload."""
    from oracle import sfft_oracle as orc
    (N0, N1, w, DK, DB, storage), d = make_workload(name, 0)
    s = min(sample_side, N0, N1)
    crop = {k: np.ascontiguousarray(v[:s, :s]) for k, v in d.items()}
    P = orc.ssc_params(s, s, w, DK, DB, True)
    best = None
    for _ in range(repeats):
        t0 = time.time()
        orc.gss(crop['REF'], crop['SCI'], crop['mREF'], crop['mSCI'], P)
        dt = time.time() - t0
        best = dt if best is None else min(best, dt)
    return (s * s / 1e6) / best, best, s, orc.WORKERS


def run_reference(args):
    """The reference arm: the reference's algorithm for this path on the host cores (the NumPy port in oracle/, which is
    validated against the unmodified reference on the golden fixtures; the reference itself is Python + numba + pyFFTW
    and OOMs above 2048^2).  W warm-up steps, then exactly K timed steps; a step is one GSS of a bounded
    crop of the workload's pair (the crop is halved if K + W steps would not finish within a few minutes)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    # the same host threads at every N: torch.distributed.run exports OMP_NUM_THREADS=1 to its workers, which would
    # throttle the BLAS / FFT pools of this leg only when it is launched under torchrun
    ncores = os.cpu_count() or 1
    os.environ['SFFT_ORACLE_WORKERS'] = str(ncores)
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=ncores)
    except Exception:
        pass
    from oracle import sfft_oracle as orc
    orc.WORKERS = ncores
    name = args.workload
    if name.startswith('c3'):
        return run_reference_c3(args, ncores)
    N0, N1, w, DK, DB, storage, sid = WORKLOADS[name]
    W, K = max(0, args.warmup), max(1, args.steps)
    (_, _, _, _, _, _), d = make_workload(name, 0)
    side = min(args.cpu_sample, N0, N1)
    budget_s = 240.0
    while True:
        crop = {k: np.ascontiguousarray(v[:side, :side]) for k, v in d.items()}
        P = orc.ssc_params(side, side, w, DK, DB, True)

        def step():
            t0 = time.time()
            orc.gss(crop['REF'], crop['SCI'], crop['mREF'], crop['mSCI'], P)
            return time.time() - t0
        t_first = step()                                   # untimed probe (also warms the FFT plans)
        if t_first * (K + W) <= budget_s or side <= max(64, 8 * w):
            break
        side //= 2
    for _ in range(W):
        step()
    t0 = time.time()
    for _ in range(K):
        step()
    dt = (time.time() - t0) / K
    v = (side * side / 1e6) / dt
    sample = '%dx%d crop of the %s pair (full size %dx%d: same_config %s), KerHW=%d DK=%d DB=%d, fp64 NumPy port of the reference NumPy backend, %d host threads' % (
        side, side, name, N0, N1, str(side == N0 and side == N1).lower(), w, DK, DB, ncores)
    print(json.dumps({
        'impl': 'reference', 'metric': 'Mpix/s per 4Kx4K SFFT subtraction (GSS: fit + apply)', 'value': v,
        'unit': 'Mpix/s', 'n_gpus': args.gpus, 'steps': K, 'warmup': W,
        'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic', 'config': {'workload': name, 'sample': sample, 'crop': [side, side], 'image': [N0, N1], 'same_config': side == N0 and side == N1},
        'cpu_baseline': {'value': v, 'unit': 'Mpix/s', 'cores': ncores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': v, 'unit': 'Mpix/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument('--cpu-sample', type=int, default=2048, help='side of the crop timed by the CPU legs (halved until K + W steps fit the budget)')
    ap.add_argument('--no-config4', action='store_true', help='skip the shared-template sub-benchmark of the default run')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-pipeline', action='store_true', help='e2e leg: blocking sfftb_gss calls only')
    ap.add_argument('--e2e-steps', type=int, default=0, help='steps of the host-buffer leg (0 = min(steps, 20))')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from sfft_b200 import _lib as B
    from sfft_b200.plan import Plan

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    numa = pin_to_gpu_numa_node(local)          # before any pinned allocation: first touch places the host buffers
    if args.workload.startswith('c3'):
        return run_c3(args, world, rank, local, numa)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    W = max(3, args.warmup)
    K = max(1, args.steps)

    (N0, N1, w, DK, DB, storage), d = make_workload(args.workload, rank)
    npdt = np.float32 if storage == 'fp32' else np.float64
    tdt = torch.float32 if storage == 'fp32' else torch.float64
    code = B.F32 if storage == 'fp32' else B.F64
    host = {k: torch.from_numpy(np.ascontiguousarray(v.astype(npdt))).pin_memory() for k, v in d.items()}
    devt = {k: v.to(dev) for k, v in host.items()}
    diff_d = torch.empty((N0, N1), dtype=tdt, device=dev)
    diff_h = torch.empty((N0, N1), dtype=tdt).pin_memory()
    sol_d = torch.empty(0, dtype=torch.float64, device=dev)

    plan = Plan(N0, N1, w, w, DK, DB, True, device=local, storage=storage)
    sol_d = torch.empty(plan.NEQ, dtype=torch.float64, device=dev)
    sol_h = np.empty(plan.NEQ, np.float64)
    stream = torch.cuda.current_stream(dev)
    plan.bind_torch_stream(stream)
    plan.set_timing(True)
    L = B.lib()

    shared = args.workload.startswith('c4_template')
    bcast_ms = None
    if shared:
        from sfft_b200.batch import TemplateBatch
        tb = TemplateBatch(plan, rank, world)
        torch.cuda.synchronize(dev)
        t0 = time.time()
        if rank == 0:
            tb.set_template(devt['REF'], devt['mREF'])
        else:
            tb.set_template()
        torch.cuda.synchronize(dev)
        bcast_ms = (time.time() - t0) * 1e3

    # shared-template tiles resident in HBM: two tiles in flight on two plans that share the template state
    # (sfftb_gss_template_submit with device pointers; each plan on its own stream), no host round trip between tiles
    dpipe = None
    if shared and not args.no_pipeline:
        from sfft_b200.batch import TemplatePipeline
        dpipe = TemplatePipeline(N0, N1, w, DK, DB, True, device=local, storage=storage, stream_ptr=None, first_plan=plan)
        plan.set_stream(0)                             # every plan on its own stream (see TemplatePipeline)
        dpipe.set_template()
        dpipe.plans[1].set_timing(True)
        d_diff = [diff_d, torch.empty_like(diff_d)]
        d_sol = [sol_d, torch.empty_like(sol_d)]
        d_state = {'k': 0, 'busy': [False, False]}

    def drain_device():
        if ppipe is not None:
            nd = len(ppipe['plans'])
            for d_ in range(nd):
                slot = (ppipe['k'] + d_) % nd
                if ppipe['busy'][slot]:
                    ppipe['plans'][slot].gss_finish()
                    ppipe['busy'][slot] = False
        if dpipe is not None:
            for slot in range(2):
                if d_state['busy'][slot]:
                    dpipe.plans[slot].gss_finish()
                    d_state['busy'][slot] = False

    # device-resident pairs of the plain workloads: PIPE_DEPTH plans, each on its own stream, driven round-robin through
    # sfftb_gss_submit_device / sfftb_gss_finish with the SM partition of sfftb_plan_set_partition -- the latency-bound
    # Cholesky of pair k runs on SOLVER_SMS SMs while the row and column passes of pair k + 1 run on the others.  Every pair
    # is a complete GSS (fit + apply) with its own outputs; results are bit-identical to one pair at a time.
    ppipe = None
    global PIPE_DEPTH, SOLVER_SMS
    big = N0 * N1 > 8192 * 8192 and 'SFFTB_BENCH_DEPTH' not in os.environ
    # (16384^2: the solve is < 10 % of a pair, a plan holds ~25 GB, and two pairs in flight measured slower: 110 vs 88 ms)
    if not shared and not args.no_pipeline and not big:
        plans = [plan] + [Plan(N0, N1, w, w, DK, DB, True, device=local, storage=storage) for _ in range(PIPE_DEPTH - 1)]
        for pl in plans:
            pl.set_stream(0)
            pl.set_partition(SOLVER_SMS)
            pl.set_timing(True)
        ppipe = {'plans': plans, 'busy': [False] * PIPE_DEPTH, 'k': 0,
                 'diff': [diff_d] + [torch.empty_like(diff_d) for _ in range(PIPE_DEPTH - 1)],
                 'sol': [sol_d] + [torch.empty_like(sol_d) for _ in range(PIPE_DEPTH - 1)]}

    def step_device():
        if ppipe is not None:
            slot = ppipe['k'] % len(ppipe['plans'])
            pl = ppipe['plans'][slot]
            if ppipe['busy'][slot]:
                pl.gss_finish()
                for k_, v_ in pl.timings().items():
                    stage[k_] = stage.get(k_, 0.0) + v_
                stage['_n'] = stage.get('_n', 0) + 1
            pl.gss_submit_device(devt['REF'].data_ptr(), devt['SCI'].data_ptr(), devt['mREF'].data_ptr(), devt['mSCI'].data_ptr(),
                                 code, ppipe['sol'][slot].data_ptr(), ppipe['diff'][slot].data_ptr(), code)
            ppipe['busy'][slot] = True
            ppipe['k'] += 1
            return
        if dpipe is not None:
            slot = d_state['k'] % 2
            if d_state['busy'][slot]:
                dpipe.plans[slot].gss_finish()
            dpipe.plans[slot].gss_template_submit_device(devt['SCI'].data_ptr(), devt['mSCI'].data_ptr(), code,
                                                         d_sol[slot].data_ptr(), d_diff[slot].data_ptr(), code)
            d_state['busy'][slot] = True
            d_state['k'] += 1
            return
        if shared:
            plan.gss_template_device(devt['SCI'].data_ptr(), devt['mSCI'].data_ptr(), code, sol_d.data_ptr(),
                                     diff_d.data_ptr(), code)
            return
        plan.gss_device(devt['REF'].data_ptr(), devt['SCI'].data_ptr(), devt['mREF'].data_ptr(), devt['mSCI'].data_ptr(),
                        code, sol_d.data_ptr(), diff_d.data_ptr(), code)

    def step_host():
        if shared:
            B.check(L.sfftb_gss_template(plan._h, host['SCI'].data_ptr(), host['mSCI'].data_ptr(), B.MEM_HOST, code,
                                         sol_h.ctypes.data, B.MEM_HOST, diff_h.data_ptr(), B.MEM_HOST, code))
            return
        B.check(L.sfftb_gss(plan._h, host['REF'].data_ptr(), host['SCI'].data_ptr(), host['mREF'].data_ptr(),
                            host['mSCI'].data_ptr(), B.MEM_HOST, code, sol_h.ctypes.data, B.MEM_HOST,
                            diff_h.data_ptr(), B.MEM_HOST, code))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    stage = {}
    for _ in range(max(W, 4) if (dpipe is not None or ppipe is not None) else W):
        step_device()
    drain_device()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    count_launches = lambda: (plan.launch_count + (dpipe.plans[1].launch_count if dpipe is not None else 0) +
                              (sum(pl.launch_count for pl in ppipe['plans'][1:]) if ppipe is not None else 0))
    l0 = count_launches()
    stage = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(K):
        step_device()
        if ppipe is None:
            for k, v in plan.timings().items():
                stage[k] = stage.get(k, 0.0) + v
            stage['_n'] = stage.get('_n', 0) + 1
    drain_device()                                     # every pair / tile of the timed region has completed
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1) / K
    launches = count_launches() - l0
    nst = max(1, stage.pop('_n', 1))
    stage = {k: v / nst for k, v in stage.items()}
    # Kernel durations for the roofline: with several pairs in flight the per-stage events of a plan also count the time its
    # kernels wait for SMs held by another pair, so the stages are timed once more with ONE pair at a time on the whole GPU
    # (same plan, same buffers, partition off) right after the timed region.
    stage_overlapped, serial_ms = None, None
    if ppipe is not None:
        stage_overlapped = stage
        plan.set_partition(0)
        acc, NS = {}, 5
        torch.cuda.synchronize(dev)
        t0s = time.perf_counter()
        for _ in range(NS):
            plan.gss_device(devt['REF'].data_ptr(), devt['SCI'].data_ptr(), devt['mREF'].data_ptr(), devt['mSCI'].data_ptr(),
                            code, sol_d.data_ptr(), diff_d.data_ptr(), code)
            for k_, v_ in plan.timings().items():
                acc[k_] = acc.get(k_, 0.0) + v_
        torch.cuda.synchronize(dev)
        serial_ms = (time.perf_counter() - t0s) * 1e3 / NS
        stage = {k_: v_ / NS for k_, v_ in acc.items()}
        plan.set_partition(SOLVER_SMS)

    # end-to-end leg: pinned host buffers in, host difference image out, through the public host-buffer API.
    # (a) one blocking sfftb_gss call per step (latency of a single pair);  (b) the same steps through PairPipeline
    # (sfftb_gss_submit / sfftb_gss_finish on two plans sharing the compute stream): every step still copies its four
    # images in and its difference image out inside the timed region, but the copies of step k + 1 overlap step k.
    KE = args.e2e_steps or min(K, 20)
    for _ in range(2):
        step_host()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(stream)
    for _ in range(KE):
        step_host()
    e3.record(stream)
    barrier()
    ms_e2e_single = e2.elapsed_time(e3) / KE
    ms_e2e, e2e_mode = ms_e2e_single, 'one blocking sfftb_gss call per step'
    ms_e2e_full, ms_e2e_delta, delta_bytes = None, None, 0
    if not shared and not args.no_pipeline:
        from sfft_b200.batch import PairPipeline
        E2E_DEPTH = int(os.environ.get('SFFTB_BENCH_E2E_DEPTH', '3'))
        pipe = PairPipeline(N0, N1, w, DK, DB, True, device=local, storage=storage, depth=E2E_DEPTH, solver_sms=SOLVER_SMS)
        diff_hs = [diff_h] + [torch.empty((N0, N1), dtype=tdt).pin_memory() for _ in range(E2E_DEPTH - 1)]
        sol_hs = [torch.empty(plan.NEQ, dtype=torch.float64).pin_memory() for _ in range(E2E_DEPTH)]

        def step_pipe(k):
            pipe.submit(host['REF'], host['SCI'], host['mREF'], host['mSCI'], Solution_out=sol_hs[k % E2E_DEPTH], DIFF_out=diff_hs[k % E2E_DEPTH])
        for k in range(E2E_DEPTH + 1):
            step_pipe(k)
        pipe.drain()
        barrier()
        t0 = time.perf_counter()
        e2.record(stream)
        for k in range(KE):
            step_pipe(k)
        pipe.drain()                                   # the last difference images are in host memory when this returns
        e3.record(stream)
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3 / KE
        # the copies run on the plans' copy streams, so the honest clock is the host's (submit of the first step ->
        # last result in host memory); the events on the compute stream are kept as a cross-check
        ms_e2e_full = max(wall_ms, e2.elapsed_time(e3) / KE)
        ms_e2e, e2e_mode = ms_e2e_full, 'PairPipeline: sfftb_gss_submit/finish on %d plans (own streams, SM partition), copies of step k+1 under step k' % E2E_DEPTH
        # the same steps with the masked pair sent as sparse deltas against the unmasked pair (sfftb_gss_submit_delta): the
        # masked images differ from the unmasked ones only inside the masked stamps, so two images instead of four cross
        # the link; the deltas are part of the step's input and are copied inside the timed region like the images
        from sfft_b200.batch import sparse_delta
        dI = sparse_delta(host['REF'].numpy(), host['mREF'].numpy())
        dJ = sparse_delta(host['SCI'].numpy(), host['mSCI'].numpy())
        dI = tuple(torch.from_numpy(a).pin_memory().numpy() for a in dI)
        dJ = tuple(torch.from_numpy(a).pin_memory().numpy() for a in dJ)
        delta_bytes = sum(a.nbytes for a in dI + dJ)

        def step_delta(k):
            pipe.submit_delta(host['REF'], host['SCI'], dI, dJ, Solution_out=sol_hs[k % E2E_DEPTH], DIFF_out=diff_hs[k % E2E_DEPTH])
        for k in range(E2E_DEPTH + 1):
            step_delta(k)
        pipe.drain()
        barrier()
        t0 = time.perf_counter()
        e2.record(stream)
        for k in range(KE):
            step_delta(k)
        pipe.drain()
        e3.record(stream)
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3 / KE
        ms_e2e_delta = max(wall_ms, e2.elapsed_time(e3) / KE)
        pipe.close()
    if shared and not args.no_pipeline:
        from sfft_b200.batch import TemplatePipeline
        tp = dpipe                                     # the two plans of the device leg (template state already cloned)
        diff_hs = [diff_h, torch.empty((N0, N1), dtype=tdt).pin_memory()]
        sol_hs = [torch.empty(plan.NEQ, dtype=torch.float64).pin_memory() for _ in range(2)]

        def step_tpipe(k):
            tp.submit(host['SCI'], host['mSCI'], Solution_out=sol_hs[k % 2], DIFF_out=diff_hs[k % 2])
        for k in range(4):                              # includes the first (factorising) tile of the second plan
            step_tpipe(k)
        tp.drain()
        barrier()
        t0 = time.perf_counter()
        e2.record(stream)
        for k in range(KE):
            step_tpipe(k)
        tp.drain()
        e3.record(stream)
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3 / KE
        ms_e2e, e2e_mode = max(wall_ms, e2.elapsed_time(e3) / KE), 'TemplatePipeline: sfftb_gss_template_submit/finish on two plans sharing the template state'
        tp.close()
        dpipe = None
    clocks = sampler.stop()
    esz = 4 if storage == 'fp32' else 8
    probe = link_probe(torch, dist, dev, world, N0 * N1 * esz)
    c4 = None
    if args.workload == DEFAULT_WORKLOAD and not args.no_config4:
        c4 = run_config4(torch, dist, B, world, rank, local, dev)

    t = torch.tensor([ms, ms_e2e, ms_e2e_single, ms_e2e_delta or 0.0, ms_e2e_full or 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e, ms_e2e_single = float(t[0]), float(t[1]), float(t[2])
    ms_e2e_delta = float(t[3]) if ms_e2e_delta else None
    ms_e2e_full = float(t[4]) if ms_e2e_full else None
    mpix = N0 * N1 / 1e6
    value = world * mpix / (ms / 1e3)
    e2e_value = world * mpix / (ms_e2e / 1e3)

    if rank == 0:
        csz = 2 * esz
        NH = N1 // 2 + 1
        Fij = (DK + 1) * (DK + 2) // 2
        peak, which = hbm_peak()
        # dominant kernel: the fit column pass.  Bytes it must move in this layout: read the (DK+1)+1 stored
        # row-spectrum planes once, write the lag partials.
        npairs = Fij * (Fij + 1) // 2
        Fpq = (DB + 1) * (DB + 2) // 2
        nrowsK = npairs * (4 * w + 1) + Fij * (2 * w + 1)
        seg_path = 4 * w + 32 <= 256      # fit_seg3_kernel (one launch for DK <= 2, three plane-range launches for DK = 3)
        nrowsL = (Fij * Fpq * (2 * w + 1) + Fpq) if seg_path else (Fij * (DB + 1) * (2 * w + 1) + (DB + 1))
        kname = 'fit_seg4_kernel' if seg_path else 'fit_col_kernel'
        npass = 3 if (seg_path and DK == 3) else 1          # every plane-range launch streams the stored planes once
        alg_bytes = (DK + 2) * NH * N0 * csz + (nrowsK + nrowsL) * NH * 16 / npass       # per launch
        t_kernel = (stage.get('fit_cols_kernel') or stage.get('fit_cols', 0.0)) / 1e3 / npass      # average launch duration of the kernel
        achieved = alg_bytes / t_kernel / 1e9 if t_kernel > 0 else None      # = bytes per launch / average launch time
        # whole-step algorithmic bytes: (4 n_pl + 7) * N0 * N1 * s with n_pl = Fij + 1
        step_bytes = (4 * (Fij + 1) + 7) * N0 * N1 * esz
        # end to end.  Headline: the pipelined host-buffer API with the masked pair sent as sparse deltas against the unmasked
        # pair (sfftb_gss_submit_delta; bit-identical results, tests/test_gpu_parity.py::test_pair_pipeline_submit_finish) --
        # the masked images of the packets ARE the unmasked ones with stamps replaced, so this is the input a caller has.
        # The four-full-images form of the same step and the link ceiling measured in this run are reported beside it.
        h2d_full = (2 if shared else 4) * N0 * N1 * esz
        d2h = N0 * N1 * esz + plan.NEQ * 8
        four = {'value': e2e_value, 'unit': 'Mpix/s', 'ms_per_step': ms_e2e, 'mode': e2e_mode,
                'h2d_bytes_per_step': h2d_full, 'd2h_bytes_per_step': d2h}
        if probe:
            four['link_bound_ms_per_step'] = max(h2d_full / (probe['h2d_gbs_per_gpu'] * 1e6), d2h / (probe['d2h_gbs_per_gpu'] * 1e6))
            four['frac_of_link_bound'] = four['link_bound_ms_per_step'] / ms_e2e
        e2e_obj = dict(four)
        if ms_e2e_delta:
            h2d_delta = 2 * N0 * N1 * esz + delta_bytes
            e2e_obj = {'value': world * mpix / (ms_e2e_delta / 1e3), 'unit': 'Mpix/s', 'ms_per_step': ms_e2e_delta,
                       'mode': 'PairPipeline.submit_delta: sfftb_gss_submit_delta / sfftb_gss_finish on %d plans, own streams, SM partition (I, J + sparse deltas of ' % E2E_DEPTH +
                               'mI, mJ in, DIFF + Solution out; copies of step k+1 under step k)',
                       'h2d_bytes_per_step': h2d_delta, 'd2h_bytes_per_step': d2h}
            if probe:
                b2 = max(h2d_delta / (probe['h2d_gbs_per_gpu'] * 1e6), d2h / (probe['d2h_gbs_per_gpu'] * 1e6))
                e2e_obj['link_bound_ms_per_step'] = b2
                e2e_obj['frac_of_link_bound'] = b2 / ms_e2e_delta
                # both directions are busy in a pipeline: while the difference image goes home each way only gets the duplex
                # rate the probe measured (at N = 8 the host side of this box gives 8 GB/s each way per GPU against 23 / 12
                # one way), the rest of the input then moves at the one-way rate
                dup = probe['duplex_each_way_gbs_per_gpu'] * 1e6
                small, big_ = min(h2d_delta, d2h), max(h2d_delta, d2h)
                one_way = (probe['h2d_gbs_per_gpu'] if h2d_delta >= d2h else probe['d2h_gbs_per_gpu']) * 1e6
                b3 = small / dup + (big_ - small) / one_way
                e2e_obj['duplex_link_bound_ms_per_step'] = b3
                e2e_obj['frac_of_duplex_link_bound'] = b3 / ms_e2e_delta
            e2e_obj['four_image_form'] = four
        e2e_obj.update({'steps': KE, 'single_call_ms': ms_e2e_single, 'single_call_value': world * mpix / (ms_e2e_single / 1e3),
                        'host_numa_node': numa})
        if probe:
            e2e_obj['link_probe'] = probe
        out = {
            'metric': 'Mpix/s per 4Kx4K SFFT subtraction (GSS: fit + apply)', 'value': value, 'unit': 'Mpix/s',
            'n_gpus': world, 'steps': K, 'warmup': W, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': args.workload, 'image': [N0, N1], 'KerHW': w, 'KerPolyOrder': DK, 'BGPolyOrder': DB,
                       'storage': storage, 'arithmetic': 'fp64', 'pairs_per_step_per_gpu': 1,
                       'l2_policy': 'working set per step (inputs 4x%.0f MB + spectra %.0f MB) exceeds the 126 MB L2' % (
                           N0 * N1 * esz / 1e6, (DK + 2) * NH * N0 * csz / 1e6),
                       'fold': plan.dims['fold'], 'sub_len': plan.dims['sub_len'],
                       'shared_template': shared, 'template_prepare_broadcast_ms': bcast_ms,
                       'tiles_in_flight': 2 if (shared and not args.no_pipeline) else 1,
                       'pairs_in_flight': PIPE_DEPTH if ppipe is not None else 1,
                       'solver_partition_sms': SOLVER_SMS if ppipe is not None else 0,
                       'device_leg': ('%d plans, each on its own stream, sfftb_gss_submit_device / sfftb_gss_finish round-robin; SM '
                                      'partition: Cholesky on %d SMs, row / column passes on the others, so the solve of pair k runs '
                                      'beside the transforms of pair k+1 (stage_ms are per-pair event times under that overlap)' % (
                                          PIPE_DEPTH, SOLVER_SMS)) if ppipe is not None else 'one blocking call per step'},
            'stage_ms': stage,
            'stage_ms_note': ('one pair at a time on the whole GPU (5 blocking calls after the timed region): kernel durations; '
                              'the timed region itself keeps %d pairs in flight' % PIPE_DEPTH) if stage_overlapped is not None else 'timed region',
            'stage_ms_pairs_in_flight': stage_overlapped,
            'one_pair_at_a_time_ms': serial_ms,
            'assembly_solve_ms': stage.get('fit_cols', 0) + stage.get('fit_reduce_fill', 0) + stage.get('fit_solve', 0),
            'solver': plan.last_solver,
            'clocks': clocks,
            'e2e': e2e_obj,
            'gpu_launches': launches,
            'roofline': {'bound': 'hbm', 'kernel': kname, 'achieved': achieved, 'peak': peak,
                         'note': 'the path is fp64-issue bound on B200 (64 DFMA/clk/SM, DESIGN.md section 4); the HBM '
                                 'fraction is reported as the contract asks',
                         'peak_source': which, 'unit': 'GB/s', 'frac': (achieved / peak) if achieved else None,
                         'traffic': None, 'algorithmic_bytes_per_launch': alg_bytes, 'launches_per_step_of_kernel': npass,
                         'kernel_ms': t_kernel * 1e3,
                         'step_algorithmic_bytes': step_bytes,
                         'step_frac': step_bytes / (ms / 1e3) / 1e9 / peak},
        }
        if c4 is not None:
            out['config4'] = c4
        tr = os.path.join(ROOT, 'profiles', 'traffic.json')
        if os.path.exists(tr):
            try:
                trd = json.load(open(tr)).get(args.workload, {})
                out['roofline']['traffic'] = trd.get(kname)
                if trd.get('limiters', {}).get(kname):
                    # what actually limits the kernel (from the committed ncu capture, not measured in this run)
                    out['roofline']['ncu_limiters'] = trd['limiters'][kname]
            except Exception:
                pass
        if world == 1 and not args.no_cpu_baseline:
            v, dt, s, cores = cpu_port_mpix(args.workload, args.cpu_sample)
            out['cpu_baseline'] = {'value': v, 'unit': 'Mpix/s', 'cores': cores, 'kind': 'port', 'seconds': dt,
                                   'sample': '%dx%d crop of the same pair, same KerHW/orders, fp64 NumPy port of the '
                                             'reference NumPy backend (oracle/sfft_oracle.py), one GSS' % (s, s)}
        else:
            out['cpu_baseline'] = None
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
