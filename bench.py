#!/usr/bin/env python
"""Benchmark of the shifted-correlation hot path (BASELINE.json metric: atom*origin*lag evals/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (every N): BASELINE config 4, the north-star case -- ISF at q = 22.7 nm^-1 plus the chi4 susceptibility
through ``correlation.chi4`` (``shifted_correlation(isf, average=False, segments=20, window=0.2, skip=0.1)`` and
``N * var(axis=0)``, examples/plot_chi4.py:29-37) on 1 000 000 atoms x 20 000 frames of a periodic random walk with
nojump unwrapping: 20 origins x 81 log-spaced lags = 1.62e9 evaluations per step.  STRONG scaling: the same
trajectory is sharded by contiguous atom blocks over the N GPUs (one process per GPU); the per-(origin, lag) partial
sums are added over the GPUs by the finish kernel of every call over NVLink peer memory (no NCCL call per step).
The 240 GB trajectory is generated on the device segment by segment and streamed through the jump scan once; only
the frames the schedules touch stay resident (frames + 1-byte / packed 2-bit image counts).

One step = one ``chi4`` call.  ``value``: rows resident in HBM.  ``e2e``: the frames the schedule touches sit in
page-locked HOST memory next to the trajectory's jump matrices (what a reader with a frame index and the
``.nojump.npz`` cache hands over); every step uploads them, rebuilds the image counts, correlates and reads the
result back.  ``roofline``: the correlation kernel alone (CUDA events around it on its stream, inside the timed
region), per GPU; ``roofline.through_api`` the same bytes over ``ms_per_step``; ``roofline.sustained`` a >= 2 s
back-to-back loop.  ``parity``: rank 0 re-computes a sub-block of its shard with the CPU oracle (every N) and, for
N > 1, checks the peer-memory sums against a torch.distributed all-reduce of the per-rank partials (integer counts:
exactly).  ``cpu_baseline`` (N = 1): the UNMODIFIED reference (baseline/_ref, stub-imported) on an atom sub-block
of the same trajectory.  ``config2`` (N = 1): the round-1 workload (ISF, 100 000 atoms x 10 000 frames, 100 origins)
with the other kernels of the path, for continuity.

``--impl reference`` times the reference alone (pure Python/NumPy, single-threaded: 1 core) on the same kind of
sub-block; its ``value`` is the correlation with the jump matrices prepared, like ``value`` of our arm.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from functools import partial

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

_OUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout, but libraries write there too (NCCL prints its version banner
    to file descriptor 1).  Keep the real stdout for the result line and point fd 1 at stderr for everyone else."""
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)


def emit(line):
    out = _OUT if _OUT is not None else sys.stdout
    out.write(json.dumps(line) + '\n')
    out.flush()


Q = 22.7
METRIC = 'atom*origin*lag evals/s (ISF + chi4 via shifted_correlation, nojump unwrap)'
UNIT = 'evals/s'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--atoms', type=int, default=1000000, help='atoms in total (sharded over the GPUs)')
    ap.add_argument('--frames', type=int, default=20000)
    ap.add_argument('--segments', type=int, default=20)
    ap.add_argument('--window', type=float, default=0.2)
    ap.add_argument('--skip', type=float, default=0.1)
    ap.add_argument('--block-atoms', type=int, default=125000, help='atoms per generator block (one seed each)')
    ap.add_argument('--parity-atoms', type=int, default=4096, help='atom sub-block checked against the oracle')
    ap.add_argument('--cpu-atoms', type=int, default=0, help='atom sub-block timed on the CPU (0 = from the budget)')
    ap.add_argument('--cpu-seconds', type=float, default=120.0, help='time budget of the CPU arm')
    ap.add_argument('--e2e-steps', type=int, default=3)
    ap.add_argument('--sustain-seconds', type=float, default=2.5)
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-config2', action='store_true')
    ap.add_argument('--no-sustained', action='store_true')
    return ap.parse_args()


def config_of(a):
    """The workload description both arms print (identical, so that the driver can match them)."""
    from mdevaluate_b200.correlation import log_indices
    starts = np.unique(np.linspace(a.frames * a.skip, a.frames * (1 - a.window), num=a.segments, endpoint=False, dtype=int))
    S, L = len(starts), len(log_indices(0, int(a.frames * a.window)))
    per_gpu = a.atoms // max(1, a.gpus)
    return {'workload': 'config4: ISF q=%.1f + chi4, %d atoms x %d frames, nojump, segments=%d, window=%g, skip=%g, '
                        'log lags' % (Q, a.atoms, a.frames, a.segments, a.window, a.skip),
            'atoms_total': a.atoms, 'frames': a.frames, 'segments': a.segments, 'window': a.window, 'skip': a.skip,
            'q': Q, 'origins': int(S), 'lags': int(L), 'evals_per_step': int(a.atoms) * int(S) * int(L),
            'parallelism': 'atom-shard x%d (strong scaling: %d atoms per GPU)' % (a.gpus, per_gpu),
            'l2': 'inputs larger than L2: each step streams %.1f GB of frame rows per GPU'
                  % (12.0 * per_gpu * S * (L + 1) / 1e9)}


def cpu_sample_atoms(a):
    """Atoms of the sub-block the CPU arm runs: sized so that warm-up + timed steps take about ``cpu_seconds``
    (measured here: a reference step costs ~1.6 s + 2.3e-4 s per atom on this schedule), the same in both arms."""
    if a.cpu_atoms:
        return min(a.cpu_atoms, a.atoms)
    per_step = a.cpu_seconds / max(1, a.steps + a.warmup)
    n = int((per_step - 1.6) / 2.3e-4) // 1024 * 1024
    return int(min(max(n, 2048), 16384, a.atoms))


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.FIELDS, '--format=csv,noheader,nounits',
                 '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def window(self, t0, t1):
        sm, mx, reasons, power = [], [], set(), []
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        for ts, line in list(self.lines):
            if not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons),
                'power_w_max': float(max(power)), 'samples': len(sm)}

    def stop(self):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()


def measured_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as fh:
            return float(json.load(fh)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


def stored_traffic(key):
    """DRAM bytes per launch of the dominant kernel from a committed ``ncu --set full`` capture of exactly this
    kernel and shape (profiles/traffic.json names the capture); None when there is none -- it is not measured in
    this run (ncu cannot run inside a timed bench)."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as fh:
            entry = json.load(fh).get(key)
        return (entry['bytes'], entry.get('source')) if entry else (None, None)
    except Exception:
        return None, None


# ------------------------------------------------------------------------------------------------
# the reference on the CPU
# ------------------------------------------------------------------------------------------------
def host_random_walk(n_atoms, n_frames, box, seed):
    """The synthetic walk generated on the host (torch CPU ops) for the stand-alone reference arm."""
    from mdevaluate_b200.synthetic import RandomWalkStream
    stream = RandomWalkStream(n_frames, box, 0, 1, n_atoms, seed=seed, device='cpu', frame_block=256,
                              capture=(0, n_atoms))
    for _ in stream.device_segments():
        pass
    return stream.captured, stream.box_diag, stream.times


def reference_timer(pos, box, times, a):
    """``(step, info)``: ``step()`` runs one ``shifted_correlation(isf)`` of the reference (or, where baseline/_ref is
    absent, of its NumPy port) on the host arrays and returns ``(result, seconds)``; the jump matrices are prepared
    once, outside (their cost is in ``info``)."""
    from baseline import reference as ref
    kw = dict(segments=a.segments, window=a.window, skip=a.skip)
    if ref.available():
        md = ref.load()
        coords, jump_s = ref.coordinates(md, pos, box, times)
        frames = coords.nojump
        fn = partial(md.correlation.isf, q=Q)

        def step():
            t0 = time.perf_counter()
            _, r = md.correlation.shifted_correlation(fn, frames, **kw)
            return np.asarray(r), time.perf_counter() - t0
        return step, {'kind': 'reference', 'parse_jumps_s': jump_s,
                      'what': 'mdevaluate (baseline/_ref, unmodified, stub-imported): shifted_correlation(isf) on '
                              'coords.nojump with reader.parse_jumps matrices prepared beforehand'}
    from oracle import md_oracle as orc
    t0 = time.perf_counter()
    unw = orc.UnwrappedTrajectory(pos, box, times)
    jump_s = time.perf_counter() - t0

    def step():
        t0 = time.perf_counter()
        _, r = orc.shifted_correlation(partial(orc.isf, q=Q), unw, **kw)
        return np.asarray(r), time.perf_counter() - t0
    return step, {'kind': 'port', 'parse_jumps_s': jump_s,
                  'what': 'oracle/md_oracle.py (NumPy port; baseline/_ref is absent)'}


def reference_arm(a):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from mdevaluate_b200.synthetic import box_length
    n = cpu_sample_atoms(a)
    pos, box, times = host_random_walk(n, a.frames, box_length(a.atoms), seed=20240)
    step, info = reference_timer(pos, box, times, a)
    for _ in range(a.warmup):
        step()
    wall = 0.0
    shape = None
    for _ in range(a.steps):
        r, s = step()
        wall += s
        shape = r.shape
    evals = n * shape[0] * shape[1]
    value = evals * a.steps / wall
    sample = ('%d-atom sub-block of the workload (all %d frames, S=%d, L=%d) per step; %s; parse_jumps (one-off, not '
              'timed) %.1f s' % (n, a.frames, shape[0], shape[1], info['what'], info['parse_jumps_s']))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': a.gpus, 'steps': a.steps,
        'warmup': a.warmup, 'ms_per_step': 1e3 * wall / a.steps, 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': config_of(a),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': 1, 'kind': info['kind'], 'sample': sample,
                         'host_cores': os.cpu_count()},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(torch, dev):
    """Run this process (and allocate its page-locked buffers) on the NUMA node the GPU hangs off: the host->device
    copies of 8 ranks otherwise all cross one memory controller (round-1 SCALE: end-to-end efficiency 0.42)."""
    try:
        props = torch.cuda.get_device_properties(dev)
        bus = '%04x:%02x:%02x.0' % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        base = '/sys/bus/pci/devices/%s' % bus
        with open(base + '/local_cpulist') as fh:
            cpulist = fh.read().strip()
        with open(base + '/numa_node') as fh:
            node = int(fh.read().strip())
        cpus = set()
        for part in cpulist.split(','):
            lo, _, hi = part.partition('-')
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        use = (cpus & allowed) or allowed
        os.sched_setaffinity(0, use)
        return {'numa_node': node, 'cpus': len(use), 'pci': bus}
    except Exception as exc:
        return {'numa_node': None, 'error': '%s: %s' % (type(exc).__name__, exc)}


def ours_arm(a):
    import torch
    import torch.distributed as dist

    import mdevaluate_b200 as md
    from mdevaluate_b200 import parallel
    from mdevaluate_b200._cuda import engine
    from mdevaluate_b200.synthetic import RandomWalkStream, box_length
    from oracle import md_oracle as orc

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py --impl ours needs a CUDA device (there is no CPU fallback)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    numa = bind_to_gpu_numa_node(torch, dev)
    if world > 1:
        import datetime
        # a stuck collective must fail loudly, not hang the run
        dist.init_process_group('nccl', device_id=dev, timeout=datetime.timedelta(seconds=300))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    c = md.correlation
    N, T = a.atoms, a.frames
    n_blocks = max(1, N // a.block_atoms)
    if N % a.block_atoms or n_blocks % world:
        raise SystemExit('--atoms must be a multiple of --block-atoms and the blocks must divide over the GPUs')
    per_rank = n_blocks // world
    n_local = per_rank * a.block_atoms
    n_par = min(a.parity_atoms, n_local) // 16 * 16
    n_cpu = cpu_sample_atoms(a) if (rank == 0 and world == 1 and not a.no_cpu) else 0
    capture = max(n_par, n_cpu) if rank == 0 else 0
    stream = RandomWalkStream(T, box_length(N), rank * per_rank, per_rank, a.block_atoms, seed=20240, device=dev,
                              capture=(0, capture) if capture else None)
    box, times = stream.box_diag, stream.times
    coords = md.Coordinates(stream)
    if world > 1:
        parallel.shard_atoms(coords, n_atoms_global=N)
    frames = coords.nojump
    kw = dict(segments=a.segments, window=a.window, skip=a.skip)

    # ---- one pass over the trajectory: jump scan, the frames of the schedules are kept --------------------
    res = coords._resident()
    res.force_compact = True
    starts, idx = c._schedule(frames, c.log_indices, a.segments, a.window, a.skip)
    starts_b, idx_b = c._schedule(frames, c.log_indices, 10, 0.5, None)          # S=10 / L=84, the default schedule
    ids = np.unique(np.concatenate([(starts[:, None] + idx[None, :]).ravel(), (starts_b[:, None] + idx_b[None, :]).ravel()]))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    t_build = time.time()
    store = res.store_for(ids, nojump=True)
    torch.cuda.synchronize()
    t_build = time.time() - t_build
    S, L = len(starts), len(idx)
    evals_step = N * S * L
    peak, peak_src = measured_peak()
    alg_bytes = 12.0 * n_local * S * (L + 1)            # per GPU and launch
    fn = partial(c.isf, q=Q)

    def step():
        return c.chi4(frames, Q, **kw)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()               # nvidia-smi needs ~0.3 s to deliver its first sample: started before the warm-up
    t_axis, x4 = step()
    time.sleep(0.3)
    for i in range(a.warmup):         # W untimed steps, immediately before the timed region (no idle gap: the
        engine.time_slot = i % 1024   # clocks of an idle GPU drop within a few hundred ms), kernel events on as in
        step()                        # the timed steps

    # ---- timed region: K steps, rows resident in HBM -----------------------------------------------------
    launches0 = engine.launches()
    barrier()
    w0 = time.time()
    ev[0].record()
    for i in range(a.steps):
        engine.time_slot = i % 1024
        t_axis, x4 = step()
    ev[1].record()
    barrier()
    w1 = time.time()
    engine.time_slot = -1
    elapsed_ms = max_over_ranks(ev[0].elapsed_time(ev[1]))
    launches = engine.launches() - launches0
    plan = engine.last_plan()
    kernel_ms = [plan.kernel_ms(i % 1024) for i in range(max(0, a.steps - 1024), a.steps)]
    k_ms = max_over_ranks(float(np.mean(kernel_ms)))
    ms_per_step = elapsed_ms / a.steps
    value = evals_step * a.steps / (elapsed_ms * 1e-3)
    clocks = sampler.window(w0, w1) if rank == 0 else None
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    moved = (13.0 if store.packed is not None else 15.0) * n_local * S * (L + 1)
    traffic, traffic_src = stored_traffic('isf:%d:%d:%d' % (n_local, S, L))
    roofline = {
        'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
        'traffic': traffic, 'traffic_source': traffic_src,
        'kernel': 'corr_bulk_kernel<%s image counts, SINC, constant box> (float32 displacements)'
                  % ('packed' if store.packed is not None else 'byte'),
        'kernel_ms': k_ms, 'kernel_ms_min': float(np.min(kernel_ms)), 'kernel_ms_median': float(np.median(kernel_ms)),
        'kernel_ms_max': float(np.max(kernel_ms)), 'per_gpu': True,
        'algorithmic_bytes_per_launch': alg_bytes, 'moved_bytes_per_launch': moved, 'peak_source': peak_src,
        'through_api': {'ms_per_step': ms_per_step, 'frac': alg_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                        'host_overhead_ms': ms_per_step - k_ms},
        'burst_or_sustained': 'burst: %d back-to-back steps, %.3f s' % (a.steps, elapsed_ms * 1e-3),
        'note': 'algorithmic bytes = 12 B (float32 xyz) per atom*origin*(lag+1) of this GPU\'s shard; the kernel also '
                'reads the image counts (packed: 1 B per atom and frame, 13 B moved per 12 B algorithmic)'}

    # ---- sustained: >= sustain_seconds back to back, clocks alongside ---------------------------------------
    if not a.no_sustained:
        n_sus = int(min(20000, max(a.steps, a.sustain_seconds / (ms_per_step * 1e-3))))
        barrier()
        s0 = time.time()
        ev[0].record()
        for i in range(n_sus):
            engine.time_slot = i % 1024
            step()
        ev[1].record()
        barrier()
        s1 = time.time()
        engine.time_slot = -1
        sus_ms = max_over_ranks(ev[0].elapsed_time(ev[1])) / n_sus
        tail = [plan.kernel_ms(i % 1024) for i in range(max(0, n_sus - 1024), n_sus)]
        sus_k = max_over_ranks(float(np.mean(tail)))
        roofline['sustained'] = {'steps': n_sus, 'seconds': (s1 - s0), 'ms_per_step': sus_ms, 'kernel_ms': sus_k,
                                 'frac': alg_bytes / (sus_k * 1e-3) / 1e9 / peak,
                                 'through_api_frac': alg_bytes / (sus_ms * 1e-3) / 1e9 / peak,
                                 'value': evals_step / (sus_ms * 1e-3),
                                 'clocks': sampler.window(s0, s1) if rank == 0 else None}

    # ---- where a step's time beyond its kernel goes: the plan's timeline (host clock + the GPU's %globaltimer) ----
    try:
        pairs = [engine.clock_pair() for _ in range(30)]
        g_ns, before, seen = min(pairs, key=lambda p_: p_[2] - p_[1])
        offset = g_ns - (before + seen) / 2.0
        rows_tl = []
        for _ in range(30):
            t0 = time.perf_counter_ns()
            step()
            t1 = time.perf_counter_ns()
            h0, h1, h2, g0, g1, g2 = plan.timeline()
            rows_tl.append((t1 - t0, h2 - h0, h1 - h0, g0 - offset - h0, g1 - g0, g2 - g1, h2 - (g2 - offset)))
        med = np.median(np.array(rows_tl, dtype=np.float64), axis=0) / 1e3
        roofline['through_api']['timeline_us'] = {
            'step (chi4 call, host clock)': med[0], 'mdb_corr_plan_run': med[1], 'launch call': med[2],
            'launch entered -> first CTA': med[3], 'correlation kernel start -> finish kernel start': med[4],
            'finish kernel': med[5], 'flag store -> host saw it': med[6], 'python around the run': med[0] - med[1],
            'step beyond the correlation kernel': med[0] - med[4],
            'clock_offset_uncertainty': (seen - before) / 2e3,
            'note': 'medians of 30 steps WITHOUT kernel events, taken after the sustained leg (power-capped clocks); '
                    'GPU and host clocks related by mdb_clock_pair.  host_overhead_ms above is ms_per_step minus the '
                    'event-timed kernel and so includes what the two cudaEventRecords per step cost'}
    except Exception as exc:
        roofline['through_api']['timeline_us'] = {'error': '%s: %s' % (type(exc).__name__, exc)}

    # ---- the second schedule of config 4 (S=10, L=84) and the ISF through the float64 chain ------------------
    def rate(fn_, kw_, reps=5, **tune):
        for k_, v_ in tune.items():
            setattr(engine.tuning, k_, v_)
        try:
            engine.time_slot = 0
            ms, out = [], None
            for _ in range(reps):
                out = c.shifted_correlation(fn_, frames, **kw_)[1]
                ms.append(engine.last_plan().kernel_ms(0))
        finally:
            engine.time_slot = -1
            for k_ in tune:
                setattr(engine.tuning, k_, 0)
        return max_over_ranks(float(min(ms))), out
    other = {}
    k_b, _ = rate(fn, dict(segments=10, window=0.5))
    Sb, Lb = len(starts_b), len(idx_b)
    other['isf_S10_L84'] = {'kernel_ms': k_b, 'evals_per_s': N * Sb * Lb / (k_b * 1e-3),
                            'frac': 12.0 * n_local * Sb * (Lb + 1) / (k_b * 1e-3) / 1e9 / peak}
    k64, f64 = rate(fn, kw, isf_float64=1)
    _, f32 = rate(fn, kw, reps=1)
    other['isf_f64'] = {'kernel_ms': k64, 'frac': alg_bytes / (k64 * 1e-3) / 1e9 / peak,
                        'max_abs_diff_to_f32_path': float(np.abs(f64 - f32).max()),
                        'note': 'the same call with displacements through the float64 chain (tuning.isf_float64)'}

    # ---- parity -------------------------------------------------------------------------------------------
    parity = None
    if rank == 0 and n_par:
        sub = stream.captured[:, :n_par]
        unw = orc.UnwrappedTrajectory(sub, box, times)
        _, ref = orc.shifted_correlation(partial(orc.isf, q=Q), unw, **kw)
        ref_x4 = n_par * ref.var(axis=0)
        rows = store.rows_for(starts[:, None] + idx[None, :])
        spec = c._kernel_spec(fn)
        sub_plan = engine.CorrPlan(store.atom_block(0, n_par), rows, spec.corr)
        sums, _ = sub_plan.run()
        ours = spec.finalize(sums, None, n_par, False)
        sub_plan.close()
        err = np.abs(ours - ref)
        tol = 1e-5 * np.abs(ref) + 1e-7
        counts_ok = bool(np.array_equal(store.atom_block(0, n_par).image_counts(rows[0, :4]),
                                        unw.counts[(starts[0] + idx[:4])]))
        parity = {'sub_block_atoms': n_par, 'max_abs_err': float(err.max()),
                  'max_rel_err': float((err / np.maximum(np.abs(ref), 1e-300)).max()),
                  'within_tolerance': bool((err <= tol).all()), 'tolerance': '1e-5*|ref| + 1e-7',
                  'chi4_max_abs_err': float(np.abs(n_par * ours.var(axis=0) - ref_x4).max()),
                  'image_counts_bit_exact': counts_ok,
                  'oracle': 'oracle/md_oracle.py on the first %d atoms of rank 0\'s shard, all %d frames' % (n_par, T)}
    if world > 1:
        # the peer-memory sums against torch.distributed: sums to rounding, integer counts exactly
        bins = np.linspace(0, 3.0, 501)
        rows = store.rows_for(starts[:, None] + idx[None, :])
        checks = {}
        for name, f in (('isf', fn), ('van_hove', partial(c.van_hove_self, bins=bins))):
            spec = c._kernel_spec(f)
            local_plan = engine.CorrPlan(store, rows, spec.corr)
            s_loc, h_loc = local_plan.run()
            s_ref = torch.from_numpy(np.array(s_loc)).to(dev)
            dist.all_reduce(s_ref)
            peer_plan = engine.CorrPlan(store, rows, spec.corr, comm=coords.comm.p2p)
            s_peer, h_peer = peer_plan.run()
            denom = max(1e-300, float(s_ref.abs().max()))
            checks[name + '_sums_max_rel_diff'] = float((torch.from_numpy(np.array(s_peer)).to(dev) - s_ref).abs().max()) / denom
            if h_loc is not None:
                h_ref = torch.from_numpy(np.array(h_loc)).to(dev)
                dist.all_reduce(h_ref)
                checks[name + '_counts_equal'] = bool(torch.equal(torch.from_numpy(np.array(h_peer)).to(dev), h_ref))
                checks[name + '_counts_total'] = int(h_ref.sum().item())
            local_plan.close()
            peer_plan.close()
        ok = torch.tensor([1 if (checks['isf_sums_max_rel_diff'] < 1e-12 and checks['van_hove_counts_equal']) else 0],
                          device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        checks['all_ranks_agree'] = bool(int(ok.item()))
        checks['peer_memory_comm'] = coords.comm.p2p is not None
        if parity is not None:
            parity['multi_gpu'] = checks
            parity['within_tolerance'] = bool(parity['within_tolerance'] and checks['all_ranks_agree'])

    # ---- CPU baseline: the reference on a sub-block (rank 0, N = 1) -------------------------------------------
    cpu_baseline = None
    if n_cpu:
        cpu_step, info = reference_timer(np.ascontiguousarray(stream.captured[:, :n_cpu]), box, times, a)
        cpu_step()
        r_cpu, s_cpu = cpu_step()
        cpu_baseline = {'value': n_cpu * S * L / s_cpu, 'unit': UNIT, 'cores': 1, 'kind': info['kind'],
                        'host_cores': os.cpu_count(),
                        'sample': '%d-atom sub-block of the workload (all %d frames, S=%d, L=%d), one timed step after '
                                  'one warm-up; %s; parse_jumps (one-off, not timed) %.1f s'
                                  % (n_cpu, T, S, L, info['what'], info['parse_jumps_s'])}

    # ---- end to end ---------------------------------------------------------------------------------------------
    e2e = None
    if not a.no_e2e:
        try:
            e2e = measure_e2e(a, torch, md, c, parallel, res, store, stream, starts, idx, kw, world, N, n_local, barrier,
                              max_over_ranks, ev, evals_step, dev)
        except Exception as exc:      # e.g. not enough page-locked host memory: report it, keep the line
            e2e = {'value': None, 'unit': UNIT, 'h2d_bytes_per_step': None, 'd2h_bytes_per_step': int(S * L * 3 * 8),
                   'error': '%s: %s' % (type(exc).__name__, exc)}
            if world > 1:
                raise
    sampler.stop()

    config2 = None
    if world == 1 and not a.no_config2:
        del store
        res._compact.clear()
        torch.cuda.empty_cache()
        config2 = config2_leg(a, torch, md, c, engine, orc, dev, peak)

    if rank == 0:
        cfg = config_of(a)
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup,
            'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
            # the arithmetic of the timed path: float32 displacements + float32 sinc, summed in float64
            # (csrc/corr_kernels.cu::process_tile_f32d); other_kernels.isf_f64 is the float64 chain
            'dtype': 'f32', 'data': 'synthetic', 'config': cfg,
            'roofline': roofline, 'cpu_baseline': cpu_baseline, 'e2e': e2e, 'gpu_launches': int(launches),
            'clocks': clocks, 'parity': parity, 'chi4_max': float(np.max(x4)), 'other_kernels': other,
            'setup': {'trajectory_pass_s': t_build, 'kept_frames': int(len(ids)), 'numa': numa,
                      'image_rows': 'packed' if 'packed' in roofline['kernel'] else 'bytes'},
            'config2': config2,
        }
        emit(line)
    if world > 1:
        parallel.release_peer_comms()
        dist.destroy_process_group()


def measure_e2e(a, torch, md, c, parallel, res, store, stream, starts, idx, kw, world, N, n_local, barrier,
                max_over_ranks, ev, evals_step, dev):
    """End to end: the frames the schedule touches in page-locked host memory + the trajectory's jump matrices
    (scipy CSR, as ``reader.generate_nojump_matrixes`` leaves them); uploaded and turned into image counts every step."""
    from scipy import sparse
    from mdevaluate_b200.reader import SparseFrameReader
    T = a.frames
    ids = np.unique((starts[:, None] + idx[None, :]).ravel())
    rows = torch.from_numpy(store.rows_for(ids).astype(np.int64)).to(dev)
    n3 = 3 * n_local
    host = torch.empty((len(ids), n3), dtype=torch.float32, pin_memory=True)
    step_rows = max(1, (1 << 30) // (n3 * 4))
    for r0 in range(0, len(ids), step_rows):
        host[r0:r0 + step_rows].copy_(store.coords.index_select(0, rows[r0:r0 + step_rows])[:, :n3])
    torch.cuda.synchronize()
    t0 = time.time()
    fr, comp, val = res.sparse_jumps()                  # one more pass over the generated trajectory (setup)
    mats = tuple(sparse.csr_matrix((val[comp % 3 == d], (fr[comp % 3 == d], comp[comp % 3 == d] // 3)),
                                   shape=(T, n_local), dtype=np.int8) for d in range(3))
    prep_s = time.time() - t0
    base = SparseFrameReader(ids, host.numpy().reshape(len(ids), n_local, 3), stream.box_diag, stream.times)
    base.nojump_matrixes = mats

    def e2e_step():
        coords = md.Coordinates(base.fresh())
        if world > 1:
            parallel.shard_atoms(coords, n_atoms_global=N)
        return c.chi4(coords.nojump, Q, **kw)

    t, x4 = e2e_step()          # also converts the matrices to column lists once (host side, cached page-locked)
    lists = next(iter(base.host_cache.values()))
    list_bytes = int(sum(t_.numel() * t_.element_size() for t_ in lists[1:]))
    # the floor: the same page-locked bytes copied to the device by all ranks at once, nothing else
    sink = torch.empty((min(len(ids), step_rows), n3), dtype=torch.float32, device=dev)
    barrier()
    ev[0].record()
    for r0 in range(0, len(ids), step_rows):
        sink[:len(ids) - r0].copy_(host[r0:r0 + step_rows], non_blocking=True)
    ev[1].record()
    barrier()
    floor_ms = max_over_ranks(ev[0].elapsed_time(ev[1]))
    del sink
    barrier()
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(a.e2e_steps + 1)]
    ev[0].record()
    marks[0].record()
    for i in range(a.e2e_steps):
        t, x4_e2e = e2e_step()
        marks[i + 1].record()
    ev[1].record()
    barrier()
    e2e_ms = max_over_ranks(ev[0].elapsed_time(ev[1])) / a.e2e_steps
    each_ms = [max_over_ranks(marks[i].elapsed_time(marks[i + 1])) for i in range(a.e2e_steps)]
    S, L = len(starts), len(idx)
    h2d = int(len(ids)) * n3 * 4 + list_bytes
    return {'value': evals_step / (e2e_ms * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': h2d * world,
            'd2h_bytes_per_step': int(S * L * 3 * 8) * world, 'ms_per_step': e2e_ms, 'steps': a.e2e_steps,
            'ms_each_step': each_ms,
            'h2d_GBps_per_gpu': h2d / (e2e_ms * 1e-3) / 1e9, 'matches_resident': bool(np.array_equal(x4, x4_e2e)),
            'h2d_floor': {'ms': floor_ms, 'GBps_per_gpu': len(ids) * n3 * 4 / (floor_ms * 1e-3) / 1e9,
                          'GBps_all_gpus': world * len(ids) * n3 * 4 / (floor_ms * 1e-3) / 1e9,
                          'note': 'the frame bytes alone, copied by all %d ranks at the same time: what the host\'s '
                                  'PCIe / memory system delivers' % world},
            'note': 'per step and GPU: %d frames of the schedule (page-locked host, %.2f GB) + the jump matrices as column '
                    'lists (%.0f MB) -> HBM, image counts from the matrices, packing, correlation, chi4, result read '
                    'back; matrices prepared once in %.1f s (not timed)'
                    % (len(ids), len(ids) * n3 * 4 / 1e9, list_bytes / 1e6, prep_s)}


def config2_leg(a, torch, md, c, engine, orc, dev, peak):
    """BASELINE config 2 on one GPU, as round 1 measured it: ISF through shifted_correlation on 100 000 atoms x
    10 000 frames (dense in HBM), 100 origins x 82 lags; the other kernels of the path on the same rows; the one-off
    jump scan; end to end with the whole trajectory streamed from page-locked host memory every step."""
    from mdevaluate_b200.synthetic import random_walk_device, box_length
    N, T, SEG = 100000, 10000, 100
    block, box, times = random_walk_device(N, T, seed=20240, box=box_length(N), device=dev, pitch=engine.row_pitch(N))
    coords = md.Coordinates.from_arrays(block, box, times, n_atoms=N)
    frames = coords.nojump
    res = coords._resident()
    store = res.store_for(np.arange(T), nojump=True)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    scratch = torch.empty_like(store.images)
    scan_ms = []
    for _ in range(3):
        ev[0].record()
        res._scan_all(images=scratch)
        ev[1].record()
        torch.cuda.synchronize()
        scan_ms.append(ev[0].elapsed_time(ev[1]))
    assert torch.equal(scratch, store.images)
    del scratch
    scan_ms = float(min(scan_ms))
    fn = partial(c.isf, q=Q)
    kw = dict(segments=SEG)
    t, r = c.shifted_correlation(fn, frames, **kw)
    S, L = r.shape
    alg = 12.0 * N * S * (L + 1)
    for _ in range(3):
        c.shifted_correlation(fn, frames, **kw)
    steps = 20
    torch.cuda.synchronize()
    ev[0].record()
    for i in range(steps):
        engine.time_slot = i
        c.shifted_correlation(fn, frames, **kw)
    ev[1].record()
    torch.cuda.synchronize()
    engine.time_slot = -1
    ms_per_step = ev[0].elapsed_time(ev[1]) / steps
    plan = engine.last_plan()
    k_ms = float(np.mean([plan.kernel_ms(i) for i in range(steps)]))
    traffic, traffic_src = stored_traffic('isf:%d:%d:%d' % (N, S, L))
    out = {'workload': 'config2: ISF q=%.1f, %d atoms x %d frames, nojump, segments=%d, window=0.5, log lags' % (Q, N, T, SEG),
           'value': N * S * L / (ms_per_step * 1e-3), 'ms_per_step': ms_per_step, 'steps': steps,
           'roofline': {'kernel_ms': k_ms, 'frac': alg / (k_ms * 1e-3) / 1e9 / peak,
                        'through_api_frac': alg / (ms_per_step * 1e-3) / 1e9 / peak,
                        'host_overhead_ms': ms_per_step - k_ms, 'traffic': traffic, 'traffic_source': traffic_src,
                        'algorithmic_bytes_per_launch': alg},
           'nojump_pass': {'ms': scan_ms, 'frac': 15.0 * N * T / (scan_ms * 1e-3) / 1e9 / peak,
                           'note': 'one-off jump scan over the resident trajectory: 12 B read + 3 B written per '
                                   'atom*frame; packing the counts to 2 bits is one more pass over the 3 B'}}

    def kernel_of(call, reps=3, **tune):
        for k_, v_ in tune.items():
            setattr(engine.tuning, k_, v_)
        try:
            engine.time_slot = 0
            ms = []
            for _ in range(reps):
                call()
                ms.append(engine.last_plan().kernel_ms(0))
        finally:
            engine.time_slot = -1
            for k_ in tune:
                setattr(engine.tuning, k_, 0)
        return float(min(ms))
    bins = np.linspace(0, 3.0, 501)
    vh = partial(c.van_hove_self, bins=bins)
    legs = {'msd': lambda: c.shifted_correlation(c.msd, frames, **kw),
            'non_gaussian': lambda: c.shifted_correlation(c.non_gaussian, frames, **kw),
            'van_hove_self_500bins': lambda: c.shifted_correlation(vh, frames, **kw),
            'van_hove_self+non_gaussian (one pass)': lambda: c.shifted_correlation_multi({'vh': vh, 'a2': c.non_gaussian},
                                                                                         frames, **kw)}
    other = {}
    for name, call in legs.items():
        k = kernel_of(call)
        other[name] = {'kernel_ms': k, 'evals_per_s': N * S * L / (k * 1e-3), 'frac': alg / (k * 1e-3) / 1e9 / peak}
    k = kernel_of(lambda: c.shifted_correlation(fn, frames, **kw), isf_float64=1)
    other['isf_f64'] = {'kernel_ms': k, 'frac': alg / (k * 1e-3) / 1e9 / peak}
    k = kernel_of(legs['van_hove_self+non_gaussian (one pass)'], hist_float64=1)
    other['van_hove_self+non_gaussian f64 chain'] = {'kernel_ms': k, 'frac': alg / (k * 1e-3) / 1e9 / peak}
    out['other_kernels'] = other

    # dense schedules (BASELINE config 5's shape: every frame an origin, linear lags, average=True) on the first 2 000
    # frames of the same rows: the tiled kernel (csrc/dense_kernels.cu), bound by shared memory / FP32 / XU, not by HBM
    try:
        Td = 2000
        dcoords = md.Coordinates.from_arrays(block[:Td], box, times[:Td], n_atoms=N)
        lin = lambda lo, hi: np.arange(lo, hi)
        dense = {}
        for name, f in (('msd', c.msd), ('isf', fn)):
            c.shifted_correlation(f, dcoords.nojump, segments=8, window=0.5, average=True, index_distribution=lin)
            torch.cuda.synchronize()
            ev[0].record()
            _, r = c.shifted_correlation(f, dcoords.nojump, segments=Td // 2, window=0.5, average=True,
                                         index_distribution=lin)
            ev[1].record()
            torch.cuda.synchronize()
            ms = ev[0].elapsed_time(ev[1])
            dense[name] = {'ms': ms, 'evals_per_s': float(N) * (Td // 2) * len(r) / (ms * 1e-3)}
        dense['workload'] = 'dense: %d atoms, %d origins x %d linear lags, average=True (the call, scan excluded)' % (
            N, Td // 2, Td // 2)
        dense['full_size'] = 'profiles/r2_config5_full.json: 200 000 atoms x 50 000 frames from page-locked host memory'
        out['dense'] = dense
        del dcoords
    except Exception as exc:
        out['dense'] = {'error': '%s: %s' % (type(exc).__name__, exc)}

    # parity of the sub-block against the oracle
    n_sub = 4096
    sub = block[:, :3 * n_sub].cpu().numpy().reshape(T, n_sub, 3)
    unw = orc.UnwrappedTrajectory(sub, box, times)
    _, ref = orc.shifted_correlation(partial(orc.isf, q=Q), unw, **kw)
    starts, idx = c._schedule(frames, c.log_indices, SEG, 0.5, None)
    rows = (starts[:, None] + idx[None, :]).astype(np.int32)
    spec = c._kernel_spec(fn)
    sub_plan = engine.CorrPlan(store.atom_block(0, n_sub), rows, spec.corr)
    ours = spec.finalize(sub_plan.run()[0], None, n_sub, False)
    sub_plan.close()
    err = np.abs(ours - ref)
    out['parity'] = {'sub_block_atoms': n_sub, 'max_abs_err': float(err.max()),
                     'within_tolerance': bool((err <= 1e-5 * np.abs(ref) + 1e-7).all())}
    del sub, unw

    # end to end: the whole trajectory in page-locked host memory, streamed through the scan every step
    try:
        host = torch.empty((T, N, 3), dtype=torch.float32, pin_memory=True)
        host.view(T, 3 * N).copy_(block[:, :3 * N])
        torch.cuda.synchronize()
        host_np = host.numpy()

        def e2e_step():
            hc = md.Coordinates.from_arrays(host_np, box, times)
            hc._resident().force_compact = True
            return c.shifted_correlation(fn, hc.nojump, **kw)
        e2e_step()
        torch.cuda.synchronize()
        ev[0].record()
        for _ in range(3):
            e2e_step()
        ev[1].record()
        torch.cuda.synchronize()
        e2e_ms = ev[0].elapsed_time(ev[1]) / 3
        out['e2e'] = {'value': N * S * L / (e2e_ms * 1e-3), 'ms_per_step': e2e_ms, 'h2d_bytes_per_step': T * N * 12,
                      'd2h_bytes_per_step': int(S * L * 3 * 8),
                      'note': 'all %d frames from page-locked host memory every step, jump scan overlapped with the copy' % T}
    except Exception as exc:
        out['e2e'] = {'error': '%s: %s' % (type(exc).__name__, exc)}
    return out


def main():
    a = parse_args()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if a.gpus > 1 and world == 1 and 'RANK' not in os.environ:
        # convenience: launch ourselves under torchrun the way the driver does
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(a.gpus),
               '--master-addr', '127.0.0.1', '--master-port', os.environ.get('MASTER_PORT', '29541'),
               os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    claim_stdout()
    if a.impl == 'reference':
        reference_arm(a)
    else:
        ours_arm(a)


if __name__ == '__main__':
    main()
