#!/usr/bin/env python
"""Benchmark of the shifted-correlation hot path (BASELINE.json metric: atom*origin*lag evals/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (N=1): BASELINE config 2 -- ISF at q = 22.7 nm^-1 through ``shifted_correlation`` on
100 000 atoms x 10 000 frames of a periodic random walk with nojump unwrapping, 100 origins x
82 log-spaced lags (8.2e8 evals per step).  With N GPUs every rank holds its own slab of
100 000 atoms (weak scaling, atom sharding, one all-reduce of the (S, L, 3) partial sums).

One step = one ``shifted_correlation`` call.  ``value`` is measured with the coordinates and the
int8 image counts resident in HBM; ``e2e`` is the same call on a trajectory in pinned HOST memory:
every step streams all frames to the GPU (jump scan overlapped with the copy), keeps the frames
the schedule touches, correlates, and reads the result back.  ``roofline`` is the correlation
kernel alone (CUDA events around the kernel, on its stream).  ``cpu_baseline`` is the NumPy port of
the reference (oracle/md_oracle.py, single-threaded like the reference) on an atom sub-block of
the same trajectory; the GPU result on that sub-block is compared with it (``parity``).

``--impl reference`` times that CPU port alone (the reference is pure Python/NumPy and
single-threaded: 1 core).
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
METRIC = 'atom*origin*lag evals/s (ISF via shifted_correlation, nojump unwrap)'
UNIT = 'evals/s'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--atoms', type=int, default=100000, help='atoms per GPU')
    ap.add_argument('--frames', type=int, default=10000)
    ap.add_argument('--segments', type=int, default=100)
    ap.add_argument('--window', type=float, default=0.5)
    ap.add_argument('--skip', type=float, default=None)
    ap.add_argument('--cpu-atoms', type=int, default=16384, help='atom sub-block timed on the CPU')
    ap.add_argument('--e2e-steps', type=int, default=5)
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--function', default='isf', choices=['isf', 'msd'])
    ap.add_argument('--pyprofile', action='store_true', help='cProfile the host side of 20 resident steps')
    return ap.parse_args()


def workload_name(a, world):
    return ('config2: %s q=%.1f, %d atoms/GPU x %d frames, nojump, segments=%d, window=%g, skip=%s, log lags'
            % (a.function.upper(), Q, a.atoms, a.frames, a.segments, a.window, a.skip))


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

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        for ts, line in self.lines:
            if t0 is not None and not (t0 - 0.05 <= ts <= t1 + 0.15):
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


def measured_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as fh:
            return float(json.load(fh)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


# ------------------------------------------------------------------------------------------------
# CPU port of the reference (the checker / baseline)
# ------------------------------------------------------------------------------------------------
def cpu_reference_step(pos, box, times, segments, function, window=0.5, skip=None):
    """One full reference-style evaluation on host arrays: parse_jumps + nojump + shifted_correlation."""
    from oracle import md_oracle as orc
    fn = partial(orc.isf, q=Q) if function == 'isf' else orc.msd
    t0 = time.perf_counter()
    unw = orc.UnwrappedTrajectory(pos, box, times)
    t1 = time.perf_counter()
    t, r = orc.shifted_correlation(fn, unw, segments=segments, window=window, skip=skip)
    t2 = time.perf_counter()
    return r, t1 - t0, t2 - t1


def reference_arm(a):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from mdevaluate_b200.synthetic import random_walk, box_length
    from oracle import md_oracle as orc
    starts, num = orc.schedule(a.frames, a.segments, a.window, a.skip or 0)
    S, L = len(starts), len(orc.log_indices(0, num))
    # bounded sample: aim at ~120 s for the whole run at the ~4e6 evals/s one core does on the whole step (jump
    # detection and unwrapping included; the per-frame Python overhead of the 10 000-frame pass sets a floor of a few
    # seconds per step whatever the sample)
    budget = 120.0 / max(1, a.steps + a.warmup)
    n = int(min(a.cpu_atoms, max(1024, budget * 4e6 / (S * L))))
    n = min(n, a.atoms)
    pos, box, times = random_walk(n, a.frames, seed=20240, box=box_length(a.atoms))
    evals = n * S * L
    for _ in range(a.warmup):
        cpu_reference_step(pos, box, times, a.segments, a.function, a.window, a.skip)
    t0 = time.perf_counter()
    scan_s = corr_s = 0.0
    for _ in range(a.steps):
        _, ts, tc = cpu_reference_step(pos, box, times, a.segments, a.function, a.window, a.skip)
        scan_s += ts
        corr_s += tc
    wall = time.perf_counter() - t0
    value = evals * a.steps / wall
    sample = ('%d-atom sub-block of the workload (all %d frames, S=%d, L=%d) per step: parse_jumps + nojump + '
              'shifted_correlation; correlation alone %.3g evals/s' % (n, a.frames, S, L, evals * a.steps / corr_s))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': a.gpus, 'steps': a.steps,
        'warmup': a.warmup, 'ms_per_step': 1e3 * wall / a.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': workload_name(a, 1), 'sample_atoms': n},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': 1, 'kind': 'port', 'sample': sample,
                         'host_cores': os.cpu_count()},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def ours_arm(a):
    import torch
    import torch.distributed as dist

    import mdevaluate_b200 as md
    from mdevaluate_b200 import parallel
    from mdevaluate_b200._cuda import engine
    from mdevaluate_b200.synthetic import random_walk_device, box_length

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py --impl ours needs a CUDA device (there is no CPU fallback)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        import datetime
        # a stuck collective must fail loudly, not hang the run
        dist.init_process_group('nccl', device_id=dev, timeout=datetime.timedelta(seconds=180))

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
    fn = partial(c.isf, q=Q) if a.function == 'isf' else c.msd
    N, T = a.atoms, a.frames
    L_box = box_length(N)          # every rank: the density of one slab
    block, box, times = random_walk_device(N, T, seed=20240 + rank, box=L_box, device=dev,
                                           pitch=engine.row_pitch(N))
    torch.cuda.synchronize()

    def make_coords(source, **kw):
        coords = md.Coordinates.from_arrays(source, box, times, **kw)
        if world > 1:
            parallel.shard_atoms(coords, n_atoms_global=N * world)
        return coords

    coords = make_coords(block, n_atoms=N)
    # ---- one-off unwrap pass over the resident trajectory (reported, not in the headline) ------
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    resident = coords._resident()
    store = resident.store_for(np.arange(T), nojump=True)       # builds the int8 image block (first call)
    torch.cuda.synchronize()
    scratch = torch.empty_like(store.images)
    scan_ms = []
    for _ in range(3):                                          # the same scan again, warm, into scratch
        ev[0].record()
        resident._scan_all(images=scratch)
        ev[1].record()
        torch.cuda.synchronize()
        scan_ms.append(ev[0].elapsed_time(ev[1]))
    assert torch.equal(scratch, store.images)
    del scratch
    scan_ms = float(min(scan_ms))
    peak, peak_src = measured_peak()
    scan_bytes = 15.0 * N * T
    nojump_pass = {'ms': scan_ms, 'atom_frames_per_s': N * T / (scan_ms * 1e-3),
                   'algorithmic_GBps': scan_bytes / (scan_ms * 1e-3) / 1e9,
                   'frac': scan_bytes / (scan_ms * 1e-3) / 1e9 / peak,
                   'note': 'one-off unwrap pass over the resident trajectory: 12 B read + 3 B written per atom*frame; '
                           'includes the host-side overflow check (one small D2H)'}

    frames = coords.nojump
    t_axis, res = c.shifted_correlation(fn, frames, segments=a.segments, window=a.window, skip=a.skip)
    S, L = res.shape
    evals_step = N * world * S * L
    for _ in range(a.warmup):
        c.shifted_correlation(fn, frames, segments=a.segments, window=a.window, skip=a.skip)

    if a.pyprofile and rank == 0:
        import cProfile
        import pstats
        prof = cProfile.Profile()
        prof.enable()
        for _ in range(20):
            c.shifted_correlation(fn, frames, segments=a.segments, window=a.window, skip=a.skip)
        prof.disable()
        pstats.Stats(prof, stream=sys.stderr).sort_stats('cumulative').print_stats(25)

    # ---- timed region: K steps, inputs resident in HBM ---------------------------------------------
    engine.profile_kernels(True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = engine.launches()
    kernel_ms = []
    barrier()
    w0 = time.time()
    ev[0].record()
    for _ in range(a.steps):
        t_axis, res = c.shifted_correlation(fn, frames, segments=a.segments, window=a.window, skip=a.skip)
        kernel_ms.append(engine.last_kernel_ms())
    ev[1].record()
    barrier()
    w1 = time.time()
    elapsed_ms = max_over_ranks(ev[0].elapsed_time(ev[1]))
    launches = engine.launches() - launches0
    clocks = sampler.stop(w0, w1) if rank == 0 else None
    engine.profile_kernels(False)
    value = evals_step * a.steps / (elapsed_ms * 1e-3)

    k_ms = float(np.mean(kernel_ms))
    alg_bytes = 12.0 * N * S * (L + 1)
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as fh:
            entry = json.load(fh).get('%s:%d:%d:%d' % (a.function, N, T, a.segments))
        traffic = entry['bytes'] if entry else None
    except Exception:
        traffic = None
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': traffic, 'kernel': 'corr_kernel<float, images, SINC>' if a.function == 'isf'
                else 'corr_kernel<float, images, R2>', 'kernel_ms': k_ms, 'kernel_ms_min': float(np.min(kernel_ms)),
                'kernel_ms_median': float(np.median(kernel_ms)), 'kernel_ms_max': float(np.max(kernel_ms)),
                'algorithmic_bytes_per_launch': alg_bytes, 'peak_source': peak_src,
                'moved_bytes_per_launch': 15.0 * N * S * (L + 1),
                'note': 'algorithmic bytes = 12 B (float32 xyz) per atom*origin*(lag+1); the kernel also reads 3 B of '
                        'image-count bytes per coordinate triple (nojump), so moved bytes are 1.25x algorithmic'}

    # ---- the other kernels of the path on the same rows (kernel time only; informational) ------------
    # (every rank takes part: with N > 1 each call ends in the all-reduce of the partial sums)
    other = {}
    engine.profile_kernels(True)
    vh_bins = np.linspace(0, 3.0, 501)
    extra = {'msd': c.msd, 'non_gaussian': c.non_gaussian, 'van_hove_self_500bins': partial(c.van_hove_self, bins=vh_bins)}
    if a.function != 'isf':
        extra['isf'] = partial(c.isf, q=Q)
    for name, f in extra.items():
        ms = []
        for _ in range(3):
            c.shifted_correlation(f, frames, segments=a.segments, window=a.window, skip=a.skip)
            ms.append(engine.last_kernel_ms())
        other[name] = {'kernel_ms': float(min(ms)), 'evals_per_s': N * S * L / (min(ms) * 1e-3),
                       'frac': alg_bytes / (min(ms) * 1e-3) / 1e9 / peak}
    ms = []
    for _ in range(3):
        c.shifted_correlation_multi({'vh': extra['van_hove_self_500bins'], 'a2': c.non_gaussian}, frames,
                                    segments=a.segments, window=a.window, skip=a.skip)
        ms.append(engine.last_kernel_ms())
    other['van_hove_self+non_gaussian (one pass)'] = {'kernel_ms': float(min(ms)),
                                                      'evals_per_s': N * S * L / (min(ms) * 1e-3),
                                                      'frac': alg_bytes / (min(ms) * 1e-3) / 1e9 / peak}
    engine.profile_kernels(False)

    # ---- parity + CPU baseline on an atom sub-block (rank 0, N=1 only) ------------------------------
    cpu_baseline = None
    parity = None
    if rank == 0 and world == 1 and not a.no_cpu:
        n_sub = min(a.cpu_atoms, N)
        sub = block[:, :3 * n_sub].cpu().numpy().reshape(T, n_sub, 3)
        ref, scan_s, corr_s = cpu_reference_step(sub, box, times, a.segments, a.function, a.window, a.skip)
        sub_store = engine.DeviceStore(store.coords, n_sub, images=store.images, box=store.box,
                                       box_const=store.box_const)
        starts, idx = c._schedule(frames, c.log_indices, a.segments, a.window, a.skip)
        rows = (starts[:, None] + idx[None, :]).astype(np.int32)
        spec = c._kernel_spec(fn)
        sums, _ = engine.correlate_device(sub_store, rows, spec.corr)
        ours = spec.finalize(sums.cpu().numpy(), None, n_sub, False)
        err = np.abs(ours - ref)
        tol = 1e-5 * np.abs(ref) + 1e-7 * (1.0 if a.function == 'isf' else float(np.abs(ref).max()))
        parity = {'sub_block_atoms': n_sub, 'max_abs_err': float(err.max()),
                  'max_rel_err': float((err / np.maximum(np.abs(ref), 1e-300)).max()),
                  'within_tolerance': bool((err <= tol).all()), 'tolerance': '1e-5*|ref| + 1e-7*scale'}
        evals_sub = n_sub * S * L
        cpu_baseline = {'value': evals_sub / (scan_s + corr_s), 'unit': UNIT, 'cores': 1, 'kind': 'port',
                        'host_cores': os.cpu_count(),
                        'sample': '%d-atom sub-block, all %d frames, S=%d, L=%d: parse_jumps+nojump %.2f s, '
                                  'shifted_correlation %.2f s (correlation alone %.3g evals/s)'
                                  % (n_sub, T, S, L, scan_s, corr_s, evals_sub / corr_s)}
        del sub

    # ---- end to end: trajectory in pinned host memory, streamed every step --------------------------
    e2e = None
    if not a.no_e2e:
        try:
            e2e = measure_e2e(a, torch, c, fn, block, box, times, make_coords, barrier, max_over_ranks, ev, evals_step,
                              S, L)
        except Exception as exc:      # e.g. not enough page-locked host memory: report it, keep the line
            e2e = {'value': None, 'unit': UNIT, 'h2d_bytes_per_step': int(T) * N * 12,
                   'd2h_bytes_per_step': int(S * L * 3 * 8), 'error': '%s: %s' % (type(exc).__name__, exc)}
            if world > 1:
                raise

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup,
            'ms_per_step': elapsed_ms / a.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            # the arithmetic of the timed path: ISF = float32 displacements + float32 sinc, summed in float64
            # (csrc/corr_kernels.cu::process_tile_sinc_f32); MSD = float64 throughout, like the reference
            'dtype': 'f32' if a.function == 'isf' else 'f64', 'data': 'synthetic',
            'config': {'workload': workload_name(a, world), 'atoms_total': N * world, 'frames': T, 'origins': int(S),
                       'lags': int(L), 'evals_per_step': int(evals_step), 'parallelism': 'atom-shard x%d' % world,
                       'l2': 'inputs larger than L2: each step streams %.1f GB of frame rows' % (alg_bytes / 1e9)},
            'roofline': roofline, 'cpu_baseline': cpu_baseline, 'e2e': e2e, 'gpu_launches': int(launches),
            'clocks': clocks, 'parity': parity, 'nojump_pass': nojump_pass, 'other_kernels': other,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def measure_e2e(a, torch, c, fn, block, box, times, make_coords, barrier, max_over_ranks, ev, evals_step, S, L):
    """End to end: the trajectory sits in pinned host memory and is streamed to the GPU every step."""
    T, N = a.frames, a.atoms
    host = torch.empty((T, N, 3), dtype=torch.float32, pin_memory=True)
    host.view(T, 3 * N).copy_(block[:, :3 * N])
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    host_np = host.numpy()

    def e2e_step():
        hc = make_coords(host_np)
        hc._resident().force_compact = True
        return c.shifted_correlation(fn, hc.nojump, segments=a.segments, window=a.window, skip=a.skip)

    e2e_step()
    barrier()
    ev[0].record()
    for _ in range(a.e2e_steps):
        _, res_e2e = e2e_step()
    ev[1].record()
    barrier()
    e2e_ms = max_over_ranks(ev[0].elapsed_time(ev[1])) / a.e2e_steps
    e2e = {'value': evals_step / (e2e_ms * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': int(T) * N * 12,
           'd2h_bytes_per_step': int(S * L * 3 * 8), 'ms_per_step': e2e_ms, 'steps': a.e2e_steps,
           'note': 'pinned host (T,N,3) float32 -> HBM in time segments overlapped with the jump scan, '
                   'needed frames kept, correlation, result read back'}
    return e2e


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
