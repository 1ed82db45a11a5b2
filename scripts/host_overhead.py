"""Where the time of one resident shifted_correlation call goes beyond its kernel (config-2 shape, S origins).
usage: python scripts/host_overhead.py [S] [atoms] [frames]"""
import ctypes as C
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from functools import partial
import numpy as np, torch
import mdevaluate_b200 as md
from mdevaluate_b200 import correlation as c
from mdevaluate_b200._cuda import engine, lib as _lib
from mdevaluate_b200.synthetic import random_walk_device, box_length
S = int(sys.argv[1]) if len(sys.argv) > 1 else 20
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
T = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
block, box, times = random_walk_device(N, T, seed=1, box=box_length(N), pitch=engine.row_pitch(N))
coords = md.Coordinates.from_arrays(block, box, times, n_atoms=N).nojump
fn = partial(c.isf, q=22.7)
for _ in range(5):
    c.shifted_correlation(fn, coords, segments=S)
torch.cuda.synchronize()
n = 300


def timeit(label, f, reps=n):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    dt = (time.perf_counter() - t0) / reps * 1e6
    print('%-34s %8.1f us' % (label, dt), flush=True)
    return dt


engine.time_slot = 0
c.shifted_correlation(fn, coords, segments=S)
k_us = engine.last_plan().kernel_ms(0) * 1e3
engine.time_slot = -1
print('%-34s %8.1f us' % ('correlation kernel (events)', k_us))
total = timeit('shifted_correlation (cached call)', lambda: c.shifted_correlation(fn, coords, segments=S))
timeit('chi4 (call + var)', lambda: c.chi4(coords, 22.7, segments=S, window=0.5, skip=None))
plan = engine.last_plan()
run = timeit('plan.run()', lambda: plan.run())
lib = _lib.load()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
t_launch = t_wait = 0.0
for _ in range(n):
    t0 = time.perf_counter()
    lib.mdb_corr_plan_launch(plan.handle, st, -1)
    t1 = time.perf_counter()
    lib.mdb_corr_plan_wait(plan.handle)
    t2 = time.perf_counter()
    t_launch += t1 - t0
    t_wait += t2 - t1
print('%-34s %8.1f us' % ('  mdb_corr_plan_launch (2 launches)', t_launch / n * 1e6))
print('%-34s %8.1f us' % ('  mdb_corr_plan_wait', t_wait / n * 1e6))
print('%-34s %8.1f us' % ('  wait - kernel', t_wait / n * 1e6 - k_us))
for mode, label in ((1, 'mapped writes + cudaStreamSynchronize'), (2, 'cudaMemcpyAsync + cudaStreamSynchronize'), (0, 'mapped writes + flag polling')):
    lib.mdb_corr_plan_set_wait(plan.handle, mode)
    timeit('plan.run(), %s' % label, lambda: plan.run())
# the timeline of a run: host clock and the GPU's clock, related by a clock pair measured here
pairs = [engine.clock_pair() for _ in range(50)]
gap = min(seen - before for g, before, seen in pairs)
g, before, seen = min(pairs, key=lambda p: p[2] - p[1])
offset = g - (before + seen) / 2          # gpu clock - host clock, +- gap / 2
print('clock pair: launch-to-seen of a 1-thread kernel %.1f us (best of 50) -> clock offset known to +-%.1f us'
      % (gap / 1e3, gap / 2e3))
rows = []
for _ in range(n):
    plan.run()
    h0, h1, h2, g0, g1, g2 = plan.timeline()
    rows.append((h1 - h0, g0 - offset - h0, g1 - g0, g2 - g1, h2 - (g2 - offset), h2 - h0))
rows = np.median(np.array(rows, dtype=np.float64), axis=0) / 1e3
for label, v in zip(('launch call (2 kernels enqueued)', 'launch entered -> first CTA of the correlation kernel',
                     'correlation kernel start -> finish kernel start', 'finish kernel (start -> flag store)',
                     'flag store -> host saw it', 'plan.run() total (host clock)'), rows):
    print('  timeline: %-56s %8.1f us' % (label, v))
key, res = c._call_key(fn, coords, c.log_indices, S, 0.5, None)
timeit('_call_key', lambda: c._call_key(fn, coords, c.log_indices, S, 0.5, None))
entry = c._call_cache[key]
timeit('res.serves', lambda: res.serves(entry.store_ref(), True))
sums = np.array(plan.run()[0])
timeit('finalize', lambda: entry.spec.finalize(sums, None, N, False))
timeit('coords.nojump (copy)', lambda: coords.nojump)
print('%-34s %8.1f us' % ('call - kernel', total - k_us))
# an empty kernel launch + event sync for scale
x = torch.zeros(1, device='cuda')
timeit('torch tiny op + synchronize', lambda: (x.add_(1), torch.cuda.synchronize()))
