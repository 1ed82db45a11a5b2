"""Where the host time of one resident shifted_correlation call goes (S=10, config-2 shape)."""
import sys, time
sys.path.insert(0, '/root/repo')
from functools import partial
import numpy as np, torch
import mdevaluate_b200 as md
from mdevaluate_b200 import correlation as c
from mdevaluate_b200._cuda import engine
from mdevaluate_b200.synthetic import random_walk_device, box_length
N, T, S = 100000, 10000, int(sys.argv[1]) if len(sys.argv) > 1 else 10
block, box, times = random_walk_device(N, T, seed=1, box=box_length(N), pitch=engine.row_pitch(N))
coords = md.Coordinates.from_arrays(block, box, times, n_atoms=N).nojump
fn = partial(c.isf, q=22.7)
for _ in range(5):
    c.shifted_correlation(fn, coords, segments=S)
torch.cuda.synchronize()
n = 200
t0 = time.perf_counter()
for _ in range(n):
    c.shifted_correlation(fn, coords, segments=S)
t1 = time.perf_counter()
print('call total      %.1f us' % ((t1 - t0) / n * 1e6))
engine.profile_kernels(True)
c.shifted_correlation(fn, coords, segments=S)
print('kernel          %.1f us' % (engine.last_kernel_ms() * 1e3))
engine.profile_kernels(False)
def timeit(label, f):
    t0 = time.perf_counter()
    for _ in range(n): f()
    print('%-15s %.1f us' % (label, (time.perf_counter() - t0) / n * 1e6))
spec = c._kernel_spec(fn)
timeit('_kernel_spec', lambda: c._kernel_spec(fn))
timeit('_schedule', lambda: c._schedule(coords, c.log_indices, S, 0.5, None))
starts, idx = c._schedule(coords, c.log_indices, S, 0.5, None)
ids = starts[:, None] + idx[None, :]
timeit('frame_ids', lambda: (starts[:, None] + idx[None, :]).max())
timeit('_times', lambda: c._times(coords, idx))
timeit('_provider', lambda: c._provider(coords, ids))
store, rows, n_atoms, f32, comm = c._provider(coords, ids)
plan = c._plan_for(store, rows, spec.corr, None, None)
timeit('plan lookup', lambda: c._plan_for(store, rows, spec.corr, None, None))
timeit('plan.run', lambda: plan.run())
sums = np.array(plan.run()[0])
timeit('finalize', lambda: spec.finalize(sums, None, n_atoms, False))
