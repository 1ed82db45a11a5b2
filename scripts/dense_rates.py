"""Dense schedule (every frame an origin, linear lags, average=True): tiled kernel vs streaming kernel."""
import sys, time
sys.path.insert(0, '/root/repo')
from functools import partial
import numpy as np, torch
import mdevaluate_b200 as md
from mdevaluate_b200 import correlation as c
from mdevaluate_b200._cuda import engine
from mdevaluate_b200.synthetic import random_walk_device, box_length
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
block, box, times = random_walk_device(N, T, seed=1, box=box_length(N), pitch=engine.row_pitch(N))
coords = md.Coordinates.from_arrays(block, box, times, n_atoms=N).nojump
dense = lambda a, b: np.arange(a, b)
S = T // 2
ref = {}
for name, fn in (('msd', c.msd), ('isf', partial(c.isf, q=22.7))):
    c.shifted_correlation(fn, coords, segments=8, average=True, index_distribution=dense)      # warm (scan etc.)
    engine.tuning.dense_float64 = 1
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    t, r64 = c.shifted_correlation(fn, coords, segments=S, window=0.5, average=True, index_distribution=dense)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    engine.tuning.dense_float64 = 0
    print('%s tiled f64 S=%d L=%d N=%d: %.3f s  %.3e evals/s' % (name, S, len(r64), N, dt, N * S * len(r64) / dt))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    t, r = c.shifted_correlation(fn, coords, segments=S, window=0.5, average=True, index_distribution=dense)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    L = len(r)
    print('%s tiled v2  S=%d L=%d N=%d: %.3f s  %.3e evals/s   max rel diff to the float64 tile %.2e' % (name, S, L, N, dt, N * S * L / dt, np.abs(r / r64 - 1)[1:].max() if name == 'msd' else np.abs(r - r64).max()))
    save = c.DENSE_MIN_PAIRS
    c.DENSE_MIN_PAIRS = 1 << 62
    S2 = max(8, S // 16)
    t0 = time.perf_counter()
    t, r2 = c.shifted_correlation(fn, coords, segments=S2, window=0.5, average=True, index_distribution=dense)
    torch.cuda.synchronize()
    dt2 = time.perf_counter() - t0
    c.DENSE_MIN_PAIRS = save
    print('%s streaming S=%d L=%d N=%d: %.3f s  %.3e evals/s' % (name, S2, L, N, dt2, N * S2 * L / dt2))
