"""BASELINE config 5 in shape, scaled in time: dense all-lag MSD / ISF (every frame an origin, linear lags,
average=True) on 200 000 atoms, the trajectory in pinned HOST memory, processed in atom slabs whose rows are
uploaded while the previous slab's kernel runs (the full 50 000 frames = 120 GB take ~4 GPU-minutes per
function; T scales the work quadratically).

    python scripts/config5_scaled.py [T=5000] [slab_atoms=102400]
"""
import sys
import time
sys.path.insert(0, '/root/repo')
from functools import partial
import numpy as np, torch
import mdevaluate_b200 as md
from mdevaluate_b200 import correlation as c
from mdevaluate_b200._cuda import engine
from mdevaluate_b200.synthetic import random_walk_device, box_length

N = 200000
T = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
SLAB = int(sys.argv[2]) if len(sys.argv) > 2 else 102400
block, box, times = random_walk_device(N, T, seed=5, box=box_length(N), pitch=engine.row_pitch(N))
host = torch.empty((T, N, 3), dtype=torch.float32, pin_memory=True)
host.view(T, 3 * N).copy_(block[:, :3 * N])
torch.cuda.synchronize()
dense = lambda a, b: np.arange(a, b)
kw = dict(segments=T // 2, window=0.5, average=True, index_distribution=dense)
resident = md.Coordinates.from_arrays(block, box, times, n_atoms=N).nojump
c.shifted_correlation(c.msd, resident, segments=8, average=True, index_distribution=dense)     # scan, warm-up
out = {}
for name, fn in (('msd', c.msd), ('isf', partial(c.isf, q=22.7))):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    t, r_res = c.shifted_correlation(fn, resident, **kw)
    dt_res = time.perf_counter() - t0
    S = T // 2
    L = len(r_res)
    evals = float(N) * S * L
    c.DENSE_SLAB_ATOMS = SLAB
    t0 = time.perf_counter()
    t, r_ooc = c.shifted_correlation(fn, md.Coordinates.from_arrays(host.numpy(), box, times).nojump, **kw)
    dt_ooc = time.perf_counter() - t0
    c.DENSE_SLAB_ATOMS = None
    err = float(np.max(np.abs(r_ooc - r_res) / np.maximum(np.abs(r_res), 1e-300)))
    print('%s: S=%d L=%d N=%d  %.3e evals | resident %.3f s = %.3e evals/s | host, %d-atom slabs %.3f s = %.3e evals/s '
          '(H2D %.1f GB) | max rel diff %.2e' % (name, S, L, N, evals, dt_res, evals / dt_res, SLAB, dt_ooc,
                                                 evals / dt_ooc, T * N * 12 / 1e9, err))
