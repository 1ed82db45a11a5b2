"""Rate of correlation.overlap through shifted_correlation (mdb_overlap) next to scipy's cKDTree on the host.
usage: python scripts/overlap_rate.py [atoms] [frames] [segments] [radius]"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from functools import partial
import numpy as np
import torch
import mdevaluate_b200 as md
from mdevaluate_b200 import correlation as c
from mdevaluate_b200._cuda import engine
from mdevaluate_b200.coordinates import CoordinatesKDTree
from mdevaluate_b200.synthetic import random_walk_device, box_length

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
S = int(sys.argv[3]) if len(sys.argv) > 3 else 10
radius = float(sys.argv[4]) if len(sys.argv) > 4 else 0.11
block, box, times = random_walk_device(N, T, seed=1, box=box_length(N), pitch=engine.row_pitch(N))
coords = md.Coordinates.from_arrays(block, box, times, n_atoms=N)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for label, frames, tree in (('raw frames, periodic tree', coords, CoordinatesKDTree(coords, boxsize=float(box[0]))),
                            ('raw frames, open tree', coords, CoordinatesKDTree(coords)),
                            ('nojump frames, open tree', coords.nojump, CoordinatesKDTree(coords.nojump))):
    fn = partial(c.overlap, crds_tree=tree, radius=radius)
    c.shifted_correlation(fn, frames, segments=S)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        ev[0].record()
        t, r = c.shifted_correlation(fn, frames, segments=S)
        ev[1].record()
        torch.cuda.synchronize()
        best = min(best, ev[0].elapsed_time(ev[1]))
    evals = N * r.shape[0] * r.shape[1]
    print('%-28s S=%d L=%d  %8.2f ms  %.3e atom*origin*lag evals/s   overlap(lag 1)=%.4f  (last lag)=%.4f'
          % (label, r.shape[0], r.shape[1], best, evals / (best * 1e-3), r[:, 1].mean(), r[:, -1].mean()), flush=True)
# the same with the atoms queried in file order (scattered candidate loads)
engine.overlap_file_order = True
tree = CoordinatesKDTree(coords, boxsize=float(box[0]))
fn = partial(c.overlap, crds_tree=tree, radius=radius)
c.shifted_correlation(fn, coords, segments=S)
torch.cuda.synchronize()
ev[0].record()
t, r2 = c.shifted_correlation(fn, coords, segments=S)
ev[1].record()
torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1])
print('%-28s S=%d L=%d  %8.2f ms  %.3e atom*origin*lag evals/s' % ('periodic tree, file order', r2.shape[0], r2.shape[1], ms,
                                                                  N * r2.size / (ms * 1e-3)))
engine.overlap_file_order = False
# the host: scipy's cKDTree as the reference uses it, one (origin, lag) pair
try:
    from scipy.spatial import cKDTree
    a = block[0, :3 * N].cpu().numpy().reshape(N, 3)
    b = block[T // 4, :3 * N].cpu().numpy().reshape(N, 3)
    t0 = time.perf_counter()
    tree = cKDTree(a, boxsize=float(box[0]))
    t1 = time.perf_counter()
    d = tree.query(b)[0]
    t2 = time.perf_counter()
    print('scipy cKDTree (1 core): build %.1f ms, query of %d atoms %.1f ms -> %.3e evals/s (query only)'
          % ((t1 - t0) * 1e3, N, (t2 - t1) * 1e3, N / (t2 - t1)))
except ImportError:
    print('scipy is not installed here')
