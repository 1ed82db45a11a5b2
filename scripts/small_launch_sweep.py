"""Kernel time of the ISF launch on the config-4 per-GPU shapes (chi4 schedule: S=20, L=81) for different atom-block
splits and ring depths.  Usage: python scripts/small_launch_sweep.py [atoms] [frames]
The trajectory is built compact (only the schedule's frames are kept) so that 10^6 atoms fit one GPU."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from functools import partial
import numpy as np, torch
import mdevaluate_b200 as md
from mdevaluate_b200._cuda import engine
from mdevaluate_b200.synthetic import random_walk_device, box_length

N = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
dev = torch.device('cuda', 0)
block, box, times = random_walk_device(N, T, seed=1, box=box_length(N), device=dev, pitch=engine.row_pitch(N))
coords = md.Coordinates.from_arrays(block, box, times, n_atoms=N)
c = md.correlation
fn = partial(c.isf, q=22.7)
frames = coords.nojump
kw = dict(segments=20, window=0.2, skip=0.1)
engine.profile_kernels(True)


def run(label):
    for _ in range(3):
        t, r = c.shifted_correlation(fn, frames, **kw)
    ms = []
    for _ in range(15):
        t, r = c.shifted_correlation(fn, frames, **kw)
        ms.append(engine.last_kernel_ms())
    S, L = r.shape
    k = float(np.median(ms))
    alg = 12.0 * N * S * (L + 1)
    print('%-28s N=%d S=%d L=%d kernel median %.4f ms (min %.4f)  frac %.3f' % (label, N, S, L, k, min(ms),
                                                                               alg / k / 1e6 / 6453.1), flush=True)


run('default')
for ab in (1, 2, 3, 4, 5, 6, 8, 10, 12, 16, 24, 31, 41, 62, 123):
    os.environ['MDB_ABLOCKS'] = str(ab)
    run('MDB_ABLOCKS=%d' % ab)
os.environ.pop('MDB_ABLOCKS')
for st in (3, 4, 5, 6):
    os.environ['MDB_STAGES'] = str(st)
    run('MDB_STAGES=%d' % st)
os.environ.pop('MDB_STAGES')
os.environ['MDB_SINC_F64'] = '1'
run('MDB_SINC_F64=1')
os.environ.pop('MDB_SINC_F64')
fn = c.msd
run('msd default')
