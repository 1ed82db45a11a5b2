"""Dense schedule runs for ncu captures: MSD then ISF through the tiled kernel.  usage: profile_dense.py [atoms] [frames]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from functools import partial
import numpy as np, torch
import mdevaluate_b200 as md
from mdevaluate_b200 import correlation as c
from mdevaluate_b200._cuda import engine
from mdevaluate_b200.synthetic import random_walk_device, box_length
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
block, box, times = random_walk_device(N, T, seed=1, box=box_length(N), pitch=engine.row_pitch(N))
coords = md.Coordinates.from_arrays(block, box, times, n_atoms=N).nojump
dense = lambda a, b: np.arange(a, b)
for fn in (c.msd, partial(c.isf, q=22.7)):
    t, r = c.shifted_correlation(fn, coords, segments=T // 2, window=0.5, average=True, index_distribution=dense)
    torch.cuda.synchronize()
    print(r[:3])
