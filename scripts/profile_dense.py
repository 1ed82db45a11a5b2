import sys
sys.path.insert(0, '/root/repo')
from functools import partial
import numpy as np, torch
import mdevaluate_b200 as md
from mdevaluate_b200 import correlation as c
from mdevaluate_b200._cuda import engine
from mdevaluate_b200.synthetic import random_walk_device, box_length
N, T = 100000, 1000
block, box, times = random_walk_device(N, T, seed=1, box=box_length(N), pitch=engine.row_pitch(N))
coords = md.Coordinates.from_arrays(block, box, times, n_atoms=N).nojump
dense = lambda a, b: np.arange(a, b)
for _ in range(2):
    t, r = c.shifted_correlation(partial(c.isf, q=22.7), coords, segments=T // 2, window=0.5, average=True,
                                 index_distribution=dense)
torch.cuda.synchronize()
print(r[:3])
