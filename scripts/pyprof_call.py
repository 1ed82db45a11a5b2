"""Python-side cost of a cached chi4 / shifted_correlation call: cProfile over many calls on a small resident block
(the kernel is a few microseconds, so the profile is the host path).  usage: python scripts/pyprof_call.py [calls]"""
import cProfile
import os
import pstats
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from functools import partial
import torch
import mdevaluate_b200 as md
from mdevaluate_b200 import correlation as c
from mdevaluate_b200._cuda import engine
from mdevaluate_b200.synthetic import random_walk_device, box_length
calls = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
N, T = 4096, 2000
block, box, times = random_walk_device(N, T, seed=1, box=box_length(N), pitch=engine.row_pitch(N))
coords = md.Coordinates.from_arrays(block, box, times, n_atoms=N).nojump
kw = dict(segments=20, window=0.2, skip=0.1)
fn = partial(c.isf, q=22.7)
for _ in range(20):
    c.chi4(coords, 22.7, **kw)
    c.shifted_correlation(fn, coords, average=False, **kw)
torch.cuda.synchronize()
plan = engine.last_plan()
for label, f in (('plan.run()', plan.run), ('shifted_correlation', lambda: c.shifted_correlation(fn, coords, average=False, **kw)),
                 ('chi4', lambda: c.chi4(coords, 22.7, **kw))):
    t0 = time.perf_counter()
    for _ in range(calls):
        f()
    print('%-24s %7.1f us per call' % (label, (time.perf_counter() - t0) / calls * 1e6))
engine.time_slot = 0
t0 = time.perf_counter()
for _ in range(calls):
    c.chi4(coords, 22.7, **kw)
print('%-24s %7.1f us per call' % ('chi4, kernel events on', (time.perf_counter() - t0) / calls * 1e6))
engine.time_slot = -1
pr = cProfile.Profile()
pr.enable()
for _ in range(calls):
    c.chi4(coords, 22.7, **kw)
pr.disable()
st = pstats.Stats(pr)
st.sort_stats('tottime').print_stats(22)
