"""A/B of the no-image-change shortcut on one box: the ISF kernel of config 4 (and config 2's shape) with
tuning.variant bit 2 set / cleared, alternating.  usage: python scripts/shortcut_ab.py [atoms] [frames] [segments] [window] [skip]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from functools import partial
import numpy as np
import torch
import mdevaluate_b200 as md
from mdevaluate_b200 import correlation as c
from mdevaluate_b200._cuda import engine
from mdevaluate_b200.synthetic import random_walk_device, box_length

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 4400
S = int(sys.argv[3]) if len(sys.argv) > 3 else 20
window = float(sys.argv[4]) if len(sys.argv) > 4 else 0.9
block, box, times = random_walk_device(N, T, seed=1, box=box_length(N), pitch=engine.row_pitch(N))
coords = md.Coordinates.from_arrays(block, box, times, n_atoms=N).nojump
bins = np.linspace(0, 3.0, 501)
for name, fn in (('isf', partial(c.isf, q=22.7)), ('van_hove', partial(c.van_hove_self, bins=bins))):
    out = {}
    for rep in range(3):
        for variant in (0, 4):
            engine.tuning.variant = variant
            ms = []
            for _ in range(12):
                engine.time_slot = 0
                t, r = c.shifted_correlation(fn, coords, segments=S, window=window)
                ms.append(engine.last_plan().kernel_ms(0))
            engine.time_slot = -1
            out.setdefault(variant, []).append(float(np.median(ms[2:])))
            out.setdefault('r%d' % variant, r)
    engine.tuning.variant = 0
    L = len(t)
    gb = 12.0 * N * r.shape[0] * (L + 1) / 1e9
    print('%-9s N=%d S=%d L=%d  %.1f GB algorithmic: shortcut %s ms, general formula %s ms, identical results: %s'
          % (name, N, r.shape[0], L, gb, ['%.3f' % v for v in out[0]], ['%.3f' % v for v in out[4]],
             np.array_equal(out['r0'], out['r4'])), flush=True)
