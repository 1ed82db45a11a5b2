"""Debug: where do van Hove counts differ between the kernel variants (and the oracle)?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from functools import partial
from mdevaluate_b200._cuda import lib as _l
if len(sys.argv) > 1:
    _l.LIB_PATH = os.path.abspath(sys.argv[1])
import mdevaluate_b200 as md
from mdevaluate_b200._cuda import engine
from mdevaluate_b200.synthetic import random_walk
from oracle import md_oracle as orc
c = md.correlation
pos, box, times = random_walk(20000, 300, seed=3, sigma=0.05, box=14.4)
big = md.Coordinates.from_arrays(pos, box, times)
bins = np.linspace(0, 3.0, 501)
vh = partial(c.van_hove_self, bins=bins)
unw = orc.UnwrappedTrajectory(pos, box, times)
_, ref = orc.shifted_correlation(partial(orc.van_hove_self, bins=bins), unw, segments=20)
ref = np.rint(ref * 20000).astype(np.int64)
res = {}
for name, packed, h64 in (('packed f32d', True, 0), ('packed f64', True, 1), ('bytes f32d', False, 0), ('bytes f64', False, 1)):
    engine.tuning.use_packed, engine.tuning.hist_float64 = packed, h64
    for rep in range(3):
        _, got = c.shifted_correlation(vh, big.nojump, segments=20)
        got = np.rint(got * 20000).astype(np.int64)
        d = got - ref
        bad = np.argwhere(d != 0)
        print('%-12s rep %d: total %d (ref %d)  mismatching cells %d  sum of |diff| %d' % (name, rep, got.sum(), ref.sum(), len(bad), np.abs(d).sum()))
        for s, l, b in bad[:6]:
            print('     origin %d lag %d bin %d: got %d ref %d' % (s, l, b, got[s, l, b], ref[s, l, b]))
        per_pair = got.sum(axis=-1) - ref.sum(axis=-1)
        print('     (origin, lag) pairs with a wrong total: %d  (min %d max %d)' % ((per_pair != 0).sum(), per_pair.min(), per_pair.max()))
