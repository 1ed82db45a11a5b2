"""Four-point susceptibility from the per-origin ISF spread (reference: examples/plot_chi4.py:29-40)."""
import os
import sys
from functools import partial

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mdevaluate_b200 as md
from mdevaluate_b200.synthetic import random_walk

positions, box, times = random_walk(20000, 4000)
OW = md.Coordinates.from_arrays(positions, box, times).nojump

t, Fqt = md.correlation.shifted_correlation(
    partial(md.correlation.isf, q=22.7),
    OW,
    average=False,
    window=0.2,
    skip=0.1,
    segments=20
)
chi4 = len(OW[0]) * Fqt.var(axis=0)
smoothed = md.utils.runningmean(chi4, 5)
print('ISF per origin:', Fqt.shape, ' chi4:', chi4.shape)
for i in range(2, len(t) - 2, max(1, len(t) // 12)):
    print('t = %7.1f   chi4 = %8.4f   smoothed = %8.4f' % (t[i], chi4[i], smoothed[i - 2]))
# the same through the helper: md.correlation.chi4(OW, q=22.7, smooth=5)
