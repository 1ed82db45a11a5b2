"""Spatially resolved ISF: atoms are picked by where they are at each time origin
(reference: examples/plot_spatialisf.py:87-99, doc/special/spatial.rst)."""
import os
import sys
from functools import partial

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mdevaluate_b200 as md
from mdevaluate_b200.synthetic import random_walk

positions, box, times = random_walk(20000, 2000)
traj = md.Coordinates.from_arrays(positions, box, times).nojump
func = partial(md.correlation.isf, q=22.7)
L = float(box[2])


def height(frame):
    """z of every atom inside the periodic box (the frames are unwrapped: fold back)."""
    return np.asarray(frame)[:, 2] % L


bins = np.linspace(0, L, 5)
for lo, hi in zip(bins[:-1], bins[1:]):
    selector = partial(md.coordinates.spatial_selector, transform=height, rmin=lo, rmax=hi)
    t, S = md.correlation.shifted_correlation(
        func, traj, segments=50, skip=0.1, average=True,
        correlation=md.correlation.subensemble_correlation(selector),
        description='%g,%g' % (lo, hi))
    print('slab z = %5.2f .. %5.2f nm:  F(q, t=%g) = %.4f   F(q, t=%g) = %.4f' % (lo, hi, t[10], S[10], t[-1], S[-1]))

# several slabs in ONE call, with the number of atoms in each: multi_subensemble_correlation
def slabs(frame):
    z = height(frame)
    return np.array([(z >= lo) & (z < hi) for lo, hi in zip(bins[:-1], bins[1:])])


t, S, count = md.correlation.shifted_correlation(
    func, traj, segments=50, skip=0.1, average=True,
    correlation=md.correlation.multi_subensemble_correlation(slabs))
print('multi: result', S.shape, ' mean atoms per slab', count)
