"""ISF of a periodic random walk (reference: examples/plot_isf.py:15-19)."""
import os
import sys
from functools import partial

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mdevaluate_b200 as md
from mdevaluate_b200.synthetic import random_walk

SIGMA, Q = 0.01, 22.7
positions, box, times = random_walk(20000, 2000, sigma=SIGMA)       # (T, N, 3) float32, wrapped into the box
OW = md.Coordinates.from_arrays(positions, box, times)              # instead of md.open(...).subset(atom_name='OW')

t, S = md.correlation.shifted_correlation(
    partial(md.correlation.isf, q=Q),
    OW.nojump,
    average=True
)
theory = np.exp(-Q * Q * SIGMA * SIGMA * t / 2)                      # Gaussian walk: F(q, t) = exp(-q^2 sigma^2 t / 2)
for i in range(0, len(t), max(1, len(t) // 12)):
    print('t = %7.1f   F(q,t) = %.5f   (Gaussian walk: %.5f)' % (t[i], S[i], theory[i]))
tau = t[np.argmin(np.abs(S - 1 / np.e))]
print('1/e time: %.1f frames (theory %.1f)' % (tau, 2 / (Q * Q * SIGMA * SIGMA)))
