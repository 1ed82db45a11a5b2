"""Are the van Hove counts of a large shard reproducible and complete?  usage: hist_determinism.py [lib.so] [atoms] [frames]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from functools import partial
from mdevaluate_b200._cuda import lib as _l
if len(sys.argv) > 1 and sys.argv[1].endswith('.so'):
    _l.LIB_PATH = os.path.abspath(sys.argv[1])
import mdevaluate_b200 as md
from mdevaluate_b200._cuda import engine
from mdevaluate_b200.synthetic import random_walk_device, box_length
c = md.correlation
N = int(sys.argv[2]) if len(sys.argv) > 2 else 500000
T = int(sys.argv[3]) if len(sys.argv) > 3 else 2400
block, box, times = random_walk_device(N, T, seed=5, box=box_length(N), pitch=engine.row_pitch(N))
coords = md.Coordinates.from_arrays(block, box, times, n_atoms=N).nojump
bins = np.linspace(0, 3.0, 501)
store = coords._resident().store_for(np.arange(T), nojump=True)
starts, idx = c._schedule(coords, c.log_indices, 20, 0.5, None)
rows = store.rows_for(starts[:, None] + idx[None, :])
print('lib', _l.LIB_PATH, 'N', N, 'S', len(starts), 'L', len(idx), 'packed' if store.packed is not None else 'bytes')
variants = [('packed counts, float32 displacements', {}), ('byte counts', {'use_packed': False}),
            ('float64 chain', {'hist_float64': 1})]
for label, tune in variants if '--all' in sys.argv else variants[:1]:
  for k_, v_ in tune.items():
    setattr(engine.tuning, k_, v_)
  print(label, flush=True)
  for name, fn in (('van_hove', partial(c.van_hove_self, bins=bins)),):
    spec = c._kernel_spec(fn)
    first = None
    for rep in range(12):
          plan = engine.CorrPlan(store, rows, spec.corr)
          s, h = plan.run()
          h = np.array(h)
          plan.close()
          tot = int(h.sum())
          if first is None:
              first = h
          diff = int(np.abs(h - first).sum())
          per_pair = h.sum(axis=-1)
          print('  rep %2d total %d  (N*S*L = %d)  pairs short of N: %d  |diff to rep 0| %d' % (rep, tot, N * h.shape[0] * h.shape[1], int((per_pair != N).sum()), diff), flush=True)
  for k_ in tune:
    setattr(engine.tuning, k_, True if k_ == 'use_packed' else 0)
