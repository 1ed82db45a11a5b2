"""Debug: histogram of an unaligned sub-store (direct-load kernel) vs an aligned copy (bulk kernel) vs the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mdevaluate_b200 as md
from mdevaluate_b200._cuda import engine
from mdevaluate_b200.synthetic import random_walk_device, box_length
from oracle import md_oracle as orc
from functools import partial
N, T = 20000, 600
dev = torch.device('cuda', 0)
block, box, times = random_walk_device(N, T, seed=1, box=box_length(N), device=dev, pitch=engine.row_pitch(N), sigma=0.05)
coords = md.Coordinates.from_arrays(block, box, times, n_atoms=N)
c = md.correlation
store = coords._resident().store_for(np.arange(T), nojump=True)
starts, idx = c._schedule(coords, c.log_indices, 5, 0.5, None)
rows = (starts[:, None] + idx[None, :]).astype(np.int32)
bins = np.linspace(0, 8.0, 501)
for flags in (8, 15, 11):
    spec = engine.CorrSpec(flags, q=22.7, n_bins=500, bin_mode=0, dx=float(np.diff(bins).mean()))
    full_s, full_h = engine.correlate_device(store, rows, spec)
    n0 = 5000 + 4 * 123
    a = engine.DeviceStore(store.coords, n0, images=store.images, box=store.box, box_const=store.box_const)
    off = 3 * n0
    b = engine.DeviceStore(store.coords[:, off:], N - n0, images=store.images[:, off:], box=store.box, box_const=store.box_const)
    b2 = engine.DeviceStore(store.coords[:, off:].contiguous(), N - n0, images=store.images[:, off:].contiguous(), box=store.box, box_const=store.box_const)
    sa, ha = engine.correlate_device(a, rows, spec)
    sb, hb = engine.correlate_device(b, rows, spec)
    sb2, hb2 = engine.correlate_device(b2, rows, spec)
    print('flags', flags, 'full total', int(full_h.sum()), 'a', int(ha.sum()), 'b(ldg)', int(hb.sum()), 'b2(bulk)', int(hb2.sum()),
          'a+b==full', bool(torch.equal(ha + hb, full_h)), 'a+b2==full', bool(torch.equal(ha + hb2, full_h)),
          'sums b vs b2 maxdiff', float((sb - sb2).abs().max()), 'sums a+b2 vs full', float((sa + sb2 - full_s).abs().max()))
