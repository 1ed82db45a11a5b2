"""Small runs of every kernel family for compute-sanitizer (memcheck / racecheck); compares with the oracle so that a
hazard that changes a number is seen as well.  usage: compute-sanitizer --tool <tool> python scripts/sanitizer_cases.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from functools import partial
import numpy as np
import mdevaluate_b200 as md
from mdevaluate_b200._cuda import engine
from mdevaluate_b200.synthetic import random_walk
from oracle import md_oracle as orc

c = md.correlation
N, T = 2600, 160          # 2.5 tiles of 1024 atoms: bulk kernel, full and masked tiles
bins = np.linspace(0, 2.0, 101)
for box, label in ((6.0, 'packed image counts'), (1.3, 'byte image counts')):
    pos, bx, times = random_walk(N, T, seed=7, sigma=0.05 if box > 2 else 0.12, box=box)
    coords = md.Coordinates.from_arrays(pos, bx, times)
    unw = orc.UnwrappedTrajectory(pos, bx, times)
    store = coords._resident().store_for(np.arange(T), nojump=True)
    assert np.array_equal(store.image_counts(), unw.counts)
    print(label, 'packed' if store.packed is not None else 'bytes', flush=True)
    kw = dict(segments=5, window=0.5)
    with np.errstate(all='ignore'):
        for tune in ({}, {'isf_float64': 1, 'hist_float64': 1}):
            for k, v in tune.items():
                setattr(engine.tuning, k, v)
            for ours, theirs, exact in ((c.msd, orc.msd, False), (partial(c.isf, q=22.7), partial(orc.isf, q=22.7), False),
                                        (c.non_gaussian, orc.non_gaussian, False),
                                        (partial(c.van_hove_self, bins=bins), partial(orc.van_hove_self, bins=bins), True)):
                _, r = c.shifted_correlation(ours, coords.nojump, **kw)
                _, ref = orc.shifted_correlation(theirs, unw, **kw)
                if exact:
                    assert np.array_equal(r, ref)
                else:
                    ok = np.isnan(ref) | (np.abs(r - ref) <= 1e-5 * np.abs(ref) + 1e-7 * max(1.0, np.nanmax(np.abs(ref))))
                    assert ok.all()
            _, out = c.shifted_correlation_multi({'vh': partial(c.van_hove_self, bins=bins), 'a2': c.non_gaussian},
                                                 coords.nojump, **kw)
            _, ref = orc.shifted_correlation(partial(orc.van_hove_self, bins=bins), unw, **kw)
            assert np.array_equal(out['vh'], ref)
            for k in tune:
                setattr(engine.tuning, k, 0)
    sel = c.subensemble_correlation(lambda f: np.where(np.asarray(f)[:, 0] % box < 0.5 * box)[0])
    _, r = c.shifted_correlation(partial(c.isf, q=22.7), coords.nojump, correlation=sel, **kw)
    # raw frames, irregular bins (direct exact path), dense schedule, the peer-memory exchange on one GPU
    _, r = c.shifted_correlation(partial(c.van_hove_self, bins=np.array([0, 0.1, 0.25, 0.6, 2.0])), coords, **kw)
    dense = lambda a, b: np.arange(a, b)
    c.DENSE_MIN_PAIRS = 1
    _, r = c.shifted_correlation(c.msd, coords.nojump, segments=40, window=0.5, average=True, index_distribution=dense)
    _, r = c.shifted_correlation(partial(c.isf, q=22.7), coords.nojump, segments=40, window=0.5, average=True,
                                 index_distribution=dense)
    comm = engine.Comm(0, 1, lambda h: [h], capacity_words=1 << 15)
    starts, idx = c._schedule(coords, c.log_indices, 5, 0.5, None)
    rows = store.rows_for(starts[:, None] + idx[None, :])
    for fn in (partial(c.isf, q=22.7), partial(c.van_hove_self, bins=bins)):
        plan = engine.CorrPlan(store, rows, c._kernel_spec(fn).corr, comm=comm)
        plan.run(), plan.run()
        plan.close()
    comm.close()
    jumps = md.reader.parse_jumps(coords)
    frame = np.asarray(coords.nojump[T - 1])
    assert np.array_equal(frame, unw[T - 1])
print('sanitizer cases ok')
