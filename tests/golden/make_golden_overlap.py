"""Golden vectors for correlation.overlap on a coordinates.CoordinatesKDTree from the REAL reference (build container
only; needs scipy, as the reference does).

    python tests/golden/make_golden_overlap.py

Same loader as make_golden.py.  Writes tests/golden/overlap.npz: a small random walk in a periodic box and the overlap
through shifted_correlation for an open tree on raw frames, a periodic tree (``boxsize``), a tree of a selection
(``selector``) and an open tree on nojump frames, plus nearest distances of one frame pair for the oracle.
"""
import os
import sys
from functools import partial

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import load_reference, reference_coordinates          # noqa: E402


def half_selector(frame):
    """The atoms in the lower half of the box along x (a mask that changes from origin to origin)."""
    return np.where(np.asarray(frame)[:, 0] < 0.75)[0]


def main():
    md = load_reference()
    rng = np.random.default_rng(4242)
    n_atoms, n_frames, box_len = 400, 160, 1.5
    x = rng.uniform(0, box_len, size=(n_atoms, 3))
    pos = np.empty((n_frames, n_atoms, 3), dtype=np.float32)
    for t in range(n_frames):
        x = x + rng.normal(0, 0.012, size=x.shape)
        w = np.mod(x, box_len).astype(np.float32)
        w[w >= np.float32(box_len)] -= np.float32(box_len)
        pos[t] = w
    box = np.array([box_len] * 3, dtype=np.float32)
    times = np.arange(n_frames) * 0.25
    coords, reader = reference_coordinates(md, pos, box, times)
    corr, crd = md.correlation, md.coordinates
    out = {'pos': pos, 'box': box, 'times': times}
    kw = dict(segments=8, window=0.4)
    for radius in (0.03, 0.06):
        tree = crd.CoordinatesKDTree(coords)
        t, r = corr.shifted_correlation(partial(corr.overlap, crds_tree=tree, radius=radius), coords, **kw)
        out['open_%g' % radius] = r
        tree = crd.CoordinatesKDTree(coords, boxsize=box_len)
        t, r = corr.shifted_correlation(partial(corr.overlap, crds_tree=tree, radius=radius), coords, **kw)
        out['periodic_%g' % radius] = r
    tree = crd.CoordinatesKDTree(coords, selector=half_selector, boxsize=box_len)
    t, r = corr.shifted_correlation(partial(corr.overlap, crds_tree=tree, radius=0.05), coords, **kw)
    out['selector_0.05'] = r
    tree = crd.CoordinatesKDTree(coords.nojump)
    t, r = corr.shifted_correlation(partial(corr.overlap, crds_tree=tree, radius=0.05), coords.nojump, average=True, **kw)
    out['nojump_avg_0.05'] = r
    out['t'] = t
    # nearest distances of one frame pair (scipy through the reference's tree class), open and periodic
    out['dist_open'] = crd.CoordinatesKDTree(coords)[3].query(coords[90])[0]
    out['dist_periodic'] = crd.CoordinatesKDTree(coords, boxsize=box_len)[3].query(coords[90])[0]
    np.savez_compressed(os.path.join(HERE, 'overlap.npz'), **out)
    print('wrote overlap.npz', {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
