"""Pin the CPU oracle (oracle/md_oracle.py) to vectors produced by the real reference.

The fixtures under tests/golden/ were written by tests/golden/make_golden.py, which
imports /root/reference/mdevaluate itself.  Integer results must be identical; float
results use the same NumPy expressions and must be identical too (rtol 0).
"""
from functools import partial

import numpy as np
import pytest

from oracle import md_oracle as orc
from mdevaluate_b200.synthetic import random_walk


def same(a, b):
    np.testing.assert_array_equal(np.asarray(a), np.asarray(b))


def test_log_indices_golden(golden):
    g = golden('schedule')
    for key in g.files:
        _, first, last = key.split('_')
        same(orc.log_indices(int(first), int(last)), g[key])


def test_histogram_golden(golden):
    g = golden('histogram')
    for k in ('b0', 'b1', 'b2'):
        h, c = orc.histogram(g['data'], g[k])
        same(h, g['h_' + k])
        same(c, g['c_' + k])
    same(orc.runningmean(g['data'][:40], 5), g['runmean5'])


def test_histogram_matches_numpy_like_reference_test():
    # the reference's own test (test/test_utils.py:36-46), seeded
    data = np.random.default_rng(5).random(100)
    for bins in (np.linspace(0, 1), np.linspace(0.3, 1.5)):
        same(orc.histogram(data, bins)[0], np.histogram(data, bins=bins)[0])


@pytest.fixture(scope='module')
def small(golden):
    g = golden('walk_small')
    raw = orc.ArrayTrajectory(g['pos'], g['box'], g['times'])
    unw = orc.UnwrappedTrajectory(g['pos'], g['box'], g['times'])
    return g, raw, unw


def test_jump_table_golden(small):
    g, raw, unw = small
    same(unw.jumps, g['jumps'])
    assert np.count_nonzero(g['jumps']) > 500      # the fixture really exercises jumps


def test_unwrapped_frames_golden(small):
    g, raw, unw = small
    for i in (0, 1, 17, 100, 239):
        f = unw[i]
        assert f.dtype == np.float64
        same(f, g['nojump_%d' % i])


def _functions(g):
    return {
        'msd': orc.msd,
        'isf': partial(orc.isf, q=22.7),
        'ng': orc.non_gaussian,
        'vh': partial(orc.van_hove_self, bins=g['bins']),
        'vh_off': partial(orc.van_hove_self, bins=g['bins_off']),
        'vh_irr': partial(orc.van_hove_self, bins=g['bins_irr']),
    }


@pytest.mark.parametrize('name', ['msd', 'isf', 'ng', 'vh', 'vh_off', 'vh_irr'])
@pytest.mark.parametrize('mode', ['nojump', 'raw'])
def test_shifted_correlation_golden(small, name, mode):
    g, raw, unw = small
    frames = unw if mode == 'nojump' else raw
    with np.errstate(all='ignore'):
        t, r = orc.shifted_correlation(_functions(g)[name], frames, segments=12, window=0.4)
    same(t, g['t_' + mode])
    same(r, g['%s_%s' % (name, mode)])
    assert r.dtype == g['%s_%s' % (name, mode)].dtype


def test_chi4_average_slice_dense_golden(small):
    g, raw, unw = small
    fns = _functions(g)
    t, r = orc.shifted_correlation(fns['isf'], unw, segments=20, window=0.2, skip=0.1, average=False)
    same(r, g['chi4_fqt'])
    same(t, g['chi4_t'])
    same(orc.chi4(g['pos'].shape[1], r), g['chi4'])
    t, r = orc.shifted_correlation(fns['msd'], unw, segments=7, window=0.3, average=True)
    same(r, g['msd_avg'])
    same(t, g['msd_avg_t'])
    t, r = orc.shifted_correlation(fns['msd'], unw[24:], segments=7, window=0.3, average=True)
    same(r, g['msd_slice24'])
    t, r = orc.shifted_correlation(fns['isf'], unw, segments=5, window=0.25,
                                   index_distribution=orc.lin_indices)
    same(r, g['isf_dense'])


def test_selection_golden(small):
    g, raw, unw = small
    sel = g['sel']
    frames = orc.UnwrappedTrajectory(g['pos'], g['box'], g['times'], selection=sel)
    t, r = orc.shifted_correlation(orc.msd, frames, segments=6, window=0.4)
    same(r, g['msd_sel'])


def test_config1_golden(golden):
    """BASELINE config 1: MSD, 1000 atoms x 1000 frames, 100 origins, 70 log lags."""
    g = golden('config1')
    pos, box, times = random_walk(1000, 1000, seed=20240)
    same(box, g['box'])
    assert np.float64(pos.astype(np.float64).sum()) == g['pos_sum']
    same(pos[::97, ::53], g['pos_probe'])
    unw = orc.UnwrappedTrajectory(pos, box, times)
    assert np.count_nonzero(unw.jumps) == g['jump_nnz']
    same(unw.counts[-1], g['count_checksum'])
    t, r = orc.shifted_correlation(orc.msd, unw, segments=100, window=0.5)
    assert r.shape == (100, 70)
    same(t, g['t'])
    same(r, g['msd'])
    raw = orc.ArrayTrajectory(pos, box, times)
    t, r = orc.shifted_correlation(partial(orc.isf, q=22.7), raw, segments=100, window=0.5)
    same(r, g['isf_raw'])
    assert r.dtype == np.float32
    t, r = orc.shifted_correlation(partial(orc.van_hove_self, bins=g['vh_bins']), unw, segments=10, window=0.5)
    same(r, g['vh'])


# ---- sub-ensembles (correlation.py:27-81, 153-189) ---------------------------------------------------
def _selectors(L):
    def slab(frame, lo=0.0, hi=0.4):
        z = np.asarray(frame)[:, 2] % L
        return np.where((z >= lo * L) & (z < hi * L))[0]

    def slabs_masks(frame):
        z = np.asarray(frame)[:, 2] % L
        edges = np.array([0.0, 0.3, 0.65, 1.0]) * L
        return np.array([(z >= edges[i]) & (z < edges[i + 1]) for i in range(3)])

    def slabs_ragged(frame):
        z = np.asarray(frame)[:, 2] % L
        out = np.empty(3, dtype=object)
        out[0] = np.where(z < 0.5 * L)[0]
        out[1] = np.where(z > 2 * L)[0]
        out[2] = np.where(z >= 0.5 * L)[0]
        return out
    return slab, slabs_masks, slabs_ragged


def test_subensemble_golden(golden):
    g = golden('subensemble')
    unw = orc.UnwrappedTrajectory(g['pos'], g['box'], g['times'])
    slab, slabs_masks, slabs_ragged = _selectors(float(g['box'][0]))
    fns = {'msd': orc.msd, 'isf': partial(orc.isf, q=22.7), 'ng': orc.non_gaussian,
           'vh': partial(orc.van_hove_self, bins=g['bins'])}
    with np.errstate(all='ignore'):
        for name, fn in fns.items():
            t, r = orc.shifted_correlation(fn, unw, segments=9, window=0.4,
                                           correlation=orc.subensemble_pairwise(slab))
            same(r, g['sub_' + name])
            t, r = orc.shifted_correlation(fn, unw, segments=9, window=0.4, average=True,
                                           correlation=orc.subensemble_pairwise(partial(slab, lo=0.5, hi=0.9)))
            same(r, g['sub_avg_' + name])
            for tag, sel in (('masks', slabs_masks), ('ragged', slabs_ragged)):
                if name == 'vh' and tag == 'ragged':
                    continue
                t, r, cnt = orc.shifted_correlation(fn, unw, segments=9, window=0.4,
                                                    correlation=orc.multi_subensemble_pairwise(sel))
                same(r, g['multi_%s_%s' % (tag, name)])
                same(cnt, g['multi_%s_count' % tag])
                t, r, cnt = orc.shifted_correlation(fn, unw, segments=9, window=0.4, average=True,
                                                    correlation=orc.multi_subensemble_pairwise(sel))
                same(r, g['multi_avg_%s_%s' % (tag, name)])
                same(cnt, g['multi_avg_%s_count' % tag])
    same(t, g['sub_t'])


def test_rotational_autocorrelation_golden(golden):
    g = golden('rotational')
    for tag in ('f64', 'f32'):
        frames = orc.ArrayTrajectory(g['vec_' + tag], np.ones(3) * 100, np.arange(150) * 2.0)
        for order in (1, 2, 4):
            t, r = orc.shifted_correlation(partial(orc.rotational_autocorrelation, order=order), frames, segments=8,
                                           window=0.5)
            same(r, g['rot%d_%s' % (order, tag)])
        same(t, g['rot_t'])


def _half_selector(frame):
    return np.where(np.asarray(frame)[:, 0] < 0.75)[0]


def test_overlap_golden(golden):
    """correlation.overlap on a CoordinatesKDTree (correlation.py:288-312, coordinates.py:415-455): the oracle's brute
    force restatement of cKDTree's nearest distance equals the reference + scipy bit for bit -- open and periodic
    trees, a tree of a selection, nojump frames."""
    g = golden('overlap')
    pos, box, times = g['pos'], g['box'], g['times']
    raw = orc.ArrayTrajectory(pos, box.astype(np.float64), times)
    kw = dict(segments=8, window=0.4)
    np.testing.assert_array_equal(orc.nearest_distances(raw[3], raw[90]), g['dist_open'])
    np.testing.assert_array_equal(orc.nearest_distances(raw[3], raw[90], float(box[0])), g['dist_periodic'])
    for radius in (0.03, 0.06):
        t, r = orc.shifted_correlation(partial(orc.overlap, crds_tree=orc.FrameTrees(raw), radius=radius), raw, **kw)
        np.testing.assert_array_equal(t, g['t'])
        np.testing.assert_array_equal(r, g['open_%g' % radius])
        trees = orc.FrameTrees(raw, boxsize=float(box[0]))
        _, r = orc.shifted_correlation(partial(orc.overlap, crds_tree=trees, radius=radius), raw, **kw)
        np.testing.assert_array_equal(r, g['periodic_%g' % radius])
    trees = orc.FrameTrees(raw, selector=_half_selector, boxsize=float(box[0]))
    _, r = orc.shifted_correlation(partial(orc.overlap, crds_tree=trees, radius=0.05), raw, **kw)
    np.testing.assert_array_equal(r, g['selector_0.05'])
    unw = orc.UnwrappedTrajectory(pos, box, times)
    _, r = orc.shifted_correlation(partial(orc.overlap, crds_tree=orc.FrameTrees(unw), radius=0.05), unw, average=True,
                                   **kw)
    np.testing.assert_array_equal(r, g['nojump_avg_0.05'])


def test_nearest_distances_equal_scipy():
    """The restatement against scipy's cKDTree itself (the third-party code the reference calls), where scipy is
    installed: points outside a periodic box, a box periodic in two dimensions only, float32 input."""
    spatial = pytest.importorskip('scipy.spatial')
    rng = np.random.default_rng(8)
    for n, length, box in ((1500, 2.0, None), (1500, 2.0, 2.0), (600, 1.0, np.array([1.0, 0.0, 1.0]))):
        pts = (rng.random((n, 3)) * length).astype(np.float32)
        qry = (pts + rng.normal(0, 0.05, (n, 3))).astype(np.float32)
        want = spatial.cKDTree(pts, boxsize=box).query(qry)[0]
        np.testing.assert_array_equal(orc.nearest_distances(pts, qry, box), want)
