"""Parity of the CUDA path with the CPU oracle / the reference's golden vectors (needs a GPU).

Bars (BASELINE.json north_star):
* integer image shifts: bit-exact;
* van Hove counts: exact (the kernel replays the reference's float64 / float32 operations,
  so not even bin-edge ties are allowed to move here);
* scalar correlations: |a - b| <= 1e-5 |ref| + atol, atol = 1e-7 * scale of the quantity
  (the oracle's own float32 <-> float64 spread is 1e-7 abs on the ISF).
  In nojump mode the displacement arithmetic is float64 like the oracle's, so MSD / alpha_2
  are additionally held to 1e-11.
"""
from functools import partial

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import md_oracle as orc

RTOL = 1e-5


def close(ours, ref, scale=1.0, rtol=RTOL, atol=1e-7):
    ours = np.asarray(ours, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert ours.shape == ref.shape
    nan = np.isnan(ref)
    assert (np.isnan(ours) == nan).all()
    err = np.abs(ours - ref)[~nan]
    tol = (rtol * np.abs(ref) + atol * scale)[~nan]
    assert (err <= tol).all(), 'max excess %.3e (max err %.3e)' % ((err - tol).max(), err.max())


@pytest.fixture(scope='module')
def md():
    import mdevaluate_b200
    return mdevaluate_b200


@pytest.fixture(scope='module')
def small(golden, md):
    g = golden('walk_small')
    coords = md.Coordinates.from_arrays(g['pos'], g['box'], g['times'])
    return g, coords


def _fns(md, g):
    c = md.correlation
    return {
        'msd': c.msd,
        'isf': partial(c.isf, q=22.7),
        'ng': c.non_gaussian,
        'vh': partial(c.van_hove_self, bins=g['bins']),
        'vh_off': partial(c.van_hove_self, bins=g['bins_off']),
        'vh_irr': partial(c.van_hove_self, bins=g['bins_irr']),
    }


def test_image_counts_bit_exact(small, md):
    g, coords = small
    res = coords._resident()
    jumps = res.dense_jumps()
    assert jumps.dtype == np.int8
    np.testing.assert_array_equal(jumps, g['jumps'])
    store = res.store_for(np.arange(len(coords)), nojump=True)
    n3 = 3 * res.n_atoms
    images = store.images[:, :n3].cpu().numpy().astype(np.int64).reshape(len(coords), -1, 3) - 128
    np.testing.assert_array_equal(images, np.cumsum(g['jumps'], axis=0, dtype=np.int64))


def test_parse_jumps_api(small, md):
    g, coords = small
    from scipy import sparse
    md.reader.generate_nojump_matrixes(coords)
    mats = coords.frames.nojump_matrixes
    assert len(mats) == 3 and all(sparse.issparse(m) for m in mats)
    dense = np.stack([m.toarray() for m in mats], axis=-1)
    np.testing.assert_array_equal(dense, g['jumps'])


def test_unwrapped_frames_bit_exact(small, md):
    g, coords = small
    for i in (0, 1, 17, 100, 239):
        f = coords.nojump[i]
        assert isinstance(f, md.CoordinateFrame) and f.dtype == np.float64 and f.mode == 'nojump'
        np.testing.assert_array_equal(np.asarray(f), g['nojump_%d' % i])
        assert f.time == g['times'][i] and f.step == i


@pytest.mark.parametrize('mode', ['nojump', 'raw'])
@pytest.mark.parametrize('name', ['msd', 'isf', 'ng'])
def test_scalar_correlations_golden(small, md, name, mode):
    g, coords = small
    frames = coords.nojump if mode == 'nojump' else coords
    t, r = md.correlation.shifted_correlation(_fns(md, g)[name], frames, segments=12, window=0.4)
    ref = g['%s_%s' % (name, mode)]
    np.testing.assert_array_equal(t, g['t_' + mode])
    assert r.dtype == ref.dtype
    scale = {'msd': float(np.nanmax(g['msd_nojump'])), 'isf': 1.0, 'ng': 1.0}[name]
    # raw float32 frames: the reference itself computes in float32, alpha_2 = O(1) ratio of float32 means
    close(r, ref, scale=scale, atol=1e-6 if (mode == 'raw' and name == 'ng') else 1e-7)
    if mode == 'nojump' and name != 'isf':
        close(r, ref, scale=scale, rtol=1e-11, atol=1e-13)


@pytest.mark.parametrize('mode', ['nojump', 'raw'])
@pytest.mark.parametrize('name', ['vh', 'vh_off', 'vh_irr'])
def test_van_hove_exact(small, md, name, mode):
    g, coords = small
    frames = coords.nojump if mode == 'nojump' else coords
    t, r = md.correlation.shifted_correlation(_fns(md, g)[name], frames, segments=12, window=0.4)
    ref = g['%s_%s' % (name, mode)]
    assert r.dtype == np.float64 and r.shape == ref.shape
    n = g['pos'].shape[1]
    np.testing.assert_array_equal(np.rint(r * n).astype(np.int64), np.rint(ref * n).astype(np.int64))
    np.testing.assert_array_equal(r, ref)


def test_chi4_average_slice_dense(small, md):
    g, coords = small
    c = md.correlation
    t, fqt = c.shifted_correlation(partial(c.isf, q=22.7), coords.nojump, segments=20, window=0.2, skip=0.1,
                                   average=False)
    np.testing.assert_array_equal(t, g['chi4_t'])
    close(fqt, g['chi4_fqt'])
    close(len(coords[0]) * fqt.var(axis=0), g['chi4'], atol=1e-6)
    t2, x4 = c.chi4(coords.nojump, 22.7)
    np.testing.assert_array_equal(x4, len(coords[0]) * fqt.var(axis=0))
    t, r = c.shifted_correlation(c.msd, coords.nojump, segments=7, window=0.3, average=True)
    np.testing.assert_array_equal(t, g['msd_avg_t'])
    close(r, g['msd_avg'], rtol=1e-11, atol=1e-13)
    t, r = c.shifted_correlation(c.msd, coords.nojump[24:], segments=7, window=0.3, average=True)
    close(r, g['msd_slice24'], rtol=1e-11, atol=1e-13)
    t, r = c.shifted_correlation(partial(c.isf, q=22.7), coords.nojump, segments=5, window=0.25,
                                 index_distribution=lambda a, b: np.arange(a, b))
    close(r, g['isf_dense'])


def test_selection(small, md):
    g, _ = small
    coords = md.Coordinates(md.MemoryReader(g['pos'], g['box'], g['times']), atom_filter=g['sel'])
    t, r = md.correlation.shifted_correlation(md.correlation.msd, coords.nojump, segments=6, window=0.4)
    close(r, g['msd_sel'], rtol=1e-11, atol=1e-13)


def test_config1_golden(golden, md):
    """BASELINE config 1 against the reference's own output: MSD 1000 x 1000, S=100, L=70."""
    from mdevaluate_b200.synthetic import random_walk
    g = golden('config1')
    pos, box, times = random_walk(1000, 1000, seed=20240)
    coords = md.Coordinates.from_arrays(pos, box, times)
    c = md.correlation
    t, r = c.shifted_correlation(c.msd, coords.nojump, segments=100, window=0.5)
    assert r.shape == (100, 70)
    np.testing.assert_array_equal(t, g['t'])
    close(r, g['msd'], scale=float(g['msd'].max()), rtol=1e-11, atol=1e-13)
    t, r = c.shifted_correlation(partial(c.isf, q=22.7), coords.nojump, segments=100, window=0.5)
    close(r, g['isf'])
    print('config1 isf max abs err %.3e' % np.abs(r - g['isf']).max())
    t, r = c.shifted_correlation(c.non_gaussian, coords.nojump, segments=100, window=0.5)
    close(r, g['ng'], rtol=1e-10, atol=1e-12)
    t, r = c.shifted_correlation(partial(c.isf, q=22.7), coords, segments=100, window=0.5)
    assert r.dtype == np.float32
    close(r, g['isf_raw'], atol=2e-7)
    t, r = c.shifted_correlation(partial(c.van_hove_self, bins=g['vh_bins']), coords.nojump, segments=10,
                                 window=0.5)
    np.testing.assert_array_equal(r, g['vh'])
    t, r = c.shifted_correlation(partial(c.van_hove_self, bins=g['vh_bins']), coords, segments=10, window=0.5)
    np.testing.assert_array_equal(r, g['vh_raw'])
    # dense jump statistics of the whole trajectory
    res = coords._resident()
    store = res.store_for(np.arange(1000), nojump=True)
    last = store.images[999, :3000].cpu().numpy().reshape(1000, 3).astype(np.int64) - 128
    np.testing.assert_array_equal(last, g['count_checksum'])


def test_multi_function_single_pass(small, md):
    g, coords = small
    c = md.correlation
    fns = _fns(md, g)
    t, out = c.shifted_correlation_multi({'vh': fns['vh'], 'ng': fns['ng'], 'msd': fns['msd'], 'isf': fns['isf']},
                                         coords.nojump, segments=12, window=0.4)
    np.testing.assert_array_equal(out['vh'], g['vh_nojump'])
    close(out['ng'], g['ng_nojump'], rtol=1e-11, atol=1e-13)
    close(out['msd'], g['msd_nojump'], rtol=1e-11, atol=1e-13)
    close(out['isf'], g['isf_nojump'])


def test_direct_pair_calls_and_histogram(golden, md):
    g = golden('walk_small')
    a = g['nojump_17']
    b = g['nojump_100']
    c = md.correlation
    with np.errstate(all='ignore'):
        close(c.msd(a, b), orc.msd(a, b), rtol=1e-12, atol=0)
        close(c.isf(a, b, q=22.7), orc.isf(a, b, 22.7))
        close(c.non_gaussian(a, b), orc.non_gaussian(a, b), rtol=1e-11)
        np.testing.assert_array_equal(c.van_hove_self(a, b, bins=g['bins']), orc.van_hove_self(a, b, g['bins']))
        assert np.isnan(c.non_gaussian(a, a))
        assert c.isf(a, a, q=22.7) == 1.0 and c.msd(a, a) == 0.0
    h = golden('histogram')
    for k in ('b0', 'b1', 'b2'):
        hist, centres = md.utils.histogram(h['data'], h[k])
        np.testing.assert_array_equal(hist, h['h_' + k])
        np.testing.assert_array_equal(centres, h['c_' + k])


def test_generic_frames_protocol(golden, md):
    """Any object with len / [i] -> (N,3) array with .time works (float64 unwrapped frames here)."""
    g = golden('walk_small')
    unw = orc.UnwrappedTrajectory(g['pos'], g['box'], g['times'])
    c = md.correlation
    t, r = c.shifted_correlation(c.msd, unw, segments=12, window=0.4)
    np.testing.assert_array_equal(t, g['t_nojump'])
    close(r, g['msd_nojump'], rtol=1e-11, atol=1e-13)
    t, r = c.shifted_correlation(partial(c.van_hove_self, bins=g['bins']), unw, segments=12, window=0.4)
    np.testing.assert_array_equal(r, g['vh_nojump'])


def test_ragged_sizes_and_streaming(md):
    """Atom counts that are not multiples of 4, unaligned rows, the compact/streamed store."""
    from mdevaluate_b200.synthetic import random_walk
    c = md.correlation
    for n in (1, 3, 5, 31, 130):
        pos, box, times = random_walk(n, 60, seed=n, sigma=0.2, box=2.0)
        unw = orc.UnwrappedTrajectory(pos, box, times)
        for forced in (False, True):
            coords = md.Coordinates.from_arrays(pos, box, times)
            coords._resident().force_compact = forced
            coords._resident().SEGMENT_BYTES = 4096          # many tiny segments: carried scan state
            with np.errstate(all='ignore'):
                t0, ref = orc.shifted_correlation(orc.msd, unw, segments=5, window=0.5)
                t, r = c.shifted_correlation(c.msd, coords.nojump, segments=5, window=0.5)
                close(r, ref, rtol=1e-11, atol=1e-13)
                bins = np.linspace(0, 3.0, 41)
                _, ref = orc.shifted_correlation(partial(orc.van_hove_self, bins=bins), unw, segments=5, window=0.5)
                _, r = c.shifted_correlation(partial(c.van_hove_self, bins=bins), coords.nojump, segments=5,
                                             window=0.5)
                np.testing.assert_array_equal(r, ref)


def test_errors(md, small):
    g, coords = small
    c = md.correlation
    with pytest.raises(AssertionError):
        c.shifted_correlation(c.msd, coords, window=0.6, skip=0.5)
    with pytest.raises(NotImplementedError):
        c.shifted_correlation(lambda a, b: 0.0, coords)
    with pytest.raises(TypeError):
        c.shifted_correlation(c.isf, coords)
    with pytest.raises(IndexError):
        coords[len(coords)]
    with pytest.raises(IndexError):
        c.shifted_correlation(c.msd, coords, index_distribution=lambda a, b: np.array([0, 10 * b]))
    # autosave keyword arguments of the reference decorator are accepted and ignored
    c.shifted_correlation(c.msd, coords, description='x', autoload=False)


def test_c_abi_argument_checks(md):
    import ctypes as C
    from mdevaluate_b200._cuda import lib as L
    lib = L.load()
    a = L.CorrArgs()
    assert lib.mdb_shifted_corr(C.byref(a), None) == 1
    assert b'positive' in lib.mdb_last_error()
    sm, major, minor = C.c_int(), C.c_int(), C.c_int()
    assert lib.mdb_device_info(C.byref(sm), C.byref(major), C.byref(minor)) == 0
    assert sm.value > 0 and major.value >= 10


def test_per_frame_box(md):
    """NPT-like trajectory: the box changes every frame.  Jumps are detected with frame 0's box
    (reader.py:167), each frame is unwrapped with its own box (pbc.py:237, cache-free definition)."""
    from mdevaluate_b200.synthetic import random_walk
    pos, box, times = random_walk(203, 90, seed=11, sigma=0.15, box=2.5)
    rng = np.random.default_rng(1)
    boxes = (box[None, :] * (1 + 0.002 * rng.standard_normal((90, 3)))).astype(np.float32)
    boxes[0] = box
    unw = orc.UnwrappedTrajectory(pos, boxes, times)
    coords = md.Coordinates.from_arrays(pos, boxes, times)
    c = md.correlation
    np.testing.assert_array_equal(np.asarray(coords.nojump[57]), np.asarray(unw[57]))
    with np.errstate(all='ignore'):
        for ours, theirs in ((c.msd, orc.msd), (c.non_gaussian, orc.non_gaussian)):
            _, ref = orc.shifted_correlation(theirs, unw, segments=6, window=0.5)
            _, r = c.shifted_correlation(ours, coords.nojump, segments=6, window=0.5)
            close(r, ref, rtol=1e-11, atol=1e-13)
        bins = np.linspace(0, 3.0, 61)
        _, ref = orc.shifted_correlation(partial(orc.van_hove_self, bins=bins), unw, segments=6, window=0.5)
        _, r = c.shifted_correlation(partial(c.van_hove_self, bins=bins), coords.nojump, segments=6, window=0.5)
        np.testing.assert_array_equal(r, ref)
        _, ref = orc.shifted_correlation(partial(orc.isf, q=22.7), unw, segments=6, window=0.5)
        _, r = c.shifted_correlation(partial(c.isf, q=22.7), coords.nojump, segments=6, window=0.5)
        close(r, ref)


# ---- sub-ensembles: SURVEY.md section 8f-1 (correlation.py:27-81, 153-189) ---------------------------
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


def _check(name, got, ref):
    if name == 'vh':
        np.testing.assert_array_equal(got, ref)
    elif name == 'isf':
        close(got, ref)
    else:
        close(got, ref, rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize('name', ['msd', 'isf', 'ng', 'vh'])
def test_subensemble_correlation_golden(golden, md, name):
    g = golden('subensemble')
    coords = md.Coordinates.from_arrays(g['pos'], g['box'], g['times'])
    c = md.correlation
    slab, slabs_masks, slabs_ragged = _selectors(float(g['box'][0]))
    fn = {'msd': c.msd, 'isf': partial(c.isf, q=22.7), 'ng': c.non_gaussian,
          'vh': partial(c.van_hove_self, bins=g['bins'])}[name]
    t, r = c.shifted_correlation(fn, coords.nojump, segments=9, window=0.4,
                                 correlation=c.subensemble_correlation(slab))
    np.testing.assert_array_equal(t, g['sub_t'])
    _check(name, r, g['sub_' + name])
    t, r = c.shifted_correlation(fn, coords.nojump, segments=9, window=0.4, average=True,
                                 correlation=c.subensemble_correlation(partial(slab, lo=0.5, hi=0.9)))
    _check(name, r, g['sub_avg_' + name])
    for tag, sel in (('masks', slabs_masks), ('ragged', slabs_ragged)):
        if name == 'vh' and tag == 'ragged':
            continue
        t, r, cnt = c.shifted_correlation(fn, coords.nojump, segments=9, window=0.4,
                                          correlation=c.multi_subensemble_correlation(sel))
        np.testing.assert_array_equal(cnt, g['multi_%s_count' % tag])
        assert r.shape == g['multi_%s_%s' % (tag, name)].shape
        _check(name, r, g['multi_%s_%s' % (tag, name)])
        t, r, cnt = c.shifted_correlation(fn, coords.nojump, segments=9, window=0.4, average=True,
                                          correlation=c.multi_subensemble_correlation(sel))
        np.testing.assert_array_equal(cnt, g['multi_avg_%s_count' % tag])
        if name == 'vh':
            np.testing.assert_allclose(r, g['multi_avg_%s_%s' % (tag, name)], rtol=1e-13, atol=0)
        else:
            _check(name, r, g['multi_avg_%s_%s' % (tag, name)])


def test_spatial_selector_and_strategy_objects(golden, md):
    g = golden('subensemble')
    coords = md.Coordinates.from_arrays(g['pos'], g['box'], g['times'])
    sel = md.coordinates.spatial_selector(coords[5], lambda f: np.asarray(f)[:, 0], 0.2, 0.9)
    np.testing.assert_array_equal(sel[0], g['spatial_selector_5'][0])
    c = md.correlation
    # the strategy objects also work stand-alone, pair by pair, like the reference's generators
    slab = _selectors(float(g['box'][0]))[0]
    idx = [3, 4, 9]
    vals = list(c.subensemble_correlation(slab)(c.msd, (coords.nojump[i] for i in idx)))
    unw = orc.UnwrappedTrajectory(g['pos'], g['box'], g['times'])
    ref = list(orc.subensemble_pairwise(slab)(orc.msd, (unw[i] for i in idx)))
    np.testing.assert_allclose(vals, ref, rtol=1e-12)
    with pytest.raises(NotImplementedError):
        c.shifted_correlation(c.msd, coords, correlation=c.subensemble_correlation(lambda f: [0, 0, 1]))


def test_rotational_autocorrelation_golden(golden, md):
    """P_l of unit-vector scalar products (correlation.py:223-236)."""
    g = golden('rotational')
    c = md.correlation
    for tag, tol in (('f64', 1e-12), ('f32', 2e-6)):
        coords = md.Coordinates.from_arrays(g['vec_' + tag], np.ones(3) * 100, np.arange(150) * 2.0)
        for order in (1, 2, 4):
            t, r = c.shifted_correlation(partial(c.rotational_autocorrelation, order=order), coords, segments=8,
                                         window=0.5)
            ref = g['rot%d_%s' % (order, tag)]
            assert r.dtype == ref.dtype
            np.testing.assert_array_equal(t, g['rot_t'])
            np.testing.assert_allclose(r, ref, rtol=tol, atol=tol)
    a, b = g['vec_f64'][3], g['vec_f64'][40]
    np.testing.assert_allclose(c.rotational_autocorrelation(a, b, order=2), orc.rotational_autocorrelation(a, b, 2),
                               rtol=1e-12)
    np.testing.assert_allclose(c.rotational_autocorrelation(a, a, order=3), 1.0, rtol=1e-12)


def test_whole0_reference_frame(md):
    """Row 0 of the jump table is taken against frame0.whole (reader.py:166): molecules that were
    wrapped apart in frame 0 start with a non-zero image count."""
    from mdevaluate_b200.synthetic import random_walk
    pos, box, times = random_walk(57, 80, seed=2, sigma=0.1, box=2.0)
    whole0 = pos[0].copy()
    whole0[::5, 0] += box[0]            # these atoms sit one box further in the whole molecule
    whole0[3::7, 2] -= box[2]
    unw = orc.UnwrappedTrajectory(pos, box, times, whole0=whole0)
    coords = md.Coordinates(md.MemoryReader(pos, box, times, whole0=whole0))
    store = coords._resident().store_for(np.arange(80), nojump=True)
    images = store.images[:, :3 * 57].cpu().numpy().astype(np.int64).reshape(80, 57, 3) - 128
    np.testing.assert_array_equal(images, unw.counts)
    assert np.count_nonzero(images[0]) > 0
    np.testing.assert_array_equal(np.asarray(coords.nojump[33]), np.asarray(unw[33]))
    c = md.correlation
    _, ref = orc.shifted_correlation(orc.msd, unw, segments=4)
    _, r = c.shifted_correlation(c.msd, coords.nojump, segments=4)
    close(r, ref, rtol=1e-11, atol=1e-13)


def test_origin_batching_for_dense_schedules(small, md, monkeypatch):
    """Dense schedules have S*L pairs far beyond one launch's buffers: origins are batched."""
    g, coords = small
    c = md.correlation
    dense = lambda a, b: np.arange(a, b)
    kw = dict(segments=60, window=0.3, index_distribution=dense)
    t0, full = c.shifted_correlation(partial(c.isf, q=22.7), coords.nojump, **kw)
    t0, full_avg = c.shifted_correlation(partial(c.isf, q=22.7), coords.nojump, average=True, **kw)
    monkeypatch.setattr(c, 'MAX_PAIRS_PER_LAUNCH', 7 * full.shape[1])          # 7 origins per launch
    t1, batched = c.shifted_correlation(partial(c.isf, q=22.7), coords.nojump, **kw)
    t1, batched_avg = c.shifted_correlation(partial(c.isf, q=22.7), coords.nojump, average=True, **kw)
    np.testing.assert_array_equal(batched, full)
    np.testing.assert_array_equal(batched_avg, full_avg)
    unw = orc.UnwrappedTrajectory(g['pos'], g['box'], g['times'])
    _, ref = orc.shifted_correlation(partial(orc.isf, q=22.7), unw, segments=60, window=0.3, average=True,
                                     index_distribution=orc.lin_indices)
    close(batched_avg, ref)


# ---- dense schedules: the tiled kernel (csrc/dense_kernels.cu), BASELINE config 5 -------------------
@pytest.mark.parametrize('n_atoms,n_frames,segments,window,first', [
    (96, 240, 120, 0.5, 0),      # every frame an origin (step 1), 120 lags
    (131, 200, 50, 0.5, 0),      # origin step 2, atoms not a multiple of 64
    (70, 180, 30, 0.35, 3),      # origin step 3-4 is irregular here -> falls back unless arithmetic
    (300, 150, 75, 0.5, 2),      # lags start at 2
    (65, 240, 30, 0.5, 0),       # origin step 4 (9 origins per block)
    (200, 180, 30, 0.5, 1),      # origin step 3
    (64, 400, 200, 0.5, 0),      # 200 lags = 6 full lag blocks + 8
])
def test_dense_schedule_kernel(md, monkeypatch, n_atoms, n_frames, segments, window, first):
    from mdevaluate_b200.synthetic import random_walk
    from mdevaluate_b200._cuda import engine
    pos, box, times = random_walk(n_atoms, n_frames, seed=n_atoms, sigma=0.08, box=1.7)
    coords = md.Coordinates.from_arrays(pos, box, times)
    unw = orc.UnwrappedTrajectory(pos, box, times)
    c = md.correlation
    dense = lambda a, b: np.arange(a + first, b)
    calls = []
    real = engine.dense_device
    monkeypatch.setattr(engine, 'dense_device', lambda *a, **k: (calls.append(1), real(*a, **k))[1])
    monkeypatch.setattr(c, 'DENSE_MIN_PAIRS', 1)
    for name, ours, theirs in (('msd', c.msd, orc.msd), ('isf', partial(c.isf, q=22.7), partial(orc.isf, q=22.7))):
        for mode, frames, ref_frames in (('nojump', coords.nojump, unw),
                                         ('raw', coords, orc.ArrayTrajectory(pos, box, times))):
            n0 = len(calls)
            t, got = c.shifted_correlation(ours, frames, segments=segments, window=window, average=True,
                                           index_distribution=dense)
            starts, _ = c._schedule(frames, dense, segments, window, None)
            arithmetic = len(starts) < 2 or (np.all(np.diff(starts) == np.diff(starts)[0]) and np.diff(starts)[0] <= 4)
            assert (len(calls) > n0) == bool(arithmetic)           # the tiled kernel really ran (or really not)
            t0, ref = orc.shifted_correlation(theirs, ref_frames, segments=segments, window=window, average=True,
                                              index_distribution=dense)
            np.testing.assert_array_equal(t, t0)
            assert got.dtype == ref.dtype and got.shape == ref.shape
            # float32 offsets in shared memory (dense_corr_kernel_v2): every displacement good to ~1e-7 relative
            close(got, ref, scale=float(np.abs(ref).max()), rtol=2e-6, atol=2e-7 if mode == 'raw' else 1e-7)
            if name == 'msd' and mode == 'nojump':
                engine.tuning.dense_float64 = 1            # the float64 tile: the reference's arithmetic
                try:
                    _, got64 = c.shifted_correlation(ours, frames, segments=segments, window=window, average=True,
                                                     index_distribution=dense)
                finally:
                    engine.tuning.dense_float64 = 0
                close(got64, ref, rtol=1e-11, atol=1e-13)
            # and it agrees with the streaming kernel on the same schedule
            monkeypatch.setattr(c, 'DENSE_MIN_PAIRS', 1 << 60)
            _, stream = c.shifted_correlation(ours, frames, segments=segments, window=window, average=True,
                                              index_distribution=dense)
            monkeypatch.setattr(c, 'DENSE_MIN_PAIRS', 1)
            close(got, stream, scale=float(np.abs(ref).max()), rtol=2e-6, atol=2e-7 if name == 'isf' else 1e-12)


def test_dense_kernel_on_a_compact_store(md, monkeypatch):
    """The tiled kernel follows the frame -> row map of a compact (streamed) store."""
    from mdevaluate_b200.synthetic import random_walk
    pos, box, times = random_walk(90, 300, seed=4, sigma=0.08, box=1.7)
    c = md.correlation
    monkeypatch.setattr(c, 'DENSE_MIN_PAIRS', 1)
    dense = lambda a, b: np.arange(a, b)
    unw = orc.UnwrappedTrajectory(pos, box, times)
    kw = dict(segments=40, window=0.3, skip=0.2, average=True, index_distribution=dense)
    _, ref = orc.shifted_correlation(partial(orc.isf, q=22.7), unw, **kw)
    coords = md.Coordinates.from_arrays(pos, box, times)
    res = coords._resident()
    res.force_compact = True
    res.SEGMENT_BYTES = 16384
    _, got = c.shifted_correlation(partial(c.isf, q=22.7), coords.nojump, **kw)
    store = next(iter(res._compact.values()))
    assert store.frame_to_row is not None and store.coords.shape[0] < 300        # really a compact store
    close(got, ref)


@pytest.mark.parametrize('source', ['host', 'pinned', 'device'])
def test_dense_schedule_in_atom_slabs(md, monkeypatch, source):
    """Out-of-core dense pass (BASELINE config 5 is larger than HBM): atom slabs whose rows stream from the
    host while the previous slab's kernel runs give the sums of the single pass."""
    import torch
    from mdevaluate_b200.synthetic import random_walk
    from mdevaluate_b200._cuda import engine
    N, T = 2500, 200
    pos, box, times = random_walk(N, T, seed=9, sigma=0.08, box=2.5)
    c = md.correlation
    dense = lambda a, b: np.arange(a, b)
    kw = dict(segments=T // 2, window=0.5, average=True, index_distribution=dense)

    def make():
        if source == 'host':
            return md.Coordinates.from_arrays(pos, box, times)
        if source == 'pinned':
            pinned = torch.from_numpy(pos).pin_memory()
            return md.Coordinates.from_arrays(pinned.numpy(), box, times)
        pitch = engine.row_pitch(N)
        block = torch.zeros((T, pitch), dtype=torch.float32, device='cuda')
        block[:, :3 * N] = torch.from_numpy(pos.reshape(T, 3 * N)).cuda()
        return md.Coordinates.from_arrays(block, box, times, n_atoms=N)

    calls = []
    real = engine.dense_device
    monkeypatch.setattr(engine, 'dense_device', lambda *a, **k: (calls.append(a[0].n_atoms), real(*a, **k))[1])
    unw = orc.UnwrappedTrajectory(pos, box, times)
    _, ref = orc.shifted_correlation(orc.msd, unw, **kw)
    for name, fn in (('msd', c.msd), ('isf', partial(c.isf, q=22.7))):
        for mode in ('nojump', 'raw'):
            coords = make()
            frames = coords.nojump if mode == 'nojump' else coords
            del calls[:]
            _, one = c.shifted_correlation(fn, frames, **kw)
            assert calls == [N]
            monkeypatch.setattr(c, 'DENSE_SLAB_ATOMS', 1024)
            coords = make()                                     # nothing resident yet: the slabs load themselves
            frames = coords.nojump if mode == 'nojump' else coords
            del calls[:]
            t, got = c.shifted_correlation(fn, frames, **kw)
            monkeypatch.setattr(c, 'DENSE_SLAB_ATOMS', None)
            assert calls == [1024, 1024, N - 2048]
            assert got.dtype == one.dtype
            close(got, one, scale=float(np.abs(one).max()), rtol=1e-6 if mode == 'raw' else 1e-12,
                  atol=1e-6 if mode == 'raw' else 1e-13)
            if name == 'msd' and mode == 'nojump':
                close(got, ref, rtol=2e-6, atol=1e-12)           # float32 offsets (dense_corr_kernel_v2)


def test_custom_reader_protocol(golden, md):
    """Coordinates over ANY reader with the reference protocol (len, reader[i].positions/.box/.time,
    reader.py:254-286): frames are pulled one by one into pinned staging and streamed through the scan."""
    g = golden('walk_small')
    pos, box, times = g['pos'], g['box'], g['times']          # NpzFile decompresses on every access

    class Frame:
        def __init__(self, positions, box, time):
            self.positions, self.box, self.time = positions, box, time
            self.coordinates = positions

    class Reader:
        filename = 'custom'

        def __len__(self):
            return len(pos)

        def __getitem__(self, i):
            return Frame(pos[i], np.diag(box), times[i])

    coords = md.Coordinates(Reader())
    c = md.correlation
    for forced in (False, True):
        coords = md.Coordinates(Reader())
        coords._resident().force_compact = forced
        t, r = c.shifted_correlation(c.msd, coords.nojump, segments=12, window=0.4)
        np.testing.assert_array_equal(t, g['t_nojump'])
        close(r, g['msd_nojump'], rtol=1e-11, atol=1e-13)
        t, r = c.shifted_correlation(partial(c.van_hove_self, bins=g['bins']), coords, segments=12, window=0.4)
        np.testing.assert_array_equal(r, g['vh_raw'])
    np.testing.assert_array_equal(np.asarray(coords.nojump[100]), g['nojump_100'])


def test_store_released_without_garbage_collector(small, md):
    """The HBM store of a trajectory goes away with the last Coordinates by reference counting alone
    (no reader <-> store cycle): a streamed step must not leave GBs behind for the garbage collector."""
    import gc
    import weakref
    g, _ = small
    gc.collect()
    gc.disable()
    try:
        coords = md.Coordinates.from_arrays(g['pos'], g['box'], g['times'])
        md.correlation.shifted_correlation(md.correlation.msd, coords.nojump, segments=4)
        res = weakref.ref(coords._resident())
        store = weakref.ref(coords._resident().store_for(np.arange(4), nojump=True))
        del coords
        assert res() is None and store() is None
    finally:
        gc.enable()


# ---- stored jump matrices (the reference's .nojump.npz cache) and autosave -----------------------------
class FileLikeReader:
    """Frames only through ``reader[i]`` (like a trajectory file), counting the reads."""

    def __init__(self, md, g, filename='<file-like>'):
        self._md, self._g = md, g
        self.filename = filename
        self.reads = []
        self.nojump_matrixes = None

    def __len__(self):
        return len(self._g['pos'])

    def __getitem__(self, i):
        i = int(i)
        self.reads.append(i)
        box = self._g['box']
        return self._md.reader.MemoryFrame(self._g['pos'][i], np.diag(box).astype(box.dtype), self._g['times'][i], i)


def _golden_file(name):
    import os
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', name)


def _ref_matrices(md):
    return md.reader.read_nojump_file_matrices(_golden_file('ref.nojump.npz'))[1]


def test_stored_jump_matrices_replace_the_scan(small, md):
    """With the reference's jump-matrix cache attached only the frames of the schedule are read, and the
    image counts / unwrapped frames / correlations are those of the scan (bit-exact integers)."""
    g, _ = small
    reader = FileLikeReader(md, g)
    reader.nojump_matrixes = _ref_matrices(md)
    coords = md.Coordinates(reader)
    T = len(coords)
    t, r = md.correlation.shifted_correlation(md.correlation.msd, coords.nojump, segments=12, window=0.4)
    np.testing.assert_array_equal(t, g['t_nojump'])
    close(r, g['msd_nojump'], scale=float(np.nanmax(g['msd_nojump'])), rtol=1e-11, atol=1e-13)
    starts, idx = md.correlation._schedule(coords, md.correlation.log_indices, 12, 0.4, None)
    needed = set((starts[:, None] + idx[None, :]).ravel().tolist())
    assert set(reader.reads) <= needed | {0} and len(set(reader.reads)) < T
    # image counts of the kept frames
    res = coords._resident()
    keep = np.array(sorted(needed))
    store = res.store_for(keep, nojump=True)
    rows = store.rows_for(keep).astype(np.int64)
    n3 = 3 * res.n_atoms
    images = store.images[:, :n3].cpu().numpy().astype(np.int64)[rows].reshape(len(keep), -1, 3) - 128
    np.testing.assert_array_equal(images, np.cumsum(g['jumps'], axis=0, dtype=np.int64)[keep])
    # unwrapped frames, also one that is not part of the schedule's store
    for i in (0, 17, 100, 239):
        np.testing.assert_array_equal(np.asarray(coords.nojump[i]), g['nojump_%d' % i])


def test_stored_jump_matrices_dense_selection_and_authority(small, md):
    g, _ = small
    mats = _ref_matrices(md)
    want = np.cumsum(g['jumps'], axis=0, dtype=np.int64)
    # in-memory trajectory: dense residency, every frame's counts from the matrices
    reader = md.MemoryReader(g['pos'], g['box'], g['times'])
    reader.nojump_matrixes = mats
    coords = md.Coordinates(reader)
    store = coords._resident().store_for(np.arange(len(coords)), nojump=True)
    images = store.images[:, :3 * 96].cpu().numpy().astype(np.int64).reshape(len(coords), 96, 3) - 128
    np.testing.assert_array_equal(images, want)
    # behind an atom selection (columns of the matrices are picked like m[..., selection], pbc.py:243)
    sel = g['sel']
    fr = FileLikeReader(md, g)
    fr.nojump_matrixes = mats
    t, r = md.correlation.shifted_correlation(md.correlation.msd, md.Coordinates(fr, atom_filter=sel).nojump,
                                              segments=6, window=0.4)
    close(r, g['msd_sel'], scale=float(np.nanmax(g['msd_sel'])), rtol=1e-11, atol=1e-13)
    # the matrices are authoritative: an extra jump in row 0 (e.g. correct_nojump_matrixes_for_whole) shifts
    # that component's count in every frame
    from scipy import sparse
    extra = sparse.lil_matrix(mats[0].shape, dtype=np.int8)
    extra[0, 3] = 1
    changed = md.MemoryReader(g['pos'], g['box'], g['times'])
    changed.nojump_matrixes = ((mats[0] + extra.tocsr()).astype(np.int8), mats[1], mats[2])
    store2 = md.Coordinates(changed)._resident().store_for(np.arange(len(coords)), nojump=True)
    images2 = store2.images[:, :3 * 96].cpu().numpy().astype(np.int64).reshape(len(coords), 96, 3) - 128
    shifted = want.copy()
    shifted[:, 3, 0] += 1
    np.testing.assert_array_equal(images2, shifted)
    # a matrix that does not fit the trajectory is refused
    bad = md.MemoryReader(g['pos'], g['box'], g['times'])
    bad.nojump_matrixes = tuple(m[:, :50] for m in mats)
    with pytest.raises(ValueError):
        md.Coordinates(bad)._resident().store_for(np.arange(4), nojump=True)


def test_images_from_jumps_unaligned_rows_and_overflow(small, md):
    """The scalar store path (rows that are not word-aligned), kept-frame subsets and the overflow flag."""
    import torch
    from mdevaluate_b200._cuda import engine
    g, _ = small
    mats = _ref_matrices(md)
    indptr, frames, jumps = (torch.from_numpy(a).cuda() for a in md.reader.jump_columns(mats))
    want = np.cumsum(g['jumps'], axis=0, dtype=np.int64).reshape(len(g['pos']), -1)
    keep = np.unique(np.random.default_rng(11).integers(0, len(g['pos']), 50)).astype(np.int32)
    for pitch in (288, 289, 291):
        buf = torch.full((len(keep) * pitch + 1,), 7, dtype=torch.uint8, device='cuda')
        images = buf[1:].view(len(keep), pitch) if pitch == 288 else buf[:len(keep) * pitch].view(len(keep), pitch)
        flag = torch.zeros(1, dtype=torch.int32, device='cuda')
        engine.images_from_jumps(indptr, frames, jumps, 96, images, keep=torch.from_numpy(keep).cuda(), overflow=flag)
        got = images.cpu().numpy()
        np.testing.assert_array_equal(got[:, :288].astype(np.int64) - 128, want[keep])
        assert (got[:, 288:] == 7).all() and int(flag.item()) == 0
    big = jumps.clone()
    big[:] = 100                                            # two crossings of 100 images: past one byte
    images = torch.empty((len(g['pos']), 288), dtype=torch.uint8, device='cuda')
    flag = torch.zeros(1, dtype=torch.int32, device='cuda')
    engine.images_from_jumps(indptr, frames, big, 96, images, overflow=flag)
    assert int(flag.item()) == 1


def test_generate_nojump_matrixes_writes_the_cache(small, md, tmp_path):
    """generate_nojump_matrixes on a named trajectory: GPU scan -> CSR matrices -> `.traj.nojump.npz` in the
    reference's layout; a new reader of the same trajectory picks the file up (reader.py:189-202, 283-285)."""
    g, _ = small
    name = str(tmp_path / 'walk.xtc')
    coords = md.Coordinates(md.MemoryReader(g['pos'], g['box'], g['times'], filename=name))
    md.reader.generate_nojump_matrixes(coords)
    path = tmp_path / '.walk.xtc.nojump.npz'
    assert path.exists()
    with np.load(path, allow_pickle=True) as a, np.load(_golden_file('ref.nojump.npz'), allow_pickle=True) as b:
        for key in b.files:
            if key != 'checksum':
                assert a[key].dtype == b[key].dtype and np.array_equal(a[key], b[key]), key
    again = md.MemoryReader(g['pos'], g['box'], g['times'], filename=name)
    dense = np.stack([m.toarray() for m in again.nojump_matrixes], axis=-1)
    np.testing.assert_array_equal(dense, g['jumps'])
    t, r = md.correlation.shifted_correlation(md.correlation.msd, md.Coordinates(again).nojump, segments=12, window=0.4)
    close(r, g['msd_nojump'], scale=float(np.nanmax(g['msd_nojump'])), rtol=1e-11, atol=1e-13)


def test_autosave_round_trip(small, md, tmp_path):
    """autosave.enable(): the first call computes on the GPU and writes shifted_correlation_<fn>_<desc>.npz,
    the second returns the stored arrays without touching the GPU (doc/special/autosave.rst)."""
    from mdevaluate_b200._cuda import engine
    g, _ = small
    coords = md.Coordinates(md.MemoryReader(g['pos'], g['box'], g['times'], filename=str(tmp_path / 'walk.xtc')))
    fn = partial(md.correlation.isf, q=22.7)
    md.autosave.enable('data', verbose=False)
    try:
        t1, r1 = md.correlation.shifted_correlation(fn, coords.nojump, segments=12, window=0.4, average=True,
                                                    description='OW')
        assert (tmp_path / 'data' / 'shifted_correlation_isf_OW.npz').exists()
        before = engine.launches()
        t2, r2 = md.correlation.shifted_correlation(fn, coords.nojump, segments=12, window=0.4, average=True,
                                                    description='OW')
        assert engine.launches() == before
        np.testing.assert_array_equal(t1, t2)
        np.testing.assert_array_equal(r1, r2)
        t3, r3 = md.correlation.shifted_correlation(fn, coords.nojump, segments=12, window=0.4, description='OW')
        assert engine.launches() > before and r3.shape == g['isf_nojump'].shape     # other arguments: recomputed
        close(r3, g['isf_nojump'])
        t4, r4 = md.correlation.shifted_correlation(fn, coords.nojump, segments=12, window=0.4, description='OW')
        np.testing.assert_array_equal(r3, r4)                                        # ragged (times, result) file
        np.testing.assert_array_equal(t3, t4)
    finally:
        md.autosave.disable()


@pytest.fixture
def tuning():
    """engine.tuning with every switch back at its default after the test."""
    from mdevaluate_b200._cuda import engine
    saved = {f: getattr(engine.tuning, f) for f in engine.Tuning.FIELDS + ('use_packed',)}
    yield engine.tuning
    for f, v in saved.items():
        setattr(engine.tuning, f, v)


def test_isf_float32_displacement_path_against_float64_path(md, tuning):
    """ISF alone with a constant float32 box computes displacements in float32 (operand on the far side of
    the boundary moved first, csrc/corr_kernels.cu::process_tile_f32d); tuning.isf_float64 forces the float64
    chain the other functions use.  Both against the oracle, and against each other, in a small box where
    most atoms cross several images (|n0 - nt| up to 3 and more: byte image counts) and in a large one (packed
    counts)."""
    from mdevaluate_b200.synthetic import random_walk
    c = md.correlation
    for n_atoms, box, sigma in ((1500, 1.3, 0.12), (1100, 14.4, 0.05)):
        pos, bx, times = random_walk(n_atoms, 160, seed=n_atoms, sigma=sigma, box=box)
        unw = orc.UnwrappedTrajectory(pos, bx, times)
        if box < 2:
            assert np.abs(unw.counts[:-60] - unw.counts[60:]).max() >= 3
        _, ref = orc.shifted_correlation(partial(orc.isf, q=22.7), unw, segments=9, window=0.5)
        coords = md.Coordinates.from_arrays(pos, bx, times)
        _, fast = c.shifted_correlation(partial(c.isf, q=22.7), coords.nojump, segments=9, window=0.5)
        store = coords._resident().store_for(np.arange(160), nojump=True)
        assert (store.packed is not None) == (np.abs(unw.counts).max() <= 1)
        tuning.isf_float64 = 1
        _, slow = c.shifted_correlation(partial(c.isf, q=22.7), coords.nojump, segments=9, window=0.5)
        tuning.isf_float64 = 0
        assert not np.array_equal(fast, slow)             # really two code paths
        close(fast, ref)
        close(slow, ref)
        assert np.abs(fast - slow).max() < 5e-8
        # sub-ensembles run the masked instantiation of the same path
        sel = c.subensemble_correlation(lambda f: np.where(np.asarray(f)[:, 0] % box < 0.5 * box)[0])
        _, fast_m = c.shifted_correlation(partial(c.isf, q=22.7), coords.nojump, segments=9, window=0.5, correlation=sel)
        _, ref_m = orc.shifted_correlation(partial(orc.isf, q=22.7), unw, segments=9, window=0.5,
                                           correlation=orc.subensemble_pairwise(
                                               lambda f: np.where(np.asarray(f)[:, 0] % box < 0.5 * box)[0]))
        close(fast_m, ref_m)


def test_packed_image_counts_equal_the_byte_rows(md, tuning):
    """Counts within [-1, 1] are re-coded as one word per 4 atoms (MDB_IMAGE_PACKED): the packed rows decode to the
    scan's byte rows, and every kernel gives bit-identical results from either format (the scale 4^d of a packed
    field is folded into the box: products and FMAs are exact)."""
    from mdevaluate_b200.synthetic import random_walk
    c = md.correlation
    pos, bx, times = random_walk(2311, 200, seed=11, sigma=0.05, box=6.0)
    coords = md.Coordinates.from_arrays(pos, bx, times)
    store = coords._resident().store_for(np.arange(200), nojump=True)
    assert store.packed is not None and store.images is not None
    unw = orc.UnwrappedTrajectory(pos, bx, times)
    assert np.abs(unw.counts).max() == 1
    np.testing.assert_array_equal(store.image_counts(), unw.counts)
    only_packed = type(store)(store.coords, store.n_atoms, packed=store.packed, box=store.box,
                              box_const=store.box_const)
    np.testing.assert_array_equal(only_packed.image_counts(), unw.counts)
    bins = np.linspace(0, 2.0, 201)
    fns = {'msd': c.msd, 'ng': c.non_gaussian, 'isf': partial(c.isf, q=22.7), 'vh': partial(c.van_hove_self, bins=bins)}
    sel = c.subensemble_correlation(lambda f: np.where(np.asarray(f)[:, 1] % 6.0 < 2.5)[0])
    for isf64, hist64 in ((0, 0), (1, 1)):
        tuning.isf_float64, tuning.hist_float64 = isf64, hist64
        got = {}
        for use in (True, False):
            tuning.use_packed = use
            for name, fn in fns.items():
                got[name, use] = c.shifted_correlation(fn, coords.nojump, segments=11, window=0.4)[1]
                got[name + '_sel', use] = c.shifted_correlation(fn, coords.nojump, segments=11, window=0.4,
                                                                correlation=sel)[1]
            got['fused', use] = c.shifted_correlation_multi({'vh': fns['vh'], 'ng': fns['ng']}, coords.nojump,
                                                            segments=11, window=0.4)[1]
        for name in list(fns) + [n + '_sel' for n in fns]:
            np.testing.assert_array_equal(got[name, True], got[name, False], err_msg=name)
        np.testing.assert_array_equal(got['fused', True]['vh'], got['fused', False]['vh'])
        np.testing.assert_array_equal(got['fused', True]['ng'], got['fused', False]['ng'])
        # the unwrapped float64 frame from a packed row
    tuning.use_packed = True
    frame = coords.nojump[150]
    np.testing.assert_array_equal(np.asarray(frame), unw[150])
    import torch
    from mdevaluate_b200._cuda import engine, lib
    row = engine.nojump_apply(store.coords[150], store.packed[150], store.box[150].cpu().numpy(), store.n_atoms,
                              lib.IMAGE_PACKED)
    np.testing.assert_array_equal(row.cpu().numpy().reshape(-1, 3), unw[150])


def test_image_counts_beyond_one_byte(md):
    """An atom that crosses more than 127 images: the reference has no limit (pbc.py:235-237 sums int8 jumps into
    int64); the scan falls back to int16 rows and the direct-load kernel."""
    rng = np.random.default_rng(3)
    n_atoms, n_frames, box = 40, 700, 1.0
    x = rng.uniform(0, box, size=(n_atoms, 3))
    drift = np.zeros((n_atoms, 3))
    drift[:7, 0] = 0.3                     # 210 box lengths in 700 frames
    drift[3:9, 2] = -0.27
    walk = x[None] + np.cumsum(rng.normal(0, 0.02, size=(n_frames, n_atoms, 3)) + drift[None], axis=0)
    pos = np.mod(walk, box).astype(np.float32)
    pos[pos >= np.float32(box)] -= np.float32(box)
    bx = np.array([box] * 3, dtype=np.float32)
    times = np.arange(n_frames, dtype=np.float64)
    unw = orc.UnwrappedTrajectory(pos, bx, times)
    assert np.abs(unw.counts).max() > 127
    c = md.correlation
    for compact in (False, True):
        coords = md.Coordinates.from_arrays(pos, bx, times)
        coords._resident().force_compact = compact
        _, r = c.shifted_correlation(c.msd, coords.nojump, segments=7, window=0.5)
        _, ref = orc.shifted_correlation(orc.msd, unw, segments=7, window=0.5)
        close(r, ref, scale=float(np.nanmax(ref)), rtol=1e-11, atol=1e-13)
        _, f = c.shifted_correlation(partial(c.isf, q=7.0), coords.nojump, segments=7, window=0.5)
        _, fref = orc.shifted_correlation(partial(orc.isf, q=7.0), unw, segments=7, window=0.5)
        close(f, fref)
    store = coords._resident().store_for(np.arange(n_frames), nojump=True)
    assert store.images is None and store.packed is None and store.images16 is not None
    np.testing.assert_array_equal(np.asarray(coords.nojump[650]), unw[650])
    # per-step jumps stay int8 whatever the running counts do
    jumps = md.reader.parse_jumps(md.Coordinates.from_arrays(pos, bx, times))
    dense = np.zeros((n_frames, n_atoms, 3), dtype=np.int64)
    for d, m in enumerate(jumps):
        dense[m.row, m.col, d] = m.data
    np.testing.assert_array_equal(np.cumsum(dense, axis=0), unw.counts)


def test_generate_nojump_matrixes_from_a_subset_covers_all_atoms(small, md):
    """The jump matrices belong to the reader and are indexed by every Coordinates' own selection afterwards
    (pbc.py:243): generating them through a subset still scans all atoms."""
    g, _ = small
    reader = md.MemoryReader(g['pos'], g['box'], g['times'])
    subset = md.Coordinates(reader, atom_filter=g['sel'])
    md.reader.generate_nojump_matrixes(subset)
    dense = np.stack([m.toarray() for m in reader.nojump_matrixes], axis=-1)
    np.testing.assert_array_equal(dense, g['jumps'])
    t, r = md.correlation.shifted_correlation(md.correlation.msd, subset.nojump, segments=6, window=0.4)
    close(r, g['msd_sel'], scale=float(np.nanmax(g['msd_sel'])), rtol=1e-11, atol=1e-13)


@pytest.mark.parametrize('n_atoms,n_frames,segments', [(96, 240, 120), (131, 200, 50), (300, 150, 75), (1000, 130, 65)])
def test_dense_staging_variants_agree(md, monkeypatch, tuning, n_atoms, n_frames, segments):
    """Warp-per-row staging of the dense kernel (the default, csrc/dense_kernels.cu::stage_rows_warp) and the CTA-wide
    staging loop (tuning.dense_stage_cta): same sums bit for bit (the staged rows are identical, the evaluation code
    is shared)."""
    from mdevaluate_b200.synthetic import random_walk
    pos, box, times = random_walk(n_atoms, n_frames, seed=n_atoms, sigma=0.08, box=1.7)
    c = md.correlation
    monkeypatch.setattr(c, 'DENSE_MIN_PAIRS', 1)
    dense = lambda a, b: np.arange(a, b)
    for fn in (c.msd, partial(c.isf, q=22.7)):
        for mode in ('nojump', 'raw'):
            coords = md.Coordinates.from_arrays(pos, box, times)
            frames = coords.nojump if mode == 'nojump' else coords
            tuning.dense_stage_cta = 1
            _, ref = c.shifted_correlation(fn, frames, segments=segments, window=0.5, average=True, index_distribution=dense)
            tuning.dense_stage_cta = 0
            _, got = c.shifted_correlation(fn, frames, segments=segments, window=0.5, average=True, index_distribution=dense)
            np.testing.assert_array_equal(got, ref)


def test_van_hove_float32_estimate_is_exact(small, md, tuning):
    """van Hove with a constant float32 box (the default path): float32 displacements feed the bin estimate, the
    reference's float64 arithmetic decides the samples near an edge -- the counts must equal those of the float64
    chain (tuning.hist_float64) and of the reference."""
    from mdevaluate_b200.synthetic import random_walk
    g, coords = small
    c = md.correlation
    t, r = c.shifted_correlation(partial(c.van_hove_self, bins=g['bins']), coords.nojump, segments=12, window=0.4)
    np.testing.assert_array_equal(r, g['vh_nojump'])
    sel = c.subensemble_correlation(lambda f: np.where(np.asarray(f)[:, 2] % 1.5 < 0.6)[0])
    _, masked = c.shifted_correlation(partial(c.van_hove_self, bins=g['bins']), coords.nojump, segments=9, window=0.4,
                                      correlation=sel)
    tuning.hist_float64 = 1
    t, r64 = c.shifted_correlation(partial(c.van_hove_self, bins=g['bins']), coords.nojump, segments=12, window=0.4)
    np.testing.assert_array_equal(r64, g['vh_nojump'])
    _, masked_ref = c.shifted_correlation(partial(c.van_hove_self, bins=g['bins']), coords.nojump, segments=9, window=0.4,
                                          correlation=sel)
    tuning.hist_float64 = 0
    np.testing.assert_array_equal(masked, masked_ref)
    # a box of config-2 size, 500 bins, enough samples to land within 1e-6 of bin edges; alone and next to alpha_2
    pos, box, times = random_walk(20000, 300, seed=3, sigma=0.05, box=14.4)
    big = md.Coordinates.from_arrays(pos, box, times)
    bins = np.linspace(0, 3.0, 501)
    vh = partial(c.van_hove_self, bins=bins)
    _, got = c.shifted_correlation(vh, big.nojump, segments=20)
    _, fused = c.shifted_correlation_multi({'vh': vh, 'a2': c.non_gaussian}, big.nojump, segments=20)
    _, a2 = c.shifted_correlation(c.non_gaussian, big.nojump, segments=20)
    tuning.hist_float64 = 1
    _, ref = c.shifted_correlation(vh, big.nojump, segments=20)
    _, fused64 = c.shifted_correlation_multi({'vh': vh, 'a2': c.non_gaussian}, big.nojump, segments=20)
    tuning.hist_float64 = 0
    np.testing.assert_array_equal(got, ref)
    np.testing.assert_array_equal(fused['vh'], ref)
    np.testing.assert_array_equal(fused64['vh'], ref)
    assert got.sum(axis=-1).min() > 0.99                     # (nearly) every atom lands in a bin
    # alpha_2 next to the histogram: float32 displacements, good to 1e-7 relative per atom (north star: 1e-5);
    # through the float64 chain it equals the stand-alone kernel to rounding
    close(fused['a2'][:, 1:], a2[:, 1:], atol=1e-7)
    close(fused64['a2'][:, 1:], a2[:, 1:], rtol=1e-11, atol=1e-13)


def test_peer_comm_of_one_rank(small, md):
    """A communicator of world size 1 drives the whole peer-memory path of the finish kernel on one GPU (stores into
    the inbox, flags, wait, ordered sum): same results as without it, histograms included."""
    from mdevaluate_b200._cuda import engine
    g, coords = small
    c = md.correlation
    comm = engine.Comm(0, 1, lambda h: [h], capacity_words=1 << 16)
    try:
        res = coords._resident()
        store = res.store_for(np.arange(len(coords)), nojump=True)
        starts, idx = c._schedule(coords, c.log_indices, 12, 0.4, None)
        rows = store.rows_for(starts[:, None] + idx[None, :])
        for fn in (c.msd, partial(c.isf, q=22.7), partial(c.van_hove_self, bins=g['bins'])):
            spec = c._kernel_spec(fn)
            plain = engine.CorrPlan(store, rows, spec.corr)
            peer = engine.CorrPlan(store, rows, spec.corr, comm=comm)
            for _ in range(3):                               # both parities of the inbox
                s0, h0 = plain.run()
                s1, h1 = peer.run()
                np.testing.assert_array_equal(s0, s1)
                if h0 is not None:
                    np.testing.assert_array_equal(h0, h1)
            plain.close()
            peer.close()
    finally:
        comm.close()


def test_repeated_calls_reuse_their_plan(small, md, tuning):
    """The second call of the same request on the same trajectory is a dictionary lookup and two kernel launches
    (correlation._cached_call): same numbers, fresh arrays, and no stale answers when a switch or the jump matrices
    change."""
    g, coords = small
    c = md.correlation
    fn = partial(c.isf, q=22.7)
    kw = dict(segments=12, window=0.4)
    t1, r1 = c.shifted_correlation(fn, coords.nojump, **kw)
    assert any(k[5] is c.isf for k in c._call_cache)
    t2, r2 = c.shifted_correlation(fn, coords.nojump, **kw)
    np.testing.assert_array_equal(r1, r2)
    np.testing.assert_array_equal(t1, t2)
    assert r2 is not r1 and t2 is not t1
    r2[:] = 0                                                    # results belong to the caller
    _, r3 = c.shifted_correlation(fn, coords.nojump, **kw)
    np.testing.assert_array_equal(r1, r3)
    _, avg = c.shifted_correlation(fn, coords.nojump, average=True, **kw)
    np.testing.assert_allclose(avg, r1.sum(axis=0) / len(r1), rtol=1e-14)
    close(r1, g['isf_nojump'])
    tuning.isf_float64 = 1                                       # another kernel path: not the cached plan
    _, r64 = c.shifted_correlation(fn, coords.nojump, **kw)
    tuning.isf_float64 = 0
    assert not np.array_equal(r64, r1) and np.abs(r64 - r1).max() < 5e-8
    _, raw = c.shifted_correlation(c.msd, coords, **kw)          # raw mode of the same trajectory: its own entry
    _, raw2 = c.shifted_correlation(c.msd, coords, **kw)
    np.testing.assert_array_equal(raw, raw2)
    assert raw.dtype == np.float32
    _, unw = c.shifted_correlation(c.msd, coords.nojump, **kw)
    assert unw.dtype == np.float64 and not np.allclose(unw, raw)
    # replacing the reader's jump matrices invalidates what was derived from the scan (they are authoritative)
    md.reader.generate_nojump_matrixes(coords)
    _, r4 = c.shifted_correlation(fn, coords.nojump, **kw)
    np.testing.assert_array_equal(r4, r1)
    mats = coords.frames.nojump_matrixes
    coords.frames.nojump_matrixes = tuple(m * 0 for m in mats)   # "no atom ever jumps"
    _, r5 = c.shifted_correlation(fn, coords.nojump, **kw)
    assert not np.array_equal(r5, r1)
    coords.frames.nojump_matrixes = None
    coords.frames._nojump_matrixes = None


def test_vectors_and_rotational_autocorrelation_on_the_device(golden, md):
    """coordinates.vectors mapped over the frames + rotational_autocorrelation (coordinates.py:524-563,
    correlation.py:223-236): the vector rows are built on the device (mdb_vectors) in the reference's arithmetic --
    float64 from the nojump frames, float32 with the minimum image from raw frames -- and feed the Legendre kernel."""
    from mdevaluate_b200._cuda import engine
    from mdevaluate_b200.coordinates import vectors
    g = golden('vectors')
    c = md.correlation
    coords = md.Coordinates(md.MemoryReader(g['pos'], g['box'], g['times']))
    launches = engine.launches()
    vec_nj = vectors(coords.nojump, atoms_a=g['oxygen'], atoms_b=g['hydrogens'], normed=True)
    for order in (1, 2):
        t, r = c.shifted_correlation(partial(c.rotational_autocorrelation, order=order), vec_nj, segments=9, window=0.4)
        np.testing.assert_array_equal(t, g['t'])
        close(r, g['rot%d_nojump' % order], rtol=1e-11, atol=1e-13)
    vec_raw = vectors(coords, atoms_a=g['oxygen'], atoms_b=g['oxygen'] + 1, normed=True, box=g['box'])
    _, r = c.shifted_correlation(partial(c.rotational_autocorrelation, order=2), vec_raw, segments=9, window=0.4)
    close(r, g['rot2_raw'], atol=1e-6)
    assert engine.launches() > launches
    # the rows themselves
    res = coords._resident()
    store = res.store_for(np.arange(len(coords)), nojump=True)
    rows = store.rows_for(np.array([0, 17, 139]))
    got = engine.vector_rows(store, rows, g['oxygen'][None, :], g['hydrogens'], True)
    assert got.coords.dtype.is_floating_point and got.coords.element_size() == 8
    np.testing.assert_allclose(got.coords[:, :150].cpu().numpy().reshape(3, 50, 3), g['vec_nojump'], rtol=1e-15, atol=0)
    raw = engine.vector_rows(store.without_images(), rows, g['oxygen'][None, :], (g['oxygen'] + 1)[None, :], True,
                             pbc_box=g['box'])
    assert raw.coords.element_size() == 4
    np.testing.assert_allclose(raw.coords[:, :150].cpu().numpy().reshape(3, 50, 3), g['vec_raw'], rtol=2e-7, atol=0)
    # host frames of the map agree with the device rows (nojump frames come back through mdb_nojump_apply)
    np.testing.assert_allclose(np.asarray(vec_nj[17]), g['vec_nojump'][1], rtol=1e-15, atol=0)


def test_multi_subensembles_in_one_pass(md, monkeypatch):
    """Spatially resolved ISF (examples/plot_spatialisf.py): disjoint sub-ensembles are accumulated by ONE kernel pass
    per group of four (mdb_corr_args.n_segments) -- bit-identical to one masked pass per sub-ensemble, and equal to the
    oracle's multi_subensemble_correlation."""
    import time
    from mdevaluate_b200.synthetic import random_walk
    from mdevaluate_b200._cuda import engine
    c = md.correlation
    pos, box, times = random_walk(6000, 220, seed=12, sigma=0.05, box=5.0)
    coords = md.Coordinates.from_arrays(pos, box, times)
    L = float(box[0])
    edges = np.linspace(0, L, 7)                       # six slabs along z: two passes of the one-pass kernel

    def slabs(frame):
        z = np.asarray(frame)[:, 2] % L
        return np.array([(z >= lo) & (z < hi) for lo, hi in zip(edges[:-1], edges[1:])])
    strategy = c.multi_subensemble_correlation(slabs)
    fn = partial(c.isf, q=22.7)
    kw = dict(segments=9, window=0.4, correlation=strategy)
    launches = engine.launches()
    t, one, n_one = c.shifted_correlation(fn, coords.nojump, **kw)
    used_one = engine.launches() - launches
    monkeypatch.setattr(c, 'MULTI_SEGMENTS_PER_PASS', 1)
    launches = engine.launches()
    t2, many, n_many = c.shifted_correlation(fn, coords.nojump, **kw)
    used_many = engine.launches() - launches
    np.testing.assert_array_equal(one, many)
    np.testing.assert_array_equal(n_one, n_many)
    assert used_one < used_many                         # 2 correlation passes instead of 6
    unw = orc.UnwrappedTrajectory(pos, box, times)
    _, ref, n_ref = orc.shifted_correlation(partial(orc.isf, q=22.7), unw, segments=9, window=0.4,
                                            correlation=orc.multi_subensemble_pairwise(slabs))
    np.testing.assert_array_equal(n_one, n_ref)
    close(one, np.asarray(ref))
    # overlapping selections cannot share a pass: still correct
    def overlapping(frame):
        z = np.asarray(frame)[:, 2] % L
        return np.array([z < 0.6 * L, z > 0.4 * L])
    monkeypatch.setattr(c, 'MULTI_SEGMENTS_PER_PASS', 4)
    _, ov, n_ov = c.shifted_correlation(fn, coords.nojump, segments=9, window=0.4,
                                        correlation=c.multi_subensemble_correlation(overlapping))
    _, ov_ref, n_ov_ref = orc.shifted_correlation(partial(orc.isf, q=22.7), unw, segments=9, window=0.4,
                                                  correlation=orc.multi_subensemble_pairwise(overlapping))
    np.testing.assert_array_equal(n_ov, n_ov_ref)
    close(ov, np.asarray(ov_ref))


def _half_selector(frame):
    return np.where(np.asarray(frame)[:, 0] < 0.75)[0]


def test_overlap_equals_the_reference(golden, md):
    """correlation.overlap through shifted_correlation (correlation.py:288-312 on coordinates.CoordinatesKDTree,
    coordinates.py:415-455): integer counts / tree size, so the values are IDENTICAL to the reference + scipy's
    cKDTree (tests/golden/overlap.npz) -- open tree, periodic tree, tree of a per-origin selection, nojump frames."""
    from mdevaluate_b200._cuda import engine
    from mdevaluate_b200.coordinates import CoordinatesKDTree
    g = golden('overlap')
    c = md.correlation
    coords = md.Coordinates(md.MemoryReader(g['pos'], g['box'], g['times']))
    box_len = float(g['box'][0])
    kw = dict(segments=8, window=0.4)
    launches = engine.launches()
    for radius in (0.03, 0.06):
        t, r = c.shifted_correlation(partial(c.overlap, crds_tree=CoordinatesKDTree(coords), radius=radius), coords, **kw)
        np.testing.assert_array_equal(t, g['t'])
        np.testing.assert_array_equal(r, g['open_%g' % radius])
        tree = CoordinatesKDTree(coords, boxsize=box_len)
        _, r = c.shifted_correlation(partial(c.overlap, crds_tree=tree, radius=radius), coords, **kw)
        np.testing.assert_array_equal(r, g['periodic_%g' % radius])
    assert engine.launches() > launches
    tree = CoordinatesKDTree(coords, selector=_half_selector, boxsize=box_len)
    _, r = c.shifted_correlation(partial(c.overlap, crds_tree=tree, radius=0.05), coords, **kw)
    np.testing.assert_array_equal(r, g['selector_0.05'])
    _, r = c.shifted_correlation(partial(c.overlap, crds_tree=CoordinatesKDTree(coords.nojump), radius=0.05),
                                 coords.nojump, average=True, **kw)
    np.testing.assert_array_equal(r, g['nojump_avg_0.05'])
    # called like a plain function (one tree, one frame), as the reference allows
    onset, frame = coords[3], coords[90]
    for tree, key in ((CoordinatesKDTree(coords), 'dist_open'), (CoordinatesKDTree(coords, boxsize=box_len), 'dist_periodic')):
        for radius in (0.02, 0.05, float(np.sort(g[key])[200])):      # the last one: a distance that occurs exactly
            assert c.overlap(onset, frame, tree, radius) == (g[key] <= radius).sum() / len(onset)
    assert tree.cache_info.currsize == 1
    tree.clear_cache()
    assert tree.cache_info.currsize == 0


def test_overlap_edge_cases(md, monkeypatch):
    """Radii from nothing to everything, a radius beyond half the periodic box (one or two cells per dimension), an
    empty tree, atoms outside a periodic box (cKDTree raises ValueError), frames far outside an open tree's bounds --
    against the oracle's brute force in cKDTree's arithmetic."""
    import oracle.md_oracle as orc
    from mdevaluate_b200.coordinates import CoordinatesKDTree, DeviceTree
    c = md.correlation
    rng = np.random.default_rng(5)
    n, length = 3000, 2.0
    pts = (rng.random((n, 3)) * length).astype(np.float32)
    qry = (pts + rng.normal(0, 0.08, (n, 3))).astype(np.float32)
    qry[:50] += 7.0                                     # far outside an open tree; wrapped back by a periodic one
    from mdevaluate_b200._cuda import engine
    for box in (None, length, np.array([length, 0.0, length])):
        want = orc.nearest_distances(pts, qry, box)
        tree = DeviceTree(pts, box)
        for radius in (0.0, 1e-9, 0.01, 0.05, 0.31, 0.9, 1.5, 5.0, 100.0, -1.0):
            assert tree.count_within(qry, radius) == int((want <= radius).sum()), (box, radius)
        # the atoms queried in file order instead of the tree's cell order (what a tree of a selection gets)
        monkeypatch.setattr(engine, 'overlap_file_order', True)
        for radius in (0.05, 0.31):
            assert tree.count_within(qry, radius) == int((want <= radius).sum()), (box, radius)
        monkeypatch.setattr(engine, 'overlap_file_order', False)
    # float64 points and frames
    pts64 = rng.random((800, 3)) * length
    qry64 = pts64 + rng.normal(0, 0.05, (800, 3))
    want = orc.nearest_distances(pts64, qry64)
    assert DeviceTree(pts64).count_within(qry64, 0.07) == int((want <= 0.07).sum())
    # one tree atom, no tree atom
    assert DeviceTree(pts[:1]).count_within(qry, 0.3) == int((orc.nearest_distances(pts[:1], qry) <= 0.3).sum())
    assert DeviceTree(pts[:0]).count_within(qry, 0.3) == 0
    with pytest.raises(ValueError):
        DeviceTree(pts + 1.0, length).count_within(qry, 0.1)
    # all atoms in a plane (zero extent along z)
    flat = pts.copy()
    flat[:, 2] = 0.5
    want = orc.nearest_distances(flat, qry)
    assert DeviceTree(flat).count_within(qry, 0.2) == int((want <= 0.2).sum())
    # an empty selection in shifted_correlation: 0 / 0 like the reference
    pos, box3, times = md.synthetic.random_walk(200, 60, seed=3, sigma=0.02, box=1.2)
    coords = md.Coordinates.from_arrays(pos, box3, times)
    tree = CoordinatesKDTree(coords, selector=lambda f: np.zeros(len(f), dtype=bool))
    _, r = c.shifted_correlation(partial(c.overlap, crds_tree=tree, radius=0.1), coords, segments=3, window=0.3)
    assert np.isnan(r).all()


@pytest.mark.parametrize('n_atoms', [100_000])
def test_overlap_at_size(md, n_atoms):
    """Water-like density (33.4 atoms / nm^3), radius 0.11 nm: the schedule's counts against the oracle on a sample of
    (origin, lag) pairs, the trivial lag 0 (every atom overlaps itself) and the decay with the lag."""
    import oracle.md_oracle as orc
    from mdevaluate_b200.coordinates import CoordinatesKDTree
    from mdevaluate_b200.synthetic import random_walk, box_length
    c = md.correlation
    pos, box, times = random_walk(n_atoms, 120, seed=9, sigma=0.02, box=box_length(n_atoms))
    coords = md.Coordinates.from_arrays(pos, box, times)
    tree = CoordinatesKDTree(coords, boxsize=float(box[0]))
    t, r = c.shifted_correlation(partial(c.overlap, crds_tree=tree, radius=0.11), coords, segments=4, window=0.5)
    assert r.shape[0] == 4 and np.all(r[:, 0] == 1.0)
    mean = r.mean(axis=0)
    assert np.all(np.diff(mean)[:8] <= 0) and np.all(np.diff(mean)[2:8] < 0) and 0.0 < mean[-1] < mean[3]
    start, idx = orc.schedule(len(coords), 4, 0.5)[0], orc.log_indices(0, 60)
    sub = slice(0, 2000)
    for s, l in ((0, 1), (2, 7), (3, len(idx) - 1)):
        want = orc.nearest_distances(pos[start[s]], pos[start[s] + idx[l]][sub], float(box[0]))
        orc_count = int((want <= 0.11).sum())
        one = md.coordinates.DeviceTree(pos[start[s]], float(box[0])).count_within(pos[start[s] + idx[l]][sub], 0.11)
        assert one == orc_count, (s, l)
    # the whole frame pairs: sums of sub-blocks add up (atoms are independent queries)
    s, l = 1, 5
    whole = int(round(r[s, l] * n_atoms))
    parts = sum(md.coordinates.DeviceTree(pos[start[s]], float(box[0])).count_within(pos[start[s] + idx[l]][a:a + 25000], 0.11)
                for a in range(0, n_atoms, 25000))
    assert whole == parts
