"""CPU-side checks: schedule, function recognition, finalisation, the C ABI surface, sharding.

No kernel is launched here (there is no GPU in the build container); the CUDA parity tests are
in test_gpu_parity.py.
"""
import ctypes as C
import os
import re
from functools import partial

import numpy as np
import pytest

import mdevaluate_b200 as md
from mdevaluate_b200 import correlation as corr
from mdevaluate_b200._cuda import lib as L
from oracle import md_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Frames:
    def __init__(self, n, start=None):
        self.n = n
        self._slice = slice(start, None)

    def __len__(self):
        return self.n


def test_log_indices_golden(golden):
    g = golden('schedule')
    for key in g.files:
        _, first, last = key.split('_')
        np.testing.assert_array_equal(corr.log_indices(int(first), int(last)), g[key])


@pytest.mark.parametrize('n,segments,window,skip', [
    (1000, 100, 0.5, None), (10000, 10, 0.5, None), (20000, 20, 0.2, 0.1), (240, 12, 0.4, None),
    (50000, 25000, 0.5, None), (977, 13, 0.37, 0.21), (10, 10, 0.5, None),
])
def test_schedule_matches_oracle(n, segments, window, skip):
    starts, idx = corr._schedule(_Frames(n), corr.log_indices, segments, window, skip)
    ref_starts, num = orc.schedule(n, segments, window, 0 if skip is None else skip)
    np.testing.assert_array_equal(starts, ref_starts)
    np.testing.assert_array_equal(idx, orc.log_indices(0, num))


def test_schedule_sizes_of_baseline_configs():
    # T=1000 -> L=70, T=10000 -> L=82, T=20000 -> L=84, chi4 schedule -> L=81
    for n, lags in ((1000, 70), (10000, 82), (20000, 84)):
        starts, idx = corr._schedule(_Frames(n), corr.log_indices, 10, 0.5, None)
        assert len(idx) == lags and idx[0] == 0
    starts, idx = corr._schedule(_Frames(20000), corr.log_indices, 20, 0.2, 0.1)
    assert len(idx) == 81 and starts[0] == 2000 and starts[-1] == 15300 and len(starts) == 20


def test_schedule_skip_from_slice_and_assert():
    starts, _ = corr._schedule(_Frames(240, start=24), corr.log_indices, 7, 0.3, None)
    ref, _ = orc.schedule(240, 7, 0.3, 24 / 240)
    np.testing.assert_array_equal(starts, ref)
    with pytest.raises(AssertionError):
        corr._schedule(_Frames(100), corr.log_indices, 10, 0.6, 0.5)


def test_function_recognition():
    assert corr._kernel_spec(corr.msd).corr.flags == L.SUM_R2
    assert corr._kernel_spec(corr.non_gaussian).corr.flags == L.SUM_R2 | L.SUM_R4
    sp = corr._kernel_spec(partial(corr.isf, q=22.7))
    assert sp.corr.flags == L.SUM_SINC and sp.corr.q == 22.7
    sp = corr._kernel_spec(partial(partial(corr.isf, q=1.0), q=3.0, box=None))
    assert sp.corr.q == 3.0
    sp = corr._kernel_spec(partial(corr.van_hove_self, bins=np.linspace(0, 1, 501)))
    assert sp.corr.flags == L.HIST and sp.corr.n_bins == 500 and sp.corr.bin_mode == L.BINS_UNIFORM
    assert sp.corr.dx == np.diff(np.linspace(0, 1, 501)).mean()
    sp = corr._kernel_spec(partial(corr.van_hove_self, bins=np.linspace(0.3, 1.5)))
    assert sp.corr.bin_mode == L.BINS_EDGES and sp.corr.n_bins == 49
    # the oracle's (and the reference's) functions are recognised by name
    assert corr._kernel_spec(partial(orc.isf, q=2.0)).name == 'isf'
    with pytest.raises(NotImplementedError):
        corr._kernel_spec(np.mean)
    with pytest.raises(TypeError):
        corr._kernel_spec(corr.isf)
    with pytest.raises(TypeError):
        corr._kernel_spec(partial(corr.msd, q=1.0))
    with pytest.raises(NotImplementedError):
        corr.shifted_correlation(corr.msd, _Frames(10), correlation=lambda f, fr: None)


def test_finalize_matches_oracle_functions(golden):
    """finalize() on exact per-pair sums reproduces the oracle's function values."""
    g = golden('walk_small')
    a, b = g['nojump_17'], g['nojump_100']
    d2 = ((a - b) ** 2).sum(axis=1)
    sums = np.zeros((1, 1, 3))
    sums[0, 0] = [d2.sum(), (d2 ** 2).sum(), np.sinc(np.sqrt(d2) * 22.7 / np.pi).sum()]
    n = len(a)
    np.testing.assert_allclose(corr._Spec('msd').finalize(sums, None, n, False)[0, 0], orc.msd(a, b), rtol=1e-14)
    np.testing.assert_allclose(corr._Spec('isf', q=22.7).finalize(sums, None, n, False)[0, 0],
                               orc.isf(a, b, 22.7), rtol=1e-13)
    np.testing.assert_allclose(corr._Spec('non_gaussian').finalize(sums, None, n, False)[0, 0],
                               orc.non_gaussian(a, b), rtol=1e-13)
    bins = g['bins']
    counts = orc.histogram(np.sqrt(d2), bins)[0]
    vh = corr._Spec('van_hove_self', bins=bins).finalize(None, counts[None, None, :], n, False)[0, 0]
    np.testing.assert_array_equal(vh, orc.van_hove_self(a, b, bins))
    assert corr._Spec('msd').finalize(sums, None, n, True).dtype == np.float32
    zero = np.zeros((1, 1, 3))
    assert np.isnan(corr._Spec('non_gaussian').finalize(zero, None, n, False)[0, 0])


def test_average_is_sequential_sum_over_origins():
    vals = np.random.default_rng(0).random((7, 5))
    acc = 0
    for row in vals:
        acc = acc + row
    np.testing.assert_array_equal(corr._average(vals, 7, True), acc / 7)
    assert corr._average(vals, 7, False) is vals


def test_c_abi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, 'include', 'mdevaluate_b200.h')).read()
    declared = set(re.findall(r'\b(mdb_[a-z_0-9]+)\s*\(', header))
    assert declared == set(L.EXPORTS)
    lib = L.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.mdb_abi_version() == L.ABI_VERSION


def test_c_abi_struct_layout_matches_header():
    # field order/count of the ctypes mirrors vs the header structs
    header = open(os.path.join(ROOT, 'include', 'mdevaluate_b200.h')).read()
    for cname, ctype in (('mdb_corr_args', L.CorrArgs), ('mdb_nojump_args', L.NojumpArgs),
                         ('mdb_dense_args', L.DenseArgs), ('mdb_jumpmat_args', L.JumpMatArgs)):
        body = re.search(r'typedef struct %s \{(.*?)\} %s;' % (cname, cname), header, re.S).group(1)
        body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
        fields = [re.split(r'[\s\*]+', f.strip())[-1].split('[')[0] for f in body.split(';') if f.strip()]
        assert fields == [f[0] for f in ctype._fields_]


def test_argument_errors_without_gpu():
    lib = L.load()
    a = L.CorrArgs()
    assert lib.mdb_shifted_corr(C.byref(a), None) == 1            # MDB_ERR_ARGUMENT, no CUDA call made
    assert lib.mdb_last_error()
    nj = L.NojumpArgs()
    assert lib.mdb_nojump_scan(C.byref(nj), None) == 1
    assert lib.mdb_nojump_apply(None, None, None, 0, None, None) == 1
    assert lib.mdb_dense_corr(C.byref(L.DenseArgs()), None) == 1
    assert lib.mdb_nojump_images_from_jumps(C.byref(L.JumpMatArgs()), None) == 1
    assert lib.mdb_copy_rows_h2d(None, 0, None, 0, 0, 0, None) == 1
    buf = (C.c_char * 64)()
    assert lib.mdb_copy_rows_h2d(buf, 8, buf, 16, 32, 2, None) == 1     # width > pitch
    assert lib.mdb_copy_rows_h2d(buf, 8, buf, 8, 0, 0, None) == 0       # nothing to copy: no CUDA call


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    pos = np.zeros((4, 2, 3), dtype=np.float32)
    coords = md.Coordinates.from_arrays(pos, np.ones(3, dtype=np.float32))
    with pytest.raises(md.CudaPathError):
        corr.shifted_correlation(corr.msd, coords, segments=1)
    with pytest.raises(md.CudaPathError):
        corr.msd(pos[0], pos[1])


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'mdevaluate_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in text and 'from oracle' not in text, f


def test_memory_reader_and_coordinates_protocol():
    pos = np.arange(5 * 4 * 3, dtype=np.float32).reshape(5, 4, 3)
    box = np.array([2.0, 3.0, 4.0], dtype=np.float32)
    times = np.arange(5) * 0.5
    rd = md.MemoryReader(pos, box, times)
    assert len(rd) == 5 and rd[2].time == 1.0 and rd[-1].step == 4
    np.testing.assert_array_equal(rd[1].box, np.diag(box))
    with pytest.raises(md.NojumpError):
        rd.nojump_matrixes
    with pytest.raises(IndexError):
        rd[5]
    coords = md.Coordinates(rd, atom_filter=np.array([True, False, True, True]))
    f = coords[3]
    assert isinstance(f, md.CoordinateFrame) and f.shape == (3, 3) and f.step == 3 and f.time == 1.5
    assert len(coords[2:]) == 5 and coords[2:]._slice == slice(2, None)
    assert [fr.step for fr in coords[1::2]] == [1, 3]
    assert coords.nojump.mode == 'nojump' and coords.mode is None
    with pytest.raises(md.coordinates.UnknownCoordinatesMode):
        coords.get_mode('pbc')
    per_frame = md.MemoryReader(pos, np.tile(box, (5, 1)))
    assert not per_frame.box_constant and per_frame[0].box.dtype == np.float32


def test_atom_blocks_cover_everything():
    from mdevaluate_b200.parallel import atom_block
    for n in (1, 7, 100, 1000003):
        for w in (1, 2, 3, 8):
            blocks = [atom_block(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            for (a0, a1), (b0, b1) in zip(blocks, blocks[1:]):
                assert a1 == b0 and (a0 % 4 == 0 or a0 == n) and (b0 % 4 == 0 or b0 == n)


def test_selector_masks():
    n = 10
    for sel, expect in (
            (np.array([1, 3, 4]), [1, 3, 4]),
            ((np.array([0, 9]),), [0, 9]),                       # what np.where / mask2indices return
            (slice(2, 8, 3), [2, 5]),
            (np.arange(n) % 2 == 0, [0, 2, 4, 6, 8]),
            (np.array([], dtype=int), []),
            (np.array([-1]), [9])):
        mask, count = corr._selector_mask(sel, n)
        assert mask.dtype == np.uint8 and count == len(expect)
        np.testing.assert_array_equal(np.nonzero(mask)[0], expect)
    with pytest.raises(NotImplementedError):
        corr._selector_mask(np.array([1, 1, 2]), n)
    sels, shape = corr._selector_array(np.array([[0, 1], [2, 3], [4, 5]]))
    assert shape == (3,)
    ragged = np.empty(2, dtype=object)
    ragged[0], ragged[1] = [0, 1, 2], []
    sels, shape = corr._selector_array(ragged)
    assert shape == (2,)
    masks = np.array([[True, False], [False, True]])
    assert corr._selector_array(masks)[1] == (2,)
    np.testing.assert_array_equal(md.utils.mask2indices([True, False, True])[0], [0, 2])
    assert md.utils.mask2indices([[True, True, False], [True, False, True]]).shape == (2, 1, 2)


def test_frame_ids_follow_python_indexing():
    starts = np.array([0, 5])
    ids = corr._frame_ids(starts, np.array([0, 1, 4]), 10)
    np.testing.assert_array_equal(ids, [[0, 1, 4], [5, 6, 9]])
    ids = corr._frame_ids(starts, np.array([0, -1]), 10)          # negative frames count from the end
    np.testing.assert_array_equal(ids, [[0, 9], [5, 4]])
    with pytest.raises(IndexError):
        corr._frame_ids(starts, np.array([0, 5]), 10)
    with pytest.raises(IndexError):
        corr._frame_ids(starts, np.array([0, -11]), 10)


def test_strategy_objects_carry_their_selector():
    sel = lambda frame: [0]
    sub = corr.subensemble_correlation(sel)
    assert sub._mdb_strategy == ('sub', sel, corr.correlation)
    multi = corr.multi_subensemble_correlation(sel)
    assert multi.has_counter and multi._mdb_strategy == ('multi', sel)
    # a sub-ensemble strategy around a custom inner strategy is not on the GPU path
    custom = corr.subensemble_correlation(sel, correlation_function=lambda f, fr: iter(()))
    with pytest.raises(NotImplementedError):
        corr.shifted_correlation(corr.msd, _Frames(10), correlation=custom)


def test_rotational_spec():
    sp = corr._kernel_spec(partial(corr.rotational_autocorrelation, order=4))
    assert sp.corr.flags == L.SUM_LEGENDRE and sp.corr.order == 4
    with pytest.raises(ValueError):
        corr._kernel_spec(partial(corr.rotational_autocorrelation, order=2.5))


def test_dense_slab_plan():
    """Atom slabs of an out-of-core dense pass (correlation._dense_slabs): one slab when the trajectory is or
    fits in HBM, otherwise whole 1024-atom tiles sized so that two slabs fit the memory budget."""
    from mdevaluate_b200 import correlation as c

    class FakeResident:
        def __init__(self, n_atoms, n_frames, budget, device=False):
            self.n_atoms, self.n_frames, self._budget, self._device = n_atoms, n_frames, budget, device

        def on_device(self):
            return self._device

        def memory_budget(self):
            return self._budget

    # BASELINE config 5: 200 000 atoms x 50 000 frames, 15 B per atom and frame with image counts = 150 GB
    slabs = c._dense_slabs(FakeResident(200000, 50000, int(125e9)), nojump=True)
    assert slabs == [(0, 82944), (82944, 165888), (165888, 200000)]
    per_atom = 50000 * 15
    assert 2 * 82944 * per_atom <= 125e9 < 2 * (82944 + 1024) * per_atom
    assert c._dense_slabs(FakeResident(200000, 50000, int(160e9)), nojump=True) == [(0, 200000)]      # fits
    assert c._dense_slabs(FakeResident(200000, 50000, int(125e9)), nojump=False) == [(0, 200000)]     # 120 GB raw
    assert c._dense_slabs(FakeResident(200000, 50000, int(1e9), device=True), nojump=True) == [(0, 200000)]
    tiny = c._dense_slabs(FakeResident(5000, 50000, int(1e6)), nojump=True)                           # never below a tile
    assert tiny[0] == (0, 1024) and tiny[-1] == (4096, 5000)
    try:
        c.DENSE_SLAB_ATOMS = 2500                                                                     # forced: rounded to tiles
        assert c._dense_slabs(FakeResident(5000, 100, int(1e12), device=True), nojump=True) == [
            (0, 2048), (2048, 4096), (4096, 5000)]
    finally:
        c.DENSE_SLAB_ATOMS = None


def test_vectors_map_on_raw_frames_matches_the_reference(golden):
    """coordinates.vectors through CoordinatesMap on raw (float32) frames, on the host: the reference's numbers
    (tests/golden/make_golden_vectors.py), minimum image included."""
    import mdevaluate_b200 as md
    from mdevaluate_b200.coordinates import CoordinatesMap, vectors
    g = golden('vectors')
    coords = md.Coordinates(md.MemoryReader(g['pos'], g['box'], g['times']))
    vec = vectors(coords, atoms_a=g['oxygen'], atoms_b=g['oxygen'] + 1, normed=True, box=g['box'])
    assert isinstance(vec, CoordinatesMap) and len(vec) == len(coords)
    for k, i in enumerate((0, 17, 139)):
        frame = vec[i]
        assert frame.step == i and frame.dtype == np.float32
        np.testing.assert_array_equal(np.asarray(frame), g['vec_raw'][k])
    unnormed = vectors(coords, atoms_a=g['oxygen'], atoms_b=g['oxygen'] + 2, box=g['box'])
    for k, i in enumerate((5, 100)):
        np.testing.assert_array_equal(np.asarray(unnormed[i]), g['vec_raw_unnormed'][k])
    assert vec[10:]._slice == slice(10, None) and vec.description.startswith('vectors')


def test_overlap_request_and_tree_cache():
    """Host side of correlation.overlap: the partial is recognised, its arguments are checked like a Python call, the
    tree class keeps the reference's constructor and LRU cache (coordinates.py:415-461) -- no device needed."""
    from functools import partial
    import mdevaluate_b200 as md
    from mdevaluate_b200 import correlation as c
    from mdevaluate_b200.coordinates import CoordinatesKDTree, DeviceTree
    pos, box, times = md.synthetic.random_walk(30, 12, seed=1, sigma=0.02, box=1.0)
    coords = md.Coordinates.from_arrays(pos, box, times)
    tree = CoordinatesKDTree(coords, selector=lambda f: np.arange(5), boxsize=1.0, maxcache=4)
    assert c._overlap_request(partial(c.msd)) is None
    got = c._overlap_request(partial(partial(c.overlap, crds_tree=tree), radius=0.1))
    assert got[0] is tree and got[1] == 0.1
    with pytest.raises(TypeError):
        c._overlap_request(partial(c.overlap, crds_tree=tree))
    with pytest.raises(TypeError):
        c._overlap_request(partial(c.overlap, crds_tree=tree, radius=0.1, q=3))
    with pytest.raises(NotImplementedError):
        c._overlap_request(partial(c.overlap, crds_tree=object(), radius=0.1))
    node = tree[3]
    assert isinstance(node, DeviceTree) and node.n == 5 and tree[3] is node
    np.testing.assert_array_equal(node.data, np.asarray(coords[3])[:5])
    assert tree.cache_info.hits == 1 and tree.cache_info.currsize == 1
    tree.clear_cache()
    assert tree.cache_info.currsize == 0
    assert md.checksum.checksum(tree) == md.checksum.checksum(CoordinatesKDTree(coords, selector=tree.selector, boxsize=1.0))
    assert md.checksum.checksum(tree) != md.checksum.checksum(CoordinatesKDTree(coords, boxsize=1.0))
