"""On-disk caches against fixtures the REAL reference wrote (tests/golden/make_golden_cache.py):
checksums (checksum.py:41-96), the jump-matrix cache ``.{traj}.nojump.npz`` (reader.py:125-241) and the
autosave result files (autosave.py:85-183).  CPU only: formats, names and host logic."""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, 'golden')
sys.path.insert(0, GOLDEN)

from make_golden_cache import Named, STRIP_SOURCE, checksum_cases  # noqa: E402

from mdevaluate_b200 import autosave, checksum as cs, reader as rd  # noqa: E402


@pytest.fixture(scope='module')
def ref_sums():
    with open(os.path.join(GOLDEN, 'checksum.json')) as fh:
        return json.load(fh)


# ---- checksums ------------------------------------------------------------------------------------
@pytest.mark.parametrize('name', sorted(checksum_cases()))
def test_checksum_equals_reference(ref_sums, name):
    assert str(cs.checksum(*checksum_cases()[name])) == ref_sums[name]


def test_checksum_properties(ref_sums):
    assert cs.checksum('42') == cs.checksum(42)                      # test/test_checksum.py:10
    a = np.array([1, 2, 3])
    assert cs.checksum(a) == cs.checksum(a.tobytes())                # test/test_checksum.py:16-17
    salt = cs.SALT
    try:
        first = cs.checksum(999)
        cs.SALT = b'999'
        assert first != cs.checksum(999)                             # test/test_checksum.py:11-13
    finally:
        cs.SALT = salt
    assert cs.strip_comments(STRIP_SOURCE) == ref_sums['strip_comments']


def test_version_decorator(ref_sums):
    @cs.version(3)
    def versioned(x):
        return x
    assert str(cs.checksum(versioned)) == ref_sums['version_decorator']
    assert versioned(5) == 5

    @cs.version(1)
    def f1():
        pass
    first = cs.checksum(f1)

    @cs.version(1)
    def f1(x, y):  # noqa: F811
        return x + y
    assert first == cs.checksum(f1)                                  # test/test_checksum.py:22-35

    @cs.version(2)
    def f1(x, y):  # noqa: F811
        pass
    assert first != cs.checksum(f1)                                  # test/test_checksum.py:37-41


def test_function_checksum_follows_source_and_free_variables():
    scale = 2.0

    def f(x):
        # a comment does not count
        return scale * x

    def g(x):
        return scale * x
    first = cs.checksum(f)
    assert first != cs.checksum(g)            # the name is part of the source
    scale = 3.0
    assert first != cs.checksum(f)            # free variables are part of the checksum


def test_package_functions_have_fixed_checksums():
    from functools import partial
    from mdevaluate_b200 import correlation as c
    assert cs.checksum(c.msd) == cs.checksum(str(cs.checksum('msd', 1)))
    assert cs.checksum(partial(c.isf, q=22.7)) != cs.checksum(partial(c.isf, q=22.8))
    assert cs.checksum(c.log_indices) != cs.checksum(c.correlation)


# ---- the nojump cache -----------------------------------------------------------------------------
class PlainReader:
    """A file-like reader: frames only through ``reader[i]`` (no in-memory block), a name, a checksum."""

    def __init__(self, pos, box, times, filename, token=987654321):
        self._pos, self._box, self._times = pos, box, times
        self.filename = filename
        self.token = token
        self.reads = []
        self._nojump_matrixes = None

    def __len__(self):
        return len(self._pos)

    def __getitem__(self, i):
        self.reads.append(int(i))
        return rd.MemoryFrame(self._pos[i], np.diag(self._box).astype(self._box.dtype), self._times[i], int(i))

    def __checksum__(self):
        return self.token

    @property
    def nojump_matrixes(self):
        if self._nojump_matrixes is None:
            raise rd.NojumpError('Nojump Data not available: {}'.format(self.filename))
        return self._nojump_matrixes

    @nojump_matrixes.setter
    def nojump_matrixes(self, mats):
        self._nojump_matrixes = mats


def test_reference_nojump_file_is_read(golden):
    g = golden('walk_small')
    stored, mats = rd.read_nojump_file(os.path.join(GOLDEN, 'ref.nojump.npz'), trust_pickle=True)
    assert stored == cs.checksum(rd.NOJUMP_MAGIC, cs.checksum(Named(987654321)))
    dense = np.stack([m.toarray() for m in mats], axis=-1)
    assert dense.dtype == np.int8
    np.testing.assert_array_equal(dense, g['jumps'])


def test_nojump_file_written_like_the_reference(golden, tmp_path):
    g = golden('walk_small')
    reader = PlainReader(g['pos'], g['box'], g['times'], str(tmp_path / 'walk_small.xtc'))
    _, mats = rd.read_nojump_file(os.path.join(GOLDEN, 'ref.nojump.npz'), trust_pickle=True)
    rd.save_nojump_matrixes(reader, mats)
    ours = tmp_path / '.walk_small.xtc.nojump.npz'
    assert rd.nojump_filename(reader) == str(ours) and ours.exists()
    with np.load(ours, allow_pickle=True) as a, np.load(os.path.join(GOLDEN, 'ref.nojump.npz'),
                                                        allow_pickle=True) as b:
        # every entry of the reference's layout, plus the pickle-free twin of the checksum
        assert sorted(a.files) == sorted(b.files + ['checksum_str'])
        assert str(a['checksum_str'][()]) == str(a['checksum'][()])
        for key in b.files:
            assert a[key].dtype == b[key].dtype and a[key].shape == b[key].shape, key
            assert np.array_equal(a[key], b[key]), key


def test_nojump_file_load_rules(golden, tmp_path):
    g = golden('walk_small')
    name = str(tmp_path / 'walk_small.xtc')
    reader = PlainReader(g['pos'], g['box'], g['times'], name)
    with pytest.raises(rd.NojumpError):
        reader.nojump_matrixes
    _, mats = rd.read_nojump_file(os.path.join(GOLDEN, 'ref.nojump.npz'), trust_pickle=True)
    rd.save_nojump_matrixes(reader, mats)
    fresh = PlainReader(g['pos'], g['box'], g['times'], name)
    assert rd.load_nojump_matrixes(fresh) is True
    assert (fresh.nojump_matrixes[2] != mats[2]).nnz == 0
    other = PlainReader(g['pos'], g['box'], g['times'], name, token=1)      # another trajectory: ignored, kept
    assert rd.load_nojump_matrixes(other) is False and os.path.exists(rd.nojump_filename(other))
    assert rd.load_nojump_matrixes(other, verify=False) is True            # a cache the reference wrote
    with open(rd.nojump_filename(reader), 'wb') as fh:                       # a corrupt file is removed
        fh.write(b'not a zip file')
    assert rd.load_nojump_matrixes(PlainReader(g['pos'], g['box'], g['times'], name)) is False
    assert not os.path.exists(str(tmp_path / '.walk_small.xtc.nojump.npz'))


def test_memory_reader_picks_up_its_cache(golden, tmp_path):
    from mdevaluate_b200 import MemoryReader
    g = golden('walk_small')
    name = str(tmp_path / 'traj.xtc')
    reader = MemoryReader(g['pos'], g['box'], g['times'], filename=name)
    with pytest.raises(rd.NojumpError):
        reader.nojump_matrixes
    _, mats = rd.read_nojump_file(os.path.join(GOLDEN, 'ref.nojump.npz'), trust_pickle=True)
    rd.save_nojump_matrixes(reader, mats)
    again = MemoryReader(g['pos'], g['box'], g['times'], filename=name)       # reader.py:283-285
    assert (again.nojump_matrixes[0] != mats[0]).nnz == 0
    moved = g['pos'].copy()
    moved[-1, 0, 0] += 0.25
    changed = MemoryReader(moved, g['box'], g['times'], filename=name)        # other data, same name
    with pytest.raises(rd.NojumpError):
        changed.nojump_matrixes


@pytest.mark.parametrize('filtered', [False, True])
def test_jump_columns_reproduce_the_image_counts(golden, filtered):
    """The column lists the kernel walks, replayed in NumPy: running counts of any kept frames equal the
    prefix sums of the reference's jump matrices (pbc.py:235-237), also behind an atom selection."""
    g = golden('walk_small')
    _, mats = rd.read_nojump_file(os.path.join(GOLDEN, 'ref.nojump.npz'), trust_pickle=True)
    sel = np.nonzero(g['sel'])[0] if filtered else None
    indptr, frames, jumps = rd.jump_columns(mats, sel)
    jumps_ref = g['jumps'] if sel is None else g['jumps'][:, sel]
    T, N = jumps_ref.shape[:2]
    assert indptr.dtype == np.int64 and frames.dtype == np.int32 and jumps.dtype == np.int8
    assert indptr.shape == (3 * N + 1,) and indptr[-1] == np.count_nonzero(jumps_ref)
    keep = np.unique(np.random.default_rng(5).integers(0, T, 40))
    counts = np.zeros((len(keep), 3 * N), dtype=np.int64)
    for c in range(3 * N):
        f, j = frames[indptr[c]:indptr[c + 1]], jumps[indptr[c]:indptr[c + 1]]
        assert (np.diff(f) > 0).all()
        counts[:, c] = [j[f <= k].sum(dtype=np.int64) for k in keep]
    want = np.cumsum(jumps_ref, axis=0, dtype=np.int64)[keep].reshape(len(keep), 3 * N)
    np.testing.assert_array_equal(counts, want)


# ---- autosave ---------------------------------------------------------------------------------------
class Frames:
    description = 'OW'

    def __init__(self, reader):
        self.frames = reader

    def __checksum__(self):
        return 555


@pytest.fixture
def autosave_on():
    yield
    autosave.disable()


def test_autosave_file_like_the_reference(golden, tmp_path, autosave_on):
    g = golden('walk_small')
    with open(os.path.join(GOLDEN, 'autosave.json')) as fh:
        want = json.load(fh)
    reader = PlainReader(g['pos'], g['box'], g['times'], str(tmp_path / 'walk_small.xtc'))
    calls = []

    @autosave.autosave_data(nargs=2, kwargs_keys=('scale',), version='toy-1')
    def toy(function, frames, scale=1.0):
        calls.append(scale)
        return np.arange(4.0) * scale, np.arange(8.0).reshape(4, 2) * scale

    # off: plain call, no file
    toy(Named('fn'), Frames(reader), scale=2.0, description='run1')
    assert calls == [2.0] and not (tmp_path / 'autosave_dir').exists()

    autosave.enable('autosave_dir', verbose=False)
    first = toy(Named('fn'), Frames(reader), scale=2.0, description='run1')
    path = tmp_path / 'autosave_dir' / want['filename']
    assert os.listdir(tmp_path / 'autosave_dir') == [want['filename']]
    with np.load(path, allow_pickle=True) as a, np.load(os.path.join(GOLDEN, 'ref_autosave.npz'),
                                                        allow_pickle=True) as b:
        assert sorted(b.files) == ['checksum', 'data']
        assert sorted(a.files) == ['checksum', 'checksum_str', 'data', 'item_0', 'item_1', 'n_items']
        assert str(a['checksum'][()]) == str(b['checksum'][()]) == want['checksum'] == str(a['checksum_str'][()])
        assert a['data'].dtype == b['data'].dtype and a['data'].shape == b['data'].shape
        assert all(np.array_equal(x, y) for x, y in zip(a['data'], b['data']))
    # second call: loaded, not recomputed; same values
    second = toy(Named('fn'), Frames(reader), scale=2.0, description='run1')
    assert calls == [2.0, 2.0]
    assert all(np.array_equal(x, y) for x, y in zip(first, second))
    # other arguments: recomputed, file replaced; autoload=False: recomputed
    toy(Named('fn'), Frames(reader), scale=3.0, description='run1')
    toy(Named('fn'), Frames(reader), scale=3.0, description='run1', autoload=False)
    assert calls == [2.0, 2.0, 3.0, 3.0]
    with autosave.disabled():
        toy(Named('fn'), Frames(reader), scale=3.0, description='run1')
    assert calls[-1] == 3.0 and len(calls) == 5
    assert autosave.autosave_directory == 'autosave_dir'
    # a file the reference wrote for the same call holds pickled entries only: it is ignored (recomputed and
    # replaced) unless the user opts in to unpickling; then it is accepted as it is
    import shutil
    shutil.copy(os.path.join(GOLDEN, 'ref_autosave.npz'), path)
    toy(Named('fn'), Frames(reader), scale=2.0, description='run1')
    assert len(calls) == 6
    shutil.copy(os.path.join(GOLDEN, 'ref_autosave.npz'), path)
    autosave.trust_pickle = True
    try:
        loaded = toy(Named('fn'), Frames(reader), scale=2.0, description='run1')
    finally:
        autosave.trust_pickle = False
    assert len(calls) == 6 and all(np.array_equal(x, y) for x, y in zip(first, loaded))
    ragged = autosave.save_data(str(tmp_path / 'ragged.npz'), 7, (np.arange(3), np.arange(4), np.arange(5)))
    assert ragged is None
    back = autosave.load_data(str(tmp_path / 'ragged.npz'))
    assert back.dtype == object and [len(x) for x in back] == [3, 4, 5]


def test_autosave_name_of_shifted_correlation(golden, tmp_path, autosave_on):
    """shifted_correlation_<function>_<description>.npz below <trajectory dir>/<autosave dir>
    (autosave.py:89-104; doc/special/autosave.rst)."""
    from functools import partial
    from mdevaluate_b200 import Coordinates, MemoryReader, correlation as c
    g = golden('walk_small')
    coords = Coordinates(MemoryReader(g['pos'], g['box'], g['times'], filename=str(tmp_path / 'traj.xtc')))
    autosave.enable('data', verbose=False)
    name = autosave.get_filename(c.shifted_correlation, 0, 'OW', partial(c.isf, q=22.7), coords.nojump)
    assert name == str(tmp_path / 'data' / 'shifted_correlation_isf_OW.npz')
    # the relevant arguments decide the checksum: mode, slice, q, schedule keywords
    def key(fn, frames, **kw):
        return cs.checksum('shifted_correlation-1', fn, frames, *kw.values())
    base = key(partial(c.isf, q=22.7), coords.nojump, segments=10)
    assert base == key(partial(c.isf, q=22.7), coords.nojump, segments=10)
    assert base != key(partial(c.isf, q=22.7), coords, segments=10)
    assert base != key(partial(c.isf, q=22.7), coords.nojump[24:], segments=10)
    assert base != key(partial(c.isf, q=20.0), coords.nojump, segments=10)
    assert base != key(partial(c.isf, q=22.7), coords.nojump, segments=11)


def test_correct_nojump_matrixes_for_whole(golden, tmp_path):
    """Row 0 of the stored matrices := jumps of frame 0 against frame0.whole (reader.py:244-251); for the
    single-atom residues of this package `whole` is the identity, so row 0 becomes zero -- and the file is
    rewritten."""
    from mdevaluate_b200 import Coordinates, MemoryReader
    g = golden('walk_small')
    name = str(tmp_path / 'traj.xtc')
    reader = MemoryReader(g['pos'], g['box'], g['times'], filename=name)
    _, mats = rd.read_nojump_file(os.path.join(GOLDEN, 'ref.nojump.npz'), trust_pickle=True)
    from scipy import sparse
    seeded = sparse.lil_matrix(mats[1].shape, dtype=np.int8)
    seeded[0, 5] = -1
    reader.nojump_matrixes = (mats[0], (mats[1] + seeded.tocsr()).astype(np.int8), mats[2])
    assert reader.nojump_matrixes[1][0, 5] == -1
    rd.correct_nojump_matrixes_for_whole(Coordinates(reader))
    fixed = reader.nojump_matrixes
    assert all(sparse.isspmatrix_csr(m) and m.dtype == np.int8 for m in fixed)
    assert all(m[0].nnz == 0 for m in fixed)
    assert (fixed[1][1:] != mats[1][1:]).nnz == 0                      # the other rows are untouched
    _, stored = rd.read_nojump_file(rd.nojump_filename(reader))
    assert all((a != b).nnz == 0 for a, b in zip(stored, fixed))


class _Boom:
    """Unpickling this object would run code (what ADVICE.md warns about)."""
    fired = []

    def __reduce__(self):
        return (_Boom.fired.append, ('unpickled',))


def test_cache_files_are_never_unpickled_by_default(golden, tmp_path, autosave_on):
    """A cache file next to a trajectory is data from someone else's directory: neither the automatic pick-up of
    ``.{traj}.nojump.npz`` nor an autosave lookup may unpickle its entries."""
    from mdevaluate_b200 import MemoryReader
    g = golden('walk_small')
    name = str(tmp_path / 'traj.xtc')
    _, mats = rd.read_nojump_file_matrices(os.path.join(GOLDEN, 'ref.nojump.npz'))
    evil = {'checksum': np.array(_Boom(), dtype=object)}
    for d, mat in enumerate(mats):
        evil['shape'] = mat.shape
        for attr in rd.CSR_ATTRS:
            evil['{}_{}'.format(attr, d)] = getattr(mat, attr)
    cache = tmp_path / '.traj.xtc.nojump.npz'
    np.savez(cache, **evil)
    reader = MemoryReader(g['pos'], g['box'], g['times'], filename=name)          # auto pick-up: ignored
    with pytest.raises(rd.NojumpError):
        reader.nojump_matrixes
    assert rd.load_nojump_matrixes(reader) is False and cache.exists()
    assert rd.load_nojump_matrixes(reader, verify=False) is True                  # matrices only, checksum untouched
    assert _Boom.fired == []
    with pytest.raises(rd.PickledChecksum):
        rd.read_nojump_file(str(cache))
    result = tmp_path / 'result.npz'
    np.savez(result, checksum=np.array(_Boom(), dtype=object), data=np.arange(3.0))
    assert autosave.verify_file(str(result), 1234) is False
    with pytest.raises(ValueError):
        autosave.load_data(str(result))
    assert _Boom.fired == []
