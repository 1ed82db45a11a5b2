"""Generate golden vectors from the REAL reference (run in the build container only).

    python tests/golden/make_golden.py

Imports /root/reference/mdevaluate with stub modules for its absent I/O
dependencies (pygmx, dask -- SURVEY.md section 8c; none is touched by the hot
path), drives ``reader.parse_jumps`` / ``pbc.nojump`` /
``correlation.shifted_correlation`` with a fake in-memory reader, and stores
inputs and outputs as small ``.npz`` fixtures next to this file.  The fixtures
are committed; /root/reference does not exist on the GPU box.
"""
import os
import sys
import types
import warnings
from functools import partial

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def load_reference():
    os.environ.setdefault('HOME', '/root')
    warnings.simplefilter('ignore')

    def stub(name, **attrs):
        mod = types.ModuleType(name)
        mod.__dict__.update(attrs)
        sys.modules[name] = mod

    class _Err(Exception):
        pass

    stub('pygmx', TPXReader=object, read_xtcframe=lambda *a, **k: None, open=lambda *a, **k: None)
    stub('pygmx.gromacs', _atoms_from_grofile=None, load_indices=None)
    stub('pygmx.errors', InvalidMagicException=_Err, InvalidIndexException=_Err, FileTypeError=_Err)
    stub('dask', delayed=lambda *a, **k: (lambda f: f), __version__='0.15.0')
    stub('dask.array')
    sys.path.insert(0, '/root/reference')
    import mdevaluate
    return mdevaluate


class FakeFrame:
    def __init__(self, positions, box, time):
        self.positions = positions
        self.box = box
        self.time = time
        self.coordinates = positions


class FakeReader:
    """The reader protocol Coordinates consumes (reader.py:254-286)."""
    filename = '/nonexistent/synthetic.xtc'

    def __init__(self, positions, box_diag, times):
        self.pos = positions
        self.box_diag = box_diag
        self.times = times
        self._nojump_matrixes = None

    def __len__(self):
        return self.pos.shape[0]

    def __getitem__(self, i):
        box = self.box_diag if self.box_diag.ndim == 1 else self.box_diag[i]
        return FakeFrame(self.pos[i], np.diag(box).astype(self.pos.dtype), self.times[i])

    @property
    def nojump_matrixes(self):
        return self._nojump_matrixes

    @nojump_matrixes.setter
    def nojump_matrixes(self, mats):
        self._nojump_matrixes = mats


class FakeAtoms:
    def __init__(self, m):
        self.masses = np.ones(m)
        self.charges = np.zeros(m)


class FakeSubset:
    def __init__(self, m, selection=None):
        self.selection = np.ones(m, bool) if selection is None else selection
        n = int(self.selection.sum())
        self.residue_ids = np.arange(1, n + 1)
        self.atoms = FakeAtoms(m)
        self.description = 'synthetic'


def reference_coordinates(md, positions, box_diag, times, selection=None):
    from scipy import sparse
    reader = FakeReader(positions, box_diag, times)
    T, M = positions.shape[:2]
    coords_all = md.coordinates.Coordinates(reader, atom_subset=FakeSubset(M))
    jd = md.reader.parse_jumps(coords_all)
    reader.nojump_matrixes = tuple(
        sparse.csr_matrix((np.array(m.data), (np.array(m.row), np.array(m.col))), shape=(T, M)) for m in jd
    )
    coords = md.coordinates.Coordinates(reader, atom_subset=FakeSubset(M, selection))
    return coords, reader


def dense_jumps(reader):
    return np.stack([m.toarray() for m in reader.nojump_matrixes], axis=-1).astype(np.int8)


def main():
    md = load_reference()
    from mdevaluate_b200.synthetic import random_walk
    corr = md.correlation

    # ---- G0: schedule / index goldens ------------------------------------------------
    sched = {}
    for (first, last) in [(0, 1), (0, 7), (0, 100), (0, 500), (0, 5000), (0, 10000), (0, 25000), (3, 977)]:
        sched['log_%d_%d' % (first, last)] = corr.log_indices(first, last)
    np.savez_compressed(os.path.join(HERE, 'schedule.npz'), **sched)

    # ---- G1: small periodic walk with many jumps, positions stored -------------------
    pos, box, times = random_walk(96, 240, seed=7, sigma=0.06, box=1.5)
    times = times * 0.5 + 3.0           # non-trivial time axis
    coords, reader = reference_coordinates(md, pos, box, times)
    g1 = {'pos': pos, 'box': box, 'times': times, 'jumps': dense_jumps(reader)}
    for i in (0, 1, 17, 100, 239):
        g1['nojump_%d' % i] = np.asarray(coords.nojump[i])
    bins = np.linspace(0, 1.2, 61)
    bins_off = np.linspace(0.05, 1.5, 30)
    bins_irr = np.concatenate([[0.0], np.sort(np.random.default_rng(3).uniform(0.01, 1.4, 25)), [1.45]])
    fns = {
        'msd': corr.msd,
        'isf': partial(corr.isf, q=22.7),
        'ng': corr.non_gaussian,
        'vh': partial(corr.van_hove_self, bins=bins),
        'vh_off': partial(corr.van_hove_self, bins=bins_off),
        'vh_irr': partial(corr.van_hove_self, bins=bins_irr),
    }
    g1.update(bins=bins, bins_off=bins_off, bins_irr=bins_irr)
    with np.errstate(all='ignore'):
        for name, fn in fns.items():
            for mode, frames in (('nojump', coords.nojump), ('raw', coords)):
                t, r = corr.shifted_correlation(fn, frames, segments=12, window=0.4)
                g1['%s_%s' % (name, mode)] = r
                g1['t_%s' % mode] = t
        t, r = corr.shifted_correlation(fns['isf'], coords.nojump, segments=20, window=0.2, skip=0.1,
                                        average=False)
        g1['chi4_fqt'] = r
        g1['chi4'] = len(coords[0]) * r.var(axis=0)
        g1['chi4_t'] = t
        t, r = corr.shifted_correlation(fns['msd'], coords.nojump, segments=7, window=0.3, average=True)
        g1['msd_avg'] = r
        g1['msd_avg_t'] = t
        t, r = corr.shifted_correlation(fns['msd'], coords.nojump[24:], segments=7, window=0.3, average=True)
        g1['msd_slice24'] = r
        t, r = corr.shifted_correlation(fns['isf'], coords.nojump, segments=5, window=0.25,
                                        index_distribution=lambda a, b: np.arange(a, b))
        g1['isf_dense'] = r
    # selection of a subset of atoms (coordinates.py:261 mask gather, pbc.py:242)
    sel = np.zeros(96, bool)
    sel[5:60:2] = True
    coords_sel, _ = reference_coordinates(md, pos, box, times, selection=sel)
    with np.errstate(all='ignore'):
        t, r = corr.shifted_correlation(fns['msd'], coords_sel.nojump, segments=6, window=0.4)
    g1['sel'] = sel
    g1['msd_sel'] = r
    np.savez_compressed(os.path.join(HERE, 'walk_small.npz'), **g1)

    # ---- G1b: sub-ensembles picked by the origin frame (correlation.py:27-81, 153-189) ----------
    L = float(box[0])

    def slab(frame, lo=0.0, hi=0.4):            # atoms whose wrapped z lies in [lo, hi) * L at the origin
        z = np.asarray(frame)[:, 2] % L
        return np.where((z >= lo * L) & (z < hi * L))[0]

    def slabs_masks(frame):                     # (3, N) boolean masks
        z = np.asarray(frame)[:, 2] % L
        edges = np.array([0.0, 0.3, 0.65, 1.0]) * L
        return np.array([(z >= edges[i]) & (z < edges[i + 1]) for i in range(3)])

    def slabs_ragged(frame):                    # object array of index arrays, one of them empty
        z = np.asarray(frame)[:, 2] % L
        out = np.empty(3, dtype=object)
        out[0] = np.where(z < 0.5 * L)[0]
        out[1] = np.where(z > 2 * L)[0]
        out[2] = np.where(z >= 0.5 * L)[0]
        return out

    gs = {'pos': pos, 'box': box, 'times': times, 'bins': bins}
    with np.errstate(all='ignore'):
        for name in ('msd', 'isf', 'ng', 'vh'):
            fn = fns[name]
            t, r = corr.shifted_correlation(fn, coords.nojump, segments=9, window=0.4,
                                            correlation=corr.subensemble_correlation(slab))
            gs['sub_' + name] = r
            t, r = corr.shifted_correlation(fn, coords.nojump, segments=9, window=0.4, average=True,
                                            correlation=corr.subensemble_correlation(partial(slab, lo=0.5, hi=0.9)))
            gs['sub_avg_' + name] = r
            for tag, sel in (('masks', slabs_masks), ('ragged', slabs_ragged)):
                if name == 'vh' and tag == 'ragged':
                    continue
                t, r, cnt = corr.shifted_correlation(fn, coords.nojump, segments=9, window=0.4,
                                                     correlation=corr.multi_subensemble_correlation(sel))
                gs['multi_%s_%s' % (tag, name)] = r
                gs['multi_%s_count' % tag] = cnt
                t, r, cnt = corr.shifted_correlation(fn, coords.nojump, segments=9, window=0.4, average=True,
                                                     correlation=corr.multi_subensemble_correlation(sel))
                gs['multi_avg_%s_%s' % (tag, name)] = r
                gs['multi_avg_%s_count' % tag] = cnt
        gs['sub_t'] = t
        rr = md.coordinates.spatial_selector(coords[5], lambda f: np.asarray(f)[:, 0], 0.2, 0.9)
        gs['spatial_selector_5'] = rr
    np.savez_compressed(os.path.join(HERE, 'subensemble.npz'), **gs)

    # ---- G1c: rotational autocorrelation of unit vectors (correlation.py:223-236) ---------------
    rng = np.random.default_rng(21)
    v = rng.standard_normal((150, 77, 3))
    for t in range(1, 150):
        v[t] = v[t - 1] + 0.15 * rng.standard_normal((77, 3))
        v[t] /= np.linalg.norm(v[t], axis=1, keepdims=True)
    v[0] /= np.linalg.norm(v[0], axis=1, keepdims=True)
    gr = {}
    for dt, tag in ((np.float64, 'f64'), (np.float32, 'f32')):
        vv = v.astype(dt)
        vc, _ = reference_coordinates(md, vv, np.ones(3, dtype=dt) * 100, np.arange(150) * 2.0)
        gr['vec_' + tag] = vv
        for order in (1, 2, 4):
            t, r = corr.shifted_correlation(partial(corr.rotational_autocorrelation, order=order), vc, segments=8,
                                            window=0.5)
            gr['rot%d_%s' % (order, tag)] = r
        gr['rot_t'] = t
    np.savez_compressed(os.path.join(HERE, 'rotational.npz'), **gr)

    # ---- G2: BASELINE config 1 (1000 x 1000, MSD, S=100, L=70), outputs only ---------
    pos, box, times = random_walk(1000, 1000, seed=20240)
    coords, reader = reference_coordinates(md, pos, box, times)
    g2 = {'box': box, 'pos_sum': np.float64(pos.astype(np.float64).sum()),
          'pos_probe': pos[::97, ::53].copy()}
    jumps = dense_jumps(reader)
    g2['jump_nnz'] = np.int64(np.count_nonzero(jumps))
    g2['count_checksum'] = np.cumsum(jumps, axis=0, dtype=np.int64)[-1].astype(np.int64)
    with np.errstate(all='ignore'):
        for name, fn in (('msd', corr.msd), ('isf', partial(corr.isf, q=22.7)), ('ng', corr.non_gaussian)):
            t, r = corr.shifted_correlation(fn, coords.nojump, segments=100, window=0.5)
            g2[name] = r
            t, r32 = corr.shifted_correlation(fn, coords, segments=100, window=0.5)
            g2[name + '_raw'] = r32
        g2['t'] = t
        vb = np.linspace(0, 1.0, 501)
        t, r = corr.shifted_correlation(partial(corr.van_hove_self, bins=vb), coords.nojump, segments=10,
                                        window=0.5)
        g2['vh_bins'] = vb
        g2['vh'] = r
        t, r = corr.shifted_correlation(partial(corr.van_hove_self, bins=vb), coords, segments=10, window=0.5)
        g2['vh_raw'] = r
    np.savez_compressed(os.path.join(HERE, 'config1.npz'), **g2)

    # ---- G3: utils.histogram on fixed data (test/test_utils.py:36-46 pattern) --------
    rng = np.random.default_rng(11)
    data = rng.random(1000)
    edge_data = np.concatenate([data, np.arange(0, 50) * 0.02, [0.3, 0.6, 0.999999, 1.0]])
    b0 = np.linspace(0, 1)
    b1 = np.linspace(0.3, 1.5)
    b2 = np.linspace(0, 1.0, 501)
    g3 = {'data': edge_data, 'b0': b0, 'b1': b1, 'b2': b2}
    for k, b in (('b0', b0), ('b1', b1), ('b2', b2)):
        h, c = md.utils.histogram(edge_data, b)
        g3['h_' + k] = h
        g3['c_' + k] = c
    g3['runmean5'] = md.utils.runningmean(data[:40], 5)
    np.savez_compressed(os.path.join(HERE, 'histogram.npz'), **g3)
    print('golden vectors written to', HERE)


if __name__ == '__main__':
    main()
