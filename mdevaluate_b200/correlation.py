"""``shifted_correlation`` and its per-atom correlation functions on the B200.

Drop-in for the part of mdevaluate/correlation.py that SURVEY.md section 8 puts on the hot path:
``shifted_correlation`` (86-201), ``log_indices`` (16-18), the ``correlation`` strategy (21-24),
``msd`` (203-208), ``isf`` (211-220), ``van_hove_self`` (239-248), ``non_gaussian`` (355-362) and
the chi4 recipe of examples/plot_chi4.py:29-37.

The schedule (origins, lags, times) is computed on the host with the SAME NumPy expressions as
the reference, so it is reproduced bit for bit.  The S x L x N evaluation loop runs in ONE CUDA
launch (csrc/corr_kernels.cu) that returns per-(origin, lag) sums; ratios, 1/N, the average over
origins and variances are taken afterwards, as the reference averages function values per origin
(correlation.py:191-199), not moments.

There is no CPU fallback: functions other than the four kernels (or a custom ``correlation``
strategy) raise ``NotImplementedError``; a missing CUDA library or device raises
``CudaPathError``.
"""
from __future__ import annotations

import functools
import weakref
from itertools import chain

import numpy as np

from ._cuda import engine
from ._cuda import lib as _lib
from ._cuda.lib import CudaPathError  # noqa: F401  (re-exported)
from .autosave import autosave_data
from .checksum import checksum
from .utils import bin_rule, runningmean  # noqa: F401

__all__ = ['log_indices', 'correlation', 'subensemble_correlation', 'multi_subensemble_correlation',
           'shifted_correlation', 'msd', 'isf', 'van_hove_self', 'non_gaussian', 'rotational_autocorrelation',
           'chi4', 'shifted_correlation_multi', 'overlap']


# --------------------------------------------------------------------------------------------
# schedule
# --------------------------------------------------------------------------------------------
def log_indices(first, last, num=100):
    """Logarithmically spaced integer lags in ``[first, last]`` (correlation.py:16-18)."""
    ls = np.logspace(0, np.log10(last - first + 1), num=num)
    return np.unique(np.int_(ls) - 1 + first)


def correlation(function, frames):
    """The default pairing strategy: first frame of ``frames`` against all, itself first
    (correlation.py:21-24).  ``shifted_correlation`` recognises this object and runs the whole
    origin x lag loop on the GPU; calling it directly evaluates pair by pair (each pair is one
    small kernel launch)."""
    iterator = iter(frames)
    start_frame = next(iterator)
    return map(lambda f: function(start_frame, f), chain([start_frame], iterator))


def subensemble_correlation(selector_function, correlation_function=correlation):
    """Correlate only the atoms ``selector_function(origin_frame)`` picks (correlation.py:27-35).

    Passed as ``correlation=`` to :func:`shifted_correlation` the selection becomes a per-origin
    atom mask inside the CUDA kernel (``mdb_corr_args.select``); called directly it evaluates pair
    by pair like the reference."""
    def c(function, frames):
        iterator = iter(frames)
        start_frame = next(iterator)
        selector = selector_function(start_frame)
        subensemble = map(lambda f: f[selector], chain([start_frame], iterator))
        return correlation_function(function, subensemble)
    c._mdb_strategy = ('sub', selector_function, correlation_function)
    return c


def multi_subensemble_correlation(selector_function):
    """Several sub-ensembles per origin and their sizes (correlation.py:38-81).  ``selector_function``
    returns an array whose entries are selectors (index lists, slices, masks); with
    :func:`shifted_correlation` every sub-ensemble is one masked pass of the CUDA kernel and the
    result is ``(times, result, count)`` as in the reference (correlation.py:153-189)."""
    def cmulti(function, frames):
        iterator = iter(frames)
        start_frame = next(iterator)
        selectors, sel_shape = _selector_array(selector_function(start_frame))
        f_values = np.zeros(sel_shape + np.shape(function(start_frame, start_frame)),)
        count = np.zeros(sel_shape, dtype=int)
        first = [True]

        def cc(act_frame):
            for index in np.ndindex(sel_shape):
                sel = selectors[index]
                sf_sel = start_frame[sel]
                if first[0]:
                    count[index] = len(sf_sel)
                f_values[index] = function(sf_sel, act_frame[sel]) if count[index] != 0 else 0
            first[0] = False
            return np.asarray(f_values.copy())
        return map(cc, chain([start_frame], iterator)), count
    cmulti.has_counter = True
    cmulti._mdb_strategy = ('multi', selector_function)
    return cmulti


def _selector_array(selectors):
    """The reference's reading of a multi-selector result (correlation.py:63-66)."""
    selectors = np.asarray(selectors)
    sel_shape = selectors.shape
    if sel_shape[-1] == 0:
        selectors = np.asarray(selectors, int)
    if selectors.dtype != object:
        sel_shape = sel_shape[:-1]
    return selectors, sel_shape


def _selector_mask(selector, n_atoms):
    """Any NumPy row selector (index array, tuple from np.where, slice, boolean mask) as a 0/1 byte
    mask over the atoms; the picked atoms keep their multiplicity 1 (duplicates are refused)."""
    mask = np.zeros(n_atoms, dtype=np.uint8)
    probe = np.arange(n_atoms)[selector]
    if probe.ndim != 1:
        raise NotImplementedError('selectors must pick rows (atoms) of the frame')
    mask[probe] = 1
    if int(mask.sum()) != probe.size:
        raise NotImplementedError('selectors that pick an atom more than once are not supported on the B200 path')
    return mask, probe.size


# --------------------------------------------------------------------------------------------
# the four per-atom functions: markers for shifted_correlation, GPU-evaluated when called
# --------------------------------------------------------------------------------------------
def msd(start, frame):
    """Mean square displacement (correlation.py:203-208)."""
    return _pair_value(_Spec('msd'), start, frame)


def isf(start, frame, q, box=None):
    """Incoherent intermediate scattering function, isotropic average ``sin(qr)/(qr)``
    (correlation.py:211-220).  Bind ``q`` with ``functools.partial(isf, q=22.7)``."""
    return _pair_value(_Spec('isf', q=q), start, frame)


def van_hove_self(start, end, bins):
    """Self part of the van Hove function: histogram of ``|r_i(0) - r_i(t)|`` over ``bins``,
    divided by the number of atoms (correlation.py:239-248)."""
    return _pair_value(_Spec('van_hove_self', bins=bins), start, end)


def non_gaussian(onset, frame):
    """Non-Gaussian parameter ``3/5 <r^4>/<r^2>^2 - 1`` (correlation.py:355-362); NaN at lag 0."""
    return _pair_value(_Spec('non_gaussian'), onset, frame)


def rotational_autocorrelation(onset, frame, order=2):
    """Legendre polynomial ``P_order`` of the scalar products of the vectors in ``onset`` and ``frame``,
    averaged (correlation.py:223-236).  The frames hold vectors (e.g. unit bond vectors)."""
    return _pair_value(_Spec('rotational_autocorrelation', order=order), onset, frame)


def overlap(onset, frame, crds_tree, radius):
    """Overlap with a reference configuration (correlation.py:288-312): the fraction -- relative to the size of
    the tree of frame ``onset.step`` -- of the atoms of ``frame`` that have a tree atom within ``radius``.

    ``crds_tree`` is a :class:`mdevaluate_b200.coordinates.CoordinatesKDTree`.  Through ``shifted_correlation``
    (``partial(overlap, crds_tree=..., radius=...)``) the whole schedule is one call of ``mdb_overlap``."""
    tree = crds_tree[onset.step]
    with np.errstate(divide='ignore', invalid='ignore'):
        return np.int64(tree.count_within(frame, radius)) / tree.n


def _overlap_pair(tree, frame, radius):
    """One tree against one frame (both host arrays): ``mdb_overlap`` with a single (origin, lag) pair."""
    frame = np.asarray(frame)
    if frame.ndim != 2 or frame.shape[1] != 3:
        raise ValueError('frame must be an (N, 3) array')
    if tree.n == 0 or len(frame) == 0:
        return 0
    tree_store, _ = _store_from_host_frames([tree.data])
    store, _ = _store_from_host_frames([frame])
    counts, _, _, bad = engine.overlap_counts(tree_store, np.zeros(1, np.int32), store, np.zeros((1, 1), np.int32),
                                              radius, tree.boxsize)
    if bad:
        raise ValueError('Some input data are outside of the periodic box (cKDTree needs 0 <= x < boxsize).')
    return int(counts[0, 0])


def _overlap_request(function):
    """``(tree, radius)`` if ``function`` is ``overlap`` with its tree and radius bound, else None."""
    kwargs = {}
    fn = function
    while isinstance(fn, functools.partial):
        if fn.args:
            return None
        merged = dict(fn.keywords)
        merged.update(kwargs)
        kwargs = merged
        fn = fn.func
    module = getattr(fn, '__module__', '') or ''
    if getattr(fn, '__name__', None) != 'overlap' or not (fn is overlap or module.endswith('correlation')
                                                          or module.endswith('md_oracle')):
        return None
    extra = set(kwargs) - {'crds_tree', 'radius'}
    if extra:
        raise TypeError('overlap() got unexpected keyword arguments %s' % sorted(extra))
    if 'crds_tree' not in kwargs or 'radius' not in kwargs:
        raise TypeError("overlap() needs 'crds_tree' and 'radius' bound with functools.partial")
    from .coordinates import CoordinatesKDTree
    if not isinstance(kwargs['crds_tree'], CoordinatesKDTree):
        raise NotImplementedError('overlap on the B200 path needs a mdevaluate_b200.coordinates.CoordinatesKDTree; '
                                  'there is no CPU fallback')
    return kwargs['crds_tree'], float(kwargs['radius'])


MAX_OVERLAP_TREE_ATOMS = 1 << 27      # tree atoms (origins x atoms) one mdb_overlap call sorts; more: origin batches


def _overlap_schedule(frames, tree, radius, frame_ids):
    """The overlap of every (origin, lag) of a schedule: values (S, L) float64."""
    from .coordinates import Coordinates
    if getattr(frames, 'comm', None) is not None:
        raise NotImplementedError('overlap needs every atom of the reference configuration on one GPU: '
                                  'atom-sharded coordinates are not supported')
    store, rows, _, _, _ = _provider(frames, frame_ids)
    # the step of the origin frames (CoordinateFrame.step, coordinates.py:266): the frame number for Coordinates
    if isinstance(frames, Coordinates):
        steps = np.asarray(frame_ids[:, 0], dtype=np.int64)
    else:
        steps = np.array([int(frames[int(i)].step) for i in frame_ids[:, 0]], dtype=np.int64)
    tree_frames = tree.frames
    if isinstance(tree_frames, Coordinates):
        n_tree_frames = len(tree_frames)
        if ((steps < -n_tree_frames) | (steps >= n_tree_frames)).any():
            raise IndexError('frame out of range')
        steps = np.where(steps < 0, steps + n_tree_frames, steps)
    tree_store, tree_rows, _, _, tree_comm = _provider(tree_frames, steps[:, None])
    if tree_comm is not None:
        raise NotImplementedError('overlap needs every atom of the reference configuration on one GPU')
    select = None
    sizes_host = None
    if tree.has_selector:
        picked = [_selector_mask(tree.selector(tree_frames[int(i)]), tree_store.n_atoms) for i in steps]
        masks = np.stack([m for m, _ in picked])
        sizes_host = np.array([c for _, c in picked])
        torch = _lib.require_cuda()
        pitch = (masks.shape[1] + 3) // 4 * 4
        host = np.zeros((masks.shape[0], pitch), dtype=np.uint8)
        host[:, :masks.shape[1]] = masks
        select = torch.from_numpy(host).to(tree_store.device)
    S = len(steps)
    batch = max(1, MAX_OVERLAP_TREE_ATOMS // max(1, tree_store.n_atoms))
    counts = np.empty(rows.shape, dtype=np.int64)
    sizes = np.empty(S, dtype=np.int64)
    for s0 in range(0, S, batch):
        s1 = min(S, s0 + batch)
        cnt, n_tree, _, bad = engine.overlap_counts(tree_store, tree_rows[s0:s1, 0], store, rows[s0:s1], radius,
                                                    tree.boxsize, None if select is None else select[s0:s1])
        if bad:
            raise ValueError('Some input data are outside of the periodic box (cKDTree needs 0 <= x < boxsize).')
        counts[s0:s1] = cnt
        sizes[s0:s1] = n_tree
    if sizes_host is not None and not np.array_equal(sizes, sizes_host):      # pragma: no cover
        raise CudaPathError('tree sizes on the device differ from the selector\'s')
    with np.errstate(divide='ignore', invalid='ignore'):
        return counts / sizes[:, None]


_KNOWN = {'msd': msd, 'isf': isf, 'van_hove_self': van_hove_self, 'non_gaussian': non_gaussian,
          'rotational_autocorrelation': rotational_autocorrelation}


class _Spec:
    """A recognised correlation function with its bound parameters."""

    def __init__(self, name, q=None, bins=None, order=2):
        self.name = name
        self.q = q
        self.order = order
        self.bins = None if bins is None else np.asarray(bins)
        if name == 'isf':
            if q is None:
                raise TypeError("isf() missing required argument 'q' (bind it with functools.partial)")
            self.corr = engine.CorrSpec(_lib.SUM_SINC, q=float(q))
        elif name == 'msd':
            self.corr = engine.CorrSpec(_lib.SUM_R2)
        elif name == 'non_gaussian':
            self.corr = engine.CorrSpec(_lib.SUM_R2 | _lib.SUM_R4)
        elif name == 'rotational_autocorrelation':
            if int(order) != order or not 0 <= order <= 32:
                raise ValueError('order must be an integer in 0..32')
            self.corr = engine.CorrSpec(_lib.SUM_LEGENDRE, order=int(order))
        elif name == 'van_hove_self':
            if bins is None:
                raise TypeError("van_hove_self() missing required argument 'bins'")
            mode, dx, nb = bin_rule(self.bins)
            self.corr = engine.CorrSpec(_lib.HIST, n_bins=nb, bin_mode=mode, dx=dx,
                                        edges=self.bins.astype(np.float64))
        else:  # pragma: no cover
            raise NotImplementedError(name)

    def finalize(self, sums, hist, n_atoms, float32):
        """Per-(origin, lag) function values from the shard-summed kernel output.  ``n_atoms`` is the
        number of correlated atoms: a scalar, or one count per origin (sub-ensembles, shape (S,))."""
        if np.ndim(n_atoms):
            n_atoms = np.asarray(n_atoms)
            if self.name == 'van_hove_self':
                if (n_atoms == 0).any():
                    raise ZeroDivisionError('division by zero')      # the reference's ``1 / len(start)``
                return (1 / n_atoms)[:, None, None] * hist
            n_atoms = n_atoms[:, None]
        elif n_atoms and self.name == 'isf':
            out = sums[..., 2] / n_atoms
            return out.astype(np.float32) if float32 else out
        with np.errstate(divide='ignore', invalid='ignore'):
            if self.name in ('msd', 'rotational_autocorrelation'):
                out = sums[..., 0] / n_atoms
            elif self.name == 'isf':
                out = sums[..., 2] / n_atoms
            elif self.name == 'non_gaussian':
                m2 = sums[..., 0] / n_atoms
                m4 = sums[..., 1] / n_atoms
                out = 3 / 5 * m4 / m2 ** 2 - 1
            else:
                return 1 / n_atoms * hist          # always float64 (correlation.py:248)
        # the reference computes in the frames' dtype: float32 frames give float32 results (the Legendre
        # polynomial has float64 coefficients, so rotational_autocorrelation is always float64)
        if float32 and self.name != 'rotational_autocorrelation':
            return out.astype(np.float32)
        return out


def _kernel_spec(function):
    """Map a user callable (one of the four functions, possibly wrapped in ``partial``) to a _Spec."""
    kwargs = {}
    fn = function
    while isinstance(fn, functools.partial):
        if fn.args:
            raise NotImplementedError('positional partial arguments are not supported; bind q / bins by keyword')
        merged = dict(fn.keywords)
        merged.update(kwargs)
        kwargs = merged
        fn = fn.func
    name = getattr(fn, '__name__', None)
    module = getattr(fn, '__module__', '') or ''
    ours = _KNOWN.get(name) is fn
    # the reference's own functions (and restatements of them) carry the same names
    theirs = name in _KNOWN and (module.endswith('correlation') or module.endswith('md_oracle'))
    if not (ours or theirs):
        raise NotImplementedError(
            'shifted_correlation on the B200 path supports msd, isf, van_hove_self, non_gaussian and '
            'rotational_autocorrelation '
            '(optionally via functools.partial); got %r. There is no CPU fallback.' % (function,))
    kwargs.pop('box', None)      # isf's ``box`` argument is unused by the reference (correlation.py:211)
    if name == 'isf':
        extra = set(kwargs) - {'q'}
    elif name == 'van_hove_self':
        extra = set(kwargs) - {'bins'}
    elif name == 'rotational_autocorrelation':
        extra = set(kwargs) - {'order'}
    else:
        extra = set(kwargs)
    if extra:
        raise TypeError('%s() got unexpected keyword arguments %s' % (name, sorted(extra)))
    return _Spec(name, **kwargs)


# --------------------------------------------------------------------------------------------
# frame providers: where the frames come from
# --------------------------------------------------------------------------------------------
def _store_from_host_frames(frame_list):
    """Upload a list of host (N, 3) arrays as the rows of a compact store (generic frames objects)."""
    torch = _lib.require_cuda()
    first = np.asarray(frame_list[0])
    if first.ndim != 2 or first.shape[1] != 3:
        raise ValueError('frames must be (N, 3) arrays')
    dtype = np.float32 if first.dtype == np.float32 else np.float64
    n_atoms = first.shape[0]
    pitch = engine.row_pitch(n_atoms)
    host = torch.zeros((len(frame_list), pitch), dtype=torch.float32 if dtype == np.float32 else torch.float64)
    view = host[:, :3 * n_atoms].numpy().reshape(len(frame_list), n_atoms, 3)
    for r, f in enumerate(frame_list):
        f = np.asarray(f)
        if f.shape != first.shape:
            raise ValueError('all frames must have the same shape')
        view[r] = f
    dev = torch.device('cuda', torch.cuda.current_device())
    return engine.DeviceStore(host.to(dev), n_atoms), dtype == np.float32


def _provider(frames, frame_ids):
    """Return ``(store, rows (S,L) int32, n_atoms, float32_results, comm)`` for the needed frames."""
    from .coordinates import Coordinates
    ids = np.asarray(frame_ids)
    if isinstance(frames, Coordinates):
        res = frames._resident()
        nojump = frames.mode == 'nojump'
        # repeated calls: the store and the row table of this schedule, as long as the resident still serves that
        # store (and, for nojump, its image counts still come from the reader's current jump matrices)
        key = (nojump, ids.shape, ids.tobytes())
        hit = res.lookup.get(key)
        if hit is not None and res.serves(hit[0], nojump):
            store, rows = hit
        else:
            store = res.store_for(ids.ravel(), nojump=nojump)
            rows = store.rows_for(ids)
            if len(res.lookup) > 32:
                res.lookup.clear()
            res.lookup[key] = (store, rows)
        _provider.base = (weakref.ref(res), weakref.ref(store))
        if not nojump:
            store = store.without_images()
        float32 = (not nojump) and res.dtype == np.float32
        return store, rows, res.n_atoms, float32, frames.comm
    from .coordinates import CoordinatesMap
    if isinstance(frames, CoordinatesMap):
        mapped = _vector_map_provider(frames, ids)
        if mapped is not None:
            return mapped
    # generic frames protocol (coordinates.py:257-293): pull the needed frames through __getitem__
    _provider.base = None
    uniq, inverse = np.unique(ids.ravel(), return_inverse=True)
    store, float32 = _store_from_host_frames([frames[int(i)] for i in uniq])
    rows = inverse.reshape(ids.shape).astype(np.int32)
    return store, rows, store.n_atoms, float32, getattr(frames, 'comm', None)


def _index_table(selector, n_atoms):
    """A frame subscript (index list, 2-D index list, boolean mask, slice) as an int array ``(m, n_vec)``."""
    probe = np.arange(n_atoms)[selector]
    if probe.ndim == 1:
        return probe[None, :]
    if probe.ndim == 2:
        return probe
    raise NotImplementedError('atom selections of vectors() must be one- or two-dimensional')


def _vector_map_provider(frames, ids):
    """``coordinates.vectors`` mapped over a Coordinates object: the vector rows of the needed frames are built on
    the device from the resident frame rows (``engine.vector_rows``).  None for any other map."""
    from . import coordinates as crd
    fn = frames.function
    base = frames.coordinates
    if not (isinstance(fn, functools.partial) and fn.func is crd.vectors.frame_function and not fn.args
            and isinstance(base, crd.Coordinates)):
        return None
    kw = dict(fn.keywords)
    if set(kw) - {'atoms_a', 'atoms_b', 'normed', 'box'} or 'atoms_a' not in kw or 'atoms_b' not in kw:
        return None
    box = kw.get('box')
    if box is not None and np.ndim(box) != 1:
        return None
    res = base._resident()
    if res.dtype != np.float32:
        return None
    nojump = base.mode == 'nojump'
    uniq, inverse = np.unique(ids.ravel(), return_inverse=True)
    store = res.store_for(uniq, nojump=nojump)
    if not nojump:
        store = store.without_images()
    vec_store = engine.vector_rows(store, store.rows_for(uniq), _index_table(kw['atoms_a'], res.n_atoms),
                                   _index_table(kw['atoms_b'], res.n_atoms), bool(kw.get('normed', False)), pbc_box=box)
    _provider.base = None
    rows = inverse.reshape(ids.shape).astype(np.int32)
    return vec_store, rows, vec_store.n_atoms, not nojump, base.comm


def _times(frames, idx):
    reader = getattr(frames, 'frames', None)
    times = getattr(reader, 'times', None)
    if times is not None:
        return np.asarray(times)[idx] - np.asarray(times)[0]
    from .coordinates import Coordinates
    if isinstance(frames, Coordinates):
        # CoordinateFrame.time is the reader frame's time (coordinates.py:117-119): no need to gather /
        # unwrap the positions of every lag frame just to read it
        return np.array([reader[int(i)].time for i in idx]) - reader[0].time
    return np.array([frames[int(i)].time for i in idx]) - frames[0].time       # correlation.py:151


_schedule_cache = {}
MAX_PAIRS_PER_LAUNCH = 1 << 22      # (origin, lag) pairs per kernel launch; more are batched by origin


def _schedule(frames, index_distribution, segments, window, skip):
    """Origins and lags, the reference's expressions verbatim (correlation.py:131-144).
    The default ``log_indices`` is pure, so its schedules are memoised (arrays are read-only)."""
    if skip is None:
        try:
            skip = frames._slice.start / len(frames)
        except (TypeError, AttributeError):
            skip = 0
    assert window + skip < 1
    key = None
    if index_distribution is log_indices:
        try:
            key = (len(frames), segments, float(window), float(skip))
            hit = _schedule_cache.get(key)
            if hit is not None:
                return hit
        except TypeError:
            key = None

    start_frames = np.unique(np.linspace(
        len(frames) * skip, len(frames) * (1 - window),
        num=segments, endpoint=False, dtype=int
    ))
    num_frames = int(len(frames) * (window))
    idx = np.asarray(index_distribution(0, num_frames))
    if key is not None:
        start_frames.setflags(write=False)
        idx.setflags(write=False)
        if len(_schedule_cache) > 256:
            _schedule_cache.clear()
        _schedule_cache[key] = (start_frames, idx)
    return start_frames, idx


DENSE_MIN_PAIRS = 2048          # below this the streaming kernel is just as good
DENSE_SLAB_ATOMS = None         # atoms per slab of an out-of-core dense run; None = from the free HBM


def _dense_slabs(res, nojump):
    """Atom ranges of one dense pass.  A trajectory that is (or fits) in HBM is one slab.  A larger one --
    BASELINE config 5: 200 000 atoms x 50 000 frames = 120 GB of frames + 30 GB of image counts -- is cut
    into atom slabs (atoms are independent, the per-lag sums add up): every slab's rows stream from the
    host in time segments while the tiled kernel still works on the previous slab, so two slabs are
    resident at a time."""
    n = res.n_atoms
    slab = DENSE_SLAB_ATOMS
    if slab is None:
        if res.on_device():
            return [(0, n)]
        per_atom = res.n_frames * (12 + (3 if nojump else 0))
        budget = res.memory_budget()
        if per_atom * n <= budget:
            return [(0, n)]
        slab = budget // 2 // per_atom
    slab = max(1024, int(slab) // 1024 * 1024)          # whole 1024-atom tiles (bulk copies, alignment)
    return [(a0, min(n, a0 + slab)) for a0 in range(0, n, slab)]


def _dense_sums(res, nojump, spec, first_frame, last_frame, first_origin, step, S, L):
    """Device ``(L, 2)`` sums of one resident trajectory (or atom slab); None if its rows are unaligned."""
    store = res.store_for(np.arange(first_frame, last_frame + 1), nojump=nojump)
    if not nojump:
        store = store.without_images()
    if store.has_images and store.images is None:
        return None                                   # int16 image counts: the streaming kernel reads them
    if store.coords.stride(0) % 4 or store.coords.data_ptr() % 16 or (
            store.images is not None and (store.images.stride(0) % 4 or store.images.data_ptr() % 4)):
        return None                                   # unaligned rows: the streaming kernel has a scalar path
    return engine.dense_device(store, first_origin, step, S, 0, L, spec.corr.flags, q=spec.corr.q)


def _dense_average(frames, spec, start_frames, idx):
    """``average=True`` over a dense schedule -- origins in arithmetic progression, lags a contiguous
    range (``index_distribution=lambda a, b: np.arange(a, b)``) -- for msd / isf: the tiled kernel
    (csrc/dense_kernels.cu) reuses every frame row for a whole block of (origin, lag) pairs and sums
    over the origins on the device.  Returns ``None`` when the schedule or function does not qualify."""
    from .coordinates import Coordinates
    if spec.name not in ('msd', 'isf') or not isinstance(frames, Coordinates):
        return None
    S, L = len(start_frames), len(idx)
    if S * L < DENSE_MIN_PAIRS or L < 2 or np.any(np.diff(idx) != 1) or idx[0] < 0:
        return None
    step = 1
    if S > 1:
        d = np.diff(start_frames)
        step = int(d[0])
        if step < 1 or step > engine.DENSE_MAX_ORIGIN_STEP or np.any(d != step):
            return None
    if start_frames[0] < 0 or start_frames[-1] + idx[-1] >= len(frames):
        return None                                   # let the general path raise / wrap indices
    res = frames._resident()
    nojump = frames.mode == 'nojump'
    if res.dtype != np.float32:
        return None
    # the origin of segment s is the FIRST frame of its iterator, start_s + idx[0] (correlation.py:21-24)
    span = (int(start_frames[0]), int(start_frames[-1] + idx[-1]), int(start_frames[0] + idx[0]), step, S, L)
    slabs = _dense_slabs(res, nojump)
    if len(slabs) == 1:
        sums = _dense_sums(res, nojump, spec, *span)
        if sums is None:
            return None
    else:
        torch = _lib.require_cuda()
        sums = None
        in_flight = []                                # (event after the slab's kernel, the slab)
        for a0, a1 in slabs:
            if len(in_flight) == 2:
                # two slabs are resident: the older one's kernel must be through before its rows are
                # released (the next slab's upload, on another stream, may be carved from them)
                kernel_done, old = in_flight.pop(0)
                kernel_done.synchronize()
                del old
            slab = res.atom_slab(a0, a1)
            part = _dense_sums(slab, nojump, spec, *span)       # uploads under the previous slab's kernel
            if part is None:
                return None
            sums = part if sums is None else sums.add_(part)
            kernel_done = torch.cuda.Event()
            kernel_done.record(torch.cuda.current_stream(res.device))
            in_flight.append((kernel_done, slab))
        del in_flight
    n_atoms = res.n_atoms
    if frames.comm is not None:
        import torch.distributed as dist
        group, n_atoms = frames.comm
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    sums = sums.cpu().numpy()
    out = sums[:, 0 if spec.name == 'msd' else 1] / n_atoms / S
    return out.astype(np.float32) if (not nojump) else out


def _global_counts(frames, counts):
    """Sub-ensemble sizes: summed over the shards of an atom-sharded trajectory (every rank runs the
    selector on its own atoms)."""
    comm = getattr(frames, 'comm', None)
    if comm is None:
        return counts
    from . import parallel
    torch = _lib.require_cuda()
    return parallel.allreduce_counts(counts, comm, torch.device('cuda', torch.cuda.current_device()))


def _frame_ids(start_frames, idx, n_frames):
    """(S, L) trajectory frame of every (origin, lag); Python index rules as in ``frames[i]``."""
    frame_ids = start_frames[:, None] + idx[None, :]
    if frame_ids.size:
        if frame_ids.max() >= n_frames or frame_ids.min() < -n_frames:
            raise IndexError('shifted_correlation: frame %d outside the trajectory (%d frames)'
                             % (frame_ids.max() if frame_ids.max() >= n_frames else frame_ids.min(), n_frames))
        if frame_ids.min() < 0:
            frame_ids = np.where(frame_ids < 0, frame_ids + n_frames, frame_ids)
    return frame_ids


def _merged_spec(specs):
    """One kernel request for all functions of a call."""
    flags = 0
    hist_spec = None
    q = 0.0
    order = 0
    if len(specs) > 1 and any(sp.name == 'rotational_autocorrelation' for sp in specs):
        raise NotImplementedError('rotational_autocorrelation runs alone (its frames hold vectors)')
    for sp in specs:
        flags |= sp.corr.flags
        if sp.corr.flags & _lib.HIST:
            if hist_spec is not None:
                raise NotImplementedError('one histogram per launch')
            hist_spec = sp.corr
        if sp.corr.flags & _lib.SUM_SINC:
            q = sp.corr.q
        order = max(order, sp.corr.order)
    corr = engine.CorrSpec(flags, q=q, order=order)
    if hist_spec is not None:
        corr.n_bins, corr.bin_mode, corr.dx, corr.edges = (hist_spec.n_bins, hist_spec.bin_mode, hist_spec.dx,
                                                           hist_spec.edges)
    return corr


def _spec_key(corr):
    return (corr.flags, corr.q, corr.order, corr.n_bins, corr.bin_mode, corr.dx,
            None if corr.edges is None else corr.edges.tobytes())


MAX_PLANS_PER_STORE = 16
MULTI_SEGMENTS_PER_PASS = 4      # sub-ensembles one kernel pass accumulates (shared memory: 16 KiB each); 1 = one pass each


def _plan_for(store, rows, corr, select_dev, p2p):
    """The cached :class:`engine.CorrPlan` of this request on this store (plans own their device / host buffers,
    so a repeated call allocates nothing)."""
    if select_dev is not None:
        return engine.CorrPlan(store, rows, corr, select=select_dev, comm=p2p)       # masks differ per call
    key = ('plan', rows.shape, rows.tobytes(), _spec_key(corr), id(p2p), engine.tuning.key())
    plan = store._cache.get(key)
    if plan is None:
        plans = [k for k in store._cache if isinstance(k, tuple) and k and k[0] == 'plan']
        if len(plans) >= MAX_PLANS_PER_STORE:
            for k in plans:
                store._cache.pop(k).close()
        plan = store._cache[key] = engine.CorrPlan(store, rows, corr, comm=p2p)
    return plan


def _run(frames, frame_ids, specs, select=None, n_segments=1):
    """Kernel launch + shard reduction; returns host ``(sums, hist, n_atoms_global, float32)``.
    ``select`` is an optional host uint8 ``(S, N)`` per-origin atom mask -- or, with ``n_segments`` > 1, the number
    1..n_segments of the sub-ensemble every atom belongs to (0: none); ``sums`` is then ``(S, L, n_segments, 3)``.
    ``sums`` / ``hist`` are views of the plan's page-locked result block: read them before the next call."""
    corr = _merged_spec(specs)
    store, rows, n_atoms, float32, comm = _provider(frames, frame_ids)
    select_dev = None
    if select is not None:
        torch = _lib.require_cuda()
        padded = np.zeros((select.shape[0], (select.shape[1] + 3) // 4 * 4), dtype=np.uint8)
        padded[:, :select.shape[1]] = select
        select_dev = torch.from_numpy(padded).to(store.device)
    p2p = getattr(comm, 'p2p', None) if comm is not None else None
    if n_segments > 1:
        plan = engine.CorrPlan(store, rows, corr, select=select_dev, comm=p2p, n_segments=n_segments)
    else:
        plan = _plan_for(store, rows, corr, select_dev, p2p)
    sums, hist = plan.run()
    _run.last = ((weakref.ref(plan), _provider.base, float32)
                 if (select is None and (comm is None or p2p is not None) and _provider.base is not None) else None)
    if comm is not None:
        if p2p is None:      # no peer memory (gloo / IPC refused): the torch.distributed all-reduce
            from . import parallel
            torch = _lib.require_cuda()
            dev = store.device if parallel.backend_of(comm) == 'nccl' else 'cpu'
            s_t = torch.from_numpy(np.array(sums)).to(dev)
            h_t = torch.from_numpy(np.array(hist)).to(dev) if hist is not None else None
            s_t, h_t, _ = parallel.allreduce_partials(s_t, h_t, n_atoms, comm)
            sums = s_t.cpu().numpy()
            hist = h_t.cpu().numpy() if h_t is not None else None
        n_atoms = comm[1]
    if select_dev is not None:
        sums = np.array(sums)
        hist = np.array(hist) if hist is not None else None
        plan.close()
    return sums, hist, n_atoms, float32


def _average(values, n_origins, average):
    if not average:
        return values
    result = 0
    for row in values:                      # correlation.py:191-199: sequential sum, then / S
        result = result + row
    return result / n_origins


def _fixed_checksum(function, version_nr):
    """Key the autosave files of this package's own functions by name and version instead of by source
    text (what checksum.version does, without wrapping: the functions are recognised by identity)."""
    fixed = checksum(function.__name__, version_nr)
    function.__checksum__ = lambda: fixed


for _f in (log_indices, correlation, msd, isf, van_hove_self, non_gaussian, rotational_autocorrelation, overlap):
    _fixed_checksum(_f, 1)
del _f


# --------------------------------------------------------------------------------------------
# repeated calls: everything a call needs besides the kernel launch, remembered per (trajectory, function, schedule)
# --------------------------------------------------------------------------------------------
class _CachedCall:
    """Weak references only: the plan lives in its store's cache, the store in its resident, the resident on the
    reader -- remembering a call must not keep gigabytes of HBM alive."""
    __slots__ = ('res_ref', 'store_ref', 'plan_ref', 'spec', 'times', 'n_origins', 'float32', 'tuning', 'comm')


_call_cache = {}
MAX_CACHED_CALLS = 64


def _call_key(function, frames, index_distribution, segments, window, skip):
    """Hashable identity of a plain call with the default lag distribution on a Coordinates object, or None."""
    from .coordinates import Coordinates
    if index_distribution is not log_indices or not isinstance(frames, Coordinates):
        return None, None
    res = frames.__dict__.get('_resident_obj')
    if res is None:
        return None, None
    fn, kws = function, ()
    if isinstance(fn, functools.partial):
        if fn.args or isinstance(fn.func, functools.partial):
            return None, None
        kws = tuple(sorted(fn.keywords.items()))
        fn = fn.func
    sl = frames._slice
    key = (id(res), frames.mode, sl.start, sl.stop, sl.step, fn, kws, segments, window, skip, id(frames.comm))
    try:
        hash(key)
    except TypeError:          # e.g. bins given as an array: the general path has its own caches
        return None, None
    return key, res


def _cached_call(key, res, frames, average):
    entry = _call_cache.get(key)
    if entry is None:
        return None
    store, plan = entry.store_ref(), entry.plan_ref()
    if (entry.res_ref() is not res or entry.comm is not frames.comm or store is None or plan is None
            or plan.handle is None or entry.tuning != engine.tuning.key()
            or not res.serves(store, frames.mode == 'nojump')):
        del _call_cache[key]
        return None
    sums, hist = plan.run()
    n_atoms = res.n_atoms if entry.comm is None else entry.comm[1]
    values = entry.spec.finalize(sums, hist, n_atoms, entry.float32)
    return entry.times.copy(), _average(values, entry.n_origins, average)


def _remember_call(key, res, frames, spec, times, n_origins):
    last = getattr(_run, 'last', None)
    if last is None or last[1][0]() is not res:
        return
    entry = _CachedCall()
    entry.plan_ref, (entry.res_ref, entry.store_ref), entry.float32 = last
    entry.spec, entry.times, entry.n_origins = spec, np.array(times), n_origins
    entry.tuning, entry.comm = engine.tuning.key(), frames.comm
    if len(_call_cache) >= MAX_CACHED_CALLS:
        _call_cache.clear()
    _call_cache[key] = entry


def shifted_correlation(function, frames, index_distribution=log_indices, correlation=correlation,
                        segments=10, window=0.5, skip=None, average=False):
    """Calculate the time series of a correlation function over shifted time origins.

    Same call as the reference (correlation.py:83-201), including its autosave decorator: with
    ``autosave.enable(dir)`` the result is written to / read from
    ``shifted_correlation_<function>_<description>.npz``; ``description=`` and ``autoload=`` are the
    decorator's extra keyword arguments (autosave.py:157-158).

    Returns ``(times, result)`` with ``result`` of shape ``(segments, lags[, bins])`` or, with
    ``average=True``, ``(lags[, bins])``.
    """
    strategy = getattr(correlation, '_mdb_strategy', None)
    plain = correlation is globals()['correlation']
    if not plain and not (strategy and (strategy[0] == 'multi' or strategy[2] is globals()['correlation'])):
        raise NotImplementedError('only the default strategy, subensemble_correlation(selector) and '
                                  'multi_subensemble_correlation(selector) are on the B200 path; '
                                  'there is no CPU fallback')
    call_key = res_key = None
    if plain:
        call_key, res_key = _call_key(function, frames, index_distribution, segments, window, skip)
        if call_key is not None:
            hit = _cached_call(call_key, res_key, frames, average)
            if hit is not None:
                return hit
    request = _overlap_request(function)
    if request is not None:
        if not plain:
            raise NotImplementedError('overlap is evaluated with the default correlation strategy only')
        start_frames, idx = _schedule(frames, index_distribution, segments, window, skip)
        frame_ids = _frame_ids(start_frames, idx, len(frames))
        values = _overlap_schedule(frames, request[0], request[1], frame_ids)
        return _times(frames, idx), _average(values, len(start_frames), average)
    spec = _kernel_spec(function)
    start_frames, idx = _schedule(frames, index_distribution, segments, window, skip)
    frame_ids = _frame_ids(start_frames, idx, len(frames))
    times = _times(frames, idx)
    if plain and average:
        dense = _dense_average(frames, spec, start_frames, idx)
        if dense is not None:
            return times, dense
    if plain:
        n_origins = len(start_frames)
        batch = max(1, MAX_PAIRS_PER_LAUNCH // max(1, len(idx)))
        if batch >= n_origins:
            sums, hist, n_atoms, float32 = _run(frames, frame_ids, [spec])
            values = spec.finalize(sums, hist, n_atoms, float32)
            if call_key is not None:
                _remember_call(call_key, res_key, frames, spec, times, n_origins)
            return times, _average(values, n_origins, average)
        # very many (origin, lag) pairs (dense schedules): origins go through in batches so that the
        # per-pair device buffers stay bounded; the average is accumulated origin by origin as above
        acc = 0 if average else []
        for s0 in range(0, n_origins, batch):
            sums, hist, n_atoms, float32 = _run(frames, frame_ids[s0:s0 + batch], [spec])
            values = spec.finalize(sums, hist, n_atoms, float32)
            if average:
                for row in values:
                    acc = acc + row
            else:
                acc.append(values)
        return times, (acc / n_origins if average else np.concatenate(acc, axis=0))

    # ---- sub-ensembles: the selector sees every origin frame on the host, the kernel gets masks ----
    from .coordinates import Coordinates
    if isinstance(frames, Coordinates):
        # the schedule's store first: the origin frames below (and every batch) are then served from it instead of
        # one single-frame store -- for a streamed trajectory one pass over all frames -- per origin
        frames._resident().store_for(frame_ids.ravel(), nojump=frames.mode == 'nojump')
    origins = [frames[int(i)] for i in frame_ids[:, 0]]
    n_all = len(origins[0])
    if strategy[0] == 'sub':
        picked = [_selector_mask(strategy[1](o), n_all) for o in origins]
        masks = np.stack([m for m, _ in picked])
        counts = _global_counts(frames, np.array([c for _, c in picked]))
        sums, hist, _, float32 = _run(frames, frame_ids, [spec], select=masks)
        values = spec.finalize(sums, hist, counts, float32)
        return times, _average(values, len(start_frames), average)

    sel_shape = None
    per_origin = []
    for o in origins:
        selectors, shape = _selector_array(strategy[1](o))
        if sel_shape is None:
            sel_shape = shape
        elif shape != sel_shape:
            raise ValueError('the selector function must return the same number of sub-ensembles for every origin')
        per_origin.append([_selector_mask(selectors[index], n_all) for index in np.ndindex(sel_shape)])
    n_sub = len(per_origin[0])
    counts = np.array([[c for _, c in row] for row in per_origin]).reshape((len(origins),) + sel_shape)
    counts = _global_counts(frames, counts)
    values = None
    indices = list(np.ndindex(sel_shape))

    def put(index, sums, hist, float32):
        nonlocal values
        cnt = counts[(slice(None),) + index]
        safe = np.where(cnt == 0, 1, cnt)
        vals = spec.finalize(sums, hist, safe, float32)                  # (S, L[, B])
        vals = np.where((cnt == 0).reshape((-1,) + (1,) * (vals.ndim - 1)), 0, vals)   # ``... else 0``
        if values is None:
            values = np.zeros((len(origins),) + sel_shape + vals.shape[1:], dtype=np.float64)
        values[(slice(None),) + index] = vals

    # ONE pass for several sub-ensembles (ISF, disjoint selections -- e.g. radial shells): the selector table holds
    # the number of the sub-ensemble of every atom and the kernel keeps one accumulator per sub-ensemble
    # (mdb_corr_args.n_segments).  Anything else: one masked pass per sub-ensemble.
    todo = list(range(n_sub))
    all_masks = np.stack([np.stack([row[k][0] for row in per_origin]) for k in range(n_sub)])     # (n_sub, S, N)
    if spec.name == 'isf' and n_sub > 1 and MULTI_SEGMENTS_PER_PASS > 1 and int(all_masks.sum(axis=0).max()) <= 1:
        try:
            for g0 in range(0, n_sub, MULTI_SEGMENTS_PER_PASS):
                group = list(range(g0, min(n_sub, g0 + MULTI_SEGMENTS_PER_PASS)))
                seg = np.zeros(all_masks.shape[1:], dtype=np.uint8)
                for j, k in enumerate(group):
                    seg[all_masks[k] != 0] = j + 1
                sums, _, _, float32 = _run(frames, frame_ids, [spec], select=seg, n_segments=max(2, len(group)))
                for j, k in enumerate(group):
                    put(indices[k], sums[:, :, j, :], None, float32)
                    todo.remove(k)
        except CudaPathError as exc:
            if 'status 3' not in str(exc):          # MDB_ERR_UNSUPPORTED: this store / box has no one-pass kernel
                raise
    for k in todo:
        sums, hist, _, float32 = _run(frames, frame_ids, [spec], select=all_masks[k])
        put(indices[k], sums, hist, float32)
    # values: (S, *sel_shape, L, ...); the reference's per-origin array is (L, *sel_shape, ...)
    cdim = len(sel_shape)
    if not average:
        return times, values, counts                                     # correlation.py:178-189
    for i in range(len(origins)):                                        # correlation.py:155-176
        act_result = np.moveaxis(values[i], cdim, 0)
        act_count = counts[i]
        if i == 0:
            count = act_count.copy()
            bt = (np.newaxis,)
            for j in range(act_result.ndim - 1):
                bt += (np.newaxis,) if j >= cdim else (slice(None),)
            result = act_result * act_count[bt]
        else:
            result += act_result * act_count[bt]
            count += act_count
    np.divide(result, count[bt], out=result, where=count[bt] != 0)
    result = np.moveaxis(result, 0, cdim)
    return times, result, count / len(origins)


# the autosave decorator of the reference (correlation.py:83-85): result files are keyed by the version tag; the
# source-based key of this package's own implementation is its name and version (like the functions above)
_fixed_checksum(shifted_correlation, 1)
shifted_correlation = autosave_data(nargs=2, kwargs_keys=(
    'index_distribution', 'correlation', 'segments', 'window', 'skip', 'average'
), version='shifted_correlation-1')(shifted_correlation)


def shifted_correlation_multi(functions, frames, index_distribution=log_indices, segments=10, window=0.5,
                              skip=None, average=False):
    """Several correlation functions over the SAME schedule from one pass over the coordinates.

    Extension of the reference API: ``functions`` is a dict ``name -> callable``; the kernel
    accumulates everything in one launch (e.g. van Hove histogram + non-Gaussian parameter,
    BASELINE config 3).  Returns ``(times, {name: result})``.
    """
    specs = {k: _kernel_spec(f) for k, f in functions.items()}
    start_frames, idx = _schedule(frames, index_distribution, segments, window, skip)
    frame_ids = _frame_ids(start_frames, idx, len(frames))
    qs = {sp.q for sp in specs.values() if sp.name == 'isf'}
    if len(qs) > 1:
        raise NotImplementedError('one q per launch')
    times = _times(frames, idx)
    sums, hist, n_atoms, float32 = _run(frames, frame_ids, list(specs.values()))
    out = {k: _average(sp.finalize(sums, hist, n_atoms, float32), len(start_frames), average)
           for k, sp in specs.items()}
    return times, out


def chi4(frames, q, index_distribution=log_indices, segments=20, window=0.2, skip=0.1, smooth=None):
    """Four-point susceptibility from the per-origin ISF spread (examples/plot_chi4.py:29-37):
    ``N * Fqt.var(axis=0)`` of the non-averaged ISF; ``smooth=k`` applies ``runningmean(chi4, k)``."""
    t, fqt = shifted_correlation(functools.partial(isf, q=q), frames, index_distribution=index_distribution,
                                 segments=segments, window=window, skip=skip, average=False)
    if getattr(frames, 'comm', None) is not None:
        n_atoms = frames.comm[1]
    elif hasattr(frames, '_resident'):
        n_atoms = frames._resident().n_atoms          # len(OW[0]) without fetching (and unwrapping) frame 0
    else:
        n_atoms = len(frames[0])
    # fqt.var(axis=0), operation by operation as NumPy's _var evaluates it (same bits), without its Python overhead
    fqt = np.asarray(fqt)
    if fqt.ndim == 2 and fqt.dtype == np.float64 and len(fqt):
        dev = fqt - np.add.reduce(fqt, axis=0, keepdims=True) / len(fqt)
        x4 = n_atoms * (np.add.reduce(dev * dev, axis=0) / len(fqt))
    else:
        x4 = n_atoms * fqt.var(axis=0)
    if smooth:
        return t, x4, runningmean(x4, smooth)
    return t, x4


# --------------------------------------------------------------------------------------------
# direct pair evaluation (calling msd(a, b) etc. like plain functions)
# --------------------------------------------------------------------------------------------
def _pair_store(start, frame):
    start = np.asarray(start)
    frame = np.asarray(frame)
    if start.shape != frame.shape:
        raise ValueError('start and frame must have the same shape')
    common = np.float32 if (start.dtype == np.float32 and frame.dtype == np.float32) else np.float64
    return _store_from_host_frames([start.astype(common, copy=False), frame.astype(common, copy=False)])


def _pair_value(spec, start, frame):
    store, float32 = _pair_store(start, frame)
    rows = np.array([[0, 1]], dtype=np.int32)
    sums, hist = engine.correlate_device(store, rows, spec.corr)
    sums = sums.cpu().numpy()
    hist = hist.cpu().numpy() if hist is not None else None
    value = spec.finalize(sums, hist, store.n_atoms, float32)[0, 1]
    return value if value.ndim else value[()]


def _pair_counts(start, end, bins):
    """Integer histogram of ``|start - end|`` (used by utils.histogram)."""
    spec = _Spec('van_hove_self', bins=bins)
    store, _ = _pair_store(start, end)
    rows = np.array([[0, 1]], dtype=np.int32)
    _, hist = engine.correlate_device(store, rows, spec.corr)
    return hist.cpu().numpy()[0, 1]
