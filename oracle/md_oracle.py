"""CPU oracle for the shifted-correlation hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a NumPy restatement of the algorithm of the reference package
(mdevaluate, read-only under /root/reference) for the one path this repository
accelerates.  It exists so that the CUDA path can be checked on machines where
the reference itself is absent (the GPU box).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it; nothing under ``mdevaluate_b200/`` does.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the real
reference (with stubbed I/O dependencies) in the build container, runs it on
seeded synthetic trajectories and stores inputs + outputs under
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every function
below against those vectors (bit-for-bit for integer results and for the float
results that use the same NumPy expressions).

Each function cites the reference lines it follows (paths relative to
/root/reference/).  Arithmetic is delegated to NumPy/SciPy exactly as the
reference does (numpy 2.3.5 / scipy 1.18.1 in this image; the reference pins
neither, ``setup.py:26``).
"""
from __future__ import annotations

from itertools import chain

import numpy as np

__all__ = [
    "log_indices", "lin_indices", "schedule", "pairwise", "shifted_correlation",
    "msd", "isf", "van_hove_self", "non_gaussian", "histogram", "runningmean",
    "chi4", "jump_table", "image_counts", "unwrap_frame", "TimedFrame",
    "ArrayTrajectory", "UnwrappedTrajectory", "nearest_distances", "NearestTree", "FrameTrees", "overlap",
]


# --------------------------------------------------------------------------- #
# lag / origin schedule
# --------------------------------------------------------------------------- #
def log_indices(first, last, num=100):
    """Log-spaced integer lags in [first, last].  mdevaluate/correlation.py:16-18."""
    spaced = np.logspace(0, np.log10(last - first + 1), num=num)
    return np.unique(np.int_(spaced) - 1 + first)


def lin_indices(first, last):
    """Dense linear lags (BASELINE config 5 passes ``lambda a, b: np.arange(a, b)``)."""
    return np.arange(first, last)


def schedule(n_frames, segments=10, window=0.5, skip=0.0):
    """Origins and window length.  mdevaluate/correlation.py:136-142.

    Returns ``(start_frames, num_frames)``; the caller applies
    ``index_distribution(0, num_frames)`` (correlation.py:144).
    """
    assert window + skip < 1
    starts = np.unique(np.linspace(
        n_frames * skip, n_frames * (1 - window),
        num=segments, endpoint=False, dtype=int,
    ))
    return starts, int(n_frames * window)


def pairwise(function, frames):
    """The default ``correlation`` strategy.  mdevaluate/correlation.py:21-24.

    The first frame of the iterator is the origin; the origin is also the first
    frame evaluated, so element 0 is always ``function(origin, origin)``.
    """
    it = iter(frames)
    origin = next(it)
    return map(lambda f: function(origin, f), chain([origin], it))


def subensemble_pairwise(selector_function, correlation_function=None):
    """``subensemble_correlation``: the origin frame picks the atoms.  mdevaluate/correlation.py:27-35."""
    inner = pairwise if correlation_function is None else correlation_function

    def strategy(function, frames):
        it = iter(frames)
        origin = next(it)
        selector = selector_function(origin)
        return inner(function, map(lambda f: f[selector], chain([origin], it)))
    return strategy


def multi_subensemble_pairwise(selector_function):
    """``multi_subensemble_correlation``: several sub-ensembles per origin, with their sizes.
    mdevaluate/correlation.py:38-81 (the strategy is flagged ``has_counter``)."""
    def strategy(function, frames):
        it = iter(frames)
        origin = next(it)
        selectors = np.asarray(selector_function(origin))
        sel_shape = selectors.shape
        if sel_shape[-1] == 0:
            selectors = np.asarray(selectors, int)
        if selectors.dtype != object:
            sel_shape = sel_shape[:-1]
        values = np.zeros(sel_shape + function(origin, origin).shape,)
        count = np.zeros(sel_shape, dtype=int)
        first = [True]

        def per_frame(frame):
            for index in np.ndindex(sel_shape):
                sel = selectors[index]
                picked = origin[sel]
                if first[0]:
                    count[index] = len(picked)
                values[index] = function(picked, frame[sel]) if count[index] != 0 else 0
            first[0] = False
            return np.asarray(values.copy())
        return map(per_frame, chain([origin], it)), count
    strategy.has_counter = True
    return strategy


def shifted_correlation(function, frames, index_distribution=log_indices,
                        segments=10, window=0.5, skip=None, average=False, correlation=None):
    """Driver.  mdevaluate/correlation.py:131-201 (``has_counter`` branch: 153-189)."""
    if correlation is not None and getattr(correlation, 'has_counter', False):
        return _shifted_correlation_counted(function, frames, index_distribution, segments, window, skip, average,
                                            correlation)
    strategy = pairwise if correlation is None else correlation
    if skip is None:
        try:
            skip = frames._slice.start / len(frames)
        except (TypeError, AttributeError):
            skip = 0
    starts, num_frames = schedule(len(frames), segments, window, skip)
    idx = index_distribution(0, num_frames)

    times = np.array([frames[i].time for i in idx]) - frames[0].time

    result = 0 if average else []
    for s in starts:
        row = list(strategy(function, map(frames.__getitem__, idx + s)))
        if average:
            result += np.array(row)
        else:
            result.append(row)
    result = np.array(result)
    if average:
        result = result / len(starts)
    return times, result


def _shifted_correlation_counted(function, frames, index_distribution, segments, window, skip, average, strategy):
    """The ``has_counter`` branch.  mdevaluate/correlation.py:153-189."""
    if skip is None:
        try:
            skip = frames._slice.start / len(frames)
        except (TypeError, AttributeError):
            skip = 0
    starts, num_frames = schedule(len(frames), segments, window, skip)
    idx = index_distribution(0, num_frames)
    times = np.array([frames[i].time for i in idx]) - frames[0].time

    def correlate(s):
        return strategy(function, map(frames.__getitem__, idx + s))

    if average:
        for i, s in enumerate(starts):
            act_result, act_count = correlate(s)
            act_result = np.array(list(act_result))
            act_count = np.array(act_count)
            if i == 0:
                count = act_count
                cdim = act_count.ndim
                rdim = act_result.ndim
                bt = (np.newaxis,)
                for j in range(rdim - 1):
                    bt += (np.newaxis,) if j >= cdim else (slice(None),)
                result = act_result * act_count[bt]
            else:
                result += act_result * act_count[bt]
                count += act_count
        np.divide(result, count[bt], out=result, where=count[bt] != 0)
        result = np.moveaxis(result, 0, cdim)
        count = count / len(starts)
        return times, result, count
    count = []
    result = []
    for s in starts:
        act_result, act_count = correlate(s)
        result.append(list(act_result))
        count.append(act_count)
    count = np.asarray(count)
    cdim = count.ndim
    result = np.moveaxis(np.asarray(result), 1, cdim)
    return times, result, count


# --------------------------------------------------------------------------- #
# per-pair kernels
# --------------------------------------------------------------------------- #
def msd(start, frame):
    """mdevaluate/correlation.py:203-208."""
    vec = start - frame
    return (vec ** 2).sum(axis=1).mean()


def isf(start, frame, q, box=None):
    """Isotropic incoherent scattering function.  mdevaluate/correlation.py:211-220."""
    vec = start - frame
    distance = (vec ** 2).sum(axis=1) ** .5
    return np.sinc(distance * q / np.pi).mean()


def rotational_autocorrelation(onset, frame, order=2):
    """mdevaluate/correlation.py:223-236."""
    from scipy.special import legendre
    scalar_prod = (onset * frame).sum(axis=-1)
    poly = legendre(order)
    return poly(scalar_prod).mean()


def runningmean(data, nav):
    """mdevaluate/utils.py:150-161."""
    return np.convolve(data, np.ones((nav,)) / nav, mode='valid')


def histogram(data, bins):
    """mdevaluate/utils.py:328-338 (bincount fast path / np.histogram)."""
    widths = np.diff(bins)
    dx = widths.mean()
    if bins.min() == 0 and widths.std() < 1e-6:
        hist = np.bincount((data // dx).astype(int), minlength=len(widths))[:len(widths)]
    else:
        hist = np.histogram(data, bins=bins)[0]
    return hist, runningmean(bins, 2)


def van_hove_self(start, end, bins):
    """mdevaluate/correlation.py:239-248."""
    vec = start - end
    delta_r = ((vec) ** 2).sum(axis=-1) ** .5
    return 1 / len(start) * histogram(delta_r, bins)[0]


def non_gaussian(onset, frame):
    """mdevaluate/correlation.py:355-362 (NaN at lag 0 is intended)."""
    r_2 = ((frame - onset) ** 2).sum(axis=-1)
    return 3 / 5 * (r_2 ** 2).mean() / r_2.mean() ** 2 - 1


def chi4(n_atoms, fqt):
    """examples/plot_chi4.py:29-37: ``len(OW[0]) * Fqt.var(axis=0)`` on the (S, L) ISF."""
    return n_atoms * np.asarray(fqt).var(axis=0)


# --------------------------------------------------------------------------- #
# overlap with a reference configuration (correlation.py:288-312 on coordinates.CoordinatesKDTree, :415-455)
# --------------------------------------------------------------------------- #
def nearest_distances(points, queries, boxsize=None, chunk=256):
    """Distance of every query point to its nearest tree point: what ``scipy.spatial.cKDTree(points,
    boxsize=boxsize).query(queries)[0]`` returns (k=1, p=2), restated by brute force.

    scipy (1.18.1 here; the reference pins no version, ``setup.py:26``) is a third-party dependency of the reference,
    not part of its tree.  Its arithmetic, from ckdtree/src (``distance_base.h``, ``distance.h``, ``query.cxx``):
    everything in float64; ``s = dx*dx; s += dy*dy; s += dz*dz`` (no fused multiply-add); a periodic tree wraps the
    query point, ``x - floor(x / L) * L`` (``BoxDist1D::wrap_position``), and every difference, ``d < -L/2: d + L;
    d > L/2: d - L`` (``wrap_distance``); the result is ``sqrt(min s)``.  ``tests/test_oracle_golden.py`` checks this
    function against scipy itself, bit for bit.
    """
    points = np.ascontiguousarray(points, dtype=np.float64)
    queries = np.ascontiguousarray(queries, dtype=np.float64)
    out = np.full(len(queries), np.inf)
    if len(points) == 0:
        return out
    box = None
    if boxsize is not None:
        box = np.broadcast_to(np.asarray(boxsize, dtype=np.float64), (3,))
        periodic = box > 0
        safe = np.where(periodic, box, 1.0)
        wrapped = queries - np.floor(queries / safe) * safe
        wrapped = np.where(wrapped >= safe, wrapped - safe, wrapped)
        wrapped = np.where(wrapped < 0, wrapped + safe, wrapped)
        queries = np.where(periodic, wrapped, queries)
        half = 0.5 * box
    for c0 in range(0, len(queries), chunk):
        d = queries[c0:c0 + chunk, None, :] - points[None, :, :]
        if box is not None:
            d = np.where(periodic & (d < -half), d + box, np.where(periodic & (d > half), d - box, d))
        s = d[..., 0] * d[..., 0]
        s = s + d[..., 1] * d[..., 1]
        s = s + d[..., 2] * d[..., 2]
        out[c0:c0 + chunk] = np.sqrt(s.min(axis=1))
    return out


class NearestTree:
    """Stands for one ``cKDTree``: ``.n`` points and ``.query(x)[0]`` nearest distances."""

    def __init__(self, points, boxsize=None):
        self.data = np.asarray(points, dtype=np.float64)
        self.n = len(self.data)
        self.boxsize = boxsize

    def query(self, x):
        return nearest_distances(self.data, x, self.boxsize), None


class FrameTrees:
    """coordinates.CoordinatesKDTree (coordinates.py:415-455): ``trees[i]`` is the tree of
    ``frames[i][selector(frames[i])]``."""

    def __init__(self, frames, selector=None, boxsize=None):
        self.frames = frames
        self.selector = selector if selector is not None else (lambda x: slice(None))
        self.boxsize = boxsize

    def __getitem__(self, index):
        frame = self.frames[index]
        return NearestTree(np.asarray(frame)[self.selector(frame)], self.boxsize)


def overlap(onset, frame, crds_tree, radius):
    """mdevaluate/correlation.py:288-312."""
    tree = crds_tree[onset.step]
    return (tree.query(frame)[0] <= radius).sum() / tree.n


# --------------------------------------------------------------------------- #
# nojump: producer (reader.parse_jumps) and consumer (pbc.nojump)
# --------------------------------------------------------------------------- #
def jump_table(positions, box_diag, whole0=None):
    """Per-step image jumps, dense.  mdevaluate/reader.py:165-186.

    ``positions`` is (T, M, 3) in the reader's dtype (float32), ``box_diag`` the
    diagonal of frame 0's box (reader.py:167: one box for ALL frames).  Row 0 is
    taken against ``frame0.whole`` (reader.py:166); ``whole0=None`` means every
    residue is a single atom, for which ``whole`` is the identity.
    Returns int8 (T, M, 3); the reference stores the same numbers as three
    (T x M) CSR matrices (reader.py:199-201).
    """
    positions = np.asarray(positions)
    box_diag = np.asarray(box_diag)
    prev = positions[0] if whole0 is None else np.asarray(whole0)
    out = np.empty(positions.shape, dtype=np.int8)
    for t in range(positions.shape[0]):
        curr = positions[t]
        out[t] = ((curr - prev) / box_diag).round().astype(np.int8)
        prev = curr
    return out


def image_counts(jumps):
    """Inclusive prefix sum over frames: ``m[:t+1].sum(axis=0)``.  mdevaluate/pbc.py:235-237,244-246."""
    return np.cumsum(jumps, axis=0, dtype=np.int64)


def unwrap_frame(frame, counts, box_diag):
    """``frame - counts * box`` in float64.  mdevaluate/pbc.py:235-237,247.

    ``counts`` is int64 (M, 3) (sum of the int8 CSR rows), ``box_diag`` the
    CURRENT frame's box diagonal in its own dtype (float32 from the reader):
    int64 * float32 promotes to float64, and so does float32 - float64.
    """
    delta = np.asarray(counts) * np.asarray(box_diag)
    return frame - delta


# --------------------------------------------------------------------------- #
# minimal frames protocol (coordinates.py:105-189, 257-293) for in-memory data
# --------------------------------------------------------------------------- #
class TimedFrame(np.ndarray):
    """(N, 3) array that carries ``.time`` / ``.step`` / ``.box`` like CoordinateFrame."""

    def __array_finalize__(self, obj):
        if obj is None:
            return
        self.time = getattr(obj, 'time', None)
        self.step = getattr(obj, 'step', None)
        self.box = getattr(obj, 'box', None)


class ArrayTrajectory:
    """Raw (wrapped) frames of an in-memory (T, N, 3) array; ``len`` is the full length."""

    def __init__(self, positions, box_diag, times=None, selection=None):
        self.positions = np.asarray(positions)
        self.box_diag = np.asarray(box_diag)
        T = self.positions.shape[0]
        self.times = np.arange(T, dtype=float) if times is None else np.asarray(times)
        self.selection = selection
        self._slice = slice(None)

    def __len__(self):
        return self.positions.shape[0]

    def _raw(self, i):
        frame = self.positions[i]
        if self.selection is not None:
            frame = frame[self.selection]
        return frame

    def _wrap(self, arr, i):
        out = np.asarray(arr).view(TimedFrame)
        out.time = self.times[i]
        out.step = i
        box = self.box_diag if self.box_diag.ndim == 1 else self.box_diag[i]
        out.box = np.diag(box)
        return out

    def __getitem__(self, i):
        if isinstance(i, slice):
            import copy
            sl = copy.copy(self)
            sl._slice = i
            return sl
        i = int(i)
        if i < 0:
            i += len(self)
        return self._wrap(self._raw(i), i)


class UnwrappedTrajectory(ArrayTrajectory):
    """``coords.nojump``: every frame is ``raw - image_counts * box`` (float64)."""

    def __init__(self, positions, box_diag, times=None, selection=None, whole0=None):
        super().__init__(positions, box_diag, times, selection)
        box0 = self.box_diag if self.box_diag.ndim == 1 else self.box_diag[0]
        self.jumps = jump_table(self.positions, box0, whole0)
        self.counts = image_counts(self.jumps)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return super().__getitem__(i)
        i = int(i)
        if i < 0:
            i += len(self)
        box = self.box_diag if self.box_diag.ndim == 1 else self.box_diag[i]
        full = unwrap_frame(self.positions[i], self.counts[i], box)
        if self.selection is not None:
            full = full[self.selection]
        return self._wrap(full, i)
