"""Coordinates / CoordinateFrame: the frames protocol ``shifted_correlation`` consumes.

Mirrors the part of mdevaluate/coordinates.py the hot path touches (``CoordinateFrame``
105-189, ``Coordinates`` 192-311): ``len`` is the FULL trajectory length even for a sliced
object (292-293), ``coords[i]`` is a ``CoordinateFrame`` with ``.time/.box/.step/.mode``,
``coords[a:b:c]`` is a copy carrying ``_slice`` (286-290), ``coords.nojump`` is a copy in
``'nojump'`` mode (199-215).  Selections, maps and KD-trees are out of scope.

Frames handed back to Python are host NumPy arrays; the correlation kernels never go through
them -- ``shifted_correlation`` asks ``_resident()`` for the HBM store directly.
"""
from copy import copy

import numpy as np

from .logging import logger
from .pbc import nojump


class UnknownCoordinatesMode(Exception):
    pass


def spatial_selector(frame, transform, rmin, rmax):
    """Atoms with ``rmin <= transform(frame) <= rmax`` (coordinates.py:85-102); a selector for
    ``correlation.subensemble_correlation``.  Runs on the host frame of each origin."""
    from .utils import mask2indices
    r = transform(frame)
    selector = (rmin <= r) & (rmax >= r)
    return mask2indices(selector)


def spherical_radius(frame, origin=None):
    """Distance of every atom from ``origin`` (default: box centre) (coordinates.py:49-56)."""
    if origin is None:
        origin = frame.box.diagonal() / 2
    return ((frame - origin) ** 2).sum(axis=-1) ** 0.5


class CoordinateFrame(np.ndarray):
    _known_modes = ('pbc', 'whole', 'nojump')

    @property
    def box(self):
        return np.array(self.coordinates.frames[self.step].box)

    @property
    def volume(self):
        return self.box.diagonal().cumprod()[-1]

    @property
    def time(self):
        return self.coordinates.frames[self.step].time

    @property
    def selection(self):
        return self.coordinates.atom_filter

    @property
    def pbc(self):
        frame = self % self.box.diagonal()
        frame.mode = 'pbc'
        return frame

    @property
    def whole(self):
        # single-atom residues: ``whole`` is the identity (pbc.py:163-209 only moves bonded atoms)
        frame = self.copy()
        frame.mode = 'whole'
        return frame

    @property
    def nojump(self):
        if self.mode != 'nojump':
            if self.mode is not None:
                logger.warning('Combining Nojump with other Coordinate modes is not supported and '
                               'may cause unexpected results.')
            frame = nojump(self)
            frame.mode = 'nojump'
            return frame
        return self

    def __new__(subtype, shape, dtype=float, buffer=None, offset=0, strides=None, order=None,
                coordinates=None, step=None, box=None, mode=None):
        obj = np.ndarray.__new__(subtype, shape, dtype, buffer, offset, strides)
        obj.coordinates = coordinates
        obj.step = step
        obj.mode = mode
        return obj

    def __array_finalize__(self, obj):
        if obj is None:
            return
        self.coordinates = getattr(obj, 'coordinates', None)
        self.step = getattr(obj, 'step', None)
        self.mode = getattr(obj, 'mode', None)


class Coordinates:
    """Trajectory data for the evaluation functions, backed by an HBM store.

    Args:
        frames: a reader (``len``, ``reader[i].positions/.box/.time``), e.g. ``reader.MemoryReader``
        atom_filter (opt.): boolean mask or index array selecting atoms
        atom_subset (opt.): object with a ``.selection`` mask (the reference's AtomSubset)
        mode (opt.): ``None`` or ``'nojump'``
    """

    def __init__(self, frames, atom_filter=None, atom_subset=None, mode=None):
        self._mode = mode
        self.frames = frames
        self._slice = slice(None)
        self.comm = None            # set by parallel.shard_atoms: (process group, global atom count)
        assert atom_filter is None or atom_subset is None, "Cannot use both: subset and filter"
        if atom_filter is not None:
            self.atom_filter = np.asarray(atom_filter)
            self.atom_subset = None
        elif atom_subset is not None:
            self.atom_filter = np.asarray(atom_subset.selection)
            self.atom_subset = atom_subset
            self.atoms = getattr(atom_subset, 'atoms', None)
        else:
            self.atom_filter = np.ones(shape=(len(frames[0].coordinates),), dtype=bool)
            self.atom_subset = None

    @classmethod
    def from_arrays(cls, positions, box, times=None, n_atoms=None, mode=None):
        """Build Coordinates over an in-memory trajectory (host NumPy or a CUDA tensor)."""
        from .reader import MemoryReader
        return cls(MemoryReader(positions, box, times, n_atoms=n_atoms), mode=mode)

    # -- modes ------------------------------------------------------------------------------
    def get_mode(self, mode):
        if mode not in (None, 'nojump'):
            raise UnknownCoordinatesMode('mode %r is outside the CUDA path (only None / nojump)' % (mode,))
        other = copy(self)
        other._mode = mode
        return other

    @property
    def nojump(self):
        return self.get_mode('nojump')

    @property
    def mode(self):
        return self._mode

    @mode.setter
    def mode(self, val):
        if val in CoordinateFrame._known_modes:
            logger.warning('Changing the Coordinates mode directly is deprecated. '
                           'Use Coordinates.%s instead, which returns a copy.', val)
            if val != 'nojump':
                raise UnknownCoordinatesMode('mode %r is outside the CUDA path' % (val,))
            self._mode = val
        else:
            raise UnknownCoordinatesMode('No such mode: {}'.format(val))

    # -- HBM residency ------------------------------------------------------------------------
    def _resident(self):
        """The HBM residency manager shared by every view of (reader, selection)."""
        res = self.__dict__.get('_resident_obj')
        if res is not None:
            return res
        from ._cuda.resident import Resident
        cache = self.frames.__dict__.setdefault('_mdb_resident', {}) if hasattr(self.frames, '__dict__') else {}
        key = None
        if self.atom_filter is not None:
            # the atoms themselves, not the bytes of whatever array selects them: a bool mask and an index array
            # (or arrays of different shape) with identical bytes must not share a store
            af = np.asarray(self.atom_filter)
            index = np.nonzero(af)[0] if af.dtype == bool else af.astype(np.int64).ravel()
            key = ('atoms', len(index), index.tobytes())
        res = cache.get(key)
        if res is None:
            res = Resident(self.frames, self.atom_filter)
            cache[key] = res
        self._resident_obj = res      # copies of this object (slices, mode views) inherit the handle
        return res

    # -- frames protocol ----------------------------------------------------------------------
    def get_frame(self, fnr):
        fnr = int(fnr)
        n = len(self)
        if not -n <= fnr < n:
            raise IndexError('frame %d out of range' % fnr)
        pos = np.asarray(self.frames[fnr].positions)
        if self.atom_filter is not None:
            pos = pos[self.atom_filter]
        frame = pos.view(CoordinateFrame)
        frame.coordinates = self
        frame.step = fnr
        if self.mode is not None:
            frame = getattr(frame, self.mode)
        return frame

    def __iter__(self):
        for i in range(len(self))[self._slice]:
            yield self[i]

    def __getitem__(self, item):
        if isinstance(item, slice):
            sliced = copy(self)
            sliced._slice = item
            return sliced
        return self.get_frame(item)

    def __len__(self):
        return len(self.frames)

    def __checksum__(self):
        # coordinates.py:295-296
        from .checksum import checksum
        return checksum(self.frames, self.atom_filter, self._slice, self.mode)

    def __repr__(self):
        return "Coordinates <{}>: {} atoms, mode={}".format(
            getattr(self.frames, 'filename', '?'), int(np.count_nonzero(self.atom_filter))
            if self.atom_filter.dtype == bool else len(self.atom_filter), self._mode)

    @property
    def description(self):
        return getattr(self.atom_subset, 'description', '')


class CoordinatesMap:
    """A function applied to every frame of a Coordinates object (coordinates.py:332-393): ``map[i]`` is
    ``function(coordinates[i])`` as a :class:`CoordinateFrame`.  ``shifted_correlation`` pulls such frames through the
    generic frames protocol -- except for :func:`vectors`, whose rows are built on the device (``mdb_vectors``)."""

    def __init__(self, coordinates, function):
        from functools import partial
        self.coordinates = coordinates
        self.frames = self.coordinates.frames
        self.atom_subset = self.coordinates.atom_subset
        self.function = function
        if isinstance(function, partial):
            self._description = self.function.func.__name__
        else:
            self._description = self.function.__name__

    def __iter__(self):
        for frame in self.coordinates:
            step = frame.step
            frame = self.function(frame)
            if not isinstance(frame, CoordinateFrame):
                frame = frame.view(CoordinateFrame)
                frame.coordinates = self
                frame.step = step
            yield frame

    def __getitem__(self, item):
        if isinstance(item, slice):
            return self.__class__(self.coordinates[item], self.function)
        frame = self.function(self.coordinates.__getitem__(item))
        if not isinstance(frame, CoordinateFrame):
            frame = frame.view(CoordinateFrame)
        frame.coordinates = self
        frame.step = item
        return frame

    def __len__(self):
        return len(self.coordinates.frames)

    def __checksum__(self):
        from .checksum import checksum
        return checksum(self.coordinates, self.function)

    @property
    def _slice(self):
        return self.coordinates._slice

    @property
    def comm(self):
        return self.coordinates.comm

    @property
    def description(self):
        return '{}_{}'.format(self._description, self.coordinates.description)

    @description.setter
    def description(self, desc):
        self._description = desc

    @property
    def nojump(self):
        return CoordinatesMap(self.coordinates.nojump, self.function)


class DeviceTree:
    """What ``CoordinatesKDTree[i]`` hands out: the selected atoms of frame i -- the points scipy's cKDTree would be
    built from -- with the one query ``overlap`` needs, evaluated on the device (``mdb_overlap``)."""

    def __init__(self, points, boxsize=None):
        self.data = np.asarray(points)
        if self.data.ndim != 2 or self.data.shape[1] != 3:
            raise ValueError('a tree holds (n, 3) points')
        self.n = len(self.data)
        self.boxsize = None if boxsize is None else np.broadcast_to(np.asarray(boxsize, dtype=np.float64), (3,))

    def count_within(self, frame, radius):
        """``(tree.query(frame)[0] <= radius).sum()`` of the cKDTree this object stands for."""
        from . import correlation
        return correlation._overlap_pair(self, frame, radius)


class CoordinatesKDTree:
    """The trees of the reference configurations ``overlap`` compares with (coordinates.py:415-461): same
    constructor, ``tree[i]`` is built from ``frames[i][selector(frames[i])]``, cached by an LRU cache.

    There is no KD-tree on the device: :func:`mdevaluate_b200.correlation.shifted_correlation` recognises
    ``partial(overlap, crds_tree=CoordinatesKDTree(...), radius=...)`` and evaluates the whole schedule with a cell
    grid per origin (``csrc/overlap_kernels.cu``), in cKDTree's arithmetic.  ``ckdtree`` is accepted and ignored:
    both scipy trees return the same nearest distances."""

    def clear_cache(self):
        """Clear the LRU cache."""
        self._get_tree_at_index.cache_clear()

    @property
    def cache_info(self):
        """Return info about the state of the cache."""
        return self._get_tree_at_index.cache_info()

    def _get_tree_at_index(self, index):
        frame = self.frames[index]
        return DeviceTree(np.asarray(frame)[self.selector(frame)], self.boxsize)

    def __init__(self, frames, selector=None, boxsize=None, maxcache=128, ckdtree=True):
        from functools import lru_cache
        self.has_selector = selector is not None
        self.selector = selector if selector is not None else (lambda x: slice(None))
        self.frames = frames
        self.boxsize = boxsize
        self._get_tree_at_index = lru_cache(maxsize=maxcache)(self._get_tree_at_index)

    def __getitem__(self, index):
        return self._get_tree_at_index(index)

    def __checksum__(self):
        from .checksum import checksum
        return checksum(self.selector if self.has_selector else None, self.frames,
                        None if self.boxsize is None else np.asarray(self.boxsize, dtype=np.float64))


def map_coordinates(func):
    """Decorator: ``func(coordinates, **kwargs)`` becomes a :class:`CoordinatesMap` of ``func`` over the frames
    (coordinates.py:464-468)."""
    from functools import partial, wraps

    @wraps(func)
    def wrapped(coordinates, **kwargs):
        return CoordinatesMap(coordinates, partial(func, **kwargs))
    wrapped.frame_function = func
    return wrapped


def _pbc_diff(v1, v2, box):
    """pbc.pbc_diff for no box / a rectangular box (pbc.py:34-56)."""
    if box is None:
        return v1 - v2
    box = np.asarray(box)
    if box.ndim != 1:
        raise NotImplementedError('triclinic boxes are outside this package')
    v = v1 - v2
    s = v / box
    return box * (s - s.round())


@map_coordinates
def vectors(coordinates, atoms_a, atoms_b, normed=False, box=None):
    """Vectors between the atoms of two subsets (coordinates.py:524-563).  ``atoms_a`` / ``atoms_b`` are any NumPy
    subscripts of a frame; a two-dimensional index list takes the mean over its first axis.  ``box`` (a box diagonal)
    applies the minimum image convention.  Used as ``vectors(coords, atoms_a=..., atoms_b=..., normed=True)`` it is
    the frames argument of ``shifted_correlation(partial(rotational_autocorrelation, order=2), ...)``; the rows are
    then computed on the device."""
    coords_a = coordinates[atoms_a]
    if len(coords_a.shape) > 2:
        coords_a = coords_a.mean(axis=0)
    coords_b = coordinates[atoms_b]
    if len(coords_b.shape) > 2:
        coords_b = coords_b.mean(axis=0)
    vec = _pbc_diff(coords_a, coords_b, box=box)
    norm = np.linalg.norm(vec, axis=-1).reshape(-1, 1) if normed else 1
    vec.reference = coords_a
    return vec / norm
