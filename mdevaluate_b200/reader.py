"""Trajectory readers for the CUDA path and the producer side of nojump.

The reference's readers decode files (pygmx / MDAnalysis, reader.py:51-124) -- out of scope here.
``MemoryReader`` offers the same reader protocol (reader.py:254-286: ``len``, ``reader[i]`` ->
frame with ``.positions/.box/.time``, ``.filename``, ``.nojump_matrixes``) over an in-memory
trajectory that lives in host memory (NumPy, optionally pinned) or already in HBM (torch).

``parse_jumps`` / ``generate_nojump_matrixes`` keep the reference signatures
(reader.py:165-202) but run the jump detection on the GPU (csrc/nojump_kernels.cu).
"""
from __future__ import annotations

import builtins
import os
import pickle
from collections import namedtuple
from zipfile import BadZipFile

import numpy as np

from .checksum import checksum
from .logging import logger


class NojumpError(Exception):
    """Raised when jump matrices are requested but were never generated (reader.py:43-44)."""


class MemoryFrame:
    """What ``reader[i]`` returns: positions (M, 3), box (3, 3) and time."""
    __slots__ = ('positions', 'box', 'time', 'step')

    def __init__(self, positions, box, time, step):
        self.positions = positions
        self.box = box
        self.time = time
        self.step = step

    @property
    def coordinates(self):  # coordinates.py:254 uses ``frames[0].coordinates`` for the default filter
        return self.positions


def _box_diagonals(box, n_frames):
    """Normalise a box argument to ``(constant?, diag array)``; keeps the dtype (float32 stays float32)."""
    box = np.asarray(box)
    if box.ndim == 1 and box.shape[0] == 3:
        return True, box.copy()
    if box.ndim == 2 and box.shape == (3, 3):
        return True, box.diagonal().copy()
    if box.ndim == 2 and box.shape == (n_frames, 3):
        return False, box.copy()
    if box.ndim == 3 and box.shape == (n_frames, 3, 3):
        return False, np.ascontiguousarray(np.diagonal(box, axis1=1, axis2=2))
    raise ValueError('box must have shape (3,), (3,3), (T,3) or (T,3,3)')


class MemoryReader:
    """Reader protocol over an in-memory trajectory.

    Args:
        positions: ``(T, M, 3)`` float32 NumPy array (host), or a torch CUDA tensor ``(T, pitch)``
            float32 holding ``3*M`` interleaved xyz values per row (then pass ``n_atoms``).
        box: box diagonal ``(3,)``, matrix ``(3,3)``, or per frame ``(T,3)`` / ``(T,3,3)``.
        times: ``(T,)`` frame times; default ``arange(T)``.
        whole0: optional ``(M, 3)`` positions of frame 0 with molecules made whole (the reference's
            ``trajectory[0].whole``, reader.py:166): the jumps of frame 0 are taken against it.  Default:
            frame 0 itself, which is what ``whole`` returns for single-atom residues.
    """

    def __init__(self, positions, box, times=None, n_atoms=None, filename='<memory>', whole0=None):
        self.filename = filename
        self._nojump_matrixes = None
        self.whole0 = None if whole0 is None else np.ascontiguousarray(whole0, dtype=np.float32)
        self.device_block = None
        self.host_block = None
        if isinstance(positions, np.ndarray):
            if positions.ndim != 3 or positions.shape[2] != 3:
                raise ValueError('positions must have shape (T, M, 3)')
            self.host_block = positions
            self.n_frames, self.n_atoms = positions.shape[:2]
            self.dtype = positions.dtype
        else:  # torch tensor on a CUDA device
            import torch
            if not (isinstance(positions, torch.Tensor) and positions.is_cuda and positions.dim() == 2):
                raise TypeError('positions must be a (T,M,3) NumPy array or a 2-D CUDA tensor (T, pitch)')
            if n_atoms is None:
                raise ValueError('n_atoms is required for a device block')
            if positions.shape[1] < 3 * n_atoms:
                raise ValueError('row pitch smaller than 3*n_atoms')
            self.device_block = positions
            self.n_frames, self.n_atoms = positions.shape[0], int(n_atoms)
            self.dtype = np.dtype('float32') if positions.dtype == torch.float32 else np.dtype('float64')
        self.box_constant, self.box_diag = _box_diagonals(box, self.n_frames)
        self.times = (np.arange(self.n_frames, dtype=np.float64) if times is None
                      else np.asarray(times, dtype=np.float64))
        if self.times.shape != (self.n_frames,):
            raise ValueError('times must have shape (T,)')
        # like BaseReader.__init__ (reader.py:283-285): pick up a stored nojump cache next to the trajectory
        if not filename.startswith('<') and os.path.exists(nojump_filename(self)):
            load_nojump_matrixes(self)

    def __checksum__(self):
        """Identity of the trajectory for the on-disk caches (the reference hashes the file name and the
        frame index of its file reader, reader.py:288-295): name, shape, box and the first and last frame."""
        def frame_bytes(i):
            return np.ascontiguousarray(self[i].positions).tobytes()
        return checksum(self.filename, str((self.n_frames, self.n_atoms)), np.asarray(self.box_diag),
                        frame_bytes(0), frame_bytes(self.n_frames - 1))

    # -- reader protocol ---------------------------------------------------------------------
    def __len__(self):
        return self.n_frames

    def box_of(self, i):
        return self.box_diag if self.box_constant else self.box_diag[i]

    def __getitem__(self, i):
        i = int(i)
        if i < 0:
            i += self.n_frames
        if not 0 <= i < self.n_frames:
            raise IndexError('frame %d out of range' % i)
        if self.host_block is not None:
            pos = self.host_block[i]
        else:
            pos = self.device_block[i, :3 * self.n_atoms].cpu().numpy().reshape(self.n_atoms, 3)
        diag = self.box_of(i)
        return MemoryFrame(pos, np.diag(diag).astype(diag.dtype), self.times[i], i)

    @property
    def nojump_matrixes(self):
        if self._nojump_matrixes is None:
            raise NojumpError('Nojump Data not available: {}'.format(self.filename))
        return self._nojump_matrixes

    @nojump_matrixes.setter
    def nojump_matrixes(self, mats):
        self._nojump_matrixes = mats


class SparseFrameReader:
    """Reader protocol over only the frames an analysis touches, held in page-locked host memory -- what a file reader
    with a frame index delivers cheaply -- plus the trajectory's jump matrices (``nojump_matrixes``, e.g. from the
    ``.nojump.npz`` cache), which make it complete for nojump: the GPU path then uploads just these frames
    (``_cuda/resident.py``: strided DMA out of ``pinned_rows``) and derives their image counts from the matrices.

    Args:
        frame_ids: ascending trajectory frames that are present.
        positions: ``(len(frame_ids), M, 3)`` float32 host array (page-locked for DMA speed).
        box, times: as for :class:`MemoryReader`; ``times`` covers ALL ``n_frames`` frames.
    """

    def __init__(self, frame_ids, positions, box, times, filename='<sparse>', host_cache=None):
        self.filename = filename
        self.frame_ids = np.ascontiguousarray(frame_ids, dtype=np.int64)
        if positions.ndim != 3 or positions.shape[2] != 3 or positions.shape[0] != len(self.frame_ids):
            raise ValueError('positions must have shape (len(frame_ids), M, 3)')
        self.positions = positions
        self.times = np.asarray(times, dtype=np.float64)
        self.n_frames = len(self.times)
        self.n_atoms = positions.shape[1]
        self.dtype = positions.dtype
        self.box_constant, self.box_diag = _box_diagonals(box, self.n_frames)
        self._slot = np.full((self.n_frames,), -1, dtype=np.int64)
        self._slot[self.frame_ids] = np.arange(len(self.frame_ids))
        self._nojump_matrixes = None
        self.device_block = None
        self.host_block = None
        self.whole0 = None
        self.host_cache = {} if host_cache is None else host_cache

    def fresh(self):
        """A new reader over the same host data (nothing cached on the GPU side is shared, the host-side caches are)."""
        other = SparseFrameReader(self.frame_ids, self.positions, self.box_diag, self.times, self.filename,
                                  host_cache=self.host_cache)
        other._nojump_matrixes = self._nojump_matrixes
        return other

    def __len__(self):
        return self.n_frames

    def box_of(self, i):
        return self.box_diag if self.box_constant else self.box_diag[i]

    def __getitem__(self, i):
        i = int(i)
        if i < 0:
            i += self.n_frames
        if not 0 <= i < self.n_frames:
            raise IndexError('frame %d out of range' % i)
        slot = self._slot[i]
        if slot < 0:
            if i == 0:         # Coordinates only asks frame 0 for the number of atoms
                return MemoryFrame(np.zeros((self.n_atoms, 3), dtype=self.dtype), np.diag(self.box_of(0)),
                                   self.times[0], 0)
            raise IndexError('frame %d is not held by this reader' % i)
        diag = self.box_of(i)
        return MemoryFrame(self.positions[slot], np.diag(diag).astype(diag.dtype), self.times[i], i)

    def pinned_rows(self, frame_ids):
        """Page-locked ``(n, 3M)`` torch view of exactly these frames, or None (not all present in one run / not
        page-locked)."""
        import torch
        slots = self._slot[np.asarray(frame_ids, dtype=np.int64)]
        if len(slots) == 0 or slots[0] < 0 or np.any(np.diff(slots) != 1):
            return None
        flat = torch.from_numpy(self.positions).view(len(self.frame_ids), 3 * self.n_atoms)
        return flat[int(slots[0]):int(slots[-1]) + 1] if flat.is_pinned() else None

    @property
    def nojump_matrixes(self):
        if self._nojump_matrixes is None:
            raise NojumpError('Nojump Data not available: {}'.format(self.filename))
        return self._nojump_matrixes

    @nojump_matrixes.setter
    def nojump_matrixes(self, mats):
        self._nojump_matrixes = mats


SparseData = namedtuple('SparseData', ['data', 'row', 'col'])


def parse_jumps(trajectory):
    """Per-step periodic jumps of every atom, as three COO triplets (reference: reader.py:165-186).

    ``trajectory`` is a :class:`~mdevaluate_b200.coordinates.Coordinates`.  The detection
    ``round((x_t - x_{t-1}) / box_0)`` runs on the GPU (bit-exact int8, csrc/nojump_kernels.cu);
    the triplets are extracted from the dense result on the host for API compatibility.
    """
    frames, comps, vals = trajectory._resident().sparse_jumps()      # non-zero jumps, sorted by frame and component
    out = []
    for d in range(3):
        pick = (comps % 3) == d
        out.append(SparseData(data=vals[pick], row=frames[pick].astype(np.int64), col=comps[pick] // 3))
    return tuple(out)


def generate_nojump_matrixes(trajectory):
    """Create the three CSR jump matrices, attach them to the reader and -- when the trajectory has a
    file name -- store them in the reference's cache format next to it (reader.py:189-202)."""
    from scipy import sparse
    logger.info('generate Nojump Matrixes for: {}'.format(trajectory))
    res = trajectory._resident()
    if res.n_atoms != res.n_atoms_all:
        # the matrices belong to the READER and are indexed by each Coordinates' selection afterwards
        # (pbc.py:243), so they always cover all atoms, whatever subset this Coordinates shows
        from .coordinates import Coordinates
        trajectory = Coordinates(trajectory.frames)
        res = trajectory._resident()
    jump_data = parse_jumps(trajectory)
    n_frames = len(trajectory)
    n_atoms = res.n_atoms
    reader = trajectory.frames
    reader.nojump_matrixes = tuple(
        sparse.csr_matrix((np.array(m.data), (m.row, m.col)), shape=(n_frames, n_atoms)) for m in jump_data
    )
    if not str(getattr(reader, 'filename', '<memory>')).startswith('<'):
        save_nojump_matrixes(reader)


# --------------------------------------------------------------------------------------------------
# the on-disk nojump cache `.{trajectory}.nojump.npz` (reader.py:125-241)
# --------------------------------------------------------------------------------------------------
CSR_ATTRS = ('data', 'indices', 'indptr')
NOJUMP_MAGIC = 2016


def is_writeable(fname):
    """Whether a file can be created next to ``fname`` -- tested by creating one (reader.py:125-140)."""
    fdir = os.path.dirname(fname)
    if not os.access(fdir, os.W_OK):
        return False
    probe = os.path.join(fdir, str(np.random.randint(999999999)))
    while os.path.exists(probe):
        probe = os.path.join(fdir, str(np.random.randint(999999999)))
    try:
        with builtins.open(probe, 'w'):
            pass
        os.remove(probe)
        return True
    except PermissionError:
        return False


def nojump_filename(reader):
    """``<dir>/.<trajectory>.nojump.npz``, or the same below ``~/.mdevaluate/nojump`` when the trajectory's
    directory is read-only (reader.py:143-158)."""
    directory, fname = os.path.split(reader.filename)
    local = os.path.join(directory, '.{}.nojump.npz'.format(fname))
    if os.path.exists(local) or is_writeable(directory):
        return local
    fallback = os.path.join(os.path.join(os.environ['HOME'], '.mdevaluate/nojump'), directory.lstrip('/'),
                            '.{}.nojump.npz'.format(local))
    logger.info('Saving nojump to {}, since original location is not writeable.'.format(fallback))
    os.makedirs(os.path.dirname(fallback), exist_ok=True)
    return fallback


def _nojump_checksum(reader):
    return checksum(NOJUMP_MAGIC, checksum(reader))


def save_nojump_matrixes(reader, matrixes=None):
    """Write the jump matrices in the reference's format (reader.py:205-213): ``checksum``, ``shape`` and
    ``data_d / indices_d / indptr_d`` of the three CSR matrices, uncompressed npz -- plus ``checksum_str``, the
    checksum in decimal, which is what this package reads back (no pickle)."""
    if matrixes is None:
        matrixes = reader.nojump_matrixes
    csum = _nojump_checksum(reader)
    data = {'checksum': csum, 'checksum_str': np.array(str(int(csum)))}
    for d, mat in enumerate(matrixes):
        data['shape'] = mat.shape
        for attr in CSR_ATTRS:
            data['{}_{}'.format(attr, d)] = getattr(mat, attr)
    np.savez(nojump_filename(reader), **data)


class PickledChecksum(Exception):
    """The cache file carries its checksum only as a pickled object (a file the reference wrote)."""


def read_nojump_file(filename, trust_pickle=False):
    """``(checksum, (Jx, Jy, Jz))`` of a ``.nojump.npz`` file written by this package or by the reference.

    The checksum is a 160-bit Python integer.  This package stores it in decimal (``checksum_str``) and reads the
    file with ``allow_pickle=False``.  The reference stores it as a pickled object entry only: unpickling runs code
    from the file, so that happens solely with ``trust_pickle=True`` -- for cache files you (or the reference on
    your behalf) wrote.  Without it such a file raises :class:`PickledChecksum`; the matrices themselves are plain
    arrays and never need pickle."""
    from scipy import sparse
    with np.load(filename, allow_pickle=False) as data:
        shape = tuple(int(n) for n in data['shape'])
        mats = tuple(
            sparse.csr_matrix(tuple(data['{}_{}'.format(attr, d)] for attr in CSR_ATTRS), shape=shape)
            for d in range(3)
        )
        if 'checksum_str' in data:
            return int(str(data['checksum_str'][()])), mats
        if 'checksum' not in data:
            return None, mats
    if not trust_pickle:
        raise PickledChecksum(filename)
    with np.load(filename, allow_pickle=True) as data:
        return data['checksum'][()], mats


def load_nojump_matrixes(reader, filename=None, verify=True, trust_pickle=False):
    """Attach the stored jump matrices to ``reader`` if the file belongs to it (reader.py:216-241): a
    file with another checksum is ignored, an unreadable one is removed.  ``filename`` reads a specific
    file; ``verify=False`` accepts a foreign checksum (a cache the reference wrote through its own file
    reader hashes that reader's frame index, which only the reference can recompute) and does not look at the
    stored checksum at all.  ``trust_pickle=True`` allows unpickling the checksum entry of a file the reference
    wrote (see :func:`read_nojump_file`); the automatic pick-up in ``MemoryReader.__init__`` never does.  Returns
    True when the matrices were attached."""
    zipname = nojump_filename(reader) if filename is None else filename
    try:
        try:
            stored, mats = read_nojump_file(zipname, trust_pickle=trust_pickle)
        except PickledChecksum:
            if verify:
                logger.info('Ignoring %s: its checksum is a pickled entry (pass trust_pickle=True to '
                            'load_nojump_matrixes for cache files the reference wrote)', zipname)
                return False
            stored, mats = read_nojump_file_matrices(zipname)
    except (AttributeError, BadZipFile, OSError, KeyError, ValueError, EOFError, pickle.UnpicklingError):
        logger.info('Removing zip-File: %s', zipname)
        if filename is None:
            os.remove(zipname)
        return False
    if verify and stored != _nojump_checksum(reader):
        logger.info('Invlaid Nojump Data: {}'.format(zipname))
        return False
    reader.nojump_matrixes = mats
    logger.info('Loaded Nojump Matrixes: {}'.format(zipname))
    return True


def read_nojump_file_matrices(filename):
    """``(None, matrices)``: the three CSR matrices of a cache file without touching its checksum entry."""
    from scipy import sparse
    with np.load(filename, allow_pickle=False) as data:
        shape = tuple(int(n) for n in data['shape'])
        mats = tuple(
            sparse.csr_matrix(tuple(data['{}_{}'.format(attr, d)] for attr in CSR_ATTRS), shape=shape)
            for d in range(3)
        )
    return None, mats


def correct_nojump_matrixes_for_whole(trajectory):
    """Row 0 of the matrices := jumps of frame 0 against ``frame0.whole`` (reader.py:244-251)."""
    reader = trajectory.frames
    frame = trajectory[0]
    box = frame.box.diagonal()
    cor = ((frame - frame.whole) / box).round().astype(np.int8)
    mats = tuple(m.tolil() for m in reader.nojump_matrixes)
    for d in range(3):
        mats[d][0] = cor[:, d]
    reader.nojump_matrixes = tuple(m.tocsr() for m in mats)
    if not str(getattr(reader, 'filename', '<memory>')).startswith('<'):
        save_nojump_matrixes(reader)


def jump_columns(matrixes, atom_filter=None):
    """The three jump matrices (frames x atoms, any scipy sparse format) as ONE column-compressed list over
    the interleaved components ``c = 3*atom + d`` -- the layout ``mdb_nojump_images_from_jumps`` walks:
    ``indptr`` int64 ``(3N+1,)``, ``frames`` int32 ``(nnz,)`` ascending inside a column, ``jumps`` int8.
    ``atom_filter`` (index array) selects columns first, like ``m[..., selection]`` in pbc.py:243."""
    from scipy import sparse
    cols = []
    for m in matrixes:
        m = sparse.csc_matrix(m)
        if atom_filter is not None:
            m = m[:, np.asarray(atom_filter)]
        m.sort_indices()
        if m.nnz and (np.abs(m.data.astype(np.int64)) > 127).any():
            raise ValueError('jump matrices must hold int8 jumps')
        cols.append(m)
    n_atoms = cols[0].shape[1]
    counts = np.stack([np.diff(m.indptr) for m in cols], axis=1).reshape(-1).astype(np.int64)
    indptr = np.zeros(3 * n_atoms + 1, dtype=np.int64)
    np.cumsum(counts, out=indptr[1:])
    frames = np.empty(int(indptr[-1]), dtype=np.int32)
    jumps = np.empty(int(indptr[-1]), dtype=np.int8)
    for d, m in enumerate(cols):
        per_col = np.diff(m.indptr)
        col_of = np.repeat(np.arange(n_atoms, dtype=np.int64), per_col)
        dest = indptr[3 * col_of + d] + (np.arange(m.nnz, dtype=np.int64) - m.indptr[col_of])
        frames[dest] = m.indices
        jumps[dest] = m.data
    return indptr, frames, jumps
