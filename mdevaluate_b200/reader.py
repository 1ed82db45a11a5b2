"""Trajectory readers for the CUDA path and the producer side of nojump.

The reference's readers decode files (pygmx / MDAnalysis, reader.py:51-124) -- out of scope here.
``MemoryReader`` offers the same reader protocol (reader.py:254-286: ``len``, ``reader[i]`` ->
frame with ``.positions/.box/.time``, ``.filename``, ``.nojump_matrixes``) over an in-memory
trajectory that lives in host memory (NumPy, optionally pinned) or already in HBM (torch).

``parse_jumps`` / ``generate_nojump_matrixes`` keep the reference signatures
(reader.py:165-202) but run the jump detection on the GPU (csrc/nojump_kernels.cu).
"""
from __future__ import annotations

from collections import namedtuple

import numpy as np

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


SparseData = namedtuple('SparseData', ['data', 'row', 'col'])


def parse_jumps(trajectory):
    """Per-step periodic jumps of every atom, as three COO triplets (reference: reader.py:165-186).

    ``trajectory`` is a :class:`~mdevaluate_b200.coordinates.Coordinates`.  The detection
    ``round((x_t - x_{t-1}) / box_0)`` runs on the GPU (bit-exact int8, csrc/nojump_kernels.cu);
    the triplets are extracted from the dense result on the host for API compatibility.
    """
    jumps = trajectory._resident().dense_jumps()          # (T, N, 3) int8 on the host
    out = []
    for d in range(3):
        row, col = np.nonzero(jumps[:, :, d])
        out.append(SparseData(data=jumps[row, col, d], row=row, col=col))
    return tuple(out)


def generate_nojump_matrixes(trajectory):
    """Create the three CSR jump matrices and attach them to the reader (reader.py:189-202).

    The on-disk ``.nojump.npz`` cache of the reference (reader.py:205-241) is not written here.
    """
    from scipy import sparse
    logger.info('generate Nojump Matrixes for: {}'.format(trajectory))
    jump_data = parse_jumps(trajectory)
    n_frames = len(trajectory)
    n_atoms = trajectory._resident().n_atoms
    trajectory.frames.nojump_matrixes = tuple(
        sparse.csr_matrix((np.array(m.data), (m.row, m.col)), shape=(n_frames, n_atoms)) for m in jump_data
    )
