"""The UNMODIFIED reference (mdevaluate) as the CPU arm of bench.py -- test/benchmark infrastructure only.

``install()`` (build container only) puts the reference package under ``baseline/_ref`` (git-ignored, shipped to the
GPU box with the repository snapshot).  ``pip install --target`` fails on the reference's ``version='dev'``
(``packaging.version.InvalidVersion``, recorded in DESIGN.md), so the package directory -- pure Python, no build
step -- is copied as it is, which is all a ``--target`` install of it would do.  Nothing under ``mdevaluate_b200/``
imports this module.

``load()`` imports it with stub modules for its absent I/O dependencies (``pygmx``, ``dask`` -- SURVEY.md section 8c;
none is touched by the hot path).  ``coordinates()`` wraps an in-memory trajectory in the reader protocol the
reference consumes and prepares the jump matrices with the reference's own ``reader.parse_jumps``.
"""
import os
import shutil
import subprocess
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, '_ref')
SOURCE = '/root/reference'


def install(force=False):
    """Copy the reference package to baseline/_ref (after trying the contract's pip command).  Returns a note."""
    target = os.path.join(REF_DIR, 'mdevaluate')
    if os.path.isdir(target) and not force:
        return 'present'
    if not os.path.isdir(os.path.join(SOURCE, 'mdevaluate')):
        return 'reference sources absent (%s)' % SOURCE
    os.makedirs(REF_DIR, exist_ok=True)
    tmp = '/tmp/mdevaluate_ref_copy'
    shutil.rmtree(tmp, ignore_errors=True)
    shutil.copytree(SOURCE, tmp)
    cmd = [sys.executable, '-m', 'pip', 'install', '--no-index', '--no-build-isolation', '--no-deps', '--find-links',
           '/opt/wheelhouse', '--target', REF_DIR, tmp]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode == 0 and os.path.isdir(target):
        return 'pip install --target ok'
    shutil.rmtree(target, ignore_errors=True)
    shutil.copytree(os.path.join(SOURCE, 'mdevaluate'), target)
    return 'pip refused the package (version "dev" is not PEP 440); package directory copied unmodified'


def available():
    return os.path.isdir(os.path.join(REF_DIR, 'mdevaluate'))


def load():
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
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import mdevaluate
    if not os.path.abspath(mdevaluate.__file__).startswith(REF_DIR):
        raise ImportError('mdevaluate resolved to %s, not to baseline/_ref' % mdevaluate.__file__)
    return mdevaluate


class _Frame:
    def __init__(self, positions, box, time):
        self.positions = positions
        self.box = box
        self.time = time
        self.coordinates = positions


class _Reader:
    """The reader protocol Coordinates consumes (reference reader.py:254-286)."""
    filename = '/nonexistent/synthetic.xtc'

    def __init__(self, positions, box_diag, times):
        self.pos = positions
        self.box_diag = box_diag
        self.times = times
        self._nojump_matrixes = None

    def __len__(self):
        return self.pos.shape[0]

    def __getitem__(self, i):
        return _Frame(self.pos[i], np.diag(self.box_diag).astype(self.pos.dtype), self.times[i])

    @property
    def nojump_matrixes(self):
        return self._nojump_matrixes

    @nojump_matrixes.setter
    def nojump_matrixes(self, mats):
        self._nojump_matrixes = mats


class _Atoms:
    def __init__(self, m):
        self.masses = np.ones(m)
        self.charges = np.zeros(m)


class _Subset:
    def __init__(self, m):
        self.selection = np.ones(m, bool)
        self.residue_ids = np.arange(1, m + 1)
        self.atoms = _Atoms(m)
        self.description = 'synthetic'


def coordinates(md, positions, box_diag, times):
    """``(Coordinates, seconds spent in parse_jumps)`` of the reference over an in-memory trajectory, with the jump
    matrices built by the reference itself (reader.py:165-201, without the npz write at :202)."""
    import time
    from scipy import sparse
    reader = _Reader(positions, box_diag, times)
    n_frames, n_atoms = positions.shape[:2]
    coords = md.coordinates.Coordinates(reader, atom_subset=_Subset(n_atoms))
    t0 = time.perf_counter()
    jumps = md.reader.parse_jumps(coords)
    reader.nojump_matrixes = tuple(
        sparse.csr_matrix((np.array(m.data), (np.array(m.row), np.array(m.col))), shape=(n_frames, n_atoms))
        for m in jumps)
    return coords, time.perf_counter() - t0
