"""mdevaluate_b200 -- the shifted-correlation hot path of mdevaluate on NVIDIA B200 (sm_100a).

Drop-in for ``mdevaluate.correlation.shifted_correlation`` with ``msd`` / ``isf`` /
``van_hove_self`` / ``non_gaussian`` (+ chi4) and the ``nojump`` unwrapping they depend on::

    import mdevaluate_b200 as md
    coords = md.Coordinates.from_arrays(positions, box, times)          # (T, N, 3) float32
    t, S = md.correlation.shifted_correlation(partial(md.correlation.isf, q=22.7), coords.nojump,
                                              average=True)

Everything outside that path (file readers, selections, pair/tree analyses) is out of scope.
"""
from . import autosave, checksum, correlation, coordinates, pbc, reader, utils, parallel, synthetic  # noqa: F401
from .coordinates import Coordinates, CoordinateFrame  # noqa: F401
from .reader import MemoryReader, NojumpError  # noqa: F401
from ._cuda.lib import CudaPathError  # noqa: F401
from .logging import logger  # noqa: F401

__version__ = '0.1.0'
