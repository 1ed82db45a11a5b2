"""nojump unwrapping, consumer side (reference: mdevaluate/pbc.py:212-247).

The reference rebuilds ``delta = cumsum(jumps) * box`` on the host with a 128-entry cache per
reader; here the cumulative image counts live in HBM next to the coordinates (see
``_cuda/resident.py``) and a frame is unwrapped by one small kernel.
"""
from .logging import logger

NOJUMP_CACHESIZE = 128   # kept for API compatibility; the HBM store replaces the delta cache


def nojump(frame, usecache=True):
    """Return the nojump coordinates of a :class:`CoordinateFrame` as float64 (pbc.py:215-247)."""
    coords = frame.coordinates
    if coords is None or frame.step is None:
        raise ValueError('nojump needs a CoordinateFrame that knows its Coordinates and step')
    step = frame.step % len(coords)
    out = coords._resident().unwrapped_frame(step).view(type(frame))
    out.coordinates = coords
    out.step = frame.step
    out.mode = frame.mode
    return out
