"""Host-side helpers that sit at the edge of the hot path (reference: mdevaluate/utils.py).

Only tiny O(bins) / O(lags) arrays are touched here; the per-atom work of ``histogram`` runs in
the CUDA kernel (csrc/corr_kernels.cu) with the binning rule selected by :func:`bin_rule`.
"""
import numpy as np

from ._cuda import lib as _lib


def runningmean(data, nav):
    """Running mean over ``nav`` points, shape ``(N-(nav-1),)`` (utils.py:150-161)."""
    return np.convolve(data, np.ones((nav,)) / nav, mode='valid')


def mask2indices(mask):
    """Selected indices of a mask; for a 2-D mask along the second axis (utils.py:102-118)."""
    mask = np.array(mask)
    if len(mask.shape) == 1:
        indices = np.where(mask)
    else:
        indices = np.array([np.where(m) for m in mask])
    return indices


def bin_rule(bins):
    """Which rule ``utils.histogram`` (utils.py:328-338) applies to these bin edges.

    Returns ``(mode, dx, n_bins)``: the bincount fast path ``(r // dx)`` when the edges start at 0
    and are equally spaced (``np.diff(bins).std() < 1e-6``), otherwise ``np.histogram`` semantics.
    ``dx`` is ``np.diff(bins).mean()`` exactly as the reference computes it.
    """
    bins = np.asarray(bins)
    widths = np.diff(bins)
    dx = widths.mean()
    if bins.min() == 0 and widths.std() < 1e-6:
        return _lib.BINS_UNIFORM, float(dx), len(widths)
    return _lib.BINS_EDGES, float(dx), len(widths)


def histogram(data, bins):
    """Histogram of ``data`` with the reference's rule; counts come from the CUDA kernel.

    ``data`` is treated as distances from the origin (the kernel histograms ``|r|`` of
    displacement vectors), so it must be non-negative.  Returns ``(hist, bin_centres)``.
    """
    from . import correlation
    data = np.asarray(data)
    if data.ndim != 1:
        raise ValueError('histogram expects 1-D data')
    if data.size and data.min() < 0:
        raise ValueError('the CUDA histogram bins distances; negative data is not supported')
    start = np.zeros((data.size, 3), dtype=data.dtype if data.dtype in (np.float32, np.float64) else np.float64)
    end = start.copy()
    end[:, 0] = data
    hist = correlation._pair_counts(start, end, np.asarray(bins))
    return hist, runningmean(np.asarray(bins), 2)
