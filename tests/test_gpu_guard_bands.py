"""Out-of-bounds check of every kernel family without compute-sanitizer (the pool refuses to run it:
profiles/r2_compute_sanitizer_closed.log).

Every device buffer the host side allocates through torch -- frame rows, image counts, jump rows, select masks,
workspaces, sums, histograms, the overlap kernel's cell lists -- is placed between two 4 KiB guard bands filled with a
canary, and its body is poisoned (0xFF bytes: NaN as a float, -1 as a count) unless the caller asked for zeros.
A store outside a buffer hits a canary; a load outside a buffer (or of padding the kernel should mask) reads poison
and shows up in the comparison with the oracle.  Shapes are chosen off every tile size (atoms not a multiple of 4,
64 or 1024; rows not a multiple of the ring depth)."""
from functools import partial

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GUARD = 4096
CANARY = 0xA5


class GuardedTorch:
    """Stands in for the torch module inside mdevaluate_b200: device allocations get guard bands."""

    def __init__(self, torch):
        self._t = torch
        self.blocks = []

    def __getattr__(self, name):
        return getattr(self._t, name)

    def _guarded(self, size, dtype, device, zero):
        t = self._t
        if len(size) == 1 and isinstance(size[0], (tuple, list, t.Size)):
            size = tuple(size[0])
        dtype = dtype or t.float32
        n = int(np.prod(size, dtype=np.int64)) * t.empty((), dtype=dtype).element_size()
        padded = (n + 255) // 256 * 256
        raw = t.full((padded + 2 * GUARD,), CANARY, dtype=t.uint8, device=device)
        body = raw[GUARD:GUARD + n]
        body.fill_(0 if zero else 0xFF)
        self.blocks.append((raw, n))
        return body.view(dtype).reshape(tuple(size))

    def _on_device(self, device):
        return device is not None and self._t.device(device).type == 'cuda'

    def empty(self, *size, dtype=None, device=None, **kw):
        if not self._on_device(device) or kw:
            return self._t.empty(*size, dtype=dtype, device=device, **kw)
        return self._guarded(size, dtype, device, zero=False)

    def zeros(self, *size, dtype=None, device=None, **kw):
        if not self._on_device(device) or kw:
            return self._t.zeros(*size, dtype=dtype, device=device, **kw)
        return self._guarded(size, dtype, device, zero=True)

    def check(self):
        assert self.blocks, 'nothing was allocated through the guarded module'
        for raw, n in self.blocks:
            head = raw[:GUARD]
            tail = raw[GUARD + n:]
            assert bool((head == CANARY).all()) and bool((tail == CANARY).all()), \
                'a kernel wrote outside a %d-byte buffer' % n
        return len(self.blocks)


@pytest.fixture(scope='module')
def md():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    import mdevaluate_b200
    return mdevaluate_b200


@pytest.fixture
def guarded(monkeypatch):
    import torch
    from mdevaluate_b200._cuda import lib as _lib
    _lib.require_cuda()
    proxy = GuardedTorch(torch)
    monkeypatch.setattr(_lib, 'require_cuda', lambda: proxy)
    yield proxy


def _close(got, ref, rtol=1e-5, atol=1e-7):
    ref = np.asarray(ref, dtype=np.float64)
    scale = max(1.0, float(np.nanmax(np.abs(ref)))) if ref.size else 1.0
    ok = np.isnan(ref) | (np.abs(np.asarray(got, dtype=np.float64) - ref) <= rtol * np.abs(ref) + atol * scale)
    assert ok.all()


@pytest.mark.parametrize('n_atoms,box,sigma', [(2601, 6.0, 0.05), (1027, 1.3, 0.12), (61, 1.0, 0.2)])
def test_streaming_kernels_stay_inside_their_buffers(guarded, md, n_atoms, box, sigma):
    """Scan, packing, every correlation kernel (packed / byte / 16-bit image counts, float32 and float64 chains, fused
    histogram, masks, several sub-ensembles per pass), the one-shot API with caller-owned outputs, jump matrices."""
    import oracle.md_oracle as orc
    from mdevaluate_b200._cuda import engine
    from mdevaluate_b200.synthetic import random_walk
    c = md.correlation
    T = 163
    pos, bx, times = random_walk(n_atoms, T, seed=7, sigma=sigma, box=box)
    coords = md.Coordinates.from_arrays(pos, bx, times)
    unw = orc.UnwrappedTrajectory(pos, bx, times)
    store = coords._resident().store_for(np.arange(T), nojump=True)
    assert np.array_equal(store.image_counts(), unw.counts)
    bins = np.linspace(0, 2.0, 101)
    kw = dict(segments=5, window=0.5)
    cases = ((c.msd, orc.msd, False), (partial(c.isf, q=22.7), partial(orc.isf, q=22.7), False),
             (c.non_gaussian, orc.non_gaussian, False),
             (partial(c.van_hove_self, bins=bins), partial(orc.van_hove_self, bins=bins), True),
             (partial(c.van_hove_self, bins=np.array([0, 0.1, 0.25, 0.6, 2.0])),
              partial(orc.van_hove_self, bins=np.array([0, 0.1, 0.25, 0.6, 2.0])), True))
    with np.errstate(all='ignore'):
        for tune in ({}, {'isf_float64': 1, 'hist_float64': 1}, {'use_packed': False}):
            for k, v in tune.items():
                setattr(engine.tuning, k, v)
            try:
                for ours, theirs, exact in cases:
                    _, ref = orc.shifted_correlation(theirs, unw, **kw)
                    _, r = c.shifted_correlation(ours, coords.nojump, **kw)
                    np.testing.assert_array_equal(r, ref) if exact else _close(r, ref)
                    # the one-shot entry point: sums, histogram and workspace are torch buffers (guarded)
                    spec = c._kernel_spec(ours)
                    starts, idx = c._schedule(coords, c.log_indices, 5, 0.5, None)
                    rows = store.rows_for(starts[:, None] + idx[None, :])
                    sums, hist = engine.correlate_device(store, rows, spec.corr)
                    one = spec.finalize(sums.cpu().numpy(), None if hist is None else hist.cpu().numpy(), n_atoms, False)
                    np.testing.assert_array_equal(one, ref) if exact else _close(one, ref)
                _, out = c.shifted_correlation_multi({'vh': partial(c.van_hove_self, bins=bins), 'a2': c.non_gaussian},
                                                     coords.nojump, **kw)
                np.testing.assert_array_equal(out['vh'], orc.shifted_correlation(partial(orc.van_hove_self, bins=bins),
                                                                                 unw, **kw)[1])
            finally:
                for k in tune:
                    setattr(engine.tuning, k, True if k == 'use_packed' else 0)
        # raw float32 frames
        raw = orc.ArrayTrajectory(pos, bx, times)
        _, r = c.shifted_correlation(c.msd, coords, **kw)
        _close(r, orc.shifted_correlation(orc.msd, raw, **kw)[1], rtol=2e-5)
    # sub-ensembles: a mask per origin, and four slabs in one pass
    half = lambda f: np.where(np.asarray(f)[:, 0] % box < 0.5 * box)[0]
    _, r = c.shifted_correlation(partial(c.isf, q=22.7), coords.nojump, correlation=c.subensemble_correlation(half), **kw)
    _, ref = orc.shifted_correlation(partial(orc.isf, q=22.7), unw, correlation=orc.subensemble_pairwise(half), **kw)
    _close(r, ref)
    edges = np.linspace(0, box, 5)
    slabs = lambda f: np.array([(np.asarray(f)[:, 2] % box >= lo) & (np.asarray(f)[:, 2] % box < hi)
                                for lo, hi in zip(edges[:-1], edges[1:])])
    _, r, n = c.shifted_correlation(partial(c.isf, q=22.7), coords.nojump,
                                    correlation=c.multi_subensemble_correlation(slabs), **kw)
    _, ref, n_ref = orc.shifted_correlation(partial(orc.isf, q=22.7), unw,
                                            correlation=orc.multi_subensemble_pairwise(slabs), **kw)
    np.testing.assert_array_equal(n, n_ref)
    _close(r, ref)
    # jump matrices off the device, unwrapped frames back
    md.reader.parse_jumps(coords)
    np.testing.assert_array_equal(np.asarray(coords.nojump[T - 1]), unw[T - 1])
    assert guarded.check() > 10


def test_dense_and_overlap_kernels_stay_inside_their_buffers(guarded, md, monkeypatch):
    """Dense schedules (float32-offset tiles and the float64 tile kernel) and the overlap kernels' cell lists."""
    import oracle.md_oracle as orc
    from mdevaluate_b200._cuda import engine
    from mdevaluate_b200.coordinates import CoordinatesKDTree
    from mdevaluate_b200.synthetic import random_walk
    c = md.correlation
    monkeypatch.setattr(c, 'DENSE_MIN_PAIRS', 1)
    n_atoms, T, box = 1091, 150, 2.0
    pos, bx, times = random_walk(n_atoms, T, seed=11, sigma=0.06, box=box)
    coords = md.Coordinates.from_arrays(pos, bx, times)
    unw = orc.UnwrappedTrajectory(pos, bx, times)
    dense = lambda a, b: np.arange(a, b)
    kw = dict(segments=41, window=0.45, average=True, index_distribution=dense)
    for tune in ({}, {'dense_float64': 1}):
        for k, v in tune.items():
            setattr(engine.tuning, k, v)
        try:
            for ours, theirs in ((c.msd, orc.msd), (partial(c.isf, q=22.7), partial(orc.isf, q=22.7))):
                _, r = c.shifted_correlation(ours, coords.nojump, **kw)
                _close(r, orc.shifted_correlation(theirs, unw, **kw)[1])
        finally:
            for k in tune:
                setattr(engine.tuning, k, 0)
    raw = orc.ArrayTrajectory(pos, bx.astype(np.float64), times)
    for boxsize in (None, box):
        tree = CoordinatesKDTree(coords, boxsize=boxsize)
        _, r = c.shifted_correlation(partial(c.overlap, crds_tree=tree, radius=0.09), coords, segments=4, window=0.4)
        _, ref = orc.shifted_correlation(partial(orc.overlap, crds_tree=orc.FrameTrees(raw, boxsize=boxsize), radius=0.09),
                                         raw, segments=4, window=0.4)
        np.testing.assert_array_equal(r, ref)
    assert guarded.check() > 5
