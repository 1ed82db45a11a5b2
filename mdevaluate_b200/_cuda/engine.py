"""Device-side plumbing: HBM stores of frame rows and the calls into the C ABI.

torch is used for device memory, streams and (in parallel.py) NCCL -- plumbing only.
Layout in HBM:

* ``coords``  float32 ``(n_rows, pitch)``: one frame per row, xyz interleaved, ``pitch`` a multiple
  of 4 floats so that every row is 16-byte aligned (float4 streaming);
* ``images``  uint8 ``(n_rows, img_pitch)``: cumulative periodic image counts per component stored
  excess-128 (``count + IMAGE_BIAS``), same element order as ``coords`` (only for nojump);
* ``box``     float64 ``(n_rows, 3)``: box diagonal of the frame stored in each row.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib as _lib


IMAGE_BIAS = 128


def _round_up(x, m):
    return (x + m - 1) // m * m


def row_pitch(n_atoms):
    """Row length in elements: 3N rounded up to 32 floats (128 B) so rows start on cache lines."""
    return _round_up(3 * n_atoms, 32)


class DeviceStore:
    """Frame rows resident in HBM plus the map trajectory frame -> row."""

    def __init__(self, coords, n_atoms, images=None, box=None, frame_to_row=None, n_frames=None, box_const=None):
        self.coords = coords            # torch (n_rows, pitch) float32 / float64
        self.n_atoms = int(n_atoms)
        self.images = images            # torch uint8 (n_rows, img_pitch), excess-128, or None
        self.box = box                  # torch float64 (n_rows, 3) or None
        self.box_const = box_const      # 3 floats when every row has the same box
        self.frame_to_row = frame_to_row  # np.int32 (T,) with -1 = absent, or None = identity
        self.n_frames = n_frames if n_frames is not None else coords.shape[0]
        self._cache = {}

    @property
    def device(self):
        return self.coords.device

    def without_images(self):
        """The same rows read as raw (wrapped) coordinates: a view that shares all buffers."""
        if self.images is None:
            return self
        view = self._cache.get('raw_view')
        if view is None:
            view = DeviceStore(self.coords, self.n_atoms, frame_to_row=self.frame_to_row, n_frames=self.n_frames)
            self._cache['raw_view'] = view
        return view

    def rows_for(self, frame_ids):
        frame_ids = np.asarray(frame_ids)
        if self.frame_to_row is None:
            rows = frame_ids.astype(np.int32)
        else:
            rows = self.frame_to_row[frame_ids]
            if (rows < 0).any():
                raise KeyError('frame not resident in this store')
        return np.ascontiguousarray(rows, dtype=np.int32)

    def has_frames(self, frame_ids):
        if self.frame_to_row is None:
            return True
        return bool((self.frame_to_row[np.asarray(frame_ids)] >= 0).all())


class CorrSpec:
    """What one correlation launch has to accumulate (built by correlation._kernel_spec)."""

    def __init__(self, flags, q=0.0, n_bins=0, bin_mode=_lib.BINS_UNIFORM, dx=0.0, edges=None, order=0):
        self.flags = int(flags)
        self.q = float(q)
        self.order = int(order)
        self.n_bins = int(n_bins)
        self.bin_mode = int(bin_mode)
        self.dx = float(dx)
        self.edges = None if edges is None else np.ascontiguousarray(edges, dtype=np.float64)


_launch_counter = 0


def launches():
    """Total number of kernels this process enqueued through the C ABI (bench.py reports it)."""
    return _launch_counter


def _count(n):
    global _launch_counter
    _launch_counter += n


def profile_kernels(enable):
    """Bracket the correlation kernel of every following launch with CUDA events (bench.py)."""
    _lib.check(_lib.load().mdb_corr_profile(1 if enable else 0), 'mdb_corr_profile')


def last_kernel_ms():
    ms = C.c_float(0.0)
    _lib.check(_lib.load().mdb_corr_last_kernel_ms(C.byref(ms)), 'mdb_corr_last_kernel_ms')
    return float(ms.value)


def correlate_device(store, rows, spec, stream=None, select=None):
    """Launch the correlation kernels; returns device tensors ``(sums (S,L,3) f64, hist (S,L,B) i64|None)``.

    ``rows`` is an int32 ``(S, L)`` array of storage rows (host NumPy, cached on the device per store).
    """
    torch = _lib.require_cuda()
    lib = _lib.load()
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    S, L = rows.shape
    dev = store.device
    key = ('rows', rows.shape, rows.tobytes())       # the bytes themselves: no hash-collision risk
    rows_dev = store._cache.get(key)
    if rows_dev is None:
        rows_dev = torch.from_numpy(rows).to(dev)
        if len(store._cache) > 64:
            store._cache.clear()
        store._cache[key] = rows_dev
    want_hist = bool(spec.flags & _lib.HIST)
    sums = torch.empty((S, L, 3), dtype=torch.float64, device=dev)
    hist = torch.empty((S, L, spec.n_bins), dtype=torch.int64, device=dev) if want_hist else None
    edges_dev = None
    if want_hist and spec.bin_mode == _lib.BINS_EDGES:
        ekey = ('edges', spec.edges.tobytes())
        edges_dev = store._cache.get(ekey)
        if edges_dev is None:
            edges_dev = torch.from_numpy(spec.edges).to(dev)
            store._cache[ekey] = edges_dev

    a = _lib.CorrArgs()
    a.coords = store.coords.data_ptr()
    a.pitch = store.coords.stride(0)
    a.coord_type = _lib.F32 if store.coords.dtype == torch.float32 else _lib.F64
    if store.images is not None:
        a.images = store.images.data_ptr()
        a.img_pitch = store.images.stride(0)
        if store.box_const is not None:
            a.box = None
            a.box_const[0], a.box_const[1], a.box_const[2] = store.box_const
        else:
            a.box = store.box.data_ptr()
    a.n_atoms = store.n_atoms
    a.n_origins = S
    a.n_lags = L
    a.rows = rows_dev.data_ptr()
    if select is not None:       # uint8 (S, sel_pitch): nonzero = atom takes part for that origin
        if select.shape[0] != S or select.stride(0) % 4 or select.stride(0) < store.n_atoms:
            raise ValueError('select must be (S, pitch) uint8 with pitch >= n_atoms and a multiple of 4')
        a.select = select.data_ptr()
        a.sel_pitch = select.stride(0)
    a.flags = spec.flags
    a.q = spec.q
    a.legendre_order = spec.order
    a.n_bins = spec.n_bins
    a.bin_mode = spec.bin_mode
    a.dx = spec.dx
    a.edges = edges_dev.data_ptr() if edges_dev is not None else None
    a.sums = sums.data_ptr()
    a.hist = hist.data_ptr() if hist is not None else None
    skey = ('ws_bytes', S, L, spec.flags, spec.n_bins, a.coord_type, store.images is not None)
    need = store._cache.get(skey)
    if need is None:
        nbytes = C.c_size_t(0)
        _lib.check(lib.mdb_corr_workspace_bytes(C.byref(a), C.byref(nbytes)), 'mdb_corr_workspace_bytes')
        need = store._cache[skey] = nbytes.value
    wkey = 'workspace'
    ws = store._cache.get(wkey)
    if ws is None or ws.numel() < need:
        ws = torch.empty((max(need, 1),), dtype=torch.uint8, device=dev)
        store._cache[wkey] = ws
    a.workspace = ws.data_ptr()
    a.workspace_bytes = ws.numel()
    st = stream if stream is not None else torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        _lib.check(lib.mdb_shifted_corr(C.byref(a), C.c_void_p(st.cuda_stream)), 'mdb_shifted_corr')
    _count(lib.mdb_corr_last_launches())
    return sums, hist


DENSE_MAX_ORIGIN_STEP = 4        # beyond that neighbouring origins share too few rows to pay off


def dense_device(store, first_origin, origin_step, n_origins, first_lag, n_lags, flags, q=0.0, stream=None):
    """Dense schedule (arithmetic origins x contiguous lags), summed over origins: device ``(L, 2)`` float64
    with the sums of r^2 and sinc(q r) over all origins and atoms (mdb_dense_corr)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = store.device
    if store.coords.dtype != torch.float32:
        raise _lib.CudaPathError('the dense kernel takes float32 rows')
    sums = torch.empty((n_lags, 2), dtype=torch.float64, device=dev)
    a = _lib.DenseArgs()
    a.coords = store.coords.data_ptr()
    a.pitch = store.coords.stride(0)
    if store.images is not None:
        a.images = store.images.data_ptr()
        a.img_pitch = store.images.stride(0)
        if store.box_const is not None:
            a.box_const[0], a.box_const[1], a.box_const[2] = store.box_const
        else:
            a.box = store.box.data_ptr()
    if store.frame_to_row is not None:
        rows_dev = store._cache.get('frame_rows')
        if rows_dev is None:
            rows_dev = store._cache['frame_rows'] = torch.from_numpy(
                np.ascontiguousarray(store.frame_to_row, dtype=np.int32)).to(dev)
        a.frame_rows = rows_dev.data_ptr()
    a.n_atoms = store.n_atoms
    a.first_origin, a.origin_step, a.n_origins = int(first_origin), int(origin_step), int(n_origins)
    a.first_lag, a.n_lags = int(first_lag), int(n_lags)
    a.flags = int(flags)
    a.q = float(q)
    a.sums = sums.data_ptr()
    nbytes = C.c_size_t(0)
    _lib.check(lib.mdb_dense_workspace_bytes(C.byref(a), C.byref(nbytes)), 'mdb_dense_workspace_bytes')
    ws = store._cache.get('workspace')
    if ws is None or ws.numel() < nbytes.value:
        ws = store._cache['workspace'] = torch.empty((max(nbytes.value, 1),), dtype=torch.uint8, device=dev)
    a.workspace = ws.data_ptr()
    a.workspace_bytes = ws.numel()
    st = stream if stream is not None else torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        _lib.check(lib.mdb_dense_corr(C.byref(a), C.c_void_p(st.cuda_stream)), 'mdb_dense_corr')
    _count(2)
    return sums


def nojump_scan(coords_seg, n_atoms, box0, box_is_f32, prev, cum, first_segment, images=None, jumps=None,
                out_row=None, coords_out=None, overflow=None, whole0=None, stream=None):
    """One segment of the jump scan (mdb_nojump_scan).  All tensors are device tensors."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    a = _lib.NojumpArgs()
    a.coords = coords_seg.data_ptr()
    a.pitch = coords_seg.stride(0)
    a.n_frames = coords_seg.shape[0]
    a.n_atoms = n_atoms
    a.box0[0], a.box0[1], a.box0[2] = (float(b) for b in box0)
    a.box_is_f32 = 1 if box_is_f32 else 0
    a.first_segment = 1 if first_segment else 0
    a.whole0 = whole0.data_ptr() if whole0 is not None else None
    a.prev = prev.data_ptr()
    a.cum = cum.data_ptr()
    a.out_row = out_row.data_ptr() if out_row is not None else None
    out = images if images is not None else jumps
    if images is not None and jumps is not None and images.stride(0) != jumps.stride(0):
        raise ValueError('images and jumps must share a row pitch')
    a.images = images.data_ptr() if images is not None else None
    a.jumps = jumps.data_ptr() if jumps is not None else None
    a.img_pitch = out.stride(0) if out is not None else 0
    a.coords_out = coords_out.data_ptr() if coords_out is not None else None
    a.pitch_out = coords_out.stride(0) if coords_out is not None else 0
    a.overflow = overflow.data_ptr() if overflow is not None else None
    dev = coords_seg.device
    st = stream if stream is not None else torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        _lib.check(lib.mdb_nojump_scan(C.byref(a), C.c_void_p(st.cuda_stream)), 'mdb_nojump_scan')
    _count(1)


def images_from_jumps(indptr, frames, jumps, n_atoms, images, keep=None, overflow=None, stream=None):
    """Image counts of the kept frames from jump matrices re-sorted by component column
    (mdb_nojump_images_from_jumps).  All tensors are device tensors; ``keep`` int32 ascending or None."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    if indptr.dtype != torch.int64 or indptr.numel() != 3 * n_atoms + 1:
        raise ValueError('indptr must be int64 with 3*n_atoms + 1 entries')
    if frames.dtype != torch.int32 or jumps.dtype != torch.int8 or frames.numel() != jumps.numel():
        raise ValueError('frames must be int32 and jumps int8, one entry per stored jump')
    a = _lib.JumpMatArgs()
    a.indptr = indptr.data_ptr()
    a.nnz = frames.numel()
    a.frames = frames.data_ptr() if a.nnz else None
    a.jumps = jumps.data_ptr() if a.nnz else None
    a.n_atoms = n_atoms
    a.keep = keep.data_ptr() if keep is not None else None
    a.n_keep = keep.numel() if keep is not None else images.shape[0]
    a.images = images.data_ptr()
    a.img_pitch = images.stride(0)
    a.overflow = overflow.data_ptr() if overflow is not None else None
    dev = images.device
    st = stream if stream is not None else torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        _lib.check(lib.mdb_nojump_images_from_jumps(C.byref(a), C.c_void_p(st.cuda_stream)),
                   'mdb_nojump_images_from_jumps')
    _count(1)


def copy_rows_h2d(src_ptr, src_pitch_bytes, dst, width_bytes, n_rows, stream):
    """Strided DMA of ``n_rows`` rows from page-locked host memory into the rows of device tensor ``dst``
    (mdb_copy_rows_h2d) on ``stream``."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    with torch.cuda.device(dst.device):
        _lib.check(lib.mdb_copy_rows_h2d(C.c_void_p(src_ptr), src_pitch_bytes, C.c_void_p(dst.data_ptr()),
                                         dst.stride(0) * dst.element_size(), width_bytes, n_rows,
                                         C.c_void_p(stream.cuda_stream)), 'mdb_copy_rows_h2d')


def nojump_apply(coords_row, images_row, box_diag, n_atoms):
    """float64 unwrapped frame ``x - n*box`` of one stored row (device tensor (3N,))."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = coords_row.device
    out = torch.empty((3 * n_atoms,), dtype=torch.float64, device=dev)
    box = (C.c_double * 3)(*[float(b) for b in box_diag])
    st = torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        _lib.check(lib.mdb_nojump_apply(C.c_void_p(coords_row.data_ptr()), C.c_void_p(images_row.data_ptr()),
                                        box, n_atoms, C.c_void_p(out.data_ptr()), C.c_void_p(st.cuda_stream)),
                   'mdb_nojump_apply')
    _count(1)
    return out
