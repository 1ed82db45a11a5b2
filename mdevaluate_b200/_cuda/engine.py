"""Device-side plumbing: HBM stores of frame rows and the calls into the C ABI.

torch is used for device memory, streams and (in parallel.py) NCCL -- plumbing only.
Layout in HBM:

* ``coords``  float32 ``(n_rows, pitch)``: one frame per row, xyz interleaved, ``pitch`` a multiple
  of 4 floats so that every row is 16-byte aligned (float4 streaming);
* ``images``  uint8 ``(n_rows, img_pitch)``: cumulative periodic image counts per component stored
  excess-128 (``count + IMAGE_BIAS``), same element order as ``coords`` (only for nojump);
* ``packed``  int32 ``(n_rows, packed_pitch)``: the same counts when none leaves [-1, 1], one word per group of
  4 atoms (byte a = atom a, bits 0-1 / 2-3 / 4-5 = n_x+1, n_y+1, n_z+1): what the correlation kernels stream;
* ``images16`` int16 ``(n_rows, pitch)``: plain counts, the fallback beyond +-127;
* ``box``     float64 ``(n_rows, 3)``: box diagonal of the frame stored in each row.
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import lib as _lib


IMAGE_BIAS = 128


def _round_up(x, m):
    return (x + m - 1) // m * m


def row_pitch(n_atoms):
    """Row length in elements: 3N rounded up to 32 floats (128 B) so rows start on cache lines."""
    return _round_up(3 * n_atoms, 32)


class Tuning:
    """Explicit kernel switches (``mdb_tuning`` of the C ABI); 0 = the measured default.  Process-wide, read at every
    launch / plan creation.  ``engine.tuning.isf_float64 = 1`` e.g. sends the ISF through the float64 chain."""
    FIELDS = ('ring_stages', 'atom_blocks', 'isf_float64', 'hist_float64', 'scan_direct', 'dense_stage_cta',
              'dense_float64', 'variant')

    def __init__(self):
        for f in self.FIELDS:
            setattr(self, f, 0)
        self.use_packed = True          # host-side: hand the packed image rows to the kernels when a store has them

    def __setattr__(self, name, value):
        object.__setattr__(self, name, value)
        object.__setattr__(self, '_key', None)

    def fill(self, struct):
        for f in self.FIELDS:
            setattr(struct, f, int(getattr(self, f)))

    def key(self):
        key = self._key
        if key is None:
            key = tuple(int(getattr(self, f)) for f in self.FIELDS) + (bool(self.use_packed),)
            object.__setattr__(self, '_key', key)
        return key


tuning = Tuning()


def packed_pitch(n_atoms):
    """Row length of packed image rows in 32-bit words: one word per 4 atoms, rounded up to 128 bytes."""
    return _round_up((n_atoms + 3) // 4, 32)


class DeviceStore:
    """Frame rows resident in HBM plus the map trajectory frame -> row."""

    def __init__(self, coords, n_atoms, images=None, box=None, frame_to_row=None, n_frames=None, box_const=None,
                 packed=None, images16=None):
        self.coords = coords            # torch (n_rows, pitch) float32 / float64
        self.n_atoms = int(n_atoms)
        self.images = images            # torch uint8 (n_rows, img_pitch), excess-128, or None
        self.packed = packed            # torch int32 (n_rows, packed_pitch) 2-bit counts, or None
        self.images16 = images16        # torch int16 (n_rows, pitch) counts beyond +-127, or None
        self.box = box                  # torch float64 (n_rows, 3) or None
        self.box_const = box_const      # 3 floats when every row has the same box
        self.frame_to_row = frame_to_row  # np.int32 (T,) with -1 = absent, or None = identity
        self.n_frames = n_frames if n_frames is not None else coords.shape[0]
        self.source = None              # what the image counts were derived from (Resident: jump matrices / scan)
        self._cache = {}

    @property
    def device(self):
        return self.coords.device

    @property
    def has_images(self):
        return self.images is not None or self.packed is not None or self.images16 is not None

    def without_images(self):
        """The same rows read as raw (wrapped) coordinates: a view that shares all buffers."""
        if not self.has_images:
            return self
        view = self._cache.get('raw_view')
        if view is None:
            view = DeviceStore(self.coords, self.n_atoms, frame_to_row=self.frame_to_row, n_frames=self.n_frames)
            self._cache['raw_view'] = view
        return view

    def atom_block(self, a0, a1):
        """The columns of atoms [a0, a1) as a store that shares all buffers (a0 a multiple of 16 keeps every row
        16-byte aligned in all formats)."""
        if a0 % 16:
            raise ValueError('atom blocks must start at a multiple of 16 atoms')
        return DeviceStore(self.coords[:, 3 * a0:], a1 - a0,
                           images=None if self.images is None else self.images[:, 3 * a0:],
                           packed=None if self.packed is None else self.packed[:, a0 // 4:],
                           images16=None if self.images16 is None else self.images16[:, 3 * a0:],
                           box=self.box, frame_to_row=self.frame_to_row, n_frames=self.n_frames,
                           box_const=self.box_const)

    def image_counts(self, rows=None):
        """Host int64 ``(n_rows, n_atoms, 3)`` image counts of the given storage rows (all rows by default), from
        whichever format the store holds -- for tests and diagnostics."""
        n3 = 3 * self.n_atoms
        sel = slice(None) if rows is None else np.asarray(rows, dtype=np.int64)
        if self.images is not None:
            arr = self.images[:, :n3].cpu().numpy()[sel].astype(np.int64) - IMAGE_BIAS
        elif self.images16 is not None:
            arr = self.images16[:, :n3].cpu().numpy()[sel].astype(np.int64)
        elif self.packed is not None:
            words = self.packed[:, :(self.n_atoms + 3) // 4].cpu().numpy()[sel].view(np.uint32)
            c = np.arange(12)
            shifts = (8 * (c // 3) + 2 * (c % 3)).astype(np.uint32)
            arr = ((words[:, :, None] >> shifts[None, None, :]) & 3).astype(np.int64) - 1
            arr = arr.reshape(arr.shape[0], -1)[:, :n3]
        else:
            raise ValueError('the store holds no image counts')
        return arr.reshape(arr.shape[0], self.n_atoms, 3)

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


time_slot = -1            # >= 0: plans bracket their correlation kernel with the CUDA event pair of that slot
last_plan = None          # weak reference to the plan that ran last (its kernel timings: CorrPlan.kernel_ms)


def profile_kernels(enable):
    """Bracket the correlation kernel of every following plan run with CUDA events (slot 0)."""
    global time_slot
    time_slot = 0 if enable else -1


def last_kernel_ms():
    """Duration of the correlation kernel of the last plan run (after ``profile_kernels(True)``)."""
    plan = last_plan() if last_plan is not None else None
    if plan is None or time_slot < 0:
        raise _lib.CudaPathError('no profiled launch (call profile_kernels(True) first)')
    return plan.kernel_ms(time_slot)


def _fill_corr_args(a, store, rows_dev, S, L, spec, select, edges_dev, n_segments=1):
    """The part of mdb_corr_args that describes the rows and the request (outputs are set by the caller)."""
    torch = _lib.require_cuda()
    a.coords = store.coords.data_ptr()
    a.pitch = store.coords.stride(0)
    a.coord_type = _lib.F32 if store.coords.dtype == torch.float32 else _lib.F64
    if store.has_images:
        if store.images is not None:
            a.images = store.images.data_ptr()
            a.img_pitch = store.images.stride(0)
        if store.packed is not None and tuning.use_packed:
            a.packed = store.packed.data_ptr()
            a.packed_pitch = store.packed.stride(0)
        if store.images16 is not None:
            a.images16 = store.images16.data_ptr()
            a.img16_pitch = store.images16.stride(0)
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
    tuning.fill(a.tuning)
    a.n_segments = int(n_segments)


def _rows_on_device(store, rows):
    torch = _lib.require_cuda()
    key = ('rows', rows.shape, rows.tobytes())       # the bytes themselves: no hash-collision risk
    rows_dev = store._cache.get(key)
    if rows_dev is None:
        rows_dev = torch.from_numpy(rows).to(store.device)
        if len(store._cache) > 64:
            for k in [k for k in store._cache if isinstance(k, tuple) and k and k[0] in ('rows', 'plan')]:
                plan = store._cache.pop(k)
                if isinstance(plan, CorrPlan):
                    plan.close()
        store._cache[key] = rows_dev
    return rows_dev, key


def _edges_on_device(store, spec):
    torch = _lib.require_cuda()
    if not (spec.flags & _lib.HIST) or spec.bin_mode != _lib.BINS_EDGES:
        return None
    ekey = ('edges', spec.edges.tobytes())
    edges_dev = store._cache.get(ekey)
    if edges_dev is None:
        edges_dev = store._cache[ekey] = torch.from_numpy(spec.edges).to(store.device)
    return edges_dev


def correlate_device(store, rows, spec, stream=None, select=None):
    """Launch the correlation kernels; returns device tensors ``(sums (S,L,3) f64, hist (S,L,B) i64|None)``.

    ``rows`` is an int32 ``(S, L)`` array of storage rows (host NumPy, cached on the device per store).  One-shot
    form (fresh output tensors per call); repeated calls of the same request go through :class:`CorrPlan`.
    """
    torch = _lib.require_cuda()
    lib = _lib.load()
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    S, L = rows.shape
    dev = store.device
    rows_dev, _ = _rows_on_device(store, rows)
    want_hist = bool(spec.flags & _lib.HIST)
    sums = torch.empty((S, L, 3), dtype=torch.float64, device=dev)
    hist = torch.empty((S, L, spec.n_bins), dtype=torch.int64, device=dev) if want_hist else None
    a = _lib.CorrArgs()
    _fill_corr_args(a, store, rows_dev, S, L, spec, select, _edges_on_device(store, spec))
    a.sums = sums.data_ptr()
    a.hist = hist.data_ptr() if hist is not None else None
    nbytes = C.c_size_t(0)
    _lib.check(lib.mdb_corr_workspace_bytes(C.byref(a), C.byref(nbytes)), 'mdb_corr_workspace_bytes')
    need = nbytes.value
    ws = store._cache.get('workspace')
    if ws is None or ws.numel() < need:
        ws = torch.empty((max(need, 1),), dtype=torch.uint8, device=dev)
        store._cache['workspace'] = ws
    a.workspace = ws.data_ptr()
    a.workspace_bytes = ws.numel()
    st = stream if stream is not None else torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        _lib.check(lib.mdb_shifted_corr(C.byref(a), C.c_void_p(st.cuda_stream)), 'mdb_shifted_corr')
    _count(lib.mdb_corr_last_launches())
    return sums, hist


def clock_pair():
    """(gpu_ns, host_before_ns, host_seen_ns): one tiny kernel stores the GPU's clock into mapped host memory."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    g, a, b = C.c_ulonglong(0), C.c_ulonglong(0), C.c_ulonglong(0)
    _lib.check(lib.mdb_clock_pair(C.c_void_p(torch.cuda.current_stream().cuda_stream), C.byref(g), C.byref(a), C.byref(b)), 'mdb_clock_pair')
    return int(g.value), int(a.value), int(b.value)


class CorrPlan:
    """One repeated correlation request on one store (``mdb_corr_plan``): the C side owns workspace, device sums /
    histogram and a page-locked host block; :meth:`run` is two kernel launches and returns NumPy views of that host
    block (valid until the next run).  With ``comm`` (a :class:`Comm`) the sums come back added over the GPUs."""

    def __init__(self, store, rows, spec, select=None, comm=None, n_segments=1):
        torch = _lib.require_cuda()
        lib = _lib.load()
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        self.shape = rows.shape
        S, L = rows.shape
        # no reference to the store itself: cached plans live IN store._cache, and a cycle would keep gigabytes of
        # HBM waiting for the garbage collector; the tensors the C side points into are kept alive explicitly
        self.comm = comm
        self._keep = (_rows_on_device(store, rows)[0], _edges_on_device(store, spec), select, store.coords,
                      store.images, store.packed, store.images16, store.box)
        a = _lib.CorrArgs()
        _fill_corr_args(a, store, self._keep[0], S, L, spec, select, self._keep[1], n_segments)
        handle = C.c_void_p()
        with torch.cuda.device(store.device):
            _lib.check(lib.mdb_corr_plan_create(C.byref(a), comm.handle if comm is not None else None,
                                                C.byref(handle)), 'mdb_corr_plan_create')
        self.handle = handle
        self._lib = lib
        # the raw handle of torch's current stream on the store's device, without building a Stream object per call
        self._raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', None)
        self._stream_of = torch.cuda.current_stream
        self._device = store.device
        self._device_index = store.device.index if store.device.index is not None else torch.cuda.current_device()
        n_seg = max(1, int(n_segments))
        n = S * L * 3 * n_seg
        ptr = lib.mdb_corr_plan_sums(handle)
        self.sums = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(n,))
        self.sums = self.sums.reshape((S, L, 3) if n_seg == 1 else (S, L, n_seg, 3))
        self.hist = None
        if spec.flags & _lib.HIST:
            hp = lib.mdb_corr_plan_hist(handle)
            self.hist = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_int64)),
                                              shape=(S * L * spec.n_bins,)).reshape(S, L, spec.n_bins)
        self.launches = lib.mdb_corr_plan_launches(handle) + (2 if spec.flags & _lib.HIST else 0)
        self._weak = weakref.ref(self)

    def run(self, slot=None):
        """Launch on the current stream, wait, and return ``(sums, hist)`` -- views of the plan's host block."""
        global last_plan
        if self._raw_stream is not None:
            st = self._raw_stream(self._device_index)
        else:
            st = self._stream_of(self._device).cuda_stream
        rc = self._lib.mdb_corr_plan_run(self.handle, C.c_void_p(st), time_slot if slot is None else slot)
        if rc:
            _lib.check(rc, 'mdb_corr_plan_run')
        _count(self.launches)
        last_plan = self._weak
        return self.sums, self.hist

    def kernel_ms(self, time_slot):
        ms = C.c_float(0.0)
        _lib.check(self._lib.mdb_corr_plan_kernel_ms(self.handle, time_slot, C.byref(ms)), 'mdb_corr_plan_kernel_ms')
        return float(ms.value)

    def timeline(self):
        """Where the time of the last run went (ns): host clock at launch entry / launch return / result seen, and
        the GPU's clock at correlation-kernel start / finish-kernel start / finish-kernel end (different origins,
        see :func:`clock_pair`)."""
        out = (C.c_ulonglong * 6)()
        _lib.check(self._lib.mdb_corr_plan_timeline(self.handle, out), 'mdb_corr_plan_timeline')
        return [int(v) for v in out]

    def close(self):
        if self.handle is not None:
            self._lib.mdb_corr_plan_destroy(self.handle)
            self.handle = None
            self.sums = self.hist = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Comm:
    """Peer-memory communicator (``mdb_comm``): every rank's inbox is mapped into every other rank with CUDA IPC;
    the finish kernel of a plan adds the partial sums of all GPUs over NVLink.  ``exchange(handle_bytes) ->
    [bytes of rank 0, ...]`` carries the 64-byte IPC handles once (``torch.distributed.all_gather_object``)."""

    def __init__(self, rank, world, exchange, capacity_words=1 << 20, timeout_ms=20000):
        _lib.require_cuda()
        lib = _lib.load()
        self.rank, self.world = int(rank), int(world)
        self.capacity_words = int(capacity_words)
        handle = C.c_void_p()
        _lib.check(lib.mdb_comm_create(self.rank, self.world, self.capacity_words, int(timeout_ms), C.byref(handle)),
                   'mdb_comm_create')
        self.handle = handle
        self._lib = lib
        if self.world > 1:
            mine = C.create_string_buffer(_lib.COMM_HANDLE_BYTES)
            _lib.check(lib.mdb_comm_handle(handle, mine), 'mdb_comm_handle')
            everyone = exchange(bytes(mine.raw))
            if len(everyone) != self.world or any(len(h) != _lib.COMM_HANDLE_BYTES for h in everyone):
                raise ValueError('exchange() must return the %d-byte handle of every rank' % _lib.COMM_HANDLE_BYTES)
            blob = C.create_string_buffer(b''.join(everyone), self.world * _lib.COMM_HANDLE_BYTES)
            _lib.check(lib.mdb_comm_connect(handle, blob), 'mdb_comm_connect')

    def close(self):
        if self.handle is not None:
            self._lib.mdb_comm_destroy(self.handle)
            self.handle = None


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
    tuning.fill(a.tuning)
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
                out_row=None, coords_out=None, overflow=None, whole0=None, stream=None, images16=None):
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
    if images16 is not None:
        a.images16 = images16.data_ptr()
        a.img16_pitch = images16.stride(0)
    tuning.fill(a.tuning)
    dev = coords_seg.device
    st = stream if stream is not None else torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        _lib.check(lib.mdb_nojump_scan(C.byref(a), C.c_void_p(st.cuda_stream)), 'mdb_nojump_scan')
    _count(1)


def images_from_jumps(indptr, frames, jumps, n_atoms, images, keep=None, overflow=None, stream=None, images16=None):
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
    out = images if images is not None else images16
    a.n_keep = keep.numel() if keep is not None else out.shape[0]
    if images is not None:
        a.images = images.data_ptr()
        a.img_pitch = images.stride(0)
    else:
        a.images16 = images16.data_ptr()
        a.img16_pitch = images16.stride(0)
    a.overflow = overflow.data_ptr() if overflow is not None else None
    dev = out.device
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


def nojump_apply(coords_row, images_row, box_diag, n_atoms, image_format=_lib.IMAGE_BYTES):
    """float64 unwrapped frame ``x - n*box`` of one stored row (device tensor (3N,)); ``images_row`` in any of the
    three image formats."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = coords_row.device
    out = torch.empty((3 * n_atoms,), dtype=torch.float64, device=dev)
    box = (C.c_double * 3)(*[float(b) for b in box_diag])
    st = torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        _lib.check(lib.mdb_nojump_apply_format(C.c_void_p(coords_row.data_ptr()), C.c_void_p(images_row.data_ptr()),
                                               int(image_format), box, n_atoms, C.c_void_p(out.data_ptr()),
                                               C.c_void_p(st.cuda_stream)), 'mdb_nojump_apply_format')
    _count(1)
    return out


def pack_images(images, n_atoms):
    """Byte image rows -> packed rows (``mdb_images_pack``).  Returns ``(packed int32 (n_rows, packed_pitch),
    max |count|)``; the packed rows are valid only if the maximum is <= 1."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = images.device
    packed = torch.zeros((images.shape[0], packed_pitch(n_atoms)), dtype=torch.int32, device=dev)
    worst = torch.zeros((1,), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        _lib.check(lib.mdb_images_pack(C.c_void_p(images.data_ptr()), images.stride(0), images.shape[0], n_atoms,
                                       C.c_void_p(packed.data_ptr()), packed.stride(0), C.c_void_p(worst.data_ptr()),
                                       C.c_void_p(st.cuda_stream)), 'mdb_images_pack')
    _count(1)
    return packed, int(worst.item())


def vector_rows(store, rows, atoms_a, atoms_b, normed, pbc_box=None, stream=None):
    """Rows of vectors between atoms (``mdb_vectors``) for the storage rows ``rows`` of ``store``: a new
    :class:`DeviceStore` of ``len(rows)`` rows x ``n_vec`` "atoms" (float64 when the store carries image counts --
    nojump frames are float64 -- else float32), row i = vectors of ``rows[i]``.  ``atoms_a`` / ``atoms_b``: int arrays
    ``(n_vec,)`` or ``(m, n_vec)`` (mean over m)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = store.device
    if store.coords.dtype != torch.float32:
        raise _lib.CudaPathError('vector rows are built from float32 frame rows')
    idx = []
    for atoms in (atoms_a, atoms_b):
        arr = np.ascontiguousarray(np.atleast_2d(np.asarray(atoms, dtype=np.int64)), dtype=np.int32)
        if arr.ndim != 2 or arr.min() < 0 or arr.max() >= store.n_atoms:
            raise IndexError('atom indices out of range')
        idx.append(arr)
    if idx[0].shape[1] != idx[1].shape[1]:
        raise ValueError('atoms_a and atoms_b must select the same number of atoms')
    n_vec = idx[0].shape[1]
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    f64 = store.has_images
    out = torch.zeros((len(rows), row_pitch(n_vec)), dtype=torch.float64 if f64 else torch.float32, device=dev)
    keep = [torch.from_numpy(a).to(dev) for a in idx] + [torch.from_numpy(rows).to(dev)]
    a = _lib.VectorsArgs()
    a.coords = store.coords.data_ptr()
    a.pitch = store.coords.stride(0)
    if f64:
        if store.images is not None:
            img, a.image_format = store.images, _lib.IMAGE_BYTES
        elif store.images16 is not None:
            img, a.image_format = store.images16, _lib.IMAGE_I16
        else:
            img, a.image_format = store.packed, _lib.IMAGE_PACKED
        a.images = img.data_ptr()
        a.img_pitch = img.stride(0)
        if store.box_const is not None:
            a.box_const[0], a.box_const[1], a.box_const[2] = store.box_const
        else:
            a.box = store.box.data_ptr()
    a.rows = keep[2].data_ptr()
    a.n_rows = len(rows)
    a.atoms_a, a.m_a = keep[0].data_ptr(), idx[0].shape[0]
    a.atoms_b, a.m_b = keep[1].data_ptr(), idx[1].shape[0]
    a.n_vec = n_vec
    a.normed = 1 if normed else 0
    if pbc_box is not None:
        a.use_pbc = 1
        a.pbc_box[0], a.pbc_box[1], a.pbc_box[2] = (float(b) for b in pbc_box)
    a.out_type = _lib.F64 if f64 else _lib.F32
    a.out = out.data_ptr()
    a.out_pitch = out.stride(0)
    st = stream if stream is not None else torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        _lib.check(lib.mdb_vectors(C.byref(a), C.c_void_p(st.cuda_stream)), 'mdb_vectors')
    _count(1)
    return DeviceStore(out, n_vec)


def _fill_frame_rows(a, prefix, store):
    """The fields of mdb_overlap_args that describe one trajectory's rows (``prefix`` '' or 'ref_')."""
    torch = _lib.require_cuda()
    setattr(a, prefix + 'coords', store.coords.data_ptr())
    setattr(a, prefix + 'pitch', store.coords.stride(0))
    setattr(a, 'ref_type' if prefix else 'coord_type', _lib.F32 if store.coords.dtype == torch.float32 else _lib.F64)
    if not store.has_images:
        return
    if store.packed is not None and tuning.use_packed:
        images, fmt = store.packed, _lib.IMAGE_PACKED
    elif store.images is not None:
        images, fmt = store.images, _lib.IMAGE_BYTES
    else:
        images, fmt = store.images16, _lib.IMAGE_I16
    setattr(a, prefix + 'images', images.data_ptr())
    setattr(a, prefix + 'image_format', fmt)
    setattr(a, prefix + 'img_pitch', images.stride(0))
    box_const = getattr(a, prefix + 'box_const')
    if store.box_const is not None:
        box_const[0], box_const[1], box_const[2] = store.box_const
    else:
        setattr(a, prefix + 'box', store.box.data_ptr())


overlap_file_order = False      # measurement switch: query atoms in file order instead of the tree's cell order


def overlap_counts(tree_store, tree_rows, store, rows, radius, boxsize=None, tree_select=None, stream=None):
    """``mdb_overlap``: for every (origin, lag) the number of atoms of frame ``rows[s, l]`` of ``store`` that have an
    atom of the tree of origin s -- row ``tree_rows[s]`` of ``tree_store``, restricted to ``tree_select[s]`` (uint8
    device tensor (S, pitch)) -- within ``radius`` (cKDTree arithmetic, periodic with ``boxsize``).
    Returns ``(counts (S, L) int64, tree sizes (S,), bounds (S, 6), out_of_box flag)`` as NumPy arrays."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = store.device
    if tree_store.device != dev:
        raise _lib.CudaPathError('the tree and the frames must live on the same device')
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    tree_rows = np.ascontiguousarray(tree_rows, dtype=np.int32).ravel()
    S, L = rows.shape
    if tree_rows.shape[0] != S:
        raise ValueError('one tree row per origin')
    a = _lib.OverlapArgs()
    _fill_frame_rows(a, 'ref_', tree_store)
    _fill_frame_rows(a, '', store)
    a.n_ref = tree_store.n_atoms
    a.n_atoms = store.n_atoms
    rows_dev = torch.from_numpy(rows).to(dev)
    tree_rows_dev = torch.from_numpy(tree_rows).to(dev)
    a.rows = rows_dev.data_ptr()
    a.ref_rows = tree_rows_dev.data_ptr()
    if tree_select is not None:
        if tree_select.shape[0] != S or tree_select.stride(0) < tree_store.n_atoms or tree_select.dtype != torch.uint8:
            raise ValueError('tree_select must be (S, pitch) uint8 with pitch >= the atoms of a tree row')
        a.ref_select = tree_select.data_ptr()
        a.ref_sel_pitch = tree_select.stride(0)
    a.n_origins, a.n_lags = S, L
    a.file_order = 1 if overlap_file_order else 0
    a.radius = float(radius)
    if boxsize is not None:
        b = np.broadcast_to(np.asarray(boxsize, dtype=np.float64), (3,))
        a.boxsize[0], a.boxsize[1], a.boxsize[2] = [float(v) for v in b]
    counts = torch.empty((S, L), dtype=torch.int64, device=dev)
    sizes = torch.empty((S,), dtype=torch.int32, device=dev)
    bounds = torch.empty((S, 6), dtype=torch.float64, device=dev)
    bad = torch.zeros((1,), dtype=torch.int32, device=dev)
    a.counts, a.tree_sizes, a.bounds, a.out_of_box = counts.data_ptr(), sizes.data_ptr(), bounds.data_ptr(), bad.data_ptr()
    need = C.c_size_t(0)
    _lib.check(lib.mdb_overlap_workspace_bytes(C.byref(a), C.byref(need)), 'mdb_overlap_workspace_bytes')
    ws = torch.empty((max(int(need.value), 1),), dtype=torch.uint8, device=dev)
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    st = stream if stream is not None else torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        _lib.check(lib.mdb_overlap(C.byref(a), C.c_void_p(st.cuda_stream)), 'mdb_overlap')
    _count(lib.mdb_overlap_launches())
    return counts.cpu().numpy(), sizes.cpu().numpy(), bounds.cpu().numpy(), bool(bad.item())

