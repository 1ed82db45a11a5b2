"""Bringing a trajectory into HBM: dense residency, compaction of needed frames, nojump scan.

One ``Resident`` exists per (reader, atom selection).  It answers "give me a DeviceStore that
holds these frames (with image counts if nojump)" with the cheapest of:

* dense   -- every frame of the trajectory is (or is made) resident; image counts for every
             frame come from ONE scan over the trajectory.  Chosen when it fits the memory budget
             (always for a trajectory that already lives in HBM).
* compact -- only the frames a schedule touches are kept.  The trajectory is streamed through
             the GPU in time segments (pinned host memory -> HBM, copy overlapped with the scan of
             the previous segment); the scan carries (prev, cum) across segments and writes the
             kept frames and their image counts straight into the compact store.
"""
from __future__ import annotations

import weakref

import numpy as np

from . import engine
from . import lib as _lib
from ..logging import logger


def _mem_budget(torch, device):
    """Bytes a new store may take: 70 % of what the driver reports free plus what torch's allocator holds
    cached but unused (a store that was just released is still 'reserved' for the driver)."""
    free, _total = torch.cuda.mem_get_info(device)
    cached = torch.cuda.memory_reserved(device) - torch.cuda.memory_allocated(device)
    return int((free + max(0, cached)) * 0.70)


class Resident:
    SEGMENT_BYTES = 512 << 20        # staging buffer per stream slot for host -> HBM streaming

    def __init__(self, reader, atom_filter, device=None):
        torch = _lib.require_cuda()
        # Coordinates caches this object ON the reader (one Resident per reader and selection), so a strong
        # reference back would form a cycle and the HBM buffers (GBs) would wait for the garbage collector
        # instead of being released with the last Coordinates: measured as 70 ms of cudaMalloc per
        # end-to-end step.  The reader outlives us through Coordinates.frames.
        self._reader_ref = None
        self._reader_obj = reader
        if hasattr(reader, '__dict__'):
            try:
                self._reader_ref = weakref.ref(reader)
                self._reader_obj = None
            except TypeError:
                pass
        self.n_frames = len(reader)
        first = reader[0].positions
        m_all = len(first)
        self.atom_filter = None
        if atom_filter is not None:
            atom_filter = np.asarray(atom_filter)
            if atom_filter.dtype == bool:
                if not atom_filter.all():
                    self.atom_filter = np.nonzero(atom_filter)[0]
            else:
                self.atom_filter = atom_filter.astype(np.int64)
        # a selection that is one contiguous run of atoms is copied as a slice, not gathered
        self._filter_run = None
        if self.atom_filter is not None and len(self.atom_filter) and np.all(np.diff(self.atom_filter) == 1):
            self._filter_run = (int(self.atom_filter[0]), int(self.atom_filter[-1]) + 1)
        self.n_atoms_all = m_all
        self._jump_lists = None         # (matrices, indptr, frames, jumps) on the device
        self.n_atoms = m_all if self.atom_filter is None else len(self.atom_filter)
        self.dtype = np.asarray(first).dtype
        block = getattr(reader, 'device_block', None)
        if device is None:
            device = block.device if block is not None else torch.device('cuda', torch.cuda.current_device())
        self.device = torch.device(device)
        self.pitch = engine.row_pitch(self.n_atoms)
        self._dense = None              # DeviceStore with all frames (raw)
        self._dense_images = False
        self._compact = {}              # (mode, hash of frame ids) -> DeviceStore
        self.force_compact = False
        self.pack_images = True         # re-code byte image rows as packed rows when no count leaves [-1, 1]
        self.lookup = {}                # correlation._provider: schedule -> (store, rows) of repeated calls

    @property
    def reader(self):
        if self._reader_ref is None:
            return self._reader_obj
        reader = self._reader_ref()
        if reader is None:
            raise ReferenceError('the trajectory reader of this store no longer exists')
        return reader

    # -- box / time helpers --------------------------------------------------------------------
    def _box_of(self, i):
        if hasattr(self.reader, 'box_of'):
            return np.asarray(self.reader.box_of(i))
        return np.asarray(self.reader[i].box).diagonal()

    def _box_rows(self, frame_ids):
        torch = _lib.require_cuda()
        if getattr(self.reader, 'box_constant', False):
            b = np.tile(np.asarray(self.reader.box_diag, dtype=np.float64), (len(frame_ids), 1))
        else:
            b = np.stack([np.asarray(self._box_of(int(i)), dtype=np.float64) for i in frame_ids])
        return torch.from_numpy(np.ascontiguousarray(b)).to(self.device)

    def _box_const(self):
        if getattr(self.reader, 'box_constant', False):
            return tuple(float(b) for b in np.asarray(self.reader.box_diag, dtype=np.float64))
        return None

    # -- segment source: a device float32 (n, pitch') view of frames [t0, t1) -------------------
    def _segment_frames(self):
        per_frame = self.pitch * 4
        return max(1, min(self.n_frames, self.SEGMENT_BYTES // per_frame))

    def _host_segment(self, t0, t1, out):
        """Fill pinned ``out[:t1-t0, :3N]`` with frames [t0, t1) of the selected atoms."""
        n3 = 3 * self.n_atoms
        block = getattr(self.reader, 'host_block', None)
        view = out[:t1 - t0, :n3].numpy().reshape(t1 - t0, self.n_atoms, 3)
        if block is not None:
            if self.atom_filter is None:
                np.copyto(view, block[t0:t1], casting='same_kind')
            elif self._filter_run is not None:
                np.copyto(view, block[t0:t1, self._filter_run[0]:self._filter_run[1]], casting='same_kind')
            else:
                np.take(block[t0:t1], self.atom_filter, axis=1, out=view)
        else:
            for t in range(t0, t1):
                pos = np.asarray(self.reader[t].positions)
                view[t - t0] = pos if self.atom_filter is None else pos[self.atom_filter]

    def _segments(self):
        """Yield ``(t0, t1, device_tensor)`` over the whole trajectory, double-buffered."""
        torch = _lib.require_cuda()
        generated = getattr(self.reader, 'device_segments', None)
        if generated is not None:
            # a reader that produces its frames on the device in time order (synthetic.RandomWalkStream)
            if self.atom_filter is not None:
                raise _lib.CudaPathError('a generated stream is read with all of its atoms')
            yield from generated(self.device)
            return
        block = getattr(self.reader, 'device_block', None)
        if block is not None and block.dtype == torch.float32 and (self.atom_filter is None or self._filter_run):
            # already in HBM: views of the block (all atoms, or the columns of one contiguous run of atoms)
            c0, c1 = (0, block.shape[1]) if self.atom_filter is None else (3 * self._filter_run[0],
                                                                            3 * self._filter_run[1])
            seg = max(1, min(self.n_frames, (2 << 30) // (block.stride(0) * 4)))
            for t0 in range(0, self.n_frames, seg):
                t1 = min(self.n_frames, t0 + seg)
                yield t0, t1, block[t0:t1, c0:c1]
            return
        if self.dtype != np.float32:
            raise _lib.CudaPathError('nojump needs float32 wrapped positions (got %s)' % self.dtype)
        seg = self._segment_frames()
        main = torch.cuda.current_stream(self.device)
        copy_stream = torch.cuda.Stream(self.device)
        n3 = 3 * self.n_atoms
        direct = self._pinned_block()
        pinned = ([] if direct is not None else
                  [torch.empty((seg, self.pitch), dtype=torch.float32).pin_memory() for _ in range(2)])
        staged = [torch.empty((seg, self.pitch), dtype=torch.float32, device=self.device) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]
        starts = list(range(0, self.n_frames, seg))

        def issue(k):
            slot = k % 2
            t0 = starts[k]
            t1 = min(self.n_frames, t0 + seg)
            if k >= 2:
                consumed[slot].synchronize()      # host buffer and device slot are free again
            with torch.cuda.stream(copy_stream):
                if direct is not None:
                    # the trajectory itself is page-locked: DMA straight out of it
                    staged[slot][:t1 - t0, :n3].copy_(direct[t0:t1], non_blocking=True)
                else:
                    self._host_segment(t0, t1, pinned[slot])
                    staged[slot][:t1 - t0].copy_(pinned[slot][:t1 - t0], non_blocking=True)
                ready[slot].record(copy_stream)

        if starts:
            issue(0)
        for k, t0 in enumerate(starts):
            slot = k % 2
            t1 = min(self.n_frames, t0 + seg)
            if k + 1 < len(starts):
                issue(k + 1)
            main.wait_event(ready[slot])
            yield t0, t1, staged[slot][:t1 - t0]
            consumed[slot].record(main)
        main.synchronize()

    def _pinned_block(self):
        """The whole host trajectory as a page-locked ``(T, 3N)`` tensor if it can be DMA'd as it is."""
        torch = _lib.require_cuda()
        host_block = getattr(self.reader, 'host_block', None)
        if host_block is None or self.atom_filter is not None or not host_block.flags.c_contiguous:
            return None
        if host_block.dtype != np.float32:
            return None
        as_t = torch.from_numpy(host_block).view(self.n_frames, 3 * self.n_atoms)
        return as_t if as_t.is_pinned() else None

    def _pinned_run(self):
        """``(address of the run's first value in frame 0, bytes per frame)`` when the selection is one
        contiguous run of atoms of a page-locked C-contiguous float32 host trajectory, else None."""
        torch = _lib.require_cuda()
        host_block = getattr(self.reader, 'host_block', None)
        if (host_block is None or self._filter_run is None or host_block.dtype != np.float32
                or not host_block.flags.c_contiguous):
            return None
        if not torch.from_numpy(host_block).is_pinned():
            return None
        row_bytes = host_block.shape[1] * 12
        return host_block.ctypes.data + self._filter_run[0] * 12, row_bytes

    def _upload_rows(self, coords):
        """Host trajectory -> the rows of ``coords`` (float32), entirely on a copy stream: nothing is queued
        on the compute stream, so the upload of one atom slab runs under the kernel of the previous one.
        The caller guarantees that no kernel still reads memory ``coords`` may have been carved from."""
        torch = _lib.require_cuda()
        main = torch.cuda.current_stream(self.device)
        copy_stream = torch.cuda.Stream(self.device)
        n3 = 3 * self.n_atoms
        seg = self._segment_frames()
        starts = range(0, self.n_frames, seg)
        direct = self._pinned_block()
        run = self._pinned_run()
        if direct is not None:
            with torch.cuda.stream(copy_stream):
                for t0 in starts:
                    t1 = min(self.n_frames, t0 + seg)
                    coords[t0:t1, :n3].copy_(direct[t0:t1], non_blocking=True)
        elif run is not None:
            # the columns of a contiguous run of atoms out of the page-locked trajectory: strided DMA
            base, row_bytes = run
            for t0 in starts:
                t1 = min(self.n_frames, t0 + seg)
                engine.copy_rows_h2d(base + t0 * row_bytes, row_bytes, coords[t0:t1], n3 * 4, t1 - t0, copy_stream)
        else:
            pinned = [torch.empty((seg, self.pitch), dtype=torch.float32).pin_memory() for _ in range(2)]
            sent = [torch.cuda.Event() for _ in range(2)]
            for k, t0 in enumerate(starts):
                t1 = min(self.n_frames, t0 + seg)
                slot = k % 2
                if k >= 2:
                    sent[slot].synchronize()          # the staging buffer has left the host
                self._host_segment(t0, t1, pinned[slot])
                with torch.cuda.stream(copy_stream):
                    coords[t0:t1].copy_(pinned[slot][:t1 - t0], non_blocking=True)
                    sent[slot].record(copy_stream)
        copy_stream.synchronize()                     # only the copies: the compute stream keeps running
        main.wait_stream(copy_stream)

    # -- dense residency -------------------------------------------------------------------------
    def _dense_bytes(self, with_images):
        per = self.pitch * 4 + (self.pitch if with_images else 0)
        return per * self.n_frames

    def _ensure_dense(self):
        torch = _lib.require_cuda()
        if self._dense is not None:
            return self._dense
        block = getattr(self.reader, 'device_block', None)
        if block is not None and self.atom_filter is None:
            coords = block
        else:
            if self.dtype not in (np.float32, np.float64):
                raise _lib.CudaPathError('positions must be float32 or float64')
            tdt = torch.float32 if self.dtype == np.float32 else torch.float64
            coords = torch.empty((self.n_frames, self.pitch), dtype=tdt, device=self.device)
            if tdt == torch.float32 and block is not None and self._filter_run:
                for t0, t1, seg in self._segments():          # a contiguous run of atoms of a device block
                    coords[t0:t1, :seg.shape[1]].copy_(seg, non_blocking=True)
            elif tdt == torch.float32:
                self._upload_rows(coords)
            else:
                n3 = 3 * self.n_atoms
                for t in range(self.n_frames):
                    pos = np.asarray(self.reader[t].positions)
                    pos = pos if self.atom_filter is None else pos[self.atom_filter]
                    coords[t, :n3] = torch.from_numpy(np.ascontiguousarray(pos).reshape(-1)).to(self.device)
        self._dense = engine.DeviceStore(coords, self.n_atoms, n_frames=self.n_frames)
        return self._dense

    def _scan_all(self, images=None, jumps=None, out_row=None, coords_out=None, images16=None):
        """Run the jump scan over the whole trajectory (segment by segment, carrying state).  Returns True when a
        running count left the range of the requested image rows (one byte: +-127, int16: +-32767)."""
        torch = _lib.require_cuda()
        n3 = 3 * self.n_atoms
        prev = torch.zeros((n3,), dtype=torch.float32, device=self.device)
        cum = torch.zeros((n3,), dtype=torch.int32, device=self.device)
        overflow = torch.zeros((1,), dtype=torch.int32, device=self.device)
        box0 = self._box_of(0)
        box_is_f32 = box0.dtype == np.float32
        whole0 = None
        ref0 = getattr(self.reader, 'whole0', None)
        if ref0 is not None:      # frame0.whole of the selected atoms (reader.py:166)
            ref0 = ref0 if self.atom_filter is None else ref0[self.atom_filter]
            if ref0.shape != (self.n_atoms, 3):
                raise ValueError('whole0 must have shape (M, 3)')
            whole0 = torch.from_numpy(np.ascontiguousarray(ref0.reshape(-1))).to(self.device)
        dense = self._dense
        if dense is not None and dense.coords.dtype == torch.float32:
            source = ((0, self.n_frames, dense.coords),)
        else:
            source = self._segments()
        for t0, t1, seg in source:
            rows = None
            if out_row is not None:
                rows = out_row[t0:t1]
                img_seg, jmp_seg, i16_seg = images, jumps, images16
            else:
                img_seg = images[t0:t1] if images is not None else None
                jmp_seg = jumps[t0:t1] if jumps is not None else None
                i16_seg = images16[t0:t1] if images16 is not None else None
            engine.nojump_scan(seg, self.n_atoms, box0, box_is_f32, prev, cum, first_segment=(t0 == 0),
                               images=img_seg, jumps=jmp_seg, out_row=rows, coords_out=coords_out,
                               overflow=overflow, whole0=whole0, images16=i16_seg)
        return bool(int(overflow.item()))

    @staticmethod
    def _too_far():
        return _lib.CudaPathError('an atom crossed more than 32767 periodic images; the image counts overflow')

    def _compress_images(self, store):
        """Byte image rows -> the packed rows the correlation kernels prefer (when no count leaves [-1, 1])."""
        if store.images is None or not self.pack_images:
            return
        packed, worst = engine.pack_images(store.images, self.n_atoms)
        if worst <= 1:
            store.packed = packed
        else:
            logger.debug('image counts reach %d: keeping one byte per component', worst)

    # -- stored jump matrices (the reference's nojump cache) ----------------------------------------
    def _matrices(self):
        """``reader.nojump_matrixes`` if the reader carries them (loaded from a ``.nojump.npz`` cache or set by
        ``generate_nojump_matrixes``), else None.  When present they DEFINE the unwrapping, as in the reference
        (pbc.py:233-237 reads nothing else), and only the frames a schedule touches have to be read."""
        mats = getattr(self.reader, '_nojump_matrixes', None)
        if mats is None:
            try:
                mats = self.reader.nojump_matrixes
            except Exception:
                return None
        if mats is None:
            return None
        if len(mats) != 3:
            raise ValueError('nojump_matrixes must hold three matrices (x, y, z)')
        for m in mats:
            if m.shape[0] != self.n_frames or m.shape[1] != self.n_atoms_all:
                raise ValueError('nojump matrix of shape %s does not fit a trajectory of %d frames x %d atoms'
                                 % (m.shape, self.n_frames, self.n_atoms_all))
        return mats

    def _images_from_matrices(self, mats, n_rows, keep=None):
        """Image rows (rows = ``keep`` ascending, or every frame) from the jump matrices: ``(images uint8 | None,
        images16 | None)`` -- int16 rows only when a count leaves one byte."""
        torch = _lib.require_cuda()
        from ..reader import jump_columns
        lists = self._jump_lists
        if lists is None or lists[0] is not mats:
            # the column lists are a pure function of the matrices: readers that share a `host_cache` dict (clones of
            # one trajectory, reader.SparseFrameReader.fresh) convert them once and keep them page-locked
            cache = getattr(self.reader, 'host_cache', None)
            key = ('jump_columns', id(mats), None if self.atom_filter is None else self.atom_filter.tobytes())
            host = cache.get(key) if cache is not None else None
            if host is None or host[0] is not mats:
                arrays = jump_columns(mats, self.atom_filter)
                host = (mats,) + tuple(torch.from_numpy(np.ascontiguousarray(a)) for a in arrays)
                if cache is not None:
                    host = (mats,) + tuple(t.pin_memory() for t in host[1:])
                    cache[key] = host
            lists = self._jump_lists = (mats,) + tuple(t.to(self.device, non_blocking=True) for t in host[1:])
        overflow = torch.zeros((1,), dtype=torch.int32, device=self.device)
        keep_dev = None
        if keep is not None:
            keep_dev = torch.from_numpy(np.ascontiguousarray(keep, dtype=np.int32)).to(self.device)
        images = torch.empty((n_rows, self.pitch), dtype=torch.uint8, device=self.device)
        engine.images_from_jumps(lists[1], lists[2], lists[3], self.n_atoms, images, keep=keep_dev, overflow=overflow)
        if not int(overflow.item()):
            return images, None
        del images
        images16 = torch.empty((n_rows, self.pitch), dtype=torch.int16, device=self.device)
        engine.images_from_jumps(lists[1], lists[2], lists[3], self.n_atoms, None, keep=keep_dev, overflow=overflow,
                                 images16=images16)
        if int(overflow.item()):
            raise self._too_far()
        return None, images16

    def _ensure_dense_images(self):
        torch = _lib.require_cuda()
        store = self._ensure_dense()
        mats = self._matrices()
        if store.has_images and store.source is not mats:
            # the reader's jump matrices were replaced (generate / load / correct_nojump_matrixes_for_whole):
            # they are authoritative (pbc.py:233-237), the stored counts are stale
            store.images = store.packed = store.images16 = None
            store._cache.clear()
        if not store.has_images:
            if store.coords.dtype != torch.float32:
                raise _lib.CudaPathError('nojump needs float32 wrapped positions')
            if mats is not None:
                store.images, store.images16 = self._images_from_matrices(mats, self.n_frames)
            else:
                images = torch.empty((self.n_frames, self.pitch), dtype=torch.uint8, device=self.device)
                if self._scan_all(images=images):
                    # beyond +-127 images: int16 rows from the general scan kernel (the reference has no limit,
                    # pbc.py:235-237 sums int8 jumps into int64)
                    del images
                    store.images16 = torch.empty((self.n_frames, self.pitch), dtype=torch.int16, device=self.device)
                    if self._scan_all(images16=store.images16):
                        raise self._too_far()
                else:
                    store.images = images
            store.source = mats
            store.box = self._box_rows(range(self.n_frames))
            store.box_const = self._box_const()
            self._compress_images(store)
        return store

    def sparse_jumps(self):
        """Per-step jumps of every frame in coordinate form, host arrays ``(frame int32, component int64, jump int8)``
        sorted by frame and component (component = 3 * atom + dimension): the scan runs segment by segment and only
        the non-zero entries leave the device, so this also serves trajectories that never fit in memory."""
        torch = _lib.require_cuda()
        n3 = 3 * self.n_atoms
        prev = torch.zeros((n3,), dtype=torch.float32, device=self.device)
        cum = torch.zeros((n3,), dtype=torch.int32, device=self.device)
        box0 = self._box_of(0)
        whole0 = None
        ref0 = getattr(self.reader, 'whole0', None)
        if ref0 is not None:
            ref0 = ref0 if self.atom_filter is None else ref0[self.atom_filter]
            whole0 = torch.from_numpy(np.ascontiguousarray(ref0.reshape(-1))).to(self.device)
        dense = self._dense
        if dense is not None and dense.coords.dtype == torch.float32:
            step = max(1, (256 << 20) // self.pitch)
            source = ((t0, min(self.n_frames, t0 + step), dense.coords[t0:t0 + step])
                      for t0 in range(0, self.n_frames, step))
        else:
            source = self._segments()
        frames, comps, vals = [], [], []
        for t0, t1, seg in source:
            jumps = torch.empty((t1 - t0, self.pitch), dtype=torch.int8, device=self.device)
            engine.nojump_scan(seg, self.n_atoms, box0, box0.dtype == np.float32, prev, cum, first_segment=(t0 == 0),
                               jumps=jumps, whole0=whole0)
            nz = torch.nonzero(jumps[:, :n3])
            if nz.numel():
                frames.append((nz[:, 0] + t0).to(torch.int32).cpu().numpy())
                comps.append(nz[:, 1].cpu().numpy())
                vals.append(jumps[:, :n3][nz[:, 0], nz[:, 1]].cpu().numpy())
            del jumps, nz
        if not frames:
            return (np.zeros(0, np.int32), np.zeros(0, np.int64), np.zeros(0, np.int8))
        return np.concatenate(frames), np.concatenate(comps), np.concatenate(vals)

    def dense_jumps(self):
        """Per-step jumps of every frame as a host int8 array (T, N, 3) -- for parse_jumps."""
        torch = _lib.require_cuda()
        n3 = 3 * self.n_atoms
        jumps = torch.empty((self.n_frames, self.pitch), dtype=torch.int8, device=self.device)
        self._scan_all(jumps=jumps)            # per-step jumps only: the running counts are not stored
        return jumps[:, :n3].cpu().numpy().reshape(self.n_frames, self.n_atoms, 3)

    # -- compact residency -------------------------------------------------------------------------
    def _build_compact(self, frame_ids, nojump):
        torch = _lib.require_cuda()
        frame_ids = np.asarray(frame_ids, dtype=np.int64)
        n_rows = len(frame_ids)
        f2r = np.full((self.n_frames,), -1, dtype=np.int32)
        f2r[frame_ids] = np.arange(n_rows, dtype=np.int32)
        n3 = 3 * self.n_atoms
        mats = self._matrices() if nojump else None
        if nojump and mats is None:
            # no stored jumps: every frame streams through the scan, the kept ones are written as they pass
            coords = torch.empty((n_rows, self.pitch), dtype=torch.float32, device=self.device)
            images = torch.empty((n_rows, self.pitch), dtype=torch.uint8, device=self.device)
            out_row = torch.from_numpy(f2r).to(self.device)
            images16 = None
            if self._scan_all(images=images, out_row=out_row, coords_out=coords):
                images = None
                images16 = torch.empty((n_rows, self.pitch), dtype=torch.int16, device=self.device)
                if self._scan_all(images16=images16, out_row=out_row):
                    raise self._too_far()
            store = engine.DeviceStore(coords, self.n_atoms, images=images, images16=images16,
                                       box=self._box_rows(frame_ids), frame_to_row=f2r, n_frames=self.n_frames,
                                       box_const=self._box_const())
            store.source = None
            self._compress_images(store)
            return store
        tdt = torch.float32 if self.dtype == np.float32 else torch.float64
        if nojump and tdt != torch.float32:
            raise _lib.CudaPathError('nojump needs float32 wrapped positions (got %s)' % self.dtype)
        coords = torch.empty((n_rows, self.pitch), dtype=tdt, device=self.device)
        block = getattr(self.reader, 'device_block', None)
        if block is not None and self.atom_filter is None:
            idx = torch.from_numpy(frame_ids).to(self.device)
            w = min(block.shape[1], self.pitch)
            coords[:, :w] = block.index_select(0, idx)[:, :w]
        elif self.atom_filter is None and getattr(self.reader, 'pinned_rows', None) is not None and \
                self.reader.pinned_rows(frame_ids) is not None:
            # the reader holds exactly these frames in page-locked memory: DMA straight out of it
            src = self.reader.pinned_rows(frame_ids)              # torch (n_rows, 3N), pinned
            step = max(1, (1 << 30) // (n3 * 4))
            for r0 in range(0, n_rows, step):
                coords[r0:r0 + step, :n3].copy_(src[r0:r0 + step], non_blocking=True)
        else:
            # only the kept frames are read from the reader
            chunk = max(1, (256 << 20) // (self.pitch * coords.element_size()))
            host = torch.empty((min(chunk, n_rows), self.pitch), dtype=tdt).pin_memory()
            for r0 in range(0, n_rows, chunk):
                r1 = min(n_rows, r0 + chunk)
                view = host[:r1 - r0, :n3].numpy().reshape(r1 - r0, self.n_atoms, 3)
                for r in range(r0, r1):
                    pos = np.asarray(self.reader[int(frame_ids[r])].positions)
                    view[r - r0] = pos if self.atom_filter is None else pos[self.atom_filter]
                coords[r0:r1].copy_(host[:r1 - r0], non_blocking=True)
                torch.cuda.current_stream(self.device).synchronize()
        if not nojump:
            return engine.DeviceStore(coords, self.n_atoms, frame_to_row=f2r, n_frames=self.n_frames)
        images, images16 = self._images_from_matrices(mats, n_rows, keep=frame_ids)
        store = engine.DeviceStore(coords, self.n_atoms, images=images, images16=images16,
                                   box=self._box_rows(frame_ids), frame_to_row=f2r, n_frames=self.n_frames,
                                   box_const=self._box_const())
        store.source = mats
        self._compress_images(store)
        return store

    # -- public ---------------------------------------------------------------------------------
    def on_device(self):
        """Whether the trajectory already lives in HBM (a device block or a dense store)."""
        block = getattr(self.reader, 'device_block', None)
        return self._dense is not None or (block is not None and self.atom_filter is None)

    def memory_budget(self):
        return _mem_budget(_lib.require_cuda(), self.device)

    def atom_slab(self, a0, a1):
        """A Resident over atoms [a0, a1) of this selection (out-of-core passes, slab by slab).  Not cached
        on the reader: its HBM store lives as long as the caller holds it."""
        own = np.arange(self.n_atoms_all) if self.atom_filter is None else self.atom_filter
        return Resident(self.reader, own[a0:a1], device=self.device)

    def serves(self, store, nojump):
        """Whether ``store`` is still one of this resident's stores and up to date (correlation._provider)."""
        if self.force_compact and store is self._dense:
            return False
        if store is not self._dense and not any(store is s for s in self._compact.values()):
            return False
        return not nojump or (store.has_images and store.source is self._matrices())

    def store_for(self, frame_ids, nojump):
        """A DeviceStore holding (at least) ``frame_ids``; with image counts if ``nojump``."""
        if (self._dense is not None and not self.force_compact
                and (not nojump or (self._dense.has_images and self._dense.source is self._matrices()))):
            return self._dense                      # hot path: everything is resident already
        torch = _lib.require_cuda()
        frame_ids = np.unique(np.asarray(frame_ids, dtype=np.int64))
        if frame_ids.size and (frame_ids[0] < 0 or frame_ids[-1] >= self.n_frames):
            raise IndexError('frame index out of range')
        block = getattr(self.reader, 'device_block', None)
        # a file-backed reader whose jumps are already stored: read only the frames the schedule touches
        # (what the reference does, reader.py:254-286 + pbc.py:215-247) instead of the whole trajectory
        sparse_io = (nojump and self._dense is None and block is None
                     and getattr(self.reader, 'host_block', None) is None and self._matrices() is not None)
        if not self.force_compact and not sparse_io:
            on_device = self._dense is not None or (block is not None and self.atom_filter is None)
            have_images = self._dense is not None and self._dense.has_images
            extra = 0 if on_device else self._dense_bytes(False)
            if nojump and not have_images:
                extra += self.pitch * self.n_frames
            if extra <= _mem_budget(torch, self.device):
                return self._ensure_dense_images() if nojump else self._ensure_dense()
        mode = 'nojump' if nojump else 'raw'
        key = (mode, frame_ids.tobytes())
        if nojump:
            # stored jump matrices are authoritative: stores whose counts came from other matrices (or from the
            # scan, before matrices were loaded) are stale
            mats = self._matrices()
            for k in [k for k, st in self._compact.items() if k[0] == 'nojump' and st.source is not mats]:
                del self._compact[k]
        store = self._compact.get(key)
        if store is None:
            for (have_mode, _), have in self._compact.items():       # e.g. single frames of a schedule's store
                if have_mode == mode and have.has_frames(frame_ids):
                    return have
            logger.debug('building compact store of %d frames (nojump=%s)', len(frame_ids), nojump)
            if len(frame_ids) > 4 or len(self._compact) > 16:
                self._compact.clear()    # one schedule-sized compact store at a time: they are large
            store = self._build_compact(frame_ids, nojump)
            self._compact[key] = store
        return store

    def unwrapped_frame(self, i):
        """float64 (N, 3) nojump frame ``x - n*box`` (pbc.py:235-247) as a host array."""
        store = self.store_for([i], nojump=True)
        row = int(store.rows_for([i])[0])
        n3 = 3 * self.n_atoms
        if store.images is not None:
            img, fmt = store.images[row], _lib.IMAGE_BYTES
        elif store.images16 is not None:
            img, fmt = store.images16[row], _lib.IMAGE_I16
        else:
            img, fmt = store.packed[row], _lib.IMAGE_PACKED
        out = engine.nojump_apply(store.coords[row], img, store.box[row].cpu().numpy(), self.n_atoms, fmt)
        return out.cpu().numpy().reshape(self.n_atoms, 3)

    def raw_frame(self, i):
        pos = np.asarray(self.reader[i].positions)
        return pos if self.atom_filter is None else pos[self.atom_filter]
