"""ctypes binding of libmdevaluate_b200.so (the C ABI in include/mdevaluate_b200.h).

There is no CPU fallback: if the shared library cannot be loaded (and cannot be
built because nvcc is absent) every compute entry point raises ``CudaPathError``.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libmdevaluate_b200.so')

# flags (keep in sync with include/mdevaluate_b200.h)
F32, F64 = 0, 1
SUM_R2, SUM_R4, SUM_SINC, HIST, SUM_LEGENDRE = 1, 2, 4, 8, 16
BINS_UNIFORM, BINS_EDGES = 0, 1
IMAGE_BYTES, IMAGE_PACKED, IMAGE_I16 = 1, 2, 3
ERR_TIMEOUT = 5
COMM_HANDLE_BYTES = 64
PLAN_TIME_SLOTS = 1024
ABI_VERSION = 2

EXPORTS = (
    'mdb_last_error', 'mdb_abi_version', 'mdb_device_info',
    'mdb_corr_workspace_bytes', 'mdb_shifted_corr', 'mdb_corr_last_launches',
    'mdb_corr_profile', 'mdb_corr_last_kernel_ms',
    'mdb_nojump_scan', 'mdb_nojump_apply', 'mdb_nojump_images_from_jumps',
    'mdb_dense_workspace_bytes', 'mdb_dense_corr', 'mdb_copy_rows_h2d',
    'mdb_corr_plan_create', 'mdb_corr_plan_run', 'mdb_corr_plan_launch', 'mdb_corr_plan_wait',
    'mdb_corr_plan_kernel_ms', 'mdb_corr_plan_sums', 'mdb_corr_plan_hist', 'mdb_corr_plan_sums_device',
    'mdb_corr_plan_launches', 'mdb_corr_plan_destroy', 'mdb_corr_plan_set_wait', 'mdb_corr_plan_timeline',
    'mdb_clock_pair',
    'mdb_comm_create', 'mdb_comm_handle', 'mdb_comm_connect', 'mdb_comm_destroy',
    'mdb_images_pack', 'mdb_nojump_apply_format', 'mdb_vectors', 'mdb_overlap_workspace_bytes', 'mdb_overlap',
    'mdb_overlap_launches',
)


class CudaPathError(RuntimeError):
    """The CUDA path is unavailable or a C-ABI call failed."""


class Tuning(C.Structure):
    """mdb_tuning: explicit kernel switches, 0 = the measured default."""
    _fields_ = [('ring_stages', C.c_int32), ('atom_blocks', C.c_int32), ('isf_float64', C.c_int32),
                ('hist_float64', C.c_int32), ('scan_direct', C.c_int32), ('dense_stage_cta', C.c_int32),
                ('dense_float64', C.c_int32), ('variant', C.c_int32)]


class CorrArgs(C.Structure):
    _fields_ = [
        ('coords', C.c_void_p), ('pitch', C.c_int64), ('coord_type', C.c_int32),
        ('images', C.c_void_p), ('img_pitch', C.c_int64), ('box', C.c_void_p), ('box_const', C.c_double * 3),
        ('n_atoms', C.c_int32), ('n_origins', C.c_int32), ('n_lags', C.c_int32),
        ('rows', C.c_void_p), ('select', C.c_void_p), ('sel_pitch', C.c_int64), ('flags', C.c_int32), ('q', C.c_double),
        ('legendre_order', C.c_int32),
        ('n_bins', C.c_int32), ('bin_mode', C.c_int32), ('dx', C.c_double), ('edges', C.c_void_p),
        ('sums', C.c_void_p), ('hist', C.c_void_p),
        ('workspace', C.c_void_p), ('workspace_bytes', C.c_size_t),
        ('packed', C.c_void_p), ('packed_pitch', C.c_int64), ('images16', C.c_void_p), ('img16_pitch', C.c_int64),
        ('tuning', Tuning), ('n_segments', C.c_int32),
    ]


class NojumpArgs(C.Structure):
    _fields_ = [
        ('coords', C.c_void_p), ('pitch', C.c_int64), ('n_frames', C.c_int32), ('n_atoms', C.c_int32),
        ('box0', C.c_double * 3), ('box_is_f32', C.c_int32), ('first_segment', C.c_int32),
        ('whole0', C.c_void_p), ('prev', C.c_void_p), ('cum', C.c_void_p), ('out_row', C.c_void_p),
        ('images', C.c_void_p), ('img_pitch', C.c_int64), ('jumps', C.c_void_p),
        ('coords_out', C.c_void_p), ('pitch_out', C.c_int64), ('overflow', C.c_void_p),
        ('images16', C.c_void_p), ('img16_pitch', C.c_int64), ('tuning', Tuning),
    ]


class JumpMatArgs(C.Structure):
    _fields_ = [
        ('indptr', C.c_void_p), ('frames', C.c_void_p), ('jumps', C.c_void_p), ('nnz', C.c_int64),
        ('n_atoms', C.c_int32), ('keep', C.c_void_p), ('n_keep', C.c_int32),
        ('images', C.c_void_p), ('img_pitch', C.c_int64), ('overflow', C.c_void_p),
        ('images16', C.c_void_p), ('img16_pitch', C.c_int64),
    ]


class DenseArgs(C.Structure):
    _fields_ = [
        ('coords', C.c_void_p), ('pitch', C.c_int64), ('images', C.c_void_p), ('img_pitch', C.c_int64),
        ('box', C.c_void_p), ('box_const', C.c_double * 3), ('frame_rows', C.c_void_p),
        ('n_atoms', C.c_int32), ('first_origin', C.c_int32), ('origin_step', C.c_int32), ('n_origins', C.c_int32),
        ('first_lag', C.c_int32), ('n_lags', C.c_int32), ('flags', C.c_int32), ('q', C.c_double),
        ('sums', C.c_void_p), ('workspace', C.c_void_p), ('workspace_bytes', C.c_size_t),
        ('packed', C.c_void_p), ('packed_pitch', C.c_int64), ('tuning', Tuning),
    ]


class VectorsArgs(C.Structure):
    _fields_ = [
        ('coords', C.c_void_p), ('pitch', C.c_int64), ('images', C.c_void_p), ('image_format', C.c_int32),
        ('img_pitch', C.c_int64), ('box', C.c_void_p), ('box_const', C.c_double * 3), ('rows', C.c_void_p),
        ('n_rows', C.c_int32), ('atoms_a', C.c_void_p), ('m_a', C.c_int32), ('atoms_b', C.c_void_p), ('m_b', C.c_int32),
        ('n_vec', C.c_int32), ('normed', C.c_int32), ('use_pbc', C.c_int32), ('pbc_box', C.c_double * 3),
        ('out_type', C.c_int32), ('out', C.c_void_p), ('out_pitch', C.c_int64),
    ]


class OverlapArgs(C.Structure):
    _fields_ = [
        ('ref_coords', C.c_void_p), ('ref_pitch', C.c_int64), ('ref_type', C.c_int32), ('n_ref', C.c_int32),
        ('ref_images', C.c_void_p), ('ref_image_format', C.c_int32), ('reserved0', C.c_int32),
        ('ref_img_pitch', C.c_int64), ('ref_box', C.c_void_p), ('ref_box_const', C.c_double * 3),
        ('ref_rows', C.c_void_p), ('ref_select', C.c_void_p), ('ref_sel_pitch', C.c_int64),
        ('coords', C.c_void_p), ('pitch', C.c_int64), ('coord_type', C.c_int32), ('n_atoms', C.c_int32),
        ('images', C.c_void_p), ('image_format', C.c_int32), ('file_order', C.c_int32), ('img_pitch', C.c_int64),
        ('box', C.c_void_p), ('box_const', C.c_double * 3), ('rows', C.c_void_p),
        ('n_origins', C.c_int32), ('n_lags', C.c_int32), ('radius', C.c_double), ('boxsize', C.c_double * 3),
        ('counts', C.c_void_p), ('tree_sizes', C.c_void_p), ('bounds', C.c_void_p), ('out_of_box', C.c_void_p),
        ('workspace', C.c_void_p), ('workspace_bytes', C.c_size_t),
    ]


_lock = threading.Lock()
_lib = None


def _declare(lib):
    lib.mdb_last_error.restype = C.c_char_p
    lib.mdb_last_error.argtypes = []
    lib.mdb_abi_version.restype = C.c_int
    lib.mdb_device_info.argtypes = [C.POINTER(C.c_int)] * 3
    lib.mdb_corr_workspace_bytes.argtypes = [C.POINTER(CorrArgs), C.POINTER(C.c_size_t)]
    lib.mdb_shifted_corr.argtypes = [C.POINTER(CorrArgs), C.c_void_p]
    lib.mdb_corr_last_launches.restype = C.c_int
    lib.mdb_corr_profile.argtypes = [C.c_int]
    lib.mdb_corr_last_kernel_ms.argtypes = [C.POINTER(C.c_float)]
    lib.mdb_nojump_scan.argtypes = [C.POINTER(NojumpArgs), C.c_void_p]
    lib.mdb_nojump_images_from_jumps.argtypes = [C.POINTER(JumpMatArgs), C.c_void_p]
    lib.mdb_copy_rows_h2d.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p]
    lib.mdb_dense_workspace_bytes.argtypes = [C.POINTER(DenseArgs), C.POINTER(C.c_size_t)]
    lib.mdb_dense_corr.argtypes = [C.POINTER(DenseArgs), C.c_void_p]
    lib.mdb_nojump_apply.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.c_int32, C.c_void_p,
                                     C.c_void_p]
    lib.mdb_corr_plan_create.argtypes = [C.POINTER(CorrArgs), C.c_void_p, C.POINTER(C.c_void_p)]
    lib.mdb_corr_plan_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    lib.mdb_corr_plan_launch.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    lib.mdb_corr_plan_wait.argtypes = [C.c_void_p]
    lib.mdb_corr_plan_kernel_ms.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_float)]
    for name in ('mdb_corr_plan_sums', 'mdb_corr_plan_hist', 'mdb_corr_plan_sums_device'):
        getattr(lib, name).argtypes = [C.c_void_p]
    lib.mdb_corr_plan_launches.argtypes = [C.c_void_p]
    lib.mdb_corr_plan_timeline.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong)]
    lib.mdb_clock_pair.argtypes = [C.c_void_p] + [C.POINTER(C.c_ulonglong)] * 3
    lib.mdb_corr_plan_set_wait.argtypes = [C.c_void_p, C.c_int32]
    lib.mdb_corr_plan_destroy.argtypes = [C.c_void_p]
    lib.mdb_comm_create.argtypes = [C.c_int32, C.c_int32, C.c_size_t, C.c_int32, C.POINTER(C.c_void_p)]
    lib.mdb_comm_handle.argtypes = [C.c_void_p, C.c_void_p]
    lib.mdb_comm_connect.argtypes = [C.c_void_p, C.c_void_p]
    lib.mdb_comm_destroy.argtypes = [C.c_void_p]
    lib.mdb_vectors.argtypes = [C.POINTER(VectorsArgs), C.c_void_p]
    lib.mdb_overlap_workspace_bytes.argtypes = [C.POINTER(OverlapArgs), C.POINTER(C.c_size_t)]
    lib.mdb_overlap.argtypes = [C.POINTER(OverlapArgs), C.c_void_p]
    lib.mdb_overlap_launches.restype = C.c_int
    lib.mdb_images_pack.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p,
                                    C.c_void_p]
    lib.mdb_nojump_apply_format.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_double), C.c_int32,
                                            C.c_void_p, C.c_void_p]
    pointers = ('mdb_last_error', 'mdb_corr_plan_sums', 'mdb_corr_plan_hist', 'mdb_corr_plan_sums_device')
    for name in pointers[1:]:
        getattr(lib, name).restype = C.c_void_p
    for name in EXPORTS:
        if name not in pointers:
            getattr(lib, name).restype = C.c_int        # every other entry point returns an int status / count


def load(build_if_missing=True):
    """Load (building first if the library is absent and nvcc exists) and return the CDLL."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH) and build_if_missing:
            try:
                from ..csrc.build import build
                build()
            except Exception as exc:  # pragma: no cover - depends on toolchain
                raise CudaPathError(
                    'libmdevaluate_b200.so is missing and could not be built (%s); run '
                    '`python -m mdevaluate_b200.csrc.build`. There is no CPU fallback.' % exc) from exc
        if not os.path.exists(LIB_PATH):
            raise CudaPathError('libmdevaluate_b200.so not found at %s; there is no CPU fallback' % LIB_PATH)
        try:
            lib = C.CDLL(LIB_PATH)
        except OSError as exc:
            raise CudaPathError('cannot load %s: %s' % (LIB_PATH, exc)) from exc
        missing = [n for n in EXPORTS if not hasattr(lib, n)]
        if missing:
            raise CudaPathError('libmdevaluate_b200.so lacks symbols: %s' % ', '.join(missing))
        _declare(lib)
        if lib.mdb_abi_version() != ABI_VERSION:
            raise CudaPathError('ABI version mismatch: library %d, binding %d'
                                % (lib.mdb_abi_version(), ABI_VERSION))
        _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        msg = load().mdb_last_error()
        raise CudaPathError('%s failed (status %d): %s' % (what, rc, msg.decode() if msg else ''))


def require_cuda():
    """Torch with a visible CUDA device, or a loud error (no CPU fallback on the product path)."""
    import torch
    if not torch.cuda.is_available():
        raise CudaPathError('no CUDA device visible: mdevaluate_b200 has no CPU fallback '
                            '(the CPU oracle lives in oracle/ and is test infrastructure only)')
    return torch
