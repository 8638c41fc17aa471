"""ctypes binding of the C ABI declared in include/waterentropy_b200.h.

The product has no CPU fallback: if the shared library is missing, or no CUDA
device is present when a context is created, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _build

c_int32_p = C.POINTER(C.c_int32)


class we_topology(C.Structure):
    _fields_ = [
        ("n_atoms", C.c_int32),
        ("masses", C.c_void_p), ("charges", C.c_void_p),
        ("resid", C.c_void_p), ("resname_id", C.c_void_p), ("type_id", C.c_void_p),
        ("is_water", C.c_void_p),
        ("n_bonds", C.c_int32), ("bonds", C.c_void_p),
        ("n_solute_ua", C.c_int32), ("solute_ua", C.c_void_p),
        ("n_centres", C.c_int32), ("centres", C.c_void_p),
        ("frag_ptr", C.c_void_p), ("frag_atoms", C.c_void_p),
        ("bin_of_atom", C.c_void_p), ("centre_class", C.c_void_p),
        ("n_pair_bins", C.c_int32), ("n_classes", C.c_int32), ("recipe", C.c_int32),
    ]


class we_options(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("chunk_frames", C.c_int32), ("table_log2", C.c_int32),
        ("keep_records", C.c_int32), ("keep_debug", C.c_int32),
        ("search_radius", C.c_float), ("max_box", C.c_float), ("keep_shells", C.c_int32),
        ("force_upload", C.c_int32),
    ]


class we_diag(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "wide_searches", "shrink_retries", "table_entries",
        "kernel_launches", "frames_done", "recorded", "fast_path_mismatches")]


EXPORTS = [
    "we_last_error", "we_version", "we_create", "we_destroy", "we_submit", "we_submit_device",
    "we_sync", "we_get_diag", "we_error_location", "we_n_bins", "we_copy_cov", "we_device_tables",
    "we_n_label_entries", "we_copy_label_entries", "we_frame_record_count", "we_copy_frame_records",
    "we_n_heavy", "we_n_donors", "we_copy_heavy_idx", "we_copy_donors", "we_debug_size",
    "we_debug_copy", "we_reset", "we_test_angle_cos", "we_test_powf2", "we_test_pow2", "we_test_distance",
    "we_profile", "we_kernel_times", "we_span_start", "we_span_stop",
    "we_label_entries_device", "we_merge_label_entries", "we_snapshot_cov", "we_wait_cov",
    "we_copy_frame_shells",
]

WE_ERR_DUPLICATE_COORD = -1
WE_ERR_NO_NEIGHBOURS = -2

_lib = None


def library_path() -> str:
    return os.environ.get("WE_LIB_PATH") or _build.SO  # WE_LIB_PATH: tuning builds only


def load():
    """Load libwaterentropy_b200.so (must have been built: `python -c 'import __graft_entry__ as g; g.build()'`)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: the CUDA extension has not been built and waterentropy_b200 "
            "has no CPU fallback.  Run __graft_entry__.build() (needs nvcc)."
        )
    L = C.CDLL(path)
    V, I, I64 = C.c_void_p, C.c_int32, C.c_int64
    L.we_last_error.restype = C.c_char_p
    L.we_version.restype = I
    L.we_create.restype = I
    L.we_create.argtypes = [C.POINTER(we_topology), C.POINTER(we_options), C.POINTER(V)]
    L.we_destroy.restype = None
    L.we_destroy.argtypes = [V]
    for name in ("we_submit", "we_submit_device"):
        fn = getattr(L, name)
        fn.restype = I
        fn.argtypes = [V, V, V, V, V, I]
    L.we_sync.restype = I
    L.we_sync.argtypes = [V]
    L.we_reset.restype = I
    L.we_reset.argtypes = [V]
    L.we_get_diag.restype = I
    L.we_get_diag.argtypes = [V, C.POINTER(we_diag)]
    L.we_error_location.restype = I
    L.we_error_location.argtypes = [V, C.POINTER(I64), C.POINTER(I)]
    for name in ("we_n_bins", "we_n_label_entries", "we_n_heavy", "we_n_donors"):
        fn = getattr(L, name)
        fn.restype = I
        fn.argtypes = [V]
    L.we_copy_cov.restype = I
    L.we_copy_cov.argtypes = [V, V, V]
    L.we_device_tables.restype = I
    L.we_device_tables.argtypes = [V, C.POINTER(V), C.POINTER(V)]
    L.we_copy_label_entries.restype = I
    L.we_copy_label_entries.argtypes = [V, V, I]
    L.we_snapshot_cov.restype = I
    L.we_snapshot_cov.argtypes = [V]
    L.we_wait_cov.restype = I
    L.we_wait_cov.argtypes = [V, I, V, V]
    L.we_label_entries_device.restype = I
    L.we_label_entries_device.argtypes = [V, C.POINTER(V)]
    L.we_merge_label_entries.restype = I
    L.we_merge_label_entries.argtypes = [V, V, I]
    L.we_frame_record_count.restype = I
    L.we_frame_record_count.argtypes = [V, I64]
    L.we_copy_frame_records.restype = I
    L.we_copy_frame_records.argtypes = [V, I64, V, I]
    L.we_copy_frame_shells.restype = I
    L.we_copy_frame_shells.argtypes = [V, I64, V, I]
    L.we_copy_heavy_idx.restype = I
    L.we_copy_heavy_idx.argtypes = [V, V]
    L.we_copy_donors.restype = I
    L.we_copy_donors.argtypes = [V, V, V]
    L.we_debug_size.restype = I64
    L.we_debug_size.argtypes = [V, I64, I]
    L.we_debug_copy.restype = I
    L.we_debug_copy.argtypes = [V, I64, I, V, I64]
    L.we_profile.restype = I
    L.we_profile.argtypes = [V, I]
    L.we_kernel_times.restype = I
    L.we_kernel_times.argtypes = [V, V, V]
    L.we_span_start.restype = I
    L.we_span_start.argtypes = [V]
    L.we_span_stop.restype = I
    L.we_span_stop.argtypes = [V, C.POINTER(C.c_double)]
    L.we_test_angle_cos.restype = I
    L.we_test_angle_cos.argtypes = [I, V, V, I, V]
    L.we_test_powf2.restype = I
    L.we_test_powf2.argtypes = [I, V, I, V]
    L.we_test_pow2.restype = I
    L.we_test_pow2.argtypes = [I, V, I, V]
    L.we_test_distance.restype = I
    L.we_test_distance.argtypes = [I, V, V, I, V]
    _lib = L
    return L


class WaterEntropyError(RuntimeError):
    pass


def check(rc: int):
    """Map C-ABI status codes to the reference's exception behaviour."""
    if rc >= 0:
        return rc
    msg = load().we_last_error().decode("utf-8", "replace")
    if rc in (WE_ERR_DUPLICATE_COORD, WE_ERR_NO_NEIGHBOURS):
        # maths/trig.py:36-42 raises ValueError in both situations
        raise ValueError(msg)
    raise WaterEntropyError(f"waterentropy_b200 error {rc}: {msg}")
