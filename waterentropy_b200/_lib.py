"""ctypes binding of the C ABI declared in include/waterentropy_b200.h.

The product has no CPU fallback: if the shared library is missing, or no CUDA
device is present when a context is created, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _build

c_int32_p = C.POINTER(C.c_int32)
WE_VERSION = 202  # include/waterentropy_b200.h; a stale build is refused


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
        ("force_upload", C.c_int32), ("shells_only", C.c_int32), ("eig_flavour", C.c_int32),
        ("pos_upload", C.c_int32),
    ]


class we_diag(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "wide_searches", "shrink_retries", "table_entries",
        "kernel_launches", "frames_done", "recorded", "fast_path_mismatches", "eig_fallbacks",
        "h2d_dma_bytes", "light_atoms_fetched", "host_pack_us")]


EXPORTS = [
    "we_last_error", "we_version", "we_create", "we_destroy", "we_submit", "we_submit_device",
    "we_sync", "we_get_diag", "we_error_location", "we_n_bins", "we_copy_cov", "we_device_tables",
    "we_n_label_entries", "we_copy_label_entries", "we_frame_record_count", "we_copy_frame_records",
    "we_n_heavy", "we_n_donors", "we_copy_heavy_idx", "we_copy_donors", "we_debug_size",
    "we_debug_copy", "we_reset", "we_test_angle_cos", "we_test_powf2", "we_test_pow2", "we_test_distance",
    "we_profile", "we_kernel_times", "we_span_start", "we_span_stop",
    "we_label_entries_device", "we_merge_label_entries", "we_snapshot_cov", "we_wait_cov",
    "we_copy_frame_shells", "we_test_eig3", "we_frames_retired", "we_wait_frames_retired",
    "we_molecule_force_torque", "we_molecule_force_torque_multi", "we_batch_frames", "we_vib_entropy", "we_orient_entropy",
    "we_copy_frame_records_range", "we_convert_be_f32",
]

WE_ERR_DUPLICATE_COORD = -1
WE_ERR_NO_NEIGHBOURS = -2

# we_options.eig_flavour: which OpenBLAS kernel family the principal-axis step replays
EIG_OPENBLAS_AVX512 = 0
EIG_OPENBLAS_AVX2 = 1
_EIG_NAMES = {"avx512": EIG_OPENBLAS_AVX512, "skylakex": EIG_OPENBLAS_AVX512, "0": EIG_OPENBLAS_AVX512,
              "avx2": EIG_OPENBLAS_AVX2, "haswell": EIG_OPENBLAS_AVX2, "zen": EIG_OPENBLAS_AVX2, "1": EIG_OPENBLAS_AVX2}
_AVX512_CORES = ("SKYLAKEX", "COOPERLAKE", "SAPPHIRERAPIDS")
_AVX2_CORES = ("HASWELL", "ZEN", "BROADWELL", "EXCAVATOR")
_host_flavour = None


def numpy_openblas_core():
    """Name of the kernel set NumPy's bundled OpenBLAS dispatched to on this host
    (`openblas_get_corename`), or None when NumPy is not running on an OpenBLAS we can query."""
    import numpy as np

    np.linalg.eig(np.eye(2))  # make sure the LAPACK library is mapped
    paths = []
    try:
        with open("/proc/self/maps") as fh:
            for line in fh:
                path = line.rsplit(" ", 1)[-1].strip()
                if "openblas" in os.path.basename(path).lower() and path not in paths:
                    paths.append(path)
    except OSError:
        return None
    for path in paths:
        try:
            lib = C.CDLL(path)
        except OSError:
            continue
        for sym in ("scipy_openblas_get_corename64_", "scipy_openblas_get_corename", "openblas_get_corename64_",
                    "openblas_get_corename"):
            fn = getattr(lib, sym, None)
            if fn is not None:
                fn.restype = C.c_char_p
                name = fn()
                return name.decode() if name else None
    return None


def host_eig_flavour():
    """(flavour, source, pinned): the OpenBLAS kernel family `numpy.linalg.eig` runs on THIS host, which
    is what the reference's eigenvector signs follow (maths/transformations.py:121-124).  `WE_EIG_FLAVOUR`
    (avx512 | avx2) overrides.  `pinned` is False when the host's BLAS could not be identified; the
    AVX-512 kernels (the ones the golden vectors were made with) are replayed then."""
    global _host_flavour
    env = os.environ.get("WE_EIG_FLAVOUR", "").strip().lower()
    if env:
        if env not in _EIG_NAMES:
            raise ValueError(f"WE_EIG_FLAVOUR must be one of {sorted(_EIG_NAMES)}, got {env!r}")
        return _EIG_NAMES[env], f"WE_EIG_FLAVOUR={env}", True
    if _host_flavour is None:
        core = numpy_openblas_core()
        up = (core or "").upper()
        if any(k in up for k in _AVX512_CORES):
            _host_flavour = (EIG_OPENBLAS_AVX512, f"numpy OpenBLAS core {core}", True)
        elif any(k in up for k in _AVX2_CORES):
            _host_flavour = (EIG_OPENBLAS_AVX2, f"numpy OpenBLAS core {core}", True)
        else:
            _host_flavour = (EIG_OPENBLAS_AVX512, f"numpy BLAS core {core!r} not restated", False)
    return _host_flavour

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
    if L.we_version() != WE_VERSION:
        raise ImportError(f"{path} is ABI version {L.we_version()}, this package needs {WE_VERSION}: rebuild it "
                          "(__graft_entry__.build())")
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
    L.we_convert_be_f32.restype = I
    L.we_convert_be_f32.argtypes = [V, V, I, C.c_uint64, I, C.c_float, I]
    L.we_reset.restype = I
    L.we_reset.argtypes = [V]
    L.we_get_diag.restype = I
    L.we_get_diag.argtypes = [V, C.POINTER(we_diag)]
    L.we_error_location.restype = I
    L.we_error_location.argtypes = [V, C.POINTER(I64), C.POINTER(I)]
    for name in ("we_n_bins", "we_n_label_entries", "we_n_heavy", "we_n_donors", "we_batch_frames"):
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
    L.we_copy_frame_records_range.restype = I64
    L.we_copy_frame_records_range.argtypes = [V, I64, I, V, V, I64]
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
    L.we_frames_retired.restype = I64
    L.we_frames_retired.argtypes = [V]
    L.we_wait_frames_retired.restype = I
    L.we_wait_frames_retired.argtypes = [V, I64]
    L.we_molecule_force_torque.restype = I
    L.we_molecule_force_torque.argtypes = [I, V, V, V, V, I, I, V]
    L.we_molecule_force_torque_multi.restype = I
    L.we_molecule_force_torque_multi.argtypes = [I, V, V, V, I, V, V, I, V, V, V, V, V, V, V, V, V, I, V, V]
    L.we_vib_entropy.restype = I
    L.we_vib_entropy.argtypes = [V, C.c_double, C.c_double, V]
    L.we_orient_entropy.restype = I
    L.we_orient_entropy.argtypes = [V, V, I, I, V, V, V]
    L.we_test_eig3.restype = I
    L.we_test_eig3.argtypes = [I, V, I, I, V]
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


def convert_be_f32(src_addrs, dst_addrs, n_each, op=0, factor=1.0, n_threads=0):
    """Big-endian float32 blocks at `src_addrs` -> native float32 at `dst_addrs` (`n_each` values per block),
    `op` 0 copy / 1 multiply / 2 divide by `factor` in float32 (we_convert_be_f32: the threaded byte swap +
    unit conversion of the trajectory readers)."""
    import numpy as np

    src = np.ascontiguousarray(src_addrs, dtype=np.uint64)
    dst = np.ascontiguousarray(dst_addrs, dtype=np.uint64)
    assert src.shape == dst.shape and src.ndim == 1
    check(load().we_convert_be_f32(src.ctypes.data, dst.ctypes.data, len(src), int(n_each), int(op), float(factor),
                                   int(n_threads)))
