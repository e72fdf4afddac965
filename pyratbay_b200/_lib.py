"""ctypes binding of the C ABI declared in include/pb200_lbl.h.

The product path has no CPU fallback: if the CUDA library is missing or no device is
usable, every entry point raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PB200_LIB selects an alternative build of the same library (kernel-tuning experiments).
_LIB_PATH = os.environ.get("PB200_LIB") or os.path.join(_HERE, "libpb200_lbl.so")
_lib = None

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int64_p = ctypes.POINTER(ctypes.c_int64)
c_void_p = ctypes.c_void_p

EXPORTS = [
    "pb200_last_error", "pb200_version", "pb200_device_count", "pb200_voigt_grid",
    "pb200_engine_create", "pb200_engine_destroy", "pb200_engine_set_grid",
    "pb200_engine_build_voigt", "pb200_engine_set_voigt", "pb200_engine_profile_len",
    "pb200_engine_get_profile", "pb200_engine_set_species", "pb200_engine_set_partition",
    "pb200_engine_set_lines", "pb200_engine_line_stats", "pb200_extinction_batch_host",
    "pb200_extinction_batch_dev", "pb200_engine_last_timing", "pb200_engine_launch_count",
    "pb200_engine_dense_units", "pb200_engine_dense_ms",
    "pb200_interp_ec", "pb200_interp_ec_per_mol", "pb200_interp_ec_dev",
    "pb200_engine_stream", "pb200_bench_fp64", "pb200_bench_l2",
    "pb200_nearest_thresholds", "pb200_selftest_exact",
    "pb200_optical_depth", "pb200_optical_depth_dev",
    "pb200_table_create", "pb200_table_destroy", "pb200_table_interp",
    "pb200_table_launch_count", "pb200_regrid_table_dev",
]


class PB200Error(RuntimeError):
    """Raised when a C-ABI call returns a negative status."""


def lib_path():
    return _LIB_PATH


def load():
    """Load libpb200_lbl.so (raises if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise PB200Error(
            f"CUDA library not found at {_LIB_PATH}. Build it with "
            "`python -m pyratbay_b200.build` (there is no CPU fallback).")
    lib = ctypes.CDLL(_LIB_PATH)
    lib.pb200_last_error.restype = ctypes.c_char_p
    lib.pb200_version.restype = ctypes.c_char_p
    lib.pb200_engine_profile_len.restype = ctypes.c_int64
    lib.pb200_engine_profile_len.argtypes = [c_void_p]
    lib.pb200_engine_launch_count.restype = ctypes.c_int64
    lib.pb200_engine_launch_count.argtypes = [c_void_p]
    lib.pb200_engine_dense_units.restype = ctypes.c_int64
    lib.pb200_engine_dense_units.argtypes = [c_void_p]
    lib.pb200_engine_dense_ms.restype = ctypes.c_double
    lib.pb200_engine_dense_ms.argtypes = [c_void_p]
    lib.pb200_engine_stream.restype = c_void_p
    lib.pb200_engine_stream.argtypes = [c_void_p]
    lib.pb200_table_destroy.restype = None
    lib.pb200_table_destroy.argtypes = [c_void_p]
    lib.pb200_table_launch_count.restype = ctypes.c_int64
    lib.pb200_table_launch_count.argtypes = [c_void_p]
    lib.pb200_engine_destroy.restype = None
    lib.pb200_engine_destroy.argtypes = [c_void_p]
    _lib = lib
    return lib


def check(status):
    if status != 0:
        msg = load().pb200_last_error().decode(errors="replace")
        raise PB200Error(f"pb200 error {status}: {msg}")


def device_count():
    return int(load().pb200_device_count())


def require_device():
    if device_count() < 1:
        raise PB200Error(
            "no usable CUDA device: pyratbay_b200 runs only on a GPU (no CPU fallback)")
