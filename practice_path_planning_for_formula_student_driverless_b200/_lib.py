"""Loader of csrc/libraceline_b200.so (the C ABI of include/raceline_b200.h).

The product path has no CPU fallback: if the CUDA extension is missing every call fails loudly here.
"""
from __future__ import annotations

import ctypes as C
import os

from ._abi import RlBatchDesc, RlBatchOut, RlGeomDesc, RlGeomOut, RlJobStats, RlParams

_HERE = os.path.dirname(os.path.abspath(__file__))
# RL_LIB_VARIANT selects a tuning / debug build made by `build.py --suffix` (e.g. "_dbg"); unset = the product library
LIB_PATH = os.path.join(_HERE, "csrc", "libraceline_b200" + os.environ.get("RL_LIB_VARIANT", "") + ".so")
_LIB = None

# every symbol include/raceline_b200.h declares
ABI_SYMBOLS = [
    "rl_abi_version", "rl_status_string", "rl_device_count", "rl_default_params", "rl_create", "rl_destroy",
    "rl_set_stream", "rl_set_option", "rl_last_error", "rl_host_alloc", "rl_host_free", "rl_job_sample_offsets", "rl_plan_for_track", "rl_solve_batch",
    "rl_batch_create", "rl_batch_upload", "rl_batch_solve", "rl_batch_download", "rl_batch_sync", "rl_batch_device_outputs",
    "rl_batch_launches_per_solve", "rl_batch_destroy", "rl_compute_min_curvature_raceline",
    "rl_compute_min_time_raceline", "rl_geom_row_offsets", "rl_centerline_geom_batch", "rl_synth_tracks", "rl_last_kernel_ms", "rl_debug_check_failures",
    "rl_measure_fp64_peak",
]


class ExtensionMissing(RuntimeError):
    pass


def lib():
    """Return the loaded shared library with argtypes set; raise if it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ExtensionMissing(
            f"{LIB_PATH} is missing: build it with `python -m practice_path_planning_for_formula_student_driverless_b200.build` "
            "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    dp = C.POINTER(C.c_double)
    vp = C.c_void_p
    L.rl_abi_version.restype = C.c_int
    L.rl_status_string.argtypes = [C.c_int]
    L.rl_status_string.restype = C.c_char_p
    L.rl_device_count.restype = C.c_int
    L.rl_default_params.argtypes = [C.POINTER(RlParams)]
    L.rl_default_params.restype = C.c_int
    L.rl_create.argtypes = [C.c_int, C.POINTER(C.c_int)]
    L.rl_create.restype = vp
    L.rl_destroy.argtypes = [vp]
    L.rl_destroy.restype = None
    L.rl_set_stream.argtypes = [vp, vp]
    L.rl_set_stream.restype = C.c_int
    L.rl_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
    L.rl_set_option.restype = C.c_int
    L.rl_last_error.argtypes = [vp]
    L.rl_last_error.restype = C.c_char_p
    L.rl_host_alloc.argtypes = [C.c_size_t]
    L.rl_host_alloc.restype = vp
    L.rl_host_free.argtypes = [vp]
    L.rl_host_free.restype = None
    L.rl_job_sample_offsets.argtypes = [C.POINTER(RlBatchDesc), C.POINTER(C.c_int64)]
    L.rl_job_sample_offsets.restype = C.c_int
    L.rl_plan_for_track.argtypes = [C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.rl_plan_for_track.restype = C.c_int
    L.rl_solve_batch.argtypes = [vp, C.POINTER(RlBatchDesc), C.POINTER(RlBatchOut)]
    L.rl_solve_batch.restype = C.c_int
    L.rl_batch_create.argtypes = [vp, C.POINTER(RlBatchDesc), C.POINTER(C.c_int)]
    L.rl_batch_create.restype = vp
    L.rl_batch_upload.argtypes = [vp, C.POINTER(RlBatchDesc)]
    L.rl_batch_upload.restype = C.c_int
    L.rl_batch_solve.argtypes = [vp]
    L.rl_batch_solve.restype = C.c_int
    L.rl_batch_download.argtypes = [vp, C.POINTER(RlBatchOut)]
    L.rl_batch_download.restype = C.c_int
    L.rl_batch_device_outputs.argtypes = [vp, C.POINTER(RlBatchOut)]
    L.rl_batch_device_outputs.restype = C.c_int
    L.rl_batch_sync.argtypes = [vp]
    L.rl_batch_sync.restype = C.c_int
    L.rl_batch_launches_per_solve.argtypes = [vp]
    L.rl_batch_launches_per_solve.restype = C.c_int
    L.rl_batch_destroy.argtypes = [vp]
    L.rl_batch_destroy.restype = None
    single = [vp, dp, C.c_int, dp, C.c_int, dp, C.c_int, C.c_double, C.c_double, C.c_int, C.POINTER(RlParams),
              dp, dp, dp, dp, dp]
    L.rl_compute_min_curvature_raceline.argtypes = single + [C.POINTER(RlJobStats)]
    L.rl_compute_min_curvature_raceline.restype = C.c_int
    L.rl_compute_min_time_raceline.argtypes = single + [dp, dp, dp, C.POINTER(RlJobStats)]
    L.rl_compute_min_time_raceline.restype = C.c_int
    L.rl_geom_row_offsets.argtypes = [C.POINTER(RlGeomDesc), C.POINTER(C.c_int64)]
    L.rl_geom_row_offsets.restype = C.c_int
    L.rl_centerline_geom_batch.argtypes = [vp, C.POINTER(RlGeomDesc), C.POINTER(RlGeomOut)]
    L.rl_centerline_geom_batch.restype = C.c_int
    L.rl_synth_tracks.argtypes = [C.c_uint64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, dp]
    L.rl_synth_tracks.restype = C.c_int
    L.rl_debug_check_failures.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.rl_debug_check_failures.restype = C.c_int
    L.rl_last_kernel_ms.argtypes = [vp]
    L.rl_last_kernel_ms.restype = C.c_double
    L.rl_measure_fp64_peak.argtypes = [vp, dp]
    L.rl_measure_fp64_peak.restype = C.c_int
    _LIB = L
    return L
