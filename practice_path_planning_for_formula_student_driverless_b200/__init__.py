"""B200-native batched raceline solver: the min-curvature and min-time stages of
tjsdn3065/Practice_path_planning_for_formula_student_driverless as hand-written sm_100a CUDA kernels
behind the C ABI of include/raceline_b200.h.  This package is the thin host-side mirror of the
reference's solver interface; importing it does not load the CUDA library (the first call does, and
fails loudly if it has not been built -- there is no CPU fallback).
"""
from ._abi import (RL_ABI_VERSION, RL_ERR_ARG, RL_ERR_CUDA, RL_ERR_NODEVICE, RL_ERR_NOMEM, RL_ERR_UNSUPPORTED,
                   RL_MAX_OUTER_LOG, RL_OK, RL_STAGE_EVAL, RL_STAGE_MINCURV, RL_STAGE_MINTIME, RlBatchDesc, RlBatchOut, RlJob,
                   RlJobStats, RlParams)
from .solver import (CenterlineGeom, Config, Context, DeviceBatch, PackedBatch, PackedGeom, PinnedPool, RacelineError, Result, Track,
                     centerline_geom_batch, compute_min_curvature_raceline, debug_compare_paths, path_length, compute_min_time_raceline, default_context,
                     polyline_edges, ring_edges, solve_batch, synth_tracks)

__all__ = [
    "RL_ABI_VERSION", "RL_OK", "RL_ERR_ARG", "RL_ERR_CUDA", "RL_ERR_UNSUPPORTED", "RL_ERR_NOMEM", "RL_ERR_NODEVICE",
    "RL_STAGE_MINCURV", "RL_STAGE_MINTIME", "RL_STAGE_EVAL", "RL_MAX_OUTER_LOG", "RlParams", "RlJob", "RlJobStats", "RlBatchDesc",
    "RlBatchOut", "Config", "Context", "DeviceBatch", "PackedBatch", "PackedGeom", "PinnedPool", "RacelineError", "Result", "Track",
    "compute_min_curvature_raceline", "compute_min_time_raceline", "default_context", "polyline_edges", "ring_edges",
    "solve_batch", "synth_tracks", "CenterlineGeom", "centerline_geom_batch", "debug_compare_paths", "path_length",
]
