"""ctypes mirrors of the POD structs in include/raceline_b200.h (no library loading here)."""
from __future__ import annotations

import ctypes as C

RL_ABI_VERSION = 2
RL_STAGE_MINCURV = 1  # compute_min_curvature_raceline, reference src/main.cpp:683
RL_STAGE_MINTIME = 2  # compute_min_time_raceline, reference src/main.cpp:905
RL_STAGE_EVAL = 3  # profile of a given path: heading/curvature + v(s) + lap (the debug block's laps, src/main.cpp:1464-1477)
RL_MAX_OUTER_LOG = 32

RL_OK = 0
RL_ERR_ARG = -1
RL_ERR_CUDA = -2
RL_ERR_UNSUPPORTED = -3
RL_ERR_NOMEM = -4
RL_ERR_NODEVICE = -5

_D = C.c_double
_I = C.c_int32


class RlParams(C.Structure):
    """rl_params: the cfg::Config fields the hot path reads (reference src/main.cpp:77-113)."""

    _fields_ = [
        ("veh_width_arg", _D), ("veh_width_m", _D), ("safety_margin_m", _D), ("lambda_smooth", _D),
        ("step_init", _D), ("step_min", _D), ("armijo_c", _D), ("kappa_eps", _D), ("v_cap_mps", _D),
        ("mass_kg", _D), ("Cd", _D), ("A_front_m2", _D), ("rho_air", _D), ("c_rr", _D), ("P_max_W", _D),
        ("a_total_max", _D), ("a_lat_max", _D), ("a_long_acc_cap", _D), ("a_long_brake_cap", _D),
        ("w_time_gain", _D), ("time_gamma_power", _D), ("inv_v_gain", _D),
        ("max_outer_iters", _I), ("max_inner_iters", _I), ("max_vpass_iters", _I),
        ("time_weight_use_inv_v", _I), ("use_total_ge_lat", _I), ("reserved", _I),
    ]


class RlJob(C.Structure):
    _fields_ = [("track", _I), ("param", _I), ("stage", _I), ("reserved", _I)]


class RlJobStats(C.Structure):
    _fields_ = [
        ("status", _I), ("n", _I), ("outer_done", _I), ("accepted", _I), ("backtracks", _I), ("evals", _I),
        ("vpass_rounds", _I), ("exist_scans", _I), ("ray_tests", C.c_int64), ("lap_time", _D),
        ("J0", _D * RL_MAX_OUTER_LOG), ("Jend", _D * RL_MAX_OUTER_LOG), ("lap_outer", _D * RL_MAX_OUTER_LOG),
        ("acc_outer", _I * RL_MAX_OUTER_LOG), ("bt_outer", _I * RL_MAX_OUTER_LOG),
    ]


class RlBatchDesc(C.Structure):
    _fields_ = [
        ("n_tracks", _I), ("n_params", _I), ("n_jobs", _I), ("reserved", _I),
        ("samp_off", C.c_void_p), ("seg_off", C.c_void_p), ("center_xy", C.c_void_p), ("seg", C.c_void_p),
        ("track_L", C.c_void_p), ("track_closed", C.c_void_p), ("params", C.c_void_p), ("jobs", C.c_void_p),
    ]


class RlBatchOut(C.Structure):
    _fields_ = [
        ("xy", C.c_void_p), ("heading", C.c_void_p), ("curvature", C.c_void_p), ("alpha_total", C.c_void_p),
        ("alpha_last", C.c_void_p), ("v", C.c_void_p), ("ax", C.c_void_p), ("stats", C.c_void_p),
    ]


class RlGeomDesc(C.Structure):
    _fields_ = [
        ("n_tracks", _I), ("emit_closed_duplicate", _I), ("mid_off", C.c_void_p), ("mids_xy", C.c_void_p),
        ("samples", C.c_void_p), ("track_closed", C.c_void_p), ("seg_off", C.c_void_p), ("seg", C.c_void_p),
        ("params", C.c_void_p),
    ]


class RlGeomOut(C.Structure):
    _fields_ = [
        ("xy", C.c_void_p), ("s_rel", C.c_void_p), ("heading", C.c_void_p), ("curvature", C.c_void_p),
        ("dist_inner", C.c_void_p), ("dist_outer", C.c_void_p), ("width", C.c_void_p), ("v_kappa", C.c_void_p),
        ("track_L", C.c_void_p), ("track_s0", C.c_void_p),
    ]


PARAM_DOUBLE_FIELDS = [n for n, t in RlParams._fields_ if t is _D]
PARAM_INT_FIELDS = [n for n, t in RlParams._fields_ if t is _I]
