"""TEST INFRASTRUCTURE: ctypes binding of oracle/liboracle.so (the C restatement of the reference hot path).

Only tests/, bench.py's cpu_baseline leg and __graft_entry__.smoke() may import this module, and
only as the checker.  The product package never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from practice_path_planning_for_formula_student_driverless_b200._abi import (
    RL_STAGE_MINCURV,
    RL_STAGE_MINTIME,
    RlJobStats,
    RlParams,
)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    """Compile liboracle.so (and the _ref/ binaries when /root/reference is present)."""
    so = os.path.join(_HERE, "liboracle.so")
    syn = os.path.join(_HERE, "libsynth_tracks.so")
    if force or not os.path.exists(so) or not os.path.exists(syn) or not os.path.exists(os.path.join(_HERE, "liboracle_fma.so")) or os.path.getmtime(so) < max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("raceline_oracle.c", "geom_oracle.c", "raceline_oracle.h")):
        subprocess.run(["make", "-C", _HERE, "port"], check=True, capture_output=True)
    return so


_SYNTH = None


def synth_tracks(n_tracks, n_samples, m_per_ring, seed_base=0xB200, first_id=0, threads=0):
    """The bench workload generator (csrc/synth_tracks.cpp) through oracle/libsynth_tracks.so: the reference arm of
    bench.py makes its inputs without loading the CUDA library.  Returns (center_xy, seg, L)."""
    global _SYNTH
    if _SYNTH is None:
        build()
        _SYNTH = C.CDLL(os.path.join(_HERE, "libsynth_tracks.so"))
        dp = C.POINTER(C.c_double)
        _SYNTH.rl_synth_tracks.argtypes = [C.c_uint64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, dp]
        _SYNTH.rl_synth_tracks.restype = C.c_int
    center = np.empty((n_tracks * n_samples, 2))
    seg = np.empty((n_tracks * 2 * m_per_ring, 4))
    L = np.empty(n_tracks)
    dp = C.POINTER(C.c_double)
    rc = _SYNTH.rl_synth_tracks(C.c_uint64(seed_base), C.c_int64(first_id), n_tracks, n_samples, m_per_ring, threads,
                                center.ctypes.data_as(dp), seg.ctypes.data_as(dp), L.ctypes.data_as(dp))
    if rc != 0:
        raise RuntimeError(f"rl_synth_tracks failed ({rc})")
    return center, seg, L


def build_ref():
    """Build oracle/_ref/* from /root/reference if it is mounted; prebuilt files are used otherwise."""
    if os.path.exists("/root/reference/src/main.cpp"):
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, capture_output=True)
        # the drop-in demo binary links against the CUDA library, which must have been built first
        subprocess.run(["make", "-C", _HERE, "dropin"], check=False, capture_output=True)
    return os.path.join(_HERE, "_ref")


def ref_binary(name="ref_harness"):
    p = os.path.join(_HERE, "_ref", name)
    return p if os.path.exists(p) else None


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        dp = C.POINTER(C.c_double)
        L.orc_default_params.argtypes = [C.POINTER(RlParams)]
        L.orc_default_params.restype = None
        L.orc_solve.argtypes = [C.c_int, dp, C.c_int, dp, C.c_int, dp, C.c_int, C.c_double, C.c_int,
                                C.POINTER(RlParams), dp, dp, dp, dp, dp, dp, dp, C.POINTER(RlJobStats)]
        L.orc_solve.restype = C.c_int
        L.orc_eval_cost_grad.argtypes = [dp, dp, dp, dp, dp, C.c_double, C.c_double, dp, C.c_int, C.c_int, dp, dp]
        L.orc_eval_cost_grad.restype = C.c_double
        L.orc_velocity_profile.argtypes = [C.POINTER(RlParams), dp, C.c_int, C.c_double, C.c_int, dp, dp]
        L.orc_velocity_profile.restype = C.c_double
        L.orc_geom_rows.argtypes = [C.c_int, C.c_int, C.c_int]
        L.orc_geom_rows.restype = C.c_int
        L.orc_centerline_geom.argtypes = [dp, C.c_int, C.c_int, C.c_int, C.c_int, dp, C.c_int, dp, C.c_int, C.POINTER(RlParams),
                                          dp, dp, dp, dp, dp, dp, dp, dp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_centerline_geom.restype = C.c_int
        L.orc_time_weights.argtypes = [C.POINTER(RlParams), dp, dp, C.c_int, dp]
        L.orc_time_weights.restype = None
        L.orc_corridor.argtypes = [dp, dp, C.c_int, dp, C.c_int, dp, C.c_int, C.c_double, dp, dp]
        L.orc_corridor.restype = None
        L.orc_normals.argtypes = [dp, C.c_int, C.c_int, dp]
        L.orc_normals.restype = None
        L.orc_heading_curv.argtypes = [dp, C.c_int, C.c_double, C.c_int, dp, dp]
        L.orc_heading_curv.restype = None
        L.orc_lin_geom.argtypes = [dp, dp, C.c_int, C.c_double, C.c_int, dp, dp, dp, dp]
        L.orc_lin_geom.restype = None
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def default_params() -> RlParams:
    p = RlParams()
    lib().orc_default_params(C.byref(p))
    return p


_FMA = None


def _fma_lib():
    """liboracle_fma.so: the same C restatement compiled with FMA contraction (a rounding-sensitivity yardstick)."""
    global _FMA
    if _FMA is None:
        build()
        F = C.CDLL(os.path.join(_HERE, "liboracle_fma.so"))
        F.orc_solve.argtypes = lib().orc_solve.argtypes
        F.orc_solve.restype = C.c_int
        _FMA = F
    return _FMA


def solve(stage, center_xy, inner_seg, outer_seg, L, closed=True, params=None, rounding="reference"):
    """Run one stage on the CPU.  Returns a dict with the reference Result fields + 'stats'.
    rounding="fma": the FMA-contracted build (NOT the reference's bits; see _fma_lib)."""
    center_xy = _c(center_xy).reshape(-1, 2)
    inner_seg = _c(inner_seg).reshape(-1, 4)
    outer_seg = _c(outer_seg).reshape(-1, 4)
    n = center_xy.shape[0]
    params = params if params is not None else default_params()
    out = {k: np.zeros(n) for k in ("heading", "curvature", "alpha_total", "alpha_last", "v", "ax")}
    out["xy"] = np.zeros((n, 2))
    st = RlJobStats()
    rc = (lib() if rounding == "reference" else _fma_lib()).orc_solve(int(stage), _p(center_xy), n, _p(inner_seg), inner_seg.shape[0], _p(outer_seg),
                         outer_seg.shape[0], float(L), int(bool(closed)), C.byref(params), _p(out["xy"]),
                         _p(out["heading"]), _p(out["curvature"]), _p(out["alpha_total"]), _p(out["alpha_last"]),
                         _p(out["v"]), _p(out["ax"]), C.byref(st))
    if rc != 0:
        raise RuntimeError(f"orc_solve failed: {rc}")
    out["stats"] = st
    out["lap_time"] = st.lap_time
    return out


def centerline_geom(mids_xy, samples, inner_seg, outer_seg, closed=True, emit_closed_duplicate=True, params=None):
    """The stage before the path on the CPU (geom_oracle.c): centre line + width/geometry rows of one track."""
    mids_xy = _c(mids_xy).reshape(-1, 2)
    inner_seg = _c(inner_seg).reshape(-1, 4)
    outer_seg = _c(outer_seg).reshape(-1, 4)
    params = params if params is not None else default_params()
    L = lib()
    rows = L.orc_geom_rows(int(samples), int(bool(closed)), int(bool(emit_closed_duplicate)))
    out = {k: np.zeros(rows) for k in ("s_rel", "heading", "curvature", "dist_inner", "dist_outer", "width", "v_kappa")}
    out["xy"] = np.zeros((rows, 2))
    Lv, s0 = C.c_double(0.0), C.c_double(0.0)
    rc = L.orc_centerline_geom(_p(mids_xy), mids_xy.shape[0], int(samples), int(bool(closed)), int(bool(emit_closed_duplicate)),
                               _p(inner_seg), inner_seg.shape[0], _p(outer_seg), outer_seg.shape[0], C.byref(params),
                               _p(out["xy"]), _p(out["s_rel"]), _p(out["heading"]), _p(out["curvature"]), _p(out["dist_inner"]),
                               _p(out["dist_outer"]), _p(out["width"]), _p(out["v_kappa"]), C.byref(Lv), C.byref(s0))
    if rc != rows:
        raise RuntimeError(f"orc_centerline_geom failed: {rc}")
    out["L"], out["s0"] = Lv.value, s0.value
    return out


def eval_cost_grad(A1, A2, N0, W, gamma2, h, lam, alpha, closed=True):
    A1, A2, N0, W, alpha = map(_c, (A1, A2, N0, W, alpha))
    g2 = _c(gamma2) if gamma2 is not None else None
    n = alpha.size
    grad = np.zeros(n)
    work = np.zeros(8 * n)
    J = lib().orc_eval_cost_grad(_p(A1), _p(A2), _p(N0), _p(W), _p(g2), float(h), float(lam), _p(alpha), n,
                                 int(bool(closed)), _p(grad), _p(work))
    return J, grad


def velocity_profile(params, kappa, h, closed=True):
    kappa = _c(kappa)
    n = kappa.size
    v, ax = np.zeros(n), np.zeros(n)
    lap = lib().orc_velocity_profile(C.byref(params), _p(kappa), n, float(h), int(bool(closed)), _p(v), _p(ax))
    return v, ax, lap


def time_weights(params, kappa, v):
    kappa, v = _c(kappa), _c(v)
    g2 = np.zeros(kappa.size)
    lib().orc_time_weights(C.byref(params), _p(kappa), _p(v), kappa.size, _p(g2))
    return g2


def normals(pxy, closed=True):
    pxy = _c(pxy).reshape(-1, 2)
    out = np.zeros_like(pxy)
    lib().orc_normals(_p(pxy), pxy.shape[0], int(bool(closed)), _p(out))
    return out


def heading_curv(pxy, h, closed=True):
    pxy = _c(pxy).reshape(-1, 2)
    hd, kp = np.zeros(pxy.shape[0]), np.zeros(pxy.shape[0])
    lib().orc_heading_curv(_p(pxy), pxy.shape[0], float(h), int(bool(closed)), _p(hd), _p(kp))
    return hd, kp


def lin_geom(pxy, nxy, h, closed=True):
    pxy, nxy = _c(pxy).reshape(-1, 2), _c(nxy).reshape(-1, 2)
    n = pxy.shape[0]
    A1, A2, N0, W = (np.zeros(n) for _ in range(4))
    lib().orc_lin_geom(_p(pxy), _p(nxy), n, float(h), int(bool(closed)), _p(A1), _p(A2), _p(N0), _p(W))
    return A1, A2, N0, W


def corridor(pxy, nxy, inner_seg, outer_seg, guard):
    pxy, nxy = _c(pxy).reshape(-1, 2), _c(nxy).reshape(-1, 2)
    inner_seg, outer_seg = _c(inner_seg).reshape(-1, 4), _c(outer_seg).reshape(-1, 4)
    n = pxy.shape[0]
    lo, hi = np.zeros(n), np.zeros(n)
    lib().orc_corridor(_p(pxy), _p(nxy), n, _p(inner_seg), inner_seg.shape[0], _p(outer_seg), outer_seg.shape[0],
                       float(guard), _p(lo), _p(hi))
    return lo, hi


__all__ = ["RL_STAGE_MINCURV", "RL_STAGE_MINTIME", "build", "build_ref", "ref_binary", "default_params", "solve",
           "eval_cost_grad", "velocity_profile", "time_weights", "normals", "heading_curv", "lin_geom", "corridor"]
