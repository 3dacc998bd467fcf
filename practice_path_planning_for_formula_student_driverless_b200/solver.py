"""Host-side mirror of the reference's solver interface, over the C ABI (include/raceline_b200.h).

Reference boundary (tjsdn3065/Practice_path_planning_for_formula_student_driverless, src/main.cpp):

    raceline_min_curv::compute_min_curvature_raceline(center, innerE, outerE, veh_width, L, closed) -> Result   :683
    raceline_min_time::compute_min_time_raceline(center, innerE, outerE, veh_width, L, closed)      -> Result   :905
    cfg::Config                                                                                                :47-119
    edges::ringEdges / edges::polylineEdges                                                                     :251 / :256

Same names, same argument meaning, same Result fields; Config is passed explicitly instead of the
process-global cfg::get().  The batched form (`solve_batch`) is the data-parallel version of the two
calls at main.cpp:1347 and main.cpp:1397.  Everything here runs through the CUDA library; there is no
CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import weakref
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from ._abi import (RL_OK, RL_STAGE_EVAL, RL_STAGE_MINCURV, RL_STAGE_MINTIME, RlBatchDesc, RlBatchOut, RlJob, RlJobStats,
                   RlParams)
from ._lib import lib


class RacelineError(RuntimeError):
    def __init__(self, status, detail=""):
        self.status = status
        msg = lib().rl_status_string(status).decode()
        super().__init__(f"raceline_b200: {msg} ({status}){': ' + detail if detail else ''}")


# ------------------------------------------------------------------------------------------------
# cfg::Config (main.cpp:47-119): the fields the min-curv / min-time stages read
# ------------------------------------------------------------------------------------------------
@dataclass
class Config:
    is_closed_track: bool = True          # main.cpp:54
    veh_width_m: float = 1.0              # :77
    safety_margin_m: float = 0.05         # :78
    lambda_smooth: float = 1.6e-3         # :81
    max_outer_iters: int = 14             # :82
    max_inner_iters: int = 120            # :83
    step_init: float = 0.65               # :84
    step_min: float = 1e-6                # :85
    armijo_c: float = 1e-5                # :86
    kappa_eps: float = 1e-6               # :89
    v_cap_mps: float = 27.0               # :90
    mass_kg: float = 255.0                # :93
    Cd: float = 0.30                      # :94
    A_front_m2: float = 1.00              # :95
    rho_air: float = 1.225                # :96
    c_rr: float = 0.015                   # :97
    P_max_W: float = 80000.0              # :98
    mu: float = 1.17                      # :101
    a_total_max: Optional[float] = None   # :102  mu*9.81, evaluated once at construction like the reference
    a_lat_max: float = 11.0               # :103
    a_long_acc_cap: float = 8.0           # :104
    a_long_brake_cap: float = 11.0        # :105
    w_time_gain: float = 1.0              # :108
    time_gamma_power: float = 2.0         # :109
    time_weight_use_inv_v: bool = False   # :110
    inv_v_gain: float = 0.1               # :111
    max_vpass_iters: int = 6              # :112
    use_total_ge_lat: bool = True         # :113

    def __post_init__(self):
        if self.a_total_max is None:
            self.a_total_max = self.mu * 9.81

    def to_params(self, veh_width: Optional[float] = None) -> RlParams:
        p = RlParams()
        p.veh_width_arg = self.veh_width_m if veh_width is None else float(veh_width)
        for name in ("veh_width_m", "safety_margin_m", "lambda_smooth", "step_init", "step_min", "armijo_c", "kappa_eps",
                     "v_cap_mps", "mass_kg", "Cd", "A_front_m2", "rho_air", "c_rr", "P_max_W", "a_total_max", "a_lat_max",
                     "a_long_acc_cap", "a_long_brake_cap", "w_time_gain", "time_gamma_power", "inv_v_gain"):
            setattr(p, name, float(getattr(self, name)))
        p.max_outer_iters = int(self.max_outer_iters)
        p.max_inner_iters = int(self.max_inner_iters)
        p.max_vpass_iters = int(self.max_vpass_iters)
        p.time_weight_use_inv_v = int(bool(self.time_weight_use_inv_v))
        p.use_total_ge_lat = int(bool(self.use_total_ge_lat))
        return p


def ring_edges(R) -> np.ndarray:
    """edges::ringEdges (main.cpp:251): closed ring of cones -> (M,4) segments x0,y0,x1,y1."""
    R = np.ascontiguousarray(R, dtype=np.float64).reshape(-1, 2)
    return np.concatenate([R, np.roll(R, -1, axis=0)], axis=1) if len(R) else np.zeros((0, 4))


def polyline_edges(R) -> np.ndarray:
    """edges::polylineEdges (main.cpp:256): open cone chain -> (M-1,4) segments."""
    R = np.ascontiguousarray(R, dtype=np.float64).reshape(-1, 2)
    return np.concatenate([R[:-1], R[1:]], axis=1) if len(R) >= 2 else np.zeros((0, 4))


@dataclass
class Track:
    """The solver arguments that describe one track (main.cpp:683-686)."""
    center_xy: np.ndarray      # (N,2) centre samples, closing duplicate dropped (main.cpp:1681-1683)
    inner_seg: np.ndarray      # (M_in,4)  innerE
    outer_seg: np.ndarray      # (M_out,4) outerE
    L: float                   # CL.L
    closed: bool = True


@dataclass
class Result:
    """raceline_min_curv::Result (main.cpp:677-681) / raceline_min_time::Result (main.cpp:897-903)."""
    raceline: np.ndarray
    heading: np.ndarray
    curvature: np.ndarray
    alpha_total: np.ndarray
    alpha_last: np.ndarray
    v: Optional[np.ndarray] = None
    ax: Optional[np.ndarray] = None
    lap_time: float = 0.0
    stats: Optional[RlJobStats] = None


def _f64(a, cols=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a.reshape(-1, cols) if cols else a


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


class PinnedPool:
    """Page-locked numpy arrays (rl_host_alloc) so that H2D/D2H copies are asynchronous.

    The arrays are views of memory the pool owns: `close()` frees it, so it refuses to run while an array handed out
    by `empty()` is still referenced (drop the arrays -- and the PackedBatch built on them -- first, or pass
    force=True at interpreter teardown)."""

    def __init__(self):
        self._ptrs = []
        self._live = []   # weak references to the base arrays handed out

    def empty(self, shape, dtype):
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) if np.ndim(shape) else int(shape)
        nbytes = max(1, n * dtype.itemsize)
        p = lib().rl_host_alloc(nbytes)
        if not p:
            raise MemoryError("rl_host_alloc failed")
        self._ptrs.append(p)
        buf = (C.c_char * nbytes).from_address(p)
        base = np.frombuffer(buf, dtype=dtype, count=n)
        self._live.append(weakref.ref(base))
        return base.reshape(shape)

    def copy(self, a):
        out = self.empty(a.shape, a.dtype)
        out[...] = a
        return out

    def live_arrays(self) -> int:
        return sum(1 for r in self._live if r() is not None)

    def close(self, force=False):
        if not force and self.live_arrays():
            raise RuntimeError(f"PinnedPool.close(): {self.live_arrays()} array(s) of this pool are still referenced; "
                               "their memory would be freed under them")
        for p in self._ptrs:
            lib().rl_host_free(p)
        self._ptrs, self._live = [], []


class PackedBatch:
    """Host arrays in the rl_batch_desc layout + the matching output arrays."""

    def __init__(self, tracks: Sequence[Track], params: Sequence[RlParams], jobs, pool: Optional[PinnedPool] = None):
        alloc = (lambda shape, dt: pool.empty(shape, dt)) if pool else (lambda shape, dt: np.empty(shape, dtype=dt))
        self.n_tracks, self.n_params = len(tracks), len(params)
        ns = [int(_f64(t.center_xy, 2).shape[0]) for t in tracks]
        mi = [int(_f64(t.inner_seg, 4).shape[0]) for t in tracks]
        mo = [int(_f64(t.outer_seg, 4).shape[0]) for t in tracks]
        self.samp_off = alloc(self.n_tracks + 1, np.int64)
        self.samp_off[0] = 0
        self.samp_off[1:] = np.cumsum(ns)
        self.seg_off = alloc(2 * self.n_tracks + 1, np.int64)
        self.seg_off[0] = 0
        inter = np.empty(2 * self.n_tracks, dtype=np.int64)
        inter[0::2], inter[1::2] = mi, mo
        self.seg_off[1:] = np.cumsum(inter)
        self.center_xy = alloc((int(self.samp_off[-1]), 2), np.float64)
        self.seg = alloc((int(self.seg_off[-1]), 4), np.float64)
        self.track_L = alloc(self.n_tracks, np.float64)
        self.track_closed = alloc(self.n_tracks, np.int32)
        for t, tr in enumerate(tracks):
            self.center_xy[self.samp_off[t]:self.samp_off[t + 1]] = _f64(tr.center_xy, 2)
            self.seg[self.seg_off[2 * t]:self.seg_off[2 * t + 1]] = _f64(tr.inner_seg, 4)
            self.seg[self.seg_off[2 * t + 1]:self.seg_off[2 * t + 2]] = _f64(tr.outer_seg, 4)
            self.track_L[t] = float(tr.L)
            self.track_closed[t] = int(bool(tr.closed))
        self._finish(params, jobs, alloc, pool)

    @classmethod
    def from_arrays(cls, samp_off, seg_off, center_xy, seg, track_L, track_closed, params, jobs, pool=None):
        """Wrap arrays that are already in the packed layout (no per-track copies)."""
        self = cls.__new__(cls)
        alloc = (lambda shape, dt: pool.empty(shape, dt)) if pool else (lambda shape, dt: np.empty(shape, dtype=dt))
        self.n_tracks, self.n_params = len(track_L), len(params)
        self.samp_off = np.ascontiguousarray(samp_off, dtype=np.int64)
        self.seg_off = np.ascontiguousarray(seg_off, dtype=np.int64)
        self.center_xy = _f64(center_xy, 2)
        self.seg = _f64(seg, 4)
        self.track_L = _f64(track_L)
        self.track_closed = np.ascontiguousarray(track_closed, dtype=np.int32)
        self._finish(params, jobs, alloc, pool)
        return self

    def _finish(self, params, jobs, alloc, pool=None):
        jobs = np.asarray(jobs, dtype=np.int64).reshape(-1, 3)
        self.n_jobs = jobs.shape[0]
        self.params = (RlParams * max(1, self.n_params))(*params)
        self.jobs_np = jobs
        self.jobs = (RlJob * max(1, self.n_jobs))()
        for j, (t, p, s) in enumerate(jobs):
            self.jobs[j].track, self.jobs[j].param, self.jobs[j].stage = int(t), int(p), int(s)
        self.desc = RlBatchDesc()
        self.desc.n_tracks, self.desc.n_params, self.desc.n_jobs = self.n_tracks, self.n_params, self.n_jobs
        self.desc.samp_off = _ptr(self.samp_off) or self.samp_off.ctypes.data
        self.desc.seg_off = _ptr(self.seg_off) or self.seg_off.ctypes.data
        self.desc.center_xy = _ptr(self.center_xy)
        self.desc.seg = _ptr(self.seg)
        self.desc.track_L = _ptr(self.track_L)
        self.desc.track_closed = _ptr(self.track_closed)
        self.desc.params = C.cast(self.params, C.c_void_p)
        self.desc.jobs = C.cast(self.jobs, C.c_void_p)
        ns = np.diff(self.samp_off)
        self.job_off = np.zeros(self.n_jobs + 1, dtype=np.int64)
        if self.n_jobs:
            ok = (jobs[:, 0] >= 0) & (jobs[:, 0] < self.n_tracks)      # bad indices are reported by the C ABI
            self.job_off[1:] = np.cumsum(np.where(ok, ns[np.clip(jobs[:, 0], 0, max(0, self.n_tracks - 1))], 0))
        rows = int(self.job_off[-1])
        self.rows = rows
        self.out_xy = alloc((rows, 2), np.float64)
        self.out_heading = alloc(rows, np.float64)
        self.out_curvature = alloc(rows, np.float64)
        self.out_alpha_total = alloc(rows, np.float64)
        self.out_alpha_last = alloc(rows, np.float64)
        self.out_v = alloc(rows, np.float64)
        self.out_ax = alloc(rows, np.float64)
        if pool is not None:      # page-locked, so that the library copies the counters straight into it
            raw = pool.empty(max(1, self.n_jobs) * C.sizeof(RlJobStats), np.uint8)
            raw[:] = 0
            self._stats_raw = raw
            self.out_stats = (RlJobStats * max(1, self.n_jobs)).from_buffer(raw)
        else:
            self.out_stats = (RlJobStats * max(1, self.n_jobs))()
        self.out = RlBatchOut()
        self.out.xy, self.out.heading, self.out.curvature = _ptr(self.out_xy), _ptr(self.out_heading), _ptr(self.out_curvature)
        self.out.alpha_total, self.out.alpha_last = _ptr(self.out_alpha_total), _ptr(self.out_alpha_last)
        self.out.v, self.out.ax = _ptr(self.out_v), _ptr(self.out_ax)
        self.out.stats = C.cast(self.out_stats, C.c_void_p)

    @property
    def h2d_bytes(self):
        return int(self.samp_off.nbytes + self.seg_off.nbytes + self.center_xy.nbytes + self.seg.nbytes +
                   self.track_L.nbytes + self.track_closed.nbytes + C.sizeof(RlParams) * self.n_params +
                   C.sizeof(RlJob) * self.n_jobs)

    @property
    def d2h_bytes(self):
        return int(self.rows * 8 * 8 + C.sizeof(RlJobStats) * self.n_jobs)

    def result(self, j) -> Result:
        a, b = int(self.job_off[j]), int(self.job_off[j + 1])
        st = self.out_stats[j]
        mt = int(self.jobs_np[j, 2]) in (RL_STAGE_MINTIME, RL_STAGE_EVAL)
        return Result(self.out_xy[a:b], self.out_heading[a:b], self.out_curvature[a:b], self.out_alpha_total[a:b],
                      self.out_alpha_last[a:b], self.out_v[a:b] if mt else None, self.out_ax[a:b] if mt else None,
                      float(st.lap_time), st)


class Context:
    """One context per device: replaces the reference's process-global cfg::get() state (main.cpp:120)."""

    def __init__(self, device: int = 0):
        st = C.c_int(0)
        self._h = lib().rl_create(int(device), C.byref(st))
        if not self._h:
            raise RacelineError(st.value, "rl_create")
        self.device = device
        self._batches = weakref.WeakSet()   # live DeviceBatch objects: closed with (before) the context

    def set_stream(self, cuda_stream: Optional[int]):
        lib().rl_set_stream(self._h, C.c_void_p(cuda_stream) if cuda_stream else None)

    def set_option(self, name: str, value: int):
        """rl_set_option: tuning knobs / test hooks of the host plan ("solve_chunks", "chunk_streams", "geom_chunks",
        "max_chain", "force_chain", "force_cluster", "no_few_search", "debug_inject"; 0 = automatic; see
        include/raceline_b200.h)."""
        self._check(lib().rl_set_option(self._h, name.encode(), int(value)), "rl_set_option")

    def _check(self, st, what=""):
        if st != RL_OK:
            raise RacelineError(st, (what + " " + lib().rl_last_error(self._h).decode()).strip())

    def solve_batch(self, batch: PackedBatch) -> PackedBatch:
        """Host buffers in, host buffers out (rl_solve_batch)."""
        self._check(lib().rl_solve_batch(self._h, C.byref(batch.desc), C.byref(batch.out)), "rl_solve_batch")
        return batch

    def debug_check_failures(self):
        """(count, first failure code, CTA of the first failure) of the -DRL_DEBUG_CHECKS build; raises on the product
        build (RL_ERR_UNSUPPORTED)."""
        out = (C.c_uint64 * 2)()
        self._check(lib().rl_debug_check_failures(self._h, out), "rl_debug_check_failures")
        return int(out[0]), int(out[1] & 0xffffffff), int(out[1] >> 32)

    def last_kernel_ms(self) -> float:
        """Device time of the kernels of the last centerline_geom_batch call on this context (copies excluded)."""
        return float(lib().rl_last_kernel_ms(self._h))

    def fp64_peak_tflops(self) -> float:
        v = C.c_double(0)
        self._check(lib().rl_measure_fp64_peak(self._h, C.byref(v)))
        return v.value

    def close(self):
        if self._h:
            for b in list(self._batches):
                b.close()
            lib().rl_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _CudaView:
    """A device array described through __cuda_array_interface__ (what torch.as_tensor needs to wrap it in place)."""

    def __init__(self, ptr, shape, strides=None, typestr="<f8", owner=None):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3,
                                         "strides": tuple(strides) if strides else None}
        self._owner = owner


class DeviceBatch:
    """A device-resident batch (rl_batch_*): upload once, solve many times."""

    def __init__(self, ctx: Context, batch: PackedBatch):
        st = C.c_int(0)
        self.ctx, self.host = ctx, batch
        self._h = lib().rl_batch_create(ctx._h, C.byref(batch.desc), C.byref(st))
        if not self._h:
            ctx._check(st.value, "rl_batch_create")
        ctx._batches.add(self)

    def upload(self):
        self.ctx._check(lib().rl_batch_upload(self._h, C.byref(self.host.desc)), "rl_batch_upload")

    def solve(self):
        self.ctx._check(lib().rl_batch_solve(self._h), "rl_batch_solve")

    def download(self):
        self.ctx._check(lib().rl_batch_download(self._h, C.byref(self.host.out)), "rl_batch_download")

    def sync(self):
        self.ctx._check(lib().rl_batch_sync(self._h), "rl_batch_sync")

    @property
    def launches_per_solve(self):
        return lib().rl_batch_launches_per_solve(self._h)

    def device_tensor(self, name: str):
        """The batch's DEVICE output array `name` as a torch tensor that shares its memory (no copy): "xy" (rows, 2),
        "heading" / "curvature" / "alpha_total" / "alpha_last" / "v" / "ax" (rows,), or "lap_time" (n_jobs,) -- a strided
        view of the per-job counters.  For device-side consumers such as the final gather over NCCL (sharding.py)."""
        import torch
        dev = RlBatchOut()
        self.ctx._check(lib().rl_batch_device_outputs(self._h, C.byref(dev)), "rl_batch_device_outputs")
        rows, nj = self.host.rows, self.host.n_jobs
        if name == "lap_time":
            view = _CudaView(dev.stats + RlJobStats.lap_time.offset, (nj,), (C.sizeof(RlJobStats),), owner=self)
        elif name == "xy":
            view = _CudaView(dev.xy, (rows, 2), owner=self)
        elif name in ("heading", "curvature", "alpha_total", "alpha_last", "v", "ax"):
            view = _CudaView(getattr(dev, name), (rows,), owner=self)
        else:
            raise KeyError(name)
        return torch.as_tensor(view, device=torch.device("cuda", self.ctx.device))

    def close(self):
        if self._h:
            lib().rl_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_DEFAULT_CTX: List[Optional[Context]] = [None]


def default_context() -> Context:
    if _DEFAULT_CTX[0] is None:
        _DEFAULT_CTX[0] = Context(0)
    return _DEFAULT_CTX[0]


def _single(stage, center, innerE, outerE, veh_width, L, closed, cfg, ctx):
    cfg = cfg or Config()
    ctx = ctx or default_context()
    center = _f64(center, 2)
    innerE, outerE = _f64(innerE, 4), _f64(outerE, 4)
    n = center.shape[0]
    p = cfg.to_params(veh_width)
    dp = C.POINTER(C.c_double)
    xy, hd, kp, at, al = np.zeros((n, 2)), np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n)
    st = RlJobStats()

    def P(a):
        return a.ctypes.data_as(dp)

    common = [ctx._h, P(center), n, P(innerE), innerE.shape[0], P(outerE), outerE.shape[0], float(veh_width), float(L),
              int(bool(closed)), C.byref(p), P(xy), P(hd), P(kp), P(at), P(al)]
    if stage == RL_STAGE_MINCURV:
        rc = lib().rl_compute_min_curvature_raceline(*common, C.byref(st))
        ctx._check(rc, "rl_compute_min_curvature_raceline")
        return Result(xy, hd, kp, at, al, stats=st)
    v, ax, lap = np.zeros(n), np.zeros(n), C.c_double(0)
    rc = lib().rl_compute_min_time_raceline(*common, P(v), P(ax), C.byref(lap), C.byref(st))
    ctx._check(rc, "rl_compute_min_time_raceline")
    return Result(xy, hd, kp, at, al, v, ax, lap.value, st)


def compute_min_curvature_raceline(center, innerE, outerE, veh_width, L, closed, cfg: Optional[Config] = None,
                                   ctx: Optional[Context] = None) -> Result:
    """raceline_min_curv::compute_min_curvature_raceline (main.cpp:683-764) on the GPU."""
    return _single(RL_STAGE_MINCURV, center, innerE, outerE, veh_width, L, closed, cfg, ctx)


def compute_min_time_raceline(center, innerE, outerE, veh_width, L, closed, cfg: Optional[Config] = None,
                              ctx: Optional[Context] = None) -> Result:
    """raceline_min_time::compute_min_time_raceline (main.cpp:905-1052) on the GPU."""
    return _single(RL_STAGE_MINTIME, center, innerE, outerE, veh_width, L, closed, cfg, ctx)


def solve_batch(tracks: Sequence[Track], configs: Sequence[Config], jobs, ctx: Optional[Context] = None,
                veh_width: Optional[float] = None) -> List[Result]:
    """Solve jobs = [(track index, config index, stage), ...] in one batched call."""
    ctx = ctx or default_context()
    pb = PackedBatch(tracks, [c.to_params(veh_width) for c in configs], jobs)
    ctx.solve_batch(pb)
    return [pb.result(j) for j in range(pb.n_jobs)]


def path_length(P, closed=True) -> float:
    """The `path_length` lambda of the reference's debug block (main.cpp:1445-1450)."""
    P = _f64(P, 2)
    if P.shape[0] <= 1:
        return 0.0
    d = np.hypot(P[1:, 0] - P[:-1, 0], P[1:, 1] - P[:-1, 1])
    tot = 0.0
    for v in d:                      # left-to-right sum like the reference
        tot += float(v)
    if closed and P.shape[0] >= 2:
        tot += float(np.hypot(P[0, 0] - P[-1, 0], P[0, 1] - P[-1, 1]))
    return tot


DEBUG_COLUMNS = ("s", "cx", "cy", "mt_x", "mt_y", "mc_x", "mc_y", "d_mt_signed_m", "d_mc_signed_m", "d_mt_abs_m", "d_mc_abs_m",
                 "kappa_mt", "v_mt", "ax_mt", "alat_mt", "alat_ratio", "gamma", "a_acc_cap", "a_brk_cap", "a_power_cap")


def debug_compare_paths(center, mt: "Result", mc_raceline, L, s0=0.0, closed=True, cfg: Optional["Config"] = None,
                        ctx: Optional["Context"] = None):
    """The reference's debug comparison (pipeline::compute_mintime_and_save, main.cpp:1440-1593): the lap of the centre
    line and of the min-curvature path under the same dynamics (two RL_STAGE_EVAL jobs on the GPU, main.cpp:1464-1477)
    and the 20 columns of <base>_debug_compare_paths.csv (main.cpp:1493-1495) with their summary statistics.
    `mc_raceline` is what the reference re-reads from <base>_raceline.csv (closing duplicate dropped), or None."""
    cfg = cfg or Config()
    center = _f64(center, 2)
    N = min(mt.raceline.shape[0], center.shape[0])
    tracks = [Track(center, np.zeros((0, 4)), np.zeros((0, 4)), L, closed)]
    jobs = [(0, 0, RL_STAGE_EVAL)]
    mc = None
    if mc_raceline is not None and len(mc_raceline):
        mc = _f64(mc_raceline, 2)
        tracks.append(Track(mc, np.zeros((0, 4)), np.zeros((0, 4)), path_length(mc, closed), closed))   # L_mc, Hmc (1472-1473)
        jobs.append((1, 0, RL_STAGE_EVAL))
    ev = solve_batch(tracks, [cfg], jobs, ctx=ctx)
    out = {"lap_center": ev[0].lap_time, "lap_mincurv": ev[1].lap_time if mc is not None else -1.0, "lap_mintime": mt.lap_time}
    # centre-line normals (normals_from_points_generic, main.cpp:581-593)
    if closed:
        t = (np.roll(center, -1, axis=0) - np.roll(center, 1, axis=0)) * 0.5
    else:
        t = np.empty_like(center)
        t[1:-1] = (center[2:] - center[:-2]) * 0.5
        t[0], t[-1] = center[1] - center[0], center[-1] - center[-2]
    tn = np.hypot(t[:, 0], t[:, 1])
    t[tn < 1e-15] = (1.0, 0.0)
    nv = np.stack([-t[:, 1], t[:, 0]], axis=1)
    ln = np.sqrt(nv[:, 0] * nv[:, 0] + nv[:, 1] * nv[:, 1])
    nc = np.where((ln < 1e-15)[:, None], 0.0, nv / np.maximum(ln, 1e-300)[:, None])
    k = np.arange(N)
    col = {"s": (s0 + L * (k / float(max(1, N)))) - s0, "cx": center[:N, 0], "cy": center[:N, 1],
           "mt_x": mt.raceline[:N, 0], "mt_y": mt.raceline[:N, 1]}
    mcx = np.full(N, np.nan)
    mcy = np.full(N, np.nan)
    if mc is not None:
        m = min(N, mc.shape[0])
        mcx[:m], mcy[:m] = mc[:m, 0], mc[:m, 1]
    col["mc_x"], col["mc_y"] = mcx, mcy
    col["d_mt_signed_m"] = (col["mt_x"] - col["cx"]) * nc[:N, 0] + (col["mt_y"] - col["cy"]) * nc[:N, 1]
    col["d_mc_signed_m"] = (mcx - col["cx"]) * nc[:N, 0] + (mcy - col["cy"]) * nc[:N, 1]
    col["d_mt_abs_m"], col["d_mc_abs_m"] = np.abs(col["d_mt_signed_m"]), np.abs(col["d_mc_signed_m"])
    kap, v, ax = mt.curvature[:N], mt.v[:N], mt.ax[:N]
    alat = v * v * np.abs(kap)
    col.update({"kappa_mt": kap, "v_mt": v, "ax_mt": ax, "alat_mt": alat,
                "alat_ratio": np.minimum(1.0, alat / cfg.a_total_max) if cfg.a_total_max > 1e-9 else np.zeros(N)})
    vk = np.sqrt(cfg.a_lat_max / np.maximum(np.abs(kap), cfg.kappa_eps))
    col["gamma"] = 1.0 + cfg.w_time_gain * np.clip((vk - v) / np.maximum(1e-6, vk), 0.0, 1.0)       # main.cpp:1525-1527
    a_res = np.sqrt(np.maximum(0.0, cfg.a_total_max * cfg.a_total_max - alat * alat))               # ax_caps, main.cpp:1530-1539
    Fd = 0.5 * cfg.rho_air * cfg.Cd * cfg.A_front_m2 * v * v
    Fr = cfg.mass_kg * 9.81 * cfg.c_rr
    with np.errstate(divide="ignore", invalid="ignore"):
        a_pow = np.where((cfg.P_max_W > 0) & (v > 1e-6), cfg.P_max_W / (cfg.mass_kg * v) - (Fd + Fr) / cfg.mass_kg, 1e9)
    col["a_acc_cap"] = np.maximum(0.0, np.minimum(np.minimum(a_res, cfg.a_long_acc_cap), a_pow))
    col["a_brk_cap"] = np.maximum(0.0, np.minimum(a_res, cfg.a_long_brake_cap) + (Fd + Fr) / cfg.mass_kg)
    col["a_power_cap"] = np.maximum(0.0, a_pow)
    out["columns"] = col
    a = col["d_mt_abs_m"]
    out["mt_offset_mean"], out["mt_offset_rms"] = float(a.sum() / max(1, N)), float(np.sqrt((a * a).sum() / max(1, N)))
    out["mt_offset_max"], out["mt_offset_argmax"] = (float(a.max()), int(a.argmax())) if N else (0.0, 0)
    if out["lap_mincurv"] > 0.0:
        out["lap_gain_vs_mincurv_pct"] = (out["lap_mincurv"] - mt.lap_time) / out["lap_mincurv"] * 100.0   # main.cpp:1586-1588
    return out


@dataclass
class CenterlineGeom:
    """One track's rows of <base>_with_geom.csv (main.cpp:1304) plus CL.L / CL.s0 (main.cpp:1255-1259)."""
    xy: np.ndarray
    s: np.ndarray
    heading: np.ndarray
    curvature: np.ndarray
    dist_inner: np.ndarray
    dist_outer: np.ndarray
    width: np.ndarray
    v_kappa: np.ndarray
    L: float
    s0: float
    samples: int

    @property
    def center_for_opt(self) -> np.ndarray:
        """The centre line the solver stages take (main.cpp:1681-1683): the first `samples` rows."""
        return self.xy[:self.samples]


class PackedGeom:
    """Host arrays in the rl_geom_desc layout + the matching output arrays (page-locked when a PinnedPool is given):
    pack once, then `run(ctx)` is exactly one rl_centerline_geom_batch call."""

    def __init__(self, mids: Sequence[np.ndarray], samples: Sequence[int], inner_rings: Sequence[np.ndarray],
                 outer_rings: Sequence[np.ndarray], closed=True, cfg: Optional[Config] = None, emit_closed_duplicate=True,
                 pool: Optional[PinnedPool] = None):
        from ._abi import RlGeomDesc, RlGeomOut
        cfg = cfg or Config()
        nt = len(mids)
        self.n_tracks = nt
        self.closed_arr = np.ascontiguousarray(np.broadcast_to(np.asarray(closed, dtype=bool), (nt,)).astype(np.int32))
        mids = [_f64(m, 2) for m in mids]
        segs = []
        for a, b in zip(inner_rings, outer_rings):
            segs += [_f64(a, 4), _f64(b, 4)]
        self.mid_off = np.zeros(nt + 1, dtype=np.int64)
        self.mid_off[1:] = np.cumsum([m.shape[0] for m in mids])
        self.seg_off = np.zeros(2 * nt + 1, dtype=np.int64)
        self.seg_off[1:] = np.cumsum([g.shape[0] for g in segs])
        self.mids_xy = np.ascontiguousarray(np.concatenate(mids, axis=0)) if nt else np.zeros((0, 2))
        self.seg = np.ascontiguousarray(np.concatenate(segs, axis=0)) if segs else np.zeros((0, 4))
        zeros = (lambda shape: pool.empty(shape, np.float64)) if pool else (lambda shape: np.zeros(shape))
        if pool:
            self.mids_xy, self.seg = pool.copy(self.mids_xy), pool.copy(self.seg)
        self.samples = np.ascontiguousarray(np.asarray(samples, dtype=np.int32).reshape(nt))
        self.params = cfg.to_params()
        d = RlGeomDesc()
        d.n_tracks, d.emit_closed_duplicate = nt, int(bool(emit_closed_duplicate))
        d.mid_off, d.mids_xy, d.samples, d.track_closed = _ptr(self.mid_off), _ptr(self.mids_xy), _ptr(self.samples), _ptr(self.closed_arr)
        d.seg_off, d.seg, d.params = _ptr(self.seg_off), _ptr(self.seg), C.cast(C.pointer(self.params), C.c_void_p)
        self.desc = d
        self.off = np.zeros(nt + 1, dtype=np.int64)
        st = lib().rl_geom_row_offsets(C.byref(d), self.off.ctypes.data_as(C.POINTER(C.c_int64)))
        if st != RL_OK:
            raise RacelineError(st, "rl_geom_row_offsets")
        rows = int(self.off[nt])
        self.rows = rows
        self.arr = {k: zeros(rows) for k in ("s_rel", "heading", "curvature", "dist_inner", "dist_outer", "width", "v_kappa")}
        self.xy, self.L, self.s0 = zeros((rows, 2)), np.zeros(nt), np.zeros(nt)
        o = RlGeomOut()
        o.xy, o.track_L, o.track_s0 = _ptr(self.xy), _ptr(self.L), _ptr(self.s0)
        for k, a in self.arr.items():
            setattr(o, k, _ptr(a))
        self.out = o

    @property
    def h2d_bytes(self):
        return int(self.mids_xy.nbytes + self.seg.nbytes + self.mid_off.nbytes + self.seg_off.nbytes + self.samples.nbytes + self.closed_arr.nbytes)

    @property
    def d2h_bytes(self):
        return int(self.rows * 9 * 8 + self.n_tracks * 16)

    def run(self, ctx: "Context"):
        ctx._check(lib().rl_centerline_geom_batch(ctx._h, C.byref(self.desc), C.byref(self.out)), "rl_centerline_geom_batch")
        return self

    def results(self) -> List["CenterlineGeom"]:
        out, arr = [], self.arr
        for t in range(self.n_tracks):
            a, b = int(self.off[t]), int(self.off[t + 1])
            out.append(CenterlineGeom(self.xy[a:b], arr["s_rel"][a:b], arr["heading"][a:b], arr["curvature"][a:b], arr["dist_inner"][a:b],
                                      arr["dist_outer"][a:b], arr["width"][a:b], arr["v_kappa"][a:b], float(self.L[t]), float(self.s0[t]),
                                      int(self.samples[t])))
        return out


def centerline_geom_batch(mids: Sequence[np.ndarray], samples: Sequence[int], inner_rings: Sequence[np.ndarray],
                          outer_rings: Sequence[np.ndarray], closed=True, cfg: Optional[Config] = None,
                          ctx: Optional[Context] = None, emit_closed_duplicate=True,
                          pool: Optional[PinnedPool] = None) -> List[CenterlineGeom]:
    """pipeline::make_centerline + the per-sample body of pipeline::compute_geom_and_save (main.cpp:1270-1335) for a
    batch of tracks: ordered mid points in, centre line + heading/curvature/ring distances/width/v_kappa rows out.
    `inner_rings` / `outer_rings` are segment arrays (ring_edges / polyline_edges of the *_from_mids points).
    With a PinnedPool the packed inputs and the output rows are page-locked (copies at PCIe speed)."""
    ctx = ctx or default_context()
    return PackedGeom(mids, samples, inner_rings, outer_rings, closed, cfg, emit_closed_duplicate, pool).run(ctx).results()


def synth_tracks(n_tracks, n_samples, m_per_ring=None, seed_base=0xB200, first_id=0, threads=0, pool=None):
    """Deterministic synthetic closed tracks; returns packed arrays (center_xy, seg, L)."""
    m = int(round(n_samples / 2.2)) if m_per_ring is None else int(m_per_ring)
    alloc = (lambda shape: pool.empty(shape, np.float64)) if pool else (lambda shape: np.empty(shape, dtype=np.float64))
    center = alloc((n_tracks * n_samples, 2))
    seg = alloc((n_tracks * 2 * m, 4))
    L = alloc(n_tracks)
    dp = C.POINTER(C.c_double)
    rc = lib().rl_synth_tracks(C.c_uint64(seed_base), C.c_int64(first_id), n_tracks, n_samples, m, threads,
                               center.ctypes.data_as(dp), seg.ctypes.data_as(dp), L.ctypes.data_as(dp))
    if rc != RL_OK:
        raise RacelineError(rc, "rl_synth_tracks")
    return center, seg, L, m
