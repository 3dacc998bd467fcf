#!/usr/bin/env python
"""bench.py -- raceline solves/s (min-curv + min-time) on N B200s, next to the host-CPU reference.

One step = one pass of the hot path over the whole job: 65,536 synthetic closed tracks, N = 2048 samples,
M = 931 cones per ring, default Config (BASELINE.json configs[3]) -- at every GPU count (strong scaling: rank r of G
owns tracks [r*65536/G, (r+1)*65536/G), no collective on the data path).  One solve = min-curvature stage +
min-time stage of one track (2 jobs).

    python bench.py [--gpus N] [--steps K] [--warmup W]              our arm (CUDA, through the C ABI)
    python bench.py --impl reference [--steps K] [--warmup W]        the reference's own CPU code, all host cores

`value`   : solves/s with inputs resident in HBM (kernel launches only, CUDA events, max over ranks).
`e2e`     : solves/s through rl_solve_batch with pinned HOST buffers: H2D + kernels + D2H inside the timed region,
            the same number of steps.
`roofline`: FP64 pipe.  achieved = algorithmic flops (SURVEY.md 8d formulas, from the counters the kernel
            returns) / kernel time; peak = DFMA throughput measured in this run by rl_measure_fp64_peak (reported next
            to SMs x 64 DFMA/clk x 2 x SM clock).
`cpu_baseline`: oracle/_ref/ref_harness (the unmodified reference TU) on a bounded sample of the same tracks,
            one pinned process per host core (rank 0, N = 1 only).
`parity`  : the GPU results of exactly those tracks against the reference's outputs (north-star tolerances); the run
            FAILS when they are out of tolerance.
`extra`   : the other BASELINE configs, each with its own roofline: configs[4] (N = 16,384, 8192 tracks in total),
            configs[2] (4096-Config sweep on competition_map2), configs[1] (the 7 shipped maps), configs[0]
            (single-track latency), plus the stage before the path (centre line + width/geometry).
`final_gather` (N > 1): lap times and raceline rasters of all problems gathered over NCCL from the device buffers.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SAMPLES = 2048
M_PER_RING = 931
SEED_BASE = 0xB200
TRACKS_TOTAL = 65536
WORKLOAD = ("BASELINE configs[3]: 65,536 synthetic closed tracks, N=2048 samples, M=931 cones/ring, default Config, "
            "min-curv + min-time per track")
METRIC = "raceline_solves_per_s"
UNIT = "solves/s"
MAPS = ["training_map", "competition_map1", "competition_map2", "competition_map3",
        "competition_map_testday1", "competition_map_testday2", "competition_map_testday3"]
# north-star tolerances (BASELINE.json)
TOL_ALPHA, TOL_KAPPA, TOL_V, TOL_LAP_REL = 1e-4, 1e-6, 1e-4, 1e-5


# ------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(local, world):
    """Multi-GPU runs: keep this rank (and the pinned buffers it allocates next) on the NUMA node its GPU hangs off.
    Returns what was done for the JSON line; a box with one node (or without the sysfs entries) is left alone."""
    info = {"gpu_node": None, "bound_cpus": None}
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        dev = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{dev}/numa_node").read().strip())
        info["gpu_node"] = node
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        if node < 0 or len(nodes) < 2 or world < 2:
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        mine = cpus & set(os.sched_getaffinity(0))
        if mine:
            os.sched_setaffinity(0, mine)
            info["bound_cpus"] = len(mine)
    except Exception:
        pass
    return info


def host_cores():
    try:
        return sorted(os.sched_getaffinity(0))
    except AttributeError:
        return list(range(os.cpu_count() or 1))


def host_mem_available_bytes():
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable:"):
                return int(ln.split()[1]) * 1024
    except Exception:
        pass
    return None


def run_reference_cpu(center, seg, L, n_tracks, cores, n=None, m=None, keep_results=False):
    """Solve n_tracks (both stages) with the reference's own CPU code, one pinned process per core.
    Returns (solves_per_s, wall_s, kind, solver_ms_sum, results or None); results = per-job dicts in job order."""
    from oracle import batchfile, oracle
    n, m = n or N_SAMPLES, m or M_PER_RING
    exe = oracle.ref_binary("ref_harness")
    p = oracle.default_params()
    jobs = np.array([[t, 0, st] for t in range(n_tracks) for st in (1, 2)])   # RL_STAGE_MINCURV, RL_STAGE_MINTIME
    nproc = min(len(cores), n_tracks)
    cuts = [(n_tracks * i) // nproc for i in range(nproc + 1)]   # contiguous track slices per process
    if exe:
        row = np.array([float(getattr(p, k)) for k in batchfile.PARAM_FIELDS])
        with tempfile.TemporaryDirectory() as td:
            bf = os.path.join(td, "batch.bin")
            batchfile.write_rlb1(bf, np.arange(n_tracks + 1) * n, np.arange(2 * n_tracks + 1) * m, L[:n_tracks],
                                 np.ones(n_tracks), center[:n_tracks * n], seg[:n_tracks * 2 * m], row[None, :], jobs)
            t0 = time.perf_counter()
            procs = []
            for i in range(nproc):
                a, b = cuts[i], cuts[i + 1]
                cmd = ["taskset", "-c", str(cores[i]), exe, "solve", bf, os.path.join(td, f"r{i}.bin"), str(2 * a), str(2 * (b - a))]
                procs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True))
            outs = [pr.communicate()[0] for pr in procs]
            wall = time.perf_counter() - t0
            if any(pr.returncode != 0 for pr in procs):
                raise RuntimeError("ref_harness failed")
            solver_ms = sum(float(o.split("solver_ms=")[1]) for o in outs)
            res = None
            if keep_results:
                res = []
                for i in range(nproc):
                    res.extend(batchfile.read_rlr1(os.path.join(td, f"r{i}.bin")))
        return n_tracks / wall, wall, "reference", solver_ms, res
    # fall back to the C restatement (still the checker, still CPU): one process per core through multiprocessing
    import multiprocessing as mp
    c3, s4 = center.reshape(-1, n, 2), seg.reshape(-1, 2, m, 4)

    def work(a, b, q):
        t0 = time.perf_counter()
        out = []
        for t in range(a, b):
            for st in (1, 2):
                o = oracle.solve(st, c3[t], s4[t, 0], s4[t, 1], L[t], True, p)
                if keep_results:
                    out.append({"n": n, "stage": st, "lap_time": o.get("lap_time", 0.0), "backtracks": o["stats"].backtracks,
                                "xy": o["xy"], **{k: o[k] for k in ("heading", "curvature", "alpha_total", "alpha_last", "v", "ax") if k in o}})
        q.put((a, (time.perf_counter() - t0) * 1e3, out))

    oracle.lib()
    q = mp.Queue()
    t0 = time.perf_counter()
    ps = [mp.Process(target=work, args=(cuts[i], cuts[i + 1], q)) for i in range(nproc)]
    for pr in ps:
        pr.start()
    got = sorted((q.get() for _ in ps), key=lambda g: g[0])
    for pr in ps:
        pr.join()
    wall = time.perf_counter() - t0
    res = [r for g in got for r in g[2]] if keep_results else None
    return n_tracks / wall, wall, "port", sum(g[1] for g in got), res


class ClockSampler:
    """nvidia-smi clocks / throttle reasons of one GPU during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self._stop, self._th = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._th = threading.Thread(target=self._run, daemon=True)
        self._th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._th.join(timeout=10)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(self.rows), "reasons": reasons}


def algorithmic_flops(stats, n_jobs, stages):
    """F_alg = F_pgd + F_geom + F_v + F_ray, from the counters the kernel returns; only work the
    kernel executed is credited (ray term: 17 flop x exact ray/segment tests actually run)."""
    f = 0.0
    for j in range(n_jobs):
        st = stats[j]
        n, mt = st.n, stages[j] == 2
        per_eval = 35.0 if mt else 33.0
        trials = max(0, st.evals - st.outer_done)
        f += n * (st.evals * per_eval + trials * 7.0)          # cost/grad evaluations + line-search trials
        f += n * 70.0 * (st.outer_done + 1)                    # normals, lin-geom, heading/curvature per outer
        if mt:
            f += min(5040.0 * n, 28.0 * n * st.vpass_rounds)   # v(s) sweeps (reference: 15 calls x 6 x 2 sweeps x 28 flop)
        f += 17.0 * st.ray_tests
    return f


def solve_jobs(n_tracks):
    jobs = np.empty((2 * n_tracks, 3), dtype=np.int64)
    jobs[0::2, 0] = jobs[1::2, 0] = np.arange(n_tracks)
    jobs[:, 1] = 0
    jobs[0::2, 2], jobs[1::2, 2] = 1, 2
    return jobs


def fp64_peak_info(ctx, clk_mhz):
    """Measured DFMA throughput next to the arithmetic ceiling SMs x 64 DFMA/clk x 2 flop x SM clock."""
    import torch
    peak = ctx.fp64_peak_tflops()
    sms = torch.cuda.get_device_properties(ctx.device).multi_processor_count
    nominal = sms * 64 * 2 * (clk_mhz or 1965.0) * 1e6 / 1e12
    return peak, {"measured_tflops": peak, "sms": sms, "dfma_per_clk_per_sm": 64, "sm_mhz": clk_mhz,
                  "sms_x_64_x_2_x_clock_tflops": nominal, "measured_over_nominal": peak / nominal if nominal else None}


class DeviceTimer:
    """CUDA-event timing on the stream the library launches on."""

    def __init__(self, stream):
        import torch
        self.torch, self.stream = torch, stream
        self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def __enter__(self):
        self.e0.record(self.stream)
        return self

    def __exit__(self, *a):
        self.e1.record(self.stream)
        self.torch.cuda.synchronize()
        self.ms = self.e0.elapsed_time(self.e1)


# ------------------------------------------------------------------------------------------------
# the secondary workloads (BASELINE configs[4], [2], [1], [0] and the stage before the path)
# ------------------------------------------------------------------------------------------------
def extra_resident(rl, ctx, stream, name, workload, pb, jobs, fp64_peak, steps, warmup, n_solves, world, dist, also_e2e=True):
    """Resident + end-to-end throughput and the FP64 roofline of one packed batch; max over ranks, sum of the work."""
    import torch
    dev = rl.DeviceBatch(ctx, pb)
    for _ in range(warmup):
        dev.solve()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    with DeviceTimer(stream) as t:
        for _ in range(steps):
            dev.solve()
    ms = t.ms
    dev.download(); dev.sync()
    flops = algorithmic_flops(pb.out_stats, pb.n_jobs, jobs[:, 2])
    launches = dev.launches_per_solve
    dev.close()
    ms_e2e = None
    if also_e2e:
        ctx.solve_batch(pb)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        with DeviceTimer(stream) as t:
            for _ in range(steps):
                ctx.solve_batch(pb)
        ms_e2e = t.ms
    red = torch.tensor([ms, ms_e2e or 0.0], dtype=torch.float64, device="cuda")
    tot = torch.tensor([flops, float(n_solves)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms, ms_e2e_r = float(red[0]), float(red[1])
    solves = float(tot[1])
    kernel_ms = ms / steps
    achieved = (float(tot[0]) / world) / (kernel_ms * 1e-3) / 1e12
    out = {"name": name, "workload": workload, "metric": METRIC, "unit": UNIT, "value": solves * steps / (ms * 1e-3),
           "ms_per_step": kernel_ms, "steps": steps, "solves_per_step": solves, "gpu_launches_per_step": launches,
           "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                        "frac": achieved / fp64_peak if fp64_peak else None, "flops_per_step_per_gpu": float(tot[0]) / world}}
    if also_e2e:
        out["e2e"] = {"value": solves * steps / (ms_e2e_r * 1e-3), "unit": UNIT, "h2d_bytes_per_step": pb.h2d_bytes,
                      "d2h_bytes_per_step": pb.d2h_bytes}
    return out


def run_extras(args, rl, ctx, stream, fp64_peak, world, rank, dist):
    import torch
    from practice_path_planning_for_formula_student_driverless_b200 import sharding
    extras = []
    cfg = rl.Config()
    threads = max(1, len(host_cores()) // max(1, world))
    # ---- configs[4]: long-track stress, N = 16,384, 8192 tracks in total, one 8-CTA cluster per job ----
    if args.long_tracks_total > 0:
        n, m = 16384, 7447
        lo, hi = sharding.shard_bounds(args.long_tracks_total, world, rank)
        pool = rl.PinnedPool()
        center, seg, L, _ = rl.synth_tracks(hi - lo, n, m, seed_base=SEED_BASE, first_id=lo, threads=threads, pool=pool)
        jobs = solve_jobs(hi - lo)
        pb = rl.PackedBatch.from_arrays(np.arange(hi - lo + 1, dtype=np.int64) * n, np.arange(2 * (hi - lo) + 1, dtype=np.int64) * m,
                                        center, seg, L, np.ones(hi - lo, np.int32), [cfg.to_params()], jobs, pool=pool)
        e = extra_resident(rl, ctx, stream, "configs[4]", f"long-track stress: {args.long_tracks_total} synthetic closed tracks, "
                           "N=16384 samples, M=7447 cones/ring, default Config, one 8-CTA cluster per job", pb, jobs, fp64_peak,
                           max(1, args.extra_steps), 1, hi - lo, world, dist, also_e2e=False)
        e["mean_lap_s"] = float(np.mean([pb.out_stats[j].lap_time for j in range(1, pb.n_jobs, 2)]))
        extras.append(e)
        del pb, center, seg, L
        pool.close(force=True)
    if rank != 0:
        return extras
    # ---- configs[2]: 8^4 Config sweep on competition_map2, one copy of the geometry ----
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "competition_map2.npz")))
    tr = rl.Track(g["center_xy"], g["inner_seg"], g["outer_seg"], float(g["L"]))
    lam = np.geomspace(4e-4, 6.4e-3, 8); mus = np.linspace(1.15, 1.6, 8); pw = np.linspace(20e3, 80e3, 8); wt = np.linspace(0.0, 3.5, 8)
    cfgs = [rl.Config(lambda_smooth=float(a), mu=float(b), a_total_max=float(b) * 9.81, P_max_W=float(c), w_time_gain=float(d))
            for a in lam for b in mus for c in pw for d in wt]
    jobs = np.array([[0, k, st] for k in range(len(cfgs)) for st in (1, 2)], dtype=np.int64)
    pool = rl.PinnedPool()
    pb = rl.PackedBatch([tr], [c.to_params() for c in cfgs], jobs, pool=pool)
    e = extra_resident(rl, ctx, stream, "configs[2]", "Config sweep on competition_map2 (N=252): 4096 combos of lambda_smooth x mu x "
                       "P_max_W x w_time_gain, min-curv + min-time each, one copy of the geometry", pb, jobs, fp64_peak,
                       max(10, args.extra_steps), 3, len(cfgs), 1, None)
    laps = np.array([pb.out_stats[j].lap_time for j in range(1, pb.n_jobs, 2)])
    e["best_combo"] = int(np.argmin(laps)); e["best_lap_s"] = float(laps.min()); e["worst_lap_s"] = float(laps.max())
    extras.append(e)
    del pb
    pool.close(force=True)
    # ---- configs[1]: the 7 shipped maps batched, checked against the reference goldens ----
    gs = [dict(np.load(os.path.join(ROOT, "tests", "golden", m + ".npz"))) for m in MAPS]
    tracks = [rl.Track(q["center_xy"], q["inner_seg"], q["outer_seg"], float(q["L"])) for q in gs]
    jobs = np.array([[t, 0, st] for t in range(len(tracks)) for st in (1, 2)], dtype=np.int64)
    pb = rl.PackedBatch(tracks, [cfg.to_params()], jobs)
    e = extra_resident(rl, ctx, stream, "configs[1]", "the 7 shipped maps (N=187..261) batched on one B200, min-curv + min-time each",
                       pb, jobs, fp64_peak, 20, 3, len(tracks), 1, None)
    worst = {"dalpha": 0.0, "dkappa": 0.0, "dv": 0.0, "lap_rel": 0.0, "bt_equal": True}
    for t, q in enumerate(gs):
        for st, pre in ((1, "mc_"), (2, "mt_")):
            r = pb.result(2 * t + (st - 1))
            worst["dalpha"] = max(worst["dalpha"], float(np.max(np.abs(r.alpha_total - q[pre + "alpha_total"]))))
            worst["dkappa"] = max(worst["dkappa"], float(np.max(np.abs(r.curvature - q[pre + "curvature"]))))
            worst["bt_equal"] &= bool(r.stats.backtracks == int(np.sum(q[pre + "bt"])))
            if st == 2:
                worst["dv"] = max(worst["dv"], float(np.max(np.abs(r.v - q["mt_v"]))))
                worst["lap_rel"] = max(worst["lap_rel"], abs(r.lap_time - float(q["mt_lap_time"])) / float(q["mt_lap_time"]))
    e["parity_vs_reference_goldens"] = worst
    if not (worst["dalpha"] <= TOL_ALPHA and worst["dkappa"] <= TOL_KAPPA and worst["dv"] <= TOL_V and worst["lap_rel"] <= TOL_LAP_REL and worst["bt_equal"]):
        raise SystemExit(f"bench.py: shipped maps out of tolerance against the reference goldens: {worst}")
    extras.append(e)
    # ---- configs[0]: single-track latency through the single-problem entry points (host buffers in and out) ----
    q = gs[0]
    lat = []
    for it in range(110):
        t0 = time.perf_counter()
        rl.compute_min_curvature_raceline(q["center_xy"], q["inner_seg"], q["outer_seg"], cfg.veh_width_m, float(q["L"]), True, cfg, ctx=ctx)
        rmt = rl.compute_min_time_raceline(q["center_xy"], q["inner_seg"], q["outer_seg"], cfg.veh_width_m, float(q["L"]), True, cfg, ctx=ctx)
        if it >= 10:
            lat.append((time.perf_counter() - t0) * 1e3)
    lat.sort()
    pb1 = rl.PackedBatch([tracks[0]], [cfg.to_params()], jobs[:2])
    with_batch = []
    for it in range(110):
        t0 = time.perf_counter()
        ctx.solve_batch(pb1)
        if it >= 10:
            with_batch.append((time.perf_counter() - t0) * 1e3)
    with_batch.sort()
    fl = algorithmic_flops(pb1.out_stats, 2, jobs[:2, 2])
    p50b = with_batch[len(with_batch) // 2]
    extras.append({"name": "configs[0]", "workload": "training_map (N=216), one track, min-curv + min-time, host buffers in and out, 100 calls",
                   "metric": "single_track_latency_ms", "unit": "ms", "higher_is_better": False,
                   "two_calls_p50": lat[len(lat) // 2], "two_calls_p99": lat[min(len(lat) - 1, int(0.99 * len(lat)))],
                   "one_batch_call_p50": p50b, "one_batch_call_p99": with_batch[min(len(with_batch) - 1, int(0.99 * len(with_batch)))],
                   "value": p50b, "lap_s": float(rmt.lap_time),
                   "roofline": {"bound": "fp64", "achieved": fl / (p50b * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                                "frac": fl / (p50b * 1e-3) / 1e12 / fp64_peak if fp64_peak else None,
                                "note": "two CTAs on a 148-SM device: latency, not throughput"}})
    # ---- the stage before the path: centre line + width/geometry rows, batched ----
    try:
        ng = args.geom_tracks
        c8, s8, L8, m8 = rl.synth_tracks(ng, N_SAMPLES, M_PER_RING, seed_base=SEED_BASE)
        c8, s8 = c8.reshape(ng, N_SAMPLES, 2), s8.reshape(ng, 2, m8, 4)
        # ordered mid points = every 2nd..3rd centre sample (931 mid points for 2048 rows, the reference's ratio)
        idx = np.linspace(0, N_SAMPLES, M_PER_RING, endpoint=False).astype(int)
        mids = [c8[t][idx] for t in range(ng)]
        inner = [s8[t, 0] for t in range(ng)]
        outer = [s8[t, 1] for t in range(ng)]
        gpool = rl.PinnedPool()
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs")
        except Exception:
            hbm_peak = None
        pg = rl.PackedGeom(mids, [N_SAMPLES] * ng, inner, outer, closed=True, cfg=cfg, pool=gpool)
        ctx.set_option("geom_chunks", 1)    # one upload | the two kernels | one download: the kernels' own device time
        pg.run(ctx); pg.run(ctx)
        kms = ctx.last_kernel_ms()
        ctx.set_option("geom_chunks", 0)    # the call as a user gets it: a pipeline of track ranges
        walls = []
        for it in range(4):
            t0 = time.perf_counter()
            pg.run(ctx)                     # one rl_centerline_geom_batch call: pinned H2D + two kernels + pinned D2H
            walls.append(time.perf_counter() - t0)
        wall = min(walls[1:])
        rows = ng * N_SAMPLES
        # algorithmic HBM bytes per track: mid points and both rings in (16 B + 2 x 32 B per cone), 9 doubles per row out
        alg_bytes = ng * (16.0 * M_PER_RING + 64.0 * M_PER_RING + 72.0 * N_SAMPLES)
        extras.append({"name": "stage_before_the_path", "workload": f"rl_centerline_geom_batch: {ng} tracks, 931 mid points -> 2048 rows each, "
                       "spline fit + uniform resample + distancesToRings (main.cpp:1270-1335)",
                       "metric": "tracks_per_s", "unit": "tracks/s", "value": ng / (kms * 1e-3), "kernel_ms": kms,
                       "rows_per_s": rows / (kms * 1e-3),
                       "e2e": {"value": ng / wall, "unit": "tracks/s", "ms_wall": wall * 1e3, "h2d_bytes_per_step": pg.h2d_bytes,
                               "d2h_bytes_per_step": pg.d2h_bytes,
                               "note": "one rl_centerline_geom_batch call: pinned H2D + kernels + pinned D2H of 9 output columns, "
                                       f"pipelined over {max(1, min(8, ng // 1024))} track range(s)"},
                       "roofline": {"bound": "hbm", "achieved": alg_bytes / (kms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                    "frac": (alg_bytes / (kms * 1e-3) / 1e9 / hbm_peak) if hbm_peak else None,
                                    "note": "two kernels (centerline_kernel, ring_distance_kernel); FP64 ray tests and a sequential "
                                            "tridiagonal solve per track, far from either roof"}})
        del pg
        gpool.close(force=True)
    except Exception as ex:   # the stage is a "next" row: report, do not fail the headline
        extras.append({"name": "stage_before_the_path", "error": repr(ex)})
    return extras


# ------------------------------------------------------------------------------------------------
def parity_block(pb, ref_results, n_ref):
    """GPU results of the first n_ref tracks against the reference's outputs for the same tracks."""
    par = {"tracks": n_ref, "max_dalpha": 0.0, "max_dkappa": 0.0, "max_dv": 0.0, "lap_rel": 0.0, "bt_equal": True, "bt_compared": 0,
           "tolerances": {"alpha_m": TOL_ALPHA, "kappa_1_per_m": TOL_KAPPA, "v_mps": TOL_V, "lap_rel": TOL_LAP_REL},
           "note": "backtracks compared where the reference logs them (min-time '[PG] bt=' lines, main.cpp:1019-1022)"}
    for j in range(2 * n_ref):
        r, o = pb.result(j), ref_results[j]
        par["max_dalpha"] = max(par["max_dalpha"], float(np.max(np.abs(r.alpha_total - o["alpha_total"]))))
        par["max_dkappa"] = max(par["max_dkappa"], float(np.max(np.abs(r.curvature - o["curvature"]))))
        if o["stage"] == 2:
            par["max_dv"] = max(par["max_dv"], float(np.max(np.abs(r.v - o["v"]))))
            par["lap_rel"] = max(par["lap_rel"], abs(r.lap_time - o["lap_time"]) / o["lap_time"])
        if o.get("backtracks", -1) >= 0:
            par["bt_compared"] += 1
            par["bt_equal"] &= bool(r.stats.backtracks == o["backtracks"])
    par["ok"] = bool(par["max_dalpha"] <= TOL_ALPHA and par["max_dkappa"] <= TOL_KAPPA and par["max_dv"] <= TOL_V and
                     par["lap_rel"] <= TOL_LAP_REL and par["bt_equal"])
    return par


def bench_ours(args):
    import torch
    import torch.distributed as dist

    import practice_path_planning_for_formula_student_driverless_b200 as rl
    from practice_path_planning_for_formula_student_driverless_b200 import sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local, world)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    total = args.tracks_total
    # pinned host buffers of the end-to-end arm: 92 KB in + 262 KB out per track
    need = total * (16 * N_SAMPLES + 64 * M_PER_RING + 128 * N_SAMPLES) * 1.15
    avail = host_mem_available_bytes()
    reduced = None
    if avail is not None and need > 0.8 * avail:
        reduced = total
        total = max(world * 296, int(total * 0.8 * avail / need) // (world * 296) * (world * 296))
    lo, hi = sharding.shard_bounds(total, world, rank)
    tpg = hi - lo
    pool = rl.PinnedPool()
    t0 = time.time()
    threads = max(1, len(host_cores()) // max(1, world))
    center, seg, L, m = rl.synth_tracks(tpg, N_SAMPLES, M_PER_RING, seed_base=SEED_BASE, first_id=lo, threads=threads, pool=pool)
    gen_s = time.time() - t0
    jobs = solve_jobs(tpg)
    cfg = rl.Config()
    pb = rl.PackedBatch.from_arrays(np.arange(tpg + 1, dtype=np.int64) * N_SAMPLES, np.arange(2 * tpg + 1, dtype=np.int64) * m,
                                    center, seg, L, np.ones(tpg, np.int32), [cfg.to_params()], jobs, pool=pool)
    ctx = rl.Context(local)
    # a real (non-default) stream shared by the library and the timing events: rl_set_stream(NULL) would
    # select the context's own stream, which torch.cuda.Event on the default stream cannot see
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)
    if args.solve_chunks:
        ctx.set_option("solve_chunks", args.solve_chunks)
    if args.chunk_streams:
        ctx.set_option("chunk_streams", args.chunk_streams)
    dev = rl.DeviceBatch(ctx, pb)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident-input throughput ----
    for _ in range(args.warmup):
        dev.solve()
    barrier()
    with ClockSampler(local) as clk:
        with DeviceTimer(stream) as tm:
            for _ in range(args.steps):
                dev.solve()
    ms = tm.ms
    barrier()
    dev.download(); dev.sync()
    launches = dev.launches_per_solve * args.steps
    flops_step = algorithmic_flops(pb.out_stats, pb.n_jobs, jobs[:, 2])
    lap_mean = float(np.mean([pb.out_stats[j].lap_time for j in range(1, pb.n_jobs, 2)]))
    diag = {"exist_scans_per_job": float(np.mean([pb.out_stats[j].exist_scans for j in range(pb.n_jobs)])),
            "ray_tests_per_job": float(np.mean([pb.out_stats[j].ray_tests for j in range(pb.n_jobs)])),
            "vpass_rounds_per_mt_job": float(np.mean([pb.out_stats[j].vpass_rounds for j in range(1, pb.n_jobs, 2)]))}
    clocks = clk.summary()
    fp64_peak, fp64_info = fp64_peak_info(ctx, clocks.get("sm_mhz"))

    # ---- the final gather over NCCL, from the device buffers (north_star: lap times and rasters) ----
    gather = None
    if world > 1:
        laps_dev = dev.device_tensor("lap_time")[1::2]
        xy_dev = dev.device_tensor("xy").view(pb.n_jobs, N_SAMPLES, 2)[1::2].reshape(-1, 2)     # min-time racelines
        sharding.gather_lap_times(laps_dev, total); sharding.gather_rasters(xy_dev, total, N_SAMPLES)    # warm NCCL
        barrier()
        with DeviceTimer(stream) as tg:
            all_laps = sharding.gather_lap_times(laps_dev, total)
        with DeviceTimer(stream) as tr:
            all_xy = sharding.gather_rasters(xy_dev, total, N_SAMPLES)
        mine = np.array([pb.out_stats[j].lap_time for j in range(1, pb.n_jobs, 2)])
        ok = bool(np.array_equal(all_laps[lo:hi], mine)) and bool(torch.equal(all_xy[lo * N_SAMPLES:hi * N_SAMPLES], xy_dev))
        flag = torch.tensor([1.0 if ok else 0.0, tg.ms, tr.ms], dtype=torch.float64, device="cuda")
        mn = flag.clone(); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        mx = flag.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        gather = {"backend": "nccl", "lap_bytes_total": 8 * total, "lap_us": float(mx[1]) * 1e3,
                  "raster_bytes_total": 16 * N_SAMPLES * total, "raster_us": float(mx[2]) * 1e3,
                  "raster_gbs_per_rank": 16 * N_SAMPLES * total / (float(mx[2]) * 1e-3) / 1e9,
                  "every_rank_holds_all": bool(float(mn[0]) == 1.0), "mean_lap_all_s": float(np.mean(all_laps))}
        del all_xy
        if not gather["every_rank_holds_all"]:
            raise SystemExit("bench.py: the NCCL final gather returned something else than the ranks' own results")

    # keep the resident results of the parity sample before the end-to-end arm overwrites the host arrays
    n_ref = 0
    if world == 1 and not args.no_cpu_baseline:
        n_ref = min(tpg, len(host_cores()) * args.cpu_tracks_per_core)
    dev.close()

    # ---- end to end: pinned host buffers through rl_solve_batch, the same number of steps ----
    ctx.solve_batch(pb)   # warm the context's device buffers
    barrier()
    with DeviceTimer(stream) as te:
        for _ in range(args.steps):
            ctx.solve_batch(pb)
    ms_e2e = te.ms
    barrier()

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
    fl = torch.tensor([flops_step], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(fl, op=dist.ReduceOp.SUM)
    ms, ms_e2e = float(t[0]), float(t[1])
    value = total * args.steps / (ms * 1e-3)
    e2e_value = total * args.steps / (ms_e2e * 1e-3)
    kernel_ms = ms / max(1, launches)                      # one solve_kernel launch per step and rank
    achieved_tf = (float(fl[0]) / world) / (kernel_ms * 1e-3) / 1e12
    alg_bytes = tpg * (128.0 * N_SAMPLES + 64.0 * M_PER_RING)   # inputs + outputs per solve
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    traffic = None
    try:
        # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture, scaled to this launch's track count
        traffic = json.load(open(os.path.join(ROOT, "profiles", "solve_kernel_traffic.json")))["dram_bytes_per_track"] * tpg
    except Exception:
        pass

    out = None
    if rank == 0:
        step_s = ms_e2e / args.steps * 1e-3
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "tracks_total": total, "tracks_per_gpu": tpg,
                       "n_samples": N_SAMPLES, "m_per_ring": M_PER_RING, "seed_base": SEED_BASE,
                       "l2": "inputs larger than L2 (%.0f MB per GPU)" % (pb.h2d_bytes / 1e6), "parallelism": f"dp{world}",
                       "mean_lap_s": lap_mean, "gen_s": round(gen_s, 2), "kernel_diag": diag},
            "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved_tf / fp64_peak if fp64_peak else None, "traffic": traffic,
                         "peak_source": "DFMA throughput measured in this run (rl_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 entry",
                         "fp64_peak": fp64_info,
                         "flops_per_launch": float(fl[0]) / world, "kernel_ms": kernel_ms,
                         "hbm_achieved_gbs": alg_bytes / (kernel_ms * 1e-3) / 1e9, "hbm_peak_gbs": peaks.get("hbm_gbs")},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": pb.h2d_bytes, "d2h_bytes_per_step": pb.d2h_bytes,
                    "steps": args.steps, "ms_per_step": ms_e2e / args.steps,
                    "host_gbs_per_rank": {"h2d": pb.h2d_bytes / step_s / 1e9, "d2h": pb.d2h_bytes / step_s / 1e9},
                    "numa": numa},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if reduced:
            out["config"]["tracks_total_requested"] = reduced
            out["config"]["note"] = "track count reduced to fit the host's available memory for the pinned end-to-end buffers"
        if gather:
            out["final_gather"] = gather
        if n_ref:
            cores = host_cores()
            sps, wall, kind, solver_ms, res = run_reference_cpu(center, seg, L, n_ref, cores, keep_results=True)
            out["cpu_baseline"] = {"value": sps, "unit": UNIT, "cores": min(len(cores), n_ref), "kind": kind,
                                   "sample": f"first {n_ref} tracks of the GPU batch, both stages, wall {wall:.1f} s "
                                             f"(sum of solver-only time {solver_ms / 1e3:.1f} core-s)"}
            out["parity"] = parity_block(pb, res, n_ref)    # pb holds the end-to-end arm's results of all tracks
    del dev
    if not args.no_extras:
        extras = run_extras(args, rl, ctx, stream, fp64_peak, world, rank, dist if world > 1 else None)
        if out is not None:
            out["extra"] = extras
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()
    del pb, center, seg, L
    pool.close(force=True)
    if out is not None:
        args.emit(out)
        if "parity" in out and not out["parity"]["ok"]:
            raise SystemExit(f"bench.py: GPU results out of tolerance against the reference: {out['parity']}")


def bench_reference(args):
    """The reference's own CPU implementation on the host cores; inputs from oracle/libsynth_tracks.so (the CUDA
    library is not loaded in this arm)."""
    from oracle import oracle
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    cores = host_cores()
    per_step = len(cores) * args.ref_tracks_per_core
    center, seg, L = oracle.synth_tracks(per_step, N_SAMPLES, M_PER_RING, seed_base=SEED_BASE, first_id=0, threads=len(cores))
    times, kind = [], "port"
    for it in range(args.warmup + args.steps):
        sps, wall, kind, _, _ = run_reference_cpu(center, seg, L, per_step, cores)
        if it >= args.warmup:
            times.append(wall)
    tot = sum(times)
    value = per_step * len(times) / tot
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_samples": N_SAMPLES, "m_per_ring": M_PER_RING,
                   "seed_base": SEED_BASE, "tracks_per_step": per_step,
                   "note": "each step = a bounded sample of the workload: the first tracks of the 65,536 (the full set is ~48 core-hours)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": min(len(cores), per_step), "kind": kind,
                         "sample": f"{per_step} tracks per step (the first tracks of the GPU batch), one pinned process per core"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    args.emit(out)


class StdoutOnlyJson:
    """Libraries (NCCL prints its version banner on stdout) must not put anything next to the ONE JSON line the
    driver parses: file descriptor 1 points at stderr until the result is printed."""

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, obj):
        sys.stdout.flush()
        os.dup2(self._saved, 1)
        print(json.dumps(obj), flush=True)
        os.dup2(2, 1)

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self._saved, 1)
        os.close(self._saved)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--tracks-total", type=int, default=TRACKS_TOTAL, help="tracks of the whole job (all GPUs together)")
    ap.add_argument("--tracks-per-gpu", type=int, default=0, help="development: tracks_total = this x GPUs")
    ap.add_argument("--long-tracks-total", type=int, default=8192, help="configs[4]: N=16384 tracks of the whole job (0 = skip)")
    ap.add_argument("--extra-steps", type=int, default=1)
    ap.add_argument("--geom-tracks", type=int, default=8192)
    ap.add_argument("--cpu-tracks-per-core", type=int, default=4)
    ap.add_argument("--ref-tracks-per-core", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--solve-chunks", type=int, default=0, help="development: rl_set_option('solve_chunks')")
    ap.add_argument("--chunk-streams", type=int, default=0, help="development: rl_set_option('chunk_streams')")
    args = ap.parse_args()
    if args.tracks_per_gpu > 0:
        args.tracks_total = args.tracks_per_gpu * int(os.environ.get("WORLD_SIZE", "1"))
    with StdoutOnlyJson() as out:
        args.emit = out.emit
        if args.impl == "reference":
            bench_reference(args)
        else:
            bench_ours(args)


if __name__ == "__main__":
    main()
