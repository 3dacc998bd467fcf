#!/usr/bin/env python
"""bench.py -- raceline solves/s (min-curv + min-time) on N B200s, next to the host-CPU reference.

One step = one pass of the hot path over this rank's batch of synthetic closed tracks
(BASELINE.json configs[3]: N = 2048 samples, M = 931 cones per ring; 8192 tracks per GPU, i.e.
65,536 at 8 GPUs -- weak scaling, problems are independent and sharded with no collective).
One solve = min-curvature stage + min-time stage of one track (2 jobs).

    python bench.py [--gpus N] [--steps K] [--warmup W]              our arm (CUDA, through the C ABI)
    python bench.py --impl reference [--steps K] [--warmup W]        the reference's own CPU code, all host cores

`value`  : solves/s with inputs resident in HBM (kernel launches only, CUDA events, max over ranks).
`e2e`    : solves/s through rl_solve_batch with pinned HOST buffers: H2D + kernels + D2H inside the timed region.
`roofline`: FP64 pipe.  achieved = algorithmic flops (SURVEY.md 8d formulas, from the counters the kernel
           returns) / kernel time; peak = DFMA throughput measured in this run by rl_measure_fp64_peak.
`cpu_baseline`: oracle/_ref/ref_harness (the unmodified reference TU) on a bounded sample of the same tracks,
           one pinned process per host core.
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
WORKLOAD = ("BASELINE configs[3]: synthetic closed tracks, N=2048 samples, M=931 cones/ring, default Config, "
            "min-curv + min-time per track")
METRIC = "raceline_solves_per_s"
UNIT = "solves/s"


# ------------------------------------------------------------------------------------------------
def host_cores():
    try:
        return sorted(os.sched_getaffinity(0))
    except AttributeError:
        return list(range(os.cpu_count() or 1))


def make_tracks(first_id, n_tracks, pool=None):
    import practice_path_planning_for_formula_student_driverless_b200 as rl
    center, seg, L, m = rl.synth_tracks(n_tracks, N_SAMPLES, M_PER_RING, seed_base=SEED_BASE, first_id=first_id, pool=pool)
    return center, seg, L, m


def run_reference_cpu(center, seg, L, n_tracks, cores, tag="ref"):
    """Solve n_tracks (both stages) with the reference's own CPU code, one pinned process per core.
    Returns (solves_per_s, wall_s, kind, solver_ms_sum)."""
    from oracle import batchfile, oracle
    from practice_path_planning_for_formula_student_driverless_b200 import RL_STAGE_MINCURV, RL_STAGE_MINTIME
    n, m = N_SAMPLES, M_PER_RING
    exe = oracle.ref_binary("ref_harness")
    p = oracle.default_params()
    jobs = np.array([[t, 0, st] for t in range(n_tracks) for st in (RL_STAGE_MINCURV, RL_STAGE_MINTIME)])
    nproc = min(len(cores), n_tracks)
    # contiguous track slices per process
    cuts = [(n_tracks * i) // nproc for i in range(nproc + 1)]
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
        return n_tracks / wall, wall, "reference", solver_ms
    # fall back to the C restatement (still the checker, still CPU): one process per core through multiprocessing
    import multiprocessing as mp
    c3, s4 = center.reshape(-1, n, 2), seg.reshape(-1, 2, m, 4)

    def work(a, b, q):
        t0 = time.perf_counter()
        for t in range(a, b):
            for st in (RL_STAGE_MINCURV, RL_STAGE_MINTIME):
                oracle.solve(st, c3[t], s4[t, 0], s4[t, 1], L[t], True, p)
        q.put((time.perf_counter() - t0) * 1e3)

    oracle.lib()
    q = mp.Queue()
    t0 = time.perf_counter()
    ps = [mp.Process(target=work, args=(cuts[i], cuts[i + 1], q)) for i in range(nproc)]
    for pr in ps:
        pr.start()
    ms = sum(q.get() for _ in ps)
    for pr in ps:
        pr.join()
    wall = time.perf_counter() - t0
    return n_tracks / wall, wall, "port", ms


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


# ------------------------------------------------------------------------------------------------
def bench_ours(args):
    import torch
    import torch.distributed as dist

    import practice_path_planning_for_formula_student_driverless_b200 as rl

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tpg = args.tracks_per_gpu
    pool = rl.PinnedPool()
    t0 = time.time()
    threads = max(1, len(host_cores()) // max(1, world))
    center, seg, L, m = rl.synth_tracks(tpg, N_SAMPLES, M_PER_RING, seed_base=SEED_BASE, first_id=rank * tpg, threads=threads,
                                        pool=pool)
    gen_s = time.time() - t0
    jobs = np.empty((2 * tpg, 3), dtype=np.int64)
    jobs[0::2, 0] = jobs[1::2, 0] = np.arange(tpg)
    jobs[:, 1] = 0
    jobs[0::2, 2], jobs[1::2, 2] = rl.RL_STAGE_MINCURV, rl.RL_STAGE_MINTIME
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
    dev = rl.DeviceBatch(ctx, pb)
    fp64_peak = ctx.fp64_peak_tflops()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident-input throughput ----
    for _ in range(args.warmup):
        dev.solve()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record(stream)
        for _ in range(args.steps):
            dev.solve()
        e1.record(stream)
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    barrier()
    dev.download(); dev.sync()
    launches = dev.launches_per_solve * args.steps
    flops_step = algorithmic_flops(pb.out_stats, pb.n_jobs, jobs[:, 2])
    lap_mean = float(np.mean([pb.out_stats[j].lap_time for j in range(1, pb.n_jobs, 2)]))
    diag = {"exist_scans_per_job": float(np.mean([pb.out_stats[j].exist_scans for j in range(pb.n_jobs)])),
            "ray_tests_per_job": float(np.mean([pb.out_stats[j].ray_tests for j in range(pb.n_jobs)])),
            "vpass_rounds_per_mt_job": float(np.mean([pb.out_stats[j].vpass_rounds for j in range(1, pb.n_jobs, 2)]))}

    # ---- end to end: pinned host buffers through rl_solve_batch ----
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    ctx.solve_batch(pb)   # warm the context's device buffers
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(stream)
    for _ in range(e2e_steps):
        ctx.solve_batch(pb)
    f1.record(stream)
    torch.cuda.synchronize()
    ms_e2e = f0.elapsed_time(f1)
    barrier()

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
    fl = torch.tensor([flops_step], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(fl, op=dist.ReduceOp.SUM)
    ms, ms_e2e = float(t[0]), float(t[1])
    total_solves_step = tpg * world
    value = total_solves_step * args.steps / (ms * 1e-3)
    e2e_value = total_solves_step * e2e_steps / (ms_e2e * 1e-3)
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

    if N_SAMPLES != 2048:
        traffic = None     # the committed ncu capture is of the N=2048 kernel
    out = None
    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "tracks_per_gpu": tpg, "tracks_total": total_solves_step,
                       "n_samples": N_SAMPLES, "m_per_ring": M_PER_RING, "seed_base": SEED_BASE,
                       "l2": "inputs larger than L2 (%.0f MB per GPU)" % (pb.h2d_bytes / 1e6), "parallelism": f"dp{world}",
                       "mean_lap_s": lap_mean, "gen_s": round(gen_s, 2), "kernel_diag": diag},
            "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved_tf / fp64_peak if fp64_peak else None, "traffic": traffic,
                         "peak_source": "DFMA throughput measured in this run (rl_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 entry",
                         "flops_per_launch": float(fl[0]) / world, "kernel_ms": kernel_ms,
                         "hbm_achieved_gbs": alg_bytes / (kernel_ms * 1e-3) / 1e9, "hbm_peak_gbs": peaks.get("hbm_gbs")},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": pb.h2d_bytes, "d2h_bytes_per_step": pb.d2h_bytes,
                    "steps": e2e_steps},
            "gpu_launches": launches,
            "clocks": clk.summary(),
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = host_cores()
            n_ref = min(tpg, len(cores) * args.cpu_tracks_per_core)
            sps, wall, kind, solver_ms = run_reference_cpu(center, seg, L, n_ref, cores)
            out["cpu_baseline"] = {"value": sps, "unit": UNIT, "cores": min(len(cores), n_ref), "kind": kind,
                                   "sample": f"first {n_ref} tracks of the GPU batch, both stages, wall {wall:.1f} s "
                                             f"(sum of solver-only time {solver_ms / 1e3:.1f} core-s)"}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    dev.close(); ctx.close(); pool.close()
    if out is not None:
        args.emit(out)


def bench_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    cores = host_cores()
    per_step = len(cores) * args.ref_tracks_per_core
    center, seg, L, m = make_tracks(0, per_step)
    times, kind = [], "port"
    for it in range(args.warmup + args.steps):
        sps, wall, kind, _ = run_reference_cpu(center, seg, L, per_step, cores)
        if it >= args.warmup:
            times.append(wall)
    tot = sum(times)
    value = per_step * len(times) / tot
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_samples": N_SAMPLES, "m_per_ring": M_PER_RING,
                   "seed_base": SEED_BASE, "tracks_per_step": per_step},
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
    ap.add_argument("--tracks-per-gpu", type=int, default=8192)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-tracks-per-core", type=int, default=4)
    ap.add_argument("--ref-tracks-per-core", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-baseline-long", action="store_true")
    ap.add_argument("--workload", default="n2048", choices=["n2048", "n16384"],
                    help="n2048 = BASELINE configs[3] (the metric's configuration); n16384 = configs[4], the long-track stress "
                         "(one 8-CTA cluster per job; 1024 tracks per GPU = 8192 at 8 GPUs)")
    args = ap.parse_args()
    global N_SAMPLES, M_PER_RING, WORKLOAD
    if args.workload == "n16384":
        N_SAMPLES, M_PER_RING = 16384, 7447
        WORKLOAD = ("BASELINE configs[4]: long-track stress, synthetic closed tracks, N=16384 samples, M=7447 cones/ring, "
                    "default Config, min-curv + min-time per track, one 8-CTA cluster per job")
        if args.tracks_per_gpu == 8192:
            args.tracks_per_gpu = 1024
        if not args.cpu_baseline_long:
            args.no_cpu_baseline = True      # ~3 core-minutes per solve on the CPU: opt in with --cpu-baseline-long
        args.cpu_tracks_per_core = 1
    with StdoutOnlyJson() as out:
        args.emit = out.emit
        if args.impl == "reference":
            bench_reference(args)
        else:
            bench_ours(args)


if __name__ == "__main__":
    main()
