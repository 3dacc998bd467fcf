#!/usr/bin/env python
"""Development tool: time shares of the phases of solve_kernel.

    python -m practice_path_planning_for_formula_student_driverless_b200.build -D RL_PHASE_TIMERS --suffix _ph
    RL_LIB_VARIANT=_ph python tools/phase_report.py [--tracks 592] [--n 2048] [--m 931]

Needs the -DRL_PHASE_TIMERS build (thread 0 of every CTA accumulates clock64() cycles per phase into
stats.J0[16..24]); the product library carries no timers.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import practice_path_planning_for_formula_student_driverless_b200 as rl  # noqa: E402

NAMES = ["setup + first corridor", "linearisation + park", "v(s) + time weights", "stencil coefficients", "PGD loop",
         "path back + update", "corridor update pass", "corridor search fallback + staging", "final geometry / v(s) / stores"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tracks", type=int, default=592)
    ap.add_argument("--n", type=int, default=2048)
    ap.add_argument("--m", type=int, default=931)
    ap.add_argument("--sweep", action="store_true", help="BASELINE configs[2]: competition_map2 x 4096 Configs")
    ap.add_argument("--chain", type=int, default=0, help="rl_set_option('force_chain')")
    a = ap.parse_args()
    ctx = rl.Context(0)
    if a.chain:
        ctx.set_option("force_chain", a.chain)
    if a.sweep:
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        g = dict(np.load(os.path.join(root, "tests", "golden", "competition_map2.npz")))
        tr = rl.Track(g["center_xy"], g["inner_seg"], g["outer_seg"], float(g["L"]))
        lam = np.geomspace(4e-4, 6.4e-3, 8); mus = np.linspace(1.15, 1.6, 8); pw = np.linspace(20e3, 80e3, 8); wt = np.linspace(0.0, 3.5, 8)
        cfgs = [rl.Config(lambda_smooth=float(x), mu=float(b), a_total_max=float(b) * 9.81, P_max_W=float(c), w_time_gain=float(d))
                for x in lam for b in mus for c in pw for d in wt]
        jobs = np.array([[0, k, st] for k in range(len(cfgs)) for st in (1, 2)], dtype=np.int64)
        pb = rl.PackedBatch([tr], [c.to_params() for c in cfgs], jobs)
    else:
        center, seg, L, m = rl.synth_tracks(a.tracks, a.n, a.m)
        jobs = np.empty((2 * a.tracks, 3), dtype=np.int64)
        jobs[0::2, 0] = jobs[1::2, 0] = np.arange(a.tracks)
        jobs[:, 1] = 0
        jobs[0::2, 2], jobs[1::2, 2] = rl.RL_STAGE_MINCURV, rl.RL_STAGE_MINTIME
        pb = rl.PackedBatch.from_arrays(np.arange(a.tracks + 1, dtype=np.int64) * a.n, np.arange(2 * a.tracks + 1, dtype=np.int64) * m,
                                        center, seg, L, np.ones(a.tracks, np.int32), [rl.Config().to_params()], jobs)
    import time
    ctx.solve_batch(pb)
    t0 = time.perf_counter(); ctx.solve_batch(pb); dt = time.perf_counter() - t0
    print(f"one rl_solve_batch call: {dt * 1e3:.1f} ms for {pb.n_jobs} jobs; vpass rounds per min-time job "
          f"{np.mean([pb.out_stats[j].vpass_rounds for j in range(1, pb.n_jobs, 2)]):.0f}")
    for stage, sl in (("min-curv", slice(0, None, 2)), ("min-time", slice(1, None, 2))):
        idx = range(pb.n_jobs)[sl]
        ph = np.array([[pb.out_stats[j].J0[16 + i] for i in range(9)] for j in idx])
        tot = ph.sum(axis=1).mean()
        print(f"{stage}: mean {tot / 1e6:.3f} Mcycles per job, evals {np.mean([pb.out_stats[j].evals for j in idx]):.0f}")
        for i, nm in enumerate(NAMES):
            print(f"   {nm:38s} {100 * ph[:, i].mean() / tot:6.2f} %   {ph[:, i].mean() / 1e3:9.1f} kcycles")
    if a.n > 4096:   # cluster kernel: per-rank corridor times (update pass | fallback + staging | outers with a fallback)
        for stage, sl in (("min-curv", slice(0, None, 2)), ("min-time", slice(1, None, 2))):
            idx = range(pb.n_jobs)[sl]
            upd = np.array([[pb.out_stats[j].Jend[16 + r] for r in range(16)] for j in idx]).mean(axis=0)
            fbk = np.array([[pb.out_stats[j].lap_outer[16 + r] for r in range(16)] for j in idx]).mean(axis=0)
            nfb = np.array([[pb.out_stats[j].acc_outer[16 + r] for r in range(16)] for j in idx]).mean(axis=0)
            print(f"{stage}: per rank kcycles per job: update " + " ".join(f"{x / 1e3:.0f}" for x in upd[:8]))
            print(f"{stage}:                        fallback " + " ".join(f"{x / 1e3:.0f}" for x in fbk[:8]))
            print(f"{stage}:           outers with a fallback " + " ".join(f"{x:.1f}" for x in nfb[:8]))
            fewt = np.array([[pb.out_stats[j].Jend[24 + r] for r in range(8)] for j in idx]).mean(axis=0)
            fewn = np.array([[pb.out_stats[j].bt_outer[24 + r] for r in range(8)] for j in idx]).mean(axis=0)
            tile = np.array([[pb.out_stats[j].acc_outer[24 + r] for r in range(8)] for j in idx]).mean(axis=0)
            print(f"{stage}: few-sample search kcycles per job " + " ".join(f"{x / 1e3:.0f}" for x in fewt) + " | samples " + " ".join(f"{x:.1f}" for x in fewn)
                  + " | tile-streaming searches " + " ".join(f"{x:.2f}" for x in tile))
    if a.n > 4096:
        idx = range(pb.n_jobs)
        tf = np.array([[pb.out_stats[j].J0[25 + k] for k in range(5)] for j in idx]).mean(axis=0)
        print("rank 0 few-sample search, kcycles per job: setup %.0f | sweep 1 %.0f | sweep 2 %.0f | combine + state %.0f | cone %.0f" % tuple(tf / 1e3))
    ctx.close()


if __name__ == "__main__":
    main()
