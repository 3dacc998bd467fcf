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
    a = ap.parse_args()
    center, seg, L, m = rl.synth_tracks(a.tracks, a.n, a.m)
    jobs = np.empty((2 * a.tracks, 3), dtype=np.int64)
    jobs[0::2, 0] = jobs[1::2, 0] = np.arange(a.tracks)
    jobs[:, 1] = 0
    jobs[0::2, 2], jobs[1::2, 2] = rl.RL_STAGE_MINCURV, rl.RL_STAGE_MINTIME
    pb = rl.PackedBatch.from_arrays(np.arange(a.tracks + 1, dtype=np.int64) * a.n, np.arange(2 * a.tracks + 1, dtype=np.int64) * m,
                                    center, seg, L, np.ones(a.tracks, np.int32), [rl.Config().to_params()], jobs)
    ctx = rl.Context(0)
    ctx.solve_batch(pb)
    for stage, sl in (("min-curv", slice(0, None, 2)), ("min-time", slice(1, None, 2))):
        idx = range(pb.n_jobs)[sl]
        ph = np.array([[pb.out_stats[j].J0[16 + i] for i in range(9)] for j in idx])
        tot = ph.sum(axis=1).mean()
        print(f"{stage}: mean {tot / 1e6:.3f} Mcycles per job, evals {np.mean([pb.out_stats[j].evals for j in idx]):.0f}")
        for i, nm in enumerate(NAMES):
            print(f"   {nm:38s} {100 * ph[:, i].mean() / tot:6.2f} %   {ph[:, i].mean() / 1e3:9.1f} kcycles")
    ctx.close()


if __name__ == "__main__":
    main()
