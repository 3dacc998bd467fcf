#!/usr/bin/env python
"""Development: per-launch device time of the BASELINE configs[2] sweep (one-warp class), with SM clocks."""
import os, subprocess, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import practice_path_planning_for_formula_student_driverless_b200 as rl

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
g = dict(np.load(os.path.join(root, "tests", "golden", "competition_map2.npz")))
tr = rl.Track(g["center_xy"], g["inner_seg"], g["outer_seg"], float(g["L"]))
lam = np.geomspace(4e-4, 6.4e-3, 8); mus = np.linspace(1.15, 1.6, 8); pw = np.linspace(20e3, 80e3, 8); wt = np.linspace(0.0, 3.5, 8)
cfgs = [rl.Config(lambda_smooth=float(x), mu=float(b), a_total_max=float(b) * 9.81, P_max_W=float(c), w_time_gain=float(d))
        for x in lam for b in mus for c in pw for d in wt]
jobs = np.array([[0, k, st] for k in range(len(cfgs)) for st in (1, 2)], dtype=np.int64)
pb = rl.PackedBatch([tr], [c.to_params() for c in cfgs], jobs)
ctx = rl.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
for chain in (int(a) for a in (sys.argv[1:] or ["0"])):
    ctx.set_option("force_chain", chain)
    dev = rl.DeviceBatch(ctx, pb)
    ms = []
    for it in range(12):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); dev.solve(); e1.record(stream); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for it in range(10):
        dev.solve()
    e1.record(stream); torch.cuda.synchronize()
    clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
    dev.download(); dev.sync()
    ev = np.mean([pb.out_stats[j].evals for j in range(pb.n_jobs)]); vr = np.mean([pb.out_stats[j].vpass_rounds for j in range(1, pb.n_jobs, 2)])
    print(f"force_chain {chain}: launches per solve {dev.launches_per_solve}; single launches ms {[round(x, 1) for x in ms]}; 10 back to back: {e0.elapsed_time(e1) / 10:.1f} ms each; "
          f"{clk}; evals/job {ev:.0f}, vpass rounds/MT job {vr:.0f}")
    dev.close()
