#!/usr/bin/env python
"""Development: locate where the GPU result leaves the oracle's on one adversarial track (tests/trackgen.py)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import practice_path_planning_for_formula_student_driverless_b200 as rl
import trackgen
from oracle import oracle

kind, n, seed = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
gkw = eval(sys.argv[4]) if len(sys.argv) > 4 else {}
stage = int(sys.argv[5]) if len(sys.argv) > 5 else 1
c, inner, outer, L = trackgen.make_track(seed, n, kind, **gkw)
ctx = rl.Context(0)
print("case", kind, n, seed, gkw, "stage", stage)
prev_ok = None
for k in range(1, 15):
    cfg = rl.Config(max_outer_iters=k)
    tr = rl.Track(c, inner, outer, L)
    r = rl.solve_batch([tr], [cfg], [(0, 0, stage)], ctx=ctx)[0]
    o = oracle.solve(stage, c, inner, outer, L, True, cfg.to_params())
    da = np.abs(r.alpha_total - o["alpha_total"])
    i = int(np.argmax(da))
    print(f"outers {k:2d}: max|dalpha_total| {da.max():.3e} at sample {i}  J0[k-1] gpu {r.stats.J0[k-1]:.17g} orc {o['stats'].J0[k-1]:.17g}  "
          f"Jend gpu {r.stats.Jend[k-1]:.17g} orc {o['stats'].Jend[k-1]:.17g}  bt {r.stats.bt_outer[k-1]}/{o['stats'].bt_outer[k-1]} acc {r.stats.acc_outer[k-1]}/{o['stats'].acc_outer[k-1]}")
    if da.max() > 1e-8 and prev_ok is not None:
        # the corridor of outer k-1 (0-based) is built from the path after k-1 outers = prev_ok's path
        P = prev_ok["xy"]
        nrm = oracle.normals(P, True)
        guard = cfg.veh_width_m * 0.5 + cfg.safety_margin_m
        lo, hi = oracle.corridor(P, nrm, inner, outer, guard)
        al_g, al_o = r.alpha_last, o["alpha_last"]
        viol = np.where((al_g > hi + 1e-12) | (al_g < lo - 1e-12))[0]
        print("   samples where the GPU's last alpha leaves the oracle's box:", viol[:20], "count", len(viol))
        for j in viol[:6]:
            print(f"     sample {j}: lo {lo[j]:.9f} hi {hi[j]:.9f} gpu alpha_last {al_g[j]:.9f} oracle {al_o[j]:.9f}")
        atb = np.where((np.abs(al_o - hi) < 1e-12) | (np.abs(al_o - lo) < 1e-12))[0]
        dd = np.abs(al_g - al_o)
        print("   oracle samples at a bound:", len(atb), " max |dalpha_last| among them %.3e, elsewhere %.3e" % (dd[atb].max() if len(atb) else 0, np.delete(dd, atb).max()))
        break
    if da.max() <= 1e-8:
        prev_ok = o
ctx.close()
