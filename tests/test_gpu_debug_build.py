"""The debug-checks build (libraceline_b200_dbg.so, -DRL_DEBUG_CHECKS) in place of compute-sanitizer, which is not
available on the pool: shared-memory phase hand-over checks, guard zones between the regions, bounds asserts
(csrc/raceline_kernels.cuh).  The build runs in a subprocess (one library per process): the shipped maps, synthetic
tracks of every size class, an open track and a cluster case must solve with ZERO failed checks and the usual parity;
with the fault injection on, the checker must fire."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

SCRIPT = r'''
import json, sys
import numpy as np
sys.path.insert(0, "tests")
import practice_path_planning_for_formula_student_driverless_b200 as rl
from conftest import MAPS, load_golden, assert_result_close
from oracle import oracle
MC, MT = rl.RL_STAGE_MINCURV, rl.RL_STAGE_MINTIME
ctx = rl.Context(0)
out = {}
# shipped maps, chained (T = 32 class), against the reference goldens
gs = [load_golden(m) for m in MAPS]
tracks = [rl.Track(g["center_xy"], g["inner_seg"], g["outer_seg"], g["L"]) for g in gs]
jobs = [(t, 0, st) for t in range(len(tracks)) for st in (MC, MT)]
ctx.set_option("force_chain", 2)
res = rl.solve_batch(tracks, [rl.Config()], jobs, ctx=ctx)
for (t, _, st), r in zip(jobs, res):
    pre = "mc_" if st == MC else "mt_"
    assert_result_close(r, gs[t], pre, st == MT, tag=(MAPS[t], pre, "debug build"))
    assert r.stats.backtracks == int(gs[t][pre + "bt"].sum())
# synthetic tracks of the other size classes (64 ... 512 threads), chained, against the oracle
for n in (300, 700, 1500, 2048, 3000):
    c, s, L, m = rl.synth_tracks(1, n, seed_base=0xDB60 + n)
    tr = rl.Track(c.reshape(n, 2), s.reshape(2, m, 4)[0], s.reshape(2, m, 4)[1], float(L[0]))
    rs = rl.solve_batch([tr], [rl.Config()], [(0, 0, MC), (0, 0, MT)], ctx=ctx)
    for st, r in zip((MC, MT), rs):
        o = oracle.solve(st, tr.center_xy, tr.inner_seg, tr.outer_seg, tr.L, True, rl.Config().to_params())
        assert np.max(np.abs(r.alpha_total - o["alpha_total"])) < 1e-4 and r.stats.accepted == o["stats"].accepted, (n, st)
# an open track and a forced 2-CTA cluster
g = load_golden("open_competition_map1")
r = rl.solve_batch([rl.Track(g["center_xy"], g["inner_seg"], g["outer_seg"], g["L"], closed=False)], [rl.Config()], [(0, 0, MC), (0, 0, MT)], ctx=ctx)
assert r[1].stats.accepted == g["mt_accepted"]
ctx.set_option("force_cluster", 2)
c, s, L, m = rl.synth_tracks(1, 1500, seed_base=0xDB99)
tr = rl.Track(c.reshape(1500, 2), s.reshape(2, m, 4)[0], s.reshape(2, m, 4)[1], float(L[0]))
rc = rl.solve_batch([tr], [rl.Config()], [(0, 0, MC), (0, 0, MT)], ctx=ctx)
o = oracle.solve(MT, tr.center_xy, tr.inner_seg, tr.outer_seg, tr.L, True, rl.Config().to_params())
assert np.max(np.abs(rc[1].alpha_total - o["alpha_total"])) < 1e-4
ctx.set_option("force_cluster", 0)
out["clean"] = ctx.debug_check_failures()
# fault injection: warp 0 skips the hand-over of the coefficient phase -> the checker must see it
ctx.set_option("debug_inject", 1)
c, s, L, m = rl.synth_tracks(1, 2048, seed_base=0xDB61)
tr = rl.Track(c.reshape(2048, 2), s.reshape(2, m, 4)[0], s.reshape(2, m, 4)[1], float(L[0]))
rl.solve_batch([tr], [rl.Config()], [(0, 0, MC)], ctx=ctx)
out["injected"] = ctx.debug_check_failures()
ctx.set_option("debug_inject", 0)
print("RESULT " + json.dumps(out))
'''


def test_debug_checks_build_is_clean_and_has_teeth():
    lib = os.path.join(ROOT, "practice_path_planning_for_formula_student_driverless_b200", "csrc", "libraceline_b200_dbg.so")
    if not os.path.exists(lib):
        pytest.skip("libraceline_b200_dbg.so not built (python -m ...build -D RL_DEBUG_CHECKS --suffix _dbg)")
    env = dict(os.environ, RL_LIB_VARIANT="_dbg")
    r = subprocess.run([sys.executable, "-c", SCRIPT], cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1]
    out = json.loads(line[7:])
    assert out["clean"][0] == 0, f"debug checks failed: count {out['clean'][0]}, first code {out['clean'][1]} in CTA {out['clean'][2]}"
    assert out["injected"][0] > 0 and 1000 <= out["injected"][1] < 2000, out


def test_product_build_carries_no_checks(ctx):
    import practice_path_planning_for_formula_student_driverless_b200 as rl
    with pytest.raises(rl.RacelineError) as e:
        ctx.debug_check_failures()
    assert e.value.status == rl.RL_ERR_UNSUPPORTED
