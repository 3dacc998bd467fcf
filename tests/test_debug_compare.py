"""The reference's debug comparison (pipeline::compute_mintime_and_save, main.cpp:1440-1593):
the lap of the centre line and of the min-curvature path under the same dynamics (RL_STAGE_EVAL jobs: heading/curvature
of a GIVEN path + the v(s) profile, main.cpp:1464-1477) and the 20 columns of <base>_debug_compare_paths.csv.
Goldens: `ref_harness debug` = the CSV the reference writes itself + the laps recomputed through its own functions."""
import numpy as np
import pytest

import practice_path_planning_for_formula_student_driverless_b200 as rl
from conftest import TOL_LAP_REL, load_golden
from oracle import oracle

CASES = ["debug_training_map", "debug_competition_map3"]
EMPTY = np.zeros((0, 4))


@pytest.mark.parametrize("name", CASES)
def test_eval_stage_oracle_bitwise(name):
    g = load_golden(name)
    c = oracle.solve(rl.RL_STAGE_EVAL, g["center_xy"], EMPTY, EMPTY, g["L"], True)
    m = oracle.solve(rl.RL_STAGE_EVAL, g["mc_xy"], EMPTY, EMPTY, g["L_mc"], True)
    assert c["lap_time"] == g["lap_center"] and m["lap_time"] == g["lap_mincurv"]
    assert np.array_equal(c["xy"], g["center_xy"]) and np.all(c["alpha_total"] == 0.0)
    assert abs(rl.path_length(g["mc_xy"], True) - g["L_mc"]) <= 1e-12 * g["L_mc"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_debug_compare_vs_reference_csv(ctx, name):
    g = load_golden(name)
    cfg = rl.Config()
    tr = rl.Track(g["center_xy"], g["inner_seg"], g["outer_seg"], g["L"])
    mt = rl.solve_batch([tr], [cfg], [(0, 0, rl.RL_STAGE_MINTIME)], ctx=ctx)[0]
    d = rl.debug_compare_paths(g["center_xy"], mt, g["mc_xy"], g["L"], g["s0"], True, cfg, ctx)
    assert abs(d["lap_center"] - g["lap_center"]) <= TOL_LAP_REL * g["lap_center"]
    assert abs(d["lap_mincurv"] - g["lap_mincurv"]) <= TOL_LAP_REL * g["lap_mincurv"]
    assert abs(d["lap_mintime"] - g["lap_mintime"]) <= TOL_LAP_REL * g["lap_mintime"]
    ref = g["columns"]
    assert ref.shape == (g["rows"], 20) and len(rl.solver.DEBUG_COLUMNS) == 20
    # the CSV has 9 decimals; positions/offsets follow the alpha tolerance, dynamics columns the v / kappa tolerances
    tol = {"s": 1e-8, "cx": 1e-8, "cy": 1e-8, "mt_x": 1e-4, "mt_y": 1e-4, "mc_x": 1e-8, "mc_y": 1e-8, "d_mt_signed_m": 1e-4,
           "d_mc_signed_m": 1e-8, "d_mt_abs_m": 1e-4, "d_mc_abs_m": 1e-8, "kappa_mt": 1e-6, "v_mt": 1e-4, "ax_mt": 2e-3,
           "alat_mt": 2e-3, "alat_ratio": 1e-3, "gamma": 1e-4, "a_acc_cap": 2e-3, "a_brk_cap": 2e-3, "a_power_cap": 2e-3}
    for c, k in enumerate(rl.solver.DEBUG_COLUMNS):
        err = float(np.max(np.abs(d["columns"][k] - ref[:, c])))
        assert err <= tol[k], (name, k, err)
    assert d["mt_offset_max"] == pytest.approx(float(np.max(ref[:, 9])), abs=1e-4)
    assert d["lap_gain_vs_mincurv_pct"] == pytest.approx((g["lap_mincurv"] - g["lap_mintime"]) / g["lap_mincurv"] * 100.0, abs=1e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [5, 300, 2048, 5000])
def test_eval_stage_vs_oracle(ctx, n, monkeypatch):
    """RL_STAGE_EVAL on every kernel family (single CTA, cluster), batched with optimisation jobs on the same track"""
    center, seg, L, m = rl.synth_tracks(1, max(n, 16), seed_base=0xEA0 + n)
    center, seg = center.reshape(-1, 2)[:n], seg.reshape(2, m, 4)
    tr = rl.Track(center, seg[0], seg[1], L[0] * n / max(n, 16))
    cfg = rl.Config()
    jobs = [(0, 0, rl.RL_STAGE_MINCURV), (0, 0, rl.RL_STAGE_EVAL), (0, 0, rl.RL_STAGE_MINTIME)] if n <= 2048 else [(0, 0, rl.RL_STAGE_EVAL)]
    res = rl.solve_batch([tr], [cfg], jobs, ctx=ctx)
    r = res[1] if n <= 2048 else res[0]
    o = oracle.solve(rl.RL_STAGE_EVAL, tr.center_xy, tr.inner_seg, tr.outer_seg, tr.L, True, cfg.to_params())
    assert np.array_equal(r.raceline, tr.center_xy) and np.all(r.alpha_total == 0.0)
    assert np.max(np.abs(r.curvature - o["curvature"])) <= 1e-9 and np.max(np.abs(r.v - o["v"])) <= 1e-6
    assert abs(r.lap_time - o["lap_time"]) <= TOL_LAP_REL * o["lap_time"]
    if n <= 2048:   # the jobs around it are unaffected by sharing a chain with an EVAL job
        o2 = oracle.solve(rl.RL_STAGE_MINTIME, tr.center_xy, tr.inner_seg, tr.outer_seg, tr.L, True, cfg.to_params())
        assert abs(res[2].lap_time - o2["lap_time"]) <= TOL_LAP_REL * o2["lap_time"]
        assert res[2].stats.accepted == o2["stats"].accepted
