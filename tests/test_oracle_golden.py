"""CPU tests: the oracle (C restatement) against the golden vectors generated from the reference itself."""
import numpy as np
import pytest

from conftest import MAPS, load_golden
from oracle import oracle
from practice_path_planning_for_formula_student_driverless_b200 import RL_STAGE_MINCURV, RL_STAGE_MINTIME

# known answers of BASELINE.md section 2, recomputed with the full-precision L the front end produces
LAPS = {"training_map": 30.0525156, "competition_map1": 24.4126773, "competition_map2": 36.0096013,
        "competition_map3": 28.3506013, "competition_map_testday1": 23.7601223,
        "competition_map_testday2": 26.2850779, "competition_map_testday3": 36.9488450}
SHAPES = {"training_map": (216, 98), "competition_map1": (187, 85), "competition_map2": (252, 115),
          "competition_map3": (197, 90), "competition_map_testday1": (259, 119),
          "competition_map_testday2": (199, 91), "competition_map_testday3": (261, 118)}


@pytest.mark.parametrize("name", MAPS)
def test_golden_shapes_and_laps(name):
    g = load_golden(name)
    assert (g["n"], g["m_inner"]) == SHAPES[name] and g["m_inner"] == g["m_outer"]
    assert abs(g["mt_lap_time"] - LAPS[name]) < 5e-7
    # every outer iteration ran the full 120 accepted steps on the shipped maps
    assert g["mc_bt"].size == 1680 and g["mt_bt"].size == 1680
    assert 13 <= int(g["mc_bt"].sum()) <= 26 and 28 <= int(g["mt_bt"].sum()) <= 42


@pytest.mark.parametrize("name", MAPS + ["competition_map2_n1000"])
def test_oracle_matches_reference_bitwise(name):
    g = load_golden(name)
    p = oracle.default_params()
    for stage, pre, extra in ((RL_STAGE_MINCURV, "mc_", ()), (RL_STAGE_MINTIME, "mt_", ("v", "ax"))):
        r = oracle.solve(stage, g["center_xy"], g["inner_seg"], g["outer_seg"], g["L"], g["closed"], p)
        for k in ("xy", "heading", "curvature", "alpha_total", "alpha_last") + extra:
            assert np.array_equal(r[k], g[pre + k]), (name, pre + k, np.abs(r[k] - g[pre + k]).max())
        bt = g[pre + "bt"]
        assert r["stats"].accepted == bt.size and r["stats"].backtracks == int(bt.sum())
    assert r["lap_time"] == g["mt_lap_time"]


def test_oracle_matches_reference_on_config_sweep():
    base = load_golden("competition_map2")
    sw = load_golden("sweep_competition_map2")
    from oracle.batchfile import PARAM_FIELDS
    for j, (t, pi, stage) in enumerate(sw["jobs"]):
        p = oracle.default_params()
        for name, val in zip(PARAM_FIELDS, sw["params_rows"][pi]):
            cur = getattr(p, name)
            setattr(p, name, int(val) if isinstance(cur, int) else float(val))
        r = oracle.solve(int(stage), base["center_xy"], base["inner_seg"], base["outer_seg"], base["L"], True, p)
        for k in ("xy", "alpha_total", "curvature") + (("v",) if stage == RL_STAGE_MINTIME else ()):
            assert np.array_equal(r[k], sw[f"j{j}_{k}"]), (j, k)
        assert r["stats"].accepted == sw[f"j{j}_accepted"] and r["stats"].backtracks == sw[f"j{j}_backtracks"]


def test_eval_cost_grad_is_gradient_of_cost():
    """finite-difference check of the restated eval_cost_grad (main.cpp:654-675 / 866-895)."""
    rng = np.random.default_rng(0)
    n, h, lam = 40, 1.7, 1.6e-3
    A1, A2, N0 = rng.normal(size=n) * 0.1, 1 + 0.1 * rng.normal(size=n), rng.normal(size=n) * 0.05
    W = 1 + 0.05 * rng.normal(size=n)
    g2 = 1 + rng.random(n)
    a = rng.normal(size=n) * 0.2
    for closed in (True, False):
        for gam in (None, g2):
            J, grad = oracle.eval_cost_grad(A1, A2, N0, W, gam, h, lam, a, closed)
            for i in (0, 1, 7, n - 2, n - 1):
                e = np.zeros(n); e[i] = 1e-6
                Jp, _ = oracle.eval_cost_grad(A1, A2, N0, W, gam, h, lam, a + e, closed)
                Jm, _ = oracle.eval_cost_grad(A1, A2, N0, W, gam, h, lam, a - e, closed)
                assert abs((Jp - Jm) / 2e-6 - grad[i]) < 1e-6 * max(1.0, abs(grad[i]))


def test_velocity_profile_properties():
    p = oracle.default_params()
    rng = np.random.default_rng(1)
    kappa = 0.15 * np.sin(np.linspace(0, 12 * np.pi, 300, endpoint=False)) + 0.01 * rng.normal(size=300)
    v, ax, lap = oracle.velocity_profile(p, kappa, 1.7, True)
    vk = np.minimum(p.v_cap_mps, np.sqrt(p.a_lat_max / np.maximum(np.abs(kappa), p.kappa_eps)))
    assert np.all(v <= vk + 1e-12) and np.all(v > 0)
    assert abs(lap - np.sum(1.7 / v)) < 1e-9
    # zero iterations leave the curvature-limited profile
    p.max_vpass_iters = 0
    v0, _, _ = oracle.velocity_profile(p, kappa, 1.7, True)
    assert np.array_equal(v0, vk)


def test_corridor_fallback_when_ray_misses():
    """a ring the ray never hits falls back to the nearest-segment distance (main.cpp:696)."""
    ring = np.array([[2.0, -1, 2, 1], [2, 1, 3, 1], [3, 1, 3, -1], [3, -1, 2, -1]])   # box to the right
    far = np.array([[-50.0, -60, -50, 60]])                                           # wall far to the left
    lo, hi = oracle.corridor(np.array([[0.0, 0.0]]), np.array([[1.0, 0.0]]), ring, far, 0.5)
    assert abs(hi[0] - 1.5) < 1e-12          # +n hits the box at 2
    assert abs(lo[0] + 1.5) < 1e-12          # -n misses the box: nearest distance 2 < wall at 50


@pytest.mark.parametrize("name", ["open_competition_map1", "open_training_map"])
def test_oracle_matches_reference_bitwise_open_track(name):
    """open-track mode (DiffOpsOpen main.cpp:560-579, one-sided normals/derivatives, no wrap in the v(s) passes)."""
    g = load_golden(name)
    p = oracle.default_params()
    for stage, pre, extra in ((RL_STAGE_MINCURV, "mc_", ()), (RL_STAGE_MINTIME, "mt_", ("v", "ax"))):
        r = oracle.solve(stage, g["center_xy"], g["inner_seg"], g["outer_seg"], g["L"], False, p)
        for k in ("xy", "heading", "curvature", "alpha_total", "alpha_last") + extra:
            assert np.array_equal(r[k], g[pre + k]), (name, pre + k)
        assert r["stats"].accepted == g[pre + "accepted"] and r["stats"].backtracks == g[pre + "backtracks"]
    assert r["lap_time"] == g["mt_lap_time"]
