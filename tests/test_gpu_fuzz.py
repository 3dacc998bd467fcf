"""Randomised parity sweep: random track shapes, sizes and Configs (fixed seeds), min-curv + min-time chained in one
CTA, against the pinned oracle.  Exercises the certificate / update / chain machinery of the corridor on many more
geometries than the hand-picked cases."""
import numpy as np
import pytest

import practice_path_planning_for_formula_student_driverless_b200 as rl
from conftest import TOL_LAP_REL, assert_result_close
from test_gpu_parity import MC, MT, oracle_ref, stalled

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("block", range(8))
def test_random_tracks_and_configs(ctx, block):
    rng = np.random.default_rng(0xF022 + block)
    tracks, cfgs, jobs = [], [], []
    for t in range(12):
        n = int(rng.integers(24, 900))
        m = int(max(8, round(n / rng.uniform(1.6, 3.0))))
        center, seg, L, mm = rl.synth_tracks(1, n, m_per_ring=m, seed_base=int(rng.integers(1, 1 << 30)))
        seg = seg.reshape(2, mm, 4)
        tracks.append(rl.Track(center.reshape(n, 2), seg[0], seg[1], L[0]))
        cfgs.append(rl.Config(lambda_smooth=float(10 ** rng.uniform(-3.6, -2.0)), safety_margin_m=float(rng.uniform(0.0, 0.45)),
                              veh_width_m=float(rng.uniform(0.7, 1.5)), w_time_gain=float(rng.uniform(0.0, 3.0)),
                              P_max_W=float(rng.uniform(15e3, 90e3)), step_init=float(rng.choice([0.3, 0.65, 1.5, 6.0])),
                              max_outer_iters=int(rng.integers(3, 15)), time_weight_use_inv_v=bool(rng.integers(0, 2))))
        jobs += [(t, t, MC), (t, t, MT)]
    ctx.set_option("force_chain", 2)
    res = rl.solve_batch(tracks, cfgs, jobs, ctx=ctx)
    ctx.set_option("force_chain", 0)
    for (t, c, st), r in zip(jobs, res):
        prm = cfgs[c].to_params()
        o = oracle_ref(st, tracks[t], prm)
        assert_result_close(r, o, "o_", st == MT, tag=("fuzz", block, t, st))
        assert r.stats.accepted == o["stats"].accepted, (block, t, st)
        if not stalled(o["stats"]):
            assert r.stats.backtracks == o["stats"].backtracks, (block, t, st)
        if st == MT:
            assert abs(r.lap_time - o["lap"]) <= TOL_LAP_REL * max(o["lap"], 1e-9)


@pytest.mark.parametrize("block", range(2))
def test_random_long_tracks_on_the_cluster_path(ctx, block):
    """the same sweep through the cluster kernel (2 CTAs per job), chained"""
    rng = np.random.default_rng(0xF0C2 + block)
    tracks, cfgs, jobs = [], [], []
    for t in range(5):
        n = int(rng.integers(1024, 2600))
        m = int(round(n / rng.uniform(1.8, 2.8)))
        center, seg, L, mm = rl.synth_tracks(1, n, m_per_ring=m, seed_base=int(rng.integers(1, 1 << 30)))
        seg = seg.reshape(2, mm, 4)
        tracks.append(rl.Track(center.reshape(n, 2), seg[0], seg[1], L[0]))
        cfgs.append(rl.Config(lambda_smooth=float(10 ** rng.uniform(-3.4, -2.2)), safety_margin_m=float(rng.uniform(0.0, 0.4)),
                              veh_width_m=float(rng.uniform(0.8, 1.4)), w_time_gain=float(rng.uniform(0.0, 2.5)),
                              max_outer_iters=int(rng.integers(3, 9))))
        jobs += [(t, t, MC), (t, t, MT)]
    ctx.set_option("force_cluster", 2)
    ctx.set_option("force_chain", 2)
    res = rl.solve_batch(tracks, cfgs, jobs, ctx=ctx)
    ctx.set_option("force_chain", 0)
    ctx.set_option("force_cluster", 0)
    for (t, c, st), r in zip(jobs, res):
        o = oracle_ref(st, tracks[t], cfgs[c].to_params())
        assert_result_close(r, o, "o_", st == MT, tag=("fuzz cluster", block, t, st))
        assert r.stats.accepted == o["stats"].accepted, (block, t, st)
        if not stalled(o["stats"]):
            assert r.stats.backtracks == o["stats"].backtracks, (block, t, st)


def test_random_open_tracks(ctx):
    """open paths: random arcs of random synthetic tracks with polyline rings, chained"""
    rng = np.random.default_rng(0xF0D3)
    tracks, cfgs, jobs = [], [], []
    for t in range(10):
        nn = int(rng.integers(60, 900))
        n = int(rng.integers(20, nn - 8))
        center, seg, L, m = rl.synth_tracks(1, nn, seed_base=int(rng.integers(1, 1 << 30)))
        center, seg = center.reshape(nn, 2), seg.reshape(2, m, 4)
        keep = max(3, int(m * n / nn))
        tracks.append(rl.Track(center[:n], rl.polyline_edges(seg[0, :keep + 2, :2]), rl.polyline_edges(seg[1, :keep + 2, :2]),
                               L[0] * n / nn, closed=False))
        cfgs.append(rl.Config(lambda_smooth=float(10 ** rng.uniform(-3.4, -2.2)), safety_margin_m=float(rng.uniform(0.0, 0.3)),
                              w_time_gain=float(rng.uniform(0.0, 2.5)), max_outer_iters=int(rng.integers(3, 15))))
        jobs += [(t, t, MC), (t, t, MT)]
    ctx.set_option("force_chain", 2)
    res = rl.solve_batch(tracks, cfgs, jobs, ctx=ctx)
    ctx.set_option("force_chain", 0)
    for (t, c, st), r in zip(jobs, res):
        o = oracle_ref(st, tracks[t], cfgs[c].to_params())
        assert_result_close(r, o, "o_", st == MT, tag=("fuzz open", t, st))
        assert r.stats.accepted == o["stats"].accepted, (t, st)
        if not stalled(o["stats"]):
            assert r.stats.backtracks == o["stats"].backtracks, (t, st)
