"""Job chains: one CTA (or one cluster) works through consecutive jobs on the same track and keeps what the corridor code
learnt (anchors, clearances, parity bits, certificates), so every job after the first starts with a corridor UPDATE.
The host only forms chains in large batches; the "force_chain" option (rl_set_option) makes it form them in small ones so that the chained path
is checked against the oracle and, bit for bit, against the unchained path."""
import numpy as np
import pytest

import practice_path_planning_for_formula_student_driverless_b200 as rl
from conftest import MAPS, TOL_LAP_REL, assert_result_close, load_golden
from test_gpu_parity import MC, MT, oracle_ref, stalled, track_of

pytestmark = pytest.mark.gpu


def _same_bits(a, b):
    assert np.array_equal(a.raceline, b.raceline) and np.array_equal(a.alpha_total, b.alpha_total)
    assert np.array_equal(a.curvature, b.curvature) and np.array_equal(a.heading, b.heading)
    assert a.stats.accepted == b.stats.accepted and a.stats.backtracks == b.stats.backtracks and a.stats.evals == b.stats.evals
    if a.v is not None and b.v is not None:
        assert np.array_equal(a.v, b.v) and a.lap_time == b.lap_time


def test_chained_shipped_maps(ctx, goldens):
    """MC + MT + a second Config per map in one chain of three jobs; the reference goldens and the unchained bits"""
    tracks = [track_of(goldens[n]) for n in MAPS]
    cfgs = [rl.Config(), rl.Config(lambda_smooth=3.2e-3, w_time_gain=2.0, veh_width_m=1.2)]
    jobs = [(t, c, st) for t in range(len(tracks)) for (c, st) in ((0, MC), (0, MT), (1, MT))]
    plain = rl.solve_batch(tracks, cfgs, jobs, ctx=ctx)
    ctx.set_option("force_chain", 3)
    chained = rl.solve_batch(tracks, cfgs, jobs, ctx=ctx)
    ctx.set_option("force_chain", 0)
    for (t, c, st), a, b in zip(jobs, chained, plain):
        _same_bits(a, b)
        if c == 0:
            g, pre = goldens[MAPS[t]], ("mc_" if st == MC else "mt_")
            assert_result_close(a, g, pre, st == MT, tag=(MAPS[t], pre, "chained"))
            assert a.stats.backtracks == int(g[pre + "bt"].sum())
    # jobs after the first of a chain did not search the rings again: far fewer exact ray tests
    assert chained[1].stats.ray_tests < plain[1].stats.ray_tests


@pytest.mark.parametrize("n", [40, 300, 1500, 2048])
def test_chained_synthetic_vs_oracle(ctx, n):
    center, seg, L, m = rl.synth_tracks(2, n, seed_base=0xC4A1 + n)
    center, seg = center.reshape(2, n, 2), seg.reshape(2, 2, m, 4)
    tracks = [rl.Track(center[i], seg[i, 0], seg[i, 1], L[i]) for i in range(2)]
    cfgs = [rl.Config(), rl.Config(P_max_W=30000.0, safety_margin_m=0.2)]
    jobs = [(0, 0, MC), (0, 0, MT), (0, 1, MC), (0, 1, MT), (1, 0, MT), (1, 1, MT)]
    ctx.set_option("force_chain", 4)
    res = rl.solve_batch(tracks, cfgs, jobs, ctx=ctx)
    ctx.set_option("force_chain", 0)
    for (t, c, st), r in zip(jobs, res):
        o = oracle_ref(st, tracks[t], cfgs[c].to_params())
        assert_result_close(r, o, "o_", st == MT, tag=("chain", n, t, c, st))
        assert r.stats.accepted == o["stats"].accepted, (n, t, c, st)
        if not stalled(o["stats"]):
            assert r.stats.backtracks == o["stats"].backtracks, (n, t, c, st)
        if st == MT:
            assert abs(r.lap_time - o["lap"]) <= TOL_LAP_REL * o["lap"]


def test_chained_open_track(ctx):
    g = load_golden("open_competition_map1")
    tr = rl.Track(g["center_xy"], g["inner_seg"], g["outer_seg"], g["L"], closed=False)
    ctx.set_option("force_chain", 2)
    res = rl.solve_batch([tr], [rl.Config()], [(0, 0, MC), (0, 0, MT)], ctx=ctx)
    ctx.set_option("force_chain", 0)
    for st, pre, r in ((MC, "mc_", res[0]), (MT, "mt_", res[1])):
        assert_result_close(r, g, pre, st == MT, tag=("open chained", pre))
        assert r.stats.accepted == g[pre + "accepted"] and r.stats.backtracks == g[pre + "backtracks"]


@pytest.mark.parametrize("cs,n", [(2, 1100), (4, 3000)])
def test_chained_cluster_vs_oracle(ctx, cs, n):
    """the cluster kernel's chains: MC then MT of the same long track in one cluster"""
    center, seg, L, m = rl.synth_tracks(1, n, seed_base=0xC4B2 + n)
    tr = rl.Track(center.reshape(n, 2), seg.reshape(2, m, 4)[0], seg.reshape(2, m, 4)[1], L[0])
    cfg = rl.Config()
    jobs = [(0, 0, MC), (0, 0, MT)]
    ctx.set_option("force_cluster", cs)
    plain = rl.solve_batch([tr], [cfg], jobs, ctx=ctx)
    ctx.set_option("force_chain", 2)
    chained = rl.solve_batch([tr], [cfg], jobs, ctx=ctx)
    ctx.set_option("force_chain", 0)
    ctx.set_option("force_cluster", 0)
    for (t, c, st), a, b in zip(jobs, chained, plain):
        _same_bits(a, b)
        o = oracle_ref(st, tr, cfg.to_params())
        assert_result_close(a, o, "o_", st == MT, tag=("cluster chain", cs, n, st))
        assert a.stats.accepted == o["stats"].accepted
    assert chained[1].stats.ray_tests < plain[1].stats.ray_tests


@pytest.mark.parametrize("n,cs", [(700, 0), (1100, 2), (4100, 8)])
def test_ring_order_does_not_matter(ctx, n, cs):
    """rings rotated to another start cone and traversed the other way round (still vertex chains): anchors, windows and
    chunk-local vertex ranges work on ring indices, the results must not care"""
    center, seg, L, m = rl.synth_tracks(1, n, seed_base=0xC4C3 + n)
    seg = seg.reshape(2, m, 4)
    inner = np.roll(seg[0], -(m // 3), axis=0)                              # another start cone
    outer = seg[1][::-1][:, [2, 3, 0, 1]].copy()                            # the other way round: segment i = (end, start) reversed
    tr0 = rl.Track(center.reshape(n, 2), seg[0], seg[1], L[0])
    tr1 = rl.Track(center.reshape(n, 2), inner, outer, L[0])
    cfg = rl.Config()
    if cs:
        ctx.set_option("force_cluster", cs)
    res = rl.solve_batch([tr0, tr1], [cfg], [(0, 0, MT), (1, 0, MT), (1, 0, MC)], ctx=ctx)
    if cs:
        ctx.set_option("force_cluster", 0)
    o = oracle_ref(MT, tr1, cfg.to_params())
    assert_result_close(res[1], o, "o_", True, tag=("ring order", n, cs))
    assert res[1].stats.accepted == o["stats"].accepted and res[1].stats.backtracks == o["stats"].backtracks
    # same geometry, same answer (the searches take minima over the same set of segments)
    assert np.max(np.abs(res[0].alpha_total - res[1].alpha_total)) < 1e-9
    assert abs(res[0].lap_time - res[1].lap_time) <= 1e-9 * res[0].lap_time
    omc = oracle_ref(MC, tr1, cfg.to_params())
    assert_result_close(res[2], omc, "o_", False, tag=("ring order mc", n, cs))


@pytest.mark.parametrize("n,cs", [(300, 0), (700, 0), (1100, 2)])
def test_segment_soup_rings(ctx, n, cs):
    """rings as an unordered set of segments (no vertex chain): the update path and the parity shortcut do not apply,
    the searching path answers every build; chained with a second job"""
    center, seg, L, m = rl.synth_tracks(1, n, seed_base=0xC4D4 + n)
    seg = seg.reshape(2, m, 4)
    rng = np.random.default_rng(n)
    tr = rl.Track(center.reshape(n, 2), seg[0][rng.permutation(m)], seg[1][rng.permutation(m)], L[0])
    cfg = rl.Config()
    if cs:
        ctx.set_option("force_cluster", cs)
    ctx.set_option("force_chain", 2)
    res = rl.solve_batch([tr], [cfg], [(0, 0, MC), (0, 0, MT)], ctx=ctx)
    ctx.set_option("force_chain", 0)
    if cs:
        ctx.set_option("force_cluster", 0)
    for st, r in zip((MC, MT), res):
        o = oracle_ref(st, tr, cfg.to_params())
        assert_result_close(r, o, "o_", st == MT, tag=("soup", n, cs, st))
        assert r.stats.accepted == o["stats"].accepted
        if not stalled(o["stats"]):
            assert r.stats.backtracks == o["stats"].backtracks


@pytest.mark.parametrize("n,m", [(300, 400), (2048, 1400)])
def test_rings_larger_than_one_tile(ctx, n, m):
    """more cones per ring than one shared-memory tile of the job's size class holds: the tiled corridor path of the
    single-CTA kernel (no per-sample state), also as the second job of a chain"""
    center, seg, L, mm = rl.synth_tracks(1, n, m_per_ring=m, seed_base=0xC4E5 + n)
    assert mm == m
    seg = seg.reshape(2, m, 4)
    tr = rl.Track(center.reshape(n, 2), seg[0], seg[1], L[0])
    cfg = rl.Config()
    ctx.set_option("force_chain", 2)
    res = rl.solve_batch([tr], [cfg], [(0, 0, MC), (0, 0, MT)], ctx=ctx)
    ctx.set_option("force_chain", 0)
    for st, r in zip((MC, MT), res):
        o = oracle_ref(st, tr, cfg.to_params())
        assert_result_close(r, o, "o_", st == MT, tag=("tiled", n, m, st))
        assert r.stats.accepted == o["stats"].accepted
        if not stalled(o["stats"]):
            assert r.stats.backtracks == o["stats"].backtracks


@pytest.mark.parametrize("cs", [0, 2])
def test_negative_guard_disables_parity_shortcut(ctx, cs):
    """veh_width/2 + safety_margin < 0 lets the path cross a ring, so 'inside a closed ring => every ray hits it' may not
    be used (parity_ok); results still follow the reference"""
    n = 1100 if cs else 600
    center, seg, L, m = rl.synth_tracks(1, n, seed_base=0xC4F6 + n)
    seg = seg.reshape(2, m, 4)
    tr = rl.Track(center.reshape(n, 2), seg[0], seg[1], L[0])
    cfg = rl.Config(safety_margin_m=-0.9, max_outer_iters=6)
    if cs:
        ctx.set_option("force_cluster", cs)
    ctx.set_option("force_chain", 2)
    res = rl.solve_batch([tr], [cfg], [(0, 0, MC), (0, 0, MT)], ctx=ctx)
    ctx.set_option("force_chain", 0)
    if cs:
        ctx.set_option("force_cluster", 0)
    for st, r in zip((MC, MT), res):
        o = oracle_ref(st, tr, cfg.to_params())
        assert_result_close(r, o, "o_", st == MT, tag=("neg guard", cs, st))
        assert r.stats.accepted == o["stats"].accepted
        if not stalled(o["stats"]):
            assert r.stats.backtracks == o["stats"].backtracks
