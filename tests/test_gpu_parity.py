"""GPU parity tests: the CUDA path, called through the C ABI, against the reference goldens and the oracle.

Tolerances are BASELINE.json's: alpha 1e-4 m, kappa 1e-6 1/m, v 1e-4 m/s, lap time 1e-5 relative, and
identical accepted-step / Armijo-backtrack counts on the shipped maps.
"""
import ctypes as C

import numpy as np
import pytest

import practice_path_planning_for_formula_student_driverless_b200 as rl
from conftest import MAPS, TOL_ALPHA, TOL_LAP_REL, assert_result_close, load_golden
from oracle import oracle
from oracle.batchfile import PARAM_FIELDS

pytestmark = pytest.mark.gpu

MC, MT = rl.RL_STAGE_MINCURV, rl.RL_STAGE_MINTIME


def track_of(g):
    return rl.Track(g["center_xy"], g["inner_seg"], g["outer_seg"], g["L"], bool(g["closed"]))


def oracle_ref(stage, tr, params):
    r = oracle.solve(stage, tr.center_xy, tr.inner_seg, tr.outer_seg, tr.L, tr.closed, params)
    d = {"o_" + k: r[k] for k in ("xy", "heading", "curvature", "alpha_total", "alpha_last", "v", "ax")}
    d["lap"] = r["lap_time"]
    d["stats"] = r["stats"]
    return d


def stalled(st):
    """True when some outer iteration of the oracle made no progress beyond rounding noise (e.g. the N=16
    synthetic track is an exact circle): the Armijo test then compares numbers that differ in the last bit
    and the backtrack count is not a property of the algorithm any more.  Every call is tallied; the totals are
    printed at the end of the session (conftest.pytest_terminal_summary)."""
    from conftest import STALLED_TALLY
    hit = any(abs(st.J0[o] - st.Jend[o]) <= 1e-12 * abs(st.J0[o]) for o in range(st.outer_done))
    STALLED_TALLY["asked"] += 1
    STALLED_TALLY["stalled"] += int(hit)
    return hit


def test_shipped_maps_batched_parity(ctx, goldens):
    """BASELINE config 2: all shipped maps batched on one GPU, every output column against the reference."""
    names = MAPS + ["competition_map2_n1000"]
    gs = [goldens.get(n) or load_golden(n) for n in names]
    tracks = [track_of(g) for g in gs]
    jobs = [(t, 0, st) for t in range(len(tracks)) for st in (MC, MT)]
    res = rl.solve_batch(tracks, [rl.Config()], jobs, ctx=ctx)
    worst = {}
    for (t, _, st), r in zip(jobs, res):
        g, pre = gs[t], ("mc_" if st == MC else "mt_")
        errs = assert_result_close(r, g, pre, st == MT, tag=(names[t], pre))
        for k, v in errs.items():
            worst[k] = max(worst.get(k, 0.0), v)
        bt = g[pre + "bt"]
        assert r.stats.status == 0 and r.stats.n == g["n"] and r.stats.outer_done == 14
        assert r.stats.accepted == bt.size, (names[t], pre, r.stats.accepted)
        assert r.stats.backtracks == int(bt.sum()), (names[t], pre, r.stats.backtracks, int(bt.sum()))
        per_outer = [int(r.stats.bt_outer[o]) for o in range(14)]
        assert per_outer == [int(bt[120 * o:120 * (o + 1)].sum()) for o in range(14)], (names[t], pre)
        assert r.stats.evals == 14 + bt.size + int(bt.sum())
        if st == MT:
            assert abs(r.lap_time - g["mt_lap_time"]) <= TOL_LAP_REL * g["mt_lap_time"]
    print("worst abs errors vs reference:", worst)


def test_single_problem_entry_points(ctx, goldens):
    """rl_compute_min_curvature_raceline / rl_compute_min_time_raceline: the reference's own signatures."""
    g = goldens["training_map"]
    cfg = rl.Config()
    r = rl.compute_min_curvature_raceline(g["center_xy"], g["inner_seg"], g["outer_seg"], cfg.veh_width_m, g["L"], True,
                                          cfg, ctx)
    assert_result_close(r, g, "mc_", False)
    r = rl.compute_min_time_raceline(g["center_xy"], g["inner_seg"], g["outer_seg"], cfg.veh_width_m, g["L"], True, cfg, ctx)
    assert_result_close(r, g, "mt_", True)
    assert abs(r.lap_time - g["mt_lap_time"]) <= TOL_LAP_REL * g["mt_lap_time"]
    # empty input -> empty Result, no error (main.cpp:689 / 912)
    e = rl.compute_min_time_raceline(np.zeros((0, 2)), g["inner_seg"], g["outer_seg"], 1.0, 1.0, True, cfg, ctx)
    assert e.raceline.shape == (0, 2) and e.lap_time == 0.0


def test_config_sweep_parity(ctx, goldens):
    """BASELINE config 3 spot check: off-default Configs (lambda, mu->a_total_max, P_max, w_time_gain, inv-v weights)."""
    base = goldens["competition_map2"]
    sw = load_golden("sweep_competition_map2")
    params = []
    for row in sw["params_rows"]:
        p = rl.RlParams()
        for name, val in zip(PARAM_FIELDS, row):
            cur = getattr(p, name)
            setattr(p, name, int(val) if isinstance(cur, int) else float(val))
        params.append(p)
    pb = rl.PackedBatch([track_of(base)], params, sw["jobs"])
    ctx.solve_batch(pb)
    for j, (_, pi, st) in enumerate(sw["jobs"]):
        r = pb.result(j)
        ref = {"xy": sw[f"j{j}_xy"], "heading": sw[f"j{j}_heading"], "curvature": sw[f"j{j}_curvature"],
               "alpha_total": sw[f"j{j}_alpha_total"], "alpha_last": sw[f"j{j}_alpha_last"], "v": sw[f"j{j}_v"],
               "ax": sw[f"j{j}_ax"]}
        assert_result_close(r, ref, "", st == MT, tag=("sweep", j))
        assert r.stats.accepted == sw[f"j{j}_accepted"] and r.stats.backtracks == sw[f"j{j}_backtracks"], j
        if st == MT:
            assert abs(r.lap_time - sw[f"j{j}_lap_time"]) <= TOL_LAP_REL * sw[f"j{j}_lap_time"]


@pytest.mark.parametrize("n", [16, 37, 64, 100, 127, 128, 130, 200, 255, 256, 257, 300, 512, 700, 1024, 1500, 2048])
def test_synthetic_tracks_vs_oracle(ctx, n):
    """ragged N across every size class (and the exact-fit kernels at N = T*K) against the pinned oracle."""
    center, seg, L, m = rl.synth_tracks(2, n, seed_base=0xB200 + 7 * n)
    center, seg = center.reshape(2, n, 2), seg.reshape(2, 2, m, 4)
    tracks = [rl.Track(center[i], seg[i, 0], seg[i, 1], L[i]) for i in range(2)]
    cfg = rl.Config()
    res = rl.solve_batch(tracks, [cfg], [(0, 0, MC), (0, 0, MT), (1, 0, MT)], ctx=ctx)
    for (t, st), r in zip([(0, MC), (0, MT), (1, MT)], res):
        o = oracle_ref(st, tracks[t], cfg.to_params())
        assert_result_close(r, o, "o_", st == MT, tag=(n, t, st))
        assert r.stats.accepted == o["stats"].accepted, (n, t, st)
        if not stalled(o["stats"]):
            assert r.stats.backtracks == o["stats"].backtracks, (n, t, st)
        if st == MT:
            assert abs(r.lap_time - o["lap"]) <= TOL_LAP_REL * o["lap"]


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 9])
def test_tiny_tracks_vs_oracle(ctx, n):
    th = 2 * np.pi * np.arange(n) / max(n, 1)
    R = 12.0
    center = np.stack([R * np.cos(th), R * np.sin(th)], axis=1)
    ring_in = rl.ring_edges(np.stack([(R - 1.7) * np.cos(np.linspace(0, 2 * np.pi, 24, endpoint=False)),
                                      (R - 1.7) * np.sin(np.linspace(0, 2 * np.pi, 24, endpoint=False))], axis=1))
    ring_out = rl.ring_edges(np.stack([(R + 1.8) * np.cos(np.linspace(0, 2 * np.pi, 30, endpoint=False)),
                                       (R + 1.8) * np.sin(np.linspace(0, 2 * np.pi, 30, endpoint=False))], axis=1))
    tr = rl.Track(center, ring_in, ring_out, 2 * np.pi * R)
    cfg = rl.Config()
    for st in (MC, MT):
        r = rl.solve_batch([tr], [cfg], [(0, 0, st)], ctx=ctx)[0]
        o = oracle_ref(st, tr, cfg.to_params())
        assert_result_close(r, o, "o_", st == MT, tag=("tiny", n, st))
        assert r.stats.accepted == o["stats"].accepted and r.stats.backtracks == o["stats"].backtracks, (n, st)


def test_edge_cases(ctx, goldens):
    g = goldens["competition_map1"]
    tr = track_of(g)
    # a ring the rays can miss entirely / an empty ring / zero iterations
    no_outer = rl.Track(tr.center_xy, tr.inner_seg, np.zeros((0, 4)), tr.L)
    far = rl.Track(tr.center_xy, tr.inner_seg, tr.outer_seg[:3], tr.L)
    cfg0 = rl.Config(max_outer_iters=0)
    cfg1 = rl.Config(max_inner_iters=0, max_outer_iters=3)
    cfg2 = rl.Config(max_vpass_iters=0)
    cfg3 = rl.Config(step_init=40.0)          # forces many Armijo backtracks
    cfg4 = rl.Config(step_init=40.0, step_min=10.0, max_outer_iters=2)   # step falls below step_min -> inner loop stops
    tracks = [tr, no_outer, far]
    cfgs = [rl.Config(), cfg0, cfg1, cfg2, cfg3, cfg4]
    jobs = [(1, 0, MC), (1, 0, MT), (2, 0, MC), (2, 0, MT), (0, 1, MT), (0, 2, MC), (0, 2, MT), (0, 3, MT), (0, 4, MC),
            (0, 4, MT), (0, 5, MC)]
    res = rl.solve_batch(tracks, cfgs, jobs, ctx=ctx)
    for (t, c, st), r in zip(jobs, res):
        o = oracle_ref(st, tracks[t], cfgs[c].to_params())
        assert_result_close(r, o, "o_", st == MT, tag=("edge", t, c, st))
        assert r.stats.accepted == o["stats"].accepted and r.stats.backtracks == o["stats"].backtracks, (t, c, st)
        assert r.stats.evals == o["stats"].evals, (t, c, st)
        if st == MT:
            assert abs(r.lap_time - o["lap"]) <= TOL_LAP_REL * max(o["lap"], 1e-9)


def test_errors_and_unsupported(ctx, goldens):
    from practice_path_planning_for_formula_student_driverless_b200._lib import lib
    tr = track_of(goldens["training_map"])
    with pytest.raises(rl.RacelineError) as e:
        rl.solve_batch([tr], [rl.Config()], [(3, 0, MC)], ctx=ctx)
    assert e.value.status == rl.RL_ERR_ARG
    with pytest.raises(rl.RacelineError) as e:
        rl.solve_batch([tr], [rl.Config()], [(0, 0, 7)], ctx=ctx)
    assert e.value.status == rl.RL_ERR_ARG
    big = rl.Track(np.zeros((40000, 2)), tr.inner_seg, tr.outer_seg, 72000.0)    # N > 32,768: beyond one 16-CTA cluster
    with pytest.raises(rl.RacelineError) as e:
        rl.solve_batch([big], [rl.Config()], [(0, 0, MC)], ctx=ctx)
    assert e.value.status == rl.RL_ERR_UNSUPPORTED
    big_open = rl.Track(np.zeros((40000, 2)), tr.inner_seg, tr.outer_seg, 72000.0, closed=False)   # open tracks have the same cap
    with pytest.raises(rl.RacelineError) as e:
        rl.solve_batch([big_open], [rl.Config()], [(0, 0, MC)], ctx=ctx)
    assert e.value.status == rl.RL_ERR_UNSUPPORTED
    assert lib().rl_solve_batch(ctx._h, None, None) == rl.RL_ERR_ARG
    # the context is still usable after errors
    r = rl.solve_batch([tr], [rl.Config()], [(0, 0, MC)], ctx=ctx)[0]
    assert r.stats.accepted == 1680


def test_full_size_properties(ctx):
    """BASELINE config 4 shape (N = 2048, M = 931/ring): size-independent properties on a few hundred tracks."""
    nt, n = 296, 2048
    center, seg, L, m = rl.synth_tracks(nt, n)
    samp_off = np.arange(nt + 1, dtype=np.int64) * n
    seg_off = np.arange(2 * nt + 1, dtype=np.int64) * m
    cfg = rl.Config()
    jobs = [(t, 0, st) for t in range(nt) for st in (MC, MT)]
    pb = rl.PackedBatch.from_arrays(samp_off, seg_off, center, seg, L, np.ones(nt, np.int32), [cfg.to_params()], jobs)
    dev = rl.DeviceBatch(ctx, pb)
    dev.solve(); dev.download(); dev.sync()
    first = [pb.result(j) for j in range(pb.n_jobs)]
    a_tot = pb.out_alpha_total.copy(); xy = pb.out_xy.copy(); v = pb.out_v.copy()
    for j, r in enumerate(first):
        st = r.stats
        assert st.status == 0 and st.outer_done == 14 and st.accepted <= 14 * 120
        for o in range(14):
            assert st.Jend[o] <= st.J0[o] * (1 + 1e-12)                  # Armijo: the cost never increases
        assert np.all(np.isfinite(r.raceline)) and np.all(np.isfinite(r.curvature))
        # the raceline stays on the track: |offset from the centre line| < ring offset (+ chord sagitta slack)
        c0 = center.reshape(nt, n, 2)[jobs[j][0]]
        assert np.max(np.linalg.norm(r.raceline - c0, axis=1)) < 1.75 + 0.25
        if jobs[j][2] == MT:
            assert np.all(r.v > 0) and np.all(r.v <= cfg.v_cap_mps + 1e-12)
            assert abs(r.lap_time - np.sum((L[jobs[j][0]] / n) / r.v)) <= 1e-9 * r.lap_time
            vk = np.sqrt(cfg.a_lat_max / np.maximum(np.abs(r.curvature), cfg.kappa_eps))
            assert np.all(r.v <= vk * (1 + 1e-12))
    # determinism: a second solve of the resident batch reproduces every bit
    dev.solve(); dev.download(); dev.sync()
    assert np.array_equal(a_tot, pb.out_alpha_total) and np.array_equal(xy, pb.out_xy) and np.array_equal(v, pb.out_v)
    # spot-check three tracks against the oracle at full size
    for t in (0, 137, 295):
        tr = rl.Track(center.reshape(nt, n, 2)[t], seg.reshape(nt, 2, m, 4)[t, 0], seg.reshape(nt, 2, m, 4)[t, 1], L[t])
        for k, stg in enumerate((MC, MT)):
            o = oracle_ref(stg, tr, cfg.to_params())
            r = first[2 * t + k]
            assert_result_close(r, o, "o_", stg == MT, tag=("full", t, stg))
            assert r.stats.accepted == o["stats"].accepted and r.stats.backtracks == o["stats"].backtracks
    dev.close()


def test_rigid_motion_and_mirror_invariance(ctx, goldens):
    """alpha is a geometric quantity: rotating/translating the inputs leaves it unchanged, mirroring negates it."""
    g = goldens["competition_map3"]
    th, sh = 0.7, np.array([13.0, -4.0])
    Rm = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])

    def xf(P, M, s):
        return P.reshape(-1, 2) @ M.T + s

    def seg_xf(S, M, s):
        return np.concatenate([xf(S[:, :2], M, s), xf(S[:, 2:], M, s)], axis=1)

    mir = np.array([[1.0, 0.0], [0.0, -1.0]])
    t0 = track_of(g)
    t1 = rl.Track(xf(g["center_xy"], Rm, sh), seg_xf(g["inner_seg"], Rm, sh), seg_xf(g["outer_seg"], Rm, sh), g["L"])
    t2 = rl.Track(xf(g["center_xy"], mir, 0), seg_xf(g["inner_seg"], mir, 0), seg_xf(g["outer_seg"], mir, 0), g["L"])
    res = rl.solve_batch([t0, t1, t2], [rl.Config()], [(0, 0, MT), (1, 0, MT), (2, 0, MT)], ctx=ctx)
    assert np.max(np.abs(res[0].alpha_total - res[1].alpha_total)) < TOL_ALPHA
    assert np.max(np.abs(res[0].alpha_total + res[2].alpha_total)) < TOL_ALPHA
    assert abs(res[0].lap_time - res[1].lap_time) < TOL_LAP_REL * res[0].lap_time
    assert abs(res[0].lap_time - res[2].lap_time) < TOL_LAP_REL * res[0].lap_time


@pytest.mark.parametrize("name,shuffled", [("competition_map3", False), ("competition_map1", True), ("competition_map2", True),
                                           ("competition_map3", True), ("competition_map_testday1", True),
                                           ("competition_map_testday2", True)])
def test_dropin_binary_matches_reference_binary(tmp_path, goldens, name, shuffled):
    """The reference's own main() with ONLY its two solver calls redirected to the CUDA library
    (oracle/_ref/fsd_path_b200, see oracle/Makefile `dropin`) against the unmodified reference binary:
    same cone files in, same CSV columns out (s,x,y,heading_rad,curvature,alpha_last,v_mps,ax_mps2; main.cpp:1364, 1420).
    `shuffled`: the cones in random order, like the reference's own csv/*_shuffled.csv sets (its front end --
    Delaunay, MST ordering -- has to put them back in order before the solver stages run)."""
    import os
    import subprocess
    ref, ours = oracle.ref_binary("fsd_path"), oracle.ref_binary("fsd_path_b200")
    if not (ref and ours):
        pytest.skip("oracle/_ref binaries not built (they are built where /root/reference is mounted)")
    g = goldens[name]
    rng = np.random.default_rng(len(name) * 131 + 7)
    inner, outer = g["inner_seg"][:, :2], g["outer_seg"][:, :2]
    if shuffled:
        inner, outer = inner[rng.permutation(len(inner))], outer[rng.permutation(len(outer))]
    np.savetxt(tmp_path / "inner.csv", inner, delimiter=",", fmt="%.17g")
    np.savetxt(tmp_path / "outer.csv", outer, delimiter=",", fmt="%.17g")
    outs = {}
    for tag, exe in (("ref", ref), ("b200", ours)):
        d = tmp_path / tag
        d.mkdir()
        r = subprocess.run([exe, str(tmp_path / "inner.csv"), str(tmp_path / "outer.csv"), str(d / "centerline.csv")],
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        lap = [l for l in r.stderr.splitlines() if "[mintime] Estimated laptime" in l]
        outs[tag] = (d, lap)
    assert outs["ref"][1] == outs["b200"][1] and outs["ref"][1], outs
    for name, tols in (("centerline_raceline_with_geom.csv", [1e-6, 1e-4, 1e-4, 1e-6, 1e-6, 1e-4, 1e-4]),
                       ("centerline_mintime_with_geom.csv", [1e-6, 1e-4, 1e-4, 1e-6, 1e-6, 1e-4, 1e-4, 1e-3])):
        a = np.loadtxt(outs["ref"][0] / name, delimiter=",", skiprows=1)
        b = np.loadtxt(outs["b200"][0] / name, delimiter=",", skiprows=1)
        assert a.shape == b.shape and a.shape[0] > 100
        for c, tol in enumerate(tols):
            assert np.max(np.abs(a[:, c] - b[:, c])) <= tol + 2e-9, (name, c, np.max(np.abs(a[:, c] - b[:, c])))
    # the stages the drop-in does not touch are byte-identical
    for name in ("centerline.csv", "centerline_with_geom.csv", "centerline_inner_from_mids.csv"):
        assert open(outs["ref"][0] / name).read() == open(outs["b200"][0] / name).read()


def test_pipelined_host_path_matches_resident_path(ctx):
    """rl_solve_batch cuts large batches into chunks (H2D | kernels | D2H overlap); same bits as one resident solve."""
    nt, n = 1300, 64
    center, seg, L, m = rl.synth_tracks(nt, n, seed_base=77)
    samp_off = np.arange(nt + 1, dtype=np.int64) * n
    seg_off = np.arange(2 * nt + 1, dtype=np.int64) * m
    jobs = [(t, 0, st) for t in range(nt) for st in (MC, MT)]                 # 2600 jobs -> 2 chunks
    cfg = rl.Config()
    pb = rl.PackedBatch.from_arrays(samp_off, seg_off, center, seg, L, np.ones(nt, np.int32), [cfg.to_params()], jobs)
    ctx.solve_batch(pb)
    got = {k: getattr(pb, "out_" + k).copy() for k in ("xy", "heading", "curvature", "alpha_total", "alpha_last", "v", "ax")}
    laps = np.array([pb.out_stats[j].lap_time for j in range(pb.n_jobs)])
    acc = np.array([pb.out_stats[j].accepted for j in range(pb.n_jobs)])
    dev = rl.DeviceBatch(ctx, pb)
    dev.solve(); dev.download(); dev.sync()
    for k, v in got.items():
        ref = getattr(pb, "out_" + k)
        mask = np.ones(len(ref), bool)
        if k in ("v", "ax"):                       # rows of min-curv jobs are left untouched
            mask = np.repeat(np.array([j[2] == MT for j in jobs]), n)
        assert np.array_equal(v[mask], ref[mask]), k
    assert np.array_equal(laps, [pb.out_stats[j].lap_time for j in range(pb.n_jobs)])
    assert np.array_equal(acc, [pb.out_stats[j].accepted for j in range(pb.n_jobs)])
    dev.close()
    tr = rl.Track(center.reshape(nt, n, 2)[1299], seg.reshape(nt, 2, m, 4)[1299, 0], seg.reshape(nt, 2, m, 4)[1299, 1], L[1299])
    o = oracle_ref(MT, tr, cfg.to_params())
    assert_result_close(pb.result(2 * 1299 + 1), o, "o_", True, tag="last job of the last chunk")


@pytest.mark.parametrize("name", ["open_competition_map1", "open_training_map"])
def test_open_track_parity_vs_reference(ctx, name):
    """closed=false (cfg is_closed_track, main.cpp:54): one-sided stencils at the ends, polyline edges, no wrap."""
    g = load_golden(name)
    tr = rl.Track(g["center_xy"], g["inner_seg"], g["outer_seg"], g["L"], closed=False)
    res = rl.solve_batch([tr], [rl.Config()], [(0, 0, MC), (0, 0, MT)], ctx=ctx)
    for st, pre, r in ((MC, "mc_", res[0]), (MT, "mt_", res[1])):
        assert_result_close(r, g, pre, st == MT, tag=(name, pre))
        assert r.stats.accepted == g[pre + "accepted"] and r.stats.backtracks == g[pre + "backtracks"], (name, pre)
    assert abs(res[1].lap_time - g["mt_lap_time"]) <= TOL_LAP_REL * g["mt_lap_time"]


@pytest.mark.parametrize("n", [1, 2, 3, 4, 7, 33, 100, 128, 200, 256, 257, 600, 2048])
def test_open_tracks_vs_oracle(ctx, n):
    """open paths of every size class: an arc of a synthetic track, rings opened at the same place."""
    nn = max(n, 16)
    center, seg, L, m = rl.synth_tracks(1, nn + 8, seed_base=0xB200 + 3 * n)
    center, seg = center.reshape(nn + 8, 2), seg.reshape(2, m, 4)
    keep = max(3, int(m * n / (nn + 8)))
    tr = rl.Track(center[:n], rl.polyline_edges(seg[0, :keep + 2, :2]), rl.polyline_edges(seg[1, :keep + 2, :2]),
                  L[0] * n / (nn + 8), closed=False)
    cfg = rl.Config()
    res = rl.solve_batch([tr], [cfg], [(0, 0, MC), (0, 0, MT)], ctx=ctx)
    for st, r in zip((MC, MT), res):
        o = oracle_ref(st, tr, cfg.to_params())
        assert_result_close(r, o, "o_", st == MT, tag=("open", n, st))
        assert r.stats.accepted == o["stats"].accepted, (n, st)
        if not stalled(o["stats"]):
            assert r.stats.backtracks == o["stats"].backtracks, (n, st)
        if st == MT:
            assert abs(r.lap_time - o["lap"]) <= TOL_LAP_REL * max(o["lap"], 1e-9)


def test_config_sweep_4096(ctx, goldens):
    """BASELINE configs[2]: competition_map2 x 8^4 grid over lambda_smooth x mu x P_max_W x w_time_gain (SURVEY 8d ranges;
    a_total_max = mu * 9.81 set per combo, the reference evaluates it once at construction, main.cpp:102).
    All 8192 jobs in one batch sharing one copy of the geometry; 64 combos spot-checked against the pinned oracle."""
    import time
    g = goldens["competition_map2"]
    tr = track_of(g)
    lam = np.logspace(np.log10(4e-4), np.log10(6.4e-3), 8)
    mu = np.linspace(1.15, 1.6, 8)
    pmax = np.linspace(20e3, 80e3, 8)
    wtg = np.linspace(0.0, 3.5, 8)
    cfgs = [rl.Config(lambda_smooth=float(a), a_total_max=float(b) * 9.81, P_max_W=float(c), w_time_gain=float(d))
            for a in lam for b in mu for c in pmax for d in wtg]
    assert len(cfgs) == 4096
    jobs = [(0, p, st) for p in range(4096) for st in (MC, MT)]
    pb = rl.PackedBatch([tr], [c.to_params() for c in cfgs], jobs)
    t0 = time.perf_counter()
    ctx.solve_batch(pb)
    dt = time.perf_counter() - t0
    print(f"4096-combo sweep: {4096 / dt:.0f} solves/s end to end")
    laps = np.array([pb.out_stats[2 * p + 1].lap_time for p in range(4096)])
    assert np.all(np.isfinite(laps)) and np.all(laps > 10.0) and np.all(laps < 80.0)
    assert all(pb.out_stats[j].status == 0 and pb.out_stats[j].outer_done == 14 for j in range(0, 8192, 37))
    # the grid really is 4096 different problems: (almost) no two combos predict the same lap
    assert np.unique(np.round(laps, 9)).size > 2000
    rng = np.random.default_rng(4096)
    for p in rng.choice(4096, size=64, replace=False):
        prm = cfgs[p].to_params()
        for k, st in enumerate((MC, MT)):
            o = oracle_ref(st, tr, prm)
            r = pb.result(2 * int(p) + k)
            assert_result_close(r, o, "o_", st == MT, tag=("sweep4096", int(p), st))
            assert r.stats.accepted == o["stats"].accepted, (int(p), st)
            if not stalled(o["stats"]):
                assert r.stats.backtracks == o["stats"].backtracks, (int(p), st)
            if st == MT:
                assert abs(r.lap_time - o["lap"]) <= TOL_LAP_REL * o["lap"]
