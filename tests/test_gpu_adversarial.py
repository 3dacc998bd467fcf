"""Adversarial-geometry parity suite (tests/trackgen.py): jittered cones, varying width, hairpins, sections a few metres
apart, rings as segment soup -- everything the smooth constant-width bench tracks do not have -- against the oracle
(oracle/raceline_oracle.c, pinned bit for bit to the reference; the corridor of main.cpp:478-512, 694-711 makes no
assumption about the rings, so neither may the certificates / anchors / clearances of the kernels).

Also: the shipped maps pushed through the GPU centre-line stage at 500 / 1000 / 2048 samples and then solved.
Every case runs both stages as one chain (the second job starts from the first job's corridor state).
"""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import practice_path_planning_for_formula_student_driverless_b200 as rl
import trackgen
from conftest import TOL_LAP_REL, assert_result_close, load_golden
from oracle import oracle
from test_gpu_parity import MC, MT, stalled

pytestmark = pytest.mark.gpu

CASES = [   # (kind, n, generator keywords, Config keywords)
    ("flower", 2048, dict(jitter=0.12, width=(1.25, 3.0)), {}),
    ("flower", 1000, dict(jitter=0.20), dict(safety_margin_m=0.2)),
    ("paperclip", 300, dict(jitter=0.05), {}),
    ("paperclip", 700, dict(jitter=0.10, width=(1.3, 2.0), soup=(True, False)), {}),
    ("serpentine", 1500, dict(jitter=0.08), {}),
    ("serpentine", 2048, dict(width=(1.5, 2.2)), dict(veh_width_m=1.3)),
    ("hourglass", 900, dict(jitter=0.05), {}),
    ("hourglass", 2048, dict(width=(1.4, 1.9), soup=(False, True)), {}),
    ("flower", 3000, dict(jitter=0.10, width=(1.25, 2.5)), {}),
    ("serpentine", 400, dict(jitter=0.15, reverse_outer=True), dict(w_time_gain=2.0)),
    ("paperclip", 1200, dict(jitter=0.2, width=(1.25, 3.0), soup=(True, True)), {}),
    ("hourglass", 2048, dict(jitter=0.18, width=(1.25, 2.2)), dict(lambda_smooth=4e-4)),
]


def _oracle_all(work):
    """work: [(stage, track, params)] -> [(reference-rounded result, FMA-rounded result)]; the C restatement releases
    the GIL inside ctypes: one thread per host core."""
    def one(w):
        st, tr, p = w
        return (oracle.solve(st, tr.center_xy, tr.inner_seg, tr.outer_seg, tr.L, tr.closed, p),
                oracle.solve(st, tr.center_xy, tr.inner_seg, tr.outer_seg, tr.L, tr.closed, p, rounding="fma"))
    with ThreadPoolExecutor(max_workers=16) as ex:
        return list(ex.map(one, work))


def _dev(a, b):
    m = lambda k: float(np.max(np.abs(np.asarray(a[k]) - np.asarray(b[k])))) if len(a[k]) else 0.0
    return {"alpha": max(m("alpha_total"), m("alpha_last")), "xy": m("xy"), "kappa": m("curvature"), "v": m("v"),
            "lap": abs(a["lap_time"] - b["lap_time"]) / max(1e-30, abs(b["lap_time"])) if b["lap_time"] else 0.0}


def _compare(r, pair, st, tag, tally):
    """North-star tolerances -- unless the case is ROUNDING-SENSITIVE: on some of these tracks an outer iteration that
    leaves through the |dJ| < 1e-10 exit (main.cpp:740) before its first backtrack lets the reference's PGD amplify
    last-bit differences by ~1.7x per step (1e-17 -> 1e-6 m in 48 steps), so the reference's own algorithm, merely
    re-rounded (FMA-contracted build of the oracle), lands as far from the reference as the GPU does.  Such a job is
    held to 4x that yardstick instead, and counted."""
    o, f = pair
    g = {"alpha_total": r.alpha_total, "alpha_last": r.alpha_last, "xy": r.raceline, "curvature": r.curvature,
         "v": r.v if r.v is not None else o["v"], "lap_time": r.lap_time if st == MT else o["lap_time"]}
    e, y = _dev(g, o), _dev(f, o)
    tol = {"alpha": 1e-4, "xy": 1e-4, "kappa": 1e-6, "v": 1e-4, "lap": TOL_LAP_REL}
    sensitive = any(y[k] > 0.05 * tol[k] for k in tol) or f["stats"].accepted != o["stats"].accepted
    tally["jobs"] += 1
    if not sensitive:
        d = {"o_" + k: o[k] for k in ("xy", "heading", "curvature", "alpha_total", "alpha_last", "v", "ax")}
        assert_result_close(r, d, "o_", st == MT, tag=tag)
        if r.stats.accepted != o["stats"].accepted:
            # Only an outer iteration that leaves through the |dJ| < 1e-10 exit (main.cpp:740; it never fires on the
            # shipped maps) may differ, and by one step: whether the step whose dJ is about 1e-10 is the last one is
            # decided in the last bits of J.
            for k in range(o["stats"].outer_done):
                a, b = int(r.stats.acc_outer[k]), int(o["stats"].acc_outer[k])
                assert a == b or (abs(a - b) == 1 and min(a, b) < 120 and r.stats.bt_outer[k] == o["stats"].bt_outer[k]), (tag, k, a, b)
            tally["exit_knife_edge"] = tally.get("exit_knife_edge", 0) + 1
        if stalled(o["stats"]):
            tally["stalled"] += 1
        else:
            assert r.stats.backtracks == o["stats"].backtracks, (tag, r.stats.backtracks, o["stats"].backtracks)
        if st == MT:
            assert abs(r.lap_time - o["lap_time"]) <= TOL_LAP_REL * o["lap_time"], tag
        return
    tally["rounding_sensitive"] += 1
    tally["notes"].append((tag, {k: (e[k], y[k]) for k in e}))
    for k in tol:
        assert e[k] <= max(tol[k], 4.0 * y[k]), (tag, k, "gpu vs reference", e[k], "re-rounded reference vs reference", y[k])
    assert e["alpha"] <= 1e-3 and e["lap"] <= 1e-4, (tag, e)       # sensitive, not wrong


def _run(ctx, cases, seed, force_cluster=0):
    tracks, cfgs, jobs = [], [], []
    for i, (kind, n, gkw, ckw) in enumerate(cases):
        c, inner, outer, L = trackgen.make_track(seed + i, n, kind, **gkw)
        tracks.append(rl.Track(c, inner, outer, L))
        cfgs.append(rl.Config(**ckw))
        jobs += [(i, i, MC), (i, i, MT)]
    ctx.set_option("force_chain", 2)
    if force_cluster:
        ctx.set_option("force_cluster", force_cluster)
    res = rl.solve_batch(tracks, cfgs, jobs, ctx=ctx)
    ctx.set_option("force_chain", 0); ctx.set_option("force_cluster", 0)
    ref = _oracle_all([(st, tracks[t], cfgs[c].to_params()) for (t, c, st) in jobs])
    tally = {"stalled": 0, "jobs": 0, "fallback_scans": 0, "rounding_sensitive": 0, "notes": []}
    for (t, c, st), r, o in zip(jobs, res, ref):
        _compare(r, o, st, (cases[t][0], cases[t][1], "cluster" if force_cluster else "cta", st), tally)
        if st == MT:      # second job of its chain: no first build, so every full existence search here is the fallback path
            tally["fallback_scans"] += r.stats.exist_scans
    return tally


def test_adversarial_tracks_single_cta(ctx):
    tally = _run(ctx, CASES, seed=9100)
    print("adversarial (one CTA per chain):", tally)
    assert tally["fallback_scans"] > 0, "no case reached the searching fallback of the corridor update: the suite lost its teeth"
    assert tally["stalled"] <= 2 and tally["rounding_sensitive"] <= tally["jobs"] // 3, tally


@pytest.mark.parametrize("cs", [2, 4])
def test_adversarial_tracks_cluster(ctx, cs):
    lo, hi = 512 * cs, 2048 * cs
    cases = [c for c in CASES if lo <= c[1] <= hi]
    assert len(cases) >= 3
    tally = _run(ctx, cases, seed=9100, force_cluster=cs)
    print(f"adversarial ({cs}-CTA cluster per chain):", tally)
    assert tally["stalled"] <= 2 and tally["rounding_sensitive"] <= tally["jobs"] // 3, tally


def test_second_seed_without_chains(ctx):
    """the same generators, other seeds, every job on its own (first build + certificates from scratch each time)"""
    tracks, jobs = [], []
    picks = [CASES[k] for k in (2, 3, 6, 9, 10)]
    for i, (kind, n, gkw, _) in enumerate(picks):
        c, inner, outer, L = trackgen.make_track(7700 + i, n, kind, **gkw)
        tracks.append(rl.Track(c, inner, outer, L))
        jobs += [(i, 0, MT)]
    res = rl.solve_batch(tracks, [rl.Config()], jobs, ctx=ctx)
    ref = _oracle_all([(MT, tracks[t], rl.Config().to_params()) for (t, _, _) in jobs])
    tally = {"stalled": 0, "jobs": 0, "rounding_sensitive": 0, "notes": []}
    for (t, _, st), r, o in zip(jobs, res, ref):
        _compare(r, o, st, ("unchained", picks[t][0], picks[t][1]), tally)
    print("adversarial (unchained):", tally)


@pytest.mark.parametrize("samples", [500, 1000, 2048])
def test_shipped_maps_through_the_gpu_centerline_stage(ctx, samples):
    """mid points of the shipped maps -> rl_centerline_geom_batch at `samples` rows -> both solver stages, against the
    oracle fed with the SAME centre line (real cone noise at every size class up to N = 2048)."""
    names = ["geom_training_map", "geom_competition_map1", "geom_competition_map_testday2"]
    gs = [load_golden(n) for n in names]
    geo = rl.centerline_geom_batch([g["mids_xy"] for g in gs], [samples] * len(gs), [g["inner_seg"] for g in gs],
                                   [g["outer_seg"] for g in gs], closed=True, ctx=ctx)
    tracks = [rl.Track(q.center_for_opt, g["inner_seg"], g["outer_seg"], q.L) for q, g in zip(geo, gs)]
    jobs = [(t, 0, st) for t in range(len(tracks)) for st in (MC, MT)]
    ctx.set_option("force_chain", 2)
    res = rl.solve_batch(tracks, [rl.Config()], jobs, ctx=ctx)
    ctx.set_option("force_chain", 0)
    ref = _oracle_all([(st, tracks[t], rl.Config().to_params()) for (t, _, st) in jobs])
    tally = {"stalled": 0, "jobs": 0, "rounding_sensitive": 0, "notes": []}
    for (t, _, st), r, o in zip(jobs, res, ref):
        _compare(r, o, st, (names[t], samples, st), tally)
    print("shipped maps at", samples, "samples:", tally)
    assert tally["stalled"] == 0
