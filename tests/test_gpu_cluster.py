"""GPU parity tests of the long-track path: one thread-block cluster per job (raceline_cluster.cuh).

The cluster kernel serves N in (4096, 16384] (BASELINE configs[4] is N = 16,384).  the "force_cluster" option (rl_set_option) routes
shorter closed tracks through it as well, so that every cluster size and the ragged / exact-fit chunkings can be
checked against the pinned oracle at sizes the CPU finishes in seconds.  Same tolerances as test_gpu_parity.py.
"""
import os

import numpy as np
import pytest

import practice_path_planning_for_formula_student_driverless_b200 as rl
from conftest import TOL_ALPHA, TOL_LAP_REL, assert_result_close
from test_gpu_parity import MC, MT, oracle_ref, stalled

pytestmark = pytest.mark.gpu


def _tracks(nt, n, seed):
    center, seg, L, m = rl.synth_tracks(nt, n, seed_base=seed)
    center, seg = center.reshape(nt, n, 2), seg.reshape(nt, 2, m, 4)
    return [rl.Track(center[i], seg[i, 0], seg[i, 1], L[i]) for i in range(nt)]


def _check(r, tr, st, cfg, tag):
    o = oracle_ref(st, tr, cfg.to_params())
    assert_result_close(r, o, "o_", st == MT, tag=tag)
    assert r.stats.status == 0 and r.stats.n == tr.center_xy.shape[0]
    assert r.stats.accepted == o["stats"].accepted, tag
    if not stalled(o["stats"]):
        assert r.stats.backtracks == o["stats"].backtracks, tag
        assert [int(x) for x in r.stats.bt_outer[:14]] == [int(x) for x in o["stats"].bt_outer[:14]], tag
    if st == MT:
        assert abs(r.lap_time - o["lap"]) <= TOL_LAP_REL * o["lap"], tag


@pytest.mark.parametrize("cs,n", [(2, 1024), (2, 1100), (2, 4096), (4, 2500), (4, 3000), (8, 4100)])
def test_forced_cluster_vs_oracle(ctx, cs, n):
    """every cluster size, ragged and exact-fit chunks, on tracks the single-CTA kernels could also solve"""
    ctx.set_option("force_cluster", cs)
    tracks = _tracks(2, n, 0xC100 + n)
    cfg = rl.Config()
    jobs = [(0, 0, MC), (0, 0, MT), (1, 0, MT)]
    res = rl.solve_batch(tracks, [cfg], jobs, ctx=ctx)
    ctx.set_option("force_cluster", 0)
    for (t, _, st), r in zip(jobs, res):
        _check(r, tracks[t], st, cfg, ("cluster", cs, n, t, st))
    # and the same bits as the single-CTA kernel on the same problem where that kernel exists
    if n <= 4096:
        res1 = rl.solve_batch(tracks, [cfg], jobs, ctx=ctx)
        for a, b in zip(res, res1):
            assert a.stats.accepted == b.stats.accepted and a.stats.backtracks == b.stats.backtracks
            assert np.max(np.abs(a.alpha_total - b.alpha_total)) < 1e-9
            if a.v is not None and b.v is not None and a.lap_time:
                assert abs(a.lap_time - b.lap_time) <= 1e-9 * b.lap_time


def test_forced_cluster_config_variants(ctx):
    """inverse-speed weights (a third cluster-wide reduction), heavy backtracking and early stops on the cluster path"""
    ctx.set_option("force_cluster", 2)
    tracks = _tracks(1, 1300, 0xC177)
    cfgs = [rl.Config(time_weight_use_inv_v=True, inv_v_gain=0.4), rl.Config(step_init=40.0),
            rl.Config(step_init=40.0, step_min=10.0, max_outer_iters=2), rl.Config(max_vpass_iters=1, P_max_W=20000.0)]
    jobs = [(0, 0, MT), (0, 1, MC), (0, 1, MT), (0, 2, MC), (0, 3, MT)]
    res = rl.solve_batch(tracks, cfgs, jobs, ctx=ctx)
    for (t, c, st), r in zip(jobs, res):
        _check(r, tracks[t], st, cfgs[c], ("cluster-cfg", c, st))


def test_long_track_natural_dispatch(ctx):
    """N = 5000 (4 CTAs, ragged chunks) takes the cluster kernel without any hook; batched with short tracks"""
    long_tr = _tracks(1, 5000, 0xC500)[0]
    short = _tracks(1, 300, 0xC501)[0]
    cfg = rl.Config()
    jobs = [(0, 0, MC), (1, 0, MT), (0, 0, MT)]
    res = rl.solve_batch([long_tr, short], [cfg], jobs, ctx=ctx)
    for (t, _, st), r in zip(jobs, res):
        _check(r, [long_tr, short][t], st, cfg, ("long", t, st))


def test_baseline_config5_shape(ctx):
    """BASELINE configs[4] shape: N = 16,384 samples, M = 7447 cones per ring, 8 CTAs per track.  Size-independent
    properties on a few tracks, bit-reproducibility, and (unless RL_SKIP_SLOW) one min-time solve against the oracle."""
    nt, n = 3, 16384
    center, seg, L, m = rl.synth_tracks(nt, n, seed_base=0xB200)
    assert m == 7447
    samp_off = np.arange(nt + 1, dtype=np.int64) * n
    seg_off = np.arange(2 * nt + 1, dtype=np.int64) * m
    cfg = rl.Config()
    jobs = [(t, 0, st) for t in range(nt) for st in (MC, MT)]
    pb = rl.PackedBatch.from_arrays(samp_off, seg_off, center, seg, L, np.ones(nt, np.int32), [cfg.to_params()], jobs)
    dev = rl.DeviceBatch(ctx, pb)
    dev.solve(); dev.download(); dev.sync()
    first = [pb.result(j) for j in range(pb.n_jobs)]
    a_tot = pb.out_alpha_total.copy(); xy = pb.out_xy.copy()
    for j, r in enumerate(first):
        st = r.stats
        assert st.status == 0 and st.n == n and st.outer_done == 14 and 0 < st.accepted <= 14 * 120
        for o in range(14):
            assert st.Jend[o] <= st.J0[o] * (1 + 1e-12)
        c0 = center.reshape(nt, n, 2)[jobs[j][0]]
        assert np.all(np.isfinite(r.raceline)) and np.max(np.linalg.norm(r.raceline - c0, axis=1)) < 2.0
        if jobs[j][2] == MT:
            assert np.all(r.v > 0) and np.all(r.v <= cfg.v_cap_mps + 1e-12)
            assert abs(r.lap_time - np.sum((L[jobs[j][0]] / n) / r.v)) <= 1e-9 * r.lap_time
    dev.solve(); dev.download(); dev.sync()
    assert np.array_equal(a_tot, pb.out_alpha_total) and np.array_equal(xy, pb.out_xy)
    dev.close()
    if os.environ.get("RL_SKIP_SLOW"):
        return
    tr = rl.Track(center.reshape(nt, n, 2)[1], seg.reshape(nt, 2, m, 4)[1, 0], seg.reshape(nt, 2, m, 4)[1, 1], L[1])
    _check(first[3], tr, MT, cfg, ("config5", 1, MT))


def test_track_longer_than_16384_on_a_16_cta_cluster(ctx):
    """The reference solver has no length cap (N = center.size(), main.cpp:689).  Closed tracks of 16,385 .. 32,768
    samples run on a 16-CTA thread-block cluster (a non-portable cluster size the launch has to opt in to); the test
    skips where the device does not schedule such clusters.  Few cones per ring so that the brute-force oracle
    (O(N*M) per corridor build) stays within seconds."""
    n, m_ring = 20000, 400
    center, seg, L, m = rl.synth_tracks(1, n, m_ring, seed_base=0x16C7)
    tr = rl.Track(center.reshape(n, 2), seg.reshape(2, m, 4)[0], seg.reshape(2, m, 4)[1], float(L[0]))
    cfg = rl.Config()
    try:
        res = rl.solve_batch([tr], [cfg], [(0, 0, MC), (0, 0, MT)], ctx=ctx)
    except rl.RacelineError as e:
        if e.status == rl.RL_ERR_UNSUPPORTED:
            pytest.skip("this device does not schedule 16-CTA clusters of the kernel")
        raise
    for st, r in zip((MC, MT), res):
        assert r.stats.status == 0 and r.stats.n == n
        _check(r, tr, st, cfg, ("16-CTA cluster", n, st))


@pytest.mark.parametrize("cs,n", [(4, 6000), (8, 16384)])
def test_few_sample_search_equals_tile_streaming_bit_for_bit(ctx, cs, n):
    """The samples the update path flags are rebuilt either by the box-culled whole-CTA search (corridor_search_few_c) or by
    streaming every ring tile again (corridor_search_c).  Both are exact: hits, point distances and therefore the corridor
    are the same numbers, so a chained min-curv + min-time solve must agree bit for bit -- and the few-sample path must
    actually have run (its cone certificates count as existence scans, like the tile path's)."""
    ctx.set_option("force_cluster", cs if n <= 4096 * 2 else 0)
    tracks = _tracks(1, n, 0xFE00 + n)
    cfg = rl.Config()
    jobs = [(0, 0, MC), (0, 0, MT)]
    try:
        a = rl.solve_batch(tracks, [cfg], jobs, ctx=ctx)
        ctx.set_option("no_few_search", 1)
        b = rl.solve_batch(tracks, [cfg], jobs, ctx=ctx)
    finally:
        ctx.set_option("no_few_search", 0)
        ctx.set_option("force_cluster", 0)
    for ra, rb in zip(a, b):
        assert ra.stats.status == 0 and rb.stats.status == 0
        assert ra.stats.accepted == rb.stats.accepted and ra.stats.backtracks == rb.stats.backtracks and ra.stats.evals == rb.stats.evals
        for name in ("raceline", "alpha_total", "alpha_last", "heading", "curvature"):
            assert np.array_equal(getattr(ra, name), getattr(rb, name)), (name, cs, n)
        if ra.v is not None:
            assert np.array_equal(ra.v, rb.v) and np.array_equal(ra.ax, rb.ax) and ra.lap_time == rb.lap_time
        assert ra.stats.exist_scans > 0 and rb.stats.exist_scans > 0


def _open_arc(n, seed):
    """an arc of a synthetic track as an OPEN path: the first n samples, rings cut at the same place (polylines)"""
    center, seg, L, m = rl.synth_tracks(1, n + 8, seed_base=seed)
    center, seg = center.reshape(n + 8, 2), seg.reshape(2, m, 4)
    keep = max(3, int(m * n / (n + 8)))
    return rl.Track(center[:n], rl.polyline_edges(seg[0, :keep + 2, :2]), rl.polyline_edges(seg[1, :keep + 2, :2]),
                    L[0] * n / (n + 8), closed=False)


@pytest.mark.parametrize("cs,n", [(2, 1100), (2, 4096), (4, 3000), (8, 4100), (0, 5000)])
def test_open_tracks_on_the_cluster_path(ctx, cs, n):
    """DiffOpsOpen (main.cpp:560-579) on a thread-block cluster: one-sided ends in the first / last chunk only, no wrap in
    the v(s) sweeps, ax = 0 at the last sample.  Forced clusters on open arcs the single-CTA open kernel also solves
    (oracle + the single-CTA result), and one open track longer than 4096 samples through the natural dispatch."""
    tr = _open_arc(n, 0x0E00 + n)
    cfg = rl.Config()
    jobs = [(0, 0, MC), (0, 0, MT)]
    ctx.set_option("force_cluster", cs)
    try:
        res = rl.solve_batch([tr], [cfg], jobs, ctx=ctx)
    finally:
        ctx.set_option("force_cluster", 0)
    for (t, _, st), r in zip(jobs, res):
        _check(r, tr, st, cfg, ("open cluster", cs, n, st))
    if n <= 4096:
        res1 = rl.solve_batch([tr], [cfg], jobs, ctx=ctx)
        for a, b in zip(res, res1):
            assert a.stats.accepted == b.stats.accepted and a.stats.backtracks == b.stats.backtracks
            assert np.max(np.abs(a.alpha_total - b.alpha_total)) < 1e-9
            if a.v is not None and a.lap_time:
                assert abs(a.lap_time - b.lap_time) <= 1e-9 * b.lap_time
