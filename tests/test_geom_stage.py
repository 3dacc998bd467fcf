"""The stage before the path: centre line (natural cubic spline fit + uniform resample,
main.cpp:404-474, 1270-1279) and the width/geometry rows of <base>_with_geom.csv (main.cpp:1288-1335).

Goldens (tests/golden/geom_*.npz) come from `ref_harness geom`, which drives the reference's own Spline1D /
distancesToRings and checks itself against the CSV the reference writes.  CPU tests pin the C restatement
(oracle/geom_oracle.c) to them bit for bit; GPU tests compare the CUDA kernels, through the C ABI, with both.
The CUDA unit is compiled without FMA contraction, so everything except atan2 / pow is expected to agree to the
last bit; the tolerances below are what the test enforces.
"""
import numpy as np
import pytest

import practice_path_planning_for_formula_student_driverless_b200 as rl
from conftest import load_golden
from oracle import oracle

GEOM_CASES = ["geom_training_map", "geom_competition_map1", "geom_competition_map_testday2", "geom_competition_map2_n1000",
              "geom_open_competition_map3"]
TOL_XY = 1e-9          # m
TOL_ANGLE = 1e-12      # rad
TOL_CURV = 1e-12       # 1/m (relative to max(1, |kappa|))
TOL_DIST = 1e-9        # m


@pytest.mark.parametrize("name", GEOM_CASES)
def test_geom_oracle_matches_reference_bitwise(name):
    g = load_golden(name)
    r = oracle.centerline_geom(g["mids_xy"], g["samples"], g["inner_seg"], g["outer_seg"], bool(g["closed"]),
                               bool(g["emit_closed_duplicate"]))
    assert r["xy"].shape[0] == g["rows"]
    assert r["L"] == g["L"] and r["s0"] == g["s0"]
    assert np.array_equal(r["xy"][:, 0], g["x"]) and np.array_equal(r["xy"][:, 1], g["y"])
    for k in ("s_rel", "heading", "curvature", "dist_inner", "dist_outer", "width", "v_kappa"):
        assert np.array_equal(r[k], g[k]), k
    # the centre line handed to the solver stages is the first `samples` rows
    assert np.array_equal(g["center_xy"][:g["samples"]], r["xy"][:g["samples"]])


def test_geom_oracle_properties():
    """a circle of mid points: constant curvature 1/R, width = ring gap, heading tangent to the circle"""
    R, n = 40.0, 120
    th = 2 * np.pi * np.arange(n) / n
    mids = np.stack([R * np.cos(th), R * np.sin(th)], axis=1)
    ring = lambda r, m: rl.ring_edges(np.stack([r * np.cos(2 * np.pi * np.arange(m) / m), r * np.sin(2 * np.pi * np.arange(m) / m)], axis=1))
    r = oracle.centerline_geom(mids, 200, ring(R - 1.5, 300), ring(R + 1.5, 300), True, True)
    assert abs(r["L"] - 2 * R * n * np.sin(np.pi / n) * (n - 1) / n) < 1e-9          # chord length of the first n-1 gaps
    k = r["curvature"][5:-5]
    assert np.all(np.abs(k - 1.0 / R) < 2e-4)
    assert np.all(np.abs(r["width"] - 3.0) < 0.02)


def _check_gpu(res, ref, tag):
    assert res.xy.shape == ref["xy"].shape, tag
    assert np.max(np.abs(res.xy - ref["xy"])) <= TOL_XY, tag
    assert abs(res.L - ref["L"]) <= 1e-12 * ref["L"] and abs(res.s0 - ref["s0"]) <= 1e-12 * max(1.0, ref["s0"]), tag
    assert np.max(np.abs(res.s - ref["s_rel"])) <= 1e-9, tag
    dh = np.abs(np.angle(np.exp(1j * (res.heading - ref["heading"]))))
    assert np.max(dh) <= TOL_ANGLE, (tag, np.max(dh))
    assert np.max(np.abs(res.curvature - ref["curvature"]) / np.maximum(1.0, np.abs(ref["curvature"]))) <= TOL_CURV, tag
    for k in ("dist_inner", "dist_outer", "width"):
        assert np.max(np.abs(getattr(res, k) - ref[k])) <= TOL_DIST, (tag, k)
    assert np.max(np.abs(res.v_kappa - ref["v_kappa"])) <= 1e-9, tag


@pytest.mark.gpu
def test_geom_gpu_vs_reference_goldens(ctx):
    """all golden tracks in ONE batched call (closed and open, 187 ... 1000 samples)"""
    gs = [load_golden(n) for n in GEOM_CASES]
    res = rl.centerline_geom_batch([g["mids_xy"] for g in gs], [g["samples"] for g in gs], [g["inner_seg"] for g in gs],
                                   [g["outer_seg"] for g in gs], closed=[bool(g["closed"]) for g in gs], ctx=ctx)
    for name, g, r in zip(GEOM_CASES, gs, res):
        ref = {"xy": np.stack([g["x"], g["y"]], axis=1), **{k: g[k] for k in ("s_rel", "heading", "curvature", "dist_inner",
                                                                             "dist_outer", "width", "v_kappa", "L", "s0")}}
        _check_gpu(r, ref, name)
        assert r.samples == g["samples"] and r.center_for_opt.shape[0] == g["samples"]
        # bitwise where no libm function is involved
        assert np.array_equal(r.xy, ref["xy"]) and np.array_equal(r.dist_inner, g["dist_inner"]), name


@pytest.mark.gpu
@pytest.mark.parametrize("n_mid,samples", [(3, 5), (7, 40), (64, 64), (500, 2500), (2000, 3000), (4090, 1000)])
def test_geom_gpu_vs_oracle_synthetic(ctx, n_mid, samples):
    """mid points sub-sampled from synthetic centre lines: spline sizes up to the 4090-point limit, more rows than one
    2048-row chunk, rings larger than one shared-memory tile"""
    n = max(n_mid, 16)
    center, seg, L, m = rl.synth_tracks(2, n, m_per_ring=max(8, min(3000, n)), seed_base=0x6E0 + n_mid)
    center, seg = center.reshape(2, n, 2), seg.reshape(2, 2, m, 4)
    mids = [center[i][:n_mid] for i in range(2)]
    cfg = rl.Config()
    for closed in (True, False):
        res = rl.centerline_geom_batch(mids, [samples, samples + 3], [seg[0, 0], seg[1, 0]], [seg[0, 1], seg[1, 1]], closed=closed,
                                       cfg=cfg, ctx=ctx)
        for i, r in enumerate(res):
            o = oracle.centerline_geom(mids[i], samples + 3 * i, seg[i, 0], seg[i, 1], closed, True, cfg.to_params())
            _check_gpu(r, o, (n_mid, samples, closed, i))


@pytest.mark.gpu
def test_geom_gpu_errors_and_edge_cases(ctx, goldens):
    g = load_golden("geom_training_map")
    with pytest.raises(rl.RacelineError) as e:      # fewer than 3 mid points: the reference does not fit a spline either
        rl.centerline_geom_batch([g["mids_xy"][:2]], [10], [g["inner_seg"]], [g["outer_seg"]], ctx=ctx)
    assert e.value.status == rl.RL_ERR_ARG
    with pytest.raises(rl.RacelineError) as e:
        rl.centerline_geom_batch([np.zeros((5000, 2))], [10], [g["inner_seg"]], [g["outer_seg"]], ctx=ctx)
    assert e.value.status == rl.RL_ERR_UNSUPPORTED
    # an empty ring: both distances of that ring are 0 (main.cpp:523)
    r = rl.centerline_geom_batch([g["mids_xy"]], [g["samples"]], [np.zeros((0, 4))], [g["outer_seg"]], ctx=ctx)[0]
    o = oracle.centerline_geom(g["mids_xy"], g["samples"], np.zeros((0, 4)), g["outer_seg"], True, True)
    assert np.all(r.dist_inner == 0.0)
    _check_gpu(r, o, "empty inner ring")
    # a ring no ray can hit (three far segments): the point-ring distance takes over (main.cpp:519)
    far = g["outer_seg"][:3] + 500.0
    r = rl.centerline_geom_batch([g["mids_xy"]], [g["samples"]], [g["inner_seg"]], [far], ctx=ctx)[0]
    o = oracle.centerline_geom(g["mids_xy"], g["samples"], g["inner_seg"], far, True, True)
    _check_gpu(r, o, "far outer ring")
    # the centre line feeds the solver stages: same result as solving on the golden centre line
    tr = rl.Track(r.center_for_opt, g["inner_seg"], g["outer_seg"], r.L)
    mt = rl.solve_batch([tr], [rl.Config()], [(0, 0, rl.RL_STAGE_MINTIME)], ctx=ctx)[0]
    ref = goldens["training_map"]
    assert abs(mt.lap_time - ref["mt_lap_time"]) <= 1e-5 * ref["mt_lap_time"]


@pytest.mark.gpu
def test_packed_geom_pinned_reuse_and_kernel_time(ctx):
    """PackedGeom: pack once (page-locked), run the C-ABI call repeatedly -- the shape bench.py times; same rows as the
    one-shot wrapper, and the library reports the device time of its two kernels."""
    gs = [load_golden(n) for n in GEOM_CASES[:3]]
    args = ([g["mids_xy"] for g in gs], [int(g["samples"]) for g in gs], [g["inner_seg"] for g in gs], [g["outer_seg"] for g in gs])
    once = rl.centerline_geom_batch(*args, closed=True, ctx=ctx)
    pool = rl.PinnedPool()
    pg = rl.PackedGeom(*args, closed=True, pool=pool)
    for _ in range(3):
        pg.run(ctx)
    assert ctx.last_kernel_ms() > 0.0
    for a, b in zip(once, pg.results()):
        assert np.array_equal(a.xy, b.xy) and np.array_equal(a.width, b.width) and a.L == b.L
    assert pg.d2h_bytes == pg.rows * 72 + 16 * len(gs)
    del pg, b
    pool.close(force=True)


@pytest.mark.gpu
def test_geom_batch_pipelined_ranges_equal_one_call(ctx):
    """rl_centerline_geom_batch as a pipeline of track ranges (what large batches get automatically) against the
    single upload | kernels | download call on the same batch: every output column and L, s0 bit for bit."""
    gs = [load_golden(n) for n in GEOM_CASES[:3]] * 4 + [load_golden(GEOM_CASES[0])]          # 13 tracks, ragged shapes
    args = ([g["mids_xy"] for g in gs], [int(g["samples"]) for g in gs], [g["inner_seg"] for g in gs], [g["outer_seg"] for g in gs])
    ctx.set_option("geom_chunks", 1)
    one = rl.centerline_geom_batch(*args, closed=True, ctx=ctx)
    assert ctx.last_kernel_ms() > 0.0
    for chunks in (2, 5, 13, 16):
        ctx.set_option("geom_chunks", chunks)
        many = rl.centerline_geom_batch(*args, closed=True, ctx=ctx)
        assert ctx.last_kernel_ms() < 0.0          # kernels and copies overlap: no separate kernel time
        for a, b in zip(one, many):
            for col in ("xy", "s", "heading", "curvature", "dist_inner", "dist_outer", "width", "v_kappa"):
                assert np.array_equal(getattr(a, col), getattr(b, col)), (chunks, col)
            assert a.L == b.L and a.s0 == b.s0
    ctx.set_option("geom_chunks", 0)
