"""CPU tests: the C-ABI library loads and exports every declared symbol, host-side logic, sharding over gloo."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import practice_path_planning_for_formula_student_driverless_b200 as rl
from conftest import ROOT
from practice_path_planning_for_formula_student_driverless_b200 import _lib, sharding
from practice_path_planning_for_formula_student_driverless_b200._abi import RlJobStats, RlParams


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "raceline_b200.h")).read()
    declared = set(re.findall(r"^[a-z_ \*0-9]*?\b(rl_[a-z0-9_]+)\(", hdr, flags=re.M))
    assert declared == set(_lib.ABI_SYMBOLS), declared ^ set(_lib.ABI_SYMBOLS)
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.rl_abi_version() == rl.RL_ABI_VERSION
    assert L.rl_status_string(rl.RL_ERR_UNSUPPORTED) == b"unsupported problem shape"


def test_struct_layouts_match_header():
    assert C.sizeof(RlParams) == 22 * 8 + 6 * 4
    assert C.sizeof(RlJobStats) == 8 * 4 + 8 + 8 + 3 * 8 * rl.RL_MAX_OUTER_LOG + 2 * 4 * rl.RL_MAX_OUTER_LOG
    assert RlJobStats.lap_time.offset == 40 and RlJobStats.J0.offset == 48


def test_default_params_match_reference_config():
    p = RlParams()
    assert _lib.lib().rl_default_params(C.byref(p)) == 0
    q = rl.Config().to_params()
    from oracle import oracle
    o = oracle.default_params()
    for name, _ in RlParams._fields_:
        assert getattr(p, name) == getattr(q, name) == getattr(o, name), name
    assert abs(p.a_total_max - 1.17 * 9.81) < 1e-15 and p.max_outer_iters == 14 and p.max_inner_iters == 120


def test_no_device_fails_loudly_without_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(rl.RacelineError) as e:
        rl.Context(0)
    assert e.value.status == rl.RL_ERR_NODEVICE


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "practice_path_planning_for_formula_student_driverless_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("no CPU fallback", ""), f


def test_edges_and_packing():
    ring = np.array([[0.0, 0], [1, 0], [1, 1]])
    e = rl.ring_edges(ring)
    assert e.shape == (3, 4) and np.array_equal(e[2], [1, 1, 0, 0])
    assert rl.polyline_edges(ring).shape == (2, 4)
    t0 = rl.Track(np.zeros((5, 2)), e, e[:2], 5.0)
    t1 = rl.Track(np.ones((3, 2)), e[:1], e, 3.0)
    pb = rl.PackedBatch([t0, t1], [rl.Config().to_params()], [(1, 0, 1), (0, 0, 2), (1, 0, 2)])
    assert list(pb.samp_off) == [0, 5, 8] and list(pb.seg_off) == [0, 3, 5, 6, 9]
    assert list(pb.job_off) == [0, 3, 8, 11] and pb.rows == 11
    off = (C.c_int64 * 4)()
    assert _lib.lib().rl_job_sample_offsets(C.byref(pb.desc), off) == 0 and list(off) == [0, 3, 8, 11]


def test_synth_tracks_deterministic_and_well_formed():
    a = rl.synth_tracks(3, 512, seed_base=99)
    b = rl.synth_tracks(3, 512, seed_base=99, threads=1)
    c = rl.synth_tracks(1, 512, seed_base=99, first_id=2)
    for x, y in zip(a[:3], b[:3]):
        assert np.array_equal(x, y)
    assert np.array_equal(a[0].reshape(3, 512, 2)[2], c[0].reshape(512, 2))
    center, seg, L, m = a
    assert m == round(512 / 2.2)
    P = center.reshape(3, 512, 2)[0]
    d = np.linalg.norm(np.roll(P, -1, axis=0) - P, axis=1)
    assert abs(L[0] - 512 * 1.8) < 1e-2 and d.max() < 1.81 and d.min() > 1.7
    s = seg.reshape(3, 2, m, 4)[0]
    assert np.allclose(s[0, :-1, 2:], s[0, 1:, :2]) and np.allclose(s[0, -1, 2:], s[0, 0, :2])   # closed ring edges
    w = np.linalg.norm(s[0, :, :2] - s[1, :, :2], axis=1)
    assert np.allclose(w, 3.5, atol=1e-9)


def test_shard_bounds_cover_and_invert():
    for n in (0, 1, 7, 4096, 65536):
        for g in (1, 2, 3, 4, 8):
            cuts = [sharding.shard_bounds(n, g, r) for r in range(g)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[r][1] == cuts[r + 1][0] for r in range(g - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1
            for p in (0, n // 3, n - 1):
                if 0 <= p < n:
                    a, b = sharding.shard_bounds(n, g, sharding.owner_of(p, n, g))
                    assert a <= p < b


def _gloo_worker(rank, world, port, n, q):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = sharding.shard_bounds(n, world, rank)
    laps = 20.0 + ((np.arange(lo, hi) * 7919) % 101) * 0.01      # "lap time" of each owned problem
    allv = sharding.gather_lap_times(laps, n)
    best = sharding.best_of_sweep(laps, n)
    q.put((rank, allv, best))
    dist.barrier()
    dist.destroy_process_group()


def test_final_gather_over_gloo_world_size_2():
    import torch.multiprocessing as mp
    n, world, port = 37, 2, 29731
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = 20.0 + ((np.arange(n) * 7919) % 101) * 0.01
    for _, allv, best in got:
        assert np.array_equal(allv, expect)
        assert best[0] == int(np.argmin(expect)) and best[1] == expect.min()


def test_library_reads_no_environment_variables():
    """Tuning knobs / test hooks go through rl_set_option; the C ABI never calls getenv."""
    csrc = os.path.join(ROOT, "practice_path_planning_for_formula_student_driverless_b200", "csrc")
    for f in os.listdir(csrc):
        if f.endswith((".cu", ".cuh", ".h", ".cpp")):
            assert "getenv" not in open(os.path.join(csrc, f)).read(), f


def test_reference_arm_generator_equals_the_product_generator():
    """bench.py --impl reference makes its tracks with oracle/libsynth_tracks.so (the same source file compiled without
    CUDA) so that the CUDA library is not loaded in that arm: same seed, same bits."""
    from oracle import oracle
    c0, s0, L0, m = rl.synth_tracks(3, 300, 136, seed_base=0xB200, first_id=5)
    c1, s1, L1 = oracle.synth_tracks(3, 300, 136, seed_base=0xB200, first_id=5)
    assert np.array_equal(c0, c1) and np.array_equal(s0, s1) and np.array_equal(L0, L1)


def test_ref_harness_solve_reports_the_logged_backtracks(tmp_path):
    """RLR2 result files carry the backtracks the reference logs ('[PG] .. bt=' lines, min-time only): bench.py's
    `parity.bt_equal` rests on them."""
    from conftest import load_golden
    from oracle import batchfile, oracle
    exe = oracle.ref_binary("ref_harness")
    if not exe:
        pytest.skip("oracle/_ref/ref_harness not built (needs /root/reference at build time)")
    import subprocess
    g = load_golden("training_map")
    n, mi, mo = g["center_xy"].shape[0], g["inner_seg"].shape[0], g["outer_seg"].shape[0]
    p = oracle.default_params()
    row = np.array([float(getattr(p, k)) for k in batchfile.PARAM_FIELDS])
    bf, out = str(tmp_path / "b.bin"), str(tmp_path / "r.bin")
    batchfile.write_rlb1(bf, [0, n], [0, mi, mi + mo], [g["L"]], [1], g["center_xy"], np.concatenate([g["inner_seg"], g["outer_seg"]]),
                         row[None, :], [[0, 0, 1], [0, 0, 2]])
    subprocess.run([exe, "solve", bf, out], check=True, capture_output=True)
    res = batchfile.read_rlr1(out)
    assert res[0]["backtracks"] == -1                       # the reference logs nothing for min-curv
    assert res[1]["backtracks"] == int(g["mt_bt"].sum())
    assert np.array_equal(res[1]["alpha_total"], g["mt_alpha_total"]) and res[1]["lap_time"] == g["mt_lap_time"]


def test_pinned_pool_refuses_to_free_memory_under_live_arrays():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("pinned allocations need a CUDA driver")
    pool = rl.PinnedPool()
    a = pool.empty((4, 2), np.float64)
    with pytest.raises(RuntimeError):
        pool.close()
    del a
    pool.close()


def test_size_class_plan_boundaries():
    """rl_plan_for_track: the host plan's choice of kernel per track length (the same class_for_n / cluster_size_for_n
    calls plan_batch makes), at every boundary: K = 4 classes for 128 .. 1024 samples (every thread of a multi-warp CTA
    owns at least two samples), K = 8 elsewhere, clusters of 4 / 8 / 16 CTAs for long closed tracks."""
    L = _lib.lib()

    def plan(n, closed=1, max_cs=16):
        t, k, cs = C.c_int32(), C.c_int32(), C.c_int32()
        st = L.rl_plan_for_track(n, closed, max_cs, C.byref(t), C.byref(k), C.byref(cs))
        return st, t.value, k.value, cs.value

    for n, want in [(0, (32, 8)), (1, (32, 8)), (127, (32, 8)), (128, (64, 4)), (216, (64, 4)), (256, (64, 4)), (257, (128, 4)),
                    (261, (128, 4)), (512, (128, 4)), (513, (256, 4)), (1024, (256, 4)), (1025, (256, 8)), (2048, (256, 8)),
                    (2049, (512, 8)), (4096, (512, 8))]:
        for closed in (0, 1):
            assert plan(n, closed) == (0, want[0], want[1], 0), (n, closed)
    assert plan(4097) == (0, 256, 8, 4) and plan(8192) == (0, 256, 8, 4) and plan(8193) == (0, 256, 8, 8)
    assert plan(16384) == (0, 256, 8, 8) and plan(16385) == (0, 256, 8, 16) and plan(32768) == (0, 256, 8, 16)
    assert plan(16385, max_cs=8)[0] == rl.RL_ERR_UNSUPPORTED and plan(32769)[0] == rl.RL_ERR_UNSUPPORTED
    assert plan(4097, closed=0) == (0, 256, 8, 4) and plan(20000, closed=0) == (0, 256, 8, 16)   # open tracks take clusters too
    assert plan(-1)[0] == rl.RL_ERR_ARG
    assert L.rl_plan_for_track(100, 1, 8, None, None, None) == rl.RL_ERR_ARG


def test_every_option_of_rl_set_option_is_documented_and_reset_between_tests():
    """The option names rl_set_option accepts (csrc/raceline_api.cu) are exactly the ones the header documents, the
    Python wrapper's docstring lists them, and the GPU test fixture puts every plan option back to automatic."""
    src = open(os.path.join(ROOT, "practice_path_planning_for_formula_student_driverless_b200", "csrc", "raceline_api.cu")).read()
    body = src[src.index("int rl_set_option("):src.index("const char* rl_last_error")]
    accepted = set(re.findall(r'strcmp\(name, "([a-z_]+)"\)', body))
    hdr = open(os.path.join(ROOT, "include", "raceline_b200.h")).read()
    doc = hdr[hdr.index("Tuning knobs and test hooks"):hdr.index("int rl_set_option(")]
    documented = set(re.findall(r'^ \*   "([a-z_]+)"', doc, flags=re.M))
    assert accepted == documented, accepted ^ documented
    assert {"solve_chunks", "chunk_streams", "geom_chunks"} <= accepted
    conf = open(os.path.join(ROOT, "tests", "conftest.py")).read()
    reset = set(re.findall(r'"([a-z_]+)"', conf[conf.index("def _plan_options_back_to_automatic"):conf.index("def angle_diff")]))
    assert accepted - {"no_few_search", "debug_inject"} <= reset, accepted - reset
