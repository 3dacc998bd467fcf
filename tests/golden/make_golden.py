#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container only).

For every shipped map this runs oracle/_ref/ref_harness[_instr] (which #include the reference
src/main.cpp): the reference front end produces the hot-path inputs (centre samples, ring segments,
L) and the reference solvers produce the outputs.  Before anything is written, the C restatement
(oracle/raceline_oracle.c) must reproduce those outputs BIT FOR BIT -- that is the oracle's pin.

Also written:
  sweep_competition_map2.npz  -- 16 non-default Config combos (lambda_smooth, a_total_max=mu*9.81,
                                 P_max_W, w_time_gain, plus the inv-v weighting switch) solved by the
                                 reference through `ref_harness solve`; pins the port off-default.
  competition_map2_n1000.npz  -- the same map with samples forced to 1000 (a larger, ragged N).
  open_<map>.npz              -- open-track mode (closed=false, polyline edges), solved by the reference.

Usage:  python tests/golden/make_golden.py         (needs /root/reference)
"""
from __future__ import annotations

import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import batchfile, oracle  # noqa: E402
from practice_path_planning_for_formula_student_driverless_b200._abi import (  # noqa: E402
    RL_STAGE_MINCURV,
    RL_STAGE_MINTIME,
)

REF = "/root/reference"
MAPS = ["training_map", "competition_map1", "competition_map2", "competition_map3",
        "competition_map_testday1", "competition_map_testday2", "competition_map_testday3"]
OUT = os.path.join(ROOT, "tests", "golden")
FIELDS_MC = ("xy", "heading", "curvature", "alpha_total", "alpha_last")
FIELDS_MT = FIELDS_MC + ("v", "ax")


def params_row(p):
    return np.array([float(getattr(p, k)) for k in batchfile.PARAM_FIELDS])


def check_port_bitwise(d, tag):
    """The C port must equal the reference bit for bit on this map."""
    p = oracle.default_params()
    for stage, pre, fields in ((RL_STAGE_MINCURV, "mc", FIELDS_MC), (RL_STAGE_MINTIME, "mt", FIELDS_MT)):
        r = oracle.solve(stage, d["center_xy"], d["inner_seg"], d["outer_seg"], d["L"], d["closed"], p)
        for k in fields:
            a, b = r[k], d[f"{pre}_{k}"]
            if not np.array_equal(a.view(np.uint64), b.view(np.uint64)):
                raise SystemExit(f"[{tag}] oracle port differs from reference in {pre}_{k}: max|d|={np.abs(a-b).max():.3e}")
        st = r["stats"]
        bt_ref = d["mc_bt"] if pre == "mc" else d["mt_bt"]
        if st.accepted != bt_ref.size or st.backtracks != int(bt_ref.sum()):
            raise SystemExit(f"[{tag}] {pre}: accepted/backtracks {st.accepted}/{st.backtracks} vs reference "
                             f"{bt_ref.size}/{int(bt_ref.sum())}")
        per_outer = [int(x) for x in st.bt_outer[:st.outer_done]]
        ref_outer = [int(bt_ref[120 * o:120 * (o + 1)].sum()) for o in range(st.outer_done)] if bt_ref.size == 120 * st.outer_done else None
        if ref_outer is not None and per_outer != ref_outer:
            raise SystemExit(f"[{tag}] {pre}: per-outer backtracks differ")
        if pre == "mt" and r["lap_time"] != d["mt_lap_time"]:
            raise SystemExit(f"[{tag}] lap time differs")
    return True


def run_frontend(name, samples=0):
    with tempfile.TemporaryDirectory() as td:
        outs = []
        for exe in ("ref_harness", "ref_harness_instr"):
            o = os.path.join(td, exe + ".bin")
            cmd = [os.path.join(ROOT, "oracle", "_ref", exe), "frontend", f"{REF}/csv/{name}_inner.csv",
                   f"{REF}/csv/{name}_outer.csv", o] + ([str(samples)] if samples else [])
            res = subprocess.run(cmd, check=True, capture_output=True, text=True)
            print(f"  {exe}: {res.stdout.strip()}")
            outs.append(batchfile.read_rlg1(o))
    plain, instr = outs
    for k, v in plain.items():
        if k in ("mc_bt",):
            continue
        same = np.array_equal(v, instr[k]) if isinstance(v, np.ndarray) else v == instr[k]
        assert same, f"instrumented reference build changed {k}"
    plain["mc_bt"] = instr["mc_bt"]  # only the instrumented build logs min-curv backtracks
    return plain


def sweep_golden(base):
    """16 off-default combos on competition_map2, solved by the reference itself."""
    lam = [4e-4, 1.6e-3, 6.4e-3, 3.2e-3]
    mu = [1.15, 1.3, 1.45, 1.6]
    pmax = [20e3, 40e3, 60e3, 80e3]
    wt = [0.0, 1.0, 2.0, 3.5]
    rows, meta = [], []
    for i in range(16):
        p = oracle.default_params()
        p.lambda_smooth = lam[i % 4]
        p.a_total_max = mu[(i // 4) % 4] * 9.81
        p.P_max_W = pmax[(i * 7 + 1) % 4]
        p.w_time_gain = wt[(i * 5 + 2) % 4]
        if i % 5 == 0:
            p.time_weight_use_inv_v = 1
            p.inv_v_gain = 0.3
        if i == 3:
            p.time_gamma_power = 1.5
        if i == 7:
            p.use_total_ge_lat = 0
        rows.append(params_row(p))
        meta.append(p)
    rows = np.stack(rows)
    n, mi, mo = base["n"], base["m_inner"], base["m_outer"]
    jobs = np.array([[0, i, st] for i in range(16) for st in (RL_STAGE_MINCURV, RL_STAGE_MINTIME)])
    with tempfile.TemporaryDirectory() as td:
        bf, rf = os.path.join(td, "b.bin"), os.path.join(td, "r.bin")
        batchfile.write_rlb1(bf, [0, n], [0, mi, mi + mo], [base["L"]], [base["closed"]], base["center_xy"],
                             np.concatenate([base["inner_seg"], base["outer_seg"]]), rows, jobs)
        subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_harness"), "solve", bf, rf], check=True,
                       capture_output=True)
        res = batchfile.read_rlr1(rf)
    out = {"params_rows": rows, "jobs": jobs}
    for j, r in enumerate(res):
        p = meta[jobs[j, 1]]
        o = oracle.solve(int(jobs[j, 2]), base["center_xy"], base["inner_seg"], base["outer_seg"], base["L"], base["closed"], p)
        for k in (FIELDS_MT if jobs[j, 2] == RL_STAGE_MINTIME else FIELDS_MC):
            if not np.array_equal(o[k].view(np.uint64), r[k].view(np.uint64)):
                raise SystemExit(f"[sweep job {j}] oracle port differs from reference in {k}: {np.abs(o[k]-r[k]).max():.3e}")
        if jobs[j, 2] == RL_STAGE_MINTIME and o["lap_time"] != r["lap_time"]:
            raise SystemExit(f"[sweep job {j}] lap differs")
        for k in FIELDS_MT:
            out[f"j{j}_{k}"] = r[k]
        out[f"j{j}_lap_time"] = r["lap_time"]
        out[f"j{j}_accepted"] = o["stats"].accepted
        out[f"j{j}_backtracks"] = o["stats"].backtracks
    return out


def open_golden(base):
    """Open-track mode (cfg is_closed_track=false, main.cpp:54): the same map driven as an open path with
    polyline edges (edges::polylineEdges, main.cpp:256), solved by the reference itself."""
    n = base["n"]
    inner = np.concatenate([base["inner_seg"][:-1, :2], base["inner_seg"][1:, :2]], axis=1)
    outer = np.concatenate([base["outer_seg"][:-1, :2], base["outer_seg"][1:, :2]], axis=1)
    p = oracle.default_params()
    jobs = np.array([[0, 0, RL_STAGE_MINCURV], [0, 0, RL_STAGE_MINTIME]])
    with tempfile.TemporaryDirectory() as td:
        bf, rf = os.path.join(td, "b.bin"), os.path.join(td, "r.bin")
        batchfile.write_rlb1(bf, [0, n], [0, len(inner), len(inner) + len(outer)], [base["L"]], [0], base["center_xy"],
                             np.concatenate([inner, outer]), params_row(p)[None, :], jobs)
        subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_harness"), "solve", bf, rf], check=True, capture_output=True)
        res = batchfile.read_rlr1(rf)
    out = {"center_xy": base["center_xy"], "inner_seg": inner, "outer_seg": outer, "L": base["L"], "n": n}
    for pre, r, stage in (("mc", res[0], RL_STAGE_MINCURV), ("mt", res[1], RL_STAGE_MINTIME)):
        o = oracle.solve(stage, base["center_xy"], inner, outer, base["L"], False, p)
        for k in (FIELDS_MT if pre == "mt" else FIELDS_MC):
            if not np.array_equal(o[k].view(np.uint64), r[k].view(np.uint64)):
                raise SystemExit(f"[open {pre}] oracle port differs from reference in {k}: {np.abs(o[k]-r[k]).max():.3e}")
            out[f"{pre}_{k}"] = r[k]
        out[f"{pre}_accepted"] = o["stats"].accepted
        out[f"{pre}_backtracks"] = o["stats"].backtracks
    if oracle.solve(RL_STAGE_MINTIME, base["center_xy"], inner, outer, base["L"], False, p)["lap_time"] != res[1]["lap_time"]:
        raise SystemExit("[open] lap differs")
    out["mt_lap_time"] = res[1]["lap_time"]
    return out


def geom_golden(name, samples=0, open_track=False):
    """The centre-line + width/geometry stage through the reference's own functions (`ref_harness geom`, which also
    checks itself against the *_with_geom.csv the reference writes); pins geom_oracle.c bit for bit."""
    with tempfile.TemporaryDirectory() as td:
        o = os.path.join(td, "g.bin")
        cmd = [os.path.join(ROOT, "oracle", "_ref", "ref_harness"), "geom", f"{REF}/csv/{name}_inner.csv",
               f"{REF}/csv/{name}_outer.csv", o] + ([str(samples)] if (samples or open_track) else []) + (["open"] if open_track else [])
        res = subprocess.run(cmd, check=True, capture_output=True, text=True)
        print(f"  geom: {res.stdout.strip()}")
        d = batchfile.read_rgm1(o)
    r = oracle.centerline_geom(d["mids_xy"], d["samples"], d["inner_seg"], d["outer_seg"], bool(d["closed"]),
                               bool(d["emit_closed_duplicate"]))
    pairs = [("L", r["L"], d["L"]), ("s0", r["s0"], d["s0"]), ("x", r["xy"][:, 0], d["x"]), ("y", r["xy"][:, 1], d["y"])]
    pairs += [(k, r[k], d[k]) for k in ("s_rel", "curvature", "dist_inner", "dist_outer", "width", "v_kappa")]
    for k, a, b in pairs:
        if not np.array_equal(np.asarray(a), np.asarray(b)):
            raise SystemExit(f"[geom {name}] port differs from the reference in {k}: max |d| = {np.max(np.abs(np.asarray(a) - np.asarray(b))):.3e}")
    # heading goes through libm atan2 on both sides here, so it is bitwise too
    if not np.array_equal(r["heading"], d["heading"]):
        raise SystemExit(f"[geom {name}] heading differs")
    assert np.array_equal(d["center_xy"][:d["samples"], 0], d["x"][:d["samples"]])   # centre line == geometry x,y
    return d


def debug_golden(name):
    """The reference's debug comparison (`ref_harness debug`): the CSV it writes + the two extra laps; pins the EVAL stage
    of the C restatement (heading/curvature + v(s) profile of a given path) bit for bit."""
    with tempfile.TemporaryDirectory() as td:
        o = os.path.join(td, "d.bin")
        res = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_harness"), "debug", f"{REF}/csv/{name}_inner.csv",
                              f"{REF}/csv/{name}_outer.csv", o], check=True, capture_output=True, text=True)
        print(f"  debug: {res.stdout.strip()}")
        d = batchfile.read_rdg1(o)
    from practice_path_planning_for_formula_student_driverless_b200 import RL_STAGE_EVAL
    e = np.zeros((0, 4))
    lap_c = oracle.solve(RL_STAGE_EVAL, d["center_xy"], e, e, d["L"], True)["lap_time"]
    lap_m = oracle.solve(RL_STAGE_EVAL, d["mc_xy"], e, e, d["L_mc"], True)["lap_time"]
    if lap_c != d["lap_center"] or lap_m != d["lap_mincurv"]:
        raise SystemExit(f"[debug {name}] EVAL stage of the port differs: {lap_c} vs {d['lap_center']}, {lap_m} vs {d['lap_mincurv']}")
    return d


def main():
    if not os.path.exists(f"{REF}/src/main.cpp"):
        raise SystemExit("the reference is not mounted; goldens can only be regenerated in the build container")
    oracle.build(force=True)
    oracle.build_ref()
    for name in MAPS:
        print(name)
        d = run_frontend(name)
        check_port_bitwise(d, name)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print("competition_map2 @ samples=1000")
    d = run_frontend("competition_map2", samples=1000)
    check_port_bitwise(d, "competition_map2_n1000")
    np.savez_compressed(os.path.join(OUT, "competition_map2_n1000.npz"), **d)
    print("config sweep on competition_map2")
    base = dict(np.load(os.path.join(OUT, "competition_map2.npz")))
    base = {k: (v.item() if v.ndim == 0 else v) for k, v in base.items()}
    np.savez_compressed(os.path.join(OUT, "sweep_competition_map2.npz"), **sweep_golden(base))
    print("open-track mode on competition_map1 and training_map")
    for name in ("competition_map1", "training_map"):
        b = dict(np.load(os.path.join(OUT, name + ".npz")))
        b = {k: (v.item() if v.ndim == 0 else v) for k, v in b.items()}
        np.savez_compressed(os.path.join(OUT, f"open_{name}.npz"), **open_golden(b))
    print("centre line + width/geometry stage (SURVEY 8f rows 1-2)")
    for name, smp in (("training_map", 0), ("competition_map1", 0), ("competition_map_testday2", 0), ("competition_map2", 1000)):
        d = geom_golden(name, smp)
        np.savez_compressed(os.path.join(OUT, f"geom_{name}" + (f"_n{smp}" if smp else "") + ".npz"), **d)
    d = geom_golden("competition_map3", 0, open_track=True)      # cfg is_closed_track = false: no padding, polyline edges
    np.savez_compressed(os.path.join(OUT, "geom_open_competition_map3.npz"), **d)
    print("debug comparison of the min-time stage wrapper (SURVEY 8f row 3)")
    for name in ("training_map", "competition_map3"):
        np.savez_compressed(os.path.join(OUT, f"debug_{name}.npz"), **debug_golden(name))
    print("oracle port == reference, bit for bit, on every golden case")


if __name__ == "__main__":
    main()
