"""TEST INFRASTRUCTURE: binary file formats shared by oracle/ref_harness.cpp, tests and bench.py.

RLG1 (one shipped map, written by `ref_harness frontend`): hot-path inputs produced by the
reference front end (main.cpp:1613-1663) plus both solver outputs (main.cpp:677-681, 897-903).
RLB1 (a packed batch = the rl_batch_desc layout of include/raceline_b200.h) and RLR1 (results +
per-job solver wall-clock) are the `ref_harness solve` input and output.
"""
from __future__ import annotations

import numpy as np

MAGIC_G = 0x31474C52
MAGIC_B = 0x31424C52
MAGIC_R = 0x31524C52
MAGIC_R2 = 0x32524C52   # RLR1 + per-job backtracks as logged by the reference ("[PG] .. bt=", main.cpp:1019-1022)

# rl_params field order (include/raceline_b200.h); ints are stored as doubles in RLB1
PARAM_FIELDS = [
    "veh_width_arg", "veh_width_m", "safety_margin_m", "lambda_smooth", "step_init", "step_min",
    "armijo_c", "kappa_eps", "v_cap_mps", "mass_kg", "Cd", "A_front_m2", "rho_air", "c_rr", "P_max_W",
    "a_total_max", "a_lat_max", "a_long_acc_cap", "a_long_brake_cap", "w_time_gain",
    "time_gamma_power", "inv_v_gain", "max_outer_iters", "max_inner_iters", "max_vpass_iters",
    "time_weight_use_inv_v", "use_total_ge_lat", "reserved",
]


def read_rlg1(path):
    """Return a dict of numpy arrays for one `ref_harness frontend` dump."""
    raw = np.fromfile(path, dtype=np.uint8)
    hdr = raw[:40].view(np.int64)
    assert hdr[0] == MAGIC_G, "not an RLG1 file"
    n, m_in, m_out, closed = (int(hdr[1]), int(hdr[2]), int(hdr[3]), int(hdr[4]))
    pos = 40

    def f64(count):
        nonlocal pos
        out = raw[pos:pos + 8 * count].view(np.float64).copy()
        pos += 8 * count
        return out

    def i64(count):
        nonlocal pos
        out = raw[pos:pos + 8 * count].view(np.int64).copy()
        pos += 8 * count
        return out

    d = {"n": n, "m_inner": m_in, "m_outer": m_out, "closed": closed}
    d["L"] = float(f64(1)[0])
    d["s0"] = float(f64(1)[0])
    d["center_xy"] = f64(2 * n).reshape(n, 2)
    d["inner_seg"] = f64(4 * m_in).reshape(m_in, 4)
    d["outer_seg"] = f64(4 * m_out).reshape(m_out, 4)
    for st, extra in (("mc", ()), ("mt", ("v", "ax"))):
        d[st + "_xy"] = f64(2 * n).reshape(n, 2)
        for k in ("heading", "curvature", "alpha_total", "alpha_last") + extra:
            d[st + "_" + k] = f64(n)
    d["mt_lap_time"] = float(f64(1)[0])
    k = int(i64(1)[0])
    d["mc_bt"] = i64(k)
    k = int(i64(1)[0])
    d["mt_bt"] = i64(k)
    assert pos == raw.size, "trailing bytes in RLG1 file"
    return d


def write_rlb1(path, samp_off, seg_off, track_L, track_closed, center_xy, seg, params_rows, jobs):
    """params_rows: (n_params, 28) float64 in PARAM_FIELDS order; jobs: (n_jobs, 3) int (track, param, stage)."""
    samp_off = np.ascontiguousarray(samp_off, dtype=np.int64)
    seg_off = np.ascontiguousarray(seg_off, dtype=np.int64)
    params_rows = np.ascontiguousarray(params_rows, dtype=np.float64).reshape(-1, len(PARAM_FIELDS))
    jobs = np.ascontiguousarray(jobs, dtype=np.int64).reshape(-1, 3)
    nt = samp_off.size - 1
    with open(path, "wb") as f:
        np.array([MAGIC_B, nt, params_rows.shape[0], jobs.shape[0]], dtype=np.int64).tofile(f)
        samp_off.tofile(f)
        seg_off.tofile(f)
        np.ascontiguousarray(track_L, dtype=np.float64).tofile(f)
        np.ascontiguousarray(track_closed, dtype=np.int64).tofile(f)
        np.ascontiguousarray(center_xy, dtype=np.float64).tofile(f)
        np.ascontiguousarray(seg, dtype=np.float64).tofile(f)
        params_rows.tofile(f)
        jobs.tofile(f)


def read_rlr1(path):
    """Return a list of per-job dicts from a `ref_harness solve` result file."""
    raw = np.fromfile(path, dtype=np.uint8)
    hdr = raw[:16].view(np.int64)
    assert hdr[0] in (MAGIC_R, MAGIC_R2), "not an RLR1/RLR2 file"
    v2 = hdr[0] == MAGIC_R2
    pos = 16
    out = []
    for _ in range(int(hdr[1])):
        n, stage = (int(x) for x in raw[pos:pos + 16].view(np.int64))
        ms, lap = (float(x) for x in raw[pos + 16:pos + 32].view(np.float64))
        pos += 32
        bt = -1
        if v2:
            bt = int(raw[pos:pos + 8].view(np.int64)[0])
            pos += 8
        body = raw[pos:pos + 8 * 8 * n].view(np.float64)
        pos += 8 * 8 * n
        r = {"n": n, "stage": stage, "wall_ms": ms, "lap_time": lap, "backtracks": bt, "xy": body[:2 * n].reshape(n, 2).copy()}
        for i, k in enumerate(("heading", "curvature", "alpha_total", "alpha_last", "v", "ax")):
            r[k] = body[(2 + i) * n:(3 + i) * n].copy()
        out.append(r)
    assert pos == raw.size
    return out


MAGIC_M = 0x314D4752  # "RGM1": one `ref_harness geom` dump


def read_rgm1(path):
    """Return a dict of numpy arrays for one `ref_harness geom` dump (centre line + width/geometry stage)."""
    raw = np.fromfile(path, dtype=np.uint8)
    hdr = raw[:64].view(np.int64)
    assert hdr[0] == MAGIC_M, "not an RGM1 file"
    n_mid, samples, rows, m_in, m_out, closed, emit = (int(x) for x in hdr[1:8])
    pos = 64

    def f64(count):
        nonlocal pos
        out = raw[pos:pos + 8 * count].view(np.float64).copy()
        pos += 8 * count
        return out

    d = {"n_mid": n_mid, "samples": samples, "rows": rows, "m_inner": m_in, "m_outer": m_out, "closed": closed,
         "emit_closed_duplicate": emit}
    d["L"] = float(f64(1)[0])
    d["s0"] = float(f64(1)[0])
    d["mids_xy"] = f64(2 * n_mid).reshape(n_mid, 2)
    d["inner_seg"] = f64(4 * m_in).reshape(m_in, 4)
    d["outer_seg"] = f64(4 * m_out).reshape(m_out, 4)
    nc = samples + (1 if emit else 0)
    d["center_xy"] = f64(2 * nc).reshape(nc, 2)
    for k in ("s_rel", "x", "y", "heading", "curvature", "dist_inner", "dist_outer", "width", "v_kappa"):
        d[k] = f64(rows)
    assert pos == raw.size, "trailing bytes in RGM1 file"
    return d


MAGIC_D = 0x31474452  # "RDG1": one `ref_harness debug` dump


def read_rdg1(path):
    """Return a dict for one `ref_harness debug` dump (the reference's debug comparison, main.cpp:1440-1593)."""
    raw = np.fromfile(path, dtype=np.uint8)
    hdr = raw[:48].view(np.int64)
    assert hdr[0] == MAGIC_D, "not an RDG1 file"
    n, n_mc, rows, m_in, m_out = (int(x) for x in hdr[1:6])
    pos = 48

    def f64(count):
        nonlocal pos
        out = raw[pos:pos + 8 * count].view(np.float64).copy()
        pos += 8 * count
        return out

    d = {"n": n, "rows": rows}
    for k in ("L", "s0", "lap_center", "lap_mincurv", "lap_mintime", "L_mc"):
        d[k] = float(f64(1)[0])
    d["center_xy"] = f64(2 * n).reshape(n, 2)
    d["mc_xy"] = f64(2 * n_mc).reshape(n_mc, 2)
    d["inner_seg"] = f64(4 * m_in).reshape(m_in, 4)
    d["outer_seg"] = f64(4 * m_out).reshape(m_out, 4)
    d["columns"] = f64(20 * rows).reshape(rows, 20)
    assert pos == raw.size, "trailing bytes in RDG1 file"
    return d
