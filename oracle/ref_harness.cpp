// ref_harness.cpp -- TEST INFRASTRUCTURE.  Drives the UNMODIFIED reference
// translation unit (REF_MAIN_CPP, by default /root/reference/src/main.cpp) as a
// function-level oracle and as the CPU baseline.  The reference source is
// #included where it lies; nothing of it is copied into this repository, and
// the binary is only ever written to oracle/_ref/ (git-ignored).
//
//   ref_harness frontend <inner.csv> <outer.csv> <out.bin> [samples]
//       runs the reference front end (main.cpp:1613-1663: load, DT, mids, MST,
//       rings, spline) and both solver stages exactly as main() does
//       (main.cpp:1678-1696) and dumps hot-path inputs + outputs as raw doubles.
//       If [samples] > 0 the dynamic sample count (main.cpp:1638-1641) is replaced by it.
//   ref_harness geom <inner.csv> <outer.csv> <out.bin> [samples [open]]
//       runs the reference front end up to the ordered mid points and the rings, then pipeline::make_centerline
//       (main.cpp:1270) and the per-sample body of pipeline::compute_geom_and_save (main.cpp:1311-1329) through the
//       reference's own Spline1D / distancesToRings, dumps inputs + outputs as raw doubles, and checks them against
//       the *_with_geom.csv the reference itself writes (9 decimals).
//   ref_harness debug <inner.csv> <outer.csv> <out.bin>
//       runs the two stage wrappers with cfg debug_dump = true exactly as main() does (main.cpp:1684, 1694) and dumps
//       the 20 columns of the <base>_debug_compare_paths.csv the reference writes (main.cpp:1493-1495), the min-curv
//       path it re-reads from <base>_raceline.csv, and the centre-line / min-curv laps of main.cpp:1464-1477 recomputed
//       through the reference's own functions in full precision.
//   ref_harness solve <batch.bin> <out.bin> [first_job] [n_jobs]
//       solves jobs of a packed batch file (format: oracle/batchfile.py) with
//       cfg::get() set per job from the job's parameter row; records the
//       wall-clock of each solver call (stderr discarded, the "[PG]" lines at
//       main.cpp:1019 are unconditional).
//
// Built twice by oracle/Makefile: once on the pristine source, once on a sed
// copy in oracle/_ref/ whose only change is a "[MCBT] outer it bt" stderr line
// after main.cpp:739 (the reference logs backtracks for min-time only).
#ifndef REF_MAIN_CPP
#define REF_MAIN_CPP "/root/reference/src/main.cpp"
#endif
#define main ref_main
#include REF_MAIN_CPP
#undef main

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace {

struct NullBuf : std::streambuf { int overflow(int c) override { return c; } };

// Discards the reference's stderr chatter but keeps count of the Armijo backtracks it reports: the "bt=<int>" field of
// every "[PG]" line (min-time, main.cpp:1019-1022) and the third integer of "[MCBT] outer it bt" lines (instrumented
// build only).  Cheaper than accumulating the log: one short line buffer.
struct BtCountBuf : std::streambuf {
    std::string line;
    long long bt = 0, lines = 0;
    void reset() { line.clear(); bt = 0; lines = 0; }
    void end_line()
    {
        if (line.find("[PG]") != std::string::npos) {
            const size_t q = line.rfind("bt=");
            if (q != std::string::npos) { bt += std::atoll(line.c_str() + q + 3); ++lines; }
        } else {
            const size_t p = line.find("[MCBT]");
            long long o, it, b;
            if (p != std::string::npos && std::sscanf(line.c_str() + p + 6, "%lld %lld %lld", &o, &it, &b) == 3) { bt += b; ++lines; }
        }
        line.clear();
    }
    int overflow(int c) override
    {
        if (c == '\n') end_line(); else if (c != EOF && line.size() < 512) line.push_back((char)c);
        return c;
    }
    std::streamsize xsputn(const char* s, std::streamsize n) override
    {
        for (std::streamsize i = 0; i < n; ++i) overflow((unsigned char)s[i]);
        return n;
    }
};

void put_i64(FILE* f, int64_t v) { fwrite(&v, 8, 1, f); }
void put_f64(FILE* f, double v) { fwrite(&v, 8, 1, f); }
void put_vec(FILE* f, const std::vector<double>& v) { if (!v.empty()) fwrite(v.data(), 8, v.size(), f); }
void put_pts(FILE* f, const std::vector<geom::Vec2>& v) { for (auto& p : v) { put_f64(f, p.x); put_f64(f, p.y); } }

// parse "bt=<int>" of every "[PG]" line and "[MCBT] outer it bt" lines
void parse_bt(const std::string& log, std::vector<int64_t>& pg_bt, std::vector<int64_t>& mc_bt)
{
    std::istringstream is(log);
    std::string line;
    while (std::getline(is, line)) {
        size_t p = line.find("[PG]");
        if (p != std::string::npos) {
            size_t q = line.rfind("bt=");
            if (q != std::string::npos) pg_bt.push_back(std::atoll(line.c_str() + q + 3));
            continue;
        }
        p = line.find("[MCBT]");
        if (p != std::string::npos) {
            long long o, it, bt;
            if (std::sscanf(line.c_str() + p + 6, "%lld %lld %lld", &o, &it, &bt) == 3) mc_bt.push_back(bt);
        }
    }
}

int run_frontend(int argc, char** argv)
{
    if (argc < 5) { std::fprintf(stderr, "usage: frontend inner.csv outer.csv out.bin [samples]\n"); return 1; }
    const std::string innerPath = argv[2], outerPath = argv[3], outBin = argv[4];
    const int forced_samples = (argc > 5) ? std::atoi(argv[5]) : 0;
    auto& C = cfg::get();
    C.verbose = false;
    C.debug_dump = false;
    char tmpl[] = "/tmp/ref_harness_XXXXXX";
    if (!mkdtemp(tmpl)) { std::perror("mkdtemp"); return 1; }
    const std::string base = std::string(tmpl) + "/centerline";

    std::ostringstream captured;
    std::streambuf* old = std::cerr.rdbuf(captured.rdbuf());

    // stages 1-6 exactly as main() sequences them
    auto inner = io::loadCSV_XY(innerPath);
    auto outer = io::loadCSV_XY(outerPath);
    if (inner.size() < 2 || outer.size() < 2) { std::cerr.rdbuf(old); std::fprintf(stderr, "need >=2 points per ring\n"); return 2; }
    const bool closed_mode = C.is_closed_track;
    auto tri = pipeline::buildDT(inner, outer);
    auto MF = pipeline::extract_mids_with_len_filter(tri, base);
    if (C.use_dynamic_samples) C.samples = pipeline::dynamic_samples_from_mids_count((int)MF.mids.size());
    if (forced_samples > 0) C.samples = forced_samples;
    auto OM = pipeline::order_and_align_mids_open_closed(MF.mids, closed_mode);
    auto RR = pipeline::reconstruct_rings_and_align(OM, MF, tri, base);
    auto CL = pipeline::make_centerline(OM, closed_mode, base);

    std::vector<geom::Vec2> center_for_opt = CL.center;
    if (closed_mode && center_for_opt.size() >= 2 && geom::almostEq(center_for_opt.front(), center_for_opt.back(), 1e-12))
        center_for_opt.pop_back();
    auto innerE = closed_mode ? edges::ringEdges(RR.inner_from_mids) : edges::polylineEdges(RR.inner_from_mids);
    auto outerE = closed_mode ? edges::ringEdges(RR.outer_from_mids) : edges::polylineEdges(RR.outer_from_mids);

    captured.str("");
    auto mc = raceline_min_curv::compute_min_curvature_raceline(center_for_opt, innerE, outerE, C.veh_width_m, CL.L, closed_mode);
    std::string log_mc = captured.str();
    captured.str("");
    auto mt = raceline_min_time::compute_min_time_raceline(center_for_opt, innerE, outerE, C.veh_width_m, CL.L, closed_mode);
    std::string log_mt = captured.str();
    std::cerr.rdbuf(old);

    std::vector<int64_t> pg_bt, mc_bt, dummy;
    parse_bt(log_mc, dummy, mc_bt);
    parse_bt(log_mt, pg_bt, dummy);

    FILE* f = std::fopen(outBin.c_str(), "wb");
    if (!f) { std::perror("fopen"); return 1; }
    const int64_t N = (int64_t)center_for_opt.size();
    put_i64(f, 0x31474C52);  // "RLG1"
    put_i64(f, N);
    put_i64(f, (int64_t)innerE.size());
    put_i64(f, (int64_t)outerE.size());
    put_i64(f, closed_mode ? 1 : 0);
    put_f64(f, CL.L);
    put_f64(f, CL.s0);
    put_pts(f, center_for_opt);
    for (auto& e : innerE) { put_f64(f, e.first.x); put_f64(f, e.first.y); put_f64(f, e.second.x); put_f64(f, e.second.y); }
    for (auto& e : outerE) { put_f64(f, e.first.x); put_f64(f, e.first.y); put_f64(f, e.second.x); put_f64(f, e.second.y); }
    put_pts(f, mc.raceline); put_vec(f, mc.heading); put_vec(f, mc.curvature); put_vec(f, mc.alpha_total); put_vec(f, mc.alpha_last);
    put_pts(f, mt.raceline); put_vec(f, mt.heading); put_vec(f, mt.curvature); put_vec(f, mt.alpha_total); put_vec(f, mt.alpha_last);
    put_vec(f, mt.v); put_vec(f, mt.ax);
    put_f64(f, mt.lap_time);
    put_i64(f, (int64_t)mc_bt.size()); for (auto b : mc_bt) put_i64(f, b);
    put_i64(f, (int64_t)pg_bt.size()); for (auto b : pg_bt) put_i64(f, b);
    std::fclose(f);
    std::string rm = std::string("rm -rf ") + tmpl;
    if (std::system(rm.c_str()) != 0) {}
    std::printf("N=%lld Min=%zu Mout=%zu L=%.9f lap=%.9f mc_steps=%zu mt_steps=%zu\n", (long long)N, innerE.size(), outerE.size(),
                CL.L, mt.lap_time, mc_bt.size(), pg_bt.size());
    return 0;
}


// the width/geometry stage, driven through the reference's own functions
int run_geom(int argc, char** argv)
{
    if (argc < 5) { std::fprintf(stderr, "usage: geom inner.csv outer.csv out.bin [samples [open]]\n"); return 1; }
    const std::string innerPath = argv[2], outerPath = argv[3], outBin = argv[4];
    const int forced_samples = (argc > 5) ? std::atoi(argv[5]) : 0;
    auto& C = cfg::get();
    C.verbose = false;
    C.debug_dump = false;
    if (argc > 6 && std::strcmp(argv[6], "open") == 0) C.is_closed_track = false;   // cfg is_closed_track, main.cpp:54
    char tmpl[] = "/tmp/ref_harness_XXXXXX";
    if (!mkdtemp(tmpl)) { std::perror("mkdtemp"); return 1; }
    const std::string base = std::string(tmpl) + "/centerline";
    std::ostringstream captured;
    std::streambuf* old = std::cerr.rdbuf(captured.rdbuf());
    auto inner = io::loadCSV_XY(innerPath);
    auto outer = io::loadCSV_XY(outerPath);
    const bool closed_mode = C.is_closed_track;
    auto tri = pipeline::buildDT(inner, outer);
    auto MF = pipeline::extract_mids_with_len_filter(tri, base);
    if (C.use_dynamic_samples) C.samples = pipeline::dynamic_samples_from_mids_count((int)MF.mids.size());
    if (forced_samples > 0) C.samples = forced_samples;
    auto OM = pipeline::order_and_align_mids_open_closed(MF.mids, closed_mode);
    auto RR = pipeline::reconstruct_rings_and_align(OM, MF, tri, base);
    auto CL = pipeline::make_centerline(OM, closed_mode, base);
    pipeline::compute_geom_and_save(base, CL.center, CL.spx, CL.spy, CL.s0, CL.L, closed_mode, RR.inner_from_mids, RR.outer_from_mids);
    std::cerr.rdbuf(old);
    auto innerE = closed_mode ? edges::ringEdges(RR.inner_from_mids) : edges::polylineEdges(RR.inner_from_mids);
    auto outerE = closed_mode ? edges::ringEdges(RR.outer_from_mids) : edges::polylineEdges(RR.outer_from_mids);

    // the loop of compute_geom_and_save (main.cpp:1306-1329), values kept in full precision
    const int Ncenter = (int)CL.center.size();
    const int Kmax = closed_mode ? C.samples : Ncenter;
    const int denomN = closed_mode ? C.samples : std::max(1, C.samples);
    std::vector<double> srel, xs, ys, hd, kp, din, dout, wid, vk;
    for (int k = 0; k < Kmax; ++k) {
        double si = CL.s0 + CL.L * (double(k) / double(denomN));
        double x, xp, xpp, y, yp, ypp;
        CL.spx.eval_with_deriv(si, x, xp, xpp);
        CL.spy.eval_with_deriv(si, y, yp, ypp);
        double heading = std::atan2(yp, xp);
        double speed2 = xp * xp + yp * yp;
        double denom = std::pow(std::max(1e-12, speed2), 1.5);
        double curv = (xp * ypp - yp * xpp) / denom;
        geom::Vec2 nvec = geom::normalize(geom::Vec2{-yp, xp}, 1e-12);
        double d_in = 0.0, d_out = 0.0;
        if (nvec.x != 0 || nvec.y != 0) distancesToRings({x, y}, nvec, innerE, outerE, d_in, d_out);
        double denom_k = std::max(std::fabs(curv), C.kappa_eps);
        double v_kappa = std::sqrt(C.a_lat_max / denom_k);
        if (v_kappa > C.v_cap_mps) v_kappa = C.v_cap_mps;
        srel.push_back(si - CL.s0); xs.push_back(x); ys.push_back(y); hd.push_back(heading); kp.push_back(curv);
        din.push_back(d_in); dout.push_back(d_out); wid.push_back(d_in + d_out); vk.push_back(v_kappa);
    }
    // pin against what the reference wrote itself
    {
        std::ifstream fi(base + "_with_geom.csv");
        std::string line;
        std::getline(fi, line);
        int k = 0;
        double worst = 0.0;
        while (std::getline(fi, line) && k < Kmax) {
            double v[9];
            if (std::sscanf(line.c_str(), "%lf,%lf,%lf,%lf,%lf,%lf,%lf,%lf,%lf", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5], &v[6], &v[7], &v[8]) != 9) break;
            const double mine[9] = {srel[k], xs[k], ys[k], hd[k], kp[k], din[k], dout[k], wid[k], vk[k]};
            for (int c = 0; c < 9; ++c) worst = std::max(worst, std::fabs(v[c] - mine[c]));
            ++k;
        }
        if (k != Kmax || worst > 1.0e-9) { std::fprintf(stderr, "geom: harness loop disagrees with the reference CSV (rows %d/%d, worst %.3e)\n", k, Kmax, worst); return 3; }
    }
    FILE* f = std::fopen(outBin.c_str(), "wb");
    if (!f) { std::perror("fopen"); return 1; }
    put_i64(f, 0x314D4752);  // "RGM1"
    put_i64(f, (int64_t)OM.ordered.size());
    put_i64(f, (int64_t)C.samples);
    put_i64(f, (int64_t)Kmax);
    put_i64(f, (int64_t)innerE.size());
    put_i64(f, (int64_t)outerE.size());
    put_i64(f, closed_mode ? 1 : 0);
    put_i64(f, C.emit_closed_duplicate ? 1 : 0);
    put_f64(f, CL.L);
    put_f64(f, CL.s0);
    put_pts(f, OM.ordered);
    for (auto& e : innerE) { put_f64(f, e.first.x); put_f64(f, e.first.y); put_f64(f, e.second.x); put_f64(f, e.second.y); }
    for (auto& e : outerE) { put_f64(f, e.first.x); put_f64(f, e.first.y); put_f64(f, e.second.x); put_f64(f, e.second.y); }
    put_pts(f, CL.center);                       // samples (+1 when the closing duplicate is emitted)
    put_vec(f, srel); put_vec(f, xs); put_vec(f, ys); put_vec(f, hd); put_vec(f, kp); put_vec(f, din); put_vec(f, dout); put_vec(f, wid); put_vec(f, vk);
    std::fclose(f);
    std::string rm = std::string("rm -rf ") + tmpl;
    if (std::system(rm.c_str()) != 0) {}
    std::printf("mids=%zu samples=%d rows=%d Min=%zu Mout=%zu L=%.9f s0=%.9f\n", OM.ordered.size(), C.samples, Kmax, innerE.size(), outerE.size(), CL.L, CL.s0);
    return 0;
}

// the debug comparison of pipeline::compute_mintime_and_save
int run_debug(int argc, char** argv)
{
    if (argc < 5) { std::fprintf(stderr, "usage: debug inner.csv outer.csv out.bin\n"); return 1; }
    const std::string innerPath = argv[2], outerPath = argv[3], outBin = argv[4];
    auto& C = cfg::get();
    C.verbose = false;
    C.debug_dump = true;
    char tmpl[] = "/tmp/ref_harness_XXXXXX";
    if (!mkdtemp(tmpl)) { std::perror("mkdtemp"); return 1; }
    const std::string base = std::string(tmpl) + "/centerline";
    std::ostringstream captured;
    std::streambuf* old = std::cerr.rdbuf(captured.rdbuf());
    auto inner = io::loadCSV_XY(innerPath);
    auto outer = io::loadCSV_XY(outerPath);
    const bool closed_mode = C.is_closed_track;
    auto tri = pipeline::buildDT(inner, outer);
    auto MF = pipeline::extract_mids_with_len_filter(tri, base);
    if (C.use_dynamic_samples) C.samples = pipeline::dynamic_samples_from_mids_count((int)MF.mids.size());
    auto OM = pipeline::order_and_align_mids_open_closed(MF.mids, closed_mode);
    auto RR = pipeline::reconstruct_rings_and_align(OM, MF, tri, base);
    auto CL = pipeline::make_centerline(OM, closed_mode, base);
    std::vector<geom::Vec2> center_for_opt = CL.center;
    if (closed_mode && center_for_opt.size() >= 2 && geom::almostEq(center_for_opt.front(), center_for_opt.back(), 1e-12))
        center_for_opt.pop_back();
    pipeline::compute_raceline_and_save(base, center_for_opt, CL.s0, CL.L, closed_mode, RR.inner_from_mids, RR.outer_from_mids);
    pipeline::compute_mintime_and_save(base, center_for_opt, CL.s0, CL.L, closed_mode, RR.inner_from_mids, RR.outer_from_mids);
    std::cerr.rdbuf(old);
    // the min-curv path as the debug block sees it (main.cpp:1454-1461) and the two extra laps (main.cpp:1464-1477)
    std::vector<geom::Vec2> mc = io::loadCSV_XY(base + "_raceline.csv");
    if (!mc.empty() && closed_mode && mc.size() >= 2 && geom::almostEq(mc.front(), mc.back(), 1e-12)) mc.pop_back();
    const double Hc = center_for_opt.size() > 0 ? (CL.L / std::max(1, (int)center_for_opt.size())) : 1.0;
    std::vector<double> hd, kp;
    raceline_min_curv::heading_curv_from_points_generic(center_for_opt, Hc, closed_mode, hd, kp);
    auto VPc = raceline_min_time::velocity_profile_forward_backward(kp, Hc, closed_mode);
    double Lmc = 0.0;
    for (size_t i = 0; i + 1 < mc.size(); ++i) Lmc += std::hypot(mc[i + 1].x - mc[i].x, mc[i + 1].y - mc[i].y);
    if (closed_mode && mc.size() >= 2) Lmc += std::hypot(mc[0].x - mc.back().x, mc[0].y - mc.back().y);
    const double Hmc = Lmc / std::max(1, (int)mc.size());
    raceline_min_curv::heading_curv_from_points_generic(mc, Hmc, closed_mode, hd, kp);
    auto VPmc = raceline_min_time::velocity_profile_forward_backward(kp, Hmc, closed_mode);
    // the CSV the reference wrote
    std::vector<double> cols;
    int rows = 0;
    {
        std::ifstream fi(base + "_debug_compare_paths.csv");
        std::string line;
        std::getline(fi, line);
        while (std::getline(fi, line)) {
            std::istringstream ls(line);
            std::string tok;
            int c = 0;
            while (std::getline(ls, tok, ',')) { cols.push_back(std::atof(tok.c_str())); ++c; }
            if (c != 20) { std::fprintf(stderr, "debug csv: %d columns\n", c); return 3; }
            ++rows;
        }
    }
    // the lap the min-time stage itself reported: "[mintime] Estimated laptime: x s" (main.cpp:1437) -- recompute exactly instead
    auto innerE = closed_mode ? edges::ringEdges(RR.inner_from_mids) : edges::polylineEdges(RR.inner_from_mids);
    auto outerE = closed_mode ? edges::ringEdges(RR.outer_from_mids) : edges::polylineEdges(RR.outer_from_mids);
    std::cerr.rdbuf(captured.rdbuf());
    auto mt = raceline_min_time::compute_min_time_raceline(center_for_opt, innerE, outerE, C.veh_width_m, CL.L, closed_mode);
    std::cerr.rdbuf(old);
    FILE* f = std::fopen(outBin.c_str(), "wb");
    if (!f) { std::perror("fopen"); return 1; }
    put_i64(f, 0x31474452);  // "RDG1"
    put_i64(f, (int64_t)center_for_opt.size());
    put_i64(f, (int64_t)mc.size());
    put_i64(f, (int64_t)rows);
    put_i64(f, (int64_t)innerE.size());
    put_i64(f, (int64_t)outerE.size());
    put_f64(f, CL.L); put_f64(f, CL.s0); put_f64(f, VPc.lap_time); put_f64(f, VPmc.lap_time); put_f64(f, mt.lap_time); put_f64(f, Lmc);
    put_pts(f, center_for_opt);
    put_pts(f, mc);
    for (auto& e : innerE) { put_f64(f, e.first.x); put_f64(f, e.first.y); put_f64(f, e.second.x); put_f64(f, e.second.y); }
    for (auto& e : outerE) { put_f64(f, e.first.x); put_f64(f, e.first.y); put_f64(f, e.second.x); put_f64(f, e.second.y); }
    put_vec(f, cols);
    std::fclose(f);
    std::string rm = std::string("rm -rf ") + tmpl;
    if (std::system(rm.c_str()) != 0) {}
    std::printf("N=%zu mc=%zu rows=%d lap_center=%.9f lap_mincurv=%.9f lap_mintime=%.9f\n", center_for_opt.size(), mc.size(), rows,
                VPc.lap_time, VPmc.lap_time, mt.lap_time);
    return 0;
}

template <class T> bool get_arr(FILE* f, std::vector<T>& v, size_t n) { v.resize(n); return n == 0 || fread(v.data(), sizeof(T), n, f) == n; }

// parameter row layout: the rl_params field order of include/raceline_b200.h (22 doubles, then 6 ints as doubles)
enum { P_VEH_ARG, P_VEH_M, P_SAFETY, P_LAMBDA, P_STEP_INIT, P_STEP_MIN, P_ARMIJO, P_KAPPA_EPS, P_VCAP, P_MASS, P_CD, P_AFRONT,
       P_RHO, P_CRR, P_PMAX, P_ATOT, P_ALAT, P_ACC_CAP, P_BRK_CAP, P_WTIME, P_GPOW, P_INVV_GAIN, P_MAX_OUTER, P_MAX_INNER,
       P_MAX_VPASS, P_USE_INVV, P_USE_TOT_GE_LAT, P_RESERVED, P_COUNT };

void apply_params(const double* p)
{
    auto& C = cfg::get();
    C.veh_width_m = p[P_VEH_M]; C.safety_margin_m = p[P_SAFETY]; C.lambda_smooth = p[P_LAMBDA];
    C.step_init = p[P_STEP_INIT]; C.step_min = p[P_STEP_MIN]; C.armijo_c = p[P_ARMIJO];
    C.kappa_eps = p[P_KAPPA_EPS]; C.v_cap_mps = p[P_VCAP]; C.mass_kg = p[P_MASS]; C.Cd = p[P_CD];
    C.A_front_m2 = p[P_AFRONT]; C.rho_air = p[P_RHO]; C.c_rr = p[P_CRR]; C.P_max_W = p[P_PMAX];
    C.a_total_max = p[P_ATOT]; C.a_lat_max = p[P_ALAT]; C.a_long_acc_cap = p[P_ACC_CAP];
    C.a_long_brake_cap = p[P_BRK_CAP]; C.w_time_gain = p[P_WTIME]; C.time_gamma_power = p[P_GPOW];
    C.inv_v_gain = p[P_INVV_GAIN]; C.max_outer_iters = (int)p[P_MAX_OUTER]; C.max_inner_iters = (int)p[P_MAX_INNER];
    C.max_vpass_iters = (int)p[P_MAX_VPASS]; C.time_weight_use_inv_v = p[P_USE_INVV] != 0.0;
    C.use_total_ge_lat = p[P_USE_TOT_GE_LAT] != 0.0;
}

int run_solve(int argc, char** argv)
{
    if (argc < 4) { std::fprintf(stderr, "usage: solve batch.bin out.bin [first_job] [n_jobs]\n"); return 1; }
    FILE* f = std::fopen(argv[2], "rb");
    if (!f) { std::perror("batch"); return 1; }
    int64_t hdr[4];
    if (fread(hdr, 8, 4, f) != 4 || hdr[0] != 0x31424C52) { std::fprintf(stderr, "bad batch file\n"); return 1; }
    const int64_t nt = hdr[1], np = hdr[2], nj = hdr[3];
    std::vector<int64_t> samp_off, seg_off, closed, jobs;
    std::vector<double> track_L, cxy, seg, params;
    bool ok = get_arr(f, samp_off, nt + 1) && get_arr(f, seg_off, 2 * nt + 1) && get_arr(f, track_L, nt) && get_arr(f, closed, nt);
    ok = ok && get_arr(f, cxy, 2 * (size_t)samp_off[nt]) && get_arr(f, seg, 4 * (size_t)seg_off[2 * nt]);
    ok = ok && get_arr(f, params, (size_t)np * P_COUNT) && get_arr(f, jobs, (size_t)nj * 3);
    std::fclose(f);
    if (!ok) { std::fprintf(stderr, "truncated batch file\n"); return 1; }
    int64_t first = (argc > 4) ? std::atoll(argv[4]) : 0;
    int64_t count = (argc > 5) ? std::atoll(argv[5]) : nj - first;
    if (first < 0 || first + count > nj) { std::fprintf(stderr, "job range out of bounds\n"); return 1; }

    auto& C = cfg::get();
    C.verbose = false; C.debug_dump = false;
    BtCountBuf nb; std::streambuf* old = std::cerr.rdbuf(&nb);

    FILE* o = std::fopen(argv[3], "wb");
    if (!o) { std::cerr.rdbuf(old); std::perror("out"); return 1; }
    put_i64(o, 0x32524C52);  // "RLR2": RLR1 + the backtracks the reference logged for the job (-1: it logged none)
    put_i64(o, count);
    double total_ms = 0;
    for (int64_t j = first; j < first + count; ++j) {
        const int64_t t = jobs[3 * j], pi = jobs[3 * j + 1], stage = jobs[3 * j + 2];
        const double* p = &params[(size_t)pi * P_COUNT];
        apply_params(p);
        const int64_t n = samp_off[t + 1] - samp_off[t];
        std::vector<geom::Vec2> center((size_t)n);
        for (int64_t i = 0; i < n; ++i) center[i] = {cxy[2 * (samp_off[t] + i)], cxy[2 * (samp_off[t] + i) + 1]};
        auto mk = [&](int64_t a, int64_t b) {
            std::vector<std::pair<geom::Vec2, geom::Vec2>> E((size_t)(b - a));
            for (int64_t k = a; k < b; ++k) E[k - a] = {{seg[4 * k], seg[4 * k + 1]}, {seg[4 * k + 2], seg[4 * k + 3]}};
            return E;
        };
        auto innerE = mk(seg_off[2 * t], seg_off[2 * t + 1]);
        auto outerE = mk(seg_off[2 * t + 1], seg_off[2 * t + 2]);
        const bool cl = closed[t] != 0;
        std::vector<geom::Vec2> rl; std::vector<double> hd, kp, at, al, v, ax; double lap = 0;
        nb.reset();
        auto t0 = std::chrono::steady_clock::now();
        if (stage == 1) {
            auto r = raceline_min_curv::compute_min_curvature_raceline(center, innerE, outerE, p[P_VEH_ARG], track_L[t], cl);
            rl = std::move(r.raceline); hd = std::move(r.heading); kp = std::move(r.curvature); at = std::move(r.alpha_total); al = std::move(r.alpha_last);
            v.assign((size_t)n, 0.0); ax.assign((size_t)n, 0.0);
        } else {
            auto r = raceline_min_time::compute_min_time_raceline(center, innerE, outerE, p[P_VEH_ARG], track_L[t], cl);
            rl = std::move(r.raceline); hd = std::move(r.heading); kp = std::move(r.curvature); at = std::move(r.alpha_total); al = std::move(r.alpha_last);
            v = std::move(r.v); ax = std::move(r.ax); lap = r.lap_time;
        }
        double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        total_ms += ms;
        put_i64(o, n); put_i64(o, stage); put_f64(o, ms); put_f64(o, lap); put_i64(o, nb.lines > 0 ? nb.bt : -1);
        put_pts(o, rl); put_vec(o, hd); put_vec(o, kp); put_vec(o, at); put_vec(o, al); put_vec(o, v); put_vec(o, ax);
    }
    std::fclose(o);
    std::cerr.rdbuf(old);
    std::printf("jobs=%lld solver_ms=%.3f\n", (long long)count, total_ms);
    return 0;
}

}  // namespace

int main(int argc, char** argv)
{
    if (argc >= 2 && std::strcmp(argv[1], "frontend") == 0) return run_frontend(argc, argv);
    if (argc >= 2 && std::strcmp(argv[1], "solve") == 0) return run_solve(argc, argv);
    if (argc >= 2 && std::strcmp(argv[1], "geom") == 0) return run_geom(argc, argv);
    if (argc >= 2 && std::strcmp(argv[1], "debug") == 0) return run_debug(argc, argv);
    std::fprintf(stderr, "usage: %s frontend|solve|geom|debug ...\n", argv[0]);
    return 1;
}
