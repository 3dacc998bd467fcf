// raceline_b200_adapter.hpp -- header-only C++17 host adapter over the C ABI (raceline_b200.h).
//
// Mirrors the two solver entry points of tjsdn3065/Practice_path_planning_for_formula_student_driverless
// with the SAME argument lists, so the reference pipeline swaps exactly two calls and keeps everything else
// (cfg::Config, CSV loaders/dumpers, Delaunay/MST/spline/width stages):
//
//   raceline_min_curv::compute_min_curvature_raceline(center, innerE, outerE, veh_width, L, closed)   main.cpp:683-686
//   raceline_min_time::compute_min_time_raceline(center, innerE, outerE, veh_width, L, closed)        main.cpp:905-909
//
// The templates only rely on names the reference already has: a Vec2 of two doubles (main.cpp:125),
// vector<pair<Vec2,Vec2>> edge lists (main.cpp:251), the Result structs (main.cpp:677-681, 897-903) and the
// Config field names (main.cpp:77-113).  Errors surface as std::runtime_error, like the reference's own
// stage wrappers (main.cpp:1353).  There is no CPU fallback.
#pragma once

#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "raceline_b200.h"

namespace raceline_b200 {

// cfg::Config (main.cpp:47-119) -> rl_params.  veh_width is the solver argument (main.cpp:706, 930).
template <class Cfg>
inline rl_params params_from_config(const Cfg& C, double veh_width)
{
    rl_params p;
    rl_default_params(&p);
    p.veh_width_arg = veh_width;
    p.veh_width_m = C.veh_width_m;
    p.safety_margin_m = C.safety_margin_m;
    p.lambda_smooth = C.lambda_smooth;
    p.step_init = C.step_init;
    p.step_min = C.step_min;
    p.armijo_c = C.armijo_c;
    p.kappa_eps = C.kappa_eps;
    p.v_cap_mps = C.v_cap_mps;
    p.mass_kg = C.mass_kg;
    p.Cd = C.Cd;
    p.A_front_m2 = C.A_front_m2;
    p.rho_air = C.rho_air;
    p.c_rr = C.c_rr;
    p.P_max_W = C.P_max_W;
    p.a_total_max = C.a_total_max;
    p.a_lat_max = C.a_lat_max;
    p.a_long_acc_cap = C.a_long_acc_cap;
    p.a_long_brake_cap = C.a_long_brake_cap;
    p.w_time_gain = C.w_time_gain;
    p.time_gamma_power = C.time_gamma_power;
    p.inv_v_gain = C.inv_v_gain;
    p.max_outer_iters = C.max_outer_iters;
    p.max_inner_iters = C.max_inner_iters;
    p.max_vpass_iters = C.max_vpass_iters;
    p.time_weight_use_inv_v = C.time_weight_use_inv_v ? 1 : 0;
    p.use_total_ge_lat = C.use_total_ge_lat ? 1 : 0;
    return p;
}

// One context per process, created on first use (replaces the global cfg::get() coupling, main.cpp:120).  The
// function-local static is initialised exactly once even when several threads arrive together (C++11), and the C ABI
// serialises the calls threads make through the shared context (see the threading contract in raceline_b200.h).
inline rl_ctx* context(int device = 0)
{
    struct Holder {
        rl_ctx* h = nullptr;
        explicit Holder(int dev)
        {
            int st = RL_OK;
            h = rl_create(dev, &st);
            if (!h) throw std::runtime_error(std::string("raceline_b200: rl_create failed: ") + rl_status_string(st));
        }
        ~Holder() { if (h) rl_destroy(h); }
    };
    static Holder holder(device);   // a throwing constructor leaves it uninitialised: the next call tries again
    return holder.h;
}

inline void check(int st, const char* what)
{
    if (st != RL_OK)
        throw std::runtime_error(std::string("raceline_b200: ") + what + ": " + rl_status_string(st) + " (" + rl_last_error(context()) + ")");
}

// Result = raceline_min_curv::Result (fields raceline, heading, curvature, alpha_total, alpha_last)
template <class Result, class Vec2, class Cfg>
inline Result compute_min_curvature_raceline(const std::vector<Vec2>& center,
                                             const std::vector<std::pair<Vec2, Vec2>>& innerE,
                                             const std::vector<std::pair<Vec2, Vec2>>& outerE,
                                             double veh_width, double L, bool closed, const Cfg& C)
{
    static_assert(sizeof(Vec2) == 2 * sizeof(double), "Vec2 must be two doubles");
    static_assert(sizeof(std::pair<Vec2, Vec2>) == 4 * sizeof(double), "edge = 4 doubles");
    const int n = (int)center.size();
    Result r;
    if (n == 0) return r;   // main.cpp:689
    r.raceline.resize(n); r.heading.resize(n); r.curvature.resize(n); r.alpha_total.resize(n); r.alpha_last.resize(n);
    const rl_params p = params_from_config(C, veh_width);
    check(rl_compute_min_curvature_raceline(context(), reinterpret_cast<const double*>(center.data()), n,
                                            reinterpret_cast<const double*>(innerE.data()), (int)innerE.size(),
                                            reinterpret_cast<const double*>(outerE.data()), (int)outerE.size(),
                                            veh_width, L, closed ? 1 : 0, &p, reinterpret_cast<double*>(r.raceline.data()),
                                            r.heading.data(), r.curvature.data(), r.alpha_total.data(), r.alpha_last.data(),
                                            nullptr),
          "compute_min_curvature_raceline");
    return r;
}

// Result = raceline_min_time::Result (adds v, ax, lap_time)
template <class Result, class Vec2, class Cfg>
inline Result compute_min_time_raceline(const std::vector<Vec2>& center,
                                        const std::vector<std::pair<Vec2, Vec2>>& innerE,
                                        const std::vector<std::pair<Vec2, Vec2>>& outerE,
                                        double veh_width, double L, bool closed, const Cfg& C)
{
    static_assert(sizeof(Vec2) == 2 * sizeof(double), "Vec2 must be two doubles");
    static_assert(sizeof(std::pair<Vec2, Vec2>) == 4 * sizeof(double), "edge = 4 doubles");
    const int n = (int)center.size();
    Result r;
    if (n == 0) return r;   // main.cpp:912
    r.raceline.resize(n); r.heading.resize(n); r.curvature.resize(n); r.alpha_total.resize(n); r.alpha_last.resize(n);
    r.v.resize(n); r.ax.resize(n);
    const rl_params p = params_from_config(C, veh_width);
    double lap = 0.0;
    check(rl_compute_min_time_raceline(context(), reinterpret_cast<const double*>(center.data()), n,
                                       reinterpret_cast<const double*>(innerE.data()), (int)innerE.size(),
                                       reinterpret_cast<const double*>(outerE.data()), (int)outerE.size(),
                                       veh_width, L, closed ? 1 : 0, &p, reinterpret_cast<double*>(r.raceline.data()),
                                       r.heading.data(), r.curvature.data(), r.alpha_total.data(), r.alpha_last.data(),
                                       r.v.data(), r.ax.data(), &lap, nullptr),
          "compute_min_time_raceline");
    r.lap_time = lap;
    return r;
}

}  // namespace raceline_b200
