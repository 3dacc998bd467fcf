/*
 * raceline_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, single-threaded CPU restatement of the reference's min-curvature
 * and min-time raceline stages (tjsdn3065/Practice_path_planning_for_formula_
 * student_driverless, src/main.cpp).  It exists only so that tests/, bench.py's
 * cpu_baseline leg and __graft_entry__.smoke() can check the CUDA path; nothing
 * under practice_path_planning_for_formula_student_driverless_b200/ may call it.
 *
 * Parity status: PINNED.  The reference ships no golden vectors,
 * so the pin is the reference itself: oracle/ref_harness.cpp #includes the
 * unmodified reference TU, and tests/golden/make_golden.py checks that this
 * file reproduces the reference's outputs bit for bit on all 7 shipped maps
 * (same operation order, no FMA contraction: build with -ffp-contract=off).
 *
 * Every function cites the reference lines it restates.  Floating-point
 * expressions keep the reference's association order on purpose.
 */
#include "raceline_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* std::min / std::max / std::clamp have defined NaN and tie behaviour; keep it. */
static inline double mn(double a, double b) { return (b < a) ? b : a; }
static inline double mx(double a, double b) { return (a < b) ? b : a; }
static inline double clampd(double v, double lo, double hi) { return (v < lo) ? lo : ((hi < v) ? hi : v); }

/* ------------------------------------------------------------------------
 * Ray / segment utilities                                 main.cpp:478-512
 * seg: m segments, 4 doubles each (x0,y0,x1,y1) = vector<pair<Vec2,Vec2>>
 * ---------------------------------------------------------------------- */

/* main.cpp:478-490 */
static int ray_hits_segment(double ax_, double ay_, double dx, double dy, const double* s, double* t_out)
{
    const double eps = 1e-15;
    double vx = s[2] - s[0], vy = s[3] - s[1];
    double den = dx * (-vy) + dy * (vx);
    if (fabs(den) < eps) return 0;
    double ax = s[0] - ax_, ay = s[1] - ay_;
    double inv = 1.0 / den;
    double t = (ax * (-vy) + ay * (vx)) * inv;
    double u = (dx * ay - dy * ax) * inv;
    if (t >= 0.0 && u >= -1e-12 && u <= 1.0 + 1e-12) { *t_out = t; return 1; }
    return 0;
}

/* main.cpp:491-500 */
double orc_ray_to_ring(double px, double py, double dx, double dy, const double* seg, int m)
{
    double best = INFINITY;
    for (int j = 0; j < m; ++j) {
        double t;
        if (ray_hits_segment(px, py, dx, dy, seg + 4 * (size_t)j, &t))
            if (t > 0.0 && t < best) best = t;
    }
    return best;
}

/* main.cpp:501-512 */
double orc_min_dist_to_segments(double px, double py, const double* seg, int m)
{
    double best = INFINITY;
    for (int j = 0; j < m; ++j) {
        const double* s = seg + 4 * (size_t)j;
        double abx = s[2] - s[0], aby = s[3] - s[1];
        double apx = px - s[0], apy = py - s[1];
        double denom = mx(1e-30, abx * abx + aby * aby);
        double t = clampd((abx * apx + aby * apy) / denom, 0.0, 1.0);
        double qx = s[0] + abx * t, qy = s[1] + aby * t;
        best = mn(best, hypot(px - qx, py - qy));
    }
    return best;
}

/* the safe_ray lambda, main.cpp:694-699 / 917-922 */
static double safe_ray(double px, double py, double dx, double dy, const double* seg, int m)
{
    double t = orc_ray_to_ring(px, py, dx, dy, seg, m);
    if (!isfinite(t)) t = orc_min_dist_to_segments(px, py, seg, m);
    if (!isfinite(t)) t = 0.0;
    return mx(0.0, t);
}

/* corridor build, main.cpp:701-711 and 749-756 (min-time: 925-935, 1033-1040) */
void orc_corridor(const double* pxy, const double* nxy, int n,
                  const double* inner_seg, int m_in, const double* outer_seg, int m_out,
                  double guard, double* lo, double* hi)
{
    for (int i = 0; i < n; ++i) {
        double nx = nxy[2 * i], ny = nxy[2 * i + 1], px = pxy[2 * i], py = pxy[2 * i + 1];
        double dpos = mn(safe_ray(px, py, nx, ny, inner_seg, m_in), safe_ray(px, py, nx, ny, outer_seg, m_out));
        double dneg = mn(safe_ray(px, py, -nx, -ny, inner_seg, m_in), safe_ray(px, py, -nx, -ny, outer_seg, m_out));
        hi[i] = mx(0.0, dpos - guard);
        lo[i] = -mx(0.0, dneg - guard);
        if (!isfinite(hi[i])) hi[i] = 0.0;
        if (!isfinite(lo[i])) lo[i] = 0.0;
    }
}

/* distancesToRings, main.cpp:513-524 (used by the width/geometry stage, "next" row 1) */
void orc_distances_to_rings(double px, double py, double nx, double ny,
                            const double* inner_seg, int m_in, const double* outer_seg, int m_out,
                            double* d_inner, double* d_outer)
{
    double di1 = orc_ray_to_ring(px, py, nx, ny, inner_seg, m_in);
    double di2 = orc_ray_to_ring(px, py, -nx, -ny, inner_seg, m_in);
    *d_inner = (isfinite(di1) || isfinite(di2)) ? mn(di1, di2) : orc_min_dist_to_segments(px, py, inner_seg, m_in);
    double do1 = orc_ray_to_ring(px, py, nx, ny, outer_seg, m_out);
    double do2 = orc_ray_to_ring(px, py, -nx, -ny, outer_seg, m_out);
    *d_outer = (isfinite(do1) || isfinite(do2)) ? mn(do1, do2) : orc_min_dist_to_segments(px, py, outer_seg, m_out);
    if (!isfinite(*d_inner)) *d_inner = 0.0;
    if (!isfinite(*d_outer)) *d_outer = 0.0;
}

/* ------------------------------------------------------------------------
 * Finite-difference operators                              main.cpp:545-579
 * ---------------------------------------------------------------------- */
static inline int wrapi(int i, int n) { i %= n; if (i < 0) i += n; return i; }

/* DiffOps::D1 549-551 / DiffOpsOpen::D1 563-566 */
static void d1(const double* a, double* out, int n, double h, int closed)
{
    if (closed) {
        double inv2h = 1.0 / (2 * h);
        for (int i = 0; i < n; ++i) out[i] = (a[wrapi(i + 1, n)] - a[wrapi(i - 1, n)]) * inv2h;
    } else {
        double invh = 1.0 / h, inv2h = 1.0 / (2 * h);
        for (int i = 0; i < n; ++i) out[i] = 0.0;
        if (n == 0) return;
        if (n == 1) { out[0] = 0; return; }
        out[0] = (a[1] - a[0]) * invh;
        for (int i = 1; i <= n - 2; ++i) out[i] = (a[i + 1] - a[i - 1]) * inv2h;
        out[n - 1] = (a[n - 1] - a[n - 2]) * invh;
    }
}
/* DiffOps::D2 552-554 / DiffOpsOpen::D2 573-575 */
static void d2(const double* a, double* out, int n, double h, int closed)
{
    double invh2 = 1.0 / (h * h);
    if (closed) {
        for (int i = 0; i < n; ++i) out[i] = (a[wrapi(i + 1, n)] - 2 * a[i] + a[wrapi(i - 1, n)]) * invh2;
    } else {
        for (int i = 0; i < n; ++i) out[i] = 0.0;
        if (n <= 2) return;
        for (int i = 1; i <= n - 2; ++i) out[i] = (a[i + 1] - 2 * a[i] + a[i - 1]) * invh2;
    }
}
/* DiffOps::D1T 555-557 / DiffOpsOpen::D1T 567-572 (scatter form, same += order) */
static void d1t(const double* v, double* out, int n, double h, int closed)
{
    if (closed) {
        double inv2h = 1.0 / (2 * h);
        for (int i = 0; i < n; ++i) out[i] = (v[wrapi(i - 1, n)] - v[wrapi(i + 1, n)]) * inv2h;
    } else {
        double invh = 1.0 / h, inv2h = 1.0 / (2 * h);
        for (int i = 0; i < n; ++i) out[i] = 0.0;
        if (n <= 1) return;
        out[0] += (-invh) * v[0];
        out[1] += (+invh) * v[0];
        for (int i = 1; i <= n - 2; ++i) { out[i - 1] += (-inv2h) * v[i]; out[i + 1] += (+inv2h) * v[i]; }
        out[n - 2] += (-invh) * v[n - 1];
        out[n - 1] += (+invh) * v[n - 1];
    }
}
/* DiffOps::D2T 558 (= D2) / DiffOpsOpen::D2T 576-578 */
static void d2t(const double* v, double* out, int n, double h, int closed)
{
    if (closed) { d2(v, out, n, h, 1); return; }
    double invh2 = 1.0 / (h * h);
    for (int i = 0; i < n; ++i) out[i] = 0.0;
    if (n <= 2) return;
    for (int i = 1; i <= n - 2; ++i) {
        out[i - 1] += (+invh2) * v[i];
        out[i] += (-2 * invh2) * v[i];
        out[i + 1] += (+invh2) * v[i];
    }
}

/* ------------------------------------------------------------------------
 * Per-sample geometry
 * ---------------------------------------------------------------------- */

/* normals_from_points_generic, main.cpp:581-593; geom::normalize 132 */
void orc_normals(const double* pxy, int n, int closed, double* nxy)
{
    for (int i = 0; i < n; ++i) {
        double tx, ty;
        if (n == 1) { tx = 1; ty = 0; }
        else if (closed) {
            int ip = (i + 1) % n, im = (i - 1 + n) % n;
            tx = (pxy[2 * ip] - pxy[2 * im]) * 0.5; ty = (pxy[2 * ip + 1] - pxy[2 * im + 1]) * 0.5;
        } else if (i == 0) { tx = pxy[2] - pxy[0]; ty = pxy[3] - pxy[1]; }
        else if (i == n - 1) { tx = pxy[2 * (n - 1)] - pxy[2 * (n - 2)]; ty = pxy[2 * (n - 1) + 1] - pxy[2 * (n - 2) + 1]; }
        else { tx = (pxy[2 * (i + 1)] - pxy[2 * (i - 1)]) * 0.5; ty = (pxy[2 * (i + 1) + 1] - pxy[2 * (i - 1) + 1]) * 0.5; }
        if (sqrt(tx * tx + ty * ty) < 1e-15) { tx = 1; ty = 0; }
        double nvx = -ty, nvy = tx;
        double len = sqrt(nvx * nvx + nvy * nvy);
        if (len < 1e-15) { nxy[2 * i] = 0; nxy[2 * i + 1] = 0; }
        else { nxy[2 * i] = nvx / len; nxy[2 * i + 1] = nvy / len; }
    }
}

/* the `deriv` lambda shared by main.cpp:599-613 and 625-639 */
static void derivs(const double* p, int n, int i, double h, int closed,
                   double* xp, double* yp, double* xpp, double* ypp)
{
#define PX(k) p[2 * (k)]
#define PY(k) p[2 * (k) + 1]
    if (n == 1) { *xp = 1; *yp = 0; *xpp = *ypp = 0; return; }
    if (closed) {
        int ip = (i + 1) % n, im = (i - 1 + n) % n;
        *xp = (PX(ip) - PX(im)) / (2 * h); *yp = (PY(ip) - PY(im)) / (2 * h);
        *xpp = (PX(ip) - 2 * PX(i) + PX(im)) / (h * h); *ypp = (PY(ip) - 2 * PY(i) + PY(im)) / (h * h);
    } else if (i == 0) {
        *xp = (PX(1) - PX(0)) / h; *yp = (PY(1) - PY(0)) / h;
        if (n >= 3) { *xpp = (PX(2) - 2 * PX(1) + PX(0)) / (h * h); *ypp = (PY(2) - 2 * PY(1) + PY(0)) / (h * h); }
        else *xpp = *ypp = 0;
    } else if (i == n - 1) {
        *xp = (PX(n - 1) - PX(n - 2)) / h; *yp = (PY(n - 1) - PY(n - 2)) / h;
        if (n >= 3) { *xpp = (PX(n - 1) - 2 * PX(n - 2) + PX(n - 3)) / (h * h); *ypp = (PY(n - 1) - 2 * PY(n - 2) + PY(n - 3)) / (h * h); }
        else *xpp = *ypp = 0;
    } else {
        *xp = (PX(i + 1) - PX(i - 1)) / (2 * h); *yp = (PY(i + 1) - PY(i - 1)) / (2 * h);
        *xpp = (PX(i + 1) - 2 * PX(i) + PX(i - 1)) / (h * h); *ypp = (PY(i + 1) - 2 * PY(i) + PY(i - 1)) / (h * h);
    }
#undef PX
#undef PY
}

/* heading_curv_from_points_generic, main.cpp:595-620 */
void orc_heading_curv(const double* pxy, int n, double h, int closed, double* heading, double* kappa)
{
    for (int i = 0; i < n; ++i) {
        double xp, yp, xpp, ypp;
        derivs(pxy, n, i, h, closed, &xp, &yp, &xpp, &ypp);
        heading[i] = atan2(yp, xp);
        double denom = pow(mx(1e-12, xp * xp + yp * yp), 1.5);
        kappa[i] = (xp * ypp - yp * xpp) / denom;
    }
}

/* precompute_lin_geom_generic, main.cpp:622-651 */
void orc_lin_geom(const double* pxy, const double* nxy, int n, double h, int closed,
                  double* A1, double* A2, double* N0, double* W)
{
    for (int i = 0; i < n; ++i) {
        double xp, yp, xpp, ypp;
        derivs(pxy, n, i, h, closed, &xp, &yp, &xpp, &ypp);
        double nx = nxy[2 * i], ny = nxy[2 * i + 1];
        A1[i] = nx * ypp - ny * xpp;
        A2[i] = xp * ny - yp * nx;
        N0[i] = xp * ypp - yp * xpp;
        double denom = pow(mx(1e-12, xp * xp + yp * yp), 1.5);
        W[i] = 1.0 / denom;
    }
}

/* ------------------------------------------------------------------------
 * Cost / gradient of the frozen (linearised) problem
 * eval_cost_grad_frozen        main.cpp:654-675   (gamma2 == NULL)
 * eval_cost_grad_timeweighted  main.cpp:866-895   (gamma2 != NULL)
 * work: 8*n doubles of scratch
 * ---------------------------------------------------------------------- */
double orc_eval_cost_grad(const double* A1, const double* A2, const double* N0, const double* W,
                          const double* gamma2, double h, double lambda_smooth, const double* alpha,
                          int n, int closed, double* grad, double* work)
{
    double *a1 = work, *a2 = work + n, *z = work + 2 * n, *q1 = work + 3 * n, *q2 = work + 4 * n;
    double *g1 = work + 5 * n, *g2 = work + 6 * n, *gsm = work + 7 * n;
    d1(alpha, a1, n, h, closed);
    d2(alpha, a2, n, h, closed);
    for (int i = 0; i < n; ++i) z[i] = W[i] * (N0[i] + A1[i] * a1[i] + A2[i] * a2[i]);
    double J = 0;
    if (gamma2) { for (int i = 0; i < n; ++i) J += gamma2[i] * z[i] * z[i]; }
    else        { for (int i = 0; i < n; ++i) J += z[i] * z[i]; }
    double Jsm = 0;
    for (int i = 0; i < n; ++i) Jsm += a1[i] * a1[i];
    J += lambda_smooth * Jsm;
    for (int i = 0; i < n; ++i) {
        double wz = gamma2 ? (W[i] * gamma2[i] * z[i]) : (W[i] * z[i]);
        q1[i] = A1[i] * wz; q2[i] = A2[i] * wz;
    }
    d1t(q1, g1, n, h, closed);
    d2t(q2, g2, n, h, closed);
    /* the reference recomputes D1*alpha here (main.cpp:670); a1 already holds it */
    d1t(a1, gsm, n, h, closed);
    for (int i = 0; i < n; ++i) grad[i] = 2.0 * (g1[i] + g2[i]) + 2.0 * lambda_smooth * gsm[i];
    return J;
}

/* ------------------------------------------------------------------------
 * v(s) profile                                             main.cpp:782-862
 * ---------------------------------------------------------------------- */

/* the ax_max_at lambda, main.cpp:797-824 */
void orc_ax_max_at(const rl_params* C, double vi, double ki, double* a_acc_out, double* a_brk_out)
{
    double alat = vi * vi * fabs(ki);
    double a_total = C->use_total_ge_lat ? mx(C->a_total_max, C->a_lat_max) : C->a_total_max;
    double a_res = sqrt(mx(0.0, a_total * a_total - alat * alat));
    double Fd = 0.5 * C->rho_air * C->Cd * C->A_front_m2 * vi * vi;
    double Fr = C->mass_kg * 9.81 * C->c_rr;
    double a_power = (C->P_max_W > 0 && vi > 1e-6)
                         ? (C->P_max_W / (C->mass_kg * vi) - (Fd + Fr) / C->mass_kg)
                         : 1e9;
    double a_acc = a_res;                         /* std::min({a_res, cap, a_power}) */
    if (C->a_long_acc_cap < a_acc) a_acc = C->a_long_acc_cap;
    if (a_power < a_acc) a_acc = a_power;
    a_acc = mx(0.0, a_acc);
    double a_brk = mn(a_res, C->a_long_brake_cap) + (Fd + Fr) / C->mass_kg;
    a_brk = mx(0.0, a_brk);
    *a_acc_out = a_acc; *a_brk_out = a_brk;
}

/* velocity_profile_forward_backward, main.cpp:782-862.  Returns lap time. */
double orc_velocity_profile(const rl_params* C, const double* kappa, int n, double h, int closed,
                            double* v, double* ax)
{
    if (n == 0) return 0.0;
    for (int i = 0; i < n; ++i) {
        double k = fabs(kappa[i]);
        double v_kappa = sqrt(C->a_lat_max / mx(k, C->kappa_eps));
        v[i] = mn(C->v_cap_mps, v_kappa);
    }
    int iters = C->max_vpass_iters;
    while (iters-- > 0) {
        double a_acc, a_brk;
        for (int i = 0; i + 1 < n; ++i) {
            orc_ax_max_at(C, v[i], kappa[i], &a_acc, &a_brk);
            double vf = sqrt(mx(0.0, v[i] * v[i] + 2.0 * a_acc * h));
            v[i + 1] = mn(v[i + 1], vf);
        }
        if (closed) {
            orc_ax_max_at(C, v[n - 1], kappa[n - 1], &a_acc, &a_brk);
            double vf0 = sqrt(mx(0.0, v[n - 1] * v[n - 1] + 2.0 * a_acc * h));
            v[0] = mn(v[0], vf0);
        }
        for (int i = n - 2; i >= 0; --i) {
            orc_ax_max_at(C, v[i + 1], kappa[i + 1], &a_acc, &a_brk);
            double vb = sqrt(mx(0.0, v[i + 1] * v[i + 1] + 2.0 * a_brk * h));
            v[i] = mn(v[i], vb);
        }
        if (closed) {
            orc_ax_max_at(C, v[0], kappa[0], &a_acc, &a_brk);
            double vbN = sqrt(mx(0.0, v[0] * v[0] + 2.0 * a_brk * h));
            v[n - 1] = mn(v[n - 1], vbN);
        }
    }
    double t = 0.0;
    for (int i = 0; i < n; ++i) {
        int j = (i + 1 < n) ? i + 1 : (closed ? 0 : i);
        double v0 = v[i], v1 = v[j];
        ax[i] = (v1 * v1 - v0 * v0) / (2.0 * h);
        t += h / mx(1e-6, v[i]);
    }
    return t;
}

/* gamma^2 time weights, main.cpp:950-977 */
void orc_time_weights(const rl_params* C, const double* kappa, const double* v, int n, double* gamma2)
{
    double v_avg = 0.0;
    for (int i = 0; i < n; ++i) v_avg += v[i];
    v_avg /= (double)((n > 1) ? n : 1);
    for (int i = 0; i < n; ++i) {
        double k = fabs(kappa[i]);
        double vkappa = sqrt(C->a_lat_max / mx(k, C->kappa_eps));
        double r = pow(mn(1.0, v[i] / mx(1e-6, vkappa)), 2.0);
        r = mn(1.0, mx(0.0, r));
        double corner_w = 1.0 + C->w_time_gain * pow(r, C->time_gamma_power);
        double invv_w = 1.0;
        if (C->time_weight_use_inv_v) {
            double ratio = v_avg / mx(1e-6, v[i]);
            invv_w = 1.0 + C->inv_v_gain * (ratio - 1.0);
            if (invv_w < 1.0) invv_w = 1.0;
            if (invv_w > 3.0) invv_w = 3.0;
        }
        double gamma = corner_w * invv_w;
        gamma2[i] = gamma * gamma;
    }
}

/* ------------------------------------------------------------------------
 * Stage drivers
 * compute_min_curvature_raceline  main.cpp:683-764
 * compute_min_time_raceline       main.cpp:905-1052
 * ---------------------------------------------------------------------- */
int orc_solve(int stage, const double* center_xy, int n,
              const double* inner_seg, int m_in, const double* outer_seg, int m_out,
              double L, int closed, const rl_params* C,
              double* out_xy, double* out_heading, double* out_curv,
              double* out_alpha_total, double* out_alpha_last,
              double* out_v, double* out_ax, rl_job_stats* st)
{
    if (st) { memset(st, 0, sizeof(*st)); st->n = n; }
    if (stage != RL_STAGE_MINCURV && stage != RL_STAGE_MINTIME && stage != RL_STAGE_EVAL) return RL_ERR_ARG;
    if (n == 0) return RL_OK;                                  /* main.cpp:689 / 912 */
    if (stage == RL_STAGE_EVAL) {
        /* the debug block's extra laps, main.cpp:1464-1477: heading/curvature of the given points with h = L/N, then the v(s) profile */
        const double hh = L / (double)n;
        double* tmp = (double*)malloc(sizeof(double) * 2 * (size_t)n);
        if (!tmp) return RL_ERR_NOMEM;
        for (int i = 0; i < n; ++i) { out_alpha_total[i] = 0.0; out_alpha_last[i] = 0.0; }
        memcpy(out_xy, center_xy, sizeof(double) * 2 * (size_t)n);
        orc_heading_curv(center_xy, n, hh, closed, out_heading, out_curv);
        double lap = orc_velocity_profile(C, out_curv, n, hh, closed, out_v ? out_v : tmp, out_ax ? out_ax : tmp + n);
        if (st) st->lap_time = lap;
        free(tmp);
        return RL_OK;
    }
    const int mt = (stage == RL_STAGE_MINTIME);
    const double h = L / (double)n;                            /* main.cpp:690 / 913 */

    size_t nn = (size_t)n;
    double* buf = (double*)malloc(sizeof(double) * nn * 26);
    if (!buf) return RL_ERR_NOMEM;
    double *P = buf, *nv = buf + 2 * nn, *lo = buf + 4 * nn, *hi = buf + 5 * nn;
    double *alpha = buf + 6 * nn, *a_new = buf + 7 * nn, *grad = buf + 8 * nn, *grad_new = buf + 9 * nn;
    double *A1 = buf + 10 * nn, *A2 = buf + 11 * nn, *N0 = buf + 12 * nn, *W = buf + 13 * nn;
    double *gamma2 = buf + 14 * nn, *kap = buf + 15 * nn, *hd = buf + 16 * nn, *work = buf + 17 * nn; /* 8n */
    double* vax = buf + 25 * nn; /* n: scratch ax when the caller passes none */
    double *vv = out_v ? out_v : a_new; /* v is only needed between profile and weights */

    memcpy(P, center_xy, sizeof(double) * 2 * nn);
    orc_normals(P, n, closed, nv);                             /* 692 / 915 */
    orc_corridor(P, nv, n, inner_seg, m_in, outer_seg, m_out,
                 C->veh_width_arg * 0.5 + C->safety_margin_m, lo, hi); /* 701-711 / 925-935 */

    for (int i = 0; i < n; ++i) { alpha[i] = 0.0; out_alpha_total[i] = 0.0; out_alpha_last[i] = 0.0; }

    for (int outer = 0; outer < C->max_outer_iters; ++outer) {  /* 721 / 939 */
        orc_lin_geom(P, nv, n, h, closed, A1, A2, N0, W);       /* 722 / 941 */
        const double* g2 = NULL;
        double lap_t = 0.0;
        if (mt) {
            orc_heading_curv(P, n, h, closed, hd, kap);          /* 944 */
            double* vtmp = work;                                 /* v, ax scratch: work[0..2n) */
            lap_t = orc_velocity_profile(C, kap, n, h, closed, vtmp, vtmp + nn); /* 947 */
            orc_time_weights(C, kap, vtmp, n, gamma2);           /* 950-977 */
            g2 = gamma2;
        }
        double step = C->step_init;                              /* 723 / 996 */
        double J = orc_eval_cost_grad(A1, A2, N0, W, g2, h, C->lambda_smooth, alpha, n, closed, grad, work);
        double J_prev = J;
        int acc_n = 0, bt_n = 0;
        if (st) { st->evals++; if (outer < RL_MAX_OUTER_LOG) { st->J0[outer] = J; st->lap_outer[outer] = lap_t; } }

        for (int it = 0; it < C->max_inner_iters; ++it) {       /* 727 / 1000 */
            int accepted = 0, bt = 0;
            while (bt < 20) {                                   /* 729 */
                for (int i = 0; i < n; ++i) {
                    double ai = alpha[i] - step * grad[i];
                    a_new[i] = mn(hi[i], mx(lo[i], ai));        /* 731 */
                }
                double J_new = orc_eval_cost_grad(A1, A2, N0, W, g2, h, C->lambda_smooth, a_new, n, closed, grad_new, work);
                if (st) st->evals++;
                double dec = 0.0;
                for (int i = 0; i < n; ++i) dec += grad[i] * (a_new[i] - alpha[i]); /* 733 */
                if (J_new <= J + C->armijo_c * dec) {           /* 734 */
                    double* t1 = alpha; alpha = a_new; a_new = t1;
                    double* t2 = grad; grad = grad_new; grad_new = t2;
                    J = J_new; accepted = 1; break;
                }
                step *= 0.5; bt++; bt_n++;                       /* 737 */
                if (step < C->step_min) break;
            }
            if (!accepted) break;                                /* 739 */
            acc_n++;
            if (fabs(J_prev - J) < 1e-10) break;                 /* 740 */
            J_prev = J;
        }
        if (st) {
            st->accepted += acc_n; st->backtracks += bt_n; st->outer_done = outer + 1;
            if (outer < RL_MAX_OUTER_LOG) { st->Jend[outer] = J; st->acc_outer[outer] = acc_n; st->bt_outer[outer] = bt_n; }
        }
        for (int i = 0; i < n; ++i) {                            /* 743-745 / 1027-1030 */
            out_alpha_last[i] = alpha[i];
            P[2 * i] += nv[2 * i] * alpha[i]; P[2 * i + 1] += nv[2 * i + 1] * alpha[i];
            out_alpha_total[i] += alpha[i];
        }
        orc_normals(P, n, closed, nv);                           /* 746 / 1031 */
        orc_corridor(P, nv, n, inner_seg, m_in, outer_seg, m_out,
                     C->veh_width_m * 0.5 + C->safety_margin_m, lo, hi); /* 749-756 / 1033-1040 */
        for (int i = 0; i < n; ++i) alpha[i] = 0.0;              /* 757 / 1041 */
    }

    orc_heading_curv(P, n, h, closed, out_heading, out_curv);    /* 761 / 1046 */
    memcpy(out_xy, P, sizeof(double) * 2 * nn);
    if (mt) {
        double lap = orc_velocity_profile(C, out_curv, n, h, closed, vv, out_ax ? out_ax : vax); /* 1047 */
        if (st) st->lap_time = lap;
    }
    free(buf);
    return RL_OK;
}

/* cfg::Config defaults, main.cpp:77-113 (values restated, not shared with the product library) */
void orc_default_params(rl_params* p)
{
    memset(p, 0, sizeof(*p));
    p->veh_width_arg = 1.0; p->veh_width_m = 1.0; p->safety_margin_m = 0.05;
    p->lambda_smooth = 1.6e-3; p->max_outer_iters = 14; p->max_inner_iters = 120;
    p->step_init = 0.65; p->step_min = 1e-6; p->armijo_c = 1e-5;
    p->kappa_eps = 1e-6; p->v_cap_mps = 27.0;
    p->mass_kg = 255.0; p->Cd = 0.30; p->A_front_m2 = 1.00; p->rho_air = 1.225; p->c_rr = 0.015;
    p->P_max_W = 80000.0;
    p->a_total_max = 1.17 * 9.81; p->a_lat_max = 11.0; p->a_long_acc_cap = 8.0; p->a_long_brake_cap = 11.0;
    p->w_time_gain = 1.0; p->time_gamma_power = 2.0; p->time_weight_use_inv_v = 0; p->inv_v_gain = 0.1;
    p->max_vpass_iters = 6; p->use_total_ge_lat = 1;
}
