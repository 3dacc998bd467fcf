/*
 * geom_oracle.c -- TEST INFRASTRUCTURE.  CPU restatement of the stage before the hot path:
 *
 *   centerline::Spline1D::fit / eval_with_deriv            src/main.cpp:404-446
 *   centerline::splineUniformResample                      src/main.cpp:448-474
 *   pipeline::make_centerline                              src/main.cpp:1270-1279
 *   pipeline::compute_geom_and_save, per-sample body       src/main.cpp:1306-1329
 *   distancesToRings                                        src/main.cpp:513-524   (orc_distances_to_rings)
 *
 * Pinned bit for bit by tests/golden/make_golden.py against oracle/_ref/ref_harness geom, which drives the
 * reference's own functions and checks itself against the *_with_geom.csv the reference writes.
 * Only tests/ and the checker legs of bench.py / __graft_entry__.smoke() may use it; never the product path.
 */
#include <math.h>
#include <stdlib.h>

#include "raceline_oracle.h"

typedef struct { int n; double *s, *a, *b, *c, *d; } spline1d;

/* Spline1D::triSolve, main.cpp:406-410 */
static void tri_solve(double* dl, double* dm, double* du, double* rhs, int n)
{
    for (int i = 1; i < n; ++i) { double w = dl[i - 1] / dm[i - 1]; dm[i] -= w * du[i - 1]; rhs[i] -= w * rhs[i - 1]; }
    rhs[n - 1] /= dm[n - 1];
    for (int i = n - 2; i >= 0; --i) rhs[i] = (rhs[i] - du[i] * rhs[i + 1]) / dm[i];
}

/* Spline1D::fit, main.cpp:411-426 (n >= 3) */
static int spline_fit(spline1d* sp, const double* s, const double* y, int n)
{
    sp->n = n;
    sp->s = (double*)malloc(sizeof(double) * 5 * (size_t)n);
    if (!sp->s) return -1;
    sp->a = sp->s + n; sp->b = sp->a + n; sp->c = sp->b + n; sp->d = sp->c + n;
    for (int i = 0; i < n; ++i) { sp->s[i] = s[i]; sp->a[i] = y[i]; sp->b[i] = 0.0; sp->c[i] = 0.0; sp->d[i] = 0.0; }
    double* h = (double*)malloc(sizeof(double) * 5 * (size_t)n);
    if (!h) return -1;
    double *dl = h + n, *dm = dl + n, *du = dm + n, *rhs = du + n;
    for (int i = 0; i < n - 1; ++i) { double v = s[i + 1] - s[i]; h[i] = (v > 1e-30) ? v : 1e-30; }
    for (int i = 1; i <= n - 2; ++i) {
        double hi_1 = h[i - 1], hi = h[i];
        dl[i - 1] = hi_1; dm[i - 1] = 2.0 * (hi_1 + hi); du[i - 1] = hi;
        rhs[i - 1] = 3.0 * ((y[i + 1] - y[i]) / hi - (y[i] - y[i - 1]) / hi_1);
    }
    if (n - 2 > 0) tri_solve(dl, dm, du, rhs, n - 2);
    for (int i = 1; i <= n - 2; ++i) sp->c[i] = rhs[i - 1];
    sp->c[0] = 0.0; sp->c[n - 1] = 0.0;
    for (int i = 0; i < n - 1; ++i) {
        sp->b[i] = (y[i + 1] - y[i]) / h[i] - (2.0 * sp->c[i] + sp->c[i + 1]) * h[i] / 3.0;
        sp->d[i] = (sp->c[i + 1] - sp->c[i]) / (3.0 * h[i]);
    }
    free(h);
    return 0;
}

/* Spline1D::eval_with_deriv, main.cpp:435-445 */
static void spline_eval(const spline1d* sp, double si, double* f, double* fp, double* fpp)
{
    int n = sp->n, lo = 0, hi = n - 1;
    if (si <= sp->s[0]) lo = 0;
    else if (si >= sp->s[n - 1]) lo = n - 2;
    else { while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (sp->s[mid] <= si) lo = mid; else hi = mid; } }
    double t = si - sp->s[lo];
    *f = sp->a[lo] + sp->b[lo] * t + sp->c[lo] * t * t + sp->d[lo] * t * t * t;
    *fp = sp->b[lo] + 2.0 * sp->c[lo] * t + 3.0 * sp->d[lo] * t * t;
    *fpp = 2.0 * sp->c[lo] + 6.0 * sp->d[lo] * t;
}

int orc_geom_rows(int samples, int closed, int emit_dup) { return closed ? samples : samples + (emit_dup ? 1 : 0); }

/* returns the number of rows written, or -1 (fewer than 3 mid points / allocation failure) */
int orc_centerline_geom(const double* mids_xy, int n_mid, int samples, int closed, int emit_dup,
                        const double* inner_seg, int m_in, const double* outer_seg, int m_out, const rl_params* C,
                        double* out_xy, double* s_rel, double* heading, double* curvature,
                        double* d_inner, double* d_outer, double* width, double* v_kappa, double* L_out, double* s0_out)
{
    if (n_mid < 3 || samples < 1) return -1;
    const int pad = closed ? 3 : 0;                      /* main.cpp:1273 */
    const int M = n_mid + 2 * pad;
    double* buf = (double*)malloc(sizeof(double) * 3 * (size_t)M);
    if (!buf) return -1;
    double *s = buf, *xs = s + M, *ys = xs + M;
    /* splineUniformResample, main.cpp:452-461 */
    int q = 0;
    for (int i = 0; i < pad; ++i, ++q) { xs[q] = mids_xy[2 * (n_mid - pad + i)]; ys[q] = mids_xy[2 * (n_mid - pad + i) + 1]; }
    for (int i = 0; i < n_mid; ++i, ++q) { xs[q] = mids_xy[2 * i]; ys[q] = mids_xy[2 * i + 1]; }
    for (int i = 0; i < pad; ++i, ++q) { xs[q] = mids_xy[2 * i]; ys[q] = mids_xy[2 * i + 1]; }
    s[0] = 0.0;
    for (int i = 1; i < M; ++i) { double dx = xs[i] - xs[i - 1], dy = ys[i] - ys[i - 1]; s[i] = s[i - 1] + sqrt(dx * dx + dy * dy); }
    spline1d spx, spy;
    if (spline_fit(&spx, s, xs, M) || spline_fit(&spy, s, ys, M)) { free(buf); return -1; }
    const double s0 = s[pad], s1 = s[M - pad - 1];
    const double L = (s1 - s0 > 1e-30) ? (s1 - s0) : 1e-30;       /* main.cpp:464 */
    const int rows = orc_geom_rows(samples, closed, emit_dup);
    const int denomN = closed ? samples : (samples > 1 ? samples : 1);
    for (int k = 0; k < rows; ++k) {                     /* main.cpp:1311-1329 */
        double si = s0 + L * ((double)k / (double)denomN);
        double x, xp, xpp, y, yp, ypp;
        spline_eval(&spx, si, &x, &xp, &xpp);
        spline_eval(&spy, si, &y, &yp, &ypp);
        double hd = atan2(yp, xp);
        double speed2 = xp * xp + yp * yp;
        double denom = pow((speed2 > 1e-12) ? speed2 : 1e-12, 1.5);
        double curv = (xp * ypp - yp * xpp) / denom;
        double nn = sqrt((-yp) * (-yp) + xp * xp);        /* geom::normalize(Vec2{-yp, xp}, 1e-12), main.cpp:132 */
        double nx = 0.0, ny = 0.0;
        if (!(nn < 1e-12)) { nx = -yp / nn; ny = xp / nn; }
        double di = 0.0, dout = 0.0;
        if (nx != 0 || ny != 0) orc_distances_to_rings(x, y, nx, ny, inner_seg, m_in, outer_seg, m_out, &di, &dout);
        double dk = (fabs(curv) > C->kappa_eps) ? fabs(curv) : C->kappa_eps;
        double vk = sqrt(C->a_lat_max / dk);
        if (vk > C->v_cap_mps) vk = C->v_cap_mps;
        if (out_xy) { out_xy[2 * k] = x; out_xy[2 * k + 1] = y; }
        if (s_rel) s_rel[k] = si - s0;
        if (heading) heading[k] = hd;
        if (curvature) curvature[k] = curv;
        if (d_inner) d_inner[k] = di;
        if (d_outer) d_outer[k] = dout;
        if (width) width[k] = di + dout;
        if (v_kappa) v_kappa[k] = vk;
    }
    if (L_out) *L_out = L;
    if (s0_out) *s0_out = s0;
    free(spx.s); free(spy.s); free(buf);
    return rows;
}
