// synth_tracks.cpp -- deterministic synthetic closed tracks (host only; bench/test workload generator).
//
// the reference ships 7 maps of N = 187..261 samples, so the
// large-N workloads are generated.  Track `id` is a polar "flower" curve
//     r(theta) = R0 * (1 + sum_{j<3} a_j sin(k_j theta + phi_j))
// with (k_j, a_j, phi_j) drawn from mt19937_64(seed_base + id); the amplitude scale and R0 are solved
// so that the length is N * 1.8 m (the shipped maps have h = L/N = 1.60..1.83 m) and max|kappa| hits a
// target drawn from [0.08, 0.25] 1/m.  A polar curve is star-shaped, hence free of self-intersections.  The centre line is
// the uniform arc-length resample of a 64x oversampled polyline (L = its length); the rings are the
// +-1.75 m normal offsets at M uniformly spaced arc positions, emitted as ring edges exactly like
// edges::ringEdges (reference main.cpp:251).  Same seed => same bits, on any machine.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <random>
#include <thread>
#include <vector>

#include "../../include/raceline_b200.h"

namespace {

constexpr double kPi = 3.14159265358979323846;
constexpr double kHalfWidth = 1.75;
constexpr int kOversample = 64;

struct Shape { int k[3]; double a[3]; double phi[3]; };

inline double uni(std::mt19937_64& g) { return (double)(g() >> 11) * (1.0 / 9007199254740992.0); }

// r, r', r'' of the unit polar curve r = 1 + a0 sin(k0 th + p0) + scale * (a1 sin(..) + a2 sin(..))
inline void polar(const Shape& s, double scale, double th, double& r, double& r1, double& r2)
{
    double f = 0, f1 = 0, f2 = 0;
    for (int j = 0; j < 3; ++j) {
        const double k = s.k[j], arg = k * th + s.phi[j], a = (j == 0) ? s.a[j] : scale * s.a[j];
        const double sn = std::sin(arg), cs = std::cos(arg);
        f += a * sn; f1 += a * k * cs; f2 -= a * k * k * sn;
    }
    r = 1.0 + f; r1 = f1; r2 = f2;
}

// length and max|curvature| of the unit curve, probed analytically at n_probe angles
void unit_metrics(const Shape& s, double scale, int n_probe, double& len, double& kmax)
{
    len = 0.0; kmax = 0.0;
    const double dth = 2.0 * kPi / n_probe;
    for (int i = 0; i < n_probe; ++i) {
        double r, r1, r2;
        polar(s, scale, dth * i, r, r1, r2);
        const double q = r * r + r1 * r1;
        const double sq = std::sqrt(q);
        len += sq * dth;
        kmax = std::max(kmax, std::fabs((r * r + 2.0 * r1 * r1 - r * r2) / (q * sq)));
    }
}

void make_track(uint64_t seed, int N, int M, double* center_xy, double* seg, double* L_out)
{
    std::mt19937_64 g(seed);
    Shape s;
    const double L0 = N * 1.8;                      // target length: h = L/N ~ 1.8 m like the shipped maps
    // wavelengths: large-scale shape, medium sweepers, tight corners
    s.k[0] = 2 + (int)(uni(g) * 4.0);
    s.k[2] = std::max(3, (int)std::lround(L0 / (25.0 + 20.0 * uni(g))));
    s.k[1] = std::max(2, (int)std::lround(L0 / (80.0 + 80.0 * uni(g))));
    s.a[0] = (s.k[0] * s.k[0] < N / 8) ? (0.05 + 0.10 * uni(g)) : 0.0;
    s.a[1] = (0.15 + 0.25 * uni(g)) / ((double)s.k[1] * s.k[1]);
    s.a[2] = (0.50 + 0.50 * uni(g)) / ((double)s.k[2] * s.k[2]);
    for (int j = 0; j < 3; ++j) s.phi[j] = 2.0 * kPi * uni(g);
    const double target = 0.08 + 0.17 * uni(g);     // max |kappa| in 1/m

    // Solve for the amplitude scale: with R0 = L0/len(scale) the physical max curvature is
    // kmax(scale)*len(scale)/L0; bisect on scale (bounded so the radial slope stays moderate).
    const int n_probe = N;
    double s_hi = 2.5 / std::max(1e-12, s.a[1] * s.k[1] + s.a[2] * s.k[2]);
    double s_lo = 0.0, len, kmax;
    unit_metrics(s, s_hi, n_probe, len, kmax);
    double scale = s_hi;
    if (kmax * len / L0 > target) {
        for (int it = 0; it < 14; ++it) {
            scale = 0.5 * (s_lo + s_hi);
            unit_metrics(s, scale, n_probe, len, kmax);
            if (kmax * len / L0 > target) s_hi = scale; else s_lo = scale;
        }
        scale = s_lo;
    }
    unit_metrics(s, scale, n_probe, len, kmax);
    const double R0 = L0 / len;

    // oversampled polyline (closed) and its cumulative length; sin/cos by rotation recurrence,
    // re-anchored with exact values every 64 steps
    const int Q = kOversample * N;
    std::vector<double> px((size_t)Q + 1), py((size_t)Q + 1), cs((size_t)Q + 1);
    {
        const double dth = 2.0 * kPi / (double)Q;
        double rc[4], rs[4], wc[4], ws[4];   // [0..2] harmonics, [3] theta itself
        for (int j = 0; j < 4; ++j) { const double k = (j < 3) ? s.k[j] : 1.0; wc[j] = std::cos(k * dth); ws[j] = std::sin(k * dth); }
        for (int i = 0; i < Q; ++i) {
            if ((i & 63) == 0) {
                const double th = dth * i;
                for (int j = 0; j < 4; ++j) {
                    const double arg = (j < 3) ? s.k[j] * th + s.phi[j] : th;
                    rc[j] = std::cos(arg); rs[j] = std::sin(arg);
                }
            }
            const double r = R0 * (1.0 + s.a[0] * rs[0] + scale * (s.a[1] * rs[1] + s.a[2] * rs[2]));
            px[i] = r * rc[3]; py[i] = r * rs[3];
            for (int j = 0; j < 4; ++j) {
                const double c2 = rc[j] * wc[j] - rs[j] * ws[j], s2 = rs[j] * wc[j] + rc[j] * ws[j];
                rc[j] = c2; rs[j] = s2;
            }
        }
    }
    px[Q] = px[0]; py[Q] = py[0];
    cs[0] = 0.0;
    for (int i = 1; i <= Q; ++i) {
        const double dx = px[i] - px[i - 1], dy = py[i] - py[i - 1];
        cs[i] = cs[i - 1] + std::sqrt(dx * dx + dy * dy);
    }
    const double L = cs[Q];
    *L_out = L;

    // uniform arc-length resample: centre samples
    {
        int j = 0;
        for (int i = 0; i < N; ++i) {
            const double t = L * (double)i / (double)N;
            while (j + 1 < Q && cs[j + 1] <= t) ++j;
            const double u = (t - cs[j]) / std::max(1e-300, cs[j + 1] - cs[j]);
            center_xy[2 * i] = px[j] + (px[j + 1] - px[j]) * u;
            center_xy[2 * i + 1] = py[j] + (py[j + 1] - py[j]) * u;
        }
    }
    // ring cones at M uniform arc positions, offset along the left/right normal of the local chord
    std::vector<double> rin((size_t)2 * M), rout((size_t)2 * M);
    {
        int j = 0;
        for (int c = 0; c < M; ++c) {
            const double t = L * ((double)c + 0.5) / (double)M;
            while (j + 1 < Q && cs[j + 1] <= t) ++j;
            const double u = (t - cs[j]) / std::max(1e-300, cs[j + 1] - cs[j]);
            const double x = px[j] + (px[j + 1] - px[j]) * u, y = py[j] + (py[j + 1] - py[j]) * u;
            // tangent over a wider stencil than one oversampled chord
            const int ja = (j - kOversample / 4 + Q) % Q, jb = (j + kOversample / 4) % Q;
            double tx = px[jb] - px[ja], ty = py[jb] - py[ja];
            const double tn = std::max(1e-300, std::hypot(tx, ty));
            tx /= tn; ty /= tn;
            rin[2 * c] = x - ty * kHalfWidth; rin[2 * c + 1] = y + tx * kHalfWidth;      // left of a CCW curve = inside
            rout[2 * c] = x + ty * kHalfWidth; rout[2 * c + 1] = y - tx * kHalfWidth;
        }
    }
    for (int c = 0; c < M; ++c) {
        const int d = (c + 1) % M;
        double* si = seg + 4 * (size_t)c;
        si[0] = rin[2 * c]; si[1] = rin[2 * c + 1]; si[2] = rin[2 * d]; si[3] = rin[2 * d + 1];
        double* so = seg + 4 * ((size_t)M + c);
        so[0] = rout[2 * c]; so[1] = rout[2 * c + 1]; so[2] = rout[2 * d]; so[3] = rout[2 * d + 1];
    }
}

}  // namespace

extern "C" int rl_synth_tracks(uint64_t seed_base, int64_t first_id, int n_tracks, int n_samples, int m_per_ring,
                               int n_threads, double* center_xy, double* seg, double* track_L)
{
    if (n_tracks < 0 || n_samples < 4 || m_per_ring < 3 || !center_xy || !seg || !track_L) return RL_ERR_ARG;
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    n_threads = std::min(n_threads, std::max(1, n_tracks));
    auto work = [&](int tid) {
        for (int t = tid; t < n_tracks; t += n_threads)
            make_track(seed_base + (uint64_t)(first_id + t), n_samples, m_per_ring,
                       center_xy + (size_t)2 * n_samples * t, seg + (size_t)8 * m_per_ring * t, track_L + t);
    };
    if (n_threads == 1) { work(0); return RL_OK; }
    std::vector<std::thread> th;
    for (int i = 0; i < n_threads; ++i) th.emplace_back(work, i);
    for (auto& x : th) x.join();
    return RL_OK;
}
