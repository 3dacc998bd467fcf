// raceline_geom.cu -- the stage before the path, batched over tracks:
//
//   centerline_kernel      pipeline::make_centerline (main.cpp:1270-1279) = centerline::splineUniformResample
//                          (448-474) over Spline1D::fit / eval_with_deriv (404-446), and the spline part of
//                          pipeline::compute_geom_and_save's per-sample body (1311-1320, 1325-1326)
//   ring_distance_kernel   distancesToRings (513-524) for every row: the same exact ray / point-segment
//                          formulas as the corridor of the solver stages, rings streamed through shared memory
//
// This translation unit is compiled with -fmad=false: without FMA contraction the spline fit, the cubic evaluation
// and the ray formulas round exactly like the reference's x86-64 build (only atan2 / pow go through a different
// libm).  The natural-spline solve is a sequential recurrence (Thomas algorithm); it runs on two lanes per track
// (x and y), tracks in parallel -- a batch of tracks is the parallel dimension, as everywhere in this library.
#include "raceline_kernels.cuh"

namespace rl {

namespace {

constexpr int kGeomT = 256;
constexpr int kGeomK = 8;
constexpr int kGeomMaxPts = 4096;   // padded mid points per track (6 arrays of doubles in shared memory)

__device__ __forceinline__ double hstep(const double* s, int i) { const double v = s[i + 1] - s[i]; return (v > 1e-30) ? v : 1e-30; }   // main.cpp:414

// Spline1D::eval_with_deriv (main.cpp:435-445) with b, d of main.cpp:421-424 formed on the fly from a, c
__device__ __forceinline__ void spline_eval(const double* s, const double* a, const double* c, int n, double si,
                                            double& f, double& fp, double& fpp)
{
    int lo = 0, hi = n - 1;
    if (si <= s[0]) lo = 0;
    else if (si >= s[n - 1]) lo = n - 2;
    else { while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s[mid] <= si) lo = mid; else hi = mid; } }
    const double t = si - s[lo];
    const double h = hstep(s, lo);
    const double b = (a[lo + 1] - a[lo]) / h - (2.0 * c[lo] + c[lo + 1]) * h / 3.0;
    const double d = (c[lo + 1] - c[lo]) / (3.0 * h);
    f = a[lo] + b * t + c[lo] * t * t + d * t * t * t;
    fp = b + 2.0 * c[lo] * t + 3.0 * d * t * t;
    fpp = 2.0 * c[lo] + 6.0 * d * t;
}

__global__ void __launch_bounds__(kGeomT, 1) centerline_kernel(const GeomBatch G, int n_tracks)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int trk = blockIdx.x;
    if (trk >= n_tracks) return;
    const int tid = threadIdx.x;
    const long long m0 = G.mid_off[trk];
    const int nmid = (int)(G.mid_off[trk + 1] - m0);
    const bool closed = G.closed[trk] != 0;
    const int pad = closed ? 3 : 0;                       // main.cpp:1273
    const int M = nmid + 2 * pad;
    const int samples = G.samples[trk];
    double* sX = reinterpret_cast<double*>(smem_raw);
    double* sY = sX + M;
    double* sS = sY + M;
    double* sDm = sS + M;
    double* sCx = sDm + M;
    double* sCy = sCx + M;
    // padded point list (main.cpp:452-456)
    for (int q = tid; q < M; q += kGeomT) {
        int src = q - pad;
        if (src < 0) src += nmid;
        else if (src >= nmid) src -= nmid;
        sX[q] = G.mids_xy[2 * (m0 + src)];
        sY[q] = G.mids_xy[2 * (m0 + src) + 1];
    }
    __syncthreads();
    if (tid == 0) {                                       // cumulative chord length (main.cpp:459-460): a left-to-right sum
        double acc = 0.0;
        sS[0] = 0.0;
        for (int i = 1; i < M; ++i) {
            const double dx = sX[i] - sX[i - 1], dy = sY[i] - sY[i - 1];
            acc = acc + sqrt(dx * dx + dy * dy);
            sS[i] = acc;
        }
    }
    __syncthreads();
    // tridiagonal system of the natural spline (main.cpp:415-419); row j = 0..M-3, right-hand sides kept at index j+1
    const int nsys = M - 2;
    for (int j = tid; j < nsys; j += kGeomT) {
        const double h0 = hstep(sS, j), h1 = hstep(sS, j + 1);
        sDm[j] = 2.0 * (h0 + h1);
        sCx[j + 1] = 3.0 * ((sX[j + 2] - sX[j + 1]) / h1 - (sX[j + 1] - sX[j]) / h0);
        sCy[j + 1] = 3.0 * ((sY[j + 2] - sY[j + 1]) / h1 - (sY[j + 1] - sY[j]) / h0);
    }
    __syncthreads();
    if (tid < 2 && nsys > 0) {                            // Spline1D::triSolve (main.cpp:406-410); lane 0: x, lane 1: y
        double* rhs = (tid == 0 ? sCx : sCy) + 1;
        double dm_prev = sDm[0], r_prev = rhs[0];
        for (int j = 1; j < nsys; ++j) {
            const double w = hstep(sS, j - 1) / dm_prev;          // dl[j-1] / dm[j-1]
            const double dmj = sDm[j] - w * hstep(sS, j);         // dm[j] -= w * du[j-1]
            const double rj = rhs[j] - w * r_prev;
            if (tid == 0) sDm[j] = dmj;                           // both lanes compute the same bits; one stores
            rhs[j] = rj;
            dm_prev = dmj; r_prev = rj;
        }
        __syncwarp(0x3u);
        double r_next = rhs[nsys - 1] / sDm[nsys - 1];
        rhs[nsys - 1] = r_next;
        for (int j = nsys - 2; j >= 0; --j) {
            r_next = (rhs[j] - hstep(sS, j + 1) * r_next) / sDm[j];   // du[j] = h[j+1]
            rhs[j] = r_next;
        }
        (tid == 0 ? sCx : sCy)[0] = 0.0;
        (tid == 0 ? sCx : sCy)[M - 1] = 0.0;
    }
    __syncthreads();
    const double s0 = sS[pad], s1 = sS[M - pad - 1];
    const double dl = s1 - s0;
    const double L = (dl > 1e-30) ? dl : 1e-30;           // main.cpp:464
    const long long r0 = G.row_off[trk];
    const int rows = (int)(G.row_off[trk + 1] - r0);
    const int denomN = closed ? samples : (samples > 1 ? samples : 1);
    if (tid == 0) { if (G.track_L) G.track_L[trk] = L; if (G.track_s0) G.track_s0[trk] = s0; }
    for (int k = tid; k < rows; k += kGeomT) {            // main.cpp:1311-1326
        const double si = s0 + L * ((double)k / (double)denomN);
        double x, xp, xpp, y, yp, ypp;
        spline_eval(sS, sX, sCx, M, si, x, xp, xpp);
        spline_eval(sS, sY, sCy, M, si, y, yp, ypp);
        const double hd = atan2(yp, xp);
        const double speed2 = xp * xp + yp * yp;
        const double denom = pow((speed2 > 1e-12) ? speed2 : 1e-12, 1.5);
        const double curv = (xp * ypp - yp * xpp) / denom;
        const double nn = sqrt((-yp) * (-yp) + xp * xp);  // geom::normalize(Vec2{-yp, xp}, 1e-12), main.cpp:132
        double nx = 0.0, ny = 0.0;
        if (!(nn < 1e-12)) { nx = -yp / nn; ny = xp / nn; }
        const double dk = (fabs(curv) > G.kappa_eps) ? fabs(curv) : G.kappa_eps;
        double vk = sqrt(G.a_lat_max / dk);
        if (vk > G.v_cap) vk = G.v_cap;
        G.xy[2 * (r0 + k)] = x; G.xy[2 * (r0 + k) + 1] = y;
        G.s_rel[r0 + k] = si - s0;
        G.heading[r0 + k] = hd;
        G.curvature[r0 + k] = curv;
        G.v_kappa[r0 + k] = vk;
        G.dist_inner[r0 + k] = nx;                        // the normal travels to ring_distance_kernel in the distance rows
        G.dist_outer[r0 + k] = ny;
    }
}

// distancesToRings (main.cpp:513-524) for the rows of one track: per ring the nearest +n / -n hit (one intersection
// serves both rays: den, t, u of the -n ray are -den, -t, u exactly), and the point-ring distance when neither ray
// hits.  Rings are streamed through shared memory in tiles with one bounding box per kSegBlock segments; a box the
// ray's LINE misses is skipped, everything accepted goes through the exact formulas of main.cpp:478-490 / 504-509.
__global__ void __launch_bounds__(kGeomT, 1) ring_distance_kernel(const GeomBatch G, int n_tracks)
{
    constexpr int T = kGeomT, K = kGeomK, NP = T * K;
    constexpr int CAP = ((4 * NP * 8) / (32 + 32 / SB)) / SB * SB;   // segments per tile (+ one box per SB segments)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2* sPt = reinterpret_cast<double2*>(smem_raw);
    double2* sNr = sPt + NP;
    double* sSeg = reinterpret_cast<double*>(sNr + NP);
    double* sBox = sSeg + 4 * CAP;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(sSeg + 4 * NP);
    const int trk = blockIdx.x;
    if (trk >= n_tracks) return;
    const int tid = threadIdx.x;
    const long long r0 = G.row_off[trk];
    const int rows = (int)(G.row_off[trk + 1] - r0);
    const long long segI0 = G.seg_off[2 * trk], segO0 = G.seg_off[2 * trk + 1], segE = G.seg_off[2 * trk + 2];
    const double INF = dinf();
    uint32_t bar_phase = 0;
    if (tid == 0) {
        mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    for (int c0 = 0; c0 < rows; c0 += NP) {
        const int nc = min(NP, rows - c0);
        __syncthreads();
        for (int i = tid; i < nc; i += T) {
            sPt[i] = make_double2(G.xy[2 * (r0 + c0 + i)], G.xy[2 * (r0 + c0 + i) + 1]);
            sNr[i] = make_double2(G.dist_inner[r0 + c0 + i], G.dist_outer[r0 + c0 + i]);
        }
        __syncthreads();
        double dring[2][K];
        for (int ring = 0; ring < 2; ++ring) {
            const long long base = ring ? segO0 : segI0;
            const int mr = (int)(ring ? (segE - segO0) : (segO0 - segI0));
            double pos[K], neg[K], dmin2[K];
#pragma unroll
            for (int j = 0; j < K; ++j) { pos[j] = INF; neg[j] = INF; dmin2[j] = INF; }
            const int ntiles = (mr + CAP - 1) / CAP;
            for (int pass = 0; pass < 2; ++pass) {
                if (pass == 1) {
                    bool need = false;
#pragma unroll
                    for (int j = 0; j < K; ++j) need = need || ((tid + j * T < nc) && pos[j] == INF && neg[j] == INF);
                    if (!__syncthreads_or(need)) break;
                }
                for (int tile = 0; tile < ntiles; ++tile) {
                    const int t0 = tile * CAP, nt = min(CAP, mr - t0), nblk = (nt + SB - 1) / SB;
                    if (pass == 0 || ntiles > 1) {
                        __syncthreads();
                        if (tid == 0) {
                            fence_proxy_async();
                            mbar_expect_tx(mbar, (uint32_t)nt * 32u);
                            bulk_g2s(sSeg, G.seg + 4 * (base + t0), (uint32_t)nt * 32u, mbar);
                        }
                        mbar_wait(mbar, bar_phase); bar_phase ^= 1;
                        for (int bq = tid; bq < nblk; bq += T) {
                            double xmin = INF, xmax = -INF, ymin = INF, ymax = -INF;
                            const int e = min(nt, bq * SB + SB);
                            for (int s = bq * SB; s < e; ++s) {
                                const double x0 = sSeg[4 * s], y0 = sSeg[4 * s + 1], x1 = sSeg[4 * s + 2], y1 = sSeg[4 * s + 3];
                                xmin = fmin(xmin, fmin(x0, x1)); xmax = fmax(xmax, fmax(x0, x1));
                                ymin = fmin(ymin, fmin(y0, y1)); ymax = fmax(ymax, fmax(y0, y1));
                                sSeg[4 * s + 2] = x1 - x0; sSeg[4 * s + 3] = y1 - y0;      // v = S1 - S0, main.cpp:482
                            }
                            const double cx = 0.5 * (xmin + xmax), cy = 0.5 * (ymin + ymax);
                            sBox[4 * bq] = cx; sBox[4 * bq + 1] = cy;
                            sBox[4 * bq + 2] = 0.5 * (xmax - xmin) + 1e-9 + 1e-12 * fabs(cx);
                            sBox[4 * bq + 3] = 0.5 * (ymax - ymin) + 1e-9 + 1e-12 * fabs(cy);
                        }
                        __syncthreads();
                    }
#pragma unroll
                    for (int j = 0; j < K; ++j) {
                        const int i = tid + j * T;
                        if (i >= nc) continue;
                        const double2 Pc = sPt[i];
                        const double nx = sNr[i].x, ny = sNr[i].y;
                        if (nx == 0.0 && ny == 0.0) continue;                 // main.cpp:1322: distances stay 0
                        if (pass == 0) {
                            double bp = pos[j], bn = neg[j];
                            for (int bq = 0; bq < nblk; ++bq) {
                                const double cx = sBox[4 * bq], cy = sBox[4 * bq + 1], hx = sBox[4 * bq + 2], hy = sBox[4 * bq + 3];
                                const double sc = nx * (cy - Pc.y) - ny * (cx - Pc.x);
                                const double ext = fabs(nx) * hy + fabs(ny) * hx + 1e-7;
                                if (fabs(sc) > ext) continue;                 // the line misses the box
                                const int e = min(nt, bq * SB + SB);
                                for (int s = bq * SB; s < e; ++s) {
                                    const double x0 = sSeg[4 * s], y0 = sSeg[4 * s + 1], vx = sSeg[4 * s + 2], vy = sSeg[4 * s + 3];
                                    const double den = nx * (-vy) + ny * vx;            // main.cpp:483
                                    if (fabs(den) < 1e-15) continue;                    // main.cpp:484
                                    const double ax = x0 - Pc.x, ay = y0 - Pc.y;        // main.cpp:485
                                    const double inv = 1.0 / den;
                                    const double t = (ax * (-vy) + ay * vx) * inv;      // main.cpp:486
                                    const double u = (nx * ay - ny * ax) * inv;         // main.cpp:487
                                    if (u >= -1e-12 && u <= 1.0 + 1e-12) {              // main.cpp:488
                                        if (t > 0.0) bp = fmin(bp, t);                  // +n ray, main.cpp:497
                                        else if (t < 0.0) bn = fmin(bn, -t);            // -n ray: t' = -t
                                    }
                                }
                            }
                            pos[j] = bp; neg[j] = bn;
                        } else if (pos[j] == INF && neg[j] == INF) {
                            // minDistanceToSegments_global (main.cpp:501-512), boxes pruned by the best distance so far
                            double best2 = dmin2[j], bound2 = dmin2[j] * (1.0 + 1e-9);
                            for (int bq = 0; bq < nblk; ++bq) {
                                const double cx = sBox[4 * bq], cy = sBox[4 * bq + 1], hx = sBox[4 * bq + 2], hy = sBox[4 * bq + 3];
                                const double ddx = fmax(0.0, fabs(cx - Pc.x) - hx), ddy = fmax(0.0, fabs(cy - Pc.y) - hy);
                                if (ddx * ddx + ddy * ddy > bound2) continue;
                                const int e = min(nt, bq * SB + SB);
                                for (int s = bq * SB; s < e; ++s) {
                                    const double x0 = sSeg[4 * s], y0 = sSeg[4 * s + 1], vx = sSeg[4 * s + 2], vy = sSeg[4 * s + 3];
                                    const double apx = Pc.x - x0, apy = Pc.y - y0;
                                    const double denom = fmax(1e-30, vx * vx + vy * vy);
                                    const double tt = fmin(1.0, fmax(0.0, (vx * apx + vy * apy) / denom));
                                    const double qx = x0 + vx * tt, qy = y0 + vy * tt;
                                    const double ex = Pc.x - qx, ey = Pc.y - qy;
                                    const double d2 = ex * ex + ey * ey;
                                    if (d2 < best2) { best2 = d2; bound2 = d2 * (1.0 + 1e-9); }
                                }
                            }
                            dmin2[j] = best2;
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < K; ++j) {
                // main.cpp:519-523: min of the two rays when either is finite, else the point-ring distance; non-finite -> 0
                double d = (pos[j] < INF || neg[j] < INF) ? fmin(pos[j], neg[j]) : ((dmin2[j] < INF) ? sqrt(dmin2[j]) : INF);
                if (!(d < INF)) d = 0.0;
                dring[ring][j] = d;
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const int i = tid + j * T;
            if (i >= nc) continue;
            const bool degenerate = (sNr[i].x == 0.0 && sNr[i].y == 0.0);
            const double di = degenerate ? 0.0 : dring[0][j], dout = degenerate ? 0.0 : dring[1][j];
            G.dist_inner[r0 + c0 + i] = di;
            G.dist_outer[r0 + c0 + i] = dout;
            G.width[r0 + c0 + i] = di + dout;
        }
    }
}

}  // namespace

size_t geom_centerline_smem(int max_pts) { return (size_t)max_pts * 6 * sizeof(double); }
size_t geom_ring_smem() { return (size_t)kGeomT * kGeomK * (16 + 16 + 32) + 64; }
int geom_max_points() { return kGeomMaxPts; }

int configure_geom()
{
    cudaError_t e = cudaFuncSetAttribute(centerline_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)geom_centerline_smem(kGeomMaxPts));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(ring_distance_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)geom_ring_smem());
    return (int)e;
}

int launch_geom(const GeomBatch& G, int n_tracks, int max_pts, void* stream)
{
    if (n_tracks <= 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    centerline_kernel<<<n_tracks, kGeomT, geom_centerline_smem(max_pts), s>>>(G, n_tracks);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    ring_distance_kernel<<<n_tracks, kGeomT, geom_ring_smem(), s>>>(G, n_tracks);
    return (int)cudaGetLastError();
}

}  // namespace rl
