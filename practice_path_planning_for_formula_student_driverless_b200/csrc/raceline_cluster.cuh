// raceline_cluster.cuh -- long tracks: one thread-block CLUSTER per job (BASELINE configs[4], N up to 16,384).
//
// A track that does not fit one CTA (N > 4096) is cut into CS contiguous chunks, one per CTA of a cluster of
// CS = 4 or 8 CTAs (T = 256 threads, K = 8 samples per thread, up to 2048 samples per CTA).  Each CTA runs the
// same blocked layout as solve_kernel on its chunk; what crosses a chunk boundary travels through distributed
// shared memory (st.shared::cluster into the NEIGHBOUR's shared memory, made visible by the cluster barrier):
//
//   * the 2-sample alpha halos of the cost/gradient window at the two CTA edges,
//   * the per-warp partial sums of (J, g.dalpha): every warp stores its pair into every CTA of the cluster, and
//     after the barrier every thread reduces the same CS*8 pairs in the same order -> bit-identical Armijo
//     decisions in all CTAs with ONE cluster barrier per evaluation,
//   * one path point each way (normals / derivatives are 3-point stencils),
//   * the published end values of the relaxed v(s) sweeps and their "anything changed" flags,
//   * the stencil coefficients of the neighbour's edge sample.
//
// The corridor streams each ring through shared memory in tiles (TMA bulk copies), builds the FP32 two-level
// box hierarchy per tile and runs the same exact FP64 ray / distance searches as the single-CTA fast path
// (ray_scan / dist_scan); every CTA does this for its own samples only, with no cluster traffic.
//
// Closed tracks only (is_closed_track = true in every BASELINE config; open long tracks -> RL_ERR_UNSUPPORTED).
#pragma once
#include "raceline_kernels.cuh"

namespace rl {
namespace {

constexpr int kMaxCS = 8;        // portable cluster size limit
constexpr int kcT = 256;         // threads per CTA of the cluster kernel
constexpr int kcNW = kcT / 32;
// scratch layout (bytes)
constexpr int kcBar = 0;                              // mbarrier
constexpr int kcFlag = 16;                            // int[3]: cluster-wide OR flags (rotating)
constexpr int kcHalo = 32;                            // double2[2]: path point left of the chunk | right of the chunk
constexpr int kcCoef = 64;                            // double[8]: (c0,cp,cm,-) of the left neighbour's last | right neighbour's first sample
constexpr int kcRed = 128;                            // [2][kMaxCS][NW][2] doubles
constexpr int kcExF = kcRed + 2 * kMaxCS * kcNW * 16; // [2][NW+1][2] doubles; slot NW = right neighbour CTA's first two samples
constexpr int kcExL = kcExF + 2 * (kcNW + 1) * 16;    // [2][NW+1][2] doubles; slot NW = left neighbour CTA's last two samples
constexpr int kcMisc = kcExL + 2 * (kcNW + 1) * 16;   // a few ints
constexpr int kcBytes = 3072;
static_assert(kcMisc + 64 <= kcBytes, "cluster scratch layout");

struct Clu {
    uint32_t CS, rank, left, right;
    int n0, Nloc;     // this CTA's chunk [n0, n0+Nloc) of the track
};

__device__ __forceinline__ uint32_t cl_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cl_size() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
// shared::cluster address of `p` (a pointer into THIS CTA's shared memory) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t cl_map(const void* p, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
__device__ __forceinline__ void cl_st2(uint32_t a, double x, double y)
{
    asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(x), "d"(y) : "memory");
}
__device__ __forceinline__ void cl_st1(uint32_t a, double x) { asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(a), "d"(x) : "memory"); }
__device__ __forceinline__ void cl_st_u32(uint32_t a, uint32_t x) { asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(a), "r"(x) : "memory"); }
// cluster-wide barrier; release/acquire at cluster scope orders local and distributed shared memory and global memory
__device__ __forceinline__ void cl_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// Deterministic cluster-wide sum of two values, bit-identical in every thread of every CTA.  One cluster barrier.
// sRedPh: [kMaxCS][NW][2] doubles of the current phase.
__device__ __forceinline__ void cluster_sum2(double& a, double& b, double* sRedPh, const Clu& cl, int lane, int warp)
{
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane < (int)cl.CS) cl_st2(cl_map(sRedPh + 2 * ((int)cl.rank * kcNW + warp), (uint32_t)lane), a, b);
    cl_sync();
    const int ne = (int)cl.CS * kcNW;   // <= 64 pairs
    double sa = 0.0, sb = 0.0;
    if (lane < ne) { const double2 e = *reinterpret_cast<const double2*>(sRedPh + 2 * lane); sa = e.x; sb = e.y; }
    if (lane + 32 < ne) { const double2 e = *reinterpret_cast<const double2*>(sRedPh + 2 * (lane + 32)); sa += e.x; sb += e.y; }
    a = warp_sum(sa);   // xor butterfly: every lane ends with the same bits
    b = warp_sum(sb);
}

// cluster-wide OR.  sFlag: int[3] rotating slots (all zero at kernel start); slot advances per call.
__device__ __forceinline__ bool cluster_or(bool pred, int* sFlag, int& slot, const Clu& cl, int tid, int lane)
{
    const bool any = __any_sync(kFull, pred);
    if (any && lane < (int)cl.CS) cl_st_u32(cl_map(sFlag + slot, (uint32_t)lane), 1u);
    const int nxt = (slot == 2) ? 0 : slot + 1;
    if (tid == 0) sFlag[nxt] = 0;   // last read two barriers ago, next written after this barrier
    cl_sync();
    const bool r = sFlag[slot] != 0;
    slot = nxt;
    return r;
}

// the chunk's path with one point of each neighbour chunk
struct PathView {
    const double2* sP;
    const double2* halo;   // [0] = point before the chunk, [1] = point after the chunk
    int Nloc;
    __device__ __forceinline__ double2 at(int il) const { return (il < 0) ? halo[0] : ((il >= Nloc) ? halo[1] : sP[il]); }
};
__device__ __forceinline__ void normal_c(const PathView& pv, int il, double& nx, double& ny)
{
    const double2 Pm = pv.at(il - 1), Pp = pv.at(il + 1);
    normal_from_tangent((Pp.x - Pm.x) * 0.5, (Pp.y - Pm.y) * 0.5, nx, ny);   // main.cpp:584-592
}
__device__ __forceinline__ void derivs_c(const PathView& pv, int il, double h, double& xp, double& yp, double& xpp, double& ypp)
{
    derivs_central(pv.at(il - 1), pv.at(il), pv.at(il + 1), h, xp, yp, xpp, ypp);
}
// publish the chunk's end points into the neighbours' halo slots (followed by a cluster barrier)
__device__ __forceinline__ void exchange_path_halo(const double2* sP, double2* sHalo, const Clu& cl, int tid)
{
    if (tid == 0) { const double2 p = sP[0]; cl_st2(cl_map(&sHalo[1], cl.left), p.x, p.y); }
    if (tid == 32) { const double2 p = sP[cl.Nloc - 1]; cl_st2(cl_map(&sHalo[0], cl.right), p.x, p.y); }
    cl_sync();
}

// ---- halo exchange of the trial alpha --------------------------------------------------------------------
template <int K, int MODE>
__device__ __forceinline__ Halo halo_send_c(const double (&x)[K], const Part& pt, const Clu& cl, double* sExF, double* sExL)
{
    double F0, F1, L0, L1;
    edge_values<K, MODE>(x, pt.cnt, F0, F1, L0, L1);
    Halo h;
    h.l0 = __shfl_sync(kFull, L0, pt.srcL);
    h.l1 = __shfl_sync(kFull, L1, pt.srcL);
    h.r0 = __shfl_sync(kFull, F0, pt.srcR);
    h.r1 = __shfl_sync(kFull, F1, pt.srcR);
    if (pt.lane == 0) {
        if (pt.warp == 0) cl_st2(cl_map(sExF + 2 * kcNW, cl.left), F0, F1);       // I am the right neighbour of CTA `left`
        else { sExF[2 * pt.warp] = F0; sExF[2 * pt.warp + 1] = F1; }
    }
    if (pt.lane == 31) {
        if (pt.warp == kcNW - 1) cl_st2(cl_map(sExL + 2 * kcNW, cl.right), L0, L1);
        else { sExL[2 * pt.warp] = L0; sExL[2 * pt.warp + 1] = L1; }
    }
    return h;
}
__device__ __forceinline__ void halo_recv_c(Halo& h, const Part& pt, const double* sExF, const double* sExL)
{
    if (pt.lane == 0) { const int s = (pt.warp == 0) ? kcNW : pt.warp - 1; h.l0 = sExL[2 * s]; h.l1 = sExL[2 * s + 1]; }
    if (pt.lane == 31) { const int s = (pt.warp == kcNW - 1) ? kcNW : pt.warp + 1; h.r0 = sExF[2 * s]; h.r1 = sExF[2 * s + 1]; }
}

template <int K>
struct PgdCtxC {
    const double* sC0; const double* sCp; const double* sCm; double* sSt;
    double* sRed; double* sExF; double* sExL;
    double lamJ, armijo_c, step2, J, decp;
    int ph;
};
constexpr int kcRedStride = kMaxCS * kcNW * 2;   // doubles per phase
constexpr int kcExStride = (kcNW + 1) * 2;

// pgd_half of raceline_kernels.cuh with the cluster exchange: ONE cluster barrier per evaluation
template <int K, int MODE>
__device__ __forceinline__ bool pgd_half_c(const Part& pt, const Clu& cl, const double (&xa)[K], const Halo& ha, double (&xb)[K], Halo& hb,
                                           const double* sLo, const double* sHi, const double (&cL)[3], const double (&cR)[3],
                                           PgdCtxC<K>& c)
{
    constexpr int T = kcT;
    double gh[K];
    double Jz = 0.0, Sd = 0.0;
    eval_window<T, K, MODE>(xa, ha, pt, c.sC0, c.sCp, c.sCm, cL, cR, c.lamJ, Jz, Sd, gh);
    double Jn = fma(c.lamJ, Sd, Jz);
    double dec2p = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const double xn = clamp_box(fma(-c.step2, gh[k], xa[k]), sLo[k * T], sHi[k * T]);
        dec2p = fma(gh[k], xn - xa[k], dec2p);
        xb[k] = xn;
    }
    hb = halo_send_c<K, MODE>(xb, pt, cl, c.sExF + c.ph * kcExStride, c.sExL + c.ph * kcExStride);
    double dec = c.decp;
    cluster_sum2(Jn, dec, c.sRed + c.ph * kcRedStride, cl, pt.lane, pt.warp);
    halo_recv_c(hb, pt, c.sExF + c.ph * kcExStride, c.sExL + c.ph * kcExStride);
    c.ph ^= 1;
    dec *= 2.0;                                    // gh is grad/2 (main.cpp:733)
    if (Jn <= c.J + c.armijo_c * dec) {            // Armijo accept, main.cpp:734
#pragma unroll
        for (int k = 0; k < K; ++k) c.sSt[k * T] = xa[k];
        c.decp = dec2p; c.J = Jn;
        return true;
    }
    return false;
}

// pgd_outer of raceline_kernels.cuh on a cluster (main.cpp:723-742 / 996-1026)
template <int K, int MODE>
__device__ __forceinline__ PgdOut pgd_outer_c(const Part& pt, const Clu& cl, const double* sLo, const double* sHi,
                                              const double (&cL)[3], const double (&cR)[3],
                                              const double* sC0, const double* sCp, const double* sCm, double* sSt,
                                              double* sRed, double* sExF, double* sExL, int& ph,
                                              double lamJ, double step_init, double step_min, double armijo_c, int max_inner)
{
    constexpr int T = kcT;
    PgdOut o; o.acc = 0; o.bt = 0; o.ev = 0;
    PgdCtxC<K> c;
    c.sC0 = sC0; c.sCp = sCp; c.sCm = sCm; c.sSt = sSt; c.sRed = sRed; c.sExF = sExF; c.sExL = sExL;
    c.lamJ = lamJ; c.armijo_c = armijo_c; c.ph = ph;
    c.step2 = 2.0 * step_init;
    double x[K], y[K];
    Halo hx, hy;
    hx.l0 = hx.l1 = hx.r0 = hx.r1 = 0.0;
    hy = hx;
#pragma unroll
    for (int k = 0; k < K; ++k) { x[k] = 0.0; y[k] = 0.0; sSt[k * T] = 0.0; }
    {
        double gh[K];
        double Jz = 0.0, Sd = 0.0;
        eval_window<T, K, MODE>(x, hx, pt, sC0, sCp, sCm, cL, cR, lamJ, Jz, Sd, gh);
        o.ev++;
        double Jt = fma(lamJ, Sd, Jz);
        double decp = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const double xn = clamp_box(-c.step2 * gh[k], sLo[k * T], sHi[k * T]);
            decp = fma(gh[k], xn, decp);
            x[k] = xn;
        }
        hx = halo_send_c<K, MODE>(x, pt, cl, sExF + c.ph * kcExStride, sExL + c.ph * kcExStride);
        double zero = 0.0;
        cluster_sum2(Jt, zero, sRed + c.ph * kcRedStride, cl, pt.lane, pt.warp);
        halo_recv_c(hx, pt, sExF + c.ph * kcExStride, sExL + c.ph * kcExStride);
        c.ph ^= 1;
        c.J = Jt; c.decp = decp;
        o.J0 = Jt;
    }
    double Jprev = c.J;
    int it = 0, bt = 0;
    bool in_x = true;
    while (it < max_inner) {
        const bool acc = in_x ? pgd_half_c<K, MODE>(pt, cl, x, hx, y, hy, sLo, sHi, cL, cR, c)
                              : pgd_half_c<K, MODE>(pt, cl, y, hy, x, hx, sLo, sHi, cL, cR, c);
        o.ev++;
        if (acc) {
            in_x = !in_x;
            o.acc++; it++; bt = 0;
            if (fabs(Jprev - c.J) < 1e-10) break;       // main.cpp:740
            Jprev = c.J;
        } else {
            c.step2 *= 0.5; bt++; o.bt++;               // main.cpp:737
            if (0.5 * c.step2 < step_min || bt >= 20) break;
            double a[K], gh[K];
#pragma unroll
            for (int k = 0; k < K; ++k) a[k] = sSt[k * T];
            Halo ha = halo_send_c<K, MODE>(a, pt, cl, sExF + c.ph * kcExStride, sExL + c.ph * kcExStride);
            cl_sync();
            halo_recv_c(ha, pt, sExF + c.ph * kcExStride, sExL + c.ph * kcExStride);
            c.ph ^= 1;
            double jz = 0.0, sd = 0.0;
            eval_window<T, K, MODE>(a, ha, pt, sC0, sCp, sCm, cL, cR, lamJ, jz, sd, gh);
            double decp = 0.0;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const double xn = clamp_box(fma(-c.step2, gh[k], a[k]), sLo[k * T], sHi[k * T]);
                decp = fma(gh[k], xn - a[k], decp);
                x[k] = xn;
            }
            c.decp = decp;
            hx = halo_send_c<K, MODE>(x, pt, cl, sExF + c.ph * kcExStride, sExL + c.ph * kcExStride);
            cl_sync();
            halo_recv_c(hx, pt, sExF + c.ph * kcExStride, sExL + c.ph * kcExStride);
            c.ph ^= 1;
            in_x = true;
        }
    }
    ph = c.ph;
    o.Jend = c.J;
    return o;
}

// ---- v(s) profile on a cluster: vprofile_blocked with the end values of each CTA published to its neighbour ----
// sX: 6*(T+1) doubles.  Index T of each array is the neighbour CTA's edge thread.
template <int K>
__device__ __forceinline__ void vprofile_c(const Part& pt, const Clu& cl, const VPar& q, const double (&kap)[K], double (&v)[K],
                                           int max_iters, double* sX, int* sFlag, int& fslot, int& rounds)
{
    constexpr int T = kcT, S = kcT + 1;
    double* sVL = sX;            // [2][S] last-slot value of each thread; [T] = last thread of the left neighbour CTA
    double* sVF = sX + 2 * S;    // [2][S] first-slot value; [T] = first thread of the right neighbour CTA
    double* sKF = sX + 4 * S;    // [S] kappa of the first slot; [T] = right neighbour CTA
    double* sKL = sX + 5 * S;    // [S] kappa of the last slot; [T] = left neighbour CTA
    const int cnt = pt.cnt, tid = pt.tid;
    const bool gfirst = (cl.rank == 0 && tid == 0), glast = (cl.rank == cl.CS - 1 && tid == T - 1);
    const int iL = (tid == 0) ? T : tid - 1, iR = (tid == T - 1) ? T : tid + 1;
#pragma unroll
    for (int k = 0; k < K; ++k)
        v[k] = (k < cnt) ? fmin(q.v_cap, sqrt(q.a_lat_max / fmax(fabs(kap[k]), q.kappa_eps))) : 0.0;   // main.cpp:787-794
    // region B of EVERY CTA must be free before a neighbour stores into it: the corridor (tiles in region B) has no
    // cluster barrier of its own, so a CTA that finished early could otherwise overwrite a neighbour's ring tile
    cl_sync();
    {
        double kl = kap[0];
#pragma unroll
        for (int k = 1; k < K; ++k) if (k < cnt) kl = kap[k];
        sKF[tid] = kap[0]; sKL[tid] = kl;
        if (tid == 0) cl_st1(cl_map(&sKF[T], cl.left), kap[0]);
        if (tid == T - 1) cl_st1(cl_map(&sKL[T], cl.right), kl);
    }
    cl_sync();
    const double kapL = sKL[iL], kapR = sKF[iR];
    int b = 0;
    for (int iter = 0; iter < max_iters; ++iter) {
        double v0[K], vstart[K];
#pragma unroll
        for (int k = 0; k < K; ++k) { vstart[k] = v[k]; v0[k] = v[k]; }
        // ---------------- forward sweep (main.cpp:829-833) ----------------
        {
            double last = v[0];
#pragma unroll
            for (int k = 1; k < K; ++k) if (k < cnt) last = v[k];
            sVL[b * S + tid] = last;
            if (tid == T - 1) cl_st1(cl_map(&sVL[b * S + T], cl.right), last);
        }
        cl_sync();
        for (;;) {
            const double vin = sVL[b * S + iL];
            double u = gfirst ? v0[0] : fmin(v0[0], f_acc(q, vin, kapL));
            v[0] = u;
#pragma unroll
            for (int k = 1; k < K; ++k)
                if (k < cnt) { u = fmin(v0[k], f_acc(q, u, kap[k - 1])); v[k] = u; }
            const bool changed = (u != sVL[b * S + tid]);
            sVL[(b ^ 1) * S + tid] = u;
            if (tid == T - 1) cl_st1(cl_map(&sVL[(b ^ 1) * S + T], cl.right), u);
            b ^= 1;
            ++rounds;
            if (!cluster_or(changed, sFlag, fslot, cl, tid, pt.lane)) break;
        }
        // closed-loop wrap: v[0] = min(v[0], f_acc(v[N-1])), main.cpp:834-839
        if (gfirst) v[0] = fmin(v[0], f_acc(q, sVL[b * S + iL], kapL));
        // ---------------- backward sweep (main.cpp:841-845) ----------------
#pragma unroll
        for (int k = 0; k < K; ++k) v0[k] = v[k];
        sVF[b * S + tid] = v[0];
        if (tid == 0) cl_st1(cl_map(&sVF[b * S + T], cl.left), v[0]);
        cl_sync();
        for (;;) {
            const double vin = sVF[b * S + iR];
            double u = 0.0;
#pragma unroll
            for (int k = K - 1; k >= 0; --k) {
                if (k == cnt - 1) { u = glast ? v0[k] : fmin(v0[k], f_brk(q, vin, kapR)); v[k] = u; }
                else if (k < cnt - 1) { u = fmin(v0[k], f_brk(q, u, kap[k + 1])); v[k] = u; }
            }
            const bool changed = (u != sVF[b * S + tid]);
            sVF[(b ^ 1) * S + tid] = u;
            if (tid == 0) cl_st1(cl_map(&sVF[(b ^ 1) * S + T], cl.left), u);
            b ^= 1;
            ++rounds;
            if (!cluster_or(changed, sFlag, fslot, cl, tid, pt.lane)) break;
        }
        // closed-loop wrap: v[N-1] = min(v[N-1], f_brk(v[0])), main.cpp:846-850
        if (glast) {
            const double w = f_brk(q, sVF[b * S + iR], kapR);
#pragma unroll
            for (int k = 0; k < K; ++k) if (k == cnt - 1) v[k] = fmin(v[k], w);
        }
        bool chg_iter = false;
#pragma unroll
        for (int k = 0; k < K; ++k) if (k < cnt && v[k] != vstart[k]) chg_iter = true;
        if (!cluster_or(chg_iter, sFlag, fslot, cl, tid, pt.lane)) break;
    }
}

// ax and lap time, main.cpp:854-860 (closed track: the last sample's successor is sample 0)
template <int K>
__device__ __forceinline__ double lap_and_ax_c(const Part& pt, const Clu& cl, const VPar& q, const double (&v)[K], double (&ax)[K],
                                               double* sX, double* sRedPh)
{
    constexpr int T = kcT;
    double* sVF = sX;   // [T+1]
    sVF[pt.tid] = v[0];
    if (pt.tid == 0) cl_st1(cl_map(&sVF[T], cl.left), v[0]);
    cl_sync();
    const double vnext_edge = sVF[(pt.tid == T - 1) ? T : pt.tid + 1];
    double t = 0.0, dummy = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        ax[k] = 0.0;
        if (k < pt.cnt) {
            const double v0 = v[k];
            double v1 = vnext_edge;
            if (k + 1 < K) { if (k + 1 < pt.cnt) v1 = v[k + 1]; }
            ax[k] = (v1 * v1 - v0 * v0) / (2.0 * q.h);
            t += q.h / fmax(1e-6, v0);
        }
    }
    cluster_sum2(t, dummy, sRedPh, cl, pt.lane, pt.warp);
    return t;
}

// ---- corridor of one chunk, rings streamed through region B in tiles (main.cpp:694-711, 749-756) ----------------
// Per ring: exact nearest +n / -n hits over all tiles (ray_scan, pruned by this ring's own best hit so far); a ring
// some ray misses entirely falls back to the exact point-ring distance (dist_scan over all tiles), main.cpp:696.
// Consecutive mapping (sample il = tid + j*T) inside the search, results handed to the blocked layout through region B.
template <int K>
__device__ __forceinline__ void corridor_stream_c(const Part& pt, const PathView& pv, double* sB, uint64_t* mbar, uint32_t& bar_phase,
                                                  int* sMisc, const double* __restrict__ gseg, long long segI0, long long segO0,
                                                  long long segE, double guard, double (&lo)[K], double (&hi)[K], long long& ray_tests)
{
    constexpr int T = kcT, NP = kcT * K;
    constexpr int CAP = fast_tile_cap(NP);
    const int Nl = pv.Nloc, tid = pt.tid;
    const double INF = dinf();
    const double2 org = pv.sP[0];
    double dpos[K], dneg[K];
#pragma unroll
    for (int j = 0; j < K; ++j) { dpos[j] = INF; dneg[j] = INF; }
    for (int ring = 0; ring < 2; ++ring) {
        const long long base = ring ? segO0 : segI0;
        const int mr = (int)(ring ? (segE - segO0) : (segO0 - segI0));
        if (mr == 0) {   // safe_ray on an empty ring returns 0 (main.cpp:696-697)
#pragma unroll
            for (int j = 0; j < K; ++j) { dpos[j] = fmin(dpos[j], 0.0); dneg[j] = fmin(dneg[j], 0.0); }
            continue;
        }
        double pos[K], neg[K], dmin[K];
#pragma unroll
        for (int j = 0; j < K; ++j) { pos[j] = INF; neg[j] = INF; dmin[j] = INF; }
        const int ntiles = (mr + CAP - 1) / CAP;
        RayTile tl;
        float m0 = 0.f;
        for (int pass = 0; pass < 2; ++pass) {
            if (pass == 1) {
                bool need = false;
#pragma unroll
                for (int j = 0; j < K; ++j) need = need || ((tid + j * T < Nl) && (pos[j] == INF || neg[j] == INF));
                if (!block_or<T>(need)) break;
            }
            for (int tile = 0; tile < ntiles; ++tile) {
                const int t0 = tile * CAP, nt = min(CAP, mr - t0);
                if (pass == 0 || ntiles > 1) {
                    bool chain;
                    m0 = ring_tile_build<T, K>(pt, sB, mbar, bar_phase, sMisc, gseg + 4 * (base + t0), nt, org.x, org.y, tl, chain);
                }
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    const int il = tid + j * T;
                    if (il >= Nl) continue;
                    const double2 Pc = pv.sP[il];
                    const float px = (float)(Pc.x - org.x), py = (float)(Pc.y - org.y);
                    const float m = m0 + 2e-6f * fmaxf(fabsf(px), fabsf(py));
                    if (pass == 0) {
                        double nx, ny;
                        normal_c(pv, il, nx, ny);
                        double bp = pos[j], bn = neg[j];
                        ray_scan(tl, Pc, nx, ny, px, py, (float)nx, (float)ny, m, 0, true, true, false, bp, bn, pos[j], neg[j], ray_tests);
                    } else if (pos[j] == INF || neg[j] == INF) {
                        const double d = dist_scan(tl, Pc, px, py, m, 0, fmin(dmin[j], fmin(pos[j], neg[j])));
                        dmin[j] = fmin(dmin[j], d);
                    }
                }
            }
        }
        // safe_ray + min over the two rings (main.cpp:696-705)
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const double dist = (dmin[j] < INF) ? dmin[j] : 0.0;
            const double vp = (pos[j] < INF) ? pos[j] : dist;
            const double vn = (neg[j] < INF) ? neg[j] : dist;
            dpos[j] = fmin(dpos[j], fmax(0.0, vp));
            dneg[j] = fmin(dneg[j], fmax(0.0, vn));
        }
    }
    // hi/lo (main.cpp:707-710), handed to the blocked layout through region B
    block_sync<T>();
    double* sLoS = sB;
    double* sHiS = sB + NP;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const int il = tid + j * T;
        if (il < Nl) {
            double hv = fmax(0.0, dpos[j] - guard);
            double lv = -fmax(0.0, dneg[j] - guard);
            if (!isfinite(hv)) hv = 0.0;
            if (!isfinite(lv)) lv = 0.0;
            sHiS[il] = hv; sLoS[il] = lv;
        }
    }
    block_sync<T>();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        lo[k] = 0.0; hi[k] = 0.0;
        if (k < pt.cnt) { lo[k] = sLoS[pt.start + k]; hi[k] = sHiS[pt.start + k]; }
    }
    block_sync<T>();
}

// ---- the cluster solver kernel -----------------------------------------------------------------------------
template <int K, int MODE>
__global__ void __launch_bounds__(kcT, 2)
solve_cluster_kernel(const DevBatch B, const int* __restrict__ job_list, int n_list)
{
    constexpr int T = kcT, NP = kcT * K;
    constexpr bool EXACT = (MODE == kModeExact);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2* sP = reinterpret_cast<double2*>(smem_raw);
    double* sB = reinterpret_cast<double*>(smem_raw + (size_t)NP * 16);
    unsigned char* scr = smem_raw + (size_t)NP * 16 + (size_t)NP * 32;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(scr + kcBar);
    int* sFlag = reinterpret_cast<int*>(scr + kcFlag);
    double2* sHalo = reinterpret_cast<double2*>(scr + kcHalo);
    double* sCoef = reinterpret_cast<double*>(scr + kcCoef);
    double* sRed = reinterpret_cast<double*>(scr + kcRed);
    double* sExF = reinterpret_cast<double*>(scr + kcExF);
    double* sExL = reinterpret_cast<double*>(scr + kcExL);
    int* sMisc = reinterpret_cast<int*>(scr + kcMisc);

    Clu cl;
    cl.CS = cl_size(); cl.rank = cl_rank();
    cl.left = (cl.rank == 0) ? cl.CS - 1 : cl.rank - 1;
    cl.right = (cl.rank == cl.CS - 1) ? 0 : cl.rank + 1;
    const int cid = (int)(blockIdx.x / cl.CS);
    if (cid >= n_list) return;   // uniform over the cluster
    const int jid = job_list[cid];
    const rl_job job = B.jobs[jid];
    const rl_params& C = B.params[job.param];
    rl_job_stats* st = B.stats + jid;
    const int trk = job.track;
    const long long s0 = B.samp_off[trk];
    const int N = (int)(B.samp_off[trk + 1] - s0);
    const bool ev = (job.stage == RL_STAGE_EVAL);      // profile of the given path only
    const bool mt = (job.stage == RL_STAGE_MINTIME) || ev;
    const double h = B.track_L[trk] / (double)N;
    {
        const int base = N / (int)cl.CS, rem = N % (int)cl.CS, r = (int)cl.rank;
        cl.Nloc = base + (r < rem ? 1 : 0);
        cl.n0 = r * base + min(r, rem);
    }
    const long long row0 = B.job_off[jid] + cl.n0;   // first output row of this chunk
    const int Nl = cl.Nloc;

    // ---- blocked partition of the chunk over the threads (the host guarantees Nl >= 2*T: every thread owns >= 2 samples) ----
    Part pt;
    pt.N = Nl; pt.tid = threadIdx.x; pt.lane = pt.tid & 31; pt.warp = pt.tid >> 5;
    {
        const int tid = pt.tid;
        pt.Tact = T;
        const int Kc = EXACT ? K : (Nl + T - 1) / T;
        const int nfull = EXACT ? T : Nl - T * (Kc - 1);
        pt.cnt = EXACT ? K : (tid < nfull ? Kc : Kc - 1);
        pt.start = EXACT ? tid * K : (tid < nfull ? tid * Kc : nfull * Kc + (tid - nfull) * (Kc - 1));
        pt.tL = (tid == 0) ? T - 1 : tid - 1;
        pt.tR = (tid == T - 1) ? 0 : tid + 1;
        pt.cntL = EXACT ? K : (pt.tL < nfull ? Kc : Kc - 1);
        pt.srcL = (pt.lane + 31) & 31; pt.srcR = (pt.lane + 1) & 31;
    }
    const int tid = pt.tid, cnt = pt.cnt, start = pt.start;
    PathView pv; pv.sP = sP; pv.halo = sHalo; pv.Nloc = Nl;

    uint32_t bar_phase = 0;
    if (tid == 0) {
        mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        sFlag[0] = 0; sFlag[1] = 0; sFlag[2] = 0;
        if (cl.rank == 0) {
            st->status = RL_OK; st->n = N; st->outer_done = 0; st->accepted = 0; st->backtracks = 0; st->evals = 0;
            st->vpass_rounds = 0; st->exist_scans = 0; st->ray_tests = 0; st->lap_time = 0.0;
            for (int o = 0; o < RL_MAX_OUTER_LOG; ++o) {
                st->J0[o] = 0.0; st->Jend[o] = 0.0; st->lap_outer[o] = 0.0; st->acc_outer[o] = 0; st->bt_outer[o] = 0;
            }
        }
        fence_proxy_async();
    }
    cl_sync();   // every CTA of the cluster is running and initialised before any distributed-shared-memory access

    // ---- load this chunk of the centre line (TMA bulk copy) and trade end points with the neighbours ----
    if (tid == 0) {
        mbar_expect_tx(mbar, (uint32_t)Nl * 16u);
        bulk_g2s(sP, B.center_xy + 2 * (s0 + cl.n0), (uint32_t)Nl * 16u, mbar);
    }
    mbar_wait(mbar, bar_phase); bar_phase ^= 1;
    exchange_path_halo(sP, sHalo, cl, tid);

#pragma unroll
    for (int k = 0; k < K; ++k)
        if (k < cnt) { B.alpha_total[row0 + start + k] = 0.0; B.alpha_last[row0 + start + k] = 0.0; }

    VPar q;
    q.v_cap = C.v_cap_mps; q.a_lat_max = C.a_lat_max; q.kappa_eps = C.kappa_eps;
    {
        const double a_total = C.use_total_ge_lat ? fmax(C.a_total_max, C.a_lat_max) : C.a_total_max;   // main.cpp:802-804
        q.a_tot2 = a_total * a_total;
    }
    q.kd = 0.5 * C.rho_air * C.Cd * C.A_front_m2; q.Fr = C.mass_kg * 9.81 * C.c_rr; q.mass = C.mass_kg; q.P = C.P_max_W;
    q.acc_cap = C.a_long_acc_cap; q.brk_cap = C.a_long_brake_cap; q.h = h; q.has_power = (C.P_max_W > 0);

    const long long segI0 = B.seg_off[2 * trk], segO0 = B.seg_off[2 * trk + 1], segE = B.seg_off[2 * trk + 2];
    const double inv2h = 1.0 / (2 * h), invh2 = 1.0 / (h * h);          // DiffOps, main.cpp:547
    const double lamJ = C.lambda_smooth * inv2h * inv2h;
    long long ray_tests = 0;
    int vrounds = 0, ph = 0, fslot = 0;
    int acc_total = 0, bt_total = 0, ev_total = 0;
    const int max_outer = ev ? 0 : C.max_outer_iters;

    double lo[K], hi[K];
#pragma unroll
    for (int k = 0; k < K; ++k) { lo[k] = 0.0; hi[k] = 0.0; }
    // initial corridor from the centre line: guard uses the veh_width ARGUMENT (main.cpp:706 / 930)
    if (!ev)
        corridor_stream_c<K>(pt, pv, sB, mbar, bar_phase, sMisc, B.seg, segI0, segO0, segE,
                             C.veh_width_arg * 0.5 + C.safety_margin_m, lo, hi, ray_tests);

    double* sC0 = sB + tid;
    double* sCp = sB + NP + tid;
    double* sCm = sB + 2 * NP + tid;
    double* sSt = sB + 3 * NP + tid;

    for (int outer = 0; outer < max_outer; ++outer) {
        // =================== linearisation (main.cpp:722 / 941-944) ===================
        double A1[K], A2[K], N0[K], Wd[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            A1[k] = 0.0; A2[k] = 0.0; N0[k] = 0.0; Wd[k] = 1.0;
            if (k < cnt) {
                const int il = start + k;
                double nx, ny, xp, yp, xpp, ypp;
                normal_c(pv, il, nx, ny);
                derivs_c(pv, il, h, xp, yp, xpp, ypp);
                A1[k] = nx * ypp - ny * xpp;          // main.cpp:644-646
                A2[k] = xp * ny - yp * nx;
                N0[k] = xp * ypp - yp * xpp;
                Wd[k] = pow15(xp * xp + yp * yp);     // denom; W = 1/denom (main.cpp:647-648)
            }
        }
        // ---- park the chunk of the path in global memory while the PGD runs (its region holds lo|hi) ----
        fence_proxy_async();
        block_sync<T>();
        if (tid == 0) bulk_s2g_issue(B.xy + 2 * row0, sP, (uint32_t)Nl * 16u);
        block_sync<T>();
        double* sLo = reinterpret_cast<double*>(sP) + tid;
        double* sHi = sLo + NP;
#pragma unroll
        for (int k = 0; k < K; ++k) { sLo[k * T] = lo[k]; sHi[k * T] = hi[k]; }
        double gam[K];
#pragma unroll
        for (int k = 0; k < K; ++k) gam[k] = 1.0;
        double lap_outer = 0.0;
        if (mt) {
            // ============ v(s) profile + time weights (main.cpp:944-977) ============
            double kap[K], vv[K], axd[K];
#pragma unroll
            for (int k = 0; k < K; ++k) kap[k] = (k < cnt) ? N0[k] / Wd[k] : 0.0;    // kappa, main.cpp:618
            block_sync<T>();
            vprofile_c<K>(pt, cl, q, kap, vv, C.max_vpass_iters, sB, sFlag, fslot, vrounds);
            block_sync<T>();
            lap_outer = lap_and_ax_c<K>(pt, cl, q, vv, axd, sB, sRed + ph * kcRedStride);
            ph ^= 1;
            double v_avg = 0.0;
            if (C.time_weight_use_inv_v) {            // main.cpp:951
                double sv = 0.0, dz = 0.0;
#pragma unroll
                for (int k = 0; k < K; ++k) if (k < cnt) sv += vv[k];
                cluster_sum2(sv, dz, sRed + ph * kcRedStride, cl, pt.lane, pt.warp);
                ph ^= 1;
                v_avg = sv / (double)max(1, N);
            }
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (k < cnt) {                         // main.cpp:954-975
                    const double vk = sqrt(C.a_lat_max / fmax(fabs(kap[k]), C.kappa_eps));
                    double r = fmin(1.0, vv[k] / fmax(1e-6, vk));
                    r = r * r;
                    r = fmin(1.0, fmax(0.0, r));
                    const double pw = C.time_gamma_power;
                    const double rp = (pw == 2.0) ? r * r : ((pw == 1.0) ? r : pow(r, pw));
                    const double corner_w = 1.0 + C.w_time_gain * rp;
                    double invv_w = 1.0;
                    if (C.time_weight_use_inv_v) {
                        const double ratio = v_avg / fmax(1e-6, vv[k]);
                        invv_w = 1.0 + C.inv_v_gain * (ratio - 1.0);
                        if (invv_w < 1.0) invv_w = 1.0;
                        if (invv_w > 3.0) invv_w = 3.0;
                    }
                    gam[k] = corner_w * invv_w;
                }
            }
            block_sync<T>();
        }
        // ---- stencil coefficients into region B (slot-major); edge samples also go to the neighbour CTAs ----
        {
            double f0 = 0.0, fp = 0.0, fm = 0.0, l0 = 0.0, lp = 0.0, lm = 0.0;   // first / last owned sample
#pragma unroll
            for (int k = 0; k < K; ++k) {
                double c0 = 0.0, cp = 0.0, cm = 0.0;
                if (k < cnt) {
                    const double gw = gam[k] / Wd[k];
                    c0 = gw * N0[k];
                    const double c1 = gw * A1[k] * inv2h, c2 = gw * A2[k] * invh2;
                    cp = c1 + c2; cm = c2 - c1;
                    if (k == 0) { f0 = c0; fp = cp; fm = cm; }
                    l0 = c0; lp = cp; lm = cm;
                }
                sC0[k * T] = c0; sCp[k * T] = cp; sCm[k * T] = cm;
            }
            if (tid == 0) {   // my first sample is the right-hand halo of CTA `left`
                cl_st2(cl_map(sCoef + 4, cl.left), f0, fp);
                cl_st1(cl_map(sCoef + 6, cl.left), fm);
            }
            if (tid == T - 1) {
                cl_st2(cl_map(sCoef + 0, cl.right), l0, lp);
                cl_st1(cl_map(sCoef + 2, cl.right), lm);
            }
        }
        cl_sync();
        double cL[3], cR[3];
        {
            const int kl = pt.cntL - 1;
            const double* b0 = sB + pt.tL; const double* br = sB + pt.tR;
            cL[0] = b0[kl * T]; cL[1] = b0[NP + kl * T]; cL[2] = b0[2 * NP + kl * T];
            cR[0] = br[0]; cR[1] = br[NP]; cR[2] = br[2 * NP];
            if (tid == 0) { cL[0] = sCoef[0]; cL[1] = sCoef[1]; cL[2] = sCoef[2]; }
            if (tid == T - 1) { cR[0] = sCoef[4]; cR[1] = sCoef[5]; cR[2] = sCoef[6]; }
        }
        if (!EXACT) {
            // the first unused slot mirrors the right neighbour's first sample (position cnt+1 of the window)
            if (cnt < K) { sC0[cnt * T] = cR[0]; sCp[cnt * T] = cR[1]; sCm[cnt * T] = cR[2]; }
        }
        // =================== projected gradient with Armijo (main.cpp:723-742 / 996-1026) ===================
        const PgdOut po = pgd_outer_c<K, MODE>(pt, cl, sLo, sHi, cL, cR, sC0, sCp, sCm, sSt, sRed, sExF, sExL, ph, lamJ,
                                               C.step_init, C.step_min, C.armijo_c, C.max_inner_iters);
        acc_total += po.acc; bt_total += po.bt; ev_total += po.ev;
        if (tid == 0 && cl.rank == 0 && outer < RL_MAX_OUTER_LOG) {
            st->J0[outer] = po.J0; st->Jend[outer] = po.Jend; st->lap_outer[outer] = lap_outer;
            st->acc_outer[outer] = po.acc; st->bt_outer[outer] = po.bt;
        }
        // ---- bring the chunk of the path back ----
        block_sync<T>();
        if (tid == 0) {
            bulk_wait_all();
            fence_proxy_async();
            mbar_expect_tx(mbar, (uint32_t)Nl * 16u);
            bulk_g2s(sP, B.xy + 2 * row0, (uint32_t)Nl * 16u, mbar);
        }
        mbar_wait(mbar, bar_phase); bar_phase ^= 1;
        // =================== path update (main.cpp:743-746 / 1027-1031) ===================
        double2 Pn[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (k < cnt) {
                const int il = start + k;
                const double al = sSt[k * T];
                double nx, ny;
                normal_c(pv, il, nx, ny);
                const double2 Pc = sP[il];
                Pn[k].x = Pc.x + nx * al; Pn[k].y = Pc.y + ny * al;
                B.alpha_total[row0 + il] += al;
                if (outer == max_outer - 1) B.alpha_last[row0 + il] = al;
            }
        }
        cl_sync();   // every CTA has read the old end points of its neighbours
#pragma unroll
        for (int k = 0; k < K; ++k) if (k < cnt) sP[start + k] = Pn[k];
        block_sync<T>();
        exchange_path_halo(sP, sHalo, cl, tid);
        // =================== corridor from the new path (main.cpp:749-756 / 1033-1040) ===================
        // (the reference also rebuilds it after the LAST path update, but nothing reads that corridor: skipped)
        if (outer + 1 == max_outer) continue;
        corridor_stream_c<K>(pt, pv, sB, mbar, bar_phase, sMisc, B.seg, segI0, segO0, segE,
                             C.veh_width_m * 0.5 + C.safety_margin_m, lo, hi, ray_tests);
    }

    // =================== final geometry (main.cpp:761 / 1046) ===================
    block_sync<T>();
    double lap = 0.0;
    {
        double kap[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            kap[k] = 0.0;
            if (k < cnt) {
                const int il = start + k;
                double xp, yp, xpp, ypp;
                derivs_c(pv, il, h, xp, yp, xpp, ypp);
                kap[k] = (xp * ypp - yp * xpp) / pow15(xp * xp + yp * yp);
                B.heading[row0 + il] = atan2(yp, xp);
                B.curvature[row0 + il] = kap[k];
            }
        }
        if (mt) {
            // final v(s) profile (main.cpp:1047)
            double vv[K], axd[K];
            block_sync<T>();
            vprofile_c<K>(pt, cl, q, kap, vv, C.max_vpass_iters, sB, sFlag, fslot, vrounds);
            block_sync<T>();
            lap = lap_and_ax_c<K>(pt, cl, q, vv, axd, sB, sRed + ph * kcRedStride);
            ph ^= 1;
#pragma unroll
            for (int k = 0; k < K; ++k)
                if (k < cnt) { B.v[row0 + start + k] = vv[k]; B.ax[row0 + start + k] = axd[k]; }
        }
    }
    // raceline out: TMA bulk store of the chunk
    fence_proxy_async();
    block_sync<T>();
    if (tid == 0) bulk_s2g(B.xy + 2 * row0, sP, (uint32_t)Nl * 16u);

    // counters (the last cluster barrier also keeps every CTA alive until its neighbours' remote stores are done)
    {
        double rt = (double)ray_tests, dz = 0.0;
        cluster_sum2(rt, dz, sRed + ph * kcRedStride, cl, pt.lane, pt.warp);
        ph ^= 1;
        if (tid == 0 && cl.rank == 0) {
            st->outer_done = max_outer; st->accepted = acc_total; st->backtracks = bt_total; st->evals = ev_total;
            st->vpass_rounds = vrounds; st->ray_tests = (long long)rt; st->lap_time = lap;
        }
    }
}

}  // namespace

inline size_t smem_bytes_cluster(int K) { return (size_t)kcT * K * 48 + kcBytes; }

int launch_solve_cluster(const DevBatch& B, const int* job_list, int n_list, int cs, int mode, void* stream)
{
    constexpr int K = 8;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n_list * cs));
    cfg.blockDim = dim3(kcT);
    cfg.dynamicSmemBytes = smem_bytes_cluster(K);
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e;
    if (mode == 1) e = cudaLaunchKernelEx(&cfg, solve_cluster_kernel<K, 1>, B, job_list, n_list);
    else e = cudaLaunchKernelEx(&cfg, solve_cluster_kernel<K, 0>, B, job_list, n_list);
    return (int)e;
}
int configure_solve_cluster()
{
    const int smem = (int)smem_bytes_cluster(8);
    cudaError_t e = cudaFuncSetAttribute(solve_cluster_kernel<8, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(solve_cluster_kernel<8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    return (int)e;
}

}  // namespace rl
