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
// Closed tracks (is_closed_track = true in every BASELINE config) and, MODE = open, open ones: the one-sided end stencils
// exist in the first / last chunk only.
#pragma once
#include "raceline_kernels.cuh"

namespace rl {
namespace {

constexpr int kMaxCS = 16;       // 2, 4, 8 are portable cluster sizes; 16 needs cudaFuncAttributeNonPortableClusterSizeAllowed
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
constexpr int kcEbar = kcMisc + 64;                   // uint64[2]: the two mbarriers of the per-evaluation exchange (even / odd)
constexpr int kcBytes = 4880;                         // two CTAs (+ their static UpdCtx and the 1 KB reserve each) fill the SM's 228 KB to 112 bytes
static_assert(kcEbar + 16 <= kcBytes && kcEbar % 8 == 0, "cluster scratch layout");

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
__device__ __forceinline__ void packed_warp_sum2(double& a, double& b, int lane)
{
    // one butterfly for both values (block_sum2): after the first exchange lanes 0-15 carry a, lanes 16-31 carry b
    const bool hi_half = (lane & 16) != 0;
    const double keep = hi_half ? b : a, send = hi_half ? a : b;
    double v = keep + __shfl_xor_sync(kFull, send, 16);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    a = __shfl_sync(kFull, v, 0);
    b = __shfl_sync(kFull, v, 16);
}
// sum of the `ne` (<= kMaxCS*NW = 128) partial pairs every CTA holds after an exchange; every lane of every warp of every
// CTA adds the same pairs in the same order and ends with the same bits
__device__ __forceinline__ void gather_pairs(const double* sRedPh, int ne, int lane, double& a, double& b)
{
    double sa = 0.0, sb = 0.0;
    if (lane < ne) { const double2 e = *reinterpret_cast<const double2*>(sRedPh + 2 * lane); sa = e.x; sb = e.y; }
#pragma unroll
    for (int j = 1; j < (kMaxCS * kcNW) / 32; ++j)
        if (lane + 32 * j < ne) { const double2 e = *reinterpret_cast<const double2*>(sRedPh + 2 * (lane + 32 * j)); sa += e.x; sb += e.y; }
    packed_warp_sum2(sa, sb, lane);
    a = sa; b = sb;
}
__device__ __forceinline__ void cluster_sum2(double& a, double& b, double* sRedPh, const Clu& cl, int lane, int warp)
{
    packed_warp_sum2(a, b, lane);
    if (lane < (int)cl.CS) cl_st2(cl_map(sRedPh + 2 * ((int)cl.rank * kcNW + warp), (uint32_t)lane), a, b);
    cl_sync();
    gather_pairs(sRedPh, (int)cl.CS * kcNW, lane, a, b);
}

// ---- the per-evaluation exchange without a cluster barrier -------------------------------------------------------
// barrier.cluster makes all CS*256 threads wait for the slowest CTA AND pays a cluster-scope fence (28 % of the stall
// samples of round 1's kernel).  An evaluation only needs DATA: the CS*NW partial pairs and the two alpha halos of the
// neighbour CTAs.  So every producer sends its 16 bytes with st.async, which writes into the consumer CTA's shared memory
// and completes the bytes on the consumer's own mbarrier in one instruction; a consumer waits on its own mbarrier
// only -- until ITS data is there, not until everybody is done.  Per evaluation and CTA the barrier expects
// (CS*NW + 2)*16 bytes plus one arrival per local warp (the warps' local warp-edge halos go through plain st.shared).
// Two barriers alternate (even / odd evaluation): a neighbour that is one evaluation ahead signals the OTHER barrier,
// and it cannot be two ahead because it needs this CTA's data of the evaluation in between; its bytes may land before
// this CTA has posted its expect_tx (the tx-count goes negative for a moment, which an mbarrier allows; the phase cannot
// complete meanwhile because no local warp has arrived yet).
__device__ __forceinline__ void st_async2(uint32_t remote_addr, double x, double y, uint32_t remote_bar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];" ::"r"(remote_addr), "d"(x), "d"(y),
                 "r"(remote_bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_local(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_local(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAITC_%=:\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONEC_%=;\n"
        "bra WAITC_%=;\n"
        "DONEC_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
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
    bool openL = false, openR = false;   // OPEN track: this chunk starts at the first / ends at the last sample (one-sided there)
    __device__ __forceinline__ double2 at(int il) const { return (il < 0) ? halo[0] : ((il >= Nloc) ? halo[1] : sP[il]); }
};
__device__ __forceinline__ void normal_c(const PathView& pv, int il, double& nx, double& ny)
{
    if (pv.openL && il == 0) { normal_from_tangent(pv.sP[1].x - pv.sP[0].x, pv.sP[1].y - pv.sP[0].y, nx, ny); return; }           // main.cpp:586
    if (pv.openR && il == pv.Nloc - 1) { normal_from_tangent(pv.sP[il].x - pv.sP[il - 1].x, pv.sP[il].y - pv.sP[il - 1].y, nx, ny); return; }
    const double2 Pm = pv.at(il - 1), Pp = pv.at(il + 1);
    normal_from_tangent((Pp.x - Pm.x) * 0.5, (Pp.y - Pm.y) * 0.5, nx, ny);   // main.cpp:584-592
}
__device__ __forceinline__ void derivs_c(const PathView& pv, int il, const HStep& H, double& xp, double& yp, double& xpp, double& ypp)
{
    const double2* sP = pv.sP;
    if (pv.openL && il == 0) {              // the `deriv` lambda at the ends of an open track, main.cpp:604-613 (a chunk has >= 512 samples)
        xp = (sP[1].x - sP[0].x) * H.inv_h; yp = (sP[1].y - sP[0].y) * H.inv_h;
        xpp = (sP[2].x - 2 * sP[1].x + sP[0].x) * H.invh2; ypp = (sP[2].y - 2 * sP[1].y + sP[0].y) * H.invh2;
        return;
    }
    if (pv.openR && il == pv.Nloc - 1) {
        const int n = pv.Nloc;
        xp = (sP[n - 1].x - sP[n - 2].x) * H.inv_h; yp = (sP[n - 1].y - sP[n - 2].y) * H.inv_h;
        xpp = (sP[n - 1].x - 2 * sP[n - 2].x + sP[n - 3].x) * H.invh2; ypp = (sP[n - 1].y - 2 * sP[n - 2].y + sP[n - 3].y) * H.invh2;
        return;
    }
    derivs_central(pv.at(il - 1), pv.at(il), pv.at(il + 1), H, xp, yp, xpp, ypp);
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
    uint64_t* ebar;     // the two exchange mbarriers of this CTA
    uint32_t epar;      // their wait parities (bit b: barrier b): a register copy, handed back to the kernel by pgd_outer_c
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
    const double dec2p = project_trial<T, K>(xa, gh, c.step2, sLo, sHi, xb);
    double dec = c.decp;
#ifdef RL_CLUSTER_BARRIER_EXCHANGE      // round 1's exchange (one barrier.cluster per evaluation), kept for A/B measurements
    hb = halo_send_c<K, MODE>(xb, pt, cl, c.sExF + c.ph * kcExStride, c.sExL + c.ph * kcExStride);
    cluster_sum2(Jn, dec, c.sRed + c.ph * kcRedStride, cl, pt.lane, pt.warp);
    halo_recv_c(hb, pt, c.sExF + c.ph * kcExStride, c.sExL + c.ph * kcExStride);
#else
    {
        // ---- exchange of this evaluation: halos of the next trial + partial sums, signalled through mbarrier c.ph ----
        double* sExF = c.sExF + c.ph * kcExStride; double* sExL = c.sExL + c.ph * kcExStride; double* sRedPh = c.sRed + c.ph * kcRedStride;
        uint64_t* bar = c.ebar + c.ph;
        double F0, F1, L0, L1;
        edge_values<K, MODE>(xb, pt.cnt, F0, F1, L0, L1);
        hb.l0 = __shfl_sync(kFull, L0, pt.srcL);
        hb.l1 = __shfl_sync(kFull, L1, pt.srcL);
        hb.r0 = __shfl_sync(kFull, F0, pt.srcR);
        hb.r1 = __shfl_sync(kFull, F1, pt.srcR);
        if (pt.lane == 0) {
            if (pt.warp == 0) st_async2(cl_map(sExF + 2 * kcNW, cl.left), F0, F1, cl_map(bar, cl.left));     // I am the right neighbour of CTA `left`
            else { sExF[2 * pt.warp] = F0; sExF[2 * pt.warp + 1] = F1; }
        }
        if (pt.lane == 31) {
            if (pt.warp == kcNW - 1) st_async2(cl_map(sExL + 2 * kcNW, cl.right), L0, L1, cl_map(bar, cl.right));
            else { sExL[2 * pt.warp] = L0; sExL[2 * pt.warp + 1] = L1; }
        }
        packed_warp_sum2(Jn, dec, pt.lane);
        if (pt.lane < (int)cl.CS)
            st_async2(cl_map(sRedPh + 2 * ((int)cl.rank * kcNW + pt.warp), (uint32_t)pt.lane), Jn, dec, cl_map(bar, (uint32_t)pt.lane));
        __syncwarp();
        if (pt.lane == 0) {
            if (pt.warp == 0) mbar_arrive_expect_local(bar, (uint32_t)((int)cl.CS * kcNW + 2) * 16u);
            else mbar_arrive_local(bar);
        }
        mbar_wait_cluster(bar, (c.epar >> c.ph) & 1u);
        c.epar ^= (1u << c.ph);
        gather_pairs(sRedPh, (int)cl.CS * kcNW, pt.lane, Jn, dec);
        halo_recv_c(hb, pt, sExF, sExL);
    }
#endif
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
                                              double lamJ, double step_init, double step_min, double armijo_c, int max_inner,
                                              uint64_t* ebar, uint32_t& epar)
{
    constexpr int T = kcT;
    PgdOut o; o.acc = 0; o.bt = 0; o.ev = 0;
    PgdCtxC<K> c;
    c.sC0 = sC0; c.sCp = sCp; c.sCm = sCm; c.sSt = sSt; c.sRed = sRed; c.sExF = sExF; c.sExL = sExL;
    c.lamJ = lamJ; c.armijo_c = armijo_c; c.ph = ph;
    c.ebar = ebar; c.epar = epar;
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
            double lo, hi;
            ld_pair<T>(sLo, sHi, k, lo, hi);
            const double xn = clamp_box(-c.step2 * gh[k], lo, hi);
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
                double lo, hi;
                ld_pair<T>(sLo, sHi, k, lo, hi);
                const double xn = clamp_box(fma(-c.step2, gh[k], a[k]), lo, hi);
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
    epar = c.epar;
    o.Jend = c.J;
    return o;
}

// ---- v(s) profile on a cluster: vprofile_blocked with the end values of each CTA published to its neighbour ----
// sX: 6*(T+1) doubles.  Index T of each array is the neighbour CTA's edge thread.
template <int K>
__device__ __forceinline__ void vprofile_c(const Part& pt, const Clu& cl, const VPar& q, const double (&kap)[K], double (&v)[K],
                                           int max_iters, double* sX, int* sFlag, int& fslot, int& rounds, bool closed = true)
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
        double v0[K];
        bool chg_iter = false;   // as in vprofile_blocked: v only decreases, so "changed" is gathered from the two sweeps
#pragma unroll
        for (int k = 0; k < K; ++k) v0[k] = v[k];
        // ---------------- forward sweep (main.cpp:829-833) ----------------
        {
            double last = v[0];
#pragma unroll
            for (int k = 1; k < K; ++k) if (k < cnt) last = v[k];
            sVL[b * S + tid] = last;
            if (tid == T - 1) cl_st1(cl_map(&sVL[b * S + T], cl.right), last);
        }
        cl_sync();
        // (as in vprofile_blocked: a chunk whose input did not change since the previous round is not recomputed)
        bool first_round = true;
        double vin_prev = 0.0, u = 0.0;
        for (;;) {
            const double vin = sVL[b * S + iL];
            if (first_round || vin != vin_prev) {
                u = gfirst ? v0[0] : fmin(v0[0], f_acc(q, vin, kapL));
                v[0] = u;
#pragma unroll
                for (int k = 1; k < K; ++k)
                    if (k < cnt) { u = fmin(v0[k], f_acc(q, u, kap[k - 1])); v[k] = u; }
                vin_prev = vin;
            }
            first_round = false;
            const bool changed = (u != sVL[b * S + tid]);
            sVL[(b ^ 1) * S + tid] = u;
            if (tid == T - 1) cl_st1(cl_map(&sVL[(b ^ 1) * S + T], cl.right), u);
            b ^= 1;
            ++rounds;
            if (!cluster_or(changed, sFlag, fslot, cl, tid, pt.lane)) break;
        }
        // closed-loop wrap: v[0] = min(v[0], f_acc(v[N-1])), main.cpp:834-839
        if (gfirst && closed) v[0] = fmin(v[0], f_acc(q, sVL[b * S + iL], kapL));
#pragma unroll
        for (int k = 0; k < K; ++k) if (k < cnt && v[k] != v0[k]) chg_iter = true;
        // ---------------- backward sweep (main.cpp:841-845) ----------------
#pragma unroll
        for (int k = 0; k < K; ++k) v0[k] = v[k];
        sVF[b * S + tid] = v[0];
        if (tid == 0) cl_st1(cl_map(&sVF[b * S + T], cl.left), v[0]);
        cl_sync();
        first_round = true;
        for (;;) {
            const double vin = sVF[b * S + iR];
            if (first_round || vin != vin_prev) {
                u = 0.0;
#pragma unroll
                for (int k = K - 1; k >= 0; --k) {
                    if (k == cnt - 1) { u = glast ? v0[k] : fmin(v0[k], f_brk(q, vin, kapR)); v[k] = u; }
                    else if (k < cnt - 1) { u = fmin(v0[k], f_brk(q, u, kap[k + 1])); v[k] = u; }
                }
                vin_prev = vin;
            }
            first_round = false;
            const bool changed = (u != sVF[b * S + tid]);
            sVF[(b ^ 1) * S + tid] = u;
            if (tid == 0) cl_st1(cl_map(&sVF[(b ^ 1) * S + T], cl.left), u);
            b ^= 1;
            ++rounds;
            if (!cluster_or(changed, sFlag, fslot, cl, tid, pt.lane)) break;
        }
        // closed-loop wrap: v[N-1] = min(v[N-1], f_brk(v[0])), main.cpp:846-850
        if (glast && closed) {
            const double w = f_brk(q, sVF[b * S + iR], kapR);
#pragma unroll
            for (int k = 0; k < K; ++k) if (k == cnt - 1) v[k] = fmin(v[k], w);
        }
#pragma unroll
        for (int k = 0; k < K; ++k) if (k < cnt && v[k] != v0[k]) chg_iter = true;
        if (!cluster_or(chg_iter, sFlag, fslot, cl, tid, pt.lane)) break;
    }
}

// ax and lap time, main.cpp:854-860 (closed track: the last sample's successor is sample 0)
template <int K>
__device__ __forceinline__ double lap_and_ax_c(const Part& pt, const Clu& cl, const VPar& q, const double (&v)[K], double (&ax)[K],
                                               double* sX, double* sRedPh, bool closed = true)
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
            double v1 = (!closed && cl.rank == cl.CS - 1 && pt.tid == T - 1) ? v0 : vnext_edge;   // open: j = i at the last sample (main.cpp:856)
            if (k + 1 < K) { if (k + 1 < pt.cnt) v1 = v[k + 1]; }
            ax[k] = (v1 * v1 - v0 * v0) / (2.0 * q.h);
            t += q.h / fmax(1e-6, v0);
        }
    }
    cluster_sum2(t, dummy, sRedPh, cl, pt.lane, pt.warp);
    return t;
}

// ---- corridor of one chunk ------------------------------------------------------------------------------------
// Two paths, as in solve_kernel (corridor_build_fast / corridor_update), adapted to rings that do not fit shared
// memory:
//   corridor_search_c   the SEARCHING path: each ring streamed through region B in tiles (TMA bulk copies, FP32 box
//                       hierarchy per tile); exact nearest +n / -n hits and the exact point-ring distance over ALL
//                       tiles (main.cpp:478-512, 694-711), and what later builds live on: the anchor segment of the
//                       nearest hit, the clearance of everything outside its window, and an existence certificate
//                       (FAR segment / ring-free CONE) for the ray the point-ring fallback (main.cpp:696) depends on.
//                       First build: all samples.  Later: only the samples the update path flagged.
//   corridor_update_c   the UPDATE path: the anchors of a chunk's samples span a short range of each ring, so only
//                       that range of vertices (FP64 + FP32, 24 B each) is loaded; windows + certificates settle the
//                       sample exactly like corridor_update_sample does for a whole ring (same code, LOCAL form).
// Both run per CTA on its own samples with no cluster traffic.

// ray_scan with the nearest hit's SEGMENT per direction (tile-local; untouched when this tile does not improve it).
// Both rays; pruned by this ring's own best hits so far (pos / neg), so INF afterwards means: no hit on this ring at all.
// (arguments and results by value: see ray_scan_v)
struct RaySegRes { double pos, neg; int sp, sn, tests; };
__device__ __noinline__ RaySegRes ray_scan_seg_v(const RayTile tl_in, double2 P, double nx, double ny, float px, float py, float m,
                                                 double bp, double bn)
{
    const RayTile tl = tile_in_smem(tl_in);
    const double INF = dinf();
    const float FINF = __int_as_float(0x7f800000);
    const float fnx = (float)nx, fny = (float)ny, anx = fabsf(fnx), any = fabsf(fny);
    int sp = -1, sn = -1, tests = 0;
    float bpf = (bp < INF) ? __double2float_ru(bp) : FINF, bnf = (bn < INF) ? __double2float_ru(bn) : FINF;
    for (int sb = 0; sb < tl.nsup; ++sb) {
        if (!ray_box(tl.supF[sb], px, py, fnx, fny, anx, any, m, bpf, bnf, true, true)) continue;
        const int b1 = min(tl.nblk, sb * SU + SU);
        for (int b = sb * SU; b < b1; ++b) {
            if (!ray_box(tl.boxF[b], px, py, fnx, fny, anx, any, m, bpf, bnf, true, true)) continue;
            const int s1 = min(tl.nt, b * SB + SB);
            for (int s = b * SB; s < s1; ++s) {
                const float4 f = tl.segF[s];
                const float sa = fnx * (f.y - py) - fny * (f.x - px), sbb = fnx * (f.w - py) - fny * (f.z - px);
                if (fminf(sa, sbb) > m || fmaxf(sa, sbb) < -m) continue;
                const double x0 = tl.segD[4 * s], y0 = tl.segD[4 * s + 1], vx = tl.segD[4 * s + 2], vy = tl.segD[4 * s + 3];
                const double den = nx * (-vy) + ny * vx;                         // main.cpp:483
                ++tests;
                if (fabs(den) < 1e-15) continue;                                 // main.cpp:484
                const double ax = x0 - P.x, ay = y0 - P.y;                      // main.cpp:485
                const double inv = 1.0 / den;
                const double t = (ax * (-vy) + ay * vx) * inv;                  // main.cpp:486
                const double u = (nx * ay - ny * ax) * inv;                     // main.cpp:487
                if (u >= -1e-12 && u <= 1.0 + 1e-12) {                          // main.cpp:488
                    if (t > 0.0) { if (t < bp) { bp = t; bpf = __double2float_ru(t); sp = s; } }         // +n ray, main.cpp:497
                    else if (t < 0.0) { if (-t < bn) { bn = -t; bnf = __double2float_ru(-t); sn = s; } }  // -n ray: t' = -t
                }
            }
        }
    }
    RaySegRes r;
    r.pos = bp; r.neg = bn; r.sp = sp; r.sn = sn; r.tests = tests;
    return r;
}
__device__ __forceinline__ void ray_scan_seg(const RayTile& tl, double2 P, double nx, double ny, float px, float py, float m,
                                             double& pos_io, double& neg_io, int& segp_io, int& segn_io, long long& tests_io)
{
    const RaySegRes r = ray_scan_seg_v(tl, P, nx, ny, px, py, m, pos_io, neg_io);
    pos_io = r.pos; neg_io = r.neg; tests_io += r.tests;
    if (r.sp >= 0) segp_io = r.sp;
    if (r.sn >= 0) segn_io = r.sn;
}

// crossings (mod 2) of the +x ray from P with the segments of one TILE of a vertex chain; the end vertex of the
// tile's last segment is the start of the ring's next segment (next_start), shared bit for bit like every other vertex
__device__ __noinline__ bool inside_ring_tile(const RayTile tl_in, double2 P, float px, float py, float m, double2 next_start)
{
    const RayTile tl = tile_in_smem(tl_in);
    int cnt = 0;
    for (int sb = 0; sb < tl.nsup; ++sb) {
        const float4 sx = tl.supF[sb];
        if (fabsf(sx.y - py) > sx.w + m || px > sx.x + sx.z + m) continue;
        const int b1 = min(tl.nblk, sb * SU + SU);
        for (int b = sb * SU; b < b1; ++b) {
            const float4 bx = tl.boxF[b];
            if (fabsf(bx.y - py) > bx.w + m || px > bx.x + bx.z + m) continue;
            const int s1 = min(tl.nt, b * SB + SB);
            for (int s = b * SB; s < s1; ++s) {
                const double ax = tl.segD[4 * s], ay = tl.segD[4 * s + 1];
                const double bxx = (s + 1 == tl.nt) ? next_start.x : tl.segD[4 * (s + 1)];
                const double byy = (s + 1 == tl.nt) ? next_start.y : tl.segD[4 * (s + 1) + 1];
                if ((ay > P.y) != (byy > P.y)) {
                    const double xc = ax + (P.y - ay) * (bxx - ax) / (byy - ay);
                    if (xc > P.x) ++cnt;
                }
            }
        }
    }
    return (cnt & 1) != 0;
}

struct AccC {   // one sample's accumulators while the rings stream by (local memory: the sample loops are not unrolled)
    double nx, ny, pos[2], neg[2], dist[2];
    int sp[2], sn[2];
    float clr[2], px, py;
    unsigned par;     // bit r: odd number of crossings with ring r so far (first build only)
};

template <int K>
__device__ __forceinline__ void corridor_search_c(const Part& pt, const PathView& pv, double* sB, uint64_t* mbar, uint32_t& bar_phase,
                                                  int* sMisc, unsigned* sHint, unsigned short* sClr,
                                                  const double* __restrict__ gseg, const double* __restrict__ gcenter,
                                                  unsigned long long* __restrict__ gcert, unsigned long long* __restrict__ gapex,
                                                  long long segI0, long long segO0, long long segE, double guard, unsigned mask,
                                                  bool first, bool parity_ok,
                                                  double (&loc)[K], double (&hic)[K], long long& ray_tests, int& ex_scans)
{
    // first: the build from the centre line (every sample): also settles, once per job, whether each ring is a closed
    // vertex chain and which samples lie inside it (crossing parity) -- a ray from inside a closed ring always hits it
    constexpr int T = kcT, NP = kcT * K;
    constexpr int CAP = fast_tile_cap(NP);
    const int Nl = pv.Nloc, tid = pt.tid;
    const double INF = dinf();
    const double2 org = pv.sP[0];
    AccC acc[K];
#pragma unroll 1
    for (int j = 0; j < K; ++j) {
        const int i = tid + j * T;
        if (i >= Nl || !((mask >> j) & 1u)) continue;
        AccC& a = acc[j];
        normal_c(pv, i, a.nx, a.ny);
        const double2 Pc = pv.sP[i];
        a.px = (float)(Pc.x - org.x); a.py = (float)(Pc.y - org.y);
        for (int r = 0; r < 2; ++r) { a.pos[r] = INF; a.neg[r] = INF; a.dist[r] = INF; a.sp[r] = -1; a.sn[r] = -1; a.clr[r] = 3e18f; }
        a.par = 0u;
    }
    int chain_ok[2] = {1, 1};
    float mring[2] = {0.f, 0.f};
    for (int ring = 0; ring < 2; ++ring) {
        const long long base = ring ? segO0 : segI0;
        const int mr = (int)(ring ? (segE - segO0) : (segO0 - segI0));
        if (mr == 0) continue;   // safe_ray on an empty ring returns 0 (main.cpp:696-697): handled in the combine step
        const int ntiles = (mr + CAP - 1) / CAP;
        RayTile tl;
        float m0 = 0.f;
        for (int pass = 0; pass < 2; ++pass) {
            for (int tile = 0; tile < ntiles; ++tile) {
                const int t0 = tile * CAP, nt = min(CAP, mr - t0);
                if (pass == 0 || ntiles > 1) {
                    bool chain, linked;
                    m0 = ring_tile_build<T, K>(pt, sB, mbar, bar_phase, sMisc, gseg + 4 * (base + t0), nt, org.x, org.y, tl, chain, &linked);
                    mring[ring] = fmaxf(mring[ring], m0);
                    if (first && pass == 0 && !linked) chain_ok[ring] = 0;
                }
                double2 next_start = make_double2(0.0, 0.0);
                if (first && pass == 0) {
                    // the vertex after this tile (the ring's first vertex after the last tile) must equal the tile's last end point
                    const int nxt = (t0 + nt == mr) ? 0 : t0 + nt;
                    next_start = *reinterpret_cast<const double2*>(gseg + 4 * (base + nxt));
                    const double2 last_end = *reinterpret_cast<const double2*>(gseg + 4 * (base + t0 + nt - 1) + 2);
                    if (!(next_start.x == last_end.x && next_start.y == last_end.y)) chain_ok[ring] = 0;
                }
#pragma unroll 1
                for (int j = 0; j < K; ++j) {
                    const int i = tid + j * T;
                    if (i >= Nl || !((mask >> j) & 1u)) continue;
                    AccC& a = acc[j];
                    const double2 Pc = pv.sP[i];
                    const float m = m0 + 2e-6f * fmaxf(fabsf(a.px), fabsf(a.py));
                    if (pass == 0) {
                        // exact nearest hits of both rays on this ring, and their segments (ring-global)
                        int sp = -1, sn = -1;
                        ray_scan_seg(tl, Pc, a.nx, a.ny, a.px, a.py, m, a.pos[ring], a.neg[ring], sp, sn, ray_tests);
                        if (sp >= 0) a.sp[ring] = t0 + sp;
                        if (sn >= 0) a.sn[ring] = t0 + sn;
                        if (first && inside_ring_tile(tl, Pc, a.px, a.py, m, next_start)) a.par ^= (1u << ring);
                    } else {
                        // everything outside the anchor's window is at least `clr` away; exact point-ring distance
                        const int anchor = (a.pos[ring] <= a.neg[ring]) ? a.sp[ring] : a.sn[ring];
                        if (anchor >= 0 && mr <= 8191) a.clr[ring] = fminf(a.clr[ring], clearance_scan(tl, a.px, a.py, m, anchor, t0, mr));
                        if (a.pos[ring] == INF || a.neg[ring] == INF) {     // only a ring some ray misses needs it (main.cpp:696)
                            const double d = dist_scan(tl, Pc, a.px, a.py, m, 0, fmin(a.dist[ring], fmin(a.pos[ring], a.neg[ring])));
                            a.dist[ring] = fmin(a.dist[ring], d);
                        }
                    }
                }
            }
        }
    }
    if (tid == 0) {   // the FP32 margin later updates must allow for: the largest any tile was built with
        if (__int_as_float(sMisc[10]) < mring[0]) sMisc[10] = __float_as_int(mring[0]);
        if (__int_as_float(sMisc[11]) < mring[1]) sMisc[11] = __float_as_int(mring[1]);
        if (first) { sMisc[8] = chain_ok[0]; sMisc[9] = chain_ok[1]; }     // bit 0: closed vertex chain (every thread computed the same)
    }
    const int rfl[2] = {first ? chain_ok[0] : (sMisc[8] & 1), first ? chain_ok[1] : (sMisc[9] & 1)};
    // ---- combine (main.cpp:696-710), leave the per-sample state, decide which existence certificate a sample needs ----
    const int M0 = (int)(segO0 - segI0), M1 = (int)(segE - segO0);
    unsigned need_cone = 0u;   // bit j: sample j needs a CONE certificate; coneKey[j] = ring | dir << 1
    unsigned coneKey = 0u;     // 2 bits per sample
#pragma unroll 1
    for (int j = 0; j < K; ++j) {
        const int i = tid + j * T;
        if (i >= Nl || !((mask >> j) & 1u)) continue;
        AccC& a = acc[j];
        double dpos = INF, dneg = INF;
        for (int r = 0; r < 2; ++r) {
            const int mr = r ? M1 : M0;
            double vp, vn;
            if (mr == 0) { vp = 0.0; vn = 0.0; }
            else {
                const double dist = (a.dist[r] < INF) ? a.dist[r] : 0.0;
                vp = (a.pos[r] < INF) ? a.pos[r] : dist;
                vn = (a.neg[r] < INF) ? a.neg[r] : dist;
            }
            dpos = fmin(dpos, fmax(0.0, vp));
            dneg = fmin(dneg, fmax(0.0, vn));
        }
        double hv = fmax(0.0, dpos - guard), lv = -fmax(0.0, dneg - guard);
        if (!isfinite(hv)) hv = 0.0;
        if (!isfinite(lv)) lv = 0.0;
        hic[j] = hv; loc[j] = lv;
        // state: anchors, clearances (relative to the centre-line sample, 1/4 m units, rounded down)
        const double2 Pc = pv.sP[i];
        const double cx0 = gcenter[2 * i], cy0 = gcenter[2 * i + 1];
        const float disp = __double2float_ru(sqrt((Pc.x - cx0) * (Pc.x - cx0) + (Pc.y - cy0) * (Pc.y - cy0))) * (1.f + 1e-6f);
        // parity bits: taken now (first build) or kept
        unsigned hw = first ? (((a.par & 1u) && rfl[0] ? 1u : 0u) << 26) | (((a.par & 2u) && rfl[1] ? 1u : 0u) << 27) : (sHint[i] & (3u << 26));
        unsigned cw = 0u;
        float rcv[2] = {0.f, 0.f};
        for (int r = 0; r < 2; ++r) {
            const int mr = r ? M1 : M0;
            const int anchor = (a.pos[r] <= a.neg[r]) ? a.sp[r] : a.sn[r];
            if (anchor >= 0 && mr > 2 * kWin + 1 && mr <= 8191) {
                const float cval = a.clr[r] - disp;
                const unsigned cq = (cval >= 63.75f) ? 255u : (cval > 0.f ? (unsigned)(cval * 4.f) : 0u);
                hw |= ((unsigned)anchor << (13 * r)) | (1u << (28 + r));
                cw |= cq << (8 * r);
                rcv[r] = 0.25f * (float)cq - disp - 4.f * (mring[r] * 1.001f + 2e-6f * fmaxf(fabsf(a.px), fabsf(a.py)) + 1e-5f);
            }
        }
        sHint[i] = hw; sClr[i] = (unsigned short)cw;
        // the certificate the update path will ask for: a ring without a certified near hit in some direction whose
        // point distance undercuts the other ring's hit there (see corridor_update_sample)
        for (int dir = 0; dir < 2; ++dir) {
            for (int r = 0; r < 2; ++r) {
                const double hit = dir ? a.neg[r] : a.pos[r];
                if (hit <= (double)rcv[r]) continue;                       // a near hit: no question to answer
                if (parity_ok && rfl[r] && ((hw >> (26 + r)) & 1u)) continue;  // inside a closed ring: every ray hits it
                // (asked for only while the point distance undercuts the other ring's hit, but that flips as the path
                //  moves and a search costs a pass over every tile: the certificate is made now, whichever way it stands)
                const unsigned key = ((unsigned)r << 2) | ((unsigned)dir << 3);
                if (hit < INF) {   // it hits far away: remember where
                    gcert[i] = ((unsigned long long)(unsigned)(dir ? a.sn[r] : a.sp[r]) << 32) | (unsigned long long)(key | kCertFar);
                } else { need_cone |= (1u << j); coneKey = (coneKey & ~(3u << (2 * j))) | (((unsigned)r | ((unsigned)dir << 1)) << (2 * j)); }
            }
        }
    }
    // ---- ring-free cones for the rays that miss (cone_scan over every tile of that ring) ----
    if (block_or<T>(need_cone != 0u)) {
        for (int ring = 0; ring < 2; ++ring) {
            bool mine = false;
#pragma unroll 1
            for (int j = 0; j < K; ++j) mine = mine || (((need_cone >> j) & 1u) && (int)((coneKey >> (2 * j)) & 1u) == ring);
            if (!block_or<T>(mine)) continue;
            const long long base = ring ? segO0 : segI0;
            const int mr = ring ? M1 : M0;
            const int ntiles = (mr + CAP - 1) / CAP;
            float cb[K];
#pragma unroll 1
            for (int j = 0; j < K; ++j) cb[j] = -1.f;
            for (int tile = 0; tile < ntiles; ++tile) {
                const int t0 = tile * CAP, nt = min(CAP, mr - t0);
                RayTile tl;
                bool chain;
                const float m0 = ring_tile_build<T, K>(pt, sB, mbar, bar_phase, sMisc, gseg + 4 * (base + t0), nt, org.x, org.y, tl, chain);
#pragma unroll 1
                for (int j = 0; j < K; ++j) {
                    const int i = tid + j * T;
                    if (!((need_cone >> j) & 1u) || (int)((coneKey >> (2 * j)) & 1u) != ring) continue;
                    const AccC& a = acc[j];
                    const int dir = (int)((coneKey >> (2 * j + 1)) & 1u);
                    const double sg = dir ? -1.0 : 1.0;
                    const int qx = __double2int_rn(sg * a.nx * 32767.0), qy = __double2int_rn(sg * a.ny * 32767.0);
                    double d0x, d0y;
                    cert_axis(((unsigned)qx & 0xffffu) | ((unsigned)qy << 16), d0x, d0y);
                    const double toward = dir ? a.pos[ring] : a.neg[ring];     // this ring's hit on the opposite ray
                    const double back = (toward < INF) ? fmax(0.0, toward - fmax(0.05, 0.5 * guard)) : 0.0;
                    const double2 Pc = pv.sP[i];
                    const double cx0 = gcenter[2 * i], cy0 = gcenter[2 * i + 1];
                    const float oxf = (float)((Pc.x - cx0) - back * d0x), oyf = (float)((Pc.y - cy0) - back * d0y);
                    const double axd = cx0 + (double)oxf, ayd = cy0 + (double)oyf;
                    const float m = m0 + 2e-6f * fmaxf(fabsf(a.px), fabsf(a.py));
                    cb[j] = fmaxf(cb[j], cone_scan(tl, (float)(axd - org.x), (float)(ayd - org.y), (float)d0x, (float)d0y, m, 0));
                }
            }
#pragma unroll 1
            for (int j = 0; j < K; ++j) {
                const int i = tid + j * T;
                if (!((need_cone >> j) & 1u) || (int)((coneKey >> (2 * j)) & 1u) != ring) continue;
                const AccC& a = acc[j];
                const int dir = (int)((coneKey >> (2 * j + 1)) & 1u);
                const double sg = dir ? -1.0 : 1.0;
                const unsigned key = ((unsigned)ring << 2) | ((unsigned)dir << 3);
                const int qx = __double2int_rn(sg * a.nx * 32767.0), qy = __double2int_rn(sg * a.ny * 32767.0);
                const unsigned nw1 = ((unsigned)qx & 0xffffu) | ((unsigned)qy << 16);
                double d0x, d0y;
                cert_axis(nw1, d0x, d0y);
                const double toward = dir ? a.pos[ring] : a.neg[ring];
                const double back = (toward < INF) ? fmax(0.0, toward - fmax(0.05, 0.5 * guard)) : 0.0;
                const double2 Pc = pv.sP[i];
                const double cx0 = gcenter[2 * i], cy0 = gcenter[2 * i + 1];
                const float oxf = (float)((Pc.x - cx0) - back * d0x), oyf = (float)((Pc.y - cy0) - back * d0y);
                unsigned nw0 = key | kCertNoCone;
                if (cb[j] < 0.9995f) {
                    const unsigned qc = (unsigned)ceilf((cb[j] + 1.f) * 32767.f + 0.5f);
                    if (qc < 65535u) nw0 = key | kCertCone | (qc << 16);
                }
                gcert[i] = ((unsigned long long)nw1 << 32) | (unsigned long long)nw0;
                gapex[i] = ((unsigned long long)__float_as_uint(oyf) << 32) | (unsigned long long)__float_as_uint(oxf);
                ++ex_scans;
            }
        }
    }
}

// ---- the searching path for a FEW flagged samples ------------------------------------------------------------------
// After the first build the update path flags a handful of samples per chunk and outer iteration (round-2 timers: about
// 100 per job, all because an existence certificate no longer applies), and corridor_search_c answers them by streaming
// every tile of both rings through shared memory again: about 24 tile builds = 500 k cycles for 4 samples, during which
// the other CTAs of the cluster wait at the next cluster barrier (30 % of a job).  Here the whole CTA works on ONE flagged
// sample at a time straight from global memory (the rings are L2-resident): every thread takes every T-th segment of a
// ring, runs the same FP32 side test and the same exact FP64 formulas (main.cpp:478-512) as ray_scan_seg / dist_scan /
// clearance_scan / cone_scan, and block-wide minima replace the scan order (they do not depend on it: hits, distances
// and clearances are minima over the ring; equal hits keep the lower segment index).  Results, per-sample state and
// certificates are those of corridor_search_c(first = false).
constexpr int kFewMax = 24;      // more flagged samples than this in one chunk: the tile-streaming search is cheaper

// Every job leaves an exact bounding box (xmin, ymin, xmax, ymax) per GROUP of 32 consecutive segments of each ring in
// global memory (few_boxes_build: the job's ax rows, which nothing else touches before the final profile).  A pass over a
// ring first tests the boxes (one per thread) and lists the groups that can matter -- the ray's line crosses the box, the
// box is within the radius a distance or clearance question reaches, the box is not wholly outside the 20-degree cone --
// and then one warp takes one listed group at a time, one lane per segment.  About 30 of the 233 groups of a 7447-segment
// ring survive; the culled ones provably cannot change a minimum (the tests are exact FP64 box geometry with slack).
// BODY sees a4[u], b4[u] (end points), sg4[u] (index in the ring), ok4[u] (false: past the end of the ring) for
// u < kFewW; PRED is an expression over the box `bx`.
constexpr int kFewW = 1;
constexpr int kFewMaxGroups = 1024;        // listed groups fit 4 KB of region B: rings of up to 32,768 segments
#define RL_FEW_FOREACH(gs_, mr_, gbox_, PRED, BODY)                                                                      \
    if ((mr_) > 0) {                                                                                                     \
        const int ng__ = ((mr_) + 31) >> 5;                                                                              \
        __syncthreads();               /* the list of the previous pass has been consumed */                             \
        if (tid == 0) sMisc[0] = 0;                                                                                      \
        __syncthreads();                                                                                                 \
        for (int g__ = tid; g__ < ng__; g__ += kcT) {                                                                    \
            const double2 lo__ = __ldcg(reinterpret_cast<const double2*>((gbox_) + 4 * (size_t)g__));                    \
            const double2 hi__ = __ldcg(reinterpret_cast<const double2*>((gbox_) + 4 * (size_t)g__ + 2));                \
            const double4 bx = make_double4(lo__.x, lo__.y, hi__.x, hi__.y);                                             \
            if (PRED) sCand[atomicAdd(&sMisc[0], 1)] = g__;                                                              \
        }                                                                                                                \
        __syncthreads();                                                                                                 \
        const int nc__ = sMisc[0];                                                                                       \
        for (int ci__ = warp; ci__ < nc__; ci__ += kcNW) {                                                               \
            const int g__ = sCand[ci__];                                                                                 \
            double2 a4[kFewW], b4[kFewW]; int sg4[kFewW]; bool ok4[kFewW];                                               \
            sg4[0] = 32 * g__ + lane; ok4[0] = sg4[0] < (mr_);                                                           \
            const size_t si__ = (size_t)(ok4[0] ? sg4[0] : 32 * g__);                                                    \
            a4[0] = __ldg(reinterpret_cast<const double2*>(gs_) + 2 * si__);                                             \
            b4[0] = __ldg(reinterpret_cast<const double2*>(gs_) + 2 * si__ + 1);                                         \
            { BODY }                                                                                                     \
        }                                                                                                                \
    }
// box tests (exact FP64 geometry; the slack covers the rounding of the few operations and the 1e-12 tolerance of main.cpp:488)
__device__ __forceinline__ bool few_box_line(const double4 bx, double2 P, double nx, double ny)
{
    const double cx = 0.5 * (bx.x + bx.z) - P.x, cy = 0.5 * (bx.y + bx.w) - P.y, hx = 0.5 * (bx.z - bx.x), hy = 0.5 * (bx.w - bx.y);
    return fabs(nx * cy - ny * cx) <= (fabs(nx) * hy + fabs(ny) * hx) * (1.0 + 1e-9) + 1e-9 * (fabs(cx) + fabs(cy) + 1.0);
}
__device__ __forceinline__ bool few_box_near(const double4 bx, double2 P, double R)
{
    const double dx = fmax(0.0, fmax(bx.x - P.x, P.x - bx.z)), dy = fmax(0.0, fmax(bx.y - P.y, P.y - bx.w));
    return !(dx * dx + dy * dy > R * R * (1.0 + 1e-9) + 1e-9);      // R = INF: every box
}
// false: no point of the box is seen from the apex within 20 degrees of the axis (cone_scan never looks wider)
__device__ __forceinline__ bool few_box_cone(const double4 bx, double ax, double ay, double dx, double dy)
{
    const double cx = 0.5 * (bx.x + bx.z) - ax, cy = 0.5 * (bx.y + bx.w) - ay, hx = 0.5 * (bx.z - bx.x), hy = 0.5 * (bx.w - bx.y);
    const double R2 = (hx * hx + hy * hy) * (1.0 + 1e-6) + 1e-6, r2 = cx * cx + cy * cy;
    if (r2 <= R2 * 1.001) return true;                         // apex inside (or at) the box's disk
    const double ir = 1.0 / sqrt(r2);
    const double ct = (cx * dx + cy * dy) * ir, so = sqrt(R2) * ir;
    const double co = sqrt(fmax(0.0, 1.0 - so * so));
    if (ct >= co - 1e-6) return true;                          // the axis passes through (or close to) the disk
    const double st = sqrt(fmax(0.0, 1.0 - ct * ct));
    return ct * co + st * so + 1e-5 > 0.9396926;              // cos(theta_c - omega) against cos(20 degrees), with slack
}
// the boxes of one job: groups are dealt to the warps of the whole cluster; a cluster barrier must follow before any read
__device__ __forceinline__ void few_boxes_build(const Clu& cl, int tid, const double* __restrict__ gseg, long long segI0, long long segO0,
                                                long long segE, double* __restrict__ gbox)
{
    const int lane = tid & 31, gw = (int)cl.rank * kcNW + (tid >> 5), nw = (int)cl.CS * kcNW;
    const double INF = dinf();
#pragma unroll 1
    for (int r = 0; r < 2; ++r) {
        const double* gs = gseg + 4 * (r ? segO0 : segI0);
        const int mr = (int)(r ? (segE - segO0) : (segO0 - segI0));
        const int ng = (mr + 31) >> 5;
        double* gb = gbox + (r ? 4 * (size_t)(((int)(segO0 - segI0) + 31) >> 5) : 0);
        for (int g = gw; g < ng; g += nw) {
            const int sg = 32 * g + lane;
            double xmin = INF, ymin = INF, xmax = -INF, ymax = -INF;
            if (sg < mr) {
                const double2 a = __ldg(reinterpret_cast<const double2*>(gs) + 2 * (size_t)sg), b = __ldg(reinterpret_cast<const double2*>(gs) + 2 * (size_t)sg + 1);
                xmin = fmin(a.x, b.x); xmax = fmax(a.x, b.x); ymin = fmin(a.y, b.y); ymax = fmax(a.y, b.y);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                xmin = fmin(xmin, __shfl_xor_sync(kFull, xmin, o)); ymin = fmin(ymin, __shfl_xor_sync(kFull, ymin, o));
                xmax = fmax(xmax, __shfl_xor_sync(kFull, xmax, o)); ymax = fmax(ymax, __shfl_xor_sync(kFull, ymax, o));
            }
            if (lane == 0) {
                __stcg(reinterpret_cast<double2*>(gb + 4 * (size_t)g), make_double2(xmin, ymin));
                __stcg(reinterpret_cast<double2*>(gb + 4 * (size_t)g + 2), make_double2(xmax, ymax));
            }
        }
    }
}
// FP32 margin of ONE segment's copy relative to the job origin (the tiles use the largest of their segments)
__device__ __forceinline__ float few_margin(float x0, float y0, float x1, float y1)
{
    return 2e-6f * fmaxf(fmaxf(fabsf(x0), fabsf(y0)), fmaxf(fabsf(x1), fabsf(y1))) + 1e-6f;
}
struct MinHit { double t; int s; };
__device__ __forceinline__ void minhit_take(MinHit& a, double t, int sidx)
{
    if (t < a.t || (t == a.t && sidx >= 0 && (a.s < 0 || sidx < a.s))) { a.t = t; a.s = sidx; }
}
// block-wide lexicographic minimum of (t, s); every thread returns the same pair.  scratch: 2 * kcNW doubles.
__device__ __forceinline__ MinHit block_minhit(MinHit v, double* scratch, int lane, int warp)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double t2 = __shfl_xor_sync(kFull, v.t, o);
        const int s2 = __shfl_xor_sync(kFull, v.s, o);
        minhit_take(v, t2, s2);
    }
    __syncthreads();   // the scratch may still be read from the previous reduction
    if (lane == 0) { scratch[2 * warp] = v.t; scratch[2 * warp + 1] = __longlong_as_double((long long)v.s); }
    __syncthreads();
    MinHit r; r.t = scratch[0]; r.s = (int)__double_as_longlong(scratch[1]);
#pragma unroll
    for (int w = 1; w < kcNW; ++w) minhit_take(r, scratch[2 * w], (int)__double_as_longlong(scratch[2 * w + 1]));
    return r;
}
__device__ __forceinline__ double block_min_d(double v, double* scratch, int lane, int warp)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(kFull, v, o));
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double r = scratch[0];
#pragma unroll
    for (int w = 1; w < kcNW; ++w) r = fmin(r, scratch[w]);
    return r;
}

template <int K>
__device__ __forceinline__ bool corridor_search_few_c(const Part& pt, const PathView& pv, double* sB, int* sMisc, const double* __restrict__ gbox,
                                                      unsigned* sHint, unsigned short* sClr,
                                                      const double* __restrict__ gseg, const double* __restrict__ gcenter,
                                                      unsigned long long* __restrict__ gcert, unsigned long long* __restrict__ gapex,
                                                      long long segI0, long long segO0, long long segE, double guard, unsigned mask,
                                                      bool parity_ok, double (&loc)[K], double (&hic)[K], long long& ray_tests, int& ex_scans,
                                                      double* tdbg = nullptr)
{
    constexpr int T = kcT;
    const int tid = pt.tid, lane = pt.lane, warp = pt.warp;
#ifdef RL_PHASE_TIMERS
#define RL_TF(k) do { if (tdbg && tid == 0) { const long long t__ = clock64(); tdbg[k] += (double)(t__ - tf_last); tf_last = t__; } } while (0)
    long long tf_last = clock64();
#else
#define RL_TF(k) do { } while (0)
#endif
    const double INF = dinf();
    const float FINF = __int_as_float(0x7f800000);
    int* sList = reinterpret_cast<int*>(sB);                 // region B is free between the update pass and the staging
    double* sScr = sB + 16;                                  // reduction scratch (16 doubles) behind the list
    int* sCand = reinterpret_cast<int*>(sB + 32);            // groups a pass has to visit (<= kFewMaxGroups)
    const int M[2] = {(int)(segO0 - segI0), (int)(segE - segO0)};
    const int ngr[2] = {(M[0] + 31) >> 5, (M[1] + 31) >> 5};
    if (ngr[0] > kFewMaxGroups || ngr[1] > kFewMaxGroups || !gbox) return false;
    const double* gbx[2] = {gbox, gbox + 4 * (size_t)ngr[0]};
    if (tid == 0) sMisc[0] = 0;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < K; ++j)
        if ((mask >> j) & 1u) { const int slot = atomicAdd(&sMisc[0], 1); if (slot < kFewMax) sList[slot] = tid + j * T; }
    __syncthreads();
    const int nfl = sMisc[0];
    if (nfl > kFewMax) return false;                         // uniform: the caller runs the tile-streaming search instead
    __syncthreads();                                         // every thread has read the count before a pass reuses the word
    const double2 org = pv.sP[0];
    const int rfl[2] = {sMisc[8] & 1, sMisc[9] & 1};
    const float mring[2] = {__int_as_float(sMisc[10]), __int_as_float(sMisc[11])};   // the margin the update passes allow for (>= any segment's)
    long long tests = 0;
    for (int q = 0; q < nfl; ++q) {
        const int i = sList[q];
        const double2 Pc = pv.sP[i];
        double nx, ny;
        normal_c(pv, i, nx, ny);
        const float px = (float)(Pc.x - org.x), py = (float)(Pc.y - org.y), fnx = (float)nx, fny = (float)ny;
        const float pm = 2e-6f * fmaxf(fabsf(px), fabsf(py));
        MinHit hp[2], hn[2];
        double dist[2] = {INF, INF};
        float clr[2] = {3e18f, 3e18f};
        const unsigned hw_old = sHint[i];
        RL_TF(0);
        // ---- sweep 1: exact nearest hits of both rays on both rings (ray_scan_seg without the box hierarchy); the
        //      clearance is taken along the way for the OLD anchor, which the new one usually equals ----
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            MinHit bp, bn; bp.t = INF; bp.s = -1; bn.t = INF; bn.s = -1;
            const double* gs = gseg + 4 * (r ? segO0 : segI0);
            const int mr = M[r];
            const bool clr_ok = (mr <= 8191 && mr > 2 * kWin + 1);
            int anchor_old = ((hw_old >> (28 + r)) & 1u) ? (int)((hw_old >> (13 * r)) & 0x1fffu) : -1;
            if (anchor_old >= mr || !clr_ok) anchor_old = -1;
            float best = 64.f;
            RL_FEW_FOREACH(gs, mr, gbx[r], (few_box_line(bx, Pc, nx, ny) || (anchor_old >= 0 && few_box_near(bx, Pc, 64.5))), {
                float fx0[kFewW]; float fy0[kFewW]; float fx1[kFewW]; float fy1[kFewW]; float m4[kFewW];
                unsigned pass = 0u;
#pragma unroll
                for (int u = 0; u < kFewW; ++u) {
                    fx0[u] = (float)(a4[u].x - org.x); fy0[u] = (float)(a4[u].y - org.y);
                    fx1[u] = (float)(b4[u].x - org.x); fy1[u] = (float)(b4[u].y - org.y);
                    m4[u] = few_margin(fx0[u], fy0[u], fx1[u], fy1[u]) + pm;
                    const float sa = fnx * (fy0[u] - py) - fny * (fx0[u] - px); const float sbb = fnx * (fy1[u] - py) - fny * (fx1[u] - px);
                    if (ok4[u] && !(fminf(sa, sbb) > m4[u] || fmaxf(sa, sbb) < -m4[u])) pass |= (1u << u);
                }
                if (anchor_old >= 0) {
#pragma unroll
                    for (int u = 0; u < kFewW; ++u) {
                        int dj = sg4[u] - anchor_old; if (dj < 0) dj += mr;
                        const bool outside = ok4[u] && !(dj <= kWin || dj >= mr - kWin);
                        const float d = sqrtf(seg_dist2_f(make_float4(fx0[u], fy0[u], fx1[u], fy1[u]), px, py)) - 4.f * m4[u];
                        best = outside ? fminf(best, fmaxf(d, 0.f)) : best;
                    }
                }
#pragma unroll
                for (int u = 0; u < kFewW; ++u) {
                    if ((pass >> u) & 1u) {                                                  // rare: the exact test (main.cpp:478-490)
                        const double vx = b4[u].x - a4[u].x; const double vy = b4[u].y - a4[u].y;   // main.cpp:482
                        const double den = nx * (-vy) + ny * vx;                         // main.cpp:483
                        ++tests;
                        if (!(fabs(den) < 1e-15)) {                                      // main.cpp:484
                            const double ax = a4[u].x - Pc.x; const double ay = a4[u].y - Pc.y;  // main.cpp:485
                            const double inv = 1.0 / den;
                            const double t = (ax * (-vy) + ay * vx) * inv;              // main.cpp:486
                            const double u2 = (nx * ay - ny * ax) * inv;                // main.cpp:487
                            if (u2 >= -1e-12 && u2 <= 1.0 + 1e-12) {                    // main.cpp:488
                                if (t > 0.0) minhit_take(bp, t, sg4[u]);                // +n ray, main.cpp:497
                                else if (t < 0.0) minhit_take(bn, -t, sg4[u]);          // -n ray: t' = -t
                            }
                        }
                    }
                }
            })
            hp[r] = block_minhit(bp, sScr, lane, warp);
            hn[r] = block_minhit(bn, sScr, lane, warp);
            const int anchor = (hp[r].t <= hn[r].t) ? hp[r].s : hn[r].s;
            if (anchor_old >= 0) {                           // uniform
                const float c = (float)block_min_d((double)best, sScr, lane, warp);
                if (anchor == anchor_old) clr[r] = c;
            }
        }
        RL_TF(1);
        // ---- sweep 2, only where needed: the clearance of a NEW anchor (clearance_scan) and, for a ring some ray misses,
        //      the exact point-ring distance (dist_scan; main.cpp:501-512) ----
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int mr = M[r];
            if (mr == 0) continue;
            const int anchor = (hp[r].t <= hn[r].t) ? hp[r].s : hn[r].s;
            const bool want_clr = (anchor >= 0 && mr <= 8191 && mr > 2 * kWin + 1) && !(clr[r] < 1e18f);
            const bool want_dist = (hp[r].t == INF || hn[r].t == INF);
            if (!want_clr && !want_dist) continue;
            const double* gs = gseg + 4 * (r ? segO0 : segI0);
            float best = 64.f;
            double best2 = INF;
            const double ub = fmin(hp[r].t, hn[r].t);     // a hit point lies on the ring: the point-ring distance is at most that
            const double reach = fmax(want_clr ? 64.5 : 0.0, want_dist ? ub : 0.0);
            RL_FEW_FOREACH(gs, mr, gbx[r], few_box_near(bx, Pc, reach), {
                if (want_clr) {
#pragma unroll
                    for (int u = 0; u < kFewW; ++u) {
                        int dj = sg4[u] - anchor; if (dj < 0) dj += mr;
                        const bool outside = ok4[u] && !(dj <= kWin || dj >= mr - kWin);
                        const float4 f = make_float4((float)(a4[u].x - org.x), (float)(a4[u].y - org.y), (float)(b4[u].x - org.x), (float)(b4[u].y - org.y));
                        const float m = few_margin(f.x, f.y, f.z, f.w) + pm;
                        const float d = sqrtf(seg_dist2_f(f, px, py)) - 4.f * m;
                        best = outside ? fminf(best, fmaxf(d, 0.f)) : best;
                    }
                }
                if (want_dist) {
#pragma unroll
                    for (int u = 0; u < kFewW; ++u) {
                        const double vx = b4[u].x - a4[u].x; const double vy = b4[u].y - a4[u].y;
                        const double apx = Pc.x - a4[u].x; const double apy = Pc.y - a4[u].y;
                        const double denom = fmax(1e-30, vx * vx + vy * vy);
                        const double tt = fmin(1.0, fmax(0.0, (vx * apx + vy * apy) / denom));
                        const double qx = a4[u].x + vx * tt; const double qy = a4[u].y + vy * tt;
                        const double ex = Pc.x - qx; const double ey = Pc.y - qy;
                        best2 = ok4[u] ? fmin(best2, ex * ex + ey * ey) : best2;
                    }
                }
            })
            if (want_clr) clr[r] = (float)block_min_d((double)best, sScr, lane, warp);
            if (want_dist) { const double d2 = block_min_d(best2, sScr, lane, warp); dist[r] = (d2 < INF) ? sqrt(d2) : INF; }
        }
        RL_TF(2);
        // ---- combine (main.cpp:696-710), per-sample state, certificate decision: as in corridor_search_c ----
        double dpos = INF, dneg = INF;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            double vp, vn;
            if (M[r] == 0) { vp = 0.0; vn = 0.0; }
            else {
                const double dd = (dist[r] < INF) ? dist[r] : 0.0;
                vp = (hp[r].t < INF) ? hp[r].t : dd;
                vn = (hn[r].t < INF) ? hn[r].t : dd;
            }
            dpos = fmin(dpos, fmax(0.0, vp));
            dneg = fmin(dneg, fmax(0.0, vn));
        }
        double hv = fmax(0.0, dpos - guard), lv = -fmax(0.0, dneg - guard);
        if (!isfinite(hv)) hv = 0.0;
        if (!isfinite(lv)) lv = 0.0;
#pragma unroll
        for (int j = 0; j < K; ++j)
            if (i == tid + j * T) { hic[j] = hv; loc[j] = lv; }
        const double cx0 = gcenter[2 * i], cy0 = gcenter[2 * i + 1];
        const float disp = __double2float_ru(sqrt((Pc.x - cx0) * (Pc.x - cx0) + (Pc.y - cy0) * (Pc.y - cy0))) * (1.f + 1e-6f);
        unsigned hw = hw_old & (3u << 26);     // parity bits are kept
        unsigned cw = 0u;
        float rcv[2] = {0.f, 0.f};
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int mr = M[r];
            const int anchor = (hp[r].t <= hn[r].t) ? hp[r].s : hn[r].s;
            if (anchor >= 0 && mr > 2 * kWin + 1 && mr <= 8191) {
                const float cval = clr[r] - disp;
                const unsigned cq = (cval >= 63.75f) ? 255u : (cval > 0.f ? (unsigned)(cval * 4.f) : 0u);
                hw |= ((unsigned)anchor << (13 * r)) | (1u << (28 + r));
                cw |= cq << (8 * r);
                rcv[r] = 0.25f * (float)cq - disp - 4.f * (mring[r] * 1.001f + pm + 1e-5f);
            }
        }
        __syncthreads();                          // every thread has read the old hint word
        if (tid == 0) { sHint[i] = hw; sClr[i] = (unsigned short)cw; }
        int cone_r = -1, cone_dir = 0;
        unsigned long long cert_new = 0ull;
        bool cert_set = false;
#pragma unroll
        for (int dir = 0; dir < 2; ++dir) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const MinHit& h = dir ? hn[r] : hp[r];
                if (h.t <= (double)rcv[r]) continue;                                  // a near hit: no question to answer
                if (parity_ok && rfl[r] && ((hw >> (26 + r)) & 1u)) continue;        // inside a closed ring: every ray hits it
                const unsigned key = ((unsigned)r << 2) | ((unsigned)dir << 3);
                if (h.t < INF) { cert_new = ((unsigned long long)(unsigned)h.s << 32) | (unsigned long long)(key | kCertFar); cert_set = true; }
                else { cone_r = r; cone_dir = dir; }
            }
        }
        if (cert_set && tid == 0) gcert[i] = cert_new;
        RL_TF(3);
        // ---- a ring-free cone for the ray that misses (cone_scan over the whole ring) ----
        if (cone_r >= 0) {
            const int r = cone_r, dir = cone_dir;
            const double sgn = dir ? -1.0 : 1.0;
            const int qx = __double2int_rn(sgn * nx * 32767.0), qy = __double2int_rn(sgn * ny * 32767.0);
            const unsigned nw1 = ((unsigned)qx & 0xffffu) | ((unsigned)qy << 16);
            double d0x, d0y;
            cert_axis(nw1, d0x, d0y);
            const double toward = dir ? hp[r].t : hn[r].t;     // this ring's hit on the opposite ray
            const double back = (toward < INF) ? fmax(0.0, toward - fmax(0.05, 0.5 * guard)) : 0.0;
            const float oxf = (float)((Pc.x - cx0) - back * d0x), oyf = (float)((Pc.y - cy0) - back * d0y);
            const double axd = cx0 + (double)oxf, ayd = cy0 + (double)oyf;
            const float ax = (float)(axd - org.x), ay = (float)(ayd - org.y), dx = (float)d0x, dy = (float)d0y;
            float cbest = 0.9396926f;
            const double* gs = gseg + 4 * (r ? segO0 : segI0);
            RL_FEW_FOREACH(gs, M[r], gbx[r], few_box_cone(bx, axd, ayd, d0x, d0y), {
#pragma unroll
                for (int u = 0; u < kFewW; ++u) {
                    const float fx0 = (float)(a4[u].x - org.x); const float fy0 = (float)(a4[u].y - org.y);
                    const float fx1 = (float)(b4[u].x - org.x); const float fy1 = (float)(b4[u].y - org.y);
                    const float e = 2.f * (few_margin(fx0, fy0, fx1, fy1) + pm) + 1e-6f;
                    const float ux = fx0 - ax; const float uy = fy0 - ay; const float wx = fx1 - ax; const float wy = fy1 - ay;
                    const float t1 = ux * dx + uy * dy; const float c1 = ux * dy - uy * dx;    // along / across the axis
                    const float t2 = wx * dx + wy * dy; const float c2 = wx * dy - wy * dx;
                    const float r1 = sqrtf(ux * ux + uy * uy); const float r2 = sqrtf(wx * wx + wy * wy);
                    bool none = (r1 <= 8.f * e || r2 <= 8.f * e);                               // apex (almost) on the ring
                    const bool crosses = !((c1 > e && c2 > e) || (c1 < -e && c2 < -e));
                    const float ds = c1 - c2;
                    const float tc = t1 + (t2 - t1) * __fdividef(c1, ds);
                    none = none || (crosses && (fabsf(ds) < 16.f * e || tc > -e * (4.f + 2.f * __fdividef(fabsf(t2 - t1), fabsf(ds)))));
                    const float k1 = __fdividef(t1 + e, r1) + 2e-6f; const float k2 = __fdividef(t2 + e, r2) + 2e-6f;
                    const float cnew = none ? 2.f : fmaxf(k1, k2);
                    cbest = ok4[u] ? fmaxf(cbest, cnew) : cbest;
                }
            })
            const float cb = -(float)block_min_d(-(double)cbest, sScr, lane, warp);
            if (tid == 0) {
                const unsigned key = ((unsigned)r << 2) | ((unsigned)dir << 3);
                unsigned nw0 = key | kCertNoCone;
                if (cb < 0.9995f) {
                    const unsigned qc = (unsigned)ceilf((cb + 1.f) * 32767.f + 0.5f);
                    if (qc < 65535u) nw0 = key | kCertCone | (qc << 16);
                }
                gcert[i] = ((unsigned long long)nw1 << 32) | (unsigned long long)nw0;
                gapex[i] = ((unsigned long long)__float_as_uint(oyf) << 32) | (unsigned long long)__float_as_uint(oxf);
                ++ex_scans;
            }
        }
        RL_TF(4);
    }
    (void)FINF;
    ray_tests += tests;
    __syncthreads();      // region B is handed on to the staging
    return true;
}

// the UPDATE path of one chunk: returns the mask of samples the searching path has to rebuild
template <int K>
__device__ __forceinline__ unsigned corridor_update_c(const Part& pt, const PathView& pv, double* sB, int* sMisc,
                                                      const unsigned* sHint, const unsigned short* sClr, const double2* sHalo,
                                                      const double* __restrict__ gseg, const double* __restrict__ gcenter,
                                                      const unsigned long long* __restrict__ gcert, const unsigned long long* __restrict__ gapex,
                                                      long long segI0, long long segO0, long long segE, double guard, bool parity_ok,
                                                      double (&loc)[K], double (&hic)[K], long long& ray_tests)
{
    constexpr int T = kcT, NP = kcT * K;
    const int Nl = pv.Nloc, tid = pt.tid;
    const int closed_bits = (pv.openL ? 2 : 0) | (pv.openR ? 4 : 0) | ((!pv.openL && !pv.openR) ? 1 : 0);   // see corridor_update_sample<LOCAL>
    const int M[2] = {(int)(segO0 - segI0), (int)(segE - segO0)};
#pragma unroll
    for (int j = 0; j < K; ++j) { loc[j] = 0.0; hic[j] = 0.0; }
    if (M[0] <= 2 * kWin + 1 || M[1] <= 2 * kWin + 1 || M[0] > 8191 || M[1] > 8191) return 0xffffffffu;
    // ---- the range of each ring this chunk's windows touch: anchors relative to sample 0's anchor, on the circle ----
    const unsigned h0 = sHint[0];
    if (!((h0 >> 28) & 1u) || !((h0 >> 29) & 1u)) return 0xffffffffu;       // (uniform: every thread reads the same word)
    block_sync<T>();
    if (tid == 0) { sMisc[12] = 0; sMisc[13] = 0; sMisc[14] = 0; sMisc[15] = 0; }
    block_sync<T>();
    int jref[2], lmin[2] = {0, 0}, lmax[2] = {0, 0};
#pragma unroll
    for (int r = 0; r < 2; ++r) jref[r] = (int)((h0 >> (13 * r)) & 0x1fffu);
#pragma unroll 1
    for (int j = 0; j < K; ++j) {
        const int i = tid + j * T;
        if (i >= Nl) continue;
        const unsigned hw = sHint[i];
        for (int r = 0; r < 2; ++r) {
            if (!((hw >> (28 + r)) & 1u)) continue;                          // no anchor: that sample is flagged anyway
            int rel = (int)((hw >> (13 * r)) & 0x1fffu) - jref[r];
            if (rel > M[r] / 2) rel -= M[r];
            else if (rel < -(M[r] / 2)) rel += M[r];
            lmin[r] = min(lmin[r], rel); lmax[r] = max(lmax[r], rel);
        }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { lmin[r] = min(lmin[r], __shfl_xor_sync(kFull, lmin[r], o)); lmax[r] = max(lmax[r], __shfl_xor_sync(kFull, lmax[r], o)); }
        if (pt.lane == 0) { atomicMin(&sMisc[12 + 2 * r], lmin[r]); atomicMax(&sMisc[13 + 2 * r], lmax[r]); }
    }
    block_sync<T>();
    int basev[2], nseg[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int rmin = sMisc[12 + 2 * r], rmax = sMisc[13 + 2 * r];
        nseg[r] = rmax - rmin + 2 * kWin + 1;
        basev[r] = jref[r] + rmin - kWin;
        if (nseg[r] >= M[r]) { nseg[r] = M[r]; basev[r] = 0; }
        basev[r] %= M[r];
        if (basev[r] < 0) basev[r] += M[r];
    }
    if ((size_t)(nseg[0] + nseg[1] + 2) * 24 > (size_t)NP * 32) return 0xffffffffu;     // does not fit region B
    // ---- load the vertex ranges (and verify that consecutive segments really share their vertex) ----
    double2* V0 = reinterpret_cast<double2*>(sB);
    double2* V1 = V0 + (nseg[0] + 1);
    float2* F0 = reinterpret_cast<float2*>(V1 + (nseg[1] + 1));
    float2* F1 = F0 + (nseg[0] + 1);
    const double2 org = pv.sP[0];
    int linked = 1;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const double* gs = gseg + 4 * (r ? segO0 : segI0);
        double2* V = r ? V1 : V0;
        float2* F = r ? F1 : F0;
        for (int q = tid; q <= nseg[r]; q += T) {
            int sgl = basev[r] + ((q < nseg[r]) ? q : nseg[r] - 1);
            if (sgl >= M[r]) sgl -= M[r];
            const double4 sv = *reinterpret_cast<const double4*>(gs + 4 * (size_t)sgl);
            double2 v;
            if (q < nseg[r]) {
                v = make_double2(sv.x, sv.y);
                if (q + 1 < nseg[r]) {
                    int sn = sgl + 1; if (sn >= M[r]) sn -= M[r];
                    const double2 nx2 = *reinterpret_cast<const double2*>(gs + 4 * (size_t)sn);
                    linked &= (nx2.x == sv.z && nx2.y == sv.w);
                }
            } else v = make_double2(sv.z, sv.w);
            V[q] = v; F[q] = make_float2((float)(v.x - org.x), (float)(v.y - org.y));
        }
    }
    if (tid == 0) {   // the CTA-uniform context of corridor_update_sample (static shared memory), published by the barrier below
        UpdCtx& c = s_upd;
        const unsigned char* base = reinterpret_cast<const unsigned char*>(pv.sP);    // start of the dynamic shared memory
        c.oV0 = (int)(reinterpret_cast<const unsigned char*>(V0) - base);
        c.oF0 = (int)(reinterpret_cast<const unsigned char*>(F0) - base);
        c.oHint = (int)(reinterpret_cast<const unsigned char*>(sHint) - base); c.oClr = (int)(reinterpret_cast<const unsigned char*>(sClr) - base);
        c.oHalo = (int)(reinterpret_cast<const unsigned char*>(sHalo) - base);
        c.gcenter = gcenter; c.gcert = gcert; c.gapex = gapex;
        c.ox = org.x; c.oy = org.y; c.guard = guard; c.N = Nl; c.M0 = M[0]; c.M1 = M[1]; c.rf0 = sMisc[8] & 1; c.rf1 = sMisc[9] & 1;
        c.mr0 = __int_as_float(sMisc[10]); c.mr1 = __int_as_float(sMisc[11]);
        c.parity_ok = parity_ok; c.closed = closed_bits;
        c.base0 = basev[0]; c.base1 = basev[1]; c.len0 = nseg[0]; c.len1 = nseg[1];
        c.gs0 = gseg + 4 * segI0;   // ring 1's segment records follow ring 0's (segO0 = segI0 + M0)
    }
    if (!__syncthreads_and(linked)) return 0xffffffffu;
    unsigned flagged = 0u;
    double2 cn = make_double2(0.0, 0.0);
    unsigned long long wn = 0ull, an = 0ull;
    if (tid < Nl) { cn = *reinterpret_cast<const double2*>(gcenter + 2 * tid); wn = gcert[tid]; an = gapex[tid]; }
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const int i = tid + j * T;
        const double2 cc = cn;
        const unsigned long long wc = wn, ac = an;
        if (j + 1 < K && i + T < Nl) { cn = *reinterpret_cast<const double2*>(gcenter + 2 * (i + T)); wn = gcert[i + T]; an = gapex[i + T]; }
        if (i >= Nl) continue;
        const UpdRes r = corridor_update_sample<true>(i, cc.x, cc.y, wc, ac);
        ray_tests += r.tests;
        if (r.flagged) flagged |= (1u << j);
        else { hic[j] = r.hv; loc[j] = r.lv; }
    }
    return flagged;
}

// Development build only (-DRL_PHASE_TIMERS): thread 0 of CTA 0 adds clock64() cycles per phase to stats.J0[16..24]
// (global memory; same phase numbering as solve_kernel)
#ifdef RL_PHASE_TIMERS
#define RL_PHC(i) do { if (threadIdx.x == 0 && cl.rank == 0) { const long long t__ = clock64(); st->J0[16 + (i)] += (double)(t__ - ph_last); ph_last = t__; } } while (0)
#else
#define RL_PHC(i) do { } while (0)
#endif

// ---- the cluster solver kernel -----------------------------------------------------------------------------
template <int K, int MODE>
__global__ void __launch_bounds__(kcT, 2)
solve_cluster_kernel(const DevBatch B, const int* __restrict__ job_list, const int* __restrict__ item_off, int n_items)
{
    constexpr int T = kcT, NP = kcT * K;
    constexpr bool EXACT = (MODE == kModeExact), OPEN = (MODE == kModeOpen);
    constexpr bool closed = !OPEN;
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
    unsigned* sHint = reinterpret_cast<unsigned*>(scr + kcBytes);                                   // anchors + valid bits, per sample
    unsigned short* sClr = reinterpret_cast<unsigned short*>(scr + kcBytes + (size_t)NP * 4);       // clearances, per sample

    Clu cl;
    cl.CS = cl_size(); cl.rank = cl_rank();
    cl.left = (cl.rank == 0) ? cl.CS - 1 : cl.rank - 1;
    cl.right = (cl.rank == cl.CS - 1) ? 0 : cl.rank + 1;
    const int cid = (int)(blockIdx.x / cl.CS);
    if (cid >= n_items) return;   // uniform over the cluster
    // one cluster works through one ITEM: a short chain of jobs on the same track (see solve_kernel); the corridor state
    // of every chunk (anchors, clearances, parity, certificates) carries over to the next job of the chain
    const int it0 = item_off[cid], it1 = item_off[cid + 1];
    uint32_t bar_phase = 0;
    int fslot = 0, prev_trk = -1;
    uint64_t* ebar = reinterpret_cast<uint64_t*>(scr + kcEbar);   // per-evaluation exchange barriers: one arrival per local warp + tx bytes
    uint32_t epar = 0;
    if (threadIdx.x == 0) {
        mbar_init(mbar, 1);
        mbar_init(ebar, kcNW);
        mbar_init(ebar + 1, kcNW);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        sFlag[0] = 0; sFlag[1] = 0; sFlag[2] = 0;
    }
    cl_sync();   // every CTA of the cluster is running and initialised before any distributed-shared-memory access
  for (int itj = it0; itj < it1; ++itj) {
    const int jid = job_list[itj];
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
        pt.first_chunk = (cl.rank == 0); pt.last_chunk = (cl.rank == cl.CS - 1);
    }
    const int tid = pt.tid, cnt = pt.cnt, start = pt.start;
    PathView pv; pv.sP = sP; pv.halo = sHalo; pv.Nloc = Nl;
    pv.openL = OPEN && pt.first_chunk; pv.openR = OPEN && pt.last_chunk;

    if (tid == 0) {
        if (cl.rank == 0) {
            st->status = RL_OK; st->n = N; st->outer_done = 0; st->accepted = 0; st->backtracks = 0; st->evals = 0;
            st->vpass_rounds = 0; st->exist_scans = 0; st->ray_tests = 0; st->lap_time = 0.0;
            for (int o = 0; o < RL_MAX_OUTER_LOG; ++o) {
                st->J0[o] = 0.0; st->Jend[o] = 0.0; st->lap_outer[o] = 0.0; st->acc_outer[o] = 0; st->bt_outer[o] = 0;
            }
#ifdef RL_PHASE_TIMERS
            for (int i = 0; i < 9; ++i) st->J0[16 + i] = 0.0;
#endif
        }
        fence_proxy_async();
    }
#ifdef RL_PHASE_TIMERS
    long long ph_last = clock64();
#endif
    block_sync<T>();

    // ---- load this chunk of the centre line (TMA bulk copy) and trade end points with the neighbours ----
    if (tid == 0) {
        mbar_expect_tx(mbar, (uint32_t)Nl * 16u);
        bulk_g2s(sP, B.center_xy + 2 * (s0 + cl.n0), (uint32_t)Nl * 16u, mbar);
    }
    mbar_wait(mbar, bar_phase); bar_phase ^= 1;
    // per-group bounding boxes of both rings for corridor_search_few_c: in the job's ax rows (free until the final profile),
    // built by the whole cluster and published by the cluster barrier of the halo exchange
    const long long segI0 = B.seg_off[2 * trk], segO0 = B.seg_off[2 * trk + 1], segE = B.seg_off[2 * trk + 2];
    double* gbox = nullptr;
    if (!ev && !B.no_few_search && 4ll * ((((segO0 - segI0) + 31) >> 5) + (((segE - segO0) + 31) >> 5)) + 1 <= (long long)N) {
        gbox = B.ax + ((B.job_off[jid] + 1) & ~1ll);        // 16-byte aligned (rows start at any multiple of 8 bytes)
        few_boxes_build(cl, tid, B.seg, segI0, segO0, segE, gbox);
    }
    exchange_path_halo(sP, sHalo, cl, tid);

#pragma unroll
    for (int k = 0; k < K; ++k)
        if (k < cnt) { B.alpha_total[row0 + start + k] = 0.0; B.alpha_last[row0 + start + k] = 0.0; }

    // the v(s) constants are CTA-uniform: they live behind the exchange arrays of the sweeps in region B while a profile
    // runs (thread 0 writes them before the barrier that precedes it) instead of in 25 registers for the whole job --
    // this kernel has no static shared memory to spare (two CTAs fill the SM to 112 bytes)
    VPar* const sq = reinterpret_cast<VPar*>(sB + 6 * (kcT + 1) + 2);
    const VPar& q = *sq;
    auto fill_vpar = [&]() {
        if (tid == 0) {
            VPar& w = *sq;
            w.v_cap = C.v_cap_mps; w.a_lat_max = C.a_lat_max; w.kappa_eps = C.kappa_eps;
            const double a_total = C.use_total_ge_lat ? fmax(C.a_total_max, C.a_lat_max) : C.a_total_max;   // main.cpp:802-804
            w.a_tot2 = a_total * a_total;
            w.kd = 0.5 * C.rho_air * C.Cd * C.A_front_m2; w.Fr = C.mass_kg * 9.81 * C.c_rr; w.mass = C.mass_kg; w.inv_mass = 1.0 / C.mass_kg; w.P = C.P_max_W;
            w.acc_cap = C.a_long_acc_cap; w.brk_cap = C.a_long_brake_cap; w.h = h; w.has_power = (C.P_max_W > 0);
        }
    };

    const HStep H(h);
    const double inv2h = H.inv2h, invh2 = H.invh2;                       // DiffOps, main.cpp:547
    const double lamJ = C.lambda_smooth * inv2h * inv2h;
    long long ray_tests = 0;
    int vrounds = 0, ph = 0;
    int acc_total = 0, bt_total = 0, ev_total = 0;
    const int max_outer = ev ? 0 : C.max_outer_iters;

    // initial corridor from the centre line: guard uses the veh_width ARGUMENT (main.cpp:706 / 930)
    const double* gcenter = B.center_xy + 2 * (s0 + cl.n0);
    unsigned long long* gcert = reinterpret_cast<unsigned long long*>(B.heading + row0);      // certificates: scratch in the chunk's
    unsigned long long* gapex = reinterpret_cast<unsigned long long*>(B.curvature + row0);    // heading / curvature rows until the end
    int ex_scans = 0;
    const bool parity_ok = (C.veh_width_arg * 0.5 + C.safety_margin_m >= 0.0) && (C.veh_width_m * 0.5 + C.safety_margin_m >= 0.0);
    if (!ev) {
        const double guard0 = C.veh_width_arg * 0.5 + C.safety_margin_m;
        const bool same_track = (trk == prev_trk);   // the previous job of this chain left its corridor state behind
        double loc[K], hic[K];
#pragma unroll
        for (int j = 0; j < K; ++j) { loc[j] = 0.0; hic[j] = 0.0; }
        if (!same_track) {
            for (int i = tid; i < NP; i += T) { sHint[i] = 0u; sClr[i] = 0; }
            for (int i = tid; i < Nl; i += T) gcert[i] = 0ull;
            if (tid == 0) { sMisc[8] = 0; sMisc[9] = 0; sMisc[10] = 0; sMisc[11] = 0; }
            block_sync<T>();
            corridor_search_c<K>(pt, pv, sB, mbar, bar_phase, sMisc, sHint, sClr, B.seg, gcenter, gcert, gapex, segI0, segO0, segE,
                                 guard0, 0xffffffffu, true, parity_ok, loc, hic, ray_tests, ex_scans);
        } else {
            unsigned flagged = corridor_update_c<K>(pt, pv, sB, sMisc, sHint, sClr, sHalo, B.seg, gcenter, gcert, gapex, segI0, segO0, segE,
                                                    guard0, parity_ok, loc, hic, ray_tests);
            if (block_or<T>(flagged != 0u)) {
                if (!corridor_search_few_c<K>(pt, pv, sB, sMisc, gbox, sHint, sClr, B.seg, gcenter, gcert, gapex, segI0, segO0, segE,
                                              guard0, flagged, parity_ok, loc, hic, ray_tests, ex_scans))
                    corridor_search_c<K>(pt, pv, sB, mbar, bar_phase, sMisc, sHint, sClr, B.seg, gcenter, gcert, gapex, segI0, segO0, segE,
                                         guard0, flagged, false, parity_ok, loc, hic, ray_tests, ex_scans);
            }
        }
        corridor_stage_out<T, K>(pt, sB, loc, hic);
    }
    RL_PHC(0);   // setup + first corridor of the job

    double* sC0 = sB + tid;
    double* sCp = pair_base_a(sB + NP, NP, tid);
    double* sCm = pair_base_b(sB + NP, NP, tid);
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
                derivs_c(pv, il, H, xp, yp, xpp, ypp);
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
        double* sLo = pair_base_a(reinterpret_cast<double*>(sP), NP, tid);
        double* sHi = pair_base_b(reinterpret_cast<double*>(sP), NP, tid);
        staged_bounds_home<T, K>(pt, sB, sLo, sHi);   // the corridor's bounds: staging area -> their home for the PGD
        RL_PHC(1);   // linearisation + parking the path
        double gam[K];
#pragma unroll
        for (int k = 0; k < K; ++k) gam[k] = 1.0;
        double lap_outer = 0.0;
        if (mt) {
            // ============ v(s) profile + time weights (main.cpp:944-977) ============
            double kap[K], vv[K], axd[K];
#pragma unroll
            for (int k = 0; k < K; ++k) kap[k] = (k < cnt) ? N0[k] / Wd[k] : 0.0;    // kappa, main.cpp:618
            fill_vpar();       // region B is free since the barrier that ended staged_bounds_home
            block_sync<T>();
            vprofile_c<K>(pt, cl, q, kap, vv, C.max_vpass_iters, sB, sFlag, fslot, vrounds, closed);
            block_sync<T>();
            lap_outer = lap_and_ax_c<K>(pt, cl, q, vv, axd, sB, sRed + ph * kcRedStride, closed);
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
        RL_PHC(2);   // v(s) profile + time weights
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
                    if (OPEN) {   // DiffOpsOpen, main.cpp:563-575: D1 one-sided with 1/h at the ends, D2 zero there
                        const int gi = cl.n0 + start + k;
                        if (gi == 0) { cp = 2.0 * c1; cm = 0.0; }
                        else if (gi == N - 1) { cp = 0.0; cm = -2.0 * c1; }
                    }
                    if (k == 0) { f0 = c0; fp = cp; fm = cm; }
                    l0 = c0; lp = cp; lm = cm;
                }
                sC0[k * T] = c0; st_pair<T>(sCp, sCm, k, cp, cm);
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
            cL[0] = sB[pt.tL + kl * T]; cR[0] = sB[pt.tR];
            ld_pair<T>(pair_base_a(sB + NP, NP, pt.tL), pair_base_b(sB + NP, NP, pt.tL), kl, cL[1], cL[2]);
            ld_pair<T>(pair_base_a(sB + NP, NP, pt.tR), pair_base_b(sB + NP, NP, pt.tR), 0, cR[1], cR[2]);
            if (tid == 0) { cL[0] = sCoef[0]; cL[1] = sCoef[1]; cL[2] = sCoef[2]; }
            if (tid == T - 1) { cR[0] = sCoef[4]; cR[1] = sCoef[5]; cR[2] = sCoef[6]; }
            if (OPEN) {   // nothing beyond the two ends of an open track
                if (tid == 0 && pt.first_chunk) { cL[0] = 0.0; cL[1] = 0.0; cL[2] = 0.0; }
                if (tid == T - 1 && pt.last_chunk) { cR[0] = 0.0; cR[1] = 0.0; cR[2] = 0.0; }
            }
        }
        if (!EXACT) {
            // the first unused slot mirrors the right neighbour's first sample (position cnt+1 of the window)
            if (cnt < K) { sC0[cnt * T] = cR[0]; st_pair<T>(sCp, sCm, cnt, cR[1], cR[2]); }
        }
        RL_PHC(3);   // stencil coefficients (+ their exchange)
        // =================== projected gradient with Armijo (main.cpp:723-742 / 996-1026) ===================
        const PgdOut po = pgd_outer_c<K, MODE>(pt, cl, sLo, sHi, cL, cR, sC0, sCp, sCm, sSt, sRed, sExF, sExL, ph, lamJ,
                                               C.step_init, C.step_min, C.armijo_c, C.max_inner_iters, ebar, epar);
        acc_total += po.acc; bt_total += po.bt; ev_total += po.ev;
        if (tid == 0 && cl.rank == 0 && outer < RL_MAX_OUTER_LOG) {
            st->J0[outer] = po.J0; st->Jend[outer] = po.Jend; st->lap_outer[outer] = lap_outer;
            st->acc_outer[outer] = po.acc; st->bt_outer[outer] = po.bt;
        }
        RL_PHC(4);   // projected-gradient loop
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
        RL_PHC(5);   // path back + path update + halo exchange
        // =================== corridor from the new path (main.cpp:749-756 / 1033-1040) ===================
        // (the reference also rebuilds it after the LAST path update, but nothing reads that corridor: skipped)
        if (outer + 1 == max_outer) continue;
        {
            const double guard = C.veh_width_m * 0.5 + C.safety_margin_m;
            double loc[K], hic[K];
#ifdef RL_PHASE_TIMERS
            const long long tc0 = clock64();   // per-RANK corridor times (the other timers run on rank 0 only): Jend / lap_outer / acc_outer[16 + rank]
#endif
            unsigned flagged = corridor_update_c<K>(pt, pv, sB, sMisc, sHint, sClr, sHalo, B.seg, gcenter, gcert, gapex, segI0, segO0, segE,
                                                    guard, parity_ok, loc, hic, ray_tests);
            RL_PHC(6);   // corridor update pass
#ifdef RL_PHASE_TIMERS
            const long long tc1 = clock64();
            int fb = 0;
#endif
            if (block_or<T>(flagged != 0u)) {  // per CTA: the searching path rebuilds the flagged samples (no cluster traffic inside)
#ifdef RL_PHASE_TIMERS
                const long long tf0 = clock64();
#endif
                const bool few = corridor_search_few_c<K>(pt, pv, sB, sMisc, gbox, sHint, sClr, B.seg, gcenter, gcert, gapex, segI0, segO0, segE,
                                                          guard, flagged, parity_ok, loc, hic, ray_tests, ex_scans
#ifdef RL_PHASE_TIMERS
                                                          , (cl.rank == 0) ? &st->J0[25] : nullptr
#endif
                                                          );
#ifdef RL_PHASE_TIMERS
                if (tid == 0 && (cl.rank & 7) == cl.rank) {
                    if (few) { st->Jend[24 + cl.rank] += (double)(clock64() - tf0); st->bt_outer[24 + cl.rank] += sMisc[0]; }
                    else st->acc_outer[24 + cl.rank] += 1;
                }
#endif
                if (!few)
                    corridor_search_c<K>(pt, pv, sB, mbar, bar_phase, sMisc, sHint, sClr, B.seg, gcenter, gcert, gapex, segI0, segO0, segE,
                                         guard, flagged, false, parity_ok, loc, hic, ray_tests, ex_scans);
#ifdef RL_PHASE_TIMERS
                fb = 1;
#endif
            }
            corridor_stage_out<T, K>(pt, sB, loc, hic);
#ifdef RL_PHASE_TIMERS
            if (tid == 0) {
                const long long tc2 = clock64();
                st->Jend[16 + cl.rank] += (double)(tc1 - tc0); st->lap_outer[16 + cl.rank] += (double)(tc2 - tc1); st->acc_outer[16 + cl.rank] += fb;
            }
#endif
            RL_PHC(7);   // searching path for flagged samples + staging
        }
    }

    // the next job of the chain inherits this chunk's certificates (same thread, same samples: no barrier needed)
    if (!ev && itj + 1 < it1) {
        const int njid = job_list[itj + 1];
        if (B.jobs[njid].track == trk) {
            const long long nrow0 = B.job_off[njid] + cl.n0;
            unsigned long long* ncert = reinterpret_cast<unsigned long long*>(B.heading + nrow0);
            unsigned long long* napex = reinterpret_cast<unsigned long long*>(B.curvature + nrow0);
            for (int i = tid; i < Nl; i += T) { ncert[i] = gcert[i]; napex[i] = gapex[i]; }
        }
    }
    prev_trk = ev ? -1 : trk;
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
                derivs_c(pv, il, H, xp, yp, xpp, ypp);
                kap[k] = (xp * ypp - yp * xpp) / pow15(xp * xp + yp * yp);
                B.heading[row0 + il] = atan2(yp, xp);
                B.curvature[row0 + il] = kap[k];
            }
        }
        if (mt) {
            // final v(s) profile (main.cpp:1047)
            double vv[K], axd[K];
            block_sync<T>();
            fill_vpar();
            block_sync<T>();
            vprofile_c<K>(pt, cl, q, kap, vv, C.max_vpass_iters, sB, sFlag, fslot, vrounds, closed);
            block_sync<T>();
            lap = lap_and_ax_c<K>(pt, cl, q, vv, axd, sB, sRed + ph * kcRedStride, closed);
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
        double rt = (double)ray_tests, es = (double)ex_scans;
        cluster_sum2(rt, es, sRed + ph * kcRedStride, cl, pt.lane, pt.warp);
        ph ^= 1;
        if (tid == 0 && cl.rank == 0) {
            st->outer_done = max_outer; st->accepted = acc_total; st->backtracks = bt_total; st->evals = ev_total;
            st->vpass_rounds = vrounds; st->ray_tests = (long long)rt; st->lap_time = lap; st->exist_scans = (int)es;
        }
    }
    RL_PHC(8);   // certificates hand-over, final geometry, final v(s) profile, stores
    cl_sync();   // the next job of the chain reuses the shared-memory regions of every CTA
  }
}

}  // namespace

inline size_t smem_bytes_cluster(int K) { return (size_t)kcT * K * (48 + 6) + kcBytes; }   // + per-sample corridor state

int launch_solve_cluster(const DevBatch& B, const int* job_list, const int* item_off, int n_items, int cs, int mode, void* stream)
{
    constexpr int K = 8;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n_items * cs));
    cfg.blockDim = dim3(kcT);
    cfg.dynamicSmemBytes = smem_bytes_cluster(K);
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (cs > cluster_max_size()) return (int)cudaErrorInvalidConfiguration;
    cudaError_t e;
    if (mode == 1) e = cudaLaunchKernelEx(&cfg, solve_cluster_kernel<K, 1>, B, job_list, item_off, n_items);
    else if (mode == 2) e = cudaLaunchKernelEx(&cfg, solve_cluster_kernel<K, 2>, B, job_list, item_off, n_items);
    else e = cudaLaunchKernelEx(&cfg, solve_cluster_kernel<K, 0>, B, job_list, item_off, n_items);
    return (int)e;
}
namespace { int g_cluster_max = 8; }
// largest cluster the device schedules for this kernel: 16 (non-portable size, tracks up to 32,768 samples) when the
// occupancy query says at least one such cluster fits, else the portable 8
int cluster_max_size() { return g_cluster_max; }
int configure_solve_cluster()
{
    const int smem = (int)smem_bytes_cluster(8);
    cudaError_t e = cudaFuncSetAttribute(solve_cluster_kernel<8, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(solve_cluster_kernel<8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(solve_cluster_kernel<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    g_cluster_max = 8;
    if (cudaFuncSetAttribute(solve_cluster_kernel<8, 0>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
        cudaFuncSetAttribute(solve_cluster_kernel<8, 1>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
        cudaFuncSetAttribute(solve_cluster_kernel<8, 2>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(16); cfg.blockDim = dim3(kcT); cfg.dynamicSmemBytes = (size_t)smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 16; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, solve_cluster_kernel<8, 0>, &cfg) == cudaSuccess && n >= 1) g_cluster_max = 16;
    }
    cudaGetLastError();
    return 0;
}

}  // namespace rl
