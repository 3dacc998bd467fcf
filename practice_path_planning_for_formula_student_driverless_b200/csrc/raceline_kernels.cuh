// raceline_kernels.cuh -- hand-written sm_100a kernels for the batched raceline solver.
//
// One CTA of T threads solves one job = one stage (min-curvature or min-time) of one
// (track, Config) problem, start to finish, with no host round trip:
//
//   reference stage                                   here
//   compute_min_curvature_raceline  main.cpp:683-764   solve_kernel, stage RL_STAGE_MINCURV
//   compute_min_time_raceline       main.cpp:905-1052  solve_kernel, stage RL_STAGE_MINTIME
//
// Data layout.  Samples are BLOCKED over threads: thread t owns `cnt` (<= K) consecutive
// samples, kept in registers through the projected-gradient loop.  Shared memory holds
//   sP      [T*K] double2   the current path P (persistent; TMA bulk-loaded / bulk-stored)
//   region B 4*T*K doubles  phase-dependent:
//        PGD phase : c0 | cp | cm | stash      (stencil coefficients + last accepted alpha), slot-major [k][t]
//        ray phase : ring-segment tile (x0,y0,vx,vy per segment) + one bounding box per kSegBlock segments
//        v(s) phase: neighbour-exchange arrays of the relaxation sweeps
//   scratch 2 KB            mbarrier, reduction partials and warp-edge halos (double-buffered)
//
// Per cost/gradient evaluation (main.cpp:654-675 / 866-895) a thread needs the trial alpha of its own
// samples plus a 2-sample halo on each side (the gradient of the periodic 3-point stencils has a
// 5-point footprint).  Halos travel by warp shuffle, warp-edge halos through shared memory, and the
// (J, g.dalpha) reduction shares the SAME barrier: one __syncthreads per evaluation.  The Armijo
// decision (main.cpp:734) is taken redundantly by every thread from bit-identical reduced values.
//
// Tensor cores are not used on purpose: D^T D is pentadiagonal, there is no dense contraction here.
#include <cuda_runtime.h>
#include <math_constants.h>

#include <cstdint>

#pragma once
#include "raceline_device.h"

namespace rl {
namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int SB = kSegBlock;   // segments per box
constexpr int SU = kSupBlock;   // boxes per super box (fast corridor path)

// scratch layout (bytes)
constexpr int kScrBar = 0;
constexpr int kScrRed = 64;             // [2][16][2] doubles
constexpr int kScrExF = kScrRed + 512;  // [2][16][2] doubles: first two samples of each warp's lane 0
constexpr int kScrExL = kScrExF + 512;  // [2][16][2] doubles: last two samples of each warp's lane 31
constexpr int kScrMisc = kScrExL + 512; // a few ints
constexpr int kScrBytes = kScratchBytes; // (raceline_device.h) followed by the per-sample corridor hint words [T*K]
static_assert(kScrMisc + 64 <= kScrBytes, "scratch layout");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- debug-checks build (-DRL_DEBUG_CHECKS, library variant "_dbg"; compute-sanitizer is not available on the pool) ----
// Region B of the shared memory changes owner five times per outer iteration (stencil coefficients + stash | ring tile |
// both rings' vertices | corridor staging | v(s) exchange arrays) and the path area once (path | box bounds); what
// keeps the owners apart are hand-placed block barriers.  The debug build checks that protocol and the layout:
//   * phase hand-over: every thread counts the region-B phases it has FINISHED in sDone[tid]; entering a phase (after the
//     barrier that is supposed to free the region) it checks that the thread with its lane in every warp has finished as
//     many phases as itself.  A missing or misplaced barrier lets a fast warp enter while a slow one still reads: caught
//     whenever the timing produces it, reported with the phase code.
//   * guard zones of 64 bytes with a known pattern between / after the shared-memory regions, verified at every job end;
//   * bounds asserts where a phase carves its views out of a region.
// Failures are counted in global memory (rl_debug_check_failures); the solve goes on.
#ifdef RL_DEBUG_CHECKS
constexpr int kDbgGap = 64;
constexpr int kDbgScr = 2048;                 // sDone[T <= 512] behind the ordinary scratch
constexpr unsigned kDbgCanary = 0xC0FFEE11u;
enum { kDbgPhCoef = 1, kDbgPhVsweep = 2, kDbgPhVerts = 3, kDbgPhStage = 4, kDbgPhTile = 5, kDbgPhLap = 6 };
__device__ __noinline__ void dbg_fail(int* sMisc, int code)
{
    unsigned long long* g = *reinterpret_cast<unsigned long long**>(sMisc + 24);
    if (g) { atomicAdd(g, 1ull); atomicCAS(g + 1, 0ull, (unsigned long long)(unsigned)code | ((unsigned long long)blockIdx.x << 32)); }
}
__device__ __forceinline__ int* dbg_done(int* sMisc) { return sMisc + (2048 - kScrMisc) / 4; }
template <int T>
__device__ __forceinline__ void dbg_enter(int* sMisc, int phase)
{
    const int* sDone = dbg_done(sMisc);
    const int mine = sDone[threadIdx.x], lane = threadIdx.x & 31;
    bool ok = true;
#pragma unroll 1
    for (int w = 0; w < T / 32; ++w) ok = ok && (sDone[w * 32 + lane] >= mine);
    if (!ok) dbg_fail(sMisc, 1000 + phase);
}
__device__ __forceinline__ void dbg_leave(int* sMisc, int phase)
{
    int* sDone = dbg_done(sMisc);
    if (sMisc[27] == phase && sMisc[26] > 0 && threadIdx.x < 32) return;   // fault injection (tests): warp 0 "forgets" to hand this phase over
    sDone[threadIdx.x] = sDone[threadIdx.x] + 1;
}
#define RL_DBG_ENTER(T, sMisc, ph) dbg_enter<T>(sMisc, ph)
#define RL_DBG_LEAVE(sMisc, ph) dbg_leave(sMisc, ph)
#define RL_DBG_ASSERT(sMisc, cond, code) do { if (!(cond)) dbg_fail(sMisc, code); } while (0)
#else
constexpr int kDbgGap = 0;
constexpr int kDbgScr = 0;
#define RL_DBG_ENTER(T, sMisc, ph) do { } while (0)
#define RL_DBG_LEAVE(sMisc, ph) do { } while (0)
#define RL_DBG_ASSERT(sMisc, cond, code) do { } while (0)
#endif
__device__ __forceinline__ double dinf() { return __longlong_as_double(0x7ff0000000000000LL); }

// ---- TMA 1-D bulk copies (cp.async.bulk -> SASS UBLKCP) completed through an mbarrier -------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// issue a shared->global bulk store and return once its shared-memory source may be overwritten
__device__ __forceinline__ void bulk_s2g_issue(void* gdst, const void* ssrc, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* ssrc, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int T>
__device__ __forceinline__ void block_sync()
{
    if (T == 32) __syncwarp(); else __syncthreads();
}
template <int T>
__device__ __forceinline__ int block_or(int pred)
{
    if (T == 32) return __any_sync(kFull, pred);
    return __syncthreads_or(pred);
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
// Deterministic block-wide sum of two values; the result is bit-identical in every thread.
// Contains exactly one block barrier (a warp barrier for single-warp CTAs).
// Warp stage: ONE packed butterfly for both values (after the first exchange lanes 0-15 carry `a`, lanes 16-31 carry
// `b`): 5 dependent shuffle steps instead of 10.  Block stage: the per-warp pairs are added as a balanced tree, so
// the adds of one level are independent (a serial `+=` chain costs one FP64 latency per warp).
template <int T>
__device__ __forceinline__ void block_sum2(double& a, double& b, double* sred, int lane, int warp)
{
    const bool hi_half = (lane & 16) != 0;
    const double keep = hi_half ? b : a, send = hi_half ? a : b;
    double v = keep + __shfl_xor_sync(kFull, send, 16);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    if (T > 32) {
        constexpr int NW = T / 32;
        if ((lane & 15) == 0) sred[2 * warp + (lane >> 4)] = v;
        __syncthreads();
        double2 e[NW];
        // all partial pairs are requested before the first add: `volatile` keeps the compiler from recycling two
        // registers through NW/2 dependent load -> add round trips (one LDS latency each)
        const uint32_t sa = smem_u32(sred);
#pragma unroll
        for (int w = 0; w < NW; ++w)
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(e[w].x), "=d"(e[w].y) : "r"(sa + 16u * w));
#pragma unroll
        for (int n = NW / 2; n > 0; n >>= 1) {
#pragma unroll
            for (int w = 0; w < n; ++w) { e[w].x = e[2 * w].x + e[2 * w + 1].x; e[w].y = e[2 * w].y + e[2 * w + 1].y; }
        }
        a = e[0].x; b = e[0].y;
    } else {
        a = __shfl_sync(kFull, v, 0);
        b = __shfl_sync(kFull, v, 16);
        __syncwarp();
    }
}

// how the N samples of a job are blocked over the T threads of its CTA
struct Part {
    int N, Tact, cnt, start, tL, tR, cntL, srcL, srcR, tid, lane, warp;
    // open tracks on a cluster: only the first / last chunk holds an end of the track (always true in solve_kernel)
    bool first_chunk = true, last_chunk = true;
};

// ---- per-sample geometry ---------------------------------------------------------------------------
// kernel modes: 0 = closed track, ragged N; 1 = closed track, N == T*K exactly; 2 = open track (ragged)
constexpr int kModeClosed = 0, kModeExact = 1, kModeOpen = 2;

// tangent -> unit normal with the degenerate-case rules of main.cpp:588-592
// One square root and one reciprocal: |(-ty, tx)| = |(tx, ty)|, so the reference's two length tests coincide (after its
// first fix-up t := (1, 0) the second can never fire), and nv / len is taken as nv * (1 / len) (<= 1 ulp apart).
__device__ __forceinline__ void normal_from_tangent(double tx, double ty, double& nx, double& ny)
{
    const double len = sqrt(ty * ty + tx * tx);
    if (len < 1e-15) { nx = -0.0; ny = 1.0; }       // t := (1, 0)  ->  n = (-0, 1)
    else { const double r = 1.0 / len; nx = -ty * r; ny = tx * r; }
}
// normals_from_points_generic, main.cpp:581-593 (closed: periodic central difference; open: one-sided ends)
__device__ __forceinline__ void normal_at(const double2* sP, int i, int N, bool closed, double& nx, double& ny)
{
    double tx, ty;
    if (N == 1) { tx = 1.0; ty = 0.0; }
    else if (closed) {
        const double2 Pm = sP[(i == 0) ? N - 1 : i - 1], Pp = sP[(i == N - 1) ? 0 : i + 1];
        tx = (Pp.x - Pm.x) * 0.5; ty = (Pp.y - Pm.y) * 0.5;
    } else if (i == 0) { tx = sP[1].x - sP[0].x; ty = sP[1].y - sP[0].y; }
    else if (i == N - 1) { tx = sP[N - 1].x - sP[N - 2].x; ty = sP[N - 1].y - sP[N - 2].y; }
    else { tx = (sP[i + 1].x - sP[i - 1].x) * 0.5; ty = (sP[i + 1].y - sP[i - 1].y) * 0.5; }
    normal_from_tangent(tx, ty, nx, ny);
}
// the sample spacing and its reciprocals, formed once per job: the difference quotients below multiply by them
// (the reference divides by 2h, h*h and h; <= 1 ulp apart, and six FP64 divisions per sample and outer iteration fewer)
struct HStep {
    double h, inv_h, inv2h, invh2;
    __device__ __forceinline__ HStep() {}
    __device__ __forceinline__ explicit HStep(double hh) : h(hh), inv_h(1.0 / hh), inv2h(1.0 / (2 * hh)), invh2(1.0 / (hh * hh)) {}
};
// solve_kernel keeps the current job's step constants in static shared memory (CTA-uniform; thread 0 writes them at the
// start of the job), like the v(s) constants below: read where they are needed instead of held in registers for the job
__shared__ double s_hstep[4];
// central-difference derivatives of the `deriv` lambda (interior / periodic case), main.cpp:599-603 / 625-629
__device__ __forceinline__ void derivs_central(double2 Pm, double2 Pc, double2 Pp, const HStep& H,
                                               double& xp, double& yp, double& xpp, double& ypp)
{
    xp = (Pp.x - Pm.x) * H.inv2h; yp = (Pp.y - Pm.y) * H.inv2h;
    xpp = (Pp.x - 2 * Pc.x + Pm.x) * H.invh2; ypp = (Pp.y - 2 * Pc.y + Pm.y) * H.invh2;
}
// the `deriv` lambda, main.cpp:599-613 / 625-639
__device__ __forceinline__ void derivs_at(const double2* sP, int i, int N, const HStep& H, bool closed,
                                          double& xp, double& yp, double& xpp, double& ypp)
{
    if (N == 1) { xp = 1.0; yp = 0.0; xpp = 0.0; ypp = 0.0; return; }
    if (closed || (i > 0 && i < N - 1)) {
        derivs_central(sP[(i == 0) ? N - 1 : i - 1], sP[i], sP[(i == N - 1) ? 0 : i + 1], H, xp, yp, xpp, ypp);
    } else if (i == 0) {
        xp = (sP[1].x - sP[0].x) * H.inv_h; yp = (sP[1].y - sP[0].y) * H.inv_h;
        if (N >= 3) { xpp = (sP[2].x - 2 * sP[1].x + sP[0].x) * H.invh2; ypp = (sP[2].y - 2 * sP[1].y + sP[0].y) * H.invh2; }
        else { xpp = 0.0; ypp = 0.0; }
    } else {
        xp = (sP[N - 1].x - sP[N - 2].x) * H.inv_h; yp = (sP[N - 1].y - sP[N - 2].y) * H.inv_h;
        if (N >= 3) { xpp = (sP[N - 1].x - 2 * sP[N - 2].x + sP[N - 3].x) * H.invh2; ypp = (sP[N - 1].y - 2 * sP[N - 2].y + sP[N - 3].y) * H.invh2; }
        else { xpp = 0.0; ypp = 0.0; }
    }
}
// pow(max(1e-12, q), 1.5), main.cpp:617 / 647
__device__ __forceinline__ double pow15(double q)
{
    const double s = fmax(1e-12, q);
    return s * sqrt(s);
}

// ---- v(s) profile pieces: the ax_max_at lambda, main.cpp:797-824 -----------------------------------
struct VPar {
    double v_cap, a_lat_max, kappa_eps, a_tot2, kd, Fr, mass, inv_mass, P, acc_cap, brk_cap, h;
    int has_power;
};
__shared__ VPar s_vpar;   // solve_kernel: the current job's constants (written by thread 0 at the start of the job)
__device__ __forceinline__ double f_acc(const VPar& q, double vi, double ki)
{   // forward step value sqrt(max(0, v^2 + 2 a_acc h)), main.cpp:830-831
    const double alat = vi * vi * fabs(ki);
    const double a_res = sqrt(fmax(0.0, q.a_tot2 - alat * alat));
    const double Fd = q.kd * vi * vi;
    double a_power = 1e9;
    if (q.has_power && vi > 1e-6) a_power = q.P / (q.mass * vi) - (Fd + q.Fr) * q.inv_mass;
    double a_acc = fmin(fmin(a_res, q.acc_cap), a_power);
    a_acc = fmax(0.0, a_acc);
    return sqrt(fmax(0.0, vi * vi + 2.0 * a_acc * q.h));
}
__device__ __forceinline__ double f_brk(const VPar& q, double vi, double ki)
{   // backward step value sqrt(max(0, v^2 + 2 a_brk h)), main.cpp:842-843
    const double alat = vi * vi * fabs(ki);
    const double a_res = sqrt(fmax(0.0, q.a_tot2 - alat * alat));
    const double Fd = q.kd * vi * vi;
    double a_brk = fmin(a_res, q.brk_cap) + (Fd + q.Fr) * q.inv_mass;
    a_brk = fmax(0.0, a_brk);
    return sqrt(fmax(0.0, vi * vi + 2.0 * a_brk * q.h));
}

// velocity_profile_forward_backward, main.cpp:782-862, on the blocked layout.
//
// The reference sweeps are sequential recurrences v[i+1] = min(v[i+1], f(v[i])).  Here every thread
// runs the recurrence over its own chunk from the value its neighbour published last round, and the
// rounds repeat until no published value changes.  Each round recomputes from the values at the START
// of the sweep (v0), so the fixed point is the unique solution of u[i+1] = min(v0[i+1], f(u[i])) --
// exactly the sequential sweep, for any f (f is not monotone near the friction limit, so a running
// minimum would NOT be exact).  The two closed-loop wrap updates (main.cpp:834-839, 846-850) stay single
// post-sweep updates.  An iteration that changes nothing ends the loop early (the map is deterministic).
// sX: 6*T doubles of exchange space.  Returns v[] (blocked); *rounds += relaxation rounds.
template <int T, int K>
__device__ __forceinline__ void vprofile_blocked(const Part& pt, const VPar& q, const double (&kap)[K], double (&v)[K],
                                                 int max_iters, double* sX, int& rounds, bool closed, double (&vkap)[K])
{
    double* sVL = sX;           // [2][T] last-slot value of each thread
    double* sVF = sX + 2 * T;   // [2][T] first-slot value
    double* sKF = sX + 4 * T;   // [T] kappa of first slot
    double* sKL = sX + 5 * T;   // [T] kappa of last slot
    const int cnt = pt.cnt, tid = pt.tid;
    const bool act = cnt > 0;
    const bool hasL = act && tid > 0, hasR = act && tid < pt.Tact - 1;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        vkap[k] = (k < cnt) ? sqrt(q.a_lat_max / fmax(fabs(kap[k]), q.kappa_eps)) : 0.0;   // also the v_kappa of the time weights (main.cpp:956)
        v[k] = (k < cnt) ? fmin(q.v_cap, vkap[k]) : 0.0;                                  // main.cpp:787-794
    }
    {
        double kl = kap[0];
#pragma unroll
        for (int k = 1; k < K; ++k) if (k < cnt) kl = kap[k];
        sKF[tid] = kap[0]; sKL[tid] = kl;
    }
    block_sync<T>();
    const double kapL = sKL[pt.tL], kapR = sKF[pt.tR];
    int b = 0;
    for (int iter = 0; iter < max_iters; ++iter) {
        double v0[K];
        // "did this iteration change anything": every update is a min with the old value, so v[k] at the end differs from
        // v[k] at the start exactly when the forward or the backward sweep (or a wrap update) lowered it
        bool chg_iter = false;
        // ---------------- forward sweep (main.cpp:829-833) ----------------
#pragma unroll
        for (int k = 0; k < K; ++k) v0[k] = v[k];
        {
            double last = v[0];
#pragma unroll
            for (int k = 1; k < K; ++k) if (k < cnt) last = v[k];
            sVL[b * T + tid] = last;
        }
        block_sync<T>();
        // A round recomputes the chunk from the start-of-sweep values and the neighbour's published value only: when that
        // input is the one of the previous round, so is the output.  After the first round the recurrences (two square
        // roots and a division per step, all dependent) therefore only run in the warps the change is passing through.
        bool first_round = true;
        double vin_prev = 0.0, last = 0.0;
        for (;;) {
            bool changed = false;
            if (act) {
                const double vin = sVL[b * T + pt.tL];
                if (first_round || vin != vin_prev) {
                    double u = hasL ? fmin(v0[0], f_acc(q, vin, kapL)) : v0[0];
                    v[0] = u;
#pragma unroll
                    for (int k = 1; k < K; ++k)
                        if (k < cnt) { u = fmin(v0[k], f_acc(q, u, kap[k - 1])); v[k] = u; }
                    last = u;
                    vin_prev = vin;
                }
                changed = (last != sVL[b * T + tid]);
            }
            first_round = false;
            sVL[(b ^ 1) * T + tid] = last;
            b ^= 1;
            ++rounds;
            if (!block_or<T>(changed)) break;
        }
        // closed-loop wrap: v[0] = min(v[0], f_acc(v[N-1])), main.cpp:834-839
        if (closed && tid == 0) v[0] = fmin(v[0], f_acc(q, sVL[b * T + pt.tL], kapL));
#pragma unroll
        for (int k = 0; k < K; ++k) if (k < cnt && v[k] != v0[k]) chg_iter = true;
        // ---------------- backward sweep (main.cpp:841-845) ----------------
#pragma unroll
        for (int k = 0; k < K; ++k) v0[k] = v[k];
        sVF[b * T + tid] = v[0];
        block_sync<T>();
        first_round = true;
        double first = 0.0;
        for (;;) {
            bool changed = false;
            if (act) {
                const double vin = sVF[b * T + pt.tR];
                if (first_round || vin != vin_prev) {
                    double u = 0.0;
#pragma unroll
                    for (int k = K - 1; k >= 0; --k) {
                        if (k == cnt - 1) { u = hasR ? fmin(v0[k], f_brk(q, vin, kapR)) : v0[k]; v[k] = u; }
                        else if (k < cnt - 1) { u = fmin(v0[k], f_brk(q, u, kap[k + 1])); v[k] = u; }
                    }
                    first = u;
                    vin_prev = vin;
                }
                changed = (first != sVF[b * T + tid]);
            }
            first_round = false;
            sVF[(b ^ 1) * T + tid] = first;
            b ^= 1;
            ++rounds;
            if (!block_or<T>(changed)) break;
        }
        // closed-loop wrap: v[N-1] = min(v[N-1], f_brk(v[0])), main.cpp:846-850
        if (closed && tid == pt.Tact - 1) {
            const double w = f_brk(q, sVF[b * T + pt.tR], kapR);
#pragma unroll
            for (int k = 0; k < K; ++k) if (k == cnt - 1) v[k] = fmin(v[k], w);
        }
#pragma unroll
        for (int k = 0; k < K; ++k) if (k < cnt && v[k] != v0[k]) chg_iter = true;
        if (!block_or<T>(chg_iter)) break;
    }
}

// ax and lap time, main.cpp:854-860.  Returns the block-wide lap time; ax[] per owned slot.
template <int T, int K>
__device__ __forceinline__ double lap_and_ax(const Part& pt, const VPar& q, const double (&v)[K], double (&ax)[K],
                                             double* sX, double* sred, bool closed)
{
    double* sVF = sX;
    sVF[pt.tid] = v[0];
    block_sync<T>();
    const double vnext_edge = sVF[pt.tR];
    double t = 0.0, dummy = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        ax[k] = 0.0;
        if (k < pt.cnt) {
            const double v0 = v[k];
            double v1 = (!closed && pt.tid == pt.Tact - 1) ? v0 : vnext_edge;   // open: j = i at the last sample (main.cpp:856)
            if (k + 1 < K) { if (k + 1 < pt.cnt) v1 = v[k + 1]; }
            ax[k] = (v1 * v1 - v0 * v0) / (2.0 * q.h);
            t += q.h / fmax(1e-6, v0);
        }
    }
    block_sum2<T>(t, dummy, sred, pt.lane, pt.warp);
    return t;
}

// ---- cost/gradient of the frozen problem on one thread's window --------------------------------------
// eval_cost_grad_frozen / _timeweighted, main.cpp:654-675 / 866-895, restated per sample with the time
// weight gamma folded into the coefficients:
//   zhat_i = c0_i + cp_i a[i+1] + cm_i a[i-1] - (cp_i+cm_i) a[i]        (= gamma_i * W_i (N0 + A1 D1a + A2 D2a)_i)
//   J      = sum zhat^2 + lamJ sum d^2,    d_i = a[i+1]-a[i-1]          (lamJ = lambda/(2h)^2)
//   g_i/2  = cp_{i-1} zhat_{i-1} + cm_{i+1} zhat_{i+1} - (cp_i+cm_i) zhat_i + lamJ (d_{i-1}-d_{i+1})
// x: trial alpha of the K owned slots; halos hl0,hl1 (samples start-2,start-1) and hr0,hr1 (start+cnt, +1).
// Returns gh = grad/2 and accumulates Jz = sum zhat^2, Sd = sum d^2 over the OWNED slots.
struct Halo { double l0, l1, r0, r1; };

// Shared-memory layout of two per-sample constants that are always read together -- (cp, cm) and (lo, hi) --
// slot-major [k][t].  Packed (default): the pair shares one 16-byte slot, one LDS.128 instead of two LDS.64.
// -DRL_NOPACK128 keeps the two planes apart (the layout of round 1, kept for A/B measurements).
#ifdef RL_NOPACK128
constexpr bool kPack = false;
#else
constexpr bool kPack = true;
#endif
// `region` holds 2*NP doubles; base pointers of thread t's column
__device__ __forceinline__ double* pair_base_a(double* region, int NP, int t) { return kPack ? region + 2 * t : region + t; }
__device__ __forceinline__ double* pair_base_b(double* region, int NP, int t) { return kPack ? region + 2 * t + 1 : region + NP + t; }
template <int T>
__device__ __forceinline__ void ld_pair(const double* a, const double* b, int k, double& x, double& y)
{
    if (kPack) { const double2 v = reinterpret_cast<const double2*>(a)[k * T]; x = v.x; y = v.y; }
    else { x = a[k * T]; y = b[k * T]; }
}
template <int T>
__device__ __forceinline__ void st_pair(double* a, double* b, int k, double x, double y)
{
    if (kPack) reinterpret_cast<double2*>(a)[k * T] = make_double2(x, y);
    else { a[k * T] = x; b[k * T] = y; }
}

template <int T, int K, int MODE>
__device__ __forceinline__ void eval_window(const double (&x)[K], const Halo& hh, const Part& pt,
                                            const double* __restrict__ sC0, const double* __restrict__ sCp,
                                            const double* __restrict__ sCm, const double (&cL)[3], const double (&cR)[3],
                                            double lamJ, double& Jz, double& Sd, double (&gh)[K])
{
    constexpr bool EXACT = (MODE == kModeExact), OPEN = (MODE == kModeOpen);
    const int cnt = pt.cnt;
    // open tracks (DiffOpsOpen, main.cpp:560-579): the stencils are one-sided at the two ends.  With ghost
    // values a[-1] := a[0], a[N] := a[N-1], zero stencil coefficients on the ghosts and end coefficients built by
    // the kernel, zhat keeps the closed form; the smoothing term gets weight 4 on the end samples
    // ((a1-a0)/h = 2 d/(2h)) and ghost terms e[-1] := -e[0], e[N] := -e[N-1], which is exactly D1^T D1.
    const bool endL = OPEN && (pt.tid == 0) && pt.first_chunk, endR = OPEN && (pt.tid == pt.Tact - 1) && pt.last_chunk;
    double w[K + 4];
    w[0] = hh.l0; w[1] = hh.l1;
#pragma unroll
    for (int k = 0; k < K; ++k) w[2 + k] = x[k];
    w[K + 2] = hh.r0; w[K + 3] = hh.r1;
    if (OPEN) {
        if (endL) { w[0] = x[0]; w[1] = x[0]; }
    }
    if (!EXACT) {
#pragma unroll
        for (int k = 1; k < K; ++k)
            if (cnt == k) { w[2 + k] = (OPEN && endR) ? x[k - 1] : hh.r0; w[3 + k] = (OPEN && endR) ? x[k - 1] : hh.r1; }
        if (OPEN && endR && cnt == K) { w[K + 2] = x[K - 1]; w[K + 3] = x[K - 1]; }
    }
    double z[K + 2], s[K + 2], d[K + 2], pp[K + 2], mm[K + 2];
#pragma unroll
    for (int p = 0; p < K + 2; ++p) {
        double c0, cp, cm;
        if (p == 0) { c0 = cL[0]; cp = cL[1]; cm = cL[2]; }
        else if (p == K + 1) { c0 = cR[0]; cp = cR[1]; cm = cR[2]; }
        else { c0 = sC0[(p - 1) * T]; ld_pair<T>(sCp, sCm, p - 1, cp, cm); }
        const double wm = w[p], wc = w[p + 1], wp = w[p + 2];
        s[p] = cp + cm;
        z[p] = fma(cp, wp, fma(cm, wm, fma(-s[p], wc, c0)));
        d[p] = wp - wm;
        pp[p] = cp * z[p];
        mm[p] = cm * z[p];
    }
    if (OPEN) {
        // e[p] = weight * d[p] (weight relative to lamJ); positions: p = 0 is sample start-1, p = k+1 is slot k
        double e[K + 2];
#pragma unroll
        for (int p = 0; p < K + 2; ++p) {
            const bool is_end = (endL && p == 1) || (endR && p == cnt);
            e[p] = is_end ? 4.0 * d[p] : d[p];
        }
        if (endL) e[0] = -e[1];
#pragma unroll
        for (int p = 2; p < K + 2; ++p) if (endR && p == cnt + 1) e[p] = -e[p - 1];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int p = k + 1;
            double t = pp[p - 1] + mm[p + 1];
            t = fma(-s[p], z[p], t);
            gh[k] = fma(lamJ, e[p - 1] - e[p + 1], t);
            if (k < cnt) {
                Jz = fma(z[p], z[p], Jz);
                Sd = fma(e[p], d[p], Sd);
            }
        }
        return;
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int p = k + 1;
        double t = pp[p - 1] + mm[p + 1];
        t = fma(-s[p], z[p], t);
        gh[k] = fma(lamJ, d[p - 1] - d[p + 1], t);
        if (EXACT || k < cnt) {
            Jz = fma(z[p], z[p], Jz);
            Sd = fma(d[p], d[p], Sd);
        }
    }
}

// first two / last two owned samples of a thread (what its neighbours need as halos)
template <int K, int MODE>
__device__ __forceinline__ void edge_values(const double (&x)[K], int cnt, double& F0, double& F1, double& L0, double& L1)
{
    constexpr bool EXACT = (MODE == kModeExact);
    F0 = x[0];
    F1 = x[1];
    if (EXACT) { L0 = x[K - 2]; L1 = x[K - 1]; }
    else {
        L0 = x[0]; L1 = x[1];
#pragma unroll
        for (int k = 2; k < K; ++k)
            if (k < cnt) { L0 = x[k - 1]; L1 = x[k]; }
        if (cnt <= 1) { F1 = x[0]; L0 = x[0]; L1 = x[0]; }  // N == 1: every neighbour is the sample itself
    }
}

// shuffle part of the halo exchange + publication of the warp-edge values (before the barrier)
template <int T, int K, int MODE>
__device__ __forceinline__ Halo halo_send(const double (&x)[K], const Part& pt, double* sExF, double* sExL)
{
    double F0, F1, L0, L1;
    edge_values<K, MODE>(x, pt.cnt, F0, F1, L0, L1);
    Halo h;
    h.l0 = __shfl_sync(kFull, L0, pt.srcL);
    h.l1 = __shfl_sync(kFull, L1, pt.srcL);
    h.r0 = __shfl_sync(kFull, F0, pt.srcR);
    h.r1 = __shfl_sync(kFull, F1, pt.srcR);
    if (T > 32) {
        // predicated stores: no divergent branch (BSSY / BSYNC and its branch-resolving bubble) in the hot loop
        asm volatile("{\n.reg .pred p;\nsetp.eq.s32 p, %0, 0;\n@p st.shared.v2.f64 [%1], {%2, %3};\n}" ::"r"(pt.lane),
                     "r"(smem_u32(sExF + 2 * pt.warp)), "d"(F0), "d"(F1) : "memory");
        asm volatile("{\n.reg .pred p;\nsetp.eq.s32 p, %0, 31;\n@p st.shared.v2.f64 [%1], {%2, %3};\n}" ::"r"(pt.lane),
                     "r"(smem_u32(sExL + 2 * pt.warp)), "d"(L0), "d"(L1) : "memory");
    }
    return h;
}
// after the barrier: warp-edge lanes pick their halos up from shared memory
template <int T>
__device__ __forceinline__ void halo_recv(Halo& h, const Part& pt, const double* sExF, const double* sExL)
{
    if (T > 32) {
        constexpr int NW = T / 32;
        if (pt.lane == 0) { const int wp = (pt.warp == 0) ? NW - 1 : pt.warp - 1; h.l0 = sExL[2 * wp]; h.l1 = sExL[2 * wp + 1]; }
        if (pt.lane == 31) { const int wn = (pt.warp == NW - 1) ? 0 : pt.warp + 1; h.r0 = sExF[2 * wn]; h.r1 = sExF[2 * wn + 1]; }
    }
}

struct PgdOut { double J0, Jend; int acc, bt, ev; };
// size classes up to this thread count run one copy of the half step of the projected-gradient loop (see pgd_outer)
#ifndef RL_SINGLE_COPY_MAX_T
#define RL_SINGLE_COPY_MAX_T 128
#endif
constexpr int kSingleCopyMaxT = RL_SINGLE_COPY_MAX_T;

// std::min(hi, std::max(lo, a)) of main.cpp:731 as two compare-selects (fmin/fmax cost twice as much in SASS)
__device__ __forceinline__ double clamp_box(double a, double lo, double hi)
{
    a = (lo < a) ? a : lo;
    return (a < hi) ? a : hi;
}

// Projection of the next trial onto the box (main.cpp:731): xb = min(hi, max(lo, xa - step*grad)); returns
// sum gh * (xb - xa) (main.cpp:733 for the NEXT Armijo test, in units of grad/2).
template <int T, int K>
__device__ __forceinline__ double project_trial(const double (&xa)[K], const double (&gh)[K], double step2,
                                                const double* sLo, const double* sHi, double (&xb)[K])
{
    double dec2p = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double lo, hi;
        ld_pair<T>(sLo, sHi, k, lo, hi);
        const double xn = clamp_box(fma(-step2, gh[k], xa[k]), lo, hi);
        dec2p = fma(gh[k], xn - xa[k], dec2p);
        xb[k] = xn;
    }
    return dec2p;
}

// State shared by the two halves of the ping-pong loop below.
template <int K>
struct PgdCtx {
    const double* sC0; const double* sCp; const double* sCm; double* sSt;
    double* sRed; double* sExF; double* sExL;
    double lamJ, armijo_c, step2, J, decp;
    int ph;
};

// Evaluate the trial xa (halos ha), form the next trial xb = clamp(xa - step*g) speculatively, exchange its
// halos and reduce (J(xa), g_prev.(xa - a)) through ONE barrier, then take the Armijo decision (main.cpp:734).
// Returns true when xa is accepted: the stash then holds xa and xb/hb is the next trial.
template <int T, int K, int MODE>
__device__ __forceinline__ bool pgd_half(const Part& pt, const double (&xa)[K], const Halo& ha, double (&xb)[K], Halo& hb,
                                         const double* sLo, const double* sHi, const double (&cL)[3],
                                         const double (&cR)[3], PgdCtx<K>& c)
{
    double gh[K];
    double Jz = 0.0, Sd = 0.0;
    eval_window<T, K, MODE>(xa, ha, pt, c.sC0, c.sCp, c.sCm, cL, cR, c.lamJ, Jz, Sd, gh);
    double Jn = fma(c.lamJ, Sd, Jz);
    const double dec2p = project_trial<T, K>(xa, gh, c.step2, sLo, sHi, xb);
    hb = halo_send<T, K, MODE>(xb, pt, c.sExF + c.ph * 32, c.sExL + c.ph * 32);
    double dec = c.decp;
    block_sum2<T>(Jn, dec, c.sRed + c.ph * 32, pt.lane, pt.warp);
    halo_recv<T>(hb, pt, c.sExF + c.ph * 32, c.sExL + c.ph * 32);
    c.ph ^= 1;
    dec *= 2.0;                                    // gh is grad/2: back to grad . (a_new - alpha), main.cpp:733
    if (Jn <= c.J + c.armijo_c * dec) {            // Armijo accept, main.cpp:734
#pragma unroll
        for (int k = 0; k < K; ++k) c.sSt[k * T] = xa[k];
        c.decp = dec2p; c.J = Jn;
        return true;
    }
    return false;
}

// One outer iteration's projected-gradient loop with Armijo backtracking (main.cpp:723-742 / 996-1026).
// In: coefficients and box bounds in shared memory (sC0/sCp/sCm/sLo/sHi, already offset by tid), cL/cR halo
// coefficients.
// Out: the accepted alpha of the owned slots in the stash sSt[k*T] (also already offset by tid).
//
// Steady state costs ONE barrier per evaluation: after the gradient of trial x is known, the next trial
// y = clamp(x - step*g) is formed speculatively (it is the reference's next trial whenever x is accepted,
// and x is accepted ~98% of the time), its halos are exchanged and (J(x), g_prev.(x-a)) are reduced through
// the same barrier.  The trial alternates between two register arrays (x, y) so that an accept moves no data.
// On a reject the step halves and the trial is rebuilt from the stashed accepted alpha.
template <int T, int K, int MODE>
__device__ __forceinline__ PgdOut pgd_outer(const Part& pt, const double* sLo, const double* sHi,
                                            const double (&cL)[3], const double (&cR)[3],
                                            const double* sC0, const double* sCp, const double* sCm, double* sSt,
                                            double* sRed, double* sExF, double* sExL, int& ph,
                                            double lamJ, double step_init, double step_min, double armijo_c, int max_inner)
{
    PgdOut o; o.acc = 0; o.bt = 0; o.ev = 0;
    PgdCtx<K> c;
    c.sC0 = sC0; c.sCp = sCp; c.sCm = sCm; c.sSt = sSt; c.sRed = sRed; c.sExF = sExF; c.sExL = sExL;
    c.lamJ = lamJ; c.armijo_c = armijo_c; c.ph = ph;
    c.step2 = 2.0 * step_init;   // gh is grad/2, so alpha - step*grad = alpha - step2*gh
    double x[K], y[K];
    Halo hx, hy;
    hx.l0 = hx.l1 = hx.r0 = hx.r1 = 0.0;
    hy = hx;
#pragma unroll
    for (int k = 0; k < K; ++k) { x[k] = 0.0; y[k] = 0.0; sSt[k * T] = 0.0; }
    {
        // ---- J and gradient at alpha = 0 (main.cpp:724 / 997), first trial into x ----
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
        hx = halo_send<T, K, MODE>(x, pt, sExF + c.ph * 32, sExL + c.ph * 32);
        double zero = 0.0;
        block_sum2<T>(Jt, zero, sRed + c.ph * 32, pt.lane, pt.warp);
        halo_recv<T>(hx, pt, sExF + c.ph * 32, sExL + c.ph * 32);
        c.ph ^= 1;
        c.J = Jt; c.decp = decp;
        o.J0 = Jt;
    }
    double Jprev = c.J;
    int it = 0, bt = 0;
    bool in_x = true;    // which array holds the current trial
    while (it < max_inner) {
        // Small CTAs: a dozen one-warp CTAs (or seven two-warp, four four-warp ones) sit on an SM, each somewhere else in the
        // code, and the loop is bound by instruction fetch (one-warp class, ncu: no_instruction = 2/3 of the stall samples).
        // They run ONE copy of the half step and move the new trial back (24 register moves) instead of alternating
        // between two copies: half the loop's code (+20 % at T = 32, +9 % at T = 64, +8 % at T = 128).
        bool acc;
        if (T <= kSingleCopyMaxT) {
            acc = pgd_half<T, K, MODE>(pt, x, hx, y, hy, sLo, sHi, cL, cR, c);
            if (acc) {
#pragma unroll
                for (int k = 0; k < K; ++k) x[k] = y[k];
                hx = hy;
            }
        } else
            acc = in_x ? pgd_half<T, K, MODE>(pt, x, hx, y, hy, sLo, sHi, cL, cR, c)
                       : pgd_half<T, K, MODE>(pt, y, hy, x, hx, sLo, sHi, cL, cR, c);
        o.ev++;
        if (acc) {
            if (T > kSingleCopyMaxT) in_x = !in_x;
            o.acc++; it++; bt = 0;
            if (fabs(Jprev - c.J) < 1e-10) break;       // main.cpp:740
            Jprev = c.J;
        } else {
            c.step2 *= 0.5; bt++; o.bt++;               // main.cpp:737
            if (0.5 * c.step2 < step_min || bt >= 20) break;   // not accepted: leave the inner loop (main.cpp:739)
            // rebuild the trial (into x) from the accepted alpha; its gradient is recomputed (rejects are rare)
            double a[K], gh[K];
#pragma unroll
            for (int k = 0; k < K; ++k) a[k] = sSt[k * T];
            Halo ha = halo_send<T, K, MODE>(a, pt, sExF + c.ph * 32, sExL + c.ph * 32);
            block_sync<T>();
            halo_recv<T>(ha, pt, sExF + c.ph * 32, sExL + c.ph * 32);
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
            hx = halo_send<T, K, MODE>(x, pt, sExF + c.ph * 32, sExL + c.ph * 32);
            block_sync<T>();
            halo_recv<T>(hx, pt, sExF + c.ph * 32, sExL + c.ph * 32);
            c.ph ^= 1;
            in_x = true;
        }
    }
    ph = c.ph;
    o.Jend = c.J;
    return o;
}



// ---- corridor, general path (rings of any size, streamed through shared memory in tiles) ------------------
// (main.cpp:694-711, 749-756): rays +-n against both rings, exact nearest hit
// Consecutive mapping (sample i = tid + j*T) so that a warp's 32 rays are neighbours and share boxes.
// The +n and -n rays lie on one line: den, t, u of the -n ray are exactly -den, -t, u of the +n ray
// (IEEE negation is exact), so one intersection serves both.  A box can only contain a hit if the
// line crosses it, which prunes the brute-force loop of rayToRingDistance (main.cpp:491-500) without
// changing its result.  A ring a ray misses entirely falls back to the nearest point-segment distance
// (minDistanceToSegments_global, main.cpp:501-512), pruned by box lower bounds.
// Results are returned in the BLOCKED layout through region B.
template <int T, int K>
__device__ __forceinline__ void corridor_build_tiled(const Part& pt, const double2* sP, double* sB, uint64_t* mbar, uint32_t& bar_phase, bool closed,
                                               const double* __restrict__ gseg, long long segI0, long long segO0, long long segE,
                                               double guard, double (&lo)[K], double (&hi)[K], long long& ray_tests)
{
    constexpr int NP = T * K;
    constexpr int CAP = ((4 * NP * 8) / (32 + 32 / SB)) / SB * SB;   // segments per tile (+ one box per SB segments)
    double* sSeg = sB;
    double* sBox = sB + 4 * CAP;
    const int N = pt.N, tid = pt.tid;
    const double INF = dinf();
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
        double pos[K], neg[K], dmin2[K];
#pragma unroll
        for (int j = 0; j < K; ++j) { pos[j] = INF; neg[j] = INF; dmin2[j] = INF; }
        const int ntiles = (mr + CAP - 1) / CAP;
        for (int pass = 0; pass < 2; ++pass) {
            if (pass == 1) {
                bool need = false;
#pragma unroll
                for (int j = 0; j < K; ++j) need = need || ((tid + j * T < N) && (pos[j] == INF || neg[j] == INF));
                if (!block_or<T>(need)) break;
            }
            for (int tile = 0; tile < ntiles; ++tile) {
                const int t0 = tile * CAP, nt = min(CAP, mr - t0), nblk = (nt + SB - 1) / SB;
                if (pass == 0 || ntiles > 1) {
                    block_sync<T>();
                    if (tid == 0) {
                        fence_proxy_async();
                        mbar_expect_tx(mbar, (uint32_t)nt * 32u);
                        bulk_g2s(sSeg, gseg + 4 * (base + t0), (uint32_t)nt * 32u, mbar);
                    }
                    mbar_wait(mbar, bar_phase); bar_phase ^= 1;
                    // boxes; segments become (x0, y0, vx, vy) with vx = x1-x0 as in main.cpp:482
                    for (int bq = tid; bq < nblk; bq += T) {
                        double xmin = INF, xmax = -INF, ymin = INF, ymax = -INF;
                        const int e = min(nt, bq * SB + SB);
                        for (int s = bq * SB; s < e; ++s) {
                            const double x0 = sSeg[4 * s], y0 = sSeg[4 * s + 1], x1 = sSeg[4 * s + 2], y1 = sSeg[4 * s + 3];
                            xmin = fmin(xmin, fmin(x0, x1)); xmax = fmax(xmax, fmax(x0, x1));
                            ymin = fmin(ymin, fmin(y0, y1)); ymax = fmax(ymax, fmax(y0, y1));
                            sSeg[4 * s + 2] = x1 - x0; sSeg[4 * s + 3] = y1 - y0;
                        }
                        const double cx = 0.5 * (xmin + xmax), cy = 0.5 * (ymin + ymax);
                        sBox[4 * bq] = cx; sBox[4 * bq + 1] = cy;
                        sBox[4 * bq + 2] = 0.5 * (xmax - xmin) + 1e-9 + 1e-12 * fabs(cx);
                        sBox[4 * bq + 3] = 0.5 * (ymax - ymin) + 1e-9 + 1e-12 * fabs(cy);
                    }
                    block_sync<T>();
                }
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    const int i = tid + j * T;
                    if (i >= N) continue;
                    const double2 Pc = sP[i];
                    if (pass == 0) {
                        double nx, ny;
                        normal_at(sP, i, N, closed, nx, ny);
                        double bp = pos[j], bn = neg[j];
                        for (int bq = 0; bq < nblk; ++bq) {
                            const double cx = sBox[4 * bq], cy = sBox[4 * bq + 1], hx = sBox[4 * bq + 2], hy = sBox[4 * bq + 3];
                            const double sc = nx * (cy - Pc.y) - ny * (cx - Pc.x);
                            const double ext = fabs(nx) * hy + fabs(ny) * hx + 1e-7;
                            if (fabs(sc) <= ext) {
                                const int e = min(nt, bq * SB + SB);
                                for (int s = bq * SB; s < e; ++s) {
                                    const double x0 = sSeg[4 * s], y0 = sSeg[4 * s + 1], vx = sSeg[4 * s + 2], vy = sSeg[4 * s + 3];
                                    const double den = nx * (-vy) + ny * vx;            // main.cpp:483
                                    const double ax = x0 - Pc.x, ay = y0 - Pc.y;        // main.cpp:485
                                    const double un = nx * ay - ny * ax;
                                    const double aden = fabs(den);
                                    const double us = (den > 0.0) ? un : -un;
                                    // division-free prefilter, strictly wider than the exact test below
                                    if (aden >= 1e-15 && us >= -2e-12 * aden && us <= aden + 2e-12 * aden) {
                                        const double inv = 1.0 / den;
                                        const double t = (ax * (-vy) + ay * vx) * inv;   // main.cpp:486
                                        const double u = un * inv;                       // main.cpp:487
                                        ++ray_tests;
                                        if (u >= -1e-12 && u <= 1.0 + 1e-12) {           // main.cpp:488
                                            if (t > 0.0) bp = fmin(bp, t);               // +n ray, main.cpp:497
                                            else if (t < 0.0) bn = fmin(bn, -t);         // -n ray (t' = -t)
                                        }
                                    }
                                }
                            }
                        }
                        pos[j] = bp; neg[j] = bn;
                    } else if (pos[j] == INF || neg[j] == INF) {
                        // nearest point-segment distance to this ring (main.cpp:501-512), box-pruned
                        const double ub = fmin(pos[j], neg[j]);   // a hit point lies on the ring
                        double bound2 = (ub < INF) ? ub * ub * (1.0 + 1e-9) : INF;
                        bound2 = fmin(bound2, dmin2[j] * (1.0 + 1e-9));
                        double best2 = dmin2[j];
                        for (int bq = 0; bq < nblk; ++bq) {
                            const double cx = sBox[4 * bq], cy = sBox[4 * bq + 1], hx = sBox[4 * bq + 2], hy = sBox[4 * bq + 3];
                            const double ddx = fmax(0.0, fabs(cx - Pc.x) - hx), ddy = fmax(0.0, fabs(cy - Pc.y) - hy);
                            if (ddx * ddx + ddy * ddy <= bound2) {
                                const int e = min(nt, bq * SB + SB);
                                for (int s = bq * SB; s < e; ++s) {
                                    const double x0 = sSeg[4 * s], y0 = sSeg[4 * s + 1], vx = sSeg[4 * s + 2], vy = sSeg[4 * s + 3];
                                    const double apx = Pc.x - x0, apy = Pc.y - y0;
                                    const double denom = fmax(1e-30, vx * vx + vy * vy);
                                    const double tt = fmin(1.0, fmax(0.0, (vx * apx + vy * apy) / denom));
                                    const double qx = x0 + vx * tt, qy = y0 + vy * tt;
                                    const double ex = Pc.x - qx, ey = Pc.y - qy;
                                    const double d2 = ex * ex + ey * ey;
                                    best2 = fmin(best2, d2);
                                    bound2 = fmin(bound2, d2 * (1.0 + 1e-9));
                                }
                            }
                        }
                        dmin2[j] = best2;
                    }
                }
            }
        }
        // safe_ray + min over the two rings (main.cpp:704-705)
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const double dist = (dmin2[j] < INF) ? sqrt(dmin2[j]) : 0.0;
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
        const int i = tid + j * T;
        if (i < N) {
            double hv = fmax(0.0, dpos[j] - guard);
            double lv = -fmax(0.0, dneg[j] - guard);
            if (!isfinite(hv)) hv = 0.0;
            if (!isfinite(lv)) lv = 0.0;
            sHiS[i] = hv; sLoS[i] = lv;
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

// segments of one ring that fit region B together with their FP32 copy, boxes and super boxes
__host__ __device__ constexpr int fast_tile_cap(int np)
{
    return (int)(((long long)32 * np * SB * SU) / (48 * SB * SU + 16 * SU + 16)) / (SB * SU) * (SB * SU);
}

// ---- corridor, fast path (both rings fit one tile each) ---------------------------------------------------
// Same results as corridor_build_tiled / the reference loops, far fewer tests:
//  * rays, not lines: a box is visited only if it meets the +n or -n RAY inside [0, best hit so far]; the
//    bound is shared by both rings because only min(inner, outer) enters the corridor (main.cpp:704-705);
//  * a two-level box hierarchy (SB segments per box, SB boxes per super box) tested in FP32 with
//    outward-rounded boxes and a margin that covers every rounding error; FP32 only ever SKIPS work, every
//    accepted hit comes from the FP64 formulas of main.cpp:478-490;
//  * near-first order: the scan starts at the box where this sample's nearest hit was last time (hints live
//    in shared memory across the outer iterations), so far-side crossings are pruned by the bound instead
//    of being descended divergently;
//  * "does this ray hit the ring AT ALL" (needed because a ring without any hit falls back to the
//    point-ring distance, main.cpp:696) is settled without a search when the ring is a closed chain and the
//    sample lies inside it (parity), a bit computed once per job: the path never crosses a ring between
//    corridor builds because |alpha| <= hit distance - guard (main.cpp:707-708, guard >= 0).
struct RayTile {
    const double* segD;   // [nt][4] x0,y0,vx,vy
    const float4* segF;   // [nt]    x0,y0,x1,y1 relative to the job origin
    const float4* boxF;   // [nblk]  cx,cy,hx,hy (inflated)
    const float4* supF;   // [nsup]
    int nt, nblk, nsup;
};

// The searches below are real (noinline) functions, and a pointer that crosses a call is a GENERIC pointer: its loads
// become LD.E (long scoreboard) although the tile lives in shared memory.  Rebuilding the pointers from the CTA's
// extern shared array tells the compiler the address space again (LDS).
template <class P>
__device__ __forceinline__ const P* as_smem(const P* g)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    return reinterpret_cast<const P*>(smem_raw + (smem_u32(g) - smem_u32(smem_raw)));
}
__device__ __forceinline__ RayTile tile_in_smem(const RayTile& t)
{
    RayTile r = t;
    r.segD = as_smem(t.segD); r.segF = as_smem(t.segF); r.boxF = as_smem(t.boxF); r.supF = as_smem(t.supF);
    return r;
}

__device__ __forceinline__ bool ray_box(const float4 bx, float px, float py, float nx, float ny, float anx, float any,
                                        float m, float bp, float bn, bool wp, bool wn)
{
    const float dx = bx.x - px, dy = bx.y - py;
    const float sc = nx * dy - ny * dx;
    if (fabsf(sc) > anx * bx.w + any * bx.z + m) return false;     // the line misses the box
    const float tc = nx * dx + ny * dy;
    const float r = anx * bx.z + any * bx.w + m;
    const bool pos_ok = wp && (tc + r >= 0.f) && (tc - r <= bp);
    const bool neg_ok = wn && (tc - r <= 0.f) && (-tc - r <= bn);
    return pos_ok || neg_ok;
}

// scan one ring for one sample.  bp/bn: running nearest +n / -n hit over both rings (pruning bound);
// pos_r/neg_r: nearest hit found on THIS ring (INF if none inside the bound).  hint: box to start at.
// Returns the segment of the nearest hit (-1 if none).
// Arguments and results of the real (noinline) functions travel BY VALUE: a reference parameter (the tile description,
// the in/out bounds) lives in the caller's local memory and is read and written back through L1 at every call.
struct RayRes { double bp, bn, pos, neg; int seg, tests; };
__device__ __noinline__ RayRes ray_scan_v(const RayTile tl_in, double2 P, double nx, double ny, float px, float py, float fnx, float fny,
                                          float m, int hint, bool wp, bool wn, bool first_hit_only,
                                          double bp, double bn, double pos_r, double neg_r)
{
    const RayTile tl = tile_in_smem(tl_in);
    int tests = 0;
    const double INF = dinf();
    const float anx = fabsf(fnx), any = fabsf(fny);
    float bpf = (bp < INF) ? __double2float_ru(bp) : __int_as_float(0x7f800000);
    float bnf = (bn < INF) ? __double2float_ru(bn) : __int_as_float(0x7f800000);
    int best_box = -1;   // segment index of the nearest hit found
    double best_abs = INF;
    int sb0 = hint / SU;
    if (sb0 >= tl.nsup) sb0 = 0;
    for (int q = 0; q < tl.nsup; ++q) {
        int sb = sb0 + q;
        if (sb >= tl.nsup) sb -= tl.nsup;
        if (!ray_box(tl.supF[sb], px, py, fnx, fny, anx, any, m, bpf, bnf, wp, wn)) continue;
        const int b1 = min(tl.nblk, sb * SU + SU);
        for (int b = sb * SU; b < b1; ++b) {
            if (!ray_box(tl.boxF[b], px, py, fnx, fny, anx, any, m, bpf, bnf, wp, wn)) continue;
            const int s1 = min(tl.nt, b * SB + SB);
            for (int s = b * SB; s < s1; ++s) {
                const float4 f = tl.segF[s];
                const float sa = fnx * (f.y - py) - fny * (f.x - px), sbb = fnx * (f.w - py) - fny * (f.z - px);
                if (fminf(sa, sbb) > m || fmaxf(sa, sbb) < -m) continue;       // both ends strictly on one side
                const double x0 = tl.segD[4 * s], y0 = tl.segD[4 * s + 1], vx = tl.segD[4 * s + 2], vy = tl.segD[4 * s + 3];
                const double den = nx * (-vy) + ny * vx;                         // main.cpp:483
                ++tests;
                if (fabs(den) < 1e-15) continue;                                 // main.cpp:484
                const double ax = x0 - P.x, ay = y0 - P.y;                      // main.cpp:485
                const double inv = 1.0 / den;
                const double t = (ax * (-vy) + ay * vx) * inv;                  // main.cpp:486
                const double u = (nx * ay - ny * ax) * inv;                     // main.cpp:487
                if (u >= -1e-12 && u <= 1.0 + 1e-12) {                          // main.cpp:488
                    if (t > 0.0) {                                              // +n ray, main.cpp:497
                        if (wp) {
                            pos_r = fmin(pos_r, t);
                            if (t < bp) { bp = t; bpf = __double2float_ru(t); }
                            if (t < best_abs) { best_abs = t; best_box = s; }
                            if (first_hit_only) goto done;
                        }
                    } else if (t < 0.0) {                                       // -n ray: t' = -t
                        if (wn) {
                            neg_r = fmin(neg_r, -t);
                            if (-t < bn) { bn = -t; bnf = __double2float_ru(-t); }
                            if (-t < best_abs) { best_abs = -t; best_box = s; }
                            if (first_hit_only) goto done;
                        }
                    }
                }
            }
        }
    }
done:
    RayRes r;
    r.bp = bp; r.bn = bn; r.pos = pos_r; r.neg = neg_r; r.seg = best_box; r.tests = tests;
    return r;
}
// the in/out form the callers use (inlined: its references are the caller's registers)
__device__ __forceinline__ int ray_scan(const RayTile& tl, double2 P, double nx, double ny, float px, float py, float fnx, float fny,
                                        float m, int hint, bool wp, bool wn, bool first_hit_only,
                                        double& bp_io, double& bn_io, double& pos_io, double& neg_io, long long& tests_io)
{
    const RayRes r = ray_scan_v(tl, P, nx, ny, px, py, fnx, fny, m, hint, wp, wn, first_hit_only, bp_io, bn_io, pos_io, neg_io);
    bp_io = r.bp; bn_io = r.bn; pos_io = r.pos; neg_io = r.neg; tests_io += r.tests;
    return r.seg;
}

// nearest point-segment distance to the ring (minDistanceToSegments_global, main.cpp:501-512); ub: any known
// upper bound (a hit point lies on the ring) or INF.
__device__ __noinline__ double dist_scan(const RayTile tl_in, double2 P, float px, float py, float m, int hint, double ub)
{
    const RayTile tl = tile_in_smem(tl_in);
    const double INF = dinf();
    double best2 = INF;
    float boundf = (ub < INF) ? __double2float_ru(ub) * (1.f + 1e-5f) + m : __int_as_float(0x7f800000);
    int sb0 = hint / SU;
    if (sb0 >= tl.nsup) sb0 = 0;
    for (int q = 0; q < tl.nsup; ++q) {
        int sb = sb0 + q;
        if (sb >= tl.nsup) sb -= tl.nsup;
        {
            const float4 bx = tl.supF[sb];
            const float ddx = fmaxf(0.f, fabsf(bx.x - px) - bx.z - m), ddy = fmaxf(0.f, fabsf(bx.y - py) - bx.w - m);
            if (ddx * ddx + ddy * ddy > boundf * boundf) continue;
        }
        const int b1 = min(tl.nblk, sb * SU + SU);
        for (int b = sb * SU; b < b1; ++b) {
            const float4 bx = tl.boxF[b];
            const float ddx = fmaxf(0.f, fabsf(bx.x - px) - bx.z - m), ddy = fmaxf(0.f, fabsf(bx.y - py) - bx.w - m);
            if (ddx * ddx + ddy * ddy > boundf * boundf) continue;
            const int s1 = min(tl.nt, b * SB + SB);
            for (int s = b * SB; s < s1; ++s) {
                {   // FP32 lower bound of the point-segment distance
                    const float4 f = tl.segF[s];
                    const float vx = f.z - f.x, vy = f.w - f.y, apx = px - f.x, apy = py - f.y;
                    const float tt = fminf(1.f, fmaxf(0.f, __fdividef(vx * apx + vy * apy, fmaxf(1e-30f, vx * vx + vy * vy))));
                    const float ex = apx - vx * tt, ey = apy - vy * tt;
                    const float bq = boundf + 4.f * m;
                    if (ex * ex + ey * ey > bq * bq) continue;
                }
                const double x0 = tl.segD[4 * s], y0 = tl.segD[4 * s + 1], vx = tl.segD[4 * s + 2], vy = tl.segD[4 * s + 3];
                const double apx = P.x - x0, apy = P.y - y0;
                const double denom = fmax(1e-30, vx * vx + vy * vy);
                const double tt = fmin(1.0, fmax(0.0, (vx * apx + vy * apy) / denom));
                const double qx = x0 + vx * tt, qy = y0 + vy * tt;
                const double ex = P.x - qx, ey = P.y - qy;
                const double d2 = ex * ex + ey * ey;
                if (d2 < best2) {
                    best2 = d2;
                    const float df = sqrtf(__double2float_ru(d2)) * (1.f + 1e-5f) + m;
                    boundf = fminf(boundf, df);
                }
            }
        }
    }
    return (best2 < INF) ? sqrt(best2) : INF;
}

// parity of the crossings of the +x ray from P with a CLOSED chain (vertices shared bit for bit): P inside?
__device__ __noinline__ bool inside_ring(const RayTile tl_in, double2 P, float px, float py, float m)
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
                const int sn = (s + 1 == tl.nt) ? 0 : s + 1;
                const double ax = tl.segD[4 * s], ay = tl.segD[4 * s + 1], bxx = tl.segD[4 * sn], byy = tl.segD[4 * sn + 1];
                if ((ay > P.y) != (byy > P.y)) {
                    const double xc = ax + (P.y - ay) * (bxx - ax) / (byy - ay);
                    if (xc > P.x) ++cnt;
                }
            }
        }
    }
    return (cnt & 1) != 0;
}

// load one ring into region B and build its FP32 filter structures.  Returns the FP32 margin; *closed tells
// whether the segments form a closed chain (end of j == start of j+1, bit for bit).
template <int T, int K>
__device__ __forceinline__ float ring_tile_build(const Part& pt, double* sB, uint64_t* mbar, uint32_t& bar_phase, int* sMisc,
                                                 const double* __restrict__ gseg, int mr, double ox, double oy,
                                                 RayTile& tl, bool& closed, bool* linked = nullptr, bool dbg = false)
{
    (void)dbg;
    constexpr int NP = T * K;
    constexpr int CAP = fast_tile_cap(NP);
    const int tid = pt.tid;
    double* segD = sB;
    float4* segF = reinterpret_cast<float4*>(sB + 4 * CAP);
    float4* boxF = segF + CAP;
    float4* supF = boxF + CAP / SB;
    tl.segD = segD; tl.segF = segF; tl.boxF = boxF; tl.supF = supF;
    tl.nt = mr; tl.nblk = (mr + SB - 1) / SB; tl.nsup = (tl.nblk + SU - 1) / SU;
    block_sync<T>();
    if (dbg) { RL_DBG_ENTER(T, sMisc, kDbgPhTile); RL_DBG_ASSERT(sMisc, mr <= CAP, 2001); }
    if (tid == 0) {
        sMisc[0] = 0;
        fence_proxy_async();
        mbar_expect_tx(mbar, (uint32_t)mr * 32u);
        bulk_g2s(segD, gseg, (uint32_t)mr * 32u, mbar);
    }
    mbar_wait(mbar, bar_phase); bar_phase ^= 1;
    block_sync<T>();
    float emax = 0.f;
    int ok = 1, lk = 1;
    for (int s = tid; s < mr; s += T) {
        const double x0 = segD[4 * s], y0 = segD[4 * s + 1], x1 = segD[4 * s + 2], y1 = segD[4 * s + 3];
        const int sn = (s + 1 == mr) ? 0 : s + 1;
        const bool same = (x1 == segD[4 * sn] && y1 == segD[4 * sn + 1]);
        ok &= same;
        if (s + 1 < mr) lk &= same;
        float4 f;
        f.x = (float)(x0 - ox); f.y = (float)(y0 - oy); f.z = (float)(x1 - ox); f.w = (float)(y1 - oy);
        segF[s] = f;
        emax = fmaxf(emax, fmaxf(fmaxf(fabsf(f.x), fabsf(f.y)), fmaxf(fabsf(f.z), fabsf(f.w))));
    }
    atomicMax(&sMisc[0], __float_as_int(emax));
    closed = (T == 32) ? (__all_sync(kFull, ok) != 0) : (__syncthreads_and(ok) != 0);
    if (linked) *linked = (T == 32) ? (__all_sync(kFull, lk) != 0) : (__syncthreads_and(lk) != 0);
    block_sync<T>();
    // second pass: (x1,y1) -> (vx,vy) once every thread has compared the shared vertices
    for (int s = tid; s < mr; s += T) {
        segD[4 * s + 2] -= segD[4 * s]; segD[4 * s + 3] -= segD[4 * s + 1];   // vx = x1-x0, vy = y1-y0 (main.cpp:482)
    }
    const float m = 2e-6f * __int_as_float(sMisc[0]) + 1e-6f;
    for (int b = tid; b < tl.nblk; b += T) {
        float xmin = 3e38f, xmax = -3e38f, ymin = 3e38f, ymax = -3e38f;
        const int e = min(mr, b * SB + SB);
        for (int s = b * SB; s < e; ++s) {
            const float4 f = segF[s];
            xmin = fminf(xmin, fminf(f.x, f.z)); xmax = fmaxf(xmax, fmaxf(f.x, f.z));
            ymin = fminf(ymin, fminf(f.y, f.w)); ymax = fmaxf(ymax, fmaxf(f.y, f.w));
        }
        float4 bx;
        bx.x = 0.5f * (xmin + xmax); bx.y = 0.5f * (ymin + ymax);
        bx.z = 0.5f * (xmax - xmin) * (1.f + 1e-6f) + m; bx.w = 0.5f * (ymax - ymin) * (1.f + 1e-6f) + m;
        boxF[b] = bx;
    }
    block_sync<T>();
    for (int sb = tid; sb < tl.nsup; sb += T) {
        float xmin = 3e38f, xmax = -3e38f, ymin = 3e38f, ymax = -3e38f;
        const int e = min(tl.nblk, sb * SU + SU);
        for (int b = sb * SU; b < e; ++b) {
            const float4 bx = boxF[b];
            xmin = fminf(xmin, bx.x - bx.z); xmax = fmaxf(xmax, bx.x + bx.z);
            ymin = fminf(ymin, bx.y - bx.w); ymax = fmaxf(ymax, bx.y + bx.w);
        }
        float4 sx;
        sx.x = 0.5f * (xmin + xmax); sx.y = 0.5f * (ymin + ymax);
        sx.z = 0.5f * (xmax - xmin) * (1.f + 1e-6f) + m; sx.w = 0.5f * (ymax - ymin) * (1.f + 1e-6f) + m;
        supF[sb] = sx;
    }
    block_sync<T>();
    return m;
}

// ---- per-sample corridor state (persists across the outer iterations, in shared memory) ------------------
// sHint[i]: bits 0-12 / 13-25 anchor segment j0 of ring 0 / 1 (the segment of the sample's nearest hit when the
//           state was taken), bits 26/27 sample inside ring 0/1 (parity, once per job), bits 28/29 state valid.
// sClr[i] : two bytes, clearance of ring 0 / 1 in units of 1/4 m, rounded down and measured at the CENTRE-LINE
//           position C0 of the sample: every segment of the ring OUTSIDE the index window [j0-W, j0+W] is at
//           least that far from C0.  For a later path position P with |P - C0| = disp, every outside segment is
//           at least Rc = clearance - disp away, so
//             * a window hit with t <= Rc is the nearest hit of the whole ring,
//             * no window hit means the ring has no hit with t < Rc,
//             * a window point-segment distance <= Rc is the point-ring distance,
//           and nothing but the 2W+1 window segments has to be touched.  When a certificate fails the general
//           box-hierarchy search runs for that sample and refreshes its state.
constexpr int kWin = 3;

__device__ __forceinline__ int wrap_seg(int s, int M) { return (s < 0) ? s + M : ((s >= M) ? s - M : s); }

// exact ray/segment test of main.cpp:478-490 for both rays of the line; updates the nearest +n / -n hits
__device__ __forceinline__ void seg_hit(const double* sd, double2 P, double nx, double ny, double& pos, double& neg, int s,
                                        int& seg_p, int& seg_n, long long& tests)
{
    const double x0 = sd[0], y0 = sd[1], vx = sd[2], vy = sd[3];
    const double den = nx * (-vy) + ny * vx;                         // main.cpp:483
    ++tests;
    if (fabs(den) < 1e-15) return;                                   // main.cpp:484
    const double ax = x0 - P.x, ay = y0 - P.y;                      // main.cpp:485
    const double inv = 1.0 / den;
    const double t = (ax * (-vy) + ay * vx) * inv;                  // main.cpp:486
    const double u = (nx * ay - ny * ax) * inv;                     // main.cpp:487
    if (u >= -1e-12 && u <= 1.0 + 1e-12) {                          // main.cpp:488
        if (t > 0.0) { if (t < pos) { pos = t; seg_p = s; } }       // +n ray, main.cpp:497
        else if (t < 0.0) { if (-t < neg) { neg = -t; seg_n = s; } } // -n ray: t' = -t
    }
}

// nearest +n / -n hits among the window segments [j0-W, j0+W] of the ring
__device__ __forceinline__ void window_rays(const RayTile& tl, double2 P, double nx, double ny, float px, float py, float fnx,
                                            float fny, float m, int j0, double& pos, double& neg, int& seg_p, int& seg_n,
                                            long long& tests)
{
    const int M = tl.nt;
    const int cntw = min(M, 2 * kWin + 1);
    int s = (M > 2 * kWin + 1) ? wrap_seg(j0 - kWin, M) : 0;
    for (int q = 0; q < cntw; ++q) {
        const float4 f = tl.segF[s];
        const float sa = fnx * (f.y - py) - fny * (f.x - px), sbb = fnx * (f.w - py) - fny * (f.z - px);
        if (!(fminf(sa, sbb) > m || fmaxf(sa, sbb) < -m)) seg_hit(tl.segD + 4 * s, P, nx, ny, pos, neg, s, seg_p, seg_n, tests);
        s = (s + 1 == M) ? 0 : s + 1;
    }
}

// exact distance from P to segment s (minDistanceToSegments_global body, main.cpp:504-509), squared
__device__ __forceinline__ double seg_dist2(const double* sd, double2 P)
{
    const double x0 = sd[0], y0 = sd[1], vx = sd[2], vy = sd[3];
    const double apx = P.x - x0, apy = P.y - y0;
    const double denom = fmax(1e-30, vx * vx + vy * vy);
    const double tt = fmin(1.0, fmax(0.0, (vx * apx + vy * apy) / denom));
    const double ex = P.x - (x0 + vx * tt), ey = P.y - (y0 + vy * tt);
    return ex * ex + ey * ey;
}
// FP32 lower bound material: squared distance from p to the FP32 copy of a segment
__device__ __forceinline__ float seg_dist2_f(const float4 f, float px, float py)
{
    const float vx = f.z - f.x, vy = f.w - f.y, apx = px - f.x, apy = py - f.y;
    const float tt = fminf(1.f, fmaxf(0.f, __fdividef(vx * apx + vy * apy, fmaxf(1e-30f, vx * vx + vy * vy))));
    const float ex = apx - vx * tt, ey = apy - vy * tt;
    return ex * ex + ey * ey;
}

// nearest point-segment distance among the window segments
__device__ __forceinline__ double window_dist(const RayTile& tl, double2 P, float px, float py, float m, int j0, double ub)
{
    const double INF = dinf();
    const int M = tl.nt;
    const int cntw = min(M, 2 * kWin + 1);
    int s = (M > 2 * kWin + 1) ? wrap_seg(j0 - kWin, M) : 0;
    double best2 = INF;
    float boundf = (ub < INF) ? __double2float_ru(ub) * (1.f + 1e-5f) + m : 3e18f;
    for (int q = 0; q < cntw; ++q) {
        const float bq = boundf + 4.f * m;
        if (seg_dist2_f(tl.segF[s], px, py) <= bq * bq) {
            const double d2 = seg_dist2(tl.segD + 4 * s, P);
            if (d2 < best2) { best2 = d2; boundf = fminf(boundf, sqrtf(__double2float_ru(d2)) * (1.f + 1e-5f) + m); }
        }
        s = (s + 1 == M) ? 0 : s + 1;
    }
    return (best2 < INF) ? sqrt(best2) : INF;
}

// conservative (rounded down) distance from p to every segment of the ring outside the window of j0
// s_off / m_ring: the tile holds segments [s_off, s_off + nt) of a ring of m_ring segments (a streamed ring); j0 is ring-global
__device__ __noinline__ float clearance_scan(const RayTile tl_in, float px, float py, float m, int j0, int s_off = 0, int m_ring = -1)
{
    const RayTile tl = tile_in_smem(tl_in);
    const int M = (m_ring < 0) ? tl.nt : m_ring;
    if (M <= 2 * kWin + 1) return 3e18f;
    // clearances are stored in 1/4 m up to 63.75 m: nothing farther away than that can matter, which also lets the
    // scan skip every distant box from the start (a streamed ring has its near segments in one tile of many)
    float best = 64.f;    // distance, not squared
    int sb0 = ((j0 - s_off) / SB) / SU;
    if (sb0 >= tl.nsup || sb0 < 0) sb0 = 0;
    for (int q = 0; q < tl.nsup; ++q) {
        int sb = sb0 + q;
        if (sb >= tl.nsup) sb -= tl.nsup;
        {
            const float4 bx = tl.supF[sb];
            const float ddx = fmaxf(0.f, fabsf(bx.x - px) - bx.z), ddy = fmaxf(0.f, fabsf(bx.y - py) - bx.w);
            if (ddx * ddx + ddy * ddy >= best * best) continue;
        }
        const int b1 = min(tl.nblk, sb * SU + SU);
        for (int b = sb * SU; b < b1; ++b) {
            const float4 bx = tl.boxF[b];
            const float ddx = fmaxf(0.f, fabsf(bx.x - px) - bx.z), ddy = fmaxf(0.f, fabsf(bx.y - py) - bx.w);
            if (ddx * ddx + ddy * ddy >= best * best) continue;
            const int s1 = min(tl.nt, b * SB + SB);
            for (int s = b * SB; s < s1; ++s) {
                int dj = s + s_off - j0; if (dj < 0) dj += M;
                if (dj <= kWin || dj >= M - kWin) continue;              // window segment
                const float d = sqrtf(seg_dist2_f(tl.segF[s], px, py)) - 4.f * m;
                best = fminf(best, fmaxf(d, 0.f));
            }
        }
    }
    return best;
}

// ---- existence certificates ---------------------------------------------------------------------------------
// "Does the ray that points AWAY from a ring hit that ring anywhere?" decides whether safe_ray falls back to the
// point-ring distance (main.cpp:696), and the answer is needed again at every corridor build for about half the
// samples.  It is stable between builds, so a full search leaves a certificate per sample (16 bytes, kept in the
// job's heading and curvature rows until the final geometry pass overwrites them):
//   FAR(f)      the ray hit segment f: re-test that one segment exactly; a hit anywhere is a hit;
//   CONE(A,d,c) the ray missed: the ring has no point inside the cone with apex A (on the ray's line, between the ring
//               behind the sample and the sample), axis d (the ray direction then) and cos(half-angle) = c.  A later ray
//               whose origin lies in the cone and whose direction is within the half-angle of d stays inside the
//               (convex) cone: still a miss.
// A certificate that does not apply falls back to the full search (and is rebuilt), so results never change.
constexpr unsigned kCertNone = 0u, kCertFar = 1u, kCertCone = 2u, kCertNoCone = 3u;

// can the box be skipped: no point of it is seen from the apex under a smaller angle to the axis than acos(cbest)
__device__ __forceinline__ bool cone_prune(const float4 bx, float ax, float ay, float dx, float dy, float cbest)
{
    const float cx = bx.x - ax, cy = bx.y - ay;
    const float R2 = bx.z * bx.z + bx.w * bx.w;       // boxes are already inflated by the FP32 margin
    const float r2 = cx * cx + cy * cy;
    if (r2 <= R2 * 1.001f) return false;              // apex inside (or at) the box's disk
    const float ir = rsqrtf(r2);
    const float ct = (cx * dx + cy * dy) * ir, so = sqrtf(R2) * ir * (1.f + 1e-6f);
    const float co = sqrtf(fmaxf(0.f, 1.f - so * so));
    if (ct >= co - 1e-5f) return false;               // the axis passes through (or close to) the disk
    const float st = sqrtf(fmaxf(0.f, 1.f - ct * ct));
    return ct * co + st * so + 1e-5f <= cbest;        // cos(theta_c - omega), rounded up
}

// largest cosine of the angle between the unit axis (dx,dy) and any point of the ring seen from the apex (ax,ay)
// (tile coordinates), rounded UP.  Returns 2 when no cone exists (a segment crosses the axis ahead of the apex, or
// FP32 cannot tell).  Along a segment that does not cross the forward axis the angle is extremal at its end points.
__device__ __noinline__ float cone_scan(const RayTile tl_in, float ax, float ay, float dx, float dy, float m, int hint)
{
    const RayTile tl = tile_in_smem(tl_in);
    const float e = 2.f * m + 1e-6f;
    // a cone wider than 20 degrees is never needed (and never wider than 90: only a CONVEX cone keeps
    // origin + t*direction inside), so everything outside the 20-degree cone is skipped from the start
    float cbest = 0.9396926f;
    int sb0 = hint / SU;
    if (sb0 >= tl.nsup) sb0 = 0;
    for (int q = 0; q < tl.nsup; ++q) {
        int sb = sb0 + q;
        if (sb >= tl.nsup) sb -= tl.nsup;
        if (cone_prune(tl.supF[sb], ax, ay, dx, dy, cbest)) continue;
        const int b1 = min(tl.nblk, sb * SU + SU);
        for (int b = sb * SU; b < b1; ++b) {
            if (cone_prune(tl.boxF[b], ax, ay, dx, dy, cbest)) continue;
            const int s1 = min(tl.nt, b * SB + SB);
            for (int s = b * SB; s < s1; ++s) {
                const float4 f = tl.segF[s];
                const float ux = f.x - ax, uy = f.y - ay, wx = f.z - ax, wy = f.w - ay;
                const float t1 = ux * dx + uy * dy, c1 = ux * dy - uy * dx;    // along / across the axis
                const float t2 = wx * dx + wy * dy, c2 = wx * dy - wy * dx;
                const float r1 = sqrtf(ux * ux + uy * uy), r2 = sqrtf(wx * wx + wy * wy);
                if (r1 <= 8.f * e || r2 <= 8.f * e) return 2.f;               // apex (almost) on the ring
                if (!((c1 > e && c2 > e) || (c1 < -e && c2 < -e))) {
                    // the segment meets the axis LINE: fine only if it does so clearly behind the apex
                    const float ds = c1 - c2;
                    if (fabsf(ds) < 16.f * e) return 2.f;
                    const float tc = t1 + (t2 - t1) * __fdividef(c1, ds);
                    if (tc > -e * (4.f + 2.f * __fdividef(fabsf(t2 - t1), fabsf(ds)))) return 2.f;
                }
                const float k1 = __fdividef(t1 + e, r1) + 2e-6f, k2 = __fdividef(t2 + e, r2) + 2e-6f;
                cbest = fmaxf(cbest, fmaxf(k1, k2));
            }
        }
    }
    return cbest;
}

struct ExQuery {
    double2 P; double nx, ny, cx0, cy0, ox, oy, guard, toward;
    unsigned long long cert, apex;      // the sample's certificate words (loaded early by the caller)
    float px, py, m;
    int j0, ring, dir, i, mr;
};
// unit axis of a CONE certificate from its two 16-bit components (the same bits at creation and at every check)
__device__ __forceinline__ void cert_axis(unsigned w1, double& dx, double& dy)
{
    const float fx = (float)(short)(w1 & 0xffffu), fy = (float)(short)(w1 >> 16);
    const float rs = rsqrtf(fmaxf(1.f, fx * fx + fy * fy));
    dx = (double)(fx * rs); dy = (double)(fy * rs);
}
// does the ray of direction `dir` (0: +n, 1: -n) from the sample hit the ring at all?  (certificates: see above)
// certificate word: w0 = state[1:0] | ring[2] | dir[3] | cos(half-angle) [31:16];  w1 = cone axis (2 x int16) or the
// far-hit segment; apex word: the cone apex relative to the centre-line sample, 2 floats.
struct ExRes { int tests, scans; bool ex; };
__device__ __noinline__ ExRes far_hit_exists_v(const RayTile tl, const ExQuery& qy, unsigned long long* __restrict__ gcert,
                                               unsigned long long* __restrict__ gapex)
{
    long long tests_io = 0;
    int scans_io = 0;
    ExRes res; res.tests = 0; res.scans = 0; res.ex = false;
    const double INF = dinf();
    const int i = qy.i, dir = qy.dir;
    const double2 Pc = qy.P;
    const double nx = qy.nx, ny = qy.ny;
    const unsigned w0 = (unsigned)qy.cert, w1 = (unsigned)(qy.cert >> 32);
    const unsigned key = ((unsigned)qy.ring << 2) | ((unsigned)dir << 3);
    const bool mine = ((w0 & 0xCu) == key) && ((w0 & 3u) != kCertNone);
    const double sg = dir ? -1.0 : 1.0;
    long long tests = 0;
    if (mine && (w0 & 3u) == kCertFar) {
        const int f = (int)w1;
        if (f < qy.mr) {
            double tp = INF, tn = INF; int s_p = -1, s_n = -1;
            seg_hit(tl.segD + 4 * f, Pc, nx, ny, tp, tn, f, s_p, s_n, tests);
            if ((dir ? tn : tp) < INF) { res.tests = (int)tests; res.ex = true; return res; }
        }
    } else if (mine && (w0 & 3u) == kCertCone) {
        double d0x, d0y;
        cert_axis(w1, d0x, d0y);
        const double cc = (double)(w0 >> 16) * (1.0 / 32767.0) - 1.0 + 2e-6;
        const double ux = (Pc.x - qy.cx0) - (double)__uint_as_float((unsigned)qy.apex);
        const double uy = (Pc.y - qy.cy0) - (double)__uint_as_float((unsigned)(qy.apex >> 32));
        if (ux * d0x + uy * d0y >= sqrt(ux * ux + uy * uy) * cc && sg * (nx * d0x + ny * d0y) >= cc) return res;
    }
    // ---- no certificate applies: full search, then leave a certificate for the next builds ----
    double ubp = INF, ubn = INF, pr = INF, nr = INF;
    const int hs = ray_scan(tl, Pc, nx, ny, qy.px, qy.py, (float)nx, (float)ny, qy.m, qy.j0 / SB, dir == 0, dir == 1, true, ubp, ubn, pr, nr, tests);
    tests_io += tests;
    ++scans_io;
    const bool ex = dir ? (nr < INF) : (pr < INF);
    if (!mine && (w0 & 3u) != kCertNone) {
        // the slot serves another (ring, direction) of this sample: leave it alone
    } else if (ex) {
        gcert[i] = ((unsigned long long)(unsigned)hs << 32) | (unsigned long long)(key | kCertFar);
    } else if (!(mine && (w0 & 3u) == kCertNoCone)) {
        // a miss: look for a ring-free cone about the current ray, apex on the ray's line between the ring (behind the
        // sample) and every later path position (the path keeps at least `guard` away from the ring)
        const int qx = __double2int_rn(sg * nx * 32767.0), qyy = __double2int_rn(sg * ny * 32767.0);
        const unsigned nw1 = ((unsigned)qx & 0xffffu) | ((unsigned)qyy << 16);
        double d0x, d0y;
        cert_axis(nw1, d0x, d0y);
        const double back = (qy.toward < INF) ? fmax(0.0, qy.toward - fmax(0.05, 0.5 * qy.guard)) : 0.0;
        const float oxf = (float)((Pc.x - qy.cx0) - back * d0x), oyf = (float)((Pc.y - qy.cy0) - back * d0y);
        const double axd = qy.cx0 + (double)oxf, ayd = qy.cy0 + (double)oyf;     // the apex, exactly as every later check sees it
        const float cb = cone_scan(tl, (float)(axd - qy.ox), (float)(ayd - qy.oy), (float)d0x, (float)d0y, qy.m, qy.j0 / SB);
        unsigned nw0 = key | kCertNoCone;
        if (cb < 0.9995f) {
            const unsigned qc = (unsigned)ceilf((cb + 1.f) * 32767.f + 0.5f);
            if (qc < 65535u) nw0 = key | kCertCone | (qc << 16);
        }
        gcert[i] = ((unsigned long long)nw1 << 32) | (unsigned long long)nw0;
        gapex[i] = ((unsigned long long)__float_as_uint(oyf) << 32) | (unsigned long long)__float_as_uint(oxf);
    }
    res.tests = (int)tests_io; res.scans = scans_io; res.ex = ex;
    return res;
}
__device__ __forceinline__ bool far_hit_exists(const RayTile& tl, const ExQuery& qy, unsigned long long* __restrict__ gcert,
                                               unsigned long long* __restrict__ gapex, long long& tests_io, int& scans_io)
{
    const ExRes r = far_hit_exists_v(tl, qy, gcert, gapex);
    tests_io += r.tests; scans_io += r.scans;
    return r.ex;
}

template <int T, int K>
__device__ __forceinline__ void corridor_build_fast(const Part& pt, const double2* sP, double* sB, uint64_t* mbar, uint32_t& bar_phase,
                                                    int* sMisc, unsigned* sHint, unsigned short* sClr, bool first, bool parity_ok, bool closed,
                                                    const double* __restrict__ gseg, const double* __restrict__ gcenter,
                                                    unsigned long long* __restrict__ gcert, unsigned long long* __restrict__ gapex,
                                                    long long segI0, long long segO0, long long segE,
                                                    double guard, unsigned mask, double (&loc)[K], double (&hic)[K],
                                                    long long& ray_tests, int& ex_scans)
{
    // mask: bit j = sample tid + j*T is to be (re)built; results go to loc/hic[j] (CONSECUTIVE mapping), others untouched
    const int N = pt.N, tid = pt.tid;
    const double INF = dinf();
    const float FINF = __int_as_float(0x7f800000);
    const double2 org = sP[0];
    double bp[K], bn[K], dfp[K], dfn[K];
    float lbp[K], lbn[K];          // "a hit exists on some ring, value unknown but >= lb" (parity certificate)
    unsigned flagged = 0u;         // bit j: sample j must be redone by the general search
#pragma unroll 1
    for (int j = 0; j < K; ++j)
        if ((mask >> j) & 1u) { bp[j] = INF; bn[j] = INF; dfp[j] = INF; dfn[j] = INF; lbp[j] = FINF; lbn[j] = FINF; }
    // ring order: the ring the samples lie INSIDE of goes first (its far hits are settled by parity, and the
    // bounds it produces let the other ring skip most existence questions)
    const unsigned h0 = sHint[0];
    const int first_ring = (!first && ((h0 >> 27) & 1u) && !((h0 >> 26) & 1u)) ? 1 : 0;

    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1) {
            if (!block_or<T>(flagged != 0u)) break;
#pragma unroll 1
            for (int j = 0; j < K; ++j)
                if ((flagged >> j) & 1u) { bp[j] = INF; bn[j] = INF; dfp[j] = INF; dfn[j] = INF; lbp[j] = FINF; lbn[j] = FINF; }
        }
        for (int rr = 0; rr < 2; ++rr) {
            const int ring = rr ^ first_ring;
            const long long base = ring ? segO0 : segI0;
            const int mr = (int)(ring ? (segE - segO0) : (segO0 - segI0));
            if (mr == 0) {   // safe_ray on an empty ring returns 0 (main.cpp:696-698)
#pragma unroll 1
                for (int j = 0; j < K; ++j) { dfp[j] = 0.0; dfn[j] = 0.0; }
                continue;
            }
            RayTile tl;
            bool chain, linked;   // the ring's segments form a closed chain / consecutive segments share their vertex
            const float m0 = ring_tile_build<T, K>(pt, sB, mbar, bar_phase, sMisc, gseg + 4 * base, mr, org.x, org.y, tl, chain, &linked, true);
            if (first && tid == 0) {   // per-ring constants for corridor_update
                sMisc[8 + ring] = (chain ? 1 : 0) | (linked ? 2 : 0);
                sMisc[10 + ring] = __float_as_int(m0);
            }
            const int hshift = 13 * ring;
#pragma unroll 1
            for (int j = 0; j < K; ++j) {
                const int i = tid + j * T;
                if (i >= N || !((mask >> j) & 1u)) continue;
                if (pass == 1 && !((flagged >> j) & 1u)) continue;
                const double2 Pc = sP[i];
                double nx, ny;
                normal_at(sP, i, N, closed, nx, ny);
                const float px = (float)(Pc.x - org.x), py = (float)(Pc.y - org.y), fnx = (float)nx, fny = (float)ny;
                const float m = m0 + 2e-6f * fmaxf(fabsf(px), fabsf(py));
                unsigned hw = sHint[i];
                unsigned short cw = sClr[i];
                int j0 = (int)((hw >> hshift) & 0x1fffu);
                if (j0 >= mr) j0 = 0;
                const bool inside = parity_ok && chain && ((hw >> (26 + ring)) & 1u);
                // displacement from the centre-line position the clearance refers to
                const double cx0 = gcenter[2 * i], cy0 = gcenter[2 * i + 1];
                unsigned long long cert_w = 0ull, apex_w = 0ull;
                if (!first) { cert_w = gcert[i]; apex_w = gapex[i]; }   // issued early: only the existence branch reads them
                const float disp = __double2float_ru(sqrt((Pc.x - cx0) * (Pc.x - cx0) + (Pc.y - cy0) * (Pc.y - cy0))) * (1.f + 1e-6f);
                float Rc = 0.f;
                if (pass == 0 && !first && ((hw >> (28 + ring)) & 1u)) {
                    const unsigned cq = (cw >> (8 * ring)) & 0xffu;
                    Rc = 0.25f * (float)cq - disp - 4.f * m;
                }
                double pos_r = INF, neg_r = INF;
                int seg_p = -1, seg_n = -1;
                bool ex_p, ex_n;          // does the +n / -n ray hit this ring at all
                bool have_d = false;
                double dist_r = INF;       // point-ring distance, computed on demand
                bool general = !(Rc > 0.f);
                if (!general) {
                    // ---------------- certified window search ----------------
                    window_rays(tl, Pc, nx, ny, px, py, fnx, fny, m, j0, pos_r, neg_r, seg_p, seg_n, ray_tests);
                    const double Rcd = (double)Rc;
                    // directions that need the general search: a hit that is not certified nearest, or no hit,
                    // no parity and a common bound that does not make far hits irrelevant
                    const bool open_p = (pos_r > Rcd) && ((pos_r < INF) || (!inside && !(bp[j] <= Rcd)));
                    const bool open_n = (neg_r > Rcd) && ((neg_r < INF) || (!inside && !(bn[j] <= Rcd)));
                    if (pos_r <= Rcd) bp[j] = fmin(bp[j], pos_r);
                    if (neg_r <= Rcd) bn[j] = fmin(bn[j], neg_r);
                    ex_p = (pos_r < INF); ex_n = (neg_r < INF);
                    if (open_p || open_n) {
                        const bool bounded_p = (bp[j] < INF) || (pos_r < INF), bounded_n = (bn[j] < INF) || (neg_r < INF);
                        double bq = fmin(bp[j], pos_r), bm = fmin(bn[j], neg_r), pr2 = pos_r, nr2 = neg_r;
                        ray_scan(tl, Pc, nx, ny, px, py, fnx, fny, m, j0 / SB, open_p, open_n, false, bq, bm, pr2, nr2, ray_tests);
                        if (open_p) { pos_r = pr2; bp[j] = fmin(bp[j], bq); ex_p = (pos_r < INF) || (bounded_p && inside); }
                        if (open_n) { neg_r = nr2; bn[j] = fmin(bn[j], bm); ex_n = (neg_r < INF) || (bounded_n && inside); }
                        const bool redo_p = open_p && !ex_p && bounded_p, redo_n = open_n && !ex_n && bounded_n;
                        if (redo_p || redo_n) {   // pruned by a bound and no parity: look for any hit at all
                            double ubp = INF, ubn = INF, pr = INF, nr = INF;
                            if (redo_p) { ray_scan(tl, Pc, nx, ny, px, py, fnx, fny, m, j0 / SB, true, false, true, ubp, ubn, pr, nr, ray_tests); ex_p = (pr < INF); }
                            if (redo_n) { ray_scan(tl, Pc, nx, ny, px, py, fnx, fny, m, j0 / SB, false, true, true, ubp, ubn, pr, nr, ray_tests); ex_n = (nr < INF); }
                        }
                    }
                    // directions without a window hit that were not searched: the ring has no hit with t < Rc
#pragma unroll
                    for (int dir = 0; dir < 2; ++dir) {
                        const bool open_d = dir ? open_n : open_p;
                        const double hit = dir ? neg_r : pos_r;
                        if (open_d || hit < INF) continue;
                        double& bnd = dir ? bn[j] : bp[j];
                        bool& ex = dir ? ex_n : ex_p;
                        if (inside) {
                            ex = true;                                   // parity: it hits, somewhere beyond Rc
                            if (!(bnd <= Rcd)) { if (dir) lbn[j] = fminf(lbn[j], Rc); else lbp[j] = fminf(lbp[j], Rc); }
                        } else {
                            // far hits cannot lower the minimum (bnd <= Rc); whether one EXISTS only matters if the
                            // point-ring distance would (main.cpp:696)
                            if (!have_d) {
                                dist_r = window_dist(tl, Pc, px, py, m, j0, fmin(pos_r, neg_r));
                                if (!(dist_r <= Rcd)) dist_r = dist_scan(tl, Pc, px, py, m, j0 / SB, fmin(dist_r, fmin(pos_r, neg_r)));
                                have_d = true;
                            }
                            if (dist_r >= bnd) ex = true;                // irrelevant either way: treat as settled
                            else {
                                // ---- existence of a far hit: certificate first, full search otherwise ----
                                ExQuery qy;
                                qy.P = Pc; qy.nx = nx; qy.ny = ny; qy.px = px; qy.py = py; qy.m = m; qy.j0 = j0; qy.ring = ring; qy.dir = dir;
                                qy.i = i; qy.mr = mr; qy.cx0 = cx0; qy.cy0 = cy0; qy.ox = org.x; qy.oy = org.y;
                                qy.guard = guard; qy.toward = dir ? pos_r : neg_r; qy.cert = cert_w; qy.apex = apex_w;
                                ex = far_hit_exists(tl, qy, gcert, gapex, ray_tests, ex_scans);
                                cert_w = gcert[i]; apex_w = gapex[i];     // the other direction of this ring may ask next
                            }
                        }
                    }
                } else {
                    // ---------------- general search (first build, failed certificate, redo pass) ----------------
                    const bool bounded_p = (bp[j] < INF), bounded_n = (bn[j] < INF);
                    const int hs = ray_scan(tl, Pc, nx, ny, px, py, fnx, fny, m, j0 / SB, true, true, false, bp[j], bn[j], pos_r, neg_r, ray_tests);
                    ex_p = (pos_r < INF) || (bounded_p && inside);
                    ex_n = (neg_r < INF) || (bounded_n && inside);
                    const bool redo_p = !ex_p && bounded_p, redo_n = !ex_n && bounded_n;
                    if (redo_p || redo_n) {
                        double ubp = INF, ubn = INF, pr = INF, nr = INF;
                        if (redo_p) { ray_scan(tl, Pc, nx, ny, px, py, fnx, fny, m, j0 / SB, true, false, true, ubp, ubn, pr, nr, ray_tests); ex_p = (pr < INF); }
                        if (redo_n) { ray_scan(tl, Pc, nx, ny, px, py, fnx, fny, m, j0 / SB, false, true, true, ubp, ubn, pr, nr, ray_tests); ex_n = (nr < INF); }
                    }
                    // refresh the state of this sample for this ring
                    if (first) {
                        const bool in = chain && inside_ring(tl, Pc, px, py, m);
                        hw = (hw & ~(1u << (26 + ring))) | ((in ? 1u : 0u) << (26 + ring));
                    }
                    if (hs >= 0) {
                        j0 = hs;
                        const float c = clearance_scan(tl, px, py, m, j0) - disp;
                        const unsigned cq = (c >= 63.75f) ? 255u : (c > 0.f ? (unsigned)(c * 4.f) : 0u);
                        cw = (unsigned short)((cw & ~(0xffu << (8 * ring))) | (cq << (8 * ring)));
                        hw = (hw & ~(0x1fffu << hshift)) | ((unsigned)j0 << hshift) | (1u << (28 + ring));
                    } else {
                        hw &= ~(1u << (28 + ring));
                    }
                    sHint[i] = hw; sClr[i] = cw;
                    if (first && hs >= 0 && !(parity_ok && chain && ((hw >> (26 + ring)) & 1u))) {
                        // the certificate the update passes will ask for, made now: the ray WITHOUT a near hit on a ring the
                        // sample is not inside of.  (On demand it would be made in some later build, for a few samples at a
                        // time, each time at the price of two tile builds and a round of block barriers.)
                        const float rc0 = 0.25f * (float)((cw >> (8 * ring)) & 0xffu) - disp - 4.f * m;
                        const int dirq = (pos_r <= (double)rc0) ? ((neg_r <= (double)rc0) ? -1 : 1) : ((neg_r <= (double)rc0) ? 0 : -1);
                        if (dirq >= 0) {
                            ExQuery qy;
                            qy.P = Pc; qy.nx = nx; qy.ny = ny; qy.px = px; qy.py = py; qy.m = m; qy.j0 = j0; qy.ring = ring; qy.dir = dirq;
                            qy.i = i; qy.mr = mr; qy.cx0 = cx0; qy.cy0 = cy0; qy.ox = org.x; qy.oy = org.y;
                            qy.guard = guard; qy.toward = dirq ? pos_r : neg_r; qy.cert = gcert[i]; qy.apex = gapex[i];
                            far_hit_exists(tl, qy, gcert, gapex, ray_tests, ex_scans);
                        }
                    }
                }
                // a ring some ray misses entirely: nearest point-segment distance (main.cpp:696)
                if (!ex_p || !ex_n) {
                    if (!have_d) dist_r = dist_scan(tl, Pc, px, py, m, j0 / SB, fmin(pos_r, neg_r));
                    if (!ex_p) dfp[j] = fmin(dfp[j], dist_r);
                    if (!ex_n) dfn[j] = fmin(dfn[j], dist_r);
                }
            }
            RL_DBG_LEAVE(sMisc, kDbgPhTile);     // this thread is done with the ring tile in region B
        }
        if (pass == 0) {
            // a pending "exists, value >= lb" that could undercut the minimum found: redo that sample in full
#pragma unroll 1
            for (int j = 0; j < K; ++j) {
                if (tid + j * T >= N || !((mask >> j) & 1u)) continue;
                if ((double)lbp[j] < fmin(bp[j], dfp[j]) || (double)lbn[j] < fmin(bn[j], dfn[j])) flagged |= (1u << j);
            }
        }
    }
    // hi/lo (main.cpp:704-710)
#pragma unroll 1
    for (int j = 0; j < K; ++j) {
        const int i = tid + j * T;
        if (i < N && ((mask >> j) & 1u)) {
            const double dpos = fmin(bp[j], dfp[j]), dneg = fmin(bn[j], dfn[j]);
            double hv = fmax(0.0, fmax(0.0, dpos) - guard);
            double lv = -fmax(0.0, fmax(0.0, dneg) - guard);
            if (!isfinite(hv)) hv = 0.0;
            if (!isfinite(lv)) lv = 0.0;
            hic[j] = hv; loc[j] = lv;
        }
    }
}

// corridor bounds from the consecutive mapping (sample tid + j*T) to the blocked layout, through region B
template <int T, int K>
// The bounds stay in the staging area (sample-major lo | hi at the start of region B) until the next outer iteration
// has parked the path and their home in the path area is free (staged_bounds_home): they never occupy registers across
// the linearisation.
__device__ __forceinline__ void corridor_stage_out(const Part& pt, double* sB, const double (&loc)[K], const double (&hic)[K],
                                                   int* dbgMisc = nullptr)
{
    constexpr int NP = T * K;
    (void)dbgMisc;
    block_sync<T>();
    if (dbgMisc) RL_DBG_ENTER(T, dbgMisc, kDbgPhStage);
    double* sLoS = sB;
    double* sHiS = sB + NP;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const int i = pt.tid + j * T;
        if (i < pt.N) { sHiS[i] = hic[j]; sLoS[i] = loc[j]; }
    }
    block_sync<T>();
}
// corridor_build_tiled returns the bounds in registers (blocked layout): same staging area, same hand-over
template <int T, int K>
__device__ __forceinline__ void corridor_stage_blocked(const Part& pt, double* sB, const double (&lo)[K], const double (&hi)[K], int* dbgMisc = nullptr)
{
    constexpr int NP = T * K;
    (void)dbgMisc;
    block_sync<T>();
    if (dbgMisc) RL_DBG_ENTER(T, dbgMisc, kDbgPhStage);
#pragma unroll
    for (int k = 0; k < K; ++k)
        if (k < pt.cnt) { sB[pt.start + k] = lo[k]; sB[NP + pt.start + k] = hi[k]; }
    block_sync<T>();
}
// staging area -> (lo, hi) pairs of the owned slots in the (parked) path area; unused slots get (0, 0)
template <int T, int K>
__device__ __forceinline__ void staged_bounds_home(const Part& pt, const double* sB, double* sLo, double* sHi, int* dbgMisc = nullptr)
{
    constexpr int NP = T * K;
    (void)dbgMisc;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double l = 0.0, h = 0.0;
        if (k < pt.cnt) { l = sB[pt.start + k]; h = sB[NP + pt.start + k]; }
        st_pair<T>(sLo, sHi, k, l, h);
    }
    if (dbgMisc) RL_DBG_LEAVE(dbgMisc, kDbgPhStage);
    block_sync<T>();   // region B is free again
}

// ---- corridor update, common case (every build after the first) ------------------------------------------------
// When both rings are vertex chains, everything a sample needs in the steady state lies in the 2*kWin+1 segments around
// its anchor on each ring plus its certificates:
//   * a window hit with t <= Rc (Rc = stored clearance - displacement) is that ring's nearest hit;
//   * a ring without a window hit has no hit nearer than Rc: it cannot undercut a certified hit b <= Rc of the other
//     ring; whether it has a hit AT ALL only matters when its point distance D < b (main.cpp:696), and then parity
//     (sample inside a closed ring) or the existence certificate answers;
//   * D itself is the window distance when that is <= Rc, and otherwise D > Rc >= b does not matter.
// Both rings' vertices (16 B each) sit in region B together, so one pass over the samples does both rings with scalars
// only, no searches and no calls.  A sample any certificate fails for is FLAGGED and rebuilt by corridor_build_fast,
// which also refreshes its anchors / clearances / certificates.  Returns the flag mask (bit j = sample tid + j*T).
struct UpdCtx {
    // byte offsets into the CTA's dynamic shared memory (pointers rebuilt from the extern array keep the loads LDS)
    // (ring 1's arrays follow ring 0's: V1 = V0 + len0 + 1, F1 = F0 + len0 + 1)
    int oV0, oF0, oHint, oClr, oHalo;
    int N, M0, M1, rf0, rf1;
    // LOCAL form (a chunk of a long track): shared memory holds vertices [base, base + len] of each ring only; the
    // path has one halo point each side (oHalo); FAR segments outside the range come from global memory (gs0; ring 1's
    // records follow ring 0's: gs0 + 4 * M0)
    int base0, base1, len0, len1;
    float mr0, mr1;
    int parity_ok, closed;
    const double* gcenter; const unsigned long long* gcert; const unsigned long long* gapex; const double* gs0;
    double ox, oy, guard;
};
// it lives in static shared memory, which is rounded up to the 128-byte alignment of the dynamic part: the cluster kernel's
// two CTAs per SM have exactly that much room
static_assert(sizeof(UpdCtx) <= 128, "UpdCtx must fit one 128-byte unit of static shared memory");
// ring-global segment index -> index into the chunk-local vertex arrays
__device__ __forceinline__ int loc_idx(int sg, int base, int M) { const int li = sg - base; return (li < 0) ? li + M : li; }
// one sample of corridor_update: true = FLAGGED (hv/lv untouched), else the corridor bounds hv >= 0 >= lv
// CTA-uniform, so it lives in static shared memory: thread 0 fills it before the barrier that precedes the sample loop
// (as an argument it sat in local memory and every call read its fields back through L1)
__shared__ UpdCtx s_upd;
// result of one sample, returned BY VALUE (registers): reference out-parameters of a real function live in local memory
struct UpdRes { double hv, lv; int tests; bool flagged;
#ifdef RL_PHASE_TIMERS
    int why;   // development: which rule flagged the sample (bit 0 no state, 1 clearance used up, 2 window outside the local range,
               // 3 window hit beyond Rc, 4 no certified bound, 5 FAR certificate failed, 6 CONE certificate failed, 7 no certificate)
#endif
};
#ifdef RL_PHASE_TIMERS
#define RL_WHY(b) (why |= (1 << (b)))
#else
#define RL_WHY(b) ((void)0)
#endif
template <bool LOCAL>
__device__ __noinline__ UpdRes corridor_update_sample(int i, double cx0, double cy0, unsigned long long cert_w, unsigned long long apex_w)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const UpdCtx& c = s_upd;
    const double2* sP = reinterpret_cast<const double2*>(smem_raw);
    const double2* cV0 = reinterpret_cast<const double2*>(smem_raw + c.oV0);
    const double2* cV1 = cV0 + (c.len0 + 1);
    const float2* cF0 = reinterpret_cast<const float2*>(smem_raw + c.oF0);
    const float2* cF1 = cF0 + (c.len0 + 1);
    const double INF = dinf();
    long long ray_tests = 0;
    const double2 Pc = sP[i];
    double nx, ny;
    if (LOCAL) {   // the chunk's neighbours' end points live in the halo slots (closed track)
        const double2* hl = reinterpret_cast<const double2*>(smem_raw + c.oHalo);
        // c.closed: bit 0 closed track, bit 1 / 2: this chunk holds the first / last sample of an OPEN track (one-sided there)
        if ((c.closed & 2) && i == 0) normal_from_tangent(sP[1].x - sP[0].x, sP[1].y - sP[0].y, nx, ny);
        else if ((c.closed & 4) && i == c.N - 1) normal_from_tangent(sP[i].x - sP[i - 1].x, sP[i].y - sP[i - 1].y, nx, ny);
        else {
            const double2 Pm = (i == 0) ? hl[0] : sP[i - 1], Pp = (i == c.N - 1) ? hl[1] : sP[i + 1];
            normal_from_tangent((Pp.x - Pm.x) * 0.5, (Pp.y - Pm.y) * 0.5, nx, ny);
        }
    } else normal_at(sP, i, c.N, (c.closed & 1) != 0, nx, ny);
    const unsigned hw = reinterpret_cast<const unsigned*>(smem_raw + c.oHint)[i];
    const unsigned cw = reinterpret_cast<const unsigned short*>(smem_raw + c.oClr)[i];
    // upper bound of the sample's displacement from the centre line, in FP32: the components are rounded up, the few
    // ulps sqrtf / fmaf may lose are covered by the factor (the bound only has to be conservative)
    const float dfx = __double2float_ru(fabs(Pc.x - cx0)), dfy = __double2float_ru(fabs(Pc.y - cy0));
    const float disp = sqrtf(fmaf(dfx, dfx, dfy * dfy)) * (1.f + 2e-6f);
    const float px = (float)(Pc.x - c.ox), py = (float)(Pc.y - c.oy), fnx = (float)nx, fny = (float)ny;
    const float pmax = 2e-6f * fmaxf(fabsf(px), fabsf(py)) + 1e-5f;
    bool flag = false;
#ifdef RL_PHASE_TIMERS
    int why = 0;
#endif
    double w[2][2], Rc[2], wd[2];
    bool ins[2];
#pragma unroll
    for (int ring = 0; ring < 2; ++ring) {
        const int M = ring ? c.M1 : c.M0;
        const double2* V = ring ? cV1 : cV0;
        const float2* F = ring ? cF1 : cF0;
        const int rf = ring ? c.rf1 : c.rf0;
        const float m = (ring ? c.mr1 : c.mr0) * 1.001f + pmax;      // the FP32 margin the stored clearance was taken with
        int j0 = (int)((hw >> (13 * ring)) & 0x1fffu);
        if (j0 >= M) j0 = 0;
        const unsigned cq = (cw >> (8 * ring)) & 0xffu;
        const float rc = 0.25f * (float)cq - disp - 4.f * m;
        if (!((hw >> (28 + ring)) & 1u) || !(rc > 0.f)) { flag = true; RL_WHY(((hw >> (28 + ring)) & 1u) ? 1 : 0); }
        Rc[ring] = (double)rc;
        ins[ring] = c.parity_ok && (rf & 1) && ((hw >> (26 + ring)) & 1u);
        double pos = INF, neg = INF;
        int sp_ = -1, sn_ = -1;
        int sg = wrap_seg(j0 - kWin, M);
#pragma unroll 1
        for (int q = 0; q < 2 * kWin + 1; ++q) {
            // side of the ray's line each end point lies on (FP32, margin m): both clearly on one side -> no hit
            const int li = LOCAL ? loc_idx(sg, ring ? c.base1 : c.base0, M) : sg;
            if (LOCAL && li >= (ring ? c.len1 : c.len0)) { flag = true; RL_WHY(2); }         // window outside the chunk-local range
            else {
                const float2 fa = F[li], fb = F[li + 1];
                const float sa = fnx * (fa.y - py) - fny * (fa.x - px), sb = fnx * (fb.y - py) - fny * (fb.x - px);
                if (!(fminf(sa, sb) > m || fmaxf(sa, sb) < -m)) {
                    const double2 a = V[li], b = V[li + 1];
                    const double sd[4] = {a.x, a.y, b.x - a.x, b.y - a.y};      // v = b - a as in main.cpp:482
                    seg_hit(sd, Pc, nx, ny, pos, neg, sg, sp_, sn_, ray_tests);
                }
            }
            sg = (sg + 1 == M) ? 0 : sg + 1;
        }
        w[ring][0] = pos; w[ring][1] = neg;
        wd[ring] = -1.0;   // window distance not computed yet
    }
    double dres[2];
#pragma unroll
    for (int dir = 0; dir < 2; ++dir) {
        double b = INF;
#pragma unroll
        for (int ring = 0; ring < 2; ++ring) {
            if (w[ring][dir] <= Rc[ring]) b = fmin(b, w[ring][dir]);
            else if (w[ring][dir] < INF) { flag = true; RL_WHY(3); }           // a window hit that is not certified nearest
        }
        double res = b;
#pragma unroll
        for (int ring = 0; ring < 2; ++ring) {
            if (w[ring][dir] < INF) continue;
            if (!(b <= Rc[ring])) { flag = true; RL_WHY(4); continue; }    // nothing certified undercuts this ring's far hits
            if (ins[ring]) continue;                             // parity: it hits, beyond Rc >= b
            if (wd[ring] < 0.0) {
                // exact point-ring distance over the window (minDistanceToSegments_global body, main.cpp:504-509)
                const int M = ring ? c.M1 : c.M0;
                const double2* V = ring ? cV1 : cV0;
                const float2* F = ring ? cF1 : cF0;
                const float m = (ring ? c.mr1 : c.mr0) * 1.001f + pmax;
                int j0 = (int)((hw >> (13 * ring)) & 0x1fffu);
                if (j0 >= M) j0 = 0;
                const int sg0 = wrap_seg(j0 - kWin, M);
                // FP32 first: only segments within a few margins of the FP32 minimum can hold the exact minimum
                float dminf = 3e38f;
                int sg = sg0;
#pragma unroll 1
                for (int q = 0; q < 2 * kWin + 1; ++q) {
                    const int li = LOCAL ? loc_idx(sg, ring ? c.base1 : c.base0, M) : sg;
                    if (!(LOCAL && li >= (ring ? c.len1 : c.len0))) {          // (out of range: the hit loop above already flagged)
                        const float2 fa = F[li], fb = F[li + 1];
                        dminf = fminf(dminf, seg_dist2_f(make_float4(fa.x, fa.y, fb.x, fb.y), px, py));
                    }
                    sg = (sg + 1 == M) ? 0 : sg + 1;
                }
                // the exact window distance is at least the FP32 minimum less the margin: when even that cannot undercut
                // the certified hit b, the point distance does not matter (main.cpp:696) and is not computed
                if ((double)(sqrtf(dminf) * (1.f - 2e-6f) - 8.f * m) >= b) continue;
                const float lim = sqrtf(dminf) + 8.f * m;
                const float lim2 = lim * lim * (1.f + 1e-5f);
                double best2 = INF;
                sg = sg0;
#pragma unroll 1
                for (int q = 0; q < 2 * kWin + 1; ++q) {
                    const int li = LOCAL ? loc_idx(sg, ring ? c.base1 : c.base0, M) : sg;
                    if (!(LOCAL && li >= (ring ? c.len1 : c.len0))) {
                        const float2 fa = F[li], fb = F[li + 1];
                        if (seg_dist2_f(make_float4(fa.x, fa.y, fb.x, fb.y), px, py) <= lim2) {
                            const double2 a = V[li], bb = V[li + 1];
                            const double sd[4] = {a.x, a.y, bb.x - a.x, bb.y - a.y};
                            best2 = fmin(best2, seg_dist2(sd, Pc));
                        }
                    }
                    sg = (sg + 1 == M) ? 0 : sg + 1;
                }
                wd[ring] = sqrt(best2);
            }
            if (wd[ring] <= Rc[ring] && wd[ring] < b) {
                // the point distance would undercut: does this ray hit the ring anywhere?  (certificate or flag)
                const unsigned w0 = (unsigned)cert_w, w1 = (unsigned)(cert_w >> 32);
                const unsigned key = ((unsigned)ring << 2) | ((unsigned)dir << 3);
                const unsigned stt = ((w0 & 0xCu) == key) ? (w0 & 3u) : kCertNone;
                if (stt == kCertFar) {
                    const int M = ring ? c.M1 : c.M0;
                    const double2* V = ring ? cV1 : cV0;
                    const int f = (int)w1;
                    bool hit = false;
                    if (f < M) {
                        // the segment the ray hit last time; on a long track that hit can be kilometres away and drifts
                        // along the ring from build to build, so the LOCAL form also tries the neighbours and follows it
                        const int ntry = LOCAL ? 5 : 1;
                        for (int q = 0; q < ntry && !hit; ++q) {
                            int fq = f + ((q + 1) >> 1) * ((q & 1) ? 1 : -1);       // f, f+1, f-1, f+2, f-2
                            if (fq < 0) fq += M; else if (fq >= M) fq -= M;
                            double2 a, bb;
                            const int li = LOCAL ? loc_idx(fq, ring ? c.base1 : c.base0, M) : fq;
                            if (LOCAL && li >= (ring ? c.len1 : c.len0)) {     // far away: the segment record itself (x0,y0,x1,y1)
                                const double* g = c.gs0 + 4 * ((size_t)fq + (ring ? (size_t)c.M0 : 0));
                                a = make_double2(g[0], g[1]); bb = make_double2(g[2], g[3]);
                            } else { a = V[li]; bb = V[li + 1]; }
                            const double sd[4] = {a.x, a.y, bb.x - a.x, bb.y - a.y};
                            double tp = INF, tn = INF; int s_p = -1, s_n = -1;
                            seg_hit(sd, Pc, nx, ny, tp, tn, fq, s_p, s_n, ray_tests);
                            hit = (dir ? tn : tp) < INF;
                            if (LOCAL && hit && q > 0)
                                const_cast<unsigned long long*>(c.gcert)[i] = ((unsigned long long)(unsigned)fq << 32) | (unsigned long long)w0;
                        }
                    }
                    if (!hit) { flag = true; RL_WHY(5); }
                } else if (stt == kCertCone) {
                    double d0x, d0y;
                    cert_axis(w1, d0x, d0y);
                    const double cc = (double)(w0 >> 16) * (1.0 / 32767.0) - 1.0 + 2e-6;
                    const double ux = (Pc.x - cx0) - (double)__uint_as_float((unsigned)apex_w);
                    const double uy = (Pc.y - cy0) - (double)__uint_as_float((unsigned)(apex_w >> 32));
                    const double sgd = dir ? -1.0 : 1.0;
                    if (ux * d0x + uy * d0y >= sqrt(ux * ux + uy * uy) * cc && sgd * (nx * d0x + ny * d0y) >= cc) res = fmin(res, wd[ring]);
                    else { flag = true; RL_WHY(6); }
                } else { flag = true; RL_WHY(7); }
            }
        }
        dres[dir] = res;
    }
    UpdRes r;
    r.tests = (int)ray_tests; r.flagged = flag; r.hv = 0.0; r.lv = 0.0;
#ifdef RL_PHASE_TIMERS
    r.why = why;
#endif
    if (flag) return r;
    double hv = fmax(0.0, fmax(0.0, dres[0]) - c.guard);
    double lv = -fmax(0.0, fmax(0.0, dres[1]) - c.guard);
    if (!isfinite(hv)) hv = 0.0;
    if (!isfinite(lv)) lv = 0.0;
    r.hv = hv; r.lv = lv;
    return r;
}

template <int T, int K>
__device__ __forceinline__ unsigned corridor_update(const Part& pt, const double2* sP, double* sB, const int* sMisc,
                                                    const unsigned* sHint, const unsigned short* sClr, bool parity_ok, bool closed,
                                                    const double* __restrict__ gseg, const double* __restrict__ gcenter,
                                                    const unsigned long long* __restrict__ gcert, const unsigned long long* __restrict__ gapex,
                                                    long long segI0, long long segO0, long long segE, double guard,
                                                    double (&loc)[K], double (&hic)[K], long long& ray_tests)
{
    const int N = pt.N, tid = pt.tid;
    const double INF = dinf();
    const int M0 = (int)(segO0 - segI0), M1 = (int)(segE - segO0);
    double2* V0 = reinterpret_cast<double2*>(sB);
    double2* V1 = V0 + (M0 + 1);
    float2* F0 = reinterpret_cast<float2*>(V1 + (M1 + 1));   // FP32 copies relative to the job origin: they only ever SKIP exact work
    float2* F1 = F0 + (M0 + 1);
    const double2 org = sP[0];
    block_sync<T>();   // region B is free
    RL_DBG_ENTER(T, const_cast<int*>(sMisc), kDbgPhVerts);
    RL_DBG_ASSERT(const_cast<int*>(sMisc), (size_t)(M0 + M1 + 2) * 24 <= (size_t)T * K * 32, 2002);
    for (int q = tid; q <= M0; q += T) {
        const double2 v = (q < M0) ? *reinterpret_cast<const double2*>(gseg + 4 * (segI0 + q)) : *reinterpret_cast<const double2*>(gseg + 4 * (segI0 + M0 - 1) + 2);
        V0[q] = v; F0[q] = make_float2((float)(v.x - org.x), (float)(v.y - org.y));
    }
    for (int q = tid; q <= M1; q += T) {
        const double2 v = (q < M1) ? *reinterpret_cast<const double2*>(gseg + 4 * (segO0 + q)) : *reinterpret_cast<const double2*>(gseg + 4 * (segO0 + M1 - 1) + 2);
        V1[q] = v; F1[q] = make_float2((float)(v.x - org.x), (float)(v.y - org.y));
    }
    if (tid == 0) {
        UpdCtx& c = s_upd;
        const unsigned char* base = reinterpret_cast<const unsigned char*>(sP);    // sP is the start of the dynamic shared memory
        c.oV0 = (int)(reinterpret_cast<const unsigned char*>(V0) - base);
        c.oF0 = (int)(reinterpret_cast<const unsigned char*>(F0) - base);
        c.oHint = (int)(reinterpret_cast<const unsigned char*>(sHint) - base); c.oClr = (int)(reinterpret_cast<const unsigned char*>(sClr) - base);
        c.gcenter = gcenter; c.gcert = gcert; c.gapex = gapex;
        c.ox = org.x; c.oy = org.y; c.guard = guard; c.N = N; c.M0 = M0; c.M1 = M1; c.rf0 = sMisc[8]; c.rf1 = sMisc[9];
        c.mr0 = __int_as_float(sMisc[10]); c.mr1 = __int_as_float(sMisc[11]);
        c.parity_ok = parity_ok; c.closed = closed;
        c.base0 = 0; c.base1 = 0; c.len0 = M0; c.len1 = M1; c.oHalo = 0; c.gs0 = nullptr;
    }
    block_sync<T>();
    unsigned flagged = 0u;
    // the per-sample code exists once (a real function): the instruction cache matters more than the call, and the
    // results stay in registers (static j)
    // the sample's global-memory state (centre-line point, certificate words) is fetched one sample ahead
    double2 cn = make_double2(0.0, 0.0);
    unsigned long long wn = 0ull, an = 0ull;
    if (tid < N) { cn = *reinterpret_cast<const double2*>(gcenter + 2 * tid); wn = gcert[tid]; an = gapex[tid]; }
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const int i = tid + j * T;
        loc[j] = 0.0; hic[j] = 0.0;
        const double2 cc = cn;
        const unsigned long long wc = wn, ac = an;
        if (j + 1 < K && i + T < N) { cn = *reinterpret_cast<const double2*>(gcenter + 2 * (i + T)); wn = gcert[i + T]; an = gapex[i + T]; }
        if (i >= N) continue;
        const UpdRes r = corridor_update_sample<false>(i, cc.x, cc.y, wc, ac);
        ray_tests += r.tests;
        if (r.flagged) flagged |= (1u << j);
        else { hic[j] = r.hv; loc[j] = r.lv; }
    }
    RL_DBG_LEAVE(const_cast<int*>(sMisc), kDbgPhVerts);
    return flagged;
}

// Development build only (-DRL_PHASE_TIMERS): thread 0 accumulates clock64() cycles per phase of the job in the
// scratch area and leaves them in stats.J0[16..31] (outer iterations use entries 0..13).
#ifdef RL_PHASE_TIMERS
#define RL_PH(i) do { if (threadIdx.x == 0) { const long long t__ = clock64(); sPh[i] += t__ - sPh[15]; sPh[15] = t__; } } while (0)
#else
#define RL_PH(i) do { } while (0)
#endif

// ---- the solver kernel ------------------------------------------------------------------------------------
// CTAs one SM holds of a size class: shared memory decides (228 KB per SM, 1 KB reserve per CTA, 384 B static), and the
// launch bound promises no more than that, so the small classes may use the registers the missing CTAs leave free
__host__ __device__ constexpr int class_ctas_per_sm(int T, int K)
{
    const int per = T * K * 54 + kScratchBytes + 2 * kDbgGap * 2 + kDbgScr + 384 + 1024;
    const int n = 233472 / per;
    // ptxas grants registers in tiers: a bound of 13..16 CTAs of one warp still means 128, 12 means 168 -- and the ragged
    // one-warp kernel needs more than 128 to keep its projected-gradient loop out of local memory: 12 CTAs instead of 13
    if (T == 32 && n >= 12 && n < 16) return 12;
#ifndef RL_SMALL_CLASS_CTAS
#define RL_SMALL_CLASS_CTAS 8
#endif
    if (T == 64 && K == 4) return RL_SMALL_CLASS_CTAS;     // two warps, four samples per thread: 8 CTAs keep 128 registers (13 would mean 72)
    if (K == 4) { const int r = 512 / T; return r < n ? r : n; }   // the other K = 4 classes: as many CTAs as 128 registers allow
    return n > 16 ? 16 : (n < 1 ? 1 : n);
}
template <int T, int K, int MODE>
__global__ void __launch_bounds__(T, class_ctas_per_sm(T, K))
solve_kernel(const DevBatch B, const int* __restrict__ job_list, const int* __restrict__ item_off, int n_items)
{
    constexpr int NP = T * K;
    constexpr bool EXACT = (MODE == kModeExact), OPEN = (MODE == kModeOpen);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2* sP = reinterpret_cast<double2*>(smem_raw);
    double* sB = reinterpret_cast<double*>(smem_raw + (size_t)NP * 16 + kDbgGap);
    unsigned char* scr = smem_raw + (size_t)NP * 16 + (size_t)NP * 32 + 2 * kDbgGap;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(scr + kScrBar);
    double* sRed = reinterpret_cast<double*>(scr + kScrRed);
    double* sExF = reinterpret_cast<double*>(scr + kScrExF);
    double* sExL = reinterpret_cast<double*>(scr + kScrExL);
    int* sMisc = reinterpret_cast<int*>(scr + kScrMisc);
    unsigned* sHint = reinterpret_cast<unsigned*>(scr + kScrBytes + kDbgScr + kDbgGap);
    unsigned short* sClr = reinterpret_cast<unsigned short*>(scr + kScrBytes + kDbgScr + kDbgGap + (size_t)NP * 4);
#ifdef RL_DEBUG_CHECKS
    // guard zones: after the path area, after region B, after the scratch (+ sDone), after the per-sample corridor state
    unsigned* const guard[4] = {reinterpret_cast<unsigned*>(smem_raw + (size_t)NP * 16), reinterpret_cast<unsigned*>(smem_raw + (size_t)NP * 48 + kDbgGap),
                                reinterpret_cast<unsigned*>(scr + kScrBytes + kDbgScr), reinterpret_cast<unsigned*>(scr + kScrBytes + kDbgScr + kDbgGap + (size_t)NP * 6)};
    for (int z = 0; z < 4; ++z)
        for (int q = threadIdx.x; q < kDbgGap / 4; q += T) guard[z][q] = kDbgCanary;
    for (int q = threadIdx.x; q < 512; q += T) dbg_done(sMisc)[q] = 0;
    if (threadIdx.x == 0) {
        *reinterpret_cast<unsigned long long**>(sMisc + 24) = B.dbg;
        sMisc[26] = B.dbg ? (int)B.dbg[2] : 0;       // fault injection switch (rl_set_option "debug_inject")
        sMisc[27] = kDbgPhCoef;
    }
    __syncthreads();
#endif
#ifdef RL_PHASE_TIMERS
    long long* sPh = reinterpret_cast<long long*>(scr + kScrMisc + 128);
#endif

    // One CTA works through one ITEM = a short chain of jobs on the SAME track (typically its min-curvature and its
    // min-time stage, or neighbouring Configs of a sweep).  Everything the corridor code learns about the track --
    // anchor segments, clearances, parity bits (shared memory) and the existence certificates (handed from job to job
    // in global memory) -- stays valid for the next job of the chain, whose first corridor build is then an update.
    if ((int)blockIdx.x >= n_items) return;
    const int it0 = item_off[blockIdx.x], it1 = item_off[blockIdx.x + 1];
    uint32_t bar_phase = 0;
    if (threadIdx.x == 0) {
        mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    int prev_trk = -1;
    bool fast_update = false;   // corridor_update applies: both rings are vertex chains with more segments than a window
  for (int itj = it0; itj < it1; ++itj) {
    const int jid = job_list[itj];
    const rl_job job = B.jobs[jid];
    const rl_params& C = B.params[job.param];
    rl_job_stats* st = B.stats + jid;
    const int trk = job.track;
    const long long s0 = B.samp_off[trk];
    const int N = (int)(B.samp_off[trk + 1] - s0);
    const long long row0 = B.job_off[jid];
    const bool ev = (job.stage == RL_STAGE_EVAL);      // profile of the given path only: no corridor, no optimisation
    const bool mt = (job.stage == RL_STAGE_MINTIME) || ev;
    const double h = B.track_L[trk] / (double)N;
    const bool closed = (MODE != kModeOpen);

    // ---- blocked partition of the N samples over the threads ----
    Part pt;
    pt.N = N; pt.tid = threadIdx.x; pt.lane = pt.tid & 31; pt.warp = pt.tid >> 5;
    {
        const int tid = pt.tid;
        pt.Tact = EXACT ? T : min(T, max(1, N / 2));
        const int Kc = EXACT ? K : (N + pt.Tact - 1) / pt.Tact;
        const int nfull = EXACT ? T : N - pt.Tact * (Kc - 1);
        pt.cnt = EXACT ? K : (tid < pt.Tact ? (tid < nfull ? Kc : Kc - 1) : 0);
        pt.start = EXACT ? tid * K : (tid < nfull ? tid * Kc : nfull * Kc + (tid - nfull) * (Kc - 1));
        pt.tL = (tid == 0) ? pt.Tact - 1 : (tid < pt.Tact ? tid - 1 : 0);
        pt.tR = (tid >= pt.Tact - 1) ? 0 : tid + 1;
        pt.cntL = EXACT ? K : (pt.tL < nfull ? Kc : Kc - 1);
        if (T == 32) { pt.srcL = pt.tL; pt.srcR = pt.tR; }
        else { pt.srcL = (pt.lane + 31) & 31; pt.srcR = (pt.lane + 1) & 31; }
    }
    const int tid = pt.tid, cnt = pt.cnt, start = pt.start;
#ifdef RL_PHASE_TIMERS
    if (tid == 0) { for (int i = 0; i < 16; ++i) sPh[i] = 0; sPh[15] = clock64(); }
#endif

    if (tid == 0) {
        st->status = RL_OK; st->n = N; st->outer_done = 0; st->accepted = 0; st->backtracks = 0; st->evals = 0;
        st->vpass_rounds = 0; st->exist_scans = 0; st->ray_tests = 0; st->lap_time = 0.0;
        for (int o = 0; o < RL_MAX_OUTER_LOG; ++o) {
            st->J0[o] = 0.0; st->Jend[o] = 0.0; st->lap_outer[o] = 0.0; st->acc_outer[o] = 0; st->bt_outer[o] = 0;
        }
        *reinterpret_cast<HStep*>(s_hstep) = HStep(h);
        VPar& w = s_vpar;
        w.v_cap = C.v_cap_mps; w.a_lat_max = C.a_lat_max; w.kappa_eps = C.kappa_eps;
        {
            const double a_total = C.use_total_ge_lat ? fmax(C.a_total_max, C.a_lat_max) : C.a_total_max;   // main.cpp:802-804
            w.a_tot2 = a_total * a_total;
        }
        w.kd = 0.5 * C.rho_air * C.Cd * C.A_front_m2; w.Fr = C.mass_kg * 9.81 * C.c_rr; w.mass = C.mass_kg; w.inv_mass = 1.0 / C.mass_kg; w.P = C.P_max_W;
        w.acc_cap = C.a_long_acc_cap; w.brk_cap = C.a_long_brake_cap; w.h = h; w.has_power = (C.P_max_W > 0);
        fence_proxy_async();
    }
    block_sync<T>();

    // ---- load the centre line into sP (TMA bulk copy of N*16 bytes) ----
    if (tid == 0) {
        mbar_expect_tx(mbar, (uint32_t)N * 16u);
        bulk_g2s(sP, B.center_xy + 2 * s0, (uint32_t)N * 16u, mbar);
    }
    mbar_wait(mbar, bar_phase); bar_phase ^= 1;

    // alpha_total rows are accumulated by their owning thread (blocked layout)
#pragma unroll
    for (int k = 0; k < K; ++k)
        if (k < cnt) { B.alpha_total[row0 + start + k] = 0.0; B.alpha_last[row0 + start + k] = 0.0; }

    // the v(s) constants of the job are CTA-uniform and live in static shared memory (13 values that would otherwise
    // occupy 25 registers from here to the end of the job); thread 0 writes them, the barrier below publishes them
    const VPar& q = s_vpar;

    const long long segI0 = B.seg_off[2 * trk], segO0 = B.seg_off[2 * trk + 1], segE = B.seg_off[2 * trk + 2];
    const HStep& H = *reinterpret_cast<const HStep*>(s_hstep);          // h, 1/h, 1/(2h), 1/h^2 (DiffOps, main.cpp:547)
    long long ray_tests = 0;
    int vrounds = 0, ph = 0, ex_scans = 0;
    int acc_total = 0, bt_total = 0, ev_total = 0;
    const int max_outer = ev ? 0 : C.max_outer_iters;

    // initial corridor from the centre line: guard uses the veh_width ARGUMENT (main.cpp:706 / 930)
    constexpr int CAPF = fast_tile_cap(NP);
    const bool fast_rays = !ev && (segO0 - segI0 <= CAPF) && (segE - segO0 <= CAPF) && (CAPF < 8192);
    const bool parity_ok = (C.veh_width_arg * 0.5 + C.safety_margin_m >= 0.0) && (C.veh_width_m * 0.5 + C.safety_margin_m >= 0.0);
    // per-sample existence certificates (see corridor_build_fast) live in the heading and curvature rows until the final geometry pass
    unsigned long long* gcert = reinterpret_cast<unsigned long long*>(B.heading + row0);
    unsigned long long* gapex = reinterpret_cast<unsigned long long*>(B.curvature + row0);
    const bool same_track = fast_rays && (trk == prev_trk);   // the previous job of this chain left its corridor state behind
    if (!same_track) {
        for (int i = tid; i < NP; i += T) { sHint[i] = 0u; sClr[i] = 0; }
        for (int i = tid; i < N; i += T) gcert[i] = 0ull;
        fast_update = false;
    }
    if (fast_rays) {
        const double guard0 = C.veh_width_arg * 0.5 + C.safety_margin_m;
        double loc[K], hic[K];
#pragma unroll
        for (int j = 0; j < K; ++j) { loc[j] = 0.0; hic[j] = 0.0; }
        if (!same_track) {
            if (tid == 0) { sMisc[8] = 0; sMisc[9] = 0; sMisc[10] = 0; sMisc[11] = 0; }
            corridor_build_fast<T, K>(pt, sP, sB, mbar, bar_phase, sMisc, sHint, sClr, true, parity_ok, closed, B.seg, B.center_xy + 2 * s0, gcert, gapex,
                                      segI0, segO0, segE, guard0, 0xffffffffu, loc, hic, ray_tests, ex_scans);
        } else {
            unsigned flagged = 0xffffffffu;
            if (fast_update)
                flagged = corridor_update<T, K>(pt, sP, sB, sMisc, sHint, sClr, parity_ok, closed, B.seg, B.center_xy + 2 * s0, gcert, gapex,
                                                segI0, segO0, segE, guard0, loc, hic, ray_tests);
            if (block_or<T>(flagged != 0u))
                corridor_build_fast<T, K>(pt, sP, sB, mbar, bar_phase, sMisc, sHint, sClr, false, parity_ok, closed, B.seg, B.center_xy + 2 * s0, gcert, gapex,
                                          segI0, segO0, segE, guard0, flagged, loc, hic, ray_tests, ex_scans);
        }
        corridor_stage_out<T, K>(pt, sB, loc, hic, sMisc);
        if (!same_track)
            fast_update = (sMisc[8] & 2) && (sMisc[9] & 2) && (segO0 - segI0 > 2 * kWin + 1) && (segE - segO0 > 2 * kWin + 1);
    } else if (!ev) {
        double lo[K], hi[K];
#pragma unroll
        for (int k = 0; k < K; ++k) { lo[k] = 0.0; hi[k] = 0.0; }
        corridor_build_tiled<T, K>(pt, sP, sB, mbar, bar_phase, closed, B.seg, segI0, segO0, segE,
                                   C.veh_width_arg * 0.5 + C.safety_margin_m, lo, hi, ray_tests);
        corridor_stage_blocked<T, K>(pt, sB, lo, hi, sMisc);
    }

    double* sC0 = sB + tid;
    double* sCp = pair_base_a(sB + NP, NP, tid);
    double* sCm = pair_base_b(sB + NP, NP, tid);
    double* sSt = sB + 3 * NP + tid;
    RL_PH(0);   // setup + first corridor of the job

    for (int outer = 0; outer < max_outer; ++outer) {
        // =================== linearisation (main.cpp:722 / 941-944) ===================
        double A1[K], A2[K], N0[K], Wd[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            A1[k] = 0.0; A2[k] = 0.0; N0[k] = 0.0; Wd[k] = 1.0;
            if (k < cnt) {
                const int i = start + k;
                double nx, ny, xp, yp, xpp, ypp;
                normal_at(sP, i, N, closed, nx, ny);
                derivs_at(sP, i, N, H, closed, xp, yp, xpp, ypp);
                A1[k] = nx * ypp - ny * xpp;          // main.cpp:644-646
                A2[k] = xp * ny - yp * nx;
                N0[k] = xp * ypp - yp * xpp;
                Wd[k] = 1.0 / pow15(xp * xp + yp * yp);   // W = 1/denom (main.cpp:647-648)
            }
        }
        // ---- park the path in global memory (the job's xy rows) while the PGD runs: its shared-memory
        //      region holds the box bounds lo|hi (slot-major) until the path update needs P again ----
        fence_proxy_async();
        block_sync<T>();
        if (tid == 0) bulk_s2g_issue(B.xy + 2 * row0, sP, (uint32_t)N * 16u);   // returns once the smem read is done
        block_sync<T>();
        double* sLo = pair_base_a(reinterpret_cast<double*>(sP), NP, tid);
        double* sHi = pair_base_b(reinterpret_cast<double*>(sP), NP, tid);
        staged_bounds_home<T, K>(pt, sB, sLo, sHi, sMisc);   // the corridor's bounds: staging area -> their home for the PGD
        double gam[K];
#pragma unroll
        for (int k = 0; k < K; ++k) gam[k] = 1.0;
        double lap_outer = 0.0;
        RL_PH(1);   // linearisation + parking the path
        if (mt) {
            // ============ v(s) profile + time weights (main.cpp:944-977) ============
            double kap[K], vv[K], axd[K], vkap[K];
#pragma unroll
            for (int k = 0; k < K; ++k) kap[k] = (k < cnt) ? N0[k] * Wd[k] : 0.0;    // kappa = N0 / denom, main.cpp:618
            RL_DBG_ENTER(T, sMisc, kDbgPhVsweep);
            // the linearisation (48 registers) waits behind the exchange arrays of the sweeps (first NP doubles of region B)
            // while the v(s) profile runs: own slots only, written and read back by the same thread
            // (the exchange arrays take 6 T doubles: with K = 4 that is more than one plane of T K, and only A1 and A2 fit behind them)
            constexpr int kPark0 = (6 * T > NP) ? 6 * T : NP;
            constexpr bool kParkW = (kPark0 + 3 * NP <= 4 * NP);
            static_assert(kPark0 + 2 * NP <= 4 * NP, "region B holds the exchange arrays and two parked planes");
#pragma unroll
            for (int k = 0; k < K; ++k) {
                sB[kPark0 + k * T + tid] = A1[k]; sB[kPark0 + NP + k * T + tid] = A2[k];
                if (kParkW) sB[kPark0 + 2 * NP + k * T + tid] = Wd[k];
            }
            vprofile_blocked<T, K>(pt, q, kap, vv, C.max_vpass_iters, sB, vrounds, closed, vkap);
            block_sync<T>();
            lap_outer = lap_and_ax<T, K>(pt, q, vv, axd, sB, sRed + ph * 32, closed);
            RL_DBG_LEAVE(sMisc, kDbgPhVsweep);
            ph ^= 1;
            double v_avg = 0.0;
            if (C.time_weight_use_inv_v) {            // main.cpp:951
                double sv = 0.0, dz = 0.0;
#pragma unroll
                for (int k = 0; k < K; ++k) if (k < cnt) sv += vv[k];
                block_sum2<T>(sv, dz, sRed + ph * 32, pt.lane, pt.warp);
                ph ^= 1;
                v_avg = sv / (double)max(1, N);
            }
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (k < cnt) {                         // main.cpp:954-975
                    const double vk = vkap[k];     // sqrt(a_lat_max / max(|kappa|, kappa_eps)), main.cpp:956
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
#pragma unroll
            for (int k = 0; k < K; ++k) {
                A1[k] = sB[kPark0 + k * T + tid]; A2[k] = sB[kPark0 + NP + k * T + tid];
                if (kParkW) Wd[k] = sB[kPark0 + 2 * NP + k * T + tid];
            }
            block_sync<T>();
        }
        RL_PH(2);   // v(s) profile + time weights
        // ---- stencil coefficients into region B (slot-major) ----
        RL_DBG_ENTER(T, sMisc, kDbgPhCoef);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double c0 = 0.0, cp = 0.0, cm = 0.0;
            if (k < cnt) {
                const double gw = gam[k] * Wd[k];
                c0 = gw * N0[k];
                const double c1 = gw * A1[k] * H.inv2h, c2 = gw * A2[k] * H.invh2;
                cp = c1 + c2; cm = c2 - c1;
                if (OPEN) {   // DiffOpsOpen, main.cpp:563-575: D1 one-sided with 1/h at the ends, D2 zero there
                    const int i = start + k;
                    if (N == 1) { cp = 0.0; cm = 0.0; }
                    else if (i == 0) { cp = 2.0 * c1; cm = 0.0; }
                    else if (i == N - 1) { cp = 0.0; cm = -2.0 * c1; }
                }
            }
            sC0[k * T] = c0; st_pair<T>(sCp, sCm, k, cp, cm);
        }
        block_sync<T>();
        double cL[3], cR[3];
        {
            const int kl = pt.cntL - 1;
            cL[0] = sB[pt.tL + kl * T]; cR[0] = sB[pt.tR];
            ld_pair<T>(pair_base_a(sB + NP, NP, pt.tL), pair_base_b(sB + NP, NP, pt.tL), kl, cL[1], cL[2]);
            ld_pair<T>(pair_base_a(sB + NP, NP, pt.tR), pair_base_b(sB + NP, NP, pt.tR), 0, cR[1], cR[2]);
        }
        if (OPEN) {   // nothing beyond the two ends of an open track
            if (tid == 0) { cL[0] = 0.0; cL[1] = 0.0; cL[2] = 0.0; }
            if (tid == pt.Tact - 1) { cR[0] = 0.0; cR[1] = 0.0; cR[2] = 0.0; }
        }
        if (!EXACT) {
            // the first unused slot mirrors the right neighbour's first sample (position cnt+1 of the window)
            if (cnt < K && cnt > 0) { sC0[cnt * T] = cR[0]; st_pair<T>(sCp, sCm, cnt, cR[1], cR[2]); }
        }
        RL_PH(3);   // stencil coefficients
        // =================== projected gradient with Armijo (main.cpp:723-742 / 996-1026) ===================
        const double lamJ = C.lambda_smooth * H.inv2h * H.inv2h;
        const PgdOut po = pgd_outer<T, K, MODE>(pt, sLo, sHi, cL, cR, sC0, sCp, sCm, sSt, sRed, sExF, sExL, ph, lamJ,
                                                 C.step_init, C.step_min, C.armijo_c, C.max_inner_iters);
        acc_total += po.acc; bt_total += po.bt; ev_total += po.ev;
        if (tid == 0 && outer < RL_MAX_OUTER_LOG) {
            st->J0[outer] = po.J0; st->Jend[outer] = po.Jend; st->lap_outer[outer] = lap_outer;
            st->acc_outer[outer] = po.acc; st->bt_outer[outer] = po.bt;
        }
        RL_PH(4);   // projected-gradient loop
        // ---- bring the path back ----
        block_sync<T>();
        if (tid == 0) {
            bulk_wait_all();                       // the parked copy is complete and visible
            fence_proxy_async();
            mbar_expect_tx(mbar, (uint32_t)N * 16u);
            bulk_g2s(sP, B.xy + 2 * row0, (uint32_t)N * 16u, mbar);
        }
        mbar_wait(mbar, bar_phase); bar_phase ^= 1;
        // =================== path update (main.cpp:743-746 / 1027-1031) ===================
        double2 Pn[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (k < cnt) {
                const int i = start + k;
                const double al = sSt[k * T];
                double nx, ny;
                normal_at(sP, i, N, closed, nx, ny);
                const double2 Pc = sP[i];
                Pn[k].x = Pc.x + nx * al; Pn[k].y = Pc.y + ny * al;
                B.alpha_total[row0 + i] += al;
                if (outer == max_outer - 1) B.alpha_last[row0 + i] = al;
            }
        }
        RL_DBG_LEAVE(sMisc, kDbgPhCoef);      // the stash (region B) has been read: the next corridor pass may take the region
        block_sync<T>();
#pragma unroll
        for (int k = 0; k < K; ++k) if (k < cnt) sP[start + k] = Pn[k];
        block_sync<T>();
        // =================== corridor from the new path (main.cpp:749-756 / 1033-1040) ===================
        // (the reference also rebuilds it after the LAST path update, but nothing reads that corridor: skipped)
        RL_PH(5);   // path back + path update
        if (outer + 1 == max_outer) continue;
        if (fast_rays) {
            const double guard = C.veh_width_m * 0.5 + C.safety_margin_m;
            double loc[K], hic[K];
            unsigned flagged = 0xffffffffu;
            if (fast_update)
                flagged = corridor_update<T, K>(pt, sP, sB, sMisc, sHint, sClr, parity_ok, closed, B.seg, B.center_xy + 2 * s0, gcert, gapex,
                                                segI0, segO0, segE, guard, loc, hic, ray_tests);
            else {
#pragma unroll
                for (int j = 0; j < K; ++j) { loc[j] = 0.0; hic[j] = 0.0; }
            }
            RL_PH(6);   // corridor update pass
            if (block_or<T>(flagged != 0u))   // some certificate failed (or none exists yet): the searching path rebuilds those samples
                corridor_build_fast<T, K>(pt, sP, sB, mbar, bar_phase, sMisc, sHint, sClr, false, parity_ok, closed, B.seg, B.center_xy + 2 * s0, gcert, gapex,
                                          segI0, segO0, segE, guard, flagged, loc, hic, ray_tests, ex_scans);
            corridor_stage_out<T, K>(pt, sB, loc, hic, sMisc);
            RL_PH(7);   // searching path for flagged samples + staging
        } else {
            double lo[K], hi[K];
#pragma unroll
            for (int k = 0; k < K; ++k) { lo[k] = 0.0; hi[k] = 0.0; }
            corridor_build_tiled<T, K>(pt, sP, sB, mbar, bar_phase, closed, B.seg, segI0, segO0, segE,
                                       C.veh_width_m * 0.5 + C.safety_margin_m, lo, hi, ray_tests);
            corridor_stage_blocked<T, K>(pt, sB, lo, hi, sMisc);
            RL_PH(7);
        }
    }

    // the next job of the chain inherits the certificates (same thread, same samples: no barrier needed)
    if (fast_rays && itj + 1 < it1) {
        const int njid = job_list[itj + 1];
        if (B.jobs[njid].track == trk) {
            const long long nrow0 = B.job_off[njid];
            unsigned long long* ncert = reinterpret_cast<unsigned long long*>(B.heading + nrow0);
            unsigned long long* napex = reinterpret_cast<unsigned long long*>(B.curvature + nrow0);
            for (int i = tid; i < N; i += T) { ncert[i] = gcert[i]; napex[i] = gapex[i]; }
        }
    }
    prev_trk = ev ? -1 : trk;
    // =================== final geometry (main.cpp:761 / 1046) ===================
    block_sync<T>();
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const int i = tid + j * T;
        if (i < N) {
            double xp, yp, xpp, ypp;
            derivs_at(sP, i, N, H, closed, xp, yp, xpp, ypp);
            B.heading[row0 + i] = atan2(yp, xp);
            B.curvature[row0 + i] = (xp * ypp - yp * xpp) / pow15(xp * xp + yp * yp);
        }
    }
    double lap = 0.0;
    if (mt) {
        // final v(s) profile (main.cpp:1047)
        double kap[K], vv[K], axd[K], vkap[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            kap[k] = 0.0;
            if (k < cnt) {
                const int i = start + k;
                double xp, yp, xpp, ypp;
                derivs_at(sP, i, N, H, closed, xp, yp, xpp, ypp);
                kap[k] = (xp * ypp - yp * xpp) / pow15(xp * xp + yp * yp);
            }
        }
        block_sync<T>();
        RL_DBG_ENTER(T, sMisc, kDbgPhVsweep);
        vprofile_blocked<T, K>(pt, q, kap, vv, C.max_vpass_iters, sB, vrounds, closed, vkap);
        block_sync<T>();
        lap = lap_and_ax<T, K>(pt, q, vv, axd, sB, sRed + ph * 32, closed);
        RL_DBG_LEAVE(sMisc, kDbgPhVsweep);
        ph ^= 1;
#pragma unroll
        for (int k = 0; k < K; ++k)
            if (k < cnt) { B.v[row0 + start + k] = vv[k]; B.ax[row0 + start + k] = axd[k]; }
    }
    // raceline out: TMA bulk store of sP
    fence_proxy_async();
    block_sync<T>();
    if (tid == 0) bulk_s2g(B.xy + 2 * row0, sP, (uint32_t)N * 16u);

    // counters
    {
        double rt = (double)ray_tests, es = (double)ex_scans;
        block_sum2<T>(rt, es, sRed + ph * 32, pt.lane, pt.warp);
        ph ^= 1;
        if (tid == 0) {
            st->outer_done = max_outer; st->accepted = acc_total; st->backtracks = bt_total; st->evals = ev_total;
            st->vpass_rounds = vrounds; st->ray_tests = (long long)rt; st->lap_time = lap; st->exist_scans = (int)es;
        }
    }
#ifdef RL_DEBUG_CHECKS
    if (tid == 0)
        for (int z = 0; z < 4; ++z)
            for (int q = 0; q < kDbgGap / 4; ++q)
                if (guard[z][q] != kDbgCanary) { dbg_fail(sMisc, 9000 + z); break; }
#endif
    RL_PH(8);   // certificates hand-over, final geometry, final v(s) profile, stores
#ifdef RL_PHASE_TIMERS
    if (tid == 0) for (int i = 0; i < 9; ++i) st->J0[16 + i] = (double)sPh[i];
#endif
    block_sync<T>();   // the next job of the chain reuses the shared-memory regions
  }
}


}  // namespace
}  // namespace rl

// one translation unit per thread count T instantiates its three kernels (closed ragged, closed exact-fit, open)
#define RL_INSTANTIATE_AS(T, K, NAME)                                                                                          \
    namespace rl {                                                                                                    \
    int launch_solve_##NAME(const DevBatch& B, const int* job_list, const int* item_off, int n_items, int mode, void* stream) \
    {                                                                                                                 \
        const size_t smem = smem_bytes_for_class(T, K);                                                               \
        cudaStream_t s = (cudaStream_t)stream;                                                                        \
        if (mode == 1) solve_kernel<T, K, 1><<<n_items, T, smem, s>>>(B, job_list, item_off, n_items);                \
        else if (mode == 2) solve_kernel<T, K, 2><<<n_items, T, smem, s>>>(B, job_list, item_off, n_items);           \
        else solve_kernel<T, K, 0><<<n_items, T, smem, s>>>(B, job_list, item_off, n_items);                          \
        return (int)cudaGetLastError();                                                                               \
    }                                                                                                                 \
    int configure_solve_##NAME()                                                                                         \
    {                                                                                                                 \
        const int smem = (int)smem_bytes_for_class(T, K);                                                             \
        cudaError_t e = cudaFuncSetAttribute(solve_kernel<T, K, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
        if (e == cudaSuccess)                                                                                         \
            e = cudaFuncSetAttribute(solve_kernel<T, K, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);       \
        if (e == cudaSuccess)                                                                                         \
            e = cudaFuncSetAttribute(solve_kernel<T, K, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);       \
        return (int)e;                                                                                                \
    }                                                                                                                 \
    int occupancy_solve_##NAME()                                                                                         \
    {                                                                                                                 \
        int n = 0;                                                                                                    \
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, solve_kernel<T, K, 0>, T, smem_bytes_for_class(T, K)) != cudaSuccess) { cudaGetLastError(); n = 0; } \
        return n;                                                                                                     \
    }                                                                                                                 \
    }

#define RL_INSTANTIATE(T, K) RL_INSTANTIATE_AS(T, K, T)
