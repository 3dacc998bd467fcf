// Development probe: the projected-gradient loop of raceline_kernels.cuh in isolation, on realistic stencil
// coefficients, to measure cycles per cost/gradient evaluation for loop variants / occupancies without the rest of the
// solver.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -DPROBE_V=0 -o pgd_probe tools/pgd_probe.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../practice_path_planning_for_formula_student_driverless_b200/csrc/raceline_kernels.cuh"

#ifndef PROBE_T
#define PROBE_T 256
#endif
#ifndef PROBE_K
#define PROBE_K 8
#endif
#ifndef PROBE_MINB
#define PROBE_MINB 2
#endif
#ifndef PROBE_V
#define PROBE_V 0
#endif

namespace rl {
namespace {

template <int T, int K, int MODE>
__global__ void __launch_bounds__(T, PROBE_MINB)
probe_kernel(const double* __restrict__ gc0, const double* __restrict__ gcp, const double* __restrict__ gcm, const double* __restrict__ glo,
             const double* __restrict__ ghi, int N, int outers, int max_inner, double lamJ, double step_init, double step_min,
             double armijo_c, long long* out_cyc, int* out_ev, double* out_J, double* out_alpha)
{
    constexpr int NP = T * K;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sPd = reinterpret_cast<double*>(smem_raw);                 // 2*NP doubles: lo|hi
    double* sB = reinterpret_cast<double*>(smem_raw + (size_t)NP * 16); // 4*NP doubles
    unsigned char* scr = smem_raw + (size_t)NP * 16 + (size_t)NP * 32;
    double* sRed = reinterpret_cast<double*>(scr + kScrRed);
    double* sExF = reinterpret_cast<double*>(scr + kScrExF);
    double* sExL = reinterpret_cast<double*>(scr + kScrExL);
    Part pt;
    pt.N = N; pt.tid = threadIdx.x; pt.lane = pt.tid & 31; pt.warp = pt.tid >> 5;
    const int tid = pt.tid;
    pt.Tact = T; pt.cnt = K; pt.start = tid * K;
    pt.tL = (tid == 0) ? T - 1 : tid - 1; pt.tR = (tid == T - 1) ? 0 : tid + 1; pt.cntL = K;
    pt.srcL = (pt.lane + 31) & 31; pt.srcR = (pt.lane + 1) & 31;
    double* sC0 = sB + tid;
    double* sCp = pair_base_a(sB + NP, NP, tid);
    double* sCm = pair_base_b(sB + NP, NP, tid);
    double* sSt = sB + 3 * NP + tid;
    double* sLo = pair_base_a(sPd, NP, tid);
    double* sHi = pair_base_b(sPd, NP, tid);
    for (int k = 0; k < K; ++k) {
        const int i = tid * K + k;
        sC0[k * T] = gc0[i];
        st_pair<T>(sCp, sCm, k, gcp[i], gcm[i]);
        st_pair<T>(sLo, sHi, k, glo[i], ghi[i]);
    }
    __syncthreads();
    double cL[3], cR[3];
    cL[0] = sB[pt.tL + (K - 1) * T]; cR[0] = sB[pt.tR];
    ld_pair<T>(pair_base_a(sB + NP, NP, pt.tL), pair_base_b(sB + NP, NP, pt.tL), K - 1, cL[1], cL[2]);
    ld_pair<T>(pair_base_a(sB + NP, NP, pt.tR), pair_base_b(sB + NP, NP, pt.tR), 0, cR[1], cR[2]);
    int ph = 0, ev = 0;

    double Jend = 0.0;
    const long long t0 = clock64();
    for (int o = 0; o < outers; ++o) {
        const PgdOut po = pgd_outer<T, K, MODE>(pt, sLo, sHi, cL, cR, sC0, sCp, sCm, sSt, sRed, sExF, sExL, ph, lamJ, step_init, step_min,
                                                 armijo_c, max_inner);
        ev += po.ev; Jend = po.Jend;
        __syncthreads();
    }
    const long long t1 = clock64();
    if (tid == 0) { out_cyc[blockIdx.x] = t1 - t0; out_ev[blockIdx.x] = ev; out_J[blockIdx.x] = Jend; }
    if (blockIdx.x == 0) for (int k = 0; k < K; ++k) out_alpha[tid * K + k] = sSt[k * T];
}

}  // namespace
}  // namespace rl

int main(int argc, char** argv)
{
    constexpr int T = PROBE_T, K = PROBE_K, N = T * K;
    const int outers = argc > 1 ? atoi(argv[1]) : 4;
    const int waves = argc > 2 ? atoi(argv[2]) : 2;
    const int smem_kb = argc > 3 ? atoi(argv[3]) : 110;     // 110 KB = what the solver kernel uses (2 CTAs/SM)
    const int pin_every = argc > 4 ? atoi(argv[4]) : 0;     // > 0: every pin_every-th sample gets a (nearly) closed box (active constraints)
    // a flower-shaped closed track, h = 1.8 m: linearisation exactly as solve_kernel builds it
    const double pi = 3.14159265358979323846, R0 = N * 1.8 / (2 * pi);
    std::vector<double> px(N), py(N);
    for (int i = 0; i < N; ++i) {
        const double th = 2 * pi * i / N, r = R0 * (1 + 0.02 * sin(7 * th) + 0.004 * sin(23 * th + 1.0));
        px[i] = r * cos(th); py[i] = r * sin(th);
    }
    double L = 0;
    for (int i = 0; i < N; ++i) L += hypot(px[(i + 1) % N] - px[i], py[(i + 1) % N] - py[i]);
    const double h = L / N, inv2h = 1 / (2 * h), invh2 = 1 / (h * h), lambda = 1.6e-3, lamJ = lambda * inv2h * inv2h;
    std::vector<double> c0(N), cp(N), cm(N), lo(N), hi(N);
    for (int i = 0; i < N; ++i) {
        const int a = (i + N - 1) % N, b = (i + 1) % N;
        double tx = (px[b] - px[a]) * 0.5, ty = (py[b] - py[a]) * 0.5, len = hypot(tx, ty), nx = -ty / len, ny = tx / len;
        const double xp = (px[b] - px[a]) / (2 * h), yp = (py[b] - py[a]) / (2 * h);
        const double xpp = (px[b] - 2 * px[i] + px[a]) / (h * h), ypp = (py[b] - 2 * py[i] + py[a]) / (h * h);
        const double A1 = nx * ypp - ny * xpp, A2 = xp * ny - yp * nx, N0 = xp * ypp - yp * xpp, W = pow(fmax(1e-12, xp * xp + yp * yp), 1.5);
        const double gw = 1.0 / W, c1 = gw * A1 * inv2h, c2 = gw * A2 * invh2;
        c0[i] = gw * N0; cp[i] = c1 + c2; cm[i] = c2 - c1; lo[i] = -1.2; hi[i] = 1.2;
        if (pin_every > 0 && i % pin_every == 0) { lo[i] = -1e-9; hi[i] = 1e-9; }
    }
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    auto kern = rl::probe_kernel<T, K, 1>;
    const size_t smem = (size_t)smem_kb * 1024;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { printf("smem attr failed\n"); return 1; }
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, T, smem);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
    const int grid = sms * occ * waves;
    double *d[5], *dJ, *dA; long long* dcyc; int* dev_ev;
    std::vector<double>* src[5] = {&c0, &cp, &cm, &lo, &hi};
    for (int i = 0; i < 5; ++i) { cudaMalloc(&d[i], N * 8); cudaMemcpy(d[i], src[i]->data(), N * 8, cudaMemcpyHostToDevice); }
    cudaMalloc(&dJ, grid * 8); cudaMalloc(&dA, N * 8); cudaMalloc(&dcyc, grid * 8); cudaMalloc(&dev_ev, grid * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        kern<<<grid, T, smem>>>(d[0], d[1], d[2], d[3], d[4], N, outers, 120, lamJ, 0.65, 1e-6, 1e-5, dcyc, dev_ev, dJ, dA);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = fminf(best, ms);
    }
    std::vector<long long> cyc(grid); std::vector<int> ev(grid); std::vector<double> J(grid), A(N);
    cudaMemcpy(cyc.data(), dcyc, grid * 8, cudaMemcpyDeviceToHost); cudaMemcpy(ev.data(), dev_ev, grid * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(J.data(), dJ, grid * 8, cudaMemcpyDeviceToHost); cudaMemcpy(A.data(), dA, N * 8, cudaMemcpyDeviceToHost);
    double mc = 0; for (int i = 0; i < grid; ++i) mc += (double)cyc[i] / ev[i];
    mc /= grid;
    double asum = 0, amax = 0; for (int i = 0; i < N; ++i) { asum += A[i]; amax = fmax(amax, fabs(A[i])); }
    const double evals_total = (double)ev[0] * grid;
    printf("V=%d T=%d K=%d minb=%d regs=%d spill(local)=%zuB smem=%dKB occ=%d grid=%d: evals/CTA=%d  %.1f cyc/eval/CTA  "
           "%.3f ms  %.2f Mevals/s/SM  => %.0f cyc per eval-slot/SM   J=%.17g sum(alpha)=%.17g max|alpha|=%.6f\n",
           PROBE_V, T, K, PROBE_MINB, fa.numRegs, fa.localSizeBytes, smem_kb, occ, grid, ev[0], mc, best,
           evals_total / (best * 1e-3) / sms / 1e6, 1.965e9 / (evals_total / (best * 1e-3) / sms), J[0], asum, amax);
    return 0;
}
