// raceline_dispatch.cu -- size-class dispatch over the per-T translation units + the FP64 throughput probe.
#include <cuda_runtime.h>

#include "raceline_device.h"

namespace rl {

#define RL_DECL(T)                                                                                   \
    int launch_solve_##T(const DevBatch& B, const int* job_list, const int* item_off, int n_items, int mode, void* stream); \
    int configure_solve_##T();                                                                       \
    int occupancy_solve_##T();
RL_DECL(32) RL_DECL(64) RL_DECL(128) RL_DECL(256) RL_DECL(512) RL_DECL(64x4) RL_DECL(128x4) RL_DECL(256x4)
#undef RL_DECL

namespace {
// FP64 FMA throughput probe: 8 independent chains per thread, 2 flops per FMA.
__global__ void fp64_peak_kernel(double* out, int iters)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace

namespace { int g_occ[kNumClasses] = {0, 0, 0, 0, 0, 0, 0, 0}; }

int configure_kernels()
{
    int e = configure_solve_32();
    if (!e) e = configure_solve_64();
    if (!e) e = configure_solve_128();
    if (!e) e = configure_solve_256();
    if (!e) e = configure_solve_512();
    if (!e) e = configure_solve_64x4();
    if (!e) e = configure_solve_128x4();
    if (!e) e = configure_solve_256x4();
    if (!e) e = configure_solve_cluster();
    if (!e) e = configure_geom();
    if (!e) {   // what the device really holds per SM (registers, shared memory + its per-CTA reserve): the host plan's CTA slots
        g_occ[0] = occupancy_solve_32(); g_occ[1] = occupancy_solve_64(); g_occ[2] = occupancy_solve_128();
        g_occ[3] = occupancy_solve_256(); g_occ[4] = occupancy_solve_512(); g_occ[5] = occupancy_solve_64x4(); g_occ[6] = occupancy_solve_128x4(); g_occ[7] = occupancy_solve_256x4();
    }
    return e;
}

int ctas_per_sm(int cls)
{
    if (cls >= 0 && cls < kNumClasses && g_occ[cls] > 0) return g_occ[cls];
    const int np = kClasses[cls].T * kClasses[cls].K;
    return np >= 4096 ? 1 : (4096 / np > 16 ? 16 : 4096 / np);
}

int launch_solve(const DevBatch& B, const int* job_list, int n_list, const int* item_off, int n_items, int cls, int mode, void* stream)
{
    if (n_list <= 0) return 0;
    if (cls >= kClusterClassBase) return launch_solve_cluster(B, job_list, item_off, n_items, cls - kClusterClassBase, mode, stream);
    if (cls == 5) return launch_solve_64x4(B, job_list, item_off, n_items, mode, stream);
    if (cls == 6) return launch_solve_128x4(B, job_list, item_off, n_items, mode, stream);
    if (cls == 7) return launch_solve_256x4(B, job_list, item_off, n_items, mode, stream);
    switch (kClasses[cls].T) {
        case 32: return launch_solve_32(B, job_list, item_off, n_items, mode, stream);
        case 64: return launch_solve_64(B, job_list, item_off, n_items, mode, stream);
        case 128: return launch_solve_128(B, job_list, item_off, n_items, mode, stream);
        case 256: return launch_solve_256(B, job_list, item_off, n_items, mode, stream);
        case 512: return launch_solve_512(B, job_list, item_off, n_items, mode, stream);
        default: return (int)cudaErrorInvalidValue;
    }
}

int launch_fp64_peak(double* d_out, int blocks, int threads, int iters, void* stream)
{
    fp64_peak_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(d_out, iters);
    return (int)cudaGetLastError();
}

}  // namespace rl
