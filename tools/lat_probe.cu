// Development probe: dependent-issue latencies on sm_100a of the instructions the PGD loop is made of.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/lat_probe tools/lat_probe.cu && ./gpurun_out/lat_probe
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(double* out, long long* cyc, int iters)
{
    double a = threadIdx.x * 1e-9 + 1.0, b = 0.999999, c = 1e-7;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); }
    long long t1 = clock64();
    double d = a;
    for (int i = 0; i < iters; ++i) { d = d + c; d = d + b; d = d + c; d = d + b; }
    long long t2 = clock64();
    double e = d;
    for (int i = 0; i < iters; ++i) { e = __shfl_xor_sync(0xffffffffu, e, 1); e = __shfl_xor_sync(0xffffffffu, e, 2); e = __shfl_xor_sync(0xffffffffu, e, 4); e = __shfl_xor_sync(0xffffffffu, e, 8); }
    long long t3 = clock64();
    double f = e;
    for (int i = 0; i < iters; ++i) { f = __shfl_xor_sync(0xffffffffu, f, 1) + f; f = __shfl_xor_sync(0xffffffffu, f, 2) + f; f = __shfl_xor_sync(0xffffffffu, f, 4) + f; f = __shfl_xor_sync(0xffffffffu, f, 8) + f; }
    long long t4 = clock64();
    double g = f; 
    for (int i = 0; i < iters; ++i) { g = (g < b) ? g + c : b; g = (g < c) ? c : g * b; g = (g < b) ? g + c : b; g = (g < c) ? c : g * b; }
    long long t5 = clock64();
    // 8 independent DFMA chains from one warp: issue-limited rate
    double x0 = g, x1 = g + 1, x2 = g + 2, x3 = g + 3, x4 = g + 4, x5 = g + 5, x6 = g + 6, x7 = g + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, b, c); x1 = fma(x1, b, c); x2 = fma(x2, b, c); x3 = fma(x3, b, c);
        x4 = fma(x4, b, c); x5 = fma(x5, b, c); x6 = fma(x6, b, c); x7 = fma(x7, b, c);
    }
    long long t6 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; cyc[5] = t6 - t5;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}
int main()
{
    double* d; long long* c; long long h[6];
    cudaMalloc(&d, 8 * 1024 * 256); cudaMalloc(&c, 48);
    const int iters = 4096;
    for (int warps = 1; warps <= 16; warps *= 2) {
        probe<<<1, 32 * warps>>>(d, c, iters);
        cudaMemcpy(h, c, 48, cudaMemcpyDeviceToHost);
        printf("warps/CTA %2d (1 SM): dep DFMA %.2f cyc, dep DADD %.2f, dep SHFL64 %.2f, SHFL64+DADD %.2f, DSETP+select+op %.2f, 8 indep DFMA: %.2f cyc each\n", warps,
               h[0] / (4.0 * iters), h[1] / (4.0 * iters), h[2] / (4.0 * iters), h[3] / (4.0 * iters), h[4] / (4.0 * iters), h[5] / (8.0 * iters));
    }
    return 0;
}
