// instantiation of solve_kernel<256, 4, *> (512 < N <= 1024 with four samples per thread)
#include "raceline_kernels.cuh"

RL_INSTANTIATE_AS(256, 4, 256x4)
