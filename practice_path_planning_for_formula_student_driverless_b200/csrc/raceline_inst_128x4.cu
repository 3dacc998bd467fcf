// instantiation of solve_kernel<128, 4, *> (256 < N <= 512 with four samples per thread)
#include "raceline_kernels.cuh"

RL_INSTANTIATE_AS(128, 4, 128x4)
