// instantiation of solve_kernel<128, 8, *> (compiled in parallel with the other thread counts)
#include "raceline_kernels.cuh"

RL_INSTANTIATE(128, 8)
