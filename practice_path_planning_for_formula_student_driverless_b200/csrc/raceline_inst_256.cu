// instantiation of solve_kernel<256, 8, *> (compiled in parallel with the other thread counts)
#include "raceline_kernels.cuh"

RL_INSTANTIATE(256, 8)
