// instantiation of solve_kernel<32, 8, *> (compiled in parallel with the other thread counts)
#include "raceline_kernels.cuh"

RL_INSTANTIATE(32, 8)
