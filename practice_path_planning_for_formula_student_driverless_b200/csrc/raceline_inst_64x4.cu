// instantiation of solve_kernel<64, 4, *>: the two-warp class of the small maps (compiled in parallel with the other classes)
#include "raceline_kernels.cuh"

RL_INSTANTIATE_AS(64, 4, 64x4)
