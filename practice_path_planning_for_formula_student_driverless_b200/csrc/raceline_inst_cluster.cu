// instantiation of solve_cluster_kernel<8, *> (long tracks: one thread-block cluster per job)
#include "raceline_cluster.cuh"
