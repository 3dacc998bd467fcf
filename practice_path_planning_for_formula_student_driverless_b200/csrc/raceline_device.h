// raceline_device.h -- internal types shared by the kernels and the host-side ABI implementation.
#pragma once

#include <cstdint>

#include "../../include/raceline_b200.h"

namespace rl {

// Device view of one packed batch (include/raceline_b200.h: rl_batch_desc / rl_batch_out).
struct DevBatch {
    const long long* samp_off;   // [n_tracks+1]
    const long long* seg_off;    // [2*n_tracks+1]
    const double* center_xy;     // [2*total_samples]
    const double* seg;           // [4*total_segs]
    const double* track_L;       // [n_tracks]
    const int* track_closed;     // [n_tracks]
    const rl_params* params;     // [n_params]
    const rl_job* jobs;          // [n_jobs]
    const long long* job_off;    // [n_jobs+1] output row offsets
    double* xy;                  // [2*rows]
    double* heading;
    double* curvature;
    double* alpha_total;
    double* alpha_last;
    double* v;
    double* ax;
    rl_job_stats* stats;         // [n_jobs]
    unsigned long long* dbg;     // debug-checks build only (else null): [0] failures, [1] first failure (code | CTA << 32), [2] fault injection
    int no_few_search;           // test hook (rl_set_option "no_few_search"): the cluster kernel rebuilds flagged samples by tile streaming only
};

// Size classes: one CTA of T threads solves one (track, config, stage) job, K samples per thread.
struct SizeClass {
    int T;
    int K;
};
constexpr int kNumClasses = 8;
// the last three entries are the classes with FOUR samples per thread for tracks of up to 1024 samples (round 2): twice
// the threads per track, half the projected-gradient loop -- a loop that fits the instruction cache, and fewer, larger CTAs
// at different places of the code per SM.  class_for_n prefers them when every thread still owns two samples.
constexpr SizeClass kClasses[kNumClasses] = {{32, 8}, {64, 8}, {128, 8}, {256, 8}, {512, 8}, {64, 4}, {128, 4}, {256, 4}};
constexpr int kSegBlock = 8;  // segments per bounding box in the corridor ray-cast
constexpr int kSupBlock = 8;  // boxes per super box (fast corridor path)
constexpr int kFirstK4Class = 5;
constexpr int kSmallTwoWarpClass = 5;

inline int class_for_n(int n)
{
#ifdef RL_NO_MID_K4      // A/B builds: only the two-warp class of the small maps
    constexpr int kEndK4 = kFirstK4Class + 1;
#else
    constexpr int kEndK4 = kNumClasses;
#endif
    for (int c = kFirstK4Class; c < kEndK4; ++c)     // every thread of a multi-warp CTA must own at least two samples
        if (n >= 2 * kClasses[c].T && n <= kClasses[c].T * kClasses[c].K) return c;
    for (int c = 0; c < kFirstK4Class; ++c)
        if (n <= kClasses[c].T * kClasses[c].K) return c;
    return -1;
}

// scratch area (barrier, reduction and halo exchange buffers, a few ints); the debug-checks and phase-timer builds keep
// more words there.  The product size lets the T = 128 class keep four CTAs per SM next to the kernels' static shared
// memory (job constants, 384 bytes).
#if defined(RL_DEBUG_CHECKS) || defined(RL_PHASE_TIMERS)
constexpr int kScratchBytes = 2048;
#else
constexpr int kScratchBytes = 1664;
#endif
// dynamic shared memory of one CTA of class (T,K): see layout in raceline_kernels.cu
inline size_t smem_bytes_for_class(int T, int K)
{
    const size_t np = (size_t)T * K;
    size_t b = np * 16      /* path points, double2            */
               + np * 8 * 4 /* region B: PGD coefficients+stash / ray tile+corridor staging */
               + kScratchBytes /* barriers, reduction + halo exchange scratch */
               + np * 6;    /* per-sample corridor state: anchor segments + parity bits (4 B), clearances (2 B) */
#ifdef RL_DEBUG_CHECKS
    b += 4 * 64 + 2048;     /* four guard zones + the per-thread phase counters (raceline_kernels.cuh) */
#endif
    return b;
}

// Long tracks (N > 4096): one thread-block cluster of `cs` CTAs (256 threads x 8 samples) per job, see
// raceline_cluster.cuh.  Every CTA of the cluster needs between 512 and 2048 samples.  Returns 0 when no cluster
// size up to `max_cs` fits.  `force` > 0 (test hook RL_FORCE_CLUSTER) asks for exactly that size.
constexpr int kMaxClusterSize = 16;
int cluster_max_size();                  // 16 when the device schedules 16-CTA (non-portable) clusters of the kernel, else 8
constexpr int kClusterClassBase = 100;   // ClassList.cls = kClusterClassBase + cs for cluster launches
inline int cluster_size_for_n(long long n, int force = 0, int max_cs = 8)
{
    const int sizes[4] = {2, 4, 8, 16};
    for (int i = 0; i < 4; ++i) {
        const int cs = sizes[i];
        if (cs > max_cs) break;
        if (force > 0 && cs != force) continue;
        if (force <= 0 && cs == 2) continue;   // N <= 4096 belongs to the single-CTA kernels
        if (n >= 512ll * cs && n <= 2048ll * cs) return cs;
    }
    return 0;
}
// mode 0 = ragged chunks, 1 = every chunk holds exactly 2048 samples.  Returns a cudaError_t as int.
int launch_solve_cluster(const DevBatch& B, const int* job_list, const int* item_off, int n_items, int cs, int mode, void* stream);
int configure_solve_cluster();

// ---- the stage before the path (raceline_geom.cu): centre line + width/geometry rows, batched over tracks ----
struct GeomBatch {
    const long long* mid_off;   // [n_tracks+1]
    const double* mids_xy;
    const int* samples;
    const int* closed;
    const long long* seg_off;   // [2*n_tracks+1]
    const double* seg;
    const long long* row_off;   // [n_tracks+1]
    double kappa_eps, a_lat_max, v_cap;
    int emit_dup;
    double* xy; double* s_rel; double* heading; double* curvature;
    double* dist_inner; double* dist_outer; double* width; double* v_kappa;
    double* track_L; double* track_s0;
};

size_t geom_centerline_smem(int max_pts);
int geom_max_points();          // padded mid points one track may have
int configure_geom();
int launch_geom(const GeomBatch& G, int n_tracks, int max_pts, void* stream);   // returns a cudaError_t as int

// launches job_list[0..n_list) (indices into B.jobs) with the kernel of class `cls` and mode
// 0 = closed track, 1 = closed track and every job has N == T*K, 2 = open track.  One CTA per ITEM: item k is the chain
// job_list[item_off[k] .. item_off[k+1]) of jobs on the same track (item_off has n_items+1 entries, relative to
// job_list); a cluster class gives each item to one thread-block cluster.  Returns a cudaError_t as int.
int launch_solve(const DevBatch& B, const int* job_list, int n_list, const int* item_off, int n_items, int cls, int mode, void* stream);
// CTAs of class `cls` one SM holds at a time: the occupancy the runtime reports for the kernel (set by configure_kernels)
int ctas_per_sm(int cls);
int configure_kernels();  // sets max dynamic shared memory on every instantiation
int launch_fp64_peak(double* d_out, int blocks, int threads, int iters, void* stream);

}  // namespace rl
