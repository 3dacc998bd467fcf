/*
 * raceline_b200.h -- C ABI of the B200-native batched raceline solver.
 *
 * This is the drop-in boundary for the two solver stages of
 * tjsdn3065/Practice_path_planning_for_formula_student_driverless:
 *
 *   raceline_min_curv::compute_min_curvature_raceline   (src/main.cpp:683-764)
 *   raceline_min_time::compute_min_time_raceline        (src/main.cpp:905-1052)
 *
 * The reference has no FFI/plugin layer; its boundary is those two in-TU
 * functions, called from pipeline::compute_raceline_and_save (main.cpp:1347)
 * and pipeline::compute_mintime_and_save (main.cpp:1397).  Every entry point
 * below names the reference interface it replaces.  Plain pointers and sizes
 * only; nothing throws across this boundary; there is no CPU fallback -- a
 * call without a usable sm_100a device returns RL_ERR_NODEVICE / RL_ERR_CUDA.
 *
 * Arithmetic is IEEE FP64 throughout (the reference is double-only).
 */
#ifndef RACELINE_B200_H
#define RACELINE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RL_ABI_VERSION 2

/* job stages (one solve of the BASELINE metric = one MINCURV job + one MINTIME job) */
#define RL_STAGE_MINCURV 1 /* compute_min_curvature_raceline, main.cpp:683 */
#define RL_STAGE_MINTIME 2 /* compute_min_time_raceline,      main.cpp:905 */
/* profile of a GIVEN path, no optimisation: heading_curv_from_points_generic (main.cpp:595) with h = L/N, then
 * velocity_profile_forward_backward (main.cpp:782) -> heading, curvature, v, ax, lap_time; xy = the input path, alpha = 0.
 * This is what the reference's debug block computes for the centre line and the min-curv path (main.cpp:1464-1477). */
#define RL_STAGE_EVAL 3

/* per-outer-iteration log depth kept in rl_job_stats (cfg max_outer_iters defaults to 14) */
#define RL_MAX_OUTER_LOG 32

/* status codes (0 = ok, negative = error) */
#define RL_OK 0
#define RL_ERR_ARG (-1)         /* null pointer, negative size, index out of range */
#define RL_ERR_CUDA (-2)        /* a CUDA runtime call or kernel failed */
#define RL_ERR_UNSUPPORTED (-3) /* problem shape outside what the kernels cover */
#define RL_ERR_NOMEM (-4)       /* host or device allocation failed */
#define RL_ERR_NODEVICE (-5)    /* no CUDA device / not an sm_100 part */

/*
 * Config fields the hot path reads (cfg::Config, main.cpp:47-119), one struct
 * per (track x config) problem so a Config sweep is just an array of these.
 * Defaults: rl_default_params().  NB the reference evaluates
 * a_total_max = mu*9.81 once at construction (main.cpp:102); a sweep over mu
 * must set a_total_max itself.
 */
typedef struct rl_params {
    double veh_width_arg;    /* the solver's veh_width argument: initial corridor (main.cpp:706, 930) */
    double veh_width_m;      /* cfg veh_width_m: corridor updates (main.cpp:753, 1037)               */
    double safety_margin_m;  /* main.cpp:78  */
    double lambda_smooth;    /* main.cpp:81  */
    double step_init;        /* main.cpp:84  */
    double step_min;         /* main.cpp:85  */
    double armijo_c;         /* main.cpp:86  */
    double kappa_eps;        /* main.cpp:89  */
    double v_cap_mps;        /* main.cpp:90  */
    double mass_kg;          /* main.cpp:93  */
    double Cd;               /* main.cpp:94  */
    double A_front_m2;       /* main.cpp:95  */
    double rho_air;          /* main.cpp:96  */
    double c_rr;             /* main.cpp:97  */
    double P_max_W;          /* main.cpp:98  */
    double a_total_max;      /* main.cpp:102 */
    double a_lat_max;        /* main.cpp:103 */
    double a_long_acc_cap;   /* main.cpp:104 */
    double a_long_brake_cap; /* main.cpp:105 */
    double w_time_gain;      /* main.cpp:108 */
    double time_gamma_power; /* main.cpp:109 */
    double inv_v_gain;       /* main.cpp:111 */
    int32_t max_outer_iters; /* main.cpp:82  */
    int32_t max_inner_iters; /* main.cpp:83  */
    int32_t max_vpass_iters; /* main.cpp:112 */
    int32_t time_weight_use_inv_v; /* main.cpp:110 */
    int32_t use_total_ge_lat;      /* main.cpp:113 */
    int32_t reserved;
} rl_params;

/* one unit of work: solve stage `stage` of track `track` under params[param] */
typedef struct rl_job {
    int32_t track;
    int32_t param;
    int32_t stage; /* RL_STAGE_MINCURV or RL_STAGE_MINTIME */
    int32_t reserved;
} rl_job;

/*
 * Per-job counters.  The reference only prints these to stderr ("[GN k] J0=",
 * main.cpp:725; "[MT k] J0= .. t=", 998; "[PG] .. bt=", 1019-1022) and logs no
 * backtracks at all for min-curv; here they are returned.
 */
typedef struct rl_job_stats {
    int32_t status;       /* RL_OK or an error for this job                    */
    int32_t n;            /* samples                                            */
    int32_t outer_done;   /* outer linearisations executed                      */
    int32_t accepted;     /* accepted PGD steps, all outers (main.cpp:735)      */
    int32_t backtracks;   /* Armijo step halvings, all outers (main.cpp:737)    */
    int32_t evals;        /* cost/grad evaluations (main.cpp:724, 732)          */
    int32_t vpass_rounds; /* device diagnostic: relaxation rounds of the v(s) passes */
    int32_t exist_scans;  /* device diagnostic: full "does this ray hit the ring at all" searches (main.cpp:696 semantics) */
    int64_t ray_tests;    /* device diagnostic: exact FP64 ray/segment tests run */
    double lap_time;      /* min-time: final predicted lap (main.cpp:1047); else 0 */
    double J0[RL_MAX_OUTER_LOG];        /* cost at alpha=0 per outer (main.cpp:724, 997)  */
    double Jend[RL_MAX_OUTER_LOG];      /* cost after the last accepted step per outer    */
    double lap_outer[RL_MAX_OUTER_LOG]; /* min-time: VP.lap_time per outer (main.cpp:998) */
    int32_t acc_outer[RL_MAX_OUTER_LOG]; /* accepted steps per outer */
    int32_t bt_outer[RL_MAX_OUTER_LOG];  /* backtracks per outer     */
} rl_job_stats;

/*
 * A batch of problems in packed (CSR-style) layout.  All pointers are HOST
 * pointers for the host entry points (pinned memory makes the copies
 * asynchronous) and are only read.
 *
 *   track t: centre samples  center_xy[2*samp_off[t] .. 2*samp_off[t+1])   (x,y interleaved,
 *            the reference's vector<Vec2>, closing duplicate already dropped, main.cpp:1681-1683)
 *            inner segments  seg[4*seg_off[2t]   .. 4*seg_off[2t+1])       (x0,y0,x1,y1 per segment,
 *            outer segments  seg[4*seg_off[2t+1] .. 4*seg_off[2t+2])        the reference's
 *                                                                           vector<pair<Vec2,Vec2>>,
 *                                                                           edges::ringEdges main.cpp:251)
 *            L = track_L[t] (CL.L), closed = track_closed[t] (cfg is_closed_track)
 */
typedef struct rl_batch_desc {
    int32_t n_tracks;
    int32_t n_params;
    int32_t n_jobs;
    int32_t reserved;
    const int64_t* samp_off;      /* [n_tracks+1]   */
    const int64_t* seg_off;       /* [2*n_tracks+1] */
    const double* center_xy;      /* [2*samp_off[n_tracks]]   */
    const double* seg;            /* [4*seg_off[2*n_tracks]]  */
    const double* track_L;        /* [n_tracks] */
    const int32_t* track_closed;  /* [n_tracks] */
    const rl_params* params;      /* [n_params] */
    const rl_job* jobs;           /* [n_jobs]   */
} rl_batch_desc;

/*
 * Caller-allocated outputs (the fields of the reference's Result structs,
 * main.cpp:677-681 and 897-903).  Per-sample arrays hold the jobs' samples back
 * to back in job order: job j owns rows [off[j], off[j+1]) with
 * off = rl_job_sample_offsets().  v/ax rows of MINCURV jobs are left untouched.
 * Any per-sample pointer may be NULL to skip that download.
 */
typedef struct rl_batch_out {
    double* xy;           /* raceline, x,y interleaved   [2*rows] */
    double* heading;      /* [rows] */
    double* curvature;    /* [rows] */
    double* alpha_total;  /* [rows] */
    double* alpha_last;   /* [rows] */
    double* v;            /* [rows] (MINTIME jobs) */
    double* ax;           /* [rows] (MINTIME jobs) */
    rl_job_stats* stats;  /* [n_jobs] */
} rl_batch_out;

typedef struct rl_ctx rl_ctx;     /* owns the device, streams, staging */
typedef struct rl_batch rl_batch; /* a device-resident batch */

/* ---- library / context ------------------------------------------------- */
int rl_abi_version(void);
const char* rl_status_string(int status);
int rl_device_count(void);
/* cfg::Config defaults for the fields above (main.cpp:77-113) */
int rl_default_params(rl_params* p);
/*
 * One context per device; replaces the process-global cfg::get() (main.cpp:120).
 * Threading contract: every entry point that takes a context (or a batch made from it) is internally serialised by a
 * lock inside the context, so host threads MAY share one; calls then run one after the other.  For concurrent solves
 * use one context per thread (contexts share nothing).  A batch may be destroyed after its context.
 */
rl_ctx* rl_create(int device, int* status);
void rl_destroy(rl_ctx* ctx);
/* run on a caller-owned cudaStream_t (e.g. torch's current stream); NULL = the context's own stream */
int rl_set_stream(rl_ctx* ctx, void* cuda_stream);
/*
 * Tuning knobs and test hooks of the host plan (0 = automatic):
 *   "solve_chunks"  pipeline chunks of rl_solve_batch (1..16)
 *   "chunk_streams" how those chunks are spread over the kernel streams: 1 = round robin over the device's stream
 *                   priority levels, 2 = a stream per chunk with priorities that never rise from one chunk to the
 *                   next, 3 = streams of one priority (order of launch only)
 *   "geom_chunks"   track ranges rl_centerline_geom_batch pipelines (upload | kernels | download); automatic: one per
 *                   1024 tracks, at most 8; 1 = one upload, the kernels, one download
 *   "max_chain"     longest chain of consecutive jobs on one track that one CTA / cluster works through
 *   "force_chain"   form chains of exactly this length whatever the batch size (tests)
 *   "force_cluster" route closed tracks of any length through the cluster kernel with this many CTAs (tests)
 *   "no_few_search" test hook: 1 = the long-track kernel rebuilds flagged samples by tile streaming only (results
 *                   are bit-identical either way; tests compare the two)
 *   "debug_inject"  debug-checks build only, see rl_debug_check_failures
 * Unknown names return RL_ERR_ARG.  The library reads no environment variables.
 */
int rl_set_option(rl_ctx* ctx, const char* name, int64_t value);
const char* rl_last_error(rl_ctx* ctx);

/* page-locked host memory for asynchronous copies (cudaHostAlloc / cudaFreeHost) */
void* rl_host_alloc(size_t bytes);
void rl_host_free(void* p);

/* ---- layout helpers ---------------------------------------------------- */
/* off[j] = sum of N(track(job k)) for k<j; off has n_jobs+1 entries */
int rl_job_sample_offsets(const rl_batch_desc* desc, int64_t* off);
/* Which kernel a track of n samples gets (pure host logic, no device needed; for planning and tests): *threads and
 * *samples_per_thread of the size class of the single-CTA kernels (n <= 4096), or the CTAs of the thread-block cluster
 * (256 threads x 8 samples each) in *cluster_ctas for longer tracks (open or closed), else 0 there.  Returns RL_OK,
 * RL_ERR_UNSUPPORTED when no kernel covers the shape (more than 32,768 samples), RL_ERR_ARG for n < 0 or
 * null pointers.  max_cluster_ctas: 8 (portable cluster size) or 16. */
int rl_plan_for_track(int64_t n_samples, int32_t closed, int32_t max_cluster_ctas, int32_t* threads, int32_t* samples_per_thread,
                      int32_t* cluster_ctas);

/* ---- batched solve: the data-parallel form of main.cpp:1347 + main.cpp:1397 */
/* host buffers in, host buffers out; H2D, kernels and D2H are chunk-pipelined inside */
int rl_solve_batch(rl_ctx* ctx, const rl_batch_desc* desc, const rl_batch_out* out);

/* device-resident form: upload once, solve many times, download on demand */
rl_batch* rl_batch_create(rl_ctx* ctx, const rl_batch_desc* desc, int* status);
/* async H2D of new VALUES for the shapes the batch was created with: per-track sample / segment counts, closed flags
 * and every job's (track, stage) must equal those given to rl_batch_create (RL_ERR_ARG otherwise; the plan -- size
 * classes, output rows, chains -- is frozen at creation); coordinates, L, params and job.param may change */
int rl_batch_upload(rl_batch* b, const rl_batch_desc* desc);
int rl_batch_solve(rl_batch* b);                               /* async kernel launches          */
int rl_batch_download(rl_batch* b, const rl_batch_out* out);   /* async D2H                      */
int rl_batch_sync(rl_batch* b);                                /* wait; returns first CUDA error */
/* DEVICE pointers of the batch's output arrays (same layout as rl_batch_out, rows in job order), for device-side
 * consumers: the final gather of lap times and rasters over NCCL / peer copies reads them in place.  Valid until
 * the batch is destroyed; written by rl_batch_solve on the context's stream.  The v / ax rows of MINCURV jobs hold
 * no result (the kernels may use them as scratch). */
int rl_batch_device_outputs(rl_batch* b, rl_batch_out* device_pointers);
int rl_batch_launches_per_solve(const rl_batch* b);            /* kernels one rl_batch_solve launches */
void rl_batch_destroy(rl_batch* b);

/* ---- single-problem entry points with the reference's argument meaning -- */
/*
 * compute_min_curvature_raceline(center, innerE, outerE, veh_width, L, closed) -> Result
 * (main.cpp:683-686).  veh_width/L/closed as in the reference; Config via `p`
 * (p->veh_width_arg is overwritten by veh_width).  n == 0 returns RL_OK with
 * nothing written, like the reference's empty Result (main.cpp:689).
 */
int rl_compute_min_curvature_raceline(rl_ctx* ctx, const double* center_xy, int n,
                                      const double* inner_seg, int m_inner,
                                      const double* outer_seg, int m_outer,
                                      double veh_width, double L, int closed, const rl_params* p,
                                      double* raceline_xy, double* heading, double* curvature,
                                      double* alpha_total, double* alpha_last, rl_job_stats* stats);
/*
 * compute_min_time_raceline(center, innerE, outerE, veh_width, L, closed) -> Result
 * (main.cpp:905-909); adds v, ax and lap_time (main.cpp:897-903).
 */
int rl_compute_min_time_raceline(rl_ctx* ctx, const double* center_xy, int n,
                                 const double* inner_seg, int m_inner,
                                 const double* outer_seg, int m_outer,
                                 double veh_width, double L, int closed, const rl_params* p,
                                 double* raceline_xy, double* heading, double* curvature,
                                 double* alpha_total, double* alpha_last, double* v, double* ax,
                                 double* lap_time, rl_job_stats* stats);

/* ---- the stage before the path: centre line + width/geometry ------------------- */
/*
 * pipeline::make_centerline (main.cpp:1270-1279) = centerline::splineUniformResample (main.cpp:448-474) over
 * Spline1D::fit / eval (main.cpp:404-446), followed by the per-sample body of pipeline::compute_geom_and_save
 * (main.cpp:1306-1329): spline position and derivatives, heading, curvature, distancesToRings (main.cpp:513-524),
 * width and v_kappa -- the columns s,x,y,heading_rad,curvature,dist_to_inner,dist_to_outer,width,v_kappa_mps of
 * <base>_with_geom.csv (main.cpp:1304).  Batched over tracks; the reference's host code keeps the Delaunay/MST
 * ordering that produces the mid points and the CSV writer that consumes the rows.
 *
 *   track t: ordered mid points  mids_xy[2*mid_off[t] .. 2*mid_off[t+1])   (OM.ordered, at least 3 per track)
 *            samples[t] = cfg samples (main.cpp:1639), closed[t] = closed_mode; paddingK = closed ? 3 : 0 (main.cpp:1273)
 *            rings as in rl_batch_desc (edges::ringEdges / polylineEdges of the *_from_mids points)
 *   rows of track t: closed ? samples : samples + emit_closed_duplicate   (Kmax of main.cpp:1308)
 *   the first samples[t] rows of x,y are the centre line the solver stages take (center_for_opt, main.cpp:1681-1683)
 */
typedef struct rl_geom_desc {
    int32_t n_tracks;
    int32_t emit_closed_duplicate; /* cfg emit_closed_duplicate (main.cpp:55) */
    const int64_t* mid_off;        /* [n_tracks+1]   */
    const double* mids_xy;         /* [2*mid_off[n_tracks]] */
    const int32_t* samples;        /* [n_tracks]     */
    const int32_t* track_closed;   /* [n_tracks]     */
    const int64_t* seg_off;        /* [2*n_tracks+1] */
    const double* seg;             /* [4*seg_off[2*n_tracks]] */
    const rl_params* params;       /* one Config: kappa_eps, a_lat_max, v_cap_mps (main.cpp:1325-1326) */
} rl_geom_desc;

typedef struct rl_geom_out {       /* rows back to back in track order, off = rl_geom_row_offsets() */
    double* xy;          /* [2*rows] */
    double* s_rel;       /* [rows] si - s0 */
    double* heading;     /* [rows] */
    double* curvature;   /* [rows] */
    double* dist_inner;  /* [rows] */
    double* dist_outer;  /* [rows] */
    double* width;       /* [rows] */
    double* v_kappa;     /* [rows] */
    double* track_L;     /* [n_tracks] CL.L  */
    double* track_s0;    /* [n_tracks] CL.s0 */
} rl_geom_out;

int rl_geom_row_offsets(const rl_geom_desc* desc, int64_t* off); /* off has n_tracks+1 entries */
/* host buffers in, host buffers out (any rl_geom_out pointer may be NULL); at most 4090 mid points per track */
int rl_centerline_geom_batch(rl_ctx* ctx, const rl_geom_desc* desc, const rl_geom_out* out);

/* ---- measurement helpers (bench / tests; not on the solve path) -------- */
/*
 * Deterministic synthetic closed tracks: track id
 * first_id+k is drawn from mt19937_64(seed_base + id).  Writes n_samples centre
 * points, m_per_ring inner and m_per_ring outer ring segments and L per track
 * into packed arrays of exactly the rl_batch_desc layout.  Host only.
 */
int rl_synth_tracks(uint64_t seed_base, int64_t first_id, int n_tracks, int n_samples,
                    int m_per_ring, int n_threads, double* center_xy, double* seg, double* track_L);
/*
 * Debug-checks build only (library compiled with -DRL_DEBUG_CHECKS; RL_ERR_UNSUPPORTED otherwise): the kernels verify
 * the hand-over protocol of their shared-memory regions, guard zones between the regions and the bounds of the views
 * they carve out (csrc/raceline_kernels.cuh).  out[0] = failed checks since the context was created, out[1] = the
 * first one (check code | CTA index << 32).  The option "debug_inject" = 1 makes one warp skip a hand-over so that a
 * test can see the checker fire.
 */
int rl_debug_check_failures(rl_ctx* ctx, uint64_t* out2);
/* device time (CUDA events on the context's stream) of the kernels of the last rl_centerline_geom_batch call on this
 * context, in milliseconds, copies excluded; -1 before the first call and after a call that ran as a pipeline of track
 * ranges (kernels and copies overlap there; option "geom_chunks" = 1 gives the unpipelined call) */
double rl_last_kernel_ms(rl_ctx* ctx);
/* measured DFMA throughput of this device in TFLOP/s (2 flop per FMA); the FP64 roofline denominator */
int rl_measure_fp64_peak(rl_ctx* ctx, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* RACELINE_B200_H */
