// raceline_api.cu -- the extern "C" ABI of include/raceline_b200.h over the kernels in raceline_kernels.cu.
// Host side only: validation, packing of job lists per size class, device buffers, copies, launches.
// There is no CPU fallback: without a usable device every entry point returns an error status.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "raceline_device.h"

using rl::DevBatch;

constexpr int kMaxChunks = 16;
// how the chunks of rl_solve_batch are spread over the kernel streams (see solve_batch_pipeline): 1, 2 or 3
#ifndef RL_DEFAULT_CHUNK_STREAMS
#define RL_DEFAULT_CHUNK_STREAMS 2
#endif
constexpr int kDefaultChunkStreams = RL_DEFAULT_CHUNK_STREAMS;

struct rl_ctx {
    int device = 0;
    int n_sm = 148;
    // every entry point that touches the context's streams, scratch batch or error string holds this lock: a context
    // may be shared by host threads (calls are serialised); use one context per thread for concurrency
    std::recursive_mutex mu;
    // tuning knobs and test hooks (rl_set_option); 0 = automatic
    int opt_solve_chunks = 0, opt_chunk_streams = 0, opt_geom_chunks = 0, opt_max_chain = 0, opt_force_chain = 0, opt_force_cluster = 0, opt_no_few_search = 0;
    bool pipeline_ready = false;
    struct GeomBufs* geom = nullptr;   // device buffers of rl_centerline_geom_batch, kept between calls (grow only)
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;   // timing events around the kernels of the last geometry call
    float last_kernel_ms = -1.f;
    unsigned long long* d_dbg = nullptr;   // debug-checks build: [0] failures, [1] first failure, [2] fault injection switch
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;   // own_stream or the caller's
    std::string err;
    rl_batch* scratch = nullptr;     // reused by rl_solve_batch and the single-problem entry points
    // pipeline of rl_solve_batch: H2D | kernels | D2H.  One compute stream per chunk, with DESCENDING priority:
    // chunk k+1 fills the SMs chunk k's tail leaves idle, but never delays chunk k (whose results can then go
    // down while chunk k+1 still computes).  Equal priorities would share the SMs and finish all chunks last.
    cudaStream_t s_in = nullptr, s_out = nullptr, s_k[kMaxChunks] = {};
    int n_prio = 1;
    // a D2H copy into PAGEABLE memory blocks the host until it has run, which would serialise the whole pipeline;
    // the per-job stats therefore land in this pinned staging area first
    rl_job_stats* h_stats = nullptr;
    size_t h_stats_cap = 0;
    cudaEvent_t ev_start = nullptr, ev_in[kMaxChunks] = {}, ev_k[kMaxChunks] = {}, ev_end[kMaxChunks + 2] = {};
};

namespace {

template <class T>
struct DevArr {
    T* p = nullptr;
    size_t cap = 0;   // elements
    cudaError_t ensure(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 8 + 16;
        cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct ClassList { int cls; int mode; int begin; int count; int chunk; int item_begin; int n_items; };

}  // namespace

struct GeomBufs {
    DevArr<long long> mid_off, seg_off, row_off;
    DevArr<double> mids, seg, out, L, s0;
    DevArr<int> samples, closed;
    void release()
    {
        mid_off.release(); seg_off.release(); row_off.release(); mids.release(); seg.release(); out.release();
        L.release(); s0.release(); samples.release(); closed.release();
    }
};

struct rl_batch {
    rl_ctx* ctx = nullptr;
    int n_tracks = 0, n_params = 0, n_jobs = 0;
    long long total_samples = 0, total_segs = 0, rows = 0;
    // device inputs
    DevArr<long long> d_samp_off, d_seg_off, d_job_off;
    DevArr<double> d_center, d_seg, d_L;
    DevArr<int> d_closed, d_joblist, d_itemoff;
    DevArr<rl_params> d_params;
    DevArr<rl_job> d_jobs;
    // device outputs
    DevArr<double> d_xy, d_heading, d_curv, d_atot, d_alast, d_v, d_ax;
    DevArr<rl_job_stats> d_stats;
    // host-side plan
    std::vector<long long> job_off;
    std::vector<int> joblist, itemoff;
    std::vector<ClassList> lists;
    std::vector<std::pair<int, int>> skipped;   // (job, status) for jobs no kernel covers
    int n_chunks = 1;
    std::vector<int> chunk_job0;                // [n_chunks+1] job ranges of the chunks
    std::vector<int> chunk_tmax;                // [n_chunks] highest track index a chunk touches
    // what the plan was made from (rl_batch_upload must bring the same shapes, see upload_inputs)
    std::vector<long long> plan_samp_off, plan_seg_off;
    std::vector<int> plan_closed;
    std::vector<rl_job> plan_jobs;
    // row runs [first, count) of the jobs that produce v / ax (MINTIME, EVAL), merged; per chunk [vrun0[c], vrun0[c+1])
    std::vector<std::pair<long long, long long>> vruns;
    std::vector<int> vrun0;
    int device = 0;
    void release()
    {
        d_samp_off.release(); d_seg_off.release(); d_job_off.release(); d_center.release(); d_seg.release();
        d_L.release(); d_closed.release(); d_joblist.release(); d_itemoff.release(); d_params.release(); d_jobs.release();
        d_xy.release(); d_heading.release(); d_curv.release(); d_atot.release(); d_alast.release();
        d_v.release(); d_ax.release(); d_stats.release();
    }
};

namespace {

int c_sm(const rl_batch* b);

// Length of the job chains of one launch.  A chain of c jobs on one track saves every job after the first the corridor
// search of its first build (measured: about a fifth of a job), but the launch has c times fewer items to spread over
// the CTA slots of the device.  Pick the c that minimises the modelled time: one item's cost(c) = c - (c - 1) * saving
// job units, times the waves the items need -- items / slots plus half a wave of tail when there are several waves
// (CTAs finish at different times and, in the pipelined path, the next chunk's launch fills the tail), exactly one when
// all items are resident at once.  `max_run` = the longest run of consecutive jobs on one track.
int pick_chain(size_t n_jobs, int slots, int max_chain, int max_run)
{
    constexpr double kChainSaving = 0.2;
    int best = 1;
    double best_cost = 1e300;
    for (int c = 1; c <= std::max(1, std::min(max_chain, max_run)); ++c) {
        const double items = std::ceil((double)n_jobs / c);
        const double waves = (items <= (double)slots) ? 1.0 : items / (double)std::max(1, slots) + 0.5;
        const double cost = waves * ((double)c - (double)(c - 1) * kChainSaving);
        if (cost < best_cost * (1.0 - 1e-9)) { best_cost = cost; best = c; }
    }
    return best;
}
template <class Jobs>
int longest_run(const rl_batch_desc* d, const Jobs& list)
{
    int best = 0, run = 0, last_t = -1;
    for (size_t q = 0; q < list.size(); ++q) {
        const rl_job& jb = d->jobs[list[q]];
        if (jb.stage == RL_STAGE_EVAL) { run = 0; last_t = -1; continue; }
        run = (jb.track == last_t) ? run + 1 : 1;
        last_t = jb.track;
        best = std::max(best, run);
    }
    return best;
}

int fail(rl_ctx* c, int status, const std::string& msg)
{
    if (c) c->err = msg;
    return status;
}
int cuda_fail(rl_ctx* c, cudaError_t e, const char* what)
{
    return fail(c, RL_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define RL_CUDA(ctx, call)                                          \
    do {                                                            \
        cudaError_t e__ = (call);                                   \
        if (e__ != cudaSuccess) return cuda_fail((ctx), e__, #call); \
    } while (0)

int validate_desc(rl_ctx* c, const rl_batch_desc* d)
{
    if (!d) return fail(c, RL_ERR_ARG, "null batch descriptor");
    if (d->n_tracks < 0 || d->n_params < 0 || d->n_jobs < 0) return fail(c, RL_ERR_ARG, "negative count");
    if (d->n_jobs == 0) return RL_OK;
    if (!d->samp_off || !d->seg_off || !d->track_L || !d->track_closed || !d->params || !d->jobs)
        return fail(c, RL_ERR_ARG, "null array in batch descriptor");
    if (d->n_tracks == 0 || d->n_params == 0) return fail(c, RL_ERR_ARG, "jobs without tracks or params");
    if (d->samp_off[0] != 0 || d->seg_off[0] != 0) return fail(c, RL_ERR_ARG, "offset arrays must start at 0");
    for (int t = 0; t < d->n_tracks; ++t) {
        if (d->samp_off[t + 1] < d->samp_off[t]) return fail(c, RL_ERR_ARG, "samp_off not monotone");
        if (d->seg_off[2 * t + 1] < d->seg_off[2 * t] || d->seg_off[2 * t + 2] < d->seg_off[2 * t + 1])
            return fail(c, RL_ERR_ARG, "seg_off not monotone");
        if (d->samp_off[t + 1] - d->samp_off[t] > (1ll << 30)) return fail(c, RL_ERR_ARG, "track too long");
    }
    if (d->samp_off[d->n_tracks] > 0 && !d->center_xy) return fail(c, RL_ERR_ARG, "null center_xy");
    if (d->seg_off[2 * d->n_tracks] > 0 && !d->seg) return fail(c, RL_ERR_ARG, "null seg");
    for (int j = 0; j < d->n_jobs; ++j) {
        const rl_job& jb = d->jobs[j];
        if (jb.track < 0 || jb.track >= d->n_tracks) return fail(c, RL_ERR_ARG, "job.track out of range");
        if (jb.param < 0 || jb.param >= d->n_params) return fail(c, RL_ERR_ARG, "job.param out of range");
        if (jb.stage != RL_STAGE_MINCURV && jb.stage != RL_STAGE_MINTIME && jb.stage != RL_STAGE_EVAL) return fail(c, RL_ERR_ARG, "job.stage invalid");
    }
    return RL_OK;
}

// size the buffers, classify the jobs, upload the (small) plan arrays
int plan_batch(rl_batch* b, const rl_batch_desc* d, int n_chunks = 1)
{
    rl_ctx* c = b->ctx;
    b->n_tracks = d->n_tracks; b->n_params = d->n_params; b->n_jobs = d->n_jobs;
    b->total_samples = d->n_tracks ? d->samp_off[d->n_tracks] : 0;
    b->total_segs = d->n_tracks ? d->seg_off[2 * d->n_tracks] : 0;
    b->job_off.assign((size_t)d->n_jobs + 1, 0);
    for (int j = 0; j < d->n_jobs; ++j) {
        const int t = d->jobs[j].track;
        b->job_off[j + 1] = b->job_off[j] + (d->samp_off[t + 1] - d->samp_off[t]);
    }
    b->rows = b->job_off[d->n_jobs];

    // job lists per chunk and (class, mode), in batch order; consecutive jobs on the same track form one ITEM (one CTA
    // works through the chain and keeps what it learnt about the track's corridor), as long as enough items remain to
    // fill the device several times over
    b->lists.clear(); b->joblist.clear(); b->itemoff.clear(); b->skipped.clear();
    n_chunks = std::max(1, std::min(n_chunks, std::min(kMaxChunks, d->n_jobs)));
    b->n_chunks = n_chunks;
    b->chunk_job0.assign((size_t)n_chunks + 1, 0);
    b->chunk_tmax.assign((size_t)n_chunks, -1);
    const int max_chain = c->opt_max_chain > 0 ? c->opt_max_chain : 8;   // rl_set_option("max_chain")
    const int force_chain = std::max(0, c->opt_force_chain);             // test hook: chains of that length whatever the batch size
    for (int c = 0; c <= n_chunks; ++c) {
        // The first chunk's upload and the last chunk's download are the only copies no kernel hides: with enough
        // chunks those two are a quarter of the size of the others.
        // a chunk boundary never separates the jobs of one track (they form a chain and share the track's upload)
        long long num = c, den = n_chunks;
        if (n_chunks >= 6) { den = 4ll * n_chunks - 6; num = (c == 0) ? 0 : (c == n_chunks ? den : 4ll * c - 3); }
        int j = (int)(((long long)d->n_jobs * num) / den);
        if (c > 0 && c < n_chunks) {
            j = std::max(j, b->chunk_job0[c - 1]);
            for (int moved = 0; moved < max_chain && j > b->chunk_job0[c - 1] && j < d->n_jobs && d->jobs[j].track == d->jobs[j - 1].track; ++moved) ++j;
        }
        b->chunk_job0[c] = j;
    }
    b->plan_samp_off.assign(d->samp_off, d->samp_off + d->n_tracks + 1);
    b->plan_seg_off.assign(d->seg_off, d->seg_off + 2 * d->n_tracks + 1);
    b->plan_closed.assign(d->track_closed, d->track_closed + d->n_tracks);
    b->plan_jobs.assign(d->jobs, d->jobs + d->n_jobs);
    b->vruns.clear(); b->vrun0.assign((size_t)n_chunks + 1, 0);
    for (int c = 0; c < n_chunks; ++c) {
        b->vrun0[c] = (int)b->vruns.size();
        for (int j = b->chunk_job0[c]; j < b->chunk_job0[c + 1]; ++j) {
            if (d->jobs[j].stage == RL_STAGE_MINCURV || b->job_off[j + 1] == b->job_off[j]) continue;
            if ((int)b->vruns.size() > b->vrun0[c] && b->vruns.back().first + b->vruns.back().second == b->job_off[j])
                b->vruns.back().second += b->job_off[j + 1] - b->job_off[j];
            else b->vruns.push_back({b->job_off[j], b->job_off[j + 1] - b->job_off[j]});
        }
    }
    b->vrun0[n_chunks] = (int)b->vruns.size();
    for (int c = 0; c < n_chunks; ++c) {
        std::vector<std::vector<int>> bucket(rl::kNumClasses * 3);
        std::vector<std::vector<int>> cbucket(3 * (rl::kMaxClusterSize + 1));   // [cs][mode: ragged | exact fit | open] cluster launches
        const int force_cs = b->ctx->opt_force_cluster;   // test hook: cluster kernel on shorter tracks
        for (int j = b->chunk_job0[c]; j < b->chunk_job0[c + 1]; ++j) {
            const rl_job& jb = d->jobs[j];
            const int t = jb.track;
            b->chunk_tmax[c] = std::max(b->chunk_tmax[c], t);
            const long long n = d->samp_off[t + 1] - d->samp_off[t];
            if (n == 0) { b->skipped.push_back({j, RL_OK}); continue; }   // empty Result, main.cpp:689 / 912
            const int cls = rl::class_for_n((int)n);
            const int cs = (cls < 0 || force_cs > 0) ? rl::cluster_size_for_n(n, force_cs, rl::cluster_max_size()) : 0;
            if (cs > 0) { cbucket[3 * cs + (!d->track_closed[t] ? 2 : (n == 2048ll * cs ? 1 : 0))].push_back(j); continue; }
            if (cls < 0) { b->skipped.push_back({j, RL_ERR_UNSUPPORTED}); continue; }
            const bool exact = (n == (long long)rl::kClasses[cls].T * rl::kClasses[cls].K);
            const int mode = !d->track_closed[t] ? 2 : (exact ? 1 : 0);   // open | closed exact-fit | closed ragged
            bucket[cls * 3 + mode].push_back(j);
        }
        for (int k = 0; k < rl::kNumClasses * 3; ++k) {
            if (bucket[k].empty()) continue;
            const int slots = std::max(1, c_sm(b) * rl::ctas_per_sm(k / 3));
            int chain = pick_chain(bucket[k].size(), slots, max_chain, longest_run(d, bucket[k]));
            if (force_chain > 0) chain = force_chain;
            ClassList l = {k / 3, k % 3, (int)b->joblist.size(), (int)bucket[k].size(), c, (int)b->itemoff.size(), 0};
            int run = 0, last_t = -1;
            for (size_t q = 0; q < bucket[k].size(); ++q) {
                const int t = d->jobs[bucket[k][q]].track;
                const bool eval_job = (d->jobs[bucket[k][q]].stage == RL_STAGE_EVAL);     // no corridor state to share
                if (q == 0 || t != last_t || run >= chain || eval_job) { b->itemoff.push_back((int)q); ++l.n_items; run = 0; }
                last_t = eval_job ? -1 : t; ++run;
            }
            b->itemoff.push_back((int)bucket[k].size());
            b->lists.push_back(l);
            b->joblist.insert(b->joblist.end(), bucket[k].begin(), bucket[k].end());
        }
        for (size_t k = 0; k < cbucket.size(); ++k) {
            if (cbucket[k].empty()) continue;
            const int cs = (int)(k / 3);
            const int slots = std::max(1, (c_sm(b) * 2) / cs);      // two CTAs per SM, cs CTAs per cluster
            int chain = pick_chain(cbucket[k].size(), slots, max_chain, longest_run(d, cbucket[k]));
            if (force_chain > 0) chain = force_chain;
            ClassList l = {rl::kClusterClassBase + cs, (int)(k % 3), (int)b->joblist.size(), (int)cbucket[k].size(), c, (int)b->itemoff.size(), 0};
            int run = 0, last_t = -1;
            for (size_t q = 0; q < cbucket[k].size(); ++q) {
                const int t = d->jobs[cbucket[k][q]].track;
                const bool eval_job = (d->jobs[cbucket[k][q]].stage == RL_STAGE_EVAL);
                if (q == 0 || t != last_t || run >= chain || eval_job) { b->itemoff.push_back((int)q); ++l.n_items; run = 0; }
                last_t = eval_job ? -1 : t; ++run;
            }
            b->itemoff.push_back((int)cbucket[k].size());
            b->lists.push_back(l);
            b->joblist.insert(b->joblist.end(), cbucket[k].begin(), cbucket[k].end());
        }
    }

    RL_CUDA(c, b->d_samp_off.ensure((size_t)d->n_tracks + 1));
    RL_CUDA(c, b->d_seg_off.ensure((size_t)2 * d->n_tracks + 1));
    RL_CUDA(c, b->d_job_off.ensure((size_t)d->n_jobs + 1));
    RL_CUDA(c, b->d_center.ensure((size_t)2 * b->total_samples + 2));
    RL_CUDA(c, b->d_seg.ensure((size_t)4 * b->total_segs + 4));
    RL_CUDA(c, b->d_L.ensure((size_t)d->n_tracks));
    RL_CUDA(c, b->d_closed.ensure((size_t)d->n_tracks));
    RL_CUDA(c, b->d_params.ensure((size_t)d->n_params));
    RL_CUDA(c, b->d_jobs.ensure((size_t)d->n_jobs));
    RL_CUDA(c, b->d_joblist.ensure(b->joblist.size() + 1));
    RL_CUDA(c, b->d_itemoff.ensure(b->itemoff.size() + 1));
    const size_t rows = (size_t)b->rows + 2;
    RL_CUDA(c, b->d_xy.ensure(2 * rows));
    RL_CUDA(c, b->d_heading.ensure(rows));
    RL_CUDA(c, b->d_curv.ensure(rows));
    RL_CUDA(c, b->d_atot.ensure(rows));
    RL_CUDA(c, b->d_alast.ensure(rows));
    RL_CUDA(c, b->d_v.ensure(rows));
    RL_CUDA(c, b->d_ax.ensure(rows));
    RL_CUDA(c, b->d_stats.ensure((size_t)d->n_jobs));

    cudaStream_t s = c->stream;
    RL_CUDA(c, cudaMemcpyAsync(b->d_job_off.p, b->job_off.data(), sizeof(long long) * b->job_off.size(), cudaMemcpyHostToDevice, s));
    if (!b->joblist.empty())
        RL_CUDA(c, cudaMemcpyAsync(b->d_joblist.p, b->joblist.data(), sizeof(int) * b->joblist.size(), cudaMemcpyHostToDevice, s));
    if (!b->itemoff.empty())
        RL_CUDA(c, cudaMemcpyAsync(b->d_itemoff.p, b->itemoff.data(), sizeof(int) * b->itemoff.size(), cudaMemcpyHostToDevice, s));
    RL_CUDA(c, cudaMemsetAsync(b->d_stats.p, 0, sizeof(rl_job_stats) * (size_t)d->n_jobs, s));
    for (auto& sk : b->skipped) {
        const int t = d->jobs[sk.first].track;
        int hdr[2] = {sk.second, (int)(d->samp_off[t + 1] - d->samp_off[t])};
        RL_CUDA(c, cudaMemcpyAsync(&b->d_stats.p[sk.first], hdr, sizeof(hdr), cudaMemcpyHostToDevice, s));
        RL_CUDA(c, cudaStreamSynchronize(s));   // hdr is a stack temporary
    }
    // job_off/joblist live in std::vectors owned by the batch: safe for the async copies above
    return RL_OK;
}

int upload_inputs(rl_batch* b, const rl_batch_desc* d)
{
    rl_ctx* c = b->ctx;
    cudaStream_t s = c->stream;
    // The host plan (size class and mode of every job, output row offsets, chains) was frozen by rl_batch_create and the
    // kernels size their shared-memory use by it: the descriptor must describe the SAME shapes -- per-track sample and
    // segment counts, closed flags, and the (track, stage) of every job.  Values (coordinates, L, params, and which
    // params a job uses) may change.
    if (d->n_tracks != b->n_tracks || d->n_params != b->n_params || d->n_jobs != b->n_jobs)
        return fail(c, RL_ERR_ARG, "rl_batch_upload: counts differ from rl_batch_create");
    if (d->n_jobs == 0) return RL_OK;
    if (std::memcmp(d->samp_off, b->plan_samp_off.data(), sizeof(long long) * ((size_t)d->n_tracks + 1)) != 0 ||
        std::memcmp(d->seg_off, b->plan_seg_off.data(), sizeof(long long) * ((size_t)2 * d->n_tracks + 1)) != 0)
        return fail(c, RL_ERR_ARG, "rl_batch_upload: per-track sample / segment counts differ from rl_batch_create");
    for (int t = 0; t < d->n_tracks; ++t)
        if ((d->track_closed[t] != 0) != (b->plan_closed[t] != 0))
            return fail(c, RL_ERR_ARG, "rl_batch_upload: track_closed differs from rl_batch_create");
    for (int j = 0; j < d->n_jobs; ++j)
        if (d->jobs[j].track != b->plan_jobs[j].track || d->jobs[j].stage != b->plan_jobs[j].stage)
            return fail(c, RL_ERR_ARG, "rl_batch_upload: a job's track or stage differs from rl_batch_create");
    RL_CUDA(c, cudaMemcpyAsync(b->d_samp_off.p, d->samp_off, sizeof(long long) * ((size_t)d->n_tracks + 1), cudaMemcpyHostToDevice, s));
    RL_CUDA(c, cudaMemcpyAsync(b->d_seg_off.p, d->seg_off, sizeof(long long) * ((size_t)2 * d->n_tracks + 1), cudaMemcpyHostToDevice, s));
    if (b->total_samples)
        RL_CUDA(c, cudaMemcpyAsync(b->d_center.p, d->center_xy, sizeof(double) * 2 * (size_t)b->total_samples, cudaMemcpyHostToDevice, s));
    if (b->total_segs)
        RL_CUDA(c, cudaMemcpyAsync(b->d_seg.p, d->seg, sizeof(double) * 4 * (size_t)b->total_segs, cudaMemcpyHostToDevice, s));
    RL_CUDA(c, cudaMemcpyAsync(b->d_L.p, d->track_L, sizeof(double) * (size_t)d->n_tracks, cudaMemcpyHostToDevice, s));
    RL_CUDA(c, cudaMemcpyAsync(b->d_closed.p, d->track_closed, sizeof(int) * (size_t)d->n_tracks, cudaMemcpyHostToDevice, s));
    RL_CUDA(c, cudaMemcpyAsync(b->d_params.p, d->params, sizeof(rl_params) * (size_t)d->n_params, cudaMemcpyHostToDevice, s));
    RL_CUDA(c, cudaMemcpyAsync(b->d_jobs.p, d->jobs, sizeof(rl_job) * (size_t)d->n_jobs, cudaMemcpyHostToDevice, s));
    return RL_OK;
}

int c_sm(const rl_batch* b) { return b->ctx ? b->ctx->n_sm : 148; }

DevBatch dev_view(const rl_batch* b)
{
    DevBatch B;
    B.samp_off = b->d_samp_off.p; B.seg_off = b->d_seg_off.p; B.center_xy = b->d_center.p; B.seg = b->d_seg.p;
    B.track_L = b->d_L.p; B.track_closed = b->d_closed.p; B.params = b->d_params.p; B.jobs = b->d_jobs.p;
    B.job_off = b->d_job_off.p; B.xy = b->d_xy.p; B.heading = b->d_heading.p; B.curvature = b->d_curv.p;
    B.alpha_total = b->d_atot.p; B.alpha_last = b->d_alast.p; B.v = b->d_v.p; B.ax = b->d_ax.p; B.stats = b->d_stats.p;
    B.dbg = b->ctx ? b->ctx->d_dbg : nullptr;
    B.no_few_search = b->ctx ? b->ctx->opt_no_few_search : 0;
    return B;
}

struct DevFree {
    std::vector<void*> ptrs;
    ~DevFree() { for (void* p : ptrs) cudaFree(p); }
    template <class T> cudaError_t alloc(T** p, size_t n)
    {
        *p = nullptr;
        cudaError_t e = cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T));
        if (e == cudaSuccess) ptrs.push_back(*p);
        return e;
    }
};

static_assert(sizeof(long long) == sizeof(int64_t), "int64 layout");

// D2H of the v / ax rows of chunk `k` (all chunks: k < 0): only the rows of jobs that produce them (MINTIME, EVAL) --
// the rows of MINCURV jobs are left untouched, as include/raceline_b200.h promises.  Equally long, equally spaced runs
// (the usual min-curv / min-time interleave of equally long tracks) travel as ONE strided copy.
int copy_v_rows(rl_ctx* c, const rl_batch* b, int k, double* dst, const double* src, cudaStream_t s)
{
    const int r0 = k < 0 ? 0 : b->vrun0[k], r1 = k < 0 ? (int)b->vruns.size() : b->vrun0[k + 1];
    if (r1 <= r0) return RL_OK;
    const auto* R = b->vruns.data();
    bool periodic = (r1 - r0 >= 3);
    const long long pitch = periodic ? R[r0 + 1].first - R[r0].first : 0;
    for (int r = r0; periodic && r < r1; ++r)
        periodic = (R[r].second == R[r0].second) && (R[r].first == R[r0].first + (long long)(r - r0) * pitch);
    if (periodic) {
        RL_CUDA(c, cudaMemcpy2DAsync(dst + R[r0].first, (size_t)pitch * 8, src + R[r0].first, (size_t)pitch * 8, (size_t)R[r0].second * 8,
                                     (size_t)(r1 - r0), cudaMemcpyDeviceToHost, s));
        return RL_OK;
    }
    for (int r = r0; r < r1; ++r)
        RL_CUDA(c, cudaMemcpyAsync(dst + R[r].first, src + R[r].first, (size_t)R[r].second * 8, cudaMemcpyDeviceToHost, s));
    return RL_OK;
}

}  // namespace

extern "C" {

int rl_abi_version(void) { return RL_ABI_VERSION; }

const char* rl_status_string(int s)
{
    switch (s) {
        case RL_OK: return "ok";
        case RL_ERR_ARG: return "bad argument";
        case RL_ERR_CUDA: return "CUDA error";
        case RL_ERR_UNSUPPORTED: return "unsupported problem shape";
        case RL_ERR_NOMEM: return "out of memory";
        case RL_ERR_NODEVICE: return "no usable CUDA device";
        default: return "unknown status";
    }
}

int rl_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int rl_default_params(rl_params* p)
{
    if (!p) return RL_ERR_ARG;
    std::memset(p, 0, sizeof(*p));
    // cfg::Config defaults, main.cpp:77-113
    p->veh_width_arg = 1.0; p->veh_width_m = 1.0; p->safety_margin_m = 0.05;
    p->lambda_smooth = 1.6e-3; p->step_init = 0.65; p->step_min = 1e-6; p->armijo_c = 1e-5;
    p->kappa_eps = 1e-6; p->v_cap_mps = 27.0;
    p->mass_kg = 255.0; p->Cd = 0.30; p->A_front_m2 = 1.00; p->rho_air = 1.225; p->c_rr = 0.015; p->P_max_W = 80000.0;
    p->a_total_max = 1.17 * 9.81; p->a_lat_max = 11.0; p->a_long_acc_cap = 8.0; p->a_long_brake_cap = 11.0;
    p->w_time_gain = 1.0; p->time_gamma_power = 2.0; p->inv_v_gain = 0.1;
    p->max_outer_iters = 14; p->max_inner_iters = 120; p->max_vpass_iters = 6;
    p->time_weight_use_inv_v = 0; p->use_total_ge_lat = 1;
    return RL_OK;
}

rl_ctx* rl_create(int device, int* status)
{
    int st = RL_OK;
    rl_ctx* c = nullptr;
    int n = 0, n_sm_probe = 148;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0 || device < 0 || device >= n) { cudaGetLastError(); st = RL_ERR_NODEVICE; }
    if (st == RL_OK) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) st = RL_ERR_NODEVICE;   // sm_100a only
        else n_sm_probe = prop.multiProcessorCount;
    }
    if (st == RL_OK && cudaSetDevice(device) != cudaSuccess) st = RL_ERR_CUDA;
    if (st == RL_OK) {
        c = new (std::nothrow) rl_ctx();
        if (!c) st = RL_ERR_NOMEM;
    }
    if (st == RL_OK) {
        c->device = device;
        c->n_sm = n_sm_probe;
        if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) st = RL_ERR_CUDA;
        c->stream = c->own_stream;
        if (st == RL_OK && rl::configure_kernels() != 0) st = RL_ERR_CUDA;
#ifdef RL_DEBUG_CHECKS
        if (st == RL_OK && (cudaMalloc((void**)&c->d_dbg, 3 * sizeof(unsigned long long)) != cudaSuccess ||
                            cudaMemset(c->d_dbg, 0, 3 * sizeof(unsigned long long)) != cudaSuccess)) st = RL_ERR_CUDA;
#endif
        if (st != RL_OK) { if (c->own_stream) cudaStreamDestroy(c->own_stream); delete c; c = nullptr; }
    }
    if (status) *status = st;
    return c;
}

void rl_destroy(rl_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->scratch) { rl_batch_destroy(c->scratch); c->scratch = nullptr; }
    // whatever ensure_pipeline managed to create (it may have stopped half way)
    if (c->s_in) cudaStreamDestroy(c->s_in);
    if (c->s_out) cudaStreamDestroy(c->s_out);
    if (c->ev_start) cudaEventDestroy(c->ev_start);
    for (int i = 0; i < kMaxChunks; ++i) {
        if (c->s_k[i]) cudaStreamDestroy(c->s_k[i]);
        if (c->ev_in[i]) cudaEventDestroy(c->ev_in[i]);
        if (c->ev_k[i]) cudaEventDestroy(c->ev_k[i]);
    }
    for (int i = 0; i < kMaxChunks + 2; ++i) if (c->ev_end[i]) cudaEventDestroy(c->ev_end[i]);
    if (c->h_stats) cudaFreeHost(c->h_stats);
    if (c->geom) { c->geom->release(); delete c->geom; }
    if (c->d_dbg) cudaFree(c->d_dbg);
    if (c->ev_t0) cudaEventDestroy(c->ev_t0);
    if (c->ev_t1) cudaEventDestroy(c->ev_t1);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

int rl_set_stream(rl_ctx* c, void* cuda_stream)
{
    if (!c) return RL_ERR_ARG;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
    return RL_OK;
}

int rl_set_option(rl_ctx* c, const char* name, int64_t value)
{
    if (!c || !name) return RL_ERR_ARG;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    const int v = (int)std::max<int64_t>(0, std::min<int64_t>(value, 1 << 20));
    if (!std::strcmp(name, "solve_chunks")) c->opt_solve_chunks = std::min(v, kMaxChunks);
    else if (!std::strcmp(name, "chunk_streams")) c->opt_chunk_streams = std::min(v, 3);
    else if (!std::strcmp(name, "geom_chunks")) c->opt_geom_chunks = std::min(v, kMaxChunks);
    else if (!std::strcmp(name, "max_chain")) c->opt_max_chain = v;
    else if (!std::strcmp(name, "force_chain")) c->opt_force_chain = v;
    else if (!std::strcmp(name, "force_cluster")) c->opt_force_cluster = v;
    else if (!std::strcmp(name, "no_few_search")) c->opt_no_few_search = v;
    else if (!std::strcmp(name, "debug_inject")) {
        // debug-checks build only: make one warp skip a phase hand-over so that tests can see the checker fire
        if (!c->d_dbg) return fail(c, RL_ERR_UNSUPPORTED, "rl_set_option: debug_inject needs the -DRL_DEBUG_CHECKS build");
        const unsigned long long inj = (unsigned long long)v;
        RL_CUDA(c, cudaSetDevice(c->device));
        RL_CUDA(c, cudaMemcpy(c->d_dbg + 2, &inj, sizeof(inj), cudaMemcpyHostToDevice));
    }
    else return fail(c, RL_ERR_ARG, std::string("rl_set_option: unknown option ") + name);
    return RL_OK;
}

const char* rl_last_error(rl_ctx* c) { return c ? c->err.c_str() : "null context"; }

int rl_plan_for_track(int64_t n, int32_t closed, int32_t max_cs, int32_t* threads, int32_t* spt, int32_t* cluster_ctas)
{
    if (n < 0 || !threads || !spt || !cluster_ctas) return RL_ERR_ARG;
    *threads = 0; *spt = 0; *cluster_ctas = 0;
    const int cls = (n <= (1ll << 30)) ? rl::class_for_n((int)n) : -1;     // the same calls plan_batch makes
    if (cls >= 0) { *threads = rl::kClasses[cls].T; *spt = rl::kClasses[cls].K; return RL_OK; }
    (void)closed;     // open and closed tracks are covered alike
    const int cs = rl::cluster_size_for_n(n, 0, max_cs >= 16 ? 16 : 8);
    if (cs <= 0) return RL_ERR_UNSUPPORTED;
    *threads = 256; *spt = 8; *cluster_ctas = cs;
    return RL_OK;
}

int rl_job_sample_offsets(const rl_batch_desc* d, int64_t* off)
{
    if (!d || !off || d->n_jobs < 0) return RL_ERR_ARG;
    off[0] = 0;
    for (int j = 0; j < d->n_jobs; ++j) {
        const int t = d->jobs[j].track;
        if (t < 0 || t >= d->n_tracks) return RL_ERR_ARG;
        off[j + 1] = off[j] + (d->samp_off[t + 1] - d->samp_off[t]);
    }
    return RL_OK;
}

void* rl_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void rl_host_free(void* p) { if (p) cudaFreeHost(p); }

rl_batch* rl_batch_create(rl_ctx* c, const rl_batch_desc* d, int* status)
{
    int st = c ? validate_desc(c, d) : RL_ERR_ARG;
    rl_batch* b = nullptr;
    std::unique_lock<std::recursive_mutex> lk;
    if (c) lk = std::unique_lock<std::recursive_mutex>(c->mu);
    if (st == RL_OK && cudaSetDevice(c->device) != cudaSuccess) st = RL_ERR_CUDA;
    if (st == RL_OK) {
        b = new (std::nothrow) rl_batch();
        if (!b) st = RL_ERR_NOMEM;
    }
    if (st == RL_OK) {
        b->ctx = c;
        b->device = c->device;
        st = plan_batch(b, d);
        if (st == RL_OK) st = upload_inputs(b, d);
        if (st == RL_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) st = RL_ERR_CUDA;
        if (st != RL_OK) { b->release(); delete b; b = nullptr; }
    }
    if (status) *status = st;
    return b;
}

int rl_batch_upload(rl_batch* b, const rl_batch_desc* d)
{
    if (!b || !d) return RL_ERR_ARG;
    std::lock_guard<std::recursive_mutex> lk(b->ctx->mu);
    const int st = validate_desc(b->ctx, d);
    if (st != RL_OK) return st;
    RL_CUDA(b->ctx, cudaSetDevice(b->device));
    return upload_inputs(b, d);
}

int rl_batch_solve(rl_batch* b)
{
    if (!b) return RL_ERR_ARG;
    rl_ctx* c = b->ctx;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    RL_CUDA(c, cudaSetDevice(c->device));
    const DevBatch B = dev_view(b);
    for (const ClassList& l : b->lists) {
        const int e = rl::launch_solve(B, b->d_joblist.p + l.begin, l.count, b->d_itemoff.p + l.item_begin, l.n_items, l.cls, l.mode, c->stream);
        if (e != 0) return cuda_fail(c, (cudaError_t)e, "solve_kernel launch");
    }
    return RL_OK;
}

int rl_batch_launches_per_solve(const rl_batch* b) { return b ? (int)b->lists.size() : 0; }

int rl_batch_download(rl_batch* b, const rl_batch_out* o)
{
    if (!b || !o) return RL_ERR_ARG;
    rl_ctx* c = b->ctx;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    RL_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const size_t rows = (size_t)b->rows;
    if (rows) {
        if (o->xy) RL_CUDA(c, cudaMemcpyAsync(o->xy, b->d_xy.p, 16 * rows, cudaMemcpyDeviceToHost, s));
        if (o->heading) RL_CUDA(c, cudaMemcpyAsync(o->heading, b->d_heading.p, 8 * rows, cudaMemcpyDeviceToHost, s));
        if (o->curvature) RL_CUDA(c, cudaMemcpyAsync(o->curvature, b->d_curv.p, 8 * rows, cudaMemcpyDeviceToHost, s));
        if (o->alpha_total) RL_CUDA(c, cudaMemcpyAsync(o->alpha_total, b->d_atot.p, 8 * rows, cudaMemcpyDeviceToHost, s));
        if (o->alpha_last) RL_CUDA(c, cudaMemcpyAsync(o->alpha_last, b->d_alast.p, 8 * rows, cudaMemcpyDeviceToHost, s));
        if (o->v) { const int st = copy_v_rows(c, b, -1, o->v, b->d_v.p, s); if (st != RL_OK) return st; }
        if (o->ax) { const int st = copy_v_rows(c, b, -1, o->ax, b->d_ax.p, s); if (st != RL_OK) return st; }
    }
    if (o->stats && b->n_jobs)
        RL_CUDA(c, cudaMemcpyAsync(o->stats, b->d_stats.p, sizeof(rl_job_stats) * (size_t)b->n_jobs, cudaMemcpyDeviceToHost, s));
    return RL_OK;
}

int rl_batch_device_outputs(rl_batch* b, rl_batch_out* dev)
{
    if (!b || !dev) return RL_ERR_ARG;
    dev->xy = b->d_xy.p; dev->heading = b->d_heading.p; dev->curvature = b->d_curv.p; dev->alpha_total = b->d_atot.p;
    dev->alpha_last = b->d_alast.p; dev->v = b->d_v.p; dev->ax = b->d_ax.p; dev->stats = b->d_stats.p;
    return RL_OK;
}

int rl_batch_sync(rl_batch* b)
{
    if (!b) return RL_ERR_ARG;
    std::lock_guard<std::recursive_mutex> lk(b->ctx->mu);
    RL_CUDA(b->ctx, cudaSetDevice(b->device));
    RL_CUDA(b->ctx, cudaStreamSynchronize(b->ctx->stream));
    RL_CUDA(b->ctx, cudaGetLastError());
    return RL_OK;
}

void rl_batch_destroy(rl_batch* b)
{
    if (!b) return;
    cudaSetDevice(b->device);   // the batch remembers its device: destroying it after its context is legal
    b->release();
    delete b;
}

static int ensure_pipeline(rl_ctx* c)
{
    if (c->pipeline_ready) return RL_OK;
    if (c->s_in) return fail(c, RL_ERR_CUDA, "the copy/compute pipeline of this context could not be created earlier");
    RL_CUDA(c, cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
    RL_CUDA(c, cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
    int lo_p = 0, hi_p = 0;   // numerically lowest value = highest priority
    RL_CUDA(c, cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p));
    c->n_prio = std::max(1, std::min(kMaxChunks, lo_p - hi_p + 1));
    for (int i = 0; i < kMaxChunks; ++i)
        RL_CUDA(c, cudaStreamCreateWithPriority(&c->s_k[i], cudaStreamNonBlocking, std::min(lo_p, hi_p + i)));
    RL_CUDA(c, cudaEventCreateWithFlags(&c->ev_start, cudaEventDisableTiming));
    for (int i = 0; i < kMaxChunks; ++i) {
        RL_CUDA(c, cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming));
        RL_CUDA(c, cudaEventCreateWithFlags(&c->ev_k[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < kMaxChunks + 2; ++i) RL_CUDA(c, cudaEventCreateWithFlags(&c->ev_end[i], cudaEventDisableTiming));
    c->pipeline_ready = true;   // only now: a half-built pipeline is never used (rl_destroy frees what exists)
    return RL_OK;
}

// Host buffers in, host buffers out.  The batch is cut into up to kMaxChunks job ranges; chunk k's inputs go up
// while chunk k-1 computes and chunk k-2's results come down, all bracketed by the context's stream (so events
// the caller records on it time the whole call).
// everything rl_solve_batch queues after ev_start; on an error the caller drains the streams before returning
static int solve_batch_pipeline(rl_ctx* c, rl_batch* b, const rl_batch_desc* d, const rl_batch_out* o);

int rl_solve_batch(rl_ctx* c, const rl_batch_desc* d, const rl_batch_out* o)
{
    if (!c || !o) return RL_ERR_ARG;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    int st = validate_desc(c, d);
    if (st != RL_OK) return st;
    if (d->n_jobs == 0) return RL_OK;
    RL_CUDA(c, cudaSetDevice(c->device));
    if (!c->scratch) {
        c->scratch = new (std::nothrow) rl_batch();
        if (!c->scratch) return fail(c, RL_ERR_NOMEM, "host allocation failed");
        c->scratch->ctx = c;
        c->scratch->device = c->device;
    }
    rl_batch* b = c->scratch;
    st = ensure_pipeline(c);
    if (st != RL_OK) return st;
    // chunks of at least ~2 device fills of chained items each (one item = the jobs of one track), at most kMaxChunks
    int want_chunks = std::max(1, std::min(kMaxChunks, d->n_jobs / 2048));
    if (c->opt_solve_chunks > 0) want_chunks = std::min(kMaxChunks, c->opt_solve_chunks);   // rl_set_option("solve_chunks")
    // the small plan arrays travel on the context stream before the pipeline starts
    st = plan_batch(b, d, want_chunks);
    if (st == RL_OK) st = solve_batch_pipeline(c, b, d, o);
    if (st != RL_OK) {
        // copies into the caller's buffers and reads of the caller's inputs may still be in flight on the pipeline
        // streams: nothing may outlive this call
        const std::string keep = c->err;
        if (c->s_in) cudaStreamSynchronize(c->s_in);
        if (c->s_out) cudaStreamSynchronize(c->s_out);
        for (int i = 0; i < kMaxChunks; ++i) if (c->s_k[i]) cudaStreamSynchronize(c->s_k[i]);
        cudaStreamSynchronize(c->stream);
        cudaGetLastError();
        c->err = keep;
    }
    return st;
}

static int solve_batch_pipeline(rl_ctx* c, rl_batch* b, const rl_batch_desc* d, const rl_batch_out* o)
{
    int st = RL_OK;
    cudaStream_t s0 = c->stream;
    RL_CUDA(c, cudaMemcpyAsync(b->d_samp_off.p, d->samp_off, sizeof(long long) * ((size_t)d->n_tracks + 1), cudaMemcpyHostToDevice, s0));
    RL_CUDA(c, cudaMemcpyAsync(b->d_seg_off.p, d->seg_off, sizeof(long long) * ((size_t)2 * d->n_tracks + 1), cudaMemcpyHostToDevice, s0));
    RL_CUDA(c, cudaMemcpyAsync(b->d_L.p, d->track_L, sizeof(double) * (size_t)d->n_tracks, cudaMemcpyHostToDevice, s0));
    RL_CUDA(c, cudaMemcpyAsync(b->d_closed.p, d->track_closed, sizeof(int) * (size_t)d->n_tracks, cudaMemcpyHostToDevice, s0));
    RL_CUDA(c, cudaMemcpyAsync(b->d_params.p, d->params, sizeof(rl_params) * (size_t)d->n_params, cudaMemcpyHostToDevice, s0));
    RL_CUDA(c, cudaMemcpyAsync(b->d_jobs.p, d->jobs, sizeof(rl_job) * (size_t)d->n_jobs, cudaMemcpyHostToDevice, s0));
    RL_CUDA(c, cudaEventRecord(c->ev_start, s0));
    RL_CUDA(c, cudaStreamWaitEvent(c->s_in, c->ev_start, 0));
    RL_CUDA(c, cudaStreamWaitEvent(c->s_out, c->ev_start, 0));
    for (int i = 0; i < kMaxChunks; ++i) RL_CUDA(c, cudaStreamWaitEvent(c->s_k[i], c->ev_start, 0));

    // the per-job counters go straight into the caller's array when it is page-locked (rl_host_alloc); a D2H copy into
    // PAGEABLE memory would block the host until it has run and serialise the pipeline, so those land in a pinned
    // staging area first and are copied over at the end
    rl_job_stats* stats_dst = o->stats;
    if (o->stats) {
        cudaPointerAttributes pa;
        const bool pinned = cudaPointerGetAttributes(&pa, o->stats) == cudaSuccess && pa.type == cudaMemoryTypeHost;
        cudaGetLastError();
        if (!pinned) stats_dst = nullptr;
    }
    if (o->stats && !stats_dst && (size_t)d->n_jobs > c->h_stats_cap) {
        if (c->h_stats) cudaFreeHost(c->h_stats);
        c->h_stats = nullptr; c->h_stats_cap = 0;
        const size_t want = (size_t)d->n_jobs + (size_t)d->n_jobs / 8 + 16;
        RL_CUDA(c, cudaHostAlloc((void**)&c->h_stats, want * sizeof(rl_job_stats), cudaHostAllocDefault));
        c->h_stats_cap = want;
    }
    const DevBatch B = dev_view(b);
    int up_t = 0;   // tracks [0, up_t) are already queued for upload
    size_t li = 0;
    for (int k = 0; k < b->n_chunks; ++k) {
        // ---- inputs of the tracks this chunk is the first to touch ----
        const int t1 = b->chunk_tmax[k] + 1;
        if (t1 > up_t) {
            const long long a0 = d->samp_off[up_t], a1 = d->samp_off[t1], g0 = d->seg_off[2 * up_t], g1 = d->seg_off[2 * t1];
            if (a1 > a0)
                RL_CUDA(c, cudaMemcpyAsync(b->d_center.p + 2 * a0, d->center_xy + 2 * a0, sizeof(double) * 2 * (size_t)(a1 - a0), cudaMemcpyHostToDevice, c->s_in));
            if (g1 > g0)
                RL_CUDA(c, cudaMemcpyAsync(b->d_seg.p + 4 * g0, d->seg + 4 * g0, sizeof(double) * 4 * (size_t)(g1 - g0), cudaMemcpyHostToDevice, c->s_in));
            up_t = t1;
        }
        RL_CUDA(c, cudaEventRecord(c->ev_in[k], c->s_in));
        // ---- kernels ----
        // Chunk k runs on its own stream; the streams' priorities fall over the first n_prio chunks and stay at the lowest
        // level after that, so a later chunk only fills what the earlier ones leave idle (the tail of their last wave)
        // and the chunks FINISH in order.  That order is what the one in-order download stream needs: round 2 first ran
        // chunk k on stream k mod n_prio, where chunk k + n_prio outranks chunks k+1 .. k+n_prio-1 -- the low-priority
        // chunks then finish near the end of the call and every download queued behind theirs waits with them
        // (65,536 tracks: 86 ms of exposed copies per call instead of 9 ms).
        cudaStream_t sk;
        switch (c->opt_chunk_streams ? c->opt_chunk_streams : kDefaultChunkStreams) {   // rl_set_option("chunk_streams")
        case 1: sk = c->s_k[k % c->n_prio]; break;                 // round robin over the priority levels
        case 3: { const int base = c->n_prio - 1; sk = c->s_k[base + k % (kMaxChunks - base)]; } break;   // one priority
        default: sk = c->s_k[k]; break;                            // a stream per chunk, priorities never rise with k
        }
        RL_CUDA(c, cudaStreamWaitEvent(sk, c->ev_in[k], 0));
        for (; li < b->lists.size() && b->lists[li].chunk == k; ++li) {
            const ClassList& l = b->lists[li];
            const int e = rl::launch_solve(B, b->d_joblist.p + l.begin, l.count, b->d_itemoff.p + l.item_begin, l.n_items, l.cls, l.mode, sk);
            if (e != 0) return cuda_fail(c, (cudaError_t)e, "solve_kernel launch");
        }
        RL_CUDA(c, cudaEventRecord(c->ev_k[k], sk));
        // ---- results of this chunk's rows ----
        RL_CUDA(c, cudaStreamWaitEvent(c->s_out, c->ev_k[k], 0));
        const int j0 = b->chunk_job0[k], j1 = b->chunk_job0[k + 1];
        const size_t r0 = (size_t)b->job_off[j0], nr = (size_t)(b->job_off[j1] - b->job_off[j0]);
        if (nr) {
            if (o->xy) RL_CUDA(c, cudaMemcpyAsync(o->xy + 2 * r0, b->d_xy.p + 2 * r0, 16 * nr, cudaMemcpyDeviceToHost, c->s_out));
            if (o->heading) RL_CUDA(c, cudaMemcpyAsync(o->heading + r0, b->d_heading.p + r0, 8 * nr, cudaMemcpyDeviceToHost, c->s_out));
            if (o->curvature) RL_CUDA(c, cudaMemcpyAsync(o->curvature + r0, b->d_curv.p + r0, 8 * nr, cudaMemcpyDeviceToHost, c->s_out));
            if (o->alpha_total) RL_CUDA(c, cudaMemcpyAsync(o->alpha_total + r0, b->d_atot.p + r0, 8 * nr, cudaMemcpyDeviceToHost, c->s_out));
            if (o->alpha_last) RL_CUDA(c, cudaMemcpyAsync(o->alpha_last + r0, b->d_alast.p + r0, 8 * nr, cudaMemcpyDeviceToHost, c->s_out));
            if (o->v) { st = copy_v_rows(c, b, k, o->v, b->d_v.p, c->s_out); if (st != RL_OK) return st; }
            if (o->ax) { st = copy_v_rows(c, b, k, o->ax, b->d_ax.p, c->s_out); if (st != RL_OK) return st; }
        }
        if (o->stats && j1 > j0)
            RL_CUDA(c, cudaMemcpyAsync((stats_dst ? stats_dst : c->h_stats) + j0, b->d_stats.p + j0, sizeof(rl_job_stats) * (size_t)(j1 - j0), cudaMemcpyDeviceToHost, c->s_out));
    }
    // join everything back into the context stream
    for (int i = 0; i < kMaxChunks + 2; ++i) {
        cudaStream_t tail = (i < kMaxChunks) ? c->s_k[i] : (i == kMaxChunks ? c->s_in : c->s_out);
        RL_CUDA(c, cudaEventRecord(c->ev_end[i], tail));
        RL_CUDA(c, cudaStreamWaitEvent(s0, c->ev_end[i], 0));
    }
    st = rl_batch_sync(b);
    if (st != RL_OK) return st;
    if (o->stats && !stats_dst) std::memcpy(o->stats, c->h_stats, sizeof(rl_job_stats) * (size_t)d->n_jobs);
    for (auto& sk : b->skipped)
        if (sk.second != RL_OK) return fail(c, sk.second, "a job's shape is not covered by the kernels (N too large)");
    return RL_OK;
}

static int solve_single(rl_ctx* c, int stage, const double* center_xy, int n, const double* inner_seg, int m_inner,
                        const double* outer_seg, int m_outer, double veh_width, double L, int closed, const rl_params* p,
                        double* xy, double* heading, double* curvature, double* alpha_total, double* alpha_last,
                        double* v, double* ax, double* lap_time, rl_job_stats* stats)
{
    if (!c || !p || n < 0 || m_inner < 0 || m_outer < 0) return RL_ERR_ARG;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    if (n > 0 && !center_xy) return RL_ERR_ARG;
    if ((m_inner > 0 && !inner_seg) || (m_outer > 0 && !outer_seg)) return RL_ERR_ARG;
    rl_job_stats local;
    std::memset(&local, 0, sizeof(local));
    if (n == 0) {   // the reference returns an empty Result (main.cpp:689 / 912)
        if (stats) *stats = local;
        if (lap_time) *lap_time = 0.0;
        return RL_OK;
    }
    std::vector<double> seg((size_t)4 * ((size_t)m_inner + m_outer) + 4);
    if (m_inner) std::memcpy(seg.data(), inner_seg, sizeof(double) * 4 * (size_t)m_inner);
    if (m_outer) std::memcpy(seg.data() + 4 * (size_t)m_inner, outer_seg, sizeof(double) * 4 * (size_t)m_outer);
    rl_params pp = *p;
    pp.veh_width_arg = veh_width;
    const int64_t samp_off[2] = {0, n};
    const int64_t seg_off[3] = {0, m_inner, (int64_t)m_inner + m_outer};
    const int32_t cl = closed ? 1 : 0;
    const rl_job job = {0, 0, stage, 0};
    rl_batch_desc d;
    std::memset(&d, 0, sizeof(d));
    d.n_tracks = 1; d.n_params = 1; d.n_jobs = 1;
    d.samp_off = samp_off; d.seg_off = seg_off; d.center_xy = center_xy; d.seg = seg.data();
    d.track_L = &L; d.track_closed = &cl; d.params = &pp; d.jobs = &job;
    rl_batch_out o;
    std::memset(&o, 0, sizeof(o));
    o.xy = xy; o.heading = heading; o.curvature = curvature; o.alpha_total = alpha_total; o.alpha_last = alpha_last;
    o.v = v; o.ax = ax; o.stats = &local;
    const int st = rl_solve_batch(c, &d, &o);
    if (stats) *stats = local;
    if (lap_time) *lap_time = local.lap_time;
    return st;
}

int rl_compute_min_curvature_raceline(rl_ctx* ctx, const double* center_xy, int n, const double* inner_seg, int m_inner,
                                      const double* outer_seg, int m_outer, double veh_width, double L, int closed,
                                      const rl_params* p, double* raceline_xy, double* heading, double* curvature,
                                      double* alpha_total, double* alpha_last, rl_job_stats* stats)
{
    return solve_single(ctx, RL_STAGE_MINCURV, center_xy, n, inner_seg, m_inner, outer_seg, m_outer, veh_width, L, closed, p,
                        raceline_xy, heading, curvature, alpha_total, alpha_last, nullptr, nullptr, nullptr, stats);
}

int rl_compute_min_time_raceline(rl_ctx* ctx, const double* center_xy, int n, const double* inner_seg, int m_inner,
                                 const double* outer_seg, int m_outer, double veh_width, double L, int closed,
                                 const rl_params* p, double* raceline_xy, double* heading, double* curvature,
                                 double* alpha_total, double* alpha_last, double* v, double* ax, double* lap_time,
                                 rl_job_stats* stats)
{
    return solve_single(ctx, RL_STAGE_MINTIME, center_xy, n, inner_seg, m_inner, outer_seg, m_outer, veh_width, L, closed, p,
                        raceline_xy, heading, curvature, alpha_total, alpha_last, v, ax, lap_time, stats);
}

int rl_geom_row_offsets(const rl_geom_desc* d, int64_t* off)
{
    if (!d || !off || d->n_tracks < 0 || (d->n_tracks > 0 && (!d->samples || !d->track_closed))) return RL_ERR_ARG;
    off[0] = 0;
    for (int t = 0; t < d->n_tracks; ++t) {
        if (d->samples[t] < 1) return RL_ERR_ARG;
        off[t + 1] = off[t] + d->samples[t] + ((!d->track_closed[t] && d->emit_closed_duplicate) ? 1 : 0);   // Kmax, main.cpp:1308
    }
    return RL_OK;
}

// rl_centerline_geom_batch on a large batch: the tracks are cut into n_chunks ranges; range k's mid points and rings go up
// while range k-1's two kernels run and range k-2's rows come down (the streams and events of rl_solve_batch's pipeline;
// a kernel stream per range, so the ranges finish in order).  The per-track arrays are already on the device; the
// kernels index them by track, everything else by the absolute offsets they hold, so a range is the same GeomBatch
// with the per-track pointers moved up.
static int geom_pipeline(rl_ctx* c, const rl_geom_desc* d, const rl_geom_out* o, const rl::GeomBatch& G,
                         const std::vector<long long>& row_off, int n_chunks, int max_pts)
{
    cudaStream_t s = c->stream;
    const int nt = d->n_tracks;
    RL_CUDA(c, cudaEventRecord(c->ev_start, s));
    RL_CUDA(c, cudaStreamWaitEvent(c->s_in, c->ev_start, 0));
    RL_CUDA(c, cudaStreamWaitEvent(c->s_out, c->ev_start, 0));
    for (int k = 0; k < n_chunks; ++k) RL_CUDA(c, cudaStreamWaitEvent(c->s_k[k], c->ev_start, 0));
    for (int k = 0; k < n_chunks; ++k) {
        const int t0 = (int)(((long long)nt * k) / n_chunks), t1 = (int)(((long long)nt * (k + 1)) / n_chunks);
        if (t1 <= t0) continue;
        const long long m0 = d->mid_off[t0], m1 = d->mid_off[t1], g0 = d->seg_off[2 * t0], g1 = d->seg_off[2 * t1];
        RL_CUDA(c, cudaMemcpyAsync(const_cast<double*>(G.mids_xy) + 2 * m0, d->mids_xy + 2 * m0, sizeof(double) * 2 * (size_t)(m1 - m0), cudaMemcpyHostToDevice, c->s_in));
        if (g1 > g0)
            RL_CUDA(c, cudaMemcpyAsync(const_cast<double*>(G.seg) + 4 * g0, d->seg + 4 * g0, sizeof(double) * 4 * (size_t)(g1 - g0), cudaMemcpyHostToDevice, c->s_in));
        RL_CUDA(c, cudaEventRecord(c->ev_in[k], c->s_in));
        cudaStream_t sk = c->s_k[k];
        RL_CUDA(c, cudaStreamWaitEvent(sk, c->ev_in[k], 0));
        rl::GeomBatch Gk = G;
        Gk.mid_off += t0; Gk.samples += t0; Gk.closed += t0; Gk.seg_off += 2 * t0; Gk.row_off += t0;
        Gk.track_L += t0; Gk.track_s0 += t0;
        const int e = rl::launch_geom(Gk, t1 - t0, max_pts, sk);
        if (e != 0) return cuda_fail(c, (cudaError_t)e, "geometry kernels");
        RL_CUDA(c, cudaEventRecord(c->ev_k[k], sk));
        RL_CUDA(c, cudaStreamWaitEvent(c->s_out, c->ev_k[k], 0));
        const size_t r0 = (size_t)row_off[t0], nr = (size_t)(row_off[t1] - row_off[t0]);
        if (nr) {
            if (o->xy) RL_CUDA(c, cudaMemcpyAsync(o->xy + 2 * r0, G.xy + 2 * r0, 16 * nr, cudaMemcpyDeviceToHost, c->s_out));
            if (o->s_rel) RL_CUDA(c, cudaMemcpyAsync(o->s_rel + r0, G.s_rel + r0, 8 * nr, cudaMemcpyDeviceToHost, c->s_out));
            if (o->heading) RL_CUDA(c, cudaMemcpyAsync(o->heading + r0, G.heading + r0, 8 * nr, cudaMemcpyDeviceToHost, c->s_out));
            if (o->curvature) RL_CUDA(c, cudaMemcpyAsync(o->curvature + r0, G.curvature + r0, 8 * nr, cudaMemcpyDeviceToHost, c->s_out));
            if (o->dist_inner) RL_CUDA(c, cudaMemcpyAsync(o->dist_inner + r0, G.dist_inner + r0, 8 * nr, cudaMemcpyDeviceToHost, c->s_out));
            if (o->dist_outer) RL_CUDA(c, cudaMemcpyAsync(o->dist_outer + r0, G.dist_outer + r0, 8 * nr, cudaMemcpyDeviceToHost, c->s_out));
            if (o->width) RL_CUDA(c, cudaMemcpyAsync(o->width + r0, G.width + r0, 8 * nr, cudaMemcpyDeviceToHost, c->s_out));
            if (o->v_kappa) RL_CUDA(c, cudaMemcpyAsync(o->v_kappa + r0, G.v_kappa + r0, 8 * nr, cudaMemcpyDeviceToHost, c->s_out));
        }
    }
    for (int i = 0; i < kMaxChunks + 2; ++i) {
        cudaStream_t tail = (i < kMaxChunks) ? c->s_k[i] : (i == kMaxChunks ? c->s_in : c->s_out);
        RL_CUDA(c, cudaEventRecord(c->ev_end[i], tail));
        RL_CUDA(c, cudaStreamWaitEvent(s, c->ev_end[i], 0));
    }
    // the two per-track columns come down once, at the end: callers tend to hold them in pageable memory, and a copy
    // into pageable memory inside the loop would block the host until that range's kernels are done (no pipeline left)
    if (o->track_L) RL_CUDA(c, cudaMemcpyAsync(o->track_L, G.track_L, sizeof(double) * (size_t)nt, cudaMemcpyDeviceToHost, s));
    if (o->track_s0) RL_CUDA(c, cudaMemcpyAsync(o->track_s0, G.track_s0, sizeof(double) * (size_t)nt, cudaMemcpyDeviceToHost, s));
    RL_CUDA(c, cudaStreamSynchronize(s));
    RL_CUDA(c, cudaGetLastError());
    c->last_kernel_ms = -1.f;   // copies and kernels overlap: no separate kernel time (geom_chunks = 1 measures it)
    return RL_OK;
}

// pipeline::make_centerline + the per-sample body of pipeline::compute_geom_and_save (main.cpp:1270-1335), batched
int rl_centerline_geom_batch(rl_ctx* c, const rl_geom_desc* d, const rl_geom_out* o)
{
    if (!c || !d || !o) return RL_ERR_ARG;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    if (d->n_tracks < 0) return fail(c, RL_ERR_ARG, "negative track count");
    if (d->n_tracks == 0) return RL_OK;
    if (!d->mid_off || !d->mids_xy || !d->samples || !d->track_closed || !d->seg_off || !d->params)
        return fail(c, RL_ERR_ARG, "null array in geometry descriptor");
    const int nt = d->n_tracks;
    if (d->mid_off[0] != 0 || d->seg_off[0] != 0) return fail(c, RL_ERR_ARG, "offset arrays must start at 0");
    int max_pts = 0;
    for (int t = 0; t < nt; ++t) {
        const long long nm = d->mid_off[t + 1] - d->mid_off[t];
        if (nm < 3) return fail(c, RL_ERR_ARG, "a track needs at least 3 ordered mid points (main.cpp:452)");
        if (d->seg_off[2 * t + 1] < d->seg_off[2 * t] || d->seg_off[2 * t + 2] < d->seg_off[2 * t + 1]) return fail(c, RL_ERR_ARG, "seg_off not monotone");
        const long long pts = nm + (d->track_closed[t] ? 6 : 0);
        if (pts > rl::geom_max_points()) return fail(c, RL_ERR_UNSUPPORTED, "too many mid points for one track");
        max_pts = std::max(max_pts, (int)pts);
    }
    if (d->seg_off[2 * nt] > 0 && !d->seg) return fail(c, RL_ERR_ARG, "null seg");
    std::vector<long long> row_off((size_t)nt + 1);
    {
        std::vector<int64_t> ro((size_t)nt + 1);
        const int st = rl_geom_row_offsets(d, ro.data());
        if (st != RL_OK) return fail(c, st, "samples must be >= 1");
        for (int t = 0; t <= nt; ++t) row_off[t] = ro[t];
    }
    RL_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const size_t rows = (size_t)row_off[nt], n_mid = (size_t)d->mid_off[nt], n_seg = (size_t)d->seg_off[2 * nt];
    if (!c->geom) {
        c->geom = new (std::nothrow) GeomBufs();
        if (!c->geom) return fail(c, RL_ERR_NOMEM, "host allocation failed");
    }
    if (!c->ev_t0) {
        RL_CUDA(c, cudaEventCreate(&c->ev_t0));
        RL_CUDA(c, cudaEventCreate(&c->ev_t1));
    }
    GeomBufs& gb = *c->geom;     // no allocation in the steady state: the buffers only ever grow
    RL_CUDA(c, gb.mid_off.ensure((size_t)nt + 1));
    RL_CUDA(c, gb.seg_off.ensure((size_t)2 * nt + 1));
    RL_CUDA(c, gb.row_off.ensure((size_t)nt + 1));
    RL_CUDA(c, gb.mids.ensure(2 * n_mid + 2));
    RL_CUDA(c, gb.seg.ensure(4 * n_seg + 4));
    RL_CUDA(c, gb.samples.ensure((size_t)nt));
    RL_CUDA(c, gb.closed.ensure((size_t)nt));
    RL_CUDA(c, gb.out.ensure(9 * rows + 16));      // xy (2), s_rel, heading, curvature, dist_inner, dist_outer, width, v_kappa
    RL_CUDA(c, gb.L.ensure((size_t)nt));
    RL_CUDA(c, gb.s0.ensure((size_t)nt));
    long long *d_mid_off = gb.mid_off.p, *d_seg_off = gb.seg_off.p, *d_row_off = gb.row_off.p;
    double *d_mids = gb.mids.p, *d_seg = gb.seg.p, *d_out = gb.out.p, *d_L = gb.L.p, *d_s0 = gb.s0.p;
    int *d_samples = gb.samples.p, *d_closed = gb.closed.p;
    RL_CUDA(c, cudaMemcpyAsync(d_mid_off, d->mid_off, sizeof(long long) * ((size_t)nt + 1), cudaMemcpyHostToDevice, s));
    RL_CUDA(c, cudaMemcpyAsync(d_seg_off, d->seg_off, sizeof(long long) * ((size_t)2 * nt + 1), cudaMemcpyHostToDevice, s));
    RL_CUDA(c, cudaMemcpyAsync(d_row_off, row_off.data(), sizeof(long long) * ((size_t)nt + 1), cudaMemcpyHostToDevice, s));
    // large batches run as a pipeline of track ranges (geom_pipeline below); rl_set_option("geom_chunks"): 0 = automatic
    const int n_chunks = std::min(nt, c->opt_geom_chunks > 0 ? c->opt_geom_chunks : std::max(1, std::min(8, nt / 1024)));
    if (n_chunks <= 1) {
        RL_CUDA(c, cudaMemcpyAsync(d_mids, d->mids_xy, sizeof(double) * 2 * n_mid, cudaMemcpyHostToDevice, s));
        if (n_seg) RL_CUDA(c, cudaMemcpyAsync(d_seg, d->seg, sizeof(double) * 4 * n_seg, cudaMemcpyHostToDevice, s));
    }
    RL_CUDA(c, cudaMemcpyAsync(d_samples, d->samples, sizeof(int) * (size_t)nt, cudaMemcpyHostToDevice, s));
    RL_CUDA(c, cudaMemcpyAsync(d_closed, d->track_closed, sizeof(int) * (size_t)nt, cudaMemcpyHostToDevice, s));
    rl::GeomBatch G;
    G.mid_off = d_mid_off; G.mids_xy = d_mids; G.samples = d_samples; G.closed = d_closed; G.seg_off = d_seg_off; G.seg = d_seg;
    G.row_off = d_row_off; G.kappa_eps = d->params->kappa_eps; G.a_lat_max = d->params->a_lat_max; G.v_cap = d->params->v_cap_mps;
    G.emit_dup = d->emit_closed_duplicate;
    G.xy = d_out; G.s_rel = d_out + 2 * rows; G.heading = G.s_rel + rows; G.curvature = G.heading + rows;
    G.dist_inner = G.curvature + rows; G.dist_outer = G.dist_inner + rows; G.width = G.dist_outer + rows; G.v_kappa = G.width + rows;
    G.track_L = d_L; G.track_s0 = d_s0;
    if (n_chunks > 1) {
        int st = ensure_pipeline(c);
        if (st == RL_OK) st = geom_pipeline(c, d, o, G, row_off, n_chunks, max_pts);
        if (st != RL_OK) {   // nothing queued by this call may outlive it (the caller's buffers)
            const std::string keep = c->err;
            if (c->s_in) cudaStreamSynchronize(c->s_in);
            if (c->s_out) cudaStreamSynchronize(c->s_out);
            for (int i = 0; i < kMaxChunks; ++i) if (c->s_k[i]) cudaStreamSynchronize(c->s_k[i]);
            cudaStreamSynchronize(s);
            cudaGetLastError();
            c->err = keep;
        }
        return st;
    }
    RL_CUDA(c, cudaEventRecord(c->ev_t0, s));
    const int e = rl::launch_geom(G, nt, max_pts, s);
    if (e != 0) return cuda_fail(c, (cudaError_t)e, "geometry kernels");
    RL_CUDA(c, cudaEventRecord(c->ev_t1, s));
    if (rows) {
        if (o->xy) RL_CUDA(c, cudaMemcpyAsync(o->xy, G.xy, 16 * rows, cudaMemcpyDeviceToHost, s));
        if (o->s_rel) RL_CUDA(c, cudaMemcpyAsync(o->s_rel, G.s_rel, 8 * rows, cudaMemcpyDeviceToHost, s));
        if (o->heading) RL_CUDA(c, cudaMemcpyAsync(o->heading, G.heading, 8 * rows, cudaMemcpyDeviceToHost, s));
        if (o->curvature) RL_CUDA(c, cudaMemcpyAsync(o->curvature, G.curvature, 8 * rows, cudaMemcpyDeviceToHost, s));
        if (o->dist_inner) RL_CUDA(c, cudaMemcpyAsync(o->dist_inner, G.dist_inner, 8 * rows, cudaMemcpyDeviceToHost, s));
        if (o->dist_outer) RL_CUDA(c, cudaMemcpyAsync(o->dist_outer, G.dist_outer, 8 * rows, cudaMemcpyDeviceToHost, s));
        if (o->width) RL_CUDA(c, cudaMemcpyAsync(o->width, G.width, 8 * rows, cudaMemcpyDeviceToHost, s));
        if (o->v_kappa) RL_CUDA(c, cudaMemcpyAsync(o->v_kappa, G.v_kappa, 8 * rows, cudaMemcpyDeviceToHost, s));
    }
    if (o->track_L) RL_CUDA(c, cudaMemcpyAsync(o->track_L, d_L, sizeof(double) * (size_t)nt, cudaMemcpyDeviceToHost, s));
    if (o->track_s0) RL_CUDA(c, cudaMemcpyAsync(o->track_s0, d_s0, sizeof(double) * (size_t)nt, cudaMemcpyDeviceToHost, s));
    RL_CUDA(c, cudaStreamSynchronize(s));
    RL_CUDA(c, cudaGetLastError());
    c->last_kernel_ms = -1.f;
    cudaEventElapsedTime(&c->last_kernel_ms, c->ev_t0, c->ev_t1);
    return RL_OK;
}

int rl_debug_check_failures(rl_ctx* c, uint64_t* out2)
{
    if (!c || !out2) return RL_ERR_ARG;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    if (!c->d_dbg) return RL_ERR_UNSUPPORTED;    // the product build carries no checks
    RL_CUDA(c, cudaSetDevice(c->device));
    RL_CUDA(c, cudaDeviceSynchronize());
    unsigned long long h[2] = {0, 0};
    RL_CUDA(c, cudaMemcpy(h, c->d_dbg, sizeof(h), cudaMemcpyDeviceToHost));
    out2[0] = h[0]; out2[1] = h[1];
    return RL_OK;
}

double rl_last_kernel_ms(rl_ctx* c)
{
    if (!c) return -1.0;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    return (double)c->last_kernel_ms;
}

int rl_measure_fp64_peak(rl_ctx* c, double* tflops)
{
    if (!c || !tflops) return RL_ERR_ARG;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    RL_CUDA(c, cudaSetDevice(c->device));
    cudaDeviceProp prop;
    RL_CUDA(c, cudaGetDeviceProperties(&prop, c->device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 15;
    double* d = nullptr;
    RL_CUDA(c, cudaMalloc((void**)&d, sizeof(double) * (size_t)blocks * threads));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) {
        if (e0) cudaEventDestroy(e0);
        cudaFree(d);
        return cuda_fail(c, cudaGetLastError(), "fp64 probe events");
    }
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0, c->stream);
        const int e = rl::launch_fp64_peak(d, blocks, threads, iters, c->stream);
        cudaEventRecord(e1, c->stream);
        if (e != 0 || cudaEventSynchronize(e1) != cudaSuccess) {
            const cudaError_t ce = e != 0 ? (cudaError_t)e : cudaGetLastError();
            cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
            return cuda_fail(c, ce, "fp64 probe");
        }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double fl = 2.0 * 8.0 * (double)iters * (double)blocks * threads;
        if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    *tflops = best;
    return RL_OK;
}

}  // extern "C"
