// ssb_api.cu -- kernels and the C ABI of libssb (include/ssb.h).
//
// Launch geometry of the simulator kernels: one warp per environment, 4 environments per 128-thread CTA,
// grid = ceil(B/4).  At B = 4096 that is 1024 CTAs (~7 per SM on 148 SMs), all resident at once (the one-slot
// kernels are held to 72 registers for that); the kernels are dependency chains per environment bound by
// instruction supply, so residency is what keeps the SM busy.  The Decima policy is a sequence
// of list-driven tile kernels over all environments (ssb_decima_tc.cuh).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "ssb_decima.cuh"
#include "ssb_decima_tc.cuh"
#include "ssb_decima_fused.cuh"
#include "ssb_backward.cuh"
#include "ssb_sim.cuh"

using namespace ssb;

namespace ssb {
thread_local char g_cuda_err[256] = "";  // ssb_last_cuda_error(); also written by ssb_learn.cu
}

namespace {

#ifndef SSB_WARPS_PER_CTA
#define SSB_WARPS_PER_CTA 4
#endif
constexpr int WARPS_PER_CTA = SSB_WARPS_PER_CTA;


#define CUDA_TRY(expr)                                                                   \
    do {                                                                                 \
        cudaError_t e_ = (expr);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", #expr, cudaGetErrorString(e_)); \
            return SSB_E_CUDA;                                                           \
        }                                                                                \
    } while (0)

// The stream-taking entry points launch on the handle's device whatever the caller's current device is
// (a stream of another device fails the launch with a plain CUDA error instead of an opaque one later).
#define SSB_ON_DEVICE(env)                                                   \
    do {                                                                     \
        int cur_ = -1;                                                       \
        if (cudaGetDevice(&cur_) != cudaSuccess || cur_ != (env)->device) CUDA_TRY(cudaSetDevice((env)->device)); \
    } while (0)

// Ordering between the caller's streams and the handle's own stream (the *_host entry points): every stream-taking
// entry point marks the end of what it enqueued (SSB_MARK), and a *_host call first makes its own stream wait for
// that mark (host_begin) -- work still queued on the caller's stream is never overtaken by a host-buffer call.
#define SSB_MARK(env, stream)                                                             \
    do {                                                                                  \
        if ((cudaStream_t)(stream) != (env)->own_stream) {                                \
            CUDA_TRY(cudaEventRecord((env)->ev_last, (cudaStream_t)(stream)));            \
            (env)->dirty = 1;                                                             \
        }                                                                                 \
    } while (0)

// ------------------------------------------------------------------------------------ kernels
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
k_reset(Params p, const uint64_t *seeds, const double *time_limits, const uint8_t *mask)
{
    const int b = blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= p.B) return;
    if (mask && !mask[b]) return;
    Sim sim(p, b, lane);
    const uint64_t seed = seeds ? seeds[b] : 0ull;
    if (lane == 0) { sim.h->base_seed = seed; sim.h->reset_count = 1; }
    sim.reset_w(seed, time_limits ? time_limits[b] : (p.mean_time_limit > 0.0 ? sim.sample_time_limit(seed) : INFINITY));
}

// NS: executor slots per lane of the batched fast path (1: E <= 32, 2: E <= 64), see ssb_sim.cuh
template <int NS>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, NS == 1 ? 7 : 4)
k_step(Params p, const int32_t *stage_idx, const int32_t *num_exec, const uint8_t *mask, int max_events,
       int auto_reset, uint64_t seed_step, int32_t *next_a, int32_t *next_n, int dynamic_partition)
{
    const int b = blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= p.B) return;
    if (mask && !mask[b]) return;
    Sim sim(p, b, lane);
    // next_a / next_n (optional): the built-in fair / FIFO scheduler's action for the observation this call
    // leaves behind (ssb_step_fair_host), evaluated while the state is still in cache
    auto suggest = [&]() {
        if (!next_a) return;
        int a = -1, n = 1;
        if (!sim.h->error && !sim.h->done && !sim.h->pending) sim.fair_action_w(dynamic_partition != 0, a, n);
        if (lane == 0) { next_a[b] = a; next_n[b] = n; }
    };
    if (auto_reset && !sim.h->error && (sim.h->done || sim.oh->truncated)) {
        // the caller's `if terminated or truncated: env.reset(seed=...)` (rollout_worker.py:118-120, :150-153)
        const uint64_t seed = sim.h->base_seed + seed_step * (uint64_t)sim.h->reset_count;
        const double tl = sim.next_time_limit(seed);
        const bool was_trunc = !sim.h->done;
        __syncwarp();
        if (lane == 0) { sim.h->reset_count += 1; if (was_trunc) sim.stats->episodes++; }
        sim.reset_w(seed, tl);
        if (lane == 0) sim.oh->was_reset = 1;
        __syncwarp();
        suggest();
        return;
    }
    if (lane == 0) sim.oh->error = 0;
    __syncwarp();
    sim.template step_w<NS>(stage_idx[b], num_exec[b], max_events);
    __syncwarp();
    suggest();
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
k_fair_actions(Params p, int dynamic_partition, int32_t *stage_idx, int32_t *num_exec)
{
    const int b = blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= p.B) return;
    Sim sim(p, b, lane);
    int a = -1, n = 1;
    sim.fair_action_w(dynamic_partition != 0, a, n);
    if (lane == 0) { stage_idx[b] = a; num_exec[b] = n; }
}

// fused policy + step: `num_decisions` decisions per environment in one launch
// (min 7 CTAs per SM for the one-slot kernel: 4096 envs = 1024 CTAs must all be resident at once on
// 148 SMs; at 80 registers only 6 fit and the last 136 CTAs run as a second wave, +37 % time.
// Two-slot kernel: 4 CTAs per SM = 128 registers, what it needs without spilling.)
template <int NS>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, NS == 1 ? 7 : 4)
k_rollout_fair(Params p, int num_decisions, int dynamic_partition, int auto_reset, uint64_t seed_step,
               ssb_transition *traj)
{
    const int b = blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= p.B) return;
    Sim sim(p, b, lane);
#ifdef SSB_PROFILE
    const long long t_start = clock64();
#endif
    int d = 0, fresh = 0;
    while (d < num_decisions) {
        if (sim.h->error) break;
        if (sim.h->done || sim.oh->truncated) {
            if (!auto_reset) break;
            fresh = 4;
            // rollout_worker.py:118-120: seed = base_seed + seed_step * reset_count
            const uint64_t seed = sim.h->base_seed + seed_step * (uint64_t)sim.h->reset_count;
            const double tl = sim.next_time_limit(seed);
            const bool was_trunc = !sim.h->done;
            __syncwarp();
            if (lane == 0) { sim.h->reset_count += 1; if (was_trunc) sim.stats->episodes++; }
            sim.reset_w(seed, tl);
            continue;
        }
        int a = -1, n = 1;
#ifdef SSB_PROFILE
        long long tp0 = clock64();
#endif
        sim.fair_action_w(dynamic_partition != 0, a, n);
#ifdef SSB_PROFILE
        if (lane == 0) p.prof[(size_t)b * 16 + 8] += (unsigned long long)(clock64() - tp0);
#endif
        const double wall0 = sim.oh->wall_time;
        sim.template step_w<NS>(a, n);
        if (traj && lane == 0) {  // RolloutBuffer.add (rollout_worker.py:34-40) minus the observation
            ssb_transition t;
            t.wall_time = wall0; t.reward = sim.oh->reward; t.stage_idx = a; t.num_exec = n;
            t.flags = (sim.oh->terminated ? 1 : 0) | (sim.oh->truncated ? 2 : 0) | fresh;
            t.lgprob = 0.0f;
            traj[(size_t)b * num_decisions + d] = t;
        }
        fresh = 0;
        d++;
    }
#ifdef SSB_PROFILE
    if (lane == 0) p.prof[(size_t)b * 16 + 10] += (unsigned long long)(clock64() - t_start);
#endif
}

// RolloutWorkerAsync.collect_rollout (trainers/rollout_worker.py:160-206): the rollout ends when the env's accumulated
// simulated time reaches `duration` (or after max_decisions rows), resets do not end it, and the time axis of the
// stored rows is that accumulated time.  (A kernel of its own, so that the default rollout kernel stays as measured.)
template <int NS>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, NS == 1 ? 7 : 4)
k_rollout_fair_async(Params p, int max_decisions, double duration, int dynamic_partition, uint64_t seed_step,
                     ssb_transition *traj, int32_t *num_steps, double *elapsed_out)
{
    const int b = blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= p.B) return;
    Sim sim(p, b, lane);
    int d = 0, fresh = 0;
    double elapsed = 0.0;
    while (d < max_decisions && elapsed < duration) {
        if (sim.h->error) break;
        if (sim.h->done || sim.oh->truncated) {  // :196-200, applied when the loop comes back around
            fresh = 4;
            const uint64_t seed = sim.h->base_seed + seed_step * (uint64_t)sim.h->reset_count;
            const double tl = sim.next_time_limit(seed);
            const bool was_trunc = !sim.h->done;
            __syncwarp();
            if (lane == 0) { sim.h->reset_count += 1; if (was_trunc) sim.stats->episodes++; }
            sim.reset_w(seed, tl);
            continue;
        }
        int a = -1, n = 1;
        sim.fair_action_w(dynamic_partition != 0, a, n);
        const double wall0 = sim.oh->wall_time;
        sim.template step_w<NS>(a, n);
        if (traj && lane == 0) {  // rollout_buffer.add(obs, elapsed_time, action, lgprob, reward) (:191)
            ssb_transition t;
            t.wall_time = elapsed; t.reward = sim.oh->reward; t.stage_idx = a; t.num_exec = n;
            t.flags = (sim.oh->terminated ? 1 : 0) | (sim.oh->truncated ? 2 : 0) | fresh;
            t.lgprob = 0.0f;
            traj[(size_t)b * max_decisions + d] = t;
        }
        elapsed += sim.oh->wall_time - wall0;  // the duration of this step (:194)
        fresh = 0;
        d++;
    }
    if (lane == 0) {
        if (num_steps) num_steps[b] = d;
        if (elapsed_out) elapsed_out[b] = elapsed;
    }
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32) k_decima_obs(Params p)
{
    __shared__ uint64_t Sk[WARPS_PER_CTA][64];
    const int b = blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= p.B) return;
    Sim sim(p, b, lane);
    sim.decima_obs_w(Sk[threadIdx.x >> 5]);
}

// rollout-buffer rows around one { policy ; step } call of ssb_rollout_decima
// (the row index d lives in device memory so that one captured graph serves every decision of a call)
__global__ void k_traj_pre(Params p, const int32_t *a, const int32_t *n, ssb_transition *traj, int K)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.B) return;
    const int d = *p.traj_d;
    ssb_transition t;
    t.wall_time = p.obs_hdr[b].wall_time; t.reward = 0.0; t.stage_idx = a[b]; t.num_exec = n[b];
    t.flags = p.obs_hdr[b].was_reset ? 4 : 0;
    t.lgprob = p.pol_lgprob[b];
    traj[(size_t)b * K + d] = t;
}
__global__ void k_traj_post(Params p, ssb_transition *traj, int K)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.B) return;
    const int d = *p.traj_d;
    const ssb_obs_hdr &o = p.obs_hdr[b];
    ssb_transition &t = traj[(size_t)b * K + d];
    if (o.was_reset || o.error == SSB_ENV_DONE) { t.flags = 8; t.reward = 0.0; return; }
    t.reward = o.reward;
    t.flags |= (o.terminated ? 1 : 0) | (o.truncated ? 2 : 0);
}

// collect_stats sums in a fixed order: block k reduces the envs k, k + STATS_BLOCKS, ... (one warp per env, lanes
// over its jobs) into part[k][6]; a last warp adds the blocks' partial sums in block order.
constexpr int STATS_BLOCKS = 128;
__global__ void __launch_bounds__(256) k_collect_stats_part(Params p, double *part)
{
    __shared__ double sh[8][6];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (int b = (int)blockIdx.x + w * STATS_BLOCKS; b < p.B; b += 8 * STATS_BLOCKS) {
        const EnvHdr &h = p.hdr[b];
        const JobRec *jb = p.job + (size_t)b * p.Jc;
        const double wall = h.wall_time;
        double jt = 0.0, cd = 0.0;
        int nc = 0, na = 0;
        for (int j = lane; j < h.next_arrival; j += 32) {  // jobs that have arrived
            const JobRec &J = jb[j];
            jt += fmin(J.t_completed, wall) - J.t_arrival;
            if (J.state == JOB_COMPLETED) { nc++; cd += J.t_completed - J.t_arrival; }
            na++;
        }
        for (int off = 16; off; off >>= 1) {  // fixed butterfly order
            jt += __shfl_xor_sync(0xffffffffu, jt, off);
            cd += __shfl_xor_sync(0xffffffffu, cd, off);
            nc += __shfl_xor_sync(0xffffffffu, nc, off);
            na += __shfl_xor_sync(0xffffffffu, na, off);
        }
        if (wall > 0.0) { acc[0] += jt / wall; acc[1] += 1.0; }
        acc[2] += nc; acc[3] += na; acc[4] += cd; acc[5] += wall;
    }
    if (lane == 0)
        for (int i = 0; i < 6; i++) sh[w][i] = acc[i];
    __syncthreads();
    if (threadIdx.x < 6) {
        double s = 0.0;
        for (int i = 0; i < 8; i++) s += sh[i][threadIdx.x];
        part[(size_t)blockIdx.x * 8 + threadIdx.x] = s;
    }
}
__global__ void k_collect_stats_final(const double *part, double *out)
{
    if (threadIdx.x < 8) {
        double s = 0.0;
        if (threadIdx.x < 6)
            for (int k = 0; k < STATS_BLOCKS; k++) s += part[(size_t)k * 8 + threadIdx.x];
        out[threadIdx.x] = s;
    }
}

// ---- packed observations: the rows of every env's observation back to back (what a host caller copies out)
// offsets[b] = (first node, first edge, first job) of env b in the packed arrays, offsets[B] = the totals.
// One block; every thread scans a contiguous chunk of envs, the chunk sums are scanned through shared memory.
__global__ void __launch_bounds__(1024) k_pack_scan(Params p, int32_t *offsets)
{
    __shared__ int sh[3][1024];
    const int t = threadIdx.x, per = (p.B + 1023) / 1024, lo = min(t * per, p.B), hi = min(lo + per, p.B);
    int sn = 0, se = 0, sj = 0;
    for (int b = lo; b < hi; b++) {
        const ssb_obs_hdr &o = p.obs_hdr[b];
        sn += o.num_nodes; se += o.num_edges; sj += o.num_active_jobs;
    }
    sh[0][t] = sn; sh[1][t] = se; sh[2][t] = sj;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        int a = 0, b2 = 0, c = 0;
        if (t >= off) { a = sh[0][t - off]; b2 = sh[1][t - off]; c = sh[2][t - off]; }
        __syncthreads();
        sh[0][t] += a; sh[1][t] += b2; sh[2][t] += c;
        __syncthreads();
    }
    int on = sh[0][t] - sn, oe = sh[1][t] - se, oj = sh[2][t] - sj;
    for (int b = lo; b < hi; b++) {
        offsets[3 * b] = on; offsets[3 * b + 1] = oe; offsets[3 * b + 2] = oj;
        const ssb_obs_hdr &o = p.obs_hdr[b];
        on += o.num_nodes; oe += o.num_edges; oj += o.num_active_jobs;
    }
    if (t == 1023) { offsets[3 * p.B] = sh[0][t]; offsets[3 * p.B + 1] = sh[1][t]; offsets[3 * p.B + 2] = sh[2][t]; }
}
__global__ void __launch_bounds__(128)
k_pack_copy(Params p, const int32_t *offsets, float *nodes, int32_t *edges, int32_t *dag_ptr, int32_t *supplies)
{
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= p.B) return;
    const ssb_obs_hdr &o = p.obs_hdr[b];
    const int on = offsets[3 * b], oe = offsets[3 * b + 1], oj = offsets[3 * b + 2];
    const float *sn = p.obs_nodes + (size_t)b * p.Sc * 3;
    const int32_t *se = p.obs_edges + (size_t)b * p.Mc * 2;
    for (int i = lane; i < 3 * o.num_nodes; i += 32) nodes[(size_t)on * 3 + i] = sn[i];
    for (int i = lane; i < 2 * o.num_edges; i += 32) edges[(size_t)oe * 2 + i] = se[i];
    for (int i = lane; i <= o.num_active_jobs; i += 32) dag_ptr[oj + b + i] = p.obs_dag_ptr[(size_t)b * (p.Jc + 1) + i];
    for (int i = lane; i < o.num_active_jobs; i += 32) supplies[oj + i] = p.obs_supplies[(size_t)b * p.Jc + i];
}

__global__ void k_zero_stats(ssb_stats *s, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) s[i] = ssb_stats{};
}

// ------------------------------------------------------------------------------------ workspace
struct Carver {
    char *base;
    size_t off = 0;
    template <typename T>
    T *take(size_t n)
    {
        off = (off + 255) & ~size_t(255);
        T *ptr = base ? reinterpret_cast<T *>(base + off) : nullptr;
        off += n * sizeof(T);
        return ptr;
    }
};

struct Dims {
    int TAB, RT, Sc, Mc, P, Cc, max_stages, max_edges;
};

int compute_dims(const ssb_config &c, const ssb_bank &bk, Dims &d)
{
    if (c.num_envs < 1 || c.num_executors < 1 || c.num_executors > 128) return SSB_E_INVALID;
    if (c.max_jobs < 1 || c.max_jobs > 16000) return SSB_E_INVALID;
    if (c.job_arrival_cap > c.max_jobs) return SSB_E_INVALID;
    if (!(c.job_arrival_rate > 0)) return SSB_E_INVALID;
    if (bk.num_templates != 154) return SSB_E_INVALID;  // 7 sizes x 22 queries (tpch.py:14-15)
    int ms = 0, me = 0;
    for (int t = 0; t < bk.num_templates; t++) {
        ms = bk.num_stages[t] > ms ? bk.num_stages[t] : ms;
        int ne = bk.edge_base[t + 1] - bk.edge_base[t];
        me = ne > me ? ne : me;
    }
    if (ms < 1 || ms > 64) return SSB_E_INVALID;
    d.max_stages = ms;
    d.max_edges = me;
    d.TAB = 8;
    while (d.TAB <= 4 * c.num_executors) d.TAB <<= 1;
    d.RT = 8;
    while (d.RT <= 4 * c.max_jobs) d.RT <<= 1;
    d.Sc = c.max_jobs * ms;
    d.Mc = c.max_jobs * me;
    if (d.Sc > 32000) return SSB_E_INVALID;
    d.P = 2 + c.max_jobs + d.Sc;
    d.Cc = 2 * c.num_executors + 16;
    return SSB_OK;
}

struct BankDev {
    int32_t *num_stages, *stage_base, *edge_base, *num_tasks;
    int16_t *edges;
    double *rough;
    uint64_t *parent, *child;
    uint8_t *present;
    uint2 *dur;
    double *vals;
    short4 *iv;
};

void carve(Carver &cv, const ssb_config &c, const ssb_bank &bk, const Dims &d, Params &p, BankDev &bd,
           int32_t **st_a, int32_t **st_n, uint64_t **st_seed, double **st_tl, uint8_t **st_mask)
{
    const size_t B = c.num_envs, T = bk.num_templates, TS = bk.num_template_stages;
    bd.num_stages = cv.take<int32_t>(T);
    bd.stage_base = cv.take<int32_t>(T + 1);
    bd.edge_base = cv.take<int32_t>(T + 1);
    bd.num_tasks = cv.take<int32_t>(TS);
    bd.edges = cv.take<int16_t>(2 * (size_t)bk.num_template_edges);
    bd.rough = cv.take<double>(TS);
    bd.parent = cv.take<uint64_t>(TS);
    bd.child = cv.take<uint64_t>(TS);
    bd.present = cv.take<uint8_t>(TS * 4);
    bd.dur = cv.take<uint2>(TS * 24);
    bd.vals = cv.take<double>((size_t)bk.num_values + 1);
    bd.iv = cv.take<short4>((size_t)c.num_executors + 1);
    p.hdr = cv.take<EnvHdr>(B);
    p.exec = cv.take<ExecRec>(B * c.num_executors);
    p.job = cv.take<JobRec>(B * c.max_jobs);
    p.stage = cv.take<StageRec>(B * d.Sc);
    p.active = cv.take<int16_t>(B * c.max_jobs);
    p.old_act = cv.take<int16_t>(B * c.max_jobs);
    p.reward_ord = cv.take<int16_t>(B * c.max_jobs);
    p.commits = cv.take<Commit>(B * d.Cc);
    p.pool_hdr = cv.take<PoolHdr>(B * d.P);
    p.pool_tab = cv.take<uint8_t>(B * d.P * d.TAB);
    p.scr_tab = cv.take<uint8_t>(B * 3 * d.TAB);
    p.rset = cv.take<uint16_t>(B * 2 * d.RT);
    p.trace_t = cv.take<double>(c.tape_capacity > 0 ? B * c.max_jobs : 1);
    p.trace_tmpl = cv.take<int32_t>(c.tape_capacity > 0 ? B * c.max_jobs : 1);
    p.tape = cv.take<double>(c.tape_capacity > 0 ? B * (size_t)c.tape_capacity : 1);
    p.log = cv.take<LogRow>(c.log_capacity > 0 ? B * (size_t)c.log_capacity : 1);
    p.hist = cv.take<HistRow>(c.history_capacity > 0 ? B * (size_t)c.history_capacity : 1);
    p.stats = cv.take<ssb_stats>(B);
    p.stats_part = cv.take<double>(128 * 8);
    p.prof = cv.take<unsigned long long>(B * 16);
    p.obs_hdr = cv.take<ssb_obs_hdr>(B);
    p.obs_nodes = cv.take<float>(B * d.Sc * 3);
    p.obs_edges = cv.take<int32_t>(B * d.Mc * 2);
    p.obs_dag_ptr = cv.take<int32_t>(B * (c.max_jobs + 1));
    p.obs_supplies = cv.take<int32_t>(B * c.max_jobs);
    if (c.flags & (SSB_FLAG_DECIMA_OBS | SSB_FLAG_DECIMA_POLICY)) {
        p.dec_feat = cv.take<float>(B * d.Sc * 5);
        p.dec_stage_mask = cv.take<uint8_t>(B * d.Sc);
        p.dec_frontier_mask = cv.take<uint8_t>(B * d.Sc);
        p.dec_caps = cv.take<int32_t>(B * c.max_jobs);
        p.dec_edge_bits = cv.take<uint64_t>(B * d.Mc);
        p.dec_depth = cv.take<int32_t>(B);
    }
    if (c.flags & SSB_FLAG_DECIMA_POLICY) {
        p.Epad = (c.num_executors + 3) & ~3;
        p.pol_w = cv.take<float>(dd::TOTAL);
        p.pol_h_init = cv.take<float>(B * d.Sc * 16);
        p.pol_h = cv.take<float>(B * d.Sc * 16);
        p.pol_msg = cv.take<float>(B * d.Sc * 16);
        p.pol_h_dag = cv.take<float>(B * c.max_jobs * 16);
        p.pol_g = cv.take<float>(B * c.max_jobs * 16);
        p.pol_h_glob = cv.take<float>(B * 16);
        p.pol_row_start = cv.take<int32_t>(B * d.Sc);
        p.pol_stage_logits = cv.take<float>(B * d.Sc);
        p.pol_exec_logits = cv.take<float>(B * p.Epad);
        p.pol_action = cv.take<int32_t>(B * 4);
        p.pol_lgprob = cv.take<float>(B);
        p.pol_entropy = cv.take<float>(B);
        {   // scratch that parks the live observation during ssb_decima_evaluate (same layout as a snapshot)
            size_t sb = 0;
            const size_t parts[8] = {B * sizeof(ssb_obs_hdr), B * d.Mc * 2 * sizeof(int32_t),
                                     B * (c.max_jobs + 1) * sizeof(int32_t), B * d.Sc * 5 * sizeof(float), B * d.Sc,
                                     B * c.max_jobs * sizeof(int32_t), B * d.Mc * sizeof(uint64_t), B * sizeof(int32_t)};
            for (size_t x : parts) sb += (x + 255) & ~size_t(255);
            p.pol_snap = cv.take<char>(sb);
        }
        p.traj_d = cv.take<int32_t>(4);
        p.pol_act_a = cv.take<int32_t>(B);
        p.pol_act_n = cv.take<int32_t>(B);
        p.pl_all = cv.take<int32_t>(B * d.Sc);
        p.pl_sink = cv.take<int32_t>(B * d.Sc);
        p.pl_cand = cv.take<int32_t>(B * d.Sc);
        p.pl_cand_job = cv.take<int32_t>(B * d.Sc);
        p.pl_cand_out = cv.take<int32_t>(B * d.Sc);
        p.pl_jobs = cv.take<int32_t>(B * c.max_jobs);
        p.pl_exec = cv.take<int32_t>(B * p.Epad);
        // every masked edge contributes at most one sender and one receiver entry per level it is masked at
        p.lvl_cap = (int)std::min<size_t>(4 * B * d.Mc, (size_t)0x7fffffff);
        p.pl_lvl = cv.take<int32_t>((size_t)p.lvl_cap);
        p.pl_cnt = cv.take<int32_t>(tc::CNT_TOTAL);
        p.pl_ncand = cv.take<int32_t>(B);
        p.pl_bits = cv.take<unsigned long long>(B * d.Sc * 2);
        p.pol_wblob = cv.take<float>(tc::BLOB_TOTAL);
        p.pol_wblob3 = cv.take<uint32_t>(fz::BLOB_TOTAL);
        p.pol_cand_rank = cv.take<int32_t>(B * d.Sc);
        p.fz_cursor = cv.take<int32_t>(4);
        p.as_elapsed = cv.take<double>(B);
        p.as_wall0 = cv.take<double>(B);
        p.as_rows = cv.take<int32_t>(B);
        p.as_kind = cv.take<uint8_t>(B);
        p.as_fresh = cv.take<uint8_t>(B);
        p.as_any = cv.take<int32_t>(4);
    }
    *st_a = cv.take<int32_t>(B);
    *st_n = cv.take<int32_t>(B);
    *st_seed = cv.take<uint64_t>(B);
    *st_tl = cv.take<double>(B);
    *st_mask = cv.take<uint8_t>(B);
}

}  // namespace

// How ssb_decima_policy runs: row lists + one launch per MLP pass with the TMEM-resident bf16 three-term tiles (the
// default for large batches), the whole decision of a group of envs in one persistent kernel (one launch: small
// batches, e.g. the single-env facade), or round 1's shared-memory tf32 tiles (kept for A/B measurements).
enum { POLICY_TILES = 0, POLICY_FUSED = 1, POLICY_TILES_TF32 = 2 };

struct ssb_env {
    ssb_config cfg;
    Dims dims;
    Params p;
    BankDev bank;
    int device;
    char *ws;
    size_t ws_bytes;
    cudaStream_t own_stream;
    int32_t *st_a, *st_n;  // staging for the *_host entry points
    uint64_t *st_seed;
    double *st_tl;
    uint8_t *st_mask;
    int grid;
    int num_sms;
    int dmax;           // upper bound of the message-passing depth: longest template chain - 1
    // CUDA graph of one ssb_rollout_decima decision and the arguments it was captured with
    cudaEvent_t ev;
    cudaEvent_t ev_last;  // recorded after the latest work enqueued on a caller's stream (see SSB_MARK / host_begin)
    int dirty;            // such work exists since the last *_host call
    cudaGraphExec_t dg_exec;
    ssb_transition *dg_traj;
    int dg_k, dg_events, dg_autoreset, no_graph;
    uint64_t dg_seed_step;
    int policy_mode;    // POLICY_* below (SSB_DECIMA_MODE overrides the default)
    int fused_group;    // environments per group of the fused policy kernel
    int snap_loaded;    // ssb_decima_snapshot_load: a stored observation is in place, the live one parked
    int auto_reset;     // ssb_set_autoreset
    uint64_t auto_seed_step;
};

static int host_begin(ssb_env *env)
{
    CUDA_TRY(cudaSetDevice(env->device));
    if (env->dirty) {
        CUDA_TRY(cudaStreamWaitEvent(env->own_stream, env->ev_last, 0));
        env->dirty = 0;
    }
    return SSB_OK;
}

template <int ST>
static int launch_mlp_rows(ssb_env *env, const float *x, int n, float *out, cudaStream_t s)
{
    const size_t smem = sizeof(uint32_t) * fz::Blob<ST>::WORDS;
    CUDA_TRY(cudaFuncSetAttribute(fz::k_mlp_rows<ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles = (n + 127) / 128;
    fz::k_mlp_rows<ST><<<std::min(tiles, env->num_sms * 4), 128, smem, s>>>(env->p.pol_wblob3, x, n, out);
    CUDA_TRY(cudaGetLastError());
    SSB_MARK(env, s);
    return SSB_OK;
}

struct BackwardScratch { float *gs, *ge, *d_hdag, *d_hglob, *d_hinit, *d_msg; size_t floats; };
BackwardScratch backward_scratch(const Params &p, float *base)
{
    BackwardScratch b{};
    size_t off = 0;
    auto take = [&](size_t n) { float *q = base ? base + off : nullptr; off += (n + 63) & ~size_t(63); return q; };
    b.gs = take((size_t)p.B * p.Sc);
    b.ge = take((size_t)p.B * p.Epad);
    b.d_hdag = take((size_t)p.B * p.Jc * 16);
    b.d_hglob = take((size_t)p.B * 16);
    b.d_hinit = take((size_t)p.B * p.Sc * 16);
    b.d_msg = take((size_t)p.B * p.Sc * 16);
    b.floats = off;
    return b;
}
int launch_mlp_backward(int stage, ssb_env *env, const int32_t *list, const int32_t *offset, const int32_t *count,
                        int level, const float *g_out, float *dW, const bwd::Bufs &bw, cudaStream_t s);
template <int ST>
int launch_tile(ssb_env *env, const int32_t *list, const int32_t *offset, const int32_t *count, int level,
                int ctas_per_sm, cudaStream_t s)
{
    tc::TileArgs a{list, offset, count, level};
    if (env->policy_mode == POLICY_TILES_TF32)
        tc::k_tile_mlp<ST><<<env->num_sms * ctas_per_sm, 128, tc::Smem<ST>::BYTES, s>>>(env->p, a);
    else
        fz::k_tile3<ST><<<env->num_sms * (fz::Spec<ST>::OUT > 1 ? 2 * ctas_per_sm : ctas_per_sm), 128,
                          sizeof(uint32_t) * fz::Blob<ST>::WORDS, s>>>(env->p, a);
    CUDA_TRY(cudaGetLastError());
    return SSB_OK;
}
template <int ST>
int prepare_tile_kernel()
{
    CUDA_TRY(cudaFuncSetAttribute(tc::k_tile_mlp<ST>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)tc::Smem<ST>::BYTES));
    CUDA_TRY(cudaFuncSetAttribute(fz::k_tile3<ST>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(sizeof(uint32_t) * fz::Blob<ST>::WORDS)));
    return SSB_OK;
}

int launch_mlp_backward(int stage, ssb_env *env, const int32_t *list, const int32_t *offset, const int32_t *count,
                        int level, const float *g_out, float *dW, const bwd::Bufs &bw, cudaStream_t s)
{
    CUDA_TRY(bwd::mlp_backward(stage, env->p, env->num_sms, list, offset, count, level, g_out, nullptr, nullptr, dW, bw,
                               true, s));
    return SSB_OK;
}

extern "C" {

int ssb_abi_version(void) { return SSB_ABI_VERSION; }
const char *ssb_last_cuda_error(void) { return g_cuda_err; }

int ssb_workspace_bytes(const ssb_config *cfg, const ssb_bank *bank, size_t *bytes)
{
    if (!cfg || !bank || !bytes) return SSB_E_INVALID;
    Dims d;
    int rc = compute_dims(*cfg, *bank, d);
    if (rc) return rc;
    Carver cv{nullptr};
    Params p{};
    BankDev bd{};
    int32_t *a, *n; uint64_t *s; double *tl; uint8_t *m;
    carve(cv, *cfg, *bank, d, p, bd, &a, &n, &s, &tl, &m);
    *bytes = (cv.off + 255) & ~size_t(255);
    return SSB_OK;
}

int ssb_create(const ssb_config *cfg, const ssb_bank *bk, int device, void *workspace,
               size_t workspace_bytes, ssb_env **out)
{
    if (!cfg || !bk || !workspace || !out) return SSB_E_INVALID;
    if (reinterpret_cast<uintptr_t>(workspace) & 255) return SSB_E_WORKSPACE;
    Dims d;
    int rc = compute_dims(*cfg, *bk, d);
    if (rc) return rc;
    size_t need = 0;
    ssb_workspace_bytes(cfg, bk, &need);
    if (workspace_bytes < need) return SSB_E_WORKSPACE;
    CUDA_TRY(cudaSetDevice(device));
    ssb_env *env = new (std::nothrow) ssb_env();
    if (!env) return SSB_E_INVALID;
    env->cfg = *cfg;
    env->dims = d;
    env->device = device;
    env->ws = static_cast<char *>(workspace);
    env->ws_bytes = workspace_bytes;
    Carver cv{env->ws};
    Params &p = env->p;
    p = Params{};
    carve(cv, *cfg, *bk, d, p, env->bank, &env->st_a, &env->st_n, &env->st_seed, &env->st_tl, &env->st_mask);
    p.B = cfg->num_envs; p.E = cfg->num_executors; p.Jc = cfg->max_jobs; p.Sc = d.Sc; p.Mc = d.Mc;
    p.TAB = d.TAB; p.RT = d.RT; p.P = d.P; p.Cc = d.Cc; p.max_stages = d.max_stages;
    p.tape_cap = cfg->tape_capacity; p.log_cap = cfg->log_capacity;
    p.hist_cap = cfg->history_capacity > 0 ? cfg->history_capacity : 0;
    p.job_arrival_cap = cfg->job_arrival_cap;
    p.moving_delay = cfg->moving_delay; p.warmup_delay = cfg->warmup_delay;
    p.mean_interarrival = 1 / cfg->job_arrival_rate;  // tpch.py:42
    p.beta = cfg->beta;
    env->grid = (cfg->num_envs + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    {
        const char *ng = getenv("SSB_NO_GRAPH");
        env->no_graph = ng && ng[0] == '1';
    }
    CUDA_TRY(cudaDeviceGetAttribute(&env->num_sms, cudaDevAttrMultiProcessorCount, device));
    {
        // longest chain of topological generations over the templates bounds the depth of every observation
        int dmax = 0;
        for (int t = 0; t < bk->num_templates; t++) {
            const int ns = bk->num_stages[t], base = bk->stage_base[t];
            const uint64_t all = ns >= 64 ? ~0ull : ((1ull << ns) - 1);
            uint64_t done = 0;
            int gens = 0;
            while (done != all && gens < 64) {
                uint64_t lk = 0;
                for (int s = 0; s < ns; s++)
                    if (!((done >> s) & 1) && (bk->parent_mask[base + s] & ~done) == 0) lk |= 1ull << s;
                if (!lk) break;
                done |= lk;
                gens++;
            }
            dmax = gens - 1 > dmax ? gens - 1 : dmax;
        }
        env->dmax = dmax;
    }
    if (p.pol_w) {
        env->policy_mode = cfg->num_envs <= 2 * env->num_sms ? POLICY_FUSED : POLICY_TILES;
        if (const char *pm = getenv("SSB_DECIMA_MODE")) env->policy_mode = std::max(0, std::min(atoi(pm), 2));
        CUDA_TRY(cudaFuncSetAttribute(fz::fused::k_decima_fused, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)fz::fused::Smem::BYTES));
        {   // group size: one round of groups over the resident CTAs (two per SM) when that needs at most GMAX
            // environments per group; otherwise ~3000 observation nodes per group (full 128-row tiles within a
            // level), the group count rounded to whole rounds
            const int ctas = 2 * env->num_sms, B = cfg->num_envs;
            int G = (B + ctas - 1) / ctas;
            if (G > fz::fused::GMAX) {
                const int by_nodes = std::max(1, std::min(fz::fused::GMAX, 3072 / std::max(64, d.Sc / 4)));
                const int rounds = (B + by_nodes * ctas - 1) / (by_nodes * ctas);
                G = (B + rounds * ctas - 1) / (rounds * ctas);
            }
            G = std::max(1, std::min(G, fz::fused::GMAX));
            if (const char *fg = getenv("SSB_FUSED_GROUP")) G = std::max(1, std::min(atoi(fg), fz::fused::GMAX));
            env->fused_group = G;
        }
        int rc;
        if ((rc = prepare_tile_kernel<tc::ST_PREP>()) || (rc = prepare_tile_kernel<tc::ST_SINK>()) ||
            (rc = prepare_tile_kernel<tc::ST_MSG>()) || (rc = prepare_tile_kernel<tc::ST_RCV>()) ||
            (rc = prepare_tile_kernel<tc::ST_DAG>()) || (rc = prepare_tile_kernel<tc::ST_GLOB>()) ||
            (rc = prepare_tile_kernel<tc::ST_STAGE>()) || (rc = prepare_tile_kernel<tc::ST_EXEC>()))
            return rc;

    }
    // ---- bank upload
    const size_t T = bk->num_templates, TS = bk->num_template_stages, ME = bk->num_template_edges;
    BankDev &bd = env->bank;
    CUDA_TRY(cudaMemcpy(bd.num_stages, bk->num_stages, T * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(bd.stage_base, bk->stage_base, (T + 1) * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(bd.edge_base, bk->edge_base, (T + 1) * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(bd.num_tasks, bk->num_tasks, TS * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(bd.rough, bk->rough_duration, TS * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(bd.parent, bk->parent_mask, TS * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(bd.child, bk->child_mask, TS * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(bd.vals, bk->dur_values, (size_t)bk->num_values * 8, cudaMemcpyHostToDevice));
    {
        std::vector<int16_t> e16(2 * ME);
        for (size_t i = 0; i < 2 * ME; i++) e16[i] = (int16_t)bk->edges[i];
        CUDA_TRY(cudaMemcpy(bd.edges, e16.data(), e16.size() * 2, cudaMemcpyHostToDevice));
        std::vector<uint8_t> pr(TS * 4, 0);
        for (size_t i = 0; i < TS; i++)
            for (int w = 0; w < 3; w++) pr[i * 4 + w] = bk->present[i * 3 + w];
        CUDA_TRY(cudaMemcpy(bd.present, pr.data(), pr.size(), cudaMemcpyHostToDevice));
        std::vector<uint2> du(TS * 24);
        for (size_t i = 0; i < TS * 24; i++) du[i] = make_uint2(bk->dur_off[i], bk->dur_cnt[i]);
        CUDA_TRY(cudaMemcpy(bd.dur, du.data(), du.size() * sizeof(uint2), cudaMemcpyHostToDevice));
        for (size_t i = 0; i < TS; i++)
            if (bk->num_tasks[i] < 1 || bk->num_tasks[i] > 65535) { delete env; return SSB_E_INVALID; }
        // _init_executor_intervals (tpch.py:237-262) as integer levels
        const int LV[8] = {5, 10, 20, 40, 50, 60, 80, 100};
        const int cap = cfg->num_executors;
        std::vector<int16_t> iv(2 * (cap + 1), 0);
        for (int i = 0; i <= LV[0] && i <= cap; i++) iv[2 * i] = iv[2 * i + 1] = LV[0];
        for (int i = 0; i < 7; i++) {
            for (int r = LV[i] + 1; r < LV[i + 1] && r <= cap; r++) { iv[2 * r] = LV[i]; iv[2 * r + 1] = LV[i + 1]; }
            if (LV[i + 1] > cap) break;
            iv[2 * LV[i + 1]] = iv[2 * LV[i + 1] + 1] = LV[i + 1];
        }
        if (cap > LV[7]) for (int r = LV[7] + 1; r < cap; r++) iv[2 * r] = iv[2 * r + 1] = LV[7];
        std::vector<short4> iv4(cap + 1);
        for (int r = 0; r <= cap; r++) {
            int li = -1, ri = -1;  // level value -> level index (-1: not a data level, e.g. the 0 rows)
            for (int q = 0; q < 8; q++) {
                if (LV[q] == iv[2 * r]) li = q;
                if (LV[q] == iv[2 * r + 1]) ri = q;
            }
            iv4[r] = make_short4(iv[2 * r], iv[2 * r + 1], (short)li, (short)ri);
        }
        CUDA_TRY(cudaMemcpy(bd.iv, iv4.data(), iv4.size() * sizeof(short4), cudaMemcpyHostToDevice));
    }
    p.b_num_stages = bd.num_stages; p.b_stage_base = bd.stage_base; p.b_edge_base = bd.edge_base;
    p.b_num_tasks = bd.num_tasks; p.b_edges = bd.edges; p.b_rough = bd.rough; p.b_parent = bd.parent;
    p.b_child = bd.child; p.b_present = bd.present; p.b_dur = bd.dur; p.b_vals = bd.vals; p.iv = bd.iv;
    CUDA_TRY(cudaMemset(p.hdr, 0, sizeof(EnvHdr) * (size_t)p.B));
    CUDA_TRY(cudaMemset(p.stats, 0, sizeof(ssb_stats) * (size_t)p.B));
    CUDA_TRY(cudaMemset(p.prof, 0, sizeof(unsigned long long) * 16 * (size_t)p.B));
    CUDA_TRY(cudaMemset(p.obs_hdr, 0, sizeof(ssb_obs_hdr) * (size_t)p.B));
    CUDA_TRY(cudaStreamCreateWithFlags(&env->own_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&env->ev, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&env->ev_last, cudaEventDisableTiming));
    CUDA_TRY(cudaDeviceSynchronize());
    *out = env;
    return SSB_OK;
}

int ssb_destroy(ssb_env *env)
{
    if (!env) return SSB_E_INVALID;
    cudaSetDevice(env->device);
    cudaStreamSynchronize(env->own_stream);
    cudaStreamDestroy(env->own_stream);
    if (env->dg_exec) cudaGraphExecDestroy(env->dg_exec);
    if (env->ev) cudaEventDestroy(env->ev);
    if (env->ev_last) cudaEventDestroy(env->ev_last);
    delete env;
    return SSB_OK;
}

int ssb_load_trace(ssb_env *env, int32_t b, int32_t n_jobs, const double *t_arrival, const int32_t *tmpl,
                   const double *tape, int64_t n_tape)
{
    if (!env || b < 0 || b >= env->p.B || n_jobs < 1 || n_jobs > env->p.Jc || !t_arrival || !tmpl)
        return SSB_E_INVALID;
    if (env->cfg.tape_capacity <= 0) return SSB_E_INVALID;
    if (tape && n_tape > env->cfg.tape_capacity) return SSB_E_INVALID;
    if (t_arrival[0] != 0.0) return SSB_E_INVALID;  // spark_sched_sim.py:150
    for (int j = 0; j < n_jobs; j++)
        if (tmpl[j] < 0 || tmpl[j] >= 154) return SSB_E_INVALID;
    CUDA_TRY(cudaSetDevice(env->device));
    const Params &p = env->p;
    CUDA_TRY(cudaMemcpy(p.trace_t + (size_t)b * p.Jc, t_arrival, n_jobs * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(p.trace_tmpl + (size_t)b * p.Jc, tmpl, n_jobs * 4, cudaMemcpyHostToDevice));
    int32_t tl = tape ? (int32_t)n_tape : -1;
    if (tape && n_tape > 0)
        CUDA_TRY(cudaMemcpy(p.tape + (size_t)b * p.tape_cap, tape, n_tape * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(&p.hdr[b].trace_jobs, &n_jobs, 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(&p.hdr[b].tape_len, &tl, 4, cudaMemcpyHostToDevice));
    return SSB_OK;
}

int ssb_clear_trace(ssb_env *env, int32_t b)
{
    if (!env || b < 0 || b >= env->p.B) return SSB_E_INVALID;
    CUDA_TRY(cudaSetDevice(env->device));
    int32_t z = 0;
    CUDA_TRY(cudaMemcpy(&env->p.hdr[b].trace_jobs, &z, 4, cudaMemcpyHostToDevice));
    return SSB_OK;
}

int ssb_reset(ssb_env *env, const uint64_t *seeds, const double *time_limits, const uint8_t *mask, void *stream)
{
    if (!env) return SSB_E_INVALID;
    SSB_ON_DEVICE(env);
    k_reset<<<env->grid, WARPS_PER_CTA * 32, 0, (cudaStream_t)stream>>>(env->p, seeds, time_limits, mask);
    CUDA_TRY(cudaGetLastError());
    SSB_MARK(env, stream);
    return SSB_OK;
}

static int step_launch(ssb_env *env, const int32_t *stage_idx, const int32_t *num_exec, const uint8_t *mask,
                       int32_t max_events, int32_t *next_a, int32_t *next_n, int dyn, cudaStream_t s,
                       int force_autoreset = 0, uint64_t force_seed_step = 0)
{
    const int ar = force_autoreset ? 1 : env->auto_reset;
    const uint64_t ss = force_autoreset ? force_seed_step : env->auto_seed_step;
    if (env->p.E <= 32)
        k_step<1><<<env->grid, WARPS_PER_CTA * 32, 0, s>>>(env->p, stage_idx, num_exec, mask, max_events, ar, ss, next_a,
                                                           next_n, dyn);
    else
        k_step<2><<<env->grid, WARPS_PER_CTA * 32, 0, s>>>(env->p, stage_idx, num_exec, mask, max_events, ar, ss, next_a,
                                                           next_n, dyn);
    CUDA_TRY(cudaGetLastError());
    SSB_MARK(env, s);
    return SSB_OK;
}

int ssb_step(ssb_env *env, const int32_t *stage_idx, const int32_t *num_exec, const uint8_t *mask,
             int32_t max_events, void *stream)
{
    if (!env || !stage_idx || !num_exec) return SSB_E_INVALID;
    SSB_ON_DEVICE(env);
    return step_launch(env, stage_idx, num_exec, mask, max_events, nullptr, nullptr, 0, (cudaStream_t)stream);
}

// The captured graph of ssb_rollout_decima holds Params and the auto-reset arguments BY VALUE: every setter that
// changes one of them drops the graph, the next rollout call captures it again.
static void drop_decision_graph(ssb_env *env)
{
    if (env->dg_exec) { cudaGraphExecDestroy(env->dg_exec); env->dg_exec = nullptr; }
}

int ssb_set_autoreset(ssb_env *env, int32_t enable, uint64_t seed_step)
{
    if (!env) return SSB_E_INVALID;
    env->auto_reset = enable ? 1 : 0;
    env->auto_seed_step = seed_step;
    drop_decision_graph(env);
    return SSB_OK;
}

int ssb_set_mean_time_limit(ssb_env *env, double mean_ms)
{
    if (!env || !(mean_ms >= 0.0)) return SSB_E_INVALID;
    env->p.mean_time_limit = mean_ms;
    drop_decision_graph(env);
    return SSB_OK;
}

int ssb_reset_host(ssb_env *env, const uint64_t *seeds, const double *time_limits, const uint8_t *mask,
                   ssb_obs_hdr *hdr_out)
{
    if (!env) return SSB_E_INVALID;
    {
        const int rcb = host_begin(env);  // also waits for work still queued on the caller's streams
        if (rcb) return rcb;
    }
    cudaStream_t s = env->own_stream;
    const size_t B = env->p.B;
    if (seeds) CUDA_TRY(cudaMemcpyAsync(env->st_seed, seeds, B * 8, cudaMemcpyHostToDevice, s));
    if (time_limits) CUDA_TRY(cudaMemcpyAsync(env->st_tl, time_limits, B * 8, cudaMemcpyHostToDevice, s));
    if (mask) CUDA_TRY(cudaMemcpyAsync(env->st_mask, mask, B, cudaMemcpyHostToDevice, s));
    int rc = ssb_reset(env, seeds ? env->st_seed : nullptr, time_limits ? env->st_tl : nullptr,
                       mask ? env->st_mask : nullptr, s);
    if (rc) return rc;
    if (hdr_out) CUDA_TRY(cudaMemcpyAsync(hdr_out, env->p.obs_hdr, B * sizeof(ssb_obs_hdr), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return SSB_OK;
}

int ssb_step_host(ssb_env *env, const int32_t *stage_idx, const int32_t *num_exec, const uint8_t *mask,
                  int32_t max_events, ssb_obs_hdr *hdr_out)
{
    if (!env || !stage_idx || !num_exec) return SSB_E_INVALID;
    {
        const int rcb = host_begin(env);  // also waits for work still queued on the caller's streams
        if (rcb) return rcb;
    }
    cudaStream_t s = env->own_stream;
    const size_t B = env->p.B;
    CUDA_TRY(cudaMemcpyAsync(env->st_a, stage_idx, B * 4, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(env->st_n, num_exec, B * 4, cudaMemcpyHostToDevice, s));
    if (mask) CUDA_TRY(cudaMemcpyAsync(env->st_mask, mask, B, cudaMemcpyHostToDevice, s));
    int rc = ssb_step(env, env->st_a, env->st_n, mask ? env->st_mask : nullptr, max_events, s);
    if (rc) return rc;
    if (hdr_out) CUDA_TRY(cudaMemcpyAsync(hdr_out, env->p.obs_hdr, B * sizeof(ssb_obs_hdr), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return SSB_OK;
}

int ssb_step_fair_host(ssb_env *env, const int32_t *stage_idx, const int32_t *num_exec, const uint8_t *mask,
                       int32_t max_events, int32_t dynamic_partition, ssb_obs_hdr *hdr_out, int32_t *next_stage_idx,
                       int32_t *next_num_exec)
{
    if (!env || !stage_idx || !num_exec || !next_stage_idx || !next_num_exec) return SSB_E_INVALID;
    {
        const int rcb = host_begin(env);  // also waits for work still queued on the caller's streams
        if (rcb) return rcb;
    }
    cudaStream_t s = env->own_stream;
    const size_t B = env->p.B;
    CUDA_TRY(cudaMemcpyAsync(env->st_a, stage_idx, B * 4, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(env->st_n, num_exec, B * 4, cudaMemcpyHostToDevice, s));
    if (mask) CUDA_TRY(cudaMemcpyAsync(env->st_mask, mask, B, cudaMemcpyHostToDevice, s));
    // the suggestions are written to the staging arrays the actions were read from (each env reads its action
    // before it writes its suggestion)
    int rc = step_launch(env, env->st_a, env->st_n, mask ? env->st_mask : nullptr, max_events, env->st_a, env->st_n,
                         dynamic_partition, s);
    if (rc) return rc;
    if (hdr_out) CUDA_TRY(cudaMemcpyAsync(hdr_out, env->p.obs_hdr, B * sizeof(ssb_obs_hdr), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(next_stage_idx, env->st_a, B * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(next_num_exec, env->st_n, B * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return SSB_OK;
}

// layout of the caller's scratch for the packed observation
namespace {
struct PackLayout { size_t off_offsets, off_nodes, off_edges, off_dag, off_sup, bytes; };
PackLayout pack_layout(const Params &p)
{
    PackLayout l{};
    size_t o = 0;
    auto take = [&](size_t n) { size_t at = o; o += (n + 255) & ~size_t(255); return at; };
    l.off_offsets = take(sizeof(int32_t) * 3 * ((size_t)p.B + 1));
    l.off_nodes = take(sizeof(float) * 3 * (size_t)p.B * p.Sc);
    l.off_edges = take(sizeof(int32_t) * 2 * (size_t)p.B * p.Mc);
    l.off_dag = take(sizeof(int32_t) * (size_t)p.B * (p.Jc + 1));
    l.off_sup = take(sizeof(int32_t) * (size_t)p.B * p.Jc);
    l.bytes = o;
    return l;
}
}  // namespace

int ssb_packed_obs_bytes(ssb_env *env, size_t *bytes)
{
    if (!env || !bytes) return SSB_E_INVALID;
    *bytes = pack_layout(env->p).bytes;
    return SSB_OK;
}

int ssb_get_obs_host(ssb_env *env, ssb_packed_obs *out, void *scratch, size_t scratch_bytes)
{
    if (!env || !out || !out->offsets || !scratch || (reinterpret_cast<uintptr_t>(scratch) & 255)) return SSB_E_INVALID;
    const Params &p = env->p;
    const PackLayout l = pack_layout(p);
    if (scratch_bytes < l.bytes) return SSB_E_WORKSPACE;
    {
        const int rcb = host_begin(env);  // also waits for work still queued on the caller's streams
        if (rcb) return rcb;
    }
    cudaStream_t s = env->own_stream;
    char *sc = static_cast<char *>(scratch);
    int32_t *d_off = reinterpret_cast<int32_t *>(sc + l.off_offsets);
    float *d_nodes = reinterpret_cast<float *>(sc + l.off_nodes);
    int32_t *d_edges = reinterpret_cast<int32_t *>(sc + l.off_edges), *d_dag = reinterpret_cast<int32_t *>(sc + l.off_dag),
            *d_sup = reinterpret_cast<int32_t *>(sc + l.off_sup);
    k_pack_scan<<<1, 1024, 0, s>>>(p, d_off);
    k_pack_copy<<<(p.B + 3) / 4, 128, 0, s>>>(p, d_off, d_nodes, d_edges, d_dag, d_sup);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out->offsets, d_off, sizeof(int32_t) * 3 * ((size_t)p.B + 1), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));  // the totals size the copies below
    const int64_t tn = out->offsets[3 * p.B], te = out->offsets[3 * p.B + 1], tj = out->offsets[3 * p.B + 2];
    if (tn > out->node_capacity || te > out->edge_capacity || tj > out->job_capacity) return SSB_E_WORKSPACE;
    if (out->nodes) CUDA_TRY(cudaMemcpyAsync(out->nodes, d_nodes, sizeof(float) * 3 * tn, cudaMemcpyDeviceToHost, s));
    if (out->edge_links) CUDA_TRY(cudaMemcpyAsync(out->edge_links, d_edges, sizeof(int32_t) * 2 * te, cudaMemcpyDeviceToHost, s));
    if (out->dag_ptr) CUDA_TRY(cudaMemcpyAsync(out->dag_ptr, d_dag, sizeof(int32_t) * (tj + p.B), cudaMemcpyDeviceToHost, s));
    if (out->exec_supplies) CUDA_TRY(cudaMemcpyAsync(out->exec_supplies, d_sup, sizeof(int32_t) * tj, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return SSB_OK;
}

int ssb_rollout_fair_traj(ssb_env *env, int32_t num_decisions, int32_t dynamic_partition, int32_t auto_reset,
                          uint64_t seed_step, ssb_transition *traj, void *stream)
{
    if (!env || num_decisions < 0) return SSB_E_INVALID;
    SSB_ON_DEVICE(env);
    if (env->p.E <= 32)
        k_rollout_fair<1><<<env->grid, WARPS_PER_CTA * 32, 0, (cudaStream_t)stream>>>(
            env->p, num_decisions, dynamic_partition, auto_reset, seed_step, traj);
    else
        k_rollout_fair<2><<<env->grid, WARPS_PER_CTA * 32, 0, (cudaStream_t)stream>>>(
            env->p, num_decisions, dynamic_partition, auto_reset, seed_step, traj);
    CUDA_TRY(cudaGetLastError());
    SSB_MARK(env, stream);
    return SSB_OK;
}

int ssb_rollout_fair_async(ssb_env *env, int32_t max_decisions, double rollout_duration, int32_t dynamic_partition,
                           uint64_t seed_step, ssb_transition *traj, int32_t *num_steps, double *elapsed, void *stream)
{
    if (!env || max_decisions < 0 || !(rollout_duration > 0.0)) return SSB_E_INVALID;
    SSB_ON_DEVICE(env);
    if (env->p.E <= 32)
        k_rollout_fair_async<1><<<env->grid, WARPS_PER_CTA * 32, 0, (cudaStream_t)stream>>>(
            env->p, max_decisions, rollout_duration, dynamic_partition, seed_step, traj, num_steps, elapsed);
    else
        k_rollout_fair_async<2><<<env->grid, WARPS_PER_CTA * 32, 0, (cudaStream_t)stream>>>(
            env->p, max_decisions, rollout_duration, dynamic_partition, seed_step, traj, num_steps, elapsed);
    CUDA_TRY(cudaGetLastError());
    SSB_MARK(env, stream);
    return SSB_OK;
}

int ssb_rollout_fair(ssb_env *env, int32_t num_decisions, int32_t dynamic_partition, int32_t auto_reset,
                     uint64_t seed_step, void *stream)
{
    return ssb_rollout_fair_traj(env, num_decisions, dynamic_partition, auto_reset, seed_step, nullptr, stream);
}

int ssb_fair_actions(ssb_env *env, int32_t dynamic_partition, int32_t *stage_idx, int32_t *num_exec, void *stream)
{
    if (!env || !stage_idx || !num_exec) return SSB_E_INVALID;
    SSB_ON_DEVICE(env);
    k_fair_actions<<<env->grid, WARPS_PER_CTA * 32, 0, (cudaStream_t)stream>>>(env->p, dynamic_partition,
                                                                               stage_idx, num_exec);
    CUDA_TRY(cudaGetLastError());
    SSB_MARK(env, stream);
    return SSB_OK;
}

int ssb_get_views(ssb_env *env, ssb_views *out)
{
    if (!env || !out) return SSB_E_INVALID;
    out->hdr = env->p.obs_hdr;
    out->nodes = env->p.obs_nodes;
    out->edge_links = env->p.obs_edges;
    out->dag_ptr = env->p.obs_dag_ptr;
    out->exec_supplies = env->p.obs_supplies;
    out->node_stride = env->p.Sc;
    out->edge_stride = env->p.Mc;
    out->job_stride = env->p.Jc;
    out->pad = 0;
    return SSB_OK;
}

int ssb_decima_obs(ssb_env *env, void *stream)
{
    if (!env || !env->p.dec_feat) return SSB_E_INVALID;  // needs SSB_FLAG_DECIMA_OBS
    SSB_ON_DEVICE(env);
    k_decima_obs<<<env->grid, WARPS_PER_CTA * 32, 0, (cudaStream_t)stream>>>(env->p);
    CUDA_TRY(cudaGetLastError());
    SSB_MARK(env, stream);
    return SSB_OK;
}

int ssb_get_decima_views(ssb_env *env, ssb_decima_views *out)
{
    if (!env || !out || !env->p.dec_feat) return SSB_E_INVALID;
    out->features = env->p.dec_feat;
    out->stage_mask = env->p.dec_stage_mask;
    out->frontier_mask = env->p.dec_frontier_mask;
    out->commit_caps = env->p.dec_caps;
    out->edge_bits = env->p.dec_edge_bits;
    out->depth = env->p.dec_depth;
    out->node_stride = env->p.Sc;
    out->edge_stride = env->p.Mc;
    out->job_stride = env->p.Jc;
    out->pad = 0;
    return SSB_OK;
}

int ssb_set_decima_weights(ssb_env *env, const float *weights, int32_t n_floats)
{
    if (!env || !weights || !env->p.pol_w || n_floats != dw::TOTAL) return SSB_E_INVALID;
    CUDA_TRY(cudaSetDevice(env->device));
    // state_dict order ([out][in] per Linear) -> device layout (transposed, 4-float padded; ssb_decima.cuh)
    static const int dims[7][4] = {{5, 32, 16, 16},  {16, 32, 16, 16}, {16, 32, 16, 16}, {21, 32, 16, 16},
                                   {16, 32, 16, 16}, {53, 64, 64, 1},  {36, 64, 64, 1}};
    std::vector<float> dev(dd::TOTAL, 0.0f);
    size_t src = 0, dst = 0;
    for (int m = 0; m < 7; m++) {
        for (int l = 0; l < 3; l++) {
            const int in = dims[m][l], out = dims[m][l + 1];
            for (int o = 0; o < out; o++)
                for (int i = 0; i < in; i++) dev[dst + (size_t)i * out + o] = weights[src + (size_t)o * in + i];
            src += (size_t)in * out;
            dst += dd::pad4(in * out);
            for (int o = 0; o < out; o++) dev[dst + o] = weights[src + o];
            src += out;
            dst += dd::pad4(out);
        }
    }
    if (src != (size_t)dw::TOTAL || dst != (size_t)dd::TOTAL) return SSB_E_INVALID;
    CUDA_TRY(cudaMemcpy(env->p.pol_w, dev.data(), sizeof(float) * dd::TOTAL, cudaMemcpyHostToDevice));
    // tensor-core path: per-stage blobs (canonical UMMA tiles, tf32 hi/lo halves, biases); padding stays zero
    CUDA_TRY(cudaMemset(env->p.pol_wblob, 0, sizeof(float) * tc::BLOB_TOTAL));
    tc::k_build_blob<tc::ST_PREP><<<1, 128>>>(env->p);
    tc::k_build_blob<tc::ST_SINK><<<1, 128>>>(env->p);
    tc::k_build_blob<tc::ST_MSG><<<1, 128>>>(env->p);
    tc::k_build_blob<tc::ST_RCV><<<1, 128>>>(env->p);
    tc::k_build_blob<tc::ST_DAG><<<1, 128>>>(env->p);
    tc::k_build_blob<tc::ST_GLOB><<<1, 128>>>(env->p);
    tc::k_build_blob<tc::ST_STAGE><<<1, 128>>>(env->p);
    tc::k_build_blob<tc::ST_EXEC><<<1, 128>>>(env->p);
    CUDA_TRY(cudaMemset(env->p.pol_wblob3, 0, sizeof(uint32_t) * fz::BLOB_TOTAL));
    fz::k_build_blob<tc::ST_PREP><<<1, 128>>>(env->p.pol_w, env->p.pol_wblob3);
    fz::k_build_blob<tc::ST_SINK><<<1, 128>>>(env->p.pol_w, env->p.pol_wblob3);
    fz::k_build_blob<tc::ST_MSG><<<1, 128>>>(env->p.pol_w, env->p.pol_wblob3);
    fz::k_build_blob<tc::ST_RCV><<<1, 128>>>(env->p.pol_w, env->p.pol_wblob3);
    fz::k_build_blob<tc::ST_DAG><<<1, 128>>>(env->p.pol_w, env->p.pol_wblob3);
    fz::k_build_blob<tc::ST_GLOB><<<1, 128>>>(env->p.pol_w, env->p.pol_wblob3);
    fz::k_build_blob<tc::ST_STAGE><<<1, 128>>>(env->p.pol_w, env->p.pol_wblob3);
    fz::k_build_blob<tc::ST_EXEC><<<1, 128>>>(env->p.pol_w, env->p.pol_wblob3);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaDeviceSynchronize());
    return SSB_OK;
}

// The backward pass and ssb_decima_work walk the row lists of the list-driven path; the fused forward kernel does not
// build them, so they are (re)built here from the observation, the adapter's outputs and the stored action.
__global__ void __launch_bounds__(128) k_plan_exec(Params p)
{
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= p.B) return;
    const int job_idx = p.pol_action[(size_t)b * 4 + 1];
    const int cap = job_idx >= 0 ? p.dec_caps[(size_t)b * p.Jc + job_idx] : 0;
    int base = 0;
    if (lane == 0 && cap > 0) base = atomicAdd(&p.pl_cnt[tc::CNT_EXEC], cap);
    base = __shfl_sync(0xffffffffu, base, 0);
    for (int c = lane; c < cap; c += 32) p.pl_exec[base + c] = b * p.Epad + c;
}
static int replan(ssb_env *env, cudaStream_t s)
{
    if (env->policy_mode != POLICY_FUSED) return SSB_OK;  // the list-driven forward pass left its lists in place
    const Params &p = env->p;
    const int warp_grid = (p.B + 3) / 4;
    CUDA_TRY(cudaMemsetAsync(p.pl_cnt, 0, sizeof(int32_t) * tc::CNT_TOTAL, s));
    tc::k_pol_plan_a<<<warp_grid, 128, 0, s>>>(p);
    tc::k_pol_plan_scan<<<1, 32, 0, s>>>(p);
    tc::k_pol_plan_b<<<warp_grid, 128, 0, s>>>(p);
    k_plan_exec<<<warp_grid, 128, 0, s>>>(p);
    CUDA_TRY(cudaGetLastError());
    return SSB_OK;
}

// run_adapter = false: the adapter's outputs are already in place (restored from a snapshot);
// advance_draws = false: the Philox policy stream of the envs is left where it is (pure evaluation)
static int decima_policy_impl(ssb_env *env, const int32_t *forced_stage, const int32_t *forced_num_exec,
                              int32_t *stage_idx_out, int32_t *num_exec_out, bool run_adapter, bool advance_draws,
                              cudaStream_t s, const uint8_t *active = nullptr)
{
    Params p = env->p;
    p.pol_active = active;
    if (env->policy_mode == POLICY_FUSED) {
        // the whole decision of every env in one persistent kernel (ssb_decima_fused.cuh)
        CUDA_TRY(cudaMemsetAsync(p.fz_cursor, 0, sizeof(int32_t) * 4, s));
        fz::fused::Args a{forced_stage, forced_num_exec, stage_idx_out, num_exec_out, p.fz_cursor,
                          run_adapter ? 1 : 0, advance_draws ? 1 : 0, env->fused_group};
        const int groups = (p.B + env->fused_group - 1) / env->fused_group;
        fz::fused::k_decima_fused<<<std::min(groups, 2 * env->num_sms), fz::fused::THREADS, fz::fused::Smem::BYTES, s>>>(p, a);
        CUDA_TRY(cudaGetLastError());
        SSB_MARK(env, s);
        return SSB_OK;
    }
    // observation adapter -> row lists -> one tensor-core tile pass per MLP (lists: ssb_decima_tc.cuh)
    const int32_t *cnt = p.pl_cnt;
    const int warp_grid = (p.B + 3) / 4;
    int rc;
    CUDA_TRY(cudaMemsetAsync(p.pl_cnt, 0, sizeof(int32_t) * tc::CNT_TOTAL, s));
    if (run_adapter) k_decima_obs<<<env->grid, WARPS_PER_CTA * 32, 0, s>>>(p);
    tc::k_pol_plan_a<<<warp_grid, 128, 0, s>>>(p);
    tc::k_pol_plan_scan<<<1, 32, 0, s>>>(p);
    tc::k_pol_plan_b<<<warp_grid, 128, 0, s>>>(p);
    CUDA_TRY(cudaGetLastError());
    if ((rc = launch_tile<tc::ST_PREP>(env, p.pl_all, nullptr, cnt + tc::CNT_ALL, 0, 4, s))) return rc;
    if ((rc = launch_tile<tc::ST_SINK>(env, p.pl_sink, nullptr, cnt + tc::CNT_SINK, 0, 4, s))) return rc;
    for (int k = env->dmax - 1; k >= 0; k--) {  // reversed(edge_masks) (scheduler.py:214-232)
        if ((rc = launch_tile<tc::ST_MSG>(env, p.pl_lvl, cnt + tc::OFF_LVL + 2 * k, cnt + tc::CNT_LVL + 2 * k, k, 4, s)))
            return rc;
        if ((rc = launch_tile<tc::ST_RCV>(env, p.pl_lvl, cnt + tc::OFF_LVL + 2 * k + 1, cnt + tc::CNT_LVL + 2 * k + 1,
                                          k, 4, s)))
            return rc;
    }
    if ((rc = launch_tile<tc::ST_DAG>(env, p.pl_all, nullptr, cnt + tc::CNT_ALL, 0, 4, s))) return rc;
    if ((rc = launch_tile<tc::ST_GLOB>(env, p.pl_jobs, nullptr, cnt + tc::CNT_JOBS, 0, 4, s))) return rc;
    tc::k_pol_glob_sum<<<(p.B + 7) / 8, 128, 0, s>>>(p);
    // (score heads: four CTAs per SM in the TMEM path; round 1's shared-memory tiles fit one)
    const int head_ctas = env->policy_mode == POLICY_TILES_TF32 ? 1 : 4;
    if ((rc = launch_tile<tc::ST_STAGE>(env, nullptr, nullptr, cnt + tc::CNT_CAND, 0, head_ctas, s))) return rc;
    tc::k_pol_sample_stage<<<warp_grid, 128, 0, s>>>(p, forced_stage);
    if ((rc = launch_tile<tc::ST_EXEC>(env, p.pl_exec, nullptr, cnt + tc::CNT_EXEC, 0, head_ctas, s))) return rc;
    tc::k_pol_sample_exec<<<warp_grid, 128, 0, s>>>(p, forced_num_exec, stage_idx_out, num_exec_out,
                                                    advance_draws ? 1 : 0);
    CUDA_TRY(cudaGetLastError());
    SSB_MARK(env, s);
    return SSB_OK;
}

int ssb_decima_policy(ssb_env *env, const int32_t *forced_stage, const int32_t *forced_num_exec,
                      int32_t *stage_idx_out, int32_t *num_exec_out, void *stream)
{
    if (!env || !env->p.pol_w) return SSB_E_INVALID;  // needs SSB_FLAG_DECIMA_POLICY
    SSB_ON_DEVICE(env);
    return decima_policy_impl(env, forced_stage, forced_num_exec, stage_idx_out, num_exec_out, true, true,
                              (cudaStream_t)stream);
}

// ---- stored observations (RolloutBuffer.obsns) and their re-evaluation (DecimaScheduler.evaluate_actions)
namespace {
struct SnapPart { void *ptr; size_t bytes; };
int snapshot_parts(const ssb_env *env, SnapPart *out)
{
    const Params &p = env->p;
    const size_t B = p.B;
    int n = 0;
    out[n++] = {p.obs_hdr, B * sizeof(ssb_obs_hdr)};
    out[n++] = {p.obs_edges, B * p.Mc * 2 * sizeof(int32_t)};
    out[n++] = {p.obs_dag_ptr, B * (p.Jc + 1) * sizeof(int32_t)};
    out[n++] = {p.dec_feat, B * p.Sc * 5 * sizeof(float)};
    out[n++] = {p.dec_stage_mask, B * p.Sc};
    out[n++] = {p.dec_caps, B * p.Jc * sizeof(int32_t)};
    out[n++] = {p.dec_edge_bits, B * p.Mc * sizeof(uint64_t)};
    out[n++] = {p.dec_depth, B * sizeof(int32_t)};
    return n;
}
size_t snapshot_bytes(const ssb_env *env)
{
    SnapPart parts[8];
    const int n = snapshot_parts(env, parts);
    size_t total = 0;
    for (int i = 0; i < n; i++) total += (parts[i].bytes + 255) & ~size_t(255);
    return total;
}
int snapshot_copy(const ssb_env *env, char *buf, bool to_buf, cudaStream_t s)
{
    SnapPart parts[8];
    const int n = snapshot_parts(env, parts);
    size_t off = 0;
    for (int i = 0; i < n; i++) {
        if (to_buf) CUDA_TRY(cudaMemcpyAsync(buf + off, parts[i].ptr, parts[i].bytes, cudaMemcpyDeviceToDevice, s));
        else CUDA_TRY(cudaMemcpyAsync(parts[i].ptr, buf + off, parts[i].bytes, cudaMemcpyDeviceToDevice, s));
        off += (parts[i].bytes + 255) & ~size_t(255);
    }
    return SSB_OK;
}
}  // namespace

// slot i of dst <- the stored observation of sample (src_step[i], src_env[i]); only the rows in use are copied
namespace {
struct GatherParts {
    size_t off[8];      // byte offset of each part inside a snapshot block
    size_t stride[8];   // bytes per environment
};
}  // namespace
__global__ void __launch_bounds__(128)
k_snapshot_gather(Params p, GatherParts gp, const char *src, size_t block_bytes, int num_steps, const int32_t *src_step,
                  const int32_t *src_env, char *dst)
{
    const int i = blockIdx.x, tid = threadIdx.x;
    if (i >= p.B) return;
    const int k = src_step[i], b = src_env[i];
    ssb_obs_hdr *dh = reinterpret_cast<ssb_obs_hdr *>(dst + gp.off[0]) + i;
    if (k < 0 || k >= num_steps || b < 0 || b >= p.B) {  // empty slot: an observation that takes no part
        if (tid == 0) {
            ssb_obs_hdr e = {};
            e.terminated = 1;
            *dh = e;
            reinterpret_cast<int32_t *>(dst + gp.off[7])[i] = 0;
        }
        return;
    }
    const char *blk = src + (size_t)k * block_bytes;
    const ssb_obs_hdr sh = reinterpret_cast<const ssb_obs_hdr *>(blk + gp.off[0])[b];
    if (tid == 0) {
        *dh = sh;
        reinterpret_cast<int32_t *>(dst + gp.off[7])[i] = reinterpret_cast<const int32_t *>(blk + gp.off[7])[b];
    }
    auto copy4 = [&](int part, size_t bytes) {  // 4-byte words (every part but the stage mask is int32 / f32 / u64 data)
        const uint32_t *s = reinterpret_cast<const uint32_t *>(blk + gp.off[part] + (size_t)b * gp.stride[part]);
        uint32_t *d = reinterpret_cast<uint32_t *>(dst + gp.off[part] + (size_t)i * gp.stride[part]);
        for (size_t w = tid; w < bytes / 4; w += 128) d[w] = s[w];
    };
    copy4(1, (size_t)sh.num_edges * 8);              // edge links
    copy4(2, ((size_t)sh.num_active_jobs + 1) * 4);  // dag_ptr
    copy4(3, (size_t)sh.num_nodes * 20);             // node features
    copy4(5, (size_t)sh.num_active_jobs * 4);        // commit caps
    copy4(6, (size_t)sh.num_edges * 8);              // per-edge level bits
    const uint8_t *sm = reinterpret_cast<const uint8_t *>(blk + gp.off[4] + (size_t)b * gp.stride[4]);
    uint8_t *dm = reinterpret_cast<uint8_t *>(dst + gp.off[4] + (size_t)i * gp.stride[4]);
    for (int w = tid; w < sh.num_nodes; w += 128) dm[w] = sm[w];
}

int ssb_decima_snapshot_gather(ssb_env *env, const void *snapshots, int32_t num_steps, const int32_t *src_step,
                               const int32_t *src_env, void *dst, void *stream)
{
    if (!env || !snapshots || !src_step || !src_env || !dst || num_steps < 1 || !env->p.dec_feat) return SSB_E_INVALID;
    SSB_ON_DEVICE(env);
    const Params &p = env->p;
    SnapPart parts[8];
    const int n = snapshot_parts(env, parts);
    if (n != 8) return SSB_E_INVALID;
    GatherParts gp;
    size_t off = 0;
    for (int q = 0; q < 8; q++) {
        gp.off[q] = off;
        gp.stride[q] = parts[q].bytes / (size_t)p.B;
        off += (parts[q].bytes + 255) & ~size_t(255);
    }
    k_snapshot_gather<<<p.B, 128, 0, (cudaStream_t)stream>>>(p, gp, static_cast<const char *>(snapshots), snapshot_bytes(env),
                                                             num_steps, src_step, src_env, static_cast<char *>(dst));
    CUDA_TRY(cudaGetLastError());
    SSB_MARK(env, stream);
    return SSB_OK;
}

int ssb_decima_head_adjoint(ssb_env *env, const float *grad_lgprob, const float *grad_entropy,
                            float *grad_stage_logits, float *grad_exec_logits, void *stream)
{
    if (!env || !env->p.pol_w || !grad_lgprob || !grad_entropy || !grad_stage_logits || !grad_exec_logits)
        return SSB_E_INVALID;
    SSB_ON_DEVICE(env);
    CUDA_TRY(bwd::head_adjoint(env->p, grad_lgprob, grad_entropy, grad_stage_logits, grad_exec_logits,
                               (cudaStream_t)stream));
    SSB_MARK(env, stream);
    return SSB_OK;
}

int ssb_decima_head_backward(ssb_env *env, const float *grad_stage_logits, const float *grad_exec_logits,
                             float *grad_weights, float *grad_stage_inputs, float *grad_exec_inputs,
                             float *stage_inputs, float *exec_inputs, int32_t *num_rows, void *stream)
{
    if (!env || !env->p.pol_w || !grad_stage_logits || !grad_exec_logits || !grad_weights) return SSB_E_INVALID;
    SSB_ON_DEVICE(env);
    cudaStream_t s = (cudaStream_t)stream;
    const Params &p = env->p;
    {
        const int rcp = replan(env, s);
        if (rcp) return rcp;
    }
    const bwd::Bufs none{nullptr, nullptr, nullptr, nullptr, nullptr};
    CUDA_TRY(bwd::mlp_backward(tc::ST_STAGE, p, env->num_sms, nullptr, nullptr, p.pl_cnt + tc::CNT_CAND, 0,
                               grad_stage_logits, grad_stage_inputs, stage_inputs, grad_weights, none, false, s));
    CUDA_TRY(bwd::mlp_backward(tc::ST_EXEC, p, env->num_sms, p.pl_exec, nullptr, p.pl_cnt + tc::CNT_EXEC, 0,
                               grad_exec_logits, grad_exec_inputs, exec_inputs, grad_weights, none, false, s));
    if (num_rows) {
        int32_t c[tc::CNT_OVERFLOW + 1];
        CUDA_TRY(cudaMemcpyAsync(c, p.pl_cnt, sizeof(c), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        num_rows[0] = c[tc::CNT_CAND];
        num_rows[1] = c[tc::CNT_EXEC];
    }
    SSB_MARK(env, s);
    return SSB_OK;
}

int ssb_decima_backward_bytes(ssb_env *env, size_t *bytes)
{
    if (!env || !bytes || !env->p.pol_w) return SSB_E_INVALID;
    *bytes = backward_scratch(env->p, nullptr).floats * sizeof(float);
    return SSB_OK;
}

int ssb_decima_backward(ssb_env *env, const float *grad_lgprob, const float *grad_entropy, float *grad_weights,
                        float *grad_node_embeddings, int32_t through_node_encoder, void *scratch, void *stream)
{
    if (!env || !env->p.pol_w || !grad_lgprob || !grad_entropy || !grad_weights || !grad_node_embeddings || !scratch ||
        (reinterpret_cast<uintptr_t>(scratch) & 15))
        return SSB_E_INVALID;
    SSB_ON_DEVICE(env);
    cudaStream_t s = (cudaStream_t)stream;
    const Params &p = env->p;
    const BackwardScratch b = backward_scratch(p, static_cast<float *>(scratch));
    {
        const int rcp = replan(env, s);
        if (rcp) return rcp;
    }
    CUDA_TRY(cudaMemsetAsync(grad_node_embeddings, 0, sizeof(float) * (size_t)p.B * p.Sc * 16, s));
    CUDA_TRY(cudaMemsetAsync(b.d_hdag, 0, sizeof(float) * (size_t)p.B * p.Jc * 16, s));
    CUDA_TRY(cudaMemsetAsync(b.d_hglob, 0, sizeof(float) * (size_t)p.B * 16, s));
    CUDA_TRY(bwd::head_adjoint(p, grad_lgprob, grad_entropy, b.gs, b.ge, s));
    const bwd::Bufs bw{grad_node_embeddings, b.d_hdag, b.d_hglob, b.d_hinit, b.d_msg};
    const int32_t *cnt = p.pl_cnt;
    float *gw = grad_weights;
    int rc;
    // heads first (their input gradients feed all three summaries), then the global summary, then the job summaries
    if ((rc = launch_mlp_backward(tc::ST_STAGE, env, nullptr, nullptr, cnt + tc::CNT_CAND, 0, b.gs, gw, bw, s))) return rc;
    if ((rc = launch_mlp_backward(tc::ST_EXEC, env, p.pl_exec, nullptr, cnt + tc::CNT_EXEC, 0, b.ge, gw, bw, s))) return rc;
    if ((rc = launch_mlp_backward(tc::ST_GLOB, env, p.pl_jobs, nullptr, cnt + tc::CNT_JOBS, 0, nullptr, gw, bw, s))) return rc;
    if ((rc = launch_mlp_backward(tc::ST_DAG, env, p.pl_all, nullptr, cnt + tc::CNT_ALL, 0, nullptr, gw, bw, s))) return rc;
    if (!through_node_encoder) { SSB_MARK(env, s); return SSB_OK; }
    // NodeEncoder (scheduler.py:191-234), the levels in the reverse of the forward order.  Level k's backward needs
    // the embeddings as they were BEFORE level k and the level's messages; the forward pass overwrites both in
    // place, so they are recomputed: reset (PREP), sinks, levels dmax-1 .. k+1, then level k's messages.  First
    // correct version: O(depth^2) tile passes instead of saving every level's rows.
    CUDA_TRY(cudaMemsetAsync(b.d_hinit, 0, sizeof(float) * (size_t)p.B * p.Sc * 16, s));
    CUDA_TRY(cudaMemsetAsync(b.d_msg, 0, sizeof(float) * (size_t)p.B * p.Sc * 16, s));
    for (int k = 0; k < env->dmax; k++) {
        if ((rc = launch_tile<tc::ST_PREP>(env, p.pl_all, nullptr, cnt + tc::CNT_ALL, 0, 4, s))) return rc;
        if ((rc = launch_tile<tc::ST_SINK>(env, p.pl_sink, nullptr, cnt + tc::CNT_SINK, 0, 4, s))) return rc;
        for (int j = env->dmax - 1; j > k; j--) {
            if ((rc = launch_tile<tc::ST_MSG>(env, p.pl_lvl, cnt + tc::OFF_LVL + 2 * j, cnt + tc::CNT_LVL + 2 * j, j, 4, s)))
                return rc;
            if ((rc = launch_tile<tc::ST_RCV>(env, p.pl_lvl, cnt + tc::OFF_LVL + 2 * j + 1, cnt + tc::CNT_LVL + 2 * j + 1,
                                              j, 4, s)))
                return rc;
        }
        if ((rc = launch_tile<tc::ST_MSG>(env, p.pl_lvl, cnt + tc::OFF_LVL + 2 * k, cnt + tc::CNT_LVL + 2 * k, k, 4, s)))
            return rc;
        if ((rc = launch_mlp_backward(tc::ST_RCV, env, p.pl_lvl, cnt + tc::OFF_LVL + 2 * k + 1,
                                                  cnt + tc::CNT_LVL + 2 * k + 1, k, nullptr, gw, bw, s)))
            return rc;
        if ((rc = launch_mlp_backward(tc::ST_MSG, env, p.pl_lvl, cnt + tc::OFF_LVL + 2 * k, cnt + tc::CNT_LVL + 2 * k, k,
                                                  nullptr, gw, bw, s)))
            return rc;
    }
    if ((rc = launch_mlp_backward(tc::ST_SINK, env, p.pl_sink, nullptr, cnt + tc::CNT_SINK, 0, nullptr, gw, bw, s))) return rc;
    if ((rc = launch_mlp_backward(tc::ST_PREP, env, p.pl_all, nullptr, cnt + tc::CNT_ALL, 0, nullptr, gw, bw, s))) return rc;
    SSB_MARK(env, s);
    return SSB_OK;
}

int ssb_decima_snapshot_bytes(ssb_env *env, size_t *bytes)
{
    if (!env || !bytes || !env->p.dec_feat) return SSB_E_INVALID;
    *bytes = snapshot_bytes(env);
    return SSB_OK;
}

int ssb_decima_snapshot(ssb_env *env, void *dst, void *stream)
{
    if (!env || !dst || !env->p.dec_feat) return SSB_E_INVALID;
    SSB_ON_DEVICE(env);
    k_decima_obs<<<env->grid, WARPS_PER_CTA * 32, 0, (cudaStream_t)stream>>>(env->p);  // the adapter's view of the state
    CUDA_TRY(cudaGetLastError());
    const int rcs = snapshot_copy(env, static_cast<char *>(dst), true, (cudaStream_t)stream);
    if (rcs) return rcs;
    SSB_MARK(env, stream);
    return SSB_OK;
}

int ssb_decima_snapshot_load(ssb_env *env, const void *snapshot, void *stream)
{
    if (!env || !snapshot || !env->p.pol_w || env->snap_loaded) return SSB_E_INVALID;
    SSB_ON_DEVICE(env);
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    // the live observation is parked in the handle's scratch while the stored one is worked on
    if ((rc = snapshot_copy(env, env->p.pol_snap, true, s))) return rc;
    if ((rc = snapshot_copy(env, const_cast<char *>(static_cast<const char *>(snapshot)), false, s))) return rc;
    env->snap_loaded = 1;
    SSB_MARK(env, stream);
    return SSB_OK;
}

int ssb_decima_snapshot_unload(ssb_env *env, void *stream)
{
    if (!env || !env->p.pol_w || !env->snap_loaded) return SSB_E_INVALID;
    SSB_ON_DEVICE(env);
    env->snap_loaded = 0;
    const int rcs = snapshot_copy(env, env->p.pol_snap, false, (cudaStream_t)stream);
    if (rcs) return rcs;
    SSB_MARK(env, stream);
    return SSB_OK;
}

int ssb_decima_evaluate(ssb_env *env, const void *snapshot, const int32_t *stage_sel, const int32_t *exec_sel,
                        float *lgprob_out, float *entropy_out, void *stream)
{
    if (!env || !stage_sel || !exec_sel || !env->p.pol_w) return SSB_E_INVALID;
    SSB_ON_DEVICE(env);
    if (!snapshot && !env->snap_loaded) return SSB_E_INVALID;  // NULL: the snapshot ssb_decima_snapshot_load put in place
    cudaStream_t s = (cudaStream_t)stream;
    const size_t B = env->p.B;
    int rc;
    if (snapshot && (rc = ssb_decima_snapshot_load(env, snapshot, stream))) return rc;
    // whatever happens below, a snapshot this call loaded is unloaded again (the live observation comes back)
    rc = decima_policy_impl(env, stage_sel, exec_sel, nullptr, nullptr, false, false, s);
    if (!rc && lgprob_out &&
        cudaMemcpyAsync(lgprob_out, env->p.pol_lgprob, B * 4, cudaMemcpyDeviceToDevice, s) != cudaSuccess) rc = SSB_E_CUDA;
    if (!rc && entropy_out &&
        cudaMemcpyAsync(entropy_out, env->p.pol_entropy, B * 4, cudaMemcpyDeviceToDevice, s) != cudaSuccess) rc = SSB_E_CUDA;
    if (rc == SSB_E_CUDA && !g_cuda_err[0]) snprintf(g_cuda_err, sizeof(g_cuda_err), "ssb_decima_evaluate: %s",
                                                      cudaGetErrorString(cudaGetLastError()));
    if (snapshot) {
        const int rc2 = ssb_decima_snapshot_unload(env, stream);
        if (!rc) rc = rc2;
    }
    if (!rc) SSB_MARK(env, stream);
    return rc;
}

// ---- fixed-duration Decima rollouts spanning resets (RolloutWorkerAsync.collect_rollout, rollout_worker.py:160-206)
// round = { who takes part ; policy ; row + step (or reset) ; bookkeeping }
__global__ void k_dasync_begin(Params p, double duration, int max_rows)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b == 0) *p.as_any = 0;
    if (b >= p.B) return;
    const ssb_obs_hdr &o = p.obs_hdr[b];
    int kind = 0;
    if (p.as_elapsed[b] < duration && p.as_rows[b] < max_rows && !(o.error && o.error != SSB_ENV_DONE) && !p.hdr[b].error)
        kind = (p.hdr[b].done || o.truncated) ? 2 : 1;  // the reset of :196-200, applied when the loop comes back around
    p.as_kind[b] = (uint8_t)kind;
}
__global__ void k_dasync_pre(Params p, const int32_t *a, const int32_t *n, ssb_transition *traj, int max_rows)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.B || p.as_kind[b] != 1) return;
    ssb_transition t;  // rollout_buffer.add(obs, elapsed_time, action, lgprob, reward) (:191): reward filled in below
    t.wall_time = p.as_elapsed[b]; t.reward = 0.0; t.stage_idx = a[b]; t.num_exec = n[b];
    t.flags = p.as_fresh[b] ? 4 : 0;
    t.lgprob = p.pol_lgprob[b];
    traj[(size_t)b * max_rows + p.as_rows[b]] = t;
    p.as_wall0[b] = p.obs_hdr[b].wall_time;
}
__global__ void k_dasync_post(Params p, ssb_transition *traj, int max_rows, double duration)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.B) return;
    const int kind = p.as_kind[b];
    if (kind == 2) { p.as_fresh[b] = 1; atomicAdd(p.as_any, 1); return; }
    if (kind != 1) return;
    const ssb_obs_hdr &o = p.obs_hdr[b];
    ssb_transition &t = traj[(size_t)b * max_rows + p.as_rows[b]];
    t.reward = o.reward;
    t.flags |= (o.terminated ? 1 : 0) | (o.truncated ? 2 : 0);
    p.as_elapsed[b] += o.wall_time - p.as_wall0[b];  // the duration of this step (:194)
    p.as_rows[b] += 1;
    p.as_fresh[b] = 0;
    if (p.as_elapsed[b] < duration && p.as_rows[b] < max_rows && !o.error) atomicAdd(p.as_any, 1);
}

__global__ void k_traj_next(Params p) { if (threadIdx.x == 0 && blockIdx.x == 0) *p.traj_d += 1; }

// one decision of ssb_rollout_decima, enqueued on s (captured into a CUDA graph by the caller)
static int decima_decision(ssb_env *env, int32_t num_decisions, int32_t max_events, ssb_transition *traj, cudaStream_t s)
{
    const Params &p = env->p;
    const int tb = (p.B + 127) / 128;
    int rc = ssb_decima_policy(env, nullptr, nullptr, p.pol_act_a, p.pol_act_n, s);
    if (rc) return rc;
    if (traj) k_traj_pre<<<tb, 128, 0, s>>>(p, p.pol_act_a, p.pol_act_n, traj, num_decisions);
    if ((rc = ssb_step(env, p.pol_act_a, p.pol_act_n, nullptr, max_events, s))) return rc;
    if (traj) k_traj_post<<<tb, 128, 0, s>>>(p, traj, num_decisions);
    k_traj_next<<<1, 32, 0, s>>>(p);
    CUDA_TRY(cudaGetLastError());
    return SSB_OK;
}

int ssb_rollout_decima(ssb_env *env, int32_t num_decisions, int32_t max_events, ssb_transition *traj, void *stream)
{
    if (!env || !env->p.pol_w || num_decisions < 0) return SSB_E_INVALID;
    SSB_ON_DEVICE(env);
    cudaStream_t caller = (cudaStream_t)stream, s = caller;
    // the legacy default stream cannot be captured: run on the handle's own stream, ordered after / before the
    // caller's stream with events
    const bool hop = caller == nullptr || caller == cudaStreamLegacy || caller == cudaStreamPerThread;
    if (hop && !env->no_graph) {
        s = env->own_stream;
        CUDA_TRY(cudaEventRecord(env->ev, caller));
        CUDA_TRY(cudaStreamWaitEvent(s, env->ev, 0));
    }
    CUDA_TRY(cudaMemsetAsync(env->p.traj_d, 0, sizeof(int32_t), s));
    // The ~50 launches of one decision are captured once into a CUDA graph and replayed: the kernels are short
    // (10-100 us) and strictly dependent, so the per-launch gaps are a visible share of a decision.
    // (SSB_NO_GRAPH=1 launches them one by one.)
    const bool same = env->dg_exec && env->dg_traj == traj && env->dg_k == num_decisions && env->dg_events == max_events &&
                      env->dg_autoreset == env->auto_reset && env->dg_seed_step == env->auto_seed_step;
    if (!same && !env->no_graph) {
        if (env->dg_exec) { cudaGraphExecDestroy(env->dg_exec); env->dg_exec = nullptr; }
        cudaGraph_t g = nullptr;
        CUDA_TRY(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        const int rc = decima_decision(env, num_decisions, max_events, traj, s);
        const cudaError_t ce = cudaStreamEndCapture(s, &g);
        if (rc || ce != cudaSuccess || !g) {
            if (g) cudaGraphDestroy(g);
            cudaGetLastError();
            env->no_graph = 1;  // fall back to plain launches for this handle
        } else {
            const cudaError_t ie = cudaGraphInstantiate(&env->dg_exec, g, 0);
            cudaGraphDestroy(g);
            if (ie != cudaSuccess) { env->dg_exec = nullptr; env->no_graph = 1; cudaGetLastError(); }
            env->dg_traj = traj; env->dg_k = num_decisions; env->dg_events = max_events;
            env->dg_autoreset = env->auto_reset; env->dg_seed_step = env->auto_seed_step;
        }
    }
    for (int d = 0; d < num_decisions; d++) {
        if (env->dg_exec && !env->no_graph) CUDA_TRY(cudaGraphLaunch(env->dg_exec, s));
        else {
            const int rc = decima_decision(env, num_decisions, max_events, traj, s);
            if (rc) return rc;
        }
    }
    if (s != caller) {
        CUDA_TRY(cudaEventRecord(env->ev, s));
        CUDA_TRY(cudaStreamWaitEvent(caller, env->ev, 0));
    }
    SSB_MARK(env, caller);
    return SSB_OK;
}

int ssb_decima_mlp_rows(ssb_env *env, int32_t mlp, const float *x, int32_t n_rows, float *out, void *stream)
{
    if (!env || !env->p.pol_w || !x || !out || n_rows < 1) return SSB_E_INVALID;
    SSB_ON_DEVICE(env);
    cudaStream_t s = (cudaStream_t)stream;
    switch (mlp) {
    case 0: return launch_mlp_rows<tc::ST_PREP>(env, x, n_rows, out, s);
    case 1: return launch_mlp_rows<tc::ST_MSG>(env, x, n_rows, out, s);
    case 2: return launch_mlp_rows<tc::ST_RCV>(env, x, n_rows, out, s);
    case 3: return launch_mlp_rows<tc::ST_DAG>(env, x, n_rows, out, s);
    case 4: return launch_mlp_rows<tc::ST_GLOB>(env, x, n_rows, out, s);
    case 5: return launch_mlp_rows<tc::ST_STAGE>(env, x, n_rows, out, s);
    case 6: return launch_mlp_rows<tc::ST_EXEC>(env, x, n_rows, out, s);
    default: return SSB_E_INVALID;
    }
}

int ssb_decima_work(ssb_env *env, int64_t *out)
{
    if (!env || !out || !env->p.pol_w) return SSB_E_INVALID;
    CUDA_TRY(cudaSetDevice(env->device));
    CUDA_TRY(cudaDeviceSynchronize());
    {
        const int rcp = replan(env, env->own_stream);
        if (rcp) return rcp;
    }
    CUDA_TRY(cudaDeviceSynchronize());
    std::vector<int32_t> c(tc::CNT_TOTAL);
    CUDA_TRY(cudaMemcpy(c.data(), env->p.pl_cnt, sizeof(int32_t) * tc::CNT_TOTAL, cudaMemcpyDeviceToHost));
    int64_t send = 0, recv = 0;
    for (int k = 0; k < tc::MAX_LEVELS; k++) { send += c[tc::CNT_LVL + 2 * k]; recv += c[tc::CNT_LVL + 2 * k + 1]; }
    const int64_t gnn = 16 * 32 + 32 * 16 + 16 * 16;  // one 16 -> 32 -> 16 -> 16 MLP
    out[0] = c[tc::CNT_ALL]; out[1] = c[tc::CNT_SINK]; out[2] = c[tc::CNT_CAND]; out[3] = c[tc::CNT_JOBS];
    out[4] = c[tc::CNT_EXEC]; out[5] = send; out[6] = recv;
    out[7] = out[0] * ((5 * 32 + 32 * 16 + 16 * 16) + (21 * 32 + 32 * 16 + 16 * 16)) + (out[1] + send + recv + out[3]) * gnn +
             out[2] * (53 * 64 + 64 * 64 + 64) + out[4] * (36 * 64 + 64 * 64 + 64);
    return SSB_OK;
}

int ssb_rollout_decima_async(ssb_env *env, int32_t max_decisions, double rollout_duration, uint64_t seed_step,
                             ssb_transition *traj, int32_t *num_steps, double *elapsed, void *stream)
{
    if (!env || !env->p.pol_w || max_decisions < 1 || !(rollout_duration > 0.0) || !traj) return SSB_E_INVALID;
    SSB_ON_DEVICE(env);
    cudaStream_t s = (cudaStream_t)stream;
    const Params &p = env->p;
    const int tb = (p.B + 127) / 128;
    CUDA_TRY(cudaMemsetAsync(p.as_elapsed, 0, sizeof(double) * (size_t)p.B, s));
    CUDA_TRY(cudaMemsetAsync(p.as_rows, 0, sizeof(int32_t) * (size_t)p.B, s));
    CUDA_TRY(cudaMemsetAsync(p.as_fresh, 0, (size_t)p.B, s));  // (a reset round is always followed by its env's decision)
    int any = 1, rounds = 0;
    while (any) {
        // a few rounds per host check of "is any env still inside its rollout"; finished envs cost nothing in the
        // policy (treated as absent) and are masked out of the step
        for (int r = 0; r < 8; r++, rounds++) {
            k_dasync_begin<<<tb, 128, 0, s>>>(p, rollout_duration, max_decisions);
            int rc = decima_policy_impl(env, nullptr, nullptr, p.pol_act_a, p.pol_act_n, true, true, s, p.as_kind);
            if (rc) return rc;
            k_dasync_pre<<<tb, 128, 0, s>>>(p, p.pol_act_a, p.pol_act_n, traj, max_decisions);
            // mask = as_kind (0: untouched); finished episodes are re-seeded with seed + seed_step * reset_count
            if ((rc = step_launch(env, p.pol_act_a, p.pol_act_n, p.as_kind, 0, nullptr, nullptr, 0, s, 1, seed_step))) return rc;
            k_dasync_post<<<tb, 128, 0, s>>>(p, traj, max_decisions, rollout_duration);
            CUDA_TRY(cudaGetLastError());
        }
        CUDA_TRY(cudaMemcpyAsync(&any, p.as_any, sizeof(int), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        if (rounds > 4 * max_decisions + 64) break;  // (every round either writes a row or resets an env)
    }
    if (num_steps) CUDA_TRY(cudaMemcpyAsync(num_steps, p.as_rows, sizeof(int32_t) * (size_t)p.B, cudaMemcpyDeviceToDevice, s));
    if (elapsed) CUDA_TRY(cudaMemcpyAsync(elapsed, p.as_elapsed, sizeof(double) * (size_t)p.B, cudaMemcpyDeviceToDevice, s));
    SSB_MARK(env, s);
    return SSB_OK;
}

int ssb_get_policy_views(ssb_env *env, ssb_policy_views *out)
{
    if (!env || !out || !env->p.pol_w) return SSB_E_INVALID;
    out->stage_logits = env->p.pol_stage_logits;
    out->exec_logits = env->p.pol_exec_logits;
    out->action = env->p.pol_action;
    out->lgprob = env->p.pol_lgprob;
    out->entropy = env->p.pol_entropy;
    out->node_stride = env->p.Sc;
    out->exec_stride = env->p.Epad;
    return SSB_OK;
}

int ssb_get_stats(ssb_env *env, ssb_stats **out)
{
    if (!env || !out) return SSB_E_INVALID;
    *out = env->p.stats;
    return SSB_OK;
}

int ssb_get_debug_counters(ssb_env *env, uint64_t **out)
{
    if (!env || !out) return SSB_E_INVALID;
    *out = reinterpret_cast<uint64_t *>(env->p.prof);
    return SSB_OK;
}

int ssb_collect_stats(ssb_env *env, double *out, void *stream)
{
    if (!env || !out) return SSB_E_INVALID;
    SSB_ON_DEVICE(env);
    k_collect_stats_part<<<STATS_BLOCKS, 256, 0, (cudaStream_t)stream>>>(env->p, env->p.stats_part);
    k_collect_stats_final<<<1, 32, 0, (cudaStream_t)stream>>>(env->p.stats_part, out);
    CUDA_TRY(cudaGetLastError());
    SSB_MARK(env, stream);
    return SSB_OK;
}

int ssb_reset_stats(ssb_env *env, void *stream)
{
    if (!env) return SSB_E_INVALID;
    SSB_ON_DEVICE(env);
    k_zero_stats<<<(env->p.B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(env->p.stats, env->p.B);
    CUDA_TRY(cudaGetLastError());
    SSB_MARK(env, stream);
    return SSB_OK;
}

int ssb_get_jobs(ssb_env *env, int32_t b, int32_t *n_jobs, double *t_arrival, double *t_completed,
                 int32_t *tmpl, uint8_t *state, int32_t capacity)
{
    if (!env || b < 0 || b >= env->p.B || !n_jobs) return SSB_E_INVALID;
    CUDA_TRY(cudaSetDevice(env->device));
    CUDA_TRY(cudaDeviceSynchronize());
    EnvHdr h;
    CUDA_TRY(cudaMemcpy(&h, env->p.hdr + b, sizeof(EnvHdr), cudaMemcpyDeviceToHost));
    *n_jobs = h.n_jobs;
    int n = h.n_jobs < capacity ? h.n_jobs : capacity;
    std::vector<JobRec> jr(n > 0 ? n : 1);
    if (n > 0)
        CUDA_TRY(cudaMemcpy(jr.data(), env->p.job + (size_t)b * env->p.Jc, sizeof(JobRec) * n, cudaMemcpyDeviceToHost));
    for (int j = 0; j < n; j++) {
        if (t_arrival) t_arrival[j] = jr[j].t_arrival;
        if (t_completed) t_completed[j] = jr[j].t_completed;
        if (tmpl) tmpl[j] = jr[j].tmpl;
        if (state) state[j] = jr[j].state;
    }
    return SSB_OK;
}

int ssb_get_history(ssb_env *env, int32_t b, int64_t *n_rows, double *t, int16_t *exec, int16_t *job, int64_t capacity)
{
    if (!env || b < 0 || b >= env->p.B || !n_rows) return SSB_E_INVALID;
    CUDA_TRY(cudaSetDevice(env->device));
    CUDA_TRY(cudaDeviceSynchronize());
    EnvHdr h;
    CUDA_TRY(cudaMemcpy(&h, env->p.hdr + b, sizeof(EnvHdr), cudaMemcpyDeviceToHost));
    *n_rows = h.hist_n;
    int64_t n = std::min<int64_t>(std::min<int64_t>(h.hist_n, env->p.hist_cap), capacity);
    if (n <= 0) return SSB_OK;
    std::vector<HistRow> rows(n);
    CUDA_TRY(cudaMemcpy(rows.data(), env->p.hist + (size_t)b * env->p.hist_cap, sizeof(HistRow) * n,
                        cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < n; i++) {
        if (t) t[i] = rows[i].t;
        if (exec) exec[i] = rows[i].exec;
        if (job) job[i] = rows[i].job;
    }
    return SSB_OK;
}

int ssb_get_log(ssb_env *env, int32_t b, int64_t lo, int64_t hi, int64_t *n_rows, double *t, uint8_t *type,
                int16_t *job, int16_t *stage, int32_t *task, int16_t *exec, double *t_accepted)
{
    if (!env || b < 0 || b >= env->p.B || !n_rows) return SSB_E_INVALID;
    CUDA_TRY(cudaSetDevice(env->device));
    CUDA_TRY(cudaDeviceSynchronize());
    EnvHdr h;
    CUDA_TRY(cudaMemcpy(&h, env->p.hdr + b, sizeof(EnvHdr), cudaMemcpyDeviceToHost));
    *n_rows = h.log_n;
    if (hi <= lo) return SSB_OK;
    if (env->p.log_cap <= 0 || lo < 0 || hi > h.log_n || hi > env->p.log_cap) return SSB_E_INVALID;
    std::vector<LogRow> rows(hi - lo);
    CUDA_TRY(cudaMemcpy(rows.data(), env->p.log + (size_t)b * env->p.log_cap + lo, sizeof(LogRow) * (hi - lo),
                        cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < hi - lo; i++) {
        const LogRow &r = rows[i];
        if (t) t[i] = r.t;
        if (type) type[i] = r.type;
        if (job) job[i] = r.job;
        if (stage) stage[i] = r.stage;
        if (task) task[i] = r.task;
        if (exec) exec[i] = r.exec;
        if (t_accepted) t_accepted[i] = r.t_acc;
    }
    return SSB_OK;
}

}  // extern "C"
