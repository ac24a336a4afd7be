// ssb_env.cuh -- what the translation units of libssb share: the handle, status / ordering macros and the internal
// entry points that cross units.  The simulator kernels (ssb_api.cu) and the Decima policy (ssb_policy.cu) are separate
// units ON PURPOSE: the fused rollout kernel is bound by instruction supply, and what else is instantiated in its unit
// moves its code around -- the policy kernels next to it cost the headline 5 % (16.8 M instead of 17.7 M decisions/s,
// profiles/r02_bench_n1_before_split.json); round 1 saw the same with the backward kernels.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "ssb_types.cuh"

namespace ssb {
extern thread_local char g_cuda_err[256];  // ssb_last_cuda_error()
}

#ifndef SSB_WARPS_PER_CTA
#define SSB_WARPS_PER_CTA 4
#endif
constexpr int WARPS_PER_CTA = SSB_WARPS_PER_CTA;

#define CUDA_TRY(expr)                                                                   \
    do {                                                                                 \
        cudaError_t e_ = (expr);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            snprintf(ssb::g_cuda_err, sizeof(ssb::g_cuda_err), "%s: %s", #expr, cudaGetErrorString(e_)); \
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


// How ssb_decima_policy runs: row lists + one launch per MLP pass with the TMEM-resident bf16 three-term tiles (the
// default for large batches), the whole decision of a group of envs in one persistent kernel (one launch: small
// batches, e.g. the single-env facade), or round 1's shared-memory tf32 tiles (kept for A/B measurements).
enum { POLICY_TILES = 0, POLICY_FUSED = 1, POLICY_TILES_TF32 = 2 };

struct ssb_env {
    ssb_config cfg;
    Dims dims;
    ssb::Params p;
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
    void *bw_scratch;   // ssb_decima_attach_backward_scratch
    int bw_saved;       // the last policy evaluation left its levels' input rows in bw_scratch
    int fused_group;    // environments per group of the fused policy kernel
    int snap_loaded;    // ssb_decima_snapshot_load: a stored observation is in place, the live one parked
    int auto_reset;     // ssb_set_autoreset
    uint64_t auto_seed_step;
};

// ---- internal entry points across the units (C linkage like the public ones; not exported through include/ssb.h)
extern "C" {
// ssb_api.cu
int ssb_i_step_launch(ssb_env *env, const int32_t *stage_idx, const int32_t *num_exec, const uint8_t *mask,
                      int32_t max_events, int32_t *next_a, int32_t *next_n, int dyn, cudaStream_t s,
                      int force_autoreset, uint64_t force_seed_step);
int ssb_i_decima_obs(ssb_env *env, cudaStream_t s);  // the observation adapter kernel
int ssb_i_decima_obs_cta(ssb_env *env, cudaStream_t s);  // its one-CTA-per-env version (ssb_policy.cu)
// ssb_policy.cu
void ssb_i_policy_carve(Carver &cv, const ssb_config &c, const Dims &d, ssb::Params &p);
int ssb_i_policy_init(ssb_env *env);
void ssb_i_drop_decision_graph(ssb_env *env);
}
