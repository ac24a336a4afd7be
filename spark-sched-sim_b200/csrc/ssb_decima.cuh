// ssb_decima.cuh -- Decima policy forward pass on the device, one warp per environment.
//
// Restates DecimaScheduler.schedule (schedulers/decima/scheduler.py:71-99): DAG-GNN encoder
// (NodeEncoder :173-241, DagEncoder :244-257, GlobalEncoder :260-276), stage score network
// (:279-320), executor-count score network (:323-385) and utils.sample (decima/utils.py:19-23),
// on the observation that k_decima_obs / decima_obs_w just wrote.  float32 throughout (fmaf
// accumulation starting from the bias), lanes over nodes / jobs / candidate actions; every lane
// evaluates a whole 3-layer MLP for its item with the weights broadcast from L1.
//
// NodeEncoder is NOT textbook level-synchronous message passing: for each
// level mask, from the deepest to the first, every head (child) of a masked edge sends
// mlp_msg(h), and every tail (parent) OVERWRITES h = h_init + mlp_update(sum of its masked
// children's messages).  The loop below is that literal sequence.
#pragma once
#include "ssb_sim.cuh"

namespace ssb {

// ABI layout (dw): the tensors of models/decima/model.pt in state_dict order, each [out][in] row-major
// (42 tensors, 20 802 floats).  Device layout (dd): per layer the TRANSPOSED weight [in][out] followed
// by the bias, every tensor padded to a multiple of 4 floats so that a lane fetches the weights of four
// consecutive output neurons with one 128-bit load (ssb_set_decima_weights does the re-layout).
namespace dw {
constexpr int mlp(int in, int h1, int h2, int out) { return h1 * in + h1 + h2 * h1 + h2 + out * h2 + out; }
constexpr int TOTAL = mlp(5, 32, 16, 16) + 2 * mlp(16, 32, 16, 16) + mlp(21, 32, 16, 16) + mlp(16, 32, 16, 16) +
                      mlp(53, 64, 64, 1) + mlp(36, 64, 64, 1);
static_assert(TOTAL == 20802, "Decima parameter count");
}  // namespace dw
namespace dd {
__host__ __device__ constexpr int pad4(int n) { return (n + 3) & ~3; }
__host__ __device__ constexpr int layer(int in, int out) { return pad4(in * out) + pad4(out); }
__host__ __device__ constexpr int mlp(int in, int h1, int h2, int out)
{
    return layer(in, h1) + layer(h1, h2) + layer(h2, out);
}
constexpr int PREP = 0;
constexpr int MSG = PREP + mlp(5, 32, 16, 16);
constexpr int UPD = MSG + mlp(16, 32, 16, 16);
constexpr int DAG = UPD + mlp(16, 32, 16, 16);
constexpr int GLOB = DAG + mlp(21, 32, 16, 16);
constexpr int STAGE = GLOB + mlp(16, 32, 16, 16);
constexpr int EXEC = STAGE + mlp(53, 64, 64, 1);
constexpr int TOTAL = EXEC + mlp(36, 64, 64, 1);
}  // namespace dd

template <bool TANH>
__device__ __forceinline__ float act(float x)
{
    if (TANH) return tanhf(x);
    return x > 0.0f ? x : 0.2f * x;  // LeakyReLU(negative_slope=0.2), config/decima_tpch.yaml:69-73
}

// One Linear (+ activation): out[o] = act(bias[o] + sum_i W[o][i] * in[i]), accumulated in input order
// with fmaf.  16 output accumulators live in registers at a time; Wt is the transposed weight.
template <int IN, int OUT, bool ACT, bool TANH>
__device__ __forceinline__ void dense(const float *__restrict__ Wt, const float *in, float *out)
{
    const float *bias = Wt + dd::pad4(IN * OUT);
    if (OUT == 1) {
        float s = __ldg(bias);
#pragma unroll 4
        for (int i = 0; i < IN; i++) s = fmaf(__ldg(Wt + i), in[i], s);
        out[0] = s;
        return;
    }
#pragma unroll 1
    for (int oc = 0; oc < OUT; oc += 16) {
        float acc[16];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const float4 bv = __ldg(reinterpret_cast<const float4 *>(bias + oc) + q);
            acc[4 * q] = bv.x; acc[4 * q + 1] = bv.y; acc[4 * q + 2] = bv.z; acc[4 * q + 3] = bv.w;
        }
#pragma unroll 2
        for (int i = 0; i < IN; i++) {
            const float v = in[i];
            const float4 *wp = reinterpret_cast<const float4 *>(Wt + i * OUT + oc);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const float4 wv = __ldg(wp + q);
                acc[4 * q] = fmaf(wv.x, v, acc[4 * q]);
                acc[4 * q + 1] = fmaf(wv.y, v, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(wv.z, v, acc[4 * q + 2]);
                acc[4 * q + 3] = fmaf(wv.w, v, acc[4 * q + 3]);
            }
        }
#pragma unroll
        for (int q = 0; q < 16; q++) out[oc + q] = ACT ? act<TANH>(acc[q]) : acc[q];
    }
}

// make_mlp(IN, [H1, H2], OUT): Linear / act / Linear / act / Linear (decima/utils.py:45-64)
template <int IN, int H1, int H2, int OUT, bool TANH>
__device__ __forceinline__ void mlp3(const float *__restrict__ w, const float (&in)[IN], float (&out)[OUT])
{
    float a1[H1], a2[H2];
    dense<IN, H1, true, TANH>(w, in, a1);
    dense<H1, H2, true, TANH>(w + dd::layer(IN, H1), a1, a2);
    dense<H2, OUT, false, TANH>(w + dd::layer(IN, H1) + dd::layer(H1, H2), a2, out);
}

__device__ __forceinline__ void ld16(const float *p, float (&v)[16])
{
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float4 q = reinterpret_cast<const float4 *>(p)[i];
        v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
    }
}
__device__ __forceinline__ void st16(float *p, const float (&v)[16])
{
#pragma unroll
    for (int i = 0; i < 4; i++) reinterpret_cast<float4 *>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}

struct PolicyBufs {  // per-environment slices of the policy scratch / output arrays
    float *h_init, *h, *msg;   // [Sc][16]
    float *h_dag, *g;          // [Jc][16]
    float *h_glob;             // [16]
    int32_t *row_start;        // [Sc] first observation edge whose tail is this node
    uint8_t *flag;             // [3][Sc] tail-of-any-edge / sender / receiver flags
    float *stage_logits;       // [Sc]
    float *exec_logits;        // [E]
    int32_t *action;           // [4]: stage_idx, job_idx, num_exec, number of stage candidates
    float *lgprob;             // [1]
};

// softmax-sample (or take `forced` if >= 0) over logits[0..n) ; returns the index, adds log prob
__device__ inline int sample_w(const float *logits, int n, int forced, float u, int lane, float &lgprob)
{
    float mx = -INFINITY;
    for (int i = lane; i < n; i += 32) mx = fmaxf(mx, logits[i]);
    for (int off = 16; off; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, off));
    float sum = 0.0f;
    for (int i = lane; i < n; i += 32) sum += expf(logits[i] - mx);
    for (int off = 16; off; off >>= 1) sum += __shfl_xor_sync(FULL, sum, off);
    int idx = forced;
    if (idx < 0) {  // inverse-CDF draw, sequential over the candidates (lane 0)
        idx = n - 1;
        if (lane == 0) {
            float acc = 0.0f, target = u * sum;
            for (int i = 0; i < n; i++) {
                acc += expf(logits[i] - mx);
                if (acc > target) { idx = i; break; }
            }
        }
        idx = __shfl_sync(FULL, idx, 0);
    }
    if (idx >= 0 && idx < n) lgprob += logf(expf(logits[idx] - mx) / sum);
    return idx;
}

// Must be called by all 32 lanes right after sim.decima_obs_w(); `w` = flat weights.
__device__ inline void decima_policy_w(Sim &sim, const float *__restrict__ w, const PolicyBufs &pb,
                                       int forced_stage, int forced_num_exec)
{
    const Params &p = sim.p;
    const int lane = sim.lane, b = sim.b;
    const ssb_obs_hdr &oh = *sim.oh;
    const int N = oh.num_nodes, M = oh.num_edges, Ja = oh.num_active_jobs;
    const float *x = p.dec_feat + (size_t)b * p.Sc * 5;
    const int32_t *edges = p.obs_edges + (size_t)b * p.Mc * 2;
    const uint64_t *ebits = p.dec_edge_bits + (size_t)b * p.Mc;
    const int32_t *dag_ptr = p.obs_dag_ptr + (size_t)b * (p.Jc + 1);
    const uint8_t *smask = p.dec_stage_mask + (size_t)b * p.Sc;
    const int32_t *caps = p.dec_caps + (size_t)b * p.Jc;
    const int depth = p.dec_depth[b];
    uint8_t *is_tail = pb.flag, *snd = pb.flag + p.Sc, *rcv = pb.flag + 2 * p.Sc;

    // h_init = mlp_prep(x)  (:201)
    for (int n0 = 0; n0 < N; n0 += 32) {
        const int n = n0 + lane;
        if (n < N) {
            float in[5], out[16];
#pragma unroll
            for (int i = 0; i < 5; i++) in[i] = x[n * 5 + i];
            mlp3<5, 32, 16, 16, false>(w + dd::PREP, in, out);
            st16(pb.h_init + (size_t)n * 16, out);
            if (depth == 0) st16(pb.h + (size_t)n * 16, out);  // _forward_no_mp (:236-241)
            is_tail[n] = 0;
        }
    }
    __syncwarp();
    if (depth > 0) {
        for (int e0 = 0; e0 < M; e0 += 32) {
            const int e = e0 + lane;
            if (e < M) {
                const int u = edges[2 * e];
                is_tail[u] = 1;
                if (e == 0 || edges[2 * (e - 1)] != u) pb.row_start[u] = e;  // edges are sorted by tail
            }
        }
        __syncwarp();
        // nodes that are the tail of no edge: h = mlp_update(h_init); all others start at 0 (:204-212)
        for (int n0 = 0; n0 < N; n0 += 32) {
            const int n = n0 + lane;
            if (n < N) {
                float out[16];
                if (!is_tail[n]) {
                    float in[16];
                    ld16(pb.h_init + (size_t)n * 16, in);
                    mlp3<16, 32, 16, 16, false>(w + dd::UPD, in, out);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; i++) out[i] = 0.0f;
                }
                st16(pb.h + (size_t)n * 16, out);
            }
        }
        __syncwarp();
        for (int k = depth - 1; k >= 0; k--) {  // reversed(edge_masks) (:214-232)
            for (int n = lane; n < N; n += 32) { snd[n] = 0; rcv[n] = 0; }
            __syncwarp();
            for (int e = lane; e < M; e += 32) {
                if ((ebits[e] >> k) & 1) { rcv[edges[2 * e]] = 1; snd[edges[2 * e + 1]] = 1; }
            }
            __syncwarp();
            for (int n0 = 0; n0 < N; n0 += 32) {  // msg[src_mask] = mlp_msg(h[src_mask])
                const int n = n0 + lane;
                if (n < N && snd[n]) {
                    float in[16], out[16];
                    ld16(pb.h + (size_t)n * 16, in);
                    mlp3<16, 32, 16, 16, false>(w + dd::MSG, in, out);
                    st16(pb.msg + (size_t)n * 16, out);
                }
            }
            __syncwarp();
            for (int n0 = 0; n0 < N; n0 += 32) {  // h[dst_mask] = h_init + mlp_update(adj @ msg)
                const int n = n0 + lane;
                if (n < N && rcv[n]) {
                    float agg[16], out[16], hi[16];
#pragma unroll
                    for (int i = 0; i < 16; i++) agg[i] = 0.0f;
                    for (int e = pb.row_start[n]; e < M && edges[2 * e] == n; e++) {
                        if ((ebits[e] >> k) & 1) {
                            float m[16];
                            ld16(pb.msg + (size_t)edges[2 * e + 1] * 16, m);
#pragma unroll
                            for (int i = 0; i < 16; i++) agg[i] += m[i];
                        }
                    }
                    mlp3<16, 32, 16, 16, false>(w + dd::UPD, agg, out);
                    ld16(pb.h_init + (size_t)n * 16, hi);
#pragma unroll
                    for (int i = 0; i < 16; i++) out[i] = hi[i] + out[i];
                    st16(pb.h + (size_t)n * 16, out);
                }
            }
            __syncwarp();
        }
    }
    // DagEncoder: h_dag[j] = sum over the job's nodes of mlp([x, h])  (:252-257)
    for (int n0 = 0; n0 < N; n0 += 32) {
        const int n = n0 + lane;
        if (n < N) {
            float in[21], out[16], hv[16];
#pragma unroll
            for (int i = 0; i < 5; i++) in[i] = x[n * 5 + i];
            ld16(pb.h + (size_t)n * 16, hv);
#pragma unroll
            for (int i = 0; i < 16; i++) in[5 + i] = hv[i];
            mlp3<21, 32, 16, 16, false>(w + dd::DAG, in, out);
            st16(pb.msg + (size_t)n * 16, out);  // msg buffer reused for the per-node dag terms
        }
    }
    __syncwarp();
    for (int j0 = 0; j0 < Ja; j0 += 32) {
        const int j = j0 + lane;
        if (j < Ja) {
            float s[16], gj[16];
#pragma unroll
            for (int i = 0; i < 16; i++) s[i] = 0.0f;
            for (int n = dag_ptr[j]; n < dag_ptr[j + 1]; n++) {
                float z[16];
                ld16(pb.msg + (size_t)n * 16, z);
#pragma unroll
                for (int i = 0; i < 16; i++) s[i] += z[i];
            }
            st16(pb.h_dag + (size_t)j * 16, s);
            mlp3<16, 32, 16, 16, false>(w + dd::GLOB, s, gj);  // GlobalEncoder (:265-276)
            st16(pb.g + (size_t)j * 16, gj);
        }
    }
    __syncwarp();
    if (lane < 16) {
        float s = 0.0f;
        for (int j = 0; j < Ja; j++) s += pb.g[(size_t)j * 16 + lane];
        pb.h_glob[lane] = s;
    }
    __syncwarp();
    float hg[16];
    ld16(pb.h_glob, hg);
    // StagePolicyNetwork: scores of the schedulable nodes, in node order (:293-320)
    int n_cand = 0;
    for (int n0 = 0; n0 < N; n0 += 32) {
        const int n = n0 + lane;
        const bool cand = n < N && smask[n];
        const unsigned bm = __ballot_sync(FULL, cand);
        if (cand) {
            int lo = 0, hi = Ja;  // job of node n: last j with dag_ptr[j] <= n
            while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (dag_ptr[mid] <= n) lo = mid; else hi = mid; }
            float in[53], out[1], t[16];
#pragma unroll
            for (int i = 0; i < 5; i++) in[i] = x[n * 5 + i];
            ld16(pb.h + (size_t)n * 16, t);
#pragma unroll
            for (int i = 0; i < 16; i++) in[5 + i] = t[i];
            ld16(pb.h_dag + (size_t)lo * 16, t);
#pragma unroll
            for (int i = 0; i < 16; i++) in[21 + i] = t[i];
#pragma unroll
            for (int i = 0; i < 16; i++) in[37 + i] = hg[i];
            mlp3<53, 64, 64, 1, true>(w + dd::STAGE, in, out);
            pb.stage_logits[n_cand + __popc(bm & ((1u << lane) - 1))] = out[0];
        }
        n_cand += __popc(bm);
    }
    __syncwarp();
    // draws: Philox policy stream ctr = (policy draw index, 0, 4, 0)
    const uint32_t pd = sim.h->policy_draws;
    const uint4 rw = philox4x32_10(pd, 0u, 4u, 0u, (uint32_t)sim.h->seed, (uint32_t)(sim.h->seed >> 32));
    const float u1 = ((float)(rw.x >> 8) + 0.5f) * (1.0f / 16777216.0f), u2 = ((float)(rw.y >> 8) + 0.5f) * (1.0f / 16777216.0f);
    float lgprob = 0.0f;
    const int stage_idx = n_cand > 0 ? sample_w(pb.stage_logits, n_cand, forced_stage, u1, lane, lgprob) : -1;
    // job of the chosen stage: the stage_idx-th schedulable node (scheduler.py:86-88)
    int job_idx = -1, num_exec = 0, cap = 0;
    if (stage_idx >= 0 && stage_idx < n_cand) {
        int seen = 0, node = -1;
        for (int n0 = 0; n0 < N && node < 0; n0 += 32) {
            const int n = n0 + lane;
            const unsigned bm = __ballot_sync(FULL, n < N && smask[n]);
            const int c = __popc(bm);
            if (stage_idx < seen + c) {
                unsigned m = bm;
                for (int q = stage_idx - seen; q > 0; q--) m &= m - 1;
                node = n0 + __ffs(m) - 1;
            }
            seen += c;
        }
        int lo = 0, hi = Ja;
        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (dag_ptr[mid] <= node) lo = mid; else hi = mid; }
        job_idx = lo;
        cap = caps[job_idx];
        // ExecPolicyNetwork: scores of num_exec = 0 .. cap-1 for that job (:338-385)
        for (int c0 = 0; c0 < cap; c0 += 32) {
            const int c = c0 + lane;
            if (c < cap) {
                float in[36], out[1], t[16];
                const int first = dag_ptr[job_idx];
#pragma unroll
                for (int i = 0; i < 3; i++) in[i] = x[first * 5 + i];
                ld16(pb.h_dag + (size_t)job_idx * 16, t);
#pragma unroll
                for (int i = 0; i < 16; i++) in[3 + i] = t[i];
#pragma unroll
                for (int i = 0; i < 16; i++) in[19 + i] = hg[i];
                in[35] = __fdiv_rn((float)c, (float)p.E);  // torch.arange(E) / E in float32 (:380)
                mlp3<36, 64, 64, 1, true>(w + dd::EXEC, in, out);
                pb.exec_logits[c] = out[0];
            }
        }
        __syncwarp();
        num_exec = cap > 0 ? sample_w(pb.exec_logits, cap, forced_num_exec, u2, lane, lgprob) : 0;
    }
    if (lane == 0) {
        pb.action[0] = stage_idx; pb.action[1] = job_idx; pb.action[2] = num_exec; pb.action[3] = n_cand;
        pb.lgprob[0] = lgprob;
        sim.h->policy_draws = pd + 1;
    }
    __syncwarp();
}

}  // namespace ssb
