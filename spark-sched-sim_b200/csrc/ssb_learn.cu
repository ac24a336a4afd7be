// ssb_learn.cu -- the learner-side entry points (returns, baselines, PPO loss head, Adam) as their own translation
// unit: they take plain device pointers, never an ssb_env, and their kernels stay out of the unit that holds the
// instruction-supply-bound rollout kernels (see ssb_backward.cuh).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>

#include "../../include/ssb.h"
#include "ssb_learn.cuh"

using namespace ssb;

namespace ssb {
extern thread_local char g_cuda_err[256];
}

#define CUDA_TRY(expr)                                                                   \
    do {                                                                                 \
        cudaError_t e_ = (expr);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", #expr, cudaGetErrorString(e_)); \
            return SSB_E_CUDA;                                                           \
        }                                                                                \
    } while (0)

extern "C" {

int ssb_discounted_returns(const ssb_transition *traj, const int32_t *num_steps, const double *final_wall,
                           int32_t B, int32_t stride, double beta, double *returns, void *stream)
{
    if (!traj || !num_steps || !final_wall || !returns || B < 1 || stride < 1) return SSB_E_INVALID;
    learn::k_discounted_returns<<<(B + 3) / 4, 128, 0, (cudaStream_t)stream>>>(traj, num_steps, final_wall, B, stride,
                                                                              beta, returns);
    CUDA_TRY(cudaGetLastError());
    return SSB_OK;
}

int ssb_differential_returns(const ssb_transition *traj, const int32_t *num_steps, const double *final_wall, int32_t B,
                             int32_t stride, double *window, int32_t cap, int32_t *which, int32_t *scratch,
                             double *avg_num_jobs, double *returns, void *stream)
{
    if (!traj || !num_steps || !final_wall || !window || !which || !scratch || !avg_num_jobs || !returns || B < 1 ||
        stride < 1 || cap < 1 || (*which != 0 && *which != 1))
        return SSB_E_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    int32_t *cnt = scratch, *off = scratch + B;
    const double *src = window + (size_t)*which * cap * 2;
    double *dst = window + (size_t)(1 - *which) * cap * 2;
    learn::k_diff_count<<<(B + 3) / 4, 128, 0, s>>>(traj, num_steps, final_wall, B, stride, cnt);
    learn::k_diff_scan<<<1, 32, 0, s>>>(cnt, B, off);
    learn::k_diff_keep<<<64, 256, 0, s>>>(src, dst, cap, off, B);
    learn::k_diff_fill<<<(B + 3) / 4, 128, 0, s>>>(traj, num_steps, final_wall, B, stride, off, dst, cap);
    learn::k_diff_avg<<<1, 32, 0, s>>>(dst, cap, avg_num_jobs);
    learn::k_diff_returns<<<(B + 127) / 128, 128, 0, s>>>(traj, num_steps, final_wall, B, stride, avg_num_jobs, returns);
    CUDA_TRY(cudaGetLastError());
    *which = 1 - *which;
    return SSB_OK;
}

int ssb_ppo_loss(const float *new_lgprob, const float *old_lgprob, const float *entropy, const double *returns,
                 const double *baselines, const int32_t *idx, int32_t n, float clip_range, float entropy_coeff,
                 double *scratch, float *out, float *grad_lgprob, float *grad_entropy, void *stream)
{
    if (!new_lgprob || !old_lgprob || !entropy || !returns || !baselines || !scratch || !out || n < 1)
        return SSB_E_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    double *part = scratch, *part2 = scratch + 2 * learn::PPO_BLOCKS;
    learn::k_ppo_moments<<<learn::PPO_BLOCKS, learn::PPO_THREADS, 0, s>>>(returns, baselines, idx, n, part);
    learn::k_ppo_terms<<<learn::PPO_BLOCKS, learn::PPO_THREADS, 0, s>>>(new_lgprob, old_lgprob, entropy, returns,
                                                                        baselines, idx, n, clip_range, entropy_coeff,
                                                                        part, part2, grad_lgprob, grad_entropy);
    learn::k_ppo_final<<<1, 32, 0, s>>>(part2, n, entropy_coeff, out);
    CUDA_TRY(cudaGetLastError());
    return SSB_OK;
}

int ssb_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int32_t n, int32_t step, float lr,
                  float beta1, float beta2, float eps, float max_grad_norm, double *scratch, float *grad_norm_out,
                  void *stream)
{
    if (!param || !grad || !exp_avg || !exp_avg_sq || !scratch || n < 1 || step < 1) return SSB_E_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    learn::k_grad_sqsum<<<learn::PPO_BLOCKS, learn::PPO_THREADS, 0, s>>>(grad, n, scratch);
    const int blocks = std::min(learn::PPO_BLOCKS, (n + learn::PPO_THREADS - 1) / learn::PPO_THREADS);
    learn::k_adam_step<<<blocks, learn::PPO_THREADS, 0, s>>>(param, grad, exp_avg, exp_avg_sq, n, step, lr, beta1, beta2,
                                                             eps, max_grad_norm, scratch, grad_norm_out);
    CUDA_TRY(cudaGetLastError());
    return SSB_OK;
}

int ssb_group_baselines(const ssb_transition *traj, const double *returns, const int32_t *num_steps, int32_t B,
                        int32_t stride, int32_t group_size, double *baselines, void *stream)
{
    if (!traj || !returns || !num_steps || !baselines || B < 1 || stride < 1 || group_size < 1 || group_size > 128)
        return SSB_E_INVALID;
    learn::k_group_baselines<<<(B + 3) / 4, 128, 0, (cudaStream_t)stream>>>(traj, returns, num_steps, B, stride,
                                                                           group_size, baselines);
    CUDA_TRY(cudaGetLastError());
    return SSB_OK;
}

}  // extern "C"
