// ssb_learn.cuh -- what the trainer computes from the rollout buffers before the policy update
// (trainers/trainer.py:172-212 `_preprocess_rollouts`), on the device buffers ssb_rollout_fair_traj fills:
//   * continuously discounted returns   (trainers/utils/returns_calculator.py:67-76)
//   * the interpolated group baseline    (trainers/utils/baselines.py:12-37)
// f64 throughout, the same operation order as the numpy code (no FMA: the library is built with -fmad=false).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ssb.h"

namespace ssb {
namespace learn {

// wall time of step k of env b (k == n: the time after the last step)
__device__ __forceinline__ double step_time(const ssb_transition *traj, const double *final_wall, int b, int stride,
                                            int n, int k)
{
    return k < n ? traj[(size_t)b * stride + k].wall_time : final_wall[b];
}

// R_k = r_k + exp(-beta * 1e-3 * dt_k) * R_{k+1}, backwards, R_n = 0 (returns_calculator.py:67-76).
// One thread per rollout (a serial recurrence); the exponentials of a chunk are computed by the whole warp
// first so that the chain itself is one multiply-add per step.
__global__ void __launch_bounds__(128)
k_discounted_returns(const ssb_transition *traj, const int32_t *num_steps, const double *final_wall, int B, int stride,
                     double beta, double *returns)
{
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= B) return;
    const int n = min(num_steps[b], stride);
    const double nb = -beta * 1e-3;  // `-self.beta * 1e-3 * dt` evaluates left to right
    double R = 0.0;
    for (int hi = n; hi > 0; hi -= 32) {
        const int k = hi - 1 - lane;  // lane 0 holds the latest step of the chunk
        double g = 0.0, r = 0.0;
        if (k >= 0) {
            const double dt = step_time(traj, final_wall, b, stride, n, k + 1) - step_time(traj, final_wall, b, stride, n, k);
            g = exp(nb * dt);
            r = traj[(size_t)b * stride + k].reward;
        }
        const int m = min(32, hi);
        double mine = 0.0;
        for (int i = 0; i < m; i++) {
            const double gi = __shfl_sync(0xffffffffu, g, i), ri = __shfl_sync(0xffffffffu, r, i);
            R = __dadd_rn(ri, __dmul_rn(gi, R));
            if (lane == i) mine = R;
        }
        if (k >= 0) returns[(size_t)b * stride + k] = mine;
    }
}

// np.interp(x, xp, fp) for one x (numpy/core/src/multiarray/compiled_base.c:arr_interp): clamps outside the
// range; j = the LAST index with xp[j] <= x (times repeat: same-round decisions share a wall time);
// x == xp[j] returns fp[j]; otherwise slope * (x - xp[j]) + fp[j].
__device__ inline double interp1(double x, const ssb_transition *tr, const double *fp, int n)
{
    if (n <= 0) return 0.0;
    if (x < tr[0].wall_time) return fp[0];
    if (x >= tr[n - 1].wall_time) return fp[n - 1];  // x > last: right value; x == last: fp[n - 1]
    int lo = 0, hi = n - 1;  // invariant: xp[lo] <= x < xp[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (tr[mid].wall_time <= x) lo = mid; else hi = mid;
    }
    const double x0 = tr[lo].wall_time, y0 = fp[lo];
    if (x0 == x) return y0;
    const double slope = __ddiv_rn(__dadd_rn(fp[lo + 1], -y0), __dadd_rn(tr[lo + 1].wall_time, -x0));
    return __dadd_rn(__dmul_rn(slope, __dadd_rn(x, -x0)), y0);
}

// Baseline.average (baselines.py:12-37): envs [g * R, (g + 1) * R) ran the same job sequence; the baseline of
// env i at its step time t is the mean over the group of every member's returns interpolated at t.  The mean is
// numpy's pairwise sum (sequential below 8 terms, 8 accumulators up to 128) divided by R.
__global__ void __launch_bounds__(128)
k_group_baselines(const ssb_transition *traj, const double *returns, const int32_t *num_steps, int B, int stride,
                  int R, double *baselines)
{
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= B) return;
    const int g0 = (b / R) * R, n = min(num_steps[b], stride);
    for (int k = lane; k < n; k += 32) {
        const double t = traj[(size_t)b * stride + k].wall_time;
        double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, res = 0.0;
        const int full = R < 8 ? 0 : R - (R % 8);
        for (int j = 0; j < R; j++) {
            const int e = g0 + j;
            const double y = e < B ? interp1(t, traj + (size_t)e * stride, returns + (size_t)e * stride,
                                             min(num_steps[e], stride)) : 0.0;
            if (j < full) acc[j & 7] = j < 8 ? y : __dadd_rn(acc[j & 7], y);
            else {
                if (j == full && full)
                    res = __dadd_rn(__dadd_rn(__dadd_rn(acc[0], acc[1]), __dadd_rn(acc[2], acc[3])),
                                    __dadd_rn(__dadd_rn(acc[4], acc[5]), __dadd_rn(acc[6], acc[7])));
                res = __dadd_rn(res, y);
            }
        }
        if (full == R)
            res = __dadd_rn(__dadd_rn(__dadd_rn(acc[0], acc[1]), __dadd_rn(acc[2], acc[3])),
                            __dadd_rn(__dadd_rn(acc[4], acc[5]), __dadd_rn(acc[6], acc[7])));
        baselines[(size_t)b * stride + k] = __ddiv_rn(res, (double)R);
    }
}

}  // namespace learn
}  // namespace ssb

namespace ssb {
namespace learn {

// ---- differential returns (returns_calculator.py:52-65, :78-89) --------------------------------------------------
// avg_num_jobs = (total job time) / (total time) over a window of the latest `cap` steps with dt > 0, taken over all
// rollouts in order (CircularArray.extend); then R_k = -(job_time_k - dt_k * avg_num_jobs) + R_{k+1}.

// per rollout: number of steps with a positive duration
__global__ void __launch_bounds__(128)
k_diff_count(const ssb_transition *traj, const int32_t *num_steps, const double *final_wall, int B, int stride, int32_t *cnt)
{
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= B) return;
    const int n = min(num_steps[b], stride);
    int c = 0;
    for (int k = lane; k < n; k += 32)
        c += (step_time(traj, final_wall, b, stride, n, k + 1) - step_time(traj, final_wall, b, stride, n, k)) > 0.0;
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0) cnt[b] = c;
}
// exclusive prefix of the counts (one thread: B is a few thousand) -> off[0..B], off[B] = total
__global__ void k_diff_scan(const int32_t *cnt, int B, int32_t *off)
{
    if (threadIdx.x || blockIdx.x) return;
    int run = 0;
    for (int b = 0; b < B; b++) { off[b] = run; run += cnt[b]; }
    off[B] = run;
}
// CircularArray.extend into `dst` (the other half of a ping-pong pair): the kept tail of `src`, then the new rows
__global__ void __launch_bounds__(256)
k_diff_keep(const double *src, double *dst, int cap, const int32_t *off, int B)
{
    const int total = off[B], num_new = total < cap ? total : cap, num_keep = cap - num_new;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < num_keep; i += gridDim.x * blockDim.x) {
        dst[2 * i] = src[2 * (i + num_new)];
        dst[2 * i + 1] = src[2 * (i + num_new) + 1];
    }
}
__global__ void __launch_bounds__(128)
k_diff_fill(const ssb_transition *traj, const int32_t *num_steps, const double *final_wall, int B, int stride,
            const int32_t *off, double *dst, int cap)
{
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= B) return;
    const int total = off[B], num_new = total < cap ? total : cap, num_keep = cap - num_new, first = total - num_new;
    const int n = min(num_steps[b], stride);
    int run = off[b];  // global index of this rollout's next filtered row
    for (int k0 = 0; k0 < n; k0 += 32) {
        const int k = k0 + lane;
        double dt = 0.0;
        if (k < n) dt = step_time(traj, final_wall, b, stride, n, k + 1) - step_time(traj, final_wall, b, stride, n, k);
        const bool keep = k < n && dt > 0.0;
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int g = run + __popc(m & ((1u << lane) - 1u));
            if (g >= first) {
                dst[2 * (size_t)(num_keep + g - first)] = dt;
                dst[2 * (size_t)(num_keep + g - first) + 1] = traj[(size_t)b * stride + k].reward;
            }
        }
        run += __popc(m);
    }
}
// total_time, rew_sum = buff.data.sum(0): numpy adds the rows one after the other; avg = -rew_sum / total_time
__global__ void k_diff_avg(const double *buf, int cap, double *avg_num_jobs)
{
    __shared__ double col[2];
    if (threadIdx.x < 2) {
        double s = 0.0;
        for (int i = 0; i < cap; i++) s = __dadd_rn(s, buf[2 * (size_t)i + threadIdx.x]);
        col[threadIdx.x] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) *avg_num_jobs = __ddiv_rn(-col[1], col[0]);
}
__global__ void __launch_bounds__(128)
k_diff_returns(const ssb_transition *traj, const int32_t *num_steps, const double *final_wall, int B, int stride,
               const double *avg_num_jobs, double *returns)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int n = min(num_steps[b], stride);
    const double avg = *avg_num_jobs;
    double R = 0.0;
    for (int k = n - 1; k >= 0; k--) {
        const double dt = step_time(traj, final_wall, b, stride, n, k + 1) - step_time(traj, final_wall, b, stride, n, k);
        const double job_time = -traj[(size_t)b * stride + k].reward;
        const double expected = __dmul_rn(dt, avg);
        R = __dadd_rn(-__dadd_rn(job_time, -expected), R);  // R = -(job_time - expected_job_time) + R
        returns[(size_t)b * stride + k] = R;
    }
}

}  // namespace learn
}  // namespace ssb

namespace ssb {
namespace learn {

// ---- PPO clip loss head (trainers/ppo.py:104-140 `_compute_loss`) ------------------------------------------------
// Forward values and the adjoint seeds of the backward pass (d loss / d lgprob_i, d loss / d entropy_i) for one
// mini-batch of samples.  Sums are taken in f64 in a fixed order (block partition + tree), results rounded to f32.
constexpr int PPO_BLOCKS = 128, PPO_THREADS = 256;

__device__ __forceinline__ double block_sum(double v, double *sh)
{
    for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int i = 0; i < PPO_THREADS / 32; i++) s += sh[i];
    return s;
}
__device__ __forceinline__ float ppo_advantage(const double *returns, const double *baselines, int i)
{
    return (float)(returns[i] - baselines[i]);  // `returns - baselines` in f64, then torch.tensor(...).float()
}
// part[k] = { sum a, sum a^2 } of block k's samples
__global__ void __launch_bounds__(PPO_THREADS)
k_ppo_moments(const double *returns, const double *baselines, const int32_t *idx, int n, double *part)
{
    __shared__ double sh[PPO_THREADS / 32];
    double s1 = 0.0, s2 = 0.0;
    for (int i = blockIdx.x * PPO_THREADS + threadIdx.x; i < n; i += PPO_BLOCKS * PPO_THREADS) {
        const double a = ppo_advantage(returns, baselines, idx ? idx[i] : i);
        s1 += a; s2 += a * a;
    }
    s1 = block_sum(s1, sh);
    s2 = block_sum(s2, sh);
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = s1; part[2 * blockIdx.x + 1] = s2; }
}
// per-sample terms; part2[k] = { sum min(pl1, pl2), sum entropy, sum (ratio - 1 - log_ratio) }
__global__ void __launch_bounds__(PPO_THREADS)
k_ppo_terms(const float *new_lgprob, const float *old_lgprob, const float *entropy, const double *returns,
            const double *baselines, const int32_t *idx, int n, float clip_range, float entropy_coeff,
            const double *part, double *part2, float *grad_lgprob, float *grad_entropy)
{
    __shared__ double sh[PPO_THREADS / 32];
    double s1 = 0.0, s2 = 0.0;
    for (int k = 0; k < PPO_BLOCKS; k++) { s1 += part[2 * k]; s2 += part[2 * k + 1]; }
    const double mean = s1 / n;
    // torch.std: the unbiased estimator; (advgs - mean) / (std + 1e-8) in f32 (ppo.py:116-117)
    const double var = n > 1 ? fmax(s2 - n * mean * mean, 0.0) / (n - 1) : nan("");
    const float meanf = (float)mean, denom = (float)sqrt(var) + 1e-8f;
    const float lo = 1.0f - clip_range, hi = 1.0f + clip_range, inv_n = 1.0f / (float)n;
    double t_min = 0.0, t_ent = 0.0, t_kl = 0.0;
    for (int i = blockIdx.x * PPO_THREADS + threadIdx.x; i < n; i += PPO_BLOCKS * PPO_THREADS) {
        const int r = idx ? idx[i] : i;
        const float adv = (ppo_advantage(returns, baselines, r) - meanf) / denom;
        const float log_ratio = new_lgprob[r] - old_lgprob[r];
        const float ratio = expf(log_ratio);
        const float pl1 = adv * ratio, pl2 = adv * fminf(fmaxf(ratio, lo), hi);
        t_min += fminf(pl1, pl2);
        t_ent += entropy[r];
        t_kl += (ratio - 1.0f) - log_ratio;
        // d(-mean(min(pl1, pl2))) / d lgprob: through pl1 always when it is the smaller one, through both (the clamp
        // passes the gradient) while the ratio is inside the clip range
        const bool inside = ratio >= lo && ratio <= hi;
        if (grad_lgprob) grad_lgprob[i] = (inside || pl1 < pl2) ? -inv_n * adv * ratio : 0.0f;
        if (grad_entropy) grad_entropy[i] = -entropy_coeff * inv_n;
    }
    t_min = block_sum(t_min, sh);
    t_ent = block_sum(t_ent, sh);
    t_kl = block_sum(t_kl, sh);
    if (threadIdx.x == 0) {
        part2[3 * blockIdx.x] = t_min; part2[3 * blockIdx.x + 1] = t_ent; part2[3 * blockIdx.x + 2] = t_kl;
    }
}
// out = { loss, policy_loss, entropy_loss, approx_kl_div }
__global__ void k_ppo_final(const double *part2, int n, float entropy_coeff, float *out)
{
    if (threadIdx.x || blockIdx.x) return;
    double t_min = 0.0, t_ent = 0.0, t_kl = 0.0;
    for (int k = 0; k < PPO_BLOCKS; k++) { t_min += part2[3 * k]; t_ent += part2[3 * k + 1]; t_kl += part2[3 * k + 2]; }
    const float policy_loss = (float)(-t_min / n), entropy_loss = (float)(-t_ent / n);
    out[0] = policy_loss + entropy_coeff * entropy_loss;
    out[1] = policy_loss;
    out[2] = entropy_loss;
    out[3] = (float)(t_kl / n);
}

}  // namespace learn
}  // namespace ssb

namespace ssb {
namespace learn {

// ---- parameter update (TrainableScheduler.update_parameters, schedulers/scheduler.py:37-54: loss.backward();
// clip_grad_norm_(max_grad_norm); optim.step() with torch.optim.Adam, trainer.py opt_cls / opt_kwargs) ------------
// part[k] = block k's sum of g^2 (f64)
__global__ void __launch_bounds__(PPO_THREADS) k_grad_sqsum(const float *grad, int n, double *part)
{
    __shared__ double sh[PPO_THREADS / 32];
    double s = 0.0;
    for (int i = blockIdx.x * PPO_THREADS + threadIdx.x; i < n; i += PPO_BLOCKS * PPO_THREADS) s += (double)grad[i] * grad[i];
    s = block_sum(s, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = s;
}
// clip_grad_norm_: g *= min(1, max_norm / (||g|| + 1e-6)) (a non-finite norm leaves everything untouched); then Adam (no weight decay, no amsgrad):
// m += (g - m) * (1 - b1); v = v * b2 + (1 - b2) * g * g; p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
__global__ void __launch_bounds__(PPO_THREADS)
k_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int n, int step, float lr, float beta1,
            float beta2, float eps, float max_grad_norm, const double *part, float *grad_norm_out)
{
    double ss = 0.0;
    for (int k = 0; k < PPO_BLOCKS; k++) ss += part[k];
    const float total_norm = (float)sqrt(ss);
    float coef = 1.0f;
    if (max_grad_norm > 0.0f) coef = fminf(max_grad_norm / (total_norm + 1e-6f), 1.0f);
    if (grad_norm_out && blockIdx.x == 0 && threadIdx.x == 0) *grad_norm_out = total_norm;
    // clip_grad_norm_(..., error_if_nonfinite=True) (schedulers/scheduler.py:46-48) raises on a NaN / Inf norm before
    // optim.step(): nothing is written here either; the caller sees the non-finite norm in grad_norm_out and raises
    if (!isfinite(total_norm)) return;
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    const float step_size = (float)((double)lr / bc1), bc2_sqrt = (float)sqrt(bc2);
    for (int i = blockIdx.x * PPO_THREADS + threadIdx.x; i < n; i += gridDim.x * PPO_THREADS) {
        const float g = grad[i] * coef;
        const float m = exp_avg[i] + (g - exp_avg[i]) * (1.0f - beta1);
        const float v = exp_avg_sq[i] * beta2 + (1.0f - beta2) * g * g;
        exp_avg[i] = m; exp_avg_sq[i] = v;
        param[i] = param[i] - step_size * (m / (sqrtf(v) / bc2_sqrt + eps));
    }
}

}  // namespace learn
}  // namespace ssb
