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
