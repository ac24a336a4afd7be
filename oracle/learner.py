"""CPU restatement (TEST INFRASTRUCTURE, never imported by the product) of what the reference's trainer computes
from the rollout buffers before a policy update (trainers/trainer.py:172-212):

  * discounted_returns  -- ReturnsCalculator._calc_discounted_returns (trainers/utils/returns_calculator.py:67-76)
  * group_baselines     -- Baseline.average/_average (trainers/utils/baselines.py:12-37), including the two numpy
                           routines it leans on: np.interp (compiled_base.c:arr_interp) and the pairwise summation
                           behind ndarray.mean (loops_utils.h.src:pairwise_sum)

Pinned by tests/test_learner_oracle.py against tests/golden/learner_vectors.npz, which
tests/golden/gen_learner_golden.py produced by running the reference's own classes.
"""
from __future__ import annotations

import math

import numpy as np


def discounted_returns(rewards, times, beta):
    """rewards: K floats; times: K + 1 wall times (the last one after the final step)."""
    K = len(rewards)
    out = np.zeros(K)
    R = 0.0
    for k in range(K - 1, -1, -1):
        dt = float(times[k + 1]) - float(times[k])
        R = float(rewards[k]) + math.exp(-beta * 1e-3 * dt) * R   # returns_calculator.py:73
        out[k] = R
    return out


class DifferentialReturns:
    """ReturnsCalculator with buff_cap (returns_calculator.py:24-65, :78-89): CircularArray window of the latest
    `cap` (dt, reward) rows with dt > 0 over all rollouts in order; avg_num_jobs = -sum(reward) / sum(dt) with the
    rows added one after the other (ndarray.sum(0) of a C-contiguous (cap, 2) array)."""

    def __init__(self, cap):
        self.cap = cap
        self.data = np.zeros((cap, 2))
        self.avg_num_jobs = None

    def __call__(self, rewards_list, times_list):
        dts = [np.asarray(t[1:], dtype=np.float64) - np.asarray(t[:-1], dtype=np.float64) for t in times_list]
        rows = [(float(dt), float(r)) for d, rs in zip(dts, rewards_list) for dt, r in zip(d, rs) if dt > 0]
        rows = rows[-self.cap:]
        keep = self.cap - len(rows)
        if keep > 0:
            self.data[:keep] = self.data[self.cap - keep:].copy()
        if rows:
            self.data[keep:] = np.asarray(rows)
        total_time = rew_sum = 0.0
        for i in range(self.cap):
            total_time += self.data[i, 0]
            rew_sum += self.data[i, 1]
        self.avg_num_jobs = -rew_sum / total_time
        out = []
        for d, rs in zip(dts, rewards_list):
            ret = np.zeros(len(rs))
            R = 0.0
            for k in range(len(rs) - 1, -1, -1):
                R = -((-float(rs[k])) - float(d[k]) * self.avg_num_jobs) + R
                ret[k] = R
            out.append(ret)
        return out


def interp1(x, xp, fp):
    """np.interp for one point; xp non-decreasing with repeats allowed."""
    n = len(xp)
    if x < xp[0]:
        return float(fp[0])
    if x >= xp[n - 1]:
        return float(fp[n - 1])
    lo, hi = 0, n - 1            # xp[lo] <= x < xp[hi]
    while hi - lo > 1:
        mid = (lo + hi) >> 1
        if xp[mid] <= x:
            lo = mid
        else:
            hi = mid
    if xp[lo] == x:
        return float(fp[lo])
    slope = (float(fp[lo + 1]) - float(fp[lo])) / (float(xp[lo + 1]) - float(xp[lo]))
    return slope * (float(x) - float(xp[lo])) + float(fp[lo])


def pairwise_mean(vals):
    """ndarray.mean() of up to 128 float64 values: numpy's pairwise sum, then / n."""
    n = len(vals)
    assert n <= 128
    if n < 8:
        res = -0.0
        for v in vals:
            res += float(v)
    else:
        r = [float(v) for v in vals[:8]]
        full = n - n % 8
        for i in range(8, full, 8):
            for j in range(8):
                r[j] += float(vals[i + j])
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        for i in range(full, n):
            res += float(vals[i])
    return res / n


def group_baselines(ts_list, ys_list, num_rollouts):
    """ts_list[i]: step times of rollout i (len K_i); ys_list[i]: its returns.  Consecutive groups of
    `num_rollouts` rollouts share a job sequence (baselines.py:19-24)."""
    out = []
    for g0 in range(0, len(ts_list), num_rollouts):
        members = list(range(g0, min(g0 + num_rollouts, len(ts_list))))
        for i in members:
            b = np.zeros(len(ts_list[i]))
            for k, t in enumerate(ts_list[i]):
                b[k] = pairwise_mean([interp1(t, ts_list[j], ys_list[j]) for j in members])
            out.append(b)
    return out
