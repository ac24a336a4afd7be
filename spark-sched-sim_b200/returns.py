"""What the trainer derives from the rollout buffers before a policy update, on device tensors
(trainers/trainer.py:172-212): `ReturnsCalculator` (trainers/utils/returns_calculator.py) and `Baseline`
(trainers/utils/baselines.py) with the reference's names and constructor arguments.

The rollouts are the device buffers `BatchedSparkSchedSimEnv.rollout_fair_traj` fills: a uint8 tensor holding
ssb_transition[B][stride], the number of stored steps per rollout and the wall time after the last step.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _native as nat


def _stream(t: torch.Tensor):
    return torch.cuda.current_stream(t.device).cuda_stream


class ReturnsCalculator:
    """beta: continuously discounted returns, R_k = r_k + exp(-beta * 1e-3 * dt_k) * R_{k+1}
    (returns_calculator.py:67-76).  buff_cap: differential returns against the moving estimate of the average
    number of jobs, kept over the latest buff_cap steps of positive duration (returns_calculator.py:52-65, :78-89);
    `avg_num_jobs` is then a device scalar updated by every call."""

    def __init__(self, buff_cap=None, beta=None):
        assert bool(buff_cap) ^ bool(beta), "exactly one of `buff_cap` and `beta` must be specified"
        self.buff_cap = int(buff_cap) if buff_cap else None
        self.beta = float(beta) if beta else None
        self.avg_num_jobs = None
        self._window = None
        self._which = C.c_int32(0)

    def __call__(self, traj: torch.Tensor, num_steps: torch.Tensor, final_wall: torch.Tensor, stride: int):
        B = num_steps.numel()
        assert traj.is_cuda and traj.dtype == torch.uint8 and traj.numel() >= B * stride * nat.TRANSITION_DTYPE.itemsize
        assert num_steps.dtype == torch.int32 and final_wall.dtype == torch.float64
        out = torch.zeros(B, stride, dtype=torch.float64, device=traj.device)
        if self.buff_cap:
            if self._window is None:
                self._window = torch.zeros(2, self.buff_cap, 2, dtype=torch.float64, device=traj.device)
                self.avg_num_jobs = torch.zeros(1, dtype=torch.float64, device=traj.device)
            scratch = torch.empty(2 * B + 1, dtype=torch.int32, device=traj.device)
            nat.check(nat.lib().ssb_differential_returns(
                traj.data_ptr(), num_steps.data_ptr(), final_wall.data_ptr(), B, int(stride), self._window.data_ptr(),
                self.buff_cap, C.byref(self._which), scratch.data_ptr(), self.avg_num_jobs.data_ptr(), out.data_ptr(),
                _stream(traj)), "ssb_differential_returns")
            return out
        nat.check(nat.lib().ssb_discounted_returns(traj.data_ptr(), num_steps.data_ptr(), final_wall.data_ptr(), B,
                                                   int(stride), self.beta, out.data_ptr(), _stream(traj)),
                  "ssb_discounted_returns")
        return out


class Baseline:
    """Interpolated average of the returns over the rollouts of one job sequence (baselines.py:12-37): rollouts
    [j * num_rollouts, (j + 1) * num_rollouts) form group j."""

    def __init__(self, num_sequences, num_rollouts):
        self.num_sequences = int(num_sequences)
        self.num_rollouts = int(num_rollouts)

    def __call__(self, traj: torch.Tensor, returns: torch.Tensor, num_steps: torch.Tensor):
        B, stride = returns.shape
        assert B == self.num_sequences * self.num_rollouts
        out = torch.zeros_like(returns)
        nat.check(nat.lib().ssb_group_baselines(traj.data_ptr(), returns.data_ptr(), num_steps.data_ptr(), B, stride,
                                                self.num_rollouts, out.data_ptr(), _stream(traj)),
                  "ssb_group_baselines")
        return out
