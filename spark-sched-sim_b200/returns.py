"""What the trainer derives from the rollout buffers before a policy update, on device tensors
(trainers/trainer.py:172-212): `ReturnsCalculator` (trainers/utils/returns_calculator.py) and `Baseline`
(trainers/utils/baselines.py) with the reference's names and constructor arguments.

The rollouts are the device buffers `BatchedSparkSchedSimEnv.rollout_fair_traj` fills: a uint8 tensor holding
ssb_transition[B][stride], the number of stored steps per rollout and the wall time after the last step.
"""
from __future__ import annotations

import torch

from . import _native as nat


def _stream(t: torch.Tensor):
    return torch.cuda.current_stream(t.device).cuda_stream


class ReturnsCalculator:
    """Continuously discounted returns, R_k = r_k + exp(-beta * 1e-3 * dt_k) * R_{k+1}
    (returns_calculator.py:67-76).  Differential returns (`buff_cap`) are not implemented on the device."""

    def __init__(self, buff_cap=None, beta=None):
        assert bool(buff_cap) ^ bool(beta), "exactly one of `buff_cap` and `beta` must be specified"
        if buff_cap:
            raise NotImplementedError("differential returns (returns_calculator.py:52-65) are not on the device yet")
        self.beta = float(beta)

    def __call__(self, traj: torch.Tensor, num_steps: torch.Tensor, final_wall: torch.Tensor, stride: int):
        B = num_steps.numel()
        assert traj.is_cuda and traj.dtype == torch.uint8 and traj.numel() >= B * stride * nat.TRANSITION_DTYPE.itemsize
        assert num_steps.dtype == torch.int32 and final_wall.dtype == torch.float64
        out = torch.zeros(B, stride, dtype=torch.float64, device=traj.device)
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
