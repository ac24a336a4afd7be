"""The loss head of the reference's PPO trainer (trainers/ppo.py:104-140 `_compute_loss`) on the device:
forward values and the adjoint seeds d loss / d lgprob, d loss / d entropy for the policy's backward pass."""
import torch

from . import _native as nat
from .returns import _stream


class PPOLoss:
    """clip_range / entropy_coeff as in the trainer's config (ppo.py:43-46).  __call__ takes the flat per-sample
    device arrays (new_lgprob, old_lgprob, entropy: float32; returns, baselines: float64) and an optional int32
    index tensor selecting the mini-batch; returns ({"loss", "policy_loss", "entropy_loss", "approx_kl_div"} as a
    float32 device tensor of 4, grad_lgprob, grad_entropy)."""

    KEYS = ("loss", "policy_loss", "entropy_loss", "approx_kl_div")

    def __init__(self, clip_range=0.2, entropy_coeff=0.0):
        self.clip_range = float(clip_range)
        self.entropy_coeff = float(entropy_coeff)
        self._scratch = None

    def __call__(self, new_lgprob, old_lgprob, entropy, returns, baselines, idx=None):
        for t in (new_lgprob, old_lgprob, entropy):
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        for t in (returns, baselines):
            assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()
        dev = new_lgprob.device
        if idx is not None:
            assert idx.is_cuda and idx.dtype == torch.int32 and idx.is_contiguous()
        n = int(idx.numel()) if idx is not None else int(new_lgprob.numel())
        if self._scratch is None or self._scratch.device != dev:
            self._scratch = torch.empty(640, dtype=torch.float64, device=dev)
        out = torch.empty(4, dtype=torch.float32, device=dev)
        g_lp = torch.empty(n, dtype=torch.float32, device=dev)
        g_en = torch.empty(n, dtype=torch.float32, device=dev)
        nat.check(nat.lib().ssb_ppo_loss(
            new_lgprob.data_ptr(), old_lgprob.data_ptr(), entropy.data_ptr(), returns.data_ptr(), baselines.data_ptr(),
            idx.data_ptr() if idx is not None else None, n, self.clip_range, self.entropy_coeff,
            self._scratch.data_ptr(), out.data_ptr(), g_lp.data_ptr(), g_en.data_ptr(), _stream(new_lgprob)),
            "ssb_ppo_loss")
        return out, g_lp, g_en
