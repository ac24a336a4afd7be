"""DecimaScheduler with the reference's plug-in interface (schedulers/decima/scheduler.py:22-139 and the
TrainableScheduler base, schedulers/scheduler.py:21-54): same constructor keywords, `name`, `env_wrapper_cls`,
`schedule(obs) -> (action, info)`, `evaluate_actions(obsns, actions) -> {"lgprobs", "entropies"}`,
`update_parameters(loss)`, `device`.  Everything runs on the device (tensor-core MLPs, Philox policy stream, the
backward pass, clip_grad_norm_ + Adam); `obs` for `schedule` must come from `DecimaEnvWrapper`, which attaches the env
handle the policy is evaluated on.

    scheduler = make_scheduler(cfg["agent"] | {"num_executors": 10, "state_dict_path": "models/decima/model.pt"})
    env = scheduler.env_wrapper_cls(SparkSchedSimEnv(env_cfg))
    obs, _ = env.reset(seed=1234)
    action, info = scheduler.schedule(obs)         # {"stage_idx", "job_idx", "num_exec"}, {"lgprob"}

Training (the learner's side of trainers/ppo.py:72-102) works on a batched handle and its RolloutStore:

    scheduler = DecimaScheduler(..., opt_cls="Adam", opt_kwargs={"lr": 3e-4}, max_grad_norm=0.5).bind(batched_env)
    res = scheduler.evaluate_actions(store.samples(idx), actions)     # idx: up to B dataset indices k * B + b
    out, g_lp, g_en = ppo.PPOLoss(...)(res["lgprobs"], old_lgprobs, res["entropies"], returns, baselines)
    scheduler.update_parameters(ppo.Loss(g_lp, g_en))                 # backward, clip, Adam, new weights in the policy
"""
from __future__ import annotations

from typing import Any

import numpy as np

from ..decima import DecimaEnvWrapper
from .scheduler import Scheduler


class DecimaScheduler(Scheduler):
    def __init__(self, num_executors: int, embed_dim: int = 16, gnn_mlp_kwargs: dict[str, Any] | None = None,
                 policy_mlp_kwargs: dict[str, Any] | None = None, state_dict_path: str | None = None,
                 state_dict: dict | None = None, opt_cls: str | None = None, opt_kwargs: dict[str, Any] | None = None,
                 max_grad_norm: float | None = None, **kwargs):
        self.name = "Decima"
        self.env_wrapper_cls = DecimaEnvWrapper
        self.num_executors = num_executors
        self.max_grad_norm = max_grad_norm
        if opt_cls not in (None, "Adam"):
            raise ValueError("the device learner implements torch.optim.Adam (config/decima_tpch.yaml: opt_cls 'Adam')")
        self._opt = (opt_cls, dict(opt_kwargs or {}))
        self.optim = None
        self._env = None
        self._pending = None
        gnn = (gnn_mlp_kwargs or {}).get("hid_dims", [32, 16])
        pol = (policy_mlp_kwargs or {}).get("hid_dims", [64, 64])
        if embed_dim != 16 or list(gnn) != [32, 16] or list(pol) != [64, 64]:
            raise ValueError("the device policy implements the shipped architecture (config/decima_tpch.yaml:66-77): "
                             "embed_dim 16, GNN hidden [32, 16], policy hidden [64, 64]")
        if state_dict is None:
            if not state_dict_path:
                raise ValueError("DecimaScheduler needs the weights: state_dict_path or state_dict")
            self.name += f":{state_dict_path}"
            if state_dict_path.endswith(".npz"):
                z = np.load(state_dict_path)
                state_dict = {k: z[k] for k in z.files}
            else:
                import torch

                state_dict = torch.load(state_dict_path, map_location="cpu")
        self._state_dict = state_dict
        self._loaded_into: set[int] = set()

    # ------------------------------------------------------------ TrainableScheduler (schedulers/scheduler.py:21-54)
    def bind(self, env):
        """Attaches the batched handle the learner works on: uploads the weights and, when an optimiser was
        configured, creates it over the flat parameter vector on the handle's device."""
        import torch

        from .. import ppo

        self._env = env
        env.set_decima_weights(self._state_dict)
        self._loaded_into.add(id(env))
        if self._opt[0]:
            sd = self._state_dict
            flat = np.concatenate([np.asarray(sd[k].detach().cpu().numpy() if hasattr(sd[k], "detach") else sd[k],
                                              np.float32).reshape(-1) for k in env.DECIMA_PARAM_ORDER])
            kw = self._opt[1]
            self.optim = ppo.Adam(torch.from_numpy(flat).to(env.device).contiguous(), lr=kw.get("lr", 1e-3),
                                  betas=kw.get("betas", (0.9, 0.999)), eps=kw.get("eps", 1e-8),
                                  max_grad_norm=self.max_grad_norm)
        return self

    @property
    def device(self):
        if self._env is None:
            raise ValueError("DecimaScheduler.device: bind(env) first")
        return self._env.device

    def evaluate_actions(self, obsns, actions) -> dict:
        """obsns: `RolloutStore.samples(idx)` (up to B stored observations); actions: int tensor [n, 3] of
        (stage_idx, job_idx, num_exec) in Decima's format, as RolloutBuffer keeps them.  Returns the log-probabilities
        and normalised entropies of those actions under the current weights (device float32 [n]); the evaluated
        mini-batch stays in place for `update_parameters`."""
        import torch

        env = self._env
        if env is None:
            raise ValueError("DecimaScheduler.evaluate_actions: bind(env) first")
        if self._pending is not None:
            env.decima_snapshot_unload()
            self._pending = None
        store, ks, bs, n = obsns
        actions = torch.as_tensor(actions, device=env.device).reshape(-1, 3).int()
        assert actions.shape[0] == n <= env.num_envs
        if getattr(self, "_staging", None) is None:
            self._staging = torch.empty(env.decima_snapshot_bytes(), dtype=torch.uint8, device=env.device)
        stage_sel = torch.zeros(env.num_envs, dtype=torch.int32, device=env.device)
        exec_sel = torch.zeros_like(stage_sel)
        stage_sel[:n], exec_sel[:n] = actions[:, 0], actions[:, 2]
        env.decima_snapshot_gather(store.snapshots, ks, bs, out=self._staging)
        env.decima_snapshot_load(self._staging)
        self._pending = n
        lg, en = env.decima_evaluate(None, stage_sel, exec_sel, for_backward=True)
        return {"lgprobs": lg[:n], "entropies": en[:n]}

    def update_parameters(self, loss=None) -> None:
        """loss.backward() -> clip_grad_norm_ -> optim.step() (schedulers/scheduler.py:37-54): `loss` carries the loss
        head's adjoint seeds (ppo.Loss) for the mini-batch of the last evaluate_actions call."""
        import torch

        env = self._env
        assert self.optim is not None, "DecimaScheduler was constructed without an optimiser (opt_cls)"
        assert self._pending is not None, "update_parameters follows evaluate_actions"
        n, B = self._pending, env.num_envs
        try:
            grads = torch.zeros_like(self.optim.params)
            if loss is not None:
                gl = torch.zeros(B, dtype=torch.float32, device=env.device)
                ge = torch.zeros_like(gl)
                gl[:n], ge[:n] = loss.grad_lgprob, loss.grad_entropy
                env.decima_backward(gl, ge, grads)
        finally:
            env.decima_snapshot_unload()
            self._pending = None
        self.optim.step(grads)
        env.set_decima_weights(self.optim.params)

    def schedule(self, obs: dict) -> tuple[dict, dict]:
        env = obs.get("_ssb_env")
        if env is None:
            raise ValueError("DecimaScheduler.schedule needs an observation from DecimaEnvWrapper")
        if id(env) not in self._loaded_into:
            env.set_decima_weights(self._state_dict)
            self._loaded_into.add(id(env))
        env.decima_policy()
        stage_idx, job_idx, num_exec, _ = (int(x) for x in env.pol_action[0].tolist())
        return ({"stage_idx": stage_idx, "job_idx": job_idx, "num_exec": num_exec},
                {"lgprob": float(env.pol_lgprob[0].item())})
