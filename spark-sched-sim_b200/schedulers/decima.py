"""DecimaScheduler with the reference's plug-in interface (schedulers/decima/scheduler.py:22-99): same constructor
keywords, `name`, `env_wrapper_cls`, `schedule(obs) -> (action, info)`.  The forward pass and the sampling run on the
device (ssb_decima_policy: tensor-core MLPs, Philox policy stream); `obs` must come from `DecimaEnvWrapper`, which
attaches the env handle the policy is evaluated on.  Inference only (training_mode / optimiser arguments are accepted
and ignored; the PPO update is not part of this package yet).

    scheduler = make_scheduler(cfg["agent"] | {"num_executors": 10, "state_dict_path": "models/decima/model.pt"})
    env = scheduler.env_wrapper_cls(SparkSchedSimEnv(env_cfg))
    obs, _ = env.reset(seed=1234)
    action, info = scheduler.schedule(obs)         # {"stage_idx", "job_idx", "num_exec"}, {"lgprob"}
"""
from __future__ import annotations

from typing import Any

import numpy as np

from ..decima import DecimaEnvWrapper
from .scheduler import Scheduler


class DecimaScheduler(Scheduler):
    def __init__(self, num_executors: int, embed_dim: int = 16, gnn_mlp_kwargs: dict[str, Any] | None = None,
                 policy_mlp_kwargs: dict[str, Any] | None = None, state_dict_path: str | None = None,
                 state_dict: dict | None = None, **kwargs):
        self.name = "Decima"
        self.env_wrapper_cls = DecimaEnvWrapper
        self.num_executors = num_executors
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
