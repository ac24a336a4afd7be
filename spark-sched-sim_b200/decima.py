"""Decima's env adapters with the reference's interface (schedulers/decima/env_wrapper.py:12-161),
backed by the device kernel `ssb_decima_obs` instead of numpy + networkx on the host.

    env = DecimaEnvWrapper(SparkSchedSimEnv(env_cfg, decima_obs=True))

The observation keeps the reference's keys: `dag_batch` (GraphInstance with the 5 node features),
`dag_ptr`, `stage_mask`, `exec_mask` (bool[Ja, E]), `edge_masks` (bool[depth, M]).  For batched use
read `BatchedSparkSchedSimEnv.dec_*` device tensors directly (no host round trip).
"""
from __future__ import annotations

from typing import Any

import numpy as np

from .gym_compat import ActionWrapper, GraphInstance, ObservationWrapper, Wrapper

NUM_NODE_FEATURES = 5


class DecimaActWrapper(ActionWrapper):
    """Decima's action {stage_idx, job_idx, num_exec in [0, E)} -> env action (env_wrapper.py:19-34)."""

    def __init__(self, env) -> None:
        super().__init__(env)

    def action(self, act: dict[str, Any]) -> dict[str, Any]:
        return {"stage_idx": act["stage_idx"], "num_exec": 1 + act["num_exec"]}


class DecimaObsWrapper(ObservationWrapper):
    """Base observation -> Decima observation (env_wrapper.py:37-161), computed on the device."""

    def __init__(self, env, num_tasks_scale: int = 200, work_scale: float = 1e5) -> None:
        super().__init__(env)
        if num_tasks_scale != 200 or work_scale != 1e5:
            raise ValueError("the device adapter uses the reference's default feature scales")
        self._batched = env.unwrapped._batched
        if not self._batched.has_decima_obs:
            raise ValueError("create the env with decima_obs=True to use the Decima wrappers")
        self.num_executors = env.unwrapped.num_executors

    def observation(self, obs: dict[str, Any]) -> dict[str, Any]:
        self._batched.decima_obs()
        d = self._batched.decima_obs_host(0)
        base = obs["dag_batch"]
        return {
            "dag_batch": GraphInstance(nodes=d["features"], edges=base.edges, edge_links=base.edge_links),
            "dag_ptr": obs["dag_ptr"],
            "stage_mask": d["stage_mask"],
            "exec_mask": d["exec_mask"],
            "edge_masks": d["edge_masks"],
            # handle for schedulers.DecimaScheduler, which evaluates the policy on the device (extra key:
            # the reference's schedulers only read the keys above)
            "_ssb_env": self._batched,
        }


class DecimaEnvWrapper(Wrapper):
    def __init__(self, env):
        super().__init__(DecimaObsWrapper(DecimaActWrapper(env)))
