"""Host-side heuristic policies on the base observation dict: fair / FIFO round-robin and random
(schedulers/heuristics/round_robin.py:7-49, random_scheduler.py:7-32, utils.py:5-37).  The fused
on-device equivalent of the fair/FIFO policy is `ssb_rollout_fair` / `ssb_fair_actions`."""
from __future__ import annotations

import numpy as np

from .scheduler import Scheduler


def preprocess_obs(obs: dict) -> None:
    """Adds `frontier_stages` (nodes without an incoming edge in the observed graph) and
    `schedulable_stages` (node id -> rank among schedulable nodes == valid `stage_idx`)."""
    nodes = obs["dag_batch"].nodes
    frontier = np.ones(nodes.shape[0], dtype=bool)
    frontier[obs["dag_batch"].edge_links[:, 1]] = False
    sched = nodes[:, 2].astype(bool).nonzero()[0]
    obs["frontier_stages"] = set(frontier.nonzero()[0].tolist())
    obs["schedulable_stages"] = {int(v): i for i, v in enumerate(sched)}


def find_stage(obs: dict, job_idx: int) -> int:
    """First schedulable frontier stage of the job, else its first schedulable stage, else -1."""
    selected = -1
    for node in range(obs["dag_ptr"][job_idx], obs["dag_ptr"][job_idx + 1]):
        i = obs["schedulable_stages"].get(node)
        if i is None:
            continue
        if node in obs["frontier_stages"]:
            return i
        if selected == -1:
            selected = i
    return selected


class RoundRobinScheduler(Scheduler):
    def __init__(self, num_executors, dynamic_partition=True, **kwargs):
        self.name = "Fair" if dynamic_partition else "FIFO"
        self.num_executors = num_executors
        self.dynamic_partition = dynamic_partition
        self.env_wrapper_cls = None

    def schedule(self, obs: dict) -> tuple[dict, dict]:
        preprocess_obs(obs)
        num_active_jobs = len(obs["exec_supplies"])
        if self.dynamic_partition:
            executor_cap = int(np.ceil(self.num_executors / max(1, num_active_jobs)))
        else:
            executor_cap = self.num_executors
        committable = obs["num_committable_execs"]
        src = obs["source_job_idx"]
        if src < num_active_jobs:  # the job that is releasing executors goes first
            sel = find_stage(obs, src)
            if sel != -1:
                return {"stage_idx": sel, "num_exec": committable}, {}
        for j in range(num_active_jobs):  # then jobs by order of arrival, up to their share
            if obs["exec_supplies"][j] >= executor_cap or j == src:
                continue
            sel = find_stage(obs, j)
            if sel == -1:
                continue
            return {"stage_idx": sel,
                    "num_exec": min(committable, executor_cap - obs["exec_supplies"][j])}, {}
        return {"stage_idx": -1, "num_exec": committable}, {}


class RandomScheduler(Scheduler):
    def __init__(self, seed=42, **kwargs):
        self.name = "Random"
        self.env_wrapper_cls = None
        self.set_seed(seed)

    def set_seed(self, seed):
        self.np_random = np.random.RandomState(seed)

    def schedule(self, obs: dict) -> tuple[dict, dict]:
        preprocess_obs(obs)
        job_idxs = list(range(len(obs["exec_supplies"])))
        stage_idx = -1
        while job_idxs:
            j = self.np_random.choice(job_idxs)
            stage_idx = find_stage(obs, j)
            if stage_idx != -1:
                break
            job_idxs.remove(j)
        num_exec = self.np_random.randint(1, obs["num_committable_execs"] + 1)
        return {"stage_idx": stage_idx, "num_exec": num_exec}, {}
