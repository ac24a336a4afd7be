"""SparkSchedSimEnv: single-environment, Gymnasium-shaped view of the CUDA simulator.

Same constructor, `reset`/`step` contract, observation/action dictionaries, exceptions and public
attributes as the reference class (spark_sched_sim/spark_sched_sim.py:29-245), so the reference's
schedulers (`schedule(obs)`), wrappers and `metrics` run on top of it unchanged.  Internally it is a
BatchedSparkSchedSimEnv with B = 1; every call goes through the C ABI with host buffers.

Differences that are by design:
* job sequences and task durations come from the counter-based Philox streams
  (oracle/philox_ref.py is the spec), not from numpy's PCG64 -- a seed denotes a different, equally
  distributed episode;
* `render_mode="human"` is not supported (the PyGame renderer is out of scope).
"""
from __future__ import annotations

from collections import deque
from types import SimpleNamespace
from typing import Any

import numpy as np

from .batched_env import ERROR_MESSAGES, BatchedSparkSchedSimEnv
from .gym_compat import Env, GraphInstance, spaces

NUM_NODE_FEATURES = 3


class SparkSchedSimEnv(Env):
    metadata = {"render_modes": [], "render_fps": 30}

    def __init__(self, env_cfg: dict[str, Any], bank=None, device="cuda:0", max_jobs: int | None = None,
                 tape_capacity: int = 0, log_capacity: int = 0, decima_obs: bool = True,
                 history_capacity: int = 0):
        self.num_executors: int = env_cfg["num_executors"]
        self.moving_delay = env_cfg["moving_delay"]
        self.beta: float = env_cfg.get("beta", 0)
        self.job_arrival_cap = env_cfg.get("job_arrival_cap")
        if env_cfg.get("render_mode") == "human":
            raise ValueError("pygame rendering is not available in the CUDA env")
        # accept both spellings of the sampler key (config/decima_tpch.yaml:86 vs examples.py:21)
        sampler = env_cfg.get("data_sampler_cls", "TPCHDataSampler")
        if sampler != "TPCHDataSampler" or env_cfg.get("dataset", "tpch") != "tpch":
            raise ValueError(f"'{sampler}' is not a valid data sampler.")
        self._batched = BatchedSparkSchedSimEnv(env_cfg, num_envs=1, bank=bank, device=device,
                                                max_jobs=max_jobs, tape_capacity=tape_capacity,
                                                log_capacity=log_capacity, decima_obs=decima_obs,
                                                decima_policy=decima_obs,
                                                history_capacity=history_capacity)  # one env: the buffers are tiny, so
        # the Decima wrappers / DecimaScheduler work on any env, as `scheduler.env_wrapper_cls(env)` expects
        self.wall_time: float = 0
        self.jobs: dict[int, SimpleNamespace] = {}
        self.active_job_ids: list[int] = []
        self.completed_job_ids: set[int] = set()
        self.job_duration_buff: deque[float] = deque(maxlen=200)  # survives resets (:83)
        self._auto_seed = 0
        self.action_space = spaces.Dict({
            "stage_idx": spaces.Discrete(1, start=-1),
            "num_exec": spaces.Discrete(self.num_executors, start=1),
        })
        self.observation_space = None

    # ------------------------------------------------------------ gymnasium API
    def reset(self, seed: int | None = None, options: dict[str, Any] | None = None):
        options = options or {}
        time_limit = options.get("time_limit", np.inf)
        if time_limit is np.inf and not self.job_arrival_cap:
            raise ValueError("must either have a limit on job arrivals or time.")
        if seed is None:
            self._auto_seed += 1
            seed = (1 << 40) + self._auto_seed
        # time-limited arrivals without a job cap: the handle grows until the episode's jobs fit (the reference's
        # job list is unbounded, spark_sched_sim.py:150-154)
        hdr = self._batched.reset_host(np.array([seed], np.uint64), np.array([time_limit], np.float64), grow=True)
        self._raise_on_error(int(hdr[0]["error"]))
        # per-episode host state starts empty (:145, :173); job_duration_buff survives resets (:83)
        self.jobs = {}
        self.active_job_ids = []
        self.completed_job_ids = set()
        return self._finish(hdr)[0], self.info

    def step(self, action: dict):
        # action_space.contains(): exactly these two integer entries (:276-277)
        if not (isinstance(action, dict) and action.keys() == {"stage_idx", "num_exec"}
                and all(isinstance(action[k], (int, np.integer)) and not isinstance(action[k], bool)
                        or (isinstance(action[k], np.ndarray) and action[k].shape == ()
                            and np.issubdtype(action[k].dtype, np.integer))
                        for k in action)):
            raise ValueError("invalid action: does not belong to the action space")
        hdr = self._batched.step_host(np.array([int(action["stage_idx"])], np.int32),
                                      np.array([int(action["num_exec"])], np.int32))
        self._raise_on_error(int(hdr[0]["error"]))
        obs, h = self._finish(hdr)
        return obs, float(h["reward"]), bool(h["terminated"]), False, self.info

    def close(self) -> None:
        self._batched.close()

    def load_trace(self, t_arrival, template, tape=None) -> None:
        """Parity mode: the next reset() replays this pre-sampled job sequence (and duration tape)."""
        self._batched.load_trace(0, t_arrival, template, tape)

    # ------------------------------------------------------------ attributes other layers read
    @property
    def all_jobs_complete(self) -> bool:
        return self.num_completed_jobs == len(self.jobs)

    @property
    def num_completed_jobs(self) -> int:
        return len(self.completed_job_ids)

    @property
    def num_active_jobs(self) -> int:
        return len(self.active_job_ids)

    @property
    def info(self) -> dict:
        return {"wall_time": self.wall_time}

    @property
    def executors(self) -> list:
        """`env.executors[i].history` as the renderer reads it (spark_sched_sim.py:411, executor.py:25-44); needs
        `history_capacity > 0` at construction."""
        return [SimpleNamespace(id_=i, history=h) for i, h in enumerate(self._batched.executor_histories(0))]

    @property
    def avg_job_duration(self) -> float:
        return np.mean(self.job_duration_buff).item() * 1e-3

    # ------------------------------------------------------------ internals
    def _raise_on_error(self, code: int) -> None:
        if code == 0:
            return
        if code == 2:
            raise KeyError(ERROR_MESSAGES[2])
        if code < 1000:
            raise ValueError(ERROR_MESSAGES.get(code, f"error {code}"))
        raise AssertionError(f"simulator invariant violated (libssb check at ssb_sim.cuh:{code - 1000})")

    def _finish(self, hdr):
        h = hdr[0]
        self.wall_time = float(h["wall_time"])
        o = self._batched.obs(0, hdr)
        ta, tc, tm, st = self._batched.jobs(0, with_state=True)
        newly_done = [j for j in range(len(ta)) if st[j] == 2 and j not in self.completed_job_ids]
        self.jobs = {j: SimpleNamespace(id_=j, t_arrival=float(ta[j]), t_completed=float(tc[j]),
                                        template=int(tm[j]), query_num=int(tm[j]) % 22 + 1,
                                        query_size_idx=int(tm[j]) // 22)
                     for j in range(len(ta))}
        self.active_job_ids = [j for j in range(len(ta)) if st[j] == 1]
        for j in sorted(newly_done, key=lambda j: tc[j]):
            self.completed_job_ids.add(j)
            self.job_duration_buff.append(float(tc[j] - ta[j]))
        self.job_arrival_cap = len(self.jobs)
        n_nodes = o["nodes"].shape[0]
        self.action_space["stage_idx"].n = n_nodes + 1
        edge_links = o["edge_links"].astype(np.int64)
        obs = {
            "dag_batch": GraphInstance(o["nodes"], np.zeros(len(edge_links), dtype=int), edge_links),
            "dag_ptr": o["dag_ptr"].tolist(),
            "num_committable_execs": o["num_committable_execs"],
            "source_job_idx": o["source_job_idx"],
            "exec_supplies": o["exec_supplies"].tolist(),
        }
        return obs, h
