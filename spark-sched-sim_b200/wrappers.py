"""Env wrappers with the reference's interface."""
from __future__ import annotations

import numpy as np

from .gym_compat import Wrapper


class StochasticTimeLimit(Wrapper):
    """Samples each episode's time limit from an exponential distribution and raises `truncated`
    once it is reached (spark_sched_sim/wrappers/stochastic_time_limit.py:5-31)."""

    def __init__(self, env, mean_time_limit, seed=42, verbose=False):
        super().__init__(env)
        self.mean_time_limit = mean_time_limit
        self.np_random = np.random.RandomState(seed)
        self.verbose = verbose

    def reset(self, seed=None, options=None):
        if seed:
            self.np_random = np.random.RandomState(seed)
        self.time_limit = self.np_random.exponential(self.mean_time_limit)
        if self.verbose:
            print(f"resetting. seed={seed}, timelim={int(self.time_limit * 1e-3)}s", flush=True)
        options = dict(options or {})
        options["time_limit"] = self.time_limit
        return self.env.reset(seed=seed, options=options)

    def step(self, act):
        obs, rew, term, trunc, info = self.env.step(act)
        if info["wall_time"] >= self.time_limit:
            trunc = True
        return obs, rew, term, trunc, info
