"""Minimal stand-in for `gymnasium` 0.29 -- TEST INFRASTRUCTURE ONLY.

The reference (ArchieGertsman/spark-sched-sim) imports gymnasium for its `Env` base class, the
wrapper classes, `make`/`register` and a handful of `spaces`.  gymnasium is not installed in this
image and there is no network, so `oracle/refrun.py` puts this directory on `sys.path` to execute
the UNMODIFIED reference from `/root/reference` and record golden traces.  Only the surface the
reference touches is provided:

* `Env.reset(seed=...)` seeds `self.np_random` as gymnasium does
  (`np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))`),
* `NP_RANDOM_FACTORY` -- optional hook: when set, `Env.reset(seed)` installs
  `NP_RANDOM_FACTORY(seed)` instead.  Golden generation uses it to plug the counter-based Philox
  stream (oracle/philox_ref.py) into the reference's own sampler, so that the reference, the C
  oracle and the CUDA kernels all consume the same random numbers.

Nothing in the product package imports this module.
"""
from __future__ import annotations

import importlib

import numpy as np

from . import spaces  # noqa: F401
from .envs import registration as _registration

NP_RANDOM_FACTORY = None  # type: ignore[var-annotated]


class Env:
    metadata: dict = {}
    render_mode = None
    action_space = None
    observation_space = None
    _np_random = None

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            if NP_RANDOM_FACTORY is not None:
                self._np_random = NP_RANDOM_FACTORY(seed)
            else:
                self._np_random = np.random.Generator(
                    np.random.PCG64(np.random.SeedSequence(seed))
                )
        return None

    @property
    def np_random(self):
        if self._np_random is None:
            if NP_RANDOM_FACTORY is not None:
                self._np_random = NP_RANDOM_FACTORY(None)
            else:
                self._np_random = np.random.Generator(np.random.PCG64())
        return self._np_random

    @np_random.setter
    def np_random(self, value):
        self._np_random = value

    @property
    def unwrapped(self):
        return self

    def step(self, action):
        raise NotImplementedError

    def close(self):
        pass


class Wrapper(Env):
    def __init__(self, env):
        self.env = env
        self._action_space = None
        self._observation_space = None

    # spaces fall through to the wrapped env unless overridden
    @property
    def action_space(self):
        if self._action_space is None:
            return self.env.action_space
        return self._action_space

    @action_space.setter
    def action_space(self, space):
        self._action_space = space

    @property
    def observation_space(self):
        if self._observation_space is None:
            return self.env.observation_space
        return self._observation_space

    @observation_space.setter
    def observation_space(self, space):
        self._observation_space = space

    @property
    def np_random(self):
        return self.env.np_random

    @np_random.setter
    def np_random(self, value):
        self.env.np_random = value

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def __getattr__(self, name):
        if name.startswith("_") or name == "env":
            raise AttributeError(name)
        return getattr(self.env, name)

    def reset(self, *, seed=None, options=None):
        return self.env.reset(seed=seed, options=options)

    def step(self, action):
        return self.env.step(action)

    def close(self):
        return self.env.close()


class ObservationWrapper(Wrapper):
    def reset(self, *, seed=None, options=None):
        obs, info = self.env.reset(seed=seed, options=options)
        return self.observation(obs), info

    def step(self, action):
        obs, reward, terminated, truncated, info = self.env.step(action)
        return self.observation(obs), reward, terminated, truncated, info

    def observation(self, observation):
        raise NotImplementedError


class ActionWrapper(Wrapper):
    def step(self, action):
        return self.env.step(self.action(action))

    def action(self, action):
        raise NotImplementedError


def register(id, entry_point, **kwargs):  # noqa: A002
    _registration.register(id, entry_point, **kwargs)


def make(id, **kwargs):  # noqa: A002
    """`make("module:EnvId", **kwargs)`: import `module` (which registers), build the env."""
    if ":" in id:
        module, env_id = id.split(":", 1)
        importlib.import_module(module)
    else:
        env_id = id
    entry_point = _registration.registry[env_id]
    if isinstance(entry_point, str):
        mod_name, cls_name = entry_point.split(":")
        entry_point = getattr(importlib.import_module(mod_name), cls_name)
    return entry_point(**kwargs)
