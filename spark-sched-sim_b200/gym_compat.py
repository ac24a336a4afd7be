"""Gymnasium compatibility: use the real package when it is installed, otherwise the few classes
the env's contract needs (spaces for membership tests, GraphInstance, Env/Wrapper bases).  The
observation/action dictionaries are identical either way (spark_sched_sim.py:85-125, :393-399)."""
from __future__ import annotations

from typing import NamedTuple

import numpy as np

try:  # pragma: no cover - depends on the image
    import gymnasium as gym
    from gymnasium import Env, Wrapper, ObservationWrapper, ActionWrapper
    from gymnasium.spaces import GraphInstance
    import gymnasium.spaces as spaces

    HAVE_GYMNASIUM = True
except ImportError:
    HAVE_GYMNASIUM = False

    class GraphInstance(NamedTuple):
        nodes: np.ndarray
        edges: np.ndarray
        edge_links: np.ndarray

    class _Space:
        def contains(self, x) -> bool:
            return True

        def __contains__(self, x) -> bool:
            return self.contains(x)

    class _Discrete(_Space):
        def __init__(self, n, seed=None, start=0):
            self.n, self.start = int(n), int(start)

        def contains(self, x) -> bool:
            if isinstance(x, (int, np.integer)) or (
                isinstance(x, np.ndarray) and x.shape == () and np.issubdtype(x.dtype, np.integer)
            ):
                return bool(self.start <= int(x) < self.start + self.n)
            return False

    class _Generic(_Space):
        def __init__(self, *args, **kwargs):
            self.args, self.kwargs = args, kwargs
            self.n = args[0] if args else None
            self.feature_space = args[0] if args else None

    class _Dict(_Space):
        def __init__(self, spaces=None, **kw):
            self.spaces = dict(spaces or {})
            self.spaces.update(kw)

        def __getitem__(self, k):
            return self.spaces[k]

        def contains(self, x) -> bool:
            return (isinstance(x, dict) and x.keys() == self.spaces.keys()
                    and all(self.spaces[k].contains(x[k]) for k in self.spaces))

    class spaces:  # noqa: N801 - namespace look-alike of gymnasium.spaces
        Discrete = _Discrete
        Dict = _Dict
        Box = Graph = Sequence = MultiBinary = _Generic
        GraphInstance = GraphInstance

    class Env:
        metadata: dict = {}

        def reset(self, *, seed=None, options=None):
            return None

        @property
        def unwrapped(self):
            return self

        def close(self):
            pass

    class Wrapper(Env):
        def __init__(self, env):
            self.env = env

        def __getattr__(self, name):
            if name.startswith("_") or name == "env":
                raise AttributeError(name)
            return getattr(self.env, name)

        @property
        def unwrapped(self):
            return self.env.unwrapped

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
            obs, r, term, trunc, info = self.env.step(action)
            return self.observation(obs), r, term, trunc, info

    class ActionWrapper(Wrapper):
        def step(self, action):
            return self.env.step(self.action(action))
