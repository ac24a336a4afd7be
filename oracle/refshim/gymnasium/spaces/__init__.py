"""`gymnasium.spaces` stand-in (see ../__init__.py).  Only membership tests and the attributes
the reference mutates (`Discrete.n`, `Sequence.feature_space`, `MultiBinary.n`) are modelled."""
from __future__ import annotations

from typing import NamedTuple

import numpy as np


class GraphInstance(NamedTuple):
    nodes: np.ndarray
    edges: np.ndarray
    edge_links: np.ndarray


class Space:
    def contains(self, x) -> bool:  # pragma: no cover - overridden
        return True

    def __contains__(self, x) -> bool:
        return self.contains(x)


class Discrete(Space):
    def __init__(self, n, seed=None, start=0):
        self.n = int(n)
        self.start = int(start)

    def contains(self, x) -> bool:
        # gymnasium 0.29: python ints and 0-d numpy integer scalars/arrays only
        if isinstance(x, bool):
            as_int = int(x)
        elif isinstance(x, int):
            as_int = x
        elif isinstance(x, (np.generic, np.ndarray)) and (
            np.issubdtype(x.dtype, np.integer) and x.shape == ()
        ):
            as_int = int(x)
        else:
            return False
        return bool(self.start <= as_int < self.start + self.n)


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        self.low, self.high, self.shape, self.dtype = low, high, shape, dtype


class MultiBinary(Space):
    def __init__(self, n, seed=None):
        self.n = n


class Graph(Space):
    def __init__(self, node_space, edge_space, seed=None):
        self.node_space = node_space
        self.edge_space = edge_space


class Sequence(Space):
    def __init__(self, space, seed=None, stack=False):
        self.feature_space = space
        self.stack = stack


class Dict(Space):
    def __init__(self, spaces=None, seed=None, **kwargs):
        self.spaces = dict(spaces or {})
        self.spaces.update(kwargs)

    def __getitem__(self, key):
        return self.spaces[key]

    def __setitem__(self, key, value):
        self.spaces[key] = value

    def keys(self):
        return self.spaces.keys()

    def contains(self, x) -> bool:
        if isinstance(x, dict) and x.keys() == self.spaces.keys():
            return all(self.spaces[k].contains(x[k]) for k in self.spaces)
        return False
