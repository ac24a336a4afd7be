"""Scheduler plug-ins with the reference's interface (schedulers/scheduler.py:10-18):
`schedule(obs) -> (action, info)`, attributes `name` and `env_wrapper_cls`."""
from .scheduler import Scheduler
from .heuristics import RandomScheduler, RoundRobinScheduler, find_stage, preprocess_obs
from .decima import DecimaScheduler

__all__ = ["Scheduler", "RoundRobinScheduler", "RandomScheduler", "DecimaScheduler", "make_scheduler",
           "find_stage", "preprocess_obs"]


def make_scheduler(agent_cfg):
    """schedulers/__init__.py:17-21: class name looked up in this module."""
    from copy import deepcopy

    glob = globals()
    agent_cls = agent_cfg["agent_cls"]
    assert agent_cls in glob, f"'{agent_cls}' is not a valid scheduler."
    cfg = deepcopy(agent_cfg)
    cfg.pop("agent_cls")
    return glob[agent_cls](**cfg)
