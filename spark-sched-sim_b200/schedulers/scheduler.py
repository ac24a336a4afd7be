from abc import ABC, abstractmethod


class Scheduler(ABC):
    """Interface for all schedulers (schedulers/scheduler.py:10-18)."""

    name: str
    env_wrapper_cls = None

    @abstractmethod
    def schedule(self, obs: dict) -> tuple[dict, dict]:
        ...
