import torch


class Batch:
    """Attribute + item access, `in`, `.to(device)` -- what schedulers/decima/{utils,scheduler}.py need."""

    def __init__(self, **kwargs):
        self.__dict__["_store"] = {}
        for k, v in kwargs.items():
            self._store[k] = v

    def __getattr__(self, key):
        try:
            return self.__dict__["_store"][key]
        except KeyError:
            raise AttributeError(key)

    def __setattr__(self, key, value):
        self._store[key] = value

    def __getitem__(self, key):
        return self._store[key]

    def __setitem__(self, key, value):
        self._store[key] = value

    def __contains__(self, key):
        return key in self._store

    def to(self, device, non_blocking=False):
        for k, v in self._store.items():
            if isinstance(v, torch.Tensor):
                self._store[k] = v.to(device)
        return self
