"""`gymnasium.envs.registration` stand-in (see ../__init__.py): a dict of id -> entry point."""
registry: dict = {}


def register(id, entry_point=None, **kwargs):  # noqa: A002
    registry[id] = entry_point
