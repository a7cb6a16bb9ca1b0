"""TEST INFRASTRUCTURE ONLY -- import-only stand-in for `gym` so that the reference's
gym_rad_search/envs/rad_search_env.py (imports at :12-13) loads unmodified.  No behaviour beyond what the env touches."""
from . import spaces  # noqa: F401


class Env:
    metadata: dict = {}

    def reset(self):
        raise NotImplementedError

    def step(self, action):
        raise NotImplementedError


_REGISTRY = {}


def make(id, **kwargs):
    import importlib

    mod, cls = _REGISTRY[id].split(":")
    return getattr(importlib.import_module(mod), cls)(**kwargs)
