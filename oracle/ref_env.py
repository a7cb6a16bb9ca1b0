"""TEST INFRASTRUCTURE ONLY -- runs the reference's own code, unmodified, from /root/reference.

`load_reference_env()` imports /root/reference/gym_rad_search/gym_rad_search/envs/rad_search_env.py through the
import stubs in oracle/shims (gym, matplotlib) and the exact-arithmetic `visilibity` restatement.
`load_reference_ppo()` imports /root/reference/algos/multiagent/ppo.py (PPOBuffer, discount_cumsum) through the
`ray` / `mpi4py` stubs.

/root/reference only exists in the build container; on the GPU box these loaders raise and the tests that need
them skip -- the committed vectors under tests/golden/ (made by tests/golden/make_golden.py with these loaders) stand in.
Nothing in the product package imports this file.
"""
from __future__ import annotations

import copy
import importlib
import os
import sys

import numpy as np

REFERENCE_ROOT = os.environ.get("RADSEARCH_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def reference_available() -> bool:
    return os.path.isfile(
        os.path.join(REFERENCE_ROOT, "gym_rad_search", "gym_rad_search", "envs", "rad_search_env.py")
    )


def _prepare_path() -> None:
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for p in (os.path.join(REFERENCE_ROOT, "gym_rad_search"), REFERENCE_ROOT, _SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)
    # the shims must win over anything of the same name
    sys.path.remove(_SHIMS)
    sys.path.insert(0, _SHIMS)


def load_reference_env():
    """Return the reference module gym_rad_search.envs.rad_search_env (RadSearch, get_step, ...)."""
    _prepare_path()
    return importlib.import_module("gym_rad_search.envs.rad_search_env")


def load_reference_ppo():
    """Return the reference module algos.multiagent.ppo (PPOBuffer, discount_cumsum, ...)."""
    _prepare_path()
    return importlib.import_module("algos.multiagent.ppo")


class RecordingGenerator:
    """Wraps a numpy Generator; behaves identically, and for every poisson() call records (lam, count, uniforms)
    where `uniforms` are the next `n_uniforms` doubles the bit generator would have produced from the state it had
    when poisson() was entered -- i.e. exactly the stream numpy's PTRS consumed.  That is what "the reference's
    uniforms injected" (BASELINE.json north_star) means for the kernels."""

    def __init__(self, gen: np.random.Generator, n_uniforms: int = 32):
        self._gen = gen
        self._n = n_uniforms
        self.poisson_log = []

    def poisson(self, lam, size=None):
        assert size is None
        state = copy.deepcopy(self._gen.bit_generator.state)
        k = self._gen.poisson(lam)
        probe = np.random.Generator(type(self._gen.bit_generator)())
        probe.bit_generator.state = state
        u = probe.random(self._n)
        self.poisson_log.append((float(lam), int(k), u))
        return k

    def __getattr__(self, name):
        return getattr(self._gen, name)
