"""TEST INFRASTRUCTURE ONLY -- import-only `ray` stub (algos/multiagent/ppo.py:12 imports it and never uses it)."""


def remote(*a, **k):
    raise NotImplementedError
