"""TEST INFRASTRUCTURE ONLY -- gym.envs.registration.register (gym_rad_search/__init__.py:1-5)."""


def register(id, entry_point, **kwargs):
    import gym

    gym._REGISTRY[id] = entry_point
