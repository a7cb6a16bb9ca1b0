from .rad_search_env import RadSearch, StepResult  # noqa: F401  (mirrors gym_rad_search/envs/__init__.py:1)
