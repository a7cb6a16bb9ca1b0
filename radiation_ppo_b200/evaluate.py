"""Batched Monte-Carlo evaluation over saved test scenarios.

Mirrors the measurement of /root/reference/algos/multiagent/evaluate.py::EpisodeRunner.run (E:333-475): every test
scenario (`env_<i>` of a `test_env_dict_*_v4` file, E:203, loaded by `refresh_environment`, rad_search_env.py:799-874) is
played `montecarlo_runs` times by a policy until an agent reaches the source or `steps_per_episode` steps have passed;
per scenario it reports the success count, the episode lengths and the episode returns of the successful and the
unsuccessful runs, and the source / background intensities (E:428-450, `MonteCarloResults` E:88-103).

The reference runs the scenarios one after the other, one environment, one Monte-Carlo run at a time.  Here all
scenarios x runs are ONE batched environment (100 runs x 1000 scenarios = 100,000 envs in a launch): `rs_load_scenarios`
builds them, `rs_step` advances them, and the bookkeeping is a handful of tensor ops per step.  The Monte-Carlo runs
of a scenario differ in their Poisson draws (Philox keyed by the env id) and in whatever the policy samples.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, Iterable, Optional

import numpy as np
import torch

from .envs.rad_search_env import RadSearch
from .scenario_io import scenario_arrays


@dataclass
class MonteCarloResults:
    """Per-scenario results, arrays of shape [scenarios, runs] unless noted (cf. E:88-103)."""

    success: torch.Tensor                  # bool: a terminal state was reached before the timeout
    episode_length: torch.Tensor           # int32 (`total_episode_length`)
    episode_return: torch.Tensor           # float64 sum of the float32 (team) rewards, agent 0's tally (E:402-409, 441)
    success_counter: torch.Tensor          # int32 [scenarios]
    intensity: torch.Tensor                # int32 [scenarios]
    background_intensity: torch.Tensor     # int32 [scenarios]
    completed_runs: int = 0
    extra: Dict = field(default_factory=dict)

    def summary(self) -> Dict[str, float]:
        s = self.success
        f = lambda x, m: float(x[m].float().median()) if bool(m.any()) else float("nan")     # noqa: E731
        return {"success_rate": float(s.float().mean()), "median_len_success": f(self.episode_length, s),
                "median_len_fail": f(self.episode_length, ~s), "median_ret_success": f(self.episode_return, s),
                "median_ret_fail": f(self.episode_return, ~s)}


Policy = Callable[[torch.Tensor, int], torch.Tensor]     # (obs [N, A, 11] on the device, step index) -> actions [N, A]


def uniform_policy(seed: int = 0) -> Policy:
    """The reference's `uniform_search` baseline: actions drawn uniformly from the 8 directions."""
    gen: Dict[str, torch.Generator] = {}

    def act(obs: torch.Tensor, t: int) -> torch.Tensor:
        g = gen.setdefault("g", torch.Generator(device=obs.device).manual_seed(seed))
        return torch.randint(0, 8, obs.shape[:2], generator=g, device=obs.device, dtype=torch.int32)

    return act


class MonteCarloEvaluator:
    """All scenarios x Monte-Carlo runs as one batched RadSearch.

    env_dict: the joblib dict of a `test_env_dict_*` file (or None with `scenarios` = the arrays of
    scenario_io.scenario_arrays); ids: which scenarios; env_kwargs: the reference's env kwargs
    (`obstruction_count`, `enforce_grid_boundaries`, `number_agents`, ...).  Env n = scenario n // runs, run n % runs."""

    def __init__(self, env_dict: Optional[Dict] = None, ids: Optional[Iterable[int]] = None, montecarlo_runs: int = 100,
                 steps_per_episode: int = 120, obstruction_count: int = 0, scenarios: Optional[Dict[str, np.ndarray]] = None,
                 seed: int = 0, device=None, team_mode: str = "cooperative", standardize: int = 0, **env_kwargs) -> None:
        k_max = max(int(obstruction_count), 0)
        if scenarios is None:
            if env_dict is None:
                raise ValueError("give env_dict or scenarios")
            scenarios = scenario_arrays(env_dict, ids, k_max=max(k_max, 1), with_obstacles=obstruction_count > 0)
        self.n_scenarios = int(len(scenarios["intensity"]))
        self.runs, self.steps_per_episode, self.team_mode = int(montecarlo_runs), int(steps_per_episode), team_mode
        N = self.n_scenarios * self.runs
        self.env = RadSearch(obstruction_count=obstruction_count, num_envs=N, seed=seed, device=device,
                             steps_per_episode=steps_per_episode, k_max=k_max, standardize=standardize, **env_kwargs)
        rep = lambda a: np.repeat(np.asarray(a), self.runs, axis=0)          # noqa: E731
        self._arrays = {k: rep(v) for k, v in scenarios.items()}
        if k_max > 0:
            if int(np.max(scenarios["num_obs"])) > k_max:
                raise ValueError("a scenario has more obstructions than obstruction_count")
            self._arrays["rects"] = np.ascontiguousarray(self._arrays["rects"][:, :k_max])
        self.intensity = torch.as_tensor(np.asarray(scenarios["intensity"]), dtype=torch.int32, device=self.env.device)
        self.background = torch.as_tensor(np.asarray(scenarios["bkg"]), dtype=torch.int32, device=self.env.device)

    def run(self, policy: Policy, on_step: Optional[Callable[[int, torch.Tensor, torch.Tensor], None]] = None) -> MonteCarloResults:
        env, S, R, A = self.env, self.n_scenarios, self.runs, self.env.number_agents
        N = S * R
        dev = env.device
        a = self._arrays
        obs = env.load_scenarios(a["src"], a["det"], a["intensity"], a["bkg"],
                                 rects=a["rects"] if env._cfg.k_max > 0 else None, num_obs=a["num_obs"])
        active = torch.ones(N, dtype=torch.bool, device=dev)
        success = torch.zeros(N, dtype=torch.bool, device=dev)
        length = torch.zeros(N, dtype=torch.int32, device=dev)
        ret = torch.zeros(N, dtype=torch.float64, device=dev)
        for t in range(self.steps_per_episode):
            acts = policy(obs, t).to(device=dev, dtype=torch.int32).reshape(N, A)
            obs, reward, team, done, _, _ = env.step_batch(acts, auto_reset=False)
            r = team if self.team_mode != "individual" else reward[:, 0]      # E:402-409 (agent 0's tally is reported)
            ret += torch.where(active, r, torch.zeros_like(r)).double()
            length += active.to(torch.int32)
            terminal = (done != 0).any(dim=1) & active                          # E:415-419
            success |= terminal
            active &= ~terminal
            if on_step is not None:
                on_step(t, obs, active)
            if t % 8 == 7 and not bool(active.any()):                           # every run has finished early
                break
        return MonteCarloResults(
            success=success.view(S, R), episode_length=length.view(S, R), episode_return=ret.view(S, R),
            success_counter=success.view(S, R).sum(dim=1).to(torch.int32), intensity=self.intensity,
            background_intensity=self.background, completed_runs=R)
