"""One epoch of experience collection for a batched RadSearch env: the loop of
/root/reference/algos/multiagent/train.py:321-571 (single-agent RAD-A2C branch) with every environment of the batch
stepping at once and NO per-step copies -- the policy reads its observations from the rollout buffer's row t, writes
action / value / log-probability into the same row, and the env step kernel stores the next observation, the reward and
the path-end flags straight into the buffer (RadSearch.step_batch(out=BatchedPPOBuffer.step_outputs(t))).

Caller rules reproduced (T: = train.py):
  * T:394-405  timeout / episode_over / epoch_ended: applied on the device by rs_step (auto_reset, `ended` bits);
  * T:446-487  a trajectory that ends by timeout or at the epoch's last step is bootstrapped with the critic's value of the
               observation that follows (with the recurrent state of that moment); one that only reached the source gets 0;
  * T:484      the epoch's last step asks for new obstructions (epoch_end=True);
  * T:509-511  the recurrent state of an env restarts with its episode (not at the epoch's last step: T:326 resets all);
  * T:359-527  EpRet / EpLen / DoneCount / OutOfBound bookkeeping -> rollout_stats.EpisodeStats.
The policy is the caller's (any torch module); this module only fixes the order of operations around the kernels.
"""
from __future__ import annotations

from typing import Optional, Protocol, Tuple

import torch

from . import _lib as L
from .envs.rad_search_env import RadSearch
from .ppo_buffer import BatchedPPOBuffer
from .rollout_stats import EpisodeStats


class RolloutPolicy(Protocol):
    def act(self, obs: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """obs [N, D] f32 -> (action int32 [N], state value f32 [N], log-probability f32 [N]); advances the policy's own
        recurrent state (T:345-353)."""

    def value(self, obs: torch.Tensor) -> torch.Tensor:
        """State value f32 [N] of `obs` under the current recurrent state, without advancing it (T:470-477)."""

    def reset_state(self, mask: Optional[torch.Tensor]) -> None:
        """Restart the recurrent state of the envs in `mask` (bool [N]; None = all) (T:326, 509-511)."""


class RolloutCollector:
    """Collects `buf.T` steps of every env into `buf` and finishes the trajectories (GAE) -- one epoch of T:321-571."""

    def __init__(self, env: RadSearch, buf: BatchedPPOBuffer, policy: RolloutPolicy, stats: Optional[EpisodeStats] = None):
        if env.number_agents != 1:
            raise ValueError("RolloutCollector drives single-agent envs (the RAD-A2C branch of train.py)")
        if not env.auto_reset:
            raise ValueError("RolloutCollector needs auto_reset=True (the caller rules run on the device)")
        if buf.N != env.num_envs or buf.D != L.OBS_DIM:
            raise ValueError("buffer shape does not match the env batch")
        self.env, self.buf, self.policy, self.stats = env, buf, policy, stats
        self._first = True
        self._src = None

    def collect(self, gae_variant: int = 0) -> None:
        env, buf, pol = self.env, self.buf, self.policy
        T = buf.T
        # the observation that opens the epoch: the env's reset observation the first time, afterwards the one that
        # followed the previous epoch's last step (already in the buffer's row T)
        buf.start_epoch(env.obs if self._first else None)
        self._first = False
        pol.reset_state(None)                                                        # T:326
        src = env.num_envs > 1
        for t in range(T):
            rows = buf.policy_rows(t)
            action, value, logp = pol.act(rows["obs"])
            rows["act"].copy_(action)
            rows["val"].copy_(value)
            rows["logp"].copy_(logp)
            if src:
                rows["src"].copy_(env.src_coords)                                    # T:283-285, 416 (target of the PFGRU)
            last = t == T - 1
            env.step_batch(action, epoch_end=last, out=buf.step_outputs(t))          # obs -> row t+1, reward / ended -> row t
            ended = buf.end_buf[t]
            # T:462-487: bootstrap where the trajectory was cut (timeout, or every env at the epoch's last step)
            v_next = pol.value(env.final_obs.reshape(buf.N, buf.D))
            cut = torch.ones_like(ended, dtype=torch.bool) if last else (ended & L.E_TIMEOUT) != 0
            torch.where(cut, v_next, torch.zeros_like(v_next), out=rows["boot"])
            if self.stats is not None:
                self.stats.update(buf.rew_buf[t].view(buf.N, 1), buf.rew_buf[t], env.done_flags, env.info_flags, ended)
            if not last:
                pol.reset_state(ended != 0)                                          # T:509-511
            buf.advance()
        buf.finish_paths(variant=gae_variant)
