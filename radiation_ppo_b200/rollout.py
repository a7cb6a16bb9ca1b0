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

import ctypes as C
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
        """Restart the recurrent state of the envs in `mask` (bool [N]; None = all) (T:326, 509-511).  A policy that
        exposes its state as `hidden_state` (float32 [N, H], contiguous) has it restarted in place by rs_rollout_post and
        is only called with mask=None."""


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
        # the rows of the (persistent) buffer every step writes: resolved once, not at every step
        self._rows = [buf.policy_rows(t) for t in range(buf.T)]
        self._outs = [buf.step_outputs(t) for t in range(buf.T)]
        self._outs_resolved = [env.resolve_outputs(o) for o in self._outs]

    def collect(self, gae_variant: int = 0) -> None:
        env, buf, pol = self.env, self.buf, self.policy
        T, N = buf.T, buf.N
        lib = L.load()
        stream = C.c_void_p(torch.cuda.current_stream(env.device).cuda_stream)
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())             # noqa: E731
        # the observation that opens the epoch: the env's reset observation the first time, afterwards the one that
        # followed the previous epoch's last step (already in the buffer's row T)
        buf.start_epoch(env.obs if self._first else None)
        self._first = False
        pol.reset_state(None)                                                        # T:326
        hidden = getattr(pol, "hidden_state", None)                                  # [N, H] float32, restarted in place
        st = self.stats.fused_args() if self.stats is not None else (None,) * 5
        final_obs = env.final_obs.reshape(N, buf.D)
        for t in range(T):
            rows = self._rows[t]
            action, value, logp = pol.act(rows["obs"])
            # row t of the buffer <- action, value, log-probability, source coordinates (T:416-428), one launch
            L.check(lib.rs_rollout_pre(p(action), p(value), p(logp), p(env._src), p(rows["act"]), p(rows["val"]),
                                       p(rows["logp"]), p(rows["src"]), N, stream), "rs_rollout_pre")
            last = t == T - 1
            outs = self._outs[t]
            env.step_batch(action, epoch_end=last, out=self._outs_resolved[t])       # obs -> row t+1, reward / ended -> row t
            # T:462-487 bootstrap where the trajectory was cut, T:509-511 recurrent state, T:361-391 episode statistics
            v_next = pol.value(final_obs)
            L.check(lib.rs_rollout_post(p(outs["reward"]), p(outs["ended"]), p(env.done_flags), p(env.info_flags), p(v_next),
                                        p(rows["boot"]), p(hidden), 0 if hidden is None else hidden.shape[1], p(st[0]),
                                        p(st[1]), p(st[2]), p(st[3]), p(st[4]), N, int(last), stream), "rs_rollout_post")
            if hidden is None and not last:
                pol.reset_state(outs["ended"] != 0)                                  # T:509-511
            buf.advance()
        buf.finish_paths(variant=gae_variant)
