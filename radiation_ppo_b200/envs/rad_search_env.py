"""Batched, GPU-resident RadSearch with the reference's gym API.

Mirrors /root/reference/gym_rad_search/gym_rad_search/envs/rad_search_env.py::RadSearch (constructor kwargs R:320-390,
`reset()` R:730-797, `step(action)` R:443-728, `refresh_environment` R:799-874, attributes read by
algos/multiagent/main.py:538-567 and train.py:279-533).  With ``num_envs == 1`` the methods return the reference's
4-tuples of dicts keyed by agent id; with ``num_envs > 1`` the same names hold CUDA tensors.  All work runs in the
hand-written kernels behind include/radsearch_b200.h; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import gc
import weakref
from typing import Any, Dict, NamedTuple, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from .. import _lib as L

A_SIZE = 9                      # R:66
DETECTABLE_DIRECTIONS = 8       # R:67
DET_STEP = 100.0                # R:68
DET_STEP_FRAC = 71.0            # R:69
DIST_TH = 110.0                 # R:70
EPSILON = 0.0000001             # R:74


class StepResult(NamedTuple):   # R:248-252
    observation: Dict[int, Any]
    reward: Dict[int, float]
    terminal: Dict[int, bool]
    info: Dict[int, Dict[Any, Any]]


class Discrete:
    """Minimal stand-in for gym.spaces.Discrete (R:355)."""

    def __init__(self, n: int):
        self.n = int(n)
        self.shape = ()


class Box:
    """Minimal stand-in for gym.spaces.Box (R:364)."""

    def __init__(self, low, high, shape, dtype=np.float32):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype


class _AgentView:
    """Read-only view of one agent of one environment (Agent dataclass fields R:255-300)."""

    def __init__(self, env: "RadSearch", agent_id: int):
        self._envref, self.id = weakref.ref(env), agent_id      # no reference cycle: an env is freed when dropped

    @property
    def _env(self) -> "RadSearch":
        return self._envref()

    @property
    def det_coords(self) -> Tuple[float, float]:
        d = self._env._det[self.id, 0].tolist()
        return (float(d[0]), float(d[1]))

    @property
    def prev_det_dist(self) -> float:
        return float(self._env._best[self.id, 0].item())

    @property
    def out_of_bounds_count(self) -> int:
        return int(self._env._aflags[self.id, 0].item()) & 0xFFFFFF

    @property
    def obstacle_blocking(self) -> bool:
        return bool((int(self._env._aflags[self.id, 0].item()) >> 24) & 1)


class HostStepBuffers:
    """Pinned host memory for RadSearch.step_host (one block for all outputs, plus the action array)."""

    def __init__(self, env: "RadSearch"):
        self.actions = torch.zeros((env.num_envs, env.number_agents), dtype=torch.int32).pin_memory()
        # with one agent the team reward IS the agent's reward (R:661-665): it stays on the device and `team_reward` is a
        # view of `reward` here -- 4 of 55 bytes per env-step less on the PCIe link that bounds this path
        self.nbytes = env._out_host_bytes
        self.flat = torch.zeros(self.nbytes, dtype=torch.uint8).pin_memory()
        for name, shape, dt, o, nbytes in env._out_layout:
            if o + nbytes <= self.nbytes:
                setattr(self, name, self.flat[o:o + nbytes].view(dt).view(*shape))
        if not hasattr(self, "team_reward"):
            self.team_reward = self.reward.view(env.num_envs)
        self.event = torch.cuda.Event()
        self.h2d_bytes = self.actions.numel() * 4
        self.d2h_bytes = self.nbytes

    def wait(self) -> "HostStepBuffers":
        self.event.synchronize()
        return self


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class RadSearch:
    """RadSearch environment(s) on one GPU.

    Reference kwargs (same names and meaning): ``bbox``, ``observation_area``, ``np_random``, ``obstruction_count``,
    ``enforce_grid_boundaries``, ``number_agents``, ``save_gif`` (ignored), ``DEBUG`` (unsupported).
    Additive kwargs: ``num_envs``, ``device``, ``seed`` (Philox key; drawn from ``np_random`` when omitted),
    ``env_id_offset`` (global id of env 0: results do not depend on how envs are sharded over ranks),
    ``steps_per_episode`` / ``auto_reset`` (the caller rules of train.py:394-405 applied on the device),
    ``fast_poisson`` (fp32 acceptance test in the PTRS sampler), ``count_law`` (0 reference, 1 inverse square),
    ``prefetch`` (with auto_reset: the next episode of every env is prepared ahead of time by ``rs_prepare`` on a side
    stream, so a finished env adopts it inside the step kernel instead of waiting for a reset kernel; results are
    identical either way), ``use_cuda_graph`` (with prefetch: the step / reset / prepare launches are captured once
    and replayed; the Philox step counter then lives on the device), ``standardize`` (0 off; 1 / "radteam": the count
    channel obs[..., 0] is replaced by its per-episode running z-score, StatisticStandardization of
    RADTEAM_core.py:188-277 applied in train.py's order -- update(reading) then standardize(reading), reset with the
    episode; 2 / "statbuff": the older StatBuff rule of algos/test_environment/core.py:55-79 with the clip to [-8, 8];
    the unstandardised counts stay available as ``raw_count``).
    """

    metadata = {"render.modes": ["human"], "video.frames_per_second": 5}
    continuous = False

    def __init__(
        self,
        bbox: Sequence[Sequence[float]] = ((0.0, 0.0), (2700.0, 0.0), (2700.0, 2700.0), (0.0, 2700.0)),
        observation_area: Sequence[float] = (200.0, 500.0),
        np_random: Optional[np.random.Generator] = None,
        obstruction_count: int = 0,
        enforce_grid_boundaries: bool = False,
        save_gif: bool = False,
        number_agents: int = 1,
        DEBUG: bool = False,
        *,
        num_envs: int = 1,
        device: Union[str, torch.device, None] = None,
        seed: Optional[int] = None,
        env_id_offset: int = 0,
        steps_per_episode: int = 120,
        auto_reset: bool = False,
        fast_poisson: bool = False,
        count_law: int = 0,
        k_max: Optional[int] = None,
        prefetch: bool = False,
        use_cuda_graph: bool = False,
        standardize: Union[int, str] = 0,
        prefetch_period: Optional[int] = None,
    ) -> None:
        if DEBUG:
            raise NotImplementedError("the reference's DEBUG hard-codes (R:373-378, 782-784) are not reproduced")
        if not torch.cuda.is_available():
            raise L.RadSearchLibraryError("RadSearch needs a CUDA device: the B200 kernels have no CPU fallback")
        self._lib = L.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if self.device.type != "cuda":
            raise L.RadSearchLibraryError("RadSearch needs a CUDA device: the B200 kernels have no CPU fallback")
        xs = [float(p[0]) for p in bbox]
        ys = [float(p[1]) for p in bbox]
        for v in (*xs, *ys, *observation_area):
            if float(v) != int(v):
                raise ValueError("bbox / observation_area must be integer-valued (SURVEY N1: lattice geometry)")
        self.bbox = tuple((float(p[0]), float(p[1])) for p in bbox)
        self.observation_area = (float(observation_area[0]), float(observation_area[1]))
        self.np_random = np_random if np_random is not None else np.random.default_rng(0)
        self.obstruction_count = int(obstruction_count)
        self.enforce_grid_boundaries = bool(enforce_grid_boundaries)
        self.save_gif = save_gif
        self.number_agents = int(number_agents)
        self.num_envs = int(num_envs)
        self.auto_reset = bool(auto_reset)
        self.fast_poisson = bool(fast_poisson)
        self.env_id_offset = int(env_id_offset)
        if prefetch_period is not None:
            if not 1 <= int(prefetch_period) <= 64:
                raise ValueError("prefetch_period must be in 1..64")
            self.PREFETCH_PERIOD = int(prefetch_period)      # this instance's block length (class default: 4)
        # prefetch lists an env once per block of PREFETCH_PERIOD steps: timed-out episodes must outlast two blocks (a
        # terminal episode lasts >= 9 steps; one that ends before its refill is back takes the synchronous reset)
        self.prefetch = bool(prefetch) and self.auto_reset and int(steps_per_episode) >= 2 * self.PREFETCH_PERIOD
        self.use_cuda_graph = bool(use_cuda_graph) and self.prefetch
        self.seed = int(seed) if seed is not None else int(self.np_random.integers(0, 2**63 - 1))

        cfg = L.RsConfig()
        cfg.bbox[0], cfg.bbox[1], cfg.bbox[2], cfg.bbox[3] = int(min(xs)), int(min(ys)), int(max(xs)), int(max(ys))
        cfg.obs_area[0], cfg.obs_area[1] = int(observation_area[0]), int(observation_area[1])
        cfg.enforce = int(self.enforce_grid_boundaries)
        cfg.n_agents = self.number_agents
        cfg.obstruction_count = self.obstruction_count
        cfg.count_law = int(count_law)
        cfg.max_ep_len = int(steps_per_episode)
        if k_max is None:
            k_max = 5 if self.obstruction_count == -1 else max(self.obstruction_count, 0)
        cfg.k_max = int(k_max)
        self.standardize = {"radteam": 1, "statbuff": 2}.get(standardize, standardize) if isinstance(standardize, str) \
            else int(standardize)
        if self.standardize not in (0, 1, 2):
            raise ValueError("standardize must be 0, 1 ('radteam') or 2 ('statbuff')")
        cfg.standardize = self.standardize
        self._cfg = cfg

        b0x, b0y, b1x, b1y = cfg.bbox[0], cfg.bbox[1], cfg.bbox[2], cfg.bbox[3]
        lo, hi = cfg.obs_area[0], cfg.obs_area[1]
        self.search_area = ((float(b0x + lo), float(b0y + lo)), (float(b1x - hi), float(b0y + lo)),
                            (float(b1x - hi), float(b1y - hi)), (float(b0x + lo), float(b1y - hi)))   # R:393-420
        self.max_dist = float(np.hypot(self.search_area[2][0] - self.search_area[1][0],
                                       self.search_area[2][1] - self.search_area[1][1]))              # R:423-425
        assert self.max_dist > 1000, "Maximum distance available is too small, unable to spawn source and detector 1000 cm apart"
        self.scale = 1 / self.search_area[2][1]                                                        # R:435
        self.scaled_grid_max = (1, 1)
        self.step_size = DET_STEP
        self.action_space = Discrete(A_SIZE)
        self.number_actions = A_SIZE
        self.detectable_directions = DETECTABLE_DIRECTIONS
        self.observation_space = Box(0, np.inf, shape=(L.OBS_DIM,), dtype=np.float32)
        self.background_radiation_bounds = (10, 51)
        self.radiation_intensity_bounds = (1e6, 10e6)
        self.coord_noise = False                           # settable, like the reference's dataclass field (R:570-574)
        self.epoch_end = True                                                                          # R:421
        self.epoch_cnt = 0
        self.iter_count = 0
        self._ctr = 0                      # Philox step counter: one tick per kernel call that draws numbers

        N, A, K = self.num_envs, self.number_agents, cfg.k_max
        dev = self.device
        z = lambda *s, dt=torch.int32: torch.zeros(*s, dtype=dt, device=dev)     # noqa: E731
        self._src, self._rad = z(N, 2), z(N, 2)
        self._rects = z(max(K, 1), N, 4)
        self._meta = z(N)
        self._det = z(A, N, 2)
        self._best = z(A, N, dt=torch.float64)
        self._aflags = z(A, N)
        self._dsrc = z(N, max(4 * K, 1), dt=torch.float64)        # env-major: a unit gathers its own row
        self._vis = z(max(4 * K, 1), N)
        self._status = z(N)
        self._reset_list, self._reset_count = z(N), z(1)
        self._epi = z(N)
        ptrs = [self._src, self._rad, self._rects, self._meta, self._det, self._best, self._aflags, self._dsrc, self._vis,
                self._status, self._reset_list, self._reset_count, self._epi]
        if self.prefetch:
            self._nx_src, self._nx_det, self._nx_rad = z(N, 2), z(N, 2), z(N, 2)
            self._nx_best = z(N, dt=torch.float64)
            self._nx_dsrc = z(N, max(4 * K, 1), dt=torch.float64)
            self._nx_obs = z(N, A, L.OBS_DIM, dt=torch.float32)
            self._nx_seq = z(N)
            self._refill_list, self._refill_count = z(2, N), z(2)
            ptrs += [self._nx_src, self._nx_det, self._nx_rad, self._nx_best, self._nx_dsrc, self._nx_obs, self._nx_seq,
                     self._refill_list, self._refill_count]
        else:
            ptrs += [None] * 9
        self._ctr_dev = z(1, dt=torch.int64)
        ptrs.append(self._ctr_dev)
        if self.standardize:
            self._st_mean, self._st_m2 = z(A, N, dt=torch.float64), z(A, N, dt=torch.float64)
            self.raw_count = z(N, A, dt=torch.float32)
            ptrs += [self._st_mean, self._st_m2, self.raw_count]
        else:
            self._st_mean = self._st_m2 = self.raw_count = None
            ptrs += [None] * 3
        self._ticket = z(1)
        ptrs.append(self._ticket)
        # float lower bounds of dsrc (same rows): the single-agent step kernel prunes its shortest-path search on them
        self._dsf = z(N, max(4 * K, 1), dt=torch.float32)
        self._nx_dsf = z(N, max(4 * K, 1), dt=torch.float32) if self.prefetch else None
        ptrs += [self._dsf, self._nx_dsf]
        self._st = L.RsState(*[None if t is None else t.data_ptr() for t in ptrs])
        self._blk_par, self._blk_pos = 0, 0         # refill list of the current block of steps, position in the block
        self._side = torch.cuda.Stream(device=dev, priority=-1) if self.prefetch else None
        self._ev_main = torch.cuda.Event()
        self._ev_side = [torch.cuda.Event(), torch.cuda.Event()]
        self._pending = [False, False]      # list holds envs waiting for rs_prepare
        self._inflight = [False, False]     # an rs_prepare draining the list is running on the side stream
        self._graphs = {}
        self._block_graphs = {}
        self._refill_dirty = False
        self._act_buf = z(N, A)
        self._act_block = z(self.PREFETCH_PERIOD, N, A) if self.use_cuda_graph else None
        self._ctr_dev_val = 0               # host mirror of *ctr_dev (graph replays advance both)
        # step outputs: views into ONE flat allocation (256-byte aligned segments) so that a host consumer gets them with
        # a single device->host copy (step_host)
        segs = [("obs", (N, A, L.OBS_DIM), torch.float32), ("reward", (N, A), torch.float32),
                ("done_flags", (N, A), torch.uint8), ("info_flags", (N, A), torch.uint8), ("ended", (N,), torch.uint8),
                ("team_reward", (N,), torch.float32)]        # last: not sent to the host when A == 1 (it equals reward)
        self._out_layout, off = [], 0
        for name, shape, dt in segs:
            nbytes = int(np.prod(shape)) * torch.empty((), dtype=dt).element_size()
            self._out_layout.append((name, shape, dt, off, nbytes))
            off += (nbytes + 255) // 256 * 256
        self._out_flat = torch.zeros(off, dtype=torch.uint8, device=dev)
        self._out_host_bytes = self._out_layout[-1][3] if A == 1 else off
        for name, shape, dt, o, nbytes in self._out_layout:
            setattr(self, name, self._out_flat[o:o + nbytes].view(dt).view(*shape))
        self.final_obs = z(N, A, L.OBS_DIM, dt=torch.float32)
        self._host_stream = None
        self.agents = {i: _AgentView(self, i) for i in range(A)}
        self.reset()

    # ------------------------------------------------------------------------------------------------------------
    # raw batched entry points (tensors in, tensors out; asynchronous on the current stream)
    # ------------------------------------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _base_flags(self) -> int:
        return L.F_FAST_POISSON if self.fast_poisson else 0

    def reset_batch(self, mask: Optional[torch.Tensor] = None, new_obstacles: Union[bool, torch.Tensor] = False,
                    uniforms: Optional[torch.Tensor] = None, from_list: bool = False) -> torch.Tensor:
        """Reset the selected envs (all when ``mask`` is None); returns the observation tensor [N, A, 11]."""
        flags = self._base_flags()
        new_mask = None
        if isinstance(new_obstacles, torch.Tensor):
            new_mask = new_obstacles.to(device=self.device, dtype=torch.uint8).contiguous()
        elif new_obstacles:
            flags |= L.F_NEW_OBSTACLES
        if from_list:
            flags |= L.F_RESET_LIST
        m = None if mask is None else mask.to(device=self.device, dtype=torch.uint8).contiguous()
        u = None if uniforms is None else uniforms.to(device=self.device, dtype=torch.float64).contiguous()
        self._ctr += 1
        self._quiesce_prefetch()
        with torch.cuda.device(self.device):
            L.check(self._lib.rs_reset(C.byref(self._cfg), C.byref(self._st), _ptr(m), _ptr(new_mask), _ptr(self.obs),
                                       self.num_envs, self.env_id_offset, self.seed, self._ctr, _ptr(u),
                                       0 if u is None else u.shape[-1], flags, self._stream()), "rs_reset")
            if self.prefetch:
                self._launch_prepare(0, use_list=False)      # next episodes of everybody, against the current obstructions
        return self.obs

    _OUT_SPEC = (("obs", "obs", torch.float32), ("reward", "reward", torch.float32),
                 ("team_reward", "team_reward", torch.float32), ("done", "done_flags", torch.uint8),
                 ("info", "info_flags", torch.uint8), ("ended", "ended", torch.uint8), ("final_obs", "final_obs", torch.float32))

    def _outputs(self, out: Optional[Dict[str, torch.Tensor]]):
        """The seven output tensors of a step: the env's own, or the caller's where `out` names one (same shape and dtype,
        contiguous, on this device) -- e.g. the rows of a rollout buffer, so that the kernel stores straight into them."""
        if not out:
            return (self.obs, self.reward, self.team_reward, self.done_flags, self.info_flags, self.ended, self.final_obs)
        if isinstance(out, tuple):                      # already resolved by resolve_outputs()
            return out
        unknown = set(out) - {k for k, _, _ in self._OUT_SPEC}
        if unknown:
            raise ValueError(f"unknown step outputs {sorted(unknown)}")
        res = []
        for key, attr, dt in self._OUT_SPEC:
            own = getattr(self, attr)
            t = out.get(key)
            if t is None:
                res.append(own)
                continue
            if t.dtype != dt or t.device != own.device or t.numel() != own.numel() or not t.is_contiguous():
                raise ValueError(f"out[{key!r}] must be a contiguous {dt} tensor of {own.numel()} elements on {own.device}")
            res.append(t)
        return tuple(res)

    def resolve_outputs(self, out: Dict[str, torch.Tensor]) -> tuple:
        """Validate an `out` mapping once; the returned tuple can be passed as `out=` to step_batch any number of times
        (a rollout loop resolves its T rows up front instead of at every step)."""
        return self._outputs(dict(out))

    def step_batch(self, actions: Optional[torch.Tensor], epoch_end: bool = False,
                   uniforms: Optional[torch.Tensor] = None, auto_reset: Optional[bool] = None,
                   out: Optional[Dict[str, torch.Tensor]] = None):
        """One step for every env.  actions: int32 CUDA tensor [N] or [N, A] (None = the reference's step(None) probe).
        Returns (obs, reward, team_reward, done, info, ended); with auto-reset, envs that finished have already been
        reset, `obs` holds their first observation and `final_obs` the last one of the finished episode.
        out: tensors that receive the step's outputs instead of the env's own (keys obs, reward, team_reward, done, info,
        ended, final_obs) -- BatchedPPOBuffer.step_outputs(t) hands out its rows this way, so a rollout needs no copies."""
        ar = self.auto_reset if auto_reset is None else auto_reset
        a = None
        if actions is not None:
            a = actions.to(device=self.device, dtype=torch.int32).reshape(self.num_envs, self.number_agents).contiguous()
        u = None if uniforms is None else uniforms.to(device=self.device, dtype=torch.float64).contiguous()
        outs = self._outputs(out)
        self._ctr += 1
        if ar and self.prefetch and a is not None and u is None:
            self._step_prefetch(a, epoch_end, outs if out else None)
            return outs[:6]
        self._quiesce_prefetch()
        o_obs, o_rew, o_team, o_done, o_info, o_ended, o_final = outs
        flags = self._base_flags() | (L.F_AUTO_RESET if ar else 0) | (L.F_EPOCH_END if (ar and epoch_end) else 0)
        with torch.cuda.device(self.device):
            L.check(self._lib.rs_step(C.byref(self._cfg), C.byref(self._st), _ptr(a), _ptr(o_obs), _ptr(o_rew),
                                      _ptr(o_team), _ptr(o_done), _ptr(o_info),
                                      _ptr(o_ended), _ptr(o_final) if ar else None, self.num_envs,
                                      self.env_id_offset, self.seed, self._ctr, _ptr(u),
                                      0 if u is None else u.shape[-1], flags, self._stream()), "rs_step")
            if ar:
                rflags = self._base_flags() | L.F_RESET_LIST | (L.F_NEW_OBSTACLES if epoch_end else 0)
                L.check(self._lib.rs_reset(C.byref(self._cfg), C.byref(self._st), None, None, _ptr(o_obs),
                                           self.num_envs, self.env_id_offset, self.seed, self._ctr, None, 0, rflags,
                                           self._stream()), "rs_reset")
        return outs[:6]

    # ---- prefetch machinery -------------------------------------------------------------------------------------
    # Steps are grouped in blocks of PREFETCH_PERIOD (default 4, `prefetch_period=` per instance: one rs_prepare launch and
    # one CUDA-graph launch per block; longer blocks amortise both); the envs that adopt their prefetched episode during block b are
    # appended to refill list b & 1, and one rs_prepare launch (one thread per env: throughput, not latency) drains that
    # list on a high-priority side stream while block b+1 runs.  An episode lasts >= 9 steps (source and detector
    # start >= 1000 apart, a step is <= 100.4, the goal radius is 110), so the next scenario is back in place in time;
    # if it ever is not, the env simply takes the synchronous reset path -- the resulting state is the same.
    PREFETCH_PERIOD = 4

    def _quiesce_prefetch(self) -> None:
        """Make the main stream wait for any rs_prepare in flight and forget pending refill lists (the envs in them
        take the synchronous reset path next time): called before anything that rewrites env state wholesale."""
        if not self.prefetch:
            return
        if any(self._inflight):
            torch.cuda.current_stream(self.device).wait_stream(self._side)
        if any(self._pending) or self._refill_dirty:
            self._refill_count.zero_()              # nothing stale for an unconditional rs_prepare (block graphs) to redo
            self._refill_dirty = False
        self._inflight = [False, False]
        self._pending = [False, False]
        self._blk_pos = 0

    def prefetch_all(self) -> None:
        """(Re)compute the next episode of EVERY env on the current stream (one rs_prepare launch over all envs).  Done
        automatically after reset / load_scenarios; call it after driving the C entry points directly, or after anything
        else that dropped the pending refill lists, so that the next resets adopt prefetched episodes again instead of
        going through the reset kernel."""
        if not self.prefetch:
            return
        self._quiesce_prefetch()
        with torch.cuda.device(self.device):
            self._launch_prepare(0, use_list=False)

    def _launch_step_sequence(self, a, p: int, epoch_end: bool, device_ctr: bool, first: bool, outs=None) -> None:
        """rs_step (which starts refill list p when `first`) + rs_reset(list) on the current stream; with the device
        step counter the reset kernel's last CTA advances it (RS_F_BUMP_CTR): two launches per step."""
        o_obs, o_rew, o_team, o_done, o_info, o_ended, o_final = outs if outs is not None else self._outputs(None)
        pf = L.F_PREFETCH | (L.F_PARITY1 if p else 0) | (L.F_DEVICE_CTR if device_ctr else 0)
        flags = self._base_flags() | L.F_AUTO_RESET | pf | (L.F_EPOCH_END if epoch_end else 0) | \
            (L.F_ZERO_REFILL if first else 0)
        L.check(self._lib.rs_step(C.byref(self._cfg), C.byref(self._st), _ptr(a), _ptr(o_obs), _ptr(o_rew),
                                  _ptr(o_team), _ptr(o_done), _ptr(o_info),
                                  _ptr(o_ended), _ptr(o_final), self.num_envs, self.env_id_offset, self.seed,
                                  self._ctr, None, 0, flags, self._stream()), "rs_step")
        rflags = self._base_flags() | L.F_RESET_LIST | pf | (L.F_NEW_OBSTACLES if epoch_end else 0) | \
            (L.F_BUMP_CTR if device_ctr else 0)
        L.check(self._lib.rs_reset(C.byref(self._cfg), C.byref(self._st), None, None, _ptr(o_obs), self.num_envs,
                                   self.env_id_offset, self.seed, self._ctr, None, 0, rflags, self._stream()), "rs_reset")
        self._refill_dirty = True

    def _launch_prepare(self, p: int, use_list: bool = True) -> None:
        flags = self._base_flags() | (L.F_REFILL_LIST if use_list else 0) | (L.F_PARITY1 if p else 0)
        L.check(self._lib.rs_prepare(C.byref(self._cfg), C.byref(self._st), self.num_envs, self.env_id_offset, self.seed,
                                     flags, self._stream()), "rs_prepare")

    def _step_prefetch(self, a: torch.Tensor, epoch_end: bool, outs=None) -> None:
        dev = self.device
        main = torch.cuda.current_stream(dev)
        p, first = self._blk_par, self._blk_pos == 0
        with torch.cuda.device(dev):
            if epoch_end:
                # new obstructions for everybody: nothing may be preparing scenarios against the old ones
                self._quiesce_prefetch()
                self._launch_step_sequence(a, p, True, False, True, outs)
                self._pending[p] = True                 # the reset pushed every env to list p
                self._blk_par, self._blk_pos = p ^ 1, 0
                return
            if first:
                if self._pending[p ^ 1]:                # the previous block's list -> rs_prepare on the side stream
                    self._ev_main.record(main)
                    self._side.wait_event(self._ev_main)
                    with torch.cuda.stream(self._side):
                        self._launch_prepare(p ^ 1)
                        self._ev_side[p ^ 1].record(self._side)
                    self._pending[p ^ 1], self._inflight[p ^ 1] = False, True
                if self._inflight[p]:                   # list p is about to be reused: its rs_prepare must be done
                    main.wait_event(self._ev_side[p])
                    self._inflight[p] = False
            if self.use_cuda_graph and outs is None:                     # (a captured graph stores to the env's own outputs)
                if a.data_ptr() != self._act_buf.data_ptr():             # the caller may fill action_buffer in place
                    self._act_buf.copy_(a)
                key = (p, first)
                g = self._graphs.get(key)
                if g is None:
                    g = self._graphs[key] = self._capture(p, first)
                if self._ctr_dev_val != self._ctr:
                    self._ctr_dev.fill_(self._ctr)
                    self._reset_count.zero_()           # a stream-launched step may have left entries behind
                g.replay()
                self._ctr_dev_val = self._ctr + 1
            else:
                self._launch_step_sequence(a, p, False, False, first, outs)
            self._pending[p] = True
            self._blk_pos += 1
            if self._blk_pos == self.PREFETCH_PERIOD:
                self._blk_par, self._blk_pos = p ^ 1, 0

    def capture_graphs(self) -> None:
        """Capture the four step-graph variants (refill list 0/1 x first-step-of-block) and the two block graphs ahead of
        time."""
        if not self.use_cuda_graph:
            return
        for p in (0, 1):
            for first in (False, True):
                if (p, first) not in self._graphs:
                    self._graphs[(p, first)] = self._capture(p, first)
            if p not in self._block_graphs:
                self._block_graphs[p] = self._capture_block(p)

    @property
    def action_block(self) -> torch.Tensor:
        """int32 [PREFETCH_PERIOD, N, A] device buffer step_block reads its actions from."""
        return self._act_block

    def step_block(self, actions: Optional[torch.Tensor] = None):
        """PREFETCH_PERIOD consecutive steps (+ auto-resets) replayed as ONE CUDA graph: the 2 x PERIOD kernel launches
        of the block on the current stream and, on a parallel branch, the rs_prepare that refills the next episodes of the
        envs that finished during the previous block.  For callers that have the block's actions up front (open-loop
        rollouts, Monte-Carlo evaluation with a scripted policy, throughput measurement): one graph launch per PERIOD steps
        instead of one per step.  actions: int32 [PERIOD, N(, A)] (None: action_block was filled in place).  The outputs
        left in obs / reward / ... are those of the block's last step.  Same results as PERIOD step_batch calls."""
        if not self.use_cuda_graph:
            raise ValueError("step_block needs auto_reset=True, prefetch=True, use_cuda_graph=True")
        if self._blk_pos != 0:
            raise ValueError("step_block must start at a block boundary (a multiple of PREFETCH_PERIOD steps since the "
                             "last reset / epoch end)")
        dev, P = self.device, self.PREFETCH_PERIOD
        main = torch.cuda.current_stream(dev)
        p = self._blk_par
        with torch.cuda.device(dev):
            if actions is not None and actions.data_ptr() != self._act_block.data_ptr():
                self._act_block.copy_(actions.reshape(P, self.num_envs, self.number_agents))
            if any(self._inflight):                     # a stream-launched rs_prepare: order it before the graph
                main.wait_stream(self._side)
                self._inflight = [False, False]
            g = self._block_graphs.get(p)
            if g is None:
                g = self._block_graphs[p] = self._capture_block(p)
            if self._ctr_dev_val != self._ctr + 1:
                self._ctr_dev.fill_(self._ctr + 1)
                self._reset_count.zero_()
            g.replay()
        self._ctr += P
        self._ctr_dev_val = self._ctr + 1
        self._pending[p], self._pending[p ^ 1] = True, False      # the graph's rs_prepare drained the other list
        self._refill_dirty = True
        self._blk_par = p ^ 1
        return self.obs, self.reward, self.team_reward, self.done_flags, self.info_flags, self.ended

    def _capture_block(self, p: int):
        """{rs_prepare(list p^1) on a side branch || PERIOD x (rs_step -> rs_reset(list p))} captured once per parity."""
        dev = self.device
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        cap = torch.cuda.Stream(device=dev)
        fork, join = torch.cuda.Event(), torch.cuda.Event()
        ctr = self._ctr
        gc_was_on = gc.isenabled()
        gc.disable()
        try:
            with torch.cuda.stream(cap):
                with torch.cuda.graph(g, stream=cap, capture_error_mode="relaxed"):
                    fork.record(cap)
                    self._side.wait_event(fork)
                    with torch.cuda.stream(self._side):
                        self._launch_prepare(p ^ 1)
                        join.record(self._side)
                    for i in range(self.PREFETCH_PERIOD):
                        self._launch_step_sequence(self._act_block[i], p, False, True, i == 0)
                    cap.wait_event(join)
        finally:
            if gc_was_on:
                gc.enable()
        self._ctr = ctr
        self._ctr_dev_val = -1
        return g

    def _capture(self, p: int, first: bool):
        """Capture {rs_step -> rs_reset(list), whose last CTA bumps the device step counter} once; replayed per step."""
        dev = self.device
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        cap = torch.cuda.Stream(device=dev)
        # the region holds two kernel launches and nothing else; "relaxed" + no cyclic GC inside it, so that a finalizer of
        # some unrelated object (say another env's graphs being destroyed) cannot invalidate the capture
        gc_was_on = gc.isenabled()
        gc.disable()
        try:
            with torch.cuda.stream(cap):
                with torch.cuda.graph(g, stream=cap, capture_error_mode="relaxed"):
                    self._launch_step_sequence(self._act_buf, p, False, True, first)
        finally:
            if gc_was_on:
                gc.enable()
        self._ctr_dev_val = -1                           # capture does not execute; force a refresh before the replay
        return g

    # ---- host-facing step: pinned host buffers in and out, asynchronous ------------------------------------------------
    def host_buffers(self) -> "HostStepBuffers":
        """Pinned host buffers for step_host: `.actions` int32 [N, A] to fill, and the step outputs as views of one
        pinned block (`.obs`, `.reward`, `.team_reward`, `.done_flags`, `.info_flags`, `.ended`)."""
        return HostStepBuffers(self)

    def step_host(self, hb: "HostStepBuffers", epoch_end: bool = False,
                  actions: Optional[torch.Tensor] = None) -> "HostStepBuffers":
        """One step driven from the host: copies hb.actions host->device, runs the step (+ auto-reset), copies all step
        outputs device->host in ONE transfer into hb, all asynchronously on this env's own stream; `hb.wait()` blocks
        until the results are in host memory (`actions`: another pinned int32 [N, A] tensor to send instead of
        hb.actions).  Independent env batches stepped this way overlap their copies with each
        other's kernels (the copy engines and the SMs work concurrently)."""
        if self._host_stream is None:
            self._host_stream = torch.cuda.Stream(device=self.device)
        # whatever the caller's stream did to this env before (reset, step_batch, load_scenarios) is ordered first
        self._host_stream.wait_stream(torch.cuda.current_stream(self.device))
        h_act = hb.actions if actions is None else actions.reshape(self.num_envs, self.number_agents)
        with torch.cuda.stream(self._host_stream):
            if self.use_cuda_graph and self.auto_reset and not epoch_end:
                self._act_buf.copy_(h_act, non_blocking=True)            # straight into the captured graph's action buffer
                acts = self._act_buf
            else:
                acts = h_act.to(self.device, non_blocking=True)
            self.step_batch(acts, epoch_end=epoch_end)
            hb.flat.copy_(self._out_flat[:hb.nbytes], non_blocking=True)
            hb.event.record(self._host_stream)
        return hb

    def load_scenarios(self, src, det, intensity, bkg, rects=None, num_obs=None,
                       uniforms: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Batched refresh_environment (R:799-874): inject scenarios instead of sampling them.  rects: [N, K, 4] as
        x0,y0,x1,y1 (K <= k_max), num_obs: [N].  Returns the initial observations."""
        dev = self.device
        ti = lambda x, shape: torch.as_tensor(np.asarray(x) if not isinstance(x, torch.Tensor) else x).to(   # noqa: E731
            device=dev, dtype=torch.int32).reshape(shape).contiguous()
        N = self.num_envs
        s, d = ti(src, (N, 2)), ti(det, (N, 2))
        it, bk = ti(intensity, (N,)), ti(bkg, (N,))
        if rects is None or self._cfg.k_max == 0:
            r, k_in = None, 0
            no = torch.zeros(N, dtype=torch.int32, device=dev)
        else:
            r = torch.as_tensor(np.asarray(rects) if not isinstance(rects, torch.Tensor) else rects).to(
                device=dev, dtype=torch.int32).contiguous()
            k_in = int(r.shape[1])
            no = ti(num_obs, (N,))
        u = None if uniforms is None else uniforms.to(device=dev, dtype=torch.float64).contiguous()
        self._ctr += 1
        self._quiesce_prefetch()
        with torch.cuda.device(dev):
            L.check(self._lib.rs_load_scenarios(C.byref(self._cfg), C.byref(self._st), _ptr(s), _ptr(d), _ptr(it),
                                                _ptr(bk), _ptr(r), k_in, _ptr(no), _ptr(self.obs), N,
                                                self.env_id_offset, self.seed, self._ctr, _ptr(u),
                                                0 if u is None else u.shape[-1], self._stream()), "rs_load_scenarios")
            if self.prefetch:
                self._launch_prepare(0, use_list=False)
        self.epoch_end = False
        return self.obs

    def scenario_arrays(self) -> Dict[str, np.ndarray]:
        """The current scenario of every env (source, agent 0's detector, intensities, obstructions) as the arrays of
        scenario_io.scenario_arrays: what create_envs (algos/test_environment/eval/test_env_gen.py:13-24) saves after
        env.reset().  With scenario_io.to_env_dict / save_test_env_dict this writes the reference's test-environment files
        from one batched reset instead of one gym.make + reset per scenario."""
        K = self._cfg.k_max
        return dict(src=self._src.cpu().numpy().astype(np.int32), det=self._det[0].cpu().numpy().astype(np.int32),
                    intensity=self._rad[:, 0].cpu().numpy().astype(np.int32), bkg=self._rad[:, 1].cpu().numpy().astype(np.int32),
                    rects=self._rects[:max(K, 1)].permute(1, 0, 2).contiguous().cpu().numpy().astype(np.int32),
                    num_obs=(self._meta & 0xFF).cpu().numpy().astype(np.int32))

    def shortest_path_to(self, points: torch.Tensor, variant: int = 0) -> torch.Tensor:
        """Shortest-path length from each env's source to points[n] (int tensor [N, 2]) around its obstructions
        (world.shortest_path(...).length(), R:491-493) -> float64 tensor [N]."""
        pts = points.to(device=self.device, dtype=torch.int32).reshape(self.num_envs, 2).contiguous()
        out = torch.empty(self.num_envs, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            L.check(self._lib.rs_query_shortest_path(C.byref(self._cfg), C.byref(self._st), _ptr(pts), _ptr(out),
                                                     self.num_envs, int(variant), self._stream()), "rs_query_shortest_path")
        return out

    # ------------------------------------------------------------------------------------------------------------
    # state views
    # ------------------------------------------------------------------------------------------------------------
    @property
    def action_buffer(self) -> torch.Tensor:
        """int32 [N, A] device buffer the captured step graph reads its actions from: a policy that writes its sampled
        actions here and passes this tensor to step_batch saves the device-to-device copy."""
        return self._act_buf

    @property
    def status(self) -> torch.Tensor:
        return self._status

    @property
    def num_obs(self):
        v = self._meta & 0xFF
        return int(v[0].item()) if self.num_envs == 1 else v

    @property
    def done(self):
        v = ((self._meta >> 8) & 1).bool()
        return bool(v[0].item()) if self.num_envs == 1 else v

    @property
    def steps_in_episode(self) -> torch.Tensor:
        return self._meta >> 16

    @property
    def src_coords(self):
        if self.num_envs == 1:
            s = self._src[0].tolist()
            return (float(s[0]), float(s[1]))
        return self._src

    @property
    def intensity(self):
        return int(self._rad[0, 0].item()) if self.num_envs == 1 else self._rad[:, 0]

    @property
    def bkg_intensity(self):
        return int(self._rad[0, 1].item()) if self.num_envs == 1 else self._rad[:, 1]

    @property
    def det_coords(self) -> torch.Tensor:
        """[A, N, 2] detector coordinates."""
        return self._det

    @property
    def obs_coord(self):
        """Obstruction vertex lists [(x,y),(x,y+h),(x+w,y+h),(x+w,y)] of env 0 (R:975-983); batched: rects [K,N,4]."""
        if self.num_envs > 1:
            return self._rects
        out = []
        r = self._rects[:, 0].tolist()
        for k in range(self.num_obs):
            x0, y0, x1, y1 = (float(v) for v in r[k])
            out.append([(x0, y0), (x0, y1), (x1, y1), (x1, y0)])
        return out

    def get_agent_outOfBounds_count(self, id: int) -> int:      # R:1764
        return self.agents[id].out_of_bounds_count

    def ping(self):
        return "PONG"

    def render(self, *args, **kwargs):
        """Rendering (R:1308-1762) is out of scope for the B200 path."""
        return None

    # ------------------------------------------------------------------------------------------------------------
    # gym API
    # ------------------------------------------------------------------------------------------------------------
    def _noisy(self, obs):
        """R:570-574 `coord_noise`: N(0, 5) on the detector coordinates of the observation (not of the state), scaled like
        them.  Drawn from `np_random` for one env (per agent, in agent order, as the reference does), from torch's
        generator for a batch."""
        if not self.coord_noise:
            return obs
        if self.num_envs == 1:
            for i in range(self.number_agents):
                obs[i][1:3] += self.np_random.normal(scale=5, size=2) * self.scale
            return obs
        out = obs.clone()
        out[..., 1:3] += torch.randn_like(out[..., 1:3]) * (5.0 * self.scale)
        return out

    def _pack(self, only=None):
        A = self.number_agents
        if self.num_envs > 1:
            info = {
                "out_of_bounds": (self.info_flags & L.I_OOB) != 0,
                "out_of_bounds_count": (self._aflags & 0xFFFFFF).transpose(0, 1),
                "blocked": (self.info_flags & L.I_BLOCKED) != 0,
                "collision": (self.info_flags & L.I_COLLISION) != 0,
                "scale": self.scale,
                "ended": self.ended,
                "final_observation": self.final_obs,
            }
            return (self._noisy(self.obs), {"team_reward": self.team_reward, "individual_reward": self.reward},
                    self.done_flags != 0, info)
        obs = self.obs[0].double().cpu().numpy()
        rew = self.reward[0].double().cpu().numpy()
        team = float(self.team_reward[0].item())
        done = self.done_flags[0].cpu().numpy()
        inf = self.info_flags[0].cpu().numpy()
        oobc = (self._aflags[:, 0] & 0xFFFFFF).cpu().numpy()
        r2 = lambda v: round(float(v), 2)                                   # noqa: E731  rewards are 2-decimal values
        ids = range(A) if only is None else only
        # agents a dict action did not name were not stepped by the reference: their entries stay None (R:627-630)
        pick = lambda d: {i: (d[i] if i in ids else None) for i in range(A)}   # noqa: E731
        rewards = [r2(rew[i]) for i in ids]
        if only is not None:                                                # team reward = max over the stepped agents R:661-665
            team = None
            for r in rewards:
                team = r if not team else max(team, r)
        return (
            pick(self._noisy({i: obs[i].copy() for i in range(A)})),
            {"team_reward": r2(team) if team is not None else None, "individual_reward": pick({i: r2(rew[i]) for i in range(A)})},
            pick({i: bool(done[i]) for i in range(A)}),
            pick({i: {"out_of_bounds": bool(inf[i] & L.I_OOB), "out_of_bounds_count": int(oobc[i]),
                      "blocked": bool(inf[i] & L.I_BLOCKED), "scale": self.scale} for i in range(A)}),
        )

    def reset(self):
        """R:730-797.  New obstructions are drawn only when ``epoch_end`` was set (train.py:484)."""
        new_obs = bool(self.epoch_end)
        self.reset_batch(mask=None, new_obstacles=new_obs)
        if new_obs:
            self.epoch_cnt += 1
            self.epoch_end = False
        self.iter_count = 0
        self.done_flags.zero_()
        self.info_flags.zero_()
        self.reward.zero_()
        self.team_reward.zero_()
        out = self._pack()
        if self.num_envs == 1:
            # the probe step of R:794 reports round(-0.5 * sp / max_dist, 2) with sp = prev_det_dist (R:551-567)
            ind = {i: round(-0.5 * self.agents[i].prev_det_dist / self.max_dist, 2) for i in range(self.number_agents)}
            team = None
            for r in ind.values():                                       # R:661-665
                if not team:
                    team = r
                elif team < r:
                    team = r
            out = (out[0], {"team_reward": team, "individual_reward": ind}, out[2], out[3])
        return out

    def step(self, action=None):
        """R:443-728.  ``action``: int (applied to every agent; -1 = idle), dict {agent_id: action}, None (probe), or for
        ``num_envs > 1`` an integer tensor [N] / [N, A]."""
        A = self.number_agents
        only = None
        if action is None:
            acts = None
        elif isinstance(action, torch.Tensor):
            acts = action
        elif isinstance(action, dict):
            if not action or any(i not in range(A) for i in action):
                raise ValueError("action dict must be keyed by agent ids 0..number_agents-1")
            for a in action.values():
                assert int(a) in range(A_SIZE)
            # R:645-659 steps only the agents the dict names; the others are probed without moving (action -1, which the
            # kernel treats like step(None) for that agent) and reported as None.  Difference to the reference: a named
            # agent that moves onto the cell of an unnamed one counts as a collision here.
            if len(action) < A:
                only = sorted(action)
            acts = torch.tensor([[int(action.get(i, -1)) for i in range(A)]] * self.num_envs, dtype=torch.int32)
        elif isinstance(action, (int, np.integer)):
            a = 8 if int(action) == -1 else int(action)                   # R:620-623
            assert a in range(A_SIZE)
            acts = torch.full((self.num_envs, A), a, dtype=torch.int32)
        else:
            raise ValueError("Incompatible Action type")
        self.step_batch(acts)
        if acts is not None:
            self.iter_count += 1
        return self._pack(only)

    def refresh_environment(self, env_dict: Dict, id: int, num_obs: int = 0):
        """R:799-874: load scenario ``env_<id>`` of a saved test-environment dict into every env of this instance."""
        from ..scenario_io import scenario_arrays

        arr = scenario_arrays(env_dict, [id] * self.num_envs, k_max=self._cfg.k_max, with_obstacles=num_obs > 0)
        self.load_scenarios(**arr)
        self.iter_count = 1
        self.done_flags.zero_()
        obs = self._pack()[0]
        return obs
