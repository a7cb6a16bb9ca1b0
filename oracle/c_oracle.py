"""TEST INFRASTRUCTURE ONLY -- ctypes/numpy front-end of the C restatement (oracle/radsearch_oracle.c).

Used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the checker and the
CPU baseline.  The product package never imports it.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build_oracle

MAX_K, MAX_A, OBS_DIM = 8, 8, 11

ENV_DTYPE = np.dtype(
    [
        ("num_obs", np.int32),
        ("rect", np.int32, (MAX_K, 4)),
        ("src", np.int32, (2,)),
        ("intensity", np.int32),
        ("bkg", np.int32),
        ("det", np.int32, (MAX_A, 2)),
        ("best", np.float64, (MAX_A,)),
        ("sp", np.float64, (MAX_A,)),
        ("euc", np.float64, (MAX_A,)),
        ("oob", np.int32, (MAX_A,)),
        ("oob_count", np.int32, (MAX_A,)),
        ("blocked", np.int32, (MAX_A,)),
        ("collision", np.int32, (MAX_A,)),
        ("los_blocked", np.int32, (MAX_A,)),
        ("done", np.int32),
        ("iter_count", np.int32),
        ("ep_len", np.int32),
        ("status", np.uint32),
        ("episode", np.uint32),
    ],
    align=True,
)

OUT_DTYPE = np.dtype(
    [
        ("obs", np.float64, (MAX_A, OBS_DIM)),
        ("reward", np.float64, (MAX_A,)),
        ("team_reward", np.float64),
        ("done", np.int32, (MAX_A,)),
        ("lam", np.float64, (MAX_A,)),
    ],
    align=True,
)


class Config(C.Structure):
    _fields_ = [
        ("bbox", C.c_int32 * 4),
        ("obs_area", C.c_int32 * 2),
        ("enforce", C.c_int32),
        ("n_agents", C.c_int32),
        ("obstruction_count", C.c_int32),
        ("count_law", C.c_int32),
        ("max_ep_len", C.c_int32),
    ]


class Rng(C.Structure):
    _fields_ = [
        ("u", C.POINTER(C.c_double)),
        ("n_u", C.c_int32),
        ("pos", C.c_int32),
        ("key", C.c_uint32 * 2),
        ("ctr", C.c_uint32 * 4),
        ("buf", C.c_uint32 * 4),
        ("have", C.c_int32),
        ("status", C.POINTER(C.c_uint32)),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = build_oracle.LIB
        if not os.path.exists(path) or os.path.exists(build_oracle.SRC):
            path = build_oracle.build()
        _lib = C.CDLL(path)
        _lib.orc_poisson.restype = C.c_int64
        _lib.orc_poisson.argtypes = [C.POINTER(Rng), C.c_double]
        _lib.orc_loggam.restype = C.c_double
        _lib.orc_loggam.argtypes = [C.c_double]
        _lib.orc_round2.restype = C.c_double
        _lib.orc_round2.argtypes = [C.c_double]
        _lib.orc_rng_double.restype = C.c_double
        _lib.orc_rng_u32.restype = C.c_uint32
        _lib.orc_rng_below.restype = C.c_uint32
        _lib.orc_shortest_path.restype = C.c_double
        _lib.orc_rollout.restype = C.c_int64
        assert _lib.orc_sizeof_env() == ENV_DTYPE.itemsize, (_lib.orc_sizeof_env(), ENV_DTYPE.itemsize)
        assert _lib.orc_sizeof_out() == OUT_DTYPE.itemsize, (_lib.orc_sizeof_out(), OUT_DTYPE.itemsize)
    return _lib


def default_config(**kw) -> Config:
    c = Config()
    lib().orc_default_config(C.byref(c))
    for k, v in kw.items():
        if k in ("bbox", "obs_area"):
            for i, x in enumerate(v):
                getattr(c, k)[i] = int(x)
        else:
            setattr(c, k, int(v))
    return c


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t) if a is not None else None


def i32(*v):
    return (C.c_int32 * len(v))(*[int(x) for x in v])


def philox(ctr, key):
    out = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out)
    return list(out)


def poisson_injected(lam: float, uniforms: np.ndarray):
    """numpy Generator.poisson(lam) replayed on an explicit stream of next_double() values -> (count, status)."""
    u = np.ascontiguousarray(uniforms, dtype=np.float64)
    st = C.c_uint32(0)
    r = Rng()
    lib().orc_rng_inject(C.byref(r), _p(u, C.POINTER(C.c_double)), C.c_int32(len(u)), C.byref(st))
    k = lib().orc_poisson(C.byref(r), C.c_double(lam))
    return int(k), int(st.value)


def poisson_philox(lam: float, seed: int, env_id: int, domain: int, agent: int, step_ctr: int) -> int:
    st = C.c_uint32(0)
    r = Rng()
    lib().orc_rng_philox(C.byref(r), C.c_uint64(seed), C.c_uint32(env_id), C.c_uint32(domain), C.c_uint32(agent),
                         C.c_uint64(step_ctr), C.byref(st))
    return int(lib().orc_poisson(C.byref(r), C.c_double(lam)))


def round2(x: float) -> float:
    return float(lib().orc_round2(C.c_double(x)))


def seg_hits_open_rect(p, q, r) -> bool:
    return bool(lib().orc_seg_hits_open_rect(i32(*p), i32(*q), i32(*r)))


def seg_touches_seg(a, b, c, d) -> bool:
    return bool(lib().orc_seg_touches_seg(i32(*a), i32(*b), i32(*c), i32(*d)))


def los_blocked_rect(p, q, r) -> bool:
    return bool(lib().orc_los_blocked_rect(i32(*p), i32(*q), i32(*r)))


class OracleBatch:
    """N scalar environments stepped one by one (OpenMP over environments), reference semantics."""

    def __init__(self, n: int, cfg: Config | None = None, seed: int = 0, env_id0: int = 0, threads: int = 0):
        self.cfg = cfg if cfg is not None else default_config()
        self.n = int(n)
        self.A = int(self.cfg.n_agents)
        self.seed, self.env_id0, self.threads = int(seed), int(env_id0), int(threads)
        self.envs = np.zeros(self.n, dtype=ENV_DTYPE)
        self.outs = np.zeros(self.n, dtype=OUT_DTYPE)

    def load_scenarios(self, src, det, intensity, bkg, rects, num_obs):
        """refresh_environment (rad_search_env.py:799-874) for every env; rects [n, K, 4] as x0,y0,x1,y1."""
        L = lib()
        rects = np.ascontiguousarray(rects, dtype=np.int32)
        for i in range(self.n):
            L.orc_load_scenario(C.byref(self.cfg), C.c_void_p(self.envs[i : i + 1].ctypes.data), i32(*src[i]),
                                i32(*det[i]), C.c_int32(int(intensity[i])), C.c_int32(int(bkg[i])),
                                _p(rects[i]), C.c_int32(int(num_obs[i])))

    def reset(self, mask=None, new_obstacles=None, uniforms=None):
        """reset() for the selected envs; draws keyed by (seed, env id, episode number)."""
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        no = None if new_obstacles is None else np.ascontiguousarray(new_obstacles, dtype=np.uint8)
        u = None if uniforms is None else np.ascontiguousarray(uniforms, dtype=np.float64)
        n_inj = 0 if u is None else u.shape[-1]
        lib().orc_reset_batch(C.byref(self.cfg), _p(self.envs), C.c_int32(self.n), _p(m), _p(no), C.c_uint64(self.seed),
                              C.c_uint32(self.env_id0), _p(u), C.c_int32(n_inj), _p(self.outs),
                              C.c_int32(self.threads))
        return self.outs

    def step(self, actions, step_ctr: int, uniforms=None):
        """actions: int array [n, A] in 0..8, or None for step(None)."""
        a = None if actions is None else np.ascontiguousarray(actions, dtype=np.int32).reshape(self.n, self.A)
        u = None if uniforms is None else np.ascontiguousarray(uniforms, dtype=np.float64)
        n_inj = 0 if u is None else u.shape[-1]
        lib().orc_step_batch(C.byref(self.cfg), _p(self.envs), C.c_int32(self.n), _p(a), C.c_uint64(self.seed),
                             C.c_uint32(self.env_id0), C.c_uint64(step_ctr), _p(u), C.c_int32(n_inj), _p(self.outs),
                             C.c_int32(self.threads))
        return self.outs

    def rollout(self, T: int, step_ctr0: int = 0, epoch_end_last: bool = True):
        """T steps of the per-env loop (random actions, timeout / done resets; new obstructions after the last step
        when epoch_end_last)."""
        chk = C.c_double(0.0)
        n = lib().orc_rollout(C.byref(self.cfg), _p(self.envs), C.c_int32(self.n), C.c_int32(T), C.c_uint64(self.seed),
                              C.c_uint32(self.env_id0), C.c_uint64(step_ctr0), C.c_int32(int(epoch_end_last)),
                              C.c_int32(self.threads), C.byref(chk))
        return int(n), float(chk.value)

    def shortest_path(self, i: int, det) -> float:
        return float(lib().orc_shortest_path(C.c_void_p(self.envs[i : i + 1].ctypes.data), i32(*det)))

    def sensors(self, i: int, agent: int = 0) -> np.ndarray:
        out = np.zeros(8)
        lib().orc_sensors(C.byref(self.cfg), C.c_void_p(self.envs[i : i + 1].ctypes.data), C.c_int32(agent), _p(out))
        return out


class Stat(C.Structure):
    _fields_ = [("mean", C.c_double), ("m2", C.c_double), ("std", C.c_double), ("count", C.c_int32)]


class Standardizer:
    """[n] independent running standardisers (RADTEAM_core.py:188-277 mode 1, test_environment StatBuff + clip mode 2)
    driven the way train.py drives them: update(x) then standardize(x); reset(mask) at episode ends."""

    def __init__(self, n: int, mode: int = 1):
        self.n, self.mode = int(n), int(mode)
        self.s = (Stat * self.n)()
        self.reset()

    def reset(self, mask=None):
        for i in range(self.n):
            if mask is None or mask[i]:
                lib().orc_stat_reset(C.byref(self.s[i]))

    def update_standardize(self, x, mask=None) -> np.ndarray:
        f = lib().orc_stat_update_standardize
        f.restype = C.c_double
        x = np.asarray(x, np.float64).reshape(self.n)
        z = np.zeros(self.n)
        for i in range(self.n):
            if mask is None or mask[i]:
                z[i] = f(C.byref(self.s[i]), C.c_double(float(x[i])), C.c_int32(self.mode))
        return z

    @property
    def mean(self):
        return np.array([s.mean for s in self.s])

    @property
    def m2(self):
        return np.array([s.m2 for s in self.s])

    @property
    def std(self):
        return np.array([s.std for s in self.s])


class MapsOracle:
    """One MapsBuffer of the reference (RADTEAM_core.py:394-932) restated in C (oracle/maps_oracle.c)."""

    MAP_NAMES = ("prediction", "location", "others", "readings", "visits", "obstacles", "combined")

    def __init__(self, n_agents: int, steps_per_episode: int = 120, dims=(27, 27), resolution_accuracy: float = 22.0):
        L = lib()
        L.orc_maps_new.restype = C.c_void_p
        L.orc_maps_data.restype = C.POINTER(C.c_float)
        L.orc_maps_status.restype = C.c_uint32
        self.dims, self.A = tuple(dims), int(n_agents)
        self._h = C.c_void_p(L.orc_maps_new(C.c_int32(dims[0]), C.c_int32(dims[1]), C.c_int32(n_agents),
                                            C.c_int32(steps_per_episode), C.c_double(resolution_accuracy)))

    def __del__(self):
        try:
            lib().orc_maps_free(self._h)
        except Exception:
            pass

    def reset(self):
        lib().orc_maps_reset(self._h)

    def observation_to_map(self, obs, agent_id: int, pred) -> np.ndarray:
        """obs [A, 11] float64, pred (x, y) deflated -> the seven maps [7, X, Y] float32 (a copy)."""
        o = np.ascontiguousarray(obs, dtype=np.float64).reshape(self.A, OBS_DIM)
        p = (C.c_double * 2)(float(pred[0]), float(pred[1]))
        lib().orc_maps_observation_to_map(self._h, _p(o), C.c_int32(agent_id), p)
        return self.maps()

    def maps(self) -> np.ndarray:
        ptr = lib().orc_maps_data(self._h)
        return np.ctypeslib.as_array(ptr, shape=(7, *self.dims)).copy()

    @property
    def status(self) -> int:
        return int(lib().orc_maps_status(self._h))


def gae(rew, val, path_end, boot, gamma=0.99, lam=0.90, threads=0):
    """Batched P:391-423 over [T, N] float32 arrays -> (adv, ret) float32."""
    rew = np.ascontiguousarray(rew, dtype=np.float32)
    val = np.ascontiguousarray(val, dtype=np.float32)
    pe = np.ascontiguousarray(path_end, dtype=np.uint8)
    boot = np.ascontiguousarray(boot, dtype=np.float32)
    T, N = rew.shape
    adv = np.empty_like(rew)
    ret = np.empty_like(rew)
    lib().orc_gae(_p(rew), _p(val), _p(pe), _p(boot), _p(adv), _p(ret), C.c_int32(T), C.c_int32(N), C.c_double(gamma),
                  C.c_double(lam), C.c_int32(threads))
    return adv, ret
