"""TEST INFRASTRUCTURE ONLY -- numpy front-end for the host-emulation build of the kernel logic (tests/emu/rs_emu.cpp).

Mirrors the layout radiation_ppo_b200.envs allocates on the device, but in host memory, and calls the emulated entry
points.  Never imported by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from radiation_ppo_b200 import _lib as L

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
EMU_LIB = os.path.join(HERE, "librs_emu.so")


def build_emu(force=False):
    srcs = [os.path.join(HERE, "rs_emu.cpp"), os.path.join(HERE, "cuda_host_shim.h"),
            os.path.join(ROOT, "radiation_ppo_b200", "csrc", "rs_env_impl.cuh"),
            os.path.join(ROOT, "radiation_ppo_b200", "csrc", "rs_step_tiled.cuh"),
            os.path.join(ROOT, "radiation_ppo_b200", "csrc", "rs_step1.cuh"),
            os.path.join(ROOT, "radiation_ppo_b200", "csrc", "rs_poisson_alias.h"),
            os.path.join(ROOT, "radiation_ppo_b200", "csrc", "rs_device.cuh"),
            os.path.join(ROOT, "radiation_ppo_b200", "csrc", "rs_rcp.cuh"),
            os.path.join(ROOT, "include", "radsearch_b200.h")]
    if force or not os.path.exists(EMU_LIB) or os.path.getmtime(EMU_LIB) < max(os.path.getmtime(s) for s in srcs):
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-I" + HERE, "-o", EMU_LIB,
                        srcs[0]], check=True)
    return EMU_LIB


_emu = None


def emu():
    global _emu
    if _emu is None:
        _emu = L.declare(C.CDLL(build_emu()), prefix="emu_")
        _emu.emu_step1.restype = _emu.emu_step.restype
        _emu.emu_step1.argtypes = _emu.emu_step.argtypes
        _emu.emu_div_count_mismatches.restype = C.c_longlong
        _emu.emu_div_count_mismatches.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong]
        _emu.emu_div_const_mismatches.restype = C.c_longlong
        _emu.emu_div_const_mismatches.argtypes = [C.c_void_p, C.c_longlong, C.c_double]
        _emu.emu_round2_fast.restype = _emu.emu_round2.restype = C.c_double
        _emu.emu_round2_fast.argtypes = _emu.emu_round2.argtypes = [C.c_double]
    return _emu


def make_config(n_agents=1, obstruction_count=5, enforce=True, k_max=None, max_ep_len=120, count_law=0,
                bbox=(0, 0, 2700, 2700), obs_area=(200, 500), standardize=0):
    c = L.RsConfig()
    for i, v in enumerate(bbox):
        c.bbox[i] = v
    c.obs_area[0], c.obs_area[1] = obs_area
    c.enforce = int(enforce)
    c.n_agents = n_agents
    c.obstruction_count = obstruction_count
    c.count_law = count_law
    c.max_ep_len = max_ep_len
    c.k_max = k_max if k_max is not None else (5 if obstruction_count == -1 else max(obstruction_count, 0))
    c.standardize = standardize
    return c


def _vp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class EmuEnv:
    def __init__(self, n, cfg, seed=0, env_id0=0, tiled=False):
        # tiled=True: single-agent steps go through the multi-agent tile program's phases instead of rs_step1.cuh
        self.n, self.cfg, self.seed, self.env_id0, self.tiled = n, cfg, seed, env_id0, tiled
        A, K = cfg.n_agents, cfg.k_max
        self.A, self.K = A, K
        self.src = np.zeros((n, 2), np.int32)
        self.rad = np.zeros((n, 2), np.int32)
        self.rects = np.zeros((max(K, 1), n, 4), np.int32)
        self.meta = np.zeros(n, np.int32)
        self.det = np.zeros((A, n, 2), np.int32)
        self.best = np.zeros((A, n), np.float64)
        self.aflags = np.zeros((A, n), np.int32)
        self.dsrc = np.zeros((n, max(4 * K, 1)), np.float64)
        self.vis = np.zeros((max(4 * K, 1), n), np.uint32)
        self.status = np.zeros(n, np.uint32)
        self.reset_list = np.zeros(n, np.int32)
        self.reset_count = np.zeros(1, np.int32)
        self.epi = np.zeros(n, np.uint32)
        self.nx_src = np.zeros((n, 2), np.int32)
        self.nx_det = np.zeros((n, 2), np.int32)
        self.nx_rad = np.zeros((n, 2), np.int32)
        self.nx_best = np.zeros(n, np.float64)
        self.nx_dsrc = np.zeros((n, max(4 * K, 1)), np.float64)
        self.nx_obs = np.zeros((n, A, 11), np.float32)
        self.nx_seq = np.zeros(n, np.uint32)
        self.refill_list = np.zeros((2, n), np.int32)
        self.refill_count = np.zeros(2, np.int32)
        self.ctr_dev = np.zeros(1, np.uint64)
        self.st_mean = np.zeros((A, n), np.float64)
        self.st_m2 = np.zeros((A, n), np.float64)
        self.raw_count = np.zeros((n, A), np.float32)
        self.ticket = np.zeros(1, np.uint32)
        self.dsf = np.zeros((n, max(4 * K, 1)), np.float32)
        self.nx_dsf = np.zeros((n, max(4 * K, 1)), np.float32)
        self.st = L.RsState(*[_vp(getattr(self, f)) for f, _ in L.RsState._fields_])
        self.obs = np.zeros((n, A, 11), np.float32)
        self.final_obs = np.zeros((n, A, 11), np.float32)
        self.reward = np.zeros((n, A), np.float32)
        self.team_reward = np.zeros(n, np.float32)
        self.done = np.zeros((n, A), np.uint8)
        self.info = np.zeros((n, A), np.uint8)
        self.ended = np.zeros(n, np.uint8)

    def step(self, actions, step_ctr, uniforms=None, flags=0):
        a = None if actions is None else np.ascontiguousarray(actions, np.int32).reshape(self.n, self.A)
        u = None if uniforms is None else np.ascontiguousarray(uniforms, np.float64)
        fn = emu().emu_step1 if (self.A == 1 and not self.tiled) else emu().emu_step
        rc = fn(C.byref(self.cfg), C.byref(self.st), _vp(a), _vp(self.obs), _vp(self.reward),
                            _vp(self.team_reward), _vp(self.done), _vp(self.info), _vp(self.ended),
                            _vp(self.final_obs), self.n, self.env_id0, self.seed, step_ctr, _vp(u),
                            0 if u is None else u.shape[-1], flags)
        assert rc == 0

    def reset(self, step_ctr, mask=None, new_mask=None, uniforms=None, flags=0):
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        nm = None if new_mask is None else np.ascontiguousarray(new_mask, np.uint8)
        u = None if uniforms is None else np.ascontiguousarray(uniforms, np.float64)
        rc = emu().emu_reset(C.byref(self.cfg), C.byref(self.st), _vp(m), _vp(nm), _vp(self.obs), self.n, self.env_id0,
                             self.seed, step_ctr, _vp(u), 0 if u is None else u.shape[-1], flags)
        assert rc == 0

    def load_scenarios(self, src, det, intensity, bkg, rects, num_obs, step_ctr=0, uniforms=None):
        arrs = [np.ascontiguousarray(x, np.int32) for x in (src, det, intensity, bkg, rects, num_obs)]
        u = None if uniforms is None else np.ascontiguousarray(uniforms, np.float64)
        k_in = arrs[4].shape[1] if arrs[4].ndim == 3 else 0
        rc = emu().emu_load_scenarios(C.byref(self.cfg), C.byref(self.st), _vp(arrs[0]), _vp(arrs[1]), _vp(arrs[2]),
                                      _vp(arrs[3]), _vp(arrs[4]), k_in, _vp(arrs[5]), _vp(self.obs), self.n,
                                      self.env_id0, self.seed, step_ctr, _vp(u), 0 if u is None else u.shape[-1])
        assert rc == 0

    def prepare(self, flags=0):
        rc = emu().emu_prepare(C.byref(self.cfg), C.byref(self.st), self.n, self.env_id0, self.seed, flags)
        assert rc == 0

    def query_sp(self, pts, variant=0):
        p = np.ascontiguousarray(pts, np.int32)
        out = np.zeros(self.n, np.float64)
        rc = emu().emu_query_shortest_path(C.byref(self.cfg), C.byref(self.st), _vp(p), _vp(out), self.n, variant)
        assert rc == 0
        return out

    # views in the oracle's terms
    @property
    def num_obs(self):
        return self.meta & 0xFF

    @property
    def env_done(self):
        return (self.meta >> 8) & 1

    @property
    def ep_len(self):
        return self.meta >> 16
