"""Shared helpers of the parity tests: golden-record loading, oracle drivers and array comparisons."""
from __future__ import annotations

import os

import numpy as np

from oracle import c_oracle as co

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

STEP_FILES = {
    "ref_steps_k5_enforce": dict(n_agents=1, obstruction_count=5, enforce=1),
    "ref_steps_krand_free": dict(n_agents=1, obstruction_count=-1, enforce=0),
    "ref_steps_a3_k3": dict(n_agents=3, obstruction_count=3, enforce=1),
    "ref_steps_k0": dict(n_agents=1, obstruction_count=0, enforce=1),
    "ref_probes_k5_enforce": dict(n_agents=1, obstruction_count=5, enforce=1),
    "ref_probes_krand_free": dict(n_agents=1, obstruction_count=-1, enforce=0),
    "ref_probes_a3_k4": dict(n_agents=3, obstruction_count=4, enforce=1),
    "ref_probes_k7": dict(n_agents=1, obstruction_count=7, enforce=1),
}


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def oracle_load_records(g, cfg_kw):
    """OracleBatch with one env per golden record, every field of the reference's pre-call state restored."""
    n = len(g["is_reset"])
    A = cfg_kw["n_agents"]
    ob = co.OracleBatch(n, co.default_config(**cfg_kw))
    e = ob.envs
    e["num_obs"] = g["pre_num_obs"]
    e["rect"] = g["pre_rects"]
    e["src"] = g["pre_src"]
    e["intensity"] = g["pre_intensity"]
    e["bkg"] = g["pre_bkg"]
    e["det"][:, :A] = g["pre_det"]
    e["best"][:, :A] = g["pre_best"]
    e["sp"][:, :A] = g["pre_sp"]
    e["euc"][:, :A] = g["pre_euc"]
    e["oob_count"][:, :A] = g["pre_oob_count"]
    e["blocked"][:, :A] = g["pre_blocked"]
    e["done"] = g["pre_done"]
    e["iter_count"] = g["pre_iter_count"]
    return ob


def step_records_with_oracle(g, cfg_kw):
    """Run every golden record through the C oracle (reset records as step(None), others with their actions)."""
    ob = oracle_load_records(g, cfg_kw)
    n, A = ob.n, cfg_kw["n_agents"]
    is_reset = g["is_reset"].astype(bool)
    outs = np.zeros(n, co.OUT_DTYPE)
    envs = ob.envs.copy()
    for sel, acts in ((is_reset, None), (~is_reset, g["out_actions"])):
        idx = np.where(sel)[0]
        if len(idx) == 0:
            continue
        sub = co.OracleBatch(len(idx), ob.cfg)
        sub.envs[:] = ob.envs[idx]
        sub.step(None if acts is None else acts[idx], 0, uniforms=g["out_uniforms"][idx])
        outs[idx] = sub.outs
        envs[idx] = sub.envs
    return envs, outs


def assert_matches_golden(g, envs, outs, A, obs_exact=True, sens_rtol=0.0, label=""):
    """Compare oracle-style (envs, outs) structured arrays with the golden outputs of the reference."""
    is_reset = g["is_reset"].astype(bool)
    ne = np.testing.assert_array_equal
    ne(envs["det"][:, :A], g["out_det"], err_msg=label + " det")
    ne(envs["oob_count"][:, :A], g["out_oob_count"], err_msg=label + " oob_count")
    ne(envs["blocked"][:, :A], g["out_blocked"], err_msg=label + " blocked")
    ne(envs["best"][:, :A], g["out_best"], err_msg=label + " best")
    obs = outs["obs"][:, :A]
    ne(obs[:, :, 0], g["out_obs"][:, :, 0], err_msg=label + " counts")
    if obs_exact:
        ne(obs, g["out_obs"], err_msg=label + " obs")
    else:
        ref = g["out_obs"].astype(np.float32)
        ne(obs[:, :, 1:3].astype(np.float32), ref[:, :, 1:3], err_msg=label + " xy")
        np.testing.assert_allclose(obs[:, :, 3:], ref[:, :, 3:], rtol=sens_rtol, atol=0, err_msg=label + " sensors")
        ne(obs[:, :, 3:] == 0, ref[:, :, 3:] == 0, err_msg=label + " sensor zeros")
        ne(obs[:, :, 3:] == 1, ref[:, :, 3:] == 1, err_msg=label + " sensor ones")
    s = ~is_reset
    ne(outs["reward"][s][:, :A], g["out_reward"][s], err_msg=label + " reward")
    ne(outs["done"][s][:, :A], g["out_done"][s], err_msg=label + " done")
    ne(outs["team_reward"][s], g["out_team_reward"][s], err_msg=label + " team reward")


def gae_numpy_reference(rew, val, end, boot, gamma=0.99, lam=0.90):
    """Per-column, per-trajectory restatement of P:391-423 with scipy.signal.lfilter (discount_cumsum P:62-85)."""
    import scipy.signal

    def dcs(x, d):
        return scipy.signal.lfilter([1], [1, float(-d)], x[::-1], axis=0)[::-1]

    T, N = rew.shape
    adv = np.zeros((T, N), np.float32)
    ret = np.zeros((T, N), np.float32)
    for n in range(N):
        s = 0
        for t in range(T):
            if end[t, n] or t == T - 1:
                sl = slice(s, t + 1)
                rews = np.append(rew[sl, n], float(boot[t, n]))      # `last_state_value: float` -> float64 math
                vals = np.append(val[sl, n], float(boot[t, n]))
                deltas = rews[:-1] + gamma * vals[1:] - vals[:-1]
                adv[sl, n] = dcs(deltas, gamma * lam)
                ret[sl, n] = dcs(rews, gamma)[:-1]
                s = t + 1
    return adv, ret


def synthetic_rollout(T, N, seed=0, max_ep=120):
    """SURVEY 8(d) GAE inputs: rew ~ -0.5*U(0,1.5) with 5% +0.1, val ~ N(0,1), path ends every <= max_ep steps."""
    rng = np.random.default_rng(seed)
    rew = (-0.5 * rng.uniform(0, 1.5, (T, N))).astype(np.float32)
    rew[rng.random((T, N)) < 0.05] = 0.1
    val = rng.normal(size=(T, N)).astype(np.float32)
    end = np.zeros((T, N), np.uint8)
    boot = np.zeros((T, N), np.float32)
    for n in range(N):
        t = 0
        while t < T:
            e = min(T, t + int(rng.integers(1, max_ep + 1))) - 1
            end[e, n] = 1
            boot[e, n] = 0.0 if rng.random() < 0.3 else np.float32(rng.normal())
            t = e + 1
    end[T - 1] = 1
    return rew, val, end, boot


def compare_state(em, ob, A):
    e = ob.envs
    ne = np.testing.assert_array_equal
    ne(em.num_obs, e["num_obs"])
    for k in range(em.K):
        m = e["num_obs"] > k
        ne(em.rects[k][m], e["rect"][:, k][m])
    ne(em.src, e["src"])
    ne(em.rad[:, 0], e["intensity"])
    ne(em.rad[:, 1], e["bkg"])
    for a in range(A):
        ne(em.det[a], e["det"][:, a])
        ne(em.best[a], e["best"][:, a])
        ne(em.aflags[a] & 0xFFFFFF, e["oob_count"][:, a])
        ne((em.aflags[a] >> 24) & 1, e["blocked"][:, a])
    ne(em.env_done, e["done"])
    ne(em.status, e["status"])


def compare_obs(obs, ref64, sel=None):
    ref = ref64.astype(np.float32)
    if sel is not None:
        obs, ref = obs[sel], ref[sel]
    np.testing.assert_array_equal(obs[:, :, :3], ref[:, :, :3])
    np.testing.assert_allclose(obs[:, :, 3:], ref[:, :, 3:], rtol=1e-5, atol=0)
    np.testing.assert_array_equal(obs[:, :, 3:] == 0, ref[:, :, 3:] == 0)
    np.testing.assert_array_equal(obs[:, :, 3:] == 1, ref[:, :, 3:] == 1)


class GpuView:
    """numpy snapshot of a radiation_ppo_b200.RadSearch instance, with the attribute names of tests/emu/harness.EmuEnv."""

    def __init__(self, env):
        c = lambda x: x.detach().cpu().numpy()      # noqa: E731
        self.K = env._cfg.k_max
        self.src, self.rad, self.rects, self.meta = c(env._src), c(env._rad), c(env._rects), c(env._meta)
        self.det, self.best, self.aflags, self.status = c(env._det), c(env._best), c(env._aflags), c(env._status).astype(np.uint32)
        self.obs, self.final_obs, self.reward, self.team_reward = c(env.obs), c(env.final_obs), c(env.reward), c(env.team_reward)
        self.done, self.info, self.ended = c(env.done_flags), c(env.info_flags), c(env.ended)
        self.reset_list, self.reset_count = c(env._reset_list), c(env._reset_count)
        self.num_obs = self.meta & 0xFF
        self.env_done = (self.meta >> 8) & 1
        self.ep_len = self.meta >> 16


class StandardizedOracle:
    """The caller-side standardisation of train.py (RAD-A2C branch) around an OracleBatch: one running standardiser per
    (env, agent), `update(reading)` then `standardize(reading)` for the reset observation (T:311, 548) and for every
    next observation (T:436, 339, 469), `reset()` with the episode (T:509).  z holds the standardised count channel
    of the current observation, final_z the one of the last observation of an episode that just ended."""

    def __init__(self, ob, mode):
        self.ob, self.A, self.mode = ob, ob.A, mode
        self.st = co.Standardizer(ob.n * ob.A, mode)
        self.z = np.zeros((ob.n, ob.A))

    def _mask(self, env_mask):
        return None if env_mask is None else np.repeat(np.asarray(env_mask, bool), self.A)

    def after_reset(self, env_mask=None):
        m = self._mask(env_mask)
        self.st.reset(m)
        z = self.st.update_standardize(self.ob.outs["obs"][:, :self.A, 0].reshape(-1), m).reshape(self.ob.n, self.A)
        self.z = z if env_mask is None else np.where(np.asarray(env_mask, bool)[:, None], z, self.z)
        return self.z

    def after_step(self):
        self.z = self.st.update_standardize(self.ob.outs["obs"][:, :self.A, 0].reshape(-1)).reshape(self.ob.n, self.A)
        return self.z
