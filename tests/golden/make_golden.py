"""Generates the committed fixtures under tests/golden/ by running the UNMODIFIED reference from /root/reference
(through oracle/shims) in the build container.  Re-run: `python tests/golden/make_golden.py` (takes a few minutes).

  ref_steps_*.npz     per-call records of RadSearch.reset()/step(): full pre-state, actions, the uniforms numpy's
                      poisson consumed, and every output (observation, rewards, done, info, new state).
  ref_probes_*.npz    the same for adversarial one-step probes (detector on/near obstruction edges, corners, walls,
                      the source).
  ref_gae.npz         PPOBuffer.GAE_advantage_and_rewardsToGO outputs (scipy lfilter path) on random trajectories and
                      the known-answer vectors of unit_tests/test_PPO.py:259-286, 462-571.
  ref_reset_stats.npz marginals of 1500 reference resets (distributional parity of the Philox sampler).
  scenarios_v4.npz    the reference's saved evaluation scenarios (inputs only), first 250 of each obstruction count.
  action_lut.npz      get_step(a) for a in 0..8.
  ref_standardize.npz StatisticStandardization (RADTEAM_core.py:188-277) and StatBuff (test_environment/core.py:55-79)
                      run over count sequences in the caller's order update(x); standardize(x): z-scores and the
                      running mean / M2 / std after every reading.
  ref_maps_*.npz      MapsBuffer.observation_to_map (RADTEAM_core.py:394-932): one reference MapsBuffer per agent fed
                      with the observation dicts of an oracle-env rollout (float64) and random source predictions;
                      the seven maps of every agent's buffer after every call, and the reset points.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.ref_env import REFERENCE_ROOT, load_reference_env, load_reference_ppo  # noqa: E402
from tests.golden.ref_harness import record_episodes, record_probes  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
PRE_KEYS = ["src", "det", "intensity", "bkg", "rects", "num_obs", "best", "iter_count", "done", "oob_count", "blocked",
            "sp", "euc"]
OUT_KEYS = ["actions", "uniforms", "lam", "obs", "reward", "team_reward", "done", "oob", "oob_count", "blocked", "det",
            "sp", "best", "los"]


def pack(recs):
    d = {"is_reset": np.array([r["kind"] == "reset" for r in recs], np.uint8)}
    for k in PRE_KEYS:
        vals = []
        for r in recs:
            p = r["pre"]
            if k in p:
                vals.append(np.asarray(p[k]))
            else:  # reset records have no agent history: zeros / the post-reset value
                proto = {"iter_count": 0, "done": 0, "oob_count": np.zeros_like(r["oob_count"]),
                         "blocked": np.zeros_like(r["blocked"]), "sp": r["best"] * 0 + p["best"],
                         "euc": r["best"] * 0}[k]
                vals.append(np.asarray(proto))
        d["pre_" + k] = np.stack(vals)
    for k in OUT_KEYS:
        d["out_" + k] = np.stack([np.asarray(r[k]) for r in recs])
    return d


def main():
    only = set(sys.argv[1:])          # e.g. `python tests/golden/make_golden.py gae` regenerates one group
    want = lambda grp: not only or grp in only      # noqa: E731
    os.makedirs(OUT, exist_ok=True)
    m = load_reference_env()
    np.savez_compressed(os.path.join(OUT, "action_lut.npz"), step=np.array([m.get_step(a) for a in range(9)], np.float64))

    jobs = [
        ("ref_steps_k5_enforce", lambda: record_episodes(2, 420, obstruction_count=5, enforce=True)),
        ("ref_steps_krand_free", lambda: record_episodes(3, 360, obstruction_count=-1, enforce=False)),
        ("ref_steps_a3_k3", lambda: record_episodes(5, 260, obstruction_count=3, enforce=True, n_agents=3, idle_prob=0.2)),
        ("ref_steps_k0", lambda: record_episodes(7, 200, obstruction_count=0, enforce=True)),
        ("ref_probes_k5_enforce", lambda: record_probes(11, 12, 60, obstruction_count=5, enforce=True)),
        ("ref_probes_krand_free", lambda: record_probes(12, 8, 50, obstruction_count=-1, enforce=False)),
        ("ref_probes_a3_k4", lambda: record_probes(13, 8, 40, obstruction_count=4, enforce=True, n_agents=3)),
        ("ref_probes_k7", lambda: record_probes(14, 5, 50, obstruction_count=7, enforce=True)),
    ]
    for name, fn in jobs:
        if not (want("steps") or (want("probes") and "probes" in name)):
            continue
        recs = fn()
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **pack(recs))
        print(name, len(recs), "records", flush=True)

    if want("gae"):
        make_gae()
    if want("reset"):
        make_reset_stats(m)
    if want("scenarios"):
        make_scenarios()
    if want("standardize"):
        make_standardize()
    if want("maps"):
        make_maps()
    if want("get"):
        make_get()


def make_gae():
    # ---- GAE through the reference PPOBuffer -------------------------------------------------------------------
    ppo = load_reference_ppo()
    rng = np.random.default_rng(0)
    T = 96
    cols = []
    for c in range(24):
        buf = ppo.PPOBuffer(observation_dimension=11, max_size=T, max_episode_length=120, number_agents=1)
        rew = (-0.5 * rng.uniform(0, 1.5, T)).astype(np.float32)
        rew[rng.random(T) < 0.05] = 0.1
        val = rng.normal(size=T).astype(np.float32)
        end = np.zeros(T, np.uint8)
        boot = np.zeros(T, np.float32)
        t = 0
        while t < T:
            ln = int(rng.integers(1, 40))
            e = min(T, t + ln) - 1
            end[e] = 1
            boot[e] = 0.0 if rng.random() < 0.4 else np.float32(rng.normal())
            t = e + 1
        for t in range(T):
            buf.store(obs=np.zeros(11, np.float32), act=0, rew=rew[t], val=val[t], logp=0.0, src=np.zeros(2),
                      full_observation={}, heatmap_stacks=None, terminal=False)
            if end[t]:
                # `last_state_value: float` (P:391): a Python float -> the whole computation runs in float64.
                # (RADA2C_core.py:549 hands in a float32 ndarray instead, which keeps the deltas in float32; that
                # variant is recorded as adv32/ret32 and only has to agree to the stated 1e-5 tolerance.)
                buf.GAE_advantage_and_rewardsToGO(float(boot[t]))
        buf32 = ppo.PPOBuffer(observation_dimension=11, max_size=T, max_episode_length=120, number_agents=1)
        for t in range(T):
            buf32.store(obs=np.zeros(11, np.float32), act=0, rew=rew[t], val=val[t], logp=0.0, src=np.zeros(2),
                        full_observation={}, heatmap_stacks=None, terminal=False)
            if end[t]:
                buf32.GAE_advantage_and_rewardsToGO(np.array([boot[t]], np.float32))
        cols.append((rew, val, end, boot, buf.adv_buf.copy(), buf.ret_buf.copy(), buf32.adv_buf.copy(),
                     buf32.ret_buf.copy()))
    g = {k: np.stack([c[i] for c in cols], axis=1)
         for i, k in enumerate(["rew", "val", "end", "boot", "adv", "ret", "adv32", "ret32"])}
    # known-answer vectors of the reference's own unit tests
    g["kat_rewards"] = np.array([-0.46, -0.48, -0.46, -0.45, -0.45, -0.47, -0.48, -0.48, -0.48, -0.49])
    g["kat_values"] = np.array([-0.26629043, -0.26634163, -0.26718464, -0.26631153, -0.26637784, -0.26601458,
                                -0.26657045, -0.2666973, -0.26680088, -0.26717135])
    g["kat_last_val"] = np.float64(-0.26717135)
    rews = np.append(g["kat_rewards"], g["kat_last_val"])
    vals = np.append(g["kat_values"], g["kat_last_val"])
    deltas = rews[:-1] + 0.99 * vals[1:] - vals[:-1]
    g["kat_adv"] = ppo.discount_cumsum(deltas, 0.99 * 0.90)
    g["kat_ret"] = ppo.discount_cumsum(rews, 0.99)[:-1]
    np.savez_compressed(os.path.join(OUT, "ref_gae.npz"), **g)
    print("ref_gae done", flush=True)



def make_reset_stats(m):
    # ---- reset marginals ----------------------------------------------------------------------------------------
    from tests.golden.ref_harness import rect_of

    env = m.RadSearch(obstruction_count=-1, np_random=np.random.default_rng(99), enforce_grid_boundaries=True)
    rows = []
    for i in range(1500):
        env.epoch_end = (i % 3 == 0)
        env.reset()
        ag = env.agents[0]
        area = sum((r[2] - r[0]) * (r[3] - r[1]) for r in map(rect_of, env.poly[: env.num_obs]))
        rows.append((env.num_obs, env.src_coords[0], env.src_coords[1], ag.det_coords[0], ag.det_coords[1],
                     env.intensity, env.bkg_intensity, ag.prev_det_dist, m.dist_p(ag.det_coords, env.src_coords),
                     int(ag.intersect), area))
    np.savez_compressed(os.path.join(OUT, "ref_reset_stats.npz"), rows=np.array(rows, np.float64),
                        cols=np.array(["num_obs", "src_x", "src_y", "det_x", "det_y", "intensity", "bkg", "sp", "euc",
                                       "los_blocked", "rect_area"]))
    print("reset stats done", flush=True)



def make_scenarios():
    # ---- saved evaluation scenarios ------------------------------------------------------------------------------
    from radiation_ppo_b200.scenario_io import load_test_env_dict, scenario_arrays

    base = os.path.join(REFERENCE_ROOT, "algos", "multiagent", "evaluation", "test_environments")
    sc = {}
    for k in range(0, 8):
        path = os.path.join(base, f"test_env_dict_obs{k}_med_v4")
        if not os.path.exists(path):
            continue
        arr = scenario_arrays(load_test_env_dict(path), range(250), k_max=7)
        for key, v in arr.items():
            sc[f"obs{k}_{key}"] = v
    np.savez_compressed(os.path.join(OUT, "scenarios_v4.npz"), **sc)
    print("scenarios done")


def make_get():
    """PPOBuffer.get() (P:425-502): one reference buffer per column; the rows of its `ep_form` tensors and their lengths."""
    ppo = load_reference_ppo()
    rng = np.random.default_rng(11)
    T, N = 72, 8
    g = dict(obs=rng.normal(size=(T, N, 11)).astype(np.float32), act=rng.integers(0, 8, (T, N)).astype(np.float32),
             rew=(-0.5 * rng.uniform(0, 1.5, (T, N))).astype(np.float32), val=rng.normal(size=(T, N)).astype(np.float32),
             logp=rng.normal(size=(T, N)).astype(np.float32), src=rng.uniform(0, 2200, (T, N, 2)).astype(np.float32),
             end=np.zeros((T, N), np.uint8), boot=np.zeros((T, N), np.float32))
    rows, lens, advs = [], [], []
    for n in range(N):
        buf = ppo.PPOBuffer(observation_dimension=11, max_size=T, max_episode_length=120, number_agents=1)
        t = 0
        while t < T:
            e = min(T, t + int(rng.integers(1, 30))) - 1
            g["end"][e, n] = 1
            g["boot"][e, n] = 0.0 if rng.random() < 0.4 else np.float32(rng.normal())
            t = e + 1
        start = 0
        for t in range(T):
            buf.store(obs=g["obs"][t, n], act=g["act"][t, n], rew=g["rew"][t, n], val=g["val"][t, n], logp=g["logp"][t, n],
                      src=g["src"][t, n], full_observation={}, heatmap_stacks=None, terminal=False)
            if g["end"][t, n]:
                buf.GAE_advantage_and_rewardsToGO(float(g["boot"][t, n]))
                # train.py:493-503 stores the length of episodes that are over; an epoch cut-off leaves the tail to get()
                if t < T - 1 or n % 2 == 0:
                    buf.store_episode_length(t + 1 - start)
                start = t + 1
        data = buf.get()
        ep = [e[0].numpy() for e in data["ep_form"]]
        rows.append(np.concatenate(ep, axis=0))
        ln = [len(e) for e in ep]
        lens.append(np.array(ln + [0] * (T - len(ln)), np.int32))
        advs.append(data["adv"].numpy())
    g["ep_rows"] = np.stack(rows)           # [N, T, 17]
    g["ep_lens"] = np.stack(lens)           # [N, T] zero padded
    g["adv_norm"] = np.stack(advs, axis=1)  # [T, N] per-column (single rank) normalisation
    np.savez_compressed(os.path.join(OUT, "ref_get_epform.npz"), **g)
    print("ref_get_epform done", flush=True)


def make_maps():
    import importlib
    import warnings

    from oracle import c_oracle as co
    from oracle.ref_env import _prepare_path

    _prepare_path()
    core = importlib.import_module("algos.multiagent.NeuralNetworkCores.RADTEAM_core")
    scale = 1 / 2200.0
    ra = core.calculate_resolution_accuracy(resolution_multiplier=0.01, scale=scale)
    offset = scale * 500.0                                   # enforce_boundaries: RADTEAM_core.py:1738-1739
    for name, A, T, idle, ml, seed in (("ref_maps_a1", 1, 130, 0.05, 40, 3), ("ref_maps_a4", 4, 90, 0.15, 30, 4),
                                      ("ref_maps_a2_idle", 2, 80, 0.7, 80, 5)):
        rng = np.random.default_rng(seed)
        ob = co.OracleBatch(1, co.default_config(n_agents=A, obstruction_count=4, enforce=1, max_ep_len=ml), seed=seed)
        ob.reset()
        bufs = [core.MapsBuffer(observation_dimension=11, steps_per_episode=120, number_of_agents=A, grid_bounds=(1, 1),
                                resolution_accuracy=ra, offset=offset, resolution_multiplier=0.01) for _ in range(A)]
        assert bufs[0].map_dimensions == (27, 27)
        obs_seq, pred_seq, reset_seq, maps_seq = [], [], [], []
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for t in range(T):
                obs = ob.outs["obs"][0, :A].copy()                       # float64 [A, 11], the env's observation dict
                pred = rng.uniform(0.0, 1.2, size=(A, 2))
                d = {i: obs[i].copy() for i in range(A)}
                stacks = [np.stack(bufs[i].observation_to_map(d, i, (float(pred[i, 0]), float(pred[i, 1])))) for i in range(A)]
                obs_seq.append(obs); pred_seq.append(pred); maps_seq.append(np.stack(stacks).astype(np.float32))
                acts = rng.integers(0, 8, size=(1, A))
                acts[rng.random((1, A)) < idle] = 8
                ob.step(acts, t + 1)
                e = ob.envs
                ended = bool(e["done"][0] == 1 or e["ep_len"][0] == ml)
                reset_seq.append(ended)
                if ended:
                    # the bootstrap call on the final observation (train.py:476-480), then reset_agent (train.py:536-538)
                    fobs = ob.outs["obs"][0, :A].copy()
                    fpred = rng.uniform(0.0, 1.2, size=(A, 2))
                    fd = {i: fobs[i].copy() for i in range(A)}
                    fst = [np.stack(bufs[i].observation_to_map(fd, i, (float(fpred[i, 0]), float(fpred[i, 1])))) for i in range(A)]
                    obs_seq.append(fobs); pred_seq.append(fpred); maps_seq.append(np.stack(fst).astype(np.float32))
                    reset_seq.append(2)                                   # 2 = this record is a bootstrap call; reset follows
                    for b in bufs:
                        b.reset()
                    ob.reset(mask=np.ones(1), new_obstacles=np.zeros(1))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), obs=np.stack(obs_seq), pred=np.stack(pred_seq),
                            reset_after=_reset_flags(reset_seq),
                            maps=np.stack(maps_seq), ra=np.float64(ra), dims=np.array([27, 27]))
        print(name, len(obs_seq), "calls", flush=True)


def _reset_flags(reset_seq):
    """One flag per recorded call: 0 = nothing follows, 1 = the buffers are reset after this call (episode ended and this
    was the bootstrap call)."""
    out = []
    for r in reset_seq:
        if r == 2:
            out.append(1)
        else:
            out.append(0)
    return np.array(out, np.int32)


def make_standardize():
    import importlib

    from oracle.ref_env import _prepare_path

    _prepare_path()
    core = importlib.import_module("algos.multiagent.NeuralNetworkCores.RADTEAM_core")
    old = importlib.import_module("algos.test_environment.core")
    rng = np.random.default_rng(21)
    S, Lmax = 96, 121
    x = np.zeros((S, Lmax))
    length = np.zeros(S, np.int32)
    out = {k: np.zeros((S, Lmax)) for k in ("z1", "mean1", "m2_1", "std1", "z2", "mean2", "m2_2", "std2")}
    for s in range(S):
        n = int(rng.integers(1, Lmax + 1))
        length[s] = n
        kind = s % 6
        if kind == 0:                       # far from the source: background only, small spread (std clamps to 1)
            seq = rng.poisson(float(rng.integers(10, 51)), n)
        elif kind == 1:                     # approaching the source: counts grow by orders of magnitude
            seq = rng.poisson(np.linspace(30.0, float(rng.integers(2000, 90000)), n))
        elif kind == 2:                     # constant readings: sample variance exactly 0 (the two std rules differ)
            seq = np.full(n, int(rng.integers(10, 5000)))
        elif kind == 3:                     # spikes: |z| beyond 8 (mode 2 clips)
            seq = rng.poisson(20.0, n)
            seq[rng.integers(0, n)] = 9_000_000
        elif kind == 4:                     # tiny spread: 0 < std < 1
            seq = 1000 + (rng.random(n) < 0.1).astype(np.int64)
        else:
            seq = rng.poisson(rng.uniform(10, 9e4, n))
        a, b = core.StatisticStandardization(), old.StatBuff()
        for t, v in enumerate(seq):
            v = float(v)
            x[s, t] = v
            a.update(v)
            out["z1"][s, t], out["mean1"][s, t], out["m2_1"][s, t], out["std1"][s, t] = a.standardize(v), a.mean, a.square_dist_mean, a.std
            b.update(v)
            out["z2"][s, t] = np.clip((v - b.mu) / b.sig_obs, -8, 8)        # test_environment/ppo.py:502
            out["mean2"][s, t], out["m2_2"][s, t], out["std2"][s, t] = b.mu, b.sig_sto, b.sig_obs
    np.savez_compressed(os.path.join(OUT, "ref_standardize.npz"), x=x, length=length, **out)
    print("ref_standardize", S, "sequences", flush=True)


if __name__ == "__main__":
    main()
