"""Times the UNMODIFIED reference (Python, /root/reference through oracle/shims) on this container's cores and writes
profiles/r02_reference_python_baselines.json -- the secondary CPU baselines SURVEY.md 8(d) / BASELINE.md section 2 name:

  * RadSearch.step (rad_search_env.py:443-728) with uniform random actions, reset on done / every 120 steps, new obstructions
    every 480 steps: 1 process and P = os.cpu_count() processes (one env per process, the reference's one-env-per-rank
    model), with 5 obstructions and with none;
  * PPOBuffer.GAE_advantage_and_rewardsToGO (ppo.py:391-423, scipy lfilter) over the columns of a [480, N] rollout, 1 and P
    processes;
  * one 480-step epoch of the single-agent loop (BASELINE configs[0]: one env, no obstructions, a GRU(11 -> 24) actor-critic
    stepping on the CPU, the reference PPOBuffer storing every step, GAE at every path end; loop semantics of
    algos/test_environment/ppo.py:495-573).

/root/reference does not exist on the GPU box and the geometry library the reference binds (PyVisiLibity, C++) cannot be
built here: the obstruction case runs through the exact-rational pure-Python restatement in oracle/shims/visilibity.py, far
slower than the SWIG library, so its numbers are a LOWER bound on the reference's speed and are marked as such.  bench.py
copies this file's content into its JSON line as `cpu_baseline_reference_python` (kind "reference-python, build
container"); the C port timed live on the GPU box's cores stays the headline reference arm.

  python tools/time_reference_python.py [--seconds 20]
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _step_worker(args):
    k_obs, seconds, seed = args
    import numpy as np

    from oracle.ref_env import load_reference_env

    m = load_reference_env()
    env = m.RadSearch(obstruction_count=k_obs, np_random=np.random.default_rng(seed), enforce_grid_boundaries=True)
    env.reset()
    rng = np.random.default_rng(seed + 1)
    n, in_ep, in_epoch = 0, 0, 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        _, _, done, _ = env.step({0: int(rng.integers(0, 8))})
        n += 1
        in_ep += 1
        in_epoch += 1
        if in_epoch == 480:
            env.epoch_end = True
            in_epoch = 0
        if done[0] or in_ep == 120 or env.epoch_end:
            env.reset()
            in_ep = 0
    return n, time.perf_counter() - t0


def _gae_worker(args):
    n_cols, seed = args
    import numpy as np

    from oracle.ref_env import load_reference_ppo

    ppo = load_reference_ppo()
    T = 480
    rng = np.random.default_rng(seed)
    rew = (-0.5 * rng.uniform(0, 1.5, (T, n_cols))).astype(np.float32)
    val = rng.normal(size=(T, n_cols)).astype(np.float32)
    t0 = time.perf_counter()
    for c in range(n_cols):
        buf = ppo.PPOBuffer(observation_dimension=11, max_size=T, max_episode_length=120, number_agents=1)
        buf.rew_buf[:] = rew[:, c]
        buf.val_buf[:] = val[:, c]
        for e in range(120, T + 1, 120):                       # a path end every 120 steps (timeouts)
            buf.ptr = e
            buf.GAE_advantage_and_rewardsToGO(float(val[e - 1, c]))
    return T * n_cols, time.perf_counter() - t0


def _config1_epoch(seed=2):
    """BASELINE configs[0]: one env, no obstructions, GRU actor-critic, one 480-step epoch on the CPU."""
    import numpy as np
    import torch

    from oracle.ref_env import load_reference_env, load_reference_ppo

    m, ppo = load_reference_env(), load_reference_ppo()
    torch.manual_seed(0)
    torch.set_num_threads(1)
    env = m.RadSearch(obstruction_count=0, np_random=np.random.default_rng(seed), enforce_grid_boundaries=True)
    gru, pi, vf = torch.nn.GRUCell(11, 24), torch.nn.Linear(24, 8), torch.nn.Linear(24, 1)
    buf = ppo.PPOBuffer(observation_dimension=11, max_size=480, max_episode_length=120, number_agents=1)
    obs = env.reset()[0][0]
    h = torch.zeros(1, 24)
    in_ep = 0
    t0 = time.perf_counter()
    with torch.no_grad():
        for t in range(480):
            x = torch.as_tensor(np.asarray(obs, np.float32))[None]
            h = gru(x, h)
            dist = torch.distributions.Categorical(logits=pi(h))
            a = dist.sample()
            v, logp = float(vf(h)), float(dist.log_prob(a))
            nobs, rew, done, _ = env.step({0: int(a)})
            buf.store(obs=obs, act=int(a), rew=rew["individual_reward"][0], val=v, logp=logp,
                      src=np.array(env.src_coords, dtype="float32"), full_observation={0: obs}, heatmap_stacks=None,
                      terminal=False)
            obs = nobs[0]
            in_ep += 1
            timeout, last = in_ep == 120, t == 479
            if done[0] or timeout or last:
                boot = 0.0
                if timeout or last:
                    boot = float(vf(gru(torch.as_tensor(np.asarray(obs, np.float32))[None], h)))
                buf.GAE_advantage_and_rewardsToGO(boot)
                if done[0] or timeout:
                    buf.store_episode_length(in_ep)
                if last:
                    env.epoch_end = True
                obs = env.reset()[0][0]
                h = torch.zeros(1, 24)
                in_ep = 0
    return time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=20.0)
    ap.add_argument("--only-config1", action="store_true", help="re-time the config-1 epoch and merge it into the existing file")
    args = ap.parse_args()
    path = os.path.join(ROOT, "profiles", "r02_reference_python_baselines.json")
    if args.only_config1:
        out = json.load(open(path))
        out["config1_epoch_s"] = _config1_epoch()
        out["config1_env_steps_per_s"] = 480 / out["config1_epoch_s"]
        json.dump(out, open(path, "w"), indent=1)
        print("config1 epoch", out["config1_epoch_s"])
        return
    P = os.cpu_count() or 1
    out = {"where": "build container (the reference cannot travel to the GPU box)", "cores": P,
           "python": sys.version.split()[0],
           "note": "geometry through the exact-rational pure-Python visilibity restatement (oracle/shims): a lower bound on "
                   "the reference's speed with its C++ library"}
    with mp.get_context("spawn").Pool(P) as pool:
        for k_obs in (0, 5):
            n1, dt1 = _step_worker((k_obs, args.seconds, 2))
            res = pool.map(_step_worker, [(k_obs, args.seconds, 10 + i) for i in range(P)])
            out[f"step_k{k_obs}"] = {"unit": "env-steps/s", "one_process": n1 / dt1, "P_processes": sum(n / dt for n, dt in res),
                                     "processes": P, "seconds": args.seconds}
            print(k_obs, out[f"step_k{k_obs}"], flush=True)
        n1, dt1 = _gae_worker((256, 0))
        res = pool.map(_gae_worker, [(256, i) for i in range(P)])
        out["gae_lfilter"] = {"unit": "elements/s", "one_process": n1 / dt1, "P_processes": sum(n / dt for n, dt in res),
                              "GBps_at_17B_one_process": 17 * n1 / dt1 / 1e9,
                              "GBps_at_17B_P_processes": 17 * sum(n / dt for n, dt in res) / 1e9, "processes": P,
                              "sample": "[480, 256] columns per process, reference PPOBuffer + scipy lfilter, path end every 120 steps"}
        print(out["gae_lfilter"], flush=True)
    out["config1_epoch_s"] = _config1_epoch()
    out["config1_env_steps_per_s"] = 480 / out["config1_epoch_s"]
    print("config1 epoch", out["config1_epoch_s"], flush=True)
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
