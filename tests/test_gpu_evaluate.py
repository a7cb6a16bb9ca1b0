"""GPU test (pytest -m gpu) of the batched Monte-Carlo evaluation driver (radiation_ppo_b200/evaluate.py)
against the oracle played the reference's way: one scenario, one run at a time, until done or the timeout
(evaluate.py:333-475)."""
import numpy as np
import pytest
import torch

from oracle import c_oracle as co
from tests import parity_util as pu

pytestmark = pytest.mark.gpu

import radiation_ppo_b200 as rp  # noqa: E402


@pytest.mark.parametrize("k,A", [(5, 1), (3, 2)])
def test_monte_carlo_evaluator_matches_oracle_episodes(k, A):
    sc = pu.load_golden("scenarios_v4")
    S, R, T = 48, 5, 120
    arr = {key: sc[f"obs{k}_{key}"][:S] for key in ("src", "det", "intensity", "bkg", "rects", "num_obs")}
    ev = rp.MonteCarloEvaluator(scenarios=arr, montecarlo_runs=R, steps_per_episode=T, obstruction_count=k, seed=123,
                                enforce_grid_boundaries=True, number_agents=A)
    N = S * R
    rng = np.random.default_rng(k)
    table = rng.integers(0, 8, size=(T, N, A)).astype(np.int32)
    # bias the walk towards the source so that a good share of the runs succeeds
    src = np.repeat(arr["src"], R, axis=0).astype(np.float64) / 2200.0
    dirs = np.array([[-1, 0], [-1, 1], [0, 1], [1, 1], [1, 0], [1, -1], [0, -1], [-1, -1]], np.float64)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    tab_dev = torch.as_tensor(table, device=ev.env.device)
    src_dev = torch.as_tensor(src, dtype=torch.float32, device=ev.env.device)
    dirs_dev = torch.as_tensor(dirs, dtype=torch.float32, device=ev.env.device)
    log = []

    def policy(obs, t):
        # greedy towards the source on even steps (float32 arithmetic on the device), table action on odd steps
        to = src_dev[:, None, :] - obs[:, :, 1:3]
        greedy = (to @ dirs_dev.T).argmax(dim=2).to(torch.int32)
        a = torch.where(torch.tensor(t % 2 == 0, device=obs.device), greedy, tab_dev[t])
        log.append(a.cpu().numpy())
        return a

    res = ev.run(policy)
    # the oracle, the reference's way
    ob = co.OracleBatch(N, co.default_config(n_agents=A, obstruction_count=k, enforce=1), seed=123)
    rep = {key: np.repeat(v, R, axis=0) for key, v in arr.items()}
    ob.load_scenarios(rep["src"], rep["det"], rep["intensity"], rep["bkg"], rep["rects"], rep["num_obs"])
    active = np.ones(N, bool); success = np.zeros(N, bool); length = np.zeros(N, np.int32); ret = np.zeros(N)
    ctr0 = ev.env._ctr - len(log)
    for t in range(len(log)):
        ob.step(log[t], ctr0 + t + 1)
        team = ob.outs["team_reward"].astype(np.float32).astype(np.float64)
        ret += np.where(active, team, 0.0)
        length += active
        term = (ob.outs["done"][:, :A] != 0).any(axis=1) & active
        success |= term
        active &= ~term
    np.testing.assert_array_equal(res.success.cpu().numpy().reshape(-1), success)
    np.testing.assert_array_equal(res.episode_length.cpu().numpy().reshape(-1), length)
    np.testing.assert_allclose(res.episode_return.cpu().numpy().reshape(-1), ret, rtol=0, atol=1e-9)
    np.testing.assert_array_equal(res.success_counter.cpu().numpy(), success.reshape(S, R).sum(1))
    assert 0.1 < success.mean() < 1.0 and res.completed_runs == R
    s = res.summary()
    assert s["success_rate"] == pytest.approx(success.mean())


def test_batched_reset_writes_reference_test_environment_files(tmp_path):
    """create_envs / create_envs_snr (algos/test_environment/eval/test_env_gen.py:13-69) from ONE batched reset: the
    sampled scenarios leave as the reference's joblib dict, come back through the loader, and drive the evaluator; the
    oracle loaded with the same file agrees with the GPU env bit for bit."""
    from radiation_ppo_b200 import scenario_io as sio

    n, k = 4000, 3
    env = rp.RadSearch(obstruction_count=k, enforce_grid_boundaries=True, num_envs=n, seed=21)
    arr = env.scenario_arrays()
    assert (arr["num_obs"] == k).all() and (np.linalg.norm(arr["src"] - arr["det"], axis=1) >= 1000).all()
    idx = sio.select_by_snr(arr, 40, "none")
    picked = {key: v[idx] for key, v in arr.items()}
    path = str(tmp_path / f"test_env_dict_obs{k}")
    sio.save_test_env_dict(path, sio.to_env_dict(picked))
    d = sio.load_test_env_dict(path)
    assert len(d) == 40 and len(d["env_0"][4]) == k
    ev = rp.MonteCarloEvaluator(env_dict=d, montecarlo_runs=3, steps_per_episode=30, obstruction_count=k, seed=5,
                                enforce_grid_boundaries=True)
    res = ev.run(rp.uniform_policy(0))
    assert res.episode_length.shape == (40, 3) and int(res.episode_length.max()) <= 30
    ob = co.OracleBatch(40, co.default_config(obstruction_count=k, enforce=1))
    back = sio.scenario_arrays(d, k_max=k, with_obstacles=True)
    ob.load_scenarios(back["src"], back["det"], back["intensity"], back["bkg"], back["rects"], back["num_obs"])
    env2 = rp.RadSearch(obstruction_count=k, enforce_grid_boundaries=True, num_envs=40, seed=5)
    env2.load_scenarios(**back)
    np.testing.assert_array_equal(env2._best[0].cpu().numpy(), [ob.shortest_path(i, back["det"][i]) for i in range(40)])


def test_episode_stats_follow_the_training_loops_bookkeeping():
    """radiation_ppo_b200.EpisodeStats against the per-env bookkeeping of train.py:359-527 done in numpy on an oracle
    rollout: EpRet / EpLen of the episodes that ended by terminal or timeout (not by the epoch cut), DoneCount, OutOfBound."""
    N, A, T, ML = 1024, 2, 90, 20
    env = rp.RadSearch(obstruction_count=2, enforce_grid_boundaries=True, number_agents=A, num_envs=N, seed=12,
                       steps_per_episode=ML, auto_reset=True)
    ob = co.OracleBatch(N, co.default_config(n_agents=A, obstruction_count=2, enforce=1, max_ep_len=ML), seed=12)
    ob.reset()
    st = rp.EpisodeStats(N, A, env.device)
    rng = np.random.default_rng(4)
    ep_ret = np.zeros(N); ep_len = np.zeros(N, np.int64)
    rets, lens, done_cnt, oob_cnt = [], [], np.zeros(A), np.zeros(A)
    src = None
    for t in range(T):
        # walk towards the lower-left wall now and then so that out-of-bounds events occur
        acts = np.where(rng.random((N, A)) < 0.3, 7, rng.integers(0, 8, size=(N, A)))
        epoch_end = t == T - 1
        env.step_batch(torch.as_tensor(acts, dtype=torch.int32, device=env.device), epoch_end=epoch_end)
        st.update(env.reward, env.team_reward, env.done_flags, env.info_flags, env.ended)
        ob.step(acts, env._ctr)
        e, o = ob.envs, ob.outs
        ep_ret += o["team_reward"].astype(np.float32).astype(np.float64)
        ep_len += 1
        done_cnt += (o["done"][:, :A] != 0).sum(0)
        oob_cnt += (e["oob"][:, :A] != 0).sum(0)
        over = (e["done"] == 1) | (e["ep_len"] == ML)
        rets += list(ep_ret[over]); lens += list(ep_len[over])
        mask = over | epoch_end
        ep_ret[mask] = 0; ep_len[mask] = 0
        if mask.any():
            ob.reset(mask=mask, new_obstacles=np.full(N, epoch_end))
    s = {k: v.cpu().numpy() for k, v in st.epoch_summary().items()}
    rets, lens = np.array(rets), np.array(lens)
    assert s["Episodes"][0] == len(rets) > N
    np.testing.assert_allclose(s["AverageEpRet"], rets.mean(), rtol=1e-12)
    np.testing.assert_allclose(s["StdEpRet"], rets.std(), rtol=1e-9)
    np.testing.assert_allclose(s["MinEpRet"], rets.min(), rtol=0, atol=1e-12)
    np.testing.assert_allclose(s["MaxEpRet"], rets.max(), rtol=0, atol=1e-12)
    np.testing.assert_allclose(s["EpLen"], lens.mean(), rtol=1e-12)
    np.testing.assert_array_equal(s["DoneCount"], done_cnt)
    np.testing.assert_array_equal(s["OutOfBound"], oob_cnt)
    assert oob_cnt.sum() > 0 and float(st._acc.abs().sum()) == 0.0
