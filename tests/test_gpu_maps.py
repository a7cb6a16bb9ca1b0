"""GPU parity tests (pytest -m gpu) of the RAD-TEAM map observation kernels (rs_maps_update / rs_maps_reset) against the
golden vectors recorded from the reference's MapsBuffer and against the C restatement at batch sizes the reference
cannot run.  Bar: all map values bit-exact in float32."""
import numpy as np
import pytest
import torch

from oracle import c_oracle as co
from tests import parity_util as pu

pytestmark = pytest.mark.gpu

import radiation_ppo_b200 as rp  # noqa: E402

MAPS_FILES = {"ref_maps_a1": 1, "ref_maps_a4": 4, "ref_maps_a2_idle": 2}


def stacks_to_seven(actor, critic, i):
    """the reference's MapStack order for agent i's buffer: prediction, location, others, readings, visits, obstacles, combined"""
    return np.concatenate([actor[i], critic[:1]], axis=0)


@pytest.mark.parametrize("name", list(MAPS_FILES))
def test_maps_kernel_reproduces_reference_mapsbuffer(name):
    """tests/golden/ref_maps_*.npz: the reference's MapsBuffer.observation_to_map (RADTEAM_core.py:532-616) per agent
    buffer on oracle-env rollouts (float64 observations), incl. the bootstrap call and reset at episode ends; the kernel
    gets the float32 observations the env kernels write."""
    g = pu.load_golden(name)
    A = MAPS_FILES[name]
    mb = rp.BatchedMapsBuffer(1, A, 120)
    assert mb.map_dimensions == tuple(g["dims"]) and mb.resolution_accuracy == float(g["ra"])
    for t in range(len(g["obs"])):
        obs = torch.as_tensor(g["obs"][t].astype(np.float32)).reshape(1, A, 11)
        pred = torch.as_tensor(g["pred"][t].astype(np.float32)).reshape(1, A, 2)
        # the reference was fed float64 predictions; the kernel takes the fp32 a PFGRU produces: compare on those
        actor, critic = mb.update(obs.cuda(), pred.cuda())
        a, c = actor[0].cpu().numpy(), critic[0].cpu().numpy()
        for i in range(A):
            got = stacks_to_seven(a, c, i)
            want = g["maps"][t, i].copy()
            # prediction map: the golden cell came from the float64 prediction; recompute it from the fp32 one
            want[0] = 0
            px, py = (int(float(np.float32(g["pred"][t, i, k])) * float(g["ra"])) for k in (0, 1))
            want[0, px, py] = 1
            np.testing.assert_array_equal(got, want, err_msg=f"{name} call {t} buffer {i}")
        assert int(mb.status.sum()) == 0
        if g["reset_after"][t]:
            mb.reset()
            assert float(mb.actor_maps.abs().sum()) == 0 and float(mb.critic_maps.abs().sum()) == 0


def test_compat_mapsbuffer_interface():
    g = pu.load_golden("ref_maps_a2_idle")
    bufs = [rp.MapsBuffer(observation_dimension=11, steps_per_episode=120, number_of_agents=2) for _ in range(2)]
    assert bufs[0].map_dimensions == (27, 27) and bufs[0].base == 242
    for t in range(12):
        d = {i: g["obs"][t, i] for i in range(2)}
        for i in range(2):
            pred = tuple(float(np.float32(v)) for v in g["pred"][t, i])
            maps = bufs[i].observation_to_map(d, i, pred)
            assert len(maps) == 7
            for k in range(1, 7):
                np.testing.assert_array_equal(maps[k], g["maps"][t, i, k])


@pytest.mark.parametrize("N,A,T", [(512, 4, 70), (2048, 1, 50)])
def test_maps_kernel_matches_oracle_on_env_rollouts(N, A, T):
    """The env kernels feed the map kernels (config 4's pipeline): observations of a batched rollout with auto-reset
    -> rs_maps_update, bootstrap call on the final observation of ended episodes, rs_maps_reset for them; a sample of
    the environments is replayed through oracle/maps_oracle.c (one MapsOracle per agent buffer)."""
    ML = 25
    env = rp.RadSearch(obstruction_count=5, enforce_grid_boundaries=True, number_agents=A, num_envs=N, seed=9,
                       steps_per_episode=ML, auto_reset=True)
    mb = rp.BatchedMapsBuffer(N, A, ML, environment_scale=env.scale)
    sample = np.unique(np.random.default_rng(0).integers(0, N, 24))
    oracles = {n: [co.MapsOracle(A, ML) for _ in range(A)] for n in sample}
    gen = torch.Generator(device="cuda").manual_seed(5)
    rng = np.random.default_rng(1)

    def check(obs_t, pred_t, mask=None):
        actor, critic = mb.update(obs_t, pred_t, mask)
        a, c = actor.cpu().numpy(), critic.cpu().numpy()
        o, p = obs_t.cpu().numpy().astype(np.float64), pred_t.cpu().numpy()
        for n in sample:
            if mask is not None and not bool(mask[n]):
                continue
            # float64 observation of the reference: x * (1 / 2200) on the lattice coordinate
            o64 = o[n].copy()
            o64[:, 1:3] = np.rint(o[n, :, 1:3] / env.scale) * env.scale
            for i in range(A):
                want = oracles[n][i].observation_to_map(o64, i, (float(p[n, i, 0]), float(p[n, i, 1])))
                np.testing.assert_array_equal(stacks_to_seven(a[n], c[n], i), want, err_msg=f"env {n} buffer {i}")

    for t in range(T):
        pred = torch.rand(N, A, 2, generator=gen, device="cuda") * 1.2
        check(env.obs, pred)
        acts = torch.as_tensor(rng.integers(0, 9, size=(N, A)), dtype=torch.int32, device="cuda")
        env.step_batch(acts)
        ended = (env.ended & 4) != 0
        if bool(ended.any()):
            pred = torch.rand(N, A, 2, generator=gen, device="cuda") * 1.2
            check(env.final_obs, pred, mask=ended)                   # bootstrap call (train.py:476-480)
            mb.reset(mask=env.ended, mask_bits=4)                    # the env's ended[] with RS_E_RESET, no torch op
            for n in sample:
                if bool(ended[n]):
                    for b in oracles[n]:
                        b.reset()
    assert int(mb.status.sum()) == 0


def test_maps_unenforced_boundaries_147_grid_and_status_flags():
    """enforce_grid_boundaries=False: the 147 x 147 maps of RADTEAM_core.py:1740-1746; detectors that leave the search area
    to the low side get negative inflated coordinates, which numpy's indexing wraps once (the oracle restates that);
    predictions outside the map raise in the reference and are flagged here."""
    import radiation_ppo_b200.maps_buffer as mbm

    N, A, T = 256, 2, 60
    env = rp.RadSearch(obstruction_count=2, enforce_grid_boundaries=False, number_agents=A, num_envs=N, seed=3)
    ra = mbm.calculate_resolution_accuracy(0.01, env.scale)
    offset = env.scale * (500.0 + 120 * 100.0)
    mb = rp.BatchedMapsBuffer(N, A, 120, resolution_accuracy=ra, offset=offset, environment_scale=env.scale)
    assert mb.map_dimensions == (147, 147)
    oracles = {n: [co.MapsOracle(A, 120, (147, 147), ra) for _ in range(A)] for n in range(0, N, 16)}
    rng = np.random.default_rng(2)
    wrapped = 0
    for t in range(T):
        pred = torch.as_tensor(rng.uniform(0, 1.0, size=(N, A, 2)).astype(np.float32), device="cuda")
        actor, critic = mb.update(env.obs, pred)
        a, c = actor.cpu().numpy(), critic.cpu().numpy()
        o = env.obs.cpu().numpy().astype(np.float64)
        for n, bufs in oracles.items():
            o64 = o[n].copy()
            o64[:, 1:3] = np.rint(o[n, :, 1:3] / env.scale) * env.scale
            wrapped += int((o64[:, 1:3] < 0).any())
            for i in range(A):
                want = bufs[i].observation_to_map(o64, i, (float(pred[n, i, 0]), float(pred[n, i, 1])))
                np.testing.assert_array_equal(stacks_to_seven(a[n], c[n], i), want)
        # walk down-left so that some detectors cross x < 0 / y < 0
        acts = torch.as_tensor(rng.choice([0, 6, 7, 7], size=(N, A)), dtype=torch.int32, device="cuda")
        env.step_batch(acts)
    assert wrapped > 0 and int(mb.status.sum()) == 0
    bad = torch.full((N, A, 2), 9.0, device="cuda")                   # 9.0 * 22 = 198 > 147: outside the map
    mb.update(env.obs, bad)
    assert bool(((mb.status & 4) != 0).all())


def test_reference_unit_test_scenario_observation_to_map():
    """The scenario of the reference's own unit test (unit_tests/test_RADTEAM_core.py:576-704: three agents, two calls),
    through the reference-style MapsBuffer interface.  That test predates the source-prediction map (it indexes a six-map
    stack); the assertions are the same, with the stack order of the shipped code (RADTEAM_core.py:606-616)."""
    maps = rp.MapsBuffer(observation_dimension=11, steps_per_episode=120, number_of_agents=3)
    ra = maps.resolution_accuracy

    def deflate(c):                                       # MapsBuffer._deflate_coordinates (RADTEAM_core.py:717-746)
        return (float(c[0] / ra), float(c[1] / ra))

    def call(count, c01, c2):
        s1, s2 = deflate(c01), deflate(c2)
        obs = {0: np.array([count, s1[0], s1[1], 0., 0., 0., 0.1, 0., 0., 0., 0.], dtype=np.float32),
               1: np.array([count, s1[0], s1[1], 0., 0., 0., 0.1, 0., 0., 0., 0.], dtype=np.float32),
               2: np.array([count, s2[0], s2[1], 0., 0., 0., 0.1, 0., 0., 0., 0.], dtype=np.float32)}
        return maps.observation_to_map(obs, 0, (0.5, 0.5))

    pred, loc, others, readings, visits, obstacles, combo = call(1000.0, (0, 1), (0, 2))
    assert pred[11][11] == 1.0 and pred.sum() == 1.0
    assert loc[0][1] == 1.0 and np.delete(loc.ravel(), 1).max() == 0.0
    assert others[0][1] == 1.0 and others[0][2] == 1.0 and others.sum() == 2.0
    assert readings.max() == 0.0                          # first readings standardise to 0
    assert visits[0][1] > visits[0][2] > 0.0              # two agents visited (0, 1), one visited (0, 2)
    assert obstacles[0][1] == np.float32(0.1) and obstacles[0][2] == np.float32(0.1)
    assert combo[0][1] == 2.0 and combo[0][2] == 1.0 and combo.sum() == 3.0

    pred, loc, others, readings, visits, obstacles, combo = call(5000.0, (0, 3), (0, 4))
    assert loc[0][1] == 0.0 and loc[0][3] == 1.0
    assert others[0][1] == 0.0 and others[0][2] == 0.0 and others[0][3] == 1.0 and others[0][4] == 1.0
    assert readings[0][3] > 0.0 and readings[0][4] > 0.0
    assert all(visits[0][k] > 0.0 for k in (1, 2, 3, 4))
    assert all(obstacles[0][k] > 0.0 for k in (1, 2, 3, 4))
    assert combo[0][1] == 0.0 and combo[0][2] == 0.0 and combo[0][3] == 2.0 and combo[0][4] == 1.0
    maps.reset()
    assert all(float(np.abs(m).sum()) == 0.0 for m in call(0.0, (5, 5), (5, 5))[3:4])      # readings: a zero count -> z = 0
