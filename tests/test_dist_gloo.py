"""CPU tests of the multi-rank host logic with world_size 2 over gloo: env sharding, the two-pass global advantage
statistics (mpi_tools.py:71-95 semantics), flattened gradient averaging (mpi_pytorch.py:26-33), parameter broadcast and
the episode-statistic reduction.  The env / GAE kernels themselves need no collective (envs shard independently)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from radiation_ppo_b200 import dist as rdist


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        total = 1001
        lo, hi = rdist.shard_range(total, rank, world)
        data = torch.from_numpy(np.random.default_rng(0).normal(2.0, 3.0, total)).float()
        x = data[lo:hi]
        mean, std = rdist.global_mean_std(x.double().sum(), torch.tensor(float(x.numel())),
                                          lambda m: ((x.double() - m) ** 2).sum())
        torch.manual_seed(0)
        net = torch.nn.Linear(4, 3)
        rdist.sync_params(net)
        net(torch.full((2, 4), float(rank + 1))).sum().backward()
        g_local = net.weight.grad.clone()
        rdist.average_gradients(net.parameters())
        stats = rdist.reduce_episode_stats({"ep_ret": torch.tensor(1.5 * (rank + 1)), "episodes": torch.tensor(3.0)})
        out[rank] = dict(lo=lo, hi=hi, mean=mean.item(), std=std.item(), g_local=g_local.numpy(),
                         g_avg=net.weight.grad.numpy().copy(), w=net.weight.detach().numpy().copy(),
                         ep_ret=stats["ep_ret"].item(), episodes=stats["episodes"].item())
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    assert (r0["lo"], r0["hi"], r1["lo"], r1["hi"]) == (0, 501, 501, 1001)
    data = np.random.default_rng(0).normal(2.0, 3.0, 1001).astype(np.float32).astype(np.float64)
    for r in (r0, r1):
        assert r["mean"] == pytest.approx(data.mean(), rel=1e-12)
        assert r["std"] == pytest.approx(data.std(), rel=1e-12)            # population std
        np.testing.assert_allclose(r["g_avg"], (r0["g_local"] + r1["g_local"]) / 2, rtol=1e-6)
        assert r["ep_ret"] == pytest.approx(4.5) and r["episodes"] == 6.0
    np.testing.assert_array_equal(r0["w"], r1["w"])


def test_shard_range_partitions_everything():
    for total in (1, 7, 1024, 1048576, 1000003):
        for world in (1, 2, 3, 4, 8):
            edges = [rdist.shard_range(total, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


def test_results_do_not_depend_on_sharding():
    """Global env ids key the Philox streams: two half-size shards reproduce the full batch (checked with the oracle,
    which shares the stream definition with the kernels)."""
    from oracle import c_oracle as co

    n = 64
    full = co.OracleBatch(n, co.default_config(obstruction_count=3), seed=5, env_id0=0)
    full.reset()
    parts = []
    for r in range(2):
        lo, hi = rdist.shard_range(n, r, 2)
        p = co.OracleBatch(hi - lo, co.default_config(obstruction_count=3), seed=5, env_id0=lo)
        p.reset()
        parts.append(p)
    acts = np.random.default_rng(1).integers(0, 8, (n, 1))
    full.step(acts, 1)
    parts[0].step(acts[:32], 1)
    parts[1].step(acts[32:], 1)
    np.testing.assert_array_equal(full.outs["obs"], np.concatenate([parts[0].outs["obs"], parts[1].outs["obs"]]))
    np.testing.assert_array_equal(full.envs["det"], np.concatenate([parts[0].envs["det"], parts[1].envs["det"]]))
