"""Multi-GPU plumbing: environments shard independently across ranks (no data-path collective); torch.distributed is
used only where the reference used MPI (algos/multiagent/rl_tools/mpi_tools.py:46-95, mpi_pytorch.py:26-49):
gradient averaging, advantage statistics, episode statistics, parameter broadcast."""
from __future__ import annotations

from typing import Dict, Iterable, Tuple

import torch
import torch.distributed as dist


def shard_range(total_envs: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous env range [lo, hi) of `rank`; the Philox key uses the GLOBAL env id (lo + local index), so results
    do not depend on world_size."""
    base, rem = divmod(total_envs, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _on() -> bool:
    return dist.is_available() and dist.is_initialized()


def global_mean_std(local_sum: torch.Tensor, local_n: torch.Tensor, centered_sumsq_fn, group=None):
    """mpi_statistics_scalar (mpi_tools.py:71-95): mean = allreduce(sum)/allreduce(n), then
    std = sqrt(allreduce(sum((x-mean)^2))/n).  `centered_sumsq_fn(mean)` returns this rank's sum((x-mean)^2)."""
    sn = torch.stack([local_sum.reshape(()), local_n.reshape(())]).to(torch.float64)
    if _on():
        dist.all_reduce(sn, group=group)
    mean = sn[0] / sn[1]
    ss = centered_sumsq_fn(mean).reshape(1).to(torch.float64)
    if _on():
        dist.all_reduce(ss, group=group)
    return mean, torch.sqrt(ss[0] / sn[1])


def average_gradients(params: Iterable[torch.nn.Parameter], group=None) -> None:
    """mpi_avg_grads (mpi_pytorch.py:26-33) as ONE flattened all-reduce instead of one blocking call per tensor."""
    if not _on() or dist.get_world_size(group) == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    flat /= dist.get_world_size(group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


def sync_params(module: torch.nn.Module, group=None) -> None:
    """sync_params (mpi_pytorch.py:43-49): broadcast rank 0's parameters."""
    if not _on() or dist.get_world_size(group) == 1:
        return
    for p in module.parameters():
        dist.broadcast(p.data, src=0, group=group)


def reduce_episode_stats(stats: Dict[str, torch.Tensor], group=None) -> Dict[str, torch.Tensor]:
    """Sum-reduce a dict of scalar counters (EpRet sum, EpLen sum, episode count, DoneCount, OutOfBound) in one call."""
    keys = sorted(stats)
    v = torch.stack([stats[k].reshape(()).to(torch.float64) for k in keys])
    if _on():
        dist.all_reduce(v, group=group)
    return {k: v[i] for i, k in enumerate(keys)}
