"""TEST HELPERS (world_size-2 gloo tests; the product's exchange is csrc/ppo_dp.cuh + drone_rl_b200/ppo.py).  Host-side helpers of the multi-GPU path (one process per GPU, torch.distributed): how envs are
sharded and how per-rank sums combine.  Pure torch -- they run on CPU under gloo in the tests and
on GPU under NCCL in bench.py / ppo.py.  Envs shard by contiguous global-id ranges with NO
collective in the step; the only exchanges are the flat-gradient / advantage-sum
all-reduces of the PPO update and scalar reductions for logging and timing."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(global_envs: int, rank: int, world: int):
    """Contiguous shard [first, first + count) of rank `rank`; sizes differ by at most one."""
    base, extra = divmod(global_envs, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def allreduce_sum_(t: torch.Tensor) -> torch.Tensor:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def max_over_ranks(value: float, device=None) -> float:
    """Timing rule: a multi-GPU duration is the MAX over ranks."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def advantage_moments(stats: torch.Tensor):
    """[sum, sum of squares, count] (already all-reduced) -> (mean, 1 / (unbiased std + 1e-8)) as SB3
    normalises advantages; the same formula runs on the device in ppo_grad_kernel."""
    s, q, n = (float(x) for x in stats)
    mean = s / n
    var = max((q - s * mean) / (n - 1.0), 0.0)
    return mean, 1.0 / (var ** 0.5 + 1e-8)
