"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: sharding by global env id and the
data-parallel PPO contract -- sum-form gradients + global advantage moments, all-reduced, give the
single-process update (checked here with the float64 torch oracle standing in for the kernels;
tests/test_gpu_ppo.py::test_gradient_is_shard_additive_and_deterministic checks the kernels)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import dist_helpers as du
from oracle import drone_oracle as do
from oracle import ppo_oracle as po


def test_shard_ranges_partition_the_global_ids():
    for total, world in [(64 * 2 ** 20, 8), (4096, 2), (10, 4), (7, 8)]:
        spans = [du.shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == total
        for (f0, c0), (f1, _) in zip(spans, spans[1:]):
            assert f0 + c0 == f1
        assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _batch(seed, B):
    g = torch.Generator().manual_seed(seed)
    theta = po.init_params(0, torch.float64) + 0.05 * torch.randn(po.N_PARAMS, generator=g, dtype=torch.float64)
    obs = torch.randn(B, 15, generator=g, dtype=torch.float64)
    mean, value, log_std = po.forward(theta, obs)
    act = mean + torch.randn(B, 4, generator=g, dtype=torch.float64)
    old = po.log_prob(mean, log_std, act) + 0.1 * torch.randn(B, generator=g, dtype=torch.float64)
    adv = torch.randn(B, generator=g, dtype=torch.float64) * 2 + 1
    ret = value + torch.randn(B, generator=g, dtype=torch.float64)
    return theta, (obs, act, old, adv, ret)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B = 512
    theta, batch = _batch(5, B)
    first, count = du.shard_range(B, rank, world)
    shard = tuple(t[first:first + count] for t in batch)
    # 1. global advantage moments from all-reduced sums
    adv = shard[3]
    stats = torch.tensor([adv.sum(), (adv * adv).sum(), float(count)], dtype=torch.float64)
    du.allreduce_sum_(stats)
    mean, inv_std = du.advantage_moments(stats)
    # 2. sum-form gradient of the shard, all-reduced, scaled by 1 / global count
    p = theta.clone().requires_grad_(True)
    loss, _ = po.ppo_loss(p, *shard, adv_mean=torch.tensor(mean, dtype=torch.float64),
                          adv_std=torch.tensor(1.0 / inv_std - 1e-8, dtype=torch.float64))
    (g,) = torch.autograd.grad(loss * count, p)
    du.allreduce_sum_(g)
    g = g / B
    # 3. timing rule
    t = du.max_over_ranks(1.0 + rank)
    # 4. env sharding: each rank steps its own id range; the union equals one big env
    env = do.BatchedDroneOracle(count, do.SINGLE, seed=3, env_offset=first)
    obs = torch.from_numpy(env.reset())
    gathered = [torch.zeros(B // world, 15) for _ in range(world)]
    dist.all_gather(gathered, obs)
    if rank == 0:
        q.put((g.numpy(), mean, inv_std, t, torch.cat(gathered).numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_update_equals_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    g, mean, inv_std, t, obs = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    theta, batch = _batch(5, 512)
    p = theta.clone().requires_grad_(True)
    loss, _ = po.ppo_loss(p, *batch)
    (ref,) = torch.autograd.grad(loss, p)
    np.testing.assert_allclose(g, ref.numpy(), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(mean, float(batch[3].mean()), rtol=1e-12)
    np.testing.assert_allclose(1.0 / inv_std, float(batch[3].std()) + 1e-8, rtol=1e-12)
    assert t == 2.0
    whole = do.BatchedDroneOracle(512, do.SINGLE, seed=3)
    assert np.array_equal(obs, whole.reset())
