"""N > 1 host logic on CPU (gloo, world_size 2): the shard plan and the one-exchange-per-minibatch
protocol.  The per-rank compute is the oracle standing in for the CUDA kernels; what is tested is
that partial [grads | stats] sums with 1/m_total seeds and all-reduced advantage moments reproduce
the single-process minibatch step exactly as DESIGN.md section 6 states."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from aur_ppo_b200 import parallel


def test_shard_plan_arithmetic():
    p = parallel.ShardPlan(8, 3, 1048576, 128, 4)
    assert p.local_envs == 131072 and p.env_id0 == 393216 and p.env_ids[-1] == 524287
    assert p.batch_size == 134217728 and p.local_batch == 16777216
    assert p.minibatch_size == 33554432 and p.local_minibatch == 4194304
    one = parallel.ShardPlan(1, 0, 4, 128, 4)
    assert (one.local_envs, one.minibatch_size, one.local_minibatch) == (4, 128, 128)
    assert parallel.make_allreduce(one) is None
    with pytest.raises(ValueError):
        parallel.ShardPlan(8, 0, 100, 128, 4)
    # every env id is owned exactly once
    ids = [i for r in range(4) for i in parallel.ShardPlan(4, r, 64, 8, 2).env_ids]
    assert ids == list(range(64))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import ppo_ref as R
    from tests.helpers import flat_from_named, random_policy
    torch.set_num_threads(1)
    plan = parallel.current_plan(num_envs=8, num_steps=16, num_minibatches=2)
    assert (plan.world_size, plan.rank, plan.local_envs, plan.env_id0) == (world, rank, 8 // world, rank * (8 // world))
    reduce_fn = parallel.make_allreduce(plan)
    pol, named = random_policy(4, 2, 64, 2, False, seed=21)
    names = list(named)
    # rank-specific init, then broadcast: everyone must end with rank 0's weights
    mod = torch.nn.Linear(3, 3)
    torch.manual_seed(100 + rank); torch.nn.init.normal_(mod.weight)
    parallel.broadcast_parameters(mod, plan)
    torch.manual_seed(100); ref = torch.nn.Linear(3, 3); torch.manual_seed(100); torch.nn.init.normal_(ref.weight)
    assert torch.equal(mod.weight, ref.weight)
    # the global minibatch (same on every rank), of which this rank owns a contiguous share
    g = torch.Generator().manual_seed(5)
    m = plan.minibatch_size
    obs = torch.randn(m, 4, generator=g); act = torch.randint(0, 2, (m,), generator=g).float()
    oldlp = -0.7 + 0.2 * torch.randn(m, generator=g); adv = torch.randn(m, generator=g) * 2 + 1
    ret = torch.randn(m, generator=g); vold = torch.randn(m, generator=g)
    lo, hi = rank * plan.local_minibatch, (rank + 1) * plan.local_minibatch
    sl = slice(lo, hi)
    # exchange 1: advantage moments (sum, sum of squares, count) in fp64
    mom = torch.tensor([adv[sl].double().sum(), (adv[sl].double() ** 2).sum(), float(hi - lo)], dtype=torch.float64)
    reduce_fn(mom)
    mean = mom[0] / mom[2]
    std = torch.sqrt((mom[1] - mom[0] * mean) / (mom[2] - 1))
    advn = ((adv[sl] - mean.float()) / (std.float() + 1e-8))
    # local compute with 1/m_total seeds: sum-reduced losses == means over the whole minibatch
    pol.requires_grad_(True)
    loss, stats, _, _ = R.ppo_loss(pol, obs[sl], act[sl], oldlp[sl], advn, ret[sl], vold[sl], norm_adv=False)
    scale = (hi - lo) / m
    grads = torch.autograd.grad(loss * scale, pol.tensors())
    packed = torch.cat([torch.from_numpy(flat_from_named({n: gr.numpy() for n, gr in zip(names, grads)})),
                        torch.tensor([stats["policy_loss"] * (hi - lo), stats["value_loss"] * (hi - lo),
                                      stats["entropy"] * (hi - lo)])])
    reduce_fn(packed)          # exchange 2: ONE packed buffer
    if rank == 0:
        np.save(os.path.join(out_dir, "packed.npy"), packed.numpy())
    dist.destroy_process_group()


def test_two_rank_exchange_reproduces_single_process_step(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    packed = np.load(tmp_path / "packed.npy")
    from oracle import ppo_ref as R
    from tests.helpers import flat_from_named, random_policy
    pol, named = random_policy(4, 2, 64, 2, False, seed=21)
    g = torch.Generator().manual_seed(5)
    m = 8 * 16 // 2
    obs = torch.randn(m, 4, generator=g); act = torch.randint(0, 2, (m,), generator=g).float()
    oldlp = -0.7 + 0.2 * torch.randn(m, generator=g); adv = torch.randn(m, generator=g) * 2 + 1
    ret = torch.randn(m, generator=g); vold = torch.randn(m, generator=g)
    pol.requires_grad_(True)
    loss, stats, _, _ = R.ppo_loss(pol, obs, act, oldlp, adv, ret, vold)
    grads = torch.autograd.grad(loss, pol.tensors())
    want = flat_from_named({n: gr.numpy() for n, gr in zip(named, grads)})
    np.testing.assert_allclose(packed[:-3], want, rtol=2e-4, atol=1e-7)
    np.testing.assert_allclose(packed[-3:] / m, [stats["policy_loss"], stats["value_loss"], stats["entropy"]], rtol=1e-5)
