"""In-kernel data-parallel exchange (parallel.PeerExchange): two ranks push their advantage moments and packed
gradient sums into each other's CUDA-IPC exchange areas and the Adam kernel gathers them; the result must equal
the single-GPU update of the whole minibatch.  Runs with both ranks on cuda:0 when the box has one GPU (IPC works
across processes on one device), on cuda:0 / cuda:1 otherwise.  Control plane: gloo on 127.0.0.1."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    return port


def _worker(rank, world, port, cont, out_dir, shape=None, ahead=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = rank % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    from aur_ppo_b200 import kernels, parallel
    from tests.helpers import flat_from_named, random_policy
    obs_dim, act_dim = (3, 1) if cont else (4, 2)
    hidden, layers = 64, 2
    if shape is not None:                                   # a shape of the generic kernel (update_generic.cu)
        obs_dim, act_dim, hidden, layers = shape
    _, named = random_policy(obs_dim, act_dim, hidden, layers, cont, seed=4)
    desc = kernels.policy_desc(obs_dim, act_dim, hidden, layers, cont)
    flat0 = torch.from_numpy(flat_from_named(named)).cuda()
    B, m = 6000, 4096                                       # the global minibatch: m of B rows, same on every rank
    g = torch.Generator().manual_seed(12)
    obs = (torch.randn(B, obs_dim, generator=g) * 0.5).cuda()
    act = (torch.randn(B, act_dim, generator=g) if cont else torch.randint(0, act_dim, (B,), generator=g).float()).cuda()
    oldlp = (-0.7 + 0.2 * torch.randn(B, generator=g)).cuda(); adv = (torch.randn(B, generator=g) * 2 + 1).cuda()
    ret = torch.randn(B, generator=g).cuda(); vold = torch.randn(B, generator=g).cuda()
    idx = torch.randperm(B, generator=g)[:m].to(torch.int32).cuda()
    bufs = (obs, act, oldlp, adv, ret, vold)
    plan = parallel.ShardPlan(world, rank, num_envs=world, num_steps=m // world, num_minibatches=1)
    ex = parallel.PeerExchange(plan, desc)
    up = kernels.Updater(desc, flat0.clone(), exchange=ex)
    share = m // world
    mine = idx[rank * share:(rank + 1) * share].contiguous()
    grads, stats = [], []
    if ahead:
        # the advantage moments of all three "minibatches" of the iteration in one launch, exchanged ONCE
        # (aur_ppo_adv_moments_multi); entry 1 gets a different index list so that a wrong entry would show
        other = idx.flip(0)[rank * share:(rank + 1) * share].contiguous()
        up.prepare_moments(adv, torch.stack([mine, other, mine]).contiguous())
    for step in range(3):
        kw = dict(moments_index=2 * (step & 1)) if ahead else {}
        grads.append(up.grad(*bufs, mine, m_total=m, **kw).clone())  # local sums (before the gather)
        stats.append(up.apply(2.5e-4, 0.5).clone())
        grads.append(up.grads.clone())                                # world sums, written back by the Adam kernel
    torch.cuda.synchronize()
    assert ex.status() == 0
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), params=up.params.cpu().numpy(), stats=torch.stack(stats).cpu().numpy(),
             gsum=torch.stack(grads[1::2]).cpu().numpy())
    if rank == 0:
        ref = kernels.Updater(desc, flat0.clone())
        rg, rs = [], []
        for step in range(3):
            rg.append(ref.grad(*bufs, idx, m_total=m).clone())
            rs.append(ref.apply(2.5e-4, 0.5).clone())
        np.savez(os.path.join(out_dir, "single.npz"), params=ref.params.cpu().numpy(), stats=torch.stack(rs).cpu().numpy(),
                 gsum=torch.stack(rg).cpu().numpy())
    ex.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("cont,shape,ahead", [(False, None, False), (True, None, False), (False, (6, 3, 32, 3), False),
                                              (True, (5, 2, 128, 2), False), (False, None, True), (True, None, True),
                                              (False, (6, 3, 32, 3), True)])
def test_peer_exchange_matches_single_gpu(tmp_path, cont, shape, ahead):
    """ahead=True: the advantage moments travel once per iteration (prepare_moments) instead of once per minibatch."""
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), cont, str(tmp_path), shape, ahead), nprocs=world, join=True)
    r0, r1, one = [np.load(tmp_path / f) for f in ("rank0.npz", "rank1.npz", "single.npz")]
    # every rank ends with bit-identical parameters and statistics (same sums in the same order)
    np.testing.assert_array_equal(r0["params"], r1["params"])
    np.testing.assert_array_equal(r0["stats"], r1["stats"])
    np.testing.assert_array_equal(r0["gsum"], r1["gsum"])
    # and they equal the single-GPU update of the whole minibatch up to fp32 summation order
    scale = np.abs(one["gsum"]).max()
    np.testing.assert_allclose(r0["gsum"], one["gsum"], rtol=2e-5, atol=2e-6 * scale)
    np.testing.assert_allclose(r0["stats"], one["stats"], rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(r0["params"], one["params"], rtol=1e-5, atol=1e-6)
