"""The shape-generic minibatch update (csrc/update_generic.cu): every `--hidden_dim` / `--num_layers` of the reference CLI
(src/run_ppo.py:36,38) and observation / action widths up to 8, against the torch-autograd restatement of
src/ppo.py:220-269 (oracle/ppo_ref.py) and, on the headline shape, against the reference-generated golden file.
Tolerance: north_star's 1e-4 relative for losses and gradients."""
import os

import numpy as np
import pytest
import torch

from aur_ppo_b200 import _lib, kernels
from oracle import ppo_ref as R
from tests.helpers import flat_from_named, random_policy

pytestmark = pytest.mark.gpu


def _grad_close(got, want, rtol=1e-4):
    np.testing.assert_allclose(got, want, rtol=rtol, atol=rtol * float(np.abs(want).max()) + 1e-9)


SHAPES = [  # obs_dim, act_dim, hidden, layers, continuous, m, B
    (4, 2, 64, 3, False, 1000, 5000),
    (4, 2, 64, 1, False, 333, 1024),
    (4, 2, 64, 10, False, 700, 2048),      # the shape of the shipped actor_critic_10.pt
    (4, 2, 64, 16, False, 200, 512),
    (3, 1, 32, 2, True, 777, 4096),
    (6, 3, 128, 2, False, 900, 4096),      # Acrobot-sized
    (8, 4, 64, 2, False, 640, 2048),       # the shape of the shipped actor_critic_2.pt (state 8, 4 actions)
    (8, 8, 256, 2, True, 500, 2048),
    (2, 3, 20, 4, False, 129, 512),        # hidden not a multiple of the 32-row pass / 64-wide gradient block
    (5, 2, 200, 3, True, 257, 1024),
    (4, 2, 64, 2, False, 1000, 4096),      # outside the generic range only by choice: forced through impl 3 below
]


@pytest.mark.parametrize("obs_dim,act_dim,hidden,layers,cont,m,B", SHAPES)
def test_generic_update_vs_oracle(obs_dim, act_dim, hidden, layers, cont, m, B):
    L = _lib.lib()
    prev = L.aur_ppo_update_get_impl()
    assert L.aur_ppo_update_set_impl(3) == 0
    try:
        pol, named = random_policy(obs_dim, act_dim, hidden, layers, cont, seed=hidden + layers)
        names = list(named.keys())
        desc = kernels.policy_desc(obs_dim, act_dim, hidden, layers, cont)
        g = torch.Generator().manual_seed(m)
        b_obs = torch.randn(B, obs_dim, generator=g) * 0.7
        b_act = torch.randn(B, act_dim, generator=g) if cont else torch.randint(0, act_dim, (B,), generator=g).float()
        with torch.no_grad():
            _, lp0, _, v0 = pol.evaluate(b_obs, b_act)
        b_lp = lp0 + 0.25 * torch.randn(B, generator=g)
        b_adv = torch.randn(B, generator=g) * 3 + 0.5
        b_ret = torch.randn(B, generator=g)
        b_val = v0.flatten() + 0.4 * torch.randn(B, generator=g)
        idx = torch.randperm(B, generator=g)[:m]
        opt = R.RefAdam(pol.tensors(), lr=3e-4, eps=1e-5)
        params = torch.from_numpy(flat_from_named(named)).cuda()
        up = kernels.Updater(desc, params)
        dbuf = [t.cuda().contiguous() for t in (b_obs, b_act, b_lp, b_adv, b_ret, b_val)]
        for step in range(2):
            stats_ref, raw, _, _ = R.ppo_update_step(pol, opt, b_obs[idx], b_act[idx], b_lp[idx], b_adv[idx], b_ret[idx],
                                                     b_val[idx], max_grad_norm=0.5)
            grads = up.grad(*dbuf, idx.to(torch.int32).cuda()).clone()
            again = up.grad(*dbuf, idx.to(torch.int32).cuda()).clone()
            assert torch.equal(grads, again)                       # fixed accumulation order: bit-reproducible
            stats = up.apply(3e-4, 0.5).cpu().numpy()
            want = flat_from_named({n: r.numpy() for n, r in zip(names, raw)})
            _grad_close(grads[:up.P].cpu().numpy(), want)
            for i, k in enumerate(kernels.STAT_NAMES):
                np.testing.assert_allclose(stats[i], stats_ref[k], rtol=1e-4, atol=2e-6, err_msg=k)
            np.testing.assert_allclose(params.cpu().numpy(), flat_from_named({n: pol.p[n].detach().numpy() for n in names}),
                                       rtol=1e-5, atol=1e-6)
    finally:
        L.aur_ppo_update_set_impl(prev)


@pytest.mark.parametrize("tag", ["disc", "disc_big", "cont", "disc_nonorm", "disc_novclip"])
def test_generic_update_matches_reference_golden(golden_dir, tag):
    """The generic kernel forced onto the headline shape reproduces the reference's own gradients (tests/golden/update.npz)."""
    L = _lib.lib()
    prev = L.aur_ppo_update_get_impl()
    assert L.aur_ppo_update_set_impl(3) == 0
    try:
        g = np.load(os.path.join(golden_dir, "update.npz"))
        names = [str(n) for n in g[f"{tag}_names"]]
        named0 = {n: g[f"{tag}_p0_{n}"] for n in names}
        cont = "actor_logstd" in names
        obs, act = g[f"{tag}_obs"], g[f"{tag}_act"].astype(np.float32)
        clip, ent_c, vf_c, mgn, lr, norm_adv, clip_vloss = [float(v) for v in g[f"{tag}_hyper"]]
        desc = kernels.policy_desc(obs.shape[1], 1 if cont else 2, 64, 2, cont)
        params = torch.from_numpy(flat_from_named(named0)).cuda()
        up = kernels.Updater(desc, params)
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        bufs = [dev(obs), dev(act), dev(g[f"{tag}_oldlp"]), dev(g[f"{tag}_adv"]), dev(g[f"{tag}_ret"]), dev(g[f"{tag}_vold"])]
        idx = torch.arange(obs.shape[0], dtype=torch.int32, device="cuda")
        for step in range(2):
            grads = up.grad(*bufs, idx, clip_coeff=clip, entropy_coeff=ent_c, value_coeff=vf_c, norm_adv=bool(norm_adv),
                            clip_vloss=bool(clip_vloss)).clone()
            stats = up.apply(lr, mgn).cpu().numpy()
            _grad_close(grads[:up.P].cpu().numpy(), flat_from_named({n: g[f"{tag}_s{step}_g_{n}"] for n in names}))
            want = g[f"{tag}_s{step}_stats"]
            got = [stats[0], stats[1], stats[2], stats[7], stats[3], stats[4], stats[5], stats[6]]
            np.testing.assert_allclose(got, want, rtol=1e-4, atol=2e-6)
            np.testing.assert_allclose(params.cpu().numpy(), flat_from_named({n: g[f"{tag}_s{step}_p_{n}"] for n in names}),
                                       rtol=1e-5, atol=1e-6)
    finally:
        L.aur_ppo_update_set_impl(prev)


def test_unsupported_shapes_fail_loudly():
    for bad in (kernels.policy_desc(9, 2, 64, 2, False), kernels.policy_desc(4, 2, 66, 2, False),
                kernels.policy_desc(4, 2, 512, 2, False), kernels.policy_desc(4, 2, 256, 16, False)):
        with pytest.raises(_lib.AurError):
            kernels.Updater(bad, torch.zeros(kernels.policy_param_count(bad), device="cuda"))
