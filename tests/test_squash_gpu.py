"""Row Q: tanh-squashed Gaussian head (PPOGaussianPolicyBase.sample, src/nets/nets.py:90-105) on the device vs the
reference-generated golden vectors and the restated oracle."""
import os

import numpy as np
import pytest
import torch

from aur_ppo_b200 import kernels
from oracle import ppo_ref as R

pytestmark = pytest.mark.gpu


def test_squashed_sample_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "squash.npz"))
    x, act = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["act"]).cuda()
    y, lp, mean, ent = kernels.squashed_gaussian_sample(x[:, :5].contiguous(), x[:, 5:].contiguous(), act)
    np.testing.assert_allclose(y.cpu().numpy(), g["a"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(lp.cpu().numpy(), g["logp"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(mean.cpu().numpy(), g["mean"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(ent.cpu().numpy(), g["ent"], rtol=1e-6)


@pytest.mark.parametrize("B,A", [(1, 1), (1000, 5), (4096, 16), (65537, 3)])
def test_squashed_sample_sampled_mode_replays_on_oracle(B, A):
    g = torch.Generator().manual_seed(B + A)
    mean = torch.randn(B, A, generator=g).cuda() * 0.7
    log_std = (torch.rand(B, A, generator=g) - 0.8).cuda()
    y, lp, mo, ent, pre = kernels.squashed_gaussian_sample(mean, log_std, None, seed=5, stream_id=9, return_pre_tanh=True)
    # replay the drawn pre-tanh actions on the checker
    y2, lp2, mo2, ent2 = R.squashed_sample(mean.cpu(), log_std.cpu(), pre.cpu())
    np.testing.assert_allclose(y.cpu().numpy(), y2.numpy(), rtol=1e-6, atol=1e-7)
    # log(1 - y^2 + 1e-6) is ill-conditioned once tanh saturates (one ulp of y moves it by up to 6e-8 / (1 - y^2 + 1e-6)):
    # tight tolerance where |action| < 3, conditioning-scaled tolerance elsewhere
    tame = (pre.abs().max(dim=1).values < 3.0).cpu().numpy()
    np.testing.assert_allclose(lp.cpu().numpy()[tame], lp2.numpy()[tame], rtol=1e-5, atol=2e-5)
    cond = (1.2e-7 / (1.0 - y.double() ** 2 + 1e-6)).sum(dim=1, keepdim=True).cpu().numpy()
    assert (np.abs(lp.cpu().numpy() - lp2.numpy()) <= 2e-5 + 1e-5 * np.abs(lp2.numpy()) + 2 * cond).all()
    np.testing.assert_allclose(ent.cpu().numpy(), ent2.numpy(), rtol=1e-6)
    # and the same rows fed back as given actions reproduce the sampled outputs exactly
    y3, lp3, _, _ = kernels.squashed_gaussian_sample(mean, log_std, pre)
    assert torch.equal(y3, y) and torch.equal(lp3, lp)
    if B >= 4096:   # the noise is standard normal
        z = ((pre - mean) / log_std.exp()).double().flatten()
        assert abs(z.mean().item()) < 5 / np.sqrt(z.numel()) and abs(z.var().item() - 1) < 0.02
        y4 = kernels.squashed_gaussian_sample(mean, log_std, None, seed=5, stream_id=10)[0]
        assert not torch.equal(y4, y)


def test_squashed_sample_rejects_bad_shapes():
    from aur_ppo_b200 import _lib
    with pytest.raises(_lib.AurError):
        kernels.squashed_gaussian_sample(torch.zeros(4, 17, device="cuda"), torch.zeros(4, 17, device="cuda"))
    with pytest.raises(_lib.AurError):
        kernels.squashed_gaussian_sample(torch.zeros(4, 3, device="cuda"), torch.zeros(4, 2, device="cuda"))
