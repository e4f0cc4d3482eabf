"""The C4 restatement used as the row-X checker is exactly equivariant (CPU)."""
import math

import pytest
import torch

from oracle import equiv_ref as Q


def test_expansions_have_reference_channel_counts():
    p = Q.init_params(0)
    chans = [Q.expand_trivial_to_regular(p["actor.enc0.psi"]).shape[0]] + [
        Q.expand_regular_to_regular(p[f"actor.enc{l}.psi"]).shape[0] for l in range(1, 7)]
    assert chans == [64, 128, 256, 512, 1024, 512, 512]
    assert Q.expand_regular_to_regular(p["actor.enc1.psi"]).shape == (128, 64, 3, 3)


def test_actor_critic_are_c4_equivariant():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    p = Q.init_params(1, scale=1.5)
    obs = torch.rand(1, 1, 128, 128) * 0.32
    state = torch.tensor([1.0])
    x = Q.cat_obs(state, obs)
    with torch.no_grad():
        mean, log_std = Q.actor_forward(p, x)
        v = Q.critic_forward(p, x)
        xr = torch.rot90(x, 1, dims=(2, 3))
        mean_r, log_std_r = Q.actor_forward(p, xr)
        v_r = Q.critic_forward(p, xr)
    # invariant outputs: p, dz, dtheta means, all log-stds, value
    torch.testing.assert_close(mean_r[:, [0, 3, 4]], mean[:, [0, 3, 4]], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(log_std_r, log_std, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(v_r, v, rtol=1e-4, atol=1e-5)
    # (dx, dy) transform with the standard representation: a 90-degree rotation, same norm
    dxy, dxy_r = mean[0, 1:3], mean_r[0, 1:3]
    assert abs(dxy.norm() - dxy_r.norm()) < 1e-4 * max(1.0, float(dxy.norm()))
    rot = torch.tensor([[0.0, -1.0], [1.0, 0.0]])
    ok_pos = torch.allclose(rot @ dxy, dxy_r, rtol=1e-3, atol=1e-5)
    ok_neg = torch.allclose(rot.T @ dxy, dxy_r, rtol=1e-3, atol=1e-5)
    assert ok_pos or ok_neg
    assert float(dxy.norm()) > 1e-6        # the test is not vacuous


def test_update_loss_runs_and_has_gradients_for_every_parameter():
    torch.manual_seed(0)
    p = {k: v.requires_grad_(True) for k, v in Q.init_params(2).items()}
    B = 2
    obs = torch.rand(B, 1, 128, 128) * 0.32
    state = torch.tensor([0.0, 1.0])
    action = torch.randn(B, 5)
    loss, stats = Q.update_loss(p, state, obs, action, torch.full((B,), -5.0), torch.tensor([1.0, -1.0]),
                                torch.zeros(B), torch.zeros(B))
    loss.backward()
    assert math.isfinite(stats["loss"])
    for k, v in p.items():
        assert v.grad is not None and torch.isfinite(v.grad).all(), k
