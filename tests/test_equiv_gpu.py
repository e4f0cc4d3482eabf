"""Equivariant encoder kernels vs the torch fp32 restatement (GPU).  bf16 operands with fp32
accumulation: compared against fp32 conv2d of the SAME bf16-rounded inputs/weights, tolerance 1e-2
relative (the bf16 rounding of the stored output)."""
import pytest
import torch
import torch.nn.functional as F

from aur_ppo_b200 import kernels
from oracle import equiv_ref as Q

pytestmark = pytest.mark.gpu


def test_expand_matches_oracle_expansion():
    g = torch.Generator().manual_seed(0)
    psi = torch.randn(8, 16, 4, 3, 3, generator=g)
    bias = torch.randn(8, generator=g)
    wmat, wt, b = kernels.equiv_expand_regular(psi.cuda(), bias.cuda(), want_wt=True)
    W = Q.expand_regular_to_regular(psi).bfloat16()                 # [32, 64, 3, 3]
    want = W.permute(0, 2, 3, 1).reshape(32, 9, 64)
    assert torch.equal(wmat.cpu(), want)
    want_t = torch.flip(W, dims=(2, 3)).permute(1, 2, 3, 0).reshape(64, 9, 32)
    assert torch.equal(wt.cpu(), want_t)
    assert torch.equal(b.cpu(), Q.expand_bias_regular(bias))


@pytest.mark.parametrize("B,H,Fi,Fo,pad,pool", [(3, 16, 16, 32, 1, True), (2, 64, 16, 32, 1, True), (5, 8, 32, 64, 1, False),
                                                (4, 8, 64, 32, 0, True), (2, 32, 32, 16, 1, False)])
def test_conv_layer_matches_conv2d(B, H, Fi, Fo, pad, pool):
    g = torch.Generator().manual_seed(B * 100 + H)
    Cin, Cout = Fi * 4, Fo * 4
    psi = torch.randn(Fo, Fi, 4, 3, 3, generator=g) * (2.0 / (Cin * 9)) ** 0.5
    bias = 0.1 * torch.randn(Fo, generator=g)
    x = torch.randn(B, Cin, H, H, generator=g).bfloat16()
    wmat, _, bias_ch = kernels.equiv_expand_regular(psi.cuda(), bias.cuda())
    Hb = H + 2 * pad
    inp = torch.zeros(B, Hb, Hb, Cin, dtype=torch.bfloat16, device="cuda")
    inp[:, pad:pad + H, pad:pad + H, :] = x.permute(0, 2, 3, 1).cuda()
    Ho = Hb - 2
    Hn = Ho // 2 if pool else Ho
    out = torch.zeros(B, Hn + 2, Hn + 2, Cout, dtype=torch.bfloat16, device="cuda")
    arg = torch.zeros(B, Hn, Hn, Cout, dtype=torch.uint8, device="cuda") if pool else None
    kernels.conv3x3_bf16(inp, wmat, bias_ch, 2 if pool else 1, out, 1, arg)
    W = Q.expand_regular_to_regular(psi).bfloat16().float()
    ref = F.relu(F.conv2d(x.float(), W, Q.expand_bias_regular(bias), padding=pad))
    if pool:
        ref, idx = F.max_pool2d(ref, 2, return_indices=True)
    got = out[:, 1:1 + Hn, 1:1 + Hn, :].permute(0, 3, 1, 2).float().cpu()
    torch.testing.assert_close(got, ref, rtol=1e-2, atol=1e-2)
    assert float(out[:, 0].abs().max()) == 0 and float(out[:, :, -1].abs().max()) == 0        # halo untouched
    if pool:
        # arg-max agrees wherever the maximum is unique by a margin
        yy = idx // Ho
        xx = idx % Ho
        want_w = ((yy & 1) * 2 + (xx & 1)).permute(0, 2, 3, 1)
        agree = (arg.cpu().long() == want_w).float().mean().item()
        assert agree > 0.97, agree


def test_conv0_direct_matches_conv2d():
    g = torch.Generator().manual_seed(3)
    B = 3
    psi = torch.randn(16, 2, 3, 3, generator=g) * 0.3
    bias = 0.1 * torch.randn(16, generator=g)
    obs = torch.rand(B, 1, 128, 128, generator=g) * 0.32
    state = torch.tensor([0.0, 1.0, 1.0])
    out = torch.zeros(B, 66, 66, 64, dtype=torch.bfloat16, device="cuda")
    arg = torch.zeros(B, 64, 64, 64, dtype=torch.uint8, device="cuda")
    kernels.equiv_conv0(obs.cuda(), state.cuda(), psi.cuda(), bias.cuda(), out, arg)
    x = Q.cat_obs(state, obs)
    ref = F.max_pool2d(F.relu(F.conv2d(x, Q.expand_trivial_to_regular(psi), Q.expand_bias_regular(bias), padding=1)), 2)
    got = out[:, 1:65, 1:65, :].permute(0, 3, 1, 2).float().cpu()
    torch.testing.assert_close(got, ref, rtol=1e-2, atol=2e-3)
    assert float(out[:, 0].abs().max()) == 0


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-12))


def test_full_update_gradients_match_oracle_autograd():
    """Whole row X on a small batch: forward values, loss statistics and EVERY parameter gradient
    against torch fp32 autograd of the restated model.  bf16 operands / fp32 accumulation through
    7 + 2 layers forward and backward: gradients are compared in relative L2 norm (5e-2) and
    direction (cosine > 0.995), scalars within 2e-2."""
    from aur_ppo_b200 import equiv
    B = 8
    torch.manual_seed(0)
    params = equiv.init_params(seed=5, scale=1.3)
    g = torch.Generator().manual_seed(1)
    obs = torch.rand(B, 1, 128, 128, generator=g) * 0.32
    state = (torch.rand(B, generator=g) > 0.5).float()
    action = torch.randn(B, 5, generator=g)
    adv, ret = torch.randn(B, generator=g), torch.randn(B, generator=g)
    cpu = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in params.items()}
    with torch.no_grad():
        lp0, _, v0 = Q.evaluate(cpu, state, obs, action)
    oldlp = lp0 + 0.15 * torch.randn(B, generator=g)
    vold = v0 + 0.3 * torch.randn(B, generator=g)
    loss, stats = Q.update_loss(cpu, state, obs, action, oldlp, adv, ret, vold)
    loss.backward()

    model = equiv.EquivActorCritic(params, B)
    dev = lambda t: t.cuda().contiguous()
    st = model.loss_and_grads(dev(state), dev(obs), dev(action), dev(oldlp), dev(adv), dev(ret), dev(vold)).cpu()
    with torch.no_grad():
        lp_ref, _, v_ref = Q.evaluate(cpu, state, obs, action)
    torch.testing.assert_close(model.value.cpu(), v_ref, rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(model.logp.cpu(), lp_ref, rtol=2e-2, atol=5e-2)
    assert abs(st[0] - stats["policy_loss"]) < 2e-2 * max(1, abs(stats["policy_loss"]))
    assert abs(st[1] - stats["value_loss"]) < 2e-2 * max(1, abs(stats["value_loss"]))
    assert abs(st[2] - stats["entropy"]) < 2e-2 * max(1, abs(stats["entropy"]))
    worst = {}
    for k, p in cpu.items():
        got, want = model.grads[k].cpu(), p.grad
        rel = _rel(got, want)
        cos = float(torch.nn.functional.cosine_similarity(got.flatten(), want.flatten(), dim=0))
        worst[k] = (rel, cos)
        assert rel < 5e-2 and cos > 0.995, (k, rel, cos)
    # one optimiser step moves the parameters like torch Adam with actor-only clipping
    before = {k: v.clone() for k, v in params.items()}
    model.apply(lr=3e-4, max_grad_norm=0.5)
    moved = sum(float((params[k] - before[k]).abs().max()) for k in params)
    assert moved > 0 and all(torch.isfinite(v).all() for v in params.values())
    k = "critic.enc3.psi"
    step = (params[k] - before[k]).cpu()
    # first Adam step: |delta| = lr * |g| / (|g| + eps*...) ~ lr * sign(g) where |g| >> eps
    gk = model.grads[k].cpu()
    big = gk.abs() > 1e-3
    if big.any():
        torch.testing.assert_close(step[big], -3e-4 * torch.sign(gk[big]), rtol=2e-2, atol=1e-6)
