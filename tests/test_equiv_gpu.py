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
