"""Row X at REFERENCE precision against float64 torch -- north_star's bar for this row: losses and gradients within 1e-4 relative.
Multi-plane modes of the tensor-core entry points (aur_tc_set_precision):
  3 planes ("fp32"):  bf16 hi + mid + lo operand planes, six products: fp32-equivalent, the reference-precision mode;
  2 planes ("split"): hi + mid, three products.
MEASURED (round 2): the tensor core adds each K = 16 step into its fp32 accumulator with TRUNCATION, so a long contraction
drifts by ~2e-8 per MMA step - 1.5e-5 (2 planes) / 3.1e-5 (3 planes, twice the steps) at K = 4608 when everything accumulates
in TMEM, MORE than the operand rounding of the two-plane split.  The multi-plane kernels therefore promote the accumulator
into fp32 registers (round-to-nearest) every 32 MMA steps (conv, GEMM) or bound the steps per CTA before the fp32 atomics
(weight gradients); the drift tests below hold that.  For scale: the reference runs these convolutions
through cuDNN with torch's default `torch.backends.cudnn.allow_tf32 = True` (torch 2.1.1, src/environment.yml:540; never
changed in src/), i.e. with 10-bit TF32 operands (~5e-4) on any Ampere-or-later GPU.

Per-layer tests compare each kernel with F.conv2d / autograd evaluated in float64 on the SAME fp32 inputs.  The whole-update
test compares every parameter gradient with float64 autograd of the restated model

  (a) on IDENTICAL ROUTING (the oracle is forced to take the max-pool arg-max, ReLU masks and GroupPooling choices the device
      took: with routing fixed the network is linear, so this isolates arithmetic) at <= 1e-4 per tensor -- asserted;
  (b) with the oracle's OWN routing: the number of discrete decisions that differ is counted and printed, and the gradients are
      held to a looser bar, because ONE flipped ReLU / pool decision moves a gradient of cancelling terms by ~sqrt(1 / #active)
      -- the same happens between torch fp32 and torch fp64 on the CPU (measured in the test and printed beside it), so no
      implementation can meet 1e-4 there on random inputs; that is a property of max-pool / ReLU, not of the kernels.
"""
import math

import pytest
import torch
import torch.nn.functional as F

from aur_ppo_b200 import _lib, kernels
from oracle import equiv_ref as Q

pytestmark = pytest.mark.gpu
BAR = 1e-4
LAYER_BAR = {2: 1e-5, 3: 5e-6}           # per-kernel relative L2 vs float64, by operand planes (measured 4.5e-6 / 7e-7; wgrad 3e-6)


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-300))


def test_split_planes_round_trip():
    x = torch.randn(1000, device="cuda") * torch.logspace(-6, 6, 1000, device="cuda")
    pl = kernels.split_planes(x)
    assert pl.shape == (2, 1000) and pl.dtype == torch.bfloat16
    back = kernels.join_planes(pl)
    assert float(((back - x).abs() / x.abs()).max()) < 2.0 ** -16
    assert torch.equal(kernels.join_planes(kernels.split_planes(x, 3)), x)      # three bf16 planes hold every fp32 bit


@pytest.mark.parametrize("P", [2, 3])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 200, 512), (64, 16, 4608), (16, 512, 8)])
def test_tc_gemm_split_matches_fp64(M, N, K, P):
    g = torch.Generator().manual_seed(M + N + K)
    a, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g)
    with kernels.tc_precision(P):
        c = kernels.tc_gemm_bf16(kernels.split_planes(a.cuda(), P), kernels.split_planes(b.cuda(), P))
    want = a.double() @ b.double().T
    print(f"gemm {M}x{N}x{K} planes {P}: rel {_rel(c.cpu(), want):.2e}")
    assert _rel(c.cpu(), want) < (6e-6 if P == 2 else 2e-6), _rel(c.cpu(), want)
    # the single-plane mode on the same operands is ~300x less accurate: the split is what buys the precision
    with kernels.tc_precision(1):
        c1 = kernels.tc_gemm_bf16(a.cuda().bfloat16(), b.cuda().bfloat16())
    assert _rel(c1.cpu(), want) > 20 * _rel(c.cpu(), want)
    # wrong plane count for the current mode is an error, not a silent misread
    with pytest.raises(_lib.AurError):
        kernels.tc_gemm_bf16(kernels.split_planes(a.cuda(), P), kernels.split_planes(b.cuda(), P))


@pytest.mark.parametrize("B,H,Fi,Fo,pad,pool", [(3, 16, 16, 32, 1, True), (2, 64, 16, 32, 1, True), (5, 8, 32, 64, 1, False),
                                                (4, 8, 64, 32, 0, True), (2, 32, 32, 16, 1, False), (5, 8, 16, 32, 1, True),
                                                (1, 32, 32, 48, 1, True), (2, 8, 128, 128, 0, False),
                                                (2, 8, 256, 128, 0, False)])          # Cin = 1024: K = 9216 (layer 5)
@pytest.mark.parametrize("P", [2, 3])
def test_conv_layer_split_matches_fp64_conv2d(B, H, Fi, Fo, pad, pool, P):
    g = torch.Generator().manual_seed(B * 100 + H)
    Cin, Cout = Fi * 4, Fo * 4
    psi = torch.randn(Fo, Fi, 4, 3, 3, generator=g) * (2.0 / (Cin * 9)) ** 0.5
    bias = 0.1 * torch.randn(Fo, generator=g)
    x = torch.randn(B, Cin, H, H, generator=g)
    Hb = H + 2 * pad
    Ho = Hb - 2
    Hn = Ho // 2 if pool else Ho
    with kernels.tc_precision(P):
        wmat, _, bias_ch = kernels.equiv_expand_regular(psi.cuda(), bias.cuda())
        assert wmat.shape == (P, Cout, 9, Cin)
        inp = torch.zeros(P, B, Hb, Hb, Cin, dtype=torch.bfloat16, device="cuda")
        inp[:, :, pad:pad + H, pad:pad + H, :] = kernels.split_planes(x.permute(0, 2, 3, 1).contiguous().cuda(), P)
        out = torch.zeros(P, B, Hn + 2, Hn + 2, Cout, dtype=torch.bfloat16, device="cuda")
        arg = torch.zeros(B, Hn, Hn, Cout, dtype=torch.uint8, device="cuda") if pool else None
        kernels.conv3x3_bf16(inp, wmat, bias_ch, 2 if pool else 1, out, 1, arg)
    W = Q.expand_regular_to_regular(psi.double())
    ref = F.relu(F.conv2d(x.double(), W, Q.expand_bias_regular(bias.double()), padding=pad))
    if pool:
        ref = F.max_pool2d(ref, 2)
    got = kernels.join_planes(out)[:, 1:1 + Hn, 1:1 + Hn, :].permute(0, 3, 1, 2).cpu()
    print(f"conv Cin {Cin} Cout {Cout} H {H} planes {P}: rel {_rel(got, ref):.2e}")
    assert _rel(got, ref) < LAYER_BAR[P], _rel(got, ref)
    assert float(out[:, :, 0].abs().max()) == 0 and float(out[:, :, :, -1].abs().max()) == 0      # halos untouched, both planes
    if pool:      # routing: the stored arg-max picks the window element whose value IS the pooled value
        full = F.relu(F.conv2d(x.double(), W, Q.expand_bias_regular(bias.double()), padding=pad))
        picked = Q._windows(full).gather(-1, arg.permute(0, 3, 1, 2).long().cpu().unsqueeze(-1)).squeeze(-1)
        assert _rel(picked, ref) < LAYER_BAR[P]


@pytest.mark.parametrize("P", [2, 3])
@pytest.mark.parametrize("B,H,Cin,Cout", [(3, 16, 64, 128), (2, 32, 64, 64), (4, 8, 128, 256), (5, 8, 64, 200)])
def test_wgrad3x3_split_matches_fp64_autograd(B, H, Cin, Cout, P):
    from aur_ppo_b200.kernels import _stream
    g = torch.Generator().manual_seed(Cin + Cout)
    x = torch.randn(B, Cin, H, H, generator=g)
    dy = torch.randn(B, Cout, H, H, generator=g) * 0.1
    Hb = H + 2
    xb = torch.zeros(P, B, Hb, Hb, Cin, dtype=torch.bfloat16, device="cuda")
    xb[:, :, 1:1 + H, 1:1 + H, :] = kernels.split_planes(x.permute(0, 2, 3, 1).contiguous().cuda(), P)
    dyb = torch.zeros(P, B, Hb, Hb, Cout, dtype=torch.bfloat16, device="cuda")
    dyb[:, :, 1:1 + H, 1:1 + H, :] = kernels.split_planes(dy.permute(0, 2, 3, 1).contiguous().cuda(), P)
    dw = torch.zeros(Cout, 9, Cin, device="cuda")
    with kernels.tc_precision(P):
        rc = _lib.lib().aur_wgrad3x3_bf16(Cout, Cin, B * Hb * Hb, dyb.data_ptr(), xb.data_ptr(), -(Hb + 1), Hb, dw.data_ptr(), 0, _stream())
    _lib.check(rc, "aur_wgrad3x3_bf16")
    W = torch.zeros(Cout, Cin, 3, 3, dtype=torch.float64, requires_grad=True)
    F.conv2d(x.double(), W, padding=1).backward(dy.double())
    want = W.grad.permute(0, 2, 3, 1).reshape(Cout, 9, Cin)
    print(f"wgrad Cin {Cin} Cout {Cout} planes {P}: rel {_rel(dw.cpu(), want):.2e}")
    assert _rel(dw.cpu(), want) < LAYER_BAR[P], _rel(dw.cpu(), want)


def test_weight_gradient_accumulator_drift_at_minibatch_scale():
    """1.1 M pixel rows (what a 256-sample minibatch gives layer 1; config D has 16x that) with same-signed terms, the worst
    case for a truncating accumulator: every CTA's TMEM partial must stay short enough that the fp32 result is still within
    2e-5 of a float64 reference (here torch's float64 convolution backward on the device)."""
    from aur_ppo_b200.kernels import _stream
    B, H, Cin, Cout = 256, 64, 64, 64
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.rand(B, Cin, H, H, generator=g, device="cuda") + 0.5
    dy = torch.rand(B, Cout, H, H, generator=g, device="cuda") * 0.1 + 0.05
    Hb = H + 2
    xb = torch.zeros(2, B, Hb, Hb, Cin, dtype=torch.bfloat16, device="cuda")
    xb[:, :, 1:1 + H, 1:1 + H, :] = kernels.split_planes(x.permute(0, 2, 3, 1).contiguous())
    dyb = torch.zeros(2, B, Hb, Hb, Cout, dtype=torch.bfloat16, device="cuda")
    dyb[:, :, 1:1 + H, 1:1 + H, :] = kernels.split_planes(dy.permute(0, 2, 3, 1).contiguous())
    dw = torch.zeros(Cout, 9, Cin, device="cuda")
    with kernels.tc_precision(2):
        rc = _lib.lib().aur_wgrad3x3_bf16(Cout, Cin, B * Hb * Hb, dyb.data_ptr(), xb.data_ptr(), -(Hb + 1), Hb, dw.data_ptr(), 0, _stream())
    _lib.check(rc, "aur_wgrad3x3_bf16")
    want = torch.nn.grad.conv2d_weight(x.double(), (Cout, Cin, 3, 3), dy.double(), padding=1).permute(0, 2, 3, 1).reshape(Cout, 9, Cin)
    err = _rel(dw.cpu(), want.cpu())
    print(f"weight-gradient drift at Q = {B * Hb * Hb}: rel {err:.2e}")
    assert err < 2e-5, err


@pytest.mark.parametrize("P", [2, 3])
def test_conv0_split_forward_and_weight_gradient(P):
    g = torch.Generator().manual_seed(11)
    B = 5
    psi = (torch.randn(16, 2, 3, 3, generator=g) * 0.3)
    bias = (0.1 * torch.randn(16, generator=g))
    obs = torch.rand(B, 1, 128, 128, generator=g) * 0.32
    state = (torch.rand(B, generator=g) > 0.5).float()
    da1 = torch.randn(B, 64, 64, 64, generator=g) * 0.1                                  # upstream gradient, NHWC
    out = torch.zeros(P, B, 66, 66, 64, dtype=torch.bfloat16, device="cuda")
    arg = torch.zeros(B, 64, 64, 64, dtype=torch.uint8, device="cuda")
    ws = torch.zeros(64 * 18 + 64, device="cuda")
    dpsi, dbias = torch.zeros(16, 2, 3, 3, device="cuda"), torch.zeros(16, device="cuda")
    # (device tensors are kept in variables: a temporary's memory may be reused before the kernel has read it)
    d_obs, d_state, d_g = obs.cuda(), state.cuda(), kernels.split_planes(da1.cuda(), P)
    with kernels.tc_precision(P):
        kernels.equiv_conv0(d_obs, d_state, psi.cuda(), bias.cuda(), out, arg)
        rc = _lib.lib().aur_equiv_conv0_wgrad(d_obs.data_ptr(), d_state.data_ptr(), d_g.data_ptr(), out.data_ptr(), arg.data_ptr(), B,
                                              ws.data_ptr(), dpsi.data_ptr(), dbias.data_ptr(), kernels._stream())
    _lib.check(rc, "aur_equiv_conv0_wgrad")
    torch.cuda.synchronize()
    x = Q.cat_obs(state, obs).double()
    pd, bd = psi.double().requires_grad_(True), bias.double().requires_grad_(True)
    z = F.conv2d(x, Q.expand_trivial_to_regular(pd), Q.expand_bias_regular(bd), padding=1)
    ref = F.max_pool2d(F.relu(z), 2)
    got = kernels.join_planes(out)[:, 1:65, 1:65, :].permute(0, 3, 1, 2).cpu()
    assert _rel(got, ref) < 1e-5, _rel(got, ref)
    # gradient on the DEVICE's routing (its arg-max and its positive outputs)
    a = arg.permute(0, 3, 1, 2).long().cpu()
    pos = (out[0, :, 1:65, 1:65, :].permute(0, 3, 1, 2).float().cpu() > 0)
    y = Q._windows(z).gather(-1, a.unsqueeze(-1)).squeeze(-1) * pos.double()
    y.backward(da1.double().permute(0, 3, 1, 2))
    assert _rel(dpsi.cpu(), pd.grad) < LAYER_BAR[P] * 3, _rel(dpsi.cpu(), pd.grad)
    assert _rel(dbias.cpu(), bd.grad) < LAYER_BAR[P] * 3, _rel(dbias.cpu(), bd.grad)


def _device_route(model, B, real=None):
    """The discrete decisions the device took in its last forward, in the oracles' `route` format.
    real: real channel count per layer when the stored activations are channel-padded (plain CNN)."""
    route = {}
    inner = [(1, 65), (1, 33), (1, 17), (1, 9), (0, 8), (0, 3)]
    for net in ("actor", "critic"):
        e, layers = model.enc[net], []
        for l in range(6):
            lo, hi = inner[l]
            C = real[l] if real else e.a[l].shape[-1]
            act = e.a[l][0, :, lo:hi, lo:hi, :C].permute(0, 3, 1, 2).float().cpu()          # hi plane: > 0 <=> value > 0
            arg = e.arg[l][..., :C].permute(0, 3, 1, 2).long().cpu() if e.arg[l] is not None else None
            layers.append(dict(arg=arg, pos=act > 0))
        layers.append(dict(arg=None, pos=(e.feat[0].float().cpu() > 0).reshape(B, -1, 1, 1)))
        route[net] = layers
    return route


def _count_flips(dev_route, own_route):
    flips, total = 0, 0
    items = [(d, o) for net in ("actor", "critic") for d, o in zip(dev_route[net], own_route[net])]
    items.append((dev_route["group"], own_route["group"]))
    for d, o in items:
        diff = d["pos"] != o["pos"]
        if d["arg"] is not None:
            diff = diff | (d["pos"] & o["pos"] & (d["arg"] != o["arg"]))
        flips += int(diff.sum())
        total += diff.numel()
    return flips, total


@pytest.mark.parametrize("kind,head_scale,precision", [("equiv", 0.02, "fp32"), ("equiv", 0.1, "fp32"), ("plain", 1.0, "fp32"),
                                                       ("equiv", 0.02, "split"), ("equiv", 0.1, "split"), ("plain", 1.0, "split")])
def test_full_update_split_gradients_match_fp64_autograd(kind, head_scale, precision):
    """Bars on identical routing (measured values in brackets).  "fp32": EVERY tensor, the forward log-prob / value and the
    three loss terms within north_star's 1e-4 [<= 2.6e-5].  "split": the same 1e-4 for the critic chain [<= 2e-5] and for the
    actor at both head scales [<= 7e-5], 3e-4 for the plain CNN's actor [1.7e-4: its forward log-prob error of 1.3e-4
    ABSOLUTE is the relative error of every loss seed].  actor.head.psi_irrep is measured against the norm of the UN-projected
    head gradient it is the C4 projection of (the projection cancels ~30x on this input).
    head_scale = factor on the equivariant head filters.  The actor's log_std is a head OUTPUT there (equiv.py:88-90); at 0.1
    it reaches -2 (std 0.13) on this input, where d log_prob / d mean = diff / var amplifies a forward error ~50x into the
    loss seeds of EVERY actor gradient; at 0.02 (|log_std| < 0.5, a freshly initialised policy) it does not.
    actor.head.psi_irrep is the C4 projection of a mostly non-equivariant head gradient: a difference of nearly equal sums."""
    from aur_ppo_b200 import equiv, plain_cnn
    B = 8
    g = torch.Generator().manual_seed(1)
    obs = torch.rand(B, 1, 128, 128, generator=g) * 0.32
    for b in range(B):                                       # blocks of different height, like close_loop_block_picking heightmaps
        y, x = 20 + 9 * b, 90 - 8 * b
        obs[b, 0, y:y + 16, x:x + 16] += 0.1 + 0.02 * b
    state = (torch.rand(B, generator=g) > 0.5).float()
    action = torch.randn(B, 5, generator=g)
    adv, ret = torch.randn(B, generator=g), torch.randn(B, generator=g)
    real = None
    if kind == "equiv":
        O = Q
        params = equiv.init_params(seed=5, scale=1.1)
        for k in ("actor.head.psi_triv", "actor.head.psi_irrep", "critic.head2.w"):
            params[k].mul_(head_scale)
        make = lambda: equiv.EquivActorCritic(params, B, precision=precision)
    else:
        from oracle import cnn_ref as O
        params = {k: v.cuda().contiguous() for k, v in O.formula_params(O.param_shapes(), seed=3).items()}
        make = lambda: plain_cnn.PlainActorCritic(params, B, precision=precision)
        real = plain_cnn.REAL
    p32 = {k: v.detach().cpu().clone() for k, v in params.items()}
    p64 = {k: v.double().requires_grad_(True) for k, v in p32.items()}
    with torch.no_grad():
        lp0, _, v0 = O.evaluate(p32, state, obs, action)
    oldlp = lp0 + 0.15 * torch.randn(B, generator=g)
    vold = v0 + 0.3 * torch.randn(B, generator=g)
    model = make()
    dev = lambda t: t.cuda().contiguous()
    st = model.loss_and_grads(dev(state), dev(obs), dev(action), dev(oldlp), dev(adv), dev(ret), dev(vold)).cpu()
    args64 = [t.double() for t in (state, obs, action, oldlp, adv, ret, vold)]

    # ---- (a) identical routing: the device's decisions forced on the float64 oracle
    route = _device_route(model, B, real)
    route["keep"] = {}
    a_out, c_pre, _, _ = model._last_head
    if kind == "equiv":
        hw = (c_pre.cpu() + Q.expand_bias_regular(p32["critic.head1.bias"])).reshape(B, -1, 4)
        m = hw.max(-1).values
        route["group"] = dict(arg=(hw == m.unsqueeze(-1)).float().argmax(-1), pos=m > 0)
    else:
        route["group"] = dict(arg=None, pos=(c_pre.cpu() + p32["critic.critic.0.bias"]) > 0)
    loss, st_ref = O.update_loss(p64, *args64, route=route)
    loss.backward()
    with torch.no_grad():
        lp64, _, v64 = O.evaluate(p64, *args64[:3], route=route)
    d_lp = float((model.logp.cpu().double() - lp64).abs().max())
    d_v = float((model.value.cpu().double() - v64).abs().max())
    forced = {k: _rel(model.grads[k].cpu(), p64[k].grad) for k in p64}
    if kind == "equiv":                      # the irrep head filter: error against the un-projected gradient's norm (see docstring)
        k = "actor.head.psi_irrep"
        unproj = route["keep"]["W_actor_head"].grad[0:2]
        cancel = float(unproj.norm() / p64[k].grad.norm())
        forced[k] = float((model.grads[k].cpu().double() - p64[k].grad).norm() / unproj.norm())
        print(f"   actor.head.psi_irrep: projection cancels {cancel:.0f}x; error / |un-projected gradient| = {forced[k]:.1e}")
    worst_a = max(v for k, v in forced.items() if k.startswith("actor"))
    worst_c = max(v for k, v in forced.items() if k.startswith("critic"))
    print(f"[{kind} x{head_scale} {precision}] device vs float64 autograd on identical routing: worst actor {worst_a:.1e}, worst critic "
          f"{worst_c:.1e}; forward: max |log_prob error| {d_lp:.1e} (|log_prob| ~ {float(lp64.abs().mean()):.1f}), max |value error| {d_v:.1e}")
    print("   per tensor:", {k: f"{v:.1e}" for k, v in forced.items()})
    assert abs(st[0] - st_ref["policy_loss"]) < BAR * max(1, abs(st_ref["policy_loss"]))
    assert abs(st[1] - st_ref["value_loss"]) < BAR * max(1, abs(st_ref["value_loss"]))
    assert abs(st[2] - st_ref["entropy"]) < BAR * max(1, abs(st_ref["entropy"]))
    assert d_lp < BAR * float(lp64.abs().mean()) and d_v < BAR * max(1.0, float(v64.abs().max()))
    assert worst_c < BAR, forced
    assert worst_a < (3e-4 if (kind == "plain" and precision == "split") else BAR), forced

    # ---- (b) the oracle's own routing in float64 and in float32: flips and what they cost
    own64 = {"actor": [], "critic": []}
    q64 = {k: v.detach().clone().requires_grad_(True) for k, v in p64.items()}
    loss64, _ = O.update_loss(q64, *args64, route_out=own64)
    loss64.backward()
    own32 = {"actor": [], "critic": []}
    q32 = {k: v.clone().requires_grad_(True) for k, v in p32.items()}
    loss32, _ = O.update_loss(q32, state, obs, action, oldlp, adv, ret, vold, route_out=own32)
    loss32.backward()
    flips_dev, total = _count_flips(route, own64)
    flips_f32, _ = _count_flips(own32, own64)
    free_dev = {k: _rel(model.grads[k].cpu(), q64[k].grad) for k in q64}
    free_f32 = {k: _rel(q32[k].grad, q64[k].grad) for k in q64}
    print(f"   routing decisions differing from float64: device ({precision}) {flips_dev} of {total}, torch fp32 on the CPU {flips_f32} of {total}")
    print("   own-routing gradient error vs float64: device worst %.2e (median %.2e); torch fp32 CPU worst %.2e (median %.2e)" %
          (max(free_dev.values()), sorted(free_dev.values())[len(free_dev) // 2], max(free_f32.values()),
           sorted(free_f32.values())[len(free_f32) // 2]))
    assert flips_dev <= (max(20, 4 * max(flips_f32, 1)) if precision == "fp32" else max(400, 100 * max(flips_f32, 1))), (flips_dev, flips_f32)
    assert max(free_dev.values()) < 5e-2, free_dev

    if kind != "equiv" or precision != "fp32":
        return
    # one Adam step on these gradients (actor-only clip, robot_ppo.py:401-402)
    before = {k: v.clone() for k, v in params.items()}
    model.apply(lr=3e-4, max_grad_norm=0.5)
    norm = math.sqrt(sum(float((model.grads[q].double() ** 2).sum()) for q in model.grads if q.startswith("actor.")))
    coef = min(1.0, 0.5 / (norm + 1e-6))
    for k in ("actor.enc3.psi", "critic.enc3.psi"):
        gk = model.grads[k].cpu() * (coef if k.startswith("actor.") else 1.0)
        step = (params[k] - before[k]).cpu()
        want = -3e-4 / (1 - 0.9) * (0.1 * gk) / ((0.001 * gk * gk).sqrt() / math.sqrt(1 - 0.999) + 1e-5)
        torch.testing.assert_close(step, want, rtol=1e-4, atol=2e-8)      # (params - before) is quantised at ulp(param) ~ 4e-9


def test_config_d_full_minibatch_properties():
    """BASELINE config D at its full size (minibatch 4096, two-plane mode to bound memory) through size-independent properties:
    (1) a sample's log-prob / value do not depend on what else is in the batch: samples 0..7 of the 4096 match the float64
    oracle evaluated on those 8 alone; (2) gradient sums are additive over a split of the minibatch (checksum of checksums):
    grads(4096) == grads(first 2048) + grads(last 2048) with the same 1 / 4096 seeds (advantage normalisation off so that the
    halves share nothing but the weights) - this also exercises the weight-gradient accumulation at 17.8 M pixel rows."""
    from aur_ppo_b200 import equiv
    B = 4096
    params = equiv.init_params(seed=5, scale=1.1)
    for k in ("actor.head.psi_triv", "actor.head.psi_irrep", "critic.head2.w"):
        params[k].mul_(0.02)
    g = torch.Generator(device="cuda").manual_seed(0)
    obs = torch.rand(B, 1, 128, 128, generator=g, device="cuda") * 0.32
    state = (torch.rand(B, generator=g, device="cuda") > 0.5).float()
    action = torch.randn(B, 5, generator=g, device="cuda")
    adv, ret, vold = (torch.randn(B, generator=g, device="cuda") for _ in range(3))
    oldlp = -6.4 + 0.2 * torch.randn(B, generator=g, device="cuda")        # around the new log-probs: ratios on both sides of the clip
    p32 = {k: v.detach().cpu().clone() for k, v in params.items()}
    kw = dict(norm_adv=False)
    whole = equiv.EquivActorCritic(params, B, precision="split")
    whole.loss_and_grads(state, obs, action, oldlp, adv, ret, vold, **kw)
    lp_dev, v_dev = whole.logp[:8].cpu().double(), whole.value[:8].cpu().double()
    g_whole = {k: v.clone() for k, v in whole.grads.items()}
    del whole
    torch.cuda.empty_cache()
    p64 = {k: v.double() for k, v in p32.items()}
    with torch.no_grad():
        lp64, _, v64 = Q.evaluate(p64, state[:8].cpu().double(), obs[:8].cpu().double(), action[:8].cpu().double())
    assert float((lp_dev - lp64).abs().max()) < BAR * float(lp64.abs().mean()), (lp_dev, lp64)
    assert float((v_dev - v64).abs().max()) < BAR * max(1.0, float(v64.abs().max()))
    half = equiv.EquivActorCritic(params, B // 2, precision="split")
    acc = {k: torch.zeros_like(v) for k, v in g_whole.items()}
    for sl in (slice(0, B // 2), slice(B // 2, B)):
        half.loss_and_grads(*(t[sl].contiguous() for t in (state, obs, action, oldlp, adv, ret, vold)), m_total=B, **kw)
        for k in acc:
            acc[k] += half.grads[k]
    worst = {k: _rel(acc[k].cpu(), g_whole[k].cpu()) for k in acc}
    print("config D additivity over halves, worst relative L2:", max(worst.values()))
    assert max(worst.values()) < BAR, worst
