"""Plain CNN actor-critic update on the GPU vs oracle/cnn_ref.py, which is pinned to the reference's
own base_actor / base_critic classes by tests/golden/plain_cnn.npz.  THIS FILE: the single-plane bf16 FAST mode (below the
reference's fp32 precision), compared with the checker rounded at the same storage points (quant=True), the bars of
tests/test_equiv_gpu.py.  The reference-precision (split) mode is held to 1e-4 in tests/test_equiv_split_gpu.py."""
import math

import numpy as np
import pytest
import torch

from aur_ppo_b200 import plain_cnn
from oracle import cnn_ref as C

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-12))


def _setup(B=8, seed=1):
    g = torch.Generator().manual_seed(seed)
    obs = torch.rand(B, 1, 128, 128, generator=g) * 0.32
    # blocks of different height, like the heightmaps of close_loop_block_picking
    for b in range(B):
        y, x = 20 + 9 * b, 90 - 8 * b
        obs[b, 0, y:y + 16, x:x + 16] += 0.1 + 0.02 * b
    state = (torch.rand(B, generator=g) > 0.5).float()
    action = torch.randn(B, 5, generator=g)
    adv, ret = torch.randn(B, generator=g), torch.randn(B, generator=g)
    return obs, state, action, adv, ret, g


def test_plain_update_gradients_match_oracle_autograd():
    B = 8
    obs, state, action, adv, ret, g = _setup(B)
    cpu = {k: v.clone().requires_grad_(True) for k, v in C.formula_params(C.param_shapes(), seed=3).items()}
    with torch.no_grad():
        lp0, _, v0 = C.evaluate(cpu, state, obs, action)
    oldlp = lp0 + 0.15 * torch.randn(B, generator=g)
    vold = v0 + 0.3 * torch.randn(B, generator=g)
    loss, st_ref = C.update_loss(cpu, state, obs, action, oldlp, adv, ret, vold, quant=True)
    loss.backward()
    with torch.no_grad():
        lp_q, ent_q, v_q = C.evaluate(cpu, state, obs, action, quant=True)
        lp_f, _, v_f = C.evaluate(cpu, state, obs, action, quant=False)

    params = {k: v.detach().clone().cuda().contiguous() for k, v in cpu.items()}
    model = plain_cnn.PlainActorCritic(params, B)
    dev = lambda t: t.cuda().contiguous()
    st = model.loss_and_grads(dev(state), dev(obs), dev(action), dev(oldlp), dev(adv), dev(ret), dev(vold)).cpu()
    torch.testing.assert_close(model.value.cpu(), v_q, rtol=1e-2, atol=3e-3)
    torch.testing.assert_close(model.logp.cpu(), lp_q, rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(model.value.cpu(), v_f, rtol=3e-2, atol=2e-2)          # vs plain fp32
    torch.testing.assert_close(model.logp.cpu(), lp_f, rtol=3e-2, atol=5e-2)
    assert abs(st[0] - st_ref["policy_loss"]) < 1e-2 * max(1, abs(st_ref["policy_loss"]))
    assert abs(st[1] - st_ref["value_loss"]) < 1e-2 * max(1, abs(st_ref["value_loss"]))
    assert abs(st[2] - st_ref["entropy"]) < 1e-3 * max(1, abs(st_ref["entropy"]))
    worst = {}
    for k in cpu:
        got, want = model.grads[k].cpu(), cpu[k].grad
        assert got.shape == want.shape, k
        rel = _rel(got, want)
        cos = float(torch.nn.functional.cosine_similarity(got.flatten(), want.flatten(), dim=0))
        worst[k] = (rel, cos)
    # Bars.  Max-pool / ReLU routing is discrete: on this input the checker's own gradients move by 8-41 % (actor encoder)
    # and 8-20 % (critic encoder) in relative L2 between its fp32 and bf16-storage variants, because a few near-tie routes
    # flip (8 samples, 16-64 channels: little averaging; the actor's five signed head gradients cancel more than the
    # critic's single one).  Against the bf16-storage checker the CUDA path must stay well inside that: every tensor of the
    # critic chain, the heads and the actor's last convolution within 7e-2 / cosine 0.998 (measured <= 0.050); the actor's encoder layers 0-5,
    # where a flipped route is amplified layer by layer, within 0.2 / 0.985 (measured 0.06-0.18, less than half the
    # checker-vs-checker gap).  The same kernels are compared per layer against conv2d / autograd in test_equiv_gpu.py.
    loose = {k for k in cpu if k.startswith("actor.conv.conv.") and not k.startswith("actor.conv.conv.17.")}
    bad = {k: v for k, v in worst.items()
           if not ((v[0] < 0.2 and v[1] > 0.985) if k in loose else (v[0] < 7e-2 and v[1] > 0.998))}
    assert not bad, bad
    # one Adam step with the actor-only clip (robot_ppo.py:401-402) against torch.optim.Adam on the checker's gradients
    actor_keys = [k for k in cpu if k.startswith("actor.")]
    ref_p = {k: v.detach().clone() for k, v in cpu.items()}
    gn = math.sqrt(sum(float(model.grads[k].double().pow(2).sum()) for k in actor_keys))
    coef = min(1.0, 0.5 / (gn + 1e-6))
    model.apply(lr=3e-4, max_grad_norm=0.5)
    for k in cpu:
        gk = model.grads[k].cpu() * (coef if k in actor_keys else 1.0)
        m = 0.1 * gk
        v = 0.001 * gk * gk
        step = 3e-4 / (1 - 0.9) * m / (v.sqrt() / math.sqrt(1 - 0.999) + 1e-5)
        np.testing.assert_allclose(model.p[k].cpu().numpy(), (ref_p[k] - step).numpy(), rtol=1e-5, atol=2e-7, err_msg=k)


def test_plain_facade_evaluate_and_sampling():
    from aur_ppo_b200.models import robot_actor_critic
    B = 5
    obs, state, action, adv, ret, g = _setup(B, seed=4)
    m = robot_actor_critic("cuda", False, seed=2)
    cpu = {k: v.detach().cpu().clone() for k, v in m.tensors().items()}
    with torch.no_grad():
        lp_ref, ent_ref, v_ref = C.evaluate(cpu, state, obs, action, quant=True)
    scaled, unscaled, lp, ent, val = m.evaluate(state, obs, action.cuda())
    torch.testing.assert_close(lp.cpu(), lp_ref, rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(ent.cpu(), ent_ref, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(val.cpu().reshape(-1), v_ref, rtol=1e-2, atol=3e-3)
    assert torch.equal(m.value(state, obs), val)
    un_t, sc_t = m.decodeActions(*[action.cuda()[:, i] for i in range(5)])
    assert torch.equal(sc_t, scaled) and torch.equal(un_t, unscaled)
    s1, u1, lp1, _, _ = m.evaluate(state, obs)
    s3, u3, lp3, _, _ = m.evaluate(state, obs, u1)
    assert torch.equal(u3, u1) and torch.equal(s3, s1) and torch.equal(lp3, lp1)
    # reference checkpoints load by name: base_actor / base_critic state_dict keys under actor. / critic.
    sd = m.reference_state_dicts()
    assert "conv.conv.0.weight" in sd["actor_state"] and "critic.2.bias" in sd["critic_state"] and sd["actor_logstd"].shape == (1, 5)
    m.load_reference_state_dicts(sd)
    assert m.engine(8).p["actor.conv.conv.9.weight"].data_ptr() == m.tensors()["actor.conv.conv.9.weight"].data_ptr()
