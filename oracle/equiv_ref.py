"""CPU/torch restatement of the equivariant actor-critic update (row X) -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference's layers are `e2cnn.nn.R2Conv` over `gspaces.Rot2dOnR2(4)`
(src/nets/equiv.py:1,16-59,70-80,134-150); e2cnn==0.2.3 (setup.txt:25, environment.yml:489) is a
third-party dependency that is neither vendored nor installable here, and the reference has no
tests or golden vectors for this path.  What is restated:

  * the ARCHITECTURE exactly as the reference composes it: EquivariantEncoder128 (equiv.py:12-62:
    7 x [R2Conv 3x3, ReLU, (PointwiseMaxPool 2)], regular fields 16,32,64,128,256,128,128, padding
    1,1,1,1,1,0,0), EquivariantActor head (equiv.py:74-91: 1x1 conv to irrep(1) + 8 trivial, mean =
    [inv0, dx, dy, inv1, inv2], log_std clamped to [-20, 2]), EquivariantCritic head (equiv.py:138-150:
    1x1 regular->regular, ReLU, GroupPooling, 1x1 trivial->trivial), robot_actor_critic.evaluate
    (robot_actor_critic.py:104-131: gripper state tiled into a second image channel, Normal, summed
    log-prob / entropy, no tanh) and the loss of robot_ppo.update (robot_ppo.py:329-408);
  * the EQUIVARIANCE CONSTRAINT exactly: on a 3x3 (or 1x1) grid the 90-degree rotations map the grid
    onto itself, so the full solution space of C4-steerable kernels is the p4 group convolution
    (Cohen & Welling 2016): W[(o,r),(i,s),y,x] = psi[o,i,(s-r)%4, R_r^{-1}(y,x)].  e2cnn instead
    uses a band-limited basis of that same space (fewer free parameters per field pair), so WEIGHTS
    ARE NOT INTERCHANGEABLE with e2cnn checkpoints; outputs are equivariant under the same group
    action (checked in tests/test_equiv_oracle.py).

Channel order of a regular field type: channel = field * 4 + r  (r = rotation index), as e2cnn lays
out `n * [regular_repr]`.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

ENC_FIELDS = [16, 32, 64, 128, 256, 128, 128]      # n_out = 128: n/8, n/4, n/2, n, 2n, n, n
ENC_PAD = [1, 1, 1, 1, 1, 0, 0]
ENC_POOL = [True, True, True, True, False, True, False]
N_ACT = 5


def rot90_grid(w: torch.Tensor, r: int) -> torch.Tensor:
    """Spatial part of the group action on a filter: rotate the last two dims by r * 90 degrees (CCW)."""
    return torch.rot90(w, r, dims=(-2, -1))


def expand_trivial_to_regular(psi: torch.Tensor) -> torch.Tensor:
    """psi [Fo, Ci, k, k] -> W [Fo*4, Ci, k, k]: W[(o,r)] = rot_r(psi[o])."""
    return torch.stack([rot90_grid(psi, r) for r in range(4)], dim=1).reshape(psi.shape[0] * 4, *psi.shape[1:])


def expand_regular_to_regular(psi: torch.Tensor) -> torch.Tensor:
    """psi [Fo, Fi, 4, k, k] -> W [Fo*4, Fi*4, k, k]: W[(o,r),(i,s)] = rot_r(psi[o,i,(s-r)%4])."""
    Fo, Fi, _, k, _ = psi.shape
    rows = []
    for r in range(4):
        rolled = torch.roll(psi, shifts=r, dims=2)                # index s <- psi[(s - r) % 4]
        rows.append(rot90_grid(rolled, r))
    W = torch.stack(rows, dim=1)                                   # [Fo, 4(r), Fi, 4(s), k, k]
    return W.reshape(Fo * 4, Fi * 4, k, k)


def expand_regular_to_trivial(psi: torch.Tensor) -> torch.Tensor:
    """1x1: psi [Co, Fi] -> W [Co, Fi*4]: invariant output = the same weight on all 4 group channels."""
    return psi.unsqueeze(-1).expand(-1, -1, 4).reshape(psi.shape[0], -1)


def expand_regular_to_irrep1(psi: torch.Tensor) -> torch.Tensor:
    """1x1: psi [Fi, 2] -> W [2, Fi*4]: W[:, (i,s)] = R(s * 90deg) psi[i]  (standard representation)."""
    c = torch.tensor([1.0, 0.0, -1.0, 0.0]); s = torch.tensor([0.0, 1.0, 0.0, -1.0])
    a, b = psi[:, 0:1], psi[:, 1:2]                                # [Fi,1]
    row0 = (c * a - s * b).reshape(-1)                             # x component
    row1 = (s * a + c * b).reshape(-1)
    return torch.stack([row0, row1], 0)


def expand_bias_regular(b: torch.Tensor) -> torch.Tensor:
    return b.repeat_interleave(4)


def init_params(seed: int = 0, obs_channels: int = 2, scale: float = 1.0) -> Dict[str, torch.Tensor]:
    """He-style init of the free parameters psi of one actor + one critic (separate encoders)."""
    g = torch.Generator().manual_seed(seed)
    p: Dict[str, torch.Tensor] = {}
    for net in ("actor", "critic"):
        cin_f = None
        for l, fo in enumerate(ENC_FIELDS):
            if l == 0:
                fan_in = obs_channels * 9
                p[f"{net}.enc{l}.psi"] = torch.randn(fo, obs_channels, 3, 3, generator=g) * scale * math.sqrt(2.0 / fan_in)
            else:
                fan_in = cin_f * 4 * 9
                p[f"{net}.enc{l}.psi"] = torch.randn(fo, cin_f, 4, 3, 3, generator=g) * scale * math.sqrt(2.0 / fan_in)
            p[f"{net}.enc{l}.bias"] = 0.01 * torch.randn(fo, generator=g)
            cin_f = fo
    F_last = ENC_FIELDS[-1]
    p["actor.head.psi_irrep"] = torch.randn(F_last, 2, generator=g) * math.sqrt(1.0 / (F_last * 4))
    p["actor.head.psi_triv"] = torch.randn(2 * N_ACT - 2, F_last, generator=g) * math.sqrt(1.0 / (F_last * 4))
    p["actor.head.bias_triv"] = 0.01 * torch.randn(2 * N_ACT - 2, generator=g)
    p["critic.head1.psi"] = torch.randn(F_last, F_last, 4, 1, 1, generator=g) * math.sqrt(2.0 / (F_last * 4))
    p["critic.head1.bias"] = 0.01 * torch.randn(F_last, generator=g)
    p["critic.head2.w"] = torch.randn(1, F_last, generator=g) * math.sqrt(1.0 / F_last)
    p["critic.head2.bias"] = 0.01 * torch.randn(1, generator=g)
    return p


def bf16_ste(x: torch.Tensor) -> torch.Tensor:
    """Round to bf16 in the forward pass, identity in the backward pass (straight-through)."""
    return x + (x.bfloat16().float() - x).detach()


def _windows(z: torch.Tensor) -> torch.Tensor:
    """[B,C,H,W] -> [B,C,H/2,W/2,4] with the 2x2 window flattened as w = wy * 2 + wx."""
    B, C, H, W = z.shape
    return z.reshape(B, C, H // 2, 2, W // 2, 2).permute(0, 1, 2, 4, 3, 5).reshape(B, C, H // 2, W // 2, 4)


def encoder_forward(p: Dict[str, torch.Tensor], net: str, x: torch.Tensor, collect: List = None, quant: bool = False,
                    route: List = None, route_out: List = None) -> torch.Tensor:
    """x [B,2,128,128] -> [B,512] (regular fields at 1x1).

    quant=True mirrors WHERE the CUDA path stores bf16 (expanded weights of layers 1-6, every stored
    activation) so that the discrete routing decisions (max-pool / GroupPooling arg-max, ReLU masks)
    of the two paths coincide; the arithmetic stays fp32.

    route: per layer a dict(arg=[B,C,Hp,Wp] window index or None, pos=[B,C,Hp,Wp] bool): FORCES the discrete
    decisions (which window element a max-pool takes, which outputs the ReLU keeps) instead of taking them
    from this pass's own values -- with routing fixed the network is a linear map, so two implementations
    that agree on the routing must agree on every gradient to rounding.  route_out: filled with this
    pass's own routing in the same format (to count decisions that differ)."""
    for l in range(len(ENC_FIELDS)):
        psi = p[f"{net}.enc{l}.psi"]
        W = expand_trivial_to_regular(psi) if l == 0 else expand_regular_to_regular(psi)
        if quant and l > 0:
            W = bf16_ste(W)
        x = F.conv2d(x, W, expand_bias_regular(p[f"{net}.enc{l}.bias"]), padding=ENC_PAD[l])
        if route_out is not None:
            with torch.no_grad():
                if ENC_POOL[l]:
                    zw = _windows(x)
                    m, a = zw.max(-1)
                    # first maximum in scan order (torch's max_pool2d backward); .max returns an arbitrary one on exact ties
                    a = (zw == m.unsqueeze(-1)).float().argmax(-1)
                    route_out.append(dict(arg=a, pos=m > 0, margin=m))
                else:
                    route_out.append(dict(arg=None, pos=x > 0, margin=x.detach()))
        if route is not None:
            r = route[l]
            if ENC_POOL[l]:
                x = _windows(x).gather(-1, r["arg"].unsqueeze(-1)).squeeze(-1)
            x = x * r["pos"].to(x.dtype)
            if collect is not None:
                collect.append(x)
            continue
        x = F.relu(x)
        if ENC_POOL[l]:
            x = F.max_pool2d(x, 2)
        if quant:
            x = bf16_ste(x)
        if collect is not None:
            collect.append(x)
    return x.reshape(x.shape[0], -1)


def actor_forward(p, cat_obs, collect=None, quant=False, route=None, route_out=None):
    """EquivariantActor.forward (equiv.py:82-91) -> (mean [B,5], log_std [B,5])."""
    feat = encoder_forward(p, "actor", cat_obs, collect, quant, route=route["actor"] if route else None,
                           route_out=route_out["actor"] if route_out is not None else None)
    W = torch.cat([expand_regular_to_irrep1(p["actor.head.psi_irrep"]), expand_regular_to_trivial(p["actor.head.psi_triv"])], 0)
    if quant:
        W = bf16_ste(W)
    if route is not None and "keep" in route and W.requires_grad:  # expose the UN-projected head gradient to the tests
        W.retain_grad()
        route["keep"]["W_actor_head"] = W
    bias = torch.cat([torch.zeros(2, dtype=feat.dtype), p["actor.head.bias_triv"]])
    out = feat @ W.T + bias                                        # [B,10]
    dxy, inv_act = out[:, 0:2], out[:, 2:N_ACT]
    mean = torch.cat((inv_act[:, 0:1], dxy, inv_act[:, 1:]), dim=1)
    log_std = torch.clamp(out[:, N_ACT:], min=-20, max=2)
    return mean, log_std


def critic_forward(p, cat_obs, collect=None, quant=False, route=None, route_out=None):
    """EquivariantCritic.forward (equiv.py:153-157) -> value [B]."""
    feat = encoder_forward(p, "critic", cat_obs, collect, quant, route=route["critic"] if route else None,
                           route_out=route_out["critic"] if route_out is not None else None)
    W1 = expand_regular_to_regular(p["critic.head1.psi"]).reshape(feat.shape[1], feat.shape[1])
    if quant:
        W1 = bf16_ste(W1)
    hpre = feat @ W1.T + expand_bias_regular(p["critic.head1.bias"])
    if route_out is not None:
        with torch.no_grad():
            hw = hpre.reshape(hpre.shape[0], -1, 4)
            m = hw.max(-1).values
            route_out["group"] = dict(arg=(hw == m.unsqueeze(-1)).float().argmax(-1), pos=m > 0, margin=m)
    if route is not None:
        g = route["group"]
        pooled = hpre.reshape(hpre.shape[0], -1, 4).gather(-1, g["arg"].unsqueeze(-1)).squeeze(-1) * g["pos"].to(hpre.dtype)
        return (pooled @ p["critic.head2.w"].T + p["critic.head2.bias"]).reshape(-1)
    h = F.relu(hpre)
    pooled = h.reshape(h.shape[0], -1, 4).max(dim=2).values        # GroupPooling: max over the group channels
    return (pooled @ p["critic.head2.w"].T + p["critic.head2.bias"]).reshape(-1)


def cat_obs(state: torch.Tensor, obs: torch.Tensor) -> torch.Tensor:
    """robot_actor_critic.py:106-107: tile the gripper state into a second image channel."""
    tile = state.reshape(state.size(0), 1, 1, 1).repeat(1, 1, obs.shape[2], obs.shape[3])
    return torch.cat([obs, tile], dim=1)


def evaluate(p, state, obs, action, quant=False, route=None, route_out=None):
    """robot_actor_critic.evaluate with `action` given -> (log_prob [B], entropy [B], value [B])."""
    x = cat_obs(state, obs)
    mean, log_std = actor_forward(p, x, quant=quant, route=route, route_out=route_out)
    std = torch.exp(log_std)
    var = std ** 2
    log_prob = -((action - mean) ** 2) / (2 * var) - std.log() - math.log(math.sqrt(2 * math.pi))
    entropy = 0.5 + 0.5 * math.log(2 * math.pi) + std.log()
    return log_prob.sum(1), entropy.sum(1), critic_forward(p, x, quant=quant, route=route, route_out=route_out)


def update_loss(p, state, obs, action, oldlp, adv, ret, vold, true_action=None, clip_coeff=0.2, ent_c=0.01, vf_c=0.5,
                norm_adv=True, clip_vloss=True, expert_weight=0.0, quant=False, route=None, route_out=None):
    """Loss of robot_ppo.update (robot_ppo.py:345-398)."""
    newlogprob, entropy, newvalue = evaluate(p, state, obs, action, quant=quant, route=route, route_out=route_out)
    log_ratio = newlogprob - oldlp
    ratio = log_ratio.exp()
    mb_adv = adv
    if norm_adv:
        mb_adv = (mb_adv - mb_adv.mean()) / (mb_adv.std() + 1e-8)
    policy_loss = torch.max(-mb_adv * ratio, -mb_adv * torch.clamp(ratio, 1 - clip_coeff, 1 + clip_coeff)).mean()
    if clip_vloss:
        v_un = (newvalue - ret) ** 2
        v_cl = (vold + torch.clamp(newvalue - vold, -clip_coeff, clip_coeff) - ret) ** 2
        value_loss = 0.5 * torch.max(v_un, v_cl).mean()
    else:
        value_loss = 0.5 * ((newvalue - ret) ** 2).mean()
    value_loss = value_loss * vf_c
    entropy_loss = entropy.mean()
    loss = policy_loss - ent_c * entropy_loss + value_loss
    if true_action is not None and expert_weight:
        loss = loss + expert_weight * F.mse_loss(action, true_action)     # constant w.r.t. the parameters
    return loss, dict(policy_loss=policy_loss.item(), value_loss=value_loss.item(), entropy=entropy_loss.item(),
                      loss=loss.item())
