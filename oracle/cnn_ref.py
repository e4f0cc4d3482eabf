"""CPU restatement of the reference's plain CNN actor-critic -- TEST INFRASTRUCTURE ONLY (tests/, never the product path).

Follows src/nets/base_cnns.py:20-84 (`base_encoder`: seven 3x3 Conv2d + ReLU, MaxPool2d(2) after layers 0-3 and 5,
padding 1,1,1,1,1,0,0; `base_actor.mean_linear`; `base_critic.critic`), src/models/robot_actor_critic.py:41-51,104-131
(`evaluate`: state tiled into a second image channel, Normal(mean, exp(actor_logstd)), summed log-prob / entropy) and
the loss of src/robot_ppo.py:345-398.  Parameters are keyed by the reference modules' state_dict names under
`actor.` / `critic.` plus `actor_logstd`.

PINNED: tests/golden/plain_cnn.npz holds outputs, loss terms and per-tensor gradient norms produced by the reference's
OWN `base_actor` / `base_critic` classes (oracle/gen_golden_cnn.py imports them from /root/reference);
tests/test_plain_cnn_oracle.py checks this file against it.

quant=True rounds to bf16 (straight-through) where the CUDA path stores bf16: the weights of layers 1-6 and of the two
head matrices, and every stored activation - so that max-pool / ReLU routing coincides; arithmetic stays fp32."""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

CONV_IDX = [0, 3, 6, 9, 12, 14, 17]
PADS = [1, 1, 1, 1, 1, 0, 0]
POOL = [True, True, True, True, False, True, False]
N_ACT = 5


def bf16_ste(x: torch.Tensor) -> torch.Tensor:
    return x + (x.bfloat16().float() - x).detach()


def formula_params(shapes: Dict[str, tuple], seed: int = 11) -> Dict[str, torch.Tensor]:
    """Deterministic parameters for the golden file (torch's CPU generator is platform independent): He-scaled normal
    weights (x1.3 so that the signal survives seven ReLU layers), small normal biases, actor_logstd in [-0.5, 0.1]."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shp in shapes.items():
        fan_in = int(torch.tensor(shp[1:]).prod()) if len(shp) > 1 else 0
        if name == "actor_logstd":
            out[name] = torch.rand(shp, generator=g) * 0.6 - 0.5
        elif fan_in:
            out[name] = torch.randn(shp, generator=g) * (1.3 * math.sqrt(2.0 / fan_in))
        else:
            out[name] = 0.05 * torch.randn(shp, generator=g)
    for k in ("actor.mean_linear.weight", "critic.critic.2.weight"):
        out[k] = out[k] * 0.1                       # keep means / values O(1)
    return out


def param_shapes() -> Dict[str, tuple]:
    real = [16, 32, 64, 128, 256, 256, 128]
    shp = {}
    for net in ("actor", "critic"):
        cin = 2
        for l, co in enumerate(real):
            shp[f"{net}.conv.conv.{CONV_IDX[l]}.weight"] = (co, cin, 3, 3)
            shp[f"{net}.conv.conv.{CONV_IDX[l]}.bias"] = (co,)
            cin = co
        if net == "actor":
            shp["actor.mean_linear.weight"] = (N_ACT, 128)
            shp["actor.mean_linear.bias"] = (N_ACT,)
        else:
            shp["critic.critic.0.weight"] = (128, 128)
            shp["critic.critic.0.bias"] = (128,)
            shp["critic.critic.2.weight"] = (1, 128)
            shp["critic.critic.2.bias"] = (1,)
    shp["actor_logstd"] = (1, N_ACT)
    return shp


def cat_obs(state: torch.Tensor, obs: torch.Tensor) -> torch.Tensor:
    tile = state.reshape(state.size(0), 1, 1, 1).repeat(1, 1, obs.shape[2], obs.shape[3])
    return torch.cat([obs, tile], dim=1)


def _windows(z: torch.Tensor) -> torch.Tensor:
    """[B,C,H,W] -> [B,C,H/2,W/2,4] with the 2x2 window flattened as w = wy * 2 + wx."""
    B, C, H, W = z.shape
    return z.reshape(B, C, H // 2, 2, W // 2, 2).permute(0, 1, 2, 4, 3, 5).reshape(B, C, H // 2, W // 2, 4)


def encoder_forward(p, net: str, x: torch.Tensor, quant: bool = False, route=None, route_out=None) -> torch.Tensor:
    """route / route_out: forced / recorded discrete decisions per layer (max-pool arg-max, ReLU mask), see
    oracle/equiv_ref.py::encoder_forward."""
    for l in range(7):
        W, b = p[f"{net}.conv.conv.{CONV_IDX[l]}.weight"], p[f"{net}.conv.conv.{CONV_IDX[l]}.bias"]
        if quant and l > 0:
            W = bf16_ste(W)
        x = F.conv2d(x, W, b, padding=PADS[l])
        if route_out is not None:
            with torch.no_grad():
                if POOL[l]:
                    zw = _windows(x)
                    m = zw.max(-1).values
                    route_out.append(dict(arg=(zw == m.unsqueeze(-1)).float().argmax(-1), pos=m > 0))
                else:
                    route_out.append(dict(arg=None, pos=x > 0))
        if route is not None:
            r = route[l]
            if POOL[l]:
                x = _windows(x).gather(-1, r["arg"].unsqueeze(-1)).squeeze(-1)
            x = x * r["pos"].to(x.dtype)
            continue
        x = F.relu(x)
        if POOL[l]:
            x = F.max_pool2d(x, 2)
        if quant:
            x = bf16_ste(x)
    return x.reshape(x.shape[0], -1)


def evaluate(p, state, obs, action, quant: bool = False, route=None, route_out=None):
    """robot_actor_critic.evaluate (equivariant=False) with `action` given -> (log_prob [B], entropy [B], value [B])."""
    x = cat_obs(state, obs)
    q = bf16_ste if quant else (lambda t: t)
    fa = encoder_forward(p, "actor", x, quant, route["actor"] if route else None, route_out["actor"] if route_out is not None else None)
    fc = encoder_forward(p, "critic", x, quant, route["critic"] if route else None, route_out["critic"] if route_out is not None else None)
    mean = fa @ q(p["actor.mean_linear.weight"]).T + p["actor.mean_linear.bias"]
    log_std = p["actor_logstd"].expand_as(mean)
    std = torch.exp(log_std)
    var = std ** 2
    log_prob = -((action - mean) ** 2) / (2 * var) - std.log() - math.log(math.sqrt(2 * math.pi))
    entropy = 0.5 + 0.5 * math.log(2 * math.pi) + std.log()
    hpre = fc @ q(p["critic.critic.0.weight"]).T + p["critic.critic.0.bias"]
    if route_out is not None:
        route_out["group"] = dict(arg=None, pos=hpre.detach() > 0)
    h = hpre * route["group"]["pos"].to(hpre.dtype) if route is not None else F.relu(hpre)
    value = (h @ p["critic.critic.2.weight"].T + p["critic.critic.2.bias"]).reshape(-1)
    return log_prob.sum(1), entropy.sum(1), value


def update_loss(p, state, obs, action, oldlp, adv, ret, vold, clip_coeff=0.2, ent_c=0.01, vf_c=0.5, norm_adv=True,
                clip_vloss=True, quant: bool = False, route=None, route_out=None):
    """Loss of robot_ppo.update (robot_ppo.py:345-398) without the behaviour-cloning term (constant in the parameters)."""
    newlogprob, entropy, newvalue = evaluate(p, state, obs, action, quant=quant, route=route, route_out=route_out)
    ratio = (newlogprob - oldlp).exp()
    mb_adv = (adv - adv.mean()) / (adv.std() + 1e-8) if norm_adv else adv
    policy_loss = torch.max(-mb_adv * ratio, -mb_adv * torch.clamp(ratio, 1 - clip_coeff, 1 + clip_coeff)).mean()
    if clip_vloss:
        v_un = (newvalue - ret) ** 2
        v_cl = (vold + torch.clamp(newvalue - vold, -clip_coeff, clip_coeff) - ret) ** 2
        value_loss = 0.5 * torch.max(v_un, v_cl).mean()
    else:
        value_loss = 0.5 * ((newvalue - ret) ** 2).mean()
    value_loss = value_loss * vf_c
    loss = policy_loss - ent_c * entropy.mean() + value_loss
    return loss, dict(policy_loss=policy_loss.item(), value_loss=value_loss.item(), entropy=entropy.mean().item(), loss=loss.item())


def golden_inputs(B: int = 2):
    g = torch.Generator().manual_seed(7)
    obs = torch.rand(B, 1, 128, 128, generator=g) * 0.32
    state = torch.tensor([0.0, 1.0] * (B // 2) + [1.0] * (B % 2))
    action = torch.randn(B, N_ACT, generator=g)
    adv = torch.tensor([0.7, -1.3] * (B // 2) + [0.2] * (B % 2))
    ret = torch.tensor([0.5, -0.25] * (B // 2) + [0.1] * (B % 2))
    return obs, state, action, adv, ret
