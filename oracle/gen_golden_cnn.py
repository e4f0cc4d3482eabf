"""Generates tests/golden/plain_cnn.npz from the reference's OWN classes (run in the build container only:
needs /root/reference; the fixture travels, the reference does not).

  python oracle/gen_golden_cnn.py

`base_actor` / `base_critic` (src/nets/base_cnns.py:57-84) are instantiated unmodified, loaded with the deterministic
formula parameters of oracle/cnn_ref.py, and evaluated exactly as robot_actor_critic.evaluate (equivariant=False,
src/models/robot_actor_critic.py:104-131) and the loss of robot_ppo.update (src/robot_ppo.py:345-398) do; the file stores
the inputs' seed-derived tensors, log-probs, entropies, values, loss terms and the L2 norm + 4 probe entries of every
parameter gradient (the full gradients would be 10 MB)."""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import cnn_ref as C  # noqa: E402
from src.nets.base_cnns import base_actor, base_critic  # noqa: E402  (the reference's own code)


def main():
    torch.manual_seed(0)
    p = C.formula_params(C.param_shapes())
    actor, critic = base_actor(), base_critic()
    actor.load_state_dict({k[len("actor."):]: v for k, v in p.items() if k.startswith("actor.")})
    critic.load_state_dict({k[len("critic."):]: v for k, v in p.items() if k.startswith("critic.")})
    actor_logstd = torch.nn.Parameter(p["actor_logstd"].clone())
    obs, state, action, adv, ret = C.golden_inputs(2)
    # robot_actor_critic.evaluate, equivariant = False (robot_actor_critic.py:104-131)
    state_tile = state.reshape(state.size(0), 1, 1, 1).repeat(1, 1, obs.shape[2], obs.shape[3])
    cat_obs = torch.cat([obs, state_tile], dim=1)
    action_mean = actor(cat_obs)
    action_logstd = actor_logstd.expand_as(action_mean)
    dist = torch.distributions.Normal(action_mean, torch.exp(action_logstd))
    log_prob = dist.log_prob(action).sum(1)
    entropy = dist.entropy().sum(1)
    value = critic(cat_obs)
    oldlp = log_prob.detach() + torch.tensor([0.1, -0.3])
    vold = value.detach().reshape(-1) + torch.tensor([0.05, -0.4])
    # robot_ppo.update loss (robot_ppo.py:345-398), clip 0.2, entropy 0.01, value 0.5, norm_adv, clip_vloss
    newvalue = value.view(-1)
    ratio = (log_prob - oldlp).exp()
    mb_adv = (adv - adv.mean()) / (adv.std() + 1e-8)
    pl = torch.max(-mb_adv * ratio, -mb_adv * torch.clamp(ratio, 0.8, 1.2)).mean()
    v_un = (newvalue - ret) ** 2
    v_cl = (vold + torch.clamp(newvalue - vold, -0.2, 0.2) - ret) ** 2
    vl = 0.5 * torch.max(v_un, v_cl).mean() * 0.5
    loss = pl - 0.01 * entropy.mean() + vl
    loss.backward()
    out = dict(logp=log_prob.detach().numpy(), entropy=entropy.detach().numpy(), value=newvalue.detach().numpy(),
               oldlp=oldlp.numpy(), vold=vold.numpy(), policy_loss=pl.item(), value_loss=vl.item(), loss=loss.item(),
               mean=action_mean.detach().numpy())
    named = {"actor." + k: v for k, v in actor.named_parameters()}
    named.update({"critic." + k: v for k, v in critic.named_parameters()})
    named["actor_logstd"] = actor_logstd
    names = list(C.param_shapes().keys())
    out["grad_names"] = np.array(names)
    out["grad_norm"] = np.array([float(named[n].grad.norm()) for n in names])
    out["grad_probe"] = np.array([named[n].grad.reshape(-1)[[0, named[n].numel() // 3, named[n].numel() // 2, -1]].numpy() for n in names])
    path = os.path.join(ROOT, "tests", "golden", "plain_cnn.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", "loss", loss.item(), "logp", log_prob.tolist(), "value", newvalue.tolist())


if __name__ == "__main__":
    main()
