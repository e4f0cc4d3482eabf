"""Generate tests/golden/*.npz from the REFERENCE'S OWN CODE (run in the build
container only: /root/reference does not exist on the GPU box).

TEST INFRASTRUCTURE ONLY.  Usage:  python oracle/gen_golden.py

The reference is imported unmodified from /root/reference under four import
shims: stub modules `gym`, `matplotlib(.pyplot)`, and
alias modules `nets` / `models` that make ppo.py:3 and ppo.py:7 resolvable.
Vectors written:

  gae.npz     ppo.run_gae / ppo.normal_advantage (ppo.py:125-157) on the
              SURVEY known-answer case and on seeded random [T,N] cases
  model.npz   actor_critic.evaluate / .value (models/actor_critic.py:31-51)
              for the discrete and the continuous policy, fixed parameters
  update.npz  one minibatch step restating ppo.py:220-269 around the
              reference actor_critic + torch.optim.Adam(eps=1e-5): loss parts,
              diagnostics, gradients (pre- and post-clip), updated parameters,
              for two consecutive steps (exercises Adam bias correction)
  squash.npz  PPOGaussianPolicyBase.sample (nets/nets.py:90-105) on fixed
              mean/log_std/action
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def import_reference():
    sys.path.insert(0, REF)
    for name in ["gym", "matplotlib", "matplotlib.pyplot"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    import src.nets.nets as ref_nets
    nets_alias = types.ModuleType("nets")
    for n in ("discrete_net", "continuous_net", "critic"):
        setattr(nets_alias, n, getattr(ref_nets, n))
    sys.modules["nets"] = nets_alias
    import src.models.actor_critic as ref_ac
    models_alias = types.ModuleType("models")
    models_alias.actor_critic = ref_ac.actor_critic
    sys.modules["models"] = models_alias
    import src.ppo as ref_ppo
    return ref_ppo, ref_ac, ref_nets


def fill_params(model):
    """i-th parameter tensor <- 0.1*sin(0.37*k + i), k the flat index."""
    with torch.no_grad():
        for i, p in enumerate(model.parameters()):
            k = torch.arange(p.numel(), dtype=torch.float64)
            p.copy_((0.1 * torch.sin(0.37 * k + i)).to(torch.float32).reshape(p.shape))


def gae_cases(ref_ppo):
    out = {}
    cases = []
    # SURVEY known-answer case
    T, N = 4, 2
    values = torch.tensor([[0.5 * np.sin(np.float32(2 * t + n)) for n in range(N)] for t in range(T)], dtype=torch.float32)
    term = torch.zeros(T, N); term[2, 1] = 1.0
    cases.append(("ka", torch.ones(T, N), values, term, torch.tensor([0.25, -0.75]), torch.tensor([0.0, 1.0]), 0.99, 0.95))
    g = torch.Generator().manual_seed(0)
    for name, T, N, pd, gamma, lam in [("r0", 128, 4, 0.05, 0.99, 0.95), ("r1", 37, 13, 0.2, 0.9, 0.8),
                                       ("r2", 256, 64, 1 / 200, 0.99, 0.95), ("r3", 1, 5, 0.5, 0.99, 0.95),
                                       ("r4", 2048, 3, 0.01, 0.999, 1.0)]:
        rew = torch.rand(T, N, generator=g)
        val = torch.randn(T, N, generator=g)
        term = (torch.rand(T, N, generator=g) < pd).float()
        nv = torch.randn(N, generator=g)
        nd = (torch.rand(N, generator=g) < pd).float()
        cases.append((name, rew, val, term, nv, nd, gamma, lam))
    for name, rew, val, term, nv, nd, gamma, lam in cases:
        fake = types.SimpleNamespace(
            buffer=types.SimpleNamespace(rewards=rew, values=val, terminals=term),
            num_steps=rew.shape[0], gamma=gamma, gae_lambda=lam)
        ret, adv = ref_ppo.ppo.run_gae(fake, nv, nd)
        ret2, adv2 = ref_ppo.ppo.normal_advantage(fake, nv, nd)
        for k, v in dict(rew=rew, val=val, term=term, nv=nv, nd=nd, gamma=np.float64(gamma), lam=np.float64(lam),
                         gae_ret=ret, gae_adv=adv, mc_ret=ret2, mc_adv=adv2).items():
            out[f"{name}_{k}"] = np.asarray(v)
    out["names"] = np.array([c[0] for c in cases])
    return out


def model_cases(ref_ac):
    out = {}
    for tag, sd, ad, cont, nl in [("disc", 4, 2, False, 2), ("cont", 3, (1,), True, 2), ("disc3", 4, 2, False, 3),
                                  ("cont2", 5, (2,), True, 2)]:
        m = ref_ac.actor_critic(sd, ad, 64, nl, 0.0, cont)
        fill_params(m)
        if cont:
            with torch.no_grad():
                m.actor_logstd.copy_(torch.linspace(-0.3, 0.2, m.actor_logstd.numel()).reshape(1, -1))
        B = 16
        obs = (0.7 * torch.cos(torch.arange(B * sd, dtype=torch.float64) * 1.3)).float().reshape(B, sd)
        if cont:
            A = int(np.prod(ad))
            act = (torch.sin(torch.arange(B * A, dtype=torch.float64) * 0.9)).float().reshape(B, A)
        else:
            act = (torch.arange(B) % ad).long()
        with torch.no_grad():
            a, lp, ent, v = m.evaluate(obs, act)
            val = m.value(obs)
            head = m.actor(obs)
        out[f"{tag}_names"] = np.array([n for n, _ in m.named_parameters()])
        for n, p in m.named_parameters():
            out[f"{tag}_p_{n}"] = p.detach().numpy()
        for k, v_ in dict(obs=obs, act=act, logp=lp, ent=ent, value=v, value_flat=val, head=head).items():
            out[f"{tag}_{k}"] = v_.numpy()
    return out


def update_cases(ref_ac):
    """ppo.py:220-269 restated around the reference model (the update is inline in
    train() and not callable)."""
    out = {}
    for tag, sd, ad, cont, B, norm_adv, clip_vloss in [("disc", 4, 2, False, 8, True, True),
                                                       ("disc_big", 4, 2, False, 300, True, True),
                                                       ("cont", 3, (1,), True, 64, True, True),
                                                       ("disc_nonorm", 4, 2, False, 32, False, True),
                                                       ("disc_novclip", 4, 2, False, 32, True, False)]:
        m = ref_ac.actor_critic(sd, ad, 64, 2, 0.0, cont)
        fill_params(m)
        clip_coeff, ent_c, vf_c, mgn, lr = 0.2, 0.01, 0.5, 0.5, 2.5e-4
        opt = torch.optim.Adam(m.parameters(), lr=lr, eps=1e-5)
        if tag == "disc":
            obs = (0.05 * torch.cos(torch.arange(B * sd, dtype=torch.float32))).reshape(B, sd)
            act = (torch.arange(B) % 2).float()
            oldlp = torch.full((B,), -0.7); adv = torch.sin(torch.arange(B, dtype=torch.float32))
            ret = torch.ones(B); vold = torch.full((B,), 0.5)
        else:
            g = torch.Generator().manual_seed(7)
            obs = torch.randn(B, sd, generator=g) * 0.5
            if cont:
                act = torch.randn(B, 1, generator=g)
                oldlp = -0.9 - 0.5 * act.flatten() ** 2 + 0.1 * torch.randn(B, generator=g)
            else:
                act = torch.randint(0, 2, (B,), generator=g).float()
                oldlp = -0.69 + 0.3 * torch.randn(B, generator=g)
            adv = torch.randn(B, generator=g) * 2
            ret = torch.randn(B, generator=g)
            vold = ret + 0.3 * torch.randn(B, generator=g)
        names = [n for n, _ in m.named_parameters()]
        out[f"{tag}_names"] = np.array(names)
        for n, p in m.named_parameters():
            out[f"{tag}_p0_{n}"] = p.detach().numpy().copy()
        for k, v_ in dict(obs=obs, act=act, oldlp=oldlp, adv=adv, ret=ret, vold=vold).items():
            out[f"{tag}_{k}"] = v_.numpy()
        out[f"{tag}_hyper"] = np.array([clip_coeff, ent_c, vf_c, mgn, lr, float(norm_adv), float(clip_vloss)])
        for step in range(2):
            _, newlogprob, entropy, newvalue = m.evaluate(obs, act if cont else act)
            log_ratio = newlogprob - oldlp
            ratio = log_ratio.exp()
            with torch.no_grad():
                old_approx_kl = (-log_ratio).mean()
                approx_kl = ((ratio - 1) - log_ratio).mean()
                clipfrac = ((ratio - 1.0).abs() > clip_coeff).float().mean()
            mb_adv = adv
            if norm_adv:
                mb_adv = (mb_adv - mb_adv.mean()) / (mb_adv.std() + 1e-8)
            loss_one = -mb_adv * ratio
            loss_two = -mb_adv * torch.clamp(ratio, 1 - clip_coeff, 1 + clip_coeff)
            policy_loss = torch.max(loss_one, loss_two).mean()
            newvalue = newvalue.view(-1)
            if clip_vloss:
                v_loss_unclipped = (newvalue - ret) ** 2
                v_clipped = vold + torch.clamp(newvalue - vold, -clip_coeff, clip_coeff)
                v_loss_clipped = (v_clipped - ret) ** 2
                value_loss = 0.5 * torch.max(v_loss_unclipped, v_loss_clipped).mean()
            else:
                value_loss = 0.5 * ((newvalue - vold) ** 2).mean()     # ppo.py:261 (uses b_values)
            entropy_loss = entropy.mean()
            loss = policy_loss - ent_c * entropy_loss + value_loss * vf_c
            opt.zero_grad()
            loss.backward()
            for n, p in m.named_parameters():
                out[f"{tag}_s{step}_g_{n}"] = p.grad.detach().numpy().copy()
            gnorm = torch.nn.utils.clip_grad_norm_(m.parameters(), mgn)
            opt.step()
            for n, p in m.named_parameters():
                out[f"{tag}_s{step}_p_{n}"] = p.detach().numpy().copy()
            out[f"{tag}_s{step}_stats"] = np.array([policy_loss.item(), value_loss.item(), entropy_loss.item(), loss.item(),
                                                    old_approx_kl.item(), approx_kl.item(), clipfrac.item(), gnorm.item()])
            out[f"{tag}_s{step}_newlogp"] = newlogprob.detach().numpy().copy()
            out[f"{tag}_s{step}_newvalue"] = newvalue.detach().numpy().copy()
    return out


def squash_cases(ref_nets):
    class P(ref_nets.PPOGaussianPolicyBase):
        def forward(self, x):
            return x[:, :5], x[:, 5:]
    B = 12
    k = torch.arange(B * 10, dtype=torch.float64)
    x = torch.cat([(torch.sin(0.7 * k[:B * 5])).reshape(B, 5), (0.5 * torch.cos(0.3 * k[B * 5:]) - 0.5).reshape(B, 5)], 1).float()
    act = (1.5 * torch.sin(1.1 * k[:B * 5] + 0.3)).reshape(B, 5).float()
    with torch.no_grad():
        a, lp, mean, ent = P().sample(x, act)
    return dict(x=x.numpy(), act=act.numpy(), a=a.numpy(), logp=lp.numpy(), mean=mean.numpy(), ent=ent.numpy())


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)
    ref_ppo, ref_ac, ref_nets = import_reference()
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "gae.npz"), **gae_cases(ref_ppo))
    np.savez_compressed(os.path.join(OUT, "model.npz"), **model_cases(ref_ac))
    np.savez_compressed(os.path.join(OUT, "update.npz"), **update_cases(ref_ac))
    np.savez_compressed(os.path.join(OUT, "squash.npz"), **squash_cases(ref_nets))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
