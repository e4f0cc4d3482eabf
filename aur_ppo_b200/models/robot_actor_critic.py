"""The reference's robot_actor_critic (src/models/robot_actor_critic.py:19-157), equivariant branch, as an
nn.Module over the sm_100a kernels: same constructor arguments, `evaluate(state, obs, action=None)` ->
(actions, unscaled_actions, log_prob [B], entropy [B], value [B]), `value(state, obs)`, `decodeActions`,
`getActionFromPlan`, `test_action`.

The module owns the parameters (the p4 group-convolution filters psi of aur_ppo_b200/equiv.py, one
nn.Parameter each under `.actor` / `.critic`); `engine(batch)` hands the SAME storage to `EquivActorCritic`, so
the minibatch update of robot_ppo.update (src/robot_ppo.py:329-408) trains this module in place.  Inference runs
encoders and head GEMMs on tcgen05 and decodes the heads in `head_eval_kernel` (Normal sampling from Philox,
summed log-prob / entropy, action scaling).  No autograd graph is built and there is no CPU path: gradients come
from `engine(batch).loss_and_grads(...)`.

`equivariant=False` selects the plain `base_actor` / `base_critic` CNNs (src/nets/base_cnns.py:20-84, SURVEY.md §8(f)
rank 3) on the same kernels through `aur_ppo_b200/plain_cnn.py::PlainActorCritic`; its parameters carry the reference
modules' own state_dict names, see `reference_state_dicts()`.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import numpy as np
import torch
from torch import nn

from .. import _lib
from .. import plain_cnn
from ..equiv import ENC_FIELDS, N_ACT, EquivActorCritic, init_params
from ..kernels import _ptr, _stream, squashed_gaussian_sample, tc_gemm_bf16, tc_precision


class _PsiNet(nn.Module):
    """Parameter holder of one net: attribute names are the engine's keys with '.' -> '_'."""

    def __init__(self, net: str, params: Dict[str, torch.Tensor]):
        super().__init__()
        self._keys = []
        for k, v in params.items():
            if k.startswith(net + ".") or (net == "actor" and k == "actor_logstd"):
                name = k.replace(".", "_") if k == "actor_logstd" else k[len(net) + 1:].replace(".", "_")
                self.register_parameter(name, nn.Parameter(v, requires_grad=False))
                self._keys.append((k, name))

    def tensors(self) -> Dict[str, torch.Tensor]:
        """The Parameters themselves (requires_grad = False): the update engine re-points their storage into its flat
        parameter buffer (EquivActorCritic.__init__), so the objects must be the module's own, not `.data` aliases."""
        return {k: getattr(self, name) for k, name in self._keys}


class robot_actor_critic(nn.Module):
    def __init__(self, device, equivariant: bool, dx=0.02, dy=0.02, dz=0.02, dr=np.pi / 8, n_a=5, tau=0.001, seed: int = 0) -> None:
        super().__init__()
        if n_a != N_ACT:
            raise _lib.AurError("robot_actor_critic: the equivariant actor head has 5 action dims (equiv.py:70-80)")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.AurError("robot_actor_critic needs a CUDA device (no CPU path)")
        # robot_actor_critic.py:23-27: p_range is an int64 tensor [0, 1], the others fp32
        self.p_range = torch.tensor([0, 1])
        self.dtheta_range = torch.tensor([-dr, dr])
        self.dx_range = torch.tensor([-dx, dx])
        self.dy_range = torch.tensor([-dy, dy])
        self.dz_range = torch.tensor([-dz, dz])
        self.n_a = n_a
        self.equivariant = equivariant
        p = init_params(seed, self.device) if equivariant else plain_cnn.init_params(seed, self.device)
        self.actor = _PsiNet("actor", p)
        self.critic = _PsiNet("critic", p)
        self._engines: Dict[int, EquivActorCritic] = {}
        self._calls = 0
        self.seed = seed
        lohi = []
        for r in (self.p_range, self.dx_range, self.dy_range, self.dz_range, self.dtheta_range):
            lohi += [float(r[0].float()), float(r[1].float())]
        self._ranges = (ctypes.c_float * 10)(*lohi)

    def forward(self, act):
        pass

    # ------------------------------------------------------------------ engine
    def tensors(self) -> Dict[str, torch.Tensor]:
        t = self.actor.tensors()
        t.update(self.critic.tensors())
        return t

    def engine(self, batch: int, **kw) -> EquivActorCritic:
        """The update engine for minibatches of `batch` samples, sharing this module's parameter storage."""
        if batch not in self._engines:
            cls = EquivActorCritic if self.equivariant else plain_cnn.PlainActorCritic
            self._engines[batch] = cls(self.tensors(), batch, **kw)
        return self._engines[batch]

    def _run(self, state, obs, action, want_actor: bool, want_critic: bool):
        B = obs.shape[0]
        if obs.dim() != 4 or obs.shape[1:] != (1, 128, 128):
            raise _lib.AurError("robot_actor_critic: obs must be [B,1,128,128] (close_loop_block_picking heightmap)")
        Bp = (B + 7) // 8 * 8
        obs_d = obs.to(self.device, torch.float32)
        st_d = state.to(self.device, torch.float32).reshape(B)
        if Bp != B:
            obs_d = torch.cat([obs_d, obs_d.new_zeros(Bp - B, 1, 128, 128)])
            st_d = torch.cat([st_d, st_d.new_zeros(Bp - B)])
        obs_d, st_d = obs_d.contiguous(), st_d.contiguous()
        e = self.engine(Bp)
        a_out = c_pre = None
        with tc_precision(e.P):
            e._expand()
            if want_actor:
                e._encoder_forward("actor", st_d, obs_d)
                a_out = tc_gemm_bf16(e.enc["actor"].feat, e._w["actor.head"][0])
            if want_critic:
                e._encoder_forward("critic", st_d, obs_d)
                c_pre = tc_gemm_bf16(e.enc["critic"].feat, e._w["critic.head1"][0])
        dev = self.device
        f = lambda *s: torch.empty(*s, device=dev)
        out = dict(unscaled=f(Bp, 5), scaled=f(Bp, 5), logp=f(Bp), ent=f(Bp), value=f(Bp), mean=f(Bp, 5), logstd=f(Bp, 5))
        act_d = None
        if action is not None:
            act_d = action.to(dev, torch.float32).reshape(B, 5)
            if Bp != B:
                act_d = torch.cat([act_d, act_d.new_zeros(Bp - B, 5)])
            act_d = act_d.contiguous()
        self._calls += 1
        outs = (out["unscaled"].data_ptr(), out["scaled"].data_ptr(), out["logp"].data_ptr(), out["ent"].data_ptr(),
                out["value"].data_ptr(), out["mean"].data_ptr(), out["logstd"].data_ptr())
        with torch.cuda.device(dev):
            if self.equivariant:
                a_bias = torch.cat([torch.zeros(2, device=dev), e.p["actor.head.bias_triv"]]).contiguous()
                w2 = e.p["critic.head2.w"].reshape(-1).contiguous()
                rc = _lib.lib().aur_equiv_head_eval(
                    Bp, _ptr(a_out), a_bias.data_ptr(), _ptr(c_pre), e._w["critic.head1"][2].data_ptr(), w2.data_ptr(),
                    e.p["critic.head2.bias"].data_ptr(), _ptr(act_d), self.seed & 0xFFFFFFFFFFFFFFFF, self._calls, self._ranges,
                    *outs, _stream())
            else:
                w2 = e.p["critic.critic.2.weight"].reshape(-1).contiguous()
                rc = _lib.lib().aur_plain_head_eval(
                    Bp, _ptr(a_out), e.p["actor.mean_linear.bias"].data_ptr(), e.p["actor_logstd"].data_ptr(), _ptr(c_pre),
                    e._w["critic.head1"][2].data_ptr(), w2.data_ptr(), e.p["critic.critic.2.bias"].data_ptr(), _ptr(act_d),
                    self.seed & 0xFFFFFFFFFFFFFFFF, self._calls, self._ranges, *outs, _stream())
        _lib.check(rc, "aur_equiv_head_eval")
        return {k: v[:B] for k, v in out.items()}

    # ------------------------------------------------------------------ reference API
    def value(self, state, obs):
        """robot_actor_critic.py:57-60 -> critic(cat_obs), [B,1] like EquivariantCritic.forward's reshape(batch, -1)."""
        return self._run(state, obs, None, False, True)["value"].reshape(-1, 1)

    def decodeActions(self, *args):
        """robot_actor_critic.py:63-82 (elementwise torch on [B] vectors, as in the reference)."""
        rng = [self.p_range, self.dx_range, self.dy_range, self.dz_range, self.dtheta_range][:len(args)]
        scaled = [0.5 * (u + 1) * (r[1] - r[0]) + r[0] for u, r in zip(args, rng)]
        return torch.stack(list(args), dim=1), torch.stack(scaled, dim=1)

    def getActionFromPlan(self, plan):
        """robot_actor_critic.py:85-102."""
        rng = [self.p_range, self.dx_range, self.dy_range, self.dz_range, self.dtheta_range][:self.n_a]
        un = []
        for i, r in enumerate(rng):
            a = plan[:, i].clamp(*r)
            un.append(2 * (a - r[0]) / (r[1] - r[0]) - 1)
        return self.decodeActions(*un)

    def evaluate(self, state, obs, action=None):
        """robot_actor_critic.py:104-131 -> (actions, unscaled_actions, log_prob.sum(1), entropy.sum(1), value [B,1])."""
        o = self._run(state, obs, action, True, True)
        return o["scaled"], o["unscaled"], o["logp"], o["ent"], o["value"].reshape(-1, 1)

    def evaluate_pretrain(self, state, obs, action=None):
        """robot_actor_critic.py:134-149 (the actor alone, for the behaviour-cloning phase): action = tanh(Normal(mean,
        exp(logstd)).rsample()) - or tanh(action) when one is given - through decodeActions -> (scaled, unscaled), both fp16
        as in the reference.  The draw is the squashed-Gaussian kernel's Philox stream (aur_squashed_gaussian_sample)."""
        o = self._run(state, obs, None, True, False)
        if action is None:
            self._calls += 1
            y = squashed_gaussian_sample(o["mean"].contiguous(), o["logstd"].contiguous(), None, seed=self.seed, stream_id=self._calls)[0]
        else:
            y = torch.tanh(action.to(self.device, torch.float32).reshape(-1, self.n_a))
        unscaled, scaled = self.decodeActions(*[y[:, i] for i in range(self.n_a)])
        return scaled.to(torch.float16), unscaled.to(torch.float16)

    @staticmethod
    def pretrain_loss(b_actions, b_true_actions, mb_inds):
        """robot_ppo.py:291-307 `pretrain_update`: mse_loss(b_actions[mb], b_true_actions[mb]).  In the reference the stored
        actions are buffer tensors with no graph back to the policy (`.requires_grad_(True)` makes them leaves), so
        `expert_loss.backward()` reaches no parameter and `optimizer.step()` leaves the policy unchanged: the phase's only
        numerical product is this scalar, which is what is restated here."""
        return torch.mean((b_actions[mb_inds].float() - b_true_actions[mb_inds].float()) ** 2)

    def test_action(self, state, obs):
        """robot_actor_critic.py:152-157: decodeActions(tanh(mean))."""
        mean = torch.tanh(self._run(state, obs, None, True, False)["mean"])
        return self.decodeActions(*[mean[:, i] for i in range(self.n_a)])

    # robot_ppo.py:502-507 checkpoint layout
    def checkpoint_dict(self, optimizer_state=None) -> dict:
        return {"actor_state": self.actor.state_dict(), "critic_state": self.critic.state_dict(), "optimizer_state": optimizer_state}

    def load_checkpoint_dict(self, d: dict) -> None:
        self.actor.load_state_dict(d["actor_state"])
        self.critic.load_state_dict(d["critic_state"])

    # plain CNN only: the reference modules' own state_dicts (base_actor / base_critic keys, robot_actor_critic.actor_logstd)
    def reference_state_dicts(self) -> dict:
        if self.equivariant:
            raise _lib.AurError("reference_state_dicts: e2cnn checkpoints are not interchangeable (DESIGN.md section 5)")
        t = self.tensors()
        return {"actor_state": {k[len("actor."):]: v.detach().clone() for k, v in t.items() if k.startswith("actor.")},
                "critic_state": {k[len("critic."):]: v.detach().clone() for k, v in t.items() if k.startswith("critic.")},
                "actor_logstd": t["actor_logstd"].detach().clone()}

    def load_reference_state_dicts(self, d: dict) -> None:
        t = self.tensors()
        with torch.no_grad():
            for k, v in d["actor_state"].items():
                t["actor." + k].copy_(v)
            for k, v in d["critic_state"].items():
                t["critic." + k].copy_(v)
            if "actor_logstd" in d:
                t["actor_logstd"].copy_(d["actor_logstd"])
