"""Plain (non-equivariant) CNN actor-critic update on the B200 - SURVEY.md §8(f) rank 3: host-side mirror of
`robot_actor_critic(equivariant=False).evaluate` (src/models/robot_actor_critic.py:41-51,104-131) over `base_actor` /
`base_critic` / `base_encoder` (src/nets/base_cnns.py:20-84) and of the minibatch step of `robot_ppo.update`
(src/robot_ppo.py:329-408).

The encoder has the spatial structure of the equivariant one (3x3 convolutions, pad 1 x5 / pad 0 x2, 2x2 max-pools
after layers 0-3 and 5) with channels 2 -> 16 -> 32 -> 64 -> 128 -> 256 -> 256 -> 128, so it runs on the SAME sm_100a
kernels (`EquivActorCritic`'s forward / backward machinery, csrc/equiv*.cu) with channel counts padded up to the
64-channel K chunk of the implicit-GEMM convolution:

* layer 0 runs the direct layer-0 kernels in their PLAIN instantiation (`aur_plain_conv0`, `aur_plain_conv0_wgrad`): the
  16 filters land in channels 0..15 of the 64-channel buffer the 64-wide contraction of layer 1 reads, the other 48
  channels stay zero (a first version passed the filters as 16 equivariant "fields" and paid 4x the layer-0 work);
* layers 1-2 (16 -> 32 -> 64 real channels) run as 64 -> 64 contractions with zero padding;
* layer 6 (3x3 valid -> 1x1) and the heads are dense GEMMs on tcgen05 (`aur_tc_gemm_bf16`).

Parameters are keyed by the reference modules' own state_dict names (`actor.conv.conv.0.weight`, ...,
`actor.mean_linear.weight`, `critic.critic.0.weight`, `actor_logstd`), so reference checkpoints load without renaming.
Scatter / gather between those tensors and the padded contraction matrices is torch indexing on <= 0.6 M elements
(plumbing).  No CPU path.
"""
from __future__ import annotations

import ctypes
import math
from typing import Dict, Optional

import torch

from . import _lib
from .equiv import EquivActorCritic, N_ACT, _chk
from .kernels import _stream, adv_moments, tc_gemm_bf16, tc_precision

CONV_IDX = [0, 3, 6, 9, 12, 14, 17]                 # positions of the Conv2d layers in base_encoder.conv
REAL = [16, 32, 64, 128, 256, 256, 128]             # real output channels of the seven convolutions
PAD = [64, 64, 64, 128, 256, 256]                   # stored (padded) channels of the activations a[0..5]


def conv_key(net: str, l: int, what: str) -> str:
    return f"{net}.conv.conv.{CONV_IDX[l]}.{what}"


def init_params(seed: int = 0, device="cuda") -> Dict[str, torch.Tensor]:
    """weights_init of the reference (base_cnns.py:12-17): xavier_normal_ convolutions, xavier_uniform_ / zero-bias linears;
    convolution biases keep torch's default U(-1/sqrt(fan_in), 1/sqrt(fan_in))."""
    g = torch.Generator().manual_seed(seed)
    p: Dict[str, torch.Tensor] = {}
    for net in ("actor", "critic"):
        cin = 2
        for l, co in enumerate(REAL):
            fan_in, fan_out = cin * 9, co * 9
            p[conv_key(net, l, "weight")] = torch.randn(co, cin, 3, 3, generator=g) * math.sqrt(2.0 / (fan_in + fan_out))
            p[conv_key(net, l, "bias")] = (torch.rand(co, generator=g) * 2 - 1) / math.sqrt(fan_in)
            cin = co

    def xavier_uniform(o, i):
        a = math.sqrt(6.0 / (i + o))
        return (torch.rand(o, i, generator=g) * 2 - 1) * a
    p["actor.mean_linear.weight"] = xavier_uniform(N_ACT, 128)
    p["actor.mean_linear.bias"] = torch.zeros(N_ACT)
    p["critic.critic.0.weight"] = xavier_uniform(128, 128)
    p["critic.critic.0.bias"] = torch.zeros(128)
    p["critic.critic.2.weight"] = xavier_uniform(1, 128)
    p["critic.critic.2.bias"] = torch.zeros(1)
    p["actor_logstd"] = torch.zeros(1, N_ACT)
    return {k: v.to(device).contiguous() for k, v in p.items()}


class PlainActorCritic(EquivActorCritic):
    CH = tuple(PAD)
    FEAT = 128
    D_HEAD = 267

    def __init__(self, params: Dict[str, torch.Tensor], batch: int, lr: float = 3e-4, eps: float = 1e-5, betas=(0.9, 0.999),
                 split: bool = False, precision: Optional[str] = None):
        super().__init__(params, batch, lr, eps, betas, split=split, precision=precision)
        dev = self.dev
        # input-channel positions of each layer's real channels inside the padded activation feeding it
        self._in_pos = [None] + [torch.arange(REAL[l - 1], device=dev) for l in range(1, 7)]

    # ------------------------------------------------------------------ weights
    def _layer0_params(self, net: str):
        return self.p[conv_key(net, 0, "weight")], self.p[conv_key(net, 0, "bias")]

    def _expand(self):
        w = {}
        for net in ("actor", "critic"):
            for l in range(1, 6):
                W, b = self.p[conv_key(net, l, "weight")], self.p[conv_key(net, l, "bias")]
                co, ci = W.shape[0], W.shape[1]
                Cin, Cout = self.CH[l - 1], self.CH[l]
                dense = torch.zeros(Cout, Cin, 3, 3, device=self.dev)
                dense[:co, self._in_pos[l]] = W
                wm = self._bf(dense.permute(0, 2, 3, 1).reshape(Cout, 9, Cin))
                wt = self._bf(torch.flip(dense, dims=(2, 3)).permute(1, 2, 3, 0).reshape(Cin, 9, Cout))
                bias = torch.zeros(Cout, device=self.dev)
                bias[:co] = b
                w[f"{net}.{l}"] = (wm, wt, bias)
            W6, b6 = self.p[conv_key(net, 6, "weight")], self.p[conv_key(net, 6, "bias")]
            wm6 = self._bf(W6.permute(0, 2, 3, 1).reshape(128, 9 * 256))
            w[f"{net}.6"] = (wm6, wm6.transpose(1, 2).contiguous(), b6.contiguous())
        Wa = torch.zeros(16, 128, device=self.dev)
        Wa[:N_ACT] = self.p["actor.mean_linear.weight"]
        w["actor.head"] = (self._bf(Wa), self._bf(Wa.t().contiguous()))
        W1 = self.p["critic.critic.0.weight"]
        w["critic.head1"] = (self._bf(W1), self._bf(W1.t().contiguous()), self.p["critic.critic.0.bias"].contiguous())
        self._w = w

    # ---------------------------------------------------------------- gradients
    def _store_wgrad(self, net: str, l: int, dw: torch.Tensor, Cout: int, Cin: int):
        if l == 6:
            g = dw.reshape(128, 3, 3, 256).permute(0, 3, 1, 2)
        else:
            co = REAL[l]
            g = dw.reshape(Cout, 3, 3, Cin)[:co][:, :, :, self._in_pos[l]].permute(0, 3, 1, 2)
        self.grads[conv_key(net, l, "weight")].copy_(g)

    def _store_bgrad(self, net: str, l: int, dy2d: torch.Tensor, Q: int, Cout: int):
        out = torch.zeros(Cout, device=self.dev)           # the kernel accumulates
        with torch.cuda.device(self.dev):
            for pl in range(self.P):
                _chk(_lib.lib().aur_colsum_bf16(Q, Cout, dy2d[pl].data_ptr(), 1, out.data_ptr(), _stream()), "aur_colsum_bf16")
        self.grads[conv_key(net, l, "bias")].copy_(out[:REAL[l]])

    def _bgrad_begin(self, net: str, l: int, C: int):
        return torch.zeros(C, device=self.dev), 1

    def _bgrad_end(self, net: str, l: int, acc: torch.Tensor):
        self.grads[conv_key(net, l, "bias")].copy_(acc[:REAL[l]])

    def _conv0(self, net: str, state, obs, e):
        W0, b0 = self._layer0_params(net)
        with torch.cuda.device(self.dev):
            _chk(_lib.lib().aur_plain_conv0(obs.data_ptr(), state.data_ptr(), W0.data_ptr(), b0.data_ptr(), self.B, e.a[0].data_ptr(),
                                            e.arg[0].data_ptr(), _stream()), "aur_plain_conv0")

    def _layer0_wgrad(self, net: str, state, obs, dprev, e):
        with torch.cuda.device(self.dev):
            _chk(_lib.lib().aur_plain_conv0_wgrad(obs.data_ptr(), state.data_ptr(), dprev.data_ptr(), e.a[0].data_ptr(),
                                                  e.arg[0].data_ptr(), self.B, self.ws.data_ptr(),
                                                  self.grads[conv_key(net, 0, "weight")].data_ptr(),
                                                  self.grads[conv_key(net, 0, "bias")].data_ptr(), _stream()), "aur_plain_conv0_wgrad")

    # ------------------------------------------------------------------- update
    def _loss_and_grads(self, state, obs, action, oldlp, adv, ret, vold, clip_coeff, entropy_coeff, value_coeff, norm_adv,
                        clip_vloss) -> torch.Tensor:
        L = _lib.lib()
        B = self.B
        self._flat["g"].zero_()
        self.stats.zero_(); self.d_head.zero_()
        a_out, c_pre = self._forward(state, obs)
        if norm_adv:
            adv_moments(adv, self.moments, self._mom_ws)
        d_a_out = self._empty(B, 16)
        d_c_h = self._empty(B, 128)
        h = _lib.PlainHeadArgs()
        h.B, h.clip_vloss, h.m_total = B, int(bool(clip_vloss)), getattr(self, "_m_total", B)
        h.a_out, h.a_bias, h.actor_logstd = a_out.data_ptr(), self.p["actor.mean_linear.bias"].data_ptr(), self.p["actor_logstd"].data_ptr()
        h.c_pre, h.c_bias1 = c_pre.data_ptr(), self._w["critic.head1"][2].data_ptr()
        w2 = self.p["critic.critic.2.weight"].reshape(-1).contiguous()
        h.c_w2, h.c_b2 = w2.data_ptr(), self.p["critic.critic.2.bias"].data_ptr()
        h.action, h.oldlp, h.adv, h.ret, h.vold = (t.data_ptr() for t in (action, oldlp, adv, ret, vold))
        h.adv_moments = self.moments.data_ptr() if norm_adv else None
        h.clip_coeff, h.entropy_coeff, h.value_coeff = float(clip_coeff), float(entropy_coeff), float(value_coeff)
        h.d_a_out, h.d_c_h, h.d_head, h.stats = d_a_out.data_ptr(), d_c_h.data_ptr(), self.d_head.data_ptr(), self.stats.data_ptr()
        h.value_out, h.logp_out = self.value.data_ptr(), self.logp.data_ptr()
        with torch.cuda.device(self.dev):
            _chk(L.aur_plain_head_loss(ctypes.byref(h), _stream()), "aur_plain_head_loss")
        self._last_head = (a_out, c_pre, d_a_out, d_c_h)
        fa, fc = self.enc["actor"].feat, self.enc["critic"].feat
        dWa = tc_gemm_bf16(self._t(d_a_out), self._t(fa))                       # [16,128]
        self.grads["actor.mean_linear.weight"].copy_(dWa[:N_ACT])
        self.grads["actor.mean_linear.bias"].copy_(self.d_head[0:5])
        self.grads["actor_logstd"].copy_(self.d_head[5:10].reshape(1, N_ACT))
        dW1 = tc_gemm_bf16(self._t(d_c_h), self._t(fc))                         # [128,128]
        self.grads["critic.critic.0.weight"].copy_(dW1)
        self.grads["critic.critic.2.weight"].copy_(self.d_head[10:138].reshape(1, 128))
        self.grads["critic.critic.2.bias"].copy_(self.d_head[138:139])
        self.grads["critic.critic.0.bias"].copy_(self.d_head[139:267])
        dfa = tc_gemm_bf16(d_a_out, self._w["actor.head"][1])                   # [B,128] fp32
        dfc = tc_gemm_bf16(d_c_h, self._w["critic.head1"][1])
        self._encoder_backward("actor", state, obs, dfa)
        self._encoder_backward("critic", state, obs, dfc)
        return self.stats / B

    # apply(): the base class - robot_ppo.py:401 clips `self.policy.actor.parameters()`, i.e. every actor.* tensor (actor_logstd
    # is a parameter of the policy module, not of `.actor`, so it is not clipped), then one Adam over everything
