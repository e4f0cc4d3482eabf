"""Equivariant actor-critic update on the B200 (row X): host-side mirror of
`robot_actor_critic.evaluate` (src/models/robot_actor_critic.py:104-131) over
`EquivariantActor` / `EquivariantCritic` (src/nets/equiv.py:65-157) and of the minibatch step of
`robot_ppo.update` (src/robot_ppo.py:329-408: PPO loss, clip_grad_norm_ on the ACTOR only, one Adam
over actor + critic, eps=1e-5).

Every convolution / dense contraction runs on tcgen05 tensor cores through the C ABI
(aur_conv3x3_bf16, aur_wgrad3x3_bf16, aur_tc_gemm_bf16); activations are bf16 NHWC with explicit
halos, accumulation is fp32, parameters / gradients / Adam moments are fp32.

Operand precisions (aur_tc_set_precision), `precision=`:
  "fp32"   3 planes: every bf16 tensor is a stack hi / mid / lo (3 x 8 = 24 mantissa bits) and every contraction issues the
           six products above 2^-24, accumulated in TMEM in chunks of 32 MMA steps that are promoted into fp32 registers (the
           tensor core's own accumulator truncates): fp32-EQUIVALENT, like the reference (equiv.py / robot_ppo.py compute in
           fp32).  Measured ~7e-7 per layer; every gradient tensor within 3e-5 of float64 autograd on identical routing, 5
           routing decisions of 9.0 M differ from float64 (torch fp32 on the CPU: 4).  6x the MMA work;
  "split"  2 planes (hi / mid, three products, the same promotion): ~5e-6 per layer, gradients within 1e-4 except where the
           problem amplifies the forward error; still ~100x tighter than the TF32 the reference gets from cuDNN's default; 3x;
  "bf16"   single-plane bf16 operands: the fast mode, below the reference's precision (1e-2 class).
Internally every bf16 buffer carries a leading plane dimension P (1, 2 or 3).  The free parameters are
the p4 group-convolution filters psi (see oracle/equiv_ref.py for the restated architecture and why
weights are not interchangeable with e2cnn checkpoints).  No CPU path.

Host-side torch ops are used only for plumbing on tiny tensors: the expansion / projection of the
two 1x1 head filters (<= 262k elements) and buffer zeroing.
"""
from __future__ import annotations

import ctypes
import math
from typing import Dict, List, Optional

import torch

from . import _lib
from .kernels import _ptr, _stream, adv_moments, conv3x3_bf16, equiv_conv0, equiv_expand_regular, tc_gemm_bf16, tc_precision

ENC_FIELDS = [16, 32, 64, 128, 256, 128, 128]
PRECISIONS = {"bf16": 1, "split": 2, "fp32": 3}          # operand planes (aur_tc_set_precision)
N_ACT = 5


def _chk(rc, what):
    _lib.check(rc, what)


def init_params(seed: int = 0, device="cuda", scale: float = 1.0) -> Dict[str, torch.Tensor]:
    """He-style init of psi for one actor and one critic (separate encoders), fp32 on `device`."""
    g = torch.Generator().manual_seed(seed)
    p: Dict[str, torch.Tensor] = {}
    for net in ("actor", "critic"):
        cin_f = None
        for l, fo in enumerate(ENC_FIELDS):
            if l == 0:
                p[f"{net}.enc{l}.psi"] = torch.randn(fo, 2, 3, 3, generator=g) * scale * math.sqrt(2.0 / 18)
            else:
                p[f"{net}.enc{l}.psi"] = torch.randn(fo, cin_f, 4, 3, 3, generator=g) * scale * math.sqrt(2.0 / (cin_f * 36))
            p[f"{net}.enc{l}.bias"] = 0.01 * torch.randn(fo, generator=g)
            cin_f = fo
    F = ENC_FIELDS[-1]
    p["actor.head.psi_irrep"] = torch.randn(F, 2, generator=g) * math.sqrt(1.0 / (F * 4))
    p["actor.head.psi_triv"] = torch.randn(2 * N_ACT - 2, F, generator=g) * math.sqrt(1.0 / (F * 4))
    p["actor.head.bias_triv"] = 0.01 * torch.randn(2 * N_ACT - 2, generator=g)
    p["critic.head1.psi"] = torch.randn(F, F, 4, 1, 1, generator=g) * math.sqrt(2.0 / (F * 4))
    p["critic.head1.bias"] = 0.01 * torch.randn(F, generator=g)
    p["critic.head2.w"] = torch.randn(1, F, generator=g) * math.sqrt(1.0 / F)
    p["critic.head2.bias"] = 0.01 * torch.randn(1, generator=g)
    return {k: v.to(device).contiguous() for k, v in p.items()}


class _Enc:
    """Activation / gradient buffers of one encoder for a fixed batch size; ch = channels of the six stored activations
    (NHWC bf16, halo where the next layer pads), feat = encoder output width."""

    def __init__(self, B: int, dev, ch=(64, 128, 256, 512, 1024, 512), feat: int = 512, planes: int = 1):
        bf = lambda *s: torch.zeros(planes, *s, dtype=torch.bfloat16, device=dev)
        u8 = lambda *s: torch.zeros(*s, dtype=torch.uint8, device=dev)
        self.a = [bf(B, 66, 66, ch[0]), bf(B, 34, 34, ch[1]), bf(B, 18, 18, ch[2]), bf(B, 10, 10, ch[3]), bf(B, 8, 8, ch[4]),
                  bf(B, 3, 3, ch[5])]
        self.arg = [u8(B, 64, 64, ch[0]), u8(B, 32, 32, ch[1]), u8(B, 16, 16, ch[2]), u8(B, 8, 8, ch[3]), None, u8(B, 3, 3, ch[5])]
        self.feat = bf(B, feat)


class EquivActorCritic:
    CH = (64, 128, 256, 512, 1024, 512)      # channels of the stored activations a[0..5]
    FEAT = 512                               # encoder output width
    D_HEAD = 651

    def __init__(self, params: Dict[str, torch.Tensor], batch: int, lr: float = 3e-4, eps: float = 1e-5,
                 betas=(0.9, 0.999), split: bool = False, precision: Optional[str] = None):
        if precision is None:
            precision = "split" if split else "bf16"
        if precision not in PRECISIONS:
            raise _lib.AurError(f"precision must be one of {sorted(PRECISIONS)}")
        self.precision = precision
        self.P = PRECISIONS[precision]
        self.split = self.P > 1
        if batch % 8:
            raise _lib.AurError("batch must be a multiple of 8 (16-byte rows for the TMA weight-gradient maps)")
        _lib.lib()
        self.p = params
        self.dev = next(iter(params.values())).device
        if self.dev.type != "cuda":
            raise _lib.AurError("EquivActorCritic needs CUDA parameters (no CPU fallback)")
        self.B = batch
        self.enc = {"actor": _Enc(batch, self.dev, self.CH, self.FEAT, self.P),
                    "critic": _Enc(batch, self.dev, self.CH, self.FEAT, self.P)}
        # ONE flat fp32 buffer each for parameters, gradients and the Adam moments, clipped group (`actor.*`, robot_ppo.py:401)
        # first: the clip norm is one launch and Adam two, instead of one per tensor.  The caller's tensors keep their identity -
        # their storage is re-pointed into the flat buffer (as actor_critic.flat_parameters does), so a module sharing them
        # (models.robot_actor_critic) sees every update.
        keys = [k for k in params if k.startswith("actor.")] + [k for k in params if not k.startswith("actor.")]
        total = sum(params[k].numel() for k in keys)
        self._n_clip = sum(params[k].numel() for k in keys if k.startswith("actor."))
        self._flat = {name: torch.zeros(total, device=self.dev) for name in ("g", "m1", "m2")}
        # (a second engine over the same tensors - another batch size - finds them already laid out and shares the buffer)
        first, o, laid_out = params[keys[0]], 0, True
        for k in keys:
            v = params[k]
            if v.dtype != torch.float32:
                raise _lib.AurError(f"parameter {k} must be float32")
            laid_out = laid_out and v.is_contiguous() and v.data_ptr() == first.data_ptr() + 4 * o and \
                v.untyped_storage().data_ptr() == first.untyped_storage().data_ptr()
            o += v.numel()
        if laid_out and first.untyped_storage().nbytes() >= 4 * (first.storage_offset() + total):
            self._flat["p"] = torch.empty(0, device=self.dev).set_(first.untyped_storage(), first.storage_offset(), (total,))
        else:
            self._flat["p"] = torch.zeros(total, device=self.dev)
            laid_out = False
        self.grads, self.m1, self.m2 = {}, {}, {}
        o = 0
        for k in keys:
            v, n = params[k], params[k].numel()
            if not laid_out:
                self._flat["p"][o:o + n].copy_(v.detach().reshape(-1))
                v.data = self._flat["p"][o:o + n].view(v.shape)
            self.grads[k] = self._flat["g"][o:o + n].view(v.shape)
            self.m1[k] = self._flat["m1"][o:o + n].view(v.shape)
            self.m2[k] = self._flat["m2"][o:o + n].view(v.shape)
            o += n
        self.lr, self.eps, self.betas, self.step_count = lr, eps, betas, 0
        self.stats = torch.zeros(8, device=self.dev)
        self.d_head = torch.zeros(self.D_HEAD, device=self.dev)
        self.moments = torch.zeros(3, dtype=torch.float64, device=self.dev)
        self.sumsq = torch.zeros(1, dtype=torch.float64, device=self.dev)
        self.ws = torch.zeros(1 << 20, device=self.dev)
        desc = _lib.PolicyDesc(4, 2, 64, 2, 0)                  # only sizes the advantage-moment workspace
        self._mom_ws = torch.zeros((int(_lib.lib().aur_ppo_update_workspace_bytes(ctypes.byref(desc))) + 3) // 4, device=self.dev)
        self._cs = (torch.tensor([1.0, 0.0, -1.0, 0.0], device=self.dev), torch.tensor([0.0, 1.0, 0.0, -1.0], device=self.dev))
        self._a_bias = torch.zeros(10, device=self.dev)
        self.value = torch.zeros(batch, device=self.dev)
        self.logp = torch.zeros(batch, device=self.dev)
        self._idx11 = None
        self._w = {}
        self._zcache = {}

    # ------------------------------------------------------------------- planes
    def _bf(self, x: torch.Tensor) -> torch.Tensor:
        """fp32 -> [P, ...] bf16 planes (tiny host-side tensors only: head filters)."""
        from .kernels import split_planes
        return split_planes(x, self.P)

    def _empty(self, *shape) -> torch.Tensor:
        return torch.empty(self.P, *shape, dtype=torch.bfloat16, device=self.dev)

    # ------------------------------------------------------------------ weights
    def _expand(self):
        """psi -> bf16 contraction matrices (forward, backward-data) and per-channel biases."""
        w = {}
        for net in ("actor", "critic"):
            for l in range(1, 6):
                wm, wt, b = equiv_expand_regular(self.p[f"{net}.enc{l}.psi"], self.p[f"{net}.enc{l}.bias"], want_wt=True)
                w[f"{net}.{l}"] = (wm.reshape(self.P, *wm.shape[-3:]), wt.reshape(self.P, *wt.shape[-3:]), b)
            wm, _, b = equiv_expand_regular(self.p[f"{net}.enc6.psi"], self.p[f"{net}.enc6.bias"])
            wm = wm.reshape(self.P, 512, 4608)
            w[f"{net}.6"] = (wm, wm.transpose(1, 2).contiguous(), b)
        # heads (tiny, torch plumbing): actor [16,512] (10 used), critic head-1 [512,512]
        pi, pt = self.p["actor.head.psi_irrep"], self.p["actor.head.psi_triv"]
        c, s = self._cs
        a_, b_ = pi[:, 0:1], pi[:, 1:2]
        Wa = torch.zeros(16, 512, device=self.dev)
        Wa[0] = (c * a_ - s * b_).reshape(-1)
        Wa[1] = (s * a_ + c * b_).reshape(-1)
        Wa[2:10] = pt.unsqueeze(-1).expand(-1, -1, 4).reshape(8, 512)
        w["actor.head"] = (self._bf(Wa), self._bf(Wa.t().contiguous()))
        psi1 = self.p["critic.head1.psi"].reshape(128, 128, 4)
        if self._idx11 is None:
            o = torch.arange(128, device=self.dev).view(128, 1, 1, 1); r = torch.arange(4, device=self.dev).view(1, 4, 1, 1)
            i = torch.arange(128, device=self.dev).view(1, 1, 128, 1); s_ = torch.arange(4, device=self.dev).view(1, 1, 1, 4)
            self._idx11 = ((o * 128 + i) * 4 + ((s_ - r) % 4)).reshape(512, 512)
        W1 = psi1.reshape(-1)[self._idx11]
        w["critic.head1"] = (self._bf(W1), self._bf(W1.t().contiguous()),
                             self.p["critic.head1.bias"].repeat_interleave(4).contiguous())
        self._w = w

    # ------------------------------------------------------------------ forward
    def _layer0_params(self, net: str):
        return self.p[f"{net}.enc0.psi"], self.p[f"{net}.enc0.bias"]

    def _conv0(self, net: str, state, obs, e):
        psi0, bias0 = self._layer0_params(net)
        equiv_conv0(obs, state, psi0, bias0, e.a[0], e.arg[0])

    def _encoder_forward(self, net: str, state, obs):
        e, w = self.enc[net], self._w
        self._conv0(net, state, obs, e)
        for l, (epi, off) in zip(range(1, 6), [(2, 1), (2, 1), (2, 1), (1, 0), (2, 0)]):
            wm, _, b = w[f"{net}.{l}"]
            conv3x3_bf16(e.a[l - 1], wm, b, epi, e.a[l], off, e.arg[l])
        wm6, _, b6 = w[f"{net}.6"]
        pre = tc_gemm_bf16(e.a[5].reshape(self.P, self.B, 9 * self.CH[5]), wm6)
        L = _lib.lib()
        with torch.cuda.device(self.dev):
            _chk(L.aur_bias_relu_bf16(self.B, self.FEAT, pre.data_ptr(), b6.data_ptr(), e.feat.data_ptr(), _stream()), "aur_bias_relu_bf16")

    def forward(self, state: torch.Tensor, obs: torch.Tensor):
        """Encoders + head GEMMs; returns (actor head output [B,16] fp32, critic head-1 pre-activation [B,512] fp32)."""
        with tc_precision(self.P):
            return self._forward(state, obs)

    def _forward(self, state: torch.Tensor, obs: torch.Tensor):
        self._expand()
        self._encoder_forward("actor", state, obs)
        self._encoder_forward("critic", state, obs)
        a_out = tc_gemm_bf16(self.enc["actor"].feat, self._w["actor.head"][0])
        c_pre = tc_gemm_bf16(self.enc["critic"].feat, self._w["critic.head1"][0])
        return a_out, c_pre

    # ----------------------------------------------------------------- backward
    def _t(self, x: torch.Tensor) -> torch.Tensor:
        """[P,R,C] -> [P,C,R], plane by plane."""
        P, R, C = x.shape
        out = torch.empty(P, C, R, dtype=torch.bfloat16, device=self.dev)
        with torch.cuda.device(self.dev):
            for pl in range(P):
                _chk(_lib.lib().aur_transpose_bf16(R, C, x[pl].data_ptr(), out[pl].data_ptr(), _stream()), "aur_transpose_bf16")
        return out

    def _cast(self, g: torch.Tensor, ref: Optional[torch.Tensor]) -> torch.Tensor:
        """fp32 [..] -> [P, ..] bf16 planes, masked by ref > 0 (ref: [P, ..] planes, the hi plane decides)."""
        out = self._empty(*g.shape)
        with torch.cuda.device(self.dev):
            _chk(_lib.lib().aur_relu_mask_bf16(g.numel(), g.data_ptr(), _ptr(ref), out.data_ptr(), _stream()), "aur_relu_mask_bf16")
        return out

    def _wgrad(self, net: str, l: int, dy_buf: torch.Tensor, x_buf: torch.Tensor, base_off: int, bias_done: bool = False):
        """dpsi_l, dbias_l from the haloed output-gradient buffer and the layer's input buffer."""
        L = _lib.lib()
        B, Hb, Wb, Cin = x_buf.shape[-4:]
        Cout = dy_buf.shape[-1]
        Q = B * Hb * Wb
        dw = torch.zeros(Cout, 9, Cin, device=self.dev)
        with torch.cuda.device(self.dev):
            _chk(L.aur_wgrad3x3_bf16(Cout, Cin, Q, dy_buf.data_ptr(), x_buf.data_ptr(), base_off, Wb, dw.data_ptr(), 0, _stream()),
                 "aur_wgrad3x3_bf16")
        self._store_wgrad(net, l, dw, Cout, Cin)
        if not bias_done:
            self._store_bgrad(net, l, dy_buf.reshape(self.P, Q, Cout), Q, Cout)

    # dense gradient of a layer's contraction matrix [Cout, 9, Cin] (layer 6: [Cout, 9 * Cin]) -> the free parameters
    def _store_wgrad(self, net: str, l: int, dw: torch.Tensor, Cout: int, Cin: int):
        with torch.cuda.device(self.dev):
            _chk(_lib.lib().aur_equiv_project_regular(dw.data_ptr(), Cout // 4, Cin // 4, self.grads[f"{net}.enc{l}.psi"].data_ptr(),
                                                      _stream()), "aur_equiv_project_regular")

    def _store_bgrad(self, net: str, l: int, dy2d: torch.Tensor, Q: int, Cout: int):
        with torch.cuda.device(self.dev):
            for pl in range(self.P):                       # the kernel accumulates: hi + mid
                _chk(_lib.lib().aur_colsum_bf16(Q, Cout, dy2d[pl].data_ptr(), 4, self.grads[f"{net}.enc{l}.bias"].data_ptr(),
                                                _stream()), "aur_colsum_bf16")

    def _layer0_wgrad(self, net: str, state, obs, dprev, e):
        with torch.cuda.device(self.dev):
            _chk(_lib.lib().aur_equiv_conv0_wgrad(obs.data_ptr(), state.data_ptr(), dprev.data_ptr(), e.a[0].data_ptr(),
                                                  e.arg[0].data_ptr(), self.B, self.ws.data_ptr(),
                                                  self.grads[f"{net}.enc0.psi"].data_ptr(),
                                                  self.grads[f"{net}.enc0.bias"].data_ptr(), _stream()), "aur_equiv_conv0_wgrad")

    def _halo_zeros(self, tag: str, *shape) -> torch.Tensor:
        """bf16 buffer whose halo must be zero and whose interior is fully overwritten by the kernel that fills it (un-pool,
        masked backward-data): allocated and zeroed ONCE per shape - the halo is never written, so it stays zero, and a
        per-call torch.zeros of these buffers was 19 GB of fill traffic per 4096-sample update.  Shared by the two nets
        (their backward passes run one after the other on one stream)."""
        key = (tag,) + shape                        # one buffer per role: two live buffers may share a shape
        t = self._zcache.get(key)
        if t is None:
            t = torch.zeros(self.P, *shape, dtype=torch.bfloat16, device=self.dev)
            self._zcache[key] = t
        return t

    def _unpool(self, dpool, act, aoff, arg, C, Hp, dHb, doff, tag: str, bias=None):
        """bias = (net, l): also accumulate that layer's bias gradient from the values written (no second pass over dy)."""
        out = self._halo_zeros(tag, self.B, dHb, dHb, C)
        acc, group = self._bgrad_begin(bias[0], bias[1], C) if bias is not None else (None, 1)
        with torch.cuda.device(self.dev):
            _chk(_lib.lib().aur_unpool_relu_bwd_colsum(self.B, Hp, Hp, C, dpool.data_ptr(), act.data_ptr(), act.shape[-3], act.shape[-2],
                                                       aoff, arg.data_ptr(), out.data_ptr(), dHb, dHb, doff, group, _ptr(acc), _stream()),
                 "aur_unpool_relu_bwd_colsum")
        if bias is not None:
            self._bgrad_end(bias[0], bias[1], acc)
        return out

    # where a fused bias-gradient accumulation lands: (fp32 accumulator [C / group], group); _bgrad_end finalises it
    def _bgrad_begin(self, net: str, l: int, C: int):
        return self.grads[f"{net}.enc{l}.bias"], 4

    def _bgrad_end(self, net: str, l: int, acc: torch.Tensor):
        pass

    def _encoder_backward(self, net: str, state, obs, dfeat: torch.Tensor):
        """dfeat: fp32 [B, FEAT] gradient wrt the encoder output (post-ReLU features)."""
        e, w, B, CH = self.enc[net], self._w, self.B, self.CH
        dz6 = self._cast(dfeat, e.feat)                                         # through the last ReLU
        wm6, wm6t, _ = w[f"{net}.6"]
        # layer 6 (dense 3x3 -> 1x1): weight gradient [FEAT, 9 CH5] = dz6^T a6 ; data gradient = dz6 W6
        dz6_cm = self._t(dz6)
        dW6 = tc_gemm_bf16(dz6_cm, self._t(e.a[5].reshape(self.P, B, 9 * CH[5])))
        self._store_wgrad(net, 6, dW6, self.FEAT, CH[5])
        self._store_bgrad(net, 6, dz6, B, self.FEAT)
        da6 = self._cast(tc_gemm_bf16(dz6, wm6t), None).reshape(self.P, B, 3, 3, CH[5])
        # layer 5 (pad 0, pooled): un-pool into a 2-halo buffer (backward-data) and into the input geometry (weights)
        dy5_d = self._unpool(da6, e.a[5], 0, e.arg[5], CH[5], 3, 10, 2, "dy5_d")
        dy5_w = self._unpool(da6, e.a[5], 0, e.arg[5], CH[5], 3, 8, 0, "dy5_w", bias=(net, 5))
        self._wgrad(net, 5, dy5_w, e.a[4], 0, bias_done=True)
        dy4 = self._halo_zeros("dy4", B, 10, 10, CH[4])
        conv3x3_bf16(dy5_d, w[f"{net}.5"][1], None, 3, dy4, 1, None, relu_ref=e.a[4], ref_off=0)   # x ReLU mask of layer 4
        # layer 4 (pad 1, ReLU only)
        self._wgrad(net, 4, dy4, e.a[3], -(10 + 1))
        da4 = self._empty(B, 8, 8, CH[3])
        conv3x3_bf16(dy4, w[f"{net}.4"][1], None, 0, da4, 0)
        # layers 3, 2, 1 (pad 1, pooled)
        dprev = da4
        for l, Hp in ((3, 8), (2, 16), (1, 32)):
            C = CH[l]
            Hb = 2 * Hp + 2
            dy = self._unpool(dprev, e.a[l], 1, e.arg[l], C, Hp, Hb, 1, f"dy{l}", bias=(net, l))
            self._wgrad(net, l, dy, e.a[l - 1], -(Hb + 1), bias_done=True)
            Cin = e.a[l - 1].shape[-1]
            dprev = self._empty(B, 2 * Hp, 2 * Hp, Cin)
            conv3x3_bf16(dy, w[f"{net}.{l}"][1], None, 0, dprev, 0)
        # layer 0 (direct)
        self._layer0_wgrad(net, state, obs, dprev, e)

    # ------------------------------------------------------------------- update
    def loss_and_grads(self, state, obs, action, oldlp, adv, ret, vold, clip_coeff=0.2, entropy_coeff=0.01,
                       value_coeff=0.5, norm_adv=True, clip_vloss=True, m_total: Optional[int] = None) -> torch.Tensor:
        """Forward + loss + full backward; gradients land in self.grads.  Returns the stats tensor (means).
        m_total: size of the whole minibatch this batch is a part of (means divide by it; default: this batch)."""
        self._m_total = int(m_total) if m_total else self.B
        with tc_precision(self.P):
            return self._loss_and_grads(state, obs, action, oldlp, adv, ret, vold, clip_coeff, entropy_coeff, value_coeff,
                                        norm_adv, clip_vloss)

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
        d_c_h = self._empty(B, 512)
        a_bias = self._a_bias
        a_bias[2:10].copy_(self.p["actor.head.bias_triv"])
        h = _lib.EquivHeadArgs()
        h.B, h.clip_vloss, h.m_total = B, int(bool(clip_vloss)), getattr(self, "_m_total", B)
        h.a_out, h.a_bias, h.c_pre = a_out.data_ptr(), a_bias.data_ptr(), c_pre.data_ptr()
        h.c_bias1 = self._w["critic.head1"][2].data_ptr()
        w2 = self.p["critic.head2.w"].reshape(-1).contiguous()
        h.c_w2, h.c_b2 = w2.data_ptr(), self.p["critic.head2.bias"].data_ptr()
        h.action, h.oldlp, h.adv, h.ret, h.vold = (t.data_ptr() for t in (action, oldlp, adv, ret, vold))
        h.adv_moments = self.moments.data_ptr() if norm_adv else None
        h.clip_coeff, h.entropy_coeff, h.value_coeff = float(clip_coeff), float(entropy_coeff), float(value_coeff)
        h.d_a_out, h.d_c_h, h.d_head, h.stats = d_a_out.data_ptr(), d_c_h.data_ptr(), self.d_head.data_ptr(), self.stats.data_ptr()
        h.value_out, h.logp_out = self.value.data_ptr(), self.logp.data_ptr()
        with torch.cuda.device(self.dev):
            _chk(L.aur_equiv_head_loss(ctypes.byref(h), _stream()), "aur_equiv_head_loss")
        self._last_head = (a_out, c_pre, d_a_out, d_c_h)
        # ---- head parameter gradients (contractions on tensor cores, projection = tiny torch plumbing)
        fa, fc = self.enc["actor"].feat, self.enc["critic"].feat
        dWa = tc_gemm_bf16(self._t(d_a_out), self._t(fa))                       # [16,512]
        c, s = self._cs
        g0, g1 = dWa[0].reshape(128, 4), dWa[1].reshape(128, 4)
        self.grads["actor.head.psi_irrep"].copy_(torch.stack([(c * g0 + s * g1).sum(1), (-s * g0 + c * g1).sum(1)], 1))
        self.grads["actor.head.psi_triv"].copy_(dWa[2:10].reshape(8, 128, 4).sum(2))
        self.grads["actor.head.bias_triv"].copy_(self.d_head[2:10])
        dW1 = tc_gemm_bf16(self._t(d_c_h), self._t(fc))                         # [512,512]
        self.grads["critic.head1.psi"].reshape(-1).index_add_(0, self._idx11.reshape(-1), dW1.reshape(-1))
        self.grads["critic.head1.bias"].copy_(self.d_head[139:651].reshape(128, 4).sum(1))
        self.grads["critic.head2.w"].copy_(self.d_head[10:138].reshape(1, 128))
        self.grads["critic.head2.bias"].copy_(self.d_head[138:139])
        # ---- gradients wrt the encoder features, then the two encoders
        dfa = tc_gemm_bf16(d_a_out, self._w["actor.head"][1])                   # [B,512] fp32
        dfc = tc_gemm_bf16(d_c_h, self._w["critic.head1"][1])
        self._encoder_backward("actor", state, obs, dfa)
        self._encoder_backward("critic", state, obs, dfc)
        st = self.stats / B
        return st

    def apply(self, lr: Optional[float] = None, max_grad_norm: float = 0.5):
        """clip_grad_norm_ on the ACTOR parameters only (robot_ppo.py:401), then Adam over everything."""
        L = _lib.lib()
        self.step_count += 1
        lr = self.lr if lr is None else lr
        self.sumsq.zero_()
        F, nc = self._flat, self._n_clip
        nr = F["p"].numel() - nc
        with torch.cuda.device(self.dev):
            _chk(L.aur_sumsq_f32(nc, F["g"].data_ptr(), self.sumsq.data_ptr(), _stream()), "aur_sumsq_f32")
            _chk(L.aur_adam_flat(nc, F["p"].data_ptr(), F["g"].data_ptr(), F["m1"].data_ptr(), F["m2"].data_ptr(), lr, self.betas[0],
                                 self.betas[1], self.eps, self.step_count, self.sumsq.data_ptr(), max_grad_norm, _stream()), "aur_adam_flat")
            if nr:
                _chk(L.aur_adam_flat(nr, F["p"].data_ptr() + 4 * nc, F["g"].data_ptr() + 4 * nc, F["m1"].data_ptr() + 4 * nc,
                                     F["m2"].data_ptr() + 4 * nc, lr, self.betas[0], self.betas[1], self.eps, self.step_count, None,
                                     max_grad_norm, _stream()), "aur_adam_flat")

    def update(self, state, obs, action, oldlp, adv, ret, vold, lr=None, max_grad_norm=0.5, **kw) -> torch.Tensor:
        st = self.loss_and_grads(state, obs, action, oldlp, adv, ret, vold, **kw)
        self.apply(lr, max_grad_norm)
        return st
