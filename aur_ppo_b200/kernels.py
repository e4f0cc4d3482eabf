"""Tensor-level wrappers over the C ABI.  Every function requires CUDA tensors
and raises if the extension is missing: there is no CPU path."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.AurError(f"{name} must be a CUDA tensor (the hot path has no CPU fallback)")
    if t.dtype != torch.float32:
        raise _lib.AurError(f"{name} must be float32, got {t.dtype}")
    if not t.is_contiguous():
        raise _lib.AurError(f"{name} must be contiguous")
    return t


def gae(rewards: torch.Tensor, values: torch.Tensor, terminals: torch.Tensor, next_value: torch.Tensor,
        next_done: torch.Tensor, gamma: float, gae_lambda: float, use_gae: bool = True,
        out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """(returns, advantages) exactly as ppo.run_gae / ppo.normal_advantage return them
    (src/ppo.py:125-157); inputs are the [T,N] fp32 rollout buffers."""
    rewards, values, terminals = _f32c(rewards, "rewards"), _f32c(values, "values"), _f32c(terminals, "terminals")
    next_value, next_done = _f32c(next_value, "next_value"), _f32c(next_done, "next_done")
    if rewards.dim() != 2 or values.shape != rewards.shape or terminals.shape != rewards.shape:
        raise _lib.AurError("rewards/values/terminals must all be [T,N]")
    T, N = rewards.shape
    if next_value.numel() != N or next_done.numel() != N:
        raise _lib.AurError("next_value/next_done must have N elements")
    if out is None:
        ret, adv = torch.empty_like(rewards), torch.empty_like(rewards)
    else:
        ret, adv = out
        _f32c(ret, "returns"), _f32c(adv, "advantages")
    with torch.cuda.device(rewards.device):
        rc = _lib.lib().aur_gae_f32(T, N, rewards.data_ptr(), values.data_ptr(), terminals.data_ptr(),
                                    next_value.data_ptr(), next_done.data_ptr(), float(gamma), float(gae_lambda),
                                    int(bool(use_gae)), adv.data_ptr(), ret.data_ptr(), _stream())
    _lib.check(rc, "aur_gae_f32")
    return ret, adv


# ---------------------------------------------------------------------------- policy
def policy_desc(obs_dim: int, act_dim: int, hidden_dim: int, num_layers: int, continuous: bool) -> _lib.PolicyDesc:
    return _lib.PolicyDesc(int(obs_dim), int(act_dim), int(hidden_dim), int(num_layers), int(bool(continuous)))


def policy_param_count(desc: _lib.PolicyDesc) -> int:
    import ctypes
    return int(_lib.lib().aur_policy_param_count(ctypes.byref(desc)))


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def policy_evaluate(desc: _lib.PolicyDesc, params: torch.Tensor, obs: torch.Tensor,
                    actions: Optional[torch.Tensor] = None, seed: int = 0, row0: int = 0, step: int = 0, greedy: bool = False):
    """(action, log_prob, entropy, value[B]) of actor_critic.evaluate (models/actor_critic.py:34-51), no grad.
    greedy: arg-max action / Normal mean instead of a sample (checkpoint evaluation)."""
    import ctypes
    params, obs = _f32c(params, "params"), _f32c(obs, "obs")
    B = obs.shape[0]
    A = desc.act_dim if desc.continuous else 1
    if actions is not None:
        actions = _f32c(actions, "actions")
    act = torch.empty((B, A) if desc.continuous else (B,), device=obs.device, dtype=torch.float32)
    logp, ent, val = (torch.empty(B, device=obs.device, dtype=torch.float32) for _ in range(3))
    with torch.cuda.device(obs.device):
        rc = _lib.lib().aur_policy_act(ctypes.byref(desc), params.data_ptr(), B, obs.data_ptr(), _ptr(actions),
                                       int(bool(greedy)), seed, row0, step, act.data_ptr(), logp.data_ptr(), ent.data_ptr(),
                                       val.data_ptr(), _stream())
    _lib.check(rc, "aur_policy_act")
    return act, logp, ent, val


def sincos_f64(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    if not x.is_cuda or x.dtype != torch.float64 or not x.is_contiguous():
        raise _lib.AurError("x must be a contiguous CUDA float64 tensor")
    s, c = torch.empty_like(x), torch.empty_like(x)
    with torch.cuda.device(x.device):
        rc = _lib.lib().aur_sincos_f64(x.numel(), x.data_ptr(), s.data_ptr(), c.data_ptr(), _stream())
    _lib.check(rc, "aur_sincos_f64")
    return s, c


# --------------------------------------------------------------------------- rollout
class RolloutBuffers:
    """The reference's torch_buffer (src/ppo.py:20-39) on device, same shapes and dtypes."""

    def __init__(self, T: int, N: int, obs_dim: int, act_shape: tuple, device):
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=device)
        self.states = z(T, N, obs_dim)
        self.actions = z(T, N, *act_shape)
        self.log_probs, self.rewards, self.terminals, self.values = z(T, N), z(T, N), z(T, N), z(T, N)
        self.next_value = z(N)


def rollout(env, desc: _lib.PolicyDesc, params: torch.Tensor, buf: RolloutBuffers, seed: int, step0: int,
            actions_in: Optional[torch.Tensor] = None) -> None:
    """T fused steps (ppo.py:201-205): fills `buf`, advances `env`, leaves critic(next_obs) in buf.next_value."""
    import ctypes
    T, N = buf.rewards.shape
    a = _lib.RolloutArgs()
    a.env_kind, a.wrappers, a.N, a.T = env.kind, int(env.wrappers), N, T
    a.policy = desc
    a.params = _f32c(params, "params").data_ptr()
    a.env = env.state_struct()
    a.obs_buf, a.act_buf, a.logp_buf = buf.states.data_ptr(), buf.actions.data_ptr(), buf.log_probs.data_ptr()
    a.val_buf, a.rew_buf, a.done_buf = buf.values.data_ptr(), buf.rewards.data_ptr(), buf.terminals.data_ptr()
    a.next_obs, a.next_done, a.next_value = env.next_obs.data_ptr(), env.next_done.data_ptr(), buf.next_value.data_ptr()
    a.actions_in = None if actions_in is None else _f32c(actions_in, "actions_in").data_ptr()
    a.seed, a.step0, a.env_id0 = int(seed), int(step0), int(env.env_id0)
    a.log = env.log_struct(T)
    a.gamma = env.gamma
    with torch.cuda.device(params.device):
        rc = _lib.lib().aur_rollout(ctypes.byref(a), _stream())
    _lib.check(rc, "aur_rollout")


# ---------------------------------------------------------------------------- update
STAT_NAMES = ["policy_loss", "value_loss", "entropy", "old_approx_kl", "approx_kl", "clipfrac", "grad_norm", "loss"]
NUM_STATS = 16


def squashed_gaussian_sample(mean: torch.Tensor, log_std: torch.Tensor, action: Optional[torch.Tensor] = None, seed: int = 0,
                             stream_id: int = 0, return_pre_tanh: bool = False):
    """PPOGaussianPolicyBase.sample (src/nets/nets.py:90-105) -> (action, log_prob [B,1], tanh(mean), entropy [B,A])."""
    mean, log_std = _f32c(mean, "mean"), _f32c(log_std, "log_std")
    if mean.dim() != 2 or mean.shape != log_std.shape:
        raise _lib.AurError("squashed_gaussian_sample: mean and log_std must be [B,A]")
    if action is not None and _f32c(action, "action").shape != mean.shape:
        raise _lib.AurError("squashed_gaussian_sample: action must be [B,A]")
    B, A = mean.shape
    y, mo, ent = torch.empty_like(mean), torch.empty_like(mean), torch.empty_like(mean)
    lp = torch.empty(B, 1, device=mean.device)
    pre = torch.empty_like(mean) if return_pre_tanh else None
    with torch.cuda.device(mean.device):
        rc = _lib.lib().aur_squashed_gaussian_sample(B, A, mean.data_ptr(), log_std.data_ptr(), _ptr(action),
                                                     seed & 0xFFFFFFFFFFFFFFFF, stream_id & 0xFFFFFFFFFFFFFFFF, y.data_ptr(),
                                                     lp.data_ptr(), mo.data_ptr(), ent.data_ptr(), _ptr(pre), _stream())
    _lib.check(rc, "aur_squashed_gaussian_sample")
    return (y, lp, mo, ent, pre) if return_pre_tanh else (y, lp, mo, ent)


def shuffle_indices(n: int, seed: int, stream_id: int, device="cuda", out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`np.random.shuffle(arange(n))` of ppo.py:214-215 as a keyed bijection computed on the device -> int32 [n]."""
    if out is None:
        out = torch.empty(n, dtype=torch.int32, device=device)
    if out.dtype != torch.int32 or out.numel() != n or not out.is_contiguous() or not out.is_cuda:
        raise _lib.AurError("shuffle_indices: out must be a contiguous CUDA int32 tensor of n elements")
    _lib.check(_lib.lib().aur_shuffle_indices(n, seed & 0xFFFFFFFFFFFFFFFF, stream_id & 0xFFFFFFFFFFFFFFFF, out.data_ptr(),
                                              _stream()), "aur_shuffle_indices")
    return out


def pack_records(b_obs, b_actions, b_logprobs, b_advantages, b_returns, b_values, out=None):
    """Gather-friendly copy of the flattened batch: (rec_actor, rec_critic), each [B,8] fp32 (one 32-byte sector per
    sample).  Supports obs_dim <= 4 and one or two action columns; returns None otherwise (use the plain arrays)."""
    for n, t in (("obs", b_obs), ("actions", b_actions), ("logprobs", b_logprobs), ("advantages", b_advantages),
                 ("returns", b_returns), ("values", b_values)):
        _f32c(t, n)
    B, obs_dim = b_obs.shape[0], b_obs.shape[1]
    act_w = 1 if b_actions.dim() == 1 else b_actions.shape[1]
    if obs_dim > 4 or act_w > 2:
        return None
    if out is None:
        out = (torch.empty(B, 8, device=b_obs.device), torch.empty(B, 8, device=b_obs.device))
    with torch.cuda.device(b_obs.device):
        rc = _lib.lib().aur_ppo_pack_records(B, obs_dim, act_w, b_obs.data_ptr(), b_actions.data_ptr(), b_logprobs.data_ptr(),
                                             b_advantages.data_ptr(), b_returns.data_ptr(), b_values.data_ptr(),
                                             out[0].data_ptr(), out[1].data_ptr(), _stream())
    _lib.check(rc, "aur_ppo_pack_records")
    return out


class Updater:
    """Device state of the optimiser side of ppo.train (src/ppo.py:80,213-269): Adam moments,
    packed gradient buffer, workspace.  `allreduce` (optional) is called on the fp64 advantage
    moments and on the packed fp32 [grads | stats] buffer -- the only exchanges of a
    data-parallel run."""

    def __init__(self, desc: _lib.PolicyDesc, params: torch.Tensor, eps: float = 1e-5, betas=(0.9, 0.999),
                 allreduce=None, exchange=None):
        import ctypes
        self.desc, self.params = desc, _f32c(params, "params")
        dev = params.device
        self.P = policy_param_count(desc)
        if params.numel() != self.P:
            raise _lib.AurError(f"flat parameter buffer has {params.numel()} elements, policy needs {self.P}")
        self.exp_avg = torch.zeros(self.P, device=dev)
        self.exp_avg_sq = torch.zeros(self.P, device=dev)
        self.grads = torch.zeros(self.P + NUM_STATS, device=dev)
        self.moments = torch.zeros(3, dtype=torch.float64, device=dev)
        self.stats = torch.zeros(NUM_STATS, device=dev)
        ws = int(_lib.lib().aur_ppo_update_workspace_bytes(ctypes.byref(desc)))
        if ws < 0:
            _lib.check(ws, "aur_ppo_update_workspace_bytes")
        self.workspace = torch.zeros((ws + 3) // 4, dtype=torch.float32, device=dev)
        self.eps, self.betas, self.step_count, self.allreduce = float(eps), betas, 0, allreduce
        # exchange (parallel.PeerExchange): the kernels all-reduce over peer memory themselves; excludes `allreduce`
        self.exchange = exchange
        if exchange is not None and allreduce is not None:
            raise _lib.AurError("give either a peer exchange or an allreduce callable, not both")
        self.moments_all = None          # [n_mb, 3] fp64: moments of every minibatch of the iteration (prepare_moments)
        self._mom_seq = 0

    MAX_MINIBATCHES = 512                # AUR_DP_MAX_MINIBATCHES

    def prepare_moments(self, b_advantages: torch.Tensor, idx_all: torch.Tensor) -> None:
        """Advantage moments of ALL the minibatches of an iteration in one launch (and, data-parallel, ONE exchange):
        idx_all [n_mb, m] int32 = the index lists of every epoch's minibatches.  Afterwards `grad(..., moments_index=j)`
        normalises minibatch j with entry j instead of computing (and exchanging) its moments itself."""
        import ctypes
        if idx_all.dtype != torch.int32 or idx_all.dim() != 2 or not idx_all.is_cuda or not idx_all.is_contiguous():
            raise _lib.AurError("idx_all must be a contiguous CUDA int32 tensor [n_mb, m]")
        n_mb, m = idx_all.shape
        if n_mb > self.MAX_MINIBATCHES:
            raise _lib.AurError(f"at most {self.MAX_MINIBATCHES} minibatches per prepare_moments call")
        if self.moments_all is None or self.moments_all.shape[0] != n_mb:
            self.moments_all = torch.zeros(n_mb, 3, dtype=torch.float64, device=self.params.device)
        self._mom_seq += 1
        with torch.cuda.device(self.params.device):
            dp = ctypes.addressof(self.exchange.ctx) if self.exchange is not None else None
            rc = _lib.lib().aur_ppo_adv_moments_multi(n_mb, m, idx_all.data_ptr(), m, _f32c(b_advantages, "advantages").data_ptr(),
                                                      self.moments_all.data_ptr(), self.workspace.data_ptr(), dp, self._mom_seq, _stream())
            _lib.check(rc, "aur_ppo_adv_moments_multi")
            if self.allreduce is not None:
                self.allreduce(self.moments_all)

    def grad(self, b_obs, b_actions, b_logprobs, b_advantages, b_returns, b_values, idx: Optional[torch.Tensor],
             m_total: Optional[int] = None, idx_offset: int = 0, m_local: Optional[int] = None, clip_coeff=0.2,
             entropy_coeff=0.01, value_coeff=0.5, norm_adv=True, clip_vloss=True, records=None,
             moments_index: Optional[int] = None) -> torch.Tensor:
        """Phases 1-2 (moments, gather+fwd+loss+bwd+reduce) -> packed [grads | stat sums] on device.
        moments_index: use entry j of the last prepare_moments() call instead of computing this minibatch's moments."""
        import ctypes
        L = _lib.lib()
        if idx is not None:
            if idx.dtype != torch.int32 or not idx.is_cuda or not idx.is_contiguous():
                raise _lib.AurError("idx must be a contiguous CUDA int32 tensor")
            m_local = idx.numel()
        elif m_local is None:
            raise _lib.AurError("give idx or m_local")
        m_total = int(m_total if m_total is not None else m_local)
        for n, t in (("obs", b_obs), ("actions", b_actions), ("logprobs", b_logprobs), ("advantages", b_advantages),
                     ("returns", b_returns), ("values", b_values)):
            _f32c(t, n)
        with torch.cuda.device(self.params.device):
            st = _stream()
            dp, seq = (ctypes.addressof(self.exchange.ctx), self.exchange.next_seq()) if self.exchange is not None else (None, 0)
            self._seq = seq
            ahead = norm_adv and moments_index is not None
            if ahead and (self.moments_all is None or not 0 <= moments_index < self.moments_all.shape[0]):
                raise _lib.AurError("moments_index needs a matching prepare_moments() call")
            if norm_adv and not ahead:
                _lib.check(L.aur_ppo_adv_moments_dp(m_local, _ptr(idx), idx_offset, b_advantages.data_ptr(),
                                                    self.moments.data_ptr(), self.workspace.data_ptr(), dp, seq, st),
                           "aur_ppo_adv_moments")
                if self.allreduce is not None:
                    self.allreduce(self.moments)
            a = _lib.UpdateArgs()
            a.policy, a.norm_adv, a.clip_vloss = self.desc, int(bool(norm_adv)), int(bool(clip_vloss))
            a.m_local, a.m_total, a.idx, a.idx_offset = m_local, m_total, _ptr(idx), idx_offset
            a.obs, a.actions, a.logprobs = b_obs.data_ptr(), b_actions.data_ptr(), b_logprobs.data_ptr()
            a.advantages, a.returns, a.values = b_advantages.data_ptr(), b_returns.data_ptr(), b_values.data_ptr()
            a.params = self.params.data_ptr()
            a.clip_coeff, a.entropy_coeff, a.value_coeff = float(clip_coeff), float(entropy_coeff), float(value_coeff)
            a.adv_moments = self.moments.data_ptr() if norm_adv else None
            if ahead:
                a.adv_moments = self.moments_all.data_ptr() + 24 * moments_index
                if dp is not None:
                    a.mom_seq, a.mom_index = self._mom_seq, moments_index
            a.workspace, a.grads_out = self.workspace.data_ptr(), self.grads.data_ptr()
            a.dp, a.dp_seq = dp, seq
            if records is not None:
                a.rec_actor, a.rec_critic = records[0].data_ptr(), records[1].data_ptr()
            _lib.check(L.aur_ppo_update_grad(ctypes.byref(a), st), "aur_ppo_update_grad")
            if self.allreduce is not None:
                self.allreduce(self.grads)
        self._m_total, self._ent_c, self._vf_c = m_total, float(entropy_coeff), float(value_coeff)
        return self.grads

    def apply(self, lr: float, max_grad_norm: float = 0.5, stats_out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Phase 3: clip_grad_norm_ + Adam in place on the flat parameters; returns the stats tensor (device).
        stats_out: optional fp32 [NUM_STATS] row the kernel writes instead of self.stats (no extra copy kernel)."""
        import ctypes
        self.step_count += 1
        stats = self.stats if stats_out is None else _f32c(stats_out, "stats_out")
        if stats.numel() < NUM_STATS:
            raise _lib.AurError(f"stats_out needs {NUM_STATS} elements")
        with torch.cuda.device(self.params.device):
            dp = ctypes.addressof(self.exchange.ctx) if self.exchange is not None else None
            rc = _lib.lib().aur_ppo_update_apply_dp(ctypes.byref(self.desc), self.params.data_ptr(), self.grads.data_ptr(),
                                                    self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), float(lr),
                                                    self.betas[0], self.betas[1], self.eps, self.step_count,
                                                    float(max_grad_norm), self._m_total, self._ent_c, self._vf_c,
                                                    stats.data_ptr(), dp, getattr(self, "_seq", 0) if dp else 0, _stream())
        _lib.check(rc, "aur_ppo_update_apply")
        return stats

    def step(self, *args, lr: float, max_grad_norm: float = 0.5, **kw) -> torch.Tensor:
        self.grad(*args, **kw)
        return self.apply(lr, max_grad_norm)


# ----------------------------------------------------------------------- tensor cores
class tc_precision:
    """`with tc_precision(P):` runs the row-X entry points on P operand planes (aur_tc_set_precision): P = 2 / 3 makes
    every bf16 tensor a stack [P, ...] of hi / mid (/ lo) planes.  Restores the previous mode on exit."""

    def __init__(self, planes: int):
        self.planes, self.prev = int(planes), None

    def __enter__(self):
        self.prev = _lib.lib().aur_tc_set_precision(self.planes)
        if self.prev < 0:
            _lib.check(self.prev, "aur_tc_set_precision")
        return self

    def __exit__(self, *exc):
        _lib.lib().aur_tc_set_precision(self.prev)
        return False


def tc_planes() -> int:
    return int(_lib.lib().aur_tc_get_precision())


def split_planes(x: torch.Tensor, planes: int = 2) -> torch.Tensor:
    """fp32 tensor -> [planes, ...] bf16 (hi = bf16(x), mid = bf16(x - hi), lo = bf16(x - hi - mid)): the layout the
    multi-plane entry points take."""
    out, r = [], x.float()
    for _ in range(planes):
        h = r.bfloat16()
        out.append(h)
        r = r - h.float()
    return torch.stack(out).contiguous()


def join_planes(x: torch.Tensor) -> torch.Tensor:
    """[P, ...] bf16 planes -> fp32 value (sum of the planes)."""
    return x.float().sum(0)


def _check_planes(t: torch.Tensor, base_dims: int, name: str) -> int:
    """Plane count of a bf16 tensor argument: `base_dims` dims = one plane, one more leading dim = a plane stack."""
    if t.dtype != torch.bfloat16 or not t.is_cuda or not t.is_contiguous():
        raise _lib.AurError(f"{name} must be a contiguous CUDA bf16 tensor")
    P = 1 if t.dim() == base_dims else (t.shape[0] if t.dim() == base_dims + 1 else -1)
    if P != tc_planes():
        raise _lib.AurError(f"{name}: {tuple(t.shape)} does not match the current tensor-core precision ({tc_planes()} plane(s))")
    return P


def _planes_shape(P: int, *shape):
    return (P,) + tuple(shape) if P > 1 else tuple(shape)


def tc_gemm_bf16(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a [M,K] bf16, b [N,K] bf16 (or [2,M,K] / [2,N,K] plane stacks in split mode) -> a @ b.T as fp32 [M,N]
    (tcgen05, fp32 accumulate)."""
    _check_planes(a, 2, "a"), _check_planes(b, 2, "b")
    M, K = a.shape[-2:]
    N = b.shape[-2]
    c = torch.empty(M, N, dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        rc = _lib.lib().aur_tc_gemm_bf16(M, N, K, a.data_ptr(), b.data_ptr(), c.data_ptr(), _stream())
    _lib.check(rc, "aur_tc_gemm_bf16")
    return c


def equiv_expand_regular(psi: torch.Tensor, bias_f: Optional[torch.Tensor] = None, want_wt: bool = False):
    """psi [Fo,Fi,4,3,3] fp32 -> (wmat [Fo*4, 9, Fi*4] bf16, wt [Fi*4, 9, Fo*4] bf16 | None, bias [Fo*4] | None);
    split mode: wmat / wt are [2, ...] plane stacks."""
    psi = _f32c(psi, "psi")
    Fo, Fi = psi.shape[0], psi.shape[1]
    P = tc_planes()
    wmat = torch.empty(_planes_shape(P, Fo * 4, 9, Fi * 4), dtype=torch.bfloat16, device=psi.device)
    wt = torch.empty(_planes_shape(P, Fi * 4, 9, Fo * 4), dtype=torch.bfloat16, device=psi.device) if want_wt else None
    bias = torch.empty(Fo * 4, dtype=torch.float32, device=psi.device) if bias_f is not None else None
    with torch.cuda.device(psi.device):
        rc = _lib.lib().aur_equiv_expand_regular(psi.data_ptr(), Fo, Fi, _ptr(bias_f), wmat.data_ptr(), _ptr(wt),
                                                 _ptr(bias), _stream())
    _lib.check(rc, "aur_equiv_expand_regular")
    return wmat, wt, bias


def conv3x3_bf16(inp: torch.Tensor, wmat: torch.Tensor, bias: Optional[torch.Tensor], epilogue: int,
                 out: torch.Tensor, out_off: int, pool_arg: Optional[torch.Tensor] = None,
                 relu_ref: Optional[torch.Tensor] = None, ref_off: int = 0) -> torch.Tensor:
    """inp [B,Hb,Wb,Cin] bf16 (halo included) -> valid 3x3 conv (+bias/ReLU/pool) into out[:, off:, off:, :].
    Split mode: inp / wmat / out (and relu_ref) are [2, ...] plane stacks."""
    import ctypes
    _check_planes(inp, 4, "inp"), _check_planes(wmat, 3, "wmat"), _check_planes(out, 4, "out")
    B, Hb, Wb, Cin = inp.shape[-4:]
    Cout = wmat.shape[-3]
    a = _lib.ConvArgs(B, Hb, Wb, Cin, Cout, epilogue, out.shape[-3], out.shape[-2], out_off, 0, inp.data_ptr(),
                      wmat.data_ptr(), _ptr(bias), out.data_ptr(), _ptr(pool_arg), _ptr(relu_ref),
                      relu_ref.shape[-3] if relu_ref is not None else 0, relu_ref.shape[-2] if relu_ref is not None else 0,
                      ref_off, 0)
    with torch.cuda.device(inp.device):
        rc = _lib.lib().aur_conv3x3_bf16(ctypes.byref(a), _stream())
    _lib.check(rc, "aur_conv3x3_bf16")
    return out


def equiv_conv0(obs: torch.Tensor, state: torch.Tensor, psi: torch.Tensor, bias_f: torch.Tensor, out: torch.Tensor,
                pool_arg: Optional[torch.Tensor] = None) -> torch.Tensor:
    B = obs.shape[0]
    _check_planes(out, 4, "out")
    with torch.cuda.device(obs.device):
        rc = _lib.lib().aur_equiv_conv0(_f32c(obs, "obs").data_ptr(), _f32c(state, "state").data_ptr(),
                                        _f32c(psi, "psi").data_ptr(), _f32c(bias_f, "bias").data_ptr(), B, out.data_ptr(),
                                        _ptr(pool_arg), _stream())
    _lib.check(rc, "aur_equiv_conv0")
    return out


def adv_moments(adv: torch.Tensor, out: torch.Tensor, workspace: torch.Tensor) -> torch.Tensor:
    """(sum, sum of squares, count) of `adv` in fp64 -> out [3] (aur_ppo_adv_moments over the whole array)."""
    with torch.cuda.device(adv.device):
        rc = _lib.lib().aur_ppo_adv_moments(adv.numel(), None, 0, _f32c(adv, "adv").data_ptr(), out.data_ptr(),
                                            workspace.data_ptr(), _stream())
    _lib.check(rc, "aur_ppo_adv_moments")
    return out
