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
