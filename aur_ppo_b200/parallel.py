"""Data-parallel plan of the PPO hot path: env columns are sharded over
ranks, rollout and GAE need no communication, and each minibatch exchanges one packed fp32 buffer
[P gradient sums | 16 statistic sums] plus three fp64 advantage moments.  Gradient seeds carry
1/m_total, so the sum over ranks is the mean over the whole minibatch.

On the GPU box the exchange is done by the update kernels themselves over NVLink / NVSwitch peer
memory (`PeerExchange`: CUDA-IPC-mapped exchange areas, push + flag, gather inside the Adam kernel);
torch.distributed (NCCL) only carries the set-up: parameter broadcast, IPC handles, barriers.
`make_allreduce` (a library SUM all-reduce between the kernels) is kept for `exchange="nccl"` and for
the gloo CPU tests of the host logic."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional

import torch


@dataclass(frozen=True)
class ShardPlan:
    world_size: int
    rank: int
    num_envs: int          # global (the reference's --num_envs)
    num_steps: int
    num_minibatches: int

    def __post_init__(self):
        if self.num_envs % self.world_size:
            raise ValueError(f"num_envs={self.num_envs} must divide over world_size={self.world_size}")
        if (self.local_envs * self.num_steps) % self.num_minibatches:
            raise ValueError("local batch must divide into num_minibatches")

    @property
    def local_envs(self) -> int:
        return self.num_envs // self.world_size

    @property
    def env_id0(self) -> int:
        """First GLOBAL env id this rank owns: seeds the env (gym seed = env id, ppo.py:188) and keys Philox."""
        return self.rank * self.local_envs

    @property
    def env_ids(self) -> range:
        return range(self.env_id0, self.env_id0 + self.local_envs)

    @property
    def batch_size(self) -> int:
        return self.num_envs * self.num_steps

    @property
    def local_batch(self) -> int:
        return self.local_envs * self.num_steps

    @property
    def minibatch_size(self) -> int:
        return self.batch_size // self.num_minibatches

    @property
    def local_minibatch(self) -> int:
        return self.minibatch_size // self.world_size


def current_plan(num_envs: int, num_steps: int, num_minibatches: int) -> ShardPlan:
    ws, rk = 1, 0
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        ws, rk = torch.distributed.get_world_size(), torch.distributed.get_rank()
    return ShardPlan(ws, rk, num_envs, num_steps, num_minibatches)


def make_allreduce(plan: ShardPlan, group=None) -> Optional[Callable[[torch.Tensor], None]]:
    """SUM all-reduce used on the advantage moments and on the packed [grads | stats] buffer."""
    if plan.world_size == 1:
        return None

    def _reduce(t: torch.Tensor) -> None:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM, group=group)
    return _reduce


def broadcast_parameters(module: torch.nn.Module, plan: ShardPlan, src: int = 0) -> None:
    """Identical initial weights on every rank (after that, identical updates keep them in sync)."""
    if plan.world_size > 1:
        for p in module.parameters():
            torch.distributed.broadcast(p.data, src)


class PeerExchange:
    """Exchange areas of the in-kernel gradient all-reduce: one cudaMalloc'ed area per rank, mapped into every
    rank of the node with CUDA IPC (handles travel through torch.distributed)."""

    def __init__(self, plan: ShardPlan, desc, group=None):
        import ctypes
        from . import _lib
        if plan.world_size > _lib.DP_MAX_RANKS:
            raise _lib.AurError(f"PeerExchange supports up to {_lib.DP_MAX_RANKS} ranks of one node")
        L = _lib.lib()
        self.plan, self._L, self._opened = plan, L, []
        self.bytes = int(L.aur_dp_area_bytes(ctypes.byref(desc)))
        own, handle = ctypes.c_void_p(), ctypes.create_string_buffer(_lib.DP_HANDLE_BYTES)
        _lib.check(L.aur_dp_alloc(self.bytes, ctypes.byref(own), handle), "aur_dp_alloc")
        self.own = own.value
        handles = [None] * plan.world_size
        torch.distributed.all_gather_object(handles, bytes(handle.raw), group=group)
        self.ctx = _lib.DpCtx()
        self.ctx.world, self.ctx.rank = plan.world_size, plan.rank
        for r, h in enumerate(handles):
            if r == plan.rank:
                self.ctx.peer[r] = self.own
            else:
                p = ctypes.c_void_p()
                _lib.check(L.aur_dp_open(ctypes.create_string_buffer(h, _lib.DP_HANDLE_BYTES), ctypes.byref(p)), "aur_dp_open")
                self.ctx.peer[r] = p.value
                self._opened.append(p.value)
        self.seq = 0
        torch.distributed.barrier(group=group)

    def next_seq(self) -> int:
        self.seq += 1
        return self.seq

    def status(self) -> int:
        return int(self._L.aur_dp_status(self.own, None))

    def wait_stats(self, reset: bool = False) -> dict:
        """Time this rank's kernels spent spinning on peers' flags (measured on the device, globaltimer) since the last reset."""
        import ctypes
        from . import _lib
        out = (ctypes.c_uint64 * 6)()
        _lib.check(self._L.aur_dp_wait_stats(self.own, out, int(reset), None), "aur_dp_wait_stats")
        return {"grad_spin_us_sum": out[0] / 1e3, "grad_spins": int(out[1]), "moment_spin_us_sum": out[2] / 1e3,
                "moment_spins": int(out[3]), "adam_wall_wait_us": out[4] / 1e3, "adam_launches": int(out[5])}

    def close(self) -> None:
        if self._L is None:
            return
        torch.cuda.synchronize()
        if torch.distributed.is_initialized():
            torch.distributed.barrier()
        for p in self._opened:
            self._L.aur_dp_close(p)
        self._L.aur_dp_free(self.own)
        self._opened, self._L = [], None
