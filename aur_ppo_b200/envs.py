"""Device-resident vector env: the host-side holder of the state the rollout kernel steps.

Mirrors how the reference drives gym.vector.SyncVectorEnv (src/ppo.py:66-68,110,188):
`reset(seed=list)` then T fused steps per `rollout()` call.  Env i is seeded like
gym/utils/seeding.py does -- PCG64(SeedSequence(seed_i)) -- with the 128-bit stream state
expanded on the host by NumPy and advanced on the device afterwards.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .kernels import _stream, _ptr

CARTPOLE, PENDULUM, MOUNTAINCAR, ACROBOT, MOUNTAINCAR_CONT = 0, 1, 2, 3, 4
ENV_IDS = {"CartPole-v1": CARTPOLE, "Pendulum-v1": PENDULUM, "MountainCar-v0": MOUNTAINCAR, "Acrobot-v1": ACROBOT,
           "MountainCarContinuous-v0": MOUNTAINCAR_CONT}
OBS_DIM = {CARTPOLE: 4, PENDULUM: 3, MOUNTAINCAR: 2, ACROBOT: 6, MOUNTAINCAR_CONT: 2}
PHYS_DIM = {CARTPOLE: 4, PENDULUM: 2, MOUNTAINCAR: 2, ACROBOT: 4, MOUNTAINCAR_CONT: 2}
CONTINUOUS = {PENDULUM, MOUNTAINCAR_CONT}


def pcg64_seed_states(seeds: Sequence[int]) -> np.ndarray:
    """[4,n] uint64 (state_hi, state_lo, inc_hi, inc_lo) of np.random.PCG64(SeedSequence(seed))."""
    out = np.empty((4, len(seeds)), dtype=np.uint64)
    m = (1 << 64) - 1
    for i, s in enumerate(seeds):
        st = np.random.PCG64(np.random.SeedSequence(int(s))).state["state"]
        out[0, i], out[1, i] = st["state"] >> 64, st["state"] & m
        out[2, i], out[3, i] = st["inc"] >> 64, st["inc"] & m
    return out


class DeviceVecEnv:
    def __init__(self, gym_id: str, num_envs: int, wrappers: bool = False, device="cuda", env_id0: int = 0,
                 log_capacity: Optional[int] = None, gamma: float = 0.99):
        if gym_id not in ENV_IDS:
            raise _lib.AurError(f"gym_id {gym_id!r} has no device kernel (compiled: {sorted(ENV_IDS)}); no CPU fallback")
        self.gym_id, self.kind, self.num_envs = gym_id, ENV_IDS[gym_id], int(num_envs)
        self.wrappers = bool(wrappers) and self.kind in CONTINUOUS
        self.device = torch.device(device)
        self.env_id0 = int(env_id0)
        self.obs_dim = OBS_DIM[self.kind]
        self.gamma = float(gamma)
        n = self.num_envs
        dev = self.device
        self.phys = torch.zeros(PHYS_DIM[self.kind], n, dtype=torch.float64, device=dev)
        self.pcg = torch.zeros(4, n, dtype=torch.int64, device=dev)
        self.elapsed = torch.zeros(n, dtype=torch.int32, device=dev)
        self.ep_return = torch.zeros(n, dtype=torch.float32, device=dev)
        self.ep_length = torch.zeros(n, dtype=torch.int32, device=dev)
        self.norm = torch.zeros(2 * OBS_DIM[self.kind] + 5, n, dtype=torch.float64, device=dev) if self.wrappers else None
        self.next_obs = torch.zeros(n, self.obs_dim, dtype=torch.float32, device=dev)
        self.next_done = torch.zeros(n, dtype=torch.float32, device=dev)
        # full episode log: opt-in (tests / small runs); the training loop reads `first_finished`
        cap = int(log_capacity) if log_capacity is not None else 0
        self.log_entries = torch.zeros(max(cap, 1), 4, dtype=torch.int32, device=dev)
        self.log_count = torch.zeros(1, dtype=torch.int32, device=dev)
        self.log_capacity = cap
        self.first_finished = None        # [T] uint64 keys, sized at the first rollout
        self.totals = torch.zeros(3, dtype=torch.float64, device=dev)

    # ------------------------------------------------------------------ C structs
    def state_struct(self) -> _lib.EnvState:
        return _lib.EnvState(self.phys.data_ptr(), self.pcg.data_ptr(), self.elapsed.data_ptr(),
                             self.ep_return.data_ptr(), self.ep_length.data_ptr(), _ptr(self.norm))

    def log_struct(self, T: int) -> _lib.EpisodeLog:
        if self.first_finished is None or self.first_finished.numel() != T:
            self.first_finished = torch.empty(T, dtype=torch.int64, device=self.device)
        self.first_finished.fill_(-1)          # all ones
        return _lib.EpisodeLog(self.log_entries.data_ptr() if self.log_capacity else None,
                               self.log_count.data_ptr() if self.log_capacity else None,   # full log (tests): a returning atomic
                               self.log_capacity, 0, self.first_finished.data_ptr(), self.totals.data_ptr())

    def first_finished_episodes(self):
        """Per step of the LAST rollout, the first finished env in env order (what ppo.py:114-122 logs):
        numpy arrays (t, local_env, return, length), one D2H copy of T*8 bytes."""
        keys = self.first_finished.cpu().numpy().view(np.uint64)
        t = np.nonzero(keys != np.uint64(0xFFFFFFFFFFFFFFFF))[0]
        k = keys[t]
        env = (k >> np.uint64(42)).astype(np.int64)
        length = ((k >> np.uint64(32)) & np.uint64(1023)).astype(np.int64)
        ret = (k & np.uint64(0xFFFFFFFF)).astype(np.uint32).view(np.float32)
        return t, env, ret, length

    # ---------------------------------------------------------------------- reset
    def reset(self, seed: Sequence[int]):
        """envs.reset(seed=[...]) (src/ppo.py:188) -> next_obs [N,obs] fp32 on device."""
        if len(seed) != self.num_envs:
            raise _lib.AurError("need one seed per env")
        st = pcg64_seed_states(seed)
        self.pcg.copy_(torch.from_numpy(st.view(np.int64)))
        with torch.cuda.device(self.device):
            es = self.state_struct()
            rc = _lib.lib().aur_env_reset(self.kind, self.num_envs, int(self.wrappers), ctypes.byref(es),
                                          self.next_obs.data_ptr(), self.next_done.data_ptr(), _stream())
        _lib.check(rc, "aur_env_reset")
        self.log_count.zero_()
        return self.next_obs, {}

    def drain_episodes(self) -> List[tuple]:
        """Finished episodes since the last drain as (step, env, return, length), ordered by (step, env)."""
        cnt = int(self.log_count.item())
        kept = min(cnt, self.log_capacity)
        rows = self.log_entries[:kept].cpu().numpy()
        self.log_count.zero_()
        out = [(int(r[0]), int(r[1]), float(np.int32(r[2]).view(np.float32)), int(r[3])) for r in rows]
        out.sort(key=lambda r: (r[0], r[1]))
        self.dropped_episodes = cnt - kept
        return out

    def close(self):
        pass
