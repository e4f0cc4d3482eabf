"""Checkpoint consumer + evaluation loop: the reference's `test` class (src/test.py:17-61) on the device envs.

The reference loads a whole-module pickle (`torch.load('actor_critic.pt')`, test.py:22), plays `episodes`
episodes of one gym env with `agent.act(state)` (a SAMPLED action, test.py:51) and collects the episode
lengths (test.py:58).  Here the episodes run side by side: one device env per episode, every env plays its
first episode to the end (terminated, or truncated by the TimeLimit), stepped by the same CUDA env kernel
the training rollout uses (aur_rollout with the chosen actions), the policy evaluated by aur_policy_act.

Differences from test.py, all deliberate:
  * gym 0.26 API (reset(seed) -> (obs, info), five-value step) as src/ppo.py uses it; test.py itself still
    has the 0.21 calls (`state = self.env.reset()`, four-value step) and would not run against the pinned gym;
  * test.py:57 updates `state` AFTER its step loop (an indentation slip), so its agent keeps seeing the
    reset observation; here the policy sees the current observation every step;
  * `mode="greedy"` (arg-max / Normal mean) in addition to the reference's sampled play.
Legacy checkpoints (Linear layers at Sequential indices 0,3,6 with Dropout in between) and current ones
(0,2,4) both load (compat.load_policy).  No CPU path.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib, compat, kernels
from .envs import DeviceVecEnv, OBS_DIM, ENV_IDS, CONTINUOUS

_TIME_LIMIT = {"CartPole-v1": 500, "Pendulum-v1": 200, "MountainCar-v0": 200, "Acrobot-v1": 500,
               "MountainCarContinuous-v0": 999}


class test:
    def __init__(self, model, env: str = "CartPole-v1", render_mode=None, device="cuda"):
        if not torch.cuda.is_available():
            raise _lib.AurError("aur_ppo_b200.test needs a CUDA device: evaluation runs on the device envs (no CPU fallback)")
        _lib.lib()
        if env not in ENV_IDS:
            raise _lib.AurError(f"gym_id {env!r} has no device kernel (compiled: {sorted(ENV_IDS)})")
        self.gym_id = env
        self.device = torch.device(device)
        self.agent = (compat.load_policy(model, map_location="cpu") if isinstance(model, (str, bytes)) or hasattr(model, "read")
                      else model).to(self.device)
        shape = self.agent.kernel_shape()
        if shape[0] != OBS_DIM[ENV_IDS[env]]:
            raise _lib.AurError(f"checkpoint expects {shape[0]} observations, {env} has {OBS_DIM[ENV_IDS[env]]}")
        if bool(shape[4]) != (ENV_IDS[env] in CONTINUOUS):
            raise _lib.AurError(f"checkpoint is {'continuous' if shape[4] else 'discrete'}, {env} is not")
        self.desc = kernels.policy_desc(*shape)
        self.flat = self.agent.flat_parameters()
        self.episode_lengths: List[int] = []
        self.episode_returns: List[float] = []

    def moving_average(self, data, window_size):
        return np.convolve(data, np.ones(window_size) / window_size, mode="valid")

    def run(self, episodes: int, max_length: int = 10000, mode: str = "sampled", seed: int = 0,
            env_seeds: Optional[Sequence[int]] = None, wrappers: bool = False):
        """Plays `episodes` episodes (one per device env) -> list of episode lengths (test.py:58); the returns are
        kept in `self.episode_returns`.  mode: "sampled" (test.py:51) or "greedy"."""
        if mode not in ("sampled", "greedy"):
            raise _lib.AurError("mode must be 'sampled' or 'greedy'")
        N = int(episodes)
        env = DeviceVecEnv(self.gym_id, N, wrappers=wrappers, device=self.device, log_capacity=4 * N + 16)
        seeds = list(env_seeds) if env_seeds is not None else list(range(seed, seed + N))
        env.reset(seeds)
        cont = bool(self.desc.continuous)
        A = self.desc.act_dim
        buf = kernels.RolloutBuffers(1, N, env.obs_dim, (A,) if cont else (), self.device)
        horizon = min(int(max_length), _TIME_LIMIT[self.gym_id])
        length = np.zeros(N, np.int64)
        ret = np.zeros(N, np.float64)
        finished = np.zeros(N, bool)
        for t in range(horizon):
            act, _, _, _ = kernels.policy_evaluate(self.desc, self.flat, env.next_obs, seed=seed, row0=0, step=t,
                                                   greedy=(mode == "greedy"))
            kernels.rollout(env, self.desc, self.flat, buf, seed=seed, step0=t, actions_in=act.reshape(1, N, *buf.actions.shape[2:]))
            for (_, e, r, l) in env.drain_episodes():                  # RecordEpisodeStatistics entries of this step
                if not finished[e]:
                    finished[e], length[e], ret[e] = True, l, r
            if finished.all():
                break
        # episodes cut by max_length before the env ended them (test.py:50 `for _ in range(max_length)`)
        if not finished.all():
            open_len = env.ep_length.cpu().numpy()
            open_ret = env.ep_return.cpu().numpy()
            for e in np.nonzero(~finished)[0]:
                length[e], ret[e] = open_len[e], open_ret[e]
        self.episode_lengths = [int(x) for x in length]
        self.episode_returns = [float(x) for x in ret]
        return self.episode_lengths


def evaluate(path, gym_id: str = "CartPole-v1", episodes: int = 100, mode: str = "sampled", seed: int = 0, max_length: int = 10000):
    """Load a reference (or own) `actor_critic*.pt` and play it on the device envs -> dict(lengths, returns, mean_length, mean_return)."""
    t = test(path, gym_id)
    lengths = t.run(episodes, max_length=max_length, mode=mode, seed=seed)
    return dict(lengths=lengths, returns=t.episode_returns, mean_length=float(np.mean(lengths)),
                mean_return=float(np.mean(t.episode_returns)))


if __name__ == "__main__":          # src/test.py:60-61: test('actor_critic.pt', 'CartPole-v1', render_mode=None).run(1000)
    import argparse
    ap = argparse.ArgumentParser(description="play a saved actor_critic*.pt on the device envs (src/test.py)")
    ap.add_argument("model", nargs="?", default="actor_critic.pt")
    ap.add_argument("--gym_id", default="CartPole-v1")
    ap.add_argument("--episodes", type=int, default=1000)
    ap.add_argument("--mode", choices=["sampled", "greedy"], default="sampled")
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    out = evaluate(a.model, a.gym_id, a.episodes, a.mode, a.seed)
    print(f"{a.episodes} episodes of {a.gym_id} ({a.mode}): mean length {out['mean_length']:.1f}, mean return {out['mean_return']:.1f}, "
          f"min {min(out['lengths'])}, max {max(out['lengths'])}")
