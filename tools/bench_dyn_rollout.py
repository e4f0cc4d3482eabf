"""Times aur_rollout for policy widths other than 64 (the runtime-width path) and for the extra env ids.
Usage: PYTHONPATH=. python tools/bench_dyn_rollout.py   -> one JSON line per case."""
import json

import torch

from aur_ppo_b200 import envs as denv, kernels

CASES = [("CartPole-v1", 4, 2, 64, 2, False), ("CartPole-v1", 4, 2, 32, 2, False), ("CartPole-v1", 4, 2, 128, 2, False),
         ("CartPole-v1", 4, 2, 256, 2, False), ("CartPole-v1", 4, 2, 64, 4, False), ("Acrobot-v1", 6, 3, 64, 2, False),
         ("MountainCar-v0", 2, 3, 64, 2, False), ("MountainCarContinuous-v0", 2, 1, 64, 2, True), ("Pendulum-v1", 3, 1, 64, 2, True)]
N, T = 65536, 128
for gym_id, obs_dim, act_dim, H, NL, cont in CASES:
    desc = kernels.policy_desc(obs_dim, act_dim, H, NL, cont)
    P = kernels.policy_param_count(desc)
    g = torch.Generator(device="cuda").manual_seed(0)
    flat = (torch.rand(P, device="cuda", generator=g) - 0.5) * (2.0 / H ** 0.5)
    env = denv.DeviceVecEnv(gym_id, N, wrappers=cont)
    env.reset(list(range(N)))
    buf = kernels.RolloutBuffers(T, N, obs_dim, (act_dim,) if cont else (), "cuda")
    for _ in range(2):
        kernels.rollout(env, desc, flat, buf, seed=1, step0=0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 3
    for i in range(n):
        kernels.rollout(env, desc, flat, buf, seed=1, step0=(i + 2) * T)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(json.dumps({"gym_id": gym_id, "hidden": H, "layers": NL, "num_envs": N, "num_steps": T, "ms": ms,
                      "env_steps_per_s": N * T / ms * 1e3}))
