"""Times aur_rollout (actor + env steps + the batched critic pass) at hidden 128 / 256 beside the 64-wide headline shape, tensor-core
kernels vs the runtime-width SIMT kernel.  Usage: python tools/bench_wide_rollout.py   -> one JSON line per case."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from aur_ppo_b200 import _lib, envs as denv, kernels

CASES = [("CartPole-v1", 4, 2, 64, False, 1), ("CartPole-v1", 4, 2, 128, False, 1), ("CartPole-v1", 4, 2, 128, False, 0),
         ("Pendulum-v1", 3, 1, 128, True, 1), ("Pendulum-v1", 3, 1, 128, True, 0), ("CartPole-v1", 4, 2, 256, False, 1),
         ("Pendulum-v1", 3, 1, 256, True, 1)]
if os.environ.get("SKIP_SIMT"):
    CASES = [c for c in CASES if c[5] == 1]
N, T = 65536, 128
L = _lib.lib()
for gym_id, obs_dim, act_dim, H, cont, impl in CASES:
    L.aur_rollout_set_impl(impl)
    desc = kernels.policy_desc(obs_dim, act_dim, H, 2, cont)
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
    print(json.dumps({"gym_id": gym_id, "hidden": H, "layers": 2, "kernels": "tcgen05" if impl else "simt", "num_envs": N,
                      "num_steps": T, "ms": round(ms, 3), "env_steps_per_s": N * T / ms * 1e3}), flush=True)
L.aur_rollout_set_impl(1)
