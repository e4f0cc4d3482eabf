"""End-to-end training runs through the reference's CLI defaults (run_ppo.py:14-51) on the device path: CartPole-v1 (discrete
defaults: 4 envs, 500k steps) and Pendulum-v1 (the continuous override: 1 env, T = 2048, 2M steps, 32 minibatches x 10 epochs).
Writes the moving-average return curves to gpurun_out/ for profiles/r2_training_curves.md."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from aur_ppo_b200 import run_ppo
from aur_ppo_b200.ppo import ppo

out = {}
for name, argv in (("CartPole-v1", ["--gym_id", "CartPole-v1"]), ("Pendulum-v1", ["--gym_id", "Pendulum-v1", "--continuous", "True"]),
                   ("CartPole-v1 4096 envs", ["--gym_id", "CartPole-v1", "--num_envs", "4096", "--total_timesteps", "20971520"])):
    params = run_ppo.params_from_args(run_ppo.build_parser().parse_args(argv + ["--no_tensorboard", "--no_save"]))
    t0 = time.time()
    agent = ppo(params)
    rets, lens, xs = agent.train()
    dt = time.time() - t0
    rets = np.asarray(rets, dtype=np.float64)
    k = max(len(rets) // 10, 1)
    curve = [float(rets[i:i + k].mean()) for i in range(0, len(rets) - k + 1, k)][:10]
    out[name] = {"num_envs": params["num_envs"], "num_steps": params["num_steps"], "total_timesteps": params["total_timesteps"],
                 "logged_episodes": int(len(rets)), "mean_return_by_tenth_of_training": curve, "seconds": round(dt, 1),
                 "env_steps_per_s_wall": round(params["total_timesteps"] / dt), "last_stats": agent.last_stats}
    print(name, out[name], flush=True)
os.makedirs("gpurun_out/r2s", exist_ok=True)
json.dump(out, open("gpurun_out/r2s/training_curves.json", "w"), indent=1)
