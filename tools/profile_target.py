"""Short program for ncu: each hot-path kernel a few times at BASELINE config-B shapes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from aur_ppo_b200 import envs as denv, kernels
from tools.microbench import _policy

which = sys.argv[1:] or ["rollout", "gae", "update"]
N, T = 65536, 128
desc, flat = _policy(False)
env = denv.DeviceVecEnv("CartPole-v1", N)
env.reset(list(range(N)))
buf = kernels.RolloutBuffers(T, N, 4, (), "cuda")
for i in range(3 if "rollout" in which else 1):
    kernels.rollout(env, desc, flat, buf, seed=1, step0=i * T)
out = (torch.empty_like(buf.rewards), torch.empty_like(buf.rewards))
if "gae" in which:
    for i in range(4):
        kernels.gae(buf.rewards, buf.values, buf.terminals, buf.next_value, env.next_done, 0.99, 0.95, True, out)
if "update" in which:
    ret, adv = kernels.gae(buf.rewards, buf.values, buf.terminals, buf.next_value, env.next_done, 0.99, 0.95, True, out)
    up = kernels.Updater(desc, flat.clone())
    idx = torch.randperm(N * T, device="cuda")[: N * T // 4].to(torch.int32)
    flat_bufs = (buf.states.reshape(-1, 4), buf.actions.reshape(-1), buf.log_probs.reshape(-1), adv.reshape(-1),
                 ret.reshape(-1), buf.values.reshape(-1))
    rec = kernels.pack_records(*flat_bufs)
    idx_all = torch.stack([torch.randperm(N * T, device="cuda")[: N * T // 4].to(torch.int32) for _ in range(4)]).contiguous()
    up.prepare_moments(flat_bufs[3], idx_all)
    for i in range(3):
        up.step(*flat_bufs, idx_all[i], lr=2.5e-4, records=rec, moments_index=i)
torch.cuda.synchronize()
print("profile target done")
