"""The reference's CPU training path restated (oracle/gym_restated.py + oracle/ppo_ref.py: per-env Python gym objects,
torch-CPU actor_critic math, run_gae loop, minibatch loop with clip_grad_norm_ + Adam and the lr anneal of ppo.py:195-198)
run to the END of training on the reference's CLI defaults, for the learning curve that tools/train_curves.py measures on the
device.  The two arms draw different random numbers (torch CPU generator vs Philox), so the curves agree statistically, not
step by step; the step-by-step agreement is what tests/ checks.  Test infrastructure: imports oracle/."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import gym_restated as G
from oracle import ppo_ref as R
from aur_ppo_b200.models.actor_critic import actor_critic

CFG = {"CartPole-v1": dict(N=4, T=128, total=500000, nm=4, epochs=4, lr=2.5e-4, ent=0.01, cont=False, O=4, A=2),
       "Pendulum-v1": dict(N=1, T=2048, total=2000000, nm=32, epochs=10, lr=3e-4, ent=0.0, cont=True, O=3, A=1)}


def run(gym_id, seed=1, total=None):
    c = CFG[gym_id]
    N, T, O, cont = c["N"], c["T"], c["O"], c["cont"]
    total = total or c["total"]
    torch.manual_seed(seed)
    np.random.seed(seed)
    mod = actor_critic(O, c["A"], 64, 2, 0.0, cont)                     # the reference's initialisation (layer_init)
    pol = R.RefPolicy({k: v.detach().clone() for k, v in mod.state_dict().items()}, cont)
    pol.requires_grad_(False)
    opt = R.RefAdam(pol.tensors(), lr=c["lr"], eps=1e-5)
    envs = G.SyncVectorEnv([G.make_env(gym_id, cont) for _ in range(N)], O)
    next_obs = torch.from_numpy(envs.reset(seed=list(range(N)))[0])
    next_done = torch.zeros(N)
    obs = torch.zeros(T, N, O); actions = torch.zeros(T, N, c["A"]) if cont else torch.zeros(T, N)
    logps = torch.zeros(T, N); rewards = torch.zeros(T, N); dones = torch.zeros(T, N); values = torch.zeros(T, N)
    batch = N * T
    mb = batch // c["nm"]
    num_updates = total // batch
    rets = []
    t0 = time.time()
    for update in range(1, num_updates + 1):
        opt.lr = R.lr_anneal(c["lr"], update, num_updates)
        for t in range(T):
            obs[t], dones[t] = next_obs, next_done
            with torch.no_grad():
                a, lp, _, v = pol.evaluate(next_obs)
            values[t], actions[t], logps[t] = v.flatten(), a, lp
            o, r, term, trunc, info = envs.step(a.numpy())
            rewards[t] = torch.tensor(r).view(-1)
            next_obs, next_done = torch.from_numpy(o), torch.from_numpy(term.astype(np.float32))
            for fi in info.get("final_info", []):                       # ppo.py:114-122: the first finished env of the step
                if fi is not None and "episode" in fi:
                    rets.append(float(fi["episode"]["r"]))
                    break
        with torch.no_grad():
            ret, adv = R.gae(rewards, values, dones, pol.value(next_obs), next_done, 0.99, 0.95)
        b = (obs.reshape(-1, O), actions.reshape(-1, c["A"]) if cont else actions.reshape(-1), logps.reshape(-1),
             adv.reshape(-1), ret.reshape(-1), values.reshape(-1))
        inds = np.arange(batch)
        for ep in range(c["epochs"]):
            np.random.shuffle(inds)
            for s in range(0, batch, mb):
                mi = torch.from_numpy(inds[s:s + mb])
                R.ppo_update_step(pol, opt, b[0][mi], b[1][mi], b[2][mi], b[3][mi], b[4][mi], b[5][mi], ent_c=c["ent"])
    dt = time.time() - t0
    rets = np.asarray(rets)
    k = max(len(rets) // 10, 1)
    curve = [float(rets[i:i + k].mean()) for i in range(0, len(rets) - k + 1, k)][:10]
    return {"num_envs": N, "num_steps": T, "total_timesteps": total, "logged_episodes": int(len(rets)),
            "mean_return_by_tenth_of_training": curve, "seconds": round(dt, 1), "env_steps_per_s_wall": round(total / dt)}


if __name__ == "__main__":
    torch.set_num_threads(int(os.environ.get("CPU_THREADS", "4")))
    out = {}
    for gym_id in sys.argv[1:] or ["CartPole-v1"]:
        out[gym_id] = run(gym_id)
        print(gym_id, out[gym_id], flush=True)
    os.makedirs("profiles", exist_ok=True)
    name = "profiles/r2_training_curves_cpu_%s.json" % "_".join(k.split("-")[0].lower() for k in out)
    json.dump(out, open(name, "w"), indent=1)
