"""Per-kernel timing on one GPU (CUDA events, L2 flushed between iterations)."""
import argparse
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from aur_ppo_b200 import kernels

PEAK = 6450.9
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timeit(fn, iters=20, warmup=5, flush=True):
    scratch = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if flush else None
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        if flush:
            scratch.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def bench_gae(args):
    for T, N in [(128, 65536), (256, 65536), (128, 131072), (512, 131072), (2048, 131072), (128, 1048576), (128, 4096)]:
        rew = torch.rand(T, N, device="cuda"); val = torch.randn(T, N, device="cuda")
        term = (torch.rand(T, N, device="cuda") < 1 / 200).float()
        nv = torch.randn(N, device="cuda"); nd = torch.zeros(N, device="cuda")
        out = (torch.empty_like(rew), torch.empty_like(rew))
        med, best = timeit(lambda: kernels.gae(rew, val, term, nv, nd, 0.99, 0.95, True, out))
        byt = 20 * T * N + 8 * N
        print(json.dumps({"kernel": "gae", "T": T, "N": N, "us_median": med * 1e6, "us_best": best * 1e6,
                          "GBps_median": byt / med / 1e9, "frac_of_measured_peak": byt / med / 1e9 / PEAK}))
        del rew, val, term, out


def _policy(cont=False):
    from tests.helpers import flat_from_named, random_policy
    _, named = random_policy(3 if cont else 4, 1 if cont else 2, 64, 2, cont, seed=1)
    desc = kernels.policy_desc(3 if cont else 4, 1 if cont else 2, 64, 2, cont)
    return desc, torch.from_numpy(flat_from_named(named)).cuda()


def bench_rollout(args):
    from aur_ppo_b200 import envs as denv
    for gym_id, N, T in [("CartPole-v1", 65536, 128), ("CartPole-v1", 131072, 128), ("CartPole-v1", 4096, 128),
                         ("Pendulum-v1", 65536, 256)]:
        cont = gym_id.startswith("Pend")
        desc, flat = _policy(cont)
        env = denv.DeviceVecEnv(gym_id, N, wrappers=cont)
        env.reset(list(range(N)))
        buf = kernels.RolloutBuffers(T, N, env.obs_dim, (1,) if cont else (), "cuda")
        step = [0]
        def run():
            kernels.rollout(env, desc, flat, buf, seed=1, step0=step[0]); step[0] += T
        med, best = timeit(run, iters=10, warmup=3)
        print(json.dumps({"kernel": "rollout", "env": gym_id, "N": N, "T": T, "ms_median": med * 1e3,
                          "env_steps_per_s": N * T / med, "HBM_GBps": 36 * N * T / med / 1e9,
                          "fp32_TFLOPs": 17792 * N * T / med / 1e12}))


def bench_update(args):
    desc, flat = _policy(False)
    for B, m in [(8388608, 2097152), (8388608, 524288), (512, 128)]:
        g = torch.Generator(device="cuda").manual_seed(3)
        dbuf = [torch.randn(B, 4, generator=g, device="cuda") * 0.5, torch.randint(0, 2, (B,), generator=g, device="cuda").float(),
                -0.7 + 0.1 * torch.randn(B, generator=g, device="cuda"), torch.randn(B, generator=g, device="cuda"),
                torch.randn(B, generator=g, device="cuda"), torch.randn(B, generator=g, device="cuda")]
        idx = torch.randperm(B, generator=g, device="cuda")[:m].to(torch.int32)
        up = kernels.Updater(desc, flat.clone())
        rec = kernels.pack_records(*dbuf)
        med, best = timeit(lambda: up.step(*dbuf, idx, lr=2.5e-4, records=rec), iters=10, warmup=3)
        print(json.dumps({"kernel": "update(moments+grad+reduce+adam)", "B": B, "m": m, "ms_median": med * 1e3,
                          "samples_per_s": m / med, "fp32_TFLOPs": 53400 * m / med / 1e12, "gather_GBps": 40 * m / med / 1e9}))


def bench_equiv(args):
    """Config D: equivariant actor-critic update on synthetic close_loop_block_picking-shaped obs, minibatch 4096."""
    from aur_ppo_b200 import equiv
    B = int(os.environ.get("EQUIV_B", "4096"))
    params = equiv.init_params(seed=0)
    for k in ("actor.head.psi_triv", "actor.head.psi_irrep", "critic.head2.w"):
        params[k].mul_(0.1)
    g = torch.Generator(device="cuda").manual_seed(0)
    obs = torch.rand(B, 1, 128, 128, generator=g, device="cuda") * 0.32
    state = (torch.rand(B, generator=g, device="cuda") > 0.5).float()
    action = torch.randn(B, 5, generator=g, device="cuda")
    adv, ret, vold = (torch.randn(B, generator=g, device="cuda") for _ in range(3))
    oldlp = torch.full((B,), -7.0, device="cuda")
    model = equiv.EquivActorCritic(params, B)
    iters = int(os.environ.get("EQUIV_ITERS", "3"))
    def fwd():
        model.forward(state, obs)
    def upd():
        model.update(state, obs, action, oldlp, adv, ret, vold)
    med_f, _ = timeit(fwd, iters=iters, warmup=1, flush=False)
    med_u, _ = timeit(upd, iters=iters, warmup=1, flush=False)
    flops_fwd = 2 * 2.80e9 * B          # two encoders
    flops_upd = 3 * flops_fwd
    print(json.dumps({"kernel": "equiv update (2 encoders fwd+bwd+Adam)", "B": B, "ms_forward": med_f * 1e3, "ms_update": med_u * 1e3,
                      "samples_per_s": B / med_u, "TFLOPs_forward": flops_fwd / med_f / 1e12, "TFLOPs_update": flops_upd / med_u / 1e12,
                      "frac_of_bf16_sustained_peak": flops_upd / med_u / 1e12 / 1414.7,
                      "mem_GB": torch.cuda.max_memory_allocated() / 1e9}))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="*", default=["gae", "rollout", "update"])
    a = ap.parse_args()
    for w in a.which:
        globals()["bench_" + w](a)
