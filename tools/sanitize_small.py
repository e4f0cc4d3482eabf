"""One small launch of every kernel family, meant to run under compute-sanitizer (race detection):

  compute-sanitizer --tool memcheck  --error-exitcode 7 python tools/sanitize_small.py
  compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_small.py

Shapes are tiny so that the 10-100x slowdown of the tools stays within minutes; results are still checked (smoke()
compares with the oracle)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import __graft_entry__ as G  # noqa: E402
from aur_ppo_b200 import _lib, envs as denv, equiv, kernels, plain_cnn  # noqa: E402


def mlp_families():
    L = _lib.lib()
    G.smoke()                                   # env_reset, rollout_tc + critic_values_tc, gae_bulk, shuffle, moments, ppo_grad_tc<2>, reduce, adam
    torch.manual_seed(0)
    N, T = 256, 16
    for gym_id, cont, act in (("CartPole-v1", False, 2), ("Pendulum-v1", True, 1), ("Acrobot-v1", False, 3)):
        env = denv.DeviceVecEnv(gym_id, N, wrappers=cont)
        env.reset(list(range(N)))
        desc = kernels.policy_desc(env.obs_dim, act, 64, 2, cont)
        P = kernels.policy_param_count(desc)
        flat = (torch.rand(P, device="cuda") - 0.5) * 0.2
        buf = kernels.RolloutBuffers(T, N, env.obs_dim, (act,) if cont else (), "cuda")
        for impl in (0, 1):                     # SIMT and tensor-core rollout
            L.aur_rollout_set_impl(impl)
            kernels.rollout(env, desc, flat, buf, seed=1, step0=0)
        L.aur_rollout_set_impl(1)
        ret, adv = kernels.gae(buf.rewards, buf.values, buf.terminals, buf.next_value, env.next_done, 0.99, 0.95)
        ret2, adv2 = kernels.gae(buf.rewards[:, :250].contiguous(), buf.values[:, :250].contiguous(), buf.terminals[:, :250].contiguous(),
                                 buf.next_value[:250].contiguous(), env.next_done[:250].contiguous(), 0.99, 0.95)   # ragged: column kernel
        fb = [buf.states.reshape(-1, env.obs_dim), buf.actions.reshape(-1, act) if cont else buf.actions.reshape(-1),
              buf.log_probs.reshape(-1), adv.reshape(-1), ret.reshape(-1), buf.values.reshape(-1)]
        idx = kernels.shuffle_indices(T * N, seed=1, stream_id=0)
        rec = kernels.pack_records(fb[0], fb[1], fb[2], fb[3], fb[4], fb[5])
        for impl in (0, 1, 2, 3):               # simt, tc<2>, tc<4>, generic
            L.aur_ppo_update_set_impl(impl)
            up = kernels.Updater(desc, flat.clone())
            up.grad(*fb, idx[:1024].contiguous(), records=rec)
            up.apply(2.5e-4, 0.5)
        L.aur_ppo_update_set_impl(1)
        kernels.policy_evaluate(desc, flat, env.next_obs)
        kernels.policy_evaluate(desc, flat, env.next_obs, greedy=True)
    m, ls = torch.randn(64, 5, device="cuda"), torch.zeros(64, 5, device="cuda")
    kernels.squashed_gaussian_sample(m, ls, seed=1)
    torch.cuda.synchronize()
    print("mlp families ok")


def cnn_families(split):
    B = 8
    g = torch.Generator(device="cuda").manual_seed(0)
    obs = torch.rand(B, 1, 128, 128, generator=g, device="cuda") * 0.32
    state = (torch.rand(B, generator=g, device="cuda") > 0.5).float()
    action = torch.randn(B, 5, generator=g, device="cuda")
    adv, ret, vold = (torch.randn(B, generator=g, device="cuda") for _ in range(3))
    oldlp = torch.full((B,), -7.0, device="cuda")
    for make in (lambda: equiv.EquivActorCritic(equiv.init_params(seed=0), B, precision=('fp32' if split else 'bf16')),
                 lambda: plain_cnn.PlainActorCritic(plain_cnn.init_params(seed=0), B, precision=('fp32' if split else 'bf16'))):
        model = make()
        model.update(state, obs, action, oldlp, adv, ret, vold)
        torch.cuda.synchronize()
        del model
    print("cnn families ok (split=%s)" % split)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "mlp"):
        mlp_families()
    if which in ("all", "cnn"):
        cnn_families(False)
    if which in ("all", "cnn_split"):
        cnn_families(True)
