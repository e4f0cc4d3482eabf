"""Whole iterations of the outer loop (src/ppo.py:192-273) on the device against the restatement, on the same inputs:
rollout -> GAE -> `num_update_epochs` x `num_minibatches` updates with the learning-rate anneal, twice in a row.

* the env trajectory of both iterations replays bit for bit on the checker (oracle/envs.c in its own trig mode) from the actions
  the device sampled;
* the stored log-probs / values are the restated model's at the parameters the iteration started from (2e-5);
* returns / advantages equal `run_gae` on the device's own rewards / values bit for bit;
* the parameter change of each iteration equals the restated update loop's (same keyed shuffles, same minibatch order, Adam with
  the annealed learning rate) within 1e-4 relative L2 of the CHANGE (parameters within 1e-6 absolute).

Covers the host orchestration (ppo.run_update: shuffle stream ids, moments of all minibatches ahead, minibatch slicing, anneal)
for the 64-wide fused kernels and for the wide tensor-core paths (128 / 256 units)."""
import numpy as np
import pytest
import torch

from aur_ppo_b200 import kernels, run_ppo
from oracle import envs as E
from oracle import ppo_ref as R
from tests.helpers import flat_from_named

pytestmark = pytest.mark.gpu


def _params(**kw):
    p = run_ppo.params_from_args(run_ppo.build_parser().parse_args([]))
    p.update(tensorboard=False, save=False)
    p.update(kw)
    return p


def _named_from_module(policy):
    return {k: v.detach().cpu().numpy().copy() for k, v in policy.state_dict().items()}


@pytest.mark.parametrize("gym_id,cont,hidden,layers", [("CartPole-v1", False, 64, 2), ("CartPole-v1", False, 128, 2),
                                                       ("Pendulum-v1", True, 64, 2), ("Pendulum-v1", True, 128, 2),
                                                       ("CartPole-v1", False, 256, 2), ("CartPole-v1", False, 128, 3)])
def test_two_iterations_vs_oracle(gym_id, cont, hidden, layers):
    from aur_ppo_b200.ppo import ppo
    N, T, NM, EP, U = 24, 48, 4, 2, 2
    torch.manual_seed(3)
    p = _params(gym_id=gym_id, continuous=cont, hidden_dim=hidden, num_layers=layers, num_envs=N, num_steps=T, num_minibatches=NM,
                num_update_epochs=EP, total_timesteps=N * T * U, learning_rate=1e-3, entropy_coeff=0.01 if not cont else 0.0)
    agent = ppo(p)
    obs_dim = 3 if cont else 4
    kind = E.PENDULUM if cont else E.CARTPOLE
    # the prelude of train() (ppo.py:180-190)
    agent.envs.reset(seed=list(agent.plan.env_ids))
    agent._env_step = 0
    agent._stats_rows = torch.zeros(EP * NM, kernels.NUM_STATS, device=agent.device)

    names = list(agent.policy.state_dict().keys())
    pol = R.RefPolicy(_named_from_module(agent.policy), cont)
    pol.requires_grad_(False)
    opt = R.RefAdam(pol.tensors(), lr=1e-3, eps=1e-5)
    shuffle_count = agent._shuffle_count
    acts_all, per_iter = [], []
    batch, mb = N * T, N * T // NM
    for u in range(1, U + 1):
        before = agent.flat.detach().cpu().numpy().copy()
        out = agent.run_update(u)
        torch.cuda.synchronize()
        b = agent.buffer
        dev = {k: getattr(b, k).cpu().numpy().copy() for k in ("states", "actions", "log_probs", "rewards", "terminals", "values")}
        dev["next_value"] = b.next_value.cpu().numpy().copy()
        dev["next_done"] = agent.envs.next_done.cpu().numpy().copy()
        dev["returns"], dev["advantages"] = agent._returns.cpu().numpy().copy(), agent._advantages.cpu().numpy().copy()
        acts_all.append(dev["actions"])
        # ---- the restated model at the parameters this iteration started from
        ot = torch.from_numpy(dev["states"].reshape(-1, obs_dim))
        at = torch.from_numpy(dev["actions"].reshape(-1, 1) if cont else dev["actions"].reshape(-1))
        with torch.no_grad():
            _, lp, _, v = pol.evaluate(ot, at)
        np.testing.assert_allclose(dev["log_probs"].reshape(-1), lp.numpy(), rtol=2e-5, atol=5e-6)
        np.testing.assert_allclose(dev["values"].reshape(-1), v.numpy().reshape(-1), rtol=2e-5, atol=5e-6)
        # ---- run_gae on the device's own rewards / values: bit for bit
        with torch.no_grad():
            ret, adv = R.gae(torch.from_numpy(dev["rewards"]), torch.from_numpy(dev["values"]), torch.from_numpy(dev["terminals"]),
                             torch.from_numpy(dev["next_value"]), torch.from_numpy(dev["next_done"]), 0.99, 0.95)
        assert np.array_equal(ret.numpy(), dev["returns"]) and np.array_equal(adv.numpy(), dev["advantages"])
        # ---- the update loop (ppo.py:208-269) with the device's keyed shuffles
        opt.lr = R.lr_anneal(1e-3, u, U)
        assert abs(agent.optimizer.param_groups[0]["lr"] - opt.lr) < 1e-15
        flat_b = (ot, at if not cont else at.reshape(-1, 1), torch.from_numpy(dev["log_probs"].reshape(-1)), adv.reshape(-1),
                  ret.reshape(-1), torch.from_numpy(dev["values"].reshape(-1)))
        ref_stats = None
        for ep in range(EP):
            inds = R.feistel_shuffle(batch, agent.shuffle_seed, shuffle_count)
            shuffle_count += 1
            for s in range(0, batch, mb):
                mi = torch.from_numpy(inds[s:s + mb].astype(np.int64))
                ref_stats = R.ppo_update_step(pol, opt, *[t[mi] for t in flat_b], max_grad_norm=0.5, ent_c=p["entropy_coeff"])[0]
        after = agent.flat.detach().cpu().numpy()
        want = flat_from_named({n: pol.p[n].detach().numpy() for n in names})
        d_dev, d_ref = after - before, want - before
        rel = float(np.linalg.norm(d_dev - d_ref) / np.linalg.norm(d_ref))
        print(f"{gym_id} hidden {hidden} x {layers}, iteration {u}: |param change| {np.linalg.norm(d_ref):.3e}, relative L2 error of the change {rel:.2e}")
        assert rel < 1e-4                      # measured 8e-6 .. 1.9e-5 (profiles/r2_whole_iteration_parity.txt)
        np.testing.assert_allclose(after, want, rtol=0, atol=1e-6)
        last = out["stats"][-1].cpu().numpy()
        for i, k in enumerate(kernels.STAT_NAMES):
            np.testing.assert_allclose(last[i], ref_stats[k], rtol=2e-3, atol=2e-5, err_msg=k)
        # keep the restated parameters ON the device's (the comparison above is per iteration, not cumulative)
        with torch.no_grad():
            for n, t in zip(names, pol.tensors()):
                t.copy_(agent.policy.state_dict()[n].detach().cpu())
        per_iter.append(dev)
    assert agent._shuffle_count == shuffle_count
    # ---- both iterations' trajectory on the checker, from the sampled actions
    acts = np.concatenate(acts_all).reshape(U * T, N, -1)
    acts = acts if cont else acts[..., 0].astype(np.int64)
    cv = E.CVecEnv(kind, N, wrappers=cont, trig=E.TRIG_CR)
    cur, _ = cv.reset(list(range(N)))
    cur_done = np.zeros(N, np.float32)
    for t in range(U * T):
        dev = per_iter[t // T]
        assert np.array_equal(dev["states"][t % T], cur) and np.array_equal(dev["terminals"][t % T], cur_done), t
        cur, r, term, trunc, info = cv.step(acts[t])
        assert np.array_equal(dev["rewards"][t % T], r.astype(np.float32)), t
        cur_done = term.astype(np.float32)
    assert np.array_equal(agent.envs.next_obs.cpu().numpy(), cur)
