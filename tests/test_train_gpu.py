"""End-to-end drop-in: ppo(params).train() on the GPU (reference config A and a wide config)."""
import os

import numpy as np
import pytest
import torch

from aur_ppo_b200 import compat, run_ppo

pytestmark = pytest.mark.gpu


def _params(**kw):
    p = run_ppo.params_from_args(run_ppo.build_parser().parse_args([]))
    p.update(tensorboard=False, save=False)
    p.update(kw)
    return p


def test_reference_config_a_learns_cartpole(tmp_path, monkeypatch):
    """CartPole-v1, num_envs=4, num_minibatches=4, 2-layer MLP (run_ppo.py defaults): the reference's
    only published outcome is the learning curve, so check the return climbs."""
    from aur_ppo_b200.ppo import ppo
    monkeypatch.chdir(tmp_path)
    torch.manual_seed(1)
    agent = ppo(_params(total_timesteps=80000, save=True))
    for attr in ("policy", "optimizer", "buffer", "envs", "batch_size", "minibatch_size", "num_updates"):
        assert hasattr(agent, attr)
    assert (agent.batch_size, agent.minibatch_size, agent.num_updates) == (512, 128, 156)
    assert agent.buffer.states.shape == (128, 4, 4) and agent.buffer.actions.shape == (128, 4)
    assert agent.buffer.actions.dtype == torch.float32
    rets, lens, xs = agent.train()
    assert len(rets) == len(lens) == len(xs) > 50
    assert xs == sorted(xs) and xs[-1] <= 80000
    first, last = np.mean(rets[:20]), np.mean(rets[-20:])
    assert first < 60 and last > 3 * first, (first, last)
    assert abs(agent.optimizer.param_groups[0]["lr"] - 2.5e-4 * (1 - 155 / 156)) < 1e-12      # lr anneal ppo.py:195-198
    st = agent.last_stats
    assert np.isfinite([st[k] for k in ("value_loss", "policy_loss", "entropy", "approx_kl", "clipfrac")]).all()
    # checkpoint: whole-module pickle, reference file name (ppo.py:296), loadable, same weights
    assert os.path.exists("actor_critic_2.pt")
    m = compat.load_policy("actor_critic_2.pt")
    for (n, a), (_, b) in zip(agent.policy.named_parameters(), m.named_parameters()):
        assert torch.equal(a.detach().cpu(), b.detach().cpu()), n
    obs = torch.zeros(3, 4)
    a, lp, v = m.act(obs)            # the checkpoint consumer of test.py:49-51
    assert a.shape == (3,)


def test_wide_config_runs_and_improves():
    from aur_ppo_b200.ppo import ppo
    torch.manual_seed(1)
    agent = ppo(_params(num_envs=2048, total_timesteps=2048 * 128 * 12))
    rets, lens, xs = agent.train()
    assert agent.num_updates == 12 and len(rets) > 100
    assert np.mean(rets[-50:]) > 1.5 * np.mean(rets[:50])


def test_hidden_128_runs_the_tensor_core_kernels_and_improves():
    """`-d 128` (src/run_ppo.py:36): rollout_tc_kernel<ENV, 128> + critic_values_tc_kernel<128> + the layer-wise update
    (update_wide.cu) behind the unchanged ppo(params).train()."""
    from aur_ppo_b200 import _lib
    from aur_ppo_b200.ppo import ppo
    torch.manual_seed(1)
    agent = ppo(_params(num_envs=2048, hidden_dim=128, total_timesteps=2048 * 128 * 12))
    L = _lib.lib()
    assert L.aur_ppo_update_get_wide() == 1 and L.aur_rollout_get_impl() == 1
    L.aur_launch_count_reset()
    rets, lens, xs = agent.train()
    # 12 iterations x 16 minibatches x >= 14 launches of the layer-wise path (the fused generic kernel needs 4)
    assert L.aur_launch_count() > 12 * 16 * 14
    assert agent.num_updates == 12 and len(rets) > 100
    assert np.mean(rets[-50:]) > 1.5 * np.mean(rets[:50])
    assert np.isfinite(list(agent.last_stats.values())).all()


def test_pendulum_continuous_runs():
    from aur_ppo_b200.ppo import ppo
    torch.manual_seed(1)
    p = _params(gym_id="Pendulum-v1", continuous=True, num_envs=512, num_steps=256, num_minibatches=8,
                num_update_epochs=4, total_timesteps=512 * 256 * 4, learning_rate=3e-4, entropy_coeff=0.0)
    agent = ppo(p)
    assert agent.buffer.actions.shape == (256, 512, 1) and agent.policy.actor_logstd.shape == (1, 1)
    rets, lens, xs = agent.train()
    assert all(l == 200 for l in lens) and len(rets) > 0
    assert np.isfinite(list(agent.last_stats.values())).all()


def test_unsupported_configs_fail_loudly():
    from aur_ppo_b200 import _lib
    from aur_ppo_b200.ppo import ppo
    with pytest.raises(_lib.AurError):
        ppo(_params(gym_id="LunarLander-v2"))                    # Box2D: no device kernel
    with pytest.raises(_lib.AurError):
        ppo(_params(continuous=True))                    # CartPole is discrete
    with pytest.raises(_lib.AurError, match="hidden_dim"):
        ppo(_params(hidden_dim=66, total_timesteps=512)).train()


@pytest.mark.parametrize("kw", [dict(hidden_dim=32), dict(hidden_dim=128, num_layers=3), dict(num_layers=1), dict(num_layers=4)])
def test_other_widths_and_depths_train_through_the_drop_in_api(kw):
    """`-d` / `-nl` (src/run_ppo.py:36,38) other than 64 / 2: runtime-width rollout + shape-generic update; CartPole
    returns improve within a short budget like the 64 x 2 config."""
    from aur_ppo_b200.ppo import ppo
    torch.manual_seed(1)
    agent = ppo(_params(num_envs=64, total_timesteps=64 * 128 * 12, **kw))
    rets, lens, xs = agent.train()
    assert len(rets) > 20 and np.isfinite(list(agent.last_stats.values())).all()
    k = max(5, len(rets) // 5)
    assert np.mean(rets[-k:]) > np.mean(rets[:k]) + 5, (np.mean(rets[:k]), np.mean(rets[-k:]))


def test_mountaincar_runs_through_the_drop_in_api():
    """--gym_id MountainCar-v0 ((f) rank 4): obs 2 / 3 actions flow through rollout, GAE, packed records and the update."""
    from aur_ppo_b200.ppo import ppo
    torch.manual_seed(1)
    agent = ppo(_params(gym_id="MountainCar-v0", num_envs=256, total_timesteps=256 * 128 * 3))
    assert agent.buffer.states.shape == (128, 256, 2) and agent.policy.actor.net[-1].out_features == 3
    rets, lens, xs = agent.train()
    assert agent.num_updates == 3 and len(rets) > 0
    assert all(l <= 200 for l in lens) and all(r == -float(l) for r, l in zip(rets, lens))     # reward -1 per step
    assert np.isfinite([agent.last_stats[k] for k in ("value_loss", "policy_loss", "entropy", "approx_kl")]).all()


def test_acrobot_runs_through_the_drop_in_api():
    """--gym_id Acrobot-v1 ((f) rank 4): obs 6 / 3 actions through the runtime-width rollout, GAE and the shape-generic
    update; the policy learns to swing up (episodes get shorter than the 500-step limit)."""
    from aur_ppo_b200.ppo import ppo
    torch.manual_seed(1)
    agent = ppo(_params(gym_id="Acrobot-v1", num_envs=16, total_timesteps=16 * 128 * 200))
    assert agent.buffer.states.shape == (128, 16, 6)
    rets, lens, xs = agent.train()
    assert len(rets) > 20 and np.isfinite(list(agent.last_stats.values())).all()
    k = max(5, len(rets) // 5)
    assert np.mean(lens[-k:]) < 0.6 * np.mean(lens[:k]), (np.mean(lens[:k]), np.mean(lens[-k:]))


def test_mountaincar_continuous_runs_through_the_drop_in_api():
    """--gym_id MountainCarContinuous-v0 --continuous True: wrapper stack with two observation dims, Normal policy, packed
    records with obs 2 / one action column; statistics stay finite and episodes are logged (999-step limit)."""
    from aur_ppo_b200.ppo import ppo
    torch.manual_seed(1)
    p = _params(gym_id="MountainCarContinuous-v0", continuous=True, num_envs=64, num_steps=256, num_minibatches=8,
                num_update_epochs=4, total_timesteps=64 * 256 * 8, learning_rate=3e-4, entropy_coeff=0.0)
    agent = ppo(p)
    assert agent.buffer.actions.shape == (256, 64, 1) and agent.buffer.states.shape == (256, 64, 2)
    rets, lens, xs = agent.train()
    assert len(rets) > 0 and all(0 < l <= 999 for l in lens)
    assert np.isfinite(list(agent.last_stats.values())).all()
