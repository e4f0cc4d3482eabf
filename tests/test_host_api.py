"""Host-side mirror of the reference interface: CLI flags/defaults (run_ppo.py:14-81), model API
(models/actor_critic.py), checkpoint layout (ppo.py:296).  CPU only."""
import io
import os
import zipfile

import numpy as np
import pytest
import torch

from aur_ppo_b200 import compat, run_ppo
from aur_ppo_b200.models.actor_critic import actor_critic
from oracle import ppo_ref as R
from tests.helpers import flat_from_named

REF_KEYS = ['gym_id', 'seed', 'num_steps', 'gae', 'total_timesteps', 'anneal_lr', 'gae_lambda', 'num_update_epochs',
            'num_envs', 'num_minibatches', 'entropy_coeff', 'value_coeff', 'clip_coeff', 'clip_vloss', 'max_grad_norm',
            'target_kl', 'norm_adv', 'capture_video', 'hidden_dim', 'continuous', 'learning_rate', 'exp_name',
            'num_layers', 'dropout', 'gamma', 'track']


def test_cli_defaults_match_reference():
    p = run_ppo.params_from_args(run_ppo.build_parser().parse_args([]))
    assert list(p.keys()) == REF_KEYS
    assert (p['gym_id'], p['seed'], p['num_steps'], p['total_timesteps']) == ('CartPole-v1', 1.0, 128, 500000)
    assert (p['num_envs'], p['num_minibatches'], p['num_update_epochs']) == (4, 4, 4)
    assert (p['gae_lambda'], p['gamma'], p['clip_coeff'], p['entropy_coeff'], p['value_coeff']) == (0.95, 0.99, 0.2, 0.01, 0.5)
    assert (p['learning_rate'], p['max_grad_norm'], p['hidden_dim'], p['num_layers'], p['dropout']) == (2.5e-4, 0.5, 64, 2, 0.0)
    assert p['gae'] is True and p['anneal_lr'] is True and p['clip_vloss'] is True and p['norm_adv'] is True
    assert p['continuous'] is False and p['target_kl'] is None and p['track'] is False and p['capture_video'] is False


def test_cli_short_flags_and_bool_parsing():
    a = run_ppo.build_parser().parse_args(['-id', 'CartPole-v1', '-ne', '65536', '-nm', '4', '-ns', '128', '-nl', '2',
                                           '-d', '64', '-do', '0.0', '-t', '1000', '-gae', 'false', '-al', 'no',
                                           '-cvl', '0', '-na', 'False', '-tkl', '0.015', '-g', '0.9'])
    p = run_ppo.params_from_args(a)
    assert p['num_envs'] == 65536 and p['total_timesteps'] == 1000 and p['gamma'] == 0.9
    assert p['gae'] is False and p['anneal_lr'] is False and p['clip_vloss'] is False and p['target_kl'] == 0.015
    assert p['norm_adv'] is True      # `type=bool` quirk of the reference: 'False' is a non-empty string
    with pytest.raises(SystemExit):
        run_ppo.build_parser().parse_args(['-gae', 'maybe'])


def test_continuous_override_and_opt_out():
    p = run_ppo.params_from_args(run_ppo.build_parser().parse_args(['-c', 'true', '-id', 'Pendulum-v1', '-ne', '64']))
    assert (p['num_envs'], p['num_steps'], p['num_minibatches'], p['num_update_epochs']) == (1, 2048, 32, 10)
    assert (p['learning_rate'], p['total_timesteps'], p['entropy_coeff']) == (3e-4, 2000000, 0)
    p = run_ppo.params_from_args(run_ppo.build_parser().parse_args(
        ['-c', 'true', '-id', 'Pendulum-v1', '-ne', '65536', '-ns', '256', '--no_continuous_override']))
    assert (p['num_envs'], p['num_steps'], p['continuous']) == (65536, 256, True)


def test_strtobool_matches_distutils_table():
    for v in ("y", "yes", "t", "true", "on", "1", "TRUE"):
        assert run_ppo.strtobool(v) == 1
    for v in ("n", "no", "f", "false", "off", "0"):
        assert run_ppo.strtobool(v) == 0
    with pytest.raises(ValueError):
        run_ppo.strtobool("2")


@pytest.mark.parametrize("tag", ["disc", "cont", "disc3"])
def test_actor_critic_matches_reference_golden(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "model.npz"))
    names = [str(n) for n in g[f"{tag}_names"]]
    cont = "actor_logstd" in names
    sd = g[f"{tag}_obs"].shape[1]
    nl = len([n for n in names if n.startswith("critic.") and n.endswith("weight")]) - 1
    ad = (g[f"{tag}_act"].shape[1],) if cont else 2
    m = actor_critic(sd, ad, 64, nl, 0.0, cont)
    assert [n for n, _ in m.named_parameters()] == names            # same parameter names and order
    m.load_state_dict({n: torch.from_numpy(g[f"{tag}_p_{n}"]) for n in names})
    obs, act = torch.from_numpy(g[f"{tag}_obs"]), torch.from_numpy(g[f"{tag}_act"])
    with torch.no_grad():
        a, lp, ent, v = m.evaluate(obs, act)
        a2, lp2, ent2, v2 = m.get_action_and_value(obs, act)
    np.testing.assert_allclose(lp.numpy(), g[f"{tag}_logp"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(ent.numpy(), g[f"{tag}_ent"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(v.numpy(), g[f"{tag}_value"], rtol=1e-6, atol=1e-7)
    assert v.shape == (obs.shape[0], 1) and m.value(obs).shape == (obs.shape[0],) and m.get_value(obs).shape == (obs.shape[0],)
    assert torch.equal(lp, lp2)
    s_a, s_lp, s_v = m.act(obs)
    assert s_a.shape[0] == obs.shape[0] and s_lp.shape == (obs.shape[0],)
    # flat buffer: canonical order, parameters are views of it
    flat = m.flat_parameters()
    np.testing.assert_array_equal(flat.numpy(), flat_from_named({n: g[f"{tag}_p_{n}"] for n in names}))
    flat.mul_(2.0)
    assert torch.equal(m.actor.net[0].weight.detach(), 2 * torch.from_numpy(g[f"{tag}_p_actor.net.0.weight"]))
    assert m.flat_parameters().data_ptr() == flat.data_ptr()


def test_init_matches_reference_scheme():
    torch.manual_seed(1)
    m = actor_critic(4, 2, 64, 2, 0.0, False)
    assert sum(p.numel() for p in m.parameters()) == 9155
    w = m.actor.net[0].weight.detach()          # [64,4] orthogonal columns scaled by sqrt(2)
    np.testing.assert_allclose((w.T @ w).numpy(), 2 * np.eye(4), atol=1e-5)
    wl = m.actor.net[4].weight.detach()         # head gain 0.01
    np.testing.assert_allclose((wl @ wl.T).numpy(), 1e-4 * np.eye(2), atol=1e-8)
    wc = m.critic.net[4].weight.detach()
    np.testing.assert_allclose((wc @ wc.T).numpy(), np.eye(1), atol=1e-6)
    assert all(float(b.abs().max()) == 0 for n, b in m.named_parameters() if n.endswith("bias"))
    mc = actor_critic(3, (1,), 64, 2, 0.0, True)
    assert sum(p.numel() for p in mc.parameters()) == 8963 and mc.actor_logstd.shape == (1, 1)


def test_checkpoint_pickle_carries_reference_class_paths(tmp_path):
    torch.manual_seed(0)
    m = actor_critic(4, 2, 64, 2, 0.0, False)
    m.flat_parameters()
    path = str(tmp_path / "actor_critic_2.pt")
    compat.save_policy(m, path)
    data = zipfile.ZipFile(path).read([n for n in zipfile.ZipFile(path).namelist() if n.endswith("data.pkl")][0])
    assert b"models.actor_critic" in data and b"nets.nets" in data and b"aur_ppo_b200" not in data
    m2 = compat.load_policy(path)
    assert [n for n, _ in m2.named_parameters()] == [n for n, _ in m.named_parameters()]
    for (n, a), (_, b) in zip(m.named_parameters(), m2.named_parameters()):
        assert torch.equal(a, b), n
    for attr in ("state_dim", "action_dim", "hidden_dim", "continuous", "num_layers", "dropout"):
        assert getattr(m2, attr) == getattr(m, attr)


def test_ppo_class_fails_loudly_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from aur_ppo_b200 import _lib
    from aur_ppo_b200.ppo import ppo
    p = run_ppo.params_from_args(run_ppo.build_parser().parse_args([]))
    with pytest.raises(_lib.AurError, match="no CPU fallback"):
        ppo(p)


def test_robot_actor_critic_refuses_cpu():
    """No CPU path behind the robot_actor_critic facade (either model family)."""
    from aur_ppo_b200 import _lib
    from aur_ppo_b200.models import robot_actor_critic
    for equivariant in (True, False):
        with pytest.raises(_lib.AurError):
            robot_actor_critic("cpu", equivariant)
