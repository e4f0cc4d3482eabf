"""Checkpoint consumer: the reference's shipped whole-module pickles
(tests/golden/ckpt/*.pt, copied by oracle/gen_golden_ckpt.py) load through compat.load_policy, evaluate like the
reference's own classes did (tests/golden/ckpt.npz holds the outputs of the reference's code on the same pickles),
run through the CUDA policy kernels, and play CartPole on the device envs like src/test.py:17-61."""
import os

import numpy as np
import pytest
import torch

from aur_ppo_b200 import compat

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CKPT = os.path.join(GOLD, "ckpt")
TAGS = ["actor_critic", "actor_critic_2", "actor_critic_10"]


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "ckpt.npz"))


def test_shipped_copies_are_identical(gold):
    # plots/actor_critic.pt and src/models/saved/actor_critic.pt are byte-identical: one fixture covers both
    assert bool(gold["saved_equals_plots"])
    for t in ("logp", "entropy", "value", "logits"):
        assert np.array_equal(gold[f"actor_critic_{t}"], gold[f"actor_critic_saved_{t}"])


@pytest.mark.parametrize("tag", TAGS)
def test_reference_pickle_loads_and_evaluates_like_the_reference(tag, gold):
    m = compat.load_policy(os.path.join(CKPT, tag + ".pt"))
    assert type(m).__name__ == "actor_critic" and list(m.state_dict().keys()) == [str(k) for k in gold[f"{tag}_keys"]]
    obs, act = torch.from_numpy(gold[f"{tag}_obs"]), torch.from_numpy(gold[f"{tag}_act"])
    with torch.no_grad():
        a, lp, ent, val = m.evaluate(obs, act)
        assert torch.equal(a, act)
        np.testing.assert_allclose(lp.numpy(), gold[f"{tag}_logp"], rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(ent.numpy(), gold[f"{tag}_entropy"], rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(val.numpy().reshape(-1), gold[f"{tag}_value"], rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(m.value(obs).numpy(), gold[f"{tag}_value2"], rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(m.get_value(obs).numpy(), gold[f"{tag}_value2"], rtol=2e-5, atol=2e-6)
        a2, lp2, v2 = m.act(obs)                                   # test.py:51 call shape
        assert a2.shape == (64,) and lp2.shape == (64,) and v2.shape == (64, 1)
    # legacy layout: Linear layers at Sequential indices 0, 3, 6, ... (Dropout in between)
    assert "actor.net.3.weight" in m.state_dict()
    shape = m.kernel_shape()
    assert shape[2] == 64 and shape[3] == (10 if tag.endswith("_10") else 2)


def test_round_trip_keeps_the_reference_class_paths(tmp_path):
    m = compat.load_policy(os.path.join(CKPT, "actor_critic.pt"))
    p = tmp_path / "actor_critic_2.pt"
    compat.save_policy(m, str(p))
    raw = open(p, "rb").read()
    assert b"models.actor_critic" in raw and b"aur_ppo_b200" not in raw
    m2 = compat.load_policy(str(p))
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k


# ------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("tag", TAGS)
def test_policy_kernels_on_reference_checkpoints(tag, gold):
    """aur_policy_act (the 64-wide kernel for actor_critic.pt, the runtime-shape kernel for the 8-observation and the
    10-layer checkpoints) vs the outputs of the reference's own classes."""
    from aur_ppo_b200 import kernels
    m = compat.load_policy(os.path.join(CKPT, tag + ".pt")).cuda()
    desc = kernels.policy_desc(*m.kernel_shape())
    flat = m.flat_parameters()
    obs = torch.from_numpy(gold[f"{tag}_obs"]).cuda()
    act = torch.from_numpy(gold[f"{tag}_act"]).float().cuda()
    a, lp, ent, val = kernels.policy_evaluate(desc, flat, obs, act)
    tol = dict(rtol=1e-4, atol=1e-5) if tag.endswith("_10") else dict(rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(lp.cpu().numpy(), gold[f"{tag}_logp"], **tol)
    np.testing.assert_allclose(ent.cpu().numpy(), gold[f"{tag}_entropy"], **tol)
    np.testing.assert_allclose(val.cpu().numpy(), gold[f"{tag}_value"], **tol)
    # greedy = torch.argmax of the reference's logits (ties broken to the first maximum), away from near-ties
    g, _, _, _ = kernels.policy_evaluate(desc, flat, obs, greedy=True)
    logits = gold[f"{tag}_logits"]
    top2 = np.sort(logits, axis=1)[:, -2:]
    clear = (top2[:, 1] - top2[:, 0]) > 1e-4
    assert clear.sum() > 32
    assert np.array_equal(g.cpu().numpy()[clear].astype(np.int64), logits.argmax(1)[clear])


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["sampled", "greedy"])
def test_play_reference_checkpoint_on_device_envs(mode):
    """src/test.py:17-61 on the device envs: the shipped CartPole policy plays 256 episodes; every episode is replayed on the
    CPU restatement of gym's CartPole with the same seeds and the actions the policy chose on the device."""
    from aur_ppo_b200.test import test as Player
    from oracle import envs as E
    t = Player(os.path.join(CKPT, "actor_critic.pt"), "CartPole-v1")
    N = 256
    lengths = t.run(N, mode=mode, seed=3)
    assert len(lengths) == N and min(lengths) >= 8 and max(lengths) <= 500
    assert all(abs(r - l) < 1e-3 for r, l in zip(t.episode_returns, lengths))      # CartPole: reward 1 per step
    # the shipped policy is a trained one: it balances far longer than a random policy (~22 steps)
    assert np.mean(lengths) > 100, np.mean(lengths)
    # replay on the checker: same seeds, actions recomputed with the reference-equivalent torch module on the SAME
    # observations the checker produces; greedy play is deterministic, so the lengths must agree exactly
    if mode == "greedy":
        m = compat.load_policy(os.path.join(CKPT, "actor_critic.pt"))
        cv = E.CVecEnv(E.CARTPOLE, N, wrappers=False, trig=E.TRIG_CR)
        obs, _ = cv.reset(list(range(3, 3 + N)))
        want = np.zeros(N, np.int64)
        done = np.zeros(N, bool)
        near_tie = np.zeros(N, bool)
        for step in range(500):
            with torch.no_grad():
                logits = m.actor(torch.from_numpy(obs)).numpy()
            near_tie |= (np.abs(logits[:, 0] - logits[:, 1]) < 1e-4) & ~done
            obs, r, term, trunc, info = cv.step(logits.argmax(1))
            fin = (term | trunc) & ~done
            want[fin] = step + 1
            done |= fin
            if done.all():
                break
        ok = ~near_tie
        assert ok.sum() > N // 2
        assert np.array_equal(np.asarray(lengths)[ok], want[ok])


@pytest.mark.gpu
def test_player_rejects_mismatched_checkpoint():
    from aur_ppo_b200 import _lib
    from aur_ppo_b200.test import test as Player
    with pytest.raises(_lib.AurError):
        Player(os.path.join(CKPT, "actor_critic_2.pt"), "CartPole-v1")           # 8 observations: not a CartPole policy
