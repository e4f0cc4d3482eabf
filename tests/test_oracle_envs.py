"""Pins the env checker (oracle/envs.c, oracle/gym_restated.py).  CPU only.

The gym boundary is PARITY UNPINNED (third-party gym==0.26.2 is absent and the
reference holds no env fixtures); what can be pinned is pinned here: NumPy's
PCG64 stream, the published CartPole reset(seed=0) answer, C == Python
restatement bit for bit, and the deterministic sincos against a 200-bit
reference."""
import math

import numpy as np
import pytest

from oracle import envs as E
from oracle import gym_restated as G


def test_cartpole_reset_seed0_known_answer():
    env = G.CartPoleEnv()
    obs, _ = env.reset(seed=0)
    full = [0.013696168732145436, -0.02302132862361297, -0.045902647606380534, -0.04834723644714709]
    np.testing.assert_array_equal(obs, np.array(full, np.float32))
    np.testing.assert_allclose(obs, [0.01369617, -0.02302133, -0.04590265, -0.04834723], rtol=2e-7)
    assert env.state[0] == 0.013696168732145436
    cv = E.CVecEnv(E.CARTPOLE, 3)
    o, _ = cv.reset([0, 1, 2])
    np.testing.assert_array_equal(o[0], obs)


def test_pcg64_stream_matches_numpy():
    for seed in (0, 1, 12345, 2**40 + 7):
        st = E.pcg64_seed_states([seed])[0]
        out = np.empty(64, np.float64)
        E.lib().orc_pcg64_doubles(int(st[0]), int(st[1]), int(st[2]), int(st[3]), 64, out.ctypes.data)
        want = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed))).random(64)
        np.testing.assert_array_equal(out, want)


@pytest.mark.parametrize("kind,gid,cont", [(E.CARTPOLE, "CartPole-v1", False), (E.PENDULUM, "Pendulum-v1", True),
                                           (E.PENDULUM, "Pendulum-v1", False), (E.MOUNTAINCAR, "MountainCar-v0", False),
                                           (E.ACROBOT, "Acrobot-v1", False),
                                           (E.MOUNTAINCAR_CONT, "MountainCarContinuous-v0", True),
                                           (E.MOUNTAINCAR_CONT, "MountainCarContinuous-v0", False)])
def test_c_checker_equals_python_restatement(kind, gid, cont):
    N = 6
    pv = G.SyncVectorEnv([G.make_env(gid, cont) for _ in range(N)], E.OBS_DIM[kind])
    cv = E.CVecEnv(kind, N, wrappers=cont, trig=E.TRIG_LIBM)
    o1, _ = pv.reset(seed=list(range(N)))
    o2, _ = cv.reset(list(range(N)))
    np.testing.assert_array_equal(o1, o2)
    rng = np.random.default_rng(5)
    episodes = 0
    for t in range(1100 if kind == E.MOUNTAINCAR_CONT else 700):
        a = (rng.integers(0, 2, N) if kind == E.CARTPOLE else rng.integers(0, 3, N) if kind in (E.MOUNTAINCAR, E.ACROBOT)
             else rng.normal(0, 1.5, (N, 1)).astype(np.float32))
        r1, r2 = pv.step(a), cv.step(a)
        for k in range(4):
            np.testing.assert_array_equal(r1[k], r2[k], err_msg=f"t={t} field={k}")
        if "final_info" in r1[4]:
            for i, it in enumerate(r1[4]["final_info"]):
                if it is not None:
                    e2 = r2[4]["final_info"][i]["episode"]
                    assert it["episode"]["r"] == e2["r"] and it["episode"]["l"] == e2["l"]
                    episodes += 1
    assert episodes > 0


def test_timelimit_truncation_keeps_done_zero():
    """ppo.py:110 discards `truncated`: a TimeLimit reset leaves done = 0."""
    cv = E.CVecEnv(E.PENDULUM, 2, wrappers=True)
    cv.reset([0, 1])
    for t in range(200):
        _, _, term, trunc, info = cv.step(np.zeros(2, np.float32))
        assert not term.any()
    assert trunc.all() and "final_info" in info and info["final_info"][0]["episode"]["l"] == 200


def test_det_sincos_correctly_rounded_sample():
    mp = pytest.importorskip("mpmath")
    mp.mp.prec = 200
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.uniform(-0.3, 0.3, 1500), rng.uniform(-100, 100, 1500), rng.uniform(-1.6e6, 1.6e6, 500),
                        [0.0, -0.0, 1e-300, math.pi / 4, -math.pi / 4, math.pi / 2, 3 * math.pi, 1e-9]])
    s, c = E.det_sincos(x)
    bad = 0
    for xi, si, ci in zip(x, s, c):
        bad += (si != float(mp.sin(mp.mpf(float(xi))))) + (ci != float(mp.cos(mp.mpf(float(xi)))))
    assert bad <= 1, bad     # correctly rounded except within ~2^-14 ulp of a tie


def test_det_vs_libm_discrepancy_is_last_bit_only():
    rng = np.random.default_rng(0)
    x = rng.uniform(-0.3, 0.3, 400_000)
    s, c = E.det_sincos(x)
    s2, c2 = E.libm_sincos(x)
    ulp = np.maximum(np.abs(s - s2) / np.spacing(np.abs(s2)), np.abs(c - c2) / np.spacing(np.abs(c2)))
    assert ulp.max() <= 1.0
    assert np.mean(s != s2) < 0.01 and np.mean(c != c2) < 0.01


def test_det_sincos_vs_independent_correctly_rounded_reference():
    """The product's sin / cos (det_sincos.h) against the checker's OWN reference (libquadmath rounded once = correctly
    rounded): a census of 2e6 random points per env range; the full 1e8-point census is profiles/r2_sincos_census.md
    (det: 6e-8 .. 1.4e-6 of the points differ in the last bit; glibc libm: 0.14 .. 0.28 %)."""
    rng = np.random.default_rng(5)
    for lo, hi in ((-0.42, 0.42), (-85.0, 85.0), (-6.3, 6.3), (-3.6, 1.8)):
        x = rng.uniform(lo, hi, 2_000_000)
        s, c = E.det_sincos(x)
        s2, c2 = E.cr_sincos(x)
        s3, c3 = E.libm_sincos(x)
        det_bad = int((s != s2).sum() + (c != c2).sum())
        libm_bad = int((s3 != s2).sum() + (c3 != c2).sum())
        assert det_bad <= 20, (lo, hi, det_bad)                       # <= 5e-6 of 4e6 values
        assert libm_bad > 50 * max(det_bad, 1), (lo, hi, libm_bad)     # the libm gym calls is the less reproducible one
        bad = (s != s2) | (c != c2)
        if bad.any():                                                 # and where it differs it is the neighbouring double
            assert np.all(np.abs(s[bad] - s2[bad]) <= np.spacing(np.abs(s2[bad])))
            assert np.all(np.abs(c[bad] - c2[bad]) <= np.spacing(np.abs(c2[bad])))


@pytest.mark.parametrize("kind,N,T,wrappers,cont,nact", [(E.CARTPOLE, 1024, 700, False, False, 2), (E.PENDULUM, 256, 600, True, True, 1),
                                                       (E.PENDULUM, 256, 600, False, True, 1), (E.ACROBOT, 128, 600, False, False, 3),
                                                       (E.MOUNTAINCAR, 256, 400, False, False, 3),
                                                       (E.MOUNTAINCAR_CONT, 128, 1100, True, True, 1)])
def test_checker_trajectories_identical_with_product_and_independent_trig(kind, N, T, wrappers, cont, nact):
    """The GPU parity tests compare the CUDA kernels (det_sincos.h) with the checker in TRIG_CR mode (libquadmath), i.e.
    with trigonometry the product does not share.  Here the checker itself runs both ways: every observation, reward,
    flag and the final fp64 state (incl. the wrappers' running statistics) must be identical bit for bit."""
    rng = np.random.default_rng(kind * 7 + 1)
    a = E.CVecEnv(kind, N, wrappers=wrappers, trig=E.TRIG_DET)
    b = E.CVecEnv(kind, N, wrappers=wrappers, trig=E.TRIG_CR)
    oa, _ = a.reset(list(range(N))); ob, _ = b.reset(list(range(N)))
    assert np.array_equal(oa, ob)
    for t in range(T):
        act = rng.normal(0, 1, (N, 1)).astype(np.float32) if cont else rng.integers(0, nact, N)
        ra, rb = a.step(act), b.step(act)
        for u, v in zip(ra[:4], rb[:4]):
            assert np.array_equal(u, v), (t,)
    assert np.array_equal(a.phys(), b.phys())


def test_det_and_libm_trajectories_agree_on_float32_observations():
    """Measures what the sincos choice changes at the interface the policy sees."""
    N, T = 64, 600
    a_env = E.CVecEnv(E.CARTPOLE, N, trig=E.TRIG_DET)
    b_env = E.CVecEnv(E.CARTPOLE, N, trig=E.TRIG_LIBM)
    a_env.reset(list(range(N))); b_env.reset(list(range(N)))
    rng = np.random.default_rng(11)
    obs_mismatch = done_mismatch = 0
    for t in range(T):
        a = rng.integers(0, 2, N)
        ra, rb = a_env.step(a), b_env.step(a)
        obs_mismatch += int((ra[0] != rb[0]).sum())
        done_mismatch += int((ra[2] != rb[2]).sum())
    assert done_mismatch == 0
    assert obs_mismatch <= 4       # fp64 last-bit differences almost never cross a float32 rounding boundary


def test_mountaincar_restatement_behaviour():
    """MountainCar-v0: reward -1 per step, left wall stops the car, full throttle right never reaches the flag from
    rest (the classic under-powered car), the energy-pumping policy does within the 200-step TimeLimit."""
    env = G.make_env("MountainCar-v0", False)
    obs, _ = env.reset(seed=0)
    assert obs.dtype == np.float32 and -0.6 <= obs[0] <= -0.4 and obs[1] == 0
    for t in range(200):
        obs, r, term, trunc, info = env.step(2)
        assert r == -1.0 and not term
    assert trunc and info["episode"]["l"] == 200
    env = G.make_env("MountainCar-v0", False)
    obs, _ = env.reset(seed=0)
    for t in range(200):
        obs, r, term, trunc, info = env.step(2 if obs[1] >= 0 else 0)
        if term:
            break
    assert term and obs[0] >= 0.5 and 80 < t < 200
    env = G.make_env("MountainCar-v0", False)
    obs, _ = env.reset(seed=1)
    for t in range(60):
        obs, *_ = env.step(0)
    assert obs[0] >= -1.2 and (obs[0] > -1.2 or obs[1] >= 0)


def test_acrobot_restatement_behaviour():
    """Acrobot-v1: reward -1 until the tip passes the bar (then 0 and terminated), angles stay wrapped to [-pi, pi],
    velocities bounded by 4 pi / 9 pi, torque-free motion from rest at the bottom stays at the bottom, an energy-pumping
    controller swings up well inside the 500-step limit while zero torque never does."""
    env = G.make_env("Acrobot-v1", False)
    env.reset(seed=3)
    inner = env.env.env
    inner.state = np.zeros(4)
    for _ in range(20):
        obs, r, term, trunc, _ = env.step(1)
        assert r == -1.0 and not term
    np.testing.assert_allclose(inner.state, 0.0, atol=1e-12)
    env.reset(seed=4)
    steps, term = 0, False
    while not term and steps < 500:
        a = 2 if inner.state[3] > 0 else 0          # torque along the actuated joint's velocity pumps energy in
        obs, r, term, trunc, _ = env.step(a)
        steps += 1
        assert abs(inner.state[0]) <= np.pi and abs(inner.state[1]) <= np.pi
        assert abs(inner.state[2]) <= 4 * np.pi and abs(inner.state[3]) <= 9 * np.pi
        assert obs.dtype == np.float32 and obs.shape == (6,)
        np.testing.assert_allclose(obs[0] ** 2 + obs[1] ** 2, 1.0, atol=1e-6)
    assert term and r == 0.0 and steps < 200, steps     # good Acrobot-v1 policies finish in roughly 60-120 steps
    env.reset(seed=5)
    for t in range(500):
        obs, r, term, trunc, info = env.step(1)
        assert not term
    assert trunc and info["episode"]["l"] == 500 and info["episode"]["r"] == -500.0


def test_mountaincar_continuous_restatement_behaviour():
    """MountainCarContinuous-v0: reward -0.1 a^2 per step (+100 at the flag), the state is float32 after every step, full
    throttle from the valley floor does not reach the flag but rocking with the velocity does, TimeLimit 999."""
    env = G.make_env("MountainCarContinuous-v0", False)
    env.reset(seed=0)
    inner = env.env.env
    for _ in range(999):
        obs, r, term, trunc, info = env.step(np.array([1.0], np.float32))
        assert inner.state.dtype == np.float32 and not term
        assert abs(r + 0.1) < 1e-15
    assert trunc and info["episode"]["l"] == 999
    env.reset(seed=1)
    steps, term = 0, False
    while not term and steps < 999:
        a = 1.0 if inner.state[1] >= 0 else -1.0
        obs, r, term, trunc, _ = env.step(np.array([a], np.float32))
        steps += 1
    assert term and steps < 300 and abs(r - 99.9) < 1e-12, (steps, r)
    # out-of-range actions: the env clips the force but charges the raw action (no ClipAction without the wrapper stack)
    env.reset(seed=2)
    _, r, _, _, _ = env.step(np.array([3.0], np.float32))
    assert abs(r + 0.9) < 1e-12
    wenv = G.make_env("MountainCarContinuous-v0", True)
    wenv.reset(seed=2)
    w_inner = wenv
    while hasattr(w_inner, "env"):
        w_inner = w_inner.env
    _, r_w, _, _, _ = wenv.step(np.array([3.0], np.float32))
    np.testing.assert_allclose(w_inner.state, inner.state)           # same force either way
