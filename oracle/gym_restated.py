"""CPU restatement of the gym==0.26.2 pieces the reference's PPO path calls.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Never by the product path.

PARITY UNPINNED at this boundary: gym 0.26.2 (reference pin:
src/environment.yml:504) is a third-party dependency that is neither vendored
under /root/reference nor installable here, and the reference has no tests or
golden vectors for env transitions.  This file restates the published
algorithm of

  gym/envs/classic_control/cartpole.py    (CartPoleEnv.step / reset)
  gym/envs/classic_control/pendulum.py    (PendulumEnv.step / reset / _get_obs)
  gym/envs/classic_control/mountain_car.py, acrobot.py (step / reset; rk4, wrap, bound)
  gym/wrappers/time_limit.py              (TimeLimit)
  gym/wrappers/record_episode_statistics.py
  gym/wrappers/clip_action.py, normalize.py, transform_observation.py,
  gym/wrappers/transform_reward.py
  gym/vector/sync_vector_env.py           (reset(seed=list), step + autoreset)
  gym/utils/seeding.py                    (np_random -> PCG64(SeedSequence(seed)))

anchored on the reference's own call sites: ppo.py:66-68 (SyncVectorEnv of
make_env thunks), ppo.py:87-97 (wrapper stack), ppo.py:110 (step), ppo.py:188
(reset(seed=list(range(num_envs)))).  The one externally known answer,
CartPole reset(seed=0) -> [0.01369617, -0.02302133, -0.04590265, -0.04834723],
is checked in tests/test_oracle_envs.py.

Arithmetic notes (reference pins NumPy 1.24.3, environment.yml:277):
  * `x ** 2` on Python / NumPy float64 scalars is restated as `x * x`
    (SURVEY.md section 7.2 item 1).
  * NumPy 1.24 value-based promotion makes `python_float * np.float32 scalar`
    a float64; this container has NumPy 2.x (NEP 50) where it would stay
    float32, so the promotions are written out explicitly with float().
  * `trig="libm"` uses math.sin/math.cos (what gym does); `trig="det"` uses
    the deterministic routine shared with the GPU kernels through the C
    checker (oracle/envs.c -> orc_sincos).
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np


def np_random(seed: Optional[int]):
    """gym/utils/seeding.py: Generator(PCG64(SeedSequence(seed)))."""
    seed_seq = np.random.SeedSequence(seed)
    return np.random.Generator(np.random.PCG64(seed_seq))


def _libm_sincos(x: float) -> Tuple[float, float]:
    return math.sin(x), math.cos(x)


class CartPoleEnv:
    """gym CartPole-v1 physics (cartpole.py), euler integrator."""

    def __init__(self, sincos: Callable[[float], Tuple[float, float]] = _libm_sincos):
        self.gravity = 9.8
        self.masscart = 1.0
        self.masspole = 0.1
        self.total_mass = self.masspole + self.masscart
        self.length = 0.5
        self.polemass_length = self.masspole * self.length
        self.force_mag = 10.0
        self.tau = 0.02
        self.theta_threshold_radians = 12 * 2 * math.pi / 360
        self.x_threshold = 2.4
        self.state = None
        self.steps_beyond_terminated = None
        self.np_random = None
        self._sincos = sincos
        self.obs_dim = 4

    def reset(self, seed: Optional[int] = None):
        if seed is not None or self.np_random is None:
            self.np_random = np_random(seed)
        self.state = tuple(float(v) for v in self.np_random.uniform(low=-0.05, high=0.05, size=(4,)))
        self.steps_beyond_terminated = None
        return np.array(self.state, dtype=np.float32), {}

    def step(self, action):
        x, x_dot, theta, theta_dot = self.state
        force = self.force_mag if action == 1 else -self.force_mag
        sintheta, costheta = self._sincos(theta)
        temp = (force + self.polemass_length * (theta_dot * theta_dot) * sintheta) / self.total_mass
        thetaacc = (self.gravity * sintheta - costheta * temp) / (
            self.length * (4.0 / 3.0 - self.masspole * (costheta * costheta) / self.total_mass)
        )
        xacc = temp - self.polemass_length * thetaacc * costheta / self.total_mass
        x = x + self.tau * x_dot
        x_dot = x_dot + self.tau * xacc
        theta = theta + self.tau * theta_dot
        theta_dot = theta_dot + self.tau * thetaacc
        self.state = (x, x_dot, theta, theta_dot)
        terminated = bool(
            x < -self.x_threshold
            or x > self.x_threshold
            or theta < -self.theta_threshold_radians
            or theta > self.theta_threshold_radians
        )
        if not terminated:
            reward = 1.0
        elif self.steps_beyond_terminated is None:
            self.steps_beyond_terminated = 0
            reward = 1.0
        else:
            self.steps_beyond_terminated += 1
            reward = 0.0
        return np.array(self.state, dtype=np.float32), reward, terminated, False, {}


class MountainCarEnv:
    """gym MountainCar-v0 (mountain_car.py): fp64 state, reward -1 per step, goal at position 0.5."""

    def __init__(self, sincos: Callable[[float], Tuple[float, float]] = _libm_sincos):
        self.min_position = -1.2
        self.max_position = 0.6
        self.max_speed = 0.07
        self.goal_position = 0.5
        self.goal_velocity = 0
        self.force = 0.001
        self.gravity = 0.0025
        self.state = None
        self.np_random = None
        self._sincos = sincos
        self.obs_dim = 2

    def reset(self, seed: Optional[int] = None):
        if seed is not None or self.np_random is None:
            self.np_random = np_random(seed)
        self.state = np.array([self.np_random.uniform(low=-0.6, high=-0.4), 0])
        return np.array(self.state, dtype=np.float32), {}

    def step(self, action):
        position, velocity = self.state
        velocity += (action - 1) * self.force + self._sincos(3 * position)[1] * (-self.gravity)
        velocity = np.clip(velocity, -self.max_speed, self.max_speed)
        position += velocity
        position = np.clip(position, self.min_position, self.max_position)
        if position == self.min_position and velocity < 0:
            velocity = 0
        terminated = bool(position >= self.goal_position and velocity >= self.goal_velocity)
        reward = -1.0
        self.state = (position, velocity)
        return np.array(self.state, dtype=np.float32), reward, terminated, False, {}


class AcrobotEnv:
    """gym Acrobot-v1 (acrobot.py): two-link pendulum, "book" dynamics, torque noise 0, one RK4 step of dt = 0.2.

    Written the way the gym source reads (NumPy fp64 arrays through rk4 / _dsdt / wrap / bound) so that the C checker
    (oracle/envs.c, scalar code) is an independent restatement.  One deliberate difference: right after reset the state
    is a float32 array and NumPy would evaluate the observation's sin / cos in float32; here they are evaluated in
    fp64 and rounded (at most one ulp apart, rarely), because NumPy's float32 trig kernels are platform specific."""

    dt = 0.2
    LINK_LENGTH_1 = 1.0
    LINK_LENGTH_2 = 1.0
    LINK_MASS_1 = 1.0
    LINK_MASS_2 = 1.0
    LINK_COM_POS_1 = 0.5
    LINK_COM_POS_2 = 0.5
    LINK_MOI = 1.0
    MAX_VEL_1 = 4 * math.pi
    MAX_VEL_2 = 9 * math.pi
    AVAIL_TORQUE = [-1.0, 0.0, +1]

    def __init__(self, sincos: Callable[[float], Tuple[float, float]] = _libm_sincos):
        self.state = None
        self.np_random = None
        self._sincos = sincos
        self.obs_dim = 6

    def _sin(self, x):
        return self._sincos(float(x))[0]

    def _cos(self, x):
        return self._sincos(float(x))[1]

    def reset(self, seed: Optional[int] = None):
        if seed is not None or self.np_random is None:
            self.np_random = np_random(seed)
        self.state = self.np_random.uniform(low=-0.1, high=0.1, size=(4,)).astype(np.float32)
        return self._get_ob(), {}

    def step(self, a):
        s = self.state
        torque = self.AVAIL_TORQUE[int(a)]
        s_augmented = np.append(np.asarray(s, dtype=np.float64), torque)
        ns = self._rk4(s_augmented, [0, self.dt])
        ns[0] = self._wrap(ns[0], -math.pi, math.pi)
        ns[1] = self._wrap(ns[1], -math.pi, math.pi)
        ns[2] = min(max(ns[2], -self.MAX_VEL_1), self.MAX_VEL_1)
        ns[3] = min(max(ns[3], -self.MAX_VEL_2), self.MAX_VEL_2)
        self.state = ns
        terminated = self._terminal()
        reward = -1.0 if not terminated else 0.0
        return self._get_ob(), reward, terminated, False, {}

    def _get_ob(self):
        s = self.state
        return np.array([self._cos(s[0]), self._sin(s[0]), self._cos(s[1]), self._sin(s[1]), s[2], s[3]], dtype=np.float32)

    def _terminal(self):
        s = self.state
        return bool(-self._cos(s[0]) - self._cos(float(s[1]) + float(s[0])) > 1.0)

    def _dsdt(self, s_augmented):
        m1, m2 = self.LINK_MASS_1, self.LINK_MASS_2
        l1 = self.LINK_LENGTH_1
        lc1, lc2 = self.LINK_COM_POS_1, self.LINK_COM_POS_2
        I1 = I2 = self.LINK_MOI
        g = 9.8
        pi = math.pi
        a = float(s_augmented[-1])
        theta1, theta2, dtheta1, dtheta2 = (float(v) for v in s_augmented[:-1])
        sin, cos = self._sin, self._cos
        d1 = m1 * (lc1 * lc1) + m2 * (l1 * l1 + lc2 * lc2 + 2 * l1 * lc2 * cos(theta2)) + I1 + I2
        d2 = m2 * (lc2 * lc2 + l1 * lc2 * cos(theta2)) + I2
        phi2 = m2 * lc2 * g * cos(theta1 + theta2 - pi / 2.0)
        phi1 = (-m2 * l1 * lc2 * (dtheta2 * dtheta2) * sin(theta2) - 2 * m2 * l1 * lc2 * dtheta2 * dtheta1 * sin(theta2)
                + (m1 * lc1 + m2 * l1) * g * cos(theta1 - pi / 2) + phi2)
        ddtheta2 = (a + d2 / d1 * phi1 - m2 * l1 * lc2 * (dtheta1 * dtheta1) * sin(theta2) - phi2) / \
                   (m2 * (lc2 * lc2) + I2 - (d2 * d2) / d1)
        ddtheta1 = -(d2 * ddtheta2 + phi1) / d1
        return dtheta1, dtheta2, ddtheta1, ddtheta2, 0.0

    def _rk4(self, y0, t):
        yout = np.zeros((len(t), len(y0)), np.float64)
        yout[0] = y0
        for i in np.arange(len(t) - 1):
            this = t[i]
            dt = t[i + 1] - this
            dt2 = dt / 2.0
            y0 = yout[i]
            k1 = np.asarray(self._dsdt(y0))
            k2 = np.asarray(self._dsdt(y0 + dt2 * k1))
            k3 = np.asarray(self._dsdt(y0 + dt2 * k2))
            k4 = np.asarray(self._dsdt(y0 + dt * k3))
            yout[i + 1] = y0 + dt / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
        return yout[-1][:4]

    @staticmethod
    def _wrap(x, m, M):
        diff = M - m
        while x > M:
            x = x - diff
        while x < m:
            x = x + diff
        return x


class Continuous_MountainCarEnv:
    """gym MountainCarContinuous-v0 (continuous_mountain_car.py): the state is re-created as a float32 array every step,
    the arithmetic in between is float64 (NumPy 1.24: a float32 scalar combined with a Python float or int gives float64;
    written out with float() because this container has NumPy 2)."""

    def __init__(self, sincos: Callable[[float], Tuple[float, float]] = _libm_sincos):
        self.min_action = -1.0
        self.max_action = 1.0
        self.min_position = -1.2
        self.max_position = 0.6
        self.max_speed = 0.07
        self.goal_position = 0.45
        self.goal_velocity = 0
        self.power = 0.0015
        self.state = None
        self.np_random = None
        self._sincos = sincos
        self.obs_dim = 2

    def reset(self, seed: Optional[int] = None):
        if seed is not None or self.np_random is None:
            self.np_random = np_random(seed)
        self.state = np.array([self.np_random.uniform(low=-0.6, high=-0.4), 0])
        return np.array(self.state, dtype=np.float32), {}

    def step(self, action):
        position = float(self.state[0])
        velocity = float(self.state[1])
        a0 = np.float32(np.asarray(action, dtype=np.float32).reshape(-1)[0])
        force = min(max(a0, self.min_action), self.max_action)
        velocity += float(force) * self.power - 0.0025 * self._sincos(3 * position)[1]
        if velocity > self.max_speed:
            velocity = self.max_speed
        if velocity < -self.max_speed:
            velocity = -self.max_speed
        position += velocity
        if position > self.max_position:
            position = self.max_position
        if position < self.min_position:
            position = self.min_position
        if position == self.min_position and velocity < 0:
            velocity = 0
        terminated = bool(position >= self.goal_position and velocity >= self.goal_velocity)
        reward = 0
        if terminated:
            reward = 100.0
        reward -= math.pow(float(a0), 2) * 0.1
        self.state = np.array([position, velocity], dtype=np.float32)
        return self.state, reward, terminated, False, {}


def angle_normalize(x: float) -> float:
    # ((x + pi) % (2 pi)) - pi ; Python float % == NumPy float64 % (fmod + sign fix)
    return ((x + math.pi) % (2 * math.pi)) - math.pi


class PendulumEnv:
    """gym Pendulum-v1 physics (pendulum.py), g=10."""

    def __init__(self, sincos: Callable[[float], Tuple[float, float]] = _libm_sincos):
        self.max_speed = 8.0
        self.max_torque = 2.0
        self.dt = 0.05
        self.g = 10.0
        self.m = 1.0
        self.l = 1.0
        self.state = None
        self.np_random = None
        self._sincos = sincos
        self.obs_dim = 3

    def _get_obs(self):
        theta, thetadot = self.state
        s, c = self._sincos(theta)
        return np.array([c, s, thetadot], dtype=np.float32)

    def reset(self, seed: Optional[int] = None):
        if seed is not None or self.np_random is None:
            self.np_random = np_random(seed)
        high = np.array([math.pi, 1.0])
        st = self.np_random.uniform(low=-high, high=high)
        self.state = (float(st[0]), float(st[1]))
        return self._get_obs(), {}

    def step(self, u):
        th, thdot = self.state
        g, m, l, dt = self.g, self.m, self.l, self.dt
        u32 = np.float32(np.clip(np.asarray(u, dtype=np.float32), -self.max_torque, self.max_torque).reshape(-1)[0])
        usq32 = np.float32(u32 * u32)               # u**2 stays float32
        an = angle_normalize(th)
        costs = an * an + 0.1 * (thdot * thdot) + 0.001 * float(usq32)
        s, _ = self._sincos(th)
        newthdot = thdot + (3 * g / (2 * l) * s + 3.0 / (m * (l * l)) * float(u32)) * dt
        newthdot = min(max(newthdot, -self.max_speed), self.max_speed)
        newth = th + newthdot * dt
        self.state = (newth, newthdot)
        return self._get_obs(), -costs, False, False, {}


class TimeLimit:
    def __init__(self, env, max_episode_steps: int):
        self.env = env
        self._max_episode_steps = max_episode_steps
        self._elapsed_steps = 0

    def reset(self, seed=None):
        self._elapsed_steps = 0
        return self.env.reset(seed=seed)

    def step(self, action):
        obs, rew, terminated, truncated, info = self.env.step(action)
        self._elapsed_steps += 1
        if self._elapsed_steps >= self._max_episode_steps:
            truncated = True
        return obs, rew, terminated, truncated, info


class RecordEpisodeStatistics:
    """Single-env form: float32 return accumulator, int32 length."""

    def __init__(self, env):
        self.env = env
        self.episode_return = np.float32(0.0)
        self.episode_length = 0

    def reset(self, seed=None):
        out = self.env.reset(seed=seed)
        self.episode_return = np.float32(0.0)
        self.episode_length = 0
        return out

    def step(self, action):
        obs, rew, terminated, truncated, info = self.env.step(action)
        self.episode_return = np.float32(self.episode_return + np.float32(rew))
        self.episode_length += 1
        if terminated or truncated:
            info = dict(info)
            info["episode"] = {"r": self.episode_return, "l": self.episode_length}
            self.episode_return = np.float32(0.0)
            self.episode_length = 0
        return obs, rew, terminated, truncated, info


class ClipAction:
    def __init__(self, env, low: float, high: float):
        self.env, self.low, self.high = env, np.float32(low), np.float32(high)

    def reset(self, seed=None):
        return self.env.reset(seed=seed)

    def step(self, action):
        return self.env.step(np.clip(np.asarray(action, dtype=np.float32), self.low, self.high))


class RunningMeanStd:
    """gym/wrappers/normalize.py with batch_count == 1 (one env per wrapper)."""

    def __init__(self, shape=()):
        self.mean = np.zeros(shape, "float64")
        self.var = np.ones(shape, "float64")
        self.count = 1e-4

    def update1(self, x):
        # batch_mean = x, batch_var = 0, batch_count = 1
        delta = np.asarray(x, dtype=np.float64) - self.mean
        tot_count = self.count + 1
        new_mean = self.mean + delta * 1 / tot_count
        m_a = self.var * self.count
        m_b = 0.0 * 1
        M2 = m_a + m_b + np.square(delta) * self.count * 1 / tot_count
        self.mean, self.var, self.count = new_mean, M2 / tot_count, tot_count


class NormalizeObservation:
    def __init__(self, env, shape, epsilon=1e-8):
        self.env, self.epsilon = env, epsilon
        self.obs_rms = RunningMeanStd(shape)

    def normalize(self, obs):
        self.obs_rms.update1(obs)
        return (np.asarray(obs, dtype=np.float64) - self.obs_rms.mean) / np.sqrt(self.obs_rms.var + self.epsilon)

    def reset(self, seed=None):
        obs, info = self.env.reset(seed=seed)
        return self.normalize(obs), info

    def step(self, action):
        obs, rew, terminated, truncated, info = self.env.step(action)
        return self.normalize(obs), rew, terminated, truncated, info


class ClipObservation:
    """TransformObservation(env, lambda obs: np.clip(obs, -10, 10)) (ppo.py:95)."""

    def __init__(self, env):
        self.env = env

    def reset(self, seed=None):
        obs, info = self.env.reset(seed=seed)
        return np.clip(obs, -10, 10), info

    def step(self, action):
        obs, rew, terminated, truncated, info = self.env.step(action)
        return np.clip(obs, -10, 10), rew, terminated, truncated, info


class NormalizeReward:
    def __init__(self, env, gamma=0.99, epsilon=1e-8):
        self.env, self.gamma, self.epsilon = env, gamma, epsilon
        self.return_rms = RunningMeanStd(())
        self.returns = 0.0

    def reset(self, seed=None):
        return self.env.reset(seed=seed)

    def step(self, action):
        obs, rew, terminated, truncated, info = self.env.step(action)
        self.returns = self.returns * self.gamma + float(rew)
        self.return_rms.update1(self.returns)
        rew = float(rew) / math.sqrt(float(self.return_rms.var) + self.epsilon)
        if terminated or truncated:
            self.returns = 0.0
        return obs, rew, terminated, truncated, info


class ClipReward:
    """TransformReward(env, lambda r: np.clip(r, -10, 10)) (ppo.py:97)."""

    def __init__(self, env):
        self.env = env

    def reset(self, seed=None):
        return self.env.reset(seed=seed)

    def step(self, action):
        obs, rew, terminated, truncated, info = self.env.step(action)
        return obs, float(min(max(rew, -10.0), 10.0)), terminated, truncated, info


def make_env(gym_id: str, continuous: bool, sincos=_libm_sincos):
    """The reference's make_env thunk (ppo.py:85-99) without video capture."""
    if gym_id == "CartPole-v1":
        env = TimeLimit(CartPoleEnv(sincos), 500)
        obs_shape = (4,)
    elif gym_id == "Pendulum-v1":
        env = TimeLimit(PendulumEnv(sincos), 200)
        obs_shape = (3,)
    elif gym_id == "MountainCar-v0":
        env = TimeLimit(MountainCarEnv(sincos), 200)
        obs_shape = (2,)
    elif gym_id == "Acrobot-v1":
        env = TimeLimit(AcrobotEnv(sincos), 500)
        obs_shape = (6,)
    elif gym_id == "MountainCarContinuous-v0":
        env = TimeLimit(Continuous_MountainCarEnv(sincos), 999)
        obs_shape = (2,)
    else:
        raise ValueError(f"unsupported gym_id {gym_id!r}")
    env = RecordEpisodeStatistics(env)
    if continuous:
        bound = 1.0 if gym_id == "MountainCarContinuous-v0" else 2.0     # env.action_space.low / high
        env = ClipAction(env, -bound, bound)
        env = NormalizeObservation(env, obs_shape)
        env = ClipObservation(env)
        env = NormalizeReward(env)
        env = ClipReward(env)
    return env


class SyncVectorEnv:
    """gym/vector/sync_vector_env.py: serial stepping, autoreset on
    terminated-or-truncated, float32 observation buffer, float64 rewards."""

    def __init__(self, envs: Sequence, obs_dim: int):
        self.envs = list(envs)
        self.num_envs = len(self.envs)
        self.observations = np.zeros((self.num_envs, obs_dim), dtype=np.float32)
        self._rewards = np.zeros((self.num_envs,), dtype=np.float64)
        self._terminateds = np.zeros((self.num_envs,), dtype=np.bool_)
        self._truncateds = np.zeros((self.num_envs,), dtype=np.bool_)

    def reset(self, seed: Optional[List[int]] = None):
        if seed is None:
            seed = [None] * self.num_envs
        for i, (env, s) in enumerate(zip(self.envs, seed)):
            obs, _ = env.reset(seed=s)
            self.observations[i] = obs
        return np.copy(self.observations), {}

    def step(self, actions):
        final_info = [None] * self.num_envs
        any_final = False
        for i, (env, action) in enumerate(zip(self.envs, actions)):
            obs, self._rewards[i], self._terminateds[i], self._truncateds[i], info = env.step(action)
            if self._terminateds[i] or self._truncateds[i]:
                final_info[i] = info
                any_final = True
                obs, _ = env.reset()
            self.observations[i] = obs
        infos = {"final_info": final_info} if any_final else {}
        return (np.copy(self.observations), np.copy(self._rewards), np.copy(self._terminateds),
                np.copy(self._truncateds), infos)
