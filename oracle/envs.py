"""ctypes front-end of the C env checker (oracle/envs.c) -- TEST INFRASTRUCTURE ONLY.

`CVecEnv` mirrors gym.vector.SyncVectorEnv as the reference drives it
(ppo.py:66-68,110,188).  Seeding follows gym/utils/seeding.py: env i gets
PCG64(SeedSequence(seed_i)); the 128-bit (state, inc) pair is expanded by the
installed NumPy and handed to C, which then runs the PCG64 stream itself.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

CARTPOLE, PENDULUM, MOUNTAINCAR, ACROBOT, MOUNTAINCAR_CONT = 0, 1, 2, 3, 4
OBS_DIM = {CARTPOLE: 4, PENDULUM: 3, MOUNTAINCAR: 2, ACROBOT: 6, MOUNTAINCAR_CONT: 2}
PHYS_DIM = {CARTPOLE: 4, PENDULUM: 2, MOUNTAINCAR: 2, ACROBOT: 4, MOUNTAINCAR_CONT: 2}
CONTINUOUS = (PENDULUM, MOUNTAINCAR_CONT)
TRIG_LIBM, TRIG_DET, TRIG_CR = 0, 1, 2       # TRIG_CR: libquadmath rounded once to double, independent of the product


def build() -> str:
    subprocess.run(["make", "-s", "-C", _HERE, "liborc.so"], check=True)
    return os.path.join(_HERE, "liborc.so")


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liborc.so")
        src = os.path.join(_HERE, "envs.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            build()
        L = ctypes.CDLL(path)
        L.orc_vec_create.restype = ctypes.c_void_p
        L.orc_vec_create.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_double]
        L.orc_vec_destroy.argtypes = [ctypes.c_void_p]
        L.orc_vec_reset.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_vec_step.argtypes = [ctypes.c_void_p] * 8
        L.orc_vec_get_phys.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.orc_vec_get_norm.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.orc_sincos.argtypes = [ctypes.c_int64] + [ctypes.c_void_p] * 3
        L.orc_libm_sincos.argtypes = [ctypes.c_int64] + [ctypes.c_void_p] * 3
        L.orc_cr_sincos.argtypes = [ctypes.c_int64] + [ctypes.c_void_p] * 3
        L.orc_pcg64_doubles.argtypes = [ctypes.c_uint64] * 4 + [ctypes.c_int64, ctypes.c_void_p]
        _LIB = L
    return _LIB


def pcg64_seed_states(seeds) -> np.ndarray:
    """[n,4] uint64: state_hi, state_lo, inc_hi, inc_lo of PCG64(SeedSequence(seed))."""
    out = np.empty((len(seeds), 4), dtype=np.uint64)
    m = (1 << 64) - 1
    for i, s in enumerate(seeds):
        st = np.random.PCG64(np.random.SeedSequence(int(s))).state["state"]
        out[i] = [st["state"] >> 64, st["state"] & m, st["inc"] >> 64, st["inc"] & m]
    return out


def det_sincos(x: np.ndarray):
    x = np.ascontiguousarray(x, dtype=np.float64)
    s, c = np.empty_like(x), np.empty_like(x)
    lib().orc_sincos(x.size, x.ctypes.data, s.ctypes.data, c.ctypes.data)
    return s, c


def libm_sincos(x: np.ndarray):
    x = np.ascontiguousarray(x, dtype=np.float64)
    s, c = np.empty_like(x), np.empty_like(x)
    lib().orc_libm_sincos(x.size, x.ctypes.data, s.ctypes.data, c.ctypes.data)
    return s, c


def cr_sincos(x: np.ndarray):
    """Correctly rounded sin / cos (113-bit libquadmath evaluation rounded once): the checker's independent reference."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    s, c = np.empty_like(x), np.empty_like(x)
    lib().orc_cr_sincos(x.size, x.ctypes.data, s.ctypes.data, c.ctypes.data)
    return s, c


def det_sincos_scalar(x: float):
    s, c = det_sincos(np.array([x]))
    return float(s[0]), float(c[0])


class CVecEnv:
    def __init__(self, kind: int, num_envs: int, wrappers: bool = False, trig: int = TRIG_DET, gamma: float = 0.99):
        self.kind, self.num_envs = kind, num_envs
        self.obs_dim = OBS_DIM[kind]
        self._h = lib().orc_vec_create(kind, num_envs, int(wrappers), trig, gamma)
        n = num_envs
        self._obs = np.zeros((n, self.obs_dim), np.float32)
        self._rew = np.zeros(n, np.float64)
        self._term = np.zeros(n, np.uint8)
        self._trunc = np.zeros(n, np.uint8)
        self._fret = np.zeros(n, np.float32)
        self._flen = np.zeros(n, np.int32)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_vec_destroy(self._h)
            self._h = None

    def reset(self, seed):
        st = pcg64_seed_states(seed)
        lib().orc_vec_reset(self._h, st.ctypes.data, self._obs.ctypes.data)
        return self._obs.copy(), {}

    def step(self, actions):
        if self.kind not in CONTINUOUS:
            a = np.ascontiguousarray(actions, dtype=np.int32)
        else:
            a = np.ascontiguousarray(np.asarray(actions, dtype=np.float32).reshape(self.num_envs))
        lib().orc_vec_step(self._h, a.ctypes.data, self._obs.ctypes.data, self._rew.ctypes.data,
                           self._term.ctypes.data, self._trunc.ctypes.data,
                           self._fret.ctypes.data, self._flen.ctypes.data)
        info = {}
        if self._flen.any():
            info["final_info"] = [({"episode": {"r": self._fret[i], "l": int(self._flen[i])}} if self._flen[i] else None)
                                  for i in range(self.num_envs)]
        return (self._obs.copy(), self._rew.copy(), self._term.astype(bool), self._trunc.astype(bool), info)

    def phys(self) -> np.ndarray:
        out = np.zeros((self.num_envs, PHYS_DIM[self.kind]), np.float64)
        lib().orc_vec_get_phys(self._h, out.ctypes.data)
        return out[:, :PHYS_DIM[self.kind]]

    def norm_stats(self) -> np.ndarray:
        out = np.zeros((self.num_envs, 2 * self.obs_dim + 5), np.float64)
        lib().orc_vec_get_norm(self._h, out.ctypes.data)
        return out
