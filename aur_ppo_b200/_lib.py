"""ctypes binding of libaurppo.so (the C ABI declared in include/aur_ppo.h).

There is no fallback: if the library is missing or a call fails, an exception is
raised.  Build it with `python -c "import __graft_entry__ as g; g.build()"` or
`make -C aur_ppo_b200/csrc`.
"""
from __future__ import annotations

import ctypes
import os
import re
import subprocess
from ctypes import c_double, c_int, c_int32, c_int64, c_uint64, c_void_p, c_char_p

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libaurppo.so")
HEADER_PATH = os.path.join(os.path.dirname(_PKG), "include", "aur_ppo.h")

_lib = None


class AurError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into aur_ppo_b200/libaurppo.so (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", os.path.join(_PKG, "csrc")], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise AurError("building libaurppo.so failed")
    return LIB_PATH


def declared_symbols() -> list:
    """Every function include/aur_ppo.h declares (used by the export test)."""
    src = open(HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(aur_[a-z0-9_]+)\s*\(", src)))


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AurError(f"{LIB_PATH} is missing: the CUDA extension has not been built and there is no CPU "
                       "fallback (run __graft_entry__.build())")
    L = ctypes.CDLL(LIB_PATH)
    L.aur_abi_version.restype = c_int
    L.aur_last_error.restype = c_char_p
    L.aur_launch_count.restype = c_int64
    L.aur_launch_count_reset.restype = None
    L.aur_gae_f32.restype = c_int
    L.aur_gae_f32.argtypes = [c_int32, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_double,
                              c_int32, c_void_p, c_void_p, c_void_p]
    L.aur_gae_kernel_kind.restype = c_int
    L.aur_gae_kernel_kind.argtypes = [c_int32, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().aur_last_error().decode(errors="replace")
        raise AurError(f"{what} failed (rc={rc}): {msg}")
