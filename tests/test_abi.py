"""CPU-side checks of the C-ABI library: it loads and exports every symbol
include/aur_ppo.h declares.  No compute calls (no GPU here)."""
import ctypes
import os

import pytest

from aur_ppo_b200 import _lib


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.lib()


def test_exports_every_declared_symbol(built):
    names = _lib.declared_symbols()
    assert "aur_gae_f32" in names and len(names) >= 6
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/aur_ppo.h but not exported"


def test_abi_version_and_error_string(built):
    assert built.aur_abi_version() == 3
    assert isinstance(built.aur_last_error(), bytes)


def test_argument_errors_do_not_need_a_gpu(built):
    rc = built.aur_gae_f32(-1, 4, None, None, None, None, None, 0.99, 0.95, 1, None, None, None)
    assert rc == -1 and b"negative" in built.aur_last_error()
    rc = built.aur_gae_f32(4, 4, None, None, None, None, None, 0.99, 0.95, 1, None, None, None)
    assert rc == -1 and b"null" in built.aur_last_error()
    assert built.aur_gae_f32(0, 4, None, None, None, None, None, 0.99, 0.95, 1, None, None, None) == 0


def test_wrappers_refuse_cpu_tensors(built):
    import torch
    from aur_ppo_b200 import kernels
    z = torch.zeros(4, 4)
    with pytest.raises(_lib.AurError):
        kernels.gae(z, z, z, torch.zeros(4), torch.zeros(4), 0.99, 0.95)
