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


def test_header_is_plain_c_and_a_c_consumer_links(tmp_path, built):
    """The boundary is a C ABI: include/aur_ppo.h must compile as C99 (no C++, no torch types) and a C program must link against
    the shared library and run without a GPU (it only asks for the ABI version and provokes an argument error)."""
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "consumer.c"
    src.write_text('#include <stdio.h>\n#include <string.h>\n#include "include/aur_ppo.h"\n'
                   'int main(void) {\n'
                   '  if (aur_abi_version() != AUR_ABI_VERSION) return 1;\n'
                   '  if (aur_gae_f32(4, 8, 0, 0, 0, 0, 0, 0.99, 0.95, 1, 0, 0, 0) != AUR_ERR_ARG) return 2;\n'
                   '  if (strlen(aur_last_error()) == 0) return 3;\n'
                   '  if (aur_ppo_update_set_wide(2) != AUR_ERR_ARG || aur_ppo_update_set_wide(1) != 0) return 4;\n'
                   '  printf("abi %d\\n", aur_abi_version());\n  return 0;\n}\n')
    exe = tmp_path / "consumer"
    lib_dir = os.path.join(root, "aur_ppo_b200")
    subprocess.run([cc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", root, str(src), "-o", str(exe), "-L", lib_dir, "-laurppo",
                    "-Wl,-rpath," + lib_dir], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert r.stdout.strip() == f"abi {built.aur_abi_version()}"
