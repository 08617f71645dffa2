"""The C-ABI library loads without a GPU and exports every function include/mono_abi.h declares;
without a CUDA device the one entry point that needs it fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "mono_abi.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mono_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_boundary():
    names = declared_functions()
    for must in ("mono_ctx_create", "mono_ode_step", "mono_pde_step", "mono_split_step", "mono_set_halo", "mono_comm_init",
                 "mono_stim_add", "mono_ksp_info"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from beat_b200 import _lib

    lib = _lib.load_library()
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.mono_abi_version() == 1


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from beat_b200 import _lib

    with pytest.raises(RuntimeError, match="(?i)cuda|device"):
        _lib.Context(0)


def test_binding_table_covers_the_header():
    from beat_b200 import _lib

    assert sorted(_lib.SIGNATURES) == declared_functions()


def test_null_context_is_refused_not_dereferenced():
    """Every entry point that takes a context answers MONO_E_INVALID to NULL (no GPU needed to check that)."""
    from beat_b200 import _lib

    lib = _lib.load_library()
    checked = 0
    for name, (res, args) in _lib.SIGNATURES.items():
        if not args or args[0] is not ctypes.c_void_p or res is not ctypes.c_int or name in ("mono_ctx_destroy", "mono_host_free"):
            continue
        zeros = [None if a is ctypes.c_void_p or hasattr(a, "contents") else 0 for a in args]
        assert getattr(lib, name)(*zeros) == -1, name
        assert b"NULL" in lib.mono_last_error(None), name
        checked += 1
    assert checked >= 50
    assert lib.mono_ctx_destroy(None) == 0
