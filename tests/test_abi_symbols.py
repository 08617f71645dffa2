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
