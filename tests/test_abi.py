"""The C-ABI library loads and exports every symbol include/hyres_b200.h declares (no compute
calls: this runs without a GPU)."""
import os
import re

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "hyres_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hyres_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(build_lib):
    from hyres_b200 import _lib
    names = _declared()
    assert len(names) >= 35
    lib = _lib.lib()
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/hyres_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes prototype in _lib.SIGNATURES"
    assert sorted(_lib.SIGNATURES) == names
    assert _lib.missing_symbols() == []


def test_signatures_have_no_torch_types():
    src = open(os.path.join(ROOT, "include", "hyres_b200.h")).read()
    assert 'extern "C"' in src
    code = re.sub(r"/\*.*?\*/", "", src, flags=re.S)  # comments may mention PyTorch layouts
    assert "torch" not in code.lower() and "at::" not in code and "std::" not in code


def test_version_and_no_device_is_an_error(build_lib):
    import torch
    from hyres_b200 import _lib
    lib = _lib.lib()
    assert lib.hyres_version() >= 100
    if not torch.cuda.is_available():
        assert lib.hyres_device_check(0) != 0  # fails loudly: no CPU fallback behind compute calls
        assert lib.hyres_last_error()


def test_library_is_sm100a_only(build_lib):
    import subprocess
    from hyres_b200 import _lib
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs
