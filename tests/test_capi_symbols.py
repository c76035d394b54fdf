"""CPU: the C-ABI library loads and exports every symbol include/strawboat_b200.h declares;
without a GPU the product refuses to compute (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "strawboat_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sb_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported():
    from strawboat_b200 import _capi
    lib = ctypes.CDLL(_capi.LIB_PATH)
    names = header_functions()
    assert len(names) >= 10
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/strawboat_b200.h but not exported"
    assert sorted(_capi.EXPORTS) == names


def test_no_cpu_fallback():
    import torch
    import strawboat_b200 as sb
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(sb.StrawboatError) as e:
        sb.Context(0)
    assert e.value.code == sb._capi.SB_CUDA


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "strawboat_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                text = open(os.path.join(dirpath, f)).read()
                assert "sbo" not in re.findall(r"^\s*(?:import|from)\s+(\w+)", text, flags=re.M), f
                assert "sb_oracle" not in text and "libsb_oracle" not in text, f
