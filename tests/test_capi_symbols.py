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


def test_ctypes_mirror_matches_the_header(tmp_path):
    """struct sizes and field offsets of the ctypes mirror == what a C compiler sees in the header"""
    import subprocess
    from strawboat_b200 import _capi
    structs = {"sb_page_meta": _capi.PageMeta, "sb_leaf": _capi.Leaf, "sb_column_in": _capi.ColumnIn, "sb_column_out": _capi.ColumnOut,
               "sb_stats": _capi.Stats, "sb_write_options": _capi.WriteOptions, "sb_leaf_array": _capi.LeafArray,
               "sb_encoded_column": _capi.EncodedColumn, "sb_page_info": _capi.PageInfo, "sb_gather_stats": _capi.GatherStats, "sb_out_buffers": _capi.OutBuffers, "sb_column_sizes": _capi.ColumnSizes, "sb_nested_level": _capi.NestedLevel,
               "sb_field": _capi.Field, "struct ArrowArray": _capi.ArrowArray, "struct ArrowSchema": _capi.ArrowSchema}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "strawboat_b200.h"', 'int main(void) {']
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    got = dict(line.rsplit(None, 1) for line in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"
