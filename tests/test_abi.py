"""CPU: the C-ABI library loads and exports every symbol include/declip_b200.h declares (no compute calls)."""
import os
import re

from declip_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "declip_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(dc_[a-z0-9_]+)\s*\(", src))


def test_library_exports_every_declared_symbol():
    import ctypes
    lib = ctypes.CDLL(_lib.so_path())
    declared = _declared()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), "missing export: " + name


def test_ctypes_signatures_cover_header():
    assert _lib.missing_symbols() == []
    declared = _declared()
    assert declared <= set(_lib.SIGNATURES) | {"dc_launch_count"}, declared - set(_lib.SIGNATURES)


def test_version_and_error_slot_without_gpu():
    lib = _lib.load()
    assert lib.dc_version() == 100
    assert isinstance(_lib.last_error(), str)


def test_product_path_fails_loudly_without_cuda():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from declip_b200 import ops
    with pytest.raises(RuntimeError):
        ops.gemm(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))


def test_no_oracle_import_in_product():
    pkg = os.path.join(ROOT, "declip_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
