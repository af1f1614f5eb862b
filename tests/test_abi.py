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


def test_split_k_heuristic_is_wave_aware():
    """Host logic of the weight-gradient GEMMs (csrc/internal.h choose_splits): never more waves x k-blocks than the
    old fill-the-machine-twice rule, and the layer shapes of the bench land on the expected factors."""
    from declip_b200 import _lib
    lib = _lib.load()

    def cost(tiles, kb, workers, s):
        kbps = -(-kb // s)
        eff = -(-kb // kbps)
        return -(-tiles * eff // workers) * (kbps + 3)

    def old(tiles, kb, workers):
        if tiles >= workers:
            return 1
        return max(1, min(-(-2 * workers // tiles), (kb + 3) // 4))
    cases = {"vit qkv": (27, 400, 8), "vit out": (9, 400, 8), "vit fc": (36, 400, 2), "txt qkv": (12, 616, 6),
             "txt fc": (16, 616, 9), "patch embed": (9, 392, 8)}
    for name, (tiles, kb, want) in cases.items():
        s = lib.dc_gemm_choose_splits(tiles, kb, 74)
        assert s == want, (name, s)
        assert cost(tiles, kb, 74, s) <= cost(tiles, kb, 74, old(tiles, kb, 74)), name
    assert lib.dc_gemm_choose_splits(300, 12, 74) == 1            # enough tiles: no split
    assert lib.dc_gemm_choose_splits(1, 3, 74) == 1               # fewer than 4 k-blocks per split are never made
    for tiles in range(1, 74):
        for kb in (8, 64, 400, 616, 1000):
            s = lib.dc_gemm_choose_splits(tiles, kb, 74)
            assert 1 <= s <= max(1, (kb + 3) // 4) and cost(tiles, kb, 74, s) <= cost(tiles, kb, 74, 1)
