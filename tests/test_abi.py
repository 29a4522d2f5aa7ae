"""CPU-only: the C-ABI libraries load and export every symbol that include/*.h declares (no compute calls without a GPU),
and the product refuses to run without a CUDA device instead of falling back to anything."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header, macro):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(macro + r"\s+[\w\s\*]+?\b(\w+)\s*\(", src)))


def test_cocg_exports_every_declared_symbol(cocg):
    names = declared("cocg.h", "COCG_API")
    assert len(names) >= 35
    lib = ctypes.CDLL(cocg.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(cocg.SYMBOLS) == names, "the ctypes binding and include/cocg.h disagree"
    assert lib.cocg_version() == 100


def test_cohost_exports_every_declared_symbol(cocg):
    names = declared("cohost.h", "COHOST_API")
    lib = cocg.load_host()
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(cocg.HOST_SYMBOLS) == names, "the ctypes binding and include/cohost.h disagree"


def test_no_cpu_fallback(cocg):
    """Without a GPU, creating a context fails loudly; with one this test is skipped (the GPU suite covers it)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(cocg.CocgError, match="no CUDA device|CUDA"):
        cocg.Context(cocg.BN254, 0)


def test_product_does_not_import_the_oracle():
    """oracle/ is test infrastructure: nothing under the package may import it."""
    pkg = os.path.join(ROOT, "collaborative-circom_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "oracle/" not in txt or f.endswith((".cu", ".cuh", ".cpp", ".hpp")) and "oracle/c" not in txt, f


def test_shard_ranges_partition_the_msm(cocg):
    from importlib import import_module
    dist = import_module("collaborative-circom_b200.distributed")
    for n in (0, 1, 7, 1000, (1 << 20) - 2):
        for world in (1, 2, 3, 8):
            spans = [dist.shard_range(n, r, world) for r in range(world)]
            pos = 0
            for off, ln in spans:
                assert off == min(pos, n) and off + ln <= n
                pos = off + ln
            assert pos == n
