"""CPU: the C-ABI library loads and exports every symbol include/oov_b200.h declares; the ctypes
table covers them all; the product path refuses CPU tensors (no fallback).  No compute calls."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "oov_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(oov_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib_path():
    import __graft_entry__ as ge
    ge.build()
    import oov_b200
    return oov_b200._lib.LIB_PATH


def test_header_declares_symbols():
    syms = declared_symbols()
    assert "oov_fullsort_topk" in syms and "oov_lsh_embed" in syms and "oov_dhe_hash" in syms
    assert len(syms) >= 25


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for s in declared_symbols():
        assert hasattr(lib, s), f"liboov_b200.so does not export {s}"


def test_ctypes_table_matches_header(lib_path):
    import oov_b200
    assert sorted(oov_b200._lib.SIGNATURES) == declared_symbols()
    lib = oov_b200._lib.load()
    assert b"sm_100a" in lib.oov_version()
    assert lib.oov_launch_count() >= 0


def test_struct_layouts():
    import oov_b200
    assert ctypes.sizeof(oov_b200._lib.OovRows) == 80          # 8 x 8-byte + 2 x (4+4) ... matches struct oov_rows
    assert ctypes.sizeof(oov_b200._lib.OovDheNet) == 80


def test_no_cpu_fallback():
    from oov_b200 import ops
    feat = torch.randn(4, 8)
    planes = torch.randn(16, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.lsh_bits(feat, planes, torch.arange(4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.fullsort_topk(torch.randn(2, 16), torch.randn(8, 16), 3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.dhe_hash(torch.arange(3), torch.zeros(2, 16, dtype=torch.uint8))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    import oov_b200
    monkeypatch.setattr(oov_b200._lib, "_lib", None)
    monkeypatch.setattr(oov_b200._lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        oov_b200._lib.load()


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the product package may reference it."""
    pkg = os.path.join(ROOT, "improving-inductive-oov-recsys_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, os.path.join(dirpath, f)
