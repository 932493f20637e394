"""CPU: the C-ABI library loads and exports every symbol include/keyhunt_b200.h declares; host-only entry
points behave; the product fails loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import keyhunt_b200 as K

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "keyhunt_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kh_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = K.load_library()
    names = _declared()
    assert len(names) >= 19
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(K.EXPORTS) == names


def test_bloom_params_match_oracle(oracle):
    for e in [1000, 10000, 12345, 262144, 1000000, 8388608, 33554432]:
        h = oracle.bloom_new(e)
        assert K.bloom_params(e).as_dict() == oracle.bloom_desc(h)
        oracle.bloom_free(h)
    with pytest.raises(K.KhError):
        K.bloom_params(999)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(K.KhError) as ei:
        K.KeyHunt(0)
    assert ei.value.code == -1          # KH_ENODEV


def test_struct_layouts_match_header():
    assert C.sizeof(K.BloomDesc) == 32
    assert C.sizeof(K._Hit) == 32 * 3 + 20 + 4 + 8
    assert C.sizeof(K._KeyInfo) == 64 + 60 + 4
    assert C.sizeof(K.BsgsDesc) == 40 + 3 * 32
    assert C.sizeof(K.Stats) == 3 * 8 + 6 * 8


def test_product_does_not_touch_oracle():
    """the product path must not import/link anything under oracle/ or tests/"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "keyhunt_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inc")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f)).read()
                assert "kh_oracle" not in txt and "_oracle" not in txt and "libkh_ref" not in txt, os.path.join(dirpath, f)
                assert "devsim" not in txt or f.endswith(".cuh"), os.path.join(dirpath, f)   # headers only mention it in comments
