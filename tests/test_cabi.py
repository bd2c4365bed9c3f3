"""The C-ABI library loads without a GPU and exports every symbol include/eqxv_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "eqxv_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(eqxv_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    assert "eqxv_conv2d_igemm_bf16" in syms and "eqxv_attention_fwd_bf16" in syms
    assert len(syms) >= 25


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in include/eqxv_b200.h but not exported: {missing}"


def test_binding_table_matches_header(lib_path):
    from eqxvision_b200 import _lib

    bound = set(_lib.SIGNATURES) | set(_lib._NON_STATUS)
    assert bound == set(declared_symbols())
    lib = _lib.load()
    assert lib.eqxv_version().decode().startswith("eqxv_b200")


def test_no_cpu_fallback_without_device(lib_path):
    """without a GPU eqxv_init must fail loudly (status + message), never fall back"""
    import torch

    from eqxvision_b200 import _lib

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.EqxvError) as ei:
        _lib.call("eqxv_init", 0)
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)


def test_product_code_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under eqxvision_b200/ may reference it"""
    pkg = os.path.join(ROOT, "eqxvision_b200")
    offenders = []
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M):
                    offenders.append(os.path.join(dp, f))
    assert not offenders, offenders
