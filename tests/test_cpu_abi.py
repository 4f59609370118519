"""CPU-side checks: the C-ABI library builds/loads and exports every symbol include/mmssl_b200.h
declares; the product path refuses to run without the CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from mmssl_b200 import build
    return build.build()


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "mmssl_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mmssl_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(lib_path):
    lib = ctypes.CDLL(lib_path)
    names = _declared_symbols()
    assert len(names) > 35
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mmssl_b200.h but not exported"


def test_ctypes_table_matches_header(lib_path):
    from mmssl_b200 import _lib
    assert sorted(_lib.exported_symbols()) == _declared_symbols()
    lib = _lib.load(require_device=False)
    assert lib.mmssl_abi_version() == 1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib_path):
    from mmssl_b200 import _lib
    with pytest.raises(_lib.MmsslLibraryError):
        _lib.load(require_device=True)
    from mmssl_b200.functional import SpMMFunction
    from mmssl_b200.graph import BipartiteGraph
    e = torch.zeros(0, dtype=torch.int64)
    with pytest.raises(_lib.MmsslLibraryError):
        BipartiteGraph(e, e, torch.zeros(0), (3, 3))


def test_product_code_never_imports_oracle():
    pkg = os.path.join(ROOT, "mmssl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f
