"""CPU-only: the C-ABI library builds for sm_100a, loads, exports every symbol the header
declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    from gpis_b200 import build, _lib
    build.build()
    return _lib.load()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "gpis_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gpis_[a-z_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    from gpis_b200 import _lib
    names = header_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_lib.SYMBOLS) == names


def test_config_struct_matches_header_layout():
    from gpis_b200 import _lib
    # 2 i32, 3 i64, 8 i32, 2+4+5 doubles
    assert C.sizeof(_lib.GpisConfig) == 8 + 24 + 32 + 8 * 11


def test_status_strings(lib):
    assert lib.gpis_status_string(0) == b"ok"
    assert b"no CUDA device" in lib.gpis_status_string(-6)
    assert lib.gpis_version().startswith(b"gpis_b200")


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gpis_b200 import GpisEngine, _lib
    with pytest.raises(_lib.GpisError) as ei:
        GpisEngine(3, 128, 8, ell=2.0)
    assert ei.value.status == _lib.ERR_NO_DEVICE


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gpis_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("no oracle", ""), f
