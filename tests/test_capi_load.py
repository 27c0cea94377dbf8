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
    return sorted(set(re.findall(r"\b(gpis_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    from gpis_b200 import _lib
    names = header_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_lib.SYMBOLS) == names


def test_config_struct_matches_header_layout():
    from gpis_b200 import _lib
    # version 0.1 prefix: 2 i32, 3 i64, 8 i32, 2+4+5 doubles; then 4 i32 of mode fields
    assert _lib.CONFIG_SIZE_V1 == 8 + 24 + 32 + 8 * 11
    assert C.sizeof(_lib.GpisConfig) == _lib.CONFIG_SIZE_V1 + 16
    assert _lib.GpisConfig.loop_mode.offset == _lib.CONFIG_SIZE_V1
    hdr = open(os.path.join(ROOT, "include", "gpis_b200.h")).read()
    assert "#define GPIS_CONFIG_SIZE_V1 %d" % _lib.CONFIG_SIZE_V1 in hdr


def test_library_reads_no_environment_variable():
    """Behaviour is selected by gpis_config fields only (no getenv inside the product library)."""
    src = os.path.join(ROOT, "gpis_b200", "csrc")
    for f in os.listdir(src):
        assert "getenv" not in open(os.path.join(src, f)).read(), f


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


def _build_c_consumer(tmp_path):
    import subprocess
    exe = str(tmp_path / "abi_check")
    src = os.path.join(ROOT, "tests", "c_abi", "abi_check.c")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", exe, src,
                    "-ldl", "-lm"], check=True)
    return exe


def test_header_is_valid_c_and_a_c_caller_gets_no_device(lib, tmp_path):
    """include/gpis_b200.h compiles as C99 (-Wall -Werror) and a plain-C caller that dlopens the
    library gets GPIS_ERR_NO_DEVICE on a box without a GPU."""
    import subprocess
    import torch
    from gpis_b200 import _lib
    exe = _build_c_consumer(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu-marked C test")
    r = subprocess.run([exe, _lib.LIB_PATH], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "sizeof(gpis_config)=168" in r.stdout and "create status=-6" in r.stdout


@pytest.mark.gpu
def test_c_caller_runs_a_selection_through_the_abi(lib, tmp_path):
    """The same C program on a GPU: lattice selection + posterior with host buffers, no Python in the loop."""
    import subprocess
    from gpis_b200 import _lib
    exe = _build_c_consumer(tmp_path)
    r = subprocess.run([exe, _lib.LIB_PATH, "gpu"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "selected=30 in_model=29 first=123" in r.stdout
