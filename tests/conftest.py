import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_sessionstart(session):
    """Make sure the C-ABI library exists and is current (nvcc cross-compiles without a GPU)."""
    try:
        from gpis_b200 import build
        build.build()
    except Exception as e:  # the tests that need it will fail loudly
        print("gpis_b200 build failed:", e)
    if os.environ.get("GPIS_EMULATE_TESTS") == "1":
        _swap_in_the_emulated_library()


def _swap_in_the_emulated_library():
    """DEVELOPMENT AID, never used by the driver: GPIS_EMULATE_TESTS=1 runs the `-m gpu` test files on a box
    without a GPU against the library compiled for the CPU emulation (tests/emu, test_emu_library.py) --
    e.g. `GPIS_EMULATE_TESTS=1 python -m pytest tests/test_gpu_parity.py -m gpu -k golden`.  Every CUDA thread
    is an OS thread there: only the small cases are practical, and tests that hand torch CUDA tensors to the
    engine cannot run.  Passing this way says the indexing is right; it is not a GPU result."""
    import subprocess
    import torch
    emu, src = os.path.join(ROOT, "tests", "emu"), os.path.join(ROOT, "gpis_b200", "csrc")
    lib = os.path.join(ROOT, "build", "libgpis_b200_emu.so")
    deps = [os.path.join(src, f) for f in os.listdir(src)] + [os.path.join(emu, f) for f in ("cuda_emu.h", "cuda_rt_emu.h")]
    os.makedirs(os.path.dirname(lib), exist_ok=True)
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.run(["g++", "-std=c++20", "-O2", "-fPIC", "-shared", "-pthread", "-x", "c++", "-DGPIS_EMULATE",
                        "-Wno-unknown-pragmas", "-I" + emu, "-o", lib, os.path.join(src, "gpis_capi.cu")], check=True)
    from gpis_b200 import _lib
    _lib.LIB_PATH, _lib._lib = lib, None
    os.environ.setdefault("GPIS_EMU_SMS", "3")
    torch.cuda.is_available = lambda: True          # the fixtures of the gpu tests assert it
    import gpis_b200.engine as eng
    eng._stream = lambda: None                       # no CUDA stream to pass
    print("GPIS_EMULATE_TESTS=1: gpu tests run against", lib)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def golden_model(g):
    """hyp / inputs of a stored gpModel in oracle form."""
    hyp = {"cov": g["model_hyp_cov"], "mean": g["model_hyp_mean"], "lik": float(g["model_hyp_lik"])}
    m = g["model_training_x"].shape[0]
    D = g["model_training_x"].shape[1]
    y = g["model_training_y"]
    return hyp, g["model_training_x"], y[:m], y[m:].reshape(m, D, order="F"), g["model_beta"]


@pytest.fixture(scope="session")
def golden():
    return load_golden
