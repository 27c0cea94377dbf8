"""DEVELOPMENT AID, never loaded by the default test run or the driver: a pytest plugin that runs the
`-m gpu` test files on a box without a GPU against the library compiled for the CPU emulation
(tests/emu, tests/test_emu_library.py):

    PYTHONPATH=tests/emu python -m pytest -p emulated_gpu_plugin tests/test_gpu_parity.py -m gpu -k golden

Every CUDA thread is an OS thread there: only the small cases are practical, and tests that hand torch
CUDA tensors to the engine cannot run.  Passing this way says the indexing is right; it is not a GPU result."""
import os

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_sessionstart(session):
    _swap_in_the_emulated_library()


def _swap_in_the_emulated_library():
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
    print("emulated_gpu_plugin: gpu tests run against", lib)


