"""Race detection for the kernels that compile under the CPU emulation (tests/emu): every CUDA thread is
an OS thread there and __syncthreads() / the warp collectives are the only synchronisation, so a barrier
missing between a shared-memory (or same-block global-memory) write and another thread's read is a data
race ThreadSanitizer reports.  Covers gpis_sample.cuh (12 kernels) and both forms of gpis_batch.cuh."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT

EMU = os.path.join(ROOT, "tests", "emu")
EXE = os.path.join(ROOT, "build", "race_check")


@pytest.fixture(scope="module")
def race_check():
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    r = subprocess.run(["g++", "-std=c++20", "-O1", "-g", "-fsanitize=thread", "-pthread", "-Wno-unknown-pragmas",
                        "-I" + EMU, "-o", EXE, os.path.join(EMU, "race_check.cpp")], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("ThreadSanitizer build not available: " + r.stderr[-300:])
    return EXE


def test_detector_reports_a_missing_barrier(race_check):
    r = subprocess.run([race_check, "--self-test"], capture_output=True, text=True, timeout=300)
    assert "self-test done" in r.stdout
    assert "WARNING: ThreadSanitizer: data race" in r.stderr and "k_missing_barrier" in r.stderr
    assert r.returncode != 0


def test_sampling_and_batch_kernels_have_no_data_race(race_check):
    r = subprocess.run([race_check], capture_output=True, text=True, timeout=1500)
    assert "race_check done" in r.stdout, r.stdout[-500:] + r.stderr[-2000:]
    assert "ThreadSanitizer" not in r.stderr, r.stderr[:4000]
    assert r.returncode == 0


def test_whole_library_lattice_backend_has_no_data_race():
    """gpis_capi.cu itself under the emulation + ThreadSanitizer: gpis_create / set_grid_candidates /
    gpis_select on tiny lattices -- the lattice backend's one-pass and two-pass appends, the warp-specialised
    bulk-copy / mbarrier / DMMA update GEMM and the epilogue."""
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    exe = os.path.join(ROOT, "build", "race_check_lib")
    r = subprocess.run(["g++", "-std=c++20", "-O1", "-g", "-fsanitize=thread", "-pthread", "-x", "c++", "-DGPIS_EMULATE",
                        "-Wno-unknown-pragmas", "-I" + EMU, "-o", exe, os.path.join(EMU, "race_check_lib.cpp")],
                       capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("ThreadSanitizer build not available: " + r.stderr[-300:])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=2400, env=dict(os.environ, GPIS_EMU_SMS="2"))
    assert "race_check_lib done" in r.stdout, r.stdout[-500:] + r.stderr[-2000:]
    assert "ThreadSanitizer" not in r.stderr, r.stderr[:4000]
    assert r.returncode == 0
