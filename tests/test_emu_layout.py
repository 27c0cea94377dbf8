"""Host-side check of the grid backend's per-candidate state layouts (tile-major, accumulator-fragment order for
128 x 128 and 64 x 64 tiles and for the persistent kernel's 16 x 16 x 16 cubes) against the kernel's own
fragment-slot -> lattice-index map: tests/emu/layout_check.cpp compiled by g++ from the product headers."""
import os
import subprocess

from conftest import ROOT


def test_state_layouts_are_bijections_and_match_the_epilogue_addressing():
    exe = os.path.join(ROOT, "build", "layout_check")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["g++", "-std=c++20", "-O1", "-DGPIS_EMULATE", "-Wno-unknown-pragmas", "-pthread",
                    "-I" + os.path.join(ROOT, "tests", "emu"), os.path.join(ROOT, "tests", "emu", "layout_check.cpp"), "-o", exe],
                   check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "layouts ok" in r.stdout, r.stdout + r.stderr
