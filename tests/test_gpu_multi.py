"""Sharded selection on real GPUs (skipped with fewer than two): two processes, one per GPU, NCCL
for the handle exchange, then the in-kernel NVLink peer exchange -- picks and posterior must be
bit-identical to the single-GPU run, for both backends and both exchange transports."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _bench(n, *extra, grid=48, active=300):
    cmd = [sys.executable]
    if n > 1:
        cmd += ["-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                "--master-port", "29571"]
    cmd += [os.path.join(ROOT, "bench.py"), "--gpus", str(n), "--grid", str(grid), "--active", str(active), "--steps", "1",
            "--warmup", "1", "--no-cpu-baseline", "--no-tf32", "--no-profile", "--no-sampling", "--no-extras", *extra]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("backend", ["grid", "panel"])
def test_two_gpu_runs_match_single_gpu(backend):
    one = _bench(1, "--backend", backend)
    for exchange in ("peer", "nccl"):
        two = _bench(2, "--backend", backend, "--exchange", exchange)
        assert two["n_gpus"] == 2 and two["n_selected"] == one["n_selected"] == 300
        assert two["ids_crc32"] == one["ids_crc32"], (backend, exchange)
        for a, b in zip(two["posterior_checksum"], one["posterior_checksum"]):
            assert abs(a - b) <= 1e-9 * max(1.0, abs(b))


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")
def test_two_gpu_cube_tiles_with_an_odd_slab_offset():
    """The persistent kernel's 16 x 16 x 16 cubes on slabs that start at an odd z index (45 of 90 slices each: the
    z-axis table rows are then copied from the even index below the cube's first slice) and end in a partial cube."""
    one = _bench(1, "--backend", "grid", grid=90, active=200)
    two = _bench(2, "--backend", "grid", "--exchange", "peer", "--equal-slabs", grid=90, active=200)
    assert two["n_gpus"] == 2 and two["n_selected"] == one["n_selected"] == 200
    assert two["ids_crc32"] == one["ids_crc32"]
    for a, b in zip(two["posterior_checksum"], one["posterior_checksum"]):
        assert abs(a - b) <= 1e-9 * max(1.0, abs(b))
