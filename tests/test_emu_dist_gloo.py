"""N > 1 with the REAL engine on the CPU: two processes over gloo, each driving the emulated library
(tests/emu, see test_emu_library.py) through gpis_b200.dist.ShardedSelector's step protocol -- candidates
sharded (z-slabs for the lattice backend, index ranges for the panel backend), factor replicated, one
fixed-size record per rank all-gathered per step, every rank appending the same row.  On B200s the same
protocol runs over NCCL or the in-kernel peer exchange (tests/test_gpu_multi.py)."""
import ctypes as C
import os
import socket
import subprocess

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from oracle import gpis_oracle as O

EMU_LIB = os.path.join(ROOT, "build", "libgpis_b200_emu.so")
DIMS, K, FIRST = (8, 7, 4), 10, 21
HYP = {"cov": np.array([np.log(2.0), 0.0]), "mean": np.zeros(4), "lik": 0.5 * np.log(0.05)}
TRAIN = {"activeSetSize": K, "beta": 10.0, "eps": 1e-2, "levelSet": 0.0}


def lattice(with_normals=False):
    idx = np.indices(DIMS[::-1]).reshape(3, -1)[::-1]
    pts = np.ascontiguousarray(idx.T.astype(np.float64))
    rng = np.random.default_rng(4)
    c = np.array([3.6, 3.1, 1.7])
    dist_c = np.linalg.norm(pts - c, axis=1)
    sd = (dist_c - 2.4) / 2.0
    tsdf = np.clip(sd, -1, 1) + 1e-3 * rng.standard_normal(len(pts))
    if not with_normals:
        return pts, tsdf
    nrm = (pts - c) / np.maximum(dist_c, 1e-9)[:, None] / 2.0
    nrm[np.abs(sd) >= 1.0] = 0.0
    return pts, tsdf, nrm


def host_exchange(eng, world):
    """The engine's exchange buffers are host memory in the emulated build: view them as CPU tensors."""
    send_p, recv_p, nbytes = eng.exchange_buffers()
    view = lambda p, n: torch.from_numpy(np.ctypeslib.as_array((C.c_double * (n // 8)).from_address(int(p))))
    return view(send_p, nbytes), view(recv_p, nbytes * world)


def _worker(rank, world, port, backend, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), GPIS_EMU_SMS="2")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gpis_b200 import GpisEngine, _lib
        from gpis_b200 import dist as gdist
        _lib.LIB_PATH, _lib._lib = EMU_LIB, None
        grad = backend.endswith("-gradients")
        pts, tsdf, nrm = lattice(True)
        nx, ny, nz = DIMS
        N = len(pts)
        cfg = dict(ell=2.0, noise=0.05, level=0.0, eps=1e-2, beta=10.0, rank=rank, world_size=world, total_candidates=N,
                   use_gradients=grad)
        if backend.startswith("grid"):
            z0, z1 = gdist.slab_range(nz, rank, world)
            lo, hi = nx * ny * z0, nx * ny * z1
            eng = GpisEngine(3, hi - lo, K, id_base=lo, **cfg)
            eng.set_grid_candidates((nx, ny, z1 - z0), (0.0, 0.0, 0.0), (1.0, 1.0, 1.0), tsdf[lo:hi],
                                    normals=nrm[lo:hi] if grad else None, z0=z0, nz_total=nz)
        else:
            lo, hi = gdist.shard_range(N, rank, world)
            eng = GpisEngine(3, hi - lo, K, id_base=lo, **cfg)
            eng.set_candidates(pts[lo:hi], tsdf[lo:hi], normals=nrm[lo:hi] if grad else None)
        sel = gdist.ShardedSelector(eng, rank, world, lo, hi, exchange=host_exchange)
        sel.select(FIRST, K)
        ids, n_model = eng.active()
        mu, var = eng.candidate_posterior()
        L, alpha = eng.factor()
        amb, hi_m, lo_m, ac = eng.levelset()
        q.put((rank, lo, hi, ids.tolist(), int(n_model), mu, var, L, hi_m, lo_m))
    finally:
        dist.destroy_process_group()


@pytest.fixture(scope="module")
def emu_lib():
    emu, src = os.path.join(ROOT, "tests", "emu"), os.path.join(ROOT, "gpis_b200", "csrc")
    deps = [os.path.join(src, f) for f in os.listdir(src)] + [os.path.join(emu, f) for f in ("cuda_emu.h", "cuda_rt_emu.h")]
    os.makedirs(os.path.dirname(EMU_LIB), exist_ok=True)
    if not os.path.exists(EMU_LIB) or any(os.path.getmtime(d) > os.path.getmtime(EMU_LIB) for d in deps):
        subprocess.run(["g++", "-std=c++20", "-O2", "-fPIC", "-shared", "-pthread", "-x", "c++", "-DGPIS_EMULATE",
                        "-Wno-unknown-pragmas", "-I" + emu, "-o", EMU_LIB, os.path.join(src, "gpis_capi.cu")], check=True)
    return EMU_LIB


@pytest.mark.parametrize("backend", ["grid", "panel", "grid-gradients"])
def test_two_ranks_real_engine_over_gloo(emu_lib, backend):
    world = 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, backend, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=900) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    pts, tsdf, nrm = lattice(True)
    ora = O.select_straddle_incremental({"points": pts, "tsdf": tsdf, "normals": nrm if backend.endswith("-gradients") else None,
                                         "noise": None}, TRAIN, FIRST, hyp=HYP, prune=(backend == "panel"))
    assert res[0][3] == res[1][3] == list(ora["order"])              # every rank reports the oracle's sequence
    assert res[0][4] == res[1][4] == ora["n_in_model"]
    assert np.array_equal(res[0][7], res[1][7])                      # replicated factor: bit-identical on both ranks
    r = ora["gp"].r
    assert np.max(np.abs(res[0][7] - ora["gp"].L[:r, :r])) < 1e-11
    high = np.concatenate([x[8] for x in res]); low = np.concatenate([x[9] for x in res])
    assert np.array_equal(high, ora["high"]) and np.array_equal(low, ora["low"])
    mu = np.concatenate([x[5] for x in res]); var = np.concatenate([x[6] for x in res])
    live = ~(high | low) if backend == "panel" else np.ones(len(pts), bool)
    fm, fv, _ = O.predict_points(ora["gp"], pts[live])
    assert np.max(np.abs(mu[live] - fm)) < 1e-10 and np.max(np.abs(var[live] - fv)) < 1e-10
