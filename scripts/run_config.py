"""Run one BASELINE.json-style configuration end to end on the grid backend and print a JSON line
(selection + full posterior grid), e.g. config 4: --shape sphere --grid 256 --active 8000 --gradients
under torchrun on 8 GPUs.  Not the headline bench (bench.py is); used for the other configs."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from gpis_b200 import GpisEngine
from gpis_b200 import dist as gdist
from gpis_b200.synth import box_tsdf, first_index, sphere_tsdf

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="sphere")
ap.add_argument("--grid", type=int, default=64)
ap.add_argument("--active", type=int, default=1000)
ap.add_argument("--ell", type=float, default=4.0)
ap.add_argument("--gradients", action="store_true")
ap.add_argument("--steps", type=int, default=1)
a = ap.parse_args()
world, rank, lr = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
g = a.grid
if a.shape == "sphere":
    out = sphere_tsdf(g, with_normals=a.gradients)
    tsdf, nrm = out[1], (out[2] if a.gradients else None)
else:
    tsdf, nrm = box_tsdf(g)[1], None
N = g ** 3
# function-only models up to 4096 rows run the persistent kernel on 16 x 16 x 16 cubes: cube-aligned slab cuts
z0, z1 = gdist.slab_range_balanced(g, rank, world, a.ell, align=1 if (a.gradients or a.active > 4096) else 16)
lo, hi = g * g * z0, g * g * z1
eng = GpisEngine(3, hi - lo, a.active, ell=a.ell, noise=0.05, level=0.0, eps=1e-2, beta=10.0,
                 use_gradients=a.gradients, rank=rank, world_size=world, total_candidates=N, id_base=lo, device=lr)
sel = gdist.ShardedSelector(eng, rank, world, lo, hi)
if world > 1:
    sel.enable_peer_exchange()
first = first_index(N)
t_nrm = None if nrm is None else np.ascontiguousarray(nrm[lo:hi].T)
times = []
for it in range(a.steps + 1):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    eng.set_grid_candidates((g, g, z1 - z0), (0.0, 0.0, 0.0), (1.0, 1.0, 1.0), tsdf[lo:hi], normals=t_nrm, z0=z0, nz_total=g, soa=True)
    sel.select(first, a.active)
    mu, var = eng.candidate_posterior()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    times.append(time.perf_counter() - t0)
ids, nm = eng.active()
cs = torch.tensor([mu.sum(), var.sum(), float(var.min())], dtype=torch.float64, device=dev)
if world > 1:
    mn = cs[2:].clone()
    dist.all_reduce(cs, op=dist.ReduceOp.SUM)
    dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    cs[2] = mn[0]
if rank == 0:
    t = min(times[1:]) if a.steps else times[0]
    R = nm * (4 if a.gradients else 1)
    print(json.dumps({"config": f"{a.shape} {g}^3, {a.active}-pt active set, gradients={a.gradients}", "n_gpus": world,
                      "seconds": t, "first_run_seconds": times[0], "n_selected": int(len(ids)), "n_in_model": int(nm),
                      "model_rows": int(R), "pts_selected_per_s": len(ids) / t, "grid_pts": N,
                      "fp64_tflops": N * float(R) * R / t / 1e12,
                      "ids_crc32": int(__import__("zlib").crc32(np.asarray(ids, np.int64).tobytes())),
                      "posterior_checksum": [float(cs[0]), float(cs[1])], "min_var": float(cs[2])}))
if world > 1:
    dist.destroy_process_group()
