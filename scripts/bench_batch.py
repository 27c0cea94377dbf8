"""BASELINE.json config 5: batch of 512 independent 32^3 object GPIS fits (M = 256 per object --
SURVEY 8(d)'s assumption, BASELINE.json gives no M).  Prints objects/s for this rank's share."""
import argparse
import json
import os
import sys
import time
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from gpis_b200 import _lib
from gpis_b200.batch import select_batch, select_batch_fused, superellipsoid_tsdf

ap = argparse.ArgumentParser()
ap.add_argument("--objects", type=int, default=512)
ap.add_argument("--grid", type=int, default=32)
ap.add_argument("--active", type=int, default=256)
ap.add_argument("--workers", type=int, nargs="*", default=[8])
ap.add_argument("--fused", action="store_true", help="object-batched kernel (gpis_batch_select): one launch for the batch")
ap.add_argument("--variants", nargs="*", default=["plain", "tuned"], help="inner-loop forms of the fused kernel to time")
ap.add_argument("--pooled", action="store_true", help="engine/stream pool (select_batch); default when --fused is not given")
a = ap.parse_args()
g = a.grid
tsdfs = [superellipsoid_tsdf(g, k) for k in range(a.objects)]
cfg = dict(ell=2.5, noise=0.05, level=0.0, eps=1e-2, beta=10.0)
ref = None
if a.fused:
    stack = np.ascontiguousarray(np.stack(tsdfs))
    for variant in a.variants:
        kern = {"plain": _lib.BATCH_PLAIN, "tuned": _lib.BATCH_DFMA, "dmma": _lib.BATCH_DMMA}[variant]
        cfg["kernel"] = kern                                  # gpis_config.batch_kernel
        select_batch_fused(stack[:8], (g, g, g), a.active, 3254, keep_posterior=False, **cfg)       # warm-up
        best = None
        for _ in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = select_batch_fused(stack, (g, g, g), a.active, 3254, **cfg)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        crc = zlib.crc32(np.concatenate([np.asarray(r["ids"], np.int64) for r in res]).tobytes())
        ref = crc if ref is None else ref
        dev = res[0]["device_ms"] / 1e3
        npts = sum(len(r["ids"]) for r in res)
        print(json.dumps({"config": f"{a.objects} x {g}^3 objects, {a.active}-pt active set each",
                          "mode": "fused (gpis_batch_select), %s inner loops" % variant,
                          "device_seconds": dev, "e2e_seconds": best, "objects_per_s_device": a.objects / dev,
                          "objects_per_s_e2e": a.objects / best, "pts_selected_per_s_e2e": npts / best,
                          "h2d_bytes": int(stack.nbytes), "d2h_bytes": int(stack.nbytes * 2 + stack.size + a.objects * a.active * 8),
                          "fp64_tflops_device": sum(2.0 * g ** 3 * r["n_in_model"] * (r["n_in_model"] + 1) / 2 for r in res) / dev / 1e12,
                          "ids_crc32": crc, "same_picks_as_first_setting": crc == ref}), flush=True)
for workers in (a.workers if (a.pooled or not a.fused) else []):
    select_batch(tsdfs[:workers], (g, g, g), a.active, 3254, n_workers=workers, keep_posterior=False, **cfg)   # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = select_batch(tsdfs, (g, g, g), a.active, 3254, n_workers=workers, **cfg)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    crc = zlib.crc32(np.concatenate([np.asarray(r["ids"], np.int64) for r in res]).tobytes())
    ref = crc if ref is None else ref
    print(json.dumps({"config": f"{a.objects} x {g}^3 objects, {a.active}-pt active set each", "workers": workers,
                      "seconds": dt, "objects_per_s": a.objects / dt,
                      "pts_selected_per_s": sum(len(r["ids"]) for r in res) / dt, "ids_crc32": crc,
                      "same_picks_as_first_setting": crc == ref}), flush=True)
