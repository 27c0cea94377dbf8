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

from gpis_b200.batch import select_batch, superellipsoid_tsdf

ap = argparse.ArgumentParser()
ap.add_argument("--objects", type=int, default=512)
ap.add_argument("--grid", type=int, default=32)
ap.add_argument("--active", type=int, default=256)
ap.add_argument("--workers", type=int, nargs="*", default=[8])
a = ap.parse_args()
g = a.grid
tsdfs = [superellipsoid_tsdf(g, k) for k in range(a.objects)]
cfg = dict(ell=2.5, noise=0.05, level=0.0, eps=1e-2, beta=10.0)
ref = None
for workers in a.workers:
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
