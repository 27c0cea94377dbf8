"""The reference's own published benchmark: level-set selection of 100 active points
from a 2-D g x g grid, g = 25 .. 800, function values only, with the straddle rule of
select_active_set_straddle.m (beta 10, eps 1e-2, latent variance) on both arms so that the picks can be
compared (ids_crc32).  The published seconds are arrays in matlab/random_code_blocks.m:328-330 (hardware
unstated; MATLAB for the "Serial" column, `GpuActiveSetSelector::SelectChol` on Fermi-era CUDA for "GPU Chol").

GPU arm (default): both backends through the C ABI, device time = CUDA events inside gpis_select, e2e =
host wall clock with host buffers (H2D of the TSDF, selection, D2H of ids + posterior grid).
--cpu: the literal per-step-refit oracle (implementation A) on 25^2, 50^2, 100^2 on the
host cores -- needs no GPU."""
import argparse
import json
import os
import sys
import time
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

PUBLISHED = {25: (0.2078, 0.0887), 50: (0.4642, 0.2583), 100: (1.9013, 1.0540), 200: (15.3608, 4.0220),
             400: (64.0320, 15.5049), 800: (251.9920, 60.1130)}
M, SIGMA, NOISE, TOL = 100, 16.0, 0.05, 0.01          # ell = sqrt(sigma) = 4 grid cells


def circle_tsdf(g, seed=0):
    """Off-centre disc on the g x g lattice (x fastest), truncated SDF + 1e-3 jitter (no exact ties)."""
    idx = np.indices((g, g)).reshape(2, -1)[::-1].T.astype(np.float64)
    c = np.array([0.497, 0.489]) * g
    sd = (np.linalg.norm(idx - c, axis=1) - 0.317 * g) / 4.0
    return idx, np.clip(sd, -1.0, 1.0) + 1e-3 * np.random.default_rng(seed).standard_normal(g * g)


def gpu_arm(grids, reps):
    import torch
    from gpis_b200 import GpisEngine, _lib
    from gpis_b200.synth import first_index
    for g in grids:
        pts, tsdf = circle_tsdf(g)
        n = g * g
        first = first_index(n)
        row = {"grid": g, "candidates": n, "active_set": M, "published_serial_matlab_s": PUBLISHED[g][0],
               "published_gpu_chol_s": PUBLISHED[g][1]}
        ids_ref = None
        for backend in ("grid", "panel"):
            eng = GpisEngine(2, n, M, ell=float(np.sqrt(SIGMA)), sf2=1.0, noise=NOISE, level=0.0, eps=1e-2, beta=10.0)
            dev_ms, e2e_ms = [], []
            for _ in range(reps + 1):                      # first pass warms up
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                if backend == "grid":
                    eng.set_grid_candidates((g, g), (0.0, 0.0), (1.0, 1.0), tsdf)
                else:
                    eng.set_candidates(pts, tsdf)
                eng.select(first, M)
                ids, n_model = eng.active()
                mu, var = eng.candidate_posterior()
                e2e_ms.append((time.perf_counter() - t0) * 1e3)
                dev_ms.append(eng.timing()["select_ms"])
            if ids_ref is None:
                ids_ref = ids
            row[backend + "_select_device_ms"] = min(dev_ms[1:])
            row[backend + "_e2e_ms"] = min(e2e_ms[1:])
            row[backend + "_n_selected"] = int(len(ids))
            row[backend + "_ids_crc32"] = int(zlib.crc32(np.asarray(ids, np.int64).tobytes()))
            row[backend + "_same_picks_as_grid"] = bool(np.array_equal(ids, ids_ref))
            eng.close()
        row["speedup_vs_published_gpu_chol_e2e"] = PUBLISHED[g][1] * 1e3 / row["grid_e2e_ms"]
        row["pts_selected_per_s_e2e"] = row["grid_n_selected"] / (row["grid_e2e_ms"] / 1e3)
        print(json.dumps(row), flush=True)


def cpu_arm(grids):
    from oracle import gpis_oracle as O
    hyp = {"cov": np.array([0.5 * np.log(SIGMA), 0.0]), "mean": np.zeros(3), "lik": 0.5 * np.log(NOISE)}
    tr = {"activeSetSize": M, "beta": 10.0, "eps": 1e-2, "levelSet": 0.0}
    for g in grids:
        pts, tsdf = circle_tsdf(g)
        shape = {"points": pts, "tsdf": tsdf, "normals": None, "noise": None}
        first = 766020790 % (g * g)
        out = {"grid": g, "candidates": g * g, "active_set": M, "published_serial_matlab_s": PUBLISHED[g][0],
               "cores": os.cpu_count()}
        t0 = time.perf_counter()
        r = O.select_active_set_straddle(shape, tr, first, hyp=hyp)
        out["literal_oracle_s"] = time.perf_counter() - t0
        out["literal_n_selected"] = int(len(r["order"]))
        out["ids_crc32"] = int(zlib.crc32(np.asarray(r["order"], np.int64).tobytes()))
        t0 = time.perf_counter()
        r2 = O.select_straddle_incremental(shape, tr, first, hyp=hyp)
        out["incremental_oracle_s"] = time.perf_counter() - t0
        out["same_picks"] = bool(np.array_equal(r["order"], r2["order"]))
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--grids", type=int, nargs="*", default=None)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    if a.cpu:
        cpu_arm(a.grids or [25, 50, 100])
    else:
        gpu_arm(a.grids or [25, 50, 100, 200, 400, 800], a.reps)
