"""Sequence parity at BASELINE sizes, in two halves so that no GPU time is spent on the CPU oracle:
  --dump    (on the GPU box)  run config 2 (64^3 sphere, 1000 picks) and config 3 (128^3 box, 4000 picks)
                              on the grid backend and write the pick sequences to gpurun_out/
  --verify  (anywhere)        run the CPU restatement (oracle/cpu_baseline.py) on the same inputs and
                              compare: all 1000 picks of config 2, the first --prefix picks of config 3
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpis_b200.synth import box_tsdf, first_index, sphere_tsdf

HYP = {"cov": np.array([np.log(4.0), 0.0]), "mean": np.zeros(4), "lik": 0.5 * np.log(0.05)}
CASES = {"c2": (sphere_tsdf, 64, 1000), "c3": (box_tsdf, 128, 4000)}

ap = argparse.ArgumentParser()
ap.add_argument("--dump", action="store_true")
ap.add_argument("--verify", action="store_true")
ap.add_argument("--prefix", type=int, default=300)
a = ap.parse_args()
out = os.path.join(ROOT, "gpurun_out")
os.makedirs(out, exist_ok=True)
if a.dump:
    from gpis_b200 import GpisEngine
    for name, (gen, g, K) in CASES.items():
        pts, tsdf = gen(g)
        eng = GpisEngine(3, g ** 3, K, ell=4.0, noise=0.05, level=0.0, eps=1e-2, beta=10.0)
        eng.set_grid_candidates((g, g, g), (0, 0, 0), (1, 1, 1), tsdf)
        eng.select(first_index(g ** 3), K)
        ids, nm = eng.active()
        np.save(os.path.join(out, f"r01_ids_{name}.npy"), ids)
        print(name, len(ids), nm)
if a.verify:
    from oracle import cpu_baseline as B
    res = {}
    for name, (gen, g, K) in CASES.items():
        ids = np.load(os.path.join(out, f"r01_ids_{name}.npy"))
        pts, tsdf = gen(g)
        steps = K - 1 if name == "c2" else a.prefix - 1
        t0 = time.perf_counter()
        cpu = B.select_compact(pts, tsdf, HYP, {"activeSetSize": K, "beta": 10.0, "eps": 1e-2, "levelSet": 0.0},
                               first_index(g ** 3), max_steps=steps)
        n = len(cpu["order"])
        same = bool(np.array_equal(cpu["order"], ids[:n]))
        res[name] = {"grid": g, "gpu_picks": int(len(ids)), "cpu_picks_compared": int(n), "identical": same,
                     "first_difference": None if same else int(np.argmin(cpu["order"] == ids[:n])),
                     "cpu_seconds": round(time.perf_counter() - t0, 1), "cpu_cores": os.cpu_count()}
        print(name, res[name], flush=True)
    json.dump(res, open(os.path.join(ROOT, "profiles", "r01_sequence_parity.json"), "w"), indent=1)
