"""Development check (GPU): the persistent selection kernel (loop_mode AUTO) against the per-step path
(loop_mode STEPS) on growing lattices -- picks, model size, posterior grid, flags, factor -- and timings."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpis_b200 import GpisEngine, _lib
from gpis_b200.synth import box_tsdf, first_index, sphere_tsdf

cases = [(sphere_tsdf, 16, 60, 2.5), (sphere_tsdf, 24, 150, 3.0), (sphere_tsdf, 64, 1000, 4.0), (box_tsdf, 128, 4000, 4.0)]
if len(sys.argv) > 1:
    cases = cases[:int(sys.argv[1])]
out = []
for gen, g, K, ell in cases:
    pts, tsdf = gen(g)
    N = g ** 3
    res = {}
    for name, mode in (("steps", _lib.LOOP_STEPS), ("persist", _lib.LOOP_AUTO)):
        eng = GpisEngine(3, N, K, ell=ell, noise=0.05, level=0.0, eps=1e-2, beta=10.0, loop_mode=mode, spin_timeout_ms=5000)
        eng.set_grid_candidates((g, g, g), (0, 0, 0), (1, 1, 1), tsdf)
        best = None
        for rep in range(2):
            t0 = time.perf_counter()
            eng.select(first_index(N), K)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        ids, nm = eng.active()
        mu, var = eng.candidate_posterior()
        amb, hi, lo, ac = eng.levelset()
        L, alpha = eng.factor()
        t = eng.timing()
        t["persist"] = eng.persist_timing()
        res[name] = dict(ids=ids, nm=nm, mu=mu, var=var, hi=hi, lo=lo, ac=ac, L=L, alpha=alpha, wall=best, t=t,
                         launches=eng.stats()["kernel_launches"])
        eng.close()
    a, b = res["steps"], res["persist"]
    n = min(len(a["ids"]), len(b["ids"]))
    same = bool(len(a["ids"]) == len(b["ids"]) and np.array_equal(a["ids"], b["ids"]))
    row = {"grid": g, "K": K, "picks_equal": same, "n_steps": len(a["ids"]), "n_persist": len(b["ids"]),
           "first_diff": None if same else int(np.argmin(a["ids"][:n] == b["ids"][:n])),
           "nm": [int(a["nm"]), int(b["nm"])],
           "max_dmu": float(np.max(np.abs(a["mu"] - b["mu"]))), "max_dvar": float(np.max(np.abs(a["var"] - b["var"]))),
           "flags_equal": bool(np.array_equal(a["hi"], b["hi"]) and np.array_equal(a["lo"], b["lo"]) and np.array_equal(a["ac"], b["ac"])),
           "max_dL": float(np.max(np.abs(a["L"] - b["L"]))) if a["L"].shape == b["L"].shape else None,
           "max_dalpha": float(np.max(np.abs(a["alpha"] - b["alpha"]))) if a["alpha"].shape == b["alpha"].shape else None,
           "select_ms": [a["t"]["select_ms"], b["t"]["select_ms"]],
           "persist_phase_ms": {k: round(v, 3) for k, v in b["t"]["persist"].items()},
           "launches": [int(a["launches"]), int(b["launches"])]}
    print(json.dumps(row), flush=True)
    out.append(row)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "persist_check.json"), "w"), indent=1)
