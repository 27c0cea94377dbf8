"""Generates tests/golden/seq_*.npz: pick sequences of the CPU restatement (oracle/cpu_baseline.py, the
incremental form of select_active_set_straddle.m:62-126 that tests/test_oracle_selection.py ties to the
literal loop) at BASELINE.json's full sizes, plus -- from a dense refit (create_gpis.m / gp_mean.m /
gp_cov.m restated in oracle/cpu_baseline.py::dense_model, predict_dense) on the model points -- the posterior
at a seeded sample of lattice points.  Test infrastructure: the -m gpu tests compare the engine's picks and
its posterior grid with these files.  CPU only.  Config 3: oracle/cpu_baseline.py::select_separable (V never stored; minutes on 8 cores), checked against
the picks select_compact reaches before it runs out of memory (3801 of 4000 on the 62 GB box, 45 minutes).

  python tests/golden/make_sequence_fixtures.py c2|c3|c4s [--steps K]

c4s is BASELINE config 4's CODE PATH at a size the CPU finishes (config 4 itself, 256^3 with 32 000 model rows, is
beyond it): 48^3 sphere WITH gradient observations, K = 1100 -- the selection stops by itself after 719 points
(2876 model rows) because nothing is left unclassified.  Gradient models always take the two-pass append
(k_grid_point / k_grid_solve<4,3> / k_grid_wrow<4>) and the NB = 4 update GEMM, exactly as at 256^3.  Oracle:
oracle/gpis_oracle.py::select_straddle_incremental (se_cov_derivative.m:41-92 blocks, noise leak included).
"""
import argparse
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from gpis_b200.synth import box_tsdf, first_index, sphere_tsdf   # noqa: E402  (input generators only)
from oracle import cpu_baseline as B                             # noqa: E402

HYP = {"cov": np.array([np.log(4.0), 0.0]), "mean": np.zeros(4), "lik": 0.5 * np.log(0.05)}
CASES = {"c2": (sphere_tsdf, 64, 1000), "c3": (box_tsdf, 128, 4000)}

ap = argparse.ArgumentParser()
ap.add_argument("case", choices=sorted(CASES) + ["c4s"])
ap.add_argument("--steps", type=int, default=None)
ap.add_argument("--sample", type=int, default=20000)
ap.add_argument("--compact", action="store_true", help="c3: run select_compact (as far as memory allows) instead of select_separable")
a = ap.parse_args()
if a.case == "c4s":
    from oracle import gpis_oracle as O
    g, K = 48, 1100
    pts, tsdf, nrm = sphere_tsdf(g, with_normals=True)
    N = g ** 3
    t0 = time.perf_counter()
    ora = O.select_straddle_incremental({"points": pts, "tsdf": tsdf, "normals": nrm, "noise": None},
                                        {"activeSetSize": K, "beta": 10.0, "eps": 1e-2, "levelSet": 0.0},
                                        first_index(N), hyp=HYP)
    secs = time.perf_counter() - t0
    print("picks", len(ora["order"]), "in model", ora["n_in_model"], "rows", ora["gp"].r, "seconds", round(secs, 1), flush=True)
    rng = np.random.default_rng(12345)
    sample = np.sort(rng.choice(N, size=min(a.sample, N), replace=False))
    mu, var, nr = O.predict_points(ora["gp"], pts[sample], want_normals=True)
    out = os.path.join(HERE, "seq_c4s.npz")
    np.savez_compressed(out, order=ora["order"].astype(np.int32), n_in_model=int(ora["n_in_model"]), sample=sample.astype(np.int32),
                        mean=mu, var=var, normals=nr, high=np.packbits(ora["high"]), low=np.packbits(ora["low"]),
                        cpu_seconds=secs, cpu_cores=os.cpu_count(), grid=g, K=K, min_top2_gap=ora["min_top2_gap"])
    print("wrote", out, os.path.getsize(out), "bytes")
    sys.exit(0)
gen, g, K = CASES[a.case]
pts, tsdf = gen(g)
N = g ** 3
train = {"activeSetSize": K, "beta": 10.0, "eps": 1e-2, "levelSet": 0.0}
ck_path = os.path.join(ROOT, "build", f"seq_{a.case}_checkpoint.npy")


def checkpoint(order, step):
    np.save(ck_path, np.asarray(order))
    print(f"step {step}: {time.perf_counter() - t0:.0f} s", flush=True)


t0 = time.perf_counter()
prefix_checked = 0
if a.case == "c3":
    # select_compact keeps V = L^-1 K for the unclassified candidates: at 128^3 it runs out of memory near step 3900
    # of 3999 on this 62 GB box (its picks up to there are kept in build/seq_c3_checkpoint.npy by an earlier run of this
    # script with --compact).  The whole sequence comes from select_separable (same rule, V never stored) and is
    # checked against that prefix.
    if a.compact:
        res = B.select_compact(pts, tsdf, HYP, train, first_index(N), max_steps=a.steps, checkpoint=checkpoint)
    else:
        res = B.select_separable((g, g, g), tsdf, HYP, train, first_index(N), max_steps=a.steps,
                                 checkpoint=lambda o, s: print(f"step {s}: {time.perf_counter() - t0:.0f} s", flush=True))
        if os.path.exists(ck_path):
            ck = np.load(ck_path)
            n = min(len(ck), len(res["order"]))
            assert np.array_equal(ck[:n], res["order"][:n]), "select_separable and select_compact disagree"
            prefix_checked = int(n)
            print("picks equal to select_compact's on its first", n, flush=True)
else:
    res = B.select_compact(pts, tsdf, HYP, train, first_index(N), max_steps=a.steps, checkpoint=checkpoint)
order = res["order"]
n_model = int(res["n_in_model"])
print("picks", len(order), "in model", n_model, "seconds", round(res["seconds"], 1), flush=True)
# dense refit on the model points (selection order), posterior at a seeded sample of lattice points
X = pts[order[:n_model]]
L, alpha = B.dense_model(X, tsdf[order[:n_model]], HYP)
rng = np.random.default_rng(12345)
sample = np.sort(rng.choice(N, size=min(a.sample, N), replace=False))
mu = np.empty(sample.size)
var = np.empty(sample.size)
for s0 in range(0, sample.size, 4096):
    sl = slice(s0, s0 + 4096)
    mu[sl], var[sl] = B.predict_dense(X, (L, alpha), pts[sample[sl]], HYP)
out = os.path.join(HERE, f"seq_{a.case}.npz")
np.savez_compressed(out, order=order.astype(np.int32), n_in_model=n_model, scored=res["scored"].astype(np.int32),
                    sample=sample.astype(np.int32), mean=mu, var=var, alpha_head=alpha[:64],
                    cpu_seconds=res["seconds"], cpu_cores=os.cpu_count(), grid=g, K=K,
                    step_seconds=res["step_seconds"].astype(np.float32), prefix_checked_against_select_compact=prefix_checked)
print("wrote", out, os.path.getsize(out), "bytes")
