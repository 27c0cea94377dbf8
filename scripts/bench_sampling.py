"""Posterior shape sampling (gpis_sample_shapes) timed on the device, next to the oracle on the host.

Two workloads, one JSON line each:
  * "reference": what create_experiment_object.m:23-26 runs and loofa_construction.mat stores a
    samplingTime for -- the stored loofa gpModel, the 25 x 25 grid with derivatives (1875 values),
    1000 samples;
  * "gpis3d": Gpis3D.sample_sdfs on a g^3 grid (function values only) from a model of m points.
Device times are CUDA events inside the C ABI (gpis_get_sampling_timing); the e2e time is the
host-side call with host buffers (H2D of the draws, D2H of the samples).  Not the headline bench."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch

import gpis_b200
from gpis_b200.sdf_io import Gpis3D

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=25)
ap.add_argument("--points", type=int, default=500)
ap.add_argument("--samples", type=int, default=16)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--cpu", action="store_true", help="also time the oracle (tests / bench infrastructure) on the host")
a = ap.parse_args()


def timed(eng, pts, z, deriv, jitter):
    best = None
    for _ in range(a.reps + 1):                      # first call warms up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        samples, logpdf = eng.sample_shapes(pts, z, use_derivative=deriv, jitter=jitter)
        dt = (time.perf_counter() - t0) * 1e3
        t = eng.sampling_timing()
        t["e2e_ms"] = dt
        if best is None or t["cov_ms"] + t["chol_ms"] + t["draw_ms"] < best["cov_ms"] + best["chol_ms"] + best["draw_ms"]:
            best = t
    return best, samples


def flops(R, nv, ns):
    return {"cov": 2.0 * R * R * nv + 2.0 * nv * nv * R, "chol": nv ** 3 / 3.0, "draw": 2.0 * ns * nv * nv}


def line(name, R, nv, ns, t, extra):
    f = flops(R, nv, ns)
    out = {"workload": name, "model_rows": R, "values": nv, "samples": ns, **{k: round(v, 3) for k, v in t.items()},
           "cov_tflops": f["cov"] / t["cov_ms"] * 1e-9, "chol_tflops": f["chol"] / t["chol_ms"] * 1e-9,
           "draw_tflops": f["draw"] / t["draw_ms"] * 1e-9, "dtype": "f64", **extra}
    print(json.dumps(out), flush=True)


# ---- reference workload: stored loofa model, 25 x 25 grid with derivatives, 1000 samples
from conftest import golden_model, load_golden                      # fixtures only (tests/golden)
g = load_golden("loofa")
hyp, X, y, dy, beta = golden_model(g)
mdl = gpis_b200.create_gpis(X, y, dy, beta, False, hyp)
ax = np.arange(1.0, 26.0)
XX, YY = np.meshgrid(ax, ax)
pts = np.stack([XX.reshape(-1, order="F"), YY.reshape(-1, order="F")], axis=1)
z = np.random.default_rng(0).standard_normal((1000, 1875))
jit = 1e-12
try:
    t, _ = timed(mdl.engine, pts, z, True, jit)
except gpis_b200._lib.GpisError:
    jit = 1e-10
    t, _ = timed(mdl.engine, pts, z, True, jit)
extra = {"jitter": jit, "reference_samplingTime_s": 1.282790291,
         "reference_note": "constructionResults.samplingTime stored in data/google_update_versions/loofa_construction.mat (MATLAB, the authors' machine)"}
if a.cpu:
    from oracle import gpis_oracle as O
    ora = O.create_gpis(X, y, dy, beta, hyp)
    t0 = time.perf_counter()
    O.sample_shapes(ora, 25, z, 1, jit)
    extra["cpu_port_s"] = time.perf_counter() - t0
    extra["cpu_cores"] = os.cpu_count()
line("reference: loofa gpModel, 25x25 grid + derivatives", 3 * X.shape[0], 1875, 1000, t, extra)

# ---- Gpis3D.sample_sdfs on a g^3 grid
rng = np.random.default_rng(1)
gdim = a.grid
P = rng.uniform(0, gdim - 1, (a.points, 3))
sd = np.linalg.norm(P - (gdim - 1) / 2.0, axis=1) - 0.3 * gdim
gp3 = Gpis3D(P, sd, (gdim, gdim, gdim), meas_variance=0.01, kernel_hyps=(1.0, 3.0))
n = gdim ** 3
z = torch.randn((a.samples, n), dtype=torch.float64, device="cuda")
t, s = timed(gp3.engine, gp3.grid_pts_, z, False, 1e-8)
extra = {"jitter": 1e-8, "finite": bool(torch.isfinite(s).all())}
if a.cpu:
    from oracle import gpis_oracle as O
    hyp3 = {"cov": np.array([np.log(3.0), 0.0]), "mean": np.zeros(4), "lik": 0.5 * np.log(0.01)}
    ora = O.create_gpis(P, sd, None, None, hyp3)
    t0 = time.perf_counter()
    M, _, K = O.gp_mean(ora, gp3.grid_pts_, False)
    Cv = O.gp_cov(ora, gp3.grid_pts_, K, False)
    O.sample_from_posterior(M, Cv, z.cpu().numpy(), 1e-8)
    extra["cpu_port_s"] = time.perf_counter() - t0
    extra["cpu_cores"] = os.cpu_count()
line("Gpis3D.sample_sdfs %d^3 grid, %d points" % (gdim, a.points), a.points, n, a.samples, t, extra)
