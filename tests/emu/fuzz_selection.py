"""Development tool (not collected by pytest): random small selection problems through every implementation
of the path -- the oracle, the lattice backend, the panel backend and the object-batched kernel (all three
forms), the last three inside the emulated library (see tests/test_emu_library.py) -- looking for edge cases
(degenerate axes, K >= N, selections that stop at once, predictive variance, non-zero mean / level).
Usage:  python tests/emu/fuzz_selection.py [n_cases] [seed]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import emulated_gpu_plugin                                # noqa: E402  (builds / swaps in the emulated library)

emulated_gpu_plugin._swap_in_the_emulated_library()
import gpis_b200                                          # noqa: E402
from gpis_b200 import _lib                                # noqa: E402
from gpis_b200.batch import select_batch_fused            # noqa: E402
from oracle import gpis_oracle as O                      # noqa: E402


def lattice_points(dims):
    idx = np.indices(tuple(dims)[::-1]).reshape(len(dims), -1)[::-1]
    return np.ascontiguousarray(idx.T.astype(np.float64))


def one_case(rng, case):
    D = int(rng.integers(2, 4))
    dims = tuple(int(rng.integers(1, 9)) for _ in range(D))
    pts = lattice_points(dims)
    N = len(pts)
    K = int(rng.integers(1, min(N, 12) + 1))
    first = int(rng.integers(0, N))
    kind = rng.choice(["blob", "plane", "far", "noise"])
    c = np.array(dims) * rng.uniform(0.2, 0.8, D)
    if kind == "blob":
        t = (np.linalg.norm(pts - c, axis=1) - rng.uniform(0.8, 3.0)) / rng.uniform(1.0, 3.0)
    elif kind == "plane":
        t = (pts @ rng.standard_normal(D) - rng.uniform(0, 4)) / 3.0
    elif kind == "far":
        t = np.full(N, rng.choice([-0.9, 0.9]))
    else:
        t = rng.standard_normal(N)
    tsdf = np.clip(t, -1, 1) + 1e-3 * rng.standard_normal(N)
    ell, sf2, noise = rng.uniform(0.8, 3.0), rng.uniform(0.5, 2.0), 10 ** rng.uniform(-3, -1)
    mean = np.concatenate([0.02 * rng.standard_normal(D), [0.1 * rng.standard_normal()]])
    level, eps, beta = 0.1 * rng.standard_normal(), 10 ** rng.uniform(-3, -1), rng.uniform(0.5, 12.0)
    hyp = {"cov": np.array([np.log(ell), 0.5 * np.log(sf2)]), "mean": mean, "lik": 0.5 * np.log(noise)}
    tr = {"activeSetSize": K, "beta": beta, "eps": eps, "levelSet": level}
    shape = {"points": pts, "tsdf": tsdf, "normals": None, "noise": None}
    tag = f"case {case}: D={D} dims={dims} K={K} first={first} kind={kind}"
    ora = O.select_straddle_incremental(shape, tr, first, hyp=hyp, prune=False)
    fm, fv, _ = O.predict_points(ora["gp"], pts)
    for backend in ("grid", "panel"):
        mdl, pred, test = gpis_b200.select_active_set(shape, tr, first_index=first, hyp=hyp, backend=backend)
        assert np.array_equal(mdl["order"], ora["order"]), (tag, backend, mdl["order"], ora["order"])
        assert mdl["n_in_model"] == ora["n_in_model"], (tag, backend)
        assert np.array_equal(pred["high"], ora["high"]) and np.array_equal(pred["low"], ora["low"]), (tag, backend, "masks")
        mu, var = mdl.engine.candidate_posterior()
        # the panel backend retires a 64-candidate tile once nothing in it competes any more (classified or
        # already selected): its running posterior is exact for the candidates still competing
        live = np.ones(N, bool) if backend == "grid" else ~(ora["high"] | ora["low"] | ora["active"])
        if live.any():
            assert np.max(np.abs(mu[live] - fm[live])) < 1e-9 and np.max(np.abs(var[live] - fv[live])) < 1e-9, (tag, backend, "posterior")
        mdl.engine.close()
    cfg = dict(ell=ell, sf2=sf2, mean_w=mean[:D], mean_c=mean[D], noise=noise, level=level, eps=eps, beta=beta)
    from gpis_b200 import _lib as glib
    for form in (glib.BATCH_PLAIN, glib.BATCH_DFMA, glib.BATCH_DMMA):
        res = select_batch_fused([tsdf, tsdf], dims, K, first, kernel=form, **cfg)
        for r in res:
            assert np.array_equal(r["ids"], ora["order"]) and r["n_in_model"] == ora["n_in_model"], (tag, "batch", form)
            assert np.max(np.abs(r["mean"] - fm)) < 1e-9 and np.max(np.abs(r["var"] - fv)) < 1e-9, (tag, "batch", form)
            assert np.array_equal((r["flags"] & 2) > 0, ora["high"]) and np.array_equal((r["flags"] & 4) > 0, ora["low"]), (tag, form)
    return tag, len(ora["order"])


def one_gradient_case(rng, case):
    """Normal observations (D + 1 rows per point) and per-point noise: lattice and panel backends."""
    D = int(rng.integers(2, 4))
    dims = tuple(int(rng.integers(2, 7)) for _ in range(D))
    pts = lattice_points(dims)
    N = len(pts)
    K = int(rng.integers(2, min(N, 7) + 1))
    first = int(rng.integers(0, N))
    c = np.array(dims) * rng.uniform(0.3, 0.7, D)
    dist = np.linalg.norm(pts - c, axis=1)
    trunc = rng.uniform(1.0, 3.0)
    sd = (dist - rng.uniform(0.8, 2.5)) / trunc
    tsdf = np.clip(sd, -1, 1) + 1e-3 * rng.standard_normal(N)
    nrm = (pts - c) / np.maximum(dist, 1e-9)[:, None] / trunc
    nrm[np.abs(sd) >= 1.0] = 0.0
    per_point = bool(rng.integers(0, 2))
    nz = 10 ** rng.uniform(-3, -1.3, N) if per_point else None
    ell, noise = rng.uniform(1.0, 3.0), 10 ** rng.uniform(-3, -1.5)
    mean = np.concatenate([0.02 * rng.standard_normal(D), [0.05 * rng.standard_normal()]])
    hyp = {"cov": np.array([np.log(ell), 0.0]), "mean": mean, "lik": 0.5 * np.log(noise)}
    tr = {"activeSetSize": K, "beta": rng.uniform(1.0, 12.0), "eps": 10 ** rng.uniform(-3, -1), "levelSet": 0.05 * rng.standard_normal()}
    shape = {"points": pts, "tsdf": tsdf, "normals": nrm, "noise": nz}
    tag = f"gradient case {case}: D={D} dims={dims} K={K} first={first} per_point_noise={per_point}"
    ora = O.select_straddle_incremental(shape, tr, first, hyp=hyp, prune=False)
    fm, fv, fn = O.predict_points(ora["gp"], pts, want_normals=True)
    for backend in ("grid", "panel"):
        mdl, pred, test = gpis_b200.select_active_set(shape, tr, first_index=first, hyp=hyp, backend=backend)
        assert np.array_equal(mdl["order"], ora["order"]), (tag, backend, mdl["order"], ora["order"])
        assert mdl["n_in_model"] == ora["n_in_model"], (tag, backend)
        assert np.array_equal(pred["high"], ora["high"]) and np.array_equal(pred["low"], ora["low"]), (tag, backend, "masks")
        assert np.max(np.abs(pred["tsdf"] - fm[test])) < 1e-8 and np.max(np.abs(pred["noise"] - fv[test])) < 1e-8, (tag, backend)
        assert np.max(np.abs(pred["normals"] - fn[test])) < 1e-8, (tag, backend, "normals")
        r = ora["gp"].r
        assert np.max(np.abs(mdl["L"] - ora["gp"].L[:r, :r])) < 1e-9, (tag, backend, "factor")
        mdl.engine.close()
    return tag, len(ora["order"])


def one_sampling_case(rng, case):
    """Dense fit, prediction at arbitrary points, full posterior covariance and draws against the oracle."""
    D = int(rng.integers(2, 4))
    m = int(rng.integers(1, 40))
    grad = bool(rng.integers(0, 2))
    X = rng.uniform(0, 6, (m, D))
    y = np.linalg.norm(X - 3.0, axis=1) - 2.0 + 0.01 * rng.standard_normal(m)
    dy = rng.standard_normal((m, D)) * 0.3 if grad else None
    per_point = bool(rng.integers(0, 2))
    nzv = 10 ** rng.uniform(-3, -1.3, m) if per_point else None
    ell, sf, noise = rng.uniform(0.8, 3.0), rng.uniform(0.7, 1.4), 10 ** rng.uniform(-3, -1.5)
    mean = np.concatenate([0.02 * rng.standard_normal(D), [0.05 * rng.standard_normal()]])
    hyp = {"cov": np.array([np.log(ell), np.log(sf)]), "mean": mean, "lik": 0.5 * np.log(noise)}
    ora = O.create_gpis(X, y, dy, nzv, hyp)
    mdl = gpis_b200.create_gpis(X, y, dy, nzv, False, hyp)
    nq = int(rng.integers(1, 75))
    q = rng.uniform(-1, 7, (nq, D))
    deriv = bool(rng.integers(0, 2)) and grad            # the reference's gp_mean / gp_cov need matching derivative flags
    tag = f"sampling case {case}: D={D} m={m} grad={grad} per_point_noise={per_point} nq={nq} deriv={deriv}"
    mu_f, _, _ = O.gp_mean(ora, q, grad)
    pm, pv, pn = gpis_b200.predict_points(mdl, q, True)
    assert np.max(np.abs(pm - mu_f[:nq])) < 1e-9, (tag, "mean")
    assert np.max(np.abs(pv - O.gp_var_diag(ora, q))) < 1e-9, (tag, "var")
    if grad:
        assert np.max(np.abs(pn - mu_f[nq:].reshape(nq, D, order="F"))) < 1e-9, (tag, "normals")
    if True:
        MEAN, _, Kx = O.gp_mean(ora, q, grad)
        COV = O.gp_cov(ora, q, Kx, grad)
        if grad and not deriv:                             # function block of the joint posterior
            MEAN, COV = MEAN[:nq], COV[:nq, :nq]
        mean_g, cov_g = mdl.engine.posterior_cov(q, use_derivative=deriv)
        assert np.max(np.abs(mean_g - MEAN)) < 1e-9 and np.max(np.abs(cov_g - COV)) < 1e-10, (tag, "cov")
        ns = int(rng.integers(1, 4))
        z = rng.standard_normal((ns, len(MEAN)))
        try:
            ref, lp, _ = O.sample_from_posterior(MEAN, COV, z, 1e-8)
        except np.linalg.LinAlgError:
            return tag + " (oracle: not PD)", 0
        draws, logpdf = mdl.engine.sample_shapes(q, z, use_derivative=deriv, jitter=1e-8)
        assert np.max(np.abs(draws - ref)) < 1e-6 and np.max(np.abs(logpdf - lp)) < 1e-5 * max(1.0, np.max(np.abs(lp))), (tag, "draws")
    mdl.engine.close()
    return tag, nq


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    t0 = time.time()
    for case in range(n):
        tag, picks = one_case(rng, case)
        print(f"ok  {tag} -> {picks} picks  ({time.time() - t0:.0f} s)", flush=True)
    for case in range(max(1, n // 3)):
        tag, picks = one_gradient_case(rng, case)
        print(f"ok  {tag} -> {picks} picks  ({time.time() - t0:.0f} s)", flush=True)
    for case in range(max(1, n // 3)):
        tag, nq = one_sampling_case(rng, case)
        print(f"ok  {tag}  ({time.time() - t0:.0f} s)", flush=True)
    print("fuzz done:", n, "cases")
