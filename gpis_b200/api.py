"""Host-side mirror of the reference's operator interface for this path.

Same names, argument meaning and outputs as the MATLAB functions
(matlab/set_selection/select_active_set.m, matlab/core/{create_gpis, gp_mean, gp_cov,
predict_2d_grid, create_sparse_gpis, evaluate_errors}.m); structs are dicts.  All GP
arithmetic runs in the CUDA engine through the C ABI; nothing here computes a
covariance on the host.

Differences that are deliberate and documented in DESIGN.md:
  * indices are 0-based;
  * the first point is an argument (`first_index`), ties go to the lowest index
    (the reference draws both at random: select_active_set_straddle.m:36-37,110-111);
  * model rows are point-major [f, d1 f, .., dD f] per point in SELECTION order
    (the reference refits in index order, dimension-major: create_gpis.m:50-53); the
    posterior is identical, `alpha_dimension_major()` gives the reference's ordering;
  * gp_cov returns the diagonal (what every caller keeps: straddle.m:96,
    predict_2d_grid.m:16).
"""
import numpy as np

from . import _lib
from .engine import GpisEngine
from .synth import first_index as _default_first


def default_hyp(D):
    """hyp as select_active_set_straddle.m:58 sets it."""
    return {"cov": np.array([0.6, 0.0]), "mean": np.zeros(D + 1), "lik": float(np.log(0.1))}


def _hyp_to_cfg(hyp, D):
    hm = np.asarray(hyp["mean"], np.float64).reshape(-1)
    return dict(ell=float(np.exp(np.float64(hyp["cov"][0]))), sf2=float(np.exp(2.0 * np.float64(hyp["cov"][1]))),
                mean_w=hm[:D], mean_c=float(hm[D]), noise=float(np.exp(2.0 * float(hyp["lik"]))))


class GpModel(dict):
    """gpModel struct (create_gpis.m:55-100) backed by a live engine handle."""

    @property
    def engine(self):
        return self["_engine"]

    def alpha_dimension_major(self):
        """alpha reordered to create_gpis.m's [f_1..f_m, d1 f_1..d1 f_m, ...]."""
        nb = self.engine.nb
        a = np.asarray(self["alpha"]).reshape(-1, nb)
        return np.concatenate([a[:, k] for k in range(nb)])


def _fill_model(model, eng, hyp, pts, y, dy, beta, ids):
    L, alpha = eng.factor()
    model.update({"_engine": eng, "hyp": hyp, "training_x": pts, "training_y": y, "training_dy": dy,
                  "beta": beta, "L": L, "alpha": alpha, "sl": 1.0, "ids": ids,
                  "use_gradients": eng.use_gradients})
    return model


def create_gpis(points, tsdf, normals, noise, trainHyp=False, hyp=None, **_ignored):
    """create_gpis.m:1-101 (dense fit on the given points).  trainHyp only resets hyp in the
    reference (the minimize call is commented out, :46); pass hyp explicitly."""
    points = np.asarray(points, np.float64)
    m, D = points.shape
    hyp = default_hyp(D) if hyp is None else hyp
    use_grad = normals is not None and np.size(normals) > 0
    use_noise = noise is not None and np.size(noise) > 0
    eng = GpisEngine(D, 0, max(m, 1), use_gradients=use_grad, **_hyp_to_cfg(hyp, D))
    eng.fit(points, tsdf, normals if use_grad else None, noise if use_noise else None)
    beta = np.asarray(noise, np.float64).reshape(-1) if use_noise else np.exp(2.0 * hyp["lik"])
    return _fill_model(GpModel(), eng, hyp, points, np.asarray(tsdf, np.float64).reshape(-1),
                       np.asarray(normals, np.float64).reshape(m, D) if use_grad else None, beta, np.arange(m))


def gp_mean(gpModel, x, use_derivative):
    """gp_mean.m:1-12: mu = Mx + Kxxp' alpha; with use_derivative the normals follow the TSDF,
    dimension-major, as in the reference."""
    mu, _, nrm = gpModel.engine.predict(x, want_normals=bool(use_derivative))
    mu = np.asarray(mu)
    if not use_derivative:
        return mu
    return np.concatenate([mu, np.asarray(nrm).reshape(-1, order="F")])


def gp_cov(gpModel, x, Kxxp=None, use_derivative=True, full=False):
    """gp_cov.m:1-21.  full=False: the diagonal over the function values (latent variance, :6) --
    all that select_active_set_straddle.m:95-96 and predict_2d_grid.m:15-16 keep.  full=True: the
    whole symmetrised posterior covariance of [f; d1 f; ..] (dimension-major) as the reference
    returns it (the consumer is sample_shapes.m:14)."""
    if full:
        return np.asarray(gpModel.engine.posterior_cov(x, use_derivative)[1])
    _, var, _ = gpModel.engine.predict(x)
    return np.asarray(var)


def sample_shapes(gpModel, gridDim, numSamples, resolution=1, z=None, seed=None, jitter=1e-12, return_logpdf=False):
    """mc_sampling/sample_shapes.m:1-33: draws of [tsdf; normals] over meshgrid(1:res:g, 1:res:g)
    from the posterior, and their densities.  The reference draws with MATLAB's randn; here the
    standard-normal draws are `z` [numSamples, 3 P] or come from torch's generator seeded with
    `seed` (on the device).  Returns (shape_samples, pdfs) -- pdfs = exp(logpdf) overflows to inf
    exactly as in the reference's stored constructionResults -- plus logpdf if asked."""
    import torch
    ax = np.arange(1, gridDim + 1, resolution, dtype=np.float64)
    X, Y = np.meshgrid(ax, ax)                                                  # :10-11
    pts = np.stack([X.reshape(-1, order="F"), Y.reshape(-1, order="F")], axis=1)
    P = pts.shape[0]
    if z is None:
        gen = torch.Generator(device="cuda")
        gen.manual_seed(0 if seed is None else int(seed))
        z = torch.randn((int(numSamples), 3 * P), dtype=torch.float64, device="cuda", generator=gen)
    samples, logpdf = gpModel.engine.sample_shapes(pts, z, use_derivative=True, jitter=jitter)
    if torch.is_tensor(samples):
        samples, logpdf = samples.cpu().numpy(), logpdf.cpu().numpy()
    shapes = [{"tsdf": s[:P].copy(), "normals": s[P:].reshape(P, 2, order="F")} for s in samples]   # :25-29
    with np.errstate(over="ignore"):
        pdfs = np.exp(logpdf)
    return (shapes, pdfs, logpdf) if return_logpdf else (shapes, pdfs)


def predict_points(gpModel, x, want_normals=True):
    mu, var, nrm = gpModel.engine.predict(x, want_normals=want_normals)
    return np.asarray(mu), np.asarray(var), (np.asarray(nrm) if want_normals else None)


def predict_2d_grid(gpModel, gridDim, thresh):
    """predict_2d_grid.m:1-27.  Points are [X(:) Y(:)] of meshgrid(1:g, 1:g) (:5-6)."""
    X, Y = np.meshgrid(np.arange(1, gridDim + 1), np.arange(1, gridDim + 1))
    pts = np.stack([X.reshape(-1, order="F"), Y.reshape(-1, order="F")], axis=1).astype(np.float64)
    tsdf, var, nrm = predict_points(gpModel, pts, True)
    pred = {"gridDim": gridDim, "normals": nrm, "tsdf": tsdf, "noise": var, "points": pts,
            "surfaceThresh": thresh}
    inside = tsdf < 0
    pred["com"] = pts[inside].mean(axis=0) if inside.any() else np.full(2, np.nan)
    si = np.flatnonzero(np.abs(tsdf) < thresh)
    surf = {"points": pts[si], "tsdf": tsdf[si], "normals": nrm[si], "noise": var[si]}
    return pred, surf


def predict_grid(gpModel, dims, want_normals=False):
    """Dense posterior over an integer voxel grid 0..dims-1 (x fastest), the 3-D counterpart of
    predict_2d_grid and of Gpis3D.predict_locations (src/grasp_selection/gpis.py:114-123)."""
    from .synth import grid_points
    pts = grid_points(dims)
    mu, var, nrm = predict_points(gpModel, pts, want_normals)
    return {"dims": tuple(dims), "points": pts, "tsdf": mu, "noise": var, "normals": nrm}


def detect_lattice(points):
    """(dims, origin, spacing) if `points` is a full regular lattice listed x fastest
    (IJK_TO_LINEAR, active_set_selection_types.h:6), else None."""
    pts = np.asarray(points, np.float64)
    N, D = pts.shape
    dims, org, sp = [], [], []
    for d in range(D):
        u = np.unique(pts[:, d])
        h = (u[-1] - u[0]) / (u.size - 1) if u.size > 1 else 1.0
        if u.size > 1 and not np.allclose(np.diff(u), h, rtol=0, atol=1e-12 * max(1.0, abs(h))):
            return None
        dims.append(int(u.size)); org.append(float(u[0])); sp.append(float(h))
    if int(np.prod(dims)) != N:
        return None
    idx, stride = np.arange(N), 1
    for d in range(D):
        jd = (idx // stride) % dims[d]
        stride *= dims[d]
        if not np.allclose(org[d] + sp[d] * jd, pts[:, d], rtol=0, atol=1e-12 * max(1.0, abs(sp[d]))):
            return None
    return dims, org, sp


def select_active_set_straddle(shapeParams, trainingParams, first_index=None, hyp=None, predict=True,
                               backend="auto", precision=_lib.PREC_FP64):
    """select_active_set_straddle.m:1-147 on the GPU engine.

    shapeParams: points (N,D), tsdf (N,), normals (N,D) or None/empty, noise (N,) or None/empty
    trainingParams: activeSetSize, beta, eps, levelSet; optionally the discrete_gp_lse.m variant of
      the rule: betaSchedule = ("gotovos", delta) for beta_t = 2 log(N t pi^2 / (6 delta))
      (matlab/lse/discrete_gp_lse.m:87) and predictiveVariance = True for the variance GPML's gp()
      returns (:297-299).  Its tolerance `tol` (:200-201) is `eps` here: ucb_max < tol is
      hi - eps < h, ucb_min < tol is lo + eps > h.
    Returns (gpModel, predShape, testIndices) like the reference; predShape additionally
    carries the level-set state (ambiguity, high, low)."""
    pts = np.asarray(shapeParams["points"], np.float64)
    N, D = pts.shape
    tsdf = np.asarray(shapeParams["tsdf"], np.float64).reshape(-1)
    normals = shapeParams.get("normals")
    use_grad = normals is not None and np.size(normals) > 0
    noise = shapeParams.get("noise")
    use_noise = noise is not None and np.size(noise) > 0
    hyp = default_hyp(D) if hyp is None else hyp
    K = int(trainingParams["activeSetSize"])
    first = _default_first(N) if first_index is None else int(first_index)
    sched = trainingParams.get("betaSchedule")
    if sched is not None and sched[0] != "gotovos":
        raise ValueError("betaSchedule must be ('gotovos', delta)")
    eng = GpisEngine(D, N, K, use_gradients=use_grad, level=float(trainingParams["levelSet"]),
                     eps=float(trainingParams["eps"]), beta=float(trainingParams["beta"]), precision=precision,
                     beta_mode=_lib.BETA_GOTOVOS if sched is not None else _lib.BETA_FIXED,
                     beta_delta=float(sched[1]) if sched is not None else 0.01,
                     variance_mode=_lib.VAR_PREDICTIVE if trainingParams.get("predictiveVariance") else _lib.VAR_LATENT,
                     **_hyp_to_cfg(hyp, D))
    lat = detect_lattice(pts) if backend in ("auto", "grid") else None
    if backend == "grid" and lat is None:
        raise ValueError("backend='grid' needs a full regular lattice listed x fastest")
    if lat is not None:
        eng.set_grid_candidates(lat[0], lat[1], lat[2], tsdf, normals if use_grad else None,
                                noise if use_noise else None)
    else:
        eng.set_candidates(pts, tsdf, normals if use_grad else None, noise if use_noise else None)
    eng.select(first, K)
    ids, n_in_model = eng.active()
    mids = ids[:n_in_model]
    nz = np.asarray(noise, np.float64).reshape(-1)[mids] if use_noise else np.exp(2.0 * hyp["lik"])
    model = _fill_model(GpModel(), eng, hyp, pts[mids], tsdf[mids],
                        np.asarray(normals, np.float64).reshape(N, D)[mids] if use_grad else None, nz, mids)
    model["order"] = ids
    model["n_in_model"] = n_in_model
    model["backend"] = "grid" if lat is not None else "panel"
    active = np.zeros(N, bool)
    active[ids] = True
    test = np.flatnonzero(~active)                                      # straddle.m:132
    amb, high, low, _ = eng.levelset()
    pred = {"points": pts[test], "ambiguity": amb, "high": high, "low": low,
            "gridDim": shapeParams.get("gridDim"), "vertices": shapeParams.get("vertices"),
            "com": shapeParams.get("com")}
    if predict and test.size:
        mu, var, nrm = predict_points(model, pts[test], use_grad)       # straddle.m:136-145
        pred.update({"tsdf": mu, "noise": var, "normals": nrm})
    return model, pred, test


def select_active_set(shapeParams, trainingParams, **kw):
    """select_active_set.m:1-26.  Only the level-set method is on this path; the Random /
    Subsample / Entropy baselines live in the un-vendored GPML toolbox (out of scope)."""
    method = trainingParams.get("activeSetMethod", "LevelSet")
    if method != "LevelSet":
        raise NotImplementedError(f"activeSetMethod {method!r} is outside the accelerated path "
                                  "(select_active_set_{random,subsample,entropy}.m)")
    return select_active_set_straddle(shapeParams, trainingParams, **kw)


def evaluate_errors(predicted, true, K):
    """evaluate_errors.m:1-12 (std(x,1) divides by N)."""
    diff = np.asarray(predicted, np.float64) - np.asarray(true, np.float64)
    a = np.abs(diff)
    return {"modelSize": K, "rawError": diff, "absError": a, "meanError": a.mean(axis=0),
            "medianError": np.median(a, axis=0), "rmsError": np.sqrt((a * a).mean(axis=0)),
            "stdError": a.std(axis=0)}


def create_sparse_gpis(shapeParams, trainingParams, **kw):
    """create_sparse_gpis.m:1-58 without the plotting: selection, held-out errors, grid."""
    import time
    t0 = time.perf_counter()
    model, pred, test = select_active_set(shapeParams, trainingParams, **kw)
    sel_time = time.perf_counter() - t0
    tsdf_err = evaluate_errors(pred["tsdf"], np.asarray(shapeParams["tsdf"]).reshape(-1)[test],
                               trainingParams["activeSetSize"])
    nrm_err = None
    if pred.get("normals") is not None:
        nrm_err = evaluate_errors(pred["normals"], np.asarray(shapeParams["normals"])[test],
                                  trainingParams["activeSetSize"])
    grid = None
    if np.asarray(shapeParams["points"]).shape[1] == 2 and shapeParams.get("gridDim"):
        grid, _ = predict_2d_grid(model, int(shapeParams["gridDim"]), trainingParams.get("surfaceThresh", 0.1))
    return model, grid, tsdf_err, nrm_err, sel_time
