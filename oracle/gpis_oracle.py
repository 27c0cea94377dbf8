"""NumPy fp64 restatement of the reference's MATLAB GPIS path (the oracle).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

The reference implementation of this path is MATLAB + the GPML toolbox; neither
MATLAB/Octave nor GPML is available (GPML is not vendored in the reference, no
version is pinned anywhere in it).  This module restates, in NumPy fp64:

* ``covSEiso`` (GPML, Rasmussen & Nickisch; published definition
  ``sf^2 exp(-|x-z|^2 / (2 ell^2))`` with ``hyp = [log ell, log sf]``), called at
  ``matlab/core/create_gpis.m:32,78`` and ``matlab/core/se_cov_derivative.m:27``;
* ``meanSum{meanLinear, meanConst}`` (GPML; ``w'x + c``), called at
  ``matlab/core/create_gpis.m:31`` and ``matlab/core/linear_mean_derivative.m:9``;
* ``solve_chol`` (GPML; ``L\\(L'\\b)`` for upper ``L``), called at
  ``matlab/core/create_gpis.m:98`` and ``matlab/core/gp_cov.m:16``;
* the reference's own ``se_cov_derivative.m``, ``linear_mean_derivative.m``,
  ``create_gpis.m``, ``gp_mean.m``, ``gp_cov.m``, ``predict_2d_grid.m``,
  ``select_active_set_straddle.m`` and ``evaluate_errors.m``.

Parity pinning: ``tests/test_oracle_golden.py`` checks this module against the
reference's stored ``gpModel`` structs (``Kxx``, ``L``, ``alpha``) and the one
stored ``predGrid`` (committed as ``tests/golden/*.npz`` by
``tests/golden/make_golden.py``).  The *selection index sequence* is pinned by
no reference data (first index is ``rand()``, ties are ``randperm``, no sequence
is stored): here ``first_index`` is an explicit input and ties go to the lowest
index.  That part of parity is "GPU vs this oracle".
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla


# --------------------------------------------------------------------------
# covariance / mean  (se_cov_derivative.m, linear_mean_derivative.m)
# --------------------------------------------------------------------------
def cov_se_iso(hyp_cov, x, z=None):
    """GPML covSEiso: sf^2 exp(-|x-z|^2/(2 ell^2)); hyp_cov = [log ell, log sf]."""
    x = np.asarray(x, np.float64)
    z = x if z is None else np.asarray(z, np.float64)
    ell = np.exp(np.float64(hyp_cov[0]))
    sf2 = np.exp(2.0 * np.float64(hyp_cov[1]))
    d2 = np.zeros((x.shape[0], z.shape[0]))
    for d in range(x.shape[1]):
        diff = x[:, d][:, None] - z[:, d][None, :]
        d2 += diff * diff
    return sf2 * np.exp(-d2 / (2.0 * ell * ell))


def se_cov_derivative(hyp_cov, beta, x, y=None):
    """Joint covariance of [f; d1 f; ...; dD f], dimension-major blocks.

    Follows matlab/core/se_cov_derivative.m:1-95.  With one point set the noise
    (scalar, or a vector used as diag) is added to the f-f block FIRST (:28-35)
    and the derivative blocks are formed from that noisy block (:47,:85-87) --
    the "noise leak" that turns each gradient diagonal into (sf2+beta_i)/ell^2.
    With two point sets no noise is added (:4-8).
    """
    x = np.asarray(x, np.float64)
    add_noise = y is None
    y = x if y is None else np.asarray(y, np.float64)
    M, D = x.shape
    N = y.shape[0]
    K = np.zeros((M + D * M, N + D * N))
    Kff = cov_se_iso(hyp_cov, x, y)
    if add_noise:
        b = np.asarray(beta, np.float64)
        if b.ndim >= 1 and b.size > 1:          # size(beta,1) > 1  (:11-13)
            Kff = Kff + np.diag(b.reshape(-1))
        else:
            Kff = Kff + np.float64(b.reshape(-1)[0] if b.ndim else b) * np.eye(M, N)
    K[:M, :N] = Kff
    ell2 = np.exp(np.float64(hyp_cov[0])) ** 2
    # diff_d[i, j] = y[j, d] - x[i, d]   (joint_mesh - joint_mesh')(1:M, M+1:M+N)
    diffs = [y[:, d][None, :] - x[:, d][:, None] for d in range(D)]
    for d in range(D):                          # derivative wrt first arg (:41-52)
        K[M + d * M:M + (d + 1) * M, :N] = diffs[d] * Kff / ell2
    for d in range(D):                          # derivative wrt second arg (:58-68)
        K[:M, N + d * N:N + (d + 1) * N] = -K[M + d * M:M + (d + 1) * M, :N]
    for di in range(D):                         # second derivatives (:74-92)
        for dj in range(D):
            K[M + dj * M:M + (dj + 1) * M, N + di * N:N + (di + 1) * N] = (
                (ell2 * (1.0 if di == dj else 0.0) - diffs[di] * diffs[dj]) * Kff / ell2 ** 2)
    return K


def linear_mean(hyp_mean, x):
    """meanSum{meanLinear, meanConst}: w'x + c with hyp_mean = [w; c]."""
    x = np.asarray(x, np.float64)
    hm = np.asarray(hyp_mean, np.float64).reshape(-1)
    D = x.shape[1]
    return x @ hm[:D] + hm[D]


def linear_mean_derivative(hyp_mean, x):
    """matlab/core/linear_mean_derivative.m:1-16 (dimension-major)."""
    x = np.asarray(x, np.float64)
    M, D = x.shape
    hm = np.asarray(hyp_mean, np.float64).reshape(-1)
    out = np.zeros(M + D * M)
    out[:M] = linear_mean(hm, x)
    for d in range(D):
        out[M + d * M:M + (d + 1) * M] = hm[d]
    return out


# --------------------------------------------------------------------------
# fit / predict  (create_gpis.m, gp_mean.m, gp_cov.m, predict_2d_grid.m)
# --------------------------------------------------------------------------
def default_hyp(D, cov=(0.6, 0.0), lik=np.log(0.1)):
    """hyp as select_active_set_straddle.m:58 sets it."""
    return {"cov": np.asarray(cov, np.float64), "mean": np.zeros(D + 1), "lik": float(lik)}


def create_gpis(points, tsdf, normals, noise, hyp):
    """matlab/core/create_gpis.m:1-101 with trainHyp=false (the minimize call is
    commented out in the reference, :46, so trainHyp only resets hyp)."""
    points = np.asarray(points, np.float64)
    tsdf = np.asarray(tsdf, np.float64).reshape(-1)
    m, D = points.shape
    use_gradients = normals is not None and np.size(normals) > 0
    use_noise = noise is not None and np.size(noise) > 0
    y = tsdf
    if use_gradients:
        normals = np.asarray(normals, np.float64).reshape(m, D)
        y = np.concatenate([tsdf, normals.reshape(-1, order="F")])      # :50-53
    beta = np.asarray(noise, np.float64).reshape(-1) if use_noise else np.exp(2.0 * hyp["lik"])
    if use_gradients:
        Kxx = se_cov_derivative(hyp["cov"], beta, points)                # :74
        Mx = linear_mean_derivative(hyp["mean"], points)
        Q = Kxx
    else:
        Kxx = cov_se_iso(hyp["cov"], points)
        Mx = linear_mean(hyp["mean"], points)
        Q = Kxx + (np.diag(beta) if use_noise else beta * np.eye(m))
    if use_gradients or use_noise or beta < 1e-6:                        # :89-95
        L = sla.cholesky(Q, lower=False)
        sl = 1.0
    else:
        L = sla.cholesky(Kxx / beta + np.eye(m), lower=False)
        sl = float(beta)
    alpha = sla.cho_solve((L, False), y - Mx) / sl                       # :98
    return {"hyp": hyp, "training_x": points, "training_y": y, "beta": beta,
            "Kxx": Kxx, "Q": Q, "Mx": Mx, "L": L, "sl": sl, "alpha": alpha,
            "use_gradients": use_gradients}


def gp_mean(model, x, use_derivative):
    """matlab/core/gp_mean.m:1-12.  Returns (mu, Mx, Kxxp)."""
    hyp = model["hyp"]
    if use_derivative:
        Mx = linear_mean_derivative(hyp["mean"], x)
        Kxxp = se_cov_derivative(hyp["cov"], model["beta"], model["training_x"], x)
    else:
        Mx = linear_mean(hyp["mean"], x)
        Kxxp = cov_se_iso(hyp["cov"], model["training_x"], x)
    return Mx + Kxxp.T @ model["alpha"], Mx, Kxxp


def gp_cov(model, x, Kxxp, use_derivative):
    """matlab/core/gp_cov.m:1-21, the FULL posterior covariance (small inputs only)."""
    hyp = model["hyp"]
    if use_derivative:
        Kxx = se_cov_derivative(hyp["cov"], 0.0, x)
        if Kxxp is None:
            Kxxp = se_cov_derivative(hyp["cov"], model["beta"], model["training_x"], x)
    else:
        Kxx = cov_se_iso(hyp["cov"], x)
        if Kxxp is None:
            Kxxp = cov_se_iso(hyp["cov"], model["training_x"], x)
    v = sla.cho_solve((model["L"], False), Kxxp) / model["sl"]
    S = Kxx - Kxxp.T @ v
    return (S + S.T) / 2.0


def gp_var_diag(model, x, Kxxp=None, use_derivative=None):
    """diag(gp_cov(...))(1:numTest): what every caller keeps
    (select_active_set_straddle.m:95-96, predict_2d_grid.m:15-16).
    sf2 - |L^-T k*|^2 over the function columns; latent variance (gp_cov.m:6)."""
    hyp = model["hyp"]
    x = np.asarray(x, np.float64)
    n = x.shape[0]
    if use_derivative is None:
        use_derivative = model["use_gradients"]
    if Kxxp is None:
        Kxxp = (se_cov_derivative(hyp["cov"], model["beta"], model["training_x"], x)
                if use_derivative else cov_se_iso(hyp["cov"], model["training_x"], x))
    kf = Kxxp[:, :n]
    w = sla.solve_triangular(model["L"], kf, trans="T", lower=False)
    sf2 = np.exp(2.0 * np.float64(hyp["cov"][1]))
    return sf2 - np.sum(w * w, axis=0) / model["sl"]


def predict_2d_grid(model, grid_dim, thresh):
    """matlab/core/predict_2d_grid.m:1-27 (points = [X(:) Y(:)] of meshgrid(1:g))."""
    X, Y = np.meshgrid(np.arange(1, grid_dim + 1), np.arange(1, grid_dim + 1))
    pts = np.stack([X.reshape(-1, order="F"), Y.reshape(-1, order="F")], axis=1).astype(np.float64)
    n = pts.shape[0]
    all_mu, _, Kxxp = gp_mean(model, pts, True)
    tsdf = all_mu[:n]
    out = {"gridDim": grid_dim, "tsdf": tsdf,
           "normals": all_mu[n:].reshape(n, 2, order="F"),
           "noise": gp_var_diag(model, pts, Kxxp, True), "points": pts,
           "surfaceThresh": thresh}
    inside = tsdf < 0
    out["com"] = pts[inside].mean(axis=0) if inside.any() else np.full(2, np.nan)
    out["surface"] = np.flatnonzero(np.abs(tsdf) < thresh)
    return out


def evaluate_errors(predicted, true, K):
    """matlab/core/evaluate_errors.m:1-12."""
    diff = np.asarray(predicted, np.float64) - np.asarray(true, np.float64)
    a = np.abs(diff)
    return {"modelSize": K, "rawError": diff, "absError": a, "meanError": a.mean(axis=0),
            "medianError": np.median(a, axis=0), "rmsError": np.sqrt((a * a).mean(axis=0)),
            "stdError": a.std(axis=0, ddof=0)}  # std(x,1): weight flag 1 = divide by N


# --------------------------------------------------------------------------
# posterior shape sampling  (mc_sampling/sample_shapes.m)
# --------------------------------------------------------------------------
def sample_grid_points(grid_dim, resolution=1):
    """sample_shapes.m:10-11: [X(:) Y(:)] of meshgrid(1:res:g, 1:res:g)."""
    ax = np.arange(1, grid_dim + 1, resolution, dtype=np.float64)
    X, Y = np.meshgrid(ax, ax)
    return np.stack([X.reshape(-1, order="F"), Y.reshape(-1, order="F")], axis=1)


def sample_from_posterior(mean, cov, z, jitter=1e-12):
    """mvnrnd(MEAN, COV + 1e-12 I, n) and mvnpdf of the draws (sample_shapes.m:20,31) with the
    standard-normal draws ``z`` [n x d] made explicit (MATLAB's randn stream is not
    reproducible outside MATLAB).  mvnrnd (Statistics Toolbox, not vendored; restated from its
    documented algorithm) forms T = cholcov(Sigma), T'T = Sigma, and returns z T + mu; mvnpdf
    evaluates -(1/2)|X0 / T|^2 - sum(log diag T) - d/2 log(2 pi) with X0 = X - mu, i.e. |z|^2 for
    a draw generated this way.  cholcov's eigen-decomposition fallback for matrices chol()
    rejects is NOT restated: LinAlgError propagates.  Returns (samples [n x d], logpdf [n], T)."""
    mean = np.asarray(mean, np.float64).reshape(-1)
    z = np.atleast_2d(np.asarray(z, np.float64))
    d = mean.shape[0]
    T = sla.cholesky(np.asarray(cov, np.float64) + jitter * np.eye(d), lower=False)
    samples = z @ T + mean[None, :]
    logpdf = -0.5 * np.sum(z * z, axis=1) - np.sum(np.log(np.diag(T))) - 0.5 * d * np.log(2.0 * np.pi)
    return samples, logpdf, T


def sample_shapes(model, grid_dim, z, resolution=1, jitter=1e-12):
    """matlab/mc_sampling/sample_shapes.m:1-33.  ``z`` [numSamples x 3 P] replaces randn.
    Returns (list of {tsdf [P], normals [P x 2]}, logpdf [numSamples]); the reference's ``pdfs``
    is exp(logpdf) (all +inf in every stored constructionResults)."""
    pts = sample_grid_points(grid_dim, resolution)
    P = pts.shape[0]
    MEAN, _, Kxxp = gp_mean(model, pts, True)                            # :15
    COV = gp_cov(model, pts, Kxxp, True)                                 # :14
    samples, logpdf, _ = sample_from_posterior(MEAN, COV, z, jitter)     # :20
    shapes = [{"tsdf": s[:P].copy(), "normals": s[P:].reshape(P, 2, order="F")} for s in samples]  # :25-29
    return shapes, logpdf


# --------------------------------------------------------------------------
# selection  (select_active_set_straddle.m)
# --------------------------------------------------------------------------
def _argmax_lowest(score):
    """Reference picks randomly among exact maxima (straddle.m:109-111);
    the oracle fixes: lowest index."""
    return int(np.flatnonzero(score == score.max())[0])


def select_active_set_straddle(shape, train, first_index, hyp=None, trace=None):
    """LITERAL restatement of matlab/set_selection/select_active_set_straddle.m:62-145:
    refit from scratch every iteration on the active points in index order,
    dimension-major rows.  O(N m^3); small inputs only.

    shape: dict(points N x D, tsdf N, normals N x D or None, noise N or None)
    train: dict(activeSetSize K, beta, eps, levelSet)
    Returns dict(order, active, high, low, model, n_in_model, testIndices, pred*).
    """
    pts = np.asarray(shape["points"], np.float64)
    tsdf = np.asarray(shape["tsdf"], np.float64).reshape(-1)
    N, D = pts.shape
    normals = shape.get("normals")
    use_grad = normals is not None and np.size(normals) > 0
    if use_grad:
        normals = np.asarray(normals, np.float64).reshape(N, D)
    noise = shape.get("noise")
    use_noise = noise is not None and np.size(noise) > 0
    if use_noise:
        noise = np.asarray(noise, np.float64).reshape(-1)
    hyp = default_hyp(D) if hyp is None else hyp
    h, s, eps, K = float(train["levelSet"]), np.sqrt(float(train["beta"])), float(train["eps"]), int(train["activeSetSize"])

    active = np.zeros(N, bool); high = np.zeros(N, bool); low = np.zeros(N, bool)
    active[first_index] = True
    order = [int(first_index)]
    model = None
    train_idx = np.flatnonzero(active)
    min_gap = np.inf
    for _i in range(2, K + 1):
        train_idx = np.flatnonzero(active)                               # :70-76
        unc = ~active & ~high & ~low                                     # :79-81
        test_idx = np.flatnonzero(unc)
        if test_idx.size == 0:                                           # :84-86
            break
        model = create_gpis(pts[train_idx], tsdf[train_idx],
                            normals[train_idx] if use_grad else None,
                            noise[train_idx] if use_noise else None, hyp)
        tp = pts[test_idx]
        mu, _, Kxxp = gp_mean(model, tp, use_grad)                        # :93-94
        mu = mu[:test_idx.size]
        var = gp_var_diag(model, tp, Kxxp, use_grad)                      # :95-96
        hi = mu + s * var                                                # :106-107 (variance!)
        lo = mu - s * var
        amb = np.minimum(hi - h, h - lo)                                 # :108
        b = _argmax_lowest(amb)
        if amb.size > 1:
            top2 = np.partition(amb, -2)[-2:]
            min_gap = min(min_gap, float(top2[1] - top2[0]))
        if trace is not None:
            trace.append({"test_idx": test_idx, "mu": mu, "var": var, "amb": amb})
        active[test_idx[b]] = True                                       # :121
        order.append(int(test_idx[b]))
        high[test_idx[lo + eps > h]] = True                              # :122-125
        low[test_idx[hi - eps < h]] = True
    # final model: training set as of the START of the last executed iteration (:129-130)
    model = create_gpis(pts[train_idx], tsdf[train_idx],
                        normals[train_idx] if use_grad else None,
                        noise[train_idx] if use_noise else None, hyp)
    test_idx = np.flatnonzero(~active)                                   # :132
    out = {"order": np.asarray(order), "active": active, "high": high, "low": low,
           "model": model, "n_in_model": int(train_idx.size), "testIndices": test_idx,
           "min_top2_gap": min_gap}
    if test_idx.size:
        mu, _, Kxxp = gp_mean(model, pts[test_idx], use_grad)
        out["pred_tsdf"] = mu[:test_idx.size]
        if use_grad:
            out["pred_normals"] = mu[test_idx.size:].reshape(test_idx.size, D, order="F")
        out["pred_noise"] = gp_var_diag(model, pts[test_idx], Kxxp, use_grad)
    return out


class IncrementalGP:
    """Point-major incremental GP with the same posterior as create_gpis/gp_mean/
    gp_cov, used for the sizes where the literal loop is O(days).

    Rows are appended per point as [f, d1 f, .., dD f] (point-major).  The
    posterior is invariant to this permutation of se_cov_derivative's
    dimension-major rows; only rounding differs.  Maintains lower L (Q = L L'),
    g = L^-1 (y - m) and, for a fixed candidate set, V = L^-1 K(rows, cand),
    mu and var.
    """

    def __init__(self, cand_pts, hyp, max_points, use_grad, dtype=np.float64):
        self.P = np.asarray(cand_pts, np.float64)
        self.N, self.D = self.P.shape
        self.nb = self.D + 1 if use_grad else 1
        self.use_grad = use_grad
        self.ell2 = float(np.exp(np.float64(hyp["cov"][0])) ** 2)
        self.sf2 = float(np.exp(2.0 * np.float64(hyp["cov"][1])))
        self.hm = np.asarray(hyp["mean"], np.float64).reshape(-1)
        R = max_points * self.nb
        self.L = np.zeros((R, R))
        self.g = np.zeros(R)
        self.V = np.zeros((R, self.N), dtype)
        self.X = np.zeros((max_points, self.D))
        self.r = 0
        self.m = 0
        self.mu = self.P @ self.hm[:self.D] + self.hm[self.D]
        self.var = np.full(self.N, self.sf2)

    def _rows_vs_points(self, x, Z):
        """cov of rows [f(x), d1 f(x), ..] (nb) against f(Z_j): nb x n."""
        diff = Z - x[None, :]                       # z - x  (se_cov_derivative.m:45-47)
        k = self.sf2 * np.exp(-np.sum(diff * diff, axis=1) / (2.0 * self.ell2))
        if not self.use_grad:
            return k[None, :]
        return np.vstack([k[None, :], (diff.T * k[None, :]) / self.ell2])

    def _rows_vs_rows(self, x, Z):
        """cov of the nb rows at x against all nb rows at each Z_j, point-major
        columns: nb x (n*nb).  Block (a, b): a = row kind at x, b = kind at z."""
        n = Z.shape[0]
        diff = Z - x[None, :]
        k = self.sf2 * np.exp(-np.sum(diff * diff, axis=1) / (2.0 * self.ell2))
        if not self.use_grad:
            return k[None, :]
        D, nb = self.D, self.nb
        out = np.zeros((nb, n, nb))
        out[0, :, 0] = k
        for d in range(D):
            out[1 + d, :, 0] = diff[:, d] * k / self.ell2          # [d_d f(x), f(z)]
            out[0, :, 1 + d] = -diff[:, d] * k / self.ell2         # [f(x), d_d f(z)]
        for di in range(D):                                       # col kind
            for dj in range(D):                                   # row kind
                out[1 + dj, :, 1 + di] = (self.ell2 * (di == dj) - diff[:, di] * diff[:, dj]) * k / self.ell2 ** 2
        return out.reshape(nb, n * nb)

    def append(self, x, y_f, y_grad, noise, cand_mask=None):
        """Append one point (nb rows).  noise: variance for this point."""
        nb, r, D = self.nb, self.r, self.D
        x = np.asarray(x, np.float64)
        Qn = self._rows_vs_rows(x, x[None, :]).reshape(nb, nb).copy()
        Qn[0, 0] += noise
        if self.use_grad:
            for d in range(D):                                    # noise-leak quirk
                Qn[1 + d, 1 + d] += noise / self.ell2
        yv = np.concatenate([[y_f], y_grad]) if self.use_grad else np.array([y_f])
        mv = np.concatenate([[x @ self.hm[:D] + self.hm[D]], self.hm[:D]]) if self.use_grad \
            else np.array([x @ self.hm[:D] + self.hm[D]])
        if r:
            Qo = self._rows_vs_rows(x, self.X[:self.m]).T          # r x nb
            Xs = sla.solve_triangular(self.L[:r, :r], Qo, lower=True)
        else:
            Xs = np.zeros((0, nb))
        S = Qn - Xs.T @ Xs
        Ln = np.linalg.cholesky(S)
        self.L[r:r + nb, :r] = Xs.T
        self.L[r:r + nb, r:r + nb] = Ln
        gn = sla.solve_triangular(Ln, yv - mv - Xs.T @ self.g[:r], lower=True)
        self.g[r:r + nb] = gn
        cols = slice(None) if cand_mask is None else cand_mask
        Kc = self._rows_vs_points(x, self.P[cols])                 # nb x n
        T = Kc - Xs.T @ self.V[:r, cols]
        Vn = sla.solve_triangular(Ln, T, lower=True)
        self.V[r:r + nb, cols] = Vn
        self.var[cols] -= np.sum(Vn.astype(np.float64) ** 2, axis=0)
        self.mu[cols] += gn @ Vn
        self.X[self.m] = x
        self.r += nb
        self.m += 1

    def alpha(self):
        return sla.solve_triangular(self.L[:self.r, :self.r], self.g[:self.r], lower=True, trans="T")


def select_straddle_incremental(shape, train, first_index, hyp=None, prune=True, trace=None):
    """Same greedy rule as select_active_set_straddle, restated with rank-1
    (block rank-nb) updates so that it runs at 32^3..128^3.  With prune=True
    only unclassified candidates are updated (the reference only predicts on
    them, straddle.m:79-81,93-96); classified ones are frozen.
    Returns the same keys as the literal version, minus the refit model.

    train may carry the discrete_gp_lse.m variant of the rule (matlab/lse/discrete_gp_lse.m):
      "betaSchedule": ("gotovos", delta)  -> beta_t = 2 log(N t pi^2 / (6 delta)), t = 1, 2, ..  (:87)
      "predictiveVariance": True          -> the variance GPML's gp() returns second (:297-299), which GPML
                                             documents as the predictive variance (latent + noise); GPML is
                                             not in the reference tree, so this reading is unverified there.
    Its classification `ucb_max < tol -> below`, `ucb_min < tol -> above` (:200-201) with
    ucb_max = hi - h, ucb_min = h - lo (:300-301) is the straddle's `hi - eps < h`, `lo + eps > h`
    with eps = tol, so "eps" carries tol."""
    pts = np.asarray(shape["points"], np.float64)
    tsdf = np.asarray(shape["tsdf"], np.float64).reshape(-1)
    N, D = pts.shape
    normals = shape.get("normals")
    use_grad = normals is not None and np.size(normals) > 0
    if use_grad:
        normals = np.asarray(normals, np.float64).reshape(N, D)
    noise = shape.get("noise")
    use_noise = noise is not None and np.size(noise) > 0
    hyp = default_hyp(D) if hyp is None else hyp
    nz = np.asarray(noise, np.float64).reshape(-1) if use_noise else np.full(N, np.exp(2.0 * hyp["lik"]))
    h, s, eps, K = float(train["levelSet"]), np.sqrt(float(train["beta"])), float(train["eps"]), int(train["activeSetSize"])
    sched = train.get("betaSchedule")
    pred_var = bool(train.get("predictiveVariance", False))

    gp = IncrementalGP(pts, hyp, K, use_grad)
    active = np.zeros(N, bool); high = np.zeros(N, bool); low = np.zeros(N, bool)
    active[first_index] = True
    order = [int(first_index)]
    pending = int(first_index)
    min_gap = np.inf
    n_in_model = 0
    amb_full = np.full(N, np.nan)
    for _i in range(2, K + 1):
        unc = ~active & ~high & ~low
        # the model of this iteration contains every pick so far (straddle.m:70-76)
        gp.append(pts[pending], tsdf[pending], normals[pending] if use_grad else None,
                  nz[pending], cand_mask=(unc if prune else None))
        n_in_model = gp.m
        if not unc.any():
            break
        idx = np.flatnonzero(unc)
        mu, var = gp.mu[idx], gp.var[idx]
        if pred_var:
            var = var + nz[idx]
        if sched is not None and sched[0] == "gotovos":                   # discrete_gp_lse.m:87
            s = np.sqrt(2.0 * np.log(N * (_i - 1) * np.pi ** 2 / (6.0 * float(sched[1]))))
        hi = mu + s * var
        lo = mu - s * var
        amb = np.minimum(hi - h, h - lo)
        amb_full[idx] = amb
        b = _argmax_lowest(amb)
        if amb.size > 1:
            top2 = np.partition(amb, -2)[-2:]
            min_gap = min(min_gap, float(top2[1] - top2[0]))
        if trace is not None:
            trace.append({"test_idx": idx, "mu": mu.copy(), "var": var.copy(), "amb": amb})
        pending = int(idx[b])
        active[pending] = True
        order.append(pending)
        high[idx[lo + eps > h]] = True
        low[idx[hi - eps < h]] = True
    return {"order": np.asarray(order), "active": active, "high": high, "low": low,
            "gp": gp, "n_in_model": n_in_model, "testIndices": np.flatnonzero(~active),
            "min_top2_gap": min_gap, "ambiguity": amb_full}


def predict_points(gp_or_model, x, want_normals=False):
    """Posterior mean / latent variance (/ mean gradient) at arbitrary points
    from an IncrementalGP (point-major) -- gp_mean.m:11 and the diagonal of
    gp_cov.m:17 restated for the incremental factor."""
    gp = gp_or_model
    x = np.asarray(x, np.float64)
    n = x.shape[0]
    r, nb, D = gp.r, gp.nb, gp.D
    a = gp.alpha()
    mu = x @ gp.hm[:D] + gp.hm[D]
    var = np.full(n, gp.sf2)
    nrm = np.tile(gp.hm[:D], (n, 1)) if want_normals else None
    Kfull = np.zeros((r, n))
    for i in range(gp.m):
        Kfull[i * nb:(i + 1) * nb] = gp._rows_vs_points(gp.X[i], x)
    mu += a @ Kfull
    W = sla.solve_triangular(gp.L[:r, :r], Kfull, lower=True)
    var -= np.sum(W * W, axis=0)
    if want_normals:
        # d/dx_d of the mean: sum_rows cov(d_d f(x), row) alpha_row
        for i in range(gp.m):
            blk = gp._rows_vs_rows(gp.X[i], x).reshape(nb, n, nb if gp.use_grad else 1) \
                if gp.use_grad else None
            if gp.use_grad:
                # blk[a, j, 1+d] = cov(row a at X_i, d_d f(x_j))
                for d in range(D):
                    nrm[:, d] += a[i * nb:(i + 1) * nb] @ blk[:, :, 1 + d]
            else:
                diff = x - gp.X[i][None, :]
                k = gp.sf2 * np.exp(-np.sum(diff * diff, axis=1) / (2.0 * gp.ell2))
                for d in range(D):
                    nrm[:, d] += a[i] * (-diff[:, d] * k / gp.ell2)
    return mu, var, nrm
