"""CPU baseline for bench.py: the oracle's incremental straddle selection, restated for speed
(function-only observations, scalar noise), plus a dense prediction sample.

TEST INFRASTRUCTURE / REPORTED BASELINE ONLY -- see oracle/__init__.py.  Same algorithm as
oracle.gpis_oracle.select_straddle_incremental (which restates
matlab/set_selection/select_active_set_straddle.m:62-126); tests/test_oracle_selection.py checks
that both produce the same index sequence.  Differences are purely mechanical: the columns of
V = L^-1 K(active, candidates) are kept compacted to the unclassified candidates (the reference
only predicts on those, straddle.m:79-81,93-96) and re-compacted when a quarter of them has been
classified, so BLAS works on contiguous memory with all host threads.
"""
import time

import numpy as np
import scipy.linalg as sla


def select_compact(points, tsdf, hyp, train, first_index, max_steps=None, time_budget_s=None, checkpoint=None):
    """checkpoint: optional callable(order_list, step) invoked every 100 steps (long fixture runs)."""
    pts = np.asarray(points, np.float64)
    y = np.asarray(tsdf, np.float64).reshape(-1)
    N, D = pts.shape
    ell2 = float(np.exp(np.float64(hyp["cov"][0])) ** 2)
    sf2 = float(np.exp(2.0 * np.float64(hyp["cov"][1])))
    noise = float(np.exp(2.0 * hyp["lik"]))
    hm = np.asarray(hyp["mean"], np.float64).reshape(-1)
    h, s, eps, K = float(train["levelSet"]), np.sqrt(float(train["beta"])), float(train["eps"]), int(train["activeSetSize"])
    steps = K - 1 if max_steps is None else min(K - 1, int(max_steps))

    cols = np.arange(N)                       # candidate index of each compact column
    live = np.ones(N, bool)                   # per compact column: still unclassified & not active
    P = pts.copy()
    V = np.zeros((steps + 1, N))
    mu = pts @ hm[:D] + hm[D]
    var = np.full(N, sf2)
    Lf = np.zeros((steps + 1, steps + 1))
    g = np.zeros(steps + 1)
    active_pos = {}
    order = [int(first_index)]
    pending = int(first_index)
    pend_col = int(first_index)
    live[pend_col] = False
    t0 = time.perf_counter()
    step_times = []
    scored = []
    r = 0
    for it in range(steps):
        ts = time.perf_counter()
        # append the pending pick: its column of V is L^-1 k(active, x)
        xnew = pts[pending]
        ell = V[:r, pend_col].copy()
        d = np.sqrt(sf2 + noise - ell @ ell)
        Lf[r, :r] = ell
        Lf[r, r] = d
        g[r] = (y[pending] - (xnew @ hm[:D] + hm[D]) - ell @ g[:r]) / d
        diff = P - xnew[None, :]
        k = sf2 * np.exp(-np.einsum("ij,ij->i", diff, diff) / (2.0 * ell2))
        vnew = (k - ell @ V[:r]) / d
        V[r] = vnew
        var -= vnew * vnew
        mu += vnew * g[r]
        r += 1
        if not live.any():
            break
        hi = mu + s * var
        lo = mu - s * var
        amb = np.where(live, np.minimum(hi - h, h - lo), -np.inf)
        scored.append(int(live.sum()))
        mx = amb.max()
        ties = np.flatnonzero(amb == mx)
        b = ties[np.argmin(cols[ties])]                 # lowest candidate index among exact ties
        pend_col, pending = int(b), int(cols[b])
        order.append(pending)
        cls = live & ((lo + eps > h) | (hi - eps < h))
        live[cls] = False
        live[b] = False
        # re-compact when a quarter of the stored columns is dead (keep the pending column)
        if live.sum() + 1 < 0.75 * live.size:
            keep = live.copy()
            keep[b] = True
            idx = np.flatnonzero(keep)
            Vn = np.zeros((steps + 1, idx.size))        # untouched rows stay unmapped (calloc)
            Vn[:r] = V[:r, idx]
            V = Vn
            P, cols, mu, var, live = P[idx], cols[idx], mu[idx], var[idx], live[idx]
            pend_col = int(np.flatnonzero(cols == pending)[0])
        step_times.append(time.perf_counter() - ts)
        if checkpoint is not None and (it + 1) % 100 == 0:
            checkpoint(order, it + 1)
        if time_budget_s is not None and time.perf_counter() - t0 > time_budget_s:
            break
    return {"order": np.asarray(order), "n_in_model": r, "L": Lf[:r, :r], "g": g[:r],
            "seconds": time.perf_counter() - t0, "step_seconds": np.asarray(step_times),
            "scored": np.asarray(scored), "X": pts[np.asarray(order[:r])]}


def predict_dense(X, yres_alpha_L, query, hyp):
    """gp_mean + diag(gp_cov) at `query` for a function-only model given (L lower, alpha)."""
    L, alpha = yres_alpha_L
    ell2 = float(np.exp(np.float64(hyp["cov"][0])) ** 2)
    sf2 = float(np.exp(2.0 * np.float64(hyp["cov"][1])))
    hm = np.asarray(hyp["mean"], np.float64).reshape(-1)
    D = X.shape[1]
    d2 = (np.sum(X * X, axis=1)[:, None] + np.sum(query * query, axis=1)[None, :] - 2.0 * X @ query.T)
    Ks = sf2 * np.exp(-np.maximum(d2, 0.0) / (2.0 * ell2))
    mu = query @ hm[:D] + hm[D] + alpha @ Ks
    W = sla.solve_triangular(L, Ks, lower=True, overwrite_b=True, check_finite=False)
    return mu, sf2 - np.einsum("ij,ij->j", W, W)


def dense_model(X, y, hyp):
    ell2 = float(np.exp(np.float64(hyp["cov"][0])) ** 2)
    sf2 = float(np.exp(2.0 * np.float64(hyp["cov"][1])))
    noise = float(np.exp(2.0 * hyp["lik"]))
    hm = np.asarray(hyp["mean"], np.float64).reshape(-1)
    D = X.shape[1]
    d2 = (np.sum(X * X, axis=1)[:, None] + np.sum(X * X, axis=1)[None, :] - 2.0 * X @ X.T)
    Q = sf2 * np.exp(-np.maximum(d2, 0.0) / (2.0 * ell2)) + noise * np.eye(X.shape[0])
    L = np.linalg.cholesky(Q)
    alpha = sla.cho_solve((L, True), y - (X @ hm[:D] + hm[D]))
    return L, alpha


def select_separable(dims, tsdf, hyp, train, first_index, max_steps=None, checkpoint=None, time_budget_s=None):
    """The same greedy selection for candidates on the integer lattice 0..dims-1 (x fastest), restated so that
    the CPU can hold BASELINE config 3 in memory: V = L^-1 K(active, candidates) is never stored (at 128^3 /
    4000 rows it is 60+ GB and select_compact above runs out of memory near step 3900).  The new row of V is
    formed from W = L^-1 directly,
        v_new[z, y, x] = sum_rho W[r, rho] Tz[rho, z] Ty[rho, y] Tx[rho, x],
    with the separable SE factors T_d[rho, j] = exp(-(j - p_rho,d)^2 / (2 ell^2)) -- one dense matrix product per
    step.  Same rule (select_active_set_straddle.m:62-126: variance-scaled bounds, lowest index on ties, flags
    only ever set), same order of the model rows.  tests/golden/make_sequence_fixtures.py checks its picks against
    select_compact's on the steps the latter completes."""
    nx, ny, nz = (int(d) for d in dims)
    N = nx * ny * nz
    y = np.asarray(tsdf, np.float64).reshape(-1)
    ell2 = float(np.exp(np.float64(hyp["cov"][0])) ** 2)
    sf2 = float(np.exp(2.0 * np.float64(hyp["cov"][1])))
    noise = float(np.exp(2.0 * hyp["lik"]))
    hm = np.asarray(hyp["mean"], np.float64).reshape(-1)
    h, s, eps, K = float(train["levelSet"]), np.sqrt(float(train["beta"])), float(train["eps"]), int(train["activeSetSize"])
    steps = K - 1 if max_steps is None else min(K - 1, int(max_steps))
    ax = [np.arange(n, dtype=np.float64) for n in (nx, ny, nz)]
    idx = np.arange(N)
    coords = np.stack([idx % nx, (idx // nx) % ny, idx // (nx * ny)], axis=1).astype(np.float64)
    mu = coords @ hm[:3] + hm[3]
    var = np.full(N, sf2)
    live = np.ones(N, bool)
    T = [np.zeros((steps + 1, n)) for n in (nx, ny, nz)]
    W = np.zeros((steps + 1, steps + 1))
    X = np.zeros((steps + 1, 3))
    g = np.zeros(steps + 1)
    order = [int(first_index)]
    pending = int(first_index)
    live[pending] = False
    scored, step_times = [], []
    t0 = time.perf_counter()
    r = 0
    for it in range(steps):
        ts = time.perf_counter()
        p = coords[pending]
        diff = X[:r] - p[None, :]
        kvec = sf2 * np.exp(-np.einsum("ij,ij->i", diff, diff) / (2.0 * ell2))
        ell = W[:r, :r] @ kvec
        d = np.sqrt(sf2 + noise - ell @ ell)
        W[r, :r] = -(ell @ W[:r, :r]) / d
        W[r, r] = 1.0 / d
        g[r] = (y[pending] - (p @ hm[:3] + hm[3]) - ell @ g[:r]) / d
        X[r] = p
        for a in range(3):
            T[a][r] = np.exp(-(ax[a] - p[a]) ** 2 / (2.0 * ell2))
        w = W[r, :r + 1] * sf2
        B = ((w[:, None] * T[2][:r + 1])[:, :, None] * T[1][:r + 1][:, None, :]).reshape(r + 1, nz * ny)
        vnew = (B.T @ T[0][:r + 1]).reshape(-1)                      # [(z, y), x] = lattice order, x fastest
        var -= vnew * vnew
        mu += vnew * g[r]
        r += 1
        if not live.any():
            break
        hi = mu + s * var
        lo = mu - s * var
        amb = np.where(live, np.minimum(hi - h, h - lo), -np.inf)
        scored.append(int(live.sum()))
        b = int(np.flatnonzero(amb == amb.max())[0])                  # lowest index among exact ties
        pending = b
        order.append(b)
        live[live & ((lo + eps > h) | (hi - eps < h))] = False
        live[b] = False
        step_times.append(time.perf_counter() - ts)
        if checkpoint is not None and (it + 1) % 100 == 0:
            checkpoint(order, it + 1)
        if time_budget_s is not None and time.perf_counter() - t0 > time_budget_s:
            break
    return {"order": np.asarray(order), "n_in_model": r, "seconds": time.perf_counter() - t0,
            "step_seconds": np.asarray(step_times), "scored": np.asarray(scored), "X": X[:r], "mu": mu, "var": var}
