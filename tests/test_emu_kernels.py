"""Kernels of gpis_b200/csrc run one header at a time on the CPU emulation of the CUDA execution model in
tests/emu (g++ -DGPIS_EMULATE) and compared with the oracle: posterior covariance / blocked Cholesky / draws
(gpis_sample.cuh, with the launch sequences the shipped library calls), the object-batched selection kernel
(gpis_batch.cuh), the dense prediction kernels (gpis_predict.cuh) and the panel backend's selection loop
(gpis_select.cuh).  This checks indexing, barrier structure and launch arithmetic where there is no GPU;
the `-m gpu` tests are the same comparisons through the C ABI on the device, and tests/test_emu_library.py
runs the whole library (gpis_capi.cu) under the same emulation."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import scipy.linalg as sla

from conftest import ROOT, golden_model, load_golden
from oracle import gpis_oracle as O

EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_SO = os.path.join(ROOT, "build", "libemu_kernels.so")
EMU_DEPS = [os.path.join(EMU_DIR, "emu_kernels.cpp"), os.path.join(EMU_DIR, "cuda_emu.h"),
            os.path.join(ROOT, "gpis_b200", "csrc", "gpis_batch_kernel.inc"),
            os.path.join(ROOT, "gpis_b200", "csrc", "gpis_sample.cuh"),
            os.path.join(ROOT, "gpis_b200", "csrc", "gpis_common.cuh"),
            os.path.join(ROOT, "gpis_b200", "csrc", "gpis_batch.cuh"),
            os.path.join(ROOT, "gpis_b200", "csrc", "gpis_launch.cuh"),
            os.path.join(ROOT, "gpis_b200", "csrc", "gpis_predict.cuh"),
            os.path.join(ROOT, "gpis_b200", "csrc", "gpis_select.cuh")]


@pytest.fixture(scope="module")
def emu():
    os.makedirs(os.path.dirname(EMU_SO), exist_ok=True)
    if not os.path.exists(EMU_SO) or any(os.path.getmtime(d) > os.path.getmtime(EMU_SO) for d in EMU_DEPS):
        subprocess.run(["g++", "-std=c++20", "-O2", "-fPIC", "-shared", "-pthread", "-Wno-unknown-pragmas",
                        "-I" + EMU_DIR, "-o", EMU_SO, EMU_DEPS[0]], check=True)
    lib = C.CDLL(EMU_SO)
    vp = C.c_void_p
    lib.emu_sample_shapes.restype = C.c_int
    lib.emu_sample_shapes.argtypes = [C.c_int, C.c_int, C.c_int, vp, C.c_int, vp, C.c_longlong, vp, vp, C.c_longlong,
                                      C.c_int, C.c_double, C.c_double, vp, C.c_double, C.c_int, vp, C.c_double,
                                      vp, vp, vp, vp, vp, C.POINTER(C.c_longlong)]
    lib.emu_chol_upper.restype = C.c_int
    lib.emu_chol_upper.argtypes = [vp, C.c_longlong, C.c_double, vp, C.POINTER(C.c_longlong)]
    return lib


def run_emu(lib, gp, query, deriv, z, jitter):
    """Feed the emulated device code what gpis_capi.cu feeds the real one: active points (SoA),
    Wt[k][row] = (L^-1)[row][k] zero-padded, alpha, the query (SoA)."""
    D, R, m = gp.D, gp.r, gp.m
    ldw = (R + 255) // 256 * 256
    W = sla.solve_triangular(gp.L[:R, :R], np.eye(R), lower=True) if R else np.zeros((0, 0))
    Wt = np.zeros((max(R, 1), ldw))
    Wt[:R, :R] = W.T
    alpha = np.ascontiguousarray(gp.alpha()) if R else np.zeros(1)
    AX = np.ascontiguousarray(gp.X[:max(m, 1)].T)
    q = np.ascontiguousarray(np.asarray(query, np.float64).T)
    nq = q.shape[1]
    nv = nq * (D + 1 if deriv else 1)
    z = np.ascontiguousarray(np.asarray(z, np.float64).reshape(-1, nv))
    ns = z.shape[0]
    mean, cov, chol = np.empty(nv), np.empty((nv, nv)), np.empty((nv, nv))
    samples, logpdf = np.empty((ns, nv)), np.empty(ns)
    mw = np.ascontiguousarray(gp.hm[:D])
    launches = C.c_longlong()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.emu_sample_shapes(D, int(gp.use_grad), R, p(AX), AX.shape[1], p(Wt), ldw, p(alpha), p(q), nq, int(deriv),
                               float(np.sqrt(gp.ell2)), gp.sf2, p(mw), float(gp.hm[D]), ns, p(z), jitter,
                               p(mean), p(cov), p(chol), p(samples), p(logpdf), C.byref(launches))
    return rc, mean, cov, chol, samples, logpdf, launches.value


def chol_launches(nv, blk=64, outer=256):
    """Launches of chol_upper_device: diagonal prep, per 64-block a diagonal kernel (+ block row
    + update inside the 256-wide outer block), per outer block one trailing update, zeroing."""
    npad = (nv + blk - 1) // blk * blk
    n = 1
    for O in range(0, npad, outer):
        oe = min(O + outer, npad)
        for o in range(O, oe, blk):
            n += 1
            if npad - o - blk > 0:
                n += 1 + (1 if oe - o - blk > 0 else 0)
        n += 1 if npad - oe > 0 else 0
    return n + 1


def incremental(hyp, X, y, dy, beta, use_grad):
    gp = O.IncrementalGP(np.zeros((0, X.shape[1])), hyp, max(X.shape[0], 1), use_grad)
    for i in range(X.shape[0]):
        gp.append(X[i], y[i], dy[i] if use_grad else None, beta[i] if np.ndim(beta) else beta)
    return gp


def test_emulated_kernels_match_sample_shapes_on_the_loofa_model(emu):
    """2-D model with gradient rows (a third of the reference's stored loofa gpModel), derivative queries on the
    sample_shapes.m grid at resolution 3: 81 points, 243 values -> four 64-blocks with padding."""
    g = load_golden("loofa")
    hyp, X, y, dy, beta = golden_model(g)
    keep = np.arange(0, X.shape[0], 3)                      # 42 of the 126 points: the emulation is slow
    X, y, dy, beta = X[keep], y[keep], dy[keep], beta[keep]
    model = O.create_gpis(X, y, dy, beta, hyp)
    gp = incremental(hyp, X, y, dy, beta, True)
    pts = O.sample_grid_points(25, 3)
    z = np.random.default_rng(3).standard_normal((3, 3 * len(pts)))
    MEAN, _, Kxxp = O.gp_mean(model, pts, True)
    COV = O.gp_cov(model, pts, Kxxp, True)
    shapes, lp = O.sample_shapes(model, 25, z, resolution=3, jitter=1e-12)
    rc, mean, cov, chol, samples, logpdf, launches = run_emu(emu, gp, pts, True, z, 1e-12)
    assert rc == 0
    assert np.max(np.abs(mean - MEAN)) < 1e-11
    assert np.max(np.abs(cov - COV)) < 1e-12 and np.array_equal(cov, cov.T)
    assert np.all(np.tril(chol, -1) == 0.0)
    assert np.max(np.abs(chol.T @ chol - (COV + 1e-12 * np.eye(len(MEAN))))) < 1e-12
    ref = np.stack([np.concatenate([s["tsdf"], s["normals"].reshape(-1, order="F")]) for s in shapes])
    assert np.max(np.abs(samples - ref)) < 1e-7
    assert np.max(np.abs(logpdf - lp)) < 1e-6 * np.max(np.abs(lp))
    assert launches == 3 + 2 + 1 + chol_launches(len(MEAN)) + 4           # the count gpis_capi.cu reports


def test_emulated_kernels_3d_function_only_and_empty_model(emu):
    """3-D function-only model (Gpis3D.sample_sdfs shape): no derivative values, 5^3 = 125 points."""
    rng = np.random.default_rng(5)
    X = rng.uniform(0, 4, (40, 3))
    y = np.linalg.norm(X - 2.0, axis=1) - 1.5
    hyp = {"cov": np.array([np.log(1.3), np.log(0.9)]), "mean": np.array([0.01, -0.02, 0.03, 0.2]), "lik": np.log(0.1)}
    model = O.create_gpis(X, y, None, None, hyp)
    gp = incremental(hyp, X, y, None, np.exp(2 * hyp["lik"]), False)
    ax = np.arange(5.0)
    pts = np.stack(np.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(-1, 3)
    z = rng.standard_normal((2, 125))
    MEAN, _, Kxxp = O.gp_mean(model, pts, False)
    COV = O.gp_cov(model, pts, Kxxp, False)
    ref, lp, T = O.sample_from_posterior(MEAN, COV, z, 1e-10)
    rc, mean, cov, chol, samples, logpdf, _ = run_emu(emu, gp, pts, False, z, 1e-10)
    assert rc == 0
    assert np.max(np.abs(mean - MEAN)) < 1e-12 and np.max(np.abs(cov - COV)) < 1e-12
    assert np.max(np.abs(samples - ref)) < 1e-8 and np.max(np.abs(logpdf - lp)) < 1e-6 * np.max(np.abs(lp))
    # derivative queries against a function-only model (rows have no gradient kind)
    pts2 = pts[::7]
    rc, mean, cov, chol, _, _, _ = run_emu(emu, gp, pts2, True, np.zeros((0, 4 * len(pts2))), 1e-10)
    assert rc == 0
    Kss = O.se_cov_derivative(hyp["cov"], 0.0, pts2)
    Ks = O.se_cov_derivative(hyp["cov"], 0.0, X, pts2)[:len(X)]           # function rows of the training block
    Q = model["Q"]
    S = Kss - Ks.T @ np.linalg.solve(Q, Ks)
    assert np.max(np.abs(cov - (S + S.T) / 2)) < 1e-11
    assert np.max(np.abs(chol.T @ chol - cov - 1e-10 * np.eye(len(cov)))) < 1e-12
    # a model with no rows: prior mean and prior covariance
    empty = O.IncrementalGP(np.zeros((0, 3)), hyp, 1, False)
    rc, mean, cov, chol, _, _, _ = run_emu(emu, empty, pts2, False, np.zeros((0, len(pts2))), 1e-10)
    assert rc == 0
    assert np.allclose(mean, pts2 @ hyp["mean"][:3] + hyp["mean"][3], rtol=0, atol=1e-15)
    assert np.max(np.abs(cov - O.cov_se_iso(hyp["cov"], pts2))) < 1e-15


def test_emulated_cholesky_reports_a_bad_pivot(emu):
    """Duplicate query points make COV exactly singular; without jitter the second copy's pivot is <= 0
    (or rounds to a tiny positive number) and the status must carry its index instead of NaNs."""
    hyp = {"cov": np.array([0.0, 0.0]), "mean": np.zeros(3), "lik": np.log(0.1)}
    gp = O.IncrementalGP(np.zeros((0, 2)), hyp, 1, False)
    pts = np.array([[0.0, 0.0], [1.0, 0.5], [0.0, 0.0], [2.0, 2.0]])
    rc, *_ = run_emu(emu, gp, pts, False, np.zeros((0, 4)), 0.0)
    assert rc == 3                                                            # 1-based index of the pivot
    rc, _, cov, chol, *_ = run_emu(emu, gp, pts, False, np.zeros((0, 4)), 1e-9)
    assert rc == 0 and np.max(np.abs(chol.T @ chol - cov - 1e-9 * np.eye(4))) < 1e-15


def test_emulated_two_level_cholesky_across_outer_blocks(emu):
    """330 values -> 384 padded: the first 256-wide outer block updates the trailing 128 x 128 with
    K = 256, the second one is factored by 64-blocks; compared with LAPACK and by reconstruction."""
    rng = np.random.default_rng(8)
    n = 330
    B = rng.standard_normal((n, 40))
    A = B @ B.T / 40.0 + np.diag(rng.uniform(0.01, 0.1, n))
    A = (A + A.T) / 2
    T = np.empty((n, n))
    launches = C.c_longlong()
    rc = emu.emu_chol_upper(A.ctypes.data_as(C.c_void_p), n, 1e-9, T.ctypes.data_as(C.c_void_p), C.byref(launches))
    assert rc == 0 and launches.value == chol_launches(n)
    ref = sla.cholesky(A + 1e-9 * np.eye(n), lower=False)
    assert np.all(np.tril(T, -1) == 0.0)
    assert np.max(np.abs(T - ref)) < 1e-12
    assert np.max(np.abs(T.T @ T - A - 1e-9 * np.eye(n))) < 1e-14


# ------------------------------------------------------------------------------------------------
# gpis_batch.cuh: one thread block runs the whole selection of one object
# ------------------------------------------------------------------------------------------------
def run_emu_batch(lib, dims, tsdfs, first, K, hyp, train, n_blocks, variance_mode=0, fast=0):
    vp = C.c_void_p
    lib.emu_batch_select.restype = C.c_int
    lib.emu_batch_select.argtypes = [vp, vp, vp, C.c_int, vp, C.c_longlong, C.c_int] + [C.c_double] * 6 + \
        [C.c_int, vp, C.c_double, C.c_int, C.c_int] + [vp] * 7
    B, N = len(tsdfs), int(np.prod(dims))
    d3 = np.ascontiguousarray(np.array(list(dims) + [1] * (3 - len(dims)), np.int32))
    org, sp = np.zeros(3), np.ones(3)
    t = np.ascontiguousarray(np.stack(tsdfs).astype(np.float64))
    ids = np.full((B, K), -1, np.int64)
    n_sel, n_model, err = (np.zeros(B, np.int32) for _ in range(3))
    mu, var, flags = np.empty((B, N)), np.empty((B, N)), np.empty((B, N), np.uint8)
    mw = np.zeros(3)
    mw[:len(dims)] = hyp["mean"][:len(dims)]
    p = lambda a: a.ctypes.data_as(vp)
    rc = lib.emu_batch_select(p(d3), p(org), p(sp), B, p(t), first, K, float(np.exp(hyp["cov"][0])),
                              float(np.exp(2 * hyp["cov"][1])), float(np.exp(2 * hyp["lik"])), train["levelSet"],
                              train["eps"], train["beta"], variance_mode, p(mw), float(hyp["mean"][len(dims)]), n_blocks, int(fast),
                              p(ids), p(n_sel), p(n_model), p(err), p(mu), p(var), p(flags))
    assert rc == 0
    return ids, n_sel, n_model, err, mu, var, flags


def lattice_points(dims):
    idx = np.indices(tuple(dims)[::-1]).reshape(len(dims), -1)[::-1]
    return np.ascontiguousarray(idx.T.astype(np.float64))


@pytest.mark.parametrize("fast", [0, 1, 2])
@pytest.mark.parametrize("dims,K,n_blocks", [((10, 9, 7), 24, 2), ((13, 12), 30, 3), ((6, 5, 19), 14, 1)])
def test_emulated_batch_kernel_matches_the_oracle_per_object(emu, dims, K, n_blocks, fast):
    """Three objects over two or three blocks (a block takes its objects one after another, reusing
    its W workspace): picks, model size, posterior of EVERY lattice point and the level-set flags
    against the incremental oracle; K - 1 points in the model (straddle.m:129-130)."""
    rng = np.random.default_rng(7)
    pts = lattice_points(dims)
    D = len(dims)
    hyp = {"cov": np.array([np.log(2.2), np.log(1.1)]), "mean": np.concatenate([0.01 * np.arange(1, D + 1), [0.05]]),
           "lik": 0.5 * np.log(0.03)}
    train = {"activeSetSize": K, "beta": 10.0, "eps": 1e-2, "levelSet": 0.0}
    tsdfs = []
    for k in range(3):
        c = np.array(dims) * rng.uniform(0.4, 0.6, D)
        tsdfs.append(np.clip((np.linalg.norm(pts - c, axis=1) - 0.3 * min(dims) - 0.3 * k) / 2.0, -1, 1)
                     + 1e-3 * rng.standard_normal(len(pts)))
    first = 17
    ids, n_sel, n_model, err, mu, var, flags = run_emu_batch(emu, dims, tsdfs, first, K, hyp, train, n_blocks, fast=fast)
    assert not err.any()
    for k in range(3):
        shape = {"points": pts, "tsdf": tsdfs[k], "normals": None, "noise": None}
        ora = O.select_straddle_incremental(shape, train, first, hyp=hyp, prune=False)
        assert n_sel[k] == len(ora["order"]) and np.array_equal(ids[k, :n_sel[k]], ora["order"]), k
        assert n_model[k] == ora["n_in_model"]
        fm, fv, _ = O.predict_points(ora["gp"], pts)
        assert np.max(np.abs(mu[k] - fm)) < 1e-10 and np.max(np.abs(var[k] - fv)) < 1e-10
        assert np.array_equal((flags[k] & 1) > 0, ora["active"])
        assert np.array_equal((flags[k] & 2) > 0, ora["high"]) and np.array_equal((flags[k] & 4) > 0, ora["low"])
    assert not (np.array_equal(mu[0], mu[1]) and np.array_equal(mu[1], mu[2]))      # the objects are different objects


def test_emulated_batch_kernel_early_stop_and_k_one(emu):
    """A flat TSDF far from the level set classifies everything after the first point: the selection stops
    with the points picked so far; K = 1 selects the first point and fits nothing."""
    dims = (6, 5, 4)
    pts = lattice_points(dims)
    hyp = {"cov": np.array([np.log(3.0), 0.0]), "mean": np.zeros(4), "lik": 0.5 * np.log(0.01)}
    train = {"activeSetSize": 12, "beta": 1.0, "eps": 1e-2, "levelSet": 0.0}
    tsdf = np.full(len(pts), 0.9) + 1e-4 * np.arange(len(pts))
    shape = {"points": pts, "tsdf": tsdf, "normals": None, "noise": None}
    ora = O.select_straddle_incremental(shape, train, 3, hyp=hyp, prune=False)
    ids, n_sel, n_model, err, mu, var, flags = run_emu_batch(emu, dims, [tsdf], 3, 12, hyp, train, 1)
    assert len(ora["order"]) < 12                                  # the oracle stops early too
    assert n_sel[0] == len(ora["order"]) and np.array_equal(ids[0, :n_sel[0]], ora["order"])
    assert n_model[0] == ora["n_in_model"] and err[0] == 0
    ids, n_sel, n_model, err, mu, var, flags = run_emu_batch(emu, dims, [tsdf], 3, 1, hyp, train, 1)
    assert n_sel[0] == 1 and ids[0, 0] == 3 and n_model[0] == 0
    assert np.all(var[0] == 1.0) and flags[0, 3] == 1 and flags[0].sum() == 1


# ------------------------------------------------------------------------------------------------
# gpis_predict.cuh: factor inverse by columns, alpha, fused dense prediction (gp_mean / diag gp_cov)
# ------------------------------------------------------------------------------------------------
def run_emu_predict(lib, gp, query, want_nrm):
    vp = C.c_void_p
    lib.emu_predict.restype = C.c_int
    lib.emu_predict.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, vp, vp, C.c_longlong, C.c_double,
                                C.c_double, vp, C.c_double, C.c_int, vp, vp, vp, vp]
    D, R, m = gp.D, gp.r, gp.m
    L = np.ascontiguousarray(gp.L[:max(R, 1), :max(R, 1)])
    g = np.ascontiguousarray(gp.g[:max(R, 1)])
    AX = np.ascontiguousarray(gp.X[:max(m, 1)].T)
    q = np.ascontiguousarray(np.asarray(query, np.float64).T)
    nq = q.shape[1]
    mean, var, nrm, alpha = np.empty(nq), np.empty(nq), np.empty((D, nq)), np.empty(max(R, 1))
    mw = np.ascontiguousarray(gp.hm[:D])
    p = lambda a: a.ctypes.data_as(vp)
    rc = lib.emu_predict(D, int(gp.use_grad), R, AX.shape[1], p(AX), p(L), L.shape[1], p(g), p(q), nq,
                         float(np.sqrt(gp.ell2)), gp.sf2, p(mw), float(gp.hm[D]), int(want_nrm), p(mean), p(var), p(nrm),
                         p(alpha))
    assert rc == 0
    return mean, var, nrm.T, alpha[:R]


def test_emulated_predict_kernel_matches_the_stored_predgrid(emu):
    """The reference's stored loofa gpModel (126 points with normals = 378 rows: two super-blocks of the
    triangular GEMM) predicted at 40 points of its stored predGrid: TSDF, normals and variance against the
    reference's own numbers, alpha against the stored gpModel.alpha."""
    g = load_golden("loofa")
    hyp, X, y, dy, beta = golden_model(g)
    gp = incremental(hyp, X, y, dy, beta, True)
    sel = np.arange(3, 625, 16)[:40]
    mean, var, nrm, alpha = run_emu_predict(emu, gp, g["pred_points"][sel], True)
    assert np.max(np.abs(mean - g["pred_tsdf"][sel])) < 1e-10
    assert np.max(np.abs(var - g["pred_noise"][sel])) < 1e-10
    assert np.max(np.abs(nrm - g["pred_normals"][sel])) < 1e-10
    m, D = X.shape
    a = alpha.reshape(m, D + 1)                                   # point-major -> the reference's dimension-major
    a_dm = np.concatenate([a[:, 0]] + [a[:, 1 + d] for d in range(D)])
    assert np.max(np.abs(a_dm - g["model_alpha"])) / np.max(np.abs(g["model_alpha"])) < 1e-9


def test_emulated_predict_kernel_3d_function_only_ragged(emu):
    """3-D, function values only, 45 rows (a partial k-chunk) and 37 queries (a partial query tile)."""
    rng = np.random.default_rng(9)
    X = rng.uniform(0, 5, (45, 3))
    y = np.linalg.norm(X - 2.4, axis=1) - 1.5
    hyp = {"cov": np.array([np.log(1.6), np.log(1.2)]), "mean": np.array([0.02, 0.0, -0.01, 0.1]), "lik": 0.5 * np.log(0.02)}
    gp = incremental(hyp, X, y, None, np.exp(2 * hyp["lik"]), False)
    q = rng.uniform(0, 5, (37, 3))
    mean, var, nrm, _ = run_emu_predict(emu, gp, q, True)
    fm, fv, fn = O.predict_points(gp, q, want_normals=True)
    assert np.max(np.abs(mean - fm)) < 1e-11 and np.max(np.abs(var - fv)) < 1e-11 and np.max(np.abs(nrm - fn)) < 1e-11


# ------------------------------------------------------------------------------------------------
# gpis_select.cuh: the panel backend's selection loop (arbitrary points, gradient observations)
# ------------------------------------------------------------------------------------------------
def run_emu_panel(lib, shape, train, first, hyp, n_ctas=3):
    vp = C.c_void_p
    lib.emu_panel_select.restype = C.c_int
    lib.emu_panel_select.argtypes = [C.c_int, C.c_int, C.c_longlong, vp, vp, vp, vp, C.c_int, C.c_longlong] + \
        [C.c_double] * 2 + [vp] + [C.c_double] * 5 + [C.c_int] + [vp] * 9
    pts = np.asarray(shape["points"], np.float64)
    N, D = pts.shape
    use_grad = shape.get("normals") is not None and np.size(shape["normals"]) > 0
    use_noise = shape.get("noise") is not None and np.size(shape["noise"]) > 0
    K = int(train["activeSetSize"])
    nb = D + 1 if use_grad else 1
    P = np.ascontiguousarray(pts.T)
    t = np.ascontiguousarray(np.asarray(shape["tsdf"], np.float64).reshape(-1))
    nr = np.ascontiguousarray(np.asarray(shape["normals"], np.float64).reshape(N, D).T) if use_grad else None
    nz = np.ascontiguousarray(np.asarray(shape["noise"], np.float64).reshape(-1)) if use_noise else None
    ids = np.full(K + 1, -1, np.int64)
    n_sel, n_model, err = C.c_int(), C.c_int(), C.c_int()
    mu, var, flags = np.empty(N), np.empty(N), np.empty(N, np.uint8)
    L, g = np.empty((K * nb, K * nb)), np.empty(K * nb)
    mw = np.ascontiguousarray(np.asarray(hyp["mean"], np.float64)[:D])
    p = lambda a: a.ctypes.data_as(vp) if a is not None else None
    rc = lib.emu_panel_select(D, int(use_grad), N, p(P), p(t), p(nr), p(nz), K, int(first), float(np.exp(hyp["cov"][0])),
                              float(np.exp(2 * hyp["cov"][1])), p(mw), float(hyp["mean"][D]), float(np.exp(2 * hyp["lik"])),
                              float(train["levelSet"]), float(train["eps"]), float(train["beta"]), n_ctas, p(ids),
                              C.byref(n_sel), C.byref(n_model), C.byref(err), p(mu), p(var), p(flags), p(L), p(g))
    assert rc == 0
    return ids[:n_sel.value], n_model.value, err.value, mu, var, flags, L, g


def test_emulated_panel_backend_on_the_reference_tape_shape(emu):
    """BASELINE config 1 in kind: the reference's stored 2-D `tape` shape (508 candidates, normals,
    per-point noise) through k_append / k_update -- first 30 picks, factor, g, level-set masks and the
    posterior of the unclassified candidates against the incremental oracle."""
    g = load_golden("results_tape")
    shape = {"points": g["shape_points"], "tsdf": g["shape_tsdf"], "normals": g["shape_normals"], "noise": g["shape_noise"]}
    train = {"activeSetSize": 30, "beta": 10.0, "eps": 1e-2, "levelSet": 0.0}
    hyp = O.default_hyp(2)
    ora = O.select_straddle_incremental(shape, train, 17, hyp=hyp)
    ids, n_model, err, mu, var, flags, L, gv = run_emu_panel(emu, shape, train, 17, hyp)
    assert err == 0 and np.array_equal(ids, ora["order"]) and n_model == ora["n_in_model"] == 29
    r = ora["gp"].r
    assert np.max(np.abs(L[:r, :r] - ora["gp"].L[:r, :r])) < 1e-11 and np.max(np.abs(gv[:r] - ora["gp"].g[:r])) < 1e-11
    assert np.array_equal((flags & 1) > 0, ora["active"])
    assert np.array_equal((flags & 2) > 0, ora["high"]) and np.array_equal((flags & 4) > 0, ora["low"])
    live = (flags == 0)
    fm, fv, _ = O.predict_points(ora["gp"], shape["points"][live])
    assert live.sum() > 50 and np.max(np.abs(mu[live] - fm)) < 1e-10 and np.max(np.abs(var[live] - fv)) < 1e-10


def test_emulated_panel_backend_3d_function_only_with_dead_tiles(emu):
    """3-D points in random order (not a lattice), function values only, several CTAs; candidates far from
    the level set classify early, so whole 64-candidate tiles drop out of the live list."""
    rng = np.random.default_rng(2)
    pts = rng.uniform(0, 9, (700, 3))
    pts = pts[np.argsort(np.linalg.norm(pts - 4.5, axis=1))]            # tiles of similar radius
    tsdf = np.clip((np.linalg.norm(pts - 4.5, axis=1) - 2.5) / 1.5, -1, 1) + 1e-3 * rng.standard_normal(700)
    shape = {"points": pts, "tsdf": tsdf, "normals": None, "noise": None}
    hyp = {"cov": np.array([np.log(1.5), 0.0]), "mean": np.zeros(4), "lik": 0.5 * np.log(0.02)}
    train = {"activeSetSize": 45, "beta": 4.0, "eps": 1e-2, "levelSet": 0.0}
    ora = O.select_straddle_incremental(shape, train, 123, hyp=hyp)
    ids, n_model, err, mu, var, flags, L, gv = run_emu_panel(emu, shape, train, 123, hyp, n_ctas=4)
    assert err == 0 and np.array_equal(ids, ora["order"]) and n_model == ora["n_in_model"]
    assert np.array_equal((flags & 2) > 0, ora["high"]) and np.array_equal((flags & 4) > 0, ora["low"])
    live = (flags == 0)
    fm, fv, _ = O.predict_points(ora["gp"], pts[live])
    assert np.max(np.abs(mu[live] - fm)) < 1e-10 and np.max(np.abs(var[live] - fv)) < 1e-10
    assert ((flags & 6) > 0).sum() > 128                                # enough classified to retire tiles
