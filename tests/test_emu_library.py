"""The WHOLE library under the CPU emulation: gpis_b200/csrc/gpis_capi.cu compiled by g++ with -DGPIS_EMULATE
against tests/emu (execution model + CUDA-runtime stand-ins) into build/libgpis_b200_emu.so, driven through
the unmodified Python host layer.  What nvcc compiles is untouched by this (its preprocessed translation
unit and the library's SASS were compared before and after the refactor that made it possible), and the
package never loads this library: the tests swap `_lib.LIB_PATH` themselves.  Small sizes only -- every CUDA
thread is an OS thread here.  The same checks at real sizes are the `-m gpu` tests."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden_model, load_golden
from oracle import gpis_oracle as O

EMU = os.path.join(ROOT, "tests", "emu")
SRC = os.path.join(ROOT, "gpis_b200", "csrc")
EMU_LIB = os.path.join(ROOT, "build", "libgpis_b200_emu.so")


@pytest.fixture(scope="module")
def G():
    deps = [os.path.join(SRC, f) for f in os.listdir(SRC)] + [os.path.join(EMU, f) for f in ("cuda_emu.h", "cuda_rt_emu.h")]
    os.makedirs(os.path.dirname(EMU_LIB), exist_ok=True)
    if not os.path.exists(EMU_LIB) or any(os.path.getmtime(d) > os.path.getmtime(EMU_LIB) for d in deps):
        subprocess.run(["g++", "-std=c++20", "-O2", "-fPIC", "-shared", "-pthread", "-x", "c++", "-DGPIS_EMULATE",
                        "-Wno-unknown-pragmas", "-I" + EMU, "-o", EMU_LIB, os.path.join(SRC, "gpis_capi.cu")], check=True)
    import gpis_b200
    from gpis_b200 import _lib
    saved = (_lib.LIB_PATH, _lib._lib)
    _lib.LIB_PATH, _lib._lib = EMU_LIB, None
    os.environ.setdefault("GPIS_EMU_SMS", "3")
    _lib.load()
    yield gpis_b200
    _lib.LIB_PATH, _lib._lib = saved


def lattice_points(dims):
    idx = np.indices(tuple(dims)[::-1]).reshape(len(dims), -1)[::-1]
    return np.ascontiguousarray(idx.T.astype(np.float64))


def blob(dims, seed, with_normals=False):
    rng = np.random.default_rng(seed)
    pts = lattice_points(dims)
    c = np.array(dims) * rng.uniform(0.42, 0.58, len(dims))
    dist = np.linalg.norm(pts - c, axis=1)
    r, trunc = 0.3 * min(d for d in dims if d > 1) + 0.4, 2.0
    sd = (dist - r) / trunc
    tsdf = np.clip(sd, -1, 1) + 1e-3 * rng.standard_normal(len(pts))
    if not with_normals:
        return pts, tsdf
    nrm = (pts - c) / np.maximum(dist, 1e-12)[:, None] / trunc
    nrm[np.abs(sd) >= 1.0] = 0.0
    return pts, tsdf, nrm


TRAIN = {"beta": 10.0, "eps": 1e-2, "levelSet": 0.0}


def test_grid_backend_3d_through_the_emulated_library(G):
    """Lattice backend, function values: one-pass solve kernels, the warp-specialised bulk-copy / mbarrier /
    DMMA update GEMM (stream-K over 3 'SMs'), the epilogue -- picks, model, full posterior grid, masks."""
    dims = (10, 9, 4)
    pts, tsdf = blob(dims, 1)
    hyp = {"cov": np.array([np.log(2.0), 0.0]), "mean": np.array([0.01, 0.0, -0.01, 0.05]), "lik": 0.5 * np.log(0.05)}
    tr = dict(TRAIN, activeSetSize=12)
    shape = {"points": pts, "tsdf": tsdf, "normals": None, "noise": None}
    ora = O.select_straddle_incremental(shape, tr, 17, hyp=hyp, prune=False)
    mdl, pred, test = G.select_active_set(shape, tr, first_index=17, hyp=hyp, backend="grid")
    assert np.array_equal(mdl["order"], ora["order"]) and mdl["n_in_model"] == ora["n_in_model"]
    assert np.array_equal(pred["high"], ora["high"]) and np.array_equal(pred["low"], ora["low"])
    mu, var = mdl.engine.candidate_posterior()
    fm, fv, _ = O.predict_points(ora["gp"], pts)
    assert np.max(np.abs(mu - fm)) < 1e-10 and np.max(np.abs(var - fv)) < 1e-10
    L, alpha = mdl.engine.factor()
    r = ora["gp"].r
    assert np.max(np.abs(L - ora["gp"].L[:r, :r])) < 1e-11
    assert mdl.engine.stats()["kernel_launches"] > 3 * 10


def test_grid_backend_gradients_2d_through_the_emulated_library(G):
    """2-D lattice with normal observations: the two-pass append (point, solve, W-row kernels), 3 rows per
    point in the update GEMM."""
    dims = (9, 8)
    pts, tsdf, nrm = blob(dims, 2, with_normals=True)
    hyp = {"cov": np.array([np.log(1.8), 0.0]), "mean": np.zeros(3), "lik": 0.5 * np.log(0.02)}
    tr = dict(TRAIN, activeSetSize=8)
    shape = {"points": pts, "tsdf": tsdf, "normals": nrm, "noise": None}
    ora = O.select_straddle_incremental(shape, tr, 5, hyp=hyp, prune=False)
    mdl, pred, test = G.select_active_set(shape, tr, first_index=5, hyp=hyp, backend="grid")
    assert np.array_equal(mdl["order"], ora["order"]) and mdl["n_in_model"] == ora["n_in_model"]
    mu, var = mdl.engine.candidate_posterior()
    fm, fv, fn = O.predict_points(ora["gp"], pts, want_normals=True)
    assert np.max(np.abs(mu - fm)) < 1e-10 and np.max(np.abs(var - fv)) < 1e-10
    assert np.max(np.abs(pred["normals"] - fn[test])) < 1e-10


def test_grid_backend_gradients_3d_row_lists_prune_through_the_emulated_library(G):
    """3-D lattice with normal observations and a length scale much shorter than the z extent: the per-step path's
    ROW LISTS (k_grid_lists + the list form of k_grid_gemm) leave out the model rows farther than 9.85 ell from a
    slice -- the posterior still equals the oracle's dense sums."""
    dims = (5, 4, 26)
    pts, tsdf, nrm = blob(dims, 4, with_normals=True)
    hyp = {"cov": np.array([np.log(0.7), 0.0]), "mean": np.zeros(4), "lik": 0.5 * np.log(0.02)}
    tr = dict(TRAIN, activeSetSize=9)
    shape = {"points": pts, "tsdf": tsdf, "normals": nrm, "noise": None}
    ora = O.select_straddle_incremental(shape, tr, 11, hyp=hyp, prune=False)
    mdl, pred, test = G.select_active_set(shape, tr, first_index=11, hyp=hyp, backend="grid")
    assert np.array_equal(mdl["order"], ora["order"]) and mdl["n_in_model"] == ora["n_in_model"]
    zs = pts[mdl["order"][:mdl["n_in_model"]], 2]
    assert zs.max() - zs.min() > 2 * 9.85 * 0.7           # some rows are out of reach of some slices
    mu, var = mdl.engine.candidate_posterior()
    fm, fv, fn = O.predict_points(ora["gp"], pts, want_normals=True)
    assert np.max(np.abs(mu - fm)) < 1e-10 and np.max(np.abs(var - fv)) < 1e-10
    assert np.array_equal(pred["high"], ora["high"]) and np.array_equal(pred["low"], ora["low"])


def test_dense_fit_predict_and_sampling_through_the_emulated_library(G):
    """create_gpis (the append kernel fed by gpis_fit), gp_mean / gp_cov / predict kernels, the full posterior
    covariance and the draws, on a third of the reference's stored loofa model against its stored predGrid
    positions (oracle values: the subset is not what the reference fitted)."""
    g = load_golden("loofa")
    hyp, X, y, dy, beta = golden_model(g)
    keep = np.arange(0, X.shape[0], 3)
    X, y, dy, beta = X[keep], y[keep], dy[keep], beta[keep]
    ora = O.create_gpis(X, y, dy, beta, hyp)
    mdl = G.create_gpis(X, y, dy, beta, False, hyp)
    a = mdl.alpha_dimension_major()
    assert np.max(np.abs(a - ora["alpha"])) / np.max(np.abs(ora["alpha"])) < 1e-9
    q = g["pred_points"][7::29][:20]
    mu, _, _ = O.gp_mean(ora, q, True)
    assert np.max(np.abs(G.gp_mean(mdl, q, True) - mu)) < 1e-10
    assert np.max(np.abs(G.gp_cov(mdl, q) - O.gp_var_diag(ora, q))) < 1e-10
    COV = O.gp_cov(ora, q, None, True)
    assert np.max(np.abs(G.gp_cov(mdl, q, None, True, full=True) - COV)) < 1e-11
    z = np.random.default_rng(0).standard_normal((2, 60))
    ref, lp, _ = O.sample_from_posterior(mu, COV, z, 1e-10)
    draws, logpdf = mdl.engine.sample_shapes(q, z, use_derivative=True, jitter=1e-10)
    assert np.max(np.abs(draws - ref)) < 1e-7 and np.max(np.abs(logpdf - lp)) < 1e-6 * np.max(np.abs(lp))


def test_batch_and_panel_through_the_emulated_library(G):
    """gpis_batch_select and the panel backend through the C ABI (host buffers in and out), and the error
    convention: statuses, never exits."""
    from gpis_b200.batch import select_batch_fused
    dims = (7, 6, 5)
    objs = [blob(dims, 10 + k)[1] for k in range(4)]
    pts = lattice_points(dims)
    cfg = dict(ell=2.0, noise=0.05, level=0.0, eps=1e-2, beta=10.0)
    hyp = {"cov": np.array([np.log(2.0), 0.0]), "mean": np.zeros(4), "lik": 0.5 * np.log(0.05)}
    tr = dict(TRAIN, activeSetSize=9)
    res = select_batch_fused(objs, dims, 9, 11, **cfg)
    from gpis_b200 import _lib as glib
    for form in (glib.BATCH_PLAIN, glib.BATCH_DMMA):     # the other forms of the kernel (gpis_config.batch_kernel)
        other = select_batch_fused(objs, dims, 9, 11, kernel=form, **cfg)
        for a, b in zip(res, other):
            assert np.array_equal(a["ids"], b["ids"]) and np.max(np.abs(a["mean"] - b["mean"])) < 1e-12
    for k in (0, 3):
        shape = {"points": pts, "tsdf": objs[k], "normals": None, "noise": None}
        ora = O.select_straddle_incremental(shape, tr, 11, hyp=hyp, prune=False)
        assert np.array_equal(res[k]["ids"], ora["order"]) and res[k]["n_in_model"] == ora["n_in_model"]
        fm, fv, _ = O.predict_points(ora["gp"], pts)
        assert np.max(np.abs(res[k]["mean"] - fm)) < 1e-10 and np.max(np.abs(res[k]["var"] - fv)) < 1e-10
    shape = {"points": pts, "tsdf": objs[1], "normals": None, "noise": None}
    ora = O.select_straddle_incremental(shape, tr, 11, hyp=hyp)
    mdl, pred, test = G.select_active_set(shape, tr, first_index=11, hyp=hyp, backend="panel")
    assert np.array_equal(mdl["order"], ora["order"]) and np.array_equal(mdl["order"], res[1]["ids"])
    mu, var, _ = O.predict_points(ora["gp"], pts[test])
    assert np.max(np.abs(pred["tsdf"] - mu)) < 1e-10 and np.max(np.abs(pred["noise"] - var)) < 1e-10
    with pytest.raises(G._lib.GpisError) as ei:
        mdl.engine.select(10 ** 9, 5)
    assert ei.value.status == G._lib.ERR_INVALID
    with pytest.raises(G._lib.GpisError):
        G.GpisEngine(3, 10, 4, ell=2.0, precision=G._lib.PREC_TF32)          # tcgen05 is not emulated


def test_tsdf_builder_and_float_entry_points_through_the_emulated_library(G):
    """gpis_build_tsdf (auto_tsdf.m:133-262) bit for bit against the oracle on a stored reference shape and a
    random image, and the fp32 entry points of the reference ABI, on the CPU emulation."""
    import ctypes as C
    from oracle import tsdf_oracle as T
    from test_oracle_tsdf import REF_VP
    z = np.load(os.path.join(ROOT, "tests", "golden", "tsdf_shapes.npz"))
    inside = z["tape_inside"]
    u = np.random.default_rng(2).uniform(size=(3,) + inside.shape)
    sp, ref = G.auto_tsdf(inside, REF_VP, u=u), T.auto_tsdf(inside, REF_VP, u)
    assert np.array_equal(sp["fullTsdf"], z["tape_fullTsdf"]) and np.array_equal(sp["fullNormals"], z["tape_fullNormals"])
    rng = np.random.default_rng(5)
    img = rng.uniform(size=(13, 9)) < 0.6
    img[3:9, 2:7] = True
    u2 = rng.uniform(size=(3, 13, 9))
    vp = dict(REF_VP, noiseGradMode="BLTR", specularNoise=True, sparsityRate=0.3, edgeWin=1, horizScale=0.7, vertScale=1.3,
              y_thresh1_low=1.5, y_thresh1_high=6, x_thresh1_low=0, x_thresh1_high=3)
    for a, b in ((sp, ref), (G.auto_tsdf(img, vp, u=u2), T.auto_tsdf(img, vp, u2))):
        for k in ("tsdf", "noise", "normals", "points", "fullTsdf", "fullNormals", "fullNoise", "validIndices"):
            assert np.array_equal(a[k], b[k]), k
    # float ABI: candidates in as float, active set out as float
    from gpis_b200 import _lib
    lib = _lib.load()
    dims = (6, 5, 3)
    pts, tsdf = blob(dims, 3)
    N, K = len(pts), 8
    t32 = tsdf.astype(np.float32)
    d32 = np.array(dims, np.int32)
    org, spc = np.zeros(3), np.ones(3)
    ref_eng = G.GpisEngine(3, N, K, ell=2.0, noise=0.05)
    ref_eng.set_grid_candidates(dims, (0, 0, 0), (1, 1, 1), t32.astype(np.float64))
    ref_eng.select(7, K)
    eng = G.GpisEngine(3, N, K, ell=2.0, noise=0.05)
    eng._ck(lib.gpis_set_grid_candidates_f32(eng.h, d32.ctypes.data, org.ctypes.data, spc.ctypes.data, 0, dims[2],
                                             t32.ctypes.data, None, None, None))
    eng.select(7, K)
    ids, nm = eng.active()
    assert np.array_equal(ids, ref_eng.active()[0])
    ain, atg = np.full(3 * K, -1, np.float32), np.full(K, -1, np.float32)
    ns, nmm = C.c_int32(), C.c_int32()
    eng._ck(lib.gpis_get_active_f32(eng.h, K, ain.ctypes.data, atg.ctypes.data, C.byref(ns), C.byref(nmm), None))
    assert ns.value == len(ids) and nmm.value == nm
    assert np.array_equal(ain.reshape(3, K)[:, :nm], pts[ids[:nm]].T.astype(np.float32))
    assert np.array_equal(atg[:nm], t32[ids[:nm]])
    pm, pv = np.empty(N, np.float32), np.empty(N, np.float32)
    eng._ck(lib.gpis_get_candidate_posterior_f32(eng.h, pm.ctypes.data, pv.ctypes.data, None))
    mu, var = eng.candidate_posterior()
    assert np.array_equal(pm, mu.astype(np.float32)) and np.array_equal(pv, var.astype(np.float32))
