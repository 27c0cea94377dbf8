"""GPU parity: the CUDA engine (through the C ABI) against the oracle and the golden vectors."""
import numpy as np
import pytest

from conftest import golden_model, load_golden
from oracle import gpis_oracle as O

pytestmark = pytest.mark.gpu

TRAIN = {"activeSetSize": 150, "beta": 10.0, "eps": 1e-2, "levelSet": 0.0}


@pytest.fixture(scope="module")
def G():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import gpis_b200
    return gpis_b200


def shape_of(g):
    return {"points": g["shape_points"], "tsdf": g["shape_tsdf"], "normals": g["shape_normals"],
            "noise": g["shape_noise"], "gridDim": 25}


def sphere(g, seed=0, ell=3.0):
    from gpis_b200.synth import sphere_tsdf
    pts, tsdf = sphere_tsdf(g, trunc=3.0, seed=seed)
    hyp = {"cov": np.array([np.log(ell), 0.0]), "mean": np.zeros(4), "lik": 0.5 * np.log(0.05)}
    return {"points": pts, "tsdf": tsdf, "normals": None, "noise": None}, hyp


# ---------------------------------------------------------------- golden vectors
@pytest.mark.parametrize("name", ["loofa", "results_tape", "results_marker_occ"])
def test_dense_fit_matches_stored_alpha(G, name):
    g = load_golden(name)
    hyp, X, y, dy, beta = golden_model(g)
    mdl = G.create_gpis(X, y, dy, beta, False, hyp)
    a = mdl.alpha_dimension_major()
    rel = np.max(np.abs(a - g["model_alpha"])) / np.max(np.abs(g["model_alpha"]))
    assert rel < 1e-9, rel


def test_predict_2d_grid_matches_stored_predgrid(G):
    g = load_golden("loofa")
    hyp, X, y, dy, beta = golden_model(g)
    mdl = G.create_gpis(X, y, dy, beta, False, hyp)
    pg, surf = G.predict_2d_grid(mdl, 25, float(g["pred_thresh"]))
    assert np.array_equal(pg["points"], g["pred_points"])
    assert np.max(np.abs(pg["tsdf"] - g["pred_tsdf"])) < 1e-9
    assert np.max(np.abs(pg["normals"] - g["pred_normals"])) < 1e-9
    assert np.max(np.abs(pg["noise"] - g["pred_noise"])) < 1e-9           # fp64 bar: 1e-5 relative
    assert np.allclose(pg["com"], g["pred_com"], rtol=1e-12)
    e = G.evaluate_errors(pg["tsdf"], g["shape_fullTsdf"], 625)
    assert abs(e["meanError"] - g["err_mean"]) < 1e-9 and abs(e["rmsError"] - g["err_rms"]) < 1e-9
    # gp_mean stacks [tsdf; normals(:)] like the reference
    stacked = G.gp_mean(mdl, g["pred_points"], True)
    assert stacked.shape == (625 * 3,)
    assert np.max(np.abs(stacked[625:] - g["pred_normals"].reshape(-1, order="F"))) < 1e-9


# ---------------------------------------------------------------- selection sequence
@pytest.mark.parametrize("first", [17, 300])
def test_straddle_sequence_2d_with_gradients(G, first):
    g = load_golden("results_tape")
    ora = O.select_straddle_incremental(shape_of(g), TRAIN, first)
    mdl, pred, test = G.select_active_set(shape_of(g), dict(TRAIN, activeSetMethod="LevelSet"), first_index=first)
    n = min(len(ora["order"]), len(mdl["order"]))
    div = int(np.argmin(np.r_[ora["order"][:n] == mdl["order"][:n], False]))
    assert np.array_equal(ora["order"], mdl["order"]), f"diverged at step {div}"
    assert mdl["n_in_model"] == ora["n_in_model"]
    assert np.array_equal(pred["high"], ora["high"]) and np.array_equal(pred["low"], ora["low"])
    assert np.array_equal(test, ora["testIndices"])
    mu, var, nrm = O.predict_points(ora["gp"], g["shape_points"][test], want_normals=True)
    assert np.max(np.abs(pred["tsdf"] - mu)) < 1e-8
    assert np.max(np.abs(pred["noise"] - var)) < 1e-8
    assert np.max(np.abs(pred["normals"] - nrm)) < 1e-8
    L = ora["gp"].L[:ora["gp"].r, :ora["gp"].r]
    assert np.max(np.abs(mdl["L"] - L)) < 1e-9
    assert np.max(np.abs(mdl["alpha"] - ora["gp"].alpha())) < 1e-7 * np.max(np.abs(mdl["alpha"]))


def test_straddle_sequence_matches_literal_matlab_order(G):
    """Against the LITERAL restatement (refit in index order each step) on a 2-D reference shape."""
    g = load_golden("results_marker_occ")
    lit = O.select_active_set_straddle(shape_of(g), TRAIN, 40)
    mdl, pred, test = G.select_active_set(shape_of(g), TRAIN, first_index=40)
    assert np.array_equal(lit["order"], mdl["order"])
    assert mdl["n_in_model"] == lit["n_in_model"]
    assert np.max(np.abs(pred["tsdf"] - lit["pred_tsdf"])) < 1e-8
    assert np.max(np.abs(pred["noise"] - lit["pred_noise"])) < 1e-8


@pytest.mark.parametrize("g,K,first", [(20, 160, 4321), (24, 200, 777)])
def test_straddle_sequence_3d_function_only(G, g, K, first):
    shape, hyp = sphere(g)
    tr = {"activeSetSize": K, "beta": 10.0, "eps": 1e-2, "levelSet": 0.0}
    ora = O.select_straddle_incremental(shape, tr, first, hyp=hyp)
    mdl, pred, test = G.select_active_set(shape, tr, first_index=first, hyp=hyp)
    n = min(len(ora["order"]), len(mdl["order"]))
    div = int(np.argmin(np.r_[ora["order"][:n] == mdl["order"][:n], False]))
    assert np.array_equal(ora["order"], mdl["order"]), f"diverged at step {div} of {n}"
    assert mdl["n_in_model"] == ora["n_in_model"]
    assert np.array_equal(pred["high"], ora["high"]) and np.array_equal(pred["low"], ora["low"])
    mu, var, _ = O.predict_points(ora["gp"], shape["points"][test])
    assert np.max(np.abs(pred["tsdf"] - mu)) < 1e-8
    assert np.max(np.abs(pred["noise"] - var) / np.maximum(np.abs(var), 1e-3)) < 1e-5
    # live candidates: the running posterior equals the dense prediction
    live = ~(pred["high"] | pred["low"])
    live[mdl["order"]] = False
    if live.any():
        cm, cv = mdl.engine.candidate_posterior()
        fm, fv, _ = O.predict_points(ora["gp"], shape["points"][live])
        # the running posterior includes the last pick only when it is in the model
        if mdl["n_in_model"] == len(mdl["order"]):
            assert np.max(np.abs(cm[live] - fm)) < 1e-8 and np.max(np.abs(cv[live] - fv)) < 1e-8


def test_k_limit_leaves_last_pick_out_of_model(G):
    g = load_golden("results_tape")
    tr = dict(TRAIN, activeSetSize=20)
    ora = O.select_straddle_incremental(shape_of(g), tr, 5)
    mdl, _, test = G.select_active_set(shape_of(g), tr, first_index=5)
    assert len(mdl["order"]) == 20 and mdl["n_in_model"] == 19          # straddle.m:129-130
    assert np.array_equal(mdl["order"], ora["order"])
    assert test.size == 508 - 20


# ---------------------------------------------------------------- edge cases
def test_edge_sizes_and_early_stop(G):
    rng = np.random.default_rng(5)
    for N in [1, 7, 63, 64, 65, 130]:
        pts = rng.uniform(0, 6, size=(N, 3))
        tsdf = np.clip(np.linalg.norm(pts - 3.0, axis=1) - 1.7, -1, 1)
        shape = {"points": pts, "tsdf": tsdf, "normals": None, "noise": None}
        hyp = {"cov": np.array([np.log(1.3), 0.0]), "mean": np.zeros(4), "lik": 0.5 * np.log(0.02)}
        tr = {"activeSetSize": min(N, 40) if N > 1 else 1, "beta": 4.0, "eps": 0.05, "levelSet": 0.0}
        ora = O.select_straddle_incremental(shape, tr, 0, hyp=hyp)
        mdl, pred, test = G.select_active_set(shape, tr, first_index=0, hyp=hyp)
        assert np.array_equal(ora["order"], mdl["order"]), N
        assert mdl["n_in_model"] == ora["n_in_model"], N
        assert test.size == N - len(mdl["order"])


def test_mean_function_and_sf2(G):
    """Non-zero affine mean and sf2 != 1 (hyp.mean / hyp.cov(2))."""
    rng = np.random.default_rng(11)
    X = rng.uniform(0, 10, size=(40, 2))
    y = np.sin(X[:, 0]) + 0.3 * X[:, 1]
    dy = np.stack([np.cos(X[:, 0]), np.full(40, 0.3)], axis=1)
    nz = rng.uniform(0.01, 0.05, 40)
    hyp = {"cov": np.array([0.4, 0.3]), "mean": np.array([0.1, -0.2, 0.7]), "lik": np.log(0.1)}
    ref = O.create_gpis(X, y, dy, nz, hyp)
    mdl = G.create_gpis(X, y, dy, nz, False, hyp)
    q = rng.uniform(0, 10, size=(77, 2))
    mu_ref, _, Kxxp = O.gp_mean(ref, q, True)
    var_ref = O.gp_var_diag(ref, q, Kxxp, True)
    mu, var, nrm = G.predict_points(mdl, q, True)
    assert np.max(np.abs(mu - mu_ref[:77])) < 1e-9
    assert np.max(np.abs(nrm - mu_ref[77:].reshape(77, 2, order="F"))) < 1e-9
    assert np.max(np.abs(var - var_ref)) < 1e-9
    # function-only + scalar noise (create_gpis.m:92-94 branch, sl = beta)
    ref2 = O.create_gpis(X, y, None, None, hyp)
    mdl2 = G.create_gpis(X, y, None, None, False, hyp)
    mu2, var2, _ = G.predict_points(mdl2, q, False)
    m2, _, K2 = O.gp_mean(ref2, q, False)
    assert np.max(np.abs(mu2 - m2)) < 1e-9
    assert np.max(np.abs(var2 - O.gp_var_diag(ref2, q, K2, False))) < 1e-9


def test_errors_are_reported_not_fatal(G):
    eng = G.GpisEngine(3, 100, 8, ell=2.0)
    with pytest.raises(G._lib.GpisError) as ei:
        eng.select(0, 4)                       # no candidates yet
    assert ei.value.status == G._lib.ERR_STATE
    with pytest.raises(G._lib.GpisError):
        eng.select_begin(1000)
    eng.close()


def test_gpu_active_set_selector_mirror(G, tmp_path):
    """GpuActiveSetSelector.SelectFromGrid / SelectChol semantics (beta schedule, predictive variance,
    sign classification) against the same rule evaluated by the oracle's GP arithmetic."""
    w = h = 18
    rng = np.random.default_rng(4)
    yy, xx = np.indices((h, w)).astype(np.float64)
    tsdf = np.clip((np.hypot(xx - 8.6, yy - 9.3) - 5.2) / 2.0, -1, 1) + 1e-3 * rng.standard_normal((h, w))
    csv = tmp_path / "t.csv"
    np.savetxt(csv, tsdf, delimiter=",", fmt="%.17g")
    sel = G.GpuActiveSetSelector()
    K, sigma, beta, tol, first = 40, 6.0, 0.05, 0.01, 37
    assert sel.SelectFromGrid(str(csv), K, sigma, beta, w, h, 1, 64, tol, first_index=first, out_dir=str(tmp_path / "out"))
    ids = sel.last["ids"]
    # oracle restatement of the same loop
    pts = np.stack([xx.reshape(-1), yy.reshape(-1)], axis=1)
    y = np.loadtxt(csv, delimiter=",").reshape(-1)
    hyp = {"cov": np.array([0.5 * np.log(sigma), 0.0]), "mean": np.zeros(3), "lik": 0.5 * np.log(beta)}
    gp = O.IncrementalGP(pts, hyp, K, False)
    N = pts.shape[0]
    active = np.zeros(N, bool); hi_f = np.zeros(N, bool); lo_f = np.zeros(N, bool)
    order, pending = [first], first
    active[first] = True
    for it in range(K - 1):
        gp.append(pts[pending], y[pending], None, beta)
        unc = ~active & ~hi_f & ~lo_f
        if not unc.any():
            break
        s = np.sqrt(2.0 * np.log(N * np.pi ** 2 * (it + 1) ** 2 / (6.0 * tol)))
        idx = np.flatnonzero(unc)
        var = gp.var[idx] + beta
        mu = gp.mu[idx]
        hi, lo = mu + s * var, mu - s * var
        amb = np.minimum(hi, -lo)
        pending = int(idx[np.flatnonzero(amb == amb.max())[0]])
        active[pending] = True
        order.append(pending)
        hi_f[idx[lo > 0]] = True
        lo_f[idx[hi < 0]] = True
    assert np.array_equal(ids, np.asarray(order))
    a = np.loadtxt(tmp_path / "out" / "alpha.csv", delimiter=",")
    assert np.max(np.abs(a - gp.alpha()[:len(a)])) < 1e-8 * np.max(np.abs(a))
    assert sel.last["errors"].mean > 0


def test_engine_reuse_across_backends_and_calls(G):
    """One handle, many calls: panel -> grid -> panel candidates, repeated selections, a dense fit in
    between; every result must equal a fresh engine's (no stale state survives a re-set)."""
    from gpis_b200.synth import sphere_tsdf
    pts, tsdf = sphere_tsdf(12, trunc=3.0)
    g, K = 12, 50
    cfg = dict(ell=2.5, noise=0.05, level=0.0, eps=1e-2, beta=10.0)

    def fresh(kind):
        e = G.GpisEngine(3, g ** 3, K, **cfg)
        if kind == "grid":
            e.set_grid_candidates((g, g, g), (0, 0, 0), (1, 1, 1), tsdf)
        else:
            e.set_candidates(pts, tsdf)
        e.select(321, K)
        ids, nm = e.active()
        q = pts[::37]
        m, v, _ = e.predict(q)
        return ids, nm, m, v

    ref_ids, ref_nm, ref_m, ref_v = fresh("grid")
    pid, pnm, pm, pv = fresh("panel")
    assert np.array_equal(ref_ids, pid) and ref_nm == pnm
    assert np.max(np.abs(ref_m - pm)) < 1e-9 and np.max(np.abs(ref_v - pv)) < 1e-9
    eng = G.GpisEngine(3, g ** 3, K, **cfg)
    q = pts[::37]
    for kind in ["panel", "grid", "grid", "panel", "fit", "grid"]:
        if kind == "fit":
            eng.fit(pts[ref_ids[:ref_nm]], tsdf[ref_ids[:ref_nm]])
            m, v, _ = eng.predict(q)
            assert np.max(np.abs(m - ref_m)) < 1e-8 and np.max(np.abs(v - ref_v)) < 1e-8
            continue
        if kind == "grid":
            eng.set_grid_candidates((g, g, g), (0, 0, 0), (1, 1, 1), tsdf)
        else:
            eng.set_candidates(pts, tsdf)
        eng.select(321, K)
        ids, nm = eng.active()
        assert np.array_equal(ids, ref_ids) and nm == ref_nm, kind
        m, v, _ = eng.predict(q)
        assert np.max(np.abs(m - ref_m)) < 1e-9 and np.max(np.abs(v - ref_v)) < 1e-9, kind
        amb, hi, lo, ac = eng.levelset()
        assert ac.sum() == len(ids) and not (hi & lo & ~ac).all()
    # argument errors are statuses
    with pytest.raises(G._lib.GpisError):
        eng.select(321, K + 5)
    with pytest.raises(G._lib.GpisError):
        eng.select(g ** 3 + 7, K)
    eng.close()


def test_float_entry_points_match_the_double_ones(G):
    """The fp32 ABI of the reference (gpu_active_set_selector.hpp:41-46): float candidates in, float active set and
    posterior out -- the same selection as the double entry points on the float-rounded inputs."""
    import ctypes as C
    from gpis_b200 import _lib
    lib = _lib.load()
    g, K = 14, 40
    shape, hyp = sphere(g, ell=2.5)
    pts32 = np.ascontiguousarray(shape["points"].T.astype(np.float32))          # SoA like the reference
    t32 = shape["tsdf"].astype(np.float32)
    N = g ** 3
    ref = G.GpisEngine(3, N, K, ell=2.5, noise=0.05, eps=1e-2, beta=10.0)
    ref.set_candidates(pts32.astype(np.float64), t32.astype(np.float64), soa=True)
    ref.select(100, K)
    ids_ref, nm_ref = ref.active()
    mu_ref, var_ref, _ = ref.predict(shape["points"][:500])
    for lattice in (False, True):
        eng = G.GpisEngine(3, N, K, ell=2.5, noise=0.05, eps=1e-2, beta=10.0)
        if lattice:
            dims = np.array([g, g, g], np.int32)
            org, sp = np.zeros(3), np.ones(3)
            eng._ck(lib.gpis_set_grid_candidates_f32(eng.h, dims.ctypes.data, org.ctypes.data, sp.ctypes.data, 0, g,
                                                     t32.ctypes.data, None, None, None))
        else:
            eng._ck(lib.gpis_set_candidates_f32(eng.h, pts32.ctypes.data, t32.ctypes.data, None, None, None, None))
        eng.select(100, K)
        ids, nm = eng.active()
        assert np.array_equal(ids, ids_ref) and nm == nm_ref
        ain = np.full(3 * K, -1, np.float32)
        atg = np.full(K, -1, np.float32)
        ns, nmm = C.c_int32(), C.c_int32()
        eng._ck(lib.gpis_get_active_f32(eng.h, K, ain.ctypes.data, atg.ctypes.data, C.byref(ns), C.byref(nmm), None))
        assert ns.value == len(ids) and nmm.value == nm
        assert np.array_equal(ain.reshape(3, K)[:, :nm], pts32[:, ids[:nm]])           # active_inputs[a + d*maxSize]
        assert np.array_equal(atg[:nm], t32[ids[:nm]]) and np.all(atg[nm:] == 0)
        q32 = np.ascontiguousarray(shape["points"][:500].T.astype(np.float32))
        m32, v32 = np.empty(500, np.float32), np.empty(500, np.float32)
        eng._ck(lib.gpis_predict_f32(eng.h, q32.ctypes.data, 500, m32.ctypes.data, v32.ctypes.data, None, None))
        assert np.array_equal(m32, np.asarray(mu_ref).astype(np.float32)) or np.max(np.abs(m32 - mu_ref)) < 1e-6
        assert np.max(np.abs(v32 - var_ref)) < 1e-6
        if lattice:
            pm, pv = np.empty(N, np.float32), np.empty(N, np.float32)
            eng._ck(lib.gpis_get_candidate_posterior_f32(eng.h, pm.ctypes.data, pv.ctypes.data, None))
            mu, var = eng.candidate_posterior()
            assert np.array_equal(pm, mu.astype(np.float32)) and np.array_equal(pv, var.astype(np.float32))
        eng.close()


def test_levelset_ambiguity_agrees_across_backends_on_unclassified_candidates(G):
    """gpis_get_levelset: the straddle score of every still-unclassified candidate is the same on the panel and the
    lattice backend (and equals the oracle's); classified candidates differ by design (see the header)."""
    shape, hyp = sphere(18, ell=3.0)
    tr = {"activeSetSize": 90, "beta": 10.0, "eps": 1e-2, "levelSet": 0.0}
    ora = O.select_straddle_incremental(shape, tr, 1234, hyp=hyp, prune=False)
    out = {}
    for backend in ("panel", "grid"):
        mdl, pred, _ = G.select_active_set(shape, tr, first_index=1234, hyp=hyp, backend=backend, predict=False)
        amb, hi, lo, ac = mdl.engine.levelset()
        out[backend] = (amb, hi | lo | ac)
        assert np.array_equal(mdl["order"], ora["order"])
    unc = ~out["panel"][1]
    assert np.array_equal(out["panel"][1], out["grid"][1]) and unc.sum() > 100
    # the last pick is selected but not in the model (straddle.m:129-130): scores are those of the last scoring
    assert np.max(np.abs(out["panel"][0][unc] - out["grid"][0][unc])) < 1e-9
    assert np.max(np.abs(out["grid"][0][unc] - ora["ambiguity"][unc])) < 1e-9
