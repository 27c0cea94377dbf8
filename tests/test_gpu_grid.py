"""GPU parity of the grid (separable, DMMA) backend against the oracle."""
import numpy as np
import pytest

from oracle import gpis_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import torch
    assert torch.cuda.is_available()
    import gpis_b200
    return gpis_b200


def sphere(g, ell=3.0, with_normals=False):
    from gpis_b200.synth import sphere_tsdf
    out = sphere_tsdf(g, trunc=3.0, with_normals=with_normals)
    hyp = {"cov": np.array([np.log(ell), 0.0]), "mean": np.zeros(4), "lik": 0.5 * np.log(0.05)}
    shape = {"points": out[0], "tsdf": out[1], "normals": out[2] if with_normals else None, "noise": None}
    return shape, hyp


def first_divergence(a, b):
    n = min(len(a), len(b))
    return int(np.argmin(np.r_[np.asarray(a[:n]) == np.asarray(b[:n]), False]))


@pytest.mark.parametrize("g,K,first", [(20, 160, 4321), (24, 200, 777)])
def test_grid_backend_sequence_and_posterior_grid(G, g, K, first):
    shape, hyp = sphere(g)
    tr = {"activeSetSize": K, "beta": 10.0, "eps": 1e-2, "levelSet": 0.0}
    ora = O.select_straddle_incremental(shape, tr, first, hyp=hyp, prune=False)
    mdl, pred, test = G.select_active_set(shape, tr, first_index=first, hyp=hyp, backend="grid")
    assert mdl["backend"] == "grid"
    assert np.array_equal(ora["order"], mdl["order"]), f"diverged at {first_divergence(ora['order'], mdl['order'])}"
    assert mdl["n_in_model"] == ora["n_in_model"]
    assert np.array_equal(pred["high"], ora["high"]) and np.array_equal(pred["low"], ora["low"])
    # the running posterior IS the full posterior grid (no candidate is ever frozen)
    mu, var = mdl.engine.candidate_posterior()
    fm, fv, _ = O.predict_points(ora["gp"], shape["points"])
    assert np.max(np.abs(mu - fm)) < 1e-8
    assert np.max(np.abs(var - fv) / np.maximum(np.abs(fv), 1e-3)) < 1e-5          # north_star fp64 bar
    # factor and dense prediction through W' maintained by the grid backend
    L = ora["gp"].L[:ora["gp"].r, :ora["gp"].r]
    assert np.max(np.abs(mdl["L"] - L)) < 1e-9
    assert np.max(np.abs(pred["tsdf"] - fm[test])) < 1e-8 and np.max(np.abs(pred["noise"] - fv[test])) < 1e-8
    # same picks as the panel backend
    mdl2, _, _ = G.select_active_set(shape, tr, first_index=first, hyp=hyp, backend="panel")
    assert mdl2["backend"] == "panel" and np.array_equal(mdl2["order"], mdl["order"])


def test_grid_backend_with_gradient_observations(G):
    shape, hyp = sphere(12, ell=2.5, with_normals=True)
    tr = {"activeSetSize": 70, "beta": 10.0, "eps": 1e-2, "levelSet": 0.0}
    ora = O.select_straddle_incremental(shape, tr, 500, hyp=hyp, prune=False)
    mdl, pred, test = G.select_active_set(shape, tr, first_index=500, hyp=hyp, backend="grid")
    assert np.array_equal(ora["order"], mdl["order"]), f"diverged at {first_divergence(ora['order'], mdl['order'])}"
    mu, var = mdl.engine.candidate_posterior()
    fm, fv, fn = O.predict_points(ora["gp"], shape["points"], want_normals=True)
    assert np.max(np.abs(mu - fm)) < 1e-8 and np.max(np.abs(var - fv)) < 1e-8
    assert np.max(np.abs(pred["normals"] - fn[test])) < 1e-8


def test_grid_backend_2d_lattice_and_odd_sizes(G):
    rng = np.random.default_rng(2)
    for dims in [(25, 25), (7, 5), (33, 9), (130, 3)]:
        nx, ny = dims
        ii = np.indices((ny, nx)).reshape(2, -1)[::-1].T.astype(np.float64) + 1.0     # 1-based like meshgrid(1:g)
        c = np.array([nx * 0.52, ny * 0.47])
        tsdf = np.clip((np.linalg.norm(ii - c, axis=1) - min(dims) * 0.3) / 2.0, -1, 1) + 1e-3 * rng.standard_normal(ii.shape[0])
        shape = {"points": ii, "tsdf": tsdf, "normals": None, "noise": None}
        hyp = {"cov": np.array([0.6, 0.0]), "mean": np.zeros(3), "lik": 0.5 * np.log(0.05)}
        tr = {"activeSetSize": min(40, ii.shape[0]), "beta": 10.0, "eps": 1e-2, "levelSet": 0.0}
        ora = O.select_straddle_incremental(shape, tr, 3, hyp=hyp, prune=False)
        mdl, pred, test = G.select_active_set(shape, tr, first_index=3, hyp=hyp, backend="grid")
        assert np.array_equal(ora["order"], mdl["order"]), dims
        mu, var = mdl.engine.candidate_posterior()
        fm, fv, _ = O.predict_points(ora["gp"], ii)
        assert np.max(np.abs(mu - fm)) < 1e-8 and np.max(np.abs(var - fv)) < 1e-8


def test_two_slab_shards_on_one_gpu_match_single_engine(G):
    """world_size = 2 emulated in one process: the records are exchanged by device copies."""
    import torch
    from gpis_b200.dist import engine_exchange_tensors
    g, K, first = 16, 90, 1000
    shape, hyp = sphere(g)
    tr = {"activeSetSize": K, "beta": 10.0, "eps": 1e-2, "levelSet": 0.0}
    ref, _, _ = G.select_active_set(shape, tr, first_index=first, hyp=hyp, backend="grid")
    cfg = dict(ell=float(np.exp(hyp["cov"][0])), noise=0.05, level=0.0, eps=1e-2, beta=10.0)
    z_split = 7
    engs, views = [], []
    for r, (z0, z1) in enumerate([(0, z_split), (z_split, g)]):
        n = g * g * (z1 - z0)
        e = G.GpisEngine(3, n, K, rank=r, world_size=2, total_candidates=g ** 3, id_base=g * g * z0, **cfg)
        e.set_grid_candidates((g, g, z1 - z0), (0, 0, 0), (1, 1, 1), shape["tsdf"][g * g * z0:g * g * z1], z0=z0, nz_total=g)
        engs.append(e)
        views.append(engine_exchange_tensors(e, 2))

    def gather():
        for send, recv in views:
            n = send.numel()
            for r2, (s2, _) in enumerate(views):
                recv[r2 * n:(r2 + 1) * n].copy_(s2)
        torch.cuda.synchronize()

    for r, e in enumerate(engs):
        lo = g * g * (0 if r == 0 else z_split)
        hi = g * g * (z_split if r == 0 else g)
        e.select_begin(first - lo if lo <= first < hi else -1)
    gather()
    for _ in range(K - 1):
        for e in engs:
            e.step_append()
            e.step_update()
        gather()
    for e in engs:
        e.select_end()
    ids0, m0 = engs[0].active()
    ids1, m1 = engs[1].active()
    assert np.array_equal(ids0, ids1) and m0 == m1
    assert np.array_equal(ids0, ref["order"]) and m0 == ref["n_in_model"]
    mu = np.concatenate([e.candidate_posterior()[0] for e in engs])
    rmu, _ = ref.engine.candidate_posterior()
    assert np.max(np.abs(mu - rmu)) < 1e-10


def test_c2_size_two_backends_agree_and_match_cpu_prefix(G):
    """BASELINE config 2 at full size (64^3 sphere, 262144 candidates, 1000-point active set):
    the two independent CUDA implementations (HBM panel backend, separable DMMA grid backend) pick
    the same 1000 points; the first 150 picks match the CPU restatement; size-independent
    properties of the posterior hold on the whole grid."""
    from gpis_b200.synth import sphere_tsdf, first_index
    from oracle import cpu_baseline as B
    pts, tsdf = sphere_tsdf(64)
    hyp = {"cov": np.array([np.log(4.0), 0.0]), "mean": np.zeros(4), "lik": 0.5 * np.log(0.05)}
    tr = {"activeSetSize": 1000, "beta": 10.0, "eps": 1e-2, "levelSet": 0.0}
    shape = {"points": pts, "tsdf": tsdf, "normals": None, "noise": None}
    first = first_index(pts.shape[0])
    assert first == 36022
    mg, pg, tg = G.select_active_set(shape, tr, first_index=first, hyp=hyp, backend="grid", predict=False)
    mp, pp, tp = G.select_active_set(shape, tr, first_index=first, hyp=hyp, backend="panel", predict=False)
    assert np.array_equal(mg["order"], mp["order"]) and len(mg["order"]) == 1000 and mg["n_in_model"] == 999
    assert np.array_equal(pg["high"], pp["high"]) and np.array_equal(pg["low"], pp["low"])
    cpu = B.select_compact(pts, tsdf, hyp, tr, first, max_steps=150)
    assert np.array_equal(cpu["order"], mg["order"][:len(cpu["order"])])
    mu, var = mg.engine.candidate_posterior()
    assert var.min() > -1e-9 and var.max() <= 1.0 + 1e-12      # latent variance within [0, sf2]
    act = mg["order"][:mg["n_in_model"]]
    assert np.all(var[act] < 0.05 + 1e-9)                       # at an observed point var < noise
    assert np.max(np.abs(mu[act] - tsdf[act])) < 0.2            # and the mean interpolates within noise
    # dense prediction through W' (grid) and through the inverted factor (panel) agree on a sample
    q = pts[::997]
    m1, v1, _ = mg.engine.predict(q)
    m2, v2, _ = mp.engine.predict(q)
    assert np.max(np.abs(m1 - m2)) < 1e-9 and np.max(np.abs(v1 - v2)) < 1e-9
    assert np.max(np.abs(m1 - mu[::997])) < 1e-9 and np.max(np.abs(v1 - var[::997])) < 1e-9


def test_real_sdf_fixture_through_the_grid_backend(G):
    """The reference's own 25^3 SDF (data/test/sdf/Co_clean_dim_25.sdf): selection + posterior grid
    on its lattice against the oracle, and the Gpis3D host class against the oracle's dense fit."""
    from conftest import load_golden
    from gpis_b200.sdf_io import Sdf3D, Gpis3D, select_from_sdf
    from gpis_b200.synth import grid_points
    g = load_golden("sdf_co_clean_dim_25")
    dims = tuple(int(v) for v in g["dims"])
    sdf = Sdf3D(g["values"].astype(np.float64).reshape(dims, order="F"), g["origin"], float(g["resolution"]))
    K, trunc, ell = 120, 4.0 * sdf.resolution, 3.0
    out = select_from_sdf(sdf, K, trunc=trunc, ell=ell, first_index=5000)
    pts = grid_points(dims)
    hyp = {"cov": np.array([np.log(ell), 0.0]), "mean": np.zeros(4), "lik": 0.5 * np.log(0.05)}
    tr = {"activeSetSize": K, "beta": 10.0, "eps": 1e-2, "levelSet": 0.0}
    ora = O.select_straddle_incremental({"points": pts, "tsdf": sdf.tsdf(trunc), "normals": None, "noise": None},
                                        tr, 5000, hyp=hyp, prune=False)
    assert np.array_equal(out["ids"], ora["order"]) and out["n_in_model"] == ora["n_in_model"]
    fm, fv, _ = O.predict_points(ora["gp"], pts)
    assert np.max(np.abs(out["mean"].reshape(-1, order="F") - fm)) < 1e-8
    assert np.max(np.abs(out["var"].reshape(-1, order="F") - fv)) < 1e-8
    # Gpis3D(points, sdf_meas, dims, ...): dense fit on a random subset + predict_locations
    rng = np.random.default_rng(0)
    sub = rng.choice(pts.shape[0], 300, replace=False)
    gp3 = Gpis3D(pts[sub], sdf.tsdf(trunc)[sub], (6, 5, 4), meas_variance=0.05, kernel_hyps=[1.0, ell])
    ref = O.create_gpis(pts[sub], sdf.tsdf(trunc)[sub], None, None, hyp)
    q = gp3.grid_pts_
    m_ref, _, Kx = O.gp_mean(ref, q, False)
    assert gp3.mean_sdf().shape == (6, 5, 4)
    assert np.max(np.abs(gp3.mean_pred_[:, 0] - m_ref)) < 1e-8
    assert np.max(np.abs(gp3.var_pred_[:, 0] - (O.gp_var_diag(ref, q, Kx, False) + 0.05))) < 1e-8


def test_cube_tiles_with_per_point_noise_and_predictive_variance(G):
    """The persistent kernel's 16 x 16 x 16 cubes on a lattice that ends in partial cubes (24 x 20 x 24), per-point
    noise and the predictive-variance score (max_subset_buffers.cu:113-115: the cell's own noise joins the variance
    before scoring) -- picks, model size and level-set masks against the oracle loop; the per-step kernels agree."""
    from gpis_b200.synth import grid_points
    dims = (24, 20, 24)
    pts = grid_points(dims)
    c = np.array([11.3, 9.6, 12.2])
    tsdf = np.clip(np.linalg.norm((pts - c) * np.array([1.0, 1.2, 0.9]), axis=1) - 7.5, -3.0, 3.0)
    rng = np.random.default_rng(11)
    noise = 0.02 + 0.06 * rng.uniform(size=pts.shape[0])
    hyp = {"cov": np.array([np.log(3.0), 0.0]), "mean": np.zeros(4), "lik": 0.5 * np.log(0.05)}
    shape = {"points": pts, "tsdf": tsdf, "normals": None, "noise": noise}
    tr = {"activeSetSize": 130, "beta": 10.0, "eps": 2e-2, "levelSet": 0.0, "predictiveVariance": True}
    ora = O.select_straddle_incremental(shape, tr, 777, hyp=hyp, prune=False)
    mdl, pred, _ = G.select_active_set(shape, tr, first_index=777, hyp=hyp, backend="grid", predict=False)
    assert mdl.engine.stats()["kernel_launches"] < 40          # the persistent kernel ran (one launch per selection)
    assert np.array_equal(mdl["order"], ora["order"]), first_divergence(mdl["order"], ora["order"])
    assert mdl["n_in_model"] == ora["n_in_model"]
    assert np.array_equal(pred["high"], ora["high"]) and np.array_equal(pred["low"], ora["low"])
    eng = G.GpisEngine(3, pts.shape[0], 130, ell=3.0, noise=0.05, level=0.0, eps=2e-2, beta=10.0,
                       variance_mode=G._lib.VAR_PREDICTIVE, loop_mode=G._lib.LOOP_STEPS)
    eng.set_grid_candidates(dims, (0, 0, 0), (1, 1, 1), tsdf, noise=noise)
    eng.select(777, 130)
    assert np.array_equal(eng.active()[0], ora["order"])


def test_gpis2d_host_class_matches_the_oracle_dense_fit(G):
    """Gpis2D (src/grasp_selection/gpis.py:125-160 shape of the 2-D class): dense fit on scattered 2-D points, the grid of
    its `dims` predicted at construction, predict_locations with the full covariance, surface_points, and the dimension
    check -- against the oracle's create_gpis / gp_mean / gp_cov."""
    from gpis_b200.sdf_io import Gpis2D
    rng = np.random.default_rng(3)
    X = rng.uniform(0.0, 9.0, size=(120, 2))
    y = np.hypot(X[:, 0] - 4.5, X[:, 1] - 4.0) - 3.0 + 0.01 * rng.standard_normal(120)
    ell, noise = 2.0, 0.02
    gp2 = Gpis2D(X, y, (10, 9), meas_variance=noise, kernel_hyps=(1.0, ell))
    hyp = {"cov": np.array([np.log(ell), 0.0]), "mean": np.zeros(3), "lik": 0.5 * np.log(noise)}
    ref = O.create_gpis(X, y, None, None, hyp)
    q = gp2.grid_pts_
    assert q.shape == (90, 2) and gp2.mean_sdf().shape == (10, 9) and gp2.variance().shape == (10, 9)
    m_ref, _, Kx = O.gp_mean(ref, q, False)
    assert np.max(np.abs(gp2.mean_pred_[:, 0] - m_ref)) < 1e-8
    assert np.max(np.abs(gp2.var_pred_[:, 0] - (O.gp_var_diag(ref, q, Kx, False) + noise))) < 1e-8
    m, c = gp2.predict_locations(q[:20], full_cov=True)
    cov_ref = O.gp_cov(ref, q[:20], None, False)
    assert np.max(np.abs(m[:, 0] - m_ref[:20])) < 1e-8
    assert np.max(np.abs(c - (cov_ref + noise * np.eye(20)))) < 1e-8
    sp = gp2.surface_points(0.3)
    assert np.array_equal(sp, q[np.abs(m_ref) < 0.3]) and len(sp) > 0
    with pytest.raises(ValueError):
        Gpis2D(X, y, (10, 9, 2))


def test_batch_of_independent_objects_matches_solo_runs(G):
    """Config 5 in miniature: a batch of independent small objects over a pool of engines/streams."""
    from gpis_b200.batch import select_batch, superellipsoid_tsdf
    g, K, B = 16, 60, 12
    tsdfs = [superellipsoid_tsdf(g, k) for k in range(B)]
    cfg = dict(ell=2.5, noise=0.05, level=0.0, eps=1e-2, beta=10.0)
    res = select_batch(tsdfs, (g, g, g), K, 3254 % g ** 3, n_workers=4, **cfg)
    assert [r["object"] for r in res] == list(range(B))
    for k in (0, 5, 11):
        # the pool's engines run the per-step launches (several handles share the device): the same form
        # solo is bit-identical; the persistent kernel (the solo default) sums the factor append in another order
        eng = G.GpisEngine(3, g ** 3, K, loop_mode=G._lib.LOOP_STEPS, append_mode=G._lib.APPEND_TWO_KERNELS, **cfg)
        eng.set_grid_candidates((g, g, g), (0, 0, 0), (1, 1, 1), tsdfs[k])
        eng.select(3254 % g ** 3, K)
        ids, nm = eng.active()
        mu, var = eng.candidate_posterior()
        assert np.array_equal(ids, res[k]["ids"]) and nm == res[k]["n_in_model"]
        assert np.array_equal(mu, res[k]["mean"]) and np.array_equal(var, res[k]["var"])   # bit-identical
        eng.close()
        eng = G.GpisEngine(3, g ** 3, K, **cfg)
        eng.set_grid_candidates((g, g, g), (0, 0, 0), (1, 1, 1), tsdfs[k])
        eng.select(3254 % g ** 3, K)
        ids, nm = eng.active()
        mu, var = eng.candidate_posterior()
        assert np.array_equal(ids, res[k]["ids"]) and nm == res[k]["n_in_model"]
        assert np.max(np.abs(mu - res[k]["mean"])) < 1e-12 and np.max(np.abs(var - res[k]["var"])) < 1e-12
        eng.close()
    # objects differ from each other
    assert not np.array_equal(res[0]["ids"], res[1]["ids"])


def test_tf32_mode_tcgen05(G):
    """GPIS_PREC_TF32: the update GEMM runs on tcgen05 (3xTF32, TMEM accumulators).  north_star's bar
    for this mode is 1e-3 relative on the posterior; the picks need not equal the fp64 sequence, so
    the posterior is checked against the oracle's dense GP on the points THIS run selected."""
    shape, hyp = sphere(20)
    tr = {"activeSetSize": 160, "beta": 10.0, "eps": 1e-2, "levelSet": 0.0}
    m64, _, _ = G.select_active_set(shape, tr, first_index=4321, hyp=hyp, backend="grid", predict=False)
    m32, _, _ = G.select_active_set(shape, tr, first_index=4321, hyp=hyp, backend="grid", predict=False,
                                    precision=G._lib.PREC_TF32)
    ids = m32["order"][:m32["n_in_model"]]
    assert len(set(ids.tolist())) == len(ids) and len(m32["order"]) == len(m64["order"])
    same = int(np.argmin(np.r_[m32["order"] == m64["order"], False]))
    assert same >= 20, f"TF32 picks diverge from fp64 already at step {same}"
    ref = O.create_gpis(shape["points"][ids], shape["tsdf"][ids], None, None, hyp)
    fm, _, Kx = O.gp_mean(ref, shape["points"], False)
    fv = O.gp_var_diag(ref, shape["points"], Kx, False)
    mu, var = m32.engine.candidate_posterior()
    assert np.max(np.abs(mu - fm)) < 1e-3 * max(1.0, np.max(np.abs(fm)))
    assert np.max(np.abs(var - fv) / np.maximum(np.abs(fv), 1e-2)) < 1e-3
    print("tf32: first divergence from fp64 picks at", same, "of", len(m64["order"]),
          "max |dmu|", np.max(np.abs(mu - fm)), "max rel dvar", np.max(np.abs(var - fv) / np.maximum(np.abs(fv), 1e-2)))


def test_tf32_mode_gradients_2d_and_odd_sizes(G):
    """TF32 mode beyond the headline shape: gradient rows (NB = 4 passes), 2-D lattices, axes that
    are not multiples of the tile (per-row bulk copies instead of one 16 KB copy), tiny models."""
    shape, hyp = sphere(12, ell=2.5, with_normals=True)
    tr = {"activeSetSize": 60, "beta": 10.0, "eps": 1e-2, "levelSet": 0.0}
    m, _, _ = G.select_active_set(shape, tr, first_index=500, hyp=hyp, backend="grid", predict=False,
                                  precision=G._lib.PREC_TF32)
    ids = m["order"][:m["n_in_model"]]
    ref = O.create_gpis(shape["points"][ids], shape["tsdf"][ids], shape["normals"][ids], None, hyp)
    fm, _, Kx = O.gp_mean(ref, shape["points"], True)
    fv = O.gp_var_diag(ref, shape["points"], Kx, True)
    mu, var = m.engine.candidate_posterior()
    n = shape["points"].shape[0]
    assert np.max(np.abs(mu - fm[:n])) < 1e-3
    assert np.max(np.abs(var - fv) / np.maximum(np.abs(fv), 1e-2)) < 1e-3
    rng = np.random.default_rng(7)
    for dims in [(25, 25), (130, 3), (7, 5)]:
        nx, ny = dims
        ii = np.indices((ny, nx)).reshape(2, -1)[::-1].T.astype(np.float64) + 1.0
        c = np.array([nx * 0.52, ny * 0.47])
        tsdf = np.clip((np.linalg.norm(ii - c, axis=1) - min(dims) * 0.3) / 2.0, -1, 1) + 1e-3 * rng.standard_normal(ii.shape[0])
        sh = {"points": ii, "tsdf": tsdf, "normals": None, "noise": None}
        hy = {"cov": np.array([0.6, 0.0]), "mean": np.zeros(3), "lik": 0.5 * np.log(0.05)}
        t2 = {"activeSetSize": min(30, ii.shape[0]), "beta": 10.0, "eps": 1e-2, "levelSet": 0.0}
        m2, _, _ = G.select_active_set(sh, t2, first_index=3, hyp=hy, backend="grid", predict=False,
                                       precision=G._lib.PREC_TF32)
        i2 = m2["order"][:m2["n_in_model"]]
        r2 = O.create_gpis(ii[i2], tsdf[i2], None, None, hy)
        f2, _, K2 = O.gp_mean(r2, ii, False)
        v2 = O.gp_var_diag(r2, ii, K2, False)
        mu2, var2 = m2.engine.candidate_posterior()
        assert np.max(np.abs(mu2 - f2)) < 1e-3, dims
        assert np.max(np.abs(var2 - v2) / np.maximum(np.abs(v2), 1e-2)) < 1e-3, dims
    # the panel backend does not implement this mode and says so
    with pytest.raises(G._lib.GpisError):
        G.select_active_set(shape, tr, first_index=500, hyp=hyp, backend="panel", precision=G._lib.PREC_TF32)


def test_one_dimensional_lattice_both_precisions(G):
    rng = np.random.default_rng(9)
    n = 300
    x = np.arange(n, dtype=np.float64)[:, None] * 0.5
    tsdf = np.clip(np.sin(x[:, 0] / 7.0) * 1.5, -1, 1) + 1e-3 * rng.standard_normal(n)
    shape = {"points": x, "tsdf": tsdf, "normals": None, "noise": None}
    hyp = {"cov": np.array([np.log(3.0), 0.0]), "mean": np.zeros(2), "lik": 0.5 * np.log(0.05)}
    tr = {"activeSetSize": 40, "beta": 10.0, "eps": 1e-2, "levelSet": 0.0}
    ora = O.select_straddle_incremental(shape, tr, 17, hyp=hyp, prune=False)
    m, _, _ = G.select_active_set(shape, tr, first_index=17, hyp=hyp, backend="grid", predict=False)
    assert m["backend"] == "grid" and np.array_equal(m["order"], ora["order"])
    fm, fv, _ = O.predict_points(ora["gp"], x)
    mu, var = m.engine.candidate_posterior()
    assert np.max(np.abs(mu - fm)) < 1e-8 and np.max(np.abs(var - fv)) < 1e-8
    mt, _, _ = G.select_active_set(shape, tr, first_index=17, hyp=hyp, backend="grid", predict=False,
                                   precision=G._lib.PREC_TF32)
    it = mt["order"][:mt["n_in_model"]]
    ref = O.create_gpis(x[it], tsdf[it], None, None, hyp)
    f2, _, K2 = O.gp_mean(ref, x, False)
    mu2, var2 = mt.engine.candidate_posterior()
    assert np.max(np.abs(mu2 - f2)) < 1e-3
    assert np.max(np.abs(var2 - O.gp_var_diag(ref, x, K2, False))) < 1e-3
