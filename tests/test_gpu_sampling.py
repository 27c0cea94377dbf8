"""GPU parity of the full posterior covariance and posterior shape sampling (SURVEY 8f rank 3:
gp_cov.m in full, mc_sampling/sample_shapes.m, Gpis3D.sample_sdfs) through the C ABI, against the
oracle's restatement.  The standard-normal draws are inputs on both sides, so the comparison is
deterministic; tolerances on draws from nearly singular covariances are wider than the fp64 bar
because two correct Cholesky factorisations of a matrix with eigenvalues at the jitter level differ
by far more than an ulp (the property T'T = COV + jitter I is checked to 1e-11 instead)."""
import numpy as np
import pytest

from conftest import golden_model, load_golden
from oracle import gpis_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import gpis_b200
    return gpis_b200


@pytest.fixture(scope="module")
def loofa(G):
    g = load_golden("loofa")
    hyp, X, y, dy, beta = golden_model(g)
    return G.create_gpis(X, y, dy, beta, False, hyp), O.create_gpis(X, y, dy, beta, hyp)


def stack(shapes):
    return np.stack([np.concatenate([s["tsdf"], s["normals"].reshape(-1, order="F")]) for s in shapes])


@pytest.mark.parametrize("res,tol", [(3, 1e-7), (2, 1e-5)])
def test_sample_shapes_matches_oracle(G, loofa, res, tol):
    mdl, ora = loofa
    pts = O.sample_grid_points(25, res)
    z = np.random.default_rng(3).standard_normal((5, 3 * len(pts)))
    ref, lp = O.sample_shapes(ora, 25, z, resolution=res, jitter=1e-12)
    shapes, pdfs, logpdf = G.sample_shapes(mdl, 25, 5, resolution=res, z=z, jitter=1e-12, return_logpdf=True)
    assert len(shapes) == 5 and shapes[0]["tsdf"].shape == (len(pts),) and shapes[0]["normals"].shape == (len(pts), 2)
    assert np.max(np.abs(stack(shapes) - stack(ref))) < tol
    assert np.max(np.abs(logpdf - lp)) < 1e-5 * np.max(np.abs(lp))
    with np.errstate(over="ignore"):
        assert np.array_equal(np.isinf(pdfs), np.isinf(np.exp(lp)))
    # full gp_cov.m and gp_mean.m
    MEAN, _, Kxxp = O.gp_mean(ora, pts, True)
    COV = O.gp_cov(ora, pts, Kxxp, True)
    cov = G.gp_cov(mdl, pts, None, True, full=True)
    assert np.max(np.abs(cov - COV)) < 1e-11 and np.array_equal(cov, cov.T)
    assert np.max(np.abs(G.gp_mean(mdl, pts, True) - MEAN)) < 1e-10
    assert np.max(np.abs(np.diag(cov)[:len(pts)] - G.gp_cov(mdl, pts))) < 1e-11


def test_sample_shapes_reference_configuration(G, loofa):
    """create_experiment_object.m:23-26 as stored in loofa_construction.mat: the full 25 x 25 grid
    (1875 values), jitter 1e-12; every stored pdf is +inf.  1201 of the 1875 eigenvalues of COV lie
    below the jitter, so a non-positive pivot is a legitimate outcome (mvnrnd then falls back to an
    eigendecomposition; this engine reports GPIS_ERR_NOT_PD): accept it and move to 1e-10."""
    mdl, ora = loofa
    pts = O.sample_grid_points(25, 1)
    z = np.random.default_rng(11).standard_normal((4, 1875))
    MEAN, _, Kxxp = O.gp_mean(ora, pts, True)
    COV = O.gp_cov(ora, pts, Kxxp, True)
    jitter = 1e-12
    try:
        samples, logpdf, T = mdl.engine.sample_shapes(pts, z, use_derivative=True, jitter=jitter, want_chol=True)
    except G._lib.GpisError as e:
        assert e.status == G._lib.ERR_NOT_PD and "raise jitter" in str(e)
        jitter = 1e-10
        samples, logpdf, T = mdl.engine.sample_shapes(pts, z, use_derivative=True, jitter=jitter, want_chol=True)
    assert np.all(np.tril(T, -1) == 0.0) and np.all(np.diag(T) > 0.0)
    assert np.max(np.abs(T.T @ T - COV - jitter * np.eye(1875))) < 1e-11
    assert np.max(np.abs(samples - (z @ T + MEAN))) < 1e-10
    ref, lp, _ = O.sample_from_posterior(MEAN, COV, z, jitter)
    assert np.max(np.abs(samples - ref)) < 1e-3
    assert np.max(np.abs(logpdf - lp)) < 1e-3 * np.max(np.abs(lp))
    with np.errstate(over="ignore"):
        assert np.all(np.isinf(np.exp(logpdf)))                    # constructionResults.pdfs: all inf
    t = mdl.engine.sampling_timing()
    assert t["cov_ms"] > 0 and t["chol_ms"] > 0 and t["draw_ms"] > 0


def test_gpis3d_sample_sdfs_and_full_cov(G):
    from gpis_b200.sdf_io import Gpis3D
    rng = np.random.default_rng(5)
    X = rng.uniform(0, 5, (60, 3))
    y = np.linalg.norm(X - 2.5, axis=1) - 1.6
    gp3 = Gpis3D(X, y, (6, 6, 6), meas_variance=0.01, kernel_hyps=(0.8, 1.4))
    hyp = {"cov": np.array([np.log(1.4), 0.5 * np.log(0.8)]), "mean": np.zeros(4), "lik": 0.5 * np.log(0.01)}
    ora = O.create_gpis(X, y, None, None, hyp)
    pts = gp3.grid_pts_
    MEAN, _, Kxxp = O.gp_mean(ora, pts, False)
    COV = O.gp_cov(ora, pts, Kxxp, False)
    z = rng.standard_normal((3, 216))
    ref, _, _ = O.sample_from_posterior(MEAN, COV, z, 1e-8)
    sdfs = gp3.sample_sdfs(3, z=z, jitter=1e-8)
    assert len(sdfs) == 3 and sdfs[0].shape == (6, 6, 6)
    assert np.max(np.abs(np.stack([s.reshape(-1) for s in sdfs]) - ref)) < 1e-7
    m, c = gp3.predict_locations(pts[:50], full_cov=True)
    assert m.shape == (50, 1) and c.shape == (50, 50)
    assert np.max(np.abs(c - COV[:50, :50] - 0.01 * np.eye(50))) < 1e-11
    # draws from the device generator: right shape, finite, mean over many draws near the posterior mean
    many = np.stack([s.reshape(-1) for s in gp3.sample_sdfs(400, seed=1, jitter=1e-8)])
    sd = np.sqrt(np.maximum(np.diag(COV), 0) + 1e-8)
    assert np.all(np.isfinite(many)) and np.max(np.abs(many.mean(axis=0) - MEAN) / (sd / 20.0 + 1e-3)) < 6.0


def test_larger_covariance_properties_and_device_pointers(G):
    """2000 query points of a 3-D model taken from a selection (the engine's own W), device tensors in
    and out: diag(COV) = the predicted variance, MEAN = the predicted mean, T'T = COV + jitter I."""
    import torch
    from gpis_b200.synth import sphere_tsdf
    pts, tsdf = sphere_tsdf(16, trunc=3.0)
    hyp = {"cov": np.array([np.log(2.5), 0.0]), "mean": np.zeros(4), "lik": 0.5 * np.log(0.05)}
    shape = {"points": pts, "tsdf": tsdf, "normals": None, "noise": None}
    tr = {"activeSetSize": 120, "beta": 10.0, "eps": 1e-2, "levelSet": 0.0}
    mdl, _, _ = G.select_active_set(shape, tr, first_index=1234, hyp=hyp, predict=False)
    rng = np.random.default_rng(0)
    q = torch.as_tensor(rng.uniform(2, 13, (2000, 3)), device="cuda")
    mean, cov = mdl.engine.posterior_cov(q, use_derivative=False)
    assert mean.is_cuda and cov.is_cuda
    mu, var, _ = mdl.engine.predict(q)
    assert float((mean - mu).abs().max()) < 1e-10
    assert float((torch.diagonal(cov) - var).abs().max()) < 1e-10
    assert bool(torch.equal(cov, cov.t()))
    gen = torch.Generator(device="cuda")
    gen.manual_seed(4)
    z = torch.randn((2, 2000), dtype=torch.float64, device="cuda", generator=gen)
    samples, logpdf, T = mdl.engine.sample_shapes(q, z, jitter=1e-9, want_chol=True)
    eye = torch.eye(2000, device="cuda", dtype=torch.float64)
    assert float((T.t() @ T - cov - 1e-9 * eye).abs().max()) < 1e-11
    assert float((samples - (z @ T + mean)).abs().max()) < 1e-10
    lp = -0.5 * (z * z).sum(1) - torch.log(torch.diagonal(T)).sum() - 1000.0 * np.log(2 * np.pi)
    assert float((logpdf - lp).abs().max()) < 1e-7 * float(lp.abs().max())


def test_singular_covariance_is_reported_not_fatal(G):
    """Duplicate query points under the prior (an engine with no rows): the third pivot is exactly 0."""
    eng = G.GpisEngine(2, 0, 4, ell=1.0, sf2=1.0, mean_c=0.3, noise=0.01)
    pts = np.array([[0.0, 0.0], [1.0, 0.5], [0.0, 0.0], [2.0, 2.0]])
    with pytest.raises(G._lib.GpisError) as ei:
        eng.sample_shapes(pts, np.zeros((1, 4)), jitter=0.0)
    assert ei.value.status == G._lib.ERR_NOT_PD and "pivot 3" in str(ei.value)
    samples, logpdf, T = eng.sample_shapes(pts, np.zeros((1, 4)), jitter=1e-9, want_chol=True)   # the handle stays usable
    assert np.max(np.abs(samples[0] - 0.3)) < 1e-15
    K = O.cov_se_iso(np.array([0.0, 0.0]), pts)
    assert np.max(np.abs(T.T @ T - K - 1e-9 * np.eye(4))) < 1e-15
    mean, cov = eng.posterior_cov(pts)
    assert np.max(np.abs(cov - K)) < 1e-15 and np.max(np.abs(mean - 0.3)) < 1e-15
    with pytest.raises(G._lib.GpisError):
        eng.sample_shapes(pts, np.zeros((1, 4)), jitter=-1.0)
    eng.close()
