"""gpis_batch_select (one thread block runs the whole selection of one object; BASELINE config 5)
against solo engines and the oracle.  Named to run after the other GPU tests."""
import numpy as np
import pytest

from oracle import gpis_oracle as O

CFG = dict(ell=2.5, noise=0.05, level=0.0, eps=1e-2, beta=10.0)
HYP = {"cov": np.array([np.log(2.5), 0.0]), "mean": np.zeros(4), "lik": 0.5 * np.log(0.05)}


def lattice_points(dims):
    idx = np.indices(tuple(dims)[::-1]).reshape(len(dims), -1)[::-1]
    return np.ascontiguousarray(idx.T.astype(np.float64))


def test_arguments_are_validated_before_the_device_is_touched():
    """Runs on the CPU box too: bad shapes are GPIS_ERR_INVALID; a valid call without a GPU is
    GPIS_ERR_NO_DEVICE (there is no CPU path)."""
    import torch
    from gpis_b200 import _lib
    from gpis_b200.batch import select_batch_fused
    for dims, K in [((33, 4, 4), 4), ((4, 4, 4), 257), ((4, 4, 4), 0)]:
        with pytest.raises(_lib.GpisError) as ei:
            select_batch_fused([np.zeros(int(np.prod(dims)))], dims, K, 0, **CFG)
        assert ei.value.status == _lib.ERR_INVALID
    if not torch.cuda.is_available():
        with pytest.raises(_lib.GpisError) as ei:
            select_batch_fused([np.zeros(64)], (4, 4, 4), 4, 0, **CFG)
        assert ei.value.status == _lib.ERR_NO_DEVICE


@pytest.fixture(params=["plain", "tuned", "dmma"])
def variant(request):
    """The forms of the kernel's inner loops (gpis_config.batch_kernel)."""
    from gpis_b200 import _lib
    return {"plain": _lib.BATCH_PLAIN, "tuned": _lib.BATCH_DFMA, "dmma": _lib.BATCH_DMMA}[request.param]


@pytest.mark.gpu
def test_fused_batch_matches_solo_engines_and_oracle(variant):
    import gpis_b200
    from gpis_b200.batch import select_batch_fused, superellipsoid_tsdf
    g, K, B = 16, 60, 7
    first = 3254 % g ** 3
    tsdfs = [superellipsoid_tsdf(g, k) for k in range(B)]
    res = select_batch_fused(tsdfs, (g, g, g), K, first, kernel=variant, **CFG)
    assert [r["object"] for r in res] == list(range(B)) and res[0]["device_ms"] > 0
    for k in (0, 3, 6):
        eng = gpis_b200.GpisEngine(3, g ** 3, K, **CFG)
        eng.set_grid_candidates((g, g, g), (0, 0, 0), (1, 1, 1), tsdfs[k])
        eng.select(first, K)
        ids, nm = eng.active()
        mu, var = eng.candidate_posterior()
        amb, hi, lo, ac = eng.levelset()
        assert np.array_equal(ids, res[k]["ids"]) and nm == res[k]["n_in_model"], k
        assert np.max(np.abs(mu - res[k]["mean"])) < 1e-10 and np.max(np.abs(var - res[k]["var"])) < 1e-10
        f = res[k]["flags"]
        assert np.array_equal((f & 1) > 0, ac) and np.array_equal((f & 2) > 0, hi) and np.array_equal((f & 4) > 0, lo)
        eng.close()
    pts = lattice_points((g, g, g))
    ora = O.select_straddle_incremental({"points": pts, "tsdf": tsdfs[1], "normals": None, "noise": None},
                                        {"activeSetSize": K, "beta": 10.0, "eps": 1e-2, "levelSet": 0.0}, first, hyp=HYP,
                                        prune=False)
    assert np.array_equal(res[1]["ids"], ora["order"]) and res[1]["n_in_model"] == ora["n_in_model"]
    fm, fv, _ = O.predict_points(ora["gp"], pts)
    assert np.max(np.abs(res[1]["mean"] - fm)) < 1e-9 and np.max(np.abs(res[1]["var"] - fv)) < 1e-9
    assert not np.array_equal(res[0]["ids"], res[1]["ids"])


@pytest.mark.gpu
def test_fused_batch_more_objects_than_sms_2d_and_device_input(variant):
    """200 small objects (> 148 blocks: blocks take several objects in turn, reusing their workspace),
    a 2-D lattice with odd sizes, and a torch CUDA tensor as input."""
    import torch
    from gpis_b200.batch import select_batch_fused
    rng = np.random.default_rng(3)
    dims, K = (13, 11), 25
    pts = lattice_points(dims)
    tsdfs = np.stack([np.clip((np.linalg.norm(pts - np.array(dims) * rng.uniform(0.35, 0.65, 2), axis=1)
                               - rng.uniform(2.5, 4.5)) / 2.0, -1, 1) + 1e-3 * rng.standard_normal(len(pts))
                      for _ in range(200)])
    res = select_batch_fused(torch.as_tensor(tsdfs, device="cuda"), dims, K, 5, kernel=variant, **CFG)
    assert len(res) == 200
    hyp = {"cov": HYP["cov"], "mean": np.zeros(3), "lik": HYP["lik"]}
    for k in (0, 147, 148, 199):
        ora = O.select_straddle_incremental({"points": pts, "tsdf": tsdfs[k], "normals": None, "noise": None},
                                            {"activeSetSize": K, "beta": 10.0, "eps": 1e-2, "levelSet": 0.0}, 5, hyp=hyp,
                                            prune=False)
        assert np.array_equal(res[k]["ids"], ora["order"]) and res[k]["n_in_model"] == ora["n_in_model"], k
        fm, fv, _ = O.predict_points(ora["gp"], pts)
        assert np.max(np.abs(res[k]["mean"] - fm)) < 1e-9 and np.max(np.abs(res[k]["var"] - fv)) < 1e-9
    # sharding the OBJECTS across ranks: rank r of 3 gets objects r::3
    part = select_batch_fused(tsdfs, dims, K, 5, rank=1, world=3, keep_posterior=False, kernel=variant, **CFG)
    assert [r["object"] for r in part] == list(range(1, 200, 3))
    assert all(np.array_equal(r["ids"], res[r["object"]]["ids"]) for r in part)
