"""gpis_build_tsdf (matlab/core/auto_tsdf.m:133-262 on the device) against the oracle restatement, bit for bit,
on the reference's stored shapes and on random images with every branch of the noise model."""
import os

import numpy as np
import pytest

from oracle import tsdf_oracle as T
from test_oracle_tsdf import REF_VP

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tsdf_shapes.npz")


@pytest.fixture(scope="module")
def G():
    import torch
    assert torch.cuda.is_available()
    import gpis_b200
    return gpis_b200


def same(sp, ref):
    for k in ("tsdf", "noise", "normals", "points", "all_points", "fullTsdf", "fullNormals", "fullNoise", "validIndices"):
        assert np.array_equal(sp[k], ref[k]), k
    assert np.allclose(sp["com"], ref["com"], atol=1e-12, equal_nan=True)


@pytest.mark.parametrize("name", ["tape", "bowl", "deodorant", "marker", "oatmeal", "stapler"])
def test_stored_shapes_levels_and_normals_bit_exact(G, name):
    z = np.load(GOLDEN)
    inside = z[name + "_inside"]
    rng = np.random.default_rng(3)
    u = rng.uniform(size=(3,) + inside.shape)
    sp = G.auto_tsdf(inside, REF_VP, u=u)
    assert np.array_equal(sp["fullTsdf"], z[name + "_fullTsdf"])           # the reference's stored values
    assert np.array_equal(sp["fullNormals"], z[name + "_fullNormals"])
    same(sp, T.auto_tsdf(inside, REF_VP, u))


def test_random_images_every_noise_branch(G):
    rng = np.random.default_rng(11)
    for case, (rows, cols) in enumerate([(1, 1), (1, 7), (9, 1), (2, 2), (31, 17), (64, 64), (200, 131)]):
        inside = rng.uniform(size=(rows, cols)) < 0.55
        if rows > 8 and cols > 8:                                            # some solid blobs so that -1 / -0.5 levels exist
            inside[rows // 4:rows // 4 + 6, cols // 4:cols // 4 + 7] = True
            inside[rows // 2:rows // 2 + 5, cols // 2:cols // 2 + 5] = False
        u = rng.uniform(size=(3, rows, cols))
        for mode in ("None", "TLBR", "TRBL", "BLTR", "BRTL"):
            vp = dict(REF_VP, noiseGradMode=mode, specularNoise=bool(case % 2), sparsityRate=0.3, edgeWin=case % 3,
                      interiorRate=0.5, horizScale=0.7, vertScale=1.3, y_thresh1_low=1.5, y_thresh1_high=rows / 2,
                      x_thresh1_low=0, x_thresh1_high=cols / 3, y_thresh2_low=rows - 3, y_thresh2_high=rows,
                      x_thresh2_low=cols - 4, x_thresh2_high=cols)
            same(G.auto_tsdf(inside, vp, u=u), T.auto_tsdf(inside, vp, u))
    # no draws given: 0.5 everywhere
    inside = rng.uniform(size=(20, 20)) < 0.5
    same(G.auto_tsdf(inside, REF_VP), T.auto_tsdf(inside, REF_VP, np.full((3, 20, 20), 0.5)))


def test_tsdf_feeds_the_selection_path(G):
    """create_experiment_object.m's order of things: auto_tsdf -> select_active_set on the measured points."""
    z = np.load(GOLDEN)
    sp = G.auto_tsdf(z["marker_inside"], REF_VP, seed=5)
    tr = {"activeSetSize": 60, "beta": 10.0, "eps": 1e-2, "levelSet": 0.0}
    hyp = {"cov": np.array([0.6, 0.0]), "mean": np.zeros(3), "lik": float(np.log(0.1))}
    shape = {"points": sp["points"], "tsdf": sp["tsdf"], "normals": sp["normals"], "noise": sp["noise"]}
    from oracle import gpis_oracle as O
    ora = O.select_straddle_incremental(shape, tr, 17, hyp=hyp)
    mdl, _, _ = G.select_active_set(shape, tr, first_index=17, hyp=hyp, predict=False)
    assert np.array_equal(mdl["order"], ora["order"])
