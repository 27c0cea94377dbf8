"""The TSDF-builder restatement (oracle/tsdf_oracle.py, matlab/core/auto_tsdf.m:133-262) against the
reference's stored shapeParams (tests/golden/tsdf_shapes.npz from data/google_objects/*.mat)."""
import os

import numpy as np
import pytest

from oracle import tsdf_oracle as T

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tsdf_shapes.npz")
# matlab/grasp_metric/setup_and_run_gpis_experiment.m:24-57 (dim = 25)
REF_VP = dict(y_thresh1_low=12.5, y_thresh2_low=25, y_thresh3_low=25, y_thresh1_high=30, y_thresh2_high=0, y_thresh3_high=0,
              x_thresh1_low=15, x_thresh2_low=25, x_thresh3_low=25, x_thresh1_high=25, x_thresh2_high=0, x_thresh3_high=0,
              occlusionScale=1000, noiseScale=0.05, interiorRate=0.4, specularNoise=False, sparsityRate=0.0, edgeWin=0,
              noiseGradMode="None", horizScale=1, vertScale=1)


def flat(a):
    return np.asarray(a).reshape(-1, order="F")


@pytest.mark.parametrize("name", ["tape", "bowl", "deodorant", "marker", "oatmeal", "stapler"])
def test_levels_and_normals_reproduce_the_stored_shapes_exactly(name):
    z = np.load(GOLDEN)
    inside = z[name + "_inside"]
    t = T.tsdf_levels(inside)
    assert np.array_equal(flat(t), z[name + "_fullTsdf"])                       # five levels, bit-exact
    gx, gy = T.central_gradients(t)
    assert np.array_equal(np.stack([flat(gx), flat(gy)], axis=1), z[name + "_fullNormals"])
    # the measured subset is a selection of the full grid in column-major order, with the stored noise scale
    pts = z[name + "_points"]
    lin = ((pts[:, 0] - 1) * inside.shape[0] + (pts[:, 1] - 1)).astype(int)
    assert np.all(np.diff(lin) > 0)
    assert np.array_equal(z[name + "_tsdf"], z[name + "_fullTsdf"][lin])
    assert np.array_equal(z[name + "_normals"], z[name + "_fullNormals"][lin])
    assert np.all(z[name + "_noise"] == 0.05)


def test_noise_model_deterministic_part_on_the_marker_shape():
    """The marker shape was stored with the parameters of setup_and_run_gpis_experiment.m: every point the
    model must measure (not occluded, tsdf >= -0.5) is there, nothing the model must occlude is, and the
    interior points (drawn with rand() > 1 - interiorRate) are a subset of the interior."""
    z = np.load(GOLDEN)
    inside = z["marker_inside"]
    g = inside.shape[0]
    t = T.tsdf_levels(inside)
    pts = z["marker_points"]
    valid = np.zeros((g, g), bool)
    valid[(pts[:, 1] - 1).astype(int), (pts[:, 0] - 1).astype(int)] = True
    must = T.noise_model(t, REF_VP, np.zeros((3, g, g))) < 1000
    may = T.noise_model(t, REF_VP, np.ones((3, g, g))) < 1000
    assert np.all(valid[must]) and np.all(may[valid])
    frac = (valid & ~must).sum() / (may & ~must).sum()
    assert 0.2 < frac < 0.6                                    # interiorRate 0.4
    # end to end with explicit draws that reproduce the stored subset
    u = np.zeros((3, g, g))
    u[0][valid & ~must] = 1.0
    sp = T.auto_tsdf(inside, REF_VP, u)
    assert np.array_equal(sp["points"], pts) and np.array_equal(sp["tsdf"], z["marker_tsdf"])
    assert np.array_equal(sp["normals"], z["marker_normals"]) and np.array_equal(sp["noise"], z["marker_noise"])
    assert np.allclose(sp["com"], z["marker_com"], atol=1e-12)


def test_noise_gradient_modes_and_specular_branch():
    rng = np.random.default_rng(0)
    inside = np.zeros((12, 9), bool)
    inside[3:9, 2:7] = True
    t = T.tsdf_levels(inside)
    u = rng.uniform(size=(3, 12, 9))
    vp = dict(REF_VP, y_thresh1_low=0, y_thresh1_high=4, x_thresh1_low=0, x_thresh1_high=3, specularNoise=True, sparsityRate=0.3,
              edgeWin=1, horizScale=0.5, vertScale=2.0)
    base = T.noise_model(t, dict(vp, noiseGradMode="None"), u)
    for mode in ("TLBR", "TRBL", "BLTR", "BRTL"):
        n = T.noise_model(t, dict(vp, noiseGradMode=mode), u)
        i, j = 6, 8                                               # an exterior pixel away from the window of the shape
        w = {"TLBR": 0.5 * (j + 1) + 2.0 * (i + 1), "TRBL": 0.5 * (9 - j) + 2.0 * (i + 1),
             "BLTR": 0.5 * (j + 1) + 2.0 * (12 - i), "BRTL": 0.5 * (9 - j) + 2.0 * (12 - i)}[mode]
        assert np.isclose(n[i, j], base[i, j] * w)
    assert base[3, 2] == 1000 and base[1, 1] != 1000 and base[5, 4] in (0.05, 1000.0)   # occluded only where tsdf < 0.6; deep interior
