"""The reference's published 2-D selection benchmark (scripts/bench_published.py):
the picks of both GPU backends against the LITERAL restatement of select_active_set_straddle.m (refit
from scratch every step) at the published sizes the literal loop finishes in seconds."""
import importlib.util
import json
import os
import zlib

import numpy as np
import pytest

from conftest import ROOT
from oracle import gpis_oracle as O

spec = importlib.util.spec_from_file_location("bench_published", os.path.join(ROOT, "scripts", "bench_published.py"))
BP = importlib.util.module_from_spec(spec)
spec.loader.exec_module(BP)

HYP = {"cov": np.array([0.5 * np.log(BP.SIGMA), 0.0]), "mean": np.zeros(3), "lik": 0.5 * np.log(BP.NOISE)}
TRAIN = {"activeSetSize": BP.M, "beta": 10.0, "eps": 1e-2, "levelSet": 0.0}


def literal_picks(g):
    pts, tsdf = BP.circle_tsdf(g)
    shape = {"points": pts, "tsdf": tsdf, "normals": None, "noise": None}
    return shape, O.select_active_set_straddle(shape, TRAIN, 766020790 % (g * g), hyp=HYP)


@pytest.mark.parametrize("g", [25, 50])
def test_recorded_gpu_picks_equal_the_literal_oracle(g):
    """profiles/r01_published_benchmark_gpu.json was written on a B200; its pick checksums must be the
    literal CPU loop's (no GPU needed to check that)."""
    rows = {json.loads(l)["grid"]: json.loads(l)
            for l in open(os.path.join(ROOT, "profiles", "r01_published_benchmark_gpu.json"))}
    _, ora = literal_picks(g)
    crc = zlib.crc32(np.asarray(ora["order"], np.int64).tobytes())
    assert rows[g]["grid_ids_crc32"] == crc and rows[g]["panel_ids_crc32"] == crc
    assert rows[g]["grid_n_selected"] == len(ora["order"])


@pytest.mark.gpu
@pytest.mark.parametrize("g", [25, 50, 100])
@pytest.mark.parametrize("backend", ["grid", "panel"])
def test_gpu_picks_and_posterior_equal_the_literal_oracle(g, backend):
    import gpis_b200
    shape, ora = literal_picks(g)
    mdl, pred, test = gpis_b200.select_active_set(shape, TRAIN, first_index=766020790 % (g * g), hyp=HYP, backend=backend)
    assert np.array_equal(mdl["order"], ora["order"])
    assert np.array_equal(test, ora["testIndices"])
    assert np.max(np.abs(pred["tsdf"] - ora["pred_tsdf"])) < 1e-8
    assert np.max(np.abs(pred["noise"] - ora["pred_noise"])) < 1e-8
