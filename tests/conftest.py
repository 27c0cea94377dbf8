import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_sessionstart(session):
    """Make sure the C-ABI library exists and is current (nvcc cross-compiles without a GPU)."""
    try:
        from gpis_b200 import build
        build.build()
    except Exception as e:  # the tests that need it will fail loudly
        print("gpis_b200 build failed:", e)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def golden_model(g):
    """hyp / inputs of a stored gpModel in oracle form."""
    hyp = {"cov": g["model_hyp_cov"], "mean": g["model_hyp_mean"], "lik": float(g["model_hyp_lik"])}
    m = g["model_training_x"].shape[0]
    D = g["model_training_x"].shape[1]
    y = g["model_training_y"]
    return hyp, g["model_training_x"], y[:m], y[m:].reshape(m, D, order="F"), g["model_beta"]


@pytest.fixture(scope="session")
def golden():
    return load_golden
