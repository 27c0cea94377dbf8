"""Host-side mirror of the reference's C++ operator, `GpuActiveSetSelector`
(include/gpu_active_set_selector.hpp:32-47, src/gpu_active_set_selector.cpp) and of the `GPIS`
command-line driver's 8-field config (src/main.cpp:25-67).

Same method names, argument order and meaning.  Behaviour follows the C++ path where it is
well-defined and the CPU semantics where the C++ code is defective:
  * kernel exp(-|x-z|^2 / (2 sigma)) with sigma as a FLOAT (the reference truncates it to int);
  * variance k(x,x) + beta - |L^-1 k|^2 (max_subset_buffers.cu:113-115; the reference's TRSM uses
    the wrong transpose), score sqrt(beta_t) * var - |mu - level| (:126), sign classification;
  * beta_t = 2 log(N pi^2 (k+1)^2 / (6 tol)) (gpu_active_set_selector.cpp:493,550), level 0;
  * first index = glibc srand(1000); rand() % N  (main.cpp:77, :470) unless given;
  * activeInputs / activeTargets ARE filled (the reference never writes them), and the three CSV
    files it drops in the working directory are written only when `out_dir` is given.
"""
import os

import numpy as np

from . import _lib
from .engine import GpisEngine
from .synth import first_index as _glibc_first

ENTROPY, LEVEL_SET = 0, 1


class GaussianProcessHyperparams:
    """include/active_set_selection_types.h:10-13."""

    def __init__(self, beta=0.1, sigma=1.0):
        self.beta, self.sigma = float(beta), float(sigma)


class PredictionError:
    """include/gpu_active_set_selector.hpp:11-17; `std` is the second moment, as in :161."""

    def __init__(self, abs_err):
        a = np.asarray(abs_err, np.float64)
        n = a.size
        self.mean = float(a.mean()) if n else 0.0
        self.std = float((a * a).mean()) if n else 0.0
        self.median = float(np.median(a)) if n else 0.0
        self.min = float(a.min()) if n else 0.0
        self.max = float(a.max()) if n else 0.0


def read_config(path):
    """main.cpp:25-67: eight whitespace-separated fields, one per non-comment line
    (csv, setSize, sigma, beta, width, height, depth, batch).  The reference's switch has no
    breaks, so each line also overwrites the later fields; the last assignment wins, which is
    what a well-formed file gives here directly."""
    vals = []
    with open(path) as f:
        for line in f:
            if line.startswith("#") or not line.strip():
                continue
            vals.append(line.split()[0])
            if len(vals) == 8:
                break
    if len(vals) < 8:
        raise ValueError("Illegal configuration - too few params")
    return {"csv": vals[0], "setSize": int(vals[1]), "sigma": float(vals[2]), "beta": float(vals[3]),
            "width": int(vals[4]), "height": int(vals[5]), "depth": int(vals[6]), "batch": int(vals[7])}


def read_csv(path, width, height, depth, store_depth):
    """ReadCsv (gpu_active_set_selector.cpp:75-113): `depth` slabs of `height` rows of `width`
    values; inputs are the integer lattice (x = column, y = row, z = slab), SoA."""
    vals = np.loadtxt(path, delimiter=",", ndmin=2)[:height * depth, :width]
    targets = np.ascontiguousarray(vals.reshape(-1))                  # IJK_TO_LINEAR: x fastest
    n = width * height * depth
    idx = np.arange(n)
    cols = [idx % width, (idx // width) % height]
    if store_depth:
        cols.append(idx // (width * height))
    return np.ascontiguousarray(np.stack(cols).astype(np.float64)), targets


def write_csv(path, buf, width, height):
    """WriteCsv (:54-72): `height` lines of `width` values from a column-major buffer."""
    b = np.asarray(buf, np.float64).reshape(width, height)
    np.savetxt(path, b.T, delimiter=",", fmt="%.9g")


class GpuActiveSetSelector:
    def __init__(self):
        self.last = None

    def SelectFromGrid(self, csvFilename, setSize, sigma, beta, width, height, depth, batchSize, tolerance,
                       storeDepth=False, first_index=None, out_dir=None):
        inputDim = 3 if storeDepth else 2
        n = width * height * depth
        inputs, targets = read_csv(csvFilename, width, height, depth, storeDepth)
        hypers = GaussianProcessHyperparams(beta, sigma)
        activeInputs = np.zeros(n * inputDim, np.float32)
        activeTargets = np.zeros(n, np.float32)
        lattice = (width, height, depth) if storeDepth else ((width, height) if depth == 1 else None)
        return self.SelectChol(setSize, inputs.reshape(-1), targets, LEVEL_SET, hypers, inputDim, 1, n, tolerance,
                               batchSize, activeInputs, activeTargets, first_index=first_index, out_dir=out_dir,
                               _lattice=lattice)

    def SelectChol(self, maxSize, inputPoints, targetPoints, mode, hypers, inputDim, targetDim, numPoints, tolerance,
                   batchSize, activeInputs=None, activeTargets=None, first_index=None, out_dir=None, _lattice=None):
        if mode != LEVEL_SET:
            raise NotImplementedError("only LEVEL_SET is implemented (the reference hard-wires it, :191)")
        if targetDim != 1:
            raise ValueError("targetDim must be 1")
        pts = np.asarray(inputPoints, np.float64).reshape(inputDim, numPoints)            # SoA, :96-101
        y = np.asarray(targetPoints, np.float64).reshape(-1)[:numPoints]
        first = _glibc_first(numPoints) if first_index is None else int(first_index)
        eng = GpisEngine(inputDim, numPoints, maxSize, ell=float(np.sqrt(hypers.sigma)), sf2=1.0, noise=hypers.beta,
                         level=0.0, eps=0.0, beta_mode=_lib.BETA_SELECTCHOL, beta_delta=float(tolerance),
                         variance_mode=_lib.VAR_PREDICTIVE)
        if _lattice is not None:
            eng.set_grid_candidates(_lattice, (0.0,) * inputDim, (1.0,) * inputDim, y)
        else:
            eng.set_candidates(pts, y, soa=True)
        eng.select(first, maxSize)
        ids, n_model = eng.active()
        k = len(ids)
        if activeInputs is not None:                   # active_inputs[a + d*max_active] (active_set_buffers.cu:200)
            ai = np.asarray(activeInputs).reshape(-1)
            for d in range(inputDim):
                ai[d * maxSize:d * maxSize + k] = pts[d, ids]
        if activeTargets is not None:
            np.asarray(activeTargets).reshape(-1)[:k] = y[ids]
        _, alpha = eng.factor()
        # final prediction over all points + error statistics on the non-active ones (:568-581)
        if _lattice is not None:
            mu, var = eng.candidate_posterior()
        else:
            mu, var, _ = eng.predict(pts.T)
        active = np.zeros(numPoints, bool)
        active[ids] = True
        errors = PredictionError(np.abs(mu[~active] - y[~active]))
        if out_dir is not None:                        # inputs.csv / targets.csv / alpha.csv (:584-586)
            os.makedirs(out_dir, exist_ok=True)
            write_csv(os.path.join(out_dir, "inputs.csv"), pts[:, ids[:n_model]].T.reshape(-1, order="F"), inputDim, n_model)
            write_csv(os.path.join(out_dir, "targets.csv"), y[ids[:n_model]], 1, n_model)
            write_csv(os.path.join(out_dir, "alpha.csv"), alpha, 1, n_model)
        self.last = {"ids": ids, "n_in_model": n_model, "alpha": alpha, "mu": mu, "sigma": var, "errors": errors,
                     "engine": eng, "select_ms": eng.timing()["select_ms"]}
        return True
