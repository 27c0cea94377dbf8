"""SDF file I/O and the Gpis2D/Gpis3D host classes -- the callers and data formats either side of
the hot path.

`SdfFile` mirrors src/grasp_selection/sdf_file.py:13-77: a `.sdf` text file is
    line 1  nx ny nz          line 2  ox oy oz          line 3  resolution
    then nx*ny*nz values, x fastest (the reader loops k, j, i outermost to innermost, :52-56)
and a `.csv` file is a 2-D grid (np.loadtxt, :71).  x-fastest is also the engine's lattice order
(IJK_TO_LINEAR), so an SDF feeds `gpis_set_grid_candidates` without any reshuffle.

`Gpis3D` / `Gpis2D` mirror the constructor / `predict_locations` / `surface_points` surface of
src/grasp_selection/gpis.py:53-123,214-303.  The reference fits with GPy (RBF kernel, hyper-
parameter optimisation) -- GPy is not on this path and its arithmetic is unpinned; here the given
`kernel_hyps = [variance, lengthscale]` are used as they are (no `optimize()`), and all GP
arithmetic runs in the CUDA engine.
"""
import os

import numpy as np

from .engine import GpisEngine


class Sdf3D:
    def __init__(self, sdf_data, origin=np.zeros(3), resolution=1.0):
        self.data = np.asarray(sdf_data, np.float64)          # data[i][j][k]
        self.origin = np.asarray(origin, np.float64)
        self.resolution = float(resolution)

    @property
    def dimensions(self):
        return self.data.shape

    def flat(self):
        """Values in lattice order, x fastest."""
        return np.ascontiguousarray(self.data.reshape(-1, order="F"))

    def tsdf(self, trunc):
        """Truncated, normalised signed distance in [-1, 1] (trunc in the file's units)."""
        return np.clip(self.flat() / float(trunc), -1.0, 1.0)


class Sdf2D:
    def __init__(self, sdf_data):
        self.data = np.asarray(sdf_data, np.float64)
        self.origin = np.zeros(2)
        self.resolution = 1.0

    @property
    def dimensions(self):
        return self.data.shape


class SdfFile:
    def __init__(self, file_name):
        self.file_name_ = file_name
        ext = os.path.splitext(file_name)[1]
        if ext == ".sdf":
            self.use_3d_ = True
        elif ext == ".csv":
            self.use_3d_ = False
        else:
            raise ValueError("Extension %s invalid for SDFs" % ext)

    def read(self):
        return self._read_3d() if self.use_3d_ else self._read_2d()

    def _read_3d(self):
        with open(self.file_name_) as f:
            nx, ny, nz = [int(t) for t in f.readline().split()]
            origin = np.array([float(t) for t in f.readline().split()])
            resolution = float(f.readline())
            vals = np.loadtxt(f, dtype=np.float64, ndmin=1)
        if vals.size != nx * ny * nz:
            raise IOError("%s: expected %d values, found %d" % (self.file_name_, nx * ny * nz, vals.size))
        return Sdf3D(vals.reshape((nx, ny, nz), order="F"), origin, resolution)

    def _read_2d(self):
        if not os.path.exists(self.file_name_):
            raise IOError("File does not exist")
        return Sdf2D(np.loadtxt(self.file_name_, delimiter=","))

    def write(self, sdf):
        """The reference leaves this as a TODO (sdf_file.py:74-77); same format as read()."""
        if self.use_3d_:
            nx, ny, nz = sdf.dimensions
            with open(self.file_name_, "w") as f:
                f.write("%d %d %d\n" % (nx, ny, nz))
                f.write("%.17g %.17g %.17g\n" % tuple(sdf.origin))
                f.write("%.17g\n" % sdf.resolution)
                np.savetxt(f, sdf.flat(), fmt="%.17g")
        else:
            np.savetxt(self.file_name_, sdf.data, delimiter=",", fmt="%.17g")


class _GpisBase:
    def __init__(self, points, sdf_meas, dims, meas_variance=0.1, kernel_name="rbf", kernel_hyps=(1.0, 3.0),
                 origin=None, resolution=1.0):
        if kernel_name != "rbf":
            raise ValueError("Invalid Kernel specified")
        self.pts_ = np.asarray(points, np.float64)
        self.data_ = np.asarray(sdf_meas, np.float64).reshape(-1)
        self.var_, self.dims_ = float(meas_variance), tuple(int(d) for d in dims)
        self.origin_, self.resolution_ = origin, resolution
        D = self.pts_.shape[1]
        self.engine = GpisEngine(D, 0, max(self.pts_.shape[0], 1), ell=float(kernel_hyps[1]),
                                 sf2=float(kernel_hyps[0]), noise=self.var_)
        self.engine.fit(self.pts_, self.data_)
        idx = np.indices(self.dims_)                                  # gpis.py:95-96 (C order)
        self.grid_pts_ = np.stack([a.reshape(-1) for a in idx], axis=1).astype(np.float64)
        self.mean_pred_, self.var_pred_ = self.predict_locations(self.grid_pts_)

    def predict_locations(self, points, full_cov=False):
        """gpis.py:114-123: (means, vars) at the given points; with full_cov the second output is
        the whole predictive covariance (latent + meas_variance on the diagonal, as GPy returns)."""
        if full_cov:
            m, c = self.engine.posterior_cov(points, use_derivative=False)
            c = np.asarray(c)
            c[np.diag_indices_from(c)] += self.var_
            return np.asarray(m)[:, None], c
        m, v, _ = self.engine.predict(points)
        # GPy's predict returns the predictive variance of a noisy observation (latent + noise)
        return np.asarray(m)[:, None], (np.asarray(v) + self.var_)[:, None]

    def sample_sdfs(self, num_samples=1, full_cov=True, z=None, seed=None, jitter=1e-8):
        """gpis.py:98-112: draws of the SDF over the whole grid from the posterior, one array of
        shape dims per sample (C order, like grid_pts_).  GPy's posterior_samples (not vendored,
        version unpinned) is restated as the latent posterior + a diagonal `jitter`; the draws z
        [num_samples, prod(dims)] may be given, else they come from torch's generator.  full_cov is
        accepted and ignored, as in the reference (gpis.py:108 never reads it)."""
        import torch
        n = self.grid_pts_.shape[0]
        if z is None:
            gen = torch.Generator(device="cuda")
            gen.manual_seed(0 if seed is None else int(seed))
            z = torch.randn((int(num_samples), n), dtype=torch.float64, device="cuda", generator=gen)
        samples, _ = self.engine.sample_shapes(self.grid_pts_, z, use_derivative=False, jitter=jitter)
        if torch.is_tensor(samples):
            samples = samples.cpu().numpy()
        return [s.reshape(self.dims_) for s in samples]

    def mean_sdf(self):
        return self.mean_pred_.reshape(self.dims_)

    def variance(self):
        return self.var_pred_.reshape(self.dims_)

    def surface_points(self, surface_thresh=0.05):
        return self.grid_pts_[np.abs(self.mean_pred_[:, 0]) < surface_thresh]


class Gpis3D(_GpisBase):
    def __init__(self, points, sdf_meas, dims, **kw):
        if len(dims) != 3:
            raise ValueError("Dimensions must be 3D!")
        super().__init__(points, sdf_meas, dims, **kw)


class Gpis2D(_GpisBase):
    def __init__(self, points, sdf_meas, dims, **kw):
        if len(dims) != 2:
            raise ValueError("Dimensions must be 2D!")
        super().__init__(points, sdf_meas, dims, **kw)


def select_from_sdf(sdf, active_set_size, trunc=None, ell=None, noise=0.05, beta=10.0, eps=1e-2, level=0.0,
                    first_index=None):
    """Level-set active-set selection + posterior grid straight from an Sdf3D (voxel units):
    the lattice of the file is the candidate lattice of the grid backend."""
    from .synth import first_index as glibc_first
    nx, ny, nz = sdf.dimensions
    n = nx * ny * nz
    trunc = 4.0 * sdf.resolution if trunc is None else trunc
    ell = 4.0 if ell is None else ell
    eng = GpisEngine(3, n, active_set_size, ell=ell, noise=noise, level=level, eps=eps, beta=beta)
    eng.set_grid_candidates((nx, ny, nz), (0.0, 0.0, 0.0), (1.0, 1.0, 1.0), sdf.tsdf(trunc))
    eng.select(glibc_first(n) if first_index is None else first_index, active_set_size)
    ids, n_model = eng.active()
    mu, var = eng.candidate_posterior()
    return {"engine": eng, "ids": ids, "n_in_model": n_model,
            "mean": mu.reshape((nx, ny, nz), order="F"), "var": var.reshape((nx, ny, nz), order="F")}
