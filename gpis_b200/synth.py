"""Synthetic TSDF inputs of BASELINE.json's configs.

Shapes are deliberately off-centre and jittered: a centred shape on a symmetric grid
ties the straddle score exactly at every step."""
import numpy as np


def grid_points(dims):
    """Integer voxel coordinates, x fastest (IJK_TO_LINEAR, active_set_selection_types.h:6)."""
    dims = tuple(int(d) for d in dims)
    idx = np.indices(dims[::-1]).reshape(len(dims), -1)[::-1]
    return np.ascontiguousarray(idx.T.astype(np.float64))


def sphere_tsdf(g, radius=None, centre=None, trunc=4.0, jitter=1e-3, seed=0, with_normals=False):
    pts = grid_points((g, g, g))
    s = g / 64.0
    c = np.array([31.637, 31.289, 31.819]) * s if centre is None else np.asarray(centre, float)
    r = 20.3 * s if radius is None else radius
    dist = np.linalg.norm(pts - c, axis=1)
    sd = (dist - r) / trunc
    tsdf = np.clip(sd, -1.0, 1.0) + jitter * np.random.default_rng(seed).standard_normal(pts.shape[0])
    if not with_normals:
        return pts, tsdf
    nrm = (pts - c) / np.maximum(dist, 1e-12)[:, None] / trunc
    nrm[np.abs(sd) >= 1.0] = 0.0
    return pts, tsdf, nrm


def box_tsdf(g, trunc=4.0, jitter=1e-3, seed=0):
    pts = grid_points((g, g, g))
    s = g / 128.0
    c = np.array([63.637, 63.289, 63.819]) * s
    half = np.array([37.3, 29.1, 22.7]) * s
    q = np.abs(pts - c) - half
    sd = np.linalg.norm(np.maximum(q, 0.0), axis=1) + np.minimum(q.max(axis=1), 0.0)
    tsdf = np.clip(sd / trunc, -1.0, 1.0) + jitter * np.random.default_rng(seed).standard_normal(pts.shape[0])
    return pts, tsdf


def first_index(n):
    """glibc srand(1000); rand() % N, as main.cpp:77 + gpu_active_set_selector.cpp:470."""
    return 766020790 % int(n)


CONFIGS = {
    # name: (generator, grid, ell, noise, M)
    "c2_sphere64": dict(shape="sphere", g=64, ell=4.0, noise=0.05, K=1000),
    "c3_box128": dict(shape="box", g=128, ell=4.0, noise=0.05, K=4000),
}


def make(name_or_shape, g=None):
    cfg = dict(CONFIGS[name_or_shape]) if name_or_shape in CONFIGS else dict(shape=name_or_shape, g=g, ell=4.0, noise=0.05, K=0)
    if g is not None:
        cfg["g"] = g
    pts, tsdf = (sphere_tsdf if cfg["shape"] == "sphere" else box_tsdf)(cfg["g"])
    return pts, tsdf, cfg
