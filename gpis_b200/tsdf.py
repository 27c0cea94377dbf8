"""Host-side mirror of matlab/core/auto_tsdf.m (from the rasterised shape on, :133-262): the step before the
selection path.  `auto_tsdf` returns the reference's shapeParams struct as a dict; the image processing, the
noise model, the normals and the subsampling run in gpis_build_tsdf (gpis_b200/csrc/gpis_tsdf.cuh) -- nothing
here computes on the host beyond gathering the measured subset out of the full arrays.

The reference rasterises a polygon with vision.ShapeInserter (Computer Vision toolbox, :129-131); that stage is
outside this path -- pass the occupancy image (True inside).  Its rand() draws (interior subsampling, specular
noise) are the explicit input `u` [3, rows, cols] (or a seed for numpy's generator)."""
import ctypes as C

import numpy as np

from . import _lib

GRAD_MODES = {"None": 0, "TLBR": 1, "TRBL": 2, "BLTR": 3, "BRTL": 4}


def _params(vp):
    p = _lib.GpisTsdfParams()
    p.struct_size = C.sizeof(_lib.GpisTsdfParams)
    p.edge_win = int(vp["edgeWin"])
    p.specular_noise = int(bool(vp["specularNoise"]))
    p.noise_grad_mode = GRAD_MODES[vp.get("noiseGradMode", "None")]
    p.noise_scale, p.occlusion_scale = float(vp["noiseScale"]), float(vp["occlusionScale"])
    p.interior_rate, p.sparsity_rate = float(vp["interiorRate"]), float(vp["sparsityRate"])
    p.horiz_scale, p.vert_scale = float(vp.get("horizScale", 1.0)), float(vp.get("vertScale", 1.0))
    for k in range(3):
        p.occ_y_low[k], p.occ_y_high[k] = float(vp["y_thresh%d_low" % (k + 1)]), float(vp["y_thresh%d_high" % (k + 1)])
        p.occ_x_low[k], p.occ_x_high[k] = float(vp["x_thresh%d_low" % (k + 1)]), float(vp["x_thresh%d_high" % (k + 1)])
    return p


def build_tsdf(inside, varParams, u=None, device=-1):
    """gpis_build_tsdf on a [rows, cols] occupancy image: (tsdf, Gx, Gy, noise) as [rows, cols] arrays and the
    0-based column-major linear indices of the measured pixels."""
    lib = _lib.load()
    img = np.asfortranarray(np.asarray(inside) != 0, dtype=np.uint8)
    rows, cols = img.shape
    n = rows * cols
    uu = None if u is None else np.ascontiguousarray(np.stack([np.asarray(a, np.float64).reshape(rows, cols).reshape(-1, order="F")
                                                                  for a in u]))
    tsdf, noise, nrm = np.empty(n), np.empty(n), np.empty((2, n))
    idx = np.empty(n, np.int32)
    nv = C.c_int32()
    prm = _params(varParams)
    st = lib.gpis_build_tsdf(img.ctypes.data_as(C.c_void_p), rows, cols, C.byref(prm),
                             None if uu is None else uu.ctypes.data_as(C.c_void_p), tsdf.ctypes.data_as(C.c_void_p),
                             nrm.ctypes.data_as(C.c_void_p), noise.ctypes.data_as(C.c_void_p), idx.ctypes.data_as(C.c_void_p),
                             C.byref(nv), int(device), None)
    _lib.check(lib, None, st)
    sq = lambda a: a.reshape(rows, cols, order="F")
    return sq(tsdf), sq(nrm[0]), sq(nrm[1]), sq(noise), idx[:nv.value].astype(np.int64)


def auto_tsdf(inside, varParams, u=None, seed=None, com=None, device=-1):
    """shapeParams of auto_tsdf.m:246-262 for the rasterised shape `inside` ([rows, cols], True inside)."""
    inside = np.asarray(inside) != 0
    rows, cols = inside.shape
    if u is None and seed is not None:
        u = np.random.default_rng(seed).uniform(size=(3, rows, cols))
    tsdf, gx, gy, noise, valid = build_tsdf(inside, varParams, u, device)
    f = lambda a: np.asarray(a).reshape(-1, order="F")
    X, Y = np.meshgrid(np.arange(1, cols + 1), np.arange(1, rows + 1))           # :222-223
    sp = {"gridDim": rows, "vertices": None,
          "tsdf": f(tsdf)[valid], "noise": f(noise)[valid], "normals": np.stack([f(gx)[valid], f(gy)[valid]], axis=1),
          "points": np.stack([f(X)[valid], f(Y)[valid]], axis=1).astype(np.float64),
          "all_points": np.stack([f(X), f(Y)], axis=1).astype(np.float64),
          "fullTsdf": f(tsdf), "fullNormals": np.stack([f(gx), f(gy)], axis=1), "fullNoise": f(noise), "validIndices": valid}
    neg = sp["tsdf"] < 0
    sp["com"] = np.asarray(com, np.float64) if com is not None else (sp["points"][neg].mean(axis=0) if neg.any() else np.full(2, np.nan))
    return sp
