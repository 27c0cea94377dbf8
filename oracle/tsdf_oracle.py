"""CPU restatement of matlab/core/auto_tsdf.m:129-262 (everything after the Vision-toolbox rasteriser):
3x3 dilate / erode masks -> five-level TSDF, the per-pixel measurement-noise model, central-difference
normals, and the subsampling to the measured points.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  The GPU path (gpis_build_tsdf) is compared with this.

Pinned by reference data (tests/golden/tsdf_shapes.npz, from data/google_objects/*.mat): with the occupancy
image recovered as fullTsdf <= 0, `tsdf_levels` reproduces the stored fullTsdf and `central_gradients` the
stored fullNormals EXACTLY on six shapes.  The first stage of auto_tsdf.m (vision.ShapeInserter rasterising
the polygon, :129-131) needs MATLAB's Computer Vision toolbox and is not restated: the occupancy image is
the input.  The noise model draws rand() in a data-dependent order (:170-189) that cannot be reproduced, so
the uniform draws are explicit inputs (u[0]: interior, u[1]: specular scale, u[2]: sparsity) -- that stage
is PARITY UNPINNED by reference data beyond its deterministic properties (occlusion windows, the level
thresholds, the stored noiseScale of every measured point), which tests/test_oracle_tsdf.py checks.

Images are MATLAB matrices: element (i, j) = (row, column) = (y, x), flattened column-major (i + j*rows),
1-based i, j in the noise model's thresholds.
"""
import numpy as np

GRAD_MODES = {"None": 0, "TLBR": 1, "TRBL": 2, "BLTR": 3, "BRTL": 4}


def _window_any(mask, half):
    """True where any element of `mask` within the (2*half+1)^2 window (clipped at the border, as imdilate /
    imerode do) is True."""
    r, c = mask.shape
    out = np.zeros_like(mask, dtype=bool)
    for di in range(-half, half + 1):
        for dj in range(-half, half + 1):
            i0, i1 = max(0, di), min(r, r + di)
            j0, j1 = max(0, dj), min(c, c + dj)
            out[i0 - di:i1 - di, j0 - dj:j1 - dj] |= mask[i0:i1, j0:j1]
    return out


def tsdf_levels(inside):
    """auto_tsdf.m:133-157.  inside: bool [rows, cols] (the rasterised shape, image value 102; outside 255).
    imdilate = window max (255 if any outside), imerode = window min (102 if any inside); J_2d is the
    dilation of the dilation = the 5x5 window."""
    inside = np.asarray(inside, bool)
    outside = ~inside
    out_di = _window_any(outside, 1)          # J_d == 255
    in_di = ~out_di                           # J_d == 102
    in_er = _window_any(inside, 1)            # J_e == 102
    out_er = ~in_er                           # J_e == 255
    out_di2 = _window_any(outside, 2)         # J_2d == 255
    in_di2 = ~out_di2
    tsdf = np.zeros(inside.shape)
    tsdf[out_di & inside] = 0.0               # surfaceMask        :146, :153
    tsdf[out_di2 & in_di] = -0.5              # surfaceMaskIn      :148, :154
    tsdf[outside & in_er] = 0.5               # surfaceMaskOut     :147, :155
    tsdf[in_di2] = -1.0                       # insideMask         :149, :156
    tsdf[out_er] = 1.0                        # outsideMask        :150, :157
    return tsdf


def noise_model(tsdf, var_params, u):
    """auto_tsdf.m:159-218 with the rand() draws made explicit: u [3, rows, cols] uniform(0, 1)."""
    vp = var_params
    rows, cols = tsdf.shape
    noise = np.ones((rows, cols))
    mode = vp.get("noiseGradMode", "None")
    ew = int(vp["edgeWin"])
    for i in range(1, rows + 1):
        for j in range(1, cols + 1):
            t = tsdf[i - 1, j - 1]
            win = tsdf[max(1, i - ew) - 1:min(rows, i + ew), max(1, j - ew) - 1:min(cols, j + ew)]
            occluded = any(i > vp["y_thresh%d_low" % k] and i <= vp["y_thresh%d_high" % k] and
                           j > vp["x_thresh%d_low" % k] and j <= vp["x_thresh%d_high" % k] for k in (1, 2, 3))
            if t < 0.6 and occluded:
                noise[i - 1, j - 1] = vp["occlusionScale"]
            elif t < -0.5:
                noise[i - 1, j - 1] = vp["noiseScale"] if u[0, i - 1, j - 1] > 1 - vp["interiorRate"] else vp["occlusionScale"]
            else:
                nv = 1.0
                if vp["specularNoise"] and np.min(np.abs(win)) < 0.6:
                    nv = u[1, i - 1, j - 1]
                    if u[2, i - 1, j - 1] > 1 - vp["sparsityRate"]:
                        nv = vp["occlusionScale"] / vp["noiseScale"]
                ns, hs, vs = vp["noiseScale"], vp.get("horizScale", 1.0), vp.get("vertScale", 1.0)
                if mode == "TLBR":
                    noise[i - 1, j - 1] = nv * (ns * hs * j + ns * vs * i)
                elif mode == "TRBL":
                    noise[i - 1, j - 1] = nv * (ns * hs * (cols - j + 1) + ns * vs * i)
                elif mode == "BLTR":
                    noise[i - 1, j - 1] = nv * (ns * hs * j + ns * vs * (rows - i + 1))
                elif mode == "BRTL":
                    noise[i - 1, j - 1] = nv * (ns * hs * (cols - j + 1) + ns * vs * (rows - i + 1))
                else:
                    noise[i - 1, j - 1] = nv * ns
    return noise


def central_gradients(tsdf):
    """imgradientxy(tsdf, 'CentralDifference') (:221): interior (f(k+1) - f(k-1)) / 2, one-sided at the border.
    Gx along the columns (x), Gy along the rows (y)."""
    t = np.asarray(tsdf, np.float64)
    gx = np.zeros_like(t)
    gy = np.zeros_like(t)
    if t.shape[1] > 1:
        gx[:, 1:-1] = (t[:, 2:] - t[:, :-2]) / 2
        gx[:, 0] = t[:, 1] - t[:, 0]
        gx[:, -1] = t[:, -1] - t[:, -2]
    if t.shape[0] > 1:
        gy[1:-1, :] = (t[2:, :] - t[:-2, :]) / 2
        gy[0, :] = t[1, :] - t[0, :]
        gy[-1, :] = t[-1, :] - t[-2, :]
    return gx, gy


def auto_tsdf(inside, var_params, u, com=None):
    """auto_tsdf.m:133-262 from the rasterised shape on: returns shapeParams (a dict with the reference's
    field names).  Points are [X(:) Y(:)] of meshgrid(1:cols, 1:rows): x = column, y = row, column-major."""
    inside = np.asarray(inside, bool)
    rows, cols = inside.shape
    tsdf = tsdf_levels(inside)
    noise = noise_model(tsdf, var_params, np.asarray(u, np.float64))
    gx, gy = central_gradients(tsdf)
    X, Y = np.meshgrid(np.arange(1, cols + 1), np.arange(1, rows + 1))
    f = lambda a: np.asarray(a).reshape(-1, order="F")
    valid = np.flatnonzero(f(noise) < var_params["occlusionScale"])            # :229
    sp = {"gridDim": rows, "tsdf": f(tsdf)[valid], "noise": f(noise)[valid],
          "normals": np.stack([f(gx)[valid], f(gy)[valid]], axis=1),
          "points": np.stack([f(X)[valid], f(Y)[valid]], axis=1).astype(np.float64),
          "all_points": np.stack([f(X), f(Y)], axis=1).astype(np.float64),
          "fullTsdf": f(tsdf), "fullNormals": np.stack([f(gx), f(gy)], axis=1), "fullNoise": f(noise), "validIndices": valid}
    neg = sp["tsdf"] < 0
    sp["com"] = np.asarray(com, np.float64) if com is not None else (sp["points"][neg].mean(axis=0) if neg.any() else np.full(2, np.nan))
    return sp
