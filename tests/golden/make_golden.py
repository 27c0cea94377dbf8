"""Generate tests/golden/*.npz from the reference's stored MATLAB data.

Run in the build container only (it reads /root/reference, which does not exist
on the GPU box):  python tests/golden/make_golden.py

The reference has no test suite; what pins this path's arithmetic is saved
output: gpModel structs (training_x, training_y, beta, hyp, Kxx, L, alpha) in
results/*.mat and data/google_update_versions/loofa_gpis.mat, and one matching
predGrid in loofa_construction.mat.  Arrays are copied
verbatim (cast to float64: MATLAB stores integer-valued doubles as uint8).
Kxx/L are kept in full for tape and loofa; for the other four models only
alpha, diag(L) and three rows of Kxx/L are kept to bound the fixture size.
"""
import os
import sys

import numpy as np
import scipy.io as sio

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def f64(a):
    return np.asarray(a, dtype=np.float64)


def model_dict(gm, full):
    d = {
        "training_x": f64(gm.training_x), "training_y": f64(gm.training_y),
        "beta": f64(gm.beta), "hyp_cov": f64(gm.hyp.cov), "hyp_mean": f64(gm.hyp.mean),
        "hyp_lik": f64(gm.hyp.lik), "alpha": f64(gm.alpha), "Mx": f64(gm.Mx), "sl": f64(gm.sl),
    }
    K, L = f64(gm.Kxx), f64(gm.L)
    assert np.array_equal(K, f64(gm.Q))
    if full:
        d["Kxx"], d["L"] = K, L
    else:
        rows = np.array([0, K.shape[0] // 2, K.shape[0] - 1])
        d["rows"], d["Kxx_rows"], d["L_rows"], d["L_diag"] = rows, K[rows], L[rows], np.diag(L).copy()
    return d


def shape_dict(sp):
    return {"points": f64(sp.points), "tsdf": f64(sp.tsdf), "normals": f64(sp.normals),
            "noise": f64(sp.noise), "gridDim": f64(sp.gridDim)}


def main():
    if not os.path.isdir(REF):
        sys.exit("reference tree not present; fixtures are already committed")
    for name in ["tape", "tape_occ", "marker_0.05", "marker_occ", "deodorant", "deodorant_occ"]:
        m = sio.loadmat(f"{REF}/results/{name}.mat", squeeze_me=True, struct_as_record=False)
        d = {("model_" + k): v for k, v in model_dict(m["gpModel"], full=(name == "tape")).items()}
        d.update({("shape_" + k): v for k, v in shape_dict(m["shapeParams"]).items()})
        np.savez_compressed(f"{OUT}/results_{name.replace('.', '_')}.npz", **d)
    g = sio.loadmat(f"{REF}/data/google_update_versions/loofa_gpis.mat", squeeze_me=True, struct_as_record=False)
    c = sio.loadmat(f"{REF}/data/google_update_versions/loofa_construction.mat", squeeze_me=True, struct_as_record=False)
    s = sio.loadmat(f"{REF}/data/google_update_versions/loofa.mat", squeeze_me=True, struct_as_record=False)
    d = {("model_" + k): v for k, v in model_dict(g["gpModel"], full=True).items()}
    d.update({("shape_" + k): v for k, v in shape_dict(s["shapeParams"]).items()})
    pg = c["constructionResults"].predGrid
    d.update({"pred_tsdf": f64(pg.tsdf), "pred_normals": f64(pg.normals), "pred_noise": f64(pg.noise),
              "pred_points": f64(pg.points), "pred_com": f64(pg.com), "pred_thresh": f64(pg.surfaceThresh)})
    te = c["constructionResults"].tsdfReconError
    d.update({"err_mean": f64(te.meanError), "err_median": f64(te.medianError),
              "err_rms": f64(te.rmsError), "err_std": f64(te.stdError),
              "shape_fullTsdf": f64(s["shapeParams"].fullTsdf)})
    np.savez_compressed(f"{OUT}/loofa.npz", **d)
    # what the reference stored about its posterior shape sampling of this model
    # (create_experiment_object.m:23-26,38-39): the densities of the 1000 draws (every one +inf) and the
    # wall-clock times of the dense fit and of sample_shapes
    cr = c["constructionResults"]
    np.savez_compressed(f"{OUT}/loofa_sampling.npz", pdfs=f64(cr.pdfs), samplingTime=f64(cr.samplingTime),
                        constructionTime=f64(cr.constructionTime))
    # real 3-D SDF (header asserted in sdf_file.py:91-99) and a 50x50 2-D csv (:116-117)
    with open(f"{REF}/data/test/sdf/Co_clean_dim_25.sdf") as f:
        nx, ny, nz = [int(t) for t in f.readline().split()]
        origin = np.array([float(t) for t in f.readline().split()])
        res = float(f.readline())
        vals = np.loadtxt(f)
    csv2d = np.loadtxt(f"{REF}/data/test/sdf/medium_black_spring_clamp_optimized_poisson_texture_mapped_mesh_clean_0.csv", delimiter=",")
    np.savez_compressed(f"{OUT}/sdf_co_clean_dim_25.npz", dims=np.array([nx, ny, nz]), origin=origin, resolution=res,
                        values=vals.astype(np.float32), csv2d_shape=np.array(csv2d.shape),
                        csv2d_checksum=np.array([csv2d.sum(), np.abs(csv2d).max()]))
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(f"{OUT}/{f}"))


if __name__ == "__main__":
    main()
