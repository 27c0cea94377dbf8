"""tests/golden/tsdf_shapes.npz from the reference's stored shapeParams (data/google_objects/*.mat): the
occupancy image (fullTsdf <= 0), fullTsdf, fullNormals, the measured points and their noise, for the six
shapes that carry them.  Run in the build container (needs /root/reference):
    python tests/golden/make_tsdf_golden.py"""
import os

import numpy as np
import scipy.io as sio

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/data/google_objects"
out = {}
names = ["tape", "bowl", "deodorant", "marker", "oatmeal", "stapler"]
for name in names:
    m = sio.loadmat(os.path.join(REF, name + ".mat"), squeeze_me=True, struct_as_record=False)["shapeParams"]
    g = int(m.gridDim)
    ft = np.asarray(m.fullTsdf, np.float64)
    out[name + "_inside"] = (ft <= 0).reshape(g, g, order="F")
    out[name + "_fullTsdf"] = ft
    out[name + "_fullNormals"] = np.asarray(m.fullNormals, np.float64)
    out[name + "_points"] = np.asarray(m.points, np.float64)
    out[name + "_tsdf"] = np.asarray(m.tsdf, np.float64)
    out[name + "_normals"] = np.asarray(m.normals, np.float64)
    out[name + "_noise"] = np.asarray(m.noise, np.float64)
    out[name + "_com"] = np.asarray(m.com, np.float64)
out["names"] = np.array(names)
np.savez_compressed(os.path.join(HERE, "tsdf_shapes.npz"), **out)
print("wrote tsdf_shapes.npz", os.path.getsize(os.path.join(HERE, "tsdf_shapes.npz")), "bytes")
