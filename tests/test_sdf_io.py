"""SDF file format (sdf_file.py) -- CPU part: header / ordering conventions on the reference's
own fixture data/test/sdf/Co_clean_dim_25.sdf (values committed as tests/golden/sdf_co_clean_dim_25.npz)."""
import numpy as np
import pytest

from conftest import load_golden
from gpis_b200.sdf_io import Sdf3D, SdfFile


def test_sdf_header_and_order_roundtrip(tmp_path):
    g = load_golden("sdf_co_clean_dim_25")
    nx, ny, nz = [int(v) for v in g["dims"]]
    vals = g["values"].astype(np.float64)
    path = tmp_path / "Co_clean_dim_25.sdf"
    with open(path, "w") as f:
        f.write(f"{nx} {ny} {nz}\n-1.04021 -1.1007 -1.01391\n0.0852112\n")
        np.savetxt(f, vals, fmt="%.9g")
    sdf = SdfFile(str(path)).read()
    # the asserts of the reference's own self-test, sdf_file.py:91-99
    assert sdf.dimensions == (25, 25, 25)
    assert tuple(sdf.origin) == (-1.04021, -1.1007, -1.01391) and sdf.resolution == 0.0852112
    assert tuple(g["origin"]) == (-1.04021, -1.1007, -1.01391) and float(g["resolution"]) == 0.0852112
    # x fastest: the n-th value of the file is data[i][j][k] with n = i + nx*(j + ny*k) (:52-56)
    for (i, j, k) in [(0, 0, 0), (24, 0, 0), (3, 7, 11), (24, 24, 24)]:
        assert sdf.data[i, j, k] == pytest.approx(vals[i + nx * (j + ny * k)], rel=1e-8)
    assert np.allclose(sdf.flat(), vals, rtol=1e-8)
    out = tmp_path / "rt.sdf"
    SdfFile(str(out)).write(sdf)
    again = SdfFile(str(out)).read()
    assert np.array_equal(again.data, sdf.data) and np.array_equal(again.origin, sdf.origin)
    assert tuple(g["csv2d_shape"]) == (50, 50)                                  # sdf_file.py:116-117
    with pytest.raises(ValueError):
        SdfFile("shape.obj")


def test_2d_csv(tmp_path):
    a = np.arange(12.0).reshape(3, 4)
    p = tmp_path / "s.csv"
    np.savetxt(p, a, delimiter=",")
    s = SdfFile(str(p)).read()
    assert s.dimensions == (3, 4) and np.array_equal(s.data, a)
