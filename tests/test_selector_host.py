"""Host logic of the GpuActiveSetSelector mirror: config / CSV conventions (CPU only)."""
import numpy as np

from gpis_b200 import selector as S


def test_read_config_and_csv_roundtrip(tmp_path):
    w, h, d = 5, 4, 3
    vals = np.arange(w * h * d, dtype=np.float64).reshape(h * d, w) * 0.25
    csv = tmp_path / "tsdf.csv"
    np.savetxt(csv, vals, delimiter=",")
    cfg = tmp_path / "gpis.cfg"
    cfg.write_text(f"# comment\n{csv}\n100\n1.5\n0.1\n{w}\n{h}\n{d}\n64\n")
    c = S.read_config(str(cfg))
    assert c == {"csv": str(csv), "setSize": 100, "sigma": 1.5, "beta": 0.1, "width": w, "height": h, "depth": d, "batch": 64}
    inputs, targets = S.read_csv(str(csv), w, h, d, True)
    # IJK_TO_LINEAR(i, j, k) = i + w*j + w*h*k; x = column, y = row in slab, z = slab
    for (i, j, k) in [(0, 0, 0), (4, 3, 2), (2, 1, 1)]:
        lin = i + w * j + w * h * k
        assert tuple(inputs[:, lin]) == (i, j, k)
        assert targets[lin] == vals[k * h + j, i]
    inputs2, _ = S.read_csv(str(csv), w, h, d, False)
    assert inputs2.shape == (2, w * h * d)
    out = tmp_path / "o.csv"
    S.write_csv(str(out), np.arange(6.0), 2, 3)         # column-major buffer, 3 lines of 2
    assert np.array_equal(np.loadtxt(out, delimiter=","), np.array([[0, 3], [1, 4], [2, 5.0]]))


def test_prediction_error_matches_reference_definition():
    e = S.PredictionError([1.0, 2.0, 4.0])
    assert e.mean == 7.0 / 3 and e.std == 7.0 and e.median == 2.0 and e.min == 1.0 and e.max == 4.0
