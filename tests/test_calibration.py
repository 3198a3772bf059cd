"""The reference's calibration JSON formats (include/visnav/serialization.h:92-174)."""
import os

import numpy as np
import pytest

import pba_b200 as pb

HERE = os.path.dirname(os.path.abspath(__file__))
REAL = "/root/reference/data/euroc_calib/calibration-double-sphere.json"


def test_double_sphere_schema_fixture():
    c = pb.load_calibration(os.path.join(HERE, "data", "calib_ds_schema.json"))
    assert c.models == ["ds", "ds"]
    assert list(c.calib_model) == [pb.CAM_DS, pb.CAM_DS]
    np.testing.assert_allclose(c.intrinsics[1], [348.75, 349.0, 372.5, 238.0, -0.21, 0.56, 0, 0])
    # Sophus parameter order qx qy qz qw tx ty tz
    np.testing.assert_allclose(c.T_i_c[1], [0.001, -0.002, 0.0005, 0.9999973750, 0.11, -0.002, 0.0007])


def test_generic_schema_round_trip(tmp_path):
    c = pb.Calibration(np.array([[0, 0, 0, 1, 0, 0, 0], [0, 0, 0.1, 0.99498744, 1, 2, 3]]), ["pinhole", "kb4"],
                       [[370.3, 370.3, 375.5, 239.5, 0, 0, 0, 0], [379.0, 379.1, 375.5, 239.5, 0.0069, -0.0014, -0.00027, -0.00045]],
                       [752, 752], [480, 480])
    path = tmp_path / "opt_calib.json"
    pb.save_calibration(path, c)
    d = pb.load_calibration(path)
    assert d.models == ["pinhole", "kb4"] and d.widths == [752, 752] and d.heights == [480, 480]
    np.testing.assert_array_equal(d.intrinsics, c.intrinsics)
    np.testing.assert_array_equal(d.T_i_c, c.T_i_c)


def test_unknown_model_is_rejected():
    with pytest.raises(ValueError, match="is not implemented"):
        pb.Calibration(np.zeros((1, 7)), ["fisheye624"], np.zeros((1, 8)))


def test_initialize_from_double_sphere():
    ds = [370.0, 371.0, 375.5, 239.5, -0.2, 0.55, 0, 0]
    np.testing.assert_array_equal(pb.initialize_from_double_sphere("ds", ds), ds)
    np.testing.assert_array_equal(pb.initialize_from_double_sphere("pinhole", ds), [370.0, 371.0, 375.5, 239.5, 0, 0, 0, 0])
    np.testing.assert_array_equal(pb.initialize_from_double_sphere("eucm", ds), [370.0, 371.0, 375.5, 239.5, 0.5, 1.0, 0, 0])


@pytest.mark.skipif(not os.path.exists(REAL), reason="reference data only present in the build container")
def test_real_euroc_calibration_file():
    c = pb.load_calibration(REAL)
    assert c.models == ["ds", "ds"]
    np.testing.assert_allclose(c.intrinsics[0][:6], [370.3418125824944, 370.3418125824944, 375.5, 239.5, 0.0, 0.5])
    np.testing.assert_allclose(c.T_i_c[0], [0, 0, 0, 1, 0, 0, 0])
