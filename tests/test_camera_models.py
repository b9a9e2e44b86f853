"""Oracle restatement of PinholeCamera::backProject3 with aslam's four distortion models
(camera-pinhole.cc:47-63; distortion-fisheye.cc:119-143, distortion-radtan.cc:96-118,
distortion-equidistant.cc:144-173): back-projecting the forward-distorted projection of a ray gives
the ray back, to the accuracy the reference's own stopping rule leaves (one Gauss-Newton step after
|e|^2 <= 1e-8)."""
import numpy as np
import pytest

from oracle import pyoracle as po
from test_gpu_ransac import EQUIDISTANT, RADTAN, distort


@pytest.mark.parametrize("model,coeffs,tol", [(0, (0, 0, 0, 0), 1e-15), (1, (0.9, 0, 0, 0), 1e-14),
                                              (2, RADTAN, 1e-7), (3, EQUIDISTANT, 1e-7)])
def test_back_project_inverts_the_forward_model(model, coeffs, tol):
    rng = np.random.default_rng(model)
    cam = po.make_camera(700.0, 701.0, 703.0, 530.0, None, None, model, coeffs)
    xy = rng.uniform(-0.9, 0.9, (4000, 2))
    xd, yd = distort(model, coeffs, xy[:, 0], xy[:, 1])
    b = po.back_project(cam, np.stack([700.0 * xd + 703.0, 701.0 * yd + 530.0], 1))
    assert np.abs(np.linalg.norm(b, axis=1) - 1).max() < 1e-15  # bearing.normalize()
    assert np.abs(b[:, :2] / b[:, 2:] - xy).max() < tol


def test_image_centre_special_cases():
    for model, coeffs in [(1, (0.9, 0, 0, 0)), (2, RADTAN), (3, EQUIDISTANT)]:
        cam = po.make_camera(700.0, 701.0, 703.0, 530.0, None, None, model, coeffs)
        b = po.back_project(cam, np.array([[703.0, 530.0], [703.0 + 0.3, 530.0 - 0.2]]))
        assert b[0].tolist() == [0.0, 0.0, 1.0]
        assert np.isfinite(b).all()
