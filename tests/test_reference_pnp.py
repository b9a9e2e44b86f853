"""The reference's own test of the geometric-verification entry point, restated:
  aslam_cv2/aslam_cv_geometric_vision/test/test_pnp_pose_estimator_test.cc:101-186
  TEST_P(VariableCameraAngle, MultiPinholeCameraP3pInterface), angles {0, pi/6, pi/2, -pi/2, -pi/6}
(two fisheye pinhole cameras fu = fv = 200, 640 x 480, w = 0.95; rig extrinsics q_C_B = AngleAxis(pi/6 i, Y),
p_C_B = (2i-1, i-1, 5i); 800 points createRandomVisiblePoint(i + 50) under srand(1), the first 20 with wrong
landmarks; absoluteMultiPoseRansacPinholeCam(pixel_sigma 0.8, 500 iterations, fixed seed)), with the test's
expectations: position and quaternion within 1e-5 of the ground truth and exactly 780 inliers. The random
keypoints follow Eigen 3.3's setRandom (-1 + 2 rand()/RAND_MAX per coefficient) on glibc's rand(), which is
what the reference binary draws on this platform. The reference runs the test with the non-linear refinement
on; the loop-closure path (--lc_nonlinear_refinement_p3p=false) takes the RANSAC model itself, which must
meet the same tolerance on this noise-free data. (The single-camera TEST_P uses opengv's central KNEIP
solver, which is not on the loop-closure path.)
CPU part: the oracle. GPU part: the CUDA kernels, same expectations + bit-equality with the oracle."""
import ctypes
import ctypes.util

import numpy as np
import pytest

from oracle import pyoracle as po

ANGLES = [0.0, np.pi / 6.0, np.pi / 2.0, -np.pi / 2.0, -np.pi / 6.0]
FU = FV = 200.0
RU, RV = 640, 480
W = 0.95
NUM_POINTS, NUM_OUTLIERS = 800, 20


def rot_y(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


def fisheye_distort(xy):
    r_u = np.hypot(*xy)
    r_rd = 2 * np.tan(W / 2) / W if r_u * r_u < 1e-5 else np.arctan(2 * np.tan(W / 2) * r_u) / (r_u * W)
    return xy * r_rd


def scenario(angle):
    libc = ctypes.CDLL(ctypes.util.find_library("c"))
    libc.srand(1)
    RAND_MAX = 2147483647
    cams_dict, ocams = [], []
    for i in range(2):
        R_C_B, p_C_B = rot_y(np.pi / 6.0 * i), np.array([2.0 * i - 1, i - 1.0, i * 5.0])
        R_B_C, t_B_C = R_C_B.T, -R_C_B.T @ p_C_B
        cams_dict.append(dict(fu=FU, fv=FV, cu=RU / 2, cv=RV / 2, R_B_C=R_B_C, t_B_C=t_B_C, distortion=1,
                              dist=(W, 0, 0, 0)))
        ocams.append(po.make_camera(FU, FV, RU / 2, RV / 2, R_B_C, t_B_C, 1, (W, 0, 0, 0)))
    R_G_B, p_G_B = rot_y(angle), np.array([1.0, 2.0, 3.0])
    kp = np.zeros((NUM_POINTS, 2))
    lm = np.zeros((NUM_POINTS, 3))
    ci = (np.arange(NUM_POINTS) % 2).astype(np.int32)
    border = min(RU, RV) * 0.1
    for i in range(NUM_POINTS):
        c = cams_dict[ci[i]]
        out = np.array([-1.0 + 2.0 * libc.rand() / RAND_MAX, -1.0 + 2.0 * libc.rand() / RAND_MAX])
        y = np.array([border + abs(out[0]) * (RU - border * 2.0), border + abs(out[1]) * (RV - border * 2.0)])
        ray = po.back_project(ocams[ci[i]], y[None])[0]  # camera frame (backProject3 incl. undistortion)
        p_C = ray / np.linalg.norm(ray) * (i + 50)
        d = fisheye_distort(p_C[:2] / p_C[2])            # project3
        kp[i] = [FU * d[0] + RU / 2, FV * d[1] + RV / 2]
        if i < NUM_OUTLIERS:                             # unsigned integer arithmetic of the test
            p_C = np.array([i // 10, i // 4, ((i - 1) & 0xFFFFFFFF) // 3], np.float64)
        p_B = c["R_B_C"] @ p_C + c["t_B_C"]              # T_G_C = T_G_B * T_C_B^-1
        lm[i] = R_G_B @ p_B + p_G_B
    return kp, ci, lm, cams_dict, ocams, R_G_B, p_G_B


def quat_xyzw(R):
    w = np.sqrt(max(0.0, 1 + R[0, 0] + R[1, 1] + R[2, 2])) / 2
    q = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1], 4 * w * w]) / (4 * w)
    return q


def expectations(T, inliers, R_G_B, p_G_B):
    assert np.abs(T[:, 3] - p_G_B).max() < 1e-5                      # EIGEN_MATRIX_NEAR(position, 1e-5)
    q, qe = quat_xyzw(T[:, :3]), quat_xyzw(R_G_B)
    assert min(np.abs(q - qe).max(), np.abs(q + qe).max()) < 1e-5    # EIGEN_MATRIX_NEAR(coeffs, 1e-5)
    assert len(inliers) == NUM_POINTS - NUM_OUTLIERS                 # EXPECT_EQ(inliers.size(), 780)
    assert inliers.tolist() == list(range(NUM_OUTLIERS, NUM_POINTS))


@pytest.mark.parametrize("angle", ANGLES)
def test_oracle_passes_the_reference_multi_camera_pnp_test(angle):
    kp, ci, lm, _, ocams, R_G_B, p_G_B = scenario(angle)
    # keypoints are the projections of the points they were back-projected to
    out = po.handle_loop_closure(kp, ci, np.arange(NUM_POINTS, dtype=np.int32), lm, ocams, pixel_sigma=0.8,
                                 num_iters=500, seed=12345)
    assert out["ransac_success"]
    expectations(out["T"], out["inliers"], R_G_B, p_G_B)


@pytest.mark.gpu
def test_device_passes_the_reference_multi_camera_pnp_test():
    from maplab_b200 import capi
    from helpers import small_world
    _, blob, _, _ = small_world()
    det = capi.Detector(blob)
    probs = [scenario(a) for a in ANGLES]
    cams = capi.make_cameras(probs[0][3])
    rs = capi.default_ransac_settings(ransac_pixel_sigma=0.8, num_ransac_iters=500)
    offsets = np.arange(len(ANGLES) + 1, dtype=np.int64) * NUM_POINTS
    ki = np.tile(np.arange(NUM_POINTS, dtype=np.int32), len(ANGLES))
    res, flags = det.pnp_ransac_batch(cams, offsets, np.concatenate([p[0] for p in probs]),
                                      np.concatenate([p[1] for p in probs]), ki,
                                      np.concatenate([p[2] for p in probs]), rs)
    for i, (kp, ci, lm, _, ocams, R_G_B, p_G_B) in enumerate(probs):
        T = res[i]["T_G_I"].reshape(3, 4)
        inl = np.nonzero(flags[offsets[i]:offsets[i + 1]])[0]
        assert res[i]["ransac_success"]
        expectations(T, inl, R_G_B, p_G_B)
        exp = po.handle_loop_closure(kp, ci, ki[:NUM_POINTS], lm, ocams, pixel_sigma=0.8, num_iters=500,
                                     seed=rs.seed, rng_mapping=rs.rng_mapping)
        assert int(res[i]["iterations"]) == exp["iterations"]
        assert np.array_equal(T, exp["T"])
