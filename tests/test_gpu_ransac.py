"""Parity of kernel 4 (batched GP3P-RANSAC + handleLoopClosure gates) with the oracle, given the
same explicit random stream: sample sequence/iterations, best model indices, inlier lists and
verdicts bit-exact; transforms within 1e-6 m / 1e-6 rad (north_star) — in practice identical."""
import numpy as np
import pytest

from maplab_b200 import capi, synthetic
from oracle import pyoracle as po
from helpers import small_world

pytestmark = pytest.mark.gpu


def _rot(rng, scale=1.0):
    a = rng.normal(size=3) * scale
    th = np.linalg.norm(a)
    if th < 1e-12:
        return np.eye(3)
    k = a / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


RADTAN = (-0.28, 0.07, 2e-4, 2e-5)                 # EuRoC-like radial-tangential coefficients
EQUIDISTANT = (-0.0434, 0.00657, -0.00743, 0.00205)  # the Alphasense camera of maplab's common_test_map


def distort(model, coeffs, x, y):
    """Forward models of aslam's distortions (distortion-{fisheye,radtan,equidistant}.cc)."""
    if model == 1:
        w = coeffs[0]
        ru = np.hypot(x, y)
        rd = np.arctan(ru * 2 * np.tan(w / 2)) / w
        return x * rd / ru, y * rd / ru
    if model == 2:
        k1, k2, p1, p2 = coeffs
        r2 = x * x + y * y
        rad = k1 * r2 + k2 * r2 * r2
        return (x + x * rad + 2 * p1 * x * y + p2 * (r2 + 2 * x * x),
                y + y * rad + 2 * p2 * x * y + p1 * (r2 + 2 * y * y))
    if model == 3:
        k1, k2, k3, k4 = coeffs
        r = np.hypot(x, y)
        th = np.arctan(r)
        thd = th * (1 + k1 * th ** 2 + k2 * th ** 4 + k3 * th ** 6 + k4 * th ** 8)
        return x * thd / r, y * thd / r
    return x, y


def _cams(rng, n_cams, fisheye=False, model=0, coeffs=(0, 0, 0, 0)):
    cams = []
    for c in range(n_cams):
        R = np.eye(3) if c == 0 else _rot(rng, 0.3)
        t = np.zeros(3) if c == 0 else rng.uniform(-0.2, 0.2, 3)
        d = dict(fu=400.0 + 5 * c, fv=402.0, cu=376.0, cv=240.0, R_B_C=R, t_B_C=t)
        if fisheye:
            d.update(distortion=1, dist=(0.9, 0, 0, 0))
        elif model:
            d.update(distortion=model, dist=tuple(coeffs))
        cams.append(d)
    return cams


def _problem(rng, cams, n, outlier_frac, noise=0.7, dup_frac=0.0):
    """Random rig pose, points in front of each camera, noisy keypoints, wrong landmarks."""
    R_G_B, t_G_B = _rot(rng, 0.5), rng.uniform(-3, 3, 3)
    ci = rng.integers(0, len(cams), n).astype(np.int32)
    kp = np.zeros((n, 2))
    lm = np.zeros((n, 3))
    for i in range(n):
        c = cams[ci[i]]
        pc = np.array([rng.uniform(-2, 2), rng.uniform(-1.5, 1.5), rng.uniform(3, 10)])
        x, y = pc[0] / pc[2], pc[1] / pc[2]
        x, y = distort(c.get("distortion", 0), c.get("dist"), x, y)
        kp[i] = [c["fu"] * x + c["cu"], c["fv"] * y + c["cv"]]
        pb = np.asarray(c["R_B_C"]) @ pc + np.asarray(c["t_B_C"])
        lm[i] = R_G_B @ pb + t_G_B
    kp += rng.normal(0, noise, kp.shape)
    bad = rng.random(n) < outlier_frac
    lm[bad] += rng.uniform(-4, 4, (int(bad.sum()), 3))
    ki = np.arange(n, dtype=np.int32)
    ndup = int(dup_frac * n)
    if ndup:  # several candidate landmarks for the same keypoint (A22 picks the best)
        src = rng.integers(0, n, ndup)
        dst = rng.integers(0, n, ndup)
        ki[dst] = ki[src]
        ci[dst] = ci[src]
        kp[dst] = kp[src]
    return kp, ci, ki, lm, (R_G_B, t_G_B)


def _check(cams, problems, bit_identical=True, **rs_kw):
    m, blob, _, _ = small_world()
    det = capi.Detector(blob)
    rs = capi.default_ransac_settings(**rs_kw)
    offsets = np.concatenate([[0], np.cumsum([len(p[1]) for p in problems])]).astype(np.int64)
    kp = np.concatenate([p[0] for p in problems]) if problems else np.zeros((0, 2))
    ci = np.concatenate([p[1] for p in problems]) if problems else np.zeros(0, np.int32)
    ki = np.concatenate([p[2] for p in problems]) if problems else np.zeros(0, np.int32)
    lm = np.concatenate([p[3] for p in problems]) if problems else np.zeros((0, 3))
    res, flags = det.pnp_ransac_batch(capi.make_cameras(cams), offsets, kp, ci, ki, lm, rs)
    ocams = [po.make_camera(c["fu"], c["fv"], c["cu"], c["cv"], c["R_B_C"], c["t_B_C"],
                            c.get("distortion", 0), c.get("dist", (0, 0, 0, 0))) for c in cams]
    n_acc = 0
    for i, p in enumerate(problems):
        exp = po.handle_loop_closure(p[0], p[1], p[2], p[3], ocams, rs.min_inlier_count,
                                     rs.min_inlier_ratio, rs.ransac_pixel_sigma, rs.num_ransac_iters,
                                     rs.seed, rs.rng_mapping)
        r = res[i]
        f = flags[offsets[i]:offsets[i + 1]]
        assert bool(r["accepted"]) == exp["accepted"], i
        assert bool(r["ransac_success"]) == exp["ransac_success"], i
        assert int(r["iterations"]) == exp["iterations"], i
        assert int(r["num_inliers"]) == exp["num_inliers"], i
        if exp["ransac_success"]:
            assert r["model_indices"].tolist() == exp["model_indices"].tolist(), i
            assert np.nonzero(f)[0].tolist() == exp["inliers"].tolist(), i
            assert np.nonzero(f == 3)[0].tolist() == sorted(exp["best_per_keypoint"].tolist()), i
            T, Te = r["T_G_I"].reshape(3, 4), exp["T"]
            assert np.abs(T[:, 3] - Te[:, 3]).max() <= 1e-6          # 1e-6 m
            dR = T[:, :3] @ Te[:, :3].T
            ang = np.arccos(np.clip((np.trace(dR) - 1) / 2, -1, 1))
            assert ang <= 1e-6                                         # 1e-6 rad
            if bit_identical:
                assert np.array_equal(T, Te), "expected bit-identical fp64 arithmetic"
        assert abs(float(r["inlier_ratio"]) - exp["inlier_ratio"]) <= 1e-15
        n_acc += int(exp["accepted"])
    return res, n_acc


@pytest.mark.parametrize("n_cams,mapping", [(1, 1), (2, 1), (2, 0)])
def test_ransac_matches_oracle(n_cams, mapping):
    rng = np.random.default_rng(100 + n_cams + mapping)
    cams = _cams(rng, n_cams)
    problems = []
    for n, out in [(60, 0.3), (200, 0.5), (35, 0.1), (120, 0.75), (500, 0.4), (25, 0.95), (80, 0.0)]:
        problems.append(_problem(rng, cams, n, out)[:4])
    res, n_acc = _check(cams, problems, rng_mapping=mapping)
    assert n_acc >= 4
    # the recovered pose of a clean problem is the ground truth (sanity of the solver itself)
    kp, ci, ki, lm, (R, t) = _problem(rng, cams, 100, 0.2, noise=0.2)
    res, _ = _check(cams, [(kp, ci, ki, lm)], rng_mapping=mapping)
    T = res[0]["T_G_I"].reshape(3, 4)
    assert np.abs(T[:, 3] - t).max() < 0.05 and np.abs(T[:, :3] - R).max() < 0.02


def test_ransac_gates_and_edge_cases():
    rng = np.random.default_rng(7)
    cams = _cams(rng, 1)
    problems = [
        _problem(rng, cams, 9, 0.0)[:4],      # fewer matches than lc_min_inlier_count: no RANSAC
        _problem(rng, cams, 0, 0.0)[:4],      # empty constraint
        _problem(rng, cams, 40, 1.0)[:4],     # outliers only -> rejected
        _problem(rng, cams, 64, 0.2, dup_frac=0.4)[:4],  # several landmarks per keypoint
        _problem(rng, cams, 10, 0.0)[:4],     # exactly the minimum
    ]
    res, _ = _check(cams, problems)
    assert res[0]["accepted"] == 0 and res[0]["iterations"] == 0 and res[1]["accepted"] == 0
    # ratio gate and a tiny minimum count (n < 4 cannot be sampled: iterations = INT_MAX)
    _check(cams, [_problem(rng, cams, 50, 0.6)[:4], _problem(rng, cams, 3, 0.0)[:4]],
           min_inlier_ratio=0.5, min_inlier_count=2)
    _check(cams, [_problem(rng, cams, 90, 0.5)[:4]], num_ransac_iters=5, seed=777)
    _check(cams, [], seed=1)


@pytest.mark.parametrize("model,coeffs", [(2, RADTAN), (3, EQUIDISTANT)])
def test_ransac_radtan_and_equidistant_cameras(model, coeffs):
    # iteratively undistorted keypoints (Gauss-Newton on the forward model, <= 30 iterations, 1e-8)
    rng = np.random.default_rng(40 + model)
    cams = _cams(rng, 2, model=model, coeffs=coeffs)
    problems = [_problem(rng, cams, n, out) for n, out in [(80, 0.3), (300, 0.5), (40, 0.0), (150, 0.6)]]
    # one keypoint exactly at the principal point: the special case around the image centre
    problems[0][0][0] = [cams[problems[0][1][0]]["cu"], cams[problems[0][1][0]]["cv"]]
    # equidistant: atan() / pow() of CUDA and glibc differ in the last ulp, so the undistorted bearings (and the
    # poses) agree to ~1e-15 instead of bit for bit; north_star's bound is 1e-6 m / 1e-6 rad
    _, n_acc = _check(cams, problems, bit_identical=(model != 3))
    assert n_acc >= 3


def test_ransac_fisheye_camera():
    rng = np.random.default_rng(9)
    cams = _cams(rng, 2, fisheye=True)
    _check(cams, [_problem(rng, cams, 150, 0.3)[:4], _problem(rng, cams, 70, 0.5)[:4]])


def test_delta_pose_gate_matches_oracle():
    # handleLoopClosure's topological gate (loop-closure-handler.cc:424-455): off by default, needs
    # the query vertices' current poses, rejects closures that move the vertex too far
    from helpers import frames_of, small_world
    m, blob, _, q = small_world(num_queries=12)
    det = capi.Detector(blob, capi.default_settings(num_nearest_neighbors=6))
    det.insert_batch(frames_of(m["frames"]), det.project(m["bits"]), m["landmarks"])
    det.set_landmark_positions(m["landmark_xyz"])
    cams = capi.make_cameras([synthetic.camera_dict()])
    qframes = frames_of(q["frames"])
    base = det.query_batch(qframes, q["bits"], q["keypoints"], cams)["results"]
    assert base["accepted"].sum() >= 6
    # priors = ground truth poses perturbed by growing offsets / yaw angles
    rng = np.random.default_rng(3)
    priors = q["T_G_I"].copy()
    for i in range(len(priors)):
        priors[i, :, 3] += rng.normal(size=3) * 0.08 * i
        a = 0.01 * i
        Ry = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
        priors[i, :, :3] = priors[i, :, :3] @ Ry
    for max_pos, max_rot in [(0.5, -1.0), (-1.0, 3.0), (0.6, 4.0), (1e9, 1e9)]:
        rs = capi.default_ransac_settings(max_delta_position_m=max_pos, max_delta_rotation_deg=max_rot)
        det.set_query_priors(priors)
        got = det.query_batch(qframes, q["bits"], q["keypoints"], cams, rs=rs)["results"]
        exp = base["accepted"].copy()
        for i in range(len(exp)):
            if exp[i]:
                ok, _, _ = po.delta_pose_gate(priors[i], base["T_G_I"][i], max_pos, max_rot)
                exp[i] = 1 if ok else 0
        assert np.array_equal(got["accepted"], exp)
        assert got["T_G_I"].tobytes() == base["T_G_I"].tobytes()     # the gate only changes the verdict
    assert 0 < exp.sum()                                              # (1e9, 1e9): nothing rejected
    rs = capi.default_ransac_settings(max_delta_position_m=0.5)
    det.set_query_priors(priors)
    assert det.query_batch(qframes, q["bits"], q["keypoints"], cams, rs=rs)["results"]["accepted"].sum() < base["accepted"].sum()
    with pytest.raises(capi.MlcError):                                # priors were consumed by the last call
        det.query_batch(qframes, q["bits"], q["keypoints"], cams, rs=rs)
