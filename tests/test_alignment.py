"""Mission-level alignment: common::transformationRansac + LS quaternion average.
CPU: the oracle's restatement of libstdc++'s uniform_int_distribution<int>(0, n-1) is pinned against
the real std::uniform_int_distribution of this image's g++ (a tiny program compiled on the fly), and the
oracle RANSAC recovers a planted alignment. GPU: mlc_transformation_ransac == oracle (inlier set and
count exact, pose to 1e-9; the quaternion sign is not defined by the reference)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as po


def _rot(rv):
    rv = np.asarray(rv, np.float64)
    a = np.linalg.norm(rv)
    if a == 0:
        return np.array([0, 0, 0, 1.0])
    return np.concatenate([np.sin(a / 2) * rv / a, [np.cos(a / 2)]])


def _qmul(a, b):  # Hamilton, (x, y, z, w)
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz])


def _samples(n, seed, outlier_every=3):
    rng = np.random.default_rng(seed)
    q_true, p_true = _rot([0.02, -0.01, 0.7]), np.array([3.0, -2.0, 0.5])
    qs, ps = [], []
    for i in range(n):
        if i % outlier_every == 0:
            qs.append(_rot(rng.normal(size=3)))
            ps.append(rng.normal(size=3) * 5)
        else:
            qs.append(_qmul(q_true, _rot(rng.normal(size=3) * 0.01)))
            ps.append(p_true + rng.normal(size=3) * 0.05)
    return np.array(qs), np.array(ps), q_true, p_true


def _angle(qa, qb):
    return 2 * np.arctan2(np.linalg.norm(_qmul(qa, qb * [-1, -1, -1, 1])[:3]), abs(_qmul(qa, qb * [-1, -1, -1, 1])[3]))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_uniform_index_matches_this_libstdcxx(tmp_path):
    src = tmp_path / "u.cc"
    src.write_text('#include <random>\n#include <cstdio>\nint main(int c, char** v){ unsigned seed = atoi(v[1]); int n = atoi(v[2]);'
                   ' std::mt19937 g(seed); std::uniform_int_distribution<> d(0, n - 1);'
                   ' for (int i = 0; i < 64; ++i) printf("%d ", d(g)); return 0; }\n')
    exe = tmp_path / "u"
    subprocess.check_call(["g++", "-O1", "-include", "cstdlib", "-o", str(exe), str(src)])
    major = int(subprocess.check_output(["g++", "-dumpversion"]).decode().split(".")[0])
    mapping = 1 if major >= 11 else 0
    for seed, n in [(12345, 7), (1, 1000), (99, 3), (2024, 1 << 20), (5, 3000000000 % (1 << 31))]:
        ref = [int(x) for x in subprocess.check_output([str(exe), str(seed), str(n)]).decode().split()]
        assert po.uniform_indices(seed, mapping, n, 64).tolist() == ref


def test_oracle_transformation_ransac_recovers_planted_alignment():
    q, p, q_true, p_true = _samples(90, 0)
    oq, op, inl = po.transformation_ransac(q, p, 2000, 0.174, 2.0, 42)
    assert len(inl) == 60 and all(i % 3 != 0 for i in inl)
    assert np.abs(op - p_true).max() < 0.05 and _angle(oq, q_true) < 0.01
    # one sample: returned as is (geometry-inl.h:128-132); no iterations: {0} wins
    oq, op, inl = po.transformation_ransac(q[:1], p[:1], 2000, 0.174, 2.0, 42)
    assert np.array_equal(oq, q[0]) and np.array_equal(op, p[0]) and inl.tolist() == [0]
    oq, op, inl = po.transformation_ransac(q, p, 0, 0.174, 2.0, 42)
    assert inl.tolist() == [0] and np.array_equal(op, p[0])
    yaw = po.yaw_only(oq)
    assert yaw[0] == 0 and yaw[1] == 0 and abs(np.linalg.norm(yaw) - 1) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("n,seed,mapping,iters", [(90, 42, 1, 2000), (1500, 7, 1, 2000), (300, 3, 0, 50),
                                                  (2, 1, 1, 10), (1, 1, 1, 10), (64, 5, 1, 0)])
def test_device_transformation_ransac_matches_oracle(n, seed, mapping, iters):
    from maplab_b200 import capi
    from helpers import small_world
    _, blob, _, _ = small_world()
    det = capi.Detector(blob)
    q, p, _, _ = _samples(n, seed)
    eq, ep, einl = po.transformation_ransac(q, p, iters, 0.174, 2.0, seed, mapping)
    gq, gp, ginl = det.transformation_ransac(q, p, num_iterations=iters, seed=seed, rng_mapping=mapping)
    assert ginl.tolist() == einl.tolist()
    assert np.abs(gp - ep).max() <= 1e-9
    assert min(np.abs(gq - eq).max(), np.abs(gq + eq).max()) <= 1e-9
    with pytest.raises(capi.MlcError):
        det.transformation_ransac(q[:0], p[:0])


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_opengv_random_stream_matches_this_libstdcxx(tmp_path):
    # opengv's SampleConsensusProblem::rnd(): uniform_int_distribution<int>(0, INT_MAX) over mt19937
    # (SampleConsensusProblem.hpp:34-46, :157-161) — the oracle's stream, which the device
    # RANSAC consumes, against the real distribution of this image's libstdc++
    src = tmp_path / "r.cc"
    src.write_text('#include <random>\n#include <climits>\n#include <cstdio>\n#include <cstdlib>\n'
                   'int main(int c, char** v){ std::mt19937 g(atoi(v[1])); std::uniform_int_distribution<int> d(0, INT_MAX);'
                   ' for (int i = 0; i < 256; ++i) printf("%d ", d(g)); return 0; }\n')
    exe = tmp_path / "r"
    subprocess.check_call(["g++", "-O1", "-o", str(exe), str(src)])
    major = int(subprocess.check_output(["g++", "-dumpversion"]).decode().split(".")[0])
    mapping = 1 if major >= 11 else 0
    for seed in (12345, 1, 987654321):
        ref = [int(x) for x in subprocess.check_output([str(exe), str(seed)]).decode().split()]
        assert po.rng_stream(seed, mapping, 256).tolist() == ref


# ---- the reference's own tests of common::transformationRansac, restated -------------------------------
# common/maplab-common/test/test_geometry.cc:10-106 (GeometryTransformationRansac: OnlyInliers,
# InliersAndOutliers, OnlyOutliers). The fixture draws its samples from std::mt19937(++seed_), seed_ = 20,
# through std::uniform_real_distribution<double> — libstdc++'s generate_canonical takes two 32-bit draws,
# (lo + hi * 2^32) / 2^64 — reproduced here on numpy's legacy MT19937 (same init_genrand seeding).
class _GeometryFixture:
    def __init__(self):
        self.seed, self.q, self.p = 20, [], []

    @staticmethod
    def _uniform(rs, a, b):
        lo, hi = (int(x) for x in rs.randint(0, 2 ** 32, size=2, dtype=np.uint64))
        return a + (b - a) * ((lo + hi * 2.0 ** 32) / 2.0 ** 64)

    def _add(self, rs, metres, rad):
        pos = [self._uniform(rs, -metres, metres) for _ in range(3)]
        angle = self._uniform(rs, -rad, rad)  # AngleAxisd(angle, UnitX)
        self.q.append([np.sin(angle / 2), 0.0, 0.0, np.cos(angle / 2)])
        self.p.append(pos)

    def add_inliers(self, n):
        self.seed += 1
        rs = np.random.RandomState(self.seed)
        for _ in range(n):
            self._add(rs, 0.01, 0.01)

    def add_outlier(self):
        self.seed += 1
        self._add(np.random.RandomState(self.seed), 5.0, 0.5)

    def arrays(self):
        return np.array(self.q), np.array(self.p)


def _geometry_cases():
    a = _GeometryFixture(); a.add_inliers(20)
    b = _GeometryFixture(); b.add_inliers(20); [b.add_outlier() for _ in range(3)]
    c = _GeometryFixture(); [c.add_outlier() for _ in range(20)]
    return a, b, c


def _expect_geometry(run):
    """run(q, p) -> (quaternion, position, inliers) with the test's settings: 40 iterations, thresholds 0.1 rad /
    0.1 m, seed 42."""
    only_inliers, mixed, only_outliers = _geometry_cases()
    for fx, expected in ((only_inliers, 20), (mixed, 20)):
        q, p, inl = run(*fx.arrays())
        assert len(inl) == expected                                  # EXPECT_EQ(kNumInliers, num_inliers)
        assert np.abs(p).max() < 1e-1 and _angle(q, np.array([0, 0, 0, 1.0])) < 1e-1   # NEAR identity, 1e-1
        assert all(i < 20 for i in inl)
    q, p, inl = run(*only_outliers.arrays())
    assert len(inl) <= 1                                             # EXPECT_GE(1, num_inliers)


def test_oracle_passes_the_reference_transformation_ransac_tests():
    _expect_geometry(lambda q, p: po.transformation_ransac(q, p, 40, 0.1, 0.1, 42))


@pytest.mark.gpu
def test_device_passes_the_reference_transformation_ransac_tests():
    from maplab_b200 import capi
    from helpers import small_world
    _, blob, _, _ = small_world()
    det = capi.Detector(blob)
    _expect_geometry(lambda q, p: det.transformation_ransac(q, p, num_iterations=40, seed=42,
                                                            max_orientation_error_rad=0.1, max_position_error_m=0.1))


# ---- tail of detectLoopClosuresMissionToDatabase: inlier gate + yaw-only projection (host code of the library)
def _rpy_matrix(roll, pitch, yaw):
    cx, sx, cy, sy, cz, sz = np.cos(roll), np.sin(roll), np.cos(pitch), np.sin(pitch), np.cos(yaw), np.sin(yaw)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx  # RollPitchYawToRotationMatrix (geometry-inl.h:64-84)


def _quat_of(R):
    w = np.sqrt(max(1e-300, 1 + np.trace(R))) / 2
    if w > 1e-3:
        return np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1], 4 * w * w]) / (4 * w)
    z = np.sqrt(max(0.0, 1 - R[0, 0] - R[1, 1] + R[2, 2])) / 2
    return np.array([(R[0, 2] + R[2, 0]) / (4 * z), (R[1, 2] + R[2, 1]) / (4 * z), z, (R[1, 0] - R[0, 1]) / (4 * z)])


def test_yaw_only_projection_and_inlier_gate():
    from maplab_b200 import capi
    rng = np.random.default_rng(0)
    yaws = np.concatenate([rng.uniform(-np.pi, np.pi, 200), [0.0, 2.2, -2.2, 3.1, -3.1, 2 * np.pi / 3 + 1e-9]])
    for yaw in yaws:
        roll, pitch = rng.uniform(-0.2, 0.2, 2)
        q = _quat_of(_rpy_matrix(roll, pitch, yaw))
        got, exp = capi.alignment_yaw_only(q), po.yaw_only(q)
        assert np.abs(got - exp).max() <= 1e-15                      # library == oracle (independent restatements)
        assert got[0] == 0 and got[1] == 0 and abs(np.linalg.norm(got) - 1) < 1e-12
        # the yaw of the input survives: R = Rz(yaw)
        assert abs((1 - 2 * got[2] ** 2) - np.cos(yaw)) < 1e-9 and abs(2 * got[2] * got[3] - np.sin(yaw)) < 1e-9
        # Eigen's matrix -> quaternion sign: w >= 0 while 1 + 2 cos(yaw) > 0, else z > 0
        assert (got[3] >= 0) if 1 + 2 * np.cos(yaw) > 0 else (got[2] > 0)
        assert np.array_equal(capi.alignment_yaw_only(-q), got)      # q and -q are the same rotation
    # kNumInliersThreshold = max(min_inlier_count, int(samples * ratio)), defaults 10 / 0.2
    assert capi.alignment_enough_inliers(10, 30) and not capi.alignment_enough_inliers(9, 30)
    assert capi.alignment_enough_inliers(20, 100) and not capi.alignment_enough_inliers(19, 100)
    assert capi.alignment_enough_inliers(0, 0, 0, 0.0)
