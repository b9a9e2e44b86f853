"""Real data end to end: maplab's own test map (tools/maplab-test-data/test_maps/common_test_map — real BRISK
descriptors of a 5-camera Alphasense rig, triangulated landmarks, equidistant pinhole cameras) with the SHIPPED
BRISK vocabulary. tests/golden/extract_real_map.py made the fixture (two of the five cameras, the keypoints `lc`
would feed the detector). The vertices are split like a multi-session map: even vertices form the database
mission, odd vertices are the query mission (another mission id, so the 10 s time filter does not apply) —
LoopDetectorNode::detectLoopClosuresMissionToDatabase's situation (LCH/src/loop-detector-node.cc:875-1005).
CPU: the oracle relocalises the query vertices against the map's own bundle-adjusted poses (the ground truth the
reference map carries). GPU: the CUDA path returns the oracle's matches / verdicts / poses."""
import os

import numpy as np
import pytest

from maplab_b200 import capi
from oracle import pyoracle as po

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
BLOB = open(os.path.join(GOLDEN, "brisk_quantizer_top10.dat"), "rb").read()


def world():
    d = np.load(os.path.join(GOLDEN, "real_map_brisk.npz"))
    fr = d["frames"]  # timestamp, vertex, camera, descriptors
    off = np.concatenate([[0], np.cumsum(fr[:, 3])])
    is_db = fr[:, 1] % 2 == 0

    def rows(mask):
        return np.concatenate([np.arange(off[i], off[i + 1]) for i in np.nonzero(mask)[0]])

    db, q = fr[is_db], fr[~is_db]
    return dict(
        db_frames=capi.make_frames(db[:, 0], db[:, 1], np.zeros(len(db), np.int64), db[:, 2], db[:, 3]),
        db_rows=rows(is_db),
        q_frames=capi.make_frames(q[:, 0], q[:, 1], np.ones(len(q), np.int64), q[:, 2], q[:, 3]),
        q_rows=rows(~is_db), q_vertices=q[::2, 1], bits=d["bits"], keypoints=d["keypoints"],
        landmarks=d["landmarks"].astype(np.int64), landmark_xyz=d["landmark_xyz"], T_G_I=d["T_G_I"],
        cameras=d["cameras"])


def camera_dicts(w):
    out = []
    for c in w["cameras"]:  # fu fv cu cv | k1..k4 | T_B_C row-major
        T = c[8:24].reshape(4, 4)
        out.append(dict(fu=c[0], fv=c[1], cu=c[2], cv=c[3], R_B_C=T[:3, :3], t_B_C=T[:3, 3], distortion=3,
                        dist=tuple(c[4:8])))
    return out


def oracle_side(w, proj_db):
    ora = po.Engine(BLOB, po.default_settings())
    at = 0
    f = w["db_frames"]
    for i in range(len(f)):
        n = int(f["num_descriptors"][i])
        rows = w["db_rows"][at:at + n]
        ora.insert(int(f["timestamp_ns"][i]), int(f["vertex_id"][i]), int(f["frame_index"][i]), 0, proj_db[at:at + n],
                   w["landmarks"][rows])
        at += n
    cams = [po.make_camera(c["fu"], c["fv"], c["cu"], c["cv"], c["R_B_C"], c["t_B_C"], 3, c["dist"])
            for c in camera_dicts(w)]
    exp = po.query_batch(ora, w["q_frames"], w["bits"][w["q_rows"]], w["keypoints"][w["q_rows"]],
                         w["landmark_xyz"], cams, num_threads=4)
    return ora, exp


def test_shipped_brisk_vocabulary_and_fixture_shapes():
    w = world()
    assert w["bits"].shape[1] == 48 and len(w["db_frames"]) == 128 and len(w["q_frames"]) == 126
    assert int(w["db_frames"]["num_descriptors"].sum()) + int(w["q_frames"]["num_descriptors"].sum()) == len(w["bits"])


def test_oracle_relocalises_the_real_query_vertices():
    w = world()
    ora0 = po.Engine(BLOB, po.default_settings())
    _, exp = oracle_side(w, ora0.project(w["bits"][w["db_rows"]]))
    acc = exp["accepted"].astype(bool)
    assert len(acc) == 63 and acc.sum() >= 60
    # recovered T_G_I against the pose the map itself stores for that vertex
    T, Tg = exp["T"][acc], w["T_G_I"][w["q_vertices"]][acc]
    err = np.linalg.norm(T[:, :, 3] - Tg[:, :, 3], axis=1)
    assert np.median(err) < 0.03 and err.max() < 0.15
    ang = [np.degrees(np.arccos(np.clip((np.trace(a[:, :3] @ b[:, :3].T) - 1) / 2, -1, 1))) for a, b in zip(T, Tg)]
    assert max(ang) < 1.0
    assert exp["num_inliers"][acc].min() >= 10


@pytest.mark.gpu
def test_device_path_equals_oracle_on_the_real_map():
    w = world()
    det = capi.Detector(BLOB, capi.default_settings())
    proj_db = det.project(w["bits"][w["db_rows"]])
    ora, exp = oracle_side(w, proj_db)
    assert np.array_equal(proj_db, ora.project(w["bits"][w["db_rows"]]))  # 384-bit BRISK projection, bit-exact
    det.insert_batch(w["db_frames"], proj_db, w["landmarks"][w["db_rows"]])
    det.set_landmark_positions(w["landmark_xyz"])
    qbits, qkp = w["bits"][w["q_rows"]], w["keypoints"][w["q_rows"]]
    qproj = det.project(qbits)
    k = det.num_neighbors()
    idx, dist = det.knn(qproj, k)
    oidx, odist = ora.knn(qproj, k)
    assert np.array_equal(idx, oidx) and np.array_equal(dist.view(np.uint32), odist.view(np.uint32))
    out = det.query_batch(w["q_frames"], qbits, qkp, capi.make_cameras(camera_dicts(w)), want_matches=True)
    res = out["results"]
    assert np.array_equal(res["accepted"], exp["accepted"]) and res["accepted"].sum() >= 60
    assert np.array_equal(np.diff(out["offsets"]), exp["num_matches"])
    assert np.array_equal(res["num_inliers"], exp["num_inliers"])
    assert np.array_equal(res["iterations"], exp["iterations"])
    ok = exp["ransac_success"].astype(bool)
    T, Te = res["T_G_I"].reshape(-1, 3, 4)[ok], exp["T"][ok]
    assert np.abs(T[:, :, 3] - Te[:, :, 3]).max() <= 1e-6                     # 1e-6 m
    for a, b in zip(T, Te):
        assert np.arccos(np.clip((np.trace(a[:, :3] @ b[:, :3].T) - 1) / 2, -1, 1)) <= 1e-6   # 1e-6 rad
    # precision of the structure matches handed to RANSAC (inliers and outliers): for most of them the
    # matched landmark is the query keypoint's own landmark (measured 0.74)
    q_landmark = w["landmarks"][w["q_rows"]]
    q_off = np.concatenate([[0], np.cumsum(w["q_frames"]["num_descriptors"])])
    m = out["matches"]
    same = q_landmark[q_off[m["query_frame"]] + m["query_keypoint"]] == m["landmark"]
    assert same.mean() > 0.6
