"""maplab_b200/vi_map_io.py against the reference's own files (present in the build container only — the tests
skip on the GPU box, where the committed fixture stands in): the loader reproduces tests/golden/real_map_brisk.npz
from maplab's common_test_map, and projection_matrix_*.dat (common::Serialize) equal the matrices inside the
shipped quantizer files."""
import os

import numpy as np
import pytest

from maplab_b200 import synthetic, vi_map_io

MAP = "/root/reference/tools/maplab-test-data/test_maps/common_test_map/vi_map"
SHARE = "/root/reference/algorithms/loopclosure/matching-based-loopclosure/share"
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
needs_reference = pytest.mark.skipif(not os.path.isdir(MAP), reason="reference checkout not mounted")


@needs_reference
def test_loader_reproduces_the_committed_fixture():
    vi_map = vi_map_io.load_vi_map(MAP)
    assert len(vi_map["vertices"]) == 127 and len(vi_map["missions"]) == 1
    x = vi_map_io.loop_closure_inputs(vi_map, (0, 2))
    d = np.load(os.path.join(GOLDEN, "real_map_brisk.npz"))
    assert np.array_equal(x["frames"], d["frames"]) and np.array_equal(x["bits"], d["bits"])
    assert np.array_equal(x["keypoints"], d["keypoints"]) and np.array_equal(x["T_G_I"], d["T_G_I"])
    used = np.unique(x["landmarks"])
    assert np.array_equal(x["landmark_xyz"][used], d["landmark_xyz"])
    assert np.array_equal(np.searchsorted(used, x["landmarks"]), d["landmarks"])
    assert (np.diff(x["frames"][::2, 0]) > 0).all()  # pose-graph (time) order
    # every kept landmark lies in front of a camera that observes it, a few metres away
    cams = vi_map_io.cameras_of(vi_map["sensors"], (0, 2))
    assert [c["distortion"] for c in cams] == [3, 3] and abs(np.linalg.det(cams[0]["R_B_C"]) - 1) < 1e-9
    off = np.concatenate([[0], np.cumsum(x["frames"][:, 3])])
    f = 10
    T = np.vstack([x["T_G_I"][x["frames"][f, 1]], [0, 0, 0, 1]])
    c = cams[x["frames"][f, 2]]
    T_B_C = np.vstack([np.hstack([c["R_B_C"], c["t_B_C"][:, None]]), [0, 0, 0, 1]])
    p_G = x["landmark_xyz"][x["landmarks"][off[f]:off[f + 1]]]
    p_C = (np.linalg.inv(T @ T_B_C) @ np.hstack([p_G, np.ones((len(p_G), 1))]).T).T[:, :3]
    assert (p_C[:, 2] > 0).mean() > 0.95


@needs_reference
def test_all_cameras_and_all_keypoint_counts():
    x = vi_map_io.loop_closure_inputs(vi_map_io.load_vi_map(MAP))
    assert len(x["frames"]) == 127 * 5 and int(x["frames"][:, 3].sum()) == len(x["bits"]) == 101999
    assert x["bits"].shape[1] == 48 and x["landmarks"].max() + 1 == len(x["landmark_xyz"]) == 14304


@needs_reference
@pytest.mark.parametrize("name", ["freak", "brisk"])
def test_projection_matrix_files_equal_the_quantizer_matrices(name):
    P = vi_map_io.load_projection_matrix(os.path.join(SHARE, f"projection_matrix_{name}.dat"))
    v = synthetic.parse_vocabulary(open(os.path.join(SHARE, f"inverted_multi_index_quantizer_{name}.dat"), "rb").read())
    assert np.array_equal(P, v["P"])
    if name == "brisk":  # the committed trimmed quantizer = first target_dim rows
        t = synthetic.parse_vocabulary(open(os.path.join(GOLDEN, "brisk_quantizer_top10.dat"), "rb").read())
        assert np.array_equal(t["P"], P[:10]) and np.array_equal(t["W1"], v["W1"]) and np.array_equal(t["W2"], v["W2"])


def test_projection_matrix_reader_rejects_garbage(tmp_path):
    p = tmp_path / "m.dat"
    p.write_bytes(b"\x02\x00\x00\x00\x02\x00\x00\x00" + b"\x00" * 12)
    with pytest.raises(ValueError):
        vi_map_io.load_projection_matrix(str(p))
    p.write_bytes(np.array([2, 3], np.int32).tobytes() + np.arange(6, dtype=np.float32).tobytes())
    assert vi_map_io.load_projection_matrix(str(p)).tolist() == [[0, 2, 4], [1, 3, 5]]


# ---- the library's own reader of the vertices files (mlc_vi_map_count / mlc_vi_map_read, C++, no libprotobuf)
# against the protobuf runtime
def _compare_with_runtime(proto_bytes):
    from maplab_b200 import capi
    got = capi.vi_map_read_vertices(proto_bytes)
    msg = vi_map_io._vi_map_class()()
    msg.ParseFromString(proto_bytes)
    V = len(msg.vertices)
    assert got["vertex_id"].tolist() == [list(i.uint) for i in msg.vertex_ids]
    assert got["mission_id"].tolist() == [list(v.mission_id.uint) for v in msg.vertices]
    assert np.array_equal(got["T_M_I"], np.array([list(v.T_M_I) for v in msg.vertices]).reshape(V, 7))
    frames = [f for v in msg.vertices for f in v.n_visual_frame.frames]
    assert got["vertex_num_frames"].tolist() == [len(v.n_visual_frame.frames) for v in msg.vertices]
    assert got["frame_timestamp_ns"].tolist() == [f.timestamp for f in frames]
    assert got["frame_is_valid"].tolist() == [int(vi_map_io.frame_is_usable(f)) for f in frames]
    assert got["frame_num_keypoints"].tolist() == [len(f.keypoint_measurements) // 2 for f in frames]
    kp = np.concatenate([np.zeros((0, 2))] + [np.array(f.keypoint_measurements).reshape(-1, 2) for f in frames])
    assert np.array_equal(got["keypoint_measurement"], kp)
    desc = [vi_map_io.descriptors_of(f) for f in frames]
    width = got["keypoint_descriptor"].shape[1]
    assert np.array_equal(got["keypoint_descriptor"],
                          np.concatenate([np.zeros((0, width), np.uint8)] + [d for d in desc if d.size]))
    ids = [list(l.uint) + [0] * (2 - len(l.uint)) for f in frames for l in f.landmark_ids]
    assert got["keypoint_landmark_id"].tolist() == ids
    lms = [l for v in msg.vertices for l in v.landmark_store.landmarks]
    assert got["vertex_num_landmarks"].tolist() == [len(v.landmark_store.landmarks) for v in msg.vertices]
    assert got["landmark_id"].tolist() == [list(l.id.uint) for l in lms]
    assert np.array_equal(got["landmark_p_B"], np.array([list(l.position) for l in lms]).reshape(len(lms), 3))
    assert got["landmark_quality"].tolist() == [l.quality for l in lms]
    return got


def _toy_map(rng, vertices=3, frames=2, bytes_per_desc=48):
    import struct
    msg = vi_map_io._vi_map_class()()
    for v in range(vertices):
        msg.vertex_ids.add().uint.extend([int(rng.integers(1, 2 ** 63)), 2 ** 64 - 1 - v])
        vert = msg.vertices.add()
        vert.T_M_I.extend(rng.normal(size=7).tolist())
        vert.mission_id.uint.extend([7, 9])
        for f in range(frames):
            fr = vert.n_visual_frame.frames.add()
            n = int(rng.integers(0, 6)) if (v, f) != (0, 0) else 4
            fr.timestamp = 1_600_000_000_000_000_000 + 1000 * v + f
            # frame validity as the reference reads it: (v + f) % 4 == 0: is_valid absent (= valid), 1: present and
            # true, 2: present and false (invalidated), 3: no frame id (the frame "has been un-set")
            if (v + f) % 4 != 3:
                fr.id.uint.extend([int(rng.integers(1, 2 ** 62)), 77])
            if (v + f) % 4 in (1, 2):
                fr.is_valid = (v + f) % 4 == 1
            fr.keypoint_measurements.extend(rng.uniform(0, 700, 2 * n).tolist())
            data = rng.integers(0, 256, (n, bytes_per_desc), dtype=np.uint8)
            fr.keypoint_descriptors = struct.pack("<qiiii", 1, bytes_per_desc, n, 0, 1) + data.tobytes()
            for i in range(n):
                lid = fr.landmark_ids.add()
                if i % 3:
                    lid.uint.extend([int(rng.integers(1, 2 ** 62)), int(rng.integers(1, 2 ** 62))])
                else:
                    lid.uint.extend([0, 0])  # invalid id
        for _ in range(int(rng.integers(0, 4))):
            lm = vert.landmark_store.landmarks.add()
            lm.id.uint.extend([int(rng.integers(1, 2 ** 62)), 5])
            lm.position.extend(rng.normal(size=3).tolist())
            lm.quality = int(rng.integers(0, 3))
    return msg


def test_cxx_reader_equals_protobuf_runtime_on_constructed_maps():
    rng = np.random.default_rng(0)
    for bytes_per_desc in (48, 64):
        got = _compare_with_runtime(_toy_map(rng, bytes_per_desc=bytes_per_desc).SerializeToString())
        assert got["keypoint_descriptor"].shape[1] == bytes_per_desc
    _compare_with_runtime(b"")  # an empty message is an empty map


def test_unset_and_invalidated_frames_stay_out_of_the_detector_inputs():
    """addVertexToDatabase / queryVertexInDatabase only take frames with isVisualFrameSet && isVisualFrameValid
    (loop-detector-node.cc:279-280, :694-695); both readers must drop the others and agree."""
    from maplab_b200 import capi
    rng = np.random.default_rng(5)
    msg = _toy_map(rng, vertices=6, frames=2)
    for v in msg.vertices:  # make every keypoint's landmark a good landmark of its vertex
        del v.landmark_store.landmarks[:]
        for fr in v.n_visual_frame.frames:
            for lid in fr.landmark_ids:
                if any(lid.uint):
                    lm = v.landmark_store.landmarks.add()
                    lm.id.uint.extend(list(lid.uint))
                    lm.position.extend(rng.normal(size=3).tolist())
                    lm.quality = 2
    missions = {(7, 9): np.eye(4)}
    ids = [tuple(i.uint) for i in msg.vertex_ids]
    exp = vi_map_io.loop_closure_inputs(dict(vertices=list(zip(ids, msg.vertices)), missions=missions, sensors=None))
    usable = [vi_map_io.frame_is_usable(f) for v in msg.vertices for f in v.n_visual_frame.frames]
    assert 0 < sum(usable) < len(usable) and len(exp["frames"]) == sum(usable)
    got = vi_map_io.loop_closure_inputs_native(capi.vi_map_read_vertices(msg.SerializeToString()), missions)
    for key in ("frames", "missions", "bits", "keypoints", "landmarks", "landmark_xyz", "T_G_I"):
        assert np.array_equal(got[key], exp[key]), key
    assert len(exp["bits"]) > 0


def test_cxx_reader_rejections():
    import gzip
    from maplab_b200 import capi
    rng = np.random.default_rng(1)
    msg = _toy_map(rng)
    good = msg.SerializeToString()
    with pytest.raises(capi.MlcError, match="gzip"):
        capi.vi_map_read_vertices(gzip.compress(good))
    with pytest.raises(capi.MlcError):
        capi.vi_map_read_vertices(good[:len(good) // 2])
    broken = vi_map_io._vi_map_class()()
    broken.CopyFrom(msg)
    broken.vertices[0].n_visual_frame.frames[0].landmark_ids.add().uint.extend([1, 2])
    with pytest.raises(capi.MlcError, match="differ in number"):
        capi.vi_map_read_vertices(broken.SerializeToString())
    broken.CopyFrom(msg)
    del broken.vertex_ids[-1]
    with pytest.raises(capi.MlcError, match="vertex_ids and vertices"):
        capi.vi_map_read_vertices(broken.SerializeToString())
    broken.CopyFrom(msg)
    fr0 = broken.vertices[0].n_visual_frame.frames[0]
    fr0.keypoint_descriptors = (2).to_bytes(8, "little") + fr0.keypoint_descriptors[8:]
    with pytest.raises(capi.MlcError, match="several descriptor types"):
        capi.vi_map_read_vertices(broken.SerializeToString())
    broken.CopyFrom(msg)
    broken.vertices[0].n_visual_frame.frames[0].keypoint_descriptors = b"\x00" * 30
    with pytest.raises(capi.MlcError, match="descriptor matrix"):
        capi.vi_map_read_vertices(broken.SerializeToString())


@needs_reference
def test_cxx_reader_equals_protobuf_runtime_on_the_reference_map():
    for name in ("vertices0", "vertices2"):
        got = _compare_with_runtime(vi_map_io.read_proto_bytes(os.path.join(MAP, name)))
        assert got["keypoint_descriptor"].shape[1] == 48 and (got["landmark_quality"] == 2).all()
        assert (got["vertex_num_frames"] == 5).all()


@needs_reference
def test_oracle_relocalises_every_query_vertex_with_the_full_rig():
    # map folder -> vi_map_io -> oracle, all five cameras (the committed fixture keeps two): even vertices are the
    # database mission, odd vertices the query mission; every query vertex is relocalised within 6 cm of the pose
    # the bundle-adjusted reference map stores for it (median ~1.3 cm)
    from maplab_b200 import capi
    from oracle import pyoracle as po
    blob = open(os.path.join(GOLDEN, "brisk_quantizer_top10.dat"), "rb").read()
    vm = vi_map_io.load_vi_map(MAP)
    x = vi_map_io.loop_closure_inputs(vm)
    fr = x["frames"]
    off = np.concatenate([[0], np.cumsum(fr[:, 3])])
    is_db = fr[:, 1] % 2 == 0
    ora = po.Engine(blob, po.default_settings())
    proj = ora.project(x["bits"])
    for i in np.nonzero(is_db)[0]:
        s, e = off[i], off[i + 1]
        ora.insert(int(fr[i, 0]), int(fr[i, 1]), int(fr[i, 2]), 0, proj[s:e], x["landmarks"][s:e])
    cams = [po.make_camera(c["fu"], c["fv"], c["cu"], c["cv"], c["R_B_C"], c["t_B_C"], c["distortion"], c["dist"])
            for c in vi_map_io.cameras_of(vm["sensors"])]
    q = fr[~is_db]
    rows = np.concatenate([np.arange(off[i], off[i + 1]) for i in np.nonzero(~is_db)[0]])
    frames = capi.make_frames(q[:, 0], q[:, 1], np.ones(len(q), np.int64), q[:, 2], q[:, 3])
    exp = po.query_batch(ora, frames, x["bits"][rows], x["keypoints"][rows], x["landmark_xyz"], cams, num_threads=4)
    acc = exp["accepted"].astype(bool)
    assert len(acc) == 63 and acc.all()
    err = np.linalg.norm(exp["T"][:, :, 3] - x["T_G_I"][q[::5, 1]][:, :, 3], axis=1)
    assert np.median(err) < 0.02 and err.max() < 0.08
    assert exp["num_inliers"].min() >= 20


SUBMAPS = "/root/reference/tools/maplab-test-data/test_maps/submap_test"


@pytest.mark.skipif(not os.path.isdir(SUBMAPS), reason="reference checkout not mounted")
def test_oracle_closes_loops_across_the_reference_submaps():
    # seven consecutive ~30 s submaps of one real stereo (radial-tangential cameras) session in a common frame:
    # all of them in one database, each its own mission; every submap queried against it (the time filter only
    # applies inside a mission). Closures into the neighbouring submaps agree with the stored poses to centimetres;
    # the last submap closes the loop back to the first ones and exposes the odometry drift (0.5-0.8 m), on which
    # the mission alignment (transformationRansac on T_G_M samples) finds a consensus.
    from maplab_b200 import capi
    from oracle import pyoracle as po
    blob = open(os.path.join(GOLDEN, "brisk_quantizer_top10.dat"), "rb").read()
    maps = []
    for i in range(7):
        vm = vi_map_io.load_vi_map(os.path.join(SUBMAPS, f"submap_{i}", "vi_map"))
        maps.append((vm, vi_map_io.loop_closure_inputs(vm)))
    cam_dicts = vi_map_io.cameras_of(maps[0][0]["sensors"])
    assert [c["distortion"] for c in cam_dicts] == [2, 2]
    cams = [po.make_camera(c["fu"], c["fv"], c["cu"], c["cv"], c["R_B_C"], c["t_B_C"], c["distortion"], c["dist"])
            for c in cam_dicts]
    ora = po.Engine(blob, po.default_settings())
    xyz, lm_off = [], 0
    for i, (_, x) in enumerate(maps):
        proj, fr = ora.project(x["bits"]), x["frames"]
        off = np.concatenate([[0], np.cumsum(fr[:, 3])])
        for f in range(len(fr)):
            s, e = off[f], off[f + 1]
            ora.insert(int(fr[f, 0]), 1000 * i + int(fr[f, 1]), int(fr[f, 2]), i, proj[s:e], x["landmarks"][s:e] + lm_off)
        xyz.append(x["landmark_xyz"])
        lm_off += len(x["landmark_xyz"])
    xyz = np.concatenate(xyz)
    errors = []
    for i, (_, x) in enumerate(maps):
        fr = x["frames"]
        frames = capi.make_frames(fr[:, 0], 1000 * i + fr[:, 1], np.full(len(fr), i, np.int64), fr[:, 2], fr[:, 3])
        exp = po.query_batch(ora, frames, x["bits"], x["keypoints"], xyz, cams, num_threads=4)
        acc = exp["accepted"].astype(bool)
        assert acc.sum() >= 2, i
        errors.append(np.linalg.norm(exp["T"][acc][:, :, 3] - x["T_G_I"][acc][:, :, 3], axis=1))
        if i == 6:
            T_lc, T_map = exp["T"][acc], x["T_G_I"][acc]
    for i in (2, 3, 4, 5):  # middle of the session: only the neighbouring submaps are in view
        assert errors[i].max() < 0.1
    drift = errors[6][errors[6] > 0.3]
    assert len(drift) >= 3 and drift.max() < 1.0
    # T_G_M samples of the last submap: T_G_I(loop closure) * T_G_I(stored)^-1 -> consensus of the drifted ones
    quats, poss = [], []
    for a, b in zip(T_lc, T_map):
        R = a[:, :3] @ b[:, :3].T
        quats.append(_quat(R))
        poss.append(a[:, 3] - R @ b[:, 3])
    q, p, inl = po.transformation_ransac(np.array(quats), np.array(poss), 2000, 0.174, 2.0, 42)
    assert len(inl) >= 3 and 0.3 < np.linalg.norm(p) < 1.5


def _quat(R):
    w = np.sqrt(max(0.0, 1 + np.trace(R))) / 2
    return np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1], 4 * w * w]) / (4 * w)


@needs_reference
def test_native_reader_feeds_the_same_inputs_as_the_protobuf_runtime():
    # map folder -> C++ reader -> detector inputs == map folder -> protobuf runtime -> detector inputs
    vm = vi_map_io.load_vi_map(MAP)
    exp = vi_map_io.loop_closure_inputs(vm, (0, 2))
    missions = vi_map_io.load_missions_native(MAP)
    assert missions.keys() == vm["missions"].keys()
    assert all(np.array_equal(missions[k], vm["missions"][k]) for k in missions)
    got = vi_map_io.loop_closure_inputs_native(vi_map_io.load_vertices_native(MAP), missions, (0, 2))
    for key in ("frames", "missions", "bits", "keypoints", "landmarks", "landmark_xyz", "T_G_I"):
        assert np.array_equal(got[key], exp[key]), key
    assert got["vertex_ids"] == exp["vertex_ids"]
    d = np.load(os.path.join(GOLDEN, "real_map_brisk.npz"))
    assert np.array_equal(got["bits"], d["bits"]) and np.array_equal(got["keypoints"], d["keypoints"])


def test_cxx_missions_reader_on_a_constructed_message():
    from maplab_b200 import capi
    msg = vi_map_io._vi_map_class()()
    rng = np.random.default_rng(2)
    base_ids = [[11, 12], [21, 22], [31, 32]]
    Ts = rng.normal(size=(3, 7))
    for b, T in zip(base_ids, Ts):
        msg.mission_base_frame_ids.add().uint.extend(b)
        msg.mission_base_frames.add().T_G_M.extend(T.tolist())
    for m, b in (([1, 2], 2), ([3, 4], 0)):  # missions point at base frames out of order
        msg.mission_ids.add().uint.extend(m)
        msg.missions.add().baseframe_id.uint.extend(base_ids[b])
    ids, T = capi.vi_map_read_missions(msg.SerializeToString())
    assert ids.tolist() == [[1, 2], [3, 4]] and np.array_equal(T, Ts[[2, 0]])
    msg.missions[1].baseframe_id.uint[0] = 99
    with pytest.raises(capi.MlcError, match="without base frame"):
        capi.vi_map_read_missions(msg.SerializeToString())
    ids, T = capi.vi_map_read_missions(b"")
    assert len(ids) == 0
