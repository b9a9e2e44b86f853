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
