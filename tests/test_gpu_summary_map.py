"""addLocalizationSummaryMapToDatabase + localization against it (SURVEY 8f rank 2;
LCH/src/loop-detector-node.cc:341-432, :434-475): a summary map is assembled from the synthetic
map in the order createLocalizationSummaryMapFromLandmarkList produces it
(map-structure/localization-summary-map/src/localization-summary-map-creation.cc:63-204:
observations landmark-major, observers numbered by first appearance), written by the REAL protobuf
runtime, ingested by the CUDA path from the file bytes and by the oracle from libprotobuf's decoded
arrays; database, kNN results and accepted localizations must be identical."""
import numpy as np
import pytest

import summary_map_proto as smp
from maplab_b200 import capi, synthetic
from oracle import pyoracle as po
from helpers import fill_oracle, frames_of, small_world

pytestmark = pytest.mark.gpu


def summary_map_of(m, proj):
    """Arrays of the summary map over all landmarks of the synthetic map."""
    kf_of_desc = np.repeat(np.arange(len(m["frames"]["num_descriptors"])), m["frames"]["num_descriptors"])
    order = np.argsort(m["landmarks"], kind="stable")  # landmark-major, observation order kept
    seen, first = np.unique(kf_of_desc[order], return_index=True)
    rank = np.empty(len(seen), np.int64)
    rank[np.argsort(first, kind="stable")] = np.arange(len(seen))
    observer_of_kf = np.full(kf_of_desc.max() + 1, -1, np.int64)
    observer_of_kf[seen] = rank  # observer index = order of first appearance
    kf_of_observer = np.empty(len(seen), np.int64)
    kf_of_observer[rank] = seen
    return dict(G_landmark_position=m["landmark_xyz"].T.astype(np.float32),
                G_observer_position=m["kf_pos"][kf_of_observer].T.astype(np.float32),
                descriptors=np.ascontiguousarray(proj[order].T),
                observer_indices=observer_of_kf[kf_of_desc[order]].astype(np.uint32),
                observation_to_landmark_index=m["landmarks"][order].astype(np.uint32)), order


@pytest.mark.parametrize("packed", [False, True])
def test_summary_map_database_and_localization_match_oracle(packed):
    m, blob, _, q = small_world(num_queries=12)
    kw = dict(num_nearest_neighbors=6)
    det = capi.Detector(blob, capi.default_settings(**kw))
    ora = po.Engine(blob, po.default_settings(**kw))
    proj = det.project(m["bits"])
    arrays, order = summary_map_of(m, proj)
    file_bytes = smp.encode(packed=packed, **arrays)
    MISSION, V0, L0 = 77, 1000, 40
    sizes = det.add_summary_map(file_bytes, MISSION, V0, L0)
    assert sizes["num_observations"] == len(order) and sizes["num_landmarks"] == len(m["landmark_xyz"])
    assert ora.add_summary_map(smp.decode(file_bytes, packed), MISSION, V0, L0) == 0
    assert det.num_descriptors() == ora.num_descriptors() == len(order)
    assert det.num_entries() == ora.num_entries() == sizes["num_observers"]

    # the database holds the observations regrouped by observer: kNN indices / distances identical
    qproj = det.project(q["bits"])
    idx, dist = det.knn(qproj, 6)
    eidx, edist = ora.knn(qproj, 6)
    assert np.array_equal(idx, eidx) and np.array_equal(dist.view(np.uint32), edist.view(np.uint32))
    assert (idx >= 0).mean() > 0.5

    # localization: landmark positions are the summary map's floats cast to double, at ids L0 + l;
    # database timestamps are 0 and the mission id is the map's own, so no time filter applies
    cam = synthetic.camera_dict()
    xyz = np.full((L0 + len(m["landmark_xyz"]), 3), np.nan)
    xyz[L0:] = m["landmark_xyz"].astype(np.float32).astype(np.float64)
    qframes = frames_of(q["frames"])
    out = det.query_batch(qframes, q["bits"], q["keypoints"], capi.make_cameras([cam]), want_matches=True)
    exp = po.query_batch(ora, qframes, q["bits"], q["keypoints"], xyz,
                         [po.make_camera(cam["fu"], cam["fv"], cam["cu"], cam["cv"])])
    res = out["results"]
    assert np.array_equal(res["accepted"], exp["accepted"])
    assert np.array_equal(res["num_inliers"], exp["num_inliers"])
    assert np.array_equal(res["iterations"], exp["iterations"])
    assert np.array_equal(np.diff(out["offsets"]), exp["num_matches"])
    ok = exp["ransac_success"].astype(bool)
    assert np.array_equal(res["T_G_I"].reshape(-1, 3, 4)[ok], exp["T"][ok])
    acc = res["accepted"].astype(bool)
    assert acc.sum() >= len(acc) // 2
    T = res["T_G_I"].reshape(-1, 3, 4)[acc]
    assert np.abs(T[:, :, 3] - q["T_G_I"][acc][:, :, 3]).max() < 0.2
    mt = out["matches"]
    assert (mt["landmark"] >= L0).all() and (mt["db_vertex"] >= V0).all()


def test_summary_map_next_to_a_regular_mission_keeps_earlier_landmarks():
    m, blob, _, q = small_world(num_queries=8)
    det = capi.Detector(blob, capi.default_settings(num_nearest_neighbors=6))
    ora = po.Engine(blob, po.default_settings(num_nearest_neighbors=6))
    proj = det.project(m["bits"])
    frames = frames_of(m["frames"])
    half = len(frames) // 2
    n_half = int(frames["num_descriptors"][:half].sum())
    det.insert_batch(frames[:half], proj[:n_half], m["landmarks"][:n_half])
    det.set_landmark_positions(m["landmark_xyz"])
    fill_oracle(ora, frames[:half], proj[:n_half], m["landmarks"][:n_half])
    arrays, _ = summary_map_of(m, proj)
    L = len(m["landmark_xyz"])
    file_bytes = capi.summary_map_serialize(**arrays)
    assert file_bytes == smp.encode(**arrays)
    det.add_summary_map(file_bytes, 500, 10_000, L + 3)  # ids L .. L+2 stay without a position
    assert ora.add_summary_map(smp.decode(file_bytes), 500, 10_000, L + 3) == 0
    xyz = np.full((2 * L + 3, 3), np.nan)
    xyz[:L] = m["landmark_xyz"]
    xyz[L + 3:] = m["landmark_xyz"].astype(np.float32).astype(np.float64)
    cam = synthetic.camera_dict()
    qframes = frames_of(q["frames"])
    out = det.query_batch(qframes, q["bits"], q["keypoints"], capi.make_cameras([cam]))
    exp = po.query_batch(ora, qframes, q["bits"], q["keypoints"], xyz,
                         [po.make_camera(cam["fu"], cam["fv"], cam["cu"], cam["cv"])])
    assert np.array_equal(out["results"]["accepted"], exp["accepted"])
    assert np.array_equal(out["results"]["num_inliers"], exp["num_inliers"])
    ok = exp["ransac_success"].astype(bool)
    assert np.array_equal(out["results"]["T_G_I"].reshape(-1, 3, 4)[ok], exp["T"][ok])


def test_summary_map_rejections():
    m, blob, _, _ = small_world()
    det = capi.Detector(blob, capi.default_settings())
    good = dict(G_landmark_position=np.zeros((3, 4), np.float32), G_observer_position=np.zeros((3, 2), np.float32),
                descriptors=np.zeros((10, 3), np.float32), observer_indices=np.array([0, 1, 1], np.uint32),
                observation_to_landmark_index=np.array([0, 3, 2], np.uint32))
    ora = po.Engine(blob, po.default_settings())
    cases = [("observer index out of range", dict(observer_indices=np.array([0, 2, 1], np.uint32)), 2),
             ("landmark index out of range", dict(observation_to_landmark_index=np.array([0, 4, 2], np.uint32)), 5),
             ("fewer descriptors than observations", dict(descriptors=np.zeros((10, 2), np.float32)), 3),
             ("No observers", dict(G_observer_position=np.zeros((3, 0), np.float32),
                                   observer_indices=np.zeros(0, np.uint32),
                                   observation_to_landmark_index=np.zeros(0, np.uint32)), 1),
             ("dimensionality", dict(descriptors=np.zeros((8, 3), np.float32)), None)]
    for why, change, check in cases:
        arrays = dict(good, **change)
        with pytest.raises(capi.MlcError, match=why):
            det.add_summary_map(smp.encode(**arrays), 1, 0, 0)
        assert det.num_entries() == 0 and det.num_descriptors() == 0  # nothing half-inserted
        if check is not None:
            assert ora.add_summary_map(arrays, 1, 0, 0) == check  # the CHECK the reference would die on
    with pytest.raises(capi.MlcError, match="malformed"):
        det.add_summary_map(b"\x0a\x05abc", 1, 0, 0)
    assert det.add_summary_map(smp.encode(**good), 1, 0, 0)["num_observers"] == 2
    assert det.num_entries() == 2 and det.num_descriptors() == 3


def test_create_summary_map_equals_reference_construction():
    # createLocalizationSummaryMapFromLandmarkList: observations landmark-major, projected on the device,
    # observers numbered by first appearance, positions cast to float; the file bytes must be what
    # libprotobuf writes for the arrays of the numpy restatement (summary_map_of)
    m, blob, _, _ = small_world()
    det = capi.Detector(blob, capi.default_settings())
    proj = det.project(m["bits"])
    arrays, order = summary_map_of(m, proj)
    kf_of_desc = np.repeat(np.arange(len(m["frames"]["num_descriptors"])), m["frames"]["num_descriptors"])
    counts = np.bincount(m["landmarks"], minlength=len(m["landmark_xyz"]))
    file_bytes = det.create_summary_map(m["landmark_xyz"], counts, m["bits"][order], 1000 + 7 * kf_of_desc[order],
                                        m["kf_pos"][kf_of_desc[order]])
    assert file_bytes == smp.encode(**arrays)
    got = capi.summary_map_parse(file_bytes)
    assert np.array_equal(got["descriptors"], proj[order].T)
    # and it feeds the database like any other summary map
    det2 = capi.Detector(blob, capi.default_settings())
    sizes = det2.add_summary_map(file_bytes, 5, 0, 0)
    assert sizes["num_observations"] == len(order) and det2.num_descriptors() == len(order)
    with pytest.raises(capi.MlcError, match="do not add up"):
        det.create_summary_map(m["landmark_xyz"], counts[::-1] * 0 + 1, m["bits"][order], kf_of_desc[order],
                               m["kf_pos"][kf_of_desc[order]])


def test_reference_summary_map_creation_test():
    # map-structure/localization-summary-map/test/test_localization_summary_map_test.cc:28-86
    # (LocalizationSummaryCreationFromMapTest): landmark 1 at (0, 0, 1) stored in v1 and observed by v2, v3; landmark 2
    # at (0, 0, 2) stored in v2 and observed by v3; the summary map of {landmark 1, landmark 2} must hold 5
    # observations with observation_to_landmark_index = 0 0 0 1 1 and observer indices = 0 1 2 1 2.
    _, blob, _, _ = small_world()
    det = capi.Detector(blob, capi.default_settings())
    v1, v2, v3 = 101, 202, 303
    bits = np.random.default_rng(0).integers(0, 256, (5, 64), dtype=np.uint8)
    file_bytes = det.create_summary_map([[0, 0, 1], [0, 0, 2]], [3, 2], bits, [v1, v2, v3, v2, v3], np.zeros((5, 3)))
    got = capi.summary_map_parse(file_bytes)
    assert got["G_landmark_position"].shape == (3, 2)                              # ASSERT_EQ(cols, 2)
    assert got["G_landmark_position"][:, 0].tolist() == [0, 0, 1]
    assert got["G_landmark_position"][:, 1].tolist() == [0, 0, 2]
    assert got["descriptors"].shape[1] == 5                                        # projectedDescriptors().cols()
    assert got["observation_to_landmark_index"].tolist() == [0, 0, 0, 1, 1]
    assert got["observer_indices"].tolist() == [0, 1, 2, 1, 2]
    assert got["G_observer_position"].shape == (3, 3) and not got["G_observer_position"].any()
    assert np.array_equal(got["descriptors"], det.project(bits).T)
