"""Parity of kernel 3 (match gathering + time filter, votes, top-fraction selection, covisibility
components, uniqueness) with the oracle's LoopDetector::Find: the accepted match lists must be
identical (integer outputs, bit-exact), per query vertex, in canonical order."""
import numpy as np
import pytest

from maplab_b200 import capi, synthetic
from maplab_b200.loop_detector import LoopDetector, ProjectedImage
from oracle import pyoracle as po
from helpers import fill_oracle, frames_of, small_world

pytestmark = pytest.mark.gpu


def _pair(blob, m, **kw):
    okw = dict(kw)
    det = capi.Detector(blob, capi.default_settings(**kw))
    ora = po.Engine(blob, po.default_settings(**okw))
    proj = det.project(m["bits"])
    frames = frames_of(m["frames"])
    det.insert_batch(frames, proj, m["landmarks"])
    fill_oracle(ora, frames, proj, m["landmarks"])
    return det, ora


def _oracle_find(ora, frames, proj_list):
    """frames: FRAME_DTYPE rows of ONE vertex."""
    return ora.find(int(frames["vertex_id"][0]), int(frames["mission_id"][0]),
                    [(int(f["timestamp_ns"]), int(f["frame_index"]), p) for f, p in zip(frames, proj_list)])


def _as_rows(matches, frames):
    return [(int(frames["frame_index"][r["query_frame"]]), int(r["query_keypoint"]), int(r["db_descriptor"]),
             int(r["db_keyframe"]), int(r["db_vertex"]), int(r["landmark"])) for r in matches]


def _check_batch(det, ora, qframes, qproj):
    matches, offsets = det.find_batch(qframes, proj=qproj)
    # split the batch by vertex like the shim does
    starts = np.concatenate([[0], np.cumsum(qframes["num_descriptors"])])
    groups, f = [], 0
    while f < len(qframes):
        g = [f]
        while f + 1 < len(qframes) and qframes["vertex_id"][f + 1] == qframes["vertex_id"][f]:
            f += 1
            g.append(f)
        groups.append(g)
        f += 1
    assert len(offsets) == len(groups) + 1
    total = 0
    for gi, g in enumerate(groups):
        exp = _oracle_find(ora, qframes[g], [qproj[starts[i]:starts[i + 1]] for i in g])
        got = _as_rows(matches[offsets[gi]:offsets[gi + 1]], qframes)
        assert got == [tuple(int(x) for x in r) for r in exp], f"vertex group {gi}"
        total += len(exp)
    return total


@pytest.mark.parametrize("kw", [
    dict(),                                                # reference defaults (auto k)
    dict(num_nearest_neighbors=6),
    dict(num_nearest_neighbors=8, min_verify_matches_num=3, fraction_best_scores=0.5),
    dict(num_nearest_neighbors=4, min_verify_matches_num=40),
    dict(num_nearest_neighbors=10, fraction_best_scores=0.01),  # floor(size*fraction) < 4 -> 4
    dict(scoring=1),                                       # probabilistic score (scoring.h:92-187)
    dict(scoring=1, num_nearest_neighbors=8, fraction_best_scores=0.5, min_verify_matches_num=3),
    dict(scoring=1, num_nearest_neighbors=10, fraction_best_scores=0.1),
])
def test_find_matches_oracle_single_camera(kw):
    m, blob, _, q = small_world(num_queries=12)
    det, ora = _pair(blob, m, **kw)
    qframes = frames_of(q["frames"])
    qproj = det.project(q["bits"])
    total = _check_batch(det, ora, qframes, qproj)
    if kw.get("min_verify_matches_num", 10) < 40:
        assert total > 0, "test world should produce loop-closure matches"


def test_time_and_mission_filter():
    # same mission as the database and timestamps close to the revisited keyframes: neighbours
    # closer than lc_min_image_time_seconds are dropped, the others kept
    m, blob, _, q = small_world(num_queries=12)
    det, ora = _pair(blob, m, num_nearest_neighbors=6, min_image_time_seconds=3.0)
    qframes = frames_of(q["frames"])
    qframes["mission_id"] = 0
    qframes["timestamp_ns"] = (q["revisit"].astype(np.int64) + 2) * 1_000_000_000 + 123
    qproj = det.project(q["bits"])
    _check_batch(det, ora, qframes, qproj)
    raw = ora.find_frame_trace(int(qframes["timestamp_ns"][0]), int(qframes["vertex_id"][0]), 0, 0,
                               qproj[:500])
    assert 0 < len(raw["raw"]) < (raw["knn_idx"] >= 0).sum()  # the filter dropped some, kept some


def test_multi_camera_vertex_second_pass():
    m, blob, _, q = small_world(num_queries=12)
    det, ora = _pair(blob, m, num_nearest_neighbors=6, min_verify_matches_num=4)
    qframes = frames_of(q["frames"])
    # pair up frames: (0,1) one vertex, (2,3) next ... camera index = frame_index
    qframes["vertex_id"] = 5000 + np.arange(len(qframes)) // 2
    qframes["frame_index"] = np.arange(len(qframes)) % 2
    qproj = det.project(q["bits"])
    total = _check_batch(det, ora, qframes, qproj)
    assert total > 0
    # mixed batch: single-camera vertices and a three-camera vertex
    qframes["vertex_id"] = [1, 2, 2, 2, 3, 4, 4, 5, 6, 7, 8, 8]
    qframes["frame_index"] = [0, 0, 1, 2, 0, 0, 1, 0, 0, 0, 0, 1]
    _check_batch(det, ora, qframes, qproj)


def test_reference_api_mirror_and_preconditions():
    m, blob, _, q = small_world(num_queries=12)
    ld = LoopDetector(blob, capi.default_settings(num_nearest_neighbors=6))
    ora = po.Engine(blob, po.default_settings(num_nearest_neighbors=6))
    assert ld.Find([]).shape == (0,)  # empty list -> empty result (matching-based-engine.cc:52-56)
    frames = frames_of(m["frames"])
    proj = ld.ProjectDescriptors(m["bits"])
    at = 0
    for f in frames:
        n = int(f["num_descriptors"])
        ld.Insert(ProjectedImage(int(f["timestamp_ns"]), int(f["vertex_id"]), int(f["frame_index"]),
                                 int(f["mission_id"]), proj[at:at + n], m["landmarks"][at:at + n]))
        at += n
    fill_oracle(ora, frames, proj, m["landmarks"])
    assert ld.NumEntries() == len(frames) and ld.NumDescriptors() == len(proj)
    qf = q["frames"]
    qp = ld.ProjectDescriptors(q["bits"][:500])
    img = ProjectedImage(int(qf["timestamp_ns"][0]), int(qf["vertex_id"][0]), 0, int(qf["mission_id"][0]), qp)
    got = ld.Find([img])
    exp = ora.find(img.vertex_id, img.mission_id, [(img.timestamp_nanoseconds, 0, qp)])
    assert [tuple(int(x) for x in r) for r in exp] == _as_rows(got, frames_of(q["frames"])[:1])
    other = ProjectedImage(0, 99, 1, 0, qp)
    with pytest.raises(capi.MlcError):  # CHECK: all images share one vertex (engine.cc:57)
        ld.Find([img, other])
    with pytest.raises(capi.MlcError):  # CHECK_EQ(descriptors.cols(), landmarks.size())
        ld.Insert(ProjectedImage(0, 1, 0, 0, qp, np.arange(3)))
    ld.Clear()
    assert ld.NumEntries() == 0 and len(ld.Find([img])) == 0


def _any_detector():
    m, blob, _, _ = small_world(num_queries=12)
    return capi.Detector(blob, capi.default_settings())


def test_reference_golden_scoring_on_device():
    # test_scoring.cc:94-140 through the device scoring functions (mlc_score)
    import json
    import os
    t = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_goldens.json")))["scoring"]
    det = _any_detector()
    acc = det.score(t["num_matches"], t["num_descriptors"], t["num_db"], False)
    assert acc.tolist() == t["expected_accumulation"]
    prob = det.score(t["num_matches"], t["num_descriptors"], t["num_db"], True)
    assert np.allclose(prob, t["expected_probabilistic"], atol=t["tolerance"])
    assert len(det.score([], [], 50, True)) == 0
    assert len(det.score([1], [1], 0, True)) == 0  # empty database: no scores (scoring.h:108-112)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_probabilistic_score_matches_oracle(seed):
    # random vote tables incl. pdf underflow (FLT_MAX / the in-loop +inf patch, quirk 7):
    # zeros, FLT_MAX and +inf at identical positions, the rest within 1e-5 relative
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 400))
    votes = rng.integers(1, 60, n).astype(np.uint64)
    num_desc = rng.integers(50, 600, n).astype(np.uint64)
    for j in rng.integers(0, n, 6):  # underflowing ids with rising / falling / equal vote counts
        votes[j] = int(rng.choice([900, 1200, 1200, 1500, 2500]))
    num_db = int(num_desc.sum()) * [1, 3, 50, 2000][seed]
    det = _any_detector()
    got = det.score(votes, num_desc, num_db, True)
    exp = po.score(votes.tolist(), num_desc.tolist(), num_db, True)
    special = (exp == 0) | np.isinf(exp) | (exp == np.finfo(np.float32).max)
    assert np.array_equal(got[special], exp[special])
    assert np.array_equal(special, (got == 0) | np.isinf(got) | (got == np.finfo(np.float32).max))
    assert np.allclose(got[~special], exp[~special], rtol=1e-5, atol=0)
    if seed == 0:
        assert np.isinf(exp).any() or (exp == np.finfo(np.float32).max).any(), "underflow branch not exercised"


def test_find_matches_oracle_imipq_engine():
    # --lc_detector_engine=imipq (+ probabilistic scoring): the whole Find on PQ distances
    m, _, voc, q = small_world(num_queries=12)
    blob = synthetic.add_product_quantizer(voc, 10, 16)
    for kw in (dict(engine=1, num_nearest_neighbors=6), dict(engine=1, scoring=1, num_nearest_neighbors=8)):
        det, ora = _pair(blob, m, **kw)
        qframes = frames_of(q["frames"])
        qproj = det.project(q["bits"])
        assert _check_batch(det, ora, qframes, qproj) > 0
