"""The `hnsw` engine slot (--lc_detector_engine hnsw, loop_closure::HSNWIndexInterface): float descriptors,
squared L2. The reference answers approximately with hnswlib; the B200 engine searches exhaustively.
Anchors: (1) tests/golden/hnsw_reference.npz = neighbour lists of the REAL vendored hnswlib (compiled from the
reference checkout into oracle/_ref, tests/golden/make_hnsw_golden.py) with the reference's parameters; (2) the
oracle's exact search (oracle/exact_knn.cc). The exact lists must explain the reference's: same order
convention (farthest first), same distances for the neighbours both found (1e-5 relative), never a worse k-th
neighbour, and the reference must recall most of them. On the GPU the engine equals the oracle bit for bit."""
import os

import numpy as np
import pytest

from maplab_b200 import capi
from oracle import pyoracle as po

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "hnsw_reference.npz"))
REF_SO = os.path.join(os.path.dirname(os.path.dirname(__file__)), "oracle", "_ref", "libhnsw_ref.so")


def _check_against_reference(idx, dist):
    ridx, rdist = G["idx"], G["dist"]
    k = idx.shape[1]
    assert (np.diff(dist, axis=1) <= 0).all() and (np.diff(rdist, axis=1) <= 0).all()   # farthest first, both
    recall = np.mean([len(set(idx[i]) & set(ridx[i])) / k for i in range(len(idx))])
    assert recall > 0.9, recall
    for i in range(len(idx)):
        exact = dict(zip(idx[i].tolist(), dist[i].tolist()))
        for j, d in zip(ridx[i].tolist(), rdist[i].tolist()):
            if j in exact:
                assert abs(exact[j] - d) <= 1e-5 * max(d, 1e-6)    # same distance for the same neighbour
        assert dist[i, 0] <= rdist[i, 0] * (1 + 1e-5)                # the exact k-th neighbour is never farther
    return recall


def test_exact_oracle_explains_the_reference_lists():
    idx, dist = po.exact_knn(G["db"], G["q"], 8)
    _check_against_reference(idx, dist)
    with pytest.raises(ValueError):
        po.exact_knn(G["db"][:5], G["q"][:1], 8)


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref not built (needs the reference checkout)")
def test_golden_is_what_the_compiled_reference_returns():
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "make_hnsw_golden", os.path.join(os.path.dirname(__file__), "golden", "make_hnsw_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    db, q = mod.world()
    assert np.array_equal(db, G["db"]) and np.array_equal(q, G["q"])
    idx, dist = mod.reference_knn(db, q, 8)
    assert np.array_equal(idx, G["idx"]) and np.array_equal(dist, G["dist"])


def _detector(dim, **kw):
    return capi.Detector(None, capi.default_settings(engine=2, float_descriptor_dim=dim, **kw))


@pytest.mark.gpu
def test_device_engine_equals_oracle_and_explains_the_reference():
    db, q = G["db"], G["q"]
    det = _detector(db.shape[1])
    # descriptors arrive as bytes and are "projected" by reinterpretation (hnsw-index-interface.h:155-163)
    proj = det.project(db.view(np.uint8).reshape(len(db), -1))
    assert np.array_equal(proj, db)
    half = 2500
    det.insert(0, 1, 0, 0, proj[:half], np.arange(half))
    det.insert(1, 2, 0, 0, proj[half:], np.arange(half, len(db)))
    assert det.num_descriptors() == len(db)
    for k in (1, 8, 16):
        idx, dist = det.knn(q, k)
        oidx, odist = po.exact_knn(db, q, k)
        assert np.array_equal(idx, oidx) and np.array_equal(dist, odist)
    idx, dist = det.knn(q, 8)
    _check_against_reference(idx, dist)
    with pytest.raises(capi.MlcError):      # CHECK_LT(num_neighbors, ef_query)
        _detector(8, hnsw_ef_query=4).knn(np.zeros((1, 8), np.float32), 4)
    small = _detector(db.shape[1])
    small.insert(0, 1, 0, 0, db[:3], np.arange(3))
    with pytest.raises(capi.MlcError):      # fewer descriptors than neighbours: the reference's CHECK_EQ aborts
        small.knn(q[:2], 8)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(70_000, 4100, 48, 6), (1500, 3, 7, 4), (200_000, 64, 128, 10)])
def test_device_engine_shapes(shape):
    """Many queries (no database split), a handful (16 splits + merge), odd dimensions, ties."""
    n_db, n_q, dim, k = shape
    rng = np.random.default_rng(n_db)
    db = rng.standard_normal((n_db, dim)).astype(np.float32)
    db[100:110] = db[50:60]                      # duplicate descriptors: distance ties, broken by index
    q = np.concatenate([db[50:60], rng.standard_normal((n_q, dim)).astype(np.float32)])[:max(n_q, 10)]
    det = _detector(dim)
    det.insert(0, 1, 0, 0, db, np.arange(n_db))
    idx, dist = det.knn(q, k)
    sample = np.arange(len(q))[:: max(len(q) // 200, 1)]
    oidx, odist = po.exact_knn(db, q[sample], k)
    assert np.array_equal(idx[sample], oidx) and np.array_equal(dist[sample], odist)
    assert (dist[:10, -1] == 0).all() and (dist[:10, -2] == 0).all()   # the duplicate pair, nearest LAST


@pytest.mark.gpu
def test_find_and_verify_on_float_descriptors():
    """The rest of the path on top of the engine: Find (voting / covisibility) and the fused query with RANSAC,
    against the oracle fed with the exact neighbour lists the same way."""
    from maplab_b200 import synthetic
    from helpers import frames_of, small_world
    m, blob, _, q = small_world(num_queries=6)
    # learned-descriptor stand-in: the projected binary descriptors ARE the float descriptors of this map
    ref_det = capi.Detector(blob, capi.default_settings(num_nearest_neighbors=6))
    db = ref_det.project(m["bits"])
    qd = ref_det.project(q["bits"])
    det = _detector(db.shape[1], num_nearest_neighbors=6)
    frames = frames_of(m["frames"])
    det.insert_batch(frames, db, m["landmarks"])
    det.set_landmark_positions(m["landmark_xyz"])
    qframes = frames_of(q["frames"])
    out = det.query_batch(qframes, qd.view(np.uint8).reshape(len(qd), -1), q["keypoints"],
                          capi.make_cameras([synthetic.camera_dict()]), want_matches=True)
    res = out["results"]
    assert res["accepted"].sum() >= 5
    T = res["T_G_I"].reshape(-1, 3, 4)[res["accepted"].astype(bool)]
    assert np.abs(T[:, :, 3] - q["T_G_I"][res["accepted"].astype(bool)][:, :, 3]).max() < 0.2
    # the matches come from the exact neighbour lists
    idx, _ = det.knn(qd, 6)
    nd = np.concatenate([[0], np.cumsum(qframes["num_descriptors"])])
    for mt in out["matches"][:200]:
        row = nd[mt["query_frame"]] + mt["query_keypoint"]
        assert mt["db_descriptor"] in idx[row]
