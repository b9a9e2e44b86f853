"""Shard-aware database build: a shard is handed only the descriptors it owns (descriptor i lives on
shard i % G) — from host rows (mlc_insert_batch_owned), from device rows (mlc_insert_batch_device) or
filtered out of all rows (mlc_insert_batch) — and answers exactly alike; Insert's uniqueness CHECK
(matching-based-engine.cc:244-252)."""
import numpy as np
import pytest

from maplab_b200 import capi, synthetic
from oracle import pyoracle as po
from helpers import fill_oracle, frames_of, small_world

pytestmark = pytest.mark.gpu


def test_owned_row_inserts_equal_filtered_inserts():
    import torch
    m, blob, _, q = small_world()
    frames = frames_of(m["frames"])
    full = capi.Detector(blob)
    proj = full.project(m["bits"])
    qp = full.project(q["bits"])
    G, k = 3, 6
    n = len(proj)
    for r in range(G):
        s = capi.default_settings(shard_rank=r, shard_count=G)
        a = capi.Detector(blob, s)
        a.insert_batch(frames, proj, m["landmarks"])
        ia, da = a.knn(qp, k)
        # host rows of the owned descriptors only, in two batches (the second starts at an arbitrary offset)
        b = capi.Detector(blob, s)
        half = len(frames) // 2
        nd = int(frames["num_descriptors"][:half].sum())
        own = np.arange(n) % G == r
        assert b.num_owned_in_range(0, nd) == own[:nd].sum()
        b.insert_batch_owned(frames[:half], proj[:nd][own[:nd]], m["landmarks"][:nd])
        assert b.num_owned_in_range(nd, n - nd) == own[nd:].sum()
        b.insert_batch_owned(frames[half:], proj[nd:][own[nd:]], m["landmarks"][nd:])
        ib, db = b.knn(qp, k)
        assert np.array_equal(ia, ib) and np.array_equal(da, db)
        # device rows
        c = capi.Detector(blob, s)
        p_d = torch.from_numpy(proj[:nd][own[:nd]]).cuda()
        l_d = torch.from_numpy(m["landmarks"][:nd]).cuda()
        c.insert_batch_device(frames[:half], p_d.data_ptr(), len(p_d), l_d.data_ptr())
        p_d2 = torch.from_numpy(proj[nd:][own[nd:]]).cuda()
        l_d2 = torch.from_numpy(m["landmarks"][nd:]).cuda()
        c.insert_batch_device(frames[half:], p_d2.data_ptr(), len(p_d2), l_d2.data_ptr())
        ic, dc = c.knn(qp, k)
        assert np.array_equal(ia, ic) and np.array_equal(da, dc)
        assert c.num_descriptors() == n and c.num_entries() == len(frames)
        with pytest.raises(capi.MlcError):  # wrong number of owned rows
            c.insert_batch_device(capi.make_frames([1], [10**9], [0], [0], [7]), p_d.data_ptr(), 7, 0)


def test_device_insert_single_shard_full_query_equals_host_insert():
    import torch
    m, blob, _, q = small_world(num_queries=8)
    frames = frames_of(m["frames"])
    a = capi.Detector(blob, capi.default_settings(num_nearest_neighbors=6))
    proj = a.project(m["bits"])
    a.insert_batch(frames, proj, m["landmarks"])
    a.set_landmark_positions(m["landmark_xyz"])
    b = capi.Detector(blob, capi.default_settings(num_nearest_neighbors=6))
    bits_d = torch.from_numpy(m["bits"]).cuda()
    proj_d = torch.empty((len(proj), b.dim), dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    b.project_device(bits_d.data_ptr(), 64, len(proj), proj_d.data_ptr(), st)
    lm_d = torch.from_numpy(m["landmarks"]).cuda()
    b.insert_batch_device(frames, proj_d.data_ptr(), len(proj), lm_d.data_ptr(), st)
    b.set_landmark_positions(m["landmark_xyz"])
    cams = capi.make_cameras([synthetic.camera_dict()])
    qf = frames_of(q["frames"])
    ra = a.query_batch(qf, q["bits"], q["keypoints"], cams, want_matches=True)
    rb = b.query_batch(qf, q["bits"], q["keypoints"], cams, want_matches=True)
    assert ra["results"].tobytes() == rb["results"].tobytes()
    assert ra["matches"].tobytes() == rb["matches"].tobytes()
    assert ra["results"]["accepted"].sum() > 0


def test_insert_refuses_a_keyframe_that_is_already_in_the_database():
    m, blob, _, q = small_world()
    det = capi.Detector(blob)
    ora = po.Engine(blob)
    proj = det.project(m["bits"][:30])
    det.insert(5, 77, 0, 1, proj[:10], np.arange(10))
    ora.insert(5, 77, 0, 1, proj[:10], np.arange(10))
    det.insert(5, 77, 1, 1, proj[10:20], np.arange(10))       # other camera of the same vertex: fine
    ora.insert(5, 77, 1, 1, proj[10:20], np.arange(10))
    with pytest.raises(capi.MlcError, match="already in the database"):
        det.insert(9, 77, 0, 1, proj[20:], np.arange(10))
    with pytest.raises(ValueError):
        ora.insert(9, 77, 0, 1, proj[20:], np.arange(10))
    # a refused batch leaves the database untouched, also when the duplicate sits inside the batch
    with pytest.raises(capi.MlcError):
        det.insert_batch(capi.make_frames([1, 2], [88, 88], [1, 1], [0, 0], [5, 5]), proj[:10], np.arange(10))
    assert det.num_descriptors() == ora.num_descriptors() == 20 and det.num_entries() == ora.num_entries() == 2
    det.clear()
    det.insert(9, 77, 0, 1, proj[20:], np.arange(10))          # after Clear the id is free again
    assert det.num_descriptors() == 10
