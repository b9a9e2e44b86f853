"""Database persistence: an index written by mlc_save_index and read back by
mlc_load_index must answer exactly like the index that was saved — kNN lists, accepted matches,
verdicts and poses byte-identical — also after further inserts, for both engines and for a shard."""
import numpy as np
import pytest

from maplab_b200 import capi, synthetic
from helpers import frames_of, small_world

pytestmark = pytest.mark.gpu


def _fill(det, m, upto=None):
    frames = frames_of(m["frames"])
    proj = det.project(m["bits"])
    if upto is None:
        det.insert_batch(frames, proj, m["landmarks"])
    else:
        nd = int(frames["num_descriptors"][:upto].sum())
        det.insert_batch(frames[:upto], proj[:nd], m["landmarks"][:nd])
    det.set_landmark_positions(m["landmark_xyz"])
    return frames, proj


@pytest.mark.parametrize("kw", [dict(), dict(engine=1), dict(shard_rank=1, shard_count=3)])
def test_saved_index_answers_identically(tmp_path, kw):
    m, blob, voc, q = small_world(num_queries=12)
    if kw.get("engine") == 1:
        blob = synthetic.add_product_quantizer(voc, 10, 16)
    a = capi.Detector(blob, capi.default_settings(num_nearest_neighbors=6, **kw))
    _fill(a, m)
    path = tmp_path / "index.mlc"
    a.save_index(path)
    b = capi.Detector(blob, capi.default_settings(num_nearest_neighbors=6, **kw))
    b.load_index(path)
    assert (b.num_descriptors(), b.num_entries()) == (a.num_descriptors(), a.num_entries())
    qp = a.project(q["bits"])
    ia, da = a.knn(qp, 6)
    ib, db = b.knn(qp, 6)
    assert np.array_equal(ia, ib) and np.array_equal(da, db) and (ia >= 0).any()
    if "shard_count" not in kw:
        cams = capi.make_cameras([synthetic.camera_dict()])
        qframes = frames_of(q["frames"])
        ra = a.query_batch(qframes, q["bits"], q["keypoints"], cams, want_matches=True)
        rb = b.query_batch(qframes, q["bits"], q["keypoints"], cams, want_matches=True)
        assert ra["results"].tobytes() == rb["results"].tobytes()
        assert ra["matches"].tobytes() == rb["matches"].tobytes()
        assert ra["results"]["accepted"].sum() > 0


def test_insert_after_load_equals_one_build(tmp_path):
    m, blob, _, q = small_world(num_queries=12)
    full = capi.Detector(blob, capi.default_settings(num_nearest_neighbors=6))
    frames, proj = _fill(full, m)
    half = len(frames) // 2
    first = capi.Detector(blob, capi.default_settings(num_nearest_neighbors=6))
    _fill(first, m, upto=half)
    first.save_index(tmp_path / "half.mlc")
    b = capi.Detector(blob, capi.default_settings(num_nearest_neighbors=6))
    b.load_index(tmp_path / "half.mlc")
    nd = int(frames["num_descriptors"][:half].sum())
    b.insert_batch(frames[half:], proj[nd:], m["landmarks"][nd:])
    qp = full.project(q["bits"])
    i0, d0 = full.knn(qp, 6)
    i1, d1 = b.knn(qp, 6)
    assert np.array_equal(i0, i1) and np.array_equal(d0, d1)


def test_load_rejects_foreign_or_damaged_files(tmp_path):
    m, blob, voc, _ = small_world(num_queries=12)
    a = capi.Detector(blob)
    _fill(a, m)
    path = tmp_path / "index.mlc"
    a.save_index(path)
    other_blob, _ = synthetic.make_vocabulary(m["bits"][:5000], num_words=32, seed=99)
    with pytest.raises(capi.MlcError):
        capi.Detector(other_blob).load_index(path)                       # another vocabulary
    with pytest.raises(capi.MlcError):
        capi.Detector(blob, capi.default_settings(shard_rank=0, shard_count=2)).load_index(path)
    data = path.read_bytes()
    (tmp_path / "cut.mlc").write_bytes(data[:len(data) // 2])
    c = capi.Detector(blob)
    with pytest.raises(capi.MlcError):
        c.load_index(tmp_path / "cut.mlc")                               # truncated
    assert c.num_descriptors() == 0                                      # untouched by the failed load
    (tmp_path / "junk.mlc").write_bytes(b"not an index" * 10)
    with pytest.raises(capi.MlcError):
        c.load_index(tmp_path / "junk.mlc")
    with pytest.raises(capi.MlcError):
        c.load_index(tmp_path / "missing.mlc")
