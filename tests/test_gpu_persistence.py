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


def test_load_validates_what_the_header_claims(tmp_path):
    """A file whose magic and vocabulary hash match but whose contents were damaged must be refused with
    an error (no exception through the C boundary, no out-of-bounds device reads later) and must leave
    the detector's database as it was."""
    import struct
    m, blob, _, q = small_world(num_queries=4)
    a = capi.Detector(blob, capi.default_settings(num_nearest_neighbors=6))
    _fill(a, m)
    path = tmp_path / "index.mlc"
    a.save_index(path)
    data = bytearray(path.read_bytes())
    hdr = struct.Struct("<8sQiiiiIiqqqqQ")
    h = list(hdr.unpack_from(data, 0))
    n, no, nkf, nxyz, list_bytes = h[8], h[9], h[10], h[11], h[12]
    num_cells = h[6]
    kf_bytes = 40 * nkf
    off_lm = hdr.size + kf_bytes
    off_gidx = off_lm + 8 * n
    off_desc = off_gidx + 4 * no
    off_cells = off_desc + 40 * no
    off_info = off_cells + 4 * no
    off_lists = off_info + 8 * num_cells
    assert off_lists + list_bytes + 24 * nxyz == len(data)
    qp = a.project(q["bits"])
    ref = a.knn(qp, 6)

    def damaged(mutate):
        d = bytearray(data)
        mutate(d)
        p = tmp_path / "damaged.mlc"
        p.write_bytes(bytes(d))
        c = capi.Detector(blob, capi.default_settings(num_nearest_neighbors=6))
        _fill(c, m)
        with pytest.raises(capi.MlcError):
            c.load_index(p)
        got = c.knn(qp, 6)   # database untouched and still consistent
        assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])

    def set_header(i, v):
        def f(d):
            hh = list(h)
            hh[i] = v
            hdr.pack_into(d, 0, *hh)
        return f

    damaged(set_header(8, 2**40))            # absurd descriptor count: no bad_alloc, an error
    damaged(set_header(8, n + 1))            # sizes no longer match the file
    damaged(set_header(10, nkf - 1))
    damaged(set_header(7, 3))                # list_dim of the other engine
    damaged(set_header(12, list_bytes + 16))
    # a cell whose list would run past the list block
    nonempty = np.frombuffer(data, "<u4", 2 * num_cells, off_info).reshape(-1, 2)
    c0 = int(np.nonzero(nonempty[:, 1])[0][0])
    damaged(lambda d: struct.pack_into("<I", d, off_info + 8 * c0 + 4, 2**30))
    damaged(lambda d: struct.pack_into("<I", d, off_info + 8 * c0, 2**31))
    # an inverted-list entry naming a descriptor outside the database
    start16 = int(nonempty[c0, 0])
    damaged(lambda d: struct.pack_into("<I", d, off_lists + 16 * start16 + 4 * 10, n + 5))
    # keyframe table that does not tile the descriptors
    damaged(lambda d: struct.pack_into("<i", d, hdr.size + 40 * 1 + 28, 3))
    # the undamaged file still loads
    c = capi.Detector(blob, capi.default_settings(num_nearest_neighbors=6))
    c.load_index(path)
    got = c.knn(qp, 6)
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])
