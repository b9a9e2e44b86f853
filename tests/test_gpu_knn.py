"""Parity of kernel 2 (coarse word search + inverted-list scan) with the oracle: visited cells,
cell assignment and neighbour indices bit-exact (ties by index), distances bit-exact (the spec
allows 1e-5 relative; the canonical fp32 order makes them identical)."""
import json
import os

import numpy as np
import pytest

from maplab_b200 import capi, synthetic
from oracle import pyoracle as po
from helpers import fill_oracle, frames_of, small_world

pytestmark = pytest.mark.gpu
G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_goldens.json")))


def _golden_blob(t):
    W1 = np.asarray(t["words1"], np.float32)
    W2 = np.asarray(t["words2"], np.float32)
    P = np.zeros((2 * W1.shape[0], 16), np.float32)  # projection unused here
    return synthetic.serialize_vocabulary(P, W1, W2)


def test_reference_golden_find_closest_words():
    # test_inverted-multi-index-common.cc:152-238 through the device kd-tree traversal
    f = G["find_closest_words"]
    blob = _golden_blob(f)
    det = capi.Detector(blob, capi.default_settings(knn_epsilon=f["epsilon"]))
    for q, nc, exp in zip(f["queries"], f["num_closest"], f["expected"]):
        nc_dev = min(nc, 16)
        cells = det.coarse_cells(np.asarray([q], np.float32), nc_dev)[0]
        exp_cells = [a * 5 + b for a, b in exp][:nc_dev]
        assert cells[:len(exp_cells)].tolist() == exp_cells


def test_reference_golden_imi_add_and_knn():
    # test_inverted-multi-index.cc:53-227
    t = G["imi"]
    blob = _golden_blob(t)
    det = capi.Detector(blob, capi.default_settings(knn_epsilon=t["epsilon"]))
    desc = np.asarray(t["descriptors"], np.float32).T
    q = np.asarray(t["query_descriptors"], np.float32).T
    assert det.coarse_cells(desc, 1)[:, 0].tolist() == t["nearest_word_per_descriptor"]
    det.insert(0, 0, 0, 0, desc, np.arange(50))
    assert det.num_descriptors() == 50
    idx, dist = det.knn(q, 10)
    nearest = t["nearest_word_per_descriptor"]
    w1, w2 = po.colmajor(t["words1"]), po.colmajor(t["words2"])
    for i in range(10):
        words = po.find_closest_words(w1, 10, w2, 5, 3, q[i], 10, eps=t["epsilon"])
        activated = {int(a) * 5 + int(b) for a, b in words}
        gt = sorted((po.squared_distance(desc[j], q[i]), j) for j in range(50) if nearest[j] in activated)
        n = min(10, len(gt))
        assert idx[i, :n].tolist() == [j for _, j in gt[:n]]
        assert dist[i, :n].tolist() == [np.float32(d) for d, _ in gt[:n]]
        assert (idx[i, n:] == -1).all() and np.isinf(dist[i, n:]).all()


@pytest.mark.parametrize("eps,radius", [(2.0, 20.0), (0.0, 20.0), (2.0, 1.5)])
def test_cells_and_knn_match_oracle(eps, radius):
    m, blob, _, q = small_world()
    s = capi.default_settings(knn_epsilon=eps, knn_max_radius=radius)
    det = capi.Detector(blob, s)
    ora = po.Engine(blob, po.default_settings(knn_epsilon=eps, knn_max_radius=radius))
    proj = det.project(m["bits"])
    frames = frames_of(m["frames"])
    det.insert_batch(frames, proj, m["landmarks"])
    fill_oracle(ora, frames, proj, m["landmarks"])
    assert det.num_descriptors() == ora.num_descriptors() and det.num_entries() == ora.num_entries()
    assert det.num_neighbors() == ora.num_neighbors()
    qp = det.project(q["bits"])
    # P3: visited cells
    imi = po.IMI(*_words(blob), 5, 10, eps=eps, radius=radius)
    assert np.array_equal(det.coarse_cells(qp, 10), imi.visited_cells(qp, len(qp)))
    # P2: cell of every database descriptor
    exp_cells = np.array([imi.cell_of(d) for d in proj[:2000]])
    assert np.array_equal(det.coarse_cells(proj[:2000], 1)[:, 0], exp_cells)
    for k in (1, 3, 8, 10):
        idx, dist = det.knn(qp, k)
        oidx, odist = ora.knn(qp, k)
        assert np.array_equal(idx, oidx)
        assert np.array_equal(dist, odist)
    st = det.last_scan_stats()
    assert st["algorithmic_bytes"] == 44 * st["entries"] and st["entries"] > 0


def _words(blob):
    v = synthetic.parse_vocabulary(blob)
    return (po.colmajor(v["W1"]), v["W1"].shape[1], po.colmajor(v["W2"]), v["W2"].shape[1])


def test_ties_broken_by_index_and_missing_neighbours_trail():
    m, blob, _, q = small_world()
    det = capi.Detector(blob)
    ora = po.Engine(blob)
    proj = det.project(m["bits"][:40])
    dup = np.concatenate([proj, proj, proj])  # identical descriptors -> distance ties
    frames = capi.make_frames([0], [0], [0], [0], [len(dup)])
    det.insert_batch(frames, dup, np.arange(len(dup)))
    ora.insert(0, 0, 0, 0, dup, np.arange(len(dup)))
    idx, dist = det.knn(proj, 8)
    oidx, odist = ora.knn(proj, 8)
    assert np.array_equal(idx, oidx) and np.array_equal(dist, odist)
    assert (idx == -1).any()  # few entries per visited cell: lists shorter than k
    for row_i, row_d in zip(idx, dist):
        miss = row_i == -1
        assert np.isinf(row_d[miss]).all()
        if miss.any():
            assert miss[np.argmax(miss):].all()  # trailing


def test_empty_database_and_empty_query():
    m, blob, _, q = small_world()
    det = capi.Detector(blob)
    qp = det.project(q["bits"][:5])
    idx, dist = det.knn(qp, 3)
    assert (idx == -1).all() and np.isinf(dist).all()
    idx, dist = det.knn(np.zeros((0, 10), np.float32), 3)
    assert idx.shape == (0, 3)
    det.insert(0, 0, 0, 0, qp, np.arange(5))
    idx, _ = det.knn(qp, 1)
    assert idx[:, 0].tolist() == [0, 1, 2, 3, 4]
    det.clear()
    assert det.num_descriptors() == 0 and det.num_entries() == 0
    idx, _ = det.knn(qp, 1)
    assert (idx == -1).all()


def test_sharded_lists_merge_to_the_single_index_result():
    """per-shard top-k merged by (distance, index) == single-index result."""
    import torch
    m, blob, _, q = small_world()
    full = capi.Detector(blob)
    proj = full.project(m["bits"])
    frames = frames_of(m["frames"])
    full.insert_batch(frames, proj, m["landmarks"])
    qp = full.project(q["bits"])
    k, G_ = 6, 3
    ref_idx, ref_dist = full.knn(qp, k)
    n = len(qp)
    idx_l = torch.empty((G_, n, k), dtype=torch.int32, device="cuda")
    dist_l = torch.empty((G_, n, k), dtype=torch.float32, device="cuda")
    for r in range(G_):
        sh = capi.Detector(blob, capi.default_settings(shard_rank=r, shard_count=G_))
        sh.insert_batch(frames, proj, m["landmarks"])
        i, d = sh.knn(qp, k)
        assert ((i == -1) | (i % G_ == r)).all()
        idx_l[r] = torch.from_numpy(i).cuda()
        dist_l[r] = torch.from_numpy(d).cuda()
    out_i = torch.empty((n, k), dtype=torch.int32, device="cuda")
    out_d = torch.empty((n, k), dtype=torch.float32, device="cuda")
    full.merge_topk_device(idx_l.data_ptr(), dist_l.data_ptr(), G_, n, k, out_i.data_ptr(),
                           out_d.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(out_i.cpu().numpy(), ref_idx)
    assert np.array_equal(out_d.cpu().numpy(), ref_dist)


def test_reference_golden_imipq_add_and_knn():
    # test_inverted-multi-index-product-quantization.cc:68-146 (template <int, 4, 1, 2>) through the
    # device PQ encode + scan kernels
    t = G["imipq"]
    W1 = np.asarray(t["words1"], np.float32)
    W2 = np.asarray(t["words2"], np.float32)
    Q1 = np.asarray(t["quantizer_centers_1"], np.float32)[None, :]
    Q2 = np.asarray(t["quantizer_centers_2"], np.float32)[None, :]
    blob = synthetic.serialize_vocabulary(np.zeros((4, 16), np.float32), W1, W2, pq=(4, 2, 1, Q1, Q2))
    det = capi.Detector(blob, capi.default_settings(engine=1, knn_epsilon=t["epsilon"],
                                                    num_closest_words=t["num_closest_words"]))
    desc = np.asarray(t["descriptors"], np.float32).T
    det.insert(0, 0, 0, 0, desc, np.arange(len(desc)))
    idx, dist = det.knn(np.asarray([t["query"]], np.float32), t["num_neighbors"])
    assert idx[0].tolist() == t["expected_indices"]
    for got, exp in zip(dist[0], t["expected_distances"]):
        assert (np.isinf(got) if exp == "inf" else got == np.float32(exp))


@pytest.mark.parametrize("ncomp,ncent,k", [(10, 16, 6), (2, 256, 10), (10, 3, 1)])
def test_imipq_knn_matches_oracle(ncomp, ncent, k):
    # imipq engine: indices AND distances bit-identical to the oracle's
    # InvertedMultiProductQuantizationIndex restatement on a seeded synthetic map
    m, _, voc, q = small_world()
    blob = synthetic.add_product_quantizer(voc, ncomp, ncent)
    det = capi.Detector(blob, capi.default_settings(engine=1))
    ora = po.Engine(blob, po.default_settings(engine=1))
    proj = det.project(m["bits"])
    frames = frames_of(m["frames"])
    det.insert_batch(frames, proj, m["landmarks"])
    fill_oracle(ora, frames, proj, m["landmarks"])
    qp = det.project(q["bits"])
    idx, dist = det.knn(qp, k)
    oidx, odist = ora.knn(qp, k)
    assert np.array_equal(idx, oidx)
    assert np.array_equal(dist, odist)
    assert (idx >= 0).any()
    st = det.last_scan_stats()
    bits = int(np.ceil(np.log2(ncent)))
    assert st["algorithmic_bytes"] == st["entries"] * (4 + (ncomp * bits + 7) // 8)
