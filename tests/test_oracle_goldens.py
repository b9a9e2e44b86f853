"""Pin the CPU oracle to the reference's own golden vectors.

Fixtures: tests/golden/reference_goldens.json, extracted from the reference's gtest
files by tests/golden/extract_goldens.py.
"""
import json
import math
import os

import numpy as np
import pytest

from oracle import pyoracle as po

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_goldens.json")))


def test_insert_neighbor_works():
    # test_inverted-multi-index-common.cc:17-64
    for case in G["insert_neighbor"]:
        idx = [i for i, _ in case["inserts"]]
        dist = [d for _, d in case["inserts"]]
        oi, od = po.insert_neighbors(idx, dist, case["k"])
        assert len(oi) == len(case["expected"])
        for (ed, ei), gi, gd in zip(case["expected"], oi, od):
            assert gi == ei
            assert gd == np.float32(ed)


def test_multi_sequence_algorithm_works():
    # test_inverted-multi-index-common.cc:66-150
    m = G["multi_sequence"]
    for n in m["prefixes"]:
        got = po.multi_sequence(m["indices_1"], m["distances_1"], m["indices_2"], m["distances_2"], n)
        assert got.tolist() == m["expected"][:n]


def test_find_closest_words_works():
    # test_inverted-multi-index-common.cc:152-238 (epsilon = 0.2, default radius)
    f = G["find_closest_words"]
    w1, w2 = po.colmajor(f["words1"]), po.colmajor(f["words2"])
    for q, nc, exp in zip(f["queries"], f["num_closest"], f["expected"]):
        got = po.find_closest_words(w1, 10, w2, 5, 3, q, nc, eps=f["epsilon"])
        assert got.tolist() == exp


def test_imi_add_descriptors_works():
    # test_inverted-multi-index.cc:53-133
    t = G["imi"]
    imi = po.IMI(po.colmajor(t["words1"]), 10, po.colmajor(t["words2"]), 5, 3, 10, eps=t["epsilon"])
    desc = np.asarray(t["descriptors"], np.float32).T  # [50][6]
    imi.add(desc, 50)
    assert imi.num_descriptors() == 50
    assert imi.num_files() == t["num_cells"]
    nearest = t["nearest_word_per_descriptor"]
    first_touch = {}
    for w in nearest:
        first_touch.setdefault(w, len(first_touch))
    counts = {}
    for i, w in enumerate(nearest):
        assert imi.cell_of(desc[i]) == w
        b = imi.bucket_of_word(w)
        assert b == first_touch[w]
        idx, d = imi.file(b)
        j = counts.get(b, 0)
        assert idx[j] == i
        assert np.array_equal(d[j], desc[i])
        counts[b] = j + 1
    imi.clear()
    assert imi.num_descriptors() == 0


def test_imi_get_n_nearest_neighbors_works():
    # test_inverted-multi-index.cc:135-227: result == linear search over the activated cells.
    # NOTE the reference runs this test at the DEFAULT epsilon for the add (flag order in
    # gtest is file order: AddDescriptorsWorks sets 0.2 first and the flag is sticky).
    t = G["imi"]
    w1, w2 = po.colmajor(t["words1"]), po.colmajor(t["words2"])
    imi = po.IMI(w1, 10, w2, 5, 3, 10, eps=t["epsilon"])
    desc = np.asarray(t["descriptors"], np.float32).T
    q = np.asarray(t["query_descriptors"], np.float32).T
    imi.add(desc, 50)
    nearest = t["nearest_word_per_descriptor"]
    idx, dist = imi.knn(q, 10, 10)
    for i in range(10):
        words = po.find_closest_words(w1, 10, w2, 5, 3, q[i], 10, eps=t["epsilon"])
        activated = {int(a) * 5 + int(b) for a, b in words}
        gt = sorted((po.squared_distance(desc[j], q[i]), j) for j in range(50) if nearest[j] in activated)
        n = min(10, len(gt))
        for j in range(n):
            assert idx[i, j] == gt[j][1]
            assert dist[i, j] == np.float32(gt[j][0])
            # EXPECT_FLOAT_EQ against a plain float evaluation (4 ULP)
            plain = np.float32(np.sum((desc[gt[j][1]].astype(np.float64) - q[i]) ** 2))
            assert abs(dist[i, j] - plain) <= 4 * np.spacing(plain)
        assert (idx[i, n:] == -1).all() and np.isinf(dist[i, n:]).all()


def test_product_quantization():
    # test_product-quantization.cc:23-86
    t = G["pq"]
    centers = po.colmajor(t["centers"])
    vecs = np.asarray(t["vectors"], np.float32).T  # [4 vectors][4 dims]
    codes = po.pq_quantize(centers, 2, 2, 5, vecs, 4)
    assert codes.T.tolist() == t["quantized_vectors"]
    lut = po.pq_fill_lut(centers, 2, 2, 5, t["query_vector"])
    assert lut.tolist() == t["expected_lut"]
    qv = np.asarray(t["quantized_vectors"], np.int32).T
    assert po.pq_distances(t["expected_lut"], qv).tolist() == t["expected_distances"]
    assert po.pq_distances(t["expected_lut"], qv, add_to=t["add_initial"]).tolist() == t["expected_added"]


def test_imipq():
    # test_inverted-multi-index-product-quantization.cc:68-146, template <int,4,1,2>
    t = G["imipq"]
    idx = po.IMIPQ(po.colmajor(t["words1"]), 4, po.colmajor(t["words2"]), 4, 2,
                   t["quantizer_centers_1"], t["quantizer_centers_2"], 4, 1, 2,
                   t["num_closest_words"], eps=t["epsilon"])
    desc = np.asarray(t["descriptors"], np.float32).T
    idx.add(desc, 5)
    for w, e in zip(t["activated_product_words"], t["expected_map_entries"]):
        assert idx.bucket_of_word(w) == e
    assert idx.num_files() == 4
    counter = 0
    for b in range(4):
        ids, codes = idx.file(b)
        assert len(ids) == t["expected_num_entries_per_inverted_file"][b]
        for j in range(len(ids)):
            assert ids[j] == counter
            assert codes[j].tolist() == t["expected_quantized_descriptors"][counter]
            counter += 1
    gi, gd = idx.knn(np.asarray([t["query"]], np.float32), 1, t["num_neighbors"])
    assert gi[0].tolist() == t["expected_indices"]
    for got, exp in zip(gd[0], t["expected_distances"]):
        if exp == "inf":
            assert math.isinf(got)
        else:
            assert got == np.float32(exp)


def test_scoring():
    # test_scoring.cc:94-140
    t = G["scoring"]
    acc = po.score(t["num_matches"], t["num_descriptors"], t["num_db"], False)
    assert acc.tolist() == t["expected_accumulation"]
    prob = po.score(t["num_matches"], t["num_descriptors"], t["num_db"], True)
    assert np.allclose(prob, t["expected_probabilistic"], atol=t["tolerance"])
    assert len(po.score([], [], 50, False)) == 0
    assert len(po.score([], [], 50, True)) == 0


def test_add_descriptor_bucket_order():
    # test_inverted-multi-index-common.cc:240-296: buckets are created in first-touch order.
    t = G["add_descriptor"]
    first = {}
    files = []
    for w, did in zip(t["word_index_per_descriptor"], t["descriptor_ids"]):
        if w not in first:
            first[w] = len(files)
            files.append([])
        files[first[w]].append(did)
    assert [first[w] for w in t["expected_word_indices"]] == t["expected_word_index_mapped_values"]
    assert files == t["expected_indices"]
