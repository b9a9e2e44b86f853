#!/usr/bin/env python3
"""Extract the golden vectors of the reference's own gtest files into JSON.

Runs in the dev container only (reads /root/reference, which does not exist on
the GPU box). The JSON files it writes are committed next to this script and
are what tests/ load. Sources:
  MBL/test/test_inverted-multi-index-common.cc   :17-64, :66-150, :152-238, :240-296
  MBL/test/test_inverted-multi-index.cc          :27-133, :135-227
  MBL/test/test_product-quantization.cc          :10-86
  MBL/test/test_inverted-multi-index-product-quantization.cc :35-146
  MBL/test/test_scoring.cc                       :40-55
with MBL = algorithms/loopclosure/matching-based-loopclosure.
Eigen's comma initialiser fills row-major; matrices are stored here row-major as
nested lists [rows][cols].
"""
import json
import os
import re

MBL = "/root/reference/algorithms/loopclosure/matching-based-loopclosure/test/"
OUT = os.path.dirname(os.path.abspath(__file__))


def strip_comments(s):
    return re.sub(r"//[^\n]*", "", s)


def nums(s):
    return [float(x) for x in re.findall(r"-?\d+\.?\d*(?:e-?\d+)?", s)]


def comma_inits(src):
    """name -> list of number lists, in order of appearance."""
    out = {}
    for m in re.finditer(r"(\w+(?:\[\d+\])?)\s*<<\s*([-\d\.,\s e]+);", src):
        out.setdefault(m.group(1), []).append(nums(m.group(2)))
    return out


def reshape(flat, rows, cols):
    assert len(flat) == rows * cols, (len(flat), rows, cols)
    return [flat[r * cols:(r + 1) * cols] for r in range(rows)]


def pairs_after(src, name):
    m = re.search(name + r"\s*=\s*\{(.*?)\};", src, re.S)
    body = m.group(1)
    return [[int(a), int(b)] for a, b in re.findall(r"make_pair\((-?\d+),\s*(-?\d+)\)", body)]


def float_pairs_after(src, name):
    m = re.search(name + r"\s*=\s*\{(.*?)\};", src, re.S)
    return [[float(a), int(b)] for a, b in re.findall(r"make_pair\(([-\d\.]+),\s*(-?\d+)\)", m.group(1))]


def int_list_after(src, name):
    m = re.search(name + r"\s*=\s*\{([^;]*?)\};", src, re.S)
    return [int(x) for x in re.findall(r"-?\d+", m.group(1))]


def main():
    g = {}
    # ------------------------------------------------------------------ common
    src = strip_comments(open(MBL + "test_inverted-multi-index-common.cc").read())
    ins = re.findall(r"InsertNeighbor\((\d+),\s*([\d\.]+),\s*(\d+),", src)
    seq1 = [(int(i), float(d)) for i, d, k in ins if k == "5"]
    seq2 = [(int(i), float(d)) for i, d, k in ins if k == "10"]
    g["insert_neighbor"] = [
        {"k": 5, "inserts": seq1, "expected": float_pairs_after(src, "expected_neighbors")},
        {"k": 10, "inserts": seq2, "expected": float_pairs_after(src, "expected_neighbors2")},
    ]
    ci = comma_inits(src)
    g["multi_sequence"] = {
        "distances_1": ci["distances_1"][0], "indices_1": [int(x) for x in ci["indices_1"][0]],
        "distances_2": ci["distances_2"][0], "indices_2": [int(x) for x in ci["indices_2"][0]],
        "expected": pairs_after(src, "expected_closest_words"),
        "prefixes": [24, 10, 5, 3, 2, 1, 0],
    }
    g["find_closest_words"] = {
        "epsilon": 0.2,
        "words1": reshape(ci["words1"][0], 3, 10), "words2": reshape(ci["words2"][0], 3, 5),
        "queries": [ci["query1"][0], ci["query2"][0], ci["query3"][0]],
        "num_closest": [15, 200, 3],
        "expected": [pairs_after(src, "expected_closest_words1"),
                     pairs_after(src, "expected_closest_words2"),
                     pairs_after(src, "expected_closest_words3")],
    }
    g["add_descriptor"] = {
        "descriptors": reshape(ci["descriptors"][0], 6, 5),
        "word_index_per_descriptor": int_list_after(src, "word_index_per_descriptor"),
        "descriptor_ids": int_list_after(src, "descriptor_ids"),
        "expected_word_indices": int_list_after(src, "expected_word_indices"),
        "expected_word_index_mapped_values": int_list_after(src, "expected_word_index_mapped_values"),
        "expected_indices": [[19, 17], [5, 4], [6]],
    }
    # --------------------------------------------------------------------- imi
    src = strip_comments(open(MBL + "test_inverted-multi-index.cc").read())
    ci = comma_inits(src)
    g["imi"] = {
        "epsilon": 0.2,
        "words1": reshape(ci["words1_"][0], 3, 10), "words2": reshape(ci["words2_"][0], 3, 5),
        "descriptors": reshape(ci["descriptors"][0], 6, 50),
        "nearest_word_per_descriptor": int_list_after(src, "nearest_word_per_descriptor"),
        "query_descriptors": reshape(ci["query_descriptors"][0], 6, 10),
        "num_cells": 32, "num_neighbors": 10, "num_closest_words": 10,
    }
    assert ci["descriptors"][0] == ci["descriptors"][1]
    # ---------------------------------------------------------------------- pq
    src = strip_comments(open(MBL + "test_product-quantization.cc").read())
    ci = comma_inits(src)
    g["pq"] = {
        "centers": reshape(ci["centers_"][0], 2, 10),
        "quantized_vectors": reshape([int(x) for x in ci["quantized_vectors_"][0]], 2, 4),
        "vectors": reshape(ci["vectors"][0], 4, 4),
        "query_vector": ci["query_vector"][0],
        "expected_lut": reshape(ci["expected_lut"][0], 2, 5),
        "expected_distances": ci["expected_distances"][0],
        "add_initial": ci["distances"][0],
        "expected_added": ci["expected_distances"][1],
    }
    # ------------------------------------------------------------------- imipq
    src = strip_comments(open(MBL + "test_inverted-multi-index-product-quantization.cc").read())
    ci = comma_inits(src)
    g["imipq"] = {
        "epsilon": 0.2,
        "descriptors": reshape(ci["descriptors_"][0], 4, 5),
        "words1": reshape(ci["words1_"][0], 2, 4), "words2": reshape(ci["words2_"][0], 2, 4),
        "quantizer_centers_1": ci["quantizer_centers_1_"][0],
        "quantizer_centers_2": ci["quantizer_centers_2_"][0],
        "num_closest_words": 16,
        "activated_product_words": int_list_after(src, "activated_product_words"),
        "expected_map_entries": int_list_after(src, "expected_map_entries"),
        "expected_quantized_descriptors": [[int(x) for x in ci["expected_quantized_descriptors[%d]" % i][0]]
                                           for i in range(5)],
        "expected_num_entries_per_inverted_file": int_list_after(src, "expected_num_entries_per_inverted_file"),
        "query": ci["query_descriptor"][0],
        "num_neighbors": 8,
        "expected_indices": int_list_after(src, "expected_indices"),
        "expected_distances": [3.5, 8.5, 10.5, 15.5, 28.5, "inf", "inf", "inf"],
    }
    # ----------------------------------------------------------------- scoring
    g["scoring"] = {
        "num_matches": [2, 10, 6], "num_descriptors": [10, 10, 10], "num_db": 50,
        "expected_accumulation": [2, 10, 6],
        "expected_probabilistic": [0, 3.12392, 1.08807], "tolerance": 1e-4,
    }
    with open(os.path.join(OUT, "reference_goldens.json"), "w") as f:
        json.dump(g, f, indent=1)
    print("wrote reference_goldens.json with keys", sorted(g))


if __name__ == "__main__":
    main()
