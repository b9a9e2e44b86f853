#!/usr/bin/env python
"""BASELINE config 5 as a one-off measurement (not a bench line): IMI kNN over 100 M projected
descriptors on ONE B200 — 1 M query descriptors, k = 10, nw = 10, W = 1000 x 1000 cells.
Database and queries are vocabulary-conditioned (a random word pair + Gaussian residual), which
gives the near-uniform ~100 entries per cell SURVEY §8d assumes. Prints one JSON line with the scan
kernel's CUDA-event time, algorithmic bytes and fraction of the measured HBM peak.

    python profiles/microbench/knn_100m.py [--db 100000000] [--queries 1000000]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def conditioned(rng, W1, W2, n, sigma):
    out = np.empty((n, 10), np.float32)
    for s in range(0, n, 1 << 22):
        e = min(n, s + (1 << 22))
        i1 = rng.integers(0, W1.shape[1], e - s)
        i2 = rng.integers(0, W2.shape[1], e - s)
        out[s:e, :5] = W1.T[i1]
        out[s:e, 5:] = W2.T[i2]
        out[s:e] += rng.standard_normal((e - s, 10), dtype=np.float32) * np.float32(sigma)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--db", type=int, default=100_000_000)
    ap.add_argument("--queries", type=int, default=1_000_000)
    ap.add_argument("--k", type=int, default=10)
    a = ap.parse_args()
    import torch
    from maplab_b200 import capi, synthetic
    rng = np.random.default_rng(1)
    # vocabulary: 1000 well-separated words per half (a seeded lattice-like random set)
    W1 = (rng.standard_normal((5, 1000)) * 3.0).astype(np.float32)
    W2 = (rng.standard_normal((5, 1000)) * 3.0).astype(np.float32)
    P = np.zeros((10, 512), np.float32)
    blob = synthetic.serialize_vocabulary(P, W1, W2, 10)
    det = capi.Detector(blob, capi.default_settings(num_nearest_neighbors=a.k))
    t0 = time.time()
    per_kf = 500
    nkf = a.db // per_kf
    n = nkf * per_kf
    proj = conditioned(rng, W1, W2, n, 0.15)
    frames = capi.make_frames(np.arange(nkf, dtype=np.int64) * 10**9, np.arange(nkf, dtype=np.int64),
                              np.zeros(nkf, np.int64), np.zeros(nkf, np.int32), np.full(nkf, per_kf, np.int32))
    t_gen = time.time() - t0
    t0 = time.time()
    det.insert_batch(frames, proj, np.arange(n, dtype=np.int64))
    del proj
    det.initialize()
    t_build = time.time() - t0
    q = conditioned(rng, W1, W2, a.queries, 0.15)
    dev = torch.device("cuda", 0)
    q_d = torch.from_numpy(q).to(dev)
    idx = torch.empty((a.queries, a.k), dtype=torch.int32, device=dev)
    dst = torch.empty((a.queries, a.k), dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    ms, st = [], None
    for i in range(6):
        flush.fill_(i)
        torch.cuda.synchronize()
        det.knn_device(q_d.data_ptr(), a.queries, a.k, idx.data_ptr(), dst.data_ptr(), stream)
        torch.cuda.synchronize()
        st = det.last_scan_stats()
        if i >= 2:
            ms.append(st["scan_ms"])
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    scan_ms = float(np.mean(ms))
    found = float((idx[:, a.k - 1] >= 0).float().mean().item())
    print(json.dumps({
        "workload": f"IMI kNN microbench: {n} database descriptors, {a.queries} query descriptors, k={a.k}, nw=10, "
                    f"W=1000x1000 cells, vocabulary-conditioned data, 1 B200",
        "scan_ms": scan_ms, "entries_per_query": st["entries"] / a.queries,
        "algorithmic_bytes_per_launch": st["algorithmic_bytes"],
        "achieved_gbs": st["algorithmic_bytes"] / (scan_ms * 1e-3) / 1e9,
        "frac_of_measured_hbm_peak": st["algorithmic_bytes"] / (scan_ms * 1e-3) / 1e9 / peak,
        "query_descriptors_per_s": a.queries / (scan_ms * 1e-3),
        "queries_with_k_neighbours": found, "generate_s": round(t_gen, 1), "insert_and_build_s": round(t_build, 1)}))


if __name__ == "__main__":
    main()
