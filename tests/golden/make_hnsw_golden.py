#!/usr/bin/env python
"""Golden vectors of the reference's HNSW engine: the REAL vendored hnswlib, compiled from the reference
checkout into oracle/_ref/libhnsw_ref.so (`make -C oracle ref`, build container only), run on a seeded
set of float descriptors exactly as loop_closure::HSNWIndexInterface does (M = 12, ef_construction = 50,
ef_query = 50, single-threaded insertion). Writes tests/golden/hnsw_reference.npz (database, queries, the
reference's neighbour lists) so that the GPU box — where /root/reference does not exist — can check the
B200 engine against them.      python tests/golden/make_hnsw_golden.py"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def reference_knn(db, q, k, M=12, ef_construction=50, ef_query=50):
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libhnsw_ref.so"))
    lib.hnsw_ref_create.restype = C.c_void_p
    h = C.c_void_p(lib.hnsw_ref_create(db.shape[1], C.c_int64(len(db)), M, ef_construction, ef_query))
    lib.hnsw_ref_add(h, db.ctypes.data_as(C.c_void_p), C.c_int64(len(db)))
    idx = np.zeros((len(q), k), np.int32)
    dist = np.zeros((len(q), k), np.float32)
    rc = lib.hnsw_ref_knn(h, q.ctypes.data_as(C.c_void_p), C.c_int64(len(q)), k, idx.ctypes.data_as(C.c_void_p),
                          dist.ctypes.data_as(C.c_void_p))
    lib.hnsw_ref_destroy(h)
    assert rc == 0
    return idx, dist


def world(seed=3, n_db=6000, n_q=300, dim=32, clusters=40):
    """Clustered descriptors (learned descriptors of revisited places are close to each other)."""
    rng = np.random.default_rng(seed)
    centres = rng.standard_normal((clusters, dim)).astype(np.float32) * 2.0
    db = (centres[rng.integers(0, clusters, n_db)] + 0.5 * rng.standard_normal((n_db, dim))).astype(np.float32)
    q = (db[rng.integers(0, n_db, n_q)] + 0.2 * rng.standard_normal((n_q, dim))).astype(np.float32)
    return np.ascontiguousarray(db), np.ascontiguousarray(q)


if __name__ == "__main__":
    db, q = world()
    idx, dist = reference_knn(db, q, 8)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "hnsw_reference.npz"), db=db, q=q, idx=idx, dist=dist)
    print("hnsw_reference.npz:", db.shape, q.shape, idx.shape)
