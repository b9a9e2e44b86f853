"""ctypes loader for the CPU ORACLE (test infrastructure — NOT product code).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module. The product package maplab_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liblc_oracle.so")
_lib = None

c_int_p = C.POINTER(C.c_int)
c_float_p = C.POINTER(C.c_float)
c_double_p = C.POINTER(C.c_double)
c_i64_p = C.POINTER(C.c_int64)
c_u64_p = C.POINTER(C.c_uint64)
c_u8_p = C.POINTER(C.c_uint8)
c_u32_p = C.POINTER(C.c_uint32)


def build(force=False):
    srcs = [f for f in os.listdir(_HERE) if f.endswith((".cc", ".h", ".inc"))]
    newest = max(os.path.getmtime(os.path.join(_HERE, f)) for f in srcs)
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < newest:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.lco_kdtree_create.restype = C.c_void_p
        _lib.lco_imi_create.restype = C.c_void_p
        _lib.lco_imipq_create.restype = C.c_void_p
        _lib.lco_engine_create.restype = C.c_void_p
        _lib.lco_squared_distance.restype = C.c_float
        _lib.lco_kdtree_knn.restype = C.c_ulong
        _lib.lco_ransac_threshold.restype = C.c_double
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a, t):
    return a.ctypes.data_as(t)


def colmajor(m):
    """[rows][cols] nested list / 2-D array -> flat column-major float32."""
    return _f32(np.asarray(m, dtype=np.float32).T.reshape(-1))


# ---------------------------------------------------------------- primitives
def insert_neighbors(idx, dist, k):
    idx, dist = _i32(idx), _f32(dist)
    oi = np.zeros(k, np.int32)
    od = np.zeros(k, np.float32)
    n = lib().lco_insert_neighbors(_p(idx, c_int_p), _p(dist, c_float_p), len(idx), k,
                                   _p(oi, c_int_p), _p(od, c_float_p))
    return oi[:n], od[:n]


def multi_sequence(i1, d1, i2, d2, num_words):
    i1, d1, i2, d2 = _i32(i1), _f32(d1), _i32(i2), _f32(d2)
    out = np.zeros((max(len(i1) * len(i2), 1), 2), np.int32)
    n = lib().lco_multi_sequence(_p(i1, c_int_p), _p(d1, c_float_p), len(i1), _p(i2, c_int_p),
                                 _p(d2, c_float_p), len(i2), num_words, _p(out, c_int_p))
    return out[:n]


class KdTree:
    def __init__(self, cloud_colmajor, dim, n):
        self.cloud = _f32(cloud_colmajor)
        self.dim, self.n = dim, n
        self.h = C.c_void_p(lib().lco_kdtree_create(_p(self.cloud, c_float_p), dim, n))

    def knn(self, q, k, eps, radius):
        q = _f32(q)
        idx = np.zeros(k, np.int32)
        d = np.zeros(k, np.float32)
        lib().lco_kdtree_knn(self.h, _p(q, c_float_p), k, C.c_float(eps), C.c_float(radius),
                             _p(idx, c_int_p), _p(d, c_float_p))
        return idx, d

    def export(self):
        nn = lib().lco_kdtree_num_nodes(self.h)
        nodes = np.zeros((nn, 3), np.uint32)
        buckets = np.zeros(self.n, np.int32)
        lib().lco_kdtree_export(self.h, _p(nodes, c_u32_p), _p(buckets, c_int_p))
        return nodes, buckets

    def __del__(self):
        if lib is not None and self.h:
            lib().lco_kdtree_destroy(self.h)
            self.h = None


def find_closest_words(w1_cm, n1, w2_cm, n2, sub_dim, query, num_closest, eps=2.0, radius=20.0):
    w1_cm, w2_cm, query = _f32(w1_cm), _f32(w2_cm), _f32(query)
    out = np.zeros((max(num_closest, 1), 2), np.int32)
    n = lib().lco_find_closest_words(_p(w1_cm, c_float_p), n1, _p(w2_cm, c_float_p), n2, sub_dim,
                                     _p(query, c_float_p), num_closest, C.c_float(eps),
                                     C.c_float(radius), _p(out, c_int_p))
    return out[:n]


def squared_distance(a, b):
    a, b = _f32(a), _f32(b)
    return float(lib().lco_squared_distance(_p(a, c_float_p), _p(b, c_float_p), len(a)))


class IMI:
    def __init__(self, w1_cm, n1, w2_cm, n2, sub_dim, nw, eps=2.0, radius=20.0):
        self.w1, self.w2 = _f32(w1_cm), _f32(w2_cm)
        self.dim, self.nw = 2 * sub_dim, nw
        self.h = C.c_void_p(lib().lco_imi_create(_p(self.w1, c_float_p), n1, _p(self.w2, c_float_p),
                                                 n2, sub_dim, nw, C.c_float(eps), C.c_float(radius)))

    def add(self, desc_cm, n):
        d = _f32(desc_cm)
        lib().lco_imi_add(self.h, _p(d, c_float_p), n)

    def num_descriptors(self):
        return lib().lco_imi_num_descriptors(self.h)

    def knn(self, q_cm, n, k):
        q = _f32(q_cm)
        idx = np.zeros((n, k), np.int32)
        dist = np.zeros((n, k), np.float32)
        lib().lco_imi_knn(self.h, _p(q, c_float_p), n, k, _p(idx, c_int_p), _p(dist, c_float_p))
        return idx, dist

    def cell_of(self, desc):
        d = _f32(desc)
        return lib().lco_imi_cell_of(self.h, _p(d, c_float_p))

    def visited_cells(self, q_cm, n):
        q = _f32(q_cm)
        cells = np.zeros((n, self.nw), np.int32)
        lib().lco_imi_visited_cells(self.h, _p(q, c_float_p), n, self.nw, _p(cells, c_int_p))
        return cells

    def num_files(self):
        return lib().lco_imi_num_files(self.h)

    def bucket_of_word(self, w):
        return lib().lco_imi_bucket_of_word(self.h, w)

    def file(self, bucket):
        n = lib().lco_imi_file_size(self.h, bucket)
        idx = np.zeros(n, np.int32)
        desc = np.zeros((n, self.dim), np.float32)
        lib().lco_imi_file_get(self.h, bucket, _p(idx, c_int_p), _p(desc, c_float_p))
        return idx, desc

    def clear(self):
        lib().lco_imi_clear(self.h)

    def __del__(self):
        if self.h:
            lib().lco_imi_destroy(self.h)
            self.h = None


class IMIPQ:
    def __init__(self, w1_cm, n1, w2_cm, n2, sub_dim, qc1, qc2, ncomp, dim_per_comp, ncenters, nw,
                 eps=2.0, radius=20.0):
        self.keep = [_f32(w1_cm), _f32(w2_cm), _f32(qc1), _f32(qc2)]
        self.dim, self.ncomp = 2 * sub_dim, ncomp
        self.h = C.c_void_p(lib().lco_imipq_create(
            _p(self.keep[0], c_float_p), n1, _p(self.keep[1], c_float_p), n2, sub_dim,
            _p(self.keep[2], c_float_p), _p(self.keep[3], c_float_p), ncomp, dim_per_comp, ncenters,
            nw, C.c_float(eps), C.c_float(radius)))

    def add(self, desc_cm, n):
        d = _f32(desc_cm)
        lib().lco_imipq_add(self.h, _p(d, c_float_p), n)

    def knn(self, q_cm, n, k):
        q = _f32(q_cm)
        idx = np.zeros((n, k), np.int32)
        dist = np.zeros((n, k), np.float32)
        lib().lco_imipq_knn(self.h, _p(q, c_float_p), self.dim, n, k, _p(idx, c_int_p),
                            _p(dist, c_float_p))
        return idx, dist

    def bucket_of_word(self, w):
        return lib().lco_imipq_bucket_of_word(self.h, w)

    def num_files(self):
        return lib().lco_imipq_num_files(self.h)

    def file(self, bucket):
        n = lib().lco_imipq_file_size(self.h, bucket)
        idx = np.zeros(n, np.int32)
        codes = np.zeros((n, self.ncomp), np.int32)
        lib().lco_imipq_file_get(self.h, bucket, _p(idx, c_int_p), _p(codes, c_int_p))
        return idx, codes

    def __del__(self):
        if self.h:
            lib().lco_imipq_destroy(self.h)
            self.h = None


def pq_quantize(centers_cm, ncomp, dim_per_comp, ncenters, vecs_cm, n):
    c, v = _f32(centers_cm), _f32(vecs_cm)
    codes = np.zeros((n, ncomp), np.int32)
    lib().lco_pq_quantize(_p(c, c_float_p), ncomp, dim_per_comp, ncenters, _p(v, c_float_p), n,
                          _p(codes, c_int_p))
    return codes


def pq_fill_lut(centers_cm, ncomp, dim_per_comp, ncenters, vec):
    c, v = _f32(centers_cm), _f32(vec)
    lut = np.zeros((ncomp, ncenters), np.float32)
    lib().lco_pq_fill_lut(_p(c, c_float_p), ncomp, dim_per_comp, ncenters, _p(v, c_float_p),
                          _p(lut, c_float_p))
    return lut


def pq_distances(lut, codes, add_to=None):
    lut = _f32(lut)
    codes = _i32(codes)
    n, ncomp = codes.shape
    dist = _f32(add_to).copy() if add_to is not None else np.zeros(n, np.float32)
    lib().lco_pq_distances(ncomp, lut.shape[1], _p(lut, c_float_p), _p(codes, c_int_p), n,
                           _p(dist, c_float_p), 1 if add_to is not None else 0)
    return dist


def score(num_matches, num_desc, num_db, probabilistic):
    m = np.ascontiguousarray(num_matches, np.uint64)
    d = np.ascontiguousarray(num_desc, np.uint64)
    s = np.zeros(len(m), np.float32)
    n = C.c_int(0)
    lib().lco_score(int(probabilistic), _p(m, c_u64_p), _p(d, c_u64_p), len(m), C.c_uint64(num_db),
                    _p(s, c_float_p), C.byref(n))
    return s[:n.value]


def project(P_cm, rows, cols, target_dim, raw, float_mode=False):
    """raw: [n][bytes] uint8. Returns [n][target_dim] float32."""
    P = _f32(P_cm)
    raw = np.ascontiguousarray(raw, np.uint8)
    n, nbytes = raw.shape
    out = np.zeros((n, target_dim), np.float32)
    lib().lco_project(_p(P, c_float_p), rows, cols, target_dim, _p(raw, c_u8_p), nbytes, n,
                      _p(out, c_float_p), int(float_mode))
    return out


def quantize_projection(P_cm, rows, cols, target_dim):
    P = _f32(P_cm)
    p_int = np.zeros((target_dim, cols), np.int32)
    shift = np.zeros(target_dim, np.int32)
    lib().lco_quantize_projection(_p(P, c_float_p), rows, cols, target_dim, _p(p_int, c_int_p),
                                  _p(shift, c_int_p))
    return p_int, shift


def vocab_parse(blob, want_pq=False):
    b = np.frombuffer(blob, np.uint8)
    dims = np.zeros(11, np.int32)
    rc = lib().lco_vocab_parse(_p(b, c_u8_p), C.c_uint64(len(b)), int(want_pq), _p(dims, c_int_p))
    return rc, dims


class Settings(C.Structure):
    _fields_ = [("num_closest_words", C.c_int), ("num_nearest_neighbors", C.c_int),
                ("scoring", C.c_int), ("engine", C.c_int),
                ("min_image_time_seconds", C.c_double), ("min_verify_matches_num", C.c_uint64),
                ("fraction_best_scores", C.c_float), ("knn_epsilon", C.c_float),
                ("knn_max_radius", C.c_float)]


def default_settings(**kw):
    s = Settings(10, -1, 0, 0, 10.0, 10, 0.25, 2.0, 20.0)
    for k, v in kw.items():
        setattr(s, k, v)
    return s


class Engine:
    """Oracle LoopDetector. Descriptors are [n][dim] float32 row-per-descriptor."""

    def __init__(self, vocab_blob, settings=None):
        self.s = settings or default_settings()
        self.blob = np.frombuffer(bytes(vocab_blob), np.uint8)
        self.h = C.c_void_p(lib().lco_engine_create(C.byref(self.s), _p(self.blob, c_u8_p),
                                                    C.c_uint64(len(self.blob))))
        assert self.h, "vocabulary parse failed"
        rc, dims = vocab_parse(bytes(vocab_blob), self.s.engine == 1)
        self.dim = int(dims[1])

    def project(self, raw):
        raw = np.ascontiguousarray(raw, np.uint8)
        n, nbytes = raw.shape
        out = np.zeros((n, self.dim), np.float32)
        lib().lco_engine_project(self.h, _p(raw, c_u8_p), nbytes, n, _p(out, c_float_p))
        return out

    def insert(self, ts, vertex, frame_index, mission, proj, landmarks):
        proj = _f32(proj)
        lm = np.ascontiguousarray(landmarks, np.int64)
        rc = lib().lco_engine_insert(self.h, C.c_int64(ts), C.c_int64(vertex), frame_index,
                                     C.c_int64(mission), self.dim, _p(proj, c_float_p), len(lm),
                                     _p(lm, c_i64_p))
        if rc != 0:
            raise ValueError("Insert: keyframe id already in the database (matching-based-engine.cc:244-252)")

    def add_summary_map(self, arrays, mission_id, first_vertex_id, first_landmark_id):
        """addLocalizationSummaryMapToDatabase on deserialized arrays (Eigen shapes: descriptors
        dim x N, G_*_position 3 x n). Returns 0, or the number of the CHECK that would abort."""
        d = np.asarray(arrays["descriptors"], np.float32)
        desc = np.ascontiguousarray(d.T)
        oi = np.ascontiguousarray(arrays["observer_indices"], np.uint32)
        ol = np.ascontiguousarray(arrays["observation_to_landmark_index"], np.uint32)
        u32p = C.POINTER(C.c_uint32)
        return lib().lco_engine_add_summary_map(
            self.h, d.shape[0], C.c_int64(np.asarray(arrays["G_observer_position"]).shape[1]),
            C.c_int64(np.asarray(arrays["G_landmark_position"]).shape[1]), C.c_int64(len(oi)),
            C.c_int64(d.shape[1]), _p(desc, c_float_p), _p(oi, u32p), _p(ol, u32p), C.c_int64(len(ol)),
            C.c_int64(mission_id), C.c_int64(first_vertex_id), C.c_int64(first_landmark_id))

    def num_descriptors(self):
        return lib().lco_engine_num_descriptors(self.h)

    def num_entries(self):
        return lib().lco_engine_num_entries(self.h)

    def num_neighbors(self):
        return lib().lco_engine_num_neighbors(self.h)

    def clear(self):
        lib().lco_engine_clear(self.h)

    def knn(self, q, k):
        q = _f32(q)
        n = q.shape[0]
        idx = np.zeros((n, k), np.int32)
        dist = np.zeros((n, k), np.float32)
        lib().lco_engine_knn(self.h, _p(q, c_float_p), n, k, _p(idx, c_int_p), _p(dist, c_float_p))
        return idx, dist

    def find(self, vertex, mission, frames):
        """frames: list of (ts, frame_index, proj[n][dim]). Returns matches [m][6] int64."""
        ts = np.array([f[0] for f in frames], np.int64)
        fi = np.array([f[1] for f in frames], np.int32)
        nd = np.array([len(f[2]) for f in frames], np.int32)
        proj = _f32(np.concatenate([np.asarray(f[2], np.float32).reshape(-1, self.dim) for f in frames]))
        cap = int(nd.sum()) * max(self.num_neighbors(), 1) + 8
        out = np.zeros((cap, 6), np.int64)
        n = lib().lco_engine_find(self.h, len(frames), _p(ts, c_i64_p), C.c_int64(vertex),
                                  _p(fi, c_int_p), C.c_int64(mission), self.dim, _p(nd, c_int_p),
                                  _p(proj, c_float_p), _p(out, c_i64_p), cap)
        return out[:n]

    def find_frame_trace(self, ts, vertex, frame_index, mission, proj, make_unique=True):
        proj = _f32(proj)
        n = proj.shape[0]
        k = self.num_neighbors()
        counts = np.zeros(5, np.int32)
        knn_idx = np.zeros((n, k), np.int32)
        knn_dist = np.zeros((n, k), np.float32)
        raw = np.zeros((n * k + 1, 6), np.int64)
        filt = np.zeros((n * k + 1, 6), np.int64)
        cand = np.zeros(n * k + 1, np.int32)
        votes = np.zeros(n * k + 1, np.int32)
        sel = np.zeros(n * k + 1, np.int32)
        lib().lco_engine_find_frame_trace(
            self.h, C.c_int64(ts), C.c_int64(vertex), frame_index, C.c_int64(mission), self.dim, n,
            _p(proj, c_float_p), int(make_unique), _p(counts, c_int_p), _p(knn_idx, c_int_p),
            _p(knn_dist, c_float_p), _p(raw, c_i64_p), _p(cand, c_int_p), _p(votes, c_int_p),
            _p(sel, c_int_p), _p(filt, c_i64_p))
        return dict(k=int(counts[0]), knn_idx=knn_idx, knn_dist=knn_dist, raw=raw[:counts[1]],
                    cand=cand[:counts[2]], votes=votes[:counts[2]], selected=sel[:counts[3]],
                    filtered=filt[:counts[4]])

    def __del__(self):
        if self.h:
            lib().lco_engine_destroy(self.h)
            self.h = None


def query_batch(engine, frames, bits, keypoints, landmark_xyz, cams, min_inlier_count=10,
                min_inlier_ratio=0.0, pixel_sigma=2.0, num_iters=100, seed=12345, rng_mapping=1,
                num_threads=1):
    """Whole query path on the CPU for a batch (frames: dict/structured array with timestamp_ns,
    vertex_id, mission_id, frame_index, num_descriptors). Returns dict of per-vertex arrays."""
    ts = np.ascontiguousarray(frames["timestamp_ns"], np.int64)
    vx = np.ascontiguousarray(frames["vertex_id"], np.int64)
    ms = np.ascontiguousarray(frames["mission_id"], np.int64)
    fi = np.ascontiguousarray(frames["frame_index"], np.int32)
    nd = np.ascontiguousarray(frames["num_descriptors"], np.int32)
    bits = np.ascontiguousarray(bits, np.uint8)
    kp = _f64(keypoints)
    xyz = _f64(landmark_xyz)
    arr = (Camera * len(cams))(*cams)
    nf = len(ts)
    sc = np.zeros((nf, 5), np.int32)
    T = np.zeros((nf, 3, 4), np.float64)
    st = np.zeros(3, np.float64)
    nv = lib().lco_query_batch(engine.h, nf, _p(ts, c_i64_p), _p(vx, c_i64_p), _p(ms, c_i64_p),
                               _p(fi, c_int_p), _p(nd, c_int_p), _p(bits, c_u8_p), bits.shape[1],
                               _p(kp, c_double_p), _p(xyz, c_double_p), C.c_int64(len(xyz)), arr,
                               len(cams), min_inlier_count, C.c_double(min_inlier_ratio),
                               C.c_double(pixel_sigma), num_iters, C.c_uint32(seed), rng_mapping,
                               num_threads, _p(sc, c_int_p), _p(T, c_double_p), _p(st, c_double_p))
    sc, T = sc[:nv], T[:nv]
    return dict(accepted=sc[:, 0], num_inliers=sc[:, 1], iterations=sc[:, 2], num_matches=sc[:, 3],
                ransac_success=sc[:, 4], T=T, stage_seconds=dict(project=st[0], find=st[1], verify=st[2]))


# ------------------------------------------------------------------ geometry
class Camera(C.Structure):
    _fields_ = [("fu", C.c_double), ("fv", C.c_double), ("cu", C.c_double), ("cv", C.c_double),
                ("distortion", C.c_int), ("dist", C.c_double * 4), ("R_B_C", C.c_double * 9),
                ("t_B_C", C.c_double * 3)]


def make_camera(fu, fv, cu, cv, R_B_C=None, t_B_C=None, distortion=0, dist=(0, 0, 0, 0)):
    c = Camera()
    c.fu, c.fv, c.cu, c.cv = fu, fv, cu, cv
    c.distortion = distortion
    R = np.eye(3) if R_B_C is None else np.asarray(R_B_C, np.float64)
    t = np.zeros(3) if t_B_C is None else np.asarray(t_B_C, np.float64)
    for i in range(4):
        c.dist[i] = dist[i]
    for i in range(9):
        c.R_B_C[i] = R.reshape(-1)[i]
    for i in range(3):
        c.t_B_C[i] = t[i]
    return c


def gp3p_solve(f, v, p):
    """f, v, p: 3x3 arrays with one column per point. Returns [n][3][4]."""
    fa, va, pa = (_f64(np.asarray(x, np.float64).T.reshape(-1)) for x in (f, v, p))
    out = np.zeros((8, 3, 4), np.float64)
    n = lib().lco_gp3p_solve(_p(fa, c_double_p), _p(va, c_double_p), _p(pa, c_double_p),
                             _p(out, c_double_p))
    return out[:n]


def rng_stream(seed, mapping, n):
    out = np.zeros(n, np.int32)
    lib().lco_rng_stream(C.c_uint32(seed), mapping, n, _p(out, c_int_p))
    return out


def back_project(cam, kps):
    kps = _f64(kps)
    n = kps.shape[0]
    out = np.zeros((n, 3), np.float64)
    lib().lco_back_project(C.byref(cam), _p(kps, c_double_p), n, _p(out, c_double_p))
    return out


def ransac_threshold(cams, pixel_sigma):
    arr = (Camera * len(cams))(*cams)
    return float(lib().lco_ransac_threshold(arr, len(cams), C.c_double(pixel_sigma)))


def handle_loop_closure(keypoints, frame_index, keypoint_index, landmarks, cams,
                        min_inlier_count=10, min_inlier_ratio=0.0, pixel_sigma=2.0, num_iters=100,
                        seed=12345, rng_mapping=1):
    kp = _f64(keypoints)
    n = kp.shape[0]
    fi, ki, lm = _i32(frame_index), _i32(keypoint_index), _f64(landmarks)
    arr = (Camera * len(cams))(*cams)
    sc = np.zeros(10, np.int32)
    ratio = C.c_double(0)
    T = np.zeros((3, 4), np.float64)
    inl = np.zeros(max(n, 1), np.int32)
    inl_d = np.zeros(max(n, 1), np.float64)
    best = np.zeros(max(n, 1), np.int32)
    lib().lco_handle_loop_closure(n, _p(kp, c_double_p), _p(fi, c_int_p), _p(ki, c_int_p),
                                  _p(lm, c_double_p), arr, len(cams), min_inlier_count,
                                  C.c_double(min_inlier_ratio), C.c_double(pixel_sigma), num_iters,
                                  C.c_uint32(seed), rng_mapping, _p(sc, c_int_p), C.byref(ratio),
                                  _p(T, c_double_p), _p(inl, c_int_p), _p(inl_d, c_double_p),
                                  _p(best, c_int_p))
    return dict(accepted=bool(sc[0]), num_inliers=int(sc[1]), ransac_success=bool(sc[2]),
                iterations=int(sc[3]), inliers=inl[:sc[4]].copy(), inlier_distances=inl_d[:sc[4]].copy(),
                model_indices=sc[5:9].copy(), best_per_keypoint=best[:sc[9]].copy(),
                inlier_ratio=ratio.value, T=T)


def delta_pose_gate(T_map, T_ransac, max_pos_m, max_rot_deg):
    """handleLoopClosure's topological gate (loop-closure-handler.cc:424-455).
    Returns (passes, delta_position_m, delta_rotation_deg)."""
    a, b = _f64(np.asarray(T_map).reshape(12)), _f64(np.asarray(T_ransac).reshape(12))
    d = np.zeros(2, np.float64)
    lib().lco_delta_pose_gate.restype = C.c_int
    ok = lib().lco_delta_pose_gate(_p(a, c_double_p), _p(b, c_double_p), C.c_double(max_pos_m),
                                   C.c_double(max_rot_deg), _p(d, c_double_p))
    return bool(ok), float(d[0]), float(d[1])


def transformation_ransac(quats_xyzw, positions, num_iterations, thr_rad, thr_m, seed, rng_mapping=1):
    """common::transformationRansac (geometry-inl.h:113-182). Returns (quat xyzw, position, inlier indices)."""
    q = _f64(np.asarray(quats_xyzw).reshape(-1, 4))
    p = _f64(np.asarray(positions).reshape(-1, 3))
    n = len(q)
    oq, op = np.zeros(4, np.float64), np.zeros(3, np.float64)
    inl = np.zeros(n, np.int32)
    lib().lco_transformation_ransac.restype = C.c_int
    k = lib().lco_transformation_ransac(_p(q, c_double_p), _p(p, c_double_p), n, num_iterations,
                                        C.c_double(thr_rad), C.c_double(thr_m), C.c_uint32(seed), rng_mapping,
                                        _p(oq, c_double_p), _p(op, c_double_p), _p(inl, c_int_p))
    return oq, op, inl[:k].copy()


def uniform_indices(seed, mapping, n, count):
    """libstdc++ uniform_int_distribution<int>(0, n-1) over mt19937(seed): first `count` draws."""
    out = np.zeros(count, np.int32)
    lib().lco_uniform_indices(C.c_uint32(seed), mapping, C.c_uint32(n), count, _p(out, c_int_p))
    return out


def yaw_only(q_xyzw):
    q = _f64(np.asarray(q_xyzw).reshape(4))
    out = np.zeros(4, np.float64)
    lib().lco_yaw_only(_p(q, c_double_p), _p(out, c_double_p))
    return out


# ------------------------------------------------------------------ hnsw engine slot
def exact_knn(db, q, k):
    """Exact k nearest neighbours of float descriptors (ground truth of the `hnsw` engine slot), in the order
    the reference interface returns them: descending (distance, index). Raises if k > len(db)."""
    db, q = _f32(db), _f32(q)
    idx = np.zeros((len(q), k), np.int32)
    dist = np.zeros((len(q), k), np.float32)
    rc = lib().lco_exact_knn(_p(db, c_float_p), C.c_int64(len(db)), _p(q, c_float_p), C.c_int64(len(q)),
                             db.shape[1], k, _p(idx, c_int_p), _p(dist, c_float_p))
    if rc != 0:
        raise ValueError("fewer than k descriptors in the index (the reference CHECKs result.size() == k)")
    return idx, dist
