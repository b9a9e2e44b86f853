"""ctypes binding of libmaplab_lc_b200.so (include/maplab_lc_b200.h).

The library is built in-tree by ``maplab_b200.build.build()`` (``__graft_entry__.build``). There
is no CPU fallback: if the shared library or a CUDA device is missing every call fails loudly.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmaplab_lc_b200.so")
_lib = None

EXPORTS = [
    "mlc_last_error", "mlc_version", "mlc_kernel_launch_count", "mlc_default_settings",
    "mlc_default_ransac_settings", "mlc_create", "mlc_destroy", "mlc_clear", "mlc_num_entries",
    "mlc_num_descriptors", "mlc_num_neighbors", "mlc_target_dim", "mlc_project",
    "mlc_project_device", "mlc_insert", "mlc_insert_batch", "mlc_insert_batch_owned",
    "mlc_insert_batch_device", "mlc_num_owned_in_range", "mlc_comm_unique_id", "mlc_comm_init",
    "mlc_comm_destroy", "mlc_comm_nccl_version", "mlc_sharded_query_batch", "mlc_sharded_query_batch_device",
    "mlc_sharded_knn_device", "mlc_initialize", "mlc_knn",
    "mlc_knn_device", "mlc_coarse_cells", "mlc_merge_topk_device", "mlc_last_scan_stats",
    "mlc_find_batch", "mlc_find_batch_bits", "mlc_find_from_knn_device", "mlc_pnp_ransac_batch",
    "mlc_coarse_device", "mlc_scan_device", "mlc_last_stage_ms", "mlc_set_landmark_positions", "mlc_set_landmark_positions_device", "mlc_query_batch", "mlc_query_batch_device", "mlc_query_from_knn_device",
    "mlc_score", "mlc_save_index", "mlc_load_index", "mlc_set_query_priors",
    "mlc_default_alignment_settings", "mlc_transformation_ransac",
    "mlc_summary_map_parse", "mlc_summary_map_serialize", "mlc_add_summary_map", "mlc_create_summary_map",
    "mlc_vi_map_count", "mlc_vi_map_read", "mlc_vi_map_missions", "mlc_alignment_enough_inliers", "mlc_alignment_yaw_only",
]


class Settings(C.Structure):
    _fields_ = [("num_closest_words", C.c_int32), ("num_nearest_neighbors", C.c_int32),
                ("scoring", C.c_int32), ("engine", C.c_int32),
                ("min_image_time_seconds", C.c_double), ("min_verify_matches_num", C.c_uint64),
                ("fraction_best_scores", C.c_float), ("knn_epsilon", C.c_float),
                ("knn_max_radius", C.c_float), ("device", C.c_int32), ("shard_rank", C.c_int32),
                ("shard_count", C.c_int32), ("shard_mode", C.c_int32), ("float_descriptor_dim", C.c_int32),
                ("hnsw_m", C.c_int32), ("hnsw_ef_construction", C.c_int32), ("hnsw_ef_query", C.c_int32),
                ("pad_", C.c_int32)]


class Frame(C.Structure):
    _fields_ = [("timestamp_ns", C.c_int64), ("vertex_id", C.c_int64), ("mission_id", C.c_int64),
                ("frame_index", C.c_int32), ("num_descriptors", C.c_int32)]


FRAME_DTYPE = np.dtype([("timestamp_ns", "<i8"), ("vertex_id", "<i8"), ("mission_id", "<i8"),
                        ("frame_index", "<i4"), ("num_descriptors", "<i4")])
MATCH_DTYPE = np.dtype([("query_frame", "<i4"), ("query_keypoint", "<i4"), ("db_descriptor", "<i4"),
                        ("db_keyframe", "<i4"), ("db_vertex", "<i8"), ("landmark", "<i8")])
CAMERA_DTYPE = np.dtype([("fu", "<f8"), ("fv", "<f8"), ("cu", "<f8"), ("cv", "<f8"),
                         ("distortion", "<i4"), ("pad_", "<i4"), ("dist", "<f8", (4,)),
                         ("R_B_C", "<f8", (9,)), ("t_B_C", "<f8", (3,))])
POSE_DTYPE = np.dtype([("accepted", "<i4"), ("ransac_success", "<i4"), ("num_inliers", "<i4"),
                       ("num_ransac_inliers", "<i4"), ("iterations", "<i4"),
                       ("model_indices", "<i4", (4,)), ("pad_", "<i4"), ("inlier_ratio", "<f8"),
                       ("T_G_I", "<f8", (12,))])


class RansacSettings(C.Structure):
    _fields_ = [("min_inlier_count", C.c_int32), ("num_ransac_iters", C.c_int32),
                ("min_inlier_ratio", C.c_double), ("ransac_pixel_sigma", C.c_double),
                ("seed", C.c_uint32), ("rng_mapping", C.c_int32),
                ("max_delta_position_m", C.c_double), ("max_delta_rotation_deg", C.c_double)]


class AlignmentSettings(C.Structure):
    _fields_ = [("num_iterations", C.c_int32), ("rng_mapping", C.c_int32),
                ("max_orientation_error_rad", C.c_double), ("max_position_error_m", C.c_double),
                ("seed", C.c_uint32), ("pad_", C.c_uint32)]


class SummaryMapSizes(C.Structure):
    _fields_ = [("num_landmarks", C.c_int64), ("num_observers", C.c_int64), ("num_observations", C.c_int64),
                ("descriptor_rows", C.c_int64), ("descriptor_cols", C.c_int64)]


class ViMapCounts(C.Structure):
    _fields_ = [("num_vertices", C.c_int64), ("num_frames", C.c_int64), ("num_keypoints", C.c_int64),
                ("num_landmarks", C.c_int64), ("descriptor_bytes", C.c_int32), ("pad_", C.c_int32)]


class ViMapArrays(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "vertex_id", "mission_id", "T_M_I", "vertex_num_frames", "vertex_num_landmarks", "frame_timestamp_ns",
        "frame_num_keypoints", "frame_is_valid", "keypoint_measurement", "keypoint_descriptor",
        "keypoint_landmark_id", "landmark_id", "landmark_p_B", "landmark_quality")]


class MlcError(RuntimeError):
    pass


def lib():
    """Load the shared library (raises if it has not been built — no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MlcError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; "
                           "g.build()'` (needs nvcc); there is no CPU fallback")
        _lib = C.CDLL(LIB_PATH)
        _lib.mlc_last_error.restype = C.c_char_p
        _lib.mlc_kernel_launch_count.restype = C.c_uint64
        _lib.mlc_num_entries.restype = C.c_int64
        _lib.mlc_num_descriptors.restype = C.c_int64
        _lib.mlc_num_owned_in_range.restype = C.c_int64
        _lib.mlc_destroy.restype = None
        _lib.mlc_default_settings.restype = None
        _lib.mlc_default_ransac_settings.restype = None
    return _lib


def _check(rc):
    if rc != 0:
        raise MlcError(lib().mlc_last_error().decode())


def _ptr(a):
    return C.c_void_p(a.ctypes.data) if a is not None and a.size else C.c_void_p(0)


def default_settings(**kw):
    s = Settings()
    lib().mlc_default_settings(C.byref(s))
    for k, v in kw.items():
        if not hasattr(s, k):
            raise AttributeError(k)
        setattr(s, k, v)
    return s


def default_ransac_settings(**kw):
    s = RansacSettings()
    lib().mlc_default_ransac_settings(C.byref(s))
    for k, v in kw.items():
        if not hasattr(s, k):
            raise AttributeError(k)
        setattr(s, k, v)
    return s


def summary_map_parse(blob):
    """LocalizationSummaryMap::deserialize on the serialized proto (host only, no device needed).
    Matrices come back in Eigen's shape: G_landmark_position 3 x L, descriptors rows x cols."""
    blob = bytes(blob)
    sz = SummaryMapSizes()
    null = C.c_void_p(0)
    _check(lib().mlc_summary_map_parse(blob, C.c_size_t(len(blob)), C.byref(sz), null, null, null, null, null))
    lm = np.zeros((sz.num_landmarks, 3), np.float32)
    ob = np.zeros((sz.num_observers, 3), np.float32)
    desc = np.zeros((sz.descriptor_cols, sz.descriptor_rows), np.float32)
    oi = np.zeros(sz.num_observations, np.uint32)
    ol = np.zeros(sz.num_observations, np.uint32)
    _check(lib().mlc_summary_map_parse(blob, C.c_size_t(len(blob)), C.byref(sz), _ptr(lm), _ptr(ob), _ptr(desc),
                                       _ptr(oi), _ptr(ol)))
    return {"G_landmark_position": lm.T, "G_observer_position": ob.T, "descriptors": desc.T,
            "observer_indices": oi, "observation_to_landmark_index": ol}


def summary_map_serialize(G_landmark_position, G_observer_position, descriptors, observer_indices,
                          observation_to_landmark_index):
    """LocalizationSummaryMap::serialize: the bytes of the `localization_summary_map` file.
    G_landmark_position 3 x L, G_observer_position 3 x O, descriptors dim x N (Eigen shapes)."""
    lm = np.ascontiguousarray(np.asarray(G_landmark_position, np.float32).reshape(3, -1).T)
    ob = np.ascontiguousarray(np.asarray(G_observer_position, np.float32).reshape(3, -1).T)
    d = np.asarray(descriptors, np.float32)
    desc = np.ascontiguousarray(d.T)
    oi = np.ascontiguousarray(observer_indices, np.uint32)
    ol = np.ascontiguousarray(observation_to_landmark_index, np.uint32)
    assert len(oi) == len(ol)
    sz = SummaryMapSizes(len(lm), len(ob), len(oi), d.shape[0], d.shape[1])
    need = C.c_size_t(0)
    lib().mlc_summary_map_serialize(C.byref(sz), _ptr(lm), _ptr(ob), _ptr(desc), _ptr(oi), _ptr(ol),
                                    C.c_void_p(0), C.c_size_t(0), C.byref(need))
    out = np.zeros(max(need.value, 1), np.uint8)
    _check(lib().mlc_summary_map_serialize(C.byref(sz), _ptr(lm), _ptr(ob), _ptr(desc), _ptr(oi), _ptr(ol),
                                           _ptr(out), C.c_size_t(need.value), C.byref(need)))
    return out[:need.value].tobytes()


def vi_map_read_vertices(proto_bytes):
    """One `vertices<N>` message of a saved vi_map (already inflated, vi_map_io.read_proto_bytes) -> dict of
    arrays (mlc_vi_map_arrays). Host only."""
    blob = bytes(proto_bytes)
    c = ViMapCounts()
    _check(lib().mlc_vi_map_count(blob, C.c_size_t(len(blob)), C.byref(c)))
    V, F, K, L, B = c.num_vertices, c.num_frames, c.num_keypoints, c.num_landmarks, c.descriptor_bytes
    out = dict(vertex_id=np.zeros((V, 2), np.uint64), mission_id=np.zeros((V, 2), np.uint64),
               T_M_I=np.zeros((V, 7), np.float64), vertex_num_frames=np.zeros(V, np.int32),
               vertex_num_landmarks=np.zeros(V, np.int32), frame_timestamp_ns=np.zeros(F, np.int64),
               frame_num_keypoints=np.zeros(F, np.int32), frame_is_valid=np.zeros(F, np.uint8),
               keypoint_measurement=np.zeros((K, 2), np.float64), keypoint_descriptor=np.zeros((K, B), np.uint8),
               keypoint_landmark_id=np.zeros((K, 2), np.uint64), landmark_id=np.zeros((L, 2), np.uint64),
               landmark_p_B=np.zeros((L, 3), np.float64), landmark_quality=np.zeros(L, np.int32))
    arrays = ViMapArrays(**{k: (v.ctypes.data if v.size else None) for k, v in out.items()})
    _check(lib().mlc_vi_map_read(blob, C.c_size_t(len(blob)), C.byref(c), C.byref(arrays)))
    return out


def vi_map_read_missions(proto_bytes):
    """The `missions` message of a saved vi_map -> (mission ids [M][2] uint64, T_G_M [M][7])."""
    blob = bytes(proto_bytes)
    n = C.c_int64(0)
    _check(lib().mlc_vi_map_missions(blob, C.c_size_t(len(blob)), C.c_int64(0), C.c_void_p(0), C.c_void_p(0),
                                     C.byref(n)))
    ids, T = np.zeros((n.value, 2), np.uint64), np.zeros((n.value, 7), np.float64)
    _check(lib().mlc_vi_map_missions(blob, C.c_size_t(len(blob)), C.c_int64(n.value), _ptr(ids), _ptr(T), C.byref(n)))
    return ids, T


def alignment_yaw_only(quat_xyzw):
    """Yaw-only projection of the mission alignment (loop-detector-node.cc:944-955), host only."""
    q = np.ascontiguousarray(quat_xyzw, np.float64).reshape(4)
    out = np.zeros(4, np.float64)
    _check(lib().mlc_alignment_yaw_only(q.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)))
    return out


def alignment_enough_inliers(num_inliers, num_samples, min_inlier_count=10, min_inlier_ratio=0.2):
    return bool(lib().mlc_alignment_enough_inliers(C.c_int32(num_inliers), C.c_int64(num_samples),
                                                   C.c_int32(min_inlier_count), C.c_double(min_inlier_ratio)))


def comm_unique_id():
    """ncclGetUniqueId through the library (rank 0; hand the 128 bytes to the other ranks)."""
    buf = C.create_string_buffer(128)
    _check(lib().mlc_comm_unique_id(buf))
    return buf.raw


def kernel_launch_count():
    return int(lib().mlc_kernel_launch_count())


def make_frames(ts, vertex, mission, frame_index, num_descriptors):
    n = len(num_descriptors)
    f = np.zeros(n, FRAME_DTYPE)
    f["timestamp_ns"] = ts
    f["vertex_id"] = vertex
    f["mission_id"] = mission
    f["frame_index"] = frame_index
    f["num_descriptors"] = num_descriptors
    return f


def make_cameras(cams):
    """cams: list of dicts(fu, fv, cu, cv, R_B_C(3x3), t_B_C(3), distortion, dist)."""
    out = np.zeros(len(cams), CAMERA_DTYPE)
    for i, c in enumerate(cams):
        out[i]["fu"], out[i]["fv"], out[i]["cu"], out[i]["cv"] = c["fu"], c["fv"], c["cu"], c["cv"]
        out[i]["distortion"] = c.get("distortion", 0)
        out[i]["dist"] = np.asarray(c.get("dist", (0, 0, 0, 0)), np.float64)
        out[i]["R_B_C"] = np.asarray(c.get("R_B_C", np.eye(3)), np.float64).reshape(-1)
        out[i]["t_B_C"] = np.asarray(c.get("t_B_C", np.zeros(3)), np.float64)
    return out


class Detector:
    """Thin owner of an ``mlc_detector*``; numpy in, numpy out."""

    def __init__(self, vocab_blob, settings=None):
        self.settings = settings or default_settings()
        blob = bytes(vocab_blob) if vocab_blob is not None else b""
        self._h = C.c_void_p()
        _check(lib().mlc_create(C.byref(self.settings), blob, C.c_size_t(len(blob)),
                                C.byref(self._h)))
        self.dim = int(lib().mlc_target_dim(self._h))

    def close(self):
        if getattr(self, "_h", None):
            lib().mlc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- LoopDetector surface -------------------------------------------------
    def clear(self):
        _check(lib().mlc_clear(self._h))

    def num_entries(self):
        return int(lib().mlc_num_entries(self._h))

    def num_descriptors(self):
        return int(lib().mlc_num_descriptors(self._h))

    def num_neighbors(self):
        return int(lib().mlc_num_neighbors(self._h))

    def project(self, bits):
        bits = np.ascontiguousarray(bits, np.uint8)
        n, nbytes = bits.shape
        out = np.empty((n, self.dim), np.float32)
        _check(lib().mlc_project(self._h, _ptr(bits), nbytes, C.c_int64(n), _ptr(out)))
        return out

    def insert_batch(self, frames, proj, landmarks=None):
        frames = np.ascontiguousarray(frames, FRAME_DTYPE)
        proj = np.ascontiguousarray(proj, np.float32).reshape(-1, self.dim)
        assert int(frames["num_descriptors"].sum()) == proj.shape[0]
        lm = None if landmarks is None else np.ascontiguousarray(landmarks, np.int64)
        _check(lib().mlc_insert_batch(self._h, _ptr(frames), C.c_int64(len(frames)), _ptr(proj),
                                      _ptr(lm) if lm is not None else C.c_void_p(0)))

    def num_owned_in_range(self, first, count):
        return int(lib().mlc_num_owned_in_range(self._h, C.c_int64(first), C.c_int64(count)))

    def insert_batch_owned(self, frames, proj_owned, landmarks=None):
        """Sharded build: rows of the descriptors this shard owns only (ascending global index)."""
        frames = np.ascontiguousarray(frames, FRAME_DTYPE)
        proj = np.ascontiguousarray(proj_owned, np.float32).reshape(-1, self.dim)
        total = int(frames["num_descriptors"].sum())
        assert self.num_owned_in_range(self.num_descriptors(), total) == proj.shape[0]
        lm = None if landmarks is None else np.ascontiguousarray(landmarks, np.int64)
        _check(lib().mlc_insert_batch_owned(self._h, _ptr(frames), C.c_int64(len(frames)), _ptr(proj),
                                            _ptr(lm) if lm is not None else C.c_void_p(0)))

    def insert_batch_device(self, frames, proj_owned_ptr, num_owned, landmarks_ptr=0, stream=0):
        frames = np.ascontiguousarray(frames, FRAME_DTYPE)
        _check(lib().mlc_insert_batch_device(self._h, _ptr(frames), C.c_int64(len(frames)),
                                             C.c_void_p(proj_owned_ptr), C.c_int64(num_owned),
                                             C.c_void_p(landmarks_ptr), C.c_void_p(stream)))

    def insert(self, ts, vertex, frame_index, mission, proj, landmarks=None):
        proj = np.ascontiguousarray(proj, np.float32).reshape(-1, self.dim)
        self.insert_batch(make_frames([ts], [vertex], [mission], [frame_index], [proj.shape[0]]),
                          proj, landmarks)

    def initialize(self):
        _check(lib().mlc_initialize(self._h))

    def knn(self, q, k):
        q = np.ascontiguousarray(q, np.float32).reshape(-1, self.dim)
        n = q.shape[0]
        idx = np.empty((n, k), np.int32)
        dist = np.empty((n, k), np.float32)
        _check(lib().mlc_knn(self._h, _ptr(q), C.c_int64(n), k, _ptr(idx), _ptr(dist)))
        return idx, dist

    def coarse_cells(self, q, nw):
        q = np.ascontiguousarray(q, np.float32).reshape(-1, self.dim)
        cells = np.empty((q.shape[0], nw), np.int32)
        _check(lib().mlc_coarse_cells(self._h, _ptr(q), C.c_int64(q.shape[0]), nw, _ptr(cells)))
        return cells

    def last_scan_stats(self):
        b, e, ms = C.c_uint64(), C.c_uint64(), C.c_double()
        _check(lib().mlc_last_scan_stats(self._h, C.byref(b), C.byref(e), C.byref(ms)))
        return dict(algorithmic_bytes=b.value, entries=e.value, scan_ms=ms.value)

    # device-pointer variants (plumbing by torch: pass tensor.data_ptr())
    def project_device(self, bits_ptr, bytes_per_desc, n, out_ptr, stream=0):
        _check(lib().mlc_project_device(self._h, C.c_void_p(bits_ptr), bytes_per_desc, C.c_int64(n),
                                        C.c_void_p(out_ptr), C.c_void_p(stream)))

    def knn_device(self, q_ptr, n, k, idx_ptr, dist_ptr, stream=0):
        _check(lib().mlc_knn_device(self._h, C.c_void_p(q_ptr), C.c_int64(n), k, C.c_void_p(idx_ptr),
                                    C.c_void_p(dist_ptr), C.c_void_p(stream)))

    def coarse_device(self, q_ptr, n, nw, cells_ptr, stream=0):
        _check(lib().mlc_coarse_device(self._h, C.c_void_p(q_ptr), C.c_int64(n), nw, C.c_void_p(cells_ptr),
                                       C.c_void_p(stream)))

    def scan_device(self, q_ptr, cells_ptr, n, k, idx_ptr, dist_ptr, stream=0):
        _check(lib().mlc_scan_device(self._h, C.c_void_p(q_ptr), C.c_void_p(cells_ptr), C.c_int64(n), k,
                                     C.c_void_p(idx_ptr), C.c_void_p(dist_ptr), C.c_void_p(stream)))

    def last_stage_ms(self):
        ms = (C.c_double * 5)()
        _check(lib().mlc_last_stage_ms(self._h, ms))
        return dict(zip(("project", "coarse", "scan", "vote_cluster", "ransac"), [float(x) for x in ms]))

    def set_query_priors(self, T_G_I):
        """Current poses of the query vertices of the next query call (delta-pose gate)."""
        t = np.ascontiguousarray(T_G_I, np.float64).reshape(-1, 12)
        _check(lib().mlc_set_query_priors(self._h, t.ctypes.data_as(C.c_void_p), C.c_int64(len(t))))

    def transformation_ransac(self, quats_xyzw, positions, **kw):
        """common::transformationRansac (geometry-inl.h:113-182) on the device.
        Returns (quaternion xyzw, position, ascending inlier indices)."""
        s = AlignmentSettings()
        lib().mlc_default_alignment_settings(C.byref(s))
        for k, v in kw.items():
            if not hasattr(s, k):
                raise AttributeError(k)
            setattr(s, k, v)
        q = np.ascontiguousarray(quats_xyzw, np.float64).reshape(-1, 4)
        p = np.ascontiguousarray(positions, np.float64).reshape(-1, 3)
        oq, op = np.zeros(4, np.float64), np.zeros(3, np.float64)
        inl = np.zeros(max(len(q), 1), np.int32)
        cnt = C.c_int32(0)
        _check(lib().mlc_transformation_ransac(self._h, q.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p),
                                               C.c_int64(len(q)), C.byref(s), oq.ctypes.data_as(C.c_void_p),
                                               op.ctypes.data_as(C.c_void_p), inl.ctypes.data_as(C.c_void_p),
                                               C.byref(cnt)))
        return oq, op, inl[:cnt.value].copy()

    def save_index(self, path):
        _check(lib().mlc_save_index(self._h, str(path).encode()))

    def load_index(self, path):
        _check(lib().mlc_load_index(self._h, str(path).encode()))

    def score(self, num_matches, num_descriptors, num_db, probabilistic):
        """scoring::compute*Score on the device (scoring.h:38-59, :92-187)."""
        m = np.ascontiguousarray(num_matches, np.uint64)
        d = np.ascontiguousarray(num_descriptors, np.uint64)
        out = np.zeros(len(m), np.float32)
        _check(lib().mlc_score(self._h, 1 if probabilistic else 0, m.ctypes.data_as(C.c_void_p),
                               d.ctypes.data_as(C.c_void_p), len(m), C.c_int64(num_db),
                               out.ctypes.data_as(C.c_void_p)))
        return out if num_db > 0 else out[:0]

    def merge_topk_device(self, idx_lists_ptr, dist_lists_ptr, num_lists, n, k, idx_ptr, dist_ptr,
                          stream=0):
        _check(lib().mlc_merge_topk_device(self._h, C.c_void_p(idx_lists_ptr),
                                           C.c_void_p(dist_lists_ptr), num_lists, C.c_int64(n), k,
                                           C.c_void_p(idx_ptr), C.c_void_p(dist_ptr),
                                           C.c_void_p(stream)))

    def find_batch(self, frames, proj=None, bits=None, capacity=None):
        """Returns (matches[MATCH_DTYPE], offsets[num_vertices+1])."""
        frames = np.ascontiguousarray(frames, FRAME_DTYPE)
        nf = len(frames)
        total = int(frames["num_descriptors"].sum())
        k = max(self.num_neighbors(), 1)
        cap = capacity if capacity is not None else total * k + 16
        matches = np.zeros(cap, MATCH_DTYPE)
        offsets = np.zeros(nf + 1, np.int64)
        nv, nm = C.c_int64(), C.c_int64()
        if bits is not None:
            bits = np.ascontiguousarray(bits, np.uint8)
            assert bits.shape[0] == total
            _check(lib().mlc_find_batch_bits(self._h, _ptr(frames), C.c_int64(nf), _ptr(bits),
                                             bits.shape[1], _ptr(matches), C.c_int64(cap),
                                             _ptr(offsets), C.byref(nv), C.byref(nm)))
        else:
            proj = np.ascontiguousarray(proj, np.float32).reshape(-1, self.dim)
            assert proj.shape[0] == total
            _check(lib().mlc_find_batch(self._h, _ptr(frames), C.c_int64(nf), _ptr(proj),
                                        _ptr(matches), C.c_int64(cap), _ptr(offsets), C.byref(nv),
                                        C.byref(nm)))
        if nm.value > cap:
            raise MlcError(f"match capacity {cap} too small for {nm.value} matches")
        return matches[:nm.value], offsets[:nv.value + 1]

    def find_from_knn_device(self, frames, idx_ptr, dist_ptr, k, capacity=None):
        frames = np.ascontiguousarray(frames, FRAME_DTYPE)
        nf = len(frames)
        total = int(frames["num_descriptors"].sum())
        cap = capacity if capacity is not None else total * k + 16
        matches = np.zeros(cap, MATCH_DTYPE)
        offsets = np.zeros(nf + 1, np.int64)
        nv, nm = C.c_int64(), C.c_int64()
        _check(lib().mlc_find_from_knn_device(self._h, _ptr(frames), C.c_int64(nf), C.c_void_p(idx_ptr),
                                              C.c_void_p(dist_ptr), k, _ptr(matches), C.c_int64(cap),
                                              _ptr(offsets), C.byref(nv), C.byref(nm)))
        return matches[:nm.value], offsets[:nv.value + 1]

    def create_summary_map(self, G_landmark_position, observations_per_landmark, bits, observer_key,
                           G_observer_position):
        """createLocalizationSummaryMapFromLandmarkList + serialize: the `localization_summary_map` file
        bytes for landmark-major observations (G_landmark_position [L][3], bits [N][bytes],
        observer_key [N], G_observer_position [N][3])."""
        lm = np.ascontiguousarray(G_landmark_position, np.float64).reshape(-1, 3)
        cnt = np.ascontiguousarray(observations_per_landmark, np.int64)
        bits = np.ascontiguousarray(bits, np.uint8)
        key = np.ascontiguousarray(observer_key, np.int64)
        pos = np.ascontiguousarray(G_observer_position, np.float64).reshape(-1, 3)
        assert len(cnt) == len(lm) and len(key) == len(bits) == len(pos)
        need = C.c_size_t(0)
        args = (self._h, C.c_int64(len(lm)), _ptr(lm), _ptr(cnt), C.c_int64(len(bits)), _ptr(bits),
                C.c_int(bits.shape[1] if bits.ndim == 2 else 0), _ptr(key), _ptr(pos))
        lib().mlc_create_summary_map(*args, C.c_void_p(0), C.c_size_t(0), C.byref(need))
        if need.value == 0:
            raise MlcError(lib().mlc_last_error().decode())
        out = np.zeros(need.value, np.uint8)
        _check(lib().mlc_create_summary_map(*args, _ptr(out), C.c_size_t(need.value), C.byref(need)))
        return out.tobytes()

    def add_summary_map(self, blob, mission_id, first_vertex_id, first_landmark_id):
        """LoopDetectorNode::addLocalizationSummaryMapToDatabase on the serialized proto."""
        blob = bytes(blob)
        sz = SummaryMapSizes()
        _check(lib().mlc_add_summary_map(self._h, blob, C.c_size_t(len(blob)), C.c_int64(mission_id),
                                         C.c_int64(first_vertex_id), C.c_int64(first_landmark_id), C.byref(sz)))
        return {k: int(getattr(sz, k)) for k, _ in SummaryMapSizes._fields_}

    def set_landmark_positions(self, xyz):
        xyz = np.ascontiguousarray(xyz, np.float64).reshape(-1, 3)
        _check(lib().mlc_set_landmark_positions(self._h, _ptr(xyz), C.c_int64(len(xyz))))

    def set_landmark_positions_device(self, xyz_ptr, n):
        _check(lib().mlc_set_landmark_positions_device(self._h, C.c_void_p(xyz_ptr), C.c_int64(n)))

    def _query(self, fn, frames, a0, a1, a2, cams, rs, want_matches, want_flags, extra=(), k=None):
        frames = np.ascontiguousarray(frames, FRAME_DTYPE)
        cams = np.ascontiguousarray(cams, CAMERA_DTYPE)
        rs = rs or default_ransac_settings()
        nf = len(frames)
        total = int(frames["num_descriptors"].sum())
        k = max(k if k is not None else self.num_neighbors(), 1)  # the buffers must hold total * k matches
        res = np.zeros(max(nf, 1), POSE_DTYPE)
        nv, nm = C.c_int64(), C.c_int64()
        cap = total * k + 16 if want_matches else 0
        matches = np.zeros(cap, MATCH_DTYPE) if want_matches else None
        offsets = np.zeros(nf + 1, np.int64)
        flags = np.zeros(total * k + 16, np.uint8) if want_flags else None
        _check(fn(self._h, _ptr(frames), C.c_int64(nf), a0, a1, a2, *extra, _ptr(cams), len(cams),
                  C.byref(rs), _ptr(res), C.byref(nv),
                  _ptr(matches) if matches is not None else C.c_void_p(0), C.c_int64(cap),
                  _ptr(offsets), C.byref(nm), _ptr(flags) if flags is not None else C.c_void_p(0)))
        out = dict(results=res[:nv.value], offsets=offsets[:nv.value + 1], num_matches=nm.value)
        if want_matches:
            out["matches"] = matches[:nm.value]
        if want_flags:
            out["inlier_flags"] = flags[:nm.value]
        return out

    def query_batch(self, frames, bits, keypoints, cams, rs=None, want_matches=False, want_flags=False):
        """Fused host-buffer query (the e2e call): returns dict(results, offsets, num_matches, ...)."""
        bits = np.ascontiguousarray(bits, np.uint8)
        kp = np.ascontiguousarray(keypoints, np.float64).reshape(-1, 2)
        assert len(bits) == len(kp)
        return self._query(lib().mlc_query_batch, frames, _ptr(bits), bits.shape[1], _ptr(kp), cams, rs,
                           want_matches, want_flags)

    def query_batch_device(self, frames, bits_ptr, bytes_per_desc, keypoints_ptr, cams, rs=None,
                           want_matches=False, want_flags=False):
        return self._query(lib().mlc_query_batch_device, frames, C.c_void_p(bits_ptr), bytes_per_desc,
                           C.c_void_p(keypoints_ptr), cams, rs, want_matches, want_flags)

    # -- multi-GPU (one process per GPU; see include/maplab_lc_b200.h "Multi-GPU") ------------
    def comm_init(self, id128):
        id128 = bytes(id128)
        assert len(id128) == 128
        _check(lib().mlc_comm_init(self._h, id128))

    def comm_init_torch(self, group=None):
        """Communicator over the ranks of a torch.distributed group (plumbing only: the 128-byte NCCL
        id travels by a torch broadcast, everything after that is the library's own NCCL traffic)."""
        import torch
        import torch.distributed as dist
        rank = dist.get_rank(group)
        buf = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = torch.frombuffer(bytearray(comm_unique_id()), dtype=torch.uint8).clone()
        if dist.get_backend(group) == "nccl":
            buf = buf.cuda()
        dist.broadcast(buf, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        self.comm_init(buf.cpu().numpy().tobytes())

    def comm_destroy(self):
        _check(lib().mlc_comm_destroy(self._h))

    def comm_nccl_version(self):
        return int(lib().mlc_comm_nccl_version(self._h))

    def sharded_query_batch(self, frames, bits, keypoints, cams, rs=None, want_matches=False, want_flags=False):
        bits = np.ascontiguousarray(bits, np.uint8)
        kp = np.ascontiguousarray(keypoints, np.float64).reshape(-1, 2)
        assert len(bits) == len(kp)
        return self._query(lib().mlc_sharded_query_batch, frames, _ptr(bits), bits.shape[1], _ptr(kp), cams, rs,
                           want_matches, want_flags)

    def sharded_query_batch_device(self, frames, bits_ptr, bytes_per_desc, keypoints_ptr, cams, rs=None,
                                   want_matches=False, want_flags=False):
        return self._query(lib().mlc_sharded_query_batch_device, frames, C.c_void_p(bits_ptr), bytes_per_desc,
                           C.c_void_p(keypoints_ptr), cams, rs, want_matches, want_flags)

    def sharded_knn_device(self, q_ptr, n, k, idx_ptr, dist_ptr):
        _check(lib().mlc_sharded_knn_device(self._h, C.c_void_p(q_ptr), C.c_int64(n), k, C.c_void_p(idx_ptr),
                                            C.c_void_p(dist_ptr)))

    def query_from_knn_device(self, frames, idx_ptr, dist_ptr, k, keypoints_ptr, cams, rs=None,
                              want_matches=False, want_flags=False):
        return self._query(lib().mlc_query_from_knn_device, frames, C.c_void_p(idx_ptr),
                           C.c_void_p(dist_ptr), k, cams, rs, want_matches, want_flags,
                           extra=(C.c_void_p(keypoints_ptr),), k=k)

    def pnp_ransac_batch(self, cams, offsets, keypoints, camera_index, keypoint_index, landmarks,
                         rs=None, want_flags=True):
        rs = rs or default_ransac_settings()
        cams = np.ascontiguousarray(cams, CAMERA_DTYPE)
        offsets = np.ascontiguousarray(offsets, np.int64)
        nprob = len(offsets) - 1
        kp = np.ascontiguousarray(keypoints, np.float64).reshape(-1, 2)
        ci = np.ascontiguousarray(camera_index, np.int32)
        ki = np.ascontiguousarray(keypoint_index, np.int32)
        lm = np.ascontiguousarray(landmarks, np.float64).reshape(-1, 3)
        res = np.zeros(nprob, POSE_DTYPE)
        flags = np.zeros(max(len(ci), 1), np.uint8) if want_flags else None
        _check(lib().mlc_pnp_ransac_batch(self._h, C.byref(rs), _ptr(cams), len(cams),
                                          C.c_int64(nprob), _ptr(offsets), _ptr(kp), _ptr(ci),
                                          _ptr(ki), _ptr(lm), _ptr(res),
                                          _ptr(flags) if flags is not None else C.c_void_p(0)))
        return res, (flags[:len(ci)] if flags is not None else None)
