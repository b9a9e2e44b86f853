"""Readers for the files either side of the loop-closure path, so that real maplab maps and
the shipped projection / quantizer files can feed the B200 detector without the reference's map stack:

  * `load_projection_matrix`  — projection_matrix_{freak,brisk}.dat: one Eigen matrix written by common::Serialize
                                (maplab-common/binary-serialization.h: int32 rows, int32 cols, column-major floats);
  * `load_vi_map`             — a saved vi_map folder (vi-map/src/vi-map-serialization.cc): gzip'd proto2 files
                                `vertices<N>`, `missions`, plus `sensors.yaml`; read with the protobuf RUNTIME through
                                descriptors declared here (vi-map/proto/vi-map/vi_map.proto, aslam-serialization/
                                proto/aslam-serialization/visual-frame.proto, aslam/common/id.proto — no protoc needed);
  * `loop_closure_inputs`     — what LoopDetectorNode::addVertexToDatabase / queryVertexInDatabase hand to the
                                detector per visual frame (convertFrameToProjectedImage,
                                LCH/src/loop-detector-node.cc:119-200): the keypoints with a valid landmark id whose
                                landmark is well constrained (quality kGood), their raw descriptors, measurements and
                                landmarks (dense numbers), landmark positions in the global frame
                                (T_G_M * T_M_I(storing vertex) * p_B, vi_map::VIMap::getLandmark_G_p), vertex poses and
                                the cameras (mlc_camera fields) of the n-camera rig.
Host-side plumbing only; nothing here touches the device."""
import functools
import glob
import gzip
import os
import struct

import numpy as np


def load_projection_matrix(path):
    raw = open(path, "rb").read()
    rows, cols = struct.unpack_from("<ii", raw, 0)
    if rows <= 0 or cols <= 0 or len(raw) != 8 + 4 * rows * cols:
        raise ValueError(f"{path}: not a serialized float matrix")
    return np.frombuffer(raw, np.float32, rows * cols, 8).reshape(cols, rows).T.copy()


@functools.lru_cache(maxsize=1)
def _vi_map_class():
    """The fields of the three .proto files that the loop-closure inputs need (others are skipped as unknown)."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    F = descriptor_pb2.FieldDescriptorProto
    O, R = F.LABEL_OPTIONAL, F.LABEL_REPEATED

    def fld(msg, name, number, ftype, label, type_name=None):
        f = msg.field.add()
        f.name, f.number, f.type, f.label = name, number, ftype, label
        if type_name:
            f.type_name = type_name

    fd = descriptor_pb2.FileDescriptorProto()
    fd.name, fd.package, fd.syntax = "maplab_b200_vi_map_subset.proto", "mlc_vi_map", "proto2"
    m = fd.message_type.add(); m.name = "Id"                      # aslam.proto.Id
    fld(m, "uint", 1, F.TYPE_UINT64, R)
    m = fd.message_type.add(); m.name = "VisualFrame"             # aslam.proto.VisualFrame
    fld(m, "id", 1, F.TYPE_MESSAGE, O, ".mlc_vi_map.Id")
    fld(m, "timestamp", 2, F.TYPE_INT64, O)
    fld(m, "keypoint_measurements", 3, F.TYPE_DOUBLE, R)
    fld(m, "keypoint_descriptors", 5, F.TYPE_BYTES, O)
    fld(m, "descriptor_types", 11, F.TYPE_INT32, R)
    fld(m, "landmark_ids", 7, F.TYPE_MESSAGE, R, ".mlc_vi_map.Id")
    fld(m, "is_valid", 9, F.TYPE_BOOL, O)
    m = fd.message_type.add(); m.name = "VisualNFrame"            # aslam.proto.VisualNFrame
    fld(m, "frames", 2, F.TYPE_MESSAGE, R, ".mlc_vi_map.VisualFrame")
    m = fd.message_type.add(); m.name = "Landmark"                # vi_map.proto.Landmark
    fld(m, "id", 1, F.TYPE_MESSAGE, O, ".mlc_vi_map.Id")
    fld(m, "position", 2, F.TYPE_DOUBLE, R)
    fld(m, "quality", 7, F.TYPE_INT32, O)
    m = fd.message_type.add(); m.name = "LandmarkStore"
    fld(m, "landmarks", 1, F.TYPE_MESSAGE, R, ".mlc_vi_map.Landmark")
    m = fd.message_type.add(); m.name = "ViwlsVertex"             # vi_map.proto.ViwlsVertex
    fld(m, "T_M_I", 3, F.TYPE_DOUBLE, R)
    fld(m, "n_visual_frame", 7, F.TYPE_MESSAGE, O, ".mlc_vi_map.VisualNFrame")
    fld(m, "landmark_store", 8, F.TYPE_MESSAGE, O, ".mlc_vi_map.LandmarkStore")
    fld(m, "mission_id", 14, F.TYPE_MESSAGE, O, ".mlc_vi_map.Id")
    m = fd.message_type.add(); m.name = "MissionBaseframe"
    fld(m, "T_G_M", 1, F.TYPE_DOUBLE, R)
    m = fd.message_type.add(); m.name = "Mission"
    fld(m, "baseframe_id", 1, F.TYPE_MESSAGE, O, ".mlc_vi_map.Id")
    m = fd.message_type.add(); m.name = "VIMap"                   # vi_map.proto.VIMap
    fld(m, "vertex_ids", 1, F.TYPE_MESSAGE, R, ".mlc_vi_map.Id")
    fld(m, "vertices", 2, F.TYPE_MESSAGE, R, ".mlc_vi_map.ViwlsVertex")
    fld(m, "mission_ids", 5, F.TYPE_MESSAGE, R, ".mlc_vi_map.Id")
    fld(m, "missions", 6, F.TYPE_MESSAGE, R, ".mlc_vi_map.Mission")
    fld(m, "mission_base_frame_ids", 7, F.TYPE_MESSAGE, R, ".mlc_vi_map.Id")
    fld(m, "mission_base_frames", 8, F.TYPE_MESSAGE, R, ".mlc_vi_map.MissionBaseframe")
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("mlc_vi_map.VIMap"))


def transform(q_xyzw_p):
    """eigen_proto::serialize(Transformation) (maplab-common/eigen-proto-inl.h:165-177): quaternion coeffs
    (x, y, z, w), then position -> 4 x 4."""
    x, y, z, w, px, py, pz = q_xyzw_p
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = R, [px, py, pz]
    return T


def descriptors_of(frame):
    """aslam serialises the uchar descriptor matrix as a 24-byte header (…, rows, cols, …) followed by the
    column-major data: one descriptor per column -> [n][bytes]."""
    blob = frame.keypoint_descriptors
    if not blob:
        return np.zeros((0, 0), np.uint8)
    rows, cols = struct.unpack_from("<ii", blob, 8)
    if len(blob) - rows * cols != 24:
        raise ValueError("unexpected descriptor blob layout")
    return np.frombuffer(blob, np.uint8, rows * cols, 24).reshape(cols, rows)


def read_proto_bytes(path):
    """The serialized message of a file written by common::proto_serialization_helper::serializeProtoToFile
    (maplab-common/src/proto-serialization-helper.cc:104-140): a gzip stream with --proto_use_compression (default),
    the plain message otherwise."""
    raw = open(path, "rb").read()
    return gzip.decompress(raw) if raw[:2] == b"\x1f\x8b" else raw


def _read(path, cls):
    msg = cls()
    msg.ParseFromString(read_proto_bytes(path))
    return msg


def load_vi_map(folder):
    """-> dict(vertices [(id, proto)], missions {mission id: T_G_M 4x4}, sensors (parsed yaml or None))."""
    cls = _vi_map_class()
    vertices = []
    files = sorted(glob.glob(os.path.join(folder, "vertices*")), key=lambda p: int(p.rsplit("vertices", 1)[1]))
    if not files:
        raise FileNotFoundError(f"no vertices<N> files under {folder}")
    for path in files:
        m = _read(path, cls)
        if len(m.vertex_ids) != len(m.vertices):
            raise ValueError(f"{path}: vertex_ids / vertices mismatch (CHECK_EQ of deserializeVertices)")
        vertices += [(tuple(i.uint), v) for i, v in zip(m.vertex_ids, m.vertices)]
    mm = _read(os.path.join(folder, "missions"), cls)
    base = {tuple(i.uint): transform(b.T_G_M) for i, b in zip(mm.mission_base_frame_ids, mm.mission_base_frames)}
    missions = {tuple(i.uint): base[tuple(ms.baseframe_id.uint)] for i, ms in zip(mm.mission_ids, mm.missions)}
    sensors = None
    sensors_path = os.path.join(folder, "sensors.yaml")
    if os.path.exists(sensors_path):
        import yaml
        sensors = yaml.safe_load(open(sensors_path))
    return dict(vertices=vertices, missions=missions, sensors=sensors)


DISTORTION = {"none": 0, "fisheye": 1, "radial-tangential": 2, "radtan": 2, "equidistant": 3}


def cameras_of(sensors, camera_indices=None):
    """mlc_camera dicts (capi.make_cameras) of the NCAMERA rig in sensors.yaml."""
    rig = [s for s in sensors["sensors"] if s["sensor_type"] == "NCAMERA"][0]
    cams = []
    for ci, c in enumerate(rig["cameras"]):
        if camera_indices is not None and ci not in camera_indices:
            continue
        cam = c["camera"]
        if cam["type"] != "pinhole":
            raise ValueError("only pinhole cameras are on the GP3P path of this library")
        fu, fv, cu, cv = cam["intrinsics"]["data"]
        dist = cam.get("distortion") or {"type": "none", "parameters": {"data": []}}
        params = list(dist["parameters"]["data"]) + [0.0] * 4
        T = np.array(c["T_B_C"]["data"], np.float64).reshape(4, 4)
        cams.append(dict(fu=fu, fv=fv, cu=cu, cv=cv, distortion=DISTORTION[dist["type"]], dist=tuple(params[:4]),
                         R_B_C=T[:3, :3], t_B_C=T[:3, 3]))
    return cams


def frame_is_usable(fr):
    """isVisualFrameSet && isVisualFrameValid of addVertexToDatabase / queryVertexInDatabase
    (loop-detector-node.cc:279-280, :694-695): a frame without a valid id has been un-set, and a frame is
    valid unless `is_valid` is present and false (visual-frame-serialization.cc:95-96, :166-168)."""
    return any(fr.id.uint) and not (fr.HasField("is_valid") and not fr.is_valid)


def loop_closure_inputs(vi_map, camera_indices=None):
    """Arrays for mlc_insert_batch / mlc_query_batch over all vertices in pose-graph (time) order.
    frames: [F][4] int64 (timestamp_ns, vertex number, frame index within `camera_indices`, descriptors)."""
    vertices = sorted(vi_map["vertices"], key=lambda iv: iv[1].n_visual_frame.frames[0].timestamp)
    mission_numbers = {}
    T_G_I = []
    for _, v in vertices:
        T_G_I.append(vi_map["missions"][tuple(v.mission_id.uint)] @ transform(v.T_M_I))
    T_G_I = np.stack(T_G_I)
    landmark_number, landmark_xyz = {}, []
    for vi, (_, v) in enumerate(vertices):
        for lm in v.landmark_store.landmarks:
            if lm.quality != 2:  # vi_map::Landmark::Quality::kGood == isLandmarkWellConstrained's cached verdict
                continue
            landmark_number[tuple(lm.id.uint)] = len(landmark_xyz)
            landmark_xyz.append((T_G_I[vi] @ np.array(list(lm.position) + [1.0]))[:3])
    frames, missions, bits, keypoints, landmarks = [], [], [], [], []
    for vi, (_, v) in enumerate(vertices):
        mission = mission_numbers.setdefault(tuple(v.mission_id.uint), len(mission_numbers))
        cams = range(len(v.n_visual_frame.frames)) if camera_indices is None else camera_indices
        for slot, ci in enumerate(cams):
            fr = v.n_visual_frame.frames[ci]
            if not frame_is_usable(fr):
                continue
            desc = descriptors_of(fr)
            kp = np.array(fr.keypoint_measurements).reshape(-1, 2)
            if not (len(desc) == len(kp) == len(fr.landmark_ids)):
                raise ValueError("keypoints / descriptors / landmark ids differ in length (CHECK_EQ of "
                                 "convertFrameToProjectedImage)")
            keep = [i for i, l in enumerate(fr.landmark_ids) if tuple(l.uint) in landmark_number]
            frames.append((fr.timestamp, vi, slot, len(keep)))
            missions.append(mission)
            bits.append(desc[keep])
            keypoints.append(kp[keep])
            landmarks += [landmark_number[tuple(fr.landmark_ids[i].uint)] for i in keep]
    width = max((b.shape[1] for b in bits if b.size), default=0)
    bits = [b if b.size else np.zeros((0, width), np.uint8) for b in bits]
    return dict(frames=np.array(frames, np.int64), missions=np.array(missions, np.int64), bits=np.concatenate(bits),
                keypoints=np.concatenate(keypoints), landmarks=np.array(landmarks, np.int64),
                landmark_xyz=np.array(landmark_xyz), T_G_I=T_G_I[:, :3, :],
                vertex_ids=[i for i, _ in vertices])


def load_vertices_native(folder):
    """All `vertices<N>` files of a map folder through the library's own C++ reader (mlc_vi_map_count /
    mlc_vi_map_read, no protobuf runtime): the arrays of capi.vi_map_read_vertices, concatenated."""
    from . import capi
    files = sorted(glob.glob(os.path.join(folder, "vertices*")), key=lambda p: int(p.rsplit("vertices", 1)[1]))
    if not files:
        raise FileNotFoundError(f"no vertices<N> files under {folder}")
    parts = [capi.vi_map_read_vertices(read_proto_bytes(p)) for p in files]
    widths = {p["keypoint_descriptor"].shape[1] for p in parts if p["keypoint_descriptor"].size}
    if len(widths) > 1:
        raise ValueError("descriptor sizes differ between the vertices files")
    width = widths.pop() if widths else 0
    for p in parts:
        if not p["keypoint_descriptor"].size:
            p["keypoint_descriptor"] = np.zeros((0, width), np.uint8)
    return {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}


def load_missions_native(folder):
    """{mission id tuple: T_G_M 4x4} through the library's C++ reader of the `missions` file."""
    from . import capi
    ids, T = capi.vi_map_read_missions(read_proto_bytes(os.path.join(folder, "missions")))
    return {tuple(int(w) for w in i): transform(t) for i, t in zip(ids, T)}


def loop_closure_inputs_native(arrays, missions, camera_indices=None):
    """loop_closure_inputs() on the arrays of load_vertices_native; `missions` = {mission id tuple: T_G_M 4x4}
    (load_vi_map(folder)["missions"]). Same output, same order."""
    V = len(arrays["vertex_num_frames"])
    frame_start = np.concatenate([[0], np.cumsum(arrays["vertex_num_frames"])])
    kp_start = np.concatenate([[0], np.cumsum(arrays["frame_num_keypoints"])])
    lm_start = np.concatenate([[0], np.cumsum(arrays["vertex_num_landmarks"])])
    order = sorted(range(V), key=lambda v: arrays["frame_timestamp_ns"][frame_start[v]])
    T_G_I = np.stack([missions[tuple(int(w) for w in arrays["mission_id"][v])] @ transform(arrays["T_M_I"][v])
                      for v in order])
    landmark_number, landmark_xyz = {}, []
    for vi, v in enumerate(order):
        for l in range(lm_start[v], lm_start[v + 1]):
            if arrays["landmark_quality"][l] != 2:
                continue
            landmark_number[tuple(int(w) for w in arrays["landmark_id"][l])] = len(landmark_xyz)
            landmark_xyz.append((T_G_I[vi] @ np.append(arrays["landmark_p_B"][l], 1.0))[:3])
    mission_numbers = {}
    frames, mission_of_frame, keep_rows, landmarks = [], [], [], []
    for vi, v in enumerate(order):
        mission = mission_numbers.setdefault(tuple(int(w) for w in arrays["mission_id"][v]), len(mission_numbers))
        cams = range(arrays["vertex_num_frames"][v]) if camera_indices is None else camera_indices
        for slot, ci in enumerate(cams):
            f = frame_start[v] + ci
            if not arrays["frame_is_valid"][f]:  # un-set or invalidated frame: not in the database, not queried
                continue
            kept = 0
            for k in range(kp_start[f], kp_start[f + 1]):
                number = landmark_number.get(tuple(int(w) for w in arrays["keypoint_landmark_id"][k]))
                if number is not None:
                    keep_rows.append(k)
                    landmarks.append(number)
                    kept += 1
            frames.append((int(arrays["frame_timestamp_ns"][f]), vi, slot, kept))
            mission_of_frame.append(mission)
    keep_rows = np.array(keep_rows, np.int64)
    return dict(frames=np.array(frames, np.int64), missions=np.array(mission_of_frame, np.int64),
                bits=arrays["keypoint_descriptor"][keep_rows], keypoints=arrays["keypoint_measurement"][keep_rows],
                landmarks=np.array(landmarks, np.int64), landmark_xyz=np.array(landmark_xyz), T_G_I=T_G_I[:, :3, :],
                vertex_ids=[tuple(int(w) for w in arrays["vertex_id"][v]) for v in order])
