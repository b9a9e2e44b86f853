#!/usr/bin/env python
"""Builds tests/golden/real_map_brisk.npz + brisk_quantizer_top10.dat from maplab's own test map
(tools/maplab-test-data/test_maps/common_test_map: 127 vertices of a 5-camera Alphasense rig, real
BRISK descriptors, triangulated landmarks) and its shipped BRISK vocabulary. Run in the build
container, where the reference checkout is mounted; the GPU box only sees the committed outputs.

    python tests/golden/extract_real_map.py

What is kept is what `lc` feeds the loop detector (LoopDetectorNode::addVertexToDatabase /
convertFrameToProjectedImage, LCH/src/loop-detector-node.cc:119-200, :273-339): per visual frame the
keypoints with a valid landmark id whose landmark is well constrained (quality kGood — all of this
map's landmarks are), their raw descriptors, measurements and landmark ids; landmark positions in the
global frame (T_G_M * T_M_I(storing vertex) * p_B, vi_map::VIMap::getLandmark_G_p); vertex poses;
the cameras of sensors.yaml. Only cameras CAMS are kept to bound the fixture size.
The map files are gzip'd proto2 messages (vi-map/proto/vi-map/vi_map.proto,
aslam-serialization/visual-frame.proto); they are read with the protobuf runtime through
descriptors declared below (no protoc in the image).
The quantizer file holds the full 384 x 384 BRISK projection; the path uses its first
lc_target_dimensionality = 10 rows (descriptor-projection.cc:45-49), so only those are stored."""
import gzip
import os
import struct
import sys

import numpy as np
import yaml
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
MAP = "/root/reference/tools/maplab-test-data/test_maps/common_test_map/vi_map/"
QUANTIZER = ("/root/reference/algorithms/loopclosure/matching-based-loopclosure/share/"
             "inverted_multi_index_quantizer_brisk.dat")
CAMS = (0, 2)
F = descriptor_pb2.FieldDescriptorProto


def _fld(msg, name, number, ftype, label, type_name=None):
    f = msg.field.add()
    f.name, f.number, f.type, f.label = name, number, ftype, label
    if type_name:
        f.type_name = type_name


def vi_map_class():
    """The fields of vi_map.proto / visual-frame.proto / id.proto this extraction reads (the rest
    are skipped as unknown fields)."""
    O, R = F.LABEL_OPTIONAL, F.LABEL_REPEATED
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name, fd.package, fd.syntax = "vi_map_subset.proto", "x", "proto2"
    m = fd.message_type.add(); m.name = "Id"
    _fld(m, "uint", 1, F.TYPE_UINT64, R)
    m = fd.message_type.add(); m.name = "VisualFrame"
    _fld(m, "timestamp", 2, F.TYPE_INT64, O)
    _fld(m, "keypoint_measurements", 3, F.TYPE_DOUBLE, R)
    _fld(m, "keypoint_descriptors", 5, F.TYPE_BYTES, O)
    _fld(m, "descriptor_types", 11, F.TYPE_INT32, R)
    _fld(m, "landmark_ids", 7, F.TYPE_MESSAGE, R, ".x.Id")
    _fld(m, "is_valid", 9, F.TYPE_BOOL, O)
    m = fd.message_type.add(); m.name = "VisualNFrame"
    _fld(m, "frames", 2, F.TYPE_MESSAGE, R, ".x.VisualFrame")
    m = fd.message_type.add(); m.name = "Landmark"
    _fld(m, "id", 1, F.TYPE_MESSAGE, O, ".x.Id")
    _fld(m, "position", 2, F.TYPE_DOUBLE, R)
    _fld(m, "quality", 7, F.TYPE_INT32, O)
    m = fd.message_type.add(); m.name = "LandmarkStore"
    _fld(m, "landmarks", 1, F.TYPE_MESSAGE, R, ".x.Landmark")
    m = fd.message_type.add(); m.name = "ViwlsVertex"
    _fld(m, "T_M_I", 3, F.TYPE_DOUBLE, R)
    _fld(m, "n_visual_frame", 7, F.TYPE_MESSAGE, O, ".x.VisualNFrame")
    _fld(m, "landmark_store", 8, F.TYPE_MESSAGE, O, ".x.LandmarkStore")
    _fld(m, "mission_id", 14, F.TYPE_MESSAGE, O, ".x.Id")
    m = fd.message_type.add(); m.name = "MissionBaseframe"
    _fld(m, "T_G_M", 1, F.TYPE_DOUBLE, R)
    m = fd.message_type.add(); m.name = "VIMap"
    _fld(m, "vertex_ids", 1, F.TYPE_MESSAGE, R, ".x.Id")
    _fld(m, "vertices", 2, F.TYPE_MESSAGE, R, ".x.ViwlsVertex")
    _fld(m, "mission_base_frames", 8, F.TYPE_MESSAGE, R, ".x.MissionBaseframe")
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("x.VIMap"))


def transform(q_xyzw_p):
    """eigen_proto::serialize(Transformation): quaternion coeffs (x, y, z, w), then position."""
    x, y, z, w, px, py, pz = q_xyzw_p
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = R, [px, py, pz]
    return T


def descriptors_of(frame):
    """aslam serialises the uchar descriptor matrix as a 24-byte header (…, rows, cols, …) followed
    by the column-major data: one descriptor per column."""
    blob = frame.keypoint_descriptors
    rows, cols = struct.unpack_from("<ii", blob, 8)
    data = np.frombuffer(blob, np.uint8, rows * cols, len(blob) - rows * cols)
    assert len(blob) - rows * cols == 24
    return data.reshape(cols, rows)


def main():
    VIMap = vi_map_class()
    vertices = []
    for name in ("vertices0", "vertices1", "vertices2"):
        m = VIMap()
        m.ParseFromString(gzip.decompress(open(MAP + name, "rb").read()))
        vertices += list(m.vertices)
    m = VIMap()
    m.ParseFromString(gzip.decompress(open(MAP + "missions", "rb").read()))
    T_G_M = transform(m.mission_base_frames[0].T_G_M)
    vertices.sort(key=lambda v: v.n_visual_frame.frames[0].timestamp)  # pose-graph order
    T_G_I = np.stack([T_G_M @ transform(v.T_M_I) for v in vertices])

    landmark_number, landmark_xyz = {}, []
    for vi, v in enumerate(vertices):
        for lm in v.landmark_store.landmarks:
            assert lm.quality == 2  # Landmark::Quality::kGood
            landmark_number[tuple(lm.id.uint)] = len(landmark_xyz)
            landmark_xyz.append((T_G_I[vi] @ np.array(list(lm.position) + [1.0]))[:3])

    frames, bits, keypoints, landmarks = [], [], [], []
    for vi, v in enumerate(vertices):
        for ci in CAMS:
            fr = v.n_visual_frame.frames[ci]
            assert fr.is_valid and set(fr.descriptor_types) == {0}
            desc = descriptors_of(fr)
            kp = np.array(fr.keypoint_measurements).reshape(-1, 2)
            assert len(desc) == len(kp) == len(fr.landmark_ids)
            keep = [i for i, l in enumerate(fr.landmark_ids) if tuple(l.uint) in landmark_number]
            frames.append((fr.timestamp, vi, CAMS.index(ci), len(keep)))
            bits.append(desc[keep])
            keypoints.append(kp[keep])
            landmarks += [landmark_number[tuple(fr.landmark_ids[i].uint)] for i in keep]
    landmarks = np.array(landmarks, np.int64)
    used = np.unique(landmarks)  # renumber to the landmarks the kept cameras observe
    renumber = np.full(len(landmark_xyz), -1, np.int64)
    renumber[used] = np.arange(len(used))

    sensors = yaml.safe_load(open(MAP + "sensors.yaml"))
    rig = [s for s in sensors["sensors"] if s["sensor_type"] == "NCAMERA"][0]
    cams = []
    for ci in CAMS:
        c = rig["cameras"][ci]
        assert c["camera"]["type"] == "pinhole" and c["camera"]["distortion"]["type"] == "equidistant"
        cams.append(np.concatenate([np.array(c["camera"]["intrinsics"]["data"], np.float64),
                                    np.array(c["camera"]["distortion"]["parameters"]["data"], np.float64),
                                    np.array(c["T_B_C"]["data"], np.float64).ravel()]))
    out = os.path.join(HERE, "real_map_brisk.npz")
    np.savez_compressed(
        out, frames=np.array(frames, np.int64), bits=np.concatenate(bits), keypoints=np.concatenate(keypoints),
        landmarks=renumber[landmarks].astype(np.int32), landmark_xyz=np.array(landmark_xyz)[used],
        T_G_I=T_G_I[:, :3, :], cameras=np.stack(cams))
    print(out, os.path.getsize(out), "bytes;", len(frames), "frames,", len(landmarks), "descriptors,", len(used),
          "landmarks")

    from maplab_b200 import synthetic
    v = synthetic.parse_vocabulary(open(QUANTIZER, "rb").read())
    print("BRISK quantizer: version", v["version"], "target_dim", v["target_dim"], "P", v["P"].shape, "W1",
          v["W1"].shape, "W2", v["W2"].shape)
    top = synthetic.serialize_vocabulary(np.ascontiguousarray(v["P"][:v["target_dim"]]), v["W1"], v["W2"],
                                         v["target_dim"], v["version"])
    dst = os.path.join(HERE, "brisk_quantizer_top10.dat")
    open(dst, "wb").write(top)
    print(dst, len(top), "bytes")


if __name__ == "__main__":
    main()
