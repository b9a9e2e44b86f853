"""The two proto2 messages of maplab's localization summary map, declared to the REAL protobuf
runtime (google.protobuf, present in the image; protoc is not) so that the tests can check the
library's own wire decoder / encoder against libprotobuf's:
  common/maplab-common/proto/maplab-common/eigen.proto:8-12        (common.proto.MatrixXf)
  map-structure/localization-summary-map/proto/localization-summary-map/
      localization-summary-map.proto:4-14                          (summary_map.proto.*)
`packed=True` declares the repeated scalars [packed=true] (what a proto3-era writer would emit):
readers must accept both encodings."""
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

F = descriptor_pb2.FieldDescriptorProto


def _field(msg, name, number, ftype, label, type_name=None, packed=False):
    f = msg.field.add()
    f.name, f.number, f.type, f.label = name, number, ftype, label
    if type_name:
        f.type_name = type_name
    if packed:
        f.options.packed = True


def messages(packed=False, suffix=""):
    pool = descriptor_pool.DescriptorPool()
    eigen = descriptor_pb2.FileDescriptorProto()
    eigen.name = "maplab-common/eigen%s.proto" % suffix
    eigen.package = "common.proto"
    eigen.syntax = "proto2"
    m = eigen.message_type.add()
    m.name = "MatrixXf"
    _field(m, "rows", 1, F.TYPE_UINT32, F.LABEL_OPTIONAL)
    _field(m, "cols", 2, F.TYPE_UINT32, F.LABEL_OPTIONAL)
    _field(m, "data", 3, F.TYPE_FLOAT, F.LABEL_REPEATED, packed=packed)
    pool.Add(eigen)

    sm = descriptor_pb2.FileDescriptorProto()
    sm.name = "localization-summary-map/localization-summary-map%s.proto" % suffix
    sm.package = "summary_map.proto"
    sm.syntax = "proto2"
    sm.dependency.append(eigen.name)
    u = sm.message_type.add()
    u.name = "UncompressedLocalizationSummaryMap"
    _field(u, "descriptors", 1, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, ".common.proto.MatrixXf")
    _field(u, "G_observer_position", 2, F.TYPE_FLOAT, F.LABEL_REPEATED, packed=packed)
    _field(u, "observer_indices", 3, F.TYPE_UINT32, F.LABEL_REPEATED, packed=packed)
    _field(u, "observation_to_landmark_index", 4, F.TYPE_UINT32, F.LABEL_REPEATED, packed=packed)
    t = sm.message_type.add()
    t.name = "LocalizationSummaryMap"
    _field(t, "G_landmark_position", 1, F.TYPE_FLOAT, F.LABEL_REPEATED, packed=packed)
    _field(t, "uncompressed_map", 2, F.TYPE_MESSAGE, F.LABEL_OPTIONAL,
           ".summary_map.proto.UncompressedLocalizationSummaryMap")
    pool.Add(sm)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("summary_map.proto.LocalizationSummaryMap"))


def encode(G_landmark_position, G_observer_position, descriptors, observer_indices,
           observation_to_landmark_index, packed=False):
    """LocalizationSummaryMap::serialize (src/localization-summary-map.cc:33-52) through libprotobuf:
    eigen_proto::serialize writes column-major data and always sets rows / cols."""
    import numpy as np
    msg = messages(packed)()
    msg.G_landmark_position.extend(np.asarray(G_landmark_position, np.float32).T.ravel().tolist())
    u = msg.uncompressed_map
    u.SetInParent()
    d = np.asarray(descriptors, np.float32)
    u.descriptors.rows, u.descriptors.cols = d.shape
    u.descriptors.data.extend(d.T.ravel().tolist())
    u.G_observer_position.extend(np.asarray(G_observer_position, np.float32).T.ravel().tolist())
    u.observer_indices.extend(int(x) for x in observer_indices)
    u.observation_to_landmark_index.extend(int(x) for x in observation_to_landmark_index)
    return msg.SerializeToString(deterministic=True)


def decode(blob, packed=False):
    import numpy as np
    msg = messages(packed)()
    msg.ParseFromString(bytes(blob))
    u = msg.uncompressed_map
    return {
        "G_landmark_position": np.array(msg.G_landmark_position, np.float32).reshape(-1, 3).T,
        "G_observer_position": np.array(u.G_observer_position, np.float32).reshape(-1, 3).T,
        "descriptors": np.array(u.descriptors.data, np.float32).reshape(u.descriptors.cols, u.descriptors.rows).T,
        "observer_indices": np.array(u.observer_indices, np.uint32),
        "observation_to_landmark_index": np.array(u.observation_to_landmark_index, np.uint32),
    }
