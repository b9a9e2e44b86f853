// Host-side reader / writer of maplab's localization summary map file and the grouping step that
// feeds it to the database.
//   format:   map-structure/localization-summary-map/proto/localization-summary-map/
//             localization-summary-map.proto:4-14 (+ common.proto.MatrixXf, maplab-common/proto/
//             maplab-common/eigen.proto:8-12), written / read by LocalizationSummaryMap::serialize /
//             deserialize (map-structure/localization-summary-map/src/localization-summary-map.cc:33-93)
//   consumer: LoopDetectorNode::addLocalizationSummaryMapToDatabase
//             (algorithms/loopclosure/loop-closure-handler/src/loop-detector-node.cc:341-432)
// There is no protoc / libprotobuf in the build: the proto2 wire format of these two messages is
// decoded directly (varint / 32-bit / length-delimited fields, packed and unpacked repeated scalars,
// unknown fields skipped, last-one-wins / merge semantics for repeated occurrences).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace mlc {

struct SummaryMap {
  std::vector<float> G_landmark_position;  // 3 x L column-major (LocalizationSummaryMap field 1)
  bool has_uncompressed_map = false;       // field 2
  uint32_t descriptor_rows = 0, descriptor_cols = 0;  // MatrixXf rows / cols
  std::vector<float> descriptors;                      // rows x cols column-major (MatrixXf data)
  std::vector<float> G_observer_position;              // 3 x O column-major
  std::vector<uint32_t> observer_indices;              // per observation
  std::vector<uint32_t> observation_to_landmark_index; // per observation

  int64_t num_landmarks() const { return static_cast<int64_t>(G_landmark_position.size() / 3); }
  int64_t num_observers() const { return static_cast<int64_t>(G_observer_position.size() / 3); }
  int64_t num_observations() const { return static_cast<int64_t>(observer_indices.size()); }

  // Decode + the CHECKs of eigen_proto::deserialize (eigen-proto-inl.h:23-36, :75-85) and of
  // LocalizationSummaryMap::deserialize ("Unsupported localization summary map format").
  bool Parse(const void* blob, size_t size, std::string* err);
  // Encode exactly as libprotobuf writes these proto2 messages: fields in number order, repeated
  // scalars unpacked (no [packed=true] in the .proto), optional rows / cols always set by
  // eigen_proto::serialize (eigen-proto-inl.h:99-111).
  void Serialize(std::vector<uint8_t>* out) const;
  size_t SerializedSize() const;  // without building the bytes (descriptor VALUES do not change the size)
};

// addLocalizationSummaryMapToDatabase's regrouping (loop-detector-node.cc:368-424): observation i
// goes to observer observer_indices[i], observations keep their order inside an observer; one
// ProjectedImage per observer (also for observers without observations), descriptor column =
// projected_descriptors.col(observation), landmark = landmark index of the observation.
struct SummaryMapImages {
  std::vector<int32_t> num_descriptors;  // per observer
  std::vector<float> proj;               // concatenated, dim floats per descriptor
  std::vector<int64_t> landmark_index;   // concatenated, index into G_landmark_position
};
bool GroupSummaryMapByObserver(const SummaryMap& map, SummaryMapImages* out, std::string* err);

}  // namespace mlc
