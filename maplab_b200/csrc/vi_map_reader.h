// Host-side reader of the `vertices<N>` files of a saved vi_map: the message
// vi_map.proto.VIMap with its vertex_ids / vertices fields
//   map-structure/vi-map/proto/vi-map/vi_map.proto:27-43 (ViwlsVertex), :82-100 (Landmark, LandmarkStore), :161-176
//   common/aslam-serialization/proto/aslam-serialization/visual-frame.proto:5-26 (VisualFrame, VisualNFrame)
//   aslam/common/id.proto (Id: repeated uint64)
// as written by vi_map::serialization::serializeVertices (vi-map/src/vi-map-serialization.cc:27-43) and read back
// by deserializeVertices (:107-121). Only what the loop-closure path consumes is kept: ids, poses, per visual
// frame the keypoint measurements, raw descriptors and observed landmark ids, and the landmark stores. The input
// is the serialized message (the file on disk is a gzip stream of it, see mlc_vi_map_count).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace mlc {

struct ViMapVertices {
  // per vertex
  std::vector<uint64_t> vertex_id, mission_id;  // 2 words each (aslam::HashId)
  std::vector<double> T_M_I;                    // 7 each: quaternion x y z w, position
  std::vector<int32_t> vertex_num_frames, vertex_num_landmarks;
  // per visual frame (vertex-major)
  std::vector<int64_t> frame_timestamp_ns;
  std::vector<int32_t> frame_num_keypoints;
  std::vector<uint8_t> frame_is_valid;
  // per keypoint (frame-major)
  std::vector<double> keypoint_measurement;    // 2 each
  std::vector<uint8_t> keypoint_descriptor;    // descriptor_bytes each
  std::vector<uint64_t> keypoint_landmark_id;  // 2 words each
  // per landmark of the landmark stores (vertex-major)
  std::vector<uint64_t> landmark_id;
  std::vector<double> landmark_p_B;  // 3 each, in the frame of the storing vertex
  std::vector<int32_t> landmark_quality;
  int32_t descriptor_bytes = 0;

  int64_t num_vertices() const { return static_cast<int64_t>(vertex_num_frames.size()); }
  int64_t num_frames() const { return static_cast<int64_t>(frame_num_keypoints.size()); }
  int64_t num_keypoints() const { return static_cast<int64_t>(keypoint_measurement.size() / 2); }
  int64_t num_landmarks() const { return static_cast<int64_t>(landmark_quality.size()); }
  bool Parse(const void* proto, size_t size, std::string* err);
};

// The `missions` file of the same folder (vi_map.proto:102-131, :161-176: VIMap.mission_ids / missions /
// mission_base_frame_ids / mission_base_frames): per mission its id and the T_G_M of its base frame
// (7 doubles: quaternion x y z w, position).
struct ViMapMissions {
  std::vector<uint64_t> mission_id;  // 2 words each
  std::vector<double> T_G_M;         // 7 each
  int64_t num_missions() const { return static_cast<int64_t>(mission_id.size() / 2); }
  bool Parse(const void* proto, size_t size, std::string* err);
};

}  // namespace mlc
