// ORACLE (test infrastructure only; see lc_oracle.h): CPU restatement of
//   LoopDetectorNode::addLocalizationSummaryMapToDatabase
//     algorithms/loopclosure/loop-closure-handler/src/loop-detector-node.cc:341-432
// on the arrays a deserialized summary_map::LocalizationSummaryMap holds
// (map-structure/localization-summary-map/src/localization-summary-map.cc:53-93). The reference
// decodes the file with libprotobuf; the tests decode it with the real google.protobuf runtime and
// hand the arrays in, so the product's own wire decoder is checked against libprotobuf's.
// Ids: the reference draws random 128-bit vertex / mission ids and hash-seeded landmark ids; here,
// as everywhere at this boundary, they are dense numbers handed out by the caller.
#include <cstdint>
#include <vector>

#include "lc_oracle.h"

using namespace lc_oracle;

extern "C" int lco_engine_add_summary_map(void* h, int dim, int64_t num_observers, int64_t num_landmarks,
                                          int64_t num_observations, int64_t descriptor_cols,
                                          const float* projected_descriptors /* dim x cols col-major */,
                                          const uint32_t* observer_indices,
                                          const uint32_t* observation_to_landmark_index,
                                          int64_t num_observation_to_landmark, int64_t mission_id,
                                          int64_t first_vertex_id, int64_t first_landmark_id) {
  LoopDetector* detector = static_cast<LoopDetector*>(h);
  if (num_observers == 0) return 1;  // LOG(FATAL) "No observers in the summary map found."
  // :368-380 accumulate the observation indices per observer
  std::vector<std::vector<int>> observer_observations(static_cast<size_t>(num_observers));
  for (int64_t i = 0; i < num_observations; ++i) {
    const int observer_index = static_cast<int>(observer_indices[i]);
    if (!(observer_index < static_cast<int>(observer_observations.size()))) return 2;  // CHECK_LT
    observer_observations[observer_index].push_back(static_cast<int>(i));
  }
  // :398-424 one ProjectedImage per observer
  for (size_t observer_idx = 0; observer_idx < observer_observations.size(); ++observer_idx) {
    ProjectedImage image;
    image.timestamp_ns = 0;
    image.mission_id = mission_id;
    image.vertex_id = first_vertex_id + static_cast<int64_t>(observer_idx);
    image.frame_index = 0;  // kFrameIndex
    image.dim = dim;
    const std::vector<int>& observations = observer_observations[observer_idx];
    image.landmarks.resize(observations.size());
    image.projected_descriptors.resize(static_cast<size_t>(dim) * observations.size());
    for (size_t i = 0; i < observations.size(); ++i) {
      const int observation_index = observations[i];
      if (!(observation_index < descriptor_cols)) return 3;  // CHECK_LT
      for (int r = 0; r < dim; ++r)
        image.projected_descriptors[i * dim + r] =
            projected_descriptors[static_cast<size_t>(observation_index) * dim + r];
      if (!(observation_index < num_observation_to_landmark)) return 4;  // CHECK_LT
      const size_t landmark_index = observation_to_landmark_index[observation_index];
      if (!(landmark_index < static_cast<size_t>(num_landmarks))) return 5;  // CHECK_LT
      image.landmarks[i] = first_landmark_id + static_cast<int64_t>(landmark_index);
    }
    detector->Insert(image);
  }
  return 0;  // loop_detector_->Initialize() is a no-op for the CPU engines
}
