// maplab_lc_b200_shim.h — header-only C++ shim over the C-ABI with the method names, argument meaning
// and error behaviour of matching_based_loopclosure::LoopDetector
// (algorithms/loopclosure/matching-based-loopclosure/include/matching-based-loopclosure/matching-based-engine.h:18-52)
// and loop_closure::IndexInterface (…/index-interface.h:9-36). Eigen-free (POD views) so that it
// compiles anywhere; the Eigen-typed adapter a maplab maintainer adds on top is in INTEGRATION.md.
// A violated precondition / device error throws std::runtime_error (the reference CHECK-aborts);
// define MAPLAB_LC_B200_ABORT to abort instead.
#ifndef MAPLAB_LC_B200_SHIM_H_
#define MAPLAB_LC_B200_SHIM_H_

#include <cstdint>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "maplab_lc_b200.h"

namespace maplab_lc_b200 {

inline void Check(int rc) {
  if (rc == 0) return;
#ifdef MAPLAB_LC_B200_ABORT
  std::abort();
#else
  throw std::runtime_error(mlc_last_error());
#endif
}

// Dense numbering of 128-bit maplab ids (aslam::HashId: two uint64 words).
struct Id128 {
  uint64_t hi, lo;
  bool operator==(const Id128& o) const { return hi == o.hi && lo == o.lo; }
};
struct Id128Hash {
  size_t operator()(const Id128& id) const { return static_cast<size_t>(id.hi ^ id.lo); }
};
class IdTable {
 public:
  int64_t Number(const Id128& id) {
    auto it = map_.find(id);
    if (it != map_.end()) return it->second;
    const int64_t n = static_cast<int64_t>(ids_.size());
    map_.emplace(id, n);
    ids_.push_back(id);
    return n;
  }
  const Id128& Id(int64_t number) const { return ids_.at(static_cast<size_t>(number)); }
  size_t size() const { return ids_.size(); }

 private:
  std::unordered_map<Id128, int64_t, Id128Hash> map_;
  std::vector<Id128> ids_;
};

// loop_closure::ProjectedImage (descriptor-projection.h:23-31) with dense ids and POD storage.
struct ProjectedImage {
  int64_t timestamp_nanoseconds = 0;
  int64_t vertex_id = 0;  // dense number of keyframe_id.vertex_id
  int32_t frame_index = 0;
  int64_t mission_id = 0;
  std::vector<float> projected_descriptors;  // dim x n column-major == n rows of dim floats
  std::vector<int64_t> landmarks;            // dense landmark numbers (database images)
};

class LoopDetector {
 public:
  LoopDetector(const mlc_settings& settings, const void* quantizer_blob, size_t size) {
    Check(mlc_create(&settings, quantizer_blob, size, &d_));
    dim_ = mlc_target_dim(d_);
  }
  ~LoopDetector() { mlc_destroy(d_); }
  LoopDetector(const LoopDetector&) = delete;
  LoopDetector& operator=(const LoopDetector&) = delete;

  void Initialize() { Check(mlc_initialize(d_)); }
  void Clear() { Check(mlc_clear(d_)); }
  size_t NumEntries() const { return static_cast<size_t>(mlc_num_entries(d_)); }
  int NumDescriptors() const { return static_cast<int>(mlc_num_descriptors(d_)); }
  int dim() const { return dim_; }
  mlc_detector* handle() { return d_; }

  // descriptors: (bytes x n) column-major uchar == n rows of `bytes`; out: dim x n column-major.
  void ProjectDescriptors(const uint8_t* descriptors, int bytes_per_descriptor, int64_t n,
                          float* projected) const {
    Check(mlc_project(d_, descriptors, bytes_per_descriptor, n, projected));
  }

  void Insert(const ProjectedImage& image) {
    const int64_t n = dim_ ? static_cast<int64_t>(image.projected_descriptors.size()) / dim_ : 0;
    if (!image.landmarks.empty() && static_cast<int64_t>(image.landmarks.size()) != n)
      Check(Fail("Insert: projected_descriptors.cols() != landmarks.size()"));
    mlc_frame f{image.timestamp_nanoseconds, image.vertex_id, image.mission_id, image.frame_index,
                static_cast<int32_t>(n)};
    Check(mlc_insert(d_, &f, image.projected_descriptors.data(),
                     image.landmarks.empty() ? nullptr : image.landmarks.data()));
  }

  // All images belong to one vertex (CHECK at matching-based-engine.cc:57). `parallelize_if_possible`
  // is accepted for signature parity; the device path is batched either way.
  void Find(const std::vector<const ProjectedImage*>& images, bool /*parallelize_if_possible*/,
            std::vector<mlc_match>* matches) const {
    matches->clear();
    if (images.empty()) return;
    std::vector<mlc_frame> frames;
    std::vector<float> proj;
    for (const ProjectedImage* im : images) {
      if (im->vertex_id != images[0]->vertex_id) Check(Fail("Find: images of different vertices"));
      const int64_t n = static_cast<int64_t>(im->projected_descriptors.size()) / dim_;
      frames.push_back(mlc_frame{im->timestamp_nanoseconds, im->vertex_id, im->mission_id,
                                 im->frame_index, static_cast<int32_t>(n)});
      proj.insert(proj.end(), im->projected_descriptors.begin(), im->projected_descriptors.end());
    }
    const int64_t cap = static_cast<int64_t>(proj.size() / dim_) * 16 + 16;
    matches->resize(static_cast<size_t>(cap));
    std::vector<int64_t> offsets(frames.size() + 1);
    int64_t nv = 0, nm = 0;
    Check(mlc_find_batch(d_, frames.data(), static_cast<int64_t>(frames.size()), proj.data(),
                         matches->data(), cap, offsets.data(), &nv, &nm));
    matches->resize(static_cast<size_t>(nm));
  }

  // LoopDetectorNode::queryVertexInDatabase for a batch of vertices (loop-detector-node.cc:668-766,
  // :819-873): frames of one vertex adjacent and in ascending camera index; one verdict per vertex.
  // `inlier_matches` (optional) receives, per accepted vertex, the best-per-keypoint inlier matches
  // (the reference's inlier_structure_matches).
  void QueryBatch(const std::vector<mlc_frame>& frames, const uint8_t* descriptors, int bytes_per_descriptor,
                  const double* keypoints_uv, const std::vector<mlc_camera>& cameras,
                  const mlc_ransac_settings& ransac, std::vector<mlc_pose_result>* verdicts,
                  std::vector<std::vector<mlc_match>>* inlier_matches = nullptr) const {
    QueryImpl(&mlc_query_batch, frames, descriptors, bytes_per_descriptor, keypoints_uv, cameras, ransac, verdicts,
              inlier_matches);
  }

  // ---- the database sharded over the GPUs of one box: one process (and one LoopDetector created with
  // shard_rank / shard_count) per GPU ----
  // Rank 0 makes the communicator id and hands the 128 bytes to the other ranks (MPI_Bcast, a file, ...).
  static std::vector<char> MakeCommunicatorId() {
    std::vector<char> id(MLC_COMM_ID_BYTES);
    Check(mlc_comm_unique_id(id.data()));
    return id;
  }
  void InitCommunicator(const std::vector<char>& id) {  // collective over the shard_count ranks
    if (id.size() != MLC_COMM_ID_BYTES) Check(Fail("InitCommunicator: the id is 128 bytes"));
    Check(mlc_comm_init(d_, id.data()));
  }
  // Sharded build: `image` describes ALL descriptors of the keyframe, `owned_projected` holds the rows of the
  // descriptors this shard owns (NumOwnedInRange(NumDescriptors(), n) rows, ascending keypoint index).
  int64_t NumOwnedInRange(int64_t first, int64_t count) const { return mlc_num_owned_in_range(d_, first, count); }
  void InsertOwned(const ProjectedImage& image, int64_t num_keypoints, const float* owned_projected) {
    mlc_frame f{image.timestamp_nanoseconds, image.vertex_id, image.mission_id, image.frame_index,
                static_cast<int32_t>(num_keypoints)};
    Check(mlc_insert_batch_owned(d_, &f, 1, owned_projected, image.landmarks.empty() ? nullptr : image.landmarks.data()));
  }
  // QueryBatch for THIS rank's slice of the step's query vertices against the whole sharded database
  // (collective: every rank calls it once per step, possibly with no frames).
  void ShardedQueryBatch(const std::vector<mlc_frame>& frames, const uint8_t* descriptors, int bytes_per_descriptor,
                         const double* keypoints_uv, const std::vector<mlc_camera>& cameras,
                         const mlc_ransac_settings& ransac, std::vector<mlc_pose_result>* verdicts,
                         std::vector<std::vector<mlc_match>>* inlier_matches = nullptr) const {
    QueryImpl(&mlc_sharded_query_batch, frames, descriptors, bytes_per_descriptor, keypoints_uv, cameras, ransac,
              verdicts, inlier_matches);
  }
  // map_->getVertex_T_G_I of the query vertices of the next QueryBatch (loop-closure-handler.cc:436-437);
  // needed only with --lc_max_delta_position_m / --lc_max_delta_rotation_deg.
  void SetQueryPriors(const double* T_G_I_3x4_rowmajor, int64_t num_vertices) {
    Check(mlc_set_query_priors(d_, T_G_I_3x4_rowmajor, num_vertices));
  }
  void SetLandmarkPositions(const double* xyz, int64_t n) { Check(mlc_set_landmark_positions(d_, xyz, n)); }
  // LoopDetectorNode::addLocalizationSummaryMapToDatabase (LCH/src/loop-detector-node.cc:341-432) on
  // the bytes of the `localization_summary_map` file; dense ids as in mlc_add_summary_map.
  mlc_summary_map_sizes AddLocalizationSummaryMapToDatabase(const void* file_bytes, size_t size, int64_t mission_id,
                                                            int64_t first_vertex_id, int64_t first_landmark_id) {
    mlc_summary_map_sizes sizes{};
    Check(mlc_add_summary_map(d_, file_bytes, size, mission_id, first_vertex_id, first_landmark_id, &sizes));
    return sizes;
  }
  // summary_map::createLocalizationSummaryMapFromLandmarkList + serialize: bytes of the `localization_summary_map` file.
  std::vector<uint8_t> CreateLocalizationSummaryMap(int64_t num_landmarks, const double* G_landmark_position,
                                                    const int64_t* observations_per_landmark, int64_t num_observations,
                                                    const uint8_t* descriptors, int bytes_per_descriptor,
                                                    const int64_t* observer_key, const double* G_observer_position) {
    size_t need = 0;
    mlc_create_summary_map(d_, num_landmarks, G_landmark_position, observations_per_landmark, num_observations,
                           descriptors, bytes_per_descriptor, observer_key, G_observer_position, nullptr, 0, &need);
    if (need == 0) Check(1);
    std::vector<uint8_t> bytes(need);
    Check(mlc_create_summary_map(d_, num_landmarks, G_landmark_position, observations_per_landmark, num_observations,
                                 descriptors, bytes_per_descriptor, observer_key, G_observer_position, bytes.data(),
                                 bytes.size(), &need));
    return bytes;
  }
  void SaveIndex(const std::string& path) { Check(mlc_save_index(d_, path.c_str())); }
  void LoadIndex(const std::string& path) { Check(mlc_load_index(d_, path.c_str())); }
  // common::transformationRansac (geometry-inl.h:113-182); returns the number of inliers.
  int TransformationRansac(const double* quats_xyzw, const double* positions, int64_t n,
                           const mlc_alignment_settings& settings, double out_quat_xyzw[4], double out_position[3],
                           std::vector<int32_t>* inlier_indices = nullptr) const {
    std::vector<int32_t> tmp(static_cast<size_t>(n > 0 ? n : 1));
    int32_t num = 0;
    Check(mlc_transformation_ransac(d_, quats_xyzw, positions, n, &settings, out_quat_xyzw, out_position,
                                    tmp.data(), &num));
    if (inlier_indices) inlier_indices->assign(tmp.begin(), tmp.begin() + num);
    return num;
  }

  // Tail of detectLoopClosuresMissionToDatabase (loop-detector-node.cc:936-955): inlier gate, yaw-only projection.
  static bool EnoughAlignmentInliers(int num_inliers, int64_t num_samples, int min_inlier_count = 10,
                                     double min_inlier_ratio = 0.2) {
    return mlc_alignment_enough_inliers(num_inliers, num_samples, min_inlier_count, min_inlier_ratio) != 0;
  }
  static void YawOnly(const double quat_xyzw[4], double out_quat_xyzw[4]) {
    Check(mlc_alignment_yaw_only(quat_xyzw, out_quat_xyzw));
  }

  // loop_closure::IndexInterface::GetNNearestNeighborsForFeatures (index-interface.h:27-35).
  void GetNNearestNeighborsForFeatures(const float* query_features, int64_t n, int num_neighbors,
                                       int32_t* indices, float* distances) const {
    Check(mlc_knn(d_, query_features, n, num_neighbors, indices, distances));
  }

 private:
  using QueryFn = int (*)(mlc_detector*, const mlc_frame*, int64_t, const uint8_t*, int, const double*,
                          const mlc_camera*, int, const mlc_ransac_settings*, mlc_pose_result*, int64_t*, mlc_match*,
                          int64_t, int64_t*, int64_t*, uint8_t*);
  void QueryImpl(QueryFn fn, const std::vector<mlc_frame>& frames, const uint8_t* descriptors,
                 int bytes_per_descriptor, const double* keypoints_uv, const std::vector<mlc_camera>& cameras,
                 const mlc_ransac_settings& ransac, std::vector<mlc_pose_result>* verdicts,
                 std::vector<std::vector<mlc_match>>* inlier_matches) const {
    int64_t n = 0;
    for (const mlc_frame& f : frames) n += f.num_descriptors;
    verdicts->assign(frames.size() + 1, mlc_pose_result{});
    const int64_t cap = n * 16 + 16;
    std::vector<mlc_match> matches(inlier_matches ? static_cast<size_t>(cap) : 0);
    std::vector<uint8_t> flags(inlier_matches ? static_cast<size_t>(cap) : 0);
    std::vector<int64_t> offsets(frames.size() + 1, 0);
    int64_t nv = 0, nm = 0;
    Check(fn(d_, frames.data(), static_cast<int64_t>(frames.size()), descriptors, bytes_per_descriptor, keypoints_uv,
             cameras.data(), static_cast<int>(cameras.size()), &ransac, verdicts->data(), &nv,
             inlier_matches ? matches.data() : nullptr, cap, offsets.data(), &nm,
             inlier_matches ? flags.data() : nullptr));
    verdicts->resize(static_cast<size_t>(nv));
    if (inlier_matches) {
      inlier_matches->assign(static_cast<size_t>(nv), {});
      for (int64_t v = 0; v < nv; ++v) {
        if (!(*verdicts)[static_cast<size_t>(v)].accepted) continue;
        for (int64_t i = offsets[v]; i < offsets[v + 1]; ++i)
          if (flags[static_cast<size_t>(i)] == 3)
            (*inlier_matches)[static_cast<size_t>(v)].push_back(matches[static_cast<size_t>(i)]);
      }
    }
  }
  static int Fail(const char* msg) {
    last_shim_error() = msg;
    throw std::runtime_error(msg);
  }
  static std::string& last_shim_error() {
    static thread_local std::string e;
    return e;
  }
  mlc_detector* d_ = nullptr;
  int dim_ = 0;
};

}  // namespace maplab_lc_b200
#endif  // MAPLAB_LC_B200_SHIM_H_
