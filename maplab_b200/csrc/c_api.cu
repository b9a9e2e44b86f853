// extern "C" surface of libmaplab_lc_b200.so (include/maplab_lc_b200.h).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/maplab_lc_b200.h"
#include "detector.h"
#include "summary_map.h"
#include "vi_map_reader.h"

struct mlc_detector {
  mlc::Detector impl;
};

namespace {
thread_local std::string g_last_error;
int Fail(const std::string& msg) {
  g_last_error = msg;
  return 1;
}
#define MLC_REQUIRE(cond, msg) \
  do {                         \
    if (!(cond)) return Fail(msg); \
  } while (0)
}  // namespace

extern "C" {

const char* mlc_last_error(void) { return g_last_error.c_str(); }
int mlc_version(void) { return 100; }
uint64_t mlc_kernel_launch_count(void) { return mlc::g_kernel_launches.load(); }

void mlc_default_settings(mlc_settings* s) {
  s->num_closest_words = 10;
  s->num_nearest_neighbors = -1;
  s->scoring = 0;
  s->engine = 0;
  s->min_image_time_seconds = 10.0;
  s->min_verify_matches_num = 10;
  s->fraction_best_scores = 0.25f;
  s->knn_epsilon = 2.0f;
  s->knn_max_radius = 20.0f;
  s->device = -1;
  s->shard_rank = 0;
  s->shard_count = 1;
  s->shard_mode = 0;
  s->float_descriptor_dim = 0;
  s->hnsw_m = 12;
  s->hnsw_ef_construction = 50;
  s->hnsw_ef_query = 50;
  s->pad_ = 0;
}

void mlc_default_ransac_settings(mlc_ransac_settings* s) {
  s->min_inlier_count = 10;
  s->num_ransac_iters = 100;
  s->min_inlier_ratio = 0.0;
  s->ransac_pixel_sigma = 2.0;
  s->seed = 12345u;
  s->rng_mapping = 1;
  s->max_delta_position_m = -1.0;
  s->max_delta_rotation_deg = -1.0;
}

int mlc_create(const mlc_settings* settings, const void* vocab_blob, size_t vocab_size,
               mlc_detector** out) {
  MLC_REQUIRE(settings && out && (vocab_blob || settings->engine == 2), "mlc_create: null argument");
  *out = nullptr;
  mlc_detector* d = new (std::nothrow) mlc_detector();
  MLC_REQUIRE(d, "out of memory");
  std::string err;
  if (!d->impl.Create(*settings, vocab_blob, vocab_size, &err)) {
    delete d;
    return Fail(err);
  }
  *out = d;
  return 0;
}

void mlc_destroy(mlc_detector* d) { delete d; }

int mlc_clear(mlc_detector* d) {
  MLC_REQUIRE(d, "null detector");
  std::string err;
  return d->impl.Clear(&err) ? 0 : Fail(err);
}
int64_t mlc_num_entries(const mlc_detector* d) { return d ? d->impl.NumEntries() : -1; }
int64_t mlc_num_descriptors(const mlc_detector* d) { return d ? d->impl.NumDescriptors() : -1; }
int mlc_num_neighbors(const mlc_detector* d) { return d ? d->impl.NumNeighbors() : -1; }
int mlc_target_dim(const mlc_detector* d) { return d ? d->impl.dim() : -1; }

int mlc_project(mlc_detector* d, const uint8_t* bits, int bytes_per_desc, int64_t n, float* out) {
  MLC_REQUIRE(d && (n == 0 || (bits && out)), "mlc_project: null argument");
  std::string err;
  return d->impl.Project(bits, bytes_per_desc, n, out, &err) ? 0 : Fail(err);
}
int mlc_project_device(mlc_detector* d, const uint8_t* d_bits, int bytes_per_desc, int64_t n,
                       float* d_out, void* stream) {
  MLC_REQUIRE(d && (n == 0 || (d_bits && d_out)), "mlc_project_device: null argument");
  MLC_REQUIRE(bytes_per_desc > 0 && (bytes_per_desc % 16 == 0 || d->impl.exact_engine()),
              "bytes per descriptor must be a multiple of 16");
  std::string err;
  return d->impl.ProjectDevice(d_bits, bytes_per_desc, n, d_out, static_cast<cudaStream_t>(stream), &err)
             ? 0
             : Fail(err);
}

int mlc_insert(mlc_detector* d, const mlc_frame* frame, const float* proj, const int64_t* landmarks) {
  return mlc_insert_batch(d, frame, 1, proj, landmarks);
}
int mlc_insert_batch(mlc_detector* d, const mlc_frame* frames, int64_t num_frames, const float* proj,
                     const int64_t* landmarks) {
  MLC_REQUIRE(d && frames && num_frames >= 0, "mlc_insert: null argument");
  std::string err;
  return d->impl.InsertBatch(frames, num_frames, proj, landmarks, false, &err) ? 0 : Fail(err);
}
int mlc_insert_batch_owned(mlc_detector* d, const mlc_frame* frames, int64_t num_frames, const float* proj_owned,
                           const int64_t* landmarks) {
  MLC_REQUIRE(d && frames && num_frames >= 0, "mlc_insert_batch_owned: null argument");
  std::string err;
  return d->impl.InsertBatch(frames, num_frames, proj_owned, landmarks, true, &err) ? 0 : Fail(err);
}
int mlc_insert_batch_device(mlc_detector* d, const mlc_frame* frames, int64_t num_frames, const float* d_proj_owned,
                            int64_t num_owned, const int64_t* d_landmarks, void* stream) {
  MLC_REQUIRE(d && frames && num_frames >= 0, "mlc_insert_batch_device: null argument");
  std::string err;
  return d->impl.InsertBatchDevice(frames, num_frames, d_proj_owned, num_owned, d_landmarks,
                                   static_cast<cudaStream_t>(stream), &err)
             ? 0
             : Fail(err);
}
int64_t mlc_num_owned_in_range(const mlc_detector* d, int64_t first, int64_t count) {
  return d ? d->impl.OwnedInRange(first, count) : -1;
}
int mlc_comm_unique_id(void* id128) {
  MLC_REQUIRE(id128, "mlc_comm_unique_id: null argument");
  std::string err;
  return mlc::CommUniqueId(id128, &err) ? 0 : Fail(err);
}
int mlc_comm_init(mlc_detector* d, const void* id128) {
  MLC_REQUIRE(d && id128, "mlc_comm_init: null argument");
  std::string err;
  return d->impl.CommInit(id128, &err) ? 0 : Fail(err);
}
int mlc_comm_destroy(mlc_detector* d) {
  MLC_REQUIRE(d, "null detector");
  d->impl.CommDestroy();
  return 0;
}
int mlc_comm_nccl_version(const mlc_detector* d) { return d ? d->impl.CommVersion() : -1; }
int mlc_sharded_knn_device(mlc_detector* d, const float* d_q, int64_t n_q, int k, int32_t* d_idx, float* d_dist) {
  MLC_REQUIRE(d && (n_q == 0 || (d_q && d_idx && d_dist)), "mlc_sharded_knn_device: null argument");
  std::string err;
  return d->impl.ShardedKnnDevice(d_q, n_q, k, d_idx, d_dist, &err) ? 0 : Fail(err);
}
static int ShardedQuery(mlc_detector* d, const mlc_frame* frames, int64_t num_frames, const uint8_t* bits,
                        int bytes_per_desc, const double* keypoints, bool on_device, const mlc_camera* cams,
                        int num_cams, const mlc_ransac_settings* rs, mlc_pose_result* results,
                        int64_t* num_vertices, mlc_match* matches, int64_t capacity, int64_t* match_offsets,
                        int64_t* num_matches, uint8_t* inlier_flags) {
  MLC_REQUIRE(d && num_vertices && cams && num_cams > 0 && (num_frames == 0 || (frames && results)),
              "mlc_sharded_query_batch: null argument");
  MLC_REQUIRE(bytes_per_desc > 0 && bytes_per_desc % 16 == 0, "bytes per descriptor must be a multiple of 16");
  mlc_ransac_settings def;
  mlc_default_ransac_settings(&def);
  std::string err;
  return d->impl.ShardedQueryBatch(frames, num_frames, bits, bytes_per_desc, keypoints, on_device, cams, num_cams,
                                   rs ? *rs : def, results, num_vertices, matches, capacity, match_offsets,
                                   num_matches, inlier_flags, &err)
             ? 0
             : Fail(err);
}
int mlc_sharded_query_batch(mlc_detector* d, const mlc_frame* frames, int64_t num_frames, const uint8_t* bits,
                            int bytes_per_desc, const double* keypoints, const mlc_camera* cams, int num_cams,
                            const mlc_ransac_settings* rs, mlc_pose_result* results, int64_t* num_vertices,
                            mlc_match* matches, int64_t capacity, int64_t* match_offsets, int64_t* num_matches,
                            uint8_t* inlier_flags) {
  return ShardedQuery(d, frames, num_frames, bits, bytes_per_desc, keypoints, false, cams, num_cams, rs, results,
                      num_vertices, matches, capacity, match_offsets, num_matches, inlier_flags);
}
int mlc_sharded_query_batch_device(mlc_detector* d, const mlc_frame* frames, int64_t num_frames,
                                   const uint8_t* d_bits, int bytes_per_desc, const double* d_keypoints,
                                   const mlc_camera* cams, int num_cams, const mlc_ransac_settings* rs,
                                   mlc_pose_result* results, int64_t* num_vertices, mlc_match* matches,
                                   int64_t capacity, int64_t* match_offsets, int64_t* num_matches,
                                   uint8_t* inlier_flags) {
  return ShardedQuery(d, frames, num_frames, d_bits, bytes_per_desc, d_keypoints, true, cams, num_cams, rs, results,
                      num_vertices, matches, capacity, match_offsets, num_matches, inlier_flags);
}
int mlc_initialize(mlc_detector* d) {
  MLC_REQUIRE(d, "null detector");
  std::string err;
  return d->impl.Initialize(&err) ? 0 : Fail(err);
}

int mlc_knn(mlc_detector* d, const float* q, int64_t n_q, int k, int32_t* idx, float* dist) {
  MLC_REQUIRE(d && (n_q == 0 || (q && idx && dist)), "mlc_knn: null argument");
  std::string err;
  return d->impl.Knn(q, n_q, k, idx, dist, &err) ? 0 : Fail(err);
}
int mlc_knn_device(mlc_detector* d, const float* d_q, int64_t n_q, int k, int32_t* d_idx,
                   float* d_dist, void* stream) {
  MLC_REQUIRE(d && (n_q == 0 || (d_q && d_idx && d_dist)), "mlc_knn_device: null argument");
  std::string err;
  return d->impl.KnnDevice(d_q, n_q, k, d_idx, d_dist, static_cast<cudaStream_t>(stream), &err)
             ? 0
             : Fail(err);
}
int mlc_coarse_cells(mlc_detector* d, const float* q, int64_t n, int nw, int32_t* cells) {
  MLC_REQUIRE(d && (n == 0 || (q && cells)), "mlc_coarse_cells: null argument");
  std::string err;
  return d->impl.CoarseCells(q, n, nw, cells, &err) ? 0 : Fail(err);
}
int mlc_coarse_device(mlc_detector* d, const float* d_q, int64_t n, int nw, int32_t* d_cells, void* stream) {
  MLC_REQUIRE(d && (n == 0 || (d_q && d_cells)), "mlc_coarse_device: null argument");
  std::string err;
  return d->impl.CoarseDevice(d_q, n, nw, d_cells, static_cast<cudaStream_t>(stream), &err) ? 0 : Fail(err);
}
int mlc_scan_device(mlc_detector* d, const float* d_q, const int32_t* d_cells, int64_t n_q, int k,
                    int32_t* d_idx, float* d_dist, void* stream) {
  MLC_REQUIRE(d && (n_q == 0 || (d_q && d_cells && d_idx && d_dist)), "mlc_scan_device: null argument");
  std::string err;
  return d->impl.ScanDevice(d_q, d_cells, n_q, k, d_idx, d_dist, static_cast<cudaStream_t>(stream), &err)
             ? 0
             : Fail(err);
}
int mlc_last_stage_ms(mlc_detector* d, double* ms5) {
  MLC_REQUIRE(d && ms5, "null argument");
  std::string err;
  return d->impl.LastStageMs(ms5, &err) ? 0 : Fail(err);
}
int mlc_merge_topk_device(mlc_detector* d, const int32_t* d_idx_lists, const float* d_dist_lists,
                          int num_lists, int64_t n_q, int k, int32_t* d_idx, float* d_dist,
                          void* stream) {
  MLC_REQUIRE(d && d_idx_lists && d_dist_lists && d_idx && d_dist, "mlc_merge_topk_device: null argument");
  std::string err;
  return d->impl.MergeTopkDevice(d_idx_lists, d_dist_lists, num_lists, n_q, k, d_idx, d_dist,
                                 static_cast<cudaStream_t>(stream), &err)
             ? 0
             : Fail(err);
}
int mlc_score(mlc_detector* d, int scoring, const uint64_t* num_matches, const uint64_t* num_descriptors,
              int n, int64_t num_db_descriptors, float* scores) {
  MLC_REQUIRE(d && (n == 0 || (num_matches && num_descriptors && scores)), "mlc_score: null argument");
  std::string err;
  return d->impl.Score(scoring, num_matches, num_descriptors, n, num_db_descriptors, scores, &err) ? 0 : Fail(err);
}
int mlc_last_scan_stats(mlc_detector* d, uint64_t* algorithmic_bytes, uint64_t* entries_scanned,
                        double* scan_kernel_ms) {
  MLC_REQUIRE(d && algorithmic_bytes && entries_scanned && scan_kernel_ms, "null argument");
  std::string err;
  return d->impl.LastScanStats(algorithmic_bytes, entries_scanned, scan_kernel_ms, &err) ? 0 : Fail(err);
}

int mlc_find_batch(mlc_detector* d, const mlc_frame* frames, int64_t num_frames, const float* proj,
                   mlc_match* matches, int64_t capacity, int64_t* match_offsets,
                   int64_t* num_vertices, int64_t* num_matches) {
  MLC_REQUIRE(d && num_vertices && num_matches && (num_frames == 0 || (frames && match_offsets)),
              "mlc_find_batch: null argument");
  std::string err;
  return d->impl.FindBatch(frames, num_frames, proj, nullptr, 0, matches, capacity, match_offsets,
                           num_vertices, num_matches, &err)
             ? 0
             : Fail(err);
}
int mlc_find_batch_bits(mlc_detector* d, const mlc_frame* frames, int64_t num_frames,
                        const uint8_t* bits, int bytes_per_desc, mlc_match* matches,
                        int64_t capacity, int64_t* match_offsets, int64_t* num_vertices,
                        int64_t* num_matches) {
  MLC_REQUIRE(d && num_vertices && num_matches && (num_frames == 0 || (frames && match_offsets)),
              "mlc_find_batch_bits: null argument");
  MLC_REQUIRE(bytes_per_desc > 0 && (bytes_per_desc % 16 == 0 || d->impl.exact_engine()),
              "bytes per descriptor must be a multiple of 16");
  std::string err;
  return d->impl.FindBatch(frames, num_frames, nullptr, bits, bytes_per_desc, matches, capacity,
                           match_offsets, num_vertices, num_matches, &err)
             ? 0
             : Fail(err);
}

int mlc_find_from_knn_device(mlc_detector* d, const mlc_frame* frames, int64_t num_frames,
                             const int32_t* d_idx, const float* d_dist, int k, mlc_match* matches,
                             int64_t capacity, int64_t* match_offsets, int64_t* num_vertices,
                             int64_t* num_matches) {
  MLC_REQUIRE(d && num_vertices && num_matches && (num_frames == 0 || (frames && match_offsets)),
              "mlc_find_from_knn_device: null argument");
  MLC_REQUIRE(k > 0 && k <= 16, "k must be in 1..16");
  std::string err;
  return d->impl.FindFromKnn(frames, num_frames, d_idx, d_dist, k, matches, capacity, match_offsets,
                             num_vertices, num_matches, &err)
             ? 0
             : Fail(err);
}

int mlc_pnp_ransac_batch(mlc_detector* d, const mlc_ransac_settings* rs, const mlc_camera* cams,
                         int num_cams, int64_t num_problems, const int64_t* offsets,
                         const double* keypoints, const int32_t* camera_index,
                         const int32_t* keypoint_index, const double* landmarks,
                         mlc_pose_result* results, uint8_t* inlier_flags) {
  MLC_REQUIRE(d && rs && cams && num_cams > 0 && num_problems >= 0, "mlc_pnp_ransac_batch: bad argument");
  MLC_REQUIRE(num_problems == 0 || (offsets && keypoints && camera_index && keypoint_index && landmarks && results),
              "mlc_pnp_ransac_batch: null argument");
  std::string err;
  return d->impl.PnpRansacBatch(*rs, cams, num_cams, num_problems, offsets, keypoints, camera_index,
                                keypoint_index, landmarks, results, inlier_flags, &err)
             ? 0
             : Fail(err);
}

int mlc_set_landmark_positions(mlc_detector* d, const double* xyz, int64_t n) {
  MLC_REQUIRE(d, "null detector");
  std::string err;
  return d->impl.SetLandmarkPositions(xyz, n, &err) ? 0 : Fail(err);
}
int mlc_set_landmark_positions_device(mlc_detector* d, const double* d_xyz, int64_t n) {
  MLC_REQUIRE(d, "null detector");
  std::string err;
  return d->impl.SetLandmarkPositionsDevice(d_xyz, n, &err) ? 0 : Fail(err);
}

int mlc_set_query_priors(mlc_detector* d, const double* T_G_I, int64_t num_vertices) {
  MLC_REQUIRE(d && (num_vertices == 0 || T_G_I) && num_vertices >= 0, "mlc_set_query_priors: bad argument");
  d->impl.SetQueryPriors(T_G_I, num_vertices);
  return 0;
}
void mlc_default_alignment_settings(mlc_alignment_settings* s) {
  s->num_iterations = 2000;
  s->rng_mapping = 1;
  s->max_orientation_error_rad = 0.174;
  s->max_position_error_m = 2.0;
  s->seed = 0u;
  s->pad_ = 0u;
}
int mlc_transformation_ransac(mlc_detector* d, const double* quats_xyzw, const double* positions, int64_t n,
                              const mlc_alignment_settings* settings, double* out_quat_xyzw, double* out_position,
                              int32_t* inlier_indices, int32_t* num_inliers) {
  MLC_REQUIRE(d && quats_xyzw && positions && settings && out_quat_xyzw && out_position && num_inliers,
              "mlc_transformation_ransac: null argument");
  std::string err;
  return d->impl.TransformationRansac(quats_xyzw, positions, n, *settings, out_quat_xyzw, out_position,
                                      inlier_indices, num_inliers, &err)
             ? 0
             : Fail(err);
}
int mlc_summary_map_parse(const void* blob, size_t size, mlc_summary_map_sizes* sizes,
                          float* G_landmark_position, float* G_observer_position, float* descriptors,
                          uint32_t* observer_indices, uint32_t* observation_to_landmark_index) {
  MLC_REQUIRE(sizes, "mlc_summary_map_parse: null sizes");
  mlc::SummaryMap m;
  std::string err;
  if (!m.Parse(blob, size, &err)) return Fail(err);
  sizes->num_landmarks = m.num_landmarks();
  sizes->num_observers = m.num_observers();
  sizes->num_observations = m.num_observations();
  sizes->descriptor_rows = m.descriptor_rows;
  sizes->descriptor_cols = m.descriptor_cols;
  auto copy = [](auto* dst, const auto& v) {
    if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(v[0]));
  };
  copy(G_landmark_position, m.G_landmark_position);
  copy(G_observer_position, m.G_observer_position);
  copy(descriptors, m.descriptors);
  copy(observer_indices, m.observer_indices);
  // observation_to_landmark_index is a separate repeated field: it has its own length
  if (observation_to_landmark_index &&
      m.observation_to_landmark_index.size() != m.observer_indices.size())
    return Fail("summary map: observer_indices and observation_to_landmark_index differ in length");
  copy(observation_to_landmark_index, m.observation_to_landmark_index);
  return 0;
}
int mlc_summary_map_serialize(const mlc_summary_map_sizes* sizes, const float* G_landmark_position,
                              const float* G_observer_position, const float* descriptors,
                              const uint32_t* observer_indices,
                              const uint32_t* observation_to_landmark_index, void* out,
                              size_t capacity, size_t* out_size) {
  MLC_REQUIRE(sizes && out_size, "mlc_summary_map_serialize: null argument");
  MLC_REQUIRE(sizes->num_landmarks >= 0 && sizes->num_observers >= 0 && sizes->num_observations >= 0 &&
                  sizes->descriptor_rows >= 0 && sizes->descriptor_cols >= 0 &&
                  sizes->descriptor_rows <= 0xFFFFFFFFLL && sizes->descriptor_cols <= 0xFFFFFFFFLL,
              "mlc_summary_map_serialize: bad sizes");
  const size_t nd = static_cast<size_t>(sizes->descriptor_rows) * static_cast<size_t>(sizes->descriptor_cols);
  MLC_REQUIRE((sizes->num_landmarks == 0 || G_landmark_position) &&
                  (sizes->num_observers == 0 || G_observer_position) && (nd == 0 || descriptors) &&
                  (sizes->num_observations == 0 || (observer_indices && observation_to_landmark_index)),
              "mlc_summary_map_serialize: null array");
  mlc::SummaryMap m;
  m.has_uncompressed_map = true;
  m.G_landmark_position.assign(G_landmark_position, G_landmark_position + 3 * sizes->num_landmarks);
  m.G_observer_position.assign(G_observer_position, G_observer_position + 3 * sizes->num_observers);
  m.descriptor_rows = static_cast<uint32_t>(sizes->descriptor_rows);
  m.descriptor_cols = static_cast<uint32_t>(sizes->descriptor_cols);
  m.descriptors.assign(descriptors, descriptors + nd);
  m.observer_indices.assign(observer_indices, observer_indices + sizes->num_observations);
  m.observation_to_landmark_index.assign(observation_to_landmark_index,
                                         observation_to_landmark_index + sizes->num_observations);
  std::vector<uint8_t> bytes;
  m.Serialize(&bytes);
  *out_size = bytes.size();
  if (!out || capacity < bytes.size()) return Fail("mlc_summary_map_serialize: output buffer too small");
  if (!bytes.empty()) std::memcpy(out, bytes.data(), bytes.size());
  return 0;
}
int mlc_create_summary_map(mlc_detector* d, int64_t num_landmarks, const double* G_landmark_position,
                           const int64_t* observations_per_landmark, int64_t num_observations,
                           const uint8_t* bits, int bytes_per_desc, const int64_t* observer_key,
                           const double* G_observer_position, void* out, size_t capacity, size_t* out_size) {
  MLC_REQUIRE(d && out_size && num_landmarks > 0 && G_landmark_position && observations_per_landmark,
              "mlc_create_summary_map: null argument or no landmarks (CHECK(!landmark_ids.empty()))");
  MLC_REQUIRE(num_observations > 0 && num_observations <= 0x7FFFFFFFLL && bits && observer_key && G_observer_position,
              "mlc_create_summary_map: no landmark observations for summary map");
  MLC_REQUIRE(bytes_per_desc > 0 && bytes_per_desc % 16 == 0,
              "mlc_create_summary_map: bytes per descriptor must be a positive multiple of 16");
  mlc::SummaryMap m;
  m.has_uncompressed_map = true;
  m.G_landmark_position.resize(static_cast<size_t>(num_landmarks) * 3);
  for (size_t i = 0; i < m.G_landmark_position.size(); ++i)
    m.G_landmark_position[i] = static_cast<float>(G_landmark_position[i]);
  // observation -> landmark index: landmark-major runs
  m.observation_to_landmark_index.reserve(static_cast<size_t>(num_observations));
  for (int64_t l = 0; l < num_landmarks; ++l) {
    MLC_REQUIRE(observations_per_landmark[l] >= 0, "mlc_create_summary_map: negative observation count");
    MLC_REQUIRE(static_cast<int64_t>(m.observation_to_landmark_index.size()) + observations_per_landmark[l] <=
                    num_observations,
                "mlc_create_summary_map: observation counts exceed num_observations");
    m.observation_to_landmark_index.insert(m.observation_to_landmark_index.end(),
                                           static_cast<size_t>(observations_per_landmark[l]), static_cast<uint32_t>(l));
  }
  MLC_REQUIRE(static_cast<int64_t>(m.observation_to_landmark_index.size()) == num_observations,
              "mlc_create_summary_map: observation counts do not add up to num_observations");
  // observers in order of first appearance (frame_id_to_index of the reference)
  std::unordered_map<int64_t, uint32_t> observer_of_key;
  m.observer_indices.resize(static_cast<size_t>(num_observations));
  for (int64_t i = 0; i < num_observations; ++i) {
    auto it = observer_of_key.find(observer_key[i]);
    if (it == observer_of_key.end()) {
      it = observer_of_key.emplace(observer_key[i], static_cast<uint32_t>(observer_of_key.size())).first;
      for (int c = 0; c < 3; ++c) m.G_observer_position.push_back(static_cast<float>(G_observer_position[3 * i + c]));
    }
    m.observer_indices[i] = it->second;
  }
  m.descriptor_rows = static_cast<uint32_t>(mlc_target_dim(d));
  m.descriptor_cols = static_cast<uint32_t>(num_observations);
  *out_size = m.SerializedSize();
  if (!out || capacity < *out_size) return Fail("mlc_create_summary_map: output buffer too small");
  m.descriptors.resize(static_cast<size_t>(num_observations) * m.descriptor_rows);
  std::string err;
  if (!d->impl.Project(bits, bytes_per_desc, num_observations, m.descriptors.data(), &err)) return Fail(err);
  std::vector<uint8_t> bytes;
  m.Serialize(&bytes);
  if (bytes.size() != *out_size) return Fail("mlc_create_summary_map: internal size mismatch");
  std::memcpy(out, bytes.data(), bytes.size());
  return 0;
}
int mlc_add_summary_map(mlc_detector* d, const void* blob, size_t size, int64_t mission_id,
                        int64_t first_vertex_id, int64_t first_landmark_id,
                        mlc_summary_map_sizes* sizes) {
  MLC_REQUIRE(d && (blob || size == 0), "mlc_add_summary_map: null argument");
  std::string err;
  int64_t s4[5] = {0, 0, 0, 0, 0};
  if (!d->impl.AddSummaryMap(blob, size, mission_id, first_vertex_id, first_landmark_id, s4, &err))
    return Fail(err);
  if (sizes) {
    sizes->num_landmarks = s4[0];
    sizes->num_observers = s4[1];
    sizes->num_observations = s4[2];
    sizes->descriptor_rows = s4[3];
    sizes->descriptor_cols = s4[4];
  }
  return 0;
}
namespace {
void Counts(const mlc::ViMapVertices& m, mlc_vi_map_counts* c) {
  c->num_vertices = m.num_vertices();
  c->num_frames = m.num_frames();
  c->num_keypoints = m.num_keypoints();
  c->num_landmarks = m.num_landmarks();
  c->descriptor_bytes = m.descriptor_bytes;
  c->pad_ = 0;
}
}  // namespace
int mlc_vi_map_count(const void* proto, size_t size, mlc_vi_map_counts* counts) {
  MLC_REQUIRE(counts, "mlc_vi_map_count: null counts");
  mlc::ViMapVertices m;
  std::string err;
  if (!m.Parse(proto, size, &err)) return Fail(err);
  Counts(m, counts);
  return 0;
}
int mlc_vi_map_read(const void* proto, size_t size, const mlc_vi_map_counts* counts, const mlc_vi_map_arrays* out) {
  MLC_REQUIRE(counts && out, "mlc_vi_map_read: null argument");
  mlc::ViMapVertices m;
  std::string err;
  if (!m.Parse(proto, size, &err)) return Fail(err);
  mlc_vi_map_counts now;
  Counts(m, &now);
  if (std::memcmp(&now, counts, sizeof(now)) != 0) return Fail("mlc_vi_map_read: counts do not belong to these bytes");
  auto copy = [](auto* dst, const auto& v) {
    if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(v[0]));
  };
  copy(out->vertex_id, m.vertex_id);
  copy(out->mission_id, m.mission_id);
  copy(out->T_M_I, m.T_M_I);
  copy(out->vertex_num_frames, m.vertex_num_frames);
  copy(out->vertex_num_landmarks, m.vertex_num_landmarks);
  copy(out->frame_timestamp_ns, m.frame_timestamp_ns);
  copy(out->frame_num_keypoints, m.frame_num_keypoints);
  copy(out->frame_is_valid, m.frame_is_valid);
  copy(out->keypoint_measurement, m.keypoint_measurement);
  copy(out->keypoint_descriptor, m.keypoint_descriptor);
  copy(out->keypoint_landmark_id, m.keypoint_landmark_id);
  copy(out->landmark_id, m.landmark_id);
  copy(out->landmark_p_B, m.landmark_p_B);
  copy(out->landmark_quality, m.landmark_quality);
  return 0;
}
int mlc_vi_map_missions(const void* proto, size_t size, int64_t capacity, uint64_t* mission_id, double* T_G_M,
                        int64_t* num_missions) {
  MLC_REQUIRE(num_missions && capacity >= 0, "mlc_vi_map_missions: bad argument");
  mlc::ViMapMissions m;
  std::string err;
  if (!m.Parse(proto, size, &err)) return Fail(err);
  *num_missions = m.num_missions();
  const int64_t n = std::min<int64_t>(capacity, m.num_missions());
  if (n > 0) {
    MLC_REQUIRE(mission_id && T_G_M, "mlc_vi_map_missions: null output");
    std::memcpy(mission_id, m.mission_id.data(), sizeof(uint64_t) * 2 * n);
    std::memcpy(T_G_M, m.T_G_M.data(), sizeof(double) * 7 * n);
  }
  return 0;
}
int mlc_alignment_enough_inliers(int32_t num_inliers, int64_t num_samples, int32_t min_inlier_count,
                                 double min_inlier_ratio) {
  const int by_ratio = static_cast<int>(static_cast<double>(num_samples) * min_inlier_ratio);
  const int threshold = min_inlier_count > by_ratio ? min_inlier_count : by_ratio;
  return num_inliers >= threshold ? 1 : 0;
}
int mlc_alignment_yaw_only(const double quat_xyzw[4], double out_quat_xyzw[4]) {
  MLC_REQUIRE(quat_xyzw && out_quat_xyzw, "mlc_alignment_yaw_only: null argument");
  // first column of Eigen's Quaternion::toRotationMatrix
  const double x = quat_xyzw[0], y = quat_xyzw[1], z = quat_xyzw[2], w = quat_xyzw[3];
  const double ty = 2.0 * y, tz = 2.0 * z;
  const double twy = ty * w, twz = tz * w, txy = ty * x, txz = tz * x, tyy = ty * y, tzz = tz * z;
  const double r00 = 1.0 - (tyy + tzz), r10 = txy + twz, r20 = txz - twy;
  const double pitch = std::atan2(-r20, std::sqrt(r00 * r00 + r10 * r10));
  const double cp = std::cos(pitch);
  const double yaw = std::fabs(cp) > 1.0e-12 ? std::atan2(r10 / cp, r00 / cp) : 0.0;
  // Rz(yaw) * Ry(0) * Rx(0) = [c -s 0; s c 0; 0 0 1], then Eigen's matrix -> quaternion branches
  const double c = std::cos(yaw), sn = std::sin(yaw);
  const double trace = c + c + 1.0;
  out_quat_xyzw[0] = 0.0;
  out_quat_xyzw[1] = 0.0;
  if (trace > 0.0) {
    const double t = std::sqrt(trace + 1.0);
    out_quat_xyzw[3] = 0.5 * t;
    out_quat_xyzw[2] = (sn - (-sn)) * (0.5 / t);
  } else {  // R22 = 1 is the largest diagonal entry
    const double t = std::sqrt(1.0 - c - c + 1.0);
    out_quat_xyzw[2] = 0.5 * t;
    out_quat_xyzw[3] = (sn - (-sn)) * (0.5 / t);
  }
  return 0;
}
int mlc_save_index(mlc_detector* d, const char* path) {
  MLC_REQUIRE(d && path, "mlc_save_index: null argument");
  std::string err;
  return d->impl.SaveIndex(path, &err) ? 0 : Fail(err);
}
int mlc_load_index(mlc_detector* d, const char* path) {
  MLC_REQUIRE(d && path, "mlc_load_index: null argument");
  std::string err;
  return d->impl.LoadIndex(path, &err) ? 0 : Fail(err);
}

static int QueryImpl(mlc_detector* d, const mlc_frame* frames, int64_t num_frames, const uint8_t* bits,
                     int bytes_per_desc, const double* keypoints, bool on_device,
                     const mlc_camera* cams, int num_cams, const mlc_ransac_settings* rs,
                     mlc_pose_result* results, int64_t* num_vertices, mlc_match* matches,
                     int64_t capacity, int64_t* match_offsets, int64_t* num_matches,
                     uint8_t* inlier_flags) {
  MLC_REQUIRE(d && rs && cams && num_cams > 0 && num_vertices, "mlc_query_batch: null argument");
  MLC_REQUIRE(num_frames == 0 || (frames && bits && keypoints && results), "mlc_query_batch: null argument");
  MLC_REQUIRE(bytes_per_desc > 0 && (bytes_per_desc % 16 == 0 || d->impl.exact_engine()),
              "bytes per descriptor must be a multiple of 16");
  std::string err;
  return d->impl.QueryBatch(frames, num_frames, bits, bytes_per_desc, keypoints, on_device, cams,
                            num_cams, *rs, results, num_vertices, matches, capacity, match_offsets,
                            num_matches, inlier_flags, &err)
             ? 0
             : Fail(err);
}
int mlc_query_batch(mlc_detector* d, const mlc_frame* frames, int64_t num_frames, const uint8_t* bits,
                    int bytes_per_desc, const double* keypoints, const mlc_camera* cams, int num_cams,
                    const mlc_ransac_settings* rs, mlc_pose_result* results, int64_t* num_vertices,
                    mlc_match* matches, int64_t capacity, int64_t* match_offsets,
                    int64_t* num_matches, uint8_t* inlier_flags) {
  return QueryImpl(d, frames, num_frames, bits, bytes_per_desc, keypoints, false, cams, num_cams, rs,
                   results, num_vertices, matches, capacity, match_offsets, num_matches, inlier_flags);
}
int mlc_query_batch_device(mlc_detector* d, const mlc_frame* frames, int64_t num_frames,
                           const uint8_t* d_bits, int bytes_per_desc, const double* d_keypoints,
                           const mlc_camera* cams, int num_cams, const mlc_ransac_settings* rs,
                           mlc_pose_result* results, int64_t* num_vertices, mlc_match* matches,
                           int64_t capacity, int64_t* match_offsets, int64_t* num_matches,
                           uint8_t* inlier_flags) {
  return QueryImpl(d, frames, num_frames, d_bits, bytes_per_desc, d_keypoints, true, cams, num_cams, rs,
                   results, num_vertices, matches, capacity, match_offsets, num_matches, inlier_flags);
}
int mlc_query_from_knn_device(mlc_detector* d, const mlc_frame* frames, int64_t num_frames,
                              const int32_t* d_idx, const float* d_dist, int k,
                              const double* d_keypoints, const mlc_camera* cams, int num_cams,
                              const mlc_ransac_settings* rs, mlc_pose_result* results,
                              int64_t* num_vertices, mlc_match* matches, int64_t capacity,
                              int64_t* match_offsets, int64_t* num_matches, uint8_t* inlier_flags) {
  MLC_REQUIRE(d && rs && cams && num_cams > 0 && num_vertices, "mlc_query_from_knn_device: null argument");
  MLC_REQUIRE(num_frames == 0 || (frames && d_idx && d_dist && d_keypoints && results),
              "mlc_query_from_knn_device: null argument");
  MLC_REQUIRE(k > 0 && k <= 16, "k must be in 1..16");
  std::string err;
  return d->impl.QueryFromKnn(frames, num_frames, d_idx, d_dist, k, d_keypoints, cams, num_cams, *rs,
                              results, num_vertices, matches, capacity, match_offsets, num_matches,
                              inlier_flags, &err)
             ? 0
             : Fail(err);
}

}  // extern "C"
