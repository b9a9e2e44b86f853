/* maplab_lc_b200.h — C-ABI of the B200-native loop-closure query path.
 *
 * Drop-in boundary for maplab's loop-closure query path. Every entry point
 * replaces one call of the reference (paths relative to the maplab checkout):
 *
 *   MBL = algorithms/loopclosure/matching-based-loopclosure
 *   LCH = algorithms/loopclosure/loop-closure-handler
 *   DP  = algorithms/loopclosure/descriptor-projection
 *
 * Conventions
 *  - plain pointers and sizes only; the library never keeps a caller pointer past the call;
 *  - every function returns 0 on success, non-zero on failure (mlc_last_error() has the text).
 *    The reference aborts (glog CHECK) where this library returns an error; the C++ shim
 *    (maplab_b200/csrc/loop_detector.h) turns non-zero into an exception / abort;
 *  - descriptors are one per row: `bits` is n x bytes_per_desc (== a column-major
 *    (bytes_per_desc x n) aslam DescriptorsT), projected descriptors are n x dim float
 *    (== column-major dim x n Eigen::MatrixXf);
 *  - 128-bit maplab ids (vertex / mission / landmark HashIds) cross the boundary as dense
 *    int64 numbers assigned by the shim;
 *  - there is NO CPU fallback: without a CUDA device / the sm_100a kernels every compute call fails.
 */
#ifndef MAPLAB_LC_B200_H_
#define MAPLAB_LC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mlc_detector mlc_detector; /* opaque */

/* Flags that change results (SURVEY.md §8b "Flags"); defaults = the reference's gflags. */
typedef struct mlc_settings {
  int32_t num_closest_words;       /* --lc_num_words_for_nn_search (10), MBL/src/detector-settings.cc:40 */
  int32_t num_nearest_neighbors;   /* --lc_num_neighbors (-1 = auto), MBL/src/detector-settings.cc:36 */
  int32_t scoring;                 /* --lc_scoring_function: 0 accumulation, 1 probabilistic */
  int32_t engine;                  /* --lc_detector_engine: 0 imi, 1 imipq, 2 hnsw (float descriptors; the search is
                                      EXACT on the device, see hnsw_* below) */
  double min_image_time_seconds;   /* --lc_min_image_time_seconds (10.0) */
  uint64_t min_verify_matches_num; /* --lc_min_verify_matches_num (10) */
  float fraction_best_scores;      /* --lc_fraction_best_scores (0.25) */
  float knn_epsilon;               /* --lc_knn_epsilon (2.0), loopclosure-common/src/flags.cc:9 */
  float knn_max_radius;            /* --lc_knn_max_radius (20.0), loopclosure-common/src/flags.cc:12 */
  int32_t device;                  /* CUDA device ordinal; -1 = current device */
  int32_t shard_rank;              /* this process' shard of the inverted lists (0 .. shard_count-1) */
  int32_t shard_count;             /* number of shards (GPUs); descriptor i lives on shard i % count */
  int32_t shard_mode;              /* 0: by descriptor index (i % count, the default: perfectly balanced, a shard
                                      is handed its own descriptors only). 1: by cell (hash(cell) % count — every
                                      shard keeps whole inverted lists; EXPERIMENTAL: every shard must be handed
                                      ALL descriptors and keeps the cells it owns) */
  /* engine 2 (loop_closure::HSNWIndexInterface, MBL/include/.../hnsw-index-interface.h; flags at
   * MBL/src/detector-settings.cc:36-47): descriptors are `float_descriptor_dim` floats (the reference reinterprets
   * the descriptor bytes, :155-163 — mlc_project does the same); no vocabulary file is read (vocab_blob may be NULL).
   * The device search is exhaustive, so hnsw_m / hnsw_ef_construction do not change results; hnsw_ef_query keeps the
   * reference's CHECK_LT(num_neighbors, ef_query). Neighbours come back as the reference interface returns them:
   * farthest first (popped from a max-heap, :141-151). */
  int32_t float_descriptor_dim;
  int32_t hnsw_m;                  /* --lc_hnsw_m (12) */
  int32_t hnsw_ef_construction;    /* --lc_hnsw_ef_construction (50) */
  int32_t hnsw_ef_query;           /* --lc_hnsw_ef_query (50) */
  int32_t pad_;
} mlc_settings;

/* One query (or database) frame header. */
typedef struct mlc_frame {
  int64_t timestamp_ns;  /* ProjectedImage::timestamp_nanoseconds, DP/.../descriptor-projection.h:23-31 */
  int64_t vertex_id;     /* dense number of ProjectedImage::keyframe_id.vertex_id */
  int64_t mission_id;    /* dense number of ProjectedImage::mission_id */
  int32_t frame_index;   /* ProjectedImage::keyframe_id.frame_index */
  int32_t num_descriptors;
} mlc_frame;

/* One structure match (vi_map::FrameKeyPointToStructureMatch, vi-map/loop-constraint.h:15-30). */
typedef struct mlc_match {
  int32_t query_frame;    /* index of the query frame in the batch */
  int32_t query_keypoint; /* keypoint_id_query.keypoint_index */
  int32_t db_descriptor;  /* global descriptor index in the database */
  int32_t db_keyframe;    /* insertion number of keyframe_id_result */
  int64_t db_vertex;      /* keyframe_id_result.vertex_id */
  int64_t landmark;       /* landmark_result */
} mlc_match;

/* Pinhole camera of the query n-camera rig (aslam::PinholeCamera + T_B_C). */
typedef struct mlc_camera {
  double fu, fv, cu, cv;
  int32_t distortion; /* aslam::Distortion of the camera: 0 none, 1 fisheye (FOV, dist[0] = w),
                         2 radtan (k1, k2, p1, p2), 3 equidistant (k1..k4) */
  int32_t pad_;
  double dist[4];
  double R_B_C[9]; /* row-major */
  double t_B_C[3];
} mlc_camera;

typedef struct mlc_ransac_settings {
  int32_t min_inlier_count;  /* --lc_min_inlier_count (10), LCH/src/loop-closure-handler.cc:10-16 */
  int32_t num_ransac_iters;  /* --lc_num_ransac_iters (100) */
  double min_inlier_ratio;   /* --lc_min_inlier_ratio (0.0) */
  double ransac_pixel_sigma; /* --lc_ransac_pixel_sigma (2.0) */
  uint32_t seed;             /* opengv seed when --lc_use_random_pnp_seed=false: 12345 */
  int32_t rng_mapping;       /* libstdc++ uniform_int_distribution mapping: 1 = GCC>=11, 0 = GCC<=10 */
  /* Topological gate, LCH/src/loop-closure-handler.cc:18-24, :424-455 (negative = off, the default):
   * a closure whose T_G_I differs from the query vertex' current pose (mlc_set_query_priors) by more
   * than this is rejected. */
  double max_delta_position_m;   /* --lc_max_delta_position_m (-1) */
  double max_delta_rotation_deg; /* --lc_max_delta_rotation_deg (-1) */
} mlc_ransac_settings;

/* Per-query result of geometric verification (LCH/src/loop-closure-handler.cc:235-550). */
typedef struct mlc_pose_result {
  int32_t accepted;      /* handleLoopClosure() return value */
  int32_t ransac_success;
  int32_t num_inliers;   /* best-per-keypoint inliers */
  int32_t num_ransac_inliers;
  int32_t iterations;
  int32_t model_indices[4];
  int32_t pad_;
  double inlier_ratio;
  double T_G_I[12]; /* 3x4 row-major [R|t] */
} mlc_pose_result;

const char* mlc_last_error(void);
int mlc_version(void);
/* Number of CUDA kernels this library has launched in this process (bench.py: gpu_launches). */
uint64_t mlc_kernel_launch_count(void);

void mlc_default_settings(mlc_settings* s);
void mlc_default_ransac_settings(mlc_ransac_settings* s);

/* LoopDetector::LoopDetector(settings) + vocabulary load
 * (MBL/src/matching-based-engine.cc:28-36, :287-317; InvertedMultiIndexVocabulary::Load,
 * MBL/include/matching-based-loopclosure/inverted-multi-index-interface.h:37-47, :72-88).
 * vocab_blob = bytes of the quantizer file named by --lc_projected_quantizer_filename. */
int mlc_create(const mlc_settings* settings, const void* vocab_blob, size_t vocab_size,
               mlc_detector** out);
void mlc_destroy(mlc_detector* d);

/* LoopDetector::Clear, MBL/src/matching-based-engine.cc:255-262 */
int mlc_clear(mlc_detector* d);
/* LoopDetector::NumEntries / NumDescriptors, MBL/include/.../matching-based-engine.h:44-50 */
int64_t mlc_num_entries(const mlc_detector* d);
int64_t mlc_num_descriptors(const mlc_detector* d);
/* getNumNeighborsToSearch, MBL/src/matching-based-engine.cc:319-338 */
int mlc_num_neighbors(const mlc_detector* d);
int mlc_target_dim(const mlc_detector* d);

/* LoopDetector::ProjectDescriptors -> descriptor_projection::ProjectDescriptorBlock
 * (MBL/src/matching-based-engine.cc:38-42, DP/src/descriptor-projection.cc:15-50).
 * bits: n x bytes_per_desc (host); out: n x dim (host). Kernel 1 (tcgen05 GEMM). */
int mlc_project(mlc_detector* d, const uint8_t* bits, int bytes_per_desc, int64_t n, float* out);
/* Same with device pointers on `stream` (cudaStream_t as void*); no host sync. */
int mlc_project_device(mlc_detector* d, const uint8_t* d_bits, int bytes_per_desc, int64_t n,
                       float* d_out, void* stream);

/* LoopDetector::Insert(ProjectedImage::Ptr), MBL/src/matching-based-engine.cc:217-253.
 * proj: num_descriptors x dim floats (host); landmarks: num_descriptors ids (may be NULL).
 * Like the reference's CHECK (:244-252) a keyframe (vertex id, frame index) that is already in the
 * database is refused (non-zero return, nothing inserted). */
int mlc_insert(mlc_detector* d, const mlc_frame* frame, const float* proj, const int64_t* landmarks);
/* Bulk variant: `num_frames` frames whose descriptors are concatenated in `proj`/`landmarks`. */
int mlc_insert_batch(mlc_detector* d, const mlc_frame* frames, int64_t num_frames, const float* proj,
                     const int64_t* landmarks);

/* Sharded database build (SURVEY.md section 8e "DB build shards trivially too"): descriptor i of the
 * database lives on shard i % shard_count, so a process only has to project — and hand over — the
 * descriptors its shard owns. Keyframe headers and landmark numbers cover ALL descriptors of the batch
 * (they are replicated metadata, like the reference's per-keyframe ProjectedImage copies,
 * MBL/src/matching-based-engine.cc:227-251); `proj_owned` holds one row per OWNED descriptor of the
 * batch in ascending global index (mlc_num_owned_in_range(d, mlc_num_descriptors(d), batch total) rows).
 * With shard_count == 1 these equal mlc_insert_batch. The _device variant takes device pointers
 * (d_landmarks may be NULL) — the database lives in HBM, nothing is staged on the host. */
int mlc_insert_batch_owned(mlc_detector* d, const mlc_frame* frames, int64_t num_frames,
                           const float* proj_owned, const int64_t* landmarks);
int mlc_insert_batch_device(mlc_detector* d, const mlc_frame* frames, int64_t num_frames,
                            const float* d_proj_owned, int64_t num_owned, const int64_t* d_landmarks,
                            void* stream);
/* Number of global descriptor indices in [first, first + count) owned by this shard. */
int64_t mlc_num_owned_in_range(const mlc_detector* d, int64_t first, int64_t count);

/* LoopDetector::Initialize (MBL/include/.../matching-based-engine.h:24) — here: freeze the
 * inserted keyframes and build the device index (cell assignment = FindClosestWords(desc, 1),
 * imilib/inverted-multi-index.h:77-94, then cell-sorted inverted lists). Called implicitly by
 * the first query after an insert. */
int mlc_initialize(mlc_detector* d);

/* IndexInterface::GetNNearestNeighborsForFeatures (MBL/include/.../index-interface.h:27-35;
 * imilib/inverted-multi-index.h:100-161): q n_q x dim, idx/dist n_q x k (host, caller-sized).
 * Missing neighbours are (-1, +inf), trailing. Kernels 2a + 2b. */
int mlc_knn(mlc_detector* d, const float* q, int64_t n_q, int k, int32_t* idx, float* dist);
int mlc_knn_device(mlc_detector* d, const float* d_q, int64_t n_q, int k, int32_t* d_idx,
                   float* d_dist, void* stream);
/* Parity checkpoints P2/P3: cell of each descriptor at insert (nw = 1) / visited cells per query
 * (nw = settings.num_closest_words; -1 for pairs with a missing word). cells: n x nw (host). */
int mlc_coarse_cells(mlc_detector* d, const float* q, int64_t n, int nw, int32_t* cells);
/* Kernel 2a and kernel 2b separately, device buffers (sharded path: visit lists are computed once
 * per query slice, exchanged, then every shard scans its own inverted lists for all queries).
 * d_cells: n x nw int32 (nw = settings.num_closest_words for the scan). */
int mlc_coarse_device(mlc_detector* d, const float* d_q, int64_t n, int nw, int32_t* d_cells, void* stream);
int mlc_scan_device(mlc_detector* d, const float* d_q, const int32_t* d_cells, int64_t n_q, int k,
                    int32_t* d_idx, float* d_dist, void* stream);
/* Merge `num_lists` per-shard top-k lists (each n_q x k, concatenated list-major, device) into
 * the k smallest by (distance, index) — the consumer of the NCCL all-gather. */
int mlc_merge_topk_device(mlc_detector* d, const int32_t* d_idx_lists, const float* d_dist_lists,
                          int num_lists, int64_t n_q, int k, int32_t* d_idx, float* d_dist,
                          void* stream);

/* scoring::computeAccumulationScore / computeProbabilisticScore
 * (MBL/include/matching-based-loopclosure/scoring.h:38-59, :92-187) over an explicit id list, on the
 * device: scores[i] for an id with num_matches[i] votes that owns num_descriptors[i] of the
 * num_db_descriptors database descriptors (scoring: 0 accumulation, 1 probabilistic; ids are scored
 * in the order given, which matters for the reference's in-loop +inf patch). An empty database or
 * id list yields no scores. mlc_find_batch applies the function selected in mlc_settings.scoring. */
int mlc_score(mlc_detector* d, int scoring, const uint64_t* num_matches, const uint64_t* num_descriptors,
              int n, int64_t num_db_descriptors, float* scores);
/* Algorithmic bytes the last mlc_knn* call scanned: sum over (query, visited cell present in
 * this shard) of list_length * bytes_per_entry (44 imi / 9 imipq) — SURVEY.md §8d. */
int mlc_last_scan_stats(mlc_detector* d, uint64_t* algorithmic_bytes, uint64_t* entries_scanned,
                        double* scan_kernel_ms);

/* Batched LoopDetector::Find (MBL/src/matching-based-engine.cc:48-168) for `num_frames` query
 * frames (frames of one vertex must be adjacent). proj: concatenated projected descriptors.
 * Output: matches in canonical order, match_offsets[num_vertices+1] (per query vertex, in order
 * of first appearance), written up to `capacity` entries; *num_matches = total found. */
int mlc_find_batch(mlc_detector* d, const mlc_frame* frames, int64_t num_frames, const float* proj,
                   mlc_match* matches, int64_t capacity, int64_t* match_offsets,
                   int64_t* num_vertices, int64_t* num_matches);
/* Same, starting from binary descriptors (project -> kNN -> vote/cluster on the device). */
int mlc_find_batch_bits(mlc_detector* d, const mlc_frame* frames, int64_t num_frames,
                        const uint8_t* bits, int bytes_per_desc, mlc_match* matches,
                        int64_t capacity, int64_t* match_offsets, int64_t* num_vertices,
                        int64_t* num_matches);

/* Kernel 3 only, on kNN lists that already live on the device (n_q x k for the concatenated
 * descriptors of `frames`; e.g. the merged result of the cross-shard all-gather). */
int mlc_find_from_knn_device(mlc_detector* d, const mlc_frame* frames, int64_t num_frames,
                             const int32_t* d_idx, const float* d_dist, int k, mlc_match* matches,
                             int64_t capacity, int64_t* match_offsets, int64_t* num_vertices,
                             int64_t* num_matches);

/* Batched geometric verification: LoopClosureHandler::handleLoopClosure gates +
 * PnpPoseEstimator::absoluteMultiPoseRansacPinholeCam (LCH/src/loop-closure-handler.cc:235-480,
 * aslam_cv2/aslam_cv_geometric_vision/src/pnp-pose-estimator.cc:75-132, :193-280).
 * Problem p covers correspondences [offsets[p], offsets[p+1]): keypoints 2 x n (u,v per match),
 * camera index, keypoint index, landmark position G_p (3 per match).
 * inlier_flags (may be NULL): per correspondence 0 = outlier, 1 = RANSAC inlier,
 * 3 = RANSAC inlier kept as best match of its keypoint. */
int mlc_pnp_ransac_batch(mlc_detector* d, const mlc_ransac_settings* rs, const mlc_camera* cams,
                         int num_cams, int64_t num_problems, const int64_t* offsets,
                         const double* keypoints, const int32_t* camera_index,
                         const int32_t* keypoint_index, const double* landmarks,
                         mlc_pose_result* results, uint8_t* inlier_flags);

/* CUDA-event times (ms) of the stages of the last mlc_query_batch* call:
 * {project, coarse word search, list scan, vote/cluster, gather + RANSAC}. */
int mlc_last_stage_ms(mlc_detector* d, double* ms5);

/* Landmark positions in the global frame by dense landmark id (what handleLoopClosure reads
 * through vi_map::VIMap::getLandmark_G_p, LCH/src/loop-closure-handler.cc:272-366). xyz: n x 3. */
int mlc_set_landmark_positions(mlc_detector* d, const double* xyz, int64_t n);
/* Same from a device array (copied). */
int mlc_set_landmark_positions_device(mlc_detector* d, const double* d_xyz, int64_t n);

/* Current poses T_G_I (3x4 row-major [R|t], map_->getVertex_T_G_I(query_vertex_id),
 * LCH/src/loop-closure-handler.cc:436-437) of the query vertices of the NEXT mlc_query_* /
 * mlc_pnp_ransac_batch call, one per query vertex / problem in batch order; copied, consumed by that
 * call. Only read when a max_delta_* limit is >= 0; with a limit set and no (or a wrong number of)
 * priors the call fails, like the reference's CHECK(query_vertex_id.isValid()). */
int mlc_set_query_priors(mlc_detector* d, const double* T_G_I, int64_t num_vertices);

/* Mission-level alignment after the queries.
 * common::transformationRansac (common/maplab-common/include/maplab-common/geometry-inl.h:113-182),
 * called by LoopDetectorNode::detectLoopClosuresMissionToDatabase
 * (LCH/src/loop-detector-node.cc:923-933) on the per-vertex T_G_M samples: every draw picks one
 * sample as hypothesis, inliers are the samples within both thresholds of it, the first hypothesis
 * whose inlier count beats the best so far (initially {0}) wins; the result is the least-squares
 * quaternion average (geometry.cc:9-31; the sign of the quaternion is not defined) and the mean
 * position of the winner's inliers. Quaternions are Eigen coeffs() order (x, y, z, w). */
typedef struct mlc_alignment_settings {
  int32_t num_iterations;           /* --anchor_transform_ransac_num_interations (2000) */
  int32_t rng_mapping;              /* libstdc++ uniform_int_distribution: 1 = GCC>=11, 0 = GCC<=10 */
  double max_orientation_error_rad; /* --anchor_transform_ransac_max_orientation_error_rad (0.174) */
  double max_position_error_m;      /* --anchor_transform_ransac_max_position_error_m (2.0) */
  uint32_t seed;                    /* the reference seeds from std::random_device */
  uint32_t pad_;
} mlc_alignment_settings;
void mlc_default_alignment_settings(mlc_alignment_settings* s);
/* inlier_indices: room for n entries (ascending), may be NULL. */
int mlc_transformation_ransac(mlc_detector* d, const double* quats_xyzw, const double* positions, int64_t n,
                              const mlc_alignment_settings* settings, double* out_quat_xyzw, double* out_position,
                              int32_t* inlier_indices, int32_t* num_inliers);

/* Tail of LoopDetectorNode::detectLoopClosuresMissionToDatabase after transformationRansac
 * (LCH/src/loop-detector-node.cc:936-955), host only:
 * mlc_alignment_enough_inliers: num_inliers >= max(--anchor_transform_min_inlier_count,
 *   int(num_samples * --anchor_transform_min_inlier_ratio)) -> 1, else 0;
 * mlc_alignment_yaw_only: "the datasets should be gravity-aligned so only yaw-axis rotation is necessary":
 *   RotationMatrixToRollPitchYaw (maplab-common/geometry-inl.h:41-61), roll = pitch = 0,
 *   RollPitchYawToRotationMatrix (:64-84), back to a quaternion as Eigen converts a matrix (x, y, z, w). */
int mlc_alignment_enough_inliers(int32_t num_inliers, int64_t num_samples, int32_t min_inlier_count,
                                 double min_inlier_ratio);
int mlc_alignment_yaw_only(const double quat_xyzw[4], double out_quat_xyzw[4]);

/* Database persistence (no reference counterpart: maplab rebuilds the loop-closure database for
 * every `lc` / `aam` / `relax` invocation and per mission, LCH/src/loop-detector-node.cc:273-339,
 * vi-map-merger.cc:71-78). mlc_save_index writes the built index — keyframe headers, projected
 * descriptors, landmark numbers, inverted lists, cell table, landmark positions — to one file;
 * mlc_load_index replaces the detector's database with it (the detector must have been created with
 * the same vocabulary, engine and sharding: checked, non-zero otherwise). Queries on a loaded
 * index return exactly what they returned on the index that was saved. */
int mlc_save_index(mlc_detector* d, const char* path);
int mlc_load_index(mlc_detector* d, const char* path);

/* Fused batched query = LoopDetectorNode::queryVertexInDatabase for a batch of vertices
 * (LCH/src/loop-detector-node.cc:668-766, :819-873): project -> kNN -> Find -> correspondence
 * assembly -> handleLoopClosure verdict + T_G_I, all on the device. keypoints: 2 doubles per
 * query descriptor (same order as bits); frame_index of each frame = its camera in `cams`.
 * results: one per query vertex (in order of first appearance). matches / match_offsets /
 * num_matches / inlier_flags may be NULL. */
int mlc_query_batch(mlc_detector* d, const mlc_frame* frames, int64_t num_frames, const uint8_t* bits,
                    int bytes_per_desc, const double* keypoints, const mlc_camera* cams, int num_cams,
                    const mlc_ransac_settings* rs, mlc_pose_result* results, int64_t* num_vertices,
                    mlc_match* matches, int64_t capacity, int64_t* match_offsets,
                    int64_t* num_matches, uint8_t* inlier_flags);
/* Same with `d_bits` / `d_keypoints` already resident on the device. */
int mlc_query_batch_device(mlc_detector* d, const mlc_frame* frames, int64_t num_frames,
                           const uint8_t* d_bits, int bytes_per_desc, const double* d_keypoints,
                           const mlc_camera* cams, int num_cams, const mlc_ransac_settings* rs,
                           mlc_pose_result* results, int64_t* num_vertices, mlc_match* matches,
                           int64_t capacity, int64_t* match_offsets, int64_t* num_matches,
                           uint8_t* inlier_flags);
/* Kernels 3 + 4 on (merged) kNN lists that live on the device: the per-rank tail of the sharded
 * path after the NCCL all-gather. d_keypoints: 2 per query descriptor. */
int mlc_query_from_knn_device(mlc_detector* d, const mlc_frame* frames, int64_t num_frames,
                              const int32_t* d_idx, const float* d_dist, int k,
                              const double* d_keypoints, const mlc_camera* cams, int num_cams,
                              const mlc_ransac_settings* rs, mlc_pose_result* results,
                              int64_t* num_vertices, mlc_match* matches, int64_t capacity,
                              int64_t* match_offsets, int64_t* num_matches, uint8_t* inlier_flags);

/* ---- Multi-GPU: the database sharded over the GPUs of one box ----------------
 * One process per GPU, each with a detector created with shard_rank = its rank and shard_count = the
 * number of ranks; the inverted lists are sharded (descriptor i on shard i % shard_count), vocabulary
 * and metadata replicated. This replaces the reference's fan-out of query vertices over host threads
 * (LoopDetectorNode::detectLoopClosuresVerticesToDatabase, LCH/src/loop-detector-node.cc:819-873;
 * common::ParallelProcess, common/maplab-common/include/maplab-common/parallel-process.h:49-92): every
 * rank brings ITS slice of the step's query vertices and gets the verdicts of that slice.
 *
 * Communicator: rank 0 calls mlc_comm_unique_id and hands the 128 bytes to the other ranks by any means
 * (MPI_Bcast, a file, torch.distributed); then every rank calls mlc_comm_init (collective, =
 * ncclCommInitRank over shard_count ranks). NCCL is bound at run time (libnccl.so.2). */
#define MLC_COMM_ID_BYTES 128
int mlc_comm_unique_id(void* id128);
int mlc_comm_init(mlc_detector* d, const void* id128);
int mlc_comm_destroy(mlc_detector* d);
int mlc_comm_nccl_version(const mlc_detector* d); /* ncclGetVersion of the NCCL in use, 0 before init */
/* queryVertexInDatabase for this rank's slice of a batch against the sharded database — COLLECTIVE:
 * every rank calls it once per step (a rank without queries passes num_frames = 0). Arguments and
 * results as mlc_query_batch / mlc_query_batch_device, for the slice. Per step each rank projects and
 * coarse-searches its slice, one grouped NCCL all-gather carries (projected query, visit list) to all
 * shards, every shard scans its lists block by block — its own slice first, while the all-gather is in
 * flight — and each block's per-shard top-k lists leave for their owner (one grouped send/recv per
 * block) while the next block is scanned; the owner merges the shard lists by (distance, index), which
 * equals the single-index result, and runs voting / clustering / RANSAC on its slice. */
int mlc_sharded_query_batch(mlc_detector* d, const mlc_frame* frames, int64_t num_frames, const uint8_t* bits,
                            int bytes_per_desc, const double* keypoints, const mlc_camera* cams, int num_cams,
                            const mlc_ransac_settings* rs, mlc_pose_result* results, int64_t* num_vertices,
                            mlc_match* matches, int64_t capacity, int64_t* match_offsets, int64_t* num_matches,
                            uint8_t* inlier_flags);
int mlc_sharded_query_batch_device(mlc_detector* d, const mlc_frame* frames, int64_t num_frames,
                                   const uint8_t* d_bits, int bytes_per_desc, const double* d_keypoints,
                                   const mlc_camera* cams, int num_cams, const mlc_ransac_settings* rs,
                                   mlc_pose_result* results, int64_t* num_vertices, mlc_match* matches,
                                   int64_t capacity, int64_t* match_offsets, int64_t* num_matches,
                                   uint8_t* inlier_flags);
/* IndexInterface::GetNNearestNeighborsForFeatures of this rank's n_q query descriptors (device,
 * n_q x dim) against the whole sharded database — COLLECTIVE like the call above. */
int mlc_sharded_knn_device(mlc_detector* d, const float* d_q, int64_t n_q, int k, int32_t* d_idx, float* d_dist);

/* Localization summary maps.
 * File format: proto2 summary_map.proto.LocalizationSummaryMap (map-structure/localization-summary-map/
 * proto/localization-summary-map/localization-summary-map.proto:4-14, common.proto.MatrixXf of
 * common/maplab-common/proto/maplab-common/eigen.proto:8-12), the message inside the file
 * "localization_summary_map" written by LocalizationSummaryMap::saveToFolder (src/localization-summary-map.cc:33-52,
 * :121-140). `blob` is the serialized message: with --proto_use_compression (the default) the file on disk is that
 * message as a gzip stream (maplab-common/src/proto-serialization-helper.cc:63-76, :120-131), to be inflated by the
 * caller. The wire format is decoded here (no libprotobuf); matrices are column-major as in eigen_proto. */
typedef struct mlc_summary_map_sizes {
  int64_t num_landmarks;    /* G_landmark_position.cols() */
  int64_t num_observers;    /* G_observer_position.cols() */
  int64_t num_observations; /* observer_indices.rows() */
  int64_t descriptor_rows;  /* projected_descriptors.rows() */
  int64_t descriptor_cols;  /* projected_descriptors.cols() */
} mlc_summary_map_sizes;
/* LocalizationSummaryMap::deserialize (src/localization-summary-map.cc:53-93), host only (no device
 * needed). Every output array may be NULL; call once with NULL arrays for the sizes. Arrays:
 * G_landmark_position 3 x L, G_observer_position 3 x O, descriptors rows x cols (all column-major
 * float), observer_indices / observation_to_landmark_index one uint32 per observation. */
int mlc_summary_map_parse(const void* blob, size_t size, mlc_summary_map_sizes* sizes,
                          float* G_landmark_position, float* G_observer_position, float* descriptors,
                          uint32_t* observer_indices, uint32_t* observation_to_landmark_index);
/* LocalizationSummaryMap::serialize (src/localization-summary-map.cc:33-52), host only: writes the
 * bytes libprotobuf writes for these arrays. Returns the size needed through *out_size; fills `out`
 * when capacity suffices (non-zero return otherwise, *out_size still set). */
int mlc_summary_map_serialize(const mlc_summary_map_sizes* sizes, const float* G_landmark_position,
                              const float* G_observer_position, const float* descriptors,
                              const uint32_t* observer_indices,
                              const uint32_t* observation_to_landmark_index, void* out,
                              size_t capacity, size_t* out_size);
/* summary_map::createLocalizationSummaryMapFromLandmarkList (map-structure/localization-summary-map/src/
 * localization-summary-map-creation.cc:63-204) + LocalizationSummaryMap::serialize: the observations of the
 * chosen landmarks, landmark-major (all observations of landmark 0, then of landmark 1, ...), are projected on
 * the device (kernel 1 = descriptor_projection::ProjectDescriptor with the detector's projection matrix and
 * target dimensionality); observers are numbered in order of first appearance of `observer_key` (any number
 * that identifies the observing visual frame, the reference keys on vi_map::VisualFrameIdentifier) and take the
 * G_p_I of that first observation; positions are cast to float like setGLandmarkPosition / setGObserverPosition.
 *   G_landmark_position: 3 doubles per landmark; observations_per_landmark: one count per landmark (sum =
 *   num_observations); bits: num_observations x bytes_per_desc; observer_key / G_observer_position
 *   (3 doubles): per observation.
 * Writes the serialized proto into `out` (size needed through *out_size; non-zero return and nothing
 * written when capacity is too small — the required size depends only on the counts and ids, so a first call
 * with out = NULL costs no device work). */
int mlc_create_summary_map(mlc_detector* d, int64_t num_landmarks, const double* G_landmark_position,
                           const int64_t* observations_per_landmark, int64_t num_observations,
                           const uint8_t* bits, int bytes_per_desc, const int64_t* observer_key,
                           const double* G_observer_position, void* out, size_t capacity, size_t* out_size);
/* LoopDetectorNode::addLocalizationSummaryMapToDatabase (LCH/src/loop-detector-node.cc:341-432):
 * one database image per observer (timestamp 0, frame index 0, mission `mission_id` — the reference
 * draws a random mission id per summary map; the caller hands out a dense one that no other mission
 * uses), vertex id of observer o = first_vertex_id + o, landmark id of landmark l =
 * first_landmark_id + l; the landmark positions (float, cast to double like
 * LocalizationSummaryMap::getGLandmarkPosition) are written into the landmark table at those ids;
 * ends with LoopDetector::Initialize. `sizes` may be NULL. */
int mlc_add_summary_map(mlc_detector* d, const void* blob, size_t size, int64_t mission_id,
                        int64_t first_vertex_id, int64_t first_landmark_id,
                        mlc_summary_map_sizes* sizes);

/* Saved vi_maps: one `vertices<N>` file = the message vi_map.proto.VIMap with its
 * vertex_ids / vertices (map-structure/vi-map/proto/vi-map/vi_map.proto:27-43, :82-100, :161-176;
 * aslam-serialization/visual-frame.proto:5-26), written by vi_map::serialization::serializeVertices
 * (vi-map/src/vi-map-serialization.cc:27-43). `proto` is the serialized message — the file on disk is a gzip stream
 * of it (proto-serialization-helper.cc:120-131), inflate it first. Host only; what the loop-closure path consumes
 * (LoopDetectorNode::addVertexToDatabase / queryVertexInDatabase, LCH/src/loop-detector-node.cc:273-339, :668-766). */
typedef struct mlc_vi_map_counts {
  int64_t num_vertices, num_frames, num_keypoints, num_landmarks;
  int32_t descriptor_bytes; /* 48 BRISK, 64 FREAK; 0 without keypoints */
  int32_t pad_;
} mlc_vi_map_counts;
typedef struct mlc_vi_map_arrays { /* caller-allocated from mlc_vi_map_counts; any pointer may be NULL */
  uint64_t* vertex_id;             /* 2 words per vertex (aslam::HashId) */
  uint64_t* mission_id;            /* 2 per vertex */
  double* T_M_I;                   /* 7 per vertex: quaternion x y z w, position (eigen-proto-inl.h:165-177) */
  int32_t* vertex_num_frames;      /* per vertex */
  int32_t* vertex_num_landmarks;   /* per vertex: size of its landmark store */
  int64_t* frame_timestamp_ns;     /* per visual frame, vertex-major */
  int32_t* frame_num_keypoints;
  uint8_t* frame_is_valid;         /* 1 = the frame is set (valid frame id) and not invalidated: what
                                      addVertexToDatabase / queryVertexInDatabase require of a frame */
  double* keypoint_measurement;    /* 2 per keypoint, frame-major */
  uint8_t* keypoint_descriptor;    /* descriptor_bytes per keypoint */
  uint64_t* keypoint_landmark_id;  /* 2 per keypoint; (0, 0) = no landmark */
  uint64_t* landmark_id;           /* 2 per stored landmark, vertex-major */
  double* landmark_p_B;            /* 3 per landmark: position in the storing vertex' frame */
  int32_t* landmark_quality;       /* vi_map::Landmark::Quality: 0 unknown, 1 bad, 2 good */
} mlc_vi_map_arrays;
int mlc_vi_map_count(const void* proto, size_t size, mlc_vi_map_counts* counts);
/* `counts` must be what mlc_vi_map_count returned for the same bytes (checked). */
int mlc_vi_map_read(const void* proto, size_t size, const mlc_vi_map_counts* counts, const mlc_vi_map_arrays* out);
/* The `missions` file of the folder (vi_map.proto:102-131: missions and their base frames): up to `capacity`
 * missions, mission_id 2 words and T_G_M 7 doubles (quaternion x y z w, position) each; *num_missions is always
 * the number in the file (call with capacity 0 for the count). */
int mlc_vi_map_missions(const void* proto, size_t size, int64_t capacity, uint64_t* mission_id, double* T_G_M,
                        int64_t* num_missions);

#ifdef __cplusplus
}
#endif
#endif /* MAPLAB_LC_B200_H_ */
