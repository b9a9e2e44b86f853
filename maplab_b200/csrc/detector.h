// Host-side engine of the B200 loop-closure path: the state behind the C-ABI handle. Mirrors
// matching_based_loopclosure::LoopDetector (matching-based-loopclosure/src/matching-based-engine.cc)
// with the database resident in HBM.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <mutex>
#include <string>
#include <unordered_set>
#include <vector>

#include "../../include/maplab_lc_b200.h"
#include "device_index.h"
#include "vocabulary.h"

namespace mlc {

// Grow-only device buffer.
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t Reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void Free() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T>
  T* as() const {
    return static_cast<T*>(p);
  }
};

// Growing device array that keeps its contents (the database lives in HBM; the host only stages).
struct GrowBuf {
  void* p = nullptr;
  size_t cap = 0;   // bytes allocated
  size_t used = 0;  // bytes holding data
  // Room for `extra` more bytes behind `used`; contents are preserved.
  cudaError_t Extend(size_t extra, cudaStream_t stream) {
    const size_t need = used + extra;
    if (need <= cap) return cudaSuccess;
    size_t want = need + need / 2 + 256;
    void* np = nullptr;
    cudaError_t e = cudaMalloc(&np, want);
    if (e != cudaSuccess) {  // no room for the slack: try the exact size
      want = need + 256;
      e = cudaMalloc(&np, want);
      if (e != cudaSuccess) return e;
    }
    if (used > 0) {
      e = cudaMemcpyAsync(np, p, used, cudaMemcpyDeviceToDevice, stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
      if (e != cudaSuccess) {
        cudaFree(np);
        return e;
      }
    }
    if (p) cudaFree(p);
    p = np;
    cap = want;
    return cudaSuccess;
  }
  unsigned char* end() const { return static_cast<unsigned char*>(p) + used; }
  void Free() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = used = 0;
  }
  template <typename T>
  T* as() const {
    return static_cast<T*>(p);
  }
};

struct KeyframeMeta {
  int64_t ts, vertex, mission;
  int32_t frame_index, first_descriptor, num_descriptors;
};

// ncclGetUniqueId through the run-time bound NCCL (sharded.cu); id128: 128 bytes.
bool CommUniqueId(void* id128, std::string* err);

class Detector {
 public:
  Detector() = default;
  ~Detector();
  bool Create(const mlc_settings& s, const void* blob, size_t size, std::string* err);

  bool Clear(std::string* err);
  int64_t NumEntries() const { return static_cast<int64_t>(keyframes_.size()); }
  int64_t NumDescriptors() const { return num_desc_; }
  int64_t NumOwnedDescriptors() const { return num_own_ + static_cast<int64_t>(pend_gidx_.size()); }
  int NumNeighbors() const;
  int dim() const { return s_.engine == 2 ? s_.float_descriptor_dim : vocab_.target_dim; }
  bool exact_engine() const { return s_.engine == 2; }

  bool Project(const uint8_t* bits, int bytes_per_desc, int64_t n, float* out, std::string* err);
  bool ProjectDevice(const uint8_t* d_bits, int bytes_per_desc, int64_t n, float* d_out,
                     cudaStream_t stream, std::string* err);
  // proj_is_owned_rows: `proj` holds only the rows of the descriptors this shard owns (ascending
  // global index) instead of one row per descriptor.
  bool InsertBatch(const mlc_frame* frames, int64_t num_frames, const float* proj,
                   const int64_t* landmarks, bool proj_is_owned_rows, std::string* err);
  // Same with device pointers: d_proj_owned = the owned rows, d_landmarks = all rows (or null).
  bool InsertBatchDevice(const mlc_frame* frames, int64_t num_frames, const float* d_proj_owned,
                         int64_t num_owned, const int64_t* d_landmarks, cudaStream_t stream,
                         std::string* err);
  // Global descriptor indices this shard owns among [first, first + count) (descriptor i lives on
  // shard i % shard_count): count and, if `out` is given, the indices in ascending order.
  int64_t OwnedInRange(int64_t first, int64_t count) const;
  bool Initialize(std::string* err);
  bool Knn(const float* q, int64_t n_q, int k, int32_t* idx, float* dist, std::string* err);
  bool KnnDevice(const float* d_q, int64_t n_q, int k, int32_t* d_idx, float* d_dist,
                 cudaStream_t stream, std::string* err);
  // Kernel 2a / 2b separately on device buffers (sharded path: cells are computed once per query
  // slice, all-gathered, then every shard scans its lists for all queries).
  bool CoarseDevice(const float* d_q, int64_t n, int nw, int32_t* d_cells, cudaStream_t stream,
                    std::string* err);
  bool ScanDevice(const float* d_q, const int32_t* d_cells, int64_t n_q, int k, int32_t* d_idx,
                  float* d_dist, cudaStream_t stream, std::string* err);
  // CUDA-event stage times of the last fused query: project, coarse, scan, vote_cluster, ransac.
  bool LastStageMs(double* ms5, std::string* err);
  bool CoarseCells(const float* q, int64_t n, int nw, int32_t* cells, std::string* err);
  bool MergeTopkDevice(const int32_t* d_idx_lists, const float* d_dist_lists, int num_lists,
                       int64_t n_q, int k, int32_t* d_idx, float* d_dist, cudaStream_t stream,
                       std::string* err);
  bool LastScanStats(uint64_t* bytes, uint64_t* entries, double* ms, std::string* err);
  // scoring::compute*Score over an explicit id list, on the device (scoring.h:38-59, :92-187).
  bool Score(int scoring, const uint64_t* votes, const uint64_t* num_desc, int n, int64_t num_db,
             float* scores, std::string* err);
  bool FindBatch(const mlc_frame* frames, int64_t num_frames, const float* proj,
                 const uint8_t* bits, int bytes_per_desc, mlc_match* matches, int64_t capacity,
                 int64_t* match_offsets, int64_t* num_vertices, int64_t* num_matches,
                 std::string* err);
  // Kernel 3 on kNN results that already live on the device (after the cross-shard merge).
  bool FindFromKnn(const mlc_frame* frames, int64_t num_frames, const int32_t* d_idx,
                   const float* d_dist, int k, mlc_match* matches, int64_t capacity,
                   int64_t* match_offsets, int64_t* num_vertices, int64_t* num_matches,
                   std::string* err);
  bool FindOnDevice(const mlc_frame* frames, int64_t num_frames, const int32_t* d_idx,
                    const float* d_dist, int k, std::vector<long long>* fin_off, std::string* err);
  bool RansacOnDevice(const mlc_ransac_settings& rs, const mlc_camera* cams, int num_cams,
                      int64_t num_problems, int64_t total, const int64_t* d_offsets,
                      const double* d_keypoints, const int32_t* d_camera_index,
                      const int32_t* d_keypoint_index, const double* d_landmarks,
                      mlc_pose_result* results, uint8_t* inlier_flags, std::string* err);
  // Landmark positions in the global frame by dense landmark id (vi_map::Landmark::get_p_G of
  // the landmark store, read by loop-closure-handler.cc:272-366); replicated on every shard.
  bool SetLandmarkPositions(const double* xyz, int64_t n, std::string* err);
  bool SetLandmarkPositionsDevice(const double* d_xyz, int64_t n, std::string* err);
  // LoopDetectorNode::addLocalizationSummaryMapToDatabase (LCH/src/loop-detector-node.cc:341-432):
  // one database image per observer of a serialized LocalizationSummaryMap (timestamp 0, one
  // mission id for the whole map, frame index 0, pre-projected descriptors), then Initialize().
  // Observer o gets vertex id first_vertex_id + o, landmark l gets id first_landmark_id + l and its
  // position (LocalizationSummaryMap::getGLandmarkPosition: float -> double) is written into the
  // landmark table at that id. sizes5 = {landmarks, observers, observations, descriptor rows, descriptor cols}.
  bool AddSummaryMap(const void* blob, size_t size, int64_t mission_id, int64_t first_vertex_id,
                     int64_t first_landmark_id, int64_t* sizes5, std::string* err);
  // Database persistence: the built index (inverted lists, cell table, metadata,
  // landmark positions) as one file, so that later runs skip projection + cell assignment + sort.
  void SetQueryPriors(const double* T_G_I, int64_t n) {
    std::lock_guard<std::recursive_mutex> lock(mu_);
    priors_.assign(T_G_I, T_G_I + 12 * n);
    have_priors_ = true;
  }
  bool TransformationRansac(const double* quats, const double* positions, int64_t n,
                            const mlc_alignment_settings& as, double* out_quat, double* out_pos,
                            int32_t* inlier_indices, int32_t* num_inliers, std::string* err);
  bool SaveIndex(const char* path, std::string* err);
  bool LoadIndex(const char* path, std::string* err);
  // Fused query: project -> kNN -> kernel 3 -> correspondence gather -> kernel 4.
  bool QueryBatch(const mlc_frame* frames, int64_t num_frames, const uint8_t* bits, int bytes_per_desc,
                  const double* keypoints, bool inputs_on_device, const mlc_camera* cams, int num_cams,
                  const mlc_ransac_settings& rs, mlc_pose_result* results, int64_t* num_vertices,
                  mlc_match* matches, int64_t capacity, int64_t* match_offsets, int64_t* num_matches,
                  uint8_t* inlier_flags, std::string* err);
  bool QueryFromKnn(const mlc_frame* frames, int64_t num_frames, const int32_t* d_idx,
                    const float* d_dist, int k, const double* d_keypoints, const mlc_camera* cams,
                    int num_cams, const mlc_ransac_settings& rs, mlc_pose_result* results,
                    int64_t* num_vertices, mlc_match* matches, int64_t capacity,
                    int64_t* match_offsets, int64_t* num_matches, uint8_t* inlier_flags,
                    std::string* err);
  // ---- multi-GPU (sharded.cu): NCCL communicator over the shard_count ranks ----
  bool CommInit(const void* id128, std::string* err);
  void CommDestroy();
  int CommVersion() const;
  bool ShardedKnnDevice(const float* d_q, int64_t n_s, int k, int32_t* d_idx, float* d_dist, std::string* err);
  bool ShardedQueryBatch(const mlc_frame* frames, int64_t num_frames, const uint8_t* bits, int bytes_per_desc,
                         const double* keypoints, bool inputs_on_device, const mlc_camera* cams, int num_cams,
                         const mlc_ransac_settings& rs, mlc_pose_result* results, int64_t* num_vertices,
                         mlc_match* matches, int64_t capacity, int64_t* match_offsets, int64_t* num_matches,
                         uint8_t* inlier_flags, std::string* err);
  bool PnpRansacBatch(const mlc_ransac_settings& rs, const mlc_camera* cams, int num_cams,
                      int64_t num_problems, const int64_t* offsets, const double* keypoints,
                      const int32_t* camera_index, const int32_t* keypoint_index,
                      const double* landmarks, mlc_pose_result* results, uint8_t* inlier_flags,
                      std::string* err);

 private:
  bool Cuda(cudaError_t e, const char* what, std::string* err) const;
  bool Nccl(int result, const char* what, std::string* err) const;
  bool ShardedSliceSizes(int64_t n_s, int64_t* n_max, std::string* err);
  bool ShardedKnnOnSlice(int64_t n_s, int64_t n_max, int k, std::string* err);
  void* comm_ = nullptr;  // ncclComm_t
  cudaStream_t comm_stream_ = nullptr;
  cudaEvent_t ev_comm_[4] = {nullptr, nullptr, nullptr, nullptr};
  static constexpr int kMaxShards = 16;
  cudaEvent_t ev_scan_[2 * kMaxShards] = {};
  DevBuf sh_counts_, sh_q_all_, sh_cells_all_, sh_pidx_, sh_pdist_, sh_ridx_, sh_rdist_;
  const int32_t* last_cells_ = nullptr;  // visit list of the last scan (for mlc_last_scan_stats)
  int last_scan_launches_ = 0;           // > 0: the scan ran as that many launches timed by ev_scan_
  bool EnsureIndex(std::string* err);
  bool UploadTrees(std::string* err);
  bool UploadKeyframeReplicas(std::string* err);

  mlc_settings s_{};
  VocabularyFile vocab_;
  uint64_t vocab_hash_ = 0;  // FNV-1a of the vocabulary blob (index files are tied to it)
  FixedProjection fp_;
  KdTreeHost tree1_, tree2_;
  int device_ = 0, sm_count_ = 148;
  cudaStream_t stream_ = nullptr;
  // host-buffer queries: the H2D copies run on their own stream in chunks, so that projection and
  // coarse search of chunk i overlap the copy of chunk i+1
  cudaStream_t copy_stream_ = nullptr;
  cudaStream_t ransac_stream_[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // extra streams of the RANSAC problem groups
  cudaEvent_t ev_ransac_[13] = {};  // [0, 5): group g + 1 joined; [6]: fork; [7, 13): last enqueued round of group g done
  int* h_remaining_ = nullptr;  // pinned round counters
  std::vector<double> priors_;  // T_G_I of the query vertices of the next call (delta-pose gate)
  bool have_priors_ = false;
  bool corr_grouped_ = false;  // next RansacOnDevice call: correspondences grouped by (camera, keypoint)
  static constexpr int kCopyChunks = 8;  // upper bound; 2 are used (tuned on B200, PCIe gen5)
  cudaEvent_t ev_copy_[kCopyChunks + 1] = {};
  cudaEvent_t ev0_ = nullptr, ev1_ = nullptr;
  cudaEvent_t ev_stage_[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool stage_valid_ = false;
  std::recursive_mutex mu_;  // queries are serialised on the detector's stream (SURVEY §8b Threading)

  ProjectionDevice proj_;
  void* d_tree_blob_ = nullptr;
  CoarseParams coarse_{};
  PqParams pq_{};            // engine 1 (imipq)
  void* d_pq_blob_ = nullptr;
  bool UploadPq(std::string* err);
  // engine dispatch of the inverted-list scan (kernel 2b)
  cudaError_t LaunchScan(const float* d_q, int64_t n_q, const int32_t* d_cells, int nw, int k,
                         int32_t* d_idx, float* d_dist, cudaStream_t stream);

  // database. What Insert keeps (matching-based-engine.cc:227-251) lives in HBM; the host holds the
  // keyframe headers, the uniqueness set of Insert's CHECK (:244-252) and a bounded staging area for
  // host-side inserts that is flushed to the device in blocks.
  std::vector<KeyframeMeta> keyframes_;
  struct KeyframeKey {
    int64_t vertex;
    int32_t frame_index;
    bool operator==(const KeyframeKey& o) const { return vertex == o.vertex && frame_index == o.frame_index; }
  };
  struct KeyframeKeyHash {
    size_t operator()(const KeyframeKey& k) const {
      uint64_t h = static_cast<uint64_t>(k.vertex) * 0x9E3779B97F4A7C15ull + static_cast<uint32_t>(k.frame_index);
      return static_cast<size_t>(h ^ (h >> 29));
    }
  };
  std::unordered_set<KeyframeKey, KeyframeKeyHash> keyframe_keys_;
  uint64_t max_lm_key_ = 0;  // max over (uint64)(landmark number + 1): bit budget of kernel 3's landmark sort
  bool UpdateMaxLandmarkDevice(const int64_t* d_lm, int64_t n, std::string* err);
  int64_t num_desc_ = 0;  // descriptors in the database (all shards)
  int64_t num_own_ = 0;   // rows of this shard already on the device
  std::vector<float> pend_desc_;      // staged owned rows (n x dim)
  std::vector<int32_t> pend_gidx_;    // their global descriptor indices
  std::vector<int64_t> pend_lm_;      // staged landmark numbers (all rows since the last flush)
  bool FlushPending(std::string* err);
  bool index_dirty_ = true;

  // database (device)
  DeviceLists lists_;
  GrowBuf d_own_desc_;   // float dim per owned descriptor, ascending global index
  GrowBuf d_own_gidx_;   // int32 global descriptor index per owned row
  GrowBuf d_desc_lm_;    // int64 landmark number per descriptor (all shards: replicated)
  DevBuf d_db_cells_;    // int32 cell per owned row (P2)
  DevBuf d_desc_kf_;     // int32 keyframe number per descriptor (replicated; rebuilt from the headers)
  DevBuf d_kf_meta_;     // KeyframeMeta

  // per-call scratch
  DevBuf d_q_, d_cells_, d_idx_, d_dist_, d_bits_, d_stats_;
  DevBuf d_covis_[8];
  DevBuf d_ransac_[4];
  DevBuf d_coarse_scratch_;
  bool CoarseChunks(const float* d_q, int64_t n, int nw, int32_t* d_cells, cudaStream_t stream, std::string* err);
  DevBuf d_query_[4];
  DevBuf d_landmark_xyz_;
  int64_t num_landmark_xyz_ = 0;
  std::vector<int32_t> rnd_host_;
  uint32_t rnd_seed_ = 0;
  int rnd_mapping_ = -1;
  int64_t last_nq_ = 0;
  int last_nw_ = 0;
  bool last_valid_ = false;
  double last_scan_ms_ = 0;
  friend struct DetectorAccess;
};

}  // namespace mlc
