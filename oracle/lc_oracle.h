// lc_oracle.h — CPU ORACLE (test infrastructure, NOT product code).
//
// A dependency-free C++17 restatement of maplab's loop-closure query path
// (SURVEY.md §8a rows A1–A22). Only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py may build, load or call
// anything in oracle/. The product path (maplab_b200/) never links it.
//
// Parity status: PINNED for A5–A10, A12 by the reference's own golden vectors
// (tests/golden/*.json, extracted from the reference's gtest files by
// tests/golden/extract_goldens.py); A18–A21 (GP3P-RANSAC) meet the reference's own
// multi-camera PnP test restated in tests/test_reference_pnp.py, and
// common::transformationRansac the fixture tests of test_geometry.cc
// (tests/test_alignment.py). UNPINNED by any reference test for A1–A4 (projection),
// A11/A13–A17 (Find / covisibility / handler gates): there the anchor is ground truth
// on REAL data — on a fixture cut from maplab's own test map (real BRISK descriptors,
// shipped vocabulary) the whole path relocalises the query vertices within centimetres
// of the poses the map stores (tests/test_real_map.py, tests/test_vi_map_io.py) — see
// DESIGN.md section 2.
//
// All paths cited are relative to /root/reference.
#pragma once
#include <cstdint>
#include <cstddef>
#include <limits>
#include <set>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

namespace lc_oracle {

// ---------------------------------------------------------------------------
// T2: vocabulary file (common::Serialize format).
// algorithms/loopclosure/matching-based-loopclosure/include/
//   matching-based-loopclosure/inverted-multi-index-interface.h:26-47, 59-88
// common/maplab-common/include/maplab-common/binary-serialization.h:25-41,128-161
// ---------------------------------------------------------------------------
struct Matrix {  // column-major float matrix
  int rows = 0, cols = 0;
  std::vector<float> data;
  float& at(int r, int c) { return data[static_cast<size_t>(c) * rows + r]; }
  float at(int r, int c) const { return data[static_cast<size_t>(c) * rows + r]; }
  const float* col(int c) const { return data.data() + static_cast<size_t>(c) * rows; }
};

struct Vocabulary {
  int version = 100;
  int target_dim = 10;
  Matrix projection;  // (>=target_dim) x Kp
  Matrix words1;      // sub_dim x W1
  Matrix words2;      // sub_dim x W2
  // Optional PQ block (serialization version 200).
  bool has_pq = false;
  int pq_num_components = 0, pq_num_centers = 0, pq_dim_per_comp = 0;
  Matrix pq_centers1, pq_centers2;
};
bool ParseVocabulary(const uint8_t* blob, size_t size, bool want_pq, Vocabulary* out,
                     std::string* err);
std::vector<uint8_t> SerializeVocabulary(const Vocabulary& v);

// ---------------------------------------------------------------------------
// A1/A2: bit unpack + projection.
// descriptor-projection/include/descriptor-projection/descriptor-projection.h:45-116
// descriptor-projection/src/descriptor-projection.cc:15-50
// Canonical arithmetic (DESIGN.md "exact projection"): every row d of P is
// quantised to a 27-bit fixed-point grid (3 balanced base-512 digits), the dot
// product with the {0,1} bits is accumulated EXACTLY in integers and rounded
// once to fp32. Order-independent, hence reproducible bit-for-bit on tensor cores.
// ---------------------------------------------------------------------------
struct FixedPointProjection {
  int target_dim = 0, kp = 0;  // kp = number of descriptor bits consumed
  std::vector<int32_t> p_int;  // target_dim x kp, row-major: P_int[d][k]
  std::vector<int> shift;      // per row: P ~= p_int * 2^-shift
};
FixedPointProjection QuantizeProjection(const Matrix& P, int target_dim);
// Split into balanced base-512 digits, each in [-256, 256].
void SplitDigits(int32_t v, int32_t digits[3]);
// raw: bytes_per_desc x n, column-major (one descriptor per column).
// out: target_dim x n column-major.
void ProjectDescriptorBlock(const uint8_t* raw, int bytes_per_desc, int n,
                            const FixedPointProjection& fp, float* out);
// Plain fp32 k-ascending reference (documents the distance to a float GEMM).
void ProjectDescriptorBlockFloat(const uint8_t* raw, int bytes_per_desc, int n,
                                 const Matrix& P, int target_dim, float* out);

// ---------------------------------------------------------------------------
// A6: libnabo kd-tree (KDTreeUnbalancedPtInLeavesImplicitBoundsStackOpt<float,
// IndexHeapBruteForceVector>), dependencies/internal/libnabo/nabo/
// kdtree_cpu.cpp:110-272 (build), :368-447 (recurseKnn), index_heap.h:263-363.
// ---------------------------------------------------------------------------
struct KdNode {
  // Leaf: dim == tree dim, child_or_size = bucket size, bucket_index valid.
  // Inner: dim = cut dimension, child_or_size = right child, cut_val valid.
  uint32_t dim;
  uint32_t child_or_size;
  union {
    float cut_val;
    uint32_t bucket_index;
  };
};
struct KdTree {
  int dim = 0;
  int num_points = 0;
  std::vector<float> cloud;            // dim x n col-major copy
  std::vector<KdNode> nodes;
  std::vector<int> bucket_point_index;  // bucket entry -> point index
  void Build(const float* cloud_col_major, int dim, int n, int bucket_size = 8);
  // k results sorted ascending; unfilled = (-1, +inf). Returns leaf points touched.
  unsigned long Knn(const float* query, int k, float epsilon, float max_radius,
                    int* indices, float* dists2) const;
};

// ---------------------------------------------------------------------------
// A7, A10: imilib/inverted-multi-index-common.h:54-72, :84-134, :148-188
// ---------------------------------------------------------------------------
void InsertNeighbor(int index, float distance, int num_neighbors,
                    std::vector<std::pair<float, int>>* nn);
void MultiSequenceAlgorithm(const int* idx1, const float* d1, int n1, const int* idx2,
                            const float* d2, int n2, int num_words,
                            std::vector<std::pair<int, int>>* closest_words);
struct SearchParams {
  float knn_epsilon = 2.0f;     // FLAGS_lc_knn_epsilon (loopclosure-common/src/flags.cc:9-13)
  float knn_max_radius = 20.0f;  // FLAGS_lc_knn_max_radius
};
void FindClosestWords(const float* query, int sub_dim, int num_closest_words,
                      const KdTree& t1, const KdTree& t2, const SearchParams& sp,
                      std::vector<std::pair<int, int>>* closest_words);

// Squared L2 distance in the canonical fp32 order (DESIGN.md "distance order":
// SSE packet-of-4 accumulation + horizontal add, then scalar remainder).
float SquaredDistance(const float* a, const float* b, int dim);

// ---------------------------------------------------------------------------
// A5, A8: imilib/inverted-multi-index.h:77-94, :100-161
// ---------------------------------------------------------------------------
class InvertedMultiIndex {
 public:
  InvertedMultiIndex(const Matrix& words1, const Matrix& words2, int num_closest_words,
                     const SearchParams& sp = SearchParams());
  void AddDescriptors(const float* desc_col_major, int n);
  void GetNNearestNeighbors(const float* query, int k, int* indices, float* distances) const;
  int GetNumDescriptorsInIndex() const { return max_db_descriptor_index_; }
  void Clear();
  void SetNumClosestWords(int n) { num_closest_words_ = n; }
  // Visited cells (word_index) in order, incl. cells absent from the index (P3).
  void VisitedCells(const float* query, std::vector<int>* cells) const;
  int CellOfDescriptor(const float* desc) const;  // P2
  // Introspection for the golden tests.
  const std::unordered_map<int, int>& word_index_map() const { return word_index_map_; }
  struct InvFile {
    std::vector<float> descriptors;  // dim per entry
    std::vector<int> indices;
  };
  const std::vector<InvFile>& inverted_files() const { return inverted_files_; }
  int dim() const { return 2 * sub_dim_; }
  const KdTree& tree1() const { return t1_; }
  const KdTree& tree2() const { return t2_; }
  int num_words2() const { return w2_; }

 private:
  int sub_dim_, w1_, w2_, num_closest_words_;
  SearchParams sp_;
  KdTree t1_, t2_;
  std::unordered_map<int, int> word_index_map_;
  std::vector<InvFile> inverted_files_;
  int max_db_descriptor_index_ = 0;
};

// ---------------------------------------------------------------------------
// A9: imilib/product-quantization.h:81-152 and
// imilib/inverted-multi-product-quantization-index.h:66-302
// ---------------------------------------------------------------------------
struct ProductQuantizer {  // one per coarse word
  int num_components = 0, dim_per_comp = 0, num_centers = 0;
  std::vector<float> centers;  // dim_per_comp x (num_components*num_centers) col-major
  void Quantize(const float* vec, int* codes) const;
  void FillLUT(const float* vec, float* lut /*num_components x num_centers, row-major*/) const;
  float ComputeDistance(const float* lut, const int* codes) const;
};
class InvertedMultiPQIndex {
 public:
  InvertedMultiPQIndex(const Matrix& words1, const Matrix& words2, const Matrix& qc1,
                       const Matrix& qc2, int num_components, int dim_per_comp,
                       int num_centers, int num_closest_words,
                       const SearchParams& sp = SearchParams());
  void AddDescriptors(const float* desc_col_major, int n);
  void GetNNearestNeighbors(const float* query, int k, int* indices, float* distances) const;
  int GetNumDescriptorsInIndex() const { return max_db_descriptor_index_; }
  void Clear();
  struct InvFile {
    std::vector<int> codes;  // num_components per entry
    std::vector<int> indices;
  };
  const std::unordered_map<int, int>& word_index_map() const { return word_index_map_; }
  const std::vector<InvFile>& inverted_files() const { return inverted_files_; }

 private:
  int sub_dim_, w1_, w2_, ncomp_, half_ncomp_, dim_per_comp_, ncenters_, num_closest_words_;
  SearchParams sp_;
  Matrix words1_, words2_;
  KdTree t1_, t2_;
  std::vector<ProductQuantizer> q1_, q2_;
  std::unordered_map<int, int> word_index_map_;
  std::vector<InvFile> inverted_files_;
  int max_db_descriptor_index_ = 0;
};

// ---------------------------------------------------------------------------
// A12: matching-based-loopclosure/scoring.h:38-59, :92-187
// ---------------------------------------------------------------------------
double BinomialPdf(double n, double p, double k);  // Boost.Math binomial pdf restated
// ids in caller iteration order. scores out, same order.
void ComputeAccumulationScore(const std::vector<size_t>& num_matches, std::vector<float>* scores);
void ComputeProbabilisticScore(const std::vector<size_t>& num_matches,
                               const std::vector<size_t>& num_descriptors_per_id,
                               size_t num_descriptors_in_database, std::vector<float>* scores);

// ---------------------------------------------------------------------------
// A4, A11–A15: matching-based-engine.{h,cc}, -inl.h. Ids are dense integers:
// the shim maps 128-bit HashIds to them.
// ---------------------------------------------------------------------------
struct EngineSettings {
  int num_closest_words_for_nn_search = 10;  // lc_num_words_for_nn_search
  double min_image_time_seconds = 10.0;      // lc_min_image_time_seconds
  size_t min_verify_matches_num = 10;        // lc_min_verify_matches_num
  float fraction_best_scores = 0.25f;        // lc_fraction_best_scores
  int num_nearest_neighbors = -1;            // lc_num_neighbors
  int scoring = 0;                           // 0 accumulation, 1 probabilistic
  int engine = 0;                            // 0 imi, 1 imipq
  SearchParams search;
};
struct ProjectedImage {  // descriptor-projection.h:23-31
  int64_t timestamp_ns = 0;
  int64_t vertex_id = 0;  // dense vertex number
  int frame_index = 0;
  int64_t mission_id = 0;
  int dim = 0;
  std::vector<float> projected_descriptors;  // dim x n col-major
  std::vector<int64_t> landmarks;            // n (database images)
};
struct Match {  // vi_map::FrameKeyPointToStructureMatch + bookkeeping
  int query_frame_index;  // frame index of the query keyframe within its vertex
  int query_keypoint;
  int db_descriptor;  // global descriptor index (canonical tie-break key)
  int db_keyframe;    // insertion number of the result keyframe
  int64_t db_vertex;
  int64_t landmark;
};
class LoopDetector {
 public:
  LoopDetector(const EngineSettings& s, const Vocabulary& v);
  void ProjectDescriptors(const uint8_t* raw, int bytes_per_desc, int n, float* out) const;
  // false = the CHECK of matching-based-engine.cc:244-252 would abort (keyframe id already in the
  // database); nothing is inserted then.
  bool Insert(const ProjectedImage& image);
  void Clear();
  size_t NumEntries() const { return keyframes_.size(); }
  int NumDescriptors() const;
  int NumNeighborsToSearch() const;  // matching-based-engine.cc:319-338
  // All images must belong to one vertex. Output: canonical-order matches
  // (query frame index, keypoint, db descriptor).
  void Find(const std::vector<const ProjectedImage*>& images, std::vector<Match>* out) const;
  // Stage outputs for parity checkpoints P4..P7 of a single query frame.
  struct FrameTrace {
    std::vector<int> knn_indices;      // k x n col-major
    std::vector<float> knn_distances;  // k x n
    std::vector<Match> raw_matches;    // after time filter (P5), scan order
    std::vector<int> cand_keyframes;   // P6: candidate keyframes (ascending)
    std::vector<int> cand_votes;       //     votes per candidate
    std::vector<int> selected_keyframes;  // top-fraction set (ascending)
    std::vector<Match> filtered;       // P7 canonical order
  };
  void FindFrame(const ProjectedImage& q, bool make_unique, FrameTrace* trace) const;
  void KnnBatch(const float* q, int n, int k, int* idx, float* dist) const;
  const FixedPointProjection& fixed_projection() const { return fp_; }

 private:
  struct Keyframe {
    int64_t ts, vertex, mission;
    int frame_index, first_descriptor, num_descriptors;
    std::vector<int64_t> landmarks;
  };
  void CovisFilterKeyframes(const std::vector<Match>& in, bool make_unique,
                            std::vector<Match>* out, FrameTrace* trace) const;
  EngineSettings s_;
  Vocabulary v_;
  FixedPointProjection fp_;
  std::vector<Keyframe> keyframes_;
  std::set<std::pair<int64_t, int>> keyframe_ids_;  // (vertex, frame index): Insert's uniqueness CHECK
  std::vector<int> desc_to_keyframe_;  // global descriptor -> keyframe number
  InvertedMultiIndex* imi_ = nullptr;
  InvertedMultiPQIndex* imipq_ = nullptr;
};
// Generic covisibility component filter on integer ids (A14/A15). group_of(m)
// picks db_keyframe (keyframe pass) or db_vertex (vertex pass).
void CovisComponents(const std::vector<Match>& matches, bool by_vertex,
                     const std::vector<int64_t>* relevant_ids /*null = all*/,
                     size_t min_verify_matches_num, bool make_unique,
                     std::vector<Match>* out);

// ---------------------------------------------------------------------------
// A17–A22: geometric verification.
// ---------------------------------------------------------------------------
struct Camera {  // pinhole; distortion: 0 none, 1 fisheye(FOV) w, 2 equidistant k1..k4, 3 radtan k1 k2 p1 p2
  double fu, fv, cu, cv;
  int distortion = 0;
  double dist[4] = {0, 0, 0, 0};
  double R_B_C[9];  // row-major rotation body<-camera
  double t_B_C[3];
};
void BackProject3(const Camera& c, const double kp[2], double bearing[3]);
double RansacThreshold(const std::vector<Camera>& cams, double pixel_sigma);

// std::mt19937 + uniform_int_distribution<int>(0, INT_MAX) restated (F11).
struct RansacRng {
  explicit RansacRng(uint32_t seed = 12345u, int mapping = 1 /*1: libstdc++>=11, 0: <=10*/);
  int Next();
  uint32_t mt[624];
  int idx;
  int mapping;
  uint32_t NextU32();
};

struct RansacResult {
  bool success = false;
  int iterations = 0;
  int model_indices[4] = {-1, -1, -1, -1};
  double T[12];  // 3x4 row-major [R|t], body pose in G
  std::vector<int> inliers;
  std::vector<double> inlier_distances;
};
// GP3P minimal solver (opengv/src/absolute_pose/modules/main.cpp:375-436).
// f (already rotated into the body frame), v, p are 3x3 column-per-point.
// Returns up to 8 [R|t] (row-major 3x4).
int Gp3pSolve(const double f[9], const double v[9], const double p[9], double solutions[8][12]);
// RANSAC over n correspondences. rand_stream: explicit pre-drawn ints (may be null
// -> RansacRng(seed, mapping)).
void AbsoluteMultiPoseRansac(const double* bearings /*3 x n col-major*/,
                             const int* cam_idx, const double* points /*3 x n*/, int n,
                             const std::vector<Camera>& cams, double threshold,
                             int max_iterations, RansacRng* rng, RansacResult* out);
struct VerifyInput {
  int n = 0;
  std::vector<double> keypoints;  // 2 x n
  std::vector<int> frame_index;   // camera index per match
  std::vector<int> keypoint_index;
  std::vector<double> landmarks;  // 3 x n
};
struct VerifyResult {
  bool accepted = false;
  int num_inliers = 0;
  double inlier_ratio = 0;
  RansacResult ransac;
  std::vector<int> best_inlier_per_keypoint;  // match indices kept (A22), ascending (frame,keypoint)
};
struct HandlerSettings {
  int min_inlier_count = 10;       // lc_min_inlier_count
  double min_inlier_ratio = 0.0;   // lc_min_inlier_ratio
  double ransac_pixel_sigma = 2.0;  // lc_ransac_pixel_sigma
  int num_ransac_iters = 100;      // lc_num_ransac_iters
  uint32_t seed = 12345u;
  int rng_mapping = 1;
};
void HandleLoopClosure(const VerifyInput& in, const std::vector<Camera>& cams,
                       const HandlerSettings& hs, VerifyResult* out);
// Topological gate of handleLoopClosure (loop-closure-handler.cc:424-455): T_I_I_ransac =
// T_G_I(map)^-1 * T_G_I_ransac; reject when its position norm / AngleAxis angle exceed the limits
// (negative limit = check off). Transforms are 3x4 row-major [R|t].
void DeltaPose(const double* T_G_I_map, const double* T_G_I_ransac, double* delta_position_m,
               double* delta_rotation_deg);
bool DeltaPoseGate(const double* T_G_I_map, const double* T_G_I_ransac, double max_delta_position_m,
                   double max_delta_rotation_deg);

// ---------------------------------------------------------------------------
// mission-level alignment (oracle/alignment.cc)
// ---------------------------------------------------------------------------
int UniformIndex(RansacRng* rng, uint32_t n, int mapping);
double AngularDistance(const double* qa_xyzw, const double* qb_xyzw);
bool PoseIsInlier(const double* qa, const double* pa, const double* qb, const double* pb, double thr_rad,
                  double thr_m);
void LsAverageQuaternion(const double* quats, const int* members, int n, double out[4]);
// returns num_inliers; inlier_indices has room for n entries
int TransformationRansac(const double* quats, const double* positions, int n, int num_iterations,
                         double thr_rad, double thr_m, uint32_t seed, int rng_mapping, double out_quat[4],
                         double out_pos[3], int* inlier_indices);
void YawOnly(const double q[4], double out[4]);

}  // namespace lc_oracle
