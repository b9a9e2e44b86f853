// Host-side vocabulary handling of the B200 loop-closure path: quantizer-file parsing, the exact
// fixed-point form of the projection matrix, and the libnabo-compatible kd-tree over the coarse
// words that the device traverses (kernel 2a).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace mlc {

struct MatrixF {  // column-major, as serialized by common::Serialize(Eigen::Matrix)
  int rows = 0, cols = 0;
  std::vector<float> v;
  float at(int r, int c) const { return v[static_cast<size_t>(c) * rows + r]; }
};

// InvertedMultiIndexVocabulary / InvertedMultiIndexProductVocabulary
// (matching-based-loopclosure/inverted-multi-index-interface.h:26-47, :59-88).
struct VocabularyFile {
  int version = 0, target_dim = 0;
  MatrixF projection, words1, words2;
  bool has_pq = false;
  int pq_components = 0, pq_centers = 0, pq_dim_per_comp = 0;
  MatrixF pq_centers1, pq_centers2;
  bool Parse(const void* blob, size_t size, bool want_pq, std::string* err);
};

// Exact projection: row d of P becomes integers p = rint(P * 2^shift_d), |p| <= 2^26. The dot
// product with the descriptor bits is an exact integer (tensor cores: 4 balanced base-256 int8
// digits, s32 accumulation) rounded ONCE to fp32. See DESIGN.md "Kernel 1".
struct FixedProjection {
  int dim = 0, kp = 0;          // output dims, descriptor bits consumed
  std::vector<int32_t> p_int;   // dim x kp row-major
  std::vector<int32_t> shift;   // per output dim
  void Build(const MatrixF& P, int target_dim);
};
constexpr int kProjDigits = 4;   // int8 digits per fixed-point value
constexpr int kProjNPad = 48;    // UMMA N (>= dim * digits = 40, multiple of 16)
void SplitDigitsBase256(int32_t v, int8_t d[kProjDigits]);

// libnabo KDTreeUnbalancedPtInLeavesImplicitBoundsStackOpt (nabo/kdtree_cpu.cpp:110-272),
// flattened for the device: node i = {dim | leaf flag, right child or bucket size, cut value or
// bucket start}.
struct KdNodeDev {
  uint32_t dim;            // cut dimension; == tree dim for a leaf
  uint32_t child_or_size;  // inner: index of the right child (left child = i + 1); leaf: #points
  uint32_t cut_or_bucket;  // inner: float bits of the cut value; leaf: first bucket entry
};
struct KdTreeHost {
  int dim = 0, num_points = 0, max_depth = 0;
  std::vector<KdNodeDev> nodes;
  std::vector<int32_t> bucket_points;  // bucket entry -> word index
  std::vector<float> cloud;            // dim x n column-major (word per column)
  void Build(const MatrixF& words, int bucket_size = 8);
};

}  // namespace mlc
