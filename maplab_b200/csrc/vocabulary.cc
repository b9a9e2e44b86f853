#include "vocabulary.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

namespace mlc {
namespace {

class BlobReader {
 public:
  BlobReader(const void* p, size_t n) : p_(static_cast<const uint8_t*>(p)), left_(n) {}
  bool Int(int* out) {
    if (left_ < sizeof(int)) return false;
    std::memcpy(out, p_, sizeof(int));
    p_ += sizeof(int);
    left_ -= sizeof(int);
    return true;
  }
  // common::Deserialize(Eigen::Matrix): int rows, int cols, raw column-major scalars
  // (maplab-common/binary-serialization.h:128-161).
  bool Mat(MatrixF* m) {
    int r = 0, c = 0;
    if (!Int(&r) || !Int(&c) || r < 0 || c < 0) return false;
    const size_t count = static_cast<size_t>(r) * static_cast<size_t>(c);
    if (left_ / sizeof(float) < count) return false;
    m->rows = r;
    m->cols = c;
    m->v.resize(count);
    std::memcpy(m->v.data(), p_, count * sizeof(float));
    p_ += count * sizeof(float);
    left_ -= count * sizeof(float);
    return true;
  }

 private:
  const uint8_t* p_;
  size_t left_;
};

}  // namespace

bool VocabularyFile::Parse(const void* blob, size_t size, bool want_pq, std::string* err) {
  BlobReader r(blob, size);
  if (!r.Int(&version) || !r.Int(&target_dim)) {
    *err = "vocabulary: truncated header";
    return false;
  }
  if (!r.Mat(&projection) || !r.Mat(&words1) || !r.Mat(&words2)) {
    *err = "vocabulary: truncated projection / word matrices";
    return false;
  }
  if (target_dim <= 0 || target_dim % 2 != 0 || projection.rows < target_dim ||
      words1.rows != target_dim / 2 || words2.rows != target_dim / 2 || words1.cols <= 0 ||
      words2.cols <= 0) {
    *err = "vocabulary: inconsistent dimensions";
    return false;
  }
  has_pq = false;
  if (want_pq) {
    int tag = 0;
    // InvertedMultiIndexProductVocabulary::Load checks its own serialization version (200).
    if (!r.Int(&tag) || tag != 200) {
      *err = "This vocabulary file was saved with a different version.";
      return false;
    }
    if (!r.Int(&pq_components) || !r.Int(&pq_centers) || !r.Int(&pq_dim_per_comp) ||
        !r.Mat(&pq_centers1) || !r.Mat(&pq_centers2)) {
      *err = "vocabulary: truncated product-quantizer block";
      return false;
    }
    has_pq = true;
  }
  return true;
}

void FixedProjection::Build(const MatrixF& P, int target_dim) {
  dim = target_dim;
  kp = P.cols;
  p_int.assign(static_cast<size_t>(dim) * kp, 0);
  shift.assign(dim, 0);
  for (int d = 0; d < dim; ++d) {
    float row_max = 0.f;
    for (int k = 0; k < kp; ++k) row_max = std::fmax(row_max, std::fabs(P.at(d, k)));
    int e = 0;  // smallest exponent with row_max <= 2^e
    if (row_max > 0.f) {
      int ex = 0;
      const float mant = std::frexp(row_max, &ex);
      e = (mant == 0.5f) ? ex - 1 : ex;
    }
    shift[d] = 26 - e;
    for (int k = 0; k < kp; ++k) {
      const double scaled = std::ldexp(static_cast<double>(P.at(d, k)), shift[d]);
      p_int[static_cast<size_t>(d) * kp + k] = static_cast<int32_t>(std::nearbyint(scaled));
    }
  }
}

void SplitDigitsBase256(int32_t v, int8_t d[kProjDigits]) {
  int64_t rest = v;
  for (int j = 0; j < kProjDigits - 1; ++j) {
    int64_t low = ((rest % 256) + 256) % 256;
    if (low >= 128) low -= 256;
    d[j] = static_cast<int8_t>(low);
    rest = (rest - low) / 256;
  }
  d[kProjDigits - 1] = static_cast<int8_t>(rest);  // |v| <= 2^26 => |rest| <= 5
}

// ---------------------------------------------------------------------------------------------
// kd-tree build. Same splitting rule as libnabo (sliding midpoint on the widest dimension of the
// node's implicit bounds, two partition passes, bucket size 8) so that the device traversal
// visits leaves in the reference's order.
// ---------------------------------------------------------------------------------------------
namespace {
struct TreeBuilder {
  KdTreeHost* tree;
  int bucket_size;
  std::vector<int32_t> order;  // permutation of the points, partitioned in place

  float Value(int d, int point) const {
    return tree->cloud[static_cast<size_t>(point) * tree->dim + d];
  }

  uint32_t Node(int first, int last, std::vector<float> lo, std::vector<float> hi, int depth) {
    tree->max_depth = std::max(tree->max_depth, depth);
    const int count = last - first;
    const uint32_t self = static_cast<uint32_t>(tree->nodes.size());
    if (count <= bucket_size) {
      KdNodeDev leaf;
      leaf.dim = static_cast<uint32_t>(tree->dim);
      leaf.child_or_size = static_cast<uint32_t>(count);
      leaf.cut_or_bucket = static_cast<uint32_t>(tree->bucket_points.size());
      for (int i = first; i < last; ++i) tree->bucket_points.push_back(order[i]);
      tree->nodes.push_back(leaf);
      return self;
    }
    // Widest side of the implicit bounds; only a strictly positive extent can win.
    int cut_dim = 0;
    float widest = 0.f;
    for (int d = 0; d < tree->dim; ++d) {
      const float extent = hi[d] - lo[d];
      if (extent > widest) {
        widest = extent;
        cut_dim = d;
      }
    }
    const float ideal = (hi[cut_dim] + lo[cut_dim]) / 2;
    float vmin = std::numeric_limits<float>::max();
    float vmax = std::numeric_limits<float>::lowest();
    for (int i = first; i < last; ++i) {
      const float x = Value(cut_dim, order[i]);
      vmin = std::min(x, vmin);
      vmax = std::max(x, vmax);
    }
    const float cut = ideal < vmin ? vmin : (ideal > vmax ? vmax : ideal);

    int32_t* p = order.data() + first;
    // pass 1: [< cut | >= cut]
    int l = 0, r = count - 1;
    for (;;) {
      while (l < count && Value(cut_dim, p[l]) < cut) ++l;
      while (r >= 0 && Value(cut_dim, p[r]) >= cut) --r;
      if (l > r) break;
      std::swap(p[l], p[r]);
      ++l;
      --r;
    }
    const int below = l;
    // pass 2 on the upper part: [== cut | > cut]
    r = count - 1;
    for (;;) {
      while (l < count && Value(cut_dim, p[l]) <= cut) ++l;
      while (r >= below && Value(cut_dim, p[r]) > cut) --r;
      if (l > r) break;
      std::swap(p[l], p[r]);
      ++l;
      --r;
    }
    const int below_or_equal = l;
    int left_count;
    if (ideal < vmin) {
      left_count = 1;
    } else if (ideal > vmax) {
      left_count = count - 1;
    } else if (below > count / 2) {
      left_count = below;
    } else if (below_or_equal < count / 2) {
      left_count = below_or_equal;
    } else {
      left_count = count / 2;
    }

    KdNodeDev inner;
    inner.dim = static_cast<uint32_t>(cut_dim);
    inner.child_or_size = 0;
    std::memcpy(&inner.cut_or_bucket, &cut, sizeof(float));
    tree->nodes.push_back(inner);
    std::vector<float> left_hi(hi), right_lo(lo);
    left_hi[cut_dim] = cut;
    right_lo[cut_dim] = cut;
    Node(first, first + left_count, lo, left_hi, depth + 1);
    const uint32_t right = Node(first + left_count, last, right_lo, hi, depth + 1);
    tree->nodes[self].child_or_size = right;
    return self;
  }
};
}  // namespace

void KdTreeHost::Build(const MatrixF& words, int bucket_size) {
  dim = words.rows;
  num_points = words.cols;
  cloud = words.v;
  nodes.clear();
  bucket_points.clear();
  max_depth = 0;
  if (num_points <= bucket_size) {
    KdNodeDev leaf;
    leaf.dim = static_cast<uint32_t>(dim);
    leaf.child_or_size = static_cast<uint32_t>(num_points);
    leaf.cut_or_bucket = 0;
    for (int i = 0; i < num_points; ++i) bucket_points.push_back(i);
    nodes.push_back(leaf);
    return;
  }
  // libnabo initialises the upper bound with numeric_limits<T>::min() (a tiny POSITIVE number,
  // nabo/nabo.cpp:72-75), which matters when all coordinates of a dimension are negative.
  std::vector<float> lo(dim, std::numeric_limits<float>::max());
  std::vector<float> hi(dim, std::numeric_limits<float>::min());
  TreeBuilder b;
  b.tree = this;
  b.bucket_size = bucket_size;
  b.order.resize(num_points);
  for (int i = 0; i < num_points; ++i) {
    b.order[i] = i;
    for (int d = 0; d < dim; ++d) {
      const float x = cloud[static_cast<size_t>(i) * dim + d];
      lo[d] = std::min(lo[d], x);
      hi[d] = std::max(hi[d], x);
    }
  }
  b.Node(0, num_points, lo, hi, 1);
}

}  // namespace mlc
