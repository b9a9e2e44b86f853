// ORACLE (test infrastructure). Restatement of the libnabo kd-tree that maplab's
// FindClosestWords uses (A6): dependencies/internal/libnabo/nabo/kdtree_cpu.cpp
// :110-272 (buildNodes), :192-238 (ctor), :339-365 (onePointKnn), :368-447
// (recurseKnn); heap: nabo/index_heap.h:263-363; bounds init: nabo/nabo.cpp:68-81.
// All arithmetic is fp32 with sequential adds and no fused multiply-add.
#include <algorithm>
#include <cassert>
#include <cmath>
#include <limits>

#include "lc_oracle.h"

namespace lc_oracle {
namespace {

struct Builder {
  KdTree* t;
  int bucket_size;
  std::vector<int> pts;

  float Coeff(int d, int idx) const { return t->cloud[static_cast<size_t>(idx) * t->dim + d]; }

  // kdtree_cpu.cpp:110-236
  unsigned BuildNodes(int first, int last, std::vector<float> min_values,
                      std::vector<float> max_values) {
    const int count = last - first;
    const unsigned pos = static_cast<unsigned>(t->nodes.size());
    if (count <= bucket_size) {
      const uint32_t init_buckets_size = static_cast<uint32_t>(t->bucket_point_index.size());
      for (int i = 0; i < count; ++i) t->bucket_point_index.push_back(pts[first + i]);
      KdNode n;
      n.dim = static_cast<uint32_t>(t->dim);
      n.child_or_size = static_cast<uint32_t>(count);
      n.bucket_index = init_buckets_size;
      t->nodes.push_back(n);
      return pos;
    }
    // argMax over (max - min): first strictly larger than 0 wins (kdtree_cpu.cpp:78-92).
    unsigned cut_dim = 0;
    {
      float max_val = 0.f;
      for (int i = 0; i < t->dim; ++i) {
        const float v = max_values[i] - min_values[i];
        if (v > max_val) {
          max_val = v;
          cut_dim = static_cast<unsigned>(i);
        }
      }
    }
    const float ideal_cut_val = (max_values[cut_dim] + min_values[cut_dim]) / 2;
    float min_v = std::numeric_limits<float>::max();
    float max_v = std::numeric_limits<float>::lowest();
    for (int i = first; i < last; ++i) {
      const float val = Coeff(cut_dim, pts[i]);
      min_v = std::min(val, min_v);
      max_v = std::max(val, max_v);
    }
    float cut_val;
    if (ideal_cut_val < min_v)
      cut_val = min_v;
    else if (ideal_cut_val > max_v)
      cut_val = max_v;
    else
      cut_val = ideal_cut_val;

    int l = 0, r = count - 1;
    while (true) {
      while (l < count && Coeff(cut_dim, pts[first + l]) < cut_val) ++l;
      while (r >= 0 && Coeff(cut_dim, pts[first + r]) >= cut_val) --r;
      if (l > r) break;
      std::swap(pts[first + l], pts[first + r]);
      ++l;
      --r;
    }
    const int br1 = l;
    r = count - 1;
    while (true) {
      while (l < count && Coeff(cut_dim, pts[first + l]) <= cut_val) ++l;
      while (r >= br1 && Coeff(cut_dim, pts[first + r]) > cut_val) --r;
      if (l > r) break;
      std::swap(pts[first + l], pts[first + r]);
      ++l;
      --r;
    }
    const int br2 = l;
    int left_count;
    if (ideal_cut_val < min_v)
      left_count = 1;
    else if (ideal_cut_val > max_v)
      left_count = count - 1;
    else if (br1 > count / 2)
      left_count = br1;
    else if (br2 < count / 2)
      left_count = br2;
    else
      left_count = count / 2;
    assert(left_count > 0 && left_count < count);

    std::vector<float> left_max(max_values);
    left_max[cut_dim] = cut_val;
    std::vector<float> right_min(min_values);
    right_min[cut_dim] = cut_val;

    KdNode n;
    n.dim = 0;
    n.child_or_size = 0;
    n.cut_val = cut_val;
    t->nodes.push_back(n);
    BuildNodes(first, first + left_count, min_values, left_max);
    const unsigned right_child = BuildNodes(first + left_count, last, right_min, max_values);
    t->nodes[pos].dim = cut_dim;
    t->nodes[pos].child_or_size = right_child;
    return pos;
  }
};

// IndexHeapBruteForceVector (index_heap.h:263-363): a sorted array; head = last.
struct LinearHeap {
  int k;
  int* idx;
  float* val;
  void Reset() {
    for (int i = 0; i < k; ++i) {
      idx[i] = -1;
      val[i] = std::numeric_limits<float>::infinity();
    }
  }
  float Head() const { return val[k - 1]; }
  void ReplaceHead(int index, float value) {
    int i;
    for (i = k - 1; i > 0; --i) {
      if (val[i - 1] > value) {
        val[i] = val[i - 1];
        idx[i] = idx[i - 1];
      } else {
        break;
      }
    }
    val[i] = value;
    idx[i] = index;
  }
};

struct Searcher {
  const KdTree* t;
  const float* query;
  LinearHeap heap;
  float off[16];
  float max_error2, max_radius2;
  unsigned long touched = 0;

  // kdtree_cpu.cpp:368-447, allowSelfMatch = true, no statistics.
  void Recurse(unsigned n, float rd) {
    const KdNode& node = t->nodes[n];
    const uint32_t cd = node.dim;
    if (cd == static_cast<uint32_t>(t->dim)) {
      const uint32_t bs = node.child_or_size;
      for (uint32_t i = 0; i < bs; ++i) {
        const int pidx = t->bucket_point_index[node.bucket_index + i];
        const float* d_ptr = &t->cloud[static_cast<size_t>(pidx) * t->dim];
        float dist = 0;
        for (int j = 0; j < t->dim; ++j) {
          const float diff = query[j] - d_ptr[j];
          dist += diff * diff;
        }
        if ((dist <= max_radius2) && (dist < heap.Head())) heap.ReplaceHead(pidx, dist);
      }
      touched += bs;
      return;
    }
    const unsigned right_child = node.child_or_size;
    const float old_off = off[cd];
    const float new_off = query[cd] - node.cut_val;
    if (new_off > 0) {
      Recurse(right_child, rd);
      rd += -old_off * old_off + new_off * new_off;
      if ((rd <= max_radius2) && (rd * max_error2 < heap.Head())) {
        off[cd] = new_off;
        Recurse(n + 1, rd);
        off[cd] = old_off;
      }
    } else {
      Recurse(n + 1, rd);
      rd += -old_off * old_off + new_off * new_off;
      if ((rd <= max_radius2) && (rd * max_error2 < heap.Head())) {
        off[cd] = new_off;
        Recurse(right_child, rd);
        off[cd] = old_off;
      }
    }
  }
};

}  // namespace

void KdTree::Build(const float* cloud_col_major, int dim_in, int n, int bucket_size) {
  dim = dim_in;
  num_points = n;
  cloud.assign(cloud_col_major, cloud_col_major + static_cast<size_t>(dim) * n);
  nodes.clear();
  bucket_point_index.clear();
  assert(dim <= 16);
  if (n <= bucket_size) {  // single-bucket tree, kdtree_cpu.cpp:201-208
    for (int i = 0; i < n; ++i) bucket_point_index.push_back(i);
    KdNode nd;
    nd.dim = static_cast<uint32_t>(dim);
    nd.child_or_size = static_cast<uint32_t>(n);
    nd.bucket_index = 0;
    nodes.push_back(nd);
    return;
  }
  // nabo.cpp:72-75: minBound = +max, maxBound = numeric_limits<T>::min() (tiny positive!).
  std::vector<float> min_bound(dim, std::numeric_limits<float>::max());
  std::vector<float> max_bound(dim, std::numeric_limits<float>::min());
  Builder b;
  b.t = this;
  b.bucket_size = bucket_size;
  b.pts.reserve(n);
  for (int i = 0; i < n; ++i) {
    b.pts.push_back(i);
    for (int d = 0; d < dim; ++d) {
      const float v = cloud[static_cast<size_t>(i) * dim + d];
      min_bound[d] = std::min(min_bound[d], v);
      max_bound[d] = std::max(max_bound[d], v);
    }
  }
  b.BuildNodes(0, n, min_bound, max_bound);
}

unsigned long KdTree::Knn(const float* query, int k, float epsilon, float max_radius,
                          int* indices, float* dists2) const {
  Searcher s;
  s.t = this;
  s.query = query;
  s.heap = LinearHeap{k, indices, dists2};
  s.heap.Reset();
  for (int d = 0; d < 16; ++d) s.off[d] = 0.f;
  s.max_radius2 = max_radius * max_radius;
  s.max_error2 = (1 + epsilon) * (1 + epsilon);
  s.Recurse(0, 0.f);
  return s.touched;
}

}  // namespace lc_oracle
