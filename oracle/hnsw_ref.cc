// ORACLE / test infrastructure: C wrapper around the REFERENCE's own HNSW engine, compiled from the
// sources where they lie (the vendored hnswlib under /root/reference/algorithms/loopclosure/
// matching-based-loopclosure/include/matching-based-loopclosure/hnswlib — header-only, no dependencies)
// into oracle/_ref/libhnsw_ref.so by `make -C oracle ref`. Nothing of the reference is copied into this
// repository. It restates loop_closure::HSNWIndexInterface
// (matching-based-loopclosure/include/matching-based-loopclosure/hnsw-index-interface.h): the same
// HierarchicalNSW<float>(L2Space(dim), N, M, ef_construction), setEf(ef_query), addPoint(descriptor, label)
// and searchKnn, with the result popped from the max-heap exactly as :141-151 does (farthest neighbour
// first). The reference inserts from getNumHardwareThreads() threads at once (:44-66), which makes its
// graph irreproducible; here the points are added one after the other (the single-thread case).
#include <cstddef>
#include <cstdint>
#include <queue>

#include "matching-based-loopclosure/hnswlib/hnswlib.h"

namespace {
struct Index {
  hnswlib::L2Space space;
  hnswlib::HierarchicalNSW<float> graph;
  size_t dim, count = 0;
  Index(size_t d, size_t max_elements, size_t M, size_t ef_construction, size_t ef_query)
      : space(d), graph(&space, max_elements, M, ef_construction), dim(d) {
    graph.setEf(ef_query);
  }
};
}  // namespace

extern "C" {
void* hnsw_ref_create(int dim, int64_t max_elements, int M, int ef_construction, int ef_query) {
  return new Index(static_cast<size_t>(dim), static_cast<size_t>(max_elements), static_cast<size_t>(M),
                   static_cast<size_t>(ef_construction), static_cast<size_t>(ef_query));
}
void hnsw_ref_destroy(void* h) { delete static_cast<Index*>(h); }
// descriptors: n rows of dim floats; labels continue from the number of points already added
void hnsw_ref_add(void* h, const float* descriptors, int64_t n) {
  Index* ix = static_cast<Index*>(h);
  for (int64_t i = 0; i < n; ++i) ix->graph.addPoint(descriptors + i * ix->dim, ix->count++);
}
// idx / dist: n rows of k entries, in the reference's order (result.top() first = farthest first).
// Returns 0, or 1 when a search returned fewer than k results (the reference's CHECK_EQ would abort).
int hnsw_ref_knn(void* h, const float* queries, int64_t n, int k, int32_t* idx, float* dist) {
  Index* ix = static_cast<Index*>(h);
  for (int64_t i = 0; i < n; ++i) {
    std::priority_queue<std::pair<float, size_t>> result = ix->graph.searchKnn(queries + i * ix->dim, k);
    if (static_cast<int>(result.size()) != k) return 1;
    for (int j = 0; j < k; ++j) {
      idx[i * k + j] = static_cast<int32_t>(result.top().second);
      dist[i * k + j] = result.top().first;
      result.pop();
    }
  }
  return 0;
}
}
