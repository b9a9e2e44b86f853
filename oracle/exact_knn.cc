// ORACLE (test infrastructure): exact k nearest neighbours of float descriptors by squared L2 distance —
// the ground truth of the `hnsw` engine slot (loop_closure::HSNWIndexInterface answers the same question
// approximately, hnsw-index-interface.h:119-155; hnswlib's scalar distance is space_l2.h:6-20: the sum of
// squared differences accumulated in index order). Result order as the reference interface returns it:
// popped from a max-heap of (distance, label) pairs, i.e. descending (distance, label).
#include <algorithm>
#include <cstdint>
#include <utility>
#include <vector>

extern "C" int lco_exact_knn(const float* db, int64_t n_db, const float* q, int64_t n_q, int dim, int k,
                             int32_t* idx, float* dist) {
  if (k > n_db) return 1;  // the reference CHECKs result.size() == num_neighbors
  std::vector<std::pair<float, int32_t>> all(static_cast<size_t>(n_db));
  for (int64_t i = 0; i < n_q; ++i) {
    const float* qi = q + i * dim;
    for (int64_t j = 0; j < n_db; ++j) {
      const float* x = db + j * dim;
      float acc = 0.f;
      for (int d = 0; d < dim; ++d) {
        const float t = qi[d] - x[d];
        acc += t * t;  // -ffp-contract=off: multiply, then add
      }
      all[static_cast<size_t>(j)] = {acc, static_cast<int32_t>(j)};
    }
    std::partial_sort(all.begin(), all.begin() + k, all.end());  // ascending (distance, label)
    for (int r = 0; r < k; ++r) {
      idx[i * k + r] = all[static_cast<size_t>(k - 1 - r)].second;
      dist[i * k + r] = all[static_cast<size_t>(k - 1 - r)].first;
    }
  }
  return 0;
}
