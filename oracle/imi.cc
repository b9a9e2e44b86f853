// ORACLE (test infrastructure). Inverted multi-index (A5, A7, A8, A10) and the
// product-quantised variant (A9). Reference files (relative to
// algorithms/loopclosure/matching-based-loopclosure/include/matching-based-loopclosure/imilib/):
//   inverted-multi-index-common.h, inverted-multi-index.h,
//   inverted-multi-product-quantization-index.h, product-quantization.h
#include <algorithm>
#include <cassert>
#include <cmath>
#include <functional>
#include <limits>
#include <queue>
#include <tuple>

#include "lc_oracle.h"

namespace lc_oracle {

// inverted-multi-index-common.h:54-72
void InsertNeighbor(int index, float distance, int num_neighbors,
                    std::vector<std::pair<float, int>>* nn) {
  const int num_found = static_cast<int>(nn->size());
  if (num_found >= num_neighbors) {
    if (nn->back().first < distance) return;
  }
  const std::pair<float, int> neighbor(distance, index);
  auto it = std::lower_bound(nn->begin(), nn->end(), neighbor);
  nn->insert(it, neighbor);
  if (num_found >= num_neighbors) nn->resize(num_neighbors);
}

// inverted-multi-index-common.h:84-134
void MultiSequenceAlgorithm(const int* idx1, const float* d1, int n1, const int* idx2,
                            const float* d2, int n2, int num_words,
                            std::vector<std::pair<int, int>>* closest_words) {
  closest_words->clear();
  std::vector<bool> pair_used(static_cast<size_t>(n1) * n2, false);
  typedef std::tuple<float, int, int> T;
  std::priority_queue<T, std::vector<T>, std::greater<T>> pq;
  pq.emplace(d1[0] + d2[0], 0, 0);
  while (!pq.empty() && static_cast<int>(closest_words->size()) < num_words) {
    const int i1 = std::get<1>(pq.top());
    const int i2 = std::get<2>(pq.top());
    const int word_index = i1 * n2 + i2;
    pair_used[word_index] = true;
    closest_words->emplace_back(idx1[i1], idx2[i2]);
    pq.pop();
    if ((i1 + 1) < n1) {
      if (i2 == 0 || pair_used[word_index + n2 - 1]) pq.emplace(d1[i1 + 1] + d2[i2], i1 + 1, i2);
    }
    if ((i2 + 1) < n2) {
      if (i1 == 0 || pair_used[word_index - n2 + 1]) pq.emplace(d1[i1] + d2[i2 + 1], i1, i2 + 1);
    }
  }
}

// inverted-multi-index-common.h:148-188
void FindClosestWords(const float* query, int sub_dim, int num_closest_words, const KdTree& t1,
                      const KdTree& t2, const SearchParams& sp,
                      std::vector<std::pair<int, int>>* closest_words) {
  assert(num_closest_words > 0);
  const int k1 = std::min(t1.num_points, num_closest_words);
  const int k2 = std::min(t2.num_points, num_closest_words);
  std::vector<int> i1(k1), i2(k2);
  std::vector<float> d1(k1), d2(k2);
  t1.Knn(query, k1, sp.knn_epsilon, sp.knn_max_radius, i1.data(), d1.data());
  t2.Knn(query + sub_dim, k2, sp.knn_epsilon, sp.knn_max_radius, i2.data(), d2.data());
  MultiSequenceAlgorithm(i1.data(), d1.data(), k1, i2.data(), d2.data(), k2, num_closest_words,
                         closest_words);
}

// (a - b).squaredNorm() for a fixed-size float vector, Eigen 3.3 SSE3 order:
// packets of four squared differences are added lane-wise, reduced with two
// horizontal adds ((l0+l1)+(l2+l3)), the scalar remainder is accumulated
// left-to-right and added last. Eigen's exact order is not pinned by any
// reference test; every admissible order is within 1e-6 relative.
float SquaredDistance(const float* a, const float* b, int dim) {
  const int vec = (dim / 4) * 4;
  if (vec == 0) {
    float s = 0.f;
    for (int i = 0; i < dim; ++i) {
      const float d = a[i] - b[i];
      const float sq = d * d;
      s = (i == 0) ? sq : s + sq;
    }
    return s;
  }
  float lane[4];
  for (int l = 0; l < 4; ++l) {
    const float d = a[l] - b[l];
    lane[l] = d * d;
  }
  for (int p = 4; p < vec; p += 4) {
    for (int l = 0; l < 4; ++l) {
      const float d = a[p + l] - b[p + l];
      lane[l] = lane[l] + d * d;
    }
  }
  float res = (lane[0] + lane[1]) + (lane[2] + lane[3]);
  if (vec < dim) {
    float rem = 0.f;
    for (int i = vec; i < dim; ++i) {
      const float d = a[i] - b[i];
      const float sq = d * d;
      rem = (i == vec) ? sq : rem + sq;
    }
    res = res + rem;
  }
  return res;
}

// ------------------------------- IMI ---------------------------------------
InvertedMultiIndex::InvertedMultiIndex(const Matrix& words1, const Matrix& words2,
                                       int num_closest_words, const SearchParams& sp)
    : sub_dim_(words1.rows),
      w1_(words1.cols),
      w2_(words2.cols),
      num_closest_words_(num_closest_words),
      sp_(sp) {
  assert(words1.rows == words2.rows && words1.cols > 0 && words2.cols > 0);
  t1_.Build(words1.data.data(), sub_dim_, w1_);
  t2_.Build(words2.data.data(), sub_dim_, w2_);
}

void InvertedMultiIndex::Clear() {
  inverted_files_.clear();
  word_index_map_.clear();
  max_db_descriptor_index_ = 0;
}

int InvertedMultiIndex::CellOfDescriptor(const float* desc) const {
  std::vector<std::pair<int, int>> cw;
  FindClosestWords(desc, sub_dim_, 1, t1_, t2_, sp_, &cw);
  assert(!cw.empty());
  // Quirk 1: no word inside lc_knn_max_radius in one of the halves. The reference
  // would compute an aliased / negative cell id; the canonical definition is "not indexed".
  if (cw[0].first < 0 || cw[0].second < 0) return -1;
  return cw[0].first * w2_ + cw[0].second;
}

// inverted-multi-index.h:77-94 + common.h:202-229 (AddDescriptor)
void InvertedMultiIndex::AddDescriptors(const float* desc, int n) {
  const int dim = 2 * sub_dim_;
  for (int i = 0; i < n; ++i) {
    const float* d = desc + static_cast<size_t>(i) * dim;
    const int word_index = CellOfDescriptor(d);
    if (word_index < 0) {  // quirk 1: unreachable descriptor keeps its index but is not stored
      ++max_db_descriptor_index_;
      continue;
    }
    auto it = word_index_map_.find(word_index);
    if (it == word_index_map_.end()) {
      word_index_map_.emplace(word_index, static_cast<int>(inverted_files_.size()));
      InvFile f;
      f.descriptors.assign(d, d + dim);
      f.indices.push_back(max_db_descriptor_index_);
      inverted_files_.push_back(std::move(f));
    } else {
      InvFile& f = inverted_files_[it->second];
      f.descriptors.insert(f.descriptors.end(), d, d + dim);
      f.indices.push_back(max_db_descriptor_index_);
    }
    ++max_db_descriptor_index_;
  }
}

void InvertedMultiIndex::VisitedCells(const float* query, std::vector<int>* cells) const {
  std::vector<std::pair<int, int>> cw;
  FindClosestWords(query, sub_dim_, num_closest_words_, t1_, t2_, sp_, &cw);
  cells->clear();
  for (const auto& w : cw) {
    // Quirk 1: pairs with a -1 member (fewer than nw words within
    // the radius) are skipped instead of aliasing another cell.
    if (w.first < 0 || w.second < 0) {
      cells->push_back(-1);
      continue;
    }
    cells->push_back(w.first * w2_ + w.second);
  }
}

// inverted-multi-index.h:100-161
void InvertedMultiIndex::GetNNearestNeighbors(const float* query, int k, int* indices,
                                              float* distances) const {
  const int dim = 2 * sub_dim_;
  std::vector<int> cells;
  VisitedCells(query, &cells);
  std::vector<std::pair<float, int>> nn;
  nn.reserve(k + 1);
  for (int cell : cells) {
    if (cell < 0) continue;
    auto it = word_index_map_.find(cell);
    if (it == word_index_map_.end()) continue;
    const InvFile& f = inverted_files_[it->second];
    const size_t num = f.indices.size();
    for (size_t j = 0; j < num; ++j) {
      const float dist = SquaredDistance(&f.descriptors[j * dim], query, dim);
      InsertNeighbor(f.indices[j], dist, k, &nn);
    }
  }
  for (size_t i = 0; i < nn.size(); ++i) {
    indices[i] = nn[i].second;
    distances[i] = nn[i].first;
  }
  for (int i = static_cast<int>(nn.size()); i < k; ++i) {
    indices[i] = -1;
    distances[i] = std::numeric_limits<float>::infinity();
  }
}

// ------------------------------- PQ ----------------------------------------
// product-quantization.h:81-103: argmin over centres of the squared distance,
// first minimum wins (Eigen minCoeff visitor).
void ProductQuantizer::Quantize(const float* vec, int* codes) const {
  for (int j = 0; j < num_components; ++j) {
    const float* comp = vec + j * dim_per_comp;
    int best = 0;
    float best_d = 0.f;
    for (int c = 0; c < num_centers; ++c) {
      const float* ctr = &centers[static_cast<size_t>(j * num_centers + c) * dim_per_comp];
      const float d = SquaredDistance(ctr, comp, dim_per_comp);
      if (c == 0 || d < best_d) {
        best_d = d;
        best = c;
      }
    }
    codes[j] = best;
  }
}
// product-quantization.h:107-125
void ProductQuantizer::FillLUT(const float* vec, float* lut) const {
  for (int i = 0; i < num_components; ++i) {
    const float* comp = vec + i * dim_per_comp;
    for (int c = 0; c < num_centers; ++c) {
      const float* ctr = &centers[static_cast<size_t>(i * num_centers + c) * dim_per_comp];
      lut[i * num_centers + c] = SquaredDistance(ctr, comp, dim_per_comp);
    }
  }
}
// product-quantization.h:144-152: sequential sum from 0.0f.
float ProductQuantizer::ComputeDistance(const float* lut, const int* codes) const {
  float dist = 0.0f;
  for (int j = 0; j < num_components; ++j) dist += lut[j * num_centers + codes[j]];
  return dist;
}

InvertedMultiPQIndex::InvertedMultiPQIndex(const Matrix& words1, const Matrix& words2,
                                           const Matrix& qc1, const Matrix& qc2,
                                           int num_components, int dim_per_comp,
                                           int num_centers, int num_closest_words,
                                           const SearchParams& sp)
    : sub_dim_(words1.rows),
      w1_(words1.cols),
      w2_(words2.cols),
      ncomp_(num_components),
      half_ncomp_(num_components / 2),
      dim_per_comp_(dim_per_comp),
      ncenters_(num_centers),
      num_closest_words_(num_closest_words),
      sp_(sp),
      words1_(words1),
      words2_(words2) {
  assert(num_components % 2 == 0);
  assert(sub_dim_ == half_ncomp_ * dim_per_comp_);
  t1_.Build(words1.data.data(), sub_dim_, w1_);
  t2_.Build(words2.data.data(), sub_dim_, w2_);
  const int cols_per_pq = half_ncomp_ * ncenters_;
  assert(qc1.rows == dim_per_comp && qc1.cols == cols_per_pq * w1_);
  assert(qc2.rows == dim_per_comp && qc2.cols == cols_per_pq * w2_);
  auto fill = [&](const Matrix& qc, int nwords, std::vector<ProductQuantizer>* q) {
    q->resize(nwords);
    for (int i = 0; i < nwords; ++i) {
      ProductQuantizer& pq = (*q)[i];
      pq.num_components = half_ncomp_;
      pq.dim_per_comp = dim_per_comp_;
      pq.num_centers = ncenters_;
      const float* src = qc.data.data() + static_cast<size_t>(i) * cols_per_pq * dim_per_comp_;
      pq.centers.assign(src, src + static_cast<size_t>(cols_per_pq) * dim_per_comp_);
    }
  };
  fill(qc1, w1_, &q1_);
  fill(qc2, w2_, &q2_);  // quirk 6: the reference loops to words_1.cols(); assumes W1 == W2
}

void InvertedMultiPQIndex::Clear() {
  inverted_files_.clear();
  word_index_map_.clear();
  max_db_descriptor_index_ = 0;
}

// inverted-multi-product-quantization-index.h:136-175
void InvertedMultiPQIndex::AddDescriptors(const float* desc, int n) {
  const int dim = 2 * sub_dim_;
  std::vector<std::pair<int, int>> cw;
  std::vector<float> res(sub_dim_);
  std::vector<int> codes(ncomp_);
  for (int i = 0; i < n; ++i) {
    const float* d = desc + static_cast<size_t>(i) * dim;
    FindClosestWords(d, sub_dim_, 1, t1_, t2_, sp_, &cw);
    const int word1 = cw[0].first, word2 = cw[0].second;
    if (word1 < 0 || word2 < 0) {  // quirk 1, as in InvertedMultiIndex::AddDescriptors
      ++max_db_descriptor_index_;
      continue;
    }
    const int word_index = word1 * w2_ + word2;
    for (int j = 0; j < sub_dim_; ++j) res[j] = d[j] - words1_.at(j, word1);
    q1_[word1].Quantize(res.data(), codes.data());
    for (int j = 0; j < sub_dim_; ++j) res[j] = d[sub_dim_ + j] - words2_.at(j, word2);
    q2_[word2].Quantize(res.data(), codes.data() + half_ncomp_);
    auto it = word_index_map_.find(word_index);
    if (it == word_index_map_.end()) {
      word_index_map_.emplace(word_index, static_cast<int>(inverted_files_.size()));
      InvFile f;
      f.codes = codes;
      f.indices.push_back(max_db_descriptor_index_);
      inverted_files_.push_back(std::move(f));
    } else {
      InvFile& f = inverted_files_[it->second];
      f.codes.insert(f.codes.end(), codes.begin(), codes.end());
      f.indices.push_back(max_db_descriptor_index_);
    }
    ++max_db_descriptor_index_;
  }
}

// inverted-multi-product-quantization-index.h:181-288
void InvertedMultiPQIndex::GetNNearestNeighbors(const float* query, int k, int* indices,
                                                float* distances) const {
  std::vector<std::pair<int, int>> cw;
  FindClosestWords(query, sub_dim_, num_closest_words_, t1_, t2_, sp_, &cw);
  std::vector<std::pair<float, int>> nn;
  nn.reserve(k + 1);
  std::unordered_map<int, std::vector<float>> cache1, cache2;
  std::vector<float> res(sub_dim_);
  for (const auto& w : cw) {
    const int word1 = w.first, word2 = w.second;
    if (word1 < 0 || word2 < 0) continue;  // quirk 1
    auto it = word_index_map_.find(word1 * w2_ + word2);
    if (it == word_index_map_.end()) continue;
    auto c1 = cache1.find(word1);
    if (c1 == cache1.end()) {
      for (int j = 0; j < sub_dim_; ++j) res[j] = query[j] - words1_.at(j, word1);
      std::vector<float> lut(static_cast<size_t>(half_ncomp_) * ncenters_);
      q1_[word1].FillLUT(res.data(), lut.data());
      c1 = cache1.emplace(word1, std::move(lut)).first;
    }
    auto c2 = cache2.find(word2);
    if (c2 == cache2.end()) {
      for (int j = 0; j < sub_dim_; ++j) res[j] = query[sub_dim_ + j] - words2_.at(j, word2);
      std::vector<float> lut(static_cast<size_t>(half_ncomp_) * ncenters_);
      q2_[word2].FillLUT(res.data(), lut.data());
      c2 = cache2.emplace(word2, std::move(lut)).first;
    }
    const InvFile& f = inverted_files_[it->second];
    const int num = static_cast<int>(f.indices.size());
    for (int j = 0; j < num; ++j) {
      const int* codes = &f.codes[static_cast<size_t>(j) * ncomp_];
      float distance = q1_[word1].ComputeDistance(c1->second.data(), codes);
      distance += q2_[word2].ComputeDistance(c2->second.data(), codes + half_ncomp_);
      InsertNeighbor(f.indices[j], distance, k, &nn);
    }
  }
  for (size_t i = 0; i < nn.size(); ++i) {
    indices[i] = nn[i].second;
    distances[i] = nn[i].first;
  }
  for (int i = static_cast<int>(nn.size()); i < k; ++i) {
    indices[i] = -1;
    distances[i] = std::numeric_limits<float>::infinity();
  }
}

}  // namespace lc_oracle
