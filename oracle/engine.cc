// ORACLE (test infrastructure). LoopDetector (A4, A11-A15) and scoring (A12).
// Reference: algorithms/loopclosure/matching-based-loopclosure/
//   src/matching-based-engine.cc:48-168 (Find), :170-215 (getMatchForDescriptorIndex),
//   :217-253 (Insert), :319-338 (getNumNeighborsToSearch)
//   include/matching-based-loopclosure/matching-based-engine-inl.h:45-183
//   (doCovisibilityFiltering), :219-254 (computeRelevantIdsForFiltering)
//   include/matching-based-loopclosure/scoring.h:38-59, :92-187
//
// Canonical orders replacing the reference's std::unordered_* iteration /
// nth_element tie behaviour:
//  (2) top-fraction ties : score descending, then keyframe insertion number ascending
//  (3) component ties    : larger size, then the component with the smallest member id
//  (4) make_matches_unique: the duplicate with the smallest db descriptor index survives
//  (5) output order      : (query frame index, query keypoint, db descriptor index)
#include <algorithm>
#include <cassert>
#include <functional>
#include <tuple>
#include <cmath>
#include <map>
#include <numeric>
#include <set>

#include "lc_oracle.h"

namespace lc_oracle {

// boost::math::pdf(binomial(n, p), k) = C(n,k) p^k (1-p)^(n-k)
// (boost/math/distributions/binomial.hpp: special cases then ibeta_derivative).
double BinomialPdf(double n, double p, double k) {
  if (p == 0) return (k == 0) ? 1.0 : 0.0;
  if (p == 1) return (k == n) ? 1.0 : 0.0;
  if (n == 0) return 1.0;
  if (k == 0) return std::pow(1 - p, n);
  if (k == n) return std::pow(p, k);
  const double lg = std::lgamma(n + 1) - std::lgamma(k + 1) - std::lgamma(n - k + 1) +
                    k * std::log(p) + (n - k) * std::log1p(-p);
  return std::exp(lg);
}

void ComputeAccumulationScore(const std::vector<size_t>& num_matches, std::vector<float>* scores) {
  scores->clear();
  for (size_t m : num_matches) scores->push_back(static_cast<float>(m));
}

// scoring.h:92-187, transcribed literally including the in-loop +inf patch (quirk 7).
void ComputeProbabilisticScore(const std::vector<size_t>& num_matches,
                               const std::vector<size_t>& num_descriptors_per_id,
                               size_t num_descriptors_in_database, std::vector<float>* scores) {
  scores->clear();
  if (num_descriptors_in_database == 0u || num_matches.empty()) return;
  size_t total = 0;
  for (size_t m : num_matches) total += m;
  size_t index_inf = 0, num_matches_inf = 0;
  for (size_t i = 0; i < num_matches.size(); ++i) {
    const size_t m = num_matches[i];
    float score = 0.f;
    const double p = static_cast<double>(num_descriptors_per_id[i]) /
                     static_cast<double>(num_descriptors_in_database);
    const size_t lower_median = static_cast<size_t>(static_cast<double>(total) * p);
    if (m > lower_median) {
      const double prob = BinomialPdf(static_cast<double>(total), p, static_cast<double>(m));
      if (prob == 0.0) {
        if (m > num_matches_inf) {
          num_matches_inf = m;
          index_inf = scores->size();
        }
        score = std::numeric_limits<float>::max();
      } else {
        score = static_cast<float>(-std::log10(prob));
      }
    }
    if (num_matches_inf > 0u && index_inf < scores->size()) {
      (*scores)[index_inf] = std::numeric_limits<float>::infinity();
    }
    scores->push_back(score);
  }
}

LoopDetector::LoopDetector(const EngineSettings& s, const Vocabulary& v) : s_(s), v_(v) {
  fp_ = QuantizeProjection(v.projection, v.target_dim);
  if (s.engine == 0) {
    imi_ = new InvertedMultiIndex(v.words1, v.words2, s.num_closest_words_for_nn_search, s.search);
  } else {
    assert(v.has_pq);
    imipq_ = new InvertedMultiPQIndex(v.words1, v.words2, v.pq_centers1, v.pq_centers2,
                                      v.pq_num_components, v.pq_dim_per_comp, v.pq_num_centers,
                                      s.num_closest_words_for_nn_search, s.search);
  }
}

void LoopDetector::ProjectDescriptors(const uint8_t* raw, int bytes_per_desc, int n,
                                      float* out) const {
  ProjectDescriptorBlock(raw, bytes_per_desc, n, fp_, out);
}

int LoopDetector::NumDescriptors() const {
  return imi_ ? imi_->GetNumDescriptorsInIndex() : imipq_->GetNumDescriptorsInIndex();
}

void LoopDetector::Clear() {
  keyframes_.clear();
  keyframe_ids_.clear();
  desc_to_keyframe_.clear();
  if (imi_) imi_->Clear();
  if (imipq_) imipq_->Clear();
}

// matching-based-engine.cc:217-253
bool LoopDetector::Insert(const ProjectedImage& image) {
  const int n = image.dim ? static_cast<int>(image.projected_descriptors.size() / image.dim) : 0;
  assert(static_cast<size_t>(n) == image.landmarks.size());
  // CHECK(emplace(...).second): the keyframe id must be new (matching-based-engine.cc:244-252)
  if (!keyframe_ids_.insert({image.vertex_id, image.frame_index}).second) return false;
  Keyframe kf;
  kf.ts = image.timestamp_ns;
  kf.vertex = image.vertex_id;
  kf.mission = image.mission_id;
  kf.frame_index = image.frame_index;
  kf.first_descriptor = static_cast<int>(desc_to_keyframe_.size());
  kf.num_descriptors = n;
  kf.landmarks = image.landmarks;
  const int kf_number = static_cast<int>(keyframes_.size());
  for (int i = 0; i < n; ++i) desc_to_keyframe_.push_back(kf_number);
  if (imi_) imi_->AddDescriptors(image.projected_descriptors.data(), n);
  if (imipq_) imipq_->AddDescriptors(image.projected_descriptors.data(), n);
  keyframes_.push_back(std::move(kf));
  return true;
}

// matching-based-engine.cc:319-338
int LoopDetector::NumNeighborsToSearch() const {
  int k = s_.num_nearest_neighbors;
  if (k == -1) {
    const int n = NumDescriptors();
    if (n < 1e4)
      k = 1;
    else if (n < 1e5)
      k = 2;
    else if (n < 1e6)
      k = 3;
    else if (n < 1e7)
      k = 6;
    else
      k = 8;
  }
  return k;
}

void LoopDetector::KnnBatch(const float* q, int n, int k, int* idx, float* dist) const {
  const int dim = v_.target_dim;
  for (int i = 0; i < n; ++i) {
    if (imi_)
      imi_->GetNNearestNeighbors(q + static_cast<size_t>(i) * dim, k, idx + static_cast<size_t>(i) * k,
                                 dist + static_cast<size_t>(i) * k);
    else
      imipq_->GetNNearestNeighbors(q + static_cast<size_t>(i) * dim, k,
                                   idx + static_cast<size_t>(i) * k,
                                   dist + static_cast<size_t>(i) * k);
  }
}

namespace {
struct MatchKey {  // identity of a vi_map::FrameKeyPointToStructureMatch (operator==)
  int qf, qk, kf;
  int64_t lm;
  bool operator<(const MatchKey& o) const {
    return std::tie(qf, qk, kf, lm) < std::tie(o.qf, o.qk, o.kf, o.lm);
  }
};
MatchKey KeyOf(const Match& m) {
  return MatchKey{m.query_frame_index, m.query_keypoint, m.db_keyframe, m.landmark};
}
bool CanonicalLess(const Match& a, const Match& b) {
  return std::tie(a.query_frame_index, a.query_keypoint, a.db_descriptor) <
         std::tie(b.query_frame_index, b.query_keypoint, b.db_descriptor);
}
}  // namespace

// doCovisibilityFiltering (matching-based-engine-inl.h:45-183) in graph terms:
// connected components of the bipartite graph {relevant ids} x {landmarks};
// component size = number of DISTINCT matches of its member ids; the largest
// component is kept iff size > min_verify_matches_num.
void CovisComponents(const std::vector<Match>& matches, bool by_vertex,
                     const std::vector<int64_t>* relevant_ids, size_t min_verify_matches_num,
                     bool make_unique, std::vector<Match>* out) {
  if (matches.empty()) return;
  auto group_of = [&](const Match& m) -> int64_t {
    return by_vertex ? m.db_vertex : static_cast<int64_t>(m.db_keyframe);
  };
  std::set<int64_t> relevant;
  if (relevant_ids) relevant.insert(relevant_ids->begin(), relevant_ids->end());
  // Union-find over group ids (relevant only) through shared landmarks.
  std::map<int64_t, int64_t> parent;  // group -> parent group
  std::function<int64_t(int64_t)> find = [&](int64_t x) {
    int64_t r = x;
    while (parent[r] != r) r = parent[r];
    while (parent[x] != r) {
      const int64_t nx = parent[x];
      parent[x] = r;
      x = nx;
    }
    return r;
  };
  std::map<int64_t, int64_t> lm_first_group;
  for (const Match& m : matches) {
    const int64_t g = group_of(m);
    if (relevant_ids && !relevant.count(g)) continue;
    if (!parent.count(g)) parent[g] = g;
    auto it = lm_first_group.find(m.landmark);
    if (it == lm_first_group.end()) {
      lm_first_group[m.landmark] = g;
    } else {
      const int64_t a = find(g), b = find(it->second);
      if (a != b) {
        // Root = smaller id so that the label is the component's minimal member.
        if (a < b)
          parent[b] = a;
        else
          parent[a] = b;
      }
    }
  }
  // Distinct matches per component; representative = smallest db descriptor.
  std::map<int64_t, std::map<MatchKey, Match>> comp;
  for (const Match& m : matches) {
    const int64_t g = group_of(m);
    if (relevant_ids && !relevant.count(g)) continue;
    const int64_t root = find(g);
    auto& cm = comp[root];
    auto ins = cm.emplace(KeyOf(m), m);
    if (!ins.second && m.db_descriptor < ins.first->second.db_descriptor) ins.first->second = m;
  }
  size_t best_size = 0;
  int64_t best_root = -1;
  for (const auto& c : comp) {  // ascending root id: ties keep the smallest root
    if (c.second.size() > best_size) {
      best_size = c.second.size();
      best_root = c.first;
    }
  }
  if (!(best_size > min_verify_matches_num)) return;
  std::vector<Match> sel;
  for (const auto& km : comp[best_root]) sel.push_back(km.second);
  std::sort(sel.begin(), sel.end(), CanonicalLess);
  if (make_unique) {
    // (query keypoint, landmark) must be unique; sel is sorted by db descriptor
    // within a keypoint so the first occurrence is the smallest descriptor index.
    std::set<std::tuple<int, int, int64_t>> used;
    for (const Match& m : sel) {
      if (used.emplace(m.query_frame_index, m.query_keypoint, m.landmark).second)
        out->push_back(m);
    }
  } else {
    out->insert(out->end(), sel.begin(), sel.end());
  }
}

void LoopDetector::CovisFilterKeyframes(const std::vector<Match>& in, bool make_unique,
                                        std::vector<Match>* out, FrameTrace* trace) const {
  if (in.empty()) return;
  // Votes per candidate keyframe (vector sizes incl. duplicates: scoring.h:55).
  std::map<int, size_t> votes;
  for (const Match& m : in) ++votes[m.db_keyframe];
  std::vector<int> ids;
  std::vector<size_t> num_matches, num_desc;
  for (const auto& v : votes) {
    ids.push_back(v.first);
    num_matches.push_back(v.second);
    num_desc.push_back(static_cast<size_t>(keyframes_[v.first].num_descriptors));
  }
  std::vector<float> scores;
  if (s_.scoring == 0)
    ComputeAccumulationScore(num_matches, &scores);
  else
    ComputeProbabilisticScore(num_matches, num_desc, static_cast<size_t>(NumDescriptors()),
                              &scores);
  // computeRelevantIdsForFiltering (inl.h:219-254)
  constexpr size_t kNumMinimumScoreIdsToEvaluate = 4u;
  size_t n_eval = std::max<size_t>(
      static_cast<size_t>(static_cast<float>(scores.size()) * s_.fraction_best_scores),
      kNumMinimumScoreIdsToEvaluate);
  n_eval = std::min<size_t>(n_eval, scores.size());
  std::vector<int> order(ids.size());
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
    if (scores[a] != scores[b]) return scores[a] > scores[b];
    return ids[a] < ids[b];
  });
  std::vector<int64_t> relevant;
  for (size_t i = 0; i < n_eval; ++i) relevant.push_back(ids[order[i]]);
  std::sort(relevant.begin(), relevant.end());
  if (trace) {
    trace->cand_keyframes = ids;
    trace->cand_votes.assign(num_matches.begin(), num_matches.end());
    trace->selected_keyframes.assign(relevant.begin(), relevant.end());
  }
  CovisComponents(in, /*by_vertex=*/false, &relevant, s_.min_verify_matches_num, make_unique, out);
}

void LoopDetector::FindFrame(const ProjectedImage& q, bool make_unique, FrameTrace* trace) const {
  const int dim = v_.target_dim;
  const int n = static_cast<int>(q.projected_descriptors.size() / dim);
  const int k = NumNeighborsToSearch();
  trace->knn_indices.assign(static_cast<size_t>(k) * n, -1);
  trace->knn_distances.assign(static_cast<size_t>(k) * n, 0.f);
  KnnBatch(q.projected_descriptors.data(), n, k, trace->knn_indices.data(),
           trace->knn_distances.data());
  trace->raw_matches.clear();
  const double min_dt = s_.min_image_time_seconds * 1e9;  // kSecondsToNanoSeconds
  for (int kp = 0; kp < n; ++kp) {
    for (int j = 0; j < k; ++j) {
      const int idx = trace->knn_indices[static_cast<size_t>(kp) * k + j];
      const float d = trace->knn_distances[static_cast<size_t>(kp) * k + j];
      if (idx == -1 || d == std::numeric_limits<float>::infinity()) break;
      // getMatchForDescriptorIndex (matching-based-engine.cc:170-215)
      const int kfn = desc_to_keyframe_[idx];
      const Keyframe& kf = keyframes_[kfn];
      if (static_cast<double>(std::llabs(q.timestamp_ns - kf.ts)) < min_dt &&
          q.mission_id == kf.mission)
        continue;
      Match m;
      m.query_frame_index = q.frame_index;
      m.query_keypoint = kp;
      m.db_descriptor = idx;
      m.db_keyframe = kfn;
      m.db_vertex = kf.vertex;
      m.landmark = kf.landmarks.empty() ? -1 : kf.landmarks[idx - kf.first_descriptor];
      trace->raw_matches.push_back(m);
    }
  }
  trace->filtered.clear();
  CovisFilterKeyframes(trace->raw_matches, make_unique, &trace->filtered, trace);
}

void LoopDetector::Find(const std::vector<const ProjectedImage*>& images,
                        std::vector<Match>* out) const {
  out->clear();
  if (images.empty()) return;
  for (const ProjectedImage* im : images) {
    assert(im->vertex_id == images[0]->vertex_id);
    (void)im;
  }
  const bool use_vertex_covis_filter = images.size() > 1u;
  std::vector<Match> temporary;
  for (const ProjectedImage* im : images) {
    FrameTrace tr;
    FindFrame(*im, !use_vertex_covis_filter, &tr);
    temporary.insert(temporary.end(), tr.filtered.begin(), tr.filtered.end());
  }
  if (use_vertex_covis_filter) {
    // matching-based-engine.cc:147-165: regroup by result vertex, no scoring.
    CovisComponents(temporary, /*by_vertex=*/true, nullptr, s_.min_verify_matches_num,
                    /*make_unique=*/true, out);
  } else {
    out->swap(temporary);
  }
  std::sort(out->begin(), out->end(), CanonicalLess);
}

}  // namespace lc_oracle
