// ORACLE (test infrastructure): flat C entry points for ctypes (tests/, smoke(),
// bench.py cpu_baseline / --impl reference only).
#include <algorithm>
#include <chrono>
#include <limits>
#include <cstring>
#include <thread>

#include "lc_oracle.h"

using namespace lc_oracle;

namespace {
Matrix MakeMatrix(const float* data, int rows, int cols) {
  Matrix m;
  m.rows = rows;
  m.cols = cols;
  m.data.assign(data, data + static_cast<size_t>(rows) * cols);
  return m;
}
}  // namespace

extern "C" {

// ---- A10 ----
int lco_insert_neighbors(const int* idx, const float* dist, int n, int k, int* out_idx,
                         float* out_dist) {
  std::vector<std::pair<float, int>> nn;
  for (int i = 0; i < n; ++i) InsertNeighbor(idx[i], dist[i], k, &nn);
  for (size_t i = 0; i < nn.size(); ++i) {
    out_idx[i] = nn[i].second;
    out_dist[i] = nn[i].first;
  }
  return static_cast<int>(nn.size());
}

// ---- A7 ----
int lco_multi_sequence(const int* i1, const float* d1, int n1, const int* i2, const float* d2,
                       int n2, int num_words, int* out_pairs) {
  std::vector<std::pair<int, int>> cw;
  MultiSequenceAlgorithm(i1, d1, n1, i2, d2, n2, num_words, &cw);
  for (size_t i = 0; i < cw.size(); ++i) {
    out_pairs[2 * i] = cw[i].first;
    out_pairs[2 * i + 1] = cw[i].second;
  }
  return static_cast<int>(cw.size());
}

// ---- A6 ----
void* lco_kdtree_create(const float* cloud, int dim, int n) {
  KdTree* t = new KdTree();
  t->Build(cloud, dim, n);
  return t;
}
void lco_kdtree_destroy(void* t) { delete static_cast<KdTree*>(t); }
unsigned long lco_kdtree_knn(void* t, const float* q, int k, float eps, float radius, int* idx,
                             float* d2) {
  return static_cast<KdTree*>(t)->Knn(q, k, eps, radius, idx, d2);
}
int lco_kdtree_num_nodes(void* t) { return static_cast<int>(static_cast<KdTree*>(t)->nodes.size()); }
// nodes: 3 x uint32 per node (dim, child_or_size, cut_val bits / bucket index); buckets: point ids
void lco_kdtree_export(void* tv, uint32_t* nodes, int* buckets) {
  KdTree* t = static_cast<KdTree*>(tv);
  for (size_t i = 0; i < t->nodes.size(); ++i) {
    nodes[3 * i] = t->nodes[i].dim;
    nodes[3 * i + 1] = t->nodes[i].child_or_size;
    nodes[3 * i + 2] = t->nodes[i].bucket_index;
  }
  for (size_t i = 0; i < t->bucket_point_index.size(); ++i) buckets[i] = t->bucket_point_index[i];
}

int lco_find_closest_words(const float* w1, int n1, const float* w2, int n2, int sub_dim,
                           const float* query, int num_closest, float eps, float radius,
                           int* out_pairs) {
  KdTree t1, t2;
  t1.Build(w1, sub_dim, n1);
  t2.Build(w2, sub_dim, n2);
  SearchParams sp;
  sp.knn_epsilon = eps;
  sp.knn_max_radius = radius;
  std::vector<std::pair<int, int>> cw;
  FindClosestWords(query, sub_dim, num_closest, t1, t2, sp, &cw);
  for (size_t i = 0; i < cw.size(); ++i) {
    out_pairs[2 * i] = cw[i].first;
    out_pairs[2 * i + 1] = cw[i].second;
  }
  return static_cast<int>(cw.size());
}

float lco_squared_distance(const float* a, const float* b, int dim) {
  return SquaredDistance(a, b, dim);
}

// ---- A5/A8 ----
void* lco_imi_create(const float* w1, int n1, const float* w2, int n2, int sub_dim, int nw,
                     float eps, float radius) {
  SearchParams sp;
  sp.knn_epsilon = eps;
  sp.knn_max_radius = radius;
  return new InvertedMultiIndex(MakeMatrix(w1, sub_dim, n1), MakeMatrix(w2, sub_dim, n2), nw, sp);
}
void lco_imi_destroy(void* h) { delete static_cast<InvertedMultiIndex*>(h); }
void lco_imi_add(void* h, const float* desc, int n) {
  static_cast<InvertedMultiIndex*>(h)->AddDescriptors(desc, n);
}
void lco_imi_clear(void* h) { static_cast<InvertedMultiIndex*>(h)->Clear(); }
int lco_imi_num_descriptors(void* h) {
  return static_cast<InvertedMultiIndex*>(h)->GetNumDescriptorsInIndex();
}
void lco_imi_knn(void* h, const float* q, int n, int k, int* idx, float* dist) {
  InvertedMultiIndex* imi = static_cast<InvertedMultiIndex*>(h);
  const int dim = imi->dim();
  for (int i = 0; i < n; ++i)
    imi->GetNNearestNeighbors(q + static_cast<size_t>(i) * dim, k, idx + static_cast<size_t>(i) * k,
                              dist + static_cast<size_t>(i) * k);
}
int lco_imi_cell_of(void* h, const float* desc) {
  return static_cast<InvertedMultiIndex*>(h)->CellOfDescriptor(desc);
}
// cells: nw ints per query (padded with -2 when fewer pairs come back)
void lco_imi_visited_cells(void* h, const float* q, int n, int nw, int* cells) {
  InvertedMultiIndex* imi = static_cast<InvertedMultiIndex*>(h);
  const int dim = imi->dim();
  std::vector<int> c;
  for (int i = 0; i < n; ++i) {
    imi->VisitedCells(q + static_cast<size_t>(i) * dim, &c);
    for (int j = 0; j < nw; ++j)
      cells[static_cast<size_t>(i) * nw + j] = j < static_cast<int>(c.size()) ? c[j] : -2;
  }
}
int lco_imi_num_files(void* h) {
  return static_cast<int>(static_cast<InvertedMultiIndex*>(h)->inverted_files().size());
}
// word_index -> bucket id (first-touch order); returns -1 if absent
int lco_imi_bucket_of_word(void* h, int word_index) {
  const auto& m = static_cast<InvertedMultiIndex*>(h)->word_index_map();
  auto it = m.find(word_index);
  return it == m.end() ? -1 : it->second;
}
int lco_imi_file_size(void* h, int bucket) {
  return static_cast<int>(static_cast<InvertedMultiIndex*>(h)->inverted_files()[bucket].indices.size());
}
void lco_imi_file_get(void* h, int bucket, int* indices, float* descriptors) {
  const auto& f = static_cast<InvertedMultiIndex*>(h)->inverted_files()[bucket];
  std::memcpy(indices, f.indices.data(), f.indices.size() * sizeof(int));
  std::memcpy(descriptors, f.descriptors.data(), f.descriptors.size() * sizeof(float));
}

// ---- A9 ----
void lco_pq_quantize(const float* centers, int ncomp, int dim_per_comp, int ncenters,
                     const float* vecs, int n, int* codes) {
  ProductQuantizer pq;
  pq.num_components = ncomp;
  pq.dim_per_comp = dim_per_comp;
  pq.num_centers = ncenters;
  pq.centers.assign(centers, centers + static_cast<size_t>(ncomp) * ncenters * dim_per_comp);
  for (int i = 0; i < n; ++i)
    pq.Quantize(vecs + static_cast<size_t>(i) * ncomp * dim_per_comp, codes + static_cast<size_t>(i) * ncomp);
}
void lco_pq_fill_lut(const float* centers, int ncomp, int dim_per_comp, int ncenters,
                     const float* vec, float* lut) {
  ProductQuantizer pq;
  pq.num_components = ncomp;
  pq.dim_per_comp = dim_per_comp;
  pq.num_centers = ncenters;
  pq.centers.assign(centers, centers + static_cast<size_t>(ncomp) * ncenters * dim_per_comp);
  pq.FillLUT(vec, lut);
}
void lco_pq_distances(int ncomp, int ncenters, const float* lut, const int* codes, int n,
                      float* dist, int add) {
  ProductQuantizer pq;
  pq.num_components = ncomp;
  pq.num_centers = ncenters;
  pq.dim_per_comp = 1;
  for (int i = 0; i < n; ++i) {
    if (add) {  // ComputeAndAddDistances (product-quantization.h:159-172)
      for (int j = 0; j < ncomp; ++j) dist[i] += lut[j * ncenters + codes[static_cast<size_t>(i) * ncomp + j]];
    } else {
      dist[i] = pq.ComputeDistance(lut, codes + static_cast<size_t>(i) * ncomp);
    }
  }
}
void* lco_imipq_create(const float* w1, int n1, const float* w2, int n2, int sub_dim,
                       const float* qc1, const float* qc2, int ncomp, int dim_per_comp,
                       int ncenters, int nw, float eps, float radius) {
  SearchParams sp;
  sp.knn_epsilon = eps;
  sp.knn_max_radius = radius;
  const int cols_per_pq = (ncomp / 2) * ncenters;
  return new InvertedMultiPQIndex(MakeMatrix(w1, sub_dim, n1), MakeMatrix(w2, sub_dim, n2),
                                  MakeMatrix(qc1, dim_per_comp, cols_per_pq * n1),
                                  MakeMatrix(qc2, dim_per_comp, cols_per_pq * n2), ncomp,
                                  dim_per_comp, ncenters, nw, sp);
}
void lco_imipq_destroy(void* h) { delete static_cast<InvertedMultiPQIndex*>(h); }
void lco_imipq_add(void* h, const float* desc, int n) {
  static_cast<InvertedMultiPQIndex*>(h)->AddDescriptors(desc, n);
}
void lco_imipq_knn(void* h, const float* q, int dim, int n, int k, int* idx, float* dist) {
  for (int i = 0; i < n; ++i)
    static_cast<InvertedMultiPQIndex*>(h)->GetNNearestNeighbors(
        q + static_cast<size_t>(i) * dim, k, idx + static_cast<size_t>(i) * k, dist + static_cast<size_t>(i) * k);
}
int lco_imipq_bucket_of_word(void* h, int word_index) {
  const auto& m = static_cast<InvertedMultiPQIndex*>(h)->word_index_map();
  auto it = m.find(word_index);
  return it == m.end() ? -1 : it->second;
}
int lco_imipq_num_files(void* h) {
  return static_cast<int>(static_cast<InvertedMultiPQIndex*>(h)->inverted_files().size());
}
int lco_imipq_file_size(void* h, int bucket) {
  return static_cast<int>(static_cast<InvertedMultiPQIndex*>(h)->inverted_files()[bucket].indices.size());
}
void lco_imipq_file_get(void* h, int bucket, int* indices, int* codes) {
  const auto& f = static_cast<InvertedMultiPQIndex*>(h)->inverted_files()[bucket];
  std::memcpy(indices, f.indices.data(), f.indices.size() * sizeof(int));
  std::memcpy(codes, f.codes.data(), f.codes.size() * sizeof(int));
}

// ---- A12 ----
void lco_score(int probabilistic, const uint64_t* num_matches, const uint64_t* num_desc, int n,
               uint64_t num_db, float* scores, int* n_out) {
  std::vector<size_t> m(num_matches, num_matches + n), d(num_desc, num_desc + n);
  std::vector<float> s;
  if (probabilistic)
    ComputeProbabilisticScore(m, d, num_db, &s);
  else
    ComputeAccumulationScore(m, &s);
  *n_out = static_cast<int>(s.size());
  for (size_t i = 0; i < s.size(); ++i) scores[i] = s[i];
}

// ---- A1/A2 ----
void lco_project(const float* P, int rows, int cols, int target_dim, const uint8_t* raw,
                 int bytes_per_desc, int n, float* out, int float_mode) {
  Matrix m = MakeMatrix(P, rows, cols);
  if (float_mode) {
    ProjectDescriptorBlockFloat(raw, bytes_per_desc, n, m, target_dim, out);
  } else {
    FixedPointProjection fp = QuantizeProjection(m, target_dim);
    ProjectDescriptorBlock(raw, bytes_per_desc, n, fp, out);
  }
}
void lco_quantize_projection(const float* P, int rows, int cols, int target_dim, int32_t* p_int,
                             int* shift) {
  FixedPointProjection fp = QuantizeProjection(MakeMatrix(P, rows, cols), target_dim);
  std::memcpy(p_int, fp.p_int.data(), fp.p_int.size() * sizeof(int32_t));
  std::memcpy(shift, fp.shift.data(), fp.shift.size() * sizeof(int));
}

// ---- T2 ----
// Returns 0 on success. dims: version, target_dim, P rows, P cols, sub_dim, W1, W2, has_pq,
// ncomp, ncenters, dim_per_comp
int lco_vocab_parse(const uint8_t* blob, uint64_t size, int want_pq, int* dims) {
  Vocabulary v;
  std::string err;
  if (!ParseVocabulary(blob, size, want_pq != 0, &v, &err)) return 1;
  dims[0] = v.version;
  dims[1] = v.target_dim;
  dims[2] = v.projection.rows;
  dims[3] = v.projection.cols;
  dims[4] = v.words1.rows;
  dims[5] = v.words1.cols;
  dims[6] = v.words2.cols;
  dims[7] = v.has_pq;
  dims[8] = v.pq_num_components;
  dims[9] = v.pq_num_centers;
  dims[10] = v.pq_dim_per_comp;
  return 0;
}

// ---- engine ----
struct lco_settings {
  int num_closest_words, num_nearest_neighbors, scoring, engine;
  double min_image_time_seconds;
  uint64_t min_verify_matches_num;
  float fraction_best_scores, knn_epsilon, knn_max_radius;
};
void* lco_engine_create(const lco_settings* s, const uint8_t* blob, uint64_t size) {
  Vocabulary v;
  std::string err;
  if (!ParseVocabulary(blob, size, s->engine == 1, &v, &err)) return nullptr;
  EngineSettings es;
  es.num_closest_words_for_nn_search = s->num_closest_words;
  es.num_nearest_neighbors = s->num_nearest_neighbors;
  es.scoring = s->scoring;
  es.engine = s->engine;
  es.min_image_time_seconds = s->min_image_time_seconds;
  es.min_verify_matches_num = s->min_verify_matches_num;
  es.fraction_best_scores = s->fraction_best_scores;
  es.search.knn_epsilon = s->knn_epsilon;
  es.search.knn_max_radius = s->knn_max_radius;
  return new LoopDetector(es, v);
}
void lco_engine_destroy(void* h) { delete static_cast<LoopDetector*>(h); }
void lco_engine_clear(void* h) { static_cast<LoopDetector*>(h)->Clear(); }
void lco_engine_project(void* h, const uint8_t* raw, int bytes_per_desc, int n, float* out) {
  static_cast<LoopDetector*>(h)->ProjectDescriptors(raw, bytes_per_desc, n, out);
}
int lco_engine_insert(void* h, int64_t ts, int64_t vertex, int frame_index, int64_t mission,
                      int dim, const float* proj, int n, const int64_t* landmarks) {
  ProjectedImage im;
  im.timestamp_ns = ts;
  im.vertex_id = vertex;
  im.frame_index = frame_index;
  im.mission_id = mission;
  im.dim = dim;
  im.projected_descriptors.assign(proj, proj + static_cast<size_t>(dim) * n);
  im.landmarks.assign(landmarks, landmarks + n);
  return static_cast<LoopDetector*>(h)->Insert(im) ? 0 : 1;
}
int lco_engine_num_descriptors(void* h) { return static_cast<LoopDetector*>(h)->NumDescriptors(); }
int lco_engine_num_entries(void* h) { return static_cast<int>(static_cast<LoopDetector*>(h)->NumEntries()); }
int lco_engine_num_neighbors(void* h) { return static_cast<LoopDetector*>(h)->NumNeighborsToSearch(); }
void lco_engine_knn(void* h, const float* q, int n, int k, int* idx, float* dist) {
  static_cast<LoopDetector*>(h)->KnnBatch(q, n, k, idx, dist);
}

// One query vertex with n_frames frames. Per frame f: ts[f], frame_index[f], n_desc[f];
// descriptors concatenated. Output: matches as 6 x int64 rows
// (query_frame_index, query_keypoint, db_descriptor, db_keyframe, db_vertex, landmark).
int lco_engine_find(void* h, int n_frames, const int64_t* ts, int64_t vertex, const int* frame_index,
                    int64_t mission, int dim, const int* n_desc, const float* proj,
                    int64_t* out_matches, int max_matches) {
  std::vector<ProjectedImage> ims(n_frames);
  std::vector<const ProjectedImage*> ptrs;
  size_t off = 0;
  for (int f = 0; f < n_frames; ++f) {
    ims[f].timestamp_ns = ts[f];
    ims[f].vertex_id = vertex;
    ims[f].frame_index = frame_index[f];
    ims[f].mission_id = mission;
    ims[f].dim = dim;
    ims[f].projected_descriptors.assign(proj + off * dim, proj + (off + n_desc[f]) * dim);
    off += n_desc[f];
    ptrs.push_back(&ims[f]);
  }
  std::vector<Match> out;
  static_cast<LoopDetector*>(h)->Find(ptrs, &out);
  const int n = std::min<int>(static_cast<int>(out.size()), max_matches);
  for (int i = 0; i < n; ++i) {
    int64_t* r = out_matches + 6 * static_cast<size_t>(i);
    r[0] = out[i].query_frame_index;
    r[1] = out[i].query_keypoint;
    r[2] = out[i].db_descriptor;
    r[3] = out[i].db_keyframe;
    r[4] = out[i].db_vertex;
    r[5] = out[i].landmark;
  }
  return static_cast<int>(out.size());
}

// Stage trace of one single-camera query frame (parity checkpoints P4-P7).
// counts: [k, n_raw, n_cand, n_selected, n_filtered]
void lco_engine_find_frame_trace(void* h, int64_t ts, int64_t vertex, int frame_index,
                                 int64_t mission, int dim, int n, const float* proj,
                                 int make_unique, int* counts, int* knn_idx, float* knn_dist,
                                 int64_t* raw_matches, int* cand_kf, int* cand_votes,
                                 int* selected_kf, int64_t* filtered) {
  ProjectedImage im;
  im.timestamp_ns = ts;
  im.vertex_id = vertex;
  im.frame_index = frame_index;
  im.mission_id = mission;
  im.dim = dim;
  im.projected_descriptors.assign(proj, proj + static_cast<size_t>(dim) * n);
  LoopDetector::FrameTrace tr;
  LoopDetector* ld = static_cast<LoopDetector*>(h);
  ld->FindFrame(im, make_unique != 0, &tr);
  const int k = ld->NumNeighborsToSearch();
  counts[0] = k;
  counts[1] = static_cast<int>(tr.raw_matches.size());
  counts[2] = static_cast<int>(tr.cand_keyframes.size());
  counts[3] = static_cast<int>(tr.selected_keyframes.size());
  counts[4] = static_cast<int>(tr.filtered.size());
  std::memcpy(knn_idx, tr.knn_indices.data(), tr.knn_indices.size() * sizeof(int));
  std::memcpy(knn_dist, tr.knn_distances.data(), tr.knn_distances.size() * sizeof(float));
  auto dump = [](const std::vector<Match>& v, int64_t* o) {
    for (size_t i = 0; i < v.size(); ++i) {
      o[6 * i + 0] = v[i].query_frame_index;
      o[6 * i + 1] = v[i].query_keypoint;
      o[6 * i + 2] = v[i].db_descriptor;
      o[6 * i + 3] = v[i].db_keyframe;
      o[6 * i + 4] = v[i].db_vertex;
      o[6 * i + 5] = v[i].landmark;
    }
  };
  dump(tr.raw_matches, raw_matches);
  dump(tr.filtered, filtered);
  std::memcpy(cand_kf, tr.cand_keyframes.data(), tr.cand_keyframes.size() * sizeof(int));
  std::memcpy(cand_votes, tr.cand_votes.data(), tr.cand_votes.size() * sizeof(int));
  std::memcpy(selected_kf, tr.selected_keyframes.data(), tr.selected_keyframes.size() * sizeof(int));
}

// ---- A17-A22 ----
int lco_gp3p_solve(const double* f, const double* v, const double* p, double* solutions) {
  double sols[8][12];
  const int n = Gp3pSolve(f, v, p, sols);
  std::memcpy(solutions, sols, sizeof(double) * 12 * n);
  return n;
}
void lco_rng_stream(uint32_t seed, int mapping, int n, int* out) {
  RansacRng rng(seed, mapping);
  for (int i = 0; i < n; ++i) out[i] = rng.Next();
}
struct lco_camera {
  double fu, fv, cu, cv;
  int distortion;
  double dist[4];
  double R_B_C[9];
  double t_B_C[3];
};
static std::vector<Camera> ToCams(const lco_camera* cams, int n) {
  std::vector<Camera> out(n);
  for (int i = 0; i < n; ++i) {
    out[i].fu = cams[i].fu;
    out[i].fv = cams[i].fv;
    out[i].cu = cams[i].cu;
    out[i].cv = cams[i].cv;
    out[i].distortion = cams[i].distortion;
    std::memcpy(out[i].dist, cams[i].dist, sizeof(double) * 4);
    std::memcpy(out[i].R_B_C, cams[i].R_B_C, sizeof(double) * 9);
    std::memcpy(out[i].t_B_C, cams[i].t_B_C, sizeof(double) * 3);
  }
  return out;
}
void lco_back_project(const lco_camera* cam, const double* kp, int n, double* bearings) {
  std::vector<Camera> c = ToCams(cam, 1);
  for (int i = 0; i < n; ++i) BackProject3(c[0], kp + 2 * i, bearings + 3 * i);
}
double lco_ransac_threshold(const lco_camera* cams, int n_cams, double pixel_sigma) {
  return RansacThreshold(ToCams(cams, n_cams), pixel_sigma);
}
// out_scalars: [accepted, num_inliers, ransac_success, iterations, n_ransac_inliers, model idx x4]
void lco_handle_loop_closure(int n, const double* keypoints, const int* frame_index,
                             const int* keypoint_index, const double* landmarks,
                             const lco_camera* cams, int n_cams, int min_inlier_count,
                             double min_inlier_ratio, double pixel_sigma, int num_iters,
                             uint32_t seed, int rng_mapping, int* out_scalars, double* out_ratio,
                             double* out_T, int* out_inliers, double* out_inlier_dist,
                             int* out_best_per_keypoint) {
  VerifyInput in;
  in.n = n;
  in.keypoints.assign(keypoints, keypoints + 2 * static_cast<size_t>(n));
  in.frame_index.assign(frame_index, frame_index + n);
  in.keypoint_index.assign(keypoint_index, keypoint_index + n);
  in.landmarks.assign(landmarks, landmarks + 3 * static_cast<size_t>(n));
  HandlerSettings hs;
  hs.min_inlier_count = min_inlier_count;
  hs.min_inlier_ratio = min_inlier_ratio;
  hs.ransac_pixel_sigma = pixel_sigma;
  hs.num_ransac_iters = num_iters;
  hs.seed = seed;
  hs.rng_mapping = rng_mapping;
  VerifyResult r;
  HandleLoopClosure(in, ToCams(cams, n_cams), hs, &r);
  out_scalars[0] = r.accepted;
  out_scalars[1] = r.num_inliers;
  out_scalars[2] = r.ransac.success;
  out_scalars[3] = r.ransac.iterations;
  out_scalars[4] = static_cast<int>(r.ransac.inliers.size());
  for (int i = 0; i < 4; ++i) out_scalars[5 + i] = r.ransac.model_indices[i];
  out_scalars[9] = static_cast<int>(r.best_inlier_per_keypoint.size());
  *out_ratio = r.inlier_ratio;
  if (r.ransac.success) std::memcpy(out_T, r.ransac.T, sizeof(double) * 12);
  for (size_t i = 0; i < r.ransac.inliers.size(); ++i) {
    out_inliers[i] = r.ransac.inliers[i];
    out_inlier_dist[i] = r.ransac.inlier_distances[i];
  }
  for (size_t i = 0; i < r.best_inlier_per_keypoint.size(); ++i)
    out_best_per_keypoint[i] = r.best_inlier_per_keypoint[i];
}

// ---- whole query path for a batch of vertices, threaded like common::ParallelProcess ----
// (contiguous blocks of query vertices over num_threads std::threads:
//  common/maplab-common/include/maplab-common/parallel-process.h:49-92, driven from
//  loop-closure-handler/src/loop-detector-node.cc:867-873). Per vertex: ProjectDescriptors ->
// Find -> correspondence assembly -> HandleLoopClosure (queryVertexInDatabase, :668-766).
// out_scalars: per vertex {accepted, num_inliers, iterations, num_matches, ransac_success};
// stage_seconds: summed over threads {project, find, verify}.
int lco_query_batch(void* h, int num_frames, const int64_t* ts, const int64_t* vertex,
                    const int64_t* mission, const int* frame_index, const int* n_desc,
                    const uint8_t* bits, int bytes_per_desc, const double* keypoints,
                    const double* landmark_xyz, int64_t num_landmarks, const lco_camera* cams,
                    int n_cams, int min_inlier_count, double min_inlier_ratio, double pixel_sigma,
                    int num_iters, uint32_t seed, int rng_mapping, int num_threads,
                    int* out_scalars, double* out_T, double* stage_seconds) {
  LoopDetector* ld = static_cast<LoopDetector*>(h);
  const std::vector<Camera> cam_vec = ToCams(cams, n_cams);
  struct V {
    int first, count;
  };
  std::vector<V> verts;
  std::vector<int64_t> desc_off(num_frames + 1, 0);
  for (int f = 0; f < num_frames; ++f) {
    desc_off[f + 1] = desc_off[f] + n_desc[f];
    if (f > 0 && vertex[f] == vertex[f - 1])
      ++verts.back().count;
    else
      verts.push_back(V{f, 1});
  }
  const int nv = static_cast<int>(verts.size());
  HandlerSettings hs;
  hs.min_inlier_count = min_inlier_count;
  hs.min_inlier_ratio = min_inlier_ratio;
  hs.ransac_pixel_sigma = pixel_sigma;
  hs.num_ransac_iters = num_iters;
  hs.seed = seed;
  hs.rng_mapping = rng_mapping;
  const int dim = ld->fixed_projection().target_dim;
  std::vector<double> t_stage(3 * std::max(num_threads, 1), 0.0);
  auto work = [&](int tid, int v0, int v1) {
    using clk = std::chrono::steady_clock;
    for (int v = v0; v < v1; ++v) {
      const V& vx = verts[v];
      std::vector<ProjectedImage> ims(vx.count);
      std::vector<const ProjectedImage*> ptrs;
      auto t0 = clk::now();
      for (int c = 0; c < vx.count; ++c) {
        const int f = vx.first + c;
        ims[c].timestamp_ns = ts[f];
        ims[c].vertex_id = vertex[f];
        ims[c].frame_index = frame_index[f];
        ims[c].mission_id = mission[f];
        ims[c].dim = dim;
        ims[c].projected_descriptors.resize(static_cast<size_t>(n_desc[f]) * dim);
        ld->ProjectDescriptors(bits + desc_off[f] * bytes_per_desc, bytes_per_desc, n_desc[f],
                               ims[c].projected_descriptors.data());
        ptrs.push_back(&ims[c]);
      }
      auto t1 = clk::now();
      std::vector<Match> matches;
      ld->Find(ptrs, &matches);
      auto t2 = clk::now();
      VerifyInput in;
      in.n = static_cast<int>(matches.size());
      for (const Match& m : matches) {
        int f = vx.first;
        for (int c = 0; c < vx.count; ++c)
          if (frame_index[vx.first + c] == m.query_frame_index) f = vx.first + c;
        const int64_t qd = desc_off[f] + m.query_keypoint;
        in.keypoints.push_back(keypoints[2 * qd]);
        in.keypoints.push_back(keypoints[2 * qd + 1]);
        in.frame_index.push_back(m.query_frame_index);
        in.keypoint_index.push_back(m.query_keypoint);
        for (int a = 0; a < 3; ++a) {
          const bool known = m.landmark >= 0 && m.landmark < num_landmarks;
          in.landmarks.push_back(known ? landmark_xyz[3 * m.landmark + a]
                                       : std::numeric_limits<double>::quiet_NaN());
        }
      }
      VerifyResult r;
      HandleLoopClosure(in, cam_vec, hs, &r);
      auto t3 = clk::now();
      int* sc = out_scalars + 5 * static_cast<size_t>(v);
      sc[0] = r.accepted;
      sc[1] = r.num_inliers;
      sc[2] = r.ransac.iterations;
      sc[3] = in.n;
      sc[4] = r.ransac.success;
      if (r.ransac.success)
        std::memcpy(out_T + 12 * static_cast<size_t>(v), r.ransac.T, sizeof(double) * 12);
      else
        std::memset(out_T + 12 * static_cast<size_t>(v), 0, sizeof(double) * 12);
      t_stage[3 * tid + 0] += std::chrono::duration<double>(t1 - t0).count();
      t_stage[3 * tid + 1] += std::chrono::duration<double>(t2 - t1).count();
      t_stage[3 * tid + 2] += std::chrono::duration<double>(t3 - t2).count();
    }
  };
  if (num_threads <= 1) {
    work(0, 0, nv);
  } else {
    std::vector<std::thread> pool;
    const int per = (nv + num_threads - 1) / num_threads;
    for (int t = 0; t < num_threads; ++t) {
      const int v0 = std::min(nv, t * per), v1 = std::min(nv, (t + 1) * per);
      pool.emplace_back(work, t, v0, v1);
    }
    for (auto& th : pool) th.join();
  }
  for (int s2 = 0; s2 < 3; ++s2) {
    stage_seconds[s2] = 0;
    for (int t = 0; t < std::max(num_threads, 1); ++t) stage_seconds[s2] += t_stage[3 * t + s2];
  }
  return nv;
}

int lco_delta_pose_gate(const double* T_map, const double* T_ransac, double max_pos_m, double max_rot_deg,
                        double* delta) {
  DeltaPose(T_map, T_ransac, &delta[0], &delta[1]);
  return DeltaPoseGate(T_map, T_ransac, max_pos_m, max_rot_deg) ? 1 : 0;
}
int lco_transformation_ransac(const double* quats, const double* positions, int n, int num_iterations,
                              double thr_rad, double thr_m, uint32_t seed, int rng_mapping, double* out_quat,
                              double* out_pos, int* inlier_indices) {
  return TransformationRansac(quats, positions, n, num_iterations, thr_rad, thr_m, seed, rng_mapping,
                              out_quat, out_pos, inlier_indices);
}
void lco_uniform_indices(uint32_t seed, int mapping, uint32_t n, int count, int* out) {
  RansacRng rng(seed, mapping);
  for (int i = 0; i < count; ++i) out[i] = UniformIndex(&rng, n, mapping);
}
void lco_yaw_only(const double* q, double* out) { YawOnly(q, out); }
}  // extern "C"
