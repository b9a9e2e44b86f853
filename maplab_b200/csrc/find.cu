// Batched LoopDetector::Find (matching-based-loopclosure/src/matching-based-engine.cc:48-168):
// kNN for every query descriptor, then kernel 3 per query frame and, for multi-camera vertices,
// the vertex-level second pass (:147-165).
#include <algorithm>
#include <unordered_set>

#include "covis.h"
#include "detector.h"

namespace mlc {

bool Detector::FindBatch(const mlc_frame* frames, int64_t num_frames, const float* proj,
                         const uint8_t* bits, int bytes_per_desc, mlc_match* matches,
                         int64_t capacity, int64_t* match_offsets, int64_t* num_vertices,
                         int64_t* num_matches, std::string* err) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  *num_vertices = 0;
  *num_matches = 0;
  if (num_frames == 0) return true;  // Find on an empty list returns empty (engine.cc:52-56)
  if (s_.shard_count > 1) {
    *err = "sharded detector: merge the per-shard kNN lists first and call mlc_find_from_knn_device";
    return false;
  }
  if (!EnsureIndex(err)) return false;
  int64_t n = 0;
  for (int64_t f = 0; f < num_frames; ++f) {
    if (frames[f].num_descriptors < 0) {
      *err = "negative descriptor count";
      return false;
    }
    n += frames[f].num_descriptors;
  }
  const int k = NumNeighbors();
  const size_t qb = static_cast<size_t>(n) * dim() * 4, rb = static_cast<size_t>(n) * k * 4;
  if (!Cuda(d_q_.Reserve(qb + 16), "alloc", err) || !Cuda(d_idx_.Reserve(rb + 16), "alloc", err) ||
      !Cuda(d_dist_.Reserve(rb + 16), "alloc", err))
    return false;
  if (n > 0) {
    if (bits) {
      const size_t bb = static_cast<size_t>(n) * bytes_per_desc;
      if (!Cuda(d_bits_.Reserve(bb), "alloc", err)) return false;
      if (!Cuda(cudaMemcpyAsync(d_bits_.p, bits, bb, cudaMemcpyHostToDevice, stream_), "H2D bits", err))
        return false;
      if (!ProjectDevice(d_bits_.as<uint8_t>(), bytes_per_desc, n, d_q_.as<float>(), stream_, err))
        return false;
    } else {
      if (!proj) {
        *err = "null descriptors";
        return false;
      }
      if (!Cuda(cudaMemcpyAsync(d_q_.p, proj, qb, cudaMemcpyHostToDevice, stream_), "H2D queries", err))
        return false;
    }
    if (!KnnDevice(d_q_.as<float>(), n, k, d_idx_.as<int32_t>(), d_dist_.as<float>(), stream_, err))
      return false;
  }
  return FindFromKnn(frames, num_frames, d_idx_.as<int32_t>(), d_dist_.as<float>(), k, matches,
                     capacity, match_offsets, num_vertices, num_matches, err);
}

bool Detector::FindFromKnn(const mlc_frame* frames, int64_t num_frames, const int32_t* d_idx,
                           const float* d_dist, int k, mlc_match* matches, int64_t capacity,
                           int64_t* match_offsets, int64_t* num_vertices, int64_t* num_matches,
                           std::string* err) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  *num_vertices = 0;
  *num_matches = 0;
  if (num_frames == 0) return true;
  if (s_.scoring != 0) {
    *err = "scoring function not built: only 'accumulation' (0) is available on the device";
    return false;
  }
  if (!EnsureIndex(err)) return false;
  // ---- group adjacent frames by vertex (CHECK: all frames of a Find call share one vertex) ----
  struct Vertex {
    int first_frame, num_frames;
    int64_t slots;
  };
  std::vector<Vertex> vertices;
  std::vector<CovisItem> items(num_frames);
  std::unordered_set<int64_t> seen;
  int64_t at = 0;
  int max_slots = 0;
  for (int64_t f = 0; f < num_frames; ++f) {
    const bool same = !vertices.empty() && frames[f].vertex_id == frames[f - 1].vertex_id;
    if (!same) {
      if (!seen.insert(frames[f].vertex_id).second) {
        *err = "frames of one query vertex must be adjacent in the batch";
        return false;
      }
      vertices.push_back(Vertex{static_cast<int>(f), 0, 0});
    }
    Vertex& v = vertices.back();
    ++v.num_frames;
    const int64_t slots = static_cast<int64_t>(frames[f].num_descriptors) * k;
    v.slots += slots;
    if (slots > 8192 || frames[f].num_descriptors >= 32768 || v.num_frames > 32) {
      *err = "query frame too large for the device covisibility kernel (descriptors x k <= 8192 per frame)";
      return false;
    }
    max_slots = std::max<int>(max_slots, static_cast<int>(std::min<int64_t>(v.slots, 8192)));
    CovisItem& it = items[f];
    it.first = static_cast<int32_t>(at);
    it.count = frames[f].num_descriptors;
    it.frame = static_cast<int32_t>(f);
    it.make_unique = 1;
    it.ts = frames[f].timestamp_ns;
    it.mission = frames[f].mission_id;
    it.out_offset = at * k;
    at += frames[f].num_descriptors;
  }
  const int64_t n = at;
  std::vector<CovisItem> vitems;
  for (size_t vi = 0; vi < vertices.size(); ++vi) {
    const Vertex& v = vertices[vi];
    if (v.num_frames < 2) continue;
    for (int f = 0; f < v.num_frames; ++f) items[v.first_frame + f].make_unique = 0;
    CovisItem it{};
    it.first = v.first_frame;
    it.count = v.num_frames;
    it.frame = v.first_frame;
    it.make_unique = 1;
    it.out_offset = items[v.first_frame].out_offset;
    vitems.push_back(it);
  }
  const int nf = static_cast<int>(num_frames);
  const int nv2 = static_cast<int>(vitems.size());
  const int grid = std::min(nf, sm_count_ * 2);
  // device buffers: [0] items, [1] pass-1 matches, [2] counts (pass 1 then pass 2), [3] scratch,
  // [4] pass-1 offsets as long long, [5] final matches, [6] final offsets / vertex items
  DevBuf &b_items = d_covis_[0], &b_out = d_covis_[1], &b_cnt = d_covis_[2], &b_scr = d_covis_[3],
         &b_off = d_covis_[4], &b_final = d_covis_[5], &b_aux = d_covis_[6];
  const size_t out_elems = static_cast<size_t>(n) * k + 1;
  if (!Cuda(b_items.Reserve(sizeof(CovisItem) * (nf + nv2 + 1)), "alloc", err) ||
      !Cuda(b_out.Reserve(sizeof(mlc_match) * out_elems), "alloc", err) ||
      !Cuda(b_cnt.Reserve(sizeof(int) * (nf + nv2 + 1)), "alloc", err) ||
      !Cuda(b_scr.Reserve(sizeof(mlc_match) * CovisScratchMatches(max_slots, grid)), "alloc", err) ||
      !Cuda(b_off.Reserve(sizeof(long long) * (nf + 1)), "alloc", err))
    return false;
  std::vector<CovisItem> all_items(items);
  all_items.insert(all_items.end(), vitems.begin(), vitems.end());
  if (!Cuda(cudaMemcpyAsync(b_items.p, all_items.data(), sizeof(CovisItem) * all_items.size(),
                            cudaMemcpyHostToDevice, stream_), "H2D items", err))
    return false;
  CovisArgs a{};
  a.items = b_items.as<CovisItem>();
  a.num_items = nf;
  a.by_vertex = 0;
  a.k = k;
  a.knn_idx = d_idx;
  a.knn_dist = d_dist;
  a.desc_kf = d_desc_kf_.as<int32_t>();
  a.desc_lm = d_desc_lm_.as<int64_t>();
  a.kf_meta = d_kf_meta_.as<KeyframeMeta>();
  a.min_time_ns = s_.min_image_time_seconds * 1e9;  // kSecondsToNanoSeconds
  a.min_verify_matches_num = s_.min_verify_matches_num;
  a.fraction_best_scores = s_.fraction_best_scores;
  a.scratch = b_scr.as<mlc_match>();
  a.out_matches = b_out.as<mlc_match>();
  a.out_counts = b_cnt.as<int>();
  if (!Cuda(LaunchCovis(a, max_slots, grid, stream_), "covisibility kernel", err)) return false;
  if (nv2 > 0) {
    std::vector<long long> offs(nf);
    for (int f = 0; f < nf; ++f) offs[f] = items[f].out_offset;
    if (!Cuda(cudaMemcpyAsync(b_off.p, offs.data(), sizeof(long long) * nf, cudaMemcpyHostToDevice, stream_),
              "H2D", err))
      return false;
    CovisArgs b = a;
    b.items = b_items.as<CovisItem>() + nf;
    b.num_items = nv2;
    b.by_vertex = 1;
    b.in_matches = b_out.as<mlc_match>();
    b.in_counts = b_cnt.as<int>();
    b.in_offsets = b_off.as<long long>();
    b.out_counts = b_cnt.as<int>() + nf;
    // the offsets upload must not be overwritten before the kernel ran: offs lives until the sync below
    if (!Cuda(LaunchCovis(b, max_slots, std::min(nv2, sm_count_ * 2), stream_), "vertex covisibility kernel", err))
      return false;
    if (!Cuda(cudaStreamSynchronize(stream_), "covisibility", err)) return false;
  }
  std::vector<int> counts(nf + nv2);
  if (!Cuda(cudaMemcpyAsync(counts.data(), b_cnt.p, sizeof(int) * (nf + nv2), cudaMemcpyDeviceToHost, stream_),
            "D2H counts", err) ||
      !Cuda(cudaStreamSynchronize(stream_), "covisibility", err))
    return false;
  for (int c : counts) {
    if (c < 0) {
      *err = "more than 8192 matches survive the per-camera filter of one query vertex";
      return false;
    }
  }
  // ---- final per-vertex segments ----
  const int nvx = static_cast<int>(vertices.size());
  std::vector<CovisItem> fin(nvx);
  std::vector<int> fin_counts(nvx);
  std::vector<long long> fin_off(nvx + 1, 0);
  int multi = 0;
  for (int vi = 0; vi < nvx; ++vi) {
    const Vertex& v = vertices[vi];
    fin[vi] = items[v.first_frame];
    fin_counts[vi] = v.num_frames < 2 ? counts[v.first_frame] : counts[nf + multi++];
    fin_off[vi + 1] = fin_off[vi] + fin_counts[vi];
    match_offsets[vi] = fin_off[vi];
  }
  match_offsets[nvx] = fin_off[nvx];
  *num_vertices = nvx;
  *num_matches = fin_off[nvx];
  if (fin_off[nvx] > capacity) return true;  // caller sees num_matches > capacity and retries
  if (fin_off[nvx] == 0) return true;
  const size_t aux_bytes = sizeof(CovisItem) * nvx + sizeof(int) * nvx + sizeof(long long) * (nvx + 1) + 64;
  if (!Cuda(b_aux.Reserve(aux_bytes), "alloc", err) ||
      !Cuda(b_final.Reserve(sizeof(mlc_match) * fin_off[nvx]), "alloc", err))
    return false;
  unsigned char* aux = b_aux.as<unsigned char>();
  CovisItem* d_fin = reinterpret_cast<CovisItem*>(aux);
  long long* d_fin_off = reinterpret_cast<long long*>(aux + sizeof(CovisItem) * nvx);
  int* d_fin_cnt = reinterpret_cast<int*>(aux + sizeof(CovisItem) * nvx + sizeof(long long) * (nvx + 1));
  if (!Cuda(cudaMemcpyAsync(d_fin, fin.data(), sizeof(CovisItem) * nvx, cudaMemcpyHostToDevice, stream_), "H2D", err) ||
      !Cuda(cudaMemcpyAsync(d_fin_off, fin_off.data(), sizeof(long long) * (nvx + 1), cudaMemcpyHostToDevice, stream_), "H2D", err) ||
      !Cuda(cudaMemcpyAsync(d_fin_cnt, fin_counts.data(), sizeof(int) * nvx, cudaMemcpyHostToDevice, stream_), "H2D", err))
    return false;
  if (!Cuda(LaunchCompactMatches(b_out.as<mlc_match>(), d_fin, d_fin_cnt, d_fin_off, nvx,
                                 b_final.as<mlc_match>(), stream_), "compact matches", err))
    return false;
  if (!Cuda(cudaMemcpyAsync(matches, b_final.p, sizeof(mlc_match) * fin_off[nvx], cudaMemcpyDeviceToHost, stream_),
            "D2H matches", err))
    return false;
  return Cuda(cudaStreamSynchronize(stream_), "find", err);
}

}  // namespace mlc
