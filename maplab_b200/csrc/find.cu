// Batched LoopDetector::Find (matching-based-loopclosure/src/matching-based-engine.cc:48-168):
// kNN for every query descriptor, then kernel 3 per query frame and, for multi-camera vertices,
// the vertex-level second pass (:147-165).
#include <cstdlib>
#include <algorithm>
#include <limits>
#include <unordered_set>

#include "covis.h"
#include "detector.h"
#include "summary_map.h"

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
  std::vector<long long> fin_off;
  if (!FindOnDevice(frames, num_frames, d_idx, d_dist, k, &fin_off, err)) return false;
  const int64_t nvx = static_cast<int64_t>(fin_off.size()) - 1;
  for (int64_t v = 0; v <= nvx; ++v) match_offsets[v] = fin_off[v];
  *num_vertices = nvx;
  *num_matches = fin_off[nvx];
  if (fin_off[nvx] > capacity || fin_off[nvx] == 0) return true;  // caller retries with more room
  if (!Cuda(cudaMemcpyAsync(matches, d_covis_[5].p, sizeof(mlc_match) * fin_off[nvx],
                            cudaMemcpyDeviceToHost, stream_), "D2H matches", err))
    return false;
  return Cuda(cudaStreamSynchronize(stream_), "find", err);
}

// Kernel 3 for a batch; leaves the accepted matches of all query vertices compacted in
// d_covis_[5] (canonical order) and returns their per-vertex offsets.
bool Detector::FindOnDevice(const mlc_frame* frames, int64_t num_frames, const int32_t* d_idx,
                            const float* d_dist, int k, std::vector<long long>* fin_off_out,
                            std::string* err) {
  fin_off_out->assign(1, 0);
  if (s_.scoring != 0 && s_.scoring != 1) {
    *err = "unknown scoring function (0 accumulation, 1 probabilistic)";
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
    if (frames[f].num_descriptors < 0) {
      *err = "negative descriptor count";
      return false;
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
  const int grid = std::min(nf, sm_count_ * CovisCtasPerSm(max_slots));
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
  a.scoring = s_.scoring;
  a.group_bits = 1;
  while ((1ull << a.group_bits) <= keyframes_.size()) ++a.group_bits;
  a.landmark_bits = 1;
  while (a.landmark_bits < 64 && (max_lm_key_ >> a.landmark_bits) != 0) ++a.landmark_bits;
  a.num_db_descriptors = NumDescriptors();
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
    if (!Cuda(LaunchCovis(b, max_slots, std::min(nv2, sm_count_ * CovisCtasPerSm(max_slots)), stream_), "vertex covisibility kernel", err))
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
  std::vector<long long>& fin_off = *fin_off_out;
  fin_off.assign(nvx + 1, 0);
  int multi = 0;
  for (int vi = 0; vi < nvx; ++vi) {
    const Vertex& v = vertices[vi];
    fin[vi] = items[v.first_frame];
    fin_counts[vi] = v.num_frames < 2 ? counts[v.first_frame] : counts[nf + multi++];
    fin_off[vi + 1] = fin_off[vi] + fin_counts[vi];
  }
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
  return Cuda(LaunchCompactMatches(b_out.as<mlc_match>(), d_fin, d_fin_cnt, d_fin_off, nvx,
                                   b_final.as<mlc_match>(), stream_), "compact matches", err);
}

}  // namespace mlc

namespace mlc {
namespace {
// Correspondence assembly of handleLoopClosure (loop-closure-handler.cc:272-366): query keypoint
// measurement, camera (frame) index, keypoint index and G_landmark position of every match.
__global__ void gather_correspondences_kernel(const mlc_match* __restrict__ m, int64_t total,
                                              const int64_t* __restrict__ frame_desc_offset,
                                              const int32_t* __restrict__ frame_index,
                                              const double* __restrict__ keypoints,
                                              const double* __restrict__ lm_xyz, int64_t num_lm,
                                              double* __restrict__ out_kp, int32_t* __restrict__ out_ci,
                                              int32_t* __restrict__ out_ki, double* __restrict__ out_lm) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const mlc_match r = m[i];
  const int64_t qd = frame_desc_offset[r.query_frame] + r.query_keypoint;
  out_kp[2 * i] = keypoints[2 * qd];
  out_kp[2 * i + 1] = keypoints[2 * qd + 1];
  out_ci[i] = frame_index[r.query_frame];
  out_ki[i] = r.query_keypoint;
  const bool known = r.landmark >= 0 && r.landmark < num_lm;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  for (int c = 0; c < 3; ++c) out_lm[3 * i + c] = known ? lm_xyz[3 * r.landmark + c] : nan;
}
}  // namespace

bool Detector::SetLandmarkPositions(const double* xyz, int64_t n, std::string* err) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  if (n < 0 || (n > 0 && !xyz)) {
    *err = "bad landmark table";
    return false;
  }
  if (!Cuda(d_landmark_xyz_.Reserve(sizeof(double) * 3 * n + 16), "alloc landmarks", err)) return false;
  if (n > 0 && !Cuda(cudaMemcpyAsync(d_landmark_xyz_.p, xyz, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, stream_),
                     "H2D landmarks", err))
    return false;
  num_landmark_xyz_ = n;
  return Cuda(cudaStreamSynchronize(stream_), "landmark upload", err);
}

bool Detector::SetLandmarkPositionsDevice(const double* d_xyz, int64_t n, std::string* err) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  if (n < 0 || (n > 0 && !d_xyz)) {
    *err = "bad landmark table";
    return false;
  }
  if (!Cuda(d_landmark_xyz_.Reserve(sizeof(double) * 3 * n + 16), "alloc landmarks", err)) return false;
  if (n > 0 && !Cuda(cudaMemcpy(d_landmark_xyz_.p, d_xyz, sizeof(double) * 3 * n, cudaMemcpyDeviceToDevice),
                     "copy landmarks", err))
    return false;
  num_landmark_xyz_ = n;
  return true;
}

bool Detector::AddSummaryMap(const void* blob, size_t size, int64_t mission_id, int64_t first_vertex_id,
                             int64_t first_landmark_id, int64_t* sizes5, std::string* err) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  SummaryMap map;
  SummaryMapImages images;
  if (!map.Parse(blob, size, err) || !GroupSummaryMapByObserver(map, &images, err)) return false;
  if (map.num_observations() > 0 && static_cast<int>(map.descriptor_rows) != dim()) {
    *err = "summary map: descriptor dimensionality differs from the vocabulary's target dimensionality";
    return false;
  }
  if (first_landmark_id < 0) {
    *err = "summary map: negative landmark id";
    return false;
  }
  const int64_t observers = map.num_observers(), landmarks = map.num_landmarks();
  // landmark table: keep what is there, unknown ids in a gap stay NaN (= "no position")
  const int64_t old_n = num_landmark_xyz_, new_n = std::max(old_n, first_landmark_id + landmarks);
  std::vector<double> xyz(static_cast<size_t>(new_n) * 3, std::numeric_limits<double>::quiet_NaN());
  if (old_n > 0) {
    if (!Cuda(cudaMemcpyAsync(xyz.data(), d_landmark_xyz_.p, sizeof(double) * 3 * old_n, cudaMemcpyDeviceToHost,
                              stream_),
              "D2H landmarks", err) ||
        !Cuda(cudaStreamSynchronize(stream_), "landmark download", err))
      return false;
  }
  for (int64_t i = 0; i < 3 * landmarks; ++i)
    xyz[static_cast<size_t>(3 * first_landmark_id + i)] = static_cast<double>(map.G_landmark_position[i]);
  std::vector<mlc_frame> frames(static_cast<size_t>(observers));
  for (int64_t o = 0; o < observers; ++o) {
    frames[o].timestamp_ns = 0;  // "not relevant": the map has its own mission id
    frames[o].vertex_id = first_vertex_id + o;
    frames[o].mission_id = mission_id;
    frames[o].frame_index = 0;  // kFrameIndex
    frames[o].num_descriptors = images.num_descriptors[o];
  }
  for (int64_t& l : images.landmark_index) l += first_landmark_id;
  if (!InsertBatch(frames.data(), observers, images.proj.data(), images.landmark_index.data(), false, err))
    return false;
  if (!SetLandmarkPositions(xyz.data(), new_n, err)) return false;
  if (sizes5) {
    sizes5[0] = landmarks;
    sizes5[1] = observers;
    sizes5[2] = map.num_observations();
    sizes5[3] = map.descriptor_rows;
    sizes5[4] = map.descriptor_cols;
  }
  return EnsureIndex(err);
}

bool Detector::QueryBatch(const mlc_frame* frames, int64_t num_frames, const uint8_t* bits,
                          int bytes_per_desc, const double* keypoints, bool inputs_on_device,
                          const mlc_camera* cams, int num_cams, const mlc_ransac_settings& rs,
                          mlc_pose_result* results, int64_t* num_vertices, mlc_match* matches,
                          int64_t capacity, int64_t* match_offsets, int64_t* num_matches,
                          uint8_t* inlier_flags, std::string* err) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  *num_vertices = 0;
  if (num_matches) *num_matches = 0;
  if (num_frames == 0) return true;
  if (s_.shard_count > 1) {
    *err = "sharded detector: merge the per-shard kNN lists first and call mlc_query_from_knn_device";
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
  const uint8_t* d_bits = bits;
  const double* d_kp = keypoints;
  // Host buffers: copy in chunks on copy_stream_ (bits first, keypoints — needed only by kernel 4 —
  // last); stream_ waits for chunk i right before it projects / coarse-searches it, so the PCIe
  // transfer of the later chunks hides behind kernels 1 and 2a of the earlier ones.
  // 1/2/4/8 equal chunks measured: 207k/216k/211k/190k keyframes/s (round 1); geometric chunk sizes (1/8, 1/8, 1/4,
  // 1/2, so that only a small first copy is exposed): 250k against 254k for two halves (round 2) — the coarse
  // search loses more on small launches than the earlier start gains
  int chunks = (!inputs_on_device && n >= 65536) ? 2 : 1;
  if (const char* env = getenv("MLC_COPY_CHUNKS")) {
    const int v = atoi(env);
    if (v >= 1 && v <= kCopyChunks && !inputs_on_device) chunks = v;
  }
  const int64_t per_chunk = (n + chunks - 1) / chunks;
  if (!inputs_on_device && n > 0) {
    const size_t bb = static_cast<size_t>(n) * bytes_per_desc;
    if (!Cuda(d_bits_.Reserve(bb), "alloc", err) || !Cuda(d_query_[0].Reserve(sizeof(double) * 2 * n), "alloc", err))
      return false;
    // buffers of the previous call may still be in use on stream_ (never across a completed call,
    // every call ends synchronised) — the copy stream starts after everything enqueued so far
    if (!Cuda(cudaEventRecord(ev_copy_[kCopyChunks], stream_), "event", err) ||
        !Cuda(cudaStreamWaitEvent(copy_stream_, ev_copy_[kCopyChunks], 0), "wait", err))
      return false;
    for (int c = 0; c < chunks; ++c) {
      const int64_t s0 = c * per_chunk, cnt = std::min<int64_t>(per_chunk, n - s0);
      if (cnt <= 0) break;
      if (!Cuda(cudaMemcpyAsync(d_bits_.as<uint8_t>() + s0 * bytes_per_desc, bits + s0 * bytes_per_desc,
                                static_cast<size_t>(cnt) * bytes_per_desc, cudaMemcpyHostToDevice, copy_stream_),
                "H2D bits", err) ||
          !Cuda(cudaEventRecord(ev_copy_[c], copy_stream_), "event", err))
        return false;
    }
    if (!Cuda(cudaMemcpyAsync(d_query_[0].p, keypoints, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, copy_stream_),
              "H2D keypoints", err) ||
        !Cuda(cudaEventRecord(ev_copy_[kCopyChunks], copy_stream_), "event", err))
      return false;
    d_bits = d_bits_.as<uint8_t>();
    d_kp = d_query_[0].as<double>();
  }
  stage_valid_ = false;
  cudaEventRecord(ev_stage_[0], stream_);
  if (n > 0 && s_.engine == 2) {
    // hnsw engine: the descriptor bytes are the floats; exhaustive kNN instead of coarse search + list scan
    for (int c = 0; c < chunks && !inputs_on_device; ++c)
      if (!Cuda(cudaStreamWaitEvent(stream_, ev_copy_[c], 0), "wait", err)) return false;
    if (!ProjectDevice(d_bits, bytes_per_desc, n, d_q_.as<float>(), stream_, err)) return false;
    cudaEventRecord(ev_stage_[1], stream_);
    cudaEventRecord(ev_stage_[2], stream_);
    if (!KnnDevice(d_q_.as<float>(), n, k, d_idx_.as<int32_t>(), d_dist_.as<float>(), stream_, err)) return false;
    if (!inputs_on_device && !Cuda(cudaStreamWaitEvent(stream_, ev_copy_[kCopyChunks], 0), "wait", err))
      return false;
  } else if (n > 0) {
    const int nw = s_.num_closest_words;
    if (!Cuda(d_cells_.Reserve(static_cast<size_t>(n) * nw * 4), "alloc visit list", err)) return false;
    for (int c = 0; c < chunks; ++c) {
      const int64_t s0 = c * per_chunk, cnt = std::min<int64_t>(per_chunk, n - s0);
      if (cnt <= 0) break;
      if (!inputs_on_device && !Cuda(cudaStreamWaitEvent(stream_, ev_copy_[c], 0), "wait", err)) return false;
      if (!ProjectDevice(d_bits + s0 * bytes_per_desc, bytes_per_desc, cnt, d_q_.as<float>() + s0 * dim(), stream_, err))
        return false;
      if (c == chunks - 1) cudaEventRecord(ev_stage_[1], stream_);
      if (!CoarseDevice(d_q_.as<float>() + s0 * dim(), cnt, nw, d_cells_.as<int32_t>() + s0 * nw, stream_, err))
        return false;
    }
    cudaEventRecord(ev_stage_[2], stream_);
    if (!ScanDevice(d_q_.as<float>(), d_cells_.as<int32_t>(), n, k, d_idx_.as<int32_t>(),
                    d_dist_.as<float>(), stream_, err))
      return false;
    if (!inputs_on_device && !Cuda(cudaStreamWaitEvent(stream_, ev_copy_[kCopyChunks], 0), "wait", err))
      return false;
  } else {
    cudaEventRecord(ev_stage_[1], stream_);
    cudaEventRecord(ev_stage_[2], stream_);
  }
  cudaEventRecord(ev_stage_[3], stream_);
  const bool ok = QueryFromKnn(frames, num_frames, d_idx_.as<int32_t>(), d_dist_.as<float>(), k, d_kp,
                               cams, num_cams, rs, results, num_vertices, matches, capacity,
                               match_offsets, num_matches, inlier_flags, err);
  stage_valid_ = ok;
  return ok;
}

bool Detector::QueryFromKnn(const mlc_frame* frames, int64_t num_frames, const int32_t* d_idx,
                            const float* d_dist, int k, const double* d_keypoints,
                            const mlc_camera* cams, int num_cams, const mlc_ransac_settings& rs,
                            mlc_pose_result* results, int64_t* num_vertices, mlc_match* matches,
                            int64_t capacity, int64_t* match_offsets, int64_t* num_matches,
                            uint8_t* inlier_flags, std::string* err) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  *num_vertices = 0;
  if (num_matches) *num_matches = 0;
  if (num_frames == 0) return true;
  std::vector<long long> fin_off;
  if (!FindOnDevice(frames, num_frames, d_idx, d_dist, k, &fin_off, err)) return false;
  cudaEventRecord(ev_stage_[4], stream_);
  const int64_t nvx = static_cast<int64_t>(fin_off.size()) - 1;
  const int64_t total = fin_off[nvx];
  *num_vertices = nvx;
  if (num_matches) *num_matches = total;
  if (match_offsets)
    for (int64_t v = 0; v <= nvx; ++v) match_offsets[v] = fin_off[v];
  for (int64_t f = 0; f < num_frames; ++f) {
    if (frames[f].frame_index < 0 || frames[f].frame_index >= num_cams) {
      *err = "frame_index is the camera index of the query rig: out of range";
      return false;
    }
  }
  // per-frame tables + problem offsets
  std::vector<int64_t> desc_off(num_frames), prob_off(fin_off.begin(), fin_off.end());
  std::vector<int32_t> fidx(num_frames);
  int64_t at = 0;
  for (int64_t f = 0; f < num_frames; ++f) {
    desc_off[f] = at;
    fidx[f] = frames[f].frame_index;
    at += frames[f].num_descriptors;
  }
  DevBuf &b_tab = d_query_[1], &b_corr = d_query_[2];
  const size_t t_doff = 0, t_fidx = sizeof(int64_t) * num_frames,
               t_poff = (t_fidx + sizeof(int32_t) * num_frames + 15) & ~static_cast<size_t>(15);
  if (!Cuda(b_tab.Reserve(t_poff + sizeof(int64_t) * (nvx + 1) + 16), "alloc", err)) return false;
  unsigned char* tab = b_tab.as<unsigned char>();
  if (!Cuda(cudaMemcpyAsync(tab + t_doff, desc_off.data(), sizeof(int64_t) * num_frames, cudaMemcpyHostToDevice, stream_), "H2D", err) ||
      !Cuda(cudaMemcpyAsync(tab + t_fidx, fidx.data(), sizeof(int32_t) * num_frames, cudaMemcpyHostToDevice, stream_), "H2D", err) ||
      !Cuda(cudaMemcpyAsync(tab + t_poff, prob_off.data(), sizeof(int64_t) * (nvx + 1), cudaMemcpyHostToDevice, stream_), "H2D", err))
    return false;
  const size_t c_kp = 0, c_lm = (sizeof(double) * 2 * total + 255) & ~static_cast<size_t>(255),
               c_ci = (c_lm + sizeof(double) * 3 * total + 255) & ~static_cast<size_t>(255),
               c_ki = (c_ci + sizeof(int32_t) * total + 255) & ~static_cast<size_t>(255);
  if (!Cuda(b_corr.Reserve(c_ki + sizeof(int32_t) * total + 256), "alloc", err)) return false;
  unsigned char* corr = b_corr.as<unsigned char>();
  if (total > 0) {
    gather_correspondences_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream_>>>(
        d_covis_[5].as<mlc_match>(), total, reinterpret_cast<const int64_t*>(tab + t_doff),
        reinterpret_cast<const int32_t*>(tab + t_fidx), d_keypoints, d_landmark_xyz_.as<double>(),
        num_landmark_xyz_, reinterpret_cast<double*>(corr + c_kp), reinterpret_cast<int32_t*>(corr + c_ci),
        reinterpret_cast<int32_t*>(corr + c_ki), reinterpret_cast<double*>(corr + c_lm));
    CountLaunch();
    if (!Cuda(cudaGetLastError(), "gather correspondences", err)) return false;
    if (matches && total <= capacity &&
        !Cuda(cudaMemcpyAsync(matches, d_covis_[5].p, sizeof(mlc_match) * total, cudaMemcpyDeviceToHost, stream_),
              "D2H matches", err))
      return false;
  }
  corr_grouped_ = true;  // canonical order: (query frame, keypoint, database descriptor)
  return RansacOnDevice(rs, cams, num_cams, nvx, total, reinterpret_cast<const int64_t*>(tab + t_poff),
                        reinterpret_cast<const double*>(corr + c_kp), reinterpret_cast<const int32_t*>(corr + c_ci),
                        reinterpret_cast<const int32_t*>(corr + c_ki), reinterpret_cast<const double*>(corr + c_lm),
                        results, inlier_flags, err);
}

bool Detector::Score(int scoring, const uint64_t* votes, const uint64_t* num_desc, int n, int64_t num_db,
                     float* scores, std::string* err) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  if (n <= 0 || num_db <= 0) return true;  // scoring.h:108-116: empty database / no ids -> no scores
  if (scoring != 0 && scoring != 1) {
    *err = "unknown scoring function (0 accumulation, 1 probabilistic)";
    return false;
  }
  DevBuf &b_in = d_covis_[6], &b_out = d_covis_[7];
  if (!Cuda(b_in.Reserve(sizeof(uint64_t) * 2 * n), "alloc", err) ||
      !Cuda(b_out.Reserve(sizeof(float) * n), "alloc", err))
    return false;
  uint64_t* d_votes = b_in.as<uint64_t>();
  if (!Cuda(cudaMemcpyAsync(d_votes, votes, sizeof(uint64_t) * n, cudaMemcpyHostToDevice, stream_), "H2D", err) ||
      !Cuda(cudaMemcpyAsync(d_votes + n, num_desc, sizeof(uint64_t) * n, cudaMemcpyHostToDevice, stream_), "H2D", err))
    return false;
  if (!Cuda(LaunchScore(reinterpret_cast<const unsigned long long*>(d_votes),
                        reinterpret_cast<const unsigned long long*>(d_votes + n), n, num_db, scoring,
                        b_out.as<float>(), stream_), "score kernel", err))
    return false;
  if (!Cuda(cudaMemcpyAsync(scores, b_out.p, sizeof(float) * n, cudaMemcpyDeviceToHost, stream_), "D2H", err))
    return false;
  return Cuda(cudaStreamSynchronize(stream_), "score", err);
}

}  // namespace mlc
