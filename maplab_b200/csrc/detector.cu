#include "detector.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

namespace mlc {

std::atomic<uint64_t> g_kernel_launches{0};

namespace {
size_t Align16(size_t x) { return (x + 15) & ~static_cast<size_t>(15); }
}  // namespace

bool Detector::Cuda(cudaError_t e, const char* what, std::string* err) const {
  if (e == cudaSuccess) return true;
  *err = std::string(what) + ": " + cudaGetErrorString(e);
  return false;
}

Detector::~Detector() {
  if (d_tree_blob_) cudaFree(d_tree_blob_);
  if (d_pq_blob_) cudaFree(d_pq_blob_);
  for (int8_t* p : proj_.b_image)
    if (p) cudaFree(p);
  lists_.Free();
  DevBuf* bufs[] = {&d_db_cells_, &d_desc_kf_, &d_kf_meta_, &d_q_,    &d_cells_,
                    &d_idx_,      &d_dist_,    &d_bits_,    &d_stats_};
  for (DevBuf* b : bufs) b->Free();
  d_own_desc_.Free();
  d_own_gidx_.Free();
  d_desc_lm_.Free();
  for (DevBuf& b : d_covis_) b.Free();
  for (DevBuf& b : d_ransac_) b.Free();
  for (DevBuf& b : d_query_) b.Free();
  d_landmark_xyz_.Free();
  d_coarse_scratch_.Free();
  if (ev0_) cudaEventDestroy(ev0_);
  if (ev1_) cudaEventDestroy(ev1_);
  for (cudaEvent_t e : ev_stage_)
    if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : ev_copy_)
    if (e) cudaEventDestroy(e);
  if (h_remaining_) cudaFreeHost(h_remaining_);
  for (cudaEvent_t e : ev_ransac_)
    if (e) cudaEventDestroy(e);
  for (cudaStream_t st : ransac_stream_)
    if (st) cudaStreamDestroy(st);
  CommDestroy();
  for (DevBuf* b : {&sh_counts_, &sh_q_all_, &sh_cells_all_, &sh_pidx_, &sh_pdist_, &sh_ridx_, &sh_rdist_}) b->Free();
  for (cudaEvent_t e : ev_comm_)
    if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : ev_scan_)
    if (e) cudaEventDestroy(e);
  if (comm_stream_) cudaStreamDestroy(comm_stream_);
  if (copy_stream_) cudaStreamDestroy(copy_stream_);
  if (stream_) cudaStreamDestroy(stream_);
}

bool Detector::Create(const mlc_settings& s, const void* blob, size_t size, std::string* err) {
  s_ = s;
  if (s_.shard_count <= 0) s_.shard_count = 1;
  if (s_.shard_rank < 0 || s_.shard_rank >= s_.shard_count || s_.shard_count > kMaxShards) {
    *err = "shard_rank out of range (at most 16 shards)";
    return false;
  }
  if (s_.num_closest_words <= 0 || s_.num_closest_words > 16) {
    *err = "num_closest_words must be in 1..16";
    return false;
  }
  if (s_.engine != 0 && s_.engine != 1 && s_.engine != 2) {
    *err = "unknown detector engine (0 imi, 1 imipq, 2 hnsw)";
    return false;
  }
  if (s_.engine == 2 && (s_.float_descriptor_dim <= 0 || s_.float_descriptor_dim > 4096 || s_.shard_count > 1 ||
                         s_.hnsw_m <= 0 || s_.hnsw_ef_construction <= 0 || s_.hnsw_ef_query <= 0)) {
    *err = "hnsw engine: needs float_descriptor_dim in 1..4096, positive hnsw_* settings (detector-settings.cc:69-71) "
           "and a single shard";
    return false;
  }
  if (s_.shard_mode != 0 && s_.shard_mode != 1) {
    *err = "unknown shard_mode (0 by descriptor index, 1 by cell)";
    return false;
  }
  if (s_.engine != 2 && !vocab_.Parse(blob, size, s_.engine == 1, err)) return false;
  vocab_hash_ = 1469598103934665603ull;
  for (size_t i = 0; s_.engine != 2 && i < size; ++i)
    vocab_hash_ = (vocab_hash_ ^ static_cast<const unsigned char*>(blob)[i]) * 1099511628211ull;
  if (s_.engine != 2 && vocab_.target_dim / 2 > 8) {
    *err = "target dimensionality > 16 is not supported";
    return false;
  }
  int count = 0;
  if (!Cuda(cudaGetDeviceCount(&count), "cudaGetDeviceCount", err)) return false;
  if (count == 0) {
    *err = "no CUDA device: the B200 loop-closure path has no CPU fallback";
    return false;
  }
  if (s_.device >= 0) {
    if (!Cuda(cudaSetDevice(s_.device), "cudaSetDevice", err)) return false;
  }
  if (!Cuda(cudaGetDevice(&device_), "cudaGetDevice", err)) return false;
  cudaDeviceProp prop;
  if (!Cuda(cudaGetDeviceProperties(&prop, device_), "cudaGetDeviceProperties", err)) return false;
  if (prop.major != 10) {
    *err = "device is not sm_100 (Blackwell B200); kernels are built for sm_100a only";
    return false;
  }
  sm_count_ = prop.multiProcessorCount;
  if (!Cuda(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking), "cudaStreamCreate", err))
    return false;
  if (!Cuda(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking), "cudaStreamCreate", err))
    return false;
  for (cudaEvent_t& e : ev_copy_)
    if (!Cuda(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate", err)) return false;
  for (cudaEvent_t& e : ev_ransac_)
    if (!Cuda(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate", err)) return false;
  for (cudaStream_t& st : ransac_stream_)
    if (!Cuda(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking), "cudaStreamCreate", err)) return false;
  if (!Cuda(cudaEventCreate(&ev0_), "cudaEventCreate", err)) return false;
  if (!Cuda(cudaEventCreate(&ev1_), "cudaEventCreate", err)) return false;
  for (cudaEvent_t& e : ev_stage_)
    if (!Cuda(cudaEventCreate(&e), "cudaEventCreate", err)) return false;

  if (s_.engine == 2) return true;  // no vocabulary: the descriptors are floats, the search exhaustive
  fp_.Build(vocab_.projection, vocab_.target_dim);
  if (fp_.kp <= 512 && fp_.dim <= 12) {
    if (!Cuda(BuildProjectionDevice(fp_, &proj_), "BuildProjectionDevice", err)) return false;
  }
  tree1_.Build(vocab_.words1);
  tree2_.Build(vocab_.words2);
  if (tree1_.max_depth > 47 || tree2_.max_depth > 47) {
    *err = "vocabulary kd-tree deeper than the device traversal stack";
    return false;
  }
  if (!UploadTrees(err)) return false;
  return s_.engine == 1 ? UploadPq(err) : true;
}

// imipq: coarse words + residual quantiser centres on the device
// (InvertedMultiProductQuantizationIndex ctor, …-quantization-index.h:70-130).
bool Detector::UploadPq(std::string* err) {
  const VocabularyFile& v = vocab_;
  const int half = v.pq_components / 2, sub = v.target_dim / 2;
  if (!v.has_pq || v.pq_components <= 0 || (v.pq_components & 1) || v.pq_components > 12 ||
      v.pq_centers <= 0 || v.pq_centers > 256 || v.pq_dim_per_comp <= 0 || half * v.pq_dim_per_comp != sub) {
    *err = "imipq: unsupported product quantiser shape (components even and <= 12, centres <= 256, "
           "components/2 * dims_per_component == target_dim/2)";
    return false;
  }
  const size_t per_word = static_cast<size_t>(half) * v.pq_centers * v.pq_dim_per_comp;
  if (v.pq_centers1.v.size() != per_word * v.words1.cols || v.pq_centers2.v.size() != per_word * v.words2.cols) {
    *err = "imipq: quantiser centre matrices do not match the vocabulary";
    return false;
  }
  const size_t n_w1 = v.words1.v.size(), n_w2 = v.words2.v.size();
  const size_t n_c1 = v.pq_centers1.v.size(), n_c2 = v.pq_centers2.v.size();
  std::vector<float> blob;
  blob.reserve(n_w1 + n_w2 + n_c1 + n_c2);
  blob.insert(blob.end(), v.words1.v.begin(), v.words1.v.end());
  blob.insert(blob.end(), v.words2.v.begin(), v.words2.v.end());
  blob.insert(blob.end(), v.pq_centers1.v.begin(), v.pq_centers1.v.end());
  blob.insert(blob.end(), v.pq_centers2.v.begin(), v.pq_centers2.v.end());
  if (!Cuda(cudaMalloc(&d_pq_blob_, blob.size() * 4), "cudaMalloc(pq)", err)) return false;
  if (!Cuda(cudaMemcpy(d_pq_blob_, blob.data(), blob.size() * 4, cudaMemcpyHostToDevice), "upload pq", err))
    return false;
  const float* base = static_cast<const float*>(d_pq_blob_);
  pq_.words1 = base;
  pq_.words2 = base + n_w1;
  pq_.centers1 = base + n_w1 + n_w2;
  pq_.centers2 = base + n_w1 + n_w2 + n_c1;
  pq_.sub_dim = sub;
  pq_.half_ncomp = half;
  pq_.dim_per_comp = v.pq_dim_per_comp;
  pq_.num_centers = v.pq_centers;
  pq_.num_words1 = v.words1.cols;
  pq_.num_words2 = v.words2.cols;
  return true;
}

cudaError_t Detector::LaunchScan(const float* d_q, int64_t n_q, const int32_t* d_cells, int nw, int k,
                                 int32_t* d_idx, float* d_dist, cudaStream_t stream) {
  if (s_.engine == 1)
    return LaunchImipqScan(pq_, d_q, n_q, d_cells, nw, lists_.cell_info, lists_.lists, k, d_idx, d_dist,
                           sm_count_, stream);
  return LaunchImiScan(dim(), d_q, n_q, d_cells, nw, lists_.cell_info, lists_.lists, k, d_idx, d_dist,
                       sm_count_, stream);
}

bool Detector::UploadTrees(std::string* err) {
  CoarseParams& c = coarse_;
  size_t off = 0;
  auto place = [&](size_t bytes) {
    const size_t at = off;
    off = Align16(off + bytes);
    return static_cast<uint32_t>(at);
  };
  c.off_nodes1 = place(tree1_.nodes.size() * sizeof(KdNodeDev));
  c.off_buckets1 = place(tree1_.bucket_points.size() * 4);
  c.off_cloud1 = place(tree1_.cloud.size() * 4);
  c.off_nodes2 = place(tree2_.nodes.size() * sizeof(KdNodeDev));
  c.off_buckets2 = place(tree2_.bucket_points.size() * 4);
  c.off_cloud2 = place(tree2_.cloud.size() * 4);
  c.packed_bytes = static_cast<uint32_t>(off);
  std::vector<unsigned char> blob(off, 0);
  std::memcpy(blob.data() + c.off_nodes1, tree1_.nodes.data(), tree1_.nodes.size() * sizeof(KdNodeDev));
  std::memcpy(blob.data() + c.off_buckets1, tree1_.bucket_points.data(), tree1_.bucket_points.size() * 4);
  std::memcpy(blob.data() + c.off_cloud1, tree1_.cloud.data(), tree1_.cloud.size() * 4);
  std::memcpy(blob.data() + c.off_nodes2, tree2_.nodes.data(), tree2_.nodes.size() * sizeof(KdNodeDev));
  std::memcpy(blob.data() + c.off_buckets2, tree2_.bucket_points.data(), tree2_.bucket_points.size() * 4);
  std::memcpy(blob.data() + c.off_cloud2, tree2_.cloud.data(), tree2_.cloud.size() * 4);
  if (!Cuda(cudaMalloc(&d_tree_blob_, off), "cudaMalloc(trees)", err)) return false;
  if (!Cuda(cudaMemcpy(d_tree_blob_, blob.data(), off, cudaMemcpyHostToDevice), "upload trees", err))
    return false;
  const unsigned char* base = static_cast<const unsigned char*>(d_tree_blob_);
  c.packed = d_tree_blob_;
  c.nodes1 = reinterpret_cast<const KdNodeDev*>(base + c.off_nodes1);
  c.buckets1 = reinterpret_cast<const int32_t*>(base + c.off_buckets1);
  c.cloud1 = reinterpret_cast<const float*>(base + c.off_cloud1);
  c.nodes2 = reinterpret_cast<const KdNodeDev*>(base + c.off_nodes2);
  c.buckets2 = reinterpret_cast<const int32_t*>(base + c.off_buckets2);
  c.cloud2 = reinterpret_cast<const float*>(base + c.off_cloud2);
  c.sub_dim = vocab_.target_dim / 2;
  c.num_words1 = vocab_.words1.cols;
  c.num_words2 = vocab_.words2.cols;
  c.max_radius2 = s_.knn_max_radius * s_.knn_max_radius;
  c.max_error2 = (1 + s_.knn_epsilon) * (1 + s_.knn_epsilon);
  c.stage_in_smem = off <= 160 * 1024 ? 1 : 0;
  return true;
}

bool Detector::Clear(std::string* err) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  keyframes_.clear();
  keyframe_keys_.clear();
  max_lm_key_ = 0;
  num_desc_ = num_own_ = 0;
  pend_desc_.clear();
  pend_gidx_.clear();
  pend_lm_.clear();
  d_own_desc_.used = d_own_gidx_.used = d_desc_lm_.used = 0;
  lists_.Free();
  index_dirty_ = true;
  last_valid_ = false;
  (void)err;
  return true;
}

// getNumNeighborsToSearch, matching-based-engine.cc:319-338 (int size compared against doubles).
int Detector::NumNeighbors() const {
  int k = s_.num_nearest_neighbors;
  if (k == -1) {
    const int n = static_cast<int>(NumDescriptors());
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

bool Detector::ProjectDevice(const uint8_t* d_bits, int bytes_per_desc, int64_t n, float* d_out,
                             cudaStream_t stream, std::string* err) {
  if (s_.engine == 2) {
    // HSNWIndexInterface::ProjectDescriptors (hnsw-index-interface.h:155-163): the descriptor bytes ARE the floats
    if (bytes_per_desc != 4 * dim()) {
      *err = "hnsw engine: descriptors must be float_descriptor_dim floats";
      return false;
    }
    return n == 0 || Cuda(cudaMemcpyAsync(d_out, d_bits, static_cast<size_t>(n) * bytes_per_desc,
                                          cudaMemcpyDeviceToDevice, stream), "reinterpret descriptors", err);
  }
  if (!proj_.fp) {
    *err = "projection matrix shape unsupported by the tensor-core kernel (need <= 512 columns, <= 12 rows)";
    return false;
  }
  // ProjectDescriptorBlock: CHECK the matrix consumes no more bits than a descriptor has
  // (471-column FREAK matrices use the first 471 of 512 bits, descriptor-projection.cc:35-43).
  if (bytes_per_desc * 8 < proj_.kp) {
    *err = "descriptor shorter than the projection matrix";
    return false;
  }
  return Cuda(LaunchProjection(proj_, d_bits, bytes_per_desc, n, d_out, sm_count_, stream),
              "projection kernel", err);
}

bool Detector::Project(const uint8_t* bits, int bytes_per_desc, int64_t n, float* out,
                       std::string* err) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  if (n == 0) return true;  // descriptor-projection.cc:20-22
  if (n < 0 || bytes_per_desc <= 0 || (s_.engine != 2 && bytes_per_desc % 16 != 0)) {
    *err = "bad descriptor block shape (bytes per descriptor must be a multiple of 16)";
    return false;
  }
  const size_t in_bytes = static_cast<size_t>(n) * bytes_per_desc;
  const size_t out_bytes = static_cast<size_t>(n) * dim() * sizeof(float);
  if (!Cuda(d_bits_.Reserve(in_bytes), "alloc bits", err)) return false;
  if (!Cuda(d_q_.Reserve(out_bytes), "alloc proj", err)) return false;
  if (!Cuda(cudaMemcpyAsync(d_bits_.p, bits, in_bytes, cudaMemcpyHostToDevice, stream_), "H2D bits", err))
    return false;
  if (!ProjectDevice(d_bits_.as<uint8_t>(), bytes_per_desc, n, d_q_.as<float>(), stream_, err))
    return false;
  if (!Cuda(cudaMemcpyAsync(out, d_q_.p, out_bytes, cudaMemcpyDeviceToHost, stream_), "D2H proj", err))
    return false;
  return Cuda(cudaStreamSynchronize(stream_), "projection", err);
}

namespace {
// descriptor -> keyframe number for every descriptor of the database, from the keyframe headers
// (one CTA per keyframe; the reference keeps this as unordered_map<int, KeypointId>,
// matching-based-engine.cc:227-238).
__global__ void expand_desc_kf_kernel(const KeyframeMeta* __restrict__ kf, int32_t* __restrict__ desc_kf) {
  const KeyframeMeta m = kf[blockIdx.x];
  for (int i = threadIdx.x; i < m.num_descriptors; i += blockDim.x)
    desc_kf[static_cast<int64_t>(m.first_descriptor) + i] = static_cast<int32_t>(blockIdx.x);
}
__global__ void owned_indices_kernel(int64_t first_owned, int64_t n, int shard_count, int32_t* __restrict__ out) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n) out[i] = static_cast<int32_t>(first_owned + i * shard_count);
}
__global__ void fill_i64_kernel(int64_t* __restrict__ p, int64_t n, int64_t v) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n) p[i] = v;
}
// Sharding by cell: the owner of a cell is a hash of its number (decorrelated from the word order).
__global__ void drop_foreign_cells_kernel(int32_t* __restrict__ cells, int64_t n, int shard_rank, int shard_count) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const int32_t c = cells[i];
  if (c >= 0 && static_cast<int>(((static_cast<uint32_t>(c) * 2654435761u) >> 8) % static_cast<uint32_t>(shard_count)) != shard_rank)
    cells[i] = -1;
}
__global__ void max_landmark_key_kernel(const int64_t* __restrict__ lm, int64_t n, unsigned long long* __restrict__ out) {
  unsigned long long local = 0;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const unsigned long long key = static_cast<unsigned long long>(lm[i] + 1);
    local = key > local ? key : local;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, local, o);
    local = other > local ? other : local;
  }
  if ((threadIdx.x & 31) == 0 && local) atomicMax(out, local);  // build-time bookkeeping, not on the query path
}
constexpr size_t kPendingFlushBytes = size_t{64} << 20;  // host staging of Insert is bounded by this
}  // namespace

// Number of descriptors of [first, first + count) that live on this shard (descriptor i on shard
// i % shard_count).
int64_t Detector::OwnedInRange(int64_t first, int64_t count) const {
  if (s_.shard_mode == 1) return count;  // by cell: every shard is handed all descriptors
  const int64_t G = s_.shard_count, r = s_.shard_rank;
  auto upto = [&](int64_t x) { return x <= r ? int64_t{0} : (x - r + G - 1) / G; };  // owned in [0, x)
  return upto(first + count) - upto(first);
}

// LoopDetector::Insert, matching-based-engine.cc:217-253: consecutive global descriptor indices,
// keyframe ids must be new (CHECK :244-252).
bool Detector::InsertBatch(const mlc_frame* frames, int64_t num_frames, const float* proj,
                           const int64_t* landmarks, bool proj_is_owned_rows, std::string* err) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  const int d = dim();
  int64_t total = 0;
  for (int64_t f = 0; f < num_frames; ++f) {
    if (frames[f].num_descriptors < 0) {
      *err = "negative descriptor count";
      return false;
    }
    total += frames[f].num_descriptors;
  }
  if (total > 0 && !proj) {
    *err = "mlc_insert: null projected descriptors";
    return false;
  }
  if (NumDescriptors() + total > 2147483647LL) {
    *err = "descriptor indices are int (SURVEY H7): database would exceed 2^31-1 descriptors";
    return false;
  }
  // Insert's CHECK: "keyframe id already in the database" — checked for the whole batch before
  // anything is changed
  {
    std::unordered_set<KeyframeKey, KeyframeKeyHash> batch;
    for (int64_t f = 0; f < num_frames; ++f) {
      const KeyframeKey key{frames[f].vertex_id, frames[f].frame_index};
      if (keyframe_keys_.count(key) || !batch.insert(key).second) {
        *err = "Insert: keyframe (vertex " + std::to_string(key.vertex) + ", frame " +
               std::to_string(key.frame_index) + ") is already in the database";
        return false;
      }
    }
  }
  const int64_t G = s_.shard_mode == 1 ? 1 : s_.shard_count, r = s_.shard_mode == 1 ? 0 : s_.shard_rank;
  const int64_t base = num_desc_;
  for (int64_t f = 0; f < num_frames; ++f) {
    KeyframeMeta m;
    m.ts = frames[f].timestamp_ns;
    m.vertex = frames[f].vertex_id;
    m.mission = frames[f].mission_id;
    m.frame_index = frames[f].frame_index;
    m.first_descriptor = static_cast<int32_t>(num_desc_);
    m.num_descriptors = frames[f].num_descriptors;
    keyframes_.push_back(m);
    keyframe_keys_.insert(KeyframeKey{m.vertex, m.frame_index});
    num_desc_ += m.num_descriptors;
  }
  if (landmarks) {
    pend_lm_.insert(pend_lm_.end(), landmarks, landmarks + total);
    for (int64_t i = 0; i < total; ++i)
      max_lm_key_ = std::max<uint64_t>(max_lm_key_, static_cast<uint64_t>(landmarks[i] + 1));
  } else {
    pend_lm_.insert(pend_lm_.end(), static_cast<size_t>(total), int64_t{-1});
  }
  // first owned global index >= base
  int64_t g = base + ((r - base % G) % G + G) % G;
  if (proj_is_owned_rows || G == 1) {
    const int64_t owned = OwnedInRange(base, total);
    pend_desc_.insert(pend_desc_.end(), proj, proj + static_cast<size_t>(owned) * d);
    for (int64_t j = 0; j < owned; ++j) pend_gidx_.push_back(static_cast<int32_t>(g + j * G));
  } else {
    for (; g < base + total; g += G) {
      const float* row = proj + static_cast<size_t>(g - base) * d;
      pend_desc_.insert(pend_desc_.end(), row, row + d);
      pend_gidx_.push_back(static_cast<int32_t>(g));
    }
  }
  index_dirty_ = true;
  if (pend_desc_.size() * 4 + pend_lm_.size() * 8 >= kPendingFlushBytes) return FlushPending(err);
  return true;
}

bool Detector::UpdateMaxLandmarkDevice(const int64_t* d_lm, int64_t n, std::string* err) {
  if (n <= 0) return true;
  if (!Cuda(d_stats_.Reserve(8), "alloc", err) || !Cuda(cudaMemsetAsync(d_stats_.p, 0, 8, stream_), "memset", err))
    return false;
  const unsigned blocks = static_cast<unsigned>(std::min<int64_t>((n + 255) / 256, 1184));
  max_landmark_key_kernel<<<blocks, 256, 0, stream_>>>(d_lm, n, d_stats_.as<unsigned long long>());
  CountLaunch();
  unsigned long long v = 0;
  if (!Cuda(cudaMemcpyAsync(&v, d_stats_.p, 8, cudaMemcpyDeviceToHost, stream_), "D2H", err) ||
      !Cuda(cudaStreamSynchronize(stream_), "max landmark", err))
    return false;
  max_lm_key_ = std::max<uint64_t>(max_lm_key_, v);
  return true;
}

// Staged host rows -> device arrays (appended).
bool Detector::FlushPending(std::string* err) {
  const size_t nl = pend_lm_.size(), no = pend_gidx_.size();
  if (nl > 0) {
    if (!Cuda(d_desc_lm_.Extend(nl * 8, stream_), "grow landmark numbers", err) ||
        !Cuda(cudaMemcpyAsync(d_desc_lm_.end(), pend_lm_.data(), nl * 8, cudaMemcpyHostToDevice, stream_),
              "H2D landmark numbers", err))
      return false;
    d_desc_lm_.used += nl * 8;
  }
  if (no > 0) {
    const size_t db = no * dim() * 4;
    if (!Cuda(d_own_desc_.Extend(db, stream_), "grow descriptors", err) ||
        !Cuda(d_own_gidx_.Extend(no * 4, stream_), "grow descriptor indices", err) ||
        !Cuda(cudaMemcpyAsync(d_own_desc_.end(), pend_desc_.data(), db, cudaMemcpyHostToDevice, stream_),
              "H2D descriptors", err) ||
        !Cuda(cudaMemcpyAsync(d_own_gidx_.end(), pend_gidx_.data(), no * 4, cudaMemcpyHostToDevice, stream_),
              "H2D descriptor indices", err))
      return false;
    d_own_desc_.used += db;
    d_own_gidx_.used += no * 4;
    num_own_ += static_cast<int64_t>(no);
  }
  if (!Cuda(cudaStreamSynchronize(stream_), "insert", err)) return false;
  pend_lm_.clear();
  pend_desc_.clear();
  pend_gidx_.clear();
  return true;
}

bool Detector::InsertBatchDevice(const mlc_frame* frames, int64_t num_frames, const float* d_proj_owned,
                                 int64_t num_owned, const int64_t* d_landmarks, cudaStream_t stream,
                                 std::string* err) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  int64_t total = 0;
  for (int64_t f = 0; f < num_frames; ++f) {
    if (frames[f].num_descriptors < 0) {
      *err = "negative descriptor count";
      return false;
    }
    total += frames[f].num_descriptors;
  }
  if (NumDescriptors() + total > 2147483647LL) {
    *err = "descriptor indices are int (SURVEY H7): database would exceed 2^31-1 descriptors";
    return false;
  }
  const int64_t base = num_desc_, G = s_.shard_mode == 1 ? 1 : s_.shard_count,
                r = s_.shard_mode == 1 ? 0 : s_.shard_rank;
  const int64_t owned = OwnedInRange(base, total);
  if (num_owned != owned) {
    *err = "mlc_insert_batch_device: this shard owns " + std::to_string(owned) + " of the batch's " +
           std::to_string(total) + " descriptors, got " + std::to_string(num_owned) + " rows";
    return false;
  }
  if (owned > 0 && !d_proj_owned) {
    *err = "mlc_insert_batch_device: null projected descriptors";
    return false;
  }
  {
    std::unordered_set<KeyframeKey, KeyframeKeyHash> batch;
    for (int64_t f = 0; f < num_frames; ++f) {
      const KeyframeKey key{frames[f].vertex_id, frames[f].frame_index};
      if (keyframe_keys_.count(key) || !batch.insert(key).second) {
        *err = "Insert: keyframe (vertex " + std::to_string(key.vertex) + ", frame " +
               std::to_string(key.frame_index) + ") is already in the database";
        return false;
      }
    }
  }
  if (!FlushPending(err)) return false;  // keep the device arrays in insertion order
  if (!Cuda(cudaStreamSynchronize(stream), "caller stream", err)) return false;
  const size_t db = static_cast<size_t>(owned) * dim() * 4;
  if (!Cuda(d_desc_lm_.Extend(static_cast<size_t>(total) * 8, stream_), "grow landmark numbers", err) ||
      !Cuda(d_own_desc_.Extend(db, stream_), "grow descriptors", err) ||
      !Cuda(d_own_gidx_.Extend(static_cast<size_t>(owned) * 4, stream_), "grow descriptor indices", err))
    return false;
  if (total > 0) {
    if (d_landmarks) {
      if (!Cuda(cudaMemcpyAsync(d_desc_lm_.end(), d_landmarks, static_cast<size_t>(total) * 8,
                                cudaMemcpyDeviceToDevice, stream_), "copy landmark numbers", err))
        return false;
    } else {
      fill_i64_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream_>>>(
          reinterpret_cast<int64_t*>(d_desc_lm_.end()), total, -1);
      CountLaunch();
    }
  }
  if (owned > 0) {
    const int64_t g = base + ((r - base % G) % G + G) % G;
    if (!Cuda(cudaMemcpyAsync(d_own_desc_.end(), d_proj_owned, db, cudaMemcpyDeviceToDevice, stream_),
              "copy descriptors", err))
      return false;
    owned_indices_kernel<<<static_cast<unsigned>((owned + 255) / 256), 256, 0, stream_>>>(
        g, owned, static_cast<int>(G), reinterpret_cast<int32_t*>(d_own_gidx_.end()));
    CountLaunch();
  }
  if (!Cuda(cudaStreamSynchronize(stream_), "insert", err)) return false;
  if (d_landmarks && !UpdateMaxLandmarkDevice(reinterpret_cast<const int64_t*>(d_desc_lm_.end()), total, err)) return false;
  d_desc_lm_.used += static_cast<size_t>(total) * 8;
  d_own_desc_.used += db;
  d_own_gidx_.used += static_cast<size_t>(owned) * 4;
  num_own_ += owned;
  for (int64_t f = 0; f < num_frames; ++f) {
    KeyframeMeta m;
    m.ts = frames[f].timestamp_ns;
    m.vertex = frames[f].vertex_id;
    m.mission = frames[f].mission_id;
    m.frame_index = frames[f].frame_index;
    m.first_descriptor = static_cast<int32_t>(num_desc_);
    m.num_descriptors = frames[f].num_descriptors;
    keyframes_.push_back(m);
    keyframe_keys_.insert(KeyframeKey{m.vertex, m.frame_index});
    num_desc_ += m.num_descriptors;
  }
  index_dirty_ = true;
  return true;
}

bool Detector::Initialize(std::string* err) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  return EnsureIndex(err);
}

// Build the device index from the device-resident database: cell of every owned descriptor =
// FindClosestWords(desc, 1) (kernel 2a with one word), then cell-sorted inverted lists; the
// descriptor -> keyframe replica is regenerated from the keyframe headers.
bool Detector::EnsureIndex(std::string* err) {
  if (!index_dirty_) return true;
  if (!FlushPending(err)) return false;
  const int64_t n = NumDescriptors(), no = num_own_;
  const int d = dim();
  const uint64_t cells64 = static_cast<uint64_t>(vocab_.words1.cols) * vocab_.words2.cols;
  if (cells64 > (1ull << 28)) {
    *err = "too many cells for the dense cell table";
    return false;
  }
  const float* d_desc = d_own_desc_.as<float>();
  const int32_t* d_gidx = d_own_gidx_.as<int32_t>();
  if (s_.engine == 2) {  // exhaustive search: the descriptor rows are the index
    if (!UploadKeyframeReplicas(err)) return false;
    index_dirty_ = false;
    return true;
  }
  if (no > 0) {
    if (!Cuda(d_db_cells_.Reserve(static_cast<size_t>(no) * 4), "alloc cells", err)) return false;
    if (!CoarseChunks(d_desc, no, 1, d_db_cells_.as<int32_t>(), stream_, err)) return false;
    if (s_.shard_mode == 1 && s_.shard_count > 1) {  // by cell: keep the cells this shard owns
      drop_foreign_cells_kernel<<<static_cast<unsigned>((no + 255) / 256), 256, 0, stream_>>>(
          d_db_cells_.as<int32_t>(), no, s_.shard_rank, s_.shard_count);
      CountLaunch();
    }
  }
  bool ok = true;
  if (s_.engine == 1) {
    // imipq: entries carry the quantised residual (12 code bytes) instead of the coordinates
    uint32_t* d_codes = nullptr;
    if (no > 0) {
      ok = Cuda(cudaMalloc(&d_codes, static_cast<size_t>(no) * 12), "alloc pq codes", err) &&
           Cuda(LaunchPqEncode(pq_, d_desc, d_db_cells_.as<int32_t>(), no, d_codes, stream_), "pq encode", err);
    }
    ok = ok && Cuda(BuildImiLists(d_db_cells_.as<int32_t>(), reinterpret_cast<const float*>(d_codes), d_gidx, no,
                                  3, static_cast<uint32_t>(cells64), &lists_, stream_),
                    "build inverted lists", err);
    if (d_codes) cudaFree(d_codes);
  } else {
    ok = Cuda(BuildImiLists(d_db_cells_.as<int32_t>(), d_desc, d_gidx, no, d, static_cast<uint32_t>(cells64),
                            &lists_, stream_),
              "build inverted lists", err);
  }
  if (!ok) return false;
  if (!UploadKeyframeReplicas(err)) return false;
  (void)n;
  index_dirty_ = false;
  return true;
}

// Metadata replicas for voting / clustering (kernel 3): keyframe headers and descriptor -> keyframe.
bool Detector::UploadKeyframeReplicas(std::string* err) {
  const int64_t n = NumDescriptors();
  if (n > 0) {
    if (!Cuda(d_desc_kf_.Reserve(static_cast<size_t>(n) * 4), "alloc", err)) return false;
    if (!Cuda(d_kf_meta_.Reserve(keyframes_.size() * sizeof(KeyframeMeta)), "alloc", err)) return false;
    if (!Cuda(cudaMemcpyAsync(d_kf_meta_.p, keyframes_.data(), keyframes_.size() * sizeof(KeyframeMeta),
                              cudaMemcpyHostToDevice, stream_), "H2D", err)) return false;
    expand_desc_kf_kernel<<<static_cast<unsigned>(keyframes_.size()), 128, 0, stream_>>>(
        d_kf_meta_.as<KeyframeMeta>(), d_desc_kf_.as<int32_t>());
    CountLaunch();
    if (!Cuda(cudaGetLastError(), "expand descriptor -> keyframe", err)) return false;
  }
  return Cuda(cudaStreamSynchronize(stream_), "index build", err);
}

bool Detector::KnnDevice(const float* d_q, int64_t n_q, int k, int32_t* d_idx, float* d_dist,
                         cudaStream_t stream, std::string* err) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  if (k <= 0 || k > 16) {
    *err = "k must be in 1..16";
    return false;
  }
  if (!EnsureIndex(err)) return false;
  if (n_q == 0) return true;
  if (s_.engine == 2) {
    // CHECK_LT(num_neighbors, ef_query_) and CHECK_EQ(result.size(), num_neighbors), hnsw-index-interface.h:128, :143
    if (k >= s_.hnsw_ef_query || k > NumDescriptors()) {
      *err = "hnsw engine: num_neighbors must be < hnsw_ef_query and <= the number of descriptors in the index";
      return false;
    }
    const int splits = ExactKnnSplits(n_q, num_own_, sm_count_);
    if (!Cuda(d_coarse_scratch_.Reserve(ExactKnnScratchBytes(n_q, k, splits)), "alloc knn scratch", err)) return false;
    last_valid_ = false;
    return Cuda(LaunchExactKnn(d_own_desc_.as<float>(), num_own_, d_q, n_q, dim(), k, splits, d_coarse_scratch_.p, d_idx,
                               d_dist, stream), "exact kNN", err);
  }
  const int nw = s_.num_closest_words;
  if (!Cuda(d_cells_.Reserve(static_cast<size_t>(n_q) * nw * 4), "alloc visit list", err)) return false;
  if (!CoarseChunks(d_q, n_q, nw, d_cells_.as<int32_t>(), stream, err)) return false;
  cudaEventRecord(ev0_, stream);
  if (!Cuda(LaunchScan(d_q, n_q, d_cells_.as<int32_t>(), nw, k, d_idx, d_dist, stream), "list scan", err))
    return false;
  cudaEventRecord(ev1_, stream);
  last_cells_ = d_cells_.as<int32_t>();
  last_scan_launches_ = 0;
  last_nq_ = n_q;
  last_nw_ = nw;
  last_valid_ = true;
  return true;
}

bool Detector::Knn(const float* q, int64_t n_q, int k, int32_t* idx, float* dist, std::string* err) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  if (n_q < 0) {
    *err = "negative query count";
    return false;
  }
  if (n_q == 0) return true;
  const size_t qb = static_cast<size_t>(n_q) * dim() * 4, rb = static_cast<size_t>(n_q) * k * 4;
  if (!Cuda(d_q_.Reserve(qb), "alloc", err) || !Cuda(d_idx_.Reserve(rb), "alloc", err) ||
      !Cuda(d_dist_.Reserve(rb), "alloc", err))
    return false;
  if (!Cuda(cudaMemcpyAsync(d_q_.p, q, qb, cudaMemcpyHostToDevice, stream_), "H2D queries", err))
    return false;
  if (!KnnDevice(d_q_.as<float>(), n_q, k, d_idx_.as<int32_t>(), d_dist_.as<float>(), stream_, err))
    return false;
  if (!Cuda(cudaMemcpyAsync(idx, d_idx_.p, rb, cudaMemcpyDeviceToHost, stream_), "D2H", err)) return false;
  if (!Cuda(cudaMemcpyAsync(dist, d_dist_.p, rb, cudaMemcpyDeviceToHost, stream_), "D2H", err)) return false;
  return Cuda(cudaStreamSynchronize(stream_), "knn", err);
}

bool Detector::CoarseCells(const float* q, int64_t n, int nw, int32_t* cells, std::string* err) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  if (nw <= 0 || nw > 16 || n < 0) {
    *err = "bad arguments";
    return false;
  }
  if (s_.engine == 2) {
    *err = "the hnsw engine has no coarse quantiser";
    return false;
  }
  if (n == 0) return true;
  const size_t qb = static_cast<size_t>(n) * dim() * 4, cb = static_cast<size_t>(n) * nw * 4;
  if (!Cuda(d_q_.Reserve(qb), "alloc", err) || !Cuda(d_cells_.Reserve(cb), "alloc", err)) return false;
  if (!Cuda(cudaMemcpyAsync(d_q_.p, q, qb, cudaMemcpyHostToDevice, stream_), "H2D", err)) return false;
  if (!CoarseChunks(d_q_.as<float>(), n, nw, d_cells_.as<int32_t>(), stream_, err)) return false;
  if (!Cuda(cudaMemcpyAsync(cells, d_cells_.p, cb, cudaMemcpyDeviceToHost, stream_), "D2H", err)) return false;
  last_valid_ = false;
  return Cuda(cudaStreamSynchronize(stream_), "coarse cells", err);
}

bool Detector::CoarseDevice(const float* d_q, int64_t n, int nw, int32_t* d_cells, cudaStream_t stream,
                            std::string* err) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  if (nw <= 0 || nw > 16) {
    *err = "nw must be in 1..16";
    return false;
  }
  if (s_.engine == 2) {
    *err = "the hnsw engine has no coarse quantiser";
    return false;
  }
  return CoarseChunks(d_q, n, nw, d_cells, stream, err);
}

// Kernel 2a in chunks of <= 4 M descriptors so that the per-half word-list scratch stays bounded.
bool Detector::CoarseChunks(const float* d_q, int64_t n, int nw, int32_t* d_cells, cudaStream_t stream,
                            std::string* err) {
  const int64_t kChunk = 4 << 20;
  const int64_t first = std::min<int64_t>(n, kChunk);
  if (!Cuda(d_coarse_scratch_.Reserve(CoarseScratchBytes(coarse_, first, nw)), "alloc coarse scratch", err))
    return false;
  for (int64_t s = 0; s < n; s += kChunk) {
    const int64_t c = std::min<int64_t>(kChunk, n - s);
    if (!Cuda(LaunchCoarseWords(coarse_, d_q + s * dim(), c, nw, d_cells + s * nw, d_coarse_scratch_.p,
                                sm_count_, stream), "coarse word search", err))
      return false;
  }
  return true;
}

bool Detector::ScanDevice(const float* d_q, const int32_t* d_cells, int64_t n_q, int k, int32_t* d_idx,
                          float* d_dist, cudaStream_t stream, std::string* err) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  if (k <= 0 || k > 16) {
    *err = "k must be in 1..16";
    return false;
  }
  if (s_.engine == 2) {
    *err = "the hnsw engine has no inverted lists";
    return false;
  }
  if (!EnsureIndex(err)) return false;
  if (n_q == 0) return true;
  const int nw = s_.num_closest_words;
  cudaEventRecord(ev0_, stream);
  if (!Cuda(LaunchScan(d_q, n_q, d_cells, nw, k, d_idx, d_dist, stream), "list scan", err))
    return false;
  cudaEventRecord(ev1_, stream);
  last_cells_ = d_cells;  // mlc_last_scan_stats reads the caller's visit list: keep it until then
  last_scan_launches_ = 0;
  last_nq_ = n_q;
  last_nw_ = nw;
  last_valid_ = true;
  return true;
}

bool Detector::LastStageMs(double* ms5, std::string* err) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  if (!stage_valid_) {
    *err = "no fused query to report on";
    return false;
  }
  cudaEventSynchronize(ev_stage_[5]);
  for (int i = 0; i < 5; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ev_stage_[i], ev_stage_[i + 1]) != cudaSuccess) ms = 0.f;
    ms5[i] = ms;
  }
  return true;
}

bool Detector::MergeTopkDevice(const int32_t* d_idx_lists, const float* d_dist_lists, int num_lists,
                               int64_t n_q, int k, int32_t* d_idx, float* d_dist, cudaStream_t stream,
                               std::string* err) {
  return Cuda(LaunchMergeTopk(d_idx_lists, d_dist_lists, num_lists, n_q, k, d_idx, d_dist, stream),
              "top-k merge", err);
}

bool Detector::LastScanStats(uint64_t* bytes, uint64_t* entries, double* ms, std::string* err) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  if (!last_valid_) {
    *err = "no kNN call to report on";
    return false;
  }
  if (!Cuda(d_stats_.Reserve(8), "alloc", err)) return false;
  if (!Cuda(cudaMemsetAsync(d_stats_.p, 0, 8, stream_), "memset", err)) return false;
  if (!Cuda(LaunchScanEntries(last_cells_, last_nq_ * last_nw_, lists_.cell_info,
                              d_stats_.as<unsigned long long>(), stream_),
            "scan stats", err))
    return false;
  unsigned long long total = 0;
  if (!Cuda(cudaMemcpyAsync(&total, d_stats_.p, 8, cudaMemcpyDeviceToHost, stream_), "D2H", err)) return false;
  if (!Cuda(cudaStreamSynchronize(stream_), "scan stats", err)) return false;
  float msf = 0.f;
  if (last_scan_launches_ > 0) {  // sharded step: one launch per source rank's block
    for (int s = 0; s < last_scan_launches_; ++s) {
      float part = 0.f;
      cudaEventSynchronize(ev_scan_[2 * s + 1]);
      if (cudaEventElapsedTime(&part, ev_scan_[2 * s], ev_scan_[2 * s + 1]) == cudaSuccess) msf += part;
    }
  } else {
    cudaEventSynchronize(ev1_);
    if (cudaEventElapsedTime(&msf, ev0_, ev1_) != cudaSuccess) msf = 0.f;
  }
  *entries = total;
  if (s_.engine == 1) {
    // algorithmic entry = index + packed codes (9 B at 10 components x 16 centres)
    int bits = 1;
    while ((1 << bits) < pq_.num_centers) ++bits;
    *bytes = total * static_cast<uint64_t>(4 + (2 * pq_.half_ncomp * bits + 7) / 8);
  } else {
    *bytes = total * static_cast<uint64_t>(4 * (dim() + 1));
  }
  *ms = msf;
  return true;
}

// ---------------------------------------------------------------------------------------------
// Database persistence. The reference rebuilds its database for every `lc` / `aam` / `relax`
// invocation (loop-detector-node.cc:273-339, vi-map-merger.cc:71-78); here the built index is one
// little-endian file: header, keyframes, projected descriptors, landmark numbers, descriptor ->
// keyframe, cell per descriptor, cell table, inverted lists, landmark positions.
// ---------------------------------------------------------------------------------------------
namespace {
struct IndexFileHeader {
  char magic[8];  // "MLCIDX02"
  uint64_t vocab_hash;
  int32_t engine, dim, shard_rank, shard_count;
  uint32_t num_cells;
  int32_t list_dim;
  int64_t num_descriptors, num_owned, num_keyframes, num_landmark_xyz;
  uint64_t list_bytes;
};
bool WriteAll(FILE* f, const void* p, size_t bytes) { return bytes == 0 || fwrite(p, 1, bytes, f) == bytes; }
bool ReadAll(FILE* f, void* p, size_t bytes) { return bytes == 0 || fread(p, 1, bytes, f) == bytes; }

// Every stored descriptor index of the inverted lists must address the metadata replicas.
__global__ void validate_lists_kernel(const uint2* __restrict__ cell_info, uint32_t num_cells,
                                      const uint32_t* __restrict__ lists, int words_per_entry, int index_word,
                                      uint32_t num_descriptors, int* __restrict__ bad) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= num_cells) return;
  const uint2 ci = cell_info[c];
  const uint32_t* e = lists + (static_cast<size_t>(ci.x) << 2);
  for (uint32_t i = 0; i < ci.y; ++i)
    if (e[static_cast<size_t>(i) * words_per_entry + index_word] >= num_descriptors) *bad = 1;
}
}  // namespace

bool Detector::SaveIndex(const char* path, std::string* err) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  if (s_.engine == 2) {
    *err = "index files hold inverted lists: not available for the hnsw engine";
    return false;
  }
  if (s_.shard_mode != 0) {
    *err = "index files are written for shard_mode 0 only";
    return false;
  }
  if (!EnsureIndex(err)) return false;
  FILE* f = fopen(path, "wb");
  if (!f) {
    *err = std::string("cannot open ") + path + " for writing";
    return false;
  }
  IndexFileHeader h{};
  std::memcpy(h.magic, "MLCIDX02", 8);
  h.vocab_hash = vocab_hash_;
  h.engine = s_.engine;
  h.dim = dim();
  h.shard_rank = s_.shard_rank;
  h.shard_count = s_.shard_count;
  h.num_cells = lists_.num_cells;
  h.list_dim = lists_.dim;
  h.num_descriptors = NumDescriptors();
  h.num_owned = num_own_;
  h.num_keyframes = static_cast<int64_t>(keyframes_.size());
  h.num_landmark_xyz = num_landmark_xyz_;
  h.list_bytes = lists_.list_bytes;
  const size_t n = static_cast<size_t>(h.num_descriptors), no = static_cast<size_t>(h.num_owned);
  bool ok = WriteAll(f, &h, sizeof(h)) && WriteAll(f, keyframes_.data(), keyframes_.size() * sizeof(KeyframeMeta));
  // device-resident parts through a bounded staging buffer
  std::vector<unsigned char> stage(size_t{64} << 20);
  auto dump = [&](const void* dptr, size_t bytes) {
    for (size_t at = 0; ok && at < bytes; at += stage.size()) {
      const size_t c = std::min(stage.size(), bytes - at);
      ok = Cuda(cudaMemcpy(stage.data(), static_cast<const unsigned char*>(dptr) + at, c, cudaMemcpyDeviceToHost),
                "D2H index", err) && WriteAll(f, stage.data(), c);
    }
  };
  if (ok) dump(d_desc_lm_.p, n * 8);
  if (ok) dump(d_own_gidx_.p, no * 4);
  if (ok) dump(d_own_desc_.p, no * dim() * 4);
  if (ok) dump(d_db_cells_.p, no * 4);
  if (ok) dump(lists_.cell_info, static_cast<size_t>(lists_.num_cells) * sizeof(uint2));
  if (ok) dump(lists_.lists, lists_.list_bytes);
  if (ok) dump(d_landmark_xyz_.p, static_cast<size_t>(num_landmark_xyz_) * 24);
  ok = (fclose(f) == 0) && ok;
  if (!ok && err->empty()) *err = std::string("short write to ") + path;
  return ok;
}

bool Detector::LoadIndex(const char* path, std::string* err) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  FILE* f = fopen(path, "rb");
  if (!f) {
    *err = std::string("cannot open ") + path;
    return false;
  }
  struct Closer {
    FILE* f;
    ~Closer() { fclose(f); }
  } closer{f};
  if (s_.shard_mode != 0 || s_.engine == 2) {
    *err = "index files are read for shard_mode 0 and the imi / imipq engines only";
    return false;
  }
  IndexFileHeader h{};
  if (!ReadAll(f, &h, sizeof(h)) || std::memcmp(h.magic, "MLCIDX02", 8) != 0) {
    *err = "not an index file of this library";
    return false;
  }
  const uint64_t cells64 = static_cast<uint64_t>(vocab_.words1.cols) * vocab_.words2.cols;
  if (h.vocab_hash != vocab_hash_ || h.engine != s_.engine || h.dim != dim() || h.num_cells != cells64 ||
      h.shard_rank != s_.shard_rank || h.shard_count != s_.shard_count) {
    *err = "index file was built with another vocabulary / engine / sharding";
    return false;
  }
  // ---- the header must describe exactly this file before anything is allocated from it ----
  const int list_dim = s_.engine == 1 ? 3 : dim();
  const uint64_t wpe = static_cast<uint64_t>((list_dim + 1 + 3) & ~3);
  if (h.list_dim != list_dim || h.num_descriptors < 0 || h.num_descriptors > 2147483647LL || h.num_keyframes < 0 ||
      h.num_keyframes > h.num_descriptors + (int64_t{1} << 24) || h.num_landmark_xyz < 0 ||
      h.num_landmark_xyz > (int64_t{1} << 40) || h.num_owned != OwnedInRange(0, h.num_descriptors) ||
      (h.list_bytes & 15) != 0 || h.list_bytes > (uint64_t{1} << 36)) {
    *err = "corrupt index file header";
    return false;
  }
  const uint64_t n = static_cast<uint64_t>(h.num_descriptors), no = static_cast<uint64_t>(h.num_owned);
  const uint64_t expect = sizeof(h) + static_cast<uint64_t>(h.num_keyframes) * sizeof(KeyframeMeta) + n * 8 + no * 4 +
                          no * h.dim * 4 + no * 4 + static_cast<uint64_t>(h.num_cells) * sizeof(uint2) + h.list_bytes +
                          static_cast<uint64_t>(h.num_landmark_xyz) * 24;
  if (fseek(f, 0, SEEK_END) != 0 || static_cast<uint64_t>(ftell(f)) != expect ||
      fseek(f, static_cast<long>(sizeof(h)), SEEK_SET) != 0) {
    *err = std::string("truncated or oversized index file ") + path;
    return false;
  }
  std::vector<KeyframeMeta> kfs;
  std::vector<uint2> cell_info;
  std::vector<int32_t> gidx;
  try {
    kfs.resize(static_cast<size_t>(h.num_keyframes));
    cell_info.resize(h.num_cells);
    gidx.resize(no);
  } catch (const std::exception&) {
    *err = "out of memory reading the index file";
    return false;
  }
  if (!ReadAll(f, kfs.data(), kfs.size() * sizeof(KeyframeMeta))) {
    *err = std::string("truncated index file ") + path;
    return false;
  }
  std::unordered_set<KeyframeKey, KeyframeKeyHash> keys;
  {
    int64_t at = 0;
    for (const KeyframeMeta& m : kfs) {
      if (m.num_descriptors < 0 || m.first_descriptor != at || !keys.insert(KeyframeKey{m.vertex, m.frame_index}).second) {
        *err = "corrupt index file: keyframe table";
        return false;
      }
      at += m.num_descriptors;
    }
    if (at != h.num_descriptors) {
      *err = "corrupt index file: keyframe table does not cover the descriptors";
      return false;
    }
  }
  // ---- device side: everything is uploaded and checked before the current database is replaced ----
  GrowBuf lm, own_desc, own_gidx;
  DevBuf cells, xyz, flag;
  DeviceLists lists;
  lists.num_cells = h.num_cells;
  lists.dim = h.list_dim;
  lists.list_bytes = h.list_bytes;
  auto drop = [&]() {
    lm.Free();
    own_desc.Free();
    own_gidx.Free();
    cells.Free();
    xyz.Free();
    flag.Free();
    lists.Free();
  };
  bool ok = true;
  std::vector<unsigned char> stage;
  try {
    stage.resize(size_t{64} << 20);
  } catch (const std::exception&) {
    *err = "out of memory reading the index file";
    return false;
  }
  auto fill = [&](void* dptr, size_t bytes, void* host_copy) {
    for (size_t at = 0; ok && at < bytes; at += stage.size()) {
      const size_t c = std::min(stage.size(), bytes - at);
      ok = ReadAll(f, stage.data(), c) &&
           Cuda(cudaMemcpy(static_cast<unsigned char*>(dptr) + at, stage.data(), c, cudaMemcpyHostToDevice),
                "H2D index", err);
      if (ok && host_copy) std::memcpy(static_cast<unsigned char*>(host_copy) + at, stage.data(), c);
    }
  };
  ok = Cuda(lm.Extend(n * 8 + 16, stream_), "alloc", err) && Cuda(own_gidx.Extend(no * 4 + 16, stream_), "alloc", err) &&
       Cuda(own_desc.Extend(no * h.dim * 4 + 16, stream_), "alloc", err) && Cuda(cells.Reserve(no * 4 + 16), "alloc", err) &&
       Cuda(cudaMalloc(&lists.cell_info, sizeof(uint2) * static_cast<size_t>(h.num_cells)), "alloc", err) &&
       Cuda(cudaMalloc(&lists.lists, h.list_bytes + 16), "alloc", err) &&
       Cuda(xyz.Reserve(static_cast<size_t>(h.num_landmark_xyz) * 24 + 16), "alloc", err) &&
       Cuda(flag.Reserve(16), "alloc", err);
  if (ok) fill(lm.p, n * 8, nullptr);
  if (ok) fill(own_gidx.p, no * 4, gidx.data());
  if (ok) fill(own_desc.p, no * h.dim * 4, nullptr);
  if (ok) fill(cells.p, no * 4, nullptr);
  if (ok) fill(lists.cell_info, static_cast<size_t>(h.num_cells) * sizeof(uint2), cell_info.data());
  if (ok) fill(lists.lists, h.list_bytes, nullptr);
  if (ok) fill(xyz.p, static_cast<size_t>(h.num_landmark_xyz) * 24, nullptr);
  if (!ok) {
    drop();
    if (err->empty()) *err = std::string("truncated index file ") + path;
    return false;
  }
  // owned rows: ascending global indices of this shard
  for (uint64_t j = 0; j < no; ++j) {
    if (gidx[j] != static_cast<int64_t>(s_.shard_rank) + static_cast<int64_t>(j) * s_.shard_count) {
      drop();
      *err = "corrupt index file: descriptor indices";
      return false;
    }
  }
  // cell table: every list inside the list block, no more entries than owned rows
  {
    uint64_t entries = 0;
    for (const uint2& ci : cell_info) {
      entries += ci.y;
      if ((static_cast<uint64_t>(ci.x) << 4) + static_cast<uint64_t>(ci.y) * wpe * 4 > h.list_bytes) {
        drop();
        *err = "corrupt index file: cell table points outside the inverted lists";
        return false;
      }
    }
    if (entries > no) {
      drop();
      *err = "corrupt index file: cell table holds more entries than descriptors";
      return false;
    }
  }
  if (h.num_cells > 0) {
    int bad = 0;
    ok = Cuda(cudaMemsetAsync(flag.p, 0, 4, stream_), "memset", err);
    if (ok) {
      validate_lists_kernel<<<(h.num_cells + 127) / 128, 128, 0, stream_>>>(
          lists.cell_info, h.num_cells, lists.lists, static_cast<int>(wpe), list_dim,
          static_cast<uint32_t>(h.num_descriptors), flag.as<int>());
      CountLaunch();
      ok = Cuda(cudaMemcpyAsync(&bad, flag.p, 4, cudaMemcpyDeviceToHost, stream_), "D2H", err) &&
           Cuda(cudaStreamSynchronize(stream_), "validate index", err);
    }
    if (!ok || bad) {
      drop();
      if (err->empty()) *err = "corrupt index file: inverted lists name descriptors outside the database";
      return false;
    }
  }
  flag.Free();
  // ---- commit: replace the current database ----
  lm.used = n * 8;
  own_gidx.used = no * 4;
  own_desc.used = no * h.dim * 4;
  keyframes_.swap(kfs);
  keyframe_keys_.swap(keys);
  num_desc_ = h.num_descriptors;
  num_own_ = h.num_owned;
  pend_desc_.clear();
  pend_gidx_.clear();
  pend_lm_.clear();
  d_desc_lm_.Free();
  d_desc_lm_ = lm;
  d_own_desc_.Free();
  d_own_desc_ = own_desc;
  d_own_gidx_.Free();
  d_own_gidx_ = own_gidx;
  lists_.Free();
  lists_ = lists;
  d_db_cells_.Free();
  d_db_cells_ = cells;
  d_landmark_xyz_.Free();
  d_landmark_xyz_ = xyz;
  num_landmark_xyz_ = h.num_landmark_xyz;
  last_valid_ = false;
  index_dirty_ = true;  // until the replicas below are in place
  max_lm_key_ = 0;
  if (!UpdateMaxLandmarkDevice(d_desc_lm_.as<int64_t>(), num_desc_, err)) return false;
  if (!UploadKeyframeReplicas(err)) return false;
  index_dirty_ = false;
  return true;
}

}  // namespace mlc
