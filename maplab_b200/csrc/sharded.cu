// Multi-GPU loop-closure query step inside the library: one process per GPU, the
// inverted lists sharded over the ranks, everything else replicated. The reference's fan-out over
// query vertices is C++ (loop-closure-handler/src/loop-detector-node.cc:819-873,
// maplab-common/include/maplab-common/parallel-process.h:49-92); this is its multi-GPU counterpart
// behind the same C-ABI handle, so a C++ host gets it without any Python.
//
// Per step, on every rank r (G ranks):
//   1. project + coarse-search ITS slice of the query keyframes                      (kernels 1, 2a)
//   2. exchange 1: one grouped NCCL all-gather of (projected query, visit list) on the comm stream
//   3. scan its OWN slice against its shard while exchange 1 is in flight, then for s = 1..G-1 the
//      block of source rank (r+s)%G; as soon as block s is scanned its per-shard top-k lists leave
//      for their owner on the comm stream (one grouped send/recv per step: send to (r+s)%G, receive
//      from (r-s)%G) while the next block is scanned                                  (kernel 2b)
//   4. merge the G per-shard lists of the slice by (distance, index) — equal to the single-index
//      result because the visited cells depend only on query + vocabulary and global descriptor
//      indices keep the tie-breaks global
//   5. voting / clustering / RANSAC on the slice                                      (kernels 3, 4)
// Nothing else crosses GPUs.
//
// NCCL is bound at run time (dlopen of libnccl.so.2 when the communicator is created): the library
// itself loads on single-GPU machines without NCCL, and inside a process that already carries an NCCL
// (a torch.distributed process) the same copy is used instead of a second one.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "detector.h"

namespace mlc {

struct NcclApi {
  void* handle = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetVersion) GetVersion = nullptr;
};

namespace {
NcclApi g_nccl;
std::mutex g_nccl_mu;

bool LoadNccl(std::string* err) {
  std::lock_guard<std::mutex> lock(g_nccl_mu);
  if (g_nccl.handle) return true;
  void* h = nullptr;
  for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
    h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) {
    *err = std::string("multi-GPU queries need NCCL: ") + dlerror();
    return false;
  }
  NcclApi a;
  a.handle = h;
#define MLC_SYM(field, sym)                                             \
  a.field = reinterpret_cast<decltype(a.field)>(dlsym(h, #sym));        \
  if (!a.field) {                                                       \
    *err = "libnccl lacks " #sym;                                       \
    return false;                                                       \
  }
  MLC_SYM(GetUniqueId, ncclGetUniqueId)
  MLC_SYM(CommInitRank, ncclCommInitRank)
  MLC_SYM(CommDestroy, ncclCommDestroy)
  MLC_SYM(GetErrorString, ncclGetErrorString)
  MLC_SYM(AllGather, ncclAllGather)
  MLC_SYM(Send, ncclSend)
  MLC_SYM(Recv, ncclRecv)
  MLC_SYM(GroupStart, ncclGroupStart)
  MLC_SYM(GroupEnd, ncclGroupEnd)
  MLC_SYM(GetVersion, ncclGetVersion)
#undef MLC_SYM
  g_nccl = a;
  return true;
}
}  // namespace

bool CommUniqueId(void* id128, std::string* err) {
  static_assert(sizeof(ncclUniqueId) == 128, "mlc_comm_unique_id hands out 128 bytes");
  if (!LoadNccl(err)) return false;
  ncclUniqueId id;
  const ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r != ncclSuccess) {
    *err = std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r);
    return false;
  }
  std::memcpy(id128, &id, sizeof(id));
  return true;
}

bool Detector::Nccl(int result, const char* what, std::string* err) const {
  if (result == ncclSuccess) return true;
  *err = std::string(what) + ": " + g_nccl.GetErrorString(static_cast<ncclResult_t>(result));
  return false;
}

bool Detector::CommInit(const void* id128, std::string* err) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  if (comm_) {
    *err = "communicator already initialised";
    return false;
  }
  if (!LoadNccl(err)) return false;
  if (!Cuda(cudaSetDevice(device_), "cudaSetDevice", err)) return false;
  ncclUniqueId id;
  std::memcpy(&id, id128, sizeof(id));
  ncclComm_t comm = nullptr;
  if (!Nccl(g_nccl.CommInitRank(&comm, s_.shard_count, id, s_.shard_rank), "ncclCommInitRank", err)) return false;
  comm_ = comm;
  if (!comm_stream_ &&
      !Cuda(cudaStreamCreateWithFlags(&comm_stream_, cudaStreamNonBlocking), "cudaStreamCreate", err))
    return false;
  for (cudaEvent_t& e : ev_comm_)
    if (!e && !Cuda(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate", err)) return false;
  for (cudaEvent_t& e : ev_scan_)
    if (!e && !Cuda(cudaEventCreate(&e), "cudaEventCreate", err)) return false;
  return true;
}

void Detector::CommDestroy() {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  if (comm_) {
    cudaStreamSynchronize(comm_stream_);
    g_nccl.CommDestroy(static_cast<ncclComm_t>(comm_));
    comm_ = nullptr;
  }
}

int Detector::CommVersion() const {
  int v = 0;
  if (g_nccl.handle) g_nccl.GetVersion(&v);
  return v;
}

// Slice sizes of all ranks: ragged slices are padded to the largest (collective).
bool Detector::ShardedSliceSizes(int64_t n_s, int64_t* n_max, std::string* err) {
  ncclComm_t comm = static_cast<ncclComm_t>(comm_);
  const int G = s_.shard_count;
  if (!Cuda(sh_counts_.Reserve(sizeof(long long) * (G + 1)), "alloc", err)) return false;
  long long* d_counts = sh_counts_.as<long long>();
  const long long mine = n_s;
  std::vector<long long> counts(G);
  if (!Cuda(cudaMemcpyAsync(d_counts + G, &mine, sizeof(long long), cudaMemcpyHostToDevice, stream_), "H2D", err) ||
      !Nccl(g_nccl.AllGather(d_counts + G, d_counts, sizeof(long long), ncclChar, comm, stream_), "ncclAllGather", err) ||
      !Cuda(cudaMemcpyAsync(counts.data(), d_counts, sizeof(long long) * G, cudaMemcpyDeviceToHost, stream_), "D2H", err) ||
      !Cuda(cudaStreamSynchronize(stream_), "slice sizes", err))
    return false;
  *n_max = *std::max_element(counts.begin(), counts.end());
  return true;
}

// kNN of this rank's query slice against the WHOLE (sharded) database: steps 1-4 above. Leaves the
// merged lists of the slice in d_idx_ / d_dist_ (n_s x k). d_q_ (room for n_max rows) must hold the
// projected slice.
bool Detector::ShardedKnnOnSlice(int64_t n_s, int64_t n_max, int k, std::string* err) {
  ncclComm_t comm = static_cast<ncclComm_t>(comm_);
  const int G = s_.shard_count, r = s_.shard_rank, nw = s_.num_closest_words, d = dim();
  if (!Cuda(d_idx_.Reserve(static_cast<size_t>(std::max<int64_t>(n_max, 1)) * k * 4 + 16), "alloc", err) ||
      !Cuda(d_dist_.Reserve(static_cast<size_t>(std::max<int64_t>(n_max, 1)) * k * 4 + 16), "alloc", err))
    return false;
  last_valid_ = false;
  if (n_max == 0) {
    cudaEventRecord(ev_stage_[2], stream_);
    return true;
  }
  const size_t q_blk = static_cast<size_t>(n_max) * d * 4, c_blk = static_cast<size_t>(n_max) * nw * 4,
               l_blk = static_cast<size_t>(n_max) * k * 4;
  if (!Cuda(sh_q_all_.Reserve(q_blk * G), "alloc", err) || !Cuda(sh_cells_all_.Reserve(c_blk * G), "alloc", err) ||
      !Cuda(sh_pidx_.Reserve(l_blk * G), "alloc", err) || !Cuda(sh_pdist_.Reserve(l_blk * G), "alloc", err) ||
      !Cuda(sh_ridx_.Reserve(l_blk * G), "alloc", err) || !Cuda(sh_rdist_.Reserve(l_blk * G), "alloc", err) ||
      !Cuda(d_cells_.Reserve(c_blk), "alloc", err))
    return false;
  if (d_q_.cap < q_blk) {
    *err = "internal: query buffer smaller than the padded slice";
    return false;
  }
  // ---- 1: coarse search of the slice; padding rows visit nothing ----
  if (n_s > 0 && !CoarseChunks(d_q_.as<float>(), n_s, nw, d_cells_.as<int32_t>(), stream_, err)) return false;
  if (n_s < n_max &&
      !Cuda(cudaMemsetAsync(d_cells_.as<int32_t>() + n_s * nw, 0xFF, static_cast<size_t>(n_max - n_s) * nw * 4, stream_),
            "memset", err))
    return false;
  cudaEventRecord(ev_stage_[2], stream_);
  // ---- 2: exchange 1 on the comm stream ----
  if (!Cuda(cudaEventRecord(ev_comm_[0], stream_), "event", err) ||
      !Cuda(cudaStreamWaitEvent(comm_stream_, ev_comm_[0], 0), "wait", err))
    return false;
  if (!Nccl(g_nccl.GroupStart(), "ncclGroupStart", err) ||
      !Nccl(g_nccl.AllGather(d_q_.p, sh_q_all_.p, q_blk, ncclChar, comm, comm_stream_), "ncclAllGather", err) ||
      !Nccl(g_nccl.AllGather(d_cells_.p, sh_cells_all_.p, c_blk, ncclChar, comm, comm_stream_), "ncclAllGather", err) ||
      !Nccl(g_nccl.GroupEnd(), "ncclGroupEnd", err))
    return false;
  if (!Cuda(cudaEventRecord(ev_comm_[1], comm_stream_), "event", err)) return false;
  // ---- 3: scan + exchange 2. Three schedules (MLC_SHARD_PIPELINE, results identical):
  //   0  wait for exchange 1, ONE scan launch over all G blocks, one grouped all-to-all          (default)
  //   1  own block scanned while exchange 1 is in flight, the other blocks in one or two launches, one
  //      grouped all-to-all
  //   2  ring: block s leaves (grouped send/recv) while block s+1 is scanned
  unsigned char* pidx = sh_pidx_.as<unsigned char>();
  unsigned char* pdist = sh_pdist_.as<unsigned char>();
  unsigned char* ridx = sh_ridx_.as<unsigned char>();
  unsigned char* rdist = sh_rdist_.as<unsigned char>();
  int schedule = 0;
  if (const char* env = getenv("MLC_SHARD_PIPELINE")) {
    const int v = atoi(env);
    if (v >= 0 && v <= 2) schedule = v;
  }
  if (G == 1) schedule = 2;  // one block, nothing to exchange
  int launches = 0;
  auto scan_blocks = [&](int b0, int b1, bool from_gathered) -> bool {  // blocks [b0, b1) of the gathered order
    if (b1 <= b0) return true;
    const float* qp = from_gathered ? reinterpret_cast<const float*>(sh_q_all_.as<unsigned char>() + q_blk * b0) : d_q_.as<float>();
    const int32_t* cp = from_gathered ? reinterpret_cast<const int32_t*>(sh_cells_all_.as<unsigned char>() + c_blk * b0)
                                      : d_cells_.as<int32_t>();
    cudaEventRecord(ev_scan_[2 * launches], stream_);
    if (!Cuda(LaunchScan(qp, n_max * (b1 - b0), cp, nw, k, reinterpret_cast<int32_t*>(pidx + l_blk * b0),
                         reinterpret_cast<float*>(pdist + l_blk * b0), stream_), "list scan", err))
      return false;
    cudaEventRecord(ev_scan_[2 * launches + 1], stream_);
    ++launches;
    return true;
  };
  auto all_to_all = [&]() -> bool {  // block p of my lists -> rank p; my slice's lists from every shard
    if (!Cuda(cudaEventRecord(ev_comm_[2], stream_), "event", err) ||
        !Cuda(cudaStreamWaitEvent(comm_stream_, ev_comm_[2], 0), "wait", err))
      return false;
    if (!Nccl(g_nccl.GroupStart(), "ncclGroupStart", err)) return false;
    for (int p = 0; p < G; ++p) {
      if (p == r) continue;
      if (!Nccl(g_nccl.Send(pidx + l_blk * p, l_blk, ncclChar, p, comm, comm_stream_), "ncclSend", err) ||
          !Nccl(g_nccl.Send(pdist + l_blk * p, l_blk, ncclChar, p, comm, comm_stream_), "ncclSend", err) ||
          !Nccl(g_nccl.Recv(ridx + l_blk * p, l_blk, ncclChar, p, comm, comm_stream_), "ncclRecv", err) ||
          !Nccl(g_nccl.Recv(rdist + l_blk * p, l_blk, ncclChar, p, comm, comm_stream_), "ncclRecv", err))
        return false;
    }
    if (!Nccl(g_nccl.GroupEnd(), "ncclGroupEnd", err)) return false;
    // the own block goes straight to where the merge reads it
    return Cuda(cudaMemcpyAsync(ridx + l_blk * r, pidx + l_blk * r, l_blk, cudaMemcpyDeviceToDevice, stream_), "copy", err) &&
           Cuda(cudaMemcpyAsync(rdist + l_blk * r, pdist + l_blk * r, l_blk, cudaMemcpyDeviceToDevice, stream_), "copy", err);
  };
  if (schedule == 0) {
    if (!Cuda(cudaStreamWaitEvent(stream_, ev_comm_[1], 0), "wait", err)) return false;
    if (!scan_blocks(0, G, true) || !all_to_all()) return false;
  } else if (schedule == 1) {
    // own block from the local buffers into its place of the per-shard lists
    cudaEventRecord(ev_scan_[0], stream_);
    if (!Cuda(LaunchScan(d_q_.as<float>(), n_max, d_cells_.as<int32_t>(), nw, k, reinterpret_cast<int32_t*>(pidx + l_blk * r),
                         reinterpret_cast<float*>(pdist + l_blk * r), stream_), "list scan", err))
      return false;
    cudaEventRecord(ev_scan_[1], stream_);
    launches = 1;
    if (!Cuda(cudaStreamWaitEvent(stream_, ev_comm_[1], 0), "wait", err)) return false;
    if (!scan_blocks(0, r, true) || !scan_blocks(r + 1, G, true) || !all_to_all()) return false;
  } else {
    for (int s = 0; s < G; ++s) {
      const int src = (r + s) % G, from = (r - s + G) % G;
      const float* q_blk_p = s == 0 ? d_q_.as<float>() : reinterpret_cast<const float*>(sh_q_all_.as<unsigned char>() + q_blk * src);
      const int32_t* c_blk_p = s == 0 ? d_cells_.as<int32_t>()
                                      : reinterpret_cast<const int32_t*>(sh_cells_all_.as<unsigned char>() + c_blk * src);
      int32_t* o_idx = reinterpret_cast<int32_t*>(s == 0 ? ridx + l_blk * r : pidx + l_blk * src);
      float* o_dist = reinterpret_cast<float*>(s == 0 ? rdist + l_blk * r : pdist + l_blk * src);
      if (s == 1 && !Cuda(cudaStreamWaitEvent(stream_, ev_comm_[1], 0), "wait", err)) return false;
      cudaEventRecord(ev_scan_[2 * s], stream_);
      if (!Cuda(LaunchScan(q_blk_p, n_max, c_blk_p, nw, k, o_idx, o_dist, stream_), "list scan", err)) return false;
      cudaEventRecord(ev_scan_[2 * s + 1], stream_);
      ++launches;
      if (s == 0) continue;
      if (!Cuda(cudaEventRecord(ev_comm_[2], stream_), "event", err) ||
          !Cuda(cudaStreamWaitEvent(comm_stream_, ev_comm_[2], 0), "wait", err))
        return false;
      if (!Nccl(g_nccl.GroupStart(), "ncclGroupStart", err) ||
          !Nccl(g_nccl.Send(pidx + l_blk * src, l_blk, ncclChar, src, comm, comm_stream_), "ncclSend", err) ||
          !Nccl(g_nccl.Send(pdist + l_blk * src, l_blk, ncclChar, src, comm, comm_stream_), "ncclSend", err) ||
          !Nccl(g_nccl.Recv(ridx + l_blk * from, l_blk, ncclChar, from, comm, comm_stream_), "ncclRecv", err) ||
          !Nccl(g_nccl.Recv(rdist + l_blk * from, l_blk, ncclChar, from, comm, comm_stream_), "ncclRecv", err) ||
          !Nccl(g_nccl.GroupEnd(), "ncclGroupEnd", err))
        return false;
    }
  }
  if (!Cuda(cudaEventRecord(ev_comm_[3], comm_stream_), "event", err) ||
      !Cuda(cudaStreamWaitEvent(stream_, ev_comm_[3], 0), "wait", err))
    return false;
  // ---- 4: merge ----
  if (!Cuda(LaunchMergeTopk(sh_ridx_.as<int32_t>(), sh_rdist_.as<float>(), G, n_max, k, d_idx_.as<int32_t>(),
                            d_dist_.as<float>(), stream_), "top-k merge", err))
    return false;
  // statistics of this step's scan: all G blocks (own slice first, then the gathered ones)
  last_cells_ = sh_cells_all_.as<int32_t>();
  last_nq_ = n_max * G;
  last_nw_ = nw;
  last_scan_launches_ = launches;
  last_valid_ = true;
  return true;
}

bool Detector::ShardedKnnDevice(const float* d_q, int64_t n_s, int k, int32_t* d_idx, float* d_dist,
                                std::string* err) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  if (!comm_) {
    *err = "sharded query without a communicator: call mlc_comm_init first";
    return false;
  }
  if (k <= 0 || k > 16 || n_s < 0) {
    *err = "bad arguments";
    return false;
  }
  if (!EnsureIndex(err)) return false;
  int64_t n_max = 0;
  if (!ShardedSliceSizes(n_s, &n_max, err)) return false;
  if (!Cuda(d_q_.Reserve(static_cast<size_t>(std::max<int64_t>(n_max, 1)) * dim() * 4), "alloc", err)) return false;
  if (n_s > 0 && d_q != d_q_.as<float>() &&
      !Cuda(cudaMemcpyAsync(d_q_.p, d_q, static_cast<size_t>(n_s) * dim() * 4, cudaMemcpyDeviceToDevice, stream_),
            "copy queries", err))
    return false;
  if (!ShardedKnnOnSlice(n_s, n_max, k, err)) return false;
  if (n_s > 0 &&
      (!Cuda(cudaMemcpyAsync(d_idx, d_idx_.p, static_cast<size_t>(n_s) * k * 4, cudaMemcpyDeviceToDevice, stream_), "copy", err) ||
       !Cuda(cudaMemcpyAsync(d_dist, d_dist_.p, static_cast<size_t>(n_s) * k * 4, cudaMemcpyDeviceToDevice, stream_), "copy", err)))
    return false;
  return Cuda(cudaStreamSynchronize(stream_), "sharded knn", err);
}

// LoopDetectorNode::queryVertexInDatabase for this rank's slice of a batch of query vertices against
// the sharded database. Collective: every rank of the communicator calls it once per step.
bool Detector::ShardedQueryBatch(const mlc_frame* frames, int64_t num_frames, const uint8_t* bits,
                                 int bytes_per_desc, const double* keypoints, bool inputs_on_device,
                                 const mlc_camera* cams, int num_cams, const mlc_ransac_settings& rs,
                                 mlc_pose_result* results, int64_t* num_vertices, mlc_match* matches,
                                 int64_t capacity, int64_t* match_offsets, int64_t* num_matches,
                                 uint8_t* inlier_flags, std::string* err) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  *num_vertices = 0;
  if (num_matches) *num_matches = 0;
  if (!comm_) {
    *err = "sharded query without a communicator: call mlc_comm_init first";
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
  int64_t n_max = 0;
  if (!ShardedSliceSizes(n, &n_max, err)) return false;
  if (!Cuda(d_q_.Reserve(static_cast<size_t>(std::max<int64_t>(n_max, 1)) * dim() * 4 + 16), "alloc", err)) return false;
  const uint8_t* d_bits = bits;
  const double* d_kp = keypoints;
  if (!inputs_on_device && n > 0) {
    const size_t bb = static_cast<size_t>(n) * bytes_per_desc;
    if (!Cuda(d_bits_.Reserve(bb), "alloc", err) || !Cuda(d_query_[0].Reserve(sizeof(double) * 2 * n), "alloc", err))
      return false;
    // bits on the compute stream (needed first), keypoints — only kernel 4 reads them — on the copy stream
    if (!Cuda(cudaEventRecord(ev_copy_[kCopyChunks], stream_), "event", err) ||
        !Cuda(cudaStreamWaitEvent(copy_stream_, ev_copy_[kCopyChunks], 0), "wait", err) ||
        !Cuda(cudaMemcpyAsync(d_bits_.p, bits, bb, cudaMemcpyHostToDevice, stream_), "H2D bits", err) ||
        !Cuda(cudaMemcpyAsync(d_query_[0].p, keypoints, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, copy_stream_),
              "H2D keypoints", err) ||
        !Cuda(cudaEventRecord(ev_copy_[kCopyChunks], copy_stream_), "event", err))
      return false;
    d_bits = d_bits_.as<uint8_t>();
    d_kp = d_query_[0].as<double>();
  }
  stage_valid_ = false;
  cudaEventRecord(ev_stage_[0], stream_);
  if (n > 0 && !ProjectDevice(d_bits, bytes_per_desc, n, d_q_.as<float>(), stream_, err)) return false;
  cudaEventRecord(ev_stage_[1], stream_);
  if (!ShardedKnnOnSlice(n, n_max, k, err)) return false;  // records ev_stage_[2] after the coarse search
  if (!inputs_on_device && n > 0 && !Cuda(cudaStreamWaitEvent(stream_, ev_copy_[kCopyChunks], 0), "wait", err))
    return false;
  cudaEventRecord(ev_stage_[3], stream_);
  if (num_frames == 0) return true;
  const bool ok = QueryFromKnn(frames, num_frames, d_idx_.as<int32_t>(), d_dist_.as<float>(), k, d_kp, cams,
                               num_cams, rs, results, num_vertices, matches, capacity, match_offsets,
                               num_matches, inlier_flags, err);
  stage_valid_ = ok;
  return ok;
}

}  // namespace mlc
