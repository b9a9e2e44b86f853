// Kernel 2a of the loop-closure path: coarse word search = common::FindClosestWords
// (imilib/inverted-multi-index-common.h:148-188): two ε-approximate libnabo kd-tree searches
// (nabo/kdtree_cpu.cpp:368-447 recurseKnn, allowSelfMatch, sorted results; heap:
// nabo/index_heap.h:263-363) followed by common::MultiSequenceAlgorithm (:84-134).
//
// The kd-tree search is emulated bit-for-bit (same traversal order, same fp32 operations without
// contraction, same pruning tests), because the reference's coarse search is NOT an exact top-k
// (SURVEY F3). Work item = (query descriptor, half). To keep the warps converged the depth-first
// search runs as a per-lane state machine (descend one inner node / score one bucket point / pop
// one frame) driven in three warp-wide phases — all descending lanes walk to their leaves, all
// lanes at a leaf score their buckets, all unwinding lanes pop — so that the lanes of a warp
// execute the same code most of the time; lanes that finish pull the next work item of their
// warp's range (persistent lanes). Blocks with even/odd index own half 0 / 1 and stage only that
// half's tree in shared memory; the per-lane result heaps live in registers.
#include <cub/cub.cuh>

#include <cstdlib>

#include "device_index.h"
#include "ptx.cuh"

namespace mlc {
namespace {

constexpr int kMaxWords = 16;     // nw <= 16
constexpr int kMaxStack = 48;     // >= tree depth (checked on the host)
constexpr int kThreads = 256;

enum LaneState : int { kIdle = 0, kDescend = 1, kLeaf = 2, kPop = 3, kDone = 4 };
constexpr uint32_t kRestoreTag = 0x80000000u;

template <int D>
__device__ __forceinline__ float Pick(const float (&v)[D], uint32_t d) {
  float r = v[0];
#pragma unroll
  for (int i = 1; i < D; ++i) r = (d == static_cast<uint32_t>(i)) ? v[i] : r;
  return r;
}
template <int D>
__device__ __forceinline__ void Put(float (&v)[D], uint32_t d, float x) {
#pragma unroll
  for (int i = 0; i < D; ++i) v[i] = (d == static_cast<uint32_t>(i)) ? x : v[i];
}

// KK >= kk: capacity of the per-lane result heap, which lives in REGISTERS, right-aligned
// (entries [KK - kk, KK) are real, the ones before are -inf sentinels that never move), so that the
// k-th best ("head") is always hv[KK - 1] and an insertion is one branch-free pass.
template <int D, int KK>
__global__ void __launch_bounds__(kThreads, 4)
kd_search_kernel(CoarseParams p, const float* __restrict__ q, int64_t n, int kk1, int kk2,
                 const uint32_t* __restrict__ order, uint32_t* __restrict__ next_item, int32_t* __restrict__ out_idx,
                 float* __restrict__ out_val) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int half = blockIdx.x & 1;
  const int kk = half ? kk2 : kk1;
  const KdNodeDev* nodes = half ? p.nodes2 : p.nodes1;
  const int32_t* buckets = half ? p.buckets2 : p.buckets1;
  const float* cloud = half ? p.cloud2 : p.cloud1;
  const size_t half_base = half ? static_cast<size_t>(n) * kk1 : 0;  // [half 0: n x kk1][half 1: n x kk2]
  if (p.stage_in_smem) {
    unsigned char* dst = smem_raw;
    const uint32_t begin = half ? p.off_nodes2 : p.off_nodes1;
    const uint32_t end = half ? p.packed_bytes : p.off_nodes2;
    const unsigned char* src = reinterpret_cast<const unsigned char*>(p.packed) + begin;
    for (uint32_t i = threadIdx.x * 16; i < end - begin; i += kThreads * 16)
      *reinterpret_cast<uint4*>(dst + i) = *reinterpret_cast<const uint4*>(src + i);
    __syncthreads();
    nodes = reinterpret_cast<const KdNodeDev*>(dst + ((half ? p.off_nodes2 : p.off_nodes1) - begin));
    buckets = reinterpret_cast<const int32_t*>(dst + ((half ? p.off_buckets2 : p.off_buckets1) - begin));
    cloud = reinterpret_cast<const float*>(dst + ((half ? p.off_cloud2 : p.off_cloud1) - begin));
  }
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const uint32_t lanemask_lt = (1u << lane) - 1u;
  constexpr uint32_t kAll = 0xffffffffu;
  // Work items are handed out from one counter per half (next_item[half], zeroed by the launcher): a lane that
  // finishes takes the next unprocessed query of the whole half, so no warp runs dry while another still has a
  // range to work through (with static per-warp ranges of ~200 items 15 % of the resident warp time was idle).
  const float inf = __int_as_float(0x7f800000);
  // query coordinates and the per-dimension offsets of the traversal are indexed by the cut
  // dimension of the node: [dimension][thread] in shared memory (one LDS / STS instead of a select
  // chain over registers)
  __shared__ float qv_s[D][kThreads], off_s[D][kThreads];
  float qv[D];
  float hv[KK];
  int32_t hx[KK];
  // explicit stack: the newest entry (index sp - 1) lives in registers, entries 0 .. sp - 2 in local
  // memory — a pop hands out the register copy at once and reloads the next one in the background
  uint32_t st_tag[kMaxStack];
  float st_val[kMaxStack];
  uint32_t top_tag = 0;
  float top_val = 0.f;
  int state = kIdle, sp = 0;
  uint32_t node = 0, pos = 0, end = 0;
  float rd = 0.f;
  int64_t item = 0;

  for (;;) {
    // ---- refill idle lanes from the warp's range ----
    const uint32_t idle = __ballot_sync(kAll, state == kIdle);
    if (idle) {
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(next_item + half, static_cast<uint32_t>(__popc(idle)));
      base = __shfl_sync(kAll, base, 0);
      if (state == kIdle) {
        item = static_cast<int64_t>(base) + __popc(idle & lanemask_lt);
        const bool have = item < n;
        // processing order: queries that fall into the same leaf first are neighbours, so the lanes of a
        // warp walk (nearly) the same path; results go to the query's own row
        if (have && order) item = order[static_cast<size_t>(half) * n + item];
        if (have) {
          const float* src = q + item * (2 * D) + half * D;
#pragma unroll
          for (int d = 0; d < D; ++d) {
            qv[d] = __ldg(src + d);
            qv_s[d][tid] = qv[d];
            off_s[d][tid] = 0.f;
          }
#pragma unroll
          for (int j = 0; j < KK; ++j) {
            hv[j] = (j >= KK - kk) ? inf : -inf;
            hx[j] = -1;
          }
          node = 0;
          rd = 0.f;
          sp = 0;
          state = kDescend;
        } else {
          state = kDone;
        }
      }
    }
    if (__all_sync(kAll, state == kDone)) break;

    // ---- phase 1: every descending lane walks down to its leaf ----
    while (__any_sync(kAll, state == kDescend)) {
      if (state == kDescend) {
        const KdNodeDev nd = nodes[node];
        if (nd.dim == static_cast<uint32_t>(D)) {  // leaf
          pos = nd.cut_or_bucket;
          end = pos + nd.child_or_size;
          state = (nd.child_or_size > 0) ? kLeaf : kPop;
        } else {
          const uint32_t cd = nd.dim;
          const float old_off = off_s[cd][tid];
          const float new_off = __fsub_rn(qv_s[cd][tid], __uint_as_float(nd.cut_or_bucket));
          // rd += -old_off * old_off + new_off * new_off   (for the far child)
          if (sp > 0) {
            st_val[sp - 1] = top_val;
            st_tag[sp - 1] = top_tag;
          }
          top_val = __fadd_rn(rd, __fadd_rn(__fmul_rn(-old_off, old_off), __fmul_rn(new_off, new_off)));
          top_tag = node;  // far child and offsets are re-derived from the parent when popped
          ++sp;
          node = (new_off > 0.f) ? nd.child_or_size : node + 1;  // near child first
        }
      }
    }

    // ---- phase 2: every lane at a leaf scores its bucket, point by point ----
    while (__any_sync(kAll, state == kLeaf)) {
      if (state == kLeaf) {
        const int pidx = buckets[pos];
        const float* pt = cloud + static_cast<size_t>(pidx) * D;
        float dist = 0.f;
#pragma unroll
        for (int j = 0; j < D; ++j) {
          const float diff = __fsub_rn(qv[j], pt[j]);
          dist = __fadd_rn(dist, __fmul_rn(diff, diff));
        }
        if ((dist <= p.max_radius2) && (dist < hv[KK - 1])) {
          // IndexHeapBruteForceVector::replaceHead: the new entry goes behind every entry that is
          // not strictly larger; the largest one drops out
          bool prev_gt = false;  // hv[i - 1] > dist
          float prev_v = 0.f;
          int32_t prev_x = 0;
#pragma unroll
          for (int i = 0; i < KK; ++i) {
            const float v = hv[i];
            const int32_t x = hx[i];
            const bool gt = v > dist;
            hv[i] = prev_gt ? prev_v : (gt ? dist : v);
            hx[i] = prev_gt ? prev_x : (gt ? pidx : x);
            prev_gt = gt;
            prev_v = v;
            prev_x = x;
          }
        }
        if (++pos == end) state = kPop;
      }
    }

    // ---- phase 3: unwind until a far subtree must be visited (or the search ends) ----
    while (__any_sync(kAll, state == kPop)) {
      if (state == kPop) {
        if (sp == 0) {
          // search finished: emit the sorted heap
          int32_t* oi = out_idx + half_base + static_cast<size_t>(item) * kk;
          float* ov = out_val + half_base + static_cast<size_t>(item) * kk;
#pragma unroll
          for (int j = 0; j < KK; ++j) {
            if (j >= KK - kk) {
              oi[j - (KK - kk)] = hx[j];
              ov[j - (KK - kk)] = hv[j];
            }
          }
          state = kIdle;
        } else {
          const uint32_t tag = top_tag;
          const float val = top_val;
          bool replaced = false;
          if (tag & kRestoreTag) {
            off_s[tag & 0xFFu][tid] = val;  // leave the far subtree: restore the offset
          } else {
            const float frd = val;
            if ((frd <= p.max_radius2) && (__fmul_rn(frd, p.max_error2) < hv[KK - 1])) {
              const KdNodeDev nd = nodes[tag];
              const uint32_t cd = nd.dim;
              const float new_off = __fsub_rn(qv_s[cd][tid], __uint_as_float(nd.cut_or_bucket));
              top_tag = kRestoreTag | cd;   // the popped entry is replaced by the restore entry
              top_val = off_s[cd][tid];     // old offset (the near subtree restored it)
              replaced = true;
              off_s[cd][tid] = new_off;
              node = (new_off > 0.f) ? tag + 1 : nd.child_or_size;  // far child
              rd = frd;
              state = kDescend;
            }
          }
          if (!replaced) {
            --sp;
            if (sp > 0) {
              top_tag = st_tag[sp - 1];
              top_val = st_val[sp - 1];
            }
          }
        }
      }
    }
  }
}

// Leaf a query reaches by plain descent (no backtracking) in its half's tree: the sort key that makes
// neighbouring work items traverse alike. keys[half * n + i], vals[half * n + i] = i.
template <int D>
__global__ void __launch_bounds__(256) kd_leaf_key_kernel(CoarseParams p, const float* __restrict__ q, int64_t n,
                                                          uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (t >= 2 * n) return;
  const int half = t >= n ? 1 : 0;
  const int64_t i = t - static_cast<int64_t>(half) * n;
  const KdNodeDev* nodes = half ? p.nodes2 : p.nodes1;
  const float* src = q + i * (2 * D) + half * D;
  float qv[D];
#pragma unroll
  for (int d = 0; d < D; ++d) qv[d] = __ldg(src + d);
  uint32_t node = 0;
  for (int depth = 0; depth < kMaxStack; ++depth) {
    const KdNodeDev nd = nodes[node];
    if (nd.dim == static_cast<uint32_t>(D)) break;
    const float off = __fsub_rn(Pick<D>(qv, nd.dim), __uint_as_float(nd.cut_or_bucket));
    node = (off > 0.f) ? nd.child_or_size : node + 1;
  }
  keys[t] = (static_cast<uint32_t>(half) << 16) | node;  // < 2^16 nodes per tree (checked on the host)
  vals[t] = static_cast<uint32_t>(i);
}

// MultiSequenceAlgorithm: pop the pairs (i1, i2) in ascending (d1[i1] + d2[i2], i1, i2).
__global__ void __launch_bounds__(128)
multi_sequence_kernel(const int32_t* __restrict__ h_idx, const float* __restrict__ h_val, int64_t n,
                      int n1, int n2, int num_words, int w2, int32_t* __restrict__ cells) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  int32_t idx1[kMaxWords], idx2[kMaxWords];
  float d1[kMaxWords], d2[kMaxWords];
  for (int j = 0; j < n1; ++j) {
    idx1[j] = h_idx[i * n1 + j];
    d1[j] = h_val[i * n1 + j];
  }
  const size_t base2 = static_cast<size_t>(n) * n1;
  for (int j = 0; j < n2; ++j) {
    idx2[j] = h_idx[base2 + i * n2 + j];
    d2[j] = h_val[base2 + i * n2 + j];
  }
  uint32_t used[(kMaxWords * kMaxWords) / 32];
#pragma unroll
  for (int j = 0; j < (kMaxWords * kMaxWords) / 32; ++j) used[j] = 0;
  float pq_sum[kMaxWords + 4];
  int pq_i1[kMaxWords + 4], pq_i2[kMaxWords + 4];
  pq_sum[0] = __fadd_rn(d1[0], d2[0]);
  pq_i1[0] = 0;
  pq_i2[0] = 0;
  int pq_n = 1, emitted = 0;
  int32_t* dst = cells + i * num_words;
  while (pq_n > 0 && emitted < num_words) {
    int best = 0;
    for (int j = 1; j < pq_n; ++j) {
      const bool less = (pq_sum[j] < pq_sum[best]) ||
                        (!(pq_sum[best] < pq_sum[j]) &&
                         ((pq_i1[j] < pq_i1[best]) || (pq_i1[j] == pq_i1[best] && pq_i2[j] < pq_i2[best])));
      if (less) best = j;
    }
    const int i1 = pq_i1[best], i2 = pq_i2[best];
    --pq_n;
    pq_sum[best] = pq_sum[pq_n];
    pq_i1[best] = pq_i1[pq_n];
    pq_i2[best] = pq_i2[pq_n];
    const int word_index = i1 * n2 + i2;
    used[word_index >> 5] |= 1u << (word_index & 31);
    const int a = idx1[i1], b = idx2[i2];
    // A pair with a missing word (fewer than nw words inside the radius) is skipped (-1).
    dst[emitted++] = (a < 0 || b < 0) ? -1 : a * w2 + b;
    if (i1 + 1 < n1) {
      const int nb = word_index + n2 - 1;
      if (i2 == 0 || ((used[nb >> 5] >> (nb & 31)) & 1u)) {
        pq_sum[pq_n] = __fadd_rn(d1[i1 + 1], d2[i2]);
        pq_i1[pq_n] = i1 + 1;
        pq_i2[pq_n] = i2;
        ++pq_n;
      }
    }
    if (i2 + 1 < n2) {
      const int nb = word_index - n2 + 1;
      if (i1 == 0 || ((used[nb >> 5] >> (nb & 31)) & 1u)) {
        pq_sum[pq_n] = __fadd_rn(d1[i1], d2[i2 + 1]);
        pq_i1[pq_n] = i1;
        pq_i2[pq_n] = i2 + 1;
        ++pq_n;
      }
    }
  }
  for (int j = emitted; j < num_words; ++j) dst[j] = -1;
}

template <int D, int KK>
cudaError_t LaunchSearchK(const CoarseParams& p, const float* d_q, int64_t n, int kk1, int kk2,
                          const uint32_t* order, uint32_t* next_item, int32_t* h_idx, float* h_val, int sm_count,
                          cudaStream_t stream) {
  const uint32_t tree_bytes =
      p.stage_in_smem ? max(p.off_nodes2, p.packed_bytes - p.off_nodes2) : 0u;
  const size_t smem = tree_bytes;
  cudaError_t e = cudaFuncSetAttribute(kd_search_kernel<D, KK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  int per_sm = 1;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kd_search_kernel<D, KK>, kThreads, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  int64_t blocks_per_half = (static_cast<int64_t>(sm_count) * per_sm) / 2;
  const int64_t needed = (n + kThreads - 1) / kThreads;  // at least one query per lane to start with
  if (blocks_per_half > needed) blocks_per_half = needed;
  if (blocks_per_half < 1) blocks_per_half = 1;
  e = cudaMemsetAsync(next_item, 0, 2 * sizeof(uint32_t), stream);
  if (e != cudaSuccess) return e;
  kd_search_kernel<D, KK><<<static_cast<unsigned>(2 * blocks_per_half), kThreads, smem, stream>>>(
      p, d_q, n, kk1, kk2, order, next_item, h_idx, h_val);
  CountLaunch();
  return cudaGetLastError();
}

// Experiment, off by default (MLC_COARSE_SORT=1 switches it on; results do not depend on it): work items
// sorted by the leaf their query falls into, so that the lanes of a warp start on the same path. Measured
// on the headline step (500 k queries x 2 halves): 1.02 ms with the sort against 0.90 ms without — the
// descent + radix sort cost more than the better convergence gains. Sort buffers: keys / values, in and
// out, behind the word lists in `scratch`.
constexpr int64_t kSortMinItems = 8192;
bool SortEnabled() {
  static const bool on = [] {
    const char* env = getenv("MLC_COARSE_SORT");
    return env && atoi(env) == 1;
  }();
  return on;
}
size_t SortTempBytes(int64_t n) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, static_cast<const uint32_t*>(nullptr), static_cast<uint32_t*>(nullptr),
                                  static_cast<const uint32_t*>(nullptr), static_cast<uint32_t*>(nullptr),
                                  static_cast<int>(2 * n), 0, 17);
  return bytes;
}

template <int D>
cudaError_t LaunchSearch(const CoarseParams& p, const float* d_q, int64_t n, int kk1, int kk2,
                         int32_t* h_idx, float* h_val, void* sort_scratch, uint32_t* next_item, int sm_count,
                         cudaStream_t stream) {
  const uint32_t* order = nullptr;
  if (sort_scratch) {
    uint32_t* keys = static_cast<uint32_t*>(sort_scratch);
    uint32_t* vals = keys + 2 * n;
    uint32_t* keys_out = vals + 2 * n;
    uint32_t* vals_out = keys_out + 2 * n;
    void* tmp = vals_out + 2 * n;
    size_t tmp_bytes = SortTempBytes(n);
    kd_leaf_key_kernel<D><<<static_cast<unsigned>((2 * n + 255) / 256), 256, 0, stream>>>(p, d_q, n, keys, vals);
    CountLaunch();
    cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys_out, vals, vals_out,
                                                    static_cast<int>(2 * n), 0, 17, stream);
    if (e != cudaSuccess) return e;
    CountLaunch();
    order = vals_out;
  }
  const int kk_max = max(kk1, kk2);
  if (kk_max <= 1) return LaunchSearchK<D, 1>(p, d_q, n, kk1, kk2, order, next_item, h_idx, h_val, sm_count, stream);
  if (kk_max <= 10) return LaunchSearchK<D, 10>(p, d_q, n, kk1, kk2, order, next_item, h_idx, h_val, sm_count, stream);
  return LaunchSearchK<D, 16>(p, d_q, n, kk1, kk2, order, next_item, h_idx, h_val, sm_count, stream);
}

size_t WordListBytes(int64_t n, int kk1, int kk2) {
  return ((static_cast<size_t>(n) * (kk1 + kk2) * 4 + 127) & ~static_cast<size_t>(127)) * 2;
}

}  // namespace

size_t CoarseScratchBytes(const CoarseParams& p, int64_t n, int num_words) {
  const int kk1 = min(p.num_words1, num_words), kk2 = min(p.num_words2, num_words);
  size_t bytes = WordListBytes(n, kk1, kk2) + 256;
  if (SortEnabled() && n >= kSortMinItems) bytes += static_cast<size_t>(n) * 2 * 4 * 4 + SortTempBytes(n) + 256;
  return ((bytes + 255) & ~static_cast<size_t>(255)) + 256;  // the last 256 bytes: work-item counters of the search
}

cudaError_t LaunchCoarseWords(const CoarseParams& p, const float* d_q, int64_t n, int num_words,
                              int32_t* d_cells, void* scratch, int sm_count, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  if (num_words <= 0 || num_words > kMaxWords) return cudaErrorInvalidValue;
  const int kk1 = min(p.num_words1, num_words), kk2 = min(p.num_words2, num_words);
  int32_t* h_idx = static_cast<int32_t*>(scratch);
  float* h_val = reinterpret_cast<float*>(static_cast<unsigned char*>(scratch) +
                                          ((static_cast<size_t>(n) * (kk1 + kk2) * 4 + 127) & ~static_cast<size_t>(127)));
  void* sort_scratch = nullptr;
  if (SortEnabled() && n >= kSortMinItems && p.num_words1 < 65536 && p.num_words2 < 65536)
    sort_scratch = static_cast<unsigned char*>(scratch) + ((WordListBytes(n, kk1, kk2) + 255) & ~static_cast<size_t>(255));
  uint32_t* next_item = reinterpret_cast<uint32_t*>(static_cast<unsigned char*>(scratch) +
                                                    CoarseScratchBytes(p, n, num_words) - 256);
  cudaError_t e;
  switch (p.sub_dim) {
#define MLC_CASE(D)                                                                                \
  case D:                                                                                          \
    e = LaunchSearch<D>(p, d_q, n, kk1, kk2, h_idx, h_val, sort_scratch, next_item, sm_count, stream); \
    break;
    MLC_CASE(1) MLC_CASE(2) MLC_CASE(3) MLC_CASE(4) MLC_CASE(5) MLC_CASE(6) MLC_CASE(7) MLC_CASE(8)
#undef MLC_CASE
    default:
      return cudaErrorInvalidValue;
  }
  if (e != cudaSuccess) return e;
  multi_sequence_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, stream>>>(
      h_idx, h_val, n, kk1, kk2, num_words, p.num_words2, d_cells);
  CountLaunch();
  return cudaGetLastError();
}

}  // namespace mlc
