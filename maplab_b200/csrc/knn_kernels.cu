// Kernel 2 of the loop-closure path: inverted-multi-index kNN.
//   2a  (coarse_kernels.cu)   — FindClosestWords: kd-tree searches + multi-sequence
//   2b  imi_scan_kernel       — InvertedMultiIndex::GetNNearestNeighbors: stream the visited
//                               cells' inverted lists, squared L2 in the canonical fp32 order,
//                               top-k by (distance, index) (imilib/inverted-multi-index.h:100-161,
//                               inverted-multi-index-common.h:54-72)
//   index build                — cell assignment (2a with one word), radix sort by cell, blocked
//                               SoA inverted lists (imilib/inverted-multi-index.h:77-94)
#include <cub/cub.cuh>

#include "device_index.h"
#include "ptx.cuh"

namespace mlc {

// =============================================================================================
// 2b: inverted-list scan
// =============================================================================================
namespace {

constexpr int kMaxWords = 16;   // nw <= 16
constexpr uint64_t kEmptyKey = 0x7f800000FFFFFFFFull;  // (+inf, idx -1): sorts after every entry

// (stored - query).squaredNorm() in the canonical fp32 order of the oracle (packets of four
// squared differences accumulated lane-wise, (l0+l1)+(l2+l3), scalar remainder added last).
template <int DIM>
__device__ __forceinline__ float SquaredDistanceCanonical(const float (&a)[DIM],
                                                          const float (&b)[DIM]) {
  constexpr int kVec = (DIM / 4) * 4;
  if constexpr (kVec == 0) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < DIM; ++i) {
      const float d = __fsub_rn(a[i], b[i]);
      const float sq = __fmul_rn(d, d);
      s = (i == 0) ? sq : __fadd_rn(s, sq);
    }
    return s;
  } else {
    float lane[4];
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const float d = __fsub_rn(a[l], b[l]);
      lane[l] = __fmul_rn(d, d);
    }
#pragma unroll
    for (int p = 4; p < kVec; p += 4) {
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        const float d = __fsub_rn(a[p + l], b[p + l]);
        lane[l] = __fadd_rn(lane[l], __fmul_rn(d, d));
      }
    }
    float res = __fadd_rn(__fadd_rn(lane[0], lane[1]), __fadd_rn(lane[2], lane[3]));
    if constexpr (kVec < DIM) {
      float rem = 0.f;
#pragma unroll
      for (int i = kVec; i < DIM; ++i) {
        const float d = __fsub_rn(a[i], b[i]);
        const float sq = __fmul_rn(d, d);
        rem = (i == kVec) ? sq : __fadd_rn(rem, sq);
      }
      res = __fadd_rn(res, rem);
    }
    return res;
  }
}

// One warp per query descriptor, persistent warps. The (<= 16) visited cells' lists are flattened
// into one entry range and dealt to the lanes 64 at a time (two independent 32-entry windows per
// trip, so 2 x (DIM + 1) loads per lane are in flight); inside a block of a list the layout is
// [dim+1][block size] words, so the 32 lanes of a full block read 128 contiguous bytes per
// dimension. Query metadata is software-pipelined two deep: while query i is scanned, the cell
// table entries of query i+1 and the visit list of query i+2 are already in flight, which hides
// the dependent chain visit list -> cell table -> list entries.
constexpr int kScanThreads = 256;
constexpr int kScanCtasPerSm = 3;
constexpr uint32_t kFull = 0xffffffffu;

struct ScanEntry {
  uint64_t key;
};

template <int DIM>
__device__ __forceinline__ void LoadWindow(const uint32_t* __restrict__ lists, const uint4* seg,
                                           uint32_t base, uint32_t total, uint32_t len,
                                           uint32_t excl, int lane, uint32_t lanemask_le,
                                           float (&sv)[DIM], uint32_t* id, bool* valid) {
  // non-empty cells whose first entry falls into this window set bit (first entry - base)
  const uint32_t rel = excl - base;
  const uint32_t heads = __reduce_or_sync(kFull, (len > 0 && rel < 32u) ? (1u << rel) : 0u);
  const uint32_t before = __popc(__ballot_sync(kFull, len > 0 && excl < base));
  const uint32_t e = base + lane;
  *valid = e < total;
  if (*valid) {
    const uint4 s = seg[before + __popc(heads & lanemask_le) - 1];  // {excl, start16, len}
    const uint32_t within = e - s.x;            // entry number inside its cell
    const uint32_t blk = within >> 5;           // block of 32 entries
    const uint32_t bs = min(32u, s.z - (blk << 5));
    const uint32_t* w = lists + (static_cast<size_t>(s.y) << 2) +
                        static_cast<size_t>(blk) * ((DIM + 1) * 32) + (within & 31u);
#pragma unroll
    for (int d = 0; d < DIM; ++d) sv[d] = __uint_as_float(__ldg(w + d * bs));
    *id = __ldg(w + DIM * bs);
  }
}

template <int KT>
__device__ __forceinline__ void InsertKey(uint64_t (&best)[KT], uint64_t key) {
  if (key < best[KT - 1]) {
#pragma unroll
    for (int i = 0; i < KT; ++i) {
      const bool lt = key < best[i];
      const uint64_t lo = lt ? key : best[i];
      const uint64_t hi = lt ? best[i] : key;
      best[i] = lo;
      key = hi;
    }
  }
}

template <int DIM, int KT>
__global__ void __launch_bounds__(kScanThreads, kScanCtasPerSm)
imi_scan_kernel(const float* __restrict__ q, int64_t n_q, const int32_t* __restrict__ cells, int nw,
                const uint2* __restrict__ cell_info, const uint32_t* __restrict__ lists, int k,
                int32_t* __restrict__ out_idx, float* __restrict__ out_dist) {
  __shared__ uint4 seg_s[kScanThreads / 32][kMaxWords];
  const int lane = threadIdx.x & 31;
  uint4* seg = seg_s[threadIdx.x >> 5];
  const uint32_t lanemask_lt = (1u << lane) - 1u;
  const uint32_t lanemask_le = lanemask_lt | (1u << lane);
  const int64_t warp_stride = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  int64_t qi = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  if (qi >= n_q) return;

  // pipeline prologue: visit list of queries 0 and 1, cell table entries + descriptor of query 0
  int32_t c_next = -1;
  uint2 info = make_uint2(0u, 0u);
  float q_l = 0.f;
  {
    int32_t c = -1;
    if (lane < nw) {
      c = __ldg(cells + qi * nw + lane);
      if (qi + warp_stride < n_q) c_next = __ldg(cells + (qi + warp_stride) * nw + lane);
    }
    if (lane < DIM) q_l = __ldg(q + qi * DIM + lane);
    if (c >= 0) info = __ldg(cell_info + c);
  }

  for (; qi < n_q; qi += warp_stride) {
    // ---- prefetch: cell table of query i+1, visit list of query i+2, descriptor of query i+1
    uint2 info_next = make_uint2(0u, 0u);
    int32_t c_next2 = -1;
    float q_next_l = 0.f;
    if (c_next >= 0) info_next = __ldg(cell_info + c_next);
    if (qi + warp_stride < n_q) {
      if (lane < DIM) q_next_l = __ldg(q + (qi + warp_stride) * DIM + lane);
      if (lane < nw && qi + 2 * warp_stride < n_q)
        c_next2 = __ldg(cells + (qi + 2 * warp_stride) * nw + lane);
    }

    // ---- this query
    const uint32_t start16 = info.x, len = info.y;
    uint32_t incl = len;  // inclusive prefix sum over the (<= 16) cell lanes
#pragma unroll
    for (int o = 1; o < kMaxWords; o <<= 1) {
      const uint32_t v = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl += v;
    }
    const uint32_t total = __shfl_sync(kFull, incl, kMaxWords - 1);
    const uint32_t excl = incl - len;
    const uint32_t nonempty = __ballot_sync(kFull, len > 0);
    if (len > 0) seg[__popc(nonempty & lanemask_lt)] = make_uint4(excl, start16, len, 0u);
    float qv[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) qv[d] = __shfl_sync(kFull, q_l, d);
    __syncwarp();

    uint64_t best[KT];
#pragma unroll
    for (int i = 0; i < KT; ++i) best[i] = kEmptyKey;

    for (uint32_t base = 0; base < total; base += 64) {
      float sa[DIM], sb[DIM];
      uint32_t ida = 0xFFFFFFFFu, idb = 0xFFFFFFFFu;
      bool va = false, vb = false;
      LoadWindow<DIM>(lists, seg, base, total, len, excl, lane, lanemask_le, sa, &ida, &va);
      const bool second = base + 32 < total;  // warp-uniform
      if (second)
        LoadWindow<DIM>(lists, seg, base + 32, total, len, excl, lane, lanemask_le, sb, &idb, &vb);
      uint64_t ka = kEmptyKey, kb = kEmptyKey;
      if (va)
        ka = (static_cast<uint64_t>(__float_as_uint(SquaredDistanceCanonical<DIM>(sa, qv))) << 32) | ida;
      if (vb)
        kb = (static_cast<uint64_t>(__float_as_uint(SquaredDistanceCanonical<DIM>(sb, qv))) << 32) | idb;
      if (base == 0) {
        // first trip: the per-lane list is empty, so the two keys only need ordering
        if constexpr (KT >= 2) {
          best[0] = ka < kb ? ka : kb;
          best[1] = ka < kb ? kb : ka;
        } else {
          best[0] = ka < kb ? ka : kb;
        }
      } else {
        InsertKey<KT>(best, ka);
        if (second) InsertKey<KT>(best, kb);
      }
    }
    __syncwarp();  // every lane is done with seg before the next query overwrites it

    // ---- k rounds of warp arg-min over the lanes' heads; lane r keeps the r-th result
    uint32_t res_d = 0x7f800000u, res_i = 0xFFFFFFFFu;
    for (int r = 0; r < k; ++r) {
      const uint32_t hd = static_cast<uint32_t>(best[0] >> 32);
      const uint32_t id = static_cast<uint32_t>(best[0]);
      const uint32_t hi_min = __reduce_min_sync(kFull, hd);
      const uint32_t lo_min = __reduce_min_sync(kFull, (hd == hi_min) ? id : 0xFFFFFFFFu);
      if (lane == r) {
        res_d = hi_min;
        res_i = lo_min;
      }
      if (hi_min == 0x7f800000u && lo_min == 0xFFFFFFFFu) break;  // nothing left (warp-uniform)
      if (hd == hi_min && id == lo_min) {
#pragma unroll
        for (int i = 0; i + 1 < KT; ++i) best[i] = best[i + 1];
        best[KT - 1] = kEmptyKey;
      }
    }
    if (lane < k) {  // missing neighbours: (+inf, -1), trailing
      out_idx[qi * k + lane] = static_cast<int32_t>(res_i);
      out_dist[qi * k + lane] = __uint_as_float(res_d);
    }

    // ---- rotate the pipeline
    info = info_next;
    c_next = c_next2;
    q_l = q_next_l;
  }
}

template <int DIM>
cudaError_t LaunchScanDim(const float* q, int64_t n_q, const int32_t* cells, int nw,
                          const uint2* cell_info, const uint32_t* lists, int k, int32_t* out_idx,
                          float* out_dist, int sm_count, cudaStream_t stream) {
  const int threads = kScanThreads;
  int64_t blocks = (n_q * 32 + threads - 1) / threads;
  const int64_t cap = static_cast<int64_t>(sm_count) * kScanCtasPerSm;  // persistent: every CTA resident
  if (blocks > cap) blocks = cap;
  const unsigned g = static_cast<unsigned>(blocks);
#define MLC_SCAN(KT)                                                                         \
  imi_scan_kernel<DIM, KT><<<g, threads, 0, stream>>>(q, n_q, cells, nw, cell_info, lists, k, \
                                                      out_idx, out_dist)
  if (k <= 1) MLC_SCAN(1);
  else if (k <= 2) MLC_SCAN(2);
  else if (k <= 4) MLC_SCAN(4);
  else if (k <= 6) MLC_SCAN(6);
  else if (k <= 8) MLC_SCAN(8);
  else if (k <= 10) MLC_SCAN(10);
  else MLC_SCAN(16);
#undef MLC_SCAN
  CountLaunch();
  return cudaGetLastError();
}

}  // namespace

cudaError_t LaunchImiScan(int dim, const float* q, int64_t n_q, const int32_t* cells, int nw,
                          const uint2* cell_info, const uint32_t* lists, int k, int32_t* out_idx,
                          float* out_dist, int sm_count, cudaStream_t stream) {
  if (n_q <= 0) return cudaSuccess;
  if (k > 16 || nw > kMaxWords) return cudaErrorInvalidValue;
  switch (dim) {
    case 10:
      return LaunchScanDim<10>(q, n_q, cells, nw, cell_info, lists, k, out_idx, out_dist, sm_count,
                               stream);
    case 6:
      return LaunchScanDim<6>(q, n_q, cells, nw, cell_info, lists, k, out_idx, out_dist, sm_count,
                              stream);
    case 4:
      return LaunchScanDim<4>(q, n_q, cells, nw, cell_info, lists, k, out_idx, out_dist, sm_count,
                              stream);
    default:
      return cudaErrorInvalidValue;
  }
}

// ---------------------------------------------------------------------------------------------
// scan statistics: algorithmic entries visited = sum over (query, cell) of the list length
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void scan_entries_kernel(const int32_t* __restrict__ cells, int64_t n_visits,
                                    const uint2* __restrict__ cell_info,
                                    unsigned long long* __restrict__ total) {
  unsigned long long local = 0;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n_visits;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int32_t c = cells[i];
    if (c >= 0) local += cell_info[c].y;
  }
  typedef cub::BlockReduce<unsigned long long, 256> Reduce;
  __shared__ typename Reduce::TempStorage tmp;
  const unsigned long long sum = Reduce(tmp).Sum(local);
  if (threadIdx.x == 0 && sum) atomicAdd(total, sum);  // one per block; statistics only
}
}  // namespace

cudaError_t LaunchScanEntries(const int32_t* cells, int64_t n_visits, const uint2* cell_info,
                              unsigned long long* d_total, cudaStream_t stream) {
  if (n_visits <= 0) return cudaSuccess;
  int64_t blocks = (n_visits + 255) / 256;
  if (blocks > 1184) blocks = 1184;
  scan_entries_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(cells, n_visits, cell_info,
                                                                         d_total);
  CountLaunch();
  return cudaGetLastError();
}

// =============================================================================================
// cross-shard top-k merge (consumer of the all-gather)
// =============================================================================================
namespace {
__global__ void merge_topk_kernel(const int32_t* __restrict__ idx_lists,
                                  const float* __restrict__ dist_lists, int num_lists, int64_t n_q,
                                  int k, int32_t* __restrict__ out_idx,
                                  float* __restrict__ out_dist) {
  const int64_t qi = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (qi >= n_q) return;
  int head[16];
  for (int l = 0; l < num_lists; ++l) head[l] = 0;
  for (int r = 0; r < k; ++r) {
    uint64_t best = kEmptyKey;
    int best_l = -1;
    for (int l = 0; l < num_lists; ++l) {
      if (head[l] >= k) continue;
      const size_t at = (static_cast<size_t>(l) * n_q + qi) * k + head[l];
      const int32_t id = idx_lists[at];
      if (id < 0) continue;  // missing neighbours are trailing
      const uint64_t key =
          (static_cast<uint64_t>(__float_as_uint(dist_lists[at])) << 32) | static_cast<uint32_t>(id);
      if (key < best) {
        best = key;
        best_l = l;
      }
    }
    if (best_l < 0) {
      out_idx[qi * k + r] = -1;
      out_dist[qi * k + r] = __int_as_float(0x7f800000);
    } else {
      out_idx[qi * k + r] = static_cast<int32_t>(best & 0xFFFFFFFFu);
      out_dist[qi * k + r] = __uint_as_float(static_cast<uint32_t>(best >> 32));
      ++head[best_l];
    }
  }
}
}  // namespace

cudaError_t LaunchMergeTopk(const int32_t* idx_lists, const float* dist_lists, int num_lists,
                            int64_t n_q, int k, int32_t* out_idx, float* out_dist,
                            cudaStream_t stream) {
  if (n_q <= 0) return cudaSuccess;
  if (num_lists > 16) return cudaErrorInvalidValue;
  const int threads = 128;
  merge_topk_kernel<<<static_cast<unsigned>((n_q + threads - 1) / threads), threads, 0, stream>>>(
      idx_lists, dist_lists, num_lists, n_q, k, out_idx, out_dist);
  CountLaunch();
  return cudaGetLastError();
}

// =============================================================================================
// index build
// =============================================================================================
namespace {

// cells of this shard's descriptors -> sort keys; descriptors of other shards get the key
// `num_cells` (sorted to the end and ignored).
__global__ void shard_keys_kernel(const int32_t* __restrict__ cells, int64_t n, int shard_rank,
                                  int shard_count, uint32_t num_cells,
                                  uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const int32_t c = cells[i];
  const bool mine = (i % shard_count) == shard_rank && c >= 0;
  keys[i] = mine ? static_cast<uint32_t>(c) : num_cells;
  vals[i] = static_cast<uint32_t>(i);
}

// first/last position of every cell in the sorted key array (no atomics).
__global__ void cell_bounds_kernel(const uint32_t* __restrict__ keys, int64_t n, uint32_t num_cells,
                                   uint32_t* __restrict__ first, uint32_t* __restrict__ len) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const uint32_t c = keys[i];
  if (c >= num_cells) return;
  if (i == 0 || keys[i - 1] != c) first[c] = static_cast<uint32_t>(i);
  if (i == n - 1 || keys[i + 1] != c) len[c] = static_cast<uint32_t>(i);  // last, fixed up below
}
__global__ void cell_sizes_kernel(uint32_t num_cells, const uint32_t* __restrict__ first,
                                  uint32_t* __restrict__ len, int words_per_entry,
                                  uint32_t* __restrict__ size16) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= num_cells) return;
  uint32_t l = 0;
  if (first[c] != 0xFFFFFFFFu) l = len[c] - first[c] + 1;
  len[c] = l;
  size16[c] = (l * static_cast<uint32_t>(words_per_entry) * 4u + 15u) >> 4;
}
__global__ void cell_info_kernel(uint32_t num_cells, const uint32_t* __restrict__ start16,
                                 const uint32_t* __restrict__ len, uint2* __restrict__ info) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= num_cells) return;
  info[c] = make_uint2(start16[c], len[c]);
}
// Scatter descriptor i into block-SoA position of its cell.
__global__ void fill_lists_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                  int64_t n, uint32_t num_cells, const uint32_t* __restrict__ first,
                                  const uint2* __restrict__ info, const float* __restrict__ desc,
                                  int dim, uint32_t* __restrict__ lists) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const uint32_t c = keys[i];
  if (c >= num_cells) return;
  const uint32_t within = static_cast<uint32_t>(i) - first[c];
  const uint2 ci = info[c];
  const uint32_t blk = within >> 5;
  const uint32_t bs = min(32u, ci.y - (blk << 5));
  uint32_t* w = lists + (static_cast<size_t>(ci.x) << 2) + static_cast<size_t>(blk) * ((dim + 1) * 32) +
                (within & 31u);
  const uint32_t id = vals[i];
  const float* src = desc + static_cast<size_t>(id) * dim;
  for (int d = 0; d < dim; ++d) w[d * bs] = __float_as_uint(src[d]);
  w[dim * bs] = id;
}

}  // namespace

#define MLC_CUDA_TRY(x)                \
  do {                                 \
    cudaError_t e__ = (x);             \
    if (e__ != cudaSuccess) return e__; \
  } while (0)

cudaError_t BuildImiLists(const int32_t* d_cells, const float* d_desc, int64_t n, int dim,
                          uint32_t num_cells, int shard_rank, int shard_count, DeviceLists* out,
                          cudaStream_t stream) {
  out->Free();
  out->num_cells = num_cells;
  out->dim = dim;
  MLC_CUDA_TRY(cudaMalloc(&out->cell_info, sizeof(uint2) * static_cast<size_t>(num_cells)));
  if (n == 0) {
    MLC_CUDA_TRY(cudaMemsetAsync(out->cell_info, 0, sizeof(uint2) * static_cast<size_t>(num_cells),
                                 stream));
    MLC_CUDA_TRY(cudaMalloc(&out->lists, 16));
    out->list_bytes = 0;
    return cudaStreamSynchronize(stream);
  }
  uint32_t *keys = nullptr, *vals = nullptr, *keys_s = nullptr, *vals_s = nullptr;
  uint32_t *first = nullptr, *len = nullptr, *size16 = nullptr, *start16 = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0, tmp_scan = 0;
  cudaError_t err = cudaSuccess;
  auto cleanup = [&]() {
    cudaFree(keys);
    cudaFree(vals);
    cudaFree(keys_s);
    cudaFree(vals_s);
    cudaFree(first);
    cudaFree(len);
    cudaFree(size16);
    cudaFree(start16);
    cudaFree(tmp);
  };
#define MLC_TRY_C(x)          \
  do {                        \
    err = (x);                \
    if (err != cudaSuccess) { \
      cleanup();              \
      return err;             \
    }                         \
  } while (0)
  MLC_TRY_C(cudaMalloc(&keys, 4 * n));
  MLC_TRY_C(cudaMalloc(&vals, 4 * n));
  MLC_TRY_C(cudaMalloc(&keys_s, 4 * n));
  MLC_TRY_C(cudaMalloc(&vals_s, 4 * n));
  MLC_TRY_C(cudaMalloc(&first, 4 * static_cast<size_t>(num_cells)));
  MLC_TRY_C(cudaMalloc(&len, 4 * static_cast<size_t>(num_cells)));
  MLC_TRY_C(cudaMalloc(&size16, 4 * static_cast<size_t>(num_cells)));
  MLC_TRY_C(cudaMalloc(&start16, 4 * static_cast<size_t>(num_cells)));
  const unsigned nb = static_cast<unsigned>((n + 255) / 256);
  const unsigned cb = (num_cells + 255) / 256;
  shard_keys_kernel<<<nb, 256, 0, stream>>>(d_cells, n, shard_rank, shard_count, num_cells, keys,
                                            vals);
  CountLaunch();
  int end_bit = 1;
  while ((1ull << end_bit) <= num_cells) ++end_bit;
  MLC_TRY_C(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys_s, vals, vals_s,
                                            static_cast<int>(n), 0, end_bit, stream));
  MLC_TRY_C(cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan, size16, start16,
                                          static_cast<int>(num_cells), stream));
  if (tmp_scan > tmp_bytes) tmp_bytes = tmp_scan;
  MLC_TRY_C(cudaMalloc(&tmp, tmp_bytes));
  MLC_TRY_C(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys_s, vals, vals_s,
                                            static_cast<int>(n), 0, end_bit, stream));
  CountLaunch();
  MLC_TRY_C(cudaMemsetAsync(first, 0xFF, 4 * static_cast<size_t>(num_cells), stream));
  MLC_TRY_C(cudaMemsetAsync(len, 0, 4 * static_cast<size_t>(num_cells), stream));
  cell_bounds_kernel<<<nb, 256, 0, stream>>>(keys_s, n, num_cells, first, len);
  CountLaunch();
  cell_sizes_kernel<<<cb, 256, 0, stream>>>(num_cells, first, len, dim + 1, size16);
  CountLaunch();
  MLC_TRY_C(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, size16, start16,
                                          static_cast<int>(num_cells), stream));
  CountLaunch();
  cell_info_kernel<<<cb, 256, 0, stream>>>(num_cells, start16, len, out->cell_info);
  CountLaunch();
  uint32_t last_start = 0, last_size = 0;
  MLC_TRY_C(cudaMemcpyAsync(&last_start, start16 + (num_cells - 1), 4, cudaMemcpyDeviceToHost,
                            stream));
  MLC_TRY_C(cudaMemcpyAsync(&last_size, size16 + (num_cells - 1), 4, cudaMemcpyDeviceToHost, stream));
  MLC_TRY_C(cudaStreamSynchronize(stream));
  out->list_bytes = (static_cast<size_t>(last_start) + last_size) * 16;
  MLC_TRY_C(cudaMalloc(&out->lists, out->list_bytes + 16));
  fill_lists_kernel<<<nb, 256, 0, stream>>>(keys_s, vals_s, n, num_cells, first, out->cell_info,
                                            d_desc, dim, out->lists);
  CountLaunch();
  MLC_TRY_C(cudaGetLastError());
  MLC_TRY_C(cudaStreamSynchronize(stream));
  cleanup();
#undef MLC_TRY_C
  return cudaSuccess;
}

}  // namespace mlc
