// Kernel 2 of the loop-closure path: inverted-multi-index kNN.
//   2a  (coarse_kernels.cu)   — FindClosestWords: kd-tree searches + multi-sequence
//   2b  imi_scan_kernel       — InvertedMultiIndex::GetNNearestNeighbors: stream the visited
//                               cells' inverted lists, squared L2 in the canonical fp32 order,
//                               top-k by (distance, index) (imilib/inverted-multi-index.h:100-161,
//                               inverted-multi-index-common.h:54-72)
//   index build                — cell assignment (2a with one word), radix sort by cell, padded
//                               16-byte aligned entries (imilib/inverted-multi-index.h:77-94)
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

// Entry layout: one inverted-list entry is kWpe<DIM> 32-bit words, 16-byte aligned:
// DIM fp32 coordinates, the global descriptor index, zero padding (12 words = 48 B at DIM 10; the
// ALGORITHMIC size stays 4 * (DIM + 1) = 44 B). A lane fetches its entry with kWpe / 4 LDG.128.
template <int DIM>
constexpr int kWpe = (DIM + 1 + 3) & ~3;

// One warp per query descriptor, persistent warps. The (<= 16) visited cells' lists are flattened
// into one entry range and dealt to the lanes 64 at a time (two 32-entry windows per trip, all
// their loads issued before the first use). Top-k selection is warp-distributed: lane r holds the
// r-th smallest (distance bits, index) key seen so far. The first trip fills it with k rounds of
// warp arg-min over the (<= 64) fresh keys; later trips only touch it for keys below the current
// k-th key (ballot), inserted one by one with a shift through the lanes. Query metadata is
// software-pipelined two deep: while query i is scanned, the cell table entries of query i+1 and
// the visit list of query i+2 are already in flight, which hides the dependent chain
// visit list -> cell table -> list entries.
constexpr int kScanThreads = 256;
constexpr int kScanCtasPerSm = 4;
constexpr uint32_t kFull = 0xffffffffu;
constexpr uint32_t kInfBits = 0x7f800000u;
constexpr uint32_t kNoIndex = 0xFFFFFFFFu;

// Issue the loads of window [base, base + 32): lane l takes flattened entry base + l.
template <int DIM>
__device__ __forceinline__ bool LoadWindow(const uint4* __restrict__ lists, const uint4* seg,
                                           uint32_t base, uint32_t total, uint32_t len,
                                           uint32_t excl, int lane, uint32_t lanemask_le,
                                           uint4 (&v)[kWpe<DIM> / 4]) {
  // non-empty cells whose first entry falls into this window set bit (first entry - base)
  const uint32_t rel = excl - base;
  const uint32_t heads = __reduce_or_sync(kFull, (len > 0 && rel < 32u) ? (1u << rel) : 0u);
  const uint32_t before = __popc(__ballot_sync(kFull, len > 0 && excl < base));
  const uint32_t e = base + lane;
  const bool valid = e < total;
  if (valid) {
    const uint4 s = seg[before + __popc(heads & lanemask_le) - 1];  // {excl, start16, len}
    const uint4* p = lists + s.y + static_cast<size_t>(e - s.x) * (kWpe<DIM> / 4);
#pragma unroll
    for (int j = 0; j < kWpe<DIM> / 4; ++j) v[j] = __ldg(p + j);
  }
  return valid;
}

// (distance bits << 32 | index) of a loaded entry; NaN distances never enter a result.
template <int DIM>
__device__ __forceinline__ uint64_t EntryKey(const uint4 (&v)[kWpe<DIM> / 4], const float (&qv)[DIM],
                                             bool valid) {
  uint32_t w[kWpe<DIM>];
#pragma unroll
  for (int j = 0; j < kWpe<DIM> / 4; ++j) {
    w[4 * j + 0] = v[j].x;
    w[4 * j + 1] = v[j].y;
    w[4 * j + 2] = v[j].z;
    w[4 * j + 3] = v[j].w;
  }
  float sv[DIM];
#pragma unroll
  for (int d = 0; d < DIM; ++d) sv[d] = __uint_as_float(w[d]);
  const uint32_t bits = __float_as_uint(SquaredDistanceCanonical<DIM>(sv, qv));
  if (!valid || bits > kInfBits) return kEmptyKey;
  return (static_cast<uint64_t>(bits) << 32) | w[DIM];
}

__device__ __forceinline__ uint64_t ShflKey(uint64_t key, int src) {
  const uint32_t hi = __shfl_sync(kFull, static_cast<uint32_t>(key >> 32), src);
  const uint32_t lo = __shfl_sync(kFull, static_cast<uint32_t>(key), src);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}

// Insert the keys of the lanes named in `mask` (ascending lane order) into the warp-distributed
// sorted list `held` (lane r = r-th smallest). Warp-uniform control flow.
__device__ __forceinline__ void InsertCandidates(uint32_t mask, uint64_t key, uint64_t& held, int lane) {
  while (mask) {
    const int src = __ffs(mask) - 1;
    mask &= mask - 1;
    const uint64_t c = ShflKey(key, src);
    const uint32_t p_hi = __shfl_up_sync(kFull, static_cast<uint32_t>(held >> 32), 1);
    const uint32_t p_lo = __shfl_up_sync(kFull, static_cast<uint32_t>(held), 1);
    const uint64_t prev = (static_cast<uint64_t>(p_hi) << 32) | p_lo;
    if (c < held) held = (lane > 0 && c < prev) ? prev : c;
  }
}

template <int DIM, int KT>  // KT >= k: unroll bound of the selection rounds
__global__ void __launch_bounds__(kScanThreads, kScanCtasPerSm)
imi_scan_kernel(const float* __restrict__ q, int64_t n_q, const int32_t* __restrict__ cells, int nw,
                const uint2* __restrict__ cell_info, const uint4* __restrict__ lists, int k,
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

    uint64_t held = kEmptyKey;  // lane r: r-th smallest key so far
    for (uint32_t base = 0; base < total; base += 64) {
      uint4 va[kWpe<DIM> / 4], vb[kWpe<DIM> / 4];
      const bool in_a = LoadWindow<DIM>(lists, seg, base, total, len, excl, lane, lanemask_le, va);
      bool in_b = false;
      if (base + 32 < total)  // warp-uniform
        in_b = LoadWindow<DIM>(lists, seg, base + 32, total, len, excl, lane, lanemask_le, vb);
      uint64_t ka = EntryKey<DIM>(va, qv, in_a);
      uint64_t kb = in_b ? EntryKey<DIM>(vb, qv, true) : kEmptyKey;
      if (base == 0) {
        // k rounds of warp arg-min over the fresh keys; lane r keeps the r-th. Each lane orders
        // its two keys once, so a round only looks at the lanes' smaller keys and the winner
        // promotes its second key.
        if (kb < ka) {
          const uint64_t t = ka;
          ka = kb;
          kb = t;
        }
#pragma unroll
        for (int r = 0; r < KT; ++r) {
          if (r >= k) break;
          const uint32_t a_hi = static_cast<uint32_t>(ka >> 32), a_lo = static_cast<uint32_t>(ka);
          const uint32_t m_hi = __reduce_min_sync(kFull, a_hi);
          const uint32_t c = (a_hi == m_hi) ? a_lo : kNoIndex;
          const uint32_t m_lo = __reduce_min_sync(kFull, c);
          if (m_lo == kNoIndex) break;  // nothing left (warp-uniform)
          if (lane == r) held = (static_cast<uint64_t>(m_hi) << 32) | m_lo;
          if (c == m_lo) {
            ka = kb;
            kb = kEmptyKey;
          }
        }
      } else {
        const uint64_t kth = ShflKey(held, k - 1);
        const uint32_t ma = __ballot_sync(kFull, ka < kth);
        const uint32_t mb = __ballot_sync(kFull, kb < kth);
        InsertCandidates(ma, ka, held, lane);
        InsertCandidates(mb, kb, held, lane);
      }
    }
    __syncwarp();  // every lane is done with seg before the next query overwrites it

    if (lane < k) {  // missing neighbours: (+inf, -1), trailing
      out_idx[qi * k + lane] = static_cast<int32_t>(static_cast<uint32_t>(held));
      out_dist[qi * k + lane] = __uint_as_float(static_cast<uint32_t>(held >> 32));
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
#define MLC_SCAN(KT)                                                                             \
  imi_scan_kernel<DIM, KT><<<g, threads, 0, stream>>>(q, n_q, cells, nw, cell_info,              \
                                                      reinterpret_cast<const uint4*>(lists), k,  \
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

// =============================================================================================
// imipq engine: product-quantised residuals (imilib/inverted-multi-product-quantization-index.h,
// imilib/product-quantization.h)
// =============================================================================================
namespace {

// SquaredDistance of two runtime-length vectors in the canonical fp32 order (see
// SquaredDistanceCanonical), dim <= 8.
__device__ __forceinline__ float SquaredDistanceRuntime(const float* a, const float* b, int dim) {
  const int vec = (dim / 4) * 4;
  if (vec == 0) {
    float s = 0.f;
    for (int i = 0; i < dim; ++i) {
      const float d = __fsub_rn(a[i], b[i]);
      const float sq = __fmul_rn(d, d);
      s = (i == 0) ? sq : __fadd_rn(s, sq);
    }
    return s;
  }
  float lane[4];
  for (int l = 0; l < 4; ++l) {
    const float d = __fsub_rn(a[l], b[l]);
    lane[l] = __fmul_rn(d, d);
  }
  for (int p = 4; p < vec; p += 4)
    for (int l = 0; l < 4; ++l) {
      const float d = __fsub_rn(a[p + l], b[p + l]);
      lane[l] = __fadd_rn(lane[l], __fmul_rn(d, d));
    }
  float res = __fadd_rn(__fadd_rn(lane[0], lane[1]), __fadd_rn(lane[2], lane[3]));
  if (vec < dim) {
    float rem = 0.f;
    for (int i = vec; i < dim; ++i) {
      const float d = __fsub_rn(a[i], b[i]);
      const float sq = __fmul_rn(d, d);
      rem = (i == vec) ? sq : __fadd_rn(rem, sq);
    }
    res = __fadd_rn(res, rem);
  }
  return res;
}

// AddDescriptors of the PQ index (…-quantization-index.h:136-175): residual of each half against
// its coarse word, per component the nearest of the word's centres (first minimum wins,
// product-quantization.h:81-103). One thread per descriptor; code j lands in byte j of 12.
__global__ void pq_encode_kernel(PqParams p, const float* __restrict__ desc, const int32_t* __restrict__ cells,
                                 int64_t n, uint32_t* __restrict__ codes) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  uint32_t out[3] = {0u, 0u, 0u};
  const int32_t cell = cells[i];
  if (cell >= 0) {
    const int dim = 2 * p.sub_dim;
    const int per_word = p.half_ncomp * p.num_centers * p.dim_per_comp;
    for (int h = 0; h < 2; ++h) {
      const int w = h ? cell % p.num_words2 : cell / p.num_words2;
      const float* word = (h ? p.words2 : p.words1) + static_cast<size_t>(w) * p.sub_dim;
      const float* ctrs = (h ? p.centers2 : p.centers1) + static_cast<size_t>(w) * per_word;
      float res[8];
      for (int j = 0; j < p.sub_dim; ++j) res[j] = __fsub_rn(desc[i * dim + h * p.sub_dim + j], word[j]);
      for (int j = 0; j < p.half_ncomp; ++j) {
        int best = 0;
        float best_d = 0.f;
        for (int c = 0; c < p.num_centers; ++c) {
          const float d = SquaredDistanceRuntime(ctrs + static_cast<size_t>(j * p.num_centers + c) * p.dim_per_comp,
                                                 res + j * p.dim_per_comp, p.dim_per_comp);
          if (c == 0 || d < best_d) {
            best_d = d;
            best = c;
          }
        }
        const int at = h * p.half_ncomp + j;
        out[at >> 2] |= static_cast<uint32_t>(best) << (8 * (at & 3));
      }
    }
  }
  codes[i * 3 + 0] = out[0];
  codes[i * 3 + 1] = out[1];
  codes[i * 3 + 2] = out[2];
}

// Fallback for product quantisers whose look-up tables do not fit the fast kernel below.
// GetNNearestNeighbors of the PQ index (…-quantization-index.h:181-288): one warp per query
// descriptor. Per visited, non-empty cell (w1, w2) the lanes fill the two look-up tables
// LUT_h[component][centre] = (centre - residual_h)^2 in shared memory (FillLUT,
// product-quantization.h:107-125), then stream the cell's 16-byte entries {12 code bytes, index}:
// distance = (sequential sum over the first-half components from 0.0f) + (the same over the
// second half) (ComputeDistance, :144-152). Top-k as in imi_scan_kernel.
__global__ void __launch_bounds__(256)
imipq_scan_percell_kernel(PqParams p, const float* __restrict__ q, int64_t n_q, const int32_t* __restrict__ cells,
                  int nw, const uint2* __restrict__ cell_info, const uint4* __restrict__ lists, int k,
                  int32_t* __restrict__ out_idx, float* __restrict__ out_dist) {
  extern __shared__ float lut_s[];
  const int lane = threadIdx.x & 31;
  const int lut_size = 2 * p.half_ncomp * p.num_centers;
  float* lut = lut_s + static_cast<size_t>(threadIdx.x >> 5) * lut_size;
  const int dim = 2 * p.sub_dim;
  const int per_half = p.half_ncomp * p.num_centers;
  const int per_word = per_half * p.dim_per_comp;
  const int64_t warp_stride = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t qi = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5; qi < n_q;
       qi += warp_stride) {
    uint64_t held = kEmptyKey;
    const float* qd = q + qi * dim;
    for (int v = 0; v < nw; ++v) {
      const int32_t cell = __ldg(cells + qi * nw + v);
      if (cell < 0) continue;  // missing word
      const uint2 info = __ldg(cell_info + cell);
      if (info.y == 0) continue;  // cell not in word_index_map_
      __syncwarp();
      for (int t = lane; t < lut_size; t += 32) {
        const int h = t / per_half, r = t - h * per_half;
        const int comp = r / p.num_centers;
        const int w = h ? cell % p.num_words2 : cell / p.num_words2;
        const float* word = (h ? p.words2 : p.words1) + static_cast<size_t>(w) * p.sub_dim + comp * p.dim_per_comp;
        const float* ctr = (h ? p.centers2 : p.centers1) + static_cast<size_t>(w) * per_word +
                           static_cast<size_t>(r) * p.dim_per_comp;
        float res[8];
        for (int d = 0; d < p.dim_per_comp; ++d)
          res[d] = __fsub_rn(__ldg(qd + h * p.sub_dim + comp * p.dim_per_comp + d), __ldg(word + d));
        lut[t] = SquaredDistanceRuntime(ctr, res, p.dim_per_comp);
      }
      __syncwarp();
      const uint4* entries = lists + info.x;
      for (uint32_t base = 0; base < info.y; base += 32) {
        const uint32_t e = base + lane;
        uint64_t key = kEmptyKey;
        if (e < info.y) {
          const uint4 en = __ldg(entries + e);
          const uint32_t cw[3] = {en.x, en.y, en.z};
          float s1 = 0.0f, s2 = 0.0f;
          for (int j = 0; j < p.half_ncomp; ++j) {
            const int a1 = j, a2 = p.half_ncomp + j;
            s1 = __fadd_rn(s1, lut[j * p.num_centers + ((cw[a1 >> 2] >> (8 * (a1 & 3))) & 0xFFu)]);
            s2 = __fadd_rn(s2, lut[per_half + j * p.num_centers + ((cw[a2 >> 2] >> (8 * (a2 & 3))) & 0xFFu)]);
          }
          const uint32_t bits = __float_as_uint(__fadd_rn(s1, s2));
          if (bits <= kInfBits) key = (static_cast<uint64_t>(bits) << 32) | en.w;
        }
        const uint64_t kth = ShflKey(held, k - 1);
        InsertCandidates(__ballot_sync(kFull, key < kth), key, held, lane);
      }
    }
    if (lane < k) {
      out_idx[qi * k + lane] = static_cast<int32_t>(static_cast<uint32_t>(held));
      out_dist[qi * k + lane] = __uint_as_float(static_cast<uint32_t>(held >> 32));
    }
  }
}

// Fast path. One warp per query descriptor, persistent warps:
//   1. the (<= 16) visited cells are split into their word halves; the DISTINCT words of each half
//      get one look-up table (the reference caches LUTs per word the same way,
//      …-quantization-index.h:226-262): slot numbers by __match_any_sync, tables
//      LUT[slot][component][centre] = (centre - residual)^2 filled by all lanes together;
//   2. the lists of the visited cells are flattened into one entry range and dealt to the lanes 64
//      at a time (two windows, one LDG.128 per entry), distance = (sequential sum over the
//      first-half components from 0.0f) + (same for the second half) — the reference's order;
//   3. warp-distributed top-k as in imi_scan_kernel.

__device__ __forceinline__ float PqDistance(const PqParams& p, const float* lut1, const float* lut2, uint4 en) {
  const uint32_t cw[3] = {en.x, en.y, en.z};
  float s1 = 0.0f, s2 = 0.0f;
  for (int j = 0; j < p.half_ncomp; ++j) {
    const int a1 = j, a2 = p.half_ncomp + j;
    s1 = __fadd_rn(s1, lut1[j * p.num_centers + ((cw[a1 >> 2] >> (8 * (a1 & 3))) & 0xFFu)]);
    s2 = __fadd_rn(s2, lut2[j * p.num_centers + ((cw[a2 >> 2] >> (8 * (a2 & 3))) & 0xFFu)]);
  }
  return __fadd_rn(s1, s2);
}

__device__ __forceinline__ uint64_t PqWindowKey(const PqParams& p, const uint4* __restrict__ lists,
                                                const uint4* seg, const float* lut, int per_half, uint32_t base,
                                                uint32_t total, uint32_t len, uint32_t excl, int lane,
                                                uint32_t lanemask_le) {
  const uint32_t rel = excl - base;
  const uint32_t heads = __reduce_or_sync(kFull, (len > 0 && rel < 32u) ? (1u << rel) : 0u);
  const uint32_t before = __popc(__ballot_sync(kFull, len > 0 && excl < base));
  const uint32_t e = base + lane;
  if (e >= total) return kEmptyKey;
  const uint4 s = seg[before + __popc(heads & lanemask_le) - 1];  // {excl, start16, len, slot1 | slot2 << 8}
  const uint4 en = __ldg(lists + s.y + (e - s.x));
  const float* lut1 = lut + (s.w & 0xFFu) * per_half;
  const float* lut2 = lut + (s.w >> 8) * per_half;
  const uint32_t bits = __float_as_uint(PqDistance(p, lut1, lut2, en));
  if (bits > kInfBits) return kEmptyKey;
  return (static_cast<uint64_t>(bits) << 32) | en.w;
}

template <int KT>
__global__ void __launch_bounds__(256)
imipq_scan_kernel(PqParams p, const float* __restrict__ q, int64_t n_q, const int32_t* __restrict__ cells,
                  int nw, const uint2* __restrict__ cell_info, const uint4* __restrict__ lists, int k,
                  int32_t* __restrict__ out_idx, float* __restrict__ out_dist) {
  extern __shared__ __align__(16) unsigned char pq_smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int per_half = p.half_ncomp * p.num_centers;
  const int dim = 2 * p.sub_dim;
  // per warp: seg[16] uint4 | slot words[32] int | query[16] float | lut[2 nw * per_half] float
  const size_t per_warp = 16 * sizeof(uint4) + 32 * 4 + 16 * 4 + static_cast<size_t>(2 * nw) * per_half * 4;
  unsigned char* mine = pq_smem + warp * per_warp;
  uint4* seg = reinterpret_cast<uint4*>(mine);
  int* slot_word = reinterpret_cast<int*>(mine + 16 * sizeof(uint4));
  float* qs = reinterpret_cast<float*>(mine + 16 * sizeof(uint4) + 32 * 4);
  float* lut = qs + 16;
  const uint32_t lanemask_lt = (1u << lane) - 1u;
  const uint32_t lanemask_le = lanemask_lt | (1u << lane);
  const int64_t warp_stride = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t qi = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5; qi < n_q;
       qi += warp_stride) {
    // ---- visited cells of this query, one per lane
    int32_t c = -1;
    if (lane < nw) c = __ldg(cells + qi * nw + lane);
    uint2 info = make_uint2(0u, 0u);
    if (c >= 0) info = __ldg(cell_info + c);
    const uint32_t len = info.y;
    const bool used = len > 0;  // cell present in word_index_map_
    const int w1 = used ? c / p.num_words2 : -1 - lane;
    const int w2 = used ? c - w1 * p.num_words2 : -1 - lane;
    // ---- one LUT slot per distinct word of each half
    const uint32_t same1 = __match_any_sync(kFull, w1), same2 = __match_any_sync(kFull, w2);
    const int lead1 = __ffs(same1) - 1, lead2 = __ffs(same2) - 1;
    const uint32_t leaders1 = __ballot_sync(kFull, used && lead1 == lane);
    const uint32_t leaders2 = __ballot_sync(kFull, used && lead2 == lane);
    const int n1 = __popc(leaders1), n2 = __popc(leaders2);
    const int slot1 = __popc(leaders1 & ((1u << lead1) - 1u));
    const int slot2 = n1 + __popc(leaders2 & ((1u << lead2) - 1u));
    __syncwarp();  // the previous query's tables / segments are no longer read
    if (used && lead1 == lane) slot_word[slot1] = w1;
    if (used && lead2 == lane) slot_word[slot2] = w2;
    if (lane < dim) qs[lane] = __ldg(q + qi * dim + lane);
    // ---- flattened entry range
    uint32_t incl = len;
#pragma unroll
    for (int o = 1; o < kMaxWords; o <<= 1) {
      const uint32_t v = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl += v;
    }
    const uint32_t total = __shfl_sync(kFull, incl, kMaxWords - 1);
    const uint32_t excl = incl - len;
    const uint32_t nonempty = __ballot_sync(kFull, used);
    if (used)
      seg[__popc(nonempty & lanemask_lt)] =
          make_uint4(excl, info.x, len, static_cast<uint32_t>(slot1) | (static_cast<uint32_t>(slot2) << 8));
    __syncwarp();
    // ---- look-up tables: entry t = (slot, component, centre)
    const int entries = (n1 + n2) * per_half;
    if (p.vector_lut) {
      // one scalar per component (dim_per_comp == 1), centres 16-byte aligned and a multiple of four
      // per component: a lane fills four consecutive centres of one (slot, component) per step
      for (int t4 = lane * 4; t4 < entries; t4 += 128) {
        const int slot = __umulhi(static_cast<uint32_t>(t4), p.magic_per_half);
        const int r = t4 - slot * per_half;
        const int comp = __umulhi(static_cast<uint32_t>(r), p.magic_centers);
        const int h = slot >= n1 ? 1 : 0;
        const int w = slot_word[slot];
        const float res = __fsub_rn(qs[h * p.sub_dim + comp], __ldg((h ? p.words2 : p.words1) + static_cast<size_t>(w) * p.sub_dim + comp));
        const float4 c4 = __ldg(reinterpret_cast<const float4*>((h ? p.centers2 : p.centers1) + static_cast<size_t>(w) * per_half + r));
        float4 o;
        float d;
        d = __fsub_rn(c4.x, res); o.x = __fmul_rn(d, d);
        d = __fsub_rn(c4.y, res); o.y = __fmul_rn(d, d);
        d = __fsub_rn(c4.z, res); o.z = __fmul_rn(d, d);
        d = __fsub_rn(c4.w, res); o.w = __fmul_rn(d, d);
        *reinterpret_cast<float4*>(lut + t4) = o;
      }
    } else {
      for (int t = lane; t < entries; t += 32) {
        const int slot = __umulhi(static_cast<uint32_t>(t), p.magic_per_half);
        const int r = t - slot * per_half;
        const int comp = __umulhi(static_cast<uint32_t>(r), p.magic_centers);
        const int h = slot >= n1 ? 1 : 0;
        const int w = slot_word[slot];
        const float* word = (h ? p.words2 : p.words1) + static_cast<size_t>(w) * p.sub_dim + comp * p.dim_per_comp;
        const float* ctr = (h ? p.centers2 : p.centers1) + (static_cast<size_t>(w) * per_half + r) * p.dim_per_comp;
        float res[8];
        for (int d = 0; d < p.dim_per_comp; ++d)
          res[d] = __fsub_rn(qs[h * p.sub_dim + comp * p.dim_per_comp + d], __ldg(word + d));
        lut[t] = SquaredDistanceRuntime(ctr, res, p.dim_per_comp);
      }
    }
    __syncwarp();
    // ---- scan
    uint64_t held = kEmptyKey;
    for (uint32_t base = 0; base < total; base += 64) {
      uint64_t ka = PqWindowKey(p, lists, seg, lut, per_half, base, total, len, excl, lane, lanemask_le);
      uint64_t kb = kEmptyKey;
      if (base + 32 < total)
        kb = PqWindowKey(p, lists, seg, lut, per_half, base + 32, total, len, excl, lane, lanemask_le);
      if (base == 0) {
        if (kb < ka) {
          const uint64_t t = ka;
          ka = kb;
          kb = t;
        }
#pragma unroll
        for (int r = 0; r < KT; ++r) {
          if (r >= k) break;
          const uint32_t a_hi = static_cast<uint32_t>(ka >> 32), a_lo = static_cast<uint32_t>(ka);
          const uint32_t m_hi = __reduce_min_sync(kFull, a_hi);
          const uint32_t cc = (a_hi == m_hi) ? a_lo : kNoIndex;
          const uint32_t m_lo = __reduce_min_sync(kFull, cc);
          if (m_lo == kNoIndex) break;
          if (lane == r) held = (static_cast<uint64_t>(m_hi) << 32) | m_lo;
          if (cc == m_lo) {
            ka = kb;
            kb = kEmptyKey;
          }
        }
      } else {
        const uint64_t kth = ShflKey(held, k - 1);
        const uint32_t ma = __ballot_sync(kFull, ka < kth);
        const uint32_t mb = __ballot_sync(kFull, kb < kth);
        InsertCandidates(ma, ka, held, lane);
        InsertCandidates(mb, kb, held, lane);
      }
    }
    if (lane < k) {
      out_idx[qi * k + lane] = static_cast<int32_t>(static_cast<uint32_t>(held));
      out_dist[qi * k + lane] = __uint_as_float(static_cast<uint32_t>(held >> 32));
    }
  }
}

}  // namespace

cudaError_t LaunchPqEncode(const PqParams& p, const float* desc, const int32_t* cells, int64_t n,
                           uint32_t* codes, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  pq_encode_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, stream>>>(p, desc, cells, n, codes);
  CountLaunch();
  return cudaGetLastError();
}

cudaError_t LaunchImipqScan(const PqParams& p_in, const float* q, int64_t n_q, const int32_t* cells, int nw,
                            const uint2* cell_info, const uint32_t* lists, int k, int32_t* out_idx,
                            float* out_dist, int sm_count, cudaStream_t stream) {
  if (n_q <= 0) return cudaSuccess;
  if (k > 16 || nw > kMaxWords) return cudaErrorInvalidValue;
  PqParams p = p_in;
  const int per_half = p.half_ncomp * p.num_centers;
  p.magic_per_half = static_cast<uint32_t>((0x100000000ull + per_half - 1) / per_half);  // t / per_half, t < 2^16
  p.magic_centers = static_cast<uint32_t>((0x100000000ull + p.num_centers - 1) / p.num_centers);
  p.vector_lut = (p.dim_per_comp == 1 && p.num_centers % 4 == 0 &&
                  reinterpret_cast<uintptr_t>(p.centers1) % 16 == 0 && reinterpret_cast<uintptr_t>(p.centers2) % 16 == 0)
                     ? 1
                     : 0;
  int64_t blocks = (n_q * 32 + 255) / 256;
  const uint4* lists4 = reinterpret_cast<const uint4*>(lists);
  const size_t fast_smem = static_cast<size_t>(256 / 32) *
                           (16 * sizeof(uint4) + 32 * 4 + 16 * 4 + static_cast<size_t>(2 * nw) * per_half * 4);
  if (fast_smem <= 100 * 1024 && 2 * p.sub_dim <= 16) {
    const int per_sm = fast_smem <= 56 * 1024 ? 4 : 2;
    const int64_t cap = static_cast<int64_t>(sm_count) * per_sm;
    if (blocks > cap) blocks = cap;
    const unsigned g = static_cast<unsigned>(blocks);
#define MLC_PQ_SCAN(KT)                                                                                   \
  do {                                                                                                    \
    cudaError_t e = cudaFuncSetAttribute(imipq_scan_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         static_cast<int>(fast_smem));                                    \
    if (e != cudaSuccess) return e;                                                                       \
    imipq_scan_kernel<KT><<<g, 256, fast_smem, stream>>>(p, q, n_q, cells, nw, cell_info, lists4, k,      \
                                                         out_idx, out_dist);                              \
  } while (0)
    if (k <= 1) MLC_PQ_SCAN(1);
    else if (k <= 2) MLC_PQ_SCAN(2);
    else if (k <= 4) MLC_PQ_SCAN(4);
    else if (k <= 6) MLC_PQ_SCAN(6);
    else if (k <= 8) MLC_PQ_SCAN(8);
    else if (k <= 10) MLC_PQ_SCAN(10);
    else MLC_PQ_SCAN(16);
#undef MLC_PQ_SCAN
    CountLaunch();
    return cudaGetLastError();
  }
  // large quantisers: one cell at a time with a single pair of tables
  const size_t smem = static_cast<size_t>(256 / 32) * 2 * per_half * sizeof(float);
  if (smem > 96 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(imipq_scan_percell_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  const int64_t cap = static_cast<int64_t>(sm_count) * 4;
  if (blocks > cap) blocks = cap;
  imipq_scan_percell_kernel<<<static_cast<unsigned>(blocks), 256, smem, stream>>>(
      p, q, n_q, cells, nw, cell_info, lists4, k, out_idx, out_dist);
  CountLaunch();
  return cudaGetLastError();
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

// cells of this shard's rows -> sort keys; rows without a cell (a coarse word missing inside the
// search radius) get the key `num_cells` (sorted to the end and ignored).
__global__ void shard_keys_kernel(const int32_t* __restrict__ cells, int64_t n, uint32_t num_cells,
                                  uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const int32_t c = cells[i];
  keys[i] = c >= 0 ? static_cast<uint32_t>(c) : num_cells;
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
// Scatter descriptor i into its entry slot (16-byte aligned, padded) of its cell.
__global__ void fill_lists_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                  int64_t n, uint32_t num_cells, const uint32_t* __restrict__ first,
                                  const uint2* __restrict__ info, const float* __restrict__ desc,
                                  const int32_t* __restrict__ gidx, int dim,
                                  uint32_t* __restrict__ lists) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const uint32_t c = keys[i];
  if (c >= num_cells) return;
  const uint32_t within = static_cast<uint32_t>(i) - first[c];
  const uint2 ci = info[c];
  const int wpe = (dim + 1 + 3) & ~3;
  uint32_t* w = lists + (static_cast<size_t>(ci.x) << 2) + static_cast<size_t>(within) * wpe;
  const uint32_t row = vals[i];
  const float* src = desc + static_cast<size_t>(row) * dim;
  for (int d = 0; d < dim; ++d) w[d] = __float_as_uint(src[d]);
  w[dim] = static_cast<uint32_t>(gidx[row]);
  for (int d = dim + 1; d < wpe; ++d) w[d] = 0u;
}

}  // namespace

#define MLC_CUDA_TRY(x)                \
  do {                                 \
    cudaError_t e__ = (x);             \
    if (e__ != cudaSuccess) return e__; \
  } while (0)

cudaError_t BuildImiLists(const int32_t* d_cells, const float* d_desc, const int32_t* d_gidx, int64_t n,
                          int dim, uint32_t num_cells, DeviceLists* out, cudaStream_t stream) {
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
  shard_keys_kernel<<<nb, 256, 0, stream>>>(d_cells, n, num_cells, keys, vals);
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
  cell_sizes_kernel<<<cb, 256, 0, stream>>>(num_cells, first, len, (dim + 1 + 3) & ~3, size16);
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
                                            d_desc, d_gidx, dim, out->lists);
  CountLaunch();
  MLC_TRY_C(cudaGetLastError());
  MLC_TRY_C(cudaStreamSynchronize(stream));
  cleanup();
#undef MLC_TRY_C
  return cudaSuccess;
}

}  // namespace mlc
