// Engine `hnsw` slot (--lc_detector_engine hnsw; loop_closure::HSNWIndexInterface,
// matching-based-loopclosure/include/matching-based-loopclosure/hnsw-index-interface.h): nearest neighbours of
// FLOAT (learned) descriptors by squared L2 distance. The reference answers approximately with an hnswlib graph
// built from all host threads at once (:44-66 — its neighbour lists are not reproducible run to run); here the
// search is EXACT: every query against every database descriptor, register-tiled like a GEMM (64 queries x 128
// database points per CTA step, 4 x 8 per thread), distances accumulated in index order with separate multiply
// and add (hnswlib's scalar L2Sqr, space_l2.h:6-20, and the oracle's loop: bit-identical), k best per query kept
// warp-distributed. A batch with few queries is split over ranges of the database so that the grid fills the
// GPU; the partial lists are merged by (distance, index). Results leave in the order the reference interface
// returns them: popped from a max-heap, i.e. DESCENDING (distance, index) (:141-151).
#include "device_index.h"

namespace mlc {
namespace {

constexpr int kBM = 64;    // queries per CTA
constexpr int kBN = 128;   // database points per step
constexpr int kBK = 16;    // dimensions per shared-memory stage
constexpr int kThreads = 256;
constexpr uint64_t kEmptyKey = 0x7f800000FFFFFFFFull;
constexpr uint32_t kFull = 0xffffffffu;

__device__ __forceinline__ uint64_t ShflKey(uint64_t key, int src) {
  const uint32_t hi = __shfl_sync(kFull, static_cast<uint32_t>(key >> 32), src);
  const uint32_t lo = __shfl_sync(kFull, static_cast<uint32_t>(key), src);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}

// Insert the keys of the lanes named in `mask` into the warp-distributed sorted list (lane r = r-th smallest).
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

// grid = (query tiles, database splits). Partial lists: out[(split * n_q + query) * k + r], ascending, missing
// entries (-1, +inf) trailing.
__global__ void __launch_bounds__(kThreads, 2)
exact_knn_kernel(const float* __restrict__ db, int64_t n_db, const float* __restrict__ q, int64_t n_q, int dim, int k,
                 int64_t db_per_split, int32_t* __restrict__ out_idx, float* __restrict__ out_dist) {
  __shared__ __align__(16) float qs[kBK][kBM + 4];   // + 4: the transposing stores hit 2, not 16, lanes per bank
  __shared__ __align__(16) float xs[kBK][kBN + 4];
  __shared__ float ds[kBM][kBN + 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ty = tid >> 4, tx = tid & 15;           // 16 x 16 threads: 4 queries x 8 points each
  const int64_t q0 = static_cast<int64_t>(blockIdx.x) * kBM;
  const int64_t db_begin = static_cast<int64_t>(blockIdx.y) * db_per_split;
  const int64_t db_end = db_begin + db_per_split < n_db ? db_begin + db_per_split : n_db;
  uint64_t held[kBM / 8];                            // this warp's 8 query rows
#pragma unroll
  for (int r = 0; r < kBM / 8; ++r) held[r] = kEmptyKey;

  for (int64_t p0 = db_begin; p0 < db_end; p0 += kBN) {
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int d0 = 0; d0 < dim; d0 += kBK) {
      __syncthreads();
      // stage the next kBK dimensions of the query tile and of the database tile (k-major in shared memory)
      for (int e = tid; e < kBM * kBK; e += kThreads) {
        const int row = e / kBK, kk = e % kBK;
        const int64_t qi = q0 + row;
        qs[kk][row] = (qi < n_q && d0 + kk < dim) ? q[qi * dim + d0 + kk] : 0.f;
      }
      for (int e = tid; e < kBN * kBK; e += kThreads) {
        const int row = e / kBK, kk = e % kBK;
        const int64_t pi = p0 + row;
        xs[kk][row] = (pi < db_end && d0 + kk < dim) ? db[pi * dim + d0 + kk] : 0.f;
      }
      __syncthreads();
      const int kmax = dim - d0 < kBK ? dim - d0 : kBK;
      for (int kk = 0; kk < kmax; ++kk) {            // dimensions in index order: the reference's summation order
        const float4 a = *reinterpret_cast<const float4*>(&qs[kk][ty * 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&xs[kk][tx * 8]);
        const float4 b1 = *reinterpret_cast<const float4*>(&xs[kk][tx * 8 + 4]);
        const float av[4] = {a.x, a.y, a.z, a.w};
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float t = __fsub_rn(av[i], bv[j]);
            acc[i][j] = __fadd_rn(acc[i][j], __fmul_rn(t, t));
          }
      }
    }
    __syncthreads();                                 // the previous step's selection is done with ds
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) ds[ty * 4 + i][tx * 8 + j] = acc[i][j];
    __syncthreads();
    // selection: warp w owns query rows 8 w .. 8 w + 7; lane l looks at points l, l + 32, l + 64, l + 96
#pragma unroll
    for (int r = 0; r < kBM / 8; ++r) {
      const int row = warp * (kBM / 8) + r;
      const uint64_t kth = ShflKey(held[r], k - 1);
#pragma unroll
      for (int c = 0; c < kBN / 32; ++c) {
        const int col = c * 32 + lane;
        const int64_t pi = p0 + col;
        const uint32_t bits = __float_as_uint(ds[row][col]);
        uint64_t key = kEmptyKey;
        if (pi < db_end && bits <= 0x7f800000u) key = (static_cast<uint64_t>(bits) << 32) | static_cast<uint32_t>(pi);
        // keys below the k-th best of the row before this step; later candidates of the same step are ordered
        // by the insertion itself
        const uint32_t m = __ballot_sync(kFull, key < kth);
        InsertCandidates(m, key, held[r], lane);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < kBM / 8; ++r) {
    const int64_t qi = q0 + warp * (kBM / 8) + r;
    if (qi < n_q && lane < k) {
      const size_t at = (static_cast<size_t>(blockIdx.y) * n_q + qi) * k + lane;
      const bool found = held[r] != kEmptyKey;
      out_idx[at] = found ? static_cast<int32_t>(static_cast<uint32_t>(held[r])) : -1;
      out_dist[at] = __uint_as_float(static_cast<uint32_t>(held[r] >> 32));
    }
  }
}

// rows of k entries ascending -> descending (the reference pops a max-heap)
__global__ void reverse_rows_kernel(const int32_t* __restrict__ in_idx, const float* __restrict__ in_dist, int64_t n,
                                    int k, int32_t* __restrict__ out_idx, float* __restrict__ out_dist) {
  const int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (t >= n * k) return;
  const int64_t row = t / k;
  const int r = static_cast<int>(t % k);
  out_idx[t] = in_idx[row * k + (k - 1 - r)];
  out_dist[t] = in_dist[row * k + (k - 1 - r)];
}

}  // namespace

// Number of database splits for a batch of n_q queries (<= 16: the merge kernel's limit).
int ExactKnnSplits(int64_t n_q, int64_t n_db, int sm_count) {
  const int64_t tiles = (n_q + kBM - 1) / kBM;
  int64_t s = (2 * static_cast<int64_t>(sm_count) + tiles - 1) / tiles;
  const int64_t max_by_db = (n_db + 4 * kBN - 1) / (4 * kBN);  // at least four steps per split
  if (s > max_by_db) s = max_by_db;
  if (s > 16) s = 16;
  if (s < 1) s = 1;
  return static_cast<int>(s);
}
size_t ExactKnnScratchBytes(int64_t n_q, int k, int splits) {
  return static_cast<size_t>(splits + 1) * n_q * k * 8 + 512;  // partial lists + the merged ascending list
}

// d_idx / d_dist: n_q x k, descending (distance, index). scratch: ExactKnnScratchBytes.
cudaError_t LaunchExactKnn(const float* d_db, int64_t n_db, const float* d_q, int64_t n_q, int dim, int k, int splits,
                           void* scratch, int32_t* d_idx, float* d_dist, cudaStream_t stream) {
  if (n_q <= 0) return cudaSuccess;
  if (k < 1 || k > 32 || k > n_db || dim < 1) return cudaErrorInvalidValue;
  const size_t list = static_cast<size_t>(n_q) * k;
  int32_t* p_idx = static_cast<int32_t*>(scratch);
  float* p_dist = reinterpret_cast<float*>(p_idx + static_cast<size_t>(splits) * list);
  int32_t* m_idx = reinterpret_cast<int32_t*>(p_dist + static_cast<size_t>(splits) * list);
  float* m_dist = reinterpret_cast<float*>(m_idx + list);
  const int64_t per_split = ((n_db + splits - 1) / splits + kBN - 1) / kBN * kBN;
  const dim3 grid(static_cast<unsigned>((n_q + kBM - 1) / kBM), static_cast<unsigned>(splits));
  exact_knn_kernel<<<grid, kThreads, 0, stream>>>(d_db, n_db, d_q, n_q, dim, k, per_split, p_idx, p_dist);
  CountLaunch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const int32_t* a_idx = p_idx;
  const float* a_dist = p_dist;
  if (splits > 1) {
    e = LaunchMergeTopk(p_idx, p_dist, splits, n_q, k, m_idx, m_dist, stream);
    if (e != cudaSuccess) return e;
    a_idx = m_idx;
    a_dist = m_dist;
  }
  reverse_rows_kernel<<<static_cast<unsigned>((list + 255) / 256), 256, 0, stream>>>(a_idx, a_dist, n_q, k, d_idx, d_dist);
  CountLaunch();
  return cudaGetLastError();
}

}  // namespace mlc
