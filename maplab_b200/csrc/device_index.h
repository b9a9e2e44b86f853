// Device-side data structures and kernel launchers of the B200 loop-closure path.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>

#include "vocabulary.h"

namespace mlc {

// Launch counter (bench.py reports it as gpu_launches).
extern std::atomic<uint64_t> g_kernel_launches;
inline void CountLaunch() { g_kernel_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- kernel 2a -------------------------------------------------------------------------------
// Both kd-trees live in one packed, 16-byte aligned device blob so that a CTA can stage them in
// shared memory with one vectorised copy.
struct CoarseParams {
  const void* packed;
  uint32_t packed_bytes;
  uint32_t off_nodes1, off_buckets1, off_cloud1, off_nodes2, off_buckets2, off_cloud2;
  const KdNodeDev *nodes1, *nodes2;  // global-memory views of the same blob
  const int32_t *buckets1, *buckets2;
  const float *cloud1, *cloud2;
  int sub_dim, num_words1, num_words2;
  float max_radius2, max_error2;
  int stage_in_smem;
};
// scratch: CoarseScratchBytes(p, n, num_words) bytes for the per-half sorted word lists.
size_t CoarseScratchBytes(const CoarseParams& p, int64_t n, int num_words);
cudaError_t LaunchCoarseWords(const CoarseParams& p, const float* d_q, int64_t n, int num_words,
                              int32_t* d_cells, void* scratch, int sm_count, cudaStream_t stream);

// ---- kernel 2b -------------------------------------------------------------------------------
// Inverted lists in HBM. Cell c owns `len` consecutive entries starting at byte 16 * start16 of
// `lists`. One entry = `dim` payload words, the global descriptor index, zero padding up to a
// multiple of 16 bytes: imi = 10 fp32 coordinates + index (48 B stored, 44 B algorithmic); imipq =
// 12 code bytes + index (16 B stored; 9 B algorithmic at 10 components x 16 centres).
struct DeviceLists {
  uint2* cell_info = nullptr;  // {start16, len} per cell
  uint32_t* lists = nullptr;
  size_t list_bytes = 0;
  uint32_t num_cells = 0;
  int dim = 0;
  void Free() {
    if (cell_info) cudaFree(cell_info);
    if (lists) cudaFree(lists);
    cell_info = nullptr;
    lists = nullptr;
    list_bytes = 0;
  }
};
// d_cells / d_desc / d_gidx: one row per descriptor of THIS shard (cell, payload, global descriptor
// index); rows with a negative cell are not indexed.
cudaError_t BuildImiLists(const int32_t* d_cells, const float* d_desc, const int32_t* d_gidx, int64_t n,
                          int dim, uint32_t num_cells, DeviceLists* out, cudaStream_t stream);
cudaError_t LaunchImiScan(int dim, const float* q, int64_t n_q, const int32_t* cells, int nw,
                          const uint2* cell_info, const uint32_t* lists, int k, int32_t* out_idx,
                          float* out_dist, int sm_count, cudaStream_t stream);
// imipq engine (imilib/inverted-multi-product-quantization-index.h): device views of the coarse
// words (column-major sub_dim x W) and of the per-word residual quantiser centres
// ([word][component][centre][dim_per_comp], as serialised).
struct PqParams {
  const float *words1 = nullptr, *words2 = nullptr, *centers1 = nullptr, *centers2 = nullptr;
  int sub_dim = 0, half_ncomp = 0, dim_per_comp = 0, num_centers = 0, num_words1 = 0, num_words2 = 0;
  uint32_t magic_per_half = 0, magic_centers = 0;  // ceil(2^32 / d): t / d == __umulhi(t, magic) for small t
  int vector_lut = 0;  // LUT fill with float4 centre loads (dim_per_comp == 1, aligned, centres % 4 == 0)
};
cudaError_t LaunchPqEncode(const PqParams& p, const float* desc, const int32_t* cells, int64_t n,
                           uint32_t* codes, cudaStream_t stream);
cudaError_t LaunchImipqScan(const PqParams& p, const float* q, int64_t n_q, const int32_t* cells, int nw,
                            const uint2* cell_info, const uint32_t* lists, int k, int32_t* out_idx,
                            float* out_dist, int sm_count, cudaStream_t stream);
cudaError_t LaunchScanEntries(const int32_t* cells, int64_t n_visits, const uint2* cell_info,
                              unsigned long long* d_total, cudaStream_t stream);
cudaError_t LaunchMergeTopk(const int32_t* idx_lists, const float* dist_lists, int num_lists,
                            int64_t n_q, int k, int32_t* out_idx, float* out_dist,
                            cudaStream_t stream);

// ---- engine `hnsw` slot: exact search over float descriptors (exact_knn_kernel.cu) ---------------
int ExactKnnSplits(int64_t n_q, int64_t n_db, int sm_count);
size_t ExactKnnScratchBytes(int64_t n_q, int k, int splits);
cudaError_t LaunchExactKnn(const float* d_db, int64_t n_db, const float* d_q, int64_t n_q, int dim, int k, int splits,
                           void* scratch, int32_t* d_idx, float* d_dist, cudaStream_t stream);

// ---- kernel 1 --------------------------------------------------------------------------------
// B operand image of the projection GEMM: kProjNPad rows x (kp_padded bytes), already arranged in
// the shared-memory core-matrix layout the kernel uses (see projection_kernel.cu).
struct ProjectionDevice {
  int8_t* b_image[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // device; index = 16-byte chunks per descriptor
  const FixedProjection* fp = nullptr;  // host copy the images are built from (outlives this struct's users)
  int dim = 0;
  int kp = 0;               // descriptor bits consumed
  int32_t shift[16] = {0};  // per output dim
};
cudaError_t BuildProjectionDevice(const FixedProjection& fp, ProjectionDevice* out);
cudaError_t LaunchProjection(ProjectionDevice& pd, const uint8_t* d_bits, int bytes_per_desc,
                             int64_t n, float* d_out, int sm_count, cudaStream_t stream);

}  // namespace mlc
