// Kernel 1: binary descriptor projection (descriptor_projection::ProjectDescriptorBlock,
// descriptor-projection/src/descriptor-projection.cc:15-50; bit order of DescriptorToEigenMatrix,
// descriptor-projection/include/descriptor-projection/descriptor-projection.h:92-115).
//
// Y[n][d] = sum_k P[d][k] * bit_k(desc n) as an exact integer GEMM on the 5th-gen tensor cores:
//   A (128 x K)  = the descriptor bytes, once per bit plane i = 0..7, each copy masked with 1 << i: the u8
//                  element of (plane i, byte j) is (byte_j & (1 << i)) in {0, 2^i}. K = 8 planes x bytes per
//                  descriptor (512 for FREAK), written by the producer warps straight into the UMMA K-major
//                  core-matrix layout — one AND per 4 bytes and plane instead of a bit-to-byte expansion;
//   B (64 x K)   = the fixed-point projection matrix: the element of (plane i, byte j) is
//                  P_int[d][8 j + i] * 2^(7 - i), so that a * b = 2^7 * bit * P_int for every plane, split into
//                  5 balanced base-256 s8 digits (row 5 d + t = digit t of output dim d; other rows zero);
//   D (128 x 64) = s32 accumulators in TMEM (tcgen05.mma kind::i8), double buffered;
//   epilogue     = tcgen05.ld, recombine the digits in int64, ONE rounding to fp32 (the 2^-7 and the
//                  fixed-point scale are exact powers of two), bulk store.
// Data movement is TMA: the raw 128-descriptor tile (contiguous in global memory) arrives with one
// cp.async.bulk into a 4-stage shared-memory ring (mbarrier complete_tx), the projected tile leaves with one
// cp.async.bulk store from shared memory.
// Persistent CTAs (one per SM), warp-specialised: 4 epilogue warps, 1 MMA/TMEM warp, 8 producer warps,
// 1 TMA load warp; mbarrier pipelines between the roles.
#include <cstdlib>
#include <mutex>

#include "device_index.h"
#include "ptx.cuh"

namespace mlc {
namespace {

constexpr int kTileM = 128;
constexpr int kMaxDescBytes = 64;               // descriptor bytes (<= 512 bits)
constexpr int kMaxK = 8 * kMaxDescBytes;        // 512 u8 elements per row
constexpr int kASlotBytes = kTileM * kMaxK;     // 64 KB
constexpr int kASlots = 2;
constexpr int kALbo = 16 * 128;                 // K-adjacent core matrices: 16 row groups apart
constexpr int kASbo = 128;                      // M-adjacent 8-row groups
constexpr int kN = 64;                          // UMMA N: >= dim * digits, multiple of 16
constexpr int kDigits = 5;                      // |P_int * 2^7| <= 2^33: five balanced base-256 digits
constexpr int kBRowGroups = kN / 8;             // 8
constexpr int kBLbo = kBRowGroups * 128;        // 1024
constexpr int kBSbo = 128;
constexpr int kBBytes = (kMaxK / 16) * kBLbo;   // 32 KB
constexpr int kRawStages = 4;
constexpr int kRawBytes = kTileM * kMaxDescBytes;  // 8 KB
constexpr int kEpilogueWarps = 4;
constexpr int kMmaWarp = 4;
constexpr int kProducerWarps = 8;
constexpr int kRowGroupsPerWarp = (kTileM / 8) / kProducerWarps;  // 8-row core-matrix groups per producer warp
constexpr int kFirstProducerWarp = 5;
constexpr int kLoadWarp = kFirstProducerWarp + kProducerWarps;    // 13
constexpr int kThreads = (kLoadWarp + 1) * 32;                    // 448
constexpr int kAccStages = 2;
constexpr int kAccCols = 64;                    // TMEM columns per accumulator stage
constexpr int kTmemCols = 128;
constexpr int kMaxDim = kN / kDigits;           // 12


struct Smem {
  alignas(128) uint8_t a[kASlots][kASlotBytes];
  alignas(128) int8_t b[kBBytes];
  alignas(128) uint8_t raw[kRawStages][kRawBytes];
  alignas(128) float out[kAccStages][kTileM * kMaxDim];
  alignas(8) uint64_t full[kASlots];
  uint64_t empty[kASlots];
  uint64_t raw_full[kRawStages];
  uint64_t raw_empty[kRawStages];
  uint64_t acc_full[kAccStages];
  uint64_t acc_empty[kAccStages];
  uint32_t tmem_base;
};

struct ProjArgs {
  const uint8_t* bits;
  const int8_t* b_image;
  float* out;
  int64_t n;
  int bytes_per_desc;  // 16-byte multiple, <= 64
  int k_steps;         // MMAs per tile = (8 planes * bytes_per_desc) / 32
  int dim;
  int bulk_io;         // pointers are 16-byte aligned: tiles move with cp.async.bulk
  float scale[kMaxDim];  // 2^-(shift[d] + 7)
};

// u8 x s8 instruction descriptor: D = s32, A = u8, B = s8, both K-major, N = 64, M = 128.
constexpr uint32_t kIdesc = (2u << 4) | (0u << 7) | (1u << 10) |
                            (static_cast<uint32_t>(kN >> 3) << 17) |
                            (static_cast<uint32_t>(kTileM >> 4) << 24);

__global__ void __launch_bounds__(kThreads, 1) projection_kernel(ProjArgs args) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem& s = *reinterpret_cast<Smem*>(
      smem_raw + ((128u - (ptx::smem_u32(smem_raw) & 127u)) & 127u));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t num_tiles = (args.n + kTileM - 1) / kTileM;
  const int nkq = args.bytes_per_desc >> 4;  // 16-byte chunks per descriptor
  const uint32_t tile_bytes = static_cast<uint32_t>(kTileM) * args.bytes_per_desc;

  // ---- one-time setup ----
  for (int i = threadIdx.x * 16; i < kBBytes; i += kThreads * 16)
    *reinterpret_cast<uint4*>(s.b + i) = *reinterpret_cast<const uint4*>(args.b_image + i);
  if (warp == kMmaWarp) {
    if (lane == 0) {
      for (int i = 0; i < kASlots; ++i) {
        ptx::mbar_init(&s.full[i], kProducerWarps * 32);
        ptx::mbar_init(&s.empty[i], 1);
      }
      for (int i = 0; i < kRawStages; ++i) {
        ptx::mbar_init(&s.raw_full[i], 1);
        ptx::mbar_init(&s.raw_empty[i], kProducerWarps * 32);
      }
      for (int i = 0; i < kAccStages; ++i) {
        ptx::mbar_init(&s.acc_full[i], 1);
        ptx::mbar_init(&s.acc_empty[i], kEpilogueWarps * 32);
      }
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc(&s.tmem_base, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async_smem();  // B image written through the generic proxy
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = s.tmem_base;

  if (warp == kLoadWarp) {
    // ================= TMA loads: one bulk copy per 128-descriptor tile =================
    if (ptx::elect_one()) {
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const uint32_t st = it % kRawStages;
        const uint32_t phase = (it / kRawStages) & 1u;
        ptx::mbar_wait(&s.raw_empty[st], phase ^ 1u);
        const int64_t row0 = tile * kTileM;
        const int64_t rows = (args.n - row0) < kTileM ? (args.n - row0) : kTileM;
        const uint32_t bytes = static_cast<uint32_t>(rows) * args.bytes_per_desc;  // multiple of 16
        const uint8_t* src = args.bits + row0 * args.bytes_per_desc;
        if (args.bulk_io) {
          ptx::mbar_arrive_expect_tx(&s.raw_full[st], bytes);
          ptx::bulk_load(s.raw[st], src, bytes, &s.raw_full[st]);
        } else {
          // unaligned caller buffer: plain 4-byte copies by this thread (correct, slow; never the bench path)
          for (uint32_t i = 0; i < bytes; ++i) s.raw[st][i] = src[i];
          ptx::mbar_arrive(&s.raw_full[st]);
        }
      }
    }
    __syncwarp();
  } else if (warp >= kFirstProducerWarp) {
    // ================= plane-mask producers =================
    const int pw = warp - kFirstProducerWarp;
    const int r = lane & 7;    // row inside the 8-row core matrix
    const int kq = lane >> 3;  // 16-byte chunk of the descriptor
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const uint32_t slot = it % kASlots;
      const uint32_t phase = (it / kASlots) & 1u;
      const uint32_t st = it % kRawStages;
      const uint32_t raw_phase = (it / kRawStages) & 1u;
      const int64_t rows_left = args.n - tile * kTileM;
      ptx::mbar_wait(&s.raw_full[st], raw_phase);
      uint4 w[kRowGroupsPerWarp];
#pragma unroll
      for (int h = 0; h < kRowGroupsPerWarp; ++h) {
        const int row = (pw + h * kProducerWarps) * 8 + r;
        w[h] = make_uint4(0, 0, 0, 0);
        if (row < rows_left && kq < nkq)
          w[h] = *reinterpret_cast<const uint4*>(s.raw[st] + row * args.bytes_per_desc + kq * 16);
      }
      ptx::mbar_wait(&s.empty[slot], phase ^ 1u);
      if (kq < nkq) {
#pragma unroll
        for (int h = 0; h < kRowGroupsPerWarp; ++h) {
          const int rg = pw + h * kProducerWarps;
          uint8_t* dst = s.a[slot] + kq * kALbo + rg * kASbo + r * 16;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint32_t m = 0x01010101u << i;
            *reinterpret_cast<uint4*>(dst + i * nkq * kALbo) =
                make_uint4(w[h].x & m, w[h].y & m, w[h].z & m, w[h].w & m);
          }
        }
      }
      // the stores above consumed w[]: the ld.shared of the raw stage have returned, the TMA may overwrite it
      // (an arrive right behind the loads would issue while they are still in flight)
      ptx::mbar_arrive(&s.raw_empty[st]);
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(&s.full[slot]);
      (void)tile_bytes;
    }
  } else if (warp == kMmaWarp) {
    // ================= MMA issuer =================
    if (ptx::elect_one()) {
      const uint32_t b_addr = ptx::smem_u32(s.b);
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const uint32_t slot = it % kASlots;
        const uint32_t phase = (it / kASlots) & 1u;
        const uint32_t acc = it % kAccStages;
        const uint32_t acc_phase = (it / kAccStages) & 1u;
        ptx::mbar_wait(&s.acc_empty[acc], acc_phase ^ 1u);
        ptx::mbar_wait(&s.full[slot], phase);
        ptx::tc_fence_after();
        const uint32_t a_addr = ptx::smem_u32(s.a[slot]);
        const uint32_t d_tmem = tmem_base + acc * kAccCols;
        for (int ks = 0; ks < args.k_steps; ++ks) {
          const uint64_t a_desc = ptx::make_smem_desc(a_addr + ks * 2 * kALbo, kALbo, kASbo);
          const uint64_t b_desc = ptx::make_smem_desc(b_addr + ks * 2 * kBLbo, kBLbo, kBSbo);
          ptx::mma_i8_ss(d_tmem, a_desc, b_desc, kIdesc, ks > 0 ? 1u : 0u);
        }
        ptx::tc_commit(&s.empty[slot]);     // smem slot reusable once the MMAs have read it
        ptx::tc_commit(&s.acc_full[acc]);   // accumulator ready for the epilogue
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue =================
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it % kAccStages;
      const uint32_t acc_phase = (it / kAccStages) & 1u;
      ptx::mbar_wait(&s.acc_full[acc], acc_phase);
      ptx::tc_fence_after();
      uint32_t v[kN];
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + acc * kAccCols;
      // x8 loads, which ptxas interleaves with the arithmetic below: with ONE epilogue group that hides more than
      // the later release of the accumulator costs (22.9 against 20.0 G descriptors/s for the batched form of
      // projection_tmem_kernel)
#pragma unroll
      for (int c = 0; c < kN / 8; ++c) {
        uint32_t t8[8];
        ptx::tmem_ld_32x32b_x8(taddr + c * 8, t8);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[c * 8 + i] = t8[i];
      }
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&s.acc_empty[acc]);
      float y[kMaxDim];
#pragma unroll
      for (int d = 0; d < kMaxDim; ++d) {
        long long accv = 0;
#pragma unroll
        for (int j = kDigits - 1; j >= 0; --j)
          accv = accv * 256 + static_cast<int>(v[d * kDigits + j]);
        y[d] = __ll2float_rn(accv) * args.scale[d];  // exact power-of-two scaling
      }
      const int64_t row0 = tile * kTileM;
      const int64_t rows_left = args.n - row0;
      const int row_in_tile = warp * 32 + lane;
      const bool bulk = args.bulk_io && rows_left >= kTileM;  // whole tile: leaves with one bulk store
      if (bulk) {
        // the staging buffer of this accumulator stage was handed to the TMA two tiles ago
        if (warp == 0 && lane == 0) ptx::bulk_wait_read<1>();
        ptx::named_barrier_sync(1, kEpilogueWarps * 32);
        float* o = s.out[acc] + row_in_tile * args.dim;
#pragma unroll
        for (int d = 0; d < kMaxDim; ++d)
          if (d < args.dim) o[d] = y[d];
        ptx::fence_proxy_async_smem();
        ptx::named_barrier_sync(1, kEpilogueWarps * 32);
        if (warp == 0 && lane == 0) {
          ptx::bulk_store(args.out + row0 * args.dim, s.out[acc], static_cast<uint32_t>(kTileM) * args.dim * 4u);
          ptx::bulk_commit();
        }
      } else if (row_in_tile < rows_left) {
        float* o = args.out + (row0 + row_in_tile) * args.dim;
#pragma unroll
        for (int d = 0; d < kMaxDim; ++d)
          if (d < args.dim) o[d] = y[d];
      }
    }
    if (warp == 0 && lane == 0) ptx::bulk_wait_all();  // the bulk stores read shared memory until they retire
  }

  // ---- teardown ----
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Variant with the A operand in TENSOR MEMORY (tcgen05.mma with [a_tmem]): the plane-masked bytes never touch
// shared memory. The smem-A kernel above moves 64 KB of A per 128-descriptor tile through shared memory twice
// (producer stores, tensor-core reads); here a producer thread owns one descriptor ROW (TMEM lane), ANDs its
// 16 words with the plane masks and writes 16 columns per plane with tcgen05.st. K is padded to 64 bytes per
// plane (descriptors shorter than 64 bytes leave zero columns), so the B image is always the 64-byte one.
// TMEM columns: accumulators 2 x 64, A tiles 3 x 128 (the producers of tile t + 2 do not wait for the MMAs of tile t).
// ---------------------------------------------------------------------------------------------------------
constexpr int kTmemColsA = 512;
constexpr int kAColsPerTile = 128;   // 8 planes x 16 columns (4 u8 per column)
constexpr int kATmemBase = kAccStages * kAccCols;  // 128
constexpr int kASlotsT = 3;          // A tiles in flight: 128 + 3 x 128 = 512 columns, all of TMEM
// Warp roles: two epilogue groups of 4 warps (group g drains accumulator g = the tiles with it % 2 == g; one
// group alone is a serial chain of ~1 500 cycles per tile — tcgen05.ld, int64 recombination, staging, bulk
// store — and paced the kernel), MMA issuer, TMA loader, 8 producers.
constexpr int kEpiGroupsT = 2;
constexpr int kMmaWarpT = 4 * kEpiGroupsT;                 // 8
constexpr int kLoadWarpT = kMmaWarpT + 1;                  // 9
constexpr int kFirstProducerWarpT = kLoadWarpT + 1;        // 10
constexpr int kThreadsT = (kFirstProducerWarpT + kProducerWarps) * 32;  // 576
constexpr int kRawStagesT = 8;       // 64 KB of descriptor tiles in flight per SM (4 stages cap the loads at 3 TB/s)
static_assert(kAccStages == kEpiGroupsT, "one accumulator per epilogue group");

struct SmemT {
  alignas(128) int8_t b[kBBytes];
  alignas(128) uint8_t raw[kRawStagesT][kRawBytes];
  alignas(128) float out[kEpiGroupsT][kTileM * kMaxDim];
  alignas(8) uint64_t full[kASlotsT];
  uint64_t empty[kASlotsT];
  uint64_t raw_full[kRawStagesT];
  uint64_t raw_empty[kRawStagesT];
  uint64_t acc_full[kAccStages];
  uint64_t acc_empty[kAccStages];
  uint32_t tmem_base;
  uint32_t zero;   // see the epilogue
};

__global__ void __launch_bounds__(kThreadsT, 1) projection_tmem_kernel(ProjArgs args) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SmemT& s = *reinterpret_cast<SmemT*>(
      smem_raw + ((128u - (ptx::smem_u32(smem_raw) & 127u)) & 127u));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t num_tiles = (args.n + kTileM - 1) / kTileM;
  const int nkq = args.bytes_per_desc >> 4;

  for (int i = threadIdx.x * 16; i < kBBytes; i += kThreadsT * 16)
    *reinterpret_cast<uint4*>(s.b + i) = *reinterpret_cast<const uint4*>(args.b_image + i);
  if (warp == kMmaWarpT) {
    if (lane == 0) {
      for (int i = 0; i < kASlotsT; ++i) {
        ptx::mbar_init(&s.full[i], kProducerWarps * 32);
        ptx::mbar_init(&s.empty[i], 1);
      }
      for (int i = 0; i < kRawStagesT; ++i) {
        ptx::mbar_init(&s.raw_full[i], 1);
        ptx::mbar_init(&s.raw_empty[i], kProducerWarps * 32);
      }
      for (int i = 0; i < kAccStages; ++i) {
        ptx::mbar_init(&s.acc_full[i], 1);
        ptx::mbar_init(&s.acc_empty[i], 4 * 32);
      }
      ptx::fence_mbar_init();
      s.zero = 0;
    }
    __syncwarp();
    ptx::tmem_alloc(&s.tmem_base, kTmemColsA);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = s.tmem_base;

  if (warp == kLoadWarpT) {
    if (ptx::elect_one()) {
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const uint32_t st = it % kRawStagesT;
        const uint32_t phase = (it / kRawStagesT) & 1u;
        ptx::mbar_wait(&s.raw_empty[st], phase ^ 1u);
        const int64_t row0 = tile * kTileM;
        const int64_t rows = (args.n - row0) < kTileM ? (args.n - row0) : kTileM;
        const uint32_t bytes = static_cast<uint32_t>(rows) * args.bytes_per_desc;
        ptx::mbar_arrive_expect_tx(&s.raw_full[st], bytes);
        ptx::bulk_load(s.raw[st], args.bits + row0 * args.bytes_per_desc, bytes, &s.raw_full[st]);
      }
    }
    __syncwarp();
  } else if (warp >= kFirstProducerWarpT) {
    // ================= producers: one TMEM lane (descriptor row) per thread =================
    // The two warps of a lane quadrant split the row: 32 bytes (two 16-byte chunks) x 8 planes each. Rows are
    // 64 bytes apart, so eight lanes reading the same chunk of their rows hit two bank groups (4-way conflict,
    // and the shared-memory pipe was the busiest unit of the kernel: 77 %); here odd row PAIRS read their two
    // chunks in the other order (2-way) and swap them back in registers.
    const int quadrant = warp & 3;                               // the TMEM lanes this warp may access
    const int chunk0 = ((warp - kFirstProducerWarpT) >> 2) * 2;  // first four warps: chunks 0-1, the others 2-3
    const int row = quadrant * 32 + lane;
    const bool swapped = (row >> 1) & 1;
    const int ca = chunk0 + (swapped ? 1 : 0), cb = chunk0 + (swapped ? 0 : 1);
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const uint32_t slot = it % kASlotsT;
      const uint32_t phase = (it / kASlotsT) & 1u;
      const uint32_t st = it % kRawStagesT;
      const uint32_t raw_phase = (it / kRawStagesT) & 1u;
      const int64_t rows_left = args.n - tile * kTileM;
      ptx::mbar_wait(&s.raw_full[st], raw_phase);
      uint4 va = make_uint4(0, 0, 0, 0), vb = make_uint4(0, 0, 0, 0);
      if (row < rows_left) {
        const uint8_t* src = s.raw[st] + row * args.bytes_per_desc;
        if (ca < nkq) va = *reinterpret_cast<const uint4*>(src + ca * 16);
        if (cb < nkq) vb = *reinterpret_cast<const uint4*>(src + cb * 16);
      }
      const uint4 lo = swapped ? vb : va, hi = swapped ? va : vb;
      const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
      ptx::mbar_wait(&s.empty[slot], phase ^ 1u);
      ptx::tc_fence_after();
      const uint32_t a_tmem = tmem_base + (static_cast<uint32_t>(quadrant * 32) << 16) + kATmemBase +
                              slot * kAColsPerTile + chunk0 * 4;
#pragma unroll
      for (int pl = 0; pl < 8; ++pl) {
        const uint32_t m = 0x01010101u << pl;
        uint32_t v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = w[j] & m;
        ptx::tmem_st_32x32b_x8(a_tmem + pl * 16, v);
      }
      // Release the raw stage only now: the stores above consumed w[], so the ld.shared have RETURNED. An arrive
      // right behind the loads issues while they are in flight and the next bulk copy overwrites the stage under
      // them (measured: 0.07 % wrong rows once the source tiles come from L2).
      ptx::mbar_arrive(&s.raw_empty[st]);
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&s.full[slot]);
    }
  } else if (warp == kMmaWarpT) {
    if (ptx::elect_one()) {   // not `lane == 0`: see ptx::elect_one
      const uint32_t b_addr = ptx::smem_u32(s.b);
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const uint32_t slot = it % kASlotsT;
        const uint32_t phase = (it / kASlotsT) & 1u;
        const uint32_t acc = it % kAccStages;
        const uint32_t acc_phase = (it / kAccStages) & 1u;
        ptx::mbar_wait(&s.acc_empty[acc], acc_phase ^ 1u);
        ptx::mbar_wait(&s.full[slot], phase);
        ptx::tc_fence_after();
        const uint32_t a_tmem = tmem_base + kATmemBase + slot * kAColsPerTile;
        const uint32_t d_tmem = tmem_base + acc * kAccCols;
#pragma unroll
        for (int ks = 0; ks < 16; ++ks) {   // K = 8 planes x 64 bytes, 32 per MMA = 8 columns of A
          const uint64_t b_desc = ptx::make_smem_desc(b_addr + ks * 2 * kBLbo, kBLbo, kBSbo);
          ptx::mma_i8_ts(d_tmem, a_tmem + ks * 8, b_desc, kIdesc, ks > 0 ? 1u : 0u);
        }
        ptx::tc_commit(&s.empty[slot]);
        ptx::tc_commit(&s.acc_full[acc]);
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue: group g = warp / 4 drains accumulator g =================
    const int group = warp >> 2;
    const int quadrant = warp & 3;
    const bool issuer = quadrant == 0 && lane == 0;   // bulk groups are per-thread state: always the same thread
    uint32_t round = 0;
    for (int64_t tile = blockIdx.x + static_cast<int64_t>(group) * gridDim.x; tile < num_tiles;
         tile += static_cast<int64_t>(kEpiGroupsT) * gridDim.x, ++round) {
      ptx::mbar_wait(&s.acc_full[group], round & 1u);
      ptx::tc_fence_after();
      uint32_t v[kN];
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quadrant * 32) << 16) + group * kAccCols;
      static_assert(kN == 64, "the accumulator row is 64 columns");
      ptx::tmem_ld_32x32b_64cols<4>(taddr, v);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&s.acc_empty[group]);
      {
        // s.zero is 0. The volatile load sits behind the arrive, so no arithmetic on v[] can be scheduled between
        // the four tcgen05.ld (ptx::tmem_ld_32x32b_64cols).
        const uint32_t z = *reinterpret_cast<volatile uint32_t*>(&s.zero);
#pragma unroll
        for (int c = 0; c < kN; ++c) v[c] ^= z;
      }
      float y[kMaxDim];
#pragma unroll
      for (int d = 0; d < kMaxDim; ++d) {
        // sum_j digit_j * 256^j, |.| < 2^43: 32 x 32 -> 64-bit multiply-adds, digit 4 lands in the high word
        long long accv = static_cast<long long>(static_cast<int>(v[d * kDigits + 0]));
#pragma unroll
        for (int j = 1; j < kDigits; ++j)
          accv += static_cast<long long>(static_cast<int>(v[d * kDigits + j])) * (1ll << (8 * j));
        y[d] = __ll2float_rn(accv) * args.scale[d];
      }
      const int64_t row0 = tile * kTileM;
      const int64_t rows_left = args.n - row0;
      const int row_in_tile = quadrant * 32 + lane;
      if (rows_left >= kTileM) {
        if (issuer) ptx::bulk_wait_read<0>();      // the group's previous store has read s.out[group]
        ptx::named_barrier_sync(1 + group, 4 * 32);
        float* o = s.out[group] + row_in_tile * args.dim;
#pragma unroll
        for (int d = 0; d < kMaxDim; ++d)
          if (d < args.dim) o[d] = y[d];
        ptx::fence_proxy_async_smem();
        ptx::named_barrier_sync(1 + group, 4 * 32);
        if (issuer) {
          ptx::bulk_store(args.out + row0 * args.dim, s.out[group], static_cast<uint32_t>(kTileM) * args.dim * 4u);
          ptx::bulk_commit();
        }
      } else if (row_in_tile < rows_left) {
        float* o = args.out + (row0 + row_in_tile) * args.dim;
#pragma unroll
        for (int d = 0; d < kMaxDim; ++d)
          if (d < args.dim) o[d] = y[d];
      }
    }
    if (issuer) ptx::bulk_wait_all();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarpT) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemColsA);
  }
}

// v -> kDigits balanced base-256 digits (|digit| <= 128 except the last, which takes the rest).
void SplitDigits(int64_t v, int8_t d[kDigits]) {
  int64_t rest = v;
  for (int j = 0; j < kDigits - 1; ++j) {
    int64_t low = ((rest % 256) + 256) % 256;
    if (low >= 128) low -= 256;
    d[j] = static_cast<int8_t>(low);
    rest = (rest - low) / 256;
  }
  d[kDigits - 1] = static_cast<int8_t>(rest);  // |v| <= 2^33 => |rest| <= 3
}

}  // namespace

// B operand image for descriptors of `nkq` 16-byte chunks: K block (plane i, chunk kq) = i * nkq + kq.
cudaError_t BuildProjectionImage(const FixedProjection& fp, int nkq, int8_t** d_image) {
  if (fp.dim > kMaxDim || fp.kp > 8 * 16 * nkq || fp.kp <= 0 || nkq < 1 || nkq > kMaxDescBytes / 16)
    return cudaErrorInvalidValue;
  std::vector<int8_t> img(kBBytes, 0);
  for (int d = 0; d < fp.dim; ++d) {
    for (int bit = 0; bit < fp.kp; ++bit) {
      const int byte = bit >> 3, plane = bit & 7;
      const int kb = plane * nkq + (byte >> 4);   // K block of 16 elements
      const int k = kb * 16 + (byte & 15);
      int8_t dig[kDigits];
      SplitDigits(static_cast<int64_t>(fp.p_int[static_cast<size_t>(d) * fp.kp + bit]) << (7 - plane), dig);
      for (int j = 0; j < kDigits; ++j) {
        const int n = d * kDigits + j;  // B row
        const size_t at = static_cast<size_t>(k / 16) * kBLbo + (n / 8) * kBSbo + (n % 8) * 16 + (k % 16);
        img[at] = dig[j];
      }
    }
  }
  cudaError_t e = cudaMalloc(d_image, kBBytes);
  if (e != cudaSuccess) return e;
  return cudaMemcpy(*d_image, img.data(), kBBytes, cudaMemcpyHostToDevice);
}

cudaError_t BuildProjectionDevice(const FixedProjection& fp, ProjectionDevice* out) {
  if (fp.dim > kMaxDim || fp.dim > 16 || fp.kp > kMaxK || fp.kp <= 0) return cudaErrorInvalidValue;
  for (int8_t*& p : out->b_image) {
    if (p) cudaFree(p);
    p = nullptr;
  }
  out->fp = &fp;
  out->dim = fp.dim;
  out->kp = fp.kp;
  for (int d = 0; d < fp.dim; ++d) out->shift[d] = fp.shift[d];
  // the image of the natural descriptor size now, other sizes on first use
  const int nkq = (fp.kp + 127) / 128;
  return BuildProjectionImage(fp, nkq, &out->b_image[nkq]);
}

cudaError_t LaunchProjection(ProjectionDevice& pd, const uint8_t* d_bits, int bytes_per_desc,
                             int64_t n, float* d_out, int sm_count, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  if (bytes_per_desc % 16 != 0 || bytes_per_desc <= 0 || bytes_per_desc > kMaxDescBytes)
    return cudaErrorInvalidValue;
  if (pd.kp > bytes_per_desc * 8) return cudaErrorInvalidValue;
  // A operand through tensor memory when the tiles can move with TMA (MLC_PROJ_TMEM=0 selects the smem-A kernel);
  // it always uses the B image of 64-byte descriptors (K padded to 64 bytes per plane)
  static const bool tmem_allowed = [] {
    const char* env = getenv("MLC_PROJ_TMEM");
    return !(env && atoi(env) == 0);
  }();
  const bool aligned = reinterpret_cast<uintptr_t>(d_bits) % 16 == 0 && reinterpret_cast<uintptr_t>(d_out) % 16 == 0 &&
                       (static_cast<size_t>(kTileM) * pd.dim * 4) % 16 == 0;
  const bool use_tmem = tmem_allowed && aligned;
  const int nkq = use_tmem ? kMaxDescBytes / 16 : bytes_per_desc / 16;
  if (!pd.b_image[nkq]) {
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (!pd.b_image[nkq]) {
      cudaError_t e = BuildProjectionImage(*pd.fp, nkq, &pd.b_image[nkq]);
      if (e != cudaSuccess) return e;
    }
  }
  ProjArgs a;
  a.bits = d_bits;
  a.b_image = pd.b_image[nkq];
  a.out = d_out;
  a.n = n;
  a.bytes_per_desc = bytes_per_desc;
  a.k_steps = 8 * bytes_per_desc / 32;
  a.dim = pd.dim;
  a.bulk_io = aligned ? 1 : 0;
  for (int d = 0; d < kMaxDim; ++d)
    a.scale[d] = d < pd.dim ? ldexpf(1.0f, -(pd.shift[d] + 7)) : 0.f;
  const int64_t tiles = (n + kTileM - 1) / kTileM;
  const unsigned grid = static_cast<unsigned>(tiles < sm_count ? tiles : sm_count);
  cudaError_t e;
  if (use_tmem) {
    e = cudaFuncSetAttribute(projection_tmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(sizeof(SmemT) + 128));
    if (e != cudaSuccess) return e;
    projection_tmem_kernel<<<grid, kThreadsT, sizeof(SmemT) + 128, stream>>>(a);
  } else {
    e = cudaFuncSetAttribute(projection_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(sizeof(Smem) + 128));
    if (e != cudaSuccess) return e;
    projection_kernel<<<grid, kThreads, sizeof(Smem) + 128, stream>>>(a);
  }
  CountLaunch();
  return cudaGetLastError();
}

}  // namespace mlc
