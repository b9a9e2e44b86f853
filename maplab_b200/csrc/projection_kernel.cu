// Kernel 1: binary descriptor projection (descriptor_projection::ProjectDescriptorBlock,
// descriptor-projection/src/descriptor-projection.cc:15-50; bit order of DescriptorToEigenMatrix,
// descriptor-projection/include/descriptor-projection/descriptor-projection.h:92-115).
//
// Y[n][d] = sum_k P[d][k] * bit_k(desc n) as an exact integer GEMM on the 5th-gen tensor cores:
//   A (128 x K)  = descriptor bits expanded to u8 {0,1} by the producer warps, written straight
//                  into shared memory in the UMMA K-major core-matrix layout;
//   B (48 x K)   = the fixed-point projection matrix split into 4 balanced base-256 s8 digits
//                  (row 4*d + j = digit j of output dim d; rows >= 4*dim are zero);
//   D (128 x 48) = s32 accumulators in TMEM (tcgen05.mma kind::i8), double buffered;
//   epilogue     = tcgen05.ld, recombine the digits in int64, ONE rounding to fp32, store.
// Persistent CTAs (one per SM), warp-specialised: 4 epilogue warps, 1 MMA/TMEM warp,
// 8 bit-expansion warps; mbarrier pipelines between the roles.
#include "device_index.h"
#include "ptx.cuh"

namespace mlc {
namespace {

constexpr int kTileM = 128;
constexpr int kMaxKBytes = 512;                 // descriptor bits (<= 512)
constexpr int kASlotBytes = kTileM * kMaxKBytes;  // 64 KB: u8 per bit
constexpr int kASlots = 3;
constexpr int kALbo = 16 * 128;                 // K-adjacent core matrices: 16 row groups apart
constexpr int kASbo = 128;                      // M-adjacent 8-row groups
constexpr int kBRowGroups = kProjNPad / 8;      // 6
constexpr int kBLbo = kBRowGroups * 128;        // 768
constexpr int kBSbo = 128;
constexpr int kBBytes = (kMaxKBytes / 16) * kBLbo;  // 24576
constexpr int kEpilogueWarps = 4;
constexpr int kMmaWarp = 4;
constexpr int kProducerWarps = 8;
constexpr int kRowGroupsPerWarp = (kTileM / 8) / kProducerWarps;  // 8-row core-matrix groups per producer warp
constexpr int kFirstProducerWarp = 5;
constexpr int kThreads = (kFirstProducerWarp + kProducerWarps) * 32;  // 416
constexpr int kAccStages = 2;
constexpr int kAccCols = 64;                    // TMEM columns per accumulator stage
constexpr int kTmemCols = 128;
constexpr int kMaxDim = kProjNPad / kProjDigits;  // 12

struct Smem {
  alignas(128) uint8_t a[kASlots][kASlotBytes];
  alignas(128) int8_t b[kBBytes];
  alignas(8) uint64_t full[kASlots];
  uint64_t empty[kASlots];
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
  int k_steps;         // MMAs per tile = 8 * bytes_per_desc / 32
  int dim;
  float scale[kMaxDim];  // 2^-shift[d]
};

// u8 instruction descriptor: D = s32, A = u8, B = s8, both K-major, N = 48, M = 128.
constexpr uint32_t kIdesc = (2u << 4) | (0u << 7) | (1u << 10) |
                            (static_cast<uint32_t>(kProjNPad >> 3) << 17) |
                            (static_cast<uint32_t>(kTileM >> 4) << 24);

// 4 descriptor bits -> 4 bytes {0,1}, LSB first: bit i lands at bit 8*i (no carries since the
// partial products i + 7*j are distinct for i, j in 0..3).
__device__ __forceinline__ uint32_t Expand4(uint32_t nibble) {
  return (nibble * 0x00204081u) & 0x01010101u;
}

__global__ void __launch_bounds__(kThreads, 1) projection_kernel(ProjArgs args) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem& s = *reinterpret_cast<Smem*>(
      smem_raw + ((128u - (ptx::smem_u32(smem_raw) & 127u)) & 127u));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t num_tiles = (args.n + kTileM - 1) / kTileM;

  // ---- one-time setup ----
  for (int i = threadIdx.x * 16; i < kBBytes; i += kThreads * 16)
    *reinterpret_cast<uint4*>(s.b + i) = *reinterpret_cast<const uint4*>(args.b_image + i);
  if (warp == kMmaWarp) {
    if (lane == 0) {
      for (int i = 0; i < kASlots; ++i) {
        ptx::mbar_init(&s.full[i], kProducerWarps * 32);
        ptx::mbar_init(&s.empty[i], 1);
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

  if (warp >= kFirstProducerWarp) {
    // ================= bit-expansion producers =================
    const int pw = warp - kFirstProducerWarp;
    const int r = lane & 7;    // row inside the 8-row core matrix
    const int kq = lane >> 3;  // 128-bit quarter of the descriptor
    const int nkq = args.bytes_per_desc >> 4;
    // The descriptor words of kPrefetch tiles are in flight per warp (registers): with one tile
    // the producers were bound by the latency of their own loads (ncu r1r: long-scoreboard 4.7,
    // DRAM 12 %).
    constexpr int kPrefetch = 4;
    uint4 w[kPrefetch][kRowGroupsPerWarp];
    auto load_tile = [&](int64_t tile, uint4 (&ww)[kRowGroupsPerWarp]) {
#pragma unroll
      for (int h = 0; h < kRowGroupsPerWarp; ++h) {
        const int rg = pw + h * kProducerWarps;
        const int64_t row = tile * kTileM + rg * 8 + r;
        ww[h] = make_uint4(0, 0, 0, 0);
        if (row < args.n && kq < nkq)
          ww[h] = ptx::ldg_nc_v4(args.bits + row * args.bytes_per_desc + kq * 16);
      }
    };
#pragma unroll
    for (int p = 0; p < kPrefetch; ++p) load_tile(blockIdx.x + static_cast<int64_t>(p) * gridDim.x, w[p]);
    uint32_t it = 0;
    for (int64_t tile0 = blockIdx.x; tile0 < num_tiles; tile0 += static_cast<int64_t>(kPrefetch) * gridDim.x) {
#pragma unroll
      for (int p = 0; p < kPrefetch; ++p) {
        const int64_t tile = tile0 + static_cast<int64_t>(p) * gridDim.x;
        if (tile >= num_tiles) break;
        const uint32_t slot = it % kASlots;
        const uint32_t phase = (it / kASlots) & 1u;
        ++it;
        ptx::mbar_wait(&s.empty[slot], phase ^ 1u);
        if (kq < nkq) {
#pragma unroll
          for (int h = 0; h < kRowGroupsPerWarp; ++h) {
            const int rg = pw + h * kProducerWarps;
            uint8_t* dst = s.a[slot] + (kq * 8) * kALbo + rg * kASbo + r * 16;
            const uint32_t words[4] = {w[p][h].x, w[p][h].y, w[p][h].z, w[p][h].w};
#pragma unroll
            for (int wi = 0; wi < 4; ++wi) {
              const uint32_t x = words[wi];
              const uint32_t even = x & 0x0F0F0F0Fu;         // nibbles 0,2,4,6 in bytes 0..3
              const uint32_t odd = (x >> 4) & 0x0F0F0F0Fu;   // nibbles 1,3,5,7
              // 16 bits -> 16 bytes -> one 16-byte store; two stores per 32-bit word
#pragma unroll
              for (int half = 0; half < 2; ++half) {
                uint4 o;
                o.x = Expand4(__byte_perm(even, 0, 0x4440 + (2 * half)));
                o.y = Expand4(__byte_perm(odd, 0, 0x4440 + (2 * half)));
                o.z = Expand4(__byte_perm(even, 0, 0x4441 + (2 * half)));
                o.w = Expand4(__byte_perm(odd, 0, 0x4441 + (2 * half)));
                *reinterpret_cast<uint4*>(dst + (wi * 2 + half) * kALbo) = o;
              }
            }
          }
        }
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive(&s.full[slot]);
        load_tile(tile + static_cast<int64_t>(kPrefetch) * gridDim.x, w[p]);  // refill this register slot
      }
    }
  } else if (warp == kMmaWarp) {
    // ================= MMA issuer =================
    if (lane == 0) {
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
      uint32_t v[kProjNPad];
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + acc * kAccCols;
#pragma unroll
      for (int c = 0; c < kProjNPad / 8; ++c) {
        uint32_t t8[8];
        ptx::tmem_ld_32x32b_x8(taddr + c * 8, t8);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[c * 8 + i] = t8[i];
      }
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&s.acc_empty[acc]);
      const int64_t row = tile * kTileM + warp * 32 + lane;
      if (row < args.n) {
        float y[kMaxDim];
#pragma unroll
        for (int d = 0; d < kMaxDim; ++d) {
          long long accv = 0;
#pragma unroll
          for (int j = kProjDigits - 1; j >= 0; --j)
            accv = accv * 256 + static_cast<int>(v[d * kProjDigits + j]);
          y[d] = __ll2float_rn(accv) * args.scale[d];  // exact power-of-two scaling
        }
        float* o = args.out + row * args.dim;
        if ((args.dim & 1) == 0) {
#pragma unroll
          for (int d = 0; d < kMaxDim; d += 2)
            if (d < args.dim) *reinterpret_cast<float2*>(o + d) = make_float2(y[d], y[d + 1]);
        } else {
#pragma unroll
          for (int d = 0; d < kMaxDim; ++d)
            if (d < args.dim) o[d] = y[d];
        }
      }
    }
  }

  // ---- teardown ----
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace

cudaError_t BuildProjectionDevice(const FixedProjection& fp, ProjectionDevice* out) {
  if (fp.dim > kMaxDim || fp.dim > 16 || fp.kp > kMaxKBytes || fp.kp <= 0)
    return cudaErrorInvalidValue;
  std::vector<int8_t> img(kBBytes, 0);
  for (int d = 0; d < fp.dim; ++d) {
    for (int k = 0; k < fp.kp; ++k) {
      int8_t dig[kProjDigits];
      SplitDigitsBase256(fp.p_int[static_cast<size_t>(d) * fp.kp + k], dig);
      for (int j = 0; j < kProjDigits; ++j) {
        const int n = d * kProjDigits + j;  // B row
        const size_t at = static_cast<size_t>(k / 16) * kBLbo + (n / 8) * kBSbo + (n % 8) * 16 + (k % 16);
        img[at] = dig[j];
      }
    }
  }
  if (out->b_image) cudaFree(out->b_image);
  out->b_image = nullptr;
  cudaError_t e = cudaMalloc(&out->b_image, kBBytes);
  if (e != cudaSuccess) return e;
  e = cudaMemcpy(out->b_image, img.data(), kBBytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return e;
  out->b_bytes = kBBytes;
  out->dim = fp.dim;
  out->kp = fp.kp;
  for (int d = 0; d < fp.dim; ++d) out->shift[d] = fp.shift[d];
  return cudaSuccess;
}

cudaError_t LaunchProjection(const ProjectionDevice& pd, const uint8_t* d_bits, int bytes_per_desc,
                             int64_t n, float* d_out, int sm_count, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  if (bytes_per_desc % 16 != 0 || bytes_per_desc <= 0 || bytes_per_desc > kMaxKBytes / 8)
    return cudaErrorInvalidValue;
  if (pd.kp > bytes_per_desc * 8) return cudaErrorInvalidValue;
  ProjArgs a;
  a.bits = d_bits;
  a.b_image = pd.b_image;
  a.out = d_out;
  a.n = n;
  a.bytes_per_desc = bytes_per_desc;
  a.k_steps = bytes_per_desc * 8 / 32;
  a.dim = pd.dim;
  for (int d = 0; d < kMaxDim; ++d)
    a.scale[d] = d < pd.dim ? ldexpf(1.0f, -pd.shift[d]) : 0.f;
  cudaError_t e = cudaFuncSetAttribute(projection_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(sizeof(Smem) + 128));
  if (e != cudaSuccess) return e;
  const int64_t tiles = (n + kTileM - 1) / kTileM;
  const unsigned grid = static_cast<unsigned>(tiles < sm_count ? tiles : sm_count);
  projection_kernel<<<grid, kThreads, sizeof(Smem) + 128, stream>>>(a);
  CountLaunch();
  return cudaGetLastError();
}

}  // namespace mlc
