// Issue rate of tcgen05.mma on this GPU for the shapes kernel 1 uses: one thread issues `reps` x `kchain` MMAs
// (kchain dependent accumulations into one accumulator, like one projection tile), commits, waits. Reports
// cycles per MMA for kind::i8 / kind::f8f6f4, A in shared memory / tensor memory, N = 64 / 128 / 256, and with
// one or two accumulators alternating. Build + run: see the Makefile line at the bottom.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../maplab_b200/csrc/ptx.cuh"

using namespace mlc;

__device__ __forceinline__ void mma_f8_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_f8_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

struct Smem {
  alignas(128) uint8_t a[128 * 32 * 2];
  alignas(128) uint8_t b[256 * 32 * 2];
  uint64_t bar;
  uint32_t tmem;
};

// mode bit 0: A in TMEM; bit 1: f8 instead of i8
__global__ void __launch_bounds__(128, 1) rate_kernel(int mode, int n, int kchain, int reps, int naccs, long long* out) {
  __shared__ Smem s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (int)sizeof(s.a); i += 128) s.a[i] = 1;
  for (int i = threadIdx.x; i < (int)sizeof(s.b); i += 128) s.b[i] = 1;
  if (warp == 0) {
    if (lane == 0) { ptx::mbar_init(&s.bar, 1); ptx::fence_mbar_init(); }
    __syncwarp();
    ptx::tmem_alloc(&s.tmem, 512);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = s.tmem;
  if (warp == 0 && ptx::elect_one()) {
    const uint32_t idesc = (mode & 2) ? ((1u << 4) | ((uint32_t)(n >> 3) << 17) | (8u << 24))
                                      : ((2u << 4) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | (8u << 24));
    const uint32_t a_addr = ptx::smem_u32(s.a), b_addr = ptx::smem_u32(s.b);
    const uint32_t lbo_a = 16 * 128, lbo_b = (n / 8) * 128;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int ks = 0; ks < 16; ++ks) {
        // naccs > 0: one accumulator per chain, chains alternate; naccs < 0: -naccs accumulators INTERLEAVED
        // inside the chain (consecutive MMAs are independent)
        const uint32_t d = tmem + (naccs > 0 ? (r % naccs) : (ks & (-naccs - 1))) * n;
        const uint64_t b_desc = ptx::make_smem_desc(b_addr + (ks & 1) * 2 * lbo_b, lbo_b, 128);
        if (mode & 1) {
          const uint32_t a_t = tmem + 256 + (ks & 15) * 8;
          if (mode & 2) mma_f8_ts(d, a_t, b_desc, idesc, ks >= abs(naccs)); else ptx::mma_i8_ts(d, a_t, b_desc, idesc, ks >= abs(naccs));
        } else {
          const uint64_t a_desc = ptx::make_smem_desc(a_addr + (ks & 1) * 2 * lbo_a, lbo_a, 128);
          if (mode & 2) mma_f8_ss(d, a_desc, b_desc, idesc, ks >= abs(naccs)); else ptx::mma_i8_ss(d, a_desc, b_desc, idesc, ks >= abs(naccs));
        }
      }
    }
    ptx::tc_commit(&s.bar);
    ptx::mbar_wait(&s.bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 8);
  const char* names[4] = {"i8 A=smem", "i8 A=tmem", "f8 A=smem", "f8 A=tmem"};
  for (int mode = 0; mode < 4; ++mode)
    for (int n : {64, 128, 256})
      for (int naccs : {1, -2, -4}) {
        if (abs(naccs) * n > 256) continue;
        const int kchain = 16, reps = 64;
        for (int grid : {148}) {
          rate_kernel<<<grid, 128>>>(mode, n, kchain, reps, naccs, d_out);
          rate_kernel<<<grid, 128>>>(mode, n, kchain, reps, naccs, d_out);
          cudaError_t e = cudaDeviceSynchronize();
          long long c = 0;
          cudaMemcpy(&c, d_out, 8, cudaMemcpyDeviceToHost);
          printf("%-10s N=%3d accs=%d grid=%3d: %7.1f cycles/MMA (M=128, K=32 B)  %s\n", names[mode], n, naccs, grid,
                 double(c) / (kchain * reps), e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
      }
  return 0;
}
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/microbench/mma_rate.bin profiles/microbench/mma_rate.cu (here), then profiles/microbench/mma_rate.bin on the GPU box
