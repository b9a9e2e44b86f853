// Microbenchmark (evidence for DESIGN.md §4 kernel 2b): what HBM bandwidth does B200 deliver for the
// access pattern of an inverted-list scan — many short, randomly placed chunks — as a function of
// chunk length, per-lane vs coalesced addressing, and loads in flight per warp?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o chunk_read chunk_read.cu && ./chunk_read
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t Hash(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// MODE 0: lane l reads the 48-byte entry (chunk[l / LEN] + l % LEN) with 3 x LDG.128 (AoS-48)
// MODE 1: the same bytes, coalesced: a chunk of LEN*48 bytes is read by LEN*3 consecutive lanes
// U = independent trips (32 entries each) issued before any is consumed.
template <int MODE, int LEN, int U>
__global__ void __launch_bounds__(256, 4) read_kernel(const uint4* __restrict__ buf, uint32_t n_entries,
                                                      int trips, uint32_t* __restrict__ sink) {
  const int lane = threadIdx.x & 31;
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  uint32_t acc = 0;
  for (int t = 0; t < trips; t += U) {
    uint4 v[U][3];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t trip = gw * 7919u + static_cast<uint32_t>(t + u) * 104729u;
      if (MODE == 0) {
        const uint32_t start = Hash(trip * 8u + lane / LEN) % (n_entries - 64);
        const uint4* p = buf + static_cast<size_t>(start + lane % LEN) * 3;
#pragma unroll
        for (int j = 0; j < 3; ++j) v[u][j] = __ldg(p + j);
      } else {
        // 96 16-byte pieces per trip, piece i belongs to chunk i / (3 LEN)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int piece = j * 32 + lane;
          const uint32_t start = Hash(trip * 8u + piece / (3 * LEN)) % (n_entries - 64);
          v[u][j] = __ldg(buf + static_cast<size_t>(start) * 3 + piece % (3 * LEN));
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int j = 0; j < 3; ++j) acc += v[u][j].x ^ v[u][j].y ^ v[u][j].z ^ v[u][j].w;
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

template <int MODE, int LEN, int U>
void Run(const uint4* buf, uint32_t n_entries, uint32_t* sink, void* flush, size_t flush_bytes) {
  const int trips = 64, blocks = 148 * 4 * 4;
  float best = 1e9f;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  for (int rep = 0; rep < 4; ++rep) {
    cudaMemset(flush, rep, flush_bytes);
    cudaEventRecord(a);
    read_kernel<MODE, LEN, U><<<blocks, 256>>>(buf, n_entries, trips, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (rep > 0 && ms < best) best = ms;
  }
  const double bytes = static_cast<double>(blocks) * 8 * trips * 32 * 48;
  printf("mode %d (%s) chunk %4d B, %d trip(s) in flight/warp: %.3f ms, %.0f GB/s\n", MODE,
         MODE == 0 ? "per-lane 3xLDG.128" : "coalesced", LEN * 48, U, best, bytes / best * 1e-6);
}

int main() {
  const uint32_t n_entries = 80u << 20;  // 3.75 GiB of 48-byte entries: each chunk is touched ~0.5 times (no L2 reuse)
  uint4* buf;
  uint32_t* sink;
  void* flush;
  const size_t flush_bytes = 256u << 20;
  cudaMalloc(&buf, static_cast<size_t>(n_entries) * 48);
  cudaMemset(buf, 1, static_cast<size_t>(n_entries) * 48);
  cudaMalloc(&sink, 4);
  cudaMalloc(&flush, flush_bytes);
  Run<0, 1, 1>(buf, n_entries, sink, flush, flush_bytes);
  Run<0, 5, 1>(buf, n_entries, sink, flush, flush_bytes);
  Run<0, 5, 2>(buf, n_entries, sink, flush, flush_bytes);
  Run<0, 5, 4>(buf, n_entries, sink, flush, flush_bytes);
  Run<0, 32, 1>(buf, n_entries, sink, flush, flush_bytes);
  Run<0, 32, 2>(buf, n_entries, sink, flush, flush_bytes);
  Run<0, 32, 4>(buf, n_entries, sink, flush, flush_bytes);
  Run<1, 5, 1>(buf, n_entries, sink, flush, flush_bytes);
  Run<1, 5, 2>(buf, n_entries, sink, flush, flush_bytes);
  Run<1, 5, 4>(buf, n_entries, sink, flush, flush_bytes);
  Run<1, 32, 1>(buf, n_entries, sink, flush, flush_bytes);
  Run<1, 32, 2>(buf, n_entries, sink, flush, flush_bytes);
  Run<1, 32, 4>(buf, n_entries, sink, flush, flush_bytes);
  return 0;
}
