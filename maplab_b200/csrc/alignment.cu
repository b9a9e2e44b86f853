// Mission-level alignment that follows the loop-closure queries:
// common::transformationRansac (common/maplab-common/include/maplab-common/geometry-inl.h:113-182)
// over the per-vertex T_G_M samples of LoopDetectorNode::detectLoopClosuresMissionToDatabase
// (loop-closure-handler/src/loop-detector-node.cc:875-959), least-squares quaternion average
// (common/maplab-common/src/geometry.cc:9-31) and the yaw-only projection (:944-955).
//
// The RANSAC hypotheses are the samples themselves and the random draws do not depend on the data,
// so the device counts the inliers of EVERY sample once (n x n pose comparisons, one CTA per
// hypothesis, block reduction — no atomics) and the host replays the reference's draw sequence
// (mt19937 + libstdc++ uniform_int_distribution<int>(0, n-1), both mappings) over those counts:
// the first drawn sample whose count beats the best so far (initially {0}) wins. fp64, compiled
// with -fmad=false like the geometric verification.
#include <cub/cub.cuh>

#include <cmath>
#include <cstring>
#include <vector>

#include "detector.h"

namespace mlc {
namespace {

__device__ __forceinline__ bool PoseIsInlier(const double* qa, const double* pa, const double* qb,
                                             const double* pb, double thr_rad, double thr_m) {
  const double dx = pa[0] - pb[0], dy = pa[1] - pb[1], dz = pa[2] - pb[2];
  const double pn = sqrt(dx * dx + dy * dy + dz * dz);
  // Eigen::Quaterniond::angularDistance: d = a * conj(b), 2 * atan2(|d.vec|, |d.w|)
  const double bx = -qb[0], by = -qb[1], bz = -qb[2], bw = qb[3];
  const double w = qa[3] * bw - qa[0] * bx - qa[1] * by - qa[2] * bz;
  const double x = qa[3] * bx + qa[0] * bw + qa[1] * bz - qa[2] * by;
  const double y = qa[3] * by + qa[1] * bw + qa[2] * bx - qa[0] * bz;
  const double z = qa[3] * bz + qa[2] * bw + qa[0] * by - qa[1] * bx;
  const double ang = 2.0 * atan2(sqrt(x * x + y * y + z * z), fabs(w));
  return pn < thr_m && ang < thr_rad;
}

// counts[s] = number of samples within the thresholds of sample s (itself included)
__global__ void __launch_bounds__(256)
pose_inlier_count_kernel(const double* __restrict__ quats, const double* __restrict__ pos, int n,
                         double thr_rad, double thr_m, int* __restrict__ counts) {
  typedef cub::BlockReduce<int, 256> Reduce;
  __shared__ typename Reduce::TempStorage tmp;
  for (int s = blockIdx.x; s < n; s += gridDim.x) {
    int local = 0;
    for (int j = threadIdx.x; j < n; j += blockDim.x)
      local += PoseIsInlier(quats + 4 * s, pos + 3 * s, quats + 4 * j, pos + 3 * j, thr_rad, thr_m) ? 1 : 0;
    const int total = Reduce(tmp).Sum(local);
    if (threadIdx.x == 0) counts[s] = total;
    __syncthreads();
  }
}

// flags[j] = sample j is an inlier of the winning sample
__global__ void pose_inlier_flags_kernel(const double* __restrict__ quats, const double* __restrict__ pos,
                                         int n, int winner, double thr_rad, double thr_m,
                                         unsigned char* __restrict__ flags) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n)
    flags[j] = PoseIsInlier(quats + 4 * winner, pos + 3 * winner, quats + 4 * j, pos + 3 * j, thr_rad, thr_m) ? 1 : 0;
}

// std::mt19937
class Mt19937 {
 public:
  explicit Mt19937(uint32_t seed) : idx_(624) {
    mt_[0] = seed;
    for (int i = 1; i < 624; ++i) mt_[i] = 1812433253u * (mt_[i - 1] ^ (mt_[i - 1] >> 30)) + i;
  }
  uint32_t Next() {
    if (idx_ >= 624) {
      for (int i = 0; i < 624; ++i) {
        const uint32_t y = (mt_[i] & 0x80000000u) | (mt_[(i + 1) % 624] & 0x7fffffffu);
        mt_[i] = mt_[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      idx_ = 0;
    }
    uint32_t y = mt_[idx_++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }

 private:
  uint32_t mt_[624];
  int idx_;
};

// libstdc++ uniform_int_distribution<int>(0, n - 1)(mt19937): GCC >= 11 multiply-shift with
// rejection (mapping 1), GCC <= 10 scaling with rejection (mapping 0).
int UniformIndex(Mt19937* g, uint32_t n, int mapping) {
  if (mapping == 1) {
    uint64_t product = static_cast<uint64_t>(g->Next()) * n;
    uint32_t low = static_cast<uint32_t>(product);
    if (low < n) {
      const uint32_t threshold = static_cast<uint32_t>(-n) % n;
      while (low < threshold) {
        product = static_cast<uint64_t>(g->Next()) * n;
        low = static_cast<uint32_t>(product);
      }
    }
    return static_cast<int>(product >> 32);
  }
  const uint32_t scaling = 0xFFFFFFFFu / n;
  const uint32_t past = n * scaling;
  uint32_t ret;
  do {
    ret = g->Next();
  } while (ret >= past);
  return static_cast<int>(ret / scaling);
}

// Smallest-eigenvalue eigenvector of a symmetric 4x4 matrix (cyclic Jacobi) == the last right
// singular vector Eigen::JacobiSVD returns for A with S = A^T A (up to sign).
void SmallestEigenvector4(const double S_in[16], double v_out[4]) {
  double S[16], V[16];
  std::memcpy(S, S_in, sizeof(S));
  for (int i = 0; i < 16; ++i) V[i] = (i % 5 == 0) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < 4; ++p)
      for (int q = p + 1; q < 4; ++q) off += S[p * 4 + q] * S[p * 4 + q];
    if (off < 1e-300) break;
    for (int p = 0; p < 4; ++p)
      for (int q = p + 1; q < 4; ++q) {
        const double apq = S[p * 4 + q];
        if (apq == 0.0) continue;
        const double theta = (S[q * 4 + q] - S[p * 4 + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 4; ++k) {
          const double skp = S[k * 4 + p], skq = S[k * 4 + q];
          S[k * 4 + p] = c * skp - s * skq;
          S[k * 4 + q] = s * skp + c * skq;
        }
        for (int k = 0; k < 4; ++k) {
          const double spk = S[p * 4 + k], sqk = S[q * 4 + k];
          S[p * 4 + k] = c * spk - s * sqk;
          S[q * 4 + k] = s * spk + c * sqk;
        }
        for (int k = 0; k < 4; ++k) {
          const double vkp = V[k * 4 + p], vkq = V[k * 4 + q];
          V[k * 4 + p] = c * vkp - s * vkq;
          V[k * 4 + q] = s * vkp + c * vkq;
        }
      }
  }
  int best = 0;
  for (int i = 1; i < 4; ++i)
    if (S[i * 4 + i] < S[best * 4 + best]) best = i;
  for (int k = 0; k < 4; ++k) v_out[k] = V[k * 4 + best];
}

}  // namespace

bool Detector::TransformationRansac(const double* quats, const double* positions, int64_t n,
                                    const mlc_alignment_settings& as, double* out_quat, double* out_pos,
                                    int32_t* inlier_indices, int32_t* num_inliers, std::string* err) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  if (n <= 0 || n > (1 << 24)) {
    *err = "transformationRansac needs 1 .. 2^24 samples";  // CHECK(!T_A_B_samples.empty())
    return false;
  }
  if (as.num_iterations < 0 || as.max_orientation_error_rad < 0.0 || as.max_position_error_m < 0.0) {
    *err = "negative alignment RANSAC setting";  // CHECK_GE at loop-detector-node.cc:917-921
    return false;
  }
  if (n == 1) {  // geometry-inl.h:128-132
    std::memcpy(out_quat, quats, 4 * sizeof(double));
    std::memcpy(out_pos, positions, 3 * sizeof(double));
    *num_inliers = 1;
    if (inlier_indices) inlier_indices[0] = 0;
    return true;
  }
  const int ni = static_cast<int>(n);
  DevBuf& buf = d_ransac_[0];
  const size_t o_q = 0, o_p = sizeof(double) * 4 * n, o_c = o_p + sizeof(double) * 3 * n + 64,
               o_f = o_c + sizeof(int) * n + 64;
  if (!Cuda(buf.Reserve(o_f + n + 64), "alloc", err)) return false;
  unsigned char* base = buf.as<unsigned char>();
  double* d_q = reinterpret_cast<double*>(base + o_q);
  double* d_p = reinterpret_cast<double*>(base + o_p);
  int* d_counts = reinterpret_cast<int*>(base + ((o_c + 15) & ~size_t{15}));
  unsigned char* d_flags = base + o_f;
  if (!Cuda(cudaMemcpyAsync(d_q, quats, sizeof(double) * 4 * n, cudaMemcpyHostToDevice, stream_), "H2D", err) ||
      !Cuda(cudaMemcpyAsync(d_p, positions, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, stream_), "H2D", err))
    return false;
  const int grid = ni < sm_count_ * 8 ? ni : sm_count_ * 8;
  pose_inlier_count_kernel<<<grid, 256, 0, stream_>>>(d_q, d_p, ni, as.max_orientation_error_rad,
                                                     as.max_position_error_m, d_counts);
  CountLaunch();
  std::vector<int> counts(n);
  if (!Cuda(cudaMemcpyAsync(counts.data(), d_counts, sizeof(int) * n, cudaMemcpyDeviceToHost, stream_), "D2H", err) ||
      !Cuda(cudaStreamSynchronize(stream_), "pose inlier counts", err))
    return false;
  // replay of the reference's loop (geometry-inl.h:139-161): best starts as {0}
  Mt19937 gen(as.seed);
  int best_count = 1, winner = -1;
  for (int it = 0; it < as.num_iterations; ++it) {
    const int s = UniformIndex(&gen, static_cast<uint32_t>(n), as.rng_mapping);
    if (counts[s] > best_count) {
      best_count = counts[s];
      winner = s;
    }
  }
  std::vector<int> members;
  if (winner < 0) {
    members.push_back(0);
  } else {
    pose_inlier_flags_kernel<<<(ni + 255) / 256, 256, 0, stream_>>>(
        d_q, d_p, ni, winner, as.max_orientation_error_rad, as.max_position_error_m, d_flags);
    CountLaunch();
    std::vector<unsigned char> flags(n);
    if (!Cuda(cudaMemcpyAsync(flags.data(), d_flags, n, cudaMemcpyDeviceToHost, stream_), "D2H", err) ||
        !Cuda(cudaStreamSynchronize(stream_), "pose inlier flags", err))
      return false;
    for (int j = 0; j < ni; ++j)
      if (flags[j]) members.push_back(j);
  }
  // least-squares refinement on the inliers (geometry-inl.h:165-175, geometry.cc:9-31)
  if (members.size() == 1) {
    std::memcpy(out_quat, quats + 4 * members[0], 4 * sizeof(double));
  } else {
    double S[16] = {0};
    for (int m : members) {
      const double* q = quats + 4 * m;
      const double L[3][4] = {{q[3], q[2], -q[1], q[0]}, {-q[2], q[3], q[0], q[1]}, {q[1], -q[0], q[3], q[2]}};
      for (int r = 0; r < 3; ++r)
        for (int i = 0; i < 4; ++i)
          for (int j = 0; j < 4; ++j) S[i * 4 + j] += L[r][i] * L[r][j];
    }
    double v[4];
    SmallestEigenvector4(S, v);
    out_quat[0] = -v[0];  // quaternionInverseJPL
    out_quat[1] = -v[1];
    out_quat[2] = -v[2];
    out_quat[3] = v[3];
  }
  double p[3] = {0, 0, 0};
  for (int m : members)
    for (int k = 0; k < 3; ++k) p[k] += positions[3 * m + k];
  for (int k = 0; k < 3; ++k) out_pos[k] = p[k] / static_cast<double>(members.size());
  *num_inliers = static_cast<int32_t>(members.size());
  if (inlier_indices)
    for (size_t i = 0; i < members.size(); ++i) inlier_indices[i] = members[i];
  return true;
}

}  // namespace mlc
