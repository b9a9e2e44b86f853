// ORACLE (test infrastructure): CPU restatement of maplab's mission-level alignment step that
// follows the loop-closure query path:
//   common::transformationRansac        common/maplab-common/include/maplab-common/geometry-inl.h:113-182
//   common::ComputeLSAverageQuaternionJPL   common/maplab-common/src/geometry.cc:9-31
//   yaw-only projection                 loop-closure-handler/src/loop-detector-node.cc:923-959,
//                                       geometry-inl.h:44-84 (RotationMatrixToRollPitchYaw / RollPitchYawToRotationMatrix)
// Poses are (quaternion x, y, z, w — Eigen::Quaterniond::coeffs() order, Hamilton; position).
// Third-party arithmetic restated: Eigen::Quaterniond::angularDistance (Eigen 3.3:
// 2 * atan2(|d.vec|, |d.w|), d = a * conj(b)); Eigen::JacobiSVD's smallest right singular vector
// of A == eigenvector of the smallest eigenvalue of A^T A (4x4, cyclic Jacobi here; the sign of
// the vector is not pinned — +q and -q are the same rotation); libstdc++'s
// uniform_int_distribution<int>(0, n-1) over mt19937 in both mappings (GCC <= 10: scaling +
// rejection, GCC >= 11: Lemire's multiply-shift), SURVEY F11.
#include <cmath>
#include <cstring>

#include "lc_oracle.h"

namespace lc_oracle {

int UniformIndex(RansacRng* rng, uint32_t n, int mapping) {
  if (mapping == 1) {  // libstdc++ >= 11, _S_nd
    uint64_t product = static_cast<uint64_t>(rng->NextU32()) * n;
    uint32_t low = static_cast<uint32_t>(product);
    if (low < n) {
      const uint32_t threshold = static_cast<uint32_t>(-n) % n;
      while (low < threshold) {
        product = static_cast<uint64_t>(rng->NextU32()) * n;
        low = static_cast<uint32_t>(product);
      }
    }
    return static_cast<int>(product >> 32);
  }
  const uint32_t scaling = 0xFFFFFFFFu / n;  // libstdc++ <= 10: urngrange / uerange
  const uint32_t past = n * scaling;
  uint32_t ret;
  do {
    ret = rng->NextU32();
  } while (ret >= past);
  return static_cast<int>(ret / scaling);
}

double AngularDistance(const double* a, const double* b) {  // (x, y, z, w)
  // d = a * conj(b)
  const double bx = -b[0], by = -b[1], bz = -b[2], bw = b[3];
  const double w = a[3] * bw - a[0] * bx - a[1] * by - a[2] * bz;
  const double x = a[3] * bx + a[0] * bw + a[1] * bz - a[2] * by;
  const double y = a[3] * by + a[1] * bw + a[2] * bx - a[0] * bz;
  const double z = a[3] * bz + a[2] * bw + a[0] * by - a[1] * bx;
  return 2.0 * std::atan2(std::sqrt(x * x + y * y + z * z), std::fabs(w));
}

bool PoseIsInlier(const double* qa, const double* pa, const double* qb, const double* pb,
                  double thr_rad, double thr_m) {
  const double dx = pa[0] - pb[0], dy = pa[1] - pb[1], dz = pa[2] - pb[2];
  const double pn = std::sqrt(dx * dx + dy * dy + dz * dz);
  return pn < thr_m && AngularDistance(qa, qb) < thr_rad;
}

// Eigenvector of the smallest eigenvalue of the symmetric 4x4 matrix S (cyclic Jacobi).
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

// ComputeLSAverageQuaternionJPL: rows of A are the top 3 rows of L(q) (geometry-inl.h:95-110),
// L(q)[0:3, :] = [ q4 I - skew(q_v) | q_v ]; the average is quaternionInverseJPL of the smallest
// right singular vector, i.e. (-v_xyz, v_w).
void LsAverageQuaternion(const double* quats, const int* members, int n, double out[4]) {
  if (n == 1) {
    std::memcpy(out, quats + 4 * members[0], 4 * sizeof(double));
    return;
  }
  double S[16] = {0};
  for (int m = 0; m < n; ++m) {
    const double* q = quats + 4 * members[m];
    const double L[3][4] = {{q[3], q[2], -q[1], q[0]}, {-q[2], q[3], q[0], q[1]}, {q[1], -q[0], q[3], q[2]}};
    for (int r = 0; r < 3; ++r)
      for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) S[i * 4 + j] += L[r][i] * L[r][j];
  }
  double v[4];
  SmallestEigenvector4(S, v);
  out[0] = -v[0];
  out[1] = -v[1];
  out[2] = -v[2];
  out[3] = v[3];
}

int TransformationRansac(const double* quats, const double* positions, int n, int num_iterations,
                         double thr_rad, double thr_m, uint32_t seed, int rng_mapping,
                         double out_quat[4], double out_pos[3], int* inlier_indices) {
  if (n == 1) {
    std::memcpy(out_quat, quats, 4 * sizeof(double));
    std::memcpy(out_pos, positions, 3 * sizeof(double));
    inlier_indices[0] = 0;  // the reference leaves inlier_indices untouched here (:128-132)
    return 1;
  }
  RansacRng rng(seed, rng_mapping);
  std::vector<int> best{0};
  for (int it = 0; it < num_iterations; ++it) {
    const int s = UniformIndex(&rng, static_cast<uint32_t>(n), rng_mapping);
    std::vector<int> inl;
    for (int j = 0; j < n; ++j)
      if (PoseIsInlier(quats + 4 * s, positions + 3 * s, quats + 4 * j, positions + 3 * j, thr_rad, thr_m))
        inl.push_back(j);
    if (inl.size() > best.size()) best.swap(inl);
  }
  LsAverageQuaternion(quats, best.data(), static_cast<int>(best.size()), out_quat);
  double p[3] = {0, 0, 0};
  for (int i : best)
    for (int k = 0; k < 3; ++k) p[k] += positions[3 * i + k];
  for (int k = 0; k < 3; ++k) out_pos[k] = p[k] / static_cast<double>(best.size());
  for (size_t i = 0; i < best.size(); ++i) inlier_indices[i] = best[i];
  return static_cast<int>(best.size());
}

// loop-detector-node.cc:944-955: keep only the yaw of a rotation (quaternion x, y, z, w in/out).
void YawOnly(const double q[4], double out[4]) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double R00 = 1 - 2 * (y * y + z * z), R10 = 2 * (x * y + z * w), R20 = 2 * (x * z - y * w);
  const double R01 = 2 * (x * y - z * w), R11 = 1 - 2 * (x * x + z * z);
  const double pitch = std::atan2(-R20, std::sqrt(R00 * R00 + R10 * R10));
  double yaw;
  if (std::fabs(std::cos(pitch)) > 1.0e-12) {
    yaw = std::atan2(R10 / std::cos(pitch), R00 / std::cos(pitch));
  } else {
    yaw = 0.0;
  }
  (void)R01;
  (void)R11;
  // Eigen::Quaterniond(RollPitchYawToRotationMatrix(0, 0, yaw)) — Eigen's matrix -> quaternion conversion
  // (Eigen/src/Geometry/Quaternion.h, quaternionbase_assign_impl<Other, 3, 3>): w >= 0 while the trace is
  // positive, otherwise the largest diagonal entry (R22 = 1 here) becomes the positive component.
  const double c = std::cos(yaw), s = std::sin(yaw);
  double t = c + c + 1.0;
  out[0] = 0.0;
  out[1] = 0.0;
  if (t > 0.0) {
    t = std::sqrt(t + 1.0);
    out[3] = 0.5 * t;
    t = 0.5 / t;
    out[2] = (s + s) * t;
  } else {
    t = std::sqrt(1.0 - c - c + 1.0);
    out[2] = 0.5 * t;
    t = 0.5 / t;
    out[3] = (s + s) * t;
  }
}

}  // namespace lc_oracle
