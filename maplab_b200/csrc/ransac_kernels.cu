// Kernel 4 of the loop-closure path: batched geometric verification, one warp per query vertex.
//   LoopClosureHandler::handleLoopClosure gates      loop-closure-handler/src/loop-closure-handler.cc:235-480
//   PnpPoseEstimator::absoluteMultiPoseRansacPinholeCam  aslam_cv_geometric_vision/src/pnp-pose-estimator.cc:75-132, :193-280
//   opengv Ransac / sampler                          opengv/include/opengv/sac/implementation/Ransac.hpp:44-143,
//                                                    SampleConsensusProblem.hpp:62-82, :165-205
//   GP3P model + disambiguation + scoring            opengv/src/sac_problems/absolute_pose/AbsolutePoseSacProblem.cpp:111-199,
//                                                    opengv/src/absolute_pose/modules/main.cpp:375-436
//   getBestStructureMatchForEveryKeypoint            loop-closure-handler/src/inlier-index-with-reprojection-error.cc:7-51
//
// The sample sequence of opengv's RANSAC does not depend on the data (persistent partial
// Fisher-Yates over an explicit int stream), so the warp draws kHyp samples ahead and evaluates
// them speculatively; the sequential bookkeeping (skips, best-so-far with strict >, adaptive k) is
// then replayed in sample order. Per hypothesis:
//   * Groebner elimination: the micro-op program the oracle executes sequentially
//     (gp3p_program.inc) re-scheduled offline into waves of independent three-address ops
//     (gp3p_schedule.inc, gen_gp3p_schedule.py; dead ops removed, bit-identical results). The 32
//     lanes execute one 32-operation step at a time on a slot array in shared memory: 188 steps
//     instead of 12 253 dependent operations;
//   * 8x8 eigenvalues (Hessenberg + Francis QR) by one lane per hypothesis, then one lane per
//     (hypothesis, eigenvalue) for inverse iteration, Cayley back-substitution and the
//     disambiguation score; inlier counting with the lanes striding over the correspondences.
// fp64 throughout; this file is compiled with -fmad=false so every expression rounds exactly
// like the SSE2 build of the reference/oracle.
#include <climits>
#include <cmath>
#include <vector>

#include "detector.h"

namespace mlc {
namespace {
#include "gp3p_program.inc"
#include "gp3p_schedule.inc"

enum {
  MOP_DIVSUB = 0,
  MOP_DIV,
  MOP_NEGDIV,
  MOP_FACTOR_DIV,
  MOP_ZERO,
  MOP_SUBMUL,
  MOP_FACTOR_LOAD,
  MOP_FACTOR_INV,
  MOP_SCALE
};

constexpr int N8 = 8;
constexpr int kWarpsPerBlock = 4;
constexpr int kHyp = 16;      // hypothesis slots per problem (default)
constexpr int kHypMax = 32;   // ... when the batch is small enough to speculate deeper for free
constexpr int kRansacGroups = 6;  // problem groups pipelined on separate streams
constexpr int kHypFirst = 10; // speculated in the first round (k is still unknown) when the batch fills the GPU; 8 / 9 / 10 / 11 / 12 / 14: 1.39 / 1.23 / 1.24 / 1.25 / 1.26 / 1.33 ms per 1000 problems

struct RansacArgs {
  int first_hyp;                // hypotheses speculated per problem in the first round
  int grouped_by_keypoint;      // correspondences of one (camera, keypoint) are contiguous
  int hyp_slots;                // hypothesis slots per problem (kHyp or kHypMax)
  int64_t num_problems;
  const int64_t* offsets;
  const double* keypoints;      // 2 per correspondence
  const int32_t* camera_index;
  const int32_t* keypoint_index;
  const double* landmarks;      // 3 per correspondence
  const mlc_camera* cams;
  int num_cams;
  double threshold;             // 1 - cos(atan(sigma / mean focal)), computed on the host
  double log_one_minus_p;       // log(1 - 0.99)
  int min_inlier_count, max_iterations;
  double min_inlier_ratio;
  const int32_t* rnd_stream;    // uniform_int_distribution<int>(0, INT_MAX) draws, host generated
  int rnd_len;
  const uint4* wops;            // wave-scheduled ops {op | d<<16, a | b<<16, c | e<<16, 0}
  const unsigned short* wave_offsets;
  const short* init_table;      // GP3P_W_INIT (compacted slots)
  const short* action;          // GP3P_W_ACTION
  double* bearings;             // 3 per correspondence
  int32_t* shuffled;            // per correspondence
  mlc_pose_result* results;
  uint8_t* inlier_flags;        // may be null
};

// ---------------------------------------------------------------- GP3P elimination (per warp)
// One hypothesis, all 32 lanes. S = slot array of this warp in shared memory. fvp = {f, v, p}
// (3 x 9 doubles, column per point). Leaves the 6x8 action-matrix rows in M (row-major 8x8).
__device__ void Gp3pEliminateWarp(const RansacArgs& a, const double* fvp, double* S, double* M, int lane) {
  for (int s = lane; s < GP3P_W_NUM_SLOTS; s += 32) S[s] = 0.0;
  __syncwarp();
  for (int e = lane; e < GP3P_NUM_INIT; e += 32) {
    const short* in = a.init_table + e * 18;
    double acc = 0.0;
    const int nt = in[1];
    for (int t = 0; t < nt; ++t) {
      const short* tm = in + 2 + 4 * t;
      const double term = static_cast<double>(tm[0]) * fvp[tm[1] * 9 + tm[3] * 3 + tm[2]];
      acc = (t == 0) ? term : acc + term;
    }
    if (in[0] >= 0) S[in[0]] = acc;  // -1: value nobody reads
  }
  __syncwarp();
  // one step = 32 mutually independent operations (padded), one per lane; the next step's
  // operation word is fetched while the current one executes
  uint4 next = __ldg(a.wops + lane);
  for (int w = 0; w < GP3P_W_NUM_WAVES; ++w) {
    const uint4 m = next;
    if (w + 1 < GP3P_W_NUM_WAVES) next = __ldg(a.wops + (w + 1) * 32 + lane);
    const unsigned op = m.x & 0xFFFFu, d = m.x >> 16, x = m.y & 0xFFFFu, y = m.y >> 16,
                   c = m.z & 0xFFFFu, e = m.z >> 16;
    double val;
    switch (op) {
      case 0: val = S[x] / S[y] - S[c] / S[e]; break;
      case 1: val = S[x] / S[y]; break;
      case 2: val = -S[x] / S[y]; break;
      case 3: val = 0.0; break;
      case 4: val = S[c] - S[x] * S[y]; break;
      case 5: val = S[x] * S[y]; break;
      case 6: val = S[x]; break;
      case 7: val = 1.0 / S[x]; break;
      default: val = 0.0; break;  // 8: padding, nothing is stored
    }
    // destinations are slots that were free before this step (or the lane's own first operand,
    // updated in place), so no barrier is needed between the reads and the writes of one step
    if (op != 8u) S[d] = val;
    __syncwarp();
  }
  for (int i = lane; i < 64; i += 32) {
    const int r = i >> 3, c = i & 7;
    double v = 0.0;
    if (r < 6) {
      const int slot = a.action[i];
      v = (slot >= 0) ? -S[slot] : -0.0;
    } else if ((r == 6 && c == 0) || (r == 7 && c == 6)) {
      v = 1.0;
    }
    M[i] = v;
  }
  __syncwarp();
}

// ---------------------------------------------------------------- 8x8 real eigen-solver
// Householder Hessenberg reduction + Francis shifted QR (EISPACK orthes / hqr scheme), exactly
// the operation sequence of the oracle (oracle/pnp.cc).
__device__ __forceinline__ double SignOf(double a, double b) { return b >= 0.0 ? fabs(a) : -fabs(a); }

struct Cx {
  double re, im;
};
__device__ __forceinline__ Cx CxMul(Cx a, Cx b) {
  return Cx{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
__device__ __forceinline__ Cx CxSub(Cx a, Cx b) { return Cx{a.re - b.re, a.im - b.im}; }
__device__ __forceinline__ Cx CxDiv(Cx a, Cx b) {
  const double den = b.re * b.re + b.im * b.im;
  return Cx{(a.re * b.re + a.im * b.im) / den, (a.im * b.re - a.re * b.im) / den};
}
__device__ __forceinline__ double CxAbs2(Cx a) { return a.re * a.re + a.im * a.im; }

// Eigenvector by inverse iteration (complex LU with partial pivoting); only component ratios are
// consumed downstream.
__device__ void InverseIteration(const double M[N8][N8], Cx lambda, Cx vec[N8]) {
  Cx A[N8][N8];
  double norm = 0.0;
  for (int i = 0; i < N8; ++i)
    for (int j = 0; j < N8; ++j) {
      A[i][j] = Cx{M[i][j], 0.0};
      norm = fmax(norm, fabs(M[i][j]));
    }
  if (norm == 0.0) norm = 1.0;
  for (int i = 0; i < N8; ++i) A[i][i] = CxSub(A[i][i], lambda);
  const double tiny = norm * 2.220446049250313e-16;
  const double tiny2 = tiny * tiny;
  int perm[N8];
  for (int i = 0; i < N8; ++i) perm[i] = i;
  for (int c = 0; c < N8; ++c) {
    int piv = c;
    double best = CxAbs2(A[c][c]);
    for (int r = c + 1; r < N8; ++r) {
      const double ab = CxAbs2(A[r][c]);
      if (ab > best) {
        best = ab;
        piv = r;
      }
    }
    if (piv != c) {
      for (int j = 0; j < N8; ++j) {
        const Cx tmp = A[c][j];
        A[c][j] = A[piv][j];
        A[piv][j] = tmp;
      }
      const int tp = perm[c];
      perm[c] = perm[piv];
      perm[piv] = tp;
    }
    if (CxAbs2(A[c][c]) < tiny2) A[c][c] = Cx{tiny, 0.0};
    for (int r = c + 1; r < N8; ++r) {
      const Cx mlt = CxDiv(A[r][c], A[c][c]);
      A[r][c] = mlt;
      for (int j = c + 1; j < N8; ++j) A[r][j] = CxSub(A[r][j], CxMul(mlt, A[c][j]));
    }
  }
  Cx x[N8];
  for (int i = 0; i < N8; ++i) x[i] = Cx{1.0, 0.0};
  for (int iter = 0; iter < 3; ++iter) {
    Cx b[N8];
    if (iter == 0) {
      for (int i = 0; i < N8; ++i) b[i] = x[i];
    } else {
      for (int i = 0; i < N8; ++i) b[i] = x[perm[i]];
      for (int i = 0; i < N8; ++i)
        for (int j = 0; j < i; ++j) b[i] = CxSub(b[i], CxMul(A[i][j], b[j]));
    }
    for (int i = N8 - 1; i >= 0; --i) {
      Cx s = b[i];
      for (int j = i + 1; j < N8; ++j) s = CxSub(s, CxMul(A[i][j], x[j]));
      x[i] = CxDiv(s, A[i][i]);
    }
    double mx = 0.0;
    for (int i = 0; i < N8; ++i) mx = fmax(mx, fmax(fabs(x[i].re), fabs(x[i].im)));
    if (mx == 0.0 || !isfinite(mx)) break;
    for (int i = 0; i < N8; ++i) {
      x[i].re = x[i].re / mx;
      x[i].im = x[i].im / mx;
    }
  }
  for (int i = 0; i < N8; ++i) vec[i] = x[i];
}

// opengv::math::cayley2rot, row-major.
__device__ void Cayley2Rot(const double c[3], double R[9]) {
  const double c0 = c[0] * c[0], c1 = c[1] * c[1], c2 = c[2] * c[2];
  const double scale = 1 + c0 + c1 + c2;
  R[0] = 1 + c0 - c1 - c2;
  R[1] = 2 * (c[0] * c[1] - c[2]);
  R[2] = 2 * (c[0] * c[2] + c[1]);
  R[3] = 2 * (c[0] * c[1] + c[2]);
  R[4] = 1 - c0 + c1 - c2;
  R[5] = 2 * (c[1] * c[2] - c[0]);
  R[6] = 2 * (c[0] * c[2] - c[1]);
  R[7] = 2 * (c[1] * c[2] + c[0]);
  R[8] = 1 - c0 - c1 + c2;
  const double inv = 1 / scale;
  for (int i = 0; i < 9; ++i) R[i] = inv * R[i];
}

struct Problem {
  const double* bearings;  // 3 x n
  const int32_t* cam_idx;
  const double* points;    // 3 x n
  const mlc_camera* cams;
  int n;
};

// 1 - cos(angle) reprojection score (AbsolutePoseSacProblem.cpp:165-199).
__device__ double Distance(const Problem& pb, const double T[12], int i) {
  double tinv[3];
  for (int a = 0; a < 3; ++a) tinv[a] = -(T[0 * 4 + a] * T[3] + T[1 * 4 + a] * T[7] + T[2 * 4 + a] * T[11]);
  const double* P = pb.points + 3 * i;
  double body[3];
  for (int a = 0; a < 3; ++a)
    body[a] = T[0 * 4 + a] * P[0] + T[1 * 4 + a] * P[1] + T[2 * 4 + a] * P[2] + tinv[a] * 1.0;
  const mlc_camera& c = pb.cams[pb.cam_idx[i]];
  const double d[3] = {body[0] - c.t_B_C[0], body[1] - c.t_B_C[1], body[2] - c.t_B_C[2]};
  double r[3];
  for (int a = 0; a < 3; ++a)
    r[a] = c.R_B_C[0 * 3 + a] * d[0] + c.R_B_C[1 * 3 + a] * d[1] + c.R_B_C[2 * 3 + a] * d[2];
  const double nrm = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  r[0] /= nrm;
  r[1] /= nrm;
  r[2] /= nrm;
  const double* b = pb.bearings + 3 * i;
  return 1.0 - (r[0] * b[0] + r[1] * b[1] + r[2] * b[2]);
}

// One real-ish eigenvalue of the action matrix -> candidate pose (gp3p_main, main.cpp:396-433)
// and its disambiguation score on the 4th sample point.
__device__ void SolutionForEigenvalue(const Problem& pb, const double* Ms, double wr, double wi,
                                      const double* fvp, int fourth, double sol[12], double* score) {
  double M[N8][N8];
  for (int r = 0; r < N8; ++r)
    for (int c = 0; c < N8; ++c) M[r][c] = Ms[r * 8 + c];
  const double* f = fvp;
  const double* v = fvp + 9;
  const double* p = fvp + 18;
  Cx V[N8];
  InverseIteration(M, Cx{wr, wi}, V);
  double cay[3], n[3];
  for (int i = 0; i < 3; ++i) {
    cay[2 - i] = CxDiv(V[i + 4], V[7]).re;
    n[2 - i] = CxDiv(V[i + 1], V[7]).re;
  }
  double Rt[9];
  Cayley2Rot(cay, Rt);
  double R[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[i * 3 + j] = Rt[j * 3 + i];
  double center_cam[3] = {0, 0, 0}, center_world[3] = {0, 0, 0};
  for (int i = 0; i < 3; ++i) {
    double tmp[3], w[3];
    for (int k = 0; k < 3; ++k) w[k] = n[i] * f[i * 3 + k] + v[i * 3 + k];
    for (int k = 0; k < 3; ++k) tmp[k] = R[k * 3 + 0] * w[0] + R[k * 3 + 1] * w[1] + R[k * 3 + 2] * w[2];
    for (int k = 0; k < 3; ++k) {
      center_cam[k] = center_cam[k] + tmp[k];
      center_world[k] = center_world[k] + p[i * 3 + k];
    }
  }
  for (int k = 0; k < 3; ++k) {
    sol[k * 4 + 0] = R[k * 3 + 0];
    sol[k * 4 + 1] = R[k * 3 + 1];
    sol[k * 4 + 2] = R[k * 3 + 2];
    sol[k * 4 + 3] = center_world[k] / 3 - center_cam[k] / 3;
  }
  *score = Distance(pb, sol, fourth);
}

// Forward models + Jacobians of the two iteratively inverted distortions.
// RadTanDistortion::distortUsingExternalCoefficients (distortion-radtan.cc:14-62)
__device__ void DistortRadTan(const double* k, double* px, double* py, double F[4]) {
  const double x = *px, y = *py;
  const double k1 = k[0], k2 = k[1], p1 = k[2], p2 = k[3];
  const double mx2 = x * x, my2 = y * y, mxy = x * y, rho2 = mx2 + my2;
  const double rad = k1 * rho2 + k2 * rho2 * rho2;
  F[0] = 1.0 + rad + 2.0 * k1 * mx2 + 4.0 * k2 * rho2 * mx2 + 2.0 * p1 * y + 6.0 * p2 * x;
  F[1] = 2.0 * k1 * mxy + 4.0 * k2 * rho2 * mxy + 2.0 * p1 * x + 2.0 * p2 * y;
  F[2] = F[1];
  F[3] = 1.0 + rad + 2.0 * k1 * my2 + 4.0 * k2 * rho2 * my2 + 2.0 * p2 * x + 6.0 * p1 * y;
  *px = x + (x * rad + 2.0 * p1 * mxy + p2 * (rho2 + 2.0 * mx2));
  *py = y + (y * rad + 2.0 * p2 * mxy + p1 * (rho2 + 2.0 * my2));
}
// EquidistantDistortion::distortUsingExternalCoefficients (distortion-equidistant.cc:14-103)
__device__ void DistortEquidistant(const double* k, double* px, double* py, double F[4]) {
  const double x = *px, y = *py;
  const double k1 = k[0], k2 = k[1], k3 = k[2], k4 = k[3];
  const double x2 = x * x, y2 = y * y, r = sqrt(x2 + y2);
  if (r < 1e-10) {
    F[0] = F[1] = F[2] = F[3] = 0.0;
    return;
  }
  const double theta = atan(r), theta2 = theta * theta, theta4 = theta2 * theta2, theta6 = theta2 * theta4,
               theta8 = theta4 * theta4;
  const double poly = k1 * theta2 + k2 * theta4 + k3 * theta6 + k4 * theta8 + 1.0;
  const double thetad = theta * (1 + k1 * theta2 + k2 * theta4 + k3 * theta6 + k4 * theta8);
  const double theta3 = theta2 * theta, theta5 = theta4 * theta, theta7 = theta6 * theta;
  const double s = x2 + y2, s1 = x2 + y2 + 1.0, r3 = pow(x2 + y2, 3.0 / 2.0);
  // the reference's MATLAB-generated Jacobian, same grouping: c = x for d/du, y for d/dv
  auto dpoly = [&](double c) {
    return (k2 * c * theta3 / r * 4.0) / s1 + (k3 * c * theta5 / r * 6.0) / s1 + (k4 * c * theta7 / r * 8.0) / s1 +
           (k1 * c * theta / r * 2.0) / s1;
  };
  F[0] = theta / r * poly + x * theta / r * dpoly(x) + (x2 * poly) / (s * s1) - x2 * theta / r3 * poly;
  F[1] = x * theta / r * dpoly(y) + (x * y * poly) / (s * s1) - x * y * theta / r3 * poly;
  F[2] = y * theta / r * dpoly(x) + (x * y * poly) / (s * s1) - x * y * theta / r3 * poly;
  F[3] = theta / r * poly + y * theta / r * dpoly(y) + (y2 * poly) / (s * s1) - y2 * theta / r3 * poly;
  const double scaling = (r > 1e-8) ? thetad / r : 1.0;
  *px = x * scaling;
  *py = y * scaling;
}
// {RadTan,Equidistant}Distortion::undistortUsingExternalCoefficients (distortion-radtan.cc:96-118,
// distortion-equidistant.cc:144-173): Gauss-Newton on the forward model, du = (F^T F)^-1 F^T e, at most 30
// iterations, stop once e.e <= --acv_inv_distortion_tolerance (1e-8) AFTER applying the step.
__device__ void UndistortIterative(int model, const double* k, double* px, double* py) {
  const double yx = *px, yy = *py;
  if (model == 3 && yx * yx + yy * yy < 1e-6) return;  // equidistant: unchanged around the image centre
  double bx = yx, by = yy;
  for (int i = 0; i < 30; ++i) {
    double tx = bx, ty = by, F[4];
    if (model == 2) DistortRadTan(k, &tx, &ty, F);
    else DistortEquidistant(k, &tx, &ty, F);
    const double ex = yx - tx, ey = yy - ty;
    // F^T F (2 x 2 coefficient-wise product), its closed-form inverse, times F^T, times e
    const double a = F[0] * F[0] + F[2] * F[2], b = F[0] * F[1] + F[2] * F[3], c = F[1] * F[0] + F[3] * F[2],
                 d = F[1] * F[1] + F[3] * F[3];
    const double invdet = 1.0 / (a * d - b * c);
    const double i00 = d * invdet, i01 = -b * invdet, i10 = -c * invdet, i11 = a * invdet;
    const double m00 = i00 * F[0] + i01 * F[1], m01 = i00 * F[2] + i01 * F[3], m10 = i10 * F[0] + i11 * F[1],
                 m11 = i10 * F[2] + i11 * F[3];
    bx += m00 * ex + m01 * ey;
    by += m10 * ex + m11 * ey;
    if (ex * ex + ey * ey <= 1e-8) break;
  }
  *px = bx;
  *py = by;
}

// PinholeCamera::backProject3 + bearing normalisation (camera-pinhole.cc:47-63; undistort of
// distortion-fisheye.cc:119-143, distortion-radtan.cc:96-118, distortion-equidistant.cc:144-173).
__device__ void BackProject3(const mlc_camera& c, const double* kp, double* b) {
  double x = (kp[0] - c.cu) / c.fu;
  double y = (kp[1] - c.cv) / c.fv;
  if (c.distortion == 1) {
    const double w = c.dist[0];
    const double mul2tanwby2 = tan(w / 2.0) * 2.0;
    const double r_d = sqrt(x * x + y * y);
    if (!(mul2tanwby2 == 0 || r_d == 0)) {
      if (fabs(r_d * w) <= (89.0 * 3.14159265358979323846 / 180.0)) {
        const double r_u = tan(r_d * w) / (r_d * mul2tanwby2);
        x *= r_u;
        y *= r_u;
      }
    }
  } else if (c.distortion == 2 || c.distortion == 3) {
    UndistortIterative(c.distortion, c.dist, &x, &y);
  }
  const double nrm = sqrt(x * x + y * y + 1.0);
  b[0] = x / nrm;
  b[1] = y / nrm;
  b[2] = 1.0 / nrm;
}

// ---------------------------------------------------------------- hypothesis-parallel RANSAC
// Per round every still-running problem speculates kHyp hypotheses. The stages are separate
// kernels so that each one exposes its natural parallelism across ALL problems of the batch:
//   sample (thread / problem) -> eliminate (warp / hypothesis) -> eigenvalues (thread / hypothesis)
//   -> candidate poses (thread / (hypothesis, eigenvalue)) -> select + count inliers
//   (warp / hypothesis) -> bookkeeping replay (thread / problem).
struct ProblemState {  // sequential state of opengv::sac::Ransac::computeModel for one problem
  int iterations, best, skipped, stream_pos;
  int have_model, done, pad0, pad1;
  double k;
  double best_model[12];
  int best_sel[4];
};
struct Hypothesis {
  double fvp[27];
  double M[64];
  double wr[N8], wi[N8];
  double cand_T[N8][12];
  double cand_score[N8];
  double model[12];
  int cand_valid[N8];
  int sel[4];
  int active, eig_ok, model_ok, count;
};

__device__ __forceinline__ Problem MakeProblem(const RansacArgs& a, int64_t pi) {
  const int64_t off = a.offsets[pi];
  Problem pb;
  pb.bearings = a.bearings + 3 * off;
  pb.cam_idx = a.camera_index + off;
  pb.points = a.landmarks + 3 * off;
  pb.cams = a.cams;
  pb.n = static_cast<int>(a.offsets[pi + 1] - off);
  return pb;
}

// bearings, identity shuffle, initial state. One warp per problem.
__global__ void __launch_bounds__(128) ransac_init_kernel(RansacArgs a, ProblemState* st) {
  const int lane = threadIdx.x & 31;
  const int64_t pi = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  if (pi >= a.num_problems) return;
  const int64_t off = a.offsets[pi];
  const int n = static_cast<int>(a.offsets[pi + 1] - off);
  for (int i = lane; i < n; i += 32) {
    if (a.inlier_flags) a.inlier_flags[off + i] = 0;
    if (n >= a.min_inlier_count) {
      BackProject3(a.cams[a.camera_index[off + i]], a.keypoints + 2 * (off + i), a.bearings + 3 * (off + i));
      a.shuffled[off + i] = i;
    }
  }
  if (lane == 0) {
    ProblemState s;
    s.iterations = 0;
    s.best = -INT_MAX;
    s.skipped = 0;
    s.stream_pos = 0;
    s.have_model = 0;
    s.done = 0;
    s.pad0 = s.pad1 = 0;
    s.k = 1.0;
    for (int i = 0; i < 12; ++i) s.best_model[i] = 0.0;
    for (int i = 0; i < 4; ++i) s.best_sel[i] = -1;
    if (n < a.min_inlier_count) s.done = 1;  // handleLoopClosure bails before RANSAC
    if (!s.done && n < 4) {                  // getSamples cannot draw 4 unique indices
      s.iterations = INT_MAX;
      s.done = 1;
    }
    st[pi] = s;
  }
}

// Draw the next samples of every running problem (persistent partial Fisher-Yates). First round:
// a.first_hyp samples; later rounds: as many as the adaptive bound k still asks for (<= kHyp).
__global__ void __launch_bounds__(128) ransac_sample_kernel(RansacArgs a, ProblemState* st, Hypothesis* hyp) {
  const int64_t pi = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (pi >= a.num_problems) return;
  ProblemState& s = st[pi];
  Hypothesis* h = hyp + pi * a.hyp_slots;
  const int max_skip = a.max_iterations * 10;
  int want = a.first_hyp;
  if (s.stream_pos > 0) {
    const double need = ceil(s.k - static_cast<double>(s.iterations));
    want = need >= static_cast<double>(a.hyp_slots) ? a.hyp_slots : (need <= 1.0 ? 1 : static_cast<int>(need));
    const int room = a.max_iterations + 1 - s.iterations;  // iterations_ > max_iterations_ stops the loop
    if (want > room) want = room > 1 ? room : 1;
  }
  const bool run = !s.done && (static_cast<double>(s.iterations) < s.k && s.skipped < max_skip) &&
                   (s.stream_pos + 4 * want <= a.rnd_len);
  if (!run) {
    s.done = 1;
    for (int t = 0; t < a.hyp_slots; ++t) h[t].active = 0;
    return;
  }
  const int64_t off = a.offsets[pi];
  const int n = static_cast<int>(a.offsets[pi + 1] - off);
  int32_t* shuf = a.shuffled + off;
  if (n >= 4) {
    // The partial Fisher-Yates only ever moves entries through positions 0..3: keep those four in
    // registers, so that one sample costs one round of independent loads instead of a chain of
    // dependent global read-modify-writes.
    int32_t s4[4] = {shuf[0], shuf[1], shuf[2], shuf[3]};
    for (int t = 0; t < want; ++t) {
      int r4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) r4[i] = __ldg(a.rnd_stream + s.stream_pos + 4 * t + i);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int j = i + r4[i] % (n - i);
        if (j < 4) {
          const int32_t vi = s4[i];
          int32_t vj = s4[0];
#pragma unroll
          for (int q = 1; q < 4; ++q) vj = (j == q) ? s4[q] : vj;
#pragma unroll
          for (int q = 0; q < 4; ++q) s4[q] = (q == j) ? vi : s4[q];
          s4[i] = vj;  // after the write of position j: i == j leaves the entry unchanged
        } else {
          const int32_t tmp = shuf[j];
          shuf[j] = s4[i];
          s4[i] = tmp;
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) h[t].sel[i] = s4[i];
      h[t].active = 1;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) shuf[i] = s4[i];
  } else {
    for (int t = 0; t < want; ++t) {
      for (int i = 0; i < 4; ++i) {
        const int j = i + a.rnd_stream[s.stream_pos + 4 * t + i] % (n - i);
        const int32_t tmp = shuf[i];
        shuf[i] = shuf[j];
        shuf[j] = tmp;
      }
      for (int i = 0; i < 4; ++i) h[t].sel[i] = shuf[i];
      h[t].active = 1;
    }
  }
  for (int t = want; t < a.hyp_slots; ++t) h[t].active = 0;
  s.stream_pos += 4 * want;
}

// Groebner elimination: one warp per hypothesis, slot array in shared memory.
__global__ void __launch_bounds__(kWarpsPerBlock * 32) gp3p_eliminate_kernel(RansacArgs a, Hypothesis* hyp, int64_t num_hyp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  double* S = reinterpret_cast<double*>(smem_raw) + static_cast<size_t>(wib) * (GP3P_W_NUM_SLOTS + 27 + 64);
  double* fvp = S + GP3P_W_NUM_SLOTS;
  double* M = fvp + 27;
  const int64_t warps = static_cast<int64_t>(gridDim.x) * kWarpsPerBlock;
  for (int64_t hi = static_cast<int64_t>(blockIdx.x) * kWarpsPerBlock + wib; hi < num_hyp; hi += warps) {
    Hypothesis& h = hyp[hi];
    if (!h.active) continue;
    const Problem pb = MakeProblem(a, hi / a.hyp_slots);
    // f (bearing rotated into the body frame), v (camera offset), p (world point) of the 3 points
    if (lane < 27) {
      const int which = lane / 9, i = (lane % 9) / 3, kk = lane % 3;
      const int ci = h.sel[i];
      const mlc_camera& c = pb.cams[pb.cam_idx[ci]];
      double val;
      if (which == 0) {
        const double* b = pb.bearings + 3 * ci;
        val = c.R_B_C[kk * 3 + 0] * b[0] + c.R_B_C[kk * 3 + 1] * b[1] + c.R_B_C[kk * 3 + 2] * b[2];
      } else if (which == 1) {
        val = c.t_B_C[kk];
      } else {
        val = pb.points[3 * ci + kk];
      }
      fvp[lane] = val;
      h.fvp[lane] = val;
    }
    __syncwarp();
    Gp3pEliminateWarp(a, fvp, S, M, lane);
    for (int i = lane; i < 64; i += 32) h.M[i] = M[i];
    __syncwarp();
  }
}

// Eigenvalues of the 8x8 action matrix: EIGHT lanes per hypothesis. The matrix lives in shared
// memory; the row/column update loops — whose iterations are independent — are dealt one row or column per
// lane. The SCALAR work of orthes/hqr is dealt to the lanes too wherever its pieces are independent (same
// operations on the same operands, only evaluated by another lane, so still bit-identical to the sequential
// Hessenberg()/HqrEigenvalues()): the divisions of one step (p/x q/x r/x; p/s q/s r/s q/p r/p; the
// Householder vector a(i,m-1)/scale) run as ONE division over the lanes and are handed round with shuffles,
// and the two searches (small subdiagonal l, start row m of the double shift) test all candidates at once and
// take the first hit in the order of the sequential loop. fp64 division and sqrt are ~15-instruction sequences;
// the redundant version spent most of its instructions there.
#define A_(i, j) a[(i) * 8 + (j)]
#define GSHFL(v, src) __shfl_sync(gm, (v), (src), 8)
__device__ void HessenbergGroup(double* a, int L, unsigned gm) {
  const int high = N8 - 1;
#pragma unroll
  for (int m = 1; m <= high - 1; ++m) {
    double scale = 0.0;
#pragma unroll
    for (int i = m; i <= high; ++i) scale += fabs(A_(i, m - 1));
    if (scale != 0.0) {
      const double mine = (L >= m) ? A_(L, m - 1) / scale : 0.0;  // ort[L]
      double ort[N8];
#pragma unroll
      for (int i = 0; i < N8; ++i) ort[i] = (i >= m) ? GSHFL(mine, i) : 0.0;
      double h = 0.0;
#pragma unroll
      for (int i = high; i >= m; --i) h += ort[i] * ort[i];
      double g = sqrt(h);
      if (ort[m] > 0) g = -g;
      h -= ort[m] * g;
      ort[m] -= g;
      __syncwarp(gm);  // everyone has read column m-1
      if (L >= m) {    // column L
        const int j = L;
        double fsum = 0.0;
#pragma unroll
        for (int i = high; i >= m; --i) fsum += ort[i] * A_(i, j);
        fsum /= h;
#pragma unroll
        for (int i = m; i <= high; ++i) A_(i, j) -= fsum * ort[i];
      }
      __syncwarp(gm);
      {  // row L
        const int i = L;
        double fsum = 0.0;
#pragma unroll
        for (int j = high; j >= m; --j) fsum += ort[j] * A_(i, j);
        fsum /= h;
#pragma unroll
        for (int j = m; j <= high; ++j) A_(i, j) -= fsum * ort[j];
      }
      __syncwarp(gm);
      if (L == 0) A_(m, m - 1) = scale * g;
      __syncwarp(gm);
    }
  }
  if (L >= 2)
    for (int j = 0; j < L - 1; ++j) A_(L, j) = 0.0;
  __syncwarp(gm);
}

__device__ bool HqrGroup(double* a, double* wr, double* wi, int L, unsigned gm) {
  const int base = __ffs(gm) - 1;  // first lane of the group
  int nn, m, l, k, j, its, i, mmin;
  double z, y, x, w, t, s, r = 0, q = 0, p = 0, anorm = 0.0;
  for (i = 0; i < N8; i++)
    for (j = (i - 1 > 0 ? i - 1 : 0); j < N8; j++) anorm += fabs(A_(i, j));
  nn = N8 - 1;
  t = 0.0;
  while (nn >= 0) {
    its = 0;
    do {
      {
        // small subdiagonal element: candidates l = nn, nn-1, .. 1 on lanes 0, 1, ..; the first hit counts
        const int lj = nn - L;
        bool hit = false;
        if (lj >= 1) {
          double ss = fabs(A_(lj - 1, lj - 1)) + fabs(A_(lj, lj));
          if (ss == 0.0) ss = anorm;
          hit = fabs(A_(lj, lj - 1)) + ss == ss;
        }
        const unsigned hits = (__ballot_sync(gm, hit) >> base) & 0xFFu;
        l = 0;
        if (hits) {
          l = nn - (__ffs(hits) - 1);
          __syncwarp(gm);
          if (L == 0) A_(l, l - 1) = 0.0;
          __syncwarp(gm);
        }
      }
      x = A_(nn, nn);
      if (l == nn) {
        if (L == 0) {
          wr[nn] = x + t;
          wi[nn] = 0.0;
        }
        nn--;
      } else {
        y = A_(nn - 1, nn - 1);
        w = A_(nn, nn - 1) * A_(nn - 1, nn);
        if (l == (nn - 1)) {
          p = 0.5 * (y - x);
          q = p * p + w;
          z = sqrt(fabs(q));
          x += t;
          if (L == 0) {
            if (q >= 0.0) {
              z = p + SignOf(z, p);
              wr[nn - 1] = wr[nn] = x + z;
              if (z != 0.0) wr[nn] = x - w / z;
              wi[nn - 1] = wi[nn] = 0.0;
            } else {
              wr[nn - 1] = wr[nn] = x + p;
              wi[nn - 1] = -(wi[nn] = z);
            }
          }
          nn -= 2;
        } else {
          if (its == 60) return false;
          if (its == 10 || its == 20) {
            t += x;
            __syncwarp(gm);
            if (L <= nn) A_(L, L) -= x;
            __syncwarp(gm);
            s = fabs(A_(nn, nn - 1)) + fabs(A_(nn - 1, nn - 2));
            y = x = 0.75 * s;
            w = -0.4375 * s * s;
          }
          ++its;
          {
            // start row of the double shift: candidates m = nn-2, nn-3, .. l on lanes 0, 1, ..; the sequential
            // loop stops at the first (largest) m with m == l or a negligible subdiagonal product
            const int mj = nn - 2 - L;
            bool stop = false;
            double pj = 0.0, qj = 0.0, rj = 0.0;
            if (mj >= l) {
              const double zz = A_(mj, mj);
              const double rr = x - zz;
              double ss = y - zz;
              pj = (rr * ss - w) / A_(mj + 1, mj) + A_(mj, mj + 1);
              qj = A_(mj + 1, mj + 1) - zz - rr - ss;
              rj = A_(mj + 2, mj + 1);
              ss = fabs(pj) + fabs(qj) + fabs(rj);
              pj /= ss;
              qj /= ss;
              rj /= ss;
              if (mj == l) {
                stop = true;
              } else {
                const double uu = fabs(A_(mj, mj - 1)) * (fabs(qj) + fabs(rj));
                const double vv = fabs(pj) * (fabs(A_(mj - 1, mj - 1)) + fabs(zz) + fabs(A_(mj + 1, mj + 1)));
                stop = uu + vv == vv;
              }
            }
            const unsigned stops = (__ballot_sync(gm, stop) >> base) & 0xFFu;
            const int js = __ffs(stops) - 1;  // the lane with m == l always stops
            m = nn - 2 - js;
            p = GSHFL(pj, js);
            q = GSHFL(qj, js);
            r = GSHFL(rj, js);
          }
          __syncwarp(gm);
          if (L >= m + 2 && L <= nn) {
            A_(L, L - 2) = 0.0;
            if (L != (m + 2)) A_(L, L - 3) = 0.0;
          }
          __syncwarp(gm);
          for (k = m; k <= nn - 1; k++) {
            if (k != m) {
              p = A_(k, k - 1);
              q = A_(k + 1, k - 1);
              r = 0.0;
              if (k != (nn - 1)) r = A_(k + 2, k - 1);
              if ((x = fabs(p) + fabs(q) + fabs(r)) != 0.0) {
                // p /= x, q /= x, r /= x on lanes 0, 1, 2
                const double quo = (L == 0 ? p : (L == 1 ? q : r)) / x;
                p = GSHFL(quo, 0);
                q = GSHFL(quo, 1);
                r = GSHFL(quo, 2);
              }
            }
            if ((s = SignOf(sqrt(p * p + q * q + r * r), p)) != 0.0) {
              __syncwarp(gm);  // all lanes have read column k-1
              if (L == 0) {
                if (k == m) {
                  if (l != m) A_(k, k - 1) = -A_(k, k - 1);
                } else {
                  A_(k, k - 1) = -s * x;
                }
              }
              p += s;
              {
                // x = p / s, y = q / s, z = r / s, q /= p, r /= p on lanes 0 .. 4
                const double num = (L == 0 ? p : ((L == 1 || L == 3) ? q : r));
                const double den = L < 3 ? s : p;
                const double quo = num / den;
                x = GSHFL(quo, 0);
                y = GSHFL(quo, 1);
                z = GSHFL(quo, 2);
                q = GSHFL(quo, 3);
                r = GSHFL(quo, 4);
              }
              __syncwarp(gm);
              if (L >= k && L <= nn) {  // row modification, column L
                j = L;
                double pp = A_(k, j) + q * A_(k + 1, j);
                if (k != (nn - 1)) {
                  pp += r * A_(k + 2, j);
                  A_(k + 2, j) -= pp * z;
                }
                A_(k + 1, j) -= pp * y;
                A_(k, j) -= pp * x;
              }
              __syncwarp(gm);
              mmin = nn < k + 3 ? nn : k + 3;
              if (L >= l && L <= mmin) {  // column modification, row L
                i = L;
                double pp = x * A_(i, k) + y * A_(i, k + 1);
                if (k != (nn - 1)) {
                  pp += z * A_(i, k + 2);
                  A_(i, k + 2) -= pp * r;
                }
                A_(i, k + 1) -= pp * q;
                A_(i, k) -= pp;
              }
              __syncwarp(gm);
            }
          }
        }
      }
    } while (l < nn - 1);
  }
  return true;
}
#undef GSHFL
#undef A_

// GPW = hypotheses (8-lane groups) per warp. The QR iteration is data dependent, so the groups of
// one warp serialise each other's control flow; the kernel is latency bound with idle issue slots,
// hence fewer groups per warp (idle lanes) finish sooner.
template <int GPW>
__global__ void __launch_bounds__(128) gp3p_eigen_kernel(Hypothesis* hyp, int64_t num_hyp, int* task_count,
                                                         int32_t* tasks) {
  __shared__ double s_a[4 * GPW][64];
  __shared__ double s_w[4 * GPW][16];
  const int lane = threadIdx.x & 31;
  if (lane >= 8 * GPW) return;
  const int L = lane & 7;
  const int g = (threadIdx.x >> 5) * GPW + (lane >> 3);  // group within the block
  const unsigned gm = 0xFFu << (lane & 24);
  const int64_t hi = static_cast<int64_t>(blockIdx.x) * (4 * GPW) + g;
  if (hi >= num_hyp) return;   // whole group leaves together
  Hypothesis& h = hyp[hi];
  if (!h.active) return;
  double* a = s_a[g];
  bool finite = true;
  for (int c = 0; c < N8; ++c) {
    const double v = h.M[L * 8 + c];
    a[L * 8 + c] = v;
    if (!isfinite(v)) finite = false;
  }
  __syncwarp(gm);  // the rows written above are read by the other lanes of the group (a vote is no memory barrier)
  const bool ok_in = __all_sync(gm, finite);
  bool ok = ok_in;
  if (ok) {
    HessenbergGroup(a, L, gm);
    ok = HqrGroup(a, s_w[g], s_w[g] + 8, L, gm);
    __syncwarp(gm);
    if (ok) {
      h.wr[L] = s_w[g][L];
      h.wi[L] = s_w[g][8 + L];
    }
  }
  if (L == 0) h.eig_ok = ok ? 1 : 0;
  // The eigenvalues that yield a candidate pose (main.cpp:400: imag < 1e-4, no fabs) go to a compact task list:
  // they are ~a quarter of the (hypothesis, eigenvalue) pairs, and a candidate kernel over all pairs ran every
  // warp through the long solve for a few live lanes. The order of the list varies from run to run, the results
  // do not (a task only writes its own slot).
  const bool valid = ok && s_w[g][8 + L] < 0.0001;
  h.cand_valid[L] = valid ? 1 : 0;
  const int base_lane = lane & 24;
  const unsigned vmask = (__ballot_sync(gm, valid) >> base_lane) & 0xFFu;
  int first = 0;
  if (L == 0 && vmask) first = atomicAdd(task_count, __popc(vmask));
  first = __shfl_sync(gm, first, 0, 8);
  if (valid) tasks[first + __popc(vmask & ((1u << L) - 1u))] = static_cast<int32_t>(hi * 8 + L);
}

// Candidate pose + disambiguation score: one thread per task = (hypothesis, eigenvalue) listed by the eigenvalue kernel.
__global__ void __launch_bounds__(64) gp3p_candidate_kernel(RansacArgs a, Hypothesis* hyp, const int* task_count,
                                                            const int32_t* tasks) {
  const int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (t >= *task_count) return;
  const int32_t task = tasks[t];
  const int64_t hi = task >> 3;
  const int c = task & 7;
  Hypothesis& h = hyp[hi];
  const Problem pb = MakeProblem(a, hi / a.hyp_slots);
  double sol[12], score;
  SolutionForEigenvalue(pb, h.M, h.wr[c], h.wi[c], h.fvp, h.sel[3], sol, &score);
  for (int i = 0; i < 12; ++i) h.cand_T[c][i] = sol[i];
  h.cand_score[c] = score;
}

// Disambiguation (AbsolutePoseSacProblem.cpp:133-163) + countWithinDistance: one warp per hypothesis.
__global__ void __launch_bounds__(128) ransac_score_kernel(RansacArgs a, Hypothesis* hyp, int64_t num_hyp) {
  const int lane = threadIdx.x & 31;
  const int64_t hi = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  if (hi >= num_hyp) return;
  Hypothesis& h = hyp[hi];
  if (!h.active) return;
  // every lane evaluates the (cheap) selection redundantly
  int num = 0, first = -1, min_index = -1;
  double min_score = 1000000.0;
  for (int c = 0; c < N8; ++c) {
    if (!h.cand_valid[c]) continue;
    if (num == 0) first = c;
    ++num;
    if (h.cand_score[c] < min_score) {  // smallest score on the 4th point, first wins
      min_score = h.cand_score[c];
      min_index = c;
    }
  }
  const int pick = (num == 1) ? first : min_index;  // a single solution is accepted as is
  int count = 0;
  if (pick >= 0) {
    const Problem pb = MakeProblem(a, hi / a.hyp_slots);
    double T[12];
    for (int i = 0; i < 12; ++i) T[i] = h.cand_T[pick][i];
    for (int i = lane; i < pb.n; i += 32)
      if (Distance(pb, T, i) < a.threshold) ++count;
    for (int o = 16; o > 0; o >>= 1) count += __shfl_xor_sync(0xffffffffu, count, o);
    if (lane < 12) h.model[lane] = T[lane];
  }
  if (lane == 0) {
    h.model_ok = pick >= 0 ? 1 : 0;
    h.count = count;
  }
}

// Replay of the sequential bookkeeping of Ransac::computeModel (Ransac.hpp:64-128) in sample
// order; counts the problems that need another round.
__global__ void __launch_bounds__(128) ransac_update_kernel(RansacArgs a, ProblemState* st, const Hypothesis* hyp,
                                                            int* remaining) {
  const int64_t pi = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  int still = 0;
  if (pi < a.num_problems) {
    ProblemState& s = st[pi];
    if (!s.done) {
      const Hypothesis* h = hyp + pi * a.hyp_slots;
      const int n = static_cast<int>(a.offsets[pi + 1] - a.offsets[pi]);
      const int max_skip = a.max_iterations * 10;
      // verdicts of all speculated hypotheses first (independent loads), then the sequential replay
      // on registers; the winning model is copied once at the end
      int best_t = -1;
      bool stop = false;
      for (int t0 = 0; t0 < a.hyp_slots && !stop; t0 += 8) {
        int h_active[8], h_ok[8], h_count[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          h_active[u] = h[t0 + u].active;
          h_ok[u] = h[t0 + u].model_ok;
          h_count[u] = h[t0 + u].count;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (stop) break;
          if (!h_active[u]) {
            stop = true;
            break;
          }
          if (!(static_cast<double>(s.iterations) < s.k && s.skipped < max_skip)) {
            s.done = 1;
            stop = true;
            break;
          }
          if (!h_ok[u]) {
            ++s.skipped;
            continue;
          }
          if (h_count[u] > s.best) {
            s.best = h_count[u];
            s.have_model = 1;
            best_t = t0 + u;
            const double w = static_cast<double>(s.best) / static_cast<double>(n);
            double p_no_outliers = 1.0 - pow(w, 4.0);
            p_no_outliers = fmax(2.220446049250313e-16, p_no_outliers);
            p_no_outliers = fmin(1.0 - 2.220446049250313e-16, p_no_outliers);
            s.k = a.log_one_minus_p / log(p_no_outliers);
          }
          ++s.iterations;
          if (s.iterations > a.max_iterations) {
            s.done = 1;
            stop = true;
            break;
          }
        }
      }
      if (best_t >= 0) {
        for (int i = 0; i < 12; ++i) s.best_model[i] = h[best_t].model[i];
        for (int i = 0; i < 4; ++i) s.best_sel[i] = h[best_t].sel[i];
      }
      if (!s.done && !(static_cast<double>(s.iterations) < s.k && s.skipped < max_skip)) s.done = 1;
      still = s.done ? 0 : 1;
    }
  }
  const int any = __syncthreads_count(still);
  if (threadIdx.x == 0 && any) atomicAdd(remaining, any);  // one per block; round control only
}

// selectWithinDistance on the best model, best inlier per keypoint, handleLoopClosure gates.
__global__ void __launch_bounds__(128) ransac_finalize_kernel(RansacArgs a, const ProblemState* st) {
  const int lane = threadIdx.x & 31;
  const int64_t pi = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  if (pi >= a.num_problems) return;
  const ProblemState& s = st[pi];
  const int64_t off = a.offsets[pi];
  const Problem pb = MakeProblem(a, pi);
  const int n = pb.n;
  mlc_pose_result res;
  res.accepted = 0;
  res.ransac_success = 0;
  res.num_inliers = 0;
  res.num_ransac_inliers = 0;
  res.iterations = s.iterations;
  for (int i = 0; i < 4; ++i) res.model_indices[i] = -1;
  res.pad_ = 0;
  res.inlier_ratio = 0.0;
  for (int i = 0; i < 12; ++i) res.T_G_I[i] = 0.0;
  if (s.have_model) {
    double best_model[12];
    for (int i = 0; i < 12; ++i) best_model[i] = s.best_model[i];
    res.ransac_success = 1;
    for (int i = 0; i < 12; ++i) res.T_G_I[i] = best_model[i];
    for (int i = 0; i < 4; ++i) res.model_indices[i] = s.best_sel[i];
    int total = 0, my_best = 0;
    for (int i = lane; i < n; i += 32) {
      const double di = Distance(pb, best_model, i);
      if (!(di < a.threshold)) continue;
      ++total;
      // best inlier per (camera, keypoint): smallest score, first index wins on ties
      bool is_best = true;
      const int cam = a.camera_index[off + i], kp = a.keypoint_index[off + i];
      if (a.grouped_by_keypoint) {
        // correspondences of one (camera, keypoint) are adjacent (the fused path emits them in
        // canonical order): only the run around i has to be looked at
        for (int dir = -1; dir <= 1 && is_best; dir += 2) {
          for (int j = i + dir; j >= 0 && j < n && is_best; j += dir) {
            if (a.keypoint_index[off + j] != kp || a.camera_index[off + j] != cam) break;
            const double dj = Distance(pb, best_model, j);
            if (!(dj < a.threshold)) continue;
            if (dj < di || (dj == di && j < i)) is_best = false;
          }
        }
      } else {
        for (int j = 0; j < n && is_best; ++j) {
          if (j == i || a.keypoint_index[off + j] != kp || a.camera_index[off + j] != cam) continue;
          const double dj = Distance(pb, best_model, j);
          if (!(dj < a.threshold)) continue;
          if (dj < di || (dj == di && j < i)) is_best = false;
        }
      }
      if (is_best) ++my_best;
      if (a.inlier_flags) a.inlier_flags[off + i] = is_best ? 3 : 1;
    }
    for (int o = 16; o > 0; o >>= 1) {
      total += __shfl_xor_sync(0xffffffffu, total, o);
      my_best += __shfl_xor_sync(0xffffffffu, my_best, o);
    }
    res.num_ransac_inliers = total;
    res.num_inliers = my_best;
  }
  if (res.num_inliers >= a.min_inlier_count) {
    res.inlier_ratio = static_cast<double>(res.num_inliers) / static_cast<double>(n);
    if (!(res.inlier_ratio < a.min_inlier_ratio)) res.accepted = 1;
  }
  if (lane == 0) a.results[pi] = res;
}

// std::mt19937 + libstdc++ uniform_int_distribution<int>(0, INT_MAX).
class HostRng {
 public:
  HostRng(uint32_t seed, int mapping) : idx_(624), mapping_(mapping) {
    mt_[0] = seed;
    for (int i = 1; i < 624; ++i) mt_[i] = 1812433253u * (mt_[i - 1] ^ (mt_[i - 1] >> 30)) + i;
  }
  int Next() {
    if (mapping_ == 1) return static_cast<int>(U32() >> 1);  // libstdc++ >= 11: one draw, x >> 1
    uint32_t x;
    do {
      x = U32();
    } while (x >= 0x80000000u);  // libstdc++ <= 10: rejection, scaling 1
    return static_cast<int>(x);
  }

 private:
  uint32_t U32() {
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
  uint32_t mt_[624];
  int idx_, mapping_;
};

struct ProgramDevice {
  uint4* wops = nullptr;
  unsigned short* wave_offsets = nullptr;
  short* init_table = nullptr;
  short* action = nullptr;
};
ProgramDevice g_program;  // per process (one GPU per process)

cudaError_t EnsureProgram() {
  if (g_program.wops) return cudaSuccess;
  std::vector<uint4> packed(GP3P_W_NUM_OPS);
  for (int i = 0; i < GP3P_W_NUM_OPS; ++i) {
    const unsigned short* m = GP3P_W_OPS[i];
    packed[i] = make_uint4(m[0] | (static_cast<unsigned>(m[1]) << 16), m[2] | (static_cast<unsigned>(m[3]) << 16),
                           m[4] | (static_cast<unsigned>(m[5]) << 16), 0u);
  }
  cudaError_t e;
  if ((e = cudaMalloc(&g_program.wops, sizeof(uint4) * GP3P_W_NUM_OPS)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&g_program.wave_offsets, sizeof(GP3P_W_WAVE_OFFSETS))) != cudaSuccess) return e;
  if ((e = cudaMalloc(&g_program.init_table, sizeof(GP3P_W_INIT))) != cudaSuccess) return e;
  if ((e = cudaMalloc(&g_program.action, sizeof(GP3P_W_ACTION))) != cudaSuccess) return e;
  if ((e = cudaMemcpy(g_program.wops, packed.data(), sizeof(uint4) * GP3P_W_NUM_OPS, cudaMemcpyHostToDevice)) != cudaSuccess) return e;
  if ((e = cudaMemcpy(g_program.wave_offsets, GP3P_W_WAVE_OFFSETS, sizeof(GP3P_W_WAVE_OFFSETS), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
  if ((e = cudaMemcpy(g_program.init_table, GP3P_W_INIT, sizeof(GP3P_W_INIT), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
  return cudaMemcpy(g_program.action, GP3P_W_ACTION, sizeof(GP3P_W_ACTION), cudaMemcpyHostToDevice);
}

}  // namespace

namespace {
// T_I_I_ransac = T_G_I(map)^-1 * T_G_I_ransac; position norm and Eigen's AngleAxis(quaternion)
// angle, 2 * atan2(|q.vec|, |q.w|), of the relative rotation (matrix -> quaternion as Eigen does).
bool DeltaPoseGate(const double* A, const double* B, double max_pos_m, double max_rot_deg) {
  double R[9], d[3], p[3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0.0;
      for (int k = 0; k < 3; ++k) s += A[k * 4 + i] * B[k * 4 + j];
      R[i * 3 + j] = s;
    }
  for (int k = 0; k < 3; ++k) d[k] = B[k * 4 + 3] - A[k * 4 + 3];
  for (int i = 0; i < 3; ++i) p[i] = A[0 * 4 + i] * d[0] + A[1 * 4 + i] * d[1] + A[2 * 4 + i] * d[2];
  const double dp = std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
  double w, v[3];
  const double t = R[0] + R[4] + R[8];
  if (t > 0.0) {
    double r = std::sqrt(t + 1.0);
    w = 0.5 * r;
    r = 0.5 / r;
    v[0] = (R[7] - R[5]) * r;
    v[1] = (R[2] - R[6]) * r;
    v[2] = (R[3] - R[1]) * r;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[i * 3 + i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    double r = std::sqrt(R[i * 3 + i] - R[j * 3 + j] - R[k * 3 + k] + 1.0);
    v[i] = 0.5 * r;
    r = 0.5 / r;
    w = (R[k * 3 + j] - R[j * 3 + k]) * r;
    v[j] = (R[j * 3 + i] + R[i * 3 + j]) * r;
    v[k] = (R[k * 3 + i] + R[i * 3 + k]) * r;
  }
  const double n = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  const double deg = (n != 0.0 ? 2.0 * std::atan2(n, std::fabs(w)) : 0.0) * (180.0 / 3.14159265358979323846);
  return (max_pos_m < 0.0 || dp <= max_pos_m) && (max_rot_deg < 0.0 || deg <= max_rot_deg);
}
}  // namespace

bool Detector::PnpRansacBatch(const mlc_ransac_settings& rs, const mlc_camera* cams, int num_cams,
                              int64_t num_problems, const int64_t* offsets, const double* keypoints,
                              const int32_t* camera_index, const int32_t* keypoint_index,
                              const double* landmarks, mlc_pose_result* results,
                              uint8_t* inlier_flags, std::string* err) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  if (num_problems == 0) return true;
  const int64_t total = offsets[num_problems];
  for (int64_t p = 0; p < num_problems; ++p) {
    if (offsets[p + 1] < offsets[p]) {
      *err = "offsets must be non-decreasing";
      return false;
    }
  }
  for (int64_t i = 0; i < total; ++i) {
    if (camera_index[i] < 0 || camera_index[i] >= num_cams) {
      *err = "camera index out of range";
      return false;
    }
  }
  DevBuf& b_in = d_ransac_[0];
  size_t o = 0;
  auto place = [&](size_t bytes) {
    const size_t at = o;
    o = (o + bytes + 255) & ~static_cast<size_t>(255);
    return at;
  };
  const size_t o_off = place(sizeof(int64_t) * (num_problems + 1));
  const size_t o_kp = place(sizeof(double) * 2 * total);
  const size_t o_ci = place(sizeof(int32_t) * total);
  const size_t o_ki = place(sizeof(int32_t) * total);
  const size_t o_lm = place(sizeof(double) * 3 * total);
  if (!Cuda(b_in.Reserve(o), "alloc", err)) return false;
  unsigned char* din = b_in.as<unsigned char>();
  auto up = [&](size_t at, const void* src, size_t bytes) {
    return bytes == 0 ||
           Cuda(cudaMemcpyAsync(din + at, src, bytes, cudaMemcpyHostToDevice, stream_), "H2D", err);
  };
  if (!up(o_off, offsets, sizeof(int64_t) * (num_problems + 1)) || !up(o_kp, keypoints, sizeof(double) * 2 * total) ||
      !up(o_ci, camera_index, sizeof(int32_t) * total) || !up(o_ki, keypoint_index, sizeof(int32_t) * total) ||
      !up(o_lm, landmarks, sizeof(double) * 3 * total))
    return false;
  return RansacOnDevice(rs, cams, num_cams, num_problems, total,
                        reinterpret_cast<const int64_t*>(din + o_off),
                        reinterpret_cast<const double*>(din + o_kp),
                        reinterpret_cast<const int32_t*>(din + o_ci),
                        reinterpret_cast<const int32_t*>(din + o_ki),
                        reinterpret_cast<const double*>(din + o_lm), results, inlier_flags, err);
}

// Kernel 4 on correspondences that already live on the device; results (and flags) come back to
// the host and the stream is synchronised.
bool Detector::RansacOnDevice(const mlc_ransac_settings& rs, const mlc_camera* cams, int num_cams,
                              int64_t num_problems, int64_t total, const int64_t* d_offsets,
                              const double* d_keypoints, const int32_t* d_camera_index,
                              const int32_t* d_keypoint_index, const double* d_landmarks,
                              mlc_pose_result* results, uint8_t* inlier_flags, std::string* err) {
  if (num_problems == 0) return true;
  if (rs.num_ransac_iters < 0 || rs.num_ransac_iters > 100000 || num_cams <= 0) {
    *err = "bad RANSAC settings";
    return false;
  }
  if (!Cuda(EnsureProgram(), "upload GP3P program", err)) return false;
  // absoluteMultiPoseRansacPinholeCam: threshold from the mean focal length (pnp-pose-estimator.cc:75-132)
  double focal = 0;
  for (int i = 0; i < num_cams; ++i) focal += (cams[i].fu + cams[i].fv);
  focal /= (2.0 * static_cast<double>(num_cams));
  const double threshold = 1.0 - std::cos(std::atan(rs.ransac_pixel_sigma / focal));
  // draws: 4 per attempted sample; attempts <= max_iterations + 1 counted + 10 * max_iterations skipped
  const int rnd_len = 4 * (11 * rs.num_ransac_iters + 2 + 2 * kHypMax);  // worst case + speculation slack
  if (rnd_seed_ != rs.seed || rnd_mapping_ != rs.rng_mapping || static_cast<int>(rnd_host_.size()) != rnd_len) {
    rnd_host_.resize(rnd_len);
    HostRng rng(rs.seed, rs.rng_mapping);
    for (int i = 0; i < rnd_len; ++i) rnd_host_[i] = rng.Next();
    rnd_seed_ = rs.seed;
    rnd_mapping_ = rs.rng_mapping;
  }
  DevBuf &b_scr = d_ransac_[1], &b_hyp = d_ransac_[2], &b_out = d_ransac_[3];
  // small batches leave the GPU idle between the latency-bound rounds: deeper speculation
  // (32 instead of 16 slots per problem) then saves whole rounds
  int hyp_slots = (num_problems * kHypMax <= static_cast<int64_t>(sm_count_) * 60) ? kHypMax : kHyp;
  if (const char* env = getenv("MLC_RANSAC_SLOTS")) {
    const int v = atoi(env);
    if (v == kHyp || v == kHypMax) hyp_slots = v;
  }
  const int64_t num_hyp = num_problems * hyp_slots;
  const size_t o_cam = 0;
  const size_t o_rnd = (sizeof(mlc_camera) * num_cams + 255) & ~static_cast<size_t>(255);
  const size_t o_bear = (o_rnd + sizeof(int32_t) * rnd_len + 255) & ~static_cast<size_t>(255);
  const size_t o_shuf = (o_bear + sizeof(double) * 3 * total + 255) & ~static_cast<size_t>(255);
  if (!Cuda(b_scr.Reserve(o_shuf + sizeof(int32_t) * total + 256), "alloc", err) ||
      !Cuda(b_hyp.Reserve(sizeof(Hypothesis) * num_hyp + sizeof(ProblemState) * num_problems + 256 /* group counters */ +
                          sizeof(int32_t) * 8 * num_hyp /* candidate tasks */), "alloc", err) ||
      !Cuda(b_out.Reserve(sizeof(mlc_pose_result) * num_problems + total + 512), "alloc", err))
    return false;
  unsigned char* scr = b_scr.as<unsigned char>();
  Hypothesis* d_hyp = b_hyp.as<Hypothesis>();
  ProblemState* d_state = reinterpret_cast<ProblemState*>(b_hyp.as<unsigned char>() + sizeof(Hypothesis) * num_hyp);
  int* d_remaining = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(d_state) + sizeof(ProblemState) * num_problems);
  if (!Cuda(cudaMemcpyAsync(scr + o_cam, cams, sizeof(mlc_camera) * num_cams, cudaMemcpyHostToDevice, stream_), "H2D", err) ||
      !Cuda(cudaMemcpyAsync(scr + o_rnd, rnd_host_.data(), sizeof(int32_t) * rnd_len, cudaMemcpyHostToDevice, stream_), "H2D", err))
    return false;
  RansacArgs a;
  a.grouped_by_keypoint = corr_grouped_ ? 1 : 0;
  corr_grouped_ = false;  // one-shot, set by the fused path
  a.num_problems = num_problems;
  a.offsets = d_offsets;
  a.keypoints = d_keypoints;
  a.camera_index = d_camera_index;
  a.keypoint_index = d_keypoint_index;
  a.landmarks = d_landmarks;
  a.cams = reinterpret_cast<const mlc_camera*>(scr + o_cam);
  a.num_cams = num_cams;
  a.threshold = threshold;
  a.log_one_minus_p = std::log(1.0 - 0.99);
  a.min_inlier_count = rs.min_inlier_count;
  a.max_iterations = rs.num_ransac_iters;
  a.min_inlier_ratio = rs.min_inlier_ratio;
  a.rnd_stream = reinterpret_cast<const int32_t*>(scr + o_rnd);
  a.rnd_len = rnd_len;
  a.wops = g_program.wops;
  a.wave_offsets = g_program.wave_offsets;
  a.init_table = g_program.init_table;
  a.action = g_program.action;
  a.bearings = reinterpret_cast<double*>(scr + o_bear);
  a.shuffled = reinterpret_cast<int32_t*>(scr + o_shuf);
  a.results = b_out.as<mlc_pose_result>();
  a.inlier_flags = inlier_flags ? b_out.as<uint8_t>() + sizeof(mlc_pose_result) * num_problems : nullptr;
  const size_t smem = sizeof(double) * (GP3P_W_NUM_SLOTS + 27 + 64) * kWarpsPerBlock;
  if (!Cuda(cudaFuncSetAttribute(gp3p_eliminate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(smem)), "ransac smem", err))
    return false;
  int elim_per_sm = 4;  // resident CTAs per SM: bounded by the slot arrays in shared memory
  if (!Cuda(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&elim_per_sm, gp3p_eliminate_kernel,
                                                          kWarpsPerBlock * 32, smem), "ransac occupancy", err))
    return false;
  if (elim_per_sm < 1) elim_per_sm = 1;
  // First round: k is unknown, so speculation depth trades wasted hypotheses against one more
  // latency-bound round. A batch that does not fill the resident hypothesis slots of the
  // elimination kernel speculates the full kHyp for free.
  const int64_t resident = static_cast<int64_t>(sm_count_) * elim_per_sm * kWarpsPerBlock;
  a.hyp_slots = hyp_slots;
  a.first_hyp = (num_problems * hyp_slots <= resident + resident / 2) ? hyp_slots : kHypFirst;
  if (const char* env = getenv("MLC_RANSAC_FIRST_HYP")) {
    const int v = atoi(env);
    if (v >= 1 && v <= hyp_slots) a.first_hyp = v;
  }
  // The problems are dealt into groups that run the round pipeline on their own streams: the
  // stages are latency-bound kernels with long tails, and with staggered groups the tail of one
  // group's stage overlaps the next stage of another group.
  int groups = num_problems >= 384 ? 3 : (num_problems >= 96 ? 2 : 1);  // tuned on B200, 1000 problems
  if (const char* env = getenv("MLC_RANSAC_GROUPS")) {
    const int v = atoi(env);
    if (v >= 1 && v <= kRansacGroups) groups = v;
  }
  if (groups > num_problems) groups = static_cast<int>(num_problems);
  int* d_remaining_all = d_remaining;  // two counters per group: [2 g] running problems, [2 g + 1] candidate tasks
  int32_t* d_tasks = reinterpret_cast<int32_t*>(reinterpret_cast<unsigned char*>(d_remaining) + 256);
  // pinned, so that the per-round read-back does not block the host while it enqueues the other groups
  if (!h_remaining_ && !Cuda(cudaHostAlloc(&h_remaining_, sizeof(int) * 8, cudaHostAllocDefault), "pinned alloc", err))
    return false;
  int* remaining_h = h_remaining_;
  bool active[kRansacGroups] = {};
  RansacArgs ga[kRansacGroups];
  ProblemState* g_state[kRansacGroups];
  Hypothesis* g_hyp[kRansacGroups];
  cudaStream_t g_stream[kRansacGroups];
  if (!Cuda(cudaEventRecord(ev_ransac_[kRansacGroups], stream_), "event", err)) return false;
  for (int g = 0; g < groups; ++g) {
    const int64_t p0 = num_problems * g / groups, p1 = num_problems * (g + 1) / groups;
    ga[g] = a;
    ga[g].num_problems = p1 - p0;
    ga[g].offsets = a.offsets + p0;     // correspondence arrays stay absolute
    ga[g].results = a.results + p0;
    g_state[g] = d_state + p0;
    g_hyp[g] = d_hyp + p0 * hyp_slots;
    g_stream[g] = g == 0 ? stream_ : ransac_stream_[g - 1];
    active[g] = p1 > p0;
    if (g > 0 && !Cuda(cudaStreamWaitEvent(g_stream[g], ev_ransac_[kRansacGroups], 0), "wait", err)) return false;
  }
  auto blocks_of = [](int64_t items, int per_block) { return static_cast<unsigned>((items + per_block - 1) / per_block); };
  for (int g = 0; g < groups; ++g) {
    if (!active[g]) continue;
    ransac_init_kernel<<<blocks_of(ga[g].num_problems * 32, 128), 128, 0, g_stream[g]>>>(ga[g], g_state[g]);
    CountLaunch();
  }
  int eigen_gpw = 4;  // 1 / 2 / 4 groups per warp measured: 1.54 / 1.47 / 1.47 ms RANSAC at 1000 problems
  if (const char* env = getenv("MLC_EIGEN_GPW")) {
    const int v = atoi(env);
    if (v == 1 || v == 2 || v == 4) eigen_gpw = v;
  }
  const int max_rounds = 11 * rs.num_ransac_iters + 4;  // >= one consumed sample per round
  // One round of one group: six kernels, the count of its still-running problems to the host, an event.
  auto enqueue_round = [&](int g, int round) -> bool {
    const int64_t np = ga[g].num_problems, nh = np * hyp_slots;
    // later rounds hold few running problems: one hypothesis per warp then (no group of the warp waits for the
    // data-dependent control flow of another, and the GPU has idle warps to spare)
    int later = 1;
    if (const char* env = getenv("MLC_EIGEN_GPW_LATER")) later = atoi(env);
    const int gpw = (round == 0 || (later != 1 && later != 2 && later != 4)) ? eigen_gpw : later;
    cudaStream_t st = g_stream[g];
    const unsigned elim_blocks = static_cast<unsigned>(std::min<int64_t>(
        (nh + kWarpsPerBlock - 1) / kWarpsPerBlock, static_cast<int64_t>(sm_count_) * elim_per_sm));
    int* counters = d_remaining_all + 2 * g;
    int32_t* tasks = d_tasks + (g_hyp[g] - d_hyp) * 8;
    if (!Cuda(cudaMemsetAsync(counters, 0, 2 * sizeof(int), st), "memset", err)) return false;
    ransac_sample_kernel<<<blocks_of(np, 128), 128, 0, st>>>(ga[g], g_state[g], g_hyp[g]);
    gp3p_eliminate_kernel<<<elim_blocks, kWarpsPerBlock * 32, smem, st>>>(ga[g], g_hyp[g], nh);
    if (gpw == 1) gp3p_eigen_kernel<1><<<blocks_of(nh, 4), 128, 0, st>>>(g_hyp[g], nh, counters + 1, tasks);
    else if (gpw == 2) gp3p_eigen_kernel<2><<<blocks_of(nh, 8), 128, 0, st>>>(g_hyp[g], nh, counters + 1, tasks);
    else gp3p_eigen_kernel<4><<<blocks_of(nh, 16), 128, 0, st>>>(g_hyp[g], nh, counters + 1, tasks);
    // sized for the worst case; blocks beyond the task count leave at once
    gp3p_candidate_kernel<<<blocks_of(nh * 8, 64), 64, 0, st>>>(ga[g], g_hyp[g], counters + 1, tasks);
    ransac_score_kernel<<<blocks_of(nh * 32, 128), 128, 0, st>>>(ga[g], g_hyp[g], nh);
    ransac_update_kernel<<<blocks_of(np, 128), 128, 0, st>>>(ga[g], g_state[g], g_hyp[g], counters);
    for (int i = 0; i < 6; ++i) CountLaunch();
    return Cuda(cudaMemcpyAsync(&remaining_h[g], counters, sizeof(int), cudaMemcpyDeviceToHost, st), "D2H", err) &&
           Cuda(cudaEventRecord(ev_ransac_[7 + g], st), "event", err);
  };
  // The groups advance independently: the host polls the groups' round events, and a group whose round is
  // through either gets its next round or is finalized at once — no group waits for another one's round. A
  // large batch practically always needs a second round (the first one speculates first_hyp < k hypotheses
  // for most problems), so two rounds are enqueued up front: a round that finds every problem done costs six
  // near-empty launches, a host round trip between the rounds costs the idle GPU more.
  int upfront = (groups >= 2 && a.first_hyp < hyp_slots) ? 2 : 1;
  if (const char* env = getenv("MLC_RANSAC_UPFRONT")) {
    const int v = atoi(env);
    if (v >= 1 && v <= 4) upfront = v;
  }
  int rounds[kRansacGroups] = {};
  int live = 0;
  for (int r = 0; r < upfront; ++r)
    for (int g = 0; g < groups; ++g) {
      if (!active[g]) continue;
      if (!enqueue_round(g, rounds[g])) return false;
      ++rounds[g];
      if (r == 0) ++live;
    }
  while (live > 0) {
    for (int g = 0; g < groups; ++g) {
      if (!active[g]) continue;
      const cudaError_t q = cudaEventQuery(ev_ransac_[7 + g]);
      if (q == cudaErrorNotReady) continue;
      if (!Cuda(q, "ransac round", err)) return false;
      if (remaining_h[g] == 0) {
        active[g] = false;  // this group is done: finalize it right away on its stream
        --live;
        ransac_finalize_kernel<<<blocks_of(ga[g].num_problems * 32, 128), 128, 0, g_stream[g]>>>(ga[g], g_state[g]);
        CountLaunch();
        if (g > 0) {
          if (!Cuda(cudaEventRecord(ev_ransac_[g - 1], g_stream[g]), "event", err) ||
              !Cuda(cudaStreamWaitEvent(stream_, ev_ransac_[g - 1], 0), "wait", err))
            return false;
        }
      } else {
        if (rounds[g] >= max_rounds) {  // cannot happen: every round consumes a sample
          *err = "RANSAC round limit reached";
          return false;
        }
        if (!enqueue_round(g, rounds[g])) return false;
        ++rounds[g];
      }
    }
  }
  cudaEventRecord(ev_stage_[5], stream_);
  if (!Cuda(cudaGetLastError(), "ransac kernels", err)) return false;
  if (!Cuda(cudaMemcpyAsync(results, a.results, sizeof(mlc_pose_result) * num_problems, cudaMemcpyDeviceToHost, stream_),
            "D2H results", err))
    return false;
  if (inlier_flags && total > 0 &&
      !Cuda(cudaMemcpyAsync(inlier_flags, a.inlier_flags, total, cudaMemcpyDeviceToHost, stream_), "D2H flags", err))
    return false;
  if (!Cuda(cudaStreamSynchronize(stream_), "pnp ransac", err)) return false;
  // Topological gate of handleLoopClosure (loop-closure-handler.cc:424-455): a host-side verdict on
  // the recovered pose against the vertex' current pose; off unless a limit is >= 0.
  const bool gate = rs.max_delta_position_m >= 0.0 || rs.max_delta_rotation_deg >= 0.0;
  const bool have = have_priors_;
  have_priors_ = false;  // consumed by this call
  if (gate) {
    if (!have || static_cast<int64_t>(priors_.size()) != 12 * num_problems) {
      *err = "max_delta_* limits need mlc_set_query_priors with one pose per query vertex";
      return false;
    }
    for (int64_t p = 0; p < num_problems; ++p)
      if (results[p].accepted &&
          !DeltaPoseGate(&priors_[12 * p], results[p].T_G_I, rs.max_delta_position_m, rs.max_delta_rotation_deg))
        results[p].accepted = 0;
  }
  return true;
}

}  // namespace mlc
