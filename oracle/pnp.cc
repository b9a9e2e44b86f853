// ORACLE (test infrastructure). Geometric verification (A17-A22).
// Reference (relative to /root/reference):
//  aslam_cv2/aslam_cv_geometric_vision/src/pnp-pose-estimator.cc:75-132, :193-280
//  aslam_cv2/aslam_cv_cameras/src/camera-pinhole.cc:47-63, src/distortion-fisheye.cc:119-143
//  dependencies/3rdparty/opengv/include/opengv/sac/implementation/Ransac.hpp:44-143
//  dependencies/3rdparty/opengv/include/opengv/sac/implementation/SampleConsensusProblem.hpp
//     :34-46 (rng), :62-82 (drawIndexSample), :86-117 (getSamples), :165-205
//  dependencies/3rdparty/opengv/src/sac_problems/absolute_pose/AbsolutePoseSacProblem.cpp
//     :111-163 (GP3P branch + disambiguation), :165-199 (distances)
//  dependencies/3rdparty/opengv/src/absolute_pose/methods.cpp:191-217 (gp3p)
//  dependencies/3rdparty/opengv/src/absolute_pose/modules/main.cpp:375-436 (gp3p_main)
//  dependencies/3rdparty/opengv/src/math/cayley.cpp:35-53
//  algorithms/loopclosure/loop-closure-handler/src/loop-closure-handler.cc:235-480
//  algorithms/loopclosure/loop-closure-handler/src/inlier-index-with-reprojection-error.cc:7-51
//
// Third-party arithmetic restated: Eigen::EigenSolver<8x8> (Hessenberg + shifted
// QR; eigenvectors here by complex inverse iteration — only the scale-invariant
// ratios V(i)/V(7) are consumed), std::mt19937 and libstdc++'s
// uniform_int_distribution<int>(0, INT_MAX) (both mappings).
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <map>

#include "lc_oracle.h"

namespace lc_oracle {
namespace {
#include "gp3p_program.inc"

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

// gp3p::init + gp3p::compute as a micro-op interpreter (see gen_gp3p_program.py).
void Gp3pEliminate(const double f[9], const double v[9], const double p[9], double* S) {
  const double* src[3] = {f, v, p};  // column-major 3x3: (i,j) -> [j*3+i]
  for (int s = 0; s < GP3P_NUM_SLOTS; ++s) S[s] = 0.0;
  for (int e = 0; e < GP3P_NUM_INIT; ++e) {
    const short* in = GP3P_INIT[e];
    double acc = 0.0;
    for (int t = 0; t < in[1]; ++t) {
      const short* tm = in + 2 + 4 * t;
      const double term = static_cast<double>(tm[0]) * src[tm[1]][tm[3] * 3 + tm[2]];
      acc = (t == 0) ? term : acc + term;
    }
    S[in[0]] = acc;
  }
  double factor = 0.0;
  for (int i = 0; i < GP3P_NUM_MOPS; ++i) {
    const short* m = GP3P_MOPS[i];
    switch (m[0]) {
      case MOP_DIVSUB:
        S[m[1]] = S[m[2]] / S[m[3]] - S[m[4]] / S[m[5]];
        break;
      case MOP_DIV:
        S[m[1]] = S[m[2]] / S[m[3]];
        break;
      case MOP_NEGDIV:
        S[m[1]] = -S[m[2]] / S[m[3]];
        break;
      case MOP_FACTOR_DIV:
        factor = S[m[1]] / S[m[2]];
        break;
      case MOP_ZERO:
        S[m[1]] = 0.0;
        break;
      case MOP_SUBMUL:
        S[m[1]] = S[m[1]] - factor * S[m[2]];
        break;
      case MOP_FACTOR_LOAD:
        factor = S[m[1]];
        break;
      case MOP_FACTOR_INV:
        factor = 1.0 / S[m[1]];
        break;
      case MOP_SCALE:
        S[m[1]] = factor * S[m[1]];
        break;
    }
  }
}

// ---- real nonsymmetric eigenvalues: Householder Hessenberg + Francis QR ----
constexpr int N8 = 8;

void Hessenberg(double H[N8][N8]) {
  double ort[N8];
  const int low = 0, high = N8 - 1;
  for (int m = low + 1; m <= high - 1; ++m) {
    double scale = 0.0;
    for (int i = m; i <= high; ++i) scale += std::fabs(H[i][m - 1]);
    if (scale != 0.0) {
      double h = 0.0;
      for (int i = high; i >= m; --i) {
        ort[i] = H[i][m - 1] / scale;
        h += ort[i] * ort[i];
      }
      double g = std::sqrt(h);
      if (ort[m] > 0) g = -g;
      h -= ort[m] * g;
      ort[m] -= g;
      for (int j = m; j < N8; ++j) {
        double fsum = 0.0;
        for (int i = high; i >= m; --i) fsum += ort[i] * H[i][j];
        fsum /= h;
        for (int i = m; i <= high; ++i) H[i][j] -= fsum * ort[i];
      }
      for (int i = 0; i <= high; ++i) {
        double fsum = 0.0;
        for (int j = high; j >= m; --j) fsum += ort[j] * H[i][j];
        fsum /= h;
        for (int j = m; j <= high; ++j) H[i][j] -= fsum * ort[j];
      }
      ort[m] = scale * ort[m];
      H[m][m - 1] = scale * g;
    }
  }
  for (int i = 2; i < N8; ++i)
    for (int j = 0; j < i - 1; ++j) H[i][j] = 0.0;
}

inline double SignOf(double a, double b) { return b >= 0.0 ? std::fabs(a) : -std::fabs(a); }

// Eigenvalues of an upper Hessenberg matrix (shifted QR, EISPACK hqr scheme).
bool HqrEigenvalues(double a[N8][N8], double wr[N8], double wi[N8]) {
  int nn, m, l, k, j, its, i, mmin;
  double z, y, x, w, v, u, t, s, r = 0, q = 0, p = 0, anorm = 0.0;
  for (i = 0; i < N8; i++)
    for (j = std::max(i - 1, 0); j < N8; j++) anorm += std::fabs(a[i][j]);
  nn = N8 - 1;
  t = 0.0;
  while (nn >= 0) {
    its = 0;
    do {
      for (l = nn; l >= 1; l--) {
        s = std::fabs(a[l - 1][l - 1]) + std::fabs(a[l][l]);
        if (s == 0.0) s = anorm;
        if (std::fabs(a[l][l - 1]) + s == s) {
          a[l][l - 1] = 0.0;
          break;
        }
      }
      x = a[nn][nn];
      if (l == nn) {
        wr[nn] = x + t;
        wi[nn--] = 0.0;
      } else {
        y = a[nn - 1][nn - 1];
        w = a[nn][nn - 1] * a[nn - 1][nn];
        if (l == (nn - 1)) {
          p = 0.5 * (y - x);
          q = p * p + w;
          z = std::sqrt(std::fabs(q));
          x += t;
          if (q >= 0.0) {
            z = p + SignOf(z, p);
            wr[nn - 1] = wr[nn] = x + z;
            if (z != 0.0) wr[nn] = x - w / z;
            wi[nn - 1] = wi[nn] = 0.0;
          } else {
            wr[nn - 1] = wr[nn] = x + p;
            wi[nn - 1] = -(wi[nn] = z);
          }
          nn -= 2;
        } else {
          if (its == 60) return false;
          if (its == 10 || its == 20) {
            t += x;
            for (i = 0; i <= nn; i++) a[i][i] -= x;
            s = std::fabs(a[nn][nn - 1]) + std::fabs(a[nn - 1][nn - 2]);
            y = x = 0.75 * s;
            w = -0.4375 * s * s;
          }
          ++its;
          for (m = (nn - 2); m >= l; m--) {
            z = a[m][m];
            r = x - z;
            s = y - z;
            p = (r * s - w) / a[m + 1][m] + a[m][m + 1];
            q = a[m + 1][m + 1] - z - r - s;
            r = a[m + 2][m + 1];
            s = std::fabs(p) + std::fabs(q) + std::fabs(r);
            p /= s;
            q /= s;
            r /= s;
            if (m == l) break;
            u = std::fabs(a[m][m - 1]) * (std::fabs(q) + std::fabs(r));
            v = std::fabs(p) *
                (std::fabs(a[m - 1][m - 1]) + std::fabs(z) + std::fabs(a[m + 1][m + 1]));
            if (u + v == v) break;
          }
          for (i = m + 2; i <= nn; i++) {
            a[i][i - 2] = 0.0;
            if (i != (m + 2)) a[i][i - 3] = 0.0;
          }
          for (k = m; k <= nn - 1; k++) {
            if (k != m) {
              p = a[k][k - 1];
              q = a[k + 1][k - 1];
              r = 0.0;
              if (k != (nn - 1)) r = a[k + 2][k - 1];
              if ((x = std::fabs(p) + std::fabs(q) + std::fabs(r)) != 0.0) {
                p /= x;
                q /= x;
                r /= x;
              }
            }
            if ((s = SignOf(std::sqrt(p * p + q * q + r * r), p)) != 0.0) {
              if (k == m) {
                if (l != m) a[k][k - 1] = -a[k][k - 1];
              } else {
                a[k][k - 1] = -s * x;
              }
              p += s;
              x = p / s;
              y = q / s;
              z = r / s;
              q /= p;
              r /= p;
              for (j = k; j <= nn; j++) {
                p = a[k][j] + q * a[k + 1][j];
                if (k != (nn - 1)) {
                  p += r * a[k + 2][j];
                  a[k + 2][j] -= p * z;
                }
                a[k + 1][j] -= p * y;
                a[k][j] -= p * x;
              }
              mmin = nn < k + 3 ? nn : k + 3;
              for (i = l; i <= mmin; i++) {
                p = x * a[i][k] + y * a[i][k + 1];
                if (k != (nn - 1)) {
                  p += z * a[i][k + 2];
                  a[i][k + 2] -= p * r;
                }
                a[i][k + 1] -= p * q;
                a[i][k] -= p;
              }
            }
          }
        }
      }
    } while (l < nn - 1);
  }
  return true;
}

// Explicit complex arithmetic (formulas fixed so that the CUDA kernel can repeat
// them operation by operation; std::complex's operator/ is implementation-defined).
struct Cx {
  double re, im;
};
inline Cx CxMul(Cx a, Cx b) { return Cx{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
inline Cx CxSub(Cx a, Cx b) { return Cx{a.re - b.re, a.im - b.im}; }
inline Cx CxDiv(Cx a, Cx b) {
  const double den = b.re * b.re + b.im * b.im;
  return Cx{(a.re * b.re + a.im * b.im) / den, (a.im * b.re - a.re * b.im) / den};
}
inline double CxAbs2(Cx a) { return a.re * a.re + a.im * a.im; }

// Eigenvector of M for eigenvalue lambda by inverse iteration (complex LU with
// partial pivoting, tiny pivots replaced). Only ratios of components are used
// downstream, so the normalisation is arbitrary.
void InverseIteration(const double M[N8][N8], Cx lambda, Cx vec[N8]) {
  Cx A[N8][N8];
  double norm = 0.0;
  for (int i = 0; i < N8; ++i)
    for (int j = 0; j < N8; ++j) {
      A[i][j] = Cx{M[i][j], 0.0};
      norm = std::fmax(norm, std::fabs(M[i][j]));
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
      const double a = CxAbs2(A[r][c]);
      if (a > best) {
        best = a;
        piv = r;
      }
    }
    if (piv != c) {
      for (int j = 0; j < N8; ++j) std::swap(A[c][j], A[piv][j]);
      std::swap(perm[c], perm[piv]);
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
      for (int i = 0; i < N8; ++i) b[i] = x[i];  // first pass: U x = ones
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
    for (int i = 0; i < N8; ++i) mx = std::fmax(mx, std::fmax(std::fabs(x[i].re), std::fabs(x[i].im)));
    if (mx == 0.0 || !std::isfinite(mx)) break;
    for (int i = 0; i < N8; ++i) {
      x[i].re = x[i].re / mx;
      x[i].im = x[i].im / mx;
    }
  }
  for (int i = 0; i < N8; ++i) vec[i] = x[i];
}

// opengv::math::cayley2rot (cayley.cpp:35-53), row-major 3x3.
void Cayley2Rot(const double c[3], double R[9]) {
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
}  // namespace

int Gp3pSolve(const double f[9], const double v[9], const double p[9], double solutions[8][12]) {
  static thread_local double S[GP3P_NUM_SLOTS];
  Gp3pEliminate(f, v, p, S);
  double M[N8][N8];
  std::memset(M, 0, sizeof(M));
  for (int r = 0; r < 6; ++r)
    for (int c = 0; c < 8; ++c) {
      const int slot = GP3P_ACTION[r * 8 + c];
      M[r][c] = (slot >= 0) ? -S[slot] : -0.0;
    }
  M[6][0] = 1.0;
  M[7][6] = 1.0;
  for (int r = 0; r < N8; ++r)
    for (int c = 0; c < N8; ++c)
      if (!std::isfinite(M[r][c])) return 0;
  double H[N8][N8];
  std::memcpy(H, M, sizeof(M));
  Hessenberg(H);
  double wr[N8], wi[N8];
  if (!HqrEigenvalues(H, wr, wi)) return 0;
  int num = 0;
  for (int c = 0; c < N8; ++c) {
    if (!(wi[c] < 0.0001)) continue;  // main.cpp:400 (no fabs: negative imaginary parts pass)
    Cx V[N8];
    InverseIteration(M, Cx{wr[c], wi[c]}, V);
    double cay[3], n[3];
    for (int i = 0; i < 3; ++i) {
      cay[2 - i] = CxDiv(V[i + 4], V[7]).re;
      n[2 - i] = CxDiv(V[i + 1], V[7]).re;
    }
    double Rt[9];
    Cayley2Rot(cay, Rt);
    double R[9];  // transposeInPlace (main.cpp:416)
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) R[i * 3 + j] = Rt[j * 3 + i];
    double center_cam[3] = {0, 0, 0}, center_world[3] = {0, 0, 0};
    for (int i = 0; i < 3; ++i) {
      double tmp[3], w[3];
      for (int a = 0; a < 3; ++a) w[a] = n[i] * f[i * 3 + a] + v[i * 3 + a];
      for (int a = 0; a < 3; ++a) tmp[a] = R[a * 3 + 0] * w[0] + R[a * 3 + 1] * w[1] + R[a * 3 + 2] * w[2];
      for (int a = 0; a < 3; ++a) {
        center_cam[a] = center_cam[a] + tmp[a];
        center_world[a] = center_world[a] + p[i * 3 + a];
      }
    }
    double* sol = solutions[num++];
    for (int a = 0; a < 3; ++a) {
      sol[a * 4 + 0] = R[a * 3 + 0];
      sol[a * 4 + 1] = R[a * 3 + 1];
      sol[a * 4 + 2] = R[a * 3 + 2];
      sol[a * 4 + 3] = center_world[a] / 3 - center_cam[a] / 3;
    }
  }
  return num;
}

// ------------------------------ RNG ----------------------------------------
RansacRng::RansacRng(uint32_t seed, int mapping_in) : idx(624), mapping(mapping_in) {
  mt[0] = seed;
  for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + i;
}
uint32_t RansacRng::NextU32() {
  if (idx >= 624) {
    for (int i = 0; i < 624; ++i) {
      const uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7fffffffu);
      mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    idx = 0;
  }
  uint32_t y = mt[idx++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}
// uniform_int_distribution<int>(0, INT_MAX) over a 32-bit engine.
//  mapping 1 (libstdc++ >= 11, bits/uniform_int_dist.h _S_nd): x >> 1, one draw.
//  mapping 0 (libstdc++ <= 10): redraw while x >= 2^31, return x.
int RansacRng::Next() {
  if (mapping == 1) return static_cast<int>(NextU32() >> 1);
  uint32_t x;
  do {
    x = NextU32();
  } while (x >= 0x80000000u);
  return static_cast<int>(x);
}

// ------------------------------ camera -------------------------------------
namespace {
struct Mat2 {
  double a, b, c, d;  // [a b; c d]
};
inline Mat2 Mul(const Mat2& l, const Mat2& r) {  // Eigen fixed-size product: sum over k ascending
  return Mat2{l.a * r.a + l.b * r.c, l.a * r.b + l.b * r.d, l.c * r.a + l.d * r.c, l.c * r.b + l.d * r.d};
}
inline Mat2 Transpose(const Mat2& m) { return Mat2{m.a, m.c, m.b, m.d}; }
inline Mat2 Inverse(const Mat2& m) {  // Eigen compute_inverse_size2_helper: invdet = 1 / det
  const double invdet = 1.0 / (m.a * m.d - m.b * m.c);
  return Mat2{m.d * invdet, -m.b * invdet, -m.c * invdet, m.a * invdet};
}

// aslam_cv2/aslam_cv_cameras/src/distortion-radtan.cc:14-62
void RadTanDistort(const double* coeffs, double* point, Mat2* jac) {
  double& x = point[0];
  double& y = point[1];
  const double k1 = coeffs[0], k2 = coeffs[1], p1 = coeffs[2], p2 = coeffs[3];
  double mx2_u = x * x;
  double my2_u = y * y;
  double mxy_u = x * y;
  double rho2_u = mx2_u + my2_u;
  double rad_dist_u = k1 * rho2_u + k2 * rho2_u * rho2_u;
  const double duf_du = 1.0 + rad_dist_u + 2.0 * k1 * mx2_u + 4.0 * k2 * rho2_u * mx2_u + 2.0 * p1 * y + 6.0 * p2 * x;
  const double duf_dv = 2.0 * k1 * mxy_u + 4.0 * k2 * rho2_u * mxy_u + 2.0 * p1 * x + 2.0 * p2 * y;
  const double dvf_du = duf_dv;
  const double dvf_dv = 1.0 + rad_dist_u + 2.0 * k1 * my2_u + 4.0 * k2 * rho2_u * my2_u + 2.0 * p2 * x + 6.0 * p1 * y;
  *jac = Mat2{duf_du, duf_dv, dvf_du, dvf_dv};
  x += x * rad_dist_u + 2.0 * p1 * mxy_u + p2 * (rho2_u + 2.0 * mx2_u);
  y += y * rad_dist_u + 2.0 * p2 * mxy_u + p1 * (rho2_u + 2.0 * my2_u);
}

// aslam_cv2/aslam_cv_cameras/src/distortion-equidistant.cc:14-103
void EquidistantDistort(const double* coeffs, double* point, Mat2* jac) {
  double& x = point[0];
  double& y = point[1];
  const double k1 = coeffs[0], k2 = coeffs[1], k3 = coeffs[2], k4 = coeffs[3];
  double x2 = x * x;
  double y2 = y * y;
  double r = std::sqrt(x2 + y2);
  if (r < 1e-10) {  // keypoint remains unchanged
    *jac = Mat2{0, 0, 0, 0};
    return;
  }
  double theta = std::atan(r);
  double theta2 = theta * theta;
  double theta4 = theta2 * theta2;
  double theta6 = theta2 * theta4;
  double theta8 = theta4 * theta4;
  double thetad = theta * (1 + k1 * theta2 + k2 * theta4 + k3 * theta6 + k4 * theta8);
  double theta3 = theta2 * theta;
  double theta5 = theta4 * theta;
  double theta7 = theta6 * theta;
  const double duf_du =
      theta * 1.0 / r * (k1 * theta2 + k2 * theta4 + k3 * theta6 + k4 * theta8 + 1.0) +
      x * theta * 1.0 / r *
          ((k2 * x * theta3 * 1.0 / r * 4.0) / (x2 + y2 + 1.0) + (k3 * x * theta5 * 1.0 / r * 6.0) / (x2 + y2 + 1.0) +
           (k4 * x * theta7 * 1.0 / r * 8.0) / (x2 + y2 + 1.0) + (k1 * x * theta * 1.0 / r * 2.0) / (x2 + y2 + 1.0)) +
      ((x2) * (k1 * theta2 + k2 * theta4 + k3 * theta6 + k4 * theta8 + 1.0)) / ((x2 + y2) * (x2 + y2 + 1.0)) -
      (x2)*theta * 1.0 / std::pow(x2 + y2, 3.0 / 2.0) * (k1 * theta2 + k2 * theta4 + k3 * theta6 + k4 * theta8 + 1.0);
  const double duf_dv =
      x * theta * 1.0 / r *
          ((k2 * y * theta3 * 1.0 / r * 4.0) / (x2 + y2 + 1.0) + (k3 * y * theta5 * 1.0 / r * 6.0) / (x2 + y2 + 1.0) +
           (k4 * y * theta7 * 1.0 / r * 8.0) / (x2 + y2 + 1.0) + (k1 * y * theta * 1.0 / r * 2.0) / (x2 + y2 + 1.0)) +
      (x * y * (k1 * theta2 + k2 * theta4 + k3 * theta6 + k4 * theta8 + 1.0)) / ((x2 + y2) * (x2 + y2 + 1.0)) -
      x * y * theta * 1.0 / std::pow(x2 + y2, 3.0 / 2.0) * (k1 * theta2 + k2 * theta4 + k3 * theta6 + k4 * theta8 + 1.0);
  const double dvf_du =
      y * theta * 1.0 / r *
          ((k2 * x * theta3 * 1.0 / r * 4.0) / (x2 + y2 + 1.0) + (k3 * x * theta5 * 1.0 / r * 6.0) / (x2 + y2 + 1.0) +
           (k4 * x * theta7 * 1.0 / r * 8.0) / (x2 + y2 + 1.0) + (k1 * x * theta * 1.0 / r * 2.0) / (x2 + y2 + 1.0)) +
      (x * y * (k1 * theta2 + k2 * theta4 + k3 * theta6 + k4 * theta8 + 1.0)) / ((x2 + y2) * (x2 + y2 + 1.0)) -
      x * y * theta * 1.0 / std::pow(x2 + y2, 3.0 / 2.0) * (k1 * theta2 + k2 * theta4 + k3 * theta6 + k4 * theta8 + 1.0);
  const double dvf_dv =
      theta * 1.0 / r * (k1 * theta2 + k2 * theta4 + k3 * theta6 + k4 * theta8 + 1.0) +
      y * theta * 1.0 / r *
          ((k2 * y * theta3 * 1.0 / r * 4.0) / (x2 + y2 + 1.0) + (k3 * y * theta5 * 1.0 / r * 6.0) / (x2 + y2 + 1.0) +
           (k4 * y * theta7 * 1.0 / r * 8.0) / (x2 + y2 + 1.0) + (k1 * y * theta * 1.0 / r * 2.0) / (x2 + y2 + 1.0)) +
      ((y2) * (k1 * theta2 + k2 * theta4 + k3 * theta6 + k4 * theta8 + 1.0)) / ((x2 + y2) * (x2 + y2 + 1.0)) -
      (y2)*theta * 1.0 / std::pow(x2 + y2, 3.0 / 2.0) * (k1 * theta2 + k2 * theta4 + k3 * theta6 + k4 * theta8 + 1.0);
  *jac = Mat2{duf_du, duf_dv, dvf_du, dvf_dv};
  double scaling = (r > 1e-8) ? thetad / r : 1.0;
  x *= scaling;
  y *= scaling;
}

// distortion-radtan.cc:96-118 / distortion-equidistant.cc:144-173 (--acv_inv_distortion_tolerance = 1e-8,
// aslam_cv_cameras/src/distortion.cc:8)
void UndistortIterative(bool equidistant, const double* coeffs, double* point) {
  const int n = 30;
  const double y[2] = {point[0], point[1]};
  double ybar[2] = {y[0], y[1]};
  if (equidistant && y[0] * y[0] + y[1] * y[1] < 1e-6) return;  // special case around the image centre
  for (int i = 0; i < n; ++i) {
    double y_tmp[2] = {ybar[0], ybar[1]};
    Mat2 F;
    if (equidistant) EquidistantDistort(coeffs, y_tmp, &F);
    else RadTanDistort(coeffs, y_tmp, &F);
    const double e[2] = {y[0] - y_tmp[0], y[1] - y_tmp[1]};
    const Mat2 Ft = Transpose(F);
    const Mat2 M = Mul(Inverse(Mul(Ft, F)), Ft);  // (F^T F)^-1 F^T
    ybar[0] += M.a * e[0] + M.b * e[1];
    ybar[1] += M.c * e[0] + M.d * e[1];
    if (e[0] * e[0] + e[1] * e[1] <= 1e-8) break;
  }
  point[0] = ybar[0];
  point[1] = ybar[1];
}
}  // namespace

void BackProject3(const Camera& c, const double kp_in[2], double b[3]) {
  double x = (kp_in[0] - c.cu) / c.fu;
  double y = (kp_in[1] - c.cv) / c.fv;
  if (c.distortion == 1) {  // FisheyeDistortion::undistort (distortion-fisheye.cc:119-143)
    const double w = c.dist[0];
    const double mul2tanwby2 = std::tan(w / 2.0) * 2.0;
    const double r_d = std::sqrt(x * x + y * y);
    if (!(mul2tanwby2 == 0 || r_d == 0)) {
      if (std::fabs(r_d * w) <= (89.0 * M_PI / 180.0)) {
        const double r_u = std::tan(r_d * w) / (r_d * mul2tanwby2);
        x *= r_u;
        y *= r_u;
      }
    }
  } else if (c.distortion == 2 || c.distortion == 3) {  // RadTanDistortion / EquidistantDistortion
    double pt[2] = {x, y};
    UndistortIterative(c.distortion == 3, c.dist, pt);
    x = pt[0];
    y = pt[1];
  }
  const double nrm = std::sqrt(x * x + y * y + 1.0);  // bearing.normalize()
  b[0] = x / nrm;
  b[1] = y / nrm;
  b[2] = 1.0 / nrm;
}

double RansacThreshold(const std::vector<Camera>& cams, double pixel_sigma) {
  double focal = 0;
  for (const Camera& c : cams) focal += (c.fu + c.fv);
  focal /= (2.0 * static_cast<double>(cams.size()));
  return 1.0 - std::cos(std::atan(pixel_sigma / focal));
}

namespace {
struct Problem {
  const double* bearings;
  const int* cam_idx;
  const double* points;
  int n;
  const std::vector<Camera>* cams;

  // AbsolutePoseSacProblem::computeModelCoefficients, GP3P branch.
  bool ComputeModel(const int idx[4], double model[12]) const {
    double f[9], v[9], p[9];
    for (int i = 0; i < 3; ++i) {
      const Camera& c = (*cams)[cam_idx[idx[i]]];
      const double* b = bearings + 3 * idx[i];
      for (int a = 0; a < 3; ++a) {
        f[i * 3 + a] = c.R_B_C[a * 3 + 0] * b[0] + c.R_B_C[a * 3 + 1] * b[1] + c.R_B_C[a * 3 + 2] * b[2];
        v[i * 3 + a] = c.t_B_C[a];
        p[i * 3 + a] = points[3 * idx[i] + a];
      }
    }
    double sols[8][12];
    const int ns = Gp3pSolve(f, v, p, sols);
    if (ns == 1) {
      std::memcpy(model, sols[0], sizeof(double) * 12);
      return true;
    }
    double min_score = 1000000.0;
    int min_index = -1;
    for (int s = 0; s < ns; ++s) {
      const double score = Distance(sols[s], idx[3]);
      if (score < min_score) {
        min_score = score;
        min_index = s;
      }
    }
    if (min_index == -1) return false;
    std::memcpy(model, sols[min_index], sizeof(double) * 12);
    return true;
  }

  // 1 - cos(angle) reprojection score (AbsolutePoseSacProblem.cpp:165-199).
  double Distance(const double T[12], int i) const {
    // inverse: Rinv = R^T, tinv = -(R^T t)
    double tinv[3];
    for (int a = 0; a < 3; ++a)
      tinv[a] = -(T[0 * 4 + a] * T[3] + T[1 * 4 + a] * T[7] + T[2 * 4 + a] * T[11]);
    const double* P = points + 3 * i;
    double body[3];
    for (int a = 0; a < 3; ++a)
      body[a] = T[0 * 4 + a] * P[0] + T[1 * 4 + a] * P[1] + T[2 * 4 + a] * P[2] + tinv[a] * 1.0;
    const Camera& c = (*cams)[cam_idx[i]];
    double d[3] = {body[0] - c.t_B_C[0], body[1] - c.t_B_C[1], body[2] - c.t_B_C[2]};
    double r[3];
    for (int a = 0; a < 3; ++a)
      r[a] = c.R_B_C[0 * 3 + a] * d[0] + c.R_B_C[1 * 3 + a] * d[1] + c.R_B_C[2 * 3 + a] * d[2];
    const double nrm = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    r[0] /= nrm;
    r[1] /= nrm;
    r[2] /= nrm;
    const double* b = bearings + 3 * i;
    return 1.0 - (r[0] * b[0] + r[1] * b[1] + r[2] * b[2]);
  }
};
}  // namespace

void AbsoluteMultiPoseRansac(const double* bearings, const int* cam_idx, const double* points,
                             int n, const std::vector<Camera>& cams, double threshold,
                             int max_iterations, RansacRng* rng, RansacResult* out) {
  Problem pb{bearings, cam_idx, points, n, &cams};
  out->success = false;
  out->inliers.clear();
  out->inlier_distances.clear();
  std::vector<int> shuffled(n);
  for (int i = 0; i < n; ++i) shuffled[i] = i;
  int iterations = 0;
  int best = -INT_MAX;
  double k = 1.0;
  unsigned skipped = 0;
  const unsigned max_skip = static_cast<unsigned>(max_iterations) * 10;
  bool have_model = false;
  double model[12], best_model[12];
  const double probability = 0.99;
  while (iterations < k && skipped < max_skip) {
    if (n < 4) {  // getSamples: cannot select 4 unique points
      iterations = INT_MAX;
      break;
    }
    int sel[4];
    for (unsigned i = 0; i < 4; ++i) {
      const size_t j = i + (static_cast<size_t>(rng->Next()) % (static_cast<size_t>(n) - i));
      std::swap(shuffled[i], shuffled[j]);
    }
    for (int i = 0; i < 4; ++i) sel[i] = shuffled[i];
    if (!pb.ComputeModel(sel, model)) {
      ++skipped;
      continue;
    }
    int count = 0;
    for (int i = 0; i < n; ++i)
      if (pb.Distance(model, i) < threshold) ++count;
    if (count > best) {
      best = count;
      have_model = true;
      std::memcpy(best_model, model, sizeof(model));
      std::memcpy(out->model_indices, sel, sizeof(sel));
      const double w = static_cast<double>(best) / static_cast<double>(n);
      double p_no_outliers = 1.0 - std::pow(w, 4.0);
      p_no_outliers = std::max(std::numeric_limits<double>::epsilon(), p_no_outliers);
      p_no_outliers = std::min(1.0 - std::numeric_limits<double>::epsilon(), p_no_outliers);
      k = std::log(1.0 - probability) / std::log(p_no_outliers);
    }
    ++iterations;
    if (iterations > max_iterations) break;
  }
  out->iterations = iterations;
  if (!have_model) return;
  out->success = true;
  std::memcpy(out->T, best_model, sizeof(best_model));
  for (int i = 0; i < n; ++i) {
    const double d = pb.Distance(best_model, i);
    if (d < threshold) {
      out->inliers.push_back(i);
      out->inlier_distances.push_back(d);
    }
  }
}

// LoopClosureHandler::handleLoopClosure gates (loop-closure-handler.cc:262-455)
// without the map mutation; getBestStructureMatchForEveryKeypoint (A22).
void HandleLoopClosure(const VerifyInput& in, const std::vector<Camera>& cams,
                       const HandlerSettings& hs, VerifyResult* out) {
  *out = VerifyResult();
  if (in.n < hs.min_inlier_count) return;
  std::vector<double> bearings(static_cast<size_t>(3) * in.n);
  for (int i = 0; i < in.n; ++i)
    BackProject3(cams[in.frame_index[i]], &in.keypoints[2 * i], &bearings[3 * i]);
  const double thr = RansacThreshold(cams, hs.ransac_pixel_sigma);
  RansacRng rng(hs.seed, hs.rng_mapping);
  AbsoluteMultiPoseRansac(bearings.data(), in.frame_index.data(), in.landmarks.data(), in.n, cams,
                          thr, hs.num_ransac_iters, &rng, &out->ransac);
  // Best inlier per (frame, keypoint): first wins on ties.
  std::map<std::pair<int, int>, std::pair<int, double>> best;
  for (size_t j = 0; j < out->ransac.inliers.size(); ++j) {
    const int mi = out->ransac.inliers[j];
    const double err = out->ransac.inlier_distances[j];
    const std::pair<int, int> key(in.frame_index[mi], in.keypoint_index[mi]);
    auto it = best.find(key);
    if (it == best.end())
      best.emplace(key, std::make_pair(mi, err));
    else if (it->second.second > err)
      it->second = std::make_pair(mi, err);
  }
  out->num_inliers = static_cast<int>(best.size());
  for (const auto& b : best) out->best_inlier_per_keypoint.push_back(b.second.first);
  if (out->num_inliers < hs.min_inlier_count) return;
  out->inlier_ratio = static_cast<double>(out->num_inliers) / static_cast<double>(in.n);
  if (out->inlier_ratio < hs.min_inlier_ratio) return;
  out->accepted = true;
}

// loop-closure-handler.cc:436-442. The rotation angle is Eigen's AngleAxis(quaternion):
// 2 * atan2(|q.vec|, |q.w|), with the quaternion of R_map^T R_ransac from Eigen's matrix ->
// quaternion conversion (Shepperd's method, Eigen/src/Geometry/Quaternion.h).
void DeltaPose(const double* A, const double* B, double* delta_position_m, double* delta_rotation_deg) {
  double R[9], d[3], p[3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0.0;
      for (int k = 0; k < 3; ++k) s += A[k * 4 + i] * B[k * 4 + j];  // R_map^T * R_ransac
      R[i * 3 + j] = s;
    }
  for (int k = 0; k < 3; ++k) d[k] = B[k * 4 + 3] - A[k * 4 + 3];
  for (int i = 0; i < 3; ++i) p[i] = A[0 * 4 + i] * d[0] + A[1 * 4 + i] * d[1] + A[2 * 4 + i] * d[2];
  *delta_position_m = std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
  double w, x, y, z;
  const double t = R[0] + R[4] + R[8];
  if (t > 0.0) {
    double r = std::sqrt(t + 1.0);
    w = 0.5 * r;
    r = 0.5 / r;
    x = (R[7] - R[5]) * r;
    y = (R[2] - R[6]) * r;
    z = (R[3] - R[1]) * r;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[i * 3 + i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    double r = std::sqrt(R[i * 3 + i] - R[j * 3 + j] - R[k * 3 + k] + 1.0);
    double v[3];
    v[i] = 0.5 * r;
    r = 0.5 / r;
    w = (R[k * 3 + j] - R[j * 3 + k]) * r;
    v[j] = (R[j * 3 + i] + R[i * 3 + j]) * r;
    v[k] = (R[k * 3 + i] + R[i * 3 + k]) * r;
    x = v[0];
    y = v[1];
    z = v[2];
  }
  const double n = std::sqrt(x * x + y * y + z * z);
  const double angle = n != 0.0 ? 2.0 * std::atan2(n, std::fabs(w)) : 0.0;
  *delta_rotation_deg = angle * (180.0 / M_PI);
}

bool DeltaPoseGate(const double* T_G_I_map, const double* T_G_I_ransac, double max_delta_position_m,
                   double max_delta_rotation_deg) {
  if (!(max_delta_position_m >= 0.0 || max_delta_rotation_deg >= 0.0)) return true;
  double dp, dr;
  DeltaPose(T_G_I_map, T_G_I_ransac, &dp, &dr);
  const bool is_distance_ok = max_delta_position_m < 0.0 || dp <= max_delta_position_m;
  const bool is_rotation_ok = max_delta_rotation_deg < 0.0 || dr <= max_delta_rotation_deg;
  return is_distance_ok && is_rotation_ok;
}

}  // namespace lc_oracle
