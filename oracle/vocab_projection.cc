// ORACLE (test infrastructure). Vocabulary file format (T2) + projection (A1-A3).
#include <cmath>
#include <cstring>

#include "lc_oracle.h"

namespace lc_oracle {
namespace {
struct Reader {
  const uint8_t* p;
  size_t left;
  bool ReadInt(int* v) {
    if (left < 4) return false;
    std::memcpy(v, p, 4);
    p += 4;
    left -= 4;
    return true;
  }
  // common::Deserialize(Eigen::Matrix) — binary-serialization.h:128-161:
  // int rows; int cols; rows*cols raw scalars, column-major.
  bool ReadMatrix(Matrix* m) {
    int r, c;
    if (!ReadInt(&r) || !ReadInt(&c)) return false;
    if (r < 0 || c < 0) return false;
    const size_t bytes = static_cast<size_t>(r) * c * sizeof(float);
    if (left < bytes) return false;
    m->rows = r;
    m->cols = c;
    m->data.resize(static_cast<size_t>(r) * c);
    std::memcpy(m->data.data(), p, bytes);
    p += bytes;
    left -= bytes;
    return true;
  }
};
void PutInt(std::vector<uint8_t>* o, int v) {
  const uint8_t* b = reinterpret_cast<const uint8_t*>(&v);
  o->insert(o->end(), b, b + 4);
}
void PutMatrix(std::vector<uint8_t>* o, const Matrix& m) {
  PutInt(o, m.rows);
  PutInt(o, m.cols);
  const uint8_t* b = reinterpret_cast<const uint8_t*>(m.data.data());
  o->insert(o->end(), b, b + m.data.size() * sizeof(float));
}
}  // namespace

// InvertedMultiIndexVocabulary::Load (inverted-multi-index-interface.h:37-47) and
// InvertedMultiIndexProductVocabulary::Load (:72-88).
bool ParseVocabulary(const uint8_t* blob, size_t size, bool want_pq, Vocabulary* out,
                     std::string* err) {
  Reader r{blob, size};
  if (!r.ReadInt(&out->version) || !r.ReadInt(&out->target_dim)) {
    if (err) *err = "truncated header";
    return false;
  }
  if (!r.ReadMatrix(&out->projection) || !r.ReadMatrix(&out->words1) ||
      !r.ReadMatrix(&out->words2)) {
    if (err) *err = "truncated matrices";
    return false;
  }
  out->has_pq = false;
  if (want_pq) {
    int v200 = 0;
    if (!r.ReadInt(&v200) || v200 != 200) {
      if (err) *err = "This vocabulary file was saved with a different version.";
      return false;
    }
    if (!r.ReadInt(&out->pq_num_components) || !r.ReadInt(&out->pq_num_centers) ||
        !r.ReadInt(&out->pq_dim_per_comp) || !r.ReadMatrix(&out->pq_centers1) ||
        !r.ReadMatrix(&out->pq_centers2)) {
      if (err) *err = "truncated PQ block";
      return false;
    }
    out->has_pq = true;
  }
  return true;
}

std::vector<uint8_t> SerializeVocabulary(const Vocabulary& v) {
  std::vector<uint8_t> o;
  PutInt(&o, v.version);
  PutInt(&o, v.target_dim);
  PutMatrix(&o, v.projection);
  PutMatrix(&o, v.words1);
  PutMatrix(&o, v.words2);
  if (v.has_pq) {
    PutInt(&o, 200);
    PutInt(&o, v.pq_num_components);
    PutInt(&o, v.pq_num_centers);
    PutInt(&o, v.pq_dim_per_comp);
    PutMatrix(&o, v.pq_centers1);
    PutMatrix(&o, v.pq_centers2);
  }
  return o;
}

// Row d of P is mapped to integers p_int = rint(P * 2^shift_d) with
// |p_int| <= 2^26, shift_d = 26 - ceil(log2(max_k |P[d][k]|)).
FixedPointProjection QuantizeProjection(const Matrix& P, int target_dim) {
  FixedPointProjection fp;
  fp.target_dim = target_dim;
  fp.kp = P.cols;
  fp.p_int.assign(static_cast<size_t>(target_dim) * P.cols, 0);
  fp.shift.assign(target_dim, 0);
  for (int d = 0; d < target_dim; ++d) {
    float mx = 0.f;
    for (int k = 0; k < P.cols; ++k) mx = std::fmax(mx, std::fabs(P.at(d, k)));
    int e = 0;
    if (mx > 0.f) {
      int ex;
      const float m = std::frexp(mx, &ex);  // mx = m * 2^ex, m in [0.5, 1)
      e = (m == 0.5f) ? ex - 1 : ex;        // smallest e with mx <= 2^e
    }
    const int shift = 26 - e;
    fp.shift[d] = shift;
    for (int k = 0; k < P.cols; ++k) {
      const double scaled = std::ldexp(static_cast<double>(P.at(d, k)), shift);
      fp.p_int[static_cast<size_t>(d) * P.cols + k] =
          static_cast<int32_t>(std::nearbyint(scaled));  // round-half-even
    }
  }
  return fp;
}

void SplitDigits(int32_t v, int32_t digits[3]) {
  // Balanced base-512: each digit in [-256, 255] except the top one (<= 256).
  int32_t rest = v;
  for (int j = 0; j < 2; ++j) {
    int32_t d = ((rest % 512) + 512) % 512;  // 0..511
    if (d >= 256) d -= 512;
    digits[j] = d;
    rest = (rest - d) / 512;
  }
  digits[2] = rest;
}

// DescriptorToEigenMatrix: out[8*byte+bit] = (byte >> bit) & 1 (LSB first),
// descriptor-projection.h:92-115. ProjectDescriptorBlock uses only the first
// P.cols() bits (descriptor-projection.cc:35-49).
void ProjectDescriptorBlock(const uint8_t* raw, int bytes_per_desc, int n,
                            const FixedPointProjection& fp, float* out) {
  for (int i = 0; i < n; ++i) {
    const uint8_t* desc = raw + static_cast<size_t>(i) * bytes_per_desc;
    for (int d = 0; d < fp.target_dim; ++d) {
      const int32_t* row = fp.p_int.data() + static_cast<size_t>(d) * fp.kp;
      int64_t acc = 0;
      for (int k = 0; k < fp.kp; ++k) {
        if ((desc[k >> 3] >> (k & 7)) & 1) acc += row[k];
      }
      // One rounding: int64 -> fp32 (RNE), then exact power-of-two scaling.
      const float y = static_cast<float>(acc);
      out[static_cast<size_t>(i) * fp.target_dim + d] = std::ldexp(y, -fp.shift[d]);
    }
  }
}

void ProjectDescriptorBlockFloat(const uint8_t* raw, int bytes_per_desc, int n, const Matrix& P,
                                 int target_dim, float* out) {
  for (int i = 0; i < n; ++i) {
    const uint8_t* desc = raw + static_cast<size_t>(i) * bytes_per_desc;
    for (int d = 0; d < target_dim; ++d) {
      float acc = 0.f;
      for (int k = 0; k < P.cols; ++k) {
        const float x = ((desc[k >> 3] >> (k & 7)) & 1) ? 1.f : 0.f;
        acc += P.at(d, k) * x;
      }
      out[static_cast<size_t>(i) * target_dim + d] = acc;
    }
  }
}

}  // namespace lc_oracle
