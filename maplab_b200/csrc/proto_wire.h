// proto2 wire-format primitives shared by the file readers of this library (summary_map.cc,
// vi_map_reader.cc): key = (field_number << 3) | wire_type as a varint; wire types 0 varint,
// 1 fixed64, 2 length-delimited, 3 / 4 group start / end, 5 fixed32. Unknown fields (and known ones
// with an unexpected wire type) are skipped like libprotobuf does; repeated scalars are accepted
// packed and unpacked.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace mlc {
namespace wire {

struct Reader {
  const uint8_t* p;
  const uint8_t* end;
  bool Varint(uint64_t* v) {
    uint64_t r = 0;
    for (int shift = 0; shift < 70; shift += 7) {
      if (p == end) return false;
      const uint8_t b = *p++;
      if (shift < 64) r |= static_cast<uint64_t>(b & 0x7F) << shift;
      if (!(b & 0x80)) {
        *v = r;
        return true;
      }
    }
    return false;  // more than 10 bytes
  }
  bool Fixed32(uint32_t* v) {
    if (end - p < 4) return false;
    std::memcpy(v, p, 4);  // little endian on the wire and on the host
    p += 4;
    return true;
  }
  bool Sub(Reader* sub) {
    uint64_t len;
    if (!Varint(&len) || len > static_cast<uint64_t>(end - p)) return false;
    sub->p = p;
    sub->end = p + len;
    p += len;
    return true;
  }
  // Unknown field (or a known one with an unexpected wire type): skipped like libprotobuf does.
  // Groups nest; libprotobuf refuses more than 100 levels, so does this (a crafted file must not be able to
  // overflow the stack).
  bool Skip(uint32_t field, int wire, int depth = 0) {
    uint64_t v;
    uint32_t w;
    Reader sub;
    switch (wire) {
      case 0: return Varint(&v);
      case 1:
        if (end - p < 8) return false;
        p += 8;
        return true;
      case 2: return Sub(&sub);
      case 3:  // group: skip until the matching end-group key
        if (depth >= 100) return false;
        while (true) {
          uint64_t key;
          if (!Varint(&key)) return false;
          const int kw = static_cast<int>(key & 7);
          const uint32_t kf = static_cast<uint32_t>(key >> 3);
          if (kf == 0) return false;
          if (kw == 4) return kf == field;
          if (!Skip(kf, kw, depth + 1)) return false;
        }
      case 5: return Fixed32(&w);
      default: return false;  // 4 (stray end-group), 6, 7
    }
  }
};

inline bool Key(Reader* r, uint32_t* field, int* wire) {
  uint64_t key;
  if (!r->Varint(&key)) return false;
  *wire = static_cast<int>(key & 7);
  *field = static_cast<uint32_t>(key >> 3);
  return *field != 0 && (key >> 3) <= 0x1FFFFFFFull;
}

// repeated float: one fixed32 per key (wire 5) or a packed run (wire 2).
inline bool RepeatedFloat(Reader* r, int wire, std::vector<float>* out, bool* handled) {
  *handled = true;
  if (wire == 5) {
    uint32_t w;
    if (!r->Fixed32(&w)) return false;
    float f;
    std::memcpy(&f, &w, 4);
    out->push_back(f);
    return true;
  }
  if (wire == 2) {
    Reader sub;
    if (!r->Sub(&sub) || (sub.end - sub.p) % 4 != 0) return false;
    const size_t n = static_cast<size_t>(sub.end - sub.p) / 4, at = out->size();
    out->resize(at + n);
    if (n) std::memcpy(out->data() + at, sub.p, 4 * n);
    return true;
  }
  *handled = false;
  return true;
}
// repeated uint32: one varint per key (wire 0) or a packed run of varints (wire 2).
inline bool RepeatedU32(Reader* r, int wire, std::vector<uint32_t>* out, bool* handled) {
  *handled = true;
  uint64_t v;
  if (wire == 0) {
    if (!r->Varint(&v)) return false;
    out->push_back(static_cast<uint32_t>(v));
    return true;
  }
  if (wire == 2) {
    Reader sub;
    if (!r->Sub(&sub)) return false;
    while (sub.p != sub.end) {
      if (!sub.Varint(&v)) return false;
      out->push_back(static_cast<uint32_t>(v));
    }
    return true;
  }
  *handled = false;
  return true;
}

// repeated double: one fixed64 per key (wire 1) or a packed run (wire 2).
inline bool RepeatedDouble(Reader* r, int wire, std::vector<double>* out, bool* handled) {
  *handled = true;
  if (wire == 1) {
    if (r->end - r->p < 8) return false;
    double d;
    std::memcpy(&d, r->p, 8);
    r->p += 8;
    out->push_back(d);
    return true;
  }
  if (wire == 2) {
    Reader sub;
    if (!r->Sub(&sub) || (sub.end - sub.p) % 8 != 0) return false;
    const size_t n = static_cast<size_t>(sub.end - sub.p) / 8, at = out->size();
    out->resize(at + n);
    if (n) std::memcpy(out->data() + at, sub.p, 8 * n);
    return true;
  }
  *handled = false;
  return true;
}
// repeated uint64: one varint per key (wire 0) or a packed run of varints (wire 2).
inline bool RepeatedU64(Reader* r, int wire, std::vector<uint64_t>* out, bool* handled) {
  *handled = true;
  uint64_t v;
  if (wire == 0) {
    if (!r->Varint(&v)) return false;
    out->push_back(v);
    return true;
  }
  if (wire == 2) {
    Reader sub;
    if (!r->Sub(&sub)) return false;
    while (sub.p != sub.end) {
      if (!sub.Varint(&v)) return false;
      out->push_back(v);
    }
    return true;
  }
  *handled = false;
  return true;
}

}  // namespace wire
}  // namespace mlc
