// See summary_map.h. proto2 wire format: key = (field_number << 3) | wire_type as a varint;
// wire types 0 varint, 1 fixed64, 2 length-delimited, 3/4 group start/end, 5 fixed32.
#include "summary_map.h"

#include <cstring>

namespace mlc {
namespace {

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
  bool Skip(uint32_t field, int wire) {
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
        while (true) {
          uint64_t key;
          if (!Varint(&key)) return false;
          const int kw = static_cast<int>(key & 7);
          const uint32_t kf = static_cast<uint32_t>(key >> 3);
          if (kf == 0) return false;
          if (kw == 4) return kf == field;
          if (!Skip(kf, kw)) return false;
        }
      case 5: return Fixed32(&w);
      default: return false;  // 4 (stray end-group), 6, 7
    }
  }
};

bool Key(Reader* r, uint32_t* field, int* wire) {
  uint64_t key;
  if (!r->Varint(&key)) return false;
  *wire = static_cast<int>(key & 7);
  *field = static_cast<uint32_t>(key >> 3);
  return *field != 0 && (key >> 3) <= 0x1FFFFFFFull;
}

// repeated float: one fixed32 per key (wire 5) or a packed run (wire 2).
bool RepeatedFloat(Reader* r, int wire, std::vector<float>* out, bool* handled) {
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
bool RepeatedU32(Reader* r, int wire, std::vector<uint32_t>* out, bool* handled) {
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

bool ParseMatrixXf(Reader r, SummaryMap* m) {
  uint32_t field;
  int wire;
  while (r.p != r.end) {
    if (!Key(&r, &field, &wire)) return false;
    bool handled = false;
    uint64_t v;
    if ((field == 1 || field == 2) && wire == 0) {
      if (!r.Varint(&v)) return false;
      (field == 1 ? m->descriptor_rows : m->descriptor_cols) = static_cast<uint32_t>(v);
      handled = true;
    } else if (field == 3) {
      if (!RepeatedFloat(&r, wire, &m->descriptors, &handled)) return false;
    }
    if (!handled && !r.Skip(field, wire)) return false;
  }
  return true;
}

bool ParseUncompressed(Reader r, SummaryMap* m) {
  uint32_t field;
  int wire;
  while (r.p != r.end) {
    if (!Key(&r, &field, &wire)) return false;
    bool handled = false;
    if (field == 1 && wire == 2) {
      Reader sub;
      if (!r.Sub(&sub) || !ParseMatrixXf(sub, m)) return false;
      handled = true;
    } else if (field == 2) {
      if (!RepeatedFloat(&r, wire, &m->G_observer_position, &handled)) return false;
    } else if (field == 3) {
      if (!RepeatedU32(&r, wire, &m->observer_indices, &handled)) return false;
    } else if (field == 4) {
      if (!RepeatedU32(&r, wire, &m->observation_to_landmark_index, &handled)) return false;
    }
    if (!handled && !r.Skip(field, wire)) return false;
  }
  return true;
}

void PutVarint(uint64_t v, std::vector<uint8_t>* out) {
  while (v >= 0x80) {
    out->push_back(static_cast<uint8_t>(v) | 0x80);
    v >>= 7;
  }
  out->push_back(static_cast<uint8_t>(v));
}
size_t VarintSize(uint64_t v) {
  size_t n = 1;
  while (v >= 0x80) {
    v >>= 7;
    ++n;
  }
  return n;
}
void PutFloats(uint32_t field, const std::vector<float>& v, std::vector<uint8_t>* out) {
  const uint8_t key = static_cast<uint8_t>((field << 3) | 5);
  const size_t at = out->size();
  out->resize(at + 5 * v.size());
  uint8_t* p = out->data() + at;
  for (float f : v) {
    *p++ = key;
    std::memcpy(p, &f, 4);
    p += 4;
  }
}
void PutU32s(uint32_t field, const std::vector<uint32_t>& v, std::vector<uint8_t>* out) {
  for (uint32_t x : v) {
    out->push_back(static_cast<uint8_t>(field << 3));
    PutVarint(x, out);
  }
}
size_t U32sSize(const std::vector<uint32_t>& v) {
  size_t n = 0;
  for (uint32_t x : v) n += 1 + VarintSize(x);
  return n;
}

}  // namespace

bool SummaryMap::Parse(const void* blob, size_t size, std::string* err) {
  *this = SummaryMap();
  if (size > 0 && !blob) {
    *err = "summary map: null buffer";
    return false;
  }
  Reader r{static_cast<const uint8_t*>(blob), static_cast<const uint8_t*>(blob) + size};
  uint32_t field;
  int wire;
  while (r.p != r.end) {
    bool ok = Key(&r, &field, &wire);
    bool handled = false;
    if (ok && field == 1) {
      ok = RepeatedFloat(&r, wire, &G_landmark_position, &handled);
    } else if (ok && field == 2 && wire == 2) {
      Reader sub;
      ok = r.Sub(&sub) && ParseUncompressed(sub, this);
      has_uncompressed_map = true;
      handled = true;
    }
    if (ok && !handled) ok = r.Skip(field, wire);
    if (!ok) {
      *err = "summary map: malformed protobuf wire data (parseProtoFromFile would fail)";
      return false;
    }
  }
  if (G_landmark_position.size() % 3 != 0) {
    *err = "summary map: G_landmark_position is not 3 x L (CHECK_EQ(0, proto.size() % Rows))";
    return false;
  }
  if (!has_uncompressed_map) {
    *err = "Unsupported localization summary map format.";
    return false;
  }
  if (G_observer_position.size() % 3 != 0) {
    *err = "summary map: G_observer_position is not 3 x O (CHECK_EQ(0, proto.size() % Rows))";
    return false;
  }
  // CHECK_EQ(static_cast<int>(proto.rows() * proto.cols()), proto.data_size())
  if (static_cast<int>(descriptor_rows * descriptor_cols) != static_cast<int>(descriptors.size()) ||
      static_cast<uint64_t>(descriptor_rows) * descriptor_cols != descriptors.size()) {
    *err = "summary map: descriptors rows * cols != data size";
    return false;
  }
  return true;
}

size_t SummaryMap::SerializedSize() const {
  const size_t matrix_size = 1 + VarintSize(descriptor_rows) + 1 + VarintSize(descriptor_cols) +
                             5 * (static_cast<size_t>(descriptor_rows) * descriptor_cols);
  const size_t sub_size = 1 + VarintSize(matrix_size) + matrix_size + 5 * G_observer_position.size() +
                          U32sSize(observer_indices) + U32sSize(observation_to_landmark_index);
  return 5 * G_landmark_position.size() + 1 + VarintSize(sub_size) + sub_size;
}

void SummaryMap::Serialize(std::vector<uint8_t>* out) const {
  out->clear();
  PutFloats(1, G_landmark_position, out);
  const size_t matrix_size = 1 + VarintSize(descriptor_rows) + 1 + VarintSize(descriptor_cols) +
                             5 * descriptors.size();
  const size_t sub_size = 1 + VarintSize(matrix_size) + matrix_size + 5 * G_observer_position.size() +
                          U32sSize(observer_indices) + U32sSize(observation_to_landmark_index);
  out->reserve(out->size() + sub_size + 12);
  out->push_back((2u << 3) | 2);
  PutVarint(sub_size, out);
  out->push_back((1u << 3) | 2);
  PutVarint(matrix_size, out);
  out->push_back(1u << 3);
  PutVarint(descriptor_rows, out);
  out->push_back(2u << 3);
  PutVarint(descriptor_cols, out);
  PutFloats(3, descriptors, out);
  PutFloats(2, G_observer_position, out);
  PutU32s(3, observer_indices, out);
  PutU32s(4, observation_to_landmark_index, out);
}

bool GroupSummaryMapByObserver(const SummaryMap& map, SummaryMapImages* out, std::string* err) {
  const int64_t observers = map.num_observers(), n = map.num_observations();
  const int64_t landmarks = map.num_landmarks();
  const uint32_t dim = map.descriptor_rows;
  if (observers == 0) {
    *err = "No observers in the summary map found. Is it initialized?";
    return false;
  }
  out->num_descriptors.assign(static_cast<size_t>(observers), 0);
  for (int64_t i = 0; i < n; ++i) {
    const uint32_t o = map.observer_indices[i];
    if (o >= static_cast<uint64_t>(observers)) {
      *err = "summary map: observer index out of range (CHECK_LT(observer_index, observers))";
      return false;
    }
    if (i >= static_cast<int64_t>(map.descriptor_cols)) {
      *err = "summary map: fewer descriptors than observations (CHECK_LT(observation_index, cols))";
      return false;
    }
    if (i >= static_cast<int64_t>(map.observation_to_landmark_index.size())) {
      *err = "summary map: observation without landmark index";
      return false;
    }
    if (map.observation_to_landmark_index[i] >= static_cast<uint64_t>(landmarks)) {
      *err = "summary map: landmark index out of range (CHECK_LT(landmark_index, landmarks))";
      return false;
    }
    if (out->num_descriptors[o] == INT32_MAX) {
      *err = "summary map: too many observations of one observer";
      return false;
    }
    ++out->num_descriptors[o];
  }
  std::vector<int64_t> at(static_cast<size_t>(observers) + 1, 0);
  for (int64_t o = 0; o < observers; ++o) at[o + 1] = at[o] + out->num_descriptors[o];
  out->proj.assign(static_cast<size_t>(n) * dim, 0.f);
  out->landmark_index.assign(static_cast<size_t>(n), -1);
  for (int64_t i = 0; i < n; ++i) {  // stable: observations keep their order inside an observer
    const int64_t slot = at[map.observer_indices[i]]++;
    if (dim) std::memcpy(&out->proj[slot * dim], &map.descriptors[static_cast<size_t>(i) * dim], 4 * dim);
    out->landmark_index[slot] = map.observation_to_landmark_index[i];
  }
  return true;
}

}  // namespace mlc
