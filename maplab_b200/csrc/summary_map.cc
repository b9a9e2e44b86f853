// See summary_map.h. proto2 wire format: key = (field_number << 3) | wire_type as a varint;
// wire types 0 varint, 1 fixed64, 2 length-delimited, 3/4 group start/end, 5 fixed32.
#include "summary_map.h"

#include <cstring>

#include "proto_wire.h"

namespace mlc {
namespace {
using namespace wire;

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
