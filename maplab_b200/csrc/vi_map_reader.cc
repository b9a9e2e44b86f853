// See vi_map_reader.h.
#include "vi_map_reader.h"

#include <cstring>

#include "proto_wire.h"

namespace mlc {
namespace {
using namespace wire;

// aslam.proto.Id -> two 64-bit words (missing words read as 0 = the invalid id)
bool ParseId(Reader r, uint64_t out[2]) {
  std::vector<uint64_t> words;
  uint32_t field;
  int wt;
  while (r.p != r.end) {
    if (!Key(&r, &field, &wt)) return false;
    bool handled = false;
    if (field == 1 && !RepeatedU64(&r, wt, &words, &handled)) return false;
    if (!handled && !r.Skip(field, wt)) return false;
  }
  out[0] = words.size() > 0 ? words[0] : 0;
  out[1] = words.size() > 1 ? words[1] : 0;
  return words.size() <= 2;
}

bool ParseFrame(Reader r, ViMapVertices* m, std::string* err) {
  std::vector<double> measurements;
  std::vector<uint64_t> landmark_ids;
  const uint8_t* desc = nullptr;
  size_t desc_size = 0;
  int64_t timestamp = 0;
  // deserializeVisualFrame (aslam-serialization/src/visual-frame-serialization.cc:88-168): a frame whose id
  // is invalid "has been un-set"; a frame is valid unless is_valid is PRESENT and false
  uint8_t is_valid = 1;
  uint64_t frame_id[2] = {0, 0};
  uint32_t field;
  int wt;
  while (r.p != r.end) {
    if (!Key(&r, &field, &wt)) return false;
    bool handled = false;
    uint64_t v;
    Reader sub;
    if (field == 1 && wt == 2) {
      if (!r.Sub(&sub) || !ParseId(sub, frame_id)) return false;
      handled = true;
    } else if (field == 2 && wt == 0) {
      if (!r.Varint(&v)) return false;
      timestamp = static_cast<int64_t>(v);
      handled = true;
    } else if (field == 3) {
      if (!RepeatedDouble(&r, wt, &measurements, &handled)) return false;
    } else if (field == 5 && wt == 2) {
      if (!r.Sub(&sub)) return false;
      desc = sub.p;  // optional bytes: the last occurrence wins
      desc_size = static_cast<size_t>(sub.end - sub.p);
      handled = true;
    } else if (field == 7 && wt == 2) {
      uint64_t id[2];
      if (!r.Sub(&sub) || !ParseId(sub, id)) return false;
      landmark_ids.push_back(id[0]);
      landmark_ids.push_back(id[1]);
      handled = true;
    } else if (field == 9 && wt == 0) {
      if (!r.Varint(&v)) return false;
      is_valid = v != 0;
      handled = true;
    }
    if (!handled && !r.Skip(field, wt)) return false;
  }
  if (measurements.size() % 2 != 0) {
    *err = "vi_map: odd number of keypoint measurement coordinates";
    return false;
  }
  const size_t n = measurements.size() / 2;
  if (landmark_ids.size() != 2 * n) {
    *err = "vi_map: keypoints and observed landmark ids differ in number";
    return false;
  }
  if (n > 0 || desc_size > 0) {
    // aslam's descriptor matrix: 24-byte header with int32 rows at byte 8 and int32 cols at byte 12, then the
    // column-major uchar data (one descriptor per column)
    int32_t rows = 0, cols = 0;
    uint64_t blocks = 0;  // the DESCRIPTORS channel is a vector of matrices: leading size_t = their number
    if (desc_size >= 24) {
      std::memcpy(&blocks, desc, 8);
      std::memcpy(&rows, desc + 8, 4);
      std::memcpy(&cols, desc + 12, 4);
      if (blocks > 1) {
        *err = "vi_map: frames with several descriptor types (descriptor blocks) are not supported";
        return false;
      }
    }
    if (desc_size < 24 || blocks != 1 || rows <= 0 || cols < 0 || static_cast<size_t>(rows) * cols + 24 != desc_size ||
        static_cast<size_t>(cols) != n) {
      *err = "vi_map: descriptor matrix does not match the keypoints";
      return false;
    }
    if (m->descriptor_bytes == 0) m->descriptor_bytes = rows;
    if (m->descriptor_bytes != rows) {
      *err = "vi_map: descriptors of different sizes in one file";
      return false;
    }
    m->keypoint_descriptor.insert(m->keypoint_descriptor.end(), desc + 24, desc + desc_size);
  }
  m->keypoint_measurement.insert(m->keypoint_measurement.end(), measurements.begin(), measurements.end());
  m->keypoint_landmark_id.insert(m->keypoint_landmark_id.end(), landmark_ids.begin(), landmark_ids.end());
  m->frame_timestamp_ns.push_back(timestamp);
  m->frame_num_keypoints.push_back(static_cast<int32_t>(n));
  m->frame_is_valid.push_back((frame_id[0] != 0 || frame_id[1] != 0) ? is_valid : 0);
  return true;
}

bool ParseNFrame(Reader r, ViMapVertices* m, int32_t* num_frames, std::string* err) {
  uint32_t field;
  int wt;
  while (r.p != r.end) {
    if (!Key(&r, &field, &wt)) return false;
    if (field == 2 && wt == 2) {
      Reader sub;
      if (!r.Sub(&sub) || !ParseFrame(sub, m, err)) return false;
      ++*num_frames;
    } else if (!r.Skip(field, wt)) {
      return false;
    }
  }
  return true;
}

bool ParseLandmark(Reader r, ViMapVertices* m, std::string* err) {
  uint64_t id[2] = {0, 0};
  std::vector<double> position;
  int32_t quality = 0;  // [default = kUnknown]
  uint32_t field;
  int wt;
  while (r.p != r.end) {
    if (!Key(&r, &field, &wt)) return false;
    bool handled = false;
    Reader sub;
    uint64_t v;
    if (field == 1 && wt == 2) {
      if (!r.Sub(&sub) || !ParseId(sub, id)) return false;
      handled = true;
    } else if (field == 2) {
      if (!RepeatedDouble(&r, wt, &position, &handled)) return false;
    } else if (field == 7 && wt == 0) {
      if (!r.Varint(&v)) return false;
      quality = static_cast<int32_t>(v);
      handled = true;
    }
    if (!handled && !r.Skip(field, wt)) return false;
  }
  if (position.size() != 3) {
    *err = "vi_map: landmark position is not 3-dimensional";
    return false;
  }
  m->landmark_id.push_back(id[0]);
  m->landmark_id.push_back(id[1]);
  m->landmark_p_B.insert(m->landmark_p_B.end(), position.begin(), position.end());
  m->landmark_quality.push_back(quality);
  return true;
}

bool ParseLandmarkStore(Reader r, ViMapVertices* m, int32_t* num_landmarks, std::string* err) {
  uint32_t field;
  int wt;
  while (r.p != r.end) {
    if (!Key(&r, &field, &wt)) return false;
    if (field == 1 && wt == 2) {
      Reader sub;
      if (!r.Sub(&sub) || !ParseLandmark(sub, m, err)) return false;
      ++*num_landmarks;
    } else if (!r.Skip(field, wt)) {
      return false;
    }
  }
  return true;
}

bool ParseVertex(Reader r, ViMapVertices* m, std::string* err) {
  std::vector<double> T_M_I;
  uint64_t mission[2] = {0, 0};
  int32_t num_frames = 0, num_landmarks = 0;
  uint32_t field;
  int wt;
  while (r.p != r.end) {
    if (!Key(&r, &field, &wt)) return false;
    bool handled = false;
    Reader sub;
    if (field == 3) {
      if (!RepeatedDouble(&r, wt, &T_M_I, &handled)) return false;
    } else if (field == 7 && wt == 2) {
      if (!r.Sub(&sub) || !ParseNFrame(sub, m, &num_frames, err)) return false;
      handled = true;
    } else if (field == 8 && wt == 2) {
      if (!r.Sub(&sub) || !ParseLandmarkStore(sub, m, &num_landmarks, err)) return false;
      handled = true;
    } else if (field == 14 && wt == 2) {
      if (!r.Sub(&sub) || !ParseId(sub, mission)) return false;
      handled = true;
    }
    if (!handled && !r.Skip(field, wt)) return false;
  }
  if (T_M_I.size() != 7) {
    *err = "vi_map: T_M_I is not a 7-vector (quaternion + position)";
    return false;
  }
  m->T_M_I.insert(m->T_M_I.end(), T_M_I.begin(), T_M_I.end());
  m->mission_id.push_back(mission[0]);
  m->mission_id.push_back(mission[1]);
  m->vertex_num_frames.push_back(num_frames);
  m->vertex_num_landmarks.push_back(num_landmarks);
  return true;
}

// first Id field (number `id_field`) and first repeated-double field (number `vec_field`) of a sub-message
bool ParseIdAndVector(Reader r, uint32_t id_field, uint32_t vec_field, uint64_t id[2], std::vector<double>* vec) {
  id[0] = id[1] = 0;
  uint32_t field;
  int wt;
  while (r.p != r.end) {
    if (!Key(&r, &field, &wt)) return false;
    bool handled = false;
    Reader sub;
    if (id_field != 0 && field == id_field && wt == 2) {
      if (!r.Sub(&sub) || !ParseId(sub, id)) return false;
      handled = true;
    } else if (vec_field != 0 && field == vec_field) {
      if (!RepeatedDouble(&r, wt, vec, &handled)) return false;
    }
    if (!handled && !r.Skip(field, wt)) return false;
  }
  return true;
}

}  // namespace

bool ViMapMissions::Parse(const void* proto, size_t size, std::string* err) {
  *this = ViMapMissions();
  if (size > 0 && !proto) {
    *err = "vi_map: null buffer";
    return false;
  }
  Reader r{static_cast<const uint8_t*>(proto), static_cast<const uint8_t*>(proto) + size};
  std::vector<uint64_t> ids, base_of_mission, base_ids;
  std::vector<double> base_T;
  uint32_t field;
  int wt;
  bool ok = true;
  while (ok && r.p != r.end) {
    ok = Key(&r, &field, &wt);
    if (!ok) break;
    Reader sub;
    uint64_t id[2];
    if ((field == 5 || field == 7) && wt == 2) {
      ok = r.Sub(&sub) && ParseId(sub, id);
      if (!ok) break;
      std::vector<uint64_t>& dst = field == 5 ? ids : base_ids;
      dst.push_back(id[0]);
      dst.push_back(id[1]);
    } else if (field == 6 && wt == 2) {  // Mission: baseframe_id = 1
      std::vector<double> none;
      ok = r.Sub(&sub) && ParseIdAndVector(sub, 1, 0, id, &none);
      if (!ok) break;
      base_of_mission.push_back(id[0]);
      base_of_mission.push_back(id[1]);
    } else if (field == 8 && wt == 2) {  // MissionBaseframe: T_G_M = 1
      std::vector<double> T;
      ok = r.Sub(&sub) && ParseIdAndVector(sub, 0, 1, id, &T);
      if (ok && T.size() != 7) {
        *err = "vi_map: T_G_M is not a 7-vector (quaternion + position)";
        return false;
      }
      base_T.insert(base_T.end(), T.begin(), T.end());
    } else {
      ok = r.Skip(field, wt);
    }
  }
  if (!ok) {
    *err = "vi_map: malformed protobuf wire data (is the file still gzip-compressed?)";
    return false;
  }
  if (ids.size() != base_of_mission.size() || base_ids.size() / 2 != base_T.size() / 7) {
    *err = "vi_map: mission / base frame ids and messages differ in number";
    return false;
  }
  for (size_t m = 0; m < ids.size() / 2; ++m) {
    size_t b = 0;
    while (b < base_ids.size() / 2 &&
           !(base_ids[2 * b] == base_of_mission[2 * m] && base_ids[2 * b + 1] == base_of_mission[2 * m + 1]))
      ++b;
    if (b == base_ids.size() / 2) {
      *err = "vi_map: mission without base frame";
      return false;
    }
    mission_id.push_back(ids[2 * m]);
    mission_id.push_back(ids[2 * m + 1]);
    T_G_M.insert(T_G_M.end(), base_T.begin() + 7 * b, base_T.begin() + 7 * (b + 1));
  }
  return true;
}

bool ViMapVertices::Parse(const void* proto, size_t size, std::string* err) {
  *this = ViMapVertices();
  if (size > 0 && !proto) {
    *err = "vi_map: null buffer";
    return false;
  }
  err->clear();
  Reader r{static_cast<const uint8_t*>(proto), static_cast<const uint8_t*>(proto) + size};
  uint32_t field;
  int wt;
  bool ok = true;
  while (ok && r.p != r.end) {
    ok = Key(&r, &field, &wt);
    if (!ok) break;
    Reader sub;
    if (field == 1 && wt == 2) {
      uint64_t id[2];
      ok = r.Sub(&sub) && ParseId(sub, id);
      vertex_id.push_back(id[0]);
      vertex_id.push_back(id[1]);
    } else if (field == 2 && wt == 2) {
      ok = r.Sub(&sub) && ParseVertex(sub, this, err);
    } else {
      ok = r.Skip(field, wt);
    }
  }
  if (!ok) {
    if (err->empty()) *err = "vi_map: malformed protobuf wire data (is the file still gzip-compressed?)";
    return false;
  }
  if (vertex_id.size() != 2 * static_cast<size_t>(num_vertices())) {
    *err = "vi_map: vertex_ids and vertices differ in number (CHECK_EQ of deserializeVertices)";
    return false;
  }
  return true;
}

}  // namespace mlc
