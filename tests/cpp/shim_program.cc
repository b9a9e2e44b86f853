// Exercises include/maplab_lc_b200_shim.h end to end through the C-ABI (compiled by
// tests/test_shim_cpp.py with plain g++). Input: a binary file written by the test
//   int64 header[8] = {num_db_frames, num_db_desc, num_q_frames, num_q_desc, dim, bytes, vocab_bytes, num_landmarks}
//   vocabulary blob, db frames (mlc_frame[]), db descriptor bits, db landmark numbers (int64[]),
//   landmark positions (double[3 L]), query frames, query bits, query keypoints (double[2 nq])
// Output (stdout): "vertices accepted matches inliers" + one line per vertex; exit code 3 when the
// library reports no usable device (CPU box).
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <vector>

#include "maplab_lc_b200_shim.h"

template <typename T>
static std::vector<T> ReadVec(std::ifstream& in, size_t n) {
  std::vector<T> v(n);
  in.read(reinterpret_cast<char*>(v.data()), static_cast<std::streamsize>(sizeof(T) * n));
  return v;
}

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  std::ifstream in(argv[1], std::ios::binary);
  const std::vector<int64_t> h = ReadVec<int64_t>(in, 8);
  const std::vector<char> vocab = ReadVec<char>(in, static_cast<size_t>(h[6]));
  const auto db_frames = ReadVec<mlc_frame>(in, static_cast<size_t>(h[0]));
  const auto db_bits = ReadVec<uint8_t>(in, static_cast<size_t>(h[1] * h[5]));
  const auto db_lm = ReadVec<int64_t>(in, static_cast<size_t>(h[1]));
  const auto xyz = ReadVec<double>(in, static_cast<size_t>(3 * h[7]));
  const auto q_frames = ReadVec<mlc_frame>(in, static_cast<size_t>(h[2]));
  const auto q_bits = ReadVec<uint8_t>(in, static_cast<size_t>(h[3] * h[5]));
  const auto q_kp = ReadVec<double>(in, static_cast<size_t>(2 * h[3]));
  mlc_settings s;
  mlc_default_settings(&s);
  s.num_nearest_neighbors = 6;
  try {
    maplab_lc_b200::LoopDetector det(s, vocab.data(), vocab.size());
    const int dim = det.dim();
    size_t at = 0;
    for (const mlc_frame& f : db_frames) {  // addVertexToDatabase: project + Insert per frame
      maplab_lc_b200::ProjectedImage img;
      img.timestamp_nanoseconds = f.timestamp_ns;
      img.vertex_id = f.vertex_id;
      img.frame_index = f.frame_index;
      img.mission_id = f.mission_id;
      img.projected_descriptors.resize(static_cast<size_t>(f.num_descriptors) * dim);
      det.ProjectDescriptors(db_bits.data() + at * h[5], static_cast<int>(h[5]), f.num_descriptors,
                             img.projected_descriptors.data());
      img.landmarks.assign(db_lm.begin() + at, db_lm.begin() + at + f.num_descriptors);
      det.Insert(img);
      at += static_cast<size_t>(f.num_descriptors);
    }
    det.SetLandmarkPositions(xyz.data(), h[7]);
    det.Initialize();
    mlc_camera cam;
    std::memset(&cam, 0, sizeof(cam));
    cam.fu = cam.fv = 400.0;
    cam.cu = 376.0;
    cam.cv = 240.0;
    cam.R_B_C[0] = cam.R_B_C[4] = cam.R_B_C[8] = 1.0;
    mlc_ransac_settings rs;
    mlc_default_ransac_settings(&rs);
    std::vector<mlc_pose_result> verdicts;
    std::vector<std::vector<mlc_match>> inliers;
    det.QueryBatch(q_frames, q_bits.data(), static_cast<int>(h[5]), q_kp.data(), {cam}, rs, &verdicts, &inliers);
    int accepted = 0;
    size_t num_inlier_matches = 0;
    for (size_t v = 0; v < verdicts.size(); ++v) {
      accepted += verdicts[v].accepted;
      num_inlier_matches += inliers[v].size();
    }
    std::printf("%zu %d %zu %d\n", verdicts.size(), accepted, num_inlier_matches, det.NumDescriptors());
    for (size_t v = 0; v < verdicts.size(); ++v)
      std::printf("%d %d %d %.17g %.17g %.17g\n", verdicts[v].accepted, verdicts[v].num_inliers, verdicts[v].iterations,
                  verdicts[v].T_G_I[3], verdicts[v].T_G_I[7], verdicts[v].T_G_I[11]);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "shim: %s\n", e.what());
    return 3;
  }
  return 0;
}
