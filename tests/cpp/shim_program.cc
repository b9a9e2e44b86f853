// Exercises include/maplab_lc_b200_shim.h end to end through the C-ABI (compiled by
// tests/test_shim_cpp.py with plain g++). Input: a binary file written by the test
//   int64 header[8] = {num_db_frames, num_db_desc, num_q_frames, num_q_desc, dim, bytes, vocab_bytes, num_landmarks}
//   vocabulary blob, db frames (mlc_frame[]), db descriptor bits, db landmark numbers (int64[]),
//   landmark positions (double[3 L]), query frames, query bits, query keypoints (double[2 nq])
// Output (stdout): "vertices accepted matches inliers" + one line per vertex; exit code 3 when the
// library reports no usable device (CPU box).
// Multi-GPU mode: shim_program <world file> <rank> <world size> <id file> — one process per GPU (rank r on
// CUDA device r), the database sharded (every rank projects and inserts only the descriptors it owns), the
// NCCL communicator id handed over through <id file>, every rank queries its slice of the query vertices
// with ShardedQueryBatch and prints the lines of its slice.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <thread>
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
  const int rank = argc >= 5 ? std::atoi(argv[2]) : 0, world = argc >= 5 ? std::atoi(argv[3]) : 1;
  const bool sharded = argc >= 5;
  if (sharded) {
    s.shard_rank = rank;
    s.shard_count = world;
    s.device = rank;
  }
  try {
    maplab_lc_b200::LoopDetector det(s, vocab.data(), vocab.size());
    const int dim = det.dim();
    size_t at = 0;
    if (sharded) {
      // communicator: rank 0 writes the id file, the others wait for it
      std::vector<char> id;
      if (rank == 0) {
        id = maplab_lc_b200::LoopDetector::MakeCommunicatorId();
        std::ofstream tmp(std::string(argv[4]) + ".tmp", std::ios::binary);
        tmp.write(id.data(), static_cast<std::streamsize>(id.size()));
        tmp.close();
        std::rename((std::string(argv[4]) + ".tmp").c_str(), argv[4]);
      } else {
        for (int tries = 0; tries < 600; ++tries) {
          std::ifstream f(argv[4], std::ios::binary);
          id.assign(MLC_COMM_ID_BYTES, 0);
          if (f.read(id.data(), MLC_COMM_ID_BYTES)) break;
          id.clear();
          std::this_thread::sleep_for(std::chrono::milliseconds(100));
        }
      }
      det.InitCommunicator(id);
      // sharded build: project + hand over the owned descriptors of every keyframe only
      for (const mlc_frame& f : db_frames) {
        maplab_lc_b200::ProjectedImage img;
        img.timestamp_nanoseconds = f.timestamp_ns;
        img.vertex_id = f.vertex_id;
        img.frame_index = f.frame_index;
        img.mission_id = f.mission_id;
        img.landmarks.assign(db_lm.begin() + at, db_lm.begin() + at + f.num_descriptors);
        std::vector<uint8_t> own_bits;
        for (int i = 0; i < f.num_descriptors; ++i)
          if (static_cast<int>((at + i) % world) == rank)
            own_bits.insert(own_bits.end(), db_bits.begin() + (at + i) * h[5], db_bits.begin() + (at + i + 1) * h[5]);
        const int64_t owned = det.NumOwnedInRange(det.NumDescriptors(), f.num_descriptors);
        if (owned * h[5] != static_cast<int64_t>(own_bits.size())) throw std::runtime_error("owned row count");
        std::vector<float> own_proj(static_cast<size_t>(owned) * dim);
        det.ProjectDescriptors(own_bits.data(), static_cast<int>(h[5]), owned, own_proj.data());
        det.InsertOwned(img, f.num_descriptors, own_proj.data());
        at += static_cast<size_t>(f.num_descriptors);
      }
    }
    for (const mlc_frame& f : db_frames) {  // addVertexToDatabase: project + Insert per frame
      if (sharded) break;
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
    if (sharded) {
      // contiguous slice of the query frames (one frame per vertex in this world)
      const size_t nq = q_frames.size(), f0 = nq * rank / world, f1 = nq * (rank + 1) / world;
      size_t d0 = 0, d1 = 0;
      for (size_t f = 0; f < f1; ++f) {
        if (f < f0) d0 += q_frames[f].num_descriptors;
        d1 += q_frames[f].num_descriptors;
      }
      const std::vector<mlc_frame> slice(q_frames.begin() + f0, q_frames.begin() + f1);
      det.ShardedQueryBatch(slice, q_bits.data() + d0 * h[5], static_cast<int>(h[5]), q_kp.data() + 2 * d0, {cam}, rs,
                            &verdicts, &inliers);
      (void)d1;
    } else {
      det.QueryBatch(q_frames, q_bits.data(), static_cast<int>(h[5]), q_kp.data(), {cam}, rs, &verdicts, &inliers);
    }
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
