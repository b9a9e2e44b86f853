// Kernel 3 interface (covis_kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/maplab_lc_b200.h"
#include "detector.h"

namespace mlc {

// One unit of work of the covisibility kernel: a query frame (pass 1) or a query vertex (pass 2).
struct CovisItem {
  int32_t first;       // pass 1: first query descriptor of the frame; pass 2: first frame item
  int32_t count;       // pass 1: number of descriptors; pass 2: number of frames
  int32_t frame;       // batch frame number (pass 2: of the first frame)
  int32_t make_unique; // pass 1 only: enforce (keypoint, landmark) uniqueness
  int64_t ts, mission; // query frame timestamp / mission (pass 1)
  int64_t out_offset;  // where this item's matches go in out_matches
};

struct CovisArgs {
  const CovisItem* items;
  int num_items;
  int by_vertex;  // 0: pass 1 (group = result keyframe, top-fraction), 1: pass 2 (group = result vertex)
  int k;
  const int32_t* knn_idx;  // pass 1 inputs
  const float* knn_dist;
  const int32_t* desc_kf;
  const int64_t* desc_lm;
  const KeyframeMeta* kf_meta;
  const mlc_match* in_matches;  // pass 2 inputs = pass 1 outputs
  const int* in_counts;
  const long long* in_offsets;
  double min_time_ns;
  unsigned long long min_verify_matches_num;
  float fraction_best_scores;
  int group_bits;                // 2^group_bits > number of database keyframes
  int landmark_bits;             // 2^landmark_bits > every (landmark number + 1) of the database
  int scoring;                   // 0 accumulation, 1 probabilistic (scoring.h)
  long long num_db_descriptors;  // whole database (all shards)
  mlc_match* scratch;  // grid * (4096 | 8192) records
  mlc_match* out_matches;
  int* out_counts;
};

size_t CovisScratchMatches(int max_matches, int grid);
int CovisCtasPerSm(int max_matches);
cudaError_t LaunchCovis(const CovisArgs& a, int max_matches, int grid, cudaStream_t stream);
cudaError_t LaunchScore(const unsigned long long* votes, const unsigned long long* num_desc, int n,
                        long long num_db, int scoring, float* out, cudaStream_t stream);
cudaError_t LaunchCompactMatches(const mlc_match* in, const CovisItem* items, const int* counts,
                                 const long long* dst_offsets, int num_items, mlc_match* out,
                                 cudaStream_t stream);

}  // namespace mlc
