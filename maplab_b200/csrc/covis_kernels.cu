// Kernel 3 of the loop-closure path: match gathering with the time/mission filter, per-keyframe
// vote counting, top-fraction selection and landmark-covisibility clustering — one CTA per query
// frame (pass 1) or per multi-camera query vertex (pass 2). Everything is sort / scan / segmented
// reduction in shared memory; no atomics except one order-independent integer add (component sizes).
//
// Reference: matching-based-loopclosure/src/matching-based-engine.cc:101-123 (neighbour walk with
// `break`), :170-215 (getMatchForDescriptorIndex), :147-165 (vertex pass);
// include/matching-based-loopclosure/matching-based-engine-inl.h:45-183 (doCovisibilityFiltering),
// :219-254 (computeRelevantIdsForFiltering); scoring.h:38-59 (accumulation score), :92-187
// (probabilistic score).
// Canonical orders where the reference depends on hash-map iteration (oracle/engine.cc):
// top-fraction ties by keyframe number, component ties by smallest member, duplicates keep the
// smallest database descriptor, output sorted by (query frame, keypoint, database descriptor).
#include <cub/cub.cuh>

#include <cstdlib>

#include "covis.h"

namespace mlc {
namespace {

constexpr uint16_t kNone16 = 0xFFFFu;

template <int MAXM>
struct CovisSmem {
  unsigned long long keys[MAXM];
  uint16_t a1[MAXM + 2], a2[MAXM + 2], a3[MAXM + 2], a4[MAXM + 2];
  uint16_t a5[MAXM + 2], a6[MAXM + 2], a7[MAXM + 2], a8[MAXM + 2];
  uint8_t f1[MAXM], f2[MAXM];
  int bcast[4];
};

// Bitonic sort of keys[0, p2) in shared memory (p2 a power of two); every thread of the block must call
// it. One compare-exchange per pair and step. Pair t of a step with stride j touches the elements
// i = ((t & ~(j-1)) << 1) | (t & (j-1)) and i | j: for j <= 32 the 32 pairs of a warp stay inside one
// 64-element span that no other warp touches, so those steps only need a warp barrier; a block barrier is
// needed only around the steps with larger strides (20 instead of 66 block barriers at 2048 keys).
template <int THREADS>
__device__ __forceinline__ void BitonicSort(unsigned long long* keys, int p2) {
  for (int k2 = 2; k2 <= p2; k2 <<= 1) {
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (p2 >> 1); t += THREADS) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int ixj = i | j;
        const bool asc = (i & k2) == 0;
        const unsigned long long a = keys[i], b = keys[ixj];
        if ((a > b) == asc) {
          keys[i] = b;
          keys[ixj] = a;
        }
      }
      const int next_j = j > 1 ? (j >> 1) : k2;  // first stride of the next merge: (2 * k2) / 2
      if (j > 32 || next_j > 32)
        __syncthreads();
      else
        __syncwarp();
    }
  }
  __syncthreads();
}

__device__ __forceinline__ int NextPow2(int n) {
  int p = 2;
  while (p < n) p <<= 1;
  return p;
}

// Order-preserving exclusive scan of 0/1 flags over [0, n): thread t owns the contiguous chunk
// [t*IPT, (t+1)*IPT). pos[i] = number of set flags before i. Returns the total.
template <int THREADS, int IPT, typename FlagFn, typename OutFn>
__device__ __forceinline__ int FlagScan(int n, FlagFn flag, OutFn out,
                                        typename cub::BlockScan<int, THREADS>::TempStorage& tmp) {
  const int begin = threadIdx.x * IPT;
  int local = 0;
#pragma unroll
  for (int i = 0; i < IPT; ++i) {
    const int at = begin + i;
    if (at < n && flag(at)) ++local;
  }
  int offset = 0, total = 0;
  cub::BlockScan<int, THREADS>(tmp).ExclusiveSum(local, offset, total);
  __syncthreads();
#pragma unroll
  for (int i = 0; i < IPT; ++i) {
    const int at = begin + i;
    if (at < n) {
      const bool f = flag(at);
      out(at, offset, f);
      if (f) ++offset;
    }
  }
  __syncthreads();
  return total;
}

// boost::math::pdf(binomial(n, p), k) = C(n,k) p^k (1-p)^(n-k) (special cases first), fp64.
__device__ double BinomialPdf(double n, double p, double k) {
  if (p == 0) return (k == 0) ? 1.0 : 0.0;
  if (p == 1) return (k == n) ? 1.0 : 0.0;
  if (n == 0) return 1.0;
  if (k == 0) return pow(1 - p, n);
  if (k == n) return pow(p, k);
  const double lg = lgamma(n + 1) - lgamma(k + 1) - lgamma(n - k + 1) + k * log(p) + (n - k) * log1p(-p);
  return exp(lg);
}

// Score of one candidate id (scoring.h:139-178): -log10 of the probability that `votes` of the
// `total` votes fall on an id holding `num_desc` of the `num_db` database descriptors by chance;
// 0 below the expected count; FLT_MAX (+ *underflow) when the pdf underflows to 0.
__device__ float ProbabilisticScore(unsigned votes, unsigned total, unsigned num_desc, long long num_db,
                                    bool* underflow) {
  const double p = static_cast<double>(num_desc) / static_cast<double>(num_db);
  const unsigned long long lower_median = static_cast<unsigned long long>(static_cast<double>(total) * p);
  if (!(votes > lower_median)) return 0.f;
  const double prob = BinomialPdf(static_cast<double>(total), p, static_cast<double>(votes));
  if (prob == 0.0) {
    *underflow = true;
    return 3.402823466e+38f;
  }
  return static_cast<float>(-log10(prob));
}

template <int MAXM, int THREADS>
__global__ void __launch_bounds__(THREADS, (THREADS <= 512 ? 2 : 1)) covis_kernel(CovisArgs a) {
  constexpr int IPT = MAXM / THREADS;
  constexpr int SLOT_BITS = (MAXM == 8192) ? 13 : 12;
  constexpr unsigned long long SLOT_MASK = (1ull << SLOT_BITS) - 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CovisSmem<MAXM>& s = *reinterpret_cast<CovisSmem<MAXM>*>(smem_raw);
  __shared__ typename cub::BlockScan<int, THREADS>::TempStorage scan_tmp;
  typedef cub::BlockReduce<unsigned long long, THREADS> BlockMax;
  __shared__ typename BlockMax::TempStorage red_tmp;
  __shared__ int s_count[2][32];
  mlc_match* rec = a.scratch + static_cast<size_t>(blockIdx.x) * MAXM;
  const int tid = threadIdx.x;

  for (int w = blockIdx.x; w < a.num_items; w += gridDim.x) {
    const CovisItem item = a.items[w];
    // ------------------------------------------------------------------ load + compact (P5)
    int R = 0;
    if (!a.by_vertex) {
      // pass 1: neighbour slots of one query frame, scan order = (keypoint, neighbour rank)
      const int n_slots = item.count * a.k;
      const int32_t* idx = a.knn_idx + static_cast<size_t>(item.first) * a.k;
      const float* dst = a.knn_dist + static_cast<size_t>(item.first) * a.k;
      auto valid = [&](int slot) -> bool {
        const int32_t id = idx[slot];
        if (id < 0 || dst[slot] == __int_as_float(0x7f800000)) return false;  // trailing (-1, inf)
        const KeyframeMeta& kf = a.kf_meta[a.desc_kf[id]];
        long long dt = item.ts - kf.ts;
        if (dt < 0) dt = -dt;
        // skip iff |dt| < min_seconds * 1e9 AND same mission (matching-based-engine.cc:192-198)
        return !(static_cast<double>(dt) < a.min_time_ns && item.mission == kf.mission);
      };
      // evaluate the filter once per slot (two dependent random gathers), then compact
      for (int slot = tid; slot < n_slots; slot += THREADS) s.f1[slot] = valid(slot) ? 1 : 0;
      __syncthreads();
      R = FlagScan<THREADS, IPT>(
          n_slots, [&](int slot) { return s.f1[slot] != 0; },
          [&](int slot, int pos, bool f) {
            if (!f) return;
            const int32_t id = idx[slot];
            const int32_t kfn = a.desc_kf[id];
            mlc_match m;
            m.query_frame = item.frame;
            m.query_keypoint = slot / a.k;
            m.db_descriptor = id;
            m.db_keyframe = kfn;
            m.db_vertex = a.kf_meta[kfn].vertex;
            m.landmark = a.desc_lm[id];
            rec[pos] = m;
          },
          scan_tmp);
    } else {
      // pass 2: concatenation of the pass-1 outputs of the vertex' frames
      for (int f = 0; f < item.count; ++f) {
        const int fi = item.first + f;
        const int cnt = a.in_counts[fi];
        const mlc_match* src = a.in_matches + a.in_offsets[fi];
        for (int i = tid; i < cnt; i += THREADS)
          if (R + i < MAXM) rec[R + i] = src[i];
        R += cnt;
      }
    }
    __syncthreads();
    int out_count = 0;
    mlc_match* out = a.out_matches + item.out_offset;
    if (R > MAXM) {
      out_count = -1;  // more surviving matches than the kernel can cluster: reported to the host
    } else if (R > 0) {
      // ---------------------------------------------------------------- A: sort by group, votes
      const int p2 = NextPow2(R);
      const int fill = (!a.by_vertex && p2 > 1024) ? MAXM : p2;
      for (int i = tid; i < fill; i += THREADS) {
        unsigned long long key = ~0ull;
        if (i < R) {
          const long long g = a.by_vertex ? rec[i].db_vertex : static_cast<long long>(rec[i].db_keyframe);
          key = (static_cast<unsigned long long>(g) << SLOT_BITS) | static_cast<unsigned>(i);
        }
        s.keys[i] = key;
      }
      __syncthreads();
      if (!a.by_vertex && p2 > 1024) {
        // large frame: stable LSD radix sort over the keyframe-number bits only (the slot bits are
        // already ascending in the blocked arrangement), ~5x fewer shared-memory passes than the
        // bitonic network; its scratch aliases a3..a8, which are dead until stage B
        typedef cub::BlockRadixSort<unsigned long long, THREADS, IPT> RadixSort;
        static_assert(sizeof(typename RadixSort::TempStorage) + 16 <= 6 * sizeof(s.a3), "radix scratch");
        typename RadixSort::TempStorage& radix_tmp = *reinterpret_cast<typename RadixSort::TempStorage*>(
            (reinterpret_cast<uintptr_t>(s.a3) + 15) & ~static_cast<uintptr_t>(15));
        unsigned long long kreg[IPT];
#pragma unroll
        for (int i = 0; i < IPT; ++i) kreg[i] = s.keys[tid * IPT + i];  // MAXM >= p2; pad = ~0
        __syncthreads();
        RadixSort(radix_tmp).Sort(kreg, SLOT_BITS, SLOT_BITS + a.group_bits);
        __syncthreads();
#pragma unroll
        for (int i = 0; i < IPT; ++i) s.keys[tid * IPT + i] = kreg[i];
        __syncthreads();
      } else {
        BitonicSort<THREADS>(s.keys, p2);
      }
      // candidate rank b of every match (a1), first sorted position of every candidate (a2)
      const int Cn = FlagScan<THREADS, IPT>(
          R, [&](int i) { return i == 0 || (s.keys[i] >> SLOT_BITS) != (s.keys[i - 1] >> SLOT_BITS); },
          [&](int i, int pos, bool f) {
            const int b = f ? pos : pos - 1;
            const unsigned slot = static_cast<unsigned>(s.keys[i] & SLOT_MASK);
            s.a1[slot] = static_cast<uint16_t>(b);
            s.a8[i] = static_cast<uint16_t>(slot);  // sorted position -> match, kept until stage C
            if (f) s.a2[b] = static_cast<uint16_t>(i);
          },
          scan_tmp);
      if (tid == 0) s.a2[Cn] = static_cast<uint16_t>(R);
      __syncthreads();
      // ---------------------------------------------------------------- B: scores, top fraction
      int n_eval = Cn;
      if (!a.by_vertex) {
        // computeRelevantIdsForFiltering: max(floor(float(size) * fraction), 4), capped by size
        int want = static_cast<int>(__fmul_rn(static_cast<float>(Cn), a.fraction_best_scores));
        if (want < 4) want = 4;
        n_eval = want < Cn ? want : Cn;
        unsigned* prob_score = reinterpret_cast<unsigned*>(s.a5);  // a5/a6 are free until stage C
        if (a.scoring == 1) {
          // computeProbabilisticScore (scoring.h:92-187), candidates in ascending keyframe number
          // (the canonical replacement of the hash-map iteration order, oracle/engine.cc)
          int underflow = 0;
          for (int b = tid; b < Cn; b += THREADS) {
            const int kfn = rec[s.keys[s.a2[b]] & SLOT_MASK].db_keyframe;
            bool uf = false;
            prob_score[b] = __float_as_uint(ProbabilisticScore(
                static_cast<unsigned>(s.a2[b + 1] - s.a2[b]), static_cast<unsigned>(R),
                static_cast<unsigned>(a.kf_meta[kfn].num_descriptors), a.num_db_descriptors, &uf));
            s.f2[b] = uf ? 1 : 0;
            underflow |= uf ? 1 : 0;
          }
          if (__syncthreads_or(underflow)) {
            // quirk 7 (scoring.h:180-184): the +inf patch runs inside the per-id loop, so the id
            // whose pdf underflowed with the most votes so far becomes +inf only if the NEXT id
            // does not take that role over; the last id is never patched.
            if (tid == 0) {
              unsigned best = 0;
              int holder = -1;
              for (int b = 0; b < Cn; ++b) {
                const unsigned m = static_cast<unsigned>(s.a2[b + 1] - s.a2[b]);
                const bool takes_over = s.f2[b] && m > best;
                if (takes_over) {
                  best = m;
                  holder = b;
                } else if (holder >= 0) {
                  prob_score[holder] = 0x7f800000u;
                }
              }
            }
            __syncthreads();
          }
        }
        // The n_eval best candidates by (score descending, keyframe number ascending) — a SELECTION, the order
        // among them is never used: find the n_eval-th largest score T bit by bit (block-wide counts), take every
        // candidate above T and the first ones in keyframe order that equal T. (The bitonic sort this replaces
        // ran over next_pow2(candidates) 64-bit keys — ~1 500 candidates per frame at k = 6 — and was the largest
        // single piece of the kernel.)
        unsigned kreg[IPT];
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
          const int b = tid + i * THREADS;
          unsigned key = 0;
          if (b < Cn)  // accumulation score = votes (ordered like its float); probabilistic scores are >= 0:
            key = a.scoring == 1 ? prob_score[b] : static_cast<unsigned>(s.a2[b + 1] - s.a2[b]);  // bits order them
          kreg[i] = key;
        }
        __syncthreads();  // prob_score (a5/a6) has been read; count_tmp of a previous frame is free
        auto block_count = [&](int local, int buf) -> int {  // one barrier per call, buffers alternate
          const int wsum = __reduce_add_sync(0xffffffffu, local);
          if ((tid & 31) == 0) s_count[buf][tid >> 5] = wsum;
          __syncthreads();
          const int v = (tid & 31) < THREADS / 32 ? s_count[buf][tid & 31] : 0;
          return __reduce_add_sync(0xffffffffu, v);
        };
        unsigned T = 0;
        int buf = 0;
        for (int bit = (a.scoring == 1 ? 30 : SLOT_BITS); bit >= 0; --bit) {  // votes <= R <= MAXM = 2^SLOT_BITS
          const unsigned cand = T | (1u << bit);
          int local = 0;
#pragma unroll
          for (int i = 0; i < IPT; ++i) local += (kreg[i] >= cand) ? 1 : 0;  // padding keys are 0 < cand
          if (block_count(local, buf) >= n_eval) T = cand;
          buf ^= 1;
        }
        int above = 0;
#pragma unroll
        for (int i = 0; i < IPT; ++i) above += (kreg[i] > T) ? 1 : 0;
        const int room = n_eval - block_count(above, buf);  // >= 1 candidates equal to T are taken, in keyframe order
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
          const int b = tid + i * THREADS;
          if (b < Cn) {
            s.f1[b] = kreg[i] > T ? 1 : 0;
            s.f2[b] = kreg[i] == T ? 1 : 0;
          }
        }
        __syncthreads();
        FlagScan<THREADS, IPT>(
            Cn, [&](int b) { return s.f2[b] != 0; },
            [&](int b, int pos, bool f) { if (f && pos < room) s.f1[b] = 1; }, scan_tmp);
      } else {
        for (int b = tid; b < Cn; b += THREADS) s.f1[b] = 1;
        __syncthreads();
      }
      // dense id of the selected candidates (a3), ascending group id
      FlagScan<THREADS, IPT>(
          Cn, [&](int b) { return s.f1[b] != 0; },
          [&](int b, int pos, bool f) { s.a3[b] = f ? static_cast<uint16_t>(pos) : kNone16; }, scan_tmp);
      // a1[match] := selected dense group id or none
      for (int i = tid; i < R; i += THREADS) s.a1[i] = s.a3[s.a1[i]];
      __syncthreads();
      // ---------------------------------------------------------------- C: landmark ids, both orders
      // keys of the selected matches only (compacted in slot order), so that the remaining sorts
      // run over next_pow2(S) instead of next_pow2(R) keys
      const int S = FlagScan<THREADS, IPT>(
          R, [&](int i) { return s.a1[i] != kNone16; },
          [&](int i, int pos, bool f) {
            if (f) s.keys[pos] = (static_cast<unsigned long long>(rec[i].landmark + 1) << SLOT_BITS) | static_cast<unsigned>(i);
          },
          scan_tmp);
      const int ps = NextPow2(S);
      if (S > 1024 && a.landmark_bits <= 30) {
        // many selected matches: stable LSD radix sort over the landmark bits only (the slot bits already
        // ascend in the blocked arrangement), like stage A; its scratch aliases a2..a7, dead until the
        // scans below
        for (int j = S + tid; j < MAXM; j += THREADS) s.keys[j] = ~0ull;
        __syncthreads();
        typedef cub::BlockRadixSort<unsigned long long, THREADS, IPT> RadixSort;
        static_assert(sizeof(typename RadixSort::TempStorage) + 16 <= 6 * sizeof(s.a2), "radix scratch");
        typename RadixSort::TempStorage& radix_tmp = *reinterpret_cast<typename RadixSort::TempStorage*>(
            (reinterpret_cast<uintptr_t>(s.a2) + 15) & ~static_cast<uintptr_t>(15));
        unsigned long long kreg[IPT];
#pragma unroll
        for (int i = 0; i < IPT; ++i) kreg[i] = s.keys[tid * IPT + i];
        __syncthreads();
        RadixSort(radix_tmp).Sort(kreg, SLOT_BITS, SLOT_BITS + a.landmark_bits);
        __syncthreads();
#pragma unroll
        for (int i = 0; i < IPT; ++i) s.keys[tid * IPT + i] = kreg[i];
        __syncthreads();
      } else {
        for (int j = S + tid; j < ps; j += THREADS) s.keys[j] = ~0ull;
        __syncthreads();
        BitonicSort<THREADS>(s.keys, ps);
      }
      // landmark dense id a4[match]; offA (a5), listA (a6) = group of the match at sorted position
      const int nL = FlagScan<THREADS, IPT>(
          S, [&](int i) { return i == 0 || (s.keys[i] >> SLOT_BITS) != (s.keys[i - 1] >> SLOT_BITS); },
          [&](int i, int pos, bool f) {
            const int la = f ? pos : pos - 1;
            const int slot = static_cast<int>(s.keys[i] & SLOT_MASK);
            s.a4[slot] = static_cast<uint16_t>(la);
            s.a6[i] = s.a1[slot];
            if (f) s.a5[la] = static_cast<uint16_t>(i);
          },
          scan_tmp);
      if (tid == 0) s.a5[nL] = static_cast<uint16_t>(S);
      __syncthreads();
      // order B: selected matches by (dense group id, slot) = the stage-A order (a8) restricted to
      // the selected matches, because dense ids ascend with the group id — no sort; offB (a7)
      FlagScan<THREADS, IPT>(
          R, [&](int pp) { return s.a1[s.a8[pp]] != kNone16; },
          [&](int pp, int pos, bool f) {
            if (!f) return;
            const unsigned slot = s.a8[pp];
            const uint16_t g = s.a1[slot];
            s.keys[pos] = (static_cast<unsigned long long>(g) << SLOT_BITS) | slot;
            if (pp == 0 || s.a1[s.a8[pp - 1]] != g) s.a7[g] = static_cast<uint16_t>(pos);
          },
          scan_tmp);
      if (tid == 0) s.a7[n_eval] = static_cast<uint16_t>(S);
      // listB (a8, overwriting the stage-A order it no longer needs) = landmark id at order-B position
      for (int j = tid; j < S; j += THREADS) s.a8[j] = s.a4[s.keys[j] & SLOT_MASK];
      __syncthreads();
      // ---------------------------------------------------------------- D: components (min label)
      uint16_t* K = s.a2;  // label per selected group
      uint16_t* L = s.a3;  // label per landmark
      for (int b = tid; b < n_eval; b += THREADS) K[b] = static_cast<uint16_t>(b);
      __syncthreads();
      for (;;) {
        for (int la = tid; la < nL; la += THREADS) {
          uint16_t mn = kNone16;
          for (int i = s.a5[la]; i < s.a5[la + 1]; ++i) mn = min(mn, K[s.a6[i]]);
          L[la] = mn;
        }
        __syncthreads();
        int changed = 0;
        for (int b = tid; b < n_eval; b += THREADS) {
          uint16_t mn = K[b];
          for (int i = s.a7[b]; i < s.a7[b + 1]; ++i) mn = min(mn, L[s.a8[i]]);
          if (mn != K[b]) {
            K[b] = mn;
            changed = 1;
          }
        }
        if (!__syncthreads_or(changed)) break;
      }
      // ---------------------------------------------------------------- E: distinct matches, sizes
      // Duplicates (same query keypoint, keyframe and landmark, different database descriptor)
      // can only sit among the <= k neighbours of one keypoint, which are adjacent in `rec`.
      // (the records live in global scratch: the shared-memory ids — selected group a1, dense landmark a4,
      // query keypoint a5 — rule a neighbour out before its record is fetched)
      uint16_t* qkp = s.a5;  // offA is no longer needed
      for (int i = tid; i < R; i += THREADS) qkp[i] = static_cast<uint16_t>(rec[i].query_keypoint);
      __syncthreads();
      for (int i = tid; i < R; i += THREADS) {
        uint8_t distinct = 0;
        const uint16_t g = s.a1[i];
        if (g != kNone16) {
          distinct = 1;
          const uint16_t la = s.a4[i], kp = qkp[i];
          for (int d = -(a.k - 1); d <= a.k - 1; ++d) {
            const int j = i + d;
            if (d == 0 || j < 0 || j >= R || s.a1[j] != g || s.a4[j] != la || qkp[j] != kp) continue;
            const mlc_match me = rec[i], o = rec[j];
            if (o.query_frame == me.query_frame && o.query_keypoint == me.query_keypoint &&
                o.db_keyframe == me.db_keyframe && o.landmark == me.landmark &&
                o.db_descriptor < me.db_descriptor)
              distinct = 0;
          }
        }
        s.f2[i] = distinct;
      }
      __syncthreads();
      uint16_t* cnt = s.a3;  // distinct matches per selected group (L no longer needed)
      for (int b = tid; b < n_eval; b += THREADS) {
        int c = 0;
        for (int i = s.a7[b]; i < s.a7[b + 1]; ++i) c += s.f2[s.keys[i] & SLOT_MASK];
        cnt[b] = static_cast<uint16_t>(c);
      }
      __syncthreads();
      // component sizes: every group adds its distinct matches to its root. Integer adds in shared memory — the
      // sums do not depend on the order, so the result stays deterministic (the only atomics of the kernel; a
      // root walking all groups for its members was a 400-iteration serial loop per root). The sort keys are
      // dead from here on: their storage holds the sizes.
      int* size = reinterpret_cast<int*>(s.keys);  // (cnt is complete and the keys have been read: barrier above)
      for (int b = tid; b < n_eval; b += THREADS) size[b] = K[b] == b ? cnt[b] : 0;
      __syncthreads();
      for (int b = tid; b < n_eval; b += THREADS)
        if (K[b] != b) atomicAdd(&size[K[b]], static_cast<int>(cnt[b]));
      __syncthreads();
      unsigned long long best = 0;
      for (int r = tid; r < n_eval; r += THREADS) {
        if (K[r] != r) continue;  // not a root
        // larger size wins; ties -> smaller root
        const unsigned long long cand = (static_cast<unsigned long long>(size[r]) << 16) | (0xFFFFu - r);
        if (cand > best) best = cand;
      }
      best = BlockMax(red_tmp).Reduce(best, cub::Max());
      if (tid == 0) {
        s.bcast[0] = static_cast<int>(best >> 16);
        s.bcast[1] = 0xFFFF - static_cast<int>(best & 0xFFFFu);
      }
      __syncthreads();
      const int best_size = s.bcast[0], best_root = s.bcast[1];
      if (static_cast<unsigned long long>(best_size) > a.min_verify_matches_num) {
        // -------------------------------------------------------------- F: winners, uniqueness, order
        for (int i = tid; i < R; i += THREADS)
          s.f1[i] = (s.f2[i] && K[s.a1[i]] == best_root) ? 1 : 0;
        __syncthreads();
        const bool unique = a.by_vertex ? true : (item.make_unique != 0);
        // f2[i] := match i is emitted (winner, and the smallest database descriptor of its
        // (query keypoint, landmark) when uniqueness is enforced)
        for (int i = tid; i < R; i += THREADS) {
          uint8_t keep = s.f1[i];
          if (keep && unique) {
            const uint16_t la = s.a4[i], kp = qkp[i];
            for (int d = -(a.k - 1); d <= a.k - 1; ++d) {
              const int o_i = i + d;
              if (d == 0 || o_i < 0 || o_i >= R || !s.f1[o_i] || s.a4[o_i] != la || qkp[o_i] != kp) continue;
              const mlc_match me = rec[i], o = rec[o_i];
              if (o.query_frame == me.query_frame && o.query_keypoint == me.query_keypoint &&
                  o.landmark == me.landmark && o.db_descriptor < me.db_descriptor)
                keep = 0;
            }
          }
          s.f2[i] = keep;
        }
        __syncthreads();
        // canonical order (query frame, keypoint, database descriptor): `rec` is already ordered by
        // (query frame, keypoint) with the <= k entries of one keypoint adjacent, so the output
        // position is the compaction rank corrected by the entry's rank among its keypoint's
        // emitted entries — no sort
        out_count = FlagScan<THREADS, IPT>(
            R, [&](int i) { return s.f2[i] != 0; },
            [&](int i, int pos, bool f) {
              if (!f) return;
              const mlc_match me = rec[i];
              const uint16_t kp = qkp[i];
              int before = 0, smaller = 0;
              for (int d = -(a.k - 1); d <= a.k - 1; ++d) {
                const int o_i = i + d;
                if (d == 0 || o_i < 0 || o_i >= R || !s.f2[o_i] || qkp[o_i] != kp) continue;
                const mlc_match o = rec[o_i];
                if (o.query_frame != me.query_frame || o.query_keypoint != me.query_keypoint) continue;
                if (d < 0) ++before;
                if (o.db_descriptor < me.db_descriptor) ++smaller;
              }
              out[pos - before + smaller] = me;
            },
            scan_tmp);
      }
    }
    if (tid == 0) a.out_counts[w] = out_count;
    __syncthreads();
  }
}

// scoring::compute{Accumulation,Probabilistic}Score for an explicit id list (mlc_score): the same
// device functions the covisibility kernel uses, ids in the order given.
__global__ void score_kernel(const unsigned long long* __restrict__ votes,
                             const unsigned long long* __restrict__ num_desc, int n, long long num_db,
                             int scoring, float* __restrict__ out) {
  __shared__ unsigned long long total_s;
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int i = 0; i < n; ++i) t += votes[i];
    total_s = t;
  }
  __syncthreads();
  const unsigned total = static_cast<unsigned>(total_s);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float score = static_cast<float>(votes[i]);
    if (scoring == 1) {
      bool uf = false;
      score = ProbabilisticScore(static_cast<unsigned>(votes[i]), total, static_cast<unsigned>(num_desc[i]),
                                 num_db, &uf);
    }
    out[i] = score;
  }
  __syncthreads();
  if (scoring == 1 && threadIdx.x == 0) {
    unsigned long long best = 0;
    int holder = -1;
    for (int i = 0; i < n; ++i) {
      const bool takes_over = out[i] == 3.402823466e+38f && votes[i] > best;
      if (takes_over) {
        best = votes[i];
        holder = i;
      } else if (holder >= 0) {
        out[holder] = __int_as_float(0x7f800000);
      }
    }
  }
}

__global__ void compact_matches_kernel(const mlc_match* __restrict__ in, const CovisItem* items,
                                       const int* __restrict__ counts,
                                       const long long* __restrict__ dst_offsets, int num_items,
                                       mlc_match* __restrict__ out) {
  for (int w = blockIdx.x; w < num_items; w += gridDim.x) {
    const mlc_match* src = in + items[w].out_offset;
    mlc_match* dst = out + dst_offsets[w];
    const int n = counts[w];
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    for (int i = threadIdx.x; i < n * 2; i += blockDim.x) d4[i] = s4[i];
  }
}

}  // namespace

// Resident CTAs per SM of the kernel LaunchCovis picks (the caller sizes its grid with it).
int CovisCtasPerSm(int max_matches) {
  // one 1024-thread CTA per SM (measured on the headline step: 0.92 ms against 1.02 ms with two 512-thread
  // CTAs: the frame finishes sooner and the last wave is finer); MLC_COVIS_THREADS=512 selects the latter
  static const bool narrow = [] {
    const char* env = getenv("MLC_COVIS_THREADS");
    return env && atoi(env) == 512;
  }();
  return (max_matches > 4096 || !narrow) ? 1 : 2;
}

size_t CovisScratchMatches(int max_matches, int grid) {
  return static_cast<size_t>(max_matches > 4096 ? 8192 : 4096) * grid;
}

cudaError_t LaunchCovis(const CovisArgs& a, int max_matches, int grid, cudaStream_t stream) {
  if (a.num_items <= 0) return cudaSuccess;
  if (max_matches > 8192 || a.k > 16) return cudaErrorInvalidValue;
  cudaError_t e;
  if (max_matches <= 4096 && CovisCtasPerSm(max_matches) == 1) {
    auto fn = covis_kernel<4096, 1024>;
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(sizeof(CovisSmem<4096>)));
    if (e != cudaSuccess) return e;
    fn<<<grid, 1024, sizeof(CovisSmem<4096>), stream>>>(a);
  } else if (max_matches <= 4096) {
    auto fn = covis_kernel<4096, 512>;
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(sizeof(CovisSmem<4096>)));
    if (e != cudaSuccess) return e;
    fn<<<grid, 512, sizeof(CovisSmem<4096>), stream>>>(a);
  } else {
    auto fn = covis_kernel<8192, 1024>;
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(sizeof(CovisSmem<8192>)));
    if (e != cudaSuccess) return e;
    fn<<<grid, 1024, sizeof(CovisSmem<8192>), stream>>>(a);
  }
  CountLaunch();
  return cudaGetLastError();
}

cudaError_t LaunchScore(const unsigned long long* votes, const unsigned long long* num_desc, int n,
                        long long num_db, int scoring, float* out, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  score_kernel<<<1, 256, 0, stream>>>(votes, num_desc, n, num_db, scoring, out);
  CountLaunch();
  return cudaGetLastError();
}

cudaError_t LaunchCompactMatches(const mlc_match* in, const CovisItem* items, const int* counts,
                                 const long long* dst_offsets, int num_items, mlc_match* out,
                                 cudaStream_t stream) {
  if (num_items <= 0) return cudaSuccess;
  const int grid = num_items < 1184 ? num_items : 1184;
  compact_matches_kernel<<<grid, 128, 0, stream>>>(in, items, counts, dst_offsets, num_items, out);
  CountLaunch();
  return cudaGetLastError();
}

}  // namespace mlc
