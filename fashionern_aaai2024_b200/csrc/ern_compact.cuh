// In-kernel compaction of a candidate segment (see CandidateSink in ern_common.cuh).  Included by the scoring kernel only.
#pragma once
#include "ern_common.cuh"

namespace ern {

// Warp-cooperative in-place pruning of the segments whose owner lanes raised `full`.
// Called by a CONVERGED warp; lane L owns segment `seg` with `cnt` keys (k <= cnt <= 512).  For every full lane the
// whole warp loads that lane's keys (up to 16 per lane) and radix-searches, MSB first, for a pivot P such that at least k keys
// are >= P: every key below P is dropped (k keys of this query are >= P, so P is a valid lower bound of the query's
// final k-th best value and nothing that could still matter is lost), the survivors are rewritten densely, the owner
// gets the new count and threshold, and P is published to thr_ord for every other unit.
// The search does not have to find the exact k-th value: it starts at the highest bit in which the segment's values
// differ (one max + one min reduction skip the common prefix -- scores of a burst are close together) and stops as
// soon as no more than `keep_max` keys survive, typically after 2-4 one-bit steps instead of 32.  Only when more than
// keep_max keys carry the exact k-th VALUE does it fall through to the id half (ties -> lower id wins) and keep
// exactly k keys.
struct CompactResult {
  int cnt;
  float thr;
};
static __device__ __noinline__ CompactResult warp_compact_segment(uint64_t* seg, int cnt, float thr, uint32_t* thr_ord_q,
                                                           bool full, int k, int keep_max) {
  const unsigned kFull = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  unsigned todo = __ballot_sync(kFull, full);
  while (todo) {
    const int L = __ffs(todo) - 1;
    todo &= todo - 1;
    uint64_t* s = reinterpret_cast<uint64_t*>(__shfl_sync(kFull, reinterpret_cast<unsigned long long>(seg), L));
    const int n = __shfl_sync(kFull, cnt, L);
    __syncwarp();                                   // the owner's appends are visible to the whole warp
    constexpr int kPerLane = ERN_SEG_CAP / 32;
    uint32_t hi[kPerLane], lo[kPerLane];
    uint32_t vmax = 0u, vmin = 0xffffffffu;
#pragma unroll
    for (int j = 0; j < kPerLane; ++j) {
      const int idx = j * 32 + lane;
      const uint64_t key = idx < n ? s[idx] : 0ull;   // key 0 = empty: below every real key (ordered bits of a real value are never 0)
      hi[j] = static_cast<uint32_t>(key >> 32);
      lo[j] = static_cast<uint32_t>(key);
      if (idx < n) {
        vmax = max(vmax, hi[j]);
        vmin = min(vmin, hi[j]);
      }
    }
    vmax = __reduce_max_sync(kFull, vmax);
    vmin = __reduce_min_sync(kFull, vmin);
    // bits above `top` are common to every value of the segment
    const int top = 31 - __clz(vmax ^ vmin);          // -1 when all values are equal
    uint32_t pre = top >= 31 ? 0u : (vmax >> (top + 1)) << (top + 1);
    int rem = k;                                      // the k-th largest key is the rem-th largest of the current bucket
    int bucket = n;                                   // keys whose value matches `pre` down to the current bit
#pragma unroll 1
    for (int b = top; b >= 0 && (k - rem) + bucket > keep_max; --b) {
      const uint32_t cand = (pre | (1u << b)) >> b;
      int c = 0;
#pragma unroll
      for (int j = 0; j < kPerLane; ++j) c += ((hi[j] >> b) == cand) ? 1 : 0;
      c = __reduce_add_sync(kFull, c);
      if (c >= rem) {
        pre |= 1u << b;
        bucket = c;
      } else {
        rem -= c;
        bucket -= c;
      }
    }
    // every key with value >= pre survives: (k - rem) above the bucket + the bucket itself
    uint32_t lo_min = 0;
    if ((k - rem) + bucket > keep_max) {
      // more than keep_max keys down to the exact k-th value (all bits used): values tie across the k-th place;
      // the lower id (larger lo) wins and exactly k keys survive
      int r2 = rem;
#pragma unroll 1
      for (int b = 31; b >= 0; --b) {
        const uint32_t cand = (lo_min | (1u << b)) >> b;
        int c = 0;
#pragma unroll
        for (int j = 0; j < kPerLane; ++j) c += (hi[j] == pre && (lo[j] >> b) == cand) ? 1 : 0;
        c = __reduce_add_sync(kFull, c);
        if (c >= r2) lo_min |= 1u << b; else r2 -= c;
      }
    }
    __syncwarp();
    int base = 0;
#pragma unroll
    for (int j = 0; j < kPerLane; ++j) {
      const bool keep = hi[j] > pre || (hi[j] == pre && lo[j] >= lo_min && (hi[j] | lo[j]) != 0u);
      const unsigned m = __ballot_sync(kFull, keep);
      if (keep) s[base + __popc(m & ((1u << lane) - 1u))] = (static_cast<uint64_t>(hi[j]) << 32) | lo[j];
      base += __popc(m);
    }
    if (lane == L) {
      cnt = base;                                   // k <= cnt <= keep_max (== k on the tie path)
      thr = fmaxf(thr, ordered_to_f32(pre));
      atomicMax(thr_ord_q, pre);
    }
  }
  __syncwarp();
  CompactResult r;
  r.cnt = cnt;
  r.thr = thr;
  return r;
}

}  // namespace ern
