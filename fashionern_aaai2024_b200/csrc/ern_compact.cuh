// In-kernel compaction of a candidate segment (see CandidateSink in ern_common.cuh).  Included by the scoring kernel only.
#pragma once
#include "ern_common.cuh"

namespace ern {

// Warp-cooperative in-place selection of the k largest keys of the segments whose owner lanes raised `full`.
// Called by a CONVERGED warp; lane L owns segment `seg` with `cnt` keys (k <= cnt <= 256).  For every full lane the
// whole warp loads that lane's keys (8 per lane), radix-selects the k-th largest 32-bit ranking value bit by bit
// (one warp reduction per bit; the id half is only walked when values tie across the k-th place), rewrites the k
// survivors densely, and the owner gets cnt = k and a threshold >= the segment's k-th best value, which is also
// published to thr_ord (k keys of this query are >= it, so it is a valid lower bound of the final k-th best).
struct CompactResult {
  int cnt;
  float thr;
};
static __device__ __noinline__ CompactResult warp_compact_segment(uint64_t* seg, int cnt, float thr, uint32_t* thr_ord_q,
                                                           bool full, int k) {
  const unsigned kFull = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  unsigned todo = __ballot_sync(kFull, full);
  while (todo) {
    const int L = __ffs(todo) - 1;
    todo &= todo - 1;
    uint64_t* s = reinterpret_cast<uint64_t*>(__shfl_sync(kFull, reinterpret_cast<unsigned long long>(seg), L));
    const int n = __shfl_sync(kFull, cnt, L);
    __syncwarp();                                   // the owner's appends are visible to the whole warp
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int idx = j * 32 + lane;
      const uint64_t key = idx < n ? s[idx] : 0ull;
      hi[j] = static_cast<uint32_t>(key >> 32);
      lo[j] = static_cast<uint32_t>(key);
    }
    uint32_t pre = 0;
    int rem = k;
#pragma unroll 1
    for (int b = 31; b >= 0; --b) {
      const uint32_t cand = (pre | (1u << b)) >> b;
      int c = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) c += ((hi[j] >> b) == cand) ? 1 : 0;
      c = __reduce_add_sync(kFull, c);
      if (c >= rem) pre |= 1u << b; else rem -= c;
    }
    // pre = k-th largest ranking value; `rem` of the keys that carry exactly this value belong to the k best
    int ceq = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) ceq += (hi[j] == pre) ? 1 : 0;
    ceq = __reduce_add_sync(kFull, ceq);
    uint32_t lo_min = 0;
    if (ceq != rem) {                               // values tie across the k-th place: lower id (larger lo) wins
      int r2 = rem;
#pragma unroll 1
      for (int b = 31; b >= 0; --b) {
        const uint32_t cand = (lo_min | (1u << b)) >> b;
        int c = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) c += (hi[j] == pre && (lo[j] >> b) == cand) ? 1 : 0;
        c = __reduce_add_sync(kFull, c);
        if (c >= r2) lo_min |= 1u << b; else r2 -= c;
      }
    }
    __syncwarp();
    int base = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const bool keep = hi[j] > pre || (hi[j] == pre && lo[j] >= lo_min);
      const unsigned m = __ballot_sync(kFull, keep);
      if (keep) s[base + __popc(m & ((1u << lane) - 1u))] = (static_cast<uint64_t>(hi[j]) << 32) | lo[j];
      base += __popc(m);
    }
    if (lane == L) {
      cnt = base;                                   // == k
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
