// Shared host/device helpers of libern_b200: error plumbing, sortable candidate keys, candidate sink.
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/ern_b200.h"

namespace ern {

// ---- error plumbing (thread-local message behind ern_last_error) -------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define ERN_CUDA(expr)                                              \
  do {                                                              \
    cudaError_t _e = (expr);                                        \
    if (_e != cudaSuccess) return ::ern::cuda_fail(_e, #expr);      \
  } while (0)

#define ERN_REQUIRE(cond, ...)                                      \
  do {                                                              \
    if (!(cond)) {                                                  \
      ::ern::set_error(__VA_ARGS__);                                \
      return ERN_ERR_ARG;                                           \
    }                                                               \
  } while (0)

// ---- sortable 64-bit candidate keys ----------------------------------------------------------------
// key = (order-preserving bits of the fp32 ranking value) << 32 | (0xFFFFFFFF - global id)
// so that one unsigned compare orders by value descending, then id ascending.  key 0 = "no entry".
__host__ __device__ __forceinline__ uint32_t f32_to_ordered(float f) {
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(f);
#else
  uint32_t u;
  memcpy(&u, &f, 4);
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_f32(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}
__host__ __device__ __forceinline__ uint64_t make_key(float value, uint32_t gid) {
  return (static_cast<uint64_t>(f32_to_ordered(value)) << 32) | static_cast<uint64_t>(0xFFFFFFFFu - gid);
}
__host__ __device__ __forceinline__ float key_value(uint64_t key) {
  return ordered_to_f32(static_cast<uint32_t>(key >> 32));
}
__host__ __device__ __forceinline__ int32_t key_id(uint64_t key) {
  return static_cast<int32_t>(0xFFFFFFFFu - static_cast<uint32_t>(key));
}

// ranking value of a similarity (see ERN_RANK_* in ern_b200.h).  `s + 0.0f` canonicalises -0.
template <int kRankBy>
__device__ __forceinline__ float rank_value(float s) {
  if (kRankBy == ERN_RANK_REFERENCE) return -(1.0f - s);
  return s + 0.0f;
}

// ---- candidate sink: where the scoring kernels put (value, id) pairs that pass the threshold --------
// Per query:
//   prefix  [ERN_DENSE_ROWS] keys : slots [0, prev_count) hold the exact top-k of everything scored so far (written by
//           the selection kernel); the first ("dense") launch stores every score of the first rows here instead.
//   segs    [n_seg][seg_cap] keys : segment u has exactly ONE writer for a whole launch -- the persistent scoring unit
//           (CTA pair / CTA) number u -- which appends with a private register cursor (no atomics) and publishes its
//           count in seg_counts[q, u] when it leaves the query tile.  A segment never overflows: a gallery tile adds at
//           most TILE_G <= 256 keys, and after every tile a segment holding more than seg_cap - TILE_G keys is pruned
//           by its warp (warp_compact_segment) to the keys >= a pivot that at least k of them reach; the pivot is a
//           valid lower bound of the query's final k-th best value and is shared with every other unit through
//           thr_ord[q] (atomicMax on the order-preserving bits).  So a launch may cover any number of gallery rows in
//           any order and stays exact -- gallery order only changes how often segments are pruned.
//   thr_ord [1] : order-preserving bits of the current lower bound (f32_to_ordered(-inf) at start).
// The fp32 validation kernel, whose blocks are not persistent, treats the n_seg * seg_cap slots of a query as one
// segment with an atomic cursor in seg_counts[q, 0]; its launches are sized so that it cannot overflow.
struct CandidateSink {
  uint64_t* prefix;         // [nq, ERN_DENSE_ROWS]
  uint64_t* segs;           // [nq, n_seg, seg_cap]
  int32_t* seg_counts;      // [nq, n_seg]
  uint32_t* thr_ord;        // [nq]
  const int32_t* exclude;   // [nq] global id to drop, or nullptr
  int32_t* status;          // [4]
  int n_seg;
  int seg_cap;
  int k;
  int dense;                // 1: every row of [row_begin,row_end) is stored at prefix slot (row - row_begin)
  int64_t row_begin;        // shard-local gallery rows covered by this launch
  int64_t row_end;
  int64_t id_offset;        // global id of shard row 0
  int64_t nq;               // queries of this batch (the pointers above are already offset to its first query)
};

// dense launches: slot = row - row_begin, NaN scores and the excluded id become empty slots
__device__ __forceinline__ void sink_put_dense(const CandidateSink& s, int64_t q, int64_t row, float value,
                                               int32_t excl) {
  const uint32_t gid = static_cast<uint32_t>(row + s.id_offset);
  const bool drop = (static_cast<int32_t>(gid) == excl) || !(value == value);
  s.prefix[q * ERN_DENSE_ROWS + (row - s.row_begin)] = drop ? 0ull : make_key(value, gid);
}
// single-segment atomic append (fp32 validation kernel only)
__device__ __forceinline__ void sink_put_atomic(const CandidateSink& s, int64_t q, int64_t row, float value,
                                                int32_t excl) {
  const uint32_t gid = static_cast<uint32_t>(row + s.id_offset);
  if (static_cast<int32_t>(gid) == excl) return;
  const int pos = atomicAdd(&s.seg_counts[q * s.n_seg], 1);
  if (pos < s.n_seg * s.seg_cap) s.segs[q * static_cast<int64_t>(s.n_seg) * s.seg_cap + pos] = make_key(value, gid);
}


inline int cdiv(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }

}  // namespace ern
