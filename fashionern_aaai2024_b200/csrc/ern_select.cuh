// Parameters of the candidate-list top-k selection kernel (ern_topk.cu).
#pragma once
#include "ern_common.cuh"
namespace ern {
struct SelectParams {
  // ---- source A (merge): `n_lists` > 0 segments of k_in keys per query, laid out with the two strides
  const uint64_t* merge_src;
  int64_t list_stride;
  int64_t query_stride;
  int n_lists;
  int k_in;
  // ---- source B (a query's own candidates, see CandidateSink): prefix slots [0, prev_counts[q]) (or [0, dense_count)
  //      after the dense launch) + the published part of every segment.  The top-k goes back to prefix [0,k).
  uint64_t* prefix;          // [nq, ERN_DENSE_ROWS]
  int32_t* prev_counts;      // [nq]  in: survivors, out: min(n, k)
  const uint64_t* segs;      // [nq, n_seg, seg_cap]
  int32_t* seg_counts;       // [nq, n_seg]; reset to 0 after reading
  int n_seg;
  int seg_cap;
  int single_segment;        // fp32 validation kernel: one segment of n_seg * seg_cap slots, cursor in seg_counts[q,0]
  int dense_count;           // > 0: prefix slots [0, dense_count) are all candidates (first launch), segments unused
  int k;
  // ---- outputs (nullable)
  uint32_t* thr_ord;         // [nq] exact k-th best ranking value so far (order-preserving bits)
  float* out_scores;         // [nq, k]
  int32_t* out_ids;          // [nq, k]
  uint64_t* out_keys;        // [nq, k]
  // ---- fused candidate exchange (nullable): peer_keys[s] is rank s's gathered buffer [world, nq_total, k]; this
  //      rank's final keys are stored straight into slot `rank` of every peer's buffer over NVLink (P2P stores)
  uint64_t* const* peer_keys;
  int world;
  int rank;
  int64_t nq_total;          // queries of the whole call (row stride of the peer buffers)
  int64_t q_first;           // index of this batch's first query within the call (peer buffer row offset)
  int32_t* status;
  // [nq] (nullable) written by select_topk_warp_kernel: 0 = done by it, 1 = left to select_topk_kernel
  int32_t* sel_flags;
};
}  // namespace ern
