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
  // ---- source B (a query's own candidate list, see CandidateSink)
  uint64_t* lists;           // [nq, cap]; the compacted top-k is written back to slots [0,k)
  int cap;
  int keep;                  // survivors of earlier launches live in slots [0, prev_counts[q])
  int32_t* prev_counts;      // [nq]  in: survivors, out: min(n, k)
  int32_t* seg_counts;       // [nq, ERN_MAX_CHUNKS]; reset to 0 after reading
  int n_chunks;
  int seg_size;
  int dense_count;           // > 0: slots [0, dense_count) are all candidates (first launch), segments unused
  int k;
  // ---- outputs (nullable)
  float* thresholds;         // [nq]
  float* out_scores;         // [nq, k]
  int32_t* out_ids;          // [nq, k]
  uint64_t* out_keys;        // [nq, k]
  // ---- fused candidate exchange (nullable): peer_keys[s] is rank s's gathered buffer [world, nq, k]; this rank's
  //      final keys are stored straight into slot `rank` of every peer's buffer over NVLink (P2P stores)
  uint64_t* const* peer_keys;
  int world;
  int rank;
  int64_t nq_total;
  int32_t* status;
};
}  // namespace ern
