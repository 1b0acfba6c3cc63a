// Parameters of the candidate-list top-k selection kernel (ern_topk.cu).
#pragma once
#include "ern_common.cuh"
namespace ern {
struct SelectParams {
  // source: either `n_lists` segments (merge) or the query's own list with `counts` / dense_count
  const uint64_t* src;
  int64_t list_stride;   // between segments of one query (merge), unused otherwise
  int64_t query_stride;  // between queries
  int n_lists;
  int k_in;              // entries per segment (merge)
  const int32_t* counts_in;  // nullable: per query number of valid entries in src (filter mode)
  int dense_count;       // entries per query when counts_in == nullptr && n_lists == 1
  int cap;
  int k;
  // outputs (all nullable)
  uint64_t* list_out;    // [nq, out_stride] compacted list written back (first k entries)
  int64_t out_stride;
  int32_t* counts_out;   // [nq]
  float* thresholds;     // [nq]
  float* out_scores;     // [nq, k]
  int32_t* out_ids;      // [nq, k]
  uint64_t* out_keys;    // [nq, k]
  int32_t* status;
};
}  // namespace ern
