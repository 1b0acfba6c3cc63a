// Candidate-list compaction / k-way merge, Recall@K, CIRR subset recall and row L2-normalisation kernels.
// All of them are small HBM/latency-bound integer or reduction kernels next to the scoring GEMM.
#include "ern_internal.cuh"
#include "ern_select.cuh"
#include <cuda_fp16.h>

namespace ern {

// ------------------------------------------------------------------------------------------------------
// Top-k selection of one query's candidate keys: bitonic sort (descending) of up to ERN_LIST_CAP keys in
// shared memory.  Used (a) between scoring launches to tighten the per-query threshold, (b) to emit the
// final sorted [k] (value, id) lists, (c) to merge lists gathered from other shards / ranks.
// ------------------------------------------------------------------------------------------------------
constexpr int kSelectThreads = 256;


__device__ __forceinline__ void bitonic_sort_desc(uint64_t* a, int pow2, int tid) {
  for (int size = 2; size <= pow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int i = tid; i < (pow2 >> 1); i += kSelectThreads) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const uint64_t x = a[lo], y = a[hi];
        const bool desc = (lo & size) == 0;
        if ((x < y) == desc) {
          a[lo] = y;
          a[hi] = x;
        }
      }
    }
  }
  __syncthreads();
}

constexpr int kSmallSort = 256;  // candidates that are sorted directly / survivors of the radix select
constexpr int kMaxSegments = 256;  // >= persistent scoring units (148 CTAs on a B200)

__global__ void __launch_bounds__(kSelectThreads) select_topk_kernel(const SelectParams p, const int64_t nq) {
  __shared__ uint64_t keys[ERN_SORT_CAP];
  __shared__ uint64_t top[kSmallSort];
  __shared__ int hist[256];
  __shared__ int n_shared, m_shared, remaining_sh, ge_sh;
  __shared__ uint64_t prefix_sh;
  __shared__ int seg_off[kMaxSegments + 1];     // exclusive prefix sums of the segment counts of this query
  const int tid = threadIdx.x;
  const int k = p.k;
  // one block per query -- or, behind select_topk_warp_kernel, a small grid that walks the flags and only works on
  // the queries that kernel left over
  for (int64_t q = blockIdx.x; q < nq; q += gridDim.x) {
  if (p.sel_flags && p.sel_flags[q] == 0) continue;  // select_topk_warp_kernel has already done this query
  __syncthreads();                                   // (shared state of the previous query is no longer in use)
  // (read before the kernel's last step overwrites it; dense launches and merges have no bound yet)
  const uint32_t thr_floor = (p.n_lists == 0 && p.dense_count == 0 && p.thr_ord) ? p.thr_ord[q] : 0u;
  if (tid == 0) {
    n_shared = 0;
    m_shared = 0;
    remaining_sh = k;
    prefix_sh = 0;
    ge_sh = 0;
  }
  if (p.n_lists == 0 && p.dense_count == 0 && !p.single_segment) {
    // segment counts -> exclusive prefix sums (one coalesced read, one warp scan per 32 segments)
    const int32_t* sc = p.seg_counts + q * p.n_seg;
    if (tid < 32) {
      int carry = 0;
      for (int base = 0; base < p.n_seg; base += 32) {
        const int u = base + tid;
        const int c = u < p.n_seg ? min(sc[u], p.seg_cap) : 0;
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, incl, o);
          if (tid >= o) incl += v;
        }
        if (u < p.n_seg) seg_off[u] = carry + incl - c;
        carry += __shfl_sync(0xffffffffu, incl, 31);
      }
      if (tid == 0) seg_off[p.n_seg] = carry;
    }
  }
  __syncthreads();

  // ---- every non-empty candidate key of this query, straight from its sources (order is irrelevant) ------------
  auto scan_sources = [&](auto&& f) {
    if (p.n_lists > 0) {
      const int total = p.n_lists * p.k_in;
      for (int i = tid; i < total; i += kSelectThreads) {
        const uint64_t key = p.merge_src[(i / p.k_in) * p.list_stride + q * p.query_stride + (i % p.k_in)];
        if (key) f(key);
      }
      return;
    }
    const uint64_t* pre = p.prefix + q * ERN_DENSE_ROWS;
    const int np = p.dense_count > 0 ? p.dense_count : p.prev_counts[q];
    // keys below the query's current lower bound (tightened inside the scoring launch) cannot be among the k best
    const uint64_t floor_key = static_cast<uint64_t>(thr_floor) << 32;
    for (int i = tid; i < np; i += kSelectThreads) {
      const uint64_t key = pre[i];
      if (key && key >= floor_key) f(key);
    }
    if (p.dense_count > 0) return;
    const int32_t* sc = p.seg_counts + q * p.n_seg;
    const uint64_t* segs = p.segs + q * static_cast<int64_t>(p.n_seg) * p.seg_cap;
    if (p.single_segment) {
      const int cnt = min(sc[0], p.n_seg * p.seg_cap);
      for (int i = tid; i < cnt; i += kSelectThreads) {
        const uint64_t key = segs[i];
        if (key >= floor_key) f(key);
      }
    } else {
      // flat walk over the published entries of all segments: every thread's loads are independent of each other (a
      // warp-per-segment walk chained a count load and a key load per segment: ~10 dependent L2 round trips per query)
      const int total = seg_off[p.n_seg];
      for (int i = tid; i < total; i += kSelectThreads) {
        int lo = 0, hi = p.n_seg;                      // segment u with seg_off[u] <= i < seg_off[u + 1]
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (seg_off[mid] <= i) lo = mid; else hi = mid;
        }
        const uint64_t key = segs[static_cast<int64_t>(lo) * p.seg_cap + (i - seg_off[lo])];
        if (key >= floor_key) f(key);
      }
    }
  };

  // ---- count them; up to ERN_SORT_CAP are staged in shared memory (the usual case: all of them) ------------------
  scan_sources([&](uint64_t key) {
    const int pos = atomicAdd(&n_shared, 1);
    if (pos < ERN_SORT_CAP) keys[pos] = key;
  });
  __syncthreads();
  const int n = n_shared;
  const bool staged = n <= ERN_SORT_CAP;
  auto scan = [&](auto&& f) {
    if (staged) {
      for (int i = tid; i < n; i += kSelectThreads) f(keys[i]);
    } else {
      scan_sources(f);        // adversarially ordered galleries only: candidates are re-read from L2 per pass
    }
  };

  const uint64_t* sorted = keys;   // array whose first min(n,k) entries are the answer, descending
  int sorted_len = 0;
  if (n <= kSmallSort) {
    int pow2 = 32;
    while (pow2 < n) pow2 <<= 1;
    for (int i = n + tid; i < pow2; i += kSelectThreads) keys[i] = 0ull;
    bitonic_sort_desc(keys, pow2, tid);
    sorted_len = pow2;
  } else {
    // ---- radix select (8 bits per pass, MSB first) of the k-th largest key.  After the four passes over the
    //      ranking value it stops as soon as at most kSmallSort candidates reach the k-th value (ties included, so
    //      the result stays exact); otherwise the id half is resolved too and exactly k candidates remain.
    uint64_t mask = 0;
    for (int shift = 56; shift >= 0; shift -= 8) {
      hist[tid] = 0;
      __syncthreads();
      const uint64_t prefix = prefix_sh;
      scan([&](uint64_t key) {
        if ((key & mask) == prefix) atomicAdd(&hist[static_cast<int>((key >> shift) & 255u)], 1);
      });
      __syncthreads();
      if (tid < 32) {
        int c[8], sum = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          c[j] = hist[tid * 8 + j];
          sum += c[j];
        }
        // above = number of matching candidates in bins owned by higher lanes
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_down_sync(0xffffffffu, incl, o);
          if (tid + o < 32) incl += v;
        }
        int above = incl - sum;
        const int remaining = remaining_sh;
        __syncwarp();
        if (above < remaining && remaining <= above + sum) {
#pragma unroll
          for (int j = 7; j >= 0; --j) {
            if (above < remaining && remaining <= above + c[j]) {
              prefix_sh = prefix | (static_cast<uint64_t>(tid * 8 + j) << shift);
              remaining_sh = remaining - above;
              ge_sh = (k - (remaining - above)) + c[j];   // candidates >= the prefix found so far
              above = 1 << 30;   // found; stop matching
            } else if (above < (1 << 30)) {
              above += c[j];
            }
          }
        }
      }
      mask |= static_cast<uint64_t>(0xFFu) << shift;
      __syncthreads();
      if (shift == 32 && ge_sh <= kSmallSort) break;
    }
    const uint64_t thr64 = prefix_sh;
    scan([&](uint64_t key) {
      if (key >= thr64) {
        const int pos = atomicAdd(&m_shared, 1);
        if (pos < kSmallSort) top[pos] = key;
      }
    });
    __syncthreads();
    int m = m_shared;
    if (m > kSmallSort) {        // only possible with duplicated keys (never produced by the scoring kernels)
      if (tid == 0 && p.status) atomicAdd(&p.status[0], 1);
      m = kSmallSort;
    }
    int pow2 = 32;
    while (pow2 < m) pow2 <<= 1;
    for (int i = m + tid; i < pow2; i += kSelectThreads) top[i] = 0ull;
    bitonic_sort_desc(top, pow2, tid);
    sorted = top;
    sorted_len = pow2;
  }

  for (int j = tid; j < k; j += kSelectThreads) {
    const uint64_t key = (j < sorted_len) ? sorted[j] : 0ull;
    if (p.prefix) p.prefix[q * ERN_DENSE_ROWS + j] = key;
    if (p.out_keys) p.out_keys[q * k + j] = key;
    if (p.peer_keys) {
      const int64_t slot = (static_cast<int64_t>(p.rank) * p.nq_total + p.q_first + q) * k + j;
      for (int s = 0; s < p.world; ++s) p.peer_keys[s][slot] = key;
    }
    if (p.out_scores) p.out_scores[q * k + j] = key ? key_value(key) : -INFINITY;
    if (p.out_ids) p.out_ids[q * k + j] = key ? key_id(key) : -1;
  }
  // every pass over the sources is done (the sorts above end with a block barrier): reset the segment cursors
  if (p.n_lists == 0 && p.dense_count == 0) {
    if (p.single_segment) {
      if (tid == 0) {
        if (p.seg_counts[q * p.n_seg] > p.n_seg * p.seg_cap && p.status) atomicAdd(&p.status[0], 1);  // cannot happen
        p.seg_counts[q * p.n_seg] = 0;
      }
    } else {
      for (int u = tid; u < p.n_seg; u += kSelectThreads) p.seg_counts[q * p.n_seg + u] = 0;
    }
  }
  if (tid == 0) {
    const uint64_t kth = (k - 1 < sorted_len) ? sorted[k - 1] : 0ull;
    if (p.prev_counts) p.prev_counts[q] = n < k ? n : k;
    // fewer than k real candidates so far: no lower bound yet
    if (p.thr_ord) p.thr_ord[q] = kth ? static_cast<uint32_t>(kth >> 32) : f32_to_ordered(-INFINITY);
  }
  }  // query loop
}


// ------------------------------------------------------------------------------------------------------
// The same selection with ONE WARP per query and the candidates in registers (<= 32 per lane): the usual case between
// two scoring launches is a few hundred candidates per query, for which a 256-thread block spends its time in
// barriers, same-bin shared-memory atomics and dependent global loads (76 us per launch for 4096 queries; this one:
// a few us).  Exact top-k of the unique 64-bit keys: MSB-first one-bit radix search that skips the bits all values
// share and stops as soon as exactly k keys remain, then a 128-key warp bitonic sort.  Queries with more than
// kWarpSelMax candidates (ordered galleries) are flagged and left to select_topk_kernel, which runs right after.
// ------------------------------------------------------------------------------------------------------
constexpr int kWarpSelMax = 1024;
constexpr int kWarpsPerBlock = 8;

__global__ void __launch_bounds__(32 * kWarpsPerBlock) select_topk_warp_kernel(const SelectParams p, int64_t nq) {
  __shared__ int seg_off_s[kWarpsPerBlock][kMaxSegments + 1];
  __shared__ uint64_t sortbuf_s[kWarpsPerBlock][128];
  const unsigned kFull = 0xffffffffu;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t q = static_cast<int64_t>(blockIdx.x) * kWarpsPerBlock + w;
  if (q >= nq) return;
  const int k = p.k;
  int* seg_off = seg_off_s[w];
  uint64_t* sortbuf = sortbuf_s[w];
  const bool dense = p.dense_count > 0;
  const uint32_t thr_floor = (!dense && p.thr_ord) ? p.thr_ord[q] : 0u;
  const uint64_t floor_key = static_cast<uint64_t>(thr_floor) << 32;
  const int np = dense ? p.dense_count : p.prev_counts[q];
  int total = 0;
  if (!dense) {
    const int32_t* sc = p.seg_counts + q * p.n_seg;
    int carry = 0;
    for (int base = 0; base < p.n_seg; base += 32) {
      const int u = base + lane;
      const int c = u < p.n_seg ? min(sc[u], p.seg_cap) : 0;
      int incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += v;
      }
      if (u < p.n_seg) seg_off[u] = carry + incl - c;
      carry += __shfl_sync(kFull, incl, 31);
    }
    if (lane == 0) seg_off[p.n_seg] = carry;
    total = carry;
    __syncwarp();
  }
  if (np + total > kWarpSelMax) {
    if (lane == 0) p.sel_flags[q] = 1;
    return;
  }
  if (lane == 0) p.sel_flags[q] = 0;

  // ---- candidates -> registers (flat index lane + 32 j over [prefix | segment 0 | segment 1 | ...]) ----------------
  const uint64_t* pre = p.prefix + q * ERN_DENSE_ROWS;
  const uint64_t* segs = p.segs + q * static_cast<int64_t>(p.n_seg) * p.seg_cap;
  // pass 1 only issues the loads (the flat index grows by 32 per step, so the segment cursor only moves forward: no
  // search); pass 2 consumes them -- all of a lane's loads are in flight together
  uint64_t key[32];
  int seg_u = 0;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const int i = lane + 32 * j;
    key[j] = 0ull;
    if (i < np) {
      key[j] = pre[i];
    } else if (i < np + total) {
      const int f = i - np;
#pragma unroll 1
      while (seg_off[seg_u + 1] <= f) ++seg_u;     // (segment with seg_off[u] <= f < seg_off[u + 1]; empty ones are skipped)
      key[j] = segs[static_cast<int64_t>(seg_u) * p.seg_cap + (f - seg_off[seg_u])];
    }
  }
  uint32_t hi[32], lo[32];
  int n_real = 0;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const uint64_t kj = key[j] < floor_key ? 0ull : key[j];   // below the query's published lower bound: cannot be among the k best
    hi[j] = static_cast<uint32_t>(kj >> 32);
    lo[j] = static_cast<uint32_t>(kj);
    n_real += kj != 0ull;
  }
  n_real = __reduce_add_sync(kFull, n_real);

  // ---- which keys are among the k largest?  keep := hi > pre_hi || (hi == pre_hi && lo >= pre_lo) ------------------
  uint32_t pre_hi = 0u, pre_lo = 0u;               // n_real <= k: every real key stays (the key != 0 test below)
  if (n_real > k) {
    uint32_t vmax = 0u, vmin = 0xffffffffu;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if ((hi[j] | lo[j]) != 0u) {
        vmax = max(vmax, hi[j]);
        vmin = min(vmin, hi[j]);
      }
    }
    vmax = __reduce_max_sync(kFull, vmax);
    vmin = __reduce_min_sync(kFull, vmin);
    const int top = 31 - __clz(vmax ^ vmin);       // bits above `top` are shared by every value; -1: all values equal
    pre_hi = top >= 31 ? 0u : (vmax >> (top + 1)) << (top + 1);
    int rem = k, bucket = n_real;                  // the k-th largest key is the rem-th largest of the current bucket
#pragma unroll 1
    for (int b = top; b >= 0 && bucket != rem; --b) {
      const uint32_t cand = (pre_hi | (1u << b)) >> b;
      int c = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) c += ((hi[j] >> b) == cand) ? 1 : 0;
      c = __reduce_add_sync(kFull, c);
      if (c >= rem) { pre_hi |= 1u << b; bucket = c; } else { rem -= c; bucket -= c; }
    }
    if (bucket != rem) {
      // every value bit is decided and more keys than needed carry exactly the k-th value: the id half decides
      // (ids are unique; lower id = larger lo wins)
#pragma unroll 1
      for (int b = 31; b >= 0 && bucket != rem; --b) {
        const uint32_t cand = (pre_lo | (1u << b)) >> b;
        int c = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) c += (hi[j] == pre_hi && (lo[j] >> b) == cand) ? 1 : 0;
        c = __reduce_add_sync(kFull, c);
        if (c >= rem) { pre_lo |= 1u << b; bucket = c; } else { rem -= c; bucket -= c; }
      }
    }
  }
  const int m = n_real < k ? n_real : k;
  // ---- survivors -> 128-slot buffer ----------------------------------------------------------------------------------
  for (int i = lane; i < 128; i += 32) sortbuf[i] = 0ull;
  __syncwarp();
  int base = 0;
  uint32_t min_hi = 0xffffffffu;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const bool keep = (hi[j] | lo[j]) != 0u && (hi[j] > pre_hi || (hi[j] == pre_hi && lo[j] >= pre_lo));
    const unsigned mk = __ballot_sync(kFull, keep);
    if (keep) {
      sortbuf[base + __popc(mk & ((1u << lane) - 1u))] = (static_cast<uint64_t>(hi[j]) << 32) | lo[j];
      min_hi = min(min_hi, hi[j]);
    }
    base += __popc(mk);
  }
  __syncwarp();
  if (base != m && lane == 0 && p.status) atomicAdd(&p.status[0], 1);   // cannot happen (keys are unique)
  min_hi = __reduce_min_sync(kFull, min_hi);        // ranking value of the k-th best when m == k
  uint64_t e[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) e[r] = sortbuf[lane + 32 * r];
  // Only a launch that emits results needs them in order; between two scoring launches the prefix is a SET (both
  // selection kernels read it as one), so the 128-key bitonic sort -- a third of this kernel's instructions -- is skipped.
  const bool emit = p.out_keys || p.peer_keys || p.out_scores || p.out_ids;
  if (emit) {
#pragma unroll
    for (int size = 2; size <= 128; size <<= 1) {
#pragma unroll
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int idx = lane + 32 * r;
          const bool am_lower = (idx & stride) == 0;
          const bool desc = ((idx & ~stride) & size) == 0;
          uint64_t other;
          if (stride >= 32) {
            other = e[r ^ (stride >> 5)];
          } else {
            other = __shfl_xor_sync(kFull, e[r], stride);
          }
          const uint64_t mx = e[r] > other ? e[r] : other, mn = e[r] > other ? other : e[r];
          // (in-register partners are both updated in this pass: compute from the values before the pass)
          sortbuf[idx] = (desc == am_lower) ? mx : mn;
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 4; ++r) e[r] = sortbuf[lane + 32 * r];
        __syncwarp();
      }
    }
  }
  // ---- outputs (same contract as select_topk_kernel; the prefix is only sorted on emitting launches) ---------------
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int j = lane + 32 * r;
    if (j < k) {
      const uint64_t key = e[r];
      if (p.prefix) p.prefix[q * ERN_DENSE_ROWS + j] = key;
      if (p.out_keys) p.out_keys[q * k + j] = key;
      if (p.peer_keys) {
        const int64_t slot = (static_cast<int64_t>(p.rank) * p.nq_total + p.q_first + q) * k + j;
        for (int s2 = 0; s2 < p.world; ++s2) p.peer_keys[s2][slot] = key;
      }
      if (p.out_scores) p.out_scores[q * k + j] = key ? key_value(key) : -INFINITY;
      if (p.out_ids) p.out_ids[q * k + j] = key ? key_id(key) : -1;
    }
  }
  if (!dense)
    for (int u = lane; u < p.n_seg; u += 32) p.seg_counts[q * p.n_seg + u] = 0;
  if (lane == 0) {
    if (p.prev_counts) p.prev_counts[q] = m;
    if (p.thr_ord) p.thr_ord[q] = (m == k) ? min_hi : f32_to_ordered(-INFINITY);
  }
}

int launch_select(const SelectParams& p, int64_t nq, cudaStream_t st) {
  if (nq <= 0) return ERN_OK;
  // a query's own candidates (tensor-core path): warp-per-query kernel first; it flags the queries it leaves to the
  // block kernel (more than kWarpSelMax candidates), which returns immediately for all the others
  unsigned grid = static_cast<unsigned>(nq);
  if (p.sel_flags && p.n_lists == 0 && !p.single_segment) {
    select_topk_warp_kernel<<<cdiv(nq, kWarpsPerBlock), 32 * kWarpsPerBlock, 0, st>>>(p, nq);
    ERN_CUDA(cudaGetLastError());
    if (grid > 296u) grid = 296u;                   // leftovers are rare: two blocks per SM walk the flags
  }
  select_topk_kernel<<<grid, kSelectThreads, 0, st>>>(p, nq);
  ERN_CUDA(cudaGetLastError());
  return ERN_OK;
}

__global__ void init_state_kernel(int32_t* prev_counts, int32_t* seg_counts, uint32_t* thr_ord, int64_t nq, int n_seg,
                                  int32_t* status) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < nq) {
    prev_counts[i] = 0;
    thr_ord[i] = f32_to_ordered(-INFINITY);
  }
  if (i < nq * n_seg) seg_counts[i] = 0;
  if (i < 4 && status) status[i] = 0;
}

int launch_init_state(int32_t* prev_counts, int32_t* seg_counts, uint32_t* thr_ord, int64_t nq, int n_seg,
                      int32_t* status, cudaStream_t st) {
  const int64_t n = nq * n_seg < 4 ? 4 : nq * n_seg;
  init_state_kernel<<<cdiv(n, 256), 256, 0, st>>>(prev_counts, seg_counts, thr_ord, nq, n_seg, status);
  ERN_CUDA(cudaGetLastError());
  return ERN_OK;
}

// ------------------------------------------------------------------------------------------------------
// Recall@K: first position whose class equals the target's (run/test/test_fiq.py:51-60, test_200k.py:52-60)
// ------------------------------------------------------------------------------------------------------
constexpr int kMaxKs = 16;
struct KList {
  int32_t ks[kMaxKs];
  int nk;
};

__global__ void recall_kernel(const int32_t* __restrict__ top_ids, int64_t nq, int k,
                              const int32_t* __restrict__ class_of, int64_t n_gallery,
                              const int32_t* __restrict__ target_class, KList kl, int32_t* counts,
                              int32_t* rank_out) {
  // one warp per query
  const int64_t q = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (q >= nq) return;
  const int32_t tc = target_class[q];
  int best = k;
  for (int j = lane; j < k; j += 32) {
    const int32_t id = top_ids[q * k + j];
    if (id >= 0 && id < n_gallery && class_of[id] == tc) {
      best = j;
      break;
    }
  }
  for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
  if (lane == 0) {
    if (rank_out) rank_out[q] = best;
    for (int i = 0; i < kl.nk; ++i)
      if (best < kl.ks[i]) atomicAdd(&counts[i], 1);
  }
}

__global__ void zero_i32_kernel(int32_t* p, int n) {
  if (threadIdx.x < n) p[threadIdx.x] = 0;
}

int launch_recall(const int32_t* top_ids, int64_t nq, int k, const int32_t* class_of, int64_t n_gallery,
                  const int32_t* target_class, const int32_t* ks, int nk, int32_t* counts, int32_t* rank_out,
                  cudaStream_t st) {
  ERN_REQUIRE(nk >= 1 && nk <= kMaxKs, "nk must be in [1,%d]", kMaxKs);
  KList kl;
  kl.nk = nk;
  for (int i = 0; i < nk; ++i) kl.ks[i] = ks[i];
  zero_i32_kernel<<<1, 32, 0, st>>>(counts, nk);
  if (nq > 0) recall_kernel<<<cdiv(nq * 32, 256), 256, 0, st>>>(top_ids, nq, k, class_of, n_gallery, target_class, kl, counts, rank_out);
  ERN_CUDA(cudaGetLastError());
  return ERN_OK;
}

// ------------------------------------------------------------------------------------------------------
// CIRR subset recall (run/test/test_cirr.py:64-66,76-78): one warp per query scores the <= 8 group
// members against the query and ranks the target among the members that are not the reference image.
// ------------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float load_as_float(const T* p, int64_t i);
template <>
__device__ __forceinline__ float load_as_float<float>(const float* p, int64_t i) { return p[i]; }
template <>
__device__ __forceinline__ float load_as_float<__nv_bfloat16>(const __nv_bfloat16* p, int64_t i) {
  return __bfloat162float(p[i]);
}
template <>
__device__ __forceinline__ float load_as_float<__half>(const __half* p, int64_t i) { return __half2float(p[i]); }

constexpr int kMaxMembers = 8;

template <typename T>
__global__ void cirr_subset_kernel(const T* __restrict__ queries, int64_t nq, int64_t ldq,
                                   const T* __restrict__ gallery, int64_t n_rows, int64_t ldg, int dim,
                                   const int32_t* __restrict__ members, int m,
                                   const int32_t* __restrict__ ref_id, const int32_t* __restrict__ tgt_id,
                                   int rank_by, KList kl, int32_t* counts, int32_t* rank_out) {
  const int64_t q = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (q >= nq) return;
  float val[kMaxMembers];
  int32_t ids[kMaxMembers];
#pragma unroll
  for (int j = 0; j < kMaxMembers; ++j) {
    ids[j] = (j < m) ? members[q * m + j] : -1;
    float acc = 0.f;
    if (ids[j] >= 0 && ids[j] < n_rows) {
      for (int d = lane; d < dim; d += 32)
        acc = fmaf(load_as_float(queries, q * ldq + d), load_as_float(gallery, ids[j] * ldg + d), acc);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    val[j] = (rank_by == ERN_RANK_REFERENCE) ? -(1.0f - acc) : acc + 0.0f;
  }
  if (lane == 0) {
    const int32_t r = ref_id[q], t = tgt_id[q];
    int tpos = -1;
    for (int j = 0; j < m; ++j)
      if (ids[j] == t && ids[j] != r && ids[j] >= 0) tpos = j;
    int rank = -1;
    if (tpos >= 0) {
      rank = 0;
      for (int j = 0; j < m; ++j) {
        if (j == tpos || ids[j] < 0 || ids[j] == r || ids[j] == t) continue;
        bool dup = false;  // a member listed twice counts once
        for (int i = 0; i < j; ++i) dup |= (ids[i] == ids[j]);
        if (dup) continue;
        if (val[j] > val[tpos] || (val[j] == val[tpos] && ids[j] < t)) ++rank;
      }
    }
    if (rank_out) rank_out[q] = rank;
    if (rank >= 0)
      for (int i = 0; i < kl.nk; ++i)
        if (rank < kl.ks[i]) atomicAdd(&counts[i], 1);
  }
}

// ---- split form for row-sharded galleries: every rank scores the members it owns (0 elsewhere), the caller sums
//      the [nq, m] matrices over ranks (exactly one owner per member => the sum is exact), then ranks.
template <typename T>
__global__ void gather_scores_kernel(const T* __restrict__ queries, int64_t nq, int64_t ldq,
                                     const T* __restrict__ gallery, int64_t n_rows, int64_t ldg, int dim,
                                     int64_t id_offset, const int32_t* __restrict__ ids, int m,
                                     float* __restrict__ out) {
  const int64_t q = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (q >= nq) return;
  for (int j = 0; j < m; ++j) {
    const int64_t row = static_cast<int64_t>(ids[q * m + j]) - id_offset;
    float acc = 0.f;
    if (ids[q * m + j] >= 0 && row >= 0 && row < n_rows) {
      for (int d = lane; d < dim; d += 32)
        acc = fmaf(load_as_float(queries, q * ldq + d), load_as_float(gallery, row * ldg + d), acc);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[q * m + j] = acc;
  }
}

__global__ void cirr_rank_kernel(const float* __restrict__ scores, int64_t nq, const int32_t* __restrict__ members,
                                 int m, const int32_t* __restrict__ ref_id, const int32_t* __restrict__ tgt_id,
                                 int rank_by, KList kl, int32_t* counts, int32_t* rank_out) {
  const int64_t q = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (q >= nq) return;
  const int32_t r = ref_id[q], t = tgt_id[q];
  float val[kMaxMembers];
  int32_t ids[kMaxMembers];
  int tpos = -1;
  for (int j = 0; j < m; ++j) {
    ids[j] = members[q * m + j];
    const float s = scores[q * m + j];
    val[j] = (rank_by == ERN_RANK_REFERENCE) ? -(1.0f - s) : s + 0.0f;
    if (ids[j] == t && ids[j] != r && ids[j] >= 0) tpos = j;
  }
  int rank = -1;
  if (tpos >= 0) {
    rank = 0;
    for (int j = 0; j < m; ++j) {
      if (j == tpos || ids[j] < 0 || ids[j] == r || ids[j] == t) continue;
      bool dup = false;
      for (int i = 0; i < j; ++i) dup |= (ids[i] == ids[j]);
      if (dup) continue;
      if (val[j] > val[tpos] || (val[j] == val[tpos] && ids[j] < t)) ++rank;
    }
  }
  if (rank_out) rank_out[q] = rank;
  if (rank >= 0)
    for (int i = 0; i < kl.nk; ++i)
      if (rank < kl.ks[i]) atomicAdd(&counts[i], 1);
}

int launch_gather_scores(const void* queries, int64_t nq, int64_t ldq, const void* gallery, int64_t n_rows,
                         int64_t ldg, int dim, int dtype, int64_t id_offset, const int32_t* ids, int m, float* out,
                         cudaStream_t st) {
  ERN_REQUIRE(m >= 1 && m <= 64, "m must be in [1,64]");
  if (nq <= 0) return ERN_OK;
  const int grid = cdiv(nq * 32, 256);
  if (dtype == ERN_DTYPE_F32)
    gather_scores_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(queries), nq, ldq,
                                                      static_cast<const float*>(gallery), n_rows, ldg, dim, id_offset,
                                                      ids, m, out);
  else if (dtype == ERN_DTYPE_F16)
    gather_scores_kernel<__half><<<grid, 256, 0, st>>>(static_cast<const __half*>(queries), nq, ldq,
                                                       static_cast<const __half*>(gallery), n_rows, ldg, dim, id_offset,
                                                       ids, m, out);
  else
    gather_scores_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(queries), nq, ldq,
                                                              static_cast<const __nv_bfloat16*>(gallery), n_rows, ldg,
                                                              dim, id_offset, ids, m, out);
  ERN_CUDA(cudaGetLastError());
  return ERN_OK;
}

int launch_cirr_rank(const float* scores, int64_t nq, const int32_t* members, int m, const int32_t* ref_id,
                     const int32_t* tgt_id, int rank_by, const int32_t* ks, int nk, int32_t* counts, int32_t* rank_out,
                     cudaStream_t st) {
  ERN_REQUIRE(nk >= 1 && nk <= kMaxKs, "nk must be in [1,%d]", kMaxKs);
  ERN_REQUIRE(m >= 1 && m <= kMaxMembers, "group size m must be in [1,%d]", kMaxMembers);
  KList kl;
  kl.nk = nk;
  for (int i = 0; i < nk; ++i) kl.ks[i] = ks[i];
  zero_i32_kernel<<<1, 32, 0, st>>>(counts, nk);
  if (nq > 0) cirr_rank_kernel<<<cdiv(nq, 256), 256, 0, st>>>(scores, nq, members, m, ref_id, tgt_id, rank_by, kl, counts, rank_out);
  ERN_CUDA(cudaGetLastError());
  return ERN_OK;
}

int launch_cirr_subset(const void* queries, int64_t nq, int64_t ldq, const void* gallery, int64_t n_rows,
                       int64_t ldg, int dim, int dtype, const int32_t* members, int m, const int32_t* ref_id,
                       const int32_t* tgt_id, int rank_by, const int32_t* ks, int nk, int32_t* counts,
                       int32_t* rank_out, cudaStream_t st) {
  ERN_REQUIRE(nk >= 1 && nk <= kMaxKs, "nk must be in [1,%d]", kMaxKs);
  ERN_REQUIRE(m >= 1 && m <= kMaxMembers, "group size m must be in [1,%d]", kMaxMembers);
  KList kl;
  kl.nk = nk;
  for (int i = 0; i < nk; ++i) kl.ks[i] = ks[i];
  zero_i32_kernel<<<1, 32, 0, st>>>(counts, nk);
  if (nq > 0) {
    const int grid = cdiv(nq * 32, 256);
    if (dtype == ERN_DTYPE_F32)
      cirr_subset_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(queries), nq, ldq,
                                                      static_cast<const float*>(gallery), n_rows, ldg, dim, members,
                                                      m, ref_id, tgt_id, rank_by, kl, counts, rank_out);
    else if (dtype == ERN_DTYPE_F16)
      cirr_subset_kernel<__half><<<grid, 256, 0, st>>>(
          static_cast<const __half*>(queries), nq, ldq, static_cast<const __half*>(gallery), n_rows, ldg, dim, members,
          m, ref_id, tgt_id, rank_by, kl, counts, rank_out);
    else
      cirr_subset_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(
          static_cast<const __nv_bfloat16*>(queries), nq, ldq, static_cast<const __nv_bfloat16*>(gallery), n_rows,
          ldg, dim, members, m, ref_id, tgt_id, rank_by, kl, counts, rank_out);
  }
  ERN_CUDA(cudaGetLastError());
  return ERN_OK;
}

// ------------------------------------------------------------------------------------------------------
// Row L2-normalise (F.normalize eps 1e-12) + optional bf16 cast; one warp per row, HBM-bound.
// ------------------------------------------------------------------------------------------------------
__global__ void l2norm_rows_kernel(const float* __restrict__ x, int64_t rows, int dim, int64_t ldx, int normalize,
                                   float* __restrict__ of, int64_t ldf, uint16_t* __restrict__ ob, int64_t ldb,
                                   int out_f16) {
  const int64_t r = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* xr = x + r * ldx;
  float denom = 1.f;
  if (normalize) {
    float ss = 0.f;
    for (int d = lane; d < dim; d += 32) ss = fmaf(xr[d], xr[d], ss);
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    denom = fmaxf(sqrtf(ss), 1e-12f);  // x / max(||x||, eps), torch's F.normalize
  }
  for (int d = lane; d < dim; d += 32) {
    const float v = normalize ? xr[d] / denom : xr[d];
    if (of) of[r * ldf + d] = v;
    if (ob) ob[r * ldb + d] = out_f16 ? __half_as_ushort(__float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f)))
                                      : __bfloat16_as_ushort(__float2bfloat16_rn(v));
  }
}

int launch_l2norm_rows(const float* x, int64_t rows, int dim, int64_t ldx, int normalize, float* of, int64_t ldf,
                       void* ob, int64_t ldb, cudaStream_t st) {
  if (rows <= 0) return ERN_OK;
  l2norm_rows_kernel<<<cdiv(rows * 32, 256), 256, 0, st>>>(x, rows, dim, ldx, normalize & 1, of, ldf,
                                                           static_cast<uint16_t*>(ob), ldb, (normalize & ERN_NORM_OUT_F16) ? 1 : 0);
  ERN_CUDA(cudaGetLastError());
  return ERN_OK;
}

}  // namespace ern
