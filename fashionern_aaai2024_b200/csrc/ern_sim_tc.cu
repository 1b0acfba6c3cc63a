// bf16 similarity + streaming top-k candidate filter on the 5th-gen tensor cores (tcgen05 / TMEM / TMA).
//
// Computes, for a range of gallery rows, S = Q . G^T (fp32 accumulate) tile by tile and hands every score
// that reaches the query's current threshold to the candidate sink -- the [Q,N] score matrix of the
// reference (`1 - predicted_features @ index_features.T`, run/test/test_fiq.py:49) never exists in HBM.
//
// Shape of the computation per CTA (or CTA pair, kPair):
//   A = 128 queries x D bf16, RESIDENT in shared memory for a whole work item (160 KB at D = 640),
//   B = gallery rows streamed by TMA in 128-row x 64-column (16 KB, 128B-swizzled) stages; ring depth = 14 - D/64 (4 at D = 640, 6 at D = 512),
//   D = 128 x TILE_G fp32 accumulator in TMEM, double buffered (2 x TILE_G columns),
//   kPair: the two CTAs of a cluster issue ONE cta_group::2 UMMA of M = 256 (128 queries per CTA) by
//          N = 256 (each CTA streams half of the gallery tile), halving shared-memory operand traffic.
// Warp roles: warps 0..3 = epilogue (TMEM lane quadrant = warp id), warp 4 = TMA producer, warp 5 = TMEM allocator +
// single-thread MMA issuer (leader CTA only in a pair); epilogue thread (quadrant, lane) owns one query row: it reads
// the row's scores with tcgen05.ld (32 columns at a time), max-reduces them (FMNMX3) and only when the maximum reaches
// the row's threshold walks a bit-mask of the 32 compares, appending survivors with a private cursor into the
// segment this persistent unit owns for that query (no atomics; bursts take an out-of-line branch-free append).  A
// tile adds at most TILE_G keys to a segment; after the tile's accumulator has been handed back, a segment that could
// not take another whole tile is pruned in place by its warp (ern_compact.cuh), which also tightens the query's
// threshold for every unit (see CandidateSink): a launch is exact for any gallery order and any number of rows.
// Work items are (super tile of `tiles_per_item` gallery tiles, query tile), super-tile-major, dealt round-robin to
// the persistent units: the query tiles of a batch sweep the same few super tiles at the same time and share them
// through L2.  The epilogue is deliberately compact code (rolled chunk loop).
// Environment knobs (profiling aids, read once): ERN_FORCE_SINGLE_CTA=1, ERN_PREFETCH_TILES=n, ERN_TILES_PER_ITEM=n,
// ERN_DEBUG_FLAGS, ERN_TRACE_PTR.
#include "ern_common.cuh"
#include "ern_compact.cuh"
#include "ern_ptx.cuh"

namespace ern {
namespace simtc {

constexpr int kBlockQ = 128;
constexpr int kBlockG = 128;
constexpr int kBlockK = 64;
constexpr int kMaxKBlocks = 12;      // D <= 768 resident (the reference uses 512 and 640; 768 = ViT-L/14 leaves a 2-stage ring)
constexpr int kTotalTiles = 14;      // 16 KB smem tiles: num_kblocks hold the query tile, the rest is the gallery ring
constexpr int kMaxStages = 12;       // (D = 768 -> 2 stages, D = 640 -> 4, D = 512 -> 6, D <= 128 -> 12)
constexpr int kTileBytes = kBlockQ * kBlockK * 2;  // 16 KB: one 128 x 64 bf16 swizzled tile
constexpr int kThreads = 192;
// Warp roles.  The SM's issue arbiter favours the higher warp id within a sub-partition (wid % 4), so the two
// latency-critical single-thread roles get the highest ids and the four epilogue warps (quadrant = wid) the lowest.
constexpr int kProducerWarp = 4;
constexpr int kMmaWarp = 5;
constexpr int kSmemBytes = 1024 + kTotalTiles * kTileBytes + 512;
constexpr int kDenseAppendFrom = 7;  // survivors in one 32-column chunk above which the unrolled append path is taken

struct Params {
  CandidateSink sink;
  int num_kblocks;
  int n_qtiles;         // query tiles of (kPair ? 256 : 128) rows
  int tiles_total;      // gallery tiles of TILE_G rows in [row_begin, row_end)
  int tiles_per_item;   // gallery tiles per work item (super tile)
  int n_super;          // super tiles
  int prefetch_tiles;   // how many gallery tiles ahead of the TMA loads the L2 prefetch runs (0 = off)
  int f16_operands;     // 1: queries and gallery hold fp16 instead of bf16 (same layout, same MMA rate)
  int debug;            // profiling aid (ERN_DEBUG_FLAGS): 1 = never take the append path, 2 = skip the TMEM reads
  unsigned long long* trace;   // profiling aid (ERN_TRACE_PTR): per unit 8 counters, see tools/trace_sim.py; null = off
};

// One 32-column chunk of scores, passed BY VALUE to the out-of-line burst path so that the hot loop's register
// allocation does not see it (the copy to the callee's frame only happens on the rare path).
struct Chunk32 {
  uint32_t v[32];
};
template <int kRankBy>
static __device__ __noinline__ int dense_append(uint64_t* seg, int cnt, uint32_t mask, Chunk32 ch, int64_t row0,
                                                int64_t row_end, int64_t id_offset, int32_t excl) {
  // rows past the end of the range and the excluded id leave the mask first, then every survivor's slot follows from a
  // prefix popcount of the mask: 32 independent predicated stores, no cursor dependency chain (this warp is alone on
  // its scheduler, so a serial chain would run at instruction latency, not issue rate)
  const int64_t left = row_end - row0;                       // rows of this chunk inside the range (ragged last tile)
  if (left <= 0) mask = 0u;
  else if (left < 32) mask &= (1u << left) - 1u;
  const int64_t ex = static_cast<int64_t>(excl) - id_offset - row0;
  if (excl >= 0 && ex >= 0 && ex < 32) mask &= ~(1u << ex);
  const uint32_t gid0 = static_cast<uint32_t>(row0 + id_offset);
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if ((mask >> j) & 1u)
      seg[cnt + __popc(mask & ((1u << j) - 1u))] = make_key(rank_value<kRankBy>(__uint_as_float(ch.v[j])), gid0 + j);
  }
  return cnt + __popc(mask);
}

struct Barriers {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t a_full;
  uint64_t a_empty;
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

template <bool kPair, int kRankBy>
__global__ void __launch_bounds__(kThreads, 1)
sim_topk_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_g,
                   const Params p) {
  constexpr int kCta = kPair ? 2 : 1;
  constexpr int kTileG = kPair ? 256 : 128;   // gallery rows per accumulator tile
  constexpr int kAccCols = kTileG;            // fp32 accumulator columns per stage
  constexpr int kTmemCols = 2 * kAccCols;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_base + p.num_kblocks * kTileBytes;
  const uint32_t kStages = min(kTotalTiles - p.num_kblocks, kMaxStages);   // deeper gallery ring for smaller D
  Barriers* bars = reinterpret_cast<Barriers*>(smem_raw + (smem_base + kTotalTiles * kTileBytes - ptx::smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = kPair ? ptx::cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const int unit = kPair ? (blockIdx.x >> 1) : blockIdx.x;
  const int n_units = kPair ? (gridDim.x >> 1) : gridDim.x;
  const int n_items = p.n_qtiles * p.n_super;
  int32_t* status = p.sink.status;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxStages; ++s) {
      ptx::mbar_init(ptx::smem_u32(&bars->full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars->empty[s]), 1);
    }
    ptx::mbar_init(ptx::smem_u32(&bars->a_full), 1);
    ptx::mbar_init(ptx::smem_u32(&bars->a_empty), 1);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(ptx::smem_u32(&bars->tmem_full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars->tmem_empty[s]), 4 * kCta);  // one arrival per epilogue warp
    }
    ptx::fence_barrier_init();
    ptx::fence_proxy_async();
  }
  if (warp == kProducerWarp && lane == 0) {
    ptx::prefetch_tensormap(&tmap_q);
    ptx::prefetch_tensormap(&tmap_g);
  }
  if (warp == kMmaWarp) ptx::tmem_alloc<kCta>(ptx::smem_u32(&bars->tmem_base), kTmemCols);
  ptx::tc_fence_before();
  if (kPair) ptx::cluster_sync_all(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == kProducerWarp) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      const uint32_t a_full_dst = kPair ? ptx::mapa(ptx::smem_u32(&bars->a_full), 0) : ptx::smem_u32(&bars->a_full);
      uint32_t stage = 0, phase = 0, it = 0;
      for (int item = unit; item < n_items; item += n_units, ++it) {
        const int qt = item % p.n_qtiles;
        const int st = item / p.n_qtiles;
        const int q_row = qt * (kBlockQ * kCta) + rank * kBlockQ;
        // A buffer is free once every MMA of the previous item has retired
        ptx::mbar_wait(ptx::smem_u32(&bars->a_empty), (it & 1) ^ 1, status, 1);
        if (leader) ptx::mbar_arrive_expect_tx(ptx::smem_u32(&bars->a_full), p.num_kblocks * kTileBytes * kCta);
        for (int kb = 0; kb < p.num_kblocks; ++kb) {
          if (kPair) ptx::tma_load_2d_pair(smem_a + kb * kTileBytes, &tmap_q, kb * kBlockK, q_row, a_full_dst);
          else       ptx::tma_load_2d(smem_a + kb * kTileBytes, &tmap_q, kb * kBlockK, q_row, a_full_dst);
        }
        const int t0 = st * p.tiles_per_item;
        const int t1 = min(t0 + p.tiles_per_item, p.tiles_total);
        for (int t = t0; t < t1; ++t) {
          const int g_row = static_cast<int>(p.sink.row_begin) + t * kTileG + rank * kBlockG;
          // L2 prefetch of the gallery tile `prefetch_tiles` ahead, box by box in front of the matching loads (one
          // unit per tile, rotating over the query tiles that share it): more bytes in flight than the 4-stage ring holds
          const bool pf_on = p.prefetch_tiles > 0 && ((t + p.prefetch_tiles) % p.n_qtiles) == qt && t + p.prefetch_tiles < t1;
          const int pf_row = g_row + p.prefetch_tiles * kTileG;
          for (int kb = 0; kb < p.num_kblocks; ++kb) {
            ptx::mbar_wait(ptx::smem_u32(&bars->empty[stage]), phase ^ 1, status, 2);
            const uint32_t full = ptx::smem_u32(&bars->full[stage]);
            if (leader) ptx::mbar_arrive_expect_tx(full, kTileBytes * kCta);
            if (kPair) ptx::tma_load_2d_pair(smem_b + stage * kTileBytes, &tmap_g, kb * kBlockK, g_row, ptx::mapa(full, 0));
            else       ptx::tma_load_2d(smem_b + stage * kTileBytes, &tmap_g, kb * kBlockK, g_row, full);
            if (pf_on) ptx::tma_prefetch_2d(&tmap_g, kb * kBlockK, pf_row);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    // ================================ MMA issuer ================================
    if (leader && lane == 0) {
      // (A / B format fields of the instruction descriptor: 1 = bf16, 0 = fp16)
      const uint32_t idesc = ptx::idesc_bf16_f32(kBlockQ * kCta, kTileG) & (p.f16_operands ? ~((1u << 7) | (1u << 10)) : ~0u);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0, it = 0;
      // cycle accounting of the issuing thread (ERN_TRACE_PTR): total cycles and tiles cost nothing inside the loop;
      // the per-wait split perturbs the issue loop by ~8 % and is compiled in only with -DERN_SIM_TRACE_WAITS
      const bool tr = p.trace != nullptr;
      const long long c_begin = tr ? clock64() : 0;
      long long n_tiles = 0;
#ifdef ERN_SIM_TRACE_WAITS
      long long w_full = 0, w_acc = 0, w_a = 0, c0 = 0;
#define ERN_TW(x) x
#else
#define ERN_TW(x)
#endif
      for (int item = unit; item < n_items; item += n_units, ++it) {
        const int st = item / p.n_qtiles;
        ERN_TW(c0 = clock64();)
        ptx::mbar_wait(ptx::smem_u32(&bars->a_full), it & 1, status, 3);
        ERN_TW(w_a += clock64() - c0;)
        ptx::tc_fence_after();
        const int t0 = st * p.tiles_per_item;
        const int t1 = min(t0 + p.tiles_per_item, p.tiles_total);
        n_tiles += t1 - t0;
        for (int t = t0; t < t1; ++t) {
          ERN_TW(c0 = clock64();)
          ptx::mbar_wait(ptx::smem_u32(&bars->tmem_empty[acc]), acc_phase ^ 1, status, 4);
          ERN_TW(w_acc += clock64() - c0;)
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * kAccCols;
          for (int kb = 0; kb < p.num_kblocks; ++kb) {
            ERN_TW(c0 = clock64();)
            ptx::mbar_wait(ptx::smem_u32(&bars->full[stage]), phase, status, 5);
            ERN_TW(w_full += clock64() - c0;)
            ptx::tc_fence_after();
            const uint64_t adesc = ptx::smem_desc_sw128(smem_a + kb * kTileBytes);
            const uint64_t bdesc = ptx::smem_desc_sw128(smem_b + stage * kTileBytes);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {
              // advancing 16 bf16 (32 B) along K inside the 128 B swizzle atom = +2 in the address field
              ptx::umma_bf16<kCta>(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            }
            ptx::umma_commit<kCta>(ptx::smem_u32(&bars->empty[stage]));
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
          ptx::umma_commit<kCta>(ptx::smem_u32(&bars->tmem_full[acc]));
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        ptx::umma_commit<kCta>(ptx::smem_u32(&bars->a_empty));
      }
      if (tr) {
        unsigned long long* o = p.trace + static_cast<size_t>(unit) * 8;
        o[0] += static_cast<unsigned long long>(clock64() - c_begin);
        o[1] += static_cast<unsigned long long>(n_tiles);
#ifdef ERN_SIM_TRACE_WAITS
        o[2] += static_cast<unsigned long long>(w_full);
        o[3] += static_cast<unsigned long long>(w_acc);
        o[4] += static_cast<unsigned long long>(w_a);
#endif
      }
    }
    __syncwarp();
  } else {
    // ================================ epilogue ================================
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may read
    const int row_in_tile = quad * 32 + lane;
    const uint32_t empty_dst_base = ptx::smem_u32(&bars->tmem_empty[0]);
    uint32_t acc = 0, acc_phase = 0;
    const CandidateSink& sink = p.sink;
#ifdef ERN_SIM_TRACE_WAITS
    const bool etr = p.trace != nullptr && leader && warp == 0 && lane == 0;   // one epilogue warp of the leader CTA
    long long e_begin = etr ? clock64() : 0, e_wait = 0, ec0 = 0, ec1 = 0, e_comp = 0, n_comp = 0, n_t4 = 0, n_t8 = 0, n_t16 = 0, e_tmax = 0;
#endif
    for (int item = unit; item < n_items; item += n_units) {
      const int qt = item % p.n_qtiles;
      const int st = item / p.n_qtiles;
      const int64_t q = static_cast<int64_t>(qt) * (kBlockQ * kCta) + rank * kBlockQ + row_in_tile;
      const bool q_ok = q < sink.nq;
      const bool dense = sink.dense != 0;
      const bool live = q_ok && !dense;
      const int32_t excl = (q_ok && sink.exclude) ? sink.exclude[q] : -1;
      // this thread is the only writer of segment `unit` of query q: private cursor, plain stores
      const int64_t qs = live ? q : 0;
      uint64_t* seg = sink.segs + (qs * sink.n_seg + unit) * sink.seg_cap;
      uint32_t* thr_ord_q = sink.thr_ord + qs;
      int32_t* cnt_q = sink.seg_counts + qs * sink.n_seg + unit;
      int cnt = live ? *cnt_q : 0;
      float thr = INFINITY;                      // rows beyond the batch never append
      // A tile appends at most kTileG keys to a segment, so a segment that holds no more than `prune_at` keys when a
      // tile starts cannot overflow; it is pruned (warp_compact_segment) AFTER the tile's accumulator has been handed
      // back, i.e. in the time this warp would otherwise spend waiting for the next tile's MMAs.
      const int prune_at = sink.seg_cap - kTileG;
      const int keep_max = sink.k + (prune_at - sink.k) / 4;   // a pruning pass leaves between k and keep_max keys
      const int t0 = st * p.tiles_per_item;
      const int t1 = min(t0 + p.tiles_per_item, p.tiles_total);
      for (int t = t0; t < t1; ++t) {
        // the query's lower bound may have been tightened by any unit since the last tile (stale is harmless)
        if (live) {
          const float g = ordered_to_f32(__ldcg(thr_ord_q));
          thr = (t == t0) ? g : fmaxf(thr, g);
        }
        ERN_TW(if (etr) ec0 = clock64();)
        ptx::mbar_wait(ptx::smem_u32(&bars->tmem_full[acc]), acc_phase, status, 6);
        ERN_TW(if (etr) { e_wait += clock64() - ec0; ec1 = clock64(); })
        ptx::tc_fence_after();
        const int64_t g_base = sink.row_begin + static_cast<int64_t>(t) * kTileG;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kAccCols;
        if (!(p.debug & 2)) {
          // The chunk loop is deliberately NOT unrolled and the survivor path is a bit-mask walk: the whole
          // epilogue stays a few KB of code.  (A fully unrolled version was 250 KB and ran from L2 instruction
          // fetches: ncu showed stall_no_inst on every epilogue instruction and 6% tensor-pipe activity.)
#pragma unroll 1
          for (int c = 0; c < kAccCols / 32; ++c) {
            uint32_t v[32];
            ptx::tmem_ld_32x32(taddr + c * 32, v);
            ptx::tmem_ld_wait();
            const int64_t row0 = g_base + c * 32;
            if (dense) {
              // every row is kept (NaN scores / the excluded id become empty slots)
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if (row0 + j < sink.row_end && q_ok)
                  sink_put_dense(sink, q, row0 + j, rank_value<kRankBy>(__uint_as_float(v[j])), excl);
              }
            } else {
              float m = __uint_as_float(v[0]);
#pragma unroll
              for (int j = 1; j < 32; ++j) m = fmaxf(m, __uint_as_float(v[j]));
              if (rank_value<kRankBy>(m) >= thr && !(p.debug & 1)) {
                uint32_t mask = 0;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  mask |= (rank_value<kRankBy>(__uint_as_float(v[j])) >= thr ? 1u : 0u) << j;
                if (__popc(mask) > kDenseAppendFrom) {
                  // a burst (ordered gallery: a cluster of rows close to this query): out-of-line straight-line
                  // appends, ~8 instructions per column instead of the ~50 of a select-tree walk step
                  Chunk32 ch;
#pragma unroll
                  for (int j = 0; j < 32; ++j) ch.v[j] = v[j];
                  cnt = dense_append<kRankBy>(seg, cnt, mask, ch, row0, sink.row_end, sink.id_offset, excl);
                  mask = 0;
                }
                while (mask) {
                  const int j = __ffs(mask) - 1;
                  mask &= mask - 1;
                  // register file has no dynamic indexing: 5-level select tree picks v[j]
                  uint32_t s16[16], s8[8], s4[4], s2[2];
#pragma unroll
                  for (int i = 0; i < 16; ++i) s16[i] = (j & 16) ? v[i + 16] : v[i];
#pragma unroll
                  for (int i = 0; i < 8; ++i) s8[i] = (j & 8) ? s16[i + 8] : s16[i];
#pragma unroll
                  for (int i = 0; i < 4; ++i) s4[i] = (j & 4) ? s8[i + 4] : s8[i];
#pragma unroll
                  for (int i = 0; i < 2; ++i) s2[i] = (j & 2) ? s4[i + 2] : s4[i];
                  const float r = rank_value<kRankBy>(__uint_as_float((j & 1) ? s2[1] : s2[0]));
                  const int64_t row = row0 + j;
                  const uint32_t gid = static_cast<uint32_t>(row + sink.id_offset);
                  if (row < sink.row_end && static_cast<int32_t>(gid) != excl) seg[cnt++] = make_key(r, gid);
                }
              }
            }
          }
        }
        // accumulator stage drained: hand it back to the MMA issuer (leader CTA's barrier)
        ptx::tc_fence_before();
        __syncwarp();
        ERN_TW(if (etr) { const long long d = clock64() - ec1; if (d > 4000) ++n_t4; if (d > 8000) ++n_t8; if (d > 16000) ++n_t16; if (d > e_tmax) e_tmax = d; })
        if (lane == 0) {
          const uint32_t dst = empty_dst_base + acc * 8;
          if (kPair) ptx::mbar_arrive_cluster(ptx::mapa(dst, 0)); else ptx::mbar_arrive(dst);
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        // segments that could not take another whole tile are pruned to their best keys now
        const bool full = cnt > prune_at;
        if (__any_sync(0xffffffffu, full)) {
          ERN_TW(if (etr) ec0 = clock64();)
          const CompactResult cr = warp_compact_segment(seg, cnt, thr, thr_ord_q, full, sink.k, keep_max);
          cnt = cr.cnt;
          thr = cr.thr;
          ERN_TW(if (etr) { e_comp += clock64() - ec0; ++n_comp; })
        }
      }
      // publish how many candidates this unit holds for the query
      if (live) *cnt_q = cnt;
    }
#ifdef ERN_SIM_TRACE_WAITS
    if (etr) {
      unsigned long long* o = p.trace + static_cast<size_t>(unit) * 8;
      o[5] += static_cast<unsigned long long>(clock64() - e_begin);
      o[6] += static_cast<unsigned long long>(e_wait);
      o[7] += static_cast<unsigned long long>(e_comp);
      o[4] += static_cast<unsigned long long>(n_comp) << 40;      // (shares the slot with the query-tile wait cycles)
      unsigned long long* o2 = p.trace + static_cast<size_t>(148 + unit) * 8;   // second block: tile-time tail of this warp
      o2[0] += n_t4; o2[1] += n_t8; o2[2] += n_t16;
      if (static_cast<unsigned long long>(e_tmax) > o2[3]) o2[3] = e_tmax;
    }
#endif
  }

  // ================================ teardown ================================
  ptx::tc_fence_before();
  if (kPair) ptx::cluster_sync_all(); else __syncthreads();
  if (warp == kMmaWarp) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<kCta>(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// row-major [rows, dim] bf16 matrix -> boxes of 128 rows x 64 columns, 128-byte swizzle, zero OOB fill
int make_tmap_bf16_rows(CUtensorMap* map, const void* base, int64_t rows, int dim, int64_t ld_elems) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return ERN_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(dim), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld_elems) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kBlockK), static_cast<cuuint32_t>(kBlockQ)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld dim=%d ld=%lld base=%p)", (int)r,
              (long long)rows, dim, (long long)ld_elems, base);
    return ERN_ERR_CUDA;
  }
  return ERN_OK;
}

template <bool kPair, int kRankBy>
static int launch_one(const CUtensorMap& tq, const CUtensorMap& tg, const Params& p, int grid, cudaStream_t st) {
  auto kern = sim_topk_tc_kernel<kPair, kRankBy>;
  // the opt-in shared-memory size is a per-device function attribute
  static std::atomic<bool> configured[64];   // zero-initialised; idempotent per-device attribute set
  int dev = 0;
  ERN_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
    ERN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    if (dev >= 0 && dev < 64) configured[dev].store(true, std::memory_order_release);
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kPair ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ERN_CUDA(cudaLaunchKernelEx(&cfg, kern, tq, tg, p));
  return ERN_OK;
}

// Tiles per work item.  Items = query tiles x super tiles, dealt round-robin to the units; every item starts by
// (re)loading its query tile, which costs about 0.8 tile times.  The planner picks the length (1..64 tiles) with the
// smallest modelled makespan  rounds * (0.8 + tiles_per_item)  where rounds = ceil(items / units): long launches end up
// at 64 (many short rounds, small tail), short launches (the first steps of the schedule, small shards) get items long
// enough that one or two rounds cover the range instead of a dozen items that each reload the query tile for a single
// tile (rows [2k, 16k) of a 4096-query batch: 108 -> ~60 us).
constexpr int kMaxTilesPerItem = 64;
static int plan_tiles_per_item(int n_qtiles, int tiles_total, int units) {
  static const int forced = [] { const char* e = getenv("ERN_TILES_PER_ITEM"); return e ? atoi(e) : 0; }();
  if (forced > 0) return forced;
  int best = 1;
  double best_cost = 1e30;
  for (int tpi = 1; tpi <= kMaxTilesPerItem && tpi <= tiles_total; ++tpi) {
    const long items = static_cast<long>(n_qtiles) * ((tiles_total + tpi - 1) / tpi);
    const long rounds = (items + units - 1) / units;
    const double cost = static_cast<double>(rounds) * (0.8 + tpi);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = tpi; }
  }
  return best;
}

// persistent units (= candidate segments per query) the scoring kernel runs with for a batch of nq queries
int units_for(int64_t nq, int force_single, int sm_count) {
  const bool pair = !force_single && nq > kBlockQ;
  return pair ? sm_count / 2 : sm_count;
}

// One launch of the tensor-core scoring kernel over shard rows [sink.row_begin, sink.row_end).
int launch(const CUtensorMap& tq, const CUtensorMap& tg, const CandidateSink& sink, int dim, int rank_by,
           int force_single, int sm_count, bool f16_operands, cudaStream_t st) {
  const bool pair = !force_single && sink.nq > kBlockQ;
  const int tile_g = pair ? 256 : 128;
  Params p;
  static const int dbg = [] { const char* e = getenv("ERN_DEBUG_FLAGS"); return e ? atoi(e) : 0; }();
  p.debug = dbg;
  p.f16_operands = f16_operands ? 1 : 0;
  // L2 prefetch distance in gallery tiles.  Default: 1 when a gallery tile has a single consumer (<= 256 queries: the
  // HBM-bound / ridge regime, where the 4-stage ring alone keeps too few bytes in flight: 2.43 -> 2.35 ms at 128
  // queries x 10M rows, 2.87 -> 2.79 ms at 256), 0 otherwise (no effect at >= 512 queries).  2 or more re-fetches
  // lines that left L2 again (measured: 2.86 / 3.44 ms at distance 2 / 4).  ERN_PREFETCH_TILES overrides.
  static const int pf_env = [] { const char* e = getenv("ERN_PREFETCH_TILES"); return e ? atoi(e) : -1; }();
  const int pf = pf_env >= 0 ? pf_env : (sink.nq <= (pair ? 2 * kBlockQ : kBlockQ) ? 1 : 0);
  p.prefetch_tiles = pf;
  // ERN_TRACE_PTR: device address of a zeroed u64[8 * units] buffer owned by the profiling tool (tools/trace_sim.py)
  static unsigned long long* const trace = [] {
    const char* e = getenv("ERN_TRACE_PTR");
    return e ? reinterpret_cast<unsigned long long*>(strtoull(e, nullptr, 0)) : nullptr;
  }();
  p.trace = trace;
  p.num_kblocks = dim / kBlockK;
  p.n_qtiles = cdiv(sink.nq, pair ? 2 * kBlockQ : kBlockQ);
  p.tiles_total = cdiv(sink.row_end - sink.row_begin, tile_g);
  if (p.tiles_total <= 0) return ERN_OK;
  int units = units_for(sink.nq, force_single, sm_count);
  ERN_REQUIRE(units <= sink.n_seg, "internal: %d scoring units but %d candidate segments per query", units, sink.n_seg);
  p.tiles_per_item = plan_tiles_per_item(p.n_qtiles, p.tiles_total, units);
  p.n_super = cdiv(p.tiles_total, p.tiles_per_item);
  p.sink = sink;
  const long items = static_cast<long>(p.n_qtiles) * p.n_super;
  if (items < units) units = static_cast<int>(items);
  const int grid = pair ? 2 * units : units;
  if (pair) {
    return rank_by == ERN_RANK_REFERENCE ? launch_one<true, ERN_RANK_REFERENCE>(tq, tg, p, grid, st)
                                         : launch_one<true, ERN_RANK_SIMILARITY>(tq, tg, p, grid, st);
  }
  return rank_by == ERN_RANK_REFERENCE ? launch_one<false, ERN_RANK_REFERENCE>(tq, tg, p, grid, st)
                                       : launch_one<false, ERN_RANK_SIMILARITY>(tq, tg, p, grid, st);
}

}  // namespace simtc
}  // namespace ern
