// Fusion head (CombinerSimple.forward, models/fusion_model.py:86-94) for SMALL batches: the reference's query side
// calls it with 32 rows (run/test/test_fiq.py:132) three times per batch (models/fusion_model.py:52-54).
//
// At <= 64 rows the head is weight-bandwidth bound (59 MB of bf16 weights for 1.9 GFLOP at 32 rows: the floor is the
// time HBM needs to deliver the weights once, ~9 us), so the job is to keep every byte of the weights in flight on as
// many SMs as it takes to saturate HBM and to touch nothing else:
//   phase A   raw[B, 8D]  = relu([text | image] . [Wt ; Wi]^T + [bt ; bi])      fp32 inputs converted on the fly
//   phase B   partial[B, c] = sum over the CTA's columns of relu(raw . W1^T + b1) * w2     (h[B, 8D] never exists)
//   phase C   gate sigmoid, blend from the fp32 inputs, L2 normalise (CTA r handles row r after a second barrier)
// ONE cooperative launch: the <= 148 CTAs are co-resident, grid barriers (atomic counters behind the packed weights,
// one of 15 sets per call, left at zero by the last CTA) separate the phases -- no launch gaps, no host round trips.
// Both GEMMs split the OUTPUT COLUMNS over the CTAs (32-48 columns each, so that one wave of <= 148 CTAs covers the
// matrix) and the K range over the 8 warps of a CTA: every weight byte is read exactly once from HBM by exactly one
// warp with 16-byte streaming loads, several K blocks in flight per warp (64-100 KB per SM); no cross-CTA reduction, no
// split-K pass.  The math is bf16 x bf16 -> fp32 on the warp-level tensor path (mma.sync.m16n8k16): a 32 x 40 x 5120
// problem per CTA cannot fill a 128-row tcgen05 tile and is three orders of magnitude below the tensor roofline
// anyway.  Threads load 8 consecutive K elements of their row for both operands and use them as the fragments of two
// k16 steps: a dot product does not care in which order K is walked as long as A and B agree, and this keeps every
// global load a full 16-byte vector.
#include "ern_internal.cuh"

namespace ern {
namespace combiner {
namespace small {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// 8 consecutive K elements of one row -> four bf16x2 registers
__device__ __forceinline__ void load8_act(const __nv_bfloat16* p, bool ok, uint32_t (&r)[4]) {
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  if (ok) v = __ldcg(reinterpret_cast<const uint4*>(p));    // `raw` was written by other SMs in this launch: L2, not L1
  r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
}
__device__ __forceinline__ void load8_act(const float* p, bool ok, uint32_t (&r)[4]) {
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
  if (ok) {
    a = __ldg(reinterpret_cast<const float4*>(p));
    b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  }
  __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
  __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
  r[0] = *reinterpret_cast<uint32_t*>(&p0); r[1] = *reinterpret_cast<uint32_t*>(&p1);
  r[2] = *reinterpret_cast<uint32_t*>(&p2); r[3] = *reinterpret_cast<uint32_t*>(&p3);
}
// Weights stream HBM -> shared memory with 16-byte cp.async copies, kStages K blocks ahead of their use.  Every thread
// reads back exactly the bytes it copied, so the ring is a per-thread FIFO: no block barrier, only cp.async.wait_group.
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory"); }
__device__ __forceinline__ void lds128(uint32_t smem_addr, uint32_t (&r)[4]) {
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_addr) : "memory");
}
// K blocks in flight per thread: what fits into kRingBudget bytes of shared memory (a slot is 16 bytes per thread; a K
// block needs kNTiles weight slots and, for bf16 activations, 2 * kMTiles activation slots)
// (more is not better: with a 220 KB budget -- 6 instead of 5 stages at 32 rows, 4 instead of 3 at 64 -- every size got
// slower, 26.3 -> 27.6 us at 32 rows; profiles/r02_head_micro_ab.jsonl)
constexpr int kRingBudget = 190 * 1024;
constexpr int stages_for(int m_tiles, int n_tiles) {
  const int per_stage = (n_tiles + 2 * m_tiles) * 256 * 16;
  const int s = kRingBudget / per_stage;
  return s > 8 ? 8 : s;
}
// Early bias fetch + one fence per CTA, for every batch size.  (An A/B build that kept the old form above 32 rows was
// 2.2 us slower at 64 rows once the ring was 128-byte aligned: profiles/r02_head_align_ab.jsonl.  An earlier
// measurement that said the opposite had the ring misaligned, see fused_head_kernel.)
constexpr bool lean_form(int) { return true; }
// static shared memory of fused_head_kernel: tile[16 M][8 N + 1] + sred[40]
constexpr int static_smem_for(int m_tiles, int n_tiles) { return ((16 * m_tiles * (8 * n_tiles + 1) + 40) * 4 + 127) / 128 * 128; }

// C[rows, n0 .. n0 + 8*kNTiles) = relu(X . W^T + bias) for this CTA's column block.
//   X  : [rows, K] (TIn = float: converted to bf16 on the fly; or bf16), columns < n_split read x0, the others x1
//   W  : [N, K] bf16, K contiguous
//   !kGate: C rounded to bf16 -> out[rows, ldo]
//    kGate: partial[r, blockIdx.x] = sum_n C[r, n] * wg[n]
template <typename TIn, int kMTiles, int kNTiles, bool kGate, typename AfterPrologue>
__device__ __forceinline__ void gemm_phase(AfterPrologue after_prologue, const TIn* __restrict__ x0, const TIn* __restrict__ x1, int64_t ldx, int rows,
                                           int K, const __nv_bfloat16* __restrict__ W, int n_split,
                                           const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int64_t ldo,
                                           const float* __restrict__ wg, float* __restrict__ partial, int n_ctas,
                                           float (*tile)[8 * kNTiles + 1], uint32_t wring, float* ring_f32,
                                           uint32_t xtile, uint32_t xpitch) {
  constexpr int kCols = 8 * kNTiles;
  constexpr int kRows = 16 * kMTiles;
  constexpr int kAcc = kMTiles * kNTiles * 4;                 // fp32 accumulators per thread

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int n0 = blockIdx.x * kCols;
  const TIn* __restrict__ x = (n0 < n_split) ? x0 : x1;
  // K is split over as many warps as divide it into whole 32-element blocks (8 for K = 8D, 4 for K = 640, ...)
  const int blocks_total = K / 32;
  const int ks = (blocks_total % 8 == 0) ? 8 : (blocks_total % 4 == 0) ? 4 : (blocks_total % 2 == 0) ? 2 : 1;
  const int n_blocks = warp < ks ? blocks_total / ks : 0;
  const int k_begin = warp < ks ? warp * n_blocks * 32 : 0;

  float acc[kMTiles][kNTiles][4];
#pragma unroll
  for (int i = 0; i < kMTiles; ++i)
#pragma unroll
    for (int j = 0; j < kNTiles; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[i][j][c] = 0.f;

  const __nv_bfloat16* wrow[kNTiles];
#pragma unroll
  for (int j = 0; j < kNTiles; ++j) wrow[j] = W + static_cast<int64_t>(n0 + 8 * j + g) * K + k_begin + 8 * t;
  const TIn* xlo[kMTiles];
  const TIn* xhi[kMTiles];
  bool oklo[kMTiles], okhi[kMTiles];
#pragma unroll
  for (int i = 0; i < kMTiles; ++i) {
    const int rlo = 16 * i + g, rhi = rlo + 8;
    oklo[i] = rlo < rows;
    okhi[i] = rhi < rows;
    xlo[i] = x + static_cast<int64_t>(oklo[i] ? rlo : 0) * ldx + k_begin + 8 * t;
    xhi[i] = x + static_cast<int64_t>(okhi[i] ? rhi : 0) * ldx + k_begin + 8 * t;
  }

  // this thread's FIFO: kStages stages of kSlots x 16 bytes, slot-major with the 256 threads interleaved (conflict-free).
  // bf16 activations (`raw`, L2-resident, written by other SMs earlier in this launch) ride in the same FIFO; fp32
  // activations (phase A: K = D only) are loaded directly.
  constexpr bool kActFifo = sizeof(TIn) == 2;
  constexpr int kSlots = kNTiles + (kActFifo ? 2 * kMTiles : 0);
  constexpr int kStages = stages_for(kMTiles, kNTiles);
  const uint32_t fifo = wring + threadIdx.x * 16u;
  auto slot_addr = [&](int stage, int j) { return fifo + static_cast<uint32_t>((stage * kSlots + j) * kThreads * 16); };
  // CTAs run in lock-step: each starts its K walk at a different block so that they do not all ask the same few
  // memory channels for the same offsets at the same time (a sum does not care where the walk starts)
  const int rot = n_blocks > 0 ? static_cast<int>(blockIdx.x) % n_blocks : 0;
  auto kblock = [&](int kb) { const int b = kb + rot; return b >= n_blocks ? b - n_blocks : b; };
  auto issue = [&](int kb) {
    if (kb < n_blocks) {
      const int stage = kb % kStages;
      const int off = kblock(kb) * 32;
#pragma unroll
      for (int j = 0; j < kNTiles; ++j) cp_async16(slot_addr(stage, j), wrow[j] + off);
      if (kActFifo) {
#pragma unroll
        for (int i = 0; i < kMTiles; ++i) {
          if (oklo[i]) cp_async16(slot_addr(stage, kNTiles + 2 * i), xlo[i] + off);
          if (okhi[i]) cp_async16(slot_addr(stage, kNTiles + 2 * i + 1), xhi[i] + off);
        }
      }
    }
    cp_async_commit();                                        // (an empty group keeps the wait_group arithmetic uniform)
  };
  // The epilogue needs this CTA's kCols entries of the bias (and gate) vector only after the K loop; from a cold L2 that
  // would be an HBM round trip at the end of the phase (ncu: 8 % of the kernel's stall samples sat on the two FADDs that
  // consume them).  Thread c < kCols fetches entry c now and parks it in a register across the loop.
  constexpr bool kPreload = lean_form(kMTiles);
  float pre_bias = 0.f, pre_gate = 0.f;
  if (kPreload && threadIdx.x < kCols) {
    pre_bias = __ldg(bias + n0 + threadIdx.x);
    if (kGate) pre_gate = __ldg(wg + n0 + threadIdx.x);
  }
#pragma unroll
  for (int kb = 0; kb < kStages - 1; ++kb) issue(kb);
  after_prologue();                                           // (phase A: the input tile is built while the weights fly)
  for (int kb = 0; kb < n_blocks; ++kb) {
    issue(kb + kStages - 1);
    uint32_t ralo[kMTiles][4], rahi[kMTiles][4];
    if (!kActFifo) {
      // fp32 activations were converted into a bf16 tile in shared memory before the phase started (rows beyond the
      // batch are zero there)
      const uint32_t kofs = static_cast<uint32_t>(k_begin + 8 * t + kblock(kb) * 32) * 2u;
#pragma unroll
      for (int i = 0; i < kMTiles; ++i) {
        lds128(xtile + static_cast<uint32_t>(16 * i + g) * xpitch + kofs, ralo[i]);
        lds128(xtile + static_cast<uint32_t>(16 * i + g + 8) * xpitch + kofs, rahi[i]);
      }
    }
    cp_async_wait<kStages - 1>();                             // this thread's copies of block kb have landed
    uint32_t rb[kNTiles][4];
    const int stage = kb % kStages;
#pragma unroll
    for (int j = 0; j < kNTiles; ++j) lds128(slot_addr(stage, j), rb[j]);
    if (kActFifo) {
#pragma unroll
      for (int i = 0; i < kMTiles; ++i) {
        if (oklo[i]) lds128(slot_addr(stage, kNTiles + 2 * i), ralo[i]);
        else ralo[i][0] = ralo[i][1] = ralo[i][2] = ralo[i][3] = 0u;
        if (okhi[i]) lds128(slot_addr(stage, kNTiles + 2 * i + 1), rahi[i]);
        else rahi[i][0] = rahi[i][1] = rahi[i][2] = rahi[i][3] = 0u;
      }
    }
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
      for (int i = 0; i < kMTiles; ++i)
#pragma unroll
        for (int j = 0; j < kNTiles; ++j)
          mma_bf16(acc[i][j], ralo[i][2 * s2], rahi[i][2 * s2], ralo[i][2 * s2 + 1], rahi[i][2 * s2 + 1], rb[j][2 * s2],
                   rb[j][2 * s2 + 1]);
  }
  cp_async_wait<0>();

  // ---- sum the warps' partial tiles in a fixed order (deterministic), then the epilogue -----------------------------
  // Every thread's copies have landed and, after this barrier, nobody reads the ring (or phase A's input tile) any more:
  // the ring now holds the ks partial tiles, one per warp, in fragment order [warp][element][lane], and behind them the
  // bias / gate entries fetched before the loop.  After ONE more barrier all 256 threads add the partials up in warp
  // order -- the same order, hence the same bits, as accumulating warp after warp with a barrier in between, which is what
  // this replaced (ks + 1 barriers with one warp working at a time: 1.8 us of the 28.5 at 32 rows).
  __syncthreads();
  float* svec = ring_f32 + kWarps * (kAcc * 32);
  if (warp < ks) {
    float* mine = ring_f32 + warp * (kAcc * 32) + lane;
#pragma unroll
    for (int i = 0; i < kMTiles; ++i)
#pragma unroll
      for (int j = 0; j < kNTiles; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) mine[((i * kNTiles + j) * 4 + c) * 32] = acc[i][j][c];
  }
  if (kPreload && threadIdx.x < kCols) {
    svec[threadIdx.x] = pre_bias;
    if (kGate) svec[kCols + threadIdx.x] = pre_gate;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < kAcc * 32; idx += kThreads) {
    const int e = idx >> 5, l = idx & 31;
    float v = ring_f32[idx];
    for (int w = 1; w < ks; ++w) v += ring_f32[w * (kAcc * 32) + idx];
    const int c = e & 3, j = (e >> 2) % kNTiles, i = (e >> 2) / kNTiles;
    const int row = 16 * i + (l >> 2) + 8 * (c >> 1);
    const int col = 8 * j + 2 * (l & 3) + (c & 1);
    v = fmaxf(v + (kPreload ? svec[col] : bias[n0 + col]), 0.f);
    if (kGate) tile[row][col] = v * (kPreload ? svec[kCols + col] : wg[n0 + col]);
    else if (row < rows) out[static_cast<int64_t>(row) * ldo + n0 + col] = __float2bfloat16_rn(v);
  }
  if (kGate) {
    __syncthreads();
    if (threadIdx.x < kRows && static_cast<int>(threadIdx.x) < rows) {
      float z = 0.f;
#pragma unroll
      for (int col = 0; col < kCols; ++col) z += tile[threadIdx.x][col];
      partial[static_cast<int64_t>(threadIdx.x) * n_ctas + blockIdx.x] = z;
    }
  }
}

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

struct HeadArgs {
  const float* image;
  const float* text;
  int rows, dim;
  PackedView pv;
  __nv_bfloat16* raw;        // [rows, 8D] scratch
  float* partial;            // [rows, gridDim.x] scratch
  unsigned* sync;            // [3] zero between calls (one of 15 sets behind the packed weights)
  float* out;                // [rows, D] or null
  __nv_bfloat16* out_bf16;   // [rows, ldb] or null
  int64_t ldb;
  float* gate;               // [rows] or null
  unsigned long long* stamps;   // profiling aid or null
};

// (profiling aid, ERN_HEAD_STAMP_PTR: CTA 0 writes %globaltimer at the phase boundaries)
#define STAMP(i)                                                                                   \
  do {                                                                                             \
    if (a.stamps && blockIdx.x == 0 && threadIdx.x == 0) {                                         \
      unsigned long long t_;                                                                       \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                       \
      a.stamps[i] = t_;                                                                            \
    }                                                                                              \
  } while (0)

// The whole head in one cooperative launch (all CTAs co-resident: the grid barriers below rely on it).
template <int kMTiles, int kNTiles>
__global__ void __launch_bounds__(kThreads, 1) fused_head_kernel(const HeadArgs a) {
  constexpr int kCols = 8 * kNTiles;
  __shared__ float sred[32 + kWarps];                        // phase C: gate, per-warp sums of squares
  __shared__ float tile[16 * kMTiles][kCols + 1];
  // 128-byte aligned: a warp's 512 contiguous bytes of a slot must be four shared-memory lines, not five.  The static
  // arrays in front of it end at a multiple of 32 only; when `red` (10-20 KB, a multiple of 128 by luck) left the kernel
  // the ring slid to offset 32 mod 128 and the 64-row case went 36.3 -> 40.5 us until this attribute pinned it
  // (34.9 us with it).
  extern __shared__ __align__(128) uint8_t wring_raw[];        // stages_for(M, N) stages of (kNTiles + 2 kMTiles) slots x 256 threads x 16 B
  const uint32_t wring = static_cast<uint32_t>(__cvta_generic_to_shared(wring_raw));
  const int hid = 8 * a.dim, proj = 4 * a.dim;
  const int n_ctas = gridDim.x;

  STAMP(0);
  // ---- phase A: this CTA's columns of raw = relu([text | image] . [Wt ; Wi]^T + [bt ; bi])  (text half first, :90)
  // its input (text for the first half of the column blocks, image for the second) becomes a bf16 tile in shared
  // memory in one round trip: 16 rows x kMTiles, zero beyond the batch, row pitch 2 D + 64 bytes (conflict-free LDS.128)
  const uint32_t xpitch = static_cast<uint32_t>(a.dim) * 2u + 64u;
  const uint32_t xtile = wring + static_cast<uint32_t>(stages_for(kMTiles, kNTiles) * kNTiles * kThreads * 16);
  auto build_input_tile = [&]() {
    const float* __restrict__ x = (static_cast<int>(blockIdx.x) * kCols < proj) ? a.text : a.image;
    const int chunks_per_row = a.dim / 8;
    const int total = 16 * kMTiles * chunks_per_row;
    // twelve 8-element chunks per thread and pass (one pass covers 32 rows x 640): all their loads are in flight before
    // the first conversion
    constexpr int kBatch = 12;
    for (int c0 = threadIdx.x; c0 < total; c0 += kBatch * kThreads) {
      uint32_t v[kBatch][4];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int cc = c0 + u * kThreads;
        const int r = cc / chunks_per_row, c = cc - r * chunks_per_row;
        load8_act(x + static_cast<int64_t>(r < a.rows ? r : 0) * a.dim + c * 8, cc < total && r < a.rows, v[u]);
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int cc = c0 + u * kThreads;
        const int r = cc / chunks_per_row, c = cc - r * chunks_per_row;
        if (cc < total)
          asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(xtile + r * xpitch + c * 16), "r"(v[u][0]),
                       "r"(v[u][1]), "r"(v[u][2]), "r"(v[u][3]) : "memory");
      }
    }
    __syncthreads();
  };
  gemm_phase<float, kMTiles, kNTiles, false>(build_input_tile, a.text, a.image, a.dim, a.rows, a.dim, a.pv.wt, proj, a.pv.bt, a.raw, hid,
                                             nullptr, nullptr, n_ctas, tile, wring, reinterpret_cast<float*>(wring_raw), xtile, xpitch);
  // ---- grid barrier: every column of raw is in L2 before anybody reads a row of it
  // The block barrier orders every thread's stores before thread 0's fence, which publishes them gpu-wide together with
  // the arrival (the cooperative-groups grid.sync pattern) instead of a fence in each of the 256 threads.  Measured
  // together with the early bias fetch: -0.9 us at 1..16 rows, -2.2 us at 64 rows (profiles/r02_head_micro_ab.jsonl,
  // r02_head_align_ab.jsonl).
  constexpr bool kOneFence = lean_form(kMTiles);
  STAMP(1);
  if (!kOneFence) __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (kOneFence) __threadfence();
    atomicAdd(&a.sync[0], 1u);
    while (ld_acquire(&a.sync[0]) < static_cast<unsigned>(n_ctas)) {}
  }
  __syncthreads();
  STAMP(2);
  // ---- phase B: hidden layer + gate dot product over this CTA's columns
  gemm_phase<__nv_bfloat16, kMTiles, kNTiles, true>([] {}, a.raw, a.raw, hid, a.rows, hid, a.pv.w1, hid, a.pv.b1, nullptr, 0,
                                                    a.pv.w2, a.partial, n_ctas, tile, wring, reinterpret_cast<float*>(wring_raw), 0u, 0u);
  // ---- second grid barrier, then phase C: CTA r turns row r's partials into the fused feature row
  STAMP(3);
  if (!kOneFence) __threadfence();
  __syncthreads();
  const bool finisher = static_cast<int>(blockIdx.x) < a.rows;
  if (threadIdx.x == 0) {
    if (kOneFence) __threadfence();
    atomicAdd(&a.sync[1], 1u);
    if (finisher)
      while (ld_acquire(&a.sync[1]) < static_cast<unsigned>(n_ctas)) {}
  }
  __syncthreads();
  STAMP(4);
  if (finisher) {
    const int r = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // the row's inputs do not depend on the gate: fetch them while the partials are summed
    const float* im = a.image + static_cast<int64_t>(r) * a.dim;
    const float* tx = a.text + static_cast<int64_t>(r) * a.dim;
    float vi[4], vt[4];                                       // dim <= 1024: up to 4 elements per thread
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int d = threadIdx.x + e * kThreads;
      vi[e] = d < a.dim ? im[d] : 0.f;
      vt[e] = d < a.dim ? tx[d] : 0.f;
    }
    if (warp == 0) {                                          // same summation order as finalize_kernel (ern_combiner.cu)
      float z = 0.f;
      for (int t = lane; t < n_ctas; t += 32) z += __ldcg(&a.partial[static_cast<int64_t>(r) * n_ctas + t]);
      for (int o = 16; o > 0; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
      if (lane == 0) sred[0] = 1.0f / (1.0f + expf(-(z + a.pv.b2[0])));
    }
    __syncthreads();
    const float s = sred[0];
    float ss = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      vt[e] = s * vt[e] + (1.0f - s) * vi[e];
      ss = fmaf(vt[e], vt[e], ss);
    }
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) sred[32 + warp] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) tot += sred[32 + w];
    const float denom = fmaxf(sqrtf(tot), 1e-12f);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int d = threadIdx.x + e * kThreads;
      if (d < a.dim) {
        const float v = vt[e] / denom;
        if (a.out) a.out[static_cast<int64_t>(r) * a.dim + d] = v;
        if (a.out_bf16) a.out_bf16[static_cast<int64_t>(r) * a.ldb + d] = __float2bfloat16_rn(v);
      }
    }
    if (a.gate && threadIdx.x == 0) a.gate[r] = s;
  }
  STAMP(5);
  // the CTA that leaves last puts the three counters back to zero for the next call that draws this set
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&a.sync[2], 1u) == static_cast<unsigned>(n_ctas) - 1u) {
      a.sync[0] = 0u;
      a.sync[1] = 0u;
      a.sync[2] = 0u;
    }
  }
}
#undef STAMP

// columns per CTA: the smallest of 32 / 40 / 48 that covers N with one wave of CTAs
static int pick_ntiles(int n, int sm_count) {
  // (n / 2 is where the text half of `raw` ends and the image half begins: a column block must not straddle it)
  for (int nt = 4; nt <= 6; ++nt)
    if ((n / 2) % (8 * nt) == 0 && n / (8 * nt) <= sm_count) return nt;
  return 0;
}

int n_partials(int dim, int sm_count) {
  const int nt = pick_ntiles(8 * dim, sm_count);
  return nt ? 8 * dim / (8 * nt) : 0;
}

bool supported(int64_t rows, int dim, int sm_count) {
  // K is walked in whole 32-element blocks; the grid must be one co-resident wave; phase A's bf16 input tile must fit
  // into the part of the shared-memory ring that phase A's weight FIFO leaves free
  if (rows < 1 || rows > 64 || dim % 64 != 0) return false;
  const int nt = pick_ntiles(8 * dim, sm_count);
  if (nt == 0) return false;
  const int mt = rows <= 16 ? 1 : rows <= 32 ? 2 : 4;
  const int free_bytes = stages_for(mt, nt) * 2 * mt * kThreads * 16;
  return 16 * mt * (dim * 2 + 64) <= free_bytes;
}

template <int kMTiles, int kNTiles>
static int launch_fused(const HeadArgs& a, int grid, cudaStream_t st) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  constexpr int kRing = stages_for(kMTiles, kNTiles) * (kNTiles + 2 * kMTiles) * kThreads * 16;
  static_assert(kRing + static_smem_for(kMTiles, kNTiles) <= 232448, "ring + static shared memory exceed what a CTA may own");
  static_assert((kWarps * kMTiles * kNTiles * 4 * 32 + 16 * kNTiles) * 4 <= kRing, "the epilogue's partial tiles must fit into the ring");
  static std::atomic<bool> configured[64];                     // opt-in shared-memory size is a per-device attribute
  int dev = 0;
  ERN_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
    ERN_CUDA(cudaFuncSetAttribute(fused_head_kernel<kMTiles, kNTiles>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRing));
    if (dev >= 0 && dev < 64) configured[dev].store(true, std::memory_order_release);
  }
  cfg.dynamicSmemBytes = kRing;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;     // fails loudly instead of deadlocking if the grid cannot be co-resident
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ERN_CUDA(cudaLaunchKernelEx(&cfg, fused_head_kernel<kMTiles, kNTiles>, a));
  return ERN_OK;
}

int forward(const PackedView& pv, unsigned* sync, int dim, const float* image, const float* text, int64_t rows,
            __nv_bfloat16* raw, float* partial, float* out, void* out_bf16, int64_t ldb, float* gate, int sm_count,
            cudaStream_t st) {
  const int hid = 8 * dim;
  const int nt = pick_ntiles(hid, sm_count);
  const int grid = hid / (8 * nt);
  HeadArgs a;
  a.image = image; a.text = text; a.rows = static_cast<int>(rows); a.dim = dim; a.pv = pv; a.raw = raw;
  a.partial = partial; a.sync = sync; a.out = out; a.out_bf16 = static_cast<__nv_bfloat16*>(out_bf16); a.ldb = ldb;
  a.gate = gate;
  static unsigned long long* const stamps = [] {
    const char* e = getenv("ERN_HEAD_STAMP_PTR");
    return e ? reinterpret_cast<unsigned long long*>(strtoull(e, nullptr, 0)) : nullptr;
  }();
  a.stamps = stamps;
#define ERN_FUSED(M)                                                         \
  (nt == 4 ? launch_fused<M, 4>(a, grid, st) : nt == 5 ? launch_fused<M, 5>(a, grid, st) : launch_fused<M, 6>(a, grid, st))
  if (rows <= 16) return ERN_FUSED(1);
  if (rows <= 32) return ERN_FUSED(2);
  return ERN_FUSED(4);
#undef ERN_FUSED
}

}  // namespace small
}  // namespace combiner
}  // namespace ern
