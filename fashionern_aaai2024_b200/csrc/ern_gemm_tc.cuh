// Persistent warp-specialised tcgen05 GEMM   C = A . W^T   (A [M,K], W [N,K] bf16 K-major, fp32 accumulate in TMEM)
// with fused epilogues, shared by the fusion head (ern_combiner_tc.cu), VisualSR (ern_visualsr.cu) and the DVR
// encoder (ern_dvr.cu).
//   TMA producer warp -> ring of {128 x 64 A tile, W tile} stages (128B swizzle)
//   kPair (the mode every caller uses): the two CTAs of a cluster issue ONE tcgen05.mma.cta_group::2 of
//     M = 256 (128 rows of A per CTA) x N = 256 (each CTA streams half of the W tile) x K = 16 -- 256 x 256 tiles,
//     6-stage ring; 1-CTA mode: M = 128, N = kBlockN, 4 stages
//   accumulators double buffered in TMEM (2 x kBlockN columns)
//   4 epilogue warps per CTA: tcgen05.ld 32 columns at a time, one output row per thread, rolled chunk loop,
//     per-warp shared-memory transpose so that global stores / residual loads are full 128-byte lines
#pragma once
#include "ern_internal.cuh"
#include "ern_ptx.cuh"

namespace ern {
namespace gemmtc {

enum Epilogue {
  kEpiStoreRelu = 0,  // out_bf16[r, col0 + n] = bf16(relu(acc + bias[n]))
  kEpiGate = 1,       // partial[r, tile]  = sum_n relu(acc + bias[n]) * wg[n]
  kEpiSrGlobal = 2,   // out_f32[r, n]     = tanh(scale[n] * (acc + bias[n]) + shift[n]) * wg[n]
  kEpiSrLocal = 3,    // partial[r, tile]  = sum_n tanh(scale[r % P] * (acc + bias[n]) + shift[r % P]) * cvec[r / P, n]
  kEpiBiasBf16 = 4,   // out_bf16[r, col0 + n] = bf16(acc + bias[n])
  kEpiGeluBf16 = 5,   // out_bf16[r, col0 + n] = bf16(gelu_erf(acc + bias[n]))
  kEpiResidF32 = 6,   // out_f32[r, n] = acc + bias[n] + (residual ? residual[r, n] : 0)
  // in-batch classification loss (losses/loss.py:10-14), logits x[r, n] = alpha * acc:
  kEpiLse = 7,        // partial[tile, r] = max_n x, partial2[tile, r] = sum_n exp(x - max); diag[r] = x[r, r]
  kEpiSmGradRow = 8,  // out_bf16[r, n] = bf16((exp(x - rowvec[r]) - [r == n]) * beta * gscale[0]), 0 for n >= N
  kEpiSmGradCol = 9,  // out_bf16[r, n] = bf16((exp(x - bias[n])   - [r == n]) * beta * gscale[0]), 0 for n >= N
};

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kStages = 4;
constexpr int kABytes = kBlockM * kBlockK * 2;   // 16 KB
constexpr int kThreads = 192;
constexpr int kStagedArrays = 4;                 // per-column parameter arrays staged in smem per tile
constexpr int kStgPitch = 128 + 16;              // per-warp store-transpose buffer: 32 rows x 128 B, padded
constexpr int kStgWarpBytes = 32 * kStgPitch;

struct Params {
  int64_t m;            // rows of A / C
  int n, k;             // N multiple of kBlockN, K multiple of 64
  const float* bias;    // [N]
  __nv_bfloat16* out;   // kEpiStoreRelu
  int64_t ldo;
  int col0;
  const float* wg;      // [N]   kEpiGate / kEpiSrGlobal
  float* partial;       // [M, N / kBlockN]   kEpiGate / kEpiSrLocal
  float* out_f32;       // [M, ldo]           kEpiSrGlobal
  const float* scale;   // [N] (kEpiSrGlobal) or [P] (kEpiSrLocal): folded eval-mode BatchNorm
  const float* shift;
  const float* cvec;    // [M / P, N]         kEpiSrLocal
  int patches;          // P
  const float* residual;  // [M, ldo] fp32     kEpiResidF32 (nullable)
  float alpha, beta;      // kEpiLse / kEpiSmGrad*: logit scale, gradient coefficient
  float* partial2;        // [N / kBlockN, M]   kEpiLse (partial has the same layout there)
  float* diag;            // [M]                kEpiLse
  const float* rowvec;    // [M]                kEpiSmGradRow: per-row logsumexp
  const float* gscale;    // [1] device scalar multiplied into beta (upstream gradient), nullable
  int f16_operands;       // 1: A and W hold fp16 instead of bf16 (same 2-byte layout, same MMA rate; VisualSR)
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// erf-GELU for the bf16 path.  gelu(x) = x * Phi(x) = max(x, 0) - |x| * h with h = erfc(|x| / sqrt 2) / 2, and
// erfc(z) = 2^(z * P(z)) with a degree-4 minimax P (fitted on [0, 4.2], monotone beyond; |gelu error| < 1.2e-6 for
// all x, far below the bf16 rounding of the stored activation).  11 issue slots and ONE MUFU per element: the
// Abramowitz-Stegun form used before (rcp + ex2 = two MUFU ops, ~21 instructions) made the FFN-1 epilogue
// MUFU-bound at 2.5x the duration of its MMAs.
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  float p = fmaf(-0.00294416647f, z, 0.0295900748f);
  p = fmaf(p, z, -0.14866567f);
  p = fmaf(p, z, -0.918509346f);
  p = fmaf(p, z, -1.62788901f);
  p = fmaf(p, z, -1.0f);                         // log2(erfc(z) / 2)
  float h;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(h) : "f"(p));
  return fmaf(-fabsf(x), h, fmaxf(x, 0.f));
}
// e^x for the loss epilogues: one FMUL + MUFU.EX2 (flush-to-zero form: no denormal pre-scaling selects)
__device__ __forceinline__ float exp_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
  return y;
}
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int kBlockN, int kEpi, bool kPair>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
               const Params p) {
  // kPair: the two CTAs of a cluster issue one cta_group::2 UMMA of M = 256 (128 rows of A per CTA) x N = kBlockN
  // (each CTA streams half of the W tile): half the shared-memory operand traffic per SM of the 1-CTA form.
  constexpr int kCta = kPair ? 2 : 1;
  constexpr int kBRows = kBlockN / kCta;                 // W rows this CTA loads per stage
  constexpr int kBBytes = kBRows * kBlockK * 2;
  constexpr int kStageBytes = kABytes + kBBytes;
  constexpr int kNumStages = kPair ? 6 : kStages;
  constexpr int kTmemCols = 2 * kBlockN;
  static_assert(kBRows % 128 == 0, "W is loaded in 128-row TMA boxes");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  float* epi_smem = reinterpret_cast<float*>(gen_base + kNumStages * kStageBytes);  // [2][kStagedArrays][kBlockN]
  uint8_t* stage_smem = gen_base + kNumStages * kStageBytes + 2 * kStagedArrays * kBlockN * 4;   // [4 warps][kStgWarpBytes]
  uint64_t* bar_mem = reinterpret_cast<uint64_t*>(stage_smem + 4 * kStgWarpBytes);
  uint64_t* full_bar = bar_mem;                           // [kNumStages]
  uint64_t* empty_bar = bar_mem + kNumStages;             // [kNumStages]
  uint64_t* tmem_full_bar = bar_mem + 2 * kNumStages;     // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = kPair ? ptx::cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const int unit = kPair ? (blockIdx.x >> 1) : blockIdx.x;
  const int n_units = kPair ? (gridDim.x >> 1) : gridDim.x;
  const int m_tiles = static_cast<int>((p.m + kBlockM * kCta - 1) / (kBlockM * kCta));
  const int n_tiles = (p.n + kBlockN - 1) / kBlockN;
  const int total = m_tiles * n_tiles;
  const int kblocks = p.k / kBlockK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kNumStages; ++s) {
      ptx::mbar_init(ptx::smem_u32(&full_bar[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&empty_bar[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(ptx::smem_u32(&tmem_full_bar[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&tmem_empty_bar[s]), 4 * kCta);
    }
    ptx::fence_barrier_init();
    ptx::fence_proxy_async();
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_a);
    ptx::prefetch_tensormap(&tmap_w);
  }
  if (warp == 1) ptx::tmem_alloc<kCta>(ptx::smem_u32(tmem_slot), kTmemCols);
  ptx::tc_fence_before();
  if (kPair) ptx::cluster_sync_all(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int t = unit; t < total; t += n_units) {
        const int m0 = (t / n_tiles) * (kBlockM * kCta) + rank * kBlockM;
        // ragged last column tile with at most kBlockN / 2 real columns: the MMA warp issues half-width MMAs for it,
        // which read the first kBRows / 2 rows of each CTA's W tile -- so CTA 1's rows start kBRows / 2 further on
        const int nt0 = (t % n_tiles) * kBlockN;
        const bool narrow = kPair && nt0 + kBlockN / 2 >= p.n;
        const int n0 = nt0 + rank * (narrow ? kBRows / 2 : kBRows);
        for (int kb = 0; kb < kblocks; ++kb) {
          ptx::mbar_wait(ptx::smem_u32(&empty_bar[stage]), phase ^ 1, nullptr, 11);
          const uint32_t full = ptx::smem_u32(&full_bar[stage]);
          const uint32_t sa = smem_base + stage * kStageBytes;
          if (leader) ptx::mbar_arrive_expect_tx(full, kStageBytes * kCta);
          if (kPair) {
            const uint32_t dst = ptx::mapa(full, 0);
            ptx::tma_load_2d_pair(sa, &tmap_a, kb * kBlockK, m0, dst);
#pragma unroll
            for (int h = 0; h < kBRows / 128; ++h)
              ptx::tma_load_2d_pair(sa + kABytes + h * kABytes, &tmap_w, kb * kBlockK, n0 + h * 128, dst);
          } else {
            ptx::tma_load_2d(sa, &tmap_a, kb * kBlockK, m0, full);
#pragma unroll
            for (int h = 0; h < kBRows / 128; ++h)   // W tile as 128-row TMA boxes
              ptx::tma_load_2d(sa + kABytes + h * kABytes, &tmap_w, kb * kBlockK, n0 + h * 128, full);
          }
          if (++stage == kNumStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (leader && lane == 0) {
      constexpr uint32_t idesc_full = ptx::idesc_bf16_f32(kBlockM * kCta, kBlockN);
      constexpr uint32_t idesc_half = ptx::idesc_bf16_f32(kBlockM * kCta, kBlockN / 2);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int t = unit; t < total; t += n_units) {
        const bool narrow = kPair && (t % n_tiles) * kBlockN + kBlockN / 2 >= p.n;
        // (A / B format fields: 1 = bf16, 0 = fp16)
        const uint32_t idesc = (narrow ? idesc_half : idesc_full) & (p.f16_operands ? ~((1u << 7) | (1u << 10)) : ~0u);
        ptx::mbar_wait(ptx::smem_u32(&tmem_empty_bar[acc]), acc_phase ^ 1, nullptr, 12);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kBlockN;
        for (int kb = 0; kb < kblocks; ++kb) {
          ptx::mbar_wait(ptx::smem_u32(&full_bar[stage]), phase, nullptr, 13);
          ptx::tc_fence_after();
          const uint32_t sa = smem_base + stage * kStageBytes;
          const uint64_t adesc = ptx::smem_desc_sw128(sa);
          const uint64_t bdesc = ptx::smem_desc_sw128(sa + kABytes);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)
            ptx::umma_bf16<kCta>(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          ptx::umma_commit<kCta>(ptx::smem_u32(&empty_bar[stage]));
          if (++stage == kNumStages) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit<kCta>(ptx::smem_u32(&tmem_full_bar[acc]));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    const int quad = warp & 3;
    const int epi_tid = (warp - 2) * 32 + lane;  // 0..127
    uint32_t acc = 0, acc_phase = 0, buf = 0;
    for (int t = unit; t < total; t += n_units) {
      const int m0 = (t / n_tiles) * (kBlockM * kCta) + rank * kBlockM;
      const int nt = t % n_tiles;
      const int n0 = nt * kBlockN;
      // stage this tile's per-column parameters in shared memory, double buffered across tiles
      float* sbias = epi_smem + buf * kStagedArrays * kBlockN;
      float* swg = sbias + kBlockN;
      float* sscale = swg + kBlockN;
      float* sshift = sscale + kBlockN;
      for (int i = epi_tid; i < kBlockN; i += 128) {
        const bool in = n0 + i < p.n;
        sbias[i] = (in && p.bias) ? p.bias[n0 + i] : 0.f;
        if (kEpi == kEpiGate || kEpi == kEpiSrGlobal) swg[i] = in ? p.wg[n0 + i] : 0.f;
        if (kEpi == kEpiSrGlobal) {
          sscale[i] = in ? p.scale[n0 + i] : 0.f;
          sshift[i] = in ? p.shift[n0 + i] : 0.f;
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      ptx::mbar_wait(ptx::smem_u32(&tmem_full_bar[acc]), acc_phase, nullptr, 14);
      ptx::tc_fence_after();
      const int64_t row = static_cast<int64_t>(m0) + quad * 32 + lane;
      const bool row_ok = row < p.m;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kBlockN;
      float dot = 0.f;
      float row_scale = 0.f, row_shift = 0.f;
      const float* crow = nullptr;
      if (kEpi == kEpiSrLocal) {
        const int64_t rr = row_ok ? row : 0;
        const int pidx = static_cast<int>(rr % p.patches);
        row_scale = p.scale[pidx];
        row_shift = p.shift[pidx];
        crow = p.cvec + (rr / p.patches) * p.n + n0;
      }
      // The chunk loops are deliberately rolled: unrolling 8 chunks x 32 elements of epilogue math made a 116 KB
      // kernel that stalled on instruction fetch (ncu: stall_no_inst top, 20 % tensor pipe on the GELU GEMM).
      // Stores go through a per-warp shared-memory transpose: a thread owns one output ROW, so direct stores would
      // touch 32 different 128-byte lines per instruction (ncu: 32 sectors/request); after the transpose every
      // store instruction writes 4 complete lines.
      uint8_t* stg = stage_smem + (warp - 2) * kStgWarpBytes;
      const int64_t row_base = static_cast<int64_t>(m0) + quad * 32;
      const int srow = lane >> 3, spiece = lane & 7;
      constexpr bool kSmGrad = kEpi == kEpiSmGradRow || kEpi == kEpiSmGradCol;
      float row_lse = 0.f, coef = 0.f;
      if (kSmGrad) {
        coef = p.beta * (p.gscale ? p.gscale[0] : 1.f);
        if (kEpi == kEpiSmGradRow && row_ok) row_lse = p.rowvec[row];
      }
      float run_max = -INFINITY, run_sum = 0.f, diag_val = 0.f;   // kEpiLse
      if (kEpi == kEpiStoreRelu || kEpi == kEpiBiasBf16 || kEpi == kEpiGeluBf16 || kSmGrad) {
#pragma unroll 1
        for (int g = 0; g < kBlockN / 64; ++g) {
          if (n0 + g * 64 >= p.n) break;     // ragged last column tile: nothing more to emit
          uint32_t lo[32], hi[32];
          ptx::tmem_ld_32x32(taddr + g * 64, lo);
          ptx::tmem_ld_32x32(taddr + g * 64 + 32, hi);
          ptx::tmem_ld_wait();
          uint32_t packed[32];
          // kSmGrad: position of this row's diagonal element / number of real columns, relative to this group
          const int64_t dj64 = row - n0 - g * 64;
          const int dj = (dj64 >= 0 && dj64 < 64) ? static_cast<int>(dj64) : -1;
          const int valid = p.n - n0 - g * 64;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int cidx = g * 64 + 2 * j;
            float x0 = __uint_as_float(j < 16 ? lo[2 * j] : hi[2 * j - 32]);
            float x1 = __uint_as_float(j < 16 ? lo[2 * j + 1] : hi[2 * j - 31]);
            if (kSmGrad) {
              // The logit is rounded exactly as in the forward pass (no fma contraction), so that exp(x - lse) is
              // consistent with the lse it was reduced into (a 1-row batch gives a gradient of exactly 0).
              // Diagonal and ragged-column handling are selects, not branches (a ternary around __expf compiled to
              // 86 divergent BSSY/BSYNC regions per 64 columns and made this epilogue 3x slower than the MMAs).
              const float l0 = kEpi == kEpiSmGradRow ? row_lse : sbias[cidx];
              const float l1 = kEpi == kEpiSmGradRow ? row_lse : sbias[cidx + 1];
              float e0 = exp_fast(__fmul_rn(p.alpha, x0) - l0);
              float e1 = exp_fast(__fmul_rn(p.alpha, x1) - l1);
              e0 = (2 * j == dj) ? e0 - 1.f : e0;
              e1 = (2 * j + 1 == dj) ? e1 - 1.f : e1;
              x0 = (2 * j < valid) ? e0 * coef : 0.f;
              x1 = (2 * j + 1 < valid) ? e1 * coef : 0.f;
            } else {
              x0 += sbias[cidx];
              x1 += sbias[cidx + 1];
            }
            if (kEpi == kEpiStoreRelu) {
              x0 = fmaxf(x0, 0.f);
              x1 = fmaxf(x1, 0.f);
            } else if (kEpi == kEpiGeluBf16) {
              x0 = gelu_erf_fast(x0);
              x1 = gelu_erf_fast(x1);
            }
            packed[j] = pack_bf16x2(x0, x1);
          }
          uint4* mine = reinterpret_cast<uint4*>(stg + lane * kStgPitch);
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) mine[q4] = make_uint4(packed[4 * q4], packed[4 * q4 + 1], packed[4 * q4 + 2], packed[4 * q4 + 3]);
          __syncwarp();
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int r = srow + 4 * it;
            const uint4 val = *reinterpret_cast<const uint4*>(stg + r * kStgPitch + spiece * 16);
            if (row_base + r < p.m)
              *reinterpret_cast<uint4*>(p.out + (row_base + r) * p.ldo + p.col0 + n0 + g * 64 + spiece * 8) = val;
          }
          __syncwarp();
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < kBlockN / 32; ++c) {
          if (n0 + c * 32 >= p.n) break;
          uint32_t cur[32];
          ptx::tmem_ld_32x32(taddr + c * 32, cur);
          ptx::tmem_ld_wait();
          if (kEpi == kEpiLse) {
            // online logsumexp over this tile's columns; columns >= N (ragged last tile) are skipped
            const int valid = p.n - (n0 + c * 32);
            const int dj = static_cast<int>(row - (n0 + c * 32));
            float x[32];
            float cmax = -INFINITY;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              x[j] = j < valid ? __fmul_rn(p.alpha, __uint_as_float(cur[j])) : -INFINITY;
              cmax = fmaxf(cmax, x[j]);
              if (j == dj) diag_val = x[j];
            }
            const float new_max = fmaxf(run_max, cmax);
            float csum = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) csum += exp_fast(x[j] - new_max);
            run_sum = run_sum * exp_fast(run_max - new_max) + csum;
            run_max = new_max;
          } else if (kEpi == kEpiGate) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float h = fmaxf(__uint_as_float(cur[j]) + sbias[c * 32 + j], 0.f);
              dot = fmaf(h, swg[c * 32 + j], dot);
            }
          } else if (kEpi == kEpiSrLocal) {
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 cv = *reinterpret_cast<const float4*>(crow + c * 32 + j4 * 4);
              const float cc[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int j = j4 * 4 + e;
                const float h = tanh_fast(fmaf(row_scale, __uint_as_float(cur[j]) + sbias[c * 32 + j], row_shift));
                dot = fmaf(h, cc[e], dot);
              }
            }
          } else {
            // fp32 row outputs (kEpiResidF32 / kEpiSrGlobal): 32 columns = one 128-byte line per row
            float o[32];
            if (kEpi == kEpiResidF32) {
              if (p.residual) {
                // coalesced read of the residual tile, transposed through shared memory to row ownership
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                  const int r = srow + 4 * it;
                  uint4 val = make_uint4(0u, 0u, 0u, 0u);
                  if (row_base + r < p.m)
                    val = *reinterpret_cast<const uint4*>(p.residual + (row_base + r) * p.ldo + n0 + c * 32 + spiece * 4);
                  *reinterpret_cast<uint4*>(stg + r * kStgPitch + spiece * 16) = val;
                }
                __syncwarp();
                const float4* mine_r = reinterpret_cast<const float4*>(stg + lane * kStgPitch);
#pragma unroll
                for (int q4 = 0; q4 < 8; ++q4) {
                  const float4 r4 = mine_r[q4];
                  o[4 * q4 + 0] = r4.x;
                  o[4 * q4 + 1] = r4.y;
                  o[4 * q4 + 2] = r4.z;
                  o[4 * q4 + 3] = r4.w;
                }
                __syncwarp();
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) o[j] = 0.f;
              }
#pragma unroll
              for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(cur[j]) + sbias[c * 32 + j] + o[j];
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int cj = c * 32 + j;
                o[j] = tanh_fast(fmaf(sscale[cj], __uint_as_float(cur[j]) + sbias[cj], sshift[cj])) * swg[cj];
              }
            }
            float4* mine = reinterpret_cast<float4*>(stg + lane * kStgPitch);
#pragma unroll
            for (int q4 = 0; q4 < 8; ++q4) mine[q4] = make_float4(o[4 * q4], o[4 * q4 + 1], o[4 * q4 + 2], o[4 * q4 + 3]);
            __syncwarp();
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int r = srow + 4 * it;
              const uint4 val = *reinterpret_cast<const uint4*>(stg + r * kStgPitch + spiece * 16);
              if (row_base + r < p.m)
                *reinterpret_cast<uint4*>(p.out_f32 + (row_base + r) * p.ldo + n0 + c * 32 + spiece * 4) = val;
            }
            __syncwarp();
          }
        }
      }
      if ((kEpi == kEpiGate || kEpi == kEpiSrLocal) && row_ok) p.partial[row * n_tiles + nt] = dot;
      if (kEpi == kEpiLse && row_ok) {
        p.partial[nt * p.m + row] = run_max;      // [tile, M]: consecutive lanes own consecutive rows
        p.partial2[nt * p.m + row] = run_sum;
        if (row >= n0 && row < n0 + kBlockN && row < p.n) p.diag[row] = diag_val;
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        const uint32_t dst = ptx::smem_u32(&tmem_empty_bar[acc]);
        if (kPair) ptx::mbar_arrive_cluster(ptx::mapa(dst, 0)); else ptx::mbar_arrive(dst);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      buf ^= 1;
    }
  }

  ptx::tc_fence_before();
  if (kPair) ptx::cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<kCta>(tmem_base, kTmemCols);
  }
}

template <int kBlockN, bool kPair> constexpr int smem_bytes_v() {
  return 1024 + (kPair ? 6 : kStages) * (kABytes + (kBlockN / (kPair ? 2 : 1)) * kBlockK * 2) +
         2 * kStagedArrays * kBlockN * 4 + 4 * kStgWarpBytes + 256;
}

// number of column tiles (and of per-row partials the kEpiGate / kEpiSrLocal epilogues emit)
template <int kBlockN> static inline int n_tiles_of(int n) { return (n + kBlockN - 1) / kBlockN; }

template <int kBlockN, int kEpi, bool kPair = false>
static int launch(const CUtensorMap& ta, const CUtensorMap& tw, const Params& p, int sm_count, cudaStream_t st) {
  auto kern = gemm_tc_kernel<kBlockN, kEpi, kPair>;
  constexpr int smem = smem_bytes_v<kBlockN, kPair>();
  // the opt-in shared-memory size is a per-device function attribute
  static std::atomic<bool> configured[64];   // zero-initialised; idempotent per-device attribute set
  int dev = 0;
  ERN_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
    ERN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (dev >= 0 && dev < 64) configured[dev].store(true, std::memory_order_release);
  }
  constexpr int kCta = kPair ? 2 : 1;
  const int64_t tiles = ((p.m + kBlockM * kCta - 1) / (kBlockM * kCta)) * n_tiles_of<kBlockN>(p.n);
  int units = kPair ? sm_count / 2 : sm_count;
  if (tiles < units) units = static_cast<int>(tiles);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(units * kCta);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCta;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ERN_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tw, p));
  return ERN_OK;
}

}  // namespace gemmtc
}  // namespace ern
