// Persistent warp-specialised tcgen05 GEMM   C = A . W^T   (A [M,K], W [N,K] bf16 K-major, fp32 accumulate in TMEM)
// with fused epilogues, shared by the fusion head (ern_combiner_tc.cu) and VisualSR (ern_visualsr.cu).
//   TMA producer warp -> kStages ring of {128 x 64 A tile, kBlockN x 64 W tile} (128B swizzle)
//   one elected thread issues tcgen05.mma (M = 128, N = kBlockN, K = 16); accumulators double buffered in TMEM
//   4 epilogue warps: tcgen05.ld 32 columns at a time, one output row per thread
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
};

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kStages = 4;
constexpr int kABytes = kBlockM * kBlockK * 2;   // 16 KB
constexpr int kThreads = 192;
constexpr int kStagedArrays = 4;                 // per-column parameter arrays staged in smem per tile
template <int kBlockN> constexpr int smem_bytes() {
  return 1024 + kStages * (kABytes + kBlockN * kBlockK * 2) + 2 * kStagedArrays * kBlockN * 4 + 256;
}

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
};

struct Barriers {
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// erf-GELU for the bf16 path: erf by Abramowitz-Stegun 7.1.26 (|error| < 1.5e-7, far below the bf16 rounding of
// the stored activation) -- ~4x fewer instructions than erff in an epilogue that otherwise paces the GEMM
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erf_abs = 1.0f - poly * t * __expf(-z * z);
  const float erf_x = copysignf(erf_abs, x);
  return 0.5f * x * (1.0f + erf_x);
}
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int kBlockN, int kEpi>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
               const Params p) {
  constexpr int kBBytes = kBlockN * kBlockK * 2;
  constexpr int kStageBytes = kABytes + kBBytes;
  constexpr int kTmemCols = 2 * kBlockN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  float* epi_smem = reinterpret_cast<float*>(gen_base + kStages * kStageBytes);  // [2 buffers][kStagedArrays][kBlockN]
  Barriers* bars = reinterpret_cast<Barriers*>(gen_base + kStages * kStageBytes + 2 * kStagedArrays * kBlockN * 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = static_cast<int>((p.m + kBlockM - 1) / kBlockM);
  const int n_tiles = p.n / kBlockN;
  const int total = m_tiles * n_tiles;
  const int kblocks = p.k / kBlockK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(ptx::smem_u32(&bars->full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars->empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(ptx::smem_u32(&bars->tmem_full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars->tmem_empty[s]), 4);
    }
    ptx::fence_barrier_init();
    ptx::fence_proxy_async();
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_a);
    ptx::prefetch_tensormap(&tmap_w);
  }
  if (warp == 1) ptx::tmem_alloc<1>(ptx::smem_u32(&bars->tmem_base), kTmemCols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const int m0 = (t / n_tiles) * kBlockM;
        const int n0 = (t % n_tiles) * kBlockN;
        for (int kb = 0; kb < kblocks; ++kb) {
          ptx::mbar_wait(ptx::smem_u32(&bars->empty[stage]), phase ^ 1, nullptr, 11);
          const uint32_t full = ptx::smem_u32(&bars->full[stage]);
          const uint32_t sa = smem_base + stage * kStageBytes;
          ptx::mbar_arrive_expect_tx(full, kStageBytes);
          ptx::tma_load_2d(sa, &tmap_a, kb * kBlockK, m0, full);
#pragma unroll
          for (int h = 0; h < kBlockN / 128; ++h)   // W tile as 128-row TMA boxes
            ptx::tma_load_2d(sa + kABytes + h * kABytes, &tmap_w, kb * kBlockK, n0 + h * 128, full);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::idesc_bf16_f32(kBlockM, kBlockN);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        ptx::mbar_wait(ptx::smem_u32(&bars->tmem_empty[acc]), acc_phase ^ 1, nullptr, 12);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kBlockN;
        for (int kb = 0; kb < kblocks; ++kb) {
          ptx::mbar_wait(ptx::smem_u32(&bars->full[stage]), phase, nullptr, 13);
          ptx::tc_fence_after();
          const uint32_t sa = smem_base + stage * kStageBytes;
          const uint64_t adesc = ptx::smem_desc_sw128(sa);
          const uint64_t bdesc = ptx::smem_desc_sw128(sa + kABytes);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)
            ptx::umma_bf16<1>(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          ptx::umma_commit<1>(ptx::smem_u32(&bars->empty[stage]));
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit<1>(ptx::smem_u32(&bars->tmem_full[acc]));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    const int quad = warp & 3;
    const int epi_tid = (warp - 2) * 32 + lane;  // 0..127
    uint32_t acc = 0, acc_phase = 0, buf = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      const int m0 = (t / n_tiles) * kBlockM;
      const int nt = t % n_tiles;
      const int n0 = nt * kBlockN;
      // stage this tile's per-column parameters in shared memory, double buffered across tiles
      float* sbias = epi_smem + buf * kStagedArrays * kBlockN;
      float* swg = sbias + kBlockN;
      float* sscale = swg + kBlockN;
      float* sshift = sscale + kBlockN;
      for (int i = epi_tid; i < kBlockN; i += 128) {
        sbias[i] = p.bias[n0 + i];
        if (kEpi == kEpiGate || kEpi == kEpiSrGlobal) swg[i] = p.wg[n0 + i];
        if (kEpi == kEpiSrGlobal) {
          sscale[i] = p.scale[n0 + i];
          sshift[i] = p.shift[n0 + i];
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      ptx::mbar_wait(ptx::smem_u32(&bars->tmem_full[acc]), acc_phase, nullptr, 14);
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
      uint32_t v[2][32];
      ptx::tmem_ld_32x32(taddr, v[0]);
#pragma unroll
      for (int c = 0; c < kBlockN / 32; ++c) {
        ptx::tmem_ld_wait();
        if (c + 1 < kBlockN / 32) ptx::tmem_ld_32x32(taddr + (c + 1) * 32, v[(c + 1) & 1]);
        const uint32_t(&cur)[32] = v[c & 1];
        if (kEpi == kEpiGate) {
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
        } else if (kEpi == kEpiSrGlobal) {
          if (row_ok) {
            float4* dst = reinterpret_cast<float4*>(p.out_f32 + row * p.ldo + n0 + c * 32);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              float o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int j = c * 32 + j4 * 4 + e;
                o[e] = tanh_fast(fmaf(sscale[j], __uint_as_float(cur[j4 * 4 + e]) + sbias[j], sshift[j])) * swg[j];
              }
              dst[j4] = make_float4(o[0], o[1], o[2], o[3]);
            }
          }
        } else if (kEpi == kEpiResidF32) {
          if (row_ok) {
            float4* dst = reinterpret_cast<float4*>(p.out_f32 + row * p.ldo + n0 + c * 32);
            const float4* res = p.residual ? reinterpret_cast<const float4*>(p.residual + row * p.ldo + n0 + c * 32) : nullptr;
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              float4 r4 = res ? res[j4] : make_float4(0.f, 0.f, 0.f, 0.f);
              const int j = c * 32 + j4 * 4;
              dst[j4] = make_float4(__uint_as_float(cur[j4 * 4 + 0]) + sbias[j + 0] + r4.x,
                                    __uint_as_float(cur[j4 * 4 + 1]) + sbias[j + 1] + r4.y,
                                    __uint_as_float(cur[j4 * 4 + 2]) + sbias[j + 2] + r4.z,
                                    __uint_as_float(cur[j4 * 4 + 3]) + sbias[j + 3] + r4.w);
            }
          }
        } else {
          uint32_t packed[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float a = __uint_as_float(cur[2 * j]) + sbias[c * 32 + 2 * j];
            float b = __uint_as_float(cur[2 * j + 1]) + sbias[c * 32 + 2 * j + 1];
            if (kEpi == kEpiStoreRelu) {
              a = fmaxf(a, 0.f);
              b = fmaxf(b, 0.f);
            } else if (kEpi == kEpiGeluBf16) {
              a = gelu_erf_fast(a);
              b = gelu_erf_fast(b);
            }
            packed[j] = pack_bf16x2(a, b);
          }
          if (row_ok) {
            uint4* dst = reinterpret_cast<uint4*>(p.out + row * p.ldo + p.col0 + n0 + c * 32);
#pragma unroll
            for (int j = 0; j < 4; ++j) dst[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
          }
        }
      }
      if ((kEpi == kEpiGate || kEpi == kEpiSrLocal) && row_ok) p.partial[row * n_tiles + nt] = dot;
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bars->tmem_empty[acc]));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      buf ^= 1;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<1>(tmem_base, kTmemCols);
  }
}

template <int kBlockN, int kEpi>
static int launch(const CUtensorMap& ta, const CUtensorMap& tw, const Params& p, int sm_count, cudaStream_t st) {
  auto kern = gemm_tc_kernel<kBlockN, kEpi>;
  // the opt-in shared-memory size is a per-device function attribute
  static bool configured[64] = {};
  int dev = 0;
  ERN_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    ERN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<kBlockN>()));
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  const int64_t tiles = ((p.m + kBlockM - 1) / kBlockM) * (p.n / kBlockN);
  const int grid = static_cast<int>(tiles < sm_count ? tiles : sm_count);
  kern<<<grid, kThreads, smem_bytes<kBlockN>(), st>>>(ta, tw, p);
  ERN_CUDA(cudaGetLastError());
  return ERN_OK;
}

}  // namespace gemmtc
}  // namespace ern
