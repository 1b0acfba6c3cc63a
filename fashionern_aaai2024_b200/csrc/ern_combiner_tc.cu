// Tensor-core path of the fusion head (CombinerSimple.forward, models/fusion_model.py:86-94).
//
// One persistent, warp-specialised tcgen05 GEMM kernel   C = relu(A . W^T + bias)   with two epilogues:
//   kStore : C is rounded to bf16 and stored            -> the two projections write the halves of raw[B,8D]
//                                                          (this replaces torch.cat, :90)
//   kGate  : C is multiplied by w2 and row-reduced       -> per 256-column tile partial of the gate logit;
//                                                          h[B,8D] (:74-75) never reaches HBM
// A [M,K] and W [N,K] are bf16, K-major, fed by TMA into a 4-stage ring of 128x64 / 256x64 swizzled tiles;
// the 128 x 256 fp32 accumulator is double buffered in TMEM (512 columns) so the epilogue of tile i overlaps
// the MMAs of tile i+1.  The finaliser (gate sigmoid, blend from the fp32 inputs, L2-normalise) is shared
// with the fp32 path.
#include "ern_internal.cuh"
#include "ern_ptx.cuh"

namespace ern {
namespace combiner {

namespace tc {

constexpr int kBlockM = 128;
constexpr int kBlockN = 256;
constexpr int kBlockK = 64;
constexpr int kStages = 4;
constexpr int kABytes = kBlockM * kBlockK * 2;   // 16 KB
constexpr int kBBytes = kBlockN * kBlockK * 2;   // 32 KB (two 128-row TMA boxes)
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kThreads = 192;
constexpr int kSmemBytes = 1024 + kStages * kStageBytes + 2 * 2 * kBlockN * 4 + 256;

struct Params {
  int64_t m;            // rows of A / C
  int n, k;             // N multiple of 256, K multiple of 64
  const float* bias;    // [N]
  // kStore
  __nv_bfloat16* out;
  int64_t ldo;
  int col0;
  // kGate
  const float* wg;      // [N]
  float* partial;       // [M, n_tiles]
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

template <bool kGate>
__global__ void __launch_bounds__(kThreads, 1)
linear_relu_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                      const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  float* epi_smem = reinterpret_cast<float*>(gen_base + kStages * kStageBytes);  // [2 buffers][2 arrays][256]
  Barriers* bars = reinterpret_cast<Barriers*>(gen_base + kStages * kStageBytes + 2 * 2 * kBlockN * 4);

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
  if (warp == 1) ptx::tmem_alloc<1>(ptx::smem_u32(&bars->tmem_base), 512);
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
          ptx::tma_load_2d(sa + kABytes, &tmap_w, kb * kBlockK, n0, full);
          ptx::tma_load_2d(sa + kABytes + kABytes, &tmap_w, kb * kBlockK, n0 + 128, full);
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
      // stage this tile's bias (and gate weights) in shared memory, double buffered across tiles
      float* sbias = epi_smem + buf * 2 * kBlockN;
      float* swg = sbias + kBlockN;
      for (int i = epi_tid; i < kBlockN; i += 128) {
        sbias[i] = p.bias[n0 + i];
        if (kGate) swg[i] = p.wg[n0 + i];
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      ptx::mbar_wait(ptx::smem_u32(&bars->tmem_full[acc]), acc_phase, nullptr, 14);
      ptx::tc_fence_after();
      const int64_t row = static_cast<int64_t>(m0) + quad * 32 + lane;
      const bool row_ok = row < p.m;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kBlockN;
      float dot = 0.f;
      uint32_t v[2][32];
      ptx::tmem_ld_32x32(taddr, v[0]);
#pragma unroll
      for (int c = 0; c < kBlockN / 32; ++c) {
        ptx::tmem_ld_wait();
        if (c + 1 < kBlockN / 32) ptx::tmem_ld_32x32(taddr + (c + 1) * 32, v[(c + 1) & 1]);
        const uint32_t(&cur)[32] = v[c & 1];
        if (kGate) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float h = fmaxf(__uint_as_float(cur[j]) + sbias[c * 32 + j], 0.f);
            dot = fmaf(h, swg[c * 32 + j], dot);
          }
        } else {
          uint32_t packed[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float a = fmaxf(__uint_as_float(cur[2 * j]) + sbias[c * 32 + 2 * j], 0.f);
            const float b = fmaxf(__uint_as_float(cur[2 * j + 1]) + sbias[c * 32 + 2 * j + 1], 0.f);
            packed[j] = pack_bf16x2(a, b);
          }
          if (row_ok) {
            uint4* dst = reinterpret_cast<uint4*>(p.out + row * p.ldo + p.col0 + n0 + c * 32);
#pragma unroll
            for (int j = 0; j < 4; ++j) dst[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
          }
        }
      }
      if (kGate && row_ok) p.partial[row * n_tiles + nt] = dot;
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
    ptx::tmem_dealloc<1>(tmem_base, 512);
  }
}

template <bool kGate>
static int launch(const CUtensorMap& ta, const CUtensorMap& tw, const Params& p, int sm_count, cudaStream_t st) {
  auto kern = linear_relu_tc_kernel<kGate>;
  static bool configured = false;
  if (!configured) {
    ERN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    configured = true;
  }
  const int64_t tiles = ((p.m + kBlockM - 1) / kBlockM) * (p.n / kBlockN);
  const int grid = static_cast<int>(tiles < sm_count ? tiles : sm_count);
  kern<<<grid, kThreads, kSmemBytes, st>>>(ta, tw, p);
  ERN_CUDA(cudaGetLastError());
  return ERN_OK;
}

}  // namespace tc

static size_t al(size_t x) { return (x + 255) & ~size_t(255); }

size_t workspace_bytes_bf16(int64_t rows, int dim) {
  const size_t r = rows, d = dim;
  return al(r * d * 2) * 2 + al(r * 8 * d * 2) + al(r * (8 * d / tc::kBlockN) * 4) + 512;
}

int forward_bf16(const ern_combiner_weights* w, int dim, const float* image, const float* text, int64_t rows,
                 float* out, void* out_bf16, int64_t ldb, float* gate, void* workspace, int sm_count,
                 cudaStream_t st) {
  if (rows <= 0) return ERN_OK;
  const size_t r = rows, d = dim;
  const int proj = 4 * dim, hid = 8 * dim;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  __nv_bfloat16* img_b = reinterpret_cast<__nv_bfloat16*>(ws);
  __nv_bfloat16* txt_b = reinterpret_cast<__nv_bfloat16*>(ws + al(r * d * 2));
  __nv_bfloat16* raw = reinterpret_cast<__nv_bfloat16*>(ws + 2 * al(r * d * 2));
  float* partial = reinterpret_cast<float*>(ws + 2 * al(r * d * 2) + al(r * 8 * d * 2));
  const int n_tiles = hid / tc::kBlockN;
  PackedView pv = view_packed(w->packed_bf16, dim);

  int rc = launch_cast_bf16(image, img_b, rows * d, st);
  if (rc) return rc;
  rc = launch_cast_bf16(text, txt_b, rows * d, st);
  if (rc) return rc;

  CUtensorMap t_img, t_txt, t_raw, t_wt, t_wi, t_w1;
  if ((rc = simtc::make_tmap_bf16_rows(&t_img, img_b, rows, dim, dim))) return rc;
  if ((rc = simtc::make_tmap_bf16_rows(&t_txt, txt_b, rows, dim, dim))) return rc;
  if ((rc = simtc::make_tmap_bf16_rows(&t_raw, raw, rows, hid, hid))) return rc;
  if ((rc = simtc::make_tmap_bf16_rows(&t_wt, pv.wt, proj, dim, dim))) return rc;
  if ((rc = simtc::make_tmap_bf16_rows(&t_wi, pv.wi, proj, dim, dim))) return rc;
  if ((rc = simtc::make_tmap_bf16_rows(&t_w1, pv.w1, hid, hid, hid))) return rc;

  tc::Params p{};
  p.m = rows;
  p.n = proj;
  p.k = dim;
  p.out = raw;
  p.ldo = hid;
  // text projection -> raw[:, 0:4D], image projection -> raw[:, 4D:8D]  (text half first, fusion_model.py:90)
  p.bias = pv.bt;
  p.col0 = 0;
  if ((rc = tc::launch<false>(t_txt, t_wt, p, sm_count, st))) return rc;
  p.bias = pv.bi;
  p.col0 = proj;
  if ((rc = tc::launch<false>(t_img, t_wi, p, sm_count, st))) return rc;
  // hidden layer + gate dot product
  tc::Params g{};
  g.m = rows;
  g.n = hid;
  g.k = hid;
  g.bias = pv.b1;
  g.wg = pv.w2;
  g.partial = partial;
  if ((rc = tc::launch<true>(t_raw, t_w1, g, sm_count, st))) return rc;
  return launch_finalize(image, text, rows, dim, partial, n_tiles, pv.b2, out, out_bf16, ldb, gate, st);
}

}  // namespace combiner
}  // namespace ern
