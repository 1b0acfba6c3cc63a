// VisualSR.forward (models/fusion_model.py:141-154) -- SURVEY.md 8(f) "next" row 1: the attention pooling of the
// 13 patch embeddings that feeds the gallery-side fusion head (models/model.py:65) and the DVR module
// (models/fusion_model.py:48).  Eval mode: Dropout = identity, BatchNorm1d = per-channel affine (folded by the
// caller into scale/shift).
//
//   raw_global = mean_p local[b,p,:]                                         (:142)
//   l_emb[b,p,:] = tanh(bn_P (local[b,p,:] Wl^T + bl))                        (:145; BatchNorm1d(13): channel = p)
//   g_emb[b,:]   = tanh(bn_D (raw_global  Wg^T + bg))                         (:146; BatchNorm1d(D): channel = d)
//   logit[b,p]   = (l_emb[b,p,:] * g_emb[b,:]) . wc + bc ; w = softmax_p      (:149-150)
//   out[b,:]     = sum_p w[b,p] local[b,p,:] ;  out / (||out|| + 1e-8)        (:153-154, :136-139)
//
// The [B*13, D] x [D, D] GEMM runs on the tensor cores (ern_gemm_tc.cuh); its epilogue applies the affine, tanh and
// the product with c[b,:] = g_emb[b,:] * wc and row-reduces, so l_emb never reaches HBM.
#include <cuda_fp16.h>

#include "ern_gemm_tc.cuh"

namespace ern {
namespace visualsr {

constexpr int kMaxPatches = 32;

// The tensor-core path of VisualSR feeds its GEMMs FP16 operands, not bf16: the 13-way softmax of the attention
// logits amplifies operand rounding (bf16, 8 mantissa bits: worst output row 1.06e-2 on N(0,1) synthetic patches, a CPU
// emulation attributes ~5e-3 each to the rounding of the patches, of W_local and of the global branch); fp16 has 11
// bits and runs at the same tcgen05 rate (kind::f16 takes either).  CLIP patch features and Xavier weights sit well
// inside fp16's range; the conversion saturates at +-65504 instead of producing infinities.
__device__ __forceinline__ uint32_t pack_f16x2(float a, float b) {
  a = fminf(fmaxf(a, -65504.f), 65504.f);
  b = fminf(fmaxf(b, -65504.f), 65504.f);
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ __half to_f16(float a) { return __float2half_rn(fminf(fmaxf(a, -65504.f), 65504.f)); }

// one warp per row b: mean over patches (+ bf16 copies of the patch matrix and of the mean for the GEMMs); HBM-bound
__global__ void prepare_kernel(const float* __restrict__ local, int64_t rows, int patches, int dim,
                               float* __restrict__ mean_f32, __half* __restrict__ mean_b,
                               __half* __restrict__ local_b) {
  const int64_t b = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= rows) return;
  const float* x = local + b * patches * dim;
  const float inv = static_cast<float>(patches);
  if ((dim & 3) == 0) {
    for (int d = lane * 4; d < dim; d += 128) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int p = 0; p < patches; ++p) {
        const float4 v = *reinterpret_cast<const float4*>(x + p * dim + d);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        if (local_b) {
          *reinterpret_cast<uint2*>(local_b + (b * patches + p) * dim + d) =
              make_uint2(pack_f16x2(v.x, v.y), pack_f16x2(v.z, v.w));
        }
      }
      const float4 m = make_float4(s.x / inv, s.y / inv, s.z / inv, s.w / inv);
      if (mean_f32) *reinterpret_cast<float4*>(mean_f32 + b * dim + d) = m;
      if (mean_b) {
        *reinterpret_cast<uint2*>(mean_b + b * dim + d) = make_uint2(pack_f16x2(m.x, m.y), pack_f16x2(m.z, m.w));
      }
    }
    return;
  }
  for (int d = lane; d < dim; d += 32) {
    float s = 0.f;
    for (int p = 0; p < patches; ++p) {
      const float v = x[p * dim + d];
      s += v;
      if (local_b) local_b[(b * patches + p) * dim + d] = to_f16(v);
    }
    const float m = s / inv;
    if (mean_f32) mean_f32[b * dim + d] = m;
    if (mean_b) mean_b[b * dim + d] = to_f16(m);
  }
}

// fp32 validation GEMM, 64x64 tiles (see ern_combiner.cu for the tile structure); kLocal selects the epilogue
constexpr int kTile = 64;
constexpr int kKc = 16;
constexpr int kF32Threads = 256;

template <bool kLocal>
__global__ void __launch_bounds__(kF32Threads)
linear_tanh_f32_kernel(const float* __restrict__ X, int64_t rows, const float* __restrict__ W, int K, int N,
                       const float* __restrict__ bias, const float* __restrict__ scale,
                       const float* __restrict__ shift, const float* __restrict__ wc, float* __restrict__ out,
                       const float* __restrict__ cvec, int patches, float* __restrict__ partial, int n_tiles) {
  __shared__ float xs[kKc][kTile + 1];
  __shared__ float ws[kKc][kTile + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * kTile;
  const int n0 = blockIdx.x * kTile;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += kKc) {
    for (int e = threadIdx.x; e < kTile * kKc; e += kF32Threads) {
      const int r = e / kKc, kk = e % kKc;
      const bool kin = (k0 + kk) < K;
      xs[kk][r] = (kin && r0 + r < rows) ? X[(r0 + r) * K + k0 + kk] : 0.f;
      ws[kk][r] = (kin && n0 + r < N) ? W[static_cast<int64_t>(n0 + r) * K + k0 + kk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kKc; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = xs[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = ws[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t r = r0 + ty * 4 + i;
    const int64_t rr = r < rows ? r : 0;
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      const float lin = acc[i][j] + bias[n];
      if (kLocal) {
        const int p = static_cast<int>(rr % patches);
        dot = fmaf(tanhf(fmaf(scale[p], lin, shift[p])), cvec[(rr / patches) * N + n], dot);
      } else if (r < rows) {
        out[r * N + n] = tanhf(fmaf(scale[n], lin, shift[n])) * wc[n];
      }
    }
    if (kLocal) {
      for (int o = 8; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      if (tx == 0 && r < rows) partial[r * n_tiles + blockIdx.x] = dot;
    }
  }
}

// one warp per row b: logits -> softmax over the patches -> weighted sum of the fp32 patch features -> l2norm(+1e-8)
// HBM-bound (reads the [P, D] fp32 patch block once): float4 loads, the P softmax weights live in registers.
__global__ void finalize_kernel(const float* __restrict__ local, int64_t rows, int patches, int dim,
                                const float* __restrict__ partial, int n_tiles, const float* __restrict__ b_common,
                                float* __restrict__ out) {
  const int64_t b = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= rows) return;
  float logit = -INFINITY;
  if (lane < patches) {
    float z = 0.f;
    for (int t = 0; t < n_tiles; ++t) z += partial[(b * patches + lane) * n_tiles + t];
    logit = z + b_common[0];
  }
  float mx = logit;
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float e = lane < patches ? expf(logit - mx) : 0.f;
  float sum = e;
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float w = e / sum;
  float wp[kMaxPatches];
#pragma unroll
  for (int p = 0; p < kMaxPatches; ++p) wp[p] = __shfl_sync(0xffffffffu, w, p);
  const float* x = local + b * patches * dim;
  float* o_row = out + b * dim;
  float ss = 0.f;
  if ((dim & 3) == 0) {
    for (int d = lane * 4; d < dim; d += 128) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int p = 0; p < kMaxPatches; ++p) {
        if (p < patches) {
          const float4 v = *reinterpret_cast<const float4*>(x + p * dim + d);
          acc.x = fmaf(wp[p], v.x, acc.x);
          acc.y = fmaf(wp[p], v.y, acc.y);
          acc.z = fmaf(wp[p], v.z, acc.z);
          acc.w = fmaf(wp[p], v.w, acc.w);
        }
      }
      *reinterpret_cast<float4*>(o_row + d) = acc;
      ss = fmaf(acc.x, acc.x, fmaf(acc.y, acc.y, fmaf(acc.z, acc.z, fmaf(acc.w, acc.w, ss))));
    }
  } else {
    for (int d = lane; d < dim; d += 32) {
      float acc = 0.f;
#pragma unroll
      for (int p = 0; p < kMaxPatches; ++p)
        if (p < patches) acc = fmaf(wp[p], x[p * dim + d], acc);
      o_row[d] = acc;
      ss = fmaf(acc, acc, ss);
    }
  }
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float denom = sqrtf(ss) + 1e-8f;
  __syncwarp();
  if ((dim & 3) == 0) {
    for (int d = lane * 4; d < dim; d += 128) {
      float4 v = *reinterpret_cast<float4*>(o_row + d);
      v.x /= denom; v.y /= denom; v.z /= denom; v.w /= denom;
      *reinterpret_cast<float4*>(o_row + d) = v;
    }
  } else {
    for (int d = lane; d < dim; d += 32) o_row[d] = o_row[d] / denom;
  }
}

// prepare_kernel for the reference's shapes (P = 13, D = 128 * kIters): loops unrolled so that the 13 independent
// float4 loads of a column block are in flight together
template <int kP, int kIters>
__global__ void __launch_bounds__(256)
prepare_fixed_kernel(const float* __restrict__ local, int64_t rows, float* __restrict__ mean_f32,
                     __half* __restrict__ mean_b, __half* __restrict__ local_b) {
  constexpr int kDim = 128 * kIters;
  const int64_t b = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= rows) return;
  const float* x = local + b * kP * kDim + lane * 4;
#pragma unroll
  for (int it = 0; it < kIters; ++it) {
    float4 v[kP];
#pragma unroll
    for (int p = 0; p < kP; ++p) v[p] = *reinterpret_cast<const float4*>(x + p * kDim + it * 128);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int p = 0; p < kP; ++p) {
      s.x += v[p].x; s.y += v[p].y; s.z += v[p].z; s.w += v[p].w;
      if (local_b) {
        *reinterpret_cast<uint2*>(local_b + (b * kP + p) * kDim + lane * 4 + it * 128) =
            make_uint2(pack_f16x2(v[p].x, v[p].y), pack_f16x2(v[p].z, v[p].w));
      }
    }
    const float inv = static_cast<float>(kP);
    const float4 m = make_float4(s.x / inv, s.y / inv, s.z / inv, s.w / inv);
    if (mean_f32) *reinterpret_cast<float4*>(mean_f32 + b * kDim + lane * 4 + it * 128) = m;
    if (mean_b) {
      *reinterpret_cast<uint2*>(mean_b + b * kDim + lane * 4 + it * 128) =
          make_uint2(pack_f16x2(m.x, m.y), pack_f16x2(m.z, m.w));
    }
  }
}

static void launch_prepare_sr(const float* local, int64_t rows, int patches, int dim, float* mean_f32,
                              __half* mean_b, __half* local_b, cudaStream_t st) {
  const int blocks = cdiv(rows * 32, 256);
  const bool aligned = ((reinterpret_cast<uintptr_t>(local) | reinterpret_cast<uintptr_t>(mean_f32) |
                         reinterpret_cast<uintptr_t>(mean_b) | reinterpret_cast<uintptr_t>(local_b)) & 15u) == 0;
  if (patches == 13 && dim == 640 && aligned)
    prepare_fixed_kernel<13, 5><<<blocks, 256, 0, st>>>(local, rows, mean_f32, mean_b, local_b);
  else if (patches == 13 && dim == 512 && aligned)
    prepare_fixed_kernel<13, 4><<<blocks, 256, 0, st>>>(local, rows, mean_f32, mean_b, local_b);
  else
    prepare_kernel<<<blocks, 256, 0, st>>>(local, rows, patches, dim, mean_f32, mean_b, local_b);
}

// Specialisation for the reference's shapes (P = 13 patches, D = 128 * kIters): the first block of patch loads is
// issued before the softmax weights are known, the accumulators stay in registers through the normalisation (one
// write of the output row instead of write + re-read + write), and all loops are unrolled so that the independent
// float4 loads of a row overlap.  The generic kernel above was latency x occupancy bound (2.8 TB/s at 106 registers).
template <int kP, int kIters>
__global__ void __launch_bounds__(256)
finalize_fixed_kernel(const float* __restrict__ local, int64_t rows, const float* __restrict__ partial, int n_tiles,
                      const float* __restrict__ b_common, float* __restrict__ out) {
  constexpr int kDim = 128 * kIters;
  const int64_t b = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= rows) return;
  const float* x = local + b * kP * kDim + lane * 4;
  float4 first[kP];
#pragma unroll
  for (int p = 0; p < kP; ++p) first[p] = *reinterpret_cast<const float4*>(x + p * kDim);
  float logit = -INFINITY;
  if (lane < kP) {
    float z = 0.f;
    for (int t = 0; t < n_tiles; ++t) z += partial[(b * kP + lane) * n_tiles + t];
    logit = z + b_common[0];
  }
  float mx = logit;
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const float e = lane < kP ? expf(logit - mx) : 0.f;
  float sum = e;
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float w = e / sum;
  float wp[kP];
#pragma unroll
  for (int p = 0; p < kP; ++p) wp[p] = __shfl_sync(0xffffffffu, w, p);
  float4 acc[kIters];
#pragma unroll
  for (int it = 0; it < kIters; ++it) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int p = 0; p < kP; ++p) {
      const float4 v = it == 0 ? first[p] : *reinterpret_cast<const float4*>(x + p * kDim + it * 128);
      a.x = fmaf(wp[p], v.x, a.x);
      a.y = fmaf(wp[p], v.y, a.y);
      a.z = fmaf(wp[p], v.z, a.z);
      a.w = fmaf(wp[p], v.w, a.w);
    }
    acc[it] = a;
  }
  float ss = 0.f;
#pragma unroll
  for (int it = 0; it < kIters; ++it)
    ss = fmaf(acc[it].x, acc[it].x, fmaf(acc[it].y, acc[it].y, fmaf(acc[it].z, acc[it].z, fmaf(acc[it].w, acc[it].w, ss))));
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float denom = sqrtf(ss) + 1e-8f;
  float* o_row = out + b * kDim + lane * 4;
#pragma unroll
  for (int it = 0; it < kIters; ++it) {
    float4 v = acc[it];
    v.x /= denom; v.y /= denom; v.z /= denom; v.w /= denom;
    *reinterpret_cast<float4*>(o_row + it * 128) = v;
  }
}

static void launch_finalize_sr(const float* local, int64_t rows, int patches, int dim, const float* partial, int n_tiles,
                               const float* b_common, float* out, cudaStream_t st) {
  const int blocks = cdiv(rows * 32, 256);
  const bool aligned = ((reinterpret_cast<uintptr_t>(local) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
  if (patches == 13 && dim == 640 && aligned)
    finalize_fixed_kernel<13, 5><<<blocks, 256, 0, st>>>(local, rows, partial, n_tiles, b_common, out);
  else if (patches == 13 && dim == 512 && aligned)
    finalize_fixed_kernel<13, 4><<<blocks, 256, 0, st>>>(local, rows, partial, n_tiles, b_common, out);
  else
    finalize_kernel<<<blocks, 256, 0, st>>>(local, rows, patches, dim, partial, n_tiles, b_common, out);
}

static size_t al(size_t x) { return (x + 255) & ~size_t(255); }
constexpr int kTcBlockN = 256;   // CTA-pair tiles; a ragged last column tile (D = 640) is masked in the epilogue

size_t packed_bytes(int dim) { return al(static_cast<size_t>(dim) * dim * 2) * 2 + 256; }

__global__ void cast_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, int64_t n) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n) dst[i] = to_f16(src[i]);
}

int pack(const ern_visualsr_weights* w, int dim, void* packed, cudaStream_t st) {
  const int64_t n = static_cast<int64_t>(dim) * dim;
  uint8_t* p = static_cast<uint8_t*>(packed);
  cast_f16_kernel<<<cdiv(n, 256), 256, 0, st>>>(w->w_local, reinterpret_cast<__half*>(p), n);
  cast_f16_kernel<<<cdiv(n, 256), 256, 0, st>>>(w->w_global, reinterpret_cast<__half*>(p + al(n * 2)), n);
  ERN_CUDA(cudaGetLastError());
  return ERN_OK;
}

size_t workspace_bytes(int64_t rows, int patches, int dim, int mode) {
  const size_t r = rows, P = patches, d = dim;
  if (mode == ERN_MODE_FP32) return al(r * d * 4) * 2 + al(r * P * cdiv(dim, kTile) * 4) + 512;
  return al(r * P * d * 2) + al(r * d * 2) + al(r * d * 4) + al(r * P * gemmtc::n_tiles_of<kTcBlockN>(dim) * 4) + 512;
}

int forward(const ern_visualsr_weights* w, int dim, int patches, int mode, const float* local, int64_t rows,
            float* out, void* workspace, int sm_count, cudaStream_t st) {
  if (rows <= 0) return ERN_OK;
  const size_t r = rows, P = patches, d = dim;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const int warp_blocks = cdiv(rows * 32, 256);
  if (mode == ERN_MODE_FP32) {
    float* mean = reinterpret_cast<float*>(ws);
    float* cvec = reinterpret_cast<float*>(ws + al(r * d * 4));
    float* partial = reinterpret_cast<float*>(ws + 2 * al(r * d * 4));
    const int n_tiles = cdiv(dim, kTile);
    launch_prepare_sr(local, rows, patches, dim, mean, nullptr, nullptr, st);
    ERN_REQUIRE(cdiv(rows * patches, kTile) <= 65535, "too many rows for one fp32 VisualSR call; split the batch");
    linear_tanh_f32_kernel<false><<<dim3(n_tiles, cdiv(rows, kTile)), kF32Threads, 0, st>>>(
        mean, rows, w->w_global, dim, dim, w->b_global, w->bn_global_scale, w->bn_global_shift, w->w_common, cvec,
        nullptr, patches, nullptr, 0);
    linear_tanh_f32_kernel<true><<<dim3(n_tiles, cdiv(rows * patches, kTile)), kF32Threads, 0, st>>>(
        local, rows * patches, w->w_local, dim, dim, w->b_local, w->bn_local_scale, w->bn_local_shift, nullptr,
        nullptr, cvec, patches, partial, n_tiles);
    launch_finalize_sr(local, rows, patches, dim, partial, n_tiles, w->b_common, out, st);
    ERN_CUDA(cudaGetLastError());
    return ERN_OK;
  }
  // ---- tensor-core path
  __half* local_b = reinterpret_cast<__half*>(ws);            // fp16 operands (see pack_f16x2)
  __half* mean_b = reinterpret_cast<__half*>(ws + al(r * P * d * 2));
  float* cvec = reinterpret_cast<float*>(ws + al(r * P * d * 2) + al(r * d * 2));
  float* partial = reinterpret_cast<float*>(ws + al(r * P * d * 2) + al(r * d * 2) + al(r * d * 4));
  const int n_tiles = gemmtc::n_tiles_of<kTcBlockN>(dim);
  const uint8_t* pk = static_cast<const uint8_t*>(w->packed_bf16);
  const void* wl_b = pk;
  const void* wg_b = pk + al(d * d * 2);
  launch_prepare_sr(local, rows, patches, dim, nullptr, mean_b, local_b, st);
  ERN_CUDA(cudaGetLastError());
  CUtensorMap t_local, t_mean, t_wl, t_wg;
  int rc;
  if ((rc = simtc::make_tmap_bf16_rows(&t_local, local_b, rows * patches, dim, dim))) return rc;
  if ((rc = simtc::make_tmap_bf16_rows(&t_mean, mean_b, rows, dim, dim))) return rc;
  if ((rc = simtc::make_tmap_bf16_rows(&t_wl, wl_b, dim, dim, dim))) return rc;
  if ((rc = simtc::make_tmap_bf16_rows(&t_wg, wg_b, dim, dim, dim))) return rc;
  gemmtc::Params g{};
  g.m = rows;
  g.n = dim;
  g.k = dim;
  g.bias = w->b_global;
  g.wg = w->w_common;
  g.scale = w->bn_global_scale;
  g.shift = w->bn_global_shift;
  g.out_f32 = cvec;
  g.ldo = dim;
  g.f16_operands = 1;
  if ((rc = gemmtc::launch<kTcBlockN, gemmtc::kEpiSrGlobal, true>(t_mean, t_wg, g, sm_count, st))) return rc;
  gemmtc::Params l{};
  l.m = rows * patches;
  l.n = dim;
  l.k = dim;
  l.bias = w->b_local;
  l.scale = w->bn_local_scale;
  l.shift = w->bn_local_shift;
  l.cvec = cvec;
  l.patches = patches;
  l.partial = partial;
  l.f16_operands = 1;
  if ((rc = gemmtc::launch<kTcBlockN, gemmtc::kEpiSrLocal, true>(t_local, t_wl, l, sm_count, st))) return rc;
  launch_finalize_sr(local, rows, patches, dim, partial, n_tiles, w->b_common, out, st);
  ERN_CUDA(cudaGetLastError());
  return ERN_OK;
}

}  // namespace visualsr
}  // namespace ern
