// In-batch classification loss and its gradient  (BatchBasedClassificationLoss, losses/loss.py:6-14)
//
//   logits x = scale * P . T^T  (scale = 100, :11),   labels = arange(B) (:12),   loss = mean_i (lse_i - x_ii)  (:14)
//   d loss / d x_ij = (softmax_ij - [i == j]) / B
//   dP = scale * dX . T          dT = scale * dX^T . P
//
// bf16 mode (tensor cores): the logits never reach HBM.  Forward = the tcgen05 GEMM of ern_gemm_tc.cuh with a
// logsumexp epilogue (per 256-column tile partial max / sum-exp per row + the diagonal), then a finaliser.  Backward
// recomputes the logits twice with softmax-gradient epilogues that emit dX (row logsumexp) and dX^T (column
// logsumexp: the same GEMM with the operands swapped) as bf16, and feeds them to two plain GEMMs against the
// transposed operands.  fp32 mode (validation): FFMA GEMMs with the logits materialised.
#include "ern_gemm_f32.cuh"
#include "ern_gemm_tc.cuh"

namespace ern {
namespace bbcloss {

static size_t al(size_t x) { return (x + 255) & ~size_t(255); }
constexpr int kBlockN = 256;

// both operands in one launch (blockIdx.z): src fp32 [rows, dim] -> dst bf16 [rows, dim] and/or dst_t bf16 [dim, kp]
// (columns rows..kp-1 of the transposed copy are zero: they are the K padding of the gradient GEMMs)
struct CastJob {
  const float* src;
  int64_t lds;
  __nv_bfloat16 *dst, *dst_t;
};
__global__ void __launch_bounds__(256)
cast_transpose_kernel(CastJob j0, CastJob j1, int64_t rows, int dim, int64_t kp) {
  __shared__ float tile[32][33];
  const CastJob job = blockIdx.z ? j1 : j0;
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int64_t r = r0 + i;
    const int c = c0 + tx;
    const float v = (r < rows && c < dim) ? job.src[r * job.lds + c] : 0.f;
    tile[i][tx] = v;
    if (job.dst && r < rows && c < dim) job.dst[r * dim + c] = __float2bfloat16(v);
  }
  __syncthreads();
  if (job.dst_t) {
    for (int i = ty; i < 32; i += 8) {
      const int c = c0 + i;
      const int64_t r = r0 + tx;
      if (c < dim && r < kp) job.dst_t[static_cast<int64_t>(c) * kp + r] = __float2bfloat16(tile[tx][i]);
    }
  }
}

// fp32 transpose  src [rows, cols] (ld) -> dst [cols, rows]
__global__ void __launch_bounds__(256)
transpose_f32_kernel(const float* __restrict__ src, int64_t lds, int64_t rows, int64_t cols, float* __restrict__ dst) {
  __shared__ float tile[32][33];
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * 32;
  const int64_t c0 = static_cast<int64_t>(blockIdx.y) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8)
    tile[i][tx] = (r0 + i < rows && c0 + tx < cols) ? src[(r0 + i) * lds + c0 + tx] : 0.f;
  __syncthreads();
  for (int i = ty; i < 32; i += 8)
    if (c0 + i < cols && r0 + tx < rows) dst[(c0 + i) * rows + r0 + tx] = tile[tx][i];
}

// deterministic block sum (fixed tree) of one value per thread; result valid in thread 0
__device__ __forceinline__ float block_sum_1024(float v, float* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    v = threadIdx.x < (blockDim.x >> 5) ? scratch[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  }
  return v;
}

// combine the per-tile (max, sum-exp) partials [tile, rows] -> lse[r], row_loss[r] = lse[r] - diag[r]
__global__ void __launch_bounds__(256)
lse_combine_kernel(const float* __restrict__ pmax, const float* __restrict__ psum, int n_tiles,
                   const float* __restrict__ diag, int64_t rows, float* __restrict__ lse, float* __restrict__ row_loss) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float m = -INFINITY;
  for (int t = 0; t < n_tiles; ++t) m = fmaxf(m, pmax[t * rows + r]);
  float s = 0.f;
  for (int t = 0; t < n_tiles; ++t) s += psum[t * rows + r] * expf(pmax[t * rows + r] - m);
  const float l = m + logf(s);
  if (lse) lse[r] = l;
  row_loss[r] = l - diag[r];
}

// fp32 mode: one block per row of the materialised raw scores S (logits = scale * S)
__global__ void __launch_bounds__(256)
row_lse_f32_kernel(const float* __restrict__ S, int64_t lds, int64_t n, float scale, float* __restrict__ lse,
                   float* __restrict__ row_loss) {
  __shared__ float red[32];
  const int64_t r = blockIdx.x;
  const float* row = S + r * lds;
  float m = -INFINITY;
  for (int64_t j = threadIdx.x; j < n; j += blockDim.x) m = fmaxf(m, __fmul_rn(scale, row[j]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = red[0];
  for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  float s = 0.f;
  for (int64_t j = threadIdx.x; j < n; j += blockDim.x) s += expf(__fmul_rn(scale, row[j]) - m);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    const float l = m + logf(t);
    lse[r] = l;
    row_loss[r] = l - __fmul_rn(scale, row[r]);
  }
}

__global__ void __launch_bounds__(1024)
mean_kernel(const float* __restrict__ v, int64_t n, float* __restrict__ out) {
  __shared__ float scratch[32];
  float local = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) local += v[i];
  const float total = block_sum_1024(local, scratch);
  if (threadIdx.x == 0) out[0] = total / static_cast<float>(n);
}

// fp32 mode: dX = (exp(scale * S - lse_row) - I) * coef, written as dX [n, n] and its transpose
__global__ void __launch_bounds__(256)
grad_logits_f32_kernel(const float* __restrict__ S, int64_t lds, int64_t n, float scale, const float* __restrict__ lse,
                       float coef, const float* __restrict__ gscale, float* __restrict__ dx, float* __restrict__ dx_t) {
  __shared__ float tile[32][33];
  const float c = coef * (gscale ? gscale[0] : 1.f);
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * 32;
  const int64_t c0 = static_cast<int64_t>(blockIdx.y) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int64_t r = r0 + i, col = c0 + tx;
    float g = 0.f;
    if (r < n && col < n) {
      g = (expf(__fmul_rn(scale, S[r * lds + col]) - lse[r]) - (r == col ? 1.f : 0.f)) * c;
      dx[r * n + col] = g;
    }
    tile[i][tx] = g;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8)
    if (c0 + i < n && r0 + tx < n) dx_t[(c0 + i) * n + r0 + tx] = tile[tx][i];
}

static int64_t pad64(int64_t b) { return (b + 63) / 64 * 64; }

size_t workspace_bytes(int64_t b, int dim, int mode) {
  if (b <= 0) return 256;
  const size_t B = b, D = dim, kp = pad64(b);
  if (mode == ERN_MODE_FP32) {
    // S [B,B], dX [B,B], dX^T [B,B], P^T [D,B], T^T [D,B], row_loss [B]
    return 3 * al(B * B * 4) + 2 * al(D * B * 4) + al(B * 4) + 512;
  }
  const size_t nt = (B + kBlockN - 1) / kBlockN;
  // P, T bf16 [B,D]; P^T, T^T bf16 [D,kp]; dX, dX^T bf16 [B,kp]; partial max / sum [nt,B]; diag, row_loss [B]
  return 2 * al(B * D * 2) + 2 * al(D * kp * 2) + 2 * al(B * kp * 2) + 2 * al(B * nt * 4) + 2 * al(B * 4) + 512;
}

namespace {
struct Carve {
  __nv_bfloat16 *p, *t, *pt, *tt, *dx, *dxt;
  float *pmax, *psum, *diag, *row_loss;
};
Carve carve_bf16(void* ws, int64_t b, int dim) {
  const size_t B = b, D = dim, kp = pad64(b), nt = (B + kBlockN - 1) / kBlockN;
  uint8_t* w = static_cast<uint8_t*>(ws);
  Carve c;
  c.p = reinterpret_cast<__nv_bfloat16*>(w); w += al(B * D * 2);
  c.t = reinterpret_cast<__nv_bfloat16*>(w); w += al(B * D * 2);
  c.pt = reinterpret_cast<__nv_bfloat16*>(w); w += al(D * kp * 2);
  c.tt = reinterpret_cast<__nv_bfloat16*>(w); w += al(D * kp * 2);
  c.dx = reinterpret_cast<__nv_bfloat16*>(w); w += al(B * kp * 2);
  c.dxt = reinterpret_cast<__nv_bfloat16*>(w); w += al(B * kp * 2);
  c.pmax = reinterpret_cast<float*>(w); w += al(B * nt * 4);
  c.psum = reinterpret_cast<float*>(w); w += al(B * nt * 4);
  c.diag = reinterpret_cast<float*>(w); w += al(B * 4);
  c.row_loss = reinterpret_cast<float*>(w);
  return c;
}
struct CarveF32 {
  float *s, *dx, *dxt, *pt, *tt, *row_loss;
};
CarveF32 carve_f32(void* ws, int64_t b, int dim) {
  const size_t B = b, D = dim;
  uint8_t* w = static_cast<uint8_t*>(ws);
  CarveF32 c;
  c.s = reinterpret_cast<float*>(w); w += al(B * B * 4);
  c.dx = reinterpret_cast<float*>(w); w += al(B * B * 4);
  c.dxt = reinterpret_cast<float*>(w); w += al(B * B * 4);
  c.pt = reinterpret_cast<float*>(w); w += al(D * B * 4);
  c.tt = reinterpret_cast<float*>(w); w += al(D * B * 4);
  c.row_loss = reinterpret_cast<float*>(w);
  return c;
}
int cast_operands(const float* pred, int64_t ldp, const float* tar, int64_t ldt, int64_t b, int dim, const Carve& c,
                  bool transposed, cudaStream_t st) {
  const int64_t kp = pad64(b);
  const CastJob j0{pred, ldp, c.p, transposed ? c.pt : nullptr};
  const CastJob j1{tar, ldt, c.t, transposed ? c.tt : nullptr};
  dim3 grid(cdiv(transposed ? kp : b, 32), cdiv(dim, 32), 2);
  cast_transpose_kernel<<<grid, 256, 0, st>>>(j0, j1, b, dim, kp);
  ERN_CUDA(cudaGetLastError());
  return ERN_OK;
}
}  // namespace

int forward(const float* pred, int64_t ldp, const float* tar, int64_t ldt, int64_t b, int dim, float scale, int mode,
            float* loss, float* lse, void* workspace, int sm_count, cudaStream_t st) {
  if (mode == ERN_MODE_FP32) {
    CarveF32 c = carve_f32(workspace, b, dim);
    int rc = gemmf32::launch<gemmf32::kActNone>(pred, ldp, b, tar, dim, static_cast<int>(b), nullptr, nullptr, c.s, b,
                                                st, ldt);
    if (rc) return rc;
    // lse is needed by backward; when the caller does not want it, it lands in the (unused here) dX scratch
    float* lse_out = lse ? lse : c.dx;
    row_lse_f32_kernel<<<static_cast<unsigned>(b), 256, 0, st>>>(c.s, b, b, scale, lse_out, c.row_loss);
    ERN_CUDA(cudaGetLastError());
    mean_kernel<<<1, 1024, 0, st>>>(c.row_loss, b, loss);
    ERN_CUDA(cudaGetLastError());
    return ERN_OK;
  }
  Carve c = carve_bf16(workspace, b, dim);
  int rc;
  if ((rc = cast_operands(pred, ldp, tar, ldt, b, dim, c, false, st))) return rc;
  CUtensorMap tp, tt;
  if ((rc = simtc::make_tmap_bf16_rows(&tp, c.p, b, dim, dim))) return rc;
  if ((rc = simtc::make_tmap_bf16_rows(&tt, c.t, b, dim, dim))) return rc;
  gemmtc::Params g{};
  g.m = b;
  g.n = static_cast<int>(b);
  g.k = dim;
  g.alpha = scale;
  g.partial = c.pmax;
  g.partial2 = c.psum;
  g.diag = c.diag;
  if ((rc = gemmtc::launch<kBlockN, gemmtc::kEpiLse, true>(tp, tt, g, sm_count, st))) return rc;
  lse_combine_kernel<<<cdiv(b, 256), 256, 0, st>>>(c.pmax, c.psum, gemmtc::n_tiles_of<kBlockN>(static_cast<int>(b)),
                                                  c.diag, b, lse, c.row_loss);
  ERN_CUDA(cudaGetLastError());
  mean_kernel<<<1, 1024, 0, st>>>(c.row_loss, b, loss);
  ERN_CUDA(cudaGetLastError());
  return ERN_OK;
}

int backward(const float* pred, int64_t ldp, const float* tar, int64_t ldt, int64_t b, int dim, float scale, int mode,
             const float* lse, const float* grad_out, float* dpred, int64_t lddp, float* dtar, int64_t lddt,
             void* workspace, int sm_count, cudaStream_t st) {
  // d loss / d logits carries 1/B; the chain rule through logits = scale * P.T^T carries scale
  const float coef = scale / static_cast<float>(b);
  if (mode == ERN_MODE_FP32) {
    CarveF32 c = carve_f32(workspace, b, dim);
    const int n = static_cast<int>(b);
    int rc = gemmf32::launch<gemmf32::kActNone>(pred, ldp, b, tar, dim, n, nullptr, nullptr, c.s, b, st, ldt);
    if (rc) return rc;
    dim3 gg(cdiv(b, 32), cdiv(b, 32));
    grad_logits_f32_kernel<<<gg, 256, 0, st>>>(c.s, b, b, scale, lse, coef, grad_out, c.dx, c.dxt);
    ERN_CUDA(cudaGetLastError());
    dim3 gt(cdiv(b, 32), cdiv(dim, 32));
    transpose_f32_kernel<<<gt, 256, 0, st>>>(pred, ldp, b, dim, c.pt);
    ERN_CUDA(cudaGetLastError());
    transpose_f32_kernel<<<gt, 256, 0, st>>>(tar, ldt, b, dim, c.tt);
    ERN_CUDA(cudaGetLastError());
    if ((rc = gemmf32::launch<gemmf32::kActNone>(c.dx, b, b, c.tt, n, dim, nullptr, nullptr, dpred, lddp, st))) return rc;
    return gemmf32::launch<gemmf32::kActNone>(c.dxt, b, b, c.pt, n, dim, nullptr, nullptr, dtar, lddt, st);
  }
  Carve c = carve_bf16(workspace, b, dim);
  const int64_t kp = pad64(b);
  int rc;
  if ((rc = cast_operands(pred, ldp, tar, ldt, b, dim, c, true, st))) return rc;
  CUtensorMap tp, tt, tpt, ttt, tdx, tdxt;
  if ((rc = simtc::make_tmap_bf16_rows(&tp, c.p, b, dim, dim))) return rc;
  if ((rc = simtc::make_tmap_bf16_rows(&tt, c.t, b, dim, dim))) return rc;
  if ((rc = simtc::make_tmap_bf16_rows(&tpt, c.pt, dim, static_cast<int>(kp), kp))) return rc;
  if ((rc = simtc::make_tmap_bf16_rows(&ttt, c.tt, dim, static_cast<int>(kp), kp))) return rc;
  if ((rc = simtc::make_tmap_bf16_rows(&tdx, c.dx, b, static_cast<int>(kp), kp))) return rc;
  if ((rc = simtc::make_tmap_bf16_rows(&tdxt, c.dxt, b, static_cast<int>(kp), kp))) return rc;
  // dX[i, j]   = (softmax_ij - [i == j]) * coef        rows = queries, lse by row
  gemmtc::Params g{};
  g.m = b;
  g.n = static_cast<int>(b);
  g.k = dim;
  g.alpha = scale;
  g.beta = coef;
  g.gscale = grad_out;
  g.out = c.dx;
  g.ldo = kp;
  g.rowvec = lse;
  if ((rc = gemmtc::launch<kBlockN, gemmtc::kEpiSmGradRow, true>(tp, tt, g, sm_count, st))) return rc;
  // dX^T[j, i] = the same numbers from T . P^T          rows = targets, lse by column
  g.out = c.dxt;
  g.rowvec = nullptr;
  g.bias = lse;
  if ((rc = gemmtc::launch<kBlockN, gemmtc::kEpiSmGradCol, true>(tt, tp, g, sm_count, st))) return rc;
  // dP = dX . T   (A = dX [B, kp], W = T^T [D, kp]);   dT = dX^T . P
  gemmtc::Params d{};
  d.m = b;
  d.n = dim;
  d.k = static_cast<int>(kp);
  d.out_f32 = dpred;
  d.ldo = lddp;
  if ((rc = gemmtc::launch<kBlockN, gemmtc::kEpiResidF32, true>(tdx, ttt, d, sm_count, st))) return rc;
  d.out_f32 = dtar;
  d.ldo = lddt;
  return gemmtc::launch<kBlockN, gemmtc::kEpiResidF32, true>(tdxt, tpt, d, sm_count, st);
}

}  // namespace bbcloss
}  // namespace ern
