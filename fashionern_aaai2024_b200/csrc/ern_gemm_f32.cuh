// fp32 FFMA GEMM   out[r, n] = act(X[r,:] . W[n,:] + bias[n]) (+ residual[r, n])   -- validation-mode building block
// (64 x 64 tiles, 4 x 4 per thread, sequential-k fma accumulation).  Not a performance path.
#pragma once
#include "ern_internal.cuh"

namespace ern {
namespace gemmf32 {

constexpr int kTile = 64;
constexpr int kKc = 16;
constexpr int kThreads = 256;
enum Act { kActNone = 0, kActGelu = 1 };

template <int kAct>
__global__ void __launch_bounds__(kThreads)
linear_f32_kernel(const float* __restrict__ X, int64_t ldx, int64_t rows, const float* __restrict__ W, int K, int N,
                  const float* __restrict__ bias, const float* __restrict__ residual, float* __restrict__ out,
                  int64_t ldo, int64_t ldw) {
  __shared__ float xs[kKc][kTile + 1];
  __shared__ float ws[kKc][kTile + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * kTile;
  const int n0 = blockIdx.x * kTile;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += kKc) {
    for (int e = threadIdx.x; e < kTile * kKc; e += kThreads) {
      const int r = e / kKc, kk = e % kKc;
      const bool kin = (k0 + kk) < K;
      xs[kk][r] = (kin && r0 + r < rows) ? X[(r0 + r) * ldx + k0 + kk] : 0.f;
      ws[kk][r] = (kin && n0 + r < N) ? W[static_cast<int64_t>(n0 + r) * ldw + k0 + kk] : 0.f;
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
    if (r >= rows) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      if (kAct == kActGelu) v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
      if (residual) v += residual[r * ldo + n];
      out[r * ldo + n] = v;
    }
  }
}

// rows may exceed 65535 * 64: the row dimension is walked in slabs
template <int kAct>
static int launch(const float* X, int64_t ldx, int64_t rows, const float* W, int K, int N, const float* bias,
                  const float* residual, float* out, int64_t ldo, cudaStream_t st, int64_t ldw = 0) {
  if (ldw == 0) ldw = K;
  const int64_t slab = 65535ll * kTile;
  for (int64_t s = 0; s < rows; s += slab) {
    const int64_t r = rows - s < slab ? rows - s : slab;
    dim3 grid(cdiv(N, kTile), cdiv(r, kTile));
    linear_f32_kernel<kAct><<<grid, kThreads, 0, st>>>(X + s * ldx, ldx, r, W, K, N, bias,
                                                      residual ? residual + s * ldo : nullptr, out + s * ldo, ldo, ldw);
  }
  ERN_CUDA(cudaGetLastError());
  return ERN_OK;
}

}  // namespace gemmf32
}  // namespace ern
