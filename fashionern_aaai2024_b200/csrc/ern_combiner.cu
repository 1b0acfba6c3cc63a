// Fusion head (CombinerSimple.forward, models/fusion_model.py:86-94), fp32 validation path and the pieces
// shared with the tensor-core path: weight packing, the gate/blend/normalise finaliser.
//
//   tp  = relu(text  Wt^T + bt)                 [B,4D]      (:87)
//   ip  = relu(image Wi^T + bi)                 [B,4D]      (:88)
//   raw = [tp | ip]                             [B,8D]      (:90)  -- the two GEMMs write the two halves
//   h   = relu(raw W1^T + b1)                   [B,8D]      (:74-75) -- never stored: the epilogue reduces
//   z   = h . w2 + b2 ; s = sigmoid(z)          [B]         (:77-78)    it against w2 per 64/128-column tile
//   out = normalize(s*text + (1-s)*image)       [B,D]       (:93-94) -- from the fp32 inputs
#include "ern_internal.cuh"

namespace ern {
namespace combiner {

constexpr int kTile = 64;
constexpr int kKc = 16;
constexpr int kThreads = 256;

// out[r, col0 + n] = relu(X[r,:] . W[n,:] + bias[n])            (kGate == false)
// partial[r, blockIdx.x] = sum_{n in tile} relu(...) * wg[n]     (kGate == true)
template <bool kGate>
__global__ void __launch_bounds__(kThreads)
linear_relu_f32_kernel(const float* __restrict__ X, int64_t ldx, int64_t rows, const float* __restrict__ W, int K,
                       int N, const float* __restrict__ bias, float* __restrict__ out, int64_t ldo, int col0,
                       const float* __restrict__ wg, float* __restrict__ partial, int n_tiles) {
  __shared__ float xs[kKc][kTile + 1];
  __shared__ float ws[kKc][kTile + 1];
  const int tx = threadIdx.x & 15;  // output-column direction
  const int ty = threadIdx.x >> 4;  // row direction
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * kTile;
  const int n0 = blockIdx.x * kTile;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += kKc) {
    for (int e = threadIdx.x; e < kTile * kKc; e += kThreads) {
      const int r = e / kKc, kk = e % kKc;
      const bool kin = (k0 + kk) < K;
      xs[kk][r] = (kin && r0 + r < rows) ? X[(r0 + r) * ldx + k0 + kk] : 0.f;
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
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < N) {
        const float v = fmaxf(acc[i][j] + bias[n], 0.f);
        if (kGate) dot = fmaf(v, wg[n], dot);
        else if (r < rows) out[r * ldo + col0 + n] = v;
      }
    }
    if (kGate) {
      // the 16 threads sharing `ty` are one half-warp: fixed-order butterfly => deterministic
      for (int o = 8; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      if (tx == 0 && r < rows) partial[r * n_tiles + blockIdx.x] = dot;
    }
  }
}

// One warp per row: z = b2 + sum(partials), s = sigmoid(z), out = normalize(s*text + (1-s)*image).
__global__ void finalize_kernel(const float* __restrict__ image, const float* __restrict__ text, int64_t rows, int dim,
                                const float* __restrict__ partial, int n_tiles, const float* __restrict__ b_gate,
                                float* __restrict__ out, __nv_bfloat16* __restrict__ out_bf16, int64_t ldb,
                                float* __restrict__ gate) {
  const int64_t r = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  float z = 0.f;
  for (int t = lane; t < n_tiles; t += 32) z += partial[r * n_tiles + t];
  for (int o = 16; o > 0; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
  z += b_gate[0];
  const float s = 1.0f / (1.0f + expf(-z));
  const float* im = image + r * dim;
  const float* tx = text + r * dim;
  float ss = 0.f;
  for (int d = lane; d < dim; d += 32) {
    const float v = s * tx[d] + (1.0f - s) * im[d];
    ss = fmaf(v, v, ss);
  }
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float denom = fmaxf(sqrtf(ss), 1e-12f);
  for (int d = lane; d < dim; d += 32) {
    const float v = (s * tx[d] + (1.0f - s) * im[d]) / denom;
    if (out) out[r * dim + d] = v;
    if (out_bf16) out_bf16[r * ldb + d] = __float2bfloat16_rn(v);
  }
  if (gate && lane == 0) gate[r] = s;
}

__global__ void cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n) dst[i] = __float2bfloat16_rn(src[i]);
}

// ---- packed bf16 weights: [Wt 4DxD][Wi 4DxD][W1 8Dx8D] bf16, then fp32 [bt 4D][bi 4D][b1 8D][w2 8D][b2 1] ------------
size_t packed_bytes(int dim) {
  const size_t d = dim;
  size_t bf = (2 * 4 * d * d + 64 * d * d) * 2;
  bf = (bf + 255) & ~size_t(255);
  return bf + (4 * d + 4 * d + 8 * d + 8 * d + 1) * 4 + 256;
}
PackedView view_packed(const void* packed, int dim) {
  const size_t d = dim;
  PackedView v;
  const __nv_bfloat16* b = static_cast<const __nv_bfloat16*>(packed);
  v.wt = b;
  v.wi = b + 4 * d * d;
  v.w1 = b + 8 * d * d;
  size_t bf = (2 * 4 * d * d + 64 * d * d) * 2;
  bf = (bf + 255) & ~size_t(255);
  const float* f = reinterpret_cast<const float*>(static_cast<const uint8_t*>(packed) + bf);
  v.bt = f;
  v.bi = f + 4 * d;
  v.b1 = f + 8 * d;
  v.w2 = f + 16 * d;
  v.b2 = f + 24 * d;
  // 15 sets of 4 zeroed words behind the vectors: grid-barrier counters of the small-batch kernel (ern_combiner_small.cu)
  v.sync = reinterpret_cast<unsigned*>(const_cast<float*>(f + ((24 * d + 1 + 3) & ~size_t(3))));
  return v;
}

int pack(const ern_combiner_weights* w, int dim, void* packed, cudaStream_t st) {
  const int64_t d = dim;
  PackedView v = view_packed(packed, dim);
  auto cast = [&](const float* src, const __nv_bfloat16* dst, int64_t n) {
    cast_bf16_kernel<<<cdiv(n, 256), 256, 0, st>>>(src, const_cast<__nv_bfloat16*>(dst), n);
  };
  cast(w->w_text, v.wt, 4 * d * d);
  cast(w->w_image, v.wi, 4 * d * d);
  cast(w->w_hid, v.w1, 64 * d * d);
  ERN_CUDA(cudaGetLastError());
  ERN_CUDA(cudaMemcpyAsync(const_cast<float*>(v.bt), w->b_text, 4 * d * 4, cudaMemcpyDeviceToDevice, st));
  ERN_CUDA(cudaMemcpyAsync(const_cast<float*>(v.bi), w->b_image, 4 * d * 4, cudaMemcpyDeviceToDevice, st));
  ERN_CUDA(cudaMemcpyAsync(const_cast<float*>(v.b1), w->b_hid, 8 * d * 4, cudaMemcpyDeviceToDevice, st));
  ERN_CUDA(cudaMemcpyAsync(const_cast<float*>(v.w2), w->w_gate, 8 * d * 4, cudaMemcpyDeviceToDevice, st));
  ERN_CUDA(cudaMemcpyAsync(const_cast<float*>(v.b2), w->b_gate, 4, cudaMemcpyDeviceToDevice, st));
  ERN_CUDA(cudaMemsetAsync(v.sync, 0, 240, st));
  return ERN_OK;
}

int launch_finalize(const float* image, const float* text, int64_t rows, int dim, const float* partial, int n_tiles,
                    const float* b_gate, float* out, void* out_bf16, int64_t ldb, float* gate, cudaStream_t st) {
  if (rows <= 0) return ERN_OK;
  finalize_kernel<<<cdiv(rows * 32, 256), 256, 0, st>>>(image, text, rows, dim, partial, n_tiles, b_gate, out,
                                                        static_cast<__nv_bfloat16*>(out_bf16), ldb, gate);
  ERN_CUDA(cudaGetLastError());
  return ERN_OK;
}

// 8 elements per thread: two 16-byte loads, one 16-byte store (the scalar form above ran at 2.4 TB/s)
__global__ void cast_bf16_vec8_kernel(const float4* __restrict__ src, uint4* __restrict__ dst, int64_t n8) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n8) return;
  const float4 a = src[2 * i], b = src[2 * i + 1];
  __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
  __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
  dst[i] = make_uint4(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1),
                      *reinterpret_cast<uint32_t*>(&p2), *reinterpret_cast<uint32_t*>(&p3));
}

int launch_cast_bf16(const float* src, void* dst, int64_t n, cudaStream_t st) {
  if (n <= 0) return ERN_OK;
  const bool vec = n % 8 == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0;
  if (vec)
    cast_bf16_vec8_kernel<<<cdiv(n / 8, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(src),
                                                           static_cast<uint4*>(dst), n / 8);
  else
    cast_bf16_kernel<<<cdiv(n, 256), 256, 0, st>>>(src, static_cast<__nv_bfloat16*>(dst), n);
  ERN_CUDA(cudaGetLastError());
  return ERN_OK;
}

// ---- fp32 forward --------------------------------------------------------------------------------------------------
size_t workspace_bytes_f32(int64_t rows, int dim) {
  const size_t raw = static_cast<size_t>(rows) * 8 * dim * 4;
  const size_t part = static_cast<size_t>(rows) * cdiv(8 * dim, kTile) * 4;
  return raw + part + 512;
}

int forward_f32(const ern_combiner_weights* w, int dim, const float* image, const float* text, int64_t rows,
                float* out, void* out_bf16, int64_t ldb, float* gate, void* workspace, cudaStream_t st) {
  if (rows <= 0) return ERN_OK;
  const int proj = 4 * dim, hid = 8 * dim;
  float* raw = static_cast<float*>(workspace);
  float* partial = raw + static_cast<size_t>(rows) * hid;
  const int rblocks = cdiv(rows, kTile);
  ERN_REQUIRE(rblocks <= 65535, "too many rows for one fp32 combiner call (%lld); split the batch", (long long)rows);
  dim3 g1(cdiv(proj, kTile), rblocks);
  linear_relu_f32_kernel<false><<<g1, kThreads, 0, st>>>(text, dim, rows, w->w_text, dim, proj, w->b_text, raw, hid, 0,
                                                         nullptr, nullptr, 0);
  linear_relu_f32_kernel<false><<<g1, kThreads, 0, st>>>(image, dim, rows, w->w_image, dim, proj, w->b_image, raw, hid,
                                                         proj, nullptr, nullptr, 0);
  const int n_tiles = cdiv(hid, kTile);
  dim3 g2(n_tiles, rblocks);
  linear_relu_f32_kernel<true><<<g2, kThreads, 0, st>>>(raw, hid, rows, w->w_hid, hid, hid, w->b_hid, nullptr, 0, 0,
                                                        w->w_gate, partial, n_tiles);
  ERN_CUDA(cudaGetLastError());
  return launch_finalize(image, text, rows, dim, partial, n_tiles, w->b_gate, out, out_bf16, ldb, gate, st);
}

}  // namespace combiner
}  // namespace ern
