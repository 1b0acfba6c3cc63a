// Query-side encoder of DVR_module (models/fusion_model.py:26-49) -- SURVEY.md 8(f) "next" row 2:
//   PlusModel (:187-216): [CLS] + 13 patch embeddings + 77 token embeddings -> HF BERT encoder (2 layers, 8 heads,
//   intermediate 3072, erf-GELU, LayerNorm eps 1e-12, absolute position + token-type embeddings, all-ones mask)
//   F.normalize of the patch / token states (:38-41), nn.MultiheadAttention cross attention text -> patches (:44-46)
//   of which only the first 13 query positions are used (:47), and the mean of the normalised token states (:49).
// Outputs: cross_vision_feats[:, :13] (input of SR_module) and seq_text_mean (input of combiner_local).
// Every GEMM runs on the tensor-core kernel of ern_gemm_tc.cuh (bf16 mode) or the FFMA kernel of ern_gemm_f32.cuh
// (fp32 validation mode); embeddings+LayerNorm, attention (L = 91 per head: CUDA cores, fp32 math), row-normalise and
// LayerNorm are small dedicated kernels.  Eval mode only (dropouts are identities).
#include "ern_gemm_f32.cuh"
#include <initializer_list>

#include "ern_gemm_tc.cuh"

namespace ern {
namespace dvr {

constexpr int kMaxPerLane = 32;  // D <= 1024
constexpr float kLnEps = 1e-12f;

__device__ __forceinline__ float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// LayerNorm of one row held by a warp; two-pass (mean, then centred variance) like torch.
// Element layout: kVec (dim % 128 == 0, 16-byte aligned rows): v[4 j + e] = x[4 lane + 128 j + e] -- float4 loads and
// stores, 4x fewer memory instructions (the scalar form issued ~100 per row and ran at 4 TB/s);  otherwise
// v[i] = x[lane + 32 i].
template <bool kVec> __device__ __forceinline__ int ln_index(int lane, int i) {
  return kVec ? 4 * lane + 128 * (i >> 2) + (i & 3) : lane + 32 * i;
}
template <bool kVec>
__device__ __forceinline__ void layer_norm_row(float (&v)[kMaxPerLane], int dim, int lane, const float* w,
                                               const float* b, float* out, __nv_bfloat16* out_b) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i)
    if (ln_index<kVec>(lane, i) < dim) s += v[i];
  const float mean = warp_sum(s) / static_cast<float>(dim);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i)
    if (ln_index<kVec>(lane, i) < dim) {
      const float c = v[i] - mean;
      ss = fmaf(c, c, ss);
    }
  const float rstd = rsqrtf(warp_sum(ss) / static_cast<float>(dim) + kLnEps);
  if (kVec) {
#pragma unroll
    for (int j = 0; j < kMaxPerLane / 4; ++j) {
      const int d = 4 * lane + 128 * j;
      if (d < dim) {
        const float4 w4 = *reinterpret_cast<const float4*>(w + d);
        const float4 b4 = *reinterpret_cast<const float4*>(b + d);
        float4 y;
        y.x = (v[4 * j + 0] - mean) * rstd * w4.x + b4.x;
        y.y = (v[4 * j + 1] - mean) * rstd * w4.y + b4.y;
        y.z = (v[4 * j + 2] - mean) * rstd * w4.z + b4.z;
        y.w = (v[4 * j + 3] - mean) * rstd * w4.w + b4.w;
        if (out) *reinterpret_cast<float4*>(out + d) = y;
        if (out_b) {
          __nv_bfloat162 lo = __floats2bfloat162_rn(y.x, y.y), hi = __floats2bfloat162_rn(y.z, y.w);
          uint2 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&lo);
          pk.y = *reinterpret_cast<uint32_t*>(&hi);
          *reinterpret_cast<uint2*>(out_b + d) = pk;
        }
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i) {
      const int d = lane + 32 * i;
      if (d < dim) {
        const float y = (v[i] - mean) * rstd * w[d] + b[d];
        if (out) out[d] = y;
        if (out_b) out_b[d] = __float2bfloat16_rn(y);
      }
    }
  }
}

// one warp per token row: inputs_embeds + token_type + position -> LayerNorm  (HF BertEmbeddings with inputs_embeds)
template <bool kVec>
__global__ void embed_ln_kernel(const float* __restrict__ patches, const float* __restrict__ tokens,
                                const float* __restrict__ cls, const float* __restrict__ pos,
                                const float* __restrict__ type, const float* __restrict__ w,
                                const float* __restrict__ b, int64_t batch, int P, int T, int dim,
                                float* __restrict__ X, __nv_bfloat16* __restrict__ Xb) {
  const int L = 1 + P + T;
  const int64_t r = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= batch * L) return;
  const int64_t bi = r / L;
  const int t = static_cast<int>(r % L);
  const float* src = t == 0 ? cls : (t <= P ? patches + (bi * P + (t - 1)) * dim : tokens + (bi * T + (t - 1 - P)) * dim);
  const float* ty = type + (t <= P ? 0 : dim);
  const float* ps = pos + static_cast<int64_t>(t) * dim;
  float v[kMaxPerLane];
  if (kVec) {
#pragma unroll
    for (int j = 0; j < kMaxPerLane / 4; ++j) {
      const int d = 4 * lane + 128 * j;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (d < dim) {
        const float4 a = *reinterpret_cast<const float4*>(src + d);
        const float4 c = *reinterpret_cast<const float4*>(ty + d);
        const float4 e = *reinterpret_cast<const float4*>(ps + d);
        x = make_float4(a.x + c.x + e.x, a.y + c.y + e.y, a.z + c.z + e.z, a.w + c.w + e.w);
      }
      v[4 * j + 0] = x.x;
      v[4 * j + 1] = x.y;
      v[4 * j + 2] = x.z;
      v[4 * j + 3] = x.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i) {
      const int d = lane + 32 * i;
      v[i] = d < dim ? src[d] + ty[d] + ps[d] : 0.f;
    }
  }
  layer_norm_row<kVec>(v, dim, lane, w, b, X + r * dim, Xb ? Xb + r * dim : nullptr);
}

// out = LayerNorm(in + res) (res nullable).  In bf16 mode the residual is added HERE rather than in the epilogue of
// the preceding GEMM: a streaming kernel reads it at full bandwidth, whereas in the epilogue its global loads sat on
// the critical path of every 32-column chunk.  Same rounding sequence either way ((acc + bias) then + residual).
template <bool kVec>
__global__ void layernorm_kernel(const float* __restrict__ in, const float* res, int64_t rows, int dim,
                                 const float* __restrict__ w, const float* __restrict__ b, float* out,
                                 __nv_bfloat16* __restrict__ out_b) {   // res may alias out (in-place residual stream)
  const int64_t r = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  float v[kMaxPerLane];
  if (kVec) {
#pragma unroll
    for (int j = 0; j < kMaxPerLane / 4; ++j) {
      const int d = 4 * lane + 128 * j;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (d < dim) {
        x = *reinterpret_cast<const float4*>(in + r * dim + d);
        if (res) {
          const float4 y = *reinterpret_cast<const float4*>(res + r * dim + d);
          x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
        }
      }
      v[4 * j + 0] = x.x;
      v[4 * j + 1] = x.y;
      v[4 * j + 2] = x.z;
      v[4 * j + 3] = x.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i) {
      const int d = lane + 32 * i;
      v[i] = d < dim ? in[r * dim + d] + (res ? res[r * dim + d] : 0.f) : 0.f;
    }
  }
  layer_norm_row<kVec>(v, dim, lane, w, b, out + r * dim, out_b ? out_b + r * dim : nullptr);
}

// host-side launchers: the vector form needs dim % 128 == 0 and 16-byte aligned pointers
static bool aligned16(std::initializer_list<const void*> ptrs) {
  for (const void* q : ptrs)
    if (q && (reinterpret_cast<uintptr_t>(q) & 15u)) return false;
  return true;
}
static void launch_layernorm(const float* in, const float* res, int64_t rows, int dim, const float* w, const float* b,
                             float* out, __nv_bfloat16* out_b, cudaStream_t st) {
  const int blocks = cdiv(rows * 32, 256);
  if (dim % 128 == 0 && aligned16({in, res, w, b, out, out_b}))
    layernorm_kernel<true><<<blocks, 256, 0, st>>>(in, res, rows, dim, w, b, out, out_b);
  else
    layernorm_kernel<false><<<blocks, 256, 0, st>>>(in, res, rows, dim, w, b, out, out_b);
}

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// softmax(Q K^T * scale) V for one (batch element, head); sequences are short (<= 91), everything lives in smem
constexpr int kAttnThreads = 512;   // 16 warps on the one block that fits per SM: hides the smem/FMA latencies

template <typename T>
__global__ void __launch_bounds__(kAttnThreads)
attention_kernel(const T* __restrict__ Q, int64_t ldq, const T* __restrict__ K, int64_t ldk, const T* __restrict__ V,
                 int64_t ldv, T* __restrict__ O, int64_t ldo, int Lq, int Lk, int dh, float scale) {
  extern __shared__ float sm[];
  const int h = blockIdx.x;
  const int64_t b = blockIdx.y;
  const int pitch = dh + 1;
  float* Qs = sm;
  float* Ks = Qs + Lq * pitch;
  float* Vs = Ks + Lk * pitch;
  float* S = Vs + Lk * pitch;   // [Lq][Lk + 1]
  const int sp = Lk + 1;
  const int tid = threadIdx.x;
  // head slices are contiguous (dh elements per row): 16-byte vector loads, converted to fp32 in shared memory
  constexpr int kVec = 16 / sizeof(T);
  const int vpr = dh / kVec;   // vectors per row (dh % kVec == 0 is checked by the launcher)
  auto load_rows = [&](const T* src, int64_t ld, int rows, float* dst) {
    for (int e = tid; e < rows * vpr; e += kAttnThreads) {
      const int r = e / vpr, v = e - r * vpr;
      const uint4 raw = *reinterpret_cast<const uint4*>(src + (b * rows + r) * ld + h * dh + v * kVec);
      const T* vals = reinterpret_cast<const T*>(&raw);
#pragma unroll
      for (int u = 0; u < kVec; ++u) dst[r * pitch + v * kVec + u] = to_f(vals[u]);
    }
  };
  load_rows(Q, ldq, Lq, Qs);
  load_rows(K, ldk, Lk, Ks);
  load_rows(V, ldv, Lk, Vs);
  __syncthreads();
  // S = scale * Q K^T with 4 x 4 register tiles.  A thread's rows/columns are STRIDED (i = ib + nbi*ii, j = jb + nbj*jj)
  // so that consecutive lanes read consecutive rows of K: pitch dh+1 is odd => conflict-free; Q reads broadcast.
  {
    const int nbi = (Lq + 3) >> 2, nbj = (Lk + 3) >> 2;
    for (int blk = tid; blk < nbi * nbj; blk += kAttnThreads) {
      const int ib = blk / nbj, jb = blk % nbj;
      float acc[4][4] = {};
      const float* qp[4];
      const float* kp[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        qp[u] = Qs + min(ib + nbi * u, Lq - 1) * pitch;
        kp[u] = Ks + min(jb + nbj * u, Lk - 1) * pitch;
      }
      for (int d = 0; d < dh; ++d) {
        float qv[4], kv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          qv[u] = qp[u][d];
          kv[u] = kp[u][d];
        }
#pragma unroll
        for (int ii = 0; ii < 4; ++ii)
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) acc[ii][jj] = fmaf(qv[ii], kv[jj], acc[ii][jj]);
      }
#pragma unroll
      for (int ii = 0; ii < 4; ++ii)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int i = ib + nbi * ii, j = jb + nbj * jj;
          if (i < Lq && j < Lk) S[i * sp + j] = acc[ii][jj] * scale;
        }
    }
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  for (int i = warp; i < Lq; i += kAttnThreads / 32) {
    float mx = -INFINITY;
    for (int j = lane; j < Lk; j += 32) mx = fmaxf(mx, S[i * sp + j]);
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int j = lane; j < Lk; j += 32) {
      const float ev = expf(S[i * sp + j] - mx);
      S[i * sp + j] = ev;
      sum += ev;
    }
    sum = warp_sum(sum);
    for (int j = lane; j < Lk; j += 32) S[i * sp + j] = S[i * sp + j] / sum;
  }
  __syncthreads();
  // O = P V with 4 x 4 register tiles: rows strided as above, columns d = db + nbd*dd (consecutive lanes ->
  // consecutive d: conflict-free V reads; P reads broadcast)
  {
    const int nbi = (Lq + 3) >> 2, nbd = (dh + 3) >> 2;
    for (int blk = tid; blk < nbi * nbd; blk += kAttnThreads) {
      const int ib = blk / nbd, db = blk % nbd;
      float acc[4][4] = {};
      const float* pp[4];
      int dcol[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        pp[u] = S + min(ib + nbi * u, Lq - 1) * sp;
        dcol[u] = min(db + nbd * u, dh - 1);
      }
      for (int j = 0; j < Lk; ++j) {
        float pv[4], vv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          pv[u] = pp[u][j];
          vv[u] = Vs[j * pitch + dcol[u]];
        }
#pragma unroll
        for (int ii = 0; ii < 4; ++ii)
#pragma unroll
          for (int dd = 0; dd < 4; ++dd) acc[ii][dd] = fmaf(pv[ii], vv[dd], acc[ii][dd]);
      }
#pragma unroll
      for (int ii = 0; ii < 4; ++ii)
#pragma unroll
        for (int dd = 0; dd < 4; ++dd) {
          const int i = ib + nbi * ii, d = db + nbd * dd;
          if (i < Lq && d < dh) O[(b * Lq + i) * ldo + h * dh + d] = from_f<T>(acc[ii][dd]);
        }
    }
  }
}

// ---- bf16 attention on the (legacy) warp-level tensor path -------------------------------------------------------
// One block per (batch element, head); warp w owns score rows [16w, 16w+16).  S = Q K^T and O = P V are
// mma.sync.m16n8k16 (bf16 x bf16 -> fp32); the softmax runs on the accumulator fragments and its output is re-used
// directly as the A fragments of the second product (no shared-memory round trip).  L <= 96, head size 64 or 80.
// tcgen05 is not used here: a 91 x 91 x 80 problem per head cannot fill a 128-row UMMA tile.
constexpr int kAttnLmax = 96;

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// ldmatrix: four / two 8x8 bf16 matrices from shared memory in the mma fragment layout (lane l supplies the address
// of row l % 8 of matrix l / 8); .trans hands out the transposed fragments, i.e. B fragments from a row-major [k][n] tile
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_row) {
  const uint32_t addr = static_cast<uint32_t>(__cvta_generic_to_shared(smem_row));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
  const uint32_t addr = static_cast<uint32_t>(__cvta_generic_to_shared(smem_row));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ uint32_t pack2_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

template <int kDh>
__global__ void __launch_bounds__(192)
attention_mma_kernel(const __nv_bfloat16* __restrict__ Q, int64_t ldq, const __nv_bfloat16* __restrict__ K, int64_t ldk,
                     const __nv_bfloat16* __restrict__ V, int64_t ldv, __nv_bfloat16* __restrict__ O, int64_t ldo,
                     int Lq, int Lk, float scale) {
  // Q, K and V head slices are staged row-major with a pitch of kDh + 8 bf16 (16-byte rows at a 4-bank skew of
  // 12: every 8-row ldmatrix phase is bank-conflict free).  All fragments come from ldmatrix: x4 for the A (Q)
  // and B (K) operands of Q.K^T, x4.trans for the B operand of P.V straight from row-major V -- no transposed copy
  // (the earlier scalar transposing stores, 8 per loaded vector, were the longest phase of this kernel).
  constexpr int QP = kDh + 8;
  extern __shared__ __align__(16) uint8_t sm_raw[];
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(sm_raw);
  __nv_bfloat16* Ks = Qs + kAttnLmax * QP;
  __nv_bfloat16* Vs = Ks + kAttnLmax * QP;
  const int h = blockIdx.x;
  const int64_t b = blockIdx.y;
  const int tid = threadIdx.x;
  constexpr int kVpr = kDh / 8;
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
  for (int e = tid; e < kAttnLmax * kVpr; e += 192) {
    const int r = e / kVpr, v = e - r * kVpr;
    uint4 q4 = zero4, k4 = zero4, v4 = zero4;
    if (r < Lq) q4 = *reinterpret_cast<const uint4*>(Q + (b * Lq + r) * ldq + h * kDh + v * 8);
    if (r < Lk) {
      k4 = *reinterpret_cast<const uint4*>(K + (b * Lk + r) * ldk + h * kDh + v * 8);
      v4 = *reinterpret_cast<const uint4*>(V + (b * Lk + r) * ldv + h * kDh + v * 8);
    }
    *reinterpret_cast<uint4*>(Qs + r * QP + v * 8) = q4;
    *reinterpret_cast<uint4*>(Ks + r * QP + v * 8) = k4;
    *reinterpret_cast<uint4*>(Vs + r * QP + v * 8) = v4;
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int m0 = warp * 16;
  if (m0 >= Lq) return;
  const int ntiles = (Lk + 7) >> 3;      // 8-column score tiles
  const int ktiles = (Lk + 15) >> 4;     // 16-deep steps of P V
  // per-lane row / column offsets of the ldmatrix addresses
  const int lrow8 = lane & 7, lmat = lane >> 3;     // row within an 8x8 matrix, which of the 4 matrices

  float sacc[12][4];
#pragma unroll
  for (int nt = 0; nt < 12; ++nt)
#pragma unroll
    for (int i = 0; i < 4; ++i) sacc[nt][i] = 0.f;
#pragma unroll
  for (int kk = 0; kk < kDh / 16; ++kk) {
    // A: matrices (rows m0..+7 | m0+8..+15) x (cols kk*16..+7 | +8..+15) -> a[0..3]
    uint32_t a[4];
    ldmatrix_x4(a, Qs + (m0 + lrow8 + (lmat & 1) * 8) * QP + kk * 16 + (lmat >> 1) * 8);
#pragma unroll
    for (int np = 0; np < 6; ++np) {       // two 8-key tiles per ldmatrix.x4
      if (2 * np < ntiles) {
        // B: matrices (keys 16np..+7, d kk*16..+7), (same keys, d +8), (keys 16np+8..+15, d ..), (.., d +8)
        uint32_t bb[4];
        ldmatrix_x4(bb, Ks + (np * 16 + lrow8 + (lmat >> 1) * 8) * QP + kk * 16 + (lmat & 1) * 8);
        mma_bf16_16816(sacc[2 * np], a, bb[0], bb[1]);
        mma_bf16_16816(sacc[2 * np + 1], a, bb[2], bb[3]);
      }
    }
  }
  // softmax over the Lk valid columns of rows (m0+g) [regs 0,1] and (m0+g+8) [regs 2,3]
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < 12; ++nt) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int col = nt * 8 + 2 * t + (i & 1);
      const float v = (nt < ntiles && col < Lk) ? sacc[nt][i] * scale : -INFINITY;
      sacc[nt][i] = v;
      if (i < 2) mx0 = fmaxf(mx0, v); else mx1 = fmaxf(mx1, v);
    }
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < 12; ++nt) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float e = __expf(sacc[nt][i] - (i < 2 ? mx0 : mx1));   // exp(-inf) = 0 for masked columns
      sacc[nt][i] = e;
      if (i < 2) sum0 += e; else sum1 += e;
    }
  }
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
  const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;

  float oacc[kDh / 8][4];
#pragma unroll
  for (int dn = 0; dn < kDh / 8; ++dn)
#pragma unroll
    for (int i = 0; i < 4; ++i) oacc[dn][i] = 0.f;
#pragma unroll
  for (int kt = 0; kt < 6; ++kt) {
    if (kt < ktiles) {
      uint32_t a[4];
      a[0] = pack2_bf16(sacc[2 * kt][0] * inv0, sacc[2 * kt][1] * inv0);
      a[1] = pack2_bf16(sacc[2 * kt][2] * inv1, sacc[2 * kt][3] * inv1);
      a[2] = pack2_bf16(sacc[2 * kt + 1][0] * inv0, sacc[2 * kt + 1][1] * inv0);
      a[3] = pack2_bf16(sacc[2 * kt + 1][2] * inv1, sacc[2 * kt + 1][3] * inv1);
#pragma unroll
      for (int dp = 0; dp < kDh / 16; ++dp) {   // two 8-wide output column tiles per ldmatrix.x4.trans
        // row-major V[key][d]: matrices (keys kt*16..+7 | +8..+15) x (d 16dp..+7 | +8..+15), transposed on load
        uint32_t bb[4];
        ldmatrix_x4_trans(bb, Vs + (kt * 16 + lrow8 + (lmat & 1) * 8) * QP + dp * 16 + (lmat >> 1) * 8);
        mma_bf16_16816(oacc[2 * dp], a, bb[0], bb[1]);
        mma_bf16_16816(oacc[2 * dp + 1], a, bb[2], bb[3]);
      }
    }
  }
  const int r0 = m0 + g, r1 = m0 + g + 8;
#pragma unroll
  for (int dn = 0; dn < kDh / 8; ++dn) {
    const int col = h * kDh + dn * 8 + 2 * t;
    if (r0 < Lq) *reinterpret_cast<uint32_t*>(O + (b * Lq + r0) * ldo + col) = pack2_bf16(oacc[dn][0], oacc[dn][1]);
    if (r1 < Lq) *reinterpret_cast<uint32_t*>(O + (b * Lq + r1) * ldo + col) = pack2_bf16(oacc[dn][2], oacc[dn][3]);
  }
}

template <int kDh>
static int launch_attention_mma(const __nv_bfloat16* Q, int64_t ldq, const __nv_bfloat16* K, int64_t ldk,
                                const __nv_bfloat16* V, int64_t ldv, __nv_bfloat16* O, int64_t ldo, int64_t batch,
                                int heads, int Lq, int Lk, cudaStream_t st) {
  constexpr size_t smem = 3 * kAttnLmax * (kDh + 8) * 2;
  auto kern = attention_mma_kernel<kDh>;
  static std::atomic<bool> configured[64];   // zero-initialised; idempotent per-device attribute set
  int dev = 0;
  ERN_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
    ERN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    if (dev >= 0 && dev < 64) configured[dev].store(true, std::memory_order_release);
  }
  const float scale = 1.0f / sqrtf(static_cast<float>(kDh));
  for (int64_t b0 = 0; b0 < batch; b0 += 65535) {
    const int64_t nb = batch - b0 < 65535 ? batch - b0 : 65535;
    kern<<<dim3(heads, static_cast<unsigned>(nb)), 192, smem, st>>>(Q + b0 * Lq * ldq, ldq, K + b0 * Lk * ldk, ldk,
                                                                   V + b0 * Lk * ldv, ldv, O + b0 * Lq * ldo, ldo, Lq, Lk,
                                                                   scale);
  }
  ERN_CUDA(cudaGetLastError());
  return ERN_OK;
}

// One block per batch element: F.normalize of the patch / token states (models/fusion_model.py:38-41), the first P
// normalised token rows (the only cross-attention queries used, :47) and seq_text_mean (:49).
// 16 warps share the 90 rows of one element (the 8-warp scalar version ran at 1.9 TB/s: 12 dependent row round trips
// per warp); kVec = float4 loads / 8- or 16-byte stores (dim % 128 == 0).  Per-warp partial sums of the text rows are
// combined in a fixed order: deterministic.
constexpr int kPostWarps = 16;
template <typename T> __device__ __forceinline__ void store4(T* dst, float a, float b, float c, float d);
template <> __device__ __forceinline__ void store4<float>(float* dst, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(dst) = make_float4(a, b, c, d);
}
template <> __device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* dst, float a, float b, float c, float d) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
  *reinterpret_cast<uint2*>(dst) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
}

template <typename T, bool kVec>
__global__ void __launch_bounds__(kPostWarps * 32)
post_kernel(const float* __restrict__ X, int P, int Tn, int dim, T* __restrict__ image_norm, T* __restrict__ text_first,
            float* __restrict__ seq_text_mean) {
  extern __shared__ float post_partial[];            // [kPostWarps][dim]
  const int64_t b = blockIdx.x;
  const int L = 1 + P + Tn;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float accum[kMaxPerLane];
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) accum[i] = 0.f;
  auto load_row = [&](int t, float (&v)[kMaxPerLane]) {
    const float* x = X + (b * L + t) * dim;
    if (kVec) {
#pragma unroll
      for (int j = 0; j < kMaxPerLane / 4; ++j) {
        const int d = 4 * lane + 128 * j;
        const float4 q = d < dim ? *reinterpret_cast<const float4*>(x + d) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[4 * j + 0] = q.x;
        v[4 * j + 1] = q.y;
        v[4 * j + 2] = q.z;
        v[4 * j + 3] = q.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < kMaxPerLane; ++i) {
        const int d = lane + 32 * i;
        v[i] = d < dim ? x[d] : 0.f;
      }
    }
  };
  auto emit_row = [&](int t, const float (&v)[kMaxPerLane]) {
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i) ss = fmaf(v[i], v[i], ss);
    const float denom = fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
    const bool is_image = t <= P;
    const bool is_first = !is_image && (t - 1 - P) < P;
    T* dst = is_image ? image_norm + (b * P + (t - 1)) * dim : (is_first ? text_first + (b * P + (t - 1 - P)) * dim : nullptr);
    if (kVec) {
#pragma unroll
      for (int j = 0; j < kMaxPerLane / 4; ++j) {
        const int d = 4 * lane + 128 * j;
        if (d >= dim) continue;
        const float y0 = v[4 * j] / denom, y1 = v[4 * j + 1] / denom, y2 = v[4 * j + 2] / denom, y3 = v[4 * j + 3] / denom;
        if (!is_image) {
          accum[4 * j] += y0;
          accum[4 * j + 1] += y1;
          accum[4 * j + 2] += y2;
          accum[4 * j + 3] += y3;
        }
        if (dst) store4<T>(dst + d, y0, y1, y2, y3);
      }
    } else {
#pragma unroll
      for (int i = 0; i < kMaxPerLane; ++i) {
        const int d = lane + 32 * i;
        if (d >= dim) continue;
        const float y = v[i] / denom;
        if (!is_image) accum[i] += y;
        if (dst) dst[d] = from_f<T>(y);
      }
    }
  };
  // two rows in flight per warp: the loads of row t + 16 are issued before row t is reduced and stored
  float va[kMaxPerLane], vb[kMaxPerLane];
  int t = 1 + warp;
  if (t < L) load_row(t, va);
  while (t < L) {
    const int t1 = t + kPostWarps, t2 = t + 2 * kPostWarps;
    if (t1 < L) load_row(t1, vb);
    emit_row(t, va);
    if (t1 >= L) break;
    if (t2 < L) load_row(t2, va);
    emit_row(t1, vb);
    t = t2;
  }
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int d = ln_index<kVec>(lane, i);
    if (d < dim) post_partial[warp * dim + d] = accum[i];
  }
  __syncthreads();
  for (int d = threadIdx.x; d < dim; d += kPostWarps * 32) {
    float s = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < kPostWarps; ++w8) s += post_partial[w8 * dim + d];   // fixed order: deterministic
    seq_text_mean[b * dim + d] = s / static_cast<float>(Tn);
  }
}

template <typename T>
static int launch_post(const float* X, int64_t batch, int P, int Tn, int dim, T* image_norm, T* text_first,
                       float* seq_text_mean, cudaStream_t st) {
  const size_t smem = static_cast<size_t>(kPostWarps) * dim * sizeof(float);
  ERN_REQUIRE(smem <= 48 * 1024, "feature width %d too large for the DVR post kernel", dim);
  const bool vec = dim % 128 == 0 && aligned16({X, image_norm, text_first});
  if (vec)
    post_kernel<T, true><<<static_cast<unsigned>(batch), kPostWarps * 32, smem, st>>>(X, P, Tn, dim, image_norm, text_first, seq_text_mean);
  else
    post_kernel<T, false><<<static_cast<unsigned>(batch), kPostWarps * 32, smem, st>>>(X, P, Tn, dim, image_norm, text_first, seq_text_mean);
  ERN_CUDA(cudaGetLastError());
  return ERN_OK;
}

// ---- packed bf16 weights: per layer wq wk wv wo (DxD), wi (IxD), wo2 (DxI); then mha_in (3DxD), mha_out (DxD) --------
static size_t al(size_t x) { return (x + 255) & ~size_t(255); }
struct PackedLayout {
  size_t layer_stride, wq, wk, wv, wo, wi, wo2, bqkv, mha_in, mha_out, total;
};
static PackedLayout layout(int dim, int inter, int n_layers) {
  PackedLayout l;
  const size_t d = dim, I = inter;
  l.wq = 0;
  l.wk = al(d * d * 2);
  l.wv = 2 * al(d * d * 2);
  l.wo = 3 * al(d * d * 2);
  l.wi = 4 * al(d * d * 2);
  l.wo2 = l.wi + al(I * d * 2);
  l.bqkv = l.wo2 + al(d * I * 2);              // fp32 [3D]: query | key | value biases, for the fused QKV GEMM
  l.layer_stride = l.bqkv + al(3 * d * 4);
  l.mha_in = l.layer_stride * n_layers;
  l.mha_out = l.mha_in + al(3 * d * d * 2);
  l.total = l.mha_out + al(d * d * 2) + 256;
  return l;
}

size_t packed_bytes(int dim, int inter, int n_layers) { return layout(dim, inter, n_layers).total; }

int pack(const ern_dvr_weights* w, int dim, void* packed, cudaStream_t st) {
  const PackedLayout l = layout(dim, w->intermediate, w->n_layers);
  uint8_t* p = static_cast<uint8_t*>(packed);
  const int64_t dd = static_cast<int64_t>(dim) * dim, di = static_cast<int64_t>(dim) * w->intermediate;
  int rc = 0;
  for (int i = 0; i < w->n_layers && !rc; ++i) {
    const ern_bert_layer_weights& lw = w->layers[i];
    uint8_t* b = p + l.layer_stride * i;
    if (!rc) rc = combiner::launch_cast_bf16(lw.wq, b + l.wq, dd, st);
    if (!rc) rc = combiner::launch_cast_bf16(lw.wk, b + l.wk, dd, st);
    if (!rc) rc = combiner::launch_cast_bf16(lw.wv, b + l.wv, dd, st);
    if (!rc) rc = combiner::launch_cast_bf16(lw.wo, b + l.wo, dd, st);
    if (!rc) rc = combiner::launch_cast_bf16(lw.wi, b + l.wi, di, st);
    if (!rc) rc = combiner::launch_cast_bf16(lw.wo2, b + l.wo2, di, st);
    if (!rc) {
      float* bq = reinterpret_cast<float*>(b + l.bqkv);
      ERN_CUDA(cudaMemcpyAsync(bq, lw.bq, dim * 4, cudaMemcpyDeviceToDevice, st));
      ERN_CUDA(cudaMemcpyAsync(bq + dim, lw.bk, dim * 4, cudaMemcpyDeviceToDevice, st));
      ERN_CUDA(cudaMemcpyAsync(bq + 2 * dim, lw.bv, dim * 4, cudaMemcpyDeviceToDevice, st));
    }
  }
  if (!rc) rc = combiner::launch_cast_bf16(w->mha_in_w, p + l.mha_in, 3 * dd, st);
  if (!rc) rc = combiner::launch_cast_bf16(w->mha_out_w, p + l.mha_out, dd, st);
  return rc;
}

// ---- workspace ------------------------------------------------------------------------------------------------------
struct Work {
  float *X, *Y, *seq_dummy;
  uint8_t *Xb, *QKV, *CTX, *H, *img, *txt, *mq, *mkv, *mctx;
  size_t total;
};
static Work carve(void* base, int64_t batch, int P, int T, int dim, int inter, int mode) {
  const size_t es = mode == ERN_MODE_FP32 ? 4 : 2;
  const size_t M = static_cast<size_t>(batch) * (1 + P + T), MP = static_cast<size_t>(batch) * P, d = dim, I = inter;
  uint8_t* p = static_cast<uint8_t*>(base);
  size_t off = 0;
  Work w;
  auto take = [&](size_t bytes) { uint8_t* r = p + off; off += al(bytes); return r; };
  w.X = reinterpret_cast<float*>(take(M * d * 4));
  w.Y = reinterpret_cast<float*>(take(M * d * 4));
  w.Xb = take(mode == ERN_MODE_FP32 ? 0 : M * d * 2);
  w.QKV = take(M * 3 * d * es);
  w.CTX = take(M * d * es);
  w.H = take(M * I * es);
  w.img = take(MP * d * es);
  w.txt = take(MP * d * es);
  w.mq = take(MP * d * es);
  w.mkv = take(MP * 2 * d * es);
  w.mctx = take(MP * d * es);
  w.total = off + 256;
  return w;
}

size_t workspace_bytes(int64_t batch, int P, int T, int dim, int inter, int mode) {
  return carve(nullptr, batch, P, T, dim, inter, mode).total;
}

template <typename T>
static int run_attention(const T* Q, int64_t ldq, const T* K, int64_t ldk, const T* V, int64_t ldv, T* O, int64_t ldo,
                         int64_t batch, int heads, int Lq, int Lk, int dh, cudaStream_t st) {
  if constexpr (sizeof(T) == 2) {
    const bool vec_ok = (ldq % 8 == 0) && (ldk % 8 == 0) && (ldv % 8 == 0) && (ldo % 2 == 0);
    if (vec_ok && Lq <= kAttnLmax && Lk <= kAttnLmax) {
      if (dh == 80) return launch_attention_mma<80>(Q, ldq, K, ldk, V, ldv, O, ldo, batch, heads, Lq, Lk, st);
      if (dh == 64) return launch_attention_mma<64>(Q, ldq, K, ldk, V, ldv, O, ldo, batch, heads, Lq, Lk, st);
    }
  }
  ERN_REQUIRE(dh % (16 / static_cast<int>(sizeof(T))) == 0 && (ldq * sizeof(T)) % 16 == 0 && (ldk * sizeof(T)) % 16 == 0 &&
                  (ldv * sizeof(T)) % 16 == 0,
              "attention: head size and row strides must allow 16-byte loads (dh = %d)", dh);
  const size_t smem = (static_cast<size_t>(Lq + 2 * Lk) * (dh + 1) + static_cast<size_t>(Lq) * (Lk + 1)) * 4;
  auto kern = attention_kernel<T>;
  static size_t configured[64] = {};
  int dev = 0;
  ERN_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && configured[dev] < smem) {
    ERN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured[dev] = smem;
  }
  for (int64_t b0 = 0; b0 < batch; b0 += 65535) {
    const int64_t nb = batch - b0 < 65535 ? batch - b0 : 65535;
    kern<<<dim3(heads, static_cast<unsigned>(nb)), kAttnThreads, smem, st>>>(Q + b0 * Lq * ldq, ldq, K + b0 * Lk * ldk, ldk,
                                                                   V + b0 * Lk * ldv, ldv, O + b0 * Lq * ldo, ldo, Lq, Lk,
                                                                   dh, 1.0f / sqrtf(static_cast<float>(dh)));
  }
  ERN_CUDA(cudaGetLastError());
  return ERN_OK;
}

// bf16 GEMM helper: out = epilogue(A[M,K] . W[N,K]^T)
template <int kEpi>
static int tc_gemm(const void* A, int64_t M, int K, const void* W, int N, const float* bias, __nv_bfloat16* out_b,
                   int64_t ldo, int col0, float* out_f, const float* residual, int sm_count, cudaStream_t st) {
  CUtensorMap ta, tw;
  int rc;
  if ((rc = simtc::make_tmap_bf16_rows(&ta, A, M, K, K))) return rc;
  if ((rc = simtc::make_tmap_bf16_rows(&tw, W, N, K, K))) return rc;
  gemmtc::Params p{};
  p.m = M;
  p.n = N;
  p.k = K;
  p.bias = bias;
  p.out = out_b;
  p.ldo = ldo;
  p.col0 = col0;
  p.out_f32 = out_f;
  p.residual = residual;
  return gemmtc::launch<256, kEpi, true>(ta, tw, p, sm_count, st);
}

int encode(const ern_dvr_weights* w, int dim, int heads, int P, int T, int mode, const float* patches,
           const float* tokens, int64_t batch, float* out_cross, float* out_seq_mean, void* workspace, int sm_count,
           cudaStream_t st) {
  if (batch <= 0) return ERN_OK;
  const int L = 1 + P + T, I = w->intermediate, dh = dim / heads;
  const int64_t M = batch * L, MP = batch * P;
  const bool f32 = mode == ERN_MODE_FP32;
  Work ws = carve(workspace, batch, P, T, dim, I, mode);
  const int wb_rows = cdiv(M * 32, 256);
  int rc = 0;
  __nv_bfloat16* Xb = f32 ? nullptr : reinterpret_cast<__nv_bfloat16*>(ws.Xb);
  if (dim % 128 == 0 && aligned16({patches, tokens, w->cls_token, w->pos_emb, w->type_emb, w->emb_ln_w, w->emb_ln_b, ws.X, Xb}))
    embed_ln_kernel<true><<<wb_rows, 256, 0, st>>>(patches, tokens, w->cls_token, w->pos_emb, w->type_emb, w->emb_ln_w,
                                                   w->emb_ln_b, batch, P, T, dim, ws.X, Xb);
  else
    embed_ln_kernel<false><<<wb_rows, 256, 0, st>>>(patches, tokens, w->cls_token, w->pos_emb, w->type_emb, w->emb_ln_w,
                                                    w->emb_ln_b, batch, P, T, dim, ws.X, Xb);
  ERN_CUDA(cudaGetLastError());
  const PackedLayout pl = layout(dim, I, w->n_layers);
  const uint8_t* pk = static_cast<const uint8_t*>(w->packed_bf16);

  for (int li = 0; li < w->n_layers; ++li) {
    const ern_bert_layer_weights& lw = w->layers[li];
    if (f32) {
      float* qkv = reinterpret_cast<float*>(ws.QKV);
      float* ctx = reinterpret_cast<float*>(ws.CTX);
      float* hbuf = reinterpret_cast<float*>(ws.H);
      if ((rc = gemmf32::launch<gemmf32::kActNone>(ws.X, dim, M, lw.wq, dim, dim, lw.bq, nullptr, qkv, 3 * dim, st))) return rc;
      if ((rc = gemmf32::launch<gemmf32::kActNone>(ws.X, dim, M, lw.wk, dim, dim, lw.bk, nullptr, qkv + dim, 3 * dim, st))) return rc;
      if ((rc = gemmf32::launch<gemmf32::kActNone>(ws.X, dim, M, lw.wv, dim, dim, lw.bv, nullptr, qkv + 2 * dim, 3 * dim, st))) return rc;
      if ((rc = run_attention<float>(qkv, 3 * dim, qkv + dim, 3 * dim, qkv + 2 * dim, 3 * dim, ctx, dim, batch, heads, L, L, dh, st))) return rc;
      if ((rc = gemmf32::launch<gemmf32::kActNone>(ctx, dim, M, lw.wo, dim, dim, lw.bo, ws.X, ws.Y, dim, st))) return rc;
      launch_layernorm(ws.Y, nullptr, M, dim, lw.ln1_w, lw.ln1_b, ws.X, nullptr, st);
      if ((rc = gemmf32::launch<gemmf32::kActGelu>(ws.X, dim, M, lw.wi, dim, I, lw.bi, nullptr, hbuf, I, st))) return rc;
      if ((rc = gemmf32::launch<gemmf32::kActNone>(hbuf, I, M, lw.wo2, I, dim, lw.bo2, ws.X, ws.Y, dim, st))) return rc;
      launch_layernorm(ws.Y, nullptr, M, dim, lw.ln2_w, lw.ln2_b, ws.X, nullptr, st);
    } else {
      const uint8_t* b = pk + pl.layer_stride * li;
      __nv_bfloat16* qkv = reinterpret_cast<__nv_bfloat16*>(ws.QKV);
      __nv_bfloat16* ctx = reinterpret_cast<__nv_bfloat16*>(ws.CTX);
      __nv_bfloat16* hbuf = reinterpret_cast<__nv_bfloat16*>(ws.H);
      // one GEMM for Q | K | V: the three bf16 weight matrices are adjacent in the packed buffer ([3D, D]) when
      // D*D*2 is a multiple of the 256-byte packing alignment (true for D % 128 == 0), biases are packed likewise
      if (pl.wk == static_cast<size_t>(dim) * dim * 2) {
        if ((rc = tc_gemm<gemmtc::kEpiBiasBf16>(Xb, M, dim, b + pl.wq, 3 * dim, reinterpret_cast<const float*>(b + pl.bqkv),
                                                qkv, 3 * dim, 0, nullptr, nullptr, sm_count, st))) return rc;
      } else {
        if ((rc = tc_gemm<gemmtc::kEpiBiasBf16>(Xb, M, dim, b + pl.wq, dim, lw.bq, qkv, 3 * dim, 0, nullptr, nullptr, sm_count, st))) return rc;
        if ((rc = tc_gemm<gemmtc::kEpiBiasBf16>(Xb, M, dim, b + pl.wk, dim, lw.bk, qkv, 3 * dim, dim, nullptr, nullptr, sm_count, st))) return rc;
        if ((rc = tc_gemm<gemmtc::kEpiBiasBf16>(Xb, M, dim, b + pl.wv, dim, lw.bv, qkv, 3 * dim, 2 * dim, nullptr, nullptr, sm_count, st))) return rc;
      }
      if ((rc = run_attention<__nv_bfloat16>(qkv, 3 * dim, qkv + dim, 3 * dim, qkv + 2 * dim, 3 * dim, ctx, dim, batch, heads, L, L, dh, st))) return rc;
      if ((rc = tc_gemm<gemmtc::kEpiResidF32>(ctx, M, dim, b + pl.wo, dim, lw.bo, nullptr, dim, 0, ws.Y, nullptr, sm_count, st))) return rc;
      launch_layernorm(ws.Y, ws.X, M, dim, lw.ln1_w, lw.ln1_b, ws.X, Xb, st);   // + residual, in place over it
      if ((rc = tc_gemm<gemmtc::kEpiGeluBf16>(Xb, M, dim, b + pl.wi, I, lw.bi, hbuf, I, 0, nullptr, nullptr, sm_count, st))) return rc;
      if ((rc = tc_gemm<gemmtc::kEpiResidF32>(hbuf, M, I, b + pl.wo2, dim, lw.bo2, nullptr, dim, 0, ws.Y, nullptr, sm_count, st))) return rc;
      launch_layernorm(ws.Y, ws.X, M, dim, lw.ln2_w, lw.ln2_b, ws.X, Xb, st);
    }
    ERN_CUDA(cudaGetLastError());
  }

  // ---- normalise, seq_text_mean, cross attention (first P text positions -> P patches) ------------------------------
  const int dd = dim * dim;
  if (f32) {
    float* img = reinterpret_cast<float*>(ws.img);
    float* txt = reinterpret_cast<float*>(ws.txt);
    float* mq = reinterpret_cast<float*>(ws.mq);
    float* mkv = reinterpret_cast<float*>(ws.mkv);
    float* mctx = reinterpret_cast<float*>(ws.mctx);
    if ((rc = launch_post<float>(ws.X, batch, P, T, dim, img, txt, out_seq_mean, st))) return rc;
    if ((rc = gemmf32::launch<gemmf32::kActNone>(txt, dim, MP, w->mha_in_w, dim, dim, w->mha_in_b, nullptr, mq, dim, st))) return rc;
    if ((rc = gemmf32::launch<gemmf32::kActNone>(img, dim, MP, w->mha_in_w + dd, dim, 2 * dim, w->mha_in_b + dim, nullptr, mkv, 2 * dim, st))) return rc;
    if ((rc = run_attention<float>(mq, dim, mkv, 2 * dim, mkv + dim, 2 * dim, mctx, dim, batch, heads, P, P, dh, st))) return rc;
    if ((rc = gemmf32::launch<gemmf32::kActNone>(mctx, dim, MP, w->mha_out_w, dim, dim, w->mha_out_b, nullptr, out_cross, dim, st))) return rc;
  } else {
    __nv_bfloat16* img = reinterpret_cast<__nv_bfloat16*>(ws.img);
    __nv_bfloat16* txt = reinterpret_cast<__nv_bfloat16*>(ws.txt);
    __nv_bfloat16* mq = reinterpret_cast<__nv_bfloat16*>(ws.mq);
    __nv_bfloat16* mkv = reinterpret_cast<__nv_bfloat16*>(ws.mkv);
    __nv_bfloat16* mctx = reinterpret_cast<__nv_bfloat16*>(ws.mctx);
    if ((rc = launch_post<__nv_bfloat16>(ws.X, batch, P, T, dim, img, txt, out_seq_mean, st))) return rc;
    const uint8_t* win = pk + pl.mha_in;
    if ((rc = tc_gemm<gemmtc::kEpiBiasBf16>(txt, MP, dim, win, dim, w->mha_in_b, mq, dim, 0, nullptr, nullptr, sm_count, st))) return rc;
    if ((rc = tc_gemm<gemmtc::kEpiBiasBf16>(img, MP, dim, win + static_cast<size_t>(dd) * 2, 2 * dim, w->mha_in_b + dim, mkv, 2 * dim, 0, nullptr, nullptr, sm_count, st))) return rc;
    if ((rc = run_attention<__nv_bfloat16>(mq, dim, mkv, 2 * dim, mkv + dim, 2 * dim, mctx, dim, batch, heads, P, P, dh, st))) return rc;
    if ((rc = tc_gemm<gemmtc::kEpiResidF32>(mctx, MP, dim, pk + pl.mha_out, dim, w->mha_out_b, nullptr, dim, 0, out_cross, nullptr, sm_count, st))) return rc;
  }
  ERN_CUDA(cudaGetLastError());
  return ERN_OK;
}

}  // namespace dvr
}  // namespace ern
