// fp32 "validation mode" of the similarity + top-k candidate filter: plain FFMA tiles, no tensor cores.
// Exists so that parity with the reference's fp32 torch path (`1 - pred @ index.T`, run/test/test_fiq.py:49)
// can be shown at 1e-5 without bf16 operand rounding in the way; it feeds the same candidate sink and
// compaction pipeline as the tensor-core kernel.  Not a performance path.
#include "ern_common.cuh"

namespace ern {
namespace simf32 {

constexpr int kTile = 64;   // 64 queries x 64 gallery rows per block
constexpr int kKc = 16;     // k-chunk
constexpr int kThreads = 256;

template <int kRankBy>
__global__ void __launch_bounds__(kThreads)
sim_f32_kernel(const float* __restrict__ Q, int64_t ldq, const float* __restrict__ G, int64_t ldg, int dim,
               const CandidateSink sink) {
  __shared__ float qs[kKc][kTile + 1];
  __shared__ float gs[kKc][kTile + 1];
  const int tx = threadIdx.x & 15;  // gallery direction
  const int ty = threadIdx.x >> 4;  // query direction
  const int64_t q0 = static_cast<int64_t>(blockIdx.y) * kTile;
  const int64_t g0 = sink.row_begin + static_cast<int64_t>(blockIdx.x) * kTile;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < dim; k0 += kKc) {
    // 64 rows x 16 k per operand = 1024 elements, 4 per thread; consecutive threads read consecutive k
    for (int e = threadIdx.x; e < kTile * kKc; e += kThreads) {
      const int r = e / kKc, kk = e % kKc;
      const int64_t qr = q0 + r, gr = g0 + r;
      const bool kin = (k0 + kk) < dim;
      qs[kk][r] = (kin && qr < sink.nq) ? Q[qr * ldq + k0 + kk] : 0.f;
      gs[kk][r] = (kin && gr < sink.row_end) ? G[gr * ldg + k0 + kk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kKc; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = qs[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = gs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t q = q0 + ty * 4 + i;
    if (q >= sink.nq) continue;
    const float thr = sink.dense ? -INFINITY : ordered_to_f32(sink.thr_ord[q]);
    const int32_t excl = sink.exclude ? sink.exclude[q] : -1;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t row = g0 + tx * 4 + j;
      if (row >= sink.row_end) continue;
      const float r = rank_value<kRankBy>(acc[i][j]);
      if (sink.dense) sink_put_dense(sink, q, row, r, excl);
      else if (r >= thr) sink_put_atomic(sink, q, row, r, excl);
    }
  }
}

int launch(const float* Q, int64_t ldq, const float* G, int64_t ldg, int dim, const CandidateSink& sink, int rank_by,
           cudaStream_t st) {
  const int64_t rows = sink.row_end - sink.row_begin;
  if (rows <= 0 || sink.nq <= 0) return ERN_OK;
  const int qblocks = cdiv(sink.nq, kTile);
  ERN_REQUIRE(qblocks <= 65535, "too many queries for the fp32 validation kernel (%lld)", (long long)sink.nq);
  dim3 grid(cdiv(rows, kTile), qblocks);
  if (rank_by == ERN_RANK_REFERENCE)
    sim_f32_kernel<ERN_RANK_REFERENCE><<<grid, kThreads, 0, st>>>(Q, ldq, G, ldg, dim, sink);
  else
    sim_f32_kernel<ERN_RANK_SIMILARITY><<<grid, kThreads, 0, st>>>(Q, ldq, G, ldg, dim, sink);
  ERN_CUDA(cudaGetLastError());
  return ERN_OK;
}

}  // namespace simf32
}  // namespace ern
