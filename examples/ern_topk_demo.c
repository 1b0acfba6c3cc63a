/* Plain-C caller of libern_b200.so: the drop-in boundary does not need Python or C++.
 *
 *   gcc -std=c99 -O2 -I include -I /usr/local/cuda/include examples/ern_topk_demo.c -o ern_topk_demo \
 *       -L fashionern_aaai2024_b200 -l:libern_b200.so -L /usr/local/cuda/lib64 -lcudart -lm \
 *       -Wl,-rpath,$PWD/fashionern_aaai2024_b200
 *
 * Builds a small unit-norm bf16 gallery on the host, plants every query as an exact copy of one gallery row,
 * runs ern_sim_topk (tcgen05 path) and checks that the planted row comes back first with similarity ~1.
 */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ern_b200.h"

static unsigned short f32_to_bf16(float f) { /* round to nearest even */
  unsigned int u;
  memcpy(&u, &f, 4);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (unsigned short)(u >> 16);
}

#define CHECK_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)
#define CHECK_ERN(x) do { int rc_ = (x); if (rc_ != 0) { \
  fprintf(stderr, "libern_b200 error %d: %s\n", rc_, ern_last_error()); return 3; } } while (0)

int main(void) {
  const int nq = 300, dim = 640, k = 10;
  const long n = 50000;
  unsigned short* g = (unsigned short*)malloc((size_t)n * dim * 2);
  unsigned short* q = (unsigned short*)malloc((size_t)nq * dim * 2);
  float* row = (float*)malloc(dim * sizeof(float));
  unsigned int seed = 12345u;
  long i;
  int d, j;
  if (ern_device_check(0) != 0) { fprintf(stderr, "%s\n", ern_last_error()); return 1; }
  for (i = 0; i < n; ++i) {
    double ss = 0.0;
    for (d = 0; d < dim; ++d) {
      seed = seed * 1664525u + 1013904223u;
      row[d] = (float)((seed >> 8) & 0xFFFF) / 65536.0f - 0.5f;
      ss += (double)row[d] * row[d];
    }
    for (d = 0; d < dim; ++d) g[i * dim + d] = f32_to_bf16(row[d] / (float)sqrt(ss));
  }
  for (j = 0; j < nq; ++j) memcpy(q + (size_t)j * dim, g + (size_t)((long)j * 163 % n) * dim, dim * 2);

  void *g_dev, *q_dev, *ws;
  float* scores_dev;
  int *ids_dev, *status_dev;
  size_t ws_bytes = ern_sim_topk_workspace_bytes(nq, dim, ERN_MODE_BF16);
  CHECK_CUDA(cudaMalloc(&g_dev, (size_t)n * dim * 2));
  CHECK_CUDA(cudaMalloc(&q_dev, (size_t)nq * dim * 2));
  CHECK_CUDA(cudaMalloc(&ws, ws_bytes));
  CHECK_CUDA(cudaMalloc((void**)&scores_dev, (size_t)nq * k * 4));
  CHECK_CUDA(cudaMalloc((void**)&ids_dev, (size_t)nq * k * 4));
  CHECK_CUDA(cudaMalloc((void**)&status_dev, 16));
  CHECK_CUDA(cudaMemcpy(g_dev, g, (size_t)n * dim * 2, cudaMemcpyHostToDevice));
  CHECK_CUDA(cudaMemcpy(q_dev, q, (size_t)nq * dim * 2, cudaMemcpyHostToDevice));

  CHECK_ERN(ern_sim_topk(q_dev, nq, dim, g_dev, n, dim, dim, ERN_DTYPE_BF16, 0, NULL, k, ERN_MODE_BF16,
                         ERN_RANK_SIMILARITY, 8, scores_dev, ids_dev, NULL, status_dev, ws, ws_bytes, NULL));
  CHECK_CUDA(cudaDeviceSynchronize());

  float* scores = (float*)malloc((size_t)nq * k * 4);
  int* ids = (int*)malloc((size_t)nq * k * 4);
  int status[4];
  CHECK_CUDA(cudaMemcpy(scores, scores_dev, (size_t)nq * k * 4, cudaMemcpyDeviceToHost));
  CHECK_CUDA(cudaMemcpy(ids, ids_dev, (size_t)nq * k * 4, cudaMemcpyDeviceToHost));
  CHECK_CUDA(cudaMemcpy(status, status_dev, 16, cudaMemcpyDeviceToHost));
  int bad = status[0] != 0;
  for (j = 0; j < nq; ++j) {
    if (ids[j * k] != (int)((long)j * 163 % n) || fabsf(scores[j * k] - 1.0f) > 2e-2f) ++bad;
    for (d = 1; d < k; ++d)
      if (scores[j * k + d] > scores[j * k + d - 1]) ++bad;
  }
  printf("%s: %d queries x %ld rows, top-1 of query 0 = row %d (similarity %.4f)\n", bad ? "FAILED" : "OK", nq, n,
         ids[0], scores[0]);
  return bad ? 4 : 0;
}
