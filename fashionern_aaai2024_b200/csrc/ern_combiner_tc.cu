// Tensor-core path of the fusion head (CombinerSimple.forward, models/fusion_model.py:86-94).
//
// Uses the persistent, warp-specialised tcgen05 GEMM of ern_gemm_tc.cuh, C = relu(A . W^T + bias), with two epilogues:
//   kStore : C is rounded to bf16 and stored            -> the two projections write the halves of raw[B,8D]
//                                                          (this replaces torch.cat, :90)
//   kGate  : C is multiplied by w2 and row-reduced       -> per 256-column tile partial of the gate logit;
//                                                          h[B,8D] (:74-75) never reaches HBM
// A [M,K] and W [N,K] are bf16, K-major, fed by TMA into a 6-stage ring of 128-row x 64 swizzled tiles per CTA;
// a CTA pair issues 256 x 256 x 16 cta_group::2 MMAs, the fp32 accumulator (256 TMEM columns per CTA) is double
// buffered (512 columns) so the epilogue of tile i overlaps the MMAs of tile i+1.  The finaliser (gate sigmoid,
// blend from the fp32 inputs, L2-normalise) is shared with the fp32 path.
#include "ern_gemm_tc.cuh"

namespace ern {
namespace combiner {



constexpr int kGemmBlockN = 256;
static size_t al(size_t x) { return (x + 255) & ~size_t(255); }

size_t workspace_bytes_bf16(int64_t rows, int dim) {
  const size_t r = rows, d = dim;
  // (the small-batch path keeps one gate partial per CTA, up to 8D / 32 of them, instead of one per 256-column tile)
  const size_t n_part = rows <= 64 ? 8 * d / 32 : 8 * d / kGemmBlockN;
  return al(r * d * 2) * 2 + al(r * 8 * d * 2) + al(r * n_part * 4) + 512;
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
  const int n_tiles = hid / kGemmBlockN;
  PackedView pv = view_packed(w->packed_bf16, dim);
  if (small::supported(rows, dim, sm_count)) {
    // weight-bandwidth-bound regime (the reference's 32-row query batches): one cooperative weight-streaming launch.
    // Its grid-barrier counters are one of 15 zeroed sets behind the packed weights, taken round-robin per call and
    // left at zero by the kernel's last CTA: up to 15 forwards of one module may be in flight on different streams
    // (a memset per call instead costs 2.7 us of the 29).
    static std::atomic<unsigned> next_set{0};
    unsigned* sync = pv.sync + 4 * (next_set.fetch_add(1, std::memory_order_relaxed) % 15u);
    return small::forward(pv, sync, dim, image, text, rows, raw, partial, out, out_bf16, ldb, gate, sm_count, st);
  }

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

  gemmtc::Params p{};
  p.m = rows;
  p.n = proj;
  p.k = dim;
  p.out = raw;
  p.ldo = hid;
  // text projection -> raw[:, 0:4D], image projection -> raw[:, 4D:8D]  (text half first, fusion_model.py:90)
  p.bias = pv.bt;
  p.col0 = 0;
  if ((rc = gemmtc::launch<kGemmBlockN, gemmtc::kEpiStoreRelu, true>(t_txt, t_wt, p, sm_count, st))) return rc;
  p.bias = pv.bi;
  p.col0 = proj;
  if ((rc = gemmtc::launch<kGemmBlockN, gemmtc::kEpiStoreRelu, true>(t_img, t_wi, p, sm_count, st))) return rc;
  // hidden layer + gate dot product
  gemmtc::Params g{};
  g.m = rows;
  g.n = hid;
  g.k = hid;
  g.bias = pv.b1;
  g.wg = pv.w2;
  g.partial = partial;
  if ((rc = gemmtc::launch<kGemmBlockN, gemmtc::kEpiGate, true>(t_raw, t_w1, g, sm_count, st))) return rc;
  return launch_finalize(image, text, rows, dim, partial, n_tiles, pv.b2, out, out_bf16, ldb, gate, st);
}

}  // namespace combiner
}  // namespace ern
