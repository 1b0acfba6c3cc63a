// Prototypes of the per-file kernel launchers (host functions) used by the C ABI glue.
#pragma once
#include <cuda.h>
#include "ern_common.cuh"

namespace ern {

struct SelectParams;
int launch_select(const SelectParams& p, int64_t nq, cudaStream_t st);
int launch_init_state(int32_t* prev_counts, int32_t* seg_counts, uint32_t* thr_ord, int64_t nq, int n_seg,
                      int32_t* status, cudaStream_t st);
int launch_recall(const int32_t* top_ids, int64_t nq, int k, const int32_t* class_of, int64_t n_gallery,
                  const int32_t* target_class, const int32_t* ks, int nk, int32_t* counts, int32_t* rank_out,
                  cudaStream_t st);
int launch_cirr_subset(const void* queries, int64_t nq, int64_t ldq, const void* gallery, int64_t n_rows,
                       int64_t ldg, int dim, int dtype, const int32_t* members, int m, const int32_t* ref_id,
                       const int32_t* tgt_id, int rank_by, const int32_t* ks, int nk, int32_t* counts,
                       int32_t* rank_out, cudaStream_t st);
int launch_gather_scores(const void* queries, int64_t nq, int64_t ldq, const void* gallery, int64_t n_rows,
                         int64_t ldg, int dim, int dtype, int64_t id_offset, const int32_t* ids, int m, float* out,
                         cudaStream_t st);
int launch_cirr_rank(const float* scores, int64_t nq, const int32_t* members, int m, const int32_t* ref_id,
                     const int32_t* tgt_id, int rank_by, const int32_t* ks, int nk, int32_t* counts, int32_t* rank_out,
                     cudaStream_t st);
int launch_l2norm_rows(const float* x, int64_t rows, int dim, int64_t ldx, int normalize, float* of, int64_t ldf,
                       void* ob, int64_t ldb, cudaStream_t st);

namespace simf32 {
int launch(const float* Q, int64_t ldq, const float* G, int64_t ldg, int dim, const CandidateSink& sink, int rank_by,
           cudaStream_t st);
}
namespace simtc {
int make_tmap_bf16_rows(CUtensorMap* map, const void* base, int64_t rows, int dim, int64_t ld_elems);
int units_for(int64_t nq, int force_single, int sm_count);
int launch(const CUtensorMap& tq, const CUtensorMap& tg, const CandidateSink& sink, int dim, int rank_by,
           int force_single, int sm_count, bool f16_operands, cudaStream_t st);
}
namespace combiner {
// views into the buffer filled by ern_combiner_pack: bf16 K-major weight matrices + fp32 vectors
struct PackedView {
  const __nv_bfloat16 *wt, *wi, *w1;
  const float *bt, *bi, *b1, *w2, *b2;
  unsigned* sync;   // 15 x [4] zeroed counter sets of the small-batch kernel
};
PackedView view_packed(const void* packed, int dim);
int launch_finalize(const float* image, const float* text, int64_t rows, int dim, const float* partial, int n_tiles,
                    const float* b_gate, float* out, void* out_bf16, int64_t ldb, float* gate, cudaStream_t st);
int launch_cast_bf16(const float* src, void* dst, int64_t n, cudaStream_t st);
size_t packed_bytes(int dim);
int pack(const ern_combiner_weights* w, int dim, void* packed, cudaStream_t st);
size_t workspace_bytes_f32(int64_t rows, int dim);
int forward_f32(const ern_combiner_weights* w, int dim, const float* image, const float* text, int64_t rows,
                float* out, void* out_bf16, int64_t ldb, float* gate, void* workspace, cudaStream_t st);
size_t workspace_bytes_bf16(int64_t rows, int dim);
int forward_bf16(const ern_combiner_weights* w, int dim, const float* image, const float* text, int64_t rows,
                 float* out, void* out_bf16, int64_t ldb, float* gate, void* workspace, int sm_count,
                 cudaStream_t st);
namespace small {   // <= 64 rows: weight-streaming kernels (ern_combiner_small.cu)
bool supported(int64_t rows, int dim, int sm_count);
int n_partials(int dim, int sm_count);
int forward(const PackedView& pv, unsigned* sync, int dim, const float* image, const float* text, int64_t rows,
            __nv_bfloat16* raw, float* partial, float* out, void* out_bf16, int64_t ldb, float* gate, int sm_count,
            cudaStream_t st);
}
}

namespace dvr {
size_t packed_bytes(int dim, int inter, int n_layers);
int pack(const ern_dvr_weights* w, int dim, void* packed, cudaStream_t st);
size_t workspace_bytes(int64_t batch, int P, int T, int dim, int inter, int mode);
int encode(const ern_dvr_weights* w, int dim, int heads, int P, int T, int mode, const float* patches,
           const float* tokens, int64_t batch, float* out_cross, float* out_seq_mean, void* workspace, int sm_count,
           cudaStream_t st);
}
namespace bbcloss {
size_t workspace_bytes(int64_t b, int dim, int mode);
int forward(const float* pred, int64_t ldp, const float* tar, int64_t ldt, int64_t b, int dim, float scale, int mode,
            float* loss, float* lse, void* workspace, int sm_count, cudaStream_t st);
int backward(const float* pred, int64_t ldp, const float* tar, int64_t ldt, int64_t b, int dim, float scale, int mode,
             const float* lse, const float* grad_out, float* dpred, int64_t lddp, float* dtar, int64_t lddt,
             void* workspace, int sm_count, cudaStream_t st);
}
namespace visualsr {
size_t packed_bytes(int dim);
int pack(const ern_visualsr_weights* w, int dim, void* packed, cudaStream_t st);
size_t workspace_bytes(int64_t rows, int patches, int dim, int mode);
int forward(const ern_visualsr_weights* w, int dim, int patches, int mode, const float* local, int64_t rows,
            float* out, void* workspace, int sm_count, cudaStream_t st);
}

}  // namespace ern
