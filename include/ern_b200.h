/*
 * ern_b200.h -- C ABI of libern_b200.so: the B200-native (sm_100a) composed-retrieval scoring path.
 *
 * The reference (ChenAnno/FashionERN_AAAI2024) is pure Python/PyTorch and has no FFI layer; its
 * "operator API" for this path is (i) the module CombinerSimple (models/fusion_model.py:58-94) and
 * (ii) the tails of compute_{fiq,shoes,200k,cirr}_val_metrics (run/test/test_fiq.py:44-64,
 * test_shoes.py:44-61, test_200k.py:46-61, test_cirr.py:46-80, test_val.py:45-67 and the
 * run/valid/validate_*.py twins).  Each entry point below names the reference lines it replaces.
 * The Python mirror of those two surfaces lives in fashionern_aaai2024_b200/ and binds this
 * header with ctypes.
 *
 * Conventions
 *   - plain C types only; every pointer named *_dev / documented "device" is a CUDA device pointer
 *     owned by the caller (e.g. torch.Tensor.data_ptr()); the library allocates nothing persistent.
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered and asynchronous.
 *   - scratch is caller-provided: query the size with the matching *_workspace_bytes().
 *   - every call returns 0 on success, <0 on error; ern_last_error() returns a thread-local message.
 *   - there is no CPU fallback: on a device that is not compute capability 10.x calls fail with
 *     ERN_ERR_DEVICE.
 */
#ifndef ERN_B200_H_
#define ERN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define ERN_API __attribute__((visibility("default")))
#else
#define ERN_API
#endif

#define ERN_OK 0
#define ERN_ERR_ARG (-1)
#define ERN_ERR_CUDA (-2)
#define ERN_ERR_DEVICE (-3)
#define ERN_ERR_WORKSPACE (-4)
#define ERN_ERR_UNSUPPORTED (-5)

/* arithmetic mode */
#define ERN_MODE_BF16 0 /* bf16 operands, fp32 accumulate, tcgen05 tensor cores (the product path)   */
#define ERN_MODE_FP32 1 /* fp32 FFMA "validation mode": 1e-5 parity with the reference's fp32 torch  */

/* element type of a feature matrix handed to the library */
#define ERN_DTYPE_F32 0
#define ERN_DTYPE_BF16 1
#define ERN_DTYPE_F16 2  /* fp16 features on the tensor-core path (MODE_BF16): 11 mantissa bits instead of 8 at the same
                            tcgen05 rate; for unit-norm features the range is no concern (north_star: "bf16/fp16") */
/* ern_l2norm_rows: OR into `normalize` to make the 16-bit output fp16 instead of bf16 */
#define ERN_NORM_OUT_F16 2

/* what a candidate is ranked by */
#define ERN_RANK_SIMILARITY 0 /* s = <q, g>                                                       */
#define ERN_RANK_REFERENCE 1  /* -(1 - s) rounded in fp32: the reference's `1 - pred @ index.T`    */
                              /* (run/test/test_fiq.py:49) with its rounding-induced ties          */

/* max k of the streaming top-k; slots of one candidate segment (one per query and persistent scoring unit: room for
 * the survivors of a pruning pass plus every score of one 256-row gallery tile); max candidates the selection/merge kernel stages in shared memory (more are handled from L2); queries that
 * go through the launch schedule together (larger batches are processed ERN_QUERY_BATCH at a time) */
#define ERN_MAX_K 128
#define ERN_SEG_CAP 512
#define ERN_SORT_CAP 2048
#define ERN_QUERY_BATCH 4096
/* gallery rows scored densely (every score kept) before thresholds exist; also the slots of a query's prefix list */
#define ERN_DENSE_ROWS 256

ERN_API int ern_version(void);
ERN_API const char* ern_last_error(void);
/* 0 iff `device` exists and is sm_100-class; the product refuses to run elsewhere. */
ERN_API int ern_device_check(int device);

/* ---------------------------------------------------------------------------------------------
 * Row L2-normalise (+ optional 16-bit cast: bf16, or fp16 with normalize | ERN_NORM_OUT_F16).
 * Replaces F.normalize(index_features, dim=-1).float() (run/test/test_fiq.py:45 and twins),
 * eps = 1e-12 as torch's default.  Either output may be NULL.  ld* are row strides in elements.
 * ------------------------------------------------------------------------------------------- */
ERN_API int ern_l2norm_rows(const float* x_dev, int64_t rows, int dim, int64_t ldx, int normalize,
                    float* out_f32_dev, int64_t ld_f32, void* out_bf16_dev, int64_t ld_bf16,
                    void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fusion head: CombinerSimple.forward(image_features, text_features) in eval mode
 * (models/fusion_model.py:86-94; parameters :73-84).  Weights are torch's fp32 row-major
 * [out,in] Linear matrices; `packed_bf16` is the buffer filled by ern_combiner_pack (needed for
 * ERN_MODE_BF16 only).
 * ------------------------------------------------------------------------------------------- */
typedef struct ern_combiner_weights {
  const float* w_text;  /* text_projection_layer.0.weight  [4D, D]  */
  const float* b_text;  /* text_projection_layer.0.bias    [4D]     */
  const float* w_image; /* image_projection_layer.0.weight [4D, D]  */
  const float* b_image; /* image_projection_layer.0.bias   [4D]     */
  const float* w_hid;   /* dynamic_scalar.0.weight         [8D, 8D] */
  const float* b_hid;   /* dynamic_scalar.0.bias           [8D]     */
  const float* w_gate;  /* dynamic_scalar.3.weight         [1, 8D]  */
  const float* b_gate;  /* dynamic_scalar.3.bias           [1]      */
  const void* packed_bf16;
} ern_combiner_weights;

ERN_API size_t ern_combiner_packed_bytes(int dim);
ERN_API int ern_combiner_pack(const ern_combiner_weights* w, int dim, void* packed_dev, void* stream);
ERN_API size_t ern_combiner_workspace_bytes(int64_t rows, int dim, int mode);
/* out_f32 [rows, D] unit-norm fused features; out_bf16 (nullable) the same rounded to bf16 with row
 * stride ld_bf16 (ready to be a query/gallery operand of ern_sim_topk); gate (nullable) [rows] the
 * dynamic scalar s. */
ERN_API int ern_combiner_forward(const ern_combiner_weights* w, int dim, int mode, const float* image_dev,
                         const float* text_dev, int64_t rows, float* out_f32_dev,
                         void* out_bf16_dev, int64_t ld_bf16, float* gate_dev, void* workspace_dev,
                         size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * VisualSR.forward in eval mode (models/fusion_model.py:141-154): attention pooling of
 * the P (= 13) patch embeddings [rows, P, D] -> unit-norm [rows, D] (own l2norm with +1e-8, :136-139).
 * The eval-mode BatchNorm1d layers arrive folded: scale = gamma / sqrt(running_var + eps),
 * shift = beta - running_mean * scale (embedding_local.1 has P channels, embedding_global.1 has D).
 * ERN_MODE_BF16 needs dim % 128 == 0 and packed weights (ern_visualsr_pack).
 * ------------------------------------------------------------------------------------------- */
typedef struct ern_visualsr_weights {
  const float* w_local;         /* embedding_local.0.weight  [D, D] */
  const float* b_local;         /* embedding_local.0.bias    [D]    */
  const float* bn_local_scale;  /* folded embedding_local.1  [P]    */
  const float* bn_local_shift;  /*                           [P]    */
  const float* w_global;        /* embedding_global.0.weight [D, D] */
  const float* b_global;        /* embedding_global.0.bias   [D]    */
  const float* bn_global_scale; /* folded embedding_global.1 [D]    */
  const float* bn_global_shift; /*                           [D]    */
  const float* w_common;        /* embedding_common.weight   [1, D] */
  const float* b_common;        /* embedding_common.bias     [1]    */
  const void* packed_bf16;
} ern_visualsr_weights;

ERN_API size_t ern_visualsr_packed_bytes(int dim);
ERN_API int ern_visualsr_pack(const ern_visualsr_weights* w, int dim, void* packed_dev, void* stream);
ERN_API size_t ern_visualsr_workspace_bytes(int64_t rows, int patches, int dim, int mode);
ERN_API int ern_visualsr_forward(const ern_visualsr_weights* w, int dim, int patches, int mode,
                         const float* local_dev, int64_t rows, float* out_f32_dev, void* workspace_dev,
                         size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Query-side encoder of DVR_module in eval mode (models/fusion_model.py:26-49):
 * PlusModel = HF BERT encoder over [CLS] + P patch + T token embeddings (:187-216), F.normalize of the
 * patch/token states (:38-41), nn.MultiheadAttention cross attention of which only the first P query
 * positions are used (:44-47), mean of the normalised token states (:49).
 *   patches [B,P,D], tokens [B,T,D] fp32  ->  out_cross [B,P,D] (input of SR_module),
 *                                            out_seq_mean [B,D] (text input of combiner_local)
 * Linear weights are torch's fp32 [out,in]; ERN_MODE_BF16 needs dim % 128 == 0, intermediate % 128 == 0
 * and packed weights (ern_dvr_pack).
 * ------------------------------------------------------------------------------------------- */
typedef struct ern_bert_layer_weights {
  const float *wq, *bq, *wk, *bk, *wv, *bv; /* attention.self.{query,key,value}          */
  const float *wo, *bo, *ln1_w, *ln1_b;     /* attention.output.{dense,LayerNorm}         */
  const float *wi, *bi;                     /* intermediate.dense [I, D]                  */
  const float *wo2, *bo2, *ln2_w, *ln2_b;   /* output.{dense [D, I],LayerNorm}            */
} ern_bert_layer_weights;

#define ERN_MAX_BERT_LAYERS 4
typedef struct ern_dvr_weights {
  const float* cls_token; /* transformer_layer.cls_token [D]                                    */
  const float* pos_emb;   /* embeddings.position_embeddings.weight [>= 1+P+T, D]                */
  const float* type_emb;  /* embeddings.token_type_embeddings.weight [2, D]                     */
  const float* emb_ln_w;  /* embeddings.LayerNorm                                               */
  const float* emb_ln_b;
  int n_layers;
  int intermediate;
  ern_bert_layer_weights layers[ERN_MAX_BERT_LAYERS];
  const float* mha_in_w;  /* MR_component.in_proj_weight [3D, D]                                */
  const float* mha_in_b;  /* MR_component.in_proj_bias   [3D]                                   */
  const float* mha_out_w; /* MR_component.out_proj.weight [D, D]                                */
  const float* mha_out_b;
  const void* packed_bf16;
} ern_dvr_weights;

ERN_API size_t ern_dvr_packed_bytes(int dim, int intermediate, int n_layers);
ERN_API int ern_dvr_pack(const ern_dvr_weights* w, int dim, void* packed_dev, void* stream);
ERN_API size_t ern_dvr_workspace_bytes(int64_t batch, int patches, int tokens, int dim, int intermediate, int mode);
ERN_API int ern_dvr_encode(const ern_dvr_weights* w, int dim, int heads, int patches, int tokens, int mode,
                   const float* patches_dev, const float* tokens_dev, int64_t batch,
                   float* out_cross_dev, float* out_seq_mean_dev, void* workspace_dev,
                   size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Similarity + streaming per-query top-k over one gallery shard.
 * Replaces `distances = 1 - predicted_features @ index_features.T;
 *           sorted_indices = torch.argsort(distances, dim=-1)` (run/test/test_fiq.py:49-50 and twins)
 * for the first k columns; the [Q,N] matrix is never materialised.
 *
 *   queries   [nq, dim]   row stride ldq,  dtype (ERN_DTYPE_F32 for MODE_FP32; ERN_DTYPE_BF16 or ERN_DTYPE_F16 for
 *             MODE_BF16, the tensor-core mode: same kernels and speed, only the operand format differs)
 *   gallery   [n_rows, dim] row stride ldg, same dtype; n_rows rows of this shard
 *   id_offset global id of gallery row 0 (shard offset); ids returned are global
 *   exclude_id_dev (nullable) [nq] global id removed from query q's ranking (CIRR reference image,
 *             run/test/test_cirr.py:55-58); -1 = none
 *   out_scores [nq,k] fp32 ranking value (similarity, or -(1-s) for ERN_RANK_REFERENCE), best first;
 *   out_ids    [nq,k] int32 global ids, ties -> lower id first; missing entries: score -inf, id -1
 *   out_keys   (nullable) [nq,k] uint64 sortable (value,id) keys, the wire format of the multi-GPU
 *             candidate exchange (ern_topk_merge)
 *   growth    gallery-range growth factor of the launch schedule (>=2; 8 is the default): launch i covers rows
 *             [b, growth*b) -- tensor-core launches at most 2M rows (environment ERN_LAUNCH_MAX_ROWS) -- starting
 *             from the exact k-th best of rows [0,b) as each query's threshold.  It is a cost knob only: candidate
 *             segments prune themselves inside the kernel, so the result is exact for ANY gallery order and any growth
 *             (growth == 1: fixed (2048 - k)-row steps, a test hook).
 *   status_dev int32[4]: [0] != 0 => internal error (candidate storage inconsistent; never expected);
 *             [1..3] mbarrier watchdog diagnostics of a trapped launch.
 * MODE_BF16 requires dim % 64 == 0, dim <= 768 and 16-byte aligned rows.
 * ------------------------------------------------------------------------------------------- */
ERN_API size_t ern_sim_topk_workspace_bytes(int64_t nq, int dim, int mode);
ERN_API int ern_sim_topk(const void* queries_dev, int64_t nq, int64_t ldq, const void* gallery_dev,
                 int64_t n_rows, int64_t ldg, int dim, int dtype, int64_t id_offset,
                 const int32_t* exclude_id_dev, int k, int mode, int rank_by, int growth,
                 float* out_scores_dev, int32_t* out_ids_dev, uint64_t* out_keys_dev,
                 int32_t* status_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * ern_sim_topk fused with the multi-GPU candidate exchange: the last selection launch stores this
 * rank's [nq,k] keys straight into slot `rank` of EVERY rank's gathered buffer [world, nq, k] through
 * peer pointers (NVLink P2P stores into symmetric memory) instead of writing them locally for a
 * separate all-gather.  peer_keys_dev: device array of `world` pointers (e.g. torch symmetric
 * memory's buffer_ptrs_dev).  The caller inserts a cross-rank barrier before ern_topk_merge.
 * ------------------------------------------------------------------------------------------- */
ERN_API int ern_sim_topk_exchange(const void* queries_dev, int64_t nq, int64_t ldq, const void* gallery_dev,
                          int64_t n_rows, int64_t ldg, int dim, int dtype, int64_t id_offset,
                          const int32_t* exclude_id_dev, int k, int mode, int rank_by, int growth,
                          uint64_t* const* peer_keys_dev, int world, int rank, int32_t* status_dev,
                          void* workspace_dev, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * k-way merge of per-shard / per-rank candidate lists (after the NCCL all-gather of keys):
 * list l of query q starts at keys_dev + l*list_stride + q*query_stride and holds k_in keys.
 * n_lists * k_in <= ERN_SORT_CAP.
 * ------------------------------------------------------------------------------------------- */
ERN_API int ern_topk_merge(const uint64_t* keys_dev, int64_t nq, int n_lists, int k_in, int64_t list_stride,
                   int64_t query_stride, int k_out, float* out_scores_dev, int32_t* out_ids_dev,
                   uint64_t* out_keys_dev, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Recall@K from id membership on device.
 * Replaces the numpy string gather/compare + slice + sum of run/test/test_fiq.py:51-60 (unique
 * names) and run/test/test_200k.py:52-60 (any-hit on non-unique caption names):
 *   rank[q] = first j < k with class_of[top_ids[q,j]] == target_class[q]  (k if none)
 *   counts[i] = #{q : rank[q] < ks[i]}
 * class_of_dev [n_gallery] int32 maps a global gallery id to its name class; ks is a HOST array.
 * rank_dev (nullable) receives the per-query ranks.
 * ------------------------------------------------------------------------------------------- */
ERN_API int ern_recall_at_k(const int32_t* top_ids_dev, int64_t nq, int k, const int32_t* class_of_dev,
                    int64_t n_gallery, const int32_t* target_class_dev, const int32_t* ks, int nk,
                    int32_t* counts_dev, int32_t* rank_dev, void* stream);

/* ---------------------------------------------------------------------------------------------
 * CIRR subset recall (run/test/test_cirr.py:64-66,76-78): rank of the target among the group
 * members that are not the reference image, ordered by the same ranking value (ties -> lower id).
 *   members_dev [nq, m] int32 gallery row ids (-1 = absent); rank_dev[q] = -1 if the target is not
 *   among the surviving members (the reference asserts on that, test_cirr.py:69).
 *   counts[i] = #{q : 0 <= rank[q] < ks[i]}
 * ------------------------------------------------------------------------------------------- */
ERN_API int ern_cirr_subset_recall(const void* queries_dev, int64_t nq, int64_t ldq, const void* gallery_dev,
                           int64_t n_rows, int64_t ldg, int dim, int dtype,
                           const int32_t* members_dev, int m, const int32_t* reference_id_dev,
                           const int32_t* target_id_dev, int rank_by, const int32_t* ks, int nk,
                           int32_t* counts_dev, int32_t* rank_dev, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Split form of the CIRR subset recall for row-sharded galleries: every rank scores the
 * group members whose rows it owns -- out[q,j] = <query q, gallery row ids[q,j] - id_offset> or 0 when the
 * row is not in [id_offset, id_offset + n_rows) -- the caller sums the [nq,m] matrices over the ranks
 * (one owner per member, so the sum is exact) and ranks with ern_cirr_subset_from_scores.
 * ------------------------------------------------------------------------------------------- */
ERN_API int ern_gather_scores(const void* queries_dev, int64_t nq, int64_t ldq, const void* gallery_dev,
                      int64_t n_rows, int64_t ldg, int dim, int dtype, int64_t id_offset,
                      const int32_t* ids_dev, int m, float* out_scores_dev, void* stream);
ERN_API int ern_cirr_subset_from_scores(const float* scores_dev, int64_t nq, const int32_t* members_dev, int m,
                                const int32_t* reference_id_dev, const int32_t* target_id_dev,
                                int rank_by, const int32_t* ks, int nk, int32_t* counts_dev,
                                int32_t* rank_dev, void* stream);

/* ---------------------------------------------------------------------------------------------
 * In-batch classification loss.  Replaces BatchBasedClassificationLoss.forward
 * (losses/loss.py:10-14): logits = scale * pred . tar^T (scale = 100, :11), labels = arange(batch) (:12),
 * loss = mean cross entropy (:14) -- and its gradient (what autograd derives for run/train/train_fiq.py:134-137).
 *   pred, tar   [batch, dim] fp32, row strides ldp / ldt (elements); dim % 64 == 0, batch >= 1
 *   mode        ERN_MODE_BF16: tcgen05 GEMMs with a logsumexp / softmax-gradient epilogue, operands rounded to
 *               bf16, logits never stored;  ERN_MODE_FP32: validation path, fp32 FFMA
 *   loss_dev    [1] fp32;  lse_dev [batch] fp32 row logsumexp (nullable in forward; required by backward)
 *   backward:   grad_out_dev [1] fp32 device scalar = d(objective)/d(loss) (nullable = 1; a GradScaler factor
 *               arrives here), dpred / dtar [batch, dim] fp32 with row strides lddp / lddt (multiples of 4)
 * Workspace: ern_bbc_loss_workspace_bytes(batch, dim, mode) bytes, contents need not survive between the calls.
 * ------------------------------------------------------------------------------------------- */
ERN_API size_t ern_bbc_loss_workspace_bytes(int64_t batch, int dim, int mode);
ERN_API int ern_bbc_loss_forward(const float* pred_dev, int64_t ldp, const float* tar_dev, int64_t ldt,
                                 int64_t batch, int dim, float scale, int mode, float* loss_dev, float* lse_dev,
                                 void* workspace_dev, size_t workspace_bytes, void* stream);
ERN_API int ern_bbc_loss_backward(const float* pred_dev, int64_t ldp, const float* tar_dev, int64_t ldt,
                                  int64_t batch, int dim, float scale, int mode, const float* lse_dev,
                                  const float* grad_out_dev, float* dpred_dev, int64_t lddp, float* dtar_dev,
                                  int64_t lddt, void* workspace_dev, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ERN_B200_H_ */
