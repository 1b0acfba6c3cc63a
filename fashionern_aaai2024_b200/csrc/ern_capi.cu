// C ABI of libern_b200.so (declared in include/ern_b200.h): argument checking, launch plans, error strings.
#include <stdarg.h>
#include <string.h>

#include "ern_internal.cuh"
#include "ern_select.cuh"

namespace ern {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", static_cast<int>(e), cudaGetErrorString(e), what);
  return ERN_ERR_CUDA;
}

struct DeviceInfo {
  int device = -1;
  int major = 0, minor = 0;
  int sm_count = 0;
};
static int current_device(DeviceInfo* out) {
  static thread_local DeviceInfo cache;
  int dev = -1;
  ERN_CUDA(cudaGetDevice(&dev));
  if (cache.device != dev) {
    DeviceInfo d;
    d.device = dev;
    ERN_CUDA(cudaDeviceGetAttribute(&d.major, cudaDevAttrComputeCapabilityMajor, dev));
    ERN_CUDA(cudaDeviceGetAttribute(&d.minor, cudaDevAttrComputeCapabilityMinor, dev));
    ERN_CUDA(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev));
    cache = d;
  }
  *out = cache;
  if (cache.major != 10) {
    set_error("device %d is sm_%d%d; libern_b200 only runs on sm_100-class (B200) GPUs and has no fallback", dev,
              cache.major, cache.minor);
    return ERN_ERR_DEVICE;
  }
  return ERN_OK;
}

static size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

struct SimWorkspace {
  uint64_t* prefix;
  uint64_t* segs;
  int32_t* prev_counts;
  int32_t* seg_counts;
  uint32_t* thr_ord;
  int32_t* sel_flags;
  size_t bytes;
};
// One query batch (<= ERN_QUERY_BATCH queries) at a time goes through the launch schedule; the workspace holds the
// candidate storage of one batch: per query a 256-slot prefix and n_seg segments of seg_cap slots (CandidateSink).
static SimWorkspace carve_sim(void* base, int64_t batch, int n_seg, int seg_cap) {
  SimWorkspace w;
  uint8_t* p = static_cast<uint8_t*>(base);
  size_t off = 0;
  w.prefix = reinterpret_cast<uint64_t*>(p + off);
  off += align256(static_cast<size_t>(batch) * ERN_DENSE_ROWS * 8);
  w.segs = reinterpret_cast<uint64_t*>(p + off);
  off += align256(static_cast<size_t>(batch) * n_seg * seg_cap * 8);
  w.prev_counts = reinterpret_cast<int32_t*>(p + off);
  off += align256(static_cast<size_t>(batch) * 4);
  w.seg_counts = reinterpret_cast<int32_t*>(p + off);
  off += align256(static_cast<size_t>(batch) * n_seg * 4);
  w.thr_ord = reinterpret_cast<uint32_t*>(p + off);
  off += align256(static_cast<size_t>(batch) * 4);
  w.sel_flags = reinterpret_cast<int32_t*>(p + off);
  off += align256(static_cast<size_t>(batch) * 4);
  w.bytes = off;
  return w;
}
// slots per segment: the survivors of a pruning pass (<= 256) plus every score of one 256-row gallery tile
static int seg_cap_for(int k) { (void)k; return ERN_SEG_CAP; }

// most gallery rows of one tensor-core scoring launch (ERN_LAUNCH_MAX_ROWS overrides; 0 = no limit)
static int64_t launch_max_rows() {
  static const int64_t v = [] {
    const char* e = getenv("ERN_LAUNCH_MAX_ROWS");
    const long long x = e ? atoll(e) : (1ll << 21);
    return x <= 0 ? (1ll << 62) : static_cast<int64_t>(x);
  }();
  return v;
}

// test hook: ERN_FORCE_SINGLE_CTA=1 makes the tensor-core path use the 1-CTA kernel even for large batches
static int force_single() {
  static const int v = [] {
    const char* e = getenv("ERN_FORCE_SINGLE_CTA");
    return (e && e[0] == '1') ? 1 : 0;
  }();
  return v;
}

}  // namespace ern

using namespace ern;

extern "C" {

int ern_version(void) { return 100; }

const char* ern_last_error(void) { return g_err; }

int ern_device_check(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount");
  if (device < 0 || device >= n) {
    set_error("device %d does not exist (%d visible)", device, n);
    return ERN_ERR_DEVICE;
  }
  int major = 0;
  ERN_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  if (major != 10) {
    set_error("device %d is not sm_100-class (major %d)", device, major);
    return ERN_ERR_DEVICE;
  }
  return ERN_OK;
}

int ern_l2norm_rows(const float* x_dev, int64_t rows, int dim, int64_t ldx, int normalize, float* out_f32_dev,
                    int64_t ld_f32, void* out_bf16_dev, int64_t ld_bf16, void* stream) {
  DeviceInfo di;
  int rc = current_device(&di);
  if (rc) return rc;
  ERN_REQUIRE(x_dev && rows >= 0 && dim > 0 && ldx >= dim, "bad input matrix");
  ERN_REQUIRE(out_f32_dev || out_bf16_dev, "no output requested");
  return launch_l2norm_rows(x_dev, rows, dim, ldx, normalize, out_f32_dev, ld_f32, out_bf16_dev, ld_bf16,
                            static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------------------------------------------------
size_t ern_combiner_packed_bytes(int dim) { return combiner::packed_bytes(dim); }

int ern_combiner_pack(const ern_combiner_weights* w, int dim, void* packed_dev, void* stream) {
  DeviceInfo di;
  int rc = current_device(&di);
  if (rc) return rc;
  ERN_REQUIRE(w && packed_dev && dim > 0, "bad arguments");
  ERN_REQUIRE(w->w_text && w->b_text && w->w_image && w->b_image && w->w_hid && w->b_hid && w->w_gate && w->b_gate,
              "all eight parameter tensors are required");
  return combiner::pack(w, dim, packed_dev, static_cast<cudaStream_t>(stream));
}

size_t ern_combiner_workspace_bytes(int64_t rows, int dim, int mode) {
  if (rows < 0 || dim <= 0) return 0;
  return mode == ERN_MODE_FP32 ? combiner::workspace_bytes_f32(rows, dim) : combiner::workspace_bytes_bf16(rows, dim);
}

int ern_combiner_forward(const ern_combiner_weights* w, int dim, int mode, const float* image_dev,
                         const float* text_dev, int64_t rows, float* out_f32_dev, void* out_bf16_dev,
                         int64_t ld_bf16, float* gate_dev, void* workspace_dev, size_t workspace_bytes,
                         void* stream) {
  DeviceInfo di;
  int rc = current_device(&di);
  if (rc) return rc;
  ERN_REQUIRE(w && image_dev && text_dev && rows >= 0 && dim > 0, "bad arguments");
  ERN_REQUIRE(out_f32_dev || out_bf16_dev, "no output requested");
  ERN_REQUIRE(!out_bf16_dev || ld_bf16 >= dim, "ld_bf16 < dim");
  if (workspace_bytes < ern_combiner_workspace_bytes(rows, dim, mode) || (!workspace_dev && rows > 0)) {
    set_error("workspace too small: %zu < %zu", workspace_bytes, ern_combiner_workspace_bytes(rows, dim, mode));
    return ERN_ERR_WORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (mode == ERN_MODE_FP32) {
    ERN_REQUIRE(w->w_text && w->b_text && w->w_image && w->b_image && w->w_hid && w->b_hid && w->w_gate && w->b_gate,
                "fp32 mode needs the eight fp32 parameter tensors");
    return combiner::forward_f32(w, dim, image_dev, text_dev, rows, out_f32_dev, out_bf16_dev, ld_bf16, gate_dev,
                                 workspace_dev, st);
  }
  if (mode == ERN_MODE_BF16) {
    ERN_REQUIRE(w->packed_bf16, "bf16 mode needs packed weights (ern_combiner_pack)");
    if (dim % 64 != 0) {
      set_error("bf16 combiner needs dim %% 64 == 0 (got %d)", dim);
      return ERN_ERR_UNSUPPORTED;
    }
    return combiner::forward_bf16(w, dim, image_dev, text_dev, rows, out_f32_dev, out_bf16_dev, ld_bf16, gate_dev,
                                  workspace_dev, di.sm_count, st);
  }
  set_error("unknown mode %d", mode);
  return ERN_ERR_ARG;
}

// ---------------------------------------------------------------------------------------------------------
size_t ern_dvr_packed_bytes(int dim, int intermediate, int n_layers) {
  if (dim <= 0 || intermediate <= 0 || n_layers < 0) return 0;
  return dvr::packed_bytes(dim, intermediate, n_layers);
}

static int check_dvr_weights(const ern_dvr_weights* w) {
  ERN_REQUIRE(w && w->cls_token && w->pos_emb && w->type_emb && w->emb_ln_w && w->emb_ln_b, "embedding tensors missing");
  ERN_REQUIRE(w->n_layers >= 0 && w->n_layers <= ERN_MAX_BERT_LAYERS, "n_layers must be in [0,%d]", ERN_MAX_BERT_LAYERS);
  ERN_REQUIRE(w->intermediate > 0, "intermediate size missing");
  for (int i = 0; i < w->n_layers; ++i) {
    const ern_bert_layer_weights& l = w->layers[i];
    ERN_REQUIRE(l.wq && l.bq && l.wk && l.bk && l.wv && l.bv && l.wo && l.bo && l.ln1_w && l.ln1_b && l.wi && l.bi &&
                    l.wo2 && l.bo2 && l.ln2_w && l.ln2_b,
                "layer %d: parameter tensor missing", i);
  }
  ERN_REQUIRE(w->mha_in_w && w->mha_in_b && w->mha_out_w && w->mha_out_b, "MR_component tensors missing");
  return ERN_OK;
}

int ern_dvr_pack(const ern_dvr_weights* w, int dim, void* packed_dev, void* stream) {
  DeviceInfo di;
  int rc = current_device(&di);
  if (rc) return rc;
  if ((rc = check_dvr_weights(w))) return rc;
  ERN_REQUIRE(packed_dev && dim > 0, "bad arguments");
  return dvr::pack(w, dim, packed_dev, static_cast<cudaStream_t>(stream));
}

size_t ern_dvr_workspace_bytes(int64_t batch, int patches, int tokens, int dim, int intermediate, int mode) {
  if (batch < 0 || patches <= 0 || tokens <= 0 || dim <= 0 || intermediate <= 0) return 0;
  return dvr::workspace_bytes(batch, patches, tokens, dim, intermediate, mode);
}

int ern_dvr_encode(const ern_dvr_weights* w, int dim, int heads, int patches, int tokens, int mode,
                   const float* patches_dev, const float* tokens_dev, int64_t batch, float* out_cross_dev,
                   float* out_seq_mean_dev, void* workspace_dev, size_t workspace_bytes, void* stream) {
  DeviceInfo di;
  int rc = current_device(&di);
  if (rc) return rc;
  if ((rc = check_dvr_weights(w))) return rc;
  ERN_REQUIRE(patches_dev && tokens_dev && out_cross_dev && out_seq_mean_dev && batch >= 0, "bad arguments");
  ERN_REQUIRE(dim > 0 && dim <= 1024 && heads > 0 && dim % heads == 0, "dim must be <= 1024 and divisible by heads");
  ERN_REQUIRE(patches >= 1 && tokens >= patches && 1 + patches + tokens <= 512,
              "need 1 <= patches <= tokens and 1 + patches + tokens <= 512 (position table)");
  ERN_REQUIRE(mode == ERN_MODE_FP32 || mode == ERN_MODE_BF16, "unknown mode %d", mode);
  if (mode == ERN_MODE_BF16) {
    ERN_REQUIRE(w->packed_bf16, "bf16 mode needs packed weights (ern_dvr_pack)");
    if (dim % 128 != 0 || w->intermediate % 128 != 0) {
      set_error("bf16 DVR encoder needs dim %% 128 == 0 and intermediate %% 128 == 0 (got %d, %d)", dim, w->intermediate);
      return ERN_ERR_UNSUPPORTED;
    }
  }
  const size_t need = ern_dvr_workspace_bytes(batch, patches, tokens, dim, w->intermediate, mode);
  if (workspace_bytes < need || (!workspace_dev && batch > 0)) {
    set_error("workspace too small: %zu < %zu", workspace_bytes, need);
    return ERN_ERR_WORKSPACE;
  }
  return dvr::encode(w, dim, heads, patches, tokens, mode, patches_dev, tokens_dev, batch, out_cross_dev,
                     out_seq_mean_dev, workspace_dev, di.sm_count, static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------------------------------------------------
size_t ern_visualsr_packed_bytes(int dim) { return visualsr::packed_bytes(dim); }

int ern_visualsr_pack(const ern_visualsr_weights* w, int dim, void* packed_dev, void* stream) {
  DeviceInfo di;
  int rc = current_device(&di);
  if (rc) return rc;
  ERN_REQUIRE(w && w->w_local && w->w_global && packed_dev && dim > 0, "bad arguments");
  return visualsr::pack(w, dim, packed_dev, static_cast<cudaStream_t>(stream));
}

size_t ern_visualsr_workspace_bytes(int64_t rows, int patches, int dim, int mode) {
  if (rows < 0 || patches <= 0 || dim <= 0) return 0;
  return visualsr::workspace_bytes(rows, patches, dim, mode);
}

int ern_visualsr_forward(const ern_visualsr_weights* w, int dim, int patches, int mode, const float* local_dev,
                         int64_t rows, float* out_f32_dev, void* workspace_dev, size_t workspace_bytes, void* stream) {
  DeviceInfo di;
  int rc = current_device(&di);
  if (rc) return rc;
  ERN_REQUIRE(w && local_dev && out_f32_dev && rows >= 0 && dim > 0, "bad arguments");
  ERN_REQUIRE(patches >= 1 && patches <= 32, "patches must be in [1,32]");
  ERN_REQUIRE(w->w_local && w->b_local && w->bn_local_scale && w->bn_local_shift && w->w_global && w->b_global &&
                  w->bn_global_scale && w->bn_global_shift && w->w_common && w->b_common,
              "all ten parameter tensors are required");
  if (workspace_bytes < ern_visualsr_workspace_bytes(rows, patches, dim, mode) || (!workspace_dev && rows > 0)) {
    set_error("workspace too small: %zu < %zu", workspace_bytes, ern_visualsr_workspace_bytes(rows, patches, dim, mode));
    return ERN_ERR_WORKSPACE;
  }
  if (mode == ERN_MODE_BF16) {
    ERN_REQUIRE(w->packed_bf16, "bf16 mode needs packed weights (ern_visualsr_pack)");
    if (dim % 128 != 0) {
      set_error("bf16 VisualSR needs dim %% 128 == 0 (got %d)", dim);
      return ERN_ERR_UNSUPPORTED;
    }
  } else if (mode != ERN_MODE_FP32) {
    set_error("unknown mode %d", mode);
    return ERN_ERR_ARG;
  }
  return visualsr::forward(w, dim, patches, mode, local_dev, rows, out_f32_dev, workspace_dev, di.sm_count,
                           static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------------------------------------------------
static int sm_count_or_default() {
  int dev = -1, sms = 0;
  if (cudaGetDevice(&dev) == cudaSuccess &&
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
    return sms;
  (void)cudaGetLastError();
  return 148;   // B200
}

size_t ern_sim_topk_workspace_bytes(int64_t nq, int dim, int mode) {
  (void)dim;
  (void)mode;
  if (nq < 0) return 0;
  // two batch shapes occur: full batches and the remainder (which may run on the 1-CTA kernel with more units)
  const int sms = sm_count_or_default();
  const int64_t full = nq < ERN_QUERY_BATCH ? nq : ERN_QUERY_BATCH;
  const int64_t rem = nq % ERN_QUERY_BATCH;
  size_t bytes = carve_sim(nullptr, full, simtc::units_for(full, force_single(), sms), ERN_SEG_CAP).bytes;
  if (rem > 0) {
    const size_t b2 = carve_sim(nullptr, rem, simtc::units_for(rem, force_single(), sms), ERN_SEG_CAP).bytes;
    if (b2 > bytes) bytes = b2;
  }
  return bytes + 256;
}

static int sim_topk_impl(const void* queries_dev, int64_t nq, int64_t ldq, const void* gallery_dev, int64_t n_rows,
                         int64_t ldg, int dim, int dtype, int64_t id_offset, const int32_t* exclude_id_dev, int k,
                         int mode, int rank_by, int growth, float* out_scores_dev, int32_t* out_ids_dev,
                         uint64_t* out_keys_dev, uint64_t* const* peer_keys_dev, int world, int rank,
                         int32_t* status_dev, void* workspace_dev, size_t workspace_bytes, void* stream) {
  DeviceInfo di;
  int rc = current_device(&di);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ERN_REQUIRE(nq >= 0 && n_rows >= 0 && dim > 0, "negative sizes");
  ERN_REQUIRE(k >= 1 && k <= ERN_MAX_K, "k must be in [1,%d] (got %d)", ERN_MAX_K, k);
  ERN_REQUIRE(rank_by == ERN_RANK_SIMILARITY || rank_by == ERN_RANK_REFERENCE, "bad rank_by %d", rank_by);
  ERN_REQUIRE(growth >= 1 && growth <= 64, "growth must be in [1,64]");
  ERN_REQUIRE(status_dev != nullptr, "status_dev is required");
  ERN_REQUIRE(out_scores_dev || out_ids_dev || out_keys_dev || peer_keys_dev, "no output requested");
  ERN_REQUIRE(!peer_keys_dev || (world >= 1 && rank >= 0 && rank < world), "bad world/rank for the fused exchange");
  ERN_REQUIRE(id_offset >= 0 && id_offset + n_rows <= 0x7FFFFFFFll, "global ids must fit int32");
  ERN_REQUIRE((mode == ERN_MODE_FP32 && dtype == ERN_DTYPE_F32) ||
                  (mode == ERN_MODE_BF16 && (dtype == ERN_DTYPE_BF16 || dtype == ERN_DTYPE_F16)),
              "mode/dtype mismatch: FP32 mode takes f32 features, BF16 (tensor-core) mode takes bf16 or fp16 features");
  if (nq == 0) return ERN_OK;
  ERN_REQUIRE(queries_dev && (gallery_dev || n_rows == 0) && ldq >= dim && ldg >= dim, "bad feature matrices");
  if (workspace_bytes < ern_sim_topk_workspace_bytes(nq, dim, mode) || !workspace_dev) {
    set_error("workspace too small: %zu < %zu", workspace_bytes, ern_sim_topk_workspace_bytes(nq, dim, mode));
    return ERN_ERR_WORKSPACE;
  }
  const int esize = mode == ERN_MODE_BF16 ? 2 : 4;
  if (mode == ERN_MODE_BF16) {
    if (dim % 64 != 0 || dim > 768) {
      set_error("bf16 scoring needs dim %% 64 == 0 (zero-pad the features) and dim <= 768 (got %d)", dim);
      return ERN_ERR_UNSUPPORTED;
    }
    ERN_REQUIRE((reinterpret_cast<uintptr_t>(queries_dev) & 15) == 0 && (reinterpret_cast<uintptr_t>(gallery_dev) & 15) == 0 &&
                    (ldq * 2) % 16 == 0 && (ldg * 2) % 16 == 0,
                "bf16 feature rows must be 16-byte aligned");
  }
  CUtensorMap tq, tg;
  if (mode == ERN_MODE_BF16 && n_rows > 0) {
    rc = simtc::make_tmap_bf16_rows(&tg, gallery_dev, n_rows, dim, ldg);
    if (rc) return rc;
  }
  const int seg_cap = seg_cap_for(k);

  // ---- query batches: each runs the whole launch schedule over the shard ------------------------------------------
  for (int64_t q0 = 0; q0 < nq; q0 += ERN_QUERY_BATCH) {
    const int64_t bq = nq - q0 < ERN_QUERY_BATCH ? nq - q0 : ERN_QUERY_BATCH;
    const int n_seg = simtc::units_for(bq, force_single(), di.sm_count);
    ERN_REQUIRE(n_seg <= 256, "internal: %d scoring units exceed the selection kernel's segment table", n_seg);
    SimWorkspace ws = carve_sim(workspace_dev, bq, n_seg, seg_cap);
    const uint8_t* qbase = static_cast<const uint8_t*>(queries_dev) + static_cast<size_t>(q0) * ldq * esize;
    if (mode == ERN_MODE_BF16) {
      rc = simtc::make_tmap_bf16_rows(&tq, qbase, bq, dim, ldq);
      if (rc) return rc;
    }
    rc = launch_init_state(ws.prev_counts, ws.seg_counts, ws.thr_ord, bq, n_seg, q0 == 0 ? status_dev : nullptr, st);
    if (rc) return rc;

    CandidateSink sink;
    memset(&sink, 0, sizeof(sink));
    sink.prefix = ws.prefix;
    sink.segs = ws.segs;
    sink.seg_counts = ws.seg_counts;
    sink.thr_ord = ws.thr_ord;
    sink.exclude = exclude_id_dev ? exclude_id_dev + q0 : nullptr;
    sink.status = status_dev;
    sink.n_seg = n_seg;
    sink.seg_cap = seg_cap;
    sink.k = k;
    sink.id_offset = id_offset;
    sink.nq = bq;

    SelectParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.prefix = ws.prefix;
    sp.prev_counts = ws.prev_counts;
    sp.segs = ws.segs;
    sp.seg_counts = ws.seg_counts;
    sp.n_seg = n_seg;
    sp.seg_cap = seg_cap;
    sp.single_segment = mode == ERN_MODE_FP32 ? 1 : 0;
    sp.k = k;
    sp.thr_ord = ws.thr_ord;
    sp.sel_flags = mode == ERN_MODE_BF16 ? ws.sel_flags : nullptr;
    sp.status = status_dev;
    sp.world = world;
    sp.rank = rank;
    sp.nq_total = nq;
    sp.q_first = q0;

    // Launch schedule.  Launch 0 keeps every score of the first ERN_DENSE_ROWS rows; each later launch covers rows
    // [b, growth*b) starting from the exact k-th best of rows [0, b) as the threshold, so on an unordered gallery
    // about (growth-1)*k candidates per query reach the selection kernel after every launch, independent of the
    // gallery size.  The schedule is a cost heuristic only: the tensor-core kernel bounds its segments itself
    // (in-kernel compaction), so every launch is exact whatever the order of the rows.  The fp32 validation kernel
    // has no compaction: its launches cover at most as many rows as a query's candidate storage has slots.
    // growth == 1: fixed steps of ERN_SORT_CAP - k rows (test hook: many small launches).
    const int64_t f32_rows_max = static_cast<int64_t>(n_seg) * seg_cap;
    int64_t begin = 0;
    bool first = true;
    do {
      int64_t end;
      if (first) {
        end = ERN_DENSE_ROWS;
      } else if (growth == 1) {
        end = begin + (ERN_SORT_CAP - k);
      } else {
        end = begin * growth;
      }
      if (!first && mode == ERN_MODE_FP32 && end - begin > f32_rows_max) end = begin + f32_rows_max;
      // Long launches let the query tiles that share a super tile drift apart (static round-robin items, ~775 per unit
      // on a 58.7M-row launch): ncu then shows every gallery line fetched twice from DRAM (lts hit rate 87.5 % = 14/16).
      // Cutting a range into launches of at most 2M rows re-aligns the units every ~28 items, and the exact selection
      // after every launch keeps the thresholds tight on ordered galleries (10M clustered rows: 44.0 ms per call with
      // 8.4M-row launches, 40.4 ms with 2M; random order: 35.0 ms either way; 100M random: 357.5 vs 360.5 ms).
      if (!first && mode == ERN_MODE_BF16 && end - begin > launch_max_rows()) end = begin + launch_max_rows();
      if (end > n_rows) end = n_rows;
      sink.dense = first ? 1 : 0;
      sink.row_begin = begin;
      sink.row_end = end;
      if (end > begin) {
        if (mode == ERN_MODE_FP32) {
          rc = simf32::launch(reinterpret_cast<const float*>(qbase), ldq, static_cast<const float*>(gallery_dev), ldg,
                              dim, sink, rank_by, st);
        } else {
          rc = simtc::launch(tq, tg, sink, dim, rank_by, force_single(), di.sm_count, dtype == ERN_DTYPE_F16, st);
        }
        if (rc) return rc;
      }
      const bool last = end >= n_rows;
      sp.dense_count = first ? static_cast<int>(end - begin) : 0;
      sp.out_scores = (last && out_scores_dev) ? out_scores_dev + q0 * k : nullptr;
      sp.out_ids = (last && out_ids_dev) ? out_ids_dev + q0 * k : nullptr;
      sp.out_keys = (last && out_keys_dev) ? out_keys_dev + q0 * k : nullptr;
      sp.peer_keys = last ? peer_keys_dev : nullptr;
      rc = launch_select(sp, bq, st);
      if (rc) return rc;
      begin = end;
      first = false;
    } while (begin < n_rows);
  }
  return ERN_OK;
}

int ern_sim_topk(const void* queries_dev, int64_t nq, int64_t ldq, const void* gallery_dev, int64_t n_rows,
                 int64_t ldg, int dim, int dtype, int64_t id_offset, const int32_t* exclude_id_dev, int k, int mode,
                 int rank_by, int growth, float* out_scores_dev, int32_t* out_ids_dev, uint64_t* out_keys_dev,
                 int32_t* status_dev, void* workspace_dev, size_t workspace_bytes, void* stream) {
  return sim_topk_impl(queries_dev, nq, ldq, gallery_dev, n_rows, ldg, dim, dtype, id_offset, exclude_id_dev, k, mode,
                       rank_by, growth, out_scores_dev, out_ids_dev, out_keys_dev, nullptr, 0, 0, status_dev,
                       workspace_dev, workspace_bytes, stream);
}

int ern_sim_topk_exchange(const void* queries_dev, int64_t nq, int64_t ldq, const void* gallery_dev, int64_t n_rows,
                          int64_t ldg, int dim, int dtype, int64_t id_offset, const int32_t* exclude_id_dev, int k,
                          int mode, int rank_by, int growth, uint64_t* const* peer_keys_dev, int world, int rank,
                          int32_t* status_dev, void* workspace_dev, size_t workspace_bytes, void* stream) {
  ERN_REQUIRE(peer_keys_dev != nullptr, "peer_keys_dev is required");
  return sim_topk_impl(queries_dev, nq, ldq, gallery_dev, n_rows, ldg, dim, dtype, id_offset, exclude_id_dev, k, mode,
                       rank_by, growth, nullptr, nullptr, nullptr, peer_keys_dev, world, rank, status_dev,
                       workspace_dev, workspace_bytes, stream);
}

int ern_topk_merge(const uint64_t* keys_dev, int64_t nq, int n_lists, int k_in, int64_t list_stride,
                   int64_t query_stride, int k_out, float* out_scores_dev, int32_t* out_ids_dev, uint64_t* out_keys_dev,
                   void* stream) {
  DeviceInfo di;
  int rc = current_device(&di);
  if (rc) return rc;
  ERN_REQUIRE(keys_dev && nq >= 0 && n_lists >= 1 && k_in >= 1, "bad arguments");
  ERN_REQUIRE(static_cast<int64_t>(n_lists) * k_in <= ERN_SORT_CAP, "n_lists * k_in must be <= %d", ERN_SORT_CAP);
  ERN_REQUIRE(k_out >= 1 && k_out <= ERN_MAX_K, "k_out must be in [1,%d]", ERN_MAX_K);
  SelectParams sp;
  memset(&sp, 0, sizeof(sp));
  sp.merge_src = keys_dev;
  sp.list_stride = list_stride;
  sp.query_stride = query_stride;
  sp.n_lists = n_lists;
  sp.k_in = k_in;
  sp.k = k_out;
  sp.out_scores = out_scores_dev;
  sp.out_ids = out_ids_dev;
  sp.out_keys = out_keys_dev;
  sp.nq_total = nq;
  return launch_select(sp, nq, static_cast<cudaStream_t>(stream));
}

int ern_recall_at_k(const int32_t* top_ids_dev, int64_t nq, int k, const int32_t* class_of_dev, int64_t n_gallery,
                    const int32_t* target_class_dev, const int32_t* ks, int nk, int32_t* counts_dev, int32_t* rank_dev,
                    void* stream) {
  DeviceInfo di;
  int rc = current_device(&di);
  if (rc) return rc;
  ERN_REQUIRE(top_ids_dev && class_of_dev && target_class_dev && ks && counts_dev && k >= 1 && nq >= 0, "bad arguments");
  return launch_recall(top_ids_dev, nq, k, class_of_dev, n_gallery, target_class_dev, ks, nk, counts_dev, rank_dev,
                       static_cast<cudaStream_t>(stream));
}

int ern_cirr_subset_recall(const void* queries_dev, int64_t nq, int64_t ldq, const void* gallery_dev, int64_t n_rows,
                           int64_t ldg, int dim, int dtype, const int32_t* members_dev, int m,
                           const int32_t* reference_id_dev, const int32_t* target_id_dev, int rank_by,
                           const int32_t* ks, int nk, int32_t* counts_dev, int32_t* rank_dev, void* stream) {
  DeviceInfo di;
  int rc = current_device(&di);
  if (rc) return rc;
  ERN_REQUIRE(queries_dev && gallery_dev && members_dev && reference_id_dev && target_id_dev && ks && counts_dev,
              "bad arguments");
  ERN_REQUIRE(dtype == ERN_DTYPE_F32 || dtype == ERN_DTYPE_BF16 || dtype == ERN_DTYPE_F16, "bad dtype");
  return launch_cirr_subset(queries_dev, nq, ldq, gallery_dev, n_rows, ldg, dim, dtype, members_dev, m,
                            reference_id_dev, target_id_dev, rank_by, ks, nk, counts_dev, rank_dev,
                            static_cast<cudaStream_t>(stream));
}

int ern_gather_scores(const void* queries_dev, int64_t nq, int64_t ldq, const void* gallery_dev, int64_t n_rows,
                      int64_t ldg, int dim, int dtype, int64_t id_offset, const int32_t* ids_dev, int m,
                      float* out_scores_dev, void* stream) {
  DeviceInfo di;
  int rc = current_device(&di);
  if (rc) return rc;
  ERN_REQUIRE(queries_dev && (gallery_dev || n_rows == 0) && ids_dev && out_scores_dev, "bad arguments");
  ERN_REQUIRE(dtype == ERN_DTYPE_F32 || dtype == ERN_DTYPE_BF16 || dtype == ERN_DTYPE_F16, "bad dtype");
  return launch_gather_scores(queries_dev, nq, ldq, gallery_dev, n_rows, ldg, dim, dtype, id_offset, ids_dev, m,
                              out_scores_dev, static_cast<cudaStream_t>(stream));
}

int ern_cirr_subset_from_scores(const float* scores_dev, int64_t nq, const int32_t* members_dev, int m,
                                const int32_t* reference_id_dev, const int32_t* target_id_dev, int rank_by,
                                const int32_t* ks, int nk, int32_t* counts_dev, int32_t* rank_dev, void* stream) {
  DeviceInfo di;
  int rc = current_device(&di);
  if (rc) return rc;
  ERN_REQUIRE(scores_dev && members_dev && reference_id_dev && target_id_dev && ks && counts_dev, "bad arguments");
  return launch_cirr_rank(scores_dev, nq, members_dev, m, reference_id_dev, target_id_dev, rank_by, ks, nk, counts_dev,
                          rank_dev, static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------------------------------------------------
size_t ern_bbc_loss_workspace_bytes(int64_t batch, int dim, int mode) {
  if (batch < 0 || dim <= 0) return 0;
  return bbcloss::workspace_bytes(batch, dim, mode);
}

static int bbc_check(const float* pred, int64_t ldp, const float* tar, int64_t ldt, int64_t batch, int dim, int mode,
                     const void* workspace, size_t workspace_bytes) {
  ERN_REQUIRE(pred && tar && batch >= 1 && dim > 0 && ldp >= dim && ldt >= dim, "bad arguments");
  ERN_REQUIRE(batch <= (1 << 20), "batch too large (%lld)", (long long)batch);
  if (mode != ERN_MODE_BF16 && mode != ERN_MODE_FP32) {
    set_error("unknown mode %d", mode);
    return ERN_ERR_ARG;
  }
  if (mode == ERN_MODE_BF16 && dim % 64 != 0) {
    set_error("bf16 loss needs dim %% 64 == 0 (got %d)", dim);
    return ERN_ERR_UNSUPPORTED;
  }
  const size_t need = bbcloss::workspace_bytes(batch, dim, mode);
  if (workspace_bytes < need || !workspace) {
    set_error("workspace too small: %zu < %zu", workspace_bytes, need);
    return ERN_ERR_WORKSPACE;
  }
  return ERN_OK;
}

int ern_bbc_loss_forward(const float* pred_dev, int64_t ldp, const float* tar_dev, int64_t ldt, int64_t batch, int dim,
                         float scale, int mode, float* loss_dev, float* lse_dev, void* workspace_dev,
                         size_t workspace_bytes, void* stream) {
  DeviceInfo di;
  int rc = current_device(&di);
  if (rc) return rc;
  if ((rc = bbc_check(pred_dev, ldp, tar_dev, ldt, batch, dim, mode, workspace_dev, workspace_bytes))) return rc;
  ERN_REQUIRE(loss_dev, "loss output is required");
  return bbcloss::forward(pred_dev, ldp, tar_dev, ldt, batch, dim, scale, mode, loss_dev, lse_dev, workspace_dev,
                          di.sm_count, static_cast<cudaStream_t>(stream));
}

int ern_bbc_loss_backward(const float* pred_dev, int64_t ldp, const float* tar_dev, int64_t ldt, int64_t batch, int dim,
                          float scale, int mode, const float* lse_dev, const float* grad_out_dev, float* dpred_dev,
                          int64_t lddp, float* dtar_dev, int64_t lddt, void* workspace_dev, size_t workspace_bytes,
                          void* stream) {
  DeviceInfo di;
  int rc = current_device(&di);
  if (rc) return rc;
  if ((rc = bbc_check(pred_dev, ldp, tar_dev, ldt, batch, dim, mode, workspace_dev, workspace_bytes))) return rc;
  ERN_REQUIRE(lse_dev && dpred_dev && dtar_dev, "lse and both gradient outputs are required");
  ERN_REQUIRE(lddp >= dim && lddt >= dim && lddp % 4 == 0 && lddt % 4 == 0, "gradient row strides must be >= dim and multiples of 4");
  return bbcloss::backward(pred_dev, ldp, tar_dev, ldt, batch, dim, scale, mode, lse_dev, grad_out_dev, dpred_dev, lddp,
                           dtar_dev, lddt, workspace_dev, di.sm_count, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
