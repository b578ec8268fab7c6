/*
 * dvl_b200.h -- C ABI of the B200-native bias-measurement hot path.
 *
 * The reference (oxai/debias-vision-lang) has no FFI of its own: its hot path is plain
 * Python (debias_clip/measuring_bias.py).  This header is the boundary a drop-in
 * replacement binds instead: each entry point names the reference lines it replaces.
 * Plain C types only (no torch, no C++), every buffer is owned by the caller, the
 * library owns only the workspace inside its context.  All functions return DVL_OK (0)
 * or a negative dvl_status; nothing throws across the boundary; the text of the last
 * failure is kept per context (dvl_last_error).  A context is bound to one device and
 * is not thread-safe; use one per device/thread.  There is no CPU fallback: without an
 * sm_100 device every compute entry point fails with DVL_ERR_NO_DEVICE.
 *
 * Hot path:
 *   a1  query L2-normalisation            measuring_bias.py:79,81
 *   a2  image x query similarity          measuring_bias.py:114
 *   a3  ranking, nlargest(keep="all")     measuring_bias.py:29,57 (pandas selectn.py:106-190)
 *   a4  global label histogram            measuring_bias.py:23,56
 *   a5  MaxSkew@k                         measuring_bias.py:47-72
 *   a6  NDKL                              measuring_bias.py:17-44
 *   a7  mean over queries                 measuring_bias.py:117,119
 */
#ifndef DVL_B200_H_
#define DVL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DVL_B200_ABI_VERSION 1

#if defined(__GNUC__)
#define DVL_API __attribute__((visibility("default")))
#else
#define DVL_API
#endif

typedef struct dvl_ctx dvl_ctx;

typedef enum dvl_status {
  DVL_OK = 0,
  DVL_ERR_INVALID = -1,      /* bad argument (the reference raises AssertionError / IndexError) */
  DVL_ERR_NO_DEVICE = -2,    /* no CUDA device, or not compute capability 10.x */
  DVL_ERR_CUDA = -3,         /* a CUDA runtime / driver call failed; see dvl_last_error */
  DVL_ERR_OOM = -4,          /* workspace allocation failed */
  DVL_ERR_UNSUPPORTED = -5,  /* shape outside what the kernels take (e.g. D > 8192) */
  DVL_ERR_OVERFLOW = -6      /* a caller-provided output buffer is too small */
} dvl_status;

/* element type of an embedding matrix handed in by the caller */
typedef enum dvl_dtype { DVL_F32 = 0, DVL_BF16 = 1, DVL_F16 = 2 } dvl_dtype;

/* arithmetic of the similarity stage (a2).
 * FP32:   every row is scaled by a power of two and split into two fp16 planes (22+ significand bits);
 *         the three largest product terms are accumulated in fp32 on the tensor cores.  Error
 *         <= ~3*2^-22 of |g||t| -- that of an fp32 dot product -- and exact on data whose dot products
 *         are exactly representable (whatever the summation order).
 * BF16:   operands rounded once to bf16, fp32 accumulation (2e-2 tolerance of BASELINE.json).
 * FP32X3: every operand split exactly into three bf16 planes, six product terms: error ~2^-24 of
 *         |g||t| over the whole fp32 range, at twice the tensor work of FP32. */
typedef enum dvl_precision { DVL_PRECISION_FP32 = 0, DVL_PRECISION_BF16 = 1, DVL_PRECISION_FP32X3 = 2 } dvl_precision;

/* flags */
#define DVL_FLAG_NORMALIZE_QUERIES 0x1u /* a1: divide each query row by its L2 norm (fused into the
                                           similarity epilogue).  Off when the caller already did
                                           measuring_bias.py:79. */
#define DVL_FLAG_NORMALIZE_GALLERY 0x2u /* NOT what eval_ranking does: cosine mode */
#define DVL_FLAG_FULL_SORT         0x4u /* force the materialise-and-sort ranking route even when
                                           top_n << N (the route top_n >= N always takes) */
#define DVL_FLAG_NO_PILOT          0x10u /* streaming route: do not take an optimistic first threshold
                                           from a pilot sample (large galleries); use the progressive
                                           chunk-by-chunk scheme only */
#define DVL_FLAG_RADIX_SORT        0x20u /* full-sort route: rank with the stable LSD radix sort (segsort.cu) instead
                                           of the score-linear bucket sort (bucketsort.cu) that is tried first;
                                           same results, used for A/B measurements and tests */
#define DVL_FLAG_REUSE_GALLERY     0x40u /* dvl_similarity only: `gallery` is the matrix the preceding dvl_shard_topk on this
                                           context prepared, unchanged since (the caller vouches for it): its
                                           conversion into tensor-core operands is not repeated.  Used by the exact
                                           re-rank of flagged queries on the sharded path */
#define DVL_FLAG_STREAM_TOPK       0x8u /* force the streaming top-k route even on a small gallery
                                           (chunks of N/4 rows); needs top_n < N and top_n <= 8192 */

/* per-query result layout: 4 doubles per query */
#define DVL_MAXSKEW_EQ_OPP 0
#define DVL_MAXSKEW_DEM_PAR 1
#define DVL_NDKL_EQ_OPP 2
#define DVL_NDKL_DEM_PAR 3

#define DVL_MAX_CLASSES 16

/* stages reported by dvl_get_profile */
#define DVL_STAGE_PREP 0         /* operand split / conversion, label histogram, ln table */
#define DVL_STAGE_GEMM 1         /* tcgen05 similarity GEMM (+ fused epilogue) */
#define DVL_STAGE_SORT_HIST 2    /* radix sort: tile digit histograms */
#define DVL_STAGE_SORT_SCAN 3    /* radix sort: per-segment scans */
#define DVL_STAGE_SORT_SCATTER 4 /* radix sort: stable scatter */
#define DVL_STAGE_METRICS 5      /* MaxSkew/NDKL + mean over queries */
#define DVL_STAGE_SELECT 6       /* streaming top-k: candidate selection / merge */
#define DVL_STAGE_COPY 7         /* host<->device and layout copies issued by the library */
#define DVL_NUM_STAGES 8

/* ---- life cycle ------------------------------------------------------------------- */
DVL_API int dvl_abi_version(void);
/* Binds a context to CUDA device `device` (must be sm_100).  */
DVL_API int dvl_create(dvl_ctx** out, int device);
DVL_API int dvl_destroy(dvl_ctx* ctx);
/* Text of the last failure on this context ("" if none); valid until the next call. */
DVL_API const char* dvl_last_error(const dvl_ctx* ctx);
/* Bytes of device workspace currently held by the context. */
DVL_API size_t dvl_workspace_bytes(const dvl_ctx* ctx);
/* Number of kernels this context launched since creation (bench.py's gpu_launches). */
DVL_API uint64_t dvl_launch_count(const dvl_ctx* ctx);

/* Per-stage device timing with CUDA events recorded on the caller's stream around each stage's
 * kernels.  dvl_get_profile synchronises, writes the milliseconds and kernel-launch counts
 * accumulated per stage since the previous call (arrays of DVL_NUM_STAGES) and resets them. */
DVL_API int dvl_set_profiling(dvl_ctx* ctx, int enabled);
DVL_API int dvl_get_profile(dvl_ctx* ctx, double* stage_ms, uint64_t* stage_launches);

/* ---- labels (a4 input) ---------------------------------------------------------------------
 * The reference's labels are an int64 ndarray (datasets.py:44 -> measuring_bias.py:96,99) that every
 * per-query DataFrame and Counter re-reads (measuring_bias.py:23,56,115).  This turns DEVICE-resident
 * integer labels into the uint8 class codes the path takes, in one pass, and reports what
 * Counter(df.label) would see:
 *   out_info_host [19]  0..15 rows per class; 16 n_classes = largest label + 1; 17 out_of_range (a label
 *                       < 0 or >= DVL_MAX_CLASSES: codes are not usable, remap on the host);
 *                       18 contiguous (every class 0..n_classes-1 present -- what NDKL's
 *                       top_n_label_counts[label] indexing needs, measuring_bias.py:33-34)
 * Synchronises `stream` (the info is returned on the host). */
typedef enum dvl_label_dtype { DVL_LABEL_I64 = 0, DVL_LABEL_I32 = 1, DVL_LABEL_U8 = 2 } dvl_label_dtype;
DVL_API int dvl_encode_labels(dvl_ctx* ctx, const void* labels /* dev [n] */, int label_dtype, int64_t n,
                              uint8_t* out_codes /* dev [n] */, int64_t* out_info_host /* [19] */, void* stream);

/* ---- the hot path, device-resident inputs ----------------------------------------------
 * Replaces BOTH calls `eval_ranking(labels, image_embeddings, prompts_embeddings, ev, topn)`
 * for ev in ("maxskew", "ndkl") (measuring_bias.py:153-154 -> :102-119) with one pass:
 * similarity -> rank (score desc, lowest index first, boundary ties kept) -> label
 * histograms -> MaxSkew + NDKL -> mean over queries.
 *
 *   gallery   device [n, d] row-major, gallery_dtype    (raw image embeddings, NOT normalised)
 *   queries   device [q, d] row-major, query_dtype      (unit rows unless NORMALIZE_QUERIES)
 *   labels    device [n] uint8, values 0..n_classes-1, every class present (datasets.py:43)
 *   top_n     already resolved to an int (float topn -> ceil(n*topn), measuring_bias.py:109-110)
 *   out_mean4 HOST   double[4]: mean over queries, order DVL_MAXSKEW_EQ_OPP.. (written before return;
 *             the call synchronises `stream`)
 *   out_per_query     device double [q][4] or NULL
 *   out_ranked_len    device int32 [q] or NULL: m_q = number of ranked rows (>= min(top_n,n))
 *   out_ranked_idx    device int32 [q][ranked_cap] or NULL: gallery indices in ranked order,
 *                     first min(m_q, ranked_cap) entries of each row valid
 *   out_scores        device float [q][n] or NULL: the fp32 similarities (full-sort route only)
 *   stream    cudaStream_t (as void*) everything is enqueued on; NULL = default stream
 */
DVL_API int dvl_eval_ranking(dvl_ctx* ctx,
                     const void* gallery, int gallery_dtype,
                     const void* queries, int query_dtype,
                     const uint8_t* labels,
                     int64_t n, int32_t d, int32_t q, int32_t n_classes, int64_t top_n,
                     int precision, uint32_t flags,
                     double* out_mean4,
                     double* out_per_query,
                     int32_t* out_ranked_len, int32_t* out_ranked_idx, int64_t ranked_cap,
                     float* out_scores,
                     void* stream);

/* Same computation with HOST inputs: queries and labels are copied first, the gallery follows in row
 * blocks on the context's copy stream (pinned buffers make it asynchronous) while operand split and
 * similarity already run on the blocks that have landed; returns the four means.  This is the call
 * bench.py times for the end-to-end number of the top_n = N workloads. */
DVL_API int dvl_eval_ranking_host(dvl_ctx* ctx,
                          const void* gallery_host, int gallery_dtype,
                          const void* queries_host, int query_dtype,
                          const uint8_t* labels_host,
                          int64_t n, int32_t d, int32_t q, int32_t n_classes, int64_t top_n,
                          int precision, uint32_t flags,
                          double* out_mean4, double* out_per_query_host,
                          void* stream);

/* ---- stage entry points (same kernels, exposed for stage-isolated parity tests) -------- */

/* a3-a7 on an explicit score matrix: device float [q][ld_scores] (first n columns used).
 * Replaces compute_skew_metrics + normalized_discounted_KL (measuring_bias.py:47-72, :17-44)
 * applied to each row's DataFrame{"score","label"}. */
DVL_API int dvl_rank_and_score(dvl_ctx* ctx, const float* scores, int64_t ld_scores,
                       const uint8_t* labels, int64_t n, int32_t q, int32_t n_classes, int64_t top_n,
                       double* out_mean4, double* out_per_query,
                       int32_t* out_ranked_len, int32_t* out_ranked_idx, int64_t ranked_cap,
                       void* stream);

/* a2 only: device float out [q][ld_out] = queries . gallery^T with the precision/flags above
 * (measuring_bias.py:114 for all queries at once). */
DVL_API int dvl_similarity(dvl_ctx* ctx,
                   const void* gallery, int gallery_dtype, const void* queries, int query_dtype,
                   int64_t n, int32_t d, int32_t q, int precision, uint32_t flags,
                   float* out, int64_t ld_out, void* stream);

/* ---- gallery sharded over several GPUs ---------------------------------
 * One context per rank.  Rank r holds gallery rows [index_base, index_base + n_local) (contiguous
 * blocks, so "lowest global index" is independent of the shard layout), all queries, and the labels
 * of its rows.  The host moves three small things between the stages with its own collective
 * library (torch.distributed / NCCL all-gather, all-reduce):
 *   1. dvl_shard_topk   local streaming top-k.  Keeps in the context, per query, the sorted list
 *                       of every local row that ranks within top_n locally (ties included) and
 *                       returns tau[q] = score of the local top_n-th row (-inf if fewer rows),
 *                       tie-overflow flags, and the local label histogram (17 x u64; slot 16 is
 *                       the bad-label flag).
 *      all-gather tau, take the max over ranks: no row scoring below it can rank globally;
 *      all-reduce(sum) the label histograms; all-reduce(max) the flags.
 *   2. dvl_shard_trim   cnt[q] = local entries with score >= tau_bound[q].
 *      all-reduce(max) of max_q cnt[q] gives the common packed length `len`.
 *   3. dvl_shard_pack   out [q][len] = first cnt[q] entries, padded with ~0.
 *      all-gather the packed lists into [world][q][len].
 *   4. dvl_merge_topk   for queries [q0, q0+qn): deterministic merge of the world lists (a u64
 *                       entry is key<<32 | global index<<4 | label, so the merge is a sort of
 *                       integers), m_q with boundary ties, MaxSkew + NDKL against the GLOBAL
 *                       label histogram.  out_flags[q] bit 0: tie run too long, re-rank that query
 *                       with dvl_similarity + dvl_rank_and_score.  Bit 1: a list that arrived without padding
 *                       (possibly cut to `len`) does not end strictly behind the k-th entry of the union, so
 *                       rows that could rank may be missing: repeat the exchange with longer lists.
 *      all-gather the per-query results;  5. dvl_mean gives the four means (same reduction
 *      tree as the single-GPU path, so the numbers are bit-identical to it).
 */
DVL_API int dvl_shard_topk(dvl_ctx* ctx, const void* gallery, int gallery_dtype, const void* queries, int query_dtype,
                           const uint8_t* labels, int64_t n_local, int64_t index_base, int32_t d, int32_t q,
                           int32_t n_classes, int64_t top_n, int precision, uint32_t flags,
                           float* out_tau /* dev [q] or NULL */, uint32_t* out_flags /* dev [q] or NULL */,
                           unsigned long long* out_class_counts /* dev [17] or NULL */, int32_t* out_list_cap /* host */,
                           void* stream);      /* the three device outputs are optional: dvl_shard_stats packs the same
                                                  values from the context */
/* dvl_shard_topk with HOST inputs (the shard as the embedding cache or the encoders' host copy holds it):
 * queries and labels are copied first, the gallery shard follows in row blocks on the context's copy
 * stream, and the operand split + similarity + threshold filter of a block start as soon as the block
 * has landed -- the upload (PCIe, the end-to-end bound) runs under the compute.  The device copies stay
 * valid until the next host-input call on this context; their addresses are returned so that the rare
 * exact re-rank of a flagged query (dvl_similarity on the shard) needs no second upload. */
DVL_API int dvl_shard_topk_host(dvl_ctx* ctx, const void* gallery_host, int gallery_dtype, const void* queries_host,
                                int query_dtype, const uint8_t* labels_host, int64_t n_local, int64_t index_base,
                                int32_t d, int32_t q, int32_t n_classes, int64_t top_n, int precision, uint32_t flags,
                                float* out_tau /* dev [q] or NULL */, uint32_t* out_flags /* dev [q] or NULL */,
                                unsigned long long* out_class_counts /* dev [17] or NULL */,
                                int32_t* out_list_cap /* host */, const void** out_gallery_dev, const void** out_queries_dev,
                                const uint8_t** out_labels_dev, void* stream);
DVL_API int dvl_shard_trim(dvl_ctx* ctx, const float* tau_bound /* dev [q] */, int32_t q, int32_t* out_cnt /* dev [q] */,
                           void* stream);
DVL_API int dvl_shard_pack(dvl_ctx* ctx, const int32_t* cnt /* dev [q] */, int32_t q, int64_t len,
                           unsigned long long* out /* dev [q][len] */, void* stream);
/* The same exchange with fewer launches and no host-side glue (what sharded.evaluate_sharded uses):
 *   dvl_shard_stats      this rank's statistics record after dvl_shard_topk: [tau f32 x q][flags u32 x q]
 *                        [label counts u64 x 17] = 2q + 34 32-bit words, written into the send buffer;
 *      all-gather the records;
 *   dvl_shard_reduce     all records -> tau bound (max), flags (or), label counts (sum), kept in the context
 *                        (also on a rank that owns no rows);
 *   dvl_shard_trim_pack  dvl_shard_trim + dvl_shard_pack in one kernel against that bound; *out_need (device)
 *                        = the longest list prefix that can still rank, so a too-short `len` is detected;
 *      all-gather the packed lists;  dvl_merge_topk with class_counts_global = NULL (the reduced counts),
 *      writing straight into this rank's result record
 *          [per_query f64 x block x 4][ranked_len i32 x block][flags u32 x block][need i32], rec_bytes % 8 == 0;
 *      all-gather the records;
 *   dvl_shard_finish     mean over the q queries of all records (the reduction tree of dvl_mean: bit-identical
 *                        to the single-GPU path) -> out_mean4, and out_status3 = {any query flagged for the exact
 *                        re-rank, a label >= n_classes seen, longest list any rank needed}; ONE host
 *                        synchronisation for the whole exchange. */
DVL_API int dvl_shard_stats(dvl_ctx* ctx, int32_t* out_stats /* dev [2q+34] */, int32_t q, void* stream);
DVL_API int dvl_shard_reduce(dvl_ctx* ctx, const int32_t* stats_all /* dev [world][2q+34] */, int32_t world, int32_t q,
                             void* stream);
DVL_API int dvl_shard_trim_pack(dvl_ctx* ctx, int32_t q, int64_t len, unsigned long long* out /* dev [q][len] */,
                                int32_t* out_need /* dev [1] */, void* stream);
DVL_API int dvl_shard_finish(dvl_ctx* ctx, const void* records_all /* dev [world][rec_bytes] */, int32_t world,
                             int32_t block, int64_t rec_bytes, int32_t q, double* out_mean4 /* host */,
                             int64_t* out_status3 /* host */, void* stream);
DVL_API int dvl_merge_topk(dvl_ctx* ctx, const unsigned long long* gathered /* dev [world][q_total][len] */,
                           int32_t world, int32_t q_total, int64_t len, int32_t q0, int32_t qn, int64_t n_global,
                           const unsigned long long* class_counts_global /* dev [16], or NULL */, int32_t n_classes,
                           int64_t top_n, double* out_per_query /* dev [qn][4] */, int32_t* out_ranked_len /* dev [qn] */,
                           int32_t* out_ranked_idx /* dev [qn][ranked_cap] or NULL */, int64_t ranked_cap,
                           uint32_t* out_flags /* dev [qn] */, void* stream);
/* a7: mean over queries of a device [q][4] array into host out_mean4 (fixed reduction tree). */
DVL_API int dvl_mean(dvl_ctx* ctx, const double* per_query, int32_t q, double* out_mean4, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DVL_B200_H_ */
