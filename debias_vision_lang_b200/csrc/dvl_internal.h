// Host-side internals shared by the translation units of libdvl_b200.so.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/dvl_b200.h"

namespace dvl {

struct Ctx;

// Grow-only device buffer owned by the context.
struct DevBuf {
  void* ptr = nullptr;
  size_t bytes = 0;
};

// one prepared similarity problem: 16-bit operand planes (+ row factors) of both sides
struct Operands {
  const void* qry_planes;
  const void* gal_planes;
  int planes;
  int dp;
  const float* qry_scale;
  const float* gal_scale;
};

struct Ctx {
  int device = 0;
  int num_sms = 0;
  std::string err;
  uint64_t launches = 0;
  size_t ws_bytes = 0;
  // named workspace buffers
  DevBuf gal_planes, qry_planes, gal_inv_norm, qry_inv_norm;
  DevBuf keys[2], labs[2], idxs[2];
  DevBuf nan_cnt;
  DevBuf tile_hist, class_counts, log_tab, per_query, mean_dev, scores_tmp, seg_len, sub_cnt, partial;
  DevBuf h_gallery, h_queries, h_labels;   // staging for the host-input entry point
  // streaming top-k state
  DevBuf cand, cand_cnt, cand_tau, cand_flags, topk_labs, topk_avail, fb_queries, fb_per_query, fb_len, fb_idx;
  DevBuf merge_sorted, merge_labs, merge_avail, tab_chunks;
  // bucket sort (bucketsort.cu): bucket storage, per-bucket counts, cell -> bucket tables, per-query plans,
  // and the device word a bucket overflow raises
  DevBuf bk_store, bk_cnt, bk_cells, bk_cls, bk_lut, bk_plan, sort_flag, label_info;
  DevBuf shard_tau_bound, shard_flags_any, shard_counts, shard_out;   // reduced over the ranks (dvl_shard_reduce)
  int shard_reduced_q = 0;
  bool bucket_attr_set = false;
  bool bucket_used = false;                // a ranking of the current call went through the bucket sort
  bool force_lsd = false;                  // retry after an overflow: radix sort only
  // the gallery operand the last dvl_shard_topk prepared (DVL_FLAG_REUSE_GALLERY)
  Operands prepared_gallery = {};
  int64_t prepared_n = 0;
  int prepared_d = 0, prepared_precision = 0;
  uint32_t prepared_flags = 0;
  bool prepared_valid = false;
  bool radix_only = false;                 // DVL_FLAG_RADIX_SORT of the current call
  bool graph_bucket = false;               // the captured graph contains bucket-sort kernels
  bool select_attr_set = false;
  int64_t shard_plan_cap = 0;
  int shard_sort_cap = 0;
  int shard_q = 0;
  double* mean_pinned = nullptr;           // 4 doubles, pinned host
  int64_t log_tab_len = 0;
  // driver entry point for tensor-map encoding (resolved once, no -lcuda link dependency)
  void* encode_tiled = nullptr;
  bool gemm_attr_set = false;
  // CUDA-graph replay of the full-sort route (api.cu run_full_sort)
  uint64_t ws_generation = 0;              // bumped by every workspace (re)allocation
  std::vector<uint8_t> warm_key, graph_key;
  uint64_t warm_generation = 0, graph_generation = 0, graph_launches = 0;
  cudaGraphExec_t graph_exec = nullptr;
  cudaStream_t own_stream = nullptr;
  cudaEvent_t fence_event = nullptr;
  // host-input entry points: row blocks of the gallery travel on a copy stream, one event per block
  cudaStream_t copy_stream = nullptr;
  std::vector<cudaEvent_t> upload_events;
  // per-stage CUDA-event timing (dvl_set_profiling / dvl_get_profile)
  bool profiling = false;
  struct ProfRec { int stage; cudaEvent_t e0, e1; uint64_t launches; };
  std::vector<ProfRec> prof_recs;
  std::vector<cudaEvent_t> prof_free;
  double stage_ms[DVL_NUM_STAGES] = {0};
  uint64_t stage_launches[DVL_NUM_STAGES] = {0};
};

// Brackets the kernels of one stage with events on the launching stream when profiling is on.
struct StageScope {
  Ctx* c;
  int stage;
  cudaStream_t st;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  uint64_t launches0;
  StageScope(Ctx* ctx, int stage_id, cudaStream_t stream);
  ~StageScope();
};

int set_error(Ctx* c, int code, const char* fmt, ...);
int ensure(Ctx* c, DevBuf& b, size_t bytes);

#define DVL_CUDA(ctx, call)                                                                         \
  do {                                                                                              \
    cudaError_t e__ = (call);                                                                       \
    if (e__ != cudaSuccess)                                                                         \
      return dvl::set_error((ctx), DVL_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                            __FILE__, __LINE__);                                                    \
  } while (0)

#define DVL_TRY(expr)            \
  do {                           \
    int rc__ = (expr);           \
    if (rc__ != DVL_OK) return rc__; \
  } while (0)

#define DVL_LAUNCH_CHECK(ctx, what)                                                                  \
  do {                                                                                               \
    (ctx)->launches++;                                                                               \
    cudaError_t e__ = cudaGetLastError();                                                            \
    if (e__ != cudaSuccess)                                                                          \
      return dvl::set_error((ctx), DVL_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e__)); \
  } while (0)

// ---- prep.cu ---------------------------------------------------------------------------
// rows [r, d] of `dtype` -> 16-bit planes [planes][r][dp] (dp = padded row pitch, zero filled):
// 1 = one bf16 plane, 3 = exact bf16 triple, 2 = two fp16 planes of the power-of-two scaled row.
// row_scale [r] (required for planes == 2 or normalize): the factor the similarity epilogue applies to
// the row: 2^-e of the scaled split, times 1/||row||_2 when `normalize`.
// plane_rows: rows of a whole plane when `rows` is a row range of a larger matrix (src, dst and row_scale then
// point at the range's first row); 0 = rows.
int launch_split_rows(Ctx* c, const void* src, int dtype, int64_t rows, int d, int dp, int planes, void* dst,
                      float* row_scale, bool normalize, cudaStream_t st, int64_t plane_rows = 0);
int launch_class_counts(Ctx* c, const uint8_t* labels, int64_t n, int n_classes, unsigned long long* counts,
                        cudaStream_t st);
// integer labels -> uint8 codes + info [kMaxClasses + 2]: per-class counts, out-of-range flag, largest label
int launch_encode_labels(Ctx* c, const void* labels, int label_dtype, int64_t n, uint8_t* codes,
                         unsigned long long* info, cudaStream_t st);

// ---- gemm.cu ---------------------------------------------------------------------------
enum { kGemmStore = 0, kGemmSeed = 1, kGemmFilter = 2, kGemmProbe = 3, kGemmProbeLd = 4 /* measurement only: no epilogue / TMEM reads only */ };
struct GemmArgs {
  int mode = kGemmStore;
  int64_t row0 = 0, row1 = 0;       // gallery rows [row0, row1) handled by this launch
  const float* tau = nullptr;
  unsigned long long* cand = nullptr;
  uint32_t* cand_cnt = nullptr;
  int64_t cap = 0;
  const uint8_t* labels = nullptr;
  int64_t index_base = 0;
  int64_t gal_row_step = 1;         // > 1: the kernel's rows are a strided sample of the gallery
  int64_t gal_rows_total = 0;       //      (physical rows, needed for the plane stride)
  const void* qry_planes;           // [planes][q][dp] 16-bit elements
  const void* gal_planes;           // [planes][n][dp]
  int planes;                       // 1 (bf16), 2 (fp32 via scaled fp16x2: 3 product terms) or 3 (bf16x3: 6 terms)
  int64_t n;
  int q;
  int d;       // logical depth
  int dp;      // row pitch (elements), multiple of 8
  const float* qry_scale;  // [q] or nullptr: row factor applied in the epilogue (2^-e of the split, 1/||t_q||)
  const float* gal_scale;  // [n] or nullptr
  // outputs (any may be null)
  uint32_t* keys;   // [q][ldk]
  int64_t ldk;
  float* scores;    // [q][lds]
  int64_t lds;
};
int launch_sim_gemm(Ctx* c, const GemmArgs& a, cudaStream_t st);

// ---- segsort.cu ------------------------------------------------------------------------
// Stable ascending LSD radix sort of q independent segments of n 32-bit keys (pitch ldk).
// Payloads: the gallery label (u8, looked up by position in the first pass) and, optionally,
// the position itself (i32).  On return *out_buf says which of the two ping-pong buffers
// holds the result.
struct SortArgs {
  uint32_t* keys[2];
  uint8_t* labs[2];
  int32_t* idxs[2];   // both null when ranked indices are not requested
  const uint8_t* gallery_labels;
  int64_t n;
  int64_t ldk;
  int q;
  // When the whole segment is ranked (top_n >= n) the metrics never look at the sorted keys: all they need
  // besides the labels is the number of NaN rows (dropped from the ranking).  Then the first pass counts
  // them into nan_cnt [q] and the last pass does not write its keys (4 of its 10 bytes per pair).
  uint32_t* nan_cnt = nullptr;
};
int launch_segmented_sort(Ctx* c, const SortArgs& a, int* out_buf, cudaStream_t st);
int launch_scores_to_keys(Ctx* c, const float* scores, int64_t lds, uint32_t* keys, int64_t ldk, int64_t n, int q,
                          cudaStream_t st);

// ---- bucketsort.cu ---------------------------------------------------------------------
// Ascending sort of q segments of n ranking keys by two levels of score-linear bucketing (see the file
// header).  Emits the ranked labels and optionally ranked indices / sorted keys, packed from position 0 of
// each output row.  *overflow is OR-ed with 1 when a bucket did not fit (tie-heavy data): the outputs are
// then garbage and the caller must rank again with launch_segmented_sort.
struct BucketSortArgs {
  const uint32_t* keys = nullptr;      // [q][ldk]
  int64_t ldk = 0, n = 0;
  int q = 0;
  const uint8_t* gallery_labels = nullptr;
  uint8_t* labs_out = nullptr;         // [q][ldl]
  int64_t ldl = 0;
  int32_t* idx_out = nullptr;          // [q][ldi] or null
  int64_t ldi = 0;
  uint32_t* keys_out = nullptr;        // [q][ldko] or null
  int64_t ldko = 0;
  uint32_t* nan_cnt = nullptr;         // [q] or null: number of NaN-keyed rows per segment
  uint32_t* overflow = nullptr;        // device word
  // Fused metrics (whole-segment ranking, top_n >= n): when `partial` is set the sort also emits, per bucket,
  // the NDKL partial sums [q][B][2] (the layout finalize reads, B = bucket_sort_buckets(n)) and the class
  // counts of its ranked rows [q][B + 1][16]; labs_out may then be null.  Needs nan_cnt, the global label
  // histogram and the DXL / PWI tables of metrics.cu.
  double* partial = nullptr;
  uint32_t* ranked_cls = nullptr;
  const unsigned long long* class_counts = nullptr;
  const double* dxl = nullptr;
  const double* pwi = nullptr;
  int64_t n_gallery = 0;
  int n_classes = 0;
};
bool bucket_sort_eligible(int64_t n);
int bucket_sort_buckets(int64_t n);
size_t bucket_sort_workspace_per_query(int64_t n);
int launch_bucket_sort(Ctx* c, const BucketSortArgs& a, cudaStream_t st);

// ---- topk.cu ---------------------------------------------------------------------------
struct SelectArgs {
  unsigned long long* cand = nullptr;
  int64_t row_stride = 0, seg_stride = 0;
  int segs = 1;
  int64_t seg_len = 0;
  uint32_t* cnt = nullptr;
  int64_t cap = 0;
  float* tau = nullptr;
  uint32_t* flags = nullptr;
  int64_t k = 0;
  int sort_cap = 0;
  int64_t trigger = 0;
  bool final_pass = false;
  bool pilot_pass = false;          // only derive tau from the k-th best entry, then empty the list
  int64_t rows_total = 0;           // final pass: rows this list was filtered from (verification of an
                                    // optimistic threshold: fewer than min(k, rows_total) entries -> flag)
  unsigned long long* sorted_out = nullptr;
  uint8_t* labs = nullptr;
  int64_t ldl = 0;
  int32_t* idx_out = nullptr;
  int64_t idx_cap = 0;
  int32_t* avail = nullptr;
  int q_offset = 0;
  int q_count = 0;
  const uint8_t* labels = nullptr;   // local lists only
  int64_t index_base = 0;
};
int launch_topk_reset(Ctx* c, uint32_t* cnt, float* tau, uint32_t* flags, int q, uint32_t cnt_value, bool reset_all,
                      cudaStream_t st);
int launch_select(Ctx* c, const SelectArgs& a, cudaStream_t st);
int launch_trim_count(Ctx* c, const unsigned long long* sorted, int64_t row_stride, const int32_t* avail,
                      const float* tau_bound, int q, int32_t* out_cnt, cudaStream_t st);
int launch_pack_lists(Ctx* c, const unsigned long long* sorted, int64_t row_stride, const int32_t* cnt, int q,
                      int64_t len, unsigned long long* out, cudaStream_t st);

// gallery-sharded exchange (topk.cu): statistics record of a rank, its reduction over the ranks, trim + pack of
// the local lists in one kernel, and the final mean / status reduction over the ranks' result records
int launch_pack_stats(Ctx* c, const float* tau, const uint32_t* flags, const unsigned long long* counts, int q,
                      uint32_t* out, cudaStream_t st);
int launch_reduce_stats(Ctx* c, const uint32_t* all, int world, int q, float* tau_bound, uint32_t* flags,
                        unsigned long long* counts, cudaStream_t st);
int launch_trim_pack(Ctx* c, const unsigned long long* sorted, int64_t row_stride, const int32_t* avail,
                     const float* tau_bound, int q, int64_t len, unsigned long long* out, int32_t* need, cudaStream_t st);
int launch_shard_finish(Ctx* c, const uint8_t* rec_all, int world, int block, int64_t rec_bytes, int q,
                        const uint32_t* local_flags, const unsigned long long* counts, double* out, cudaStream_t st);

// ---- metrics.cu ------------------------------------------------------------------------
struct MetricArgs {
  const uint32_t* keys;   // sorted keys [q][ldk] (used to find m when top_n < n), may be null if top_n >= n
  const uint8_t* labs;    // ranked labels [q][ldl]
  int64_t ldk, ldl;
  int64_t n_avail;        // ranked rows available per query (n for the full sort)
  int64_t n_gallery;      // N of the whole gallery (dem_par denominator)
  int64_t top_n;
  int q;
  int n_classes;
  const unsigned long long* class_counts;  // [n_classes], global
  const int32_t* avail;                    // [q] or null: rows available per query (streaming top-k route)
  const uint32_t* nan_cnt = nullptr;       // [q] or null: NaN rows per query when keys is null (full sort, top_n >= n)
  double* per_query;                       // [q][4]
  int32_t* ranked_len;                     // [q], always written
};
// fp64 tables XL, WI, PW, PWL of metrics.cu for indices < len
int ensure_metric_tables(Ctx* c, int64_t len, cudaStream_t st);
int launch_metrics(Ctx* c, const MetricArgs& a, cudaStream_t st);
// metrics of a ranking whose NDKL partial sums and ranked class counts were already produced per bucket by the
// bucket sort (fused): ranked lengths + finalize only.  Sizes and hands out the two buffers the sort writes.
int fused_metric_buffers(Ctx* c, int q, int buckets, double** partial, uint32_t** ranked_cls);
int launch_metrics_fused(Ctx* c, const MetricArgs& a, int buckets, cudaStream_t st);
// DXL / PWI tables (device pointers) for kernels outside metrics.cu
void metric_tables(const Ctx* c, const double** dxl, const double** pwi);
int launch_mean(Ctx* c, const double* per_query, int q, double* mean_dev, cudaStream_t st);
int launch_copy_ranked(Ctx* c, const int32_t* idx_sorted, int64_t ldi, const int32_t* ranked_len, int q,
                       int32_t* out, int64_t cap, cudaStream_t st);

}  // namespace dvl
