// C ABI of libdvl_b200.so (include/dvl_b200.h): context, workspace, and the pipelines that
// chain the kernels of prep.cu / gemm.cu / segsort.cu / metrics.cu on the caller's stream.
#include <stdarg.h>
#include <stdio.h>
#include <math.h>
#include <string.h>

#include <new>
#include <vector>

#include "dvl_common.cuh"
#include "dvl_internal.h"

namespace dvl {

int set_error(Ctx* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (c != nullptr) c->err = buf;
  return code;
}

int ensure(Ctx* c, DevBuf& b, size_t bytes) {
  if (bytes <= b.bytes) return DVL_OK;
  if (b.ptr != nullptr) {
    cudaFree(b.ptr);   // implicit device sync: nothing of ours is still reading it afterwards
    c->ws_bytes -= b.bytes;
    b.ptr = nullptr;
    b.bytes = 0;
  }
  size_t want = bytes + bytes / 16;
  want = (want + 255) & ~size_t(255);
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    want = (bytes + 255) & ~size_t(255);
    e = cudaMalloc(&p, want);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_error(c, DVL_ERR_OOM, "cudaMalloc of %zu bytes failed: %s", want, cudaGetErrorString(e));
  }
  b.ptr = p;
  b.bytes = want;
  c->ws_bytes += want;
  c->ws_generation++;          // a captured graph holds workspace addresses: it is stale now
  return DVL_OK;
}

StageScope::StageScope(Ctx* ctx, int stage_id, cudaStream_t stream) : c(ctx), stage(stage_id), st(stream) {
  launches0 = c->launches;
  if (!c->profiling) return;
  auto get = [&]() {
    cudaEvent_t e = nullptr;
    if (!c->prof_free.empty()) {
      e = c->prof_free.back();
      c->prof_free.pop_back();
    } else if (cudaEventCreate(&e) != cudaSuccess) {
      e = nullptr;
    }
    return e;
  };
  e0 = get();
  e1 = get();
  if (e0 != nullptr) cudaEventRecord(e0, st);
}

StageScope::~StageScope() {
  if (!c->profiling || e0 == nullptr || e1 == nullptr) return;
  cudaEventRecord(e1, st);
  c->prof_recs.push_back({stage, e0, e1, c->launches - launches0});
}

namespace {

inline int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

int dtype_size(int dtype) {
  switch (dtype) {
    case DVL_F32: return 4;
    case DVL_BF16: return 2;
    case DVL_F16: return 2;
    default: return 0;
  }
}

int check_common(Ctx* c, int64_t n, int32_t d, int32_t q, int32_t n_classes, int64_t top_n, int precision,
                 int gdt, int qdt) {
  if (c == nullptr) return DVL_ERR_INVALID;
  if (n < 1 || q < 1 || d < 1) return set_error(c, DVL_ERR_INVALID, "n, d and q must be >= 1 (n=%lld d=%d q=%d)", (long long)n, d, q);
  if (d > 8192) return set_error(c, DVL_ERR_UNSUPPORTED, "embedding depth %d > 8192", d);
  if (n_classes < 1 || n_classes > DVL_MAX_CLASSES)
    return set_error(c, DVL_ERR_INVALID, "n_classes %d outside 1..%d", n_classes, DVL_MAX_CLASSES);
  if (top_n < 1) return set_error(c, DVL_ERR_INVALID, "top_n must be >= 1 (the reference divides by it)");
  if (precision != DVL_PRECISION_FP32 && precision != DVL_PRECISION_BF16 && precision != DVL_PRECISION_FP32X3)
    return set_error(c, DVL_ERR_INVALID, "unknown precision %d", precision);
  if (dtype_size(gdt) == 0 || dtype_size(qdt) == 0) return set_error(c, DVL_ERR_INVALID, "unknown dtype");
  return DVL_OK;
}


// bf16 embeddings in bf16 mode are already the operand TMA wants (K-major rows, 16-byte aligned
// pitch): the caller's memory is used in place and the conversion pass is skipped
bool operand_direct(const void* ptr, int dtype, int d, int precision, bool need_norm) {
  return precision == DVL_PRECISION_BF16 && dtype == DVL_BF16 && !need_norm && d % 8 == 0 &&
         (reinterpret_cast<uintptr_t>(ptr) & 15u) == 0;
}

// 16-bit planes per operand: 1 = bf16, 2 = scaled fp16 pair (fp32 mode), 3 = exact bf16 triple
int planes_of(int precision) {
  return precision == DVL_PRECISION_BF16 ? 1 : (precision == DVL_PRECISION_FP32X3 ? 3 : 2);
}

// Gallery rows that arrive from the host while the path is already running (host-input entry points):
// the rows travel in blocks on the context's copy stream, one event per block; a consumer calls
// arrive(row1) before it touches rows below row1.  The operand split of arrived rows (when the rows are not
// used in place) happens here as well, range by range, so the similarity of a chunk starts as soon as the
// chunk is on the device.
struct RowFeed {
  Ctx* c = nullptr;
  int64_t n = 0, block_rows = 0;
  const char* rows_dev = nullptr;   // staged rows on the device
  int dtype = 0, d = 0, dp = 0, planes = 0;
  void* planes_dst = nullptr;       // null: the staged rows are the operand
  float* scale = nullptr;
  bool normalize = false;
  int64_t ready = 0;                // rows [0, ready) are on the device and prepared, in compute-stream order
  int arrive(int64_t row1, cudaStream_t st);
};

int RowFeed::arrive(int64_t row1, cudaStream_t st) {
  if (row1 > n) row1 = n;
  if (row1 <= ready) return DVL_OK;
  const int64_t block = (row1 - 1) / block_rows;
  // the copies are stream-ordered: block b complete means every block before it is
  DVL_CUDA(c, cudaStreamWaitEvent(st, c->upload_events[(size_t)block], 0));
  if (planes_dst != nullptr) {
    StageScope scope(c, DVL_STAGE_PREP, st);
    const size_t elt = (size_t)(dtype == DVL_F32 ? 4 : 2);
    DVL_TRY(launch_split_rows(c, rows_dev + (size_t)ready * d * elt, dtype, row1 - ready, d, dp, planes,
                              (char*)planes_dst + (size_t)ready * dp * 2, scale != nullptr ? scale + ready : nullptr,
                              normalize, st, n));
  }
  ready = row1;
  return DVL_OK;
}

// one side of the similarity: 16-bit planes (+ the row factors of the fused scale: 2^-e of the fp16
// split, inverse row norms) in the context's workspace, or the caller's rows in place.  With a feed the
// split is deferred to RowFeed::arrive.
int prepare_side(Ctx* c, const void* rows_ptr, int dtype, int64_t rows, int d, int precision, bool need_norm,
                 DevBuf& planes_buf, DevBuf& scale_buf, const void** out_planes, const float** out_scale,
                 cudaStream_t st, RowFeed* feed = nullptr) {
  const int planes = planes_of(precision);
  const int dp = (int)round_up(d, 8);
  const bool direct = operand_direct(rows_ptr, dtype, d, precision, need_norm);
  if (!direct) DVL_TRY(ensure(c, planes_buf, (size_t)planes * rows * dp * 2));
  float* scale = nullptr;
  if (need_norm || planes == 2) {
    DVL_TRY(ensure(c, scale_buf, (size_t)rows * 4));
    scale = (float*)scale_buf.ptr;
  }
  if (feed != nullptr) {
    feed->rows_dev = (const char*)rows_ptr;
    feed->dtype = dtype;
    feed->d = d;
    feed->dp = dp;
    feed->planes = planes;
    feed->planes_dst = direct ? nullptr : planes_buf.ptr;
    feed->scale = scale;
    feed->normalize = need_norm;
  } else if (!direct) {
    StageScope scope(c, DVL_STAGE_PREP, st);
    DVL_TRY(launch_split_rows(c, rows_ptr, dtype, rows, d, dp, planes, planes_buf.ptr, scale, need_norm, st));
  }
  *out_planes = direct ? rows_ptr : planes_buf.ptr;
  *out_scale = scale;
  return DVL_OK;
}

// a1 for a query batch against an already prepared gallery (the exact re-rank of flagged queries)
int prepare_queries(Ctx* c, const void* queries, int qdt, int d, int q, int precision, uint32_t flags, Operands* op,
                    cudaStream_t st) {
  return prepare_side(c, queries, qdt, q, d, precision, (flags & DVL_FLAG_NORMALIZE_QUERIES) != 0, c->qry_planes,
                      c->qry_inv_norm, &op->qry_planes, &op->qry_scale, st);
}

// a1 + operand preparation: bf16 planes of gallery and queries, inverse norms for the fused scale
int prepare_operands(Ctx* c, const void* gallery, int gdt, const void* queries, int qdt, int64_t n, int d, int q,
                     int precision, uint32_t flags, Operands* op, cudaStream_t st, RowFeed* feed = nullptr) {
  op->planes = planes_of(precision);
  op->dp = (int)round_up(d, 8);
  DVL_TRY(prepare_side(c, gallery, gdt, n, d, precision, (flags & DVL_FLAG_NORMALIZE_GALLERY) != 0, c->gal_planes,
                       c->gal_inv_norm, &op->gal_planes, &op->gal_scale, st, feed));
  return prepare_queries(c, queries, qdt, d, q, precision, flags, op, st);
}

GemmArgs gemm_args(const Operands& op, int64_t n, int d, int q) {
  GemmArgs g;
  g.qry_planes = op.qry_planes;
  g.gal_planes = op.gal_planes;
  g.planes = op.planes;
  g.n = n;
  g.q = q;
  g.d = d;
  g.dp = op.dp;
  g.qry_scale = op.qry_scale;
  g.gal_scale = op.gal_scale;
  g.keys = nullptr;
  g.ldk = 0;
  g.scores = nullptr;
  g.lds = 0;
  g.row0 = 0;
  g.row1 = n;
  return g;
}

// similarity of prepared operands into ranking keys and/or fp32 scores (both pitched to whole 256-column tiles)
int run_similarity_prepared(Ctx* c, const Operands& op, int64_t n, int d, int q, uint32_t* keys, int64_t ldk,
                            float* scores, int64_t lds, cudaStream_t st) {
  StageScope scope(c, DVL_STAGE_GEMM, st);
  GemmArgs g = gemm_args(op, n, d, q);
  g.keys = keys;
  g.ldk = ldk;
  g.scores = scores;
  g.lds = lds;
  return launch_sim_gemm(c, g, st);
}

// similarity of prepared operands; with a feed (rows still arriving from the host) as a few row-range
// launches, each as soon as its rows are there
int run_similarity_slices(Ctx* c, const Operands& op, int64_t n, int d, int q, uint32_t* keys, int64_t ldk,
                          float* scores, int64_t lds, cudaStream_t st, RowFeed* feed) {
  if (feed == nullptr || feed->ready >= n) return run_similarity_prepared(c, op, n, d, q, keys, ldk, scores, lds, st);
  const int64_t slice = round_up((n + 3) / 4, feed->block_rows);
  for (int64_t row0 = 0; row0 < n; row0 += slice) {
    const int64_t row1 = row0 + slice < n ? row0 + slice : n;
    DVL_TRY(feed->arrive(row1, st));
    StageScope scope(c, DVL_STAGE_GEMM, st);
    GemmArgs g = gemm_args(op, n, d, q);
    g.keys = keys;
    g.ldk = ldk;
    g.scores = scores;
    g.lds = lds;
    g.row0 = row0;
    g.row1 = row1;
    DVL_TRY(launch_sim_gemm(c, g, st));
  }
  return DVL_OK;
}

int run_similarity(Ctx* c, const void* gallery, int gdt, const void* queries, int qdt, int64_t n, int d, int q,
                   int precision, uint32_t flags, uint32_t* keys, int64_t ldk, float* scores, int64_t lds,
                   cudaStream_t st, RowFeed* feed = nullptr) {
  Operands op;
  DVL_TRY(prepare_operands(c, gallery, gdt, queries, qdt, n, d, q, precision, flags, &op, st, feed));
  return run_similarity_slices(c, op, n, d, q, keys, ldk, scores, lds, st, feed);
}

// query-independent statistics: Counter(df.label) and the fp64 tables
int run_label_stats(Ctx* c, const uint8_t* labels, int64_t n, int n_classes, int64_t table_len, cudaStream_t st) {
  DVL_TRY(ensure(c, c->class_counts, sizeof(unsigned long long) * (kMaxClasses + 1)));
  StageScope scope(c, DVL_STAGE_PREP, st);
  DVL_TRY(ensure_metric_tables(c, table_len, st));
  if (labels != nullptr)
    DVL_TRY(launch_class_counts(c, labels, n, n_classes, (unsigned long long*)c->class_counts.ptr, st));
  return DVL_OK;
}

// the device word a bucket overflow raises: zero between calls (finish_enqueue reads and clears it)
int ensure_sort_flag(Ctx* c, cudaStream_t st) {
  if (c->sort_flag.ptr != nullptr) return DVL_OK;
  DVL_TRY(ensure(c, c->sort_flag, 256));
  DVL_CUDA(c, cudaMemsetAsync(c->sort_flag.ptr, 0, 256, st));
  return DVL_OK;
}

// Opt-in (DVL_FUSE_METRICS=1), off by default -- measured on config 2 and rejected: with the NDKL partial sums
// computed inside bucket_sort_kernel (fused_metrics in bucketsort.cu) the sort takes 0.90 instead of 0.28 ms and
// the partition 0.16 instead of 0.126, against 0.12 ms saved in the metrics stage (1.10 vs 0.57 ms in total).
// The metrics are bound by their fp64 table gathers, not by the one byte per ranked row they read from HBM, and
// inside the sort CTA those gathers sit on a per-thread chain of 16 rows with half the threads idle.  Kept
// because it is exact and tested (it documents what the label round trip through HBM costs: nothing).
bool fuse_metrics_enabled() {
  static const bool on = [] {
    const char* e = getenv("DVL_FUSE_METRICS");
    return e != nullptr && e[0] == '1';
  }();
  return on;
}

constexpr int kRetryLsd = 1;      // internal: a bucket of the bucket sort overflowed, rank again with the radix sort

// Full-sort route.  keys [q][ldk] already in c->keys[0] -> stable segmented sort -> metrics.
// Writes per_query [q][4] and ranked_len [q] (device); ranked indices on request.
int rank_full(Ctx* c, const uint8_t* labels, int64_t n, int64_t n_gallery, int q, int n_classes, int64_t top_n,
              int64_t ldk, double* per_query, int32_t* ranked_len, int32_t* out_ranked_idx, int64_t ranked_cap,
              cudaStream_t st) {
  const bool want_idx = out_ranked_idx != nullptr;
  if (want_idx && ranked_cap < 1) return set_error(c, DVL_ERR_INVALID, "ranked_cap must be >= 1");
  // the whole segment is ranked: the metrics need the NaN count, not the sorted keys
  const bool whole = top_n >= n;
  const bool bucket = bucket_sort_eligible(n) && !c->force_lsd && !c->radix_only;
  if (!whole || !bucket) DVL_TRY(ensure(c, c->keys[1], (size_t)q * ldk * 4));
  DVL_TRY(ensure(c, c->labs[0], (size_t)q * ldk));
  if (!bucket) DVL_TRY(ensure(c, c->labs[1], (size_t)q * ldk));
  if (want_idx) {
    DVL_TRY(ensure(c, c->idxs[0], (size_t)q * ldk * 4));
    if (!bucket) DVL_TRY(ensure(c, c->idxs[1], (size_t)q * ldk * 4));
  }
  if (whole) DVL_TRY(ensure(c, c->nan_cnt, (size_t)q * sizeof(uint32_t)));
  const uint32_t* sorted_keys = nullptr;
  const uint8_t* ranked_labs = nullptr;
  const int32_t* ranked_idx = nullptr;
  bool fused = false;
  if (bucket) {
    // two-level score-linear bucket sort (bucketsort.cu); a bucket overflow (tie-heavy data) raises
    // c->sort_flag, which finish_wait turns into one retry of the call with the radix sort
    DVL_TRY(ensure_sort_flag(c, st));
    BucketSortArgs b;
    b.keys = (const uint32_t*)c->keys[0].ptr;
    b.ldk = ldk;
    b.n = n;
    b.q = q;
    b.gallery_labels = labels;
    b.ldl = ldk;
    b.idx_out = want_idx ? (int32_t*)c->idxs[0].ptr : nullptr;
    b.ldi = ldk;
    b.keys_out = whole ? nullptr : (uint32_t*)c->keys[1].ptr;
    b.ldko = ldk;
    b.nan_cnt = whole ? (uint32_t*)c->nan_cnt.ptr : nullptr;
    b.overflow = (uint32_t*)c->sort_flag.ptr;
    // opt-in variant (see fuse_metrics_enabled): the NDKL partial sums come out of the sort itself, per bucket
    fused = whole && fuse_metrics_enabled();
    if (fused) {
      DVL_TRY(fused_metric_buffers(c, q, bucket_sort_buckets(n), &b.partial, &b.ranked_cls));
      metric_tables(c, &b.dxl, &b.pwi);
      b.class_counts = (const unsigned long long*)c->class_counts.ptr;
      b.n_gallery = n_gallery;
      b.n_classes = n_classes;
    } else {
      b.labs_out = (uint8_t*)c->labs[0].ptr;
    }
    DVL_TRY(launch_bucket_sort(c, b, st));
    c->bucket_used = true;
    sorted_keys = b.keys_out;
    ranked_labs = b.labs_out;
    ranked_idx = b.idx_out;
  } else {
  SortArgs s;
  s.keys[0] = (uint32_t*)c->keys[0].ptr;
  s.keys[1] = (uint32_t*)c->keys[1].ptr;
  s.labs[0] = (uint8_t*)c->labs[0].ptr;
  s.labs[1] = (uint8_t*)c->labs[1].ptr;
  s.idxs[0] = want_idx ? (int32_t*)c->idxs[0].ptr : nullptr;
  s.idxs[1] = want_idx ? (int32_t*)c->idxs[1].ptr : nullptr;
  s.gallery_labels = labels;
  s.n = n;
  s.ldk = ldk;
  s.q = q;
  if (whole) s.nan_cnt = (uint32_t*)c->nan_cnt.ptr;
  int buf = 0;
  DVL_TRY(launch_segmented_sort(c, s, &buf, st));
  sorted_keys = whole ? nullptr : s.keys[buf];
  ranked_labs = s.labs[buf];
  ranked_idx = s.idxs[buf];
  }
  MetricArgs m;
  m.keys = sorted_keys;
  m.nan_cnt = whole ? (const uint32_t*)c->nan_cnt.ptr : nullptr;
  m.labs = ranked_labs;
  m.ldk = ldk;
  m.ldl = ldk;
  m.n_avail = n;
  m.n_gallery = n_gallery;
  m.top_n = top_n;
  m.q = q;
  m.n_classes = n_classes;
  m.class_counts = (const unsigned long long*)c->class_counts.ptr;
  m.avail = nullptr;
  m.per_query = per_query;
  m.ranked_len = ranked_len;
  {
    StageScope scope(c, DVL_STAGE_METRICS, st);
    if (fused) DVL_TRY(launch_metrics_fused(c, m, bucket_sort_buckets(n), st));
    else DVL_TRY(launch_metrics(c, m, st));
  }
  if (want_idx) {
    StageScope scope(c, DVL_STAGE_COPY, st);
    DVL_TRY(launch_copy_ranked(c, ranked_idx, ldk, ranked_len, q, out_ranked_idx, ranked_cap, st));
  }
  return DVL_OK;
}

// mean over queries -> pinned host memory, plus the label sanity flag (stream-ordered, no host wait)
int finish_enqueue(Ctx* c, const double* per_query, int q, bool check_labels, cudaStream_t st) {
  DVL_TRY(ensure(c, c->mean_dev, 8 * sizeof(double)));
  {
    StageScope scope(c, DVL_STAGE_METRICS, st);
    DVL_TRY(launch_mean(c, per_query, q, (double*)c->mean_dev.ptr, st));
  }
  DVL_CUDA(c, cudaMemcpyAsync(c->mean_pinned, c->mean_dev.ptr, 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (check_labels)   // the out-of-range label flag travels with the class counts
    DVL_CUDA(c, cudaMemcpyAsync(c->mean_pinned + 4, (unsigned long long*)c->class_counts.ptr + kMaxClasses,
                                sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  if (c->bucket_used) {   // ... and so does the bucket-overflow word, which is cleared for the next call
    DVL_CUDA(c, cudaMemcpyAsync(c->mean_pinned + 5, c->sort_flag.ptr, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    DVL_CUDA(c, cudaMemsetAsync(c->sort_flag.ptr, 0, sizeof(uint32_t), st));
  }
  return DVL_OK;
}

// waits for the stream and hands the four means to the caller
// Returns kRetryLsd (> 0) when a bucket sort of this call overflowed: the results are not valid and the
// caller repeats the ranking with c->force_lsd set.
int finish_wait(Ctx* c, int n_classes, double* out_mean4, bool check_labels, cudaStream_t st) {
  DVL_CUDA(c, cudaStreamSynchronize(st));
  if (c->bucket_used) {
    c->bucket_used = false;
    uint32_t overflow = 0;
    memcpy(&overflow, c->mean_pinned + 5, sizeof(overflow));
    if (overflow != 0u) return kRetryLsd;
  }
  if (check_labels) {
    unsigned long long bad_flag = 0;
    memcpy(&bad_flag, c->mean_pinned + 4, sizeof(bad_flag));
    if (bad_flag != 0) return set_error(c, DVL_ERR_INVALID, "a label is >= n_classes (%d)", n_classes);
  }
  if (out_mean4 != nullptr) memcpy(out_mean4, c->mean_pinned, 4 * sizeof(double));
  return DVL_OK;
}

int finish(Ctx* c, const double* per_query, int q, int n_classes, double* out_mean4, bool check_labels,
           cudaStream_t st) {
  DVL_TRY(finish_enqueue(c, per_query, q, check_labels, st));
  return finish_wait(c, n_classes, out_mean4, check_labels, st);
}

// ---------------------------------------------------------------------------------------------
// streaming top-k route
// ---------------------------------------------------------------------------------------------
struct StreamPlan {
  int sort_cap;        // S: entries a query keeps between chunks (top_n + boundary ties)
  int64_t chunk0, chunk_max, cap, trigger;
  int64_t pilot_step;  // > 1: optimistic threshold from every pilot_step-th row (large galleries)
  int64_t pilot_rank;  // rank of the pilot sample's entry that becomes the threshold
};

// Pilot sample: ~16k rows cost <= 6 % of the main pass on galleries of >= 16x as many rows (25 % at 4x) and leave
// ~2.8 top_n survivors per query at top_n = 1000 of 1.25 M rows; 65536 rows left ~1.7 top_n but cost
// 5 % of the similarity work on config 4.
constexpr int64_t kPilotRows = 16384;
constexpr int64_t kPilotRowsMin = 4096;     // a sample this small is scored densely: no seed / filter split

bool plan_stream(int64_t n, int q, int64_t top_n, uint32_t flags, StreamPlan* plan, bool shard = false) {
  if (top_n > 8192) return false;
  if (top_n >= n && !shard) return false;        // the whole gallery is ranked: full-sort route
  int s = 2048;
  while (s < 2 * top_n) s <<= 1;
  plan->sort_cap = s;
  plan->trigger = 2 * top_n;
  int64_t chunk0 = 8192;
  if (shard || (flags & DVL_FLAG_STREAM_TOPK)) { // forced: keep several chunks even on a small gallery
    chunk0 = round_up((n + 3) / 4, 256);
    if (chunk0 > 8192) chunk0 = 8192;
  } else if (n < 4 * chunk0 || (flags & DVL_FLAG_FULL_SORT)) {
    return false;
  }
  if (chunk0 < plan->trigger) chunk0 = round_up(plan->trigger, 256);
  int64_t chunk_max = 65536;
  while (chunk_max > chunk0 && (double)q * (double)(s + chunk_max) * 8.0 > 6.0e9) chunk_max >>= 1;
  if (chunk_max < chunk0) chunk_max = chunk0;
  plan->chunk0 = chunk0;
  plan->chunk_max = chunk_max;
  plan->cap = s + chunk_max;
  plan->pilot_step = 1;
  plan->pilot_rank = 0;
  // From 65,536 rows on (every 4th row is sampled there, every n/16384-th from 262,144 rows on): below that the
  // progressive scheme's few chunks are cheaper than a pilot pass.  The mid range matters for shards: 125 k rows
  // per GPU (config 3 on 8 GPUs) took five GEMM launches and five selects progressively, three of each with a pilot.
  if (!(flags & DVL_FLAG_NO_PILOT) && n >= 4 * kPilotRows && top_n < n / 8) {
    // sample size: 1/32 of the gallery between kPilotRowsMin and kPilotRows.  A 125 k-row shard (config 3 on 8
    // GPUs) spent a third of its local time on a 16 k-row pilot (seed GEMM, select, filter GEMM, select: 121 us
    // next to the 192 us main pass); 4 k rows are one dense seed launch + one select and leave ~2.0 top_n
    // survivors instead of ~1.4.
    int64_t sample = n / 32;
    if (sample < kPilotRowsMin) sample = kPilotRowsMin;
    if (sample > kPilotRows) sample = kPilotRows;
    const int64_t step = (n + sample - 1) / sample;
    const int64_t n_view = (n + step - 1) / step;
    const double mu = (double)top_n * (double)n_view / (double)n;     // expected pilot rank of the top_n-th best
    const int64_t rank = (int64_t)ceil(mu + 4.0 * sqrt(mu + 4.0) + 8.0);
    if (rank + 1 < n_view && rank <= s / 2) {
      plan->pilot_step = step;
      plan->pilot_rank = rank;
    }
  }
  return true;
}

// Streams the local gallery through the similarity GEMM with the threshold filter and leaves, per
// query, the sorted list of every row that ranks within top_n (boundary ties included):
// c->cand [q][cap] (first avail[q] entries), labels in c->topk_labs [q][ldl], lengths in c->topk_avail.
int stream_topk(Ctx* c, const Operands& op, const uint8_t* labels, int64_t n, int64_t index_base, int d, int q,
                int64_t top_n, const StreamPlan& plan, int32_t* idx_out, int64_t idx_cap, cudaStream_t st,
                RowFeed* feed = nullptr) {
  const int64_t ldl = round_up(plan.sort_cap, 256);
  DVL_TRY(ensure(c, c->cand, (size_t)q * plan.cap * 8));
  DVL_TRY(ensure(c, c->cand_cnt, (size_t)q * 4));
  DVL_TRY(ensure(c, c->cand_tau, (size_t)q * 4));
  DVL_TRY(ensure(c, c->cand_flags, (size_t)q * 4));
  DVL_TRY(ensure(c, c->topk_labs, (size_t)q * ldl));
  DVL_TRY(ensure(c, c->topk_avail, (size_t)q * 4));
  unsigned long long* cand = (unsigned long long*)c->cand.ptr;
  uint32_t* cnt = (uint32_t*)c->cand_cnt.ptr;
  float* tau = (float*)c->cand_tau.ptr;
  uint32_t* cflags = (uint32_t*)c->cand_flags.ptr;

  SelectArgs sel;
  sel.cand = cand;
  sel.row_stride = plan.cap;
  sel.segs = 1;
  sel.cnt = cnt;
  sel.cap = plan.cap;
  sel.tau = tau;
  sel.flags = cflags;
  sel.k = top_n;
  sel.sort_cap = plan.sort_cap;
  sel.trigger = plan.trigger;
  sel.labs = (uint8_t*)c->topk_labs.ptr;
  sel.ldl = ldl;
  sel.avail = (int32_t*)c->topk_avail.ptr;
  sel.q_count = q;
  sel.labels = labels;
  sel.index_base = index_base;

  auto gemm_launch = [&](int mode, int64_t row0, int64_t row1, int64_t n_view, int64_t row_step) -> int {
    if (feed != nullptr) DVL_TRY(feed->arrive(row_step > 1 ? n : row1, st));
    StageScope scope(c, DVL_STAGE_GEMM, st);
    GemmArgs g = gemm_args(op, n_view, d, q);
    g.mode = mode;
    g.row0 = row0;
    g.row1 = row1;
    g.tau = tau;
    g.cand = cand;
    g.cand_cnt = cnt;
    g.cap = plan.cap;
    g.labels = labels;
    g.index_base = index_base;
    g.gal_row_step = row_step;
    g.gal_rows_total = n;
    return launch_sim_gemm(c, g, st);
  };
  auto select_launch = [&](int64_t k, int64_t trigger, bool final_pass, bool pilot_pass) -> int {
    StageScope scope(c, DVL_STAGE_SELECT, st);
    sel.k = k;
    sel.trigger = trigger;
    sel.final_pass = final_pass;
    sel.pilot_pass = pilot_pass;
    sel.rows_total = final_pass ? n : 0;
    sel.idx_out = final_pass ? idx_out : nullptr;
    sel.idx_cap = final_pass ? idx_cap : 0;
    return launch_select(c, sel, st);
  };

  if (plan.pilot_step > 1) {
    // ---- optimistic threshold from a strided pilot sample (every pilot_step-th row, read in place
    // through a strided TMA map).  The sample's r-th best score, r chosen 4 sigma above the rank the
    // top_n-th best is expected at, is a threshold that keeps >= top_n rows with overwhelming
    // probability; the final select VERIFIES it (fewer than top_n survivors -> the query is flagged and
    // re-ranked exactly), so the result is exact either way.  The whole gallery then streams through ONE
    // filter launch with ~2 top_n survivors per query instead of the ~6 top_n of the progressive scheme.
    const int64_t n_view = (n + plan.pilot_step - 1) / plan.pilot_step;
    // The seed chunk is written out densely (8 bytes per query and row) only to give the rest of the
    // pilot sample a first threshold: a few multiples of the pilot rank are enough (8192 rows cost
    // 0.35 ms of candidate writes + 0.29 ms of selection at Q = 10^4; 512 rows a tenth of that).
    int64_t seed_rows = round_up(4 * plan.pilot_rank > 512 ? 4 * plan.pilot_rank : 512, 256);
    if (seed_rows > plan.chunk0) seed_rows = plan.chunk0;
    if (seed_rows > n_view || n_view <= kPilotRowsMin) seed_rows = n_view;
    DVL_TRY(launch_topk_reset(c, cnt, tau, cflags, q, (uint32_t)seed_rows, true, st));
    DVL_TRY(gemm_launch(kGemmSeed, 0, seed_rows, n_view, plan.pilot_step));
    if (seed_rows < n_view) {
      DVL_TRY(select_launch(plan.pilot_rank, plan.pilot_rank + 1, false, false));
      DVL_TRY(gemm_launch(kGemmFilter, seed_rows, n_view, n_view, plan.pilot_step));
    }
    DVL_TRY(select_launch(plan.pilot_rank, 0, false, true));
    DVL_TRY(gemm_launch(kGemmFilter, 0, n, n, 1));
    return select_launch(top_n, 0, true, false);
  }

  // ---- progressive scheme: chunks of doubling size, the threshold tightens after every chunk ----
  int64_t row = 0;
  int64_t chunk = plan.chunk0;
  bool first = true;
  while (row < n) {
    const int64_t row1 = (row + chunk < n) ? row + chunk : n;
    if (first) DVL_TRY(launch_topk_reset(c, cnt, tau, cflags, q, (uint32_t)(row1 - row), true, st));
    DVL_TRY(gemm_launch(first ? kGemmSeed : kGemmFilter, row, row1, n, 1));
    row = row1;
    first = false;
    if (row < n) {
      DVL_TRY(select_launch(top_n, plan.trigger, false, false));
      chunk = (2 * chunk < plan.chunk_max) ? 2 * chunk : plan.chunk_max;
    }
  }
  return select_launch(top_n, 0, true, false);
}

// ---------------------------------------------------------------------------------------------
// full-sort route of dvl_eval_ranking as one stream-ordered sequence, and its CUDA-graph replay
// ---------------------------------------------------------------------------------------------
struct FullSortCall {
  const void* gallery; const void* queries; const uint8_t* labels;
  int gdt, qdt; int64_t n; int d, q, n_classes; int64_t top_n; int precision; uint32_t flags;
  double* per_query; int32_t* ranked_len; int32_t* ranked_idx; int64_t ranked_cap; float* out_scores;
};

// label statistics -> similarity -> ranking keys -> segmented sort -> metrics -> mean in pinned memory
int enqueue_full_sort(Ctx* c, const FullSortCall& a, cudaStream_t st, RowFeed* feed = nullptr) {
  DVL_TRY(run_label_stats(c, a.labels, a.n, a.n_classes, (a.n > a.top_n ? a.n : a.top_n) + 2, st));
  const int64_t ldk = round_up(a.n, 256);
  // The route materialises keys (and sort workspace) per query: queries are processed in batches bounded by
  // a workspace budget (DVL_WS_BUDGET_MB, default 8 GB), so Q x N does not have to fit at once.
  c->radix_only = (a.flags & DVL_FLAG_RADIX_SORT) != 0;
  const bool bucket = bucket_sort_eligible(a.n) && !c->force_lsd && !c->radix_only;
  size_t per_query_bytes = (size_t)ldk * 4 + (bucket ? bucket_sort_workspace_per_query(a.n) + (size_t)ldk * 5
                                                     : (size_t)ldk * 6 + (size_t)((a.n + 4095) / 4096) * 1024);
  if (a.ranked_idx != nullptr) per_query_bytes += (size_t)ldk * 8;
  if (a.out_scores != nullptr) per_query_bytes += (size_t)ldk * 4;
  static const size_t budget = [] {
    const char* e = getenv("DVL_WS_BUDGET_MB");
    const long long mb = e != nullptr ? atoll(e) : 0;
    return (size_t)(mb > 0 ? mb : 8192) << 20;
  }();
  int qb = (int)(budget / per_query_bytes < 1 ? 1 : (budget / per_query_bytes > (size_t)a.q ? (size_t)a.q : budget / per_query_bytes));
  Operands op;
  op.planes = planes_of(a.precision);
  op.dp = (int)round_up(a.d, 8);
  DVL_TRY(prepare_side(c, a.gallery, a.gdt, a.n, a.d, a.precision, (a.flags & DVL_FLAG_NORMALIZE_GALLERY) != 0,
                       c->gal_planes, c->gal_inv_norm, &op.gal_planes, &op.gal_scale, st, feed));
  const size_t qrow_bytes = (size_t)a.d * dtype_size(a.qdt);
  for (int q0 = 0; q0 < a.q; q0 += qb) {
    const int qn = a.q - q0 < qb ? a.q - q0 : qb;
    DVL_TRY(ensure(c, c->keys[0], (size_t)qn * ldk * 4));
    float* scores_tmp = nullptr;
    if (a.out_scores != nullptr) {
      DVL_TRY(ensure(c, c->scores_tmp, (size_t)qn * ldk * 4));
      scores_tmp = (float*)c->scores_tmp.ptr;
    }
    DVL_TRY(prepare_queries(c, (const char*)a.queries + (size_t)q0 * qrow_bytes, a.qdt, a.d, qn, a.precision, a.flags, &op,
                            st));
    DVL_TRY(run_similarity_slices(c, op, a.n, a.d, qn, (uint32_t*)c->keys[0].ptr, ldk, scores_tmp, ldk, st, feed));
    if (a.out_scores != nullptr)
      DVL_CUDA(c, cudaMemcpy2DAsync(a.out_scores + (size_t)q0 * a.n, (size_t)a.n * 4, scores_tmp, (size_t)ldk * 4,
                                    (size_t)a.n * 4, qn, cudaMemcpyDeviceToDevice, st));
    DVL_TRY(rank_full(c, a.labels, a.n, a.n, qn, a.n_classes, a.top_n, ldk, a.per_query + (size_t)q0 * 4,
                      a.ranked_len + q0, a.ranked_idx != nullptr ? a.ranked_idx + (size_t)q0 * a.ranked_cap : nullptr,
                      a.ranked_cap, st));
  }
  c->radix_only = false;
  return finish_enqueue(c, a.per_query, a.q, true, st);
}

bool graphs_enabled() {
  static const bool on = [] {
    const char* e = getenv("DVL_GRAPH");
    return !(e != nullptr && e[0] == '0');
  }();
  return on;
}

void drop_graph(Ctx* c) {
  if (c->graph_exec != nullptr) cudaGraphExecDestroy(c->graph_exec);
  c->graph_exec = nullptr;
  c->graph_key.clear();
}

// The route is ~20 short kernels: on the small galleries it serves (measure_bias on FairFace) the
// gaps between launches are a fifth of the step.  The second identical call (same buffers, shapes and
// options, workspace already sized) is captured into a CUDA graph on the context's own stream; later
// identical calls replay it.  Any workspace reallocation invalidates the graph (ws_generation).
int run_full_sort_once(Ctx* c, const FullSortCall& call, double* out_mean4, cudaStream_t st, RowFeed* feed) {
  // per-stage profiling brackets every stage with events: measured slower inside a graph (external
  // event-record nodes) than as direct launches, so profiled calls are launched directly; so are calls
  // whose gallery is still arriving from the host (they wait on the copy stream's events block by block)
  if (!graphs_enabled() || c->profiling || feed != nullptr) {
    DVL_TRY(enqueue_full_sort(c, call, st, feed));
    return finish_wait(c, call.n_classes, out_mean4, true, st);
  }
  const uint8_t* kb = reinterpret_cast<const uint8_t*>(&call);
  const std::vector<uint8_t> key(kb, kb + sizeof(call));
  if (c->graph_exec != nullptr && (c->graph_generation != c->ws_generation || c->graph_key != key)) drop_graph(c);
  if (c->graph_exec == nullptr) {
    if (c->warm_key != key || c->warm_generation != c->ws_generation) {
      // first sight of this call: run it directly (this also sizes the workspace)
      DVL_TRY(enqueue_full_sort(c, call, st));
      c->warm_key = key;
      c->warm_generation = c->ws_generation;     // after the run: growth during it is accounted for
      return finish_wait(c, call.n_classes, out_mean4, true, st);
    }
    if (c->own_stream == nullptr) {
      DVL_CUDA(c, cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
      DVL_CUDA(c, cudaEventCreateWithFlags(&c->fence_event, cudaEventDisableTiming));
    }
    const uint64_t launches0 = c->launches;
    DVL_CUDA(c, cudaStreamBeginCapture(c->own_stream, cudaStreamCaptureModeThreadLocal));
    const int rc = enqueue_full_sort(c, call, c->own_stream);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(c->own_stream, &graph);
    if (rc != DVL_OK || ce != cudaSuccess || graph == nullptr || c->ws_generation != c->warm_generation) {
      // not capturable after all (e.g. the workspace still grew): fall back to direct launches for good
      if (graph != nullptr) cudaGraphDestroy(graph);
      cudaGetLastError();
      drop_graph(c);
      c->launches = launches0;
      c->warm_key.clear();
      DVL_TRY(enqueue_full_sort(c, call, st));
      return finish_wait(c, call.n_classes, out_mean4, true, st);
    }
    const cudaError_t ie = cudaGraphInstantiate(&c->graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) {
      cudaGetLastError();
      c->graph_exec = nullptr;
      drop_graph(c);
      c->launches = launches0;
      DVL_TRY(enqueue_full_sort(c, call, st));
      return finish_wait(c, call.n_classes, out_mean4, true, st);
    }
    c->graph_launches = c->launches - launches0;
    c->launches = launches0;
    c->graph_bucket = c->bucket_used;          // the captured sequence ranks with the bucket sort (overflow word)
    c->graph_key = key;
    c->graph_generation = c->ws_generation;
  }
  // replay on the context's stream, fenced against the caller's stream on both sides
  DVL_CUDA(c, cudaEventRecord(c->fence_event, st));
  DVL_CUDA(c, cudaStreamWaitEvent(c->own_stream, c->fence_event, 0));
  DVL_CUDA(c, cudaGraphLaunch(c->graph_exec, c->own_stream));
  c->launches += c->graph_launches;
  c->bucket_used = c->graph_bucket;
  DVL_TRY(finish_wait(c, call.n_classes, out_mean4, true, c->own_stream));
  return DVL_OK;   // own_stream is idle again: work the caller enqueues next on `st` is ordered after this return
}

// Host rows -> device staging buffer in row blocks on the context's copy stream, one event per block
// (RowFeed::arrive waits on them).  `small` copies (queries, labels) go first on the same stream; the
// compute stream `st` is made to wait for them here, and the copy stream for everything enqueued on `st`
// so far (an earlier call may still be reading the staging buffers).
struct SmallCopy { void* dst; const void* src; size_t bytes; };

int start_upload(Ctx* c, const void* host_rows, void* dev_rows, int64_t n, size_t row_bytes, const SmallCopy* small,
                 int n_small, RowFeed* feed, cudaStream_t st) {
  if (c->copy_stream == nullptr) DVL_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  // blocks of >= 16 MB (PCIe efficiency; DVL_UPLOAD_BLOCK_MB overrides) and at most ~512 of them; a multiple
  // of 256 rows (GEMM tiles)
  static const size_t block_bytes = [] {
    const char* e = getenv("DVL_UPLOAD_BLOCK_MB");
    const long long mb = e != nullptr ? atoll(e) : 0;
    return (size_t)(mb > 0 ? mb : 16) << 20;
  }();
  int64_t block_rows = round_up((int64_t)(block_bytes / row_bytes) + 1, 256);
  if (block_rows * 512 < n) block_rows = round_up((n + 511) / 512, 256);
  const int64_t blocks = (n + block_rows - 1) / block_rows;
  while ((int64_t)c->upload_events.size() < blocks + 2) {
    cudaEvent_t e = nullptr;
    DVL_CUDA(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->upload_events.push_back(e);
  }
  cudaEvent_t fence = c->upload_events[(size_t)blocks], small_done = c->upload_events[(size_t)blocks + 1];
  DVL_CUDA(c, cudaEventRecord(fence, st));
  DVL_CUDA(c, cudaStreamWaitEvent(c->copy_stream, fence, 0));
  {
    StageScope scope(c, DVL_STAGE_COPY, c->copy_stream);
    for (int i = 0; i < n_small; ++i)
      DVL_CUDA(c, cudaMemcpyAsync(small[i].dst, small[i].src, small[i].bytes, cudaMemcpyHostToDevice, c->copy_stream));
    DVL_CUDA(c, cudaEventRecord(small_done, c->copy_stream));
    for (int64_t b = 0; b < blocks; ++b) {
      const int64_t r0 = b * block_rows, r1 = (r0 + block_rows < n) ? r0 + block_rows : n;
      DVL_CUDA(c, cudaMemcpyAsync((char*)dev_rows + (size_t)r0 * row_bytes, (const char*)host_rows + (size_t)r0 * row_bytes,
                                  (size_t)(r1 - r0) * row_bytes, cudaMemcpyHostToDevice, c->copy_stream));
      DVL_CUDA(c, cudaEventRecord(c->upload_events[(size_t)b], c->copy_stream));
    }
  }
  DVL_CUDA(c, cudaStreamWaitEvent(st, small_done, 0));
  feed->c = c;
  feed->n = n;
  feed->block_rows = block_rows;
  feed->ready = 0;
  return DVL_OK;
}

int run_full_sort(Ctx* c, const FullSortCall& call, double* out_mean4, cudaStream_t st, RowFeed* feed = nullptr) {
  int rc = run_full_sort_once(c, call, out_mean4, st, feed);
  if (rc != kRetryLsd) return rc;
  // a bucket overflowed (tie-heavy scores): the same call again with the radix sort, launched directly
  // (the gallery has arrived completely by now, so no feed)
  c->force_lsd = true;
  rc = enqueue_full_sort(c, call, st, nullptr);
  if (rc == DVL_OK) rc = finish_wait(c, call.n_classes, out_mean4, true, st);
  c->force_lsd = false;
  return rc;
}

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

}  // namespace
}  // namespace dvl

using dvl::Ctx;

struct dvl_ctx : public dvl::Ctx {};

extern "C" {

DVL_API int dvl_abi_version(void) { return DVL_B200_ABI_VERSION; }

DVL_API int dvl_create(dvl_ctx** out, int device) {
  if (out == nullptr) return DVL_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
    cudaGetLastError();
    return DVL_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= count) return DVL_ERR_INVALID;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return DVL_ERR_CUDA;
  if (prop.major != 10) return DVL_ERR_NO_DEVICE;   // sm_100a code only: no fallback
  dvl_ctx* c = new (std::nothrow) dvl_ctx();
  if (c == nullptr) return DVL_ERR_OOM;
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  dvl::DeviceGuard guard(device);
  if (cudaMallocHost((void**)&c->mean_pinned, 8 * sizeof(double)) != cudaSuccess) {
    delete c;
    return DVL_ERR_OOM;
  }
  *out = c;
  return DVL_OK;
}

DVL_API int dvl_destroy(dvl_ctx* c) {
  if (c == nullptr) return DVL_OK;
  dvl::DeviceGuard guard(c->device);
  cudaDeviceSynchronize();
  dvl::DevBuf* bufs[] = {&c->nan_cnt, &c->gal_planes, &c->qry_planes, &c->gal_inv_norm, &c->qry_inv_norm, &c->keys[0], &c->keys[1],
                         &c->labs[0], &c->labs[1], &c->idxs[0], &c->idxs[1], &c->tile_hist, &c->class_counts,
                         &c->log_tab, &c->per_query, &c->mean_dev, &c->scores_tmp, &c->seg_len, &c->sub_cnt, &c->partial,
                         &c->h_gallery, &c->h_queries, &c->h_labels, &c->cand, &c->cand_cnt, &c->cand_tau,
                         &c->cand_flags, &c->topk_labs, &c->topk_avail, &c->fb_queries, &c->fb_per_query, &c->fb_len,
                         &c->fb_idx, &c->merge_sorted, &c->tab_chunks, &c->bk_store, &c->bk_cnt, &c->bk_cells, &c->bk_cls, &c->bk_lut,
                         &c->bk_plan, &c->sort_flag, &c->label_info, &c->shard_tau_bound, &c->shard_flags_any,
                         &c->shard_counts, &c->shard_out, &c->merge_labs, &c->merge_avail};
  for (dvl::DevBuf* b : bufs)
    if (b->ptr != nullptr) cudaFree(b->ptr);
  if (c->mean_pinned != nullptr) cudaFreeHost(c->mean_pinned);
  if (c->graph_exec != nullptr) cudaGraphExecDestroy(c->graph_exec);
  if (c->fence_event != nullptr) cudaEventDestroy(c->fence_event);
  if (c->own_stream != nullptr) cudaStreamDestroy(c->own_stream);
  if (c->copy_stream != nullptr) cudaStreamDestroy(c->copy_stream);
  for (cudaEvent_t e : c->upload_events) cudaEventDestroy(e);
  for (auto& r : c->prof_recs) {
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  for (cudaEvent_t e : c->prof_free) cudaEventDestroy(e);
  delete c;
  return DVL_OK;
}

DVL_API const char* dvl_last_error(const dvl_ctx* c) { return c == nullptr ? "null context" : c->err.c_str(); }
DVL_API size_t dvl_workspace_bytes(const dvl_ctx* c) { return c == nullptr ? 0 : c->ws_bytes; }
DVL_API uint64_t dvl_launch_count(const dvl_ctx* c) { return c == nullptr ? 0 : c->launches; }

DVL_API int dvl_set_profiling(dvl_ctx* c, int enabled) {
  if (c == nullptr) return DVL_ERR_INVALID;
  c->profiling = enabled != 0;
  return DVL_OK;
}

DVL_API int dvl_get_profile(dvl_ctx* c, double* stage_ms, uint64_t* stage_launches) {
  if (c == nullptr) return DVL_ERR_INVALID;
  dvl::DeviceGuard guard(c->device);
  for (auto& r : c->prof_recs) {
    float ms = 0.f;
    if (cudaEventSynchronize(r.e1) == cudaSuccess && cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
      c->stage_ms[r.stage] += ms;
      c->stage_launches[r.stage] += r.launches;
    }
    c->prof_free.push_back(r.e0);
    c->prof_free.push_back(r.e1);
  }
  c->prof_recs.clear();
  for (int i = 0; i < DVL_NUM_STAGES; ++i) {
    if (stage_ms != nullptr) stage_ms[i] = c->stage_ms[i];
    if (stage_launches != nullptr) stage_launches[i] = c->stage_launches[i];
    c->stage_ms[i] = 0.0;
    c->stage_launches[i] = 0;
  }
  return DVL_OK;
}

DVL_API int dvl_encode_labels(dvl_ctx* c, const void* labels, int label_dtype, int64_t n, uint8_t* out_codes,
                              int64_t* out_info_host, void* stream) {
  if (c == nullptr) return DVL_ERR_INVALID;
  if (labels == nullptr || out_codes == nullptr || out_info_host == nullptr || n < 1)
    return dvl::set_error(c, DVL_ERR_INVALID, "bad label arguments");
  dvl::DeviceGuard guard(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  DVL_TRY(dvl::ensure(c, c->label_info, sizeof(unsigned long long) * (dvl::kMaxClasses + 2)));
  {
    dvl::StageScope scope(c, DVL_STAGE_PREP, st);
    DVL_TRY(dvl::launch_encode_labels(c, labels, label_dtype, n, out_codes, (unsigned long long*)c->label_info.ptr, st));
  }
  unsigned long long info[dvl::kMaxClasses + 2];
  DVL_CUDA(c, cudaMemcpyAsync(info, c->label_info.ptr, sizeof(info), cudaMemcpyDeviceToHost, st));
  DVL_CUDA(c, cudaStreamSynchronize(st));
  bool contiguous = info[dvl::kMaxClasses] == 0;
  for (int k = 0; k < dvl::kMaxClasses; ++k) {
    out_info_host[k] = (int64_t)info[k];
    if (k <= (int)info[dvl::kMaxClasses + 1] && info[k] == 0) contiguous = false;
  }
  out_info_host[16] = (int64_t)info[dvl::kMaxClasses + 1] + 1;
  out_info_host[17] = info[dvl::kMaxClasses] != 0 ? 1 : 0;
  out_info_host[18] = contiguous ? 1 : 0;
  return DVL_OK;
}

DVL_API int dvl_similarity(dvl_ctx* c, const void* gallery, int gdt, const void* queries, int qdt, int64_t n, int32_t d,
                   int32_t q, int precision, uint32_t flags, float* out, int64_t ld_out, void* stream) {
  DVL_TRY(dvl::check_common(c, n, d, q, 1, 1, precision, gdt, qdt));
  if (out == nullptr || ld_out < n) return dvl::set_error(c, DVL_ERR_INVALID, "bad output matrix");
  dvl::DeviceGuard guard(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t lds = dvl::round_up(n, 256);
  DVL_TRY(dvl::ensure(c, c->scores_tmp, (size_t)q * lds * 4));
  if ((flags & DVL_FLAG_REUSE_GALLERY) != 0u) {
    // the caller vouches that `gallery` is the matrix the preceding dvl_shard_topk on this context prepared and
    // that it has not changed since: only the queries are prepared, the gallery's conversion pass is skipped
    if (!c->prepared_valid || c->prepared_n != n || c->prepared_d != d || c->prepared_precision != precision ||
        c->prepared_flags != (flags & DVL_FLAG_NORMALIZE_GALLERY))
      return dvl::set_error(c, DVL_ERR_INVALID, "DVL_FLAG_REUSE_GALLERY: no matching gallery prepared by dvl_shard_topk");
    dvl::Operands op = c->prepared_gallery;
    DVL_TRY(dvl::prepare_queries(c, queries, qdt, d, q, precision, flags, &op, st));
    DVL_TRY(dvl::run_similarity_prepared(c, op, n, d, q, nullptr, 0, (float*)c->scores_tmp.ptr, lds, st));
  } else {
    c->prepared_valid = false;        // this call prepares another gallery into the same workspace
    DVL_TRY(dvl::run_similarity(c, gallery, gdt, queries, qdt, n, d, q, precision, flags, nullptr, 0,
                                (float*)c->scores_tmp.ptr, lds, st));
  }
  DVL_CUDA(c, cudaMemcpy2DAsync(out, (size_t)ld_out * 4, c->scores_tmp.ptr, (size_t)lds * 4, (size_t)n * 4, q,
                                cudaMemcpyDeviceToDevice, st));
  DVL_CUDA(c, cudaStreamSynchronize(st));
  return DVL_OK;
}

DVL_API int dvl_rank_and_score(dvl_ctx* c, const float* scores, int64_t ld_scores, const uint8_t* labels, int64_t n,
                       int32_t q, int32_t n_classes, int64_t top_n, double* out_mean4, double* out_per_query,
                       int32_t* out_ranked_len, int32_t* out_ranked_idx, int64_t ranked_cap, void* stream) {
  DVL_TRY(dvl::check_common(c, n, 1, q, n_classes, top_n, DVL_PRECISION_FP32, DVL_F32, DVL_F32));
  if (scores == nullptr || labels == nullptr || ld_scores < n) return dvl::set_error(c, DVL_ERR_INVALID, "bad inputs");
  dvl::DeviceGuard guard(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t ldk = dvl::round_up(n, 256);
  DVL_TRY(dvl::ensure(c, c->keys[0], (size_t)q * ldk * 4));
  DVL_TRY(dvl::ensure(c, c->per_query, (size_t)q * 4 * sizeof(double)));
  DVL_TRY(dvl::ensure(c, c->seg_len, (size_t)q * sizeof(int32_t)));
  DVL_TRY(dvl::run_label_stats(c, labels, n, n_classes, (n > top_n ? n : top_n) + 2, st));
  DVL_TRY(dvl::launch_scores_to_keys(c, scores, ld_scores, (uint32_t*)c->keys[0].ptr, ldk, n, q, st));
  double* per_query = out_per_query != nullptr ? out_per_query : (double*)c->per_query.ptr;
  int32_t* ranked_len = out_ranked_len != nullptr ? out_ranked_len : (int32_t*)c->seg_len.ptr;
  DVL_TRY(dvl::rank_full(c, labels, n, n, q, n_classes, top_n, ldk, per_query, ranked_len, out_ranked_idx, ranked_cap,
                         st));
  int rc = dvl::finish(c, per_query, q, n_classes, out_mean4, true, st);
  if (rc == dvl::kRetryLsd) {      // bucket overflow: the keys are still in place, rank them with the radix sort
    c->force_lsd = true;
    rc = dvl::rank_full(c, labels, n, n, q, n_classes, top_n, ldk, per_query, ranked_len, out_ranked_idx, ranked_cap, st);
    if (rc == DVL_OK) rc = dvl::finish(c, per_query, q, n_classes, out_mean4, true, st);
    c->force_lsd = false;
  }
  return rc;
}

namespace dvl {
namespace {

// Re-ranks the queries the streaming route flagged (tie run longer than its list) by the full-sort
// route and patches their rows of per_query / ranked_len / ranked_idx.  Rare by construction.
int rerank_flagged(Ctx* c, const Operands& gallery_op, const void* queries, int qdt, const uint8_t* labels,
                   int64_t n, int d, int q, int n_classes, int64_t top_n, int precision, uint32_t flags,
                   double* per_query, int32_t* ranked_len, int32_t* ranked_idx, int64_t ranked_cap,
                   cudaStream_t st) {
  std::vector<uint32_t> host_flags((size_t)q);
  DVL_CUDA(c, cudaMemcpyAsync(host_flags.data(), c->cand_flags.ptr, (size_t)q * 4, cudaMemcpyDeviceToHost, st));
  DVL_CUDA(c, cudaStreamSynchronize(st));
  std::vector<int> flagged;
  for (int i = 0; i < q; ++i)
    if (host_flags[(size_t)i] != 0u) flagged.push_back(i);
  if (flagged.empty()) return DVL_OK;
  const size_t row_bytes = (size_t)d * dtype_size(qdt);
  const int64_t ldk = round_up(n, 256);
  // batches bounded by ~4 GB of sort workspace
  int batch = (int)(4.0e9 / ((double)ldk * 20.0));
  if (batch < 1) batch = 1;
  for (size_t b0 = 0; b0 < flagged.size(); b0 += (size_t)batch) {
    const int f = (int)((flagged.size() - b0 < (size_t)batch) ? flagged.size() - b0 : (size_t)batch);
    DVL_TRY(ensure(c, c->fb_queries, (size_t)f * row_bytes));
    DVL_TRY(ensure(c, c->fb_per_query, (size_t)f * 4 * sizeof(double)));
    DVL_TRY(ensure(c, c->fb_len, (size_t)f * 4));
    if (ranked_idx != nullptr) {
      DVL_TRY(ensure(c, c->fb_idx, (size_t)f * ranked_cap * 4));
      DVL_CUDA(c, cudaMemsetAsync(c->fb_idx.ptr, 0xFF, (size_t)f * ranked_cap * 4, st));   // -1 past the ranked rows
    }
    DVL_TRY(ensure(c, c->keys[0], (size_t)f * ldk * 4));
    {
      StageScope scope(c, DVL_STAGE_COPY, st);
      for (int i = 0; i < f; ++i)
        DVL_CUDA(c, cudaMemcpyAsync((char*)c->fb_queries.ptr + (size_t)i * row_bytes,
                                    (const char*)queries + (size_t)flagged[b0 + i] * row_bytes, row_bytes,
                                    cudaMemcpyDeviceToDevice, st));
    }
    // the gallery planes of the streaming pass are still in the workspace: only the flagged queries
    // are prepared again (the streaming pass is done with its own query planes)
    Operands op = gallery_op;
    DVL_TRY(prepare_queries(c, c->fb_queries.ptr, qdt, d, f, precision, flags, &op, st));
    DVL_TRY(run_similarity_prepared(c, op, n, d, f, (uint32_t*)c->keys[0].ptr, ldk, nullptr, 0, st));
    DVL_TRY(rank_full(c, labels, n, n, f, n_classes, top_n, ldk, (double*)c->fb_per_query.ptr,
                      (int32_t*)c->fb_len.ptr, ranked_idx != nullptr ? (int32_t*)c->fb_idx.ptr : nullptr, ranked_cap,
                      st));
    StageScope scope(c, DVL_STAGE_COPY, st);
    for (int i = 0; i < f; ++i) {
      const int qi = flagged[b0 + i];
      DVL_CUDA(c, cudaMemcpyAsync(per_query + (size_t)qi * 4, (double*)c->fb_per_query.ptr + (size_t)i * 4,
                                  4 * sizeof(double), cudaMemcpyDeviceToDevice, st));
      DVL_CUDA(c, cudaMemcpyAsync(ranked_len + qi, (int32_t*)c->fb_len.ptr + i, 4, cudaMemcpyDeviceToDevice, st));
      if (ranked_idx != nullptr)
        DVL_CUDA(c, cudaMemcpyAsync(ranked_idx + (size_t)qi * ranked_cap, (int32_t*)c->fb_idx.ptr + (size_t)i * ranked_cap,
                                    (size_t)ranked_cap * 4, cudaMemcpyDeviceToDevice, st));
    }
  }
  return DVL_OK;
}

}  // namespace
}  // namespace dvl

namespace dvl {
namespace {

// dvl_eval_ranking, with the gallery either resident (feed == nullptr) or still arriving from the host
int eval_ranking_impl(dvl_ctx* c, const void* gallery, int gdt, const void* queries, int qdt, const uint8_t* labels,
                      int64_t n, int32_t d, int32_t q, int32_t n_classes, int64_t top_n, int precision,
                      uint32_t flags, double* out_mean4, double* out_per_query, int32_t* out_ranked_len,
                      int32_t* out_ranked_idx, int64_t ranked_cap, float* out_scores, void* stream, RowFeed* feed) {
  DVL_TRY(dvl::check_common(c, n, d, q, n_classes, top_n, precision, gdt, qdt));
  if (gallery == nullptr || queries == nullptr || labels == nullptr)
    return dvl::set_error(c, DVL_ERR_INVALID, "null input pointer");
  if (out_ranked_idx != nullptr && ranked_cap < 1) return dvl::set_error(c, DVL_ERR_INVALID, "ranked_cap must be >= 1");
  c->radix_only = false;      // per-call state a failed earlier call may have left behind
  c->force_lsd = false;
  c->bucket_used = false;
  c->prepared_valid = false;  // this call prepares its own gallery into the shared workspace
  dvl::DeviceGuard guard(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  DVL_TRY(dvl::ensure(c, c->per_query, (size_t)q * 4 * sizeof(double)));
  DVL_TRY(dvl::ensure(c, c->seg_len, (size_t)q * sizeof(int32_t)));
  double* per_query = out_per_query != nullptr ? out_per_query : (double*)c->per_query.ptr;
  int32_t* ranked_len = out_ranked_len != nullptr ? out_ranked_len : (int32_t*)c->seg_len.ptr;

  dvl::StreamPlan plan;
  // rows still arriving: no pilot sample (it would need the whole gallery first), progressive thresholds only
  const bool stream_route = out_scores == nullptr &&
                            dvl::plan_stream(n, q, top_n, feed != nullptr ? (flags | DVL_FLAG_NO_PILOT) : flags, &plan);
  if (stream_route) {
    DVL_TRY(dvl::run_label_stats(c, labels, n, n_classes, (n > top_n ? n : top_n) + 2, st));
    {   // the ranked lists hold up to sort_cap rows, which can exceed n on a small forced-stream gallery
      dvl::StageScope scope(c, DVL_STAGE_PREP, st);
      DVL_TRY(dvl::ensure_metric_tables(c, (int64_t)plan.sort_cap + 2, st));
    }
    dvl::Operands op;
    DVL_TRY(dvl::prepare_operands(c, gallery, gdt, queries, qdt, n, d, q, precision, flags, &op, st, feed));
    DVL_TRY(dvl::stream_topk(c, op, labels, n, 0, d, q, top_n, plan, out_ranked_idx, ranked_cap, st, feed));
    dvl::MetricArgs m;
    m.keys = nullptr;
    m.labs = (const uint8_t*)c->topk_labs.ptr;
    m.ldk = 0;
    m.ldl = dvl::round_up(plan.sort_cap, 256);
    m.n_avail = plan.sort_cap;
    m.n_gallery = n;
    m.top_n = top_n;
    m.q = q;
    m.n_classes = n_classes;
    m.class_counts = (const unsigned long long*)c->class_counts.ptr;
    m.avail = (const int32_t*)c->topk_avail.ptr;
    m.per_query = per_query;
    m.ranked_len = ranked_len;
    {
      dvl::StageScope scope(c, DVL_STAGE_METRICS, st);
      DVL_TRY(dvl::launch_metrics(c, m, st));
    }
    // flagged queries are tie-heavy by construction (a tie run longer than a candidate list): radix sort
    c->force_lsd = true;
    const int rr = dvl::rerank_flagged(c, op, queries, qdt, labels, n, d, q, n_classes, top_n, precision, flags,
                                       per_query, ranked_len, out_ranked_idx, ranked_cap, st);
    c->force_lsd = false;
    DVL_TRY(rr);
    return dvl::finish(c, per_query, q, n_classes, out_mean4, true, st);
  }

  dvl::FullSortCall call;
  memset(&call, 0, sizeof(call));       // the struct is compared bytewise: padding must be defined
  call.gallery = gallery; call.queries = queries; call.labels = labels;
  call.gdt = gdt; call.qdt = qdt; call.n = n; call.d = d; call.q = q; call.n_classes = n_classes;
  call.top_n = top_n; call.precision = precision; call.flags = flags;
  call.per_query = per_query; call.ranked_len = ranked_len; call.ranked_idx = out_ranked_idx;
  call.ranked_cap = ranked_cap; call.out_scores = out_scores;
  return dvl::run_full_sort(c, call, out_mean4, st, feed);
}

}  // namespace
}  // namespace dvl

DVL_API int dvl_eval_ranking(dvl_ctx* c, const void* gallery, int gdt, const void* queries, int qdt, const uint8_t* labels,
                     int64_t n, int32_t d, int32_t q, int32_t n_classes, int64_t top_n, int precision,
                     uint32_t flags, double* out_mean4, double* out_per_query, int32_t* out_ranked_len,
                     int32_t* out_ranked_idx, int64_t ranked_cap, float* out_scores, void* stream) {
  return dvl::eval_ranking_impl(c, gallery, gdt, queries, qdt, labels, n, d, q, n_classes, top_n, precision, flags,
                                out_mean4, out_per_query, out_ranked_len, out_ranked_idx, ranked_cap, out_scores, stream,
                                nullptr);
}

DVL_API int dvl_eval_ranking_host(dvl_ctx* c, const void* gallery_host, int gdt, const void* queries_host, int qdt,
                          const uint8_t* labels_host, int64_t n, int32_t d, int32_t q, int32_t n_classes,
                          int64_t top_n, int precision, uint32_t flags, double* out_mean4,
                          double* out_per_query_host, void* stream) {
  DVL_TRY(dvl::check_common(c, n, d, q, n_classes, top_n, precision, gdt, qdt));
  if (gallery_host == nullptr || queries_host == nullptr || labels_host == nullptr)
    return dvl::set_error(c, DVL_ERR_INVALID, "null input pointer");
  dvl::DeviceGuard guard(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t gbytes = (size_t)n * d * dvl::dtype_size(gdt);
  const size_t qbytes = (size_t)q * d * dvl::dtype_size(qdt);
  DVL_TRY(dvl::ensure(c, c->h_gallery, gbytes));
  DVL_TRY(dvl::ensure(c, c->h_queries, qbytes));
  DVL_TRY(dvl::ensure(c, c->h_labels, (size_t)n));
  // queries and labels first, then the gallery in row blocks on the copy stream: operand split and
  // similarity of a block start when the block has landed, the rest of the upload continues underneath
  dvl::RowFeed feed;
  const dvl::SmallCopy small[2] = {{c->h_queries.ptr, queries_host, qbytes}, {c->h_labels.ptr, labels_host, (size_t)n}};
  DVL_TRY(dvl::start_upload(c, gallery_host, c->h_gallery.ptr, n, (size_t)d * dvl::dtype_size(gdt), small, 2, &feed, st));
  DVL_TRY(dvl::ensure(c, c->per_query, (size_t)q * 4 * sizeof(double)));
  int rc = dvl::eval_ranking_impl(c, c->h_gallery.ptr, gdt, c->h_queries.ptr, qdt, (const uint8_t*)c->h_labels.ptr, n, d,
                                  q, n_classes, top_n, precision, flags, out_mean4, (double*)c->per_query.ptr, nullptr,
                                  nullptr, 0, nullptr, stream, &feed);
  if (rc != DVL_OK) return rc;
  if (out_per_query_host != nullptr) {
    DVL_CUDA(c, cudaMemcpyAsync(out_per_query_host, c->per_query.ptr, (size_t)q * 4 * sizeof(double),
                                cudaMemcpyDeviceToHost, st));
    DVL_CUDA(c, cudaStreamSynchronize(st));
  }
  return DVL_OK;
}

// ---------------------------------------------------------------------------------------------
// multi-GPU: gallery shards.  The host (torch.distributed over NCCL) moves the small per-query
// lists between these stages; see debias_vision_lang_b200/sharded.py.
// ---------------------------------------------------------------------------------------------
namespace dvl {
namespace {

int shard_topk_impl(dvl_ctx* c, const void* gallery, int gdt, const void* queries, int qdt, const uint8_t* labels,
                    int64_t n_local, int64_t index_base, int32_t d, int32_t q, int32_t n_classes, int64_t top_n,
                    int precision, uint32_t flags, float* out_tau, uint32_t* out_flags,
                    unsigned long long* out_class_counts, int32_t* out_list_cap, cudaStream_t st, RowFeed* feed) {
  StreamPlan plan;
  // a shard always streams: top_n >= n_local just means the whole shard is kept
  if (top_n > 8192) return set_error(c, DVL_ERR_UNSUPPORTED, "sharded ranking supports top_n <= 8192");
  // a shard with fewer than top_n rows keeps all of them and leaves tau at -inf: its worst score says
  // nothing about the global top_n-th
  if (!plan_stream(n_local, q, top_n, feed != nullptr ? (flags | DVL_FLAG_NO_PILOT) : flags, &plan, true))
    return set_error(c, DVL_ERR_UNSUPPORTED, "no streaming plan for this shard");
  DVL_TRY(run_label_stats(c, labels, n_local, n_classes, top_n + 2, st));
  Operands op;
  DVL_TRY(prepare_operands(c, gallery, gdt, queries, qdt, n_local, d, q, precision, flags, &op, st, feed));
  DVL_TRY(stream_topk(c, op, labels, n_local, index_base, d, q, top_n, plan, nullptr, 0, st, feed));
  c->shard_plan_cap = plan.cap;
  c->shard_sort_cap = plan.sort_cap;
  c->shard_q = q;
  // what DVL_FLAG_REUSE_GALLERY may pick up: the gallery operand this call prepared (planes + row factors in the
  // workspace, or the caller's bf16 rows in place), valid until the next call that prepares a gallery
  c->prepared_gallery = op;
  c->prepared_n = n_local;
  c->prepared_d = d;
  c->prepared_precision = precision;
  c->prepared_flags = flags & DVL_FLAG_NORMALIZE_GALLERY;
  c->prepared_valid = true;
  // optional: the lean exchange reads the same three things from the context (dvl_shard_stats)
  if (out_tau != nullptr) DVL_CUDA(c, cudaMemcpyAsync(out_tau, c->cand_tau.ptr, (size_t)q * 4, cudaMemcpyDeviceToDevice, st));
  if (out_flags != nullptr)
    DVL_CUDA(c, cudaMemcpyAsync(out_flags, c->cand_flags.ptr, (size_t)q * 4, cudaMemcpyDeviceToDevice, st));
  if (out_class_counts != nullptr)
    DVL_CUDA(c, cudaMemcpyAsync(out_class_counts, c->class_counts.ptr, sizeof(unsigned long long) * (kMaxClasses + 1),
                                cudaMemcpyDeviceToDevice, st));
  if (out_list_cap != nullptr) *out_list_cap = plan.sort_cap;
  return DVL_OK;    // stream-ordered: the outputs are consumed by collectives enqueued after this call
}

}  // namespace
}  // namespace dvl

DVL_API int dvl_shard_topk(dvl_ctx* c, const void* gallery, int gdt, const void* queries, int qdt, const uint8_t* labels,
                   int64_t n_local, int64_t index_base, int32_t d, int32_t q, int32_t n_classes, int64_t top_n,
                   int precision, uint32_t flags, float* out_tau, uint32_t* out_flags,
                   unsigned long long* out_class_counts, int32_t* out_list_cap, void* stream) {
  DVL_TRY(dvl::check_common(c, n_local, d, q, n_classes, top_n, precision, gdt, qdt));
  if (gallery == nullptr || queries == nullptr || labels == nullptr) return dvl::set_error(c, DVL_ERR_INVALID, "null pointer");
  dvl::DeviceGuard guard(c->device);
  return dvl::shard_topk_impl(c, gallery, gdt, queries, qdt, labels, n_local, index_base, d, q, n_classes, top_n,
                              precision, flags, out_tau, out_flags, out_class_counts, out_list_cap, (cudaStream_t)stream,
                              nullptr);
}

DVL_API int dvl_shard_topk_host(dvl_ctx* c, const void* gallery_host, int gdt, const void* queries_host, int qdt,
                        const uint8_t* labels_host, int64_t n_local, int64_t index_base, int32_t d, int32_t q,
                        int32_t n_classes, int64_t top_n, int precision, uint32_t flags, float* out_tau,
                        uint32_t* out_flags, unsigned long long* out_class_counts, int32_t* out_list_cap,
                        const void** out_gallery_dev, const void** out_queries_dev, const uint8_t** out_labels_dev,
                        void* stream) {
  DVL_TRY(dvl::check_common(c, n_local, d, q, n_classes, top_n, precision, gdt, qdt));
  if (gallery_host == nullptr || queries_host == nullptr || labels_host == nullptr)
    return dvl::set_error(c, DVL_ERR_INVALID, "null pointer");
  dvl::DeviceGuard guard(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t row_bytes = (size_t)d * dvl::dtype_size(gdt);
  const size_t qbytes = (size_t)q * d * dvl::dtype_size(qdt);
  DVL_TRY(dvl::ensure(c, c->h_gallery, (size_t)n_local * row_bytes));
  DVL_TRY(dvl::ensure(c, c->h_queries, qbytes));
  DVL_TRY(dvl::ensure(c, c->h_labels, (size_t)n_local));
  dvl::RowFeed feed;
  const dvl::SmallCopy small[2] = {{c->h_queries.ptr, queries_host, qbytes}, {c->h_labels.ptr, labels_host, (size_t)n_local}};
  DVL_TRY(dvl::start_upload(c, gallery_host, c->h_gallery.ptr, n_local, row_bytes, small, 2, &feed, st));
  if (out_gallery_dev != nullptr) *out_gallery_dev = c->h_gallery.ptr;
  if (out_queries_dev != nullptr) *out_queries_dev = c->h_queries.ptr;
  if (out_labels_dev != nullptr) *out_labels_dev = (const uint8_t*)c->h_labels.ptr;
  return dvl::shard_topk_impl(c, c->h_gallery.ptr, gdt, c->h_queries.ptr, qdt, (const uint8_t*)c->h_labels.ptr, n_local,
                              index_base, d, q, n_classes, top_n, precision, flags, out_tau, out_flags, out_class_counts,
                              out_list_cap, st, &feed);
}

DVL_API int dvl_shard_trim(dvl_ctx* c, const float* tau_bound, int32_t q, int32_t* out_cnt, void* stream) {
  if (c == nullptr) return DVL_ERR_INVALID;
  if (q != c->shard_q || c->shard_plan_cap <= 0) return dvl::set_error(c, DVL_ERR_INVALID, "dvl_shard_topk has not run");
  dvl::DeviceGuard guard(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  dvl::StageScope scope(c, DVL_STAGE_SELECT, st);
  return dvl::launch_trim_count(c, (const unsigned long long*)c->cand.ptr, c->shard_plan_cap,
                                (const int32_t*)c->topk_avail.ptr, tau_bound, q, out_cnt, st);
}

DVL_API int dvl_shard_pack(dvl_ctx* c, const int32_t* cnt, int32_t q, int64_t len, unsigned long long* out, void* stream) {
  if (c == nullptr) return DVL_ERR_INVALID;
  if (q != c->shard_q || c->shard_plan_cap <= 0) return dvl::set_error(c, DVL_ERR_INVALID, "dvl_shard_topk has not run");
  if (len < 1 || len > c->shard_sort_cap) return dvl::set_error(c, DVL_ERR_INVALID, "bad packed list length");
  dvl::DeviceGuard guard(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  dvl::StageScope scope(c, DVL_STAGE_SELECT, st);
  return dvl::launch_pack_lists(c, (const unsigned long long*)c->cand.ptr, c->shard_plan_cap, cnt, q, len, out, st);
}

DVL_API int dvl_shard_stats(dvl_ctx* c, int32_t* out_stats, int32_t q, void* stream) {
  if (c == nullptr) return DVL_ERR_INVALID;
  if (q != c->shard_q || c->shard_plan_cap <= 0) return dvl::set_error(c, DVL_ERR_INVALID, "dvl_shard_topk has not run");
  if (out_stats == nullptr) return dvl::set_error(c, DVL_ERR_INVALID, "null pointer");
  dvl::DeviceGuard guard(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  dvl::StageScope scope(c, DVL_STAGE_SELECT, st);
  return dvl::launch_pack_stats(c, (const float*)c->cand_tau.ptr, (const uint32_t*)c->cand_flags.ptr,
                                (const unsigned long long*)c->class_counts.ptr, q, (uint32_t*)out_stats, st);
}

DVL_API int dvl_shard_reduce(dvl_ctx* c, const int32_t* stats_all, int32_t world, int32_t q, void* stream) {
  if (c == nullptr) return DVL_ERR_INVALID;
  if (stats_all == nullptr || world < 1 || q < 1) return dvl::set_error(c, DVL_ERR_INVALID, "bad arguments");
  dvl::DeviceGuard guard(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  DVL_TRY(dvl::ensure(c, c->shard_tau_bound, (size_t)q * 4));
  DVL_TRY(dvl::ensure(c, c->shard_flags_any, (size_t)q * 4));
  DVL_TRY(dvl::ensure(c, c->shard_counts, sizeof(unsigned long long) * (dvl::kMaxClasses + 1)));
  c->shard_reduced_q = q;
  dvl::StageScope scope(c, DVL_STAGE_SELECT, st);
  return dvl::launch_reduce_stats(c, (const uint32_t*)stats_all, world, q, (float*)c->shard_tau_bound.ptr,
                                  (uint32_t*)c->shard_flags_any.ptr, (unsigned long long*)c->shard_counts.ptr, st);
}

DVL_API int dvl_shard_trim_pack(dvl_ctx* c, int32_t q, int64_t len, unsigned long long* out, int32_t* out_need,
                                void* stream) {
  if (c == nullptr) return DVL_ERR_INVALID;
  if (q != c->shard_q || c->shard_plan_cap <= 0) return dvl::set_error(c, DVL_ERR_INVALID, "dvl_shard_topk has not run");
  if (q != c->shard_reduced_q) return dvl::set_error(c, DVL_ERR_INVALID, "dvl_shard_reduce has not run");
  if (len < 1 || out == nullptr || out_need == nullptr) return dvl::set_error(c, DVL_ERR_INVALID, "bad arguments");
  dvl::DeviceGuard guard(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  dvl::StageScope scope(c, DVL_STAGE_SELECT, st);
  return dvl::launch_trim_pack(c, (const unsigned long long*)c->cand.ptr, c->shard_plan_cap,
                               (const int32_t*)c->topk_avail.ptr, (const float*)c->shard_tau_bound.ptr, q, len, out,
                               out_need, st);
}

DVL_API int dvl_shard_finish(dvl_ctx* c, const void* records_all, int32_t world, int32_t block, int64_t rec_bytes,
                             int32_t q, double* out_mean4, int64_t* out_status3, void* stream) {
  if (c == nullptr) return DVL_ERR_INVALID;
  if (records_all == nullptr || world < 1 || block < 1 || q < 1 || q > world * block || rec_bytes < (int64_t)block * 40 + 4 ||
      rec_bytes % 8 != 0 || out_mean4 == nullptr || out_status3 == nullptr)
    return dvl::set_error(c, DVL_ERR_INVALID, "bad arguments");
  if (q != c->shard_reduced_q) return dvl::set_error(c, DVL_ERR_INVALID, "dvl_shard_reduce has not run");
  dvl::DeviceGuard guard(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  DVL_TRY(dvl::ensure(c, c->shard_out, 8 * sizeof(double)));
  {
    dvl::StageScope scope(c, DVL_STAGE_METRICS, st);
    DVL_TRY(dvl::launch_shard_finish(c, (const uint8_t*)records_all, world, block, rec_bytes, q,
                                     (const uint32_t*)c->shard_flags_any.ptr,
                                     (const unsigned long long*)c->shard_counts.ptr, (double*)c->shard_out.ptr, st));
  }
  DVL_CUDA(c, cudaMemcpyAsync(c->mean_pinned, c->shard_out.ptr, 7 * sizeof(double), cudaMemcpyDeviceToHost, st));
  DVL_CUDA(c, cudaStreamSynchronize(st));
  memcpy(out_mean4, c->mean_pinned, 4 * sizeof(double));
  for (int i = 0; i < 3; ++i) out_status3[i] = (int64_t)c->mean_pinned[4 + i];
  return DVL_OK;
}

DVL_API int dvl_merge_topk(dvl_ctx* c, const unsigned long long* gathered, int32_t world, int32_t q_total, int64_t len,
                   int32_t q0, int32_t qn, int64_t n_global, const unsigned long long* class_counts_global,
                   int32_t n_classes, int64_t top_n, double* out_per_query, int32_t* out_ranked_len,
                   int32_t* out_ranked_idx, int64_t ranked_cap, uint32_t* out_flags, void* stream) {
  if (c == nullptr) return DVL_ERR_INVALID;
  if (class_counts_global == nullptr && c->shard_counts.ptr == nullptr)
    return dvl::set_error(c, DVL_ERR_INVALID, "no global label counts: pass them or run dvl_shard_reduce first");
  if (gathered == nullptr || out_per_query == nullptr || out_ranked_len == nullptr ||
      out_flags == nullptr || world < 1 || qn < 1 || q0 < 0 || q0 + qn > q_total || len < 1 || top_n < 1)
    return dvl::set_error(c, DVL_ERR_INVALID, "bad merge arguments");
  if (n_classes < 1 || n_classes > DVL_MAX_CLASSES) return dvl::set_error(c, DVL_ERR_INVALID, "n_classes out of range");
  dvl::DeviceGuard guard(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  int s = 2048;
  while (s < 2 * top_n && s < 16384) s <<= 1;
  const int64_t ldl = dvl::round_up(s, 256);
  DVL_TRY(dvl::ensure(c, c->merge_sorted, (size_t)qn * s * 8));
  // the merge's own ranked labels / lengths: the local lists' (topk_labs / topk_avail) must survive it -- a
  // too-short speculative list length is repaired by trimming and packing the local lists a second time
  DVL_TRY(dvl::ensure(c, c->merge_labs, (size_t)qn * ldl));
  DVL_TRY(dvl::ensure(c, c->merge_avail, (size_t)qn * 4));
  // NULL: the counts dvl_shard_reduce summed over the ranks (no copy, and this rank's own counts stay intact)
  const unsigned long long* counts_global =
      class_counts_global != nullptr ? class_counts_global : (const unsigned long long*)c->shard_counts.ptr;
  DVL_CUDA(c, cudaMemsetAsync(out_flags, 0, (size_t)qn * 4, st));
  {
    dvl::StageScope scope(c, DVL_STAGE_PREP, st);
    DVL_TRY(dvl::ensure_metric_tables(c, (s > top_n ? s : top_n) + 2, st));
  }
  dvl::SelectArgs sel;
  sel.cand = const_cast<unsigned long long*>(gathered) + (int64_t)q0 * len;
  sel.row_stride = len;
  sel.seg_stride = (int64_t)q_total * len;
  sel.segs = world;
  sel.seg_len = len;
  sel.cnt = nullptr;
  sel.cap = 0;
  sel.tau = nullptr;
  sel.flags = out_flags;
  sel.k = top_n;
  sel.sort_cap = s;
  sel.trigger = 0;
  sel.final_pass = true;
  sel.sorted_out = (unsigned long long*)c->merge_sorted.ptr;
  sel.labs = (uint8_t*)c->merge_labs.ptr;
  sel.ldl = ldl;
  sel.idx_out = out_ranked_idx;
  sel.idx_cap = ranked_cap;
  sel.avail = (int32_t*)c->merge_avail.ptr;
  sel.q_offset = 0;
  sel.q_count = qn;
  {
    dvl::StageScope scope(c, DVL_STAGE_SELECT, st);
    DVL_TRY(dvl::launch_select(c, sel, st));
  }
  dvl::MetricArgs m;
  m.keys = nullptr;
  m.labs = (const uint8_t*)c->merge_labs.ptr;
  m.ldk = 0;
  m.ldl = ldl;
  m.n_avail = s;
  m.n_gallery = n_global;
  m.top_n = top_n;
  m.q = qn;
  m.n_classes = n_classes;
  m.class_counts = counts_global;
  m.avail = (const int32_t*)c->merge_avail.ptr;
  m.per_query = out_per_query;
  m.ranked_len = out_ranked_len;
  {
    dvl::StageScope scope(c, DVL_STAGE_METRICS, st);
    DVL_TRY(dvl::launch_metrics(c, m, st));
  }
  return DVL_OK;    // stream-ordered, no host synchronisation
}

DVL_API int dvl_mean(dvl_ctx* c, const double* per_query, int32_t q, double* out_mean4, void* stream) {
  if (c == nullptr) return DVL_ERR_INVALID;
  if (per_query == nullptr || q < 1) return dvl::set_error(c, DVL_ERR_INVALID, "bad arguments");
  dvl::DeviceGuard guard(c->device);
  return dvl::finish(c, per_query, q, 0, out_mean4, false, (cudaStream_t)stream);
}

}  // extern "C"
