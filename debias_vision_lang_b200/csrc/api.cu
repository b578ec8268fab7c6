// C ABI of libdvl_b200.so (include/dvl_b200.h): context, workspace, and the pipelines that
// chain the kernels of prep.cu / gemm.cu / segsort.cu / metrics.cu on the caller's stream.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <new>

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
  if (precision != DVL_PRECISION_FP32 && precision != DVL_PRECISION_BF16)
    return set_error(c, DVL_ERR_INVALID, "unknown precision %d", precision);
  if (dtype_size(gdt) == 0 || dtype_size(qdt) == 0) return set_error(c, DVL_ERR_INVALID, "unknown dtype");
  return DVL_OK;
}

// similarity into ranking keys and/or fp32 scores (both pitched to whole 256-column tiles)
int run_similarity(Ctx* c, const void* gallery, int gdt, const void* queries, int qdt, int64_t n, int d, int q,
                   int precision, uint32_t flags, uint32_t* keys, int64_t ldk, float* scores, int64_t lds,
                   cudaStream_t st) {
  const int planes = (precision == DVL_PRECISION_FP32) ? 3 : 1;
  const int dp = (int)round_up(d, 8);
  DVL_TRY(ensure(c, c->gal_planes, (size_t)planes * n * dp * 2));
  DVL_TRY(ensure(c, c->qry_planes, (size_t)planes * q * dp * 2));
  float* gnorm = nullptr;
  float* qnorm = nullptr;
  if (flags & DVL_FLAG_NORMALIZE_GALLERY) {
    DVL_TRY(ensure(c, c->gal_inv_norm, (size_t)n * 4));
    gnorm = (float*)c->gal_inv_norm.ptr;
  }
  if (flags & DVL_FLAG_NORMALIZE_QUERIES) {
    DVL_TRY(ensure(c, c->qry_inv_norm, (size_t)q * 4));
    qnorm = (float*)c->qry_inv_norm.ptr;
  }
  {
    StageScope scope(c, DVL_STAGE_PREP, st);
    DVL_TRY(launch_split_rows(c, gallery, gdt, n, d, dp, planes, (__nv_bfloat16*)c->gal_planes.ptr, gnorm, st));
    DVL_TRY(launch_split_rows(c, queries, qdt, q, d, dp, planes, (__nv_bfloat16*)c->qry_planes.ptr, qnorm, st));
  }
  StageScope scope(c, DVL_STAGE_GEMM, st);
  GemmArgs g;
  g.qry_planes = (const __nv_bfloat16*)c->qry_planes.ptr;
  g.gal_planes = (const __nv_bfloat16*)c->gal_planes.ptr;
  g.planes = planes;
  g.n = n;
  g.q = q;
  g.d = d;
  g.dp = dp;
  g.qry_scale = qnorm;
  g.gal_scale = gnorm;
  g.keys = keys;
  g.ldk = ldk;
  g.scores = scores;
  g.lds = lds;
  return launch_sim_gemm(c, g, st);
}

// keys [q][ldk] in c->keys[0] -> sort -> metrics -> mean.  Shared by eval_ranking (full-sort
// route) and rank_and_score.
int run_sort_and_score(Ctx* c, const uint8_t* labels, int64_t n, int q, int n_classes, int64_t top_n, int64_t ldk,
                       double* out_mean4, double* out_per_query, int32_t* out_ranked_len, int32_t* out_ranked_idx,
                       int64_t ranked_cap, cudaStream_t st) {
  const bool want_idx = out_ranked_idx != nullptr;
  if (want_idx && ranked_cap < 1) return set_error(c, DVL_ERR_INVALID, "ranked_cap must be >= 1");
  DVL_TRY(ensure(c, c->keys[1], (size_t)q * ldk * 4));
  DVL_TRY(ensure(c, c->labs[0], (size_t)q * ldk));
  DVL_TRY(ensure(c, c->labs[1], (size_t)q * ldk));
  if (want_idx) {
    DVL_TRY(ensure(c, c->idxs[0], (size_t)q * ldk * 4));
    DVL_TRY(ensure(c, c->idxs[1], (size_t)q * ldk * 4));
  }
  DVL_TRY(ensure(c, c->class_counts, sizeof(unsigned long long) * (kMaxClasses + 1)));
  DVL_TRY(ensure(c, c->per_query, (size_t)q * 4 * sizeof(double)));
  DVL_TRY(ensure(c, c->seg_len, (size_t)q * sizeof(int32_t)));
  DVL_TRY(ensure(c, c->mean_dev, 8 * sizeof(double)));
  {
    StageScope scope(c, DVL_STAGE_PREP, st);
    DVL_TRY(ensure_metric_tables(c, (n > top_n ? n : top_n) + 2, st));
    DVL_TRY(launch_class_counts(c, labels, n, n_classes, (unsigned long long*)c->class_counts.ptr, st));
  }

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
  int buf = 0;
  DVL_TRY(launch_segmented_sort(c, s, &buf, st));

  int32_t* ranked_len = out_ranked_len != nullptr ? out_ranked_len : (int32_t*)c->seg_len.ptr;
  MetricArgs m;
  m.keys = s.keys[buf];
  m.labs = s.labs[buf];
  m.ldk = ldk;
  m.ldl = ldk;
  m.n_avail = n;
  m.n_gallery = n;
  m.top_n = top_n;
  m.q = q;
  m.n_classes = n_classes;
  m.class_counts = (const unsigned long long*)c->class_counts.ptr;
  m.avail = nullptr;
  m.per_query = out_per_query != nullptr ? out_per_query : (double*)c->per_query.ptr;
  m.ranked_len = ranked_len;
  {
    StageScope scope(c, DVL_STAGE_METRICS, st);
    DVL_TRY(launch_metrics(c, m, st));
    DVL_TRY(launch_mean(c, m.per_query, q, (double*)c->mean_dev.ptr, st));
  }
  if (want_idx) {
    StageScope scope(c, DVL_STAGE_COPY, st);
    DVL_TRY(launch_copy_ranked(c, s.idxs[buf], ldk, ranked_len, q, out_ranked_idx, ranked_cap, st));
  }
  DVL_CUDA(c, cudaMemcpyAsync(c->mean_pinned, c->mean_dev.ptr, 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
  // the out-of-range label flag travels with the class counts
  unsigned long long bad_flag = 0;
  DVL_CUDA(c, cudaMemcpyAsync(c->mean_pinned + 4, (unsigned long long*)c->class_counts.ptr + kMaxClasses,
                              sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  DVL_CUDA(c, cudaStreamSynchronize(st));
  memcpy(&bad_flag, c->mean_pinned + 4, sizeof(bad_flag));
  if (bad_flag != 0) return set_error(c, DVL_ERR_INVALID, "a label is >= n_classes (%d)", n_classes);
  if (out_mean4 != nullptr) memcpy(out_mean4, c->mean_pinned, 4 * sizeof(double));
  return DVL_OK;
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
  dvl::DevBuf* bufs[] = {&c->gal_planes, &c->qry_planes, &c->gal_inv_norm, &c->qry_inv_norm, &c->keys[0], &c->keys[1],
                         &c->labs[0], &c->labs[1], &c->idxs[0], &c->idxs[1], &c->tile_hist, &c->class_counts,
                         &c->log_tab, &c->per_query, &c->mean_dev, &c->scores_tmp, &c->seg_len, &c->sub_cnt, &c->partial,
                         &c->h_gallery,
                         &c->h_queries, &c->h_labels};
  for (dvl::DevBuf* b : bufs)
    if (b->ptr != nullptr) cudaFree(b->ptr);
  if (c->mean_pinned != nullptr) cudaFreeHost(c->mean_pinned);
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

DVL_API int dvl_similarity(dvl_ctx* c, const void* gallery, int gdt, const void* queries, int qdt, int64_t n, int32_t d,
                   int32_t q, int precision, uint32_t flags, float* out, int64_t ld_out, void* stream) {
  DVL_TRY(dvl::check_common(c, n, d, q, 1, 1, precision, gdt, qdt));
  if (out == nullptr || ld_out < n) return dvl::set_error(c, DVL_ERR_INVALID, "bad output matrix");
  dvl::DeviceGuard guard(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t lds = dvl::round_up(n, 256);
  DVL_TRY(dvl::ensure(c, c->scores_tmp, (size_t)q * lds * 4));
  DVL_TRY(dvl::run_similarity(c, gallery, gdt, queries, qdt, n, d, q, precision, flags, nullptr, 0,
                              (float*)c->scores_tmp.ptr, lds, st));
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
  DVL_TRY(dvl::launch_scores_to_keys(c, scores, ld_scores, (uint32_t*)c->keys[0].ptr, ldk, n, q, st));
  return dvl::run_sort_and_score(c, labels, n, q, n_classes, top_n, ldk, out_mean4, out_per_query, out_ranked_len,
                                 out_ranked_idx, ranked_cap, st);
}

DVL_API int dvl_eval_ranking(dvl_ctx* c, const void* gallery, int gdt, const void* queries, int qdt, const uint8_t* labels,
                     int64_t n, int32_t d, int32_t q, int32_t n_classes, int64_t top_n, int precision,
                     uint32_t flags, double* out_mean4, double* out_per_query, int32_t* out_ranked_len,
                     int32_t* out_ranked_idx, int64_t ranked_cap, float* out_scores, void* stream) {
  DVL_TRY(dvl::check_common(c, n, d, q, n_classes, top_n, precision, gdt, qdt));
  if (gallery == nullptr || queries == nullptr || labels == nullptr)
    return dvl::set_error(c, DVL_ERR_INVALID, "null input pointer");
  dvl::DeviceGuard guard(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t ldk = dvl::round_up(n, 256);
  DVL_TRY(dvl::ensure(c, c->keys[0], (size_t)q * ldk * 4));
  float* scores_tmp = nullptr;
  if (out_scores != nullptr) {
    DVL_TRY(dvl::ensure(c, c->scores_tmp, (size_t)q * ldk * 4));
    scores_tmp = (float*)c->scores_tmp.ptr;
  }
  DVL_TRY(dvl::run_similarity(c, gallery, gdt, queries, qdt, n, d, q, precision, flags, (uint32_t*)c->keys[0].ptr,
                              ldk, scores_tmp, ldk, st));
  if (out_scores != nullptr)
    DVL_CUDA(c, cudaMemcpy2DAsync(out_scores, (size_t)n * 4, scores_tmp, (size_t)ldk * 4, (size_t)n * 4, q,
                                  cudaMemcpyDeviceToDevice, st));
  return dvl::run_sort_and_score(c, labels, n, q, n_classes, top_n, ldk, out_mean4, out_per_query, out_ranked_len,
                                 out_ranked_idx, ranked_cap, st);
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
  {
    dvl::StageScope scope(c, DVL_STAGE_COPY, st);
    DVL_CUDA(c, cudaMemcpyAsync(c->h_gallery.ptr, gallery_host, gbytes, cudaMemcpyHostToDevice, st));
    DVL_CUDA(c, cudaMemcpyAsync(c->h_queries.ptr, queries_host, qbytes, cudaMemcpyHostToDevice, st));
    DVL_CUDA(c, cudaMemcpyAsync(c->h_labels.ptr, labels_host, (size_t)n, cudaMemcpyHostToDevice, st));
  }
  DVL_TRY(dvl::ensure(c, c->per_query, (size_t)q * 4 * sizeof(double)));
  int rc = dvl_eval_ranking(c, c->h_gallery.ptr, gdt, c->h_queries.ptr, qdt, (const uint8_t*)c->h_labels.ptr, n, d, q,
                            n_classes, top_n, precision, flags, out_mean4, (double*)c->per_query.ptr, nullptr,
                            nullptr, 0, nullptr, stream);
  if (rc != DVL_OK) return rc;
  if (out_per_query_host != nullptr) {
    DVL_CUDA(c, cudaMemcpyAsync(out_per_query_host, c->per_query.ptr, (size_t)q * 4 * sizeof(double),
                                cudaMemcpyDeviceToHost, st));
    DVL_CUDA(c, cudaStreamSynchronize(st));
  }
  return DVL_OK;
}

}  // extern "C"
