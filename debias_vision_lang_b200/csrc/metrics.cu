// Reduction stage: from the labels of the ranked rows of each query,
//   MaxSkew@k  (measuring_bias.py:47-72)  and  NDKL (measuring_bias.py:17-44),
// both desired distributions (eq_opp, dem_par), all arithmetic in float64 like the
// reference's numpy/math, then the mean over queries (measuring_bias.py:119).
//
// One CTA per query.  Each warp owns a contiguous slice of the ranked prefix and walks it
// 32 ranks at a time: per class c one ballot + popc gives every lane its prefix count
// h_c(i); with S1 = sum_c h_c ln h_c and S2 = sum_c h_c ln d_c
//     KL(h(i)/i || d) = (S1 - S2)/i - ln i
// so a rank costs C table look-ups (ln from an fp64 table) and no transcendental.
// m (the number of ranked rows, > top_n when boundary ties were kept; selectn.py:176-182)
// is found by binary search in the sorted keys.
#include "dvl_common.cuh"
#include "dvl_internal.h"

namespace dvl {

namespace {

constexpr int kMetricThreads = 256;
constexpr int kMetricWarps = kMetricThreads / 32;
constexpr double kLn2 = 0.693147180559945309417232121458176568;

// Z(top_n) = sum_{i=1..top_n} 1/log2(i+1)   (measuring_bias.py:39), deterministic tree
__global__ void __launch_bounds__(1024) ndkl_norm_kernel(const double* __restrict__ ln_tab, int64_t top_n,
                                                         double* __restrict__ z_out) {
  __shared__ double sh[32];
  double acc = 0.0;
  for (int64_t i = 1 + threadIdx.x; i <= top_n; i += blockDim.x) acc += kLn2 / ln_tab[i + 1];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = sh[threadIdx.x];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (threadIdx.x == 0) *z_out = v;
  }
}

struct MetricParams {
  const uint32_t* keys;
  const uint8_t* labs;
  int64_t ldk, ldl;
  int64_t n_avail, n_gallery, top_n;
  int n_classes;
  const unsigned long long* class_counts;
  const double* ln_tab;
  const double* z;
  double* per_query;
  int32_t* ranked_len;
};

__global__ void __launch_bounds__(kMetricThreads) metrics_kernel(const MetricParams p) {
  __shared__ unsigned int wcnt[kMetricWarps][kMaxClasses];
  __shared__ double wacc[kMetricWarps][2];
  __shared__ long long sh_m;
  const int q = blockIdx.x;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int C = p.n_classes;
  const uint8_t* lab = p.labs + (int64_t)q * p.ldl;

  if (threadIdx.x == 0) {
    long long m = p.n_avail;
    if (p.keys != nullptr) {
      const uint32_t* k = p.keys + (int64_t)q * p.ldk;
      // rows with a NaN score are dropped (selectn.py:133-134): they carry the largest key
      long long lo = 0, hi = p.n_avail;
      while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (k[mid] < kNanKey) lo = mid + 1; else hi = mid;
      }
      const long long n_valid = lo;
      m = n_valid;
      if (p.top_n < n_valid) {
        const uint32_t kth = k[p.top_n - 1];
        lo = p.top_n; hi = n_valid;            // first position whose key is > kth
        while (lo < hi) {
          const long long mid = (lo + hi) >> 1;
          if (k[mid] <= kth) lo = mid + 1; else hi = mid;
        }
        m = lo;
      }
    } else if (p.top_n < m) {
      m = p.top_n;
    }
    sh_m = m;
  }
  __syncthreads();
  const long long m = sh_m;
  // contiguous slice per warp, a multiple of 32 ranks
  long long slice = (m + kMetricWarps - 1) / kMetricWarps;
  slice = (slice + 31) & ~31ll;
  const long long begin = (long long)warp * slice;
  long long end = begin + slice;
  if (end > m) end = m;

  // phase 1: class counts of the slice
  unsigned int cnt[kMaxClasses];
#pragma unroll
  for (int c = 0; c < kMaxClasses; ++c) cnt[c] = 0;
  for (long long r0 = begin; r0 < end; r0 += 32) {
    const long long r = r0 + lane;
    const int l = (r < end) ? (int)lab[r] : 255;
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c)
      if (c < C) cnt[c] += __popc(__ballot_sync(0xffffffffu, l == c));
  }
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c) wcnt[warp][c] = cnt[c];
  }
  __syncthreads();

  // per-class constants: ln d_c for dem_par (max(count,1)/N, measuring_bias.py:27)
  double ln_d[kMaxClasses];
  unsigned int base[kMaxClasses];
  int c_present = 0;
#pragma unroll
  for (int c = 0; c < kMaxClasses; ++c) {
    ln_d[c] = 0.0;
    base[c] = 0;
    if (c < C) {
      const unsigned long long nc = p.class_counts[c];
      c_present += (nc > 0);
      ln_d[c] = log((double)(nc > 0 ? nc : 1ull) / (double)p.n_gallery);
      for (int w = 0; w < warp; ++w) base[c] += wcnt[w][c];
    }
  }
  const double ln_c = log((double)c_present);

  // phase 2: discounted KL of every prefix in the slice
  const unsigned int le_mask = (lane == 31) ? 0xffffffffu : ((2u << lane) - 1u);
  double acc_eq = 0.0, acc_dp = 0.0;
  for (long long r0 = begin; r0 < end; r0 += 32) {
    const long long r = r0 + lane;
    const bool ok = r < end;
    const int l = ok ? (int)lab[r] : 255;
    double s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c) {
      if (c < C) {
        const unsigned int mask = __ballot_sync(0xffffffffu, l == c);
        const unsigned int h = base[c] + __popc(mask & le_mask);
        base[c] += __popc(mask);
        const double hd = (double)h;
        s1 = fma(hd, p.ln_tab[h], s1);
        s2 = fma(hd, ln_d[c], s2);
      }
    }
    if (ok) {
      const long long i = r + 1;
      const double inv_i = 1.0 / (double)i;
      const double ln_i = p.ln_tab[i];
      const double w = kLn2 / p.ln_tab[i + 1];        // 1 / log2(i + 1)
      acc_eq = fma(fma(s1, inv_i, ln_c - ln_i), w, acc_eq);
      acc_dp = fma(fma(s1 - s2, inv_i, -ln_i), w, acc_dp);
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    acc_eq += __shfl_xor_sync(0xffffffffu, acc_eq, off);
    acc_dp += __shfl_xor_sync(0xffffffffu, acc_dp, off);
  }
  if (lane == 0) {
    wacc[warp][0] = acc_eq;
    wacc[warp][1] = acc_dp;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ndkl_eq = 0.0, ndkl_dp = 0.0;
    for (int w = 0; w < kMetricWarps; ++w) {
      ndkl_eq += wacc[w][0];
      ndkl_dp += wacc[w][1];
    }
    const double z = *p.z;
    // MaxSkew over the m ranked rows (measuring_bias.py:58-70)
    double ms_eq = 0.0, ms_dp = 0.0;
    for (int c = 0; c < C; ++c) {
      const unsigned long long nc = p.class_counts[c];
      if (nc == 0) continue;                   // Counter(df.label) has no such key
      unsigned int h = 0;
      for (int w = 0; w < kMetricWarps; ++w) h += wcnt[w][c];
      double pp = (double)h / (double)p.top_n;
      if (pp == 0.0) pp = 1.0 / (double)p.top_n;
      const double lp = log(pp);
      ms_eq = fmax(ms_eq, lp - log(1.0 / (double)c_present));
      ms_dp = fmax(ms_dp, lp - log((double)nc / (double)p.n_gallery));
    }
    double* out = p.per_query + (int64_t)q * 4;
    out[DVL_MAXSKEW_EQ_OPP] = ms_eq;
    out[DVL_MAXSKEW_DEM_PAR] = ms_dp;
    out[DVL_NDKL_EQ_OPP] = ndkl_eq / z;
    out[DVL_NDKL_DEM_PAR] = ndkl_dp / z;
    if (p.ranked_len != nullptr) p.ranked_len[q] = (int32_t)m;
  }
}

// mean over queries, fixed reduction tree (measuring_bias.py:119)
__global__ void __launch_bounds__(1024) mean_kernel(const double* __restrict__ per_query, int q,
                                                    double* __restrict__ mean_out) {
  __shared__ double sh[32][4];
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int i = threadIdx.x; i < q; i += blockDim.x) {
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] += per_query[(int64_t)i * 4 + k];
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], off);
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) sh[threadIdx.x >> 5][k] = acc[k];
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double v = 0.0;
    for (int w = 0; w < 32; ++w) v += sh[w][threadIdx.x];
    mean_out[threadIdx.x] = v / (double)q;
  }
}

__global__ void __launch_bounds__(256) copy_ranked_kernel(const int32_t* __restrict__ idx_sorted, int64_t ldi,
                                                          const int32_t* __restrict__ ranked_len,
                                                          int32_t* __restrict__ out, int64_t cap) {
  const int q = blockIdx.y;
  int64_t len = ranked_len[q];
  if (len > cap) len = cap;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x)
    out[(int64_t)q * cap + i] = idx_sorted[(int64_t)q * ldi + i];
}

}  // namespace

int launch_metrics(Ctx* c, const MetricArgs& a, cudaStream_t st) {
  if (a.n_classes < 1 || a.n_classes > kMaxClasses) return set_error(c, DVL_ERR_INVALID, "n_classes out of range");
  if (a.top_n < 1) return set_error(c, DVL_ERR_INVALID, "top_n must be >= 1");
  const int64_t tab_need = (a.n_avail > a.top_n ? a.n_avail : a.top_n) + 2;
  if (tab_need > c->log_tab_len) return set_error(c, DVL_ERR_INVALID, "log table too short");
  DVL_TRY(ensure(c, c->mean_dev, 8 * sizeof(double)));
  double* z = (double*)c->mean_dev.ptr + 4;
  ndkl_norm_kernel<<<1, 1024, 0, st>>>(a.log_tab, a.top_n, z);
  DVL_LAUNCH_CHECK(c, "ndkl_norm_kernel");
  MetricParams p;
  p.keys = a.keys;
  p.labs = a.labs;
  p.ldk = a.ldk;
  p.ldl = a.ldl;
  p.n_avail = a.n_avail;
  p.n_gallery = a.n_gallery;
  p.top_n = a.top_n;
  p.n_classes = a.n_classes;
  p.class_counts = a.class_counts;
  p.ln_tab = a.log_tab;
  p.z = z;
  p.per_query = a.per_query;
  p.ranked_len = a.ranked_len;
  metrics_kernel<<<(unsigned)a.q, kMetricThreads, 0, st>>>(p);
  DVL_LAUNCH_CHECK(c, "metrics_kernel");
  return DVL_OK;
}

int launch_mean(Ctx* c, const double* per_query, int q, double* mean_dev, cudaStream_t st) {
  mean_kernel<<<1, 1024, 0, st>>>(per_query, q, mean_dev);
  DVL_LAUNCH_CHECK(c, "mean_kernel");
  return DVL_OK;
}

int launch_copy_ranked(Ctx* c, const int32_t* idx_sorted, int64_t ldi, const int32_t* ranked_len, int q,
                       int32_t* out, int64_t cap, cudaStream_t st) {
  int64_t bx = (cap + 255) / 256;
  if (bx > 1024) bx = 1024;
  if (bx < 1) bx = 1;
  for (int q0 = 0; q0 < q; q0 += 65535) {
    const int qn = (q - q0) < 65535 ? (q - q0) : 65535;
    dim3 grid((unsigned)bx, (unsigned)qn);
    copy_ranked_kernel<<<grid, 256, 0, st>>>(idx_sorted + (int64_t)q0 * ldi, ldi, ranked_len + q0,
                                             out + (int64_t)q0 * cap, cap);
    DVL_LAUNCH_CHECK(c, "copy_ranked_kernel");
  }
  return DVL_OK;
}

}  // namespace dvl
