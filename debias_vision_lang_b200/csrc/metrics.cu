// Reduction stage: from the labels of the ranked rows of each query,
//   MaxSkew@k  (measuring_bias.py:47-72)  and  NDKL (measuring_bias.py:17-44),
// both desired distributions (eq_opp, dem_par), all arithmetic in float64 like the
// reference's numpy/math, then the mean over queries (measuring_bias.py:119).
//
// NDKL without transcendental or division work per ranked row, and with work per row that does
// not grow with the number of classes.  With h_c(i) the prefix class counts, w_i = 1/log2(i+1):
//     KL(h(i)/i || d) = (1/i) [sum_c h_c ln h_c - sum_c h_c ln d_c] - ln i
// Both sums only change, at rank j, in the term of row j's own class c = l_j, whose count becomes
// t_j (the row is the t_j-th of its class).  Summing by parts over i >= j:
//     sum_i w_i KL_i = sum_j T_j (DXL[t_j] - ln d_{l_j})  -  PWL[m]
//         T_j    = PWI[m] - PWI[j-1],   PWI[i] = sum_{r<=i} w_r / r
//         DXL[t] = t ln t - (t-1) ln(t-1)
//         PWL[m] = sum_{i<=m} w_i ln i,  PW[m] = sum_{i<=m} w_i   (eq_opp: ln d = -ln C gives + ln C * PW[m])
// and Z(top_n) = PW[top_n] (measuring_bias.py:39).  DXL, PWI, PW, PWL are fp64 tables built once
// per context, so a ranked row costs one ballot+popc per class (its own-class prefix count),
// two table look-ups and three fp64 FMAs.
//
// Kernels (all HBM/L2-bound byte work, one byte per ranked row):
//   ranked_len   m_q: binary search for the end of the top_n-th key's tie run (keep="all",
//                selectn.py:176-182) and for the first NaN key (dropped rows, selectn.py:133-134)
//   sub_counts   class counts of every 1024-row sub-slice (one warp each, SIMD byte compares)
//   sub_scan     exclusive prefix over the sub-slices of a query
//   ndkl_partial one warp per sub-slice: 32 rounds of warp-ballot own-class prefix counts
//   finalize     fixed-order sum of the partials, MaxSkew from the total counts, 4 doubles/query
//   mean         fixed-tree mean over the queries
#include <type_traits>

#include "dvl_common.cuh"
#include "dvl_internal.h"

namespace dvl {

namespace {

constexpr int kSub = 1024;            // ranked rows per warp-sized sub-slice
constexpr int kWarpsPerCta = 4;
constexpr int kQueriesPerCta = 8;     // ndkl_partial_kernel: the warps of a CTA take the SAME sub-slice of different queries
constexpr double kLn2 = 0.693147180559945309417232121458176568;

// ---- tables ------------------------------------------------------------------------------
// Prefix sums in a fixed blocking (chunks of kTabChunk entries, chunk totals added up in index
// order), so that an entry's value depends only on its index -- not on the table length or on
// which context built it: sharded and unsharded runs then agree bit for bit.
constexpr int kTabChunk = 256;

__device__ __forceinline__ void table_terms(int64_t i, double* w, double* wl, double* wi) {
  const double ww = kLn2 / log((double)(i + 1));
  *w = ww;
  *wl = ww * log((double)i);
  *wi = ww / (double)i;
}

__global__ void __launch_bounds__(256) table_chunk_sums_kernel(double* __restrict__ chunk_sums, int64_t len,
                                                               int64_t chunks) {
  const int64_t ch = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= chunks) return;
  const int64_t lo = ch * kTabChunk;
  const int64_t hi = (lo + kTabChunk < len) ? lo + kTabChunk : len;
  double sw = 0.0, swl = 0.0, swi = 0.0;
  for (int64_t i = (lo > 0 ? lo : 1); i < hi; ++i) {
    double w, wl, wi;
    table_terms(i, &w, &wl, &wi);
    sw += w;
    swl += wl;
    swi += wi;
  }
  chunk_sums[3 * ch] = sw;
  chunk_sums[3 * ch + 1] = swl;
  chunk_sums[3 * ch + 2] = swi;
}

__global__ void table_chunk_scan_kernel(double* __restrict__ chunk_sums, int64_t chunks) {
  const int k = threadIdx.x;          // 3 threads: one running sum each, strictly in index order
  if (k >= 3) return;
  double run = 0.0;
  for (int64_t ch = 0; ch < chunks; ++ch) {
    const double v = chunk_sums[3 * ch + k];
    chunk_sums[3 * ch + k] = run;
    run += v;
  }
}

__global__ void __launch_bounds__(256) table_fill_kernel(const double* __restrict__ chunk_sums,
                                                         double* __restrict__ dxl, double* __restrict__ pwi,
                                                         double* __restrict__ pw, double* __restrict__ pwl,
                                                         int64_t len, int64_t chunks) {
  const int64_t ch = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= chunks) return;
  const int64_t lo = ch * kTabChunk;
  const int64_t hi = (lo + kTabChunk < len) ? lo + kTabChunk : len;
  double sw = chunk_sums[3 * ch], swl = chunk_sums[3 * ch + 1], swi = chunk_sums[3 * ch + 2];
  for (int64_t i = lo; i < hi; ++i) {
    if (i == 0) {
      dxl[0] = 0.0; pwi[0] = 0.0; pw[0] = 0.0; pwl[0] = 0.0;
      continue;
    }
    double w, wl, wi;
    table_terms(i, &w, &wl, &wi);
    sw += w;
    swl += wl;
    swi += wi;
    const double di = (double)i;
    // i ln i - (i-1) ln(i-1) = ln i + (i-1) ln(1 + 1/(i-1)), without the cancellation
    dxl[i] = (i == 1) ? 0.0 : log(di) + (di - 1.0) * log1p(1.0 / (di - 1.0));
    pwi[i] = swi;
    pw[i] = sw;
    pwl[i] = swl;
  }
}

// ---- m_q ---------------------------------------------------------------------------------
// First position of [lo, hi) where `pred` (true on a prefix, false after it) stops holding, searched
// by a whole warp: 32 probes per round cut the range 33-fold (5 dependent loads at N = 10^5 where a
// binary search needs 17; 12.6 -> 5.4 us on config 2).
template <class Pred>
__device__ __forceinline__ long long warp_partition_point(long long lo, long long hi, int lane, Pred pred) {
  while (lo < hi) {
    const long long step = (hi - lo + 32) / 33;
    const long long pos = lo + (long long)(lane + 1) * step - 1;
    const bool t = (pos < hi) && pred(pos);
    const int n_true = __popc(__ballot_sync(0xffffffffu, t));   // monotone: the true probes are a prefix
    const long long nlo = lo + (long long)n_true * step;
    if (n_true < 32) {
      const long long cap = lo + (long long)(n_true + 1) * step - 1;   // this probe failed (or lies past hi)
      if (cap < hi) hi = cap;
    }
    lo = nlo;
  }
  return lo;
}

// one warp per query
__global__ void __launch_bounds__(128) ranked_len_kernel(const uint32_t* __restrict__ keys, int64_t ldk,
                                                         const int32_t* __restrict__ avail, int64_t n_avail,
                                                         int64_t top_n, int q, int32_t* __restrict__ ranked_len,
                                                         const uint32_t* __restrict__ nan_cnt) {
  const int qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (qi >= q) return;
  long long m;
  if (avail != nullptr) {
    m = avail[qi];                       // streaming top-k route: the list already ends at the tie run
  } else if (keys != nullptr) {
    const uint32_t* k = keys + (int64_t)qi * ldk;
    // first NaN key (dropped rows, selectn.py:133-134)
    const long long n_valid = warp_partition_point(0, n_avail, lane, [&](long long i) { return __ldg(k + i) < kNanKey; });
    m = n_valid;
    if (top_n < n_valid) {
      const uint32_t kth = __ldg(k + top_n - 1);
      // first position whose key is > kth: the end of the top_n-th row's tie run (keep="all")
      m = warp_partition_point(top_n, n_valid, lane, [&](long long i) { return __ldg(k + i) <= kth; });
    }
  } else if (nan_cnt != nullptr) {
    m = n_avail - (long long)nan_cnt[qi];   // whole segment ranked (top_n >= n_avail): every row but the NaN ones
  } else {
    m = top_n < n_avail ? top_n : n_avail;
  }
  if (lane == 0) ranked_len[qi] = (int32_t)m;
}

// ---- class counts of each sub-slice ---------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(32 * kWarpsPerCta)
sub_counts_kernel(const uint8_t* __restrict__ labs, int64_t ldl, const int32_t* __restrict__ ranked_len,
                  int subs, int stride, uint32_t* __restrict__ sub_cnt) {
  const int q = blockIdx.y;
  const int sub = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (sub >= subs) return;
  const long long m = ranked_len[q];
  const long long begin = (long long)sub * kSub;
  long long end = begin + kSub;
  if (end > m) end = m;
  // the row pitch is a multiple of 256 and kSub of 16: 16-byte label loads are aligned and stay inside
  // the row's padding; both loads of a lane (the whole slice of the warp) are in flight together
  const uint8_t* lab = labs + (int64_t)q * ldl;
  uint32_t w[8];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const long long r0 = begin + 512 * h + 16 * lane;
    uint4 v = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);   // labels past the end match no class
    if (r0 < end) v = __ldg(reinterpret_cast<const uint4*>(lab + r0));
    w[4 * h] = v.x; w[4 * h + 1] = v.y; w[4 * h + 2] = v.z; w[4 * h + 3] = v.w;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long left = end - (r0 + 4 * k);
      if (left < 4) w[4 * h + k] |= (left <= 0) ? 0xFFFFFFFFu : (0xFFFFFFFFu << (8 * (int)left));
    }
  }
  uint32_t cnt[C];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += __popc(__vcmpeq4(w[k], 0x01010101u * c));
    cnt[c] = acc >> 3;
  }
#pragma unroll
  for (int c = 0; c < C; ++c) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) cnt[c] += __shfl_xor_sync(0xffffffffu, cnt[c], off);
  }
  if (lane == 0) {
    uint32_t* out = sub_cnt + ((int64_t)q * stride + sub) * kMaxClasses;
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c) out[c] = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) out[c] = cnt[c];
  }
}

// exclusive prefix over sub-slices, one thread per (query, class); totals in slot `subs`
__global__ void __launch_bounds__(128) sub_scan_kernel(uint32_t* __restrict__ sub_cnt, int q, int subs) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= q * kMaxClasses) return;
  const int qi = gid / kMaxClasses, c = gid % kMaxClasses;
  uint32_t* base = sub_cnt + (int64_t)qi * (subs + 1) * kMaxClasses + c;
  uint32_t run = 0;
  int s = 0;
  for (; s + 8 <= subs; s += 8) {                 // eight loads in flight: the kernel is pure latency
    uint32_t v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = base[(int64_t)(s + j) * kMaxClasses];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      base[(int64_t)(s + j) * kMaxClasses] = run;
      run += v[j];
    }
  }
  for (; s < subs; ++s) {
    const uint32_t v = base[(int64_t)s * kMaxClasses];
    base[(int64_t)s * kMaxClasses] = run;
    run += v;
  }
  base[(int64_t)subs * kMaxClasses] = run;
}

struct PartialParams {
  const uint8_t* labs;
  int64_t ldl;
  const int32_t* ranked_len;
  int subs;
  int qn;                                 // queries of this launch
  int64_t n_gallery;
  const uint32_t* sub_pref;               // [q][subs+1][16] exclusive prefixes (+ totals)
  const unsigned long long* class_counts; // global n_c
  const double* dxl;
  const double* pwi;
  double* partial;                        // [q][subs][2]
};

// One warp per 1024-row sub-slice, eight rounds of 128 rows: a lane owns four consecutive ranked rows
// (one 32-bit label load).  The class counts of a lane's rows are packed as 8-bit fields of 64-bit
// words (classes 0-7, 8-15); ONE inclusive warp scan of those words gives every lane the number of
// rows of each class in the lanes before it (a round holds 128 rows, a field at most 128), so the
// own-class prefix count of a row -- the only thing NDKL needs from the ranking -- costs a field
// extract instead of a ballot per class.  The running per-class base of the warp lives in shared
// memory (it is indexed by the label).
// The warps of a CTA work on the same sub-slice of kQueriesPerCta consecutive queries: their PWI rows are
// the same addresses and their DXL gathers (t ~ rank x class share) cluster, so both tables are served
// from L1 instead of once per warp from L2.
template <int C>
__global__ void __launch_bounds__(32 * kQueriesPerCta) ndkl_partial_kernel(const PartialParams p) {
  constexpr int W = (C + 7) / 8;
  __shared__ double ln_d[kMaxClasses];
  __shared__ uint32_t s_base[kQueriesPerCta][kMaxClasses];
  if (threadIdx.x < kMaxClasses) {
    double v = 0.0;
    if (threadIdx.x < C) {
      const unsigned long long nc = p.class_counts[threadIdx.x];
      // dem_par: max(count, 1) / N   (measuring_bias.py:27)
      v = log((double)(nc > 0 ? nc : 1ull) / (double)p.n_gallery);
    }
    ln_d[threadIdx.x] = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  const int q = blockIdx.y * kQueriesPerCta + warp;
  const int sub = blockIdx.x;
  const int lane = threadIdx.x & 31;
  if (q >= p.qn) return;
  const long long m = p.ranked_len[q];
  const long long begin = (long long)sub * kSub;
  long long end = begin + kSub;
  if (end > m) end = m;
  double acc_eq = 0.0, acc_d = 0.0;
  if (begin < end) {
    const uint8_t* lab = p.labs + (int64_t)q * p.ldl;
    const uint32_t* pref = p.sub_pref + ((int64_t)q * (p.subs + 1) + sub) * kMaxClasses;
    uint32_t* base = s_base[warp];
    if (lane < kMaxClasses) base[lane] = pref[lane];
    __syncwarp();
    const double pwi_m = __ldg(p.pwi + m);
    // One round = 128 ranked rows.  Full rounds (all but the last of a query) run without any per-row
    // condition, so the four table gathers of a lane are in flight together; with the `k < valid` tests
    // each gather sat in its own branch region and the round paid four L2 latencies in sequence.
    // a lane's inputs of one round: four labels and the four PWI entries of its rows
    auto load_round = [&](const long long r0, uint32_t& four, double2& pa, double2& pb) {
      const long long r = r0 + 4 * lane;
      four = 0;
      pa = make_double2(0.0, 0.0);
      pb = make_double2(0.0, 0.0);
      if (r < end) {                                 // the label pitch is padded to 256, the tables to len + 1024
        four = __ldg(reinterpret_cast<const uint32_t*>(lab + r));
        pa = __ldg(reinterpret_cast<const double2*>(p.pwi + r));
        pb = __ldg(reinterpret_cast<const double2*>(p.pwi + r + 2));
      }
    };
    auto round_body = [&](const long long r0, const uint32_t four, const double2 pa, const double2 pb, auto full_tag) {
      constexpr bool kFullRound = decltype(full_tag)::value;
      const long long r = r0 + 4 * lane;             // 0-based rank of this lane's first row
      const int valid = kFullRound ? 4 : ((end - r >= 4) ? 4 : (end > r ? (int)(end - r) : 0));
      unsigned long long own[W];
#pragma unroll
      for (int w = 0; w < W; ++w) own[w] = 0ull;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t l = (four >> (8 * k)) & (kMaxClasses - 1);   // labels >= C are rejected by the caller
        if (kFullRound || k < valid) {
          if (W == 1 || l < 8) own[0] += 1ull << (8 * (l & 7));
          else own[W - 1] += 1ull << (8 * (l & 7));
        }
      }
      unsigned long long incl[W];
#pragma unroll
      for (int w = 0; w < W; ++w) {
        incl[w] = own[w];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const unsigned long long o = __shfl_up_sync(0xffffffffu, incl[w], off);
          if (lane >= off) incl[w] += o;
        }
      }
      unsigned long long run[W];                     // rows of each class before this lane's next row
#pragma unroll
      for (int w = 0; w < W; ++w) run[w] = incl[w] - own[w];
      const double pw4[4] = {pa.x, pa.y, pb.x, pb.y};
      if (kFullRound) {
        uint32_t tt[4], ll[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t l = (four >> (8 * k)) & (kMaxClasses - 1);
          const int sh = 8 * (l & 7);
          const bool hi = (W > 1) && l >= 8;
          const uint32_t before = (uint32_t)((hi ? run[W - 1] : run[0]) >> sh) & 0xFFu;
          if (hi) run[W - 1] += 1ull << sh; else run[0] += 1ull << sh;
          tt[k] = base[l] + before + 1u;             // this row is the t-th of its class
          ll[k] = l;
        }
        double dx[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) dx[k] = __ldg(p.dxl + tt[k]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const double tj = pwi_m - pw4[k];          // PWI[m] - PWI[j-1], j = r + k + 1
          acc_eq = fma(tj, dx[k], acc_eq);
          acc_d = fma(tj, ln_d[ll[k]], acc_d);
        }
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (k < valid) {
            const uint32_t l = (four >> (8 * k)) & (kMaxClasses - 1);
            const int sh = 8 * (l & 7);
            const bool hi = (W > 1) && l >= 8;
            const uint32_t before = (uint32_t)((hi ? run[W - 1] : run[0]) >> sh) & 0xFFu;
            if (hi) run[W - 1] += 1ull << sh; else run[0] += 1ull << sh;
            const uint32_t t = base[l] + before + 1u;
            const double tj = pwi_m - pw4[k];
            acc_eq = fma(tj, __ldg(p.dxl + t), acc_eq);
            acc_d = fma(tj, ln_d[l], acc_d);
          }
        }
      }
      __syncwarp();
      {                                              // lane 31's inclusive scan = the round's class totals
        unsigned long long tot = __shfl_sync(0xffffffffu, incl[0], 31);
        if (W > 1) {
          const unsigned long long tot_hi = __shfl_sync(0xffffffffu, incl[W - 1], 31);
          if (lane >= 8) tot = tot_hi;
        }
        if (lane < C) base[lane] += (uint32_t)(tot >> (8 * (lane & 7))) & 0xFFu;
      }
      __syncwarp();
    };
    // the next round's inputs are requested before the current round is processed
    uint32_t four_next;
    double2 pa_next, pb_next;
    load_round(begin, four_next, pa_next, pb_next);
    for (long long r0 = begin; r0 < end; r0 += 128) {
      const uint32_t four = four_next;
      const double2 pa = pa_next, pb = pb_next;
      if (r0 + 128 < end) load_round(r0 + 128, four_next, pa_next, pb_next);
      if (r0 + 128 <= end) round_body(r0, four, pa, pb, std::true_type{});
      else round_body(r0, four, pa, pb, std::false_type{});
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      acc_eq += __shfl_xor_sync(0xffffffffu, acc_eq, off);
      acc_d += __shfl_xor_sync(0xffffffffu, acc_d, off);
    }
  }
  if (lane == 0) {
    double* out = p.partial + ((int64_t)q * p.subs + sub) * 2;
    out[0] = acc_eq;
    out[1] = acc_eq - acc_d;
  }
}

template <int C>
void launch_counts_and_partials(const PartialParams& p, dim3 grid, int stride, uint32_t* sub_cnt_q, bool counts,
                                cudaStream_t st) {
  if (counts)
    sub_counts_kernel<C><<<grid, 32 * kWarpsPerCta, 0, st>>>(p.labs, p.ldl, p.ranked_len, p.subs, stride, sub_cnt_q);
  else
    ndkl_partial_kernel<C><<<dim3((unsigned)p.subs, (unsigned)((p.qn + kQueriesPerCta - 1) / kQueriesPerCta)),
                             32 * kQueriesPerCta, 0, st>>>(p);
}

void dispatch_classes(int n_classes, const PartialParams& p, dim3 grid, int stride, uint32_t* sub_cnt_q, bool counts,
                      cudaStream_t st) {
  switch (n_classes) {
#define DVL_CASE(C) case C: launch_counts_and_partials<C>(p, grid, stride, sub_cnt_q, counts, st); break;
    DVL_CASE(1) DVL_CASE(2) DVL_CASE(3) DVL_CASE(4) DVL_CASE(5) DVL_CASE(6) DVL_CASE(7) DVL_CASE(8)
    DVL_CASE(9) DVL_CASE(10) DVL_CASE(11) DVL_CASE(12) DVL_CASE(13) DVL_CASE(14) DVL_CASE(15) DVL_CASE(16)
#undef DVL_CASE
    default: break;
  }
}

struct FinalParams {
  const int32_t* ranked_len;
  int q;
  int n_classes;
  int subs;
  int64_t n_gallery, top_n;
  const uint32_t* sub_pref;
  const unsigned long long* class_counts;
  const double* partial;
  const double* pw;
  const double* pwl;
  double* per_query;
  int sum_counts;       // sub_pref holds per-slice COUNTS (fused bucket sort): total them into slot `subs` first
};

__global__ void __launch_bounds__(128) finalize_kernel(const FinalParams p) {
  // one warp per query: lanes sum the sub-slice partials in a fixed order (stride 32, then a
  // butterfly), lane 0 finishes NDKL and computes MaxSkew
  const int qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (qi >= p.q) return;
  const long long m = p.ranked_len[qi];
  double s_eq = 0.0, s_dp = 0.0;
  const double2* part = reinterpret_cast<const double2*>(p.partial) + (int64_t)qi * p.subs;
  for (int s = lane; s < p.subs; s += 32) {
    const double2 v = __ldg(part + s);
    s_eq += v.x;
    s_dp += v.y;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    s_eq += __shfl_xor_sync(0xffffffffu, s_eq, off);
    s_dp += __shfl_xor_sync(0xffffffffu, s_dp, off);
  }
  if (p.sum_counts) {
    uint32_t* cnt = const_cast<uint32_t*>(p.sub_pref) + (int64_t)qi * (p.subs + 1) * kMaxClasses;
    if (lane < kMaxClasses) {
      uint32_t tot = 0;
      for (int s = 0; s < p.subs; ++s) tot += cnt[(int64_t)s * kMaxClasses + lane];
      cnt[(int64_t)p.subs * kMaxClasses + lane] = tot;
    }
    __syncwarp();
  }
  if (lane != 0) return;
  int c_present = 0;
  for (int c = 0; c < p.n_classes; ++c) c_present += (p.class_counts[c] > 0);
  const double z = p.pw[p.top_n];
  const double ndkl_eq = (s_eq + log((double)c_present) * p.pw[m] - p.pwl[m]) / z;
  const double ndkl_dp = (s_dp - p.pwl[m]) / z;   // s_dp = sum_j T_j (DXL[t_j] - ln d_{l_j})
  // MaxSkew over the m ranked rows (measuring_bias.py:58-70)
  const uint32_t* tot = p.sub_pref + ((int64_t)qi * (p.subs + 1) + p.subs) * kMaxClasses;
  double ms_eq = 0.0, ms_dp = 0.0;
  for (int c = 0; c < p.n_classes; ++c) {
    const unsigned long long nc = p.class_counts[c];
    if (nc == 0) continue;                       // Counter(df.label) has no such key
    double pp = (double)tot[c] / (double)p.top_n;
    if (pp == 0.0) pp = 1.0 / (double)p.top_n;   // "no log of 0" (:64-65)
    const double lp = log(pp);
    ms_eq = fmax(ms_eq, lp - log(1.0 / (double)c_present));
    ms_dp = fmax(ms_dp, lp - log((double)nc / (double)p.n_gallery));
  }
  double* out = p.per_query + (int64_t)qi * 4;
  out[DVL_MAXSKEW_EQ_OPP] = ms_eq;
  out[DVL_MAXSKEW_DEM_PAR] = ms_dp;
  out[DVL_NDKL_EQ_OPP] = (m > 0) ? ndkl_eq : 0.0;
  out[DVL_NDKL_DEM_PAR] = (m > 0) ? ndkl_dp : 0.0;
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

int ensure_metric_tables(Ctx* c, int64_t len, cudaStream_t st) {
  // ndkl_partial_kernel reads PWI in aligned groups of four: the tables keep 4 entries of slack and
  // a length that is a multiple of 4 (the four tables sit back to back)
  if (len + 4 <= c->log_tab_len) return DVL_OK;
  const int64_t want = (len + len / 8 + 1024 + 3) / 4 * 4;
  const int64_t chunks = (want + kTabChunk - 1) / kTabChunk;
  DVL_TRY(ensure(c, c->log_tab, (size_t)want * 4 * sizeof(double)));
  DVL_TRY(ensure(c, c->tab_chunks, (size_t)chunks * 3 * sizeof(double)));
  double* base = (double*)c->log_tab.ptr;
  double* sums = (double*)c->tab_chunks.ptr;
  const unsigned blocks = (unsigned)((chunks + 255) / 256);
  table_chunk_sums_kernel<<<blocks, 256, 0, st>>>(sums, want, chunks);
  DVL_LAUNCH_CHECK(c, "table_chunk_sums_kernel");
  table_chunk_scan_kernel<<<1, 32, 0, st>>>(sums, chunks);
  DVL_LAUNCH_CHECK(c, "table_chunk_scan_kernel");
  table_fill_kernel<<<blocks, 256, 0, st>>>(sums, base, base + want, base + 2 * want, base + 3 * want, want, chunks);
  DVL_LAUNCH_CHECK(c, "table_fill_kernel");
  c->log_tab_len = want;
  return DVL_OK;
}

int launch_metrics(Ctx* c, const MetricArgs& a, cudaStream_t st) {
  if (a.n_classes < 1 || a.n_classes > kMaxClasses) return set_error(c, DVL_ERR_INVALID, "n_classes out of range");
  if (a.top_n < 1) return set_error(c, DVL_ERR_INVALID, "top_n must be >= 1");
  const int64_t tab_need = (a.n_avail > a.top_n ? a.n_avail : a.top_n) + 2;
  if (tab_need + 4 > c->log_tab_len) return set_error(c, DVL_ERR_INVALID, "metric tables too short");
  const int64_t tab = c->log_tab_len;
  const double* dxl = (const double*)c->log_tab.ptr;
  const double* pwi = dxl + tab;
  const double* pw = dxl + 2 * tab;
  const double* pwl = dxl + 3 * tab;
  // boundary ties may extend the ranked prefix to every available row
  const int subs = a.n_avail > kSub ? (int)((a.n_avail + kSub - 1) / kSub) : 1;
  DVL_TRY(ensure(c, c->sub_cnt, (size_t)a.q * (subs + 1) * kMaxClasses * sizeof(uint32_t)));
  DVL_TRY(ensure(c, c->partial, (size_t)a.q * subs * 2 * sizeof(double)));
  uint32_t* sub_cnt = (uint32_t*)c->sub_cnt.ptr;
  double* partial = (double*)c->partial.ptr;

  ranked_len_kernel<<<(a.q + 3) / 4, 128, 0, st>>>(a.keys, a.ldk, a.avail, a.n_avail, a.top_n, a.q, a.ranked_len,
                                                   a.nan_cnt);
  DVL_LAUNCH_CHECK(c, "ranked_len_kernel");
  const unsigned bx = (unsigned)((subs + kWarpsPerCta - 1) / kWarpsPerCta);
  for (int phase = 0; phase < 2; ++phase) {
    for (int q0 = 0; q0 < a.q; q0 += 65535) {
      const int qn = (a.q - q0) < 65535 ? (a.q - q0) : 65535;
      dim3 grid(bx, (unsigned)qn);
      PartialParams p;
      p.labs = a.labs + (int64_t)q0 * a.ldl;
      p.ldl = a.ldl;
      p.ranked_len = a.ranked_len + q0;
      p.subs = subs;
      p.qn = qn;
      p.n_gallery = a.n_gallery;
      p.sub_pref = sub_cnt + (int64_t)q0 * (subs + 1) * kMaxClasses;
      p.class_counts = a.class_counts;
      p.dxl = dxl;
      p.pwi = pwi;
      p.partial = partial + (int64_t)q0 * subs * 2;
      // counts are written with a (subs + 1) stride so that the scan can append the totals
      dispatch_classes(a.n_classes, p, grid, subs + 1, sub_cnt + (int64_t)q0 * (subs + 1) * kMaxClasses, phase == 0,
                       st);
      DVL_LAUNCH_CHECK(c, phase == 0 ? "sub_counts_kernel" : "ndkl_partial_kernel");
    }
    if (phase == 0) {
      sub_scan_kernel<<<(a.q * kMaxClasses + 127) / 128, 128, 0, st>>>(sub_cnt, a.q, subs);
      DVL_LAUNCH_CHECK(c, "sub_scan_kernel");
    }
  }
  FinalParams f;
  f.ranked_len = a.ranked_len;
  f.q = a.q;
  f.n_classes = a.n_classes;
  f.subs = subs;
  f.n_gallery = a.n_gallery;
  f.top_n = a.top_n;
  f.sub_pref = sub_cnt;
  f.class_counts = a.class_counts;
  f.partial = partial;
  f.pw = pw;
  f.pwl = pwl;
  f.per_query = a.per_query;
  f.sum_counts = 0;
  finalize_kernel<<<(a.q + 3) / 4, 128, 0, st>>>(f);
  DVL_LAUNCH_CHECK(c, "finalize_kernel");
  return DVL_OK;
}

void metric_tables(const Ctx* c, const double** dxl, const double** pwi) {
  *dxl = (const double*)c->log_tab.ptr;
  *pwi = *dxl + c->log_tab_len;
}

int fused_metric_buffers(Ctx* c, int q, int buckets, double** partial, uint32_t** ranked_cls) {
  DVL_TRY(ensure(c, c->sub_cnt, (size_t)q * (buckets + 1) * kMaxClasses * sizeof(uint32_t)));
  DVL_TRY(ensure(c, c->partial, (size_t)q * buckets * 2 * sizeof(double)));
  *partial = (double*)c->partial.ptr;
  *ranked_cls = (uint32_t*)c->sub_cnt.ptr;
  return DVL_OK;
}

int launch_metrics_fused(Ctx* c, const MetricArgs& a, int buckets, cudaStream_t st) {
  if (a.nan_cnt == nullptr || a.top_n < a.n_avail) return set_error(c, DVL_ERR_INVALID, "fused metrics rank whole segments");
  const int64_t tab = c->log_tab_len;
  if (a.top_n + 6 > tab || a.n_avail + 6 > tab) return set_error(c, DVL_ERR_INVALID, "metric tables too short");
  const double* dxl = (const double*)c->log_tab.ptr;
  ranked_len_kernel<<<(a.q + 3) / 4, 128, 0, st>>>(nullptr, 0, nullptr, a.n_avail, a.top_n, a.q, a.ranked_len, a.nan_cnt);
  DVL_LAUNCH_CHECK(c, "ranked_len_kernel");
  FinalParams f;
  f.ranked_len = a.ranked_len;
  f.q = a.q;
  f.n_classes = a.n_classes;
  f.subs = buckets;
  f.n_gallery = a.n_gallery;
  f.top_n = a.top_n;
  f.sub_pref = (const uint32_t*)c->sub_cnt.ptr;
  f.class_counts = a.class_counts;
  f.partial = (const double*)c->partial.ptr;
  f.pw = dxl + 2 * tab;
  f.pwl = dxl + 3 * tab;
  f.per_query = a.per_query;
  f.sum_counts = 1;
  finalize_kernel<<<(a.q + 3) / 4, 128, 0, st>>>(f);
  DVL_LAUNCH_CHECK(c, "finalize_kernel");
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
