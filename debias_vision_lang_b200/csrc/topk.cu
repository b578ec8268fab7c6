// Ranking stage, streaming top-k route (SURVEY.md row a3 for top_n << N; the north_star's
// kernel (b)).  The similarity GEMM's epilogue (gemm.cu, SEED / FILTER modes) appends to each
// query's candidate list every gallery row whose score is >= the query's running threshold
// tau_q.  This file holds the other half:
//
//   select_kernel, one CTA per query
//     * radix-select (4 x 8-bit passes over the list, shared-memory histograms) of the top_n-th
//       smallest ranking key  ->  new threshold;
//     * keeps every entry with key <= that key -- i.e. the top_n best AND every row tied with
//       the top_n-th (pandas nlargest keep="all", selectn.py:176-182) -- compacted through
//       shared memory, so tau_q only ever tightens and nothing that can still rank is dropped;
//     * final pass: bitonic sort of the kept entries in shared memory.  An entry is
//       key<<32 | global index<<4 | label, so ascending u64 order IS the reference's order
//       (score descending, lowest index first) no matter in which order the GEMM CTAs appended;
//       emits the ranked labels (for metrics.cu) and indices.
//   A query whose tie run does not fit the list (e.g. an all-zero query: every row ties) is
//   flagged and re-ranked by the full-sort route, so the result is exact for any input.
//
// The same kernel merges the per-shard lists of a multi-GPU run (segments = shards).
//
// Measured and not kept (config 4 shard, 10^4 queries x ~2.8 k candidates, final select 0.55 ms):
// skipping the radix-select passes over the key bytes all candidates of a query share (min/max of the
// keys first) and reserving compaction slots once per warp instead of once per entry -- 0.72 -> 0.83 ms
// for the three select launches, i.e. no gain: the kernel is bound by the instructions of the 55-stage
// bitonic sort of 1024 u64 entries per query, not by shared-memory atomics.
#include "dvl_common.cuh"
#include "dvl_internal.h"

namespace dvl {

namespace {

constexpr int kSelThreads = 256;

struct SelectParams {
  // candidate lists: query q's entries are `segs` runs of up to seg_len entries,
  // run s starting at cand + s*seg_stride + q*row_stride (local lists: segs = 1)
  unsigned long long* cand;
  int64_t row_stride, seg_stride;
  int segs;
  int64_t seg_len;          // entries per segment when segs > 1 (fixed); unused for segs == 1
  uint32_t* cnt;            // [q] entries in use (segs == 1); null for segs > 1
  int64_t cap;              // capacity of a local list
  float* tau;               // [q] threshold to update, or null
  uint32_t* flags;          // [q] tie-overflow flags
  int64_t k;
  int sort_cap;             // S: entries that fit the shared-memory list
  int64_t trigger;          // skip (non-final) while the list is shorter than this
  int final_pass;
  int pilot_pass;
  int rank_sort;            // final pass: the launch carries the shared memory of counting_rank_sort (1), and
                            // selection + ordering run fused (2)
  int64_t rows_total;
  // final outputs
  unsigned long long* sorted_out;   // [q][sort_cap] or null (defaults to in place for segs == 1)
  uint8_t* labs;            // [q][ldl]
  int64_t ldl;
  int32_t* idx_out;         // [q][idx_cap] or null
  int64_t idx_cap;
  int32_t* avail;           // [q] ranked rows m_q
  int q_offset;             // queries handled: blockIdx.x + q_offset
  const uint8_t* labels;    // local lists: labels of this shard, filled into the entries on the final pass
  int64_t index_base;       //              (global index of local row 0); null when entries carry labels
};

__device__ __forceinline__ unsigned long long load_entry(const SelectParams& p, int q, int64_t i) {
  if (p.segs == 1) return p.cand[(int64_t)q * p.row_stride + i];
  const int64_t s = i / p.seg_len;
  return p.cand[s * p.seg_stride + (int64_t)q * p.row_stride + (i - s * p.seg_len)];
}

// Bitonic sort of N = 256 * E u64 entries held E per thread (entry i = tid * E + j), ascending.
// Exchanges inside a thread are register swaps, inside a warp 64-bit shuffles; only the strides that
// cross warps (>= 32 * E) go through shared memory: 6 of the 55 stages at N = 1024 -- the other 49
// need no barrier and touch no shared memory.  `buf` holds the entries on entry and on return.
template <int E>
__device__ __forceinline__ void bitonic_sort_regs(unsigned long long* buf, int tid) {
  constexpr int N = kSelThreads * E;
  const int lane = tid & 31;
  unsigned long long v[E];
#pragma unroll
  for (int j = 0; j < E; ++j) v[j] = buf[tid * E + j];
  for (int size = 2; size <= N; size <<= 1) {
    for (int stride = size >> 1; stride >= E; stride >>= 1) {
      const int tstride = stride / E;                       // partner thread = tid ^ tstride, same j
      const bool lower = (tid & tstride) == 0;
      const bool asc = ((tid * E) & size) == 0;             // size > stride >= E: the bit lies in the thread part
      const bool keep_min = lower == asc;
      if (tstride >= 32) {
        __syncthreads();                                    // earlier readers of buf are done
#pragma unroll
        for (int j = 0; j < E; ++j) buf[j * kSelThreads + tid] = v[j];   // j-major: conflict-free
        __syncthreads();
#pragma unroll
        for (int j = 0; j < E; ++j) {
          const unsigned long long o = buf[j * kSelThreads + (tid ^ tstride)];
          v[j] = keep_min ? (o < v[j] ? o : v[j]) : (o > v[j] ? o : v[j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < E; ++j) {
          const unsigned long long o = __shfl_xor_sync(0xffffffffu, v[j], tstride);
          v[j] = keep_min ? (o < v[j] ? o : v[j]) : (o > v[j] ? o : v[j]);
        }
      }
    }
    // strides below E: both entries live in this thread
#pragma unroll
    for (int s = E / 2; s >= 1; s >>= 1) {
      if (s <= (size >> 1)) {
#pragma unroll
        for (int j = 0; j < E; ++j) {
          if ((j & s) == 0) {
            const bool asc = ((tid * E + j) & size) == 0;
            const unsigned long long a = v[j], b = v[j | s];
            if ((a > b) == asc) {
              v[j] = b;
              v[j | s] = a;
            }
          }
        }
      }
    }
  }
  (void)lane;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < E; ++j) buf[tid * E + j] = v[j];
  __syncthreads();
}

// ---- final ordering of a kept list without the bitonic network ----
// The kept entries of a query are a few thousand scores from one smooth tail.  Like the second level of
// bucketsort.cu: a score-linear split of [best, worst] into kSelBins sub-bins (counting sort, ~1 entry per
// sub-bin), then every entry ranks itself among the entries of its own sub-bin by its full (key, index) value.
// ~1/3 of the instructions of the 55..66-stage bitonic sort of 1024..2048 u64 entries.  Exact for the same
// reason: score -> sub-bin is monotone, so entries of different sub-bins are ordered by the sub-bin, and equal
// keys share one.  Returns false (list untouched) when a sub-bin is crowded -- heavy exact ties, e.g. scores
// on a coarse grid -- and the caller sorts bitonically as before.
constexpr int kSelBins = 2048;               // 8 consecutive sub-bins per thread
constexpr int kSelGroupMax = 48;

__device__ __forceinline__ float sel_key_score(uint32_t k) {   // score of a non-NaN key
  return __uint_as_float(k ^ (~(uint32_t)((int32_t)k >> 31) & 0x7FFFFFFFu));
}

__device__ __forceinline__ uint32_t sel_subbin(uint32_t k, float hi, float scale) {
  // (hi - s) * scale truncated with saturation (negative, NaN -> 0), capped; NaN keys (dropped rows) go last
  const uint32_t sb = min(__float2uint_rz((hi - sel_key_score(k)) * scale), (uint32_t)(kSelBins - 1));
  return k == kNanKey ? (uint32_t)(kSelBins - 1) : sb;
}

__device__ bool counting_rank_sort(unsigned long long* list, uint32_t kept, uint8_t* extra, int sort_cap, int tid) {
  uint32_t* s_key = reinterpret_cast<uint32_t*>(extra);           // [sort_cap]
  uint32_t* s_val = s_key + sort_cap;                             // [sort_cap]
  uint32_t* bins = s_val + sort_cap + 4;                          // [kSelBins]; bins[-1] = 0 (16-byte aligned)
  __shared__ uint32_t red_min[kSelThreads / 32], red_max[kSelThreads / 32], warp_tot[kSelThreads / 32];
  __shared__ uint32_t sh_big;
  const int lane = tid & 31, warp = tid >> 5;
  reinterpret_cast<uint4*>(bins)[2 * tid] = make_uint4(0u, 0u, 0u, 0u);
  reinterpret_cast<uint4*>(bins)[2 * tid + 1] = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0) { sh_big = 0u; bins[-1] = 0u; }
  uint32_t kmin = 0xFFFFFFFFu, kmax = 0u;                         // over the keys of real scores
  for (uint32_t i = tid; i < kept; i += kSelThreads) {
    const uint32_t k = (uint32_t)(list[i] >> 32);
    kmin = min(kmin, k);
    kmax = max(kmax, k == kNanKey ? 0u : k);
  }
  kmin = __reduce_min_sync(0xffffffffu, kmin);
  kmax = __reduce_max_sync(0xffffffffu, kmax);
  if (lane == 0) { red_min[warp] = kmin; red_max[warp] = kmax; }
  __syncthreads();
#pragma unroll
  for (int w = 0; w < kSelThreads / 32; ++w) {
    kmin = min(kmin, red_min[w]);
    kmax = max(kmax, red_max[w]);
  }
  float hi = 0.0f, scale = 0.0f;
  if (kmax > kmin) {                                              // two distinct real scores; +-inf ends clamped
    hi = fminf(sel_key_score(kmin), 3.4028234664e38f);
    const float lo = fmaxf(sel_key_score(kmax), -3.4028234664e38f);
    if (hi > lo) scale = (float)kSelBins / (hi - lo);
  }
  if (!(scale <= 3.4028234664e38f)) scale = 0.0f;
  for (uint32_t i = tid; i < kept; i += kSelThreads) {
    const uint32_t k = (uint32_t)(list[i] >> 32);
    atomicAdd(&bins[sel_subbin(k, hi, scale)], 1u);               // histogram: no value returned
  }
  __syncthreads();
  {
    // exclusive prefix over the sub-bins: bins[b] = first slot of sub-bin b (the scatter advances it to its end)
    const uint4 v0 = reinterpret_cast<const uint4*>(bins)[2 * tid];
    const uint4 v1 = reinterpret_cast<const uint4*>(bins)[2 * tid + 1];
    const uint32_t cnt8[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    uint32_t start[8];
    uint32_t sum = 0, mx = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      start[j] = sum;
      sum += cnt8[j];
      mx = max(mx, cnt8[j]);
    }
    uint32_t incl = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += o;
    }
    if (lane == 31) warp_tot[warp] = incl;
    if (mx > (uint32_t)kSelGroupMax) sh_big = 1u;
    __syncthreads();
    if (sh_big != 0u) return false;                               // uniform: every thread reads it after the barrier
    uint32_t base = incl - sum;
#pragma unroll
    for (int w = 0; w < kSelThreads / 32; ++w) base += (w < warp) ? warp_tot[w] : 0u;
#pragma unroll
    for (int j = 0; j < 8; ++j) start[j] += base;
    reinterpret_cast<uint4*>(bins)[2 * tid] = make_uint4(start[0], start[1], start[2], start[3]);
    reinterpret_cast<uint4*>(bins)[2 * tid + 1] = make_uint4(start[4], start[5], start[6], start[7]);
  }
  __syncthreads();
  for (uint32_t i = tid; i < kept; i += kSelThreads) {
    const unsigned long long e = list[i];
    const uint32_t k = (uint32_t)(e >> 32);
    const uint32_t dst = atomicAdd(&bins[sel_subbin(k, hi, scale)], 1u);   // next free slot of the sub-bin
    s_key[dst] = k;
    s_val[dst] = (uint32_t)e;
  }
  __syncthreads();
  for (uint32_t d = tid; d < kept; d += kSelThreads) {
    const uint32_t k = s_key[d], v = s_val[d];
    const uint32_t sb = sel_subbin(k, hi, scale);
    const uint32_t s0 = bins[(int)sb - 1], end = bins[sb];        // after the scatter bins[b] is the END of sub-bin b
    const uint32_t g = end - s0, last = end - 1u;
    uint32_t rank = s0;
#pragma unroll
    for (uint32_t j = 0; j < 4u; ++j) {                           // the first four candidates without a loop
      const uint32_t at = min(s0 + j, last);
      const uint32_t kk = s_key[at];
      bool first = kk < k;
      if (kk == k) first = s_val[at] < v;                         // itself, or an exact tie: the gallery index decides
      rank += (j < g && first) ? 1u : 0u;
    }
    for (uint32_t j = 4u; j < g; ++j) {
      const uint32_t kk = s_key[s0 + j];
      rank += (kk < k || (kk == k && s_val[s0 + j] < v)) ? 1u : 0u;
    }
    list[rank] = ((unsigned long long)k << 32) | v;
  }
  __syncthreads();
  return true;
}

// ---- final pass in one go: selection and ordering share the sub-bin histogram ----
// The 4-pass radix select, the compaction and counting_rank_sort each walk the candidate list; on a list of
// ~2 top_n candidates they were 30 % / 10 % / 25 % of the final select.  Here the score-linear histogram over ALL n
// candidates does the selection as well: the sub-bin b* in which the cumulative count crosses top_n holds the
// top_n-th entry, every candidate of a sub-bin <= b* is scattered and ranked (top_n + the few extra entries of b*),
// the rest is never touched again; the top_n-th key is then simply the key at position top_n - 1 of the ordered
// list, and the boundary tie run ends where the key changes.  Same exactness argument as counting_rank_sort.
// Returns the number of kept entries, ordered at list[0..kept) (NaN keys and padding last), *kth_out = the top_n-th
// key (0xFFFFFFFF if n <= top_n); or -1 (nothing written outside shared memory) when a sub-bin <= b* is crowded
// or the kept entries do not fit -- the caller then takes the radix-select path.
__device__ int fused_select_sort(const SelectParams& p, int q, int n, unsigned long long* list, uint8_t* extra, int tid,
                                 uint32_t* kth_out) {
  const int sort_cap = p.sort_cap;
  uint32_t* sm_keys = reinterpret_cast<uint32_t*>(list);          // [n <= 2 sort_cap] the candidates' keys
  uint32_t* s_key = reinterpret_cast<uint32_t*>(extra);           // [sort_cap]
  uint32_t* s_val = s_key + sort_cap;                             // [sort_cap]
  uint32_t* bins = s_val + sort_cap + 4;                          // [kSelBins]; bins[-1] = 0
  __shared__ uint32_t red_min[kSelThreads / 32], red_max[kSelThreads / 32], warp_tot[kSelThreads / 32];
  __shared__ uint32_t sh_big, sh_bstar, sh_kprime, sh_end;
  const int lane = tid & 31, warp = tid >> 5;
  *kth_out = 0xFFFFFFFFu;
  if (n <= 0) return 0;
  const uint32_t k = (uint32_t)(p.k < (int64_t)n ? p.k : (int64_t)n);   // >= 1
  reinterpret_cast<uint4*>(bins)[2 * tid] = make_uint4(0u, 0u, 0u, 0u);
  reinterpret_cast<uint4*>(bins)[2 * tid + 1] = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0) { sh_big = 0u; bins[-1] = 0u; sh_bstar = (uint32_t)(kSelBins - 1); sh_kprime = (uint32_t)n; }
  uint32_t kmin = 0xFFFFFFFFu, kmax = 0u;                         // over the keys of real scores
  for (int i = tid; i < n; i += kSelThreads) {
    const uint32_t key = (uint32_t)(load_entry(p, q, i) >> 32);
    sm_keys[i] = key;
    kmin = min(kmin, key);
    kmax = max(kmax, key == kNanKey ? 0u : key);
  }
  kmin = __reduce_min_sync(0xffffffffu, kmin);
  kmax = __reduce_max_sync(0xffffffffu, kmax);
  if (lane == 0) { red_min[warp] = kmin; red_max[warp] = kmax; }
  __syncthreads();
#pragma unroll
  for (int w = 0; w < kSelThreads / 32; ++w) {
    kmin = min(kmin, red_min[w]);
    kmax = max(kmax, red_max[w]);
  }
  float hi = 0.0f, scale = 0.0f;
  if (kmax > kmin) {                                              // two distinct real scores; +-inf ends clamped
    hi = fminf(sel_key_score(kmin), 3.4028234664e38f);
    const float lo = fmaxf(sel_key_score(kmax), -3.4028234664e38f);
    // real scores end in sub-bin kSelBins - 2: the last one is left to the NaN keys -- the padding of merged lists,
    // hundreds of entries.  When the lists hold just top_n real entries the top_n-th is the worst real score, and
    // sharing its sub-bin with the padding would make that sub-bin "crowded" for every query.
    if (hi > lo) scale = (float)(kSelBins - 2) / (hi - lo);
  }
  if (!(scale <= 3.4028234664e38f)) scale = 0.0f;
  for (int i = tid; i < n; i += kSelThreads) atomicAdd(&bins[sel_subbin(sm_keys[i], hi, scale)], 1u);
  __syncthreads();
  uint32_t cnt8[8], start[8];
  {
    const uint4 v0 = reinterpret_cast<const uint4*>(bins)[2 * tid];
    const uint4 v1 = reinterpret_cast<const uint4*>(bins)[2 * tid + 1];
    cnt8[0] = v0.x; cnt8[1] = v0.y; cnt8[2] = v0.z; cnt8[3] = v0.w;
    cnt8[4] = v1.x; cnt8[5] = v1.y; cnt8[6] = v1.z; cnt8[7] = v1.w;
    uint32_t sum = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      start[j] = sum;
      sum += cnt8[j];
    }
    uint32_t incl = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += o;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    uint32_t base = incl - sum;
#pragma unroll
    for (int w = 0; w < kSelThreads / 32; ++w) base += (w < warp) ? warp_tot[w] : 0u;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      start[j] += base;
      // the sub-bin in which the cumulative count crosses k (exactly one: k >= 1 and k <= n)
      if (start[j] < k && k <= start[j] + cnt8[j]) {
        sh_bstar = (uint32_t)(8 * tid + j);
        sh_kprime = start[j] + cnt8[j];
      }
    }
    reinterpret_cast<uint4*>(bins)[2 * tid] = make_uint4(start[0], start[1], start[2], start[3]);
    reinterpret_cast<uint4*>(bins)[2 * tid + 1] = make_uint4(start[4], start[5], start[6], start[7]);
  }
  __syncthreads();
  const uint32_t bstar = sh_bstar, kprime = sh_kprime;            // kprime = entries of the sub-bins <= b*
  {
    uint32_t mx = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) mx = max(mx, (uint32_t)(8 * tid + j) <= bstar ? cnt8[j] : 0u);
    if (mx > (uint32_t)kSelGroupMax) sh_big = 1u;
  }
  __syncthreads();
  if (sh_big != 0u || kprime > (uint32_t)sort_cap) return -1;     // uniform
  // scatter: every candidate of a sub-bin <= b* takes the next free slot of its sub-bin
  for (int i = tid; i < n; i += kSelThreads) {
    const uint32_t key = sm_keys[i];
    const uint32_t sb = sel_subbin(key, hi, scale);
    if (sb <= bstar) {
      const unsigned long long e = load_entry(p, q, i);           // second read: L1 / L2
      const uint32_t dst = atomicAdd(&bins[sb], 1u);
      s_key[dst] = key;
      s_val[dst] = (uint32_t)e;
    }
  }
  __syncthreads();                                                // sm_keys is dead: `list` takes the ordered entries
  for (uint32_t d = tid; d < kprime; d += kSelThreads) {
    const uint32_t key = s_key[d], v = s_val[d];
    const uint32_t sb = sel_subbin(key, hi, scale);
    const uint32_t s0 = bins[(int)sb - 1], end = bins[sb];        // after the scatter bins[b] is the END of sub-bin b
    const uint32_t g = end - s0, last = end - 1u;
    uint32_t rank = s0;
#pragma unroll
    for (uint32_t j = 0; j < 4u; ++j) {
      const uint32_t at = min(s0 + j, last);
      const uint32_t kk = s_key[at];
      bool first = kk < key;
      if (kk == key) first = s_val[at] < v;
      rank += (j < g && first) ? 1u : 0u;
    }
    for (uint32_t j = 4u; j < g; ++j) {
      const uint32_t kk = s_key[s0 + j];
      rank += (kk < key || (kk == key && s_val[s0 + j] < v)) ? 1u : 0u;
    }
    list[rank] = ((unsigned long long)key << 32) | v;
  }
  if (tid == 0) sh_end = kprime;
  __syncthreads();
  uint32_t kth = 0xFFFFFFFFu;
  if ((int64_t)n > p.k) {
    // the top_n-th entry, and the end of its tie run: positions >= top_n all lie in sub-bin b*
    kth = (uint32_t)(list[p.k - 1] >> 32);
    for (uint32_t t = (uint32_t)p.k + tid; t < kprime; t += kSelThreads)
      if ((uint32_t)(list[t] >> 32) > kth && (uint32_t)(list[t - 1] >> 32) <= kth) sh_end = t;
  }
  __syncthreads();
  *kth_out = kth;
  return (int)sh_end;
}

__global__ void __launch_bounds__(kSelThreads, 5) select_kernel(const SelectParams p) {
  extern __shared__ unsigned long long sm_list[];     // sort_cap entries
  __shared__ uint32_t hist[256];
  __shared__ uint32_t sh_prefix, sh_remaining, sh_kept;
  const int q = blockIdx.x + p.q_offset;
  const int tid = threadIdx.x;
  if ((p.flags[q] & 1u) != 0u) {                       // already handed to the full-sort route
    if (p.final_pass && tid == 0 && p.avail != nullptr) p.avail[q] = 0;
    return;
  }
  int64_t n;
  if (p.cnt != nullptr) {                              // a local list (a merge of one shard has segs == 1, too)
    n = p.cnt[q];
    // More survivors than the list holds, or (final pass after an optimistic threshold) fewer than
    // top_n of them: the list is not a superset of the ranked rows -> exact re-rank of this query.
    const bool overflow = n > p.cap;
    const bool starved = p.final_pass && p.rows_total > 0 && n < (p.k < p.rows_total ? p.k : p.rows_total);
    if (overflow || starved) {
      if (tid == 0) {
        p.flags[q] = 1u;
        p.cnt[q] = 0u;
        if (p.tau != nullptr) p.tau[q] = __int_as_float(0x7f800000);
        if (p.final_pass && p.avail != nullptr) p.avail[q] = 0;
      }
      return;
    }
  } else {
    n = (int64_t)p.segs * p.seg_len;
  }
  if (!p.final_pass && !p.pilot_pass && n < p.trigger) return;

  // ---- the k-th smallest key ----
  uint32_t kth = 0xFFFFFFFFu;
  int fast_kept = -1;
  if (p.final_pass && p.rank_sort == 2 && n <= 2 * (int64_t)p.sort_cap)
    fast_kept = fused_select_sort(p, q, (int)n, sm_list, reinterpret_cast<uint8_t*>(sm_list + p.sort_cap), tid, &kth);
  if (fast_kept < 0) kth = 0xFFFFFFFFu;
  if (fast_kept < 0 && n > p.k) {
    // short lists (the usual case) are cached once as 32-bit keys in the shared-memory list buffer,
    // so the four radix passes do not go back to L2
    uint32_t* sm_keys = reinterpret_cast<uint32_t*>(sm_list);
    const bool cached = n <= 2 * (int64_t)p.sort_cap;
    if (cached)
      for (int64_t i = tid; i < n; i += kSelThreads) sm_keys[i] = (uint32_t)(load_entry(p, q, i) >> 32);
    if (tid == 0) { sh_prefix = 0u; sh_remaining = (uint32_t)p.k; }
    uint32_t known_mask = 0u;
    for (int shift = 24; shift >= 0; shift -= 8) {
      hist[tid] = 0u;
      __syncthreads();
      const uint32_t prefix = sh_prefix;
      if (cached) {
        for (int i = tid; i < (int)n; i += kSelThreads) {
          const uint32_t key = sm_keys[i];
          if ((key & known_mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xFFu], 1u);
        }
      } else {
        for (int64_t i = tid; i < n; i += kSelThreads) {
          const uint32_t key = (uint32_t)(load_entry(p, q, i) >> 32);
          if ((key & known_mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xFFu], 1u);
        }
      }
      __syncthreads();
      if (tid < 32) {
        // warp 0 finds the digit whose cumulative count crosses `remaining`: 8 bins per lane
        const uint32_t remaining = sh_remaining;
        uint32_t local[8];
        uint32_t sum = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          local[j] = hist[tid * 8 + j];
          sum += local[j];
        }
        uint32_t incl = sum;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const uint32_t o = __shfl_up_sync(0xffffffffu, incl, off);
          if (tid >= off) incl += o;
        }
        uint32_t before = incl - sum;             // entries in lower bins
        if (before < remaining && remaining <= incl) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (before < remaining && remaining <= before + local[j]) {
              sh_prefix = prefix | ((uint32_t)(tid * 8 + j) << shift);
              sh_remaining = remaining - before;
            }
            before += local[j];
          }
        }
      }
      known_mask |= 0xFFu << shift;
      __syncthreads();
    }
    kth = sh_prefix;
  }

  if (p.cnt == nullptr && tid < p.segs) {
    // Merge of per-shard lists that were packed to a speculative common length: a list without padding may
    // have been cut short.  The merge is still exact if what was cut cannot rank: the list's last entry must
    // rank strictly behind the k-th of the union.  Otherwise the query is marked "short" (bit 1) and the
    // exchange is repeated with the length the shards reported.
    const unsigned long long last = load_entry(p, q, (int64_t)tid * p.seg_len + p.seg_len - 1);
    if (last != ~0ull && (uint32_t)(last >> 32) <= kth) atomicOr(&p.flags[q], 2u);
  }

  if (p.pilot_pass) {
    // pilot sample: its k-th best score becomes the (optimistic) threshold of the main pass; the
    // sample's own entries are dropped -- the main pass sees those rows again
    if (tid == 0) {
      p.cnt[q] = 0u;
      p.tau[q] = (n > p.k && kth != kNanKey) ? key_to_score(kth) : __int_as_float(0xff800000);
    }
    return;
  }

  // ---- keep everything that still can rank: key <= kth (ties at the boundary included) ----
  if (fast_kept < 0) {
    if (tid == 0) sh_kept = 0u;
    __syncthreads();
    for (int64_t i = tid; i < n; i += kSelThreads) {
      const unsigned long long e = load_entry(p, q, i);
      if ((uint32_t)(e >> 32) <= kth && e != ~0ull) {
        const uint32_t pos = atomicAdd(&sh_kept, 1u);
        if (pos < (uint32_t)p.sort_cap) sm_list[pos] = e;
      }
    }
    __syncthreads();
  }
  const uint32_t kept = fast_kept >= 0 ? (uint32_t)fast_kept : sh_kept;
  if (kept > (uint32_t)p.sort_cap) {
    if (tid == 0) {
      atomicOr(&p.flags[q], 1u);
      if (p.cnt != nullptr) p.cnt[q] = 0u;
      if (p.tau != nullptr) p.tau[q] = __int_as_float(0x7f800000);   // stop collecting for this query
      if (p.final_pass && p.avail != nullptr) p.avail[q] = 0;
    }
    return;
  }

  if (!p.final_pass) {
    unsigned long long* dst = p.cand + (int64_t)q * p.row_stride;
    for (uint32_t i = tid; i < kept; i += kSelThreads) dst[i] = sm_list[i];
    if (tid == 0) {
      p.cnt[q] = kept;
      if (n >= p.k && kth != kNanKey) p.tau[q] = key_to_score(kth);
    }
    return;
  }

  // ---- final: sort, emit ----
  // the exact score of the local top_n-th row: the tightest bound this shard can contribute to the exchange
  // (the running threshold is only a lower bound of it: a pilot sample's optimistic guess, or the k-th score
  // before the last chunks arrived)
  if (tid == 0 && p.tau != nullptr && n > p.k && kth != kNanKey) p.tau[q] = key_to_score(kth);
  uint32_t pow2 = 1;
  while (pow2 < kept) pow2 <<= 1;
  const uint32_t n_sort = pow2 <= 4u * kSelThreads ? 4u * kSelThreads : 8u * kSelThreads;
  if (fast_kept >= 0) {
    // selected and ordered in one go
  } else if (p.rank_sort && counting_rank_sort(sm_list, kept, reinterpret_cast<uint8_t*>(sm_list + p.sort_cap), p.sort_cap, tid)) {
    // ordered by sub-bin + in-bin rank
  } else if (pow2 <= n_sort && (uint32_t)p.sort_cap >= n_sort) {
    // up to 2048 entries (top_n <= ~1000 with its ties, the usual case): registers + warp shuffles
    for (uint32_t i = kept + tid; i < n_sort; i += kSelThreads) sm_list[i] = ~0ull;
    __syncthreads();
    if (n_sort == 4u * kSelThreads) bitonic_sort_regs<4>(sm_list, tid);
    else bitonic_sort_regs<8>(sm_list, tid);
  } else {
    for (uint32_t i = kept + tid; i < pow2; i += kSelThreads) sm_list[i] = ~0ull;
    __syncthreads();
    for (uint32_t size = 2; size <= pow2; size <<= 1) {
      for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
        for (uint32_t t = tid; t < (pow2 >> 1); t += kSelThreads) {
          const uint32_t i = 2 * t - (t & (stride - 1));
          const uint32_t j = i + stride;
          const bool up = (i & size) == 0;
          const unsigned long long a = sm_list[i], b = sm_list[j];
          if ((a > b) == up) {
            sm_list[i] = b;
            sm_list[j] = a;
          }
        }
        __syncthreads();
      }
    }
  }
  // rows with a NaN score are never ranked (selectn.py:133-134): they sort last
  if (tid == 0) sh_kept = kept;
  __syncthreads();
  for (uint32_t i = tid; i < kept; i += kSelThreads) {
    const bool nan_here = (uint32_t)(sm_list[i] >> 32) == kNanKey;
    const bool nan_prev = i > 0 && (uint32_t)(sm_list[i - 1] >> 32) == kNanKey;
    if (nan_here && !nan_prev) sh_kept = i;
  }
  __syncthreads();
  const uint32_t m = sh_kept;
  unsigned long long* sorted = (p.sorted_out != nullptr) ? p.sorted_out + (int64_t)q * p.sort_cap
                                                         : p.cand + (int64_t)q * p.row_stride;
  uint8_t* labs = p.labs + (int64_t)q * p.ldl;
  for (uint32_t i = tid; i < m; i += kSelThreads) {
    unsigned long long e = sm_list[i];
    if (p.labels != nullptr) e |= (unsigned long long)(p.labels[(int64_t)((e >> 4) & 0x0FFFFFFFull) - p.index_base] & 0xFu);
    sorted[i] = e;
    labs[i] = (uint8_t)(e & 0xFull);
    if (p.idx_out != nullptr && (int64_t)i < p.idx_cap)
      p.idx_out[(int64_t)q * p.idx_cap + i] = (int32_t)((e >> 4) & 0x0FFFFFFFull);
  }
  if (tid == 0) {
    if (p.cnt != nullptr) p.cnt[q] = m;
    p.avail[q] = (int32_t)m;
  }
}

__global__ void __launch_bounds__(256) fill_topk_state_kernel(uint32_t* cnt, float* tau, uint32_t* flags, int q,
                                                             uint32_t cnt_value, int reset_all) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= q) return;
  cnt[i] = cnt_value;
  if (reset_all) {
    tau[i] = __int_as_float(0xff800000);   // -inf: everything is a candidate
    flags[i] = 0u;
  }
}

// sorted list -> number of entries that can still rank globally: key <= key(tau_bound)
__global__ void __launch_bounds__(256) trim_count_kernel(const unsigned long long* __restrict__ sorted,
                                                         int64_t row_stride, const int32_t* __restrict__ avail,
                                                         const float* __restrict__ tau_bound, int q,
                                                         int32_t* __restrict__ out_cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= q) return;
  const uint32_t bound = score_to_key(tau_bound[i]);
  const unsigned long long* e = sorted + (int64_t)i * row_stride;
  int lo = 0, hi = avail[i];
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if ((uint32_t)(e[mid] >> 32) <= bound) lo = mid + 1; else hi = mid;
  }
  out_cnt[i] = lo;
}

__global__ void __launch_bounds__(256) pack_lists_kernel(const unsigned long long* __restrict__ sorted,
                                                         int64_t row_stride, const int32_t* __restrict__ cnt,
                                                         int64_t len, unsigned long long* __restrict__ out) {
  const int q = blockIdx.y;
  const int64_t n = cnt[q];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x)
    out[(int64_t)q * len + i] = (i < n) ? sorted[(int64_t)q * row_stride + i] : ~0ull;
}

// ---- gallery-sharded exchange: statistics, trim + pack, finish -------------------------------------
// Per-rank statistics record: [tau f32 x q][overflow flags u32 x q][label counts u64 x 17] = 2q + 34 words.
__global__ void __launch_bounds__(256) pack_stats_kernel(const float* __restrict__ tau, const uint32_t* __restrict__ flags,
                                                         const unsigned long long* __restrict__ counts, int q,
                                                         uint32_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < q) {
    out[i] = __float_as_uint(tau[i]);
    out[q + i] = flags[i];
  }
  if (i < 2 * (kMaxClasses + 1)) out[2 * q + i] = reinterpret_cast<const uint32_t*>(counts)[i];
}

// all ranks' records -> tau bound (max: no row scoring below it can rank globally), flags (or), counts (sum)
__global__ void __launch_bounds__(256) reduce_stats_kernel(const uint32_t* __restrict__ all, int world, int q,
                                                           float* __restrict__ tau_bound, uint32_t* __restrict__ flags,
                                                           unsigned long long* __restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = 2 * (int64_t)q + 2 * (kMaxClasses + 1);
  if (i < q) {
    float t = __int_as_float(0xff800000);
    uint32_t f = 0u;
    for (int w = 0; w < world; ++w) {
      t = fmaxf(t, __uint_as_float(all[w * stride + i]));
      f |= all[w * stride + q + i];
    }
    tau_bound[i] = t;
    flags[i] = f;
  }
  if (i < kMaxClasses + 1) {
    unsigned long long sum = 0ull;
    for (int w = 0; w < world; ++w) sum += reinterpret_cast<const unsigned long long*>(all + w * stride + 2 * q)[i];
    counts[i] = sum;
  }
}

// one CTA per query: entries of the sorted local list that can still rank globally (key <= key(tau bound)),
// copied to the exchange buffer and padded; the longest such prefix goes to *need
__global__ void __launch_bounds__(128) trim_pack_kernel(const unsigned long long* __restrict__ sorted, int64_t row_stride,
                                                        const int32_t* __restrict__ avail,
                                                        const float* __restrict__ tau_bound, int64_t len,
                                                        unsigned long long* __restrict__ out, int32_t* __restrict__ need) {
  __shared__ int sh_cnt;
  const int q = blockIdx.x;
  const unsigned long long* e = sorted + (int64_t)q * row_stride;
  if (threadIdx.x == 0) {
    const uint32_t bound = score_to_key(tau_bound[q]);
    int lo = 0, hi = avail[q];
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if ((uint32_t)(e[mid] >> 32) <= bound) lo = mid + 1; else hi = mid;
    }
    sh_cnt = lo;
    if (lo > 0) atomicMax(need, lo);
  }
  __syncthreads();
  const int64_t n = sh_cnt;
  for (int64_t i = threadIdx.x; i < len; i += blockDim.x) out[(int64_t)q * len + i] = (i < n) ? e[i] : ~0ull;
}

// Every rank's result record: [per_query f64 x block x 4][ranked_len i32 x block][flags u32 x block][need i32].
// Mean over the q queries with the reduction tree of mean_kernel (metrics.cu) -- so the numbers are
// bit-identical to the single-GPU path --, any tie-overflow flag, and -- only if a merge found a list cut
// short -- the longest list any rank would have sent untruncated.
__global__ void __launch_bounds__(1024) shard_finish_kernel(const uint8_t* __restrict__ rec_all, int world, int block,
                                                            int64_t rec_bytes, int q,
                                                            const uint32_t* __restrict__ local_flags,
                                                            const unsigned long long* __restrict__ counts,
                                                            double* __restrict__ out) {
  __shared__ double sh[32][4];
  __shared__ uint32_t sh_flag;
  if (threadIdx.x == 0) sh_flag = 0u;
  __syncthreads();
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  uint32_t flag = 0u;
  for (int i = threadIdx.x; i < q; i += blockDim.x) {
    const int seg = i / block, j = i - seg * block;
    const uint8_t* rec = rec_all + (int64_t)seg * rec_bytes;
    const double* pq = reinterpret_cast<const double*>(rec) + (int64_t)j * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] += pq[k];
    const uint32_t mf = reinterpret_cast<const uint32_t*>(rec + (int64_t)block * 36)[j];
    flag |= (mf & 1u) | (local_flags[i] != 0u ? 1u : 0u) | (mf & 2u);   // bit 0: exact re-rank; bit 1: a list was cut short
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
  if (flag != 0u) atomicOr(&sh_flag, flag);
  __syncthreads();
  if (threadIdx.x < 4) {
    double v = 0.0;
    for (int w = 0; w < 32; ++w) v += sh[w][threadIdx.x];
    out[threadIdx.x] = v / (double)q;
  }
  if (threadIdx.x == 4) {
    int need = 0;
    for (int w = 0; w < world; ++w)
      need = max(need, *reinterpret_cast<const int32_t*>(rec_all + (int64_t)w * rec_bytes + (int64_t)block * 40));
    out[4] = (double)((sh_flag & 1u) != 0u);
    out[5] = (double)(counts[kMaxClasses] != 0ull);
    out[6] = (sh_flag & 2u) != 0u ? (double)need : 0.0;   // the length to repeat the exchange with, if it must be
  }
}

}  // namespace

int launch_pack_stats(Ctx* c, const float* tau, const uint32_t* flags, const unsigned long long* counts, int q,
                      uint32_t* out, cudaStream_t st) {
  const int items = q > 2 * (kMaxClasses + 1) ? q : 2 * (kMaxClasses + 1);
  pack_stats_kernel<<<(items + 255) / 256, 256, 0, st>>>(tau, flags, counts, q, out);
  DVL_LAUNCH_CHECK(c, "pack_stats_kernel");
  return DVL_OK;
}

int launch_reduce_stats(Ctx* c, const uint32_t* all, int world, int q, float* tau_bound, uint32_t* flags,
                        unsigned long long* counts, cudaStream_t st) {
  const int items = q > kMaxClasses + 1 ? q : kMaxClasses + 1;
  reduce_stats_kernel<<<(items + 255) / 256, 256, 0, st>>>(all, world, q, tau_bound, flags, counts);
  DVL_LAUNCH_CHECK(c, "reduce_stats_kernel");
  return DVL_OK;
}

int launch_trim_pack(Ctx* c, const unsigned long long* sorted, int64_t row_stride, const int32_t* avail,
                     const float* tau_bound, int q, int64_t len, unsigned long long* out, int32_t* need, cudaStream_t st) {
  DVL_CUDA(c, cudaMemsetAsync(need, 0, sizeof(int32_t), st));
  trim_pack_kernel<<<(unsigned)q, 128, 0, st>>>(sorted, row_stride, avail, tau_bound, len, out, need);
  DVL_LAUNCH_CHECK(c, "trim_pack_kernel");
  return DVL_OK;
}

int launch_shard_finish(Ctx* c, const uint8_t* rec_all, int world, int block, int64_t rec_bytes, int q,
                        const uint32_t* local_flags, const unsigned long long* counts, double* out, cudaStream_t st) {
  shard_finish_kernel<<<1, 1024, 0, st>>>(rec_all, world, block, rec_bytes, q, local_flags, counts, out);
  DVL_LAUNCH_CHECK(c, "shard_finish_kernel");
  return DVL_OK;
}

int launch_topk_reset(Ctx* c, uint32_t* cnt, float* tau, uint32_t* flags, int q, uint32_t cnt_value, bool reset_all,
                      cudaStream_t st) {
  fill_topk_state_kernel<<<(q + 255) / 256, 256, 0, st>>>(cnt, tau, flags, q, cnt_value, reset_all ? 1 : 0);
  DVL_LAUNCH_CHECK(c, "fill_topk_state_kernel");
  return DVL_OK;
}

int launch_select(Ctx* c, const SelectArgs& a, cudaStream_t st) {
  if (a.sort_cap < 2 || (a.sort_cap & (a.sort_cap - 1)) != 0 || a.sort_cap > 16384)
    return set_error(c, DVL_ERR_INVALID, "select list capacity must be a power of two <= 16384");
  SelectParams p;
  p.cand = a.cand;
  p.row_stride = a.row_stride;
  p.seg_stride = a.seg_stride;
  p.segs = a.segs;
  p.seg_len = a.seg_len;
  p.cnt = a.cnt;
  p.cap = a.cap;
  p.tau = a.tau;
  p.flags = a.flags;
  p.k = a.k;
  p.sort_cap = a.sort_cap;
  p.trigger = a.trigger;
  p.final_pass = a.final_pass ? 1 : 0;
  p.pilot_pass = a.pilot_pass ? 1 : 0;
  p.rows_total = a.rows_total;
  p.sorted_out = a.sorted_out;
  p.labs = a.labs;
  p.ldl = a.ldl;
  p.idx_out = a.idx_out;
  p.idx_cap = a.idx_cap;
  p.avail = a.avail;
  p.q_offset = a.q_offset;
  p.labels = a.labels;
  p.index_base = a.index_base;
  size_t smem = (size_t)a.sort_cap * sizeof(unsigned long long);
  // final pass of a list of <= 4096 entries: ordered by counting_rank_sort, which needs a second list + its sub-bins
  static const bool bitonic_only = [] {
    const char* e = getenv("DVL_SELECT_BITONIC");
    return e != nullptr && e[0] == '1';
  }();
  // DVL_SELECT_FUSED=0: radix select + compaction + counting sort (the path fused_select_sort falls back to)
  static const bool fused_off = [] {
    const char* e = getenv("DVL_SELECT_FUSED");
    return e != nullptr && e[0] == '0';
  }();
  p.rank_sort = (a.final_pass && a.sort_cap <= 4096 && !bitonic_only) ? (fused_off ? 1 : 2) : 0;
  if (p.rank_sort) smem += (size_t)a.sort_cap * 8 + (size_t)(kSelBins + 4) * 4;
  if (!c->select_attr_set) {
    DVL_CUDA(c, cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
    c->select_attr_set = true;
  }
  select_kernel<<<(unsigned)a.q_count, kSelThreads, smem, st>>>(p);
  DVL_LAUNCH_CHECK(c, "select_kernel");
  return DVL_OK;
}

int launch_trim_count(Ctx* c, const unsigned long long* sorted, int64_t row_stride, const int32_t* avail,
                      const float* tau_bound, int q, int32_t* out_cnt, cudaStream_t st) {
  trim_count_kernel<<<(q + 255) / 256, 256, 0, st>>>(sorted, row_stride, avail, tau_bound, q, out_cnt);
  DVL_LAUNCH_CHECK(c, "trim_count_kernel");
  return DVL_OK;
}

int launch_pack_lists(Ctx* c, const unsigned long long* sorted, int64_t row_stride, const int32_t* cnt, int q,
                      int64_t len, unsigned long long* out, cudaStream_t st) {
  int64_t bx = (len + 255) / 256;
  if (bx < 1) bx = 1;
  if (bx > 64) bx = 64;
  for (int q0 = 0; q0 < q; q0 += 65535) {
    const int qn = (q - q0) < 65535 ? (q - q0) : 65535;
    dim3 grid((unsigned)bx, (unsigned)qn);
    pack_lists_kernel<<<grid, 256, 0, st>>>(sorted + (int64_t)q0 * row_stride, row_stride, cnt + q0, len,
                                            out + (int64_t)q0 * len);
    DVL_LAUNCH_CHECK(c, "pack_lists_kernel");
  }
  return DVL_OK;
}

}  // namespace dvl
