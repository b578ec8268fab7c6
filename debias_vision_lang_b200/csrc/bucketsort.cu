// Ranking stage, full-sort route, fast path (measuring_bias.py:29,57 -> pandas
// nlargest(keep="all"), which for top_n >= N is a stable descending sort of the N scores of a query).
//
// The LSD radix sort of segsort.cu moves every key through HBM eight times (four histogram reads, four
// scatter passes) and spends ~2.4 warp instructions per key and pass on digit matching.  A query's scores
// are not arbitrary 32-bit patterns, though: they are N samples of one smooth distribution.  This file
// sorts them the way that structure allows -- two levels of SCORE-LINEAR bucketing, each key touched twice:
//
//   bucket_plan_kernel       one CTA per query: a strided sample of the query's keys -> score range, a
//                            4096-cell histogram of the sample over that range, and from its prefix sums a
//                            monotone cell -> bucket table that cuts the range into B buckets of about
//                            equal mass (B ~ N / 2048).
//   bucket_partition_kernel  one CTA per 4096-key tile: cell = (smax - s) * scale, bucket = table[cell];
//                            the tile is grouped by bucket in shared memory and appended to the buckets'
//                            storage (one global atomic per tile and bucket).  An entry is
//                            key << 32 | gallery index << 4 | label -- the u64 order of the entries IS the
//                            reference's order (score descending, lowest index first), so the append order
//                            does not matter.
//   bucket_sort_kernel       one CTA per bucket (<= 4096 entries, in shared memory): a second score-linear
//                            split into 2048 sub-bins (counting sort: ~1 entry per sub-bin), then the
//                            few sub-bins holding several entries are put in order by their full u64 value.
//                            Emits the ranked labels (all the metrics need), and on request indices / keys.
//
// Monotonicity is what makes this exact: s -> (smax - s) * scale -> floor is non-increasing in s step by
// step in floating point, the cell -> bucket table is non-decreasing, so every key of bucket b ranks
// before every key of bucket b + 1, and likewise for sub-bins; equal keys always share a bucket and a
// sub-bin and are ordered by the index inside the entry.  NaN keys (rows pandas drops, selectn.py:133-134)
// go to the last sub-bin of the last bucket and sort last.
//
// Data that defeats the balance (massive exact ties: an all-zero query, a tiny score alphabet) overflows
// a bucket's 4096-entry storage; the kernels then raise a flag that travels to the host with the results,
// and the caller repeats the ranking with the radix sort of segsort.cu, which has no such limit.
#include <type_traits>

#include "dvl_common.cuh"
#include "dvl_internal.h"

namespace dvl {

namespace {

constexpr int kCells = 4096;                 // first-level cells per query
constexpr int kBucketCap = 4096;             // entries a bucket can hold
constexpr int kBucketTarget = 2048;          // planned entries per bucket
constexpr int kMaxBuckets = 256;
constexpr int kSubBins = 2048;               // second-level bins per bucket (8 per thread of the sort CTA)
constexpr int kPartThreads = 256;
constexpr int kPartItems = 16;
constexpr int kPartTile = kPartThreads * kPartItems;     // 4096 keys per partition CTA
constexpr int kSortThreads2 = 256;
constexpr int kGroupMax = 48;                // sub-bins above this size: the CTA sorts the bucket bitonically

struct SegPlan {      // per query
  float smax;         // cell = clamp((smax - s) * scale)
  float scale;
};

__device__ __forceinline__ int cell_of(uint32_t key, float smax, float scale) {
  if (key == kNanKey) return kCells - 1;
  const float s = key_to_score(key);
  // fminf/fmaxf keep the order and turn the +-inf cases into the end cells; (smax - s) is NaN only for
  // inf - inf, which fmaxf(.., 0) maps to cell 0 -- correct, s = +inf = smax is the best score
  const float f = fminf(fmaxf((smax - s) * scale, 0.0f), (float)(kCells - 1));
  return (int)f;
}

// ---------------------------------------------------------------------------------------
constexpr int kPlanThreads = 1024;

__global__ void __launch_bounds__(kPlanThreads) bucket_plan_kernel(const uint32_t* __restrict__ keys, int64_t ldk,
                                                                   int64_t n, int n_buckets, int n_samples,
                                                                   SegPlan* __restrict__ plans, uint8_t* __restrict__ luts,
                                                                   uint32_t* __restrict__ bucket_cnt,
                                                                   uint32_t* __restrict__ bucket_cells,
                                                                   uint32_t* __restrict__ bucket_cls) {
  __shared__ uint32_t hist[kCells];
  __shared__ float red_min[32], red_max[32];
  __shared__ uint32_t warp_tot[32];
  __shared__ uint32_t sh_valid;
  const int q = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t* seg = keys + (int64_t)q * ldk;
  if (tid == 0) sh_valid = 0u;
#pragma unroll
  for (int i = 0; i < kCells / kPlanThreads; ++i) hist[i * kPlanThreads + tid] = 0u;
  if (tid < n_buckets) bucket_cnt[(int64_t)q * kMaxBuckets + tid] = 0u;
  if (bucket_cls != nullptr)       // per-bucket class counts, accumulated by the partition pass (fused metrics)
    for (int i = tid; i < n_buckets * kMaxClasses; i += kPlanThreads) bucket_cls[(int64_t)q * kMaxBuckets * kMaxClasses + i] = 0u;
  // The sample: groups of 8 consecutive keys (one 32-byte sector) spread evenly over the segment, read twice
  // (range, then histogram; the second read hits L2).  Single keys at a stride below 32 would touch every
  // 128-byte line of the segment -- the whole key matrix from DRAM, twice (measured: 271 MB for config 2).
  // Four independent loads in flight per thread.
  const int64_t step = n / n_samples > 0 ? n / n_samples : 1;
  const int64_t gstep = 8 * step;
  const int samples = (int)((n + gstep - 1) / gstep) * 8;
  auto sample_pos = [&](int i) { return (int64_t)(i >> 3) * gstep + (i & 7); };
  float lo = __int_as_float(0x7f800000), hi = __int_as_float(0xff800000);
  for (int i0 = tid; i0 < samples; i0 += 4 * kPlanThreads) {
    uint32_t k[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * kPlanThreads;
      const int64_t pos = sample_pos(i);
      k[u] = (i < samples && pos < n) ? __ldg(seg + pos) : kNanKey;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float s = key_to_score(k[u]);
      if (k[u] != kNanKey && fabsf(s) <= 3.4028234664e38f) {      // finite scores define the range
        lo = fminf(lo, s);
        hi = fmaxf(hi, s);
      }
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, off));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, off));
  }
  if (lane == 0) { red_min[warp] = lo; red_max[warp] = hi; }
  __syncthreads();
  lo = red_min[lane];
  hi = red_max[lane];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, off));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, off));
  }
  float scale = 0.0f;                         // no finite sample, or all equal: one cell, one bucket
  if (hi > lo) {
    scale = (float)kCells / (hi - lo);
    if (!(fabsf(scale) <= 3.4028234664e38f)) scale = 0.0f;
  }
  if (!(hi >= lo)) hi = 0.0f;
  uint32_t n_valid = 0;
  for (int i0 = tid; i0 < samples; i0 += 4 * kPlanThreads) {
    uint32_t k[4];
    bool ok[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * kPlanThreads;
      const int64_t pos = sample_pos(i);
      ok[u] = i < samples && pos < n;
      k[u] = ok[u] ? __ldg(seg + pos) : 0u;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (ok[u]) {
        atomicAdd(&hist[cell_of(k[u], hi, scale)], 1u);
        ++n_valid;
      }
  }
  n_valid = __reduce_add_sync(0xffffffffu, n_valid);
  if (lane == 0) atomicAdd(&sh_valid, n_valid);
  __syncthreads();
  // exclusive prefix over the cells, 4 consecutive cells per thread
  const uint4 c4 = reinterpret_cast<const uint4*>(hist)[tid];
  const uint32_t local[4] = {c4.x, c4.y, c4.z, c4.w};
  const uint32_t sum = c4.x + c4.y + c4.z + c4.w;
  uint32_t incl = sum;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t o = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += o;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  uint32_t wsum = warp_tot[lane];
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t o = __shfl_up_sync(0xffffffffu, wsum, off);
    if (lane >= off) wsum += o;
  }
  uint32_t run = incl - sum + __shfl_sync(0xffffffffu, wsum - warp_tot[lane], warp);
  uint32_t packed = 0u;
  const float slices = (float)n_buckets / (float)(sh_valid > 0u ? sh_valid : 1u);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    // bucket of a cell = the equal-mass slice its first sample falls into (monotone in the prefix count)
    uint32_t b = (uint32_t)((float)run * slices);
    if (b > (uint32_t)(n_buckets - 1)) b = (uint32_t)(n_buckets - 1);
    packed |= b << (8 * j);
    run += local[j];
  }
  reinterpret_cast<uint32_t*>(luts + (int64_t)q * kCells)[tid] = packed;
  // first and last cell of every bucket (first | last << 16): the bucket's score range for the second level.
  // hist[] has been read by everyone (barrier above): its first kPlanThreads words now hold the table
  hist[tid] = packed;
  __syncthreads();
  {
    const uint32_t prev = tid > 0 ? (hist[tid - 1] >> 24) : 0xFFFFu;
    const uint32_t next = tid < kPlanThreads - 1 ? (hist[tid + 1] & 0xFFu) : 0xFFFFu;
    uint16_t* cells16 = reinterpret_cast<uint16_t*>(bucket_cells + (int64_t)q * kMaxBuckets);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t b = (packed >> (8 * j)) & 0xFFu;
      const uint32_t pb = j == 0 ? prev : ((packed >> (8 * (j - 1))) & 0xFFu);
      const uint32_t nb = j == 3 ? next : ((packed >> (8 * (j + 1))) & 0xFFu);
      if (b != pb) cells16[2 * b] = (uint16_t)(tid * 4 + j);
      if (b != nb) cells16[2 * b + 1] = (uint16_t)(tid * 4 + j);
    }
  }
  if (tid == 0) {
    plans[q].smax = hi;
    plans[q].scale = scale;
  }
}

// ---------------------------------------------------------------------------------------
struct PartSmem {          // the partition CTA's shared memory (one copy for both instantiations of the body)
  uint8_t* lut;            // [kCells] cell -> bucket
  uint32_t* hist;          // [kMaxBuckets] entries of this tile per bucket, then their staging offset
  uint32_t* gbase;         // [kMaxBuckets] where the tile's run starts in the bucket's storage
  unsigned long long* stage;   // [kPartTile] the tile grouped by bucket
  uint8_t* stage_b;        // [kPartTile] bucket of the staged entry
  uint32_t* warp_tot;
  uint32_t* cls;           // [kMaxBuckets][8] rows of this tile per (bucket, class), two 16-bit counts per word
};

template <bool kFull>
__device__ __forceinline__ void
partition_body(const uint32_t* __restrict__ keys, int64_t ldk, int64_t n, const uint8_t* __restrict__ gallery_labels,
               const SegPlan* __restrict__ plans, const uint8_t* __restrict__ luts, uint32_t* __restrict__ bucket_cnt,
               int n_buckets, unsigned long long* __restrict__ store, uint32_t* __restrict__ nan_cnt,
               uint32_t* __restrict__ overflow, uint32_t* __restrict__ bucket_cls, const PartSmem& sm) {
  uint8_t* lut = sm.lut;
  uint32_t* hist = sm.hist;
  uint32_t* gbase = sm.gbase;
  unsigned long long* stage = sm.stage;
  uint8_t* stage_b = sm.stage_b;
  uint32_t* warp_tot = sm.warp_tot;
  const int q = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  reinterpret_cast<uint4*>(lut)[tid] = __ldg(reinterpret_cast<const uint4*>(luts + (int64_t)q * kCells) + tid);
  hist[tid] = 0u;                              // kPartThreads == kMaxBuckets
  const bool count_cls = bucket_cls != nullptr;
  if (count_cls) {
    reinterpret_cast<uint4*>(sm.cls)[2 * tid] = make_uint4(0u, 0u, 0u, 0u);
    reinterpret_cast<uint4*>(sm.cls)[2 * tid + 1] = make_uint4(0u, 0u, 0u, 0u);
  }
  const float smax = plans[q].smax, scale = plans[q].scale;
  const uint32_t* seg = keys + (int64_t)q * ldk;
  const int64_t base = (int64_t)blockIdx.x * kPartTile;
  const int count = kFull ? kPartTile : (int)(n - base);
  // thread owns keys base + r*1024 + tid*4 + {0..3}; ldk and base are multiples of 4 (aligned 16-byte loads;
  // the padding of the last group is read but never used)
  uint32_t key[kPartItems];
  uint32_t lab4[kPartItems / 4];
  const bool lab_aligned = (reinterpret_cast<uintptr_t>(gallery_labels) & 3u) == 0;   // the C ABI takes any pointer
#pragma unroll
  for (int r = 0; r < kPartItems / 4; ++r) {
    const int p = r * (kPartThreads * 4) + tid * 4;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    uint32_t l = 0u;
    if (kFull || p < count) {
      v = __ldg(reinterpret_cast<const uint4*>(seg + base + p));
      if ((kFull || p + 4 <= count) && lab_aligned) {
        l = __ldg(reinterpret_cast<const uint32_t*>(gallery_labels + base + p));
      } else {
        for (int j = 0; j < 4 && p + j < count; ++j) l |= (uint32_t)gallery_labels[base + p + j] << (8 * j);
      }
    }
    key[4 * r] = v.x; key[4 * r + 1] = v.y; key[4 * r + 2] = v.z; key[4 * r + 3] = v.w;
    lab4[r] = l;
  }
  __syncthreads();
  uint32_t slot[kPartItems];                   // bucket << 16 | arrival order inside the tile's bucket run
  uint32_t nans = 0;
#pragma unroll
  for (int i = 0; i < kPartItems; ++i) {
    const int p = (i >> 2) * (kPartThreads * 4) + tid * 4 + (i & 3);
    slot[i] = 0u;
    if (kFull || p < count) {
      const uint32_t b = lut[cell_of(key[i], smax, scale)];
      nans += (key[i] == kNanKey);
      slot[i] = (b << 16) | atomicAdd(&hist[b], 1u);
      if (count_cls) {
        const uint32_t lab = (lab4[i >> 2] >> (8 * (i & 3))) & 0xFu;
        atomicAdd(&sm.cls[b * 8 + (lab >> 1)], 1u << (16 * (lab & 1u)));     // a tile holds 4096 rows: 16 bits do
      }
    }
  }
  if (nan_cnt != nullptr && nans != 0u) atomicAdd(nan_cnt + q, nans);      // NaN scores are rare
  __syncthreads();
  {
    // reserve the runs in the buckets' storage; exclusive prefix of the tile's bucket counts = staging offsets
    const uint32_t mine = hist[tid];
    uint32_t g = 0u;
    if (tid < n_buckets && mine != 0u) g = atomicAdd(bucket_cnt + (int64_t)q * kMaxBuckets + tid, mine);
    uint32_t incl = mine;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += o;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    uint32_t excl = incl - mine;
#pragma unroll
    for (int w = 0; w < kPartThreads / 32; ++w) excl += (w < warp) ? warp_tot[w] : 0u;
    if (mine != 0u && g + mine > (uint32_t)kBucketCap) atomicOr(overflow, 1u);
    gbase[tid] = g - excl;                     // storage position = gbase[b] + staging position
    hist[tid] = excl;
    if (count_cls && mine != 0u) {             // this tile's rows of bucket `tid`, per class
      uint32_t* dst = bucket_cls + ((int64_t)q * kMaxBuckets + tid) * kMaxClasses;
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        const uint32_t v = sm.cls[tid * 8 + w];
        if ((v & 0xFFFFu) != 0u) atomicAdd(dst + 2 * w, v & 0xFFFFu);
        if ((v >> 16) != 0u) atomicAdd(dst + 2 * w + 1, v >> 16);
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kPartItems; ++i) {
    const int p = (i >> 2) * (kPartThreads * 4) + tid * 4 + (i & 3);
    if (kFull || p < count) {
      const uint32_t b = slot[i] >> 16;
      const uint32_t lab = (lab4[i >> 2] >> (8 * (i & 3))) & 0xFu;
      const uint32_t sp = hist[b] + (slot[i] & 0xFFFFu);
      stage[sp] = ((unsigned long long)key[i] << 32) | (unsigned long long)(((uint32_t)(base + p) << 4) | lab);
      stage_b[sp] = (uint8_t)b;
    }
  }
  __syncthreads();
  // (Copying whole runs, one warp per bucket, so that the bucket and its storage offset are warp-uniform and the
  // bucket byte per staged entry disappears, was measured: 0.1325 instead of 0.125 ms -- runs of ~77 entries leave
  // 40 % of the lanes idle in their last iteration.)
  unsigned long long* out = store + (int64_t)q * n_buckets * kBucketCap;
#pragma unroll 4
  for (int p = tid; p < count; p += kPartThreads) {
    const uint32_t b = stage_b[p];
    const uint32_t pos = gbase[b] + (uint32_t)p;
    if (pos < (uint32_t)kBucketCap) out[b * kBucketCap + pos] = stage[p];
  }
}

__global__ void __launch_bounds__(kPartThreads, 4)
bucket_partition_kernel(const uint32_t* __restrict__ keys, int64_t ldk, int64_t n,
                        const uint8_t* __restrict__ gallery_labels, const SegPlan* __restrict__ plans,
                        const uint8_t* __restrict__ luts, uint32_t* __restrict__ bucket_cnt, int n_buckets,
                        unsigned long long* __restrict__ store, uint32_t* __restrict__ nan_cnt,
                        uint32_t* __restrict__ overflow, uint32_t* __restrict__ bucket_cls) {
  __shared__ __align__(16) uint8_t lut[kCells];
  __shared__ uint32_t hist[kMaxBuckets];
  __shared__ uint32_t gbase[kMaxBuckets];
  __shared__ __align__(16) unsigned long long stage[kPartTile];
  __shared__ uint8_t stage_b[kPartTile];
  __shared__ uint32_t warp_tot[kPartThreads / 32];
  // the per-(bucket, class) counters live in the staging buffer: they are flushed before the first entry is staged
  const PartSmem sm = {lut, hist, gbase, stage, stage_b, warp_tot, reinterpret_cast<uint32_t*>(stage)};
  if ((int64_t)(blockIdx.x + 1) * kPartTile <= n)     // all tiles of a segment but (at most) the last
    partition_body<true>(keys, ldk, n, gallery_labels, plans, luts, bucket_cnt, n_buckets, store, nan_cnt, overflow,
                         bucket_cls, sm);
  else
    partition_body<false>(keys, ldk, n, gallery_labels, plans, luts, bucket_cnt, n_buckets, store, nan_cnt, overflow,
                          bucket_cls, sm);
}

// ---------------------------------------------------------------------------------------
struct BucketSortParams {
  const unsigned long long* store;
  const uint32_t* bucket_cnt;
  const SegPlan* plans;
  const uint32_t* bucket_cells;      // [q][kMaxBuckets] first | last << 16 first-level cell of the bucket
  int n_buckets;
  uint8_t* labs_out; int64_t ldl;     // null when the metrics are fused (nobody reads the ranked labels then)
  int32_t* idx_out; int64_t ldi;
  uint32_t* keys_out; int64_t ldko;
  // fused NDKL partial sums + class counts of the ranked rows (whole-segment ranking only; see FusedMetrics)
  const uint32_t* bucket_cls;        // [q][kMaxBuckets][16] rows per class in each bucket (partition pass)
  const uint32_t* nan_cnt;           // [q]
  const unsigned long long* class_counts;   // [16] global label histogram
  const double* dxl;                 // metric tables (metrics.cu)
  const double* pwi;
  double* partial;                   // [q][n_buckets][2]
  uint32_t* ranked_cls;              // [q][n_buckets + 1][16] class counts of each bucket's ranked rows
  int64_t n, n_gallery;
  int n_classes;
};


// score of a (non-NaN) key, two instructions
__device__ __forceinline__ float key_score_fast(uint32_t k) {
  return __uint_as_float(k ^ (~(uint32_t)((int32_t)k >> 31) & 0x7FFFFFFFu));
}

// sub-bin of a key inside its bucket: (hi - s) * scale, truncated with saturation (negative and NaN -> 0),
// capped at the last sub-bin; NaN keys (dropped rows) go last.  Non-decreasing in the key.
__device__ __forceinline__ uint32_t subbin_of(uint32_t k, float hi, float scale) {
  const uint32_t sb = min(__float2uint_rz((hi - key_score_fast(k)) * scale), (uint32_t)(kSubBins - 1));
  return k == kNanKey ? (uint32_t)(kSubBins - 1) : sb;
}

constexpr int kBinsPerThread = kSubBins / kSortThreads2;      // 8 consecutive sub-bins per thread
// sub-bin sorted entries as two 32-bit arrays, sub-bins (+ one sentinel word in front), ranked labels
constexpr size_t kSortSmem = (size_t)kBucketCap * 8 + (size_t)(kSubBins + 4) * 4 + kBucketCap;

// (Measured and not kept: a persistent form -- 5 CTAs per SM walking the (query, bucket) pairs, the next bucket's keys
// requested into 8 registers per thread right after the current bucket's histogram pass so that they arrive during
// its scan / scatter / ranking.  By ncu a third of this kernel's samples sit on the first use of the loaded key, yet
// the persistent kernel ran 0.262 instead of 0.245 ms: five independent CTAs per SM in different phases already
// overlap that wait, and the one-CTA-per-bucket grid balances the tail better.)
// The kernel keeps nothing per entry in registers between its phases: every phase is a short strided loop over
// the bucket that re-reads what it needs (the bucket's entries from L1/L2, the rest from shared memory).  An
// earlier version held 16 entries per thread in registers across the phases; fully unrolled it executed 2.5x the
// instructions (register shuffling, spills, one branch region per conditional) and ran no faster for it.
// ---- NDKL partial sums of one bucket, fused into its sort (measuring_bias.py:17-44) ----
// Same arithmetic as metrics.cu's ndkl_partial_kernel (summation by parts over fp64 tables), on the bucket's
// ranked labels while they are still in shared memory: the labels never travel to HBM and back, and the
// sub_counts / sub_scan / ndkl_partial launches disappear.  Needs, per class, the number of rows ranked in
// earlier buckets (the partition pass counted rows per bucket and class) and the global ranked length
// m = n - #NaN rows (whole-segment ranking only).  A thread owns 16 consecutive ranks; class counts travel as
// 16-bit fields of W 64-bit words (W = ceil(C / 4)) through one block-wide exclusive scan.  Fixed reduction
// order: the result does not depend on scheduling.
template <int W>
__device__ __forceinline__ void fused_metrics(const BucketSortParams& p, int q, int b, int count, uint32_t before,
                                              const uint8_t* labs_s, unsigned long long* scratch, int tid) {
  __shared__ double ln_d[kMaxClasses];
  __shared__ uint32_t cbase[kMaxClasses];
  __shared__ double red[8][2];
  const int lane = tid & 31, warp = tid >> 5;
  unsigned long long* warp_scan = scratch;                      // [8][W]   (the sorted keys are dead by now)
  uint32_t* cpart = reinterpret_cast<uint32_t*>(scratch + 8 * 4);   // [16 groups][16 classes]
  // rows of each class ranked before this bucket: 16 groups of threads stride over the earlier buckets
  {
    const int c = tid & 15, g = tid >> 4;
    uint32_t acc = 0;
    const uint32_t* cls = p.bucket_cls + (int64_t)q * kMaxBuckets * kMaxClasses;
    for (int bb = g; bb < b; bb += 16) acc += __ldg(cls + bb * kMaxClasses + c);
    cpart[g * 16 + c] = acc;
  }
  if (tid < kMaxClasses) {
    double v = 0.0;
    if (tid < p.n_classes) {
      const unsigned long long nc = p.class_counts[tid];
      v = log((double)(nc > 0 ? nc : 1ull) / (double)p.n_gallery);      // dem_par: max(count, 1) / N  (:27)
    }
    ln_d[tid] = v;
  }
  __syncthreads();
  if (tid < kMaxClasses) {
    uint32_t acc = 0;
#pragma unroll
    for (int g = 0; g < 16; ++g) acc += cpart[g * 16 + tid];
    cbase[tid] = acc;
  }
  const long long m = (long long)p.n - (long long)p.nan_cnt[q];          // ranked rows of the query
  long long valid_ll = m - (long long)before;
  const int valid = valid_ll < 0 ? 0 : (valid_ll > count ? count : (int)valid_ll);   // NaN rows sit at the very end
  const int r0 = tid * 16;                                                // this thread's first local rank
  const uint4 l16 = reinterpret_cast<const uint4*>(labs_s)[tid];         // 16 labels
  const uint32_t lw[4] = {l16.x, l16.y, l16.z, l16.w};
  unsigned long long own[W];
#pragma unroll
  for (int w = 0; w < W; ++w) own[w] = 0ull;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const uint32_t l = (lw[j >> 2] >> (8 * (j & 3))) & 0xFu;
    const unsigned long long inc = (r0 + j < valid) ? (1ull << (16 * (l & 3u))) : 0ull;
#pragma unroll
    for (int w = 0; w < W; ++w) own[w] += ((l >> 2) == (uint32_t)w) ? inc : 0ull;   // labels >= C are rejected by the caller
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
    if (lane == 31) warp_scan[warp * W + w] = incl[w];
  }
  __syncthreads();
  unsigned long long run[W];                     // rows of each class among this bucket's ranks before r0
#pragma unroll
  for (int w = 0; w < W; ++w) {
    run[w] = incl[w] - own[w];
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) run[w] += (ww < warp) ? warp_scan[ww * W + w] : 0ull;
  }
  if (tid == kSortThreads2 - 1) {                // inclusive total of the last thread = class counts of the ranked rows
    uint32_t* out = p.ranked_cls + ((int64_t)q * (p.n_buckets + 1) + b) * kMaxClasses;
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c) out[c] = 0u;
#pragma unroll
    for (int w = 0; w < W; ++w) {
      const unsigned long long tot = run[w] + own[w];
#pragma unroll
      for (int f = 0; f < 4; ++f) out[4 * w + f] = (uint32_t)(tot >> (16 * f)) & 0xFFFFu;
    }
  }
  double acc_eq = 0.0, acc_d = 0.0;
  if (r0 < valid) {
    const double pwi_m = __ldg(p.pwi + m);
    const double* pw = p.pwi + before + r0;      // PWI[j - 1] of rank j = before + r0 + jj + 1
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (r0 + j < valid) {
        const uint32_t l = (lw[j >> 2] >> (8 * (j & 3))) & 0xFu;
        const int sh = 16 * (int)(l & 3u);
        unsigned long long mine = 0ull;
#pragma unroll
        for (int w = 0; w < W; ++w) {
          const bool here = (l >> 2) == (uint32_t)w;
          mine = here ? run[w] : mine;
          run[w] += here ? (1ull << sh) : 0ull;
        }
        const uint32_t t = cbase[l] + ((uint32_t)(mine >> sh) & 0xFFFFu) + 1u;   // this row is the t-th of its class
        const double tj = pwi_m - __ldg(pw + j);
        acc_eq = fma(tj, __ldg(p.dxl + t), acc_eq);
        acc_d = fma(tj, ln_d[l], acc_d);
      }
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    acc_eq += __shfl_xor_sync(0xffffffffu, acc_eq, off);
    acc_d += __shfl_xor_sync(0xffffffffu, acc_d, off);
  }
  if (lane == 0) { red[warp][0] = acc_eq; red[warp][1] = acc_d; }
  __syncthreads();
  if (tid == 0) {
    double e = 0.0, d = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { e += red[w][0]; d += red[w][1]; }
    double* out = p.partial + ((int64_t)q * p.n_buckets + b) * 2;
    out[0] = e;
    out[1] = e - d;
  }
}

template <int W>
__global__ void __launch_bounds__(kSortThreads2, 5) bucket_sort_kernel(const BucketSortParams p) {
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  // key | (index << 4 | label): the ranking compares keys and only looks at the index word of an equal key
  uint32_t* s_key = reinterpret_cast<uint32_t*>(smem_dyn);                           // [kBucketCap]
  uint32_t* s_val = s_key + kBucketCap;                                              // [kBucketCap]
  // bins[-1] = 0 is the end of "the sub-bin before the first"; the four words keep bins 16-byte aligned
  uint32_t* bins = s_val + kBucketCap + 4;                                           // [kSubBins]
  uint8_t* labs_s = reinterpret_cast<uint8_t*>(bins + kSubBins);                     // [kBucketCap]
  __shared__ uint32_t red_min[8], red_max[8];
  __shared__ uint32_t warp_tot[8];
  __shared__ uint32_t sh_big;
  const int q = blockIdx.y, b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t* cnts = p.bucket_cnt + (int64_t)q * kMaxBuckets;
  const uint32_t raw = cnts[b];
  const int count = raw > (uint32_t)kBucketCap ? kBucketCap : (int)raw;   // overflowed: the caller falls back
  if (count == 0) {
    if constexpr (W > 0) {                             // an empty bucket still owns its slots of the fused outputs
      if (tid < kMaxClasses) p.ranked_cls[((int64_t)q * (p.n_buckets + 1) + b) * kMaxClasses + tid] = 0u;
      if (tid < 2) p.partial[((int64_t)q * p.n_buckets + b) * 2 + tid] = 0.0;
    }
    return;
  }
  // ranks of this bucket's rows start after the rows of the buckets before it
  uint32_t before = 0;
  for (int i = lane; i < b; i += 32) before += cnts[i];
  before = __reduce_add_sync(0xffffffffu, before);
  // 8 consecutive sub-bins per thread: zeroed and scanned by their owner
  reinterpret_cast<uint4*>(bins)[2 * tid] = make_uint4(0u, 0u, 0u, 0u);
  reinterpret_cast<uint4*>(bins)[2 * tid + 1] = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0) { sh_big = 0u; bins[-1] = 0u; }
  const unsigned long long* src = p.store + ((int64_t)q * p.n_buckets + b) * kBucketCap;

  // ---- the bucket's score range: from the first-level cells it is made of; the two end buckets (which also
  // hold whatever lies outside the sampled range, +-inf and NaN rows) and degenerate plans measure it
  const SegPlan plan = p.plans[q];
  const uint32_t cw = p.bucket_cells[(int64_t)q * kMaxBuckets + b];
  const uint32_t c0 = cw & 0xFFFFu, c1 = cw >> 16;
  float hi = 0.0f, scale = 0.0f;
  if (c0 == 0u || c1 >= (uint32_t)(kCells - 1) || c1 < c0 || !(plan.scale > 0.0f)) {
    uint32_t kmin = 0xFFFFFFFFu, kmax = 0u;            // over the keys of real scores (a NaN key is 0xFFFFFFFF)
#pragma unroll 8
    for (int pos = tid; pos < count; pos += kSortThreads2) {
      const uint32_t k = (uint32_t)(__ldg(src + pos) >> 32);
      kmin = min(kmin, k);
      kmax = max(kmax, k == kNanKey ? 0u : k);
    }
    kmin = __reduce_min_sync(0xffffffffu, kmin);
    kmax = __reduce_max_sync(0xffffffffu, kmax);
    if (lane == 0) { red_min[warp] = kmin; red_max[warp] = kmax; }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      kmin = min(kmin, red_min[w]);
      kmax = max(kmax, red_max[w]);
    }
    if (kmax > kmin) {                                 // at least two distinct real scores; +-inf ends clamped
      hi = fminf(key_score_fast(kmin), 3.4028234664e38f);
      const float lo = fmaxf(key_score_fast(kmax), -3.4028234664e38f);
      if (hi > lo) scale = (float)kSubBins / (hi - lo);
    }
  } else {
    hi = plan.smax - (float)c0 / plan.scale;
    scale = (float)kSubBins * plan.scale / (float)(c1 - c0 + 1u);
  }
  if (!(scale <= 3.4028234664e38f)) scale = 0.0f;
  __syncthreads();

  // ---- counting sort over the sub-bins: histogram (no value returned: reductions, not atomics), prefix, then
  // every entry takes the next free slot of its sub-bin -- the order inside a sub-bin is arbitrary, the ranking
  // resolves it.  (A planned bucket is 8 rounds of 256 entries: all its loads are in flight together -- the CTA's
  // lifetime is a chain of memory round trips, and with two loads in flight this loop alone was four of them.)
#pragma unroll 8
  for (int pos = tid; pos < count; pos += kSortThreads2) {
    const uint32_t k = (uint32_t)(__ldg(src + pos) >> 32);
    atomicAdd(&bins[subbin_of(k, hi, scale)], 1u);
  }
  __syncthreads();
  {
    // exclusive prefix over the sub-bins: bins[b] = first slot of sub-bin b (the scatter advances it to its end)
    const uint4 v0 = reinterpret_cast<const uint4*>(bins)[2 * tid];
    const uint4 v1 = reinterpret_cast<const uint4*>(bins)[2 * tid + 1];
    const uint32_t cnt8[kBinsPerThread] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    uint32_t start[kBinsPerThread];
    uint32_t sum = 0, mx = 0;
#pragma unroll
    for (int j = 0; j < kBinsPerThread; ++j) {
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
    if (mx > (uint32_t)kGroupMax) sh_big = 1u;
    __syncthreads();
    uint32_t base = incl - sum;
#pragma unroll
    for (int w = 0; w < 8; ++w) base += (w < warp) ? warp_tot[w] : 0u;
#pragma unroll
    for (int j = 0; j < kBinsPerThread; ++j) start[j] += base;
    reinterpret_cast<uint4*>(bins)[2 * tid] = make_uint4(start[0], start[1], start[2], start[3]);
    reinterpret_cast<uint4*>(bins)[2 * tid + 1] = make_uint4(start[4], start[5], start[6], start[7]);
  }
  __syncthreads();
#pragma unroll 4
  for (int pos = tid; pos < count; pos += kSortThreads2) {
    const unsigned long long e = __ldg(src + pos);     // second read: L1 / L2
    const uint32_t k = (uint32_t)(e >> 32);
    const uint32_t dst = atomicAdd(&bins[subbin_of(k, hi, scale)], 1u);
    s_key[dst] = k;
    s_val[dst] = (uint32_t)e;
  }
  __syncthreads();
  if (sh_big == 0u) {
    // Entry-parallel ranking over the sub-bin sorted list: every entry counts the entries of its own sub-bin
    // that rank before it (a sub-bin holds ~1 entry).  The first four candidates are compared without a loop
    // (indices clamped into the sub-bin, results masked), a lane only loops for a sub-bin of more than four.
    const bool want_idx = p.idx_out != nullptr, want_keys = p.keys_out != nullptr;
#pragma unroll 2
    for (int d = tid; d < count; d += kSortThreads2) {
      const uint32_t k = s_key[d], v = s_val[d];
      const uint32_t sb = subbin_of(k, hi, scale);
      const uint32_t s0 = bins[(int)sb - 1], end = bins[sb];     // after the scatter bins[b] is the END of sub-bin b
      const uint32_t g = end - s0, last = end - 1u;
      uint32_t rank = s0;
#pragma unroll
      for (uint32_t j = 0; j < 4u; ++j) {
        const uint32_t at = min(s0 + j, last);
        const uint32_t kk = s_key[at];
        bool first = kk < k;
        if (kk == k) first = s_val[at] < v;            // itself, or an exact tie: the gallery index decides
        rank += (j < g && first) ? 1u : 0u;
      }
      for (uint32_t j = 4u; j < g; ++j) {
        const uint32_t kk = s_key[s0 + j];
        rank += (kk < k || (kk == k && s_val[s0 + j] < v)) ? 1u : 0u;
      }
      labs_s[rank] = (uint8_t)(v & 0xFu);
      if (want_idx) p.idx_out[(int64_t)q * p.ldi + before + rank] = (int32_t)(v >> 4);
      if (want_keys) p.keys_out[(int64_t)q * p.ldko + before + rank] = k;
    }
  } else {
    // a sub-bin is crowded (heavy ties inside the bucket): bitonic sort of the whole bucket by (key, index word)
    int pow2 = 1;
    while (pow2 < count) pow2 <<= 1;
    for (int i = count + tid; i < pow2; i += kSortThreads2) {
      s_key[i] = 0xFFFFFFFFu;
      s_val[i] = 0xFFFFFFFFu;
    }
    __syncthreads();
    for (int size = 2; size <= pow2; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int t = tid; t < (pow2 >> 1); t += kSortThreads2) {
          const int i = 2 * t - (t & (stride - 1));
          const int j = i + stride;
          const bool up = (i & size) == 0;
          const unsigned long long x = ((unsigned long long)s_key[i] << 32) | s_val[i];
          const unsigned long long y = ((unsigned long long)s_key[j] << 32) | s_val[j];
          if ((x > y) == up) {
            s_key[i] = (uint32_t)(y >> 32); s_val[i] = (uint32_t)y;
            s_key[j] = (uint32_t)(x >> 32); s_val[j] = (uint32_t)x;
          }
        }
        __syncthreads();
      }
    }
    for (int i = tid; i < count; i += kSortThreads2) {
      const uint32_t v = s_val[i];
      labs_s[i] = (uint8_t)(v & 0xFu);
      if (p.idx_out != nullptr) p.idx_out[(int64_t)q * p.ldi + before + i] = (int32_t)(v >> 4);
      if (p.keys_out != nullptr) p.keys_out[(int64_t)q * p.ldko + before + i] = s_key[i];
    }
  }
  __syncthreads();
  // ---- the bucket's ranked labels are in shared memory: fused NDKL partial sums, and / or the labels themselves
  if constexpr (W > 0) {
    fused_metrics<W>(p, q, b, count, before, labs_s, reinterpret_cast<unsigned long long*>(s_key), tid);
    if (p.labs_out == nullptr) return;
  }
  uint8_t* labs = p.labs_out + (int64_t)q * p.ldl + before;
  // 4 bytes per store where the destination allows it
  const int head = (int)((4u - (uint32_t)(reinterpret_cast<uintptr_t>(labs) & 3u)) & 3u);
  for (int i = tid; i < head && i < count; i += kSortThreads2) labs[i] = labs_s[i];
  const int words = count > head ? (count - head) >> 2 : 0;
  for (int i = tid; i < words; i += kSortThreads2) {
    const uint8_t* sp = labs_s + head + 4 * i;
    reinterpret_cast<uint32_t*>(labs + head)[i] =
        (uint32_t)sp[0] | ((uint32_t)sp[1] << 8) | ((uint32_t)sp[2] << 16) | ((uint32_t)sp[3] << 24);
  }
  for (int i = head + 4 * words + tid; i < count; i += kSortThreads2) labs[i] = labs_s[i];
}

}  // namespace

bool bucket_sort_eligible(int64_t n) {
  static const bool off = [] {
    const char* e = getenv("DVL_SORT_LSD");
    return e != nullptr && e[0] == '1';
  }();
  return !off && n >= 1 && n <= (int64_t)kMaxBuckets * kBucketTarget * 85 / 100;
}

static int buckets_for(int64_t n) {
  int b = (int)((n + kBucketTarget - 1) / kBucketTarget);
  return b < 1 ? 1 : (b > kMaxBuckets ? kMaxBuckets : b);
}

int bucket_sort_buckets(int64_t n) { return buckets_for(n); }

size_t bucket_sort_workspace_per_query(int64_t n) {
  return (size_t)buckets_for(n) * kBucketCap * 8 + kCells + kMaxBuckets * 8 + sizeof(SegPlan);
}

int launch_bucket_sort(Ctx* c, const BucketSortArgs& a, cudaStream_t st) {
  if (!bucket_sort_eligible(a.n)) return set_error(c, DVL_ERR_INVALID, "segment too long for the bucket sort");
  if (a.ldk % 4 != 0) return set_error(c, DVL_ERR_INVALID, "sort pitch must be a multiple of 4");
  const int nb = buckets_for(a.n);
  int n_samples = 64 * nb;
  if (n_samples < 1024) n_samples = 1024;
  DVL_TRY(ensure(c, c->bk_store, (size_t)a.q * nb * kBucketCap * 8));
  DVL_TRY(ensure(c, c->bk_cnt, (size_t)a.q * kMaxBuckets * 4));
  DVL_TRY(ensure(c, c->bk_cells, (size_t)a.q * kMaxBuckets * 4));
  DVL_TRY(ensure(c, c->bk_lut, (size_t)a.q * kCells));
  DVL_TRY(ensure(c, c->bk_plan, (size_t)a.q * sizeof(SegPlan)));
  const bool fused = a.partial != nullptr;
  if (fused && (a.ranked_cls == nullptr || a.nan_cnt == nullptr || a.class_counts == nullptr || a.dxl == nullptr ||
                a.pwi == nullptr || a.n_classes < 1 || a.n_classes > kMaxClasses))
    return set_error(c, DVL_ERR_INVALID, "fused metrics need the label statistics, tables and output buffers");
  if (!fused && a.labs_out == nullptr) return set_error(c, DVL_ERR_INVALID, "no output for the ranked labels");
  if (fused) DVL_TRY(ensure(c, c->bk_cls, (size_t)a.q * kMaxBuckets * kMaxClasses * 4));
  if (!c->bucket_attr_set) {
    DVL_CUDA(c, cudaFuncSetAttribute(bucket_sort_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSortSmem));
    DVL_CUDA(c, cudaFuncSetAttribute(bucket_sort_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSortSmem));
    DVL_CUDA(c, cudaFuncSetAttribute(bucket_sort_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSortSmem));
    DVL_CUDA(c, cudaFuncSetAttribute(bucket_sort_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSortSmem));
    DVL_CUDA(c, cudaFuncSetAttribute(bucket_sort_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSortSmem));
    c->bucket_attr_set = true;
  }
  if (a.nan_cnt != nullptr) DVL_CUDA(c, cudaMemsetAsync(a.nan_cnt, 0, (size_t)a.q * sizeof(uint32_t), st));
  const int tiles = (int)((a.n + kPartTile - 1) / kPartTile);
  for (int q0 = 0; q0 < a.q; q0 += 65535) {
    const int qn = (a.q - q0) < 65535 ? (a.q - q0) : 65535;
    const uint32_t* keys = a.keys + (int64_t)q0 * a.ldk;
    SegPlan* plans = (SegPlan*)c->bk_plan.ptr + q0;
    uint8_t* luts = (uint8_t*)c->bk_lut.ptr + (int64_t)q0 * kCells;
    uint32_t* cnt = (uint32_t*)c->bk_cnt.ptr + (int64_t)q0 * kMaxBuckets;
    uint32_t* cells = (uint32_t*)c->bk_cells.ptr + (int64_t)q0 * kMaxBuckets;
    uint32_t* cls = fused ? (uint32_t*)c->bk_cls.ptr + (int64_t)q0 * kMaxBuckets * kMaxClasses : nullptr;
    unsigned long long* store = (unsigned long long*)c->bk_store.ptr + (int64_t)q0 * nb * kBucketCap;
    {
      StageScope scope(c, DVL_STAGE_SORT_HIST, st);
      bucket_plan_kernel<<<(unsigned)qn, kPlanThreads, 0, st>>>(keys, a.ldk, a.n, nb, n_samples, plans, luts, cnt, cells, cls);
      DVL_LAUNCH_CHECK(c, "bucket_plan_kernel");
    }
    {
      StageScope scope(c, DVL_STAGE_SORT_SCATTER, st);
      dim3 grid((unsigned)tiles, (unsigned)qn);
      bucket_partition_kernel<<<grid, kPartThreads, 0, st>>>(keys, a.ldk, a.n, a.gallery_labels, plans, luts, cnt, nb, store,
                                                            a.nan_cnt != nullptr ? a.nan_cnt + q0 : nullptr, a.overflow, cls);
      DVL_LAUNCH_CHECK(c, "bucket_partition_kernel");
    }
    {
      StageScope scope(c, DVL_STAGE_SORT_SCAN, st);
      BucketSortParams p;
      p.store = store;
      p.bucket_cnt = cnt;
      p.plans = plans;
      p.bucket_cells = cells;
      p.n_buckets = nb;
      p.labs_out = a.labs_out != nullptr ? a.labs_out + (int64_t)q0 * a.ldl : nullptr;
      p.ldl = a.ldl;
      p.bucket_cls = cls;
      p.nan_cnt = a.nan_cnt != nullptr ? a.nan_cnt + q0 : nullptr;
      p.class_counts = a.class_counts;
      p.dxl = a.dxl;
      p.pwi = a.pwi;
      p.partial = fused ? a.partial + (int64_t)q0 * nb * 2 : nullptr;
      p.ranked_cls = fused ? a.ranked_cls + (int64_t)q0 * (nb + 1) * kMaxClasses : nullptr;
      p.n = a.n;
      p.n_gallery = a.n_gallery;
      p.n_classes = a.n_classes;
      p.idx_out = a.idx_out != nullptr ? a.idx_out + (int64_t)q0 * a.ldi : nullptr;
      p.ldi = a.ldi;
      p.keys_out = a.keys_out != nullptr ? a.keys_out + (int64_t)q0 * a.ldko : nullptr;
      p.ldko = a.ldko;
      dim3 grid((unsigned)nb, (unsigned)qn);
      switch (fused ? (a.n_classes + 3) / 4 : 0) {
        case 0: bucket_sort_kernel<0><<<grid, kSortThreads2, kSortSmem, st>>>(p); break;
        case 1: bucket_sort_kernel<1><<<grid, kSortThreads2, kSortSmem, st>>>(p); break;
        case 2: bucket_sort_kernel<2><<<grid, kSortThreads2, kSortSmem, st>>>(p); break;
        case 3: bucket_sort_kernel<3><<<grid, kSortThreads2, kSortSmem, st>>>(p); break;
        default: bucket_sort_kernel<4><<<grid, kSortThreads2, kSortSmem, st>>>(p); break;
      }
      DVL_LAUNCH_CHECK(c, "bucket_sort_kernel");
    }
  }
  return DVL_OK;
}

}  // namespace dvl
