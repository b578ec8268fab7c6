// Ranking stage, full-sort route (measuring_bias.py:29,57 ->
// pandas nlargest(keep="all") which for top_n >= N is a stable descending sort):
// Q independent segments (one per query) of N 32-bit ranking keys are sorted ascending by a
// stable LSD radix sort, 8-bit digits, four passes.  Stability + gallery order in the input
// gives the reference's tie rule (equal score -> lowest index first) without carrying the index.
// The payload is the gallery label (one byte, fetched by position in the first pass) and,
// only when the caller asks for ranked indices, the position.
//
// Per pass:  tile digit histograms -> per-segment exclusive scan over (digit, tile) ->
//            stable scatter (warp match/ballot ranking inside 4096-key tiles).
// HBM-bound byte work: coalesced 16-byte key loads, one CTA per (tile, segment).
#include <stdlib.h>

#include "dvl_common.cuh"
#include "dvl_internal.h"

namespace dvl {

namespace {

constexpr int kSortThreads = 256;
constexpr int kWarps = kSortThreads / 32;
constexpr int kItems = 16;                       // keys per thread
constexpr int kTile = kSortThreads * kItems;     // 4096 keys per CTA
constexpr int kMaxBits = 8;

// ---------------------------------------------------------------------------------------
// float scores -> ranking keys (stage entry point dvl_rank_and_score)
__global__ void __launch_bounds__(256) scores_to_keys_kernel(const float* __restrict__ scores, int64_t lds,
                                                             uint32_t* __restrict__ keys, int64_t ldk, int64_t n) {
  const int q = blockIdx.y;
  const float* s = scores + (int64_t)q * lds;
  uint32_t* k = keys + (int64_t)q * ldk;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    k[i] = score_to_key(__ldg(s + i));
}

// ---------------------------------------------------------------------------------------
// pass step 1: digit histogram of one tile.  (Two neighbouring tiles per CTA with all their loads requested
// before the first shared-memory atomic: 0.136 instead of 0.130 ms for the four launches on config 2 -- not kept.)
template <int BITS, bool kCountNan>
__global__ void __launch_bounds__(kSortThreads) tile_hist_kernel(const uint32_t* __restrict__ keys, int64_t ldk,
                                                                 int64_t n, int shift, int tiles,
                                                                 uint32_t* __restrict__ tile_hist,
                                                                 uint32_t* __restrict__ nan_cnt) {
  constexpr int RADIX = 1 << BITS;
  constexpr uint32_t MASK = RADIX - 1;
  __shared__ uint32_t hist[RADIX];
  const int q = blockIdx.y;
  const int t = blockIdx.x;
  for (int i = threadIdx.x; i < RADIX; i += kSortThreads) hist[i] = 0;
  __syncthreads();
  const uint32_t* seg = keys + (int64_t)q * ldk;
  const int64_t base = (int64_t)t * kTile;
  // ldk is a multiple of 4 and base a multiple of 4096: 16-byte loads are aligned
  uint32_t nans = 0;                             // first pass only: rows whose score is NaN (key 0xFFFFFFFF)
  if (base + kTile <= n) {
#pragma unroll
    for (int r = 0; r < kItems / 4; ++r) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(seg + base) + r * kSortThreads + threadIdx.x);
      atomicAdd(&hist[(v.x >> shift) & MASK], 1u);
      atomicAdd(&hist[(v.y >> shift) & MASK], 1u);
      atomicAdd(&hist[(v.z >> shift) & MASK], 1u);
      atomicAdd(&hist[(v.w >> shift) & MASK], 1u);
      if (kCountNan && max(max(v.x, v.y), max(v.z, v.w)) == kNanKey)     // rare: one compare per four keys
        nans += (v.x == kNanKey) + (v.y == kNanKey) + (v.z == kNanKey) + (v.w == kNanKey);
    }
  } else {
    for (int64_t i = base + threadIdx.x; i < n; i += kSortThreads) {
      const uint32_t k = __ldg(seg + i);
      atomicAdd(&hist[(k >> shift) & MASK], 1u);
      if (kCountNan) nans += (k == kNanKey);
    }
  }
  if (kCountNan && nan_cnt != nullptr && nans != 0u) atomicAdd(nan_cnt + q, nans);   // NaN scores are rare
  __syncthreads();
  uint32_t* out = tile_hist + ((int64_t)q * tiles + t) * RADIX;
  for (int i = threadIdx.x; i < RADIX; i += kSortThreads) out[i] = hist[i];
}

// pass step 2: per segment, exclusive scan in (digit major, tile minor) order, in place
template <int BITS>
__global__ void __launch_bounds__(kSortThreads) seg_scan_kernel(uint32_t* __restrict__ tile_hist, int tiles) {
  constexpr int RADIX = 1 << BITS;
  constexpr int SLABS = RADIX / kSortThreads;     // digits d = slab*256 + tid
  __shared__ uint32_t warp_tot[kWarps];
  __shared__ uint32_t carry;
  const int q = blockIdx.x;
  const int tid = threadIdx.x;
  uint32_t* h = tile_hist + (int64_t)q * tiles * RADIX;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int slab = 0; slab < SLABS; ++slab) {
    const int d = slab * kSortThreads + tid;
    uint32_t total = 0;
    for (int t = 0; t < tiles; ++t) total += h[t * RADIX + d];
    uint32_t incl = total;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, incl, off);
      if ((tid & 31) >= off) incl += o;
    }
    if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
    __syncthreads();
    uint32_t base = carry;
    for (int w = 0; w < (tid >> 5); ++w) base += warp_tot[w];
    uint32_t run = base + incl - total;
    for (int t = 0; t < tiles; ++t) {
      const uint32_t c = h[t * RADIX + d];
      h[t * RADIX + d] = run;
      run += c;
    }
    __syncthreads();
    if (tid == kSortThreads - 1) carry = base + incl;
    __syncthreads();
  }
}

// Lanes of the warp holding the same digit as this lane.
//   kRank 0: BITS ballots.  Per bit: bit test -> predicate, VOTE, SEL (all-ones if my bit is set),
//            one LOP3 accumulating the lanes that differ (vote ^ mine) -- four instructions; written in
//            PTX because the compiler's own rendering of `bit ? vote : ~vote` took seven.
//   kRank 1: one MATCH.ANY per key (kept for measurements: 47 instead of 72 instructions per key, but
//            0.23 instead of 0.148 ms per pass on config 2 -- MATCH.ANY issues far below one per clock).
template <int BITS, int kRank>
__device__ __forceinline__ uint32_t digit_peers(uint32_t digit) {
  if (kRank == 1) return __match_any_sync(0xffffffffu, digit);
  uint32_t differ = 0u;
#pragma unroll
  for (int b = 0; b < BITS; ++b) {
    uint32_t vote, mine;
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t"
        "and.b32 t, %2, %3;\n\t"
        "setp.ne.u32 p, t, 0;\n\t"
        "vote.sync.ballot.b32 %0, p, 0xffffffff;\n\t"
        "selp.b32 %1, 0xffffffff, 0, p;\n\t}"
        : "=r"(vote), "=r"(mine)
        : "r"(digit), "r"(1u << b));
    differ |= vote ^ mine;
  }
  return ~differ;
}

// pass step 3: stable scatter of one tile.
//   phase 0  all 16 key (and payload) loads of a thread are issued before anything waits on them
//   phase 1  warp-synchronous stable ranking: BITS ballots group the lanes holding equal digits, the
//            lowest lane of a group reserves the group's slots in the warp's digit counter
//   phase 2  digit totals -> exclusive bases (inside the tile, and in the output segment)
//   phase 3  keys and payloads are placed at their tile-sorted position in shared memory
//   phase 4  the tile is written out in sorted order: each digit run is a contiguous stretch of the
//            output, so neighbouring threads store neighbouring addresses
template <int BITS, bool kHasIdx>
constexpr size_t scatter_smem_bytes() {
  return (size_t)kWarps * (1 << BITS) * 2 + (size_t)(1 << BITS) * 4 + kTile * 4 + kTile +
         (kHasIdx ? kTile * 4 : 0) + 64;
}

template <int BITS, int SHIFT, int kRank, bool kFirst, bool kHasIdx, bool kFull>
__device__ __forceinline__ void
tile_scatter_body(const uint32_t* __restrict__ keys_in, const uint8_t* __restrict__ labs_in,
                  const int32_t* __restrict__ idx_in, const uint8_t* __restrict__ gallery_labels,
                  uint32_t* __restrict__ keys_out, uint8_t* __restrict__ labs_out, int32_t* __restrict__ idx_out,
                  int64_t ldk, int64_t n, int tiles, const uint32_t* __restrict__ tile_offs, int write_keys) {
  constexpr int RADIX = 1 << BITS;
  constexpr uint32_t MASK = RADIX - 1;
  constexpr int DPT = RADIX / kSortThreads;      // digits per thread in phase 2
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  uint32_t* out_off = reinterpret_cast<uint32_t*>(smem_dyn);              // [RADIX] output position of a run - dig_base
  uint32_t* s_keys = out_off + RADIX;                                     // [kTile]
  int32_t* s_idx = reinterpret_cast<int32_t*>(s_keys + kTile);            // [kTile] if kHasIdx
  uint16_t* cnt = reinterpret_cast<uint16_t*>(s_idx + (kHasIdx ? kTile : 0));   // [kWarps][RADIX]
  uint8_t* s_labs = reinterpret_cast<uint8_t*>(cnt + kWarps * RADIX);     // [kTile]
  __shared__ uint32_t warp_tot[kWarps];
  const int q = blockIdx.y;
  const int t = blockIdx.x;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t lt_mask = (1u << lane) - 1u;
  for (int i = threadIdx.x; i < kWarps * RADIX / 2; i += kSortThreads) reinterpret_cast<uint32_t*>(cnt)[i] = 0;

  const int64_t seg_off = (int64_t)q * ldk;
  // keys of this tile that exist (a full tile needs no bounds checks at all: kFull)
  const int64_t left = n - (int64_t)t * kTile;
  const int count = (kFull || left >= kTile) ? kTile : (int)left;
  // warp-striped: warp w owns keys [w*512, w*512+512) of the tile, 16 rounds of 32 consecutive keys
  const int local_base = warp * (kItems * 32) + lane;
  const int64_t warp_base = (int64_t)t * kTile + warp * (kItems * 32);
  const uint32_t* kp = keys_in + seg_off + warp_base + lane;
  const uint8_t* lp = (kFirst ? gallery_labels : labs_in + seg_off) + warp_base + lane;
  uint32_t key[kItems];
  uint32_t lab4[kItems / 4];                    // four labels per register
  int32_t idx[kHasIdx ? kItems : 1];
#pragma unroll
  for (int r = 0; r < kItems / 4; ++r) lab4[r] = 0u;
#pragma unroll
  for (int r = 0; r < kItems; ++r) {
    const bool ok = kFull || (local_base + r * 32 < count);
    key[r] = ok ? __ldg(kp + r * 32) : 0xFFFFFFFFu;
    const uint32_t l = ok ? (uint32_t)__ldg(lp + r * 32) : 0u;
    lab4[r >> 2] |= l << (8 * (r & 3));
    if (kHasIdx) idx[r] = ok ? (kFirst ? (int32_t)(warp_base + r * 32 + lane) : __ldg(idx_in + seg_off + warp_base + r * 32 + lane)) : 0;
  }
  // this tile's output offsets (used after the ranking): requested now, so their latency is hidden
  uint32_t my_offs[DPT];
  {
    const uint32_t* offs = tile_offs + ((int64_t)q * tiles + t) * RADIX + threadIdx.x * DPT;
#pragma unroll
    for (int j = 0; j < DPT; ++j) my_offs[j] = __ldg(offs + j);
  }
  __syncthreads();

  uint32_t rank2[kItems / 2];                   // two 16-bit ranks per register
#pragma unroll
  for (int r = 0; r < kItems / 2; ++r) rank2[r] = 0u;
  uint16_t* my_cnt = cnt + warp * RADIX;
  // Stable ranking inside the warp: the lanes holding the same digit form a group; every lane of the
  // group reads the warp's counter of the digit, the lowest lane bumps it by the group size.
  // Lanes past the end of the segment carry an extra digit bit, so they form their own group.
  constexpr int kBatch = (kRank == 1) ? 8 : 1;  // MATCH.ANY has a long latency: issue a batch before use
#pragma unroll
  for (int r0 = 0; r0 < kItems; r0 += kBatch) {
    uint32_t peers_b[kBatch];
#pragma unroll
    for (int j = 0; j < kBatch; ++j) {
      const int r = r0 + j;
      const bool ok = kFull || (local_base + r * 32 < count);
      const uint32_t digit = (key[r] >> SHIFT) & MASK;
      if (kRank == 1) {
        peers_b[j] = digit_peers<BITS, 1>(kFull ? digit : (digit | (ok ? 0u : (uint32_t)RADIX)));
      } else {
        uint32_t peers = digit_peers<BITS, 0>(digit);
        if (!kFull) {
          const uint32_t live = __ballot_sync(0xffffffffu, ok);
          peers &= ok ? live : ~live;
        }
        peers_b[j] = peers;
      }
    }
#pragma unroll
    for (int j = 0; j < kBatch; ++j) {
      const int r = r0 + j;
      const bool ok = kFull || (local_base + r * 32 < count);
      const uint32_t digit = (key[r] >> SHIFT) & MASK;
      const uint32_t peers = peers_b[j];
      uint32_t before = 0;
      if (ok) before = my_cnt[digit];
      __syncwarp();
      if (ok && (peers & lt_mask) == 0u) my_cnt[digit] = (uint16_t)(before + __popc(peers));
      rank2[r >> 1] |= (before + __popc(peers & lt_mask)) << (16 * (r & 1));
      __syncwarp();
    }
  }
  __syncthreads();
  {
    // thread owns DPT consecutive digits: exclusive prefix over the warps, digit totals, then the
    // exclusive scan of the totals across the block.  Afterwards cnt[w][d] = first tile position of
    // warp w's keys with digit d, out_off[d] = output position of digit d's run minus its tile position.
    uint32_t tot[DPT];
    uint32_t sum = 0;
#pragma unroll
    for (int j = 0; j < DPT; ++j) {
      const int d = threadIdx.x * DPT + j;
      uint32_t run = 0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) {
        run += cnt[w * RADIX + d];
      }
      tot[j] = run;
      sum += run;
    }
    uint32_t incl = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += o;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    uint32_t run = incl - sum;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) run += (w < warp) ? warp_tot[w] : 0u;
#pragma unroll
    for (int j = 0; j < DPT; ++j) {
      const int d = threadIdx.x * DPT + j;
      out_off[d] = my_offs[j] - run;
      uint32_t pos = run;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) {
        const uint32_t c = cnt[w * RADIX + d];
        cnt[w * RADIX + d] = (uint16_t)pos;
        pos += c;
      }
      run += tot[j];
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kItems; ++r) {
    if (kFull || (local_base + r * 32 < count)) {
      const uint32_t digit = (key[r] >> SHIFT) & MASK;
      const uint32_t p = (uint32_t)my_cnt[digit] + ((rank2[r >> 1] >> (16 * (r & 1))) & 0xFFFFu);
      s_keys[p] = key[r];
      s_labs[p] = (uint8_t)(lab4[r >> 2] >> (8 * (r & 3)));
      if (kHasIdx) s_idx[p] = idx[r];
    }
  }
  __syncthreads();
  uint32_t* ko = keys_out + seg_off;
  uint8_t* lo = labs_out + seg_off;
#pragma unroll
  for (int r = 0; r < kItems; ++r) {
    const int p = r * kSortThreads + threadIdx.x;
    if (kFull || p < count) {
      const uint32_t k = s_keys[p];
      const uint32_t dst = out_off[(k >> SHIFT) & MASK] + (uint32_t)p;
      if (write_keys) ko[dst] = k;               // uniform: the last pass of a whole-segment ranking skips it
      lo[dst] = s_labs[p];
      if (kHasIdx) idx_out[seg_off + dst] = s_idx[p];
    }
  }
}

template <int BITS, int SHIFT, int kRank, bool kFirst, bool kHasIdx>
__global__ void __launch_bounds__(kSortThreads, 4)   // 64 registers: 32 of the SM's 64 warp slots
tile_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint8_t* __restrict__ labs_in,
                    const int32_t* __restrict__ idx_in, const uint8_t* __restrict__ gallery_labels,
                    uint32_t* __restrict__ keys_out, uint8_t* __restrict__ labs_out, int32_t* __restrict__ idx_out,
                    int64_t ldk, int64_t n, int tiles, const uint32_t* __restrict__ tile_offs, int write_keys) {
  if ((int64_t)(blockIdx.x + 1) * kTile <= n)    // all tiles of a segment but (at most) the last
    tile_scatter_body<BITS, SHIFT, kRank, kFirst, kHasIdx, true>(keys_in, labs_in, idx_in, gallery_labels, keys_out,
                                                                 labs_out, idx_out, ldk, n, tiles, tile_offs, write_keys);
  else
    tile_scatter_body<BITS, SHIFT, kRank, kFirst, kHasIdx, false>(keys_in, labs_in, idx_in, gallery_labels, keys_out,
                                                                  labs_out, idx_out, ldk, n, tiles, tile_offs, write_keys);
}

int sort_rank_mode() {
  static const int mode = [] {
    const char* e = getenv("DVL_SORT_RANK");
    return (e != nullptr && e[0] == '1') ? 1 : 0;
  }();
  return mode;
}

template <int BITS, int SHIFT, int kRank, bool kFirst>
int run_pass(Ctx* c, const SortArgs& a, int cur, int tiles, uint32_t* th, cudaStream_t st) {
  constexpr int RADIX = 1 << BITS;
  const bool has_idx = a.idxs[0] != nullptr;
  for (int q0 = 0; q0 < a.q; q0 += 65535) {
    const int qn = (a.q - q0) < 65535 ? (a.q - q0) : 65535;
    const int64_t off = (int64_t)q0 * a.ldk;
    uint32_t* th_q = th + (int64_t)q0 * tiles * RADIX;
    dim3 grid((unsigned)tiles, (unsigned)qn);
    {
      StageScope scope(c, DVL_STAGE_SORT_HIST, st);
      tile_hist_kernel<BITS, kFirst><<<grid, kSortThreads, 0, st>>>(a.keys[cur] + off, a.ldk, a.n, SHIFT, tiles, th_q,
                                                                    a.nan_cnt != nullptr ? a.nan_cnt + q0 : nullptr);
      DVL_LAUNCH_CHECK(c, "tile_hist_kernel");
    }
    {
      StageScope scope(c, DVL_STAGE_SORT_SCAN, st);
      seg_scan_kernel<BITS><<<(unsigned)qn, kSortThreads, 0, st>>>(th_q, tiles);
      DVL_LAUNCH_CHECK(c, "seg_scan_kernel");
    }
    StageScope scope(c, DVL_STAGE_SORT_SCATTER, st);
    const uint32_t* ki = a.keys[cur] + off;
    const uint8_t* li = a.labs[cur] + off;
    const int32_t* ii = has_idx ? a.idxs[cur] + off : nullptr;
    uint32_t* ko = a.keys[cur ^ 1] + off;
    uint8_t* lo = a.labs[cur ^ 1] + off;
    int32_t* io = has_idx ? a.idxs[cur ^ 1] + off : nullptr;
#define DVL_SCATTER(IDX)                                                                                          \
  do {                                                                                                            \
    auto kern = tile_scatter_kernel<BITS, SHIFT, kRank, kFirst, IDX>;                                             \
    constexpr size_t smem = scatter_smem_bytes<BITS, IDX>();                                                      \
    static_assert(smem <= 48 * 1024, "above 48 KB the kernel needs cudaFuncAttributeMaxDynamicSharedMemorySize"); \
    kern<<<grid, kSortThreads, smem, st>>>(ki, li, ii, a.gallery_labels, ko, lo, io, a.ldk, a.n, tiles, th_q,     \
                                           (SHIFT + BITS < 32 || a.nan_cnt == nullptr) ? 1 : 0);                   \
  } while (0)
    if (has_idx) DVL_SCATTER(true); else DVL_SCATTER(false);
#undef DVL_SCATTER
    DVL_LAUNCH_CHECK(c, "tile_scatter_kernel");
  }
  return DVL_OK;
}

template <int kRank>
int sort_passes(Ctx* c, const SortArgs& a, int tiles, uint32_t* th, cudaStream_t st) {
  // 32-bit keys in four LSD passes of 8 bits.  Measured alternatives (config 2):
  //  * 11+11+10 bits in three passes: 0.33 vs 0.23 ms per scatter pass -- two-key runs, 2048-bin scans
  //    and tile histograms half the size of the keys;
  //  * one up-front digit histogram + decoupled look-back between the tiles of a pass instead of the
  //    per-pass histogram and scan kernels: saves 0.11 ms of histogram/scan time but the look-back adds
  //    0.09-0.19 ms to the scatter (a tile lives ~7 us; a published count takes ~1.5 us to become
  //    visible behind the SM's own scatter stores), so it was not kept;
  //  * key and label staged as one 8-byte shared-memory entry (one scattered 64-bit store instead of a
  //    32-bit and an 8-bit one, 5.4 instead of 7.3 wavefronts per 32 keys): 0.521 -> 0.518 ms, not kept.
  //    Source-level profile of the kernel (ncu, second pass): 77 warp instructions per 32 keys, 30 % of
  //    them the LOP3s of the ballot matching; stall samples are spread thin -- 17 % on the first use of
  //    the loaded keys, 11 % on the digit-counter reads, no single hot spot;
  //  * persistent CTAs that prefetch the next tile into a shared-memory staging buffer with
  //    cp.async.bulk while the current one is ranked: 0.63 vs 0.52 ms for the four passes -- the kernel
  //    is bound by shared-memory wavefronts (half of them bank conflicts of the digit-indexed counters
  //    and the scattered tile placement), and staging adds 12 % more of them.
  DVL_TRY((run_pass<8, 0, kRank, true>(c, a, 0, tiles, th, st)));
  DVL_TRY((run_pass<8, 8, kRank, false>(c, a, 1, tiles, th, st)));
  DVL_TRY((run_pass<8, 16, kRank, false>(c, a, 0, tiles, th, st)));
  DVL_TRY((run_pass<8, 24, kRank, false>(c, a, 1, tiles, th, st)));
  return DVL_OK;
}

}  // namespace

int launch_scores_to_keys(Ctx* c, const float* scores, int64_t lds, uint32_t* keys, int64_t ldk, int64_t n, int q,
                          cudaStream_t st) {
  int64_t bx = (n + 256 * 8 - 1) / (256 * 8);
  if (bx < 1) bx = 1;
  if (bx > 4096) bx = 4096;
  for (int q0 = 0; q0 < q; q0 += 65535) {
    const int qn = (q - q0) < 65535 ? (q - q0) : 65535;
    dim3 grid((unsigned)bx, (unsigned)qn);
    scores_to_keys_kernel<<<grid, 256, 0, st>>>(scores + (int64_t)q0 * lds, lds, keys + (int64_t)q0 * ldk, ldk, n);
    DVL_LAUNCH_CHECK(c, "scores_to_keys_kernel");
  }
  return DVL_OK;
}

int launch_segmented_sort(Ctx* c, const SortArgs& a, int* out_buf, cudaStream_t st) {
  const int tiles = (int)((a.n + kTile - 1) / kTile);
  if (a.ldk % 4 != 0) return set_error(c, DVL_ERR_INVALID, "sort pitch must be a multiple of 4");
  DVL_TRY(ensure(c, c->tile_hist, (size_t)a.q * tiles * (1 << kMaxBits) * sizeof(uint32_t)));
  uint32_t* th = (uint32_t*)c->tile_hist.ptr;
  if (a.nan_cnt != nullptr) DVL_CUDA(c, cudaMemsetAsync(a.nan_cnt, 0, (size_t)a.q * sizeof(uint32_t), st));
  if (sort_rank_mode() == 1)
    DVL_TRY(sort_passes<1>(c, a, tiles, th, st));
  else
    DVL_TRY(sort_passes<0>(c, a, tiles, th, st));
  *out_buf = 0;
  return DVL_OK;
}

}  // namespace dvl
