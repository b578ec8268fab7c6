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
#include "dvl_common.cuh"
#include "dvl_internal.h"

namespace dvl {

namespace {

constexpr int kSortThreads = 256;
constexpr int kWarps = kSortThreads / 32;
constexpr int kItems = 16;                       // keys per thread
constexpr int kTile = kSortThreads * kItems;     // 4096 keys per CTA
constexpr int kRadix = 256;

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
// pass step 1: digit histogram of one tile
__global__ void __launch_bounds__(kSortThreads) tile_hist_kernel(const uint32_t* __restrict__ keys, int64_t ldk,
                                                                 int64_t n, int shift, int tiles,
                                                                 uint32_t* __restrict__ tile_hist) {
  __shared__ uint32_t hist[kRadix];
  const int q = blockIdx.y;
  const int t = blockIdx.x;
  hist[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t* seg = keys + (int64_t)q * ldk;
  const int64_t base = (int64_t)t * kTile;
  // ldk is a multiple of 4 and base a multiple of 4096: 16-byte loads are aligned
  if (base + kTile <= n) {
#pragma unroll
    for (int r = 0; r < kItems / 4; ++r) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(seg + base) + r * kSortThreads + threadIdx.x);
      atomicAdd(&hist[(v.x >> shift) & 0xFF], 1u);
      atomicAdd(&hist[(v.y >> shift) & 0xFF], 1u);
      atomicAdd(&hist[(v.z >> shift) & 0xFF], 1u);
      atomicAdd(&hist[(v.w >> shift) & 0xFF], 1u);
    }
  } else {
    for (int64_t i = base + threadIdx.x; i < n; i += kSortThreads) atomicAdd(&hist[(__ldg(seg + i) >> shift) & 0xFF], 1u);
  }
  __syncthreads();
  tile_hist[((int64_t)q * tiles + t) * kRadix + threadIdx.x] = hist[threadIdx.x];
}

// pass step 2: per segment, exclusive scan in (digit major, tile minor) order, in place
__global__ void __launch_bounds__(kRadix) seg_scan_kernel(uint32_t* __restrict__ tile_hist, int tiles) {
  __shared__ uint32_t warp_tot[kRadix / 32];
  const int q = blockIdx.x;
  const int d = threadIdx.x;
  uint32_t* h = tile_hist + (int64_t)q * tiles * kRadix;
  uint32_t total = 0;
  for (int t = 0; t < tiles; ++t) total += h[t * kRadix + d];
  // exclusive scan of the 256 digit totals
  uint32_t incl = total;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t o = __shfl_up_sync(0xffffffffu, incl, off);
    if ((d & 31) >= off) incl += o;
  }
  if ((d & 31) == 31) warp_tot[d >> 5] = incl;
  __syncthreads();
  uint32_t warp_base = 0;
  for (int w = 0; w < (d >> 5); ++w) warp_base += warp_tot[w];
  uint32_t run = warp_base + incl - total;
  for (int t = 0; t < tiles; ++t) {
    const uint32_t c = h[t * kRadix + d];
    h[t * kRadix + d] = run;
    run += c;
  }
}

// pass step 3: stable scatter of one tile.
//   phase 0  all 16 key (and payload) loads of a thread are issued before anything waits on them
//   phase 1  warp-synchronous stable ranking: match.any groups equal digits, the lowest lane of a
//            group reserves the group's slots in the warp's digit counter
//   phase 2  digit totals -> exclusive bases (inside the tile, and in the output segment)
//   phase 3  keys and payloads are placed at their tile-sorted position in shared memory
//   phase 4  the tile is written out in sorted order: each digit run is a contiguous stretch of the
//            output, so neighbouring threads store neighbouring addresses
template <bool kFirst, bool kHasIdx>
__global__ void __launch_bounds__(kSortThreads, 3)
tile_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint8_t* __restrict__ labs_in,
                    const int32_t* __restrict__ idx_in, const uint8_t* __restrict__ gallery_labels,
                    uint32_t* __restrict__ keys_out, uint8_t* __restrict__ labs_out, int32_t* __restrict__ idx_out,
                    int64_t ldk, int64_t n, int shift, int tiles, const uint32_t* __restrict__ tile_offs) {
  __shared__ uint32_t cnt[kWarps][kRadix];      // per-warp digit counts, then exclusive bases over warps
  __shared__ uint32_t dig_base[kRadix];         // first tile-sorted position of each digit
  __shared__ uint32_t out_off[kRadix];          // output position of a digit run minus dig_base
  __shared__ uint32_t warp_tot[kWarps];
  __shared__ uint32_t s_keys[kTile];
  __shared__ uint8_t s_labs[kTile];
  __shared__ int32_t s_idx[kHasIdx ? kTile : 1];
  const int q = blockIdx.y;
  const int t = blockIdx.x;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t lt_mask = (1u << lane) - 1u;
  for (int i = threadIdx.x; i < kWarps * kRadix; i += kSortThreads) (&cnt[0][0])[i] = 0;

  const int64_t seg_off = (int64_t)q * ldk;
  // warp-striped: warp w owns keys [w*512, w*512+512) of the tile, 16 rounds of 32 consecutive keys
  const int64_t warp_base = (int64_t)t * kTile + warp * (kItems * 32);
  uint32_t key[kItems];
  uint32_t lab4[kItems / 4];                    // four labels per register
  int32_t idx[kHasIdx ? kItems : 1];
#pragma unroll
  for (int r = 0; r < kItems / 4; ++r) lab4[r] = 0u;
#pragma unroll
  for (int r = 0; r < kItems; ++r) {
    const int64_t pos = warp_base + r * 32 + lane;
    const bool ok = pos < n;
    key[r] = ok ? __ldg(keys_in + seg_off + pos) : 0xFFFFFFFFu;
    const uint32_t l = ok ? (uint32_t)(kFirst ? __ldg(gallery_labels + pos) : __ldg(labs_in + seg_off + pos)) : 0u;
    lab4[r >> 2] |= l << (8 * (r & 3));
    if (kHasIdx) idx[r] = ok ? (kFirst ? (int32_t)pos : __ldg(idx_in + seg_off + pos)) : 0;
  }
  __syncthreads();

  uint32_t rank2[kItems / 2];                   // two 16-bit ranks per register
#pragma unroll
  for (int r = 0; r < kItems / 2; ++r) rank2[r] = 0u;
  // Lanes holding the same digit: eight ballots (one per digit bit) ANDed together -- the MATCH.ANY
  // instruction computes the same mask but its latency/throughput made it the kernel's bottleneck
  // (ncu: 59 % of stall samples on its consumers).  Lanes past the end of the segment vote as their own
  // group (digit bit 8).
#pragma unroll
  for (int r = 0; r < kItems; ++r) {
    const bool ok = warp_base + r * 32 + lane < n;
    const uint32_t digit = (key[r] >> shift) & 0xFFu;
    uint32_t peers = __ballot_sync(0xffffffffu, ok);
    if (!ok) peers = ~peers;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const uint32_t vote = __ballot_sync(0xffffffffu, (digit >> b) & 1u);
      peers &= ((digit >> b) & 1u) ? vote : ~vote;
    }
    uint32_t before = 0;
    if (ok) before = cnt[warp][digit];            // every lane of the group reads the same counter ...
    __syncwarp();
    if (ok && (peers & lt_mask) == 0u) cnt[warp][digit] = before + __popc(peers);   // ... its lowest lane bumps it
    rank2[r >> 1] |= (before + __popc(peers & lt_mask)) << (16 * (r & 1));
    __syncwarp();
  }
  __syncthreads();
  {
    // thread d: exclusive prefix of digit d over the warps, and the digit's total
    const int d = threadIdx.x;
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      const uint32_t c = cnt[w][d];
      cnt[w][d] = run;
      run += c;
    }
    // exclusive scan of the 256 totals
    uint32_t incl = run;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += o;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    uint32_t base = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) base += (w < warp) ? warp_tot[w] : 0u;
    const uint32_t excl = base + incl - run;
    dig_base[d] = excl;
    out_off[d] = tile_offs[((int64_t)q * tiles + t) * kRadix + d] - excl;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kItems; ++r) {
    if (warp_base + r * 32 + lane < n) {
      const uint32_t digit = (key[r] >> shift) & 0xFFu;
      const uint32_t p = dig_base[digit] + cnt[warp][digit] + ((rank2[r >> 1] >> (16 * (r & 1))) & 0xFFFFu);
      s_keys[p] = key[r];
      s_labs[p] = (uint8_t)(lab4[r >> 2] >> (8 * (r & 3)));
      if (kHasIdx) s_idx[p] = idx[r];
    }
  }
  __syncthreads();
  const int64_t left = n - (int64_t)t * kTile;
  const int count = left < kTile ? (int)left : kTile;
#pragma unroll
  for (int r = 0; r < kItems; ++r) {
    const int p = r * kSortThreads + threadIdx.x;
    if (p < count) {
      const uint32_t k = s_keys[p];
      const int64_t dst = seg_off + out_off[(k >> shift) & 0xFFu] + p;
      keys_out[dst] = k;
      labs_out[dst] = s_labs[p];
      if (kHasIdx) idx_out[dst] = s_idx[p];
    }
  }
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
  const bool has_idx = a.idxs[0] != nullptr;
  if (a.ldk % 4 != 0) return set_error(c, DVL_ERR_INVALID, "sort pitch must be a multiple of 4");
  DVL_TRY(ensure(c, c->tile_hist, (size_t)a.q * tiles * kRadix * sizeof(uint32_t)));
  uint32_t* th = (uint32_t*)c->tile_hist.ptr;
  int cur = 0;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = pass * 8;
    for (int q0 = 0; q0 < a.q; q0 += 65535) {
      const int qn = (a.q - q0) < 65535 ? (a.q - q0) : 65535;
      const int64_t off = (int64_t)q0 * a.ldk;
      uint32_t* th_q = th + (int64_t)q0 * tiles * kRadix;
      dim3 grid((unsigned)tiles, (unsigned)qn);
      {
        StageScope scope(c, DVL_STAGE_SORT_HIST, st);
        tile_hist_kernel<<<grid, kSortThreads, 0, st>>>(a.keys[cur] + off, a.ldk, a.n, shift, tiles, th_q);
        DVL_LAUNCH_CHECK(c, "tile_hist_kernel");
      }
      {
        StageScope scope(c, DVL_STAGE_SORT_SCAN, st);
        seg_scan_kernel<<<(unsigned)qn, kRadix, 0, st>>>(th_q, tiles);
        DVL_LAUNCH_CHECK(c, "seg_scan_kernel");
      }
      StageScope scope(c, DVL_STAGE_SORT_SCATTER, st);
      const uint32_t* ki = a.keys[cur] + off;
      const uint8_t* li = a.labs[cur] + off;
      const int32_t* ii = has_idx ? a.idxs[cur] + off : nullptr;
      uint32_t* ko = a.keys[cur ^ 1] + off;
      uint8_t* lo = a.labs[cur ^ 1] + off;
      int32_t* io = has_idx ? a.idxs[cur ^ 1] + off : nullptr;
      if (pass == 0) {
        if (has_idx)
          tile_scatter_kernel<true, true><<<grid, kSortThreads, 0, st>>>(ki, li, ii, a.gallery_labels, ko, lo, io,
                                                                         a.ldk, a.n, shift, tiles, th_q);
        else
          tile_scatter_kernel<true, false><<<grid, kSortThreads, 0, st>>>(ki, li, ii, a.gallery_labels, ko, lo, io,
                                                                          a.ldk, a.n, shift, tiles, th_q);
      } else {
        if (has_idx)
          tile_scatter_kernel<false, true><<<grid, kSortThreads, 0, st>>>(ki, li, ii, a.gallery_labels, ko, lo, io,
                                                                          a.ldk, a.n, shift, tiles, th_q);
        else
          tile_scatter_kernel<false, false><<<grid, kSortThreads, 0, st>>>(ki, li, ii, a.gallery_labels, ko, lo, io,
                                                                           a.ldk, a.n, shift, tiles, th_q);
      }
      DVL_LAUNCH_CHECK(c, "tile_scatter_kernel");
    }
    cur ^= 1;
  }
  *out_buf = cur;
  return DVL_OK;
}

}  // namespace dvl
