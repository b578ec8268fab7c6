// Similarity stage (measuring_bias.py:114 for all queries at once):
//   S[q, i] = sum_k T[q, k] * G[i, k]
// as ONE tcgen05 GEMM instead of Q fp32 mat-vecs.  sm_100a only.
//
// Mapping.  UMMA M (TMEM lanes)  = 128 queries   -> an epilogue thread owns ONE query row
//           UMMA N (TMEM columns) = 256 gallery rows
//           K-blocks of 64 bf16 (= one 128-byte swizzle row), 4 x tcgen05.mma (K=16) each.
// Both operands are K-major in memory (row-major [rows, D]), loaded by TMA with the
// 128-byte swizzle straight into the canonical UMMA layout; no transposes anywhere.
//
// Persistent, warp-specialised CTA (one per SM):
//   warp 0      TMA producer (one lane)      full/empty mbarrier ring, kStages deep
//   warp 1      TMEM allocator + MMA issuer (one lane), double-buffered fp32 accumulators
//               (2 x 256 TMEM columns) so the epilogue of tile t overlaps the MMAs of t+1
//   warps 2..9  epilogue (a warp may only read the TMEM lane quarter warp_id % 4: two warps per quarter,
//               one per half of the tile's 256 columns -- with four warps the epilogue of the scaled
//               fp32 modes took twice the MMA time of a tile and the tensor pipe idled half the time):
//               tcgen05.ld 32 lanes x 32 columns -> fused scale (L2 normalisation, 2^-e) ->
//               STORE : ranking key / fp32 score, 128-byte row segments to HBM (full-sort route)
//               SEED  : every (key, index, label) candidate of the first gallery chunk, dense
//               FILTER: streaming top-k -- one compare per similarity against the query's
//                       running threshold tau_q held in a register; the rare survivors are
//                       appended to the query's candidate list.  The similarity matrix never
//                       reaches HBM.
//
// fp32 modes: operands arrive as several 16-bit planes each, all product terms go into the same
// accumulator, the row factors (2^-e of the fp16 split, fused L2 normalisation) are applied in the
// epilogue.
//   fp16 pair (default, three terms): k-block major -- a pipeline stage holds BOTH planes of both
//     operands for one k-block (96 KB, two stages) and feeds 12 MMAs (h1*h0', h0*h1', h0*h0'), so every
//     operand byte is fetched from L2 once per tile.  Loading a stage per term instead (1.15 MB per tile)
//     left the kernel bound by the L2 -> shared-memory path (~54 B/clk/SM), not by the tensor pipe.
//     (Requesting the next tile's gallery block into L2 a tile ahead with cp.async.bulk.prefetch.tensor
//     was measured and not kept: no gain on configs 2 and 5, 20 % slower on config 3.)
//   bf16 triple (six terms): term major, one plane pair per 48 KB stage, four stages, smallest terms first.
//
// Tiles are ordered gallery-block major so that CTAs running at the same time share the
// gallery block (read from HBM once, then L2 hits) while the small query matrix stays
// L2-resident.
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>

#include "dvl_common.cuh"
#include "dvl_internal.h"

namespace dvl {

namespace {

constexpr int BM = 128;            // queries per tile
constexpr int BN = 256;            // gallery rows per tile
constexpr int BK = 64;             // bf16 per k-block (128 B)
constexpr int UMMA_K = 16;
constexpr int kStages = 4;
constexpr int kABytes = BM * BK * 2;   // 16 KB
constexpr int kBBytes = BN * BK * 2;   // 32 KB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kTmemCols = 512;         // 2 accumulator buffers x 256 columns
constexpr int kEpiWarps = 8;           // two per TMEM lane quarter, each takes one half of the 256 columns
constexpr int kEpiThreads = 32 * kEpiWarps;
constexpr int kNumThreads = 64 + kEpiThreads;   // + TMA producer warp + MMA warp
constexpr int kEpiScratchBytes = kEpiThreads * 33 * 4;   // FILTER: one 32-word row per epilogue thread
constexpr size_t kSmemBytes =
    1024 /*align slack*/ + (size_t)kStages * kStageBytes + 256 /*barriers*/ + kEpiScratchBytes;

// plane pairs (query plane, gallery plane), smallest product magnitude first
constexpr int kMaxTerms = 6;
constexpr int kTermsX3A[6] = {0, 1, 2, 0, 1, 0};   // three bf16 planes: i + j <= 2
constexpr int kTermsX3B[6] = {2, 1, 0, 1, 0, 0};
constexpr int kTermsX2A[3] = {1, 0, 0};            // two fp16 planes: h1*h0', h0*h1', h0*h0'
constexpr int kTermsX2B[3] = {0, 1, 0};

struct KernelParams {
  int mode;          // kGemmStore / kGemmSeed / kGemmFilter
  int64_t row0;      // first gallery row of this launch (multiple of 256)
  int64_t n;         // one past the last gallery row of this launch
  int q;
  int n_qtiles;
  int64_t n_tiles;
  int k_blocks;     // ceil(d / 64)
  int n_terms;      // 1 (one plane), 3 (fp16 pair) or 6 (bf16 triple)
  int n_planes;     // 16-bit planes per operand
  int fused_planes; // 1: fp16 pair, k-block major with both planes of both operands in one stage
  int term_a[kMaxTerms], term_b[kMaxTerms];
  uint32_t idesc;   // UMMA instruction descriptor (bf16 or fp16 operands)
  const float* qry_scale;
  const float* gal_scale;   // indexed by the physical gallery row: kernel row * gal_step
  int64_t gal_step;
  uint32_t* keys;
  int64_t ldk;
  float* scores;
  int64_t lds;
  // streaming top-k
  const float* tau;            // [q] running threshold (score); a similarity survives if s >= tau
  unsigned long long* cand;    // [q][cap] entries key<<32 | global index<<4 | label
  uint32_t* cand_cnt;          // [q]
  int64_t cap;
  const uint8_t* labels;       // [n] labels of this gallery shard
  int64_t index_base;          // global index of local row 0
  int probe_pass1_only;        // measurement only
};

// The label nibble is filled in by the final select pass (topk.cu): looking it up here would put a
// dependent global load on the epilogue's slow path.
__device__ __forceinline__ unsigned long long make_entry(float s, int64_t gidx) {
  return (static_cast<unsigned long long>(score_to_key(s)) << 32) | (static_cast<unsigned long long>(gidx) << 4);
}

// v[j] for a runtime j without touching memory: a five-level select tree (31 selects).  The FILTER epilogue
// needs it for its rare survivors; staging the 32 values of a chunk in shared memory instead (indexable, the
// first version) cost ~120 KB of shared-memory traffic per tile -- on the port the tensor cores' operand reads
// and the TMA writes already saturate.
__device__ __forceinline__ uint32_t pick32(const uint32_t (&v)[32], int j) {
  uint32_t a[16], b[8], c[4], d[2];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = (j & 16) ? v[16 + i] : v[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) b[i] = (j & 8) ? a[8 + i] : a[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i] = (j & 4) ? b[4 + i] : b[i];
#pragma unroll
  for (int i = 0; i < 2; ++i) d[i] = (j & 2) ? c[2 + i] : c[i];
  return (j & 1) ? d[1] : d[0];
}

// kPair: the kernel runs as clusters of two CTAs (one TPC) that drive tcgen05.mma.cta_group::2 together: the
// pair owns 256 queries x 256 gallery rows; each CTA loads its own 128 query rows and HALF of the gallery
// tile (the tensor cores read the other half from the peer's shared memory), accumulates its own 128 x 256
// block in its own TMEM and runs its own epilogue.  Per CTA and k-block 32 KB instead of 48 KB cross
// L2 -> shared memory for the same tensor work -- the single-CTA kernel is bound by exactly that path on
// the large-Q workloads (config 4: 590 KB per tile at ~54 B/clk/SM = the measured 1340 TFLOP/s) -- and the
// ring holds six stages instead of four.  The leader CTA (rank 0) issues every MMA; the full barriers it
// waits on live in its shared memory and collect the bytes of both CTAs' TMA loads; its commits arrive on
// both CTAs' empty / accumulator-full barriers; both epilogues arrive on the leader's accumulator-empty one.
template <bool kPair>
__global__ void __launch_bounds__(kNumThreads, 1)
sim_gemm_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_g,
                const KernelParams p) {
  constexpr int kBRows = kPair ? BN / 2 : BN;              // gallery rows of a tile this CTA loads
  constexpr int kBBytesCta = kBRows * BK * 2;              // 16 or 32 KB
  constexpr int kPlaneStage = kABytes + kBBytesCta;        // one plane pair: 32 or 48 KB
  constexpr int kRingStages = kStages * kStageBytes / kPlaneStage;   // 6 or 4 (192 KB ring either way)
  extern __shared__ uint8_t smem_raw[];
  // 128-byte-swizzle atoms need 1024-byte aligned tiles
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + kStages * kStageBytes;
  // barrier addresses (same offsets in both CTAs of a pair)
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 64u + 8u * s; };
  auto tfull_bar = [&](int a) { return bar_base + 128u + 8u * a; };
  auto tempty_bar = [&](int a) { return bar_base + 144u + 8u * a; };
  const uint32_t tmem_slot = bar_base + 160u;
  uint32_t* epi_scratch = reinterpret_cast<uint32_t*>(smem_raw + (bar_base + 256u - smem_u32(smem_raw)));
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = kPair ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0u;
  // work units: a CTA, or a CTA pair, walks the tiles with this stride
  const int64_t unit = kPair ? (int64_t)(blockIdx.x >> 1) : (int64_t)blockIdx.x;
  const int64_t n_units = kPair ? (int64_t)(gridDim.x >> 1) : (int64_t)gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kRingStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), kPair ? 2 * kEpiThreads : kEpiThreads);
    }
    fence_mbar_init();
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_g);
  }
  if (warp == 1) {
    if (kPair) {
      tmem_alloc_pair(tmem_slot, kTmemCols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();   // the peer's barriers are initialised, too
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // pipeline shape: stages of one plane pair, or half as many stages holding both planes of both operands
  const int iters_per_tile = p.fused_planes ? p.k_blocks : p.k_blocks * p.n_terms;
  const int n_stages = p.fused_planes ? kRingStages / 2 : kRingStages;
  const uint32_t stage_bytes = p.fused_planes ? 2u * kPlaneStage : (uint32_t)kPlaneStage;
  // a tile index t means (query tile t % n_qtiles, gallery block t / n_qtiles); in pair mode a query tile
  // is 256 rows and this CTA takes its half
  auto my_qt = [&](int64_t t) { return kPair ? (int)(t % p.n_qtiles) * 2 + (int)cta_rank : (int)(t % p.n_qtiles); };

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t t = unit; t < p.n_tiles; t += n_units) {
        const int qt = my_qt(t);
        const int64_t gb = t / p.n_qtiles;
        const int grow = (int)(p.row0 + gb * BN) + (int)cta_rank * kBRows;   // this CTA's part of the gallery block
        for (int it = 0; it < iters_per_tile; ++it) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          // the leader's barrier counts the bytes of both CTAs
          if (leader) mbar_arrive_expect_tx(full_bar(stage), kPair ? 2u * stage_bytes : stage_bytes);
          const uint32_t fb = kPair ? (full_bar(stage) & kPeerBitMask) : full_bar(stage);
          const uint32_t sa = smem_base + stage * stage_bytes;
          auto load = [&](uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2) {
            if (kPair) tma_load_3d_pair(dst, map, fb, c0, c1, c2);
            else tma_load_3d(dst, map, fb, c0, c1, c2);
          };
          if (p.fused_planes) {
            // [A plane 0][A plane 1][B plane 0][B plane 1] of k-block `it`
            load(sa, &map_q, it * BK, qt * BM, 0);
            load(sa + kABytes, &map_q, it * BK, qt * BM, 1);
            load(sa + 2 * kABytes, &map_g, it * BK, grow, 0);
            load(sa + 2 * kABytes + kBBytesCta, &map_g, it * BK, grow, 1);
          } else {
            const int term = it / p.k_blocks;
            const int kb = it - term * p.k_blocks;
            load(sa, &map_q, kb * BK, qt * BM, p.term_a[term]);
            load(sa + kABytes, &map_g, kb * BK, grow, p.term_b[term]);
          }
          if (++stage == n_stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader CTA only in pair mode) ------------------------------
    if (lane == 0 && leader) {
      const uint32_t idesc = p.idesc;
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t acc) {
        if (kPair) umma_f16_ss_pair(d, da, db, idesc, acc);
        else umma_bf16_ss(d, da, db, idesc, acc);
      };
      auto commit = [&](uint32_t bar) {
        if (kPair) umma_commit_pair(bar, (uint16_t)0x3);
        else umma_commit(bar);
      };
      for (int64_t t = unit; t < p.n_tiles; t += n_units) {
        mbar_wait(tempty_bar(as), aphase ^ 1u);   // the epilogue(s) have drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        for (int it = 0; it < iters_per_tile; ++it) {
          mbar_wait(full_bar(stage), phase);      // TMA bytes (of both CTAs) have landed
          tc_fence_after();
          const uint32_t sa = smem_base + stage * stage_bytes;
          if (p.fused_planes) {
            const uint32_t sb = sa + 2 * kABytes;
#pragma unroll
            for (int term = 0; term < 3; ++term) {   // h1*h0', h0*h1', h0*h0': smallest first
              const uint32_t ta = sa + (term == 0 ? kABytes : 0);
              const uint32_t tb = sb + (term == 1 ? kBBytesCta : 0);
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k) {
                const uint64_t da = umma_desc_sw128(ta + k * (UMMA_K * 2));
                const uint64_t db = umma_desc_sw128(tb + k * (UMMA_K * 2));
                mma(tmem_d, da, db, (it > 0 || term > 0 || k > 0) ? 1u : 0u);
              }
            }
          } else {
            const uint32_t sb = sa + kABytes;
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t da = umma_desc_sw128(sa + k * (UMMA_K * 2));
              const uint64_t db = umma_desc_sw128(sb + k * (UMMA_K * 2));
              mma(tmem_d, da, db, (it > 0 || k > 0) ? 1u : 0u);
            }
          }
          commit(empty_bar(stage));               // smem slot free (in both CTAs) once these MMAs retire
          if (++stage == n_stages) { stage = 0; phase ^= 1u; }
        }
        commit(tfull_bar(as));                    // accumulator complete -> epilogue(s)
        as ^= 1;
        if (as == 0) aphase ^= 1u;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------ epilogue ------------------------------
    const int quarter = warp & 3;                 // TMEM lanes this warp may touch: 32*quarter..+31
    const int row_in_tile = quarter * 32 + lane;
    const int chunk0 = ((warp - 2) >> 2) * (BN / 64);   // first of this warp's four 32-column chunks
    int as = 0;
    uint32_t aphase = 0;
    for (int64_t t = unit; t < p.n_tiles; t += n_units) {
      const int qt = my_qt(t);
      const int64_t gb = t / p.n_qtiles;
      const int qrow = qt * BM + row_in_tile;
      const bool row_ok = qrow < p.q;
      const float qs = (p.qry_scale != nullptr && row_ok) ? __ldg(p.qry_scale + qrow) : 1.0f;
      float tau = __int_as_float(0x7f800000);       // +inf: padded query rows keep nothing
      if (p.mode == kGemmFilter && row_ok) tau = __ldg(p.tau + qrow);
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      if (p.mode == kGemmProbe || p.mode == kGemmProbeLd) {   // measurement only (DVL_GEMM_PROBE=1 / 2): the MMA/TMA
        if (p.mode == kGemmProbeLd) {                          // pipeline without an epilogue / with its TMEM reads only
          uint32_t v[32];
          const uint32_t ta = tmem_base + (uint32_t)(as * BN) + ((uint32_t)((warp & 3) * 32) << 16);
          uint32_t acc = 0;
#pragma unroll
          for (int ci = 0; ci < BN / 64; ++ci) {
            tmem_ld_32x32(ta + (uint32_t)((((warp - 2) >> 2) * (BN / 64) + ci) * 32), v);
            tmem_ld_wait_regs(v);
#pragma unroll
            for (int j = 0; j < 32; ++j) acc |= v[j];
          }
          if (acc == 0x12345678u) p.cand_cnt[0] = acc;         // keeps the reads alive
        }
        tc_fence_before();
        if (kPair) mbar_arrive_cluster(tempty_bar(as) & kPeerBitMask);
        else mbar_arrive(tempty_bar(as));
        as ^= 1;
        if (as == 0) aphase ^= 1u;
        continue;
      }
      const uint32_t taddr = tmem_base + (uint32_t)(as * BN) + ((uint32_t)(quarter * 32) << 16);
      const bool scaled = p.qry_scale != nullptr || p.gal_scale != nullptr;
      if (p.mode == kGemmFilter) {
        // ---- streaming top-k: one compare per similarity against tau.  Survivors (rare once tau has
        // tightened) are appended to the query's list.  Two passes over the thread's four 32-column chunks:
        //   pass 1  accumulator read + fused scale + compare -> a 32-bit survivor mask per chunk;
        //   one global atomic per TILE reserves the slots of all the thread's survivors;
        //   pass 2  the chunks that hold a survivor in any lane are read from TMEM again (reads cost nothing
        //           next to the MMAs) and the survivors are picked out of the registers (pick32) and written.
        // One atomic per chunk (the first version) put ~four global round trips of ~2000 cycles on every tile of
        // every epilogue warp -- about the time the tensor cores need for the tile -- and the kernel ran 11 %
        // below its own MMA/TMA pipeline (tools/gemm_probe.py: 1268 vs 1404 TFLOP/s without the epilogue).
        unsigned long long* list = p.cand + (int64_t)qrow * p.cap;
        uint32_t masks[BN / 64];
        uint32_t total = 0;
        auto load_scaled = [&](int ci, uint32_t (&v)[32], bool prefetch_next, uint32_t (&vnext)[32]) {
          const int chunk = chunk0 + ci;
          const int64_t col0 = p.row0 + gb * BN + chunk * 32;
          // the chunk's 32 gallery-row factors: lane j fetches column j's (one coalesced load), the element
          // loop gets them by shuffle
          float gs_lane = 1.0f;
          if (p.gal_scale != nullptr) gs_lane = (col0 + lane < p.n) ? __ldg(p.gal_scale + (col0 + lane) * p.gal_step) : 0.f;
          tmem_ld_wait_regs(v);
          if (prefetch_next) tmem_ld_32x32(taddr + (uint32_t)((chunk + 1) * 32), vnext);
          // uniform branches, outside the element loops (a shuffle under a per-element branch is neither
          // if-converted nor pipelined by the compiler)
          if (p.gal_scale != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              v[j] = __float_as_uint(__uint_as_float(v[j]) * qs * __shfl_sync(0xffffffffu, gs_lane, j));
          } else if (scaled) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * qs);
          }
        };
        {
          uint32_t vbuf[2][32];
          tmem_ld_32x32(taddr + (uint32_t)(chunk0 * 32), vbuf[0]);
#pragma unroll
          for (int ci = 0; ci < BN / 64; ++ci) {
            uint32_t (&v)[32] = vbuf[ci & 1];
            load_scaled(ci, v, ci + 1 < BN / 64, vbuf[(ci + 1) & 1]);
            uint32_t mask = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) mask |= (__uint_as_float(v[j]) >= tau ? 1u : 0u) << j;
            const int64_t left = p.n - (p.row0 + gb * BN + (chunk0 + ci) * 32);
            if (left < 32) mask &= (left <= 0) ? 0u : ((1u << left) - 1u);
            masks[ci] = mask;
            total += (uint32_t)__popc(mask);
          }
        }
        if (p.probe_pass1_only) {                  // measurement only (DVL_GEMM_PROBE=3): thresholds, no appends
          if (total == 0xFFFFFFFFu) p.cand_cnt[qrow] = total;
        } else if (__any_sync(0xffffffffu, total != 0u)) {
          uint32_t slot = 0;
          if (total != 0u) slot = atomicAdd(p.cand_cnt + qrow, total);
#pragma unroll
          for (int ci = 0; ci < BN / 64; ++ci) {
            if (!__any_sync(0xffffffffu, masks[ci] != 0u)) continue;     // warp-uniform: tcgen05.ld is warp-wide
            uint32_t v[32];
            tmem_ld_32x32(taddr + (uint32_t)((chunk0 + ci) * 32), v);
            load_scaled(ci, v, false, v);
            uint32_t mask = masks[ci];
            const int64_t col0 = p.row0 + gb * BN + (chunk0 + ci) * 32;
            while (mask != 0u) {
              const int j = __ffs(mask) - 1;
              mask &= mask - 1u;
              if ((int64_t)slot < p.cap) list[slot] = make_entry(__uint_as_float(pick32(v, j)), p.index_base + col0 + j);
              ++slot;
            }
          }
        }
      } else {
      uint32_t vbuf[2][32];                        // software pipeline as in the filter branch
      tmem_ld_32x32(taddr + (uint32_t)(chunk0 * 32), vbuf[0]);
#pragma unroll
      for (int ci = 0; ci < BN / 64; ++ci) {
        const int chunk = chunk0 + ci;
        uint32_t (&v)[32] = vbuf[ci & 1];
        const int64_t col0 = p.row0 + gb * BN + chunk * 32;
        float gs_lane = 1.0f;
        if (p.gal_scale != nullptr) gs_lane = (col0 + lane < p.n) ? __ldg(p.gal_scale + (col0 + lane) * p.gal_step) : 0.f;
        tmem_ld_wait_regs(v);
        if (ci + 1 < BN / 64) tmem_ld_32x32(taddr + (uint32_t)((chunk + 1) * 32), vbuf[(ci + 1) & 1]);
        if (p.gal_scale != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            v[j] = __float_as_uint(__uint_as_float(v[j]) * qs * __shfl_sync(0xffffffffu, gs_lane, j));
        } else if (scaled) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * qs);
        }
        if (p.mode == kGemmStore) {
          // A thread owns one query ROW of the chunk, but a row's 32 columns are only 128 bytes of output:
          // stored from the registers, every store instruction of the warp touches 32 different lines with
          // 16 bytes each.  The chunk is transposed through the warp's 32 x 33-word scratch instead, and each
          // instruction writes one row's 128 contiguous bytes (8x fewer L2 write transactions; time-neutral
          // on config 2, where key conversion + stores together are 37 of the GEMM's 140 us).
          uint32_t* tile = epi_scratch + (warp - 2) * (32 * 33);
          const int qrow0 = qt * BM + quarter * 32;             // first query row of this warp
          const int rows_ok = (p.q - qrow0 < 32) ? (p.q - qrow0) : 32;   // <= 0: nothing to store
          if (p.scores != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) tile[lane * 33 + j] = v[j];
            __syncwarp();
            float* dst = p.scores + (int64_t)qrow0 * p.lds + col0 + lane;
#pragma unroll 8
            for (int r = 0; r < 32; ++r)
              if (r < rows_ok) dst[(int64_t)r * p.lds] = __uint_as_float(tile[r * 33 + lane]);
            __syncwarp();
          }
          if (p.keys != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) tile[lane * 33 + j] = score_to_key(__uint_as_float(v[j]));
            __syncwarp();
            uint32_t* dst = p.keys + (int64_t)qrow0 * p.ldk + col0 + lane;
#pragma unroll 8
            for (int r = 0; r < 32; ++r)
              if (r < rows_ok) dst[(int64_t)r * p.ldk] = tile[r * 33 + lane];
            __syncwarp();
          }
        } else {
          // SEED, first chunk: every row is a candidate, position = column (no atomics, 256 B per thread)
          if (row_ok && col0 < p.n) {
            unsigned long long* dst = p.cand + (int64_t)qrow * p.cap + (col0 - p.row0);
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              // rows past the end of the gallery become NaN-keyed entries that are never ranked
              const unsigned long long e0 =
                  (col0 + j < p.n) ? make_entry(__uint_as_float(v[j]), p.index_base + col0 + j) : ~0ull;
              const unsigned long long e1 =
                  (col0 + j + 1 < p.n) ? make_entry(__uint_as_float(v[j + 1]), p.index_base + col0 + j + 1) : ~0ull;
              st_global_v4(dst + j, (uint32_t)e0, (uint32_t)(e0 >> 32), (uint32_t)e1, (uint32_t)(e1 >> 32));
            }
          }
        }
      }
      }
      tc_fence_before();
      if (kPair) mbar_arrive_cluster(tempty_bar(as) & kPeerBitMask);   // the leader's barrier counts both epilogues
      else mbar_arrive(tempty_bar(as));
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
  }

  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();   // neither CTA leaves while the peer may touch its memory
  if (warp == 1) {
    if (kPair) tmem_dealloc_pair(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// DVL_GEMM_PAIR=0 / 1 forces the single-CTA / CTA-pair kernel (measurements); unset: chosen by the query count
int gemm_pair_mode() {
  static const int mode = [] {
    const char* e = getenv("DVL_GEMM_PAIR");
    if (e == nullptr || e[0] == '\0') return -1;
    return e[0] == '0' ? 0 : 1;
  }();
  return mode;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// rows_view rows, `row_step` physical rows apart (row_step > 1: a strided sample of the matrix, no copy)
int make_map(Ctx* c, CUtensorMap* map, const void* base, bool fp16, int64_t rows, int d, int dp, int planes,
             int box_rows, int64_t rows_view = -1, int64_t row_step = 1) {
  if (rows_view < 0) rows_view = rows;
  if (c->encode_tiled == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess)
      return set_error(c, DVL_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable: %s", cudaGetErrorString(e));
    c->encode_tiled = fn;
  }
  // logical depth d (not the pitch): columns >= d are zero-filled by TMA
  cuuint64_t dims[3] = {(cuuint64_t)d, (cuuint64_t)rows_view, (cuuint64_t)planes};
  cuuint64_t strides[2] = {(cuuint64_t)dp * 2ull * (cuuint64_t)row_step, (cuuint64_t)rows * (cuuint64_t)dp * 2ull};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = ((EncodeTiledFn)c->encode_tiled)(map, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                                                const_cast<void*>(base),
                                                dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(c, DVL_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return DVL_OK;
}

}  // namespace

int launch_sim_gemm(Ctx* c, const GemmArgs& a, cudaStream_t st) {
  if (a.dp % 8 != 0) return set_error(c, DVL_ERR_INVALID, "plane row pitch must be a multiple of 8 elements");
  if (a.n <= 0 || a.q <= 0 || a.d <= 0) return set_error(c, DVL_ERR_INVALID, "empty similarity problem");
  if (a.row0 < 0 || a.row0 % BN != 0 || a.row1 <= a.row0 || a.row1 > a.n)
    return set_error(c, DVL_ERR_INVALID, "bad gallery row range [%lld, %lld)", (long long)a.row0, (long long)a.row1);
  if (a.mode != kGemmStore && (a.cand == nullptr || a.cand_cnt == nullptr || a.cap < 1))
    return set_error(c, DVL_ERR_INVALID, "streaming top-k launch without candidate buffers");
  if (a.mode == kGemmFilter && a.tau == nullptr) return set_error(c, DVL_ERR_INVALID, "filter launch without thresholds");
  if (a.mode == kGemmSeed && a.row1 - a.row0 > a.cap)
    return set_error(c, DVL_ERR_INVALID, "seed chunk larger than the candidate capacity");
  if (a.index_base + a.n > (1ll << 28)) return set_error(c, DVL_ERR_UNSUPPORTED, "global gallery index needs more than 28 bits");
  if (a.n > 0x7FFFFF00ll) return set_error(c, DVL_ERR_UNSUPPORTED, "gallery shard larger than 2^31 rows");
  CUtensorMap map_q, map_g;
  const bool fp16 = a.planes == 2;
  if (fp16 && (a.qry_scale == nullptr || a.gal_scale == nullptr))
    return set_error(c, DVL_ERR_INVALID, "the scaled fp16 split needs both row-scale arrays");
  DVL_TRY(make_map(c, &map_q, a.qry_planes, fp16, a.q, a.d, a.dp, a.planes, BM));
  // a.n = rows the kernel sees; when a.gal_row_step > 1 they are every gal_row_step-th row of a
  // gallery of a.gal_rows_total rows
  const int64_t total_rows = a.gal_row_step > 1 ? a.gal_rows_total : a.n;
  // CTA pairs (cta_group::2).  Measured on one B200 (DVL_GEMM_PAIR=0/1, same box, same run):
  //   config 5 (fp32 mode, 1 M x 512 x 1000): 2.49 -> 2.37 ms -- the fp16-pair pipeline gets three 64 KB stages
  //     instead of two 96 KB ones;
  //   config 3 / config 4 shard (bf16): 1.524 vs 1.541 ms / 14.06 vs 14.78 ms at 1740 vs 1635 MHz, i.e. the same
  //     cycle count: the bf16 kernel already runs at the sustained (power-capped) tensor rate of this GPU
  //     (1340-1365 TFLOP/s against a measured sustained cuBLAS peak of 1368), so halving its L2 -> shared
  //     memory traffic buys nothing;
  //   config 2 (319 queries): 0.138 -> 0.171 ms -- the second 256-row query tile is three quarters padding.
  // Hence: pairs for the fp16-pair mode when the query count fills 256-row tiles, single CTAs otherwise.
  const int64_t rows_single = (int64_t)((a.q + BM - 1) / BM) * BM, rows_pair = (int64_t)((a.q + 2 * BM - 1) / (2 * BM)) * 2 * BM;
  const bool pair = gemm_pair_mode() == 1 ||
                    (gemm_pair_mode() < 0 && fp16 && c->num_sms % 2 == 0 && a.q > BM &&
                     rows_pair * 100 <= rows_single * 104);
  DVL_TRY(make_map(c, &map_g, a.gal_planes, fp16, total_rows, a.d, a.dp, a.planes, pair ? BN / 2 : BN, a.n,
                   a.gal_row_step));
  KernelParams p;
  p.mode = a.mode;
  {
    // DVL_GEMM_PROBE=1: FILTER launches run without their epilogue (results are garbage; tools/gemm_probe.py)
    static const int probe = [] { const char* e = getenv("DVL_GEMM_PROBE"); return e != nullptr ? atoi(e) : 0; }();
    if ((probe == 1 || probe == 2) && a.mode == kGemmFilter) p.mode = probe == 2 ? kGemmProbeLd : kGemmProbe;
    p.probe_pass1_only = probe == 3 ? 1 : 0;
  }
  p.row0 = a.row0;
  p.n = a.row1;
  p.q = a.q;
  p.n_qtiles = pair ? (a.q + 2 * BM - 1) / (2 * BM) : (a.q + BM - 1) / BM;
  p.n_tiles = (int64_t)p.n_qtiles * ((a.row1 - a.row0 + BN - 1) / BN);
  p.tau = a.tau;
  p.cand = a.cand;
  p.cand_cnt = a.cand_cnt;
  p.cap = a.cap;
  p.labels = a.labels;
  p.index_base = a.index_base;
  p.k_blocks = (a.d + BK - 1) / BK;
  p.n_terms = (a.planes == 3) ? 6 : (a.planes == 2 ? 3 : 1);
  for (int i = 0; i < kMaxTerms; ++i) {
    p.term_a[i] = (a.planes == 3) ? kTermsX3A[i] : ((a.planes == 2 && i < 3) ? kTermsX2A[i] : 0);
    p.term_b[i] = (a.planes == 3) ? kTermsX3B[i] : ((a.planes == 2 && i < 3) ? kTermsX2B[i] : 0);
  }
  const int umma_m = pair ? 2 * BM : BM;
  p.idesc = fp16 ? umma_idesc_f16(umma_m, BN) : umma_idesc_bf16(umma_m, BN);
  p.fused_planes = fp16 ? 1 : 0;
  p.n_planes = a.planes;
  p.qry_scale = a.qry_scale;
  p.gal_scale = a.gal_scale;
  p.gal_step = a.gal_row_step > 1 ? a.gal_row_step : 1;
  p.keys = a.keys;
  p.ldk = a.ldk;
  p.scores = a.scores;
  p.lds = a.lds;
  if (a.keys != nullptr && (a.ldk % 4 != 0 || a.ldk < ((a.n + BN - 1) / BN) * BN))
    return set_error(c, DVL_ERR_INVALID, "key pitch must cover whole 256-column tiles");
  if (a.scores != nullptr && (a.lds % 4 != 0 || a.lds < ((a.n + BN - 1) / BN) * BN))
    return set_error(c, DVL_ERR_INVALID, "score pitch must cover whole 256-column tiles");
  if (!c->gemm_attr_set) {
    DVL_CUDA(c, cudaFuncSetAttribute(sim_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)kSmemBytes));
    DVL_CUDA(c, cudaFuncSetAttribute(sim_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)kSmemBytes));
    c->gemm_attr_set = true;
  }
  if (!pair) {
    int64_t grid = p.n_tiles < c->num_sms ? p.n_tiles : c->num_sms;
    sim_gemm_kernel<false><<<(unsigned)grid, kNumThreads, kSmemBytes, st>>>(map_q, map_g, p);
  } else {
    // one persistent CTA pair per TPC: a cluster of two CTAs
    const int64_t pairs = p.n_tiles < c->num_sms / 2 ? p.n_tiles : c->num_sms / 2;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(2 * pairs));
    cfg.blockDim = dim3(kNumThreads);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    DVL_CUDA(c, cudaLaunchKernelEx(&cfg, sim_gemm_kernel<true>, map_q, map_g, p));
  }
  DVL_LAUNCH_CHECK(c, "sim_gemm_kernel");
  return DVL_OK;
}

}  // namespace dvl
