// Operand preparation for the similarity stage and the query-independent label statistics.
//
//  * split_rows:   embeddings -> 16-bit planes for the tensor cores.
//                  bf16 mode: one bf16 plane (round to nearest even).
//                  fp32 mode (default): every row is scaled by a power of two that brings its largest
//                  magnitude into [2^14, 2^15) and split into TWO fp16 planes, x*2^e = h0 + h1 + eps with
//                  |eps| <= 2^-22 |x*2^e|; the three product terms h0*h0', h0*h1', h1*h0' accumulated in
//                  fp32 on the tensor cores reproduce the fp32 similarity of measuring_bias.py:114 to
//                  ~3*2^-22 of |g||t| (the dropped h1*h1' term and the two eps), i.e. the accuracy of an
//                  fp32 dot product itself, at half the tensor work of the three-plane split.  The row's
//                  2^-e goes into the row-scale array the GEMM epilogue multiplies by (exact).
//                  fp32x3 mode: x == b0 + b1 + b2 exactly, three bf16 planes, six product terms
//                  (i + j <= 2): ~2^-24, any fp32 range, no row scale.
//                  The row-scale array also carries 1/||row|| for the fused normalisation
//                  (measuring_bias.py:79).
//  * class_counts: Counter(df.label)  (measuring_bias.py:23,56) -- once per call instead of once
//                  per query and per metric.
#include "dvl_common.cuh"
#include "dvl_internal.h"

namespace dvl {

template <typename T>
__device__ __forceinline__ float load_as_float(const T* p);
template <>
__device__ __forceinline__ float load_as_float<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float load_as_float<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}
template <>
__device__ __forceinline__ float load_as_float<__half>(const __half* p) { return __half2float(*p); }

// Four consecutive elements of a row as floats.  `vec`: the row starts are 16-byte (fp32) / 8-byte
// (16-bit types) aligned and d is a multiple of 4, so the quad is one vector load; otherwise element
// loads with a bound on d.  Columns >= d (the padding up to the plane pitch) read as zero.
template <typename T>
__device__ __forceinline__ void load_quad(const T* __restrict__ s, int c, int d, bool vec, float (&x)[4]) {
  if (vec && c + 3 < d) {
    if constexpr (sizeof(T) == 4) {
      const float4 v = __ldcs(reinterpret_cast<const float4*>(s + c));
      x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
    } else {
      const uint2 v = __ldcs(reinterpret_cast<const uint2*>(s + c));
      const T* h = reinterpret_cast<const T*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) x[j] = load_as_float<T>(h + j);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) x[j] = (c + j < d) ? load_as_float<T>(s + c + j) : 0.f;
  }
}

// One warp per row; a lane handles 4 consecutive elements per 128-element step (coalesced 512 B per
// warp for fp32) and issues the loads of kUnroll steps before it touches any of them: at D = 512 a
// whole row is in flight per warp.  HBM-bound: 4 B read + 2*PLANES B written per element.
template <typename T, int PLANES>
__global__ void __launch_bounds__(256) split_rows_kernel(const T* __restrict__ src, int64_t rows, int d, int dp,
                                                         int64_t plane_rows, __nv_bfloat16* __restrict__ dst,
                                                         float* __restrict__ inv_norm, int vec) {
  constexpr int kUnroll = 4;
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t plane_stride = plane_rows * (int64_t)dp;   // rows of a whole plane (>= rows: row-range launches)
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += warps_total) {
    const T* s = src + row * (int64_t)d;
    __nv_bfloat16* o = dst + row * (int64_t)dp;
    float ss = 0.f;
    for (int c0 = lane * 4; c0 < dp; c0 += 128 * kUnroll) {
      float x[kUnroll][4];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) load_quad<T>(s, c0 + u * 128, d, vec != 0, x[u]);
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int c = c0 + u * 128;
        if (c >= dp) break;
        __align__(8) __nv_bfloat16 b[PLANES][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          ss = fmaf(x[u][j], x[u][j], ss);
          float r = x[u][j];
          // A +-inf / NaN element lives in the LAST plane alone (the residual inf - inf would be NaN, and every
          // product term that pairs an inf with the other operand's small residual planes could flip its sign):
          // the last plane only multiplies the other operand's leading plane, so the similarity comes out as the
          // +-inf / NaN an fp32 mat-vec produces (measuring_bias.py:114).
          const bool fin = fabsf(r) <= 3.4028234664e38f;
#pragma unroll
          for (int p = 0; p < PLANES; ++p) {
            const bool last = p == PLANES - 1;
            b[p][j] = (fin || last) ? __float2bfloat16_rn(r) : __float2bfloat16_rn(0.0f);
            r = fin ? r - __bfloat162float(b[p][j]) : r;   // exact: the residual of a round-to-8-bits fits fp32
          }
        }
#pragma unroll
        for (int p = 0; p < PLANES; ++p)
          *reinterpret_cast<uint2*>(o + p * plane_stride + c) = *reinterpret_cast<const uint2*>(b[p]);
      }
    }
    if (inv_norm != nullptr) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
      if (lane == 0) inv_norm[row] = 1.0f / sqrtf(ss);
    }
  }
}

// fp32 mode: one warp per row, two fp16 planes of the power-of-two scaled row.  Rows of up to 512
// elements stay in registers between the magnitude pass and the split pass; longer rows are read twice
// (the second time from L1/L2).
template <typename T>
__global__ void __launch_bounds__(256) split_rows_f16x2_kernel(const T* __restrict__ src, int64_t rows, int d, int dp,
                                                               int64_t plane_rows, __half* __restrict__ dst,
                                                               float* __restrict__ row_scale, int normalize, int vec) {
  constexpr int kUnroll = 4;
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t plane_stride = plane_rows * (int64_t)dp;   // rows of a whole plane (>= rows: row-range launches)
  const bool single = dp <= 128 * kUnroll;
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += warps_total) {
    const T* s = src + row * (int64_t)d;
    __half* o = dst + row * (int64_t)dp;
    float ss = 0.f, amax = 0.f;
    float x[kUnroll][4];
    for (int c0 = lane * 4; c0 < dp; c0 += 128 * kUnroll) {
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) load_quad<T>(s, c0 + u * 128, d, vec != 0, x[u]);
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          ss = fmaf(x[u][j], x[u][j], ss);
          const float a = fabsf(x[u][j]);
          amax = fmaxf(amax, (a <= 3.4028234664e38f) ? a : 0.f);   // the scale comes from the finite elements
        }
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      ss += __shfl_xor_sync(0xffffffffu, ss, off);
      amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, off));
    }
    // e = 14 - floor(log2(amax)), clamped so that 2^e and 2^-e are normal floats; an all-zero row keeps scale 1
    int e = 0;
    if (amax > 0.f) {
      const int be = (int)((__float_as_uint(amax) >> 23) & 0xFFu);   // biased exponent; 0 for subnormals
      e = 14 - (be - 127);
      e = e > 126 ? 126 : (e < -126 ? -126 : e);
    }
    const float up = __uint_as_float((uint32_t)(127 + e) << 23);
    const float down = __uint_as_float((uint32_t)(127 - e) << 23);
    for (int c0 = lane * 4; c0 < dp; c0 += 128 * kUnroll) {
      if (!single) {
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) load_quad<T>(s, c0 + u * 128, d, vec != 0, x[u]);
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int c = c0 + u * 128;
        if (c >= dp) break;
        __align__(8) __half h0[4], h1[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float xs = x[u][j] * up;            // exact (power of two; |xs| < 2^15)
          // a +-inf / NaN element goes to the second plane alone: that plane only meets the other operand's
          // leading plane (terms h1*h0', h0*h1', h0*h0'), so the similarity is the +-inf / NaN of an fp32
          // product instead of inf - inf from a residual, or an inf whose sign a tiny residual decides
          const bool fin = fabsf(x[u][j]) <= 3.4028234664e38f;
          h0[j] = __float2half_rn(fin ? xs : 0.f);
          h1[j] = __float2half_rn(fin ? xs - __half2float(h0[j]) : x[u][j]);   // the residual is exact in fp32
        }
        *reinterpret_cast<uint2*>(o + c) = *reinterpret_cast<const uint2*>(h0);
        *reinterpret_cast<uint2*>(o + plane_stride + c) = *reinterpret_cast<const uint2*>(h1);
      }
    }
    if (lane == 0) row_scale[row] = normalize ? down / sqrtf(ss) : down;
  }
}

template <typename T>
static int launch_split_t(Ctx* c, const T* src, int64_t rows, int d, int dp, int planes, void* dst, float* row_scale,
                          bool normalize, int64_t plane_rows, cudaStream_t st) {
  const int warps_per_block = 8;
  int64_t blocks = (rows + warps_per_block - 1) / warps_per_block;
  int64_t cap = (int64_t)c->num_sms * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  // whole-quad vector loads need every row start aligned to the quad: d % 4 == 0 and an aligned base
  const int vec = (d % 4 == 0 && (reinterpret_cast<uintptr_t>(src) % (4 * sizeof(T))) == 0) ? 1 : 0;
  if (planes == 1)
    split_rows_kernel<T, 1><<<(unsigned)blocks, 256, 0, st>>>(src, rows, d, dp, plane_rows, (__nv_bfloat16*)dst,
                                                             normalize ? row_scale : nullptr, vec);
  else if (planes == 3)
    split_rows_kernel<T, 3><<<(unsigned)blocks, 256, 0, st>>>(src, rows, d, dp, plane_rows, (__nv_bfloat16*)dst,
                                                             normalize ? row_scale : nullptr, vec);
  else
    split_rows_f16x2_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(src, rows, d, dp, plane_rows, (__half*)dst, row_scale,
                                                                normalize ? 1 : 0, vec);
  DVL_LAUNCH_CHECK(c, planes == 2 ? "split_rows_f16x2_kernel" : "split_rows_kernel");
  return DVL_OK;
}

int launch_split_rows(Ctx* c, const void* src, int dtype, int64_t rows, int d, int dp, int planes, void* dst,
                      float* row_scale, bool normalize, cudaStream_t st, int64_t plane_rows) {
  if (plane_rows < rows) plane_rows = rows;
  if (planes < 1 || planes > 3) return set_error(c, DVL_ERR_INVALID, "planes must be 1, 2 or 3");
  if (dp % 4 != 0) return set_error(c, DVL_ERR_INVALID, "row pitch must be a multiple of 4");
  if ((planes == 2 || normalize) && row_scale == nullptr)
    return set_error(c, DVL_ERR_INVALID, "split without a row-scale array");
  switch (dtype) {
    case DVL_F32: return launch_split_t<float>(c, (const float*)src, rows, d, dp, planes, dst, row_scale, normalize, plane_rows, st);
    case DVL_BF16:
      return launch_split_t<__nv_bfloat16>(c, (const __nv_bfloat16*)src, rows, d, dp, planes, dst, row_scale, normalize,
                                           plane_rows, st);
    case DVL_F16: return launch_split_t<__half>(c, (const __half*)src, rows, d, dp, planes, dst, row_scale, normalize, plane_rows, st);
    default: return set_error(c, DVL_ERR_INVALID, "unknown dtype %d", dtype);
  }
}

// ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) class_counts_kernel(const uint8_t* __restrict__ labels, int64_t n,
                                                           int n_classes, unsigned long long* __restrict__ counts,
                                                           unsigned int* __restrict__ bad) {
  __shared__ unsigned int sh[kMaxClasses];
  if (threadIdx.x < kMaxClasses) sh[threadIdx.x] = 0;
  __syncthreads();
  unsigned int local[kMaxClasses];
#pragma unroll
  for (int k = 0; k < kMaxClasses; ++k) local[k] = 0;
  unsigned int oob = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int l = labels[i];
    oob |= (l >= n_classes);
#pragma unroll
    for (int k = 0; k < kMaxClasses; ++k) local[k] += (l == k);
  }
#pragma unroll
  for (int k = 0; k < kMaxClasses; ++k) {
    unsigned int v = local[k];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&sh[k], v);
  }
  if (oob) atomicOr(bad, 1u);
  __syncthreads();
  if (threadIdx.x < n_classes && sh[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

// ----------------------------------------------------------------------------------------
// Integer labels as the dataset hands them over (int64 from np.array / torch.long, datasets.py:44) ->
// uint8 class codes, plus what Counter(df.label) would see: per-class counts, the largest label, and
// whether any label lies outside 0..kMaxClasses-1.  One pass, 8 bytes read + 1 byte written per row.
template <typename T>
__global__ void __launch_bounds__(256) encode_labels_kernel(const T* __restrict__ labels, int64_t n,
                                                            uint8_t* __restrict__ codes,
                                                            unsigned long long* __restrict__ info) {
  __shared__ unsigned int sh[kMaxClasses + 2];
  if (threadIdx.x < kMaxClasses + 2) sh[threadIdx.x] = 0;
  __syncthreads();
  unsigned int local[kMaxClasses];
#pragma unroll
  for (int k = 0; k < kMaxClasses; ++k) local[k] = 0;
  unsigned int oob = 0, mx = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const long long v = (long long)labels[i];
    const bool bad = v < 0 || v >= kMaxClasses;
    oob |= bad;
    const unsigned int l = bad ? 0u : (unsigned int)v;
    codes[i] = (uint8_t)(bad ? 0xFFu : l);
    mx = max(mx, l);
#pragma unroll
    for (int k = 0; k < kMaxClasses; ++k) local[k] += (!bad && l == (unsigned)k);
  }
#pragma unroll
  for (int k = 0; k < kMaxClasses; ++k) {
    unsigned int v = local[k];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&sh[k], v);
  }
  if (oob) atomicOr(&sh[kMaxClasses], 1u);
  atomicMax(&sh[kMaxClasses + 1], mx);
  __syncthreads();
  if (threadIdx.x < kMaxClasses && sh[threadIdx.x]) atomicAdd(&info[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
  if (threadIdx.x == 0) {
    if (sh[kMaxClasses]) atomicOr(&info[kMaxClasses], 1ull);
    atomicMax(&info[kMaxClasses + 1], (unsigned long long)sh[kMaxClasses + 1]);
  }
}

int launch_encode_labels(Ctx* c, const void* labels, int label_dtype, int64_t n, uint8_t* codes,
                         unsigned long long* info, cudaStream_t st) {
  DVL_CUDA(c, cudaMemsetAsync(info, 0, sizeof(unsigned long long) * (kMaxClasses + 2), st));
  int64_t blocks = (n + 256 * 8 - 1) / (256 * 8);
  if (blocks > (int64_t)c->num_sms * 8) blocks = (int64_t)c->num_sms * 8;
  if (blocks < 1) blocks = 1;
  switch (label_dtype) {
    case DVL_LABEL_I64:
      encode_labels_kernel<long long><<<(unsigned)blocks, 256, 0, st>>>((const long long*)labels, n, codes, info);
      break;
    case DVL_LABEL_I32:
      encode_labels_kernel<int><<<(unsigned)blocks, 256, 0, st>>>((const int*)labels, n, codes, info);
      break;
    case DVL_LABEL_U8:
      encode_labels_kernel<uint8_t><<<(unsigned)blocks, 256, 0, st>>>((const uint8_t*)labels, n, codes, info);
      break;
    default: return set_error(c, DVL_ERR_INVALID, "unknown label dtype %d", label_dtype);
  }
  DVL_LAUNCH_CHECK(c, "encode_labels_kernel");
  return DVL_OK;
}

int launch_class_counts(Ctx* c, const uint8_t* labels, int64_t n, int n_classes, unsigned long long* counts,
                        cudaStream_t st) {
  // counts has kMaxClasses + 1 slots; the last one is the out-of-range flag
  DVL_CUDA(c, cudaMemsetAsync(counts, 0, sizeof(unsigned long long) * (kMaxClasses + 1), st));
  int64_t blocks = (n + 256 * 16 - 1) / (256 * 16);
  if (blocks > (int64_t)c->num_sms * 8) blocks = (int64_t)c->num_sms * 8;
  if (blocks < 1) blocks = 1;
  class_counts_kernel<<<(unsigned)blocks, 256, 0, st>>>(labels, n, n_classes, counts,
                                                        reinterpret_cast<unsigned int*>(counts + kMaxClasses));
  DVL_LAUNCH_CHECK(c, "class_counts_kernel");
  return DVL_OK;
}

}  // namespace dvl
