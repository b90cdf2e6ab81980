// Corpus ingestion: row-major rows -> tiled/swizzled HBM image (rs_layout.h), dtype conversion,
// document norms (the reference recomputes all N norms on every call, retriever.py:603; here they
// are computed once), the inverse transform (used by tests and the CPU-baseline sampler) and the
// on-device synthetic corpus generator (bit-identical to oracle/retrieval_oracle.py:synth_rows).
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "rs_kernels.h"
#include "rs_layout.h"

namespace rs {

__device__ __forceinline__ uint16_t ing_f32_to_bf16(float x) {
  uint32_t u = __float_as_uint(x);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
__device__ __forceinline__ float ing_bf16_to_f32(uint16_t b) { return __uint_as_float((uint32_t)b << 16); }

// Convert 8 consecutive fp32 values of one row to the storage dtype, write the 16-byte chunk(s)
// at the swizzled position and return the sum of squares of the represented values.
__device__ __forceinline__ float store_chunk8(const float (&x)[8], long long row, int c8, int D, int dtype,
                                              int scale_exp, uint8_t* corpus) {
  const uint32_t kc_count = (uint32_t)D / RS_KCHUNK_ELEMS;
  const uint32_t planes = dtype == RS_DTYPE_FP32 ? 2u : 1u;
  uint16_t h[8], l[8];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float rep;
    if (dtype == RS_DTYPE_BF16) {
      h[i] = ing_f32_to_bf16(x[i]);
      rep = ing_bf16_to_f32(h[i]);
    } else if (dtype == RS_DTYPE_FP16) {
      __half hv = __float2half_rn(x[i]);
      h[i] = __half_as_ushort(hv);
      rep = __half2float(hv);
    } else {
      const float xs = ldexpf(x[i], scale_exp);
      __half hv = __float2half_rn(xs);
      const float r = xs - __half2float(hv);  // exact
      __half lv = __float2half_rn(ldexpf(r, RS_FP16_SPLIT_SHIFT));
      h[i] = __half_as_ushort(hv);
      l[i] = __half_as_ushort(lv);
      rep = ldexpf(__half2float(hv) + ldexpf(__half2float(lv), -RS_FP16_SPLIT_SHIFT), -scale_exp);
    }
    ss += rep * rep;
  }
  const uint64_t off = rs_corpus_offset((uint64_t)row, (uint32_t)c8 * 8u, 0, kc_count, planes);
  uint4 w;
  w.x = h[0] | ((uint32_t)h[1] << 16);
  w.y = h[2] | ((uint32_t)h[3] << 16);
  w.z = h[4] | ((uint32_t)h[5] << 16);
  w.w = h[6] | ((uint32_t)h[7] << 16);
  *reinterpret_cast<uint4*>(corpus + off) = w;
  if (dtype == RS_DTYPE_FP32) {
    w.x = l[0] | ((uint32_t)l[1] << 16);
    w.y = l[2] | ((uint32_t)l[3] << 16);
    w.z = l[4] | ((uint32_t)l[5] << 16);
    w.w = l[6] | ((uint32_t)l[7] << 16);
    *reinterpret_cast<uint4*>(corpus + off + (uint64_t)kc_count * RS_BLOCK_BYTES) = w;
  }
  return ss;
}

__device__ __forceinline__ double block_sum_128(float v) {
  __shared__ double red[4];
  double d = (double)v;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d;
  __syncthreads();
  return red[0] + red[1] + red[2] + red[3];
}

// one 128-thread block per row
__global__ void __launch_bounds__(128) tile_rows_kernel(const void* __restrict__ rows, int src_dtype, long long row0,
                                                        int D, int dtype, int scale_exp,
                                                        uint8_t* __restrict__ corpus, float* __restrict__ dnorm) {
  const long long i = blockIdx.x;
  float ss = 0.f;
  for (int c8 = threadIdx.x; c8 < D / 8; c8 += blockDim.x) {
    float x[8];
    if (src_dtype == RS_DTYPE_FP32) {
      const float4* src = reinterpret_cast<const float4*>(static_cast<const float*>(rows) + (size_t)i * D) + c8 * 2;
      float4 a = src[0], b = src[1];
      x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w;
      x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    } else {
      uint4 raw = reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(rows) + (size_t)i * D)[c8];
      const uint32_t wv[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint16_t lo = (uint16_t)(wv[j] & 0xFFFFu), hi = (uint16_t)(wv[j] >> 16);
        if (src_dtype == RS_DTYPE_BF16) {
          x[2 * j] = ing_bf16_to_f32(lo);
          x[2 * j + 1] = ing_bf16_to_f32(hi);
        } else {
          x[2 * j] = __half2float(__ushort_as_half(lo));
          x[2 * j + 1] = __half2float(__ushort_as_half(hi));
        }
      }
    }
    ss += store_chunk8(x, row0 + i, c8, D, dtype, scale_exp, corpus);
  }
  const double tot = block_sum_128(ss);
  if (threadIdx.x == 0) dnorm[row0 + i] = (float)sqrt(tot);
}

__global__ void __launch_bounds__(128) untile_rows_kernel(const uint8_t* __restrict__ corpus, long long row0, int D,
                                                          int dtype, int scale_exp, float* __restrict__ out) {
  const long long i = blockIdx.x;
  const uint32_t kc_count = (uint32_t)D / RS_KCHUNK_ELEMS;
  const uint32_t planes = dtype == RS_DTYPE_FP32 ? 2u : 1u;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    const uint64_t off = rs_corpus_offset((uint64_t)(row0 + i), (uint32_t)c, 0, kc_count, planes);
    const uint16_t h = *reinterpret_cast<const uint16_t*>(corpus + off);
    float v;
    if (dtype == RS_DTYPE_BF16) {
      v = ing_bf16_to_f32(h);
    } else if (dtype == RS_DTYPE_FP16) {
      v = __half2float(__ushort_as_half(h));
    } else {
      const uint16_t l = *reinterpret_cast<const uint16_t*>(corpus + off + (uint64_t)kc_count * RS_BLOCK_BYTES);
      v = ldexpf(__half2float(__ushort_as_half(h)) + ldexpf(__half2float(__ushort_as_half(l)), -RS_FP16_SPLIT_SHIFT),
                 -scale_exp);
    }
    out[(size_t)i * D + c] = v;
  }
}

// ---- synthetic corpus: counter-based, reproducible for any sharding (SURVEY §7 hard part 8)
__device__ __forceinline__ unsigned long long mix64(unsigned long long h, unsigned long long seed) {
  unsigned long long z = seed + (h + 1ull) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// one warp per row, 4 rows per block
__global__ void __launch_bounds__(128) synth_fill_kernel(unsigned long long seed, long long global_row0,
                                                         long long local_row0, int nrows, int D, int dtype,
                                                         int scale_exp, uint8_t* __restrict__ corpus,
                                                         float* __restrict__ dnorm) {
  const int lane = threadIdx.x & 31;
  const long long i = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (i >= nrows) return;
  const unsigned long long grow = (unsigned long long)(global_row0 + i);
  const unsigned long long hbase = grow * (unsigned long long)(D / 4);
  unsigned long long sumsq = 0;
  for (int c8 = lane; c8 < D / 8; c8 += 32) {
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const unsigned long long z = mix64(hbase + (unsigned long long)(c8 * 2 + g), seed);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const long long v = (long long)((z >> (16 * e)) & 0xFFFFull) * 2 - 65535;
        sumsq += (unsigned long long)(v * v);
      }
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) sumsq += __shfl_xor_sync(0xffffffffu, sumsq, off);
  const double rinv = __drcp_rn(__dsqrt_rn((double)sumsq));
  float ss = 0.f;
  for (int c8 = lane; c8 < D / 8; c8 += 32) {
    float x[8];
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const unsigned long long z = mix64(hbase + (unsigned long long)(c8 * 2 + g), seed);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const long long v = (long long)((z >> (16 * e)) & 0xFFFFull) * 2 - 65535;
        x[g * 4 + e] = __double2float_rn(__dmul_rn((double)v, rinv));
      }
    }
    ss += store_chunk8(x, local_row0 + i, c8, D, dtype, scale_exp, corpus);
  }
  double d = (double)ss;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
  if (lane == 0) dnorm[local_row0 + i] = (float)sqrt(d);
}

__global__ void __launch_bounds__(128) tile_minmax_kernel(const float* __restrict__ dnorm, int n_rows, int n_tiles,
                                                          float2* __restrict__ minmax) {
  const int t = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (t >= n_tiles) return;
  float mn = 3.0e38f, mx = 0.f;
  for (int r = lane; r < RS_TILE_ROWS; r += 32) {
    const long long row = (long long)t * RS_TILE_ROWS + r;
    if (row < n_rows) {
      const float v = dnorm[row];
      mn = fminf(mn, v);
      mx = fmaxf(mx, v);
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, off));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  }
  if (lane == 0) minmax[t] = make_float2(mn, mx);
}

cudaError_t launch_tile_rows(const void* rows_dev, int src_dtype, long long row0, int nrows, int D, int dtype,
                             int scale_exp, uint8_t* corpus, float* dnorm, cudaStream_t stream) {
  if (nrows <= 0) return cudaSuccess;
  tile_rows_kernel<<<nrows, 128, 0, stream>>>(rows_dev, src_dtype, row0, D, dtype, scale_exp, corpus, dnorm);
  return cudaGetLastError();
}
cudaError_t launch_untile_rows(const uint8_t* corpus, long long row0, int nrows, int D, int dtype, int scale_exp,
                               float* out_rows_dev, cudaStream_t stream) {
  if (nrows <= 0) return cudaSuccess;
  untile_rows_kernel<<<nrows, 128, 0, stream>>>(corpus, row0, D, dtype, scale_exp, out_rows_dev);
  return cudaGetLastError();
}
cudaError_t launch_synth_fill(unsigned long long seed, long long global_row0, long long local_row0, int nrows, int D,
                              int dtype, int scale_exp, uint8_t* corpus, float* dnorm, cudaStream_t stream) {
  if (nrows <= 0) return cudaSuccess;
  synth_fill_kernel<<<(nrows + 3) / 4, 128, 0, stream>>>(seed, global_row0, local_row0, nrows, D, dtype, scale_exp,
                                                         corpus, dnorm);
  return cudaGetLastError();
}
cudaError_t launch_tile_minmax(const float* dnorm, int n_rows, int n_tiles, float2* minmax, cudaStream_t stream) {
  if (n_tiles <= 0) return cudaSuccess;
  tile_minmax_kernel<<<(n_tiles + 3) / 4, 128, 0, stream>>>(dnorm, n_rows, n_tiles, minmax);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Query hand-off from the embedder: batch_processor.py:89-92 mean-pools the last hidden state on the GPU,
// copies it to the host and hands NumPy to the retriever, which copies it back.  This kernel does the pooling
// (`outputs.last_hidden_state.mean(dim=1)`: a plain mean over ALL T positions, padding included, as the reference
// computes it) straight into the fp32 query buffer a device-buffer search reads, on the caller's stream.
// One thread per (query, column): T coalesced reads, fp32 accumulation in position order.
// ------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float pool_load(const T* p);
template <>
__device__ __forceinline__ float pool_load<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float pool_load<__half>(const __half* p) { return __half2float(*p); }
template <>
__device__ __forceinline__ float pool_load<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <typename T>
__global__ void __launch_bounds__(256) mean_pool_kernel(const T* __restrict__ hidden, int T_len, int D, float* __restrict__ out) {
  const int b = blockIdx.y;
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  const T* src = hidden + (size_t)b * T_len * D + d;
  float acc = 0.f;
#pragma unroll 4
  for (int t = 0; t < T_len; ++t) acc += pool_load<T>(src + (size_t)t * D);
  out[(size_t)b * D + d] = acc / (float)T_len;
}

cudaError_t launch_mean_pool(const void* hidden_dev, int src_dtype, int B, int T_len, int D, float* out_dev, cudaStream_t stream) {
  if (B <= 0 || T_len <= 0 || D <= 0) return cudaErrorInvalidValue;
  const dim3 grid((unsigned)((D + 255) / 256), (unsigned)B);
  if (src_dtype == RS_DTYPE_FP32) mean_pool_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(hidden_dev), T_len, D, out_dev);
  else if (src_dtype == RS_DTYPE_FP16) mean_pool_kernel<__half><<<grid, 256, 0, stream>>>(static_cast<const __half*>(hidden_dev), T_len, D, out_dev);
  else if (src_dtype == RS_DTYPE_BF16) mean_pool_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(hidden_dev), T_len, D, out_dev);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

}  // namespace rs
