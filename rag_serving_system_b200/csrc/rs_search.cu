// Fused score-and-select for the retriever hot path (replaces retriever.py:594-611 + :750-769
// of the reference: SGEMM + 5 elementwise temporaries + B full-length sorts).
//
//   fused_search_kernel    ONE launch per search.  Persistent (one CTA per SM), warp-specialised:
//                            warp 0     bulk-copy (TMA engine) producer + dynamic tile scheduler
//                            warp 1     single-thread tcgen05.mma issuer, accumulators in TMEM
//                            warps 2-5  prologue: fp32 queries -> swizzled Q image in shared memory (k-major,
//                                       per-group barriers); epilogue: scores from TMEM -> running top-k,
//                                       the B x N score matrix never exists in HBM
//                            tail       per-CTA lists -> arrival counter -> CTA q merges query q
//                                       (+ multi-GPU: peer stores over NVLink, epoch flags, final merge)
//                          Two orientations (template DOCM):
//                            doc-major   (B <= 32, k <= 16): the 128 documents of a tile are the MMA's M, the
//                                        queries its N; a thread holds the 32 query scores of its document
//                            query-major (everything else): queries are M (64 or 128 TMEM lanes, one per
//                                        query), documents N = 128; a thread owns one query's top-k
//   query_prep_kernel      stand-alone query prep (streamed-Q plans, B > 32, D > 1024, RS_INLINE_PREP=0)
//   merge_local_kernel, exchange_merge_kernel, merge_topk_kernel
//                          stand-alone merges: fallbacks of the tail, rs_merge_device
//
// Query-major operand note: only Bpad A rows exist in shared memory; the remaining rows of the M=64/128
// operand alias whatever follows in shared memory and produce garbage in TMEM lanes that are never read
// (rows of D are independent).  The same holds for the doc-major N operand when Bpad < 16 or 32.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>

#include "rs_kernels.h"
#include "rs_layout.h"
#include "rs_ptx.cuh"

namespace rs {

// ------------------------------------------------------------------------------------------
// query prep
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint16_t f32_to_bf16_rn(float x) {
  uint32_t u = __float_as_uint(x);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
__device__ __forceinline__ float bf16_to_f32(uint16_t b) { return __uint_as_float((uint32_t)b << 16); }
__device__ __forceinline__ uint16_t f32_to_f16_rn(float x) { return __half_as_ushort(__float2half_rn(x)); }
__device__ __forceinline__ float f16_to_f32(uint16_t b) { return __half2float(__ushort_as_half(b)); }

// Per-query power-of-two pre-scale: brings max|q| into [1,2) so the fp16 terms stay in range.  Exact
// (a power of two), so it never changes a result; exponents are clamped to what a normal float holds.
struct QScale {
  float up, down;  // 2^e, 2^-e
  int e;
};
__device__ __forceinline__ QScale q_scale_from_max(float mx) {
  QScale s;
  s.e = 0;
  if (mx > 0.f && mx < CUDART_INF_F) s.e = max(-126, min(126, 127 - (int)((__float_as_uint(mx) >> 23) & 0xffu)));
  s.up = __uint_as_float((uint32_t)(127 + s.e) << 23);
  s.down = __uint_as_float((uint32_t)(127 - s.e) << 23);
  return s;
}
// One query element -> its tensor-core term(s) in the storage type + the value they represent.
__device__ __forceinline__ void q_convert(float x, const QScale& sc, int dtype, int terms, uint16_t& t0, uint16_t& t1,
                                          float& rep) {
  const float xs = x * sc.up;
  t1 = 0;
  if (dtype == RS_DTYPE_BF16) {
    t0 = f32_to_bf16_rn(xs);
    const float h = bf16_to_f32(t0);
    rep = h;
    if (terms == 2) {
      t1 = f32_to_bf16_rn(xs - h);
      rep = h + bf16_to_f32(t1);
    }
  } else {
    t0 = f32_to_f16_rn(xs);
    const float h = f16_to_f32(t0);
    rep = h;
    if (terms == 2) {
      t1 = f32_to_f16_rn((xs - h) * (float)(1 << RS_FP16_SPLIT_SHIFT));
      rep = h + f16_to_f32(t1) * (1.0f / (float)(1 << RS_FP16_SPLIT_SHIFT));
    }
  }
  rep *= sc.down;
}

// One CTA per (padded) query row.  Writes terms x KC blocks of the Q image, the L2 norm of the
// represented query (retriever.py:602) and the power-of-two factor that undoes the scaling.
__global__ void __launch_bounds__(256) query_prep_kernel(const float* __restrict__ q, int B, int D, int bpad,
                                                          int dtype, int terms, int corpus_scale_exp,
                                                          uint8_t* __restrict__ qimg, float* __restrict__ qnorm,
                                                          float* __restrict__ qscale) {
  const int row = blockIdx.x;
  const int kc_count = D / RS_KCHUNK_ELEMS;
  __shared__ float red[256];
  const bool live = row < B;
  // 1. max |q| -> per-query power-of-two scale (keeps fp16 terms in range; exact)
  float mx = 0.f;
  if (live)
    for (int c = threadIdx.x; c < D; c += blockDim.x) mx = fmaxf(mx, fabsf(q[(size_t)row * D + c]));
  red[threadIdx.x] = mx;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + s]);
    __syncthreads();
  }
  const QScale qsc = q_scale_from_max(red[0]);
  __syncthreads();
  // 2. split into terms, write swizzled, accumulate the norm of the represented value
  float ss = 0.f;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    const float x = live ? q[(size_t)row * D + c] : 0.f;
    uint16_t t0, t1;
    float rep;
    q_convert(x, qsc, dtype, terms, t0, t1, rep);
    ss += rep * rep;
    const int kc = c >> 6;
    const uint32_t off = rs_block_offset((uint32_t)row, (uint32_t)(c & 63));
    const size_t blk = (size_t)bpad * RS_KCHUNK_BYTES;
    *reinterpret_cast<uint16_t*>(qimg + (size_t)kc * blk + off) = t0;
    if (terms == 2) *reinterpret_cast<uint16_t*>(qimg + (size_t)(kc_count + kc) * blk + off) = t1;
  }
  red[threadIdx.x] = ss;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    qnorm[row] = sqrtf(red[0]);
    qscale[row] = ldexpf(1.0f, -(qsc.e + corpus_scale_exp));
  }
}

// ---- in-kernel query prep (resident-Q mode) ------------------------------------------------------
// Four consecutive query elements -> packed 16-bit terms (hardware cvt.rn, same round-to-nearest-even as
// f32_to_bf16_rn / __float2half_rn for finite values) + the sum of squares of the represented values.
// Compile-time dtype/terms keep this run-once code compact (it executes from a cold instruction cache).
template <bool BF16, bool TWO>
__device__ __forceinline__ void q_convert4(const float4 x, const float up, const float down, uint2& t0, uint2& t1,
                                           float& ss) {
  const float xs[4] = {x.x * up, x.y * up, x.z * up, x.w * up};
  uint32_t p[2], l[2] = {0u, 0u};
  float h[4], rep[4];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    if (BF16) {
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p[j]) : "f"(xs[2 * j + 1]), "f"(xs[2 * j]));
      h[2 * j] = __uint_as_float(p[j] << 16);
      h[2 * j + 1] = __uint_as_float(p[j] & 0xffff0000u);
    } else {
      asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p[j]) : "f"(xs[2 * j + 1]), "f"(xs[2 * j]));
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&p[j]));
      h[2 * j] = f.x;
      h[2 * j + 1] = f.y;
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) rep[j] = h[j];
  if (TWO) {
    constexpr float kUp = (float)(1 << RS_FP16_SPLIT_SHIFT), kDown = 1.0f / (float)(1 << RS_FP16_SPLIT_SHIFT);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float d0 = xs[2 * j] - h[2 * j], d1 = xs[2 * j + 1] - h[2 * j + 1];
      if (BF16) {
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l[j]) : "f"(d1), "f"(d0));
        rep[2 * j] += __uint_as_float(l[j] << 16);
        rep[2 * j + 1] += __uint_as_float(l[j] & 0xffff0000u);
      } else {
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(l[j]) : "f"(d1 * kUp), "f"(d0 * kUp));
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&l[j]));
        rep[2 * j] += f.x * kDown;
        rep[2 * j + 1] += f.y * kDown;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float r = rep[j] * down;
    ss = fmaf(r, r, ss);
  }
  t0 = make_uint2(p[0], p[1]);
  t1 = make_uint2(l[0], l[1]);
}

// One warp preps the query rows wid, wid+4, ... (at most 8: bpad <= 32) of the resident Q image.
// The work is K-MAJOR: columns are processed in groups of 256 (4 k-chunks) across all of the warp's
// rows, and the warp arrives on the group's mbarrier as soon as its part of the group is in shared
// memory, so the MMA issuer starts on the first tile's leading k-chunks while later groups are still
// being converted - the document stream (already running) is never stalled behind the whole prep.
// bf16 terms need no range pre-scale (bf16 has fp32's exponent range and a power-of-two scale does not
// change the represented value), so that type converts in a single pass over global memory; fp16
// terms first take each row's max |q| (pass A) for the power-of-two scale.  Writes the L2 norm of the
// represented query (retriever.py:602) and the factor that undoes the scaling; values are identical
// to query_prep_kernel's.  Rows >= n_queries are zero-filled.
constexpr int kPrepGroupCols = 256;  // columns per readiness group = 4 k-chunks
template <bool BF16, bool TWO>
__device__ __noinline__ void prep_queries_warp(const float* __restrict__ q, int wid, int n_queries, int D, int bpad,
                                               int corpus_scale_exp, uint8_t* qimg_smem, float* s_qnorm,
                                               float* s_qscale, uint32_t group_bar0, int lane) {
  const uint32_t blk = (uint32_t)bpad * RS_KCHUNK_BYTES;
  const uint32_t term1 = (uint32_t)(D / RS_KCHUNK_ELEMS) * blk;
  float up[8], down[8];
  int e[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    up[i] = down[i] = 1.f;
    e[i] = 0;
  }
  if (!BF16) {
    // pass A: row maxima, two rows' loads in flight together
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      float m[2] = {0.f, 0.f};
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int row = wid + 4 * (i + r);
        if (row < n_queries) {
#pragma unroll 8
          for (int c = lane * 4; c < D; c += 128) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(q + (size_t)row * D + c));
            m[r] = fmaxf(m[r], fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const QScale sc = q_scale_from_max(__uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(m[r]))));
        up[i + r] = sc.up;
        down[i + r] = sc.down;
        e[i + r] = sc.e;
      }
    }
  }
  float ss[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int n_groups = (D + kPrepGroupCols - 1) / kPrepGroupCols;
  // float4 f = lane + 32*j (j = 0,1) of a group covers columns g*256 + 4f: k-chunk 4g + lane/16 + 2j, and
  // 8 contiguous bytes inside the block at a lane-constant offset
  const uint32_t lane_col = (uint32_t)((lane * 4) & 63);
#pragma unroll 1
  for (int g = 0; g < n_groups; ++g) {
    float4 v[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = wid + 4 * i;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int c = g * kPrepGroupCols + 4 * (lane + 32 * j);
        v[i][j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < n_queries && c < D) v[i][j] = __ldg(reinterpret_cast<const float4*>(q + (size_t)row * D + c));
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = wid + 4 * i;
      if (row < bpad) {
        const uint32_t off = rs_block_offset((uint32_t)row, lane_col) + (uint32_t)(lane >> 4) * blk;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          if (g * kPrepGroupCols + 4 * (lane + 32 * j) < D) {
            uint2 t0, t1;
            q_convert4<BF16, TWO>(v[i][j], up[i], down[i], t0, t1, ss[i]);
            uint8_t* dst = qimg_smem + off + (uint32_t)(4 * g + 2 * j) * blk;
            *reinterpret_cast<uint2*>(dst) = t0;
            if (TWO) *reinterpret_cast<uint2*>(dst + term1) = t1;
          }
        }
      }
    }
    fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
    __syncwarp();
    if (lane == 0) mbar_arrive(group_bar0 + 8u * (uint32_t)g);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss[i] += __shfl_xor_sync(0xffffffffu, ss[i], o);
    const int row = wid + 4 * i;
    if (lane == 0 && row < bpad) {
      s_qnorm[row] = sqrtf(ss[i]);
      s_qscale[row] = ldexpf(1.0f, -(e[i] + corpus_scale_exp));
    }
  }
}

// ------------------------------------------------------------------------------------------
// fused score + select
// ------------------------------------------------------------------------------------------
constexpr int kProducerWarp = 0;
constexpr int kMmaWarp = 1;
constexpr int kFirstEpiWarp = 2;
constexpr int kNumEpiWarps = 4;
constexpr int kThreads = (kFirstEpiWarp + kNumEpiWarps) * 32;  // 192
constexpr int kRegK = 16;                                       // register-resident list capacity
constexpr int kBufC = 32;                                       // k > kRegK: unsorted candidate buffer per query (p.bufc <= 32)
constexpr int kMaxListK = 128;                                  // k > kRegK: sorted list in smem, <= 4 slots per lane

// words per query of the k > kRegK selection state (odd: see the epilogue)
// lists_global: only the candidate buffers stay in shared memory (+ one list-sized scratch per warp)
__host__ __device__ inline int list_row_words(int k, int bufc, bool lists_global = false) {
  return (2 * ((lists_global ? 0 : k) + bufc)) | 1;
}

constexpr int kNumEpiWarpsC = 4;
struct SmemMap {
  uint32_t q_bytes, q_res_bytes, stage_off, bar_off, dump_off, list_off, total;
};
// q_resident: the whole Q image lives in shared memory for the kernel's lifetime (small batches);
// otherwise Q k-chunks travel through the ring next to the document blocks (large batches).
__host__ __device__ inline SmemMap make_smem_map(int bpad, int kc, int terms, int stages, int cps, int k,
                                                 bool reglist, bool q_resident, int qstages, int bufc = kBufC,
                                                 bool lists_global = false, bool docm = false) {
  SmemMap m;
  // resident: the whole Q image; streamed: a private ring of qstages slots of cps k-chunks each
  m.q_res_bytes = q_resident ? (uint32_t)terms * kc * bpad * RS_KCHUNK_BYTES : 0u;
  m.q_bytes = m.q_res_bytes + (!q_resident ? (uint32_t)qstages * cps * bpad * RS_KCHUNK_BYTES : 0u);
  m.stage_off = m.q_bytes;  // multiple of 1024 because bpad % 8 == 0
  m.bar_off = m.stage_off + (uint32_t)stages * (uint32_t)cps * RS_BLOCK_BYTES;
  m.dump_off = m.bar_off + 2048;                       // [bar_off, +1 KiB) barriers, then qnorm[128], qscale[128];
                                                       // dump = rare-path chunk: 32 columns x live rows, fp32
  m.list_off = m.dump_off + (uint32_t)(docm ? 128 : (bpad + 31) / 32 * 32) * 128u;   // whole warps of 32 (M=128) or 2x16 (M=64) rows;
                                                                                       // doc-major: 4 warps x [32 queries][32 docs]
  m.total = m.list_off + (reglist ? 0u : (uint32_t)list_row_words(k, bufc, lists_global) * (uint32_t)bpad * 4u +
                                             (lists_global ? (uint32_t)kNumEpiWarpsC * 2u * (uint32_t)k * 4u : 0u));
  return m;
}

__device__ __forceinline__ float max32(const float (&v)[32]) {
  float m0 = fmaxf(fmaxf(v[0], v[1]), v[2]), m1 = fmaxf(fmaxf(v[3], v[4]), v[5]);
  float m2 = fmaxf(fmaxf(v[6], v[7]), v[8]), m3 = fmaxf(fmaxf(v[9], v[10]), v[11]);
#pragma unroll
  for (int j = 12; j < 32; j += 4) {
    m0 = fmaxf(m0, v[j]);
    m1 = fmaxf(m1, v[j + 1]);
    m2 = fmaxf(m2, v[j + 2]);
    m3 = fmaxf(m3, v[j + 3]);
  }
  return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}

// Conservative bound, in raw accumulator units, below which a document of this tile cannot beat
// the current cosine threshold (see DESIGN.md "cosine epilogue").
__device__ __forceinline__ float cosine_raw_bound(float thr_cos, float nq, float inv_qs, float ndmin, float ndmax) {
  if (thr_cos == -CUDART_INF_F) return -CUDART_INF_F;
  if (thr_cos >= 0.f) return thr_cos * nq * ndmin * inv_qs * (1.0f - 1e-6f);
  // negative threshold: the product may underflow to -0 (zero query, a threshold just below 0 from the shared bound),
  // which a raw score of 0 does not beat although its cosine (0) beats the threshold: stay strictly negative
  const float r = thr_cos * (nq * ndmax + 1e-8f) * inv_qs * (1.0f + 1e-6f);
  return r < 0.f ? r : -1.17549435e-38f;
}

// k > kRegK selection state (per query one shared-memory row of list_row_words() words, odd stride):
// a sorted top-k list (best first) plus an unsorted buffer of bufc (16 or 32) candidates that passed
// the (stale) threshold.  The sorted candidates are parked in Bs[0..32): with bufc == 16 that runs
// over the (already consumed) id half of the buffer, which is exactly 16 more words.  When a buffer fills, the whole warp merges it into the list: bitonic-sort the
// 32 candidates across lanes, then every element computes its rank in the merged order by binary
// search in the other sequence and is written straight to that position.  Ties: list elements were
// seen earlier (lower index) so they win; candidates order by (score desc, index asc).
// Must be called by the full, converged warp.  Returns the new k-th best score.
// All four pointers address ONE query's arrays (unit stride).  Candidate scores are multiplied by `scale` (an
// exact power of two, or 1) on the way in, so a list may be kept in final units while the filter runs in raw
// accumulator units; the returned k-th best is in list units.
// merge_sorted_candidates: the same merge with the (at most 32) candidates handed in as registers, one per lane
// (lanes >= n: (-inf, 0x7fffffff)); Bs is a 32-float scratch for the sorted candidate scores.
__device__ __forceinline__ float merge_sorted_candidates(float* Ls, int* Li, float* Bs, float cs, int ci, int n, int k,
                                                         int lane) {
#pragma unroll
  for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
    for (int st = size >> 1; st > 0; st >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, cs, st);
      const int oi = __shfl_xor_sync(0xffffffffu, ci, st);
      const bool desc = (lane & size) == 0;  // size == 32: always descending
      const bool lower = (lane & st) == 0;
      const bool mine_better = cs > os || (cs == os && ci < oi);
      const bool keep = (lower == desc) ? mine_better : !mine_better;
      if (!keep && !(cs == os && ci == oi)) {
        cs = os;
        ci = oi;
      }
    }
  }
  Bs[lane] = cs;  // now sorted, best first
  __syncwarp();
  float es[kMaxListK / 32];
  int ei[kMaxListK / 32], er[kMaxListK / 32];
#pragma unroll
  for (int g = 0; g < kMaxListK / 32; ++g) {
    const int pos = lane + 32 * g;
    er[g] = k;
    if (pos < k) {
      es[g] = Ls[pos];
      ei[g] = Li[pos];
      // candidates strictly better than this list element: fixed-trip binary search (count in
      // [0,32]) so the four independent searches of a lane interleave their shared-memory loads
      int c = 0;
#pragma unroll
      for (int step = 16; step > 0; step >>= 1)
        if (Bs[c + step - 1] > es[g]) c += step;
      if (c == 31 && Bs[31] > es[g]) c = 32;
      er[g] = pos + c;
    }
  }
  int lo = 0, hi = k;  // list elements at least as good as this candidate
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (Ls[mid] >= cs) lo = mid + 1; else hi = mid;
  }
  const int cr = lane + lo;
  __syncwarp();
#pragma unroll
  for (int g = 0; g < kMaxListK / 32; ++g) {
    if (er[g] < k) {
      Ls[er[g]] = es[g];
      Li[er[g]] = ei[g];
    }
  }
  if (lane < n && cr < k) {
    Ls[cr] = cs;
    Li[cr] = ci;
  }
  __syncwarp();
  return Ls[k - 1];
}
__device__ __forceinline__ float flush_candidates(float* Ls, int* Li, float* Bs, const int* Bi, int n, int k, int lane,
                                                  float scale = 1.0f) {
  const float cs = lane < n ? Bs[lane] * scale : -CUDART_INF_F;
  const int ci = lane < n ? Bi[lane] : 0x7fffffff;
  return merge_sorted_candidates(Ls, Li, Bs, cs, ci, n, k, lane);
}

// ------------------------------------------------------------------------------------------
// merge of sorted candidate lists: one warp per query, each lane walks lists lane, lane+32, ...
// Ordering everywhere is (score desc, index asc); index < 0 is padding and only occurs at the end
// of a list.  STAGED: the query's candidates are first copied (coalesced) into shared memory so
// the k selection rounds run at shared-memory latency instead of k dependent global round trips.
// ------------------------------------------------------------------------------------------
constexpr int kMergeMaxOwn = 8;  // up to 256 lists

__device__ __forceinline__ unsigned int ld_volatile_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u32(unsigned int* p, unsigned int v) {
  asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

constexpr int kMergeThreads = 128;  // four warps stage one query's candidates; warp 0 runs the k rounds

// Copy one query's candidate lists into shared memory.  Element (list l, slot j) sits at
// l*list_stride + qoff + j; the loads are issued four deep so the L2 round trips overlap.
template <typename IdxT>
__device__ __forceinline__ void stage_lists(const float* __restrict__ cand_score, const IdxT* __restrict__ cand_idx,
                                            int n_lists, int kc, long long score_list_stride,
                                            long long idx_list_stride, size_t qoff, IdxT* s_idx, float* s_score,
                                            const int nthr) {
  const int total = n_lists * kc;
  for (int e0 = threadIdx.x; e0 < total; e0 += 4 * nthr) {
    IdxT ti[4];
    float ts[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + nthr * u;
      if (e < total) {
        const int l = e / kc, j = e - l * kc;
        ti[u] = __ldcg(cand_idx + (size_t)l * idx_list_stride + qoff + j);
        ts[u] = __ldcg(cand_score + (size_t)l * score_list_stride + qoff + j);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + nthr * u;
      if (e < total) {
        s_idx[e] = ti[u];
        s_score[e] = ts[u];
      }
    }
  }
  __syncthreads();
}

// k selection rounds by ONE warp over n_lists sorted lists (score desc, index asc; index < 0 is
// padding at the end of a list).  Every list head is kept as a sortable pair: ks = order-preserving
// image of the score (0 = padding / exhausted), ki = ~index (larger = lower index).  The warp's best
// head is found with three redux.sync steps (score, index high word, index low word).  `load(l, j,
// score&, idx&)` reads element j of list l.  Writes k rows to res_* (padding: -1 / -inf).
template <typename Load>
__device__ __forceinline__ void select_rounds(Load load, int n_lists, int kc, int k, long long idx_offset,
                                              float* res_s, long long* res_i, int lane) {
  unsigned int ks[kMergeMaxOwn];
  unsigned long long ki[kMergeMaxOwn];
  int head[kMergeMaxOwn] = {};
  auto read_head = [&](int i) {
    const int l = lane + 32 * i;
    ks[i] = 0u;
    ki[i] = 0ull;
    if (l < n_lists && head[i] < kc) {
      float sc;
      long long id;
      load(l, head[i], sc, id);
      if (id >= 0) {
        unsigned int u = __float_as_uint(sc);
        ks[i] = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
        ki[i] = ~(unsigned long long)id;
      }
    }
  };
#pragma unroll
  for (int i = 0; i < kMergeMaxOwn; ++i) {
    head[i] = 0;
    read_head(i);
  }
  int r = 0;
  for (; r < k; ++r) {
    unsigned int bs = 0u;  // this lane's best head
    unsigned long long bi = 0ull;
    int bown = 0;
#pragma unroll
    for (int i = 0; i < kMergeMaxOwn; ++i)
      if (ks[i] > bs || (ks[i] == bs && ki[i] > bi)) {
        bs = ks[i];
        bi = ki[i];
        bown = i;
      }
    const unsigned int m0 = __reduce_max_sync(0xffffffffu, bs);
    if (m0 == 0u) break;  // every list exhausted: the rest is padding
    bool mine = bs == m0;
    const unsigned int m1 = __reduce_max_sync(0xffffffffu, mine ? (unsigned int)(bi >> 32) : 0u);
    mine = mine && (unsigned int)(bi >> 32) == m1;
    const unsigned int m2 = __reduce_max_sync(0xffffffffu, mine ? (unsigned int)bi : 0u);
    mine = mine && (unsigned int)bi == m2;
    if (mine) {  // exactly one lane when indices are unique
      float sc;
      long long id;
      load(lane + 32 * bown, head[bown], sc, id);
      res_s[r] = sc;
      res_i[r] = id + idx_offset;
#pragma unroll
      for (int i = 0; i < kMergeMaxOwn; ++i)
        if (i == bown) {
          ++head[i];
          read_head(i);
        }
    }
  }
  __syncwarp();
  for (int rr = r + lane; rr < k; rr += 32) {
    res_s[rr] = -CUDART_INF_F;
    res_i[rr] = -1;
  }
  __syncwarp();
}


__device__ __forceinline__ void stamp(unsigned long long* ts, int slot) {
  if (ts) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    ts[(size_t)blockIdx.x * 8 + slot] = t;
  }
}
// Per-query shared-memory workspace of a merge: result row (k indices + k scores) then the staged lists.
struct MergeSmem {
  long long* res_i;
  float* res_s;
  int* slot;  // parallel selection: the (at most k) lists whose heads rank among the k best heads
  uint8_t* stage;
};
__host__ __device__ inline size_t merge_res_bytes(int k) {  // k results (8 + 4 B) + max(k, kMaxPeers) list slots
  return ((size_t)k * 12 + (size_t)(k > kMaxPeers ? k : kMaxPeers) * 4 + 15) & ~size_t(15);
}
__device__ __forceinline__ MergeSmem merge_smem_at(uint8_t* base, int k) {
  MergeSmem m;
  m.res_i = reinterpret_cast<long long*>(base);
  m.res_s = reinterpret_cast<float*>(m.res_i + k);
  m.slot = reinterpret_cast<int*>(m.res_s + k);
  m.stage = base + merge_res_bytes(k);
  return m;
}

// ---- parallel selection for k <= 16 (every thread of the CTA takes part; no sequential rounds) -------
// The global top-k of sorted lists can only come from the lists whose HEAD is among the k best heads.
//   1. one thread per list ranks its head among all heads (brute force over n_lists <= 256 keys)
//      and the k best lists are written to slot[rank]
//   2. one thread per element of those <= k lists computes its final rank = position in its own list
//      + number of better elements in each other active list (linear scan, stops once rank >= k)
//      and writes itself to res[rank].
// Keys are unique (every document sits in exactly one list), so ranks never collide.
constexpr int kParMaxK = 16;

// sortable 64-bit key of a (score, local row id >= 0) pair: larger = better (score desc, id asc); 0 = padding
__device__ __forceinline__ unsigned long long pack_key(float sc, int id) {
  const unsigned int u = __float_as_uint(sc);
  const unsigned int ks = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ((unsigned long long)ks << 32) | (unsigned int)(0x7fffffff - id);
}
__device__ __forceinline__ void unpack_key(unsigned long long key, float& sc, int& id) {
  const unsigned int ks = (unsigned int)(key >> 32);
  sc = __uint_as_float((ks & 0x80000000u) ? (ks & 0x7fffffffu) : ~ks);
  id = 0x7fffffff - (int)(unsigned int)key;
}

// stage one query's per-CTA lists (contiguous [list][kc] scores / int32 ids at qoff) as packed keys
__device__ __forceinline__ void stage_keys(const float* __restrict__ cand_score, const int* __restrict__ cand_idx,
                                           int total, size_t qoff, unsigned long long* s_key, const int nthr) {
  for (int e0 = threadIdx.x; e0 < total; e0 += 4 * nthr) {
    int ti[4];
    float ts[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + nthr * u;
      if (e < total) {
        ti[u] = __ldcg(cand_idx + qoff + e);
        ts[u] = __ldcg(cand_score + qoff + e);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + nthr * u;
      if (e < total) s_key[e] = ti[u] < 0 ? 0ull : pack_key(ts[u], ti[u]);
    }
  }
  __syncthreads();
}

// step 2 of the parallel selection over the lists named in slot[0..n_slots) (-1 = unused slot).
// valid(pos), better(pos_a, pos_b), emit(rank, pos) work on staged element positions l*kc + j.
template <typename Valid, typename Better, typename Emit>
__device__ __forceinline__ void rank_active_lists(const int* slot, int n_slots, int kc, int k, Valid valid, Better better,
                                                  Emit emit, const int nthr) {
  for (int e = threadIdx.x; e < n_slots * kc; e += nthr) {
    const int a = e / kc, j = e - a * kc;
    const int l = slot[a];
    if (l < 0) continue;
    const int pe = l * kc + j;
    if (!valid(pe)) continue;
    int rank = j;
    for (int b = 0; b < n_slots && rank < k; ++b) {
      const int lb = slot[b];
      if (lb < 0 || b == a) continue;
      int c = 0;
      while (c < kc && rank + c < k && valid(lb * kc + c) && better(lb * kc + c, pe)) ++c;
      rank += c;
    }
    if (rank < k) emit(rank, pe);
  }
}

// local merge of packed keys: n_lists lists of kc keys -> res (k rows, padded with -1 / -inf)
__device__ __forceinline__ void select_parallel_keys(const unsigned long long* s_key, int n_lists, int kc, int k,
                                                     long long idx_offset, const MergeSmem& m, const int nthr) {
  for (int j = threadIdx.x; j < k; j += nthr) {
    m.res_s[j] = -CUDART_INF_F;
    m.res_i[j] = -1;
    m.slot[j] = -1;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < n_lists; t += nthr) {
    const unsigned long long mine = s_key[t * kc];
    if (mine != 0ull) {
      int rank = 0;
#pragma unroll 4
      for (int u = 0; u < n_lists; ++u) rank += s_key[u * kc] > mine ? 1 : 0;
      if (rank < k) m.slot[rank] = t;
    }
  }
  __syncthreads();
  rank_active_lists(
      m.slot, k, kc, k, [&](int pos) { return s_key[pos] != 0ull; },
      [&](int a, int b) { return s_key[a] > s_key[b]; },
      [&](int rank, int pos) {
        float sc;
        int id;
        unpack_key(s_key[pos], sc, id);
        m.res_s[rank] = sc;
        m.res_i[rank] = (long long)id + idx_offset;
      },
      nthr);
  __syncthreads();
}

// Local merge of one query: this GPU's per-CTA lists ([query][cta][k] layout, local int32 row ids)
// -> m.res_* (global indices).  Called by the whole CTA; warp 0 runs the rounds.  Ends with a barrier.
__device__ __noinline__ void merge_query_local(const float* __restrict__ cand_score, const int* __restrict__ cand_idx,
                                             int n_lists, int k, long long idx_offset, int q, const MergeSmem m,
                                             int nthr, unsigned long long* ts = nullptr) {
  if (k <= kParMaxK) {
    unsigned long long* s_key = reinterpret_cast<unsigned long long*>(m.stage);
    stage_keys(cand_score, cand_idx, n_lists * k, (size_t)q * n_lists * k, s_key, nthr);
    if (threadIdx.x == 0) stamp(ts, 2);  // lists staged
    select_parallel_keys(s_key, n_lists, k, k, idx_offset, m, nthr);
    if (threadIdx.x == 0) stamp(ts, 3);  // selected
    return;
  }
  int* s_idx = reinterpret_cast<int*>(m.stage);
  float* s_score = reinterpret_cast<float*>(m.stage + (size_t)n_lists * k * sizeof(int));
  stage_lists<int>(cand_score, cand_idx, n_lists, k, k, k, (size_t)q * n_lists * k, s_idx, s_score, nthr);
  if (threadIdx.x == 0) stamp(ts, 2);  // lists staged
  if (threadIdx.x < 32) {
    auto load = [&](int l, int j, float& sc, long long& id) {
      sc = s_score[l * k + j];
      id = (long long)s_idx[l * k + j];
    };
    select_rounds(load, n_lists, k, k, idx_offset, m.res_s, m.res_i, threadIdx.x & 31);
  }
  if (threadIdx.x == 0) stamp(ts, 3);  // rounds done
  __syncthreads();
}

// Candidate exchange + final merge of one query (multi-GPU).  The local row (m.res_*) is stored into this
// rank's slot of EVERY GPU's exchange buffer through peer pointers (NVLink), one 16-byte entry
// {index, score, epoch} per store; then the CTA polls the G x k entries of query q in its LOCAL buffer until
// each shows this search's epoch, and merges the G world lists into m.res_*.  The epoch rides inside the
// entry (a 16-byte store is one NVLink write, a 16-byte load one L2 sector read), so there is no flag, no
// system-scope fence and no second round trip.  A CTA only ever waits for the CTAs of the same query on the
// other GPUs.  Buffers are double-buffered by epoch parity (host side).  A peer that never shows up raises
// err_flag (the host fails the call) and its slots are treated as empty.  Ends with a barrier.
__device__ __forceinline__ void st_entry(uint4* p, long long idx, float sc, unsigned int epoch) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((unsigned int)((unsigned long long)idx & 0xffffffffull)),
               "r"((unsigned int)((unsigned long long)idx >> 32)), "r"(__float_as_uint(sc)), "r"(epoch)
               : "memory");
}
__device__ __forceinline__ uint4 ld_entry(const uint4* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
constexpr unsigned long long kPeerTimeoutNs = 30ull * 1000ull * 1000ull * 1000ull;  // 30 s
__device__ __forceinline__ void bump_epoch(unsigned int* epoch) {  // never 0: zeroed buffers must not match
  const unsigned int e = __ldcg(epoch) + 1u;
  __stcg(epoch, e == 0u ? 1u : e);
}
__device__ __noinline__ void exchange_query(const ExchangeArgs& x, int q, int k, const MergeSmem m, int nthr,
                                            unsigned int* err_flag) {
  // the search counter lives in device memory (bumped by the last CTA of a search), so a captured CUDA graph
  // can be replayed: nothing that changes from search to search is baked into the kernel arguments
  const unsigned int epoch = __ldcg(x.epoch);
  const size_t qe = (size_t)(epoch & 1u) * (size_t)x.parity_entries + (size_t)(x.q0 + q) * k;
  const int total = k * x.world;
  for (int e = threadIdx.x; e < total; e += nthr) {  // fan-out
    const int g = e / k, j = e - g * k;
    st_entry(x.dest[g] + qe + j, m.res_i[j], m.res_s[j], epoch);
  }
  long long* s_idx = reinterpret_cast<long long*>(m.stage);
  float* s_score = reinterpret_cast<float*>(m.stage + (size_t)total * sizeof(long long));
  for (int e = threadIdx.x; e < total; e += nthr) {  // collect
    const int r = e / k, j = e - r * k;
    const uint4* src = x.local + (size_t)r * x.slot_entries + qe + j;
    uint4 v = ld_entry(src);
    // A peer process may legitimately be late by seconds (its host thread was busy); only a peer that does not
    // show up for kPeerTimeoutNs is treated as gone.
    unsigned long long t0 = 0;
    int spins = 0;
    while (v.w != epoch) {
      if (++spins > 64) {
        __nanosleep(200);
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        if (now - t0 > kPeerTimeoutNs) {
          *reinterpret_cast<volatile unsigned int*>(err_flag) = 2u;
          v = make_uint4(0xffffffffu, 0xffffffffu, 0xff800000u, epoch);  // (-1, -inf)
          break;
        }
      }
      v = ld_entry(src);
    }
    s_idx[e] = (long long)(((unsigned long long)v.y << 32) | (unsigned long long)v.x);
    s_score[e] = __uint_as_float(v.z);
  }
  __syncthreads();  // staged; every thread is done reading the local row out of m.res_*
  const int world = x.world;
  if (k <= kParMaxK) {
    for (int j = threadIdx.x; j < k; j += nthr) {
      m.res_s[j] = -CUDART_INF_F;
      m.res_i[j] = -1;
    }
    for (int g = threadIdx.x; g < kMaxPeers; g += nthr) m.slot[g] = g < world ? g : -1;
    __syncthreads();
    rank_active_lists(
        m.slot, world, k, k, [&](int pos) { return s_idx[pos] >= 0; },
        [&](int a, int b) { return s_score[a] > s_score[b] || (s_score[a] == s_score[b] && s_idx[a] < s_idx[b]); },
        [&](int rank, int pos) {
          m.res_s[rank] = s_score[pos];
          m.res_i[rank] = s_idx[pos];
        },
        nthr);
    __syncthreads();
    return;
  }
  if (threadIdx.x < 32) {
    auto load = [&](int l, int j, float& sc, long long& id) {
      sc = s_score[l * k + j];
      id = s_idx[l * k + j];
    };
    select_rounds(load, world, k, k, 0, m.res_s, m.res_i, threadIdx.x & 31);
  }
  __syncthreads();
}

__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ===================== tail of a fused search: candidate merge (+ multi-GPU exchange) inside the same launch ====
// Called by every thread of every CTA once the CTA's candidate lists are in cand_* (and fenced).
__device__ __forceinline__ void fused_tail(const SearchParams& p, uint8_t* smem, const bool worker) {
  // Worker CTAs publish their lists and arrive on a device counter; CTA q (q < n_queries) waits for all
  // n_work arrivals, merges query q's lists out of L2 (ring/Q shared memory is free by now) and - multi-GPU
  // - exchanges the row with the peer GPUs and merges the world lists.  Every CTA of the grid is resident
  // (grid <= SM count, one CTA per SM), so the wait cannot starve a CTA that has not started.  The last
  // merging CTA to leave resets the counters for the launch after the next one (launches on one index are
  // stream-ordered and alternate between two counter blocks).
  if (p.tail_mode != 0) {
    if (threadIdx.x == 0 && worker) {
      __threadfence();
      atomicAdd(p.sync, 1u);
      stamp(p.debug_ts, 4);  // candidates published, arrived
    }
    if ((int)blockIdx.x < p.n_queries) {
      if (threadIdx.x == 0) {
        long long spins = 0;
        while (ld_acquire_gpu_u32(p.sync) < (unsigned int)p.n_work) {
          __nanosleep(40);
          if (++spins > (1ll << 24)) {  // ~1 s: a CTA of this grid never ran; flag it instead of hanging the GPU
            *reinterpret_cast<volatile unsigned int*>(p.err_flag) = 1u;  // mapped host memory
            break;
          }
        }
        __threadfence();
        stamp(p.debug_ts, 5);  // every worker CTA has arrived
      }
      __syncthreads();
      const MergeSmem m = merge_smem_at(smem, p.k);
      for (int q = blockIdx.x; q < ((p.debug_flags & 4) ? 0 : p.n_queries); q += gridDim.x) {
        merge_query_local(p.cand_score, p.cand_idx, p.n_lists, p.k, p.idx_offset, q, m, kThreads, p.debug_ts);
        if (p.tail_mode == 2) exchange_query(p.x, q, p.k, m, kThreads, p.err_flag);
        for (int j = threadIdx.x; j < p.k; j += kThreads) {
          p.out_score[(size_t)q * p.k + j] = m.res_s[j];
          p.out_idx[(size_t)q * p.k + j] = m.res_i[j];
        }
        __syncthreads();
      }
      if (threadIdx.x == 0) {
        stamp(p.debug_ts, 6);  // this CTA's queries merged and written
        const unsigned int mergers = (unsigned int)min((int)gridDim.x, p.n_queries);
        if (atomicAdd(p.sync + 1, 1u) == mergers - 1u) {
          atomicExch(p.sync, 0u);
          atomicExch(p.sync + 1, 0u);
          atomicExch(p.sync + 4, 0u);  // tile counter of the dynamic scheduler
          if (p.tail_mode == 2 && p.x.last_pass) bump_epoch(p.x.epoch);  // every query of this search is exchanged
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Doc-major variant (B <= 32, k <= 16, resident Q; any storage type, one or two query terms, dot or cosine): DOCUMENTS are the
// MMA's M (128 per tile, one TMEM lane each) and the QUERIES its N (16 or 32 columns).  Against the
// query-major M=64 x N=128 MMA that is half the tensor work at B=32 (no padding rows), a quarter of the
// Q-operand reads, and 16-32 accumulator columns per tile instead of 128, i.e. up to 16 tiles of TMEM
// buffering.  Epilogue: warp w owns the tile's documents 32w..32w+31; a thread holds the 32 query
// scores of ITS document.  Thresholds are warp-uniform registers thr[j]; lane j of every warp keeps
// query j's running top-k for that warp's document slice in registers.  When anything passes, the
// 32x32 scores are parked in shared memory ([query][doc]) and every lane walks its own query's passing
// documents in ascending order - lane-parallel across queries, like the query-major rare path.  At the
// end the four warps' lists of a query are merged in shared memory into the CTA's one list.
// Cosine: a thread knows its document's norm, so it turns its 32 raw scores into exact cosines
// (retriever.py:605-609, same operation order as the query-major kernel) before comparing - no
// conservative per-tile bounds are needed.
// ------------------------------------------------------------------------------------------
struct DmBars {
  uint32_t tfull0, tempty0, sched_full0, sched_empty0, qfull;
  volatile int* tile_u;
  // CTA-pair variant (fused_search_dmpair_kernel): static schedule, accumulators handed back to the LEADER's issuer
  int pair, n_pairs, rank;
  uint32_t tempty_leader0;  // shared::cluster address of the leader's tempty[0]
};
template <int KS, int NACC, bool PAIR = false, int NH = 1>
__device__ __forceinline__ void dm_epilogue(const SearchParams& p, const DmBars& b, uint32_t tmem_base, int w, int lane,
                                            float* dump /* this warp's [32][32] */, float* s_qnorm, float* s_qscale,
                                            bool inline_prep, bool worker, float* out_s /* [4][32 NH][k] */, int* out_i) {
  // NH = halves of 32 queries (2: the pair kernel with N = 64).  A thread holds its document's scores of ONE half at a
  // time; lane j keeps the running top-k of queries j and 32 + j (its warp's document slice) in registers.
  constexpr int kSched = KS;
  const int k = p.k, nq = p.n_queries;
  const uint32_t lane_base = tmem_base + ((uint32_t)(w * 32) << 16);
  float thr[NH][32];
  float ls[NH][kRegK];
  int li[NH][kRegK];
  float mythr[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) {
#pragma unroll
    for (int j = 0; j < 32; ++j) thr[h][j] = 32 * h + j < nq ? -CUDART_INF_F : CUDART_INF_F;
#pragma unroll
    for (int j = 0; j < kRegK; ++j) {
      ls[h][j] = j < k ? -CUDART_INF_F : CUDART_INF_F;  // ascending, slot 0 = current k-th best
      li[h][j] = -1;
    }
    mythr[h] = 32 * h + lane < nq ? -CUDART_INF_F : CUDART_INF_F;
  }
  const bool cosine = p.metric == RS_METRIC_COSINE;
  if (inline_prep && worker) {
    mbar_wait(b.qfull, 0);
  } else if (worker) {
    // separate prep kernel: bring the per-query constants into shared memory (every warp writes the same
    // values, so no cross-warp synchronisation is needed before this warp reads them back)
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      s_qnorm[32 * h + lane] = 32 * h + lane < nq ? p.qnorm[32 * h + lane] : 0.f;
      s_qscale[32 * h + lane] = 32 * h + lane < nq ? p.qscale[32 * h + lane] : 1.f;
    }
    __syncwarp();
  }
  int acc = 0, sslot = 0;
  uint32_t acc_ph = 0, sph = 0;
  int pu = b.pair;  // PAIR: tile pair of this iteration
  while (true) {
    int t;
    if constexpr (PAIR) {
      if (!worker || pu >= (p.n_tiles + 1) / 2) break;
      t = 2 * pu + b.rank;  // an odd tile count leaves the peer's last tile empty: t == n_tiles, every document masked
      pu += b.n_pairs;
    } else {
      mbar_wait(b.sched_full0 + 8u * sslot, sph);
      t = b.tile_u[sslot];
      __syncwarp();
      if (lane == 0) mbar_arrive(b.sched_empty0 + 8u * sslot);
      if (++sslot == kSched) {
        sslot = 0;
        sph ^= 1u;
      }
      if (t < 0) break;
    }
    const int doc0 = t * RS_TILE_ROWS + w * 32;
    const float nd = (cosine && t < p.n_tiles) ? __ldg(p.dnorm + doc0 + lane) : 1.f;  // dnorm is padded to whole tiles
    mbar_wait(b.tfull0 + 8u * acc, acc_ph);
    tc_fence_after();
    const uint32_t abase = lane_base + (uint32_t)(acc * NACC * NH) * 32u;  // buffer: [accumulator][half][32 columns]
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      float v[32];
      {
        uint32_t r[32];
        tmem_ld_x32(abase + (uint32_t)h * 32u, r);
        if (NACC == 2) {  // hi + lo accumulators (two query terms and/or an fp32 corpus stored as two planes)
          uint32_t r1[32];
          tmem_ld_x32(abase + (uint32_t)(NH + h) * 32u, r1);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaf(__uint_as_float(r1[j]), p.lo_coef, __uint_as_float(r[j]));
        } else {
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        }
      }
      if (h == NH - 1) {
        // the (last) scores are in registers: hand the accumulator back before looking at them
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (PAIR) mbar_arrive_remote(b.tempty_leader0 + 8u * acc);
          else mbar_arrive(b.tempty0 + 8u * acc);
        }
        if (++acc == p.nbuf) {
          acc = 0;
          acc_ph ^= 1u;
        }
      }
      if (p.debug_flags & 2) continue;
      if (cosine) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float denom = __fadd_rn(__fmul_rn(s_qnorm[32 * h + j], nd), 1e-8f);      // retriever.py:605,607
          v[j] = fminf(fmaxf(__fdiv_rn(v[j] * s_qscale[32 * h + j], denom), -1.0f), 1.0f);  // :607,:609
        }
      }
      if (doc0 + 32 > p.n_rows && doc0 + lane >= p.n_rows) {  // ragged last tile: padded documents never pass
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = -CUDART_INF_F;
      }
      uint32_t pm = 0u;
#pragma unroll
      for (int j = 0; j < 32; ++j) pm |= (v[j] > thr[h][j]) ? (1u << j) : 0u;
      if (__any_sync(0xffffffffu, pm != 0u)) {
#pragma unroll
        for (int j = 0; j < 32; ++j) dump[j * 32 + lane] = v[j];  // [query][doc]: conflict-free
        uint32_t mine = 0u;  // documents (lanes) of this warp that pass for query 32 h + lane
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const uint32_t m = __ballot_sync(0xffffffffu, (pm >> j) & 1u);
          if (lane == j) mine = m;
        }
        __syncwarp();
        while (mine) {  // ascending document index: strict > keeps the lower index on ties
          const int d = __ffs(mine) - 1;
          mine &= mine - 1;
          const float x = dump[lane * 32 + d];
          if (x > mythr[h]) {
            const int idx = doc0 + d;
            bool prev = true;
#pragma unroll
            for (int jj = 0; jj < kRegK; ++jj) {
              const bool lt = (jj + 1 < kRegK) && (ls[h][jj + 1 < kRegK ? jj + 1 : jj] < x);
              const float ns = lt ? ls[h][jj + 1 < kRegK ? jj + 1 : jj] : (prev ? x : ls[h][jj]);
              const int ni = lt ? li[h][jj + 1 < kRegK ? jj + 1 : jj] : (prev ? idx : li[h][jj]);
              ls[h][jj] = ns;
              li[h][jj] = ni;
              prev = lt;
            }
            mythr[h] = ls[h][0];
          }
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; ++j) thr[h][j] = __shfl_sync(0xffffffffu, mythr[h], j);
      }
    }
  }
  // this warp's lists of queries 32 h + lane, best first, final units (exact power-of-two scale)
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    const int q = 32 * h + lane;
    if (q < nq && worker) {
      const float qs = cosine ? 1.0f : s_qscale[q];  // cosine lists are already in final units
      float* os = out_s + (size_t)(w * (32 * NH) + q) * k;
      int* oi = out_i + (size_t)(w * (32 * NH) + q) * k;
#pragma unroll
      for (int j = 0; j < kRegK; ++j) {
        if (j < k) {
          os[k - 1 - j] = li[h][j] < 0 ? -CUDART_INF_F : ls[h][j] * qs;
          oi[k - 1 - j] = li[h][j];
        }
      }
    }
  }
}

// Doc-major kernels: the four warps' lists of query q (best first, shared memory at offset 0: [4][32][k] scores, then
// ids) -> the CTA's one list in cand_* (list index = blockIdx.x).  Whole CTA; ends with a barrier.
__device__ __forceinline__ void dm_publish_cta_lists(const SearchParams& p, const uint8_t* smem, const bool worker) {
  // the four warps' lists of query q (best first, shared memory) -> the CTA's one list in cand_*
  const int NQ = p.bpad > 32 ? 64 : 32;  // query rows per warp block (dm_epilogue's NH * 32)
  if (worker && (int)threadIdx.x < p.n_queries) {
    const int q = threadIdx.x, k = p.k;
    const float* ss = reinterpret_cast<const float*>(smem);
    const int* si = reinterpret_cast<const int*>(smem + 4 * NQ * kRegK * sizeof(float));
    float* cs = p.cand_score + ((size_t)q * p.n_work + blockIdx.x) * k;
    int* ci = p.cand_idx + ((size_t)q * p.n_work + blockIdx.x) * k;
    int h[4] = {0, 0, 0, 0};
    for (int r = 0; r < k; ++r) {
      int bw = -1, bi = 0;
      float bs = 0.f;
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        if (h[w] < k) {
          const int id = si[(size_t)(w * NQ + q) * k + h[w]];
          const float sc = ss[(size_t)(w * NQ + q) * k + h[w]];
          if (id >= 0 && (bw < 0 || sc > bs || (sc == bs && id < bi))) {
            bw = w;
            bs = sc;
            bi = id;
          }
        }
      }
      cs[r] = bw < 0 ? -CUDART_INF_F : bs;
      ci[r] = bw < 0 ? -1 : bi;
#pragma unroll
      for (int w = 0; w < 4; ++w)
        if (w == bw) ++h[w];
    }
    if (p.tail_mode != 0) __threadfence();
  }
  __syncthreads();
}

// One CTA walks "supertiles" of p.sub consecutive 128-document tiles.  Per supertile the K loop
// is outermost: for every 64-element k-chunk the (optional) Q blocks arrive once and are reused
// by the MMAs of all p.sub document blocks, each accumulating into its own TMEM columns.
template <int NACC, bool REGLIST, bool QRES, bool DOCM = false>
__global__ void __launch_bounds__(kThreads, 1) fused_search_kernel(const __grid_constant__ SearchParams p) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operands need 1024-byte aligned shared memory; the launcher over-allocates by 1 KiB
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr bool q_resident = QRES;
  const SmemMap map = make_smem_map(p.bpad, p.kc, p.terms, p.stages, p.cps, p.k, REGLIST, q_resident, p.qstages, p.bufc,
                                    p.lists_global != 0, DOCM);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t q_smem = smem_base;
  const uint32_t stage_smem = smem_base + map.stage_off;
  const uint32_t bar_base = smem_base + map.bar_off;
  // barrier slots (8 B each): full[s], empty[s], tfull[8], tempty[8], qfull, qring_full[], qring_empty[], sched[]
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * p.stages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * p.stages + 8 + a); };
  const uint32_t qfull_bar = bar_base + 8u * (2 * p.stages + 16);
  auto qring_full = [&](int s) { return bar_base + 8u * (2 * p.stages + 17 + s); };              // streamed Q ring
  auto qring_empty = [&](int s) { return bar_base + 8u * (2 * p.stages + 17 + p.qstages + s); };
  // tile scheduler queue: the producer warp publishes the supertile ids this CTA works on (-1 = done)
  constexpr int kSched = 8;
  auto sched_full = [&](int s) { return bar_base + 8u * (2 * p.stages + 17 + 2 * p.qstages + s); };
  auto sched_empty = [&](int s) { return bar_base + 8u * (2 * p.stages + 17 + 2 * p.qstages + kSched + s); };
  volatile int* tile_u = reinterpret_cast<volatile int*>(smem + map.bar_off + 1016 - 4 * kSched);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + map.bar_off + 1016);

  // M=128: query row r lives in TMEM lane r.  M=64 (half the tensor work, used when bpad <= 64):
  // row r lives in lane (r/16)*32 + r%16, i.e. 16 live lanes in each of the four lane quarters.
  const int rows_per_warp = p.mma_m == 64 ? 16 : 32;
  const int nact_warps = DOCM ? kNumEpiWarps : (p.bpad + rows_per_warp - 1) / rows_per_warp;
  const int S = p.sub;     // 128-document tiles per supertile
  const int NB = p.nbuf;   // accumulator buffers in TMEM (>1: the epilogue overlaps the next supertiles' MMAs)
  // doc-major: 32 query columns per tile, NB (8 or 16) buffers
  const uint32_t tmem_cols = DOCM ? (uint32_t)(NB * NACC) * 32u : (uint32_t)(NB * S * NACC) * RS_TILE_ROWS;  // 128, 256 or 512
  const int n_super = (p.n_tiles + S - 1) / S;
  const uint32_t qblk = (uint32_t)p.bpad * RS_KCHUNK_BYTES;
  // CTAs >= n_work own no tiles: they exist only so that every query has a CTA for the tail merge
  const bool worker = (int)blockIdx.x < p.n_work;
  const bool inline_prep = QRES && p.inline_prep != 0;
  float* s_qnorm = reinterpret_cast<float*>(smem + map.bar_off + 1024);
  float* s_qscale = s_qnorm + 128;

  // Overlapped launches, full-machine grids (p.pdl == 2): release the dependent grid right away.  It can only
  // become resident as this grid's CTAs leave their SMs (one CTA per SM), so the next search's CTAs start
  // one by one the moment a worker CTA of this search exits - no wait for the slowest scanner, no launch
  // latency - and since its CTAs release THEIR dependents only once all of them are resident, i.e. once every
  // CTA of this grid has exited, at most two searches are ever in flight.
  if (p.pdl == 2) griddep_launch_dependents();
  if (threadIdx.x == 0) {
    stamp(p.debug_ts, 0);  // kernel entry
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 8; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), nact_warps);
    }
    mbar_init(qfull_bar, inline_prep ? kNumEpiWarps : 1);
    if (!q_resident)
      for (int s = 0; s < p.qstages; ++s) {
        mbar_init(qring_full(s), 1);
        mbar_init(qring_empty(s), 1);
      }
    if (inline_prep)  // readiness of each 256-column group of the Q image (reuses the streamed-Q slots)
      for (int g = 0; g < 4; ++g) mbar_init(qring_full(g), kNumEpiWarps);
    for (int s = 0; s < kSched; ++s) {
      mbar_init(sched_full(s), 1);
      mbar_init(sched_empty(s), 1 + nact_warps);  // read by the MMA warp and every active epilogue warp
    }
    fence_mbar_init();
    fence_proxy_async();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(smem_u32((const void*)tmem_slot), tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const size_t tile_bytes = (size_t)p.planes * p.kc * RS_BLOCK_BYTES;

  // Both single-issuer roles run their loops with the WHOLE warp converged (uniform control flow,
  // descriptors and counters stay in uniform registers); only the copy / MMA / commit instructions
  // themselves are predicated on lane 0.  The per-stage latency of these two loops is on the
  // critical path of the ring, so they are kept as short as possible.
  const int cps = p.cps;                                   // 64-element k-chunks per ring stage (1 or 2)
  const uint32_t stage_bytes = (uint32_t)cps * RS_BLOCK_BYTES;
  if (warp == kProducerWarp) {
    // ===================== bulk-copy producer =====================
    const uint64_t pol_stream = l2_policy_evict_first();
    const uint64_t pol_keep = l2_policy_evict_last();
    if (q_resident && !inline_prep && worker && lane == 0) {
      mbar_arrive_expect_tx(qfull_bar, map.q_res_bytes);
      for (uint32_t off = 0; off < map.q_res_bytes; off += 32768u) {
        uint32_t n = min(32768u, map.q_res_bytes - off);
        bulk_g2s(q_smem + off, p.qimg + off, n, qfull_bar, pol_keep);
      }
    }
    int s = 0;
    uint32_t ph = 0;
    auto push = [&](const void* src, uint32_t bytes, uint64_t pol) {
      mbar_wait(empty_bar(s), ph ^ 1u);
      if (lane == 0) {
        mbar_arrive_expect_tx(full_bar(s), bytes);
        bulk_g2s(stage_smem + (uint32_t)s * stage_bytes, src, bytes, full_bar(s), pol);
      }
      if (++s == p.stages) {
        s = 0;
        ph ^= 1u;
      }
    };
    int qs = 0;
    uint32_t qph = 0;
    auto push_q = [&](const void* src, uint32_t bytes) {
      mbar_wait(qring_empty(qs), qph ^ 1u);
      if (lane == 0) {
        mbar_arrive_expect_tx(qring_full(qs), bytes);
        bulk_g2s(q_smem + map.q_res_bytes + (uint32_t)qs * (uint32_t)cps * qblk, src, bytes, qring_full(qs), pol_keep);
      }
      if (++qs == p.qstages) {
        qs = 0;
        qph ^= 1u;
      }
    };
    // Supertiles are handed out dynamically: the first one is static (blockIdx.x), every further one comes
    // from a device-wide counter, so CTAs that stream faster take more tiles and all CTAs finish within
    // one tile time of each other.  The counter only grows, so a CTA still sees ascending document
    // indices (the tie rule relies on that).  RS_DYN_SCHED=0: static round-robin.
    int sslot = 0;
    uint32_t sph = 0;
    int u = (worker && (int)blockIdx.x < n_super) ? (int)blockIdx.x : -1;
    while (true) {
      mbar_wait(sched_empty(sslot), sph ^ 1u);
      if (lane == 0) {
        tile_u[sslot] = u;
        mbar_arrive(sched_full(sslot));
      }
      if (++sslot == kSched) {
        sslot = 0;
        sph ^= 1u;
      }
      if (u < 0) break;
      int nxt = u + (int)gridDim.x;
      if (p.dyn_sched && lane == 0) nxt = (int)atomicAdd(p.sync + 4, 1u) + (int)gridDim.x;  // consumed after the pushes
      const int nsub = min(S, p.n_tiles - u * S);
      const uint8_t* src = p.corpus + (size_t)u * S * tile_bytes;
      for (int kc = 0; kc < p.kc; kc += cps) {
        if (!q_resident)
          for (int j = 0; j < p.terms; ++j) push_q(p.qimg + (size_t)(j * p.kc + kc) * qblk, cps * qblk);
        for (int pl = 0; pl < p.planes; ++pl)
          for (int sub = 0; sub < nsub; ++sub)
            push(src + (size_t)sub * tile_bytes + (size_t)(pl * p.kc + kc) * RS_BLOCK_BYTES, stage_bytes, pol_stream);
      }
      nxt = __shfl_sync(0xffffffffu, nxt, 0);
      u = nxt < n_super ? nxt : -1;
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer (lane 0 issues) =====================
    const uint32_t idesc = DOCM ? umma_idesc_f16(p.dtype == RS_DTYPE_BF16, RS_TILE_ROWS, p.bpad <= 16 ? 16 : 32)
                                : umma_idesc_f16(p.dtype == RS_DTYPE_BF16, p.mma_m, RS_TILE_ROWS);
    if (q_resident && !inline_prep && worker) {
      mbar_wait(qfull_bar, 0);
      tc_fence_after();
    }
    bool q_pending = inline_prep;  // first tile: wait for each k-chunk group of the Q image as it is needed
    int s = 0;
    uint32_t ph = 0;
    int acc = 0;
    uint32_t acc_ph = 0;
    int qs = 0;
    uint32_t qph = 0;
    const bool skip_mma = (p.debug_flags & 1) != 0;  // timing experiment only
    // 16-byte units: one UMMA K step (16 elements x 2 B) advances the start address by 2
    auto issue = [&](uint32_t d, uint32_t a_smem, uint32_t b_smem, bool fresh) {
      if (lane == 0 && !skip_mma) {
        // doc-major: the document block is the M operand, the Q block the N operand
        const uint64_t a = umma_smem_desc(DOCM ? b_smem : a_smem), b = umma_smem_desc(DOCM ? a_smem : b_smem);
#pragma unroll
        for (int ks = 0; ks < RS_KCHUNK_ELEMS / RS_UMMA_K; ++ks)
          umma_f16(d, a + 2u * ks, b + 2u * ks, idesc, (fresh && ks == 0) ? 0u : 1u);
      }
    };
    int sslot = 0;
    uint32_t sph = 0;
    while (true) {
      mbar_wait(sched_full(sslot), sph);
      const int u = tile_u[sslot];
      __syncwarp();
      if (lane == 0) mbar_arrive(sched_empty(sslot));
      if (++sslot == kSched) {
        sslot = 0;
        sph ^= 1u;
      }
      if (u < 0) break;
      const int nsub = min(S, p.n_tiles - u * S);
      mbar_wait(tempty_bar(acc), acc_ph ^ 1u);
      tc_fence_after();
      const uint32_t dbase = DOCM ? tmem_base + (uint32_t)(acc * NACC) * 32u : tmem_base + (uint32_t)(acc * S * NACC) * RS_TILE_ROWS;
      constexpr uint32_t kAcc2 = DOCM ? 32u : (uint32_t)RS_TILE_ROWS;  // column offset of the low-order accumulator
      for (int kc = 0; kc < p.kc; kc += cps) {
        uint32_t qa0, qa1 = 0;
        int qslot0 = 0, qslot1 = 0;
        if (q_resident) {
          if (q_pending) {
            mbar_wait(qring_full(kc >> 2), 0);
            tc_fence_after();
          }
          qa0 = q_smem + (uint32_t)kc * qblk;
          qa1 = q_smem + (uint32_t)(p.kc + kc) * qblk;
        } else {
          mbar_wait(qring_full(qs), qph);
          qa0 = q_smem + (uint32_t)qs * (uint32_t)cps * qblk;
          qslot0 = qs;
          if (++qs == p.qstages) {
            qs = 0;
            qph ^= 1u;
          }
          if (p.terms == 2) {
            mbar_wait(qring_full(qs), qph);
            qa1 = q_smem + (uint32_t)qs * (uint32_t)cps * qblk;
            qslot1 = qs;
            if (++qs == p.qstages) {
              qs = 0;
              qph ^= 1u;
            }
          }
        }
        for (int pl = 0; pl < p.planes; ++pl) {
          for (int sub = 0; sub < nsub; ++sub) {
            mbar_wait(full_bar(s), ph);
            tc_fence_after();
            const uint32_t b_smem = stage_smem + (uint32_t)s * stage_bytes;
            const uint32_t d0 = dbase + (uint32_t)(sub * NACC) * RS_TILE_ROWS;
            for (int c = 0; c < cps; ++c) {
              const bool fresh = (kc + c) == 0;
              const uint32_t bc = b_smem + (uint32_t)c * RS_BLOCK_BYTES;
              if (pl == 0) {
                issue(d0, qa0 + (uint32_t)c * qblk, bc, fresh);
                if (NACC == 2) issue(d0 + kAcc2, qa1 + (uint32_t)c * qblk, bc, fresh);
              } else {
                // lo plane of an fp32 corpus: q_hi . d_lo joins the low-order accumulator
                issue(d0 + kAcc2, qa0 + (uint32_t)c * qblk, bc, false);
              }
            }
            __syncwarp();
            if (lane == 0) umma_commit(empty_bar(s));  // smem slot is free once these MMAs have read it
            if (++s == p.stages) {
              s = 0;
              ph ^= 1u;
            }
          }
        }
        if (!q_resident && lane == 0) {  // Q chunk(s) fully consumed by this k-chunk group
          umma_commit(qring_empty(qslot0));
          if (p.terms == 2) umma_commit(qring_empty(qslot1));
        }
      }
      q_pending = false;
      __syncwarp();
      if (lane == 0) umma_commit(tfull_bar(acc));  // accumulators complete -> epilogue
      if (++acc == NB) {
        acc = 0;
        acc_ph ^= 1u;
      }
    }
  } else {
    // ===================== epilogue: one thread per query row =====================
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, 32*quarter+32) are this warp's
    if (inline_prep && worker) {
      // single-launch search: the four epilogue warps (idle until the first accumulator is ready)
      // build the resident Q image while the producer warp already streams document tiles
      static_assert(kNumEpiWarps == 4, "prep_queries_warp strides rows by 4 warps");
      const int D = p.kc * RS_KCHUNK_ELEMS, wid = warp - kFirstEpiWarp;
      const bool bf = p.dtype == RS_DTYPE_BF16;
      const uint32_t gb = qring_full(0);
      if (p.terms == 2) {
        if (bf) prep_queries_warp<true, true>(p.q_f32, wid, p.n_queries, D, p.bpad, p.corpus_scale_exp, smem, s_qnorm, s_qscale, gb, lane);
        else prep_queries_warp<false, true>(p.q_f32, wid, p.n_queries, D, p.bpad, p.corpus_scale_exp, smem, s_qnorm, s_qscale, gb, lane);
      } else {
        if (bf) prep_queries_warp<true, false>(p.q_f32, wid, p.n_queries, D, p.bpad, p.corpus_scale_exp, smem, s_qnorm, s_qscale, gb, lane);
        else prep_queries_warp<false, false>(p.q_f32, wid, p.n_queries, D, p.bpad, p.corpus_scale_exp, smem, s_qnorm, s_qscale, gb, lane);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(qfull_bar);  // this warp's norms / scales are in shared memory
      if (threadIdx.x == kFirstEpiWarp * 32) stamp(p.debug_ts, 1);  // query prep done
    }
    if constexpr (DOCM) {
      const DmBars bars{tfull_bar(0), tempty_bar(0), sched_full(0), sched_empty(0), qfull_bar, tile_u, 0, 0, 0, 0u};
      // lists of the four warps go to shared memory at offset 0: by then the last accumulator is complete, so
      // the Q image and the ring are dead
      dm_epilogue<kSched, NACC>(p, bars, tmem_base, quarter, lane, reinterpret_cast<float*>(smem + map.dump_off) + quarter * 1024,
                          s_qnorm, s_qscale, inline_prep, worker, reinterpret_cast<float*>(smem),
                          reinterpret_cast<int*>(smem + 4 * 32 * kRegK * sizeof(float)));
    } else if (quarter < nact_warps) {
      const int row = quarter * rows_per_warp + lane;
      const bool live = lane < rows_per_warp && row < p.n_queries;
      const int k = p.k;
      const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
      const bool cosine = p.metric == RS_METRIC_COSINE;
      if (inline_prep && worker) mbar_wait(qfull_bar, 0);  // all four warps' rows (and norms) are in place
      const float qs = live ? (inline_prep ? s_qscale[row] : p.qscale[row]) : 1.f;  // raw accumulator -> dot (2^n)
      const float inv_qs = 1.0f / qs;
      const float nq = live ? (inline_prep ? s_qnorm[row] : p.qnorm[row]) : 0.f;
      const float lo_coef = p.lo_coef;

      // running top-k.  k <= kRegK: registers, ascending, slot 0 is the current k-th best (the
      // threshold).  Larger k: sorted list + candidate buffer in shared memory (flush_candidates).
      float ls[REGLIST ? kRegK : 1];
      int li[REGLIST ? kRegK : 1];
      // per query one contiguous row [list scores k][list ids k][buffer scores bufc][buffer ids bufc] with an
      // ODD word stride: lanes that walk one query's slots (flush) and lanes that touch the same slot of
      // 32 different queries (append) both spread over all shared-memory banks
      // lists_global: the sorted lists live where the per-CTA candidates end up anyway (cand_* in global
      // memory, i.e. L2), in FINAL units; a flush pulls one list into a per-warp scratch, merges there and
      // writes it back.  Only the candidate buffers stay in shared memory, which gives the ring its bytes
      // back (the shared bound leaves only ~15 flushes per query per CTA).
      const bool lg = !REGLIST && p.lists_global != 0;
      const int lstride = list_row_words(k, p.bufc, lg);
      float* Ls = reinterpret_cast<float*>(smem + map.list_off);
      int* Li = reinterpret_cast<int*>(Ls + k);
      float* Bs = lg ? Ls : Ls + 2 * k;
      int* Bi = reinterpret_cast<int*>(Bs + p.bufc);
      float* Ws = Ls + (size_t)lstride * p.bpad + (size_t)quarter * 2 * k;  // per-warp scratch: k scores, k ids
      int* Wi = reinterpret_cast<int*>(Ws + k);
      auto glist_s = [&](int qrow) { return p.cand_score + ((size_t)qrow * p.n_work + blockIdx.x) * k; };
      auto glist_i = [&](int qrow) { return p.cand_idx + ((size_t)qrow * p.n_work + blockIdx.x) * k; };
      // one query's buffer -> its list; returns the new k-th best in FILTER units (raw for dot)
      auto flush_query = [&](int ql, int n) -> float {
        const int qrow = quarter * rows_per_warp + ql;
        if (!lg) return flush_candidates(Ls + (size_t)qrow * lstride, Li + (size_t)qrow * lstride, Bs + (size_t)qrow * lstride,
                                         Bi + (size_t)qrow * lstride, n, k, lane);
        const float unit = cosine ? 1.0f : __shfl_sync(0xffffffffu, qs, ql);  // list units = final units
        float* gs = glist_s(qrow);
        int* gi = glist_i(qrow);
        for (int j = lane; j < k; j += 32) {
          Ws[j] = __ldcg(gs + j);
          Wi[j] = __ldcg(gi + j);
        }
        __syncwarp();
        const float nt = flush_candidates(Ws, Wi, Bs + (size_t)qrow * lstride, Bi + (size_t)qrow * lstride, n, k, lane, unit);
        for (int j = lane; j < k; j += 32) {
          __stcg(gs + j, Ws[j]);
          __stcg(gi + j, Wi[j]);
        }
        __syncwarp();
        return cosine ? nt : nt * __shfl_sync(0xffffffffu, inv_qs, ql);  // exact: qs is a power of two
      };
      int cnt = 0;
      // per warp: [32 columns][rows_per_warp lanes] floats
      float* dump = reinterpret_cast<float*>(smem + map.dump_off) + (size_t)quarter * 32 * rows_per_warp;
      if (REGLIST) {
#pragma unroll
        for (int j = 0; j < kRegK; ++j) {
          ls[j] = j < k ? -CUDART_INF_F : CUDART_INF_F;
          li[j] = -1;
        }
      } else {
        if (live && (!lg || worker)) {
          float* l0 = lg ? glist_s(row) : Ls + (size_t)row * lstride;
          int* i0 = lg ? glist_i(row) : Li + (size_t)row * lstride;
          for (int j = 0; j < k; ++j) {
            l0[j] = -CUDART_INF_F;
            i0[j] = -1;
          }
        }
        __syncwarp();
      }
      float thr = live ? -CUDART_INF_F : CUDART_INF_F;  // in metric units (dot: raw units)
      float thr_raw = thr;
      // k > 16: bound shared between the CTAs of this GPU.  Every CTA publishes the best score it has seen
      // for each query; t = min over the n_work CTAs of those maxima means at least n_work >= k documents
      // score >= t, so anything strictly below t cannot be in the top-k.  t is ~(ln n_work)/m in tail
      // probability after m documents per CTA, the CTA's own k-th best is k/m: ~18x fewer candidates at
      // k = 100.  Ties with t are kept (the filter uses the next float below t).  The per-CTA list is then
      // no longer the CTA's exact top-k, but everything it drops is provably outside the GPU's top-k.
      const bool shared_bound = !REGLIST && p.g_best != nullptr && worker;
      float best_seen = -CUDART_INF_F, gbound = -CUDART_INF_F;
      int tiles_done = 0, next_refresh = 1;
      bool sampled = false;  // the sample pass over this CTA's first accumulator has run

      int acc = 0;
      uint32_t acc_ph = 0;
      int sslot = 0;
      uint32_t sph = 0;
      while (true) {
        mbar_wait(sched_full(sslot), sph);
        const int u = tile_u[sslot];
        __syncwarp();
        if (lane == 0) mbar_arrive(sched_empty(sslot));
        if (++sslot == kSched) {
          sslot = 0;
          sph ^= 1u;
        }
        if (u < 0) break;
        const int nsub = min(S, p.n_tiles - u * S);
        mbar_wait(tfull_bar(acc), acc_ph);
        tc_fence_after();
        if (shared_bound && !sampled) {
          // ---- sample pass over the CTA's FIRST accumulator (as in the pair kernel): before a query has k candidates its
          // own k-th best filters nothing, so every document of the first tiles would be appended and merged - at a
          // 1.25M-row shard (66 tiles per CTA) that start-up cost made k = 100 searches 1.9x (B = 32) to 2.7x (B = 128)
          // slower than k = 10 ones.  Pass A only tracks the query's best score and publishes it; once every CTA has
          // published, the filter below runs against a bound that already rejects almost everything.
          sampled = true;
          if (!(p.debug_flags & 2)) {
            for (int sub = 0; sub < nsub; ++sub) {
              const uint32_t tacc = lane_base + (uint32_t)((acc * S + sub) * NACC) * RS_TILE_ROWS;
              const int tile_row0 = (u * S + sub) * RS_TILE_ROWS;
#pragma unroll 1
              for (int c0 = 0; c0 < RS_TILE_ROWS; c0 += 32) {
                uint32_t r0[32];
                tmem_ld_x32(tacc + c0, r0);
                float v[32];
                if (NACC == 2) {
                  uint32_t r1[32];
                  tmem_ld_x32(tacc + RS_TILE_ROWS + c0, r1);
                  tmem_wait_ld();
#pragma unroll
                  for (int j = 0; j < 32; ++j) v[j] = fmaf(__uint_as_float(r1[j]), lo_coef, __uint_as_float(r0[j]));
                } else {
                  tmem_wait_ld();
#pragma unroll
                  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r0[j]);
                }
                const int valid = p.n_rows - (tile_row0 + c0);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  float x = v[j];
                  if (cosine) {
                    const float nd = p.dnorm[min(tile_row0 + c0 + j, p.n_rows - 1)];
                    const float denom = __fadd_rn(__fmul_rn(nq, nd), 1e-8f);  // retriever.py:605,607
                    x = fminf(fmaxf(__fdiv_rn(x * qs, denom), -1.0f), 1.0f);   // retriever.py:609
                  }
                  if (j >= valid) x = -CUDART_INF_F;  // padded rows of a ragged last tile
                  best_seen = fmaxf(best_seen, x);
                }
              }
            }
          }
          if (live) {
            __stcg(p.g_best + (size_t)blockIdx.x * p.bpad + row, best_seen);  // -inf: nothing to offer
            unsigned int polls = 0;
            unsigned long long t0 = 0;
            while (true) {
              bool all = true;
#pragma unroll 8
              for (int c = 0; c < p.n_work; ++c) {
                const float v = __ldcg(p.g_best + (size_t)c * p.bpad + row);
                all = all && (v == v);
              }
              if (all) break;
              __nanosleep(100);
              if ((++polls & 0xffu) == 0u) {
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (t0 == 0) t0 = now;
                if (now - t0 > 2000000000ull) break;  // 2 s: a CTA of this grid is not running; go on without the bound
              }
            }
          }
          next_refresh = 0;  // read the bound right away
        }
        if (shared_bound && tiles_done >= next_refresh) {  // after the sample, then after 1, 2, 4, ... tiles, then every 64
          next_refresh = tiles_done < 64 ? 2 * tiles_done : tiles_done + 64;
          float mn = CUDART_INF_F;
          if (live) {
#pragma unroll 8
            for (int c = 0; c < p.n_work; ++c) {
              const float v = __ldcg(p.g_best + (size_t)c * p.bpad + row);
              mn = fminf(mn, v == v ? v : -CUDART_INF_F);  // NaN = that CTA has not published yet
            }
            if (mn > -CUDART_INF_F && mn < CUDART_INF_F) {
              gbound = __uint_as_float(__float_as_uint(mn) + (mn > 0.f ? -1 : (mn < 0.f ? 1 : 0)));  // next float below
              if (mn == 0.f) gbound = -1.401298464e-45f;
              if (gbound > thr) {
                thr = gbound;
                thr_raw = cosine ? CUDART_INF_F : thr;  // cosine: recomputed per tile just below
              }
            }
          }
        }
        tiles_done += nsub;
        for (int sub = 0; sub < ((p.debug_flags & 2) ? 0 : nsub); ++sub) {
          const int t = u * S + sub;
          float ndmin = 1.f, ndmax = 1.f;
          if (cosine) {
            float2 mm = p.tile_norm_minmax[t];
            ndmin = mm.x;
            ndmax = mm.y;
            thr_raw = live ? cosine_raw_bound(thr, nq, inv_qs, ndmin, ndmax) : CUDART_INF_F;
          }
          const uint32_t tacc = lane_base + (uint32_t)((acc * S + sub) * NACC) * RS_TILE_ROWS;
          const int tile_row0 = t * RS_TILE_ROWS;
#pragma unroll 1
          for (int c0 = 0; c0 < RS_TILE_ROWS; c0 += 32) {
            float v[32];
            {
              uint32_t r0[32];
              tmem_ld_x32(tacc + c0, r0);
              if (NACC == 2) {
                uint32_t r1[32];
                tmem_ld_x32(tacc + RS_TILE_ROWS + c0, r1);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fmaf(__uint_as_float(r1[j]), lo_coef, __uint_as_float(r0[j]));
              } else {
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r0[j]);
              }
            }
            const float m = max32(v);
            if (__any_sync(0xffffffffu, m > thr_raw)) {
              // rare path: some query of this warp has a candidate in this 32-document chunk.  The
              // chunk is parked in shared memory ([column][lane], conflict-free) so that every lane
              // can walk ITS OWN passing columns with a dynamic index; the warp iterates
              // max-over-lanes(popcount) times, not once per column that passes for anybody.
              uint32_t pm = 0;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                pm |= (v[j] > thr_raw) ? (1u << j) : 0u;
                if (lane < rows_per_warp) dump[j * rows_per_warp + lane] = v[j];
              }
              if (tile_row0 + c0 + 32 > p.n_rows) {  // ragged last tile: mask padded rows
                const int valid = p.n_rows - (tile_row0 + c0);
                pm = valid <= 0 ? 0u : (pm & (0xffffffffu >> (32 - valid)));
              }
              __syncwarp();
              while (__any_sync(0xffffffffu, pm != 0u)) {
                if (pm != 0u) {
                  const int j = __ffs(pm) - 1;
                  pm &= pm - 1;
                  const float s = dump[j * rows_per_warp + lane];
                  const int idx = tile_row0 + c0 + j;
                  float x = s;  // dot metric: compare and store raw units (exact power-of-two scale)
                  if (cosine) {
                    const float nd = p.dnorm[idx];
                    const float denom = __fadd_rn(__fmul_rn(nq, nd), 1e-8f);   // retriever.py:605,607
                    x = __fdiv_rn(s * qs, denom);
                    x = fminf(fmaxf(x, -1.0f), 1.0f);                          // retriever.py:609
                  }
                  if (x > thr) {
                    if (REGLIST) {
                      bool prev = true;  // "slot j-1 was shifted" (true for j == 0)
#pragma unroll
                      for (int jj = 0; jj < kRegK; ++jj) {
                        const bool lt = (jj + 1 < kRegK) && (ls[jj + 1 < kRegK ? jj + 1 : jj] < x);
                        const float ns = lt ? ls[jj + 1 < kRegK ? jj + 1 : jj] : (prev ? x : ls[jj]);
                        const int ni = lt ? li[jj + 1 < kRegK ? jj + 1 : jj] : (prev ? idx : li[jj]);
                        ls[jj] = ns;
                        li[jj] = ni;
                        prev = lt;
                      }
                      thr = ls[0];
                      thr_raw = cosine ? cosine_raw_bound(thr, nq, inv_qs, ndmin, ndmax) : thr;
                    } else {
                      Bs[(cnt) + row * lstride] = x;  // O(1) append; threshold stays stale until the flush
                      Bi[(cnt) + row * lstride] = idx;
                      ++cnt;
                      if (p.debug_ts) atomicAdd(p.debug_ts + (size_t)blockIdx.x * 8 + 7, 1ull);  // appends (low word)
                      if (shared_bound && x > best_seen) {
                        best_seen = x;
                        __stcg(p.g_best + (size_t)blockIdx.x * p.bpad + row, x);
                      }
                    }
                  }
                }
                if (!REGLIST) {
                  uint32_t full = __ballot_sync(0xffffffffu, cnt == p.bufc);
                  while (full) {
                    const int ql = __ffs(full) - 1;
                    full &= full - 1;
                    const float nt = flush_query(ql, p.bufc);
                    if (p.debug_ts && lane == 0) atomicAdd(p.debug_ts + (size_t)blockIdx.x * 8 + 7, 1ull << 32);  // flushes
                    if (lane == ql) {
                      thr = fmaxf(nt, gbound);
                      cnt = 0;
                      thr_raw = cosine ? cosine_raw_bound(thr, nq, inv_qs, ndmin, ndmax) : thr;
                    }
                  }
                }
              }
              __syncwarp();
            }
          }
        }
        // all TMEM reads of this accumulator buffer are complete (wait::ld above) -> hand it back
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
        if (++acc == NB) {
          acc = 0;
          acc_ph ^= 1u;
        }
      }

      // a worker CTA that got no tile at all still owes the other CTAs' sample pass its entry (-inf: nothing to offer)
      if (shared_bound && !sampled && live) __stcg(p.g_best + (size_t)blockIdx.x * p.bpad + row, -CUDART_INF_F);
      if (!REGLIST) {  // drain the partially filled buffers
        uint32_t rem = __ballot_sync(0xffffffffu, cnt > 0);
        while (rem) {
          const int ql = __ffs(rem) - 1;
          rem &= rem - 1;
          const int n = __shfl_sync(0xffffffffu, cnt, ql);
          flush_query(ql, n);
        }
      }
      // per-CTA candidates, best first; dot scores leave raw units here
      griddep_wait();  // (overlapped launches) the previous search is done with cand_*
      if (live && worker) {
        float* cs = p.cand_score + ((size_t)row * p.n_work + blockIdx.x) * k;   // [query][cta][k]
        int* ci = p.cand_idx + ((size_t)row * p.n_work + blockIdx.x) * k;
        if (REGLIST) {
#pragma unroll
          for (int j = 0; j < kRegK; ++j) {
            if (j < k) {
              const float sc = cosine ? ls[j] : ls[j] * qs;
              cs[k - 1 - j] = li[j] < 0 ? -CUDART_INF_F : sc;
              ci[k - 1 - j] = li[j];
            }
          }
        } else if (!lg) {  // (lists_global: the lists already sit in cand_*, in final units)
          for (int j = 0; j < k; ++j) {
            const float raw = Ls[(j) + row * lstride];
            const int id = Li[(j) + row * lstride];
            cs[j] = id < 0 ? -CUDART_INF_F : (cosine ? raw : raw * qs);
            ci[j] = id;
          }
        }
      }
    }
  }

  if (!DOCM && p.tail_mode != 0) __threadfence();  // this thread's candidate stores are visible GPU-wide before the arrival
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }

  // Overlapped launches (p.pdl): from here on this CTA only publishes and merges, so the NEXT search's grid may
  // start filling the SMs that worker CTAs are about to leave.  Every thread first waits for the previous
  // search to be complete (its merging CTAs read the workspaces this one is about to reuse); a CTA therefore
  // releases its dependents only after its own predecessor has finished, which keeps at most two searches in
  // flight (the two blocks of sync counters alternate).  Without the launch attribute both are no-ops.
  griddep_wait();
  if (p.pdl != 2) griddep_launch_dependents();

  if constexpr (DOCM) dm_publish_cta_lists(p, smem, worker);
  fused_tail(p, smem, worker);
}

// Selection state of the wide-batch (CTA-pair) kernel: per (query, list) a 32-entry candidate buffer in global memory
constexpr int kWideBufC = 32;     // candidate-buffer entries per (query, pair)
constexpr int kWideFlushAt = 20;  // deferred flush once a buffer holds this many (a full one is flushed on the spot)

// ==========================================================================================
// Wide-batch kernel (128 < B <= 256 queries per pass share ONE pass over the shard; one query term, 16-bit storage;
// BASELINE config 5: B=256, k=100, the tensor-bound regime): the two SMs of a TPC run as a thread-block cluster of 2 and
// share ONE document stream through tcgen05.mma.cta_group::2.  (The general kernel holds one query per TMEM lane and
// would stream the shard once per 128 queries; a one-CTA form of this kernel - both 128-row halves against N=256 tile
// pairs - needs all 512 TMEM columns for one supertile, so its MMAs and its epilogue alternate: 8.0 ms per 12.5M-row
// slice where this kernel takes 5.9-6.2.  It was removed.)
//   * CTA rank r of a pair owns query rows 128r .. 128r+127 (its half of the MMA's M = 256) and loads ONE tile
//     of the pair's supertile (tiles 2u and 2u+1 = the two halves of the MMA's N = 256): per k-chunk a CTA
//     brings 16 KiB of documents from HBM and 16 KiB of its own query rows from L2 into one 32 KiB ring slot;
//     the tensor cores read the other CTA's document half through the pair link.  L2->SM traffic of Q equals the
//     document traffic.
//   * One thread of the leader (rank 0) issues M=256 x N=256 x K=16 MMAs (128 tensor cycles each, the full rate
//     of both SMs; 8 KiB of shared-memory operand reads per CTA per MMA = half the limit).  Each CTA's TMEM
//     receives ITS 128 query rows x 256 documents = 256 columns, so the 512 columns hold TWO accumulators and
//     the epilogue of supertile i overlaps the MMAs of supertile i+1 (the single-CTA wide kernel needs all 512
//     columns for one supertile and serialises the two).
//   * Barriers: full[s] of the leader counts its own bulk copies AND a relay arrive from the peer's (otherwise
//     idle) MMA warp once the peer's copies have landed; empty[s] and tfull[b] are signalled in both CTAs by
//     multicast tcgen05.commit; the epilogue warps of both CTAs arrive on the leader's tempty[b].
//   * Static schedule: pair P takes supertiles P, P + n_pairs, ... (ascending document order per CTA, which the
//     tie rule needs).  An odd tile count leaves the last supertile with one tile: rank 1 re-reads tile 2u and
//     nobody looks at columns 128-255.
//   * Epilogue: a thread owns ONE query; selection state in global memory (L2): the sorted list is the per-pair candidate
//     array the tail merges anyway (cand_*, final units), the 32-entry candidate buffer is cbuf; shared memory is all
//     ring.  The thread filters while it holds the accumulator and merges full buffers after releasing it; TMEM loads are software-pipelined (the next 32 columns are in flight while these are compared).
//     Candidate lists per query = number of worker pairs (p.n_lists).  Shared bound: a CTA publishes its
//     bound_rank-th best score (1 or 2), so that bound_rank * n_lists >= k documents are known to reach the bound.
// Every wait has a watchdog (a few seconds of wall clock, then the error word is raised and the kernel traps):
// a protocol bug or a lost peer must never hang the GPU.
// ==========================================================================================
constexpr int kPairSlots = 6;
constexpr uint32_t kPairSlotBytes = 2u * RS_BLOCK_BYTES;  // [16 KiB query rows of this CTA][16 KiB documents of this CTA's tile]
struct PairSmem {
  uint32_t bar_off, dump_off, scratch_off, total;
};
__host__ __device__ inline PairSmem make_pair_smem(int k) {
  PairSmem m;
  m.bar_off = (uint32_t)kPairSlots * kPairSlotBytes;
  m.dump_off = m.bar_off + 1024u;
  m.scratch_off = m.dump_off + 4u * 32u * 32u * 4u;
  m.total = m.scratch_off + 4u * (2u * (uint32_t)k + 64u) * 4u;
  return m;
}

template <bool CLUSTER>
__device__ __forceinline__ void mbar_wait_guarded(uint32_t bar, uint32_t parity, unsigned int* err_flag) {
  unsigned int polls = 0;
  unsigned long long t0 = 0;
  while (!(CLUSTER ? mbar_try_wait_cluster(bar, parity) : mbar_try_wait(bar, parity))) {
    if ((++polls & 0xfffu) == 0u) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      if (now - t0 > 4000000000ull) {  // 4 s
        *reinterpret_cast<volatile unsigned int*>(err_flag) = 3u;
        __threadfence_system();
        __trap();
      }
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1) fused_search_pair_kernel(const __grid_constant__ SearchParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();          // 0 = leader (issues the MMAs), 1 = peer
  const int pair = (int)blockIdx.x >> 1;
  const int n_pairs = (int)gridDim.x >> 1;
  const PairSmem map = make_pair_smem(p.k);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_base + map.bar_off;
  // a ring slot holds cps (1 or 2) consecutive k-chunks, each [16 KiB query rows][16 KiB documents]; n_slots x cps = 6
  const int cps = p.cps, n_slots = kPairSlots / cps;
  const uint32_t slot_bytes = (uint32_t)cps * kPairSlotBytes;
  auto slot_addr = [&](int s) { return smem_base + (uint32_t)s * slot_bytes; };
  auto full = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto empty = [&](int s) { return bar_base + 8u * (uint32_t)(kPairSlots + s); };  // (indices sized for cps == 1)
  auto tfull = [&](int b) { return bar_base + 8u * (uint32_t)(2 * kPairSlots + b); };
  auto tempty = [&](int b) { return bar_base + 8u * (uint32_t)(2 * kPairSlots + 2 + b); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + map.bar_off + 512);
  const int n_super = (p.n_tiles + 1) / 2;
  const bool worker = pair < n_super && (int)blockIdx.x < p.n_work;
  const size_t tile_bytes = (size_t)p.kc * RS_BLOCK_BYTES;
  constexpr uint32_t kTmemCols = 512;

  if (threadIdx.x == 0) {
    stamp(p.debug_ts, 0);
    for (int s = 0; s < n_slots; ++s) {
      mbar_init(full(s), rank == 0 ? 2u : 1u);  // leader: its own copies + the peer's relay
      mbar_init(empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull(b), 1);
      mbar_init(tempty(b), 2 * kNumEpiWarps);   // used in the leader only: the epilogue warps of both CTAs
    }
    fence_mbar_init();
    fence_proxy_async();
  }
  if (warp == kMmaWarp) {
    tmem_alloc_pair(smem_u32((const void*)tmem_slot), kTmemCols);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers are initialised before anybody signals across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kProducerWarp) {
    // ===================== bulk-copy producer (both CTAs) =====================
    if (worker) {
      const uint64_t pol_stream = l2_policy_evict_first();
      const uint64_t pol_keep = l2_policy_evict_last();
      uint64_t pol_normal;
      asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_normal));
      const bool no_hint = (p.debug_flags & 64) != 0;  // (experiment) documents without the evict_first hint
      int s = 0;
      uint32_t ph = 0;
      for (int u = pair; u < n_super; u += n_pairs) {
        const int nsub = min(2, p.n_tiles - 2 * u);
        const int tile = 2 * u + ((int)rank < nsub ? (int)rank : 0);
        const uint8_t* dsrc = p.corpus + (size_t)tile * tile_bytes;
        const uint8_t* qsrc = p.qimg + (size_t)rank * RS_BLOCK_BYTES;
        for (int kc = 0; kc < p.kc; kc += cps) {
          mbar_wait_guarded<false>(empty(s), ph ^ 1u, p.err_flag);
          if (lane == 0) {
            mbar_arrive_expect_tx(full(s), slot_bytes);
            for (int c = 0; c < cps; ++c) {
              bulk_g2s(slot_addr(s) + (uint32_t)c * kPairSlotBytes, qsrc + (size_t)(kc + c) * kPairSlotBytes, RS_BLOCK_BYTES, full(s), pol_keep);
              bulk_g2s(slot_addr(s) + (uint32_t)c * kPairSlotBytes + RS_BLOCK_BYTES, dsrc + (size_t)(kc + c) * RS_BLOCK_BYTES, RS_BLOCK_BYTES,
                       full(s), no_hint ? pol_normal : pol_stream);
            }
          }
          if (++s == n_slots) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
      // drain: the last multicast commits must have landed in this CTA's empty barriers before the pair may part
      for (int i = 0; i < n_slots; ++i) {
        mbar_wait_guarded<false>(empty(s), ph ^ 1u, p.err_flag);
        if (++s == n_slots) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
  } else if (warp == kMmaWarp) {
    if (worker && rank == 0) {
      // ===================== MMA issuer (leader): per k-chunk 4 MMAs of M=256 x N=256 x K=16 across the pair =====================
      const uint32_t idesc = umma_idesc_f16(p.dtype == RS_DTYPE_BF16, 256, 256);
      const bool skip_mma = (p.debug_flags & 1) != 0;
      int s = 0, b = 0;
      uint32_t ph = 0, acc_ph = 0;  // bit b of acc_ph = phase of accumulator b
      for (int u = pair; u < n_super; u += n_pairs) {
        mbar_wait_guarded<true>(tempty(b), ((acc_ph >> b) & 1u) ^ 1u, p.err_flag);
        tc_fence_after();
        for (int kc = 0; kc < p.kc; kc += cps) {
          mbar_wait_guarded<true>(full(s), ph, p.err_flag);
          tc_fence_after();
          if (lane == 0) {
            if (!skip_mma) {
              for (int c = 0; c < cps; ++c) {
                const uint32_t base = slot_addr(s) + (uint32_t)c * kPairSlotBytes;
                const uint64_t a = umma_smem_desc(base), bd = umma_smem_desc(base + RS_BLOCK_BYTES);
#pragma unroll
                for (int ks = 0; ks < RS_KCHUNK_ELEMS / RS_UMMA_K; ++ks)
                  umma_f16_pair(tmem_base + (uint32_t)b * 256u, a + 2u * ks, bd + 2u * ks, idesc, (kc + c == 0 && ks == 0) ? 0u : 1u);
              }
            }
            umma_commit_pair(empty(s));
          }
          __syncwarp();
          if (++s == n_slots) {
            s = 0;
            ph ^= 1u;
          }
        }
        if (lane == 0) umma_commit_pair(tfull(b));
        __syncwarp();
        acc_ph ^= 1u << b;
        b ^= 1;
      }
    } else if (worker) {
      // ===================== relay (peer): tell the leader when this CTA's half of a slot has landed =====================
      int s = 0;
      uint32_t ph = 0;
      const uint32_t full_leader = mapa_u32(full(0), 0);
      for (int u = pair; u < n_super; u += n_pairs) {
        for (int kc = 0; kc < p.kc; kc += cps) {
          mbar_wait_guarded<false>(full(s), ph, p.err_flag);
          // (a .release.cluster arrive here costs ~1 us each and serialises the relay: 3.2 ms instead of 1.2 ms for
          // the bare stream of 4M rows; the bulk copies' completion on full(s) already orders the data)
          if (lane == 0) mbar_arrive_remote(full_leader + 8u * (uint32_t)s);
          __syncwarp();
          if (++s == n_slots) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    }
  } else {
    // ===================== epilogue (both CTAs): a thread owns query row 128*rank + 32*quarter + lane =====================
    const int quarter = warp & 3;
    const int k = p.k;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const bool cosine = p.metric == RS_METRIC_COSINE;
    const int row = (int)rank * 128 + quarter * 32 + lane;
    const bool live = row < p.n_queries;
    const bool warp_live = (int)rank * 128 + quarter * 32 < p.n_queries && !(p.debug_flags & 2);
    const float qs = live ? p.qscale[row] : 1.f;
    const float inv_qs = 1.0f / qs;
    const float nq = live ? p.qnorm[row] : 0.f;
    float thr = (live && !(p.debug_flags & 32)) ? -CUDART_INF_F : CUDART_INF_F, thr_raw = thr;  // flag 32: nothing ever passes (timing)
    float best1 = -CUDART_INF_F, best2 = -CUDART_INF_F, gbound = -CUDART_INF_F;
    int cnt = 0;
    const int n_lists = p.n_lists;
    float* dump = reinterpret_cast<float*>(smem + map.dump_off) + (size_t)quarter * 32 * 32;
    float* Ws = reinterpret_cast<float*>(smem + map.scratch_off) + (size_t)quarter * (2 * k + 64);
    int* Wi = reinterpret_cast<int*>(Ws + k);
    float* Bs = Ws + 2 * k;
    auto glist_s = [&](int qrow) { return p.cand_score + ((size_t)qrow * n_lists + pair) * k; };
    auto glist_i = [&](int qrow) { return p.cand_idx + ((size_t)qrow * n_lists + pair) * k; };
    auto gbuf = [&](int qrow) { return p.cbuf + ((size_t)qrow * n_lists + pair) * kWideBufC; };
    if (worker && live) {
      float* l0 = glist_s(row);
      int* i0 = glist_i(row);
      for (int j = 0; j < k; ++j) {
        l0[j] = -CUDART_INF_F;
        i0[j] = -1;
      }
    }
    __syncwarp();
    const bool shared_bound = p.g_best != nullptr && worker;
    const int bound_rank = p.bound_rank;
    constexpr int flush_at = kWideFlushAt;  // (28 or 32 instead of 20 measured equal: 6.54 / 6.56 vs 6.61 ms)
    int supers_done = 0, next_refresh = 1;
    bool count_bound = true;  // false while the first supertile is filtered: the sample pass has counted its documents
    // merge query ql of this warp: its n buffered candidates -> its sorted list; returns the new k-th best in FILTER
    // units (raw accumulator units for dot).  Whole warp, converged.
    auto flush_query = [&](int ql, int n) -> float {
      const int qrow = (int)rank * 128 + quarter * 32 + ql;
      const float unit = cosine ? 1.0f : __shfl_sync(0xffffffffu, qs, ql);  // list units = final units
      float* gs = glist_s(qrow);
      int* gi = glist_i(qrow);
      float cs = -CUDART_INF_F;
      int ci = 0x7fffffff;
      if (lane < n) {
        const uint2 c = __ldcg(gbuf(qrow) + lane);
        cs = __uint_as_float(c.x) * unit;
        ci = (int)c.y;
      }
      for (int j = lane; j < k; j += 32) {
        Ws[j] = __ldcg(gs + j);
        Wi[j] = __ldcg(gi + j);
      }
      __syncwarp();
      const float nt = merge_sorted_candidates(Ws, Wi, Bs, cs, ci, n, k, lane);
      for (int j = lane; j < k; j += 32) {
        __stcg(gs + j, Ws[j]);
        __stcg(gi + j, Wi[j]);
      }
      __syncwarp();
      if (p.debug_ts && lane == 0) atomicAdd(p.debug_ts + (size_t)blockIdx.x * 8 + 7, 1ull << 32);  // flushes
      return cosine ? nt : nt * __shfl_sync(0xffffffffu, inv_qs, ql);  // exact: qs is a power of two
    };
    // compare 32 scores of this thread's query (columns c0..c0+31 of tile t) with its threshold; rare path: append
    auto filter32 = [&](const uint32_t (&r)[32], int t, int c0, float ndmin, float ndmax) {
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      float gm[4];  // maxima of the four groups of 8 consecutive columns
#pragma unroll
      for (int g = 0; g < 4; ++g)
        gm[g] = fmaxf(fmaxf(fmaxf(v[8 * g], v[8 * g + 1]), fmaxf(v[8 * g + 2], v[8 * g + 3])),
                      fmaxf(fmaxf(v[8 * g + 4], v[8 * g + 5]), fmaxf(v[8 * g + 6], v[8 * g + 7])));
      const float m = fmaxf(fmaxf(gm[0], gm[1]), fmaxf(gm[2], gm[3]));
      if (!__any_sync(0xffffffffu, m > thr_raw)) return;
      // rare path: only the groups in which some lane passes are parked in shared memory (8 stores per group instead of
      // 32 per chunk: the ring fills and the MMAs' operand reads already use the shared-memory bandwidth)
      const int row0 = t * RS_TILE_ROWS + c0;
      uint32_t pm = 0;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (__any_sync(0xffffffffu, gm[g] > thr_raw)) {
#pragma unroll
          for (int j = 8 * g; j < 8 * g + 8; ++j) {
            pm |= (v[j] > thr_raw) ? (1u << j) : 0u;
            dump[j * 32 + lane] = v[j];
          }
        }
      }
      if (row0 + 32 > p.n_rows) {  // ragged last tile: mask padded rows
        const int valid = p.n_rows - row0;
        pm = valid <= 0 ? 0u : (pm & (0xffffffffu >> (32 - valid)));
      }
      __syncwarp();
      while (__any_sync(0xffffffffu, pm != 0u)) {
        if (pm != 0u) {
          const int j = __ffs(pm) - 1;
          pm &= pm - 1;
          const float s = dump[j * 32 + lane];
          const int idx = row0 + j;
          float x = s;  // dot: compare and buffer raw units (exact power-of-two scale)
          if (cosine) {
            const float nd = p.dnorm[idx];
            const float denom = __fadd_rn(__fmul_rn(nq, nd), 1e-8f);  // retriever.py:605,607
            x = __fdiv_rn(s * qs, denom);
            x = fminf(fmaxf(x, -1.0f), 1.0f);                         // retriever.py:609
          }
          if (x > thr) {
            __stcg(gbuf(row) + cnt, make_uint2(__float_as_uint(x), (unsigned int)idx));
            ++cnt;
            if (p.debug_ts) atomicAdd(p.debug_ts + (size_t)blockIdx.x * 8 + 7, 1ull);  // appends (low word)
            if (shared_bound && count_bound) {  // publish this CTA's bound_rank-th best score of the query
              float pub = -CUDART_INF_F;
              if (x > best1) {
                best2 = best1;
                best1 = x;
                pub = bound_rank == 1 ? best1 : best2;
              } else if (x > best2) {
                best2 = x;
                if (bound_rank == 2) pub = best2;
              }
              if (pub > -CUDART_INF_F) __stcg(p.g_best + (size_t)pair * p.bpad + row, pub);
            }
          }
        }
        uint32_t fullm = __ballot_sync(0xffffffffu, cnt == kWideBufC);  // (rare: flushes normally happen after the release)
        while (fullm) {
          const int ql = __ffs(fullm) - 1;
          fullm &= fullm - 1;
          const float nt = flush_query(ql, kWideBufC);
          if (lane == ql) {
            thr = fmaxf(nt, gbound);
            cnt = 0;
            thr_raw = cosine ? cosine_raw_bound(thr, nq, inv_qs, ndmin, ndmax) : thr;
          }
        }
      }
      __syncwarp();
    };

    if (worker) {
      int b = 0;
      uint32_t acc_ph = 0;  // bit b = phase of accumulator b
      const bool prof = p.debug_ts != nullptr && threadIdx.x == kFirstEpiWarp * 32;
      long long t_wait = 0, t_hold = 0, t_merge = 0, tc0 = 0, tc1 = 0, tc2 = 0;
      const uint32_t tempty_leader = mapa_u32(tempty(0), 0);  // tempty(b) = tempty(0) + 8 b in the leader's window
      for (int u = pair; u < n_super; u += n_pairs) {
        const int nsub = min(2, p.n_tiles - 2 * u);
        if (prof) tc0 = clock64();
        mbar_wait_guarded<false>(tfull(b), (acc_ph >> b) & 1u, p.err_flag);
        tc_fence_after();
        if (prof) tc1 = clock64();
        if (shared_bound && u == pair) {
          // ---- sample pass over the FIRST accumulator: before a query has k candidates its own k-th best filters
          // nothing, and every document of the first supertile would be appended and merged (measured: 194 of a
          // query's 265 appends per pair, the epilogue far behind the MMAs).  So the first supertile is scanned twice
          // while its accumulator is held: pass A only tracks the query's two best scores in registers and publishes
          // the bound; once every pair has published (all are resident and publish within their first supertile),
          // the filter below runs against a bound that already rejects > 99 % of the documents.
          if (warp_live) {
            for (int sub = 0; sub < nsub; ++sub) {
              const uint32_t tacc = lane_base + (uint32_t)(b * 256 + sub * 128);
              const int trow0 = (2 * u + sub) * RS_TILE_ROWS;
#pragma unroll 1
              for (int c0 = 0; c0 < RS_TILE_ROWS; c0 += 32) {
                uint32_t r0[32];
                tmem_ld_x32(tacc + c0, r0);
                tmem_wait_ld_regs(r0);
                const int valid = p.n_rows - (trow0 + c0);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  float x = __uint_as_float(r0[j]);
                  if (cosine) {
                    const float nd = p.dnorm[min(trow0 + c0 + j, p.n_rows - 1)];
                    const float denom = __fadd_rn(__fmul_rn(nq, nd), 1e-8f);  // retriever.py:605,607
                    x = fminf(fmaxf(__fdiv_rn(x * qs, denom), -1.0f), 1.0f);   // retriever.py:609
                  }
                  if (j >= valid) x = -CUDART_INF_F;  // padded rows of a ragged last tile
                  best2 = fmaxf(best2, fminf(best1, x));
                  best1 = fmaxf(best1, x);
                }
              }
            }
          }
          if (live) {
            __stcg(p.g_best + (size_t)pair * p.bpad + row, bound_rank == 1 ? best1 : best2);  // -inf: nothing to offer
            unsigned int polls = 0;
            unsigned long long t0 = 0;
            while (true) {
              bool all = true;
#pragma unroll 8
              for (int c = 0; c < n_lists; ++c) {
                const float v = __ldcg(p.g_best + (size_t)c * p.bpad + row);
                all = all && (v == v);
              }
              if (all) break;
              __nanosleep(100);
              if ((++polls & 0xffu) == 0u) {
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (t0 == 0) t0 = now;
                if (now - t0 > 2000000000ull) break;  // 2 s: a pair of this grid is not running (device shared with another
                                                      // job); go on without the bound - the refresh below then sees an
                                                      // unpublished entry and keeps the query's own threshold: still exact
              }
            }
          }
          count_bound = false;
        } else {
          count_bound = true;
        }
        if (shared_bound && live && supers_done >= next_refresh - (u == pair ? 1 : 0)) {
          // bound refresh: after the sample, then after 1, 2, 3, 4 supertiles, then at 1.5x intervals, then every 32
          if (u != pair) next_refresh = next_refresh < 4 ? next_refresh + 1 : (next_refresh < 64 ? next_refresh + next_refresh / 2 : next_refresh + 32);
          float mn = CUDART_INF_F;
#pragma unroll 16
          for (int c = 0; c < n_lists; ++c) {
            const float v = __ldcg(p.g_best + (size_t)c * p.bpad + row);
            mn = fminf(mn, v == v ? v : -CUDART_INF_F);  // NaN = that pair has not published yet
          }
          if (mn > -CUDART_INF_F && mn < CUDART_INF_F) {
            gbound = __uint_as_float(__float_as_uint(mn) + (mn > 0.f ? -1 : (mn < 0.f ? 1 : 0)));  // next float below
            if (mn == 0.f) gbound = -1.401298464e-45f;
            if (gbound > thr) {
              thr = gbound;
              thr_raw = cosine ? CUDART_INF_F : thr;  // cosine: recomputed per tile just below
            }
          }
        }
        ++supers_done;
        if (warp_live) {
          for (int sub = 0; sub < nsub; ++sub) {
            const int t = 2 * u + sub;
            float ndmin = 1.f, ndmax = 1.f;
            if (cosine) {
              const float2 mm = p.tile_norm_minmax[t];
              ndmin = mm.x;
              ndmax = mm.y;
              thr_raw = live ? cosine_raw_bound(thr, nq, inv_qs, ndmin, ndmax) : CUDART_INF_F;
            }
            const uint32_t tacc = lane_base + (uint32_t)(b * 256 + sub * 128);
            uint32_t ra[32], rb[32];
            tmem_ld_x32(tacc, ra);
#pragma unroll 1
            for (int c0 = 0; c0 < RS_TILE_ROWS; c0 += 64) {
              tmem_wait_ld_regs(ra);
              tmem_ld_x32(tacc + c0 + 32, rb);
              filter32(ra, t, c0, ndmin, ndmax);
              tmem_wait_ld_regs(rb);
              if (c0 + 64 < RS_TILE_ROWS) tmem_ld_x32(tacc + c0 + 64, ra);
              filter32(rb, t, c0 + 32, ndmin, ndmax);
            }
          }
        }
        // every TMEM read of this supertile is complete: hand the accumulator back to the leader's issuer, THEN merge
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(tempty_leader + 8u * (uint32_t)b);
        acc_ph ^= 1u << b;
        b ^= 1;
        if (prof) tc2 = clock64();
        uint32_t due = __ballot_sync(0xffffffffu, cnt >= flush_at);
        while (due) {
          const int ql = __ffs(due) - 1;
          due &= due - 1;
          const int n = __shfl_sync(0xffffffffu, cnt, ql);
          const float nt = flush_query(ql, n);
          if (lane == ql) {
            thr = fmaxf(nt, gbound);
            cnt = 0;
            thr_raw = cosine ? CUDART_INF_F : thr;  // cosine: recomputed per tile
          }
        }
        if (prof) {
          t_wait += tc1 - tc0;
          t_hold += tc2 - tc1;
          t_merge += clock64() - tc2;
        }
      }
      if (prof) {
        p.debug_ts[(size_t)blockIdx.x * 8 + 1] = (unsigned long long)t_wait;
        p.debug_ts[(size_t)blockIdx.x * 8 + 2] = (unsigned long long)t_hold;
        p.debug_ts[(size_t)blockIdx.x * 8 + 3] = (unsigned long long)t_merge;
      }
      // drain the partially filled buffers; the lists already sit in cand_* in final units
      uint32_t rem = __ballot_sync(0xffffffffu, cnt > 0);
      while (rem) {
        const int ql = __ffs(rem) - 1;
        rem &= rem - 1;
        const int n = __shfl_sync(0xffffffffu, cnt, ql);
        flush_query(ql, n);
      }
    }
  }

  if (p.tail_mode != 0) __threadfence();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // no signal crosses the pair after this point; both CTAs are done with TMEM
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kTmemCols);
  }
  fused_tail(p, smem, worker);
}

// ==========================================================================================
// CTA-pair variant of the DOC-MAJOR kernel (two accumulators: fp32 storage and/or two query terms; B <= 32, k <= 16,
// large corpora).  The doc-major kernel keeps the whole Q image resident; with two terms that is 128 KiB at B = 32
// and leaves the document ring 64 KiB, which is what holds fp32 corpora at 0.86 of the HBM roofline (B = 16, with a
// 64 KiB image, streams at 1.0).  Here two SMs run one tcgen05.mma.cta_group::2 chain: M = 256 = the two CTAs'
// document tiles (each CTA streams its OWN tile, both planes), N = 32 = the queries, of which each CTA holds HALF
// (16 rows of the N operand: the tensor cores fetch the other half through the pair link).  The resident image
// shrinks to 64 KiB per CTA and the ring doubles.  Each CTA's TMEM receives its 128 documents x 32 queries
// (x 2 accumulators, 8 buffers), so the epilogue, the per-CTA candidate lists and the tail are the doc-major
// kernel's (dm_epilogue<.., PAIR>).  Barriers and the relay are the wide pair kernel's.  Query prep is the
// stand-alone kernel (each CTA bulk-copies its 16 rows of every k-chunk block).
// ==========================================================================================
struct DmPairSmem {
  uint32_t q_bytes, stage_off, stage_bytes, bar_off, dump_off, total;
};
__host__ __device__ inline DmPairSmem make_dmpair_smem(int kc, int terms, int planes, int stages, int cps, int bpad = 32) {
  DmPairSmem m;
  m.q_bytes = (uint32_t)(terms * kc) * (uint32_t)(bpad / 2) * RS_KCHUNK_BYTES;  // per (term, k-chunk): this CTA's bpad/2 query rows x 128 B
  m.stage_off = m.q_bytes;
  m.stage_bytes = (uint32_t)(planes * cps) * RS_BLOCK_BYTES;
  m.bar_off = m.stage_off + (uint32_t)stages * m.stage_bytes;
  m.dump_off = m.bar_off + 2048u;  // [bar_off, +1 KiB) barriers + TMEM address, then qnorm[128], qscale[128]
  m.total = m.dump_off + 4u * 32u * 32u * 4u;
  return m;
}

template <int NACC, int NH>
__global__ void __launch_bounds__(kThreads, 1) fused_search_dmpair_kernel(const __grid_constant__ SearchParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = (int)blockIdx.x >> 1;
  const int n_pairs = (int)gridDim.x >> 1;
  const int cps = p.cps, NB = p.nbuf;
  const DmPairSmem map = make_dmpair_smem(p.kc, p.terms, p.planes, p.stages, cps, 32 * NH);
  constexpr uint32_t qblk = (uint32_t)(16 * NH) * RS_KCHUNK_BYTES;  // this CTA's rows of one (term, k-chunk) block of the Q image
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t q_smem = smem_base, stage_smem = smem_base + map.stage_off, bar_base = smem_base + map.bar_off;
  auto full = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto empty = [&](int s) { return bar_base + 8u * (uint32_t)(8 + s); };
  auto tfull = [&](int a) { return bar_base + 8u * (uint32_t)(16 + a); };
  auto tempty = [&](int a) { return bar_base + 8u * (uint32_t)(24 + a); };
  const uint32_t qfull = bar_base + 8u * 32u;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + map.bar_off + 1016);
  float* s_qnorm = reinterpret_cast<float*>(smem + map.bar_off + 1024);
  float* s_qscale = s_qnorm + 128;
  const int n_super = (p.n_tiles + 1) / 2;
  const bool worker = pair < n_super && (int)blockIdx.x < p.n_work;
  const size_t tile_bytes = (size_t)p.planes * p.kc * RS_BLOCK_BYTES;
  const uint32_t tmem_cols = (uint32_t)(NB * NACC * NH) * 32u;

  if (threadIdx.x == 0) {
    stamp(p.debug_ts, 0);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full(s), rank == 0 ? 2u : 1u);  // leader: its own copies + the peer's relay
      mbar_init(empty(s), 1);
    }
    for (int a = 0; a < 8; ++a) {
      mbar_init(tfull(a), 1);
      mbar_init(tempty(a), 2 * kNumEpiWarps);  // leader only: the epilogue warps of both CTAs
    }
    mbar_init(qfull, rank == 0 ? 2u : 1u);
    fence_mbar_init();
    fence_proxy_async();
  }
  if (warp == kMmaWarp) {
    tmem_alloc_pair(smem_u32((const void*)tmem_slot), tmem_cols);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kProducerWarp) {
    if (worker) {
      const uint64_t pol_stream = l2_policy_evict_first();
      const uint64_t pol_keep = l2_policy_evict_last();
      if (lane == 0) {  // this CTA's 16 (32) rows of every (term, k-chunk) block of the Q image
        mbar_arrive_expect_tx(qfull, map.q_bytes);
        for (int blk = 0; blk < p.terms * p.kc; ++blk)
          bulk_g2s(q_smem + (uint32_t)blk * qblk, p.qimg + (size_t)blk * 2u * qblk + (size_t)rank * qblk, qblk, qfull, pol_keep);
      }
      int s = 0;
      uint32_t ph = 0;
      const uint32_t plane_bytes = (uint32_t)cps * RS_BLOCK_BYTES;
      for (int u = pair; u < n_super; u += n_pairs) {
        const int tile = (2 * u + (int)rank < p.n_tiles) ? 2 * u + (int)rank : 2 * u;  // odd tail: masked duplicate
        const uint8_t* src = p.corpus + (size_t)tile * tile_bytes;
        for (int kc = 0; kc < p.kc; kc += cps) {
          mbar_wait_guarded<false>(empty(s), ph ^ 1u, p.err_flag);
          if (lane == 0) {
            mbar_arrive_expect_tx(full(s), map.stage_bytes);
            for (int pl = 0; pl < p.planes; ++pl)
              bulk_g2s(stage_smem + (uint32_t)s * map.stage_bytes + (uint32_t)pl * plane_bytes,
                       src + (size_t)(pl * p.kc + kc) * RS_BLOCK_BYTES, plane_bytes, full(s), pol_stream);
          }
          if (++s == p.stages) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
      for (int i = 0; i < p.stages; ++i) {  // drain the last multicast commits
        mbar_wait_guarded<false>(empty(s), ph ^ 1u, p.err_flag);
        if (++s == p.stages) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
  } else if (warp == kMmaWarp) {
    if (worker && rank == 0) {
      // ===================== MMA issuer (leader): M = 256 (two document tiles) x N = 32 (queries) x K = 16 =====================
      const uint32_t idesc = umma_idesc_f16(p.dtype == RS_DTYPE_BF16, 256, 32 * NH);
      const bool skip_mma = (p.debug_flags & 1) != 0;
      mbar_wait_guarded<true>(qfull, 0, p.err_flag);
      tc_fence_after();
      int s = 0, acc = 0;
      uint32_t ph = 0, acc_ph = 0;
      const uint32_t plane_bytes = (uint32_t)cps * RS_BLOCK_BYTES;
      for (int u = pair; u < n_super; u += n_pairs) {
        mbar_wait_guarded<true>(tempty(acc), acc_ph ^ 1u, p.err_flag);
        tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)(acc * NACC * NH) * 32u, d1 = d0 + (uint32_t)NH * 32u;
        for (int kc = 0; kc < p.kc; kc += cps) {
          mbar_wait_guarded<true>(full(s), ph, p.err_flag);
          tc_fence_after();
          if (lane == 0) {
            if (!skip_mma) {
              const uint32_t st = stage_smem + (uint32_t)s * map.stage_bytes;
              for (int c = 0; c < cps; ++c) {
                const bool fresh = (kc + c) == 0;
                const uint64_t a_hi = umma_smem_desc(st + (uint32_t)c * RS_BLOCK_BYTES);
                const uint64_t b_hi = umma_smem_desc(q_smem + (uint32_t)(kc + c) * qblk);
#pragma unroll
                for (int ks = 0; ks < RS_KCHUNK_ELEMS / RS_UMMA_K; ++ks)
                  umma_f16_pair(d0, a_hi + 2u * ks, b_hi + 2u * ks, idesc, (fresh && ks == 0) ? 0u : 1u);
                if (NACC == 2) {
                  const uint64_t b_lo = umma_smem_desc(q_smem + (uint32_t)(p.kc + kc + c) * qblk);
#pragma unroll
                  for (int ks = 0; ks < RS_KCHUNK_ELEMS / RS_UMMA_K; ++ks)
                    umma_f16_pair(d1, a_hi + 2u * ks, b_lo + 2u * ks, idesc, (fresh && ks == 0) ? 0u : 1u);
                  if (p.planes == 2) {  // lo plane of an fp32 corpus: q_hi . d_lo joins the low-order accumulator
                    const uint64_t a_lo = umma_smem_desc(st + plane_bytes + (uint32_t)c * RS_BLOCK_BYTES);
#pragma unroll
                    for (int ks = 0; ks < RS_KCHUNK_ELEMS / RS_UMMA_K; ++ks)
                      umma_f16_pair(d1, a_lo + 2u * ks, b_hi + 2u * ks, idesc, 1u);
                  }
                }
              }
            }
            umma_commit_pair(empty(s));
          }
          __syncwarp();
          if (++s == p.stages) {
            s = 0;
            ph ^= 1u;
          }
        }
        if (lane == 0) umma_commit_pair(tfull(acc));
        __syncwarp();
        if (++acc == NB) {
          acc = 0;
          acc_ph ^= 1u;
        }
      }
    } else if (worker) {
      // ===================== relay (peer) =====================
      const uint32_t full_leader = mapa_u32(full(0), 0);
      mbar_wait_guarded<false>(qfull, 0, p.err_flag);
      if (lane == 0) mbar_arrive_remote(mapa_u32(qfull, 0));
      __syncwarp();
      int s = 0;
      uint32_t ph = 0;
      for (int u = pair; u < n_super; u += n_pairs) {
        for (int kc = 0; kc < p.kc; kc += cps) {
          mbar_wait_guarded<false>(full(s), ph, p.err_flag);
          if (lane == 0) mbar_arrive_remote(full_leader + 8u * (uint32_t)s);
          __syncwarp();
          if (++s == p.stages) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    }
  } else {
    const int quarter = warp & 3;
    const DmBars bars{tfull(0), tempty(0), 0u, 0u, qfull, nullptr, pair, n_pairs, (int)rank, mapa_u32(tempty(0), 0)};
    dm_epilogue<8, NACC, true, NH>(p, bars, tmem_base, quarter, lane, reinterpret_cast<float*>(smem + map.dump_off) + quarter * 1024,
                               s_qnorm, s_qscale, false, worker, reinterpret_cast<float*>(smem),
                               reinterpret_cast<int*>(smem + 4 * 32 * NH * kRegK * sizeof(float)));
  }

  if (p.tail_mode != 0) __threadfence();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, tmem_cols);
  }
  dm_publish_cta_lists(p, smem, worker);
  fused_tail(p, smem, worker);
}

// smem bytes the tail needs at offset 0 (result row + staged lists of one query)
size_t fused_tail_smem_bytes(int n_lists, int k, int world) {
  return merge_res_bytes(k) + std::max((size_t)n_lists * k * 8, (size_t)world * k * 12);
}
int fused_tail_smem_capacity(int bpad, int kc, int terms, int stages, int cps, int k, int q_resident, int qstages) {
  return (int)make_smem_map(bpad, kc, terms, stages, cps, k, k <= kRegK, q_resident != 0, qstages, kBufC, false)
      .bar_off;
}

// Plain merge: one CTA per query, lists -> out[q][0..k).
template <typename IdxT, bool STAGED>
__global__ void __launch_bounds__(kMergeThreads) merge_topk_kernel(const float* __restrict__ cand_score,
                                                                   const IdxT* __restrict__ cand_idx, int n_lists,
                                                                   int n_queries, int kc, int k, long long idx_offset,
                                                                   long long score_list_stride,
                                                                   long long idx_list_stride, long long query_stride,
                                                                   float* __restrict__ out_score,
                                                                   long long* __restrict__ out_idx) {
  extern __shared__ __align__(16) uint8_t merge_smem[];
  const int q = blockIdx.x;
  const int lane = threadIdx.x & 31;
  long long* res_i = reinterpret_cast<long long*>(merge_smem);
  float* res_s = reinterpret_cast<float*>(res_i + k);
  uint8_t* stage = merge_smem + merge_res_bytes(k);
  IdxT* s_idx = reinterpret_cast<IdxT*>(stage);
  float* s_score = reinterpret_cast<float*>(stage + (size_t)n_lists * kc * sizeof(IdxT));
  const size_t qoff = (size_t)q * query_stride;
  if (STAGED)
    stage_lists<IdxT>(cand_score, cand_idx, n_lists, kc, score_list_stride, idx_list_stride, qoff, s_idx, s_score,
                      kMergeThreads);
  if (threadIdx.x >= 32) return;  // the selection rounds are a single-warp job
  auto load = [&](int l, int j, float& sc, long long& id) {
    if (STAGED) {
      sc = s_score[l * kc + j];
      id = (long long)s_idx[l * kc + j];
    } else {
      id = (long long)__ldcg(cand_idx + (size_t)l * idx_list_stride + qoff + j);
      sc = __ldcg(cand_score + (size_t)l * score_list_stride + qoff + j);
    }
  };
  select_rounds(load, n_lists, kc, k, idx_offset, res_s, res_i, lane);
  for (int j = lane; j < k; j += 32) {
    out_score[(size_t)q * k + j] = res_s[j];
    out_idx[(size_t)q * k + j] = res_i[j];
  }
}

// Stand-alone local merge (RS_FUSED_TAIL=0 fallback): one CTA per query over the per-CTA lists.
__global__ void __launch_bounds__(kMergeThreads) merge_local_kernel(const float* __restrict__ cand_score,
                                                                    const int* __restrict__ cand_idx, int n_lists, int k,
                                                                    long long idx_offset, float* __restrict__ out_score,
                                                                    long long* __restrict__ out_idx) {
  extern __shared__ __align__(16) uint8_t merge_smem[];
  const int q = blockIdx.x;
  const MergeSmem m = merge_smem_at(merge_smem, k);
  merge_query_local(cand_score, cand_idx, n_lists, k, idx_offset, q, m, kMergeThreads);
  for (int j = threadIdx.x; j < k; j += kMergeThreads) {
    out_score[(size_t)q * k + j] = m.res_s[j];
    out_idx[(size_t)q * k + j] = m.res_i[j];
  }
}

// Multi-GPU: local merge + candidate exchange + final merge in ONE kernel (one CTA per query).
//   1. merge this GPU's per-CTA lists of query q into its local top-k (global indices)
//   2. store that row into every GPU's exchange buffer through peer pointers (NVLink) and raise
//      this rank's per-query flag there with the search epoch
//   3. wait until every rank's flag for query q shows this epoch (written by the PEER GPUs)
//   4. merge the world lists of query q from the local exchange buffer into the final row
// A CTA only ever waits for the CTAs of the same query on other GPUs, never for a CTA of its own
// grid, so no co-residency is required.  Buffers are double-buffered by epoch parity.
__global__ void __launch_bounds__(kMergeThreads) exchange_merge_kernel(const float* __restrict__ cand_score,
                                                                       const int* __restrict__ cand_idx, int n_lists,
                                                                       int k, long long idx_offset,
                                                                       float* __restrict__ out_score,
                                                                       long long* __restrict__ out_idx,
                                                                       const __grid_constant__ ExchangeArgs x, unsigned int* err_flag) {
  extern __shared__ __align__(16) uint8_t merge_smem[];
  const int q = blockIdx.x;  // query inside this pass
  const MergeSmem m = merge_smem_at(merge_smem, k);
  merge_query_local(cand_score, cand_idx, n_lists, k, idx_offset, q, m, kMergeThreads);
  exchange_query(x, q, k, m, kMergeThreads, err_flag);
  for (int j = threadIdx.x; j < k; j += kMergeThreads) {
    out_score[(size_t)q * k + j] = m.res_s[j];
    out_idx[(size_t)q * k + j] = m.res_i[j];
  }
  if (threadIdx.x == 0 && atomicAdd(x.epoch + 1, 1u) == gridDim.x - 1u) {  // last CTA of the pass
    atomicExch(x.epoch + 1, 0u);
    if (x.last_pass) bump_epoch(x.epoch);
  }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
cudaError_t launch_query_prep(const float* q_dev, int B, int D, int bpad, int dtype, int terms, int corpus_scale_exp,
                              uint8_t* qimg, float* qnorm, float* qscale, cudaStream_t stream) {
  query_prep_kernel<<<bpad, 256, 0, stream>>>(q_dev, B, D, bpad, dtype, terms, corpus_scale_exp, qimg, qnorm,
                                              qscale);
  return cudaGetLastError();
}

template <int NACC, bool REGLIST, bool QRES, bool DOCM = false>
static cudaError_t launch_fused_t(const SearchParams& p, int grid, cudaStream_t stream) {
  const SmemMap map = make_smem_map(p.bpad, p.kc, p.terms, p.stages, p.cps, p.k, REGLIST, QRES, p.qstages, p.bufc,
                                    p.lists_global != 0, DOCM);
  auto kern = fused_search_kernel<NACC, REGLIST, QRES, DOCM>;
  const int bytes = (int)map.total + 1024;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = (size_t)bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  if (p.pdl) {  // programmatic dependent launch: this grid may start while the previous kernel of the stream drains
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  return cudaLaunchKernelEx(&cfg, kern, p);
}

int fused_smem_bytes(int bpad, int kc, int terms, int stages, int cps, int k, int q_resident, int qstages, int bufc,
                     int lists_global, int docm) {
  return (int)make_smem_map(bpad, kc, terms, stages, cps, k, k <= kRegK, q_resident != 0, qstages, bufc, lists_global != 0,
                            docm != 0)
             .total + 1024;
}
int fused_reg_k() { return kRegK; }

template <int NACC, bool REGLIST>
static cudaError_t launch_fused_q(const SearchParams& p, int grid, cudaStream_t stream) {
  return p.q_resident ? launch_fused_t<NACC, REGLIST, true>(p, grid, stream)
                      : launch_fused_t<NACC, REGLIST, false>(p, grid, stream);
}

int fused_wide_bufc() { return kWideBufC; }

int fused_pair_smem_bytes(int k) { return (int)make_pair_smem(k).total + 1024; }
int fused_pair_tail_capacity(int k) { return (int)make_pair_smem(k).bar_off; }

static void pair_launch_config(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attr, int grid, int bytes, cudaStream_t stream) {
  memset(cfg, 0, sizeof(*cfg));
  cfg->gridDim = dim3((unsigned)grid);
  cfg->blockDim = dim3(kThreads);
  cfg->dynamicSmemBytes = (size_t)bytes;
  cfg->stream = stream;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg->attrs = attr;
  cfg->numAttrs = 1;
}

int fused_pair_max_clusters(int k) {
  const int bytes = fused_pair_smem_bytes(k);
  if (cudaFuncSetAttribute(fused_search_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  pair_launch_config(&cfg, attr, sms & ~1, bytes, nullptr);
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, fused_search_pair_kernel, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

static cudaError_t launch_fused_pair(const SearchParams& p, int grid, cudaStream_t stream) {
  if (p.bpad != 256 || p.terms != 1 || p.planes != 1 || p.n_queries > 256 || p.k > kMaxListK || p.cbuf == nullptr ||
      (grid & 1) || (p.n_work & 1) || p.n_lists * 2 != p.n_work || (p.cps != 1 && p.cps != 2) || p.kc % p.cps != 0 || (p.g_best && p.bound_rank != 1 && p.bound_rank != 2))
    return cudaErrorInvalidValue;
  const int bytes = fused_pair_smem_bytes(p.k);
  cudaError_t e = cudaFuncSetAttribute(fused_search_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  pair_launch_config(&cfg, attr, grid, bytes, stream);
  return cudaLaunchKernelEx(&cfg, fused_search_pair_kernel, p);
}

int fused_dmpair_smem_bytes(int kc, int terms, int planes, int stages, int cps, int bpad) {
  return (int)make_dmpair_smem(kc, terms, planes, stages, cps, bpad).total + 1024;
}
int fused_dmpair_tail_capacity(int kc, int terms, int planes, int stages, int cps, int bpad) {
  return (int)make_dmpair_smem(kc, terms, planes, stages, cps, bpad).bar_off;
}
int fused_dmpair_max_clusters(int smem_bytes) {
  if (cudaFuncSetAttribute(fused_search_dmpair_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  pair_launch_config(&cfg, attr, sms & ~1, smem_bytes, nullptr);
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, fused_search_dmpair_kernel<2, 1>, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}
template <int NACC, int NH>
static cudaError_t launch_fused_dmpair_t(const SearchParams& p, int grid, cudaStream_t stream) {
  const int bytes = fused_dmpair_smem_bytes(p.kc, p.terms, p.planes, p.stages, p.cps, p.bpad);
  cudaError_t e = cudaFuncSetAttribute(fused_search_dmpair_kernel<NACC, NH>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  pair_launch_config(&cfg, attr, grid, bytes, stream);
  return cudaLaunchKernelEx(&cfg, fused_search_dmpair_kernel<NACC, NH>, p);
}
static cudaError_t launch_fused_dmpair(const SearchParams& p, int grid, cudaStream_t stream) {
  const int nacc = (p.terms == 2 || p.planes == 2) ? 2 : 1;
  const int nh = p.bpad / 32;
  if ((p.bpad != 32 && p.bpad != 64) || p.n_queries > p.bpad || p.k > kRegK || (grid & 1) || (p.n_work & 1) ||
      p.nbuf * nacc * nh * 32 > 512 || p.nbuf < 1 || p.nbuf > 8 || p.stages < 2 || p.stages > 8 || p.kc % p.cps != 0 || p.inline_prep ||
      !((nacc == 2 && nh == 1) || (nacc == 1 && nh == 2)))
    return cudaErrorInvalidValue;
  return nacc == 2 ? launch_fused_dmpair_t<2, 1>(p, grid, stream) : launch_fused_dmpair_t<1, 2>(p, grid, stream);
}

cudaError_t launch_fused_search(const SearchParams& p, int grid, cudaStream_t stream) {
  if (p.dmpair) return launch_fused_dmpair(p, grid, stream);
  if (p.wide) return launch_fused_pair(p, grid, stream);
  const bool reglist = p.k <= kRegK;
  const int nacc = (p.terms == 2 || p.planes == 2) ? 2 : 1;
  if (p.docm) {
    if (!reglist || p.q_resident != 1 || p.bpad > 32 || p.sub != 1 || p.nbuf > 8) return cudaErrorInvalidValue;
    return nacc == 1 ? launch_fused_t<1, true, true, true>(p, grid, stream) : launch_fused_t<2, true, true, true>(p, grid, stream);
  }
  if (nacc == 1) return reglist ? launch_fused_q<1, true>(p, grid, stream) : launch_fused_q<1, false>(p, grid, stream);
  return reglist ? launch_fused_q<2, true>(p, grid, stream) : launch_fused_q<2, false>(p, grid, stream);
}

template <typename IdxT>
static cudaError_t launch_merge_t(const float* cand_score, const IdxT* cand_idx, int n_lists, int n_queries, int kc,
                                  int k, long long idx_offset, long long score_list_stride,
                                  long long idx_list_stride, long long query_stride, float* out_score,
                                  long long* out_idx, cudaStream_t stream) {
  if (n_lists > 32 * kMergeMaxOwn) return cudaErrorInvalidValue;
  const size_t res_bytes = merge_res_bytes(k);
  const size_t stage_bytes = (size_t)n_lists * kc * (sizeof(IdxT) + sizeof(float));
  if (res_bytes + stage_bytes <= 200 * 1024) {
    auto kern = merge_topk_kernel<IdxT, true>;
    const size_t bytes = res_bytes + stage_bytes;
    if (bytes > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
      if (e != cudaSuccess) return e;
    }
    kern<<<n_queries, kMergeThreads, bytes, stream>>>(cand_score, cand_idx, n_lists, n_queries, kc, k, idx_offset,
                                                      score_list_stride, idx_list_stride, query_stride, out_score,
                                                      out_idx);
  } else {
    merge_topk_kernel<IdxT, false><<<n_queries, kMergeThreads, res_bytes, stream>>>(
        cand_score, cand_idx, n_lists, n_queries, kc, k, idx_offset, score_list_stride, idx_list_stride, query_stride,
        out_score, out_idx);
  }
  return cudaGetLastError();
}

cudaError_t launch_merge_i32(const float* cand_score, const int* cand_idx, int n_lists, int n_queries, int kc, int k,
                             long long idx_offset, float* out_score, long long* out_idx, cudaStream_t stream) {
  // per-CTA candidates are stored [query][cta][k]: one query's lists are contiguous (coalesced staging)
  if (kc != k || n_lists > 32 * kMergeMaxOwn) return cudaErrorInvalidValue;
  const size_t bytes = merge_res_bytes(k) + (size_t)n_lists * k * 8;
  if (bytes > 200 * 1024) return cudaErrorInvalidValue;
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(merge_local_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
  }
  merge_local_kernel<<<n_queries, kMergeThreads, bytes, stream>>>(cand_score, cand_idx, n_lists, k, idx_offset, out_score,
                                                                  out_idx);
  return cudaGetLastError();
}
cudaError_t launch_merge_i64(const float* cand_score, const long long* cand_idx, int n_lists, int n_queries, int kc,
                             int k, long long score_list_stride, long long idx_list_stride, float* out_score,
                             long long* out_idx, cudaStream_t stream) {
  return launch_merge_t<long long>(cand_score, cand_idx, n_lists, n_queries, kc, k, 0, score_list_stride,
                                   idx_list_stride, (long long)kc, out_score, out_idx, stream);
}

cudaError_t launch_exchange_merge(const float* cand_score, const int* cand_idx, int n_lists, int n_queries, int k,
                                  long long idx_offset, float* out_score, long long* out_idx, const ExchangeArgs& x,
                                  unsigned int* err_flag, cudaStream_t stream) {
  if (n_lists > 32 * kMergeMaxOwn || x.world > kMaxPeers) return cudaErrorInvalidValue;
  const size_t res_bytes = merge_res_bytes(k);
  const size_t stage_bytes = std::max((size_t)n_lists * k * 8, (size_t)x.world * k * 12);
  const size_t bytes = res_bytes + stage_bytes;
  if (bytes > 200 * 1024) return cudaErrorInvalidValue;
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(exchange_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
  }
  exchange_merge_kernel<<<n_queries, kMergeThreads, bytes, stream>>>(cand_score, cand_idx, n_lists, k, idx_offset,
                                                                     out_score, out_idx, x, err_flag);
  return cudaGetLastError();
}

}  // namespace rs
