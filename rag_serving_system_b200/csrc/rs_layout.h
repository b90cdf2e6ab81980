// HBM / shared-memory layout shared by every kernel and by the host code.
//
// The corpus is NOT kept row-major.  It is stored in HBM as a sequence of document tiles, each
// tile being the exact shared-memory image a tcgen05 K-major SWIZZLE_128B operand needs:
//
//   tile t (128 documents)  =  planes x KC blocks, block (p, kc) = 128 rows x 128 bytes
//   row r of a block sits at   (r/8)*1024 + (r%8)*128
//   16-byte chunk c of a row sits at chunk position  c ^ (r%8)      (Swizzle<3,4,3>)
//
// so one contiguous 16 KiB bulk copy (cp.async.bulk, the TMA engine) lands an MMA-ready
// B-operand block in shared memory, and the whole scan is a perfectly sequential HBM stream.
// "planes" is 1 for bf16/fp16 storage and 2 for fp32 storage (fp16 hi plane + fp16 lo plane
// scaled by 2^11: 4 bytes per element, the same bytes as fp32, but tensor-core consumable).
//
// The query operand ("Q image") uses the same block format with Bpad rows per block.
#pragma once
#include <stdint.h>

#define RS_TILE_ROWS 128          // documents per tile  (= MMA N)
#define RS_KCHUNK_ELEMS 64        // two-byte elements per swizzle row
#define RS_KCHUNK_BYTES 128
#define RS_BLOCK_BYTES (RS_TILE_ROWS * RS_KCHUNK_BYTES)   // 16 KiB
#define RS_UMMA_K 16              // K per tcgen05.mma for 16-bit operands
#define RS_FP16_SPLIT_SHIFT 11    // lo plane / lo query term scaled by 2^11

#define RS_DTYPE_BF16 0
#define RS_DTYPE_FP16 1
#define RS_DTYPE_FP32 2

#define RS_METRIC_DOT 0
#define RS_METRIC_COSINE 1

#if defined(__CUDACC__)
#define RS_HD __host__ __device__ __forceinline__
#else
#define RS_HD static inline
#endif

// byte offset of two-byte element (row rr in [0,rows), column cc in [0,64)) inside one block
RS_HD uint32_t rs_block_offset(uint32_t rr, uint32_t cc) {
  uint32_t chunk = cc >> 3, within = cc & 7u;
  return (rr >> 3) * 1024u + (rr & 7u) * 128u + ((chunk ^ (rr & 7u)) << 4) + within * 2u;
}

// byte offset of element [plane p, row r, column c] in the tiled corpus image
RS_HD uint64_t rs_corpus_offset(uint64_t r, uint32_t c, uint32_t p, uint32_t kc_count, uint32_t planes) {
  uint64_t t = r / RS_TILE_ROWS;
  uint32_t rr = (uint32_t)(r % RS_TILE_ROWS);
  uint64_t tile_bytes = (uint64_t)planes * kc_count * RS_BLOCK_BYTES;
  return t * tile_bytes + (uint64_t)(p * kc_count + (c >> 6)) * RS_BLOCK_BYTES + rs_block_offset(rr, c & 63u);
}
