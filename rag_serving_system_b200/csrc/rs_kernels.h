// Internal interface between the C-ABI host code (rs_api.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rs {

// Arguments of the in-kernel candidate exchange (multi-GPU, peer memory over NVLink).
// Every GPU owns an exchange buffer of 16-byte entries {int64 global index, float score, uint32 epoch}:
//   entries[parity][rank r][(q0 + q) * k + j]      parity = epoch & 1, one slot of slot_entries per rank
// A rank stores its merged row of query q into ITS slot on every GPU with one 16-byte store per entry; the
// epoch travels inside the entry, so there is no separate flag, no system-scope fence and no second round
// trip: a reader polls the entry itself until it shows the search's epoch.
constexpr int kMaxPeers = 8;
struct ExchangeArgs {
  int world;                            // GPUs taking part
  int rank;                             // this GPU
  int q0;                               // first query of this pass inside the search's batch
  int last_pass;                        // 1: the launch that completes the search bumps the epoch
  unsigned int* epoch;                  // device words of the LOCAL buffer: [0] search counter (never 0; all ranks run
                                        // in lockstep), [1] departures of the stand-alone exchange kernel
  uint4* dest[kMaxPeers];               // GPU g: this rank's slot in g's buffer, parity 0 (peer pointer; dest[rank] local)
  const uint4* local;                   // local buffer, slot of rank 0, parity 0
  long long slot_entries;               // entries between consecutive ranks' slots
  long long parity_entries;             // entries between the two parities
};

struct SearchParams {
  const uint8_t* corpus;             // tiled corpus image (rs_layout.h)
  const uint8_t* qimg;               // Q image: terms x KC blocks of bpad rows
  const float* qnorm;                // [bpad] L2 norm of each query (cosine)
  const float* qscale;               // [bpad] power-of-two factor: raw accumulator -> dot product
  const float* dnorm;                // [n_tiles*128] document L2 norms (cosine)
  const float2* tile_norm_minmax;    // [n_tiles] (min, max) document norm per tile (cosine)
  float* cand_score;                 // [n_queries][grid][k] per-CTA candidates, best first
  int* cand_idx;                     // [n_queries][grid][k] local row ids, -1 = padding
  int n_rows;                        // valid documents in this shard
  int n_tiles;                       // ceil(n_rows / 128)
  int n_queries;                     // B
  int kc;                            // D / 64
  int planes;                        // 1 (bf16/fp16 storage) or 2 (fp32 as split fp16)
  int terms;                         // query terms: 1 (rounded) or 2 (hi + lo)
  int bpad;                          // B rounded up to a multiple of 8
  int k;
  int metric;
  int dtype;                         // element type of the 16-bit planes: RS_DTYPE_BF16 / RS_DTYPE_FP16
  int stages;                        // depth of the bulk-copy ring
  int qstages;                       // streamed-Q mode: slots of the private Q ring (each cps k-chunks)
  int cps;                           // 64-element k-chunks per ring stage: 1 (16 KiB) or 2 (32 KiB)
  int q_resident;                    // 1: whole Q image resident in smem; 0: Q k-chunks stream through the ring
  int sub;                           // 128-document tiles per supertile (share one Q k-chunk load)
  int mma_m;                         // 128, or 64 when bpad <= 64 (half the tensor work / operand reads)
  int nbuf;                          // TMEM accumulator buffers (1 or 2); nbuf*sub*NACC*128 <= 512 columns
  int bufc;                          // k > 16: candidate-buffer entries per query before a flush (32, or 16 to save smem)
  int docm;                          // doc-major kernel variant: documents on the MMA's M, queries on its N (B <= 32, k <= 16, dot)
  int lists_global;                  // k > 16: sorted lists in global memory (cand_*), only the buffers in shared memory
  int debug_flags;                   // timing experiments only (RS_DEBUG_FLAGS): 1 = skip MMAs, 2 = skip epilogue, 4 = skip tail merge
  float lo_coef;                     // weight of the low-order accumulator (2^-11 for fp16 terms, 1 for bf16)
  // ---- single-launch search: query prep in the prologue, candidate merge (+ exchange) in the tail
  const float* q_f32;                // fp32 queries [n_queries][D]; read when inline_prep
  int inline_prep;                   // 1: the CTA builds its own resident Q image (no query_prep_kernel launch)
  int corpus_scale_exp;              // power-of-two pre-scale of an fp32 corpus (folded into qscale)
  int n_work;                        // CTAs that own tiles and write candidate lists (= lists per query)
  int tail_mode;                     // 0: candidates only; 1: + in-kernel merge; 2: + peer exchange and final merge
  int pdl;                           // launched with programmatic stream serialization: the scan may overlap the
                                     // previous search's tail; shared workspaces are touched only after griddepcontrol.wait.
                                     // 1: dependents released after the scan; 2 (grid == SM count): at kernel entry
  unsigned int* sync;                // [0] arrivals of worker CTAs, [1] departures of merging CTAs, [4] next tile (dynamic
                                     // scheduler); the host alternates between two such blocks from launch to launch
  int dyn_sched;                     // 1: supertiles after the first come from sync[4] (work stealing); 0: static round-robin
  unsigned int* err_flag;            // mapped host word, set to 1 if the arrival wait timed out
  float* out_score;                  // tail_mode != 0: final [n_queries][k]
  long long* out_idx;
  long long idx_offset;              // global index of this shard's row 0
  ExchangeArgs x;                    // tail_mode == 2
  int wide;                          // 1: wide-batch kernel (128 < B <= 256 in one pass; fused_search_pair_kernel, CTA pairs)
  int dmpair;                        // doc-major CTA-pair kernel (fused_search_dmpair_kernel), k <= 16: two accumulators and B <= 32,
                                     // or one accumulator and 32 < B <= 64
  int n_lists;                       // candidate lists per query the tail merges: n_work, or the worker PAIRS of the pair kernel
  int bound_rank;                    // pair kernel, shared bound: a CTA publishes its 1st or 2nd best score per query
  uint2* cbuf;                       // wide: [n_queries][n_lists][32] candidate buffers {score bits, local row id}
  float* g_best;                     // k > 16: [n_work][bpad] best score seen per (CTA, query), NaN = none yet; null = off
  unsigned long long* debug_ts;      // null, or [grid][8] globaltimer stamps of the kernel's phases (RS_DEBUG_TS=1)
};

cudaError_t launch_query_prep(const float* q_dev, int B, int D, int bpad, int dtype, int terms, int corpus_scale_exp,
                              uint8_t* qimg, float* qnorm, float* qscale, cudaStream_t stream);
cudaError_t launch_fused_search(const SearchParams& p, int grid, cudaStream_t stream);
int fused_smem_bytes(int bpad, int kc, int terms, int stages, int cps, int k, int q_resident, int qstages, int bufc,
                     int lists_global, int docm = 0);
int fused_reg_k();
int fused_wide_bufc();
int fused_pair_smem_bytes(int k);
int fused_pair_tail_capacity(int k);
int fused_pair_max_clusters(int k);
int fused_dmpair_smem_bytes(int kc, int terms, int planes, int stages, int cps, int bpad);
int fused_dmpair_tail_capacity(int kc, int terms, int planes, int stages, int cps, int bpad);
int fused_dmpair_max_clusters(int smem_bytes);  // CTA pairs of the pair kernel the current device can keep resident at once (0: none)
size_t fused_tail_smem_bytes(int n_lists, int k, int world);
int fused_tail_smem_capacity(int bpad, int kc, int terms, int stages, int cps, int k, int q_resident, int qstages);
cudaError_t launch_merge_i32(const float* cand_score, const int* cand_idx, int n_lists, int n_queries, int kc, int k,
                             long long idx_offset, float* out_score, long long* out_idx, cudaStream_t stream);
cudaError_t launch_exchange_merge(const float* cand_score, const int* cand_idx, int n_lists, int n_queries, int k,
                                  long long idx_offset, float* out_score, long long* out_idx, const ExchangeArgs& x,
                                  unsigned int* err_flag, cudaStream_t stream);
cudaError_t launch_merge_i64(const float* cand_score, const long long* cand_idx, int n_lists, int n_queries, int kc,
                             int k, long long score_list_stride, long long idx_list_stride, float* out_score,
                             long long* out_idx, cudaStream_t stream);

// ingest (rs_ingest.cu)
cudaError_t launch_tile_rows(const void* rows_dev, int src_dtype, long long row0, int nrows, int D, int dtype,
                             int scale_exp, uint8_t* corpus, float* dnorm, cudaStream_t stream);
// out_rows_dev receives rows [row0, row0 + nrows) as float, row-major
cudaError_t launch_untile_rows(const uint8_t* corpus, long long row0, int nrows, int D, int dtype, int scale_exp,
                               float* out_rows_dev, cudaStream_t stream);
cudaError_t launch_synth_fill(unsigned long long seed, long long global_row0, long long local_row0, int nrows, int D,
                              int dtype, int scale_exp, uint8_t* corpus, float* dnorm, cudaStream_t stream);
// mean over the T positions of hidden [B][T][D] (fp32 / fp16 / bf16) -> out [B][D] fp32 (batch_processor.py:91)
cudaError_t launch_mean_pool(const void* hidden_dev, int src_dtype, int B, int T_len, int D, float* out_dev, cudaStream_t stream);
cudaError_t launch_tile_minmax(const float* dnorm, int n_rows, int n_tiles, float2* minmax, cudaStream_t stream);

}  // namespace rs
