/* rs_retriever.h - C ABI of the B200-native similarity + top-k backend.
 *
 * This is the drop-in boundary for the one hot path of ngray0/RAG-Serving-System: the work that
 * rag_service/core/retriever.py does through CuPy.  Each entry point names the reference
 * interface it replaces (file:line relative to the reference checkout).  The reference is pure
 * Python, so the binding a maintainer adds is a ctypes stub; the signatures use
 * plain pointers and sizes only - no torch, no Python.h, no C++ types.
 *
 * Conventions: every function returns 0 on success and a negative RS_E_* code on failure; the
 * message is available from rs_last_error() (thread-local).  No function ever falls back to a
 * CPU computation: without a usable sm_100a device rs_create fails.
 * Result ordering everywhere: score descending, ties broken by the LOWER index.
 */
#ifndef RS_RETRIEVER_H_
#define RS_RETRIEVER_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RS_ABI_VERSION 2

#if defined(__GNUC__)
#define RS_API __attribute__((visibility("default")))
#else
#define RS_API
#endif

/* storage dtype of the corpus on device (the reference stores fp32 only, retriever.py:576-578) */
#define RS_DTYPE_BF16 0
#define RS_DTYPE_FP16 1
#define RS_DTYPE_FP32 2 /* kept as two fp16 planes (hi + lo*2^-11): 4 B/element, ~2^-22 relative */

/* similarity: RS_METRIC_DOT = SimpleRetriever (retriever.py:112);
 * RS_METRIC_COSINE = CupyRetriever.distance_cosine (retriever.py:594-611): dot/(|q||d|+1e-8), clipped */
#define RS_METRIC_DOT 0
#define RS_METRIC_COSINE 1

#define RS_OK 0
#define RS_E_INVALID (-1)   /* bad argument */
#define RS_E_CUDA (-2)      /* CUDA runtime error (message has the cudaError string) */
#define RS_E_STATE (-3)     /* call out of order (e.g. search before finalize) */
#define RS_E_UNSUPPORTED (-4) /* device is not sm_100 / shape outside kernel limits */

typedef struct rs_index rs_index;

typedef struct rs_config {
  int32_t struct_size; /* = sizeof(rs_config) */
  int32_t device;      /* CUDA device ordinal that owns this shard */
  int64_t n_rows;      /* documents in THIS shard (N_A of retriever.py:590 for one GPU) */
  int64_t row_offset;  /* global index of local row 0; added to every returned index */
  int32_t dim;         /* embedding dimension D (retriever.py:591), multiple of 64, <= 4096 */
  int32_t dtype;       /* RS_DTYPE_* storage type */
  int32_t scale_exp;   /* RS_DTYPE_FP32 only: rows are multiplied by 2^scale_exp before the fp16 split */
  int32_t max_batch;   /* largest B a search call may pass */
  int32_t max_k;       /* largest k a search call may pass (<= 128) */
  int32_t reserved;
} rs_config;

typedef struct rs_stats {
  int32_t struct_size;
  int32_t sm_count;         /* SMs on the device */
  int32_t grid;             /* CTAs of the last fused kernel */
  int32_t stages;           /* depth of its bulk-copy ring */
  int32_t smem_bytes;       /* its dynamic shared memory */
  int32_t passes;           /* query sub-batches the last search needed */
  int32_t launches;         /* kernels launched by the last search (prep + fused + merge per pass) */
  int32_t query_terms;      /* 1 or 2 */
  int32_t q_resident;       /* 1: query operand resident in shared memory; 0: streamed through the ring */
  int32_t sub;              /* 128-document tiles per supertile */
  int32_t nbuf;             /* TMEM accumulator buffers */
  int32_t mma_m;            /* MMA M of the last fused kernel (64 or 128) */
  int32_t overlapped;       /* 1: the last search was launched with programmatic dependent launch (RS_OVERLAP) */
  int32_t doc_major;        /* 1: the last search ran the doc-major kernel variant */
  int64_t bytes_streamed;   /* corpus bytes read by the last search (passes * shard bytes) */
  int64_t corpus_bytes;     /* device bytes of the tiled corpus image */
  int64_t total_launches;   /* kernels launched since rs_create */
} rs_stats;

RS_API int rs_version(void);
RS_API const char* rs_last_error(void);

/* Replaces CupyRetriever.__init__'s device allocation (retriever.py:580-586). Allocates the
 * tiled corpus image, norms and all per-call workspaces. */
RS_API int rs_create(const rs_config* cfg, rs_index** out);
RS_API int rs_destroy(rs_index* ix);

/* Replaces cp.asarray(doc_embeddings) (retriever.py:582): chunked ingest of rows
 * [row0, row0+nrows) (local numbering), row-major, src_dtype = RS_DTYPE_FP32 (float) or
 * RS_DTYPE_BF16 / RS_DTYPE_FP16 (uint16 bit patterns).  Converts to the storage dtype, writes
 * the tiled image and the document norms (computed ONCE here; the reference recomputes them on
 * every call, retriever.py:603).  `host_rows` may be pageable host memory (e.g. an np.memmap
 * slice; bounced through double-buffered pinned staging so the host copy of one chunk overlaps
 * the DMA and tiling of the previous one), pinned host memory (DMA straight from it) or DEVICE
 * memory (tiled in place, no copy) - detected with cudaPointerGetAttributes. */
RS_API int rs_load_rows(rs_index* ix, const void* host_rows, int32_t src_dtype, int64_t row0, int64_t nrows);

/* Fill the shard with the synthetic unit-normalised corpus (oracle: retrieval_oracle.synth_rows);
 * local row i is global row global_row0 + i, so any sharding reproduces the same corpus. */
RS_API int rs_fill_synthetic(rs_index* ix, uint64_t seed, int64_t global_row0);

/* Must be called after the last load/fill and before the first search. */
RS_API int rs_finalize(rs_index* ix);

/* Replaces the body of CupyRetriever.batch_retrieve between the H2D copy and the D2H copy
 * (retriever.py:736-773): scores + per-row top-k for B queries.  HOST buffers; blocking.
 *   q_host      float[B*dim], row-major (batch_processor.py:92 hands exactly this)
 *   metric      RS_METRIC_*
 *   query_terms 1 = queries rounded to the storage dtype; 2 = exact hi+lo split (fp32 storage
 *               always uses 2)
 *   out_idx     int64[B*k]  global indices, -1 padding when the shard has < k rows (:767-769)
 *   out_score   float[B*k]  matching scores, -inf padding */
RS_API int rs_search(rs_index* ix, const float* q_host, int32_t B, int32_t k, int32_t metric, int32_t query_terms,
              int64_t* out_idx, float* out_score);

/* Same with DEVICE buffers, asynchronous on `stream` (a cudaStream_t passed as void*; NULL =
 * the legacy default stream).  This is the per-GPU step of the sharded search; the outputs are
 * the B x k candidates that the NCCL all-gather exchanges.
 * Streams: an index has ONE set of workspaces.  Searches are serialised by the library: a search enqueued
 * on another stream than the previous one first waits (cudaStreamWaitEvent) for the previous search, so
 * callers may mix rs_search, rs_search_device on torch's current stream, etc.; a stream handed in must stay
 * valid until the next search call.  A device-side time-out (a wait inside the kernel gave up) cannot be
 * reported by this asynchronous call: it surfaces from rs_check, or from the next search call. */
RS_API int rs_search_device(rs_index* ix, const float* q_dev, int32_t B, int32_t k, int32_t metric, int32_t query_terms,
                     int64_t* out_idx_dev, float* out_score_dev, void* stream);
/* Query hand-off from the embedder.  Replaces `outputs.last_hidden_state.mean(dim=1)` followed by
 * `.cpu().numpy()` (batch_processor.py:91-92) and the H2D copy that rs_search would do: mean over ALL T positions of
 * hidden [B][T][D] (RS_DTYPE_FP32 / FP16 / BF16, contiguous, on `device`) into the fp32 query buffer out_q_dev [B][D],
 * asynchronous on `stream`; rs_search_device on the same stream then consumes it with no host round trip. */
RS_API int rs_pool_mean_device(int32_t device, const void* hidden_dev, int32_t src_dtype, int32_t B, int32_t T, int32_t D,
                        float* out_q_dev, void* stream);
/* RS_OK, or RS_E_CUDA if a kernel of an earlier asynchronous search raised the error word (and clears it).
 * Call after synchronising the stream.  rs_search and the other host-buffer calls check it themselves and
 * fail the call whose result is affected. */
RS_API int rs_check(rs_index* ix);

/* Switches of an index (replaces nothing in the reference; tests and A/B measurements).  The RS_* environment
 * variables of the same names are read ONCE, by rs_create; this call changes one afterwards.
 * RS_OVERLAP=1 lets consecutive rs_search_device / rs_search_exchange calls on one stream overlap: the scan of
 * search i+1 starts while the tail of search i (merge, exchange) drains (programmatic dependent launch).  The
 * caller then guarantees that the queries of a search are not written by the kernel that precedes it on that
 * stream (the prologue reads them before the predecessor has finished), and rs_get_kernel_times is unavailable. */
RS_API int rs_set_option(rs_index* ix, const char* name, int32_t value);

/* Final merge of n_lists sorted candidate lists per query (the multi-GPU step):
 * list l holds scores float[B][kc] at scores_dev + l*score_list_stride and indices int64[B][kc]
 * at idx_dev + l*idx_list_stride (strides in ELEMENTS; 0 = densely packed B*kc), i.e. exactly
 * what an all-gather of the rs_search_device outputs leaves in memory (one packed buffer per
 * rank or two separate ones); writes the global top-k [B][k]. */
RS_API int rs_merge_device(int32_t device, const float* scores_dev, const int64_t* idx_dev, int32_t n_lists, int32_t B,
                    int32_t kc, int32_t k, int64_t score_list_stride, int64_t idx_list_stride,
                    int64_t* out_idx_dev, float* out_score_dev, void* stream);

/* Multi-GPU candidate exchange over NVLink peer memory, one process per GPU (an alternative to
 * the NCCL all-gather + rs_merge_device above; same results).  Each rank creates an exchange
 * buffer, exports its 64-byte CUDA IPC handle, and imports every peer's handle (exchanged by the
 * host program, e.g. with torch.distributed.all_gather).  rs_search_exchange then runs, as ONE kernel
 * launch per search (two when the batch or dimension needs the separate query-prep kernel),
 *   query prep -> fused score+select -> [local merge + peer stores of each query's k candidates into every
 *   GPU's buffer as 16-byte {index, score, epoch} entries + poll of the local entries + final merge]
 * with no collective call; every rank ends with the same global top-k.  The search epoch lives in device
 * memory, so the call can be captured in a CUDA graph.  All ranks must issue the same sequence of searches
 * (same B and k); a rank whose peers never show up gets RS_E_CUDA (rs_check / next call), not a hang. */
typedef struct rs_xchg rs_xchg;
RS_API int rs_xchg_create(int32_t device, int32_t world, int32_t rank, int32_t max_batch, int32_t max_k, rs_xchg** out);
RS_API int rs_xchg_export(rs_xchg* x, void* handle64);
RS_API int rs_xchg_import(rs_xchg* x, int32_t peer_rank, const void* handle64);
RS_API int rs_xchg_destroy(rs_xchg* x);
RS_API int rs_search_exchange(rs_index* ix, rs_xchg* x, const float* q_dev, int32_t B, int32_t k, int32_t metric,
                       int32_t query_terms, int64_t* out_idx_dev, float* out_score_dev, void* stream);
/* The same search with HOST buffers, blocking (the sharded counterpart of rs_search): pinned H2D of the
 * queries, the search with the in-kernel exchange, D2H of the result; replayed from a CUDA graph from the
 * second call of a shape on.  out_idx == out_score == NULL: this rank takes part but does not need the result
 * on the host (only one rank of a sharded service answers the caller) - no D2H. */
RS_API int rs_search_exchange_host(rs_index* ix, rs_xchg* x, const float* q_host, int32_t B, int32_t k, int32_t metric,
                            int32_t query_terms, int64_t* out_idx, float* out_score);
/* Same-process peers: make `peer`'s buffer reachable from `x` through CUDA peer access instead of an IPC handle. */
RS_API int rs_xchg_attach(rs_xchg* x, rs_xchg* peer);

/* Single-process multi-GPU (replaces nothing: the reference has one GPU; this is what `CupyRetriever(...)`
 * constructed once in one process - main.py:52-55 - drives when it owns several devices).  `shards` are n
 * indices on n DISTINCT devices holding consecutive row ranges (row_offset) of one corpus, with the same dim,
 * dtype, max_batch and max_k; the group does not own them (destroy the group first).  rs_group_search is
 * rs_search for the whole corpus: one pinned copy of the queries, then per GPU an H2D + ONE fused kernel (candidate
 * exchange and final merge inside the kernels, over NVLink peer memory) issued in parallel - the calling thread drives
 * the first shard, one worker thread per further shard (they sleep between bursts of calls) - and ONE D2H from the
 * first shard's GPU; every shard's sequence is replayed from its own CUDA graph from the second call of a shape on. */
typedef struct rs_group rs_group;
RS_API int rs_group_create(rs_index** shards, int32_t n, rs_group** out);
RS_API int rs_group_search(rs_group* g, const float* q_host, int32_t B, int32_t k, int32_t metric, int32_t query_terms,
                    int64_t* out_idx, float* out_score);
RS_API int rs_group_destroy(rs_group* g);

/* Document arena (host side).  Replaces the B*k Python list lookups of retriever.py:776-790 /
 * :789 / :675 and the per-query joins of batch_processor.py:100 (`"\n---\n".join(docs)`): the documents live ONCE in
 * a UTF-8 byte arena (document i = arena[offsets[i], offsets[i+1])), copied by rs_docs_create; rs_docs_join writes,
 * for every query i, its first min(ks[i], stride) VALID hits (idx >= 0; -1 is the padding of retriever.py:767-769)
 * joined with `sep` into `out` - one memcpy per hit, no Python objects - and out_offsets[B+1] delimits the rows.
 * ks[i] <= 0 gives an empty row (retriever.py:778-780).  *need receives the total bytes; when out_cap < *need (or
 * out == NULL) only the sizes are computed and the caller calls again with a large enough buffer. */
typedef struct rs_docs rs_docs;
RS_API int rs_docs_create(const char* arena, const int64_t* offsets, int64_t n_docs, rs_docs** out);
RS_API int rs_docs_destroy(rs_docs* d);
RS_API int rs_docs_join(const rs_docs* d, const int64_t* idx, int32_t B, int32_t stride, const int32_t* ks, const char* sep,
                 int32_t sep_len, char* out, int64_t out_cap, int64_t* out_offsets, int64_t* need);

/* Read back the represented values of rows [row0,row0+nrows) as float (inverse of the tiling);
 * used by the tests (ingest round trip) and by the CPU-baseline sampler. */
RS_API int rs_get_rows(rs_index* ix, int64_t row0, int64_t nrows, float* out_host);
/* Same rows into DEVICE memory (float[nrows*dim] on the index's device), asynchronous on `stream`: lets a test
 * recompute scores of the whole shard on the GPU with an independent matmul. */
RS_API int rs_get_rows_device(rs_index* ix, int64_t row0, int64_t nrows, float* out_dev, void* stream);
RS_API int rs_get_norms(rs_index* ix, int64_t row0, int64_t nrows, float* out_host);
/* Raw tiled image bytes [off, off+nbytes) - layout tests compare them bit-exactly with the oracle. */
RS_API int rs_get_image(rs_index* ix, int64_t off, int64_t nbytes, void* out_host);

RS_API int rs_get_stats(rs_index* ix, rs_stats* out);
/* Launch plan the host would pick for a search (no device needed; used by the CPU tests):
 * out9 = {passes, queries_per_pass, bpad, ring_stages, q_resident, sub, nbuf, mma_m, chunks_per_stage}. */
RS_API int rs_debug_plan(int32_t dim, int32_t dtype, int32_t B, int32_t k, int32_t query_terms, int32_t* out9);
/* Timing experiments only: with RS_DEBUG_TS=1 (environment at rs_create, or rs_set_option) the fused kernel stamps %globaltimer
 * (ns) at 7 phase boundaries per CTA (entry, query prep done, first accumulator, last tile, arrival,
 * all arrived, merged); out = uint64[n_ctas][8] of the last launch. */
RS_API int rs_debug_timestamps(rs_index* ix, uint64_t* out, int32_t n_ctas);
/* Durations (ms) of the last n fused score+select kernel launches, oldest first, from CUDA events
 * recorded on the launching stream around each launch (ring of 256); synchronises on them. */
RS_API int rs_get_kernel_times(rs_index* ix, float* out_ms, int32_t n);

#ifdef __cplusplus
}
#endif
#endif /* RS_RETRIEVER_H_ */
