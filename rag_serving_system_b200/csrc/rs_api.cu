// C-ABI host side of the retriever backend (include/rs_retriever.h).  Owns all device memory,
// plans each search (make_plan: passes, Q residency, ring geometry, kernel orientation; normally ONE
// launch: query prep, score+select and candidate merge / exchange inside fused_search_kernel) and
// never computes a result on the CPU.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "../../include/rs_retriever.h"
#include "rs_kernels.h"
#include "rs_layout.h"

namespace {

thread_local char g_err[512] = "";

// Live indexes per device in this process, and a counter that moves whenever that number changes.  The in-kernel tail lets
// CTAs of a grid wait for each other, which is only safe while no OTHER grid of this kind can hold SMs of the device
// (SimpleRetriever + CupyRetriever in one service process, two indexes searched from two threads): with a second live
// index on the device every search falls back to the separate merge launch - a few microseconds, never a stall.
constexpr int kMaxDevices = 64;
std::atomic<int> g_live_indexes[kMaxDevices];
std::atomic<unsigned> g_live_epoch[kMaxDevices];

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define RS_CUDA(expr)                                                                             \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return fail(RS_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

constexpr int kMaxSmem = 232448;  // 227 KiB opt-in dynamic shared memory per CTA on sm_100
constexpr int kMaxStages = 24;
constexpr int64_t kIngestChunkBytes = 64ll << 20;

struct Plan {
  int per_pass;  // queries per fused launch
  int passes;
  int bpad;
  int stages;
  int q_resident;  // whole Q image in shared memory (small batches) vs. Q k-chunks through the ring
  int sub;         // 128-document tiles per supertile
  int nbuf;        // TMEM accumulator buffers
  int mma_m;       // MMA M: 64 when the padded batch fits, else 128
  int cps;         // k-chunks per ring stage
  int bufc;        // k > 16: candidate-buffer entries per query (32; 16 when that saves a pass)
  int lg;          // k > 16: sorted lists in global memory (L2) instead of shared memory
  int qstages;     // streamed Q: slots of the private Q ring
  int wide;        // one-pass wide-batch kernel (128 < B <= 256 per pass)
};

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

// Switches of an index.  Read from the environment ONCE, in rs_create (a search never calls getenv), and
// changeable afterwards with rs_set_option.  Every non-default value is parity-tested.
struct Options {
  int docm = 1;          // RS_DOCM         1: doc-major kernel wherever it applies; 2: one-term 16-bit searches only; 0: never
  int inline_prep = 1;   // RS_INLINE_PREP  0: separate query_prep_kernel launch
  int fused_tail = 1;    // RS_FUSED_TAIL   0: separate merge / exchange kernel launch
  int dyn_sched = 1;     // RS_DYN_SCHED    0: static round-robin tiles
  int shared_bound = 1;  // RS_SHARED_BOUND k > 16: 0 = no bound shared between CTAs
  int lists_global = 1;  // RS_LISTS_GLOBAL k > 16: 0 = sorted lists in shared memory
  int graph = 1;         // RS_GRAPH        0: host-buffer searches enqueue eagerly instead of replaying a CUDA graph
  int no_events = 0;     // RS_NO_EVENTS    1: no CUDA events around the fused launch
  int dmpair = 1;        // RS_DMPAIR       1: doc-major searches with two accumulators, or 32 < B <= 64, run on CTA pairs (cta_group::2)
                         //                 when the corpus has >= 8 tiles per SM; 2: for any corpus size (tests); 0: never
  int wide = 1;          // RS_WIDE         batches above 128: 1 = ONE pass on CTA pairs (fused_search_pair_kernel); 0 = passes of 128
  int overlap = 0;       // RS_OVERLAP      1: consecutive device-buffer searches on one stream may overlap (PDL)
  int debug_flags = 0;   // RS_DEBUG_FLAGS  timing experiments: skip MMAs (1) / the epilogue (2) / the tail merge (4); pair kernel:
                         //                 thresholds at +inf so nothing is ever appended (32), documents without the L2 evict_first hint (64)
  int debug_ts = 0;      // RS_DEBUG_TS     per-CTA %globaltimer phase stamps
  // planner overrides (tests and A/B runs): -1 / 0 = let the planner decide
  int qresident = -1, cps = 0, sub = 0, bufc = 0, lean = -1, mma_m = 0, nbuf = 0, qstages = 0;
};
struct OptDef {
  const char* name;
  int Options::*field;
};
const OptDef kOptDefs[] = {
    {"RS_DOCM", &Options::docm}, {"RS_INLINE_PREP", &Options::inline_prep}, {"RS_FUSED_TAIL", &Options::fused_tail},
    {"RS_DYN_SCHED", &Options::dyn_sched}, {"RS_SHARED_BOUND", &Options::shared_bound},
    {"RS_LISTS_GLOBAL", &Options::lists_global}, {"RS_GRAPH", &Options::graph},
    {"RS_NO_EVENTS", &Options::no_events}, {"RS_WIDE", &Options::wide}, {"RS_DMPAIR", &Options::dmpair}, {"RS_OVERLAP", &Options::overlap}, {"RS_DEBUG_FLAGS", &Options::debug_flags},
    {"RS_DEBUG_TS", &Options::debug_ts}, {"RS_QRESIDENT", &Options::qresident}, {"RS_CPS", &Options::cps},
    {"RS_SUB", &Options::sub}, {"RS_BUFC", &Options::bufc}, {"RS_LEAN", &Options::lean}, {"RS_MMA_M", &Options::mma_m},
    {"RS_NBUF", &Options::nbuf}, {"RS_QSTAGES", &Options::qstages},
};
Options options_from_env() {
  Options o;
  for (const OptDef& d : kOptDefs) o.*(d.field) = env_int(d.name, o.*(d.field));
  return o;
}

}  // namespace

struct rs_index {
  rs_config cfg;
  Options opt;
  int sm_count = 0;
  int kc = 0, planes = 1, n_tiles = 0;
  int plane_dtype = RS_DTYPE_BF16;
  size_t tile_bytes = 0, corpus_bytes = 0;
  uint8_t* corpus = nullptr;
  float* dnorm = nullptr;
  float2* minmax = nullptr;
  bool finalized = false;
  // workspaces
  uint8_t* qimg = nullptr;
  float* qnorm = nullptr;
  float* qscale = nullptr;
  float* cand_score = nullptr;
  int* cand_idx = nullptr;
  float* q_dev = nullptr;           // staging for host queries
  long long* out_idx_dev = nullptr;  // staging for host outputs: [indices][scores] of one call, back to back
  void* ingest_dev = nullptr;
  void* pinned = nullptr;  // pinned host staging (queries in, results out, ingest chunks); portable across devices
  size_t pinned_bytes = 0;
  cudaStream_t stream = nullptr;
  unsigned int* sync = nullptr;      // two blocks of 64 counters (arrivals, departures, next tile); launches alternate
  unsigned int launch_parity = 0;
  float* g_best = nullptr;           // shared bound table [sm_count][256] (k > 16 and wide searches)
  void* cbuf = nullptr;              // wide searches: candidate buffers [256][grid][32] x 8 B
  bool registered = false;           // counted in g_live_indexes
  unsigned live_epoch_seen = 0;      // g_live_epoch when the cached CUDA graphs were captured
  int dm_pair_clusters = 0;          // same for the doc-major pair kernel (two-accumulator searches, B <= 32, k <= 16)
  int pair_clusters = -1;            // CTA pairs of the pair kernel this device keeps resident at once (-1: not queried)
  unsigned long long* debug_ts = nullptr;  // RS_DEBUG_TS=1: [grid][8] phase stamps of the last fused launch
  unsigned int* err_flag = nullptr;  // mapped pinned word a kernel raises when a wait times out (1: own CTAs, 2: a peer GPU)
  // searches share the workspaces above, so a search on another stream than the previous one is ordered after it
  cudaStream_t last_stream = nullptr;
  bool have_last = false;
  cudaEvent_t order_ev = nullptr;
  // CUDA-event pairs around every fused-kernel launch (on the launching stream) so callers can
  // read the kernel's own duration without a profiler
  static constexpr int kEventRing = 256;
  cudaEvent_t ev_beg[kEventRing] = {}, ev_end[kEventRing] = {};
  cudaEvent_t ingest_done[2] = {};
  // host-buffer searches replay a captured CUDA graph per (B, k, metric, terms, exchange): one launch instead
  // of six stream operations - what a small (launch-bound) corpus feels
  struct GraphEntry {
    int B, k, metric, terms, uses;
    const void* xg;
    int want_out;
    cudaGraphExec_t exec;
  };
  std::vector<GraphEntry> graphs;
  bool capturing = false;
  long long ev_count = 0;
  rs_stats stats;
  std::mutex mu;
};

// Candidate exchange buffer of one GPU.  Peers map it (CUDA IPC across processes, peer access inside one
// process) and store their B x k candidates straight into their slot; entries carry the search epoch.
//   [0, 256)  header: u32 epoch, u32 departures
//   then      entries[2 parities][world slots][max_batch * max_k] of 16 bytes {int64 index, float score, u32 epoch}
struct rs_xchg {
  int device = 0, world = 1, rank = 0, max_batch = 0, max_k = 0;
  size_t slot_entries = 0, total_bytes = 0;
  uint8_t* local = nullptr;
  uint8_t* peer[rs::kMaxPeers] = {};     // peer[g] = GPU g's buffer as seen from this GPU (peer[rank] = local)
  bool opened[rs::kMaxPeers] = {};       // mapped through an IPC handle (to be closed)
};
constexpr size_t kXchgHeaderBytes = 256;

namespace {

rs::ExchangeArgs make_exchange_args(const rs_xchg* xg, int q0, bool last_pass) {
  rs::ExchangeArgs ea;
  memset(&ea, 0, sizeof(ea));
  ea.world = xg->world;
  ea.rank = xg->rank;
  ea.q0 = q0;
  ea.last_pass = last_pass ? 1 : 0;
  ea.epoch = reinterpret_cast<unsigned int*>(xg->local);
  for (int g = 0; g < xg->world; ++g)
    ea.dest[g] = reinterpret_cast<uint4*>(xg->peer[g] + kXchgHeaderBytes) + (size_t)xg->rank * xg->slot_entries;
  ea.local = reinterpret_cast<const uint4*>(xg->local + kXchgHeaderBytes);
  ea.slot_entries = (long long)xg->slot_entries;
  ea.parity_entries = (long long)(xg->slot_entries * xg->world);
  return ea;
}

int max_cap_rows(int kc, int terms, int k, int min_stages, int cps, int q_resident, int qstages, int bufc = 32, int lg = 0) {
  // largest bpad (multiple of 8, <= 128) that leaves at least min_stages ring slots
  int best = 0;
  for (int bpad = 8; bpad <= 128; bpad += 8) {
    if (rs::fused_smem_bytes(bpad, kc, terms, min_stages, cps, k, q_resident, qstages, bufc, lg) <= kMaxSmem) best = bpad;
  }
  return best;
}

// Small batches keep the query operand resident in shared memory (one tile per accumulator,
// double-buffered TMEM).  Larger batches stream Q k-chunks through a private ring and amortise each
// chunk over `sub` document tiles that accumulate side by side in TMEM.  Measured on B200: what
// limits the stream is the number of barrier round trips per byte, so the planner prefers, in
// order: fewest passes, resident Q, the largest ring stage (cps k-chunks of 16 KiB), the largest sub.
int make_plan(const rs_index* ix, const Options& o, int B, int k, int terms, Plan* pl) {
  const int nacc = (terms == 2 || ix->planes == 2) ? 2 : 1;
  const bool lists_in_smem = k > rs::fused_reg_k();
  // k > 16: keep the sorted lists in global memory (default) so the ring keeps its bytes; RS_LISTS_GLOBAL=0: shared memory
  const int lg = lists_in_smem && o.lists_global != 0 ? 1 : 0;
  const int qstages = o.qstages > 0 ? o.qstages : 2 * terms;
  pl->qstages = qstages;
  pl->wide = 0;
  if (o.wide != 0 && B > 128 && terms == 1 && ix->planes == 1 && ix->pair_clusters != 0 && rs::fused_pair_smem_bytes(k) <= kMaxSmem) {
    // Wide batches (BASELINE config 5: B=256): both 128-row halves of the batch share ONE pass over the shard on CTA pairs
    // (fused_search_pair_kernel: M=256 x N=256 MMAs across two SMs, two TMEM accumulators); the general kernel would
    // stream the shard once per 128 queries.
    pl->wide = 1;
    pl->passes = (B + 255) / 256;
    pl->per_pass = (B + pl->passes - 1) / pl->passes;
    pl->bpad = 256;
    pl->q_resident = 0;
    pl->sub = 2;
    pl->bufc = rs::fused_wide_bufc();
    pl->lg = 1;
    pl->qstages = 0;
    pl->cps = (o.cps == 2 && ix->kc % 2 == 0) ? 2 : 1;  // k-chunks per ring slot (32 KiB each)
    pl->stages = 6 / pl->cps;
    pl->nbuf = 2;
    pl->mma_m = 256;
    return RS_OK;
  }
  long long best_cost = -1;
  for (int resident = 1; resident >= 0; --resident) {
    if (o.qresident >= 0 && resident != o.qresident) continue;
    for (int cps = 4; cps >= 1; cps >>= 1) {
      if (ix->kc % cps != 0 || (o.cps && cps != o.cps)) continue;
      if (!resident && cps == 4) continue;
      for (int sub = resident ? (o.sub ? o.sub : 1) : (nacc == 1 ? 4 : 2); sub >= 1; sub >>= 1) {
        if (o.sub && sub != o.sub) continue;
        // k > 16 only: a 16-entry candidate buffer (twice the flushes) and a 48 KiB ring are accepted
        // when they save a whole pass over the shard
        for (int bufc = 32; bufc >= (lists_in_smem ? 16 : 32); bufc >>= 1) {
          if (o.bufc && bufc != o.bufc) continue;
          for (int lean = 0; lean <= (lists_in_smem ? 1 : 0); ++lean) {
            if (o.lean >= 0 && lean != o.lean) continue;
            // ring depth: at least 64 KiB in flight (48 KiB when lean), and in streamed mode one whole
            // k-chunk group + 1
            int need = std::max(2, ((lean ? 3 : 4) + cps - 1) / cps);
            if (!resident) need = std::max(need, ix->planes * sub + 1);
            const int cap = max_cap_rows(ix->kc, terms, k, need, cps, resident, qstages, bufc, lg);
            if (cap <= 0) continue;
            const int passes = (B + cap - 1) / cap;
            const long long cost = ((long long)passes << 32) | ((long long)(1 - resident) << 28) |
                                   ((long long)(bufc == 32 ? 0 : 1) << 24) | ((long long)lean << 20) |
                                   ((long long)(4 - cps) << 16) | (long long)(8 - sub);
            if (best_cost < 0 || cost < best_cost) {
              best_cost = cost;
              pl->q_resident = resident;
              pl->cps = cps;
              pl->sub = sub;
              pl->passes = passes;
              pl->stages = need;
              pl->bufc = bufc;
              pl->lg = lg;
            }
          }
        }
      }
    }
  }
  if (best_cost < 0) return fail(RS_E_UNSUPPORTED, "B=%d k=%d dim=%d does not fit in shared memory", B, k, ix->cfg.dim);
  pl->per_pass = (B + pl->passes - 1) / pl->passes;
  pl->bpad = (pl->per_pass + 7) / 8 * 8;
  pl->mma_m = o.mma_m ? o.mma_m : (pl->bpad <= 64 ? 64 : 128);
  if (pl->mma_m != 128 && !(pl->mma_m == 64 && pl->bpad <= 64)) return fail(RS_E_INVALID, "RS_MMA_M=%d invalid for bpad=%d", pl->mma_m, pl->bpad);
  // as many accumulator buffers as the 512 TMEM columns hold: the epilogue's start-up burst (every
  // score beats an empty list) then stalls the MMA/TMA pipeline less
  pl->nbuf = o.nbuf ? o.nbuf : std::max(1, 4 / (pl->sub * nacc));
  const int cols = pl->sub * pl->nbuf * nacc;
  if (pl->nbuf < 1 || pl->nbuf > 4 || (cols != 1 && cols != 2 && cols != 4))
    return fail(RS_E_INVALID, "sub=%d nbuf=%d nacc=%d does not map onto 128/256/512 TMEM columns", pl->sub, pl->nbuf, nacc);
  auto bytes = [&](int stages) {
    return rs::fused_smem_bytes(pl->bpad, ix->kc, terms, stages, pl->cps, k, pl->q_resident, qstages, pl->bufc, pl->lg);
  };
  while (pl->stages < kMaxStages && bytes(pl->stages + 1) <= kMaxSmem) ++pl->stages;
  if (bytes(pl->stages) > kMaxSmem) return fail(RS_E_UNSUPPORTED, "B=%d k=%d does not fit in shared memory", B, k);
  return RS_OK;
}

void free_all(rs_index* ix) {
  cudaSetDevice(ix->cfg.device);
  cudaFree(ix->corpus);
  cudaFree(ix->dnorm);
  cudaFree(ix->minmax);
  cudaFree(ix->qimg);
  cudaFree(ix->qnorm);
  cudaFree(ix->qscale);
  cudaFree(ix->cand_score);
  cudaFree(ix->cand_idx);
  cudaFree(ix->q_dev);
  cudaFree(ix->out_idx_dev);
  cudaFree(ix->ingest_dev);
  cudaFree(ix->sync);
  cudaFree(ix->g_best);
  cudaFree(ix->cbuf);
  cudaFree(ix->debug_ts);
  if (ix->err_flag) cudaFreeHost(ix->err_flag);
  if (ix->pinned) cudaFreeHost(ix->pinned);
  for (auto& g : ix->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  for (int i = 0; i < 2; ++i)
    if (ix->ingest_done[i]) cudaEventDestroy(ix->ingest_done[i]);
  for (int i = 0; i < rs_index::kEventRing; ++i) {
    if (ix->ev_beg[i]) cudaEventDestroy(ix->ev_beg[i]);
    if (ix->ev_end[i]) cudaEventDestroy(ix->ev_end[i]);
  }
  if (ix->order_ev) cudaEventDestroy(ix->order_ev);
  if (ix->stream) cudaStreamDestroy(ix->stream);
}

void drop_graphs(rs_index* ix) {  // the corpus changed: cached launch sequences are stale
  for (auto& g : ix->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  ix->graphs.clear();
}

// A kernel raised the mapped error word (a wait timed out): clear it, reset the counters and fail.
int take_device_error(rs_index* ix, cudaStream_t stream) {
  const unsigned int e = *reinterpret_cast<volatile unsigned int*>(ix->err_flag);
  if (!e) return RS_OK;
  *ix->err_flag = 0;
  cudaMemsetAsync(ix->sync, 0, 512, stream);
  return fail(RS_E_CUDA, e == 2u ? "a search timed out waiting for a peer GPU's candidates (peer stalled or out of step); results are invalid"
                                 : "a search timed out waiting for its own CTAs (device shared with another job?); results are invalid");
}

// The workspaces are per index, not per stream: order this search after the previous one when the stream changed.
int order_after_previous(rs_index* ix, cudaStream_t stream) {
  if (ix->have_last && ix->last_stream != stream) {
    RS_CUDA(cudaEventRecord(ix->order_ev, ix->last_stream));
    RS_CUDA(cudaStreamWaitEvent(stream, ix->order_ev, 0));
  }
  ix->last_stream = stream;
  ix->have_last = true;
  return RS_OK;
}

// one search on device buffers; all launches go to `stream`
int search_device_locked(rs_index* ix, const float* q_dev, int B, int k, int metric, int terms, long long* out_idx,
                         float* out_score, cudaStream_t stream, rs_xchg* xg = nullptr) {
  const Options& o = ix->opt;
  if (!ix->finalized) return fail(RS_E_STATE, "rs_finalize has not been called");
  if (B <= 0 || B > ix->cfg.max_batch) return fail(RS_E_INVALID, "B=%d outside [1,%d]", B, ix->cfg.max_batch);
  if (k <= 0 || k > ix->cfg.max_k) return fail(RS_E_INVALID, "k=%d outside [1,%d]", k, ix->cfg.max_k);
  if (metric != RS_METRIC_DOT && metric != RS_METRIC_COSINE) return fail(RS_E_INVALID, "bad metric %d", metric);
  if (ix->planes == 2) terms = 2;
  if (terms != 1 && terms != 2) return fail(RS_E_INVALID, "query_terms must be 1 or 2");
  Plan pl;
  int rc = make_plan(ix, o, B, k, terms, &pl);
  if (rc != RS_OK) return rc;
  if (!ix->capturing) {
    rc = take_device_error(ix, stream);  // raised by an earlier device-buffer search (rs_check reads it sooner)
    if (rc != RS_OK) return rc;
    rc = order_after_previous(ix, stream);
    if (rc != RS_OK) return rc;
  }
  // doc-major kernel variant (documents on the MMA's M): half the tensor work and a quarter of the Q-operand reads
  // at B=32; for B <= 32, k <= 16, resident Q (RS_DOCM=0: off; RS_DOCM=2: only one-term 16-bit searches)
  const bool docm = !pl.wide && pl.q_resident == 1 && (o.docm == 1 ? true : (terms == 1 && ix->planes == 1)) &&
                    k <= rs::fused_reg_k() && pl.bpad <= 32 && pl.sub == 1 && o.docm != 0 &&
                    rs::fused_smem_bytes(pl.bpad, ix->kc, terms, pl.stages, pl.cps, k, 1, pl.qstages, pl.bufc, pl.lg, 1) <= kMaxSmem;
  if (docm) pl.nbuf = 8;
  // Two accumulators (fp32 storage / two query terms) on a large corpus: the CTA-pair form of the doc-major kernel,
  // where each CTA keeps only HALF of the Q image and the document ring doubles (fused_search_dmpair_kernel).
  const int nacc = (terms == 2 || ix->planes == 2) ? 2 : 1;
  // (B <= 16: the whole image is 64 KiB anyway, and the one-CTA kernel keeps its in-kernel query prep)
  const int dm_min_tiles = o.dmpair == 2 ? 1 : 8 * ix->sm_count;
  bool dmpair = docm && nacc == 2 && pl.bpad > 16 && o.dmpair != 0 && ix->dm_pair_clusters > 0 && ix->n_tiles >= dm_min_tiles;
  // One accumulator (16-bit storage, one query term), 32 < B <= 64: the same kernel with N = 64 - each CTA holds 32
  // queries' image (64 KiB at D = 1024), where the one-CTA kernels would need 128 KiB resident or stream Q.
  const bool dmpair64 = !pl.wide && !docm && nacc == 1 && B > 32 && B <= 64 && k <= rs::fused_reg_k() && o.docm != 0 &&
                        o.dmpair != 0 && ix->dm_pair_clusters > 0 && ix->n_tiles >= dm_min_tiles;
  if (dmpair || dmpair64) {
    const int bpad = dmpair64 ? 64 : 32;
    int cps = o.cps > 0 ? o.cps : (ix->planes == 2 ? 2 : 4);  // 64 KiB ring stages
    if (ix->kc % cps != 0) cps = 1;
    const int fixed = rs::fused_dmpair_smem_bytes(ix->kc, terms, ix->planes, 0, cps, bpad);
    const int stages = std::min(8, (kMaxSmem - fixed) / (ix->planes * cps * RS_BLOCK_BYTES));
    dmpair = stages >= 2;  // (larger dimensions: the plan made above stays)
    if (dmpair) {
      pl.bpad = bpad;
      pl.q_resident = 1;
      pl.sub = 1;
      pl.nbuf = 8;
      pl.passes = 1;
      pl.per_pass = B;
      pl.cps = cps;
      pl.stages = stages;
    }
  }
  // CTAs that own tiles.  Pair kernel: both CTAs of every pair that owns a supertile (two tiles)
  const int n_work = pl.wide ? 2 * std::min(ix->pair_clusters, (ix->n_tiles + 1) / 2)
                     : dmpair     ? 2 * std::min(ix->dm_pair_clusters, (ix->n_tiles + 1) / 2)
                                  : std::min(ix->sm_count, ix->n_tiles);
  const int n_lists = pl.wide ? n_work / 2 : n_work;  // candidate lists per query
  // single-launch search (default): query prep in the fused kernel's prologue (resident-Q plans) and the
  // candidate merge (+ exchange) in its tail.  RS_INLINE_PREP=0 / RS_FUSED_TAIL=0 fall back to the
  // separate query_prep_kernel / merge kernels.
  const bool inline_prep = !pl.wide && !dmpair && pl.q_resident == 1 && ix->cfg.dim <= 1024 && pl.bpad <= 32 && o.inline_prep != 0;
  const int tail_cap = dmpair ? rs::fused_dmpair_tail_capacity(ix->kc, terms, ix->planes, pl.stages, pl.cps, pl.bpad)
                       : pl.wide ? rs::fused_pair_tail_capacity(k)
                               : rs::fused_tail_smem_capacity(pl.bpad, ix->kc, terms, pl.stages, pl.cps, k, pl.q_resident, pl.qstages);
  const bool alone = !ix->registered || g_live_indexes[ix->cfg.device].load() <= 1;  // no other live index on this device
  const bool fused_tail = o.fused_tail != 0 && alone && rs::fused_tail_smem_bytes(n_lists, k, xg ? xg->world : 1) <= (size_t)tail_cap;
  // pair kernel: whole pairs; every resident pair when the tail needs merging CTAs for up to 256 queries
  const int grid = pl.wide ? (fused_tail ? 2 * ix->pair_clusters : n_work)
                   : dmpair     ? n_work  // >= 8 tiles per SM: far more worker CTAs than queries
                                : fused_tail ? std::max(n_work, std::min(pl.per_pass, ix->sm_count)) : n_work;
  // overlapped launches (RS_OVERLAP=1): the scan of this search may run while the previous search's tail drains.
  // Only for plans that touch no shared workspace before the tail (register lists, Q image built in shared memory).
  const bool pdl = o.overlap != 0 && !ix->capturing && inline_prep && fused_tail && pl.passes == 1 &&
                   k <= rs::fused_reg_k() && !o.debug_ts;
  // bound shared between CTAs: needs (lists x published rank) >= k documents known to reach it
  const int bound_rank = pl.wide && n_lists < k ? 2 : 1;
  const bool shared_bound = (k > rs::fused_reg_k() || pl.wide) && n_lists * bound_rank >= k && o.shared_bound != 0;
  ix->stats.grid = grid;
  ix->stats.stages = pl.stages;
  ix->stats.smem_bytes = dmpair ? rs::fused_dmpair_smem_bytes(ix->kc, terms, ix->planes, pl.stages, pl.cps, pl.bpad)
                         : pl.wide ? rs::fused_pair_smem_bytes(k)
                                 : rs::fused_smem_bytes(pl.bpad, ix->kc, terms, pl.stages, pl.cps, k, pl.q_resident, pl.qstages, pl.bufc, pl.lg, docm);
  ix->stats.q_resident = pl.q_resident;
  ix->stats.sub = pl.sub;
  ix->stats.nbuf = pl.nbuf;
  ix->stats.mma_m = dmpair ? 256 : pl.mma_m;
  ix->stats.overlapped = pdl ? 1 : 0;
  ix->stats.doc_major = (docm || dmpair) ? 1 : 0;
  ix->stats.passes = pl.passes;
  ix->stats.query_terms = terms;
  ix->stats.launches = 0;
  ix->stats.bytes_streamed = 0;
  if (xg) {
    if (B > xg->max_batch || k > xg->max_k) return fail(RS_E_INVALID, "B=%d k=%d exceed the exchange buffer (%d, %d)", B, k, xg->max_batch, xg->max_k);
    for (int g = 0; g < xg->world; ++g)
      if (!xg->peer[g]) return fail(RS_E_STATE, "exchange buffer of rank %d has not been imported", g);
  }
  for (int pass = 0; pass < pl.passes; ++pass) {
    const int b0 = pass * pl.per_pass;
    const int nb = std::min(pl.per_pass, B - b0);
    const int corpus_scale_exp = ix->planes == 2 ? ix->cfg.scale_exp : 0;
    if (!inline_prep)
      RS_CUDA(rs::launch_query_prep(q_dev + (size_t)b0 * ix->cfg.dim, nb, ix->cfg.dim, pl.bpad, ix->plane_dtype, terms,
                                    corpus_scale_exp, ix->qimg, ix->qnorm, ix->qscale, stream));
    rs::SearchParams p;
    memset(&p, 0, sizeof(p));
    p.q_f32 = q_dev + (size_t)b0 * ix->cfg.dim;
    p.inline_prep = inline_prep ? 1 : 0;
    p.corpus_scale_exp = corpus_scale_exp;
    p.n_work = n_work;
    p.n_lists = n_lists;
    p.bound_rank = bound_rank;
    p.tail_mode = fused_tail ? (xg ? 2 : 1) : 0;
    p.pdl = pdl ? (grid == ix->sm_count && n_work == grid ? 2 : 1) : 0;
    ix->launch_parity ^= 1u;
    p.sync = ix->sync + 64 * ix->launch_parity;
    p.dyn_sched = o.dyn_sched != 0 ? 1 : 0;
    if (p.dyn_sched && !fused_tail) RS_CUDA(cudaMemsetAsync(p.sync + 4, 0, sizeof(unsigned int), stream));  // the tail resets it otherwise
    p.err_flag = ix->err_flag;
    p.out_score = out_score + (size_t)b0 * k;
    p.out_idx = out_idx + (size_t)b0 * k;
    p.idx_offset = ix->cfg.row_offset;
    if (shared_bound) {
      // 0xff bytes = NaN = "this CTA has not published a score for this query yet"
      RS_CUDA(cudaMemsetAsync(ix->g_best, 0xff, (size_t)n_lists * pl.bpad * sizeof(float), stream));
      p.g_best = ix->g_best;
    }
    if (o.debug_ts) {
      if (!ix->debug_ts) RS_CUDA(cudaMalloc(&ix->debug_ts, (size_t)ix->sm_count * 8 * sizeof(unsigned long long)));
      RS_CUDA(cudaMemsetAsync(ix->debug_ts, 0, (size_t)ix->sm_count * 8 * sizeof(unsigned long long), stream));
      p.debug_ts = ix->debug_ts;
    }
    p.corpus = ix->corpus;
    p.qimg = ix->qimg;
    p.qnorm = ix->qnorm;
    p.qscale = ix->qscale;
    p.dnorm = ix->dnorm;
    p.tile_norm_minmax = ix->minmax;
    p.cand_score = ix->cand_score;
    p.cand_idx = ix->cand_idx;
    p.n_rows = (int)ix->cfg.n_rows;
    p.n_tiles = ix->n_tiles;
    p.n_queries = nb;
    p.kc = ix->kc;
    p.planes = ix->planes;
    p.terms = terms;
    p.bpad = pl.bpad;
    p.k = k;
    p.metric = metric;
    p.dtype = ix->plane_dtype;
    p.stages = pl.stages;
    p.cps = pl.cps;
    p.bufc = pl.bufc;
    p.lists_global = pl.lg;
    p.docm = (docm || dmpair) ? 1 : 0;
    p.dmpair = dmpair ? 1 : 0;
    p.wide = pl.wide;
    p.cbuf = reinterpret_cast<uint2*>(ix->cbuf);
    p.qstages = pl.qstages;
    p.q_resident = pl.q_resident;
    p.sub = pl.sub;
    p.nbuf = pl.nbuf;
    p.mma_m = pl.mma_m;
    p.debug_flags = o.debug_flags;
    p.lo_coef = ix->plane_dtype == RS_DTYPE_FP16 ? ldexpf(1.0f, -RS_FP16_SPLIT_SHIFT) : 1.0f;
    rs::ExchangeArgs ea;
    memset(&ea, 0, sizeof(ea));
    if (xg) ea = make_exchange_args(xg, b0, pass == pl.passes - 1);
    p.x = ea;
    // events around the launch would sit between two overlapped kernels and serialise them: not in that mode
    if (ix->capturing || o.no_events || pdl) {
      RS_CUDA(rs::launch_fused_search(p, grid, stream));
    } else {
      const int slot = (int)(ix->ev_count % rs_index::kEventRing);
      RS_CUDA(cudaEventRecord(ix->ev_beg[slot], stream));
      RS_CUDA(rs::launch_fused_search(p, grid, stream));
      RS_CUDA(cudaEventRecord(ix->ev_end[slot], stream));
      ++ix->ev_count;
    }
    if (!fused_tail) {
      if (!xg) {
        RS_CUDA(rs::launch_merge_i32(ix->cand_score, ix->cand_idx, n_lists, nb, k, k, ix->cfg.row_offset,
                                     out_score + (size_t)b0 * k, out_idx + (size_t)b0 * k, stream));
      } else {
        // local merge + exchange + final merge in one kernel: rows go straight into every GPU's buffer
        RS_CUDA(rs::launch_exchange_merge(ix->cand_score, ix->cand_idx, n_lists, nb, k, ix->cfg.row_offset,
                                          out_score + (size_t)b0 * k, out_idx + (size_t)b0 * k, ea, ix->err_flag, stream));
      }
    }
    ix->stats.launches += 1 + (inline_prep ? 0 : 1) + (fused_tail ? 0 : 1);
    ix->stats.bytes_streamed += (int64_t)ix->corpus_bytes;
  }
  ix->stats.total_launches += ix->stats.launches;
  return RS_OK;
}

// Host-buffer search on ix->stream: pinned H2D of the queries -> search -> ONE D2H of [indices][scores]; from the
// second call of a shape on, the sequence is replayed from a captured CUDA graph.  out_idx == nullptr: this caller
// does not need the result on the host (a non-root rank of a sharded search): no D2H, but still synchronous.
int search_host_locked(rs_index* ix, rs_xchg* xg, const float* q_host, int B, int k, int metric, int terms,
                       long long* out_idx, float* out_score) {
  if (B <= 0 || B > ix->cfg.max_batch) return fail(RS_E_INVALID, "B=%d outside [1,%d]", B, ix->cfg.max_batch);
  if (k <= 0 || k > ix->cfg.max_k) return fail(RS_E_INVALID, "k=%d outside [1,%d]", k, ix->cfg.max_k);
  if (!ix->finalized) return fail(RS_E_STATE, "rs_finalize has not been called");
  const size_t q_bytes = (size_t)B * ix->cfg.dim * sizeof(float);
  const size_t n_out = (size_t)B * k;
  char* pin = static_cast<char*>(ix->pinned);
  long long* pin_idx = reinterpret_cast<long long*>(pin + ((q_bytes + 255) & ~size_t(255)));
  float* pin_score = reinterpret_cast<float*>(pin_idx + n_out);
  const int want_out = out_idx != nullptr ? 1 : 0;
  if (ix->planes == 2) terms = 2;
  auto enqueue = [&]() -> int {  // H2D -> (prep ->) fused (-> merge) -> D2H, all on ix->stream
    RS_CUDA(cudaMemcpyAsync(ix->q_dev, pin, q_bytes, cudaMemcpyHostToDevice, ix->stream));
    float* score_dev = reinterpret_cast<float*>(ix->out_idx_dev + n_out);  // right behind this call's indices
    int rc = search_device_locked(ix, ix->q_dev, B, k, metric, terms, ix->out_idx_dev, score_dev, ix->stream, xg);
    if (rc != RS_OK) return rc;
    if (want_out)
      RS_CUDA(cudaMemcpyAsync(pin_idx, ix->out_idx_dev, n_out * (sizeof(long long) + sizeof(float)), cudaMemcpyDeviceToHost,
                              ix->stream));  // pin_score == pin_idx + n_out
    return RS_OK;
  };
  int rc = take_device_error(ix, ix->stream);
  if (rc != RS_OK) return rc;
  rc = order_after_previous(ix, ix->stream);
  if (rc != RS_OK) return rc;
  memcpy(pin, q_host, q_bytes);
  if (ix->registered) {  // an index was created or destroyed on this device: the cached launch sequences may have the wrong tail
    const unsigned ep = g_live_epoch[ix->cfg.device].load();
    if (ep != ix->live_epoch_seen) {
      drop_graphs(ix);
      ix->live_epoch_seen = ep;
    }
  }
  rs_index::GraphEntry* ge = nullptr;
  if (ix->opt.graph) {
    for (auto& g : ix->graphs)
      if (g.B == B && g.k == k && g.metric == metric && g.terms == terms && g.xg == xg && g.want_out == want_out) ge = &g;
    if (!ge && ix->graphs.size() < 64) {
      ix->graphs.push_back({B, k, metric, terms, 0, xg, want_out, nullptr});
      ge = &ix->graphs.back();
    }
  }
  if (ge && !ge->exec && ge->uses >= 1) {
    // second call with this shape: capture the sequence (the first, eager call set the kernels'
    // shared-memory attributes and validated the arguments)
    ix->capturing = true;
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaStreamBeginCapture(ix->stream, cudaStreamCaptureModeThreadLocal);
    int crc = ce == cudaSuccess ? enqueue() : RS_E_CUDA;
    cudaError_t ee = cudaStreamEndCapture(ix->stream, &graph);
    ix->capturing = false;
    if (crc == RS_OK && ee == cudaSuccess && graph && cudaGraphInstantiate(&ge->exec, graph, 0) != cudaSuccess) ge->exec = nullptr;
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    if (!ge->exec) ge->uses = -1000000;  // capture not possible here: stay on the eager path
  }
  if (ge && ge->exec) {
    RS_CUDA(cudaGraphLaunch(ge->exec, ix->stream));
    ix->stats.total_launches += ix->stats.launches;
  } else {
    rc = enqueue();
    if (rc != RS_OK) return rc;
    if (ge) ++ge->uses;
  }
  RS_CUDA(cudaStreamSynchronize(ix->stream));
  rc = take_device_error(ix, ix->stream);  // THIS call's result is invalid if a wait timed out
  if (rc != RS_OK) return rc;
  if (want_out) {
    memcpy(out_idx, pin_idx, n_out * sizeof(long long));
    memcpy(out_score, pin_score, n_out * sizeof(float));
  }
  return RS_OK;
}

}  // namespace

extern "C" {

RS_API int rs_version(void) { return RS_ABI_VERSION; }
RS_API const char* rs_last_error(void) { return g_err; }

RS_API int rs_create(const rs_config* cfg, rs_index** out) {
  if (!cfg || !out) return fail(RS_E_INVALID, "null argument");
  *out = nullptr;
  if (cfg->struct_size != (int32_t)sizeof(rs_config)) return fail(RS_E_INVALID, "rs_config.struct_size mismatch");
  if (cfg->n_rows <= 0 || cfg->n_rows > 0x7fffff00ll) return fail(RS_E_INVALID, "n_rows=%lld out of range", (long long)cfg->n_rows);
  if (cfg->dim <= 0 || cfg->dim % 64 != 0 || cfg->dim > 4096) return fail(RS_E_INVALID, "dim=%d must be a multiple of 64 in [64,4096]", cfg->dim);
  if (cfg->dtype < RS_DTYPE_BF16 || cfg->dtype > RS_DTYPE_FP32) return fail(RS_E_INVALID, "bad dtype %d", cfg->dtype);
  if (cfg->max_batch <= 0 || cfg->max_batch > 65536) return fail(RS_E_INVALID, "bad max_batch %d", cfg->max_batch);
  if (cfg->max_k <= 0 || cfg->max_k > 128) return fail(RS_E_INVALID, "max_k=%d outside [1,128]", cfg->max_k);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(RS_E_CUDA, "no CUDA device available (%s); this backend has no CPU fallback", cudaGetErrorString(e));
  if (cfg->device < 0 || cfg->device >= ndev) return fail(RS_E_INVALID, "device %d of %d", cfg->device, ndev);
  RS_CUDA(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  RS_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10)
    return fail(RS_E_UNSUPPORTED, "device %d is sm_%d%d; the kernels are sm_100a only", cfg->device, prop.major, prop.minor);

  rs_index* ix = new (std::nothrow) rs_index();
  if (!ix) return fail(RS_E_INVALID, "out of host memory");
  ix->cfg = *cfg;
  ix->opt = options_from_env();
  ix->sm_count = prop.multiProcessorCount;
  ix->kc = cfg->dim / RS_KCHUNK_ELEMS;
  ix->planes = cfg->dtype == RS_DTYPE_FP32 ? 2 : 1;
  ix->plane_dtype = cfg->dtype == RS_DTYPE_BF16 ? RS_DTYPE_BF16 : RS_DTYPE_FP16;
  ix->n_tiles = (int)((cfg->n_rows + RS_TILE_ROWS - 1) / RS_TILE_ROWS);
  ix->tile_bytes = (size_t)ix->planes * ix->kc * RS_BLOCK_BYTES;
  ix->corpus_bytes = ix->tile_bytes * ix->n_tiles;
  memset(&ix->stats, 0, sizeof(ix->stats));
  ix->stats.struct_size = sizeof(rs_stats);
  ix->stats.sm_count = ix->sm_count;
  ix->stats.corpus_bytes = (int64_t)ix->corpus_bytes;

  const int grid = std::min(ix->sm_count, ix->n_tiles);
  const int cap1 = max_cap_rows(ix->kc, 1, 1, 3, 1, 0, 2);
  if (cap1 <= 0) {
    delete ix;
    return fail(RS_E_UNSUPPORTED, "dim=%d does not fit the shared-memory pipeline", cfg->dim);
  }
  const size_t pass_rows = 256;  // queries per fused launch: up to 128 (general kernel) or 256 (wide kernel)
  const size_t qimg_bytes = (size_t)2 * ix->kc * pass_rows * RS_KCHUNK_BYTES;
  const size_t cand_elems = (size_t)grid * pass_rows * cfg->max_k;
  const size_t q_bytes = (size_t)cfg->max_batch * cfg->dim * sizeof(float);
  const size_t out_elems = (size_t)cfg->max_batch * cfg->max_k;
  ix->pinned_bytes = std::max<size_t>(kIngestChunkBytes, q_bytes + out_elems * 12 + 256);

#define RS_ALLOC(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess) {                                                                             \
      int _rc = fail(RS_E_CUDA, "%s failed: %s (corpus image %.2f GB)", #expr, cudaGetErrorString(_e),   \
                     ix->corpus_bytes / 1e9);                                                            \
      free_all(ix);                                                                                      \
      delete ix;                                                                                         \
      return _rc;                                                                                        \
    }                                                                                                    \
  } while (0)

  RS_ALLOC(cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking));
  RS_ALLOC(cudaMalloc(&ix->corpus, ix->corpus_bytes));
  RS_ALLOC(cudaMemsetAsync(ix->corpus, 0, ix->corpus_bytes, ix->stream));  // pad rows of the last tile stay zero
  RS_ALLOC(cudaMalloc(&ix->dnorm, (size_t)ix->n_tiles * RS_TILE_ROWS * sizeof(float)));
  RS_ALLOC(cudaMemsetAsync(ix->dnorm, 0, (size_t)ix->n_tiles * RS_TILE_ROWS * sizeof(float), ix->stream));
  RS_ALLOC(cudaMalloc(&ix->minmax, (size_t)ix->n_tiles * sizeof(float2)));
  RS_ALLOC(cudaMalloc(&ix->qimg, qimg_bytes));
  RS_ALLOC(cudaMalloc(&ix->qnorm, pass_rows * sizeof(float)));
  RS_ALLOC(cudaMalloc(&ix->qscale, pass_rows * sizeof(float)));
  RS_ALLOC(cudaMalloc(&ix->cand_score, cand_elems * sizeof(float)));
  RS_ALLOC(cudaMalloc(&ix->cand_idx, cand_elems * sizeof(int)));
  RS_ALLOC(cudaMalloc(&ix->q_dev, q_bytes));
  // host-buffer searches write [indices int64][scores float] back to back so ONE copy brings both home
  RS_ALLOC(cudaMalloc(&ix->out_idx_dev, out_elems * (sizeof(long long) + sizeof(float))));
  RS_ALLOC(cudaMalloc(&ix->ingest_dev, kIngestChunkBytes));
  RS_ALLOC(cudaHostAlloc(&ix->pinned, ix->pinned_bytes, cudaHostAllocPortable));
  RS_ALLOC(cudaMalloc(&ix->g_best, (size_t)prop.multiProcessorCount * pass_rows * sizeof(float)));
  // doc-major pair kernel: pairs the device keeps resident (largest shared-memory configuration)
  ix->dm_pair_clusters = std::min(rs::fused_dmpair_max_clusters(kMaxSmem), prop.multiProcessorCount / 2);
  if (cfg->max_batch > 128) {
    RS_ALLOC(cudaMalloc(&ix->cbuf, (size_t)grid * pass_rows * rs::fused_wide_bufc() * sizeof(uint2)));
    ix->pair_clusters = ix->planes == 1 ? std::min(rs::fused_pair_max_clusters(cfg->max_k), prop.multiProcessorCount / 2) : 0;
  }
  RS_ALLOC(cudaMalloc(&ix->sync, 512));
  RS_ALLOC(cudaMemsetAsync(ix->sync, 0, 512, ix->stream));
  RS_ALLOC(cudaEventCreateWithFlags(&ix->order_ev, cudaEventDisableTiming));
  RS_ALLOC(cudaHostAlloc(&ix->err_flag, 64, cudaHostAllocMapped));
  *ix->err_flag = 0;
  for (int i = 0; i < 2; ++i) RS_ALLOC(cudaEventCreateWithFlags(&ix->ingest_done[i], cudaEventDisableTiming));
  for (int i = 0; i < rs_index::kEventRing; ++i) {
    RS_ALLOC(cudaEventCreate(&ix->ev_beg[i]));
    RS_ALLOC(cudaEventCreate(&ix->ev_end[i]));
  }
  RS_ALLOC(cudaStreamSynchronize(ix->stream));
#undef RS_ALLOC
  if (cfg->device < kMaxDevices) {
    g_live_indexes[cfg->device].fetch_add(1);
    g_live_epoch[cfg->device].fetch_add(1);
    ix->registered = true;
  }
  *out = ix;
  return RS_OK;
}

RS_API int rs_destroy(rs_index* ix) {
  if (!ix) return RS_OK;
  if (ix->registered) {
    g_live_indexes[ix->cfg.device].fetch_sub(1);
    g_live_epoch[ix->cfg.device].fetch_add(1);
  }
  free_all(ix);
  delete ix;
  return RS_OK;
}

namespace {
// pageable -> pinned copy spread over a few host threads (one memcpy stream tops out near 10 GB/s)
void parallel_memcpy(void* dst, const void* src, size_t bytes) {
  const int nthreads = bytes >= (8u << 20) ? 4 : 1;
  if (nthreads == 1) {
    memcpy(dst, src, bytes);
    return;
  }
  std::thread th[4];
  const size_t slice = (bytes / nthreads + 4095) & ~size_t(4095);
  for (int i = 0; i < nthreads; ++i) {
    const size_t off = std::min(bytes, (size_t)i * slice), n = std::min(slice, bytes - off);
    th[i] = std::thread([=] { if (n) memcpy(static_cast<char*>(dst) + off, static_cast<const char*>(src) + off, n); });
  }
  for (int i = 0; i < nthreads; ++i) th[i].join();
}
}  // namespace

RS_API int rs_load_rows(rs_index* ix, const void* host_rows, int32_t src_dtype, int64_t row0, int64_t nrows) {
  if (!ix || !host_rows) return fail(RS_E_INVALID, "null argument");
  if (src_dtype < RS_DTYPE_BF16 || src_dtype > RS_DTYPE_FP32) return fail(RS_E_INVALID, "bad src_dtype %d", src_dtype);
  if (row0 < 0 || nrows < 0 || row0 + nrows > ix->cfg.n_rows)
    return fail(RS_E_INVALID, "rows [%lld,%lld) outside [0,%lld)", (long long)row0, (long long)(row0 + nrows), (long long)ix->cfg.n_rows);
  std::lock_guard<std::mutex> lock(ix->mu);
  RS_CUDA(cudaSetDevice(ix->cfg.device));
  const size_t esz = src_dtype == RS_DTYPE_FP32 ? 4 : 2;
  const size_t row_bytes = (size_t)ix->cfg.dim * esz;
  const char* src = static_cast<const char*>(host_rows);
  // where do the rows live?  device memory: tile in place; pinned host memory: DMA straight from
  // it; pageable host memory: bounce through the library's pinned staging
  cudaPointerAttributes attr;
  memset(&attr, 0, sizeof(attr));
  cudaError_t pe = cudaPointerGetAttributes(&attr, host_rows);
  if (pe != cudaSuccess) {
    cudaGetLastError();
    attr.type = cudaMemoryTypeUnregistered;
  }
  if (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged) {
    const int64_t chunk_rows = 1 << 20;
    for (int64_t r = 0; r < nrows; r += chunk_rows) {
      const int64_t n = std::min(chunk_rows, nrows - r);
      RS_CUDA(rs::launch_tile_rows(src + (size_t)r * row_bytes, src_dtype, row0 + r, (int)n, ix->cfg.dim,
                                   ix->cfg.dtype, ix->cfg.scale_exp, ix->corpus, ix->dnorm, ix->stream));
      ix->stats.total_launches += 1;
    }
    RS_CUDA(cudaStreamSynchronize(ix->stream));
    ix->finalized = false;
    drop_graphs(ix);
    return RS_OK;
  }
  const bool src_pinned = attr.type == cudaMemoryTypeHost;
  // two half-size staging buffers: the host copy of chunk i+1 overlaps the DMA + tiling of chunk i
  const size_t half = kIngestChunkBytes / 2;
  const int64_t chunk_rows = std::max<int64_t>(1, (int64_t)(half / row_bytes));
  int64_t i = 0;
  for (int64_t r = 0; r < nrows; r += chunk_rows, ++i) {
    const int64_t n = std::min(chunk_rows, nrows - r);
    const int b = (int)(i & 1);
    char* pin = static_cast<char*>(ix->pinned) + b * half;
    char* dev = static_cast<char*>(ix->ingest_dev) + b * half;
    RS_CUDA(cudaEventSynchronize(ix->ingest_done[b]));  // previous user of this half has drained
    const void* dma_src = src + (size_t)r * row_bytes;
    if (!src_pinned) {
      parallel_memcpy(pin, dma_src, (size_t)n * row_bytes);
      dma_src = pin;
    }
    RS_CUDA(cudaMemcpyAsync(dev, dma_src, (size_t)n * row_bytes, cudaMemcpyHostToDevice, ix->stream));
    RS_CUDA(rs::launch_tile_rows(dev, src_dtype, row0 + r, (int)n, ix->cfg.dim, ix->cfg.dtype, ix->cfg.scale_exp,
                                 ix->corpus, ix->dnorm, ix->stream));
    RS_CUDA(cudaEventRecord(ix->ingest_done[b], ix->stream));
    ix->stats.total_launches += 1;
  }
  RS_CUDA(cudaStreamSynchronize(ix->stream));
  ix->finalized = false;
  drop_graphs(ix);
  return RS_OK;
}

RS_API int rs_fill_synthetic(rs_index* ix, uint64_t seed, int64_t global_row0) {
  if (!ix) return fail(RS_E_INVALID, "null argument");
  std::lock_guard<std::mutex> lock(ix->mu);
  RS_CUDA(cudaSetDevice(ix->cfg.device));
  const int64_t chunk = 1 << 20;
  for (int64_t r = 0; r < ix->cfg.n_rows; r += chunk) {
    const int n = (int)std::min<int64_t>(chunk, ix->cfg.n_rows - r);
    RS_CUDA(rs::launch_synth_fill(seed, global_row0 + r, r, n, ix->cfg.dim, ix->cfg.dtype, ix->cfg.scale_exp,
                                  ix->corpus, ix->dnorm, ix->stream));
    ix->stats.total_launches += 1;
  }
  RS_CUDA(cudaStreamSynchronize(ix->stream));
  ix->finalized = false;
  drop_graphs(ix);
  return RS_OK;
}

RS_API int rs_finalize(rs_index* ix) {
  if (!ix) return fail(RS_E_INVALID, "null argument");
  std::lock_guard<std::mutex> lock(ix->mu);
  RS_CUDA(cudaSetDevice(ix->cfg.device));
  RS_CUDA(rs::launch_tile_minmax(ix->dnorm, (int)ix->cfg.n_rows, ix->n_tiles, ix->minmax, ix->stream));
  RS_CUDA(cudaStreamSynchronize(ix->stream));
  ix->stats.total_launches += 1;
  ix->finalized = true;
  return RS_OK;
}

RS_API int rs_search_device(rs_index* ix, const float* q_dev, int32_t B, int32_t k, int32_t metric, int32_t query_terms,
                     int64_t* out_idx_dev, float* out_score_dev, void* stream) {
  if (!ix || !q_dev || !out_idx_dev || !out_score_dev) return fail(RS_E_INVALID, "null argument");
  std::lock_guard<std::mutex> lock(ix->mu);
  RS_CUDA(cudaSetDevice(ix->cfg.device));
  return search_device_locked(ix, q_dev, B, k, metric, query_terms, reinterpret_cast<long long*>(out_idx_dev),
                              out_score_dev, static_cast<cudaStream_t>(stream));
}

RS_API int rs_search(rs_index* ix, const float* q_host, int32_t B, int32_t k, int32_t metric, int32_t query_terms,
              int64_t* out_idx, float* out_score) {
  if (!ix || !q_host || !out_idx || !out_score) return fail(RS_E_INVALID, "null argument");
  std::lock_guard<std::mutex> lock(ix->mu);
  RS_CUDA(cudaSetDevice(ix->cfg.device));
  return search_host_locked(ix, nullptr, q_host, B, k, metric, query_terms, reinterpret_cast<long long*>(out_idx), out_score);
}

RS_API int rs_pool_mean_device(int32_t device, const void* hidden_dev, int32_t src_dtype, int32_t B, int32_t T, int32_t D,
                               float* out_q_dev, void* stream) {
  if (!hidden_dev || !out_q_dev) return fail(RS_E_INVALID, "null argument");
  if (B <= 0 || T <= 0 || D <= 0) return fail(RS_E_INVALID, "bad shape B=%d T=%d D=%d", B, T, D);
  if (src_dtype < RS_DTYPE_BF16 || src_dtype > RS_DTYPE_FP32) return fail(RS_E_INVALID, "bad dtype %d", src_dtype);
  RS_CUDA(cudaSetDevice(device));
  RS_CUDA(rs::launch_mean_pool(hidden_dev, src_dtype, B, T, D, out_q_dev, static_cast<cudaStream_t>(stream)));
  return RS_OK;
}

RS_API int rs_check(rs_index* ix) {
  if (!ix) return fail(RS_E_INVALID, "null argument");
  std::lock_guard<std::mutex> lock(ix->mu);
  RS_CUDA(cudaSetDevice(ix->cfg.device));
  return take_device_error(ix, ix->have_last ? ix->last_stream : ix->stream);
}

RS_API int rs_set_option(rs_index* ix, const char* name, int32_t value) {
  if (!ix || !name) return fail(RS_E_INVALID, "null argument");
  std::lock_guard<std::mutex> lock(ix->mu);
  for (const OptDef& d : kOptDefs)
    if (strcmp(d.name, name) == 0) {
      ix->opt.*(d.field) = value;
      RS_CUDA(cudaSetDevice(ix->cfg.device));
      drop_graphs(ix);  // captured launch sequences were planned with the old value
      return RS_OK;
    }
  return fail(RS_E_INVALID, "unknown option %s", name);
}

RS_API int rs_merge_device(int32_t device, const float* scores_dev, const int64_t* idx_dev, int32_t n_lists, int32_t B,
                    int32_t kc, int32_t k, int64_t score_list_stride, int64_t idx_list_stride,
                    int64_t* out_idx_dev, float* out_score_dev, void* stream) {
  if (!scores_dev || !idx_dev || !out_idx_dev || !out_score_dev) return fail(RS_E_INVALID, "null argument");
  if (n_lists <= 0 || n_lists > 256 || B <= 0 || kc <= 0 || k <= 0)
    return fail(RS_E_INVALID, "bad merge shape n_lists=%d B=%d kc=%d k=%d", n_lists, B, kc, k);
  RS_CUDA(cudaSetDevice(device));
  if (score_list_stride <= 0) score_list_stride = (int64_t)B * kc;
  if (idx_list_stride <= 0) idx_list_stride = (int64_t)B * kc;
  RS_CUDA(rs::launch_merge_i64(scores_dev, reinterpret_cast<const long long*>(idx_dev), n_lists, B, kc, k,
                               score_list_stride, idx_list_stride, out_score_dev, reinterpret_cast<long long*>(out_idx_dev),
                               static_cast<cudaStream_t>(stream)));
  return RS_OK;
}

RS_API int rs_xchg_create(int32_t device, int32_t world, int32_t rank, int32_t max_batch, int32_t max_k, rs_xchg** out) {
  if (!out) return fail(RS_E_INVALID, "null argument");
  *out = nullptr;
  if (world < 1 || world > rs::kMaxPeers || rank < 0 || rank >= world) return fail(RS_E_INVALID, "world=%d rank=%d (at most %d GPUs)", world, rank, rs::kMaxPeers);
  if (max_batch <= 0 || max_k <= 0) return fail(RS_E_INVALID, "bad max_batch/max_k");
  RS_CUDA(cudaSetDevice(device));
  rs_xchg* x = new (std::nothrow) rs_xchg();
  if (!x) return fail(RS_E_INVALID, "out of host memory");
  x->device = device;
  x->world = world;
  x->rank = rank;
  x->max_batch = max_batch;
  x->max_k = max_k;
  x->slot_entries = (((size_t)max_batch * max_k) + 15) & ~size_t(15);
  x->total_bytes = kXchgHeaderBytes + 2 * x->slot_entries * world * sizeof(uint4);
  cudaError_t e = cudaMalloc(&x->local, x->total_bytes);
  if (e == cudaSuccess) e = cudaMemset(x->local, 0, x->total_bytes);
  const unsigned int one = 1u;  // epochs start at 1: zeroed entries never match
  if (e == cudaSuccess) e = cudaMemcpy(x->local, &one, sizeof(one), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    cudaFree(x->local);
    delete x;
    return fail(RS_E_CUDA, "exchange buffer allocation failed: %s", cudaGetErrorString(e));
  }
  x->peer[rank] = x->local;
  *out = x;
  return RS_OK;
}

RS_API int rs_xchg_export(rs_xchg* x, void* handle64) {
  if (!x || !handle64) return fail(RS_E_INVALID, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  RS_CUDA(cudaSetDevice(x->device));
  cudaIpcMemHandle_t h;
  RS_CUDA(cudaIpcGetMemHandle(&h, x->local));
  memcpy(handle64, &h, 64);
  return RS_OK;
}

RS_API int rs_xchg_import(rs_xchg* x, int32_t peer_rank, const void* handle64) {
  if (!x || !handle64) return fail(RS_E_INVALID, "null argument");
  if (peer_rank < 0 || peer_rank >= x->world) return fail(RS_E_INVALID, "peer_rank %d of %d", peer_rank, x->world);
  if (peer_rank == x->rank) return RS_OK;
  RS_CUDA(cudaSetDevice(x->device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  RS_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  x->peer[peer_rank] = static_cast<uint8_t*>(p);
  x->opened[peer_rank] = true;
  return RS_OK;
}

RS_API int rs_xchg_attach(rs_xchg* x, rs_xchg* peer) {
  if (!x || !peer) return fail(RS_E_INVALID, "null argument");
  if (peer->world != x->world || peer->max_batch != x->max_batch || peer->max_k != x->max_k)
    return fail(RS_E_INVALID, "exchange buffers of one group must share world, max_batch and max_k");
  if (peer->rank == x->rank) return RS_OK;
  RS_CUDA(cudaSetDevice(x->device));
  if (peer->device != x->device) {
    int can = 0;
    RS_CUDA(cudaDeviceCanAccessPeer(&can, x->device, peer->device));
    if (!can) return fail(RS_E_UNSUPPORTED, "device %d cannot access device %d's memory", x->device, peer->device);
    cudaError_t e = cudaDeviceEnablePeerAccess(peer->device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
      return fail(RS_E_CUDA, "cudaDeviceEnablePeerAccess(%d) failed: %s", peer->device, cudaGetErrorString(e));
    cudaGetLastError();
  }
  x->peer[peer->rank] = peer->local;
  return RS_OK;
}

RS_API int rs_xchg_destroy(rs_xchg* x) {
  if (!x) return RS_OK;
  cudaSetDevice(x->device);
  cudaDeviceSynchronize();
  for (int g = 0; g < x->world; ++g)
    if (x->opened[g]) cudaIpcCloseMemHandle(x->peer[g]);
  cudaFree(x->local);
  delete x;
  return RS_OK;
}

RS_API int rs_search_exchange(rs_index* ix, rs_xchg* x, const float* q_dev, int32_t B, int32_t k, int32_t metric,
                              int32_t query_terms, int64_t* out_idx_dev, float* out_score_dev, void* stream) {
  if (!ix || !x || !q_dev || !out_idx_dev || !out_score_dev) return fail(RS_E_INVALID, "null argument");
  if (x->device != ix->cfg.device) return fail(RS_E_INVALID, "index and exchange buffer live on different devices");
  std::lock_guard<std::mutex> lock(ix->mu);
  RS_CUDA(cudaSetDevice(ix->cfg.device));
  return search_device_locked(ix, q_dev, B, k, metric, query_terms, reinterpret_cast<long long*>(out_idx_dev),
                              out_score_dev, static_cast<cudaStream_t>(stream), x);
}

RS_API int rs_search_exchange_host(rs_index* ix, rs_xchg* x, const float* q_host, int32_t B, int32_t k, int32_t metric,
                                   int32_t query_terms, int64_t* out_idx, float* out_score) {
  if (!ix || !x || !q_host || (out_idx == nullptr) != (out_score == nullptr)) return fail(RS_E_INVALID, "null argument");
  if (x->device != ix->cfg.device) return fail(RS_E_INVALID, "index and exchange buffer live on different devices");
  std::lock_guard<std::mutex> lock(ix->mu);
  RS_CUDA(cudaSetDevice(ix->cfg.device));
  return search_host_locked(ix, x, q_host, B, k, metric, query_terms, reinterpret_cast<long long*>(out_idx), out_score);
}

// ---------------------------------------------------------------------------------------------------------
// Single-process multi-GPU: what a service that constructs ONE retriever in ONE process gets (main.py:52-55).
// The group owns one exchange buffer per shard (peer access inside the process).  Per search: one pinned copy of the
// queries, then every shard's launch sequence - H2D of the queries + ONE fused kernel (candidate exchange and final
// merge inside the kernels, over NVLink), + the D2H of the result on shard 0 - is issued by the host thread that drives
// that shard: the caller for shard 0, one worker thread per further shard, so the n launches go out in parallel
// (issued from one thread, as one multi-device CUDA graph, they cost ~60 us more per call at 8 GPUs).  Each sequence is
// replayed from its own single-device CUDA graph from the second call of a shape on.  The host waits on an event
// recorded behind shard 0's D2H: shard 0 has every peer's candidates by then, so every peer has long read its queries.
// Workers spin for a short while after a search (back-to-back calls find them awake), then sleep on a condition variable.
// ---------------------------------------------------------------------------------------------------------
struct rs_group {
  int n = 0;
  rs_index* ix[rs::kMaxPeers] = {};
  rs_xchg* xg[rs::kMaxPeers] = {};
  cudaEvent_t done = nullptr;  // recorded right behind the D2H of the result on shard 0's stream: what the host waits for
  struct GraphEntry {
    int B, k, metric, terms, uses;
    cudaGraphExec_t exec;
  };
  std::vector<GraphEntry> graphs[rs::kMaxPeers];  // per shard, touched only by the thread that drives the shard
  unsigned live_epoch_seen[rs::kMaxPeers] = {};
  int use_graph = 1;
  std::mutex mu;  // one search at a time
  // the job of the current search (written by the caller before `gen` moves) and the workers' state
  struct Job {
    size_t q_bytes, n_out;
    int B, k, metric, terms;
  } job = {};
  std::thread workers[rs::kMaxPeers];
  std::atomic<unsigned long long> gen{0};
  std::atomic<int> pending{0};
  std::atomic<int> sleepers{0};
  std::atomic<bool> stop{false};
  std::mutex wmu;
  std::condition_variable wcv;
  int rc[rs::kMaxPeers] = {};
  char err[rs::kMaxPeers][256] = {};
  int spin_us = 500;
};

namespace {

// One shard's launch sequence of the current job, issued by the thread that drives the shard.
int group_launch_shard(rs_group* gr, int g) {
  rs_index* ix = gr->ix[g];
  rs_index* ix0 = gr->ix[0];
  const rs_group::Job& jb = gr->job;
  std::lock_guard<std::mutex> lock(ix->mu);
  RS_CUDA(cudaSetDevice(ix->cfg.device));
  int rc = take_device_error(ix, ix->stream);
  if (rc == RS_OK) rc = order_after_previous(ix, ix->stream);
  if (rc != RS_OK) return rc;
  char* pin = static_cast<char*>(ix0->pinned);  // portable pinned memory: every GPU copies the queries from it
  long long* pin_idx = reinterpret_cast<long long*>(pin + ((jb.q_bytes + 255) & ~size_t(255)));
  auto enqueue = [&]() -> int {
    RS_CUDA(cudaMemcpyAsync(ix->q_dev, pin, jb.q_bytes, cudaMemcpyHostToDevice, ix->stream));
    float* score_dev = reinterpret_cast<float*>(ix->out_idx_dev + jb.n_out);
    int r = search_device_locked(ix, ix->q_dev, jb.B, jb.k, jb.metric, jb.terms, ix->out_idx_dev, score_dev, ix->stream,
                                 gr->n > 1 ? gr->xg[g] : nullptr);
    if (r != RS_OK) return r;
    if (g == 0) {
      RS_CUDA(cudaMemcpyAsync(pin_idx, ix->out_idx_dev, jb.n_out * (sizeof(long long) + sizeof(float)), cudaMemcpyDeviceToHost,
                              ix->stream));
      if (ix->capturing) RS_CUDA(cudaEventRecordWithFlags(gr->done, ix->stream, cudaEventRecordExternal));
      else RS_CUDA(cudaEventRecord(gr->done, ix->stream));
    }
    return RS_OK;
  };
  auto& graphs = gr->graphs[g];
  if (ix->registered) {  // an index was created or destroyed on this device: the captured tail may be the wrong kind
    const unsigned ep = g_live_epoch[ix->cfg.device].load();
    if (ep != gr->live_epoch_seen[g]) {
      for (auto& e : graphs)
        if (e.exec) cudaGraphExecDestroy(e.exec);
      graphs.clear();
      gr->live_epoch_seen[g] = ep;
    }
  }
  rs_group::GraphEntry* ge = nullptr;
  if (gr->use_graph) {
    for (auto& e : graphs)
      if (e.B == jb.B && e.k == jb.k && e.metric == jb.metric && e.terms == jb.terms) ge = &e;
    if (!ge && graphs.size() < 64) {
      graphs.push_back({jb.B, jb.k, jb.metric, jb.terms, 0, nullptr});
      ge = &graphs.back();
    }
  }
  if (ge && !ge->exec && ge->uses >= 1) {  // second call with this shape: capture (the first, eager call validated everything)
    ix->capturing = true;
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaStreamBeginCapture(ix->stream, cudaStreamCaptureModeThreadLocal);
    int crc = ce == cudaSuccess ? enqueue() : RS_E_CUDA;
    cudaError_t ee = cudaStreamEndCapture(ix->stream, &graph);
    ix->capturing = false;
    if (crc == RS_OK && ee == cudaSuccess && graph && cudaGraphInstantiate(&ge->exec, graph, 0) != cudaSuccess) ge->exec = nullptr;
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    if (!ge->exec) ge->uses = -1000000;
  }
  if (ge && ge->exec) {
    RS_CUDA(cudaGraphLaunch(ge->exec, ix->stream));
    ix->stats.total_launches += ix->stats.launches;
  } else {
    rc = enqueue();
    if (rc != RS_OK) return rc;
    if (ge) ++ge->uses;
  }
  return RS_OK;
}

void group_worker(rs_group* gr, int g) {
  unsigned long long seen = 0;
  while (true) {
    // wait for the next job: spin briefly (back-to-back searches), then sleep
    const auto t0 = std::chrono::steady_clock::now();
    unsigned spins = 0;
    while (gr->gen.load(std::memory_order_acquire) == seen && !gr->stop.load(std::memory_order_acquire)) {
#if defined(__x86_64__) || defined(__i386__)
      __builtin_ia32_pause();
#endif
      if ((++spins & 63u) == 0u &&
          std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(gr->spin_us)) {
        std::unique_lock<std::mutex> lk(gr->wmu);
        gr->sleepers.fetch_add(1);
        gr->wcv.wait(lk, [&] { return gr->gen.load(std::memory_order_acquire) != seen || gr->stop.load(std::memory_order_acquire); });
        gr->sleepers.fetch_sub(1);
      }
    }
    if (gr->stop.load(std::memory_order_acquire)) return;
    seen = gr->gen.load(std::memory_order_acquire);
    const int rc = group_launch_shard(gr, g);
    gr->rc[g] = rc;
    if (rc != RS_OK) snprintf(gr->err[g], sizeof(gr->err[g]), "%.250s", g_err);
    gr->pending.fetch_sub(1, std::memory_order_release);
  }
}

}  // namespace

RS_API int rs_group_destroy(rs_group* gr);

RS_API int rs_group_create(rs_index** shards, int32_t n, rs_group** out) {
  if (!shards || !out) return fail(RS_E_INVALID, "null argument");
  *out = nullptr;
  if (n < 1 || n > rs::kMaxPeers) return fail(RS_E_INVALID, "a group holds 1..%d shards", rs::kMaxPeers);
  for (int g = 0; g < n; ++g) {
    if (!shards[g]) return fail(RS_E_INVALID, "null shard");
    if (shards[g]->cfg.dim != shards[0]->cfg.dim || shards[g]->cfg.dtype != shards[0]->cfg.dtype ||
        shards[g]->cfg.max_batch != shards[0]->cfg.max_batch || shards[g]->cfg.max_k != shards[0]->cfg.max_k)
      return fail(RS_E_INVALID, "shards of one group must share dim, dtype, max_batch and max_k");
    for (int h = 0; h < g; ++h)  // the in-kernel exchange needs every shard's grid resident at the same time
      if (shards[h]->cfg.device == shards[g]->cfg.device) return fail(RS_E_INVALID, "two shards on device %d", shards[g]->cfg.device);
  }
  rs_group* gr = new (std::nothrow) rs_group();
  if (!gr) return fail(RS_E_INVALID, "out of host memory");
  gr->n = n;
  gr->use_graph = shards[0]->opt.graph;
  gr->spin_us = env_int("RS_GROUP_SPIN_US", 500);  // workers stay awake across back-to-back searches of up to half a millisecond
  int rc = RS_OK;
  for (int g = 0; g < n && rc == RS_OK; ++g) {
    gr->ix[g] = shards[g];
    rc = rs_xchg_create(shards[g]->cfg.device, n, g, shards[g]->cfg.max_batch, shards[g]->cfg.max_k, &gr->xg[g]);
  }
  for (int g = 0; g < n && rc == RS_OK; ++g)
    for (int h = 0; h < n && rc == RS_OK; ++h) rc = rs_xchg_attach(gr->xg[g], gr->xg[h]);
  if (rc == RS_OK) {
    cudaSetDevice(shards[0]->cfg.device);
    if (cudaEventCreateWithFlags(&gr->done, cudaEventDisableTiming) != cudaSuccess) rc = fail(RS_E_CUDA, "event creation failed");
  }
  if (rc != RS_OK) {
    rs_group_destroy(gr);
    return rc;
  }
  try {
    for (int g = 1; g < n; ++g) gr->workers[g] = std::thread(group_worker, gr, g);
  } catch (...) {
    rs_group_destroy(gr);
    return fail(RS_E_INVALID, "could not start the group's worker threads");
  }
  *out = gr;
  return RS_OK;
}

RS_API int rs_group_destroy(rs_group* gr) {
  if (!gr) return RS_OK;
  {
    std::lock_guard<std::mutex> lk(gr->wmu);
    gr->stop.store(true, std::memory_order_release);
  }
  gr->wcv.notify_all();
  for (int g = 1; g < gr->n; ++g)
    if (gr->workers[g].joinable()) gr->workers[g].join();
  for (int g = 0; g < gr->n; ++g) {
    if (gr->ix[g]) {
      cudaSetDevice(gr->ix[g]->cfg.device);
      cudaDeviceSynchronize();
    }
    for (auto& e : gr->graphs[g])
      if (e.exec) cudaGraphExecDestroy(e.exec);
  }
  for (int g = 0; g < gr->n; ++g) rs_xchg_destroy(gr->xg[g]);
  if (gr->done) cudaEventDestroy(gr->done);
  delete gr;
  return RS_OK;
}

RS_API int rs_group_search(rs_group* gr, const float* q_host, int32_t B, int32_t k, int32_t metric, int32_t query_terms,
                           int64_t* out_idx, float* out_score) {
  if (!gr || !q_host || !out_idx || !out_score) return fail(RS_E_INVALID, "null argument");
  std::lock_guard<std::mutex> glock(gr->mu);
  rs_index* ix0 = gr->ix[0];
  if (B <= 0 || B > ix0->cfg.max_batch) return fail(RS_E_INVALID, "B=%d outside [1,%d]", B, ix0->cfg.max_batch);
  if (k <= 0 || k > ix0->cfg.max_k) return fail(RS_E_INVALID, "k=%d outside [1,%d]", k, ix0->cfg.max_k);
  if (metric != RS_METRIC_DOT && metric != RS_METRIC_COSINE) return fail(RS_E_INVALID, "bad metric %d", metric);
  for (int g = 0; g < gr->n; ++g)
    if (!gr->ix[g]->finalized) return fail(RS_E_STATE, "rs_finalize has not been called on shard %d", g);
  if (ix0->planes == 2) query_terms = 2;
  if (query_terms != 1 && query_terms != 2) return fail(RS_E_INVALID, "query_terms must be 1 or 2");
  const size_t q_bytes = (size_t)B * ix0->cfg.dim * sizeof(float);
  const size_t n_out = (size_t)B * k;
  char* pin = static_cast<char*>(ix0->pinned);
  long long* pin_idx = reinterpret_cast<long long*>(pin + ((q_bytes + 255) & ~size_t(255)));
  float* pin_score = reinterpret_cast<float*>(pin_idx + n_out);
  memcpy(pin, q_host, q_bytes);
  gr->job = {q_bytes, n_out, B, k, metric, query_terms};
  // publish the job: the workers issue their shards' launches while this thread issues shard 0's
  gr->pending.store(gr->n - 1, std::memory_order_relaxed);
  {
    // (the lock orders the generation bump against a worker that is just going to sleep)
    std::lock_guard<std::mutex> lk(gr->wmu);
    gr->gen.fetch_add(1, std::memory_order_release);
  }
  if (gr->sleepers.load() > 0) gr->wcv.notify_all();
  gr->rc[0] = group_launch_shard(gr, 0);
  if (gr->rc[0] != RS_OK) snprintf(gr->err[0], sizeof(gr->err[0]), "%.250s", g_err);
  while (gr->pending.load(std::memory_order_acquire) != 0) {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#endif
  }
  for (int g = 0; g < gr->n; ++g)
    if (gr->rc[g] != RS_OK) {
      // some shard's launch failed: drain what was issued (the kernels that did start give up on their own after
      // the peer time-out), then report
      for (int h = 0; h < gr->n; ++h) {
        cudaSetDevice(gr->ix[h]->cfg.device);
        cudaStreamSynchronize(gr->ix[h]->stream);
      }
      return fail(gr->rc[g], "shard %d: %s", g, gr->err[g]);
    }
  RS_CUDA(cudaSetDevice(ix0->cfg.device));
  RS_CUDA(cudaEventSynchronize(gr->done));
  {  // shard 0 waited for every peer inside its kernel: a peer that never showed up raised shard 0's error word.  The
     // peers' own error words are read when their next launch is issued (their kernels may still be retiring).
    std::lock_guard<std::mutex> lock(ix0->mu);
    int rc = take_device_error(ix0, ix0->stream);
    if (rc != RS_OK) return rc;
  }
  memcpy(out_idx, pin_idx, n_out * sizeof(long long));
  memcpy(out_score, pin_score, n_out * sizeof(float));
  return RS_OK;
}

RS_API int rs_get_rows(rs_index* ix, int64_t row0, int64_t nrows, float* out_host) {
  if (!ix || !out_host) return fail(RS_E_INVALID, "null argument");
  if (row0 < 0 || nrows < 0 || row0 + nrows > ix->cfg.n_rows) return fail(RS_E_INVALID, "row range out of bounds");
  std::lock_guard<std::mutex> lock(ix->mu);
  RS_CUDA(cudaSetDevice(ix->cfg.device));
  const size_t row_bytes = (size_t)ix->cfg.dim * sizeof(float);
  const int64_t chunk_rows = std::max<int64_t>(1, kIngestChunkBytes / (int64_t)row_bytes);
  for (int64_t r = 0; r < nrows; r += chunk_rows) {
    const int64_t n = std::min(chunk_rows, nrows - r);
    RS_CUDA(rs::launch_untile_rows(ix->corpus, row0 + r, (int)n, ix->cfg.dim, ix->cfg.dtype, ix->cfg.scale_exp,
                                   static_cast<float*>(ix->ingest_dev), ix->stream));
    RS_CUDA(cudaMemcpyAsync(ix->pinned, ix->ingest_dev, (size_t)n * row_bytes, cudaMemcpyDeviceToHost, ix->stream));
    RS_CUDA(cudaStreamSynchronize(ix->stream));
    memcpy(reinterpret_cast<char*>(out_host) + (size_t)r * row_bytes, ix->pinned, (size_t)n * row_bytes);
    ix->stats.total_launches += 1;
  }
  return RS_OK;
}

RS_API int rs_get_rows_device(rs_index* ix, int64_t row0, int64_t nrows, float* out_dev, void* stream) {
  if (!ix || !out_dev) return fail(RS_E_INVALID, "null argument");
  if (row0 < 0 || nrows < 0 || row0 + nrows > ix->cfg.n_rows) return fail(RS_E_INVALID, "row range out of bounds");
  std::lock_guard<std::mutex> lock(ix->mu);
  RS_CUDA(cudaSetDevice(ix->cfg.device));
  const int64_t chunk_rows = 1 << 20;
  for (int64_t r = 0; r < nrows; r += chunk_rows) {
    const int64_t n = std::min(chunk_rows, nrows - r);
    RS_CUDA(rs::launch_untile_rows(ix->corpus, row0 + r, (int)n, ix->cfg.dim, ix->cfg.dtype, ix->cfg.scale_exp,
                                   out_dev + (size_t)r * ix->cfg.dim, static_cast<cudaStream_t>(stream)));
    ix->stats.total_launches += 1;
  }
  return RS_OK;
}

RS_API int rs_get_norms(rs_index* ix, int64_t row0, int64_t nrows, float* out_host) {
  if (!ix || !out_host) return fail(RS_E_INVALID, "null argument");
  if (row0 < 0 || nrows < 0 || row0 + nrows > ix->cfg.n_rows) return fail(RS_E_INVALID, "row range out of bounds");
  std::lock_guard<std::mutex> lock(ix->mu);
  RS_CUDA(cudaSetDevice(ix->cfg.device));
  RS_CUDA(cudaMemcpy(out_host, ix->dnorm + row0, (size_t)nrows * sizeof(float), cudaMemcpyDeviceToHost));
  return RS_OK;
}

RS_API int rs_get_image(rs_index* ix, int64_t off, int64_t nbytes, void* out_host) {
  if (!ix || !out_host) return fail(RS_E_INVALID, "null argument");
  if (off < 0 || nbytes < 0 || (size_t)(off + nbytes) > ix->corpus_bytes) return fail(RS_E_INVALID, "image range out of bounds");
  std::lock_guard<std::mutex> lock(ix->mu);
  RS_CUDA(cudaSetDevice(ix->cfg.device));
  RS_CUDA(cudaMemcpy(out_host, ix->corpus + off, (size_t)nbytes, cudaMemcpyDeviceToHost));
  return RS_OK;
}

RS_API int rs_get_kernel_times(rs_index* ix, float* out_ms, int32_t n) {
  if (!ix || !out_ms || n <= 0) return fail(RS_E_INVALID, "bad argument");
  std::lock_guard<std::mutex> lock(ix->mu);
  RS_CUDA(cudaSetDevice(ix->cfg.device));
  if (n > rs_index::kEventRing || n > ix->ev_count)
    return fail(RS_E_INVALID, "only %lld fused launches recorded (ring of %d)", ix->ev_count, rs_index::kEventRing);
  for (int i = 0; i < n; ++i) {
    const int slot = (int)((ix->ev_count - n + i) % rs_index::kEventRing);
    RS_CUDA(cudaEventSynchronize(ix->ev_end[slot]));
    RS_CUDA(cudaEventElapsedTime(&out_ms[i], ix->ev_beg[slot], ix->ev_end[slot]));
  }
  return RS_OK;
}

RS_API int rs_debug_timestamps(rs_index* ix, uint64_t* out, int32_t n_ctas) {
  if (!ix || !out || n_ctas <= 0 || n_ctas > ix->sm_count) return fail(RS_E_INVALID, "bad argument");
  std::lock_guard<std::mutex> lock(ix->mu);
  RS_CUDA(cudaSetDevice(ix->cfg.device));
  if (!ix->debug_ts) return fail(RS_E_STATE, "no stamps recorded: run a search with RS_DEBUG_TS=1 first");
  RS_CUDA(cudaDeviceSynchronize());
  RS_CUDA(cudaMemcpy(out, ix->debug_ts, (size_t)n_ctas * 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  return RS_OK;
}

// ---------------------------------------------------------------- document arena (host side: SURVEY §8f.3)
struct rs_docs {
  std::vector<char> arena;
  std::vector<int64_t> off;  // n + 1
};

RS_API int rs_docs_create(const char* arena, const int64_t* offsets, int64_t n_docs, rs_docs** out) {
  if (!offsets || !out || n_docs < 0 || (n_docs > 0 && offsets[n_docs] > 0 && !arena)) return fail(RS_E_INVALID, "null argument");
  *out = nullptr;
  if (offsets[0] != 0) return fail(RS_E_INVALID, "offsets[0] must be 0");
  for (int64_t i = 0; i < n_docs; ++i)
    if (offsets[i + 1] < offsets[i]) return fail(RS_E_INVALID, "offsets must not decrease (document %lld)", (long long)i);
  rs_docs* d = new (std::nothrow) rs_docs;
  if (!d) return fail(RS_E_INVALID, "out of memory");
  try {
    d->off.assign(offsets, offsets + n_docs + 1);
    d->arena.assign(arena, arena + offsets[n_docs]);
  } catch (...) {
    delete d;
    return fail(RS_E_INVALID, "out of memory");
  }
  *out = d;
  return RS_OK;
}

RS_API int rs_docs_destroy(rs_docs* d) {
  delete d;
  return RS_OK;
}

RS_API int rs_docs_join(const rs_docs* d, const int64_t* idx, int32_t B, int32_t stride, const int32_t* ks, const char* sep,
                        int32_t sep_len, char* out, int64_t out_cap, int64_t* out_offsets, int64_t* need) {
  if (!d || !idx || !ks || !out_offsets || !need || B < 0 || stride < 0 || sep_len < 0 || (sep_len > 0 && !sep))
    return fail(RS_E_INVALID, "bad argument");
  const int64_t n = (int64_t)d->off.size() - 1;
  // pass 1: sizes (and validation); pass 2: copy, only if the caller's buffer is large enough
  int64_t total = 0;
  out_offsets[0] = 0;
  for (int32_t i = 0; i < B; ++i) {
    const int32_t ki = std::min(ks[i], stride);
    int64_t bytes = 0;
    int32_t hits = 0;
    for (int32_t j = 0; j < stride && hits < ki; ++j) {
      const int64_t r = idx[(size_t)i * stride + j];
      if (r < 0) continue;  // padding (-1) of rows with fewer than k valid documents
      if (r >= n) return fail(RS_E_INVALID, "index %lld of query %d outside the %lld documents", (long long)r, i, (long long)n);
      bytes += (hits ? sep_len : 0) + (d->off[r + 1] - d->off[r]);
      ++hits;
    }
    total += bytes;
    out_offsets[i + 1] = total;
  }
  *need = total;
  if (!out || out_cap < total) return RS_OK;  // sizes only: call again with a buffer of *need bytes
  for (int32_t i = 0; i < B; ++i) {
    const int32_t ki = std::min(ks[i], stride);
    char* w = out + out_offsets[i];
    int32_t hits = 0;
    for (int32_t j = 0; j < stride && hits < ki; ++j) {
      const int64_t r = idx[(size_t)i * stride + j];
      if (r < 0) continue;
      if (hits) {
        memcpy(w, sep, (size_t)sep_len);
        w += sep_len;
      }
      const int64_t len = d->off[r + 1] - d->off[r];
      memcpy(w, d->arena.data() + d->off[r], (size_t)len);
      w += len;
      ++hits;
    }
  }
  return RS_OK;
}

RS_API int rs_debug_plan(int32_t dim, int32_t dtype, int32_t B, int32_t k, int32_t query_terms, int32_t* out9) {
  if (!out9 || dim <= 0 || dim % 64 != 0) return fail(RS_E_INVALID, "bad argument");
  rs_index tmp;
  memset(&tmp.cfg, 0, sizeof(tmp.cfg));
  tmp.cfg.dim = dim;
  tmp.kc = dim / RS_KCHUNK_ELEMS;
  tmp.planes = dtype == RS_DTYPE_FP32 ? 2 : 1;
  if (tmp.planes == 2) query_terms = 2;
  Plan pl;
  int rc = make_plan(&tmp, options_from_env(), B, k, query_terms, &pl);
  if (rc != RS_OK) return rc;
  const int v[9] = {pl.passes, pl.per_pass, pl.bpad, pl.stages, pl.q_resident, pl.sub, pl.nbuf, pl.mma_m, pl.cps};
  memcpy(out9, v, sizeof(v));
  return RS_OK;
}

RS_API int rs_get_stats(rs_index* ix, rs_stats* out) {
  if (!ix || !out) return fail(RS_E_INVALID, "null argument");
  if (out->struct_size != (int32_t)sizeof(rs_stats)) return fail(RS_E_INVALID, "rs_stats.struct_size mismatch");
  std::lock_guard<std::mutex> lock(ix->mu);
  *out = ix->stats;
  return RS_OK;
}

}  // extern "C"
