// vgpu.cu — implementation of the C ABI declared in include/vgpu.h (libvgpu.so).
//
// Host side of the B200-native scan -> filter -> group-by-aggregate path: the HBM column store,
// the planner that lowers a vgpu_plan (data, not code) to ScanParams, segment pruning with the
// reference's exact rule (src/codegen/query/filter.cc:263-335), kernel launches, the optional
// NCCL merge of per-GPU partial group tables and the extraction of the group table.
// There is NO CPU implementation of the scan in this library: without a CUDA device every entry
// point fails with VGPU_ERR_CUDA.
#include "../../include/vgpu.h"
#include "select_kernels.cuh"
#include "time_dict.h"
#include "planner.h"

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <memory>
#include <atomic>
#include <mutex>
#include <nccl.h>
#include <shared_mutex>
#include <string>
#include <vector>

using namespace vgpu;

namespace {

thread_local std::string g_err;

#define CUDA_CK(x)                                                                          \
  do {                                                                                      \
    cudaError_t e_ = (x);                                                                   \
    if (e_ != cudaSuccess)                                                                  \
      fail(e_ == cudaErrorMemoryAllocation ? VGPU_ERR_NOMEM : VGPU_ERR_CUDA,                \
           std::string(#x) + ": " + cudaGetErrorString(e_));                                \
  } while (0)

template <class F> int guard(F &&f) {
  try {
    f();
    return VGPU_OK;
  } catch (const Err &e) {
    g_err = e.msg;
    return e.code;
  } catch (const std::bad_alloc &) {
    g_err = "host allocation failed";
    return VGPU_ERR_NOMEM;
  } catch (const std::exception &e) {
    g_err = e.what();
    return VGPU_ERR_INVALID;
  }
}

uint64_t round_up(uint64_t v, uint64_t m) { return (v + m - 1) / m * m; }
uint64_t pow2_ceil(uint64_t v) {
  uint64_t p = 1;
  while (p < v) p <<= 1;
  return p;
}

// ---------------------------------------------------------------------------------------------
// NCCL through dlopen: single-GPU users never need the library
// ---------------------------------------------------------------------------------------------
struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

void nccl_load() {
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  if (g_nccl.lib) return;
  // RTLD_NOLOAD first: reuse the copy torch already mapped (same soname), else load the system one.
  void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) fail(VGPU_ERR_NCCL, std::string("cannot load libnccl.so.2: ") + dlerror());
#define NSYM(field, name)                                                  \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(lib, name)); \
  if (!g_nccl.field) fail(VGPU_ERR_NCCL, std::string("missing NCCL symbol ") + name);
  NSYM(GetUniqueId, "ncclGetUniqueId");
  NSYM(CommInitRank, "ncclCommInitRank");
  NSYM(CommDestroy, "ncclCommDestroy");
  NSYM(AllReduce, "ncclAllReduce");
  NSYM(Broadcast, "ncclBroadcast");
  NSYM(AllGather, "ncclAllGather");
  NSYM(Send, "ncclSend");
  NSYM(Recv, "ncclRecv");
  NSYM(GroupStart, "ncclGroupStart");
  NSYM(GroupEnd, "ncclGroupEnd");
  NSYM(GetErrorString, "ncclGetErrorString");
#undef NSYM
  g_nccl.lib = lib;
}
#define NCCL_CK(x)                                                                           \
  do {                                                                                       \
    ncclResult_t r_ = (x);                                                                   \
    if (r_ != ncclSuccess)                                                                   \
      fail(VGPU_ERR_NCCL, std::string(#x) + ": " + g_nccl.GetErrorString(r_));               \
  } while (0)

}  // namespace

// ---------------------------------------------------------------------------------------------
// objects behind the opaque handles
// ---------------------------------------------------------------------------------------------
// Pinned host blocks behind vgpu_result. Reference-counted: a result may be freed after
// vgpu_shutdown() of its context.
struct PinnedPool {
  std::mutex mu;
  std::vector<std::pair<void *, size_t>> free_blocks;
  ~PinnedPool() {
    for (auto &b : free_blocks) cudaFreeHost(b.first);
  }
  std::pair<void *, size_t> acquire(size_t bytes) {
    {
      std::lock_guard<std::mutex> lk(mu);
      size_t best = free_blocks.size();
      for (size_t i = 0; i < free_blocks.size(); ++i)
        if (free_blocks[i].second >= bytes && (best == free_blocks.size() || free_blocks[i].second < free_blocks[best].second))
          best = i;
      if (best != free_blocks.size()) {
        auto blk = free_blocks[best];
        free_blocks.erase(free_blocks.begin() + best);
        return blk;
      }
    }
    size_t cap = 1 << 16;
    while (cap < bytes) cap <<= 1;
    void *p = nullptr;
    if (cudaMallocHost(&p, cap) != cudaSuccess) throw std::bad_alloc();
    return {p, cap};
  }
  void release(std::pair<void *, size_t> blk) {
    std::lock_guard<std::mutex> lk(mu);
    if (free_blocks.size() >= 8) {  // keep the pool small: drop the smallest block
      size_t worst = 0;
      for (size_t i = 1; i < free_blocks.size(); ++i)
        if (free_blocks[i].second < free_blocks[worst].second) worst = i;
      if (free_blocks[worst].second < blk.second) std::swap(free_blocks[worst], blk);
      cudaFreeHost(blk.first);
      return;
    }
    free_blocks.push_back(blk);
  }
};

// Everything one in-flight query needs of its own: vgpu_query_* is re-entrant per context (the reference's generated
// query function keeps all its state on the stack and runs on `query_threads` pool threads at once,
// src/db/database.cc:28-33), so no two queries share a stream, an event or a counter block.
struct QueryScope {
  cudaStream_t s0 = nullptr, s1 = nullptr;   // main stream; side stream (early group extraction + its copies)
  cudaEvent_t ev_begin = nullptr, ev_scan0 = nullptr, ev_scan1 = nullptr, ev_end = nullptr;
  cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_c = nullptr;  // untimed ordering events between s0 and s1
  cudaEvent_t ev_phase[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // VGPU_TRACE: phase boundaries on s0
  unsigned long long *d_counters = nullptr;  // 16 x u64, layout in query_agg.inl (kC*)
  unsigned long long *h_counters = nullptr;  // pinned: [0,16) read-back, [16,32) initial image, [32,48) second read-back
  uint64_t *d_plan = nullptr;                // 64 x u64: plan-time agreement between ranks
  uint64_t *h_plan = nullptr;                // pinned
};

struct vgpu_ctx {
  int device = 0;
  int sm_count = 148;
  // column-store maintenance (put / generate / invalidate) runs on these; one such call at a time per context
  cudaStream_t stream = nullptr;
  // host -> device copies of vgpu_segment_put run on their own stream: the DMA of one segment overlaps the
  // statistics and row-mirror kernels of the previous one (which stay on `stream`, ordered by ev_copy)
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_copy = nullptr;
  std::mutex put_mu;
  // the caller's stream (vgpu_set_stream): queries are ordered after what it holds when they start, and it waits
  // for them when they end, so that the caller's own events bracket a query
  cudaStream_t user_stream = nullptr;
  // idle query scopes
  std::mutex scope_mu;
  std::vector<QueryScope *> idle_scopes;
  // multi-GPU: one collective sequence at a time per communicator (every rank must issue the same sequence)
  std::mutex comm_mu;
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  // L2 persistence (group tables are pinned in L2 while the columns stream through it)
  uint64_t l2_persist_bytes = 0, l2_window_max = 0;
  // VGPU_TUNE: bit 0 pin the group table in L2 (off: measured slower), bit 1 evict_first column streams (on),
  // bit 2 evict_last group table, bit 5 no next-chunk L2 prefetch, bit 6 build no row-major mirror,
  // bit 7 never gather from the mirror, bit 8 always gather from the mirror (tests), bit 12 no unrolled conjunction fast path
  // (always the stack interpreter), bit 13 per-lane instead of bulk L2 prefetch of the next chunk, bit 14 no
  // tightening of key domains from the predicate, bit 18 no CTA-private shared-memory copy of small dense group tables,
  // bit 19 count-distinct: never the shared-memory-set fast path, bit 20 no early group extraction on the side
  // stream, bit 21 no bucket dictionary for rolled-up time keys, bit 22 always store presence flags (never read them
  // off a COUNT accumulator), bit 23 no post-aggregation on the device (HAVING / top-N stay the host's)
  uint32_t tune = 2;
  int ctas_per_sm = VGPU_MIN_CTAS;  // resident scan CTAs per SM (compile time: with kScanThreads it fixes the register cap)
  uint32_t unit_chunks = 0;  // VGPU_UNIT_CHUNKS: 512-row chunks per dynamically scheduled work unit (0: adaptive)
  // test hooks that force the rarely taken branches at test sizes (tests/test_gpu_forced_paths.py):
  uint64_t test_pairs_cap = 0;     // VGPU_TEST_PAIRS_CAP: first-attempt capacity of the count-distinct pair regions (overflow + regrow)
  uint64_t test_hash_cap = 0;      // VGPU_TEST_HASH_CAP: first-attempt capacity of hashed group tables (x4 regrow)
  uint64_t test_bucket_pairs = 0;  // VGPU_TEST_BUCKET_PAIRS: pairs per L2-sized partition of the general dedupe path (B > 1)
  uint64_t test_small_pairs = 0;   // VGPU_TEST_SMALL_PAIRS: largest pair capacity deduplicated by one global set
  uint32_t test_set_slots = 0;     // VGPU_TEST_SET_SLOTS: slots of the shared-memory sets
  uint64_t test_expect = 0;        // "expect_pairs": pairs the fast path sizes its buckets for (too few: overflow + fallback)
  // pool of pinned host blocks that back vgpu_result (D2H at full PCIe speed, no per-query
  // cudaMallocHost); shared with the results so that they may outlive the context
  std::shared_ptr<PinnedPool> pool = std::make_shared<PinnedPool>();
  bool trace = false;
};

namespace {

struct SegmentData {
  uint8_t *slab = nullptr;
  uint64_t cap = 0;
  uint64_t nrows = 0;
  bool valid = false;
  uint32_t *bs_values[kMaxBitsetCols] = {nullptr, nullptr, nullptr, nullptr};
  uint32_t *bs_offsets[kMaxBitsetCols] = {nullptr, nullptr, nullptr, nullptr};
  uint64_t bs_n[kMaxBitsetCols] = {0, 0, 0, 0};
  uint64_t bs_vcap[kMaxBitsetCols] = {0, 0, 0, 0};  // allocated ids / offsets (elements)
  uint64_t bs_ocap[kMaxBitsetCols] = {0, 0, 0, 0};
  bool bs_has_offsets[kMaxBitsetCols] = {false, false, false, false};  // CSR offsets in use (else one id per row)
  uint64_t hi_rows = 0;        // rows [hi_rows, cap) of every column are known to be zero (no tail memsets on re-put)
  uint8_t *rows = nullptr;     // row-major mirror of every column, or nullptr
  uint64_t rows_cap = 0;       // rows the mirror was allocated for
  bool stats_pending = false;  // min/max computed on the device, not yet read back
  // ordered (to_ordered) min / max of the stored values per column; only fixed-width dimensions
  std::vector<uint64_t> omin, omax;
};

}  // namespace

struct vgpu_table {
  vgpu_ctx *ctx = nullptr;
  // queries hold it shared from planning to their last device operation; put / generate / invalidate hold it
  // exclusively (same-table writers wait for running queries, queries on it wait for the writer)
  std::shared_mutex mu;
  cudaEvent_t ev_put = nullptr;   // recorded after the device work of the last put: queries' streams wait for it
  std::vector<ColInfo> cols;
  uint32_t ndims = 0;
  uint32_t nbitsets = 0;
  uint64_t segment_size = 0;
  uint64_t row_bytes = 0;
  uint32_t row_stride = 0;     // bytes per row of the row-major mirror (0: no mirror)
  std::vector<SegmentData> segs;
  SegDesc *d_segs = nullptr;
  size_t d_segs_cap = 0;
  // per segment, per column: ordered min / max, reduced on the device at put time and read back in
  // one batch by the next query (a put never waits for its own statistics)
  unsigned long long *d_stats = nullptr;
  size_t d_stats_segs = 0;
  unsigned long long *h_stats_init = nullptr;  // pinned pattern {~0, 0} x ncols
  bool descs_dirty = true;
  bool stats_dirty = false;       // some segment's statistics are still on the device
  // vgpu_segment_put_async: converted CSR offsets that must outlive the copies still in flight
  std::vector<std::vector<uint32_t>> pending_keep;
  // scratch high-water marks so that a repeated query shape never re-runs on overflow (benign races between
  // concurrent queries: any value is a valid hint)
  std::atomic<uint64_t> hash_cap_hint{0};
  std::atomic<uint64_t> pairs_total_hint{0};   // most count-distinct pairs one query produced
  std::atomic<uint64_t> pairs_region_hint{0};  // fullest pair region of one query
  std::atomic<uint64_t> groups_hint{0};        // most groups one query returned
  std::atomic<uint32_t> distinct_general{0};   // the shared-memory-set path overflowed before: go straight to the general one
};

using Planner = PlannerT<vgpu_table>;   // planner.h

// results of select / search queries: arrays inside one pinned host block
struct vgpu_rows {
  std::shared_ptr<PinnedPool> pool;
  std::pair<void *, size_t> block{nullptr, 0};
  std::vector<const void *> cell_ptrs;
  vgpu_rows_view view{};
  ~vgpu_rows() {
    if (block.first && pool) pool->release(block);
  }
};
struct vgpu_search {
  std::vector<uint64_t> seg_offsets, codes;
  std::vector<uint32_t> first_row;
  vgpu_search_view view{};
};

struct vgpu_result {
  std::shared_ptr<PinnedPool> pool;
  std::pair<void *, size_t> block{nullptr, 0};  // pinned host memory holding every array of the view
  std::vector<const void *> key_ptrs, acc_ptrs;
  vgpu_result_view view{};
  ~vgpu_result() {
    if (block.first && pool) pool->release(block);
  }
};

namespace {

// ---------------------------------------------------------------------------------------------
// device scratch: stream-ordered pool allocations, released when the query ends
// ---------------------------------------------------------------------------------------------
struct Scratch {
  cudaStream_t stream;
  cudaStream_t side;   // a second stream that may still use the allocations when an error unwinds the query
  std::vector<void *> ptrs;
  explicit Scratch(cudaStream_t s, cudaStream_t side_stream = nullptr) : stream(s), side(side_stream) {}
  ~Scratch() {
    if (side) cudaStreamSynchronize(side);
    for (void *p : ptrs) cudaFreeAsync(p, stream);
  }
  template <class T> T *alloc(uint64_t n) {
    void *p = nullptr;
    CUDA_CK(cudaMallocAsync(&p, std::max<uint64_t>(n, 1) * sizeof(T), stream));
    ptrs.push_back(p);
    return static_cast<T *>(p);
  }
};

void destroy_scope(QueryScope *sc) {
  if (!sc) return;
  if (sc->s0) cudaStreamDestroy(sc->s0);
  if (sc->s1) cudaStreamDestroy(sc->s1);
  for (cudaEvent_t e : {sc->ev_begin, sc->ev_scan0, sc->ev_scan1, sc->ev_end, sc->ev_a, sc->ev_b, sc->ev_c})
    if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : sc->ev_phase)
    if (e) cudaEventDestroy(e);
  if (sc->d_counters) cudaFree(sc->d_counters);
  if (sc->h_counters) cudaFreeHost(sc->h_counters);
  if (sc->d_plan) cudaFree(sc->d_plan);
  if (sc->h_plan) cudaFreeHost(sc->h_plan);
  delete sc;
}

// RAII lease of a QueryScope; entering also orders the query after the caller's stream, leaving makes the
// caller's stream wait for the query
struct ScopeLease {
  vgpu_ctx *ctx;
  QueryScope *sc = nullptr;
  cudaStream_t user = nullptr;
  explicit ScopeLease(vgpu_ctx *c) : ctx(c) {
    {
      std::lock_guard<std::mutex> lk(ctx->scope_mu);
      user = ctx->user_stream;
      if (!ctx->idle_scopes.empty()) { sc = ctx->idle_scopes.back(); ctx->idle_scopes.pop_back(); }
    }
    if (!sc) {
      std::unique_ptr<QueryScope, void (*)(QueryScope *)> n(new QueryScope(), destroy_scope);
      CUDA_CK(cudaStreamCreateWithFlags(&n->s0, cudaStreamNonBlocking));
      CUDA_CK(cudaStreamCreateWithFlags(&n->s1, cudaStreamNonBlocking));
      CUDA_CK(cudaEventCreate(&n->ev_begin));
      CUDA_CK(cudaEventCreate(&n->ev_scan0));
      CUDA_CK(cudaEventCreate(&n->ev_scan1));
      CUDA_CK(cudaEventCreate(&n->ev_end));
      CUDA_CK(cudaEventCreateWithFlags(&n->ev_a, cudaEventDisableTiming));
      CUDA_CK(cudaEventCreateWithFlags(&n->ev_b, cudaEventDisableTiming));
      CUDA_CK(cudaEventCreateWithFlags(&n->ev_c, cudaEventDisableTiming));
      for (cudaEvent_t &e : n->ev_phase) CUDA_CK(cudaEventCreate(&e));
      CUDA_CK(cudaMalloc(&n->d_counters, 16 * sizeof(unsigned long long)));
      CUDA_CK(cudaMallocHost(&n->h_counters, 48 * sizeof(unsigned long long)));
      CUDA_CK(cudaMalloc(&n->d_plan, 64 * sizeof(uint64_t)));
      CUDA_CK(cudaMallocHost(&n->h_plan, 64 * sizeof(uint64_t)));
      sc = n.release();
    }
    if (user) {   // the scope's own event: concurrent queries never share one
      cudaEventRecord(sc->ev_c, user);
      cudaStreamWaitEvent(sc->s0, sc->ev_c, 0);
    }
  }
  ~ScopeLease() {
    if (!sc) return;
    // whatever happened (errors included), nothing of this query may still be running when the scope is reused
    cudaStreamSynchronize(sc->s0);
    cudaStreamSynchronize(sc->s1);
    if (user) {
      cudaEventRecord(sc->ev_c, sc->s0);
      cudaStreamWaitEvent(user, sc->ev_c, 0);
    }
    std::lock_guard<std::mutex> lk(ctx->scope_mu);
    ctx->idle_scopes.push_back(sc);
  }
};

int grid_for(uint64_t n, int threads, int sm_count, int per_sm = 8) {
  uint64_t blocks = (n + threads - 1) / threads;
  uint64_t cap = (uint64_t)sm_count * per_sm;
  return (int)std::max<uint64_t>(1, std::min(blocks, cap));
}

// return the number of kernels launched (memsets are driver operations, not our kernels)
uint32_t fill32(cudaStream_t s, int sms, void *p, uint64_t n, uint32_t v) {
  if (n == 0) return 0;
  if (v == 0) { CUDA_CK(cudaMemsetAsync(p, 0, n * 4, s)); return 0; }
  fill32_kernel<<<grid_for(n, 256, sms), 256, 0, s>>>(static_cast<uint32_t *>(p), n, v);
  CUDA_CK(cudaGetLastError());
  return 1;
}
uint32_t fill64(cudaStream_t s, int sms, void *p, uint64_t n, uint64_t v) {
  if (n == 0) return 0;
  if (v == 0) { CUDA_CK(cudaMemsetAsync(p, 0, n * 8, s)); return 0; }
  if (v == ~0ull) { CUDA_CK(cudaMemsetAsync(p, 0xff, n * 8, s)); return 0; }
  fill64_kernel<<<grid_for(n, 256, sms), 256, 0, s>>>(static_cast<uint64_t *>(p), n, v);
  CUDA_CK(cudaGetLastError());
  return 1;
}

// ---------------------------------------------------------------------------------------------
// column store
// ---------------------------------------------------------------------------------------------
void free_segment(SegmentData &sd) {
  if (sd.slab) cudaFree(sd.slab);
  if (sd.rows) cudaFree(sd.rows);
  for (int b = 0; b < kMaxBitsetCols; ++b) {
    if (sd.bs_values[b]) cudaFree(sd.bs_values[b]);
    if (sd.bs_offsets[b]) cudaFree(sd.bs_offsets[b]);
  }
  sd = SegmentData();
}

void ensure_segment(vgpu_table *t, uint32_t seg_idx, uint64_t nrows, cudaStream_t zero_stream) {
  if (nrows > t->segment_size)
    fail(VGPU_ERR_INVALID, "segment rows " + std::to_string(nrows) + " exceed segment_size " +
                               std::to_string(t->segment_size));
  if (seg_idx >= t->segs.size()) t->segs.resize(seg_idx + 1);
  SegmentData &sd = t->segs[seg_idx];
  uint64_t cap = round_up(std::max<uint64_t>(nrows, 1), kTileRows);
  // the newest segment is the one the live store still appends to (db/store.h:46-52): room for a whole segment, so
  // that vgpu_segment_update can append behind its rows without a re-put
  if (seg_idx + 1 == t->segs.size() && t->segment_size <= (1ull << 24)) cap = std::max(cap, round_up(t->segment_size, kTileRows));
  if (sd.slab == nullptr || sd.cap < cap) {
    free_segment(sd);
    if (t->row_bytes > 0) {
      CUDA_CK(cudaMalloc(&sd.slab, t->row_bytes * cap));
      // zero once: vector loads may over-read to the end of a chunk; later puts only clear what shrinks
      CUDA_CK(cudaMemsetAsync(sd.slab, 0, t->row_bytes * cap, zero_stream));
    }
    sd.cap = cap;
    sd.hi_rows = 0;
    if (t->row_stride) {
      // the mirror is an optimisation: without memory for it the columnar gathers serve every query
      if (cudaMalloc(&sd.rows, (uint64_t)t->row_stride * cap + 64) == cudaSuccess) sd.rows_cap = cap;
      else { sd.rows = nullptr; cudaGetLastError(); }
    }
    t->descs_dirty = true;
  } else {
    for (int b = 0; b < kMaxBitsetCols; ++b) sd.bs_n[b] = 0;  // buffers are kept and reused
  }
  if (sd.nrows != nrows || !sd.valid) t->descs_dirty = true;
  sd.nrows = nrows;
  sd.valid = false;
}

// per-column min/max of the stored cells (device reduction), kept in the ordered domain. Launch only:
// the values are fetched by fetch_stats() when a query needs them.
void ensure_stats_capacity(vgpu_table *t, size_t nsegs) {
  const size_t ncols = t->cols.size();
  if (t->h_stats_init == nullptr) {
    CUDA_CK(cudaMallocHost(&t->h_stats_init, 2 * ncols * sizeof(unsigned long long)));
    for (size_t c = 0; c < ncols; ++c) { t->h_stats_init[2 * c] = ~0ull; t->h_stats_init[2 * c + 1] = 0; }
  }
  if (nsegs <= t->d_stats_segs) return;
  size_t cap = std::max<size_t>(64, t->d_stats_segs * 2);
  while (cap < nsegs) cap *= 2;
  unsigned long long *n = nullptr;
  CUDA_CK(cudaMalloc(&n, cap * 2 * ncols * sizeof(unsigned long long)));
  if (t->d_stats) {
    CUDA_CK(cudaMemcpyAsync(n, t->d_stats, t->d_stats_segs * 2 * ncols * sizeof(unsigned long long),
                            cudaMemcpyDeviceToDevice, t->ctx->stream));
    CUDA_CK(cudaStreamSynchronize(t->ctx->stream));
    cudaFree(t->d_stats);
  }
  t->d_stats = n;
  t->d_stats_segs = cap;
}

void compute_stats(vgpu_table *t, uint32_t seg_idx) {
  vgpu_ctx *ctx = t->ctx;
  SegmentData &sd = t->segs[seg_idx];
  const size_t ncols = t->cols.size();
  sd.omin.assign(ncols, ~0ull);
  sd.omax.assign(ncols, 0ull);
  sd.stats_pending = false;
  std::vector<uint32_t> want;
  for (uint32_t c = 0; c < t->ndims; ++c)
    if (!t->cols[c].bitset) want.push_back(c);
  // COUNT metrics too: a count cell is the number of raw rows behind a stored row, >= 1 in every table the reference
  // builds — when the statistics confirm it, "group present" can be read off the count accumulator (query_agg.inl)
  for (uint32_t c = t->ndims; c < t->cols.size(); ++c)
    if (!t->cols[c].bitset && t->cols[c].agg == VGPU_AGG_COUNT && !type_float(t->cols[c].type) && !type_signed(t->cols[c].type)) want.push_back(c);
  if (sd.nrows == 0 || want.empty()) return;
  ensure_stats_capacity(t, t->segs.size());
  unsigned long long *d_out = t->d_stats + (size_t)seg_idx * 2 * ncols;
  CUDA_CK(cudaMemcpyAsync(d_out, t->h_stats_init, 2 * ncols * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream));
  for (size_t base = 0; base < want.size(); base += 32) {
    uint32_t n = (uint32_t)std::min<size_t>(32, want.size() - base);
    StatParams S{};
    S.slab = sd.slab;
    S.nrows = sd.nrows;
    S.ncols = n;
    for (uint32_t i = 0; i < n; ++i) {
      const ColInfo &ci = t->cols[want[base + i]];
      S.cols[i].off = ci.off_per_row * sd.cap;
      S.cols[i].width = ci.width;
      S.cols[i].sext = ci.sext;
      S.cols[i].type = ci.type;
      S.cols[i].pad = want[base + i];  // output slot = schema column
    }
    S.out = d_out;
    dim3 grid(grid_for(sd.nrows, 256, ctx->sm_count, 2), n);
    column_minmax_kernel<<<grid, 256, 0, ctx->stream>>>(S);
    CUDA_CK(cudaGetLastError());
  }
  sd.stats_pending = true;
  t->stats_dirty = true;
}

// (re)build the row-major mirror of a segment from its columns; stream-ordered after the column copies
void build_row_mirror(vgpu_table *t, uint32_t seg_idx, uint64_t row_begin = 0, uint64_t row_count = ~0ull) {
  vgpu_ctx *ctx = t->ctx;
  SegmentData &sd = t->segs[seg_idx];
  if (sd.rows == nullptr || sd.nrows == 0) return;
  if (t->cols.size() > 32) fail(VGPU_ERR_UNSUPPORTED, "row mirror supports at most 32 columns");
  row_begin = std::min(row_begin, sd.nrows);
  row_count = std::min(row_count, sd.nrows - row_begin);
  if (row_count == 0) return;
  RowsParams R{};
  R.rows = sd.rows + row_begin * t->row_stride;
  R.nrows = row_count;
  R.stride = t->row_stride;
  R.ncols = (uint32_t)t->cols.size();
  for (size_t c = 0; c < t->cols.size(); ++c) {
    const ColInfo &ci = t->cols[c];
    RowsCol &rc = R.cols[c];
    rc.row_off = ci.row_off;
    if (ci.bitset) {  // the first word a count-distinct needs: the id, or the CSR offset of the cell
      rc.width = 4;
      rc.src = reinterpret_cast<const uint8_t *>(sd.bs_has_offsets[ci.bitset_idx] ? sd.bs_offsets[ci.bitset_idx]
                                                                                    : sd.bs_values[ci.bitset_idx]) + row_begin * 4;
    } else {
      rc.width = ci.width;
      rc.src = sd.slab + ci.off_per_row * sd.cap + row_begin * ci.width;
    }
  }
  const uint32_t tile_rows = kRowsTileBytes / R.stride;
  const uint64_t ntiles = (row_count + tile_rows - 1) / tile_rows;
  const int grid = (int)std::min<uint64_t>(ntiles, (uint64_t)ctx->sm_count * 8);
  build_rows_kernel<<<grid, 256, 0, ctx->stream>>>(R);
  CUDA_CK(cudaGetLastError());
}

// one batched read-back for every segment whose statistics are still on the device
void fetch_stats(vgpu_table *t) {
  const size_t ncols = t->cols.size();
  size_t lo = t->segs.size(), hi = 0;
  for (size_t s = 0; s < t->segs.size(); ++s)
    if (t->segs[s].stats_pending) { lo = std::min(lo, s); hi = std::max(hi, s + 1); }
  if (lo >= hi) return;
  std::vector<unsigned long long> h((hi - lo) * 2 * ncols);
  CUDA_CK(cudaMemcpyAsync(h.data(), t->d_stats + lo * 2 * ncols, h.size() * sizeof(unsigned long long),
                          cudaMemcpyDeviceToHost, t->ctx->stream));
  CUDA_CK(cudaStreamSynchronize(t->ctx->stream));
  for (size_t s = lo; s < hi; ++s) {
    SegmentData &sd = t->segs[s];
    if (!sd.stats_pending) continue;
    for (size_t c = 0; c < ncols; ++c) {
      sd.omin[c] = h[(s - lo) * 2 * ncols + 2 * c];
      sd.omax[c] = h[(s - lo) * 2 * ncols + 2 * c + 1];
    }
    sd.stats_pending = false;
  }
  t->stats_dirty = false;
}

void upload_descs(vgpu_table *t) {
  if (!t->descs_dirty) return;
  size_t n = t->segs.size();
  if (n > t->d_segs_cap) {
    if (t->d_segs) cudaFree(t->d_segs);
    t->d_segs = nullptr;
    size_t cap = std::max<size_t>(16, n * 2);
    CUDA_CK(cudaMalloc(&t->d_segs, cap * sizeof(SegDesc)));
    t->d_segs_cap = cap;
  }
  std::vector<SegDesc> h(n);
  for (size_t i = 0; i < n; ++i) {
    const SegmentData &sd = t->segs[i];
    h[i].slab = sd.slab;
    h[i].nrows = sd.valid ? sd.nrows : 0;
    h[i].cap = sd.cap;
    h[i].rows = sd.rows;
    for (int b = 0; b < kMaxBitsetCols; ++b) {
      h[i].bs_values[b] = sd.bs_values[b];
      h[i].bs_offsets[b] = sd.bs_has_offsets[b] ? sd.bs_offsets[b] : nullptr;
    }
  }
  if (n) {
    CUDA_CK(cudaMemcpyAsync(t->d_segs, h.data(), n * sizeof(SegDesc), cudaMemcpyHostToDevice,
                            t->ctx->stream));
    CUDA_CK(cudaStreamSynchronize(t->ctx->stream));  // h goes out of scope
  }
  t->descs_dirty = false;
}

struct AccInfo {
  uint32_t op;         // AccOp
  uint32_t acc_width;  // 4 or 8
  uint32_t out_width;  // metric column width (8 for BITSET cardinality)
  uint64_t init;
  ncclDataType_t nccl_type;
  ncclRedOp_t nccl_op;
};

AccInfo acc_for(const ColInfo &ci) {
  AccInfo a{};
  a.out_width = ci.width;
  const uint32_t t = ci.type;
  if (ci.bitset) {
    a.op = A_DISTINCT; a.acc_width = 4; a.out_width = 8; a.init = 0;
    a.nccl_type = ncclUint32; a.nccl_op = ncclSum;
    return a;
  }
  switch (ci.agg) {
    case VGPU_AGG_SUM: case VGPU_AGG_AVG: case VGPU_AGG_COUNT:
      a.init = 0; a.nccl_op = ncclSum;
      if (t == VGPU_F32) { a.op = A_ADDF32; a.acc_width = 4; a.nccl_type = ncclFloat32; }
      else if (t == VGPU_F64) { a.op = A_ADDF64; a.acc_width = 8; a.nccl_type = ncclFloat64; }
      else if (ci.width == 8) { a.op = A_ADD64; a.acc_width = 8; a.nccl_type = ncclUint64; }
      else { a.op = A_ADD32; a.acc_width = 4; a.nccl_type = ncclUint32; }
      break;
    case VGPU_AGG_MAX: case VGPU_AGG_MIN: {
      const bool mx = ci.agg == VGPU_AGG_MAX;
      a.init = mx ? type_min_value(t) : type_max_value(t);
      a.nccl_op = mx ? ncclMax : ncclMin;
      if (t == VGPU_F32) { a.op = mx ? A_MAXF32 : A_MINF32; a.acc_width = 4; a.nccl_type = ncclFloat32; }
      else if (t == VGPU_F64) { a.op = mx ? A_MAXF64 : A_MINF64; a.acc_width = 8; a.nccl_type = ncclFloat64; }
      else if (ci.width == 8) {
        if (type_signed(t)) { a.op = mx ? A_MAXS64 : A_MINS64; a.nccl_type = ncclInt64; }
        else { a.op = mx ? A_MAXU64 : A_MINU64; a.nccl_type = ncclUint64; }
        a.acc_width = 8;
      } else {
        if (type_signed(t)) { a.op = mx ? A_MAXS32 : A_MINS32; a.nccl_type = ncclInt32; }
        else { a.op = mx ? A_MAXU32 : A_MINU32; a.nccl_type = ncclUint32; }
        a.acc_width = 4;
      }
    } break;
    default:
      fail(VGPU_ERR_INVALID, "metric column without an aggregation type");
  }
  return a;
}

struct KeyRange {
  uint64_t lo;     // widened representation (two's complement for signed)
  uint64_t range;  // number of distinct representable values, 0 means 2^64
};

unsigned __int128 range128(const KeyRange &r) {
  return r.range == 0 ? ((unsigned __int128)1 << 64) : (unsigned __int128)r.range;
}

// host copy of the device trunc for the lower bound of a rolled-up time key
uint64_t host_trunc_year_seconds(uint64_t t) {
  time_t tt = (time_t)t;
  struct tm tm;
  gmtime_r(&tt, &tm);
  tm.tm_sec = 0; tm.tm_min = 0; tm.tm_hour = 0; tm.tm_mday = 1; tm.tm_mon = 0;
  return (uint64_t)timegm(&tm);
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int vgpu_abi_version(void) { return VGPU_ABI_VERSION; }

const char *vgpu_last_error(void) { return g_err.c_str(); }

int vgpu_init(int device, vgpu_ctx **out) {
  return guard([&] {
    if (!out) fail(VGPU_ERR_INVALID, "null output pointer");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
      fail(VGPU_ERR_CUDA, std::string("no CUDA device: this library has no CPU path (") +
                              cudaGetErrorString(e) + ")");
    if (device < 0 || device >= ndev) fail(VGPU_ERR_INVALID, "device index out of range");
    CUDA_CK(cudaSetDevice(device));
    std::unique_ptr<vgpu_ctx> ctx(new vgpu_ctx());
    ctx->device = device;
    cudaDeviceProp prop;
    CUDA_CK(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    CUDA_CK(cudaFuncSetAttribute(scan_filter_groupby_kernel<VGPU_MIN_CTAS, true, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemTableBytes));
    CUDA_CK(cudaFuncSetAttribute(scan_filter_groupby_kernel<VGPU_MIN_CTAS, true, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemTableBytes));
    CUDA_CK(cudaFuncSetAttribute(scan_filter_groupby_kernel<VGPU_MIN_CTAS, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemTableBytes));
    CUDA_CK(cudaFuncSetAttribute(scan_filter_groupby_kernel<VGPU_MIN_CTAS, true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemTableBytes));
    CUDA_CK(cudaFuncSetAttribute(scan_filter_groupby_kernel<VGPU_MIN_CTAS, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemTableBytes));
    CUDA_CK(cudaFuncSetAttribute(scan_filter_groupby_kernel<VGPU_MIN_CTAS, true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemTableBytes));
    CUDA_CK(cudaFuncSetAttribute(pairs_bucket_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * kMaxSmemBuckets * 4)));
    CUDA_CK(cudaFuncSetAttribute(pairs_dedupe_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSmemSetSlots * 8)));
    CUDA_CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CUDA_CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    CUDA_CK(cudaEventCreateWithFlags(&ctx->ev_copy, cudaEventDisableTiming));
    // keep freed scratch in the pool: repeated queries never go back to the driver
    cudaMemPool_t pool;
    CUDA_CK(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t threshold = UINT64_MAX;
    CUDA_CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold));
    if (const char *e = getenv("VGPU_TUNE")) ctx->tune = (uint32_t)strtoul(e, nullptr, 0);
    ctx->trace = getenv("VGPU_TRACE") != nullptr;
    if (const char *e = getenv("VGPU_UNIT_CHUNKS")) ctx->unit_chunks = (uint32_t)strtoul(e, nullptr, 0);
    if (const char *e = getenv("VGPU_TEST_PAIRS_CAP")) ctx->test_pairs_cap = strtoull(e, nullptr, 0);
    if (const char *e = getenv("VGPU_TEST_HASH_CAP")) ctx->test_hash_cap = strtoull(e, nullptr, 0);
    if (const char *e = getenv("VGPU_TEST_BUCKET_PAIRS")) ctx->test_bucket_pairs = strtoull(e, nullptr, 0);
    if (const char *e = getenv("VGPU_TEST_SMALL_PAIRS")) ctx->test_small_pairs = strtoull(e, nullptr, 0);
    if (const char *e = getenv("VGPU_TEST_SET_SLOTS")) ctx->test_set_slots = (uint32_t)strtoul(e, nullptr, 0);
    // carve out the persisting part of L2 for group tables
    if (ctx->tune & 1u) {
      int max_persist = 0, max_window = 0;
      cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, device);
      cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, device);
      if (max_persist > 0 && max_window > 0 &&
          cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist) == cudaSuccess) {
        ctx->l2_persist_bytes = (uint64_t)max_persist;
        ctx->l2_window_max = (uint64_t)max_window;
      }
      cudaGetLastError();
    }
    *out = ctx.release();
  });
}

int vgpu_set_stream(vgpu_ctx *ctx, void *cuda_stream) {
  return guard([&] {
    if (!ctx) fail(VGPU_ERR_INVALID, "null context");
    std::lock_guard<std::mutex> lk(ctx->scope_mu);
    ctx->user_stream = static_cast<cudaStream_t>(cuda_stream);
  });
}

int vgpu_set_test_hook(vgpu_ctx *ctx, const char *name, uint64_t value) {
  return guard([&] {
    if (!ctx || !name) fail(VGPU_ERR_INVALID, "null argument");
    const std::string n(name);
    if (n == "pairs_cap") ctx->test_pairs_cap = value;
    else if (n == "hash_cap") ctx->test_hash_cap = value;
    else if (n == "bucket_pairs") ctx->test_bucket_pairs = value;
    else if (n == "small_pairs") ctx->test_small_pairs = value;
    else if (n == "set_slots") ctx->test_set_slots = (uint32_t)value;
    else if (n == "expect_pairs") ctx->test_expect = value;
    else if (n == "tune") ctx->tune = (uint32_t)value;
    else fail(VGPU_ERR_INVALID, "unknown test hook " + n);
  });
}

void vgpu_shutdown(vgpu_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
  for (QueryScope *sc : ctx->idle_scopes) destroy_scope(sc);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->ev_copy) cudaEventDestroy(ctx->ev_copy);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

// ---------------------------------------------------------------------------------------------
// column store
// ---------------------------------------------------------------------------------------------
int vgpu_table_create(vgpu_ctx *ctx, const vgpu_schema *schema, vgpu_table **out) {
  return guard([&] {
    if (!ctx || !schema || !out) fail(VGPU_ERR_INVALID, "null argument");
    *out = nullptr;
    if (schema->ncols == 0 || schema->ndims > schema->ncols || !schema->cols)
      fail(VGPU_ERR_INVALID, "malformed schema");
    if (schema->segment_size == 0) fail(VGPU_ERR_INVALID, "segment_size must be positive");
    if (schema->segment_size > (1ull << 31)) fail(VGPU_ERR_UNSUPPORTED, "segment_size above 2^31 rows");
    std::unique_ptr<vgpu_table> t(new vgpu_table());
    t->ctx = ctx;
    CUDA_CK(cudaSetDevice(ctx->device));
    CUDA_CK(cudaEventCreateWithFlags(&t->ev_put, cudaEventDisableTiming));
    t->ndims = schema->ndims;
    t->segment_size = schema->segment_size;
    uint64_t off = 0;
    for (uint32_t c = 0; c < schema->ncols; ++c) {
      const vgpu_column &sc = schema->cols[c];
      ColInfo ci{};
      ci.kind = sc.kind; ci.type = sc.type; ci.agg = sc.agg;
      ci.lit_type = sc.lit_type ? sc.lit_type - 1 : sc.type;
      if (ci.lit_type > VGPU_F64) fail(VGPU_ERR_INVALID, "bad literal type");
      const bool is_dim = c < schema->ndims;
      if (is_dim != (sc.kind <= VGPU_DIM_BOOLEAN))
        fail(VGPU_ERR_INVALID, "dimensions must precede metrics in the schema");
      switch (sc.kind) {
        case VGPU_DIM_STRING:
          if (!(sc.type <= VGPU_U64)) fail(VGPU_ERR_INVALID, "string dimension codes are unsigned");
          break;
        case VGPU_DIM_TIME:
          if (sc.type != VGPU_U32) fail(VGPU_ERR_INVALID, "time dimension is uint32 seconds");
          break;
        case VGPU_DIM_MICROTIME:
          if (sc.type != VGPU_U64) fail(VGPU_ERR_INVALID, "microtime dimension is uint64 microseconds");
          break;
        case VGPU_DIM_BOOLEAN:
          if (sc.type != VGPU_U8) fail(VGPU_ERR_INVALID, "boolean dimension is uint8");
          break;
        case VGPU_DIM_NUMERIC:
        case VGPU_METRIC_VALUE:
          break;
        case VGPU_METRIC_HIDDEN_COUNT:
          if (sc.type != VGPU_U64) fail(VGPU_ERR_INVALID, "hidden count column is uint64");
          ci.agg = VGPU_AGG_COUNT;
          break;
        case VGPU_METRIC_BITSET:
          if (sc.type != VGPU_U32 && sc.type != VGPU_U64) fail(VGPU_ERR_INVALID, "bitset ids travel as uint32 or uint64");
          break;
        default:
          fail(VGPU_ERR_INVALID, "unknown column kind");
      }
      ci.width = type_width(sc.type);
      ci.sext = type_signed(sc.type);
      if (sc.kind == VGPU_METRIC_BITSET) {
        if (t->nbitsets >= kMaxBitsetCols) fail(VGPU_ERR_UNSUPPORTED, "too many bitset metrics");
        ci.bitset = true;
        ci.bitset_idx = t->nbitsets++;
        ci.agg = VGPU_AGG_BITSET;
        ci.off_per_row = 0;
      } else {
        if (sc.kind == VGPU_METRIC_VALUE && sc.agg > VGPU_AGG_COUNT)
          fail(VGPU_ERR_INVALID, "value metric needs max/min/sum/avg/count");
        // natural alignment inside the slab: widest columns need 8-byte aligned bases; cap is a
        // multiple of 4096 rows, so off_per_row * cap is always 4096-byte aligned
        ci.off_per_row = off;
        off += ci.width;
      }
      t->cols.push_back(ci);
    }
    t->row_bytes = off;
    // Row-major mirror: cells ordered by decreasing width so that each is naturally aligned, the row
    // padded to the widest cell (8-byte cells need 8-byte aligned rows for the aligned-word gathers).
    if (!(ctx->tune & 64u) && schema->ncols <= 32) {
      uint32_t roff = 0, align = 4;
      for (uint32_t w : {8u, 4u, 2u, 1u})
        for (auto &ci : t->cols) {
          const uint32_t cw = ci.bitset ? 4u : ci.width;
          if (cw != w) continue;
          ci.row_off = roff;
          roff += cw;
          align = std::max(align, cw);
        }
      t->row_stride = (uint32_t)round_up(roff, align);
      if (t->row_stride > kRowsTileBytes / 32) t->row_stride = 0;  // very wide rows: no mirror
    }
    *out = t.release();
  });
}

void vgpu_table_free(vgpu_table *table) {
  if (!table) return;
  cudaSetDevice(table->ctx->device);
  cudaDeviceSynchronize();
  if (table->ev_put) cudaEventDestroy(table->ev_put);
  for (auto &sd : table->segs) free_segment(sd);
  if (table->d_segs) cudaFree(table->d_segs);
  if (table->d_stats) cudaFree(table->d_stats);
  if (table->h_stats_init) cudaFreeHost(table->h_stats_init);
  delete table;
}

static int put_impl(vgpu_table *t, uint32_t seg_idx, uint64_t nrows, const void *const *col_ptrs, bool wait) {
  return guard([&] {
    if (!t || (!col_ptrs && nrows)) fail(VGPU_ERR_INVALID, "null argument");
    vgpu_ctx *ctx = t->ctx;
    std::lock_guard<std::mutex> put_lk(ctx->put_mu);
    std::unique_lock<std::shared_mutex> lk(t->mu);
    CUDA_CK(cudaSetDevice(ctx->device));
    if (seg_idx > t->segs.size() + (1u << 20)) fail(VGPU_ERR_INVALID, "segment index too sparse");
    cudaStream_t cs = ctx->copy_stream;
    ensure_segment(t, seg_idx, nrows, cs);
    SegmentData &sd = t->segs[seg_idx];
    std::vector<std::vector<uint32_t>> &keep = t->pending_keep;  // converted CSR offsets must outlive the async copies
    const uint64_t old_hi = sd.hi_rows;
    for (size_t c = 0; c < t->cols.size(); ++c) {
      const ColInfo &ci = t->cols[c];
      if (ci.bitset) {
        if (nrows == 0) continue;
        const vgpu_bitset_csr *csr = static_cast<const vgpu_bitset_csr *>(col_ptrs[c]);
        if (!csr || !csr->values) fail(VGPU_ERR_INVALID, "bitset column needs a vgpu_bitset_csr");
        uint64_t nvalues = csr->offsets ? csr->offsets[nrows] : nrows;
        if (csr->offsets && csr->nvalues != nvalues) fail(VGPU_ERR_INVALID, "bitset CSR: nvalues != offsets[nrows]");
        if (!csr->offsets && csr->nvalues != nrows) fail(VGPU_ERR_INVALID, "bitset without offsets needs one id per row");
        if (nvalues >= (1ull << 32)) fail(VGPU_ERR_UNSUPPORTED, "more than 2^32 bitset ids in one segment");
        const uint32_t idw = ci.width == 8 ? 8u : 4u;   // 64-bit ids: util::Bitset<8> = Roaring64Map (bitset.h:27-31)
        bool one_per_row = idw == 4;  // cells of 64-bit ids always go through CSR offsets (MetSpec::id64)
        if (csr->offsets) {
          if (csr->offsets[0] != 0) fail(VGPU_ERR_INVALID, "bitset CSR: offsets[0] != 0");
          for (uint64_t r = 0; r < nrows; ++r) {
            if (csr->offsets[r + 1] < csr->offsets[r]) fail(VGPU_ERR_INVALID, "bitset CSR: offsets not monotone");
            if (csr->offsets[r + 1] - csr->offsets[r] != 1) one_per_row = false;
          }
        }
        // values padded to a whole tile so that speculative reads stay in bounds
        uint64_t vcap = round_up(std::max<uint64_t>(nvalues, 1), kTileRows) * (idw / 4);   // in uint32 words
        if (one_per_row) vcap = std::max<uint64_t>(vcap, sd.cap);   // appends (vgpu_segment_update) fit like in the slab
        if (sd.bs_vcap[ci.bitset_idx] < vcap) {
          if (sd.bs_values[ci.bitset_idx]) cudaFree(sd.bs_values[ci.bitset_idx]);
          sd.bs_values[ci.bitset_idx] = nullptr;
          CUDA_CK(cudaMalloc(&sd.bs_values[ci.bitset_idx], vcap * 4));
          sd.bs_vcap[ci.bitset_idx] = vcap;
          t->descs_dirty = true;
        }
        if (sd.bs_vcap[ci.bitset_idx] * 4 > nvalues * idw)
          CUDA_CK(cudaMemsetAsync(reinterpret_cast<uint8_t *>(sd.bs_values[ci.bitset_idx]) + nvalues * idw, 0,
                                  sd.bs_vcap[ci.bitset_idx] * 4 - nvalues * idw, cs));
        if (nvalues)
          CUDA_CK(cudaMemcpyAsync(sd.bs_values[ci.bitset_idx], csr->values, nvalues * idw, cudaMemcpyHostToDevice, cs));
        sd.bs_n[ci.bitset_idx] = nvalues;
        if (!one_per_row) {
          keep.emplace_back(nrows + 1);
          auto &o32 = keep.back();
          for (uint64_t r = 0; r <= nrows; ++r) o32[r] = csr->offsets ? (uint32_t)csr->offsets[r] : (uint32_t)r;
          if (sd.bs_ocap[ci.bitset_idx] < nrows + 1) {
            if (sd.bs_offsets[ci.bitset_idx]) cudaFree(sd.bs_offsets[ci.bitset_idx]);
            sd.bs_offsets[ci.bitset_idx] = nullptr;
            CUDA_CK(cudaMalloc(&sd.bs_offsets[ci.bitset_idx], (nrows + 1) * 4));
            sd.bs_ocap[ci.bitset_idx] = nrows + 1;
            t->descs_dirty = true;
          }
          CUDA_CK(cudaMemcpyAsync(sd.bs_offsets[ci.bitset_idx], o32.data(), (nrows + 1) * 4,
                                  cudaMemcpyHostToDevice, cs));
          if (!sd.bs_has_offsets[ci.bitset_idx]) t->descs_dirty = true;
          sd.bs_has_offsets[ci.bitset_idx] = true;
        } else {
          if (sd.bs_has_offsets[ci.bitset_idx]) t->descs_dirty = true;
          sd.bs_has_offsets[ci.bitset_idx] = false;
        }
        continue;
      }
      uint8_t *dst = sd.slab + ci.off_per_row * sd.cap;
      if (nrows) {
        if (!col_ptrs[c]) fail(VGPU_ERR_INVALID, "null column pointer");
        CUDA_CK(cudaMemcpyAsync(dst, col_ptrs[c], nrows * ci.width, cudaMemcpyHostToDevice, cs));
      }
      if (old_hi > nrows)   // the segment shrank: clear what the previous content left behind
        CUDA_CK(cudaMemsetAsync(dst + nrows * ci.width, 0, (old_hi - nrows) * ci.width, cs));
    }
    sd.hi_rows = nrows;
    // statistics and the row mirror follow the copies on the compute stream; the host only waits for the
    // copies (its buffers may be reused by the caller now), the kernels overlap the next segment's DMA
    CUDA_CK(cudaEventRecord(ctx->ev_copy, cs));
    CUDA_CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy, 0));
    compute_stats(t, seg_idx);
    build_row_mirror(t, seg_idx);
    CUDA_CK(cudaEventRecord(t->ev_put, ctx->stream));
    sd.valid = true;
    if (wait) {
      CUDA_CK(cudaEventSynchronize(ctx->ev_copy));
      t->pending_keep.clear();
    }
  });
}

int vgpu_segment_put(vgpu_table *t, uint32_t seg_idx, uint64_t nrows, const void *const *col_ptrs) {
  return put_impl(t, seg_idx, nrows, col_ptrs, true);
}

int vgpu_segment_put_async(vgpu_table *t, uint32_t seg_idx, uint64_t nrows, const void *const *col_ptrs) {
  return put_impl(t, seg_idx, nrows, col_ptrs, false);
}

int vgpu_table_sync(vgpu_table *t) {
  return guard([&] {
    if (!t) fail(VGPU_ERR_INVALID, "null table");
    vgpu_ctx *ctx = t->ctx;
    std::lock_guard<std::mutex> put_lk(ctx->put_mu);
    CUDA_CK(cudaSetDevice(ctx->device));
    CUDA_CK(cudaStreamSynchronize(ctx->copy_stream));
    std::unique_lock<std::shared_mutex> lk(t->mu);
    t->pending_keep.clear();
  });
}

int vgpu_host_pin(vgpu_ctx *ctx, const void *ptr, size_t bytes) {
  return guard([&] {
    if (!ctx || !ptr || !bytes) fail(VGPU_ERR_INVALID, "null argument");
    CUDA_CK(cudaSetDevice(ctx->device));
    const uintptr_t lo = reinterpret_cast<uintptr_t>(ptr) & ~uintptr_t(4095);
    const uintptr_t hi = (reinterpret_cast<uintptr_t>(ptr) + bytes + 4095) & ~uintptr_t(4095);
    cudaError_t e = cudaHostRegister(reinterpret_cast<void *>(lo), hi - lo, cudaHostRegisterDefault);
    if (e != cudaSuccess) {
      cudaGetLastError();
      fail(VGPU_ERR_CUDA, std::string("cudaHostRegister: ") + cudaGetErrorString(e));
    }
  });
}

int vgpu_host_unpin(vgpu_ctx *ctx, const void *ptr) {
  return guard([&] {
    if (!ctx || !ptr) fail(VGPU_ERR_INVALID, "null argument");
    CUDA_CK(cudaSetDevice(ctx->device));
    cudaError_t e = cudaHostUnregister(reinterpret_cast<void *>(reinterpret_cast<uintptr_t>(ptr) & ~uintptr_t(4095)));
    if (e != cudaSuccess) {
      cudaGetLastError();
      fail(VGPU_ERR_CUDA, std::string("cudaHostUnregister: ") + cudaGetErrorString(e));
    }
  });
}



int vgpu_segment_update(vgpu_table *t, uint32_t seg_idx, uint64_t row_begin, uint64_t nrows, const void *const *col_ptrs) {
  return guard([&] {
    if (!t || (!col_ptrs && nrows)) fail(VGPU_ERR_INVALID, "null argument");
    vgpu_ctx *ctx = t->ctx;
    std::lock_guard<std::mutex> put_lk(ctx->put_mu);
    std::unique_lock<std::shared_mutex> lk(t->mu);
    CUDA_CK(cudaSetDevice(ctx->device));
    if (seg_idx >= t->segs.size() || !t->segs[seg_idx].valid) fail(VGPU_ERR_STATE, "no such segment: put it first");
    SegmentData &sd = t->segs[seg_idx];
    if (row_begin > sd.nrows) fail(VGPU_ERR_INVALID, "update would leave a hole behind the segment's rows");
    if (row_begin + nrows > sd.cap || row_begin + nrows > t->segment_size)
      fail(VGPU_ERR_STATE, "update beyond the segment's device capacity: put the whole segment");
    if (nrows == 0) return;
    cudaStream_t cs = ctx->copy_stream;
    // queries that are still reading the segment finished before the exclusive lock was granted; the kernels of
    // earlier puts on ctx->stream must finish before their input is overwritten
    CUDA_CK(cudaEventRecord(ctx->ev_copy, ctx->stream));
    CUDA_CK(cudaStreamWaitEvent(cs, ctx->ev_copy, 0));
    for (size_t c = 0; c < t->cols.size(); ++c) {
      const ColInfo &ci = t->cols[c];
      if (!col_ptrs[c]) fail(VGPU_ERR_INVALID, "null column pointer");
      if (ci.bitset) {
        // one id per cell on both sides: anything else changes the CSR layout of the whole segment
        const vgpu_bitset_csr *csr = static_cast<const vgpu_bitset_csr *>(col_ptrs[c]);
        if (sd.bs_has_offsets[ci.bitset_idx] || ci.width == 8 || csr->offsets != nullptr || csr->nvalues != nrows)
          fail(VGPU_ERR_UNSUPPORTED, "partial update of a bitset column whose cells do not hold exactly one 32-bit id: put the whole segment");
        if ((row_begin + nrows) > sd.bs_vcap[ci.bitset_idx]) fail(VGPU_ERR_STATE, "update beyond the bitset capacity: put the whole segment");
        CUDA_CK(cudaMemcpyAsync(sd.bs_values[ci.bitset_idx] + row_begin, csr->values, nrows * 4, cudaMemcpyHostToDevice, cs));
        sd.bs_n[ci.bitset_idx] = std::max<uint64_t>(sd.bs_n[ci.bitset_idx], row_begin + nrows);
        continue;
      }
      CUDA_CK(cudaMemcpyAsync(sd.slab + ci.off_per_row * sd.cap + row_begin * ci.width, col_ptrs[c], nrows * ci.width,
                              cudaMemcpyHostToDevice, cs));
    }
    if (row_begin + nrows > sd.nrows) {
      sd.nrows = row_begin + nrows;
      t->descs_dirty = true;
    }
    sd.hi_rows = std::max(sd.hi_rows, sd.nrows);
    CUDA_CK(cudaEventRecord(ctx->ev_copy, cs));
    CUDA_CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy, 0));
    compute_stats(t, seg_idx);                        // whole segment, on the device: no PCIe traffic
    build_row_mirror(t, seg_idx, row_begin, nrows);   // the updated rows only
    CUDA_CK(cudaEventRecord(t->ev_put, ctx->stream));
    CUDA_CK(cudaEventSynchronize(ctx->ev_copy));
  });
}

int vgpu_segment_generate(vgpu_table *t, uint32_t seg_idx, uint64_t nrows, const vgpu_gen_col *gens,
                          uint64_t seed, uint64_t row_offset) {
  return guard([&] {
    if (!t || !gens) fail(VGPU_ERR_INVALID, "null argument");
    vgpu_ctx *ctx = t->ctx;
    std::lock_guard<std::mutex> put_lk(ctx->put_mu);
    std::unique_lock<std::shared_mutex> lk(t->mu);
    CUDA_CK(cudaSetDevice(ctx->device));
    if (t->cols.size() > 32) fail(VGPU_ERR_UNSUPPORTED, "generator supports at most 32 columns");
    ensure_segment(t, seg_idx, nrows, ctx->stream);
    SegmentData &sd = t->segs[seg_idx];
    if (sd.slab) CUDA_CK(cudaMemsetAsync(sd.slab, 0, t->row_bytes * sd.cap, ctx->stream));
    sd.hi_rows = nrows;
    t->descs_dirty = true;
    GenParams G{};
    G.slab = sd.slab;
    G.nrows = nrows;
    G.seed = seed;
    G.row_offset = row_offset;
    G.ncols = (uint32_t)t->cols.size();
    for (size_t c = 0; c < t->cols.size(); ++c) {
      const ColInfo &ci = t->cols[c];
      GenCol &gc = G.cols[c];
      if (gens[c].range == 0) fail(VGPU_ERR_INVALID, "generator range must be positive");
      if (gens[c].mode == 1 && gens[c].div == 0) fail(VGPU_ERR_INVALID, "generator div must be positive");
      gc.off = ci.off_per_row * sd.cap;
      gc.width = ci.width;
      gc.mode = gens[c].mode;
      gc.lo = gens[c].lo;
      gc.range = gens[c].range;
      gc.div = gens[c].div ? gens[c].div : 1;
      gc.gen_index = (uint32_t)c;
      gc.is_f32 = ci.type == VGPU_F32;
      gc.is_f64 = ci.type == VGPU_F64;
      gc.bitset_out = nullptr;
      if (ci.bitset && ci.width == 8) fail(VGPU_ERR_UNSUPPORTED, "the synthetic generator writes 32-bit bitset ids only");
      if (ci.bitset) {
        uint64_t vcap = round_up(std::max<uint64_t>(nrows, 1), kTileRows);
        if (sd.bs_vcap[ci.bitset_idx] < vcap) {
          if (sd.bs_values[ci.bitset_idx]) cudaFree(sd.bs_values[ci.bitset_idx]);
          sd.bs_values[ci.bitset_idx] = nullptr;
          CUDA_CK(cudaMalloc(&sd.bs_values[ci.bitset_idx], vcap * 4));
          sd.bs_vcap[ci.bitset_idx] = vcap;
        }
        CUDA_CK(cudaMemsetAsync(sd.bs_values[ci.bitset_idx], 0, sd.bs_vcap[ci.bitset_idx] * 4, ctx->stream));
        sd.bs_has_offsets[ci.bitset_idx] = false;
        sd.bs_n[ci.bitset_idx] = nrows;
        gc.bitset_out = sd.bs_values[ci.bitset_idx];
      }
    }
    if (nrows) {
      generate_kernel<<<grid_for(nrows, 256, ctx->sm_count), 256, 0, ctx->stream>>>(G);
      CUDA_CK(cudaGetLastError());
    }
    compute_stats(t, seg_idx);
    build_row_mirror(t, seg_idx);
    CUDA_CK(cudaEventRecord(t->ev_put, ctx->stream));
    sd.valid = true;
  });
}

int vgpu_segment_read(vgpu_table *t, uint32_t seg_idx, uint32_t col, void *out) {
  return guard([&] {
    if (!t || !out) fail(VGPU_ERR_INVALID, "null argument");
    vgpu_ctx *ctx = t->ctx;
    std::lock_guard<std::mutex> put_lk(ctx->put_mu);
    std::shared_lock<std::shared_mutex> lk(t->mu);
    CUDA_CK(cudaSetDevice(ctx->device));
    if (seg_idx >= t->segs.size() || !t->segs[seg_idx].valid) fail(VGPU_ERR_STATE, "no such segment");
    if (col >= t->cols.size()) fail(VGPU_ERR_INVALID, "column index out of range");
    const SegmentData &sd = t->segs[seg_idx];
    const ColInfo &ci = t->cols[col];
    if (ci.bitset) {
      if (sd.bs_n[ci.bitset_idx])
        CUDA_CK(cudaMemcpyAsync(out, sd.bs_values[ci.bitset_idx], sd.bs_n[ci.bitset_idx] * (ci.width == 8 ? 8 : 4),
                                cudaMemcpyDeviceToHost, ctx->stream));
    } else if (sd.nrows) {
      CUDA_CK(cudaMemcpyAsync(out, sd.slab + ci.off_per_row * sd.cap, sd.nrows * ci.width,
                              cudaMemcpyDeviceToHost, ctx->stream));
    }
    CUDA_CK(cudaStreamSynchronize(ctx->stream));
  });
}

int vgpu_table_invalidate(vgpu_table *t, uint32_t seg_idx) {
  return guard([&] {
    if (!t) fail(VGPU_ERR_INVALID, "null table");
    std::lock_guard<std::mutex> put_lk(t->ctx->put_mu);
    std::unique_lock<std::shared_mutex> lk(t->mu);   // no query is using the segment any more
    if (seg_idx >= t->segs.size()) fail(VGPU_ERR_STATE, "no such segment");
    CUDA_CK(cudaSetDevice(t->ctx->device));
    CUDA_CK(cudaStreamSynchronize(t->ctx->stream));
    free_segment(t->segs[seg_idx]);
    t->descs_dirty = true;
  });
}

uint32_t vgpu_table_segments(const vgpu_table *t) {
  if (!t) return 0;
  uint32_t n = 0;
  for (auto &sd : t->segs) n += sd.valid ? 1 : 0;
  return n;
}
uint64_t vgpu_table_rows(const vgpu_table *t) {
  if (!t) return 0;
  uint64_t n = 0;
  for (auto &sd : t->segs) n += sd.valid ? sd.nrows : 0;
  return n;
}
uint64_t vgpu_table_bytes(const vgpu_table *t) {
  if (!t) return 0;
  uint64_t n = 0;
  for (auto &sd : t->segs) {
    if (!sd.valid) continue;
    n += sd.cap * t->row_bytes;
    if (sd.rows) n += sd.rows_cap * t->row_stride;
    for (const ColInfo &ci : t->cols) {
      if (!ci.bitset) continue;
      n += sd.bs_n[ci.bitset_idx] * (ci.width == 8 ? 8 : 4);
      if (sd.bs_has_offsets[ci.bitset_idx]) n += (sd.nrows + 1) * 4;
    }
  }
  return n;
}

// ---------------------------------------------------------------------------------------------
// the hot path
// ---------------------------------------------------------------------------------------------
namespace {

struct QueryRun {
  // inputs
  vgpu_table *t;
  const vgpu_plan *plan;
  // planned
  Planner planner;
  std::vector<AccInfo> accs;       // per selected metric (+ hidden count last)
  std::vector<uint32_t> acc_cols;  // schema column per accumulator
  std::vector<KeyRange> ranges;
  std::vector<uint32_t> active;
  uint64_t active_rows = 0;
  uint64_t scanned_recs = 0;
  bool hash_mode = false;
  bool wide = false;  // hash on the full key tuple
  uint64_t ncells = 0;  // dense cells, or hash capacity (power of two)

  QueryRun(vgpu_table *table, const vgpu_plan *p) : t(table), plan(p), planner(table, p) {}
};

// predicate program -> the columns it streams, the prefetch table, the unrolled-conjunction flag; then the
// segment loop bookkeeping + pruning (scan.cc:42-51). Shared by the aggregate, select and search queries.
void finish_predicate_and_prune(vgpu_ctx *ctx, vgpu_table *t, QueryRun &q) {
  Planner &pl = q.planner;
  pl.build_predicate();
  pl.finish_predicate(ctx->tune);
  for (uint32_t s = 0; s < t->segs.size(); ++s) {
    const SegmentData &sd = t->segs[s];
    if (!sd.valid) continue;
    q.scanned_recs += sd.nrows;
    if (!pl.process_segment(sd, pl.root)) continue;
    q.active.push_back(s);
    q.active_rows += sd.nrows;
  }
}

// the part of ScanParams every scan-core kernel needs besides the predicate: the work list
void set_work_list(vgpu_table *t, QueryRun &q, Scratch &scratch, cudaStream_t stream) {
  ScanParams &P = q.planner.P;
  uint64_t max_rows = 0;
  for (uint32_t s : q.active) max_rows = std::max(max_rows, t->segs[s].nrows);
  P.tiles_per_seg = (uint32_t)std::max<uint64_t>(1, (max_rows + kChunkRows - 1) / kChunkRows);
  P.nactive = (uint32_t)q.active.size();
  P.total_tiles = (uint64_t)P.nactive * P.tiles_per_seg;
  P.segs = t->d_segs;
  P.tune = t->ctx->tune;
  uint32_t *d_active = scratch.alloc<uint32_t>(q.active.size());
  if (!q.active.empty())
    CUDA_CK(cudaMemcpyAsync(d_active, q.active.data(), q.active.size() * 4, cudaMemcpyHostToDevice, stream));
  P.active = d_active;
}

void validate_plan(const vgpu_table *t, const vgpu_plan *plan) {
  if (plan->nnodes && !plan->nodes) fail(VGPU_ERR_INVALID, "null predicate nodes");
  if (plan->nargs && !plan->args) fail(VGPU_ERR_INVALID, "null predicate args");
  if (plan->nkeys && !plan->keys) fail(VGPU_ERR_INVALID, "null keys");
  if (plan->nmetrics && !plan->metric_cols) fail(VGPU_ERR_INVALID, "null metric list");
  if (plan->flags & VGPU_PLAN_POST) {
    if (plan->nhnodes && !plan->hnodes) fail(VGPU_ERR_INVALID, "null HAVING nodes");
    if (plan->nhargs && !plan->hargs) fail(VGPU_ERR_INVALID, "null HAVING args");
    if (plan->sort_col != VGPU_NO_COLUMN && plan->sort_col >= t->cols.size()) fail(VGPU_ERR_INVALID, "sort column out of range");
  }
  if (plan->nkeys > kMaxKeys) fail(VGPU_ERR_UNSUPPORTED, "too many group-by keys");
  if (plan->nmetrics + (plan->need_hidden_count ? 1 : 0) > kMaxMetrics)
    fail(VGPU_ERR_UNSUPPORTED, "too many metrics");
  for (uint32_t k = 0; k < plan->nkeys; ++k) {
    const vgpu_key &key = plan->keys[k];
    if (key.col >= t->ndims) fail(VGPU_ERR_INVALID, "group-by key must be a dimension");
    if (key.nrules > VGPU_MAX_ROLLUP_RULES) fail(VGPU_ERR_INVALID, "too many rollup rules");
    const ColInfo &ci = t->cols[key.col];
    const bool rollup = key.nrules > 0 || key.query_granularity != VGPU_TU_NONE;
    if (rollup && !(ci.kind == VGPU_DIM_TIME || ci.kind == VGPU_DIM_MICROTIME))
      fail(VGPU_ERR_INVALID, "time rollup on a non-time dimension");
    if (key.query_granularity > VGPU_TU_NONE) fail(VGPU_ERR_INVALID, "bad query granularity");
    // util::Truncator has no WEEK specialisation (src/util/time.h:52-89): the reference fails to link
    if (key.query_granularity == VGPU_TU_WEEK) fail(VGPU_ERR_UNSUPPORTED, "week granularity is not supported by the reference");
    for (uint32_t r = 0; r < key.nrules; ++r) {
      if (key.rule_granularity[r] >= VGPU_TU_NONE) fail(VGPU_ERR_INVALID, "bad rollup granularity");
      if (key.rule_granularity[r] == VGPU_TU_WEEK) fail(VGPU_ERR_UNSUPPORTED, "week granularity is not supported by the reference");
    }
  }
  for (uint32_t m = 0; m < plan->nmetrics; ++m) {
    uint32_t c = plan->metric_cols[m];
    if (c < t->ndims || c >= t->cols.size()) fail(VGPU_ERR_INVALID, "metric index out of range");
    if (t->cols[c].kind == VGPU_METRIC_HIDDEN_COUNT) fail(VGPU_ERR_INVALID, "the hidden count column cannot be selected");
  }
}

int find_hidden_count(const vgpu_table *t) {
  for (size_t c = t->ndims; c < t->cols.size(); ++c)
    if (t->cols[c].kind == VGPU_METRIC_HIDDEN_COUNT) return (int)c;
  return -1;
}

}  // namespace

#include "query_agg.inl"

// ---------------------------------------------------------------------------------------------
// select / search
// ---------------------------------------------------------------------------------------------
int vgpu_query_select(vgpu_table *t, const vgpu_rows_plan *rp, vgpu_rows **out) {
  return guard([&] {
    if (!t || !rp || !out) fail(VGPU_ERR_INVALID, "null argument");
    *out = nullptr;
    if (rp->ncols > 32) fail(VGPU_ERR_UNSUPPORTED, "select of more than 32 columns");
    if (rp->ncols && !rp->cols) fail(VGPU_ERR_INVALID, "null column list");
    for (uint32_t c = 0; c < rp->ncols; ++c)
      if (rp->cols[c] >= t->cols.size()) fail(VGPU_ERR_INVALID, "select column out of range");
    vgpu_ctx *ctx = t->ctx;
    CUDA_CK(cudaSetDevice(ctx->device));
    std::shared_lock<std::shared_mutex> table_lk = lock_table_for_query(t);
    ScopeLease lease(ctx);
    QueryScope *sc = lease.sc;
    cudaStream_t stream = sc->s0;
    CUDA_CK(cudaStreamWaitEvent(stream, t->ev_put, 0));
    vgpu_plan plan{};
    plan.nnodes = rp->nnodes; plan.nargs = rp->nargs; plan.nodes = rp->nodes; plan.args = rp->args;
    validate_plan(t, &plan);
    QueryRun q(t, &plan);
    ScanParams &P = q.planner.P;
    finish_predicate_and_prune(ctx, t, q);
    Scratch scratch(stream);
    set_work_list(t, q, scratch, stream);
    std::unique_ptr<vgpu_rows> res(new vgpu_rows());
    vgpu_rows_view &view = res->view;
    view.ncols = rp->ncols;
    view.scanned_recs = q.scanned_recs;
    view.scanned_segments = q.active.size();
    CUDA_CK(cudaEventRecord(sc->ev_begin, stream));
    uint32_t launches = 0;
    const uint32_t A = P.nactive, cps = P.tiles_per_seg;
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((P.total_tiles + kWarps - 1) / kWarps, (uint64_t)ctx->sm_count * 4));
    RowsParams2 R{};
    std::vector<uint32_t> counts(P.total_tiles);
    if (P.total_tiles) {
      R.chunk_counts = scratch.alloc<uint32_t>(P.total_tiles);
      rows_count_kernel<<<grid, kThreads, 0, stream>>>(P, R);
      CUDA_CK(cudaGetLastError());
      ++launches;
      CUDA_CK(cudaMemcpyAsync(counts.data(), R.chunk_counts, P.total_tiles * 4, cudaMemcpyDeviceToHost, stream));
      CUDA_CK(cudaStreamSynchronize(stream));
    }
    // ordinals + the reference's skip / limit rules (scan.cc:107,161), segment by segment
    std::vector<uint64_t> chunk_ord(P.total_tiles), want_lo(A), want_n(A), out_base(A);
    uint64_t skipped = 0, emitted = 0, passed = 0;
    for (uint32_t si = 0; si < A; ++si) {
      uint64_t cnt = 0;
      for (uint32_t ci = 0; ci < cps; ++ci) { chunk_ord[(uint64_t)si * cps + ci] = cnt; cnt += counts[(uint64_t)si * cps + ci]; }
      passed += cnt;
      const uint64_t skip_here = std::min<uint64_t>(cnt, rp->skip > skipped ? rp->skip - skipped : 0);
      skipped += skip_here;
      const uint64_t avail = cnt - skip_here;
      uint64_t n = avail;
      if (rp->limit > 0) n = std::min<uint64_t>(avail, emitted < rp->limit ? rp->limit - emitted : 1);
      want_lo[si] = skip_here;
      want_n[si] = n;
      out_base[si] = emitted;
      emitted += n;
    }
    view.passed_rows = passed;
    view.nrows = emitted;
    if (emitted > 0xffffffffull) fail(VGPU_ERR_UNSUPPORTED, "select of more than 2^32 rows in one call: use skip / limit");
    // cells on the device, then one pinned block on the host
    std::vector<uint64_t> off_c(rp->ncols);
    uint64_t bytes = 0;
    auto out_width = [&](uint32_t c) { const ColInfo &ci = t->cols[rp->cols[c]]; return ci.bitset ? 8u : ci.width; };
    for (uint32_t c = 0; c < rp->ncols; ++c) { off_c[c] = bytes; bytes += round_up(std::max<uint64_t>(emitted * out_width(c), 1), 64); }
    res->pool = ctx->pool;
    res->block = ctx->pool->acquire(std::max<uint64_t>(bytes, 64));
    uint8_t *hb = static_cast<uint8_t *>(res->block.first);
    if (emitted) {
      uint64_t *d_ord = scratch.alloc<uint64_t>(P.total_tiles), *d_lo = scratch.alloc<uint64_t>(A),
               *d_n = scratch.alloc<uint64_t>(A), *d_base = scratch.alloc<uint64_t>(A);
      CUDA_CK(cudaMemcpyAsync(d_ord, chunk_ord.data(), P.total_tiles * 8, cudaMemcpyHostToDevice, stream));
      CUDA_CK(cudaMemcpyAsync(d_lo, want_lo.data(), A * 8, cudaMemcpyHostToDevice, stream));
      CUDA_CK(cudaMemcpyAsync(d_n, want_n.data(), A * 8, cudaMemcpyHostToDevice, stream));
      CUDA_CK(cudaMemcpyAsync(d_base, out_base.data(), A * 8, cudaMemcpyHostToDevice, stream));
      R.chunk_ord = d_ord; R.want_lo = d_lo; R.want_n = d_n; R.out_base = d_base;
      R.out_row = scratch.alloc<uint32_t>(emitted);
      R.out_seg = scratch.alloc<uint32_t>(emitted);
      rows_write_kernel<<<grid, kThreads, 0, stream>>>(P, R);
      CUDA_CK(cudaGetLastError());
      ++launches;
      if (rp->ncols) {
        GatherParams G{};
        G.segs = t->d_segs; G.out_row = R.out_row; G.out_seg = R.out_seg; G.nrows = emitted; G.ncols = rp->ncols;
        std::vector<uint8_t *> d_out(rp->ncols);
        for (uint32_t c = 0; c < rp->ncols; ++c) {
          const ColInfo &ci = t->cols[rp->cols[c]];
          d_out[c] = scratch.alloc<uint8_t>(emitted * out_width(c));
          G.cols[c].col_off = ci.off_per_row; G.cols[c].width = ci.width; G.cols[c].bitset = ci.bitset;
          G.cols[c].bitset_idx = ci.bitset_idx; G.cols[c].out = d_out[c];
        }
        rows_gather_kernel<<<grid_for(emitted, 256, ctx->sm_count), 256, 0, stream>>>(G);
        CUDA_CK(cudaGetLastError());
        ++launches;
        for (uint32_t c = 0; c < rp->ncols; ++c)
          CUDA_CK(cudaMemcpyAsync(hb + off_c[c], d_out[c], emitted * out_width(c), cudaMemcpyDeviceToHost, stream));
      }
    }
    CUDA_CK(cudaEventRecord(sc->ev_end, stream));
    CUDA_CK(cudaStreamSynchronize(stream));
    float ms = 0;
    CUDA_CK(cudaEventElapsedTime(&ms, sc->ev_begin, sc->ev_end));
    view.gpu_ms = ms;
    view.launches = launches;
    for (uint32_t c = 0; c < rp->ncols; ++c) res->cell_ptrs.push_back(hb + off_c[c]);
    view.cells = res->cell_ptrs.empty() ? nullptr : res->cell_ptrs.data();
    *out = res.release();
  });
}

int vgpu_rows_get(const vgpu_rows *rows, vgpu_rows_view *view) {
  return guard([&] {
    if (!rows || !view) fail(VGPU_ERR_INVALID, "null argument");
    *view = rows->view;
  });
}

void vgpu_rows_free(vgpu_rows *rows) { delete rows; }

int vgpu_query_search(vgpu_table *t, const vgpu_search_plan *sp, vgpu_search **out) {
  return guard([&] {
    if (!t || !sp || !out) fail(VGPU_ERR_INVALID, "null argument");
    *out = nullptr;
    if (sp->col >= t->ndims) fail(VGPU_ERR_INVALID, "search column is not a dimension");
    const ColInfo &dc = t->cols[sp->col];
    if (type_float(dc.type)) fail(VGPU_ERR_UNSUPPORTED, "search on a floating-point dimension");
    vgpu_ctx *ctx = t->ctx;
    CUDA_CK(cudaSetDevice(ctx->device));
    std::shared_lock<std::shared_mutex> table_lk = lock_table_for_query(t);
    ScopeLease lease(ctx);
    QueryScope *sc = lease.sc;
    cudaStream_t stream = sc->s0;
    CUDA_CK(cudaStreamWaitEvent(stream, t->ev_put, 0));
    vgpu_plan plan{};
    plan.nnodes = sp->nnodes; plan.nargs = sp->nargs; plan.nodes = sp->nodes; plan.args = sp->args;
    validate_plan(t, &plan);
    QueryRun q(t, &plan);
    ScanParams &P = q.planner.P;
    finish_predicate_and_prune(ctx, t, q);
    Scratch scratch(stream);
    set_work_list(t, q, scratch, stream);
    std::unique_ptr<vgpu_search> res(new vgpu_search());
    vgpu_search_view &view = res->view;
    view.scanned_recs = q.scanned_recs;
    view.scanned_segments = q.active.size();
    CUDA_CK(cudaEventRecord(sc->ev_begin, stream));
    uint32_t launches = 0;
    const uint32_t A = P.nactive;
    res->seg_offsets.assign(A + 1, 0);
    std::vector<std::vector<std::pair<uint32_t, uint64_t>>> per_seg(A);  // (first row, code)
    // batches of consecutive segments that share one dense first-row array over the dimension's value range
    const uint64_t kMaxFirstWords = 1ull << 26;  // 256 MB
    unsigned long long *d_cursor = scratch.alloc<unsigned long long>(1);
    for (uint32_t b0 = 0; b0 < A;) {
      uint64_t omin = ~0ull, omax = 0;
      uint32_t b1 = b0;
      uint64_t rows_in_batch = 0;
      while (b1 < A) {
        const SegmentData &sd = t->segs[q.active[b1]];
        uint64_t nmin = omin, nmax = omax;
        if (sd.nrows) { nmin = std::min(omin, sd.omin[sp->col]); nmax = std::max(omax, sd.omax[sp->col]); }
        const uint64_t range = nmin <= nmax ? nmax - nmin + 1 : 1;
        if (range == 0 || range > kMaxFirstWords)
          fail(VGPU_ERR_UNSUPPORTED, "search: value range of the dimension too large for the dense first-row table");
        if (b1 > b0 && range * (b1 - b0 + 1) > kMaxFirstWords) break;
        omin = nmin; omax = nmax;
        rows_in_batch += sd.nrows;
        ++b1;
      }
      if (omin > omax) { b0 = b1; continue; }  // only empty segments
      SearchParams S{};
      S.seg_begin = b0; S.seg_end = b1;
      S.col_off = dc.off_per_row; S.width = dc.width; S.sext = dc.sext;
      S.lo = from_ordered_int(omin, dc.type);
      S.range = omax - omin + 1;
      const uint64_t nfirst = S.range * (b1 - b0);
      S.first = scratch.alloc<uint32_t>(nfirst);
      CUDA_CK(cudaMemsetAsync(S.first, 0xff, nfirst * 4, stream));
      const uint64_t tiles = (uint64_t)(b1 - b0) * P.tiles_per_seg;
      const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((tiles + kWarps - 1) / kWarps, (uint64_t)ctx->sm_count * 4));
      search_first_kernel<<<grid, kThreads, 0, stream>>>(P, S);
      CUDA_CK(cudaGetLastError());
      ++launches;
      S.cap = std::min<uint64_t>(nfirst, rows_in_batch);
      S.cursor = d_cursor;
      CUDA_CK(cudaMemsetAsync(d_cursor, 0, 8, stream));
      S.out_slot = scratch.alloc<uint32_t>(S.cap);
      S.out_code = scratch.alloc<uint64_t>(S.cap);
      S.out_first = scratch.alloc<uint32_t>(S.cap);
      search_emit_kernel<<<grid_for(nfirst, 256, ctx->sm_count), 256, 0, stream>>>(S);
      CUDA_CK(cudaGetLastError());
      ++launches;
      unsigned long long n = 0;
      CUDA_CK(cudaMemcpyAsync(&n, d_cursor, 8, cudaMemcpyDeviceToHost, stream));
      CUDA_CK(cudaStreamSynchronize(stream));
      if (n > S.cap) fail(VGPU_ERR_CUDA, "search: first-row list overflow");
      std::vector<uint32_t> h_slot(n), h_first(n);
      std::vector<uint64_t> h_code(n);
      if (n) {
        CUDA_CK(cudaMemcpyAsync(h_slot.data(), S.out_slot, n * 4, cudaMemcpyDeviceToHost, stream));
        CUDA_CK(cudaMemcpyAsync(h_first.data(), S.out_first, n * 4, cudaMemcpyDeviceToHost, stream));
        CUDA_CK(cudaMemcpyAsync(h_code.data(), S.out_code, n * 8, cudaMemcpyDeviceToHost, stream));
        CUDA_CK(cudaStreamSynchronize(stream));
      }
      for (uint64_t i = 0; i < n; ++i) per_seg[h_slot[i]].emplace_back(h_first[i], h_code[i]);
      b0 = b1;
    }
    for (uint32_t si = 0; si < A; ++si) {
      auto &v = per_seg[si];
      std::sort(v.begin(), v.end());
      for (auto &e : v) { res->first_row.push_back(e.first); res->codes.push_back(e.second); }
      res->seg_offsets[si + 1] = res->codes.size();
    }
    CUDA_CK(cudaEventRecord(sc->ev_end, stream));
    CUDA_CK(cudaStreamSynchronize(stream));
    float ms = 0;
    CUDA_CK(cudaEventElapsedTime(&ms, sc->ev_begin, sc->ev_end));
    view.gpu_ms = ms;
    view.launches = launches;
    view.nsegments = A;
    view.seg_offsets = res->seg_offsets.data();
    view.codes = res->codes.data();
    view.first_row = res->first_row.data();
    *out = res.release();
  });
}

int vgpu_search_get(const vgpu_search *r, vgpu_search_view *view) {
  return guard([&] {
    if (!r || !view) fail(VGPU_ERR_INVALID, "null argument");
    *view = r->view;
  });
}

void vgpu_search_free(vgpu_search *r) { delete r; }

int vgpu_result_get(const vgpu_result *res, vgpu_result_view *view) {
  return guard([&] {
    if (!res || !view) fail(VGPU_ERR_INVALID, "null argument");
    *view = res->view;
  });
}

void vgpu_result_free(vgpu_result *res) { delete res; }

// ---------------------------------------------------------------------------------------------
// multi-GPU
// ---------------------------------------------------------------------------------------------
int vgpu_comm_unique_id(void *unique_id_128) {
  return guard([&] {
    if (!unique_id_128) fail(VGPU_ERR_INVALID, "null argument");
    nccl_load();
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    NCCL_CK(g_nccl.GetUniqueId(&id));
    memcpy(unique_id_128, &id, sizeof(id));
  });
}

int vgpu_comm_init(vgpu_ctx *ctx, int rank, int nranks, const void *unique_id_128) {
  return guard([&] {
    if (!ctx || !unique_id_128) fail(VGPU_ERR_INVALID, "null argument");
    if (nranks < 1 || rank < 0 || rank >= nranks) fail(VGPU_ERR_INVALID, "bad rank / nranks");
    std::lock_guard<std::mutex> lk(ctx->comm_mu);
    if (ctx->comm) fail(VGPU_ERR_STATE, "communicator already initialised");
    nccl_load();
    CUDA_CK(cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, unique_id_128, sizeof(id));
    NCCL_CK(g_nccl.CommInitRank(&ctx->comm, nranks, id, rank));
    ctx->rank = rank;
    ctx->nranks = nranks;
  });
}

int vgpu_comm_destroy(vgpu_ctx *ctx) {
  return guard([&] {
    if (!ctx) fail(VGPU_ERR_INVALID, "null context");
    std::lock_guard<std::mutex> lk(ctx->comm_mu);
    if (ctx->comm) {
      CUDA_CK(cudaSetDevice(ctx->device));
      CUDA_CK(cudaDeviceSynchronize());
      NCCL_CK(g_nccl.CommDestroy(ctx->comm));
      ctx->comm = nullptr;
    }
    ctx->rank = 0;
    ctx->nranks = 1;
  });
}

}  // extern "C"
