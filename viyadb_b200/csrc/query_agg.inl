// query_agg.inl — vgpu_query_agg: the aggregate query, host side (included by vgpu.cu inside extern "C").
//
// Replaces the generated `viya_query_agg` (src/codegen/query/agg_query.cc:26-75) up to `stats.aggregated_recs =
// agg_map.size()` (src/codegen/query/scan.cc:168-247). One call enqueues, on the streams of its own QueryScope:
//
//   s0   fill group table | fused scan | [NCCL: counters, dense arrays, count-distinct pairs to their owners]
//        | count-distinct dedupe (bucket + shared-memory sets) | [NCCL: per-cell counts] | late extraction
//   s1   (count-distinct queries) early extraction of keys + finished accumulators right after the scan, and their
//        copy to the host — both overlap the dedupe on s0
//
// and synchronises with the host twice: once for the counter block (overflow flags, rows passed, number of groups),
// once for the results. Sizes the device needs before it has produced them (pair regions, hash capacity, hash
// buckets) come from per-table high-water marks; an overflow flag makes the host grow and run again.
extern "C++" {
namespace {

uint64_t hint_load(const std::atomic<uint64_t> &h) { return h.load(std::memory_order_relaxed); }
void hint_raise(std::atomic<uint64_t> &h, uint64_t v) {
  uint64_t cur = h.load(std::memory_order_relaxed);
  while (cur < v && !h.compare_exchange_weak(cur, v, std::memory_order_relaxed)) {}
}

// shared lock for a query; segment statistics and the device descriptor array are brought up to date first
std::shared_lock<std::shared_mutex> lock_table_for_query(vgpu_table *t) {
  for (;;) {
    std::shared_lock<std::shared_mutex> sl(t->mu);
    if (!t->stats_dirty && !t->descs_dirty) return sl;
    sl.unlock();
    std::lock_guard<std::mutex> put_lk(t->ctx->put_mu);
    std::unique_lock<std::shared_mutex> xl(t->mu);
    fetch_stats(t);
    upload_descs(t);
  }
}

// ---------------------------------------------------------------------------------------------
// count-distinct dedupe
// ---------------------------------------------------------------------------------------------
struct PairInput {         // where the pairs of one count-distinct metric sit when the dedupe starts
  const void *pairs;       // ragged regions of 8-byte pairs, or of Pair128 when `wide`
  const uint32_t *counts;  // [nregions]
  uint32_t nregions;
  uint32_t region_cap;
  bool wide;
  uint64_t expect;         // pairs expected (high-water mark of earlier queries), 0: unknown
};
struct DistinctTarget {
  uint8_t *distinct;       // uint32 counter of cell 0
  uint32_t stride;
  const uint64_t *lookup_keys = nullptr;  // wide pairs carrying packed group keys: the owner's table
  uint64_t lookup_mask = 0;
};
enum DedupeMode { kDedupeSmall, kDedupeFast, kDedupeWide, kDedupeGeneral };

uint64_t small_pairs_limit(const vgpu_ctx *ctx) { return ctx->test_small_pairs ? ctx->test_small_pairs : (1ull << 19); }
uint32_t smem_set_slots(const vgpu_ctx *ctx) {
  return ctx->test_set_slots ? std::min<uint32_t>(std::max<uint32_t>(ctx->test_set_slots, 64), kSmemSetSlots) : kSmemSetSlots;
}

// which path, decided before any pair exists (no host round trip between the scan and the dedupe)
DedupeMode choose_dedupe_mode(const vgpu_ctx *ctx, const vgpu_table *t, const PairInput &in, uint32_t &nbuckets) {
  const uint64_t cap_total = (uint64_t)in.nregions * in.region_cap;
  nbuckets = 0;
  if (in.wide) return kDedupeWide;
  if (cap_total <= small_pairs_limit(ctx)) return kDedupeSmall;
  if ((ctx->tune & (1u << 19)) || t->distinct_general.load(std::memory_order_relaxed)) return kDedupeGeneral;
  // buckets sized for ~55 % of the set's slots; `expect` is what earlier queries really produced
  const uint64_t expect = in.expect ? in.expect : cap_total * 4 / 5;
  const uint64_t per = std::max<uint64_t>(1, (uint64_t)smem_set_slots(ctx) * 55 / 100);
  const uint64_t nb = pow2_ceil(std::max<uint64_t>(1, (expect + per - 1) / per));
  if (nb > kMaxSmemBuckets) return kDedupeGeneral;
  nbuckets = (uint32_t)std::max<uint64_t>(nb, 2);
  return kDedupeFast;
}

// sync-free paths: everything is sized from capacities, overflow raises a flag in the counter block
void dedupe_enqueue(vgpu_ctx *ctx, QueryScope *sc, Scratch &scratch, DedupeMode mode, uint32_t nbuckets, const PairInput &in,
                    const DistinctTarget &tg, uint32_t &launches, uint32_t &paths) {
  cudaStream_t stream = sc->s0;
  const uint64_t cap_total = (uint64_t)in.nregions * in.region_cap;
  if (cap_total == 0) return;
  paths |= mode == kDedupeSmall ? VGPU_DEDUPE_SMALL : mode == kDedupeWide ? VGPU_DEDUPE_WIDE : VGPU_DEDUPE_FAST;
  if (mode == kDedupeSmall) {
    const uint64_t set_cap = pow2_ceil(std::max<uint64_t>(2 * cap_total, 1024));
    uint64_t *set = scratch.alloc<uint64_t>(set_cap);
    CUDA_CK(cudaMemsetAsync(set, 0xff, set_cap * 8, stream));
    PairsDedupeParams<uint64_t> D{};
    D.pairs = static_cast<const uint64_t *>(in.pairs);
    D.counts = in.counts;
    D.nregions = in.nregions;
    D.region_cap = in.region_cap;
    D.set = set;
    D.set_mask = set_cap - 1;
    D.distinct = tg.distinct;
    D.stride = tg.stride;
    pairs_dedupe_kernel<uint64_t><<<(int)std::max<uint32_t>(1, std::min<uint32_t>(in.nregions, ctx->sm_count * 8)), 256, 0, stream>>>(D);
    CUDA_CK(cudaGetLastError());
    ++launches;
    return;
  }
  if (mode == kDedupeWide) {
    const uint64_t set_cap = pow2_ceil(std::max<uint64_t>(2 * cap_total, 1024));  // by capacity: a full set would never terminate
    Pair128 *set = scratch.alloc<Pair128>(set_cap);
    CUDA_CK(cudaMemsetAsync(set, 0xff, set_cap * sizeof(Pair128), stream));
    unsigned long long *seen = scratch.alloc<unsigned long long>(1);
    CUDA_CK(cudaMemsetAsync(seen, 0, 8, stream));
    PairsDedupeParams<Pair128> D{};
    D.pairs = static_cast<const Pair128 *>(in.pairs);
    D.counts = in.counts;
    D.nregions = in.nregions;
    D.region_cap = in.region_cap;
    D.set = set;
    D.set_mask = set_cap - 1;
    D.distinct = tg.distinct;
    D.stride = tg.stride;
    D.lookup_keys = tg.lookup_keys;
    D.lookup_mask = tg.lookup_mask;
    D.sentinel_seen = seen;
    pairs_dedupe_kernel<Pair128><<<(int)std::max<uint32_t>(1, std::min<uint32_t>(in.nregions, ctx->sm_count * 8)), 256, 0, stream>>>(D);
    CUDA_CK(cudaGetLastError());
    ++launches;
    return;
  }
  // fast path: hash buckets small enough for a shared-memory set
  const uint32_t nslots = smem_set_slots(ctx);
  const uint64_t expect = in.expect ? in.expect : cap_total * 4 / 5;
  const uint32_t limit = nslots - nslots / 4;
  const uint32_t bucket_cap = (uint32_t)std::min<uint64_t>(limit, expect / nbuckets + expect / nbuckets / 4 + 512);
  uint32_t *cursors = scratch.alloc<uint32_t>(nbuckets);
  uint64_t *buckets = scratch.alloc<uint64_t>((uint64_t)nbuckets * bucket_cap);
  CUDA_CK(cudaMemsetAsync(cursors, 0, nbuckets * 4, stream));
  PairsBucketParams A{};
  A.pairs = static_cast<const uint64_t *>(in.pairs);
  A.counts = in.counts;
  A.nregions = in.nregions;
  A.region_cap = in.region_cap;
  A.nbuckets = nbuckets;
  A.bucket_cap = bucket_cap;
  A.cursors = cursors;
  A.out = buckets;
  A.flags = sc->d_counters + kCBucketOver;
  pairs_bucket_kernel<<<(int)std::max<uint32_t>(1, std::min<uint32_t>(in.nregions, ctx->sm_count)), kBucketThreads, 2 * nbuckets * 4, stream>>>(A);
  CUDA_CK(cudaGetLastError());
  ++launches;
  PairsSmemDedupeParams D{};
  D.buckets = buckets;
  D.cursors = cursors;
  D.nbuckets = nbuckets;
  D.bucket_cap = bucket_cap;
  D.nslots = nslots;
  D.limit = limit;
  D.distinct = tg.distinct;
  D.stride = tg.stride;
  D.flags = sc->d_counters + kCSetOver;
  pairs_dedupe_smem_kernel<<<(int)std::min<uint32_t>(nbuckets, ctx->sm_count * 2), kSmemSetThreads, nslots * 8, stream>>>(D);
  CUDA_CK(cudaGetLastError());
  ++launches;
}

// General path, any size or skew (host knows `total`; synchronises): the pairs are hash-partitioned into buckets
// whose open-addressing sets stay in L2, then each bucket is inserted into one reused global set.
void dedupe_general(vgpu_ctx *ctx, QueryScope *sc, Scratch &scratch, const PairInput &in, uint64_t total, const DistinctTarget &tg,
                    uint32_t &launches, uint32_t &paths) {
  if (total == 0) return;
  paths |= VGPU_DEDUPE_GENERAL;
  cudaStream_t stream = sc->s0;
  const uint64_t bucket_pairs = ctx->test_bucket_pairs ? ctx->test_bucket_pairs : (1ull << 21);  // 2^22-slot set = 32 MB: stays in L2
  const uint32_t B = (uint32_t)std::min<uint64_t>(pow2_ceil((total + bucket_pairs - 1) / bucket_pairs), kMaxBuckets);
  PairsDedupeParams<uint64_t> D{};
  D.distinct = tg.distinct;
  D.stride = tg.stride;
  if (B <= 1) {
    const uint64_t set_cap = pow2_ceil(std::max<uint64_t>(2 * total, 1024));
    uint64_t *set = scratch.alloc<uint64_t>(set_cap);
    CUDA_CK(cudaMemsetAsync(set, 0xff, set_cap * 8, stream));
    D.pairs = static_cast<const uint64_t *>(in.pairs);
    D.counts = in.counts;
    D.nregions = in.nregions;
    D.region_cap = in.region_cap;
    D.set = set;
    D.set_mask = set_cap - 1;
    pairs_dedupe_kernel<uint64_t><<<(int)std::max<uint32_t>(1, std::min<uint32_t>(in.nregions, ctx->sm_count * 8)), 256, 0, stream>>>(D);
    CUDA_CK(cudaGetLastError());
    ++launches;
    return;
  }
  paths |= VGPU_DEDUPE_PARTITIONED;
  uint64_t bucket_cap = total / B + total / B / 8 + 8192;
  unsigned long long *cursors = scratch.alloc<unsigned long long>(kMaxBuckets + 1);
  std::vector<unsigned long long> h_cursors(kMaxBuckets + 1);
  uint64_t *buckets = nullptr;
  for (int attempt = 0; attempt < 2; ++attempt) {
    buckets = scratch.alloc<uint64_t>(bucket_cap * B);
    CUDA_CK(cudaMemsetAsync(cursors, 0, (kMaxBuckets + 1) * sizeof(unsigned long long), stream));
    PairsPartitionParams A{};
    A.pairs = static_cast<const uint64_t *>(in.pairs);
    A.counts = in.counts;
    A.total = total;
    A.nregions = in.nregions;
    A.region_cap = in.region_cap;
    A.nbuckets = B;
    A.shift = 40;
    A.bucket_cap = bucket_cap;
    A.cursors = cursors;
    A.out = buckets;
    A.overflow = cursors + kMaxBuckets;
    pairs_partition_kernel<<<(int)std::max<uint32_t>(1, std::min<uint32_t>(A.nregions, ctx->sm_count * 8)), 256, 0, stream>>>(A);
    CUDA_CK(cudaGetLastError());
    ++launches;
    CUDA_CK(cudaMemcpyAsync(h_cursors.data(), cursors, (kMaxBuckets + 1) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    CUDA_CK(cudaStreamSynchronize(stream));
    if (h_cursors[kMaxBuckets] == 0) break;
    if (attempt == 1) fail(VGPU_ERR_CUDA, "count-distinct partitioning overflowed twice");
    bucket_cap = 0;  // a skewed hash bucket: size every bucket for the largest one and scatter again
    for (uint32_t b = 0; b < B; ++b) bucket_cap = std::max<uint64_t>(bucket_cap, h_cursors[b]);
  }
  uint64_t max_n = 0;
  for (uint32_t b = 0; b < B; ++b) max_n = std::max<uint64_t>(max_n, h_cursors[b]);
  const uint64_t set_cap_max = pow2_ceil(std::max<uint64_t>(2 * max_n, 1024));
  uint64_t *set = scratch.alloc<uint64_t>(set_cap_max);
  for (uint32_t b = 0; b < B; ++b) {
    const uint64_t n = h_cursors[b];
    if (n == 0) continue;
    const uint64_t set_cap = pow2_ceil(std::max<uint64_t>(2 * n, 1024));
    CUDA_CK(cudaMemsetAsync(set, 0xff, set_cap * 8, stream));
    D.pairs = buckets + (uint64_t)b * bucket_cap;
    D.counts = nullptr;
    D.n = n;
    D.set = set;
    D.set_mask = set_cap - 1;
    pairs_dedupe_kernel<uint64_t><<<grid_for(n, 256, ctx->sm_count), 256, 0, stream>>>(D);
    CUDA_CK(cudaGetLastError());
    ++launches;
  }
}

// ---------------------------------------------------------------------------------------------
// multi-GPU
// ---------------------------------------------------------------------------------------------
// Count-distinct pairs to their owner ranks. The scan left them owner-major: slab o = regions [o * nper, (o+1) * nper)
// of `cap` elements each. Fixed-size slabs travel whole (no count round trip through the host); the counts follow
// in the same group. Must be called inside ncclGroupStart / End. recv / recv_counts: G slabs / G * nper counts.
void exchange_pairs(vgpu_ctx *ctx, cudaStream_t stream, const void *send, const uint32_t *send_counts, uint32_t nper, uint32_t cap,
                    uint32_t elem, void *recv, uint32_t *recv_counts) {
  const int G = ctx->nranks, me = ctx->rank;
  const uint64_t slab = (uint64_t)nper * cap * elem;
  const uint8_t *sb = static_cast<const uint8_t *>(send);
  uint8_t *rb = static_cast<uint8_t *>(recv);
  for (int r = 0; r < G; ++r) {
    if (r == me) {
      CUDA_CK(cudaMemcpyAsync(rb + (uint64_t)r * slab, sb + (uint64_t)r * slab, slab, cudaMemcpyDeviceToDevice, stream));
      CUDA_CK(cudaMemcpyAsync(recv_counts + (uint64_t)r * nper, send_counts + (uint64_t)r * nper, nper * 4, cudaMemcpyDeviceToDevice, stream));
      continue;
    }
    NCCL_CK(g_nccl.Send(sb + (uint64_t)r * slab, slab, ncclUint8, r, ctx->comm, stream));
    NCCL_CK(g_nccl.Recv(rb + (uint64_t)r * slab, slab, ncclUint8, r, ctx->comm, stream));
    NCCL_CK(g_nccl.Send(send_counts + (uint64_t)r * nper, nper, ncclUint32, r, ctx->comm, stream));
    NCCL_CK(g_nccl.Recv(recv_counts + (uint64_t)r * nper, nper, ncclUint32, r, ctx->comm, stream));
  }
}

// bucket sizes of every rank: matrix[r * G + o] = entries rank r holds for owner o
std::vector<uint64_t> exchange_counts(vgpu_ctx *ctx, cudaStream_t stream, unsigned long long *d_cursors, Scratch &scratch) {
  const int G = ctx->nranks;
  uint64_t *d_matrix = scratch.alloc<uint64_t>((uint64_t)G * G);
  NCCL_CK(g_nccl.AllGather(d_cursors, d_matrix, G, ncclUint64, ctx->comm, stream));
  std::vector<uint64_t> matrix((size_t)G * G);
  CUDA_CK(cudaMemcpyAsync(matrix.data(), d_matrix, matrix.size() * 8, cudaMemcpyDeviceToHost, stream));
  CUDA_CK(cudaStreamSynchronize(stream));
  return matrix;
}

// all-to-all of one bucketed array: rank `me` sends bucket o (matrix[me][o] elements) to rank o and
// receives matrix[r][me] elements from every r into recv + recv_off[r]. Must be called inside a group.
void exchange_array(vgpu_ctx *ctx, cudaStream_t stream, const void *send, uint64_t bucket_cap, uint32_t elem, void *recv,
                    const std::vector<uint64_t> &matrix, const std::vector<uint64_t> &recv_off) {
  const int G = ctx->nranks, me = ctx->rank;
  const uint8_t *sb = static_cast<const uint8_t *>(send);
  uint8_t *rb = static_cast<uint8_t *>(recv);
  for (int r = 0; r < G; ++r) {
    const uint64_t ns = matrix[(size_t)me * G + r], nr = matrix[(size_t)r * G + me];
    if (r == me) {
      if (ns) CUDA_CK(cudaMemcpyAsync(rb + recv_off[r] * elem, sb + (uint64_t)r * bucket_cap * elem, ns * elem,
                                      cudaMemcpyDeviceToDevice, stream));
      continue;
    }
    if (ns) NCCL_CK(g_nccl.Send(sb + (uint64_t)r * bucket_cap * elem, ns * elem, ncclUint8, r, ctx->comm, stream));
    if (nr) NCCL_CK(g_nccl.Recv(rb + recv_off[r] * elem, nr * elem, ncclUint8, r, ctx->comm, stream));
  }
}

// Hashed group tables across ranks: every (packed key, partial accumulators) record travels to the rank that owns
// the key, owners merge with the same Update() as the scan (store.cc:131-161), `on_owned` then runs on the owner's
// table (count-distinct: the pairs of the owned keys are deduplicated against it), and the owned groups are
// all-gathered so that every rank returns the full result. On return P / acc_ptrs / acc_cells describe a table that
// holds ALL groups of all ranks.
template <class OnOwned>
void nccl_merge_hash(vgpu_ctx *ctx, QueryScope *sc, QueryRun &q, ScanParams &P, std::vector<void *> &acc_ptrs, Scratch &scratch,
                     uint64_t &acc_cells, uint32_t &launches, OnOwned &&on_owned) {
  const int G = ctx->nranks, me = ctx->rank;
  cudaStream_t stream = sc->s0;
  const size_t nm = q.accs.size();
  // how many groups does this rank hold?
  ExtractParams E{};
  E.ncells = acc_cells;
  E.hash_mode = 1;
  E.hkeys = P.hkeys;
  E.present = P.present;
  E.hkey_stride = P.hkey_stride;
  E.present_stride = P.present_stride;
  E.count_only = 1;
  unsigned long long *d_n = scratch.alloc<unsigned long long>(1);
  CUDA_CK(cudaMemsetAsync(d_n, 0, 8, stream));
  E.counter = d_n;
  extract_groups_kernel<<<grid_for(acc_cells, 256, ctx->sm_count), 256, 0, stream>>>(E);
  CUDA_CK(cudaGetLastError());
  ++launches;
  uint64_t n_local = 0;
  CUDA_CK(cudaMemcpyAsync(&n_local, d_n, 8, cudaMemcpyDeviceToHost, stream));
  CUDA_CK(cudaStreamSynchronize(stream));

  auto run_exchange = [&](const uint64_t *keys, uint64_t nslots, uint64_t sentinel_slot, const uint8_t *sentinel_present,
                          const std::vector<void *> &src, uint64_t n_hint, uint64_t *&out_keys, std::vector<void *> &out_acc) -> uint64_t {
    const uint64_t bucket_cap = std::max<uint64_t>(n_hint, 1);
    unsigned long long *cursors = scratch.alloc<unsigned long long>(kMaxParts);
    CUDA_CK(cudaMemsetAsync(cursors, 0, kMaxParts * sizeof(unsigned long long), stream));
    uint64_t *send_keys = scratch.alloc<uint64_t>(bucket_cap * G);
    std::vector<void *> send_acc(nm);
    PartitionParams A{};
    A.keys = keys;
    A.nslots = nslots;
    A.sentinel_slot = sentinel_slot;
    A.sentinel_present = sentinel_present;
    A.nparts = (uint32_t)G;
    A.owner_shift = 0;
    A.bucket_cap = bucket_cap;
    A.cursors = cursors;
    A.out_keys = send_keys;
    A.npay = (uint32_t)nm;
    for (size_t m = 0; m < nm; ++m) {
      send_acc[m] = scratch.alloc<uint8_t>(bucket_cap * G * q.accs[m].acc_width);
      A.pay_width[m] = q.accs[m].acc_width;
      A.pay_src[m] = src[m];
      A.pay_dst[m] = send_acc[m];
    }
    partition_table_kernel<<<grid_for(nslots + 1, 256 * kPartPerThread, ctx->sm_count, 8), 256, 0, stream>>>(A);
    CUDA_CK(cudaGetLastError());
    ++launches;
    std::vector<uint64_t> matrix = exchange_counts(ctx, stream, cursors, scratch);
    std::vector<uint64_t> recv_off(G);
    uint64_t total = 0;
    for (int r = 0; r < G; ++r) { recv_off[r] = total; total += matrix[(size_t)r * G + me]; }
    out_keys = scratch.alloc<uint64_t>(total);
    out_acc.resize(nm);
    for (size_t m = 0; m < nm; ++m) out_acc[m] = scratch.alloc<uint8_t>(total * q.accs[m].acc_width);
    NCCL_CK(g_nccl.GroupStart());
    exchange_array(ctx, stream, send_keys, bucket_cap, 8, out_keys, matrix, recv_off);
    for (size_t m = 0; m < nm; ++m)
      exchange_array(ctx, stream, send_acc[m], bucket_cap, q.accs[m].acc_width, out_acc[m], matrix, recv_off);
    NCCL_CK(g_nccl.GroupEnd());
    return total;
  };

  // build a fresh table from records
  auto build_table = [&](const uint64_t *keys, const std::vector<void *> &src, uint64_t n, uint64_t cap) {
    uint64_t block_bytes = 0;
    auto carve = [&](uint64_t bytes) { uint64_t o = block_bytes; block_bytes += round_up(bytes, 256); return o; };
    const uint64_t o_h = carve(cap * 8), o_p = carve(16);
    std::vector<uint64_t> o_a(nm);
    for (size_t m = 0; m < nm; ++m) o_a[m] = carve((cap + 1) * q.accs[m].acc_width);
    uint8_t *block = scratch.alloc<uint8_t>(block_bytes);
    MergeParams M{};
    M.keys = keys;
    M.n = n;
    M.nmets = (uint32_t)nm;
    M.hkeys = reinterpret_cast<uint64_t *>(block + o_h);
    M.hmask = cap - 1;
    M.present = block + o_p;
    M.max_probe = (uint32_t)std::min<uint64_t>(cap, 1u << 20);
    M.overflow = scratch.alloc<unsigned long long>(1);
    CUDA_CK(cudaMemsetAsync(M.overflow, 0, 8, stream));
    fill64(stream, ctx->sm_count, M.hkeys, cap, kEmptyKey);
    CUDA_CK(cudaMemsetAsync(M.present, 0, 16, stream));
    for (size_t m = 0; m < nm; ++m) {
      // per-key distinct counts are disjoint between owners: merging them is a plain add
      M.ops[m] = q.accs[m].op == A_DISTINCT ? (uint32_t)A_ADD32 : q.accs[m].op;
      M.widths[m] = q.accs[m].acc_width;
      M.src[m] = src[m];
      M.acc[m] = block + o_a[m];
      if (q.accs[m].acc_width == 4) launches += fill32(stream, ctx->sm_count, M.acc[m], cap + 1, (uint32_t)q.accs[m].init);
      else launches += fill64(stream, ctx->sm_count, M.acc[m], cap + 1, q.accs[m].init);
    }
    if (n) {
      merge_records_kernel<<<grid_for(n, 256, ctx->sm_count), 256, 0, stream>>>(M);
      CUDA_CK(cudaGetLastError());
      ++launches;
    }
    P.hkeys = M.hkeys;
    P.hkey_stride = 8;
    P.hmask = M.hmask;
    P.present = M.present;
    P.present_stride = 1;
    for (size_t m = 0; m < nm; ++m) {
      acc_ptrs[m] = M.acc[m];
      P.mets[m].acc = M.acc[m];
      P.mets[m].stride = q.accs[m].acc_width;
    }
    acc_cells = cap + 1;
  };

  // 1. records to their owners, owners merge
  uint64_t *own_keys = nullptr;
  std::vector<void *> own_acc;
  const uint64_t n_owned_in = run_exchange(P.hkeys, P.hmask + 1, P.hmask + 1, P.present, acc_ptrs, n_local, own_keys, own_acc);
  build_table(own_keys, own_acc, n_owned_in, pow2_ceil(std::max<uint64_t>(2 * n_owned_in, 1024)));
  on_owned();

  // 2. all-gather the owned groups: every rank broadcasts the live records of its (merged) table
  uint64_t *send_keys = nullptr;
  std::vector<void *> send_acc;
  {
    // the partition kernel with ONE bucket compacts the merged table
    const uint64_t cap = P.hmask + 1;
    unsigned long long *cursor = scratch.alloc<unsigned long long>(kMaxParts);
    CUDA_CK(cudaMemsetAsync(cursor, 0, kMaxParts * sizeof(unsigned long long), stream));
    const uint64_t bucket_cap = std::max<uint64_t>(n_owned_in, 1);
    send_keys = scratch.alloc<uint64_t>(bucket_cap);
    send_acc.resize(nm);
    PartitionParams A{};
    A.keys = P.hkeys;
    A.nslots = cap;
    A.sentinel_slot = cap;
    A.sentinel_present = P.present;
    A.nparts = 1;
    A.owner_shift = 0;
    A.bucket_cap = bucket_cap;
    A.cursors = cursor;
    A.out_keys = send_keys;
    A.npay = (uint32_t)nm;
    for (size_t m = 0; m < nm; ++m) {
      send_acc[m] = scratch.alloc<uint8_t>(bucket_cap * q.accs[m].acc_width);
      A.pay_width[m] = q.accs[m].acc_width;
      A.pay_src[m] = acc_ptrs[m];
      A.pay_dst[m] = send_acc[m];
    }
    partition_table_kernel<<<grid_for(cap + 1, 256 * kPartPerThread, ctx->sm_count, 8), 256, 0, stream>>>(A);
    CUDA_CK(cudaGetLastError());
    ++launches;
    uint64_t *d_all = scratch.alloc<uint64_t>(G);
    NCCL_CK(g_nccl.AllGather(cursor, d_all, 1, ncclUint64, ctx->comm, stream));
    std::vector<uint64_t> counts(G);
    CUDA_CK(cudaMemcpyAsync(counts.data(), d_all, G * 8, cudaMemcpyDeviceToHost, stream));
    CUDA_CK(cudaStreamSynchronize(stream));
    std::vector<uint64_t> off(G);
    uint64_t total = 0;
    for (int r = 0; r < G; ++r) { off[r] = total; total += counts[r]; }
    uint64_t *all_keys = scratch.alloc<uint64_t>(total);
    std::vector<void *> all_acc(nm);
    for (size_t m = 0; m < nm; ++m) all_acc[m] = scratch.alloc<uint8_t>(total * q.accs[m].acc_width);
    NCCL_CK(g_nccl.GroupStart());
    for (int r = 0; r < G; ++r) {
      if (counts[r] == 0) continue;
      NCCL_CK(g_nccl.Broadcast(send_keys, all_keys + off[r], counts[r] * 8, ncclUint8, r, ctx->comm, stream));
      for (size_t m = 0; m < nm; ++m) {
        const uint32_t w = q.accs[m].acc_width;
        NCCL_CK(g_nccl.Broadcast(send_acc[m], static_cast<uint8_t *>(all_acc[m]) + off[r] * w, counts[r] * w, ncclUint8, r,
                                 ctx->comm, stream));
      }
    }
    NCCL_CK(g_nccl.GroupEnd());
    // 3. the full table, identical content on every rank (keys are disjoint between owners: plain inserts)
    build_table(all_keys, all_acc, total, pow2_ceil(std::max<uint64_t>(2 * total, 1024)));
  }
}

// Wide key tuples (hash_mode 2) across ranks: every rank compacts its live (key tuple, partial accumulators) records,
// all records are gathered on every rank (broadcast per rank: wide tables are the rare case, no owner step) and merged
// into a fresh wide table with the same Update(). On return P / acc_ptrs / acc_cells describe the merged table.
void nccl_merge_wide(vgpu_ctx *ctx, QueryScope *sc, QueryRun &q, ScanParams &P, std::vector<void *> &acc_ptrs, Scratch &scratch,
                     uint64_t &acc_cells, uint64_t local_bound, uint32_t &launches) {
  const int G = ctx->nranks;
  cudaStream_t stream = sc->s0;
  const size_t nm = q.accs.size();
  const uint32_t nkeys = std::max<uint32_t>(P.nkeys, 1);
  const uint64_t nslots = P.hmask + 1;
  const uint64_t cap = std::max<uint64_t>(1, std::min(nslots, local_bound));
  WideCompactParams W{};
  W.wstate = P.wstate;
  W.wkeys = P.wkeys;
  W.nslots = nslots;
  W.nkeys = nkeys;
  W.nmets = (uint32_t)nm;
  W.cap = cap;
  W.cursor = scratch.alloc<unsigned long long>(1);
  CUDA_CK(cudaMemsetAsync(W.cursor, 0, 8, stream));
  W.out_keys = scratch.alloc<uint64_t>(cap * nkeys);
  std::vector<void *> send_acc(nm);
  for (size_t m = 0; m < nm; ++m) {
    send_acc[m] = scratch.alloc<uint8_t>(cap * q.accs[m].acc_width);
    W.widths[m] = q.accs[m].acc_width;
    W.src[m] = acc_ptrs[m];
    W.dst[m] = send_acc[m];
  }
  wide_compact_kernel<<<grid_for(nslots, 256, ctx->sm_count), 256, 0, stream>>>(W);
  CUDA_CK(cudaGetLastError());
  ++launches;
  uint64_t *d_all = scratch.alloc<uint64_t>(G);
  NCCL_CK(g_nccl.AllGather(W.cursor, d_all, 1, ncclUint64, ctx->comm, stream));
  std::vector<uint64_t> counts(G);
  CUDA_CK(cudaMemcpyAsync(counts.data(), d_all, G * 8, cudaMemcpyDeviceToHost, stream));
  CUDA_CK(cudaStreamSynchronize(stream));
  std::vector<uint64_t> off(G);
  uint64_t total = 0;
  for (int r = 0; r < G; ++r) { off[r] = total; total += counts[r]; }
  uint64_t *all_keys = scratch.alloc<uint64_t>(total * nkeys);
  std::vector<void *> all_acc(nm);
  for (size_t m = 0; m < nm; ++m) all_acc[m] = scratch.alloc<uint8_t>(total * q.accs[m].acc_width);
  NCCL_CK(g_nccl.GroupStart());
  for (int r = 0; r < G; ++r) {
    if (counts[r] == 0) continue;
    NCCL_CK(g_nccl.Broadcast(W.out_keys, all_keys + off[r] * nkeys, counts[r] * nkeys * 8, ncclUint8, r, ctx->comm, stream));
    for (size_t m = 0; m < nm; ++m) {
      const uint32_t w = q.accs[m].acc_width;
      NCCL_CK(g_nccl.Broadcast(send_acc[m], static_cast<uint8_t *>(all_acc[m]) + off[r] * w, counts[r] * w, ncclUint8, r, ctx->comm, stream));
    }
  }
  NCCL_CK(g_nccl.GroupEnd());
  // the merged table
  const uint64_t mcap = pow2_ceil(std::max<uint64_t>(2 * total, 1024));
  uint64_t block_bytes = 0;
  auto carve = [&](uint64_t bytes) { uint64_t o = block_bytes; block_bytes += round_up(bytes, 256); return o; };
  const uint64_t o_state = carve(mcap * 4), o_keys = carve(mcap * 8 * nkeys);
  std::vector<uint64_t> o_a(nm);
  for (size_t m = 0; m < nm; ++m) o_a[m] = carve((mcap + 1) * q.accs[m].acc_width);
  uint8_t *block = scratch.alloc<uint8_t>(block_bytes);
  ScanParams T{};
  T.nkeys = P.nkeys;
  T.hmask = mcap - 1;
  T.max_probe = (uint32_t)std::min<uint64_t>(mcap, 1u << 20);
  T.wstate = reinterpret_cast<uint32_t *>(block + o_state);
  T.wkeys = reinterpret_cast<uint64_t *>(block + o_keys);
  CUDA_CK(cudaMemsetAsync(T.wstate, 0, mcap * 4, stream));
  WideMergeParams M{};
  M.keys = all_keys;
  M.n = total;
  M.nmets = (uint32_t)nm;
  M.overflow = scratch.alloc<unsigned long long>(1);
  CUDA_CK(cudaMemsetAsync(M.overflow, 0, 8, stream));
  for (size_t m = 0; m < nm; ++m) {
    M.ops[m] = q.accs[m].op;
    M.widths[m] = q.accs[m].acc_width;
    M.src[m] = all_acc[m];
    M.acc[m] = block + o_a[m];
    if (q.accs[m].acc_width == 4) launches += fill32(stream, ctx->sm_count, M.acc[m], mcap + 1, (uint32_t)q.accs[m].init);
    else launches += fill64(stream, ctx->sm_count, M.acc[m], mcap + 1, q.accs[m].init);
  }
  if (total) {
    wide_merge_kernel<<<grid_for(total, 256, ctx->sm_count), 256, 0, stream>>>(T, M);
    CUDA_CK(cudaGetLastError());
    ++launches;
  }
  P.wstate = T.wstate;
  P.wkeys = T.wkeys;
  P.hmask = T.hmask;
  for (size_t m = 0; m < nm; ++m) {
    acc_ptrs[m] = M.acc[m];
    P.mets[m].acc = M.acc[m];
    P.mets[m].stride = q.accs[m].acc_width;
  }
  acc_cells = mcap + 1;
}

}  // namespace
}  // extern "C++"

int vgpu_query_agg(vgpu_table *t, const vgpu_plan *plan, vgpu_result **out) {
  return guard([&] {
    if (!t || !plan || !out) fail(VGPU_ERR_INVALID, "null argument");
    *out = nullptr;
    vgpu_ctx *ctx = t->ctx;
    CUDA_CK(cudaSetDevice(ctx->device));
    validate_plan(t, plan);
    // several ranks: one collective sequence at a time (every rank issues the same one)
    std::unique_lock<std::mutex> comm_lk(ctx->comm_mu, std::defer_lock);
    if (ctx->nranks > 1) comm_lk.lock();
    std::shared_lock<std::shared_mutex> table_lk = lock_table_for_query(t);
    ScopeLease lease(ctx);   // destroyed (streams drained) before the table lock goes
    QueryScope *sc = lease.sc;
    cudaStream_t stream = sc->s0;
    CUDA_CK(cudaStreamWaitEvent(stream, t->ev_put, 0));  // device work of earlier puts

    QueryRun q(t, plan);
    Planner &pl = q.planner;
    ScanParams &P = pl.P;
    finish_predicate_and_prune(ctx, t, q);
    const int G = ctx->nranks;

    // ---- keys ----
    P.nkeys = plan->nkeys;
    q.ranges.resize(plan->nkeys);
    for (uint32_t k = 0; k < plan->nkeys; ++k) {
      const vgpu_key &key = plan->keys[k];
      const ColInfo &ci = t->cols[key.col];
      if (ci.bitset) fail(VGPU_ERR_INVALID, "bitset column as a key");
      KeySpec &ks = P.keys[k];
      ks.slot = (uint8_t)pl.slot_of(key.col);
      ks.rollup = key.nrules > 0 || key.query_granularity != VGPU_TU_NONE;
      ks.micro = ci.kind == VGPU_DIM_MICROTIME;
      ks.nrules = (uint8_t)key.nrules;
      ks.query_unit = (uint8_t)key.query_granularity;
      for (uint32_t r = 0; r < key.nrules; ++r) {
        ks.rule_unit[r] = (uint8_t)key.rule_granularity[r];
        ks.rule_boundary[r] = key.rule_boundary[r];
      }
      // -0.0 groups with +0.0 whatever the table layout (KeyEqual uses ==, store.cc:46-63)
      ks.fzero = type_float(ci.type);
    }
    // value range of every key over the active segments (ordered domain). With several ranks the
    // ranges — hence the cell numbering / key packing — must be the same everywhere: min-reduce them.
    std::vector<uint64_t> kmin(plan->nkeys, ~0ull), kmax(plan->nkeys, 0ull);
    for (uint32_t k = 0; k < plan->nkeys; ++k) {
      const uint32_t col = plan->keys[k].col;
      for (uint32_t s : q.active) {
        const SegmentData &sd = t->segs[s];
        if (sd.nrows == 0) continue;
        kmin[k] = std::min(kmin[k], sd.omin[col]);
        kmax[k] = std::max(kmax[k], sd.omax[col]);
      }
    }
    uint64_t global_active_rows = q.active_rows;
    uint64_t max_active_rows = q.active_rows;  // of any rank: what per-rank buffers are sized by
    if (G > 1) {
      uint64_t *h = sc->h_plan;
      const size_t n = 2 * plan->nkeys + 1;
      for (uint32_t k = 0; k < plan->nkeys; ++k) { h[2 * k] = kmin[k]; h[2 * k + 1] = ~kmax[k]; }
      h[2 * plan->nkeys] = ~q.active_rows;  // min of complements == complement of the max
      CUDA_CK(cudaMemcpyAsync(sc->d_plan, h, n * 8, cudaMemcpyHostToDevice, stream));
      NCCL_CK(g_nccl.AllReduce(sc->d_plan, sc->d_plan, n, ncclUint64, ncclMin, ctx->comm, stream));
      CUDA_CK(cudaMemcpyAsync(h, sc->d_plan, n * 8, cudaMemcpyDeviceToHost, stream));
      CUDA_CK(cudaStreamSynchronize(stream));
      for (uint32_t k = 0; k < plan->nkeys; ++k) { kmin[k] = h[2 * k]; kmax[k] = ~h[2 * k + 1]; }
      max_active_rows = ~h[2 * plan->nkeys];
      global_active_rows = max_active_rows * (uint64_t)G;  // upper bound, same on every rank
    }
    std::vector<uint64_t> tdict_values;
    for (uint32_t k = 0; k < plan->nkeys; ++k) {
      const ColInfo &ci = t->cols[plan->keys[k].col];
      const KeySpec &ks = P.keys[k];
      KeyRange kr{0, 1};
      if (type_float(ci.type)) {
        kr.lo = 0;
        kr.range = ci.width == 4 ? (1ull << 32) : 0;  // keyed by raw bits
      } else if (kmin[k] <= kmax[k]) {
        uint64_t lo = from_ordered_int(kmin[k], ci.type), hi = from_ordered_int(kmax[k], ci.type);
        bool dict = false;
        if (ks.rollup && P.tdict.npieces == 0 && !(ctx->tune & (1u << 21)) &&
            build_time_dict(plan->keys[k], ks.micro != 0, lo, hi, P.tdict, tdict_values)) {
          // the key is the rank of its rolled-up value among the attainable ones (TimeDict)
          P.tdict.key = k;
          lo = 0;
          hi = tdict_values.size() - 1;
          dict = true;
        }
        if (ks.rollup && !dict) {  // truncation only moves values down, at most to the start of their year
          if (ks.micro) lo = host_trunc_year_seconds(lo / 1000000ull) * 1000000ull;
          else lo = host_trunc_year_seconds(lo);
        }
        // A top-level conjunction restricts what a key can be for passing rows: tighten the key domain
        // (unsigned keys of at most 4 bytes, no rollup; leaf arguments are raw zero-extended values).
        uint64_t lut = 0;
        if (P.conj && !ks.rollup && !type_signed(ci.type) && ci.width <= 4 && !(ctx->tune & 16384u)) {
          lut = tighten_key_domain(P, ks.slot, lo, hi);   // planner.h
        }
        kr.lo = lo;
        kr.range = hi - lo + 1;  // wraps to 0 for the full 64-bit domain
        if (lut) { kr.lo = 0; kr.range = (uint64_t)__builtin_popcountll(lut); }
        P.keys[k].lut = lut;
      }
      q.ranges[k] = kr;
    }

    // ---- metrics ----
    const int hidden_col = find_hidden_count(t);
    if (plan->need_hidden_count && hidden_col < 0)
      fail(VGPU_ERR_INVALID, "plan needs the hidden count column but the table has none");
    for (uint32_t m = 0; m < plan->nmetrics; ++m) {
      q.acc_cols.push_back(plan->metric_cols[m]);
      q.accs.push_back(acc_for(t->cols[plan->metric_cols[m]]));
    }
    if (plan->need_hidden_count) {
      q.acc_cols.push_back((uint32_t)hidden_col);
      q.accs.push_back(acc_for(t->cols[hidden_col]));
    }
    P.nmetrics = (uint32_t)q.accs.size();
    P.ndistinct = 0;
    bool ids64 = false;
    for (uint32_t m = 0; m < P.nmetrics; ++m) {
      P.mets[m].slot = (uint8_t)pl.slot_of(q.acc_cols[m]);
      P.mets[m].op = (uint8_t)q.accs[m].op;
      if (q.accs[m].op == A_DISTINCT) {
        if (P.ndistinct >= kMaxDistinct) fail(VGPU_ERR_UNSUPPORTED, "too many count-distinct metrics in one query");
        P.distinct_met[P.ndistinct++] = (uint8_t)m;
        P.mets[m].id64 = t->cols[q.acc_cols[m]].width == 8;
        ids64 = ids64 || P.mets[m].id64;
      }
    }

    for (uint32_t k = 0; k < P.nkeys; ++k) {
      const Slot &sl = P.slots[P.keys[k].slot];
      P.keys[k].col_off = sl.off; P.keys[k].vmask = sl.vmask; P.keys[k].signbit = sl.signbit;
      P.keys[k].width = sl.width; P.keys[k].row_off = sl.row_off;
    }
    for (uint32_t m = 0; m < P.nmetrics; ++m) {
      const Slot &sl = P.slots[P.mets[m].slot];
      P.mets[m].col_off = sl.off; P.mets[m].vmask = sl.vmask; P.mets[m].signbit = sl.signbit;
      P.mets[m].width = sl.width; P.mets[m].row_off = sl.row_off;
      P.mets[m].bitset = sl.bitset; P.mets[m].bitset_idx = sl.bitset_idx;
    }
    // No predicate column at all (every row passes): the L2 prefetch of the next chunk — which otherwise pulls in the
    // predicate columns — takes the key and metric columns instead; every byte of them is used, and the per-row gathers
    // then find them in L2 (C4: the gathers were 21 % of the stall samples, profiles/r2_final_scan_ncu_c4.txt).
    if (P.nfilter_slots == 0 && plan->nnodes == 0) {
      auto add_pf = [&](uint32_t slot) {
        const Slot &sl = P.slots[slot];
        if (sl.bitset) return;
        for (uint32_t f = 0; f < P.nfilter_slots; ++f)
          if (P.pf_off[f] == sl.off) return;
        if (P.nfilter_slots >= 16) return;
        P.pf_width[P.nfilter_slots] = (uint8_t)sl.width;
        P.pf_off[P.nfilter_slots] = sl.off;
        P.filter_slots[P.nfilter_slots++] = (uint8_t)slot;
      };
      for (uint32_t k = 0; k < P.nkeys; ++k) add_pf(P.keys[k].slot);
      for (uint32_t m = 0; m < P.nmetrics; ++m) add_pf(P.mets[m].slot);
    }
    P.small_plan = P.nkeys <= 4 && P.nmetrics <= 4;
    for (uint32_t k = 0; k < P.nkeys; ++k)
      if (P.slots[P.keys[k].slot].width > 4) P.small_plan = 0;

    // ---- row-major mirror or columns for the cells of passing rows? (per chunk, in the kernel) ----
    // Bytes of DRAM atoms (64 B) each way for a 512-row chunk with n passing rows: the mirror costs the
    // atoms one row's cells span; a column costs every atom that holds at least one passing row. Columns
    // the predicate has just streamed are in L2 either way.
    P.row_stride = t->row_stride;
    P.row_thresh = 0;
    if (t->row_stride && P.small_plan && !(ctx->tune & 128u) && (plan->nnodes > 0 || (ctx->tune & 256u))) {
      bool all_mirrored = true;
      for (uint32_t s : q.active) all_mirrored = all_mirrored && t->segs[s].rows != nullptr;
      uint32_t lo_off = ~0u, hi_off = 0;
      std::vector<uint32_t> widths;
      auto payload = [&](uint32_t slot) {
        const Slot &sl = P.slots[slot];
        lo_off = std::min(lo_off, sl.row_off);
        hi_off = std::max(hi_off, sl.row_off + sl.width);
        bool streamed = false;
        for (uint32_t f = 0; f < P.nfilter_slots; ++f) streamed = streamed || P.filter_slots[f] == slot;
        if (!streamed) widths.push_back(sl.width);
      };
      for (uint32_t k = 0; k < P.nkeys; ++k) payload(P.keys[k].slot);
      for (uint32_t m = 0; m < P.nmetrics; ++m) payload(P.mets[m].slot);
      if (all_mirrored && !widths.empty()) {
        const double span = (double)(hi_off - lo_off);
        const double row_cost = 64.0 * (1.0 + (span - 1.0) / 64.0);
        for (uint32_t n = 1; n <= (uint32_t)kChunkRows; ++n) {
          double col_cost = 0;
          for (uint32_t w : widths)
            col_cost += 8.0 * w * 64.0 * (1.0 - std::pow(1.0 - (double)n / kChunkRows, 64.0 / w));
          if (n * row_cost < col_cost) P.row_thresh = n; else break;
        }
      }
      if (all_mirrored && (ctx->tune & 256u)) P.row_thresh = kChunkRows;  // tests: every batch from the mirror
    }

    // ---- dense or hash ----
    unsigned __int128 cells128 = 1;
    bool fits64 = true;
    for (auto &r : q.ranges) {
      cells128 *= range128(r);
      if (cells128 > ((unsigned __int128)1 << 64) - 2) { fits64 = false; break; }
    }
    const bool wide = !fits64;  // key tuple wider than 64 bits: hash on the full tuple
    if (wide && G > 1 && P.ndistinct)
      fail(VGPU_ERR_UNSUPPORTED, "multi-GPU count-distinct over key tuples wider than 64 bits is not implemented");
    if (wide && (plan->flags & VGPU_PLAN_FORCE_DENSE)) fail(VGPU_ERR_UNSUPPORTED, "key domain too large for a dense group table");
    const uint64_t cells = wide ? ~0ull : (uint64_t)cells128;
    uint64_t dense_limit = std::min<uint64_t>(std::max<uint64_t>(4 * global_active_rows, 1ull << 22), 1ull << 28);
    if (P.ndistinct) dense_limit = std::min<uint64_t>(dense_limit, 0xffffffffull);
    bool dense = !wide && cells <= dense_limit;
    if (plan->flags & VGPU_PLAN_FORCE_HASH) dense = false;
    if (plan->flags & VGPU_PLAN_FORCE_DENSE) {
      if (cells > (1ull << 30)) fail(VGPU_ERR_UNSUPPORTED, "key domain too large for a dense group table");
      dense = true;
    }
    q.hash_mode = !dense;
    q.wide = wide;
    if (wide) P.row_thresh = 0;  // the wide-tuple path gathers from the columns
    if (ctx->trace) {
      fprintf(stderr, "[vgpu r%d] cells=%llu dense_limit=%llu dense=%d active_rows=%llu global=%llu tdict=%u\n", ctx->rank, (unsigned long long)cells, (unsigned long long)dense_limit, (int)dense, (unsigned long long)q.active_rows, (unsigned long long)global_active_rows, P.tdict.npieces);
      for (uint32_t k = 0; k < plan->nkeys; ++k) fprintf(stderr, "[vgpu r%d]   key %u lo=%llu range=%llu kmin=%llu kmax=%llu\n", ctx->rank, k, (unsigned long long)q.ranges[k].lo, (unsigned long long)q.ranges[k].range, (unsigned long long)kmin[k], (unsigned long long)kmax[k]);
    }
    {
      uint64_t mul = 1;
      for (uint32_t k = 0; k < plan->nkeys; ++k) {
        P.keys[k].lo = q.ranges[k].lo;
        P.keys[k].mul = mul;
        mul *= q.ranges[k].range;  // the last multiplication may wrap only if it is never used
      }
    }

    // ---- post-aggregation on the device (vgpu_plan: HAVING, top-N) ----
    Planner hpl(t, plan, true);
    bool post_having = false, post_topn = false;
    uint32_t sort_src = 0xffffffffu, sort_kind = 0;
    if ((plan->flags & VGPU_PLAN_POST) && !(ctx->tune & (1u << 23)) && !q.active.empty()) {
      if (plan->nhnodes > 0) {
        hpl.build_predicate();   // throws for columns that are not selected, like the reference (query.cc:127-133)
        post_having = hpl.P.nprog <= (uint32_t)kMaxPostProg;
      }
      if (plan->sort_col != VGPU_NO_COLUMN && plan->top_k > 0) {
        const ColInfo &ci = t->cols[plan->sort_col];
        // the column's sort order must follow from its raw value: decimal integers under SmallerInt (string.h:28-49).
        // Strings, times (sorted as strings), booleans ("true" / "false"), floats and AVG (a quotient) stay the host's.
        const bool is_dim = plan->sort_col < t->ndims;
        const bool ok = !type_float(ci.type) && (is_dim ? ci.kind == VGPU_DIM_NUMERIC : (ci.bitset || (ci.agg != VGPU_AGG_AVG && ci.kind != VGPU_METRIC_HIDDEN_COUNT)));
        if (ok) {
          sort_src = hpl.source_of(plan->sort_col);
          sort_kind = type_signed(ci.type) && !ci.bitset ? 1u : 0u;
          post_topn = true;
        }
      }
    }
    const bool post = post_having || post_topn;

    // ---- work list ----
    uint64_t max_rows = 0;
    for (uint32_t s : q.active) max_rows = std::max(max_rows, t->segs[s].nrows);
    P.tiles_per_seg = (uint32_t)std::max<uint64_t>(1, (max_rows + kChunkRows - 1) / kChunkRows);
    P.nactive = (uint32_t)q.active.size();
    P.total_tiles = (uint64_t)P.nactive * P.tiles_per_seg;
    // work units: about 8 per resident warp so that dynamic scheduling evens out the tail, at most 64 chunks
    P.unit_chunks = ctx->unit_chunks;
    if (P.unit_chunks == 0) {
      const uint64_t warps = (uint64_t)ctx->sm_count * ctx->ctas_per_sm * kScanWarps;
      P.unit_chunks = (uint32_t)std::min<uint64_t>(64, std::max<uint64_t>(4, P.total_tiles / (8 * warps)));
    }
    P.units_per_seg = (P.tiles_per_seg + P.unit_chunks - 1) / P.unit_chunks;
    if ((uint64_t)P.nactive * P.units_per_seg > 0x7fffffffull) fail(VGPU_ERR_UNSUPPORTED, "too many work units");
    P.segs = t->d_segs;
    P.tune = ctx->tune;

    std::unique_ptr<vgpu_result> res(new vgpu_result());
    vgpu_result_view &view = res->view;
    view.nkeys = plan->nkeys;
    view.nmetrics = plan->nmetrics;

    CUDA_CK(cudaEventRecord(sc->ev_begin, stream));
    uint32_t launches = 0, paths = 0, paths_post = 0;
    float scan_ms_total = 0;
    bool force_general = false;   // device post-aggregation: the fast dedupe overflowed, scan again and dedupe the general way

    uint64_t hash_cap = 0;
    if (q.hash_mode) {
      uint64_t est = std::min<uint64_t>(cells, std::max<uint64_t>(global_active_rows, 1));
      uint64_t want = pow2_ceil(std::max<uint64_t>(2 * est, 1024));
      hash_cap = std::min<uint64_t>(want, 1ull << 24);
      hash_cap = std::max(hash_cap, std::min(hint_load(t->hash_cap_hint), want));
      if (ctx->test_hash_cap) hash_cap = pow2_ceil(std::max<uint64_t>(ctx->test_hash_cap, 16));
      if (hash_cap > 0xfffffffeull && P.ndistinct) fail(VGPU_ERR_UNSUPPORTED, "count-distinct over more than 2^32 groups");
    }
    // the scan grid (also the number of count-distinct pair regions per owner)
    const int scan_grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((P.total_tiles + kScanWarps - 1) / kScanWarps,
                                                                       (uint64_t)ctx->sm_count * ctx->ctas_per_sm));
    // With several ranks every buffer that takes part in an exchange must have the same size everywhere: the grid is
    // the full one (ranks with less work leave regions empty) and capacities come from agreed numbers only.
    const int region_grid = G > 1 ? ctx->sm_count * ctx->ctas_per_sm : scan_grid;
    // count-distinct: every CTA appends its (cell,id) pairs to private regions. Capacities follow the high-water
    // marks of earlier queries on this table; overflow => grow and scan again
    P.dpair_nsub = P.ndistinct && G > 1 ? (uint32_t)G : 1u;
    P.dpair_key = P.ndistinct && G > 1 && q.hash_mode;
    P.dpair_wide = P.ndistinct && (ids64 || P.dpair_key);
    const uint32_t pair_elem = P.dpair_wide ? 16u : 8u;
    uint64_t region_cap64 = 0;
    if (P.ndistinct) {
      const uint64_t nregions = (uint64_t)region_grid * P.dpair_nsub;
      const uint64_t fill_hint = hint_load(t->pairs_region_hint);
      if (fill_hint) region_cap64 = fill_hint + fill_hint / 16 + std::min<uint64_t>(fill_hint / 4, 1024) + 256;   // work units are handed out dynamically: a CTA's share varies from run to run (small tables most)
      else region_cap64 = std::max<uint64_t>(1ull << 16, max_active_rows / 16) / nregions * 5 / 4 + 1024;
      if (ctx->test_pairs_cap) region_cap64 = std::max<uint64_t>(ctx->test_pairs_cap / nregions, 4);
    }

    for (int attempt = 0;; ++attempt) {
      if (attempt > 12) fail(VGPU_ERR_NOMEM, "group table keeps overflowing");
      Scratch scratch(stream, sc->s1);
      q.ncells = q.hash_mode ? hash_cap : cells;
      const uint64_t acc_cells = q.hash_mode ? hash_cap + 1 : cells;  // + the all-ones-key cell
      P.hash_mode = q.wide ? 2u : (q.hash_mode ? 1u : 0u);
      P.max_probe = 512;
      // Group table layout. Single GPU: the fields of a cell (key, accumulators, presence flag) are
      // INTERLEAVED, so that one passing row touches one line of the table instead of one line per
      // metric array (tables beyond a few MB are DRAM-resident under the column stream: measured
      // +3..10 B/row of traffic with separate arrays). Several GPUs: one array per field, because the
      // NCCL merge reduces each array with its own type and operator.
      const bool interleave = G == 1;
      uint64_t block_bytes = 0;
      auto carve = [&](uint64_t bytes) { uint64_t o = block_bytes; block_bytes += round_up(bytes, 256); return o; };
      const uint64_t o_wstate = q.wide ? carve(hash_cap * 4) : 0;
      const uint64_t o_wkeys = q.wide ? carve(hash_cap * 8 * std::max<uint32_t>(plan->nkeys, 1)) : 0;
      std::vector<void *> acc_ptrs(q.accs.size());
      std::vector<uint32_t> acc_stride(q.accs.size());
      uint8_t *block = nullptr;
      const bool key_in_cell = q.hash_mode && !q.wide;
      if (interleave) {
        // cell = [key u64]? [8-byte accumulators] [4-byte accumulators] [presence u32]?
        uint32_t off = 0;
        uint32_t o_key = 0, o_pres = 0;
        std::vector<uint32_t> o_f(q.accs.size());
        if (key_in_cell) { o_key = off; off += 8; }
        for (size_t m = 0; m < q.accs.size(); ++m) if (q.accs[m].acc_width == 8) { o_f[m] = off; off += 8; }
        for (size_t m = 0; m < q.accs.size(); ++m) if (q.accs[m].acc_width == 4) { o_f[m] = off; off += 4; }
        const bool need_present = !q.hash_mode;
        if (need_present) { o_pres = off; off += 4; }
        uint32_t stride = off <= 4 ? 4 : off <= 8 ? 8 : off <= 16 ? 16 : off <= 32 ? 32 : (uint32_t)round_up(off, 8);
        if (stride / 4 > 48) fail(VGPU_ERR_UNSUPPORTED, "group cell too wide");
        const uint64_t o_cells = carve(acc_cells * (uint64_t)stride);
        const uint64_t o_flag = carve(16);
        block = scratch.alloc<uint8_t>(block_bytes);
        CellPattern C{};
        C.words = stride / 4;
        if (key_in_cell) { C.w[o_key / 4] = 0xffffffffu; C.w[o_key / 4 + 1] = 0xffffffffu; }
        for (size_t m = 0; m < q.accs.size(); ++m) {
          C.w[o_f[m] / 4] = (uint32_t)q.accs[m].init;
          if (q.accs[m].acc_width == 8) C.w[o_f[m] / 4 + 1] = (uint32_t)(q.accs[m].init >> 32);
          acc_ptrs[m] = block + o_cells + o_f[m];
          acc_stride[m] = stride;
        }
        const uint64_t total_words = acc_cells * (uint64_t)(stride / 4);
        fill_cells_kernel<<<grid_for(total_words, 256, ctx->sm_count), 256, 0, stream>>>(
            reinterpret_cast<uint32_t *>(block + o_cells), total_words, C);
        CUDA_CK(cudaGetLastError());
        ++launches;
        P.hkeys = key_in_cell ? reinterpret_cast<uint64_t *>(block + o_cells + o_key) : nullptr;
        P.hkey_stride = stride;
        P.hmask = q.hash_mode ? hash_cap - 1 : 0;
        if (need_present) {
          P.present = block + o_cells + o_pres;
          P.present_stride = stride;
        } else {
          P.present = block + o_flag;  // present[0] flags the all-ones key
          P.present_stride = 1;
          CUDA_CK(cudaMemsetAsync(P.present, 0, 16, stream));
        }
      } else {
        const uint64_t o_hkeys = key_in_cell ? carve(hash_cap * 8) : 0;
        const uint64_t o_present = carve(q.hash_mode ? 16 : acc_cells);
        std::vector<uint64_t> o_acc(q.accs.size());
        for (size_t m = 0; m < q.accs.size(); ++m) o_acc[m] = carve(acc_cells * q.accs[m].acc_width);
        block = scratch.alloc<uint8_t>(block_bytes);
        P.hkeys = key_in_cell ? reinterpret_cast<uint64_t *>(block + o_hkeys) : nullptr;
        P.hkey_stride = 8;
        P.hmask = q.hash_mode ? hash_cap - 1 : 0;
        if (key_in_cell) fill64(stream, ctx->sm_count, P.hkeys, hash_cap, kEmptyKey);
        P.present = block + o_present;
        P.present_stride = 1;
        CUDA_CK(cudaMemsetAsync(P.present, 0, q.hash_mode ? 16 : acc_cells, stream));
        for (size_t m = 0; m < q.accs.size(); ++m) {
          const AccInfo &a = q.accs[m];
          acc_ptrs[m] = block + o_acc[m];
          acc_stride[m] = a.acc_width;
          if (a.acc_width == 4) launches += fill32(stream, ctx->sm_count, acc_ptrs[m], acc_cells, (uint32_t)a.init);
          else launches += fill64(stream, ctx->sm_count, acc_ptrs[m], acc_cells, a.init);
        }
      }
      if (q.wide) {
        P.wstate = reinterpret_cast<uint32_t *>(block + o_wstate);
        P.wkeys = reinterpret_cast<uint64_t *>(block + o_wkeys);
        CUDA_CK(cudaMemsetAsync(P.wstate, 0, hash_cap * 4, stream));
      }
      for (size_t m = 0; m < q.accs.size(); ++m) {
        P.mets[m].acc = acc_ptrs[m];
        P.mets[m].stride = acc_stride[m];
        P.mets[m].acc_width = q.accs[m].acc_width;
      }
      // CTA-private shared-memory copy of a small dense table (see ScanParams::smem_cells)
      P.smem_cells = 0;
      uint32_t scan_dyn_smem = 0;
      if (!q.hash_mode && !(ctx->tune & 262144u)) {
        uint32_t off = 0;
        for (size_t m = 0; m < q.accs.size(); ++m) if (q.accs[m].acc_width == 8 && q.accs[m].op != A_DISTINCT) { P.mets[m].soff = off; off += 8; }
        for (size_t m = 0; m < q.accs.size(); ++m) if (q.accs[m].acc_width == 4 && q.accs[m].op != A_DISTINCT) { P.mets[m].soff = off; off += 4; }
        const uint32_t pres = off;
        off += 4;
        const uint32_t sstride = (uint32_t)round_up(off, 8);
        if (sstride <= 64 && cells * sstride <= kSmemTableBytes) {
          P.smem_cells = (uint32_t)cells;
          P.smem_stride = sstride;
          P.smem_present_off = pres;
          for (uint32_t w = 0; w < 16; ++w) P.smem_init[w] = 0;
          for (size_t m = 0; m < q.accs.size(); ++m) {
            if (q.accs[m].op == A_DISTINCT) continue;
            P.smem_init[P.mets[m].soff / 4] = (uint32_t)q.accs[m].init;
            if (q.accs[m].acc_width == 8) P.smem_init[P.mets[m].soff / 4 + 1] = (uint32_t)(q.accs[m].init >> 32);
          }
          scan_dyn_smem = (uint32_t)cells * sstride;
        }
      }
      // Presence from a COUNT accumulator: every count cell of the active segments is >= 1 (statistics reduced at put
      // time) and this rank's sums cannot wrap, so "count != 0" is exactly "some row reached the cell".
      P.skip_present = 0;
      int present_acc = -1;
      if (!q.hash_mode && !P.smem_cells && !(ctx->tune & (1u << 22))) {
        for (size_t m = 0; m < q.accs.size() && present_acc < 0; ++m) {
          const ColInfo &ci = t->cols[q.acc_cols[m]];
          if (ci.agg != VGPU_AGG_COUNT || (q.accs[m].op != A_ADD32 && q.accs[m].op != A_ADD64)) continue;
          uint64_t cmin = ~0ull, cmax = 0;
          for (uint32_t s : q.active) {
            const SegmentData &sd = t->segs[s];
            if (sd.nrows == 0) continue;
            cmin = std::min(cmin, sd.omin[q.acc_cols[m]]);
            cmax = std::max(cmax, sd.omax[q.acc_cols[m]]);
          }
          if (cmin < 1 || cmin > cmax) continue;
          const unsigned __int128 bound = (unsigned __int128)cmax * std::max<uint64_t>(q.active_rows, 1);
          if (bound >= ((unsigned __int128)1 << (8 * q.accs[m].acc_width))) continue;
          present_acc = (int)m;
        }
        P.skip_present = present_acc >= 0;
      }
      // count-distinct pair regions
      if (P.ndistinct && region_cap64 > 0xffffffffull) fail(VGPU_ERR_NOMEM, "count-distinct pair regions too large");
      P.dpair_cap = (uint32_t)region_cap64;
      const uint32_t nregions = (uint32_t)region_grid * P.dpair_nsub;
      for (uint32_t d = 0; d < P.ndistinct; ++d) {
        P.dpairs[d] = reinterpret_cast<uint64_t *>(scratch.alloc<uint8_t>(region_cap64 * nregions * pair_elem));
        P.dpair_count[d] = scratch.alloc<uint32_t>(nregions);
        CUDA_CK(cudaMemsetAsync(P.dpair_count[d], 0, nregions * 4, stream));
      }
      // counter block: zero, except the host-provided statistics that are summed across ranks with it
      {
        unsigned long long *init = sc->h_counters + 16;
        for (int i = 0; i < 16; ++i) init[i] = 0;
        init[kCScannedRecs] = q.scanned_recs;
        init[kCScannedSegs] = q.active.size();
        CUDA_CK(cudaMemcpyAsync(sc->d_counters, init, 16 * sizeof(unsigned long long), cudaMemcpyHostToDevice, stream));
      }
      P.counters = sc->d_counters;
      uint32_t *d_active = scratch.alloc<uint32_t>(q.active.size());
      if (!q.active.empty())
        CUDA_CK(cudaMemcpyAsync(d_active, q.active.data(), q.active.size() * 4, cudaMemcpyHostToDevice, stream));
      P.active = d_active;

      // ---- the fused scan ----
      CUDA_CK(cudaEventRecord(sc->ev_scan0, stream));
      if (P.total_tiles > 0) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(G > 1 && P.ndistinct ? region_grid : scan_grid);
        cfg.blockDim = dim3(kScanThreads);
        cfg.stream = stream;
        cfg.dynamicSmemBytes = scan_dyn_smem;
        cudaLaunchAttribute attr[1];
        cfg.attrs = attr;
        cfg.numAttrs = 0;
        if ((ctx->tune & 1u) && ctx->l2_persist_bytes > 0 && block_bytes > 0) {
          // pin as much of the group table as the persisting carve-out holds
          const uint64_t win = std::min<uint64_t>(block_bytes, ctx->l2_window_max);
          attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
          attr[0].val.accessPolicyWindow.base_ptr = block;
          attr[0].val.accessPolicyWindow.num_bytes = win;
          attr[0].val.accessPolicyWindow.hitRatio =
              (float)std::min(1.0, (double)ctx->l2_persist_bytes / (double)win);
          attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
          attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
          cfg.numAttrs = 1;
        }
        bool plain_keys = P.tdict.npieces == 0 && !q.wide;
        for (uint32_t k = 0; k < P.nkeys; ++k) plain_keys = plain_keys && !P.keys[k].rollup && !P.keys[k].fzero;
        // the fast instantiation: register-staged keys and metrics, 8-byte count-distinct pairs (scan_kernel.cuh)
        const bool fast = P.small_plan && !q.wide && !P.dpair_wide;
        auto launch = [&](auto kernel) { CUDA_CK(cudaLaunchKernelEx(&cfg, kernel, P)); };
        const int variant = (fast && P.conj ? 8 : 0) | (P.smem_cells ? 4 : 0) | (plain_keys ? 2 : 0) | (fast ? 1 : 0);
        switch (variant) {
          case 9: launch(scan_filter_groupby_kernel<VGPU_MIN_CTAS, false, false, true, true>); break;
          case 11: launch(scan_filter_groupby_kernel<VGPU_MIN_CTAS, false, true, true, true>); break;
          case 13: launch(scan_filter_groupby_kernel<VGPU_MIN_CTAS, true, false, true, true>); break;
          case 15: launch(scan_filter_groupby_kernel<VGPU_MIN_CTAS, true, true, true, true>); break;
          case 0: launch(scan_filter_groupby_kernel<VGPU_MIN_CTAS, false, false, false>); break;
          case 1: launch(scan_filter_groupby_kernel<VGPU_MIN_CTAS, false, false, true>); break;
          case 2: launch(scan_filter_groupby_kernel<VGPU_MIN_CTAS, false, true, false>); break;
          case 3: launch(scan_filter_groupby_kernel<VGPU_MIN_CTAS, false, true, true>); break;
          case 4: launch(scan_filter_groupby_kernel<VGPU_MIN_CTAS, true, false, false>); break;
          case 5: launch(scan_filter_groupby_kernel<VGPU_MIN_CTAS, true, false, true>); break;
          case 6: launch(scan_filter_groupby_kernel<VGPU_MIN_CTAS, true, true, false>); break;
          default: launch(scan_filter_groupby_kernel<VGPU_MIN_CTAS, true, true, true>); break;
        }
        ++launches;
      }
      CUDA_CK(cudaEventRecord(sc->ev_scan1, stream));

      // ---- where the count-distinct pairs are, and how they will be deduplicated ----
      std::vector<PairInput> pin(P.ndistinct);
      std::vector<DedupeMode> dmode(P.ndistinct);
      std::vector<uint32_t> dnb(P.ndistinct, 0);
      for (uint32_t d = 0; d < P.ndistinct; ++d) {
        pin[d].pairs = P.dpairs[d];
        pin[d].counts = P.dpair_count[d];
        pin[d].nregions = nregions;        // several ranks: what this rank receives is the same shape
        pin[d].region_cap = P.dpair_cap;
        pin[d].wide = P.dpair_wide != 0;
        pin[d].expect = hint_load(t->pairs_total_hint);
        if (ctx->test_pairs_cap) pin[d].expect = 0;
        if (ctx->test_expect) pin[d].expect = ctx->test_expect;
        dmode[d] = choose_dedupe_mode(ctx, t, pin[d], dnb[d]);
        if (force_general && dmode[d] == kDedupeFast) dmode[d] = kDedupeGeneral;
      }

      // ---- several GPUs: one exchange ----
      uint64_t acc_cells_x = acc_cells;  // group table the extraction reads (replaced by the merged one)
      unsigned long long *hc = sc->h_counters;
      bool counters_read = false, abort_attempt = false;
      auto read_counters = [&](cudaStream_t s) {
        CUDA_CK(cudaMemcpyAsync(hc, sc->d_counters, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        CUDA_CK(cudaStreamSynchronize(s));
        counters_read = true;
      };
      if (G > 1) {
        // One launch per all-reduce. Measured on 8 B200s (NCCL 2.28.9, tools/gpu_n8_debug.sh): with the all-reduces of the
        // accumulator arrays inside ONE ncclGroup the NVLS algorithm lost min-updates of the first array in ~2 % of the
        // cells (a different set every run); ungrouped, NCCL_ALGO=Ring and NCCL_NVLS_ENABLE=0 are all exact. Only the
        // point-to-point exchange of the count-distinct pairs needs a group. VGPU_NCCL_GROUP=1 restores the grouped form.
        static const bool ungroup = getenv("VGPU_NCCL_GROUP") == nullptr;
        if (!ungroup) NCCL_CK(g_nccl.GroupStart());
        NCCL_CK(g_nccl.AllReduce(sc->d_counters + kCSumFirst, sc->d_counters + kCSumFirst, kCMaxFirst - kCSumFirst, ncclUint64,
                                 ncclSum, ctx->comm, stream));
        NCCL_CK(g_nccl.AllReduce(sc->d_counters + kCMaxFirst, sc->d_counters + kCMaxFirst, kCLocalFirst - kCMaxFirst, ncclUint64,
                                 ncclMax, ctx->comm, stream));
        if (!q.hash_mode) {
          if (P.skip_present) {   // the flags are what the ranks max-reduce: read them off the local counts first
            derive_present_kernel<<<grid_for(q.ncells, 256, ctx->sm_count), 256, 0, stream>>>(
                P.present, static_cast<const uint8_t *>(acc_ptrs[present_acc]), q.accs[present_acc].acc_width, q.ncells);
            CUDA_CK(cudaGetLastError());
            ++launches;
          }
          for (size_t m = 0; m < q.accs.size(); ++m) {
            if (q.accs[m].op == A_DISTINCT) continue;
            NCCL_CK(g_nccl.AllReduce(acc_ptrs[m], acc_ptrs[m], q.ncells, q.accs[m].nccl_type, q.accs[m].nccl_op, ctx->comm, stream));
          }
          NCCL_CK(g_nccl.AllReduce(P.present, P.present, q.ncells, ncclUint8, ncclMax, ctx->comm, stream));
        }
        // pairs to their owners: what a rank receives has the shape of what it sends
        if (ungroup && P.ndistinct) NCCL_CK(g_nccl.GroupStart());
        for (uint32_t d = 0; d < P.ndistinct; ++d) {
          void *recv = scratch.alloc<uint8_t>(region_cap64 * nregions * pair_elem);
          uint32_t *recv_counts = scratch.alloc<uint32_t>(nregions);
          exchange_pairs(ctx, stream, P.dpairs[d], P.dpair_count[d], (uint32_t)region_grid, P.dpair_cap, pair_elem, recv, recv_counts);
          pin[d].pairs = recv;
          pin[d].counts = recv_counts;
        }
        if (!ungroup || P.ndistinct) NCCL_CK(g_nccl.GroupEnd());
        if (q.hash_mode) {
          // every rank must take the same grow-and-retry decision before the records travel
          read_counters(stream);
          abort_attempt = hc[kCHashOver] != 0 || hc[kCRegionOver] != 0;
          if (!abort_attempt && q.wide) {
            nccl_merge_wide(ctx, sc, q, P, acc_ptrs, scratch, acc_cells_x, std::max<uint64_t>(q.active_rows, 1), launches);
          } else if (!abort_attempt) {
            nccl_merge_hash(ctx, sc, q, P, acc_ptrs, scratch, acc_cells_x, launches, [&] {
              // count-distinct: the owner's table holds every key it owns; dedupe the pairs of those keys against it
              for (uint32_t d = 0; d < P.ndistinct; ++d) {
                DistinctTarget tg{static_cast<uint8_t *>(acc_ptrs[P.distinct_met[d]]), 4, P.hkeys, P.hmask};
                dedupe_enqueue(ctx, sc, scratch, kDedupeWide, 0, pin[d], tg, launches, paths);
              }
            });
          }
        }
      }

      if (ctx->trace) CUDA_CK(cudaEventRecord(sc->ev_phase[0], stream));   // merge / exchange enqueued
      // ---- extraction plumbing ----
      // rows the output arrays can hold: an upper bound of the number of groups known before the scan has run
      const uint64_t rows_bound = G > 1 ? std::min<uint64_t>(acc_cells_x, std::max<uint64_t>(global_active_rows, 1))
                                        : std::min<uint64_t>(acc_cells_x, std::max<uint64_t>(q.active_rows, 1));
      const bool have_rows = q.active_rows > 0 || G > 1;
      std::vector<void *> d_keys(plan->nkeys), d_accs(q.accs.size());
      ExtractParams E{};
      E.ncells = acc_cells_x;
      E.hash_mode = q.wide ? 2u : (q.hash_mode ? 1u : 0u);
      E.nkeys = plan->nkeys;
      E.hkeys = P.hkeys;
      E.present = P.present;
      E.hkey_stride = P.hkey_stride;
      E.present_stride = P.present_stride;
      E.present_width = 1;
      if (P.skip_present && G == 1) {   // single GPU: the count accumulator is the presence word
        E.present = static_cast<const uint8_t *>(acc_ptrs[present_acc]);
        E.present_stride = P.mets[present_acc].stride;
        E.present_width = q.accs[present_acc].acc_width;
      }
      E.wstate = P.wstate;
      E.wkeys = P.wkeys;
      E.counter = sc->d_counters + kCGroups;
      E.cap = rows_bound;
      uint64_t *d_dict = nullptr;
      if (P.tdict.npieces) {
        d_dict = scratch.alloc<uint64_t>(tdict_values.size());
        CUDA_CK(cudaMemcpyAsync(d_dict, tdict_values.data(), tdict_values.size() * 8, cudaMemcpyHostToDevice, stream));
      }
      for (uint32_t k = 0; k < plan->nkeys; ++k) {
        const ColInfo &ci = t->cols[plan->keys[k].col];
        d_keys[k] = scratch.alloc<uint8_t>(rows_bound * ci.width);
        E.keys[k].lo = q.ranges[k].lo;
        E.keys[k].lut = P.keys[k].lut;
        E.keys[k].dict = (P.tdict.npieces && P.tdict.key == k) ? d_dict : nullptr;
        E.keys[k].div = P.keys[k].mul;
        E.keys[k].mod = (k + 1 < plan->nkeys) ? q.ranges[k].range : 0;
        E.keys[k].width = ci.width;
        E.keys[k].out = d_keys[k];
      }
      for (size_t m = 0; m < q.accs.size(); ++m) d_accs[m] = scratch.alloc<uint8_t>(rows_bound * q.accs[m].out_width);
      auto extract_met = [&](size_t m) {
        ExtractMet em{};
        em.acc = acc_ptrs[m];
        em.stride = P.mets[m].stride;
        em.acc_width = q.accs[m].acc_width;
        em.out_width = q.accs[m].out_width;
        em.sext = q.accs[m].op != A_DISTINCT && type_signed(t->cols[q.acc_cols[m]].type) ? 1u : 0u;
        em.out = d_accs[m];
        return em;
      };
      // Count-distinct queries extract twice: keys and the finished accumulators right after the scan, on the side
      // stream (their copy to the host overlaps the dedupe), the distinct counts after the dedupe through the position
      // map the first pass left. Hashed tables merged across ranks are extracted once, at the end.
      const bool dedupe_after_sync = [&] {
        for (uint32_t d = 0; d < P.ndistinct; ++d)
          if (dmode[d] == kDedupeGeneral) return true;
        return false;
      }();
      const bool distinct_done = G > 1 && q.hash_mode;  // deduplicated inside the hash merge
      // (device post-aggregation reads every accumulator of a group at once: one extraction, after the dedupe)
      const bool early = P.ndistinct > 0 && !distinct_done && !post && acc_cells_x <= (1ull << 24) && !(ctx->tune & (1u << 20));
      uint32_t *d_pos = nullptr;
      auto run_extract = [&](cudaStream_t s, bool with_late) {
        E.nmets = 0;
        for (size_t m = 0; m < q.accs.size(); ++m)
          if (with_late || q.accs[m].op != A_DISTINCT) E.mets[E.nmets++] = extract_met(m);
        E.pos_out = with_late ? nullptr : d_pos;
        E.count_only = 0;
        if (!post) {
          extract_groups_kernel<<<grid_for(acc_cells_x, 256, ctx->sm_count), 256, 0, s>>>(E);
          CUDA_CK(cudaGetLastError());
          ++launches;
          return;
        }
        // HAVING on the raw accumulators and / or top-N on the first sort key: candidates, radix select, extraction
        E.nhprog = post_having ? hpl.P.nprog : 0;
        for (uint32_t i = 0; i < E.nhprog; ++i) E.hprog[i] = hpl.P.prog[i];
        E.sort_src = post_topn ? sort_src : 0xffffffffu;
        E.sort_kind = sort_kind;
        E.sort_desc = plan->sort_descending ? 1u : 0u;
        E.cand_cell = scratch.alloc<uint64_t>(rows_bound);
        E.cand_ord = scratch.alloc<uint64_t>(rows_bound);
        E.cand_n = sc->d_counters + kCCand;
        E.out_n = sc->d_counters + kCOut;
        unsigned long long *state = scratch.alloc<unsigned long long>(4);   // [0] threshold
        unsigned int *hist = scratch.alloc<unsigned int>(256);
        E.threshold = state;
        post_select_kernel<<<grid_for(acc_cells_x, 256, ctx->sm_count), 256, 0, s>>>(E);
        CUDA_CK(cudaGetLastError());
        ++launches;
        if (post_topn) {
          CUDA_CK(cudaMemsetAsync(hist, 0, 256 * sizeof(unsigned int), s));
          radix_pick_kernel<<<1, 32, 0, s>>>(state, -1, hist, E.cand_n, plan->top_k);   // set-up
          CUDA_CK(cudaGetLastError());
          for (int shift = 56; shift >= 0; shift -= 8) {
            radix_hist_kernel<<<grid_for(rows_bound, 256, ctx->sm_count, 4), 256, 0, s>>>(E.cand_ord, E.cand_n, state, (uint32_t)shift, hist);
            CUDA_CK(cudaGetLastError());
            radix_pick_kernel<<<1, 32, 0, s>>>(state, shift, hist, E.cand_n, plan->top_k);
            CUDA_CK(cudaGetLastError());
            launches += 2;
          }
        } else {
          CUDA_CK(cudaMemsetAsync(state, 0xff, 8, s));
        }
        post_extract_kernel<<<grid_for(rows_bound, 256, ctx->sm_count, 4), 256, 0, s>>>(E);
        CUDA_CK(cudaGetLastError());
        ++launches;
        paths_post = (post_having ? 1u : 0u) | (post_topn ? 2u : 0u);
      };
      auto run_extract_late = [&](cudaStream_t s) {
        ExtractLateParams L{};
        L.ncells = acc_cells_x;
        L.pos = d_pos;
        for (uint32_t d = 0; d < P.ndistinct; ++d) L.mets[L.nmets++] = extract_met(P.distinct_met[d]);
        extract_late_kernel<<<grid_for(acc_cells_x, 256, ctx->sm_count), 256, 0, s>>>(L);
        CUDA_CK(cudaGetLastError());
        ++launches;
      };
      auto run_dedupe = [&] {
        for (uint32_t d = 0; d < P.ndistinct; ++d) {
          DistinctTarget tg{static_cast<uint8_t *>(acc_ptrs[P.distinct_met[d]]), P.mets[P.distinct_met[d]].stride};
          dedupe_enqueue(ctx, sc, scratch, dmode[d], dnb[d], pin[d], tg, launches, paths);
        }
        if (G > 1) {  // every pair was counted on exactly one rank; a rank whose fast path overflowed tells everybody
          for (uint32_t d = 0; d < P.ndistinct; ++d)
            NCCL_CK(g_nccl.AllReduce(acc_ptrs[P.distinct_met[d]], acc_ptrs[P.distinct_met[d]], q.ncells, ncclUint32, ncclSum, ctx->comm, stream));
          NCCL_CK(g_nccl.AllReduce(sc->d_counters + kCBucketOver, sc->d_counters + kCBucketOver, 2, ncclUint64, ncclMax, ctx->comm, stream));
        }
      };

      // ---- enqueue: dedupe and extraction, then the counter block ----
      cudaStream_t cstream = stream;  // where the counter block is read
      if (!abort_attempt) {
        if (early && have_rows) {
          d_pos = scratch.alloc<uint32_t>(acc_cells_x);
          CUDA_CK(cudaEventRecord(sc->ev_a, stream));            // scan (and merge) done
          CUDA_CK(cudaStreamWaitEvent(sc->s1, sc->ev_a, 0));
          CUDA_CK(cudaMemsetAsync(d_pos, 0xff, acc_cells_x * 4, sc->s1));
          run_extract(sc->s1, false);
          CUDA_CK(cudaEventRecord(sc->ev_b, sc->s1));
          cstream = sc->s1;
          if (!dedupe_after_sync) {
            run_dedupe();
            if (ctx->trace) CUDA_CK(cudaEventRecord(sc->ev_phase[1], stream));
            CUDA_CK(cudaStreamWaitEvent(stream, sc->ev_b, 0));   // the position map
            run_extract_late(stream);
          }
        } else if (have_rows) {
          if (P.ndistinct && !distinct_done && !dedupe_after_sync) run_dedupe();
          if (!(P.ndistinct && !distinct_done && dedupe_after_sync)) run_extract(stream, true);
        }
      }
      CUDA_CK(cudaMemcpyAsync(hc, sc->d_counters, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, cstream));
      CUDA_CK(cudaStreamSynchronize(cstream));   // host sync #1
      {
        float ms = 0;
        CUDA_CK(cudaEventElapsedTime(&ms, sc->ev_scan0, sc->ev_scan1));
        scan_ms_total += ms;
      }
      const uint64_t passed = hc[kCPassed];
      if (hc[kCHashOver] != 0 || hc[kCRegionOver] != 0) {  // grow, scan again (same decision on every rank)
        CUDA_CK(cudaStreamSynchronize(stream));
        CUDA_CK(cudaStreamSynchronize(sc->s1));
        if (hc[kCHashOver] != 0) {
          if (!q.hash_mode) fail(VGPU_ERR_CUDA, "unexpected overflow flag in dense mode");
          hash_cap *= 4;
        }
        if (hc[kCRegionOver] != 0) {  // a pair region overflowed: size for the fullest one
          uint64_t most = 0;
          for (uint32_t d = 0; d < P.ndistinct; ++d) most = std::max<uint64_t>(most, hc[kCMaxFill + d]);
          region_cap64 = std::max<uint64_t>(2 * region_cap64, most + most / 8 + 256);
        }
        continue;
      }
      if (q.hash_mode) hint_raise(t->hash_cap_hint, hash_cap);
      uint64_t pairs_total[kMaxDistinct] = {0, 0};
      if (P.ndistinct) {
        uint64_t most = 0, fill = 0;
        for (uint32_t d = 0; d < P.ndistinct; ++d) {
          // several ranks: the sum over ranks / G approximates what one owner receives
          pairs_total[d] = hc[kCPairs + d];
          most = std::max<uint64_t>(most, G > 1 ? (hc[kCPairs + d] + G - 1) / G : hc[kCPairs + d]);
          fill = std::max<uint64_t>(fill, hc[kCMaxFill + d]);
        }
        hint_raise(t->pairs_total_hint, most);
        hint_raise(t->pairs_region_hint, fill);
      }
      view.passed_rows = passed;
      view.scanned_recs = hc[kCScannedRecs];
      view.scanned_segments = hc[kCScannedSegs];

      // ---- the general dedupe path needs the totals the host has just read (and synchronises itself) ----
      bool late_pending = false;  // distinct counts still to be extracted
      if (P.ndistinct && !distinct_done && have_rows && dedupe_after_sync) {
        for (uint32_t d = 0; d < P.ndistinct; ++d) {
          DistinctTarget tg{static_cast<uint8_t *>(acc_ptrs[P.distinct_met[d]]), P.mets[P.distinct_met[d]].stride};
          if (dmode[d] == kDedupeGeneral) {
            // several ranks: pairs_total is the sum over all ranks, an upper bound of what this owner received
            dedupe_general(ctx, sc, scratch, pin[d], pairs_total[d], tg, launches, paths);
          } else {
            dedupe_enqueue(ctx, sc, scratch, dmode[d], dnb[d], pin[d], tg, launches, paths);
          }
        }
        if (G > 1)
          for (uint32_t d = 0; d < P.ndistinct; ++d)
            NCCL_CK(g_nccl.AllReduce(acc_ptrs[P.distinct_met[d]], acc_ptrs[P.distinct_met[d]], q.ncells, ncclUint32, ncclSum, ctx->comm, stream));
        late_pending = true;
      }
      // ---- the fast path reports bucket / set overflow through the counter block: run the general path instead ----
      auto redo_distinct_general = [&] {
        paths |= VGPU_DEDUPE_REDONE;
        for (uint32_t d = 0; d < P.ndistinct; ++d) {
          if (dmode[d] != kDedupeFast) continue;
          uint8_t *acc = static_cast<uint8_t *>(acc_ptrs[P.distinct_met[d]]);
          const uint32_t stride = P.mets[P.distinct_met[d]].stride;
          // buckets sized from a good estimate and still too small: the data is too uneven for this table, stay general
          if (pin[d].expect && pin[d].expect >= (G > 1 ? pairs_total[d] / G : pairs_total[d]) * 4 / 5 && !ctx->test_expect)
            t->distinct_general.store(1, std::memory_order_relaxed);
          CUDA_CK(cudaMemset2DAsync(acc, stride, 0, 4, acc_cells_x, stream));  // forget the partial (and summed) counts
          DistinctTarget tg{acc, stride};
          dedupe_general(ctx, sc, scratch, pin[d], pairs_total[d], tg, launches, paths);
          if (G > 1) NCCL_CK(g_nccl.AllReduce(acc, acc, q.ncells, ncclUint32, ncclSum, ctx->comm, stream));
        }
      };

      // ---- results to the host: one pinned block, arrays 64-byte aligned ----
      uint64_t ngroups = 0;
      auto host_copy = [&](uint64_t n) {
        uint64_t bytes = 0;
        std::vector<uint64_t> off_k(plan->nkeys), off_m(plan->nmetrics);
        uint64_t off_h = 0;
        auto place = [&](uint64_t nb) { uint64_t o = bytes; bytes += round_up(std::max<uint64_t>(nb, 1), 64); return o; };
        for (uint32_t k = 0; k < plan->nkeys; ++k) off_k[k] = place(n * t->cols[plan->keys[k].col].width);
        for (uint32_t m = 0; m < plan->nmetrics; ++m) off_m[m] = place(n * q.accs[m].out_width);
        if (plan->need_hidden_count) off_h = place(n * 8);
        res->pool = ctx->pool;
        res->block = ctx->pool->acquire(bytes);
        uint8_t *hb = static_cast<uint8_t *>(res->block.first);
        // keys and finished accumulators may leave on the side stream while the dedupe still runs
        cudaStream_t es = early ? sc->s1 : stream;
        for (uint32_t k = 0; k < plan->nkeys; ++k) {
          const uint64_t nb = n * t->cols[plan->keys[k].col].width;
          if (nb) CUDA_CK(cudaMemcpyAsync(hb + off_k[k], d_keys[k], nb, cudaMemcpyDeviceToHost, es));
          res->key_ptrs.push_back(hb + off_k[k]);
        }
        for (uint32_t m = 0; m < plan->nmetrics; ++m) {
          const uint64_t nb = n * q.accs[m].out_width;
          const bool is_late = early && q.accs[m].op == A_DISTINCT;
          if (nb) CUDA_CK(cudaMemcpyAsync(hb + off_m[m], d_accs[m], nb, cudaMemcpyDeviceToHost, is_late ? stream : es));
          res->acc_ptrs.push_back(hb + off_m[m]);
        }
        if (plan->need_hidden_count) {
          if (n) CUDA_CK(cudaMemcpyAsync(hb + off_h, d_accs[plan->nmetrics], n * 8, cudaMemcpyDeviceToHost, es));
          view.hidden_count = reinterpret_cast<const uint64_t *>(hb + off_h);
        }
      };
      if (have_rows) {
        if (early) {
          ngroups = hc[kCGroups];  // counted by the early pass
          if (late_pending) {
            CUDA_CK(cudaStreamWaitEvent(stream, sc->ev_b, 0));
            run_extract_late(stream);
          }
        } else if (late_pending) {   // nothing extracted yet
          run_extract(stream, true);
          read_counters(stream);
          ngroups = hc[kCGroups];
        } else {
          ngroups = hc[kCGroups];
        }
      }
      if (ngroups > rows_bound) fail(VGPU_ERR_CUDA, "group extraction overflow");
      // device post-aggregation: the result holds the groups that survived HAVING / the top-N cut, not all of them
      const uint64_t nout = (post && have_rows) ? (uint64_t)hc[kCOut] : ngroups;
      if (nout > rows_bound) fail(VGPU_ERR_CUDA, "group extraction overflow");
      // several GPUs, VGPU_PLAN_RESULT_ON_ROOT: the merged groups go to rank 0's host only
      const uint64_t out_rows = (G > 1 && (plan->flags & VGPU_PLAN_RESULT_ON_ROOT) && ctx->rank != 0) ? 0 : nout;
      host_copy(out_rows);
      // the flags of the fast dedupe path travel behind everything on s0
      unsigned long long *hc2 = sc->h_counters + 32;
      CUDA_CK(cudaMemcpyAsync(hc2, sc->d_counters, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
      CUDA_CK(cudaEventRecord(sc->ev_end, stream));
      CUDA_CK(cudaStreamSynchronize(sc->s1));
      CUDA_CK(cudaStreamSynchronize(stream));   // host sync #2
      if (post && P.ndistinct && !distinct_done && (hc2[kCBucketOver] != 0 || hc2[kCSetOver] != 0)) {
        // HAVING / top-N already looked at incomplete distinct counts: scan again, dedupe the general way (rare)
        force_general = true;
        if (res->block.first) { res->pool->release(res->block); res->block = {nullptr, 0}; }
        res->key_ptrs.clear();
        res->acc_ptrs.clear();
        view.hidden_count = nullptr;
        continue;
      }
      if (P.ndistinct && !distinct_done && (hc2[kCBucketOver] != 0 || hc2[kCSetOver] != 0)) {
        // rare: hash buckets did not fit (very uneven data or stale hints). Redo the distinct counts the general way.
        redo_distinct_general();
        if (early) {
          run_extract_late(stream);
        } else {
          CUDA_CK(cudaMemsetAsync(sc->d_counters + kCGroups, 0, 8, stream));
          run_extract(stream, true);
        }
        for (uint32_t d = 0; d < P.ndistinct; ++d) {
          const uint32_t m = P.distinct_met[d];
          if (m < plan->nmetrics && out_rows)
            CUDA_CK(cudaMemcpyAsync(const_cast<void *>(res->acc_ptrs[m]), d_accs[m], out_rows * q.accs[m].out_width, cudaMemcpyDeviceToHost, stream));
        }
        if (!early) {  // the whole extraction was redone: the order of the groups changed with it
          for (uint32_t k = 0; k < plan->nkeys; ++k)
            if (out_rows) CUDA_CK(cudaMemcpyAsync(const_cast<void *>(res->key_ptrs[k]), d_keys[k], out_rows * t->cols[plan->keys[k].col].width, cudaMemcpyDeviceToHost, stream));
          for (uint32_t m = 0; m < plan->nmetrics; ++m)
            if (out_rows) CUDA_CK(cudaMemcpyAsync(const_cast<void *>(res->acc_ptrs[m]), d_accs[m], out_rows * q.accs[m].out_width, cudaMemcpyDeviceToHost, stream));
          if (plan->need_hidden_count && out_rows)
            CUDA_CK(cudaMemcpyAsync(const_cast<uint64_t *>(view.hidden_count), d_accs[plan->nmetrics], out_rows * 8, cudaMemcpyDeviceToHost, stream));
        }
        CUDA_CK(cudaEventRecord(sc->ev_end, stream));
        CUDA_CK(cudaStreamSynchronize(stream));
      }
      hint_raise(t->groups_hint, ngroups);
      float total_ms = 0;
      CUDA_CK(cudaEventElapsedTime(&total_ms, sc->ev_begin, sc->ev_end));
      if (ctx->trace) {   // device time of the phases on s0 (what a multi-GPU step spends outside the scan)
        float a = 0, b = 0, c = 0, d = 0, e = 0;
        cudaEventElapsedTime(&a, sc->ev_begin, sc->ev_scan0);
        cudaEventElapsedTime(&b, sc->ev_scan0, sc->ev_scan1);
        cudaEventElapsedTime(&c, sc->ev_scan1, sc->ev_phase[0]);
        if (early && !dedupe_after_sync && have_rows) {
          cudaEventElapsedTime(&d, sc->ev_phase[0], sc->ev_phase[1]);
          cudaEventElapsedTime(&e, sc->ev_phase[1], sc->ev_end);
        } else {
          cudaEventElapsedTime(&e, sc->ev_phase[0], sc->ev_end);
        }
        cudaGetLastError();
        fprintf(stderr, "[vgpu r%d] phases ms: setup %.3f scan %.3f merge/exchange %.3f dedupe(+allreduce) %.3f extract+copy %.3f total %.3f\n",
                ctx->rank, a, b, c, d, e, total_ms);
      }
      view.ngroups = out_rows;
      view.aggregated_recs = ngroups;
      view.gpu_ms = total_ms;
      view.scan_ms = scan_ms_total;
      view.launches = launches;
      view.table_mode = q.wide ? 2 : (q.hash_mode ? 1 : 0);
      view.table_cells = q.ncells;
      view.attempts = (uint32_t)attempt + 1;
      view.distinct_paths = paths;
      view.post_applied = have_rows ? paths_post : 0;
      break;
    }

    view.keys = res->key_ptrs.empty() ? nullptr : res->key_ptrs.data();
    view.accs = res->acc_ptrs.empty() ? nullptr : res->acc_ptrs.data();
    *out = res.release();
  });
}
