// scan_kernel.cuh — the fused scan -> filter -> group-by-aggregate kernel (sm_100a).
//
// Reference semantics reproduced (paths relative to the viyadb/viyadb tree):
//   segment / tuple loops   src/codegen/query/scan.cc:40-73
//   predicate               src/codegen/query/filter.cc:206-261 (branch-free &,| over typed compares)
//   key build + rollup      src/codegen/query/scan.cc:193-224, src/codegen/db/rollup.cc:77-95
//   agg_map[key].Update(m)  src/codegen/query/scan.cc:228-242, src/codegen/db/store.cc:131-161
//   count-distinct          src/util/bitset.h:26-67 (set union; cardinality at output)
//
// Execution model. Warps are fully independent (no CTA barrier in the loop). Work is handed
// out dynamically in units = runs of consecutive 512-row chunks of ONE segment, so a warp keeps the segment's
// descriptor in registers and its queue of passing rows holds bare row numbers. A chunk is 4 sub-chunks of
// 128 rows, lane l holding rows l*4..l*4+3 of each, so that every predicate-column load is one 128-bit (u32),
// 64-bit (u16) or 32-bit (u8) request per lane and 512/256/128 contiguous bytes per warp instruction — the 4
// sub-chunk loads of a column are issued back to back. Per chunk:
//   1. one bulk L2 prefetch (cp.async.bulk.prefetch.L2, TMA unit) per predicate column pulls the NEXT chunk in;
//   2. the predicate runs on 16-row register vectors: a conjunction of up to 4 vectorisable leaves fully
//      unrolled on constant-bank operands, anything else through a small stack interpreter (uniform control flow);
//   3. passing rows are compacted with one warp prefix sum into the per-warp shared-memory queue;
//   4. full batches of 32 rows are aggregated one row per lane: all key and metric cells of a row are loaded
//      before anything depends on them — from the row-major mirror when the chunk was sparse (the cells of a
//      row share one or two 64-byte DRAM atoms), from the columns otherwise — the group cell is found
//      (mixed-radix index, or open-addressing probe on the packed 64-bit key) and the accumulators are updated
//      with native RED operations, in global memory or in a CTA-private shared-memory copy of a small dense
//      table; count-distinct ids are appended to a per-CTA region and deduplicated after the scan.
// Predicate columns are read exactly once, key / metric traffic is atom-granular in the selectivity, nothing is
// written but accumulators and count-distinct pairs.
#ifndef VGPU_SCAN_KERNEL_CUH_
#define VGPU_SCAN_KERNEL_CUH_

#include "kernels.cuh"

namespace vgpu {

constexpr int kChunkRows = 512;                 // rows per warp iteration
constexpr int kSubChunk = kChunkRows / kSub;    // 128
constexpr int kWarps = kThreads / 32;     // select / search kernels
// The fused scan kernel's own CTA shape: threads per CTA x resident CTAs per SM fixes the register budget
// (65536 / (threads x CTAs), rounded down to 8). Measured on B200 (tools/gpu_ab.sh): at 256 x 3 (80 registers) the
// kernel sits on a spill cliff — harmless edits moved the C2 scan between 2.40 and 2.76 ms.
#ifndef VGPU_SCAN_THREADS
#define VGPU_SCAN_THREADS 256
#endif
constexpr int kScanThreads = VGPU_SCAN_THREADS;
constexpr int kScanWarps = kScanThreads / 32;
#ifndef VGPU_PF_DIST
#define VGPU_PF_DIST 1
#endif
constexpr uint32_t kPrefetchDist = VGPU_PF_DIST;  // how many chunks ahead of its loads a warp prefetches into L2
constexpr uint32_t kSmemTableBytes = 40 * 1024;  // CTA-private group table (3 CTAs/SM: 3 x (17 + 40) KB fit 227 KB)
constexpr int kListCap = kChunkRows + 32;       // a chunk's worth of rows plus one incomplete batch
static_assert(kVec * 32 == kSubChunk, "a lane owns kVec consecutive rows of every sub-chunk");
static_assert(kTileRows % kChunkRows == 0, "slab capacity is a whole number of chunks");

// ---------------------------------------------------------------------------------------------
// predicate interpreter: 16 rows per lane, bit (s*4+j) of the result = row s*128 + lane*4 + j
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_vec16(const uint8_t *col, uint32_t width, uint32_t row0,
                                           uint32_t (&v)[kRowsPerThread], uint64_t pol) {
  if (width == 4) {
    uint4 q[kSub];
#pragma unroll
    for (int s = 0; s < kSub; ++s) q[s] = ldg_stream128(col + (uint64_t)(row0 + s * kSubChunk) * 4, pol);
#pragma unroll
    for (int s = 0; s < kSub; ++s) {
      v[s * 4 + 0] = q[s].x; v[s * 4 + 1] = q[s].y; v[s * 4 + 2] = q[s].z; v[s * 4 + 3] = q[s].w;
    }
  } else if (width == 2) {
    uint2 q[kSub];
#pragma unroll
    for (int s = 0; s < kSub; ++s) q[s] = ldg_stream64(col + (uint64_t)(row0 + s * kSubChunk) * 2, pol);
#pragma unroll
    for (int s = 0; s < kSub; ++s) {
      v[s * 4 + 0] = q[s].x & 0xffffu; v[s * 4 + 1] = q[s].x >> 16;
      v[s * 4 + 2] = q[s].y & 0xffffu; v[s * 4 + 3] = q[s].y >> 16;
    }
  } else {
    uint32_t q[kSub];
#pragma unroll
    for (int s = 0; s < kSub; ++s) q[s] = ldg_stream32(col + (uint64_t)(row0 + s * kSubChunk), pol);
#pragma unroll
    for (int s = 0; s < kSub; ++s) {
      v[s * 4 + 0] = q[s] & 0xffu; v[s * 4 + 1] = (q[s] >> 8) & 0xffu;
      v[s * 4 + 2] = (q[s] >> 16) & 0xffu; v[s * 4 + 3] = q[s] >> 24;
    }
  }
}

__device__ __forceinline__ uint64_t bitset_card(const SegDesc &seg, uint32_t bidx, uint64_t row) {
  const uint32_t *off = seg.bs_offsets[bidx];
  if (off == nullptr) return 1;
  return (uint64_t)(__ldg(off + row + 1) - __ldg(off + row));
}

// the segment a warp is working in: a copy of its descriptor in registers (warp-uniform)
struct CurSeg {
  const uint8_t *slab;
  const uint8_t *rows;
  uint32_t cap, nrows, index;
};

// row0 = first row of the lane inside the segment (chunk_row0 + lane*4)
// kConjOnly: the caller knows the predicate is an unrolled conjunction (P.conj): the interpreter is not compiled in
template <bool kConjOnly = false>
__device__ __forceinline__ uint32_t eval_predicate(const ScanParams &P, const CurSeg &seg,
                                                   uint32_t row0, uint64_t pol) {
  uint32_t v[kRowsPerThread];
  if (kConjOnly || P.conj) {
    // The common shape — a conjunction of up to 4 vectorisable leaves — with the program counter
    // unrolled: every operand of a leaf is a constant-bank operand, nothing is interpreted.
    uint32_t m = 0xffffu;
#pragma unroll
    for (int pc = 0; pc < 4; ++pc) {
      if (pc < P.nprog) {
        const PInstr &in = P.prog[pc];
        if (pc == 0 || in.slot != P.prog[pc > 0 ? pc - 1 : 0].slot) {
          const Slot &sl = P.slots[in.slot];
          load_vec16(seg.slab + sl.off * seg.cap, sl.width, row0, v, pol);
        }
        uint32_t lm = leaf_mask16(in, v);
        if (in.neg) lm ^= 0xffffu;
        m &= lm;
      }
    }
    return m;
  }
  if constexpr (kConjOnly) {
    return 0;
  } else {
  uint32_t stk[kStackDepth];
#pragma unroll
  for (int i = 0; i < kStackDepth; ++i) stk[i] = 0;
  int cached = -1;
  for (uint32_t pc = 0; pc < P.nprog; ++pc) {
    const PInstr &in = P.prog[pc];
    const uint32_t kind = in.kind;
    if (kind <= P_OR_LEAF) {
      uint32_t m = 0;
      const uint32_t cls = in.cls;
      if (cls == C_TRUE) {
        m = 0xffffu;
      } else if (cls == C_FALSE) {
        m = 0;
      } else if (cls == C_GEN) {
        const Slot &sl = P.slots[in.slot];
        const SegDesc &sd = P.segs[seg.index];
#pragma unroll 1
        for (int s = 0; s < kSub; ++s) {
#pragma unroll 1
          for (int j = 0; j < kVec; ++j) {
            uint32_t row = row0 + s * kSubChunk + j;
            uint64_t val = 0;
            if (row < seg.nrows) {  // scalar path must not read past the logical end of CSR tables
              val = sl.bitset ? bitset_card(sd, sl.bitset_idx, row)
                              : load_elem(seg.slab + sl.off * seg.cap + (uint64_t)row * sl.width, sl.width, sl.sext);
            }
            if (gen_compare(in.gcls, in.gop, val, in.arg)) m |= 1u << (s * 4 + j);
          }
        }
      } else {
        if ((int)in.slot != cached) {
          const Slot &sl = P.slots[in.slot];
          load_vec16(seg.slab + sl.off * seg.cap, sl.width, row0, v, pol);
          cached = in.slot;
        }
        m = leaf_mask16(in, v);
      }
      if (in.neg) m ^= 0xffffu;
      if (kind == P_PUSH) {
#pragma unroll
        for (int i = kStackDepth - 1; i > 0; --i) stk[i] = stk[i - 1];
        stk[0] = m;
      } else if (kind == P_AND_LEAF) {
        stk[0] &= m;
      } else {
        stk[0] |= m;
      }
    } else {
      uint32_t r = (kind == P_AND) ? (stk[1] & stk[0]) : (stk[1] | stk[0]);
      stk[0] = r;
#pragma unroll
      for (int i = 1; i < kStackDepth - 1; ++i) stk[i] = stk[i + 1];
    }
  }
  return stk[0];
  }
}

// Wide key tuples (hash_mode 2): gather the key words of one row and find / claim its slot. Out of line
// so that its local array does not cost the common paths registers.
__device__ __noinline__ uint64_t wide_row_cell(const ScanParams &P, const CurSeg &seg, uint32_t row) {
  uint64_t kw[kMaxKeys];
  for (uint32_t k = 0; k < P.nkeys; ++k) {
    const KeySpec &ks = P.keys[k];
    const Slot &sl = P.slots[ks.slot];
    const uint8_t *a = seg.slab + sl.off * seg.cap + (uint64_t)row * sl.width;
    uint64_t val = gather_finish(gather_raw64(a), a, sl.vmask, sl.signbit);
    if (ks.rollup) val = rollup_value(val, ks);
    if (ks.fzero) val = fzero_fix(val, sl.width);
    kw[k] = val;
  }
  return wide_cell(P, kw);
}

// Key tuples outside the register-staged fast path (more than 4 keys, or 8-byte keys): gather, roll up and pack one
// row's key. Out of line: its registers must not weigh on the common path (the kernel is capped at 80).
template <bool kPlainKeys>
__device__ __noinline__ uint64_t general_row_key(const ScanParams &P, const CurSeg &seg, uint32_t row) {
  uint64_t packed = 0;
  for (uint32_t k = 0; k < P.nkeys; ++k) {
    const KeySpec &ks = P.keys[k];
    const Slot &sl = P.slots[ks.slot];
    const uint8_t *a = seg.slab + sl.off * seg.cap + (uint64_t)row * sl.width;
    uint64_t val = gather_finish(gather_raw64(a), a, sl.vmask, sl.signbit);
    if (!kPlainKeys) {
      if (P.tdict.npieces && k == P.tdict.key) val = tdict_rank(P.tdict, val);
      else if (ks.rollup) val = rollup_value(val, ks);
      if (ks.fzero) val = fzero_fix(val, sl.width);
    }
    if (ks.lut) val = (uint64_t)__popcll(ks.lut & ((1ull << val) - 1ull));
    else val -= ks.lo;
    packed += val * ks.mul;
  }
  return packed;
}
__device__ __noinline__ uint64_t hash_cell_call(const ScanParams &P, uint64_t key) { return hash_cell(P, key); }

// append one count-distinct pair to the CTA's region of the pair's owner rank; `cursor` = s_cursor[dn]
__device__ __forceinline__ void append_pair(const ScanParams &P, uint32_t *cursor, uint32_t dn, uint64_t hi, uint64_t id) {
  uint32_t sub = 0;
  if (P.dpair_nsub > 1) sub = P.dpair_key ? owner_of(hi, 0, P.dpair_nsub) : pair_owner(hi, id, P.dpair_nsub);
  const uint32_t pos = atomicAdd(cursor + sub, 1u);
  if (pos < P.dpair_cap) {
    const uint64_t at = ((uint64_t)sub * gridDim.x + blockIdx.x) * P.dpair_cap + pos;
    if (P.dpair_wide) reinterpret_cast<ulonglong2 *>(P.dpairs[dn])[at] = make_ulonglong2(hi, id);
    else P.dpairs[dn][at] = (hi << 32) | id;
  }
}
// a bitset cell that holds several ids (CSR): rare after ingesting raw rows, out of line
__device__ __noinline__ void append_csr_cell(const ScanParams &P, uint32_t *cursor, uint32_t dn, uint64_t hi, const uint32_t *vals,
                                             uint32_t lo, uint32_t hi_q, uint32_t id64) {
  for (uint32_t q = lo; q < hi_q; ++q)
    append_pair(P, cursor, dn, hi, id64 ? __ldg(reinterpret_cast<const unsigned long long *>(vals) + q) : (uint64_t)gather_u32(vals + q));
}

// ---------------------------------------------------------------------------------------------
// the fused scan kernel
// ---------------------------------------------------------------------------------------------
#ifndef VGPU_MIN_CTAS
#define VGPU_MIN_CTAS 3
#endif
// kMinCtas CTAs per SM (VGPU_MIN_CTAS, compile time): with kScanThreads it caps the registers (256 x 3 -> 80)
// kSmemTable: the instantiation that aggregates into a CTA-private shared-memory copy of a small dense group
// table (ScanParams::smem_cells != 0); the other one carries none of that code.
// kPlainKeys: no group key needs a time rollup, a bucket-dictionary lookup or the -0.0 fix (the common case: dictionary
// codes and plain integers): the per-key checks for them are compiled out of the per-row path, which is issue-bound.
// kFast: at most 4 keys of at most 4 bytes, at most 4 metrics, no wide key tuples, 8-byte count-distinct pairs — what
// nearly every query is. This instantiation holds nothing else: no call, no key loop over memory, no 16-byte pairs. The
// kernel is capped at 80 registers and sat on a spill cliff when the rare paths shared its code (tools/gpu_ab.sh:
// harmless edits moved the C2 scan between 2.40 and 2.76 ms); the other instantiation takes everything.
// kConj (fast instantiations only): the predicate is the unrolled conjunction; no interpreter, hence no call at all.
template <int kMinCtas, bool kSmemTable, bool kPlainKeys, bool kFast, bool kConj = false>
__global__ void __launch_bounds__(kScanThreads, kMinCtas)
scan_filter_groupby_kernel(const __grid_constant__ ScanParams P) {
  __shared__ uint32_t s_list[kScanWarps][kListCap];  // rows (of the warp's current segment) waiting for aggregation

  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t *list = s_list[warp];
  uint32_t npend = 0;  // rows waiting in the list (warp-uniform)
  CurSeg seg{};        // the segment of the warp's current work unit
  uint32_t my_passed = 0;
  __shared__ uint32_t s_cursor[kMaxDistinct][kMaxRanks];  // pairs this CTA appended per count-distinct metric and owner rank
  if (threadIdx.x < kMaxDistinct * kMaxRanks) (&s_cursor[0][0])[threadIdx.x] = 0;
  // CTA-private group table (small dense domains): every cell starts as the plan's initial image
  extern __shared__ __align__(16) uint8_t s_table[];
  const uint32_t s_table_a = (uint32_t)__cvta_generic_to_shared(s_table);
  if (kSmemTable) {
    const uint32_t wpc = P.smem_stride / 4, nwords = P.smem_cells * wpc;
    for (uint32_t i = threadIdx.x; i < nwords; i += kScanThreads) reinterpret_cast<uint32_t *>(s_table)[i] = P.smem_init[i % wpc];
  }
  __syncthreads();
  const bool can_overflow = P.hash_mode != 0;
  const uint64_t pol = make_stream_policy((P.tune & 2u) != 0);
  const uint64_t tpol = make_table_policy((P.tune & 4u) != 0);

  // ---- aggregate one passing row of the current segment (one row per lane) ----
  // fast path: up to 4 keys of at most 4 bytes and up to 4 metrics, staged in 12 registers
  const bool small_plan = kFast || (P.small_plan != 0 && P.hash_mode != 2);  // uniform
  auto process_row = [&](const uint32_t row, const bool rowpath) {
    uint32_t kv[4];
    uint64_t mv[4];
    const SegDesc &sd = P.segs[seg.index];  // side tables of bitset columns (uniform address)
    // position of a cell in bytes from an 8-byte aligned base: inside its column, or inside the mirror
    const uint32_t rpos = row * P.row_stride;
    if (small_plan) {
      // one DRAM round trip per row: every key and metric cell (or the first word a bitset cell
      // needs) is requested before anything depends on it; fully unrolled, so the loads issue back to
      // back and the values stay in registers. Slab, mirror and side-table bases are 8-byte aligned, so the
      // position of a cell inside its aligned word depends on the row only.
      if (rowpath) {  // warp-uniform: the cells of a row share one or two 64-byte atoms of the mirror
        const uint8_t *rb = seg.rows + (uint64_t)row * P.row_stride;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (k < P.nkeys) kv[k] = gather_raw32(rb + P.keys[k].row_off);
#pragma unroll
        for (int m = 0; m < 4; ++m)
          if (m < P.nmetrics) mv[m] = gather_raw64(rb + P.mets[m].row_off);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (k < P.nkeys) kv[k] = gather_raw32(seg.slab + P.keys[k].col_off * seg.cap + (uint64_t)row * P.keys[k].width);
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          if (m < P.nmetrics) {
            const MetSpec &ms = P.mets[m];
            const uint8_t *a = seg.slab + ms.col_off * seg.cap + (uint64_t)row * ms.width;
            if (ms.bitset) {
              const uint32_t *off = sd.bs_offsets[ms.bitset_idx];
              a = reinterpret_cast<const uint8_t *>((off == nullptr ? sd.bs_values[ms.bitset_idx] : off) + row);
            }
            mv[m] = gather_raw64(a);
          }
        }
      }
    }
    uint64_t packed = 0;
    if (!kFast && P.hash_mode == 2) {
      packed = wide_row_cell(P, seg, row);  // rare path, out of line
    } else if (small_plan) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (k < P.nkeys) {
          const KeySpec &ks = P.keys[k];
          const uint32_t pos = rowpath ? ks.row_off : row * ks.width;
          uint64_t val = ((uint64_t)(kv[k] >> ((pos & 3u) * 8u))) & ks.vmask;
          val = (val ^ ks.signbit) - ks.signbit;
          if (!kPlainKeys) {
            if (P.tdict.npieces && k == P.tdict.key) val = tdict_rank(P.tdict, val);  // rank of the rolled-up value
            else if (ks.rollup) val = rollup_value(val, ks);
            if (ks.fzero) val = fzero_fix(val, 4);
          }
          if (ks.lut) val = (uint64_t)__popcll(ks.lut & ((1ull << val) - 1ull));  // rank inside the IN list
          else val -= ks.lo;
          packed += val * ks.mul;
        }
      }
    } else if (!kFast) {
      packed = general_row_key<kPlainKeys>(P, seg, row);
    }
    uint64_t cell;
    if (!kFast && P.hash_mode == 2) {
      cell = packed;
      if (cell == kEmptyKey) {
        atomicOr(&P.counters[kCHashOver], 1ull);
        return;
      }
    } else if (P.hash_mode) {
      cell = kFast ? hash_cell(P, packed) : hash_cell_call(P, packed);
      if (cell == kEmptyKey) {
        atomicOr(&P.counters[kCHashOver], 1ull);
        return;
      }
    } else {
      cell = packed;
      if (kSmemTable) *reinterpret_cast<volatile uint32_t *>(s_table + (uint32_t)cell * P.smem_stride + P.smem_present_off) = 1u;
      else if (!P.skip_present) st_u8_hint(P.present + cell * P.present_stride, 1u, tpol);
    }
    uint32_t dn = 0;
    for (uint32_t m = 0; m < P.nmetrics; ++m) {
      const MetSpec &ms = P.mets[m];
      uint64_t pre;
      if (small_plan) {
        const uint64_t raw = m == 0 ? mv[0] : m == 1 ? mv[1] : m == 2 ? mv[2] : mv[3];
        const uint32_t pos = rowpath ? rpos + ms.row_off : row * ms.width;
        pre = (raw >> ((pos & 7u) * 8u)) & ms.vmask;
        pre = (pre ^ ms.signbit) - ms.signbit;
      } else if (!kFast) {
        const uint32_t *off = sd.bs_offsets[ms.bitset_idx];
        const uint8_t *a = ms.bitset ? reinterpret_cast<const uint8_t *>((off == nullptr ? sd.bs_values[ms.bitset_idx] : off) + row)
                                     : seg.slab + ms.col_off * seg.cap + (uint64_t)row * ms.width;
        pre = gather_finish(gather_raw64(a), a, ms.vmask, ms.signbit);
      } else {
        pre = 0;
      }
      if (ms.op != A_DISTINCT) {
        if (kSmemTable) acc_update_shared(s_table_a + (uint32_t)cell * P.smem_stride + ms.soff, ms.op, pre);
        else acc_update(reinterpret_cast<uint8_t *>(ms.acc) + cell * ms.stride, ms.op, pre, tpol);
        continue;
      }
      // count-distinct: append (cell, id) to this CTA's region (of the pair's owner rank); deduplicated after the scan
      const uint32_t *off = sd.bs_offsets[ms.bitset_idx];
      if (kFast) {   // 8-byte pairs (cell, 32-bit id); one sub-region per owner rank when several GPUs take part
        const uint32_t *vals = sd.bs_values[ms.bitset_idx];
        if (off == nullptr) {  // one id per row: `pre` is the id
          const uint32_t sub = P.dpair_nsub > 1 ? pair_owner(cell, pre, P.dpair_nsub) : 0u;
          const uint32_t pos = atomicAdd(&s_cursor[dn][sub], 1u);
          if (pos < P.dpair_cap) P.dpairs[dn][((uint64_t)sub * gridDim.x + blockIdx.x) * P.dpair_cap + pos] = (cell << 32) | pre;
        } else {               // CSR cell: `pre` is offsets[row]
          const uint32_t lo = (uint32_t)pre, hi_q = gather_u32(off + row + 1);
          for (uint32_t q = lo; q < hi_q; ++q) {
            const uint64_t id = gather_u32(vals + q);
            const uint32_t sub = P.dpair_nsub > 1 ? pair_owner(cell, id, P.dpair_nsub) : 0u;
            const uint32_t pos = atomicAdd(&s_cursor[dn][sub], 1u);
            if (pos < P.dpair_cap) P.dpairs[dn][((uint64_t)sub * gridDim.x + blockIdx.x) * P.dpair_cap + pos] = (cell << 32) | id;
          }
        }
      } else {
        const uint64_t hi = P.dpair_key ? packed : cell;
        if (off == nullptr) append_pair(P, s_cursor[dn], dn, hi, pre);   // one id per row: `pre` is the id
        else append_csr_cell(P, s_cursor[dn], dn, hi, sd.bs_values[ms.bitset_idx], (uint32_t)pre, gather_u32(off + row + 1), ms.id64);
      }
      ++dn;
    }
  };

  // the one place process_row is instantiated: `count` rows of the list, one per lane and 32 per batch
  auto batches = [&](const uint32_t count, const bool rowpath) {
#pragma unroll 1
    for (uint32_t head = 0; head < count; head += 32)
      if (head + lane < count) process_row(list[head + lane], rowpath);
  };

  // ---- work units: runs of P.unit_chunks consecutive chunks of one segment, handed out dynamically ----
  // A warp keeps its segment's descriptor in registers for a whole unit and the list holds bare row numbers;
  // units are taken in table order, so all warps together sweep a window of a few dozen segments.
  const uint32_t ups = P.units_per_seg, nunits = P.nactive * ups;
  auto grab = [&]() {
    uint32_t u = 0;
    if (lane == 0) u = atomicAdd(reinterpret_cast<unsigned int *>(&P.counters[kCUnit]), 1u);
    return __shfl_sync(0xffffffffu, u, 0);
  };
  uint32_t unit = grab();
  while (unit < nunits) {
    const uint32_t next_unit = grab();  // one ahead: its first chunk is prefetched from this unit's last one
    // a full group table / distinct set makes the host grow it and run again: stop wasting time
    if (can_overflow) {
      unsigned long long f = 0;
      if (lane == 0) f = *reinterpret_cast<volatile unsigned long long *>(&P.counters[kCHashOver]);
      if (__shfl_sync(0xffffffffu, f, 0) != 0ull) break;
    }
    const uint32_t si = unit / ups, part = unit - si * ups;
    {
      seg.index = P.active[si];
      const SegDesc &sd = P.segs[seg.index];
      seg.slab = sd.slab;
      seg.rows = sd.rows;
      seg.cap = (uint32_t)sd.cap;
      seg.nrows = (uint32_t)sd.nrows;
    }
    const uint32_t nrows = seg.nrows;
    const uint32_t c_begin = part * P.unit_chunks;
    const uint32_t c_end = min(c_begin + P.unit_chunks, (nrows + kChunkRows - 1) / kChunkRows);
    for (uint32_t ci = c_begin; ci < c_end; ++ci) {
    const uint32_t chunk_row = ci * kChunkRows;
    const uint32_t row0 = chunk_row + lane * kVec;

    // software pipeline: pull the predicate columns of this warp's NEXT chunk into L2 now (one bulk prefetch,
    // TMA unit, per column: lane f takes column f), so that its vector loads find them there instead of
    // paying a DRAM round trip per column. The chunk after the last one of a unit opens the next unit.
    if (!(P.tune & 32u) && lane < P.nfilter_slots) {
      const uint8_t *a = nullptr;
      if (ci + kPrefetchDist < c_end) {
        a = seg.slab + P.pf_off[lane] * seg.cap + (uint64_t)(chunk_row + kPrefetchDist * kChunkRows) * P.pf_width[lane];
      } else if (next_unit < nunits && (kPrefetchDist == 1 || ci + kPrefetchDist - c_end < P.unit_chunks)) {
        const uint32_t nsi = next_unit / ups, npart = next_unit - nsi * ups;
        const SegDesc &nsd = P.segs[P.active[nsi]];
        const uint32_t ahead = kPrefetchDist == 1 ? 0u : ci + kPrefetchDist - c_end;  // chunk of the next unit
        const uint32_t nrow = (npart * P.unit_chunks + ahead) * kChunkRows;
        if (nrow < (uint32_t)nsd.nrows) a = nsd.slab + P.pf_off[lane] * nsd.cap + (uint64_t)nrow * P.pf_width[lane];
      }
      if (a != nullptr)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"((uint32_t)kChunkRows * P.pf_width[lane]) : "memory");
    }

    uint32_t mask = eval_predicate<kConj>(P, seg, row0, pol);
    // rows past the end of a partially filled segment never count
    if (chunk_row + kChunkRows > nrows) {
#pragma unroll
      for (int s = 0; s < kSub; ++s) {
#pragma unroll
        for (int j = 0; j < kVec; ++j) {
          if (row0 + s * kSubChunk + j >= nrows) mask &= ~(1u << (s * 4 + j));
        }
      }
    }
    if (__ballot_sync(0xffffffffu, mask != 0) == 0) continue;
    my_passed += __popc(mask);

    // ---- compaction: one warp prefix sum of the per-lane counts, rows appended to the list ----
    const uint32_t n = __popc(mask);
    uint32_t incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    {
      uint32_t pos = npend + incl - n;
      const uint32_t base_row = chunk_row + lane * kVec;
      if (total <= 96) {  // sparse: a lane holds few passing rows, visit the set bits only
        while (mask) {
          const uint32_t b = __ffs(mask) - 1;
          mask &= mask - 1;
          list[pos++] = base_row + (b >> 2) * kSubChunk + (b & 3u);
        }
      } else {
#pragma unroll
        for (int b = 0; b < kRowsPerThread; ++b)
          if (mask & (1u << b)) list[pos++] = base_row + (b >> 2) * kSubChunk + (b & 3);
      }
    }
    __syncwarp();
    const bool rowpath = total <= P.row_thresh;  // sparse chunk: gather from the row-major mirror
    npend += total;

    // ---- hand the passing rows to the aggregation stage in full batches of 32 ----
    // The list is a small per-warp queue: rows wait (across the chunks of a unit) until a whole warp of
    // them is available, so that one DRAM round trip of gathers always serves 32 rows whatever the selectivity.
    const uint32_t head = npend & ~31u;
    if (head) {
      batches(head, rowpath);
      const uint32_t left = npend - head;  // move the incomplete batch (< 32 rows) to the front
      uint32_t t = 0;
      if (lane < left) t = list[head + lane];
      __syncwarp();
      if (lane < left) list[lane] = t;
      __syncwarp();
      npend = left;
    }
    }  // chunks of the unit
    // the unit's last, incomplete batch (the next unit is in another segment)
    if (npend) {
      batches(npend, P.row_thresh != 0);
      npend = 0;
      __syncwarp();
    }
    unit = next_unit;
  }

  // merge the CTA-private table into the global one: the same commutative Update(), once per live cell
  if (kSmemTable) {
    __syncthreads();
    for (uint32_t c = threadIdx.x; c < P.smem_cells; c += kScanThreads) {
      const uint8_t *cellp = s_table + c * P.smem_stride;
      if (*reinterpret_cast<const uint32_t *>(cellp + P.smem_present_off) == 0) continue;
      st_u8_hint(P.present + (uint64_t)c * P.present_stride, 1u, tpol);
      for (uint32_t m = 0; m < P.nmetrics; ++m) {
        const MetSpec &ms = P.mets[m];
        if (ms.op == A_DISTINCT) continue;
        const uint64_t v = ms.acc_width == 4 ? (uint64_t)*reinterpret_cast<const uint32_t *>(cellp + ms.soff)
                                             : *reinterpret_cast<const uint64_t *>(cellp + ms.soff);
        acc_update(reinterpret_cast<uint8_t *>(ms.acc) + (uint64_t)c * ms.stride, ms.op,
                   ms.acc_width == 4 && (ms.op == A_ADD32 || ms.op == A_MINS32 || ms.op == A_MAXS32) ? (uint64_t)(int64_t)(int32_t)(uint32_t)v : v, tpol);
      }
    }
  }

  // counters: one atomic per warp
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) my_passed += __shfl_down_sync(0xffffffffu, my_passed, o);
  if (lane == 0 && my_passed) atomicAdd(&P.counters[kCPassed], (unsigned long long)my_passed);
  if (P.ndistinct) {  // the only CTA-wide barrier of the kernel: publish the per-CTA pair counts
    __syncthreads();
    if (threadIdx.x < P.ndistinct * P.dpair_nsub) {
      const uint32_t d = threadIdx.x / P.dpair_nsub, sub = threadIdx.x - d * P.dpair_nsub;
      const uint32_t n = s_cursor[d][sub];
      P.dpair_count[d][sub * gridDim.x + blockIdx.x] = n < P.dpair_cap ? n : P.dpair_cap;
      if (n > P.dpair_cap) atomicOr(&P.counters[kCRegionOver], 1ull);  // the host grows the regions and scans again
      atomicAdd(&P.counters[kCPairs + d], (unsigned long long)n);
      atomicMax(&P.counters[kCMaxFill + d], (unsigned long long)n);  // fullest region: what the next query sizes them by
    }
  }
}

}  // namespace vgpu

#endif  // VGPU_SCAN_KERNEL_CUH_
