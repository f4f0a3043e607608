// kernels.cuh — hand-written sm_100a kernels of the scan -> filter -> group-by-aggregate path.
//
// Reference semantics each piece reproduces (paths relative to the viyadb/viyadb tree):
//   predicate        src/codegen/query/filter.cc:206-261 (branch-free &,| over typed compares)
//   key build        src/codegen/query/scan.cc:193-224
//   time rollup      src/codegen/db/rollup.cc:77-95, src/util/time.h:52-137 (gmtime_r/timegm, UTC)
//   Update()         src/codegen/db/store.cc:131-161 (+=, std::min, std::max, |=)
//   count-distinct   src/util/bitset.h:26-67 (set union, cardinality)
//
// This header holds the building blocks (streaming / gather loads, UTC calendar arithmetic, accumulator
// updates, hash cells; the pure arithmetic — calendar, rollup, hash mixers — is in device_arith.h, which also compiles
// for the host) and every kernel but the fused scan itself (scan_kernel.cuh): count-distinct partition /
// dedupe, multi-GPU exchange, group extraction, row-mirror build, column statistics, the synthetic generator.
#ifndef VGPU_KERNELS_CUH_
#define VGPU_KERNELS_CUH_

#include "scan_params.h"
#include "device_arith.h"
#include <cuda_runtime.h>

namespace vgpu {

constexpr uint64_t kEmptyKey = ~0ull;

// ---------------------------------------------------------------------------------------------
// streaming loads (filter columns are read exactly once: keep them out of L1)
// ---------------------------------------------------------------------------------------------
// An L2 evict-first policy: column data is read exactly once, it must not push the group table
// (pinned with an access-policy window) or anything else out of L2.
__device__ __forceinline__ uint64_t make_stream_policy(bool evict_first) {
  uint64_t pol;
  if (evict_first) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint4 ldg_stream128(const void *p, uint64_t pol) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ uint2 ldg_stream64(const void *p, uint64_t pol) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;"
               : "=r"(r.x), "=r"(r.y) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ uint32_t ldg_stream32(const void *p, uint64_t pol) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
  return r;
}

// Scalar load of one element, widened to 64 bit (zero- or sign-extended). Goes through L1 so the
// 2-8 rows sharing a 32-byte sector are served from one HBM sector.
__device__ __forceinline__ uint64_t load_elem(const uint8_t *p, uint32_t width, uint32_t sext) {
  switch (width) {
    case 1: {
      uint8_t v = __ldg(p);
      return sext ? (uint64_t)(int64_t)(int8_t)v : (uint64_t)v;
    }
    case 2: {
      uint16_t v = __ldg(reinterpret_cast<const uint16_t *>(p));
      return sext ? (uint64_t)(int64_t)(int16_t)v : (uint64_t)v;
    }
    case 4: {
      uint32_t v = __ldg(reinterpret_cast<const uint32_t *>(p));
      return sext ? (uint64_t)(int64_t)(int32_t)v : (uint64_t)v;
    }
    default:
      return __ldg(reinterpret_cast<const unsigned long long *>(p));
  }
}

// Gather flavour of load_elem for the cells of passing rows. Measured on B200 (tools/gather_probe.cu):
// a plain ld.global(.nc) miss makes L1 request the whole 128-byte line (4 sectors, ~125 B of DRAM
// traffic per 4-byte gather, whatever cudaLimitMaxL2FetchGranularity says); the .L2::64B qualifier
// halves that to one 64-byte DRAM atom. PTX offers no smaller prefetch size.
__device__ __forceinline__ uint64_t gather_elem(const uint8_t *p, uint32_t width, uint32_t sext) {
  switch (width) {
    case 1: {
      uint32_t v;
      asm volatile("ld.global.nc.L2::64B.u8 %0, [%1];" : "=r"(v) : "l"(p));
      return sext ? (uint64_t)(int64_t)(int8_t)v : (uint64_t)(v & 0xffu);
    }
    case 2: {
      uint32_t v;
      asm volatile("ld.global.nc.L2::64B.u16 %0, [%1];" : "=r"(v) : "l"(p));
      return sext ? (uint64_t)(int64_t)(int16_t)v : (uint64_t)(v & 0xffffu);
    }
    case 4: {
      uint32_t v;
      asm volatile("ld.global.nc.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(p));
      return sext ? (uint64_t)(int64_t)(int32_t)v : (uint64_t)v;
    }
    default: {
      unsigned long long v;
      asm volatile("ld.global.nc.L2::64B.u64 %0, [%1];" : "=l"(v) : "l"(p));
      return v;
    }
  }
}
// Branch-free gather of one cell of any width: load the aligned 8-byte word that contains it and
// shift / mask / sign-extend with per-slot constants. No control flow, so the loads of all key and
// metric cells of a row issue back to back (a switch on the width would put every load in its own
// basic block and serialise the DRAM round trips — measured: 8 equal stall shares, ncu r1 v3).
// Column bases are 4096-byte aligned and slab capacities whole tiles, so the aligned word is always
// inside the column.
__device__ __forceinline__ uint64_t gather_raw64(const uint8_t *p) {
  unsigned long long v;
  asm("ld.global.nc.L2::64B.u64 %0, [%1];" : "=l"(v) : "l"(reinterpret_cast<uintptr_t>(p) & ~7ull));
  return v;
}
__device__ __forceinline__ uint64_t gather_finish(uint64_t raw, const uint8_t *p, uint64_t vmask, uint64_t signbit) {
  uint64_t v = (raw >> ((reinterpret_cast<uintptr_t>(p) & 7u) * 8u)) & vmask;
  return (v ^ signbit) - signbit;
}
// 32-bit flavour for cells of at most 4 bytes (dictionary codes, time, int metrics): half the registers.
__device__ __forceinline__ uint32_t gather_raw32(const uint8_t *p) {
  uint32_t v;
  asm("ld.global.nc.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(reinterpret_cast<uintptr_t>(p) & ~3ull));
  return v;
}
__device__ __forceinline__ uint32_t gather_u32(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.global.nc.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}


// ---------------------------------------------------------------------------------------------
// accumulator update == Metrics::Update (store.cc:131-161), on native atomics
// ---------------------------------------------------------------------------------------------
// All accumulator traffic carries an L2 cache-policy operand (evict_last when the table fits L2): the
// columns stream through L2 and, with default priorities, evict the accumulator lines so that almost
// every RED goes to DRAM (ncu r1: 87 % of RED sectors missed L2 at 5e5 groups).
__device__ __forceinline__ uint64_t make_table_policy(bool evict_last) {
  uint64_t pol;
  if (evict_last) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
#define VGPU_RED(op_type, cst, addr, val, pol) \
  asm volatile("red.global." op_type ".L2::cache_hint [%0], %1, %2;" ::"l"(addr), cst(val), "l"(pol) : "memory")
// PTX wants the type last: red.global.add.L2::cache_hint.u32
#define VGPU_RED2(op, type, cst, addr, val, pol) \
  asm volatile("red.global." op ".L2::cache_hint." type " [%0], %1, %2;" ::"l"(addr), cst(val), "l"(pol) : "memory")

__device__ __forceinline__ void red_min_s32(void *a, int v, uint64_t pol) { VGPU_RED2("min", "s32", "r", a, v, pol); }
__device__ __forceinline__ void red_max_s32(void *a, int v, uint64_t pol) { VGPU_RED2("max", "s32", "r", a, v, pol); }
__device__ __forceinline__ void red_min_u32(void *a, uint32_t v, uint64_t pol) { VGPU_RED2("min", "u32", "r", a, v, pol); }
__device__ __forceinline__ void red_max_u32(void *a, uint32_t v, uint64_t pol) { VGPU_RED2("max", "u32", "r", a, v, pol); }
__device__ __forceinline__ void red_min_s64(void *a, long long v, uint64_t pol) { VGPU_RED2("min", "s64", "l", a, v, pol); }
__device__ __forceinline__ void red_max_s64(void *a, long long v, uint64_t pol) { VGPU_RED2("max", "s64", "l", a, v, pol); }
__device__ __forceinline__ void red_min_u64(void *a, unsigned long long v, uint64_t pol) { VGPU_RED2("min", "u64", "l", a, v, pol); }
__device__ __forceinline__ void red_max_u64(void *a, unsigned long long v, uint64_t pol) { VGPU_RED2("max", "u64", "l", a, v, pol); }
__device__ __forceinline__ void red_add_u32(void *a, uint32_t v, uint64_t pol) { VGPU_RED2("add", "u32", "r", a, v, pol); }
__device__ __forceinline__ void red_add_u64(void *a, unsigned long long v, uint64_t pol) { VGPU_RED2("add", "u64", "l", a, v, pol); }
__device__ __forceinline__ void red_add_f32(void *a, float v, uint64_t pol) { VGPU_RED2("add", "f32", "f", a, v, pol); }
__device__ __forceinline__ void red_add_f64(void *a, double v, uint64_t pol) { VGPU_RED2("add", "f64", "d", a, v, pol); }
__device__ __forceinline__ void st_u8_hint(uint8_t *a, uint32_t v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.u8 [%0], %1, %2;" ::"l"(a), "r"(v), "l"(pol) : "memory");
}

__device__ __forceinline__ void acc_update(void *addr, uint32_t op, uint64_t v, uint64_t pol) {
  uint32_t *a32 = reinterpret_cast<uint32_t *>(addr);
  uint64_t *a64 = reinterpret_cast<uint64_t *>(addr);
  switch (op) {
    case A_ADD32: red_add_u32(a32, (uint32_t)v, pol); break;
    case A_ADD64: red_add_u64(a64, v, pol); break;
    case A_ADDF32: red_add_f32(a32, __uint_as_float((uint32_t)v), pol); break;
    case A_ADDF64: red_add_f64(a64, __longlong_as_double((long long)v), pol); break;
    case A_MINS32: red_min_s32(a32, (int)(uint32_t)v, pol); break;
    case A_MAXS32: red_max_s32(a32, (int)(uint32_t)v, pol); break;
    case A_MINU32: red_min_u32(a32, (uint32_t)v, pol); break;
    case A_MAXU32: red_max_u32(a32, (uint32_t)v, pol); break;
    case A_MINS64: red_min_s64(a64, (long long)v, pol); break;
    case A_MAXS64: red_max_s64(a64, (long long)v, pol); break;
    case A_MINU64: red_min_u64(a64, v, pol); break;
    case A_MAXU64: red_max_u64(a64, v, pol); break;
    // IEEE order on raw bits: non-negative floats order like signed ints, negative floats in
    // reverse like unsigned ints. (NaN metrics are outside the reference's tested domain.)
    case A_MAXF32: {
      uint32_t b = (uint32_t)v;
      if (!(b >> 31)) red_max_s32(a32, (int)b, pol); else red_min_u32(a32, b, pol);
    } break;
    case A_MINF32: {
      uint32_t b = (uint32_t)v;
      if (!(b >> 31)) red_min_s32(a32, (int)b, pol); else red_max_u32(a32, b, pol);
    } break;
    case A_MAXF64: {
      if (!(v >> 63)) red_max_s64(a64, (long long)v, pol); else red_min_u64(a64, v, pol);
    } break;
    case A_MINF64: {
      if (!(v >> 63)) red_min_s64(a64, (long long)v, pol); else red_max_u64(a64, v, pol);
    } break;
    default: break;
  }
}

// The same Update() on a CTA-private accumulator in shared memory (32-bit shared-window address).
__device__ __forceinline__ void acc_update_shared(uint32_t a, uint32_t op, uint64_t v) {
  const uint32_t v32 = (uint32_t)v;
  switch (op) {
    case A_ADD32: asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v32) : "memory"); break;
    case A_ADD64: asm volatile("red.shared.add.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); break;
    case A_ADDF32: asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(a), "f"(__uint_as_float(v32)) : "memory"); break;
    case A_ADDF64: asm volatile("red.shared.add.f64 [%0], %1;" ::"r"(a), "d"(__longlong_as_double((long long)v)) : "memory"); break;
    case A_MINS32: asm volatile("red.shared.min.s32 [%0], %1;" ::"r"(a), "r"(v32) : "memory"); break;
    case A_MAXS32: asm volatile("red.shared.max.s32 [%0], %1;" ::"r"(a), "r"(v32) : "memory"); break;
    case A_MINU32: asm volatile("red.shared.min.u32 [%0], %1;" ::"r"(a), "r"(v32) : "memory"); break;
    case A_MAXU32: asm volatile("red.shared.max.u32 [%0], %1;" ::"r"(a), "r"(v32) : "memory"); break;
    case A_MINS64: asm volatile("red.shared.min.s64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); break;
    case A_MAXS64: asm volatile("red.shared.max.s64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); break;
    case A_MINU64: asm volatile("red.shared.min.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); break;
    case A_MAXU64: asm volatile("red.shared.max.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); break;
    case A_MAXF32:
      if (!(v32 >> 31)) asm volatile("red.shared.max.s32 [%0], %1;" ::"r"(a), "r"(v32) : "memory");
      else asm volatile("red.shared.min.u32 [%0], %1;" ::"r"(a), "r"(v32) : "memory");
      break;
    case A_MINF32:
      if (!(v32 >> 31)) asm volatile("red.shared.min.s32 [%0], %1;" ::"r"(a), "r"(v32) : "memory");
      else asm volatile("red.shared.max.u32 [%0], %1;" ::"r"(a), "r"(v32) : "memory");
      break;
    case A_MAXF64:
      if (!(v >> 63)) asm volatile("red.shared.max.s64 [%0], %1;" ::"r"(a), "l"(v) : "memory");
      else asm volatile("red.shared.min.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory");
      break;
    case A_MINF64:
      if (!(v >> 63)) asm volatile("red.shared.min.s64 [%0], %1;" ::"r"(a), "l"(v) : "memory");
      else asm volatile("red.shared.max.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory");
      break;
    default: break;
  }
}


// Find-or-claim the cell of `key` in the open-addressing table. The all-ones key (== the EMPTY
// marker) gets the dedicated cell `cap`. Returns ~0 on probe-limit overflow.
__device__ __forceinline__ uint64_t hash_cell(const ScanParams &P, uint64_t key) {
  if (key == kEmptyKey) {
    P.present[0] = 1;
    return P.hmask + 1;
  }
  uint64_t slot = mix64(key) & P.hmask;
  for (uint32_t probe = 0; probe < P.max_probe; ++probe) {
    uint64_t *kp = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(P.hkeys) + slot * P.hkey_stride);
    uint64_t k = *reinterpret_cast<volatile uint64_t *>(kp);
    if (k == key) return slot;
    if (k == kEmptyKey) {
      unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long *>(kp),
                                         (unsigned long long)kEmptyKey, (unsigned long long)key);
      if (old == kEmptyKey || old == key) return slot;
    }
    slot = (slot + 1) & P.hmask;
  }
  return kEmptyKey;
}

// Key tuples that do not pack into 64 bits (e.g. a float next to a double dimension,
// test/aggregation.cc NumericDimensions): one word per key, slots claimed through a state word.
// 0 free -> 1 being written -> 2 ready; readers of a slot in state 1 wait for the writer (Volta+
// independent thread scheduling makes the intra-warp case safe).
__device__ __noinline__ uint64_t wide_cell(const ScanParams &P, const uint64_t *kw) {
  uint64_t h = 0x9e3779b97f4a7c15ull;
  for (uint32_t k = 0; k < P.nkeys; ++k) h = mix64(h ^ kw[k]);
  uint64_t slot = h & P.hmask;
  for (uint32_t probe = 0; probe < P.max_probe; ++probe) {
    while (true) {
      uint32_t st = *reinterpret_cast<volatile uint32_t *>(P.wstate + slot);
      if (st == 0) {
        if (atomicCAS(P.wstate + slot, 0u, 1u) == 0u) {
          for (uint32_t k = 0; k < P.nkeys; ++k) P.wkeys[slot * P.nkeys + k] = kw[k];
          __threadfence();
          atomicExch(P.wstate + slot, 2u);
          return slot;
        }
        continue;  // somebody else claimed it: look again
      }
      if (st == 2) break;
    }
    __threadfence();
    bool same = true;
    for (uint32_t k = 0; k < P.nkeys; ++k)
      same = same && (*reinterpret_cast<volatile uint64_t *>(P.wkeys + slot * P.nkeys + k) == kw[k]);
    if (same) return slot;
    slot = (slot + 1) & P.hmask;
  }
  return kEmptyKey;
}

// ---------------------------------------------------------------------------------------------
// count-distinct, after the scan (util::Bitset |= and cardinality(), src/util/bitset.h:26-67, as a set union of
// (group, id) pairs). The scan leaves the pairs in ragged regions: one per scan CTA (and per owner rank when
// several GPUs take part). Two ways to deduplicate them, both exact:
//
//   fast path       pairs_bucket_kernel scatters the pairs into NB hash buckets small enough for a shared-memory
//                   set, pairs_dedupe_smem_kernel takes one bucket per CTA iteration: open-addressing set in
//                   shared memory, a pair seen for the first time bumps distinct[cell] with one RED.
//                   Measured on B200 (tools/dedupe_probe.cu, 2.5e7 pairs = C2): RED alone 0.15 ms; a global
//                   atomicCAS set 1.24 ms (DRAM-resident) / 0.71 + 0.20 ms (16 L2-sized partitions, round 1);
//                   this path 0.30 + 0.25 ms. Plain-store sets with CTA barriers instead of shared-memory
//                   atomics measured 2.5x SLOWER (0.62 ms) and were dropped.
//   general path    pairs_partition_kernel (L2-sized hash buckets) + pairs_dedupe_kernel<Pair> (global set):
//                   any size or skew, and 16-byte pairs (64-bit ids = Roaring64Map, bitset.h:27-31; packed 64-bit
//                   group keys when hashed group tables are merged across GPUs).
// ---------------------------------------------------------------------------------------------
struct Pair128 {
  uint64_t hi;  // cell, or packed group key
  uint64_t lo;  // id
};
__device__ __forceinline__ Pair128 cas128(Pair128 *p, Pair128 cmp, Pair128 val) {
  Pair128 r;
  asm volatile(
      "{\n .reg .b128 c, v, d;\n mov.b128 c, {%2, %3};\n mov.b128 v, {%4, %5};\n"
      " atom.global.cas.b128 d, [%6], c, v;\n mov.b128 {%0, %1}, d;\n}"
      : "=l"(r.hi), "=l"(r.lo)
      : "l"(cmp.hi), "l"(cmp.lo), "l"(val.hi), "l"(val.lo), "l"(p)
      : "memory");
  return r;
}

constexpr int kBucketThreads = 1024;
constexpr int kBucketPer = 8;          // pairs per thread and tile
constexpr uint32_t kMaxSmemBuckets = 16384;
struct PairsBucketParams {
  const uint64_t *pairs;       // regions, region r at r * region_cap
  const uint32_t *counts;      // [nregions]
  uint32_t nregions;
  uint32_t region_cap;
  uint32_t nbuckets;           // power of two <= kMaxSmemBuckets
  uint32_t bucket_cap;
  uint32_t *cursors;           // [nbuckets], zeroed; ends up holding the bucket sizes (may exceed bucket_cap)
  uint64_t *out;               // bucket b at b * bucket_cap
  unsigned long long *flags;   // set on bucket overflow
};

// Tile of 8192 pairs per CTA iteration: histogram + rank of every pair in shared memory, one global atomic per
// non-empty bucket and tile, then the scatter (the L2 write-combines the 8-byte stores of a bucket's tail).
__global__ void __launch_bounds__(kBucketThreads) pairs_bucket_kernel(const __grid_constant__ PairsBucketParams A) {
  extern __shared__ uint32_t s_hist[];  // [nbuckets] counts, then [nbuckets] bases
  uint32_t *s_base = s_hist + A.nbuckets;
  for (uint32_t r = blockIdx.x; r < A.nregions; r += gridDim.x) {
    const uint64_t *src = A.pairs + (uint64_t)r * A.region_cap;
    const uint32_t n = min(A.counts[r], A.region_cap);
    for (uint32_t t0 = 0; t0 < n; t0 += kBucketThreads * kBucketPer) {
      for (uint32_t i = threadIdx.x; i < A.nbuckets; i += kBucketThreads) s_hist[i] = 0;
      __syncthreads();
      uint64_t key[kBucketPer];
      uint32_t bkt[kBucketPer], pos[kBucketPer];
#pragma unroll
      for (int j = 0; j < kBucketPer; ++j) {
        const uint32_t i = t0 + j * kBucketThreads + threadIdx.x;
        bkt[j] = 0xffffffffu;
        if (i < n) {
          key[j] = src[i];
          bkt[j] = pair_bucket(key[j], A.nbuckets);
          pos[j] = atomicAdd(&s_hist[bkt[j]], 1u);
        }
      }
      __syncthreads();
      for (uint32_t i = threadIdx.x; i < A.nbuckets; i += kBucketThreads)
        if (s_hist[i]) s_base[i] = atomicAdd(&A.cursors[i], s_hist[i]);
      __syncthreads();
      bool over = false;
#pragma unroll
      for (int j = 0; j < kBucketPer; ++j) {
        if (bkt[j] == 0xffffffffu) continue;
        const uint32_t p = s_base[bkt[j]] + pos[j];
        if (p < A.bucket_cap) A.out[(uint64_t)bkt[j] * A.bucket_cap + p] = key[j];
        else over = true;
      }
      if (over) atomicOr(A.flags, 1ull);
      __syncthreads();
    }
  }
}

constexpr int kSmemSetThreads = 1024;
constexpr uint32_t kSmemSetSlots = 12288;   // 96 KB: two CTAs per SM (best of tools/dedupe_probe.cu)
struct PairsSmemDedupeParams {
  const uint64_t *buckets;     // bucket b at b * bucket_cap
  const uint32_t *cursors;     // [nbuckets]
  uint32_t nbuckets;
  uint32_t bucket_cap;
  uint32_t nslots;             // slots of the shared-memory set (8 bytes each)
  uint32_t limit;              // pairs a bucket may hold (<= 3/4 of the slots)
  uint8_t *distinct;           // count of cell 0 (uint32), `stride` bytes between cells
  uint32_t stride;
  unsigned long long *flags;   // set when a bucket does not fit the set
};

__global__ void __launch_bounds__(kSmemSetThreads, 2) pairs_dedupe_smem_kernel(const __grid_constant__ PairsSmemDedupeParams D) {
  extern __shared__ __align__(16) uint64_t s_set[];
  const uint64_t pol = make_table_policy(false);
  for (uint32_t b = blockIdx.x; b < D.nbuckets; b += gridDim.x) {
    const uint32_t n = D.cursors[b];
    if (n == 0) continue;   // uniform per CTA
    if (n > D.limit || n > D.bucket_cap) {
      if (threadIdx.x == 0) atomicOr(D.flags, 1ull);
      continue;
    }
    for (uint32_t i = threadIdx.x; i < D.nslots; i += kSmemSetThreads) s_set[i] = kEmptyKey;
    __syncthreads();
    const uint64_t *src = D.buckets + (uint64_t)b * D.bucket_cap;
    for (uint32_t i = threadIdx.x; i < n; i += kSmemSetThreads) {
      const uint64_t key = src[i];
      // the all-ones pair cannot live in the set: cells are < 2^32 - 1 by construction, so it never occurs
      uint32_t slot = (uint32_t)(((mix64(key ^ 0x5bd1e9955bd1e995ull) >> 32) * D.nslots) >> 32);
      while (true) {
        const unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long *>(s_set + slot),
                                                 (unsigned long long)kEmptyKey, (unsigned long long)key);
        if (old == kEmptyKey) {
          red_add_u32(D.distinct + (key >> 32) * D.stride, 1u, pol);
          break;
        }
        if (old == key) break;
        slot = slot + 1 == D.nslots ? 0 : slot + 1;
      }
    }
    __syncthreads();
  }
}

// ---- general path ----
constexpr int kMaxBuckets = 256;
struct PairsPartitionParams {
  const uint64_t *pairs;       // regions, region r at r * region_cap
  const uint32_t *counts;      // [nregions]; nullptr: `total` contiguous pairs cut into virtual regions
  uint64_t total;
  uint32_t nregions;
  uint32_t region_cap;
  uint32_t nbuckets;           // power of two <= kMaxBuckets
  uint32_t shift;              // bucket = mix64(pair) >> shift
  uint64_t bucket_cap;
  unsigned long long *cursors; // [nbuckets], zeroed
  uint64_t *out;               // bucket b at b * bucket_cap
  unsigned long long *overflow;
};

__global__ void __launch_bounds__(256) pairs_partition_kernel(const __grid_constant__ PairsPartitionParams A) {
  __shared__ uint32_t s_cnt[kMaxBuckets];
  __shared__ unsigned long long s_base[kMaxBuckets];
  constexpr int kPer = 8;
  for (uint32_t r = blockIdx.x; r < A.nregions; r += gridDim.x) {
    const uint64_t *src = A.pairs + (uint64_t)r * A.region_cap;
    const uint32_t n = A.counts ? min(A.counts[r], A.region_cap)
                                : (uint32_t)min((uint64_t)A.region_cap, A.total - (uint64_t)r * A.region_cap);
    for (uint32_t t0 = 0; t0 < n; t0 += blockDim.x * kPer) {
      for (uint32_t i = threadIdx.x; i < A.nbuckets; i += blockDim.x) s_cnt[i] = 0;
      __syncthreads();
      uint64_t key[kPer];
      uint32_t bkt[kPer], pos[kPer];
#pragma unroll
      for (int j = 0; j < kPer; ++j) {
        const uint32_t i = t0 + j * blockDim.x + threadIdx.x;
        bkt[j] = 0xffffffffu;
        if (i < n) {
          key[j] = src[i];
          bkt[j] = (uint32_t)(mix64(key[j]) >> A.shift) & (A.nbuckets - 1);
          pos[j] = atomicAdd(&s_cnt[bkt[j]], 1u);
        }
      }
      __syncthreads();
      for (uint32_t i = threadIdx.x; i < A.nbuckets; i += blockDim.x)
        if (s_cnt[i]) s_base[i] = atomicAdd(&A.cursors[i], (unsigned long long)s_cnt[i]);
      __syncthreads();
#pragma unroll
      for (int j = 0; j < kPer; ++j) {
        if (bkt[j] == 0xffffffffu) continue;
        const unsigned long long p = s_base[bkt[j]] + pos[j];
        if (p < A.bucket_cap) A.out[(uint64_t)bkt[j] * A.bucket_cap + p] = key[j];
        else atomicExch(A.overflow, 1ull);
      }
      __syncthreads();
    }
  }
}

// Pair = uint64_t (cell << 32 | id) or Pair128
template <class Pair>
struct PairsDedupeParams {
  const Pair *pairs;           // one bucket, or ragged regions when counts != nullptr
  const uint32_t *counts;      // nullptr: `n` contiguous pairs
  uint32_t nregions, region_cap;
  uint64_t n;
  Pair *set;                   // all-ones = free slot
  uint64_t set_mask;
  uint8_t *distinct;           // count of cell 0 (uint32), `stride` bytes between cells
  uint32_t stride;
  // Pair128 whose `hi` is a packed group key instead of a cell (merge of hashed group tables across GPUs): the
  // cell is the key's slot in this open-addressing table (the key is always there: its partial record arrived first)
  const uint64_t *lookup_keys;
  uint64_t lookup_mask;
  unsigned long long *sentinel_seen;  // the all-ones Pair128 cannot live in the set: counted once through this flag
};

__device__ __forceinline__ uint64_t lookup_cell(const uint64_t *keys, uint64_t mask, uint64_t key) {
  if (key == kEmptyKey) return mask + 1;   // the all-ones key lives in its dedicated cell
  uint64_t slot = mix64(key) & mask;
  while (true) {
    const uint64_t k = keys[slot];
    if (k == key) return slot;
    if (k == kEmptyKey) return kEmptyKey;  // cannot happen
    slot = (slot + 1) & mask;
  }
}

__device__ __forceinline__ void pairs_insert(const PairsDedupeParams<uint64_t> &D, uint64_t key, uint64_t pol) {
  // the all-ones pair cannot live in the set: cells are < 2^32 - 1 by construction, so it never occurs
  uint64_t slot = mix64(key ^ 0x5bd1e9955bd1e995ull) & D.set_mask;
  while (true) {
    unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long *>(D.set + slot),
                                       (unsigned long long)kEmptyKey, (unsigned long long)key);
    if (old == kEmptyKey) {
      red_add_u32(D.distinct + (key >> 32) * D.stride, 1u, pol);
      return;
    }
    if (old == key) return;
    slot = (slot + 1) & D.set_mask;
  }
}
__device__ __forceinline__ void pairs_insert(const PairsDedupeParams<Pair128> &D, Pair128 key, uint64_t pol) {
  bool fresh = false;
  if (key.hi == kEmptyKey && key.lo == kEmptyKey) {
    fresh = atomicExch(D.sentinel_seen, 1ull) == 0ull;
  } else {
    uint64_t slot = mix64(mix64(key.hi) ^ (key.lo * 0x9E3779B97F4A7C15ull)) & D.set_mask;
    const Pair128 empty{kEmptyKey, kEmptyKey};
    while (true) {
      const Pair128 old = cas128(D.set + slot, empty, key);
      if (old.hi == kEmptyKey && old.lo == kEmptyKey) { fresh = true; break; }
      if (old.hi == key.hi && old.lo == key.lo) break;
      slot = (slot + 1) & D.set_mask;
    }
  }
  if (!fresh) return;
  const uint64_t cell = D.lookup_keys ? lookup_cell(D.lookup_keys, D.lookup_mask, key.hi) : key.hi;
  if (cell != kEmptyKey) red_add_u32(D.distinct + cell * D.stride, 1u, pol);
}

template <class Pair>
__global__ void __launch_bounds__(256) pairs_dedupe_kernel(const __grid_constant__ PairsDedupeParams<Pair> D) {
  const uint64_t pol = make_table_policy(false);
  if (D.counts == nullptr) {
    // one insert in flight per thread: the kernel is bound by the L2's atomic throughput, not by latency
    // (four CAS in flight per thread measured 5 % slower on C2)
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < D.n; i += (uint64_t)gridDim.x * blockDim.x)
      pairs_insert(D, D.pairs[i], pol);
  } else {
    for (uint32_t r = blockIdx.x; r < D.nregions; r += gridDim.x) {
      const Pair *src = D.pairs + (uint64_t)r * D.region_cap;
      const uint32_t n = min(D.counts[r], D.region_cap);
      for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) pairs_insert(D, src[i], pol);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// multi-GPU exchange: scatter the live entries of an open-addressing u64 table (and the accumulator
// cells that go with them) into one bucket per owner rank. owner = mix64(key >> owner_shift) % nparts
// (count-distinct pairs: owner of the CELL, so that all ids of a group meet on one rank; group
// records: owner of the packed key). Bucket o occupies [o * bucket_cap, (o+1) * bucket_cap).
// CTA-level histogram in shared memory: one global atomic per owner per tile.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxParts = 16;
constexpr int kPartPerThread = 8;
struct PartitionParams {
  const uint64_t *keys;   // table keys, kEmptyKey = free slot
  uint64_t nslots;
  uint64_t sentinel_slot; // slot index that stands for the all-ones key when sentinel_present points at 1 (hash group tables), else ~0
  const uint8_t *sentinel_present;
  uint32_t nparts;
  uint32_t owner_shift;
  uint64_t bucket_cap;
  unsigned long long *cursors;  // [nparts], zeroed; ends up holding the bucket sizes
  uint64_t *out_keys;
  uint32_t npay;
  uint32_t pay_width[kMaxMetrics + 1];
  const void *pay_src[kMaxMetrics + 1];  // indexed by slot
  void *pay_dst[kMaxMetrics + 1];        // indexed by bucket position
};

__device__ __forceinline__ uint32_t owner_of(uint64_t key, uint32_t shift, uint32_t nparts) {
  return (uint32_t)((mix64(key >> shift) >> 17) % nparts);
}

__global__ void __launch_bounds__(256) partition_table_kernel(const __grid_constant__ PartitionParams A) {
  __shared__ uint32_t s_cnt[kMaxParts];
  __shared__ unsigned long long s_base[kMaxParts];
  const uint64_t tile = (uint64_t)blockDim.x * kPartPerThread;
  const uint64_t total = A.nslots + (A.sentinel_slot != ~0ull ? 1 : 0);
  const uint64_t ntiles = (total + tile - 1) / tile;
  for (uint64_t ti = blockIdx.x; ti < ntiles; ti += gridDim.x) {
    if (threadIdx.x < kMaxParts) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    uint64_t key[kPartPerThread];
    uint64_t slot[kPartPerThread];
    uint32_t own[kPartPerThread], pos[kPartPerThread];
#pragma unroll
    for (int j = 0; j < kPartPerThread; ++j) {
      const uint64_t i = ti * tile + (uint64_t)j * blockDim.x + threadIdx.x;
      own[j] = 0xffffffffu;
      if (i < A.nslots) {
        key[j] = A.keys[i];
        slot[j] = i;
        if (key[j] != kEmptyKey) own[j] = owner_of(key[j], A.owner_shift, A.nparts);
      } else if (i == A.nslots && A.sentinel_slot != ~0ull && A.sentinel_present[0]) {
        key[j] = kEmptyKey;  // the all-ones key lives in its dedicated cell
        slot[j] = A.sentinel_slot;
        own[j] = owner_of(kEmptyKey, A.owner_shift, A.nparts);
      }
      if (own[j] != 0xffffffffu) pos[j] = atomicAdd(&s_cnt[own[j]], 1u);
    }
    __syncthreads();
    if (threadIdx.x < A.nparts && s_cnt[threadIdx.x])
      s_base[threadIdx.x] = atomicAdd(&A.cursors[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kPartPerThread; ++j) {
      if (own[j] == 0xffffffffu) continue;
      const uint64_t p = (uint64_t)own[j] * A.bucket_cap + s_base[own[j]] + pos[j];
      A.out_keys[p] = key[j];
      for (uint32_t m = 0; m < A.npay; ++m) {
        if (A.pay_width[m] == 4)
          reinterpret_cast<uint32_t *>(A.pay_dst[m])[p] = reinterpret_cast<const uint32_t *>(A.pay_src[m])[slot[j]];
        else
          reinterpret_cast<uint64_t *>(A.pay_dst[m])[p] = reinterpret_cast<const uint64_t *>(A.pay_src[m])[slot[j]];
      }
    }
    __syncthreads();
  }
}

// Merge exchanged group records (packed key + one partial accumulator per metric) into a hash group
// table with the same commutative Update() the scan uses (store.cc:131-161).
struct MergeParams {
  const uint64_t *keys;
  uint64_t n;
  uint32_t nmets;
  uint32_t ops[kMaxMetrics + 1];
  uint32_t widths[kMaxMetrics + 1];
  const void *src[kMaxMetrics + 1];
  void *acc[kMaxMetrics + 1];
  uint64_t *hkeys;
  uint64_t hmask;
  uint8_t *present;
  uint32_t max_probe;
  unsigned long long *overflow;
};

__global__ void __launch_bounds__(256) merge_records_kernel(const __grid_constant__ MergeParams M) {
  const uint64_t pol = make_table_policy(false);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M.n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t key = M.keys[i];
    uint64_t cell;
    if (key == kEmptyKey) {
      M.present[0] = 1;
      cell = M.hmask + 1;
    } else {
      uint64_t slot = mix64(key) & M.hmask;
      cell = kEmptyKey;
      for (uint32_t probe = 0; probe < M.max_probe; ++probe) {
        unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long *>(M.hkeys + slot),
                                           (unsigned long long)kEmptyKey, (unsigned long long)key);
        if (old == kEmptyKey || old == key) { cell = slot; break; }
        slot = (slot + 1) & M.hmask;
      }
      if (cell == kEmptyKey) { atomicExch(M.overflow, 1ull); continue; }
    }
    for (uint32_t m = 0; m < M.nmets; ++m) {
      uint64_t v = M.widths[m] == 4 ? (uint64_t)reinterpret_cast<const uint32_t *>(M.src[m])[i]
                                    : reinterpret_cast<const uint64_t *>(M.src[m])[i];
      if (M.ops[m] == A_ADD32 || M.ops[m] == A_MINS32 || M.ops[m] == A_MAXS32)
        v = (uint64_t)(int64_t)(int32_t)(uint32_t)v;
      acc_update(reinterpret_cast<uint8_t *>(M.acc[m]) + cell * M.widths[m], M.ops[m], v, pol);
    }
  }
}

// Wide key tuples across ranks: compact the live slots of a wide table (state 2) into records (nkeys key words + one
// partial accumulator per metric), and merge records into a wide table with the same Update().
struct WideCompactParams {
  const uint32_t *wstate;
  const uint64_t *wkeys;
  uint64_t nslots;
  uint32_t nkeys, nmets;
  uint64_t cap;                      // records the outputs hold
  unsigned long long *cursor;
  uint64_t *out_keys;                // [cap * nkeys]
  uint32_t widths[kMaxMetrics + 1];
  const void *src[kMaxMetrics + 1];  // indexed by slot
  void *dst[kMaxMetrics + 1];        // indexed by record
};
__global__ void __launch_bounds__(256) wide_compact_kernel(const __grid_constant__ WideCompactParams W) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < W.nslots; i += (uint64_t)gridDim.x * blockDim.x) {
    if (W.wstate[i] != 2u) continue;
    const unsigned long long p = atomicAdd(W.cursor, 1ull);
    if (p >= W.cap) continue;
    for (uint32_t k = 0; k < W.nkeys; ++k) W.out_keys[p * W.nkeys + k] = W.wkeys[i * W.nkeys + k];
    for (uint32_t m = 0; m < W.nmets; ++m) {
      if (W.widths[m] == 4) reinterpret_cast<uint32_t *>(W.dst[m])[p] = reinterpret_cast<const uint32_t *>(W.src[m])[i];
      else reinterpret_cast<uint64_t *>(W.dst[m])[p] = reinterpret_cast<const uint64_t *>(W.src[m])[i];
    }
  }
}
struct WideMergeParams {
  const uint64_t *keys;              // [n * nkeys]
  uint64_t n;
  uint32_t nmets;
  uint32_t ops[kMaxMetrics + 1];
  uint32_t widths[kMaxMetrics + 1];
  const void *src[kMaxMetrics + 1];
  void *acc[kMaxMetrics + 1];
  unsigned long long *overflow;
};
// T carries the destination table (wstate, wkeys, hmask, nkeys, max_probe): wide_cell() finds or claims the slot
__global__ void __launch_bounds__(256) wide_merge_kernel(const __grid_constant__ ScanParams T, const __grid_constant__ WideMergeParams M) {
  const uint64_t pol = make_table_policy(false);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M.n; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t kw[kMaxKeys];
    for (uint32_t k = 0; k < T.nkeys; ++k) kw[k] = M.keys[i * T.nkeys + k];
    const uint64_t cell = wide_cell(T, kw);
    if (cell == kEmptyKey) { atomicExch(M.overflow, 1ull); continue; }
    for (uint32_t m = 0; m < M.nmets; ++m) {
      uint64_t v = M.widths[m] == 4 ? (uint64_t)reinterpret_cast<const uint32_t *>(M.src[m])[i]
                                    : reinterpret_cast<const uint64_t *>(M.src[m])[i];
      if (M.ops[m] == A_ADD32 || M.ops[m] == A_MINS32 || M.ops[m] == A_MAXS32) v = (uint64_t)(int64_t)(int32_t)(uint32_t)v;
      acc_update(reinterpret_cast<uint8_t *>(M.acc[m]) + cell * M.widths[m], M.ops[m], v, pol);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// group extraction: present cells -> dense SoA result
// ---------------------------------------------------------------------------------------------
struct ExtractKey {
  uint64_t lo, div, mod;  // value = lo + (packed / div) % mod   (mod == 0: no modulo)
  uint64_t lut;           // non-zero: the digit is the rank of the value among the set bits (KeySpec::lut)
  const uint64_t *dict;   // non-null: the digit is the rank of a rolled-up time value (TimeDict), dict[rank] = value
  uint32_t width;
  uint32_t pad;
  void *out;
};
struct ExtractMet {
  const void *acc;     // accumulator of cell 0
  uint32_t stride;     // bytes between cells
  uint32_t acc_width;  // 4 or 8
  uint32_t out_width;  // 1,2,4,8 (truncation == the reference's own-type wrap-around, Q4)
  uint32_t sext;       // signed integer type: sign-extend the truncated value (HAVING / sort key on the device)
  uint32_t pad;
  void *out;
};
constexpr int kMaxPostProg = 24;
struct ExtractParams {
  uint64_t ncells;  // dense cells, or hash capacity + 1 (last = sentinel-key cell)
  uint32_t hash_mode;
  uint32_t nkeys, nmets;
  const uint64_t *hkeys;
  const uint8_t *present;
  uint32_t hkey_stride, present_stride;
  uint32_t present_width;   // dense: bytes of the presence word (1: flag; 4 / 8: a COUNT accumulator, ScanParams::skip_present)
  const uint32_t *wstate;
  const uint64_t *wkeys;
  ExtractKey keys[kMaxKeys];
  ExtractMet mets[kMaxMetrics + 1];
  unsigned long long *counter;  // number of groups
  uint32_t count_only;
  uint64_t cap;                 // rows the output arrays hold: groups beyond it are counted, not written
  uint32_t *pos_out;            // optional [ncells], preset to 0xffffffff: output position of every present cell
                                // (extract_late_kernel fills in the accumulators that were not final yet)
  // ---- post-aggregation on the device (vgpu_plan: HAVING, top-N) ----
  uint32_t nhprog;              // HAVING program: PInstr with cls == C_GEN, slot = source (key k, or nkeys + metric m)
  PInstr hprog[kMaxPostProg];
  uint32_t sort_src;            // source of the first sort column, 0xffffffff: none
  uint32_t sort_kind;           // 0: unsigned (numeric order == SmallerInt order), 1: signed (SmallerInt string order)
  uint32_t sort_desc;
  uint64_t *cand_cell;          // candidates (present and passing HAVING): cell and order-preserving sort key
  uint64_t *cand_ord;
  unsigned long long *cand_n;   // number of candidates
  unsigned long long *out_n;    // rows written by post_extract_kernel
  const unsigned long long *threshold;  // post_extract_kernel: keep candidates with ord <= *threshold
};

__device__ __forceinline__ bool cell_present(const ExtractParams &E, uint64_t c, uint64_t &packed) {
  if (!E.hash_mode) {
    packed = c;
    const uint8_t *p = E.present + c * E.present_stride;
    if (E.present_width == 4) return *reinterpret_cast<const uint32_t *>(p) != 0u;
    if (E.present_width == 8) return *reinterpret_cast<const uint64_t *>(p) != 0ull;
    return *p != 0;
  }
  if (E.hash_mode == 2) {
    packed = c;
    return c + 1 < E.ncells && E.wstate[c] == 2u;
  }
  if (c == E.ncells - 1) {
    packed = kEmptyKey;
    return E.present[0] != 0;
  }
  packed = *reinterpret_cast<const uint64_t *>(reinterpret_cast<const uint8_t *>(E.hkeys) + c * E.hkey_stride);
  return packed != kEmptyKey;
}

// value of key k of the group in cell c, widened (two's complement for signed types, raw bits for floats)
__device__ __forceinline__ uint64_t group_key_value(const ExtractParams &E, uint32_t k, uint64_t c, uint64_t packed) {
  const ExtractKey &ek = E.keys[k];
  if (E.hash_mode == 2) return E.wkeys[c * E.nkeys + k];
  uint64_t q = packed / ek.div;
  if (ek.mod) q %= ek.mod;
  uint64_t v = ek.lo + q;
  if (ek.dict) v = ek.dict[q];
  if (ek.lut) {  // the q-th set bit
    uint64_t x = ek.lut;
    for (uint64_t i = 0; i < q; ++i) x &= x - 1;
    v = (uint64_t)(__ffsll((long long)x) - 1);
  }
  return v;
}
__device__ __forceinline__ uint64_t group_acc_raw(const ExtractMet &em, uint64_t c) {
  const uint8_t *ap = reinterpret_cast<const uint8_t *>(em.acc) + c * em.stride;
  return em.acc_width == 4 ? (uint64_t)*reinterpret_cast<const uint32_t *>(ap) : *reinterpret_cast<const uint64_t *>(ap);
}
// accumulator m as the reference holds it: truncated to the column's own width (Q4), sign-extended for signed types
__device__ __forceinline__ uint64_t group_acc_value(const ExtractMet &em, uint64_t c) {
  uint64_t v = group_acc_raw(em, c);
  if (em.out_width < 8) {
    v &= (1ull << (8 * em.out_width)) - 1;
    if (em.sext) { const uint64_t sb = 1ull << (8 * em.out_width - 1); v = (v ^ sb) - sb; }
  }
  return v;
}
// write group `c` to row `pos` of the result arrays
__device__ __forceinline__ void extract_cell(const ExtractParams &E, uint64_t c, uint64_t packed, uint64_t pos) {
  for (uint32_t k = 0; k < E.nkeys; ++k) {
    const ExtractKey &ek = E.keys[k];
    const uint64_t v = group_key_value(E, k, c, packed);
    switch (ek.width) {
      case 1: reinterpret_cast<uint8_t *>(ek.out)[pos] = (uint8_t)v; break;
      case 2: reinterpret_cast<uint16_t *>(ek.out)[pos] = (uint16_t)v; break;
      case 4: reinterpret_cast<uint32_t *>(ek.out)[pos] = (uint32_t)v; break;
      default: reinterpret_cast<uint64_t *>(ek.out)[pos] = v; break;
    }
  }
  for (uint32_t m = 0; m < E.nmets; ++m) {
    const ExtractMet &em = E.mets[m];
    const uint64_t v = group_acc_raw(em, c);
    switch (em.out_width) {
      case 1: reinterpret_cast<uint8_t *>(em.out)[pos] = (uint8_t)v; break;
      case 2: reinterpret_cast<uint16_t *>(em.out)[pos] = (uint16_t)v; break;
      case 4: reinterpret_cast<uint32_t *>(em.out)[pos] = (uint32_t)v; break;
      default: reinterpret_cast<uint64_t *>(em.out)[pos] = v; break;
    }
  }
}

__global__ void __launch_bounds__(256)
extract_groups_kernel(const __grid_constant__ ExtractParams E) {
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t limit = (E.ncells + 31) & ~31ull;  // whole warps iterate together
  for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < limit; c += stride) {
    uint64_t packed = 0;
    bool pres = (c < E.ncells) && cell_present(E, c, packed);
    uint32_t ballot = __ballot_sync(0xffffffffu, pres);
    if (ballot == 0) continue;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(E.counter, (unsigned long long)__popc(ballot));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (!pres || E.count_only) continue;
    uint64_t pos = base + __popc(ballot & ((1u << lane) - 1));
    if (pos >= E.cap) continue;
    if (E.pos_out) E.pos_out[c] = (uint32_t)pos;
    extract_cell(E, c, packed, pos);
  }
}

// ---------------------------------------------------------------------------------------------
// post-aggregation on the device: HAVING on raw accumulators, top-N selection on the first sort key
// ---------------------------------------------------------------------------------------------
// value of source `src` (key k, or nkeys + metric m) of the group in cell c
__device__ __forceinline__ uint64_t group_value(const ExtractParams &E, uint32_t src, uint64_t c, uint64_t packed) {
  return src < E.nkeys ? group_key_value(E, src, c, packed) : group_acc_value(E.mets[src - E.nkeys], c);
}
// FilterComparison over (agg key, accumulators), post_agg.cc:76-83: the same bitwise &,| tree as the row predicate
__device__ __forceinline__ bool group_passes(const ExtractParams &E, uint64_t c, uint64_t packed) {
  uint32_t stk = 0;  // bit i = stack entry i (depth <= kStackDepth)
  int sp = 0;
  for (uint32_t pc = 0; pc < E.nhprog; ++pc) {
    const PInstr &in = E.hprog[pc];
    if (in.kind <= P_OR_LEAF) {
      bool m = in.cls == C_TRUE ? true : in.cls == C_FALSE ? false
               : post_compare(in.gcls, in.gop, group_value(E, in.slot, c, packed), in.arg);
      if (in.neg) m = !m;
      if (in.kind == P_PUSH) { stk = (stk & ~(1u << sp)) | ((m ? 1u : 0u) << sp); ++sp; }
      else if (in.kind == P_AND_LEAF) { if (!m) stk &= ~(1u << (sp - 1)); }
      else { if (m) stk |= 1u << (sp - 1); }
    } else {
      const bool b = (stk >> (sp - 1)) & 1u, a = (stk >> (sp - 2)) & 1u;
      const bool r = in.kind == P_AND ? (a && b) : (a || b);
      --sp;
      stk = (stk & ~(1u << (sp - 1))) | ((r ? 1u : 0u) << (sp - 1));
    }
  }
  return sp == 0 ? true : (stk & 1u) != 0;
}
__device__ __forceinline__ uint64_t sort_ordinal(const ExtractParams &E, uint64_t c, uint64_t packed) {
  const uint64_t v = group_value(E, E.sort_src, c, packed);
  const uint64_t o = E.sort_kind == 1 ? smaller_int_rank((long long)v) : v;
  return E.sort_desc ? ~o : o;
}

// pass 1: count the groups, keep the candidates (present and passing HAVING) with their sort key
__global__ void __launch_bounds__(256) post_select_kernel(const __grid_constant__ ExtractParams E) {
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t limit = (E.ncells + 31) & ~31ull;
  for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < limit; c += stride) {
    uint64_t packed = 0;
    const bool pres = (c < E.ncells) && cell_present(E, c, packed);
    const uint32_t ballot = __ballot_sync(0xffffffffu, pres);
    if (ballot == 0) continue;
    if (lane == 0) atomicAdd(E.counter, (unsigned long long)__popc(ballot));
    const bool cand = pres && group_passes(E, c, packed);
    const uint32_t cb = __ballot_sync(0xffffffffu, cand);
    if (cb == 0) continue;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(E.cand_n, (unsigned long long)__popc(cb));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (!cand) continue;
    const uint64_t pos = base + __popc(cb & ((1u << lane) - 1));
    E.cand_cell[pos] = c;
    E.cand_ord[pos] = E.sort_src != 0xffffffffu ? sort_ordinal(E, c, packed) : 0ull;
  }
}

// radix select of the k-th smallest sort key among the candidates, most significant byte first: one histogram pass over
// the keys that match the prefix found so far, one single-thread step that picks the byte. state: [0] prefix,
// [1] remaining k (1-based), [2] resolved bits mask; after 8 rounds state[0] is the threshold.
__global__ void __launch_bounds__(256) radix_hist_kernel(const uint64_t *ord, const unsigned long long *n_ptr, const unsigned long long *state,
                                                         uint32_t shift, unsigned int *hist) {
  __shared__ unsigned int s_h[256];
  s_h[threadIdx.x] = 0;
  __syncthreads();
  if (state[3]) return;   // nothing to select: every candidate is kept
  const uint64_t n = *n_ptr, prefix = state[0], mask = state[2];
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t o = ord[i];
    if ((o & mask) == prefix) atomicAdd(&s_h[(o >> shift) & 0xffu], 1u);
  }
  __syncthreads();
  if (s_h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], s_h[threadIdx.x]);
}
__global__ void radix_pick_kernel(unsigned long long *state, int shift, unsigned int *hist, const unsigned long long *n_ptr,
                                  unsigned long long top_k) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (shift < 0) {                         // set-up call
    state[0] = 0; state[2] = 0;
    state[1] = top_k;
    state[3] = (top_k == 0 || *n_ptr <= top_k) ? 1 : 0;   // fewer candidates than wanted: keep them all
    if (state[3]) state[0] = ~0ull;
    return;
  }
  if (state[3]) { for (int b = 0; b < 256; ++b) hist[b] = 0; return; }
  unsigned long long k = state[1];
  int b = 0;
  for (; b < 255; ++b) {
    if (hist[b] >= k) break;
    k -= hist[b];
  }
  state[1] = k;
  state[0] |= (unsigned long long)b << shift;
  state[2] |= 0xffull << shift;
  for (int i = 0; i < 256; ++i) hist[i] = 0;
}

// pass 2: the candidates whose sort key is <= the threshold become the result
__global__ void __launch_bounds__(256) post_extract_kernel(const __grid_constant__ ExtractParams E) {
  const unsigned long long n = *E.cand_n, thr = *E.threshold;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    if (E.cand_ord[i] > thr) continue;
    const uint64_t c = E.cand_cell[i];
    uint64_t packed = 0;
    cell_present(E, c, packed);
    const unsigned long long pos = atomicAdd(E.out_n, 1ull);
    if (pos < E.cap) extract_cell(E, c, packed, pos);
  }
}

// presence flags from a COUNT accumulator (several GPUs: the flags are what the ranks max-reduce)
__global__ void __launch_bounds__(256) derive_present_kernel(uint8_t *present, const uint8_t *acc, uint32_t width, uint64_t ncells) {
  for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < ncells; c += (uint64_t)gridDim.x * blockDim.x)
    present[c] = (width == 4 ? *reinterpret_cast<const uint32_t *>(acc + c * 4) != 0u : *reinterpret_cast<const uint64_t *>(acc + c * 8) != 0ull) ? 1 : 0;
}

// accumulators that become final after the groups were extracted (count-distinct: the dedupe runs while the
// keys and the other accumulators already travel to the host)
struct ExtractLateParams {
  uint64_t ncells;
  const uint32_t *pos;   // ExtractParams::pos_out
  uint32_t nmets;
  ExtractMet mets[kMaxDistinct];
};
__global__ void __launch_bounds__(256) extract_late_kernel(const __grid_constant__ ExtractLateParams L) {
  for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < L.ncells; c += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t pos = L.pos[c];
    if (pos == 0xffffffffu) continue;
    for (uint32_t m = 0; m < L.nmets; ++m) {
      const ExtractMet &em = L.mets[m];
      const uint8_t *ap = reinterpret_cast<const uint8_t *>(em.acc) + c * em.stride;
      const uint64_t v = em.acc_width == 4 ? (uint64_t)*reinterpret_cast<const uint32_t *>(ap) : *reinterpret_cast<const uint64_t *>(ap);
      if (em.out_width == 8) reinterpret_cast<uint64_t *>(em.out)[pos] = v;
      else reinterpret_cast<uint32_t *>(em.out)[pos] = (uint32_t)v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// helpers: fills, per-column min/max (segment stats), synthetic generator
// ---------------------------------------------------------------------------------------------
// Interleaved group cells: every cell starts as the same pattern of `words` 32-bit words.
struct CellPattern {
  uint32_t words;
  uint32_t w[48];
};
__global__ void __launch_bounds__(256) fill_cells_kernel(uint32_t *p, uint64_t total_words, const __grid_constant__ CellPattern C) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_words;
       i += (uint64_t)gridDim.x * blockDim.x)
    p[i] = C.w[i % C.words];
}
__global__ void fill32_kernel(uint32_t *p, uint64_t n, uint32_t v) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x)
    p[i] = v;
}
__global__ void fill64_kernel(uint64_t *p, uint64_t n, uint64_t v) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x)
    p[i] = v;
}


struct StatCol {
  uint64_t off;
  uint32_t width, sext, type, pad;
};
struct StatParams {
  const uint8_t *slab;
  uint64_t nrows;
  uint32_t ncols;
  StatCol cols[32];
  unsigned long long *out;  // [ncols][2] ordered min, ordered max
};

__global__ void __launch_bounds__(256) column_minmax_kernel(const __grid_constant__ StatParams S) {
  const uint32_t c = blockIdx.y;
  const StatCol &sc = S.cols[c];
  uint64_t mn = ~0ull, mx = 0;
  for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < S.nrows;
       r += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t o = to_ordered(load_elem(S.slab + sc.off + r * sc.width, sc.width, sc.sext), sc.type);
    mn = min(mn, o);
    mx = max(mx, o);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = min(mn, __shfl_down_sync(0xffffffffu, mn, o));
    mx = max(mx, __shfl_down_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0 && mn <= mx) {
    atomicMin(&S.out[2 * sc.pad], (unsigned long long)mn);   // sc.pad = schema column (output slot)
    atomicMax(&S.out[2 * sc.pad + 1], (unsigned long long)mx);
  }
}

// ---------------------------------------------------------------------------------------------
// row-major mirror (put time): columns -> rows. A CTA transposes tiles of kRowsTileBytes / stride rows
// through shared memory: column cells are read coalesced along the rows, the finished rows leave as
// one contiguous coalesced block. One pass over the segment: reads row_bytes, writes stride per row.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kRowsTileBytes = 32768;
struct RowsCol {
  const uint8_t *src;  // cell of row 0
  uint32_t width;
  uint32_t row_off;
};
struct RowsParams {
  uint8_t *rows;
  uint64_t nrows;
  uint32_t stride;
  uint32_t ncols;
  RowsCol cols[32];
};

__global__ void __launch_bounds__(256) build_rows_kernel(const __grid_constant__ RowsParams R) {
  __shared__ __align__(16) uint8_t s_tile[kRowsTileBytes];
  const uint32_t tile_rows = kRowsTileBytes / R.stride;
  const uint64_t ntiles = (R.nrows + tile_rows - 1) / tile_rows;
  for (uint64_t ti = blockIdx.x; ti < ntiles; ti += gridDim.x) {
    const uint64_t row0 = ti * tile_rows;
    const uint32_t n = (uint32_t)min((uint64_t)tile_rows, R.nrows - row0);
    // padding bytes of the rows are written too: zero them once
    for (uint32_t i = threadIdx.x; i < n * R.stride / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(s_tile)[i] = 0;
    __syncthreads();
    for (uint32_t c = 0; c < R.ncols; ++c) {
      const RowsCol &rc = R.cols[c];
      for (uint32_t r = threadIdx.x; r < n; r += blockDim.x) {
        const uint8_t *src = rc.src + (row0 + r) * rc.width;
        uint8_t *dst = s_tile + r * R.stride + rc.row_off;
        switch (rc.width) {
          case 1: *dst = *src; break;
          case 2: *reinterpret_cast<uint16_t *>(dst) = *reinterpret_cast<const uint16_t *>(src); break;
          case 4: *reinterpret_cast<uint32_t *>(dst) = *reinterpret_cast<const uint32_t *>(src); break;
          default: *reinterpret_cast<uint64_t *>(dst) = *reinterpret_cast<const uint64_t *>(src); break;
        }
      }
    }
    __syncthreads();
    uint32_t *out = reinterpret_cast<uint32_t *>(R.rows + row0 * R.stride);  // stride is a multiple of 4
    for (uint32_t i = threadIdx.x; i < n * R.stride / 4; i += blockDim.x) out[i] = reinterpret_cast<uint32_t *>(s_tile)[i];
    __syncthreads();
  }
}


struct GenCol {
  uint64_t off;     // slab offset, or unused for bitset
  uint32_t width;
  uint32_t mode;
  int64_t lo;
  uint64_t range;
  uint64_t div;
  uint32_t *bitset_out;  // non-null: write uint32 ids here instead of the slab
  uint32_t gen_index;    // column index used in the hash
  uint32_t is_f32, is_f64, pad;
};
struct GenParams {
  uint8_t *slab;
  uint64_t nrows, seed, row_offset;
  uint32_t ncols;
  GenCol cols[32];
};

__global__ void __launch_bounds__(256) generate_kernel(const __grid_constant__ GenParams G) {
  for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < G.nrows;
       r += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t row = G.row_offset + r;
    for (uint32_t c = 0; c < G.ncols; ++c) {
      const GenCol &gc = G.cols[c];
      uint64_t u = gc.mode ? (row / gc.div) : splitmix64(G.seed * 0x100000001B3ULL + row * 16 + gc.gen_index);
      int64_t v = gc.lo + (int64_t)(u % gc.range);
      if (gc.bitset_out) { gc.bitset_out[r] = (uint32_t)v; continue; }
      uint8_t *p = G.slab + gc.off + r * gc.width;
      if (gc.is_f32) { *reinterpret_cast<float *>(p) = (float)v; continue; }
      if (gc.is_f64) { *reinterpret_cast<double *>(p) = (double)v; continue; }
      switch (gc.width) {
        case 1: *p = (uint8_t)v; break;
        case 2: *reinterpret_cast<uint16_t *>(p) = (uint16_t)v; break;
        case 4: *reinterpret_cast<uint32_t *>(p) = (uint32_t)v; break;
        default: *reinterpret_cast<uint64_t *>(p) = (uint64_t)v; break;
      }
    }
  }
}

}  // namespace vgpu

#endif  // VGPU_KERNELS_CUH_
