// scan_params.h — the device-side "plan": everything the fused scan kernel needs, as one POD that
// is passed by value (__grid_constant__) so every field is a uniform constant-bank read.
// Built on the host by planner code in vgpu.cu from a vgpu_plan (include/vgpu.h).
#ifndef VGPU_SCAN_PARAMS_H_
#define VGPU_SCAN_PARAMS_H_

#include <stdint.h>

namespace vgpu {

constexpr int kThreads = 256;       // threads per CTA
constexpr int kVec = 4;             // consecutive rows owned by a thread inside a sub-tile
constexpr int kSub = 4;             // sub-tiles per tile
constexpr int kRowsPerThread = kVec * kSub;
constexpr int kSubRows = kThreads * kVec;       // 1024
constexpr int kTileRows = kSubRows * kSub;      // 4096 rows per CTA iteration

constexpr int kMaxSlots = 32;   // distinct columns one query may touch
constexpr int kMaxProg = 64;    // predicate program length (leaves + combinators)
constexpr int kMaxKeys = 12;
constexpr int kMaxMetrics = 16;
constexpr int kMaxRules = 8;
constexpr int kStackDepth = 5;
constexpr int kMaxDistinct = 2;
constexpr int kMaxBitsetCols = 4;
constexpr int kMaxRanks = 8;        // GPUs of one box: owner sub-regions of a CTA's count-distinct pair region
constexpr int kMaxTimeSegs = 48;    // pieces of a rolled-up time key's bucket dictionary

// Per-query counter block (16 x u64). With several GPUs words [kCSumFirst, kCMaxFirst) are sum-reduced and words
// [kCMaxFirst, kCLocalFirst) max-reduced across ranks in one exchange, so every flag has a word of its own.
enum Counter : int {
  kCPassed = 0,        // rows whose predicate was true
  kCScannedRecs = 1,   // host-provided: QueryStats.scanned_recs of this rank
  kCScannedSegs = 2,   // host-provided: QueryStats.scanned_segments of this rank
  kCPairs = 3,         // +d: count-distinct pairs appended for distinct metric d (all CTAs)
  kCMaxFill = 5,       // +d: fullest pair region
  kCHashOver = 7,      // probe limit of the group table reached: grow x4, scan again
  kCRegionOver = 8,    // a pair region overflowed: grow, scan again
  kCBucketOver = 9,    // count-distinct fast path: a hash bucket overflowed
  kCSetOver = 10,      // count-distinct fast path: a bucket does not fit the shared-memory set
  kCError = 11,        // a rank failed before the exchange: everybody aborts
  kCGroups = 12,       // groups extracted (local)
  kCUnit = 13,         // next work unit of the scan (low 32 bits)
  kCOut = 14,          // device post-aggregation: rows written to the result
  kCCand = 15,         // device post-aggregation: groups that passed HAVING
  kCSumFirst = 0, kCMaxFirst = 5, kCLocalFirst = 12
};

// A column as the kernel sees it.
struct Slot {
  uint64_t off;      // bytes per row of all preceding fixed-width columns: the column starts at
                     // slab + off * SegDesc::cap (fixed-width columns)
  uint32_t width;    // 1,2,4,8
  uint32_t sext;     // 1: sign-extend narrow values (i8/i16/i32) when widening to 64 bit
  uint32_t bitset;   // 1: BITSET column (values/offsets come from per-segment side tables)
  uint32_t bitset_idx;  // which bitset column of the table (index into SegDesc::bs_*)
  uint32_t row_off;  // byte offset of the cell inside a row of the row-major mirror (naturally aligned)
  uint32_t pad;
  uint64_t vmask;    // value mask of the element (0xff, 0xffff, 0xffffffff, ~0)
  uint64_t signbit;  // sign bit of the element when sext, else 0:  (x ^ signbit) - signbit sign-extends
};

enum PKind : uint8_t { P_PUSH = 0, P_AND_LEAF = 1, P_OR_LEAF = 2, P_AND = 3, P_OR = 4 };
enum PCls : uint8_t {
  C_TRUE = 0,   // constant true
  C_FALSE = 1,  // constant false
  C_EQ32 = 2,   // zero-extended raw value == arg (1/2/4-byte columns)
  C_LT32 = 3,   // (zero-extended raw value ^ bias) <u arg (1/2/4-byte columns; bias maps signed types
                // to the order-preserving unsigned domain)
  C_RNG32 = 4,  // ((raw ^ bias) - arg) <u arg2  (fused lo <= x < hi, peephole)
  C_GEN = 5,    // generic per-row path: 8-byte ints, float, double, bitset cardinality
  C_LUT64 = 6   // bit `raw` of the 64-bit mask in arg (IN / NOT IN lists whose values are all < 64)
};
// generic compare element classes
enum GCls : uint8_t { G_I64 = 0, G_U64 = 1, G_F32 = 2, G_F64 = 3, G_CARD = 4 };

struct PInstr {
  uint8_t kind;   // PKind
  uint8_t cls;    // PCls
  uint8_t slot;
  uint8_t neg;    // invert the leaf mask
  uint8_t gcls;   // GCls for C_GEN
  uint8_t gop;    // vgpu_relop for C_GEN
  uint16_t pad;
  uint32_t bias;  // C_LT32/C_RNG32
  uint32_t arg2;  // C_RNG32: hi - lo
  uint64_t arg;   // prepared argument (raw bits)
};

struct KeySpec {
  uint8_t slot;
  uint8_t rollup;      // 1: apply time rollup (scan.cc:198-219)
  uint8_t micro;       // 1: value is microseconds (util::Time64), else seconds (util::Time32)
  uint8_t nrules;
  uint8_t rule_unit[kMaxRules];  // vgpu_time_unit
  uint8_t query_unit;            // vgpu_time_unit or VGPU_TU_NONE
  uint8_t fzero;                 // floating-point key: -0.0 groups with +0.0 (KeyEqual uses ==, store.cc:46-63)
  uint8_t pad[2];
  uint64_t rule_boundary[kMaxRules];
  uint64_t lo;    // subtracted from the (rolled-up) value
  uint64_t mul;   // cell index / packed key = sum (v - lo) * mul
  // copy of the key column's Slot fields: direct constant operands under the unrolled key loop
  uint64_t col_off, vmask, signbit;
  uint32_t width, row_off;
  uint64_t lut;   // non-zero: the predicate restricts this key to the codes whose bits are set (an IN list over
                  // codes < 64 in a top-level conjunction); the key is numbered by its rank in the set instead
                  // of v - lo, which shrinks the dense key domain to the values that can actually occur
};

enum AccOp : uint8_t {
  A_ADD32 = 0, A_ADD64 = 1, A_ADDF32 = 2, A_ADDF64 = 3,
  A_MINS32 = 4, A_MAXS32 = 5, A_MINU32 = 6, A_MAXU32 = 7,
  A_MINS64 = 8, A_MAXS64 = 9, A_MINU64 = 10, A_MAXU64 = 11,
  A_MINF32 = 12, A_MAXF32 = 13, A_MINF64 = 14, A_MAXF64 = 15,
  A_DISTINCT = 16
};

struct MetSpec {
  uint8_t slot;
  uint8_t op;      // AccOp
  uint8_t pad[2];
  uint32_t stride; // bytes between the accumulators of consecutive cells (== width when the table is
                   // one array per metric, == cell size when the fields of a cell are interleaved)
  void *acc;       // accumulator of cell 0
  // copy of the metric column's Slot fields (one constant load level instead of two)
  uint64_t col_off, vmask, signbit;
  uint32_t width, row_off, bitset, bitset_idx;
  uint32_t soff;       // offset of the accumulator inside a cell of the CTA-private table (smem_cells != 0)
  uint32_t acc_width;  // 4 or 8
  uint32_t id64;       // BITSET column whose ids are 64 bits wide (util::Bitset<8> = Roaring64Map, bitset.h:27-31):
  uint32_t pad2;       // the segment's id array holds uint64 and the cell always goes through its CSR offsets
};

// Bucket dictionary of a rolled-up time key. Rule boundaries cut the raw time line into regions with
// one truncation unit each (rollup.cc:77-95); inside a region the truncated values are equally spaced (minute /
// hour / day) or one per calendar month / year. When the truncated value never decreases along the raw time line,
// its RANK among all attainable values is a step function of the raw value: piece j covers raw values from
// start[j] on, and inside it the rank grows by one every step[j] seconds counted from origin[j] <= start[j] (the
// unit-aligned start of its first bucket; step 0: the whole piece is one value). The key becomes that rank — a dense domain of a few hundred buckets instead of 2^32 seconds — and the
// calendar arithmetic runs once per bucket on the host (KeySpec rollup stays the fallback).
struct TimeDict {
  uint32_t npieces;      // 0: not in use
  uint32_t key;          // index of the key it applies to
  uint32_t micro;        // values are microseconds: pieces are in seconds, value / 1e6 first
  uint32_t pad;
  uint64_t start[kMaxTimeSegs];   // ascending (seconds); start[0] <= every value of the column
  uint32_t start32[kMaxTimeSegs]; // the same when `narrow` (every value fits 32 bits): half the compare work per probe
  uint32_t narrow;
  uint32_t pad2;
  uint64_t origin[kMaxTimeSegs];
  uint32_t base[kMaxTimeSegs];
  uint32_t step[kMaxTimeSegs];    // 0, 60, 3600, 86400 (or 1 for second granularity)
};

// Per-segment descriptor (device array, one per table segment).
struct SegDesc {
  const uint8_t *slab;        // fixed-width columns
  uint64_t nrows;
  uint64_t cap;               // row capacity of the slab (multiple of kTileRows)
  const uint32_t *bs_values[kMaxBitsetCols];   // per bitset column: ids (uint32)
  const uint32_t *bs_offsets[kMaxBitsetCols];  // per bitset column: CSR offsets (nrows+1) or nullptr = 1 id/row
  const uint8_t *rows;        // row-major mirror: row r at rows + r * ScanParams::row_stride, or nullptr
};

struct ScanParams {
  // work list
  const SegDesc *segs;
  const uint32_t *active;   // indices of segments to scan
  uint32_t nactive;
  uint32_t tiles_per_seg;   // 512-row warp chunks per segment
  uint64_t total_tiles;     // nactive * tiles_per_seg
  uint32_t unit_chunks;     // chunks per work unit (a run of consecutive chunks of one segment)
  uint32_t units_per_seg;   // ceil(tiles_per_seg / unit_chunks); unit u = (active segment u / units_per_seg, part u % units_per_seg)

  // columns
  Slot slots[kMaxSlots];
  uint32_t nslots;

  uint32_t small_plan;  // <= 4 keys of <= 4 bytes and <= 4 metrics: register-staged gather path
  uint32_t tune;  // bit 1: stream the columns with an L2 evict_first policy; bit 2: group table evict_last

  // Row-major mirror: at low selectivity every key / metric cell of a passing row costs
  // a whole 64-byte DRAM atom in its column, but all of them share one or two atoms of the mirror row.
  // A batch of passing rows is gathered from the mirror when the chunk that filled it had at most
  // row_thresh passing rows (0: never) — the break-even of the two byte counts, computed by the planner.
  uint32_t row_stride;
  uint32_t row_thresh;
  uint32_t conj;              // predicate is a conjunction of at most 4 vectorisable leaves (unrolled fast path)

  // fixed-width columns the predicate reads with vector loads (prefetched one chunk ahead)
  uint32_t nfilter_slots;
  uint8_t filter_slots[kMaxSlots];
  uint8_t pf_width[kMaxSlots];   // per predicate column f: element width and bytes-per-row offset of the column
  uint64_t pf_off[kMaxSlots];    // in the slab (bulk L2 prefetch of the next chunk, lane f takes column f)

  // predicate
  uint32_t nprog;
  PInstr prog[kMaxProg];

  // group keys
  uint32_t nkeys;
  uint32_t hash_mode;      // 0: dense cells, 1: open-addressing hash on the packed 64-bit key,
                           // 2: open-addressing hash on the full key tuple (one 64-bit word per key)
  KeySpec keys[kMaxKeys];
  uint64_t *hkeys;         // hash_mode: key of slot 0, EMPTY = ~0
  uint32_t hkey_stride;    // bytes between the keys of consecutive slots
  uint32_t present_stride; // bytes between the presence flags of consecutive cells (dense)
  uint64_t hmask;          // capacity - 1
  uint8_t *present;        // dense: 1 byte per cell; hash: present[0] flags the sentinel key
  uint32_t skip_present;   // dense: no presence store per row — a COUNT accumulator whose cells are all >= 1 (segment
                           // statistics) is non-zero exactly for the groups that exist: one L2 operation less per row
  uint32_t *wstate;        // wide mode: 0 free, 1 being written, 2 ready
  uint64_t *wkeys;         // wide mode: capacity x nkeys words
  uint32_t max_probe;

  // CTA-private copy of a small dense group table in shared memory (smem_cells != 0): every CTA aggregates
  // into its own copy with shared-memory atomics and merges it into the global table once, at the end —
  // a few groups would otherwise serialise all RED operations of the GPU on a few L2 addresses
  uint32_t smem_cells, smem_stride, smem_present_off;
  uint32_t smem_init[16];   // initial image of a cell (smem_stride / 4 words)

  // metrics
  uint32_t nmetrics;
  MetSpec mets[kMaxMetrics];
  uint32_t ndistinct;                 // BITSET metrics selected (count-distinct)
  uint8_t distinct_met[kMaxDistinct]; // their indices into mets
  // count-distinct: (cell << 32 | id) pairs are appended to private regions per CTA (cursors in shared memory)
  // and deduplicated after the scan (pairs_* kernels). With several GPUs a CTA keeps one sub-region per owner
  // rank (owner = hash of the pair), laid out owner-major so that everything bound for one rank is one slab.
  // dpair_wide: 16-byte pairs {cell or packed group key, 64-bit id} (64-bit ids; hashed group tables across GPUs)
  uint64_t *dpairs[kMaxDistinct];       // region (sub, b): [(sub * gridDim.x + b) * dpair_cap, ...) elements
  uint32_t *dpair_count[kMaxDistinct];  // [dpair_nsub * gridDim.x] pairs written per region
  uint32_t dpair_cap;
  uint32_t dpair_nsub;                  // 1, or the number of ranks
  uint32_t dpair_wide;
  uint32_t dpair_key;                   // wide pairs carry the packed group key instead of the local cell (hash_mode 1)

  TimeDict tdict;

  // per-query counter block, layout kC* below
  unsigned long long *counters;
};

}  // namespace vgpu

#endif  // VGPU_SCAN_PARAMS_H_
