/*
 * vgpu.h — C ABI of libvgpu.so: the B200-native (sm_100a) replacement for ViyaDB's
 * JIT-compiled columnar scan -> filter -> hash group-by-aggregate hot loop.
 *
 * This is the drop-in boundary. Everything above it (query JSON parsing, query::AggregateQuery,
 * FilterArgsPacker, post-aggregation formatting/sorting) stays in the host language of the
 * reference (C++); everything below it is hand-written CUDA. Plain C types only: no C++ types,
 * no exceptions, no torch types. All entry points return VGPU_OK (0) or a negative vgpu_status
 * and record a message retrievable with vgpu_last_error().
 *
 * What each entry point replaces in the reference (paths relative to the viyadb/viyadb tree):
 *
 *   vgpu_table_create / vgpu_segment_put / vgpu_table_invalidate
 *       the physical column store the generated scan loop reads: the JIT-emitted
 *       `class Segment : db::SegmentBase { Dimensions d; Metrics m; SegmentStats stats; }`
 *       (src/codegen/db/store.cc:203-356), db::SegmentStore::segments_copy()
 *       (src/db/store.h:40-44) and SegmentBase::size() (src/db/segment.h:34-39).
 *       Host column memory is never owned by the library: put() copies into HBM.
 *
 *   vgpu_query_agg
 *       the generated `extern "C" viya_query_agg(db::Table&, query::RowOutput&, query::QueryStats&,
 *       std::vector<db::AnyNum> fargs, size_t skip, size_t limit, std::vector<db::AnyNum> hargs)`
 *       (src/codegen/query/agg_query.cc:26-75; fn-pointer type query::AggQueryFn,
 *       src/query/runner.h:33-35; call site src/query/runner.cc:61-62) up to and including
 *       `stats.aggregated_recs = agg_map.size()` (src/codegen/query/scan.cc:168-247), i.e.
 *         - segment loop + segment skip       scan.cc:40-51, filter.cc:263-349
 *         - row predicate                      scan.cc:56-65, filter.cc:206-261
 *         - group key build + time rollup      scan.cc:193-224, codegen/db/rollup.cc:25-95, util/time.h
 *         - agg_map[key].Update(metrics)       scan.cc:228-242, codegen/db/store.cc:31-169
 *         - count-distinct (util::Bitset |=)   util/bitset.h:26-67 over CRoaring
 *       The plan is DATA (predicate program + raw AnyNum argument bytes exactly as the reference's
 *       own FilterArgsPacker produces them, src/codegen/query/filter.cc:100-124) — nothing is
 *       compiled per query.
 *
 *   vgpu_result_*
 *       the `agg_map` the generated post-aggregation code iterates
 *       (src/codegen/query/post_agg.cc:26-147) and the four QueryStats counters
 *       (src/query/stats.h:35-58). HAVING / dict decode / formatting / sort / skip / limit remain
 *       host code.
 *
 *   vgpu_comm_* / vgpu_query_agg_partial / vgpu_partial_*
 *       multi-GPU: segments shard across the GPUs of one box, one process per GPU; the per-GPU
 *       partial group tables are merged with one NCCL exchange. The reference's only counterpart
 *       is the cluster controller's HTTP fan-out + TSV re-aggregation
 *       (src/cluster/query/agg_runner.cc:93-140), whose *semantics* (merge = the same
 *       commutative Update) are kept, not its transport.
 */
#ifndef VGPU_H_
#define VGPU_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VGPU_ABI_VERSION 3

typedef enum vgpu_status {
  VGPU_OK = 0,
  VGPU_ERR_INVALID = -1,     /* bad argument / malformed plan or schema            */
  VGPU_ERR_UNSUPPORTED = -2, /* outside the reference's own domain */
  VGPU_ERR_CUDA = -3,        /* CUDA runtime / driver failure                      */
  VGPU_ERR_NOMEM = -4,       /* device or host allocation failed                   */
  VGPU_ERR_NCCL = -5,        /* NCCL failure                                       */
  VGPU_ERR_STATE = -6        /* call order / missing segment / shut down           */
} vgpu_status;

/* Element types: db::NumericType::Type (src/db/column.h:67-78) + the UIntType code widths. */
typedef enum vgpu_type {
  VGPU_U8 = 0, VGPU_U16 = 1, VGPU_U32 = 2, VGPU_U64 = 3,
  VGPU_I8 = 4, VGPU_I16 = 5, VGPU_I32 = 6, VGPU_I64 = 7,
  VGPU_F32 = 8, VGPU_F64 = 9
} vgpu_type;

/* Column kinds: db::Dimension::DimType (column.h:153) and the two Metric classes. */
typedef enum vgpu_col_kind {
  VGPU_DIM_STRING = 0,    /* dict code, width from cardinality (column.cc:54-62)      */
  VGPU_DIM_NUMERIC = 1,
  VGPU_DIM_TIME = 2,      /* uint32 seconds                                           */
  VGPU_DIM_MICROTIME = 3, /* uint64 microseconds                                      */
  VGPU_DIM_BOOLEAN = 4,
  VGPU_METRIC_VALUE = 5,
  VGPU_METRIC_BITSET = 6,
  VGPU_METRIC_HIDDEN_COUNT = 7 /* the `uint64_t _count[]` of store.cc:286-289 */
} vgpu_col_kind;

/* Same order as db::Metric::AggregationType (column.h:246). */
typedef enum vgpu_agg {
  VGPU_AGG_MAX = 0, VGPU_AGG_MIN = 1, VGPU_AGG_SUM = 2, VGPU_AGG_AVG = 3,
  VGPU_AGG_COUNT = 4, VGPU_AGG_BITSET = 5, VGPU_AGG_NONE = 255
} vgpu_agg;

/* Same order as util::TimeUnit (src/util/time.h:31). WEEK has no Truncator specialisation in
 * the reference (time.h:52-89) and is rejected with VGPU_ERR_UNSUPPORTED. */
typedef enum vgpu_time_unit {
  VGPU_TU_YEAR = 0, VGPU_TU_MONTH = 1, VGPU_TU_WEEK = 2, VGPU_TU_DAY = 3,
  VGPU_TU_HOUR = 4, VGPU_TU_MINUTE = 5, VGPU_TU_SECOND = 6, VGPU_TU_NONE = 7
} vgpu_time_unit;

typedef struct vgpu_column {
  uint32_t kind; /* vgpu_col_kind */
  uint32_t type; /* vgpu_type; for BITSET the width the ids travel in: U32 (the reference's ubyte / ushort / uint
                    bitsets, util::Bitset<4> = Roaring) or U64 (ulong bitsets, util::Bitset<8> = Roaring64Map,
                    src/util/bitset.h:27-31) */
  uint32_t agg;  /* vgpu_agg for metrics, VGPU_AGG_NONE for dimensions */
  uint32_t lit_type; /* 0, or 1 + the vgpu_type of a filter literal on this column when it differs from `type`: a
                    BITSET metric of the reference's `ubyte` / `ushort` type travels as U32, but the AnyNum image of
                    its filter literal only defines 1 / 2 bytes (src/db/column.h:98-121) */
} vgpu_column;

/* Columns are listed as the reference indexes them: all dimensions in Table::dimensions() order,
 * then all metrics in Table::metrics() order, then (optionally) the hidden count column. */
typedef struct vgpu_schema {
  uint32_t ncols;
  uint32_t ndims;
  uint64_t segment_size; /* Table::segment_size(), src/db/table.cc:48 */
  const vgpu_column *cols;
} vgpu_schema;

/* A BITSET metric column of one segment, flattened to CSR: row r holds
 * values[offsets[r] .. offsets[r+1]). offsets == NULL means exactly one id per row.
 * values are uint32_t or uint64_t according to the column's `type`. */
typedef struct vgpu_bitset_csr {
  const uint64_t *offsets;
  const void *values;
  uint64_t nvalues;
} vgpu_bitset_csr;

/* ---- predicate program ------------------------------------------------------------------
 * Post-order encoding of the reference's filter tree AFTER FilterFactory has pushed NOT down and
 * sorted composite children by precedence (src/query/filter.cc:36-108). Leaves consume `args`
 * in exactly the order FilterArgsPacker emits them (filter.cc:100-124). */
typedef enum vgpu_node_kind {
  VGPU_NODE_RELOP = 0, /* col OP args[arg]                      (query::RelOpFilter)           */
  VGPU_NODE_IN = 1,    /* col IN / NOT IN args[arg .. arg+n)    (query::InFilter; op: 1 = IN, 0 = NOT IN) */
  VGPU_NODE_AND = 2,   /* AND of the previous n sub-expressions (query::CompositeFilter)       */
  VGPU_NODE_OR = 3,
  VGPU_NODE_EMPTY = 4  /* constant true                         (query::EmptyFilter)           */
} vgpu_node_kind;

/* Same order as query::RelOpFilter::Operator (src/query/filter.h:40-47). */
typedef enum vgpu_relop {
  VGPU_OP_EQ = 0, VGPU_OP_NE = 1, VGPU_OP_LT = 2, VGPU_OP_LE = 3, VGPU_OP_GT = 4, VGPU_OP_GE = 5
} vgpu_relop;

typedef struct vgpu_pred_node {
  uint32_t kind; /* vgpu_node_kind */
  uint32_t op;   /* vgpu_relop for RELOP; 1/0 for IN/NOT IN */
  uint32_t col;  /* schema column index */
  uint32_t arg;  /* first index into args */
  uint32_t n;    /* IN: number of values; AND/OR: number of children */
  uint32_t reserved;
} vgpu_pred_node;

#define VGPU_MAX_ROLLUP_RULES 8

/* One group-by key = one query::DimOutputColumn (src/query/query.h:119-137). */
typedef struct vgpu_key {
  uint32_t col;
  /* Time rollup, mirroring scan.cc:198-219: active when the TIME dimension has rollup rules or the
   * query column has a granularity. rule_boundary[i] is `now - after_i` in the column's unit
   * (seconds, or microseconds for microtime), computed on the host with util::Duration::add_to
   * (src/util/time.cc:49-83, codegen/db/rollup.cc:44-75); rules in TimeDimension order
   * (descending `after`, src/db/column.cc:346-349). First rule with value < boundary truncates. */
  uint32_t nrules;
  uint32_t query_granularity; /* vgpu_time_unit, VGPU_TU_NONE if absent */
  uint32_t reserved;
  uint64_t rule_boundary[VGPU_MAX_ROLLUP_RULES];
  uint32_t rule_granularity[VGPU_MAX_ROLLUP_RULES];
} vgpu_key;

typedef struct vgpu_plan {
  uint32_t nnodes;
  uint32_t nargs;
  const vgpu_pred_node *nodes;
  const uint64_t *args; /* raw 8-byte db::AnyNum images (src/db/column.h:98-121) */
  uint32_t nkeys;
  uint32_t nmetrics;
  const vgpu_key *keys;
  const uint32_t *metric_cols; /* schema column indices of the selected metrics, query order */
  uint32_t need_hidden_count;  /* AVG selected without a COUNT metric selected (scan.cc:239-241) */
  uint32_t flags;              /* VGPU_PLAN_* */
  /* ---- post-aggregation on the device (optional; replaces the HAVING test of post_agg.cc:76-83 and narrows what
   * sort.cc:24-73 has to sort). Both only REMOVE groups from the result the host then post-aggregates as before:
   *   HAVING   the same post-order node encoding as the row predicate, `col` = schema column of a SELECTED dimension or
   *            metric, literals = the raw AnyNum images FilterArgsPacker emits for query->having(). A metric compares
   *            its accumulator in the column's own type (AVG: the raw sum, Q6; BITSET: the cardinality, Q7), a
   *            dimension its (rolled-up) key value. The caller must only ask for it when the reference applies HAVING to
   *            every group: with a sort, or without skip / limit (without a sort the reference cuts the skip / limit
   *            window out of the map iteration BEFORE it tests HAVING).
   *   top-N    sort_col = schema column of the FIRST sort column, top_k = skip + limit: the result keeps every group
   *            whose first sort key is among the top_k best (ties included), so the host's exact multi-column sort on
   *            formatted strings (Q12) sees a superset of what it will output. Applied when the column's sort order can
   *            be computed from raw values — integer-typed dimensions and metrics (util::StringNumCmp::SmallerInt:
   *            length, then lexicographic, src/util/string.h:28-49), not AVG — otherwise ignored.
   * vgpu_result_view.aggregated_recs stays the number of ALL groups (QueryStats); post_applied says what ran. */
  uint32_t nhnodes;
  uint32_t nhargs;
  const vgpu_pred_node *hnodes;
  const uint64_t *hargs;
  uint32_t sort_col;           /* VGPU_NO_COLUMN: none */
  uint32_t sort_descending;
  uint64_t top_k;              /* 0: no top-N */
} vgpu_plan;

#define VGPU_NO_COLUMN 0xffffffffu
#define VGPU_PLAN_POST 8u      /* the post-aggregation fields above are filled in (older callers leave the bit clear) */

#define VGPU_PLAN_FORCE_HASH 1u  /* testing: never pick the dense group table */
#define VGPU_PLAN_FORCE_DENSE 2u /* testing: fail instead of falling back to hashing */
/* Several GPUs: only rank 0 copies the merged groups to its host (ngroups = 0 elsewhere; aggregated_recs and the other
 * QueryStats counters are still the merged ones on every rank). Without it every rank returns the full result, and G
 * simultaneous device-to-host copies of the same groups share the box's host memory bandwidth (measured at 8 GPUs on
 * 1e7 groups: 212 MB per rank, 19 ms instead of 4). */
#define VGPU_PLAN_RESULT_ON_ROOT 4u

typedef struct vgpu_ctx vgpu_ctx;
typedef struct vgpu_table vgpu_table;
typedef struct vgpu_result vgpu_result;

/* The group table of one query (library-owned host memory, valid until vgpu_result_free). */
typedef struct vgpu_result_view {
  uint64_t ngroups;
  uint32_t nkeys;
  uint32_t nmetrics;
  const void *const *keys;      /* keys[k]: array[ngroups] of the key column's element type */
  const void *const *accs;      /* accs[m]: array[ngroups]; SUM/AVG/COUNT/MIN/MAX in the metric
                                   column's own type (wrap-around like the reference, Q4);
                                   BITSET: cardinality as uint64_t                              */
  const uint64_t *hidden_count; /* array[ngroups] or NULL */
  uint64_t scanned_recs;        /* QueryStats, scan.cc:44 — all segments, pruned or not */
  uint64_t scanned_segments;    /* scan.cc:51 — processed segments only                 */
  uint64_t aggregated_recs;     /* scan.cc:246 == ngroups                               */
  uint64_t passed_rows;         /* rows whose predicate was true (for selectivity/B_alg) */
  double gpu_ms;                /* device time of all kernels of this query (CUDA events) */
  double scan_ms;               /* device time of the fused scan kernel alone             */
  uint32_t launches;            /* kernels launched for this query                        */
  uint32_t table_mode;          /* 0 dense, 1 hash64, 2 wide key tuples */
  uint64_t table_cells;         /* dense cells or hash capacity */
  uint32_t attempts;            /* scans run: > 1 after a hash-table or pair-region overflow (grow and scan again) */
  uint32_t distinct_paths;      /* count-distinct dedupe paths taken, VGPU_DEDUPE_* bits */
  uint32_t post_applied;        /* bit 0: HAVING ran on the device, bit 1: top-N selection ran on the device */
  uint32_t reserved;
} vgpu_result_view;

#define VGPU_DEDUPE_SMALL 1u    /* one global set */
#define VGPU_DEDUPE_FAST 2u     /* hash buckets + shared-memory sets */
#define VGPU_DEDUPE_WIDE 4u     /* 16-byte pairs (64-bit ids / packed group keys), one global set */
#define VGPU_DEDUPE_GENERAL 8u  /* L2-sized partitions + global sets */
#define VGPU_DEDUPE_REDONE 16u  /* the fast path overflowed and the general one ran instead */
#define VGPU_DEDUPE_PARTITIONED 32u /* the general path really cut the pairs into more than one partition */

/* ---- lifecycle ----
 * Threading (the reference runs `query_threads` queries at once next to one ingest thread, src/db/database.cc:28-33):
 * every vgpu_query_* call is re-entrant per context — each runs on streams, events and scratch of its own — and
 * holds its table's lock shared; vgpu_segment_put / _generate / _invalidate hold it exclusively, so a writer waits
 * for the queries running on that table (and they for it) while other tables are not affected. With a communicator
 * (vgpu_comm_init) queries of one context are serialised: every rank must issue the same sequence of collectives. */
int vgpu_abi_version(void);
int vgpu_init(int device, vgpu_ctx **out);
/* Order queries after what `cuda_stream` (cudaStream_t as void*) holds when they start, and make it wait for them
 * when they end: the caller's own events on that stream then bracket a query. NULL detaches. */
int vgpu_set_stream(vgpu_ctx *ctx, void *cuda_stream);
/* Test hooks: force the rarely taken branches at test sizes. name = "pairs_cap" (first capacity of the count-distinct
 * pair regions: overflow + regrow), "hash_cap" (first capacity of hashed group tables: x4 regrow), "bucket_pairs"
 * (pairs per L2-sized partition of the general dedupe path), "small_pairs" (largest capacity one global set takes),
 * "set_slots" (slots of the shared-memory sets), "expect_pairs" (pairs the fast dedupe path sizes its buckets for: too
 * few forces its overflow + the fallback), "tune" (VGPU_TUNE bits). 0 restores the default. */
int vgpu_set_test_hook(vgpu_ctx *ctx, const char *name, uint64_t value);
void vgpu_shutdown(vgpu_ctx *ctx);
const char *vgpu_last_error(void);

/* ---- column store ---- */
int vgpu_table_create(vgpu_ctx *ctx, const vgpu_schema *schema, vgpu_table **out);
void vgpu_table_free(vgpu_table *table);
/* Copy rows [0,nrows) of every column of segment seg_idx into HBM. col_ptrs[c] is the host base
 * address of the column array (any alignment; may be pageable or pinned), or a vgpu_bitset_csr*
 * for BITSET columns. Replaces any previous content of that segment. The host buffers may be reused
 * as soon as the call returns (the copies run on a dedicated stream and are waited for; the per-column
 * statistics and the row-major mirror of the segment are built asynchronously behind them and are
 * ordered before any later query). HBM footprint per row: the column widths, plus the same again
 * (rounded up to 4 or 8 bytes) for the mirror unless VGPU_TUNE bit 6 is set. */
int vgpu_segment_put(vgpu_table *table, uint32_t seg_idx, uint64_t nrows,
                     const void *const *col_ptrs);
/* The same without waiting for the copies: the call returns once they are enqueued, so that the DMA engine never idles
 * between the segments of a table load (measured: 44.6 -> 5x GB/s over PCIe 5 x16). The host buffers must stay valid and
 * unchanged until vgpu_table_sync() returns — or until a query on the table has returned, queries being ordered after
 * every earlier put on the device. Pinned sources (vgpu_host_pin, cudaHostAlloc) copy at full speed; pageable ones are
 * staged by the driver and block the call. */
int vgpu_segment_put_async(vgpu_table *table, uint32_t seg_idx, uint64_t nrows,
                           const void *const *col_ptrs);
int vgpu_table_sync(vgpu_table *table);
/* Incremental sync (the live store grows and its upserts update metric cells in place, src/codegen/db/upsert.cc:386-411):
 * replace rows [row_begin, row_begin + nrows) of an uploaded segment — col_ptrs[c] addresses the cell of row `row_begin` —
 * and/or append behind its last row (row_begin <= rows so far). Only the range crosses PCIe; the per-column statistics
 * are reduced again on the device and the row-major mirror is rebuilt for the range only. VGPU_ERR_STATE when the range
 * does not fit the segment's device buffers, VGPU_ERR_UNSUPPORTED for bitset columns whose cells do not hold exactly one
 * 32-bit id: the caller then puts the whole segment. */
int vgpu_segment_update(vgpu_table *table, uint32_t seg_idx, uint64_t row_begin, uint64_t nrows,
                        const void *const *col_ptrs);
/* Page-lock a host range for DMA (cudaHostRegister) / release it: the reference's segments are single heap objects
 * (`new Segment`, src/codegen/db/store.cc:203-356) that the adapter pins once, when it first uploads them. */
int vgpu_host_pin(vgpu_ctx *ctx, const void *ptr, size_t bytes);
int vgpu_host_unpin(vgpu_ctx *ctx, const void *ptr);
int vgpu_table_invalidate(vgpu_table *table, uint32_t seg_idx);
uint32_t vgpu_table_segments(const vgpu_table *table);
uint64_t vgpu_table_rows(const vgpu_table *table);
uint64_t vgpu_table_bytes(const vgpu_table *table);

/* Synthetic segment written directly in HBM (bench / parity at sizes no host buffer should
 * carry): value(row, c) = lo + (mode ? (row / div) : splitmix64(seed*0x100000001B3 + row*16 + c)) % range,
 * with row = row_offset + r. One entry per schema column; BITSET columns get one id per row. */
typedef struct vgpu_gen_col {
  int64_t lo;
  uint64_t range;
  uint32_t mode; /* 0 = hash, 1 = div */
  uint32_t reserved;
  uint64_t div;
} vgpu_gen_col;
int vgpu_segment_generate(vgpu_table *table, uint32_t seg_idx, uint64_t nrows,
                          const vgpu_gen_col *gens, uint64_t seed, uint64_t row_offset);
/* Copy a device-resident column back (tests): out must hold nrows * elem bytes. */
int vgpu_segment_read(vgpu_table *table, uint32_t seg_idx, uint32_t col, void *out);

/* ---- the hot path ---- */
int vgpu_query_agg(vgpu_table *table, const vgpu_plan *plan, vgpu_result **out);
int vgpu_result_get(const vgpu_result *res, vgpu_result_view *view);
void vgpu_result_free(vgpu_result *res);

/* ---- select and search queries (same segment pruning and row predicate as the aggregate query) ----
 * vgpu_query_select replaces the generated `viya_query_select` (src/codegen/query/select_query.cc:25-54,
 * fn type query::SelectQueryFn, src/query/runner.h:30-32, called at runner.cc:29-43) up to the formatting of
 * the cells: it returns, in the reference's output order, the raw cells of the rows the reference would send
 * — every passing row in (segment, tuple) order, the first `skip` of them dropped, and `limit` applied the
 * way scan.cc:161 does (the tuple loop breaks, the segment loop goes on: once `limit` rows are out, every
 * further processed segment still contributes its first passing row). */
typedef struct vgpu_rows_plan {
  uint32_t nnodes;
  uint32_t nargs;
  const vgpu_pred_node *nodes;
  const uint64_t *args;
  uint32_t ncols;
  uint32_t reserved;
  const uint32_t *cols; /* schema columns to materialise (dimensions, metrics, the hidden count), any order */
  uint64_t skip;        /* SelectQuery::skip()  */
  uint64_t limit;       /* SelectQuery::limit(), 0 = none */
} vgpu_rows_plan;

typedef struct vgpu_rows vgpu_rows;
typedef struct vgpu_rows_view {
  uint64_t nrows;
  uint32_t ncols;
  uint32_t launches;
  const void *const *cells; /* cells[c]: array[nrows] of the column's element type; BITSET: cardinality, uint64_t */
  uint64_t scanned_recs;
  uint64_t scanned_segments;
  uint64_t passed_rows;     /* passing rows of all processed segments (the reference does not count them) */
  double gpu_ms;
} vgpu_rows_view;
int vgpu_query_select(vgpu_table *table, const vgpu_rows_plan *plan, vgpu_rows **out);
int vgpu_rows_get(const vgpu_rows *rows, vgpu_rows_view *view);
void vgpu_rows_free(vgpu_rows *rows);

/* vgpu_query_search replaces the scan of the generated `viya_query_search` (src/codegen/query/scan.cc:249-299):
 * for every processed segment, in scan order, the distinct values of one integer-typed dimension among the
 * passing rows with the first row that holds each, ascending by that row. The caller replays the reference's
 * sequential part on these short lists (codes.insert, substring match on the formatted value, `limit` breaking
 * the tuple loop only) — string work that needs the host-side dictionary. */
typedef struct vgpu_search_plan {
  uint32_t nnodes;
  uint32_t nargs;
  const vgpu_pred_node *nodes;
  const uint64_t *args;
  uint32_t col; /* the dimension (schema column index); floating-point dimensions are not supported */
  uint32_t reserved;
} vgpu_search_plan;

typedef struct vgpu_search vgpu_search;
typedef struct vgpu_search_view {
  uint32_t nsegments;          /* processed segments, scan order */
  uint32_t launches;
  const uint64_t *seg_offsets; /* [nsegments + 1] into codes / first_row */
  const uint64_t *codes;       /* raw value, widened to 64 bits (sign-extended for signed types) */
  const uint32_t *first_row;   /* ascending inside a segment */
  uint64_t scanned_recs;
  uint64_t scanned_segments;
  double gpu_ms;
} vgpu_search_view;
int vgpu_query_search(vgpu_table *table, const vgpu_search_plan *plan, vgpu_search **out);
int vgpu_search_get(const vgpu_search *res, vgpu_search_view *view);
void vgpu_search_free(vgpu_search *res);

/* ---- multi-GPU (one process per GPU) ----
 * unique_id is the 128-byte ncclUniqueId produced by rank 0 (vgpu_comm_unique_id) and distributed
 * by the host's own plumbing (torch.distributed, MPI, a file...). After vgpu_comm_init,
 * vgpu_query_agg merges the per-GPU partial group tables over NCCL and every rank returns the
 * full result. */
int vgpu_comm_unique_id(void *unique_id_128);
int vgpu_comm_init(vgpu_ctx *ctx, int rank, int nranks, const void *unique_id_128);
int vgpu_comm_destroy(vgpu_ctx *ctx);

#ifdef __cplusplus
}
#endif

#endif /* VGPU_H_ */
