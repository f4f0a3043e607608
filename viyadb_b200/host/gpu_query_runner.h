// viyadb_b200/host/gpu_query_runner.h — the drop-in plug on the reference side (C++17).
//
// GpuQueryRunner is a query::QueryVisitor (src/query/query.h:229-235), a sibling of the stock
// query::QueryRunner (src/query/runner.h:40-57). Database::Query (src/db/database.cc:104-111) does
//     query::QueryRunner query_runner(*this, output);  q->Accept(query_runner);
// and a maintainer switches the aggregate path to the GPU by constructing this class instead. No
// reference file is edited; this header only *includes* reference headers and libvgpu's C ABI.
//
//   Visit(AggregateQuery*)  replaces QueryRunner::Visit(AggregateQuery*) (src/query/runner.cc:45-64):
//     1. packs filter / having literals with the reference's OWN cg::FilterArgsPacker
//        (src/codegen/query/filter.cc:100-124) — same AnyNum images, same leaf order;
//     2. walks query->filter() with a query::FilterVisitor (sibling of ComparisonBuilder,
//        filter.cc:206-261) that emits the post-order predicate program of include/vgpu.h
//        instead of C++ text;
//     3. lowers dimension_cols() (+ granularity / rollup rules, boundaries via util::Duration::add_to
//        exactly like codegen/db/rollup.cc:44-75) and metric_cols() to vgpu_key / metric indices;
//     4. calls vgpu_query_agg — nothing is compiled per query;
//     5. runs the post-aggregation of src/codegen/query/post_agg.cc:26-147 + sort.cc:24-73 on the
//        host into the caller's RowOutput and fills the four QueryStats counters.
//   Visit(SelectQuery*), Visit(SearchQuery*) go to the GPU too (rows / distinct values picked on the device,
//   formatting and the sequential search replay here); Visit(ShowTablesQuery*) delegates to the stock runner.
//
// GpuTableBinding mirrors one db::Table into HBM: column base pointers come from SegmentAccess
// (segment_access.h); segments are re-uploaded when their size() changed (ingest appends rows) or
// when the caller invalidates them (ingest updated metric cells in place, upsert.cc:386-393).
#ifndef VIYADB_B200_HOST_GPU_QUERY_RUNNER_H_
#define VIYADB_B200_HOST_GPU_QUERY_RUNNER_H_

#include "../../include/vgpu.h"
#include "codegen/query/filter.h"
#include "db/column.h"
#include "db/database.h"
#include "db/dictionary.h"
#include "db/store.h"
#include "db/table.h"
#include "query/filter.h"
#include "query/output.h"
#include "query/query.h"
#include "query/runner.h"
#include "query/stats.h"
#include "ingest_epoch.h"
#include "segment_access.h"
#include "util/format.h"
#include "util/string.h"
#include "util/time.h"
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <functional>
#include <atomic>
#include <map>
#include <set>
#include <unordered_set>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

namespace vgpu_host {

namespace db = viya::db;
namespace query = viya::query;
namespace util = viya::util;
namespace cg = viya::codegen;

// The adapter casts these enums of the reference straight into their include/vgpu.h counterparts: the orders must agree.
static_assert((int)db::Metric::AggregationType::MAX == VGPU_AGG_MAX && (int)db::Metric::AggregationType::MIN == VGPU_AGG_MIN &&
                  (int)db::Metric::AggregationType::SUM == VGPU_AGG_SUM && (int)db::Metric::AggregationType::AVG == VGPU_AGG_AVG &&
                  (int)db::Metric::AggregationType::COUNT == VGPU_AGG_COUNT &&
                  (int)db::Metric::AggregationType::BITSET == VGPU_AGG_BITSET,
              "db::Metric::AggregationType (src/db/column.h:244) vs vgpu_agg");
static_assert((int)util::TimeUnit::YEAR == VGPU_TU_YEAR && (int)util::TimeUnit::MONTH == VGPU_TU_MONTH &&
                  (int)util::TimeUnit::WEEK == VGPU_TU_WEEK && (int)util::TimeUnit::DAY == VGPU_TU_DAY &&
                  (int)util::TimeUnit::HOUR == VGPU_TU_HOUR && (int)util::TimeUnit::MINUTE == VGPU_TU_MINUTE &&
                  (int)util::TimeUnit::SECOND == VGPU_TU_SECOND,
              "util::TimeUnit (src/util/time.h:27) vs vgpu_time_unit");
static_assert((int)query::RelOpFilter::Operator::EQUAL == VGPU_OP_EQ && (int)query::RelOpFilter::Operator::NOT_EQUAL == VGPU_OP_NE &&
                  (int)query::RelOpFilter::Operator::LESS == VGPU_OP_LT && (int)query::RelOpFilter::Operator::LESS_EQUAL == VGPU_OP_LE &&
                  (int)query::RelOpFilter::Operator::GREATER == VGPU_OP_GT &&
                  (int)query::RelOpFilter::Operator::GREATER_EQUAL == VGPU_OP_GE,
              "query::RelOpFilter::Operator (src/query/filter.h:54-61) vs vgpu_relop");

inline void check(int rc, const char *what) {
  if (rc != VGPU_OK) throw std::runtime_error(std::string(what) + ": " + vgpu_last_error());
}

inline uint32_t vgpu_type_of(const db::Column *col) {
  if (col->type() == db::Column::Type::DIMENSION) {
    auto dim = static_cast<const db::Dimension *>(col);
    if (dim->dim_type() == db::Dimension::DimType::NUMERIC) {
      switch (static_cast<const db::NumDimension *>(dim)->num_type().type()) {
      case db::NumericType::BYTE: return VGPU_I8;
      case db::NumericType::UBYTE: return VGPU_U8;
      case db::NumericType::SHORT: return VGPU_I16;
      case db::NumericType::USHORT: return VGPU_U16;
      case db::NumericType::INT: return VGPU_I32;
      case db::NumericType::UINT: return VGPU_U32;
      case db::NumericType::LONG: return VGPU_I64;
      case db::NumericType::ULONG: return VGPU_U64;
      case db::NumericType::FLOAT: return VGPU_F32;
      case db::NumericType::DOUBLE: return VGPU_F64;
      }
    }
  } else {
    auto metric = static_cast<const db::Metric *>(col);
    if (metric->agg_type() != db::Metric::AggregationType::BITSET) {
      switch (static_cast<const db::NumericType &>(static_cast<const db::ValueMetric *>(metric)->num_type()).type()) {
      case db::NumericType::BYTE: return VGPU_I8;
      case db::NumericType::UBYTE: return VGPU_U8;
      case db::NumericType::SHORT: return VGPU_I16;
      case db::NumericType::USHORT: return VGPU_U16;
      case db::NumericType::INT: return VGPU_I32;
      case db::NumericType::UINT: return VGPU_U32;
      case db::NumericType::LONG: return VGPU_I64;
      case db::NumericType::ULONG: return VGPU_U64;
      case db::NumericType::FLOAT: return VGPU_F32;
      case db::NumericType::DOUBLE: return VGPU_F64;
      }
    }
  }
  switch (col->num_type().size()) {  // dictionary codes, time, boolean, bitset ids: unsigned
  case db::BaseNumType::_1: return VGPU_U8;
  case db::BaseNumType::_2: return VGPU_U16;
  case db::BaseNumType::_4: return VGPU_U32;
  default: return VGPU_U64;
  }
}

// ---------------------------------------------------------------------------------------------
// one db::Table resident in HBM
// ---------------------------------------------------------------------------------------------
class GpuTableBinding {
public:
  GpuTableBinding(vgpu_ctx *ctx, db::Table &table) : ctx_(ctx), table_(table), access_(table) {
    ndims_ = table.dimensions().size();
    nmetrics_ = table.metrics().size();
    hidden_ = access_.has_hidden_count();
    std::vector<vgpu_column> cols;
    for (auto *dim : table.dimensions()) {
      vgpu_column c{};
      switch (dim->dim_type()) {
      case db::Dimension::DimType::STRING: c.kind = VGPU_DIM_STRING; break;
      case db::Dimension::DimType::NUMERIC: c.kind = VGPU_DIM_NUMERIC; break;
      case db::Dimension::DimType::TIME:
        c.kind = static_cast<const db::TimeDimension *>(dim)->micro_precision() ? VGPU_DIM_MICROTIME : VGPU_DIM_TIME;
        break;
      case db::Dimension::DimType::BOOLEAN: c.kind = VGPU_DIM_BOOLEAN; break;
      }
      c.type = vgpu_type_of(dim);
      c.agg = VGPU_AGG_NONE;
      cols.push_back(c);
    }
    for (auto *m : table.metrics()) {
      vgpu_column c{};
      c.agg = static_cast<uint32_t>(m->agg_type());  // same enum order (include/vgpu.h)
      if (m->agg_type() == db::Metric::AggregationType::BITSET) {
        c.kind = VGPU_METRIC_BITSET;
        c.type = m->num_type().size() == db::BaseNumType::_8 ? VGPU_U64 : VGPU_U32;
        // ubyte / ushort bitsets travel as uint32, but the AnyNum image of a filter literal only defines the
        // bytes of its own type (src/db/column.h:98-121): tell the library which ones to look at
        if (m->num_type().size() == db::BaseNumType::_1) c.lit_type = VGPU_U8 + 1;
        if (m->num_type().size() == db::BaseNumType::_2) c.lit_type = VGPU_U16 + 1;
      } else {
        c.kind = VGPU_METRIC_VALUE;
        c.type = vgpu_type_of(m);
      }
      cols.push_back(c);
    }
    if (hidden_) cols.push_back(vgpu_column{VGPU_METRIC_HIDDEN_COUNT, VGPU_U64, VGPU_AGG_COUNT, 0});
    vgpu_schema schema{};
    schema.ncols = static_cast<uint32_t>(cols.size());
    schema.ndims = static_cast<uint32_t>(ndims_);
    schema.segment_size = table.segment_size();
    schema.cols = cols.data();
    check(vgpu_table_create(ctx, &schema, &handle_), "vgpu_table_create");
    for (auto &c : cols) widths_.push_back(c.kind == VGPU_METRIC_BITSET ? 0u : width_of(c.type));
    has_bitset_ = false;
    for (auto &c : cols) has_bitset_ = has_bitset_ || c.kind == VGPU_METRIC_BITSET;
    std::lock_guard<std::mutex> lk(registry_mu());
    registry()[&table_] = this;
  }
  ~GpuTableBinding() {
    {
      std::lock_guard<std::mutex> lk(registry_mu());
      registry().erase(&table_);
    }
    IngestEpoch::Forget(&table_);
    vgpu_table_free(handle_);
    for (const void *p : pinned_) vgpu_host_unpin(ctx_, p);
  }
  GpuTableBinding(const GpuTableBinding &) = delete;
  GpuTableBinding &operator=(const GpuTableBinding &) = delete;

  vgpu_table *handle() const { return handle_; }
  size_t schema_index(const db::Column *col) const {
    return col->type() == db::Column::Type::DIMENSION ? col->index() : ndims_ + col->index();
  }
  bool has_hidden_count() const { return hidden_; }
  size_t hidden_count_index() const { return ndims_ + nmetrics_; }  // last schema column (include/vgpu.h)

  // Exact change notification (incremental sync): rows [row_begin, row_end) of a segment were updated
  // in place. The upsert knows both numbers where it calls `m.Update(upsert_tuple.m, tuple_idx)`
  // (src/codegen/db/upsert.cc:386-393); an integration that reports them there never needs
  // IngestEpoch::Bump(): only the dirty ranges and the rows appended since the last upload cross PCIe.
  void MarkDirty(size_t seg_idx, size_t row_begin, size_t row_end) {
    std::lock_guard<std::mutex> lk(mu_);
    if (row_end <= row_begin) return;
    if (dirty_.size() <= seg_idx) dirty_.resize(seg_idx + 1);
    dirty_[seg_idx].emplace_back(row_begin, row_end);
  }
  static void MarkDirty(const db::Table *table, size_t seg_idx, size_t row_begin, size_t row_end) {
    std::lock_guard<std::mutex> lk(registry_mu());
    auto it = registry().find(table);
    if (it != registry().end()) it->second->MarkDirty(seg_idx, row_begin, row_end);
  }
  uint64_t partial_updates() const { return partial_updates_; }   // vgpu_segment_update calls issued so far (tests)

  // Force the next Sync() to re-upload a segment whose metric cells were updated in place.
  void Invalidate(size_t seg_idx) {
    std::lock_guard<std::mutex> lk(mu_);
    if (seg_idx < uploaded_.size()) uploaded_[seg_idx] = static_cast<size_t>(-1);
  }
  void InvalidateAll() {
    std::lock_guard<std::mutex> lk(mu_);
    std::fill(uploaded_.begin(), uploaded_.end(), static_cast<size_t>(-1));
  }

  // Snapshot semantics of scan.cc:42-44: the segment list is copied, each size() read once.
  void Sync() {
    std::lock_guard<std::mutex> lk(mu_);
    // an ingest batch finished since the last upload: cells of any segment may have been updated in place
    const uint64_t epoch = IngestEpoch::Load(&table_);
    if (epoch != epoch_seen_) {
      std::fill(uploaded_.begin(), uploaded_.end(), static_cast<size_t>(-1));
      epoch_seen_ = epoch;
    }
    auto segments = table_.store()->segments_copy();
    if (uploaded_.size() < segments.size()) uploaded_.resize(segments.size(), static_cast<size_t>(-1));
    if (dirty_.size() < segments.size()) dirty_.resize(segments.size());
    std::vector<const void *> dims(ndims_), metrics(nmetrics_);
    std::vector<uint64_t> stats(2 * ndims_ + 2);
    // CSR images of bitset cells must outlive the asynchronous copies: kept until vgpu_table_sync below
    std::vector<std::vector<uint64_t>> offsets_keep;
    std::vector<std::vector<uint32_t>> values_keep;
    std::vector<std::vector<uint64_t>> wide_keep;
    std::vector<std::unique_ptr<vgpu_bitset_csr>> csr_keep;
    bool any = false;
    for (size_t si = 0; si < segments.size(); ++si) {
      size_t size = segments[si]->size();
      if (uploaded_[si] == size && dirty_[si].empty()) continue;
      // incremental: the rows reported dirty plus the rows appended since the last upload, nothing else
      if (uploaded_[si] != static_cast<size_t>(-1) && !has_bitset_ && SyncRanges(segments[si], si, size)) {
        uploaded_[si] = size;
        dirty_[si].clear();
        continue;
      }
      dirty_[si].clear();
      // the segment is one heap object holding every fixed-width column: page-lock it once, so that this and every
      // later upload is a straight DMA (a failure only means the driver stages the copies)
      if (pinned_.insert(segments[si]).second && vgpu_host_pin(ctx_, segments[si], access_.segment_bytes()) != VGPU_OK)
        pinned_.erase(segments[si]);
      const void *hidden = nullptr;
      access_.columns()(segments[si], dims.data(), metrics.data(), &hidden, stats.data());
      std::vector<const void *> ptrs;
      for (size_t d = 0; d < ndims_; ++d) ptrs.push_back(dims[d]);
      for (auto *m : table_.metrics()) {
        if (m->agg_type() != db::Metric::AggregationType::BITSET) {
          ptrs.push_back(metrics[m->index()]);
          continue;
        }
        // flatten util::Bitset cells to CSR (ids widened by the accessor, narrowed here to uint32)
        offsets_keep.emplace_back(size + 1);
        auto &offsets = offsets_keep.back();
        uint64_t total = access_.bitset()(segments[si], m->index(), size, offsets.data(), nullptr);
        wide_keep.emplace_back(total + 1);
        auto &wide = wide_keep.back();
        access_.bitset()(segments[si], m->index(), size, offsets.data(), wide.data());
        if (m->num_type().size() == db::BaseNumType::_8) {   // util::Bitset<8> = Roaring64Map: ids travel as uint64
          csr_keep.emplace_back(new vgpu_bitset_csr{offsets.data(), wide.data(), total});
        } else {
          values_keep.emplace_back(total + 1);
          auto &values = values_keep.back();
          for (uint64_t i = 0; i < total; ++i) values[i] = static_cast<uint32_t>(wide[i]);
          csr_keep.emplace_back(new vgpu_bitset_csr{offsets.data(), values.data(), total});
        }
        ptrs.push_back(csr_keep.back().get());
      }
      if (hidden_) ptrs.push_back(hidden);
      check(vgpu_segment_put_async(handle_, static_cast<uint32_t>(si), size, ptrs.data()), "vgpu_segment_put_async");
      uploaded_[si] = size;
      any = true;
    }
    if (any) check(vgpu_table_sync(handle_), "vgpu_table_sync");
  }

private:
  static uint32_t width_of(uint32_t type) {
    switch (type) {
    case VGPU_U8: case VGPU_I8: return 1;
    case VGPU_U16: case VGPU_I16: return 2;
    case VGPU_U32: case VGPU_I32: case VGPU_F32: return 4;
    default: return 8;
    }
  }
  static std::map<const db::Table *, GpuTableBinding *> &registry() {
    static std::map<const db::Table *, GpuTableBinding *> r;
    return r;
  }
  static std::mutex &registry_mu() {
    static std::mutex m;
    return m;
  }
  // upload the dirty ranges of one segment (merged) and what was appended; false: the library wants the whole segment
  bool SyncRanges(db::SegmentBase *seg, size_t si, size_t size) {
    std::vector<std::pair<size_t, size_t>> ranges = dirty_[si];
    if (size > uploaded_[si]) ranges.emplace_back(uploaded_[si], size);
    std::sort(ranges.begin(), ranges.end());
    std::vector<std::pair<size_t, size_t>> merged;
    for (auto &r : ranges) {
      size_t b = r.first, e = std::min(r.second, size);
      if (e <= b) continue;
      if (!merged.empty() && b <= merged.back().second) merged.back().second = std::max(merged.back().second, e);
      else merged.emplace_back(b, e);
    }
    std::vector<const void *> dims(ndims_), metrics(nmetrics_);
    std::vector<uint64_t> stats(2 * ndims_ + 2);
    const void *hidden = nullptr;
    access_.columns()(seg, dims.data(), metrics.data(), &hidden, stats.data());
    for (auto &r : merged) {
      std::vector<const void *> ptrs;
      size_t c = 0;
      for (size_t d = 0; d < ndims_; ++d, ++c) ptrs.push_back(static_cast<const uint8_t *>(dims[d]) + r.first * widths_[c]);
      for (size_t m = 0; m < nmetrics_; ++m, ++c) ptrs.push_back(static_cast<const uint8_t *>(metrics[m]) + r.first * widths_[c]);
      if (hidden_) ptrs.push_back(static_cast<const uint8_t *>(hidden) + r.first * 8);
      int rc = vgpu_segment_update(handle_, static_cast<uint32_t>(si), r.first, r.second - r.first, ptrs.data());
      if (rc == VGPU_ERR_STATE || rc == VGPU_ERR_UNSUPPORTED) return false;
      check(rc, "vgpu_segment_update");
      ++partial_updates_;
    }
    return true;
  }

  std::vector<uint32_t> widths_;
  bool has_bitset_ = false;
  std::vector<std::vector<std::pair<size_t, size_t>>> dirty_;
  uint64_t partial_updates_ = 0;
  vgpu_ctx *ctx_;
  db::Table &table_;
  SegmentAccess access_;
  vgpu_table *handle_ = nullptr;
  std::set<const void *> pinned_;
  size_t ndims_ = 0, nmetrics_ = 0;
  bool hidden_ = false;
  std::vector<size_t> uploaded_;
  uint64_t epoch_seen_ = 0;
  std::mutex mu_;
};

// ---------------------------------------------------------------------------------------------
// filter tree -> predicate program (post order, reference leaf order == FilterArgsPacker order)
// ---------------------------------------------------------------------------------------------
class PredicateProgramBuilder : public query::FilterVisitor {
public:
  PredicateProgramBuilder(const db::Table &table, const GpuTableBinding &binding)
      : table_(table), binding_(binding), argidx_(0) {}

  void Visit(const query::RelOpFilter *filter) override {
    vgpu_pred_node n{};
    n.kind = VGPU_NODE_RELOP;
    n.op = static_cast<uint32_t>(filter->op());  // same enum order
    n.col = static_cast<uint32_t>(binding_.schema_index(table_.column(filter->column())));
    n.arg = argidx_++;
    nodes_.push_back(n);
  }
  void Visit(const query::InFilter *filter) override {
    vgpu_pred_node n{};
    n.kind = VGPU_NODE_IN;
    n.op = filter->equal() ? 1 : 0;
    n.col = static_cast<uint32_t>(binding_.schema_index(table_.column(filter->column())));
    n.arg = argidx_;
    n.n = static_cast<uint32_t>(filter->values().size());
    argidx_ += n.n;
    nodes_.push_back(n);
  }
  void Visit(const query::CompositeFilter *filter) override {
    for (auto f : filter->filters()) f->Accept(*this);
    vgpu_pred_node n{};
    n.kind = filter->op() == query::CompositeFilter::Operator::AND ? VGPU_NODE_AND : VGPU_NODE_OR;
    n.n = static_cast<uint32_t>(filter->filters().size());
    nodes_.push_back(n);
  }
  void Visit(const query::EmptyFilter *) override {
    vgpu_pred_node n{};
    n.kind = VGPU_NODE_EMPTY;
    nodes_.push_back(n);
  }
  const std::vector<vgpu_pred_node> &nodes() const { return nodes_; }

private:
  const db::Table &table_;
  const GpuTableBinding &binding_;
  uint32_t argidx_;
  std::vector<vgpu_pred_node> nodes_;
};

// HAVING on the host: the comparison tree of FilterComparison (filter.cc:206-261) over one group.
class HavingEvaluator : public query::FilterVisitor {
public:
  using Getter = std::function<void(const db::Column *, uint64_t &bits)>;
  HavingEvaluator(const db::Table &table, std::vector<db::AnyNum> &args, Getter getter)
      : table_(table), args_(args), getter_(std::move(getter)), argidx_(0), result_(true) {}

  bool result() const { return result_; }

  void Visit(const query::RelOpFilter *filter) override {
    const auto col = table_.column(filter->column());
    result_ = Compare(col, static_cast<int>(filter->op()), args_[argidx_++]);
  }
  void Visit(const query::InFilter *filter) override {
    const auto col = table_.column(filter->column());
    bool r = !filter->equal();
    for (size_t i = 0; i < filter->values().size(); ++i) {
      bool c = Compare(col, filter->equal() ? 0 : 1, args_[argidx_++]);
      r = filter->equal() ? (r | c) : (r & c);
    }
    result_ = r;
  }
  void Visit(const query::CompositeFilter *filter) override {
    bool is_and = filter->op() == query::CompositeFilter::Operator::AND;
    bool r = is_and;
    for (auto f : filter->filters()) {
      f->Accept(*this);  // no short circuit: every leaf consumes its arguments
      r = is_and ? (r & result_) : (r | result_);
    }
    result_ = r;
  }
  void Visit(const query::EmptyFilter *) override { result_ = true; }

private:
  template <typename T> static bool Cmp(int op, T a, T b) {
    switch (op) {
    case 0: return a == b;
    case 1: return a != b;
    case 2: return a < b;
    case 3: return a <= b;
    case 4: return a > b;
    default: return a >= b;
    }
  }
  bool Compare(const db::Column *col, int op, db::AnyNum arg) {
    uint64_t bits = 0;
    getter_(col, bits);
    bool bitset = col->type() == db::Column::Type::METRIC &&
                  static_cast<const db::Metric *>(col)->agg_type() == db::Metric::AggregationType::BITSET;
    uint32_t t = vgpu_type_of(col);
    if (bitset) {  // cardinality (uint64 in the result) against the literal of the id type
      uint64_t a = 0;
      switch (col->num_type().size()) {
      case db::BaseNumType::_1: a = arg.get_uint8_t(); break;
      case db::BaseNumType::_2: a = arg.get_uint16_t(); break;
      case db::BaseNumType::_4: a = arg.get_uint32_t(); break;
      default: a = arg.get_uint64_t(); break;
      }
      return Cmp<uint64_t>(op, bits, a);
    }
    switch (t) {
    case VGPU_U8: return Cmp<uint8_t>(op, (uint8_t)bits, arg.get_uint8_t());
    case VGPU_U16: return Cmp<uint16_t>(op, (uint16_t)bits, arg.get_uint16_t());
    case VGPU_U32: return Cmp<uint32_t>(op, (uint32_t)bits, arg.get_uint32_t());
    case VGPU_U64: return Cmp<uint64_t>(op, bits, arg.get_uint64_t());
    case VGPU_I32: return Cmp<int32_t>(op, (int32_t)(uint32_t)bits, arg.get_int32_t());
    case VGPU_I64: return Cmp<int64_t>(op, (int64_t)bits, arg.get_int64_t());
    case VGPU_F32: { float f; uint32_t b = (uint32_t)bits; std::memcpy(&f, &b, 4); return Cmp<float>(op, f, (float)arg.get_float()); }
    case VGPU_F64: { double d; std::memcpy(&d, &bits, 8); return Cmp<double>(op, d, arg.get_double()); }
    default:
      // AnyNum has no get_int8_t / get_int16_t (column.h:110-117): the reference cannot compile these
      throw std::runtime_error("filters on byte/short columns are not supported by the reference");
    }
  }

  const db::Table &table_;
  std::vector<db::AnyNum> &args_;
  Getter getter_;
  size_t argidx_;
  bool result_;
};

// ---------------------------------------------------------------------------------------------
// the visitor
// ---------------------------------------------------------------------------------------------
class GpuQueryRunner : public query::QueryVisitor {
public:
  // `bindings` outlives the runner (one per Database): table name -> HBM mirror.
  using Bindings = std::map<std::string, std::unique_ptr<GpuTableBinding>>;

  GpuQueryRunner(db::Database &database, query::RowOutput &output, vgpu_ctx *ctx, Bindings &bindings)
      : database_(database), output_(output), ctx_(ctx), bindings_(bindings), stock_(database, output),
        stats_(database.statsd()) {}

  void Visit(query::ShowTablesQuery *query) override { delegated_ = true; stock_.Visit(query); }

  // Replaces QueryRunner::Visit(SelectQuery*) (src/query/runner.cc:29-43): the device picks the rows the
  // generated viya_query_select would send (codegen/query/scan.cc:75-166, skip / limit rules included) and
  // returns their raw cells; formatting stays here, with the reference's own util::Format.
  void Visit(query::SelectQuery *query) override {
    stats_.OnBegin("select", query->table().name());
    auto &table = query->table();
    GpuTableBinding &binding = Bind(table);
    binding.Sync();
    stats_.OnCompile();
    cg::FilterArgsPacker filter_args(table);
    query->filter()->Accept(filter_args);
    std::vector<db::AnyNum> fargs = filter_args.args();
    std::vector<uint64_t> raw_args(fargs.size());
    for (size_t i = 0; i < fargs.size(); ++i) std::memcpy(&raw_args[i], &fargs[i], 8);
    PredicateProgramBuilder pred(table, binding);
    query->filter()->Accept(pred);

    auto &dim_cols = query->dimension_cols();
    auto &metric_cols = query->metric_cols();
    std::vector<uint32_t> cols;
    for (auto &dc : dim_cols) cols.push_back(static_cast<uint32_t>(binding.schema_index(dc.dim())));
    for (auto &mc : metric_cols) cols.push_back(static_cast<uint32_t>(binding.schema_index(mc.metric())));
    // AVG cells are divided by the first selected COUNT metric, else by the hidden count (scan.cc:133-154)
    int count_pos = -1;
    bool has_avg = false;
    for (size_t m = 0; m < metric_cols.size(); ++m) {
      auto agg = metric_cols[m].metric()->agg_type();
      has_avg |= agg == db::Metric::AggregationType::AVG;
      if (count_pos < 0 && agg == db::Metric::AggregationType::COUNT) count_pos = (int)(dim_cols.size() + m);
    }
    bool hidden_count = false;
    if (has_avg && count_pos < 0) {
      if (!binding.has_hidden_count()) throw std::runtime_error("AVG metric selected but the table has no count column");
      count_pos = (int)cols.size();
      cols.push_back(static_cast<uint32_t>(binding.hidden_count_index()));
      hidden_count = true;
    }
    vgpu_rows_plan plan{};
    plan.nnodes = static_cast<uint32_t>(pred.nodes().size());
    plan.nodes = pred.nodes().data();
    plan.nargs = static_cast<uint32_t>(raw_args.size());
    plan.args = raw_args.data();
    plan.ncols = static_cast<uint32_t>(cols.size());
    plan.cols = cols.data();
    plan.skip = query->skip();
    plan.limit = query->limit();
    vgpu_rows *res = nullptr;
    check(vgpu_query_select(binding.handle(), &plan, &res), "vgpu_query_select");
    std::unique_ptr<vgpu_rows, void (*)(vgpu_rows *)> guard(res, vgpu_rows_free);
    vgpu_rows_view view{};
    check(vgpu_rows_get(res, &view), "vgpu_rows_get");
    stats_.scanned_recs += view.scanned_recs;
    stats_.scanned_segments += view.scanned_segments;

    using Row = std::vector<std::string>;
    Row row(dim_cols.size() + metric_cols.size());
    util::Format fmt;
    output_.Start();
    if (query->header()) {
      for (auto &dc : dim_cols) row[dc.index()] = dc.dim()->name();
      for (auto &mc : metric_cols) row[mc.index()] = mc.metric()->name();
      output_.Send(row);
    }
    for (uint64_t r = 0; r < view.nrows; ++r) {
      for (size_t k = 0; k < dim_cols.size(); ++k) {
        auto dim = dim_cols[k].dim();
        uint64_t bits = Load(view.cells[k], (uint32_t)dim->num_type().size(), r);
        auto &cell = row[dim_cols[k].index()];
        if (dim->dim_type() == db::Dimension::DimType::STRING) {
          auto dict = static_cast<const db::StrDimension *>(dim)->dict();
          dict->lock().lock_shared();
          cell = dict->c2v()[bits];
          dict->lock().unlock_shared();
        } else if (dim->dim_type() == db::Dimension::DimType::TIME && !dim_cols[k].format().empty()) {
          cell = fmt.date(dim_cols[k].format().c_str(), (uint32_t)bits);
        } else if (dim->dim_type() == db::Dimension::DimType::BOOLEAN) {
          cell = bits ? "true" : "false";
        } else {
          cell = FormatNum(fmt, vgpu_type_of(dim), bits);
        }
      }
      for (size_t m = 0; m < metric_cols.size(); ++m) {
        auto metric = metric_cols[m].metric();
        auto &cell = row[metric_cols[m].index()];
        const void *base = view.cells[dim_cols.size() + m];
        if (metric->agg_type() == db::Metric::AggregationType::BITSET) {
          cell = fmt.num((uint64_t)Load(base, 8, r));
          continue;
        }
        uint32_t type = vgpu_type_of(metric);
        uint64_t bits = Load(base, (uint32_t)metric->num_type().size(), r);
        if (metric->agg_type() == db::Metric::AggregationType::AVG) {
          double cnt = hidden_count
                           ? (double)Load(view.cells[count_pos], 8, r)
                           : AsDouble(vgpu_type_of(metric_cols[count_pos - dim_cols.size()].metric()),
                                      Load(view.cells[count_pos],
                                           (uint32_t)metric_cols[count_pos - dim_cols.size()].metric()->num_type().size(), r));
          cell = fmt.num(AsDouble(type, bits) / cnt);
        } else {
          cell = FormatNum(fmt, type, bits);
        }
      }
      output_.Send(row);
      ++stats_.output_recs;
    }
    output_.Flush();
    stats_.OnEnd();
  }

  // Replaces QueryRunner::Visit(SearchQuery*) (src/query/runner.cc:66-80): the device finds, per processed
  // segment, the distinct values of the dimension among the passing rows with their first rows; the sequential
  // part of scan.cc:273-295 (codes.insert, substring match, `limit` breaking the tuple loop only) is replayed
  // here on those short lists, then post_agg.cc:149-166 (SendAsCol).
  void Visit(query::SearchQuery *query) override {
    auto dim = query->dimension();
    const uint32_t dim_type = vgpu_type_of(dim);
    if (dim_type == VGPU_F32 || dim_type == VGPU_F64) {  // the dense first-row table is keyed by integer values
      delegated_ = true;
      stock_.Visit(query);
      return;
    }
    stats_.OnBegin("search", query->table().name());
    auto &table = query->table();
    GpuTableBinding &binding = Bind(table);
    binding.Sync();
    stats_.OnCompile();
    cg::FilterArgsPacker filter_args(table);
    query->filter()->Accept(filter_args);
    std::vector<db::AnyNum> fargs = filter_args.args();
    std::vector<uint64_t> raw_args(fargs.size());
    for (size_t i = 0; i < fargs.size(); ++i) std::memcpy(&raw_args[i], &fargs[i], 8);
    PredicateProgramBuilder pred(table, binding);
    query->filter()->Accept(pred);
    vgpu_search_plan plan{};
    plan.nnodes = static_cast<uint32_t>(pred.nodes().size());
    plan.nodes = pred.nodes().data();
    plan.nargs = static_cast<uint32_t>(raw_args.size());
    plan.args = raw_args.data();
    plan.col = static_cast<uint32_t>(binding.schema_index(dim));
    vgpu_search *res = nullptr;
    check(vgpu_query_search(binding.handle(), &plan, &res), "vgpu_query_search");
    std::unique_ptr<vgpu_search, void (*)(vgpu_search *)> guard(res, vgpu_search_free);
    vgpu_search_view view{};
    check(vgpu_search_get(res, &view), "vgpu_search_get");
    stats_.scanned_recs += view.scanned_recs;
    stats_.scanned_segments += view.scanned_segments;

    util::Format fmt;
    std::unordered_set<uint64_t> codes;
    std::vector<std::string> values;
    std::string check_value;
    const std::string &term = query->term();
    const size_t limit = query->limit();
    for (uint32_t si = 0; si < view.nsegments; ++si) {
      for (uint64_t i = view.seg_offsets[si]; i < view.seg_offsets[si + 1]; ++i) {
        const uint64_t code = view.codes[i];
        if (!codes.insert(code).second) continue;
        if (dim->dim_type() == db::Dimension::DimType::STRING) {
          auto dict = static_cast<const db::StrDimension *>(dim)->dict();
          dict->lock().lock_shared();
          check_value = dict->c2v()[code];
          dict->lock().unlock_shared();
        } else if (dim->dim_type() == db::Dimension::DimType::BOOLEAN) {
          check_value = code ? "true" : "false";
        } else {
          check_value = FormatNum(fmt, vgpu_type_of(dim), code);
        }
        if (check_value.find(term) != std::string::npos) {
          values.push_back(check_value);
          if (limit > 0 && values.size() >= limit) break;  // the tuple loop of this segment only (scan.cc:291)
        }
      }
    }
    stats_.aggregated_recs = codes.size();
    output_.Start();
    if (query->header()) output_.Send(std::vector<std::string>{dim->name()});
    output_.SendAsCol(values);
    stats_.output_recs = values.size();
    output_.Flush();
    stats_.OnEnd();
  }

  void Visit(query::AggregateQuery *query) override {
    stats_.OnBegin("aggregate", query->table().name());
    auto &table = query->table();
    GpuTableBinding &binding = Bind(table);
    binding.Sync();
    stats_.OnCompile();  // nothing is compiled: the plan is data

    // 1. literals, packed by the reference's own code
    cg::FilterArgsPacker filter_args(table);
    query->filter()->Accept(filter_args);
    cg::FilterArgsPacker having_args(table);
    if (query->having() != nullptr) query->having()->Accept(having_args);
    std::vector<db::AnyNum> fargs = filter_args.args();
    std::vector<uint64_t> raw_args(fargs.size());
    static_assert(sizeof(db::AnyNum) == 8, "AnyNum is an 8-byte image");
    for (size_t i = 0; i < fargs.size(); ++i) std::memcpy(&raw_args[i], &fargs[i], 8);

    // 2. predicate program
    PredicateProgramBuilder pred(table, binding);
    query->filter()->Accept(pred);

    // 3. keys (+ rollup) and metrics
    std::vector<vgpu_key> keys;
    for (auto &dim_col : query->dimension_cols()) {
      vgpu_key k{};
      k.col = static_cast<uint32_t>(dim_col.dim()->index());
      k.query_granularity = VGPU_TU_NONE;
      if (dim_col.dim()->dim_type() == db::Dimension::DimType::TIME) {
        auto time_dim = static_cast<const db::TimeDimension *>(dim_col.dim());
        auto &rules = time_dim->rollup_rules();
        if (rules.size() > VGPU_MAX_ROLLUP_RULES) throw std::runtime_error("too many rollup rules");
        k.nrules = static_cast<uint32_t>(rules.size());
        for (size_t r = 0; r < rules.size(); ++r) {
          // rollup_b = Duration(unit,count).add_to((uint32_t) now, -1) [* 1000000L]   (rollup.cc:59-69)
          uint64_t b = rules[r].after().add_to((uint32_t)RollupNow(), -1);
          if (time_dim->micro_precision()) b *= 1000000UL;
          k.rule_boundary[r] = b;
          k.rule_granularity[r] = static_cast<uint32_t>(rules[r].granularity().time_unit());
        }
        if (!dim_col.granularity().empty())
          k.query_granularity = static_cast<uint32_t>(dim_col.granularity().time_unit());
      }
      keys.push_back(k);
    }
    std::vector<uint32_t> metric_cols;
    bool has_avg = false, has_count = false;
    for (auto &metric_col : query->metric_cols()) {
      metric_cols.push_back(static_cast<uint32_t>(binding.schema_index(metric_col.metric())));
      has_avg |= metric_col.metric()->agg_type() == db::Metric::AggregationType::AVG;
      has_count |= metric_col.metric()->agg_type() == db::Metric::AggregationType::COUNT;
    }
    vgpu_plan plan{};
    plan.nnodes = static_cast<uint32_t>(pred.nodes().size());
    plan.nodes = pred.nodes().data();
    plan.nargs = static_cast<uint32_t>(raw_args.size());
    plan.args = raw_args.data();
    plan.nkeys = static_cast<uint32_t>(keys.size());
    plan.keys = keys.data();
    plan.nmetrics = static_cast<uint32_t>(metric_cols.size());
    plan.metric_cols = metric_cols.data();
    plan.need_hidden_count = (has_avg && !has_count) ? 1 : 0;

    // 3b. post-aggregation on the device (include/vgpu.h: HAVING, top-N). Both only remove groups; PostAggregate below
    // stays what it was. HAVING only where the reference tests every group: with a sort, or without skip / limit
    // (without a sort it cuts the skip / limit window out of the map iteration first, post_agg.cc:26-83).
    plan.sort_col = VGPU_NO_COLUMN;
    std::vector<uint64_t> raw_hargs;
    PredicateProgramBuilder hpred(table, binding);
    if (getenv("VGPU_HOST_POST") == nullptr) {
      plan.flags |= VGPU_PLAN_POST;
      auto sort_columns = query->sort_cols();
      if (query->having() != nullptr && (!sort_columns.empty() || (query->skip() == 0 && query->limit() == 0))) {
        std::vector<db::AnyNum> hargs = having_args.args();
        raw_hargs.resize(hargs.size());
        for (size_t i = 0; i < hargs.size(); ++i) std::memcpy(&raw_hargs[i], &hargs[i], 8);
        query->having()->Accept(hpred);
        plan.nhnodes = static_cast<uint32_t>(hpred.nodes().size());
        plan.hnodes = hpred.nodes().data();
        plan.nhargs = static_cast<uint32_t>(raw_hargs.size());
        plan.hargs = raw_hargs.data();
      }
      if (!sort_columns.empty() && query->limit() > 0) {
        plan.sort_col = static_cast<uint32_t>(binding.schema_index(sort_columns[0].col()));
        plan.sort_descending = sort_columns[0].ascending() ? 0 : 1;
        plan.top_k = query->skip() + query->limit();
      }
    }

    // 4. the hot path
    vgpu_result *res = nullptr;
    check(vgpu_query_agg(binding.handle(), &plan, &res), "vgpu_query_agg");
    std::unique_ptr<vgpu_result, void (*)(vgpu_result *)> guard(res, vgpu_result_free);
    vgpu_result_view view{};
    check(vgpu_result_get(res, &view), "vgpu_result_get");
    stats_.scanned_recs += view.scanned_recs;
    stats_.scanned_segments += view.scanned_segments;
    stats_.aggregated_recs = view.aggregated_recs;

    // 5. post aggregation on the host
    PostAggregate(query, view, having_args.args());
    stats_.OnEnd();
  }

  const query::QueryStats &stats() const { return delegated_ ? stock_.stats() : stats_; }

private:
  GpuTableBinding &Bind(db::Table &table) {
    // `query_threads` queries may bind at once (src/db/database.cc:28-29): the map is shared by every runner of a database
    static std::mutex bind_mu;
    std::lock_guard<std::mutex> lk(bind_mu);
    auto it = bindings_.find(table.name());
    if (it == bindings_.end())
      it = bindings_.emplace(table.name(), std::make_unique<GpuTableBinding>(ctx_, table)).first;
    return *it->second;
  }

  static long RollupNow() {
    // VIYA_TEST_ROLLUP_TS pins "now" exactly like codegen/db/rollup.cc:47-49
    const char *test_ts = getenv("VIYA_TEST_ROLLUP_TS");
    if (test_ts != nullptr) return std::strtol(test_ts, nullptr, 10);
    return static_cast<long>(std::time(nullptr));
  }

  static uint64_t Load(const void *base, uint32_t width, uint64_t i) {
    uint64_t v = 0;
    std::memcpy(&v, static_cast<const char *>(base) + i * width, width);
    return v;
  }

  const char *FormatNum(util::Format &fmt, uint32_t type, uint64_t bits) {
    switch (type) {
    case VGPU_U8: return fmt.num((uint8_t)bits);
    case VGPU_U16: return fmt.num((uint16_t)bits);
    case VGPU_U32: return fmt.num((uint32_t)bits);
    case VGPU_U64: return fmt.num((uint64_t)bits);
    case VGPU_I8: return fmt.num((int8_t)bits);
    case VGPU_I16: return fmt.num((int16_t)bits);
    case VGPU_I32: return fmt.num((int32_t)(uint32_t)bits);
    case VGPU_I64: return fmt.num((int64_t)bits);
    case VGPU_F32: { float f; uint32_t b = (uint32_t)bits; std::memcpy(&f, &b, 4); return fmt.num(f); }
    default: { double d; std::memcpy(&d, &bits, 8); return fmt.num(d); }
    }
  }

  static double AsDouble(uint32_t type, uint64_t bits) {
    switch (type) {
    case VGPU_U8: case VGPU_U16: case VGPU_U32: case VGPU_U64: return (double)bits;
    case VGPU_I8: return (double)(int8_t)bits;
    case VGPU_I16: return (double)(int16_t)bits;
    case VGPU_I32: return (double)(int32_t)(uint32_t)bits;
    case VGPU_I64: return (double)(int64_t)bits;
    case VGPU_F32: { float f; uint32_t b = (uint32_t)bits; std::memcpy(&f, &b, 4); return (double)f; }
    default: { double d; std::memcpy(&d, &bits, 8); return d; }
    }
  }

  // src/codegen/query/post_agg.cc:26-147 + sort.cc:24-73, interpreted instead of generated
  void PostAggregate(query::AggregateQuery *query, const vgpu_result_view &view, std::vector<db::AnyNum> hargs) {
    using Row = std::vector<std::string>;
    auto &dim_cols = query->dimension_cols();
    auto &metric_cols = query->metric_cols();
    output_.Start();
    size_t n = view.ngroups;
    // skip / limit are clamped by agg_map.size() (post_agg.cc:40-47) = ALL groups, also when HAVING / top-N already ran on
    // the device and the view only holds the survivors
    size_t n_all = std::max<size_t>(n, view.aggregated_recs);
    size_t skip = std::min(n_all, query->skip());
    size_t limit = std::min(query->limit(), n_all - skip);
    auto sort_columns = query->sort_cols();
    size_t lo = 0, hi = n;
    if (sort_columns.empty()) {
      lo = skip;
      if (limit > 0) hi = lo + limit;
    }
    Row row(dim_cols.size() + metric_cols.size());
    util::Format fmt;
    if (query->header()) {
      for (auto &dc : dim_cols) row[dc.index()] = dc.dim()->name();
      for (auto &mc : metric_cols) row[mc.index()] = mc.metric()->name();
      output_.Send(row);
    }
    // count column for AVG: first selected COUNT metric, else the hidden one
    int count_metric = -1;
    for (size_t m = 0; m < metric_cols.size(); ++m)
      if (metric_cols[m].metric()->agg_type() == db::Metric::AggregationType::COUNT) { count_metric = (int)m; break; }

    std::vector<Row> post_agg;
    auto &table = query->table();
    for (size_t g = lo; g < hi; ++g) {
      if (query->having() != nullptr) {
        HavingEvaluator ev(table, hargs, [&](const db::Column *col, uint64_t &bits) {
          if (col->type() == db::Column::Type::DIMENSION) {
            for (size_t k = 0; k < dim_cols.size(); ++k)
              if (dim_cols[k].dim() == col) { bits = Load(view.keys[k], (uint32_t)col->num_type().size(), g); return; }
          } else {
            for (size_t m = 0; m < metric_cols.size(); ++m)
              if (metric_cols[m].metric() == col) {
                bool bitset = metric_cols[m].metric()->agg_type() == db::Metric::AggregationType::BITSET;
                bits = Load(view.accs[m], bitset ? 8 : (uint32_t)col->num_type().size(), g);
                return;
              }
          }
          throw std::invalid_argument("Column '" + col->name() + " is not selected");
        });
        query->having()->Accept(ev);
        if (!ev.result()) continue;
      }
      for (size_t k = 0; k < dim_cols.size(); ++k) {
        auto dim = dim_cols[k].dim();
        uint64_t bits = Load(view.keys[k], (uint32_t)dim->num_type().size(), g);
        auto &cell = row[dim_cols[k].index()];
        if (dim->dim_type() == db::Dimension::DimType::STRING) {
          auto dict = static_cast<const db::StrDimension *>(dim)->dict();
          dict->lock().lock_shared();
          cell = dict->c2v()[bits];
          dict->lock().unlock_shared();
        } else if (dim->dim_type() == db::Dimension::DimType::TIME && !dim_cols[k].format().empty()) {
          cell = fmt.date(dim_cols[k].format().c_str(), (uint32_t)bits);
        } else if (dim->dim_type() == db::Dimension::DimType::BOOLEAN) {
          cell = bits ? "true" : "false";
        } else {
          cell = FormatNum(fmt, vgpu_type_of(dim), bits);
        }
      }
      for (size_t m = 0; m < metric_cols.size(); ++m) {
        auto metric = metric_cols[m].metric();
        auto &cell = row[metric_cols[m].index()];
        if (metric->agg_type() == db::Metric::AggregationType::BITSET) {
          cell = fmt.num((uint64_t)Load(view.accs[m], 8, g));
          continue;
        }
        uint32_t type = vgpu_type_of(metric);
        uint64_t bits = Load(view.accs[m], (uint32_t)metric->num_type().size(), g);
        if (metric->agg_type() == db::Metric::AggregationType::AVG) {
          double cnt = count_metric >= 0
                           ? AsDouble(vgpu_type_of(metric_cols[count_metric].metric()),
                                      Load(view.accs[count_metric], (uint32_t)metric_cols[count_metric].metric()->num_type().size(), g))
                           : (double)view.hidden_count[g];
          // `sum / (double) count`: the quotient is a double whatever the sum's type (post_agg.cc:126-127)
          cell = fmt.num(AsDouble(type, bits) / cnt);
        } else {
          cell = FormatNum(fmt, type, bits);
        }
      }
      if (sort_columns.empty()) {
        output_.Send(row);
        ++stats_.output_recs;
      } else {
        post_agg.push_back(row);
      }
    }
    if (!sort_columns.empty()) {
      auto row_less = [&sort_columns](const Row &a, const Row &b) {
        size_t sc_size = sort_columns.size();
        for (size_t i = 0; i < sc_size; ++i) {
          auto &sc = sort_columns[i];
          size_t c = sc.index();
          bool lt, gt;
          switch (sc.col()->sort_type()) {
          case db::Column::SortType::STRING:
            lt = sc.ascending() ? a[c] < b[c] : a[c] > b[c];
            gt = sc.ascending() ? b[c] < a[c] : b[c] > a[c];
            break;
          case db::Column::SortType::INTEGER:
            lt = sc.ascending() ? util::StringNumCmp::SmallerInt(a[c], b[c]) : util::StringNumCmp::GreaterInt(a[c], b[c]);
            gt = sc.ascending() ? util::StringNumCmp::SmallerInt(b[c], a[c]) : util::StringNumCmp::GreaterInt(b[c], a[c]);
            break;
          default:
            lt = sc.ascending() ? util::StringNumCmp::SmallerFloat(a[c], b[c]) : util::StringNumCmp::GreaterFloat(a[c], b[c]);
            gt = sc.ascending() ? util::StringNumCmp::SmallerFloat(b[c], a[c]) : util::StringNumCmp::GreaterFloat(b[c], a[c]);
            break;
          }
          if (lt) return true;
          if (i < sc_size - 1 && gt) return false;
        }
        return false;
      };
      // The reference sorts everything and sends [skip, skip + limit) (sort.cc:32-73). Only that window has to be in
      // order: a partial sort of its upper end is the same rows in the same order (std::sort leaves ties unspecified
      // as well) in O(n log(skip + limit)) — what remains of a large group table when the sort key is one the device
      // cannot rank (strings, times, floats, AVG).
      if (limit > 0 && skip + limit < post_agg.size())
        std::partial_sort(post_agg.begin(), post_agg.begin() + (skip + limit), post_agg.end(), row_less);
      else
        std::sort(post_agg.begin(), post_agg.end(), row_less);
      size_t end = limit > 0 ? std::min(post_agg.size(), skip + limit) : post_agg.size();
      for (size_t i = std::min(skip, post_agg.size()); i < end; ++i) {
        output_.Send(post_agg[i]);
        ++stats_.output_recs;
      }
    }
    output_.Flush();
  }

  db::Database &database_;
  query::RowOutput &output_;
  vgpu_ctx *ctx_;
  Bindings &bindings_;
  query::QueryRunner stock_;
  query::QueryStats stats_;
  bool delegated_ = false;
};

// End of an ingest batch into `table` (input::Loader::AfterLoad in the integration): hand the rows the generated upsert
// code reported as updated in place (IngestDirty, ingest_epoch.h) to the table's resident copy. Rows appended by the batch
// need no report: Sync() sees the segments grow.
inline void FlushIngest(const db::Table *table) {
  auto &pending = IngestDirty::pending();
  for (auto &r : pending)
    if (r.table == table) GpuTableBinding::MarkDirty(table, r.seg, r.lo, r.hi);
  pending.clear();
}

// What a db::Database owns when the B200 path is selected by configuration ("gpu": true): the device context and the HBM
// copies of its tables. viyadb_b200/host/viyadb_database.patch adds `std::unique_ptr<vgpu_host::DatabaseGpu> gpu_` to
// db::Database and routes Database::Query through GpuQueryRunner when it is set.
struct DatabaseGpu {
  vgpu_ctx *ctx = nullptr;
  GpuQueryRunner::Bindings bindings;
  explicit DatabaseGpu(int device) { check(vgpu_init(device, &ctx), "vgpu_init"); }
  ~DatabaseGpu() {
    bindings.clear();   // HBM copies first, then the context they live in
    if (ctx) vgpu_shutdown(ctx);
  }
  DatabaseGpu(const DatabaseGpu &) = delete;
  DatabaseGpu &operator=(const DatabaseGpu &) = delete;
};

} // namespace vgpu_host

#endif // VIYADB_B200_HOST_GPU_QUERY_RUNNER_H_
