// viyadb_b200/host/segment_access.h — host-side adapter code (C++17, builds against the
// reference's headers; no reference file is edited).
//
// Problem: the reference's physical column store is a JIT-generated class
//   class Segment : public db::SegmentBase { Dimensions d; Metrics m; SegmentStats stats; }
// emitted by codegen::StoreDefs (src/codegen/db/store.cc:203-356). Host code outside the JIT
// never sees that type, so column base addresses are not available to a plugin.
//
// Solution: do what the reference itself does for its store functions
// (src/codegen/db/store.cc StoreFunctions, src/db/table.cc:87-91) — generate, ONCE PER TABLE
// SCHEMA (not per query), a tiny translation unit that re-emits the very same StoreDefs text and
// exports extern "C" accessors which hand out the column base pointers, per-segment min/max stats
// and flattened (CSR) bitset cells. It goes through the reference's own codegen::Compiler, so it
// lands in the same on-disk .so cache and is compiled with the same flags — the layout can never
// drift from the reference's, whatever the schema.
#ifndef VIYADB_B200_HOST_SEGMENT_ACCESS_H_
#define VIYADB_B200_HOST_SEGMENT_ACCESS_H_

#include "codegen/compiler.h"
#include "codegen/db/store.h"
#include "codegen/generator.h"
#include "db/column.h"
#include "db/database.h"
#include "db/segment.h"
#include "db/table.h"
#include <cstdint>
#include <memory>
#include <string>

namespace vgpu_host {

namespace db = viya::db;
namespace cg = viya::codegen;

// dims[i] / metrics[j] receive the base address of column arrays `T _i[segment_size]`.
// A BITSET metric slot receives the address of the util::Bitset<N> array (opaque; use the
// flatten call). hidden_count receives `uint64_t _count[]` or nullptr (store.cc:286-289).
// stats receives dmax/dmin pairs as raw 8-byte slots: stats[2*i] = dmax_i, stats[2*i+1] = dmin_i
// (zero/sign-extended or bit-copied for fp) for NUMERIC/TIME dims, untouched for others.
using SegmentColumnsFn = void (*)(db::SegmentBase *seg, const void **dims, const void **metrics,
                                  const void **hidden_count, uint64_t *stats);
// Flattens rows [0,nrows) of bitset metric `metric_idx` into CSR. With values == nullptr only
// offsets[0..nrows] are written (sizing pass). Values are widened to uint64_t.
using SegmentBitsetFn = uint64_t (*)(db::SegmentBase *seg, uint32_t metric_idx, uint64_t nrows,
                                     uint64_t *offsets, uint64_t *values);
// sizeof(Segment): the heap object that holds every fixed-width column of a segment (pinned once for DMA)
using SegmentSizeofFn = uint64_t (*)();

class SegmentAccess {
public:
  explicit SegmentAccess(db::Table &table) : table_(table) {
    lib_ = const_cast<db::Database &>(table.database()).compiler().Compile(GenerateCode());
    columns_ = lib_->GetFunction<SegmentColumnsFn>("vgpu_segment_columns");
    bitset_ = lib_->GetFunction<SegmentBitsetFn>("vgpu_segment_bitset");
    sizeof_ = lib_->GetFunction<SegmentSizeofFn>("vgpu_segment_sizeof");
  }

  SegmentColumnsFn columns() const { return columns_; }
  SegmentBitsetFn bitset() const { return bitset_; }
  uint64_t segment_bytes() const { return sizeof_(); }

  bool has_hidden_count() const {
    bool has_avg = false, has_count = false;
    for (auto *m : table_.metrics()) {
      if (m->agg_type() == db::Metric::AggregationType::AVG) has_avg = true;
      if (m->agg_type() == db::Metric::AggregationType::COUNT) has_count = true;
    }
    return has_avg && !has_count;
  }

private:
  std::string GenerateCode() const {
    cg::Code code;
    code.AddHeaders({"cstring", "cstdint", "vector", "type_traits", "db/segment.h"});
    cg::StoreDefs store_defs(table_);
    code << store_defs.GenerateCode();

    // Mirror of util::Bitset<N> (src/util/bitset.h:26-67) with the same member sequence, so the
    // private roaring_ can be iterated without touching the reference header.
    bool has_bitset = false;
    for (auto *m : table_.metrics())
      if (m->agg_type() == db::Metric::AggregationType::BITSET) has_bitset = true;
    if (has_bitset) {
      code << "template <int SizeBytes> struct VgpuBitsetMirror {\n"
              " using RoaringType = typename std::conditional<SizeBytes == 8, Roaring64Map, Roaring>::type;\n"
              " using NumType = typename std::conditional<SizeBytes == 8, uint64_t, uint32_t>::type;\n"
              " NumType cardinality_;\n"
              " RoaringType roaring_;\n"
              "};\n";
    }

    const char *vis = "__attribute__((__visibility__(\"default\")))";
    code << "extern \"C\" void vgpu_segment_columns(db::SegmentBase* sb, const void** dims, "
            "const void** metrics, const void** hidden_count, uint64_t* stats) "
         << vis << ";\n";
    code << "extern \"C\" void vgpu_segment_columns(db::SegmentBase* sb, const void** dims, "
            "const void** metrics, const void** hidden_count, uint64_t* stats) {\n"
            " auto* s = static_cast<Segment*>(sb);\n";
    for (auto *dim : table_.dimensions()) {
      auto i = std::to_string(dim->index());
      code << " dims[" << i << "] = static_cast<const void*>(s->d._" << i << ");\n";
      if (dim->dim_type() == db::Dimension::DimType::NUMERIC ||
          dim->dim_type() == db::Dimension::DimType::TIME) {
        code << " stats[" << 2 * dim->index() << "] = 0; stats[" << 2 * dim->index() + 1
             << "] = 0;\n";
        code << " std::memcpy(&stats[" << 2 * dim->index() << "], &s->stats.dmax" << i
             << ", sizeof(s->stats.dmax" << i << "));\n";
        code << " std::memcpy(&stats[" << 2 * dim->index() + 1 << "], &s->stats.dmin" << i
             << ", sizeof(s->stats.dmin" << i << "));\n";
      }
    }
    for (auto *metric : table_.metrics()) {
      auto i = std::to_string(metric->index());
      code << " metrics[" << i << "] = static_cast<const void*>(s->m._" << i << ");\n";
    }
    if (has_hidden_count()) {
      code << " *hidden_count = static_cast<const void*>(s->m._count);\n";
    } else {
      code << " *hidden_count = nullptr;\n";
    }
    code << "}\n";

    code << "extern \"C\" uint64_t vgpu_segment_sizeof() " << vis << ";\n";
    code << "extern \"C\" uint64_t vgpu_segment_sizeof() { return sizeof(Segment); }\n";

    code << "extern \"C\" uint64_t vgpu_segment_bitset(db::SegmentBase* sb, uint32_t metric_idx, "
            "uint64_t nrows, uint64_t* offsets, uint64_t* values) "
         << vis << ";\n";
    code << "extern \"C\" uint64_t vgpu_segment_bitset(db::SegmentBase* sb, uint32_t metric_idx, "
            "uint64_t nrows, uint64_t* offsets, uint64_t* values) {\n"
            " auto* s = static_cast<Segment*>(sb);\n"
            " uint64_t total = 0;\n"
            " (void)s; (void)nrows; (void)offsets; (void)values;\n"
            " switch (metric_idx) {\n";
    for (auto *metric : table_.metrics()) {
      if (metric->agg_type() != db::Metric::AggregationType::BITSET) continue;
      auto i = std::to_string(metric->index());
      auto n = std::to_string(metric->num_type().size());
      code << " case " << i << ": {\n"
           << "  static_assert(sizeof(VgpuBitsetMirror<" << n << ">) == sizeof(util::Bitset<" << n
           << ">), \"bitset mirror drifted\");\n"
           << "  std::vector<uint32_t> tmp;\n"
           << "  for (uint64_t r = 0; r < nrows; ++r) {\n"
           << "   offsets[r] = total;\n"
           << "   auto* b = reinterpret_cast<VgpuBitsetMirror<" << n << ">*>(&s->m._" << i
           << "[r]);\n"
           << "   uint64_t c = b->roaring_.cardinality();\n"
           << "   if (values != nullptr && c > 0) {\n"
           << (metric->num_type().size() == 8
                   ? "    b->roaring_.toUint64Array(values + total);\n"
                   : "    tmp.resize(c); b->roaring_.toUint32Array(tmp.data());\n"
                     "    for (uint64_t k = 0; k < c; ++k) values[total + k] = tmp[k];\n")
           << "   }\n"
           << "   total += c;\n"
           << "  }\n"
           << "  offsets[nrows] = total;\n"
           << " } break;\n";
    }
    code << " default: break;\n"
            " }\n"
            " return total;\n"
            "}\n";
    return code.str();
  }

  db::Table &table_;
  std::shared_ptr<cg::SharedLibrary> lib_;
  SegmentColumnsFn columns_;
  SegmentBitsetFn bitset_;
  SegmentSizeofFn sizeof_;
};

} // namespace vgpu_host

#endif // VIYADB_B200_HOST_SEGMENT_ACCESS_H_
