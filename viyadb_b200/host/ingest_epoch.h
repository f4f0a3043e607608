// viyadb_b200/host/ingest_epoch.h — the one thing the ingest side of ViyaDB needs to know about the B200 path
// (included by src/input/loader.cc in the integration, viyadb_database.patch; no other dependency).
#ifndef VIYADB_B200_HOST_INGEST_EPOCH_H_
#define VIYADB_B200_HOST_INGEST_EPOCH_H_

#include <atomic>
#include <cstdint>
#include <cstddef>
#include <map>
#include <mutex>
#include <vector>

namespace vgpu_host {

// ---------------------------------------------------------------------------------------------
// Ingest notifications. The reference's upsert aggregates into EXISTING tuples in place
// (src/codegen/db/upsert.cc:386-393, `m.Update(upsert_tuple.m, tuple_idx)`): metric cells — and bitset cells — of
// any segment may change without any SegmentBase::size() changing, so "size unchanged" does not mean "resident copy
// still valid". Every load batch ends in input::Loader::AfterLoad() (src/input/loader.cc:39); the integration calls
// IngestEpoch::Bump(&table_) there (viyadb_database.patch; the gtest drop-in binary, which wraps that function from
// outside the class, can only call the coarse Bump()), and a binding whose epoch is behind re-uploads its table on the
// next query.
// ---------------------------------------------------------------------------------------------
struct IngestEpoch {
  static std::atomic<uint64_t> &counter() {
    static std::atomic<uint64_t> c{1};
    return c;
  }
  // coarse: some table of this process finished an ingest batch (what a hook without access to the loader's table can say)
  static void Bump() { counter().fetch_add(1, std::memory_order_release); }
  static uint64_t Load() { return counter().load(std::memory_order_acquire); }

  // per table (`table` = the db::Table the batch went into, input::Loader::table_): only the resident copy of THAT table
  // is stale — a static 30 GB table next to a table under continuous ingest is not uploaded again after every batch
  static void Bump(const void *table) {
    std::lock_guard<std::mutex> lk(table_mu());
    ++table_counters()[table];
  }
  // what a binding of `table` compares with the value it saw at its last upload: moves when either counter moves
  static uint64_t Load(const void *table) {
    uint64_t own = 0;
    {
      std::lock_guard<std::mutex> lk(table_mu());
      auto it = table_counters().find(table);
      if (it != table_counters().end()) own = it->second;
    }
    return Load() + own;
  }
  // a table that goes away (Database::DropTable, ~Database): a later table at the same address starts from scratch
  static void Forget(const void *table) {
    std::lock_guard<std::mutex> lk(table_mu());
    table_counters().erase(table);
  }

private:
  static std::mutex &table_mu() {
    static std::mutex m;
    return m;
  }
  static std::map<const void *, uint64_t> &table_counters() {
    static std::map<const void *, uint64_t> m;
    return m;
  }
};

// Exact notifications. viyadb_database.patch makes the generated upsert code call
// vgpu_ingest_mark_dirty(table, segment, tuple) where it updates an existing tuple in place
// (src/codegen/db/upsert.cc:386-393). The rows are collected per ingest thread, adjacent ones merged, and handed to the
// table's resident copy when the batch ends (Loader::AfterLoad -> vgpu_host::FlushIngest, gpu_query_runner.h): the next
// query then moves only those rows and the appended ones across PCIe (vgpu_segment_update) — no epoch bump needed.
struct IngestDirty {
  struct Range {
    const void *table;
    size_t seg, lo, hi;
  };
  static std::vector<Range> &pending() {
    static thread_local std::vector<Range> v;   // BeforeLoad .. Load .. AfterLoad of a batch run on one thread
    return v;
  }
  static void Mark(const void *table, size_t seg, size_t row) {
    auto &v = pending();
    if (!v.empty()) {
      Range &b = v.back();
      if (b.table == table && b.seg == seg && row + 1 >= b.lo && row <= b.hi) {
        if (row < b.lo) b.lo = row;
        if (row + 1 > b.hi) b.hi = row + 1;
        return;
      }
    }
    v.push_back(Range{table, seg, row, row + 1});
  }
};

} // namespace vgpu_host

#endif // VIYADB_B200_HOST_INGEST_EPOCH_H_
