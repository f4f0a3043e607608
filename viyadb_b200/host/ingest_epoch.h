// viyadb_b200/host/ingest_epoch.h — the one thing the ingest side of ViyaDB needs to know about the B200 path
// (included by src/input/loader.cc in the integration, viyadb_database.patch; no other dependency).
#ifndef VIYADB_B200_HOST_INGEST_EPOCH_H_
#define VIYADB_B200_HOST_INGEST_EPOCH_H_

#include <atomic>
#include <cstdint>

namespace vgpu_host {

// ---------------------------------------------------------------------------------------------
// Ingest notifications. The reference's upsert aggregates into EXISTING tuples in place
// (src/codegen/db/upsert.cc:386-393, `m.Update(upsert_tuple.m, tuple_idx)`): metric cells — and bitset cells — of
// any segment may change without any SegmentBase::size() changing, so "size unchanged" does not mean "resident copy
// still valid". Every load batch ends in input::Loader::AfterLoad() (src/input/loader.cc:39); the integration calls
// IngestEpoch::Bump() there (the gtest drop-in binary wraps that very function), and a binding whose
// epoch is behind re-uploads its table on the next query.
// ---------------------------------------------------------------------------------------------
struct IngestEpoch {
  static std::atomic<uint64_t> &counter() {
    static std::atomic<uint64_t> c{1};
    return c;
  }
  static void Bump() { counter().fetch_add(1, std::memory_order_release); }
  static uint64_t Load() { return counter().load(std::memory_order_acquire); }
};

} // namespace vgpu_host

#endif // VIYADB_B200_HOST_INGEST_EPOCH_H_
