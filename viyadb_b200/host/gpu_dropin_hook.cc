// viyadb_b200/host/gpu_dropin_hook.cc — the reference's OWN gtest suite on top of GpuQueryRunner.
//
// Linked with the unmodified objects of the reference's 14 core gtest files (test/{aggregation,bitset,boolean,
// codegen,filter,index,limits,load,metrics,partitioning,search,select,sort,time}.cc) and
//   -Wl,--wrap=<Database::Database(Config const&)>   every database gets the relocatable JIT cache directory
//   -Wl,--wrap=<Database::Query(Config const&, RowOutput&)>  the one line of the integration:
//                                                     construct GpuQueryRunner instead of query::QueryRunner at
//                                                     src/db/database.cc:104-111
//   -Wl,--wrap=<Database::~Database()>                drop the HBM mirror of a database that goes away
//   -Wl,--wrap=<input::Loader::AfterLoad()>           IngestEpoch::Bump(): resident copies are stale after an ingest batch
// so every `db.Query(...)` the reference's tests issue — aggregate, select, search, show tables — runs through the C ABI
// on the GPU, against the test's own EXPECT_EQs. Built by oracle/Makefile (target gpu_tests) where the reference's
// headers are; the binary travels to the GPU box with gpurun, together with the pre-warmed JIT cache of the
// reference's upsert functions (VGPU_HOOK_MODE=prewarm runs the stock path and compiles the accessors, no GPU needed).
#include "db/database.h"
#include "db/table.h"
#include "gpu_query_runner.h"
#include "input/loader.h"
#include "query/output.h"
#include "query/query.h"
#include "query/runner.h"
#include "query/stats.h"
#include "util/config.h"
#include <cstdlib>
#include <map>
#include <mutex>
#include <nlohmann/json.hpp>
#include <string>

namespace {
std::mutex g_mu;
vgpu_ctx *g_ctx = nullptr;
std::map<const void *, vgpu_host::GpuQueryRunner::Bindings> g_bindings;

bool prewarm_mode() {
  const char *m = getenv("VGPU_HOOK_MODE");
  return m && std::string(m) == "prewarm";
}
}  // namespace

extern "C" {

void real_db_ctor(viya::db::Database *self, const viya::util::Config &conf)
    __asm__("__real__ZN4viya2db8DatabaseC1ERKNS_4util6ConfigE");
void wrap_db_ctor(viya::db::Database *self, const viya::util::Config &conf)
    __asm__("__wrap__ZN4viya2db8DatabaseC1ERKNS_4util6ConfigE");
void wrap_db_ctor(viya::db::Database *self, const viya::util::Config &conf) {
  nlohmann::json j = nlohmann::json::parse(conf.dump());
  const char *sd = getenv("VGPU_STATE_DIR");
  if (sd && !j.count("state_dir")) j["state_dir"] = std::string(sd);
  viya::util::Config relocated(j);
  real_db_ctor(self, relocated);
}

void real_db_dtor(viya::db::Database *self) __asm__("__real__ZN4viya2db8DatabaseD1Ev");
void wrap_db_dtor(viya::db::Database *self) __asm__("__wrap__ZN4viya2db8DatabaseD1Ev");
void wrap_db_dtor(viya::db::Database *self) {
  {
    std::lock_guard<std::mutex> lk(g_mu);
    g_bindings.erase(self);
  }
  real_db_dtor(self);
}

viya::db::UpsertStats real_after_load(viya::input::Loader *self) __asm__("__real__ZN4viya5input6Loader9AfterLoadEv");
viya::db::UpsertStats wrap_after_load(viya::input::Loader *self) __asm__("__wrap__ZN4viya5input6Loader9AfterLoadEv");
viya::db::UpsertStats wrap_after_load(viya::input::Loader *self) {
  viya::db::UpsertStats s = real_after_load(self);
  vgpu_host::IngestEpoch::Bump();
  return s;
}

viya::query::QueryStats real_db_query(viya::db::Database *self, const viya::util::Config &q, viya::query::RowOutput &out)
    __asm__("__real__ZN4viya2db8Database5QueryERKNS_4util6ConfigERNS_5query9RowOutputE");
viya::query::QueryStats wrap_db_query(viya::db::Database *self, const viya::util::Config &q, viya::query::RowOutput &out)
    __asm__("__wrap__ZN4viya2db8Database5QueryERKNS_4util6ConfigERNS_5query9RowOutputE");
viya::query::QueryStats wrap_db_query(viya::db::Database *self, const viya::util::Config &q, viya::query::RowOutput &out) {
  if (prewarm_mode()) {
    // authoring container, no GPU: compile what the GPU path will dlopen on the box (the accessors), run the stock path
    if (q.exists("table")) {
      try {
        vgpu_host::SegmentAccess access(*self->GetTable(q.str("table")));
      } catch (const std::exception &) {
      }
    }
    return real_db_query(self, q, out);
  }
  vgpu_host::GpuQueryRunner::Bindings *bindings;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_ctx) vgpu_host::check(vgpu_init(0, &g_ctx), "vgpu_init");
    bindings = &g_bindings[self];
  }
  viya::query::QueryFactory factory;
  std::unique_ptr<viya::query::Query> qq(factory.Create(q, *self));
  vgpu_host::GpuQueryRunner runner(*self, out, g_ctx, *bindings);
  qq->Accept(runner);
  return runner.stats();
}

}  // extern "C"
