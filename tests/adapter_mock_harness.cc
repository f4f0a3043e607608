// adapter_mock_harness.cc — TEST INFRASTRUCTURE. The C++ drop-in adapter (viyadb_b200/host/gpu_query_runner.h), UNMODIFIED
// and whole, inside a reference process without a GPU: this file DEFINES the C ABI of include/vgpu.h as a mock device —
// vgpu_table_create records the schema, vgpu_query_agg records the plan the adapter lowered and hands back a group table
// from the job file (tests/test_adapter_mock.py takes it from the oracle, which is pinned to the real reference on the
// same records) — and it is linked INSTEAD of libvgpu.so. So `query->Accept(GpuQueryRunner)` runs exactly the code a
// ViyaDB maintainer would ship: the reference's FilterArgsPacker, the predicate program, rollup boundaries through
// util::Duration, the post-aggregation request, then HAVING / formatting / sort / skip / limit on the returned groups.
// The mock computes nothing: it is not a CPU path of the product and nothing under viyadb_b200/ links it.
//
// job = {"table": {...}, "dicts": {"<string dim>": ["__exceeded", "v1", ...]}, "rollup_ts": N, "state_dir": "...",
//        "cases": [{"query": {...}, "scanned_recs": r, "scanned_segments": s,
//                   aggregate: "ngroups": n, "keys": [[bits...] per selected dimension], "accs": [[bits...] per selected
//                              metric], "hidden": [counts...] | null
//                   select:    "nrows": n, "cells": [[bits of the n rows to send] per SCHEMA column (hidden count last)]
//                   search:    "seg_offsets": [...], "codes": [...], "first_row": [...]   (vgpu_search_view)}]}
// or     {"table": ..., "state_dir": ..., "sync": [{"rows": [[...]], "notify": "none" | "epoch" | "mark"}, ...]}: every
//        step ingests its rows through the reference's own loader, notifies the binding and calls GpuTableBinding::Sync()
// or     {..., "cases": [...], "concurrent": {"threads": T, "repeat": R}}: every case from T threads at once (see below)
#include "db/database.h"
#include "db/dictionary.h"
#include "db/table.h"
#include "gpu_query_runner.h"
#include "input/simple.h"
#include "query/output.h"
#include "query/query.h"
#include "query/runner.h"
#include "util/config.h"
#include <atomic>
#include <chrono>
#include <thread>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <nlohmann/json.hpp>

using json = nlohmann::json;
namespace db = viya::db;
namespace util = viya::util;
namespace query = viya::query;

#include "mock_vgpu.h"   // the mock device: defines the C ABI of include/vgpu.h

// ---------------------------------------------------------------------------------------------
// sync mode: GpuTableBinding::Sync() against the live store of the reference. Rows go through the
// reference's own ingest (upserts merge into existing tuples IN PLACE, src/codegen/db/upsert.cc:386-393); after every
// Sync() the mock's shadow of "HBM" must equal the live segments, cell for cell.
// ---------------------------------------------------------------------------------------------
namespace {
void load_rows(db::Table *table, const json &rows) {
  struct L : viya::input::SimpleLoader {
    using viya::input::SimpleLoader::SimpleLoader;
    void Before() { BeforeLoad(); }
    void After() { AfterLoad(); }
  } l(*table);
  l.Before();
  for (auto &r : rows) {
    std::vector<std::string> row = r.get<std::vector<std::string>>();
    l.Load(row);
  }
  l.After();
}

json run_sync(const json &job, db::Database &database, db::Table *table) {
  json out = json::array();
  vgpu_host::GpuTableBinding binding(&g_ctx, *table);
  vgpu_host::SegmentAccess access(*table);
  for (auto &step : job["sync"]) {
    std::vector<size_t> sizes_before;
    for (auto *sgm : table->store()->segments_copy()) sizes_before.push_back(sgm->size());
    load_rows(table, step["rows"]);
    const std::string notify = step.value("notify", std::string("none"));
    if (notify == "mark") {
      // exact notifications, as an upsert hook would give them (here: every row that existed before the batch may have
      // been updated in place); appended rows are found by the binding itself
      for (size_t si = 0; si < sizes_before.size(); ++si) vgpu_host::GpuTableBinding::MarkDirty(table, si, 0, sizes_before[si]);
    } else if (notify == "epoch") {
      vgpu_host::IngestEpoch::Bump();
    }
    g_table->calls = json::array();
    const uint64_t partial_before = binding.partial_updates();
    binding.Sync();
    json r = compare_with_live(table, access);
    r["calls"] = g_table->calls;
    r["partial_updates"] = binding.partial_updates() - partial_before;
    r["notify"] = notify;
    out.push_back(r);
  }
  return out;
}
}  // namespace

// one case on the calling thread: arm the (thread-local) canned answer of the mock, run the query through the adapter
json run_case(const json &job, const json &c, db::Database &database, vgpu_host::GpuQueryRunner::Bindings &bindings) {
      json res;
      try {
        query::MemoryRowOutput output;
        query::QueryFactory factory;
        std::unique_ptr<query::Query> qq(factory.Create(util::Config(c["query"]), database));
        auto *aq = dynamic_cast<query::AggregateQuery *>(qq.get());
        std::vector<std::vector<char>> kbuf, abuf;
        std::vector<const void *> kptr, aptr;
        std::vector<uint64_t> hidden;
        g_canned = vgpu_result_view{};
        g_canned.scanned_recs = c.value("scanned_recs", (uint64_t)0);
        g_canned.scanned_segments = c.value("scanned_segments", (uint64_t)0);
        if (aq != nullptr) {
          const uint64_t n = c["ngroups"].get<uint64_t>();
          auto &dim_cols = aq->dimension_cols();
          auto &metric_cols = aq->metric_cols();
          kbuf.resize(dim_cols.size()); abuf.resize(metric_cols.size());
          kptr.resize(dim_cols.size()); aptr.resize(metric_cols.size());
          for (size_t k = 0; k < dim_cols.size(); ++k) {
            const uint32_t w = (uint32_t)dim_cols[k].dim()->num_type().size();
            kbuf[k].resize(n * w + 8);
            for (uint64_t g = 0; g < n; ++g) {
              uint64_t bits = c["keys"][k][g].get<uint64_t>();
              std::memcpy(kbuf[k].data() + g * w, &bits, w);
            }
            kptr[k] = kbuf[k].data();
          }
          for (size_t m = 0; m < metric_cols.size(); ++m) {
            auto metric = metric_cols[m].metric();
            const uint32_t w = metric->agg_type() == db::Metric::AggregationType::BITSET ? 8u : (uint32_t)metric->num_type().size();
            abuf[m].resize(n * w + 8);
            for (uint64_t g = 0; g < n; ++g) {
              uint64_t bits = c["accs"][m][g].get<uint64_t>();
              std::memcpy(abuf[m].data() + g * w, &bits, w);
            }
            aptr[m] = abuf[m].data();
          }
          if (!c["hidden"].is_null()) hidden = c["hidden"].get<std::vector<uint64_t>>();
          g_canned.ngroups = n;
          g_canned.nkeys = (uint32_t)dim_cols.size();
          g_canned.nmetrics = (uint32_t)metric_cols.size();
          g_canned.keys = kptr.data();
          g_canned.accs = aptr.data();
          g_canned.hidden_count = hidden.empty() ? nullptr : hidden.data();
          g_canned.aggregated_recs = n;
        } else if (dynamic_cast<query::SelectQuery *>(qq.get()) != nullptr) {
          g_sel_nrows = c["nrows"].get<uint64_t>();
          g_sel_cells = c["cells"].get<std::vector<std::vector<uint64_t>>>();
        } else if (dynamic_cast<query::SearchQuery *>(qq.get()) != nullptr) {
          g_srch_offsets = c["seg_offsets"].get<std::vector<uint64_t>>();
          g_srch_codes = c["codes"].get<std::vector<uint64_t>>();
          g_srch_rows = c["first_row"].get<std::vector<uint32_t>>();
        } else {
          throw std::runtime_error("unsupported query type");
        }
        g_plan = json();
        vgpu_host::GpuQueryRunner runner(database, output, &g_ctx, bindings);
        const auto t0 = std::chrono::steady_clock::now();
        qq->Accept(runner);
        const double host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        auto &s = runner.stats();
        json rows_out = json::array();
        if (job.value("rows_out", true)) rows_out = output.rows();
        else rows_out = {{"count", output.rows().size()}};   // timing runs: do not dump a million rows
        res = {{"rows", rows_out}, {"plan", g_plan}, {"schema", g_schema}, {"host_ms", host_ms},   // the mock's time is ~0

               {"stats", {{"scanned_segments", s.scanned_segments}, {"scanned_recs", s.scanned_recs},
                          {"aggregated_recs", s.aggregated_recs}, {"output_recs", s.output_recs}}}};
      } catch (const std::invalid_argument &e) {
        res = {{"error", e.what()}, {"error_type", "invalid_argument"}};
      } catch (const std::exception &e) {
        res = {{"error", e.what()}, {"error_type", "exception"}};
      }
      return res;
}

// ---------------------------------------------------------------------------------------------
// concurrent mode: the reference runs `query_threads` queries at once next to an ingest thread (src/db/database.cc:28-33).
// T threads run every case R times through their own GpuQueryRunner over the SHARED bindings of the database while
// another thread keeps sending ingest notifications for the table; every answer must equal the single-threaded one.
// Built with -fsanitize=thread (oracle/Makefile: adapter_mock_tsan) this is the data-race check of the adapter's host code.
// ---------------------------------------------------------------------------------------------
json run_concurrent(const json &job, db::Database &database, db::Table *table) {
  const int threads = job["concurrent"].value("threads", 4), repeat = job["concurrent"].value("repeat", 3);
  std::vector<json> baseline;
  {
    vgpu_host::GpuQueryRunner::Bindings single;
    for (auto &c : job["cases"]) baseline.push_back(run_case(job, c, database, single)["rows"]);
  }
  vgpu_host::GpuQueryRunner::Bindings bindings;   // fresh: the threads race to bind the table (GpuQueryRunner::Bind)
  std::atomic<bool> stop{false};
  std::atomic<uint64_t> mismatches{0}, queries{0}, errors{0};
  std::thread notifier([&] {
    uint64_t i = 0;
    while (!stop.load()) {
      vgpu_host::GpuTableBinding::MarkDirty(table, 0, i % 7, i % 7 + 1);
      if (i % 3 == 0) vgpu_host::IngestEpoch::Bump(table);
      if (i % 11 == 0) vgpu_host::IngestEpoch::Bump();
      vgpu_host::IngestDirty::Mark(table, 0, i % 5);
      vgpu_host::FlushIngest(table);
      ++i;
      std::this_thread::yield();
    }
  });
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t)
    pool.emplace_back([&, t] {
      for (int r = 0; r < repeat; ++r)
        for (size_t i = 0; i < job["cases"].size(); ++i) {
          const size_t k = (i + t) % job["cases"].size();   // different threads, different queries at the same time
          json res = run_case(job, job["cases"][k], database, bindings);
          ++queries;
          if (res.count("error")) ++errors;
          else if (res["rows"] != baseline[k]) ++mismatches;
        }
    });
  for (auto &th : pool) th.join();
  stop.store(true);
  notifier.join();
  bindings.clear();
  return {{"threads", threads}, {"queries", queries.load()}, {"mismatches", mismatches.load()}, {"errors", errors.load()}};
}

// ---------------------------------------------------------------------------------------------
int main(int argc, char **argv) {
  if (argc < 2) {
    std::cerr << "usage: adapter_mock_cli <job.json>\n";
    return 2;
  }
  std::ifstream in(argv[1]);
  json job;
  in >> job;
  if (job.count("rollup_ts")) {
    std::string v = std::to_string(job["rollup_ts"].get<long>()) + "L";
    setenv("VIYA_TEST_ROLLUP_TS", v.c_str(), 1);
  }
  json out;
  try {
    json dbconf;
    dbconf["state_dir"] = job.value("state_dir", std::string("/tmp/vgpu_fuzz_state"));
    dbconf["tables"] = json::array({job["table"]});
    db::Database database{util::Config(dbconf)};
    auto *table = database.GetTable(job["table"]["name"].get<std::string>());
    // dictionaries in code order, the way the generated upsert code fills them (code = c2v.size(), both maps): code 0 is
    // "__exceeded" already (dictionary.cc:22-25)
    for (auto *dim : table->dimensions()) {
      if (dim->dim_type() != db::Dimension::DimType::STRING) continue;
      auto dict = static_cast<const db::StrDimension *>(dim)->dict();
      if (!job.count("dicts")) break;   // sync mode: the reference's own ingest fills the dictionaries
      auto &vals = job["dicts"][dim->name()];
      for (size_t i = 1; i < vals.size(); ++i) {
        const std::string v = vals[i].get<std::string>();
        const uint64_t code = dict->c2v().size();
        dict->c2v().push_back(v);
        switch (dim->num_type().size()) {
        case db::BaseNumType::_1: reinterpret_cast<db::DictImpl<uint8_t> *>(dict->v2c())->insert(std::make_pair(v, (uint8_t)code)); break;
        case db::BaseNumType::_2: reinterpret_cast<db::DictImpl<uint16_t> *>(dict->v2c())->insert(std::make_pair(v, (uint16_t)code)); break;
        case db::BaseNumType::_4: reinterpret_cast<db::DictImpl<uint32_t> *>(dict->v2c())->insert(std::make_pair(v, (uint32_t)code)); break;
        default: reinterpret_cast<db::DictImpl<uint64_t> *>(dict->v2c())->insert(std::make_pair(v, (uint64_t)code)); break;
        }
      }
    }
    if (job.count("sync")) {
      out["sync"] = run_sync(job, database, table);
      std::cout << out.dump() << std::endl;
      return 0;
    }
    if (job.count("concurrent")) {
      out["concurrent"] = run_concurrent(job, database, table);
      std::cout << out.dump() << std::endl;
      return 0;
    }
    vgpu_host::GpuQueryRunner::Bindings bindings;
    out["results"] = json::array();
    for (auto &c : job["cases"]) out["results"].push_back(run_case(job, c, database, bindings));
    bindings.clear();
  } catch (const std::exception &e) {
    out["fatal"] = e.what();
    std::cout << out.dump() << std::endl;
    return 1;
  }
  std::cout << out.dump() << std::endl;
  return 0;
}
