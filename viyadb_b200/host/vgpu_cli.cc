// viyadb_b200/host/vgpu_cli.cc — the drop-in, end to end, inside ONE reference process.
//
// Builds a reference db::Database from a job file (same format as oracle/oracle_cli.cc), loads rows
// through the reference's own ingest path, then runs every query TWICE through the reference's own
// query objects:
//     q->Accept(query::QueryRunner)          stock g++-JIT path            -> "stock"
//     q->Accept(vgpu_host::GpuQueryRunner)   C ABI -> sm_100a kernels      -> "gpu"
// and prints both row sets and QueryStats. tests/test_dropin_cpp.py compares them on the GPU box.
// With "stock_only": true nothing touches the GPU (used in the authoring container to pre-warm the
// reference's JIT .so cache, which travels to the GPU box in place of /root/reference).
#include "db/database.h"
#include "db/table.h"
#include "gpu_query_runner.h"
#include "input/simple.h"
#include "query/output.h"
#include "query/query.h"
#include "query/runner.h"
#include "util/config.h"
#include <chrono>
#include <fstream>
#include <iostream>
#include <libgen.h>
#include <limits.h>
#include <nlohmann/json.hpp>
#include <unistd.h>

using json = nlohmann::json;
namespace db = viya::db;
namespace util = viya::util;
namespace query = viya::query;
namespace input = viya::input;

static json stats_json(const query::QueryStats &s) {
  return {{"scanned_segments", s.scanned_segments}, {"scanned_recs", s.scanned_recs},
          {"aggregated_recs", s.aggregated_recs}, {"output_recs", s.output_recs},
          {"whole_ms", s.whole_time.count() * 1e3}, {"compile_ms", s.compile_time.count() * 1e3}};
}

int main(int argc, char **argv) {
  if (argc < 2) {
    std::cerr << "usage: vgpu_cli <job.json>\n";
    return 2;
  }
  std::ifstream in(argv[1]);
  json job;
  in >> job;
  // the reference JIT resolves its include / library paths relative to CWD (compiler.cc:46-54)
  std::string root = job.value("ref_root", std::string());
  if (!root.empty() && chdir((root + "/build").c_str()) != 0) {
    std::cerr << "cannot chdir to " << root << "/build\n";
    return 2;
  }
  if (job.count("rollup_ts")) {
    std::string v = std::to_string(job["rollup_ts"].get<long>()) + "L";
    setenv("VIYA_TEST_ROLLUP_TS", v.c_str(), 1);
  }
  bool stock_only = job.value("stock_only", false);
  json out;
  try {
    json dbconf;
    dbconf["state_dir"] = job.value("state_dir", std::string("/tmp/viyadb_oracle"));
    dbconf["tables"] = json::array({job["table"]});
    db::Database database{util::Config(dbconf)};
    auto *table = database.GetTable(job["table"]["name"].get<std::string>());
    if (job.count("rows")) {
      struct L : input::SimpleLoader {
        using input::SimpleLoader::SimpleLoader;
        void Before() { BeforeLoad(); }
        void After() { AfterLoad(); }
      } l(*table);
      l.Before();
      for (auto &r : job["rows"]) {
        std::vector<std::string> row = r.get<std::vector<std::string>>();
        l.Load(row);
      }
      l.After();
    }
    if (job.count("generate")) {
      // same synthetic stream as oracle_cli / vgpu_segment_generate:
      // value = lo + splitmix64(seed * 0x100000001B3 + row * 16 + col) % range, as "<prefix><value>"
      auto &g = job["generate"];
      uint64_t n = g["n"].get<uint64_t>(), seed = g.value("seed", (uint64_t)42), row0 = g.value("row_offset", (uint64_t)0);
      struct L : input::SimpleLoader {
        using input::SimpleLoader::SimpleLoader;
        void Before() { BeforeLoad(); }
        void After() { AfterLoad(); }
      } l(*table);
      l.Before();
      std::vector<std::string> row(g["columns"].size());
      for (uint64_t i = 0; i < n; ++i) {
        size_t c = 0;
        for (auto &col : g["columns"]) {
          uint64_t x = seed * 0x100000001B3ULL + (row0 + i) * 16 + c + 0x9E3779B97F4A7C15ULL;
          uint64_t z = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
          z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
          z ^= z >> 31;
          int64_t v = col.value("lo", (int64_t)0) + (int64_t)(z % col.value("range", (uint64_t)1));
          row[c++] = col.value("prefix", std::string()) + std::to_string(v);
        }
        l.Load(row);
      }
      l.After();
    }
    vgpu_ctx *ctx = nullptr;
    vgpu_host::GpuQueryRunner::Bindings bindings;
    if (!stock_only) vgpu_host::check(vgpu_init(job.value("device", 0), &ctx), "vgpu_init");
    else {
      // pre-warm the accessor .so too (SegmentAccess compiles through the reference's Compiler)
      vgpu_host::SegmentAccess access(*table);
    }
    auto run_queries = [&](const char *key) {
    out[key] = json::array();
    for (auto &q : job["queries"]) {
      json res;
      for (int pass = 0; pass < (stock_only ? 1 : 2); ++pass) {
        const char *name = pass == 0 ? "stock" : "gpu";
        try {
          query::MemoryRowOutput output;
          query::QueryFactory factory;
          std::unique_ptr<query::Query> qq(factory.Create(util::Config(q), database));
          if (pass == 0) {
            query::QueryRunner runner(database, output);
            qq->Accept(runner);
            res[name] = {{"rows", output.rows()}, {"stats", stats_json(runner.stats())}};
          } else {
            vgpu_host::GpuQueryRunner runner(database, output, ctx, bindings);
            qq->Accept(runner);
            res[name] = {{"rows", output.rows()}, {"stats", stats_json(runner.stats())}};
          }
        } catch (const std::invalid_argument &e) {
          res[name] = {{"error", e.what()}, {"error_type", "invalid_argument"}};
        } catch (const std::exception &e) {
          res[name] = {{"error", e.what()}, {"error_type", "exception"}};
        }
      }
      out[key].push_back(res);
    }
    };
    run_queries("results");
    // Second ingest batch into the SAME dimension tuples: the reference's upsert updates the metric cells of existing
    // rows in place (src/codegen/db/upsert.cc:386-393), no segment size changes. The integration's one line in
    // input::Loader::AfterLoad — IngestEpoch::Bump() — tells the resident copy it is stale.
    if (job.count("reload_rows")) {
      std::vector<size_t> sizes_before;
      for (auto *sgm : table->store()->segments_copy()) sizes_before.push_back(sgm->size());
      struct L : input::SimpleLoader {
        using input::SimpleLoader::SimpleLoader;
        void Before() { BeforeLoad(); }
        void After() { AfterLoad(); }
      } l(*table);
      l.Before();
      for (auto &r : job["reload_rows"]) {
        std::vector<std::string> row = r.get<std::vector<std::string>>();
        l.Load(row);
      }
      l.After();
      if (job.value("reload_mode", std::string("epoch")) == "mark") {
        // exact notifications, as an upsert hook would give them (here: every row that existed before may have been
        // updated in place); the appended rows are found by the binding itself
        for (size_t si = 0; si < sizes_before.size(); ++si)
          vgpu_host::GpuTableBinding::MarkDirty(table, si, 0, sizes_before[si]);
      } else {
        vgpu_host::IngestEpoch::Bump();
      }
      run_queries("results_after_reload");
      uint64_t partial = 0;
      for (auto &b : bindings) partial += b.second->partial_updates();
      out["partial_updates"] = partial;
    }
    bindings.clear();
    if (ctx) vgpu_shutdown(ctx);
  } catch (const std::exception &e) {
    out["fatal"] = e.what();
    std::cout << out.dump() << std::endl;
    return 1;
  }
  std::cout << out.dump() << std::endl;
  return 0;
}
