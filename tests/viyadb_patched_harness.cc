// viyadb_patched_harness.cc — TEST INFRASTRUCTURE. ViyaDB's own db::Database, compiled from the reference's sources
// with viyadb_b200/host/viyadb_database.patch applied (oracle/Makefile, target patched_db), selected by configuration:
//     {"gpu": true}  ->  Database::Query(conf, output) runs vgpu_host::GpuQueryRunner        (the integration)
//     otherwise      ->  the stock query::QueryRunner, untouched
// Here the C ABI behind it is the mock of tests/mock_vgpu.h (records the plan, answers with the group table of the job
// file, shadows put / update calls), so the whole wiring runs without a GPU: rows go in through Database::Load-style
// ingest (input::SimpleLoader -> the patched Loader::AfterLoad), queries through the real Database::Query.
//
// job = {"table": {...}, "state_dir": "...", "gpu": true|false, "rows": [[...]] (optional, ingested first),
//        "dicts": {...} (optional: fill the dictionaries directly instead of ingesting),
//        "cases": [{"query": {...}, "ngroups": n, "keys": [...], "accs": [...], "hidden": [...] | null, ...}],
//        "reload_rows": [[...]] (optional: a second batch, then every case again), "other_table": {...} +
//        "reload_into": "<name>" (optional: the second batch goes into another table of the same database)}
#include "db/database.h"
#include "db/dictionary.h"
#include "db/table.h"
#include "gpu_query_runner.h"
#include "input/simple.h"
#include "query/output.h"
#include "util/config.h"
#include <cstring>
#include <fstream>
#include <iostream>
#include <nlohmann/json.hpp>
#include "mock_vgpu.h"   // the mock device: defines the C ABI of include/vgpu.h

namespace db = viya::db;
namespace util = viya::util;
namespace query = viya::query;

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

// the canned aggregate result of one case, kept alive while the query runs
struct Canned {
  std::vector<std::vector<char>> kbuf, abuf;
  std::vector<const void *> kptr, aptr;
  std::vector<uint64_t> hidden;
  void set(const json &c, db::Table *table) {
    const uint64_t n = c["ngroups"].get<uint64_t>();
    // element widths of the selected columns (vgpu.h: a count-distinct metric comes back as its cardinality, uint64)
    auto width_of = [&](const std::string &name, bool metric) -> uint32_t {
      auto *col = table->column(name);
      if (metric && static_cast<const db::Metric *>(col)->agg_type() == db::Metric::AggregationType::BITSET) return 8;
      return (uint32_t)col->num_type().size();
    };
    auto fill = [&](const json &arrs, const json &names, bool metric, std::vector<std::vector<char>> &buf, std::vector<const void *> &ptr) {
      buf.assign(arrs.size(), {});
      ptr.assign(arrs.size(), nullptr);
      for (size_t k = 0; k < arrs.size(); ++k) {
        const uint32_t w = width_of(names[k].get<std::string>(), metric);
        buf[k].resize(n * w + 8);
        for (uint64_t g = 0; g < n; ++g) {
          uint64_t bits = arrs[k][g].get<uint64_t>();
          std::memcpy(buf[k].data() + g * w, &bits, w);
        }
        ptr[k] = buf[k].data();
      }
    };
    fill(c["keys"], c["key_names"], false, kbuf, kptr);
    fill(c["accs"], c["acc_names"], true, abuf, aptr);
    hidden.clear();
    if (!c["hidden"].is_null()) hidden = c["hidden"].get<std::vector<uint64_t>>();
    g_canned = vgpu_result_view{};
    g_canned.ngroups = n;
    g_canned.nkeys = (uint32_t)kptr.size();
    g_canned.nmetrics = (uint32_t)aptr.size();
    g_canned.keys = kptr.data();
    g_canned.accs = aptr.data();
    g_canned.hidden_count = hidden.empty() ? nullptr : hidden.data();
    g_canned.aggregated_recs = n;
    g_canned.scanned_recs = c.value("scanned_recs", (uint64_t)0);
    g_canned.scanned_segments = c.value("scanned_segments", (uint64_t)0);
  }
};
}  // namespace

int main(int argc, char **argv) {
  if (argc < 2) {
    std::cerr << "usage: viyadb_patched_cli <job.json>\n";
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
    if (job.count("other_table")) dbconf["tables"].push_back(job["other_table"]);   // a second table of the same database
    if (job.value("gpu", false)) dbconf["gpu"] = true;      // <- the configuration switch of the patch
    db::Database database{util::Config(dbconf)};
    auto *table = database.GetTable(job["table"]["name"].get<std::string>());
    if (job.count("rows")) load_rows(table, job["rows"]);
    if (job.count("dicts")) {
      for (auto *dim : table->dimensions()) {
        if (dim->dim_type() != db::Dimension::DimType::STRING) continue;
        auto dict = static_cast<const db::StrDimension *>(dim)->dict();
        auto &vals = job["dicts"][dim->name()];
        for (size_t i = dict->c2v().size(); i < vals.size(); ++i) {
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
    }
    auto run_cases = [&](const char *key) {
      out[key] = json::array();
      for (auto &c : job["cases"]) {
        json res;
        try {
          Canned canned;
          if (c.count("ngroups")) canned.set(c, table);
          g_plan = json();
          if (g_table != nullptr) g_table->calls = json::array();
          query::MemoryRowOutput output;
          auto stats = database.Query(util::Config(c["query"]), output);     // the real entry point
          json shadow;   // after the query's Sync(): is what the device holds the live store?
          if (g_table != nullptr) {
            vgpu_host::SegmentAccess access(*table);
            shadow = compare_with_live(table, access);
          }
          res = {{"rows", output.rows()}, {"plan", g_plan}, {"device_calls", g_table ? g_table->calls : json::array()},
                 {"shadow", shadow},
                 {"stats", {{"scanned_segments", stats.scanned_segments}, {"scanned_recs", stats.scanned_recs},
                            {"aggregated_recs", stats.aggregated_recs}, {"output_recs", stats.output_recs}}}};
        } catch (const std::invalid_argument &e) {
          res = {{"error", e.what()}, {"error_type", "invalid_argument"}};
        } catch (const std::exception &e) {
          res = {{"error", e.what()}, {"error_type", "exception"}};
        }
        out[key].push_back(res);
      }
    };
    run_cases("results");
    if (job.count("reload_rows")) {
      // ends in the patched Loader::AfterLoad -> IngestEpoch::Bump(&table_): the batch may go into the OTHER table
      auto *into = job.count("reload_into") ? database.GetTable(job["reload_into"].get<std::string>()) : table;
      load_rows(into, job["reload_rows"]);
      run_cases("results_after_reload");
    }
    out["mock_device_used"] = g_table != nullptr;
  } catch (const std::exception &e) {
    out["fatal"] = e.what();
    std::cout << out.dump() << std::endl;
    return 1;
  }
  std::cout << out.dump() << std::endl;
  return 0;
}
