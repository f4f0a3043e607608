// adapter_post_harness.cc — TEST INFRASTRUCTURE. The C++ drop-in adapter's host half (viyadb_b200/host/gpu_query_runner.h:
// GpuQueryRunner::PostAggregate with HavingEvaluator, number / time / dictionary formatting, the sort on formatted strings,
// skip / limit) inside a reference process WITHOUT a GPU: the group table a device scan would return is given in the job
// file (tests/test_adapter_post.py takes it from the oracle, which is pinned to the real reference on the same records);
// the reference's own QueryFactory builds the query objects and its own FilterArgsPacker packs the HAVING literals.
// No vgpu_* entry point is called: nothing here computes a scan.
//
// job = {"table": {...}, "dicts": {"<string dim>": ["__exceeded", "v1", ...]}, "rollup_ts": N,
//        "cases": [{"query": {...}, "ngroups": n, "keys": [[bits...] per selected dimension], "accs": [[bits...] per
//                   selected metric], "hidden": [counts...] | null}]}
// bits = the cell widened to 64 bits (two's complement, raw IEEE bits, a count-distinct metric's cardinality)
#include <chrono>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <sstream>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include <nlohmann/json.hpp>
#include "db/database.h"
#include "db/dictionary.h"
#include "db/table.h"
#include "query/output.h"
#include "query/query.h"
#include "query/runner.h"
#include "util/config.h"
#define private public      // reach GpuQueryRunner::PostAggregate (a test harness, not an integration)
#include "gpu_query_runner.h"
#undef private

using json = nlohmann::json;
namespace db = viya::db;
namespace util = viya::util;
namespace query = viya::query;
namespace cg = viya::codegen;

int main(int argc, char **argv) {
  if (argc < 2) {
    std::cerr << "usage: adapter_post_cli <job.json>\n";
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
    dbconf["state_dir"] = job.value("state_dir", std::string("/tmp/viyadb_adapter_post"));
    dbconf["tables"] = json::array({job["table"]});
    db::Database database{util::Config(dbconf)};
    auto *table = database.GetTable(job["table"]["name"].get<std::string>());
    // dictionaries in code order, the way the generated upsert code fills them (code = c2v.size(), both maps): code 0 is
    // "__exceeded" already (dictionary.cc:22-25)
    for (auto *dim : table->dimensions()) {
      if (dim->dim_type() != db::Dimension::DimType::STRING) continue;
      auto dict = static_cast<const db::StrDimension *>(dim)->dict();
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
    vgpu_host::GpuQueryRunner::Bindings bindings;
    out["results"] = json::array();
    for (auto &c : job["cases"]) {
      json res;
      try {
        query::MemoryRowOutput output;
        query::QueryFactory factory;
        std::unique_ptr<query::Query> qq(factory.Create(util::Config(c["query"]), database));
        auto *aq = dynamic_cast<query::AggregateQuery *>(qq.get());
        if (aq == nullptr) throw std::runtime_error("not an aggregate query");
        std::vector<db::AnyNum> hargs;
        if (aq->having() != nullptr) {
          cg::FilterArgsPacker having_args(aq->table());
          aq->having()->Accept(having_args);
          hargs = having_args.args();
        }
        const uint64_t n = c["ngroups"].get<uint64_t>();
        auto &dim_cols = aq->dimension_cols();
        auto &metric_cols = aq->metric_cols();
        std::vector<std::vector<char>> kbuf(dim_cols.size()), abuf(metric_cols.size());
        std::vector<const void *> kptr(dim_cols.size()), aptr(metric_cols.size());
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
        std::vector<uint64_t> hidden;
        if (!c["hidden"].is_null()) hidden = c["hidden"].get<std::vector<uint64_t>>();
        vgpu_result_view view{};
        view.ngroups = n;
        view.nkeys = (uint32_t)dim_cols.size();
        view.nmetrics = (uint32_t)metric_cols.size();
        view.keys = kptr.data();
        view.accs = aptr.data();
        view.hidden_count = hidden.empty() ? nullptr : hidden.data();
        view.aggregated_recs = n;
        vgpu_host::GpuQueryRunner runner(database, output, nullptr, bindings);
        runner.PostAggregate(aq, view, hargs);
        res = {{"rows", output.rows()}, {"output_recs", runner.stats_.output_recs}};
      } catch (const std::invalid_argument &e) {
        res = {{"error", e.what()}, {"error_type", "invalid_argument"}};
      } catch (const std::exception &e) {
        res = {{"error", e.what()}, {"error_type", "exception"}};
      }
      out["results"].push_back(res);
    }
  } catch (const std::exception &e) {
    out["fatal"] = e.what();
    std::cout << out.dump() << std::endl;
    return 1;
  }
  std::cout << out.dump() << std::endl;
  return 0;
}
