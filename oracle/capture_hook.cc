// oracle/capture_hook.cc — TEST INFRASTRUCTURE (golden-vector capture), not product code.
//
// Linked into a second build of the reference's own gtest binary (oracle/Makefile: ref_capture)
// with  -Wl,--wrap=<Database::Database(Config const&)>  -Wl,--wrap=<Database::Query(...)>.
// Every `db.Query(json, output)` the reference's tests issue is executed by the UNMODIFIED
// reference (the __real_ symbol) and recorded: the database config, the query JSON, the pinned
// rollup clock, the rows the reference produced (or the exception it threw), its QueryStats
// counters, and a dump of the queried table's real segments (VGPUSEG1, same writer as oracle_cli).
// tests/golden/make_golden.py turns the capture into the committed fixtures.
//
// Output directory: $VIYA_CAPTURE_DIR (default ./capture): capture.jsonl + seg_<n>.bin
#include "db/database.h"
#include "db/table.h"
#include "query/output.h"
#include "query/stats.h"
#include "util/config.h"
#include <cstdlib>
#include <fstream>
#include <gtest/gtest.h>
#include <map>
#include <mutex>
#include <nlohmann/json.hpp>
#include <string>

#define VGPU_DUMP_ONLY 1
#include "oracle_cli.cc"  // dump_table() — same VGPUSEG1 writer

namespace {
std::mutex g_mu;
std::map<const void *, std::string> g_db_conf;
int g_seq = 0;

std::string capture_dir() {
  const char *d = getenv("VIYA_CAPTURE_DIR");
  return d ? d : "capture";
}
}  // namespace

extern "C" {

void real_db_ctor(viya::db::Database *self, const viya::util::Config &conf)
    __asm__("__real__ZN4viya2db8DatabaseC1ERKNS_4util6ConfigE");
void wrap_db_ctor(viya::db::Database *self, const viya::util::Config &conf)
    __asm__("__wrap__ZN4viya2db8DatabaseC1ERKNS_4util6ConfigE");

void wrap_db_ctor(viya::db::Database *self, const viya::util::Config &conf) {
  {
    std::lock_guard<std::mutex> lk(g_mu);
    g_db_conf[self] = conf.dump();
  }
  real_db_ctor(self, conf);
}

viya::query::QueryStats real_db_query(viya::db::Database *self, const viya::util::Config &q,
                                      viya::query::RowOutput &out)
    __asm__("__real__ZN4viya2db8Database5QueryERKNS_4util6ConfigERNS_5query9RowOutputE");
viya::query::QueryStats wrap_db_query(viya::db::Database *self, const viya::util::Config &q,
                                      viya::query::RowOutput &out)
    __asm__("__wrap__ZN4viya2db8Database5QueryERKNS_4util6ConfigERNS_5query9RowOutputE");

viya::query::QueryStats wrap_db_query(viya::db::Database *self, const viya::util::Config &q,
                                      viya::query::RowOutput &out) {
  json rec;
  int seq;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    seq = g_seq++;
    rec["db"] = json::parse(g_db_conf.count(self) ? g_db_conf[self] : std::string("{}"));
  }
  auto *info = ::testing::UnitTest::GetInstance()->current_test_info();
  rec["test"] = info ? std::string(info->test_case_name()) + "." + info->name() : std::string("?");
  rec["seq"] = seq;
  rec["query"] = json::parse(q.dump());
  const char *ts = getenv("VIYA_TEST_ROLLUP_TS");
  if (ts) rec["rollup_ts"] = std::string(ts);
  std::string dir = capture_dir();
  std::string type = q.exists("type") ? q.str("type") : "";
  if ((type == "aggregate" || type == "select" || type == "search") && q.exists("table")) {
    try {
      auto *table = self->GetTable(q.str("table"));
      std::string path = dir + "/seg_" + std::to_string(seq) + ".bin";
      dump_table(*table, path);
      rec["dump"] = "seg_" + std::to_string(seq) + ".bin";
    } catch (const std::exception &e) {
      rec["dump_error"] = e.what();
    }
  }
  auto *mem = dynamic_cast<viya::query::MemoryRowOutput *>(&out);
  size_t before = mem ? mem->rows().size() : 0;
  auto flush = [&]() {
    std::lock_guard<std::mutex> lk(g_mu);
    std::ofstream f(dir + "/capture.jsonl", std::ios::app);
    f << rec.dump() << "\n";
  };
  try {
    viya::query::QueryStats stats = real_db_query(self, q, out);
    if (mem) {
      json rows = json::array();
      for (size_t i = before; i < mem->rows().size(); ++i) rows.push_back(mem->rows()[i]);
      rec["rows"] = rows;
    }
    rec["stats"] = {{"scanned_segments", stats.scanned_segments},
                    {"scanned_recs", stats.scanned_recs},
                    {"aggregated_recs", stats.aggregated_recs},
                    {"output_recs", stats.output_recs}};
    flush();
    return stats;
  } catch (const std::invalid_argument &e) {
    rec["error"] = e.what();
    rec["error_type"] = "invalid_argument";
    flush();
    throw;
  } catch (const std::exception &e) {
    rec["error"] = e.what();
    rec["error_type"] = "exception";
    flush();
    throw;
  }
}

}  // extern "C"
