// oracle/oracle_cli.cc — TEST INFRASTRUCTURE (the oracle driver), not product code.
//
// Drives the UNMODIFIED reference implementation (compiled in place from /root/reference by
// oracle/Makefile) through its own public API — db::Database, input::SimpleLoader and
// Database::Query (src/db/database.cc:104-111), i.e. the g++-JIT path of
// src/codegen/query/{scan,filter,post_agg}.cc — and prints what it produced. Used to
//   * generate the golden fixtures under tests/golden/ (tests/golden/make_golden.py),
//   * dump the reference's real segment columns so the CUDA path can scan identical bytes,
//   * time the reference CPU path for bench.py's `cpu_baseline` / `--impl reference`.
//
// usage: oracle_cli <job.json>      (JSON result on stdout)
//
// job = {
//   "state_dir": "...",                 JIT .so cache (src/codegen/compiler.cc:93-94)
//   "rollup_ts": 1496570140,            optional; exported as VIYA_TEST_ROLLUP_TS (rollup.cc:47-49)
//   "table": { ...reference table config (src/db/table.cc:47-96)... },
//   "rows": [[...strings...], ...]      explicit rows, fed through input::SimpleLoader, and/or
//   "generate": {"n": N, "seed": S, "row_offset": R0, "columns": [ {"prefix": "a", "lo": 1, "range": 1000} ...]},
//                                       synthetic rows; one entry per INPUT column (dims, then
//                                       non-count metrics); value = lo + splitmix64(seed*K + row*16 + col) % range
//                                       (or {"mode":"div","div":D,...} → lo + (row / D) % range)
//   "dump": "path",                     optional: write the table's segments (VGPUSEG1 container)
//   "queries": [ {...reference query JSON...}, ... ],
//   "repeat": R                         run each query R times, report every timing
// }
#include "db/database.h"
#include "db/dictionary.h"
#include "db/store.h"
#include "db/table.h"
#include "input/simple.h"
#include "query/output.h"
#include "query/stats.h"
#include "util/config.h"
#include "../viyadb_b200/host/segment_access.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <libgen.h>
#include <limits.h>
#include <nlohmann/json.hpp>
#include <sstream>
#include <unistd.h>

using json = nlohmann::json;
namespace db = viya::db;
namespace util = viya::util;
namespace query = viya::query;
namespace input = viya::input;

static inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  uint64_t z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

static const char *type_code(const db::BaseNumType &t, bool is_value_metric_or_numdim) {
  if (is_value_metric_or_numdim) {
    auto &nt = static_cast<const db::NumericType &>(t);
    switch (nt.type()) {
    case db::NumericType::BYTE: return "i8";
    case db::NumericType::UBYTE: return "u8";
    case db::NumericType::SHORT: return "i16";
    case db::NumericType::USHORT: return "u16";
    case db::NumericType::INT: return "i32";
    case db::NumericType::UINT: return "u32";
    case db::NumericType::LONG: return "i64";
    case db::NumericType::ULONG: return "u64";
    case db::NumericType::FLOAT: return "f32";
    case db::NumericType::DOUBLE: return "f64";
    }
  }
  switch (t.size()) {
  case db::BaseNumType::_1: return "u8";
  case db::BaseNumType::_2: return "u16";
  case db::BaseNumType::_4: return "u32";
  default: return "u64";
  }
}

static void dump_table(db::Table &table, const std::string &path) {
  vgpu_host::SegmentAccess access(table);
  json hdr;
  hdr["segment_size"] = table.segment_size();
  hdr["dims"] = json::array();
  hdr["metrics"] = json::array();
  hdr["dicts"] = json::object();
  for (auto *dim : table.dimensions()) {
    json d;
    d["name"] = dim->name();
    const char *kind = "string";
    bool numdim = false;
    switch (dim->dim_type()) {
    case db::Dimension::STRING: kind = "string"; break;
    case db::Dimension::NUMERIC: kind = "numeric"; numdim = true; break;
    case db::Dimension::TIME:
      kind = static_cast<const db::TimeDimension *>(dim)->micro_precision() ? "microtime" : "time";
      break;
    case db::Dimension::BOOLEAN: kind = "boolean"; break;
    }
    d["kind"] = kind;
    d["type"] = type_code(dim->num_type(), numdim);
    hdr["dims"].push_back(d);
    if (dim->dim_type() == db::Dimension::STRING) {
      auto dict = static_cast<const db::StrDimension *>(dim)->dict();
      hdr["dicts"][dim->name()] = dict->c2v();
    }
  }
  static const char *agg_names[] = {"max", "min", "sum", "avg", "count", "bitset"};
  for (auto *m : table.metrics()) {
    json j;
    j["name"] = m->name();
    j["agg"] = agg_names[m->agg_type()];
    j["type"] = type_code(m->num_type(), m->agg_type() != db::Metric::BITSET);
    hdr["metrics"].push_back(j);
  }
  hdr["hidden_count"] = access.has_hidden_count();

  auto segments = table.store()->segments_copy();
  hdr["segments"] = json::array();
  size_t nd = table.dimensions().size(), nm = table.metrics().size();
  std::vector<const void *> dims(nd), metrics(nm);
  std::vector<uint64_t> stats(2 * nd + 2);
  std::string blob;
  auto append = [&blob](const void *p, size_t n) {
    blob.append(static_cast<const char *>(p), n);
    while (blob.size() % 8) blob.push_back('\0');
  };
  for (auto *s : segments) {
    size_t size = s->size();
    const void *hidden = nullptr;
    access.columns()(s, dims.data(), metrics.data(), &hidden, stats.data());
    json sj;
    sj["size"] = size;
    sj["stats"] = json::array();
    for (auto *dim : table.dimensions()) {
      if (dim->dim_type() == db::Dimension::NUMERIC || dim->dim_type() == db::Dimension::TIME) {
        sj["stats"].push_back({{"dim", dim->index()},
                               {"max_raw", stats[2 * dim->index()]},
                               {"min_raw", stats[2 * dim->index() + 1]}});
      }
    }
    sj["cols"] = json::array();
    for (auto *dim : table.dimensions()) {
      sj["cols"].push_back({{"off", blob.size()}, {"bytes", size * dim->num_type().size()}});
      append(dims[dim->index()], size * dim->num_type().size());
    }
    for (auto *m : table.metrics()) {
      if (m->agg_type() == db::Metric::BITSET) {
        std::vector<uint64_t> offsets(size + 1);
        uint64_t total = access.bitset()(s, m->index(), size, offsets.data(), nullptr);
        std::vector<uint64_t> values(total + 1);
        access.bitset()(s, m->index(), size, offsets.data(), values.data());
        size_t off0 = blob.size();
        append(offsets.data(), (size + 1) * 8);
        size_t off1 = blob.size();
        append(values.data(), total * 8);
        sj["cols"].push_back({{"off", off0}, {"bytes", (size + 1) * 8}, {"values_off", off1},
                              {"values", total}});
      } else {
        sj["cols"].push_back({{"off", blob.size()}, {"bytes", size * m->num_type().size()}});
        append(metrics[m->index()], size * m->num_type().size());
      }
    }
    if (hidden != nullptr) {
      sj["hidden_count"] = {{"off", blob.size()}, {"bytes", size * 8}};
      append(hidden, size * 8);
    }
    hdr["segments"].push_back(sj);
  }
  std::string h = hdr.dump();
  std::ofstream out(path, std::ios::binary);
  out.write("VGPUSEG1", 8);
  uint64_t hl = h.size();
  out.write(reinterpret_cast<const char *>(&hl), 8);
  out.write(h.data(), h.size());
  size_t pad = (8 - (16 + h.size()) % 8) % 8;
  out.write("\0\0\0\0\0\0\0\0", pad);
  out.write(blob.data(), blob.size());
}

#ifndef VGPU_DUMP_ONLY
int main(int argc, char **argv) {
  if (argc < 2) {
    std::cerr << "usage: oracle_cli <job.json>\n";
    return 2;
  }
  std::ifstream in(argv[1]);
  if (!in) {
    std::cerr << "cannot open " << argv[1] << "\n";
    return 2;
  }
  json job;
  in >> job;

  auto abspath = [](const std::string &p) {
    if (!p.empty() && p[0] == '/') return p;
    char cwd[PATH_MAX];
    if (getcwd(cwd, sizeof(cwd)) == nullptr) return p;
    return std::string(cwd) + "/" + p;
  };
  std::string dump_path = job.count("dump") ? abspath(job["dump"].get<std::string>()) : "";
  std::string state_dir = abspath(job.value("state_dir", std::string("/tmp/viyadb_oracle")));

  // The JIT's include/library flags are relative to CWD (compiler.cc:46-54): run from <exe dir>/root/build.
  {
    char exe[PATH_MAX];
    ssize_t n = readlink("/proc/self/exe", exe, sizeof(exe) - 1);
    if (n > 0) {
      exe[n] = '\0';
      std::string dir = dirname(exe);
      std::string build = dir + "/root/build";
      if (chdir(build.c_str()) != 0) {
        std::cerr << "cannot chdir to " << build << "\n";
        return 2;
      }
    }
  }
  if (job.count("rollup_ts")) {
    std::string v = std::to_string(job["rollup_ts"].get<long>()) + "L";
    setenv("VIYA_TEST_ROLLUP_TS", v.c_str(), 1);
  }

  json out;
  try {
    json dbconf;
    dbconf["state_dir"] = state_dir;
    dbconf["tables"] = json::array({job["table"]});
    db::Database database{util::Config(dbconf)};
    auto *table = database.GetTable(job["table"]["name"].get<std::string>());

    auto t0 = std::chrono::steady_clock::now();
    size_t loaded = 0;
    if (job.count("rows")) {
      // SimpleLoader::Load(initializer_list) = BeforeLoad; Load(row)...; AfterLoad (input/simple.cc:29-35);
      // the per-row overload is the same upsert call without the Before/After hooks, which only reset
      // per-batch stats and rollup boundaries — we go through a derived class to call them.
      struct L : input::SimpleLoader {
        using input::SimpleLoader::SimpleLoader;
        void Before() { BeforeLoad(); }
        void After() { AfterLoad(); }
      } l(*table);
      l.Before();
      for (auto &r : job["rows"]) {
        std::vector<std::string> row = r.get<std::vector<std::string>>();
        l.Load(row);
        ++loaded;
      }
      l.After();
    }
    if (job.count("generate")) {
      auto &g = job["generate"];
      uint64_t n = g["n"].get<uint64_t>();
      uint64_t seed = g.value("seed", (uint64_t)42);
      uint64_t row0 = g.value("row_offset", (uint64_t)0);
      struct ColGen { std::string prefix; int64_t lo; uint64_t range; int mode; uint64_t div; };
      std::vector<ColGen> cols;
      for (auto &c : g["columns"]) {
        ColGen cg;
        cg.prefix = c.value("prefix", std::string());
        cg.lo = c.value("lo", (int64_t)0);
        cg.range = c.value("range", (uint64_t)1);
        std::string mode = c.value("mode", std::string("hash"));
        cg.mode = mode == "div" ? 1 : 0;
        cg.div = c.value("div", (uint64_t)1);
        cols.push_back(cg);
      }
      struct L : input::SimpleLoader {
        using input::SimpleLoader::SimpleLoader;
        void Before() { BeforeLoad(); }
        void After() { AfterLoad(); }
      } l(*table);
      l.Before();
      std::vector<std::string> row(cols.size());
      for (uint64_t i = 0; i < n; ++i) {
        uint64_t r = row0 + i;
        for (size_t c = 0; c < cols.size(); ++c) {
          uint64_t u = cols[c].mode == 1 ? (r / cols[c].div)
                                         : splitmix64(seed * 0x100000001B3ULL + r * 16 + c);
          int64_t v = cols[c].lo + (int64_t)(u % cols[c].range);
          row[c] = cols[c].prefix + std::to_string(v);
        }
        l.Load(row);
        ++loaded;
      }
      l.After();
    }
    auto t1 = std::chrono::steady_clock::now();
    out["loaded_rows"] = loaded;
    out["load_ms"] = std::chrono::duration<double, std::milli>(t1 - t0).count();
    {
      size_t stored = 0;
      auto segs = table->store()->segments_copy();
      for (auto *s : segs) stored += s->size();
      out["stored_rows"] = stored;
      out["segments"] = segs.size();
    }

    if (!dump_path.empty()) dump_table(*table, dump_path);

    out["results"] = json::array();
    int repeat = job.value("repeat", 1);
    if (job.count("queries")) {
      for (auto &q : job["queries"]) {
        json res;
        try {
          json timings = json::array();
          for (int it = 0; it < repeat; ++it) {
            query::MemoryRowOutput output;
            auto stats = database.Query(util::Config(q), output);
            if (it == repeat - 1) {
              res["rows"] = output.rows();
              res["stats"] = {{"scanned_segments", stats.scanned_segments},
                              {"scanned_recs", stats.scanned_recs},
                              {"aggregated_recs", stats.aggregated_recs},
                              {"output_recs", stats.output_recs}};
            }
            timings.push_back({{"whole_ms", stats.whole_time.count() * 1e3},
                               {"compile_ms", stats.compile_time.count() * 1e3}});
          }
          res["timings"] = timings;
        } catch (const std::exception &e) {
          res["error"] = e.what();
        }
        out["results"].push_back(res);
      }
    }
  } catch (const std::exception &e) {
    out["fatal"] = e.what();
    std::cout << out.dump() << std::endl;
    return 1;
  }
  std::cout << out.dump() << std::endl;
  return 0;
}
#endif  // VGPU_DUMP_ONLY
