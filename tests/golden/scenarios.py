"""Scenario definitions shared by tests/golden/make_golden.py (-> oracle_cli, the real reference) and
tests/test_dropin_cpp.py (-> vgpu_cli: stock runner vs GpuQueryRunner in one reference process).

  * fixtures of the reference's own tests (test/db.h:33-289) with extra queries that the gtests do not
    issue (edge cases: NOT IN pruning quirk, float MAX init, wrap-around sums, missing dict values),
  * small-N twins of the benchmark configs C0-C4 (same generators as bench.py).
"""
NOW = 1496570140
T0 = 1490000000


def agg(**kw):
    q = {"type": "aggregate", "table": "events"}
    q.update(kw)
    return q


INAPP = {"name": "events",
         "dimensions": [{"name": "country"}, {"name": "event_name", "length": 20}, {"name": "install_time", "type": "uint"}],
         "metrics": [{"name": "count", "type": "count"}, {"name": "revenue", "type": "double_sum"}]}
INAPP_ROWS = [["US", "purchase", "20141112", "0.1"], ["US", "purchase", "20141113", "1.1"], ["US", "donate", "20141112", "5.0"],
              ["IL", "refund", "20141111", "1.01"], ["CH", "refund", "20141111", "1.1"], ["AZ", "refund", "20141111", "1.1"],
              ["RU", "donate", "20141112", "1.0"], ["KZ", "review", "20141113", "5.0"]]

SCENARIOS = [
    {"name": "inapp", "table": INAPP, "rows": INAPP_ROWS, "queries": [
        agg(dimensions=["event_name", "country"], metrics=["revenue"], filter={"op": "eq", "column": "country", "value": "US"}),
        agg(dimensions=["country"], metrics=["count", "revenue"], header=True),
        agg(dimensions=["country"], metrics=["count"], filter={"op": "eq", "column": "country", "value": "nowhere"}),
        agg(dimensions=["country"], metrics=["count"], filter={"op": "ne", "column": "country", "value": "nowhere"}),
        agg(dimensions=[], metrics=["count", "revenue"], filter={"op": "gt", "column": "install_time", "value": "20141111"}),
        agg(dimensions=["event_name"], metrics=["count"],
            filter={"op": "not", "filter": {"op": "or", "filters": [
                {"op": "in", "column": "country", "values": ["US", "IL"]},
                {"op": "lt", "column": "revenue", "value": "1.05"}]}}),
        agg(dimensions=["country"], metrics=["revenue", "count"], having={"op": "ge", "column": "revenue", "value": "1.1"},
            sort=[{"column": "revenue"}, {"column": "country", "ascending": True}], limit=4),
        agg(select=[{"column": "install_time"}, {"column": "count"}], sort=[{"column": "install_time", "ascending": True}], skip=1, limit=2),
    ]},
    {"name": "prune_quirk", "table": {"name": "events", "segment_size": 2,
                                      "dimensions": [{"name": "n", "type": "uint"}, {"name": "s"}],
                                      "metrics": [{"name": "count", "type": "count"}]},
     "rows": [["1", "a"], ["2", "b"], ["10", "a"], ["11", "c"], ["20", "a"]],
     "queries": [
         agg(dimensions=["n"], metrics=["count"], filter={"op": "not", "filter": {"op": "in", "column": "n", "values": ["1"]}}),
         agg(dimensions=["n"], metrics=["count"], filter={"op": "in", "column": "n", "values": ["2", "11", "99"]}),
         agg(dimensions=["s"], metrics=["count"], filter={"op": "ge", "column": "n", "value": "10"}),
         agg(dimensions=["s"], metrics=["count"], filter={"op": "or", "filters": [
             {"op": "lt", "column": "n", "value": "2"}, {"op": "eq", "column": "s", "value": "c"}]}),
         agg(dimensions=["s"], metrics=["count"], filter={"op": "and", "filters": [
             {"op": "le", "column": "n", "value": "0"}, {"op": "eq", "column": "s", "value": "a"}]}),
     ]},
    {"name": "metric_types", "table": {"name": "events", "dimensions": [{"name": "k"}],
                                       "metrics": [{"name": "bs", "type": "byte_sum"}, {"name": "ubs", "type": "ubyte_sum"},
                                                   {"name": "is", "type": "int_sum"}, {"name": "fmx", "type": "float_max"},
                                                   {"name": "fmn", "type": "float_min"}, {"name": "dmx", "type": "double_max"},
                                                   {"name": "lav", "type": "long_avg"}, {"name": "smn", "type": "short_min"},
                                                   {"name": "umx", "type": "uint_max"}]},
     "rows": [["a", "100", "200", "2000000000", "-1.5", "-1.5", "-7.25", "10", "-5", "7"],
              ["a", "100", "100", "2000000000", "-2.5", "2.5", "-0.5", "21", "3", "4000000000"],
              ["b", "-128", "255", "-5", "3.25", "0", "1e300", "-9", "-32768", "0"],
              ["b", "-1", "1", "5", "1.0", "-0.0", "2.0", "0", "32767", "1"]],
     "queries": [
         agg(dimensions=["k"], metrics=["bs", "ubs", "is", "fmx", "fmn", "dmx", "lav", "smn", "umx"]),
         agg(dimensions=[], metrics=["lav", "is"], filter={"op": "gt", "column": "lav", "value": "-100"}),
         agg(dimensions=["k"], metrics=["lav"], having={"op": "gt", "column": "lav", "value": "0"}),
     ]},
    {"name": "users_bitset", "table": {"name": "events", "dimensions": [{"name": "country"}, {"name": "event_name"}, {"name": "time", "type": "uint"}],
                                       "metrics": [{"name": "user_id", "type": "bitset"}]},
     "rows": [["US", "purchase", "1495475514", "12345"], ["RU", "support", "1495475517", "12346"],
              ["US", "openapp", "1495475632", "12347"], ["IL", "purchase", "1495475715", "12348"],
              ["KZ", "closeapp", "1495475716", "12349"], ["US", "uninstall", "1495475809", "12350"],
              ["KZ", "purchase", "1495475808", "12351"], ["US", "purchase", "1495476000", "12352"],
              ["US", "purchase", "1495475514", "12352"], ["US", "purchase", "1495475514", "777"]],
     "queries": [
         agg(dimensions=["country"], metrics=["user_id"]),
         agg(dimensions=["event_name"], metrics=["user_id"], filter={"op": "ge", "column": "user_id", "value": "1"}),
         agg(dimensions=["country", "event_name"], metrics=["user_id"], filter={"op": "gt", "column": "user_id", "value": "1"}),
         agg(dimensions=[], metrics=["user_id"], having={"op": "gt", "column": "user_id", "value": "3"}),
     ]},
    {"name": "time_rollup", "rollup_ts": NOW,
     "table": {"name": "events", "dimensions": [{"name": "country"}, {"name": "install_time", "type": "time", "format": "millis",
                                                 "rollup_rules": [{"granularity": "hour", "after": "1 days"},
                                                                  {"granularity": "day", "after": "1 weeks"},
                                                                  {"granularity": "month", "after": "1 years"}]},
                                                {"name": "mt", "type": "microtime", "format": "micros"}],
               "metrics": [{"name": "count", "type": "count"}]},
     "rows": [["US", str((NOW - d * 3600 * 7 - 11) * 1000), str((NOW - d * 86400 * 5 - 3) * 1000000 + 123456)] for d in range(120)],
     "queries": [
         agg(select=[{"column": "install_time", "format": "%Y-%m-%d %H:%M:%S"}, {"column": "count"}]),
         agg(select=[{"column": "install_time", "granularity": "day", "format": "%Y-%m-%d"}, {"column": "count"}]),
         agg(select=[{"column": "install_time", "granularity": "year", "format": "%Y-%m-%d %H:%M:%S"}, {"column": "count"}]),
         agg(select=[{"column": "mt", "granularity": "month", "format": "%Y-%m-%d %H:%M:%S"}, {"column": "count"}]),
         agg(select=[{"column": "mt", "format": "%Y-%m-%d %H:%M:%S"}, {"column": "count"}], filter={"op": "ge", "column": "mt", "value": str((NOW - 86400 * 30) * 1000000)}),
         agg(select=[{"column": "install_time", "granularity": "minute", "format": "%Y-%m-%d %H:%M:%S"}, {"column": "count"}],
             filter={"op": "gt", "column": "install_time", "value": "2017-06-01 00:00:00"}),
     ]},
    # ---- small-N twins of the benchmark configs (bench.py WORKLOADS) ----
    {"name": "c1_twin", "table": {"name": "events", "segment_size": 20000,
                                  "dimensions": [{"name": "d0"}, {"name": "d1"}, {"name": "d2"}, {"name": "d3"}],
                                  "metrics": [{"name": "count", "type": "count"}, {"name": "m1", "type": "long_sum"}, {"name": "m2", "type": "int_sum"}]},
     "generate": {"n": 70000, "seed": 42, "columns": [{"prefix": "a", "lo": 1, "range": 16}, {"prefix": "b", "lo": 1, "range": 50},
                                                       {"prefix": "c", "lo": 1, "range": 10}, {"prefix": "d", "lo": 1, "range": 1000000},
                                                       {"lo": 0, "range": 1000}, {"lo": -50, "range": 100}]},
     "queries": [agg(dimensions=["d1", "d2"], metrics=["m1", "count"], filter={"op": "eq", "column": "d0", "value": "a7"}),
                 agg(dimensions=["d0"], metrics=["m1", "m2", "count"])]},
    {"name": "c2_twin", "table": {"name": "events", "segment_size": 20000,
                                  "dimensions": [{"name": "d0"}, {"name": "d1"}, {"name": "d2"}, {"name": "d3"},
                                                 {"name": "n4", "type": "ushort"}, {"name": "t5", "type": "time", "format": "posix"}],
                                  "metrics": [{"name": "mn", "type": "int_min"}, {"name": "mx", "type": "int_max"}, {"name": "uid", "type": "bitset"}]},
     "generate": {"n": 70000, "seed": 42, "columns": [{"prefix": "a", "lo": 1, "range": 50}, {"prefix": "b", "lo": 1, "range": 20},
                                                       {"prefix": "c", "lo": 1, "range": 50}, {"prefix": "d", "lo": 1, "range": 100},
                                                       {"lo": 0, "range": 1000}, {"lo": T0, "range": 10000000},
                                                       {"lo": -2**31, "range": 2**32}, {"lo": -2**31, "range": 2**32}, {"lo": 0, "range": 1000}]},
     "queries": [agg(dimensions=["d0", "d1", "d2", "d3"], metrics=["mn", "mx", "uid"],
                     filter={"op": "and", "filters": [
                         {"op": "in", "column": "d0", "values": ["a3", "a11", "a19", "a27", "a42"]},
                         {"op": "ge", "column": "n4", "value": "250"}, {"op": "lt", "column": "n4", "value": "750"},
                         {"op": "ge", "column": "t5", "value": str(T0 + 2500000)}, {"op": "lt", "column": "t5", "value": str(T0 + 7500000)}]}),
                 agg(dimensions=["d1"], metrics=["uid", "mn"], filter={"op": "in", "column": "d0", "values": ["a3", "a11"]})]},
    {"name": "c3_twin", "table": {"name": "events", "segment_size": 20000,
                                  "dimensions": [{"name": "d0"}, {"name": "d1"}, {"name": "d2"}, {"name": "t3", "type": "time", "format": "posix"}],
                                  "metrics": [{"name": "m1", "type": "long_sum"}]},
     "generate": {"n": 70000, "seed": 42, "columns": [{"prefix": "a", "lo": 1, "range": 100}, {"prefix": "b", "lo": 1, "range": 100},
                                                       {"prefix": "c", "lo": 1, "range": 100}, {"lo": T0, "range": 4000000}, {"lo": 0, "range": 1000}]},
     "queries": [agg(dimensions=["d0", "d1", "d2"], metrics=["m1"],
                     filter={"op": "and", "filters": [{"op": "ge", "column": "t3", "value": str(T0 + 1000000)},
                                                      {"op": "lt", "column": "t3", "value": str(T0 + 2000000)}]})]},
    {"name": "c4_twin", "rollup_ts": NOW,
     "table": {"name": "events", "segment_size": 20000,
               "dimensions": [{"name": "d0"}, {"name": "t1", "type": "time", "format": "posix",
                                               "rollup_rules": [{"granularity": "hour", "after": "1 days"},
                                                                {"granularity": "day", "after": "1 weeks"},
                                                                {"granularity": "month", "after": "1 years"}]}],
               "metrics": [{"name": "m1", "type": "long_sum"}, {"name": "count", "type": "count"}]},
     "generate": {"n": 70000, "seed": 42, "columns": [{"prefix": "a", "lo": 1, "range": 200},
                                                       {"lo": NOW - 730 * 86400, "range": 730 * 86400}, {"lo": 0, "range": 1000}]},
     "queries": [agg(select=[{"column": "d0"}, {"column": "t1", "granularity": "hour", "format": "%Y-%m-%d %H:%M:%S"}, {"column": "m1"}, {"column": "count"}]),
                 agg(select=[{"column": "t1", "format": "%Y-%m-%d %H:%M:%S"}, {"column": "count"}])]},
]


# ---- select / search queries: oracle groundwork, pinned on CPU by tests/test_oracle_select.py.
# Not part of SCENARIOS: the drop-in GPU tests do not run them yet (GpuQueryRunner delegates them to the stock runner).
def sel(**kw):
    q = {"type": "select", "table": "events"}
    q.update(kw)
    return q


def search(**kw):
    q = {"type": "search", "table": "events"}
    q.update(kw)
    return q


def _table(name):
    return next(sc for sc in SCENARIOS if sc["name"] == name)


SELECT_SCENARIOS = [
    {"name": "select_inapp", "table": INAPP, "rows": INAPP_ROWS, "queries": [
        sel(dimensions=["event_name", "country"], metrics=["revenue"], filter={"op": "eq", "column": "country", "value": "US"}),
        sel(dimensions=["country"], metrics=["revenue", "count"], header=True),
        sel(dimensions=["event_name", "country"], metrics=["revenue"], limit=2),
        sel(dimensions=["event_name", "country"], metrics=["revenue"], skip=6),
        sel(select=[{"column": "*"}], filter={"op": "gt", "column": "revenue", "value": "1.05"}, skip=1, limit=3),
        sel(select=[{"column": "install_time"}, {"column": "count"}], filter={"op": "in", "column": "country", "values": ["KZ", "RU", "XX"]}),
        search(dimension="country", term="", filter={"op": "gt", "column": "revenue", "value": "1"}),
        search(dimension="event_name", term="d"),
        search(dimension="event_name", term="re", limit=1, header=True),
        search(dimension="install_time", term="2"),
        search(dimension="country", term="zzz"),
    ]},
    # three segments of two rows: `limit` only breaks the tuple loop, later segments still send one row each
    {"name": "select_multiseg", "table": _table("prune_quirk")["table"], "rows": _table("prune_quirk")["rows"], "queries": [
        sel(dimensions=["n", "s"], metrics=["count"]),
        sel(dimensions=["n", "s"], metrics=["count"], limit=1),
        sel(dimensions=["n", "s"], metrics=["count"], skip=1, limit=2),
        sel(dimensions=["n"], metrics=[], filter={"op": "ge", "column": "n", "value": "10"}, limit=1),
        sel(dimensions=["s"], metrics=["count"], filter={"op": "eq", "column": "s", "value": "a"}, skip=2),
        search(dimension="s", term="", limit=1),
        search(dimension="n", term="1"),
        search(dimension="n", term="1", limit=2, filter={"op": "ne", "column": "s", "value": "b"}),
    ]},
    {"name": "select_types", "table": _table("metric_types")["table"], "rows": _table("metric_types")["rows"], "queries": [
        sel(select=[{"column": "*"}]),
        sel(dimensions=["k"], metrics=["lav", "fmx", "dmx"], filter={"op": "lt", "column": "is", "value": "0"}),
    ]},
    {"name": "select_bitset", "table": _table("users_bitset")["table"], "rows": _table("users_bitset")["rows"], "queries": [
        sel(dimensions=["country", "event_name"], metrics=["user_id"]),
        sel(dimensions=["country"], metrics=["user_id"], filter={"op": "gt", "column": "user_id", "value": "1"}),
        search(dimension="event_name", term="p", filter={"op": "eq", "column": "country", "value": "US"}),
    ]},
    {"name": "select_time", "rollup_ts": NOW, "table": _table("time_rollup")["table"], "rows": _table("time_rollup")["rows"][:12], "queries": [
        sel(select=[{"column": "install_time", "format": "%Y-%m-%d %H:%M:%S"}, {"column": "mt"}, {"column": "count"}], limit=5),
        sel(select=[{"column": "install_time"}, {"column": "country"}], skip=9),
    ]},
    {"name": "select_c1", "table": _table("c1_twin")["table"], "generate": _table("c1_twin")["generate"], "queries": [
        sel(dimensions=["d0", "d1", "d2"], metrics=["m1", "m2", "count"], filter={"op": "eq", "column": "d0", "value": "a7"}, skip=5, limit=10),
        sel(dimensions=["d3"], metrics=["m1"], filter={"op": "and", "filters": [
            {"op": "eq", "column": "d1", "value": "b3"}, {"op": "eq", "column": "d2", "value": "c4"}]}),
        search(dimension="d1", term="b1"),
        search(dimension="d1", term="b1", limit=3, filter={"op": "eq", "column": "d0", "value": "a2"}),
        search(dimension="d2", term="c"),
    ]},
]


# ---- extra edge cases for the aggregate oracle (CPU pinning only; not run by the drop-in GPU tests) ----
EDGE_SCENARIOS = [
    {"name": "edge_empty", "table": INAPP, "rows": [], "queries": [
        agg(dimensions=["country"], metrics=["count", "revenue"]),
        agg(dimensions=[], metrics=["count"]),
        agg(dimensions=["country"], metrics=["count"], header=True, sort=[{"column": "count"}], limit=3),
    ]},
    {"name": "edge_ragged", "table": {"name": "events", "segment_size": 3,
                                      "dimensions": [{"name": "k"}, {"name": "n", "type": "int"}, {"name": "f", "type": "double"}],
                                      "metrics": [{"name": "count", "type": "count"}, {"name": "ulmx", "type": "ulong_max"},
                                                  {"name": "lmn", "type": "long_min"}, {"name": "isum", "type": "int_sum"}]},
     "rows": [["a", "-5", "0.0", "18446744073709551615", "-9223372036854775808", "2147483647"],
              ["a", "-5", "-0.0", "1", "5", "2147483647"],
              ["b", "7", "1.5", "0", "0", "-1"],
              ["b", "-2147483648", "1.5", "3", "-3", "1"],
              ["c", "2147483647", "-1e300", "9", "9", "0"],
              ["a", "0", "2.5", "4", "4", "4"],
              ["c", "1", "2.5", "18446744073709551614", "8", "-2147483648"]],
     "queries": [
         agg(dimensions=["k"], metrics=["count", "ulmx", "lmn", "isum"]),
         agg(dimensions=["n"], metrics=["count"], sort=[{"column": "n", "ascending": True}]),
         agg(dimensions=["f"], metrics=["count", "isum"]),
         agg(dimensions=["k", "n"], metrics=["ulmx"], filter={"op": "lt", "column": "n", "value": "0"}),
         agg(dimensions=["k"], metrics=["count"], filter={"op": "ge", "column": "f", "value": "1.5"}),
         agg(dimensions=["k"], metrics=["isum"], having={"op": "lt", "column": "isum", "value": "0"}),
         agg(dimensions=["k"], metrics=["count"], filter={"op": "not", "filter": {"op": "in", "column": "k", "values": ["a", "zz"]}}),
         agg(dimensions=["k"], metrics=["count"], filter={"op": "in", "column": "k", "values": ["zz", "yy"]}),
         agg(dimensions=["k"], metrics=["lmn"], filter={"op": "gt", "column": "ulmx", "value": "3"},
             sort=[{"column": "lmn", "ascending": True}, {"column": "k"}]),
     ]},
    {"name": "edge_bitset_dups", "table": {"name": "events", "segment_size": 2, "dimensions": [{"name": "c"}],
                                           "metrics": [{"name": "u", "type": "bitset"}, {"name": "count", "type": "count"}]},
     "rows": [["x", "1"], ["y", "1"], ["x", "2"], ["x", "1"], ["y", "1"], ["z", "4000000000"], ["x", "3"]],
     "queries": [
         agg(dimensions=["c"], metrics=["u", "count"]),
         agg(dimensions=[], metrics=["u"]),
         agg(dimensions=["c"], metrics=["u"], having={"op": "ge", "column": "u", "value": "2"}),
         agg(dimensions=["c"], metrics=["count"], filter={"op": "ge", "column": "u", "value": "2"}),
     ]},
    {"name": "edge_time_literals", "rollup_ts": NOW,
     "table": {"name": "events", "dimensions": [{"name": "t", "type": "time", "format": "posix"}, {"name": "b", "type": "boolean"}],
               "metrics": [{"name": "count", "type": "count"}]},
     "rows": [[str(1496275200 + 3600 * h), "true" if h % 3 else "false"] for h in range(0, 96, 5)],
     "queries": [
         agg(select=[{"column": "t", "granularity": "day", "format": "%Y-%m-%d"}, {"column": "count"}],
             filter={"op": "ge", "column": "t", "value": "2017-06-02"}),
         agg(select=[{"column": "t", "granularity": "day", "format": "%Y-%m-%d"}, {"column": "count"}],
             filter={"op": "lt", "column": "t", "value": "2017-06-02 12:00:00"}),
         agg(select=[{"column": "b"}, {"column": "count"}], filter={"op": "gt", "column": "t", "value": "1496361600"}),
         agg(select=[{"column": "t", "granularity": "month", "format": "%Y-%m"}, {"column": "b"}, {"column": "count"}],
             filter={"op": "eq", "column": "b", "value": "true"}),
         agg(dimensions=["b"], metrics=["count"], filter={"op": "ne", "column": "b", "value": "false"}),
     ]},
]
