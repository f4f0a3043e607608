"""Select / search queries on the CUDA path vs the REAL reference (golden vectors) and vs the oracle on a seeded
162k-row table: rows in the reference's output order, QueryStats exact. Through the C ABI (vgpu_query_select,
vgpu_query_search)."""
import numpy as np
import pytest

import golden_util as G
import viya_oracle
from helpers import random_table, upload

pytestmark = pytest.mark.gpu

RECS = [r for r in G.records("ref_gtest_select.jsonl") + G.records("ref_select_scenarios.jsonl")]


@pytest.fixture(scope="module")
def vdb(built_lib):
    import viyadb_b200
    return viyadb_b200


@pytest.mark.parametrize("rec", RECS, ids=[G.rec_id(r) for r in RECS])
def test_reference_select_search(vdb, rec):
    db = vdb.Database({"tables": [rec["table"]]}, device=0)
    try:
        t = db.get_table(rec["table"]["name"])
        t.load_dump(G.seg_path(rec["seg"]))
        out = vdb.MemoryRowOutput()
        if "error" in rec:
            with pytest.raises((ValueError, OverflowError, KeyError, vdb.VgpuError)):
                db.query(rec["query"], out)
            return
        stats = db.query(rec["query"], out)
        assert out.rows == rec["rows"]
        for k, v in rec["stats"].items():
            assert getattr(stats, k) == v, (k, getattr(stats, k), v)
    finally:
        db.close()


EVENTS = {"name": "events", "segment_size": 50000,
          "dimensions": [{"name": "d0"}, {"name": "d1", "cardinality": 200}, {"name": "n3", "type": "ushort"},
                         {"name": "i4", "type": "int"}, {"name": "t5", "type": "time"}, {"name": "b6", "type": "boolean"}],
          "metrics": [{"name": "count", "type": "count"}, {"name": "ls", "type": "long_sum"},
                      {"name": "savg", "type": "short_avg"}, {"name": "uid", "type": "bitset"},
                      {"name": "ds", "type": "double_sum"}]}
SPEC = {"d0": (1, 16), "d1": (1, 100), "n3": (0, 999), "i4": (-500, 500), "t5": (1490000000, 1496570140), "b6": (0, 1),
        "count": (1, 3), "ls": (-2**50, 2**50), "savg": (-3000, 3000), "uid": ("ids", 5000, 3), "ds": ("float", -1e3, 1e3)}
QUERIES = [
    {"type": "select", "dimensions": ["d0", "d1", "n3"], "metrics": ["ls", "count"],
     "filter": {"op": "eq", "column": "d0", "value": "d0_7"}, "skip": 100, "limit": 1000},
    {"type": "select", "select": [{"column": "*"}], "filter": {"op": "and", "filters": [
        {"op": "lt", "column": "n3", "value": "5"}, {"op": "gt", "column": "i4", "value": "0"}]}},
    {"type": "select", "select": [{"column": "t5", "format": "%Y-%m-%d %H:%M:%S"}, {"column": "savg"}, {"column": "uid"}, {"column": "count"}],
     "filter": {"op": "eq", "column": "b6", "value": "true"}, "limit": 49999},
    {"type": "select", "dimensions": ["i4"], "metrics": [], "skip": 149990},
    {"type": "search", "dimension": "d1", "term": "_1"},
    {"type": "search", "dimension": "d1", "term": "_1", "limit": 3, "filter": {"op": "gt", "column": "n3", "value": "900"}},
    {"type": "search", "dimension": "i4", "term": "-4", "header": True},
    {"type": "search", "dimension": "b6", "term": "e"},
]


@pytest.fixture(scope="module")
def env(vdb):
    segs, dicts, hidden = random_table(EVENTS, 4, 50000, 4321, SPEC, last_rows=12345)
    db = vdb.Database({"tables": [EVENTS]}, device=0)
    upload(db.get_table("events"), segs, dicts, hidden)
    yield db, segs, dicts, hidden
    db.close()


@pytest.mark.parametrize("qi", range(len(QUERIES)))
def test_select_search_matches_oracle(vdb, env, qi):
    db, segs, dicts, hidden = env
    q = dict(QUERIES[qi], table="events")
    out = vdb.MemoryRowOutput()
    stats = db.query(q, out)
    if q["type"] == "select":
        want = viya_oracle.run_select(EVENTS, segs, dicts, q, hidden_counts=hidden)
    else:
        want = viya_oracle.run_search(EVENTS, segs, dicts, q)
    assert out.rows == want["rows"]
    for k in ("scanned_segments", "scanned_recs", "output_recs"):
        assert getattr(stats, k) == want["stats"][k], k
    if q["type"] == "search":
        assert stats.aggregated_recs == want["stats"]["aggregated_recs"]
