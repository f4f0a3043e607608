"""The host side of the drop-in (viyadb_b200/query.py: QueryFactory, FilterArgsPacker, plan lowering, post-aggregation)
on the 504 seeded random queries of tests/golden/ref_fuzz_scenarios.jsonl — WITHOUT a GPU: the group table a device scan
would return is taken from the oracle (which tests/test_oracle_golden.py pins to the real reference on the same
records), the host formats / filters (HAVING) / sorts / cuts it, and the rows must be the reference's rows.

What this covers that needs no device: select-list and sort-column resolution, literal decoding and argument order of
filter and HAVING, rollup boundaries, the plan's shape, AVG division, number / time / dictionary formatting, the sort
on formatted strings (Q12) with skip / limit. What it does not cover is the scan itself — that is tests/test_gpu_*.py."""
import numpy as np
import pytest

import golden_util as G
import viya_oracle
import viyadb_b200 as v
from viyadb_b200 import _native as N
from viyadb_b200 import db as vdb_mod
from viyadb_b200.query import FilterArgsPacker, GpuQueryRunner, MemoryRowOutput, QueryFactory

FUZZ = [r for r in G.records("ref_fuzz_scenarios.jsonl") if "error" not in r]


def open_host_table(rec):
    """Database without a device (schema, dictionaries, plan objects only) with the dictionaries of the dump."""
    db = v.Database({"tables": [rec["table"]]}, device=None)
    t = db.get_table(rec["table"]["name"])
    hdr, _ = vdb_mod.read_dump(G.seg_path(rec["seg"]))
    for d in t.dimensions:
        if d.kind == N.DIM_STRING:
            c2v = hdr["dicts"][d.name]
            d.dict.c2v = list(c2v)
            d.dict.v2c = {s: i for i, s in enumerate(c2v)}
    return db, t


def test_fuzz_fixture_present():
    assert len(FUZZ) >= 300, "tests/golden/ref_fuzz_scenarios.jsonl missing: run tests/golden/make_golden.py fuzz"


@pytest.mark.parametrize("rec", FUZZ, ids=[G.rec_id(r) for r in FUZZ])
def test_host_post_aggregation_on_oracle_groups(rec):
    db, t = open_host_table(rec)
    q = rec["query"]
    query = QueryFactory.create(q, db)
    out = MemoryRowOutput()
    runner = GpuQueryRunner(db, out, now=rec.get("rollup_ts"))

    # ---- the plan a device would get: shape only (the scan is not run here) ----
    plan = runner.build_plan(query)
    assert plan.nkeys == len(query.dimension_cols) and plan.nmetrics == len(query.metric_cols)
    packer = FilterArgsPacker(t).visit(query.filter)
    assert plan.nnodes == len(packer.nodes) and plan.nargs == len(packer.args)
    for i in range(plan.nnodes):
        node = plan.nodes[i]
        assert node.col < len(t.dimensions) + len(t.metrics) + 1
        if node.kind in (N.NODE_RELOP, N.NODE_IN):
            assert node.arg + max(1, node.n if node.kind == N.NODE_IN else 1) <= plan.nargs
    for k in range(plan.nkeys):
        key = plan.keys[k]
        dim = query.dimension_cols[k].dim
        assert key.nrules == (len(dim.rollup_rules) if dim.kind in (N.DIM_TIME, N.DIM_MICROTIME) else 0)
        # boundaries ascend in rule order (rules are sorted by descending `after`, column.cc:346-349)
        b = [key.rule_boundary[r] for r in range(key.nrules)]
        assert b == sorted(b)

    # ---- the device's part, played by the oracle ----
    hdr, segs, dicts, hidden = viya_oracle.read_dump(G.seg_path(rec["seg"]))
    want = viya_oracle.run_query(rec["table"], segs, dicts, q, now=rec.get("rollup_ts"), hidden_counts=hidden)
    g = want["groups"]
    ngroups = want["stats"]["aggregated_recs"]
    groups = {"ngroups": ngroups, "keys": [np.asarray(k) for k in g["keys"]], "accs": [np.asarray(a) for a in g["accs"]],
              "hidden_count": g["hidden_count"]}
    runner.stats.aggregated_recs = ngroups

    # ---- the host's part ----
    hargs = FilterArgsPacker(t).visit(query.having).values if query.having is not None else []
    runner.post_aggregate(query, groups, hargs)
    ordered = bool(q.get("sort"))
    if (q.get("limit") or q.get("skip")) and not ordered:
        assert len(out.rows) == len(rec["rows"])     # only the count is defined
    elif ordered:
        assert out.rows == rec["rows"] or sorted(out.rows) == sorted(rec["rows"])
    else:
        assert sorted(out.rows) == sorted(rec["rows"])
    assert runner.stats.output_recs == rec["stats"]["output_recs"]
