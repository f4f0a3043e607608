"""The C++ drop-in adapter — whole and unmodified — on the CPU, against a mock device.

viyadb_b200/host/gpu_query_runner.h is what a ViyaDB maintainer ships: a query::QueryVisitor that packs literals with the
reference's own FilterArgsPacker, lowers the query to a vgpu_plan, calls the C ABI and post-aggregates the returned group
table (HAVING, util::Format, the sort on formatted strings through StringNumCmp, skip / limit). tests/adapter_mock_harness.cc
links it, inside a reference process, against a MOCK of include/vgpu.h instead of libvgpu.so: the mock records the plan
and answers with the group table the oracle computed (the oracle is pinned to the real reference on the same records).
For every golden record the real reference answered (gtests, scenarios, edge cases, 478 fuzz queries):

  * the rows `query->Accept(GpuQueryRunner)` sends == the reference's rows, QueryStats included;
  * the plan the C++ adapter lowered == the plan the Python mirror lowers (viyadb_b200/query.py — the host the GPU
    parity suite drives), node for node, literal for literal (the bytes of the literal's own type: the rest of an
    AnyNum image is uninitialised in the reference), rollup boundaries, post-aggregation request; same schema.

The binary links the reference's objects: built by `make -C oracle adapter_mock` (part of __graft_entry__.build()) where
/root/reference exists; without it the test skips. No GPU, no libvgpu.so: nothing here computes a scan."""
import collections
import json
import os
import subprocess
import tempfile

import numpy as np
import pytest

import golden_util as G
import viya_oracle
import viyadb_b200 as v
from viyadb_b200 import _native as N
from viyadb_b200 import db as vdb_mod
from viyadb_b200.query import GpuQueryRunner, MemoryRowOutput, QueryFactory

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.environ.get("VGPU_ADAPTER_CLI") or os.path.join(ROOT, "viyadb_b200", "host", "_build", "adapter_mock_cli")   # override: sanitizer builds
STATE = os.path.join(tempfile.gettempdir(), "vgpu_fuzz_state")
RECS = [r for name in ("ref_gtest.jsonl", "ref_scenarios.jsonl", "ref_edge_scenarios.jsonl", "ref_fuzz_scenarios.jsonl")
        for r in G.records(name) if "error" not in r and "seg" in r]
# one process per (table, segment dump): its dictionaries are rebuilt once
GROUPS = collections.OrderedDict()
for r in RECS:
    GROUPS.setdefault((json.dumps(r["table"], sort_keys=True), r["seg"], r.get("rollup_ts")), []).append(r)


@pytest.fixture(scope="module")
def cli():
    # the reference process generates and compiles code when a table is created (its JIT: g++ against its own headers),
    # so these tests only run where the reference's sources are — the authoring container, not the GPU box
    ref = os.environ.get("VIYA_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "src")):
        pytest.skip("the reference's sources are not here (its JIT needs them)")
    if not os.path.exists(CLI):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", "ref", "adapter_mock", f"REF={ref}"], check=True,
                       stdout=subprocess.DEVNULL)
    return CLI


def widen(arr):
    a = np.asarray(arr)
    if a.dtype.kind == "f":
        return a.view("<u4").astype("<u8") if a.dtype.itemsize == 4 else a.view("<u8")
    return a.astype("<i8").view("<u8") if a.dtype.kind == "i" else a.astype("<u8")


def too_big(recs):
    return any(len(r["rows"]) > 20000 for r in recs)


def run_group(cli, key):
    recs = GROUPS[key]
    hdr, _ = vdb_mod.read_dump(G.seg_path(recs[0]["seg"]))
    _, segs, dicts, hidden = viya_oracle.read_dump(G.seg_path(recs[0]["seg"]))
    cases = []
    for rec in recs:
        q = rec["query"]
        res = viya_oracle.run_query(rec["table"], segs, dicts, q, now=rec.get("rollup_ts"), hidden_counts=hidden)
        g = res["groups"]
        n = res["stats"]["aggregated_recs"]
        cases.append({"query": q, "ngroups": n,
                      "keys": [widen(k).tolist() for k in g["keys"]], "accs": [widen(a).tolist() for a in g["accs"]],
                      "hidden": None if g["hidden_count"] is None else np.asarray(g["hidden_count"]).astype("<u8").tolist(),
                      "scanned_recs": res["stats"]["scanned_recs"], "scanned_segments": res["stats"]["scanned_segments"]})
    # db::Database generates and compiles the table's Segment class when it is created, GpuTableBinding its segment
    # accessor (one g++ run each per distinct schema, cached by source hash): a cache of its own under /tmp — these .so
    # files must not travel with oracle/_ref/state
    job = {"table": recs[0]["table"], "dicts": hdr["dicts"], "cases": cases, "state_dir": STATE}
    if recs[0].get("rollup_ts") is not None:
        job["rollup_ts"] = recs[0]["rollup_ts"]
    with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
        json.dump(job, f)
        path = f.name
    try:
        # the reference's JIT resolves its include / library paths relative to the CWD (compiler.cc:46-54)
        p = subprocess.run([cli, path], capture_output=True, text=True, timeout=600,
                           cwd=os.path.join(ROOT, "oracle", "_ref", "root", "build"))
    finally:
        os.remove(path)
    if p.returncode != 0 or not p.stdout.strip():
        return {"fatal": (p.stdout[-500:], p.stderr[-500:])}
    return json.loads(p.stdout.strip().splitlines()[-1])


@pytest.fixture(scope="module")
def results(cli):
    """every (table, dump) group through one adapter process, several at a time (cold: two g++ runs per table schema)"""
    from concurrent.futures import ThreadPoolExecutor
    keys = [k for k in GROUPS if not too_big(GROUPS[k])]
    os.makedirs(STATE, exist_ok=True)
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as ex:
        outs = list(ex.map(lambda k: run_group(cli, k), keys))
    return dict(zip(keys, outs))


# ---- the Python mirror's plan and schema, as plain data ----
def mirror_schema(t):
    cols = []
    for c in t.dimensions + t.metrics:
        btype, lit = c.type, 0
        if c.kind == N.METRIC_BITSET and c.type in (N.U8, N.U16):
            btype, lit = N.U32, c.type + 1
        cols.append([c.kind, btype, c.agg, lit])
    if t.has_hidden_count:
        cols.append([N.METRIC_HIDDEN_COUNT, N.U64, N.AGG_COUNT, 0])
    return cols


def mirror_plan(rec):
    db = v.Database({"tables": [rec["table"]]}, device=None)
    t = db.get_table(rec["table"]["name"])
    hdr, _ = vdb_mod.read_dump(G.seg_path(rec["seg"]))
    for d in t.dimensions:
        if d.kind == N.DIM_STRING:
            c2v = hdr["dicts"][d.name]
            d.dict.c2v = list(c2v)
            d.dict.v2c = {s: i for i, s in enumerate(c2v)}
    query = QueryFactory.create(rec["query"], db)
    plan = GpuQueryRunner(db, MemoryRowOutput(), now=rec.get("rollup_ts")).build_plan(query)

    def nodes(arr, n):
        return [[arr[i].kind, arr[i].op, arr[i].col, arr[i].arg, arr[i].n] for i in range(n)]
    keys = []
    for k in range(plan.nkeys):
        key = plan.keys[k]
        keys.append({"col": key.col, "nrules": key.nrules, "query_granularity": key.query_granularity,
                     "rule_boundary": [key.rule_boundary[r] for r in range(key.nrules)],
                     "rule_granularity": [key.rule_granularity[r] for r in range(key.nrules)]})
    cols = t.dimensions + t.metrics
    widths = [N.TYPE_WIDTH[c.type] for c in cols] + [8]
    return {"nodes": nodes(plan.nodes, plan.nnodes), "args": [plan.args[i] for i in range(plan.nargs)], "keys": keys,
            "metric_cols": [plan.metric_cols[i] for i in range(plan.nmetrics)], "need_hidden_count": plan.need_hidden_count,
            "flags": plan.flags, "hnodes": nodes(plan.hnodes, plan.nhnodes) if plan.nhnodes else [],
            "hargs": [plan.hargs[i] for i in range(plan.nhargs)] if plan.nhargs else [],
            "sort_col": plan.sort_col, "sort_descending": plan.sort_descending, "top_k": plan.top_k}, widths, mirror_schema(t)


def masked_args(nodes, args, widths):
    """the bytes of each literal's own type (db::AnyNum leaves the rest of its 8 bytes uninitialised, column.h:98-121)"""
    out = list(args)
    for kind, op, col, arg, n in nodes:
        if kind == N.NODE_RELOP:
            span = range(arg, arg + 1)
        elif kind == N.NODE_IN:
            span = range(arg, arg + n)
        else:
            continue
        mask = (1 << (8 * widths[col])) - 1
        for i in span:
            out[i] = args[i] & mask
    return out


@pytest.mark.parametrize("key", list(GROUPS), ids=[GROUPS[k][0]["test"].split(".")[0] + f"[{len(GROUPS[k])}]" for k in GROUPS])
def test_cpp_adapter_against_mock_device(results, key):
    recs = GROUPS[key]
    if key not in results:
        pytest.skip("result too large for a JSON job file")
    out = results[key]
    assert "fatal" not in out, out.get("fatal")
    for rec, got in zip(recs, out["results"]):
        assert "error" not in got, (rec["test"], got.get("error"))
        q = rec["query"]
        # ---- rows and QueryStats == the reference's ----
        ordered = bool(q.get("sort"))
        if (q.get("limit") or q.get("skip")) and not ordered:
            assert len(got["rows"]) == len(rec["rows"]), rec["test"]     # only the count is defined
        elif ordered:
            assert got["rows"] == rec["rows"] or sorted(got["rows"]) == sorted(rec["rows"]), rec["test"]
        else:
            assert sorted(got["rows"]) == sorted(rec["rows"]), rec["test"]
        for k, val in rec["stats"].items():
            assert got["stats"][k] == val, (rec["test"], k, got["stats"][k], val)
        # ---- the plan the C++ adapter lowered == the Python mirror's ----
        want, widths, schema = mirror_plan(rec)
        plan = got["plan"]
        assert got["schema"]["cols"] == schema, rec["test"]
        assert plan["nodes"] == want["nodes"], rec["test"]
        assert masked_args(plan["nodes"], plan["args"], widths) == masked_args(want["nodes"], want["args"], widths), rec["test"]
        assert plan["hnodes"] == want["hnodes"], rec["test"]
        assert masked_args(plan["hnodes"], plan["hargs"], widths) == masked_args(want["hnodes"], want["hargs"], widths), rec["test"]
        for f in ("keys", "metric_cols", "need_hidden_count", "flags", "sort_col", "sort_descending", "top_k"):
            assert plan[f] == want[f], (rec["test"], f, plan[f], want[f])


# ------------------------------------------------------------------------------------------------
# select / search through the same mock
# ------------------------------------------------------------------------------------------------
SEL_RECS = [r for name in ("ref_gtest_select.jsonl", "ref_select_scenarios.jsonl", "ref_fuzz_select_scenarios.jsonl")
            for r in G.records(name) if "error" not in r and "seg" in r]
SEL_GROUPS = collections.OrderedDict()
for r in SEL_RECS:
    SEL_GROUPS.setdefault((json.dumps(r["table"], sort_keys=True), r["seg"], r.get("rollup_ts")), []).append(r)


def float_search(rec):
    """search on a floating-point dimension: the adapter delegates it to the stock runner (the device's first-row table
    is keyed by integer values) — nothing to mock"""
    q = rec["query"]
    if q["type"] != "search":
        return False
    d = next(d for d in rec["table"]["dimensions"] if d["name"] == q["dimension"])
    return d.get("type") in ("float", "double")


def select_case(rec, segs, dicts, hidden, dims, mets):
    q = rec["query"]
    ocols = {c.name: c for c in dims + mets}
    flt = viya_oracle.make_filter(q.get("filter"))
    case = {"query": q}
    if q["type"] == "select":
        res = viya_oracle.run_select(rec["table"], segs, dicts, q, hidden_counts=hidden)
        picked = res["picked"]
        cells = []
        for c in dims + mets:
            col = []
            for si, i in picked:
                if c.is_dim or c.agg != "bitset":
                    col.append(int(widen(np.asarray(segs[si][c.name][i:i + 1]))[0]))
                else:
                    offsets, values = segs[si][c.name]
                    col.append(len(set(np.asarray(values[int(offsets[i]):int(offsets[i + 1])]).tolist())))
            cells.append(col)
        if hidden is not None and any(h is not None for h in hidden):
            cells.append([int(hidden[si][i]) for si, i in picked])
        case.update({"nrows": len(picked), "cells": cells})
    else:
        # what vgpu_query_search returns: per processed segment the distinct values among the passing rows with the
        # first row that holds each, ascending by that row
        d = ocols[q["dimension"]]
        offsets, codes, rows = [0], [], []
        for seg in segs:
            n = viya_oracle._seg_rows(seg, dims + mets)
            if not viya_oracle.process_segment(flt, seg, n, ocols, dicts):
                continue
            if n:
                passing = np.nonzero(viya_oracle.eval_filter(flt, seg, n, ocols, dicts))[0]
                vals = widen(np.asarray(seg[d.name])[passing])
                seen = set()
                for r_, v_ in zip(passing.tolist(), vals.tolist()):
                    if v_ not in seen:
                        seen.add(v_)
                        codes.append(v_)
                        rows.append(r_)
            offsets.append(len(codes))
        res = viya_oracle.run_search(rec["table"], segs, dicts, q)
        case.update({"seg_offsets": offsets, "codes": codes, "first_row": rows})
    case["scanned_recs"] = res["stats"]["scanned_recs"]
    case["scanned_segments"] = res["stats"]["scanned_segments"]
    return case


def run_select_group(cli, key):
    recs = [r for r in SEL_GROUPS[key] if not float_search(r)]
    hdr, _ = vdb_mod.read_dump(G.seg_path(recs[0]["seg"]))
    _, segs, dicts, hidden = viya_oracle.read_dump(G.seg_path(recs[0]["seg"]))
    dims, mets = viya_oracle.parse_schema(recs[0]["table"])
    job = {"table": recs[0]["table"], "dicts": hdr["dicts"], "state_dir": STATE,
           "cases": [select_case(r, segs, dicts, hidden, dims, mets) for r in recs]}
    with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
        json.dump(job, f)
        path = f.name
    try:
        p = subprocess.run([cli, path], capture_output=True, text=True, timeout=600,
                           cwd=os.path.join(ROOT, "oracle", "_ref", "root", "build"))
    finally:
        os.remove(path)
    if p.returncode != 0 or not p.stdout.strip():
        return {"fatal": (p.stdout[-500:], p.stderr[-500:])}
    return json.loads(p.stdout.strip().splitlines()[-1])


@pytest.fixture(scope="module")
def select_results(cli):
    from concurrent.futures import ThreadPoolExecutor
    keys = [k for k in SEL_GROUPS if any(not float_search(r) for r in SEL_GROUPS[k])]
    os.makedirs(STATE, exist_ok=True)
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as ex:
        outs = list(ex.map(lambda k: run_select_group(cli, k), keys))
    return dict(zip(keys, outs))


@pytest.mark.parametrize("key", list(SEL_GROUPS), ids=[SEL_GROUPS[k][0]["test"].split(".")[0] + f"[{len(SEL_GROUPS[k])}]" for k in SEL_GROUPS])
def test_cpp_adapter_select_search_against_mock_device(select_results, key):
    if key not in select_results:
        pytest.skip("only searches on floating-point dimensions: delegated to the stock runner")
    out = select_results[key]
    assert "fatal" not in out, out.get("fatal")
    recs = [r for r in SEL_GROUPS[key] if not float_search(r)]
    for rec, got in zip(recs, out["results"]):
        assert "error" not in got, (rec["test"], got.get("error"))
        assert got["rows"] == rec["rows"], rec["test"]          # select / search output order is defined
        for k, val in rec["stats"].items():
            assert got["stats"][k] == val, (rec["test"], k, got["stats"][k], val)
        # ---- the predicate and the request the C++ adapter lowered == the Python mirror's ----
        db = v.Database({"tables": [rec["table"]]}, device=None)
        t = db.get_table(rec["table"]["name"])
        hdr, _ = vdb_mod.read_dump(G.seg_path(rec["seg"]))
        for d in t.dimensions:
            if d.kind == N.DIM_STRING:
                d.dict.c2v = list(hdr["dicts"][d.name])
                d.dict.v2c = {s: i for i, s in enumerate(d.dict.c2v)}
        query = QueryFactory.create(rec["query"], db)
        packer, _, _ = GpuQueryRunner(db, MemoryRowOutput())._predicate(query)
        widths = [N.TYPE_WIDTH[c.type] for c in t.dimensions + t.metrics] + [8]
        plan = got["plan"]
        want_nodes = [list(nd) for nd in packer.nodes]
        assert plan["nodes"] == want_nodes, rec["test"]
        assert masked_args(plan["nodes"], plan["args"], widths) == masked_args(want_nodes, packer.args, widths), rec["test"]
        if rec["query"]["type"] == "select":
            want_cols = [t.schema_index(dc.dim) for dc in query.dimension_cols] + [t.schema_index(mc.metric) for mc in query.metric_cols]
            has_avg = any(mc.metric.agg == N.AGG_AVG for mc in query.metric_cols)
            has_count = any(mc.metric.agg == N.AGG_COUNT for mc in query.metric_cols)
            if has_avg and not has_count:
                want_cols.append(t.hidden_count_index)
            assert plan["cols"] == want_cols and plan["skip"] == query.skip and plan["limit"] == query.limit, rec["test"]
        else:
            assert plan["col"] == t.schema_index(query.dimension), rec["test"]


# ------------------------------------------------------------------------------------------------
# incremental segment sync: GpuTableBinding::Sync() against the reference's live store
# ------------------------------------------------------------------------------------------------
def run_sync(cli, table, steps):
    """rows go through the reference's own ingest; after every Sync() the mock's shadow of HBM is compared with the live
    segments cell by cell. JIT cache: oracle/_ref/state, pre-warmed for the scenario tables by __graft_entry__.build()."""
    job = {"table": table, "state_dir": os.path.join(ROOT, "oracle", "_ref", "state"), "sync": steps}
    with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
        json.dump(job, f)
        path = f.name
    try:
        p = subprocess.run([cli, path], capture_output=True, text=True, timeout=600,
                           cwd=os.path.join(ROOT, "oracle", "_ref", "root", "build"))
    finally:
        os.remove(path)
    assert p.returncode == 0 and p.stdout.strip(), (p.stdout[-500:], p.stderr[-500:])
    out = json.loads(p.stdout.strip().splitlines()[-1])
    assert "fatal" not in out, out.get("fatal")
    return out["sync"]


def scenario(name):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import scenarios
    return next(s for s in scenarios.SCENARIOS if s["name"] == name)


@pytest.mark.parametrize("notify", ["epoch", "mark", "none"])
def test_sync_follows_in_place_upserts(cli, notify):
    """A second batch into the SAME dimension tuples updates metric cells in place (upsert.cc:386-393) and appends one new
    tuple. With a notification (coarse epoch, or exact dirty ranges) the resident copy is current after Sync(); "mark"
    moves only the dirty + appended rows (one vgpu_segment_update), "epoch" the whole segment. Without any notification the
    binding only sees the appended row — the in-place updates stay stale: that is why the integration has its one line in
    Loader::AfterLoad."""
    sc = scenario("inapp")
    batch2 = [r[:3] + [str(float(r[3]) * 3 + 1)] for r in sc["rows"]] + [["IL", "gift", "20141114", "7.5"]]
    first, second = run_sync(cli, sc["table"], [{"rows": sc["rows"], "notify": "none"}, {"rows": batch2, "notify": notify}])
    n1 = first["rows"]
    assert first["calls"] == [["put", 0, n1]] and first["differing_cells"] == 0 and not first["missing_or_short_segments"]
    assert second["rows"] == n1 + 1 and not second["missing_or_short_segments"]
    if notify == "epoch":
        assert second["calls"] == [["put", 0, n1 + 1]] and second["differing_cells"] == 0
    elif notify == "mark":
        assert second["calls"] == [["update", 0, 0, n1 + 1]] and second["partial_updates"] == 1
        assert second["differing_cells"] == 0
    else:
        assert second["calls"] == [["update", 0, n1, 1]]          # the appended row only
        assert second["differing_cells"] > 0                       # the updated metric cells were never re-uploaded


def test_sync_appends_across_ragged_segments(cli):
    """segment_size 2: appended rows fill the last segment and open new ones; rows that merge into existing tuples bump
    their count in place. Exact dirty ranges: every existing segment gets one update, new segments one put each."""
    sc = scenario("prune_quirk")
    batch2 = [["11", "c"], ["30", "d"], ["31", "e"], ["1", "a"]]       # two upserts into existing tuples, two new tuples
    first, second = run_sync(cli, sc["table"], [{"rows": sc["rows"], "notify": "none"}, {"rows": batch2, "notify": "mark"}])
    assert [c[0] for c in first["calls"]] == ["put"] * first["segments"] and first["differing_cells"] == 0
    assert second["rows"] == first["rows"] + 2 and second["differing_cells"] == 0 and not second["missing_or_short_segments"]
    kinds = [c[0] for c in second["calls"]]
    assert kinds.count("update") == first["segments"] and kinds.count("put") == second["segments"] - first["segments"]
    # a third batch with nothing new and no notification: nothing to move
    third = run_sync(cli, sc["table"], [{"rows": sc["rows"], "notify": "none"}, {"rows": [], "notify": "none"}])[1]
    assert third["calls"] == [] and third["differing_cells"] == 0


def test_sync_bitset_tables_take_the_whole_segment(cli):
    """bitset cells are flattened to CSR (their length changes when an upsert adds an id): no partial update"""
    sc = scenario("users_bitset")
    batch2 = [list(r) for r in sc["rows"][:3]]
    for r in batch2:
        r[-1] = str(int(r[-1]) + 1000)                                  # new ids into existing cells
    first, second = run_sync(cli, sc["table"], [{"rows": sc["rows"], "notify": "none"}, {"rows": batch2, "notify": "mark"}])
    assert first["differing_cells"] == 0 and second["differing_cells"] == 0
    assert [c[0] for c in second["calls"]] == ["put"] * second["segments"] and second["partial_updates"] == 0


# ------------------------------------------------------------------------------------------------
# concurrency (SURVEY 8b "Threading"): query_threads queries at once over the shared bindings, next to ingest notifications
# ------------------------------------------------------------------------------------------------
TSAN_CLI = os.path.join(ROOT, "viyadb_b200", "host", "_build", "adapter_mock_tsan")


def concurrent_job(name="fuzzb01", threads=6, repeat=3):
    recs = [r for r in RECS if r["test"].split(".")[0] == name]
    key = (json.dumps(recs[0]["table"], sort_keys=True), recs[0]["seg"], recs[0].get("rollup_ts"))
    hdr, _ = vdb_mod.read_dump(G.seg_path(recs[0]["seg"]))
    _, segs, dicts, hidden = viya_oracle.read_dump(G.seg_path(recs[0]["seg"]))
    cases = []
    for rec in GROUPS[key]:
        res = viya_oracle.run_query(rec["table"], segs, dicts, rec["query"], now=rec.get("rollup_ts"), hidden_counts=hidden)
        g = res["groups"]
        cases.append({"query": rec["query"], "ngroups": res["stats"]["aggregated_recs"],
                      "keys": [widen(k).tolist() for k in g["keys"]], "accs": [widen(a).tolist() for a in g["accs"]],
                      "hidden": None if g["hidden_count"] is None else np.asarray(g["hidden_count"]).astype("<u8").tolist()})
    job = {"table": recs[0]["table"], "dicts": hdr["dicts"], "cases": cases, "state_dir": STATE,
           "concurrent": {"threads": threads, "repeat": repeat}}
    if recs[0].get("rollup_ts") is not None:
        job["rollup_ts"] = recs[0]["rollup_ts"]
    return job, len(cases)


def run_concurrent(binary, job, env=None):
    with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
        json.dump(job, f)
        path = f.name
    try:
        return subprocess.run([binary, path], capture_output=True, text=True, timeout=900, env=env,
                              cwd=os.path.join(ROOT, "oracle", "_ref", "root", "build"))
    finally:
        os.remove(path)


def test_concurrent_queries_through_the_adapter(cli):
    """six threads, each running every query of a fuzz table three times through its own GpuQueryRunner over the shared
    bindings, while another thread sends ingest notifications: every answer equals the single-threaded one"""
    job, ncases = concurrent_job()
    p = run_concurrent(cli, job)
    assert p.returncode == 0 and p.stdout.strip(), (p.stdout[-500:], p.stderr[-500:])
    out = json.loads(p.stdout.strip().splitlines()[-1])["concurrent"]
    assert out["queries"] == 6 * 3 * ncases and out["mismatches"] == 0 and out["errors"] == 0


def test_concurrent_queries_under_thread_sanitizer(cli):
    """the same run in a -fsanitize=thread build of the harness (the adapter header, the reference's inline header code
    it calls and the mock are instrumented): ThreadSanitizer must report nothing"""
    if not os.path.exists(TSAN_CLI):
        ref = os.environ.get("VIYA_REFERENCE", "/root/reference")
        if not os.path.exists("/usr/bin/g++"):
            pytest.skip("no g++ with libtsan")
        r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", "ref", "adapter_mock_tsan", f"REF={ref}"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            pytest.skip("ThreadSanitizer build not available: " + r.stderr[-200:])
    job, ncases = concurrent_job(threads=8, repeat=2)
    env = dict(os.environ, TSAN_OPTIONS="exitcode=66 halt_on_error=0")
    p = run_concurrent(TSAN_CLI, job, env)
    if "FATAL: ThreadSanitizer" in p.stderr or "unexpected memory mapping" in p.stderr:
        pytest.skip("the ThreadSanitizer runtime does not start on this kernel: " + p.stderr.strip().splitlines()[0][:120])
    assert "ThreadSanitizer" not in p.stderr, p.stderr[-3000:]
    assert p.returncode == 0 and p.stdout.strip(), (p.returncode, p.stdout[-300:], p.stderr[-1500:])
    out = json.loads(p.stdout.strip().splitlines()[-1])["concurrent"]
    assert out["queries"] == 8 * 2 * ncases and out["mismatches"] == 0 and out["errors"] == 0
