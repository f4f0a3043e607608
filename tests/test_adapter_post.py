"""The C++ drop-in adapter's host half on the CPU: GpuQueryRunner::PostAggregate (viyadb_b200/host/gpu_query_runner.h —
HavingEvaluator, number / time / dictionary formatting through the reference's own util::Format, the sort on formatted
strings through its own StringNumCmp, skip / limit) inside a reference process, on every golden record the real
reference answered (gtests, scenarios, edge cases, 338 fuzz queries). The group table a device scan would return comes
from the oracle (pinned to the reference on the same records); the reference's own QueryFactory builds the query objects
and its own FilterArgsPacker packs the HAVING literals. No GPU, no vgpu_* call (tests/adapter_post_harness.cc).

The binary links the reference's objects: it is built by `make -C oracle adapter_post` (part of __graft_entry__.build())
where /root/reference exists; without it the test skips."""
import collections
import json
import os
import subprocess
import tempfile

import numpy as np
import pytest

import golden_util as G
import viya_oracle
from viyadb_b200 import db as vdb_mod

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "viyadb_b200", "host", "_build", "adapter_post_cli")
RECS = [r for name in ("ref_gtest.jsonl", "ref_scenarios.jsonl", "ref_edge_scenarios.jsonl", "ref_fuzz_scenarios.jsonl")
        for r in G.records(name) if "error" not in r and "seg" in r]
# one process per (table, segment dump): its dictionaries are rebuilt once
GROUPS = collections.OrderedDict()
for r in RECS:
    GROUPS.setdefault((json.dumps(r["table"], sort_keys=True), r["seg"], r.get("rollup_ts")), []).append(r)


@pytest.fixture(scope="module")
def cli():
    if not os.path.exists(CLI):
        ref = os.environ.get("VIYA_REFERENCE", "/root/reference")
        if not os.path.isdir(os.path.join(ref, "src")):
            pytest.skip("adapter_post_cli not built and the reference's sources are not here")
        import __graft_entry__ as g
        g.build_lib()
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", "ref", "adapter_post", f"REF={ref}"], check=True,
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
                      "hidden": None if g["hidden_count"] is None else np.asarray(g["hidden_count"]).astype("<u8").tolist()})
    # db::Database generates and compiles the table's Segment class when it is created (one g++ run per distinct schema,
    # cached by source hash): a cache of its own under /tmp — these .so files must not travel with oracle/_ref/state
    job = {"table": recs[0]["table"], "dicts": hdr["dicts"], "cases": cases,
           "state_dir": os.path.join(tempfile.gettempdir(), "vgpu_fuzz_state")}
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
    """every (table, dump) group through one adapter process, several at a time (cold: one g++ run per table schema)"""
    from concurrent.futures import ThreadPoolExecutor
    keys = [k for k in GROUPS if not too_big(GROUPS[k])]
    os.makedirs(os.path.join(tempfile.gettempdir(), "vgpu_fuzz_state"), exist_ok=True)
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as ex:
        outs = list(ex.map(lambda k: run_group(cli, k), keys))
    return dict(zip(keys, outs))


@pytest.mark.parametrize("key", list(GROUPS), ids=[GROUPS[k][0]["test"].split(".")[0] + f"[{len(GROUPS[k])}]" for k in GROUPS])
def test_cpp_adapter_post_aggregation_on_oracle_groups(results, key):
    recs = GROUPS[key]
    if key not in results:
        pytest.skip("result too large for a JSON job file")
    out = results[key]
    assert "fatal" not in out, out.get("fatal")
    for rec, got in zip(recs, out["results"]):
        assert "error" not in got, (rec["test"], got.get("error"))
        q = rec["query"]
        ordered = bool(q.get("sort"))
        if (q.get("limit") or q.get("skip")) and not ordered:
            assert len(got["rows"]) == len(rec["rows"]), rec["test"]     # only the count is defined
        elif ordered:
            assert got["rows"] == rec["rows"] or sorted(got["rows"]) == sorted(rec["rows"]), rec["test"]
        else:
            assert sorted(got["rows"]) == sorted(rec["rows"]), rec["test"]
        assert got["output_recs"] == rec["stats"]["output_recs"], rec["test"]
