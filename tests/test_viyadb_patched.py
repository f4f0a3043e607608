"""SURVEY 8f rank 4 (the part that can be built here): ViyaDB's own db::Database with the integration patch applied.

viyadb_b200/host/viyadb_database.patch is the diff a maintainer applies to ViyaDB (src/db/database.{h,cc},
src/input/loader.cc): `"gpu": true` in the database configuration routes Database::Query — the one entry point of /query,
/sql and the cluster workers (src/db/database.cc:104-111) — through vgpu_host::GpuQueryRunner, and Loader::AfterLoad tells
resident HBM copies that an ingest batch has updated cells in place. oracle/Makefile (patched_db) applies it to copies of
the reference's files, compiles them with the reference's own flags and links them with the rest of the UNMODIFIED
reference objects. Here the C ABI behind it is the mock of tests/mock_vgpu.h, so the wiring runs without a GPU:

  * with "gpu": true every golden aggregate query sent through the REAL Database::Query produces the reference's rows
    and QueryStats (the mock answers with the oracle's group table) — and the mock device saw a plan;
  * without it the very same binary runs the stock g++-JIT path and never touches the device;
  * rows ingested through the reference's loader reach the device (whole segments first, again after the next batch:
    the patched Loader::AfterLoad bumped the ingest epoch).

What cannot be built in this image is viyad itself (HTTP / SQL front-end: Boost.Asio, flex, bison are absent)."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import golden_util as G
import viya_oracle
from viyadb_b200 import db as vdb_mod

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "viyadb_b200", "host", "_build", "viyadb_patched_cli")
STATE = os.path.join(tempfile.gettempdir(), "vgpu_fuzz_state")
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


@pytest.fixture(scope="module")
def cli():
    # the reference process generates and compiles code when a table is created (its JIT: g++ against its own headers),
    # so these tests only run where the reference's sources are — the authoring container, not the GPU box
    ref = os.environ.get("VIYA_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "src")):
        pytest.skip("the reference's sources are not here (its JIT needs them)")
    if not os.path.exists(CLI):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", "ref", "patched_db", f"REF={ref}"], check=True,
                       stdout=subprocess.DEVNULL)
    return CLI


def run(cli, job):
    with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
        json.dump(job, f)
        path = f.name
    try:
        p = subprocess.run([cli, path], capture_output=True, text=True, timeout=600,
                           cwd=os.path.join(ROOT, "oracle", "_ref", "root", "build"))
    finally:
        os.remove(path)
    assert p.stdout.strip(), (p.returncode, p.stderr[-800:])
    out = json.loads(p.stdout.strip().splitlines()[-1])
    assert "fatal" not in out, out.get("fatal")
    return out


def widen(arr):
    a = np.asarray(arr)
    if a.dtype.kind == "f":
        return a.view("<u4").astype("<u8") if a.dtype.itemsize == 4 else a.view("<u8")
    return a.astype("<i8").view("<u8") if a.dtype.kind == "i" else a.astype("<u8")


def selected_names(table, q):
    dims = [d["name"] for d in table["dimensions"]]
    if "select" in q:
        names = []
        for s in q["select"]:
            names += (dims + [m["name"] for m in table["metrics"]]) if s["column"] == "*" else [s["column"]]
        return [n for n in names if n in dims], [n for n in names if n not in dims]
    return list(q.get("dimensions", [])), list(q.get("metrics", []))


def cases_for(recs):
    _, segs, dicts, hidden = viya_oracle.read_dump(G.seg_path(recs[0]["seg"]))
    cases = []
    for rec in recs:
        q = rec["query"]
        res = viya_oracle.run_query(rec["table"], segs, dicts, q, now=rec.get("rollup_ts"), hidden_counts=hidden)
        g = res["groups"]
        kn, an = selected_names(rec["table"], q)
        cases.append({"query": q, "ngroups": res["stats"]["aggregated_recs"], "key_names": kn, "acc_names": an,
                      "keys": [widen(k).tolist() for k in g["keys"]], "accs": [widen(a).tolist() for a in g["accs"]],
                      "hidden": None if g["hidden_count"] is None else np.asarray(g["hidden_count"]).astype("<u8").tolist(),
                      "scanned_recs": res["stats"]["scanned_recs"], "scanned_segments": res["stats"]["scanned_segments"]})
    return cases


def check_rows(rec, got):
    assert "error" not in got, (rec["test"], got.get("error"))
    q = rec["query"]
    ordered = bool(q.get("sort"))
    if (q.get("limit") or q.get("skip")) and not ordered:
        assert len(got["rows"]) == len(rec["rows"]), rec["test"]
    elif ordered:
        assert got["rows"] == rec["rows"] or sorted(got["rows"]) == sorted(rec["rows"]), rec["test"]
    else:
        assert sorted(got["rows"]) == sorted(rec["rows"]), rec["test"]
    for k, val in rec["stats"].items():
        assert got["stats"][k] == val, (rec["test"], k, got["stats"][k], val)


FUZZ = [r for r in G.records("ref_fuzz_scenarios.jsonl") if "error" not in r]
FUZZ_TABLES = sorted({r["test"].split(".")[0] for r in FUZZ})


@pytest.mark.parametrize("name", FUZZ_TABLES[::4])
def test_patched_database_query_runs_the_gpu_runner(cli, name):
    """"gpu": true -> the real Database::Query goes through GpuQueryRunner (mock device) and sends the reference's rows"""
    recs = [r for r in FUZZ if r["test"].split(".")[0] == name]
    hdr, _ = vdb_mod.read_dump(G.seg_path(recs[0]["seg"]))
    out = run(cli, {"table": recs[0]["table"], "dicts": hdr["dicts"], "state_dir": STATE, "gpu": True,
                    "rollup_ts": recs[0].get("rollup_ts") or 1496570140, "cases": cases_for(recs)})
    assert out["mock_device_used"]
    for rec, got in zip(recs, out["results"]):
        check_rows(rec, got)
        assert got["plan"] is not None and got["plan"]["metric_cols"] is not None     # the device saw this query's plan


def test_patched_database_without_the_switch_is_stock(cli):
    """no "gpu" key: the same binary runs the stock g++-JIT runner on really ingested rows, the device is never created"""
    import scenarios
    sc = next(s for s in scenarios.SCENARIOS if s["name"] == "inapp")
    want = {r["test"]: r for r in G.records("ref_scenarios.jsonl") if r["test"].startswith("inapp.")}
    out = run(cli, {"table": sc["table"], "rows": sc["rows"], "state_dir": os.path.join(ROOT, "oracle", "_ref", "state"),
                    "cases": [{"query": q} for q in sc["queries"]]})
    assert not out["mock_device_used"]
    for qi, got in enumerate(out["results"]):
        rec = want[f"inapp.{qi}"]
        if "error" in rec:
            assert "error" in got
            continue
        check_rows(rec, got)
        assert got["plan"] is None


def test_patched_database_ingest_reaches_the_device(cli):
    """Rows go in through the reference's loader; the first query uploads every segment. The next batch updates existing
    tuples IN PLACE and appends one: the patched upsert code reports the updated rows (vgpu_ingest_mark_dirty, resolved by
    the JIT-compiled .so from the executable), the patched Loader::AfterLoad hands them to the resident copy, and the
    following query moves exactly those rows plus the appended one — after which the device holds the live store, cell
    for cell (the mock keeps a shadow of it)."""
    import scenarios
    sc = next(s for s in scenarios.SCENARIOS if s["name"] == "inapp")
    n1 = len(sc["rows"])
    q = sc["queries"][1]
    case = {"query": q, "ngroups": 0, "key_names": ["country"], "acc_names": ["count", "revenue"], "keys": [[]], "accs": [[], []],
            "hidden": None}
    job = {"table": sc["table"], "rows": sc["rows"], "state_dir": os.path.join(ROOT, "oracle", "_ref", "state"), "gpu": True,
           "cases": [case, case]}
    # every existing tuple again, plus a new one: rows [0, n1) updated in place, row n1 appended -> one range
    batch2 = [r[:3] + [str(float(r[3]) * 3 + 1)] for r in sc["rows"]] + [["IL", "gift", "20141114", "7.5"]]
    out = run(cli, dict(job, reload_rows=batch2))
    first, again = out["results"]
    assert first["device_calls"] == [["put", 0, n1]] and first["shadow"]["differing_cells"] == 0
    assert again["device_calls"] == []                                  # nothing changed: nothing moves
    after, after_again = out["results_after_reload"]
    assert after["device_calls"] == [["update", 0, 0, n1 + 1]]
    assert after["shadow"]["rows"] == n1 + 1 and after["shadow"]["differing_cells"] == 0 and not after["shadow"]["missing_or_short_segments"]
    assert after_again["device_calls"] == []
    # two of the existing tuples only (rows 2 and 5 of the segment), nothing appended: two one-row updates
    batch3 = [sc["rows"][2][:3] + ["100.5"], sc["rows"][5][:3] + ["0.25"]]
    out = run(cli, dict(job, reload_rows=batch3))
    after = out["results_after_reload"][0]
    assert after["device_calls"] == [["update", 0, 2, 1], ["update", 0, 5, 1]]
    assert after["shadow"]["differing_cells"] == 0


def test_ingest_into_another_table_leaves_the_resident_copy_alone(cli):
    """the ingest notification is per table (IngestEpoch::Bump(&table_) in the patched Loader::AfterLoad): a batch into
    table B must not make the next query on table A upload A again — a large static table next to one under continuous
    ingest would otherwise cross PCIe after every batch"""
    import scenarios
    sc = next(s for s in scenarios.SCENARIOS if s["name"] == "inapp")
    other = next(s for s in scenarios.SCENARIOS if s["name"] == "prune_quirk")
    other_table = dict(other["table"], name="other")
    q = sc["queries"][1]
    case = {"query": q, "ngroups": 0, "key_names": ["country"], "acc_names": ["count", "revenue"], "keys": [[]], "accs": [[], []],
            "hidden": None}
    out = run(cli, {"table": sc["table"], "rows": sc["rows"], "state_dir": os.path.join(ROOT, "oracle", "_ref", "state"), "gpu": True,
                    "other_table": other_table, "cases": [case], "reload_rows": other["rows"], "reload_into": "other"})
    assert out["results"][0]["device_calls"] == [["put", 0, len(sc["rows"])]]
    assert out["results_after_reload"][0]["device_calls"] == []        # table "events" was not touched by the batch
