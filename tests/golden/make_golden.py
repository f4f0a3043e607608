#!/usr/bin/env python
"""tests/golden/make_golden.py — regenerates the committed golden fixtures. TEST INFRASTRUCTURE.

Runs only in the authoring container (needs /root/reference and oracle/_ref built by
`make -C oracle ref ref_capture`):

  1. oracle/run_capture.sh executes the reference's OWN gtest suite (unmodified sources; 73 tests
     pass) with Database::Query wrapped (oracle/capture_hook.cc). Every aggregate query of
     test/{aggregation,bitset,boolean,filter,index,limits,load,metrics,sort,time}.cc is recorded
     with the rows the reference produced, its QueryStats, and a dump of the real segments.
     -> tests/golden/ref_gtest.jsonl + tests/golden/seg/<sha1>.bin.gz
  2. oracle/_ref/oracle_cli runs extra scenarios (tests/golden/scenarios.py: small-N twins of the
     benchmark configs C0-C4 and edge cases the gtests do not cover) through the real reference.
     -> tests/golden/ref_scenarios.jsonl + seg files

  3. tests/golden/fuzz_scenarios.py: seeded random tables / rows / queries through the real reference
     -> tests/golden/ref_fuzz_scenarios.jsonl + seg files (the differential pin of the numpy oracle)

The fixtures are what travels to the GPU box (no /root/reference there).
"""
import gzip
import hashlib
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref")
SEG_DIR = os.path.join(HERE, "seg")


def store_dump(path):
    with open(path, "rb") as f:
        data = f.read()
    name = hashlib.sha1(data).hexdigest()[:16] + ".bin.gz"
    dst = os.path.join(SEG_DIR, name)
    if not os.path.exists(dst):
        with open(dst, "wb") as f:
            f.write(gzip.compress(data, 9, mtime=0))
    return name


def from_gtest_capture(run=True):
    if run:
        subprocess.run([os.path.join(ROOT, "oracle", "run_capture.sh")], check=True)
    cap = os.path.join(REF, "capture")
    out, out_sel = [], []
    for line in open(os.path.join(cap, "capture.jsonl")):
        r = json.loads(line)
        q = r["query"]
        if q.get("type") not in ("aggregate", "select", "search"):
            continue
        tables = [t for t in r["db"].get("tables", []) if t["name"] == q.get("table")]
        if not tables:
            continue
        rec = {"test": r["test"], "seq": r["seq"], "table": tables[0], "query": q}
        if "rollup_ts" in r:
            rec["rollup_ts"] = int(r["rollup_ts"].rstrip("L"))
        if "error" in r:
            rec["error"], rec["error_type"] = r["error"], r["error_type"]
        else:
            rec["rows"], rec["stats"] = r["rows"], r["stats"]
        if "dump" in r:
            rec["seg"] = store_dump(os.path.join(cap, r["dump"]))
        (out if q.get("type") == "aggregate" else out_sel).append(rec)
    for name, recs in (("ref_gtest.jsonl", out), ("ref_gtest_select.jsonl", out_sel)):
        with open(os.path.join(HERE, name), "w") as f:
            for rec in recs:
                f.write(json.dumps(rec, sort_keys=True) + "\n")
    print(f"ref_gtest.jsonl: {len(out)} aggregate queries, ref_gtest_select.jsonl: {len(out_sel)} select / search "
          f"queries from the reference's gtests")


def _run_job(jpath):
    return subprocess.run([os.path.join(REF, "oracle_cli"), jpath], capture_output=True, text=True)


def from_scenarios(which="SCENARIOS", target="ref_scenarios.jsonl", module="scenarios", workers=1, state_dir=None):
    sys.path.insert(0, HERE)
    import importlib
    from concurrent.futures import ThreadPoolExecutor
    scenarios = importlib.import_module(module)
    tmp = os.path.join(REF, "scenario_tmp")
    os.makedirs(tmp, exist_ok=True)
    out = []
    all_sc = getattr(scenarios, which)
    jobs = []
    for sc in all_sc:
        job = {"state_dir": state_dir or os.path.join(REF, "state"), "table": sc["table"], "queries": sc["queries"],
               "dump": os.path.join(tmp, sc["name"] + ".bin")}
        for k in ("rows", "generate", "rollup_ts"):
            if k in sc:
                job[k] = sc[k]
        jpath = os.path.join(tmp, sc["name"] + ".json")
        with open(jpath, "w") as f:
            json.dump(job, f)
        jobs.append((job, jpath))
    # one reference process per scenario (each distinct query is a g++ JIT compile): several at a time
    with ThreadPoolExecutor(max_workers=workers) as ex:
        procs = list(ex.map(_run_job, [jp for _, jp in jobs]))
    for sc, (job, jpath), p in zip(all_sc, jobs, procs):
        if p.returncode != 0:
            raise RuntimeError(f"oracle_cli failed on {sc['name']}: {p.stdout[-500:]} {p.stderr[-500:]}")
        res = json.loads(p.stdout.strip().splitlines()[-1])
        seg = store_dump(job["dump"])
        for qi, (q, r) in enumerate(zip(sc["queries"], res["results"])):
            rec = {"test": f"{sc['name']}.{qi}", "table": sc["table"], "query": q, "seg": seg}
            if "rollup_ts" in sc:
                rec["rollup_ts"] = sc["rollup_ts"]
            if "error" in r:
                rec["error"] = r["error"]
            else:
                rec["rows"], rec["stats"] = r["rows"], r["stats"]
            out.append(rec)
        print(f"  {sc['name']}: {res['stored_rows']} rows in {res['segments']} segments, {len(sc['queries'])} queries")
    with open(os.path.join(HERE, target), "w") as f:
        for rec in out:
            f.write(json.dumps(rec, sort_keys=True) + "\n")
    print(f"{target}: {len(out)} queries")


if __name__ == "__main__":
    os.makedirs(SEG_DIR, exist_ok=True)
    what = sys.argv[1:] or ["gtest", "scenarios"]
    if "gtest" in what:
        from_gtest_capture(run="--no-run" not in what)
    if "scenarios" in what:
        from_scenarios()
    if "select" in what or not sys.argv[1:]:
        from_scenarios("SELECT_SCENARIOS", "ref_select_scenarios.jsonl")
    if "edge" in what or not sys.argv[1:]:
        from_scenarios("EDGE_SCENARIOS", "ref_edge_scenarios.jsonl")
    if "fuzz" in what or not sys.argv[1:]:
        # a JIT cache of its own: these ~250 one-off .so files must not travel to the GPU box with oracle/_ref/state
        from_scenarios("FUZZ_SCENARIOS", "ref_fuzz_scenarios.jsonl", module="fuzz_scenarios", workers=os.cpu_count() or 1,
                       state_dir=os.path.join("/tmp", "vgpu_fuzz_state"))
    if "fuzz_select" in what or not sys.argv[1:]:
        from_scenarios("FUZZ_SELECT_SCENARIOS", "ref_fuzz_select_scenarios.jsonl", module="fuzz_scenarios",
                       workers=os.cpu_count() or 1, state_dir=os.path.join("/tmp", "vgpu_fuzz_state"))
