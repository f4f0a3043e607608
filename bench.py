#!/usr/bin/env python
"""bench.py — scanned rows/s of the scan -> filter -> group-by-aggregate hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2] [--impl reference]

A step = one pass of the hot path (vgpu_query_agg through the C ABI: prune, fused scan kernel,
count-distinct dedupe, NCCL merge when N > 1, group extraction, groups copied to the host) over the
whole resident table. One process per GPU (torchrun for N > 1); segments are sharded round-robin
across ranks (weak scaling: the per-GPU table is fixed), the only data-path collective is the merge
of the per-GPU partial group tables.

Workloads (synthetic, splitmix64(seed=42), written directly in HBM):
  c2 (default) 1e9 rows/GPU: d0 IN(5) AND n4 in [250,750) AND t5 in [T+2.5e6,T+7.5e6); group (d0,d1,d2,d3);
               min, max, count-distinct            -> the 60 %-of-roofline target config of BASELINE.json
  c0           1e7 rows: dim2 == code; group (dim1); sum        -> the reference's own CPU-runnable config
  c1           1e8 rows/GPU: d0 == code; group (d1,d2); sum, count
  c3           1e9 rows/GPU: time range; group (d0,d1,d2); sum
  c4           1e9 rows/GPU: no filter; group (d0, hour/day/month rollup of t1); sum, count (1e7 groups)

`--impl reference` times the reference's own CPU implementation (oracle/_ref/oracle_cli = the
unmodified ViyaDB sources, JIT .so cache pre-warmed in the authoring container) on the box's host
cores, as P shared-nothing single-threaded shards (the reference's own scale-out model).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T0 = 1490000000
NOW = 1496570140
SEG = 1_000_000

# per workload: table config, generator (lo, range[, mode, div]) per schema column, query, per-row bytes
WORKLOADS = {
    # C0: the reference's own CPU-runnable config (BASELINE.json configs[0]): the whole table also runs on the CPU arm
    "c0": {
        "rows": 10_000_000,
        "table": {"name": "events", "segment_size": SEG,
                  "dimensions": [{"name": "dim1"}, {"name": "dim2"}, {"name": "dim3"}],
                  "metrics": [{"name": "m1", "type": "long_sum"}, {"name": "m2", "type": "int_sum"}]},
        "gens": [(1, 1000), (1, 100), (1, 100), (0, 977), (0, 13)],
        "prefix": ["a", "b", "c", "", ""],
        "query": {"type": "aggregate", "table": "events", "dimensions": ["dim1"], "metrics": ["m1"],
                  "filter": {"op": "eq", "column": "dim2", "value": "b7"}},
        "filter_bytes": 4, "payload_bytes": 4 + 8, "full_bytes": 24, "dtype": "int64",
    },
    "c1": {
        "rows": 100_000_000,
        "table": {"name": "events", "segment_size": SEG,
                  "dimensions": [{"name": "d0"}, {"name": "d1"}, {"name": "d2"}, {"name": "d3"}],
                  "metrics": [{"name": "count", "type": "count"}, {"name": "m1", "type": "long_sum"},
                              {"name": "m2", "type": "int_sum"}]},
        "gens": [(1, 16), (1, 1000), (1, 100), (1, 1_000_000), (1, 1), (0, 1000), (-50, 100)],
        "prefix": ["a", "b", "c", "d", "", "", ""],
        "query": {"type": "aggregate", "table": "events", "dimensions": ["d1", "d2"], "metrics": ["m1", "count"],
                  "filter": {"op": "eq", "column": "d0", "value": "a7"}},
        "filter_bytes": 4, "payload_bytes": 4 + 4 + 8 + 4, "full_bytes": 32, "dtype": "int64",
    },
    "c2": {
        "rows": 1_000_000_000,
        "table": {"name": "events", "segment_size": SEG,
                  "dimensions": [{"name": "d0"}, {"name": "d1"}, {"name": "d2"}, {"name": "d3"},
                                 {"name": "n4", "type": "ushort"}, {"name": "t5", "type": "time"}],
                  "metrics": [{"name": "mn", "type": "int_min"}, {"name": "mx", "type": "int_max"},
                              {"name": "uid", "type": "bitset"}]},
        "gens": [(1, 50), (1, 20), (1, 50), (1, 100), (0, 1000), (T0, 10_000_000),
                 (-2**31, 2**32), (-2**31, 2**32), (0, 1_000_000)],
        "prefix": ["a", "b", "c", "d", "", "", "", "", ""],
        "query": {"type": "aggregate", "table": "events", "dimensions": ["d0", "d1", "d2", "d3"],
                  "metrics": ["mn", "mx", "uid"],
                  "filter": {"op": "and", "filters": [
                      {"op": "in", "column": "d0", "values": ["a3", "a11", "a19", "a27", "a42"]},
                      {"op": "ge", "column": "n4", "value": "250"}, {"op": "lt", "column": "n4", "value": "750"},
                      {"op": "ge", "column": "t5", "value": str(T0 + 2_500_000)},
                      {"op": "lt", "column": "t5", "value": str(T0 + 7_500_000)}]}},
        "filter_bytes": 4 + 2 + 4, "payload_bytes": 12 + 4 + 4 + 4, "full_bytes": 34, "dtype": "int32",
    },
    "c3": {
        "rows": 1_000_000_000,
        "table": {"name": "events", "segment_size": SEG,
                  "dimensions": [{"name": "d0"}, {"name": "d1"}, {"name": "d2"}, {"name": "t3", "type": "time"}],
                  "metrics": [{"name": "m1", "type": "long_sum"}]},
        "gens": [(1, 100), (1, 100), (1, 100), (T0, 4_000_000), (0, 1000)],
        "prefix": ["a", "b", "c", "", ""],
        "query": {"type": "aggregate", "table": "events", "dimensions": ["d0", "d1", "d2"], "metrics": ["m1"],
                  "filter": {"op": "and", "filters": [{"op": "ge", "column": "t3", "value": str(T0 + 1_000_000)},
                                                      {"op": "lt", "column": "t3", "value": str(T0 + 2_000_000)}]}},
        "filter_bytes": 4, "payload_bytes": 12 + 8, "full_bytes": 24, "dtype": "int64",
    },
    "c4": {
        "rows": 1_000_000_000,
        "table": {"name": "events", "segment_size": SEG,
                  "dimensions": [{"name": "d0"},
                                 {"name": "t1", "type": "time",
                                  "rollup_rules": [{"granularity": "hour", "after": "1 days"},
                                                   {"granularity": "day", "after": "1 weeks"},
                                                   {"granularity": "month", "after": "1 years"}]}],
                  "metrics": [{"name": "m1", "type": "long_sum"}, {"name": "count", "type": "count"}]},
        "gens": [(1, 20000), (NOW - 730 * 86400, 730 * 86400), (0, 1000), (1, 1)],
        "prefix": ["a", "", "", ""],
        "query": {"type": "aggregate", "table": "events",
                  "select": [{"column": "d0"}, {"column": "t1", "granularity": "hour"}, {"column": "m1"},
                             {"column": "count"}]},
        "filter_bytes": 0, "payload_bytes": 4 + 4 + 8 + 4, "full_bytes": 20, "dtype": "int64",
    },
}


def load_traffic(wname):
    """DRAM bytes per launch of the scan kernel from the last `ncu --set full` capture of this workload
    (profiles/traffic.json, written by tools/save_profile.sh; never measured inside a bench run)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[wname]
        return float(t["bytes_per_launch"]), t.get("source")
    except Exception:
        return None, None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t_begin, t_end):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                clk = float(f[1])
                mx = float(f[2])
            except ValueError:
                continue
            if t_begin - 0.05 <= ts <= t_end + 0.05 or not sm:
                sm.append(clk)
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# the reference CPU arm
# ------------------------------------------------------------------------------------------------
def oracle_cli_path():
    return os.path.join(ROOT, "oracle", "_ref", "oracle_cli")


def reference_job(wname, rows, row_offset, repeat, dump=None):
    w = WORKLOADS[wname]
    cols = []
    table = w["table"]
    names = [d["name"] for d in table["dimensions"]] + [m["name"] for m in table["metrics"]]
    for (name, g, prefix) in zip(names, w["gens"], w["prefix"]):
        mt = next((m for m in table["metrics"] if m["name"] == name), None)
        if mt is not None and mt["type"] == "count":
            continue  # count is not an input column
        cols.append({"prefix": prefix, "lo": g[0], "range": g[1]})
    job = {"state_dir": os.path.join(ROOT, "oracle", "_ref", "state"), "rollup_ts": NOW, "table": table,
           "generate": {"n": rows, "seed": 42, "row_offset": row_offset, "columns": cols},
           "queries": [w["query"]], "repeat": repeat}
    if dump:
        job["dump"] = dump
    return job


def run_reference(wname, rows_per_proc, procs, warmup, steps, want_rows=False, dump=None):
    """P shared-nothing single-threaded reference processes, each owning rows_per_proc rows, all
    running the same query concurrently. Returns (rows/s, ms_per_step, total_rows, detail); with want_rows also the
    formatted result rows of shard 0 (its rows are generator rows [0, rows_per_proc))."""
    cli = oracle_cli_path()
    if not os.path.exists(cli):
        return None
    tmp = tempfile.mkdtemp(prefix="vgpu_ref_")
    ps = []
    for p in range(procs):
        job = reference_job(wname, rows_per_proc, p * rows_per_proc, warmup + steps, dump if p == 0 else None)
        jp = os.path.join(tmp, f"job{p}.json")
        json.dump(job, open(jp, "w"))
        ps.append(subprocess.Popen([cli, jp], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    per_proc, scanned, rows0 = [], 0, None
    for pi, p in enumerate(ps):
        out, err = p.communicate()
        lines = out.strip().splitlines()
        if p.returncode != 0 or not lines:
            return {"error": (lines[-1] if lines else err)[-300:]}
        res = json.loads(lines[-1])
        r = res["results"][0]
        if "error" in r:
            return {"error": r["error"][-300:]}
        per_proc.append([t["whole_ms"] - t["compile_ms"] for t in r["timings"]][warmup:])
        scanned += r["stats"]["scanned_recs"]
        if pi == 0 and want_rows:
            rows0 = r["rows"]
    step_ms = [max(pp[i] for pp in per_proc) for i in range(steps)]
    ms = sum(step_ms) / len(step_ms)
    return {"value": scanned / (ms / 1e3), "ms_per_step": ms, "rows": scanned, "best_ms": min(step_ms), "rows0": rows0}


def main_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    procs = os.cpu_count() or 1
    rows_per_proc = args.ref_rows // procs
    r = run_reference(args.workload, rows_per_proc, procs, args.warmup, args.steps)
    r1 = run_reference(args.workload, rows_per_proc, 1, 1, 2)   # the reference's real per-query execution: one thread
    w = WORKLOADS[args.workload]
    line = {"impl": "reference", "metric": "scanned rows/sec", "unit": "rows/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": w["dtype"], "data": "synthetic",
            "config": {"workload": f"{args.workload} ({describe(args.workload)})", "rows": None}}
    if r is None or "error" in (r or {}):
        line["unavailable"] = "oracle/_ref/oracle_cli missing" if r is None else "reference run failed: " + r["error"]
        print(json.dumps(line))
        return 0
    sample = f"{r['rows']} rows of {args.workload} as {procs} shared-nothing single-threaded reference shards"
    line.update({"value": r["value"], "ms_per_step": r["ms_per_step"],
                 "cpu_baseline": {"value": r["value"], "unit": "rows/s", "cores": procs, "kind": "reference", "sample": sample,
                                  "value_1core": (r1 or {}).get("value"),
                                  "sample_1core": f"{rows_per_proc} rows, one single-threaded reference process"},
                 "e2e": {"value": r["value"], "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    line["config"]["rows"] = r["rows"]
    print(json.dumps(line))
    return 0


def describe(wname):
    return {"c0": "1e7 rows, 3 string dims + 2 int metrics, SELECT dim1,SUM(m1) WHERE dim2='x' GROUP BY dim1",
            "c1": "1e8 rows/GPU, eq filter, 2-key group-by, sum+count",
            "c2": "1e9 rows/GPU, IN + 2 range predicates over 3 dims, 4-key group-by, min/max + count-distinct",
            "c3": "1e9 rows/GPU, time-range filter, 3-key group-by, sum",
            "c4": "1e9 rows/GPU, no filter, (dim, rolled-up time) group-by ~1e7 groups, sum+count"}[wname]


# ------------------------------------------------------------------------------------------------
# the CUDA arm
# ------------------------------------------------------------------------------------------------
def parity_check(v, dist, rank, world, local, wname, check_rows):
    """The CUDA path against the REAL reference on the same table, inside the bench, at every N: rank 0 runs one
    reference process (oracle/_ref/oracle_cli, the unmodified ViyaDB sources) that ingests generator rows
    [0, check_rows), answers the query and dumps its segments; all ranks upload those very segments — cut into pieces,
    sharded round-robin and merged over NCCL like the timed table — and the formatted result rows must be identical
    as sorted sets."""
    w = WORKLOADS[wname]
    ref, dump = None, [None]
    if rank == 0:
        dump[0] = os.path.join(tempfile.mkdtemp(prefix="vgpu_check_"), "segments.bin")
        ref = run_reference(wname, check_rows, 1, 0, 1, want_rows=True, dump=dump[0])
        if ref is None or "error" in ref:
            dump[0] = None
    if dist is not None:
        dist.broadcast_object_list(dump, src=0)
    if dump[0] is None:
        return {"rows": check_rows, "ok": None,
                "error": "reference unavailable: " + str((ref or {}).get("error", "oracle_cli missing"))} if rank == 0 else None
    db = v.Database({"tables": [dict(w["table"])]}, device=local)
    try:
        if world > 1:
            uid = [v.Database.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            db.init_comm(rank, world, uid[0])
        t = db.get_table("events")
        chunk = max(1, check_rows // (2 * world + 1))          # ragged on purpose
        t.load_dump(dump[0], shard=(rank, world), chunk_rows=chunk)
        out = v.MemoryRowOutput()
        stats = db.query(w["query"], out, now=NOW)
    finally:
        db.close()
    if dist is not None:
        dist.barrier()
    if rank != 0:
        return None
    try:
        os.remove(dump[0])
    except OSError:
        pass
    got, want = sorted(out.rows), sorted(ref["rows0"])
    return {"rows": check_rows, "table_rows": stats.scanned_recs, "groups": len(got), "reference_groups": len(want),
            "ok": got == want, "n_gpus": world,
            "against": "unmodified reference (oracle/_ref/oracle_cli): its own segments uploaded, formatted rows compared as sorted sets"}


def run_workload(args, wname, steps, torch, v, dist, rank, world, local, full):
    """One workload on the resident table: returns the JSON line (rank 0) or None."""
    from viyadb_b200 import _native as N
    from viyadb_b200.query import GpuQueryRunner, QueryFactory
    w = WORKLOADS[wname]
    rows = args.rows or w["rows"]
    table_conf = dict(w["table"])
    db = v.Database({"tables": [table_conf]}, device=local)
    lib = N.load()
    stream = torch.cuda.current_stream()
    N.check(lib.vgpu_set_stream(db.ctx, stream.cuda_stream))
    if world > 1:
        uid = [v.Database.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        db.init_comm(rank, world, uid[0])
    t = db.get_table("events")
    # dictionaries: code k <-> "<prefix>k" (code 0 stays "__exceeded")
    for d, g, prefix in zip(t.dimensions, w["gens"], w["prefix"]):
        if d.dict is not None:
            for k in range(1, g[0] + g[1]):
                d.dict.encode(f"{prefix}{k}")
    # segments: global segment s lives on rank s % world
    nseg = (rows + SEG - 1) // SEG
    t_gen = time.time()
    for ls in range(nseg):
        gs = ls * world + rank
        n = min(SEG, rows - ls * SEG)
        t.generate_segment(ls, n, w["gens"], seed=42, row_offset=gs * SEG)
    torch.cuda.synchronize()
    t_gen = time.time() - t_gen

    query = QueryFactory.create(w["query"], db)
    # several GPUs: the merged groups are copied to rank 0's host only (one client gets one answer)
    runner = GpuQueryRunner(db, v.MemoryRowOutput(), now=NOW, flags=N.PLAN_RESULT_ON_ROOT if world > 1 else 0)
    plan = runner.build_plan(query)

    def step():
        return runner.run_plan(query, plan)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        groups = step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    scan_ms, gpu_ms, launches = [], [], 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.cudart().cudaProfilerStart()   # `ncu --profile-from-start off` lists the timed region only
    wall0 = time.time()
    ev0.record(stream)
    for _ in range(steps):
        groups = step()
        scan_ms.append(runner.stats.kernel_scan_ms)
        gpu_ms.append(runner.stats.gpu_ms)
        launches += runner.stats.launches
    ev1.record(stream)
    barrier()
    wall1 = time.time()
    torch.cuda.cudart().cudaProfilerStop()
    elapsed_ms = ev0.elapsed_time(ev1)
    if dist is not None:
        tt = torch.tensor([elapsed_ms, sum(scan_ms) / len(scan_ms), sum(gpu_ms) / len(gpu_ms)], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        elapsed_ms, scan_avg, gpu_avg = tt.tolist()
    else:
        scan_avg = sum(scan_ms) / len(scan_ms)
        gpu_avg = sum(gpu_ms) / len(gpu_ms)
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    total_rows = rows * world
    passed = runner.stats.passed_rows          # all ranks (summed in the merge)
    ngroups = runner.stats.aggregated_recs
    selectivity = passed / max(1, runner.stats.scanned_recs)
    b_alg = w["filter_bytes"] + selectivity * w["payload_bytes"]
    peak, peak_src = load_peaks()
    traffic, traffic_src = load_traffic(wname) if rows == w["rows"] else (None, None)
    alg_bytes = rows * b_alg                   # per GPU and launch
    achieved = alg_bytes / (scan_avg / 1e3) / 1e9
    step_ms = elapsed_ms / steps
    value = total_rows * steps / (elapsed_ms / 1e3)
    paths = runner.stats.distinct_paths

    # ---- post-aggregation: top-100 groups by SUM(m1) over the 1e7 groups of C4 ----
    post = None
    if wname == "c4" and rank == 0 and world == 1 and not args.no_post:
        post = measure_post_agg(v, db, w)

    # ---- e2e: host buffers -> put_segment (H2D from pinned memory) -> query -> groups on host ----
    resident = t.device_bytes
    e2e = None
    if full and not args.no_e2e:
        e2e = measure_e2e(torch, v, db, t, query, runner, plan, rows, nseg, world, dist, stream, args)
    db.close()

    line = None
    if rank == 0:
        line = {
            "metric": "scanned rows/sec", "value": value, "unit": "rows/s", "n_gpus": world, "steps": steps,
            "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": w["dtype"], "data": "synthetic",
            "config": {"workload": f"{wname} ({describe(wname)})", "rows_per_gpu": rows,
                       "segments_per_gpu": nseg, "segment_size": SEG, "groups": ngroups,
                       "selectivity": selectivity, "resident_bytes_per_gpu": resident,
                       "l2_policy": "inputs larger than L2 (table >> 126 MB), no flush needed",
                       "group_table": "dense" if runner.stats.table_mode == 0 else "hash",
                       "count_distinct_path": [n for b, n in ((1, "one global set"), (2, "hash buckets + shared-memory sets"),
                                                              (4, "16-byte pairs, global set"), (8, "L2 partitions + global sets"),
                                                              (16, "fast path overflowed, redone")) if paths & b],
                       "parallelism": f"segment-sharded x{world}, one NCCL merge of partial group tables" if world > 1 else "single GPU"},
            # three fractions of the same algorithmic bytes: over the fused scan kernel alone, over the
            # device time of the whole query (scan + count-distinct dedupe + merge + extraction, first to last
            # operation on its stream), and over the wall-clock step (what `value` is computed from)
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "frac_query_device": alg_bytes / (gpu_avg / 1e3) / 1e9 / peak,
                         "frac_step": alg_bytes / (step_ms / 1e3) / 1e9 / peak,
                         "traffic": traffic, "traffic_unit": "bytes/launch (dram__bytes_read.sum + dram__bytes_write.sum)",
                         "traffic_source": traffic_src, "kernel": "scan_filter_groupby_kernel", "kernel_ms": scan_avg,
                         "algorithmic_bytes_per_row": b_alg, "algorithmic_bytes_per_launch": alg_bytes,
                         "referenced_bytes_per_row": w["filter_bytes"] + w["payload_bytes"],   # B_full of SURVEY 8d
                         "table_bytes_per_row": w["full_bytes"],
                         "frac_of_datasheet_8TBs": achieved / 8000.0,
                         "peak_source": peak_src},
            "gpu_launches": launches, "gpu_ms_per_step": gpu_avg,
            "clocks": clocks, "e2e": e2e, "generate_s": t_gen,
        }
        if post is not None:
            line["post_agg"] = post
    return line


def main_ours(args):
    import torch
    import viyadb_b200 as v

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: viyadb_b200 has no CPU path")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    line = run_workload(args, args.workload, args.steps, torch, v, dist, rank, world, local, full=True)

    # ---- the same path against the real reference on the same rows, at every N ----
    check = None
    if not args.no_check:
        check = parity_check(v, dist, rank, world, local, args.workload, args.check_rows)

    # ---- the other named configurations, driver-witnessed at every N (shorter runs) ----
    also = {}
    if args.workload == "c2" and not args.no_also and not args.rows:
        for wname in ("c3", "c4"):
            r = run_workload(args, wname, max(3, args.steps // 4), torch, v, dist, rank, world, local, full=False)
            if r is not None:
                also[wname] = {k: r[k] for k in ("value", "unit", "ms_per_step", "n_gpus", "steps", "gpu_launches", "gpu_ms_per_step")}
                if "post_agg" in r:
                    also[wname]["post_agg"] = r["post_agg"]
                also[wname]["config"] = r["config"]
                also[wname]["roofline"] = {k: r["roofline"][k] for k in ("frac", "frac_query_device", "frac_step", "kernel_ms", "achieved",
                                                                           "algorithmic_bytes_per_row", "algorithmic_bytes_per_launch",
                                                                           "traffic", "traffic_source")}
            if not args.no_check:
                c = parity_check(v, dist, rank, world, local, wname, min(args.check_rows, 500_000))
                if r is not None:
                    also[wname]["parity_check"] = c

    # ---- CPU baseline (rank 0, N=1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        procs = os.cpu_count() or 1
        r = run_reference(args.workload, args.ref_rows // procs, procs, 1, 3)
        r1 = run_reference(args.workload, args.ref_rows // procs, 1, 1, 2)
        if r and "error" not in r:
            cpu = {"value": r["value"], "unit": "rows/s", "cores": procs, "kind": "reference",
                   "value_1core": (r1 or {}).get("value"),
                   "sample": f"{r['rows']} rows of {args.workload} as {procs} shared-nothing single-threaded shards of "
                             f"the unmodified reference (oracle/_ref/oracle_cli), query time only, mean of 3"}
        else:
            cpu = {"value": None, "unit": "rows/s", "cores": procs, "kind": "reference",
                   "sample": "unavailable: " + ("oracle_cli missing" if r is None else r["error"])}

    if rank == 0:
        line["parity_check"] = check
        line["cpu_baseline"] = cpu
        if also:
            line["also"] = also
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def measure_post_agg(v, db, w):
    """t_post of a top-N query over C4's ~1e7 groups: Database.query wall time minus the vgpu_query_agg call, with HAVING /
    top-N on the device (the host formats and sorts the survivors) and with both on the host. The host-only leg runs on a
    slice of the key space (d0 IN 60 of 20000 values: formatting 1e7 rows in the Python mirror takes minutes) and is
    scaled linearly — an estimate (a lower bound: the sort is n log n), labelled as such."""
    q = dict(w["query"], sort=[{"column": "m1"}, {"column": "d0", "ascending": True}], limit=100)
    out = {"query": "C4 + sort by SUM(m1) desc, d0 asc, limit 100"}
    for name, dev in (("device", True), ("host", False)):
        qq = dict(q)
        if not dev:
            qq["filter"] = {"op": "in", "column": "d0", "values": [f"a{k}" for k in range(1, 61)]}
        best = None
        for _ in range(2):
            rows_out = v.MemoryRowOutput()
            t0 = time.time()
            st = db.query(qq, rows_out, now=NOW, device_post=dev)
            whole = (time.time() - t0) * 1e3
            rec = {"whole_ms": whole, "scan_call_ms": st.scan_time * 1e3, "t_post_ms": whole - st.scan_time * 1e3,
                   "groups": st.aggregated_recs, "output_recs": st.output_recs, "post_applied": st.post_applied,
                   "first_row": rows_out.rows[0] if rows_out.rows else None}
            if best is None or rec["t_post_ms"] < best["t_post_ms"]:
                best = rec
        out[name] = best
    h = out["host"]
    h["t_post_ms_scaled_to_all_groups"] = h["t_post_ms"] * out["device"]["groups"] / max(1, h["groups"])
    h["note"] = "host-only leg on 60 of the 20000 d0 values, t_post scaled linearly to the full group count (estimate, lower bound)"
    out["t_post_ratio_estimate"] = h["t_post_ms_scaled_to_all_groups"] / max(1e-9, out["device"]["t_post_ms"])
    return out


def measure_e2e(torch, v, db, t, query, runner, plan, rows, nseg, world, dist, stream, args):
    """Same metric through the public API with HOST buffers: every step copies every column of every
    segment from pinned host memory into HBM (Table.put_segment -> vgpu_segment_put), runs the query
    and reads the groups back. PCIe-bound by construction."""
    import ctypes as C
    from viyadb_b200 import _native as N
    lib = N.load()
    # N = 1: the whole table is re-uploaded every step. N > 1: a bounded share per rank (default 2.5e8 rows),
    # so that the pinned host buffers of all ranks together stay modest; the table is cut down to exactly the
    # uploaded segments first, so rows scanned == rows uploaded.
    e_rows = min(rows, args.e2e_rows) if args.e2e_rows else (rows if world == 1 else min(rows, 250_000_000))
    # pinned buffers close to the GPU: bind this process to the GPU's NUMA-local cores before allocating them (and give
    # the cores back afterwards: the cpu_baseline leg that follows must see every core of the box)
    saved_affinity = os.sched_getaffinity(0)
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device()))
        numa = "cpu affinity set to the GPU's NUMA-local cores (nvmlDeviceSetCpuAffinity)"
    except Exception as e:   # noqa: BLE001
        numa = "cpu affinity not set: " + str(e)[:80]
    e_nseg = (e_rows + SEG - 1) // SEG
    e_rows = min(rows, e_nseg * SEG)
    cols = t.dimensions + t.metrics
    pinned, views = {}, {}
    h2d = 0
    for c in cols:
        width = 4 if c.kind == N.METRIC_BITSET else N.TYPE_WIDTH[c.type]
        buf = torch.empty(e_rows * width, dtype=torch.uint8, pin_memory=True)
        pinned[c.name] = buf
        views[c.name] = buf.numpy().view("<u4" if c.kind == N.METRIC_BITSET else N.NP_DTYPES[c.type])
        h2d += e_rows * width
        for s in range(e_nseg):   # fill the host buffers from the generated table (device -> pinned host)
            n = min(SEG, e_rows - s * SEG)
            N.check(lib.vgpu_segment_read(t.handle, s, t.schema_index(c), C.c_void_p(buf.data_ptr() + s * SEG * width)))
    host = []
    for s in range(e_nseg):
        n = min(SEG, e_rows - s * SEG)
        host.append(t.prepare_segment({c.name: views[c.name][s * SEG:s * SEG + n] for c in cols}))   # pointer marshalling only

    for s in range(e_nseg, nseg):
        t.invalidate(s)

    def step():
        # copies are enqueued back to back (vgpu_segment_put_async): the DMA engine never waits for the host; the query is
        # ordered after them on the device, and its result reaching the host means every copy has finished
        for s, seg in enumerate(host):
            t.put_prepared(s, seg, wait=False)
        g = runner.run_plan(query, runner.build_plan(query))
        t.sync()
        return g

    for _ in range(2):
        g = step()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    k = 3
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(k):
        g = step()
    ev1.record(stream)
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / k
    if dist is not None:
        tt = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = tt.item()
    os.sched_setaffinity(0, saved_affinity)
    d2h = sum(a.nbytes for a in g["keys"]) + sum(a.nbytes for a in g["accs"])
    scanned = runner.stats.scanned_recs   # all ranks (QueryStats are merged): the tables hold exactly the uploaded segments
    assert scanned == e_rows * world, (scanned, e_rows, world)
    return {"value": scanned / (ms / 1e3), "unit": "rows/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "rows_uploaded_per_gpu_per_step": e_rows, "rows_scanned_per_step_all_gpus": scanned, "ms_per_step": ms, "steps": k,
            "h2d_gbs_per_gpu": h2d / (ms / 1e3) / 1e9, "numa": numa,
            "note": "every step re-uploads all columns from pinned host memory (vgpu_segment_put_async, one DMA per column "
                    "and segment, no host wait in between), then runs the query and copies the groups back; PCIe-bound"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--rows", type=int, default=0, help="rows per GPU (default: the workload's)")
    ap.add_argument("--e2e-rows", type=int, default=0, help="rows per GPU re-uploaded per e2e step (0 = the whole table)")
    ap.add_argument("--ref-rows", type=int, default=16_000_000, help="total rows of the bounded CPU sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-check", action="store_true", help="skip the in-bench parity check against the reference")
    ap.add_argument("--no-post", action="store_true", help="skip the post-aggregation (top-N) measurement of the c4 run")
    ap.add_argument("--no-also", action="store_true", help="skip the extra c3 / c4 lines of the default run")
    ap.add_argument("--check-rows", type=int, default=1_000_000, help="rows of the in-bench parity check")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return main_reference(args)
    return main_ours(args)


if __name__ == "__main__":
    sys.exit(main())
