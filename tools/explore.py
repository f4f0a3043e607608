#!/usr/bin/env python
"""tools/explore.py — time query variants on the C2 table to attribute scan-kernel time to its parts."""
import copy, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import viyadb_b200 as v
from viyadb_b200.query import GpuQueryRunner, QueryFactory

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000_000
wname = sys.argv[2] if len(sys.argv) > 2 else "c2"
w = bench.WORKLOADS[wname]
db = v.Database({"tables": [w["table"]]}, device=0)
t = db.get_table("events")
for d, g, prefix in zip(t.dimensions, w["gens"], w["prefix"]):
    if d.dict is not None:
        for k in range(1, min(g[0] + g[1], 2000)):
            d.dict.encode(f"{prefix}{k}")
nseg = (rows + bench.SEG - 1) // bench.SEG
for s in range(nseg):
    t.generate_segment(s, min(bench.SEG, rows - s * bench.SEG), w["gens"], seed=42, row_offset=s * bench.SEG)
base = w["query"]
variants = {"full": base}
if wname == "c2":
    f = base["filter"]
    variants.update({
        "d0 | mn": dict(base, dimensions=["d0"], metrics=["mn"]),
        "d0 | mn,mx": dict(base, dimensions=["d0"], metrics=["mn", "mx"]),
        "d0,d1 | mn,mx": dict(base, dimensions=["d0", "d1"], metrics=["mn", "mx"]),
        "d0,d1,d2 | mn,mx": dict(base, dimensions=["d0", "d1", "d2"], metrics=["mn", "mx"]),
        "d1,d2,d3 | mn,mx": dict(base, dimensions=["d1", "d2", "d3"], metrics=["mn", "mx"]),
        "d0..d3 | mn,mx": dict(base, metrics=["mn", "mx"]),
        "d0..d3 | mn": dict(base, metrics=["mn"]),
        "d0..d3 | uid": dict(base, metrics=["uid"]),
        "d0 | mn  [filter d0 only]": dict(base, dimensions=["d0"], metrics=["mn"], filter=f["filters"][0]),
        "d0 | mn  [filter n4 only]": dict(base, dimensions=["d0"], metrics=["mn"], filter={"op": "and", "filters": f["filters"][1:3]}),
        "d0 | mn  [filter never true]": dict(base, dimensions=["d0"], metrics=["mn"],
                                             filter={"op": "and", "filters": f["filters"] + [{"op": "eq", "column": "d1", "value": "nope"}]}),
    })
for name, q in variants.items():
    query = QueryFactory.create(q, db)
    r = GpuQueryRunner(db, v.MemoryRowOutput(), now=bench.NOW)
    plan = r.build_plan(query)
    ms = []
    for i in range(6):
        g = r.run_plan(query, plan)
        ms.append(r.stats.kernel_scan_ms)
    best = min(ms[2:])
    print(f"{name:36s} scan {best:8.3f} ms  {rows/best/1e6:9.1f} Grows/s  passed {r.stats.passed_rows:10d} groups {g['ngroups']:8d}  gpu_ms {r.stats.gpu_ms:.3f}")
db.close()
