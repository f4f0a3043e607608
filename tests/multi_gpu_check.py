#!/usr/bin/env python
"""Run under torchrun on N GPUs: every workload's N-GPU result (segments sharded round-robin, NCCL merge) must equal,
on every rank, (1) the ORACLE's result (oracle/viya_oracle.py, pinned to the reference by tests/test_oracle_golden.py)
over the union of all segments — formatted rows as sorted sets and all four QueryStats counters — and (2) the
single-GPU result of the library, bit for bit. Prints MULTI_GPU_OK on rank 0."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bench
import viya_oracle
import viyadb_b200 as v
from viyadb_b200 import dist as vdist
from viyadb_b200.query import GpuQueryRunner, QueryFactory

SEG = 50_000


def run(db, w, q, flags=0):
    query = QueryFactory.create(q, db)
    r = GpuQueryRunner(db, v.MemoryRowOutput(), now=bench.NOW, flags=flags)
    g = r.run_plan(query, r.build_plan(query))
    keys = [np.array(k) for k in g["keys"]]
    accs = [np.array(a) for a in g["accs"]]
    order = np.lexsort(tuple(reversed(keys))) if keys else np.arange(g["ngroups"])
    return [k[order] for k in keys], [a[order] for a in accs], r.stats


def main():
    rank, world, local = vdist.world()
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nseg_global = 3 * world + 1          # ragged on purpose: ranks own different numbers of segments
    # low-cardinality variants: the group table is CTA-private in shared memory and merged into per-metric
    # arrays (the multi-GPU layout) at the end of the scan
    low = {"c1low": dict(bench.WORKLOADS["c1"], query=dict(bench.WORKLOADS["c1"]["query"], dimensions=["d0"], metrics=["m1", "m2", "count"])),
           "c2low": dict(bench.WORKLOADS["c2"], query=dict(bench.WORKLOADS["c2"]["query"], dimensions=["d1"], metrics=["mn", "mx", "uid"])),
           "c0": bench.WORKLOADS["c0"],
           # key tuple wider than 64 bits (float + double + code): wide-tuple hash tables, merged by gathering every rank's records
           "wide": {"table": {"name": "events", "segment_size": SEG,
                              "dimensions": [{"name": "f", "type": "float"}, {"name": "g", "type": "double"}, {"name": "d0"}],
                              "metrics": [{"name": "count", "type": "count"}, {"name": "m1", "type": "long_sum"},
                                          {"name": "mx", "type": "int_max"}]},
                    "gens": [(-3, 7), (-40, 90), (1, 30), (1, 1), (0, 1000), (-2**31, 2**32)], "prefix": ["", "", "a", "", "", ""],
                    "query": {"type": "aggregate", "table": "events", "dimensions": ["f", "g", "d0"], "metrics": ["m1", "count", "mx"],
                              "filter": {"op": "gt", "column": "m1", "value": "100"}}}}
    only = os.environ.get("VGPU_CHECK_ONLY")
    for wname, flags in (("c1", 0), ("c2", 2), ("c3", 0), ("c4", 0), ("c1", 1), ("c3", 1), ("c2", 1), ("c4", 1), ("c1low", 0), ("c2low", 0),
                         ("c0", 0), ("wide", 0)):
        if only and wname != only:
            continue
        w = low.get(wname) or bench.WORKLOADS[wname]
        conf = dict(w["table"], segment_size=SEG)
        # sharded database, attached to the communicator
        db = v.Database({"tables": [conf]}, device=local)
        vdist.init_database_comm(db, dist)
        t = db.get_table("events")
        for ls, gs in enumerate(vdist.segments_for_rank(nseg_global, rank, world)):
            n = SEG if gs < nseg_global - 1 else SEG // 3
            t.generate_segment(ls, n, w["gens"], seed=42, row_offset=gs * SEG)
        for d, g, prefix in zip(t.dimensions, w["gens"], w["prefix"]):
            if d.dict is not None:
                for k in range(1, min(g[0] + g[1], 20001)):
                    d.dict.encode(f"{prefix}{k}")
        # for the IN/eq literals to exist in the dictionaries of c1/c2 the prefixes above suffice
        keys, accs, stats = run(db, w, w["query"], flags)
        # reference: one GPU, all segments, no communicator
        db1 = v.Database({"tables": [conf]}, device=local)
        t1 = db1.get_table("events")
        for gs in range(nseg_global):
            n = SEG if gs < nseg_global - 1 else SEG // 3
            t1.generate_segment(gs, n, w["gens"], seed=42, row_offset=gs * SEG)
        for d, d1 in zip(t.dimensions, t1.dimensions):
            d1.dict = d.dict
        keys1, accs1, stats1 = run(db1, w, w["query"], flags)
        assert len(keys) == len(keys1) and all(np.array_equal(a, b) for a, b in zip(keys, keys1)), (wname, "keys differ")
        for mi, (a, b) in enumerate(zip(accs, accs1)):
            if not np.array_equal(a, b):
                bad = np.nonzero(a != b)[0]
                if rank == 0:
                    cells = [int(k.astype(np.int64)[bad[0]]) for k in keys]
                    print(f"[rank 0] bad group index range {bad.min()}..{bad.max()} (keys of first: {cells}); runs of consecutive bad: "
                          f"{int((np.diff(bad) == 1).sum())}", flush=True)
                print(f"[rank {rank}] {wname} flags={flags}: aggregate {mi} differs in {len(bad)} of {len(a)} groups; sums {int(a.astype(np.int64).sum())} vs "
                      f"{int(b.astype(np.int64).sum())}; first {a[bad[:5]].tolist()} vs {b[bad[:5]].tolist()}; paths {stats.distinct_paths} attempts {stats.attempts}", flush=True)
        assert all(np.array_equal(a, b) for a, b in zip(accs, accs1)), (wname, "aggregates differ")
        assert stats.aggregated_recs == stats1.aggregated_recs
        for k in ("scanned_recs", "scanned_segments", "passed_rows"):
            assert getattr(stats, k) == getattr(stats1, k), (wname, k, getattr(stats, k), getattr(stats1, k))
        # (1) against the oracle: the union table read back from the device, formatted rows + QueryStats
        segs = []
        for gs in range(nseg_global):
            n = SEG if gs < nseg_global - 1 else SEG // 3
            seg = {}
            for c in t1.dimensions + t1.metrics:
                col = t1.read_column(gs, c, n)
                seg[c.name] = (np.arange(n + 1, dtype="<u8"), col.astype("<u8")) if c.kind == v._native.METRIC_BITSET else col
            segs.append(seg)
        dicts = {d.name: list(d.dict.c2v) for d in t1.dimensions if d.dict is not None}
        want = viya_oracle.run_query(conf, segs, dicts, w["query"], now=bench.NOW)
        out = v.MemoryRowOutput()
        st = db.query(w["query"], out, now=bench.NOW, flags=flags)
        assert sorted(out.rows) == sorted(want["rows"]), (wname, flags, "rows differ from the oracle")
        for k in ("scanned_segments", "scanned_recs", "aggregated_recs", "output_recs"):
            assert getattr(st, k) == want["stats"][k], (wname, k, getattr(st, k), want["stats"][k])
        if rank == 0:
            print(f"{wname} flags={flags}: {stats.aggregated_recs} groups identical on {world} GPUs "
                  f"(table mode {stats.table_mode})", flush=True)
        db.close()
        db1.close()
    dist.barrier()
    if rank == 0:
        print("MULTI_GPU_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
