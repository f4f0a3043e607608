"""The REAL host planner and the REAL device leaf evaluators on the CPU: viyadb_b200/csrc/planner.h (predicate lowering
to the device program, segment pruning, key-domain tightening — the source vgpu.cu compiles) and csrc/device_arith.h
(leaf_mask16, gen_compare, post_compare — the source the kernels compile) are built with plain g++
(tests/planner_harness.cc) and run over every golden record the real reference answered — its own gtest queries, the
scenario / edge-case runs and the 504 seeded random queries of tests/golden/fuzz_scenarios.py:

  * row predicate: program(row) == the oracle's eval_filter(row) for every row of every segment, through the unrolled
    conjunction path where the planner chooses it AND through the stack interpreter;
  * segment pruning: Planner::process_segment on the statistics the device would reduce == the reference's own
    `scanned_segments` (QueryStats of the golden record) — NOT-IN quirk (Q8) included;
  * HAVING on the device: program(group) == the oracle's HAVING on the oracle's groups;
  * key-domain tightening: no passing row holds a key outside the tightened domain / lookup mask (a violation would
    index outside the dense group table).

Only the two stack machines are restated in the harness (they are entangled with device loads in the kernels)."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

import golden_util as G
import viya_oracle
import viyadb_b200 as v
from viyadb_b200 import _native as N
from viyadb_b200 import db as vdb_mod
from viyadb_b200.query import FilterArgsPacker, QueryFactory

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# aggregate records, and the select / search records too: they share the predicate and the pruning (scan.cc:40-73)
RECS = [r for name in ("ref_gtest.jsonl", "ref_scenarios.jsonl", "ref_edge_scenarios.jsonl", "ref_fuzz_scenarios.jsonl",
                       "ref_gtest_select.jsonl", "ref_select_scenarios.jsonl", "ref_fuzz_select_scenarios.jsonl")
        for r in G.records(name) if "error" not in r and "seg" in r]
U64P = C.POINTER(C.c_uint64)


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    gxx = os.environ.get("VGPU_HARNESS_CXX") or shutil.which("g++")   # VGPU_HARNESS_CXX / _FLAGS: sanitizer builds
    extra = os.environ.get("VGPU_HARNESS_FLAGS", "").split()
    if gxx is None:
        pytest.skip("g++ not available")
    so = str(tmp_path_factory.mktemp("planner") / "libplanner_harness.so")
    subprocess.run([gxx, "-std=c++17", "-O2", "-shared", "-fPIC", *extra, "-Wno-unknown-pragmas", "-Wno-unused-function", "-o", so,
                    os.path.join(ROOT, "tests", "planner_harness.cc")], check=True)
    L = C.CDLL(so)
    L.h_planner_new.restype = C.c_void_p
    L.h_planner_error.restype = C.c_char_p
    L.h_planner_error.argtypes = [C.c_void_p]
    L.h_planner_free.argtypes = [C.c_void_p]
    for f in ("h_planner_nprog", "h_planner_conj", "h_planner_max_depth", "h_planner_nslots"):
        getattr(L, f).argtypes = [C.c_void_p]
        getattr(L, f).restype = C.c_uint32
    L.h_planner_tighten.restype = C.c_uint64
    L.h_to_ordered_host.restype = C.c_uint64
    L.h_to_ordered_host.argtypes = [C.c_uint64, C.c_uint32]
    return L


def schema_arrays(t):
    """kind / type / agg / literal type per schema column, as Table._create_device_table hands them to vgpu_table_create"""
    kinds, types, aggs, lits = [], [], [], []
    for c in t.dimensions + t.metrics:
        btype, lit = c.type, 0
        if c.kind == N.METRIC_BITSET and c.type in (N.U8, N.U16):
            btype, lit = N.U32, c.type + 1
        kinds.append(c.kind); types.append(btype); aggs.append(c.agg); lits.append(lit)
    if t.has_hidden_count:
        kinds.append(N.METRIC_HIDDEN_COUNT); types.append(N.U64); aggs.append(N.AGG_COUNT); lits.append(0)
    return [np.array(a, dtype=np.uint32) for a in (kinds, types, aggs, lits)]


def new_planner(lib, t, packer, key_cols=(), metric_cols=(), having=False, tune=0):
    kinds, types, aggs, lits = schema_arrays(t)
    nodes = (N.PredNode * max(1, len(packer.nodes)))()
    for i, (kind, op, col, arg, n) in enumerate(packer.nodes):
        nodes[i] = N.PredNode(kind, op, col, arg, n, 0)
    args = np.array(packer.args or [0], dtype=np.uint64)
    kc = np.array(list(key_cols) or [0], dtype=np.uint32)
    mc = np.array(list(metric_cols) or [0], dtype=np.uint32)
    p32 = C.POINTER(C.c_uint32)
    h = lib.h_planner_new(C.c_uint32(len(kinds)), kinds.ctypes.data_as(p32), types.ctypes.data_as(p32), aggs.ctypes.data_as(p32),
                          lits.ctypes.data_as(p32), C.c_uint32(len(packer.nodes)), nodes, C.c_uint32(len(packer.args)),
                          args.ctypes.data_as(U64P), C.c_uint32(len(key_cols)), kc.ctypes.data_as(p32),
                          C.c_uint32(len(metric_cols)), mc.ctypes.data_as(p32), C.c_int(1 if having else 0), C.c_uint32(tune))
    err = lib.h_planner_error(h)
    if err:
        lib.h_planner_free(h)
        raise RuntimeError(err.decode())
    return h


def widen(arr):
    """cells as the kernel widens them: zero- / sign-extended integers, raw IEEE bits"""
    a = np.asarray(arr)
    if a.dtype.kind == "f":
        return a.view("<u4").astype("<u8") if a.dtype.itemsize == 4 else a.view("<u8").copy()
    return a.astype("<i8").view("<u8") if a.dtype.kind == "i" else a.astype("<u8")


def widened_columns(t, seg, n, hidden):
    cols = []
    for c in t.dimensions + t.metrics:
        if c.kind == N.METRIC_BITSET:
            off = np.asarray(seg[c.name][0]).astype("<u8")
            cols.append(np.ascontiguousarray(np.diff(off[:n + 1])))    # the predicate sees the cardinality (Q7)
        else:
            cols.append(np.ascontiguousarray(widen(seg[c.name][:n])))
    if t.has_hidden_count:
        cols.append(np.ascontiguousarray(np.asarray(hidden[:n] if hidden is not None else np.zeros(n)).astype("<u8")))
    return cols


def ptr_array(cols):
    arr = (U64P * len(cols))()
    for i, c in enumerate(cols):
        arr[i] = c.ctypes.data_as(U64P)
    return arr


def open_host_table(rec):
    db = v.Database({"tables": [rec["table"]]}, device=None)
    t = db.get_table(rec["table"]["name"])
    hdr, _ = vdb_mod.read_dump(G.seg_path(rec["seg"]))
    for d in t.dimensions:
        if d.kind == N.DIM_STRING:
            c2v = hdr["dicts"][d.name]
            d.dict.c2v = list(c2v)
            d.dict.v2c = {s: i for i, s in enumerate(c2v)}
    return db, t


def test_fixtures_present():
    assert len(RECS) >= 400


@pytest.mark.parametrize("rec", RECS, ids=[G.rec_id(r) for r in RECS])
def test_planner_program_matches_oracle(lib, rec):
    db, t = open_host_table(rec)
    q = rec["query"]
    query = QueryFactory.create(q, db)
    packer = FilterArgsPacker(t).visit(query.filter)
    h = new_planner(lib, t, packer)
    h_interp = new_planner(lib, t, packer, tune=4096)   # VGPU_TUNE bit 12: never the unrolled conjunction
    try:
        assert lib.h_planner_max_depth(h) <= 5
        assert lib.h_planner_conj(h_interp) == 0
        hdr, segs, dicts, hidden = viya_oracle.read_dump(G.seg_path(rec["seg"]))
        dims, mets = viya_oracle.parse_schema(rec["table"])
        ocols = {c.name: c for c in dims + mets}
        flt = viya_oracle.make_filter(q.get("filter"))
        ncols = len(t.dimensions) + len(t.metrics) + (1 if t.has_hidden_count else 0)
        scanned_segments = 0
        key_lo, key_hi, passing_keys = {}, {}, {}
        for si, seg in enumerate(segs):
            n = viya_oracle._seg_rows(seg, dims + mets)
            cols = widened_columns(t, seg, n, hidden[si] if hidden else None)
            # ---- pruning on the statistics the device reduces at put time (dimensions; ordered domain) ----
            omin = np.full(ncols, 2**64 - 1, dtype=np.uint64)
            omax = np.zeros(ncols, dtype=np.uint64)
            if n:
                for ci, c in enumerate(t.dimensions):
                    o = np.array([lib.h_to_ordered_host(int(x), c.type) for x in np.unique(cols[ci]).tolist()], dtype=np.uint64)
                    omin[ci], omax[ci] = o.min(), o.max()
            keep = lib.h_planner_process_segment(C.c_void_p(h), C.c_uint64(n), omin.ctypes.data_as(U64P), omax.ctypes.data_as(U64P))
            assert bool(keep) == bool(viya_oracle.process_segment(flt, seg, n, ocols, dicts)), ("prune", si)
            if not keep:
                continue
            scanned_segments += 1
            if n == 0:
                continue
            # ---- the row predicate: both code paths against the oracle, row by row ----
            want = np.asarray(viya_oracle.eval_filter(flt, seg, n, ocols, dicts)).astype(bool)
            pa = ptr_array(cols)
            for handle, force in ((h, 0), (h, 1), (h_interp, 0)):
                got = np.zeros(n, dtype=np.uint8)
                lib.h_planner_eval_rows(C.c_void_p(handle), C.c_uint64(n), pa, C.c_int(force), got.ctypes.data_as(C.POINTER(C.c_uint8)))
                bad = np.nonzero(got.astype(bool) != want)[0]
                assert len(bad) == 0, ("row predicate", si, int(bad[0]), force)
            for dc in (query.dimension_cols if q["type"] == "aggregate" else []):
                ci = t.schema_index(dc.dim)
                key_lo[ci] = min(key_lo.get(ci, 2**64 - 1), int(cols[ci].min()))
                key_hi[ci] = max(key_hi.get(ci, 0), int(cols[ci].max()))
                passing_keys.setdefault(ci, []).append(cols[ci][want])
        assert scanned_segments == rec["stats"]["scanned_segments"]      # the REFERENCE's own count

        # ---- key-domain tightening (dense group tables): no passing row outside the domain ----
        for ci, parts in passing_keys.items():
            lo, hi, applies = C.c_uint64(key_lo[ci]), C.c_uint64(key_hi[ci]), C.c_int(0)
            lut = lib.h_planner_tighten(C.c_void_p(h), C.c_uint32(ci), C.byref(lo), C.byref(hi), C.byref(applies))
            vals = np.concatenate(parts) if parts else np.zeros(0, np.uint64)
            if not applies.value or not len(vals):
                continue
            assert lo.value <= hi.value
            assert int(vals.min()) >= lo.value and int(vals.max()) <= hi.value, ("domain", ci, lo.value, hi.value)
            if lut:
                assert all((lut >> int(x)) & 1 for x in np.unique(vals).tolist()), ("lookup mask", ci, hex(lut))
    finally:
        lib.h_planner_free(C.c_void_p(h))
        lib.h_planner_free(C.c_void_p(h_interp))


HAVING = [r for r in RECS if "having" in r["query"]]


@pytest.mark.parametrize("rec", HAVING, ids=[G.rec_id(r) for r in HAVING])
def test_device_having_program_matches_oracle(lib, rec):
    db, t = open_host_table(rec)
    q = rec["query"]
    query = QueryFactory.create(q, db)
    hp = FilterArgsPacker(t).visit(query.having)
    key_cols = [t.schema_index(dc.dim) for dc in query.dimension_cols]
    metric_cols = [t.schema_index(mc.metric) for mc in query.metric_cols]
    h = new_planner(lib, t, hp, key_cols, metric_cols, having=True)
    try:
        hdr, segs, dicts, hidden = viya_oracle.read_dump(G.seg_path(rec["seg"]))
        res = viya_oracle.run_query(rec["table"], segs, dicts, {k: val for k, val in q.items() if k not in ("sort", "skip", "limit")},
                                    now=rec.get("rollup_ts"), hidden_counts=hidden)
        g = res["groups"]
        ngroups = res["stats"]["aggregated_recs"]
        if ngroups == 0:
            return
        # sources: selected keys, then selected metrics (raw accumulators: AVG compares the sum, BITSET the cardinality)
        sources = [np.ascontiguousarray(widen(k)) for k in g["keys"]] + [np.ascontiguousarray(widen(a)) for a in g["accs"]]
        got = np.zeros(ngroups, dtype=np.uint8)
        lib.h_planner_eval_groups(C.c_void_p(h), C.c_uint64(ngroups), ptr_array(sources), got.ctypes.data_as(C.POINTER(C.c_uint8)))
        # the oracle sent exactly the groups its HAVING kept (no sort / window in this run): count them
        header = 1 if q.get("header") else 0
        assert int(got.sum()) == len(res["rows"]) - header
    finally:
        lib.h_planner_free(C.c_void_p(h))
