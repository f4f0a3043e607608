"""CUDA path vs the REAL reference: every aggregate query of the reference's own gtest suite (and
the oracle_cli scenarios), on the very segment bytes the reference scanned (golden dumps), through
the C ABI. Integer results, keys and QueryStats bit-exact; double sums within 1e-12 relative."""
import pytest

import golden_util as G

pytestmark = pytest.mark.gpu

GTEST = G.records("ref_gtest.jsonl")
SCEN = G.records("ref_scenarios.jsonl")
EDGE = G.records("ref_edge_scenarios.jsonl")


@pytest.fixture(scope="module")
def vdb(built_lib):
    import viyadb_b200
    return viyadb_b200


def run(vdb, rec, flags=0):
    from helpers import rows_equal
    db = vdb.Database({"tables": [rec["table"]]}, device=0)
    try:
        t = db.get_table(rec["table"]["name"])
        t.load_dump(G.seg_path(rec["seg"]))
        out = vdb.MemoryRowOutput()
        if "error" in rec:
            with pytest.raises((ValueError, OverflowError, KeyError, vdb.VgpuError)):
                db.query(rec["query"], out, now=rec.get("rollup_ts"), flags=flags)
            return
        stats = db.query(rec["query"], out, now=rec.get("rollup_ts"), flags=flags)
        q = rec["query"]
        ordered = bool(q.get("sort"))
        if (q.get("limit") or q.get("skip")) and not ordered:
            assert len(out.rows) == len(rec["rows"])     # only the count is defined
        else:
            float_cols = ()
            if G.is_float_metric_query(rec):
                names = q.get("dimensions", []) + q.get("metrics", []) if "select" not in q else [s["column"] for s in q["select"]]
                mt = {m["name"]: m["type"] for m in rec["table"]["metrics"]}
                float_cols = tuple(i for i, n in enumerate(names)
                                   if mt.get(n, "").startswith(("float_", "double_")) and mt[n].endswith(("_sum", "_avg")))
            tol = 1e-5 if any(m["type"].startswith("float_") for m in rec["table"]["metrics"]) else 1e-12
            if ordered and not float_cols:
                assert out.rows == rec["rows"] or sorted(out.rows) == sorted(rec["rows"])
            else:
                assert rows_equal(out.rows, rec["rows"], float_cols, tol), (out.rows[:5], rec["rows"][:5])
        for k, v in rec["stats"].items():
            assert getattr(stats, k) == v, (k, getattr(stats, k), v)
    finally:
        db.close()


@pytest.mark.parametrize("rec", GTEST, ids=[G.rec_id(r) for r in GTEST])
def test_reference_gtests(vdb, rec):
    run(vdb, rec)


@pytest.mark.parametrize("rec", GTEST, ids=[G.rec_id(r) for r in GTEST])
def test_reference_gtests_hashed_group_table(vdb, rec):
    """Same vectors with the open-addressing group table forced (VGPU_PLAN_FORCE_HASH)."""
    if "error" in rec:
        pytest.skip("error case")
    try:
        run(vdb, rec, flags=1)
    except vdb.VgpuError as e:
        if e.code == -2:
            pytest.skip("key wider than 64 bits")
        raise


@pytest.mark.parametrize("rec", SCEN, ids=[G.rec_id(r) for r in SCEN])
def test_reference_scenarios(vdb, rec):
    run(vdb, rec)


@pytest.mark.parametrize("flags", [0, 1], ids=["auto", "force_hash"])
@pytest.mark.parametrize("rec", EDGE, ids=[G.rec_id(r) for r in EDGE])
def test_reference_edge_cases(vdb, rec, flags):
    """Empty table, ragged segments, 64-bit extremes, int wrap, bitset duplicates across segments, time literals."""
    try:
        run(vdb, rec, flags=flags)
    except vdb.VgpuError as e:
        if flags == 1 and e.code == -2:
            pytest.skip("key wider than 64 bits")
        raise

