"""The drop-in boundary, end to end in C++: one reference process (unmodified ViyaDB core, built by
oracle/Makefile) runs every scenario query through its stock query::QueryRunner (g++-JIT) and through
vgpu_host::GpuQueryRunner (viyadb_b200/host/gpu_query_runner.h -> C ABI -> CUDA). Same Database, same
segments, same query objects: the row sets and QueryStats must agree, and both must equal the golden
rows recorded from the reference. Needs the pre-warmed JIT cache (tests/dropin_util.py, run by
__graft_entry__.build() where /root/reference exists)."""
import os

import pytest

import dropin_util as D
import golden_util as G
from helpers import rows_equal

pytestmark = pytest.mark.gpu
import scenarios  # noqa: E402  (tests/golden on sys.path via dropin_util)

GOLD = {r["test"]: r for r in G.records("ref_scenarios.jsonl")}


@pytest.mark.parametrize("sc", scenarios.SCENARIOS, ids=[s["name"] for s in scenarios.SCENARIOS])
def test_gpu_runner_equals_stock_runner(sc):
    if not os.path.exists(D.CLI):
        pytest.skip("vgpu_cli not built (needs the reference headers: make -C oracle gpu_cli)")
    out = D.run_scenario(sc)
    assert "fatal" not in out, out.get("fatal")
    float_table = any(m["type"].startswith(("float_", "double_")) and m["type"].endswith(("_sum", "_avg"))
                      for m in sc["table"]["metrics"])
    for qi, (q, res) in enumerate(zip(sc["queries"], out["results"])):
        stock, gpu = res["stock"], res["gpu"]
        assert ("error" in stock) == ("error" in gpu), (qi, stock.get("error"), gpu.get("error"))
        if "error" in stock:
            continue
        gold = GOLD[f"{sc['name']}.{qi}"]
        ordered = bool(q.get("sort"))
        if (q.get("limit") or q.get("skip")) and not ordered:
            assert len(gpu["rows"]) == len(stock["rows"]) == len(gold["rows"])
        elif float_table:
            ncols = len(gpu["rows"][0]) if gpu["rows"] else 0
            fcols = tuple(range(ncols))
            assert rows_equal_numeric(gpu["rows"], stock["rows"]), (qi, gpu["rows"][:3], stock["rows"][:3])
        else:
            if ordered:
                assert gpu["rows"] == stock["rows"] or sorted(gpu["rows"]) == sorted(stock["rows"]), qi
            else:
                assert sorted(gpu["rows"]) == sorted(stock["rows"]), (qi, sorted(gpu["rows"])[:3], sorted(stock["rows"])[:3])
            assert sorted(stock["rows"]) == sorted(gold["rows"]), qi
        for k in ("scanned_segments", "scanned_recs", "aggregated_recs", "output_recs"):
            assert gpu["stats"][k] == stock["stats"][k] == gold["stats"][k], (qi, k, gpu["stats"], stock["stats"])


def rows_equal_numeric(a, b, rtol=1e-12):
    """Rows equal as sets where numeric-looking cells may differ by rtol (double sums: order of adds)."""
    if len(a) != len(b):
        return False

    def norm(r):
        out = []
        for x in r:
            try:
                out.append(float(x))
            except ValueError:
                out.append(x)
        return out
    key = lambda r: tuple(x if isinstance(x, str) else round(x, 6) for x in r)
    a, b = sorted((norm(r) for r in a), key=key), sorted((norm(r) for r in b), key=key)
    for ra, rb in zip(a, b):
        for x, y in zip(ra, rb):
            if isinstance(x, str) or isinstance(y, str):
                if x != y:
                    return False
            elif not (x == y or abs(x - y) <= rtol * max(abs(x), abs(y))):
                return False
    return True


SEL_GOLD = {r["test"]: r for r in G.records("ref_select_scenarios.jsonl")}


@pytest.mark.parametrize("sc", scenarios.SELECT_SCENARIOS, ids=[s["name"] for s in scenarios.SELECT_SCENARIOS])
def test_gpu_runner_select_search_equals_stock_runner(sc):
    """Select / search: output order is defined, so the rows of both runners and of the golden run are identical."""
    if not os.path.exists(D.CLI):
        pytest.skip("vgpu_cli not built (needs the reference headers: make -C oracle gpu_cli)")
    out = D.run_scenario(sc)
    assert "fatal" not in out, out.get("fatal")
    for qi, (q, res) in enumerate(zip(sc["queries"], out["results"])):
        stock, gpu = res["stock"], res["gpu"]
        assert ("error" in stock) == ("error" in gpu), (qi, stock.get("error"), gpu.get("error"))
        if "error" in stock:
            continue
        gold = SEL_GOLD[f"{sc['name']}.{qi}"]
        assert gpu["rows"] == stock["rows"] == gold["rows"], (qi, gpu["rows"][:3], stock["rows"][:3])
        for k in ("scanned_segments", "scanned_recs", "aggregated_recs", "output_recs"):
            assert gpu["stats"][k] == stock["stats"][k] == gold["stats"][k], (qi, k, gpu["stats"], stock["stats"])


@pytest.mark.parametrize("mode", ["epoch", "mark"])
def test_resident_copy_follows_in_place_upserts(mode):
    """A second ingest batch into the SAME dimension tuples updates metric cells of existing rows in place
    (src/codegen/db/upsert.cc:386-393): no segment grows, yet the resident HBM copy is stale. The GPU runner must
    return what the stock runner returns, before and after."""
    if not os.path.exists(D.CLI):
        pytest.skip("vgpu_cli not built (needs the reference headers: make -C oracle gpu_cli)")
    sc = next(s for s in scenarios.SCENARIOS if s["name"] == "inapp")
    # same tuples again with other metric values, plus one new tuple
    reload_rows = [r[:3] + [str(float(r[3]) * 3 + 1)] for r in sc["rows"]] + [["IL", "gift", "20141114", "7.5"]]
    out = D.run_scenario(dict(sc, reload_rows=reload_rows, reload_mode=mode))
    assert "fatal" not in out, out.get("fatal")
    # "mark": exact dirty ranges (what an upsert hook reports) -> only those rows and the appended ones are uploaded
    # again (vgpu_segment_update); "epoch": the coarse notification -> whole segments
    assert (out["partial_updates"] > 0) == (mode == "mark"), out["partial_updates"]
    changed = 0
    for key in ("results", "results_after_reload"):
        for qi, res in enumerate(out[key]):
            stock, gpu = res["stock"], res["gpu"]
            assert ("error" in stock) == ("error" in gpu), (key, qi)
            if "error" in stock:
                continue
            q = sc["queries"][qi]
            if (q.get("limit") or q.get("skip")) and not q.get("sort"):
                assert len(gpu["rows"]) == len(stock["rows"]), (key, qi)
            else:
                assert rows_equal_numeric(gpu["rows"], stock["rows"]), (key, qi, gpu["rows"][:3], stock["rows"][:3])
            for k in ("scanned_segments", "scanned_recs", "aggregated_recs", "output_recs"):
                assert gpu["stats"][k] == stock["stats"][k], (key, qi, k)
    for a, b in zip(out["results"], out["results_after_reload"]):
        if "error" not in a["stock"] and sorted(a["stock"]["rows"]) != sorted(b["stock"]["rows"]):
            changed += 1
    assert changed > 0, "the second batch must change some results, or the test proves nothing"
