"""CPU checks of bench.py's roofline bookkeeping: the per-row byte figures are the ones SURVEY.md §8(d) states
(B_alg = sum of filter widths + s * sum of the remaining key / metric widths), and the peak / traffic files load."""
import json
import os

import pytest

import bench

# workload -> (nominal selectivity, B_alg of SURVEY §8d)
SURVEY_B_ALG = {"c0": (0.01, 4.12), "c1": (1 / 16, 5.25), "c2": (0.025, 10.6), "c3": (0.25, 9.0), "c4": (1.0, 20.0)}


@pytest.mark.parametrize("wname", sorted(SURVEY_B_ALG))
def test_algorithmic_bytes_per_row(wname):
    s, want = SURVEY_B_ALG[wname]
    w = bench.WORKLOADS[wname]
    assert w["filter_bytes"] + s * w["payload_bytes"] == pytest.approx(want, rel=1e-12)


@pytest.mark.parametrize("wname", sorted(SURVEY_B_ALG))
def test_widths_follow_the_table_schema(wname):
    """filter_bytes / payload_bytes / full_bytes restate the widths of the columns the query references / the table holds."""
    w = bench.WORKLOADS[wname]
    width = {}
    for d in w["table"]["dimensions"]:
        t = d.get("type", "string")
        if t in ("time", "string"):
            width[d["name"]] = 4
        elif t == "microtime":
            width[d["name"]] = 8
        else:
            width[d["name"]] = {"byte": 1, "ubyte": 1, "short": 2, "ushort": 2, "int": 4, "uint": 4, "long": 8, "ulong": 8,
                                "float": 4, "double": 8}[d["field_type"] if "field_type" in d else d.get("num_type", t)]
    for m in w["table"]["metrics"]:
        t = m["type"]
        base = t.split("_")[0]
        width[m["name"]] = 4 if t in ("count", "bitset") else {"int": 4, "uint": 4, "long": 8, "ulong": 8, "float": 4, "double": 8,
                                                               "byte": 1, "ubyte": 1, "short": 2, "ushort": 2}[base]
    assert sum(width.values()) == w["full_bytes"]

    def filter_cols(f, out):
        if f is None:
            return out
        if "column" in f:
            out.add(f["column"])
        for sub in f.get("filters", []):
            filter_cols(sub, out)
        if "filter" in f:
            filter_cols(f["filter"], out)
        return out

    q = w["query"]
    fcols = filter_cols(q.get("filter"), set())
    dims = [d if isinstance(d, str) else d["column"] for d in q.get("dimensions", [])]
    sel = q.get("select")
    if sel is not None:
        dims = [c if isinstance(c, str) else c["column"] for c in sel]
    payload = [c for c in dict.fromkeys(dims + list(q.get("metrics", []))) if c not in fcols]
    assert sum(width[c] for c in fcols) == w["filter_bytes"]
    assert sum(width[c] for c in payload) == w["payload_bytes"]


def test_peak_and_traffic_files():
    peak, src = bench.load_peaks()
    if os.path.exists(os.path.join(bench.ROOT, "MEASURED_PEAKS.json")):
        assert src.startswith("measured") and peak == json.load(open(os.path.join(bench.ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    else:
        assert src.startswith("fallback")
    for wname in ("c2", "c3", "c4"):
        traffic, tsrc = bench.load_traffic(wname)
        assert traffic > bench.WORKLOADS[wname]["rows"] * bench.WORKLOADS[wname]["filter_bytes"]
        assert os.path.exists(os.path.join(bench.ROOT, tsrc.split(" ")[0]))
