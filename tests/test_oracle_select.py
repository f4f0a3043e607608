"""Select / search queries: the numpy oracle's run_select / run_search pinned to the REAL
reference — the select and search queries the reference's own gtest suite issues (test/select.cc, test/search.cc,
captured with their actual output) and extra oracle_cli scenarios (multi-segment skip / limit behaviour, all
column kinds). CPU only; groundwork for the GPU path of those queries."""
import pytest

import golden_util as G
import viya_oracle

GTEST = G.records("ref_gtest_select.jsonl")
SCEN = G.records("ref_select_scenarios.jsonl")
FUZZ = G.records("ref_fuzz_select_scenarios.jsonl")   # seeded random select / search queries (tests/golden/fuzz_scenarios.py)


def check(rec):
    hdr, segs, dicts, hidden = viya_oracle.read_dump(G.seg_path(rec["seg"]))
    q = rec["query"]
    if "error" in rec:
        with pytest.raises((ValueError, OverflowError, KeyError, RuntimeError)):
            if q["type"] == "select":
                viya_oracle.run_select(rec["table"], segs, dicts, q, hidden_counts=hidden)
            else:
                viya_oracle.run_search(rec["table"], segs, dicts, q)
        return
    if q["type"] == "select":
        got = viya_oracle.run_select(rec["table"], segs, dicts, q, hidden_counts=hidden)
        assert got["rows"] == rec["rows"]          # select output order is defined: (segment, tuple) order
    else:
        got = viya_oracle.run_search(rec["table"], segs, dicts, q)
        assert got["rows"] == rec["rows"]          # first-seen order of the scan
    for k, v in rec["stats"].items():
        assert got["stats"][k] == v, (k, got["stats"][k], v)


def test_golden_present():
    assert len(SCEN) >= 30, "tests/golden/ref_select_scenarios.jsonl missing: run tests/golden/make_golden.py select"


@pytest.mark.parametrize("rec", GTEST, ids=[G.rec_id(r) for r in GTEST])
def test_oracle_matches_reference_gtests(rec):
    check(rec)


@pytest.mark.parametrize("rec", SCEN, ids=[G.rec_id(r) for r in SCEN])
def test_oracle_matches_reference_scenarios(rec):
    check(rec)


@pytest.mark.parametrize("rec", FUZZ, ids=[G.rec_id(r) for r in FUZZ])
def test_oracle_matches_reference_fuzz(rec):
    check(rec)
