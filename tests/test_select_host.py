"""Host side of the select / search queries (CPU): the sequential replay of the reference's search loop over the
per-segment (first row, code) lists the device produces must reproduce the real reference's outputs — here the
lists are computed with numpy from the golden segment dumps (a test double of vgpu_query_search)."""
import numpy as np
import pytest

import golden_util as G
import viya_oracle
from viyadb_b200.query import replay_search

RECS = [r for r in G.records("ref_select_scenarios.jsonl") + G.records("ref_gtest_select.jsonl")
        if r["query"]["type"] == "search" and "error" not in r]


@pytest.mark.parametrize("rec", RECS, ids=[G.rec_id(r) for r in RECS])
def test_search_replay_matches_reference(rec):
    hdr, segs, dicts, hidden = viya_oracle.read_dump(G.seg_path(rec["seg"]))
    q = rec["query"]
    dims, mets = viya_oracle.parse_schema(rec["table"])
    cols = {c.name: c for c in dims + mets}
    d = cols[q["dimension"]]
    flt = viya_oracle.make_filter(q.get("filter"))
    lists = []
    for seg in segs:
        n = viya_oracle._seg_rows(seg, dims + mets)
        if not viya_oracle.process_segment(flt, seg, n, cols, dicts):
            continue
        first = {}
        if n:
            for i in np.nonzero(viya_oracle.eval_filter(flt, seg, n, cols, dicts))[0].tolist():
                first.setdefault(int(seg[d.name][i]), i)
        lists.append([c for c, _ in sorted(first.items(), key=lambda kv: kv[1])])

    def fmt_value(code):
        if d.kind == "string":
            return dicts[d.name][code]
        if d.kind == "boolean":
            return "true" if code else "false"
        return str(code)
    values, ncodes = replay_search(lists, fmt_value, q["term"], int(q.get("limit", 0)))
    want_rows = rec["rows"][1:] if q.get("header") else rec["rows"]
    assert [values] == want_rows
    assert ncodes == rec["stats"]["aggregated_recs"]
    assert len(values) == rec["stats"]["output_recs"]
