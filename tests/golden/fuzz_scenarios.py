"""Seeded random scenarios for the differential pin of the oracle against the REAL reference (tests/golden/make_golden.py
fuzz -> ref_fuzz_scenarios.jsonl; tests/test_oracle_golden.py). TEST INFRASTRUCTURE.

Every scenario is a random table (string / numeric / time / microtime / boolean dimensions of random widths, value metrics
of all ten numeric types x sum / min / max / avg, count, bitset), random rows over small domains (so that upserts merge
and several ragged segments exist) and random aggregate queries: nested and / or / not / in filters over dimensions and
metrics, HAVING, sort on totally ordered keys with skip / limit, time granularities over rollup rules.

Kept inside what the reference defines: no filter on byte / short columns (Q13: does not compile), no week granularity
(Q10), sums of floating-point metrics over values that add exactly (multiples of 1/4), limit / skip only under a sort
that orders every selected dimension (Q11: otherwise only the row count is defined), no skip / limit next to HAVING
(the reference reads past the end of its sorted vector when HAVING leaves fewer than skip + limit rows)."""
import random

NOW = 1496570140
INT_TYPES = {"byte": (-128, 127), "ubyte": (0, 255), "short": (-32768, 32767), "ushort": (0, 65535),
             "int": (-2**31, 2**31 - 1), "uint": (0, 2**32 - 1), "long": (-2**63, 2**63 - 1), "ulong": (0, 2**64 - 1)}
FILTERABLE = ("ubyte", "ushort", "int", "uint", "long", "ulong", "float", "double")
GRANS = ["year", "month", "day", "hour", "minute", "second"]


def _num_domain(rng, typ, small):
    """a handful of values of the type: clustered small ones, sometimes the extremes"""
    if typ in ("float", "double"):
        vals = [rng.randrange(-40, 41) / 4.0 for _ in range(small)]
        if rng.random() < 0.4:
            vals += [0.0, -0.0]
        return [repr(v) for v in vals]
    lo, hi = INT_TYPES[typ]
    vals = [rng.randrange(max(lo, -60), min(hi, 60) + 1) for _ in range(small)]
    if rng.random() < 0.35:
        vals += [lo, hi]
    if rng.random() < 0.35:
        vals.append(rng.randrange(lo, hi + 1))
    return [str(v) for v in vals]


def _make_table(rng, idx, rich=False):
    dims, mets = [], []
    gen = {}      # column name -> list of input strings to draw from
    kinds = {}    # column name -> ("str"|"num"|"time"|"micro"|"bool"|"metric"|"bitset"|"count", numeric type)
    ndims = rng.randrange(2, 6)
    for d in range(ndims):
        name = f"d{d}"
        k = rng.choices(["str", "num", "time", "micro", "bool"], [5, 5, 2, 1, 1])[0]
        if k == "time" and any(v[0] in ("time", "micro") for v in kinds.values()):
            k = "num"
        if k == "str":
            conf = {"name": name}
            card = rng.choice([None, None, 200, 40000])
            if card:
                conf["cardinality"] = card
            gen[name] = [f"{chr(97 + d)}{j}" for j in range(rng.randrange(2, 9))]
            kinds[name] = ("str", None)
        elif k == "num":
            typ = rng.choice(list(INT_TYPES) + ["float", "double"])
            conf = {"name": name, "type": typ}
            gen[name] = _num_domain(rng, typ, rng.randrange(2, 7))
            kinds[name] = ("num", typ)
        elif k == "time":
            conf = {"name": name, "type": "time", "format": "posix"}
            if rng.random() < 0.5:
                rules = rng.sample([("hour", "1 days"), ("day", "1 weeks"), ("month", "1 years"), ("day", "3 months"),
                                    ("minute", "6 hours"), ("year", "2 years")], rng.randrange(1, 4))
                conf["rollup_rules"] = [{"granularity": g, "after": a} for g, a in rules]
            span = rng.choice([3 * 86400, 40 * 86400, 800 * 86400])
            gen[name] = [str(NOW - rng.randrange(0, span)) for _ in range(rng.randrange(4, 14))]
            kinds[name] = ("time", "uint")
        elif k == "micro":
            conf = {"name": name, "type": "microtime", "format": "micros"}
            span = rng.choice([3 * 86400, 500 * 86400])
            gen[name] = [str((NOW - rng.randrange(0, span)) * 1000000 + rng.randrange(0, 1000000)) for _ in range(rng.randrange(3, 9))]
            kinds[name] = ("micro", "ulong")
        else:
            conf = {"name": name, "type": "boolean"}
            gen[name] = ["true", "false"]
            kinds[name] = ("bool", None)
        dims.append(conf)
    nmets = rng.randrange(1, 6)
    have_count = False
    for m in range(nmets):
        k = rng.choices(["value", "count", "bitset"], [7, 2, 2])[0]
        if k == "count" and have_count:
            k = "value"
        if k == "count":
            conf = {"name": "count", "type": "count"}
            if rich and rng.random() < 0.4:
                conf["max"] = 10**11                      # -> a 64-bit count column (column.cc:278-284)
            mets.append(conf)
            kinds["count"] = ("count", "uint")
            have_count = True
        elif k == "bitset":
            name = f"u{m}"
            conf = {"name": name, "type": "bitset"}
            dom = rng.choice([5, 40, 4000000000])
            if rich:
                # the id type follows "max" (column.cc:54-62, 398-400): ubyte / ushort / uint ids in a Roaring, ulong ids
                # in a Roaring64Map (bitset.h:27-31); a filter literal on the metric has that type (Q7)
                mx = rng.choice([200, 50000, None, 10**12])
                if mx:
                    conf["max"] = mx
                dom = {200: rng.choice([5, 200]), 50000: rng.choice([40, 50000]), None: dom, 10**12: rng.choice([40, 10**12])}[mx]
            gen[name] = [str(rng.randrange(0, dom)) for _ in range(rng.randrange(3, 12))]
            mets.append(conf)
            kinds[name] = ("bitset", "uint")
        else:
            typ = rng.choice(list(INT_TYPES) + ["float", "double"])
            agg = rng.choice(["sum", "min", "max", "avg"])
            name = f"m{m}"
            mets.append({"name": name, "type": f"{typ}_{agg}"})
            gen[name] = _num_domain(rng, typ, rng.randrange(3, 8))
            kinds[name] = ("metric", typ, agg)
    table = {"name": "events", "segment_size": rng.choice([3, 7, 16, 50]), "dimensions": dims, "metrics": mets}
    return table, gen, kinds


def _literal(rng, name, kinds, gen):
    k = kinds[name]
    if k[0] == "str":
        return rng.choice(gen[name] + ["nope"])
    if k[0] == "bool":
        return rng.choice(["true", "false"])
    if k[0] == "count":
        return str(rng.randrange(0, 5))
    if k[0] == "bitset":
        return str(rng.randrange(0, 4))     # compared with the cardinality (Q7)
    if k[0] in ("time", "micro"):
        v = int(rng.choice(gen[name]))
        return str(v + rng.choice([0, 0, 1, -1, 3600 * (1000000 if k[0] == "micro" else 1)]))
    typ = k[1]
    v = rng.choice(gen[name])
    if typ in ("float", "double"):
        return repr(float(v) + rng.choice([0.0, 0.0, 0.25, -0.5]))
    lo, hi = INT_TYPES[typ]
    return str(min(hi, max(lo, int(v) + rng.choice([0, 0, 1, -1, 7]))))


def _filterable(name, kinds):
    k = kinds[name]
    if k[0] in ("num", "metric"):
        return k[1] in FILTERABLE
    return True


def _filter(rng, cols, kinds, gen, depth=0):
    r = rng.random()
    if depth < 2 and r < 0.35:
        return {"op": rng.choice(["and", "or"]), "filters": [_filter(rng, cols, kinds, gen, depth + 1) for _ in range(rng.randrange(2, 4))]}
    if depth < 2 and r < 0.45:
        return {"op": "not", "filter": _filter(rng, cols, kinds, gen, depth + 1)}
    name = rng.choice(cols)
    k = kinds[name]
    if r < 0.65 and k[0] not in ("bool",):
        return {"op": "in", "column": name, "values": sorted({_literal(rng, name, kinds, gen) for _ in range(rng.randrange(1, 5))})}
    ops = ["eq", "ne"] if k[0] in ("str", "bool") else ["eq", "ne", "lt", "le", "gt", "ge"]
    return {"op": rng.choice(ops), "column": name, "value": _literal(rng, name, kinds, gen)}


def _make_query(rng, table, gen, kinds):
    dims = [d["name"] for d in table["dimensions"]]
    mets = [m["name"] for m in table["metrics"]]
    sel_d = rng.sample(dims, rng.randrange(0, min(3, len(dims)) + 1))
    sel_m = rng.sample(mets, rng.randrange(1, min(3, len(mets)) + 1))
    select = []
    for d in sel_d:
        c = {"column": d}
        if kinds[d][0] in ("time", "micro"):
            if rng.random() < 0.6:
                c["granularity"] = rng.choice(GRANS)
            c["format"] = "%Y-%m-%d %H:%M:%S"
        select.append(c)
    select += [{"column": m} for m in sel_m]
    q = {"type": "aggregate", "table": "events", "select": select}
    fcols = [c for c in dims + mets if _filterable(c, kinds)]
    if fcols and rng.random() < 0.75:
        q["filter"] = _filter(rng, fcols, kinds, gen)
    hcols = [c for c in sel_d + sel_m if _filterable(c, kinds) and kinds[c][0] not in ("time", "micro")]
    if hcols and rng.random() < 0.3:
        q["having"] = _filter(rng, hcols, kinds, gen, depth=1)
    if rng.random() < 0.45:
        # a total order: the chosen keys first, then every selected dimension (ties would make the cut undefined)
        first = rng.sample(sel_d + sel_m, rng.randrange(1, min(2, len(sel_d + sel_m)) + 1))
        keys = first + [d for d in sel_d if d not in first]
        q["sort"] = [{"column": c, "ascending": rng.random() < 0.5} for c in keys]
        # no window next to HAVING: the reference walks post_agg.begin() + skip + limit whatever HAVING left in the vector
        # (sort.cc:66-73) — undefined behaviour (a segfault in practice) when fewer rows survive
        if "having" not in q:
            if rng.random() < 0.6:
                q["limit"] = rng.randrange(1, 8)
            if rng.random() < 0.3:
                q["skip"] = rng.randrange(0, 4)
    if rng.random() < 0.15:
        q["header"] = True
    return q


def make_scenarios(seed=20261017, count=36, queries=6, prefix="fuzz", rows=(20, 260), rich=False):
    rng = random.Random(seed)
    out = []
    for i in range(count):
        table, gen, kinds = _make_table(rng, i, rich)
        nrows = rng.choice([0, 9, 40, 120, 300]) if i % 9 == 0 else rng.randrange(*rows)
        cols = [d["name"] for d in table["dimensions"]] + [m["name"] for m in table["metrics"] if m["type"] != "count"]
        data = [[rng.choice(gen[c]) for c in cols] for _ in range(nrows)]
        sc = {"name": f"{prefix}{i:02d}", "table": table, "rows": data, "rollup_ts": NOW,
              "queries": [_make_query(rng, table, gen, kinds) for _ in range(queries)]}
        out.append(sc)
    return out


# two batches: small tables with deep filters, then larger ones (more merged upserts, multi-id bitset cells, wrapped sums)
# (a third batch, "fuzzc": bitset metrics of every id width — ubyte / ushort / uint / ulong — and 64-bit count columns)
FUZZ_SCENARIOS = (make_scenarios() + make_scenarios(seed=77001, count=24, prefix="fuzzb", rows=(300, 1200))
                  + make_scenarios(seed=5150, count=24, prefix="fuzzc", rows=(40, 500), rich=True))


# ---- select / search: rows in (segment, tuple) order with the per-segment limit quirk; distinct values ----
def _make_select(rng, table, gen, kinds):
    dims = [d["name"] for d in table["dimensions"]]
    mets = [m["name"] for m in table["metrics"]]
    fcols = [c for c in dims + mets if _filterable(c, kinds)]
    if rng.random() < 0.3:
        d = rng.choice([x for x in dims if kinds[x][0] in ("str", "num", "bool") and (kinds[x][1] or "") not in ("float", "double")] or dims[:1])
        pool = gen.get(d, ["a"])
        term = rng.choice(["", rng.choice(pool)[:rng.randrange(1, 3)], rng.choice(pool)[-1:], "zz"])
        q = {"type": "search", "table": "events", "dimension": d, "term": term}
        if rng.random() < 0.5:
            q["limit"] = rng.randrange(1, 5)
    else:
        q = {"type": "select", "table": "events"}
        if rng.random() < 0.2:
            q["select"] = [{"column": "*"}]
        else:
            sel_d = rng.sample(dims, rng.randrange(1, min(3, len(dims)) + 1))
            sel_m = rng.sample(mets, rng.randrange(0, min(3, len(mets)) + 1))
            select = []
            for d in sel_d:
                c = {"column": d}
                if kinds[d][0] in ("time", "micro") and rng.random() < 0.7:
                    c["format"] = "%Y-%m-%d %H:%M:%S"
                select.append(c)
            q["select"] = select + [{"column": m} for m in sel_m]
        if rng.random() < 0.5:
            q["limit"] = rng.randrange(1, 12)
        if rng.random() < 0.35:
            q["skip"] = rng.randrange(0, 9)
    if fcols and rng.random() < 0.6:
        q["filter"] = _filter(rng, fcols, kinds, gen, depth=1)
    if rng.random() < 0.15:
        q["header"] = True
    return q


def make_select_scenarios(seed=424242, count=20, queries=6):
    rng = random.Random(seed)
    out = []
    for i in range(count):
        table, gen, kinds = _make_table(rng, i)
        nrows = rng.randrange(10, 160)
        cols = [d["name"] for d in table["dimensions"]] + [m["name"] for m in table["metrics"] if m["type"] != "count"]
        data = [[rng.choice(gen[c]) for c in cols] for _ in range(nrows)]
        out.append({"name": f"fuzzs{i:02d}", "table": table, "rows": data, "rollup_ts": NOW,
                    "queries": [_make_select(rng, table, gen, kinds) for _ in range(queries)]})
    return out


FUZZ_SELECT_SCENARIOS = make_select_scenarios()
