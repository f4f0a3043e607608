"""Shared helpers for the parity tests: seeded random tables in the reference's own column types,
fed both to the CUDA path (through the C ABI) and to the oracle."""
import numpy as np

import viya_oracle

NPT = viya_oracle.NP


def random_table(conf, nsegs, rows_per_seg, seed, spec, last_rows=None):
    """spec: column name -> (lo, hi) inclusive range | ("ids", domain, max_per_row) for bitsets |
    ("float", lo, hi). Returns (segments for the oracle, dicts, hidden_counts)."""
    dims, mets = viya_oracle.parse_schema(conf)
    rng = np.random.default_rng(seed)
    segs, hidden = [], []
    dicts = {}
    for d in dims:
        if d.kind == "string":
            lo, hi = spec[d.name]
            dicts[d.name] = ["__exceeded"] + [f"{d.name}_{i}" for i in range(1, hi + 1)]
    has_avg = any(m.agg == "avg" for m in mets)
    has_count = any(m.agg == "count" for m in mets)
    for s in range(nsegs):
        n = rows_per_seg if (last_rows is None or s < nsegs - 1) else last_rows
        seg = {}
        for c in dims + mets:
            sp = spec[c.name]
            dt = np.dtype(NPT[c.type])
            if sp[0] == "ids":
                _, domain, max_per_row = sp
                counts = rng.integers(1, max_per_row + 1, n) if max_per_row > 1 else np.ones(n, dtype=np.int64)
                offsets = np.zeros(n + 1, dtype="<u8")
                offsets[1:] = np.cumsum(counts)
                values = np.empty(int(offsets[-1]), dtype="<u8")
                for r in range(n):  # ids of one cell are distinct (a set)
                    k = int(counts[r])
                    values[int(offsets[r]):int(offsets[r + 1])] = (
                        rng.choice(domain, size=k, replace=False) if k > 1 else rng.integers(0, domain, 1))
                seg[c.name] = (offsets, values)
            elif sp[0] == "float":
                seg[c.name] = rng.uniform(sp[1], sp[2], n).astype(dt)
            elif sp[0] == "choice":
                seg[c.name] = rng.choice(np.array(sp[1], dtype=dt), n)
            else:
                lo, hi = sp
                seg[c.name] = rng.integers(lo, hi, n, endpoint=True, dtype=np.int64).astype(dt) \
                    if dt.kind != "u" or hi < 2**63 else rng.integers(lo, hi, n, endpoint=True, dtype=np.uint64).astype(dt)
        segs.append(seg)
        hidden.append(rng.integers(1, 4, n).astype("<u8") if (has_avg and not has_count) else None)
    return segs, dicts, hidden


def upload(table, segs, dicts, hidden):
    """Put oracle-format segments into the device table through the C ABI."""
    for name, c2v in dicts.items():
        d = table.dimension(name).dict
        d.c2v = list(c2v)
        d.v2c = {v: i for i, v in enumerate(c2v)}
    for i, seg in enumerate(segs):
        g = {}
        for k, v in seg.items():
            g[k] = v   # bitset cells: (offsets, ids); Table.put_segment narrows the ids to the column's width
        table.put_segment(i, g, hidden[i] if hidden else None)


def rows_equal(got, want, float_cols=(), rtol=1e-12):
    """Set equality of formatted rows; float sum columns compared with a tolerance."""
    if not float_cols:
        return sorted(got) == sorted(want)
    if len(got) != len(want):
        return False
    def key(r):
        return tuple(x for i, x in enumerate(r) if i not in float_cols)
    g, w = sorted(got, key=key), sorted(want, key=key)
    for a, b in zip(g, w):
        if key(a) != key(b):
            return False
        for i in float_cols:
            if a[i] == b[i]:
                continue
            x, y = float(a[i]), float(b[i])
            if not (x == y or abs(x - y) <= rtol * max(abs(x), abs(y))):
                return False
    return True
