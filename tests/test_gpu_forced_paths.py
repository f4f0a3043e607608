"""The branches of the device path that only large or uneven inputs reach, forced at test sizes and checked against the
oracle: the shared-memory-set dedupe of count-distinct pairs (the path the timed C2 configuration takes), the general
L2-partitioned path with more than one partition, the fast path overflowing into the general one, pair-region overflow
+ regrow, hash-table x4 regrow, 64-bit bitset ids (util::Bitset<8> = Roaring64Map, src/util/bitset.h:27-31), float
keys with -0.0 in 64-bit-packed hash tables, the bucket dictionary of rolled-up time keys, and concurrent queries next
to a concurrent put on one context (src/db/database.cc:28-33). All through the C ABI; which path ran is asserted from
vgpu_result_view.distinct_paths / attempts, so a test cannot pass on the default branch by accident."""
import threading

import numpy as np
import pytest

import viya_oracle
from helpers import random_table, rows_equal, upload

pytestmark = pytest.mark.gpu
NOW = 1496570140

MID = {"name": "mid", "segment_size": 1000000,
       "dimensions": [{"name": "d0"}, {"name": "d1", "cardinality": 200}, {"name": "n2", "type": "ushort"}],
       "metrics": [{"name": "mn", "type": "int_min"}, {"name": "uid", "type": "bitset"}]}


def mid_table(seed=99, nsegs=5, last=234567, ids=1000000):
    rng = np.random.default_rng(seed)
    dicts = {"d0": ["__exceeded"] + [f"a{i}" for i in range(1, 51)], "d1": ["__exceeded"] + [f"b{i}" for i in range(1, 21)]}
    segs = []
    for s in range(nsegs):
        n = 1000000 if s < nsegs - 1 else last
        segs.append({"d0": rng.integers(1, 51, n).astype("<u4"), "d1": rng.integers(1, 21, n).astype("<u1"),
                     "n2": rng.integers(0, 1000, n).astype("<u2"),
                     "mn": rng.integers(-2**31, 2**31, n).astype("<i4"),
                     "uid": (np.arange(n + 1, dtype="<u8"), rng.integers(0, ids, n).astype("<u8"))})
    return segs, dicts


@pytest.fixture(scope="module")
def mid(built_lib):
    import viyadb_b200 as v
    segs, dicts = mid_table()
    db = v.Database({"tables": [MID]}, device=0)
    upload(db.get_table("mid"), segs, dicts, None)
    want = {}
    yield v, db, segs, dicts, want
    db.close()


Q_ALL = {"type": "aggregate", "table": "mid", "dimensions": ["d0", "d1"], "metrics": ["mn", "uid"]}          # 4.2e6 pairs > 2^21
Q_SEL = {"type": "aggregate", "table": "mid", "dimensions": ["d0", "d1", "n2"], "metrics": ["uid", "mn"],
         "filter": {"op": "and", "filters": [{"op": "in", "column": "d0", "values": ["a3", "a11", "a19", "a27", "a42"]},
                                             {"op": "ge", "column": "n2", "value": "250"}, {"op": "lt", "column": "n2", "value": "750"}]}}


def run(mid, q, hooks=None, flags=0):
    v, db, segs, dicts, want = mid
    key = repr(sorted(q.items(), key=lambda kv: kv[0]))
    if key not in want:
        want[key] = viya_oracle.run_query(MID, segs, dicts, q, now=NOW)
    for name, val in (hooks or {}).items():
        db.set_test_hook(name, val)
    try:
        out = v.MemoryRowOutput()
        stats = db.query(q, out, now=NOW, flags=flags)
    finally:
        for name in (hooks or {}):
            db.set_test_hook(name, 2 if name == "tune" else 0)
    w = want[key]
    assert sorted(out.rows) == sorted(w["rows"])
    for k in ("scanned_segments", "scanned_recs", "aggregated_recs", "output_recs"):
        assert getattr(stats, k) == w["stats"][k], k
    return v, stats


def test_fast_path_shared_memory_sets(mid):
    """4.2e6 pairs from 1e6-row segments: buckets + shared-memory sets, one scan, no fallback."""
    from viyadb_b200 import _native as N
    v, stats = run(mid, Q_ALL)   # the first query of a table sizes its pair regions for rows / 16: one regrow here
    assert stats.distinct_paths & N.DEDUPE_FAST, stats.distinct_paths
    for _ in range(3):           # from then on the table's high-water marks size everything: one scan (work units are
        v, stats = run(mid, Q_ALL)   # handed out dynamically, so on a table this small the fullest region varies a little)
        if stats.attempts == 1:
            break
    assert stats.distinct_paths == N.DEDUPE_FAST, stats.distinct_paths
    assert stats.attempts == 1
    v, stats = run(mid, Q_SEL)   # 2.1e5 pairs over 1.25e5 groups (the C2 shape): one global set
    assert stats.distinct_paths & (N.DEDUPE_FAST | N.DEDUPE_SMALL)


def test_fast_path_small_sets_many_buckets(mid):
    from viyadb_b200 import _native as N
    v, stats = run(mid, Q_ALL, {"set_slots": 1024})   # 8192 buckets of ~500 pairs
    assert stats.distinct_paths == N.DEDUPE_FAST


def test_general_path_partitioned(mid):
    """The general path with B = 16 L2-sized partitions (what C2 ran in round 1)."""
    from viyadb_b200 import _native as N
    v, stats = run(mid, Q_ALL, {"tune": 2 | (1 << 19), "bucket_pairs": 1 << 18})
    assert stats.distinct_paths == N.DEDUPE_GENERAL | N.DEDUPE_PARTITIONED, stats.distinct_paths


def test_fast_path_overflow_falls_back(mid):
    """Buckets sized for far too few pairs: bucket overflow flag -> the general path redoes the counts."""
    from viyadb_b200 import _native as N
    v, stats = run(mid, Q_ALL, {"expect_pairs": 50000})
    assert stats.distinct_paths & N.DEDUPE_REDONE and stats.distinct_paths & N.DEDUPE_GENERAL, stats.distinct_paths
    v, stats = run(mid, Q_ALL)            # and the table is not stuck on the slow path
    assert stats.distinct_paths == N.DEDUPE_FAST


def test_pair_region_overflow_regrows(mid):
    v, stats = run(mid, Q_ALL, {"pairs_cap": 100000})
    assert stats.attempts >= 2, stats.attempts


def test_small_path_single_global_set(mid):
    from viyadb_b200 import _native as N
    v, stats = run(mid, Q_ALL, {"small_pairs": 1 << 26})
    assert stats.distinct_paths == N.DEDUPE_SMALL


def test_hash_table_regrows_x4(mid):
    from viyadb_b200 import _native as N
    v, stats = run(mid, Q_SEL, {"hash_cap": 1024}, flags=N.PLAN_FORCE_HASH)   # 1.25e5 groups: 1024 -> ... -> 2^18+
    assert stats.table_mode == 1 and stats.attempts >= 4, (stats.table_mode, stats.attempts)
    v, stats = run(mid, Q_ALL, {"hash_cap": 64, "pairs_cap": 50000}, flags=N.PLAN_FORCE_HASH)   # both overflows at once
    assert stats.attempts >= 3


# ---- 64-bit bitset ids ----
WIDE = {"name": "wide", "segment_size": 40000,
        "dimensions": [{"name": "d0"}, {"name": "n1", "type": "ushort"}],
        "metrics": [{"name": "count", "type": "count"}, {"name": "big", "type": "bitset", "max": 2**40},
                    {"name": "uid", "type": "bitset"}]}


def test_64bit_bitset_ids(built_lib):
    import viyadb_b200 as v
    from viyadb_b200 import _native as N
    rng = np.random.default_rng(5)
    dicts = {"d0": ["__exceeded"] + [f"a{i}" for i in range(1, 9)]}
    pool = np.concatenate([rng.integers(0, 2**40, 3000).astype("<u8"), np.array([2**40 - 1, 2**32, 2**32 - 1, 0, 0xFFFFFFFFFFFFFFFF], "<u8")])
    segs = []
    for s in range(3):
        n = 40000 if s < 2 else 17001
        counts = rng.integers(0, 4, n)    # cells with 0..3 ids, drawn from a pool so that groups really share ids
        offsets = np.zeros(n + 1, "<u8")
        offsets[1:] = np.cumsum(counts)
        values = pool[rng.integers(0, len(pool), int(offsets[-1]))]
        for r in np.nonzero(counts > 1)[0][:5000]:   # ids of one cell are a set
            lo, hi = int(offsets[r]), int(offsets[r + 1])
            values[lo:hi] = pool[rng.choice(len(pool), hi - lo, replace=False)]
        segs.append({"d0": rng.integers(1, 9, n).astype("<u4"), "n1": rng.integers(0, 50, n).astype("<u2"),
                     "count": np.ones(n, "<u4"), "big": (offsets, values),
                     "uid": (np.arange(n + 1, dtype="<u8"), rng.integers(0, 300, n).astype("<u8"))})
    db = v.Database({"tables": [WIDE]}, device=0)
    try:
        upload(db.get_table("wide"), segs, dicts, None)
        for q, flags in (({"dimensions": ["d0", "n1"], "metrics": ["big", "count", "uid"]}, 0),
                         ({"dimensions": ["d0"], "metrics": ["big"], "filter": {"op": "ge", "column": "big", "value": "2"}}, 0),
                         ({"dimensions": ["n1", "d0"], "metrics": ["uid", "big"]}, N.PLAN_FORCE_HASH)):
            q = dict(q, type="aggregate", table="wide")
            out = v.MemoryRowOutput()
            stats = db.query(q, out, now=NOW, flags=flags)
            want = viya_oracle.run_query(WIDE, segs, dicts, q, now=NOW)
            assert sorted(out.rows) == sorted(want["rows"])
            assert stats.aggregated_recs == want["stats"]["aggregated_recs"]
            assert stats.distinct_paths & N.DEDUPE_WIDE
    finally:
        db.close()


# ---- float keys: -0.0 == +0.0 in every table mode (KeyEqual uses ==, store.cc:46-63) ----
FKEY = {"name": "fkey", "segment_size": 5000,
        "dimensions": [{"name": "f", "type": "float"}, {"name": "s"}, {"name": "g", "type": "double"}],
        "metrics": [{"name": "count", "type": "count"}, {"name": "ls", "type": "long_sum"}]}


@pytest.mark.parametrize("dims", [["f"], ["f", "s"], ["s", "f"], ["g"], ["f", "g"]], ids=lambda d: "+".join(d))
def test_float_key_negative_zero(built_lib, dims):
    import viyadb_b200 as v
    rng = np.random.default_rng(11)
    dicts = {"s": ["__exceeded", "x", "y", "z"]}
    segs = []
    for s in range(2):
        n = 5000 if s == 0 else 1234
        segs.append({"f": rng.choice(np.array([0.0, -0.0, 1.5, -1.5, 3.25], "<f4"), n),
                     "s": rng.integers(1, 4, n).astype("<u4"),
                     "g": rng.choice(np.array([0.0, -0.0, 2.5, -2.5], "<f8"), n),
                     "count": np.ones(n, "<u4"), "ls": rng.integers(-1000, 1000, n).astype("<i8")})
    db = v.Database({"tables": [FKEY]}, device=0)
    try:
        upload(db.get_table("fkey"), segs, dicts, None)
        q = {"type": "aggregate", "table": "fkey", "dimensions": dims, "metrics": ["count", "ls"]}
        out = v.MemoryRowOutput()
        stats = db.query(q, out, now=NOW)
        want = viya_oracle.run_query(FKEY, segs, dicts, q, now=NOW)
        # the sign of a zero key that reaches the output is whichever row came first in the reference: compare it unsigned
        norm = lambda rows: sorted([[c[1:] if c in ("-0",) else c for c in r] for r in rows])
        assert norm(out.rows) == norm(want["rows"])
        assert stats.aggregated_recs == want["stats"]["aggregated_recs"]
    finally:
        db.close()


# ---- rolled-up time keys: the bucket dictionary gives the same groups as per-row calendar arithmetic ----
ROLL = {"name": "roll", "segment_size": 60000,
        "dimensions": [{"name": "d0", "cardinality": 300},
                       {"name": "t1", "type": "time",
                        "rollup_rules": [{"granularity": "hour", "after": "1 days"}, {"granularity": "day", "after": "1 weeks"},
                                         {"granularity": "month", "after": "1 years"}]},
                       {"name": "mt", "type": "microtime",
                        "rollup_rules": [{"granularity": "minute", "after": "2 hours"}, {"granularity": "day", "after": "3 days"},
                                         {"granularity": "year", "after": "2 years"}]}],
        "metrics": [{"name": "count", "type": "count"}, {"name": "ls", "type": "long_sum"}, {"name": "uid", "type": "bitset"}]}
ROLL_SPEC = {"d0": (1, 200), "t1": (NOW - 800 * 86400, NOW + 5), "mt": ((NOW - 1100 * 86400) * 1000000, (NOW + 5) * 1000000),
             "count": (1, 3), "ls": (-2**40, 2**40), "uid": ("ids", 2000, 1)}
ROLL_QUERIES = [
    {"select": [{"column": "d0"}, {"column": "t1", "granularity": "hour"}, {"column": "ls"}, {"column": "count"}]},   # the C4 shape
    {"select": [{"column": "t1"}, {"column": "count"}, {"column": "uid"}]},
    {"select": [{"column": "t1", "granularity": "month", "format": "%Y-%m"}, {"column": "count"}],
     "filter": {"op": "gt", "column": "t1", "value": str(NOW - 500 * 86400)}},
    {"select": [{"column": "t1", "granularity": "year"}, {"column": "ls"}]},
    {"select": [{"column": "mt", "granularity": "day"}, {"column": "count"}]},
    {"select": [{"column": "mt", "granularity": "hour"}, {"column": "d0"}, {"column": "count"}]},
    {"select": [{"column": "mt", "granularity": "second"}, {"column": "count"}],
     "filter": {"op": "gt", "column": "mt", "value": str((NOW - 600) * 1000000)}},
    {"select": [{"column": "t1", "granularity": "minute"}, {"column": "mt", "granularity": "month"}, {"column": "count"}]},
]


@pytest.fixture(scope="module")
def roll(built_lib):
    import viyadb_b200 as v
    segs, dicts, hidden = random_table(ROLL, 3, 60000, 77, ROLL_SPEC, last_rows=31337)
    for seg in segs:   # rows right at the rule boundaries and at month / year starts
        edge = np.array([NOW - 86400, NOW - 86400 - 1, NOW - 7 * 86400, NOW - 7 * 86400 - 1, 1464739200, 1464739199, 1483228800, 1483228799, NOW],
                        dtype="<u4")
        seg["t1"][:len(edge)] = edge
        seg["mt"][:len(edge)] = edge.astype("<u8") * 1000000 + np.arange(len(edge), dtype="<u8") * 111111
    db = v.Database({"tables": [ROLL]}, device=0)
    upload(db.get_table("roll"), segs, dicts, hidden)
    yield v, db, segs, dicts, hidden
    db.close()


@pytest.mark.parametrize("mode", ["dictionary", "calendar"])
@pytest.mark.parametrize("qi", range(len(ROLL_QUERIES)))
def test_rollup_bucket_dictionary(roll, qi, mode):
    v, db, segs, dicts, hidden = roll
    q = dict(ROLL_QUERIES[qi], type="aggregate", table="roll")
    db.set_test_hook("tune", 2 | ((1 << 21) if mode == "calendar" else 0))
    try:
        out = v.MemoryRowOutput()
        stats = db.query(q, out, now=NOW)
    finally:
        db.set_test_hook("tune", 2)
    want = viya_oracle.run_query(ROLL, segs, dicts, q, now=NOW, hidden_counts=hidden)
    assert sorted(out.rows) == sorted(want["rows"])
    assert stats.aggregated_recs == want["stats"]["aggregated_recs"]
    if mode == "dictionary" and qi == 0:
        assert stats.table_mode == 0, "the C4 shape must aggregate into a dense table through the bucket dictionary"


# ---- re-entrancy: concurrent queries and a concurrent put on one context ----
def test_concurrent_queries_and_put(mid):
    v, db, segs, dicts, want = mid
    queries = [Q_ALL, Q_SEL,
               {"type": "aggregate", "table": "mid", "dimensions": ["d1"], "metrics": ["mn"],
                "filter": {"op": "eq", "column": "d0", "value": "a7"}},
               {"type": "aggregate", "table": "mid", "dimensions": ["n2"], "metrics": ["uid"],
                "filter": {"op": "lt", "column": "n2", "value": "40"}}]
    serial = []
    for q in queries:
        out = v.MemoryRowOutput()
        db.query(q, out, now=NOW)
        serial.append(sorted(out.rows))
    # a second table on the same context takes puts while the queries run
    other = dict(MID, name="other", segment_size=100000)
    db.create_table(other)
    t2 = db.get_table("other")
    for name, c2v in dicts.items():
        t2.dimension(name).dict.c2v = list(c2v)
    errors, results = [], [[None] * 3 for _ in queries]

    def worker(i):
        try:
            for rep in range(3):
                out = v.MemoryRowOutput()
                db.query(queries[i], out, now=NOW)
                results[i][rep] = sorted(out.rows)
        except Exception as e:   # noqa: BLE001
            errors.append(e)

    def putter():
        try:
            for s in range(6):
                seg = {k: (val[0][:100001], val[1][:100000]) if isinstance(val, tuple) else val[:100000] for k, val in segs[s % len(segs)].items()}
                t2.put_segment(s, seg)
            # and a put on the table under query: replaces segment 4 with identical data (results must not change)
            db.get_table("mid").put_segment(4, segs[4])
        except Exception as e:   # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(len(queries))] + [threading.Thread(target=putter)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    for i in range(len(queries)):
        for rep in range(3):
            assert results[i][rep] == serial[i], (i, rep)
    out = v.MemoryRowOutput()
    stats = db.query({"type": "aggregate", "table": "other", "dimensions": ["d0"], "metrics": ["uid"]}, out, now=NOW)
    assert stats.scanned_recs == 600000


# ---- incremental sync: in-place updates of a row range + appended rows (vgpu_segment_update) ----
def test_segment_update_ranges(built_lib):
    import viyadb_b200 as v
    conf = {"name": "live", "segment_size": 60000,
            "dimensions": [{"name": "d0"}, {"name": "n1", "type": "ushort"}, {"name": "t2", "type": "time"}],
            "metrics": [{"name": "count", "type": "count"}, {"name": "ls", "type": "long_sum"}, {"name": "mx", "type": "int_max"},
                        {"name": "uid", "type": "bitset"}]}
    spec = {"d0": (1, 12), "n1": (0, 300), "t2": (NOW - 90 * 86400, NOW), "count": (1, 3), "ls": (-2**40, 2**40),
            "mx": (-2**31, 2**31 - 1), "uid": ("ids", 4000, 1)}
    segs, dicts, hidden = random_table(conf, 3, 50000, 31, spec, last_rows=20000)
    more, _, _ = random_table(conf, 1, 30000, 32, spec)       # replacement cells and appended rows
    db = v.Database({"tables": [conf]}, device=0)
    try:
        t = db.get_table("live")
        upload(t, segs, dicts, hidden)
        q = {"type": "aggregate", "table": "live", "select": [{"column": "d0"}, {"column": "t2", "granularity": "day"}, {"column": "count"},
                                                               {"column": "ls"}, {"column": "mx"}, {"column": "uid"}],
             "filter": {"op": "lt", "column": "n1", "value": "200"}}

        def check():
            out = v.MemoryRowOutput()
            stats = db.query(q, out, now=NOW)
            want = viya_oracle.run_query(conf, segs, dicts, q, now=NOW)
            assert sorted(out.rows) == sorted(want["rows"])
            for k in ("scanned_segments", "scanned_recs", "aggregated_recs"):
                assert getattr(stats, k) == want["stats"][k], k
        check()
        # 1. in-place update of rows [1000, 9000) of segment 1: metric cells change, dimensions stay (what an upsert does)
        for name in ("count", "ls", "mx"):
            segs[1][name] = segs[1][name].copy()
            segs[1][name][1000:9000] = more[0][name][:8000]
        ids = segs[1]["uid"][1].copy()
        ids[1000:9000] = more[0]["uid"][1][:8000]
        segs[1]["uid"] = (segs[1]["uid"][0], ids)
        part = {k: (val[1000:9000] if not isinstance(val, tuple) else val[1][1000:9000]) for k, val in segs[1].items()}
        t.update_segment(1, 1000, part)
        check()
        # 2. the last segment grows: 7000 rows appended behind its 20000 (one call), then 3 more rows and an update that
        #    straddles old and new rows
        last = segs[2]
        grown = {}
        for k, val in last.items():
            if isinstance(val, tuple):
                grown[k] = (np.arange(27001, dtype="<u8"), np.concatenate([val[1], more[0]["uid"][1][10000:17000]]))
            else:
                grown[k] = np.concatenate([val, more[0][k][10000:17000]])
        segs[2] = grown
        t.update_segment(2, 20000, {k: (val[20000:] if not isinstance(val, tuple) else val[1][20000:]) for k, val in grown.items()})
        check()
        with pytest.raises(v.VgpuError):   # a hole behind the last row is refused
            t.update_segment(2, 27005, {k: (val[:3] if not isinstance(val, tuple) else val[1][:3]) for k, val in grown.items()})
    finally:
        db.close()


# ---- post-aggregation on the device: HAVING on raw accumulators, top-N on the first sort key ----
POST_QUERIES = [
    # (query, expect HAVING on device, expect top-N on device)
    ({"dimensions": ["d1"], "metrics": ["count", "ls"], "having": {"op": "gt", "column": "count", "value": "3300"}}, True, False),
    ({"dimensions": ["d1", "d0"], "metrics": ["count", "ls"], "having": {"op": "and", "filters": [
        {"op": "gt", "column": "count", "value": "90"}, {"op": "lt", "column": "ls", "value": "0"},
        {"op": "in", "column": "d0", "values": ["d0_3", "d0_5", "d0_8"]}]}}, True, False),
    ({"dimensions": ["d1"], "metrics": ["count", "ls"], "sort": [{"column": "count"}, {"column": "d1", "ascending": True}], "limit": 7, "skip": 2}, False, True),
    ({"dimensions": ["d1", "d0"], "metrics": ["ls", "is"], "sort": [{"column": "ls", "ascending": True}], "limit": 25}, False, True),     # signed: SmallerInt order
    ({"dimensions": ["d1", "d0"], "metrics": ["ls", "is"], "sort": [{"column": "is"}], "limit": 40, "skip": 5}, False, True),              # signed, descending
    ({"dimensions": ["i4"], "metrics": ["count"], "sort": [{"column": "i4", "ascending": True}], "limit": 30}, False, True),               # signed numeric dimension
    ({"dimensions": ["d0", "n3"], "metrics": ["uid", "imax"], "having": {"op": "ge", "column": "uid", "value": "12"},
      "sort": [{"column": "uid"}, {"column": "n3", "ascending": True}, {"column": "d0"}], "limit": 15}, True, True),                       # count-distinct in HAVING and sort
    ({"dimensions": ["d0"], "metrics": ["savg", "count"], "having": {"op": "gt", "column": "savg", "value": "1000"},
      "sort": [{"column": "savg"}], "limit": 5}, True, False),                                                                               # AVG: HAVING on the raw sum, no device top-N
    ({"dimensions": ["d1"], "metrics": ["ds"], "sort": [{"column": "ds"}], "limit": 5}, False, False),                                       # float sort key: the host's
    ({"dimensions": ["d0"], "metrics": ["count"], "sort": [{"column": "d0", "ascending": True}], "limit": 5}, False, False),                 # string sort key: the host's
    ({"dimensions": ["d1"], "metrics": ["count"], "having": {"op": "gt", "column": "count", "value": "1650"}, "limit": 3}, False, False),    # window before HAVING: the host's
    ({"dimensions": ["d2", "n3"], "metrics": ["count", "umax"], "having": {"op": "ge", "column": "umax", "value": "4294000000"},
      "sort": [{"column": "umax"}, {"column": "d2"}, {"column": "n3"}], "limit": 100}, True, True),                                         # hashed table (1.6e5 groups)
]


@pytest.fixture(scope="module")
def post_env(built_lib):
    import viyadb_b200 as v
    import test_gpu_parity as TP
    segs, dicts, hidden = random_table(TP.EVENTS, 4, 50000, 1234, TP.SPEC, last_rows=12345)
    db = v.Database({"tables": [TP.EVENTS]}, device=0)
    upload(db.get_table("events"), segs, dicts, hidden)
    yield v, db, segs, dicts, hidden, TP.EVENTS
    db.close()


@pytest.mark.parametrize("qi", range(len(POST_QUERIES)))
def test_device_having_and_top_n(post_env, qi):
    v, db, segs, dicts, hidden, conf = post_env
    qd, want_having, want_topn = POST_QUERIES[qi]
    q = dict(qd, type="aggregate", table="events")
    want = viya_oracle.run_query(conf, segs, dicts, q, now=NOW, hidden_counts=hidden)
    results = {}
    for device_post in (True, False):
        out = v.MemoryRowOutput()
        stats = db.query(q, out, now=NOW, device_post=device_post)
        results[device_post] = (out.rows, stats)
        windowed = not q.get("sort") and (q.get("limit") or q.get("skip"))   # which groups fall in the window is unspecified (Q11)
        for k in ("scanned_segments", "scanned_recs", "aggregated_recs") + (() if windowed else ("output_recs",)):
            assert getattr(stats, k) == want["stats"][k], (device_post, k, getattr(stats, k), want["stats"][k])
    rows, stats = results[True]
    assert bool(stats.post_applied & 1) == want_having and bool(stats.post_applied & 2) == want_topn, stats.post_applied
    assert results[False][1].post_applied == 0
    if q.get("sort"):
        # ties of the sort key are unordered in std::sort: compare the sort-key columns row by row, the rest as sets of
        # the rows that are not tied at the cut
        names = q["dimensions"] + q["metrics"]
        cols = [names.index(sc["column"]) for sc in q["sort"]]
        def same(a, b):   # double sums: the order of the additions differs (1e-12 relative, as everywhere)
            if a == b:
                return True
            try:
                x, y = float(a), float(b)
            except ValueError:
                return False
            return abs(x - y) <= 1e-12 * max(abs(x), abs(y))
        for got in (rows, results[False][0]):
            assert len(got) == len(want["rows"])
            for r, w in zip(got, want["rows"]):
                assert all(same(r[c], w[c]) for c in cols), (r, w)
    elif windowed:
        assert len(rows) <= q["limit"] and len(results[False][0]) <= q["limit"]
    else:
        assert sorted(rows) == sorted(want["rows"])
        assert sorted(results[False][0]) == sorted(want["rows"])
