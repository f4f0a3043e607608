"""The device path's pure arithmetic, run on the CPU: viyadb_b200/csrc/device_arith.h (the source the CUDA kernels
compile) and csrc/time_dict.h (the planner's bucket dictionary) are compiled with plain g++ (tests/device_arith_harness.cc)
and checked against glibc's gmtime_r / timegm (what util::Truncator calls, src/util/time.h:52-89), against the oracle's
rollup (oracle/viya_oracle.py: rollup_key, pinned to the reference's time.cc captures) and against the string order of
util::StringNumCmp::SmallerInt (src/util/string.h:28-49).

This is a checker of arithmetic, not a CPU path: the harness exports no query entry point and libvgpu.so has none."""
import ctypes as C
import os
import shutil
import subprocess
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import viya_oracle  # noqa: E402

YEAR, MONTH, WEEK, DAY, HOUR, MINUTE, SECOND, NONE = range(8)
U64P = C.POINTER(C.c_uint64)


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    gxx = os.environ.get("VGPU_HARNESS_CXX") or shutil.which("g++")   # VGPU_HARNESS_CXX / _FLAGS: sanitizer builds
    extra = os.environ.get("VGPU_HARNESS_FLAGS", "").split()
    if gxx is None:
        pytest.skip("g++ not available")
    so = str(tmp_path_factory.mktemp("arith") / "libdevice_arith.so")
    subprocess.run([gxx, "-std=c++17", "-O2", "-shared", "-fPIC", *extra, "-Wno-unknown-pragmas", "-o", so,
                    os.path.join(ROOT, "tests", "device_arith_harness.cc")], check=True)
    return C.CDLL(so)


def _p(a, t=U64P):
    return a.ctypes.data_as(t)


def _u64(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.uint64))


def trunc_seconds(lib, t, unit, wide):
    t = _u64(t)
    out = np.empty_like(t)
    lib.h_trunc_seconds(_p(t), C.c_uint64(len(t)), C.c_uint32(unit), C.c_int(1 if wide else 0), _p(out))
    return out


def glibc_trunc(lib, t, unit):
    t = _u64(t)
    out = np.empty_like(t)
    lib.h_glibc_trunc(_p(t), C.c_uint64(len(t)), C.c_uint32(unit), _p(out))
    return out


def rollup(lib, v, micro, rules, query_unit):
    v = _u64(v)
    out = np.empty_like(v)
    units = np.array([u for _, u in rules], dtype=np.uint8)
    bounds = _u64([b for b, _ in rules])
    lib.h_rollup(_p(v), C.c_uint64(len(v)), C.c_int(int(micro)), C.c_uint32(len(rules)), _p(units, C.POINTER(C.c_uint8)),
                 _p(bounds), C.c_uint32(query_unit), _p(out))
    return out


def time_dict(lib, v, micro, rules, query_unit, lo, hi, max_values=1 << 23):
    """-> None when the planner declines, else (ranks of v, bucket values, npieces, narrow)."""
    v = _u64(v)
    ranks = np.empty_like(v)
    values = np.empty(max_values, dtype=np.uint64)
    units = np.array([u for _, u in rules], dtype=np.uint8)
    bounds = _u64([b for b, _ in rules])
    npieces, narrow = C.c_uint32(0), C.c_uint32(0)
    lib.h_time_dict.restype = C.c_longlong
    n = lib.h_time_dict(_p(v), C.c_uint64(len(v)), C.c_int(int(micro)), C.c_uint32(len(rules)), _p(units, C.POINTER(C.c_uint8)),
                        _p(bounds), C.c_uint32(query_unit), C.c_uint64(lo), C.c_uint64(hi), _p(ranks), _p(values),
                        C.c_uint64(max_values), C.byref(npieces), C.byref(narrow))
    assert n >= 0, "more bucket values than the test buffer holds"
    if n == 0:
        return None
    return ranks, values[:n].copy(), npieces.value, narrow.value


# ------------------------------------------------------------------------------------------------
# calendar truncation == glibc gmtime_r / timegm
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("unit", [YEAR, MONTH, DAY, HOUR, MINUTE, SECOND, NONE])
def test_trunc_seconds_32_equals_glibc(lib, unit):
    rng = np.random.default_rng(unit + 1)
    t = np.concatenate([rng.integers(0, 2**32, 400_000, dtype=np.uint64),
                        _u64([0, 1, 59, 60, 3599, 3600, 86399, 86400, 2**31 - 1, 2**31, 2**32 - 1]),
                        # the last / first seconds of months and years, leap days included (1972, 2000, 2100 is none)
                        _u64([68169599, 68169600, 68255999, 68256000, 951782399, 951782400, 951868799, 951868800,
                              4107542399, 4107542400, 4107628800, 946684799, 946684800, 978307199, 978307200])])
    assert np.array_equal(trunc_seconds(lib, t, unit, wide=False), glibc_trunc(lib, t, unit))


@pytest.mark.parametrize("unit", [YEAR, MONTH, DAY, HOUR, MINUTE, SECOND, NONE])
def test_trunc_seconds_64_equals_glibc(lib, unit):
    """util::Time64 holds microseconds in a uint64: its seconds part reaches 2^64 / 1e6 = 1.8e13 (year ~586000)."""
    rng = np.random.default_rng(unit + 11)
    top = 2**64 // 1_000_000
    t = np.concatenate([rng.integers(0, top, 200_000, dtype=np.uint64), rng.integers(0, 2**33, 200_000, dtype=np.uint64),
                        _u64([0, 2**32 - 1, 2**32, 2**32 + 1, 253402300799, 253402300800, top - 1, top])])
    assert np.array_equal(trunc_seconds(lib, t, unit, wide=True), glibc_trunc(lib, t, unit))


def test_trunc_seconds_every_day_boundary_of_the_32_bit_range(lib):
    """every midnight of 1970..2106 and the second before it: month and year truncation"""
    mid = np.arange(0, 2**32, 86400, dtype=np.uint64)
    t = np.concatenate([mid, mid[1:] - 1])
    for unit in (YEAR, MONTH):
        assert np.array_equal(trunc_seconds(lib, t, unit, wide=False), glibc_trunc(lib, t, unit))
        assert np.array_equal(trunc_seconds(lib, t, unit, wide=True), glibc_trunc(lib, t, unit))


def test_glibc_truncation_matches_the_oracle(lib):
    """the harness's host_trunc and the oracle's _trunc_seconds are two statements of util::Truncator"""
    rng = np.random.default_rng(5)
    t = rng.integers(0, 2**32, 3000, dtype=np.uint64)
    for unit in (YEAR, MONTH, DAY, HOUR, MINUTE):
        want = [viya_oracle._trunc_seconds(int(x), unit) for x in t]
        assert glibc_trunc(lib, t, unit).tolist() == want


# ------------------------------------------------------------------------------------------------
# rollup_value == the oracle's rollup_key (scan.cc:198-219, rollup.cc:77-95)
# ------------------------------------------------------------------------------------------------
def _random_rules(rng, now, micro, monotone):
    """rule list [(boundary, unit)] in TimeDimension order (descending `after`, i.e. ascending boundary)."""
    n = int(rng.integers(0, 5))
    spans = sorted((int(x) for x in rng.integers(3600, 3 * 365 * 86400, n)), reverse=True)
    units = [int(u) for u in rng.choice([YEAR, MONTH, DAY, HOUR, MINUTE, SECOND], n)]
    if monotone:
        units.sort()   # coarser rules for older data, as every sensible configuration has them
    rules = []
    for span, unit in zip(spans, units):
        b = now - span
        if micro:
            b *= 1_000_000
        rules.append((b, unit))
    return rules


@pytest.mark.parametrize("micro", [False, True])
def test_rollup_value_equals_the_oracle(lib, micro):
    rng = np.random.default_rng(21 + micro)
    now = 1496570140
    for case in range(40):
        rules = _random_rules(rng, now, micro, monotone=bool(case % 2))
        query_unit = int(rng.choice([NONE, NONE, YEAR, MONTH, DAY, HOUR, MINUTE, SECOND]))
        secs = np.concatenate([rng.integers(now - 4 * 365 * 86400, now + 86400, 3000),
                               np.array([b // (1_000_000 if micro else 1) + d for b, _ in rules for d in (-1, 0, 1)], dtype=np.int64)])
        v = secs.astype(np.uint64)
        if micro:
            v = v * np.uint64(1_000_000) + rng.integers(0, 1_000_000, len(v)).astype(np.uint64)
        dim = types.SimpleNamespace(micro=micro)
        want = viya_oracle.rollup_key(v, dim, rules, None if query_unit == NONE else query_unit)
        got = rollup(lib, v, micro, rules, query_unit)
        assert np.array_equal(got, want.astype(np.uint64)), (case, rules, query_unit)


# ------------------------------------------------------------------------------------------------
# the bucket dictionary: values[tdict_rank(v)] == rollup_value(v), ranks ordered like the values
# ------------------------------------------------------------------------------------------------
def _check_dict(lib, v, micro, rules, query_unit, lo, hi):
    d = time_dict(lib, v, micro, rules, query_unit, lo, hi)
    if d is None:
        return None
    ranks, values, npieces, narrow = d
    assert 1 <= npieces <= 48
    assert np.all(values[1:] > values[:-1]), "bucket values must be strictly increasing: the rank IS the order"
    assert ranks.max() < len(values)
    want = rollup(lib, v, micro, rules, query_unit)
    assert np.array_equal(values[ranks.astype(np.int64)], want), (rules, query_unit)
    return len(values)


def test_time_dict_c4_every_second(lib):
    """C4 (BASELINE.json configs[4]): rules hour > 1 day, day > 1 week, month > 1 year, query granularity hour, two years
    of seconds before the pinned clock — EVERY second of the range goes through the dictionary and through the per-row
    calendar arithmetic; both must name the same bucket. 541 buckets: 13 months, 358 days, 145 + 25 hours (the bench sees
    ~530.5 per d0 value: an hour bucket holds 2.85 rows per d0 on average, so ~6 % of them stay empty)."""
    sys.path.insert(0, ROOT)
    from viyadb_b200.timeutil import Duration
    now = 1496570140
    rules = [(Duration("1 years").add_to(now, -1), MONTH), (Duration("1 weeks").add_to(now, -1), DAY),
             (Duration("1 days").add_to(now, -1), HOUR)]
    lo, hi = now - 730 * 86400, now - 1
    seen = set()
    nvalues = None
    for a in range(lo, hi + 1, 8_000_000):
        v = np.arange(a, min(a + 8_000_000, hi + 1), dtype=np.uint64)
        nvalues = _check_dict(lib, v, False, rules, HOUR, lo, hi)
        assert nvalues is not None, "the planner must not decline C4's dictionary"
        seen.update(np.unique(rollup(lib, v, False, rules, HOUR)).tolist())
    assert nvalues == len(seen) == 541   # every dictionary entry is attainable, none is missing


@pytest.mark.parametrize("micro", [False, True])
def test_time_dict_random_rules(lib, micro):
    rng = np.random.default_rng(77 + micro)
    now = 1496570140
    scale = 1_000_000 if micro else 1
    built = declined = 0
    for case in range(120):
        rules = _random_rules(rng, now, micro, monotone=case % 4 != 3)
        query_unit = int(rng.choice([NONE, YEAR, MONTH, DAY, HOUR, MINUTE, SECOND]))
        span = int(rng.choice([3 * 3600, 5 * 86400, 400 * 86400, 4 * 365 * 86400]))
        hi_s = now + int(rng.integers(-86400, 86400))
        lo_s = hi_s - span
        secs = np.concatenate([rng.integers(lo_s, hi_s + 1, 20000),
                               np.array([lo_s, hi_s] + [min(max(b // scale + d, lo_s), hi_s) for b, _ in rules for d in (-1, 0, 1)],
                                        dtype=np.int64)])
        v = secs.astype(np.uint64) * np.uint64(scale)
        if micro:
            v = v + rng.integers(0, 1_000_000, len(v)).astype(np.uint64)
        lo, hi = int(v.min()), int(v.max())
        n = _check_dict(lib, v, micro, rules, query_unit, lo, hi)
        if n is None:
            declined += 1
        else:
            built += 1
    assert built >= 40, (built, declined)   # the property above was really exercised


def test_time_dict_declines_what_it_cannot_number(lib):
    now = 1496570140
    v = _u64([now - 10, now - 5])
    # a finer rule for older data than for newer: the truncated value is not monotone in the raw value
    assert time_dict(lib, v, False, [(now - 86400 * 30, HOUR), (now - 86400, MONTH)], NONE, now - 86400 * 400, now) is None
    # microseconds with nothing truncating them: the key keeps its raw domain
    assert time_dict(lib, v * np.uint64(1_000_000), True, [], NONE, (now - 100) * 1_000_000, now * 1_000_000) is None
    # week granularity: the reference cannot link it (Q10)
    assert time_dict(lib, v, False, [], WEEK, now - 100, now) is None
    # second granularity over two years: more buckets than the dense domain takes
    assert time_dict(lib, v, False, [], SECOND, now - 730 * 86400, now) is None


# ------------------------------------------------------------------------------------------------
# smaller_int_rank == rank under util::StringNumCmp::SmallerInt (length, then lexicographic)
# ------------------------------------------------------------------------------------------------
def _smaller_int_key(x):
    s = str(int(x))
    return (len(s), s)


def test_smaller_int_rank_orders_like_the_string_comparison(lib):
    rng = np.random.default_rng(9)
    digits = rng.integers(1, 20, 200_000)
    mag = np.array([int(rng.integers(10 ** (d - 1) if d > 1 else 0, min(10 ** d, 2**63))) for d in digits[:20000].tolist()], dtype=object)
    sign = rng.choice([-1, 1], len(mag))
    xs = [int(m) * int(s) for m, s in zip(mag, sign)]
    edges = [0, 1, -1, 9, -9, 10, -10, 99, -99, 100, -100, 2**31 - 1, -2**31, 2**32, 2**63 - 1, -2**63, -2**63 + 1,
             10**18, -10**18, 10**18 - 1, -(10**18 - 1), 9223372036854775807, 999999999999999999, -999999999999999999]
    xs = np.array(sorted(set(xs + edges + [e + d for e in edges for d in (-1, 1) if -2**63 <= e + d < 2**63])), dtype=np.int64)
    out = np.empty(len(xs), dtype=np.uint64)
    lib.h_smaller_int_rank(_p(xs, C.POINTER(C.c_longlong)), C.c_uint64(len(xs)), _p(out))
    assert len(set(out.tolist())) == len(xs), "the rank must be injective"
    by_rank = xs[np.argsort(out, kind="stable")].tolist()
    by_string = sorted(xs.tolist(), key=_smaller_int_key)
    assert by_rank == by_string
    # and it is a rank: consecutive integers of one sign and length are consecutive ranks
    r = dict(zip(xs.tolist(), out.tolist()))
    assert r[0] == 0 and r[1] == 1 and r[9] == 9
    assert r[-1] == 10 and r[-9] == 18 and r[10] == 19   # "-1".."-9" have length 2 and sort before "10"
    # the oracle's comparator is the same order
    for a, b in rng.choice(xs, (5000, 2)).tolist():
        assert (r[a] < r[b]) == viya_oracle._smaller_int(str(a), str(b))


# ------------------------------------------------------------------------------------------------
# column statistics order image, -0.0 keys, hash mixers, the generator's stream
# ------------------------------------------------------------------------------------------------
def test_to_ordered_preserves_the_order_of_every_type(lib):
    rng = np.random.default_rng(13)

    def ordered(raw, t):
        raw = _u64(raw)
        out = np.empty_like(raw)
        lib.h_to_ordered(_p(raw), C.c_uint64(len(raw)), C.c_uint32(t), _p(out))
        return out

    # signed integers arrive sign-extended to 64 bits (types 4..7 = i8, i16, i32, i64 in include/vgpu.h)
    x = np.concatenate([rng.integers(-2**63, 2**63, 100_000), np.array([-2**63, -1, 0, 1, 2**63 - 1])]).astype(np.int64)
    o = ordered(x.view(np.uint64), 7)
    idx = np.argsort(x, kind="stable")
    assert np.all(np.diff(o[idx].astype(object)) >= 0) and len(np.unique(o)) == len(np.unique(x))
    # float / double: numeric order, -0.0 == +0.0, NaN excluded (the reference compares with <, never orders NaN)
    f = np.concatenate([rng.standard_normal(100_000) * 1e6, np.array([0.0, -0.0, np.inf, -np.inf, 1e-45, -1e-45])]).astype(np.float32)
    of = ordered(f.view(np.uint32).astype(np.uint64), 8)
    idx = np.argsort(f, kind="stable")
    assert np.all(np.diff(of[idx].astype(object)) >= 0)
    assert of[f == 0].min() == of[f == 0].max()
    d = np.concatenate([rng.standard_normal(100_000) * 1e12, np.array([0.0, -0.0, np.inf, -np.inf, 5e-324, -5e-324])])
    od = ordered(d.view(np.uint64), 9)
    idx = np.argsort(d, kind="stable")
    assert np.all(np.diff(od[idx].astype(object)) >= 0)
    assert od[d == 0].min() == od[d == 0].max()
    # strictness: different values, different images
    assert len(np.unique(od)) == len(np.unique(d)) and len(np.unique(of)) == len(np.unique(f))
    # unsigned types pass through
    u = rng.integers(0, 2**64, 1000, dtype=np.uint64)
    assert np.array_equal(ordered(u, 3), u)


def test_fzero_fix(lib):
    def fix(raw, width):
        raw = _u64(raw)
        out = np.empty_like(raw)
        lib.h_fzero_fix(_p(raw), C.c_uint64(len(raw)), C.c_uint32(width), _p(out))
        return out.tolist()

    assert fix([0x80000000, 0, 0x3F800000, 0xBF800000, 0x80000001], 4) == [0, 0, 0x3F800000, 0xBF800000, 0x80000001]
    assert fix([0x8000000000000000, 0, 0x3FF0000000000000, 0x8000000000000001, 0x80000000], 8) == \
        [0, 0, 0x3FF0000000000000, 0x8000000000000001, 0x80000000]


def test_splitmix64_known_answers(lib):
    """the published splitmix64 stream from state 0 (Vigna): the synthetic generator of the bench and of oracle_cli"""
    gamma = 0x9E3779B97F4A7C15
    x = _u64([(k * gamma) & (2**64 - 1) for k in range(3)])
    out = np.empty_like(x)
    lib.h_splitmix64(_p(x), C.c_uint64(len(x)), _p(out))
    assert out.tolist() == [0xE220A8397B1DCDAF, 0x6E789E6AA1B965F4, 0x06C45D188009454F]


def test_hash_mixers_are_bijective_and_balanced(lib):
    rng = np.random.default_rng(17)
    # mix64 (murmur3 finaliser) is a bijection: no two keys may collide before the mask is applied
    x = rng.integers(0, 2**64, 500_000, dtype=np.uint64)
    x = np.unique(x)
    out = np.empty_like(x)
    lib.h_mix64(_p(x), C.c_uint64(len(x)), _p(out))
    assert len(np.unique(out)) == len(x)
    # count-distinct pairs as the scan writes them: cell << 32 | id, dense cells, ids of a small domain
    cell = rng.integers(0, 500_000, 1_000_000).astype(np.uint64)
    ident = rng.integers(0, 1_000_000, 1_000_000).astype(np.uint64)
    key = (cell << np.uint64(32)) | ident
    for nb in (16, 4096, 16384):
        b = np.empty(len(key), dtype=np.uint32)
        lib.h_pair_bucket(_p(key), C.c_uint64(len(key)), C.c_uint32(nb), _p(b, C.POINTER(C.c_uint32)))
        counts = np.bincount(b, minlength=nb)
        assert b.max() < nb
        mean = len(key) / nb
        assert counts.max() < mean + 6 * np.sqrt(mean) + 1, (nb, counts.max(), mean)   # Poisson tail: what bucket_cap assumes
    for nranks in (2, 3, 4, 8):
        o = np.empty(len(key), dtype=np.uint32)
        lib.h_pair_owner(_p(cell), _p(ident), C.c_uint64(len(key)), C.c_uint32(nranks), _p(o, C.POINTER(C.c_uint32)))
        counts = np.bincount(o, minlength=nranks)
        assert o.max() < nranks
        assert abs(counts - len(key) / nranks).max() < 6 * np.sqrt(len(key) / nranks)
        # the owner is a function of the pair alone: every copy of a pair meets on one rank
        o2 = np.empty(len(key), dtype=np.uint32)
        lib.h_pair_owner(_p(cell[::-1].copy()), _p(ident[::-1].copy()), C.c_uint64(len(key)), C.c_uint32(nranks), _p(o2, C.POINTER(C.c_uint32)))
        assert np.array_equal(o2[::-1], o)
