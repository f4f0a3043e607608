"""Host-side mirror of the reference's query layer (no GPU): filter factory (NOT push-down, precedence
sort), literal decoding, plan lowering, durations — against the reference's documented behaviour."""
import pytest

import viyadb_b200 as v
from viyadb_b200 import _native as N
from viyadb_b200.query import (AggregateQuery, CompositeFilter, FilterArgsPacker, FilterFactory, GpuQueryRunner,
                               InFilter, MemoryRowOutput, QueryFactory, RelOpFilter, parse_number)
from viyadb_b200.timeutil import Duration, parse_time_literal

TABLE = {"name": "events",
         "dimensions": [{"name": "country"}, {"name": "tiny", "cardinality": 100}, {"name": "n", "type": "uint"},
                        {"name": "b", "type": "byte"}, {"name": "t", "type": "time",
                                                         "rollup_rules": [{"granularity": "hour", "after": "1 days"},
                                                                          {"granularity": "month", "after": "1 years"},
                                                                          {"granularity": "day", "after": "1 weeks"}]},
                        {"name": "mt", "type": "microtime"}, {"name": "flag", "type": "boolean"}],
         "metrics": [{"name": "count", "type": "count"}, {"name": "rev", "type": "double_sum"}, {"name": "avg", "type": "int_avg"},
                     {"name": "uid", "type": "bitset", "max": 1000}]}


@pytest.fixture()
def db():
    return v.Database({"tables": [TABLE]}, device=None)


def test_schema_types(db):
    t = db.get_table("events")
    types = {c.name: c.type for c in t.columns()}
    assert types["country"] == N.U32 and types["tiny"] == N.U8       # column.cc:54-62, default cardinality -> u32
    assert types["t"] == N.U32 and types["mt"] == N.U64 and types["flag"] == N.U8
    assert types["count"] == N.U32 and types["uid"] == N.U16
    assert not t.has_hidden_count
    # rollup rules sorted by descending `after` (column.cc:346-349)
    assert [r.granularity for r in t.dimension("t").rollup_rules] == [1, 3, 4]


def test_not_pushdown_and_precedence_sort():
    f = FilterFactory.create({"op": "not", "filter": {"op": "and", "filters": [
        {"op": "in", "column": "country", "values": ["US"]},
        {"op": "or", "filters": [{"op": "gt", "column": "n", "value": "5"}, {"op": "eq", "column": "flag", "value": "true"}]},
        {"op": "lt", "column": "n", "value": "3"}]}})
    assert isinstance(f, CompositeFilter) and f.op == "or"
    kinds = [type(c).__name__ for c in f.filters]
    assert kinds == ["RelOpFilter", "CompositeFilter", "InFilter"]          # precedence 1 < 2 < 4
    assert f.filters[0].op == "ge" and f.filters[2].equal is False
    inner = f.filters[1]
    assert inner.op == "and" and [c.op for c in inner.filters] == ["le", "ne"]
    with pytest.raises(ValueError, match="Unsupported filter operataor"):
        FilterFactory.create({"op": "like", "column": "x", "value": "y"})


def test_args_packer_order_and_missing_dictionary_value(db):
    t = db.get_table("events")
    t.dimension("country").dict.encode("US")
    f = FilterFactory.create({"op": "and", "filters": [{"op": "in", "column": "country", "values": ["US", "nope"]},
                                                       {"op": "eq", "column": "b", "value": "-3"},
                                                       {"op": "eq", "column": "flag", "value": "true"}]})
    p = FilterArgsPacker(t).visit(f)
    # RelOps first (precedence), then the IN values; missing value -> UINT32_MAX (dictionary.cc:46-75)
    assert p.values == [-3, 1, 1, 0xFFFFFFFF]
    assert p.args[0] & 0xFF == 0xFD
    kinds = [n[0] for n in p.nodes]
    assert kinds == [N.NODE_RELOP, N.NODE_RELOP, N.NODE_IN, N.NODE_AND]


def test_number_and_time_literals():
    assert parse_number("300", N.U8) == 44 and parse_number("-1", N.U32) == 0xFFFFFFFF
    assert parse_number(" 12abc", N.I32) == 12
    with pytest.raises(ValueError):
        parse_number("abc", N.I32)
    assert parse_time_literal("1420107084", False) == 1420107084
    assert parse_time_literal("2015-01-01", False) == 1420070400
    assert parse_time_literal("2015-01-01 10:11:24", True) == 1420107084 * 1000000
    with pytest.raises(ValueError, match="Unrecognized time format"):
        parse_time_literal("yesterday", False)
    with pytest.raises(ValueError):
        parse_time_literal("2015-01-01 10:11:24.5", True)      # std::stoul(".5") throws in the reference


def test_duration_calendar_arithmetic():
    now = 1496570140            # 2017-06-04 09:55:40 UTC, the clock test/time.cc pins
    assert Duration("1 days").add_to(now, -1) == now - 86400
    assert Duration("1 weeks").add_to(now, -1) == now - 7 * 86400
    import calendar
    assert Duration("1 years").add_to(now, -1) == calendar.timegm((2016, 6, 4, 9, 55, 40))
    assert Duration("3 months").add_to(now, -1) == calendar.timegm((2017, 3, 4, 9, 55, 40))
    assert Duration("40 days").add_to(now, 1) == calendar.timegm((2017, 7, 14, 9, 55, 40))   # mday overflow normalised
    with pytest.raises(ValueError):
        Duration("0 days")


def test_plan_lowering(db):
    t = db.get_table("events")
    q = QueryFactory.create({"type": "aggregate", "table": "events",
                             "select": [{"column": "country"}, {"column": "t", "granularity": "day"}, {"column": "avg"}],
                             "filter": {"op": "ge", "column": "t", "value": "2017-01-01"}}, db)
    r = GpuQueryRunner(db, MemoryRowOutput(), now=1496570140)
    plan = r.build_plan(q)
    # AVG selected without a COUNT *selected* -> the generated code reads the hidden _count (scan.cc:239-241);
    # this table has a COUNT metric, hence no hidden column: the library rejects the plan like g++ rejects the JIT code
    assert plan.nkeys == 2 and plan.nmetrics == 1 and plan.need_hidden_count == 1
    k = plan.keys[1]
    assert k.nrules == 3 and k.query_granularity == N.TU_DAY
    import calendar
    assert [k.rule_boundary[i] for i in range(3)] == [calendar.timegm((2016, 6, 4, 9, 55, 40)), 1496570140 - 7 * 86400, 1496570140 - 86400]
    assert plan.args[0] & 0xFFFFFFFF == 1483228800


def test_query_validation_errors(db):
    with pytest.raises(ValueError, match="is not selected"):
        QueryFactory.create({"type": "aggregate", "table": "events", "dimensions": ["country"], "metrics": ["count"],
                             "sort": [{"column": "rev"}]}, db)
    with pytest.raises(ValueError, match="is not selected"):
        QueryFactory.create({"type": "aggregate", "table": "events", "dimensions": ["country"], "metrics": ["count"],
                             "having": {"op": "gt", "column": "rev", "value": "1"}}, db)
    with pytest.raises(ValueError, match="No such column"):
        QueryFactory.create({"type": "aggregate", "table": "events", "select": [{"column": "nope"}]}, db)
    with pytest.raises(NotImplementedError):
        QueryFactory.create({"type": "show", "table": "events"}, db)
    with pytest.raises(KeyError):      # config.str("dimension") throws in the reference too (query.cc:141)
        QueryFactory.create({"type": "search", "table": "events"}, db)


def test_device_post_aggregation_is_only_asked_for_where_the_reference_tests_every_group(db):
    """vgpu_plan post fields (include/vgpu.h): HAVING travels to the device with a sort, or without skip / limit — without
    a sort the reference cuts the skip / limit window out of the map iteration BEFORE it tests HAVING (post_agg.cc:26-83);
    the first sort column travels with skip + limit; device_post=False leaves everything to the host."""
    t = db.get_table("events")
    t.dimension("country").dict.encode("US")

    def plan_of(conf, **kw):
        q = QueryFactory.create(dict(conf, type="aggregate", table="events"), db)
        return GpuQueryRunner(db, MemoryRowOutput(), **kw).build_plan(q)

    base = {"dimensions": ["country", "n"], "metrics": ["count", "avg"]}
    having = {"op": "and", "filters": [{"op": "gt", "column": "count", "value": "5"}, {"op": "eq", "column": "country", "value": "US"}]}
    p = plan_of(dict(base, having=having))
    assert p.flags & N.PLAN_POST and p.nhnodes == 3 and p.nhargs == 2 and p.sort_col == N.NO_COLUMN and p.top_k == 0
    # the HAVING leaves name schema columns of SELECTED columns, literals in FilterArgsPacker order
    assert [p.hnodes[i].col for i in range(2)] == [t.schema_index(t.metric("count")), t.schema_index(t.dimension("country"))]
    p = plan_of(dict(base, having=having, limit=3))                       # window before HAVING: the host's
    assert p.nhnodes == 0 and p.top_k == 0
    p = plan_of(dict(base, having=having, limit=3, skip=2, sort=[{"column": "count"}, {"column": "n", "ascending": True}]))
    assert p.nhnodes == 3 and p.sort_col == t.schema_index(t.metric("count")) and p.sort_descending == 1 and p.top_k == 5
    p = plan_of(dict(base, sort=[{"column": "n", "ascending": True}]))    # a sort without limit: nothing to cut
    assert p.sort_col == N.NO_COLUMN and p.top_k == 0
    p = plan_of(dict(base, having=having, sort=[{"column": "count"}], limit=3), device_post=False)
    assert not (p.flags & N.PLAN_POST) and p.nhnodes == 0 and p.top_k == 0
