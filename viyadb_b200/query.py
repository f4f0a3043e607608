"""Host-side mirror of the reference's query layer for the aggregate path, ending in the C ABI.

Same names, argument meaning and error behaviour as the reference so that the parity tests read
like test/aggregation.cc & co.:

  FilterFactory / RelOpFilter / InFilter / CompositeFilter / EmptyFilter   src/query/filter.{h,cc}
  AggregateQuery (+ DimOutputColumn / MetricOutputColumn / SortColumn)     src/query/query.{h,cc}
  QueryFactory                                                             src/query/query.cc:150-173
  FilterArgsPacker / ValueDecoder                                          src/codegen/query/filter.cc:100-204
  GpuQueryRunner.visit(AggregateQuery)   replaces   QueryRunner::Visit     src/query/runner.cc:45-64
  post-aggregation (having, decode, format, sort, skip/limit)              src/codegen/query/post_agg.cc:26-147,
                                                                           sort.cc:24-73, header.cc:23-41
  MemoryRowOutput, QueryStats                                              src/query/output.h, stats.h

The scan itself (segment loop, predicate, key build with time rollup, Update()) is NOT here: it is
vgpu_query_agg in libvgpu.so. Python only lowers the query to a vgpu_plan and formats the groups.
"""
import ctypes as C
import os
import struct
import time as _time

import numpy as np

from . import _native as N
from .timeutil import parse_time_literal, time_unit_by_name


# ------------------------------------------------------------------------------------------------
# filters (src/query/filter.h:38-164, filter.cc:36-108)
# ------------------------------------------------------------------------------------------------
class Filter:
    precedence = 0


class RelOpFilter(Filter):
    precedence = 1
    OPS = ["eq", "ne", "lt", "le", "gt", "ge"]          # Operator order == vgpu_relop
    NEGATED = {"eq": "ne", "ne": "eq", "lt": "ge", "le": "gt", "gt": "le", "ge": "lt"}

    def __init__(self, op, column, value):
        self.op, self.column, self.value = op, column, value


class InFilter(Filter):
    precedence = 4

    def __init__(self, column, values, equal=True):
        self.column, self.values, self.equal = column, list(values), equal


class CompositeFilter(Filter):
    def __init__(self, op, filters):
        self.op, self.filters = op, filters            # "and" / "or"
        self.precedence = 2 if op == "and" else 3


class EmptyFilter(Filter):
    precedence = 0


class FilterFactory:
    """NOT is pushed down to the leaves; composite children are sorted by precedence
    (filter.cc:36-108). sorted() is stable like libstdc++'s insertion sort for <=16 children."""

    @staticmethod
    def create(conf, negate=False):
        if not conf or "op" not in conf:
            return EmptyFilter()
        op = conf["op"]
        if op in ("and", "or"):
            filters = [FilterFactory.create(f, negate) for f in conf["filters"]]
            filters = sorted(filters, key=lambda f: f.precedence)
            if negate:
                op = "or" if op == "and" else "and"
            return CompositeFilter(op, filters)
        if op == "not":
            return FilterFactory.create(conf["filter"], not negate)
        column = conf["column"]
        if op == "in":
            return InFilter(column, [str(v) for v in conf["values"]], not negate)
        value = str(conf["value"])
        if op in RelOpFilter.OPS:
            return RelOpFilter(RelOpFilter.NEGATED[op] if negate else op, column, value)
        raise ValueError("Unsupported filter operataor: " + op)   # sic (filter.cc:100)


def filter_columns(f, out=None):
    """query::ColumnsCollector."""
    out = set() if out is None else out
    if isinstance(f, (RelOpFilter, InFilter)):
        out.add(f.column)
    elif isinstance(f, CompositeFilter):
        for c in f.filters:
            filter_columns(c, out)
    return out


# ------------------------------------------------------------------------------------------------
# literal -> AnyNum image (ValueDecoder, src/codegen/query/filter.cc:154-204)
# ------------------------------------------------------------------------------------------------
_PACK = {N.U8: "<B", N.U16: "<H", N.U32: "<I", N.U64: "<Q", N.I8: "<b", N.I16: "<h", N.I32: "<i",
         N.I64: "<q", N.F32: "<f", N.F64: "<d"}
_INT_BITS = {N.U8: 8, N.U16: 16, N.U32: 32, N.U64: 64, N.I8: 8, N.I16: 16, N.I32: 32, N.I64: 64}


def _c_stox(value, what):
    """std::stoul / stoi / stoll: leading whitespace, optional sign, digits; trailing junk ignored."""
    s = value.lstrip()
    i = 0
    if i < len(s) and s[i] in "+-":
        i += 1
    j = i
    while j < len(s) and s[j].isdigit():
        j += 1
    if j == i:
        raise ValueError(what)              # std::invalid_argument(what)
    return int(s[:j])


def parse_number(value, vtype):
    """NumericType::Parse / UIntType::Parse (column.cc:40-52,244-268): parse with the C++ function
    of the type, then C-cast to the column type (wraps)."""
    if vtype == N.F32:
        return struct.unpack("<f", struct.pack("<f", float(value)))[0]
    if vtype == N.F64:
        return float(value)
    what = {N.I8: "stoi", N.I16: "stoi", N.I32: "stoi", N.I64: "stoll", N.U64: "stoull"}.get(vtype, "stoul")
    v = _c_stox(value, what)
    if what == "stoi" and not (-2**31 <= v < 2**31):
        raise OverflowError("stoi")         # std::out_of_range
    if what == "stoll" and not (-2**63 <= v < 2**63):
        raise OverflowError("stoll")
    bits = _INT_BITS[vtype]
    v &= (1 << bits) - 1
    if vtype in (N.I8, N.I16, N.I32, N.I64) and v >= 1 << (bits - 1):
        v -= 1 << bits
    return v


def anynum_image(v, vtype):
    """8-byte db::AnyNum image (column.h:98-121): only the low sizeof(T) bytes are meaningful."""
    raw = struct.pack(_PACK[vtype], v)
    return struct.unpack("<Q", raw + b"\0" * (8 - len(raw)))[0]


def decode_value(column, value):
    """ValueDecoder::Visit(*) -> (python value, AnyNum image)."""
    if column.is_dimension:
        if column.kind == N.DIM_STRING:
            v = column.dict.decode(value)
        elif column.kind in (N.DIM_TIME, N.DIM_MICROTIME):
            v = parse_time_literal(value, column.kind == N.DIM_MICROTIME)
        elif column.kind == N.DIM_BOOLEAN:
            v = 1 if value == "true" else 0
        else:
            v = parse_number(value, column.type)
    else:
        v = parse_number(value, column.type)
    return v, anynum_image(v, column.type)


class FilterArgsPacker:
    """Walks the filter in the reference's leaf order and packs one AnyNum per literal
    (filter.cc:100-124). Also emits the post-order predicate program of include/vgpu.h."""

    def __init__(self, table):
        self.table = table
        self.args = []      # AnyNum images
        self.values = []    # decoded python values (host-side HAVING evaluation)
        self.nodes = []     # (kind, op, col, arg, n)

    def visit(self, f):
        t = self.table
        if isinstance(f, RelOpFilter):
            col = t.column(f.column)
            v, img = decode_value(col, f.value)
            self.nodes.append((N.NODE_RELOP, RelOpFilter.OPS.index(f.op), t.schema_index(col), len(self.args), 0))
            self.args.append(img)
            self.values.append(v)
        elif isinstance(f, InFilter):
            col = t.column(f.column)
            first = len(self.args)
            for s in f.values:
                v, img = decode_value(col, s)
                self.args.append(img)
                self.values.append(v)
            self.nodes.append((N.NODE_IN, 1 if f.equal else 0, t.schema_index(col), first, len(f.values)))
        elif isinstance(f, CompositeFilter):
            for c in f.filters:
                self.visit(c)
            self.nodes.append((N.NODE_AND if f.op == "and" else N.NODE_OR, 0, 0, 0, len(f.filters)))
        else:
            self.nodes.append((N.NODE_EMPTY, 0, 0, 0, 0))
        return self


# ------------------------------------------------------------------------------------------------
# queries (src/query/query.h:50-240, query.cc:25-173)
# ------------------------------------------------------------------------------------------------
class DimOutputColumn:
    def __init__(self, conf, dim, index):
        self.dim, self.index = dim, index
        self.format = ""
        self.granularity = None
        # the (dim, index) ctor used for the "dimensions": [...] form leaves format empty (query.h:121-122)
        if conf is not None and dim.kind in (N.DIM_TIME, N.DIM_MICROTIME):
            self.format = conf.get("format", dim.format)
            if "granularity" in conf:
                self.granularity = time_unit_by_name(conf["granularity"])


class MetricOutputColumn:
    def __init__(self, conf, metric, index):
        self.metric, self.index = metric, index


class SortColumn:
    def __init__(self, col, index, ascending):
        self.col, self.index, self.ascending = col, index, ascending


class SelectQuery:
    """query::SelectQuery (src/query/query.cc:48-85)."""

    def __init__(self, conf, table):
        self.table = table
        self.header = bool(conf.get("header", False))
        self.filter = FilterFactory.create(conf.get("filter"))
        self.skip = int(conf.get("skip", 0))
        self.limit = int(conf.get("limit", 0))
        self.dimension_cols, self.metric_cols = [], []
        idx = 0
        if "select" in conf:
            for sel in conf["select"]:
                name = sel["column"]
                cols = table.columns() if name == "*" else [table.column(name)]
                for col in cols:
                    if col.is_dimension:
                        self.dimension_cols.append(DimOutputColumn(sel, col, idx))
                    else:
                        self.metric_cols.append(MetricOutputColumn(sel, col, idx))
                    idx += 1
        else:
            for name in conf.get("dimensions", []):
                self.dimension_cols.append(DimOutputColumn(None, table.dimension(name), idx))
                idx += 1
            for name in conf.get("metrics", []):
                self.metric_cols.append(MetricOutputColumn(None, table.metric(name), idx))
                idx += 1
        self.ncols = idx

    def accept(self, visitor):
        visitor.visit_select(self)


class SearchQuery:
    """query::SearchQuery (src/query/query.cc:139-142)."""

    def __init__(self, conf, table):
        self.table = table
        self.header = bool(conf.get("header", False))
        self.filter = FilterFactory.create(conf.get("filter"))
        self.dimension = table.dimension(conf["dimension"])
        self.term = conf["term"]
        self.limit = int(conf.get("limit", 0))

    def accept(self, visitor):
        visitor.visit_search(self)


def replay_search(segment_lists, fmt_value, term, limit):
    """The sequential part of the reference's search (scan.cc:273-295) over the per-segment lists of
    (first row, code) the device produced: a code is marked seen when its first occurrence in a segment is
    reached; `limit` breaks the tuple loop only, so the rest of THAT segment stays unseen while the next
    segments are still visited. Returns (values, number of distinct codes seen)."""
    seen, values = set(), []
    for codes in segment_lists:           # scan order; codes ascending by first row inside a segment
        for code in codes:
            if code in seen:
                continue
            seen.add(code)
            v = fmt_value(code)
            if term in v:
                values.append(v)
                if limit > 0 and len(values) >= limit:
                    break
    return values, len(seen)


class AggregateQuery:
    def __init__(self, conf, table):
        self.table = table
        self.header = bool(conf.get("header", False))
        self.filter = FilterFactory.create(conf.get("filter"))
        self.skip = int(conf.get("skip", 0))
        self.limit = int(conf.get("limit", 0))
        self.dimension_cols, self.metric_cols = [], []
        idx = 0
        if "select" in conf:
            for sel in conf["select"]:
                name = sel["column"]
                cols = table.columns() if name == "*" else [table.column(name)]
                for col in cols:
                    if col.is_dimension:
                        self.dimension_cols.append(DimOutputColumn(sel, col, idx))
                    else:
                        self.metric_cols.append(MetricOutputColumn(sel, col, idx))
                    idx += 1
        else:
            for name in conf.get("dimensions", []):
                self.dimension_cols.append(DimOutputColumn(None, table.dimension(name), idx))
                idx += 1
            for name in conf.get("metrics", []):
                self.metric_cols.append(MetricOutputColumn(None, table.metric(name), idx))
                idx += 1
        self.ncols = idx
        self.sort_cols = []
        for sc in conf.get("sort", []) or []:
            col = table.column(sc["column"])
            col_idx = -1
            for dc in self.dimension_cols:
                if dc.dim is col:
                    col_idx = dc.index
                    break
            if col_idx == -1:
                for mc in self.metric_cols:
                    if mc.metric is col:
                        col_idx = mc.index
                        break
            if col_idx == -1:
                raise ValueError("Sort column '" + sc["column"] + "' is not selected")
            self.sort_cols.append(SortColumn(col, col_idx, bool(sc.get("ascending", False))))
        self.having = None
        if "having" in conf:
            self.having = FilterFactory.create(conf["having"])
            names = self.column_names()
            for c in filter_columns(self.having):
                if c not in names:
                    raise ValueError("Column '" + c + " is not selected")   # sic (query.cc:131)

    def column_names(self):
        return [dc.dim.name for dc in self.dimension_cols] + [mc.metric.name for mc in self.metric_cols]

    def accept(self, visitor):
        visitor.visit_aggregate(self)


class QueryFactory:
    @staticmethod
    def create(conf, database):
        qtype = conf["type"]
        if qtype == "aggregate":
            return AggregateQuery(conf, database.get_table(conf["table"]))
        if qtype == "select":
            return SelectQuery(conf, database.get_table(conf["table"]))
        if qtype == "search":
            return SearchQuery(conf, database.get_table(conf["table"]))
        if qtype == "show":
            raise NotImplementedError("'show' queries list tables: no device work, they stay on the stock QueryRunner")
        raise ValueError("unsupported query type: " + qtype)


class MemoryRowOutput:
    """query::MemoryRowOutput (src/query/output.h:37-48)."""

    def __init__(self):
        self.rows = []

    def start(self):
        pass

    def send(self, row):
        self.rows.append(list(row))

    def flush(self):
        pass


class QueryStats:
    """query::QueryStats (src/query/stats.h:35-58) + device-side counters of our own."""

    def __init__(self):
        self.scanned_segments = 0
        self.scanned_recs = 0
        self.aggregated_recs = 0
        self.output_recs = 0
        self.compile_time = 0.0     # always 0: the plan is data, nothing is compiled per query
        self.whole_time = 0.0
        self.scan_time = 0.0        # wall time of vgpu_query_agg (groups resident on the host)
        self.gpu_ms = 0.0
        self.kernel_scan_ms = 0.0
        self.passed_rows = 0
        self.launches = 0
        self.table_mode = 0
        self.table_cells = 0
        self.attempts = 0
        self.distinct_paths = 0
        self.post_applied = 0


# ------------------------------------------------------------------------------------------------
# formatting (src/util/format.h:28-82)
# ------------------------------------------------------------------------------------------------
def fmt_num(v, vtype):
    if vtype == N.F64:
        return "%.15g" % v
    if vtype == N.F32:
        return "%g" % v            # fmt 4.x default for a float promoted to double
    return str(int(v))


def fmt_date(fmt, ts):
    return _time.strftime(fmt, _time.gmtime(int(ts) & 0xFFFFFFFF))   # Format::date takes uint32_t


def _smaller_int(a, b):
    return len(a) < len(b) if len(a) != len(b) else a < b


def _cmp_rows(sort_cols):
    """The std::sort comparator sort.cc:32-66 generates (strings; INTEGER = (length, lexicographic))."""
    import functools

    def less(a, b):
        n = len(sort_cols)
        for i, sc in enumerate(sort_cols):
            x, y = a[sc.index], b[sc.index]
            st = sc.col.sort_type
            if st == "string":
                lt = (lambda p, q: p < q) if sc.ascending else (lambda p, q: p > q)
            elif st == "integer":
                lt = _smaller_int if sc.ascending else (lambda p, q: _smaller_int(q, p))
            else:
                lt = (lambda p, q: float(p) < float(q)) if sc.ascending else (lambda p, q: float(p) > float(q))
            if lt(x, y):
                return True
            if i < n - 1 and lt(y, x):
                return False
            if i == n - 1:
                return False
        return False

    return functools.cmp_to_key(lambda a, b: -1 if less(a, b) else (1 if less(b, a) else 0))


# ------------------------------------------------------------------------------------------------
# the runner
# ------------------------------------------------------------------------------------------------
class _ResultOwner:
    """Keeps a vgpu_result alive while numpy views of its (library-owned, pinned) arrays exist."""

    def __init__(self, lib, handle):
        self._lib, self._handle = lib, handle

    def __del__(self):
        if self._handle is not None:
            self._lib.vgpu_result_free(self._handle)
            self._handle = None


class GpuQueryRunner:
    """Counterpart of query::QueryRunner (src/query/runner.cc:45-64): visit_aggregate packs the
    filter arguments exactly like the reference, lowers the query to a vgpu_plan, calls
    vgpu_query_agg and runs the host post-aggregation into the caller's RowOutput."""

    def __init__(self, database, output, flags=0, now=None, device_post=True):
        self.database = database
        self.output = output
        self.stats = QueryStats()
        self.flags = flags
        self.now = now
        self.device_post = device_post   # HAVING / top-N on the device (vgpu.h VGPU_PLAN_POST); False: the host does both
        self.last_result = None

    # "now" for rollup boundaries: VIYA_TEST_ROLLUP_TS pins it (codegen/db/rollup.cc:47-49)
    def _now(self):
        if self.now is not None:
            return int(self.now)
        env = os.environ.get("VIYA_TEST_ROLLUP_TS")
        if env:
            return int(env.rstrip("Ll"))
        return int(_time.time())

    def build_plan(self, query):
        t = query.table
        packer = FilterArgsPacker(t).visit(query.filter)
        nodes = (N.PredNode * max(1, len(packer.nodes)))()
        for i, (kind, op, col, arg, n) in enumerate(packer.nodes):
            nodes[i] = N.PredNode(kind, op, col, arg, n, 0)
        args = (C.c_uint64 * max(1, len(packer.args)))(*packer.args)
        keys = (N.Key * max(1, len(query.dimension_cols)))()
        now = None
        for i, dc in enumerate(query.dimension_cols):
            k = N.Key()
            k.col = t.schema_index(dc.dim)
            k.nrules = 0
            k.query_granularity = N.TU_NONE
            if dc.dim.kind in (N.DIM_TIME, N.DIM_MICROTIME):
                rules = dc.dim.rollup_rules
                if len(rules) > N.MAX_ROLLUP_RULES:
                    raise N.VgpuError(N.ERR_UNSUPPORTED, "too many rollup rules")
                k.nrules = len(rules)
                for r, rule in enumerate(rules):
                    if now is None:
                        now = self._now()
                    # util::Duration(unit, count).add_to((uint32_t) now, -1) [* 1000000L]  (rollup.cc:59-69)
                    b = rule.after.add_to(now & 0xFFFFFFFF, -1)
                    if dc.dim.kind == N.DIM_MICROTIME:
                        b *= 1000000
                    k.rule_boundary[r] = b
                    k.rule_granularity[r] = rule.granularity
                if dc.granularity is not None:
                    k.query_granularity = dc.granularity
            keys[i] = k
        mcols = (C.c_uint32 * max(1, len(query.metric_cols)))(
            *[t.schema_index(mc.metric) for mc in query.metric_cols])
        has_avg = any(mc.metric.agg == N.AGG_AVG for mc in query.metric_cols)
        has_count = any(mc.metric.agg == N.AGG_COUNT for mc in query.metric_cols)
        plan = N.Plan()
        plan.nnodes, plan.nargs = len(packer.nodes), len(packer.args)
        plan.nodes, plan.args = nodes, args
        plan.nkeys, plan.nmetrics = len(query.dimension_cols), len(query.metric_cols)
        plan.keys, plan.metric_cols = keys, mcols
        plan.need_hidden_count = 1 if (has_avg and not has_count) else 0   # scan.cc:239-241
        plan.flags = self.flags
        plan._keep = (nodes, args, keys, mcols)
        # ---- post-aggregation on the device (vgpu.h: HAVING, top-N); both only remove groups, post_aggregate below
        # stays what it was. HAVING only where the reference tests every group: with a sort, or without skip / limit
        # (without a sort it cuts the skip / limit window out of the map iteration first, post_agg.cc:26-83).
        plan.sort_col, plan.sort_descending, plan.top_k = N.NO_COLUMN, 0, 0
        if self.device_post:
            plan.flags |= N.PLAN_POST
            if query.having is not None and (query.sort_cols or (query.skip == 0 and query.limit == 0)):
                hp = FilterArgsPacker(t).visit(query.having)
                hnodes = (N.PredNode * max(1, len(hp.nodes)))()
                for i, (kind, op, col, arg, n) in enumerate(hp.nodes):
                    hnodes[i] = N.PredNode(kind, op, col, arg, n, 0)
                hargs = (C.c_uint64 * max(1, len(hp.args)))(*hp.args)
                plan.nhnodes, plan.nhargs, plan.hnodes, plan.hargs = len(hp.nodes), len(hp.args), hnodes, hargs
                plan._keep += (hnodes, hargs)
            if query.sort_cols and query.limit > 0:
                sc = query.sort_cols[0]
                plan.sort_col = t.schema_index(sc.col)
                plan.sort_descending = 0 if sc.ascending else 1
                plan.top_k = query.skip + query.limit
        return plan

    def run_plan(self, query, plan):
        """vgpu_query_agg -> dict of numpy arrays (copied out of the library-owned result)."""
        lib = N.load()
        t = query.table
        res = C.c_void_p()
        t0 = _time.perf_counter()
        N.check(lib.vgpu_query_agg(t.handle, C.byref(plan), C.byref(res)))
        self.stats.scan_time = _time.perf_counter() - t0
        owner = _ResultOwner(lib, res)   # the arrays below are zero-copy views of library-owned pinned memory
        view = N.ResultView()
        N.check(lib.vgpu_result_get(res, C.byref(view)))
        n = view.ngroups
        keys, accs = [], []

        def wrap(addr, dt):
            if not n:
                return np.zeros(0, dt)
            buf = (C.c_char * (n * dt.itemsize)).from_address(addr)
            arr = np.frombuffer(buf, dtype=dt, count=n)
            arr.flags.writeable = False
            return arr

        for i, dc in enumerate(query.dimension_cols):
            keys.append(wrap(view.keys[i], np.dtype(N.NP_DTYPES[dc.dim.type])))
        for i, mc in enumerate(query.metric_cols):
            dt = np.dtype("<u8") if mc.metric.agg == N.AGG_BITSET else np.dtype(N.NP_DTYPES[mc.metric.type])
            accs.append(wrap(view.accs[i], dt))
        hidden = None
        if plan.need_hidden_count:
            hidden = wrap(C.cast(view.hidden_count, C.c_void_p).value, np.dtype("<u8"))
        s = self.stats
        s.scanned_recs, s.scanned_segments = view.scanned_recs, view.scanned_segments
        s.aggregated_recs, s.passed_rows = view.aggregated_recs, view.passed_rows
        s.gpu_ms, s.kernel_scan_ms, s.launches = view.gpu_ms, view.scan_ms, view.launches
        s.table_mode, s.table_cells = view.table_mode, view.table_cells
        s.attempts, s.distinct_paths = view.attempts, view.distinct_paths
        s.post_applied = view.post_applied
        return {"ngroups": n, "keys": keys, "accs": accs, "hidden_count": hidden, "_owner": owner}

    def _predicate(self, query):
        packer = FilterArgsPacker(query.table).visit(query.filter)
        nodes = (N.PredNode * max(1, len(packer.nodes)))()
        for i, (kind, op, col, arg, n) in enumerate(packer.nodes):
            nodes[i] = N.PredNode(kind, op, col, arg, n, 0)
        args = (C.c_uint64 * max(1, len(packer.args)))(*packer.args)
        return packer, nodes, args

    # ---- select (runner.cc:29-43, codegen/query/scan.cc:75-166): device picks the rows, host formats ----
    def visit_select(self, query):
        t_begin = _time.perf_counter()
        lib = N.load()
        t = query.table
        packer, nodes, args = self._predicate(query)
        cols = [dc.dim for dc in query.dimension_cols] + [mc.metric for mc in query.metric_cols]
        sel = [t.schema_index(c) for c in cols]
        # AVG cells are divided by the first selected COUNT metric, else by the table's hidden count (scan.cc:133-154)
        count_pos = next((i for i, c in enumerate(cols) if not c.is_dimension and c.agg == N.AGG_COUNT), None)
        if count_pos is None and any((not c.is_dimension) and c.agg == N.AGG_AVG for c in cols):
            hidden = t.hidden_count_index
            if hidden is None:
                raise N.VgpuError(N.ERR_INVALID, "AVG metric selected but the table has neither COUNT nor hidden count")
            count_pos = len(sel)
            sel.append(hidden)
        plan = N.RowsPlan()
        plan.nnodes, plan.nargs, plan.nodes, plan.args = len(packer.nodes), len(packer.args), nodes, args
        carr = (C.c_uint32 * max(1, len(sel)))(*sel)
        plan.ncols, plan.cols, plan.skip, plan.limit = len(sel), carr, query.skip, query.limit
        res = C.c_void_p()
        N.check(lib.vgpu_query_select(t.handle, C.byref(plan), C.byref(res)))
        try:
            view = N.RowsView()
            N.check(lib.vgpu_rows_get(res, C.byref(view)))
            n = view.nrows
            arrays = []
            for i, si in enumerate(sel):
                col = cols[i] if i < len(cols) else None
                if col is not None and not col.is_dimension and col.agg == N.AGG_BITSET:
                    dt = np.dtype("<u8")
                elif col is None:
                    dt = np.dtype("<u8")            # hidden count
                else:
                    dt = np.dtype(N.NP_DTYPES[col.type])
                if n:
                    buf = (C.c_char * (n * dt.itemsize)).from_address(view.cells[i])
                    arrays.append(np.frombuffer(buf, dtype=dt, count=n).copy())
                else:
                    arrays.append(np.zeros(0, dt))
            s = self.stats
            s.scanned_recs, s.scanned_segments, s.passed_rows = view.scanned_recs, view.scanned_segments, view.passed_rows
            s.gpu_ms, s.launches = view.gpu_ms, view.launches
        finally:
            lib.vgpu_rows_free(res)
        out = self.output
        out.start()
        if query.header:
            row = [None] * query.ncols
            for dc in query.dimension_cols:
                row[dc.index] = dc.dim.name
            for mc in query.metric_cols:
                row[mc.index] = mc.metric.name
            out.send(row)
        ndim = len(query.dimension_cols)
        for r in range(n):
            row = [None] * query.ncols
            for i, dc in enumerate(query.dimension_cols):
                v = arrays[i][r]
                d = dc.dim
                if d.kind == N.DIM_STRING:
                    row[dc.index] = d.dict.c2v[int(v)]
                elif d.kind in (N.DIM_TIME, N.DIM_MICROTIME) and dc.format:
                    row[dc.index] = fmt_date(dc.format, int(v))
                elif d.kind == N.DIM_BOOLEAN:
                    row[dc.index] = "true" if v else "false"
                else:
                    row[dc.index] = fmt_num(v, d.type)
            for i, mc in enumerate(query.metric_cols):
                v = arrays[ndim + i][r]
                m = mc.metric
                if m.agg == N.AGG_AVG:
                    with np.errstate(divide="ignore", invalid="ignore"):
                        row[mc.index] = "%.15g" % (np.float64(v) / np.float64(arrays[count_pos][r]))
                elif m.agg == N.AGG_BITSET:
                    row[mc.index] = str(int(v))
                else:
                    row[mc.index] = fmt_num(v, m.type)
            out.send(row)
        self.stats.output_recs = n
        out.flush()
        self.stats.whole_time = _time.perf_counter() - t_begin

    # ---- search (runner.cc:66-80, scan.cc:249-299, post_agg.cc:149-166) ----
    def visit_search(self, query):
        t_begin = _time.perf_counter()
        lib = N.load()
        t = query.table
        d = query.dimension
        packer, nodes, args = self._predicate(query)
        plan = N.SearchPlan()
        plan.nnodes, plan.nargs, plan.nodes, plan.args = len(packer.nodes), len(packer.args), nodes, args
        plan.col = t.schema_index(d)
        res = C.c_void_p()
        N.check(lib.vgpu_query_search(t.handle, C.byref(plan), C.byref(res)))
        try:
            view = N.SearchView()
            N.check(lib.vgpu_search_get(res, C.byref(view)))
            lists = []
            for si in range(view.nsegments):
                lo, hi = view.seg_offsets[si], view.seg_offsets[si + 1]
                lists.append([view.codes[i] for i in range(lo, hi)])
            s = self.stats
            s.scanned_recs, s.scanned_segments = view.scanned_recs, view.scanned_segments
            s.gpu_ms, s.launches = view.gpu_ms, view.launches
        finally:
            lib.vgpu_search_free(res)
        dt = np.dtype(N.NP_DTYPES[d.type])

        def fmt_value(code):
            if d.kind == N.DIM_STRING:
                return d.dict.c2v[int(code)]
            if d.kind == N.DIM_BOOLEAN:
                return "true" if code else "false"
            v = int(code)
            if dt.kind == "i" and v >= 1 << 63:
                v -= 1 << 64                      # codes arrive sign-extended to 64 bits
            return fmt_num(v, d.type)
        values, ncodes = replay_search(lists, fmt_value, query.term, query.limit)
        out = self.output
        out.start()
        if query.header:
            out.send([d.name])
        out.send_as_col(values) if hasattr(out, "send_as_col") else out.send(values)
        self.stats.aggregated_recs = ncodes
        self.stats.output_recs = len(values)
        out.flush()
        self.stats.whole_time = _time.perf_counter() - t_begin

    def visit_aggregate(self, query):
        t_begin = _time.perf_counter()
        plan = self.build_plan(query)
        hargs = FilterArgsPacker(query.table).visit(query.having).values if query.having is not None else []
        groups = self.run_plan(query, plan)
        self.last_result = groups
        self.post_aggregate(query, groups, hargs)
        self.stats.whole_time = _time.perf_counter() - t_begin

    # ---- post_agg.cc:26-147 + sort.cc:24-73, on the host ----
    def post_aggregate(self, query, groups, hargs):
        out, stats = self.output, self.stats
        out.start()
        n = groups["ngroups"]
        # skip / limit are clamped by agg_map.size() (post_agg.cc:40-47) = all groups, also when HAVING / top-N already ran
        # on the device and `groups` only holds the survivors
        n_all = max(n, stats.aggregated_recs)
        skip = min(n_all, query.skip)
        limit = min(query.limit, n_all - skip)
        sorting = bool(query.sort_cols)
        lo, hi = 0, n
        if not sorting:
            lo = skip
            if limit > 0:
                hi = lo + limit
        if query.header:
            row = [None] * query.ncols
            for dc in query.dimension_cols:
                row[dc.index] = dc.dim.name
            for mc in query.metric_cols:
                row[mc.index] = mc.metric.name
            out.send(row)

        keys, accs = groups["keys"], groups["accs"]
        keep = np.ones(hi - lo, dtype=bool)
        if query.having is not None and hi > lo:
            cursor = [0]
            keep = self._having(query, query.having, groups, lo, hi, hargs, cursor)
        idx = np.nonzero(keep)[0] + lo

        # count column for AVG: first selected COUNT metric, else the hidden one (post_agg.cc:104-111)
        count_arr = groups["hidden_count"]
        for i, mc in enumerate(query.metric_cols):
            if mc.metric.agg == N.AGG_COUNT:
                count_arr = accs[i]
                break

        cols = [None] * query.ncols
        for i, dc in enumerate(query.dimension_cols):
            vals = keys[i][idx]
            d = dc.dim
            if d.kind == N.DIM_STRING:
                c2v = d.dict.c2v
                cols[dc.index] = [c2v[int(v)] for v in vals]
            elif d.kind in (N.DIM_TIME, N.DIM_MICROTIME) and dc.format:
                cols[dc.index] = [fmt_date(dc.format, v) for v in vals]
            elif d.kind == N.DIM_BOOLEAN:
                cols[dc.index] = ["true" if v else "false" for v in vals]
            elif d.type in (N.F32, N.F64):
                cols[dc.index] = [fmt_num(v, d.type) for v in vals]
            else:
                cols[dc.index] = [str(v) for v in vals.tolist()]
        for i, mc in enumerate(query.metric_cols):
            m = mc.metric
            vals = accs[i][idx]
            if m.agg == N.AGG_AVG:
                # value / (double) count, printed as a double (post_agg.cc:126-127)
                cnt = count_arr[idx].astype(np.float64)
                with np.errstate(divide="ignore", invalid="ignore"):
                    q = vals.astype(np.float64) / cnt
                cols[mc.index] = ["%.15g" % v for v in q]
            elif m.agg == N.AGG_BITSET:
                cols[mc.index] = [str(v) for v in vals.tolist()]
            elif m.type in (N.F32, N.F64):
                cols[mc.index] = [fmt_num(v, m.type) for v in vals]
            else:
                cols[mc.index] = [str(v) for v in vals.tolist()]
        rows = [list(r) for r in zip(*cols)] if query.ncols and len(idx) else ([[] for _ in idx] if not query.ncols else [])

        if not sorting:
            for r in rows:
                out.send(r)
            stats.output_recs += len(rows)
        else:
            rows.sort(key=_cmp_rows(query.sort_cols))
            end = min(len(rows), skip + limit) if limit > 0 else len(rows)
            for r in rows[min(skip, len(rows)):end]:
                out.send(r)
                stats.output_recs += 1
        out.flush()

    def _having(self, query, f, groups, lo, hi, hargs, cursor):
        """FilterComparison over (agg key, accumulators) (post_agg.cc:76-83): AVG compares the raw sum,
        BITSET its cardinality (filter.cc:212-218)."""
        n = hi - lo

        def column_values(name):
            for i, dc in enumerate(query.dimension_cols):
                if dc.dim.name == name:
                    return groups["keys"][i][lo:hi]
            for i, mc in enumerate(query.metric_cols):
                if mc.metric.name == name:
                    return groups["accs"][i][lo:hi]
            raise ValueError("Column '" + name + " is not selected")

        def typed(col_vals, v):
            # both sides have the column's C++ type after the reference's arg unpack (filter.cc:126-132)
            return np.array(v).astype(col_vals.dtype)

        if isinstance(f, RelOpFilter):
            vals = column_values(f.column)
            a = typed(vals, hargs[cursor[0]])
            cursor[0] += 1
            return {"eq": vals == a, "ne": vals != a, "lt": vals < a, "le": vals <= a,
                    "gt": vals > a, "ge": vals >= a}[f.op]
        if isinstance(f, InFilter):
            vals = column_values(f.column)
            r = np.zeros(n, dtype=bool) if f.equal else np.ones(n, dtype=bool)
            for _ in f.values:
                a = typed(vals, hargs[cursor[0]])
                cursor[0] += 1
                r = (r | (vals == a)) if f.equal else (r & (vals != a))
            return r
        if isinstance(f, CompositeFilter):
            r = None
            for c in f.filters:
                x = self._having(query, c, groups, lo, hi, hargs, cursor)
                r = x if r is None else ((r & x) if f.op == "and" else (r | x))
            return r
        return np.ones(n, dtype=bool)
