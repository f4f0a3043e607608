"""Host-side mirror of the reference's schema objects (db::Table, db::Column hierarchy,
db::DimensionDict) plus the HBM-resident column store behind the C ABI.

Reference: src/db/table.cc:47-96 (table config), src/db/column.cc:54-62,275-402 (type rules),
src/db/dictionary.cc:22-75 (code 0 == "__exceeded", missing value decodes to UINTn_MAX),
src/db/store.h:32-58 + src/codegen/db/store.cc:203-356 (segments of `segment_size` rows, SoA).

Only what the scan -> filter -> group-by path reads is modelled; ingest (upsert), watchers and
partitioning stay in the reference.
"""
import ctypes as C
import json
import struct

import numpy as np

from . import _native as N
from .timeutil import Duration, time_unit_by_name

_NUM_TYPES = {"byte": N.I8, "ubyte": N.U8, "short": N.I16, "ushort": N.U16, "int": N.I32, "uint": N.U32,
              "long": N.I64, "ulong": N.U64, "float": N.F32, "double": N.F64}
_AGGS = {"max": N.AGG_MAX, "min": N.AGG_MIN, "sum": N.AGG_SUM, "avg": N.AGG_AVG}
UINT_MAX = {N.U8: 0xFF, N.U16: 0xFFFF, N.U32: 0xFFFFFFFF, N.U64: 0xFFFFFFFFFFFFFFFF}


def max_value_to_uint_type(max_value):
    """src/db/column.cc:54-62 (size_t arithmetic: 0 - 1 wraps)."""
    m = (int(max_value) - 1) & 0xFFFFFFFFFFFFFFFF
    if m < 0xFF:
        return N.U8
    if m < 0xFFFF:
        return N.U16
    if m < 0xFFFFFFFF:
        return N.U32
    return N.U64


class DimensionDict:
    """db::DimensionDict: value <-> code; code 0 is the real value "__exceeded"."""

    def __init__(self, code_type, c2v=None):
        self.code_type = code_type
        self.c2v = list(c2v) if c2v else ["__exceeded"]
        self.v2c = {v: i for i, v in enumerate(self.c2v)}

    def decode(self, value):
        """DimensionDict::Decode (dictionary.cc:46-75): missing -> UINTn_MAX, so eq matches nothing."""
        return self.v2c.get(value, UINT_MAX[self.code_type])

    def encode(self, value):
        """What the upsert path does for a new value (codegen/db/upsert.cc:43-65), without the
        cardinality overflow branch: used by tests/bench to build tables."""
        code = self.v2c.get(value)
        if code is None:
            code = len(self.c2v)
            self.c2v.append(value)
            self.v2c[value] = code
        return code


class RollupRule:
    def __init__(self, conf):
        self.granularity = time_unit_by_name(conf["granularity"])
        self.after = Duration(conf["after"])


class Column:
    """db::Column; `sort_type` as in column.h:175,198,216,236,270,286."""
    is_dimension = False

    def __init__(self, name, index):
        self.name = name
        self.index = index


class Dimension(Column):
    is_dimension = True

    def __init__(self, conf, index):
        super().__init__(conf["name"], index)
        dtype = conf.get("type", "string")
        self.dim_type = dtype
        self.rollup_rules = []
        self.granularity = None
        self.format = conf.get("format", "")
        self.dict = None
        if dtype == "string":
            self.kind = N.DIM_STRING
            self.cardinality = int(conf.get("cardinality", 0xFFFFFFFF))
            self.type = max_value_to_uint_type(self.cardinality)
            self.sort_type = "string"
        elif dtype == "boolean":
            self.kind, self.type, self.sort_type = N.DIM_BOOLEAN, N.U8, "integer"
        elif dtype in ("time", "microtime"):
            self.micro = dtype == "microtime"
            self.kind = N.DIM_MICROTIME if self.micro else N.DIM_TIME
            self.type = N.U64 if self.micro else N.U32
            self.sort_type = "string"
            if "granularity" in conf:
                self.granularity = time_unit_by_name(conf["granularity"])
            elif "rollup_rules" in conf:
                rules = [RollupRule(r) for r in conf["rollup_rules"]]
                # std::sort by descending `after` (column.cc:346-349)
                self.rollup_rules = sorted(rules, key=lambda r: -r.after.seconds_from_epoch())
        else:
            self.kind = N.DIM_NUMERIC
            if dtype == "numeric":  # deprecated alias (column.cc:314-322)
                self.type = N.U64 if max_value_to_uint_type(conf.get("max", 0xFFFFFFFF)) == N.U64 else N.U32
            elif dtype in _NUM_TYPES:
                self.type = _NUM_TYPES[dtype]
            else:
                raise ValueError("Unsupported numeric type: " + dtype)
            self.sort_type = "float" if self.type in (N.F32, N.F64) else "integer"
        self.agg = N.AGG_NONE


class Metric(Column):
    def __init__(self, conf, index):
        super().__init__(conf["name"], index)
        mtype = conf["type"]
        if mtype == "bitset":
            self.kind, self.agg = N.METRIC_BITSET, N.AGG_BITSET
            self.type = max_value_to_uint_type(conf.get("max", 0xFFFFFFFF))
            self.sort_type = "integer"
        else:
            self.kind = N.METRIC_VALUE
            if mtype == "count":
                self.agg = N.AGG_COUNT
                self.type = N.U64 if max_value_to_uint_type(conf.get("max", 0xFFFFFFFF)) == N.U64 else N.U32
            else:
                base, _, agg = mtype.partition("_")
                if agg not in _AGGS:
                    raise ValueError("Unsupported metric type: " + mtype)
                if base not in _NUM_TYPES:
                    raise ValueError("Unsupported numeric type: " + base)
                self.agg, self.type = _AGGS[agg], _NUM_TYPES[base]
            self.sort_type = "float" if self.type in (N.F32, N.F64) else "integer"


class Table:
    """db::Table + its SegmentStore, with the segments resident in HBM (vgpu_table)."""

    def __init__(self, conf, database):
        self.database = database
        self.name = conf["name"]
        self.segment_size = int(conf.get("segment_size", 1000000))
        self.dimensions = [Dimension(c, i) for i, c in enumerate(conf.get("dimensions", []))]
        self.metrics = [Metric(c, i) for i, c in enumerate(conf.get("metrics", []))]
        for d in self.dimensions:
            if d.kind == N.DIM_STRING:
                d.dict = database.dicts.setdefault(d.name, DimensionDict(d.type))
        has_avg = any(m.agg == N.AGG_AVG for m in self.metrics)
        has_count = any(m.agg == N.AGG_COUNT for m in self.metrics)
        self.has_hidden_count = has_avg and not has_count  # store.cc:286-289
        # schema column order of the C ABI: dims, metrics, hidden count
        self.ncols = len(self.dimensions) + len(self.metrics) + (1 if self.has_hidden_count else 0)
        self._handle = None
        if database.ctx is not None:
            self._create_device_table()

    # ---- lookups (table.cc:110-150) ----
    def column(self, name):
        for c in self.dimensions + self.metrics:
            if c.name == name:
                return c
        raise ValueError("No such column: " + name)

    def dimension(self, name):
        for d in self.dimensions:
            if d.name == name:
                return d
        raise ValueError("No such dimension: " + name)

    def metric(self, name):
        for m in self.metrics:
            if m.name == name:
                return m
        raise ValueError("No such metric: " + name)

    def columns(self):
        return self.dimensions + self.metrics

    def column_names(self):
        return [c.name for c in self.columns()]

    def schema_index(self, column):
        return column.index if column.is_dimension else len(self.dimensions) + column.index

    @property
    def hidden_count_index(self):
        return len(self.dimensions) + len(self.metrics) if self.has_hidden_count else None

    # ---- device store ----
    def _create_device_table(self):
        lib = N.load()
        cols = (N.Column * self.ncols)()
        i = 0
        for c in self.dimensions + self.metrics:
            btype, lit = c.type, 0
            if c.kind == N.METRIC_BITSET and c.type in (N.U8, N.U16):
                btype, lit = N.U32, c.type + 1  # ids travel as uint32, filter literals keep their own width (vgpu.h)
            cols[i] = N.Column(c.kind, btype, c.agg, lit)
            i += 1
        if self.has_hidden_count:
            cols[i] = N.Column(N.METRIC_HIDDEN_COUNT, N.U64, N.AGG_COUNT, 0)
        schema = N.Schema(self.ncols, len(self.dimensions), self.segment_size, cols)
        h = C.c_void_p()
        N.check(lib.vgpu_table_create(self.database.ctx, C.byref(schema), C.byref(h)))
        self._handle = h

    @property
    def handle(self):
        if self._handle is None:
            raise RuntimeError("table has no device store (freed, or the database was opened with device=None "
                               "for plan-only use); there is no CPU scan path")
        return self._handle

    def sync(self):
        """Wait for the copies of earlier put_segment(..., wait=False) calls (vgpu_table_sync)."""
        N.check(N.load().vgpu_table_sync(self.handle))
        self._async_keep = []

    def put_segment(self, seg_idx, columns, hidden_count=None, wait=True):
        self.put_prepared(seg_idx, self.prepare_segment(columns, hidden_count), wait)

    def put_prepared(self, seg_idx, prepared, wait=True):
        """vgpu_segment_put / _put_async of host columns marshalled once by prepare_segment (a caller that uploads the same
        host buffers again and again — bench.py's end-to-end leg — pays the ctypes marshalling once)."""
        lib = N.load()
        ptrs, nrows, keep = prepared
        if wait:
            N.check(lib.vgpu_segment_put(self.handle, seg_idx, nrows, ptrs))
        else:   # the host arrays must outlive the copies: kept until sync()
            N.check(lib.vgpu_segment_put_async(self.handle, seg_idx, nrows, ptrs))
            if not hasattr(self, "_async_keep"):
                self._async_keep = []
            self._async_keep.append(keep)

    def update_segment(self, seg_idx, row_begin, columns, hidden_count=None):
        """vgpu_segment_update: replace / append rows [row_begin, row_begin + n) of an uploaded segment; `columns` hold
        the n rows of the range only."""
        ptrs, nrows, keep = self.prepare_segment(columns, hidden_count)
        N.check(N.load().vgpu_segment_update(self.handle, seg_idx, row_begin, nrows, ptrs))

    def prepare_segment(self, columns, hidden_count=None):
        """Copy one segment into HBM. `columns`: name -> numpy array (dict codes / numbers), or for a
        BITSET metric either a 1-D uint32 array (one id per row) or a pair (offsets uint64[n+1],
        values uint32[nvalues])."""
        lib = N.load()
        ptrs = (C.c_void_p * self.ncols)()
        keep = []
        nrows = None
        for c in self.dimensions + self.metrics:
            v = columns[c.name]
            si = self.schema_index(c)
            if c.kind == N.METRIC_BITSET:
                idt = "<u8" if c.type == N.U64 else "<u4"   # util::Bitset<8> (Roaring64Map) ids travel as uint64
                if isinstance(v, tuple):
                    offsets = np.ascontiguousarray(v[0], dtype="<u8")
                    values = np.ascontiguousarray(v[1], dtype=idt)
                    n = len(offsets) - 1
                    csr = N.BitsetCsr(offsets.ctypes.data_as(C.POINTER(C.c_uint64)), values.ctypes.data, len(values))
                    keep += [offsets, values]
                else:
                    values = np.ascontiguousarray(v, dtype=idt)
                    n = len(values)
                    csr = N.BitsetCsr(None, values.ctypes.data, len(values))
                    keep.append(values)
                keep.append(csr)
                ptrs[si] = C.cast(C.pointer(csr), C.c_void_p)
            else:
                arr = np.ascontiguousarray(v, dtype=N.NP_DTYPES[c.type])
                n = len(arr)
                keep.append(arr)
                ptrs[si] = arr.ctypes.data
            if nrows is None:
                nrows = n
            elif nrows != n:
                raise ValueError(f"column {c.name} has {n} rows, expected {nrows}")
        if self.has_hidden_count:
            if hidden_count is None:
                raise ValueError("table has AVG without COUNT: the hidden count column is required")
            arr = np.ascontiguousarray(hidden_count, dtype="<u8")
            if len(arr) != nrows:
                raise ValueError("hidden count length mismatch")
            keep.append(arr)
            ptrs[self.ncols - 1] = arr.ctypes.data
        return ptrs, nrows or 0, keep

    def generate_segment(self, seg_idx, nrows, gens, seed=42, row_offset=0):
        """Synthetic segment written directly in HBM (vgpu_segment_generate). `gens`: one
        (lo, range[, mode, div]) per schema column."""
        lib = N.load()
        arr = (N.GenCol * self.ncols)()
        for i, g in enumerate(gens):
            lo, rng = g[0], g[1]
            mode = g[2] if len(g) > 2 else 0
            div = g[3] if len(g) > 3 else 1
            arr[i] = N.GenCol(lo, rng, mode, 0, div)
        N.check(lib.vgpu_segment_generate(self.handle, seg_idx, nrows, arr, seed, row_offset))

    def read_column(self, seg_idx, column, nrows):
        lib = N.load()
        c = self.column(column) if isinstance(column, str) else column
        dtype = "<u4" if c.kind == N.METRIC_BITSET else N.NP_DTYPES[c.type]
        out = np.empty(nrows, dtype=dtype)
        N.check(lib.vgpu_segment_read(self.handle, seg_idx, self.schema_index(c), out.ctypes.data))
        return out

    def invalidate(self, seg_idx):
        N.check(N.load().vgpu_table_invalidate(self.handle, seg_idx))

    @property
    def segments(self):
        return N.load().vgpu_table_segments(self.handle)

    @property
    def rows(self):
        return N.load().vgpu_table_rows(self.handle)

    @property
    def device_bytes(self):
        return N.load().vgpu_table_bytes(self.handle)

    def free(self):
        if self._handle is not None:
            N.load().vgpu_table_free(self._handle)
            self._handle = None

    # ---- loading the reference's own segments (oracle_cli "dump": VGPUSEG1 container) ----
    def load_dump(self, path, shard=None, chunk_rows=None):
        """Upload segments dumped from the reference's SegmentStore so the CUDA path scans the very
        bytes the reference scanned. Returns the dump header. With chunk_rows the reference's segments are cut
        into pieces of that many rows; with shard=(rank, world) only every world-th piece is uploaded (one process
        per GPU: the pieces of all ranks together are the reference's table)."""
        hdr, blob = read_dump(path)
        for d in self.dimensions:
            if d.kind == N.DIM_STRING:
                c2v = hdr["dicts"][d.name]
                d.dict.c2v = list(c2v)
                d.dict.v2c = {v: i for i, v in enumerate(c2v)}
        ncol_file = len(hdr["dims"]) + len(hdr["metrics"])
        assert ncol_file == len(self.dimensions) + len(self.metrics), "dump/schema column count mismatch"
        piece, local = 0, 0
        for seg in hdr["segments"]:
            size = seg["size"]
            cols = {}
            for c, cj in zip(self.dimensions + self.metrics, seg["cols"]):
                if c.kind == N.METRIC_BITSET:
                    offsets = np.frombuffer(blob, dtype="<u8", count=size + 1, offset=cj["off"])
                    values = np.frombuffer(blob, dtype="<u8", count=cj["values"], offset=cj["values_off"])
                    cols[c.name] = (offsets, values)
                else:
                    cols[c.name] = np.frombuffer(blob, dtype=N.NP_DTYPES[c.type], count=size, offset=cj["off"])
            hidden = None
            if self.has_hidden_count:
                hidden = np.frombuffer(blob, dtype="<u8", count=size, offset=seg["hidden_count"]["off"])
            step = chunk_rows or max(size, 1)
            for lo in range(0, max(size, 1), step):
                hi = min(size, lo + step)
                mine = shard is None or piece % shard[1] == shard[0]
                piece += 1
                if not mine:
                    continue
                part = {}
                for name, v in cols.items():
                    if isinstance(v, tuple):
                        o = v[0][lo:hi + 1]
                        part[name] = (o - o[0], v[1][int(o[0]):int(o[-1])]) if len(o) else (np.zeros(1, "<u8"), v[1][:0])
                    else:
                        part[name] = v[lo:hi]
                self.put_segment(local, part, None if hidden is None else hidden[lo:hi])
                local += 1
        return hdr


def read_dump(path):
    with open(path, "rb") as f:
        data = f.read()
    if data[:8] != b"VGPUSEG1":
        raise ValueError("not a VGPUSEG1 container: " + path)
    (hl,) = struct.unpack("<Q", data[8:16])
    hdr = json.loads(data[16:16 + hl].decode("utf-8"))
    start = 16 + hl
    start += (8 - start % 8) % 8
    return hdr, memoryview(data)[start:]


class Database:
    """db::Database reduced to what the query path needs: tables, dictionaries, the device context.

    Database.query() is the counterpart of src/db/database.cc:104-111 with GpuQueryRunner in place
    of QueryRunner.
    """

    def __init__(self, conf=None, device=0):
        self.dicts = {}
        self.tables = {}
        self.ctx = None
        self.device = device
        if device is not None:  # device=None: schema / plan objects only (host-logic tests); no scan possible
            h = C.c_void_p()
            N.check(N.load().vgpu_init(device, C.byref(h)))
            self.ctx = h
        self.rank, self.nranks = 0, 1
        for tconf in (conf or {}).get("tables", []):
            self.create_table(tconf)

    def create_table(self, conf):
        t = Table(conf, self)
        self.tables[t.name] = t
        return t

    def get_table(self, name):
        if name not in self.tables:
            raise ValueError("No such table: " + name)
        return self.tables[name]

    def query(self, conf, output=None, **kw):
        from .query import GpuQueryRunner, QueryFactory, MemoryRowOutput
        output = output if output is not None else MemoryRowOutput()
        runner = GpuQueryRunner(self, output, **kw)
        q = QueryFactory.create(conf, self)
        q.accept(runner)
        return runner.stats

    # ---- multi-GPU: one process per GPU ----
    def set_test_hook(self, name, value):
        """vgpu_set_test_hook: force the rarely taken branches of the device path at test sizes."""
        N.check(N.load().vgpu_set_test_hook(self.ctx, name.encode(), int(value)))

    def init_comm(self, rank, nranks, unique_id):
        N.check(N.load().vgpu_comm_init(self.ctx, rank, nranks, unique_id))
        self.rank, self.nranks = rank, nranks

    @staticmethod
    def comm_unique_id():
        buf = C.create_string_buffer(128)
        N.check(N.load().vgpu_comm_unique_id(buf))
        return buf.raw

    def close(self):
        for t in self.tables.values():
            t.free()
        self.tables = {}
        if self.ctx is not None:
            N.load().vgpu_shutdown(self.ctx)
            self.ctx = None
