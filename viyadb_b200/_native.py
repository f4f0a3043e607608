"""ctypes binding of libvgpu.so (include/vgpu.h) — the C-ABI drop-in boundary.

There is no fallback: if the shared library is missing or the CUDA device is absent, every call
raises. Nothing in this package computes a query result on the CPU.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VGPU_LIB_PATH") or os.path.join(_HERE, "libvgpu.so")   # override: A/B of kernel variants

VGPU_ABI_VERSION = 3

# vgpu_status
OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_NOMEM, ERR_NCCL, ERR_STATE = 0, -1, -2, -3, -4, -5, -6
# vgpu_type
U8, U16, U32, U64, I8, I16, I32, I64, F32, F64 = range(10)
TYPE_NAMES = ["u8", "u16", "u32", "u64", "i8", "i16", "i32", "i64", "f32", "f64"]
TYPE_WIDTH = [1, 2, 4, 8, 1, 2, 4, 8, 4, 8]
NP_DTYPES = ["<u1", "<u2", "<u4", "<u8", "<i1", "<i2", "<i4", "<i8", "<f4", "<f8"]
# vgpu_col_kind
DIM_STRING, DIM_NUMERIC, DIM_TIME, DIM_MICROTIME, DIM_BOOLEAN, METRIC_VALUE, METRIC_BITSET, METRIC_HIDDEN_COUNT = range(8)
# vgpu_agg (== db::Metric::AggregationType order)
AGG_MAX, AGG_MIN, AGG_SUM, AGG_AVG, AGG_COUNT, AGG_BITSET = range(6)
AGG_NONE = 255
# vgpu_time_unit (== util::TimeUnit order)
TU_YEAR, TU_MONTH, TU_WEEK, TU_DAY, TU_HOUR, TU_MINUTE, TU_SECOND, TU_NONE = range(8)
# vgpu_node_kind
NODE_RELOP, NODE_IN, NODE_AND, NODE_OR, NODE_EMPTY = range(5)
# vgpu_relop (== query::RelOpFilter::Operator order)
OP_EQ, OP_NE, OP_LT, OP_LE, OP_GT, OP_GE = range(6)
PLAN_FORCE_HASH, PLAN_FORCE_DENSE, PLAN_RESULT_ON_ROOT, PLAN_POST = 1, 2, 4, 8
NO_COLUMN = 0xFFFFFFFF
DEDUPE_SMALL, DEDUPE_FAST, DEDUPE_WIDE, DEDUPE_GENERAL, DEDUPE_REDONE, DEDUPE_PARTITIONED = 1, 2, 4, 8, 16, 32
MAX_ROLLUP_RULES = 8


class Column(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("type", C.c_uint32), ("agg", C.c_uint32), ("lit_type", C.c_uint32)]


class Schema(C.Structure):
    _fields_ = [("ncols", C.c_uint32), ("ndims", C.c_uint32), ("segment_size", C.c_uint64),
                ("cols", C.POINTER(Column))]


class BitsetCsr(C.Structure):
    _fields_ = [("offsets", C.POINTER(C.c_uint64)), ("values", C.c_void_p), ("nvalues", C.c_uint64)]


class PredNode(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("op", C.c_uint32), ("col", C.c_uint32), ("arg", C.c_uint32),
                ("n", C.c_uint32), ("reserved", C.c_uint32)]


class Key(C.Structure):
    _fields_ = [("col", C.c_uint32), ("nrules", C.c_uint32), ("query_granularity", C.c_uint32),
                ("reserved", C.c_uint32), ("rule_boundary", C.c_uint64 * MAX_ROLLUP_RULES),
                ("rule_granularity", C.c_uint32 * MAX_ROLLUP_RULES)]


class Plan(C.Structure):
    _fields_ = [("nnodes", C.c_uint32), ("nargs", C.c_uint32), ("nodes", C.POINTER(PredNode)),
                ("args", C.POINTER(C.c_uint64)), ("nkeys", C.c_uint32), ("nmetrics", C.c_uint32),
                ("keys", C.POINTER(Key)), ("metric_cols", C.POINTER(C.c_uint32)),
                ("need_hidden_count", C.c_uint32), ("flags", C.c_uint32),
                ("nhnodes", C.c_uint32), ("nhargs", C.c_uint32), ("hnodes", C.POINTER(PredNode)),
                ("hargs", C.POINTER(C.c_uint64)), ("sort_col", C.c_uint32), ("sort_descending", C.c_uint32),
                ("top_k", C.c_uint64)]


class ResultView(C.Structure):
    _fields_ = [("ngroups", C.c_uint64), ("nkeys", C.c_uint32), ("nmetrics", C.c_uint32),
                ("keys", C.POINTER(C.c_void_p)), ("accs", C.POINTER(C.c_void_p)),
                ("hidden_count", C.POINTER(C.c_uint64)), ("scanned_recs", C.c_uint64),
                ("scanned_segments", C.c_uint64), ("aggregated_recs", C.c_uint64),
                ("passed_rows", C.c_uint64), ("gpu_ms", C.c_double), ("scan_ms", C.c_double),
                ("launches", C.c_uint32), ("table_mode", C.c_uint32), ("table_cells", C.c_uint64),
                ("attempts", C.c_uint32), ("distinct_paths", C.c_uint32), ("post_applied", C.c_uint32),
                ("reserved", C.c_uint32)]


class RowsPlan(C.Structure):
    _fields_ = [("nnodes", C.c_uint32), ("nargs", C.c_uint32), ("nodes", C.POINTER(PredNode)),
                ("args", C.POINTER(C.c_uint64)), ("ncols", C.c_uint32), ("reserved", C.c_uint32),
                ("cols", C.POINTER(C.c_uint32)), ("skip", C.c_uint64), ("limit", C.c_uint64)]


class RowsView(C.Structure):
    _fields_ = [("nrows", C.c_uint64), ("ncols", C.c_uint32), ("launches", C.c_uint32),
                ("cells", C.POINTER(C.c_void_p)), ("scanned_recs", C.c_uint64), ("scanned_segments", C.c_uint64),
                ("passed_rows", C.c_uint64), ("gpu_ms", C.c_double)]


class SearchPlan(C.Structure):
    _fields_ = [("nnodes", C.c_uint32), ("nargs", C.c_uint32), ("nodes", C.POINTER(PredNode)),
                ("args", C.POINTER(C.c_uint64)), ("col", C.c_uint32), ("reserved", C.c_uint32)]


class SearchView(C.Structure):
    _fields_ = [("nsegments", C.c_uint32), ("launches", C.c_uint32), ("seg_offsets", C.POINTER(C.c_uint64)),
                ("codes", C.POINTER(C.c_uint64)), ("first_row", C.POINTER(C.c_uint32)),
                ("scanned_recs", C.c_uint64), ("scanned_segments", C.c_uint64), ("gpu_ms", C.c_double)]


class GenCol(C.Structure):
    _fields_ = [("lo", C.c_int64), ("range", C.c_uint64), ("mode", C.c_uint32), ("reserved", C.c_uint32),
                ("div", C.c_uint64)]


# every symbol include/vgpu.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("vgpu_abi_version", C.c_int, []),
    ("vgpu_init", C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    ("vgpu_set_stream", C.c_int, [C.c_void_p, C.c_void_p]),
    ("vgpu_set_test_hook", C.c_int, [C.c_void_p, C.c_char_p, C.c_uint64]),
    ("vgpu_shutdown", None, [C.c_void_p]),
    ("vgpu_last_error", C.c_char_p, []),
    ("vgpu_table_create", C.c_int, [C.c_void_p, C.POINTER(Schema), C.POINTER(C.c_void_p)]),
    ("vgpu_table_free", None, [C.c_void_p]),
    ("vgpu_segment_put", C.c_int, [C.c_void_p, C.c_uint32, C.c_uint64, C.POINTER(C.c_void_p)]),
    ("vgpu_segment_put_async", C.c_int, [C.c_void_p, C.c_uint32, C.c_uint64, C.POINTER(C.c_void_p)]),
    ("vgpu_table_sync", C.c_int, [C.c_void_p]),
    ("vgpu_segment_update", C.c_int, [C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint64, C.POINTER(C.c_void_p)]),
    ("vgpu_host_pin", C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    ("vgpu_host_unpin", C.c_int, [C.c_void_p, C.c_void_p]),
    ("vgpu_table_invalidate", C.c_int, [C.c_void_p, C.c_uint32]),
    ("vgpu_table_segments", C.c_uint32, [C.c_void_p]),
    ("vgpu_table_rows", C.c_uint64, [C.c_void_p]),
    ("vgpu_table_bytes", C.c_uint64, [C.c_void_p]),
    ("vgpu_segment_generate", C.c_int, [C.c_void_p, C.c_uint32, C.c_uint64, C.POINTER(GenCol), C.c_uint64, C.c_uint64]),
    ("vgpu_segment_read", C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]),
    ("vgpu_query_agg", C.c_int, [C.c_void_p, C.POINTER(Plan), C.POINTER(C.c_void_p)]),
    ("vgpu_result_get", C.c_int, [C.c_void_p, C.POINTER(ResultView)]),
    ("vgpu_result_free", None, [C.c_void_p]),
    ("vgpu_query_select", C.c_int, [C.c_void_p, C.POINTER(RowsPlan), C.POINTER(C.c_void_p)]),
    ("vgpu_rows_get", C.c_int, [C.c_void_p, C.POINTER(RowsView)]),
    ("vgpu_rows_free", None, [C.c_void_p]),
    ("vgpu_query_search", C.c_int, [C.c_void_p, C.POINTER(SearchPlan), C.POINTER(C.c_void_p)]),
    ("vgpu_search_get", C.c_int, [C.c_void_p, C.POINTER(SearchView)]),
    ("vgpu_search_free", None, [C.c_void_p]),
    ("vgpu_comm_unique_id", C.c_int, [C.c_void_p]),
    ("vgpu_comm_init", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    ("vgpu_comm_destroy", C.c_int, [C.c_void_p]),
]


class VgpuError(RuntimeError):
    """Raised for any non-zero vgpu_status; mirrors the std::runtime_error the C++ adapter throws."""

    def __init__(self, code, msg):
        super().__init__(f"vgpu error {code}: {msg}")
        self.code = code
        self.msg = msg


_lib = None


def load():
    """dlopen libvgpu.so and bind every declared symbol. Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "viyadb_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.vgpu_abi_version() != VGPU_ABI_VERSION:
        raise RuntimeError("libvgpu.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise VgpuError(rc, load().vgpu_last_error().decode("utf-8", "replace"))
