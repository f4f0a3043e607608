"""viyadb_b200 — B200-native (sm_100a) scan -> filter -> group-by-aggregate path of ViyaDB.

Layout:
  csrc/       hand-written CUDA kernels + the C ABI (include/vgpu.h) -> libvgpu.so
  _native.py  ctypes binding of the C ABI (fails loudly when the library / GPU is missing)
  db.py       mirror of db::Table / Column / DimensionDict + the HBM column store
  query.py    mirror of query::AggregateQuery / FilterFactory / QueryRunner + host post-aggregation
  host/       C++ adapter sources that plug the C ABI behind the reference's query::QueryVisitor
"""
from ._native import VgpuError, load  # noqa: F401
from .db import Database, Table, DimensionDict  # noqa: F401
from .query import (AggregateQuery, FilterFactory, GpuQueryRunner, MemoryRowOutput, QueryFactory,  # noqa: F401
                    QueryStats)
