"""ctypes wrapper of oracle/libqsoracle.so (qs_oracle.h).

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs -- never by quickstep_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from quickstep_b200 import capi as A  # noqa: E402  (struct definitions of include/qsgpu.h only)
from quickstep_b200.table import HostTable, np_dtype  # noqa: E402

LIB_PATH = os.path.join(_HERE, "libqsoracle.so")


class qso_column(C.Structure):
    _fields_ = [("data", C.c_void_p), ("type", C.c_uint16), ("width", C.c_uint16)]


class qso_table(C.Structure):
    _fields_ = [("cols", C.POINTER(qso_column)), ("n_cols", C.c_uint32), ("n_rows", C.c_uint64)]


class qso_lip(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("is_anti", C.c_uint32), ("min_value", C.c_int64),
                ("max_value", C.c_int64), ("cardinality", C.c_uint64), ("words", C.POINTER(C.c_uint64))]


class qso_lip_ref(C.Structure):
    _fields_ = [("lip", C.POINTER(qso_lip)), ("attr", C.c_uint32)]


class qso_agg_result(C.Structure):
    _fields_ = [("n_groups", C.c_uint64), ("key_bytes", C.c_uint32), ("keys", C.POINTER(C.c_uint8)),
                ("values", C.POINTER(C.c_uint64)), ("is_double", C.c_uint8 * 16), ("is_null", C.c_uint8 * 16),
                ("counts", C.POINTER(C.c_int64))]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "libqsoracle.so"])


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    L = C.CDLL(LIB_PATH)
    L.qso_set_num_workers.argtypes = [C.c_int]
    L.qso_get_num_workers.restype = C.c_int
    L.qso_set_block_rows.argtypes = [C.c_uint64]
    L.qso_lip_words.restype = C.c_uint64
    L.qso_lip_words.argtypes = [C.POINTER(qso_lip)]
    L.qso_predicate.restype = C.c_int64
    L.qso_predicate.argtypes = [C.POINTER(A.qs_expr_set), C.c_int32, C.POINTER(qso_table), C.POINTER(C.c_uint64)]
    L.qso_scalar.restype = C.c_int
    L.qso_scalar.argtypes = [C.POINTER(A.qs_expr_set), C.c_int32, C.POINTER(qso_table), C.c_void_p]
    L.qso_build_lip_filter.restype = C.c_int
    L.qso_build_lip_filter.argtypes = [C.POINTER(A.qs_expr_set), C.c_int32, C.POINTER(qso_table), C.c_uint32,
                                       C.POINTER(qso_lip_ref), C.c_uint32, C.POINTER(qso_lip_ref)]
    L.qso_select.restype = C.c_int64
    L.qso_select.argtypes = [C.POINTER(A.qs_expr_set), C.c_int32, C.POINTER(qso_table), C.c_uint32,
                             C.POINTER(qso_lip_ref), C.c_uint32, C.POINTER(C.c_int32), C.POINTER(C.c_void_p)]
    L.qso_aggregate.restype = C.c_int
    L.qso_aggregate.argtypes = [C.POINTER(A.qs_expr_set), C.c_int32, C.c_uint32, C.POINTER(A.qs_aggregate),
                                C.c_uint32, C.POINTER(C.c_int32), C.POINTER(qso_table), C.c_uint32,
                                C.POINTER(qso_lip_ref), C.POINTER(qso_agg_result)]
    L.qso_agg_result_free.argtypes = [C.POINTER(qso_agg_result)]
    L.qso_hash_join.restype = C.c_int64
    L.qso_hash_join.argtypes = [C.POINTER(A.qs_expr_set), C.POINTER(qso_table), C.c_int32, C.c_uint32,
                                C.POINTER(qso_table), C.c_int32, C.c_uint32, C.c_uint32, C.POINTER(qso_lip_ref),
                                C.c_uint32, C.c_int32, C.c_uint32, C.POINTER(C.c_int32), C.POINTER(C.c_void_p),
                                C.c_uint64]
    L.qso_hash_join_nulls.restype = C.c_int64
    L.qso_hash_join_nulls.argtypes = L.qso_hash_join.argtypes + [C.POINTER(C.c_uint64)]
    L.qso_topk.restype = C.c_int64
    L.qso_topk.argtypes = [C.POINTER(qso_table), C.c_uint32, C.POINTER(A.qs_sort_key), C.c_uint64,
                           C.POINTER(C.c_uint64)]
    for fn in ("qso_decode_dict",):
        getattr(L, fn).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32]
    L.qso_decode_truncated.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32]
    L.qso_decode_strided.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32]
    L.qso_partition_of.restype = C.c_uint32
    L.qso_partition_of.argtypes = [C.c_int64, C.c_uint32]
    _lib = L
    return L


def set_workers(n: int):
    load().qso_set_num_workers(n)


def set_block_rows(n: int):
    load().qso_set_block_rows(n)


class _CTable:
    """Keeps the ctypes view of a HostTable alive."""

    def __init__(self, t: HostTable):
        self.cols = (qso_column * max(1, len(t.columns)))()
        for i, c in enumerate(t.columns):
            self.cols[i].data = c.data.ctypes.data
            self.cols[i].type = c.type
            self.cols[i].width = c.width
        self.t = qso_table(C.cast(self.cols, C.POINTER(qso_column)), len(t.columns), t.n_rows)
        self.keep = t

    def ptr(self):
        return C.byref(self.t)


class Lip:
    def __init__(self, kind, min_value=0, max_value=0, cardinality=0, is_anti=False):
        self.c = qso_lip(kind, 1 if is_anti else 0, min_value, max_value, cardinality, None)
        n = load().qso_lip_words(C.byref(self.c))
        self.words = np.zeros(max(1, n), dtype=np.uint64)
        self.n_words = n
        self.c.words = self.words.ctypes.data_as(C.POINTER(C.c_uint64))


def _lip_refs(refs):
    """refs: list of (Lip, attr_id)."""
    if not refs:
        return 0, None, None
    arr = (qso_lip_ref * len(refs))()
    for i, (lip, attr) in enumerate(refs):
        arr[i].lip = C.pointer(lip.c)
        arr[i].attr = attr
    return len(refs), arr, refs


def _i32arr(v):
    return (C.c_int32 * max(1, len(v)))(*v)


def predicate(es, root, table: HostTable):
    L = load()
    ct = _CTable(table)
    bm = np.zeros((table.n_rows + 63) // 64 + 1, dtype=np.uint64)
    n = L.qso_predicate(es.ptr(), root, ct.ptr(), bm.ctypes.data_as(C.POINTER(C.c_uint64)))
    if n < 0:
        raise RuntimeError("oracle: unsupported predicate")
    return n, bm


def bitmap_to_bool(bm: np.ndarray, n: int) -> np.ndarray:
    """MSB-first 64-bit words (utility/BitVector.hpp:934) -> bool array."""
    be = bm.astype(">u8").view(np.uint8)
    return np.unpackbits(be)[:n].astype(bool)


_TYPE_OF_RESULT = {A.QS_INT: "<i4", A.QS_LONG: "<i8", A.QS_FLOAT: "<f4", A.QS_DOUBLE: "<f8"}


def scalar(es, root, table: HostTable):
    L = load()
    ct = _CTable(table)
    buf = np.zeros(max(1, table.n_rows), dtype=np.uint64)
    ty = L.qso_scalar(es.ptr(), root, ct.ptr(), buf.ctypes.data)
    if ty < 0:
        raise RuntimeError("oracle: unsupported scalar")
    return buf.view(np.uint8)[: table.n_rows * np.dtype(_TYPE_OF_RESULT[ty]).itemsize].view(_TYPE_OF_RESULT[ty]).copy()


def build_lip_filter(es, pred_root, table, probe_refs, build_refs):
    L = load()
    ct = _CTable(table)
    np_, pa, _k1 = _lip_refs(probe_refs)
    nb, ba, _k2 = _lip_refs(build_refs)
    rc = L.qso_build_lip_filter(es.ptr() if es else None, pred_root, ct.ptr(), np_, pa, nb, ba)
    if rc != 0:
        raise RuntimeError("oracle: build_lip_filter failed")


def select(es, pred_root, table, probe_refs, project_roots, out_types):
    """out_types: list of (type_id, width).  Returns list of numpy columns."""
    L = load()
    ct = _CTable(table)
    outs = [np.zeros(max(1, table.n_rows), dtype=np_dtype(t, w)) for (t, w) in out_types]
    ptrs = (C.c_void_p * max(1, len(outs)))(*[o.ctypes.data for o in outs])
    np_, pa, _k = _lip_refs(probe_refs)
    n = L.qso_select(es.ptr(), pred_root, ct.ptr(), np_, pa, len(project_roots), _i32arr(project_roots), ptrs)
    if n < 0:
        raise RuntimeError("oracle: select failed")
    return [o[:n] for o in outs]


class AggResult:
    def __init__(self, n_groups, key_bytes, keys, values, counts, is_double, is_null):
        self.n_groups, self.key_bytes = n_groups, key_bytes
        self.keys = keys            # uint8 [n_groups, key_bytes], sorted by key bytes
        self.values = values        # list per aggregate: int64 or float64 array [n_groups]
        self.counts = counts
        self.is_double, self.is_null = is_double, is_null


def aggregate(es, pred_root, aggregates, group_by_roots, table, probe_refs=None) -> AggResult:
    """aggregates: list of (function, argument_root)."""
    L = load()
    ct = _CTable(table)
    aggs = (A.qs_aggregate * max(1, len(aggregates)))()
    for i, (f, r) in enumerate(aggregates):
        aggs[i].function = f
        aggs[i].argument_root = r
    np_, pa, _k = _lip_refs(probe_refs)
    res = qso_agg_result()
    rc = L.qso_aggregate(es.ptr(), pred_root, len(aggregates), aggs, len(group_by_roots), _i32arr(group_by_roots),
                         ct.ptr(), np_, pa, C.byref(res))
    if rc != 0:
        raise RuntimeError("oracle: aggregate failed")
    G, kb = res.n_groups, res.key_bytes
    keys = np.ctypeslib.as_array(res.keys, shape=(max(1, G * max(kb, 1)),)).copy()[: G * kb].reshape(G, kb)
    raw = np.ctypeslib.as_array(res.values, shape=(max(1, G * max(1, len(aggregates))),)).copy()
    counts = np.ctypeslib.as_array(res.counts, shape=(max(1, G),)).copy()[:G]
    vals = []
    for a in range(len(aggregates)):
        w = raw[a * G:(a + 1) * G]
        vals.append(w.view(np.float64).copy() if res.is_double[a] else w.view(np.int64).copy())
    out = AggResult(G, kb, keys, vals, counts, [bool(res.is_double[a]) for a in range(len(aggregates))],
                    [bool(res.is_null[a]) for a in range(len(aggregates))])
    L.qso_agg_result_free(C.byref(res))
    return out


def hash_join(es, build, build_pred, build_key_attr, probe, probe_pred, probe_key_attr, probe_refs, join_type,
              residual_root, project_roots, out_types, capacity):
    L = load()
    cb, cp = _CTable(build), _CTable(probe)
    outs = [np.zeros(max(1, capacity), dtype=np_dtype(t, w)) for (t, w) in out_types]
    ptrs = (C.c_void_p * max(1, len(outs)))(*[o.ctypes.data for o in outs])
    np_, pa, _k = _lip_refs(probe_refs)
    nulls = np.zeros(max(1, capacity), dtype=np.uint64)
    n = L.qso_hash_join_nulls(es.ptr(), cb.ptr(), build_pred, build_key_attr, cp.ptr(), probe_pred, probe_key_attr,
                              np_, pa, join_type, residual_root, len(project_roots), _i32arr(project_roots), ptrs,
                              capacity, nulls.ctypes.data_as(C.POINTER(C.c_uint64)))
    if n < 0:
        raise RuntimeError("oracle: hash_join failed")
    if n > capacity:
        raise RuntimeError(f"oracle: join output {n} exceeds capacity {capacity}")
    hash_join.last_nulls = nulls[:n]          # NULL mask of the rows just returned (LEFT OUTER joins)
    return [o[:n] for o in outs]


def topk(table, keys, limit):
    """keys: list of (attr, descending)."""
    L = load()
    ct = _CTable(table)
    ks = (A.qs_sort_key * max(1, len(keys)))()
    for i, (a, d) in enumerate(keys):
        ks[i].attr, ks[i].descending = a, 1 if d else 0
    ids = np.zeros(max(1, min(limit, table.n_rows)), dtype=np.uint64)
    n = L.qso_topk(ct.ptr(), len(keys), ks, limit, ids.ctypes.data_as(C.POINTER(C.c_uint64)))
    return ids[:n]


def decode_dict(codes: np.ndarray, dict_values: np.ndarray) -> np.ndarray:
    out = np.zeros(len(codes), dtype=dict_values.dtype)
    load().qso_decode_dict(out.ctypes.data, codes.ctypes.data, dict_values.ctypes.data, len(codes),
                           codes.dtype.itemsize, dict_values.dtype.itemsize)
    return out


def decode_truncated(codes: np.ndarray, value_dtype) -> np.ndarray:
    out = np.zeros(len(codes), dtype=value_dtype)
    load().qso_decode_truncated(out.ctypes.data, codes.ctypes.data, len(codes), codes.dtype.itemsize,
                                out.dtype.itemsize)
    return out


def decode_strided(slots: np.ndarray, n: int, stride: int, value_dtype) -> np.ndarray:
    out = np.zeros(n, dtype=value_dtype)
    load().qso_decode_strided(out.ctypes.data, slots.ctypes.data, n, stride, out.dtype.itemsize)
    return out


def partition_of(key: int, n_parts: int) -> int:
    return load().qso_partition_of(key, n_parts)
