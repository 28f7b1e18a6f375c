"""ctypes binding of libqshost.so (include/qshost.h): the C++ operator layer
(quickstep_b200/host/) driven as whole TPC-H queries.  Harness plumbing for tests/ and bench.py."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import capi as A

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libqshost.so")
CUSTOMER, ORDERS, LINEITEM = 0, 1, 2
BASIC_COLUMN_STORE, COMPRESSED_COLUMN_STORE, SPLIT_ROW_STORE = 0, 1, 2


class q1_row(C.Structure):
    _fields_ = [("l_returnflag", C.c_char), ("l_linestatus", C.c_char), ("pad", C.c_char * 6),
                ("sum_qty", C.c_double), ("sum_base_price", C.c_double), ("sum_disc_price", C.c_double),
                ("sum_charge", C.c_double), ("avg_qty", C.c_double), ("avg_price", C.c_double),
                ("avg_disc", C.c_double), ("count_order", C.c_int64), ("sum_disc", C.c_double)]


class q3_row(C.Structure):
    _fields_ = [("l_orderkey", C.c_int32), ("o_shippriority", C.c_int32), ("revenue", C.c_double),
                ("year", C.c_int32), ("month", C.c_uint8), ("day", C.c_uint8), ("pad", C.c_uint8 * 2)]


_VP, _U64P = C.c_void_p, C.POINTER(C.c_uint64)
SIGNATURES = {
    "qshost_db_create": [C.c_int, C.c_int, C.POINTER(_VP)],
    "qshost_db_destroy": [_VP],
    "qshost_db_set_comm": [_VP, _VP],
    "qshost_db_set_join_mode": [_VP, C.c_int],
    "qshost_db_load": [_VP, C.c_int, C.POINTER(_VP), C.c_uint64, C.c_uint64, C.c_int],
    "qshost_db_evict": [_VP, C.c_int],
    "qshost_db_stats": [_VP, C.c_int, _U64P, _U64P, _U64P],
    "qshost_db_set_code_resident": [_VP, C.c_int],
    "qshost_db_resident_coding": [_VP, C.c_int, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)],
    "qshost_set_rows_per_workorder": [C.c_uint64],
    "qshost_q1": [_VP, C.POINTER(q1_row), C.POINTER(C.c_uint32), _U64P],
    "qshost_q6": [_VP, C.POINTER(C.c_double), C.POINTER(C.c_int), _U64P],
    "qshost_q3": [_VP, C.POINTER(q3_row), C.POINTER(C.c_uint32), _U64P],
    "qshost_last_profile": [_VP, C.c_char_p, C.c_uint64],
    "qshost_result_block": [_VP, C.c_uint32, C.POINTER(_VP), _U64P, _U64P],
}
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it (make -C quickstep_b200/host). No CPU fallback.")
        A.load()                                  # libqsgpu.so first (same directory, rpath $ORIGIN)
        lib = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = C.c_int, args
        _lib = lib
    return _lib


class Database:
    """qshost_db_t: storage blocks + device residency + Worker threads."""

    def __init__(self, dev=0, num_workers=4):
        self.h = _VP()
        A.check(load().qshost_db_create(dev, num_workers, C.byref(self.h)))
        self._keep = {}

    def load(self, which, arrays, rows_per_block=0, layout=BASIC_COLUMN_STORE):
        """arrays: native-width numpy columns in schema order (host memory)."""
        arrays = [np.ascontiguousarray(a) for a in arrays]
        self._keep[which] = arrays
        ptrs = (_VP * len(arrays))(*[a.ctypes.data for a in arrays])
        A.check(load().qshost_db_load(self.h, which, ptrs, len(arrays[0]), rows_per_block, layout))

    def load_table(self, which, table, rows_per_block=0, layout=BASIC_COLUMN_STORE):
        self.load(which, [c.data for c in table.columns], rows_per_block, layout)

    def set_comm(self, comm):
        """comm: a qsgpu_comm_t (engine.Comm.h); relations loaded afterwards are this rank's partitions."""
        A.check(load().qshost_db_set_comm(self.h, comm))

    def set_join_mode(self, mode: int):
        """0: partition-wise join of co-partitioned orders / lineitem (default); 1: broadcast the build side."""
        A.check(load().qshost_db_set_join_mode(self.h, mode))

    def evict(self, which):
        A.check(load().qshost_db_evict(self.h, which))

    def set_code_resident(self, on: bool):
        """Dictionary-compressed attributes stay codes in HBM (scans run on the codes); re-stages on next use."""
        A.check(load().qshost_db_set_code_resident(self.h, 1 if on else 0))

    def resident_coding(self, which, attr):
        cw, n = C.c_uint32(0), C.c_uint32(0)
        A.check(load().qshost_db_resident_coding(self.h, which, attr, C.byref(cw), C.byref(n)))
        return cw.value, n.value

    def stats(self, which):
        b, n, r = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        A.check(load().qshost_db_stats(self.h, which, C.byref(b), C.byref(n), C.byref(r)))
        return dict(host_bytes=b.value, n_blocks=n.value, n_rows=r.value)

    def q1(self):
        rows = (q1_row * 16)()
        n, wo = C.c_uint32(16), C.c_uint64(0)
        A.check(load().qshost_q1(self.h, rows, C.byref(n), C.byref(wo)))
        out = []
        for r in rows[: n.value]:
            out.append(dict(l_returnflag=r.l_returnflag, l_linestatus=r.l_linestatus, sum_qty=r.sum_qty,
                            sum_base_price=r.sum_base_price, sum_disc_price=r.sum_disc_price,
                            sum_charge=r.sum_charge, avg_qty=r.avg_qty, avg_price=r.avg_price,
                            avg_disc=r.avg_disc, count_order=r.count_order, sum_disc=r.sum_disc))
        return out, wo.value

    def q6(self):
        rev, null, wo = C.c_double(0), C.c_int(0), C.c_uint64(0)
        A.check(load().qshost_q6(self.h, C.byref(rev), C.byref(null), C.byref(wo)))
        return rev.value, bool(null.value), wo.value

    def q3(self):
        rows = (q3_row * 16)()
        n, wo = C.c_uint32(16), C.c_uint64(0)
        A.check(load().qshost_q3(self.h, rows, C.byref(n), C.byref(wo)))
        return [(r.l_orderkey, r.revenue, (r.year, r.month, r.day), r.o_shippriority) for r in rows[: n.value]], wo.value

    def result_blocks(self):
        """The last query's result relation as host SplitRowStore blocks (reference layout): [(bytes, n_tuples)]."""
        out, i = [], 0
        while True:
            mem, nb, nt = _VP(), C.c_uint64(0), C.c_uint64(0)
            if load().qshost_result_block(self.h, i, C.byref(mem), C.byref(nb), C.byref(nt)) != 0:
                return out
            out.append((C.string_at(mem.value, nb.value), nt.value))
            i += 1

    def last_profile(self) -> str:
        buf = C.create_string_buffer(8192)
        A.check(load().qshost_last_profile(self.h, buf, 8192))
        return buf.value.decode()

    def destroy(self):
        if self.h:
            load().qshost_db_destroy(self.h)
        self.h = None


def set_rows_per_workorder(rows: int):
    A.check(load().qshost_set_rows_per_workorder(rows))
