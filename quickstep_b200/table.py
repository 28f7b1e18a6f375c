"""Host-side description of a relation's columns (numpy buffers in Quickstep's
native value layouts).  Pure bookkeeping: no query arithmetic happens here.

Layouts (types/TypeID.hpp:33-45, types/DatetimeLit.hpp:38-93):
  INT int32 | LONG int64 | FLOAT float32 | DOUBLE float64 (= SQL DECIMAL) |
  CHAR(n) n bytes NUL padded | DATE 8 bytes {int32 year; u8 month; u8 day; 2 pad}
"""
from __future__ import annotations

import numpy as np

from . import capi as A

DATE_DTYPE = np.dtype([("year", "<i4"), ("month", "u1"), ("day", "u1"), ("pad", "<u2")])
assert DATE_DTYPE.itemsize == 8

_NP = {A.QS_INT: np.dtype("<i4"), A.QS_LONG: np.dtype("<i8"), A.QS_FLOAT: np.dtype("<f4"),
       A.QS_DOUBLE: np.dtype("<f8"), A.QS_DATE: DATE_DTYPE}


def np_dtype(type_id: int, width: int = 0) -> np.dtype:
    if type_id == A.QS_CHAR:
        return np.dtype(f"S{width}")
    return _NP[type_id]


def type_width(type_id: int, width: int = 0) -> int:
    return np_dtype(type_id, width).itemsize


def make_dates(year, month, day) -> np.ndarray:
    year = np.asarray(year)
    out = np.zeros(year.shape, dtype=DATE_DTYPE)
    out["year"] = year
    out["month"] = month
    out["day"] = day
    return out


def days_to_dates(days_since_epoch) -> np.ndarray:
    """numpy datetime64[D] day numbers -> DateLit array."""
    d = np.asarray(days_since_epoch).astype("datetime64[D]")
    y = d.astype("datetime64[Y]")
    m = d.astype("datetime64[M]")
    year = y.astype(np.int64) + 1970
    month = (m.astype(np.int64) - y.astype("datetime64[M]").astype(np.int64)) + 1
    day = (d.astype(np.int64) - m.astype("datetime64[D]").astype(np.int64)) + 1
    return make_dates(year, month, day)


class Column:
    __slots__ = ("name", "type", "width", "data")

    def __init__(self, name: str, type_id: int, data: np.ndarray, width: int = 0):
        dt = np_dtype(type_id, width or (data.dtype.itemsize if type_id == A.QS_CHAR else 0))
        data = np.ascontiguousarray(data)
        if data.dtype != dt:
            if data.dtype.itemsize != dt.itemsize or type_id not in (A.QS_DATE, A.QS_CHAR):
                data = data.astype(dt)
            else:
                data = data.view(dt)
        self.name, self.type, self.width, self.data = name, type_id, dt.itemsize, data


class HostTable:
    def __init__(self, name: str, columns: list[Column]):
        self.name = name
        self.columns = columns
        n = {len(c.data) for c in columns}
        assert len(n) == 1, "ragged table"
        self.n_rows = n.pop()
        self.index = {c.name: i for i, c in enumerate(columns)}

    def attr_id(self, name: str) -> int:
        return self.index[name]

    def col(self, name: str) -> Column:
        return self.columns[self.index[name]]

    def slice(self, lo: int, hi: int) -> "HostTable":
        return HostTable(self.name, [Column(c.name, c.type, c.data[lo:hi], c.width) for c in self.columns])

    def project(self, names) -> "HostTable":
        return HostTable(self.name, [self.col(n) for n in names])

    def attr(self, es, name: str, side: int = 0) -> int:
        """ScalarAttribute node for column `name` in ExprSet `es`."""
        c = self.col(name)
        return es.attr(self.index[name], c.type, c.width, side)
