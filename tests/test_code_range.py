"""qsgpu_dictionary_code_range: the host arithmetic behind every comparison on a dictionary-coded attribute
(the translation CompressedTupleStorageSubBlock::getMatchesForPredicate does with the limit codes of a block's
sorted dictionary, storage/CompressedTupleStorageSubBlock.cpp:160-251).  Device-free: checked here against a
brute-force evaluation of the comparison over every dictionary entry, for every type, comparison and literal
position (below / equal / between / above the entries), with the reference's type promotion."""
import ctypes as C

import numpy as np
import pytest

from quickstep_b200 import capi as A
from quickstep_b200.expr import ExprSet
from quickstep_b200.table import DATE_DTYPE, make_dates

CMPS = [A.QS_EQ, A.QS_NE, A.QS_LT, A.QS_LE, A.QS_GT, A.QS_GE]
PY = {A.QS_EQ: lambda a, b: a == b, A.QS_NE: lambda a, b: a != b, A.QS_LT: lambda a, b: a < b,
      A.QS_LE: lambda a, b: a <= b, A.QS_GT: lambda a, b: a > b, A.QS_GE: lambda a, b: a >= b}


def code_range(attr_type, width, dict_values, cmp, es, lit_index):
    d = np.ascontiguousarray(dict_values)
    first, count, neg = C.c_uint32(0), C.c_uint32(0), C.c_int(0)
    c = es.c()
    A.check(A.load().qsgpu_dictionary_code_range(attr_type, width, d.ctypes.data, len(d), cmp, C.byref(c.nodes[lit_index]),
                                                c.str_pool, c.str_pool_bytes, C.byref(first), C.byref(count), C.byref(neg)))
    sel = np.zeros(len(d), dtype=bool)
    sel[first.value: first.value + count.value] = True
    return ~sel if neg.value else sel


def _unify(a, b):
    """TypeFactory::GetUnifyingType for numeric pairs (types/TypeFactory.cpp:159-180)."""
    if a == b:
        return a
    if A.QS_DOUBLE in (a, b) or {a, b} == {A.QS_LONG, A.QS_FLOAT}:
        return A.QS_DOUBLE
    if A.QS_FLOAT in (a, b):
        return A.QS_FLOAT
    return A.QS_LONG


_CAST = {A.QS_INT: int, A.QS_LONG: int, A.QS_FLOAT: np.float32, A.QS_DOUBLE: float}


@pytest.mark.parametrize("cmp", CMPS)
def test_numeric_dictionaries(cmp):
    rng = np.random.default_rng(cmp)
    cases = [(A.QS_INT, np.unique(rng.integers(-50, 50, 40)).astype(np.int32)),
             (A.QS_LONG, np.unique(rng.integers(-10**6, 10**6, 40))),
             (A.QS_FLOAT, np.unique(rng.normal(0, 10, 40).astype(np.float32))),
             (A.QS_DOUBLE, np.unique(np.round(rng.normal(0, 10, 40), 2)))]
    makers = [(A.QS_INT, lambda es, x: es.lit_int(int(x))), (A.QS_LONG, lambda es, x: es.lit_long(int(x))),
              (A.QS_FLOAT, lambda es, x: es.lit_float(float(np.float32(x)))), (A.QS_DOUBLE, lambda es, x: es.lit_double(float(x)))]
    for t, d in cases:
        probes = [d[0] - 1, d[0], d[len(d) // 2], (float(d[3]) + float(d[4])) / 2, d[-1], d[-1] + 1]
        for v in probes:
            for lt, make in makers:
                es = ExprSet()
                li = make(es, v)
                got = code_range(t, 0, d, cmp, es, li)
                lit = _CAST[lt](v)                      # the literal's own value (an int literal truncates)
                T = _CAST[_unify(t, lt)]                # both sides are cast to the unifying type, then compared
                want = np.array([PY[cmp](T(_CAST[t](x)), T(lit)) for x in d])
                assert (got == want).all(), (t, v, lt)
        es = ExprSet()
        li = es.lit_double(float("nan"))
        got = code_range(t, 0, d, cmp, es, li)
        assert got.all() if cmp == A.QS_NE else not got.any()


@pytest.mark.parametrize("cmp", CMPS)
def test_date_and_char_dictionaries(cmp):
    dates = make_dates([1994, 1994, 1995, 1995, 1998], [1, 12, 6, 6, 9], [1, 31, 16, 17, 2])
    key = lambda y, m, dd: (y, m, dd)
    for (y, m, dd) in [(1993, 5, 5), (1994, 1, 1), (1995, 6, 17), (1995, 1, 1), (1998, 9, 2), (1999, 1, 1)]:
        es = ExprSet()
        li = es.lit_date(y, m, dd)
        got = code_range(A.QS_DATE, 8, dates.view(DATE_DTYPE), cmp, es, li)
        want = np.array([PY[cmp](key(int(x["year"]), int(x["month"]), int(x["day"])), (y, m, dd)) for x in dates])
        assert (got == want).all(), (y, m, dd)
    words = np.array(sorted([b"", b"AIR", b"MAIL", b"RAIL", b"SHIP", b"TRUCK", b"REG AIR"]), dtype="S10")
    for lit in [b"", b"AIR", b"B", b"REG AIR", b"SHIP", b"ZZZ", b"MAILX"]:
        es = ExprSet()
        li = es.lit_char(lit)
        got = code_range(A.QS_CHAR, 10, words, cmp, es, li)
        want = np.array([PY[cmp](bytes(w), lit) for w in words])
        assert (got == want).all(), lit


def test_rejects_mixed_kinds():
    es = ExprSet()
    li = es.lit_char(b"x")
    with pytest.raises(A.QsGpuError):
        code_range(A.QS_INT, 0, np.array([1, 2], dtype=np.int32), A.QS_EQ, es, li)


@pytest.mark.parametrize("cmp", CMPS)
def test_char_literal_longer_than_the_attribute(cmp):
    """col CHAR(3) <cmp> 'abcdef' compares the FULL strings (AsciiStringComparators.hpp:218-251): the entry
    'abc' is less than the literal, never equal to its first three bytes."""
    d = np.array([b"ab", b"abc", b"abd", b"b"], dtype="S3")
    for lit in (b"abcdef", b"abc", b"abcd", b"ab", b"abz9"):
        es = ExprSet()
        li = es.lit_char(lit)
        got = code_range(A.QS_CHAR, 3, d, cmp, es, li)
        want = np.array([PY[cmp](bytes(x), lit) for x in d])       # Python bytes compare like strcmp on NUL-free strings
        assert (got == want).all(), (cmp, lit, got, want)
