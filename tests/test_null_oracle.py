"""The NULL oracle (oracle/qs_null_oracle.py) against the fixture of the reference's aggregation-handle unit
tests (expressions/aggregation/tests/AggregationHandleSum_unittest.cpp:164-184 createColumnVectorGeneric: a NULL, the
samples i - 10 with one NULL in the middle, a NULL; the handle must return their plain sum) and the truth tables
of comparison / negation over NULL operands."""
import numpy as np
import pytest

import qs_null_oracle as NO
from quickstep_b200 import capi as A
from quickstep_b200.expr import ExprSet
from quickstep_b200.table import Column, HostTable


@pytest.fixture(scope="module", autouse=True)
def _oracle(oracle):
    return oracle


def fixture_column(np_dtype, n_samples=100):
    vals, isnull = [0], [True]
    for i in range(n_samples):
        vals.append(i - 10 if np.dtype(np_dtype).kind == "i" else np.float32(i - 10) / np.float32(10))
        isnull.append(False)
        if i == n_samples // 2:
            vals.append(0); isnull.append(True)
    vals.append(0); isnull.append(True)
    return np.array(vals, dtype=np_dtype), np.array(isnull)


@pytest.mark.parametrize("qs_type,np_dtype", [(A.QS_INT, np.int32), (A.QS_LONG, np.int64), (A.QS_FLOAT, np.float32), (A.QS_DOUBLE, np.float64)])
def test_handle_fixture(qs_type, np_dtype):
    vals, isnull = fixture_column(np_dtype)
    t = HostTable("t", [Column("x", qs_type, vals)])
    es = ExprSet()
    x = es.attr(0, qs_type)
    aggs = [(A.QS_AGG_SUM, x), (A.QS_AGG_AVG, x), (A.QS_AGG_COUNT, x), (A.QS_AGG_MIN, x), (A.QS_AGG_MAX, x), (A.QS_AGG_COUNT, -1)]
    res = NO.aggregate(es, -1, aggs, None, t, isnull.astype(np.uint64))[None]
    good = vals[~isnull]
    seq = 0.0 if np.dtype(np_dtype).kind == "f" else 0
    for v in good:
        seq += float(v) if np.dtype(np_dtype).kind == "f" else int(v)
    assert res[0] == (seq, False) or abs(res[0][0] - seq) <= 1e-12 * abs(seq)
    if np.dtype(np_dtype).kind == "i":
        assert res[0] == (3950, False)
    assert res[2] == (100, False) and res[5] == (103, False)
    assert res[3] == (good.min().item(), False) and res[4] == (good.max().item(), False)
    assert abs(res[1][0] - seq / 100.0) <= 1e-12 * abs(seq / 100.0)
    # nothing but NULLs: the handles' finalize() is NULL, COUNT(x) is 0
    res = NO.aggregate(es, -1, aggs, None, t, np.ones(len(vals), dtype=np.uint64))[None]
    assert [r[1] for r in res] == [True, True, False, True, True, False] and res[2][0] == 0 and res[5][0] == 103


def test_comparison_truth_table():
    t = HostTable("t", [Column("x", A.QS_INT, np.array([1, 9, 0, 0], dtype=np.int32)),
                        Column("y", A.QS_INT, np.array([5, 5, 5, 0], dtype=np.int32))])
    nulls = np.array([0, 0, 0b01, 0b11], dtype=np.uint64)          # row 2: x NULL; row 3: both NULL
    es = ExprSet()
    x, y = es.attr(0, A.QS_INT), es.attr(1, A.QS_INT)
    lt = es.cmp(A.QS_LT, x, y)
    assert NO.predicate(es, lt, t, nulls).tolist() == [True, False, False, False]
    assert NO.predicate(es, es.not_(lt), t, nulls).tolist() == [False, True, True, True]     # complement, no UNKNOWN
    eq = es.cmp(A.QS_EQ, x, y)
    assert NO.predicate(es, eq, t, nulls).tolist() == [False, False, False, False]           # NULL = NULL is not true
    assert NO.predicate(es, es.or_(lt, es.cmp(A.QS_EQ, y, es.lit_int(5))), t, nulls).tolist() == [True, True, True, False]
    assert NO.null_of(es, es.add(x, es.mul(y, es.lit_int(2))), nulls).tolist() == [False, False, True, True]
