"""Pins the CPU oracle (oracle/qs_oracle.c) against the reference's own golden
vectors before anything is compared with it (CPU only, no GPU needed):

  * literal expected tables of the reference's SQL tests (LIP.test, Select.test, Partition.test),
  * closed-form expectations of AggregationOperator_unittest.cpp / HashJoinOperator_unittest.cpp,
  * outputs printed by the unmodified reference binary on dbgen SF1 (tests/golden/reference_answers.json).
"""
import json
import os

import numpy as np
import pytest

import cases as K
import oracle_tpch as OT
import tpch_data as D
from backends import OracleBackend
from quickstep_b200 import capi as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ANSWERS = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_answers.json")))


@pytest.fixture(scope="module")
def B(oracle):
    return OracleBackend()


def test_lip_test_golden(B):
    q1, total, words = K.case_lip_test(B)
    assert q1 == ANSWERS["lip_test"]["q1_x_mod_10000"]
    assert total == ANSWERS["lip_test"]["q2_sum_union"]
    # BitVectorExactFilter: bit (z - min) set for every z, MSB first
    bits = np.unpackbits(words.astype(">u8").view(np.uint8))
    assert bits[:100000:3].all() and bits.sum() == len(range(0, 100001, 3))


def test_select_test_groupby_golden(B):
    assert K.case_select_test_groupby(B) == ANSWERS["select_test_groupby"]["rows_count_g1_g2"]


def test_partition_test_join_golden(B):
    out = K.case_partition_test_join(B)
    assert sorted(i for i, _ in out) == sorted(ANSWERS["partition_test_join"]["ids"])
    chars = {i: c for i, c in out}
    assert chars[4] == b"4 2.000000" and chars[24] == b"24 4.898979" and chars[14] == b"14 3.741657"


FUNCS = {"sum": A.QS_AGG_SUM, "avg": A.QS_AGG_AVG, "min": A.QS_AGG_MIN, "max": A.QS_AGG_MAX, "count": A.QS_AGG_COUNT}


def expected_unittest(stem, func, is_expression, lo, hi, step=1):
    """Closed forms of AggregationOperator_unittest.cpp:572-602 over val in range(lo, hi, step)."""
    vals = np.arange(lo, hi, step, dtype=np.float64)
    scale = 0.1 if stem in ("FloatType", "DoubleType") else 1.0
    x = vals * scale
    e0, e1 = (2 * x, x * x) if is_expression else (x, x)
    n = len(vals)
    if func == "sum":
        return e0.sum(), e1.sum()
    if func == "avg":
        return e0.sum() / n, e1.sum() / n
    if func == "min":
        return e0.min(), e1.min()
    if func == "max":
        return e0.max(), e1.max()
    return n, n


@pytest.mark.parametrize("stem", ["IntType", "LongType", "FloatType", "DoubleType"])
@pytest.mark.parametrize("func", ["sum", "avg", "min", "max", "count"])
@pytest.mark.parametrize("is_expression", [False, True])
@pytest.mark.parametrize("with_predicate", [False, True])
def test_aggregation_unittest_closed_forms(B, stem, func, is_expression, with_predicate):
    out = K.case_agg_unittest(B, stem, FUNCS[func], is_expression, with_predicate, group_by=False)
    hi = 150 if with_predicate else 300
    e0, e1 = expected_unittest(stem, func, is_expression, 0, hi)
    assert out.n_groups == 1
    for got, exp in ((out.values[0][0], e0), (out.values[1][0], e1)):
        if stem in ("IntType", "LongType") and func in ("sum", "min", "max", "count"):
            assert int(got) == int(round(exp))
        else:
            # the reference's own tolerance (AggregationOperator_unittest.cpp:586-589)
            assert abs(float(got) - exp) <= abs(exp) * 1e-5 + 1e-12
    assert int(summation_check(hi)) == K.summation(hi - 1)


def summation_check(hi):
    return sum(range(hi))


@pytest.mark.parametrize("stem", ["IntType", "DoubleType"])
@pytest.mark.parametrize("func", ["sum", "avg", "count", "min", "max"])
@pytest.mark.parametrize("with_predicate", [False, True])
def test_aggregation_unittest_group_by(B, stem, func, with_predicate):
    """20 groups keyed (GroupBy-0, GroupBy-1); group id g holds val = g, g+20, ...
    (AggregationOperator_unittest.cpp:506-531 checkGroupByResult)."""
    out = K.case_agg_unittest(B, stem, FUNCS[func], False, with_predicate, group_by=True)
    assert out.n_groups == 20
    keys = out.keys.copy().view("<i4").reshape(-1, 2)
    hi = 150 if with_predicate else 300
    for i in range(20):
        g = int(keys[i][0]) + int(keys[i][1]) * K.K_GROUP1
        e0, _ = expected_unittest(stem, func, False, g, hi, K.K_GROUP_WIDTH)
        got = out.values[0][i]
        if stem == "IntType" and func != "avg":
            assert int(got) == int(round(e0))
        else:
            assert abs(float(got) - e0) <= abs(e0) * 1e-5 + 1e-12


@pytest.mark.parametrize("key", ["long", "int"])
def test_hash_join_unittest_match_counts(B, key):
    """HashJoinOperator_unittest.cpp:482-517: every fact tuple matches exactly one dim tuple;
    dim keys 0..99 are hit 3x each (300 fact rows over 100 keys), keys >= 100 never."""
    out = K.case_hash_join_unittest(B, key)
    assert out.n_rows == 300
    k = out.columns[0].data.astype(np.int64)
    counts = np.bincount(k, minlength=200)
    assert (counts[:100] == 3).all() and (counts[100:] == 0).all()
    for i in range(out.n_rows):
        assert out.columns[1].data[i] == str(int(k[i])).encode()


def test_join_test_left_outer_golden(B):
    """Join.test:137-165 (LEFT JOIN on INT and LONG keys): the oracle's outer join against the reference's table."""
    assert K.case_join_test_left_outer(B) == K.JOIN_TEST_LEFT_OUTER_EXPECTED


def test_hash_join_semi_anti(B):
    semi = K.case_hash_join_unittest(B, "int", A.QS_JOIN_LEFT_SEMI)
    anti = K.case_hash_join_unittest(B, "int", A.QS_JOIN_LEFT_ANTI)
    assert semi.n_rows == 300 and anti.n_rows == 0
    semi_r = K.case_hash_join_unittest(B, "int", A.QS_JOIN_LEFT_SEMI, residual=True)
    anti_r = K.case_hash_join_unittest(B, "int", A.QS_JOIN_LEFT_ANTI, residual=True)
    k = np.arange(300) % 100
    assert semi_r.n_rows == int((k % 3 != 0).sum()) and anti_r.n_rows == int((k % 3 == 0).sum())


def test_tpch_golden_sf001_shapes(golden):
    """dbgen SF0.01 fixture: row counts are dbgen's, Q1 has the four TPC-H groups."""
    assert golden["lineitem"].n_rows == 60175 and golden["orders"].n_rows == 15000 and golden["customer"].n_rows == 1500
    rows = OT.q1(golden["lineitem"])
    assert [r["l_returnflag"] + r["l_linestatus"] for r in rows] == [b"AF", b"NF", b"NO", b"RF"]
    d = golden["lineitem"].col("l_shipdate").data
    ymd = d["year"].astype(np.int64) * 10000 + d["month"].astype(np.int64) * 100 + d["day"].astype(np.int64)
    assert sum(r["count_order"] for r in rows) == int((ymd <= 19980901).sum())


@pytest.mark.skipif(not D.have_dbgen(), reason="oracle/_ref/dbgen not built (reference tree absent)")
def test_tpch_sf1_matches_reference_binary(oracle):
    """The oracle reproduces what the unmodified reference binary printed on dbgen SF1."""
    ref = ANSWERS["tpch_sf1_reference_binary"]
    tb = D.dbgen_tables(1)
    assert tb["lineitem"].n_rows == ref["lineitem_rows"]
    rev, is_null = OT.q6(tb["lineitem"])
    assert not is_null
    # Quickstep prints doubles with %.17g-like shortest round trip; block merge order may move the last ulps
    assert abs(rev - float(ref["q6_revenue_printed"])) <= 1e-9 * rev
    rows = OT.q1(tb["lineitem"])
    assert [(r["l_returnflag"] + r["l_linestatus"]).decode() for r in rows] == ref["q1_groups"]
    # TPC-H SF1 answer set values for the groups untouched by the 1998-09-01 vs -02 cutoff difference
    af = rows[0]
    assert af["count_order"] == 1478493 and af["sum_qty"] == 37734107.0
    assert abs(af["sum_base_price"] - 56586554400.73) < 0.01 and abs(af["sum_charge"] - 55909065222.83) < 0.01
    top = OT.q3(tb, D.q3_stats(tb))
    f = ref["q3_first_row"]
    assert top[0][0] == f["l_orderkey"] and abs(top[0][1] - f["revenue"]) < 1e-4
    assert "%04d-%02d-%02d" % top[0][2] == f["o_orderdate"] and top[0][3] == f["o_shippriority"]
    assert len(top) == 10


def test_oracle_matches_reference_engine_tables_sf001(golden):
    """Every cell of the result tables the UNMODIFIED reference engine printed for Q1 / Q3 / Q6 on dbgen -s 0.01
    (tests/golden/reference_engine_results.json) against the oracle over the committed dbgen columns."""
    import ref_golden as RG
    RG.check_q1(OT.q1(golden["lineitem"]), "sf0.01")
    rev, is_null = OT.q6(golden["lineitem"])
    RG.check_q6(rev, is_null, "sf0.01")
    RG.check_q3(OT.q3(golden, D.q3_stats(golden)), "sf0.01")


@pytest.mark.skipif(not D.have_dbgen(), reason="oracle/_ref/dbgen not built (reference tree absent)")
def test_oracle_matches_reference_engine_tables_sf1(oracle):
    """The same at SF1 (BASELINE.json configs[0]'s size): 4 x 10 cells of Q1, 10 x 4 of Q3, Q6."""
    import ref_golden as RG
    tb = D.dbgen_tables(1)
    RG.check_q1(OT.q1(tb["lineitem"]), "sf1")
    rev, is_null = OT.q6(tb["lineitem"])
    RG.check_q6(rev, is_null, "sf1")
    RG.check_q3(OT.q3(tb, D.q3_stats(tb)), "sf1")
