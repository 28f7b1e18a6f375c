"""Dictionary-coded attributes as a device format: every operator must give the
SAME answer over a relation whose attributes are resident as 1/2/4-byte codes as over the native columns --
i.e. the oracle's answer, and the reference's golden answers for TPC-H.

The reference evaluates comparisons with literals on the codes of a compressed stripe
(storage/CompressedTupleStorageSubBlock.cpp:160-251) and materialises values through the dictionary only where a
scalar needs them; the cases below are the parity cases of test_gpu_parity.py run through a backend that stages
base relations as blocks with per-block dictionaries (re-coded into one relation-wide dictionary on the device).
Bar: bit-exact row sets / per-row values; double SUM within 1e-9 relative."""
import numpy as np
import pytest

import cases as K
import oracle_tpch as OT
import tpch_data as D
import test_gpu_parity as P
from backends import CodedGpuBackend, OracleBackend, table_rows
from quickstep_b200 import capi as A
from quickstep_b200 import tpch as T
from quickstep_b200.expr import ExprSet
from quickstep_b200.table import Column, HostTable

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[1, 2, 4], ids=["cw-min", "cw>=2", "cw4"])
def GC(engine, request):
    b = CodedGpuBackend(engine, block_rows=997, min_cw=request.param)
    yield b
    b.close()
    engine.synchronize()


@pytest.fixture()
def GC1(engine):
    b = CodedGpuBackend(engine, block_rows=4001)
    yield b
    b.close()
    engine.synchronize()


@pytest.fixture(scope="module")
def OB(oracle):
    return OracleBackend()


# ------------------------------------------------------------------ the device format itself
def test_recode_round_trip(engine):
    """Blocks with per-block dictionaries -> relation codes -> qsgpu_relation_read gives back the native values
    of every type; the code column holds indices into the sorted relation-wide dictionary."""
    th = K.random_table(10007, seed=77)
    coded = {}
    for a, c in enumerate(th.columns):
        if c.data.dtype.kind == "f" and np.isnan(c.data).any():
            continue
        n = len(np.unique(c.data))
        coded[a] = 1 if n <= 256 else 2 if n <= 65536 else 4
    assert len(coded) >= 6
    rel = engine.Relation.from_host_coded(th, coded, block_rows=613)
    try:
        assert rel.n_rows == th.n_rows
        for a, c in enumerate(th.columns):
            got = rel.read(a)
            assert np.ascontiguousarray(got).view(np.uint8).tobytes() == np.ascontiguousarray(c.data).view(np.uint8).tobytes(), c.name
            cw, d = rel.dictionary(a)
            assert cw == coded.get(a, 0)
            if cw:
                assert (d == np.unique(c.data)).all()
        # a slice read decodes from the right offset
        a = next(iter(coded))
        assert (rel.read(a, 1234, 77) == th.columns[a].data[1234:1311]).all()
    finally:
        rel.destroy()


def test_dictionary_must_be_sorted_and_fit(engine):
    rel = engine.Relation.create([(A.QS_INT, 4)], 16)
    try:
        with pytest.raises(A.QsGpuError):
            rel.set_dictionary(0, 1, np.array([3, 2, 5], dtype=np.int32))          # not increasing
        with pytest.raises(A.QsGpuError):
            rel.set_dictionary(0, 1, np.arange(300, dtype=np.int32))               # 300 entries, 1-byte codes
        rel.set_dictionary(0, 1, np.array([2, 3, 5], dtype=np.int32))
        with pytest.raises(A.QsGpuError):
            rel.set_dictionary(0, 1, np.array([2, 3, 5], dtype=np.int32))          # already declared
    finally:
        rel.destroy()


def test_block_value_missing_from_dictionary_is_an_error(engine):
    th = HostTable("t", [Column("v", A.QS_INT, np.array([1, 2, 3, 9], dtype=np.int32))])
    rel = engine.Relation.create([(A.QS_INT, 4)], 16)
    try:
        rel.set_dictionary(0, 1, np.array([1, 2, 3], dtype=np.int32))
        d = np.array([1, 2, 3, 9], dtype=np.int32)
        mem = np.concatenate([d.view(np.uint8), np.array([0, 1, 2, 3], dtype=np.uint8), np.zeros(12, np.uint8)])
        with pytest.raises(A.QsGpuError):
            rel.stage_blocks([(mem, 4, [dict(attr=0, encoding=A.QS_ENC_DICT, offset=16, code_width=1, dict_offset=0,
                                             dict_entries=4)])])
        assert rel.n_rows == 0
    finally:
        rel.destroy()
    del th


# ------------------------------------------------------------------ comparisons on codes
@pytest.mark.parametrize("cmp", [A.QS_EQ, A.QS_NE, A.QS_LT, A.QS_LE, A.QS_GT, A.QS_GE])
def test_code_range_of_every_comparison(GC1, OB, cmp):
    """Literals below / inside / between / above the dictionary's values, on either side of the comparison and
    with a type promotion (INT attribute against a DOUBLE literal), plus CHAR and DATE attributes."""
    th = K.random_table(6000, seed=9)
    rel = GC1.relation(th)
    small = th.col("small").data
    lits = [int(small.min()) - 1, int(small.min()), int(np.median(small)), int(small.max()), int(small.max()) + 1]
    for lit in lits:
        for flipped in (False, True):
            for dbl in (False, True):
                es = ExprSet()
                l = es.lit_double(lit + 0.5) if dbl else es.lit_int(lit)
                a = th.attr(es, "small")
                p = es.cmp(cmp, l, a) if flipped else es.cmp(cmp, a, l)
                roots, schema = [th.attr(es, "i64")], [(A.QS_LONG, 8)]
                g = GC1.select(rel, es, p, None, roots, schema)
                o = OB.select(th, es, p, None, roots, schema)
                assert table_rows(g) == table_rows(o), (lit, flipped, dbl)
    # NaN literal: only != holds
    es = ExprSet()
    p = es.cmp(cmp, th.attr(es, "f64"), es.lit_double(float("nan")))
    roots, schema = [th.attr(es, "i64")], [(A.QS_LONG, 8)]
    assert table_rows(GC1.select(rel, es, p, None, roots, schema)) == table_rows(OB.select(th, es, p, None, roots, schema))
    # CHAR and DATE
    c4 = th.col("c4").data
    d = th.col("d").data
    for v in (c4[0], c4[len(c4) // 2], b"zzzz", b""):
        es = ExprSet()
        p = es.cmp(cmp, th.attr(es, "c4"), es.lit_char(bytes(v)))
        assert table_rows(GC1.select(rel, es, p, None, roots_of(es, th), schema)) == table_rows(OB.select(th, es, p, None, roots_of(es, th), schema))
    for v in (d[0], d[len(d) // 3]):
        es = ExprSet()
        p = es.cmp(cmp, th.attr(es, "d"), es.lit_date(int(v["year"]), int(v["month"]), int(v["day"])))
        assert table_rows(GC1.select(rel, es, p, None, roots_of(es, th), schema)) == table_rows(OB.select(th, es, p, None, roots_of(es, th), schema))
    assert GC1.n_coded >= 6


def roots_of(es, th):
    return [th.attr(es, "i64")]


# ------------------------------------------------------------------ the parity cases over coded relations
@pytest.mark.parametrize("n", [0, 1, 15, 16, 17, 1023, 1024, 1025, 4099, 70001])
def test_select_sizes_and_ragged_ranges(GC, OB, n):
    P.test_select_sizes_and_ragged_ranges(GC, OB, n)


@pytest.mark.parametrize("seed", range(12))
def test_random_predicates(GC1, OB, seed):
    P.test_random_predicates(GC1, OB, seed)


@pytest.mark.parametrize("seed", range(12))
def test_random_scalars_bit_exact(GC1, OB, seed):
    P.test_random_scalars_bit_exact(GC1, OB, seed)


def test_div_mod(GC, OB):
    P.test_div_mod(GC, OB)


@pytest.mark.parametrize("strategy", ["compact", "chaining", "collision_free"])
@pytest.mark.parametrize("n", [0, 1, 1000, 50000])
def test_group_by_strategies(GC, OB, strategy, n):
    P.test_group_by_strategies(GC, OB, strategy, n)


@pytest.mark.parametrize("stem", ["IntType", "DoubleType"])
@pytest.mark.parametrize("func", ["sum", "avg", "min", "max", "count"])
@pytest.mark.parametrize("with_predicate", [False, True])
@pytest.mark.parametrize("group_by", [False, True])
def test_aggregation_unittest_matrix(GC1, OB, stem, func, with_predicate, group_by):
    P.test_aggregation_unittest_matrix(GC1, OB, stem, func, True, with_predicate, group_by)


def test_lip_test_golden(GC1, OB):
    P.test_lip_test_golden(GC1, OB)


def test_select_test_groupby_golden(GC1):
    P.test_select_test_groupby_golden(GC1)


def test_lip_hash_filter_and_anti(GC1, OB):
    P.test_lip_hash_filter_and_anti(GC1, OB)


@pytest.mark.parametrize("key", ["long", "int"])
@pytest.mark.parametrize("join_type", [A.QS_JOIN_INNER, A.QS_JOIN_LEFT_SEMI, A.QS_JOIN_LEFT_ANTI])
@pytest.mark.parametrize("residual", [False, True])
def test_hash_join_unittest(GC1, OB, key, join_type, residual):
    P.test_hash_join_unittest(GC1, OB, key, join_type, residual, "open_addressing")


def test_left_outer_join_random(GC1, OB):
    P.test_left_outer_join_random(GC1, OB, "dense")


def test_join_duplicate_build_keys(GC1, OB):
    P.test_join_duplicate_build_keys(GC1, OB, "open_addressing")


# ------------------------------------------------------------------ TPC-H on a coded lineitem: reference's answers
LINEITEM_CODED = ["l_shipdate", "l_quantity", "l_discount", "l_tax", "l_returnflag", "l_linestatus", "l_extendedprice"]


def _coded_lineitem(engine, li: HostTable, block_rows=63000):
    coded = {}
    for a, c in enumerate(li.columns):
        if c.name in LINEITEM_CODED:
            n = len(np.unique(c.data))
            coded[a] = 1 if n <= 256 else 2 if n <= 65536 else 4
    return engine.Relation.from_host_coded(li, coded, block_rows=block_rows)


def _same_q1(rows, orows):
    assert len(rows) == len(orows)
    for r, o in zip(rows, orows):
        assert r["l_returnflag"] == o["l_returnflag"] and r["l_linestatus"] == o["l_linestatus"]
        assert r["count_order"] == o["count_order"]
        assert r["sum_qty"] == o["sum_qty"]            # integral doubles < 2^53: exact in any order
        for k in ("sum_base_price", "sum_disc_price", "sum_charge", "avg_qty", "avg_price", "avg_disc"):
            assert P.close(r[k], o[k]), (k, r[k], o[k])


def test_tpch_q1_q6_sf001_on_codes(engine, golden):
    """dbgen SF0.01 lineitem (the data the reference binary's golden answers belong to; the oracle is pinned to
    them in test_oracle_golden.py) staged as blocks with per-block dictionaries."""
    li = golden["lineitem"]
    rel = _coded_lineitem(engine, li, block_rows=6300)
    try:
        _same_q1(T.run_q1(rel), OT.q1(li))
        rev, is_null = T.run_q6(rel)
        orev, onull = OT.q6(li)
        assert is_null == onull and P.close(rev, orev)
    finally:
        rel.destroy()


def test_tpch_q3_sf001_on_codes(engine, golden):
    """Q3 with all three base relations coded where eligible: select + LIP probe + join build/probe + group-by
    over code tiles (keys and pass-through projections are decoded per tile in shared memory)."""
    stats = D.q3_stats(golden)
    be = CodedGpuBackend(engine, block_rows=5000)
    try:
        rels = {k: be.relation(v) for k, v in golden.items()}
        assert be.n_coded >= 8
        top = T.run_q3(rels["customer"], rels["orders"], rels["lineitem"], stats)
    finally:
        be.close()
    otop = OT.q3(golden, stats)
    assert len(top) == len(otop) == 10
    for g, o in zip(top, otop):
        assert g[0] == o[0] and g[2] == o[2] and g[3] == o[3]
        assert P.close(g[1], o[1])


def test_q1_q6_codes_equal_native_full_blocks(engine, oracle):
    """2 M synthetic lineitem rows in 63k-tuple blocks: Q1 / Q6 over codes == over native columns == oracle
    (counts exact, sums 1e-9); quantity / discount / tax take 1-byte codes, shipdate 2-byte codes."""
    arrays, _ = D.synthetic_lineitem_arrays(2_000_000, 11)
    li = D._table("lineitem", T.LINEITEM, arrays)
    nat = engine.Relation.from_host(li)
    rel = _coded_lineitem(engine, li)
    try:
        for a, c in enumerate(li.columns):
            if c.name in ("l_quantity", "l_discount", "l_tax", "l_returnflag", "l_linestatus"):
                assert rel.dictionary(a)[0] == 1
            if c.name == "l_shipdate":
                assert rel.dictionary(a)[0] == 2
        o = OT.q1(li)
        _same_q1(T.run_q1(rel), o)
        _same_q1(T.run_q1(nat), o)
        (rc, nc), (rn, nn), (ro, no) = T.run_q6(rel), T.run_q6(nat), OT.q6(li)
        assert nc == nn == no
        assert P.close(rc, rn) and P.close(rc, ro)
    finally:
        nat.destroy()
        rel.destroy()
