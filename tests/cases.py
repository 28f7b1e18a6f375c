"""Parity cases, written once against the backends API (tests/backends.py).

Sources of the cases:
  * the reference's SQL golden tests (literal expected tables):
      query_optimizer/tests/execution_generator/LIP.test:19-151
      query_optimizer/tests/execution_generator/Select.test:659-683
      query_optimizer/tests/execution_generator/Partition.test:57-75
  * the reference's operator unit tests (closed-form expectations):
      relational_operators/tests/AggregationOperator_unittest.cpp:93-106,192-206,572-602
      relational_operators/tests/HashJoinOperator_unittest.cpp:97-99,172-194,482-517
  * seeded random tables / expression trees for everything else.
"""
from __future__ import annotations

import math

import numpy as np

from quickstep_b200 import capi as A
from quickstep_b200.expr import ExprSet
from quickstep_b200.table import Column, HostTable, days_to_dates


# ------------------------------------------------------------------ LIP.test
def lip_test_tables():
    r = np.arange(0, 100001, 2, dtype=np.int32)
    s = np.arange(0, 100001, 3, dtype=np.int32)
    R = HostTable("R", [Column("x", A.QS_INT, r), Column("y", A.QS_INT, r.copy())])
    S = HostTable("S", [Column("z", A.QS_INT, s)])
    return R, S


def case_lip_test(B):
    """FilterJoin plan: BuildLIPFilter(S.z exact filter) -> Select(R, x % m = 0, LIP probe y)."""
    Rh, Sh = lip_test_tables()
    R, S = B.relation(Rh), B.relation(Sh)
    f = B.make_lip(A.QS_LIP_BITVECTOR_EXACT, A.QS_INT, int(Sh.col("z").data.min()), int(Sh.col("z").data.max()))
    B.build_lip(S, None, -1, None, [(f, 0)])

    def sel(m):
        es = ExprSet()
        p = es.cmp(A.QS_EQ, es.mod(es.attr(0, A.QS_INT), es.lit_int(m)), es.lit_int(0))
        return B.select(R, es, p, [(f, 1)], [es.attr(0, A.QS_INT)], [(A.QS_INT, 4)])

    q1 = sorted(int(v) for v in sel(10000).columns[0].data)
    # UNION ALL of two selections feeding SUM(x)
    total = 0
    for m in (5, 7):
        t = sel(m)
        es = ExprSet()
        out = B.aggregate(B.relation(t), es, -1, [(A.QS_AGG_SUM, es.attr(0, A.QS_INT))], [], A.QS_AGG_SINGLE_STATE, [])
        total += int(out.values[0][0])
    return q1, total, B.lip_words(f)


# -------------------------------------------------------- TestDatabaseLoader
def test_relation(drop_null_rows=True):
    """query_optimizer/tests/TestDatabaseLoader.cpp:141-183.  Rows with x % 10 == 0 have NULL
    int_col/double_col; nullable columns are not staged on the device, so the cases below use
    queries whose predicates reject those rows anyway and drop them up front."""
    xs = [x for x in range(25) if not (drop_null_rows and x % 10 == 0)]
    sign = [1 if x % 2 == 0 else -1 for x in xs]
    int_col = np.array([s * x for s, x in zip(sign, xs)], dtype=np.int32)
    long_col = np.array([x * x for x in xs], dtype=np.int64)
    float_col = np.array([math.sqrt(x) for x in xs], dtype=np.float32)
    double_col = np.array([s * math.sqrt(x) * x for s, x in zip(sign, xs)], dtype=np.float64)
    chars = []
    for s, x in zip(sign, xs):
        v = f"{s * x} {math.sqrt(x):.6f}"
        chars.append(v[:19].encode())
    char_col = np.array(chars, dtype="S20")
    return HostTable("test", [Column("int_col", A.QS_INT, int_col), Column("long_col", A.QS_LONG, long_col),
                              Column("float_col", A.QS_FLOAT, float_col), Column("double_col", A.QS_DOUBLE, double_col),
                              Column("char_col", A.QS_CHAR, char_col, 20)])


def case_select_test_groupby(B):
    """Select.test:659-683: COUNT(*) GROUP BY long_col/100, long_col/50 HAVING COUNT(*)>0 AND g2>5.
    Group-by expressions are projected by a Select first, then aggregated on attributes."""
    Th = test_relation(drop_null_rows=False)   # long_col is not nullable; all 25 rows count
    T = B.relation(Th.project(["long_col"]))
    es = ExprSet()
    g1 = es.div(es.attr(0, A.QS_LONG), es.lit_int(100))
    g2 = es.div(es.attr(0, A.QS_LONG), es.lit_int(50))
    t = B.select(T, es, -1, None, [g1, g2], [(A.QS_LONG, 8), (A.QS_LONG, 8)])
    e2 = ExprSet()
    having = e2.cmp(A.QS_GT, e2.attr(1, A.QS_LONG), e2.lit_int(5))
    out = B.aggregate(B.relation(t), e2, having, [(A.QS_AGG_COUNT, -1)], [e2.attr(0, A.QS_LONG), e2.attr(1, A.QS_LONG)],
                      A.QS_AGG_SEPARATE_CHAINING, [(A.QS_LONG, 8), (A.QS_LONG, 8)])
    keys = out.keys.copy().view("<i8").reshape(-1, 2)
    rows = sorted([int(out.values[0][i]), int(keys[i][0]), int(keys[i][1])] for i in range(out.n_groups))
    return sorted(rows, key=lambda r: (r[1], r[2]))


def case_partition_test_join(B):
    """Partition.test:57-75: fact(id=int_col even, score) JOIN dim(id=int_col non-null, char_col)."""
    Th = test_relation()
    ic = Th.col("int_col").data
    fact_rows = ic % 2 == 0
    fact = HostTable("fact", [Column("id", A.QS_INT, ic[fact_rows]),
                              Column("score", A.QS_DOUBLE, Th.col("double_col").data[fact_rows])])
    dim = HostTable("dim", [Column("id", A.QS_INT, ic), Column("char_col", A.QS_CHAR, Th.col("char_col").data, 20)])
    es = ExprSet()
    roots = [es.attr(0, A.QS_INT), es.attr(1, A.QS_CHAR, 20, 2)]
    out = B.hash_join(B.relation(dim), -1, 0, B.relation(fact), es, -1, 0, A.QS_JOIN_INNER, -1, roots,
                      [(A.QS_INT, 4), (A.QS_CHAR, 20)], 64)
    return sorted((int(out.columns[0].data[i]), bytes(out.columns[1].data[i])) for i in range(out.n_rows))


def join_test_tables():
    """Join.test:19-57: a(w INT, x LONG, y DOUBLE), b from a WHERE w % 2 = 0, c from a WHERE x % 3 = 0."""
    w = np.arange(20, dtype=np.int32)
    x = (w * 10).astype(np.int64)
    y = (w * 100).astype(np.float64)
    a = HostTable("a", [Column("w", A.QS_INT, w), Column("x", A.QS_LONG, x), Column("y", A.QS_DOUBLE, y)])
    mb = w % 2 == 0
    b = HostTable("b", [Column("w", A.QS_INT, w[mb]), Column("x", A.QS_LONG, x[mb] + (w[mb] // 2) % 2)])
    mc = x % 3 == 0
    c = HostTable("c", [Column("x", A.QS_LONG, x[mc]), Column("y", A.QS_DOUBLE, y[mc] + (x[mc] // 3) % 3 - 1)])
    return a, b, c


# Join.test:137-165, columns a.w, b.x, c.y of `a LEFT JOIN b ON a.w = b.w LEFT JOIN c ON a.x = c.x`
JOIN_TEST_LEFT_OUTER_EXPECTED = [
    (0, 0, -1.0), (1, None, None), (2, 21, None), (3, None, 300.0), (4, 40, None), (5, None, None), (6, 61, 601.0),
    (7, None, None), (8, 80, None), (9, None, 899.0), (10, 101, None), (11, None, None), (12, 120, 1200.0),
    (13, None, None), (14, 141, None), (15, None, 1501.0), (16, 160, None), (17, None, None), (18, 181, 1799.0),
    (19, None, None)]


def case_join_test_left_outer(B):
    """The two INT/LONG-keyed LEFT JOINs of Join.test:137-165, each as one HashOuterJoin work order; returns
    [(a.w, b.x | None, c.y | None)] ordered by a.w."""
    a, b, c = join_test_tables()
    es = ExprSet()
    out1 = B.hash_join(B.relation(b), -1, 0, B.relation(a), es, -1, 0, A.QS_JOIN_LEFT_OUTER, -1,
                       [es.attr(0, A.QS_INT), es.attr(1, A.QS_LONG, 8, 2)], [(A.QS_INT, 4), (A.QS_LONG, 8)], 64)
    es2 = ExprSet()
    out2 = B.hash_join(B.relation(c), -1, 0, B.relation(a), es2, -1, 1, A.QS_JOIN_LEFT_OUTER, -1,
                       [es2.attr(0, A.QS_INT), es2.attr(1, A.QS_DOUBLE, 8, 2)], [(A.QS_INT, 4), (A.QS_DOUBLE, 8)], 64)
    bx = {int(out1.columns[0].data[i]): (None if int(out1.nulls[i]) & 2 else int(out1.columns[1].data[i])) for i in range(out1.n_rows)}
    cy = {int(out2.columns[0].data[i]): (None if int(out2.nulls[i]) & 2 else float(out2.columns[1].data[i])) for i in range(out2.n_rows)}
    assert out1.n_rows == out2.n_rows == 20 and not any(int(n) & 1 for n in out1.nulls)
    return [(w, bx[w], cy[w]) for w in sorted(bx)]


# ------------------------------------------- AggregationOperator_unittest.cpp
K_NUM_TUPLES, K_GROUP_WIDTH, K_GROUP1 = 300, 20, 4


def agg_unittest_table(n=K_NUM_TUPLES):
    """createTuple, AggregationOperator_unittest.cpp:192-206."""
    val = np.arange(n, dtype=np.int64)
    gid = val % K_GROUP_WIDTH
    cols = [Column("GroupBy-0", A.QS_INT, (gid % K_GROUP1).astype(np.int32)),
            Column("GroupBy-1", A.QS_INT, (gid // K_GROUP1).astype(np.int32))]
    for stem, t, arr in (("IntType", A.QS_INT, val.astype(np.int32)), ("LongType", A.QS_LONG, val),
                         ("FloatType", A.QS_FLOAT, (0.1 * val).astype(np.float32)),
                         ("DoubleType", A.QS_DOUBLE, 0.1 * val)):
        cols.append(Column(stem + "-0", t, arr))
        cols.append(Column(stem + "-1", t, arr.copy()))
    return HostTable("table", cols)


def case_agg_unittest(B, stem, func, is_expression, with_predicate, group_by, predicate_value=150, n=K_NUM_TUPLES,
                      block_ranges=None):
    """setupTest / setupTestGroupBy (…unittest.cpp:236-330): two aggregates per state --
    attribute form agg(X-0), agg(X-1); expression form agg(X-0 + X-1), agg(X-0 * X-1);
    optional predicate IntType-0 < predicate_value; optional GROUP BY (GroupBy-0, GroupBy-1)."""
    Th = agg_unittest_table(n)
    T = B.relation(Th)
    es = ExprSet()
    a0, a1 = Th.attr(es, stem + "-0"), Th.attr(es, stem + "-1")
    if func == A.QS_AGG_COUNT and not is_expression:
        aggs = [(func, a0), (func, -1)]            # COUNT(attr) and COUNT(*)
    elif is_expression:
        aggs = [(func, es.add(a0, a1)), (func, es.mul(Th.attr(es, stem + "-0"), Th.attr(es, stem + "-1")))]
    else:
        aggs = [(func, a0), (func, a1)]
    pred = es.cmp(A.QS_LT, Th.attr(es, "IntType-0"), es.lit_int(predicate_value)) if with_predicate else -1
    if group_by:
        groups = [Th.attr(es, "GroupBy-0"), Th.attr(es, "GroupBy-1")]
        return B.aggregate(T, es, pred, aggs, groups, A.QS_AGG_COMPACT_KEY, [(A.QS_INT, 4), (A.QS_INT, 4)],
                           row_ranges=block_ranges)
    return B.aggregate(T, es, pred, aggs, [], A.QS_AGG_SINGLE_STATE, [], row_ranges=block_ranges)


def summation(n):
    return (n + 1) * n // 2


def summation_squares(n):
    return n * (n + 1) * (2 * n + 1) // 6


# --------------------------------------------- HashJoinOperator_unittest.cpp
def hash_join_unittest_tables(dim_rows=200, fact_rows=300):
    """HashJoinOperator_unittest.cpp:172-230: dim(long, int, char, varchar) x 200 with
    long = int = i; fact x 300 with long = int = i % (dim/2) -> every fact row matches exactly
    one dim row with key < 100; dim rows >= 100 match nothing."""
    di = np.arange(dim_rows)
    fi = np.arange(fact_rows) % (dim_rows // 2)
    dim = HostTable("dim", [Column("long", A.QS_LONG, di.astype(np.int64)), Column("int", A.QS_INT, di.astype(np.int32)),
                            Column("char", A.QS_CHAR, np.array([f"{i}".encode() for i in di], dtype="S12"), 12)])
    fact = HostTable("fact", [Column("long", A.QS_LONG, fi.astype(np.int64)), Column("int", A.QS_INT, fi.astype(np.int32)),
                              Column("char", A.QS_CHAR, np.array([f"{i}".encode() for i in fi], dtype="S12"), 12)])
    return dim, fact


def case_hash_join_unittest(B, key="long", join_type=A.QS_JOIN_INNER, residual=False):
    """LongKeyCartesianProductHashJoinTest / IntDuplicateKeyHashJoinTest shape: build on dim,
    probe fact, project (fact.key, dim.char); per-key match counts are checked by the caller."""
    dimh, facth = hash_join_unittest_tables()
    dim, fact = B.relation(dimh), B.relation(facth)
    ka = dimh.attr_id(key)
    kt = dimh.col(key).type
    es = ExprSet()
    res = -1
    if residual:   # fact.int < dim.int + 1  AND dim.int % 3 != 0  (pair predicate over both sides)
        res = es.and_(es.cmp(A.QS_LT, es.attr(1, A.QS_INT), es.add(es.attr(1, A.QS_INT, 4, 2), es.lit_int(1))),
                      es.cmp(A.QS_NE, es.mod(es.attr(1, A.QS_INT, 4, 2), es.lit_int(3)), es.lit_int(0)))
    if join_type == A.QS_JOIN_INNER:
        roots = [es.attr(ka, kt), es.attr(2, A.QS_CHAR, 12, 2)]
        schema = [(kt, 0), (A.QS_CHAR, 12)]
    else:
        roots = [es.attr(ka, kt), es.attr(2, A.QS_CHAR, 12)]
        schema = [(kt, 0), (A.QS_CHAR, 12)]
    schema = [(t, w or (8 if t == A.QS_LONG else 4)) for t, w in schema]
    return B.hash_join(dim, -1, ka, fact, es, -1, ka, join_type, res, roots, schema, 4096)


# -------------------------------------------------------------- random data
def random_table(n, seed, with_dup_keys=True):
    rng = np.random.default_rng(seed)
    cols = [Column("i32", A.QS_INT, rng.integers(-1000, 1000, size=n).astype(np.int32)),
            Column("i64", A.QS_LONG, rng.integers(-10**12, 10**12, size=n)),
            Column("f32", A.QS_FLOAT, rng.normal(0, 100, size=n).astype(np.float32)),
            Column("f64", A.QS_DOUBLE, rng.normal(0, 1e4, size=n)),
            Column("d", A.QS_DATE, days_to_dates(rng.integers(8000, 11000, size=n))),
            Column("c4", A.QS_CHAR, np.array([b"ab", b"abc", b"abcd", b"b", b"", b"zz"], dtype="S4")[rng.integers(0, 6, size=n)], 4),
            Column("k", A.QS_INT, rng.integers(0, max(2, n // 3), size=n).astype(np.int32)),
            Column("g", A.QS_CHAR, np.array([b"A", b"N", b"R"], dtype="S1")[rng.integers(0, 3, size=n)], 1),
            Column("small", A.QS_INT, rng.integers(0, 7, size=n).astype(np.int32)),
            Column("pos64", A.QS_LONG, rng.integers(0, 5000, size=n))]
    return HostTable("rnd", cols)


def random_predicate(es, t: HostTable, rng, depth=0):
    kind = rng.integers(0, 10)
    if depth < 3 and kind < 3:
        return es.and_(random_predicate(es, t, rng, depth + 1), random_predicate(es, t, rng, depth + 1))
    if depth < 3 and kind < 5:
        return es.or_(random_predicate(es, t, rng, depth + 1), random_predicate(es, t, rng, depth + 1))
    if depth < 3 and kind < 6:
        return es.not_(random_predicate(es, t, rng, depth + 1))
    op = int(rng.integers(0, 6))
    which = rng.integers(0, 8)
    if which == 0:
        return es.cmp(op, t.attr(es, "i32"), es.lit_int(int(rng.integers(-1000, 1000))))
    if which == 1:
        return es.cmp(op, t.attr(es, "i64"), es.lit_long(int(rng.integers(-10**12, 10**12))))
    if which == 2:
        return es.cmp(op, t.attr(es, "f32"), es.lit_double(float(rng.normal(0, 100))))
    if which == 3:
        return es.cmp(op, t.attr(es, "f64"), es.lit_int(int(rng.integers(-10000, 10000))))
    if which == 4:
        d = days_to_dates(rng.integers(8000, 11000, size=1))[0]
        return es.cmp(op, t.attr(es, "d"), es.lit_date(int(d["year"]), int(d["month"]), int(d["day"])))
    if which == 5:
        lit = [b"ab", b"abc", b"abcd", b"b", b"a", b"zz"][int(rng.integers(0, 6))]
        return es.cmp(op, t.attr(es, "c4"), es.lit_char(lit))
    if which == 6:   # attribute vs attribute with promotion
        return es.cmp(op, t.attr(es, "i32"), t.attr(es, "f32"))
    return es.cmp(op, es.add(t.attr(es, "i32"), t.attr(es, "small")), es.mul(t.attr(es, "small"), es.lit_int(100)))


def random_scalar(es, t: HostTable, rng, depth=0):
    """Returns (root, type_id)."""
    leaves = [("i32", A.QS_INT), ("i64", A.QS_LONG), ("f32", A.QS_FLOAT), ("f64", A.QS_DOUBLE), ("small", A.QS_INT)]
    kind = rng.integers(0, 10)
    if depth >= 3 or kind < 3:
        if rng.integers(0, 4) == 0:
            c = int(rng.integers(0, 4))
            if c == 0:
                return es.lit_int(int(rng.integers(-50, 50))), A.QS_INT
            if c == 1:
                return es.lit_long(int(rng.integers(-10**6, 10**6))), A.QS_LONG
            if c == 2:
                return es.lit_float(float(np.float32(rng.normal()))), A.QS_FLOAT
            return es.lit_double(float(rng.normal())), A.QS_DOUBLE
        n, ty = leaves[int(rng.integers(0, len(leaves)))]
        return t.attr(es, n), ty
    if kind == 3:
        a, ta = random_scalar(es, t, rng, depth + 1)
        return es.neg(a), ta
    if kind == 4:
        a, ta = random_scalar(es, t, rng, depth + 1)
        # fp -> integer casts of out-of-range values are UB in C++ (and differ by ISA): only widen
        opts = [A.QS_FLOAT, A.QS_DOUBLE] + ([A.QS_LONG] if ta in (A.QS_INT, A.QS_LONG) else [])
        to = opts[int(rng.integers(0, len(opts)))]
        return es.cast(a, to), to
    a, ta = random_scalar(es, t, rng, depth + 1)
    b, tb = random_scalar(es, t, rng, depth + 1)
    op = int(rng.integers(0, 3))   # + - *  (DIV/MOD are exercised separately with safe operands)
    if ta == tb:
        ty = ta
    elif A.QS_DOUBLE in (ta, tb) or {ta, tb} == {A.QS_LONG, A.QS_FLOAT}:
        ty = A.QS_DOUBLE
    elif A.QS_FLOAT in (ta, tb):
        ty = A.QS_FLOAT
    else:
        ty = A.QS_LONG
    return es.binary(op, a, b), ty


def width_of(ty):
    return 4 if ty in (A.QS_INT, A.QS_FLOAT) else 8
