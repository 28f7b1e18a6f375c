"""TPC-H Q1 / Q6 / Q3 as the operator DAGs the reference's (unchanged) optimizer
produces for them ("Executed plans"), expressed over the
C-ABI work-order calls.  Plan *descriptions* (expression trees, aggregate
lists) are shared with the CPU oracle so both sides evaluate the same trees.

Schema subset = the attributes the three queries touch
(benchmarks/tpch/create.sql:18-114; DECIMAL is DOUBLE, parser/SqlParser.ypp:791).
"""
from __future__ import annotations

import numpy as np

from . import capi as A
from . import engine as E
from .expr import ExprSet

LINEITEM = [("l_orderkey", A.QS_INT, 4), ("l_quantity", A.QS_DOUBLE, 8), ("l_extendedprice", A.QS_DOUBLE, 8),
            ("l_discount", A.QS_DOUBLE, 8), ("l_tax", A.QS_DOUBLE, 8), ("l_returnflag", A.QS_CHAR, 1),
            ("l_linestatus", A.QS_CHAR, 1), ("l_shipdate", A.QS_DATE, 8)]
ORDERS = [("o_orderkey", A.QS_INT, 4), ("o_custkey", A.QS_INT, 4), ("o_orderdate", A.QS_DATE, 8),
          ("o_shippriority", A.QS_INT, 4)]
CUSTOMER = [("c_custkey", A.QS_INT, 4), ("c_mktsegment", A.QS_CHAR, 10)]

# native bytes per row each query reads
Q6_BYTES_PER_ROW = 32
Q1_BYTES_PER_ROW = 42
Q3_LINEITEM_BYTES_PER_ROW = 28
Q3_ORDERS_BYTES_PER_ROW = 20
Q3_CUSTOMER_BYTES_PER_ROW = 14


def _idx(schema, name):
    for i, (n, _t, _w) in enumerate(schema):
        if n == name:
            return i
    raise KeyError(name)


def _attr(es, schema, name, side=0):
    i = _idx(schema, name)
    return es.attr(i, schema[i][1], schema[i][2], side)


# ----------------------------------------------------------------------- Q6
class Q6Plan:
    """lineitem -> Aggregate(single state) -> Finalize.
    Predicate (queries/06.sql): l_shipdate >= 1994-01-01 AND l_shipdate <
    1995-01-01 (the interval addition is constant-folded,
    ScalarBinaryExpression.hpp:91-97) AND l_discount BETWEEN 0.05 AND 0.07 AND
    l_quantity < 24 (INT literal, promoted to DOUBLE by the comparison)."""

    def __init__(self, schema=LINEITEM):
        es = ExprSet()
        sd = lambda: _attr(es, schema, "l_shipdate")
        p = es.and_(
            es.cmp(A.QS_GE, sd(), es.lit_date(1994, 1, 1)),
            es.cmp(A.QS_LT, sd(), es.lit_date(1995, 1, 1)),
            es.cmp(A.QS_GE, _attr(es, schema, "l_discount"), es.lit_double(0.05)),
            es.cmp(A.QS_LE, _attr(es, schema, "l_discount"), es.lit_double(0.07)),
            es.cmp(A.QS_LT, _attr(es, schema, "l_quantity"), es.lit_int(24)))
        rev = es.mul(_attr(es, schema, "l_extendedprice"), _attr(es, schema, "l_discount"))
        self.es, self.pred = es, p
        self.aggregates = [(A.QS_AGG_SUM, rev)]
        self.group_by = []
        self.strategy = A.QS_AGG_SINGLE_STATE


# ----------------------------------------------------------------------- Q1
class Q1Plan:
    """lineitem -> Aggregate(thread-private compact key) -> Selection -> Sort.
    After ReuseAggregateExpressions the aggregate list is SUM(qty), SUM(price),
    SUM(disc_price), SUM(charge), SUM(discount), COUNT(*); the AVGs are
    SUM/COUNT divisions in the wrapping Selection
    (rules/ReuseAggregateExpressions.cpp:59-67,229-246).  disc_price is a
    ScalarSharedExpression (ExtractCommonSubexpression)."""

    def __init__(self, schema=LINEITEM):
        es = ExprSet()
        a = lambda n: _attr(es, schema, n)
        p = es.cmp(A.QS_LE, a("l_shipdate"), es.lit_date(1998, 9, 1))
        disc_price = es.shared(es.mul(a("l_extendedprice"), es.sub(es.lit_int(1), a("l_discount"))), 0)
        charge = es.mul(disc_price, es.add(es.lit_int(1), a("l_tax")))
        self.es, self.pred = es, p
        self.aggregates = [(A.QS_AGG_SUM, a("l_quantity")), (A.QS_AGG_SUM, a("l_extendedprice")),
                           (A.QS_AGG_SUM, disc_price), (A.QS_AGG_SUM, charge),
                           (A.QS_AGG_SUM, a("l_discount")), (A.QS_AGG_COUNT, -1)]
        self.group_by = [a("l_returnflag"), a("l_linestatus")]
        self.key_schema = [(A.QS_CHAR, 1), (A.QS_CHAR, 1)]
        self.strategy = A.QS_AGG_COMPACT_KEY


def q1_rows_from_states(keys_flag, keys_status, sums, counts):
    """The wrapping Selection + Sort of Q1: AVG = SUM / COUNT (double / long ->
    double), ordered by (l_returnflag, l_linestatus).  Host-side, 4 rows."""
    order = np.lexsort((keys_status, keys_flag))
    rows = []
    for i in order:
        c = int(counts[i])
        rows.append(dict(l_returnflag=bytes(keys_flag[i]), l_linestatus=bytes(keys_status[i]),
                         sum_qty=float(sums[0][i]), sum_base_price=float(sums[1][i]),
                         sum_disc_price=float(sums[2][i]), sum_charge=float(sums[3][i]),
                         avg_qty=float(sums[0][i]) / float(c), avg_price=float(sums[1][i]) / float(c),
                         avg_disc=float(sums[4][i]) / float(c), count_order=c))
    return rows


def merge_q1_partitions(parts):
    """Merge the Q1 result rows of lineitem partitions (one list per GPU; rows carry sum_disc): SUMs and
    COUNT add, the AVGs are recomputed from the merged SUM / COUNT.  Keyed by (flag, status): the same
    group may sit at a different position in each partition's result."""
    acc = {}
    for rows in parts:
        for r in rows:
            k = (r["l_returnflag"], r["l_linestatus"])
            a = acc.setdefault(k, dict(l_returnflag=k[0], l_linestatus=k[1], sum_qty=0.0, sum_base_price=0.0,
                                       sum_disc_price=0.0, sum_charge=0.0, sum_disc=0.0, count_order=0))
            for f in ("sum_qty", "sum_base_price", "sum_disc_price", "sum_charge", "sum_disc"):
                a[f] += r[f]
            a["count_order"] += int(r["count_order"])
    out = []
    for k in sorted(acc):
        a = acc[k]
        c = float(a["count_order"])
        a.update(avg_qty=a["sum_qty"] / c, avg_price=a["sum_base_price"] / c, avg_disc=a["sum_disc"] / c)
        out.append(a)
    return out


# ----------------------------------------------------------------------- Q3
class Q3Plan:
    """[1] BuildLIPFilter(customer, c_mktsegment='BUILDING' -> exact filter on c_custkey)
       [2] Select(orders, o_orderdate < 1995-03-15, LIP probe o_custkey) -> T2(o_orderkey,o_orderdate,o_shippriority)
       [3] BuildHash(T2 on o_orderkey) + LIP build exact filter on o_orderkey
       [0] Select(lineitem, l_shipdate > 1995-03-15, LIP probe l_orderkey) -> T0(l_orderkey,l_extendedprice,l_discount)
       [4] HashJoin(probe T0, l_orderkey = o_orderkey) -> T4(l_orderkey,o_orderdate,o_shippriority,l_extendedprice,l_discount)
       [6] Aggregate(T4 group by (l_orderkey,o_orderdate,o_shippriority), SUM(price*(1-discount)))  separate chaining
       [7] Finalize  [9,10] Sort revenue desc, o_orderdate  LIMIT 10."""

    def __init__(self):
        C_, O_, L_ = CUSTOMER, ORDERS, LINEITEM
        # [1]
        e1 = ExprSet()
        self.e1, self.p1 = e1, e1.cmp(A.QS_EQ, _attr(e1, C_, "c_mktsegment"), e1.lit_char(b"BUILDING"))
        self.c_custkey = _idx(C_, "c_custkey")
        # [2]
        e2 = ExprSet()
        self.e2 = e2
        self.p2 = e2.cmp(A.QS_LT, _attr(e2, O_, "o_orderdate"), e2.lit_date(1995, 3, 15))
        self.proj2 = [_attr(e2, O_, "o_orderkey"), _attr(e2, O_, "o_orderdate"), _attr(e2, O_, "o_shippriority")]
        self.t2_schema = [(A.QS_INT, 4), (A.QS_DATE, 8), (A.QS_INT, 4)]
        self.o_custkey = _idx(O_, "o_custkey")
        # [0]
        e0 = ExprSet()
        self.e0 = e0
        self.p0 = e0.cmp(A.QS_GT, _attr(e0, L_, "l_shipdate"), e0.lit_date(1995, 3, 15))
        self.proj0 = [_attr(e0, L_, "l_orderkey"), _attr(e0, L_, "l_extendedprice"), _attr(e0, L_, "l_discount")]
        self.t0_schema = [(A.QS_INT, 4), (A.QS_DOUBLE, 8), (A.QS_DOUBLE, 8)]
        self.l_orderkey = _idx(L_, "l_orderkey")
        # [4] probe T0 (attrs 0..2), build T2 (attrs 0..2, side=2)
        e4 = ExprSet()
        self.e4 = e4
        self.proj4 = [e4.attr(0, A.QS_INT, 4), e4.attr(1, A.QS_DATE, 8, 2), e4.attr(2, A.QS_INT, 4, 2),
                      e4.attr(1, A.QS_DOUBLE, 8), e4.attr(2, A.QS_DOUBLE, 8)]
        self.t4_schema = [(A.QS_INT, 4), (A.QS_DATE, 8), (A.QS_INT, 4), (A.QS_DOUBLE, 8), (A.QS_DOUBLE, 8)]
        # [6]
        e6 = ExprSet()
        self.e6 = e6
        rev = e6.mul(e6.attr(3, A.QS_DOUBLE, 8), e6.sub(e6.lit_int(1), e6.attr(4, A.QS_DOUBLE, 8)))
        self.aggregates = [(A.QS_AGG_SUM, rev)]
        self.group_by = [e6.attr(0, A.QS_INT, 4), e6.attr(1, A.QS_DATE, 8), e6.attr(2, A.QS_INT, 4)]
        self.key_schema = [(A.QS_INT, 4), (A.QS_DATE, 8), (A.QS_INT, 4)]
        # [9,10] on the finalize output (l_orderkey, o_orderdate, o_shippriority, revenue)
        self.sort_keys = [(3, True), (1, False)]
        self.limit = 10


# ------------------------------------------------------------ device runners
def run_q6(lineitem: E.Relation, plan: Q6Plan | None = None, row_ranges=None):
    """Returns (revenue, is_null)."""
    plan = plan or Q6Plan()
    st = E.AggState(plan.strategy, plan.es, plan.pred, plan.aggregates, plan.group_by, dev=lineitem.dev)
    try:
        for lo, hi in (row_ranges or [(0, A.UINT64_MAX)]):
            st.run(lineitem, lo, hi)
        rel, mask = E.finalize_relation(st, [], [(A.QS_DOUBLE, 8)])
        v = rel.read(0)
        rel.destroy()
        return float(v[0]), bool(mask & 1)
    finally:
        st.destroy()


def run_q1(lineitem: E.Relation, plan: Q1Plan | None = None, row_ranges=None):
    """Returns the 4 (or fewer) result rows as dicts, ordered."""
    plan = plan or Q1Plan()
    st = E.AggState(plan.strategy, plan.es, plan.pred, plan.aggregates, plan.group_by, estimated=8,
                    dev=lineitem.dev)
    try:
        for lo, hi in (row_ranges or [(0, A.UINT64_MAX)]):
            st.run(lineitem, lo, hi)
        out_types = [(A.QS_DOUBLE, 8)] * 5 + [(A.QS_LONG, 8)]
        rel, _ = E.finalize_relation(st, plan.key_schema, out_types)
        cols = rel.read_all()
        rel.destroy()
        return q1_rows_from_states(cols[0], cols[1], cols[2:7], cols[7])
    finally:
        st.destroy()


def run_q3(customer: E.Relation, orders: E.Relation, lineitem: E.Relation, stats: dict,
           plan: Q3Plan | None = None, timings: dict | None = None):
    """stats: exact min/max of c_custkey and o_orderkey (what \\analyze records
    and InjectJoinFilters / AttachLIPFilters read).  Returns the top-10 rows:
    list of (l_orderkey, revenue, (y,m,d), o_shippriority)."""
    plan = plan or Q3Plan()
    dev = lineitem.dev
    objs = []
    try:
        f_cust = E.LipFilter(A.QS_LIP_BITVECTOR_EXACT, A.QS_INT, stats["c_custkey_min"], stats["c_custkey_max"], dev=dev)
        f_ord = E.LipFilter(A.QS_LIP_BITVECTOR_EXACT, A.QS_INT, stats["o_orderkey_min"], stats["o_orderkey_max"], dev=dev)
        objs += [f_cust, f_ord]
        E.build_lip_filter(customer, plan.e1, plan.p1, None, [(f_cust, plan.c_custkey)])
        t2 = E.Relation.create(plan.t2_schema, max(1, stats["orders_rows"]), dev=dev)
        objs.append(t2)
        E.select(orders, plan.e2, plan.p2, [(f_cust, plan.o_custkey)], plan.proj2, t2)
        jt = E.JoinTable(A.QS_INT, max(1024, stats.get("t2_estimate", stats["orders_rows"] // 4)), dev=dev)
        objs.append(jt)
        jt.build(t2, None, -1, 0, None, [(f_ord, 0)])
        t0 = E.Relation.create(plan.t0_schema, max(1, stats["lineitem_rows"]), dev=dev)
        objs.append(t0)
        E.select(lineitem, plan.e0, plan.p0, [(f_ord, plan.l_orderkey)], plan.proj0, t0)
        t4 = E.Relation.create(plan.t4_schema, max(1, stats.get("t4_capacity", stats["lineitem_rows"])), dev=dev)
        objs.append(t4)
        jt.probe(t0, plan.e4, -1, 0, A.QS_JOIN_INNER, -1, plan.proj4, t4)
        st = E.AggState(A.QS_AGG_SEPARATE_CHAINING, plan.e6, -1, plan.aggregates, plan.group_by,
                        estimated=max(1024, stats.get("groups_estimate", 1 << 16)), dev=dev)
        objs.append(st)
        st.run(t4)
        fin, _ = E.finalize_relation(st, plan.key_schema, [(A.QS_DOUBLE, 8)])
        objs.append(fin)
        top = E.topk(fin, plan.sort_keys, plan.limit)
        objs.append(top)
        ok, od, sp, rev = top.read(0), top.read(1), top.read(2), top.read(3)
        if timings is not None:
            timings["t2_rows"], timings["t0_rows"], timings["t4_rows"] = t2.n_rows, t0.n_rows, t4.n_rows
            timings["groups"] = fin.n_rows
        return [(int(ok[i]), float(rev[i]), (int(od[i]["year"]), int(od[i]["month"]), int(od[i]["day"])), int(sp[i]))
                for i in range(len(ok))]
    finally:
        for o in reversed(objs):
            o.destroy()
