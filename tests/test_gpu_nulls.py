"""NULL-able input attributes on the device path vs the NULL oracle (oracle/qs_null_oracle.py) and the fixtures of
the reference's aggregation-handle unit tests (expressions/aggregation/tests/AggregationHandle{Sum,Avg,Count,Min,
Max}_unittest.cpp: a column that starts with a NULL, has one in the middle and ends with one)."""
import numpy as np
import pytest

import qs_null_oracle as NO
from backends import agg_out_types
from quickstep_b200 import capi as A
from quickstep_b200.capi import QsGpuError
from quickstep_b200.expr import ExprSet
from quickstep_b200.table import Column, HostTable

pytestmark = pytest.mark.gpu

TOL = 1e-9      # relative, for double sums (BASELINE.json north_star); everything else is exact


def close(a, b):
    return a == b or (a != a and b != b) or abs(a - b) <= TOL * max(abs(a), abs(b))


def nullable_relation(engine, table: HostTable, nulls: np.ndarray, block_rows=None):
    """Device relation of `table` whose NULL-able attributes are those that have a bit anywhere in `nulls`
    (plus `extra`); NULL values are stored as zero bytes."""
    mask = int(np.bitwise_or.reduce(nulls)) if len(nulls) else 0
    for a, c in enumerate(table.columns):
        isn = ((nulls >> np.uint64(a)) & np.uint64(1)).astype(bool)
        if isn.any():
            raw = np.ascontiguousarray(c.data).copy()
            raw.view(np.uint8).reshape(len(raw), -1)[isn] = 0
            c.data = raw
    rel = engine.Relation.from_host(table, block_rows=block_rows)
    rel.set_nullable([a for a in range(len(table.columns)) if (mask >> a) & 1])
    if mask:
        rel.write_nulls(nulls)
    return rel


def run_agg(engine, rel, strategy, es, pred, aggregates, group_roots, key_schema, nullable_args, table, max_key=-1,
            work_orders=1):
    st = engine.AggState(strategy, es, pred, aggregates, group_roots, estimated=64, max_key=max_key,
                         nullable_args=nullable_args)
    try:
        n = rel.n_rows
        step = (n + work_orders - 1) // work_orders if n else 1
        for lo in range(0, max(n, 1), max(step, 1)):
            st.run(rel, lo, min(n, lo + step))
        out_types = agg_out_types(es, aggregates)
        out, mask = engine.finalize_relation(st, key_schema, out_types)
        try:
            cols = out.read_all()
            nulls = out.read_nulls()
        finally:
            out.destroy()
    finally:
        st.destroy()
    return cols, nulls, mask


def unittest_column(np_dtype, n_samples=100):
    """createColumnVectorGeneric of the reference's handle tests: NULL, the samples with one NULL in the middle, NULL."""
    vals, isnull = [0], [True]
    for i in range(n_samples):
        if np.dtype(np_dtype).kind == "i":
            vals.append(i - 10)
        else:
            vals.append(np.float32(i - 10) / np.float32(10))
        isnull.append(False)
        if i == n_samples // 2:
            vals.append(0); isnull.append(True)
    vals.append(0); isnull.append(True)
    return np.array(vals, dtype=np_dtype), np.array(isnull)


@pytest.mark.parametrize("qs_type,np_dtype", [(A.QS_INT, np.int32), (A.QS_LONG, np.int64), (A.QS_FLOAT, np.float32),
                                                (A.QS_DOUBLE, np.float64)])
def test_handle_unittest_fixture(engine, qs_type, np_dtype):
    """SUM / AVG / COUNT / MIN / MAX over the reference's unit-test column, then over an all-NULL column."""
    vals, isnull = unittest_column(np_dtype)
    t = HostTable("t", [Column("x", qs_type, vals)])
    nulls = isnull.astype(np.uint64)
    es = ExprSet()
    x = es.attr(0, qs_type)
    aggs = [(A.QS_AGG_SUM, x), (A.QS_AGG_AVG, x), (A.QS_AGG_COUNT, x), (A.QS_AGG_MIN, x), (A.QS_AGG_MAX, x), (A.QS_AGG_COUNT, -1)]
    rel = nullable_relation(engine, t, nulls)
    try:
        cols, out_nulls, mask = run_agg(engine, rel, A.QS_AGG_SINGLE_STATE, es, -1, aggs, [], [], [0, 1, 2, 3, 4], t)
    finally:
        rel.destroy()
    good = vals[~isnull]
    # the reference's own expectation: the sequential sum in the handle's precision type
    acc = np.float64(0) if np.dtype(np_dtype).kind == "f" else 0
    for v in good:
        acc = acc + (np.float64(v) if np.dtype(np_dtype).kind == "f" else int(v))
    assert mask == 0 and out_nulls[0] == 0
    if np.dtype(np_dtype).kind == "i":
        assert int(cols[0][0]) == acc == 3950
    else:
        assert close(float(cols[0][0]), float(acc))
    assert close(float(cols[1][0]), float(acc) / 100.0)
    assert int(cols[2][0]) == 100 and int(cols[5][0]) == 103
    assert cols[3][0] == good.min() and cols[4][0] == good.max()
    # every value NULL: SUM / AVG / MIN / MAX are NULL, COUNT(x) is 0 (finalize(...).isNull() in the unit tests)
    t2 = HostTable("t", [Column("x", qs_type, np.zeros(5000, dtype=np_dtype))])
    rel = nullable_relation(engine, t2, np.ones(5000, dtype=np.uint64))
    try:
        cols, out_nulls, mask = run_agg(engine, rel, A.QS_AGG_SINGLE_STATE, es, -1, aggs, [], [], [0, 1, 2, 3, 4], t2)
    finally:
        rel.destroy()
    assert mask == 0b011011 and out_nulls[0] == 0b011011
    assert int(cols[2][0]) == 0 and int(cols[5][0]) == 5000


def random_nullable_table(rng, n, groups=7):
    t = HostTable("t", [Column("g", A.QS_INT, rng.integers(0, groups, size=n).astype(np.int32)),
                        Column("x", A.QS_DOUBLE, rng.normal(10, 5, size=n)),
                        Column("y", A.QS_INT, rng.integers(-1000, 1000, size=n).astype(np.int32)),
                        Column("z", A.QS_LONG, rng.integers(-50, 50, size=n)),
                        Column("f", A.QS_FLOAT, rng.normal(size=n).astype(np.float32))])
    nulls = np.zeros(n, dtype=np.uint64)
    nulls |= (rng.random(n) < 0.3).astype(np.uint64) << np.uint64(1)
    nulls |= (rng.random(n) < 0.1).astype(np.uint64) << np.uint64(2)
    nulls |= (rng.random(n) < 0.5).astype(np.uint64) << np.uint64(4)
    return t, nulls


def check_agg(got_cols, got_nulls, n_keys, exp: dict, aggregates):
    keys = [None] * len(got_nulls) if n_keys == 0 else [int(k) for k in got_cols[0]]
    assert sorted(keys, key=lambda k: (k is None, k)) == sorted(exp.keys(), key=lambda k: (k is None, k))
    for row, k in enumerate(keys):
        for j, (val, is_null) in enumerate(exp[k]):
            got_null = bool((int(got_nulls[row]) >> (n_keys + j)) & 1)
            assert got_null == is_null, (k, j)
            if is_null:
                continue
            got = got_cols[n_keys + j][row]
            if isinstance(val, float):
                assert close(float(got), val), (k, j, got, val)
            else:
                assert int(got) == val, (k, j, got, val)


@pytest.mark.parametrize("strategy", [A.QS_AGG_SINGLE_STATE, A.QS_AGG_COMPACT_KEY, A.QS_AGG_SEPARATE_CHAINING,
                                      A.QS_AGG_COLLISION_FREE])
@pytest.mark.parametrize("work_orders", [1, 3])
def test_aggregates_skip_nulls(engine, strategy, work_orders):
    """All four aggregation strategies, a predicate over a NULL-able attribute, arguments that are NULL-able
    attributes and expressions over them; one group has only NULL arguments."""
    rng = np.random.default_rng(5 + strategy)
    t, nulls = random_nullable_table(rng, 40000)
    g = t.columns[0].data
    nulls[g == 3] |= np.uint64(1 << 1)                   # group 3: every x is NULL
    es = ExprSet()
    x, y, z, f = es.attr(1, A.QS_DOUBLE), es.attr(2, A.QS_INT), es.attr(3, A.QS_LONG), es.attr(4, A.QS_FLOAT)
    pred = es.or_(es.cmp(A.QS_GT, y, es.lit_int(-500)), es.not_(es.cmp(A.QS_LT, x, es.lit_double(12.0))))
    minmax = strategy != A.QS_AGG_COLLISION_FREE          # that table takes COUNT / SUM / AVG only (as in the reference)
    aggs = [(A.QS_AGG_SUM, x), (A.QS_AGG_AVG, x), (A.QS_AGG_COUNT, x), (A.QS_AGG_SUM, es.mul(x, es.cast(y, A.QS_DOUBLE))),
            (A.QS_AGG_COUNT, -1), (A.QS_AGG_SUM, z), (A.QS_AGG_AVG, f)]
    nullable = [0, 1, 2, 3, 6]
    if minmax:
        aggs += [(A.QS_AGG_MIN, y), (A.QS_AGG_MAX, x)]
        nullable += [7, 8]
    grouped = strategy != A.QS_AGG_SINGLE_STATE
    rel = nullable_relation(engine, t, nulls, block_rows=9973)
    try:
        cols, out_nulls, _ = run_agg(engine, rel, strategy, es, pred, aggs, [es.attr(0, A.QS_INT)] if grouped else [],
                                     [(A.QS_INT, 4)] if grouped else [], nullable, t, max_key=6, work_orders=work_orders)
    finally:
        rel.destroy()
    exp = NO.aggregate(es, pred, aggs, 0 if grouped else None, t, nulls)
    if grouped:      # the engine's own answers for such a group: SUM 0, AVG NaN, COUNT(x) 0, MAX NULL
        assert exp[3][0] == (0.0, False) and exp[3][1][0] != exp[3][1][0] and exp[3][2] == (0, False)
        if minmax:
            assert exp[3][8] == (0, True)
    check_agg(cols, out_nulls, 1 if grouped else 0, exp, aggs)


def test_predicates_and_projection(engine):
    """Select: comparisons with NULL operands are false, NOT complements, and the NULL-ness of every projected
    column (attribute or expression) arrives in the output relation's mask."""
    rng = np.random.default_rng(77)
    t, nulls = random_nullable_table(rng, 30011)
    es = ExprSet()
    x, y, z = es.attr(1, A.QS_DOUBLE), es.attr(2, A.QS_INT), es.attr(3, A.QS_LONG)
    preds = [es.cmp(A.QS_LT, x, es.lit_double(11.0)),
             es.not_(es.cmp(A.QS_LT, x, es.lit_double(11.0))),
             es.and_(es.cmp(A.QS_GE, y, es.lit_int(0)), es.cmp(A.QS_LT, es.add(x, es.cast(y, A.QS_DOUBLE)), es.lit_double(400.0))),
             es.or_(es.cmp(A.QS_EQ, z, es.lit_long(7)), es.cmp(A.QS_GT, es.cast(y, A.QS_LONG), z)),
             es.cmp(A.QS_NE, y, es.lit_int(0))]
    roots = [es.attr(0, A.QS_INT), x, es.add(x, es.cast(y, A.QS_DOUBLE)), z, es.mul(z, es.lit_long(2)), y]
    schema = [(A.QS_INT, 4), (A.QS_DOUBLE, 8), (A.QS_DOUBLE, 8), (A.QS_LONG, 8), (A.QS_LONG, 8), (A.QS_INT, 4)]
    rel = nullable_relation(engine, t, nulls, block_rows=4099)
    try:
        for pred in preds:
            out = engine.Relation.create(schema, t.n_rows)
            try:
                engine.select(rel, es, pred, None, roots, out)
                got = out.read_all()
                got_nulls = out.read_nulls()
            finally:
                out.destroy()
            keep = NO.predicate(es, pred, t, nulls)
            assert len(got_nulls) == keep.sum() and 0 < keep.sum() < t.n_rows
            import qs_oracle as O
            exp_cols = [O.scalar(es, r, t)[keep] if es.nodes[r].kind != A.QS_N_ATTRIBUTE else t.columns[es.nodes[r].a].data[keep] for r in roots]
            exp_nulls = np.zeros(int(keep.sum()), dtype=np.uint64)
            for j, r in enumerate(roots):
                exp_nulls |= NO.null_of(es, r, nulls)[keep].astype(np.uint64) << np.uint64(j)
            # tiles interleave: compare as row multisets; a NULL value's bytes are not part of the answer
            def rows(cols, nl):
                out_rows = []
                for i in range(len(nl)):
                    out_rows.append(tuple(None if (int(nl[i]) >> j) & 1 else cols[j][i].item() for j in range(len(cols))))
                return sorted(out_rows, key=repr)
            assert rows(got, got_nulls) == rows(exp_cols, exp_nulls)
    finally:
        rel.destroy()


@pytest.mark.parametrize("table", ["open", "dense"])
@pytest.mark.parametrize("join_type", [A.QS_JOIN_LEFT_ANTI, A.QS_JOIN_LEFT_OUTER])
def test_anti_and_outer_join_emit_null_key_rows(engine, table, join_type):
    """A probe row with a NULL key matches nothing: the anti join emits it, the outer join emits it NULL-padded
    (HashTable::runOverKeysFromValueAccessor, storage/HashTable.hpp:1999-2003); NULL build keys never enter the table."""
    rng = np.random.default_rng(41)
    nb, npr = 1500, 12000
    build = HostTable("b", [Column("k", A.QS_INT, rng.integers(0, 600, size=nb).astype(np.int32)),
                            Column("p", A.QS_LONG, np.arange(nb, dtype=np.int64))])
    bnull = (rng.random(nb) < 0.2).astype(np.uint64)
    probe = HostTable("p", [Column("k", A.QS_INT, rng.integers(0, 900, size=npr).astype(np.int32)),
                            Column("i", A.QS_LONG, np.arange(npr, dtype=np.int64))])
    pnull = (rng.random(npr) < 0.25).astype(np.uint64)
    es = ExprSet()
    pred = es.cmp(A.QS_GE, es.attr(1, A.QS_LONG), es.lit_long(100))          # rows 0..99 fail the probe predicate
    outer = join_type == A.QS_JOIN_LEFT_OUTER
    roots = [es.attr(1, A.QS_LONG), es.attr(0, A.QS_INT)] + ([es.attr(1, A.QS_LONG, 8, 2)] if outer else [])
    schema = [(A.QS_LONG, 8), (A.QS_INT, 4)] + ([(A.QS_LONG, 8)] if outer else [])
    brel, prel = nullable_relation(engine, build, bnull), nullable_relation(engine, probe, pnull, block_rows=2999)
    jt = engine.JoinTable(A.QS_INT, nb, dense_range=(0, 599) if table == "dense" else None)
    out = engine.Relation.create(schema, 200000)
    try:
        jt.build(brel, None, -1, 0)
        jt.probe(prel, es, pred, 0, join_type, -1, roots, out)
        got, got_nulls = out.read_all(), out.read_nulls()
    finally:
        out.destroy(); jt.destroy(); brel.destroy(); prel.destroy()
    bk, pk = build.columns[0].data, probe.columns[0].data
    keep = np.arange(npr) >= 100
    pairs = NO.join_pairs(bk, bnull.astype(bool), pk, pnull.astype(bool), keep)
    matched = {p for p, _b in pairs}
    unmatched = [p for p in range(npr) if keep[p] and p not in matched]        # NULL-key rows included
    assert sum(1 for p in unmatched if pnull[p]) > 1000
    key_of = lambda p: None if pnull[p] else int(pk[p])
    if outer:
        exp = sorted([(p, key_of(p), int(b)) for p, b in pairs] + [(p, key_of(p), None) for p in unmatched], key=repr)
        gotr = sorted(((int(got[0][i]), None if (int(got_nulls[i]) >> 1) & 1 else int(got[1][i]),
                        None if (int(got_nulls[i]) >> 2) & 1 else int(got[2][i])) for i in range(len(got_nulls))), key=repr)
    else:
        exp = sorted(((p, key_of(p)) for p in unmatched), key=repr)
        gotr = sorted(((int(got[0][i]), None if (int(got_nulls[i]) >> 1) & 1 else int(got[1][i])) for i in range(len(got_nulls))), key=repr)
    assert gotr == exp


@pytest.mark.parametrize("join_type", [A.QS_JOIN_INNER, A.QS_JOIN_LEFT_OUTER])
def test_join_build_side_nullable_attributes(engine, join_type):
    """NULL-able attributes of the BUILD side read through the join: projected as they are, inside an expression with a
    probe attribute, and in the residual predicate (a comparison with a NULL is false).  The NULL-ness comes from the
    matched build row's mask."""
    rng = np.random.default_rng(77)
    nb, npr = 900, 8000
    build = HostTable("b", [Column("k", A.QS_INT, rng.permutation(1200)[:nb].astype(np.int32)),          # unique keys
                            Column("q", A.QS_DOUBLE, rng.normal(size=nb)),
                            Column("w", A.QS_LONG, rng.integers(-5, 5, size=nb))])
    bnull = ((rng.random(nb) < 0.3).astype(np.uint64) << np.uint64(1)) | ((rng.random(nb) < 0.3).astype(np.uint64) << np.uint64(2))
    probe = HostTable("p", [Column("k", A.QS_INT, rng.integers(0, 1200, size=npr).astype(np.int32)),
                            Column("v", A.QS_DOUBLE, rng.normal(size=npr))])
    pnull = (rng.random(npr) < 0.2).astype(np.uint64) << np.uint64(1)
    es = ExprSet()
    inner = join_type == A.QS_JOIN_INNER
    bq, bw, pv = es.attr(1, A.QS_DOUBLE, 8, 2), es.attr(2, A.QS_LONG, 8, 2), es.attr(1, A.QS_DOUBLE)
    residual = es.cmp(A.QS_GE, bw, es.lit_long(-2)) if inner else -1           # an outer join takes no residual
    roots = [es.attr(0, A.QS_INT), bq, es.add(pv, bq), bw]
    schema = [(A.QS_INT, 4), (A.QS_DOUBLE, 8), (A.QS_DOUBLE, 8), (A.QS_LONG, 8)]
    brel, prel = nullable_relation(engine, build, bnull), nullable_relation(engine, probe, pnull, block_rows=1999)
    jt = engine.JoinTable(A.QS_INT, nb)
    out = engine.Relation.create(schema, 50000)
    try:
        jt.build(brel, None, -1, 0)
        jt.probe(prel, es, -1, 0, join_type, residual, roots, out)
        got, got_nulls = out.read_all(), out.read_nulls()
    finally:
        out.destroy(); jt.destroy(); brel.destroy(); prel.destroy()
    row_of = {int(k): i for i, k in enumerate(build.columns[0].data)}
    qn, wn = ((bnull >> np.uint64(1)) & np.uint64(1)).astype(bool), ((bnull >> np.uint64(2)) & np.uint64(1)).astype(bool)
    vn = ((pnull >> np.uint64(1)) & np.uint64(1)).astype(bool)
    exp = []
    for p in range(npr):
        b = row_of.get(int(probe.columns[0].data[p]))
        if b is None:
            if not inner:
                exp.append((int(probe.columns[0].data[p]), None, None, None))
            continue
        if inner and (wn[b] or build.columns[2].data[b] < -2):
            continue
        q = None if qn[b] else float(build.columns[1].data[b])
        s_ = None if (qn[b] or vn[p]) else float(probe.columns[1].data[p]) + float(build.columns[1].data[b])
        exp.append((int(probe.columns[0].data[p]), q, s_, None if wn[b] else int(build.columns[2].data[b])))
    gotr = []
    for i in range(len(got_nulls)):
        m = int(got_nulls[i])
        gotr.append((int(got[0][i]), None if (m >> 1) & 1 else float(got[1][i]), None if (m >> 2) & 1 else float(got[2][i]),
                     None if (m >> 3) & 1 else int(got[3][i])))
    assert len(exp) > 2000 and sorted(gotr, key=repr) == sorted(exp, key=repr)


@pytest.mark.parametrize("table", ["open", "dense"])
@pytest.mark.parametrize("join_type", [A.QS_JOIN_INNER, A.QS_JOIN_LEFT_SEMI])
def test_join_null_keys(engine, table, join_type):
    """NULL keys neither enter the table nor match; the probe side's NULL-able projections keep their masks."""
    rng = np.random.default_rng(31)
    nb, npr = 2000, 15000
    build = HostTable("b", [Column("k", A.QS_INT, rng.integers(0, 600, size=nb).astype(np.int32)),
                            Column("p", A.QS_LONG, np.arange(nb, dtype=np.int64))])
    bnull = (rng.random(nb) < 0.2).astype(np.uint64)
    probe = HostTable("p", [Column("k", A.QS_INT, rng.integers(0, 900, size=npr).astype(np.int32)),
                            Column("v", A.QS_DOUBLE, rng.normal(size=npr))])
    pnull = (rng.random(npr) < 0.25).astype(np.uint64) | ((rng.random(npr) < 0.4).astype(np.uint64) << np.uint64(1))
    es = ExprSet()
    inner = join_type == A.QS_JOIN_INNER
    roots = [es.attr(0, A.QS_INT), es.attr(1, A.QS_DOUBLE)] + ([es.attr(1, A.QS_LONG, 8, 2)] if inner else [])
    schema = [(A.QS_INT, 4), (A.QS_DOUBLE, 8)] + ([(A.QS_LONG, 8)] if inner else [])
    brel, prel = nullable_relation(engine, build, bnull), nullable_relation(engine, probe, pnull, block_rows=3001)
    jt = engine.JoinTable(A.QS_INT, nb, dense_range=(0, 599) if table == "dense" else None)
    out = engine.Relation.create(schema, 200000)
    try:
        jt.build(brel, None, -1, 0)
        assert jt.num_entries() == int((bnull == 0).sum())
        jt.probe(prel, es, -1, 0, join_type, -1, roots, out)
        got = out.read_all()
        got_nulls = out.read_nulls()
    finally:
        out.destroy(); jt.destroy(); brel.destroy(); prel.destroy()
    bk, pk, pv = build.columns[0].data, probe.columns[0].data, probe.columns[1].data
    pairs = NO.join_pairs(bk, bnull.astype(bool), pk, (pnull & np.uint64(1)).astype(bool))
    vnull = ((pnull >> np.uint64(1)) & np.uint64(1)).astype(bool)
    if inner:
        exp = sorted(((int(pk[p]), None if vnull[p] else float(pv[p]), int(b)) for p, b in pairs), key=repr)
        gotr = sorted(((int(got[0][i]), None if (int(got_nulls[i]) >> 1) & 1 else float(got[1][i]), int(got[2][i])) for i in range(len(got_nulls))), key=repr)
    else:
        exp = sorted({p: (int(pk[p]), None if vnull[p] else float(pv[p])) for p, _b in pairs}.values(), key=repr)
        gotr = sorted(((int(got[0][i]), None if (int(got_nulls[i]) >> 1) & 1 else float(got[1][i])) for i in range(len(got_nulls))), key=repr)
    assert len(exp) > 1000 and gotr == exp


def test_lip_filters_skip_nulls(engine):
    """A filter built from a NULL-able attribute holds the non-NULL values only; probing it drops NULL rows."""
    rng = np.random.default_rng(8)
    n = 20000
    t = HostTable("t", [Column("k", A.QS_INT, rng.integers(0, 4000, size=n).astype(np.int32))])
    nulls = (rng.random(n) < 0.5).astype(np.uint64)
    rel = nullable_relation(engine, t, nulls)
    lip = engine.LipFilter(A.QS_LIP_BITVECTOR_EXACT, A.QS_INT, 0, 3999)
    es = ExprSet()
    out = engine.Relation.create([(A.QS_INT, 4)], n)
    try:
        engine.build_lip_filter(rel, es, es.cmp(A.QS_LT, es.attr(0, A.QS_INT), es.lit_int(2000)), None, [(lip, 0)])
        words = lip.words()
        bits = np.unpackbits(words.astype(">u8").view(np.uint8))[:4000].astype(bool)
        keys = t.columns[0].data
        exp = np.zeros(4000, dtype=bool)
        exp[np.unique(keys[(nulls == 0) & (keys < 2000)])] = True
        assert (bits == exp).all()
        # NULL keys (stored as 0; bit 0 may well be set) never pass the probe
        engine.select(rel, es, -1, [(lip, 0)], [es.attr(0, A.QS_INT)], out)
        got = np.sort(out.read(0))
        want = np.sort(keys[(nulls == 0) & exp[keys]])
        assert (got == want).all() and out.read_nulls().max(initial=0) == 0
    finally:
        out.destroy(); lip.destroy(); rel.destroy()


def make_null_block_images(rng, n):
    """One block per NULL representation of the reference's formats, as raw images with stage descriptors."""
    images, expect = [], []
    # 1) column store: plain stripe + BitVector<false> over the rows (64-bit words, MSB first)
    vals = rng.integers(-10**6, 10**6, size=n).astype(np.int64)
    isn = rng.random(n) < 0.3
    words = np.zeros((n + 63) // 64, dtype=np.uint64)
    for i in np.nonzero(isn)[0]:
        words[i >> 6] |= np.uint64(1) << np.uint64(63 - (i & 63))
    img = np.zeros(64 + words.nbytes + vals.nbytes, dtype=np.uint8)
    img[64:64 + words.nbytes] = words.view(np.uint8)
    img[64 + words.nbytes:] = vals.view(np.uint8)
    images.append((img, n, [dict(attr=0, encoding=A.QS_ENC_PLAIN, offset=64 + words.nbytes, null_kind=A.QS_NULL_BITMAP,
                                 null_arg=0, null_stride=1, null_offset=64)]))
    expect.append((vals, isn))
    # 2) packed bitmap shared by several attributes: bit = row * 3 + 1 (CompressedBlockBuilder.cpp:233-236)
    vals = rng.integers(-10**6, 10**6, size=n).astype(np.int64)
    isn = rng.random(n) < 0.6
    words = np.zeros((3 * n + 63) // 64, dtype=np.uint64)
    for i in np.nonzero(isn)[0]:
        g = 3 * int(i) + 1
        words[g >> 6] |= np.uint64(1) << np.uint64(63 - (g & 63))
    img = np.concatenate([words.view(np.uint8), vals.view(np.uint8)])
    images.append((img, n, [dict(attr=0, encoding=A.QS_ENC_PLAIN, offset=words.nbytes, null_kind=A.QS_NULL_BITMAP,
                                 null_arg=1, null_stride=3, null_offset=0)]))
    expect.append((vals, isn))
    # 3) split row store slots: [null word][LONG], BitVector<true> of 1 / 2 / 4 / 8 bytes, various bits
    for width, bit in ((1, 0), (1, 5), (2, 11), (4, 17), (8, 40)):
        vals = rng.integers(-10**6, 10**6, size=n).astype(np.int64)
        isn = rng.random(n) < 0.4
        stride = width + 8 + 3
        img = rng.integers(0, 256, size=n * stride + 32).astype(np.uint8)
        for i in range(n):
            word = int(rng.integers(0, 1 << 62)) & ((1 << (8 * width)) - 1) & ~(1 << (8 * width - 1 - bit))
            if isn[i]:
                word |= 1 << (8 * width - 1 - bit)
            img[i * stride:i * stride + width] = np.frombuffer(word.to_bytes(width, "little"), dtype=np.uint8)
            img[i * stride + width:i * stride + width + 8] = vals[i:i + 1].view(np.uint8)
        images.append((img, n, [dict(attr=0, encoding=A.QS_ENC_STRIDED, offset=width, stride=stride,
                                     null_kind=A.QS_NULL_SLOT_WORD, null_arg=bit, null_stride=stride, null_width=width,
                                     null_offset=0)]))
        expect.append((vals, isn))
    # 4) dictionary-compressed stripe whose NULL code is the number of codes (CompressionDictionaryLite.hpp:40-51)
    dvals = np.sort(rng.choice(np.arange(-5000, 5000), size=200, replace=False)).astype(np.int64)
    codes = rng.integers(0, 201, size=n).astype(np.uint8)
    img = np.concatenate([dvals.view(np.uint8), codes])
    images.append((img, n, [dict(attr=0, encoding=A.QS_ENC_DICT, offset=dvals.nbytes, code_width=1, dict_offset=0,
                                 dict_entries=200, null_kind=A.QS_NULL_CODE, null_arg=200)]))
    expect.append((np.where(codes == 200, 0, dvals[np.minimum(codes, 199)]), codes == 200))
    return images, expect


def test_stage_null_representations(engine):
    rng = np.random.default_rng(91)
    n = 5000
    images, expect = make_null_block_images(rng, n)
    rel = engine.Relation.create([(A.QS_LONG, 8)], n * len(images))
    try:
        rel.set_nullable([0])
        rel.stage_blocks(images)
        assert rel.n_rows == n * len(images)
        got, got_nulls = rel.read(0), rel.read_nulls()
        for b, (vals, isn) in enumerate(expect):
            sl = slice(b * n, (b + 1) * n)
            assert (got_nulls[sl] == isn.astype(np.uint64)).all(), b
            assert (got[sl] == np.where(isn, 0, vals)).all(), b            # a NULL value is stored as zero bytes
        # ... and an aggregate over the staged relation skips exactly those rows
        es = ExprSet()
        x = es.attr(0, A.QS_LONG)
        cols, _n, mask = run_agg(engine, rel, A.QS_AGG_SINGLE_STATE, es, -1, [(A.QS_AGG_SUM, x), (A.QS_AGG_COUNT, x)], [], [], [0, 1], None)
        allv = np.concatenate([np.where(isn, 0, v) for v, isn in expect])
        alln = np.concatenate([isn for _v, isn in expect])
        assert mask == 0 and int(cols[0][0]) == int(allv.sum()) and int(cols[1][0]) == int((~alln).sum())
    finally:
        rel.destroy()


def test_topk_and_partition_carry_masks(engine):
    rng = np.random.default_rng(4)
    n = 9000
    t = HostTable("t", [Column("k", A.QS_INT, rng.permutation(n).astype(np.int32)), Column("v", A.QS_LONG, rng.integers(0, 100, size=n))])
    nulls = (rng.random(n) < 0.5).astype(np.uint64) << np.uint64(1)
    rel = nullable_relation(engine, t, nulls)
    try:
        top = engine.topk(rel, [(0, False)], 50)
        try:
            k, m = top.read(0), top.read_nulls()
            assert (k == np.arange(50)).all()
            order = np.argsort(t.columns[0].data)[:50]
            assert (m == nulls[order]).all()
        finally:
            top.destroy()
        out = engine.Relation.create(rel.schema, n)
        try:
            off = engine.hash_partition(rel, 0, 4, out)
            k, m = out.read(0), out.read_nulls()
            assert off[-1] == n
            assert (m == nulls[np.argsort(t.columns[0].data)][k]).all()
        finally:
            out.destroy()
    finally:
        rel.destroy()


@pytest.mark.parametrize("n", [3000, 40000])          # one-CTA top-k / the grid-wide (cooperative) one
def test_order_by_nullable_keys(engine, n):
    """ORDER BY on NULL-able attributes: NULLS FIRST / LAST as asked, else the reference's default -- NULLs first iff
    the key is descending (parser/ParseOrderBy.hpp:53-66).  Few distinct values, so later keys and the NULL ranks of
    several keys decide; LONG extremes are present, which a NULL must still sort strictly before / after."""
    rng = np.random.default_rng(n)
    small = n <= 16384
    # (at most 2048 rows may tie on the FIRST key at the cut -- the top-k's documented limit -- so the large case draws
    # the first key from a wide domain and keeps its NULLs under that bound)
    a = (rng.integers(-3, 4, size=n) if small else rng.integers(-10**6, 10**6, size=n)).astype(np.int64)
    a[rng.random(n) < (0.05 if small else 0.002)] = np.iinfo(np.int64).max
    a[rng.random(n) < (0.05 if small else 0.002)] = np.iinfo(np.int64).min
    b = np.round(rng.normal(size=n), 1)
    t = HostTable("t", [Column("a", A.QS_LONG, a), Column("b", A.QS_DOUBLE, b), Column("i", A.QS_LONG, rng.permutation(n).astype(np.int64))])
    nulls = ((rng.random(n) < (0.3 if small else 0.03)).astype(np.uint64)) | ((rng.random(n) < 0.3).astype(np.uint64) << np.uint64(1))
    rel = nullable_relation(engine, t, nulls)
    an, bn = (nulls & np.uint64(1)).astype(bool), ((nulls >> np.uint64(1)) & np.uint64(1)).astype(bool)
    try:
        for (da, nfa), (db, nfb) in [((False, None), (True, None)), ((True, None), (False, True)), ((False, True), (True, False)),
                                     ((True, False), (False, None))]:
            limit = 600
            top = engine.topk(rel, [(0, da, nfa), (1, db, nfb), (2, False)], limit)
            try:
                got_i, got_m = top.read(2), top.read_nulls()
            finally:
                top.destroy()

            def part(vals, isnull, desc, nf):
                nf = desc if nf is None else nf
                out = []
                for v, isn in zip(vals.tolist(), isnull.tolist()):
                    out.append((0 if nf else 2, 0) if isn else (1, -v if desc else v))
                return out
            ka, kb = part(a.astype(object), an, da, nfa), part(b, bn, db, nfb)
            order = sorted(range(n), key=lambda r: (ka[r], kb[r], int(t.columns[2].data[r])))[:limit]
            assert got_i.tolist() == [int(t.columns[2].data[r]) for r in order], (da, nfa, db, nfb)
            assert (got_m == nulls[order]).all()
    finally:
        rel.destroy()


@pytest.mark.parametrize("strategy", [A.QS_AGG_COMPACT_KEY, A.QS_AGG_SEPARATE_CHAINING])
def test_group_by_nullable_key_drops_null_rows(engine, strategy):
    """Rows whose group-by key is NULL belong to no group (PackedPayloadHashTable.hpp:861-866; the engine's own output
    is in tests/test_reference_nulls.py): as many groups as distinct non-NULL keys, counts over the rest."""
    rng = np.random.default_rng(12)
    t, nulls = random_nullable_table(rng, 20000)
    t.columns[2].data[:] = rng.integers(0, 40, size=t.n_rows)          # y: 40 distinct keys, 10 % NULL
    es = ExprSet()
    x, y = es.attr(1, A.QS_DOUBLE), es.attr(2, A.QS_INT)
    aggs = [(A.QS_AGG_COUNT, -1), (A.QS_AGG_SUM, x), (A.QS_AGG_COUNT, x)]
    rel = nullable_relation(engine, t, nulls)
    try:
        cols, out_nulls, _ = run_agg(engine, rel, strategy, es, -1, aggs, [y], [(A.QS_INT, 4)], [1, 2], t)
    finally:
        rel.destroy()
    exp = NO.aggregate(es, -1, aggs, 2, t, nulls)
    assert len(exp) == 40 and sum(v[0][0] for v in exp.values()) == int((((nulls >> np.uint64(2)) & np.uint64(1)) == 0).sum())
    check_agg(cols, out_nulls, 1, exp, aggs)


def test_refusals(engine):
    """What is not lowered for NULL-able attributes fails loudly instead of ignoring the masks."""
    rng = np.random.default_rng(2)
    t, nulls = random_nullable_table(rng, 1000)
    rel = nullable_relation(engine, t, nulls)
    es = ExprSet()
    try:
        with pytest.raises(QsGpuError):      # aggregate over a NULL-able attribute not declared as such
            st = engine.AggState(A.QS_AGG_SINGLE_STATE, es, -1, [(A.QS_AGG_SUM, es.attr(1, A.QS_DOUBLE))], [])
            try:
                st.run(rel)
            finally:
                st.destroy()
        with pytest.raises(QsGpuError):      # relation-wide dictionary codes for a NULL-able attribute
            rel.set_dictionary(2, 2, np.arange(-1000, 1000, dtype=np.int32))
    finally:
        rel.destroy()
