"""Per-row expression parity against the UNMODIFIED reference's own expression classes.

tests/golden/reference_expressions.json was written by tests/golden/make_expr_golden.cpp, a program linked against the
reference engine's libraries: 101 predicates / scalars (69 over NOT NULL attributes, 32 over NULL-able ones) built from ScalarAttribute, ScalarLiteral, ScalarUnaryExpression,
ScalarBinaryExpression, ScalarSharedExpression, ComparisonPredicate, NegationPredicate, Conjunction / DisjunctionPredicate,
evaluated by the reference's vectorised Predicate::getAllMatches / Scalar::getAllValues over 512 tuples (with NaN, +-0.0,
+-inf and a subnormal among the doubles), and lowered from their getProto() by the in-tree binding
(quickstep_b200/host/intree/ProtoLowering.hpp) into qs_node arrays.

Here the SAME node arrays are evaluated by the oracle (CPU, `-m "not gpu"`) and by the CUDA path through the C ABI
(`-m gpu`), and the results must be the reference's: match sets exact; scalar values bit for bit (SURVEY.md 8a rows P1,
P2, E1), except that two NaNs count as equal whatever their sign / payload bits (x86 SSE produces the negative default NaN
for 0.0 / 0.0, the GPU the positive one; the reference itself never looks at those bits).

The NULL-able cases run over the same tuples with every fifth value or so of five attributes NULL: the reference's answers
there are "a comparison with a NULL operand is false, NOT complements it, arithmetic over a NULL is NULL" -- the rules
oracle/qs_null_oracle.py restates and the device path implements with its per-row NULL masks.
"""
import json
import os

import numpy as np
import pytest

import qs_null_oracle as NO
import qs_oracle as O
from backends import GpuBackend, OracleBackend
from quickstep_b200 import capi as A
from quickstep_b200.expr import ExprSet
from quickstep_b200.table import Column, HostTable, np_dtype

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_expressions.json")))
CASES = [c for c in GOLDEN["cases"] if not c["nullable"]]
NULL_CASES = [c for c in GOLDEN["cases"] if c["nullable"]]
N = GOLDEN["n_rows"]
NULLS = np.frombuffer(bytes.fromhex(GOLDEN["nulls"]), dtype="<u8").copy()
RID = len(GOLDEN["columns"])          # attribute id of the row-id column appended below


def the_table() -> HostTable:
    cols = []
    for c in GOLDEN["columns"]:
        raw = np.frombuffer(bytes.fromhex(c["data"]), dtype=np.uint8)
        assert len(raw) == N * c["width"]
        cols.append(Column(c["name"], c["type"], raw.view(np_dtype(c["type"], c["width"])).copy(), c["width"]))
    cols.append(Column("rid", A.QS_INT, np.arange(N, dtype=np.int32)))
    return HostTable("t", cols)


def expr_set(case) -> ExprSet:
    """The node array exactly as the in-tree lowering produced it, plus one attribute node for the row id."""
    es = ExprSet()
    for kind, op, typ, width, a, b, lit in case["nodes"]:
        n = A.qs_node()
        n.kind, n.op, n.type, n.width, n.a, n.b = kind, op, typ, width, a, b
        n.lit.i64 = int.from_bytes(bytes.fromhex(lit), "little", signed=True)
        es.nodes.append(n)
    es.pool = bytearray(bytes.fromhex(case["pool"]))
    rid = es.attr(RID, A.QS_INT)
    return es, rid


def run_case(backend, rel, case):
    es, rid = expr_set(case)
    if case["kind"] == "predicate":
        out = backend.select(rel, es, case["root"], None, [rid], [(A.QS_INT, 4)])
        got = np.zeros(N, dtype=bool)
        ids = out.columns[0].data
        assert len(np.unique(ids)) == len(ids)
        got[ids] = True
        want = np.array([ch == "1" for ch in case["matches"]])
        assert int(want.sum()) == case["n_matches"]
        bad = np.nonzero(got != want)[0]
        assert len(bad) == 0, (case["name"], case["sql"], "rows", bad[:10])
    else:
        t, w = case["result_type"], case["result_width"]
        out = backend.select(rel, es, -1, None, [rid, case["root"]], [(A.QS_INT, 4), (t, w)])
        ids = out.columns[0].data
        assert sorted(ids.tolist()) == list(range(N))
        vals = np.ascontiguousarray(out.columns[1].data)[np.argsort(ids, kind="stable")]
        got = vals.view(np.uint8).reshape(N, w)
        want = np.frombuffer(bytes.fromhex(case["values"]), dtype=np.uint8).reshape(N, w)
        same = (got == want).all(axis=1)
        if t in (A.QS_FLOAT, A.QS_DOUBLE):
            f = np.dtype("<f4") if t == A.QS_FLOAT else np.dtype("<f8")
            same |= np.isnan(got.copy().view(f).ravel()) & np.isnan(want.copy().view(f).ravel())
        bad = np.nonzero(~same)[0]
        assert len(bad) == 0, (case["name"], case["sql"], "rows", bad[:10], got[bad[:3]], want[bad[:3]])


def check_scalar(case, ids, vals, val_nulls):
    """Rows may arrive in any order; a NULL value's bytes are not part of the answer."""
    t, w = case["result_type"], case["result_width"]
    assert sorted(ids.tolist()) == list(range(N))
    order = np.argsort(ids, kind="stable")
    got = np.ascontiguousarray(vals)[order].view(np.uint8).reshape(N, w)
    got_null = np.asarray(val_nulls, dtype=bool)[order]
    want = np.frombuffer(bytes.fromhex(case["values"]), dtype=np.uint8).reshape(N, w)
    want_null = np.array([ch == "1" for ch in case["value_nulls"]])
    assert (got_null == want_null).all(), (case["name"], np.nonzero(got_null != want_null)[0][:10])
    same = (got == want).all(axis=1) | want_null
    if t in (A.QS_FLOAT, A.QS_DOUBLE):
        f = np.dtype("<f4") if t == A.QS_FLOAT else np.dtype("<f8")
        same |= np.isnan(got.copy().view(f).ravel()) & np.isnan(want.copy().view(f).ravel())
    bad = np.nonzero(~same)[0]
    assert len(bad) == 0, (case["name"], case["sql"], "rows", bad[:10], got[bad[:3]], want[bad[:3]])


def test_golden_file_is_what_the_generator_writes():
    assert len(CASES) == 69 and len(NULL_CASES) == 32 and N == 512
    assert GOLDEN["nullable_attributes"] == [0, 2, 4, 7, 9]
    assert set(np.nonzero([(NULLS >> np.uint64(a) & np.uint64(1)).any() for a in range(RID)])[0]) == {0, 2, 4, 7, 9}
    assert sum(c["kind"] == "predicate" for c in CASES) == 39
    d = the_table().col("d").data
    assert np.isnan(d).any() and np.isinf(d).any() and (np.signbit(d) & (d == 0)).any()      # the special values are there


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_gives_the_reference_classes_results(oracle, case):
    run_case(OracleBackend(), the_table(), case)


@pytest.mark.parametrize("case", NULL_CASES, ids=[c["name"] for c in NULL_CASES])
def test_null_oracle_gives_the_reference_classes_results(oracle, case):
    t = the_table()
    es, _rid = expr_set(case)
    root = case["root"]
    if case["kind"] == "predicate":
        got = NO.predicate(es, root, t, NULLS)
        want = np.array([ch == "1" for ch in case["matches"]])
        assert (got == want).all(), (case["name"], case["sql"], np.nonzero(got != want)[0][:10])
    else:
        n = es.nodes[root]
        vals = t.columns[n.a].data if n.kind == A.QS_N_ATTRIBUTE else O.scalar(es, root, t)
        check_scalar(case, np.arange(N), vals, NO.null_of(es, root, NULLS))


@pytest.mark.gpu
@pytest.mark.parametrize("block_rows", [None, 100])
def test_cuda_path_gives_the_reference_classes_results_on_nullable_attributes(engine, block_rows):
    """The 32 NULL-able cases through qsgpu_select on a relation with a per-row NULL mask."""
    from test_gpu_nulls import nullable_relation
    rel = nullable_relation(engine, the_table(), NULLS, block_rows=block_rows)
    try:
        for case in NULL_CASES:
            es, rid = expr_set(case)
            if case["kind"] == "predicate":
                out = engine.Relation.create([(A.QS_INT, 4)], N)
                try:
                    engine.select(rel, es, case["root"], None, [rid], out)
                    ids = out.read_all()[0]
                finally:
                    out.destroy()
                got = np.zeros(N, dtype=bool)
                assert len(np.unique(ids)) == len(ids)
                got[ids] = True
                want = np.array([ch == "1" for ch in case["matches"]])
                assert (got == want).all(), (case["name"], case["sql"], np.nonzero(got != want)[0][:10])
            else:
                out = engine.Relation.create([(A.QS_INT, 4), (case["result_type"], case["result_width"])], N)
                try:
                    engine.select(rel, es, -1, None, [rid, case["root"]], out)
                    cols, nulls = out.read_all(), out.read_nulls()
                finally:
                    out.destroy()
                assert ((nulls & np.uint64(1)) == 0).all()                       # the row id is never NULL
                check_scalar(case, cols[0], cols[1], (nulls >> np.uint64(1)) & np.uint64(1))
    finally:
        rel.destroy()
        engine.synchronize()


@pytest.mark.gpu
@pytest.mark.parametrize("block_rows", [None, 100])
def test_cuda_path_gives_the_reference_classes_results(engine, block_rows):
    """All 69 cases through qsgpu_select on one staged relation (whole, and staged in 100-row blocks)."""
    G = GpuBackend(engine, block_rows=block_rows)
    try:
        rel = G.relation(the_table())
        for case in CASES:
            run_case(G, rel, case)
    finally:
        G.close()
        engine.synchronize()
