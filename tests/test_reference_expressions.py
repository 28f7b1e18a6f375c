"""Per-row expression parity against the UNMODIFIED reference's own expression classes.

tests/golden/reference_expressions.json was written by tests/golden/make_expr_golden.cpp, a program linked against the
reference engine's libraries: 69 predicates / scalars built from ScalarAttribute, ScalarLiteral, ScalarUnaryExpression,
ScalarBinaryExpression, ScalarSharedExpression, ComparisonPredicate, NegationPredicate, Conjunction / DisjunctionPredicate,
evaluated by the reference's vectorised Predicate::getAllMatches / Scalar::getAllValues over 512 tuples (with NaN, +-0.0,
+-inf and a subnormal among the doubles), and lowered from their getProto() by the in-tree binding
(quickstep_b200/host/intree/ProtoLowering.hpp) into qs_node arrays.

Here the SAME node arrays are evaluated by the oracle (CPU, `-m "not gpu"`) and by the CUDA path through the C ABI
(`-m gpu`), and the results must be the reference's: match sets exact; scalar values bit for bit (SURVEY.md 8a rows P1,
P2, E1), except that two NaNs count as equal whatever their sign / payload bits (x86 SSE produces the negative default NaN
for 0.0 / 0.0, the GPU the positive one; the reference itself never looks at those bits).
"""
import json
import os

import numpy as np
import pytest

from backends import GpuBackend, OracleBackend
from quickstep_b200 import capi as A
from quickstep_b200.expr import ExprSet
from quickstep_b200.table import Column, HostTable, np_dtype

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_expressions.json")))
CASES = GOLDEN["cases"]
N = GOLDEN["n_rows"]
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


def test_golden_file_is_what_the_generator_writes():
    assert len(CASES) == 69 and N == 512
    assert sum(c["kind"] == "predicate" for c in CASES) == 39
    d = the_table().col("d").data
    assert np.isnan(d).any() and np.isinf(d).any() and (np.signbit(d) & (d == 0)).any()      # the special values are there


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_gives_the_reference_classes_results(oracle, case):
    run_case(OracleBackend(), the_table(), case)


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
