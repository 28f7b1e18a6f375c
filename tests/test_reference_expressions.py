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

`aggregates` (66 cases): AggregateFunction::createHandle -> accumulateValueAccessor over two halves of the relation ->
mergeStates -> finalize, for SUM / AVG / MIN / MAX / COUNT over every numeric type, expressions, COUNT(*) and COUNT(CHAR),
NOT NULL and NULL-able (row A6).  `lip_filters` (24 cases): LIPFilterFactory::ReconstructFromProto, insertValueAccessor
over the tuples a predicate keeps, filterBatch over all tuples -- exact filters (INT / LONG, anti, probe values outside
the range) and identity-hash filters (negative values) (rows L1, L2, L4).

`hash_partitions` (15 cases): HashPartitionSchemeHeader::getPartitionId for INT (negative values) and LONG keys over 2 / 4 /
7 / 8 / 13 partitions -- where PartitionAwareInsertDestination sends each tuple (row f4).

`dictionaries` (205 comparisons): the reference's CompressionDictionaryBuilder builds a block dictionary from a column and
CompressionDictionary::getLimitCodesForComparisonTyped names the code range of `attribute <cmp> literal` (literals of the
attribute's type and of other types; CHAR literals shorter and longer than the attribute) -- checked against the PRODUCT's
qsgpu_dictionary_code_range (host arithmetic of libqsgpu.so, no device needed), the function every comparison on a coded
attribute goes through (row f2).

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
# "extra_" cases were added after the device runs of profiles/r4a-r4d: the oracle is held against them, the -m gpu tests
# keep the case list that ran on the hardware
EXTRA_CASES = [c for c in GOLDEN["cases"] if c["name"].startswith("extra_")]
CASES = [c for c in GOLDEN["cases"] if not c["nullable"] and not c["name"].startswith("extra_")]
NULL_CASES = [c for c in GOLDEN["cases"] if c["nullable"]]
AGGREGATES = GOLDEN["aggregates"]
LIP_FILTERS = GOLDEN["lip_filters"]
HASH_PARTITIONS = GOLDEN["hash_partitions"]
DICTIONARIES = GOLDEN["dictionaries"]
RTOL = 1e-9            # double SUM / AVG on the device (summation order differs; BASELINE.json north_star)
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
    assert len(CASES) == 69 and len(NULL_CASES) == 32 and len(EXTRA_CASES) == 17 and N == 512
    assert len(AGGREGATES) == 66 and len(LIP_FILTERS) == 24
    assert GOLDEN["nullable_attributes"] == [0, 2, 4, 7, 9]
    assert set(np.nonzero([(NULLS >> np.uint64(a) & np.uint64(1)).any() for a in range(RID)])[0].tolist()) == {0, 2, 4, 7, 9}
    assert sum(c["kind"] == "predicate" for c in CASES) == 39
    d = the_table().col("d").data
    assert np.isnan(d).any() and np.isinf(d).any() and (np.signbit(d) & (d == 0)).any()      # the special values are there


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_gives_the_reference_classes_results(oracle, case):
    run_case(OracleBackend(), the_table(), case)


@pytest.mark.parametrize("case", EXTRA_CASES, ids=[c["name"] for c in EXTRA_CASES])
def test_oracle_gives_the_reference_classes_results_on_more_type_mixes(oracle, case):
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


# ------------------------------------------------------------------------------------------- aggregation handles
_RESULT_DTYPE = {A.QS_INT: "<i4", A.QS_LONG: "<i8", A.QS_FLOAT: "<f4", A.QS_DOUBLE: "<f8"}


def reference_result(case):
    if case["result_null"]:
        return None
    return np.frombuffer(bytes.fromhex(case["result"]), dtype=_RESULT_DTYPE[case["result_type"]])[0]


def check_aggregate(case, got, got_null, exact_doubles):
    want = reference_result(case)
    assert got_null == (want is None), (case["name"], got, got_null, want)
    if want is None:
        return
    assert np.asarray(got).dtype == want.dtype, (case["name"], np.asarray(got).dtype, want.dtype)
    if want.dtype.kind == "f" and case["function"] in (A.QS_AGG_SUM, A.QS_AGG_AVG) and not exact_doubles:
        g, w = float(got), float(want)
        assert (g != g and w != w) or g == w or abs(g - w) <= RTOL * max(abs(g), abs(w)), (case["name"], g, w)
    else:
        assert got == want or (got != got and want != want), (case["name"], got, want)


@pytest.mark.parametrize("case", [c for c in AGGREGATES if not c["nullable"]], ids=[c["name"] for c in AGGREGATES if not c["nullable"]])
def test_oracle_gives_the_reference_handles_results(oracle, case):
    """Two 256-row work orders merged in order, like the generator's two accumulate + mergeStates: bit-exact, doubles too."""
    oracle.set_block_rows(256)
    try:
        es, _ = expr_set(case)
        out = OracleBackend().aggregate(the_table(), es, -1, [(case["function"], case["root"])], [], A.QS_AGG_SINGLE_STATE, [])
        check_aggregate(case, out.values[0][0], bool(out.null_mask & 1), exact_doubles=True)
    finally:
        oracle.set_block_rows(65536)


@pytest.mark.parametrize("case", [c for c in AGGREGATES if c["nullable"]], ids=[c["name"] for c in AGGREGATES if c["nullable"]])
def test_null_oracle_gives_the_reference_handles_results(oracle, case):
    es, _ = expr_set(case)
    (value, is_null), = NO.aggregate(es, -1, [(case["function"], case["root"])], None, the_table(), NULLS)[None]
    want = reference_result(case)
    assert is_null == (want is None), (case["name"], value, is_null, want)
    if want is not None:
        if want.dtype.kind == "f":
            g, w = float(value), float(want)          # numpy's pairwise sum is not the handle's sequential one
            assert (g != g and w != w) or g == w or abs(g - w) <= RTOL * max(abs(g), abs(w)), (case["name"], g, w)
        else:
            assert int(value) == int(want), (case["name"], value, want)


def argument_is_nullable(es, root):
    """Does the argument read a NULL-able attribute (what AggregateFunction::createHandle learns from the argument type)."""
    if root < 0:
        return False
    n = es.nodes[root]
    if n.kind == A.QS_N_ATTRIBUTE:
        return n.a in GOLDEN["nullable_attributes"]
    if n.kind == A.QS_N_LITERAL:
        return False
    return argument_is_nullable(es, n.a) or (n.kind == A.QS_N_BINARY and argument_is_nullable(es, n.b))


@pytest.mark.gpu
def test_cuda_path_gives_the_reference_handles_results(engine):
    """All 66 aggregate cases as SINGLE_STATE aggregations through qsgpu_agg_create / run (two work orders) / finalize."""
    from test_gpu_nulls import nullable_relation, run_agg
    t = the_table()
    G = GpuBackend(engine)
    nrel = nullable_relation(engine, the_table(), NULLS)
    try:
        rel = G.relation(t)
        for case in AGGREGATES:
            es, _ = expr_set(case)
            aggs = [(case["function"], case["root"])]
            if not case["nullable"]:
                out = G.aggregate(rel, es, -1, aggs, [], A.QS_AGG_SINGLE_STATE, [], row_ranges=[(0, 256), (256, 512)])
                check_aggregate(case, out.values[0][0], bool(out.null_mask & 1), exact_doubles=False)
            else:
                cols, out_nulls, mask = run_agg(engine, nrel, A.QS_AGG_SINGLE_STATE, es, -1, aggs, [], [],
                                                [0] if argument_is_nullable(es, case["root"]) else [], t, work_orders=2)
                check_aggregate(case, cols[0][0], bool(mask & 1), exact_doubles=False)
                assert bool(int(out_nulls[0]) & 1) == bool(mask & 1)
    finally:
        nrel.destroy()
        G.close()
        engine.synchronize()


# ------------------------------------------------------------------------------------------- LIP filters
def run_lip_case(backend, rel, case):
    es, rid = expr_set(case)
    attr_type = A.QS_INT if case["attribute_size"] == 4 else A.QS_LONG
    kind = A.QS_LIP_BITVECTOR_EXACT if case["exact"] else A.QS_LIP_SINGLE_IDENTITY_HASH
    lip = backend.make_lip(kind, attr_type, case["min_value"], case["max_value"], case["cardinality"], case["is_anti"])
    backend.build_lip(rel, es, case["root"], None, [(lip, case["build_attr"])])
    return lip, es, rid


def check_passes(case, ids):
    got = np.zeros(N, dtype=bool)
    assert len(np.unique(ids)) == len(ids)
    got[ids] = True
    want = np.array([ch == "1" for ch in case["passes"]])
    assert int(want.sum()) == case["n_passes"]
    assert (got == want).all(), (case["name"], np.nonzero(got != want)[0][:10])


@pytest.mark.parametrize("case", [c for c in LIP_FILTERS if not c["nullable"]], ids=[c["name"] for c in LIP_FILTERS if not c["nullable"]])
def test_oracle_gives_the_reference_filters_results(oracle, case):
    B = OracleBackend()
    t = the_table()
    lip, es, rid = run_lip_case(B, t, case)
    out = B.select(t, es, -1, [(lip, case["probe_attr"])], [rid], [(A.QS_INT, 4)])
    check_passes(case, out.columns[0].data)


@pytest.mark.parametrize("case", [c for c in LIP_FILTERS if c["nullable"]], ids=[c["name"] for c in LIP_FILTERS if c["nullable"]])
def test_null_rule_for_filters_gives_the_reference_filters_results(oracle, case):
    """A NULL value neither enters a filter nor passes one -- not even an anti filter (BitVectorExactFilter.hpp:113-146):
    the oracle builds from the rows whose build attribute is not NULL and the NULL probe rows are dropped afterwards."""
    B = OracleBackend()
    t = the_table()
    es, rid = expr_set(case)
    isnull = lambda a: ((NULLS >> np.uint64(a)) & np.uint64(1)).astype(bool)
    keep = NO.predicate(es, case["root"], t, NULLS) & ~isnull(case["build_attr"])
    assert int(NO.predicate(es, case["root"], t, NULLS).sum()) == case["n_built_from"]
    idx = np.nonzero(keep)[0]
    built_from = HostTable("b", [Column(c.name, c.type, c.data[idx], c.width) for c in t.columns])
    attr_type = A.QS_INT if case["attribute_size"] == 4 else A.QS_LONG
    kind = A.QS_LIP_BITVECTOR_EXACT if case["exact"] else A.QS_LIP_SINGLE_IDENTITY_HASH
    lip = B.make_lip(kind, attr_type, case["min_value"], case["max_value"], case["cardinality"], case["is_anti"])
    B.build_lip(built_from, None, -1, None, [(lip, case["build_attr"])])
    out = B.select(t, es, -1, [(lip, case["probe_attr"])], [rid], [(A.QS_INT, 4)])
    ids = out.columns[0].data
    check_passes(case, ids[~isnull(case["probe_attr"])[ids]])


@pytest.mark.gpu
def test_cuda_path_gives_the_reference_filters_results(engine):
    """The 24 LIP cases: qsgpu_lip_create from the proto's fields, qsgpu_build_lip_filter, probe inside qsgpu_select."""
    from test_gpu_nulls import nullable_relation
    G = GpuBackend(engine)
    nrel = nullable_relation(engine, the_table(), NULLS)
    try:
        rel = G.relation(the_table())
        for case in LIP_FILTERS:
            src = nrel if case["nullable"] else rel
            lip, es, rid = run_lip_case(G, src, case)
            out = engine.Relation.create([(A.QS_INT, 4)], N)
            try:
                engine.select(src, es, -1, [(lip, case["probe_attr"])], [rid], out)
                ids = out.read_all()[0]
            finally:
                out.destroy()
            check_passes(case, ids)
    finally:
        nrel.destroy()
        G.close()
        engine.synchronize()


# ------------------------------------------------------------------------------------------- hash partitioning
def partition_function(values: np.ndarray, n_parts: int) -> np.ndarray:
    """catalog/PartitionSchemeHeader.hpp:200-214 over TypedValue::getHash of one INT / LONG key: the bit pattern of the
    inline value with the rest of the 8-byte union zeroed (types/TypedValue.hpp:95-100) -- a negative INT is NOT
    sign-extended -- then `& (n - 1)` for a power of two, `% n` otherwise."""
    h = values.view(np.uint32).astype(np.uint64) if values.dtype.itemsize == 4 else values.view(np.uint64)
    n = np.uint64(n_parts)
    return (h & (n - np.uint64(1))) if n_parts & (n_parts - 1) == 0 else (h % n)


@pytest.mark.parametrize("case", HASH_PARTITIONS, ids=[f"attr{c['attr']}_n{c['n_parts']}" for c in HASH_PARTITIONS])
def test_partition_function_is_the_reference_headers(case):
    vals = the_table().columns[case["attr"]].data
    assert (partition_function(vals, case["n_parts"]) == np.array(case["partition_of_row"], dtype=np.uint64)).all()


@pytest.mark.gpu
def test_cuda_path_partitions_like_the_reference_header(engine):
    """qsgpu_hash_partition (K8 behind PartitionAwareInsertDestination): every tuple lands in the partition
    HashPartitionSchemeHeader::getPartitionId names, and the output is a permutation of the input."""
    # a partition moves every attribute of its input (at most 12): the three key attributes and the row id
    t = the_table()
    key_attrs = sorted({c["attr"] for c in HASH_PARTITIONS})
    narrow = HostTable("keys", [t.columns[a] for a in key_attrs] + [t.columns[RID]])
    G = GpuBackend(engine)
    try:
        rel = G.relation(narrow)
        for case in HASH_PARTITIONS:
            out = engine.Relation.create(rel.schema, N)
            try:
                off = engine.hash_partition(rel, key_attrs.index(case["attr"]), case["n_parts"], out)
                rids = out.read(len(key_attrs))
            finally:
                out.destroy()
            assert len(off) == case["n_parts"] + 1 and off[0] == 0 and off[-1] == N
            assert sorted(rids.tolist()) == list(range(N))
            want = np.array(case["partition_of_row"])
            for p in range(case["n_parts"]):
                part = rids[int(off[p]):int(off[p + 1])]
                assert (want[part] == p).all(), (case["attr"], case["n_parts"], p)
    finally:
        G.close()
        engine.synchronize()


# ------------------------------------------------------------------------------------------- dictionary limit codes
@pytest.mark.parametrize("d", DICTIONARIES, ids=[d["name"] for d in DICTIONARIES])
def test_code_range_is_the_reference_dictionarys_limit_codes(d):
    """qsgpu_dictionary_code_range (libqsgpu.so, host-only) against CompressionDictionary::getLimitCodesForComparisonTyped:
    the same set of codes for =, <, <=, >, >= -- and for <> as the complement of = (the reference's caller does that:
    storage/CompressedTupleStorageSubBlock.cpp:213-236)."""
    import ctypes as C
    lib = A.load()
    w, n = d["width"], d["n_codes"]
    entries = np.frombuffer(bytes.fromhex(d["entries"]), dtype=np.uint8).copy()
    assert len(entries) == n * w
    col = the_table().columns[d["attr"]].data
    assert n == len(np.unique(col))                                   # the builder's dictionary is the column's distinct values
    checked = 0
    for c in d["comparisons"]:
        es, _rid = expr_set(c)
        view = es.c()
        want = np.zeros(n, dtype=bool)
        want[c["first"]:c["second"]] = True
        for cmp, expect in ((c["cmp"], want),) + (((A.QS_NE, ~want),) if c["cmp"] == A.QS_EQ else ()):
            first, count, neg = C.c_uint32(0), C.c_uint32(0), C.c_int(0)
            A.check(lib.qsgpu_dictionary_code_range(d["type"], w if d["type"] == A.QS_CHAR else 0, entries.ctypes.data, n, cmp,
                                                    C.byref(view.nodes[c["root"]]), view.str_pool, view.str_pool_bytes, C.byref(first),
                                                    C.byref(count), C.byref(neg)))
            got = np.zeros(n, dtype=bool)
            got[first.value:first.value + count.value] = True
            if neg.value:
                got = ~got
            assert (got == expect).all(), (d["name"], c["literal"], cmp, first.value, count.value, neg.value, c["first"], c["second"])
            checked += 1
    assert checked == len(d["comparisons"]) * 6 // 5
