"""The reference's OWN plans for TPC-H Q1 / Q6 / Q3, lowered by the in-tree binding, give the engine's answers.

tests/golden/reference_plans.json was written by tests/golden/make_plan_golden.cpp, a program linked against the
unmodified reference: its parser read benchmarks/tpch/queries/{01,03,06}.sql, its optimizer and ExecutionGenerator planned
them (catalog statistics of dbgen -s 0.01), and the serialization::QueryContext they produced -- aggregation states,
predicates, scalar groups, LIP filters and deployments, join hash tables, sort configurations -- was lowered by
quickstep_b200/host/intree/{ProtoLowering,QueryContextLowering}.hpp into the C ABI's descriptions.  Every operator of the
DAG also described one work order in the reference's own serialized form (relation ids + QueryContext indices).

This test is the "no plan changes" claim of BASELINE.json's north_star, executed: a small interpreter walks the DAG in
operator order, gives each operator the lowered objects its work-order proto names, runs it with the oracle over dbgen's
SF0.01 relations, and compares the final relation with what the unmodified engine printed for the same SQL
(tests/golden/reference_engine_results.json): keys, counts and dates exact, double sums to 1e-9.

What the real plans contain that hand-written descriptions missed (and the lowering now handles): Q6's
`date '1994-01-01' + interval '1' year` arrives as an un-folded ScalarBinaryExpression over two literals (folded with
the reference's own operations, FoldStaticScalar); Q3's 'BUILDING' is a VARCHAR(8) literal compared with a CHAR(10)
attribute; Q1's AVGs are rewritten into SUM / COUNT with a SelectOperator dividing afterwards (ReuseAggregateExpressions).
"""
import json
import os
import re

import numpy as np
import pytest

import tpch_data as D
from backends import OracleBackend, agg_out_types
from quickstep_b200 import capi as A
from quickstep_b200.expr import ExprSet
from quickstep_b200.table import Column, HostTable, np_dtype
from ref_golden import ENGINE, close

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLANS = {p["query"]: p for p in json.load(open(os.path.join(ROOT, "tests", "golden", "reference_plans.json")))["plans"]}
_FIELD = re.compile(r"^\[quickstep\.serialization\.\w+\.(\w+)\]: (\S+)$")


def work_order(op):
    """The operator's serialized work order (protobuf text form) -> {field: value | [values]}."""
    out = {}
    for line in op["work_order"].split("\n"):
        line = line.strip()
        m = _FIELD.match(line)
        if m:
            k, v = m.group(1), m.group(2)
            v = {"true": True, "false": False}.get(v, v)
            if isinstance(v, str) and re.fullmatch(r"-?\d+", v):
                v = int(v)
            if k in out:
                out[k] = (out[k] if isinstance(out[k], list) else [out[k]]) + [v]
            else:
                out[k] = v
        elif line.startswith("work_order_type:"):
            out["type"] = line.split(":")[1].strip()
    for k in ("simple_selection", "join_key_attributes", "lip_filter_indexes"):
        if k in out and not isinstance(out[k], list):
            out[k] = [out[k]]
    return out


def expr_set(obj) -> ExprSet:
    es = ExprSet()
    for kind, op, typ, width, a, b, lit in obj["nodes"]:
        n = A.qs_node()
        n.kind, n.op, n.type, n.width, n.a, n.b = kind, op, typ, width, a, b
        n.lit.i64 = int.from_bytes(bytes.fromhex(lit), "little", signed=True)
        es.nodes.append(n)
    es.pool = bytearray(bytes.fromhex(obj["pool"]))
    return es


def append_exprs(es: ExprSet, obj):
    """Appends a lowered node array to `es` (ExprSet::append in host/ExprSet.hpp: a work order that evaluates several
    QueryContext entries passes ONE qs_expr_set); -> index shift of the appended nodes."""
    ges, base, pool_base = expr_set(obj), len(es.nodes), len(es.pool)
    for n in ges.nodes:
        if n.kind == A.QS_N_LITERAL:
            if n.type == A.QS_CHAR:
                n.lit.pool_offset += pool_base
        elif n.kind != A.QS_N_ATTRIBUTE:
            n.a += base
            if n.kind in (A.QS_N_BINARY, A.QS_N_COMPARISON, A.QS_N_CONJUNCTION, A.QS_N_DISJUNCTION):
                n.b += base
        es.nodes.append(n)
    es.pool += ges.pool
    es._c = None
    return base


class Interpreter:
    def __init__(self, plan, tables, backend=None, limit=None):
        # limit: the query's LIMIT (SortMergeRunOperator's top_k is not part of any serialized entry); None sorts everything
        self.plan, self.B, self.limit = plan, backend or OracleBackend(), limit
        self.schema = {r["id"]: [(t, w) for _n, t, w in r["attributes"]] for r in plan["relations"]}
        self.rel = {r["id"]: tables[r["name"]] for r in plan["relations"] if not r["temporary"]}
        self.lips = []
        for f in plan["lip_filters"]:
            assert f["kind"] == A.QS_LIP_BITVECTOR_EXACT
            self.lips.append(self.B.make_lip(f["kind"], A.QS_INT if f["attribute_size"] == 4 else A.QS_LONG, f["min_value"], f["max_value"], 0, f["is_anti"]))
        self.agg, self.built, self.trace, self.cardinality, self.nulls = {}, {}, [], {}, {}

    def destination(self, index):
        return self.plan["insert_destinations"][index]["relation_id"]

    def store(self, rel_id, cols):
        schema = self.schema[rel_id]
        assert len(cols) == len(schema), (rel_id, len(cols), len(schema))
        self.rel[rel_id] = HostTable(f"r{rel_id}", [Column(f"c{i}", t, np.ascontiguousarray(c), w) for i, ((t, w), c) in enumerate(zip(schema, cols))])

    def lip_refs(self, deployment, action):
        if deployment < 0:
            return None
        return [(self.lips[f], attr) for f, attr in self.plan["lip_filter_deployments"][deployment][action]] or None

    def order(self):
        """The DAG's edges decide who runs first (a LIP filter's builder precedes its probers whatever their indices):
        Kahn's algorithm, lowest operator index first among the ready ones."""
        ops = self.plan["operators"]
        waits = {o["index"]: 0 for o in ops}
        for o in ops:
            for consumer, _breaking in o["dependents"]:
                waits[consumer] += 1
        ready, out = sorted(i for i, w in waits.items() if w == 0), []
        while ready:
            i = ready.pop(0)
            out.append(ops[i])
            for consumer, _breaking in ops[i]["dependents"]:
                waits[consumer] -= 1
                if waits[consumer] == 0:
                    ready.append(consumer)
                    ready.sort()
        assert len(out) == len(ops)
        return out

    def run(self):
        for op in self.order():
            wo = work_order(op)
            kind = wo.get("type")
            self.trace.append((op["index"], op["name"], kind))
            if kind == "SELECT":
                src = self.rel[wo["relation_id"]]
                if wo["predicate_index"] >= 0:
                    pred = self.plan["predicates"][wo["predicate_index"]]
                    es, root = expr_set(pred), pred["root"]
                else:
                    es, root = ExprSet(), -1
                if wo["simple_projection"]:
                    roots = [es.attr(a, *self.schema[wo["relation_id"]][a]) for a in wo["simple_selection"]]
                else:
                    grp = self.plan["scalar_groups"][wo["selection_index"]]
                    base = append_exprs(es, grp)
                    roots = [r + base for r in grp["roots"]]
                dest = self.destination(wo["insert_destination_index"])
                out = self.B.select(src, es, root, self.lip_refs(wo["lip_deployment_index"], "probe"), roots, self.schema[dest])
                self.store(dest, [c.data for c in out.columns])
                if wo["relation_id"] in self.nulls:        # a scalar over a NULL is NULL (no predicate in the plans that get here)
                    import qs_null_oracle as NO
                    assert root == -1
                    nl = np.zeros(out.n_rows, dtype=np.uint64)
                    for j, r in enumerate(roots):
                        nl |= NO.null_of(es, r, self.nulls[wo["relation_id"]]).astype(np.uint64) << np.uint64(j)
                    self.nulls[dest] = nl
            elif kind == "BUILD_LIP_FILTER":
                pred = self.plan["predicates"][wo["build_side_predicate_index"]]
                self.B.build_lip(self.rel[wo["relation_id"]], expr_set(pred), pred["root"], self.lip_refs(wo["lip_deployment_index"], "probe"),
                                 self.lip_refs(wo["lip_deployment_index"], "build"))
            elif kind == "BUILD_HASH":
                pred = wo["build_predicate_index"]
                self.built[wo["join_hash_table_index"]] = (wo["relation_id"], wo["join_key_attributes"], None if pred in (-1, 4294967295) else pred)
                build_refs = self.lip_refs(wo["lip_deployment_index"], "build")
                if build_refs:      # BuildHashWorkOrder::execute also inserts the build keys into its LIP filters (BuildHashOperator.cpp:186-197)
                    assert pred in (-1, 4294967295)
                    self.B.build_lip(self.rel[wo["relation_id"]], None, -1, None, build_refs)
                self.cardinality["build_rows"] = self.rel[wo["relation_id"]].n_rows
            elif kind == "HASH_JOIN":
                join_type = {"HASH_INNER_JOIN": A.QS_JOIN_INNER, "HASH_SEMI_JOIN": A.QS_JOIN_LEFT_SEMI, "HASH_ANTI_JOIN": A.QS_JOIN_LEFT_ANTI,
                             "HASH_OUTER_JOIN": A.QS_JOIN_LEFT_OUTER}[wo["hash_join_work_order_type"]]
                build_rel, build_keys, build_pred = self.built[wo["join_hash_table_index"]]
                assert build_rel == wo["build_relation_id"]
                build, probe = self.rel[build_rel], self.rel[wo["probe_relation_id"]]
                probe_keys = wo["join_key_attributes"]
                assert len(build_keys) == len(probe_keys)
                es = ExprSet()
                bp = rp = -1
                if build_pred is not None:
                    # evaluated over the build relation alone (BuildHashWorkOrder::execute): the join sides the optimizer left
                    # on a pushed-down predicate are dropped (AttributeTypes::single_relation in the in-tree lowering)
                    bp = append_exprs(es, self.plan["predicates"][build_pred]) + self.plan["predicates"][build_pred]["root"]
                    for n in es.nodes:
                        if n.kind == A.QS_N_ATTRIBUTE:
                            n.b = 0
                if wo["residual_predicate_index"] >= 0:
                    res = self.plan["predicates"][wo["residual_predicate_index"]]
                    rp = append_exprs(es, res) + res["root"]
                grp = self.plan["scalar_groups"][wo["selection_index"]]
                base = append_exprs(es, grp)
                roots = [r + base for r in grp["roots"]]
                if len(build_keys) == 1:
                    bk, pk = build_keys[0], probe_keys[0]
                else:
                    # composite key (HashTable::putValueAccessorCompositeKey, storage/HashTable.hpp:1469): two INT attributes
                    # packed into one LONG key column appended to each side -- what qsgpu_join_*_composite does on the device
                    assert len(build_keys) == 2

                    def with_packed(t, keys):
                        hi = t.columns[keys[0]].data.astype(np.int64) << np.int64(32)
                        lo = t.columns[keys[1]].data.astype(np.int64) & np.int64(0xFFFFFFFF)
                        return HostTable(t.name, list(t.columns) + [Column("packed_key", A.QS_LONG, hi | lo)])
                    build, probe = with_packed(build, build_keys), with_packed(probe, probe_keys)
                    bk, pk = len(build.columns) - 1, len(probe.columns) - 1
                dest = self.destination(wo["insert_destination_index"])
                # the build side may carry duplicate keys: size for every pair
                out = self.B.hash_join(build, bp, bk, probe, es, -1, pk, join_type, rp, roots, self.schema[dest],
                                       max(1, probe.n_rows * 8), build_es=es)
                self.store(dest, [c.data for c in out.columns])
                self.cardinality.update(probe_rows=probe.n_rows, join_rows=out.n_rows)
            elif kind == "AGGREGATION":
                st = self.plan["aggregation_states"][wo["aggr_state_index"]]
                es = expr_set(st)
                aggs = [tuple(a) for a in st["aggregates"]]
                key_schema = [(es.nodes[r].type, es.nodes[r].width) for r in st["group_by_roots"]]
                out = self.B.aggregate(self.rel[st["relation_id"]], es, st["predicate_root"], aggs, st["group_by_roots"], st["strategy"], key_schema,
                                       self.lip_refs(wo["lip_deployment_index"], "probe"))
                self.agg[wo["aggr_state_index"]] = (out, key_schema, agg_out_types(es, aggs))
                self.cardinality["groups"] = out.n_groups
            elif kind == "FINALIZE_AGGREGATION":
                out, key_schema, val_types = self.agg[wo["aggr_state_index"]]
                dest = self.destination(wo["insert_destination_index"])
                assert [t for t, _w in self.schema[dest]] == [t for t, _w in key_schema] + [t for t, _w in val_types], (self.schema[dest], key_schema, val_types)
                cols, off = [], 0
                for t, w in key_schema:
                    dt = np_dtype(t, w)
                    cols.append(out.keys[:, off:off + dt.itemsize].copy().view(dt).reshape(-1))
                    off += dt.itemsize
                self.store(dest, cols + list(out.values))
                if out.null_mask:       # an aggregate without GROUP BY that saw no row is NULL (AggregationHandleSum.cpp:134-143)
                    assert not key_schema and out.n_groups == 1
                    self.nulls[dest] = np.array([out.null_mask << len(key_schema)], dtype=np.uint64)
            elif kind == "SORT_RUN_GENERATION":
                cfg = self.plan["sort_configs"][wo["sort_config_index"]]
                src = self.rel[wo["relation_id"]]
                keys = []
                for k, (attr, flags) in zip(cfg["keys"], cfg["qs_sort_keys"]):       # LowerSortConfiguration's qs_sort_key list
                    assert k["relation_id"] == wo["relation_id"] and attr == k["attribute_id"]
                    assert bool(flags & 1) == (not k["ascending"]) and bool(flags & 2) == k["null_first"] and bool(flags & 4) == (not k["null_first"])
                    # NOT NULL keys here: the NULL placement has nothing to order
                    keys.append((attr, bool(flags & 1)))
                top = self.B.topk(src, keys, max(1, min(self.limit or src.n_rows, src.n_rows)))
                self.store(self.destination(wo["insert_destination_index"]), [c.data for c in top.columns])
                self.sorted = self.destination(wo["insert_destination_index"])
            elif op["name"] == "SortMergeRunOperator":
                # one sorted run: merging it is the identity (the top-k LIMIT is applied when the result is compared)
                self.store(op["output_relation"], [c.data for c in self.rel[self.sorted].columns])
            else:
                assert kind in ("DESTROY_AGGREGATION_STATE", "DESTROY_HASH", "INITIALIZE_AGGREGATION", None), (op["name"], kind)
        result = max(i for i in self.rel if self.plan["relations"][[r["id"] for r in self.plan["relations"]].index(i)]["temporary"])
        self.result_nulls = self.nulls.get(result)
        return self.rel[result]


class StagingBackend:
    """The Backend interface of tests/backends.py over HOST tables for a backend that works on staged relations: inputs are
    staged (inner.relation) for every call and outputs come back as host tables, so the interpreter -- which keeps its
    temporaries on the host -- walks the same code over the oracle and over the device
    (tests/test_zz_reference_plans_on_device.py wraps GpuBackend; here it wraps OracleBackend to test the plumbing)."""

    def __init__(self, inner):
        self.G, self.staged = inner, {}

    def up(self, table):
        if id(table) not in self.staged:
            self.staged[id(table)] = (table, self.G.relation(table))          # keeps `table` alive: ids stay unique
        return self.staged[id(table)][1]

    def make_lip(self, *a, **k):
        return self.G.make_lip(*a, **k)

    def build_lip(self, rel, es, pred, probe, build):
        self.G.build_lip(self.up(rel), es if es is not None else ExprSet(), pred, probe, build)

    def select(self, rel, es, pred, probe, roots, out_schema, capacity=None):
        return self.G.select(self.up(rel), es, pred, probe, roots, out_schema, capacity)

    def hash_join(self, build, bes_pred, build_key, probe, es, probe_pred, probe_key, join_type, residual, roots, out_schema, capacity,
                  build_es=None, probe_lips=None):
        return self.G.hash_join(self.up(build), bes_pred, build_key, self.up(probe), es, probe_pred, probe_key, join_type, residual, roots,
                                out_schema, capacity, build_es=build_es, probe_lips=probe_lips)

    def aggregate(self, rel, es, pred, aggregates, group_roots, strategy, key_schema, probe=None):
        # table sizing: a compact-key state keeps the optimizer's small-group estimate (above 256 the library would take
        # the hash strategy); hash / dense tables are sized for the rows they may hold
        estimated = 16 if strategy == A.QS_AGG_COMPACT_KEY else max(16, rel.n_rows)
        return self.G.aggregate(self.up(rel), es, pred, aggregates, group_roots, strategy, key_schema, probe, estimated=estimated)

    def topk(self, rel, keys, limit):
        return self.G.topk(self.up(rel), keys, limit)

    def close(self):
        if hasattr(self.G, "close"):
            self.G.close()


@pytest.fixture(scope="module")
def tables(oracle):
    return D.golden_tables()


@pytest.mark.parametrize("query", ["q6", "q1", "q3"])
def test_staging_backend_plumbing(tables, query):
    """The adapter the device run uses, over the oracle: same answers as the direct run."""
    out = Interpreter(PLANS[query], tables, StagingBackend(OracleBackend())).run()
    {"q6": check_q6, "q1": check_q1, "q3": check_q3}[query](out)


def test_plans_are_the_optimizers(tables):
    """Shapes the rest of the repository assumes about the reference's plans, now read off the optimizer."""
    q1, q6, q3 = PLANS["q1"], PLANS["q6"], PLANS["q3"]
    names = lambda p: [o["name"] for o in p["operators"] if "DropTable" not in o["name"]]
    assert names(q6) == ["AggregationOperator", "FinalizeAggregationOperator", "DestroyAggregationStateOperator"]
    assert names(q1) == ["AggregationOperator", "FinalizeAggregationOperator", "DestroyAggregationStateOperator", "SelectOperator",
                         "SortRunGenerationOperator", "SortMergeRunOperator"]
    assert names(q3) == ["SelectOperator", "BuildLIPFilterOperator", "SelectOperator", "BuildHashOperator", "HashJoinOperator", "DestroyHashOperator",
                         "AggregationOperator", "FinalizeAggregationOperator", "DestroyAggregationStateOperator", "SortRunGenerationOperator",
                         "SortMergeRunOperator", "SelectOperator"]
    # Q6: no GROUP BY -> single state; Q1: (CHAR(1), CHAR(1)) keys -> ThreadPrivateCompactKey; Q3: SeparateChaining
    assert q6["aggregation_states"][0]["strategy"] == A.QS_AGG_SINGLE_STATE
    assert q1["aggregation_states"][0]["hash_table_impl_type"] == 4 and q1["aggregation_states"][0]["strategy"] == A.QS_AGG_COMPACT_KEY
    assert q3["aggregation_states"][0]["strategy"] == A.QS_AGG_SEPARATE_CHAINING
    # Q1's three AVGs became SUM / COUNT(*): five SUMs (the fifth is l_discount's) and one COUNT(*)
    assert [a[0] for a in q1["aggregation_states"][0]["aggregates"]] == [A.QS_AGG_SUM] * 5 + [A.QS_AGG_COUNT]
    # Q3: one exact filter over c_custkey's range, built from customer, probed by the orders select on o_custkey
    assert q3["lip_filters"] == [{"kind": A.QS_LIP_BITVECTOR_EXACT, "min_value": 1, "max_value": 1500, "attribute_size": 4, "is_anti": False}]
    assert q3["lip_filter_deployments"] == [{"build": [[0, 0]], "probe": []}, {"build": [], "probe": [[0, 1]]}]
    # Q6's interval arithmetic was folded into one DATE literal by the lowering: no node of an unsupported type is left
    lits = [n for n in q6["aggregation_states"][0]["nodes"] if n[0] == A.QS_N_LITERAL and n[2] == A.QS_DATE]
    assert sorted(int.from_bytes(bytes.fromhex(n[6])[:4], "little") for n in lits) == [1994, 1995]


def check_q6(out):
    want = ENGINE["sf0.01"]["q6"]["rows"]
    assert out.n_rows == 1 and close(out.columns[0].data[0], float(want[0][0])), (out.columns[0].data, want)


def check_q1(out):
    want = ENGINE["sf0.01"]["q1"]["rows"]
    assert out.n_rows == len(want) == 4 and len(out.columns) == 10
    for i, w in enumerate(want):
        assert out.columns[0].data[i].decode() == w[0] and out.columns[1].data[i].decode() == w[1]
        for j in range(2, 9):
            assert close(out.columns[j].data[i], float(w[j])), (i, j, out.columns[j].data[i], w[j])
        assert int(out.columns[9].data[i]) == int(w[9])


def check_q3(out):
    want = ENGINE["sf0.01"]["q3"]["rows"]
    assert out.n_rows >= len(want) == 10 and len(out.columns) == 4
    for i, w in enumerate(want):            # LIMIT 10 of the SortMergeRunOperator
        d = out.columns[2].data[i]
        assert int(out.columns[0].data[i]) == int(w[0]) and "%04d-%02d-%02d" % (d["year"], d["month"], d["day"]) == w[2]
        assert int(out.columns[3].data[i]) == int(w[3])
        assert close(out.columns[1].data[i], float(w[1])), (i, out.columns[1].data[i], w[1])


def test_q6_plan_gives_the_engines_answer(tables):
    check_q6(Interpreter(PLANS["q6"], tables).run())


def test_q1_plan_gives_the_engines_answer(tables):
    check_q1(Interpreter(PLANS["q1"], tables).run())


def test_q3_plan_gives_the_engines_answer(tables):
    check_q3(Interpreter(PLANS["q3"], tables).run())


@pytest.mark.parametrize("which", ["q3_sf10", "q3_sf100"])
def test_q3_plan_of_the_benchmarked_scale_is_the_hand_built_dag(tables, which):
    """With SF10 / SF100 statistics the optimizer adds the second LIP filter -- o_orderkey's range, built by the
    BuildHashOperator, probed by the lineitem SelectOperator -- which is the DAG quickstep_b200/tpch.py's Q3Plan and
    host/TpchPlans.cpp hand-build for bench.py.  The optimizer's plan (its filters sized for the larger key ranges) and the
    hand-built one run over the same SF0.01 relations: same rows through every operator, same answer, the engine's."""
    import oracle_tpch as OT
    plan = PLANS[which]
    assert [f["max_value"] for f in plan["lip_filters"]] == [{"q3_sf10": 1500000, "q3_sf100": 15000000}[which], {"q3_sf10": 60000000, "q3_sf100": 600000000}[which]]
    assert plan["lip_filter_deployments"] == [{"build": [[0, 0]], "probe": []}, {"build": [[1, 0]], "probe": []}, {"build": [], "probe": [[1, 0]]},
                                              {"build": [], "probe": [[0, 1]]}]
    if which == "q3_sf100":       # a 600 M-bit filter is 75 MB of words on the host: the SF10 plan is the one executed
        return
    it = Interpreter(plan, tables)
    names = [o["name"] for o in it.order() if "DropTable" not in o["name"] and "Destroy" not in o["name"]]
    assert names == ["BuildLIPFilterOperator", "SelectOperator", "BuildHashOperator", "SelectOperator", "HashJoinOperator", "AggregationOperator",
                     "FinalizeAggregationOperator", "SortRunGenerationOperator", "SortMergeRunOperator", "SelectOperator"]
    out = it.run()
    info = {}
    hand = OT.q3(tables, D.q3_stats(tables), info=info)
    assert it.cardinality == dict(build_rows=info["t2_rows"], probe_rows=info["t0_rows"], join_rows=info["t4_rows"], groups=info["groups"])
    want = ENGINE["sf0.01"]["q3"]["rows"]
    for i, (h, w) in enumerate(zip(hand, want)):
        d = out.columns[2].data[i]
        assert int(out.columns[0].data[i]) == h[0] == int(w[0]) and (int(d["year"]), int(d["month"]), int(d["day"])) == h[2]
        assert int(out.columns[3].data[i]) == h[3] == int(w[3])
        assert out.columns[1].data[i] == h[1] and close(h[1], float(w[1]))        # the two plans add the same values in the same order


def test_q1_q6_plans_do_not_depend_on_the_scale(tables):
    for q in ("q1", "q6"):
        small, big = PLANS[q], PLANS[q + "_sf100"]
        for key in ("aggregation_states", "predicates", "scalar_groups", "sort_configs"):
            assert small[key] == big[key], (q, key)
        assert [o["name"] for o in small["operators"]] == [o["name"] for o in big["operators"]]


def test_hand_built_q1_q6_trees_are_the_optimizers(tables):
    """quickstep_b200/tpch.py's Q1Plan / Q6Plan (what bench.py's operator DAGs evaluate) against the optimizer's lowered
    aggregation states: same aggregate functions in the same order, same strategy, and -- row by row over dbgen's
    lineitem -- the same predicate bits and the same argument values."""
    import qs_oracle as O
    from quickstep_b200 import tpch as T
    li = tables["lineitem"]
    for q, hand in (("q1", T.Q1Plan()), ("q6", T.Q6Plan())):
        st = PLANS[q]["aggregation_states"][0]
        es = expr_set(st)
        assert [a[0] for a in st["aggregates"]] == [f for f, _r in hand.aggregates]
        assert len(st["group_by_roots"]) == len(getattr(hand, "group_by", []))
        _n, got = O.predicate(es, st["predicate_root"], li)
        _n, want = O.predicate(hand.es, hand.pred, li)
        assert (got == want).all()
        for (f, r), (_f, hr) in zip(st["aggregates"], hand.aggregates):
            if r >= 0:
                a, b = O.scalar(es, r, li), O.scalar(hand.es, hr, li)
                assert a.dtype == b.dtype and (a.view(np.uint8) == b.view(np.uint8)).all(), (q, f, r)
        for r, hr in zip(st["group_by_roots"], getattr(hand, "group_by", [])):
            assert es.nodes[r].a == hand.es.nodes[hr].a
