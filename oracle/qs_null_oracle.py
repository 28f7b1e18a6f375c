"""CPU restatement of the reference's NULL rules on the scan / join / aggregation path.

TEST INFRASTRUCTURE ONLY (tests/ import it as the checker; nothing under quickstep_b200/ does).

A relation with NULL-able attributes is a HostTable (NULL values stored as zero bytes) plus `nulls`, one uint64
per row with bit a set when attribute a is NULL -- the layout qsgpu_relation_read_nulls returns.  Everything
here is numpy on top of qs_oracle.predicate / qs_oracle.scalar (which see only the stored bytes):

  * a scalar over a NULL operand is NULL: ArithmeticBinaryOperators.hpp:178-186 (applyToTypedValues returns
    a NULL of the result type when either operand is), UnaryOperation likewise;
  * a comparison with a NULL operand is false: LiteralComparators-inl.hpp:168-223
    (`!(cv_nullable && cv_value == nullptr) && compare(...)`), :264-290 for two attributes;
  * NOT complements the operand's matches, AND / OR intersect / unite them -- no third truth value:
    NegationPredicate.cpp:75-94, ConjunctionPredicate.cpp:138, DisjunctionPredicate.cpp:145;
  * SUM / AVG / MIN / MAX skip NULL arguments and are NULL when they saw no value (under GROUP BY only MIN / MAX
    are: the unmodified engine prints SUM = 0 and AVG = -nan for a group whose arguments are all NULL,
    tests/golden/ref_null_results.json):
    AggregationHandleSum.hpp:117-127 (`if (value.isNull()) return;`), AggregationHandleSum.cpp:134-143,
    AggregationHandleAvg.hpp:121-131 (count of non-NULL values), AggregationHandleMin.hpp:188-200,
    AggregationHandleMax.hpp:188-200;
    COUNT(x) counts the non-NULL values (AggregationHandleCount.hpp:126-133, nullable_type == true),
    COUNT(*) counts rows;
  * rows with a NULL key neither enter a join hash table nor match in it: storage/HashTable.hpp:1384,1903;
    rows with a NULL group-by key belong to no group: storage/PackedPayloadHashTable.hpp:861-866.
"""
import numpy as np

import qs_oracle as O
from quickstep_b200 import capi as A


def null_of(es, root, nulls: np.ndarray) -> np.ndarray:
    """Is the scalar rooted at `root` NULL, per row."""
    n = es.nodes[root]
    if n.kind == A.QS_N_ATTRIBUTE:
        if n.b == 2:
            return np.zeros(len(nulls), dtype=bool)
        return ((nulls >> np.uint64(n.a)) & np.uint64(1)).astype(bool)
    if n.kind in (A.QS_N_UNARY, A.QS_N_SHARED):
        return null_of(es, n.a, nulls)
    if n.kind == A.QS_N_BINARY:
        return null_of(es, n.a, nulls) | null_of(es, n.b, nulls)
    return np.zeros(len(nulls), dtype=bool)


def predicate(es, root, table, nulls: np.ndarray) -> np.ndarray:
    """Rows matching the predicate rooted at `root` (bool per row)."""
    n = es.nodes[root]
    rows = table.n_rows
    if n.kind == A.QS_N_TRUE:
        return np.ones(rows, dtype=bool)
    if n.kind == A.QS_N_FALSE:
        return np.zeros(rows, dtype=bool)
    if n.kind == A.QS_N_NEGATION:
        return ~predicate(es, n.a, table, nulls)
    if n.kind == A.QS_N_CONJUNCTION:
        return predicate(es, n.a, table, nulls) & predicate(es, n.b, table, nulls)
    if n.kind == A.QS_N_DISJUNCTION:
        return predicate(es, n.a, table, nulls) | predicate(es, n.b, table, nulls)
    assert n.kind == A.QS_N_COMPARISON
    _cnt, bm = O.predicate(es, root, table)
    m = O.bitmap_to_bool(bm, rows)
    return m & ~null_of(es, n.a, nulls) & ~null_of(es, n.b, nulls)


def aggregate(es, pred_root, aggregates, group_attr, table, nulls: np.ndarray):
    """-> {group key (python int / bytes; None without GROUP BY): [(value, is_null) per aggregate]}.
    aggregates: [(function, argument_root)]; group_attr: index of ONE non-NULL group-by attribute or None.
    Sums are accumulated in row order (double for FLOAT/DOUBLE arguments, int64 otherwise)."""
    rows = table.n_rows
    keep = predicate(es, pred_root, table, nulls) if pred_root >= 0 else np.ones(rows, dtype=bool)
    args = []
    for f, r in aggregates:
        if r < 0:
            args.append((None, np.zeros(rows, dtype=bool)))
        elif f == A.QS_AGG_COUNT:            # COUNT(x) needs only the argument's NULL-ness (x may be a CHAR attribute)
            args.append((None, null_of(es, r, nulls)))
        else:
            args.append((O.scalar(es, r, table), null_of(es, r, nulls)))
    if group_attr is None:
        groups = {None: np.nonzero(keep)[0]}
    else:
        # a row whose group-by key is NULL belongs to no group (PackedPayloadHashTable.hpp:861-866)
        keep = keep & ~((nulls >> np.uint64(group_attr)) & np.uint64(1)).astype(bool)
        keys = table.columns[group_attr].data
        groups = {}
        idx = np.nonzero(keep)[0]
        order = np.argsort(keys[idx], kind="stable")
        sk = keys[idx][order]
        cuts = np.nonzero(sk[1:] != sk[:-1])[0] + 1
        for part in np.split(idx[order], cuts):
            if len(part):
                k = keys[part[0]]
                groups[k.item() if hasattr(k, "item") else k] = np.sort(part)
    out = {}
    for k, idx in groups.items():
        res = []
        for (f, r), (vals, isnull) in zip(aggregates, args):
            if f == A.QS_AGG_COUNT:
                res.append((int(len(idx)) if r < 0 else int((~isnull[idx]).sum()), False))
                continue
            v = vals[idx][~isnull[idx]]
            if len(v) == 0:
                # no value: NULL -- except SUM / AVG under GROUP BY, whose hash-table payload is the bare running sum
                # (AggregationHandleSum.hpp:176-178, AggregationHandleAvg.hpp:180-189): 0 and 0 / 0.0 = NaN
                if group_attr is not None and f == A.QS_AGG_SUM:
                    res.append((0.0 if vals.dtype.kind == "f" else 0, False))
                elif group_attr is not None and f == A.QS_AGG_AVG:
                    res.append((float("nan"), False))
                else:
                    res.append((0, True))
                continue
            fp = v.dtype.kind == "f"
            if f == A.QS_AGG_SUM:
                res.append((float(np.add.reduce(v.astype(np.float64))) if fp else int(v.astype(np.int64).sum()), False))
            elif f == A.QS_AGG_AVG:
                s = float(np.add.reduce(v.astype(np.float64))) if fp else float(int(v.astype(np.int64).sum()))
                res.append((s / float(len(v)), False))
            elif f == A.QS_AGG_MIN:
                res.append((v.min().item(), False))
            else:
                res.append((v.max().item(), False))
        out[k] = res
    return out


def join_pairs(build_keys, build_null, probe_keys, probe_null, probe_keep=None):
    """(probe row, build row) pairs of an equi-join; NULL keys take no part on either side."""
    b_idx = np.nonzero(~build_null)[0]
    table = {}
    for i in b_idx:
        table.setdefault(int(build_keys[i]), []).append(int(i))
    pairs = []
    for p in range(len(probe_keys)):
        if probe_null[p] or (probe_keep is not None and not probe_keep[p]):
            continue
        for b in table.get(int(probe_keys[p]), ()):
            pairs.append((p, b))
    return pairs
