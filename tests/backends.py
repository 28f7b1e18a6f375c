"""One operator-level API over two executors so that every parity case is
written once: OracleBackend (CPU restatement, the checker) and GpuBackend
(libqsgpu.so through the C-ABI, the thing under test)."""
from __future__ import annotations

import numpy as np

import qs_oracle as O
from quickstep_b200 import capi as A
from quickstep_b200.table import Column, HostTable, np_dtype


def agg_out_types(es, aggregates, table_schema_types=None):
    """Result column types of finalizeAggregate for (function, argument_root) pairs."""
    from quickstep_b200.expr import ExprSet  # noqa: F401

    def stype(i):
        n = es.nodes[i]
        if n.kind in (A.QS_N_LITERAL, A.QS_N_ATTRIBUTE):
            return n.type
        if n.kind == A.QS_N_UNARY:
            return n.type if n.op == A.QS_CAST else stype(n.a)
        if n.kind == A.QS_N_SHARED:
            return stype(n.a)
        a, b = stype(n.a), stype(n.b)
        if a == b:
            return a
        if A.QS_DOUBLE in (a, b) or {a, b} == {A.QS_LONG, A.QS_FLOAT}:
            return A.QS_DOUBLE
        if A.QS_FLOAT in (a, b):
            return A.QS_FLOAT
        return A.QS_LONG

    out = []
    for f, r in aggregates:
        if f == A.QS_AGG_COUNT:
            out.append((A.QS_LONG, 8))
        elif f == A.QS_AGG_AVG:
            out.append((A.QS_DOUBLE, 8))
        else:
            t = stype(r)
            if f == A.QS_AGG_SUM:
                t = A.QS_DOUBLE if t in (A.QS_FLOAT, A.QS_DOUBLE) else A.QS_LONG
            out.append((t, 4 if t in (A.QS_INT, A.QS_FLOAT) else 8))
    return out


class AggOut:
    """Groups sorted by packed key bytes; values[j] is a numpy array per aggregate."""

    def __init__(self, keys: np.ndarray, values: list, null_mask: int = 0):
        self.keys, self.values, self.null_mask = keys, values, null_mask

    @property
    def n_groups(self):
        return len(self.keys)


def _sort_groups(keys_u8: np.ndarray, values: list):
    if keys_u8.shape[1] == 0 or len(keys_u8) == 0:
        return keys_u8, values
    order = np.lexsort(keys_u8.T[::-1])
    return keys_u8[order], [v[order] for v in values]


class OracleBackend:
    name = "oracle"

    def relation(self, table: HostTable):
        return table

    def to_host(self, rel) -> HostTable:
        return rel

    def make_lip(self, kind, attr_type, min_value=0, max_value=0, cardinality=0, is_anti=False):
        return O.Lip(kind, min_value, max_value, cardinality, is_anti)

    def lip_words(self, lip):
        return lip.words[: lip.n_words].copy()

    def build_lip(self, rel, es, pred, probe, build):
        O.build_lip_filter(es, pred, rel, probe, build)

    def select(self, rel, es, pred, probe, roots, out_schema, capacity=None) -> HostTable:
        cols = O.select(es, pred, rel, probe, roots, out_schema)
        return HostTable("sel", [Column(f"c{i}", t, cols[i], w) for i, (t, w) in enumerate(out_schema)])

    def aggregate(self, rel, es, pred, aggregates, group_roots, strategy, key_schema, probe=None, estimated=1024,
                  max_key=-1, row_ranges=None) -> AggOut:
        r = O.aggregate(es, pred, aggregates, group_roots, rel, probe)
        types = agg_out_types(es, aggregates)
        vals = []
        for j, (t, w) in enumerate(types):
            v = r.values[j]
            if t in (A.QS_INT,):
                v = v.astype(np.int32)
            elif t == A.QS_FLOAT:
                v = v.astype(np.float32)
            if len(v) == 0:          # no groups: the oracle has no state to take the type from
                v = v.astype(np_dtype(t, w))
            vals.append(v)
        mask = 0
        for j, nul in enumerate(r.is_null):
            if nul:
                mask |= 1 << j
        keys, vals = _sort_groups(r.keys, vals)
        return AggOut(keys, vals, mask)

    def hash_join(self, build, bes_pred, build_key, probe, es, probe_pred, probe_key, join_type, residual, roots,
                  out_schema, capacity, build_es=None, probe_lips=None) -> HostTable:
        # the oracle evaluates build and probe predicates out of one expression set
        assert bes_pred == -1 or build_es is es
        cols = O.hash_join(es, build, bes_pred, build_key, probe, probe_pred, probe_key, probe_lips, join_type,
                           residual, roots, out_schema, capacity)
        out = HostTable("join", [Column(f"c{i}", t, cols[i], w) for i, (t, w) in enumerate(out_schema)])
        out.nulls = np.array(O.hash_join.last_nulls, dtype=np.uint64)     # bit j of row i: column j is NULL
        return out

    def topk(self, rel, keys, limit) -> HostTable:
        ids = O.topk(rel, keys, limit).astype(np.int64)
        return HostTable("top", [Column(c.name, c.type, c.data[ids], c.width) for c in rel.columns])


class GpuBackend:
    name = "gpu"

    def __init__(self, engine, block_rows=None):
        self.E = engine
        self.block_rows = block_rows
        self._live = []

    def relation(self, table: HostTable):
        r = self.E.Relation.from_host(table, block_rows=self.block_rows)
        self._live.append(r)
        return r

    def to_host(self, rel) -> HostTable:
        return rel.to_host()

    def make_lip(self, kind, attr_type, min_value=0, max_value=0, cardinality=0, is_anti=False):
        f = self.E.LipFilter(kind, attr_type, min_value, max_value, cardinality, is_anti)
        self._live.append(f)
        return f

    def lip_words(self, lip):
        return lip.words()

    def build_lip(self, rel, es, pred, probe, build):
        self.E.build_lip_filter(rel, es, pred, probe, build)

    def select(self, rel, es, pred, probe, roots, out_schema, capacity=None) -> HostTable:
        out = self.E.Relation.create(out_schema, max(1, capacity if capacity is not None else rel.n_rows))
        try:
            self.E.select(rel, es, pred, probe, roots, out)
            return out.to_host("sel")
        finally:
            out.destroy()

    def aggregate(self, rel, es, pred, aggregates, group_roots, strategy, key_schema, probe=None, estimated=1024,
                  max_key=-1, row_ranges=None) -> AggOut:
        st = self.E.AggState(strategy, es, pred, aggregates, group_roots, estimated, max_key)
        try:
            for lo, hi in (row_ranges or [(0, A.UINT64_MAX)]):
                st.run(rel, lo, hi, probe)
            types = agg_out_types(es, aggregates)
            fin, mask = self.E.finalize_relation(st, key_schema, types)
            try:
                n = fin.n_rows
                kcols = [fin.read(i, 0, n) for i in range(len(key_schema))]
                vals = [fin.read(len(key_schema) + j, 0, n) for j in range(len(aggregates))]
            finally:
                fin.destroy()
            if key_schema:
                keys = np.concatenate([np.ascontiguousarray(k).view(np.uint8).reshape(n, k.dtype.itemsize) for k in kcols], axis=1)
            else:
                keys = np.zeros((n, 0), dtype=np.uint8)
            keys, vals = _sort_groups(keys, vals)
            return AggOut(keys, vals, mask)
        finally:
            st.destroy()

    def hash_join(self, build, bes_pred, build_key, probe, es, probe_pred, probe_key, join_type, residual, roots,
                  out_schema, capacity, build_es=None, probe_lips=None) -> HostTable:
        key_type = build.schema[build_key][0]
        dense = None
        if getattr(self, "dense_join", False):       # dense (collision-free vector style) table over [min, max]
            kv = build.read(build_key)
            dense = (int(kv.min()), int(kv.max())) if len(kv) else (0, 0)
        jt = self.E.JoinTable(key_type, max(16, build.n_rows), dense_range=dense)
        out = self.E.Relation.create(out_schema, max(1, capacity))
        try:
            jt.build(build, build_es if bes_pred >= 0 else None, bes_pred, build_key)
            jt.probe(probe, es, probe_pred, probe_key, join_type, residual, roots, out, probe_lips)
            host = out.to_host("join")
            host.nulls = out.read_nulls()
            return host
        finally:
            out.destroy()
            jt.destroy()

    def topk(self, rel, keys, limit) -> HostTable:
        top = self.E.topk(rel, keys, limit)
        try:
            return top.to_host("top")
        finally:
            top.destroy()

    def close(self):
        for o in reversed(self._live):
            o.destroy()
        self._live = []


class CodedGpuBackend(GpuBackend):
    """GpuBackend whose base relations hold every eligible attribute as dictionary codes (the compressed block
    format as a device format): blocks of `block_rows` tuples with per-block dictionaries are re-coded into
    one relation-wide dictionary by qsgpu_stage_blocks, and every operator then scans the codes.  `min_cw`
    forces wider codes than the cardinality needs (covers the 2- and 4-byte paths on small inputs)."""
    name = "gpu-coded"

    def __init__(self, engine, block_rows=1000, min_cw=1):
        super().__init__(engine, block_rows)
        self.min_cw = min_cw
        self.n_coded = 0

    def relation(self, table: HostTable):
        coded = {}
        for a, c in enumerate(table.columns):
            d = c.data
            if len(d) == 0:
                continue
            if d.dtype.kind == "f" and (np.isnan(d).any() or (np.signbit(d) & (d == 0)).any()):
                continue            # NaN has no place in a sorted dictionary; -0.0 == 0.0 would share an entry
            n = len(np.unique(d))
            coded[a] = max(self.min_cw, 1 if n <= 256 else 2 if n <= 65536 else 4)
        self.n_coded += len(coded)
        r = self.E.Relation.from_host_coded(table, coded, block_rows=self.block_rows)
        self._live.append(r)
        return r


def table_rows(t: HostTable):
    """Order-insensitive comparison helper: rows as a sorted list of byte strings."""
    n = t.n_rows
    if n == 0:
        return []
    packed = np.concatenate([np.ascontiguousarray(c.data).view(np.uint8).reshape(n, -1) for c in t.columns], axis=1)
    return sorted(bytes(r) for r in packed)
