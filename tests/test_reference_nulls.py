"""NULLs end to end against the UNMODIFIED reference engine (fixture: tests/golden/ref_null_blocks.npz and
ref_null_results.json, written by tests/golden/make_null_golden.py with oracle/_ref/quickstep_cli_shell).

The engine loaded one relation with NULL-able attributes into its three fixed-width block layouts and printed the
answers of twelve queries.  Here:
  * (CPU) the NULL oracle, fed the source rows, gives the engine's answers -- this pins oracle/qs_null_oracle.py;
  * (CPU) the block readers of oracle/ref_blocks.py find every value and every NULL where the engine put it;
  * (GPU) the engine's own block files are staged with their NULL representations (dictionary null code, per-column
    bitmap, per-tuple bitmap word), the decoded columns and masks equal the source rows, and the same queries run
    through the C ABI print the same cells: integers and NULLs exactly, doubles to 1e-9."""
import json
import os

import numpy as np
import pytest

import qs_null_oracle as NO
import ref_blocks as RB
from backends import agg_out_types
from quickstep_b200 import capi as A
from quickstep_b200.expr import ExprSet
from quickstep_b200.table import Column, HostTable

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_null_results.json")))
DATA = np.load(os.path.join(ROOT, "tests", "golden", "ref_null_blocks.npz"))
SCHEMA = [(A.QS_INT, 4), (A.QS_DOUBLE, 8), (A.QS_INT, 4), (A.QS_CHAR, 4)]
WIDTHS = [4, 8, 4, 4]
NULLABLE = [False, True, True, True]
TOL = 1e-9


def source():
    t = HostTable("t", [Column("g", A.QS_INT, DATA["g"]), Column("x", A.QS_DOUBLE, DATA["x"]), Column("y", A.QS_INT, DATA["y"]),
                        Column("c", A.QS_CHAR, DATA["c"], 4)])
    nulls = (DATA["x_null"].astype(np.uint64) << np.uint64(1)) | (DATA["y_null"].astype(np.uint64) << np.uint64(2)) | \
            (DATA["c_null"].astype(np.uint64) << np.uint64(3))
    return t, nulls


def queries():
    """name -> (ExprSet, predicate root, [(function, argument root)], group-by attribute or None, nullable aggregates)."""
    out = {}

    def new():
        es = ExprSet()
        return es, es.attr(0, A.QS_INT), es.attr(1, A.QS_DOUBLE), es.attr(2, A.QS_INT), es.attr(3, A.QS_CHAR, 4)

    es, g, x, y, c = new()
    pred = es.or_(es.cmp(A.QS_GT, y, es.lit_int(-500)), es.not_(es.cmp(A.QS_LT, x, es.lit_int(12))))
    out["single"] = (es, pred, [(A.QS_AGG_COUNT, -1), (A.QS_AGG_COUNT, x), (A.QS_AGG_SUM, x), (A.QS_AGG_AVG, x), (A.QS_AGG_MIN, y),
                                (A.QS_AGG_MAX, y), (A.QS_AGG_SUM, es.mul(x, y))], None, [1, 2, 3, 4, 5, 6])
    es, g, x, y, c = new()
    out["grouped"] = (es, -1, [(A.QS_AGG_COUNT, x), (A.QS_AGG_SUM, x), (A.QS_AGG_AVG, y), (A.QS_AGG_MIN, x), (A.QS_AGG_COUNT, c)], 0,
                      [0, 1, 2, 3, 4])
    es, g, x, y, c = new()
    out["single_all_null"] = (es, es.cmp(A.QS_EQ, g, es.lit_int(3)),
                              [(A.QS_AGG_AVG, x), (A.QS_AGG_MAX, x), (A.QS_AGG_SUM, es.mul(x, y)), (A.QS_AGG_COUNT, -1), (A.QS_AGG_SUM, x),
                               (A.QS_AGG_COUNT, x)], None, [0, 1, 2, 4, 5])
    es, g, x, y, c = new()
    out["grouped_all_null"] = (es, es.or_(es.cmp(A.QS_EQ, g, es.lit_int(3)), es.cmp(A.QS_EQ, g, es.lit_int(4))),
                               [(A.QS_AGG_AVG, x), (A.QS_AGG_MAX, x), (A.QS_AGG_SUM, es.mul(x, y)), (A.QS_AGG_COUNT, -1), (A.QS_AGG_SUM, y)], 0,
                               [0, 1, 2, 4])
    for name, build in (("not_lt", lambda es, g, x, y, c: es.not_(es.cmp(A.QS_LT, x, es.lit_int(12)))),
                        ("ge", lambda es, g, x, y, c: es.cmp(A.QS_GE, x, es.lit_int(12))),
                        ("lt", lambda es, g, x, y, c: es.cmp(A.QS_LT, x, es.lit_int(12))),
                        ("char_eq", lambda es, g, x, y, c: es.cmp(A.QS_EQ, c, es.lit_char(b"ab"))),
                        ("char_not_eq", lambda es, g, x, y, c: es.not_(es.cmp(A.QS_EQ, c, es.lit_char(b"ab"))))):
        es, g, x, y, c = new()
        out[name] = (es, build(es, g, x, y, c), [(A.QS_AGG_COUNT, -1)], None, [])
    es, g, x, y, c = new()
    out["attr_vs_attr"] = (es, es.cmp(A.QS_LT, x, y), [(A.QS_AGG_COUNT, -1), (A.QS_AGG_SUM, y)], None, [1])
    # GROUP BY a NULL-able attribute: 19 rows pass the predicate, 15 of them have a y, and the engine prints 15 groups
    es, g, x, y, c = new()
    p = es.and_(es.cmp(A.QS_EQ, g, es.lit_int(0)), es.cmp(A.QS_LT, x, es.lit_int(2)))
    out["group_by_nullable"] = (es, p, [(A.QS_AGG_COUNT, -1), (A.QS_AGG_SUM, x)], 2, [1])
    es, g, x, y, c = new()
    p = es.and_(es.cmp(A.QS_EQ, g, es.lit_int(0)), es.cmp(A.QS_LT, x, es.lit_int(2)))
    out["group_by_nullable_counts"] = (es, p, [(A.QS_AGG_COUNT, -1), (A.QS_AGG_COUNT, y)], None, [1])
    return out


def cell_matches(printed: str, value, is_null: bool) -> bool:
    if printed == "NULL":
        return is_null
    if is_null:
        return False
    if printed in ("nan", "-nan"):
        return value != value
    if isinstance(value, float) or "." in printed or "e" in printed:
        want = float(printed)
        return abs(float(value) - want) <= TOL * max(abs(want), abs(float(value))) + 1e-300
    return int(value) == int(printed)


def check_against_engine(name, table, rows):
    """rows: [[(value, is_null), ...] per output row], group rows in key order with the key as first cell."""
    want = GOLD["tables"][table][name]
    assert len(rows) == len(want), (name, table, len(rows), len(want))
    for got_row, want_row in zip(rows, want):
        assert len(got_row) == len(want_row)
        for (v, isn), printed in zip(got_row, want_row):
            assert cell_matches(printed, v, isn), (name, table, printed, v, isn)


@pytest.fixture(scope="module", autouse=True)
def _oracle(oracle):
    return oracle


@pytest.mark.parametrize("name", list(GOLD["queries"].keys()))
def test_null_oracle_gives_the_engines_answers(name):
    t, nulls = source()
    es, pred, aggs, group, _nullable = queries()[name]
    res = NO.aggregate(es, pred, aggs, group, t, nulls)
    rows = [res[None]] if group is None else [[(k, False)] + res[k] for k in sorted(res)]
    for table in GOLD["tables"]:
        check_against_engine(name, table, rows)


def block_descs(table):
    """-> (block image, n_rows, stage descriptors with the layout's NULL representation)."""
    mem = DATA["block_" + table]
    raw = mem.tobytes()
    if table == "t_row":
        info = RB.read_split_row_store(raw, WIDTHS, 0, 0, n_nullable=3)
        assert info["contiguous"] and info["null_bytes"] == 1
        descs, k = [], 0
        for a in range(4):
            d = dict(attr=a, encoding=A.QS_ENC_STRIDED, offset=info["first_slot"] + info["attr_offsets"][a], stride=info["slot_bytes"])
            if NULLABLE[a]:       # bit k (from the MSB) of the one-byte BitVector<true> at the head of the slot
                d.update(null_kind=A.QS_NULL_SLOT_WORD, null_arg=k, null_stride=info["slot_bytes"], null_width=1, null_offset=info["first_slot"])
                k += 1
            descs.append(d)
        return mem, info["n_rows"], descs
    if table == "t_col":
        info = RB.read_basic_column_store(raw, WIDTHS, NULLABLE)
        descs = []
        for a, s in enumerate(info["stripes"]):
            d = dict(attr=a, encoding=A.QS_ENC_PLAIN, offset=s["offset"])
            if s["null_offset"] is not None:
                d.update(null_kind=A.QS_NULL_BITMAP, null_arg=0, null_stride=1, null_offset=s["null_offset"])
            descs.append(d)
        return mem, info["n_rows"], descs
    info = RB.read_compressed_column_store(raw, WIDTHS)
    enc = {"dict": A.QS_ENC_DICT, "truncated": A.QS_ENC_TRUNCATED, "plain": A.QS_ENC_PLAIN}
    descs = []
    for a, s in enumerate(info["stripes"]):
        d = dict(attr=a, encoding=enc[s["encoding"]], offset=s["offset"], code_width=s["code_width"])
        if s["encoding"] == "dict":
            d.update(dict_offset=s["dict_offset"], dict_entries=s["dict_entries"])
            if NULLABLE[a]:
                d.update(null_kind=A.QS_NULL_CODE, null_arg=s["null_code"])
        elif s["null_offset"] is not None:
            d.update(null_kind=A.QS_NULL_BITMAP, null_arg=0, null_stride=1, null_offset=s["null_offset"])
        descs.append(d)
    return mem, info["n_rows"], descs


def row_multiset(cols, nulls):
    rows = []
    for i in range(len(nulls)):
        rows.append(tuple(None if (int(nulls[i]) >> a) & 1 else (cols[a][i].item() if hasattr(cols[a][i], "item") else bytes(cols[a][i]))
                          for a in range(len(cols))))
    return sorted(rows, key=repr)


def host_decode(mem, n_rows, descs):
    """numpy restatement of the three layouts' accessors, NULL representation included (the CPU check of the readers)."""
    cols, nulls = [], np.zeros(n_rows, dtype=np.uint64)
    dts = [np.dtype("<i4"), np.dtype("<f8"), np.dtype("<i4"), np.dtype("S4")]
    for a, d in enumerate(descs):
        dt = dts[a]
        if d["encoding"] == A.QS_ENC_PLAIN:
            v = np.frombuffer(mem, dtype=dt, count=n_rows, offset=d["offset"]).copy()
        elif d["encoding"] == A.QS_ENC_STRIDED:
            v = np.array([np.frombuffer(mem, dtype=dt, count=1, offset=d["offset"] + i * d["stride"])[0] for i in range(n_rows)], dtype=dt)
        elif d["encoding"] == A.QS_ENC_TRUNCATED:
            v = np.frombuffer(mem, dtype={1: "<u1", 2: "<u2", 4: "<u4"}[d["code_width"]], count=n_rows, offset=d["offset"]).astype(dt)
        else:
            codes = np.frombuffer(mem, dtype={1: "<u1", 2: "<u2", 4: "<u4"}[d["code_width"]], count=n_rows, offset=d["offset"]).astype(np.int64)
            values = np.frombuffer(mem, dtype=dt, count=d["dict_entries"], offset=d["dict_offset"])
            v = values[np.minimum(codes, d["dict_entries"] - 1)].copy()
        kind = d.get("null_kind", 0)
        isn = np.zeros(n_rows, dtype=bool)
        if kind == A.QS_NULL_CODE:
            isn = codes == d["null_arg"]
        elif kind == A.QS_NULL_BITMAP:
            words = np.frombuffer(mem, dtype="<u8", count=(n_rows + 63) // 64, offset=d["null_offset"])
            isn = np.array([(int(words[i >> 6]) >> (63 - (i & 63))) & 1 for i in range(n_rows)], dtype=bool)
        elif kind == A.QS_NULL_SLOT_WORD:
            isn = np.array([(int(mem[d["null_offset"] + i * d["null_stride"]]) >> (7 - d["null_arg"])) & 1 for i in range(n_rows)], dtype=bool)
        nulls |= isn.astype(np.uint64) << np.uint64(a)
        cols.append(v)
    return cols, nulls


@pytest.mark.parametrize("table", ["t_row", "t_col", "t_cmp"])
def test_block_readers_find_the_nulls(table):
    t, nulls = source()
    mem, n_rows, descs = block_descs(table)
    assert n_rows == t.n_rows == GOLD["rows"]
    cols, got_nulls = host_decode(mem, n_rows, descs)
    cols[3] = np.array([bytes(v).split(b"\0")[0] for v in cols[3]], dtype="S4")      # CHAR(4): up to the first NUL
    assert row_multiset(cols, got_nulls) == row_multiset([c.data for c in t.columns], nulls)
    if table == "t_cmp":     # COMPRESS ALL did compress: this fixture exercises the dictionary null code
        assert any(d.get("null_kind") == A.QS_NULL_CODE for d in descs)


@pytest.mark.gpu
@pytest.mark.parametrize("table", ["t_row", "t_col", "t_cmp"])
def test_engine_blocks_with_nulls_on_the_device(engine, table):
    t, nulls = source()
    mem, n_rows, descs = block_descs(table)
    rel = engine.Relation.create(SCHEMA, n_rows, ["g", "x", "y", "c"])
    try:
        rel.set_nullable([1, 2, 3])
        rel.stage_blocks([(mem, n_rows, descs)])
        got_nulls = rel.read_nulls()
        got = rel.read_all()
        assert row_multiset(got, got_nulls) == row_multiset([c.data for c in t.columns], nulls)
        for a in (1, 2, 3):                     # a NULL value is stored as zero bytes
            isn = ((got_nulls >> np.uint64(a)) & np.uint64(1)).astype(bool)
            assert not np.ascontiguousarray(got[a][isn]).view(np.uint8).any()
        for name, (es, pred, aggs, group, nullable) in queries().items():
            strategies = [A.QS_AGG_SINGLE_STATE] if group is None else [A.QS_AGG_COMPACT_KEY, A.QS_AGG_SEPARATE_CHAINING, A.QS_AGG_COLLISION_FREE]
            for strategy in strategies:
                if strategy == A.QS_AGG_COLLISION_FREE and any(f in (A.QS_AGG_MIN, A.QS_AGG_MAX) for f, _r in aggs):
                    continue              # that table takes COUNT / SUM / AVG only, in the reference too
                if strategy == A.QS_AGG_COLLISION_FREE and group != 0:
                    continue              # its key doubles as the array index: the small non-negative g only
                st = engine.AggState(strategy, es, pred, aggs, [es.attr(group, A.QS_INT)] if group is not None else [], estimated=16,
                                     max_key=6, nullable_args=nullable)
                try:
                    st.run(rel)
                    out, _mask = engine.finalize_relation(st, [(A.QS_INT, 4)] if group is not None else [], agg_out_types(es, aggs))
                    try:
                        cols, out_nulls = out.read_all(), out.read_nulls()
                    finally:
                        out.destroy()
                finally:
                    st.destroy()
                nk = 0 if group is None else 1
                order = np.argsort(cols[0]) if nk else [0]
                rows = []
                for i in order:
                    row = [(int(cols[0][i]), False)] if nk else []
                    for j in range(len(aggs)):
                        v = cols[nk + j][i]
                        row.append((v.item(), bool((int(out_nulls[i]) >> (nk + j)) & 1)))
                    rows.append(row)
                check_against_engine(name, table, rows)
    finally:
        rel.destroy()
