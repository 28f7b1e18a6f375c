#!/usr/bin/env python
"""One small invocation of every kernel family, for compute-sanitizer:

  compute-sanitizer --tool racecheck --log-file gpurun_out/racecheck.log python tools/racecheck_small.py
  compute-sanitizer --tool memcheck  --log-file gpurun_out/memcheck.log  python tools/racecheck_small.py

Sizes are a few thousand rows (the sanitizer slows the NVRTC-compiled scan kernels down by two to three orders of
magnitude); every result is still checked against a numpy closed form, so a race that changes an answer fails here too.
Covers: Q6 / Q1 / Q3 operator chains (scan_agg single + compact key, select + LIP probe, LIP build, join build + probe,
hash group-by, finalize, top-k in one CTA), the cooperative top-k, K8 partitioning, a left outer join, and the
NULL-able forms (aggregates, select with NULL masks, join on NULL-able keys, staging of the three NULL representations)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

import tpch_data as D  # noqa: E402
from quickstep_b200 import capi as A  # noqa: E402
from quickstep_b200 import engine as E  # noqa: E402
from quickstep_b200 import tpch as T  # noqa: E402
from quickstep_b200.expr import ExprSet  # noqa: E402
from quickstep_b200.table import Column, HostTable  # noqa: E402


def main():
    E.init([0])
    rng = np.random.default_rng(1)
    # ---- TPC-H operator chains over ~6000 lineitem rows
    tables = D.synthetic_tables(6000, seed=5)
    rels = {k: E.Relation.from_host(v) for k, v in tables.items()}
    li = tables["lineitem"]
    rev, is_null = T.run_q6(rels["lineitem"])
    c = {col.name: col.data for col in li.columns}
    sd = c["l_shipdate"].view(np.uint64)
    key = (sd & np.uint64(0xffffffff)).astype(np.int64) * 65536 + ((sd >> np.uint64(32)) & np.uint64(0xff)).astype(np.int64) * 256 + ((sd >> np.uint64(40)) & np.uint64(0xff)).astype(np.int64)
    m = (key >= 1994 * 65536 + 256 + 1) & (key < 1995 * 65536 + 256 + 1) & (c["l_discount"] >= 0.05) & (c["l_discount"] <= 0.07) & (c["l_quantity"] < 24)
    want = float((c["l_extendedprice"][m] * c["l_discount"][m]).sum())
    assert not is_null and abs(rev - want) <= 1e-9 * abs(want), (rev, want)
    rows = T.run_q1(rels["lineitem"])
    assert sum(r["count_order"] for r in rows) == int((key <= 1998 * 65536 + 9 * 256 + 1).sum())
    top = T.run_q3(rels["customer"], rels["orders"], rels["lineitem"], D.q3_stats(tables))
    assert len(top) <= 10 and all(top[i][1] >= top[i + 1][1] for i in range(len(top) - 1))
    for r in rels.values():
        r.destroy()
    print("tpch chains ok", flush=True)

    # ---- cooperative top-k (> 16 k rows), K8 hash partition
    n = 20000
    t = HostTable("s", [Column("a", A.QS_INT, rng.integers(0, 1000, size=n).astype(np.int32)), Column("b", A.QS_LONG, rng.permutation(n).astype(np.int64))])
    rel = E.Relation.from_host(t)
    topk = E.topk(rel, [(0, False), (1, True)], 100)
    order = np.lexsort((-t.columns[1].data, t.columns[0].data))[:100]
    assert (topk.read(1) == t.columns[1].data[order]).all()
    topk.destroy()
    out = E.Relation.create(rel.schema, n)
    off = E.hash_partition(rel, 0, 8, out)
    k = out.read(0)
    assert off[-1] == n and all(((k[off[p]:off[p + 1]].astype(np.uint32) & 7) == p).all() for p in range(8))
    out.destroy(); rel.destroy()
    print("top-k / partition ok", flush=True)

    # ---- left outer join with duplicates
    b = HostTable("b", [Column("k", A.QS_INT, rng.integers(0, 300, size=800).astype(np.int32)), Column("p", A.QS_LONG, np.arange(800, dtype=np.int64))])
    p = HostTable("p", [Column("k", A.QS_INT, rng.integers(0, 500, size=3000).astype(np.int32)), Column("v", A.QS_DOUBLE, rng.normal(size=3000))])
    br, pr = E.Relation.from_host(b), E.Relation.from_host(p)
    jt = E.JoinTable(A.QS_INT, 800)
    jt.build(br, None, -1, 0)
    es = ExprSet()
    out = E.Relation.create([(A.QS_INT, 4), (A.QS_LONG, 8), (A.QS_DOUBLE, 8)], 20000)
    jt.probe(pr, es, -1, 0, A.QS_JOIN_LEFT_OUTER, -1, [es.attr(0, A.QS_INT), es.attr(1, A.QS_LONG, 8, 2), es.attr(1, A.QS_DOUBLE)], out)
    counts = np.bincount(b.columns[0].data, minlength=500)
    assert out.n_rows == int(np.maximum(counts[p.columns[0].data], 1).sum())
    assert int((out.read_nulls() != 0).sum()) == int((counts[p.columns[0].data] == 0).sum())
    out.destroy(); jt.destroy(); br.destroy(); pr.destroy()
    print("outer join ok", flush=True)

    # ---- NULL-able attributes: grouped aggregate, select with masks, join on NULL-able keys, staging
    n = 5000
    t = HostTable("t", [Column("g", A.QS_INT, rng.integers(0, 5, size=n).astype(np.int32)), Column("x", A.QS_DOUBLE, rng.normal(10, 5, size=n)),
                        Column("y", A.QS_INT, rng.integers(-100, 100, size=n).astype(np.int32))])
    xn, yn = rng.random(n) < 0.3, rng.random(n) < 0.2
    t.columns[1].data[xn] = 0.0
    t.columns[2].data[yn] = 0
    nulls = (xn.astype(np.uint64) << np.uint64(1)) | (yn.astype(np.uint64) << np.uint64(2))
    rel = E.Relation.from_host(t)
    rel.set_nullable([1, 2])
    rel.write_nulls(nulls)
    es = ExprSet()
    x, y = es.attr(1, A.QS_DOUBLE), es.attr(2, A.QS_INT)
    for strategy in (A.QS_AGG_COMPACT_KEY, A.QS_AGG_SEPARATE_CHAINING):
        st = E.AggState(strategy, es, es.cmp(A.QS_GT, y, es.lit_int(-50)), [(A.QS_AGG_SUM, x), (A.QS_AGG_COUNT, x), (A.QS_AGG_MIN, y), (A.QS_AGG_COUNT, -1)],
                        [es.attr(0, A.QS_INT)], estimated=16, nullable_args=[0, 1, 2])
        st.run(rel)
        out, _ = E.finalize_relation(st, [(A.QS_INT, 4)], [(A.QS_DOUBLE, 8), (A.QS_LONG, 8), (A.QS_INT, 4), (A.QS_LONG, 8)])
        cols = out.read_all()
        keep = ~yn & (t.columns[2].data > -50)
        for i, g in enumerate(cols[0]):
            sel = keep & (t.columns[0].data == g)
            assert int(cols[4][i]) == int(sel.sum()) and int(cols[2][i]) == int((sel & ~xn).sum())
            w = float(t.columns[1].data[sel & ~xn].sum())
            assert abs(float(cols[1][i]) - w) <= 1e-9 * abs(w)
            assert int(cols[3][i]) == int(t.columns[2].data[sel].min())
        out.destroy(); st.destroy()
    out = E.Relation.create([(A.QS_INT, 4), (A.QS_DOUBLE, 8)], n)
    E.select(rel, es, es.not_(es.cmp(A.QS_LT, x, es.lit_double(11.0))), None, [es.attr(0, A.QS_INT), es.add(x, es.cast(y, A.QS_DOUBLE))], out)
    keep = ~(~xn & (t.columns[1].data < 11.0))
    assert out.n_rows == int(keep.sum()) and int(((out.read_nulls() >> np.uint64(1)) & np.uint64(1)).sum()) == int((keep & (xn | yn)).sum())
    out.destroy()
    jt = E.JoinTable(A.QS_INT, n)
    jt.build(rel, None, -1, 2)
    assert jt.num_entries() == int((~yn).sum())
    jt.destroy(); rel.destroy()
    from test_gpu_nulls import make_null_block_images
    images, expect = make_null_block_images(rng, 1500)
    rel = E.Relation.create([(A.QS_LONG, 8)], 1500 * len(images))
    rel.set_nullable([0])
    rel.stage_blocks(images)
    got_nulls = rel.read_nulls()
    assert (got_nulls == np.concatenate([isn for _v, isn in expect]).astype(np.uint64)).all()
    rel.destroy()
    print("NULL-able forms ok", flush=True)
    print("RACECHECK WORKLOAD OK", flush=True)


if __name__ == "__main__":
    main()
