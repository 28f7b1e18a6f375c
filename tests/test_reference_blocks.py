"""The reference engine's OWN physical formats as input: block files written by the unmodified quickstep_cli_shell
(compressed column store for lineitem / orders, split row store for customer) are parsed (oracle/ref_blocks.py) and
handed to qsgpu_stage_blocks as they lie.  What comes out must be what dbgen generated, and the queries over it must
print what the engine printed."""
import os
import shutil
import tempfile

import numpy as np
import pytest

import ref_golden as RG
from quickstep_b200 import capi as A
from quickstep_b200 import tpch as T
from quickstep_b200.table import Column, HostTable, np_dtype


def _decode_host(oracle, image, n_rows, d, t, w):
    """One stripe of an engine block through the oracle's decoders (CPU)."""
    dt = np_dtype(t, w)
    mem = image
    if d["encoding"] == A.QS_ENC_STRIDED:
        return oracle.decode_strided(mem[d["offset"]:], n_rows, d["stride"], dt)
    if d["encoding"] == A.QS_ENC_PLAIN:
        return np.frombuffer(mem, dtype=dt, count=n_rows, offset=d["offset"]).copy()
    cdt = {1: np.uint8, 2: np.uint16, 4: np.uint32}[d["code_width"]]
    codes = np.frombuffer(mem, dtype=cdt, count=n_rows, offset=d["offset"]).copy()
    if d["encoding"] == A.QS_ENC_TRUNCATED:
        return oracle.decode_truncated(codes, dt)
    dvals = np.frombuffer(mem, dtype=dt, count=d["dict_entries"], offset=d["dict_offset"]).copy()
    return oracle.decode_dict(codes, dvals)


def test_engine_blocks_decode_to_dbgen_rows_oracle(oracle):
    """CPU: the oracle's decoders over the engine's stripes reproduce dbgen's rows (as a multiset: the loader's workers
    fill blocks concurrently), and the engine chose the encodings this repository's block builder models."""
    blocks, exp = RG.fixture_blocks()
    images = RG.block_images(blocks)
    seen = set()
    for rel, schema in RG.SCHEMAS.items():
        cols = [[] for _ in schema]
        for image, n_rows, descs in images[rel]:
            for i, ((_nm, t, w), d) in enumerate(zip(schema, descs)):
                cols[i].append(_decode_host(oracle, image, n_rows, d, t, w))
                seen.add(d["encoding"])
        got = [np.concatenate(c) for c in cols]
        want = [exp[nm] for (nm, _t, _w) in schema]
        assert len(got[0]) == len(want[0]) > 0
        assert RG.sorted_rows(got) == RG.sorted_rows(want), rel
    assert seen == {A.QS_ENC_PLAIN, A.QS_ENC_DICT, A.QS_ENC_TRUNCATED, A.QS_ENC_STRIDED}


@pytest.mark.gpu
def test_engine_blocks_stage_on_device(engine):
    """GPU: the same block images through qsgpu_stage_blocks (one H2D per image, one decode launch)."""
    blocks, exp = RG.fixture_blocks()
    images = RG.block_images(blocks)
    for rel, schema in RG.SCHEMAS.items():
        r = engine.Relation.create([(t, w) for (_n, t, w) in schema], sum(n for _i, n, _d in images[rel]) + 8)
        try:
            r.stage_blocks(images[rel])
            got = [r.read(i) for i in range(len(schema))]
            want = [exp[nm] for (nm, _t, _w) in schema]
            assert r.n_rows == len(want[0])
            assert RG.sorted_rows(got) == RG.sorted_rows(want), rel
        finally:
            r.destroy()


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_live_engine_blocks_sf001_queries_match_engine(engine):
    """Runs the unmodified engine here (oracle/_ref/quickstep_cli_shell): loads dbgen -s 0.01, takes the block files it
    wrote, stages them on the GPU and runs Q1 / Q6 / Q3 over them -- every cell of the answers equals what the engine
    itself prints for the same database."""
    import ref_blocks as RB
    import ref_engine as R
    import tpch_data as D
    if not R.available():
        pytest.skip("oracle/_ref/quickstep_cli_shell not built (oracle/build_ref.sh)")
    store = tempfile.mkdtemp(prefix="qs_store_")
    rels = {}
    try:
        R.load("0.01", store, workers=4)
        want = {q: R.run_query(store, q, workers=4)[0] for q in ("01", "06", "03")}
        images = RG.block_images([mem for _p, mem in RB.load_blocks(store)])
        for rel, schema in RG.SCHEMAS.items():
            rels[rel] = engine.Relation.create([(t, w) for (_n, t, w) in schema], sum(n for _i, n, _d in images[rel]) + 8,
                                               [n for (n, _t, _w) in schema])
            rels[rel].stage_blocks(images[rel])
        assert rels["lineitem"].n_rows == 60175 and rels["orders"].n_rows == 15000 and rels["customer"].n_rows == 1500
        RG.ENGINE["live"] = {"q1": {"rows": want["01"]}, "q6": {"rows": want["06"]}, "q3": {"rows": want["03"]}}
        rev, is_null = T.run_q6(rels["lineitem"])
        RG.check_q6(rev, is_null, "live")
        RG.check_q1(T.run_q1(rels["lineitem"]), "live")
        okeys = rels["orders"].read(0)
        stats = dict(c_custkey_min=1, c_custkey_max=1500, o_orderkey_min=int(okeys.min()), o_orderkey_max=int(okeys.max()),
                     orders_rows=15000, lineitem_rows=60175, customer_rows=1500, t2_estimate=7500, groups_estimate=4096)
        RG.check_q3(T.run_q3(rels["customer"], rels["orders"], rels["lineitem"], stats), "live")
        # and the committed golden tables are what this engine prints today -- up to the last digits of its double sums,
        # which depend on the order its worker threads finish their blocks in (243512.79810000001 in one run,
        # 243512.79809999999 in the next)
        gold = RG.ENGINE["sf0.01"]["q3"]["rows"]
        assert len(want["03"]) == len(gold)
        for wrow, grow in zip(want["03"], gold):
            assert [wrow[0], wrow[2], wrow[3]] == [grow[0], grow[2], grow[3]]
            assert abs(float(wrow[1]) - float(grow[1])) <= 1e-9 * abs(float(grow[1]))
    finally:
        for r in rels.values():
            r.destroy()
        shutil.rmtree(store, ignore_errors=True)
