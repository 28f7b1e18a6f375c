"""The C++ operator layer (quickstep_b200/host: RelationalOperator / WorkOrder subclasses, QueryContext,
Foreman/Worker scheduling, storage blocks in the reference's physical layouts) running TPC-H Q1/Q6/Q3
end to end from HOST blocks, checked against the oracle and the reference binary's golden answers."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle_tpch as OT
import tpch_data as D
from quickstep_b200 import hostapi as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def close(a, b, tol=1e-9):
    return abs(a - b) <= tol * max(abs(a), abs(b), 1e-300)


def test_qshost_symbols_exported():
    """libqshost.so loads and exports every symbol include/qshost.h declares (no device needed)."""
    src = open(os.path.join(ROOT, "include", "qshost.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = sorted(set(re.findall(r"\b(qshost_[a-z0-9_]+)\s*\(", src)))
    lib = H.load()
    assert set(names) == set(H.SIGNATURES)
    for n in names:
        assert hasattr(lib, n)
    assert ctypes.sizeof(H.q1_row) == 80 and ctypes.sizeof(H.q3_row) == 24


def test_qshost_fails_without_device():
    n = ctypes.c_int(0)
    H.A.load().qsgpu_device_count(ctypes.byref(n))
    if n.value > 0:
        return
    h = ctypes.c_void_p()
    assert H.load().qshost_db_create(0, 2, ctypes.byref(h)) == H.A.QSGPU_ERR_NO_DEVICE


def _check_q1(rows, orows):
    assert len(rows) == len(orows)
    for r, o in zip(rows, orows):
        assert r["l_returnflag"] == o["l_returnflag"] and r["l_linestatus"] == o["l_linestatus"]
        assert r["count_order"] == o["count_order"]
        assert r["sum_qty"] == o["sum_qty"]
        for k in ("sum_base_price", "sum_disc_price", "sum_charge", "avg_qty", "avg_price", "avg_disc"):
            assert close(r[k], o[k]), (k, r[k], o[k])


def _check_q3(top, otop):
    assert len(top) == len(otop)
    for g, o in zip(top, otop):
        assert g[0] == o[0] and tuple(g[2]) == tuple(o[2]) and g[3] == o[3]
        assert close(g[1], o[1])


@pytest.mark.gpu
@pytest.mark.parametrize("layout,rows_per_block", [(H.BASIC_COLUMN_STORE, 0), (H.COMPRESSED_COLUMN_STORE, 7919),
                                                   (H.SPLIT_ROW_STORE, 10000), (H.COMPRESSED_COLUMN_STORE, 63000)])
def test_tpch_through_operator_layer(golden, layout, rows_per_block):
    """dbgen SF0.01 data cut into blocks of each physical layout; the three queries through the operator DAGs."""
    db = H.Database(0, num_workers=4)
    try:
        db.load_table(H.CUSTOMER, golden["customer"], rows_per_block, layout)
        db.load_table(H.ORDERS, golden["orders"], rows_per_block, layout)
        db.load_table(H.LINEITEM, golden["lineitem"], rows_per_block, layout)
        n = golden["lineitem"].n_rows
        st = db.stats(H.LINEITEM)
        assert st["n_rows"] == n and st["n_blocks"] == (1 if rows_per_block == 0 else -(-n // rows_per_block))
        if layout == H.COMPRESSED_COLUMN_STORE:      # flags/quantities/dates dictionary-compress well
            assert st["host_bytes"] < 0.6 * n * 46
        rev, is_null, wo = db.q6()
        orev, onull = OT.q6(golden["lineitem"])
        assert is_null == onull and close(rev, orev), (rev, orev)
        assert wo == 3                               # Aggregation (one coarse work order) + Finalize + Destroy
        rows, _ = db.q1()
        _check_q1(rows, OT.q1(golden["lineitem"]))
        top, wo3 = db.q3()
        _check_q3(top, OT.q3(golden, D.q3_stats(golden)))
        assert wo3 == 10
        # second run: blocks already resident, same answers
        rev2, _, _ = db.q6()
        assert rev2 == rev
    finally:
        db.destroy()


@pytest.mark.gpu
@pytest.mark.parametrize("rows_per_block", [4096, 63000])
def test_tpch_through_operator_layer_on_dictionary_codes(golden, rows_per_block):
    """Code-resident mode of the storage manager: dictionary-compressed attributes of the compressed-column-store
    blocks stay 1/2-byte codes in HBM (one relation-wide dictionary per attribute, block dictionaries re-coded on
    the device) and every operator of the three DAGs scans the codes.  Same answers as the oracle (and as the
    decode-at-staging mode: integer results identical, double sums 1e-9); many small work orders too."""
    db = H.Database(0, num_workers=4)
    try:
        for which, name in ((H.CUSTOMER, "customer"), (H.ORDERS, "orders"), (H.LINEITEM, "lineitem")):
            db.load_table(which, golden[name], rows_per_block, H.COMPRESSED_COLUMN_STORE)
        base_rev, _, _ = db.q6()
        base_rows, _ = db.q1()
        base_top, _ = db.q3()
        assert db.resident_coding(H.LINEITEM, 7) == (0, 0)
        db.set_code_resident(True)
        rev, is_null, wo = db.q6()
        orev, onull = OT.q6(golden["lineitem"])
        assert is_null == onull and close(rev, orev) and close(rev, base_rev) and wo == 3
        # l_shipdate 2-byte codes; l_quantity / l_discount / l_tax 1-byte codes; CHAR(1) flags stay native
        assert db.resident_coding(H.LINEITEM, 7)[0] == 2
        for a in (1, 3, 4):
            cw, n = db.resident_coding(H.LINEITEM, a)
            assert cw == 1 and 0 < n <= 50
        assert db.resident_coding(H.LINEITEM, 5) == (0, 0)
        rows, _ = db.q1()
        _check_q1(rows, OT.q1(golden["lineitem"]))
        _check_q1(rows, base_rows)
        top, _ = db.q3()
        _check_q3(top, OT.q3(golden, D.q3_stats(golden)))
        _check_q3(top, base_top)
        assert db.resident_coding(H.ORDERS, 2)[0] in (0, 2)       # o_orderdate: coded when every block compressed it
        H.set_rows_per_workorder(8192)
        try:
            rows2, _ = db.q1()
            _check_q1(rows2, base_rows)
            top2, _ = db.q3()
            _check_q3(top2, base_top)
        finally:
            H.set_rows_per_workorder(0)
        db.set_code_resident(False)
        rev3, _, _ = db.q6()
        assert rev3 == base_rev and db.resident_coding(H.LINEITEM, 7) == (0, 0)
    finally:
        db.destroy()


@pytest.mark.gpu
def test_many_work_orders_and_cold_restage(golden):
    """gpu_rows_per_workorder forces one work order per ~2 blocks; evicting the HBM image makes the next
    query stage the blocks again.  Integer results are identical, double sums within 1e-9."""
    db = H.Database(0, num_workers=8)
    try:
        for which, name in ((H.CUSTOMER, "customer"), (H.ORDERS, "orders"), (H.LINEITEM, "lineitem")):
            db.load_table(which, golden[name], 4096, H.COMPRESSED_COLUMN_STORE)
        base_rows, _ = db.q1()
        base_rev, _, _ = db.q6()
        base_top, _ = db.q3()
        H.set_rows_per_workorder(8192)
        try:
            rev, _, wo = db.q6()
            n_blocks = db.stats(H.LINEITEM)["n_blocks"]
            assert wo == -(-n_blocks // 2) + 2
            assert close(rev, base_rev)
            rows, _ = db.q1()
            _check_q1(rows, base_rows)
            top, _ = db.q3()
            _check_q3(top, base_top)
        finally:
            H.set_rows_per_workorder(0)
        db.evict(H.LINEITEM)
        rev, _, _ = db.q6()
        assert rev == base_rev
    finally:
        db.destroy()


@pytest.mark.gpu
@pytest.mark.skipif(not D.have_dbgen(), reason="oracle/_ref/dbgen not shipped")
def test_sf1_reference_answers_through_operator_layer():
    """BASELINE.json configs[0] (Q6 at SF1) + Q1/Q3 from compressed-column-store blocks of 63k rows (the
    reference's 4 MB lineitem blocks) against the values the unmodified reference binary printed."""
    import json
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_answers.json")))["tpch_sf1_reference_binary"]
    tb = D.dbgen_tables(1)
    db = H.Database(0, num_workers=8)
    try:
        for which, name in ((H.CUSTOMER, "customer"), (H.ORDERS, "orders"), (H.LINEITEM, "lineitem")):
            db.load_table(which, tb[name], 63000, H.COMPRESSED_COLUMN_STORE)
        assert db.stats(H.LINEITEM)["n_blocks"] == 96
        rev, _, _ = db.q6()
        assert close(rev, float(ref["q6_revenue_printed"])), rev
        rows, _ = db.q1()
        assert [(r["l_returnflag"] + r["l_linestatus"]).decode() for r in rows] == ref["q1_groups"]
        assert rows[0]["count_order"] == 1478493 and rows[0]["sum_qty"] == 37734107.0
        top, _ = db.q3()
        f = ref["q3_first_row"]
        assert top[0][0] == f["l_orderkey"] and abs(top[0][1] - f["revenue"]) < 1e-4
    finally:
        db.destroy()


@pytest.mark.gpu
def test_operator_classes_outside_tpch():
    """quickstep_b200/host/tests/host_gputest.cpp: HashJoinOperator as a LEFT OUTER join on a composite key and
    BuildAggregationExistenceMapOperator + collision-free aggregation, scheduled by the QueryManager over multi-block
    relations, against plain-loop expectations computed in the binary."""
    import subprocess
    exe = os.path.join(ROOT, "quickstep_b200", "lib", "qshost_gputest")
    assert os.path.exists(exe), "build it: make -C quickstep_b200/host"
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "outer_join_composite_key ok" in r.stdout and "existence_map_aggregation ok" in r.stdout, r.stdout


@pytest.mark.gpu
def test_result_relation_leaves_through_insert_destination_blocks(golden):
    """InsertDestination::bulkInsertTuples hand-off (storage/InsertDestination.cpp:202-216): the result rows of a query
    are written on the host as SplitRowStore blocks in the reference's own layout -- parsed here with the reader that
    parses the reference engine's block files (oracle/ref_blocks.py) -- and what the entry point returns is read from them."""
    import struct
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_blocks as RB
    db = H.Database(0, num_workers=2)
    try:
        for which, rel in ((H.CUSTOMER, "customer"), (H.ORDERS, "orders"), (H.LINEITEM, "lineitem")):
            db.load_table(which, golden[rel], 20000, H.COMPRESSED_COLUMN_STORE)
        rows, _ = db.q1()
        blocks = db.result_blocks()
        assert len(blocks) == 1 and blocks[0][1] == len(rows) == 4 and len(blocks[0][0]) == 2 << 20
        mem = blocks[0][0]
        d = RB.read_split_row_store(mem, [1, 1, 8, 8, 8, 8, 8, 8, 8, 8, 8], 0, 0)
        assert d["n_rows"] == 4 and d["contiguous"] and d["slot_bytes"] == 74
        for i, r in enumerate(rows):
            base = d["first_slot"] + i * d["slot_bytes"]
            assert mem[base:base + 1] == r["l_returnflag"] and mem[base + 1:base + 2] == r["l_linestatus"]
            vals = struct.unpack_from("<7dqd", mem, base + 2)
            assert vals[0] == r["sum_qty"] and vals[3] == r["sum_charge"] and vals[6] == r["avg_disc"]
            assert vals[7] == r["count_order"] and vals[8] == r["sum_disc"]
        top, _ = db.q3()
        (mem3, n3), = db.result_blocks()
        d3 = RB.read_split_row_store(mem3, [4, 8, 4, 8], 0, 0)
        assert n3 == d3["n_rows"] == len(top) == 10
        for i, t in enumerate(top):
            ok, = struct.unpack_from("<i", mem3, d3["first_slot"] + i * d3["slot_bytes"])
            rev, = struct.unpack_from("<d", mem3, d3["first_slot"] + i * d3["slot_bytes"] + 16)
            assert ok == t[0] and rev == t[1]
        rev6, is_null, _ = db.q6()
        (mem6, n6), = db.result_blocks()
        d6 = RB.read_split_row_store(mem6, [8], 0, 0)
        assert n6 == 1 and struct.unpack_from("<d", mem6, d6["first_slot"])[0] == rev6 and not is_null
    finally:
        db.destroy()
