"""world_size-2 (gloo, CPU) test of the multi-GPU host logic: block partitioning of lineitem and the
exchange + merge of per-rank partial results.  Each rank computes its partition's partial result with the
CPU oracle (the GPU box runs the CUDA path in its place, bench.py --gpus N); the merged result must equal the
oracle's answer on the whole relation."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _slice_table(tb, lo, hi):
    from quickstep_b200.table import Column, HostTable
    return HostTable(tb.name, [Column(c.name, c.type, c.data[lo:hi], c.width) for c in tb.columns])


def _oracle_q1_with_sum_disc(tb):
    import oracle_tpch as OT
    rows = OT.q1(tb)
    for r in rows:
        r["sum_disc"] = r["avg_disc"] * r["count_order"]       # the oracle keeps the AVG; exact enough at 1e-9
    return rows


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle_tpch as OT
        import qs_oracle as O
        import tpch_data as D
        from quickstep_b200 import multigpu as M
        O.load(); O.set_workers(2)
        tables = D.golden_tables()
        li = tables["lineitem"]
        lo, hi = M.block_partition(li.n_rows, world, rank, block_rows=4096)
        okeys = li.col("l_orderkey").data
        lo, hi = M.align_to_key_boundary(okeys, lo), M.align_to_key_boundary(okeys, hi)
        part = _slice_table(li, lo, hi)
        dev = torch.device("cpu")
        # Q6
        rev, is_null = OT.q6(part)
        n_pass = 0 if is_null else 1
        grev, gnull = M.allreduce_sum(0.0 if is_null else rev, n_pass, dev)
        # Q1
        rows = M.gather_merge_q1(_oracle_q1_with_sum_disc(part), dev)
        # Q3: lineitem range-partitioned on l_orderkey, orders/customer replicated
        ptab = dict(tables)
        ptab["lineitem"] = part
        top = OT.q3(ptab, D.q3_stats(tables))
        gtop = M.gather_merge_topk(top, dev)
        # LIP words: each rank sets the bits of its share of the build side
        words = torch.zeros(64, dtype=torch.int64)
        keys = np.arange(rank, 4096, world)
        for k in keys:
            words[k >> 6] |= (-(1 << 63) if (k & 63) == 0 else 1 << (63 - (k & 63)))
        M.allreduce_lip_words(words)
        if rank == 0:
            torch.save(dict(range=(lo, hi), q6=(grev, gnull), q1=rows, q3=gtop, lip=words), os.path.join(out_dir, "r0.pt"))
        torch.save((lo, hi), os.path.join(out_dir, f"range{rank}.pt"))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_block_partition_covers_everything():
    from quickstep_b200 import multigpu as M
    for n in (0, 1, 62_999, 63_000, 63_001, 6_001_215, 59_986_052):
        for world in (1, 2, 3, 4, 8):
            ranges = [M.block_partition(n, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            for a, b in zip(ranges, ranges[1:]):
                assert a[1] == b[0]
            sizes = [(-(-(hi - lo) // 63_000)) for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1


def test_q1_row_packing_roundtrip():
    from quickstep_b200 import multigpu as M
    rows = [dict(l_returnflag=b"A", l_linestatus=b"F", sum_qty=1.5, sum_base_price=2.5, sum_disc_price=3.25,
                 sum_charge=4.125, sum_disc=0.5, count_order=(1 << 40) + 7),
            dict(l_returnflag=b"N", l_linestatus=b"O", sum_qty=9.0, sum_base_price=8.0, sum_disc_price=7.0,
                 sum_charge=6.0, sum_disc=5.0, count_order=3)]
    back = M.unpack_q1_rows(M.pack_q1_rows(rows))
    assert back == rows


@pytest.mark.timeout(300)
def test_world2_gloo_partial_result_merge(tmp_path):
    import oracle_tpch as OT
    import qs_oracle as O
    import tpch_data as D
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    got = torch.load(os.path.join(tmp_path, "r0.pt"), weights_only=False)
    r0, r1 = torch.load(os.path.join(tmp_path, "range0.pt")), torch.load(os.path.join(tmp_path, "range1.pt"))
    O.load(); O.set_workers(2)
    tables = D.golden_tables()
    n = tables["lineitem"].n_rows
    assert r0[0] == 0 and r0[1] == r1[0] and r1[1] == n and 0 <= r0[1] - (r0[1] // 4096) * 4096 < 8
    close = lambda a, b: abs(a - b) <= 1e-9 * max(abs(a), abs(b))
    rev, is_null = OT.q6(tables["lineitem"])
    assert got["q6"][1] == is_null and close(got["q6"][0], rev)
    orows = OT.q1(tables["lineitem"])
    assert len(got["q1"]) == len(orows)
    for g, o in zip(got["q1"], orows):
        assert g["l_returnflag"] == o["l_returnflag"] and g["l_linestatus"] == o["l_linestatus"]
        assert g["count_order"] == o["count_order"] and g["sum_qty"] == o["sum_qty"]
        for k in ("sum_base_price", "sum_disc_price", "sum_charge", "avg_qty", "avg_price", "avg_disc"):
            assert close(g[k], o[k]), k
    otop = OT.q3(tables, D.q3_stats(tables))
    assert [t[0] for t in got["q3"]] == [t[0] for t in otop]
    for g, o in zip(got["q3"], otop):
        assert tuple(g[2]) == tuple(o[2]) and g[3] == o[3] and close(g[1], o[1])
    # every bit 0..4095 set exactly once across the two ranks -> all ones after the OR
    assert (got["lip"] == -1).all()
