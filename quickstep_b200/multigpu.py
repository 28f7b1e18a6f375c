"""Host-side logic of the multi-GPU path (one process per GPU, torch.distributed; NCCL on the GPU box,
gloo in the CPU tests): how lineitem is block-partitioned over ranks and how the per-rank partial results
are exchanged and merged.  No data-path collective is needed for the scans
themselves; only partial aggregation states / top-k candidates / LIP bit words cross ranks.

  Q6 (no GROUP BY)      all-reduce(SUM) of (sum, row count): AggregationHandleSum::mergeStates
                        (expressions/aggregation/AggregationHandleSum.cpp:109-117); count == 0 -> NULL
  Q1 (small GROUP BY)   all-gather of (packed key, SUM states, COUNT) rows, merged by key: the same group can
                        sit in a different slot on every rank (ThreadPrivateCompactKeyHashTable::mergeFrom,
                        storage/ThreadPrivateCompactKeyHashTable.cpp:306-363)
  Q3 (top-k)            lineitem is range-partitioned on l_orderkey, so groups are disjoint per rank: gather
                        each rank's top-k rows and re-select (SortMergeRunOperator's final merge)
  LIP filter            all-reduce(BOR) of the bit words when the build side is sharded
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import tpch as T

Q1_FIELDS = ("sum_qty", "sum_base_price", "sum_disc_price", "sum_charge", "sum_disc", "count_order")
Q1_MAX_GROUPS = 8          # 3 return flags x 2 line statuses, rounded up


def block_partition(n_rows: int, world: int, rank: int, block_rows: int = 63_000):
    """Rows [lo, hi) of rank `rank`: whole storage blocks, contiguous (lineitem is sorted on l_orderkey,
    benchmarks/tpch/create.sql:112, so this is a range partition on the join / group key), block counts
    differing by at most one between ranks."""
    n_blocks = -(-n_rows // block_rows)
    per, extra = divmod(n_blocks, world)
    b_lo = rank * per + min(rank, extra)
    b_hi = b_lo + per + (1 if rank < extra else 0)
    return min(n_rows, b_lo * block_rows), min(n_rows, b_hi * block_rows)


def align_to_key_boundary(sorted_keys, pos: int) -> int:
    """Moves a partition boundary forward to the next change of the (sorted) partition key, so that all rows
    of one l_orderkey land on one rank: Q3's groups are then disjoint across ranks and only top-k candidates
    need to be exchanged.  (A PartitionScheme on l_orderkey gives the reference the same guarantee,
    catalog/PartitionSchemeHeader.hpp.)"""
    n = len(sorted_keys)
    while 0 < pos < n and sorted_keys[pos] == sorted_keys[pos - 1]:
        pos += 1
    return pos


def pack_q1_rows(rows) -> torch.Tensor:
    """[Q1_MAX_GROUPS, 8] float64: valid flag, packed key (flag*256 + status), 5 SUMs, COUNT (exact < 2^53)."""
    t = torch.zeros(Q1_MAX_GROUPS, 2 + len(Q1_FIELDS), dtype=torch.float64)
    assert len(rows) <= Q1_MAX_GROUPS
    for i, r in enumerate(rows):
        key = r["l_returnflag"][0] * 256 + r["l_linestatus"][0]
        t[i] = torch.tensor([1.0, float(key)] + [float(r[f]) for f in Q1_FIELDS], dtype=torch.float64)
    return t


def unpack_q1_rows(t: torch.Tensor):
    rows = []
    for row in t.cpu().tolist():
        if row[0] == 0.0:
            continue
        k = int(row[1])
        r = dict(l_returnflag=bytes([k >> 8]), l_linestatus=bytes([k & 255]))
        r.update({f: row[2 + j] for j, f in enumerate(Q1_FIELDS)})
        r["count_order"] = int(r["count_order"])
        rows.append(r)
    return rows


def gather_merge_q1(rows, device, group=None):
    """rows: this rank's Q1 result rows (with sum_disc).  Returns the merged global rows on every rank."""
    world = dist.get_world_size(group)
    mine = pack_q1_rows(rows).to(device)
    allp = torch.zeros(world * Q1_MAX_GROUPS, mine.shape[1], dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(allp, mine, group=group)
    return T.merge_q1_partitions([unpack_q1_rows(allp)])


def allreduce_sum(value: float, n_rows: int, device, group=None):
    """Q6: (sum, contributing rows) summed over ranks; rows == 0 everywhere -> SQL NULL."""
    t = torch.tensor([value, float(n_rows)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    v, c = t.cpu().tolist()
    return v, int(c) == 0


def gather_merge_topk(top, device, limit=10, group=None):
    """top: this rank's rows (l_orderkey, revenue, (y, m, d), o_shippriority), already its local top-`limit`.
    ORDER BY revenue DESC, o_orderdate."""
    world = dist.get_world_size(group)
    mine = torch.full((limit, 5), -1.0, dtype=torch.float64)
    for i, r in enumerate(top[:limit]):
        mine[i] = torch.tensor([1.0, float(r[0]), float(r[1]), float(r[2][0] * 10000 + r[2][1] * 100 + r[2][2]), float(r[3])],
                               dtype=torch.float64)
    mine = mine.to(device)
    allt = torch.zeros(world * limit, 5, dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(allt, mine, group=group)
    rows = [r for r in allt.cpu().tolist() if r[0] == 1.0]
    rows.sort(key=lambda r: (-r[2], r[3]))
    out = []
    for r in rows[:limit]:
        d = int(r[3])
        out.append((int(r[1]), r[2], (d // 10000, (d // 100) % 100, d % 100), int(r[4])))
    return out


def allreduce_lip_words(words: torch.Tensor, group=None):
    """words: int64 view of a LIP filter's 64-bit words (device memory of qsgpu_lip_device_words on the GPU
    box).  Bitwise OR over ranks, in place."""
    dist.all_reduce(words, op=dist.ReduceOp.BOR, group=group)
    return words
