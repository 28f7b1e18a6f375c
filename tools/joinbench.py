#!/usr/bin/env python
"""BASELINE.json configs[4]: hash-join microbench, 64 Mi-row build x 1 Gi-row probe, int64 keys.

  python tools/joinbench.py                                   # 1 GPU: local build + probe
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/joinbench.py --gpus N

Build = a random permutation of 0..B-1 (unique keys) with payload 3*key+1; probe = uniform random keys over the
build domain (100 % hit).  Seeds are fixed (build 7, probe 1000+rank).  With N > 1 GPUs every rank holds 1/N of
both sides; rows are radix-partitioned by mix64(key) % N on the device (K8, qsgpu_radix_partition), exchanged
with NCCL all-to-all over NVLink, then built and probed locally (K5/K6).  The join output (the build payload of
every probe row) is materialised and summed by a single-state aggregation; the sum is checked against the
closed form 3*sum(probe keys) + #probe rows, bit-exact (int64).  Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

import torch
import torch.distributed as dist

from quickstep_b200 import capi as A
from quickstep_b200 import engine as E
from quickstep_b200.expr import ExprSet



def run(args, rank, world, local, dev):
    """The microbench proper, inside an already initialised process (torch.distributed up when world > 1, E.init
    done): bench.py calls it in-process for its `join_microbench` object.  -> the result line on rank 0, else None."""
    line = None
    LONG = (A.QS_LONG, 8)
    PAD = 64


    def buf(n):
        return torch.zeros(n + PAD, dtype=torch.int64, device=dev)


    def wrap(cols, n):
        return E.Relation.wrap([LONG] * len(cols), [c.data_ptr() for c in cols], n, dev=local, keep=cols)


    B, P = args.build_rows, args.probe_rows
    nb, npr = B // world, P // world
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    perm = torch.randperm(B, generator=g, device=dev, dtype=torch.int64)
    bkey = buf(nb)
    bkey[:nb] = perm[rank * nb:(rank + 1) * nb]
    del perm
    bpay = buf(nb)
    bpay[:nb] = bkey[:nb] * 3 + 1
    g.manual_seed(1000 + rank)
    pkey = buf(npr)
    pkey[:npr] = torch.randint(0, B, (npr,), generator=g, device=dev, dtype=torch.int64)
    expected_local = int((pkey[:npr] * 3 + 1).sum().item())
    build_rel, probe_rel = wrap([bkey, bpay], nb), wrap([pkey], npr)
    torch.cuda.synchronize()

    # probe-side expressions: project the build payload (attribute 1 of the build side)
    es = ExprSet()
    proj = [es.attr(1, A.QS_LONG, 8, 2)]
    es_sum = ExprSet()
    sum_arg = es_sum.attr(0, A.QS_LONG, 8)
    slack = 1.0 if world == 1 else 1.02            # hash partitions are balanced to well under 2 %
    cap_b, cap_p = int(nb * slack) + 4096, int(npr * slack) + 4096
    ipc = None
    if world > 1 and args.fused:
        # receive relations in IPC-exportable memory; every rank maps every other rank's columns
        mine = [E.ipc_alloc((cap_b + PAD) * 8, local), E.ipc_alloc((cap_b + PAD) * 8, local), E.ipc_alloc((cap_p + PAD) * 8, local)]
        handles = [None] * world
        dist.all_gather_object(handles, [h for (_p, h) in mine])
        peer = [[(mine[c][0] if r == rank else E.ipc_open(handles[r][c], local)) for c in range(3)] for r in range(world)]
        ipc = dict(mine=[p for (p, _h) in mine], peer=peer)
        recv_b_rel = E.Relation.wrap([LONG, LONG], ipc["mine"][:2], cap_b, dev=local)
        recv_p_rel = E.Relation.wrap([LONG], ipc["mine"][2:], cap_p, dev=local)
    elif world > 1:
        part_b, part_p = [buf(nb), buf(nb)], [buf(npr)]
        part_b_rel, part_p_rel = wrap(part_b, nb), wrap(part_p, npr)
        recv_b, recv_p = [buf(cap_b), buf(cap_b)], [buf(cap_p)]
    out_cap = cap_p if world > 1 else npr
    out_rel = E.Relation.create([LONG], out_cap, dev=local)
    if args.radix:
        rcap_b, rcap_p = (cap_b, cap_p) if world > 1 else (nb, npr)
        rad_b, rad_p = [buf(rcap_b), buf(rcap_b)], [buf(rcap_p)]
        rad_b_rel, rad_p_rel = wrap(rad_b, rcap_b), wrap(rad_p, rcap_p)


    def ev():
        e = torch.cuda.Event(enable_timing=True)
        return e


    def step():
        t = {}
        E.synchronize(local); torch.cuda.synchronize()
        w0 = time.perf_counter()
        if world > 1 and args.fused:
            # 1. counts per destination, 2. exchange them, 3. scatter straight into the peers, 4. barrier
            cb, cp = E.partition_count(build_rel, 0, world), E.partition_count(probe_rel, 0, world)
            t["partition_ms"] = (time.perf_counter() - w0) * 1e3
            w1 = time.perf_counter()
            send = torch.tensor([[int(cb[i]), int(cp[i])] for i in range(world)], dtype=torch.int64, device=dev)
            allc = torch.zeros(world, world, 2, dtype=torch.int64, device=dev)           # [sender][destination][side]
            dist.all_gather_into_tensor(allc.view(-1), send.view(-1))
            allc = allc.cpu()
            first_b = [int(allc[:rank, p, 0].sum()) for p in range(world)]
            first_p = [int(allc[:rank, p, 1].sum()) for p in range(world)]
            n_b, n_p = int(allc[:, rank, 0].sum()), int(allc[:, rank, 1].sum())
            assert n_b <= cap_b and n_p <= cap_p, (n_b, cap_b, n_p, cap_p)
            E.set_timing(True)
            E.partition_scatter_peers(build_rel, 0, world, [ipc["peer"][p][:2] for p in range(world)], first_b)
            xb = E.last_kernel_ms(A.QS_K_PARTITION)
            E.partition_scatter_peers(probe_rel, 0, world, [ipc["peer"][p][2:] for p in range(world)], first_p)
            xp = E.last_kernel_ms(A.QS_K_PARTITION)
            E.set_timing(False)
            dist.barrier()
            t["exchange_ms"] = xb + xp
            t["exchange_wall_ms"] = (time.perf_counter() - w1) * 1e3
            t["sent_bytes"] = 16 * (nb - int(cb[rank])) + 8 * (npr - int(cp[rank]))
            A.check(A.load().qsgpu_relation_set_num_rows(recv_b_rel.h, n_b))
            A.check(A.load().qsgpu_relation_set_num_rows(recv_p_rel.h, n_p))
            lb, lp = recv_b_rel, recv_p_rel
        elif world > 1:
            offs_b = E.radix_partition(build_rel, 0, world, part_b_rel)
            offs_p = E.radix_partition(probe_rel, 0, world, part_p_rel)
            E.synchronize(local)
            t["partition_ms"] = (time.perf_counter() - w0) * 1e3
            w1 = time.perf_counter()
            send = torch.tensor([[int(offs_b[i + 1] - offs_b[i]), int(offs_p[i + 1] - offs_p[i])] for i in range(world)],
                                dtype=torch.int64, device=dev)
            recv = torch.zeros_like(send)
            dist.all_to_all_single(recv, send)
            send_l, recv_l = send.cpu().tolist(), recv.cpu().tolist()
            sb, sp = [x[0] for x in send_l], [x[1] for x in send_l]
            rb, rp = [x[0] for x in recv_l], [x[1] for x in recv_l]
            n_b, n_p = sum(rb), sum(rp)
            assert n_b <= cap_b and n_p <= cap_p, (n_b, cap_b, n_p, cap_p)
            e0, e1 = ev(), ev()
            e0.record()
            for src, dst in zip(part_b, recv_b):
                dist.all_to_all_single(dst[:n_b], src[:nb], rb, sb)
            dist.all_to_all_single(recv_p[0][:n_p], part_p[0][:npr], rp, sp)
            e1.record()
            torch.cuda.synchronize()
            t["exchange_ms"] = e0.elapsed_time(e1)
            t["exchange_wall_ms"] = (time.perf_counter() - w1) * 1e3
            t["sent_bytes"] = 16 * (nb - sb[rank]) + 8 * (npr - sp[rank])
            lb, lp = wrap(recv_b, n_b), wrap(recv_p, n_p)
        else:
            lb, lp, n_b, n_p = build_rel, probe_rel, nb, npr
        ranges = [(0, A.UINT64_MAX)]
        jt = E.JoinTable(A.QS_LONG, max(n_b, 1024), dev=local, dense_range=(0, B - 1) if args.dense else None)
        if args.radix:
            # radix join: both sides regrouped by the slice of the join table their keys land in; partition p of the
            # probe side only touches slice p of the table and the rows of partition p of the build relation
            w2 = time.perf_counter()
            jt.partition(lb, 0, args.radix, rad_b_rel)
            offs = jt.partition(lp, 0, args.radix, rad_p_rel)
            E.synchronize(local)
            t["radix_partition_ms"] = (time.perf_counter() - w2) * 1e3
            lb, lp = rad_b_rel, rad_p_rel
            ranges = [(int(offs[i]), int(offs[i + 1])) for i in range(args.radix) if offs[i + 1] > offs[i]]
        E.set_timing(True)
        jt.build(lb, None, -1, 0)
        t["build_ms"] = E.last_kernel_ms(A.QS_K_JOIN_BUILD)
        A.check(A.load().qsgpu_relation_set_num_rows(out_rel.h, 0))
        t["probe_ms"] = 0.0
        for lo, hi in ranges:
            jt.probe(lp, es, -1, 0, A.QS_JOIN_INNER, -1, proj, out_rel, row_begin=lo, row_end=hi)
            t["probe_ms"] += E.last_kernel_ms(A.QS_K_JOIN_PROBE)
        E.set_timing(False)
        st = E.AggState(A.QS_AGG_SINGLE_STATE, es_sum, -1, [(A.QS_AGG_SUM, sum_arg), (A.QS_AGG_COUNT, -1)], [], dev=local)
        st.run(out_rel)
        fin, _ = E.finalize_relation(st, [], [LONG, LONG])
        total, count = int(fin.read(0)[0]), int(fin.read(1)[0])
        fin.destroy(); st.destroy(); jt.destroy()
        if world > 1 and not args.radix and not args.fused:
            lb.destroy(); lp.destroy()
        t["total_ms"] = (time.perf_counter() - w0) * 1e3
        t["n_build"], t["n_probe"] = n_b, n_p
        return t, total, count


    res = None
    for i in range(args.warmup + args.steps):
        t, total, count = step()
        if i >= args.warmup:
            res = t if res is None else {k: (res[k] + v) for k, v in t.items()}
    res = {k: v / args.steps for k, v in res.items()}
    chk = torch.tensor([total, count, expected_local, npr], dtype=torch.int64, device=dev)
    tm = torch.tensor([res.get("partition_ms", 0.0) + res.get("radix_partition_ms", 0.0), res.get("exchange_ms", 0.0), res["build_ms"], res["probe_ms"], res["total_ms"]],
                      dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(chk, op=dist.ReduceOp.SUM)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    total, count, expected, n_probe_all = chk.cpu().tolist()
    assert count == n_probe_all == P // world * world, (count, n_probe_all)
    assert total == expected, (total, expected)
    part, exch, build, probe, tot = tm.cpu().tolist()
    if rank == 0:
        nbr, npp = res["n_build"], res["n_probe"]
        line = {"metric": "hash_join_microbench_ms", "value": tot, "unit": "ms", "n_gpus": world, "steps": args.steps,
                "config": {"workload": f"{B} build rows x {P} probe rows, int64 keys, payload int64, 100% hit",
                           "join_table": "dense heads[key-min] + next[row] chains" if args.dense else "open addressing, 16 B slots, load factor <= 0.5",
                           "radix_partitions": args.radix,
                           "exchange": "none" if world == 1 else ("partition kernel writes into the peers' receive relations over NVLink (CUDA IPC)" if args.fused else "K8 partition, then NCCL all_to_all_single per column"),
                           "build_rows_per_gpu": nb, "probe_rows_per_gpu": npr},
                "rows_per_s": (B + P) / (tot * 1e-3),
                "phases_ms": {"partition": part, "exchange": exch, "build": build, "probe": probe},
                "hbm": {"build_GBps": nbr * 32 / (build * 1e-3) / 1e9, "probe_GBps": npp * (8 + 16 + 8 + 8) / (probe * 1e-3) / 1e9,
                        "note": "algorithmic bytes: build 16 B row read + 16 B slot written; probe 8 B key + 16 B slot + 8 B gathered payload + 8 B output"},
                "nvlink": None if world == 1 else {"sent_bytes_per_gpu": res["sent_bytes"], "GBps_per_gpu": res["sent_bytes"] / (exch * 1e-3) / 1e9,
                                                   "peak_GBps_per_direction": 900.0, "frac": res["sent_bytes"] / (exch * 1e-3) / 1e9 / 900.0},
                "check": {"sum_payload": total, "expected": expected, "matches": count}}
    if world > 1:
        dist.barrier()
        if ipc:
            for r in range(world):
                if r != rank:
                    for p in ipc["peer"][r]:
                        E.ipc_close(p, local)
            dist.barrier()
            recv_b_rel.destroy(); recv_p_rel.destroy()
            for p in ipc["mine"]:
                E.ipc_free(p, local)
    for r_ in (build_rel, probe_rel, out_rel):
        r_.destroy()
    return line if rank == 0 else None


def make_parser():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--build-rows", type=int, default=1 << 26)
    ap.add_argument("--probe-rows", type=int, default=1 << 30)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--radix", type=int, default=0, help="radix join: both sides are grouped by the slice of the join table "
                    "their keys land in (qsgpu_join_partition: home-slot prefix for open addressing, key range for --dense; "
                    "a power of two), so that every probe partition's slice of the table and of the build relation fits in L2")
    ap.add_argument("--fused", action="store_true", help="N > 1: the partition kernel writes straight into the peers' receive "
                    "relations over NVLink (CUDA IPC) instead of partition -> NCCL all-to-all")
    ap.add_argument("--dense", action="store_true", help="dense (collision-free vector style) join table over the key range")
    return ap


if __name__ == "__main__":
    args = make_parser().parse_args()

    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    E.init([local])
    line = run(args, rank, world, local, dev)
    if rank == 0:
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()
