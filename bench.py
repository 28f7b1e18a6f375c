#!/usr/bin/env python
"""bench.py -- TPC-H Q1 / Q6 / Q3 at SF100 (BASELINE.json's metric config) through the C++ operator layer
(libqshost.so: RelationalOperator / WorkOrder subclasses scheduled by Foreman + Workers) over the C ABI of
libqsgpu.so, plus the hash-join microbench (configs[4]).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...   # the CPU path (oracle port) on the host cores

ONE synthetic TPC-H-shaped database of `--sf` (default 100: 600,037,902 lineitem rows) whose content does not
depend on N (quickstep_b200/synth.py: fixed chunks any rank can regenerate).  N = 1 holds all of it; at N > 1
lineitem is block-partitioned on l_orderkey boundaries over the ranks and orders / customer are held in shares
(STRONG scaling: the job is the same at every N).  A "step" is one complete query, admit -> result rows on the
host, with the relations resident in HBM:
  Q1, Q6   per-rank scan + aggregate, partial states merged with NCCL inside the C ABI (qsgpu_agg_merge_all)
  Q3       customer LIP filter OR-reduced over the ranks, orders joined with lineitem partition-wise (they are
           co-partitioned on the order key) or, as `q3_broadcast_join`, with the filtered orders all-gathered to every
           rank; per-rank top-10 candidates gathered
Every query's answer is compared with the CPU oracle run over the WHOLE database (rank 0; all rows of all three
answers; counts and keys exact, double sums to 1e-9 relative).  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)

# dbgen's lineitem cardinalities (TPC-H specification)
SF_ROWS = {1: 6_001_215, 10: 59_986_052, 100: 600_037_902}
SEED = 20261018
BLOCK_ROWS = 63_000          # ~ rows of one 4 MB lineitem block
REL_TOL = 1e-9


def total_rows(args):
    if args.rows:
        return args.rows
    sf = args.sf
    return SF_ROWS.get(int(sf), int(sf * 6_000_000)) if float(sf).is_integer() else int(sf * 6_000_000)


def workload_name(args, n):
    sf = f"SF{args.sf:g}" if not args.rows else f"{n / 6e6:.2f} x SF1-sized"
    return f"TPC-H Q1 at {sf} (lineitem {n:,} rows, 42 B/row, 4 groups x 6 states)"


def n_host_workers(args, world):
    """Worker threads per rank: they poll for work while a query is in flight, so never more than this rank's share of
    the host cores (8 ranks x 4 workers would oversubscribe a 32-core host)."""
    return max(1, min(args.workers, (os.cpu_count() or 4) // world - 2))


def bench_config(args, n, world, rows_rank0, n_workers):
    """`config` of the JSON line.  Both arms print the SAME object (the reference arm times the CPU path on this very
    workload): it names the job, not the implementation that ran it."""
    return {"workload": workload_name(args, n), "total_rows": n, "rows_per_gpu": rows_rank0,
            "partitioning": f"lineitem block-partitioned on l_orderkey boundaries over {world} GPU(s); orders / customer in shares "
                            "(orders co-partitioned with lineitem: Q3 joins partition-wise; q3_broadcast_join all-gathers the build side instead)",
            "path": "libqshost.so (C++ RelationalOperator / WorkOrder layer, Foreman + Workers) -> libqsgpu.so C ABI; "
                    "cross-GPU merges inside the C ABI (NCCL)",
            "host_workers_per_rank": n_workers,
            "l2": f"inputs ({rows_rank0 * 42 / 1e9:.1f} GB per GPU) larger than L2 (126 MB); no flush needed",
            "timing": "CUDA events on the library stream around K whole queries (kernels + collectives + result read); max over ranks"}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            d = json.load(f)
        for k in ("hbm_gbs", "hbm_gb_s", "hbm_GBps"):
            if k in d:
                return float(d[k]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, dev):
        self.dev, self.rows, self.proc = dev, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.dev), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(int(f[0])); mx = max(mx, int(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------- parity (oracle)
def rel_err(a, b):
    return 0.0 if a == b else abs(a - b) / max(abs(a), abs(b), 1e-300)


def check_q1(got, want):
    """All 8 outputs of every group: keys and counts exact, the 7 double columns within REL_TOL."""
    assert len(got) == len(want), (len(got), len(want))
    worst = 0.0
    for g, w in zip(got, want):
        assert g["l_returnflag"] == w["l_returnflag"] and g["l_linestatus"] == w["l_linestatus"], (g, w)
        assert int(g["count_order"]) == int(w["count_order"]), (g, w)
        for f in ("sum_qty", "sum_base_price", "sum_disc_price", "sum_charge", "avg_qty", "avg_price", "avg_disc"):
            e = rel_err(g[f], w[f])
            assert e <= REL_TOL, (f, g[f], w[f], e)
            worst = max(worst, e)
    return {"groups": len(got), "count_order_total": sum(int(g["count_order"]) for g in got), "outputs_checked": 8 * len(got),
            "counts": "exact", "max_rel_err": worst, "tol": REL_TOL}


def check_q6(got, want):
    (gv, gnull), (wv, wnull) = got, want
    assert gnull == wnull, (got, want)
    e = rel_err(gv, wv)
    assert e <= REL_TOL, (gv, wv, e)
    return {"revenue": gv, "oracle": wv, "max_rel_err": e, "tol": REL_TOL}


def check_q3(got, want):
    """All 10 rows, in order: l_orderkey / o_orderdate / o_shippriority exact, revenue within REL_TOL.  Rows whose
    revenues tie within the tolerance may swap places; compare such runs as sets."""
    assert len(got) == len(want), (len(got), len(want))
    worst = 0.0
    gk = sorted((r[0], r[2], r[3]) for r in got)
    wk = sorted((r[0], r[2], r[3]) for r in want)
    assert gk == wk, (got, want)
    wrev = {r[0]: r[1] for r in want}
    for r in got:
        e = rel_err(r[1], wrev[r[0]])
        assert e <= REL_TOL, (r, wrev[r[0]], e)
        worst = max(worst, e)
    for a, b in zip(got, got[1:]):
        assert a[1] >= b[1] or rel_err(a[1], b[1]) <= REL_TOL, (a, b)     # ORDER BY revenue DESC
    return {"rows": len(got), "keys": "exact", "first_row": list(got[0]) if got else None, "max_rel_err": worst, "tol": REL_TOL}


def time_reference_engine(device):
    """The UNMODIFIED reference engine (oracle/_ref/quickstep_cli_shell) on this host's cores: dbgen -s 1 -> COPY -> \\analyze ->
    each query 5x with printing off, mean of the middle 3 (benchmarks/tpch/run-benchmark.sh + process.py), next to the
    oracle port over an SF1-sized relation -- the pair calibrates the port the SF100 cpu_baseline is measured with.
    -> dict, {"error": ...} or None when the engine was not built."""
    try:
        import shutil
        import tempfile
        import ref_engine as R
        if not R.available():
            return None
        cores = os.cpu_count() or 1
        store = tempfile.mkdtemp(prefix="qs_store_")
        try:
            w0 = time.perf_counter()
            R.load("1", store, workers=cores)
            load_s = time.perf_counter() - w0
            tm = R.time_queries(store, workers=cores)
            q6_rows, _ = R.run_query(store, "06", workers=cores)
        finally:
            shutil.rmtree(store, ignore_errors=True)
        import qs_oracle as O
        import oracle_tpch as OT
        from quickstep_b200 import synth as S
        O.load(); O.set_workers(cores); O.set_block_rows(BLOCK_ROWS)
        sh1 = S.db_shape(SF_ROWS[1])
        t1h = S.host_tables(S.generate_host(sh1, range(sh1["n_chunks"]), SEED, device))
        port = {}
        for nm, fn in (("q1", lambda: OT.q1(t1h["lineitem"])), ("q6", lambda: OT.q6(t1h["lineitem"]))):
            fn()
            w0 = time.perf_counter()
            for _ in range(5):
                fn()
            port[nm] = (time.perf_counter() - w0) * 1e3 / 5
        return {"binary": "oracle/_ref/quickstep_cli_shell (unmodified reference, Release, built by oracle/build_ref.sh)",
                "sf": 1, "cores": cores, "load_s": load_s, "query_ms": {"q1": tm["01"]["ms"], "q6": tm["06"]["ms"], "q3": tm["03"]["ms"]},
                "runs_ms": {k: v["runs_ms"] for k, v in tm.items()}, "q6_revenue_printed": q6_rows[0][0] if q6_rows else None,
                "procedure": "dbgen -s 1 -> COPY -> \\analyze -> each query 5x, printing off, mean of the middle 3 (benchmarks/tpch/process.py)",
                "oracle_port_sf1_ms": port,
                "port_speedup_over_engine": {k: tm[{"q1": "01", "q6": "06"}[k]]["ms"] / port[k] for k in port}}
    except Exception as ex:
        return {"error": repr(ex)}


# ------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's CPU algorithm for the path (oracle port, C) on all host cores, over the
    WHOLE lineitem relation of the same database shape, every step: measured, nothing extrapolated."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    import qs_oracle as O
    import oracle_tpch as OT
    from quickstep_b200 import synth as S
    cores = os.cpu_count() or 1
    O.load()
    O.set_workers(cores)
    O.set_block_rows(BLOCK_ROWS)
    n = total_rows(args)
    shape = S.db_shape(n)
    dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    host = S.generate_host(shape, range(shape["n_chunks"]), SEED, dev)       # data generation only (torch); not timed
    tables = S.host_tables(host)
    steps, warmup = max(1, args.steps), max(1, args.warmup)
    if n > 100_000_000:            # ~1.1-1.3 s per step at SF100: keep the whole run within a few minutes
        steps, warmup = min(steps, 60), min(warmup, 10)
    world = max(1, args.gpus)
    (_, _), (_, l_hi0), (_, _) = S.chunk_rows(shape, S.rank_chunks(shape, world, 0)[-1])     # rank 0's lineitem partition
    for _ in range(warmup):
        OT.q1(tables["lineitem"])
    t0 = time.perf_counter()
    for _ in range(steps):
        rows = OT.q1(tables["lineitem"])
    ms = (time.perf_counter() - t0) * 1e3 / steps
    line = {
        "impl": "reference", "metric": "tpch_q1_sf100_query_ms" if n == SF_ROWS[100] else "tpch_q1_query_ms", "value": ms, "unit": "ms",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(args, n, world, l_hi0, n_host_workers(args, world)),      # the job our arm ran
        "impl_path": "CPU: oracle/libqsoracle.so (C restatement of the reference's aggregation path, pinned to the unmodified "
                     "engine's answers), all host threads, over the WHOLE relation; no GPU, no partitioning",
        "rows_per_s": n / (ms * 1e-3),
        "cpu_baseline": {"value": ms, "unit": "ms", "cores": cores, "kind": "port",
                         "sample": f"oracle Q1 over the whole relation ({n:,} rows), {cores} threads, 63k-row work orders; "
                                   f"{steps} steps after {warmup} warm-up, measured (no extrapolation)"},
        "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "result_check": {"q1_groups": len(rows), "q1_count": sum(int(r["count_order"]) for r in rows)},
    }
    if world == 1 and not args.no_ref_engine:
        # kind "reference" evidence beside the port: the unmodified engine itself, at the largest scale its loader takes
        # inside a bench run (SF1; SF100 is 75 GB of .tbl text: "not run")
        del host, tables
        line["reference_engine"] = time_reference_engine(dev)
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sf", type=float, default=100.0, help="scale factor of the whole database (all GPUs together)")
    ap.add_argument("--rows", type=int, default=0, help="lineitem rows of the whole database (overrides --sf)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the oracle: no parity check, no cpu_baseline")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-join", action="store_true", help="skip the hash-join microbench (configs[4])")
    ap.add_argument("--no-ref-engine", action="store_true", help="skip the run of the unmodified reference engine at SF1")
    ap.add_argument("--no-coded", action="store_true", help="skip the code-resident (dictionary codes in HBM) runs")
    ap.add_argument("--join-build-rows", type=int, default=1 << 26)
    ap.add_argument("--join-probe-rows", type=int, default=1 << 30)
    ap.add_argument("--workers", type=int, default=4)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    args.warmup = max(args.warmup, 3)

    # Libraries (NCCL with NCCL_DEBUG set, the CUDA runtime) may print to fd 1; the contract is ONE JSON line
    # on stdout, so everything else is sent to stderr and the line is written to the saved descriptor.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import numpy as np  # noqa: F401
    import torch
    import torch.distributed as dist

    from quickstep_b200 import capi as A
    from quickstep_b200 import engine as E
    from quickstep_b200 import hostapi as H
    from quickstep_b200 import synth as S
    from quickstep_b200 import tpch as T

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    E.init([local])
    comm = E.Comm.from_torch_distributed(local) if world > 1 else None
    t_start = time.perf_counter()

    def log(msg):
        if rank == 0:
            print(f"[bench {time.perf_counter() - t_start:7.1f}s] {msg}", file=sys.stderr, flush=True)

    def barrier():
        E.synchronize(local)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ------------------------------------------------------------------ data: one database, this rank's partition
    n = total_rows(args)
    shape = S.db_shape(n)
    mine = S.rank_chunks(shape, world, rank)
    want_oracle = rank == 0 and not args.no_cpu
    if want_oracle:
        # rank 0 also holds the WHOLE database on the host for the CPU oracle; its own partition is the prefix
        full = S.generate_host(shape, range(shape["n_chunks"]), SEED, device)
        (_, o_hi), (_, l_hi), (_, c_hi) = S.chunk_rows(shape, mine[-1])
        cut = {"customer": c_hi, "orders": o_hi, "lineitem": l_hi}
        host = {rel: [a[:cut[rel]] for a in cols] for rel, cols in full.items()}
    else:
        full = None
        host = S.generate_host(shape, mine, SEED, device)
    torch.cuda.empty_cache()
    my_rows = len(host["lineitem"][0])
    log(f"generated: {n:,} lineitem rows in the database, {my_rows:,} on this rank")

    n_workers = n_host_workers(args, world)
    db = H.Database(local, num_workers=n_workers)
    if comm is not None:
        db.set_comm(comm.h)
    t_load = time.perf_counter()
    for which, rel in ((H.CUSTOMER, "customer"), (H.ORDERS, "orders"), (H.LINEITEM, "lineitem")):
        db.load(which, host[rel], BLOCK_ROWS, H.COMPRESSED_COLUMN_STORE)
    t_load = time.perf_counter() - t_load
    lst = db.stats(H.LINEITEM)
    log(f"blocks built: {lst['n_blocks']} lineitem blocks, {lst['host_bytes'] / 1e9:.2f} GB, {t_load:.1f} s")

    q1 = lambda: db.q1()[0]
    q6 = lambda: db.q6()[:2]
    q3 = lambda: db.q3()[0]

    # first use: stages the blocks (H2D + decode), compiles the kernels (NVRTC, cached on disk), and at N > 1 sets up
    # NCCL's connections (lazy: the first collectives of a process take milliseconds) -- none of it is timed
    r1, r6, r3 = q1(), q6(), q3()
    if comm is not None:
        for _ in range(8):
            comm.barrier()
            q6()
    barrier()
    log("relations resident, kernels compiled")

    def timed(step, steps, warmup):
        for _ in range(warmup):
            step()
        barrier()
        l0 = E.launch_count()
        w0 = time.perf_counter()
        E.timer_start(local)
        for _ in range(steps):
            out = step()
        dev_ms = E.timer_stop(local)
        barrier()
        wall_ms = (time.perf_counter() - w0) * 1e3
        # device events bracket the library stream (kernels + NCCL collectives); the result read is inside both clocks
        return max_over_ranks(max(dev_ms, 0.0)) / steps, max_over_ranks(wall_ms) / steps, (E.launch_count() - l0), out

    q3_steps = max(2, args.steps // 2)
    with ClockSampler(local) as clk:
        t1 = timed(q1, args.steps, args.warmup)
        t6 = timed(q6, args.steps, args.warmup)
        t3 = timed(q3, q3_steps, args.warmup)
    clocks = clk.summary()
    log(f"timed: q1 {t1[0]:.3f} ms, q6 {t6[0]:.3f} ms, q3 {t3[0]:.3f} ms")
    # N > 1: Q3 again with the build side BROADCAST (filtered orders all-gathered, every GPU builds the whole table)
    # instead of joined partition-wise: what a build side that is not co-partitioned with lineitem needs
    t3b = None
    if world > 1:
        db.set_join_mode(1)
        t3b = timed(q3, q3_steps, args.warmup)
        db.set_join_mode(0)
        check_q3(t3b[3], t3[3])
        log(f"q3 with a broadcast build side: {t3b[0]:.3f} ms")

    # ---- kernel-only times, CUDA events around each launch (timing mode: every launch waits for its kernel)
    peak, peak_src = measured_peak()

    def kernel_times(step, reps):
        E.set_timing(True)
        acc = {}
        try:
            for i in range(reps + 2):
                if i == 2:
                    E.set_timing(True)          # resets the per-family statistics after two warm passes
                step()
            for fam, name in ((A.QS_K_SCAN_AGG, "scan_agg"), (A.QS_K_SELECT, "select"), (A.QS_K_LIP, "lip_build"),
                              (A.QS_K_JOIN_BUILD, "join_build"), (A.QS_K_JOIN_PROBE, "join_probe"), (A.QS_K_GROUPBY, "groupby"),
                              (A.QS_K_TOPK, "topk")):
                s = E.kernel_ms_stats(fam)
                if s["count"]:
                    acc[name] = {"mean_per_query": s["sum"] / reps, "longest": s["max"], "launches_per_query": s["count"] / reps}
        finally:
            E.set_timing(False)
        return acc

    k1, k6, k3 = kernel_times(q1, min(args.steps, 10)), kernel_times(q6, min(args.steps, 10)), kernel_times(q3, min(args.steps, 5))
    o_rows, c_rows = len(host["orders"][0]), len(host["customer"][0])

    def roofline(kernel, kernel_ms, bytes_, note):
        kernel_ms = max_over_ranks(kernel_ms)
        ach = bytes_ / (kernel_ms * 1e-3) / 1e9
        return {"bound": "hbm", "kernel": kernel, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "peak_source": peak_src, "frac_of_nominal_8tbs": ach / 8000.0, "kernel_ms": kernel_ms,
                "algorithmic_bytes": bytes_, "traffic": None, "note": note}

    roof_q1 = roofline("qs_scan_agg_* (NVRTC instance of scan_agg_body for Q1: scan + compact-key group-by aggregation, "
                       "4 register-resident groups, 5 SUM + COUNT)", k1["scan_agg"]["mean_per_query"], my_rows * T.Q1_BYTES_PER_ROW,
                       "42 B/row x the rows of this rank's lineitem partition; slowest rank")
    roof_q6 = roofline("qs_scan_agg_* (Q6: scan + single-state SUM)", k6["scan_agg"]["mean_per_query"], my_rows * T.Q6_BYTES_PER_ROW,
                       "32 B/row x rows")
    roof_q3 = roofline("qs_scan_select_* (Q3 lineitem: l_shipdate predicate + LIP probe on l_orderkey + 3-column projection)",
                       k3["select"]["longest"], my_rows * T.Q3_LINEITEM_BYTES_PER_ROW, "28 B/row x rows; output rows (~3 % of the input) not counted")
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            per_row = tj.get("q1_scan_agg_dram_bytes_per_row")
            roof_q1["traffic"] = per_row * my_rows if per_row else None
            roof_q1["traffic_source"] = tj.get("source")
            per_row6 = tj.get("q6_scan_agg_dram_bytes_per_row")
            roof_q6["traffic"] = per_row6 * my_rows if per_row6 else None
            roof_q6["traffic_source"] = tj.get("source")
            per_row3 = tj.get("q3_lineitem_select_dram_bytes_per_row")
            roof_q3["traffic"] = per_row3 * my_rows if per_row3 else None
            roof_q3["traffic_source"] = tj.get("q3_source")
        except Exception:
            pass

    # ---- parity at the benchmarked size: the oracle over the WHOLE database (rank 0), every row of every answer
    parity, cpu = None, None
    if want_oracle:
        import qs_oracle as O
        import oracle_tpch as OT
        import tpch_data as D
        cores = os.cpu_count() or 1
        O.load(); O.set_workers(cores); O.set_block_rows(BLOCK_ROWS)
        tables = S.host_tables(full)
        w0 = time.perf_counter()
        o1 = OT.q1(tables["lineitem"])
        c_first = (time.perf_counter() - w0) * 1e3
        c_steps = 3 if world == 1 else 0
        w0 = time.perf_counter()
        for _ in range(c_steps):
            OT.q1(tables["lineitem"])
        c_ms = (time.perf_counter() - w0) * 1e3 / c_steps if c_steps else c_first
        w0 = time.perf_counter()
        o6 = OT.q6(tables["lineitem"])
        c6_ms = (time.perf_counter() - w0) * 1e3
        w0 = time.perf_counter()
        o3 = OT.q3(tables, D.q3_stats(tables))
        c3_ms = (time.perf_counter() - w0) * 1e3
        parity = {"q1": check_q1(t1[3], o1), "q6": check_q6(t6[3], o6), "q3": check_q3(t3[3], o3),
                  "against": f"CPU oracle over the whole database ({n:,} lineitem rows), merged GPU result of {world} rank(s)"}
        if world == 1:
            cpu = {"value": c_ms, "unit": "ms", "cores": cores, "kind": "port",
                   "sample": f"oracle Q1 over the whole relation ({n:,} rows), {cores} threads, 63k-row work orders: mean of "
                             f"{c_steps} runs after one warm run, measured (no extrapolation)",
                   "q6_ms": c6_ms, "q3_ms": c3_ms}
        log(f"parity ok; oracle q1 {c_ms:.0f} ms, q6 {c6_ms:.0f} ms, q3 {c3_ms:.0f} ms")
        del tables
    if world > 1:      # every rank holds the merged answer: they must agree with rank 0's bit for bit
        mine_sig = torch.tensor([float(sum(int(r["count_order"]) for r in t1[3])), t1[3][0]["sum_charge"] if t1[3] else 0.0, t6[3][0],
                                 t3[3][0][1] if t3[3] else 0.0],
                                dtype=torch.float64, device=device)
        ref_sig = mine_sig.clone()
        dist.broadcast(ref_sig, 0)
        assert torch.equal(mine_sig, ref_sig), (mine_sig, ref_sig)

    # ---- dictionary-coded residency: the blocks' dictionary-compressed attributes stay
    # 1/2-byte codes in HBM, re-coded to one relation-wide dictionary while staging; same DAGs, same answers
    coded = None
    if not args.no_coded:
        db.set_code_resident(True)
        c1r, c6r, c3r = q1(), q6(), q3()
        check_q1(c1r, t1[3]); check_q6(c6r, t6[3]); check_q3(c3r, t3[3])
        ct1, ct6 = timed(q1, args.steps, args.warmup), timed(q6, args.steps, args.warmup)
        ct3 = timed(q3, q3_steps, args.warmup)
        ck1, ck6 = kernel_times(q1, min(args.steps, 10)), kernel_times(q6, min(args.steps, 10))
        coding = {nm: db.resident_coding(H.LINEITEM, i) for i, (nm, _t, _w) in enumerate(T.LINEITEM)}
        width = {nm: w for (nm, _t, w) in T.LINEITEM}
        bpr = lambda names: sum(coding[x][0] or width[x] for x in names)
        q1_bpr = bpr(["l_shipdate", "l_returnflag", "l_linestatus", "l_quantity", "l_extendedprice", "l_discount", "l_tax"])
        q6_bpr = bpr(["l_shipdate", "l_discount", "l_quantity", "l_extendedprice"])
        kc1, kc6 = max_over_ranks(ck1["scan_agg"]["mean_per_query"]), max_over_ranks(ck6["scan_agg"]["mean_per_query"])
        coded = {"lineitem_coding": {k: {"code_bytes": v[0], "entries": v[1]} for k, v in coding.items() if v[0]},
                 "bytes_per_row": {"q1": q1_bpr, "q6": q6_bpr, "q1_native": T.Q1_BYTES_PER_ROW, "q6_native": T.Q6_BYTES_PER_ROW},
                 "query_ms": {"q1": ct1[0], "q6": ct6[0], "q3": ct3[0]},
                 "kernel_ms": {"q1_scan_agg": kc1, "q6_scan_agg": kc6},
                 "hbm_frac": {"q1": my_rows * q1_bpr / (kc1 * 1e-3) / 1e9 / peak, "q6": my_rows * q6_bpr / (kc6 * 1e-3) / 1e9 / peak},
                 "note": "same operator DAGs with the storage manager in code-resident mode; answers checked against the native run"}
        db.set_code_resident(False)
        q1()
        log(f"code-resident: q1 {ct1[0]:.3f} ms, q6 {ct6[0]:.3f} ms, q3 {ct3[0]:.3f} ms")

    # ---- e2e: HOST storage blocks -> stage (H2D + decode) -> query -> result rows (D2H).  lineitem lives on the host as
    # compressed-column-store blocks of 63,000 tuples (the reference's own format for lineitem,
    # benchmarks/tpch/create.sql:18-114; ~4 MB blocks), in one pinned slab.  Every step evicts the HBM image first.
    e2e = None
    if not args.no_e2e:
        def step_e2e():
            db.evict(H.LINEITEM)
            return q1()

        e_steps = max(2, min(args.steps, 5))
        e_ms, e_wall, _, e_rows = timed(step_e2e, e_steps, 1)
        check_q1(e_rows, t1[3])
        e2e = {"value": e_wall, "unit": "ms", "h2d_bytes_per_step": lst["host_bytes"],
               "d2h_bytes_per_step": 4 * (2 + 9 * 8), "device_ms": e_ms, "steps": e_steps, "blocks": lst["n_blocks"],
               "native_bytes": my_rows * T.Q1_BYTES_PER_ROW,
               "note": "qshost_q1 from host compressed-column-store blocks (63k tuples each, pinned slab): per rank "
                       "qsgpu_stage_blocks (H2D of the block images + one decode launch per chunk) + operator DAG + "
                       "cross-rank merge + result rows; HBM image evicted before every step; wall clock, max over ranks; "
                       "h2d bytes are this rank's",
               "block_build_s": t_load}
        log(f"e2e: {e_wall:.1f} ms per query from host blocks")

    launches_per_query = t1[2] / args.steps
    db.destroy()
    del host, full

    # ---- BASELINE.json configs[4]: hash-join microbench (64 Mi x 1 Gi int64 keys; radix-partitioned, dense table;
    # at N > 1 the partition kernel writes straight into the peers over NVLink)
    join = None
    if not args.no_join:
        import joinbench
        torch.cuda.empty_cache()
        jargs = joinbench.make_parser().parse_args([])
        jargs.build_rows, jargs.probe_rows = args.join_build_rows, args.join_probe_rows
        jargs.steps, jargs.warmup, jargs.radix, jargs.dense, jargs.fused = 3, 1, 32, True, world > 1
        try:
            join = joinbench.run(jargs, rank, world, local, device)
            if join is not None:
                ph = join["phases_ms"]
                nbr, npr = jargs.build_rows // world, jargs.probe_rows // world
                join["roofline_probe"] = {"bound": "hbm", "achieved": join["hbm"]["probe_GBps"], "peak": peak, "unit": "GB/s",
                                          "frac": join["hbm"]["probe_GBps"] / peak, "kernel_ms": ph["probe"]}
                join["roofline_partition"] = {"bound": "hbm", "achieved": (nbr * 16 + npr * 8) * 3 / (ph["partition"] * 1e-3) / 1e9 if ph["partition"] else None,
                                              "peak": peak, "unit": "GB/s", "note": "K8: rows read twice (histogram + scatter) and written once"}
                if join["roofline_partition"]["achieved"]:
                    join["roofline_partition"]["frac"] = join["roofline_partition"]["achieved"] / peak
        except Exception as ex:        # the microbench must not take the headline down with it
            join = {"error": repr(ex)}
        log("join microbench done")

    # ---- the UNMODIFIED reference engine on this host's cores (north_star: "the reference's own multi-threaded CPU path
    # timed on the B200 host's cores in the same run").  SF100 does not fit its loader in a bench run (75 GB of .tbl
    # text), so it is run at SF1 -- BASELINE.json configs[0]'s size -- with benchmarks/tpch/run-benchmark.sh's procedure
    # (5 runs, mean of the middle 3), next to the oracle port on an SF1-sized relation: the pair calibrates the port
    # the SF100 cpu_baseline is measured with.  Nothing is extrapolated.
    ref_engine = None
    if rank == 0 and world == 1 and not args.no_ref_engine:
        ref_engine = time_reference_engine(device)
        if ref_engine and "query_ms" in ref_engine:
            q = ref_engine["query_ms"]
            log(f"reference engine at SF1 on {ref_engine['cores']} cores: q1 {q['q1']:.1f} ms, q6 {q['q6']:.1f} ms, q3 {q['q3']:.1f} ms")

    if rank == 0:
        sf100 = n == SF_ROWS[100]
        line = {
            "metric": "tpch_q1_sf100_query_ms" if sf100 else "tpch_q1_query_ms", "value": t1[0], "unit": "ms", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t1[1], "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": bench_config(args, n, world, my_rows, n_workers),
            "rows_per_s": n / (t1[0] * 1e-3),
            "query_ms": {"q1": t1[0], "q6": t6[0], "q3": t3[0], **({"q3_broadcast_join": t3b[0]} if t3b else {})},
            "query_wall_ms": {"q1": t1[1], "q6": t6[1], "q3": t3[1]},
            "kernels_ms": {"q1": k1, "q6": k6, "q3": k3},
            "hbm_frac": {"q1": roof_q1["frac"], "q6": roof_q6["frac"], "q3_lineitem_select": roof_q3["frac"],
                         # Q3 is a chain of ten operators: algorithmic input bytes of the three relations over the WHOLE query
                         "q3_whole_query": (my_rows * T.Q3_LINEITEM_BYTES_PER_ROW + o_rows * T.Q3_ORDERS_BYTES_PER_ROW +
                                            c_rows * T.Q3_CUSTOMER_BYTES_PER_ROW) / (t3[0] * 1e-3) / 1e9 / peak,
                         "q1_whole_query": my_rows * T.Q1_BYTES_PER_ROW / (t1[0] * 1e-3) / 1e9 / peak,
                         "q6_whole_query": my_rows * T.Q6_BYTES_PER_ROW / (t6[0] * 1e-3) / 1e9 / peak},
            "roofline": roof_q1, "rooflines": {"q1": roof_q1, "q6": roof_q6, "q3": roof_q3},
            "dictionary_coded": coded, "join_microbench": join,
            "cpu_baseline": cpu, "reference_engine": ref_engine, "e2e": e2e, "clocks": clocks,
            "gpu_launches": t1[2], "gpu_launches_per_query": launches_per_query,
            "result_check": {"parity": parity, "q1_groups": len(t1[3]), "q1_count": sum(int(r["count_order"]) for r in t1[3]),
                             "q6_revenue": t6[3][0], "q3_first_row": list(t3[3][0]) if t3[3] else None,
                             "ranks_agree": True if world > 1 else None},
        }
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if comm is not None:
        comm.destroy()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
