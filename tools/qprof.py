#!/usr/bin/env python
"""Runs TPC-H Q1 / Q6 / Q3 through the C++ operator layer over a synthetic database, for profiling under ncu:

  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv python tools/qprof.py --sf 10
  ncu --set full --clock-control none --import-source on -k regex:qs_join_build -o rep python tools/qprof.py --sf 10 --query q3

One warm pass (staging, NVRTC, allocator) then --reps measured passes; prints wall time per query."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

import torch

import bench as B
from quickstep_b200 import engine as E
from quickstep_b200 import hostapi as H
from quickstep_b200 import synth as S

ap = argparse.ArgumentParser()
ap.add_argument("--sf", type=float, default=10.0)
ap.add_argument("--query", default="all", choices=["all", "q1", "q6", "q3"])
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--coded", action="store_true")
ap.add_argument("--profile", action="store_true", help="print the per-operator profile of the last query")
args = ap.parse_args()
args.rows = 0
n = B.total_rows(args)
E.init([0])
shape = S.db_shape(n)
host = S.generate_host(shape, range(shape["n_chunks"]), B.SEED, torch.device("cuda", 0))
db = H.Database(0, num_workers=4)
for which, rel in ((H.CUSTOMER, "customer"), (H.ORDERS, "orders"), (H.LINEITEM, "lineitem")):
    db.load(which, host[rel], B.BLOCK_ROWS, H.COMPRESSED_COLUMN_STORE)
if args.coded:
    db.set_code_resident(True)
qs = {"q1": db.q1, "q6": db.q6, "q3": db.q3}
for name, fn in qs.items():
    if args.query in ("all", name):
        fn()                                  # warm
        E.synchronize(0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.reps):
            fn()
        print(f"{name}: {(time.perf_counter() - t0) * 1e3 / args.reps:.3f} ms per query (wall)")
        if args.profile:
            print(db.last_profile())
db.destroy()
