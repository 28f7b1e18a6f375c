#!/usr/bin/env python
"""Per-operator profile of Q1 / Q6 / Q3 through the C++ operator layer (libqshost.so) on an SF10-shaped
synthetic database with blocks resident in HBM: wall ms per query plus, per operator, work orders, ms inside
execute() and ms inside getAllWorkOrders()."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

import torch

from quickstep_b200 import engine as E
from quickstep_b200 import hostapi as H
from quickstep_b200 import synth as S
from quickstep_b200 import tpch as T

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=59_986_052)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--workers", type=int, default=4)
args = ap.parse_args()
E.init([0])
dev = torch.device("cuda", 0)
cols = S.generate(args.rows, seed=1234, device=dev, key_base=0)
cols.pop("_stats")
db = H.Database(0, num_workers=args.workers)
for which, schema in ((H.CUSTOMER, T.CUSTOMER), (H.ORDERS, T.ORDERS), (H.LINEITEM, T.LINEITEM)):
    db.load(which, [cols[nm].cpu().numpy() for (nm, _t, _w) in schema], 63_000, H.COMPRESSED_COLUMN_STORE)
del cols
torch.cuda.empty_cache()
for name, fn in (("q6", db.q6), ("q1", db.q1), ("q3", db.q3)):
    for _ in range(3):
        fn()
    E.synchronize(0)
    t0 = time.perf_counter()
    for _ in range(args.reps):
        fn()
    E.synchronize(0)
    print(f"== {name}: {(time.perf_counter() - t0) * 1e3 / args.reps:.3f} ms per query (wall)")
    print(db.last_profile())
db.destroy()
