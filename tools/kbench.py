#!/usr/bin/env python
"""Kernel-only timings of the Q1 / Q6 scan+aggregate kernels (CUDA events around the launch) on an
SF10-shaped synthetic lineitem: the quick loop used while tuning; bench.py is the reported number."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

import numpy as np
import torch

from quickstep_b200 import capi as A
from quickstep_b200 import engine as E
from quickstep_b200 import synth as S
from quickstep_b200 import tpch as T

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=59_986_052)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--queries", default="q1,q6")
args = ap.parse_args()
E.init([0])
dev = torch.device("cuda", 0)
cols = S.generate(args.rows, seed=1234, device=dev, key_base=0)
cols.pop("_stats")
li = S.wrap_relations(E, cols, 0)["lineitem"]
torch.cuda.synchronize()
peak = 6542.4
for name, plan, bpr in (("q1", T.Q1Plan(), T.Q1_BYTES_PER_ROW), ("q6", T.Q6Plan(), T.Q6_BYTES_PER_ROW)):
    if name not in args.queries.split(","):
        continue
    st = E.AggState(plan.strategy, plan.es, plan.pred, plan.aggregates, plan.group_by, estimated=8, dev=0)
    E.set_timing(True)
    xs = []
    for i in range(args.reps + 3):
        st.run(li)
        if i >= 3:
            xs.append(E.last_kernel_ms(A.QS_K_SCAN_AGG))
    E.set_timing(False)
    st.destroy()
    ms = float(np.median(xs))
    gbs = args.rows * bpr / (ms * 1e-3) / 1e9
    print(f"{name}: kernel {ms:.4f} ms (min {min(xs):.4f})  {gbs:.0f} GB/s  {100 * gbs / peak:.1f}% of measured copy peak")
