#!/usr/bin/env python
"""Host-side profile of the C-ABI calls a whole query makes: wall time per entry point (call count,
total, mean) for Q1 / Q6 / Q3 on an SF10-shaped synthetic database.  Finds per-query overheads that a
kernel-only timing hides (allocation, synchronisation, result reads)."""
import argparse
import collections
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

import torch

from quickstep_b200 import capi as A
from quickstep_b200 import engine as E
from quickstep_b200 import synth as S
from quickstep_b200 import tpch as T

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=59_986_052)
ap.add_argument("--reps", type=int, default=10)
args = ap.parse_args()

E.init([0])
L = A.load()
stats = collections.OrderedDict()


class Timed:
    def __init__(self, name, fn):
        self.name, self.fn = name, fn

    def __call__(self, *a):
        t0 = time.perf_counter()
        r = self.fn(*a)
        dt = time.perf_counter() - t0
        s = stats.setdefault(self.name, [0, 0.0])
        s[0] += 1
        s[1] += dt
        return r


class Proxy:
    def __init__(self, lib):
        self._lib, self._cache = lib, {}

    def __getattr__(self, name):
        if name not in self._cache:
            self._cache[name] = Timed(name, getattr(self._lib, name))
        return self._cache[name]


proxy = Proxy(L)
A.load = lambda: proxy        # engine.py resolves the library through capi.load() on every call

dev = torch.device("cuda", 0)
cols = S.generate(args.rows, seed=1234, device=dev, key_base=0)
st = cols.pop("_stats")
rels = S.wrap_relations(E, cols, 0)
li = rels["lineitem"]
torch.cuda.synchronize()
q1p, q6p, q3p = T.Q1Plan(), T.Q6Plan(), T.Q3Plan()
for name, fn in (("q6", lambda: T.run_q6(li, q6p)), ("q1", lambda: T.run_q1(li, q1p)),
                 ("q3", lambda: T.run_q3(rels["customer"], rels["orders"], li, st, q3p))):
    for _ in range(3):
        fn()
    stats.clear()
    E.synchronize(0)
    t0 = time.perf_counter()
    for _ in range(args.reps):
        fn()
    E.synchronize(0)
    wall = (time.perf_counter() - t0) * 1e3 / args.reps
    inside = sum(v[1] for v in stats.values()) * 1e3 / args.reps
    print(f"== {name}: {wall:.3f} ms per query wall, {inside:.3f} ms inside C-ABI calls")
    for k, (n, t) in sorted(stats.items(), key=lambda kv: -kv[1][1]):
        print(f"   {k:32s} calls/query {n / args.reps:6.1f}  ms/query {t * 1e3 / args.reps:9.3f}  us/call {t * 1e6 / n:9.1f}")
