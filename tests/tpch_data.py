"""TPC-H data for the parity tests and the bench.

Two sources, both local (no network):
  * dbgen_tables(sf): the reference's vendored generator, compiled from
    /root/reference/benchmarks/tpch/dbgen into oracle/_ref/dbgen by oracle/Makefile
    (deterministic per scale factor); parsed into Quickstep's native column layouts.
  * synthetic_tables(...): a seeded numpy generator with dbgen's value domains
    (SURVEY.md section 8d "Data"): quantity 1..50, discount 0.00..0.10,
    tax 0.00..0.08, dates 1992-01-02..1998-12-01, flags R/A/N x O/F, sparse order keys.
"""
from __future__ import annotations

import os
import subprocess
import sys
import tempfile

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from quickstep_b200 import capi as A  # noqa: E402
from quickstep_b200.table import Column, HostTable, days_to_dates  # noqa: E402
from quickstep_b200 import tpch as T  # noqa: E402

DBGEN = os.path.join(_ROOT, "oracle", "_ref", "dbgen")
DISTS = os.path.join(_ROOT, "oracle", "_ref", "dists.dss")
CACHE = os.environ.get("QS_TPCH_CACHE", "/tmp/qs_tpch_cache")


def have_dbgen() -> bool:
    return os.path.exists(DBGEN) and os.path.exists(DISTS)


def _parse_dates(col) -> np.ndarray:
    d = np.asarray(col, dtype="datetime64[D]")
    return days_to_dates(d.astype(np.int64))


def _table(name, schema, arrays) -> HostTable:
    return HostTable(name, [Column(n, t, arrays[n], w) for (n, t, w) in schema])


def dbgen_tables(sf: float):
    """-> dict(customer, orders, lineitem) of HostTable, exactly dbgen's rows in dbgen's order."""
    import pandas as pd

    os.makedirs(CACHE, exist_ok=True)
    tag = os.path.join(CACHE, f"sf{sf}")
    if not os.path.exists(tag + ".npz"):
        if not have_dbgen():
            raise RuntimeError("oracle/_ref/dbgen is not built (make -C oracle ref)")
        with tempfile.TemporaryDirectory(dir=CACHE) as tmp:
            for t in ("c", "O", "L"):   # customer, orders, lineitem
                subprocess.check_call([DBGEN, "-f", "-q", "-s", str(sf), "-T", t, "-b", DISTS], cwd=tmp,
                                      stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            li = pd.read_csv(os.path.join(tmp, "lineitem.tbl"), sep="|", header=None, usecols=[0, 4, 5, 6, 7, 8, 9, 10],
                             names=["l_orderkey", "l_quantity", "l_extendedprice", "l_discount", "l_tax",
                                    "l_returnflag", "l_linestatus", "l_shipdate"], dtype={8: str, 9: str, 10: str})
            od = pd.read_csv(os.path.join(tmp, "orders.tbl"), sep="|", header=None, usecols=[0, 1, 4, 7],
                             names=["o_orderkey", "o_custkey", "o_orderdate", "o_shippriority"], dtype={4: str})
            cu = pd.read_csv(os.path.join(tmp, "customer.tbl"), sep="|", header=None, usecols=[0, 6],
                             names=["c_custkey", "c_mktsegment"], dtype={6: str})
            arrays = {}
            for c in ("l_orderkey",):
                arrays[c] = li[c].to_numpy(np.int32)
            for c in ("l_quantity", "l_extendedprice", "l_discount", "l_tax"):
                arrays[c] = li[c].to_numpy(np.float64)      # text -> double, as the reference's COPY does
            arrays["l_returnflag"] = li["l_returnflag"].to_numpy(dtype="S1")
            arrays["l_linestatus"] = li["l_linestatus"].to_numpy(dtype="S1")
            arrays["l_shipdate"] = li["l_shipdate"].to_numpy(dtype="datetime64[D]").astype(np.int64)
            arrays["o_orderkey"] = od["o_orderkey"].to_numpy(np.int32)
            arrays["o_custkey"] = od["o_custkey"].to_numpy(np.int32)
            arrays["o_orderdate"] = od["o_orderdate"].to_numpy(dtype="datetime64[D]").astype(np.int64)
            arrays["o_shippriority"] = od["o_shippriority"].to_numpy(np.int32)
            arrays["c_custkey"] = cu["c_custkey"].to_numpy(np.int32)
            arrays["c_mktsegment"] = cu["c_mktsegment"].to_numpy(dtype="S10")
            np.savez(tag + ".npz", **arrays)
    z = np.load(tag + ".npz")
    arrays = {k: z[k] for k in z.files}
    arrays["l_shipdate"] = days_to_dates(arrays["l_shipdate"])
    arrays["o_orderdate"] = days_to_dates(arrays["o_orderdate"])
    return tables_from_arrays(arrays)


def golden_tables():
    """tests/golden/tpch_sf001.npz: dbgen -s 0.01, committed (see tests/golden/make_golden.py)."""
    z = np.load(os.path.join(_ROOT, "tests", "golden", "tpch_sf001.npz"))
    arrays = {k: z[k] for k in z.files}
    arrays["l_shipdate"] = days_to_dates(arrays["l_shipdate"])
    arrays["o_orderdate"] = days_to_dates(arrays["o_orderdate"])
    return tables_from_arrays(arrays)


def tables_from_arrays(arrays):
    return dict(customer=_table("customer", T.CUSTOMER, arrays), orders=_table("orders", T.ORDERS, arrays),
                lineitem=_table("lineitem", T.LINEITEM, arrays))


_D0 = int(np.datetime64("1992-01-01").astype(np.int64))


def synthetic_lineitem_arrays(n: int, seed: int = 1, max_orderkey: int | None = None):
    """TPC-H-shaped lineitem columns (dbgen value domains), sorted on l_orderkey."""
    rng = np.random.default_rng(seed)
    n_orders = max(1, n // 4)
    ok_dense = np.sort(rng.integers(0, n_orders, size=n))
    # sparse keys: 8 used of every 32 (dbgen mk_sparse)
    okey = ((ok_dense >> 3) << 5 | (ok_dense & 7)) + 1
    qty = rng.integers(1, 51, size=n).astype(np.float64)
    price = np.round(rng.integers(90000, 10494951, size=n) / 100.0, 2)
    disc = rng.integers(0, 11, size=n) / 100.0
    tax = rng.integers(0, 9, size=n) / 100.0
    ship = _D0 + rng.integers(1, 2526, size=n)           # 1992-01-02 .. 1998-12-01
    cutoff = int(np.datetime64("1995-06-17").astype(np.int64))
    shipped_early = ship <= cutoff
    rf = np.where(shipped_early, np.where(rng.random(n) < 0.5, b"R", b"A"), b"N").astype("S1")
    ls = np.where(shipped_early, b"F", b"O").astype("S1")
    return dict(l_orderkey=okey.astype(np.int32), l_quantity=qty, l_extendedprice=price, l_discount=disc,
                l_tax=tax, l_returnflag=rf, l_linestatus=ls, l_shipdate=days_to_dates(ship)), n_orders


def synthetic_tables(n_lineitem: int, seed: int = 1):
    """customer : orders : lineitem = 0.025 : 0.25 : 1 like TPC-H."""
    rng = np.random.default_rng(seed + 7)
    arrays, n_orders = synthetic_lineitem_arrays(n_lineitem, seed)
    n_cust = max(1, n_orders // 10)
    o_dense = np.arange(n_orders)
    arrays["o_orderkey"] = (((o_dense >> 3) << 5 | (o_dense & 7)) + 1).astype(np.int32)
    arrays["o_custkey"] = rng.integers(1, n_cust + 1, size=n_orders).astype(np.int32)
    arrays["o_orderdate"] = days_to_dates(_D0 + rng.integers(0, 2406, size=n_orders))
    arrays["o_shippriority"] = np.zeros(n_orders, dtype=np.int32)
    arrays["c_custkey"] = np.arange(1, n_cust + 1, dtype=np.int32)
    segs = np.array([b"AUTOMOBILE", b"BUILDING", b"FURNITURE", b"MACHINERY", b"HOUSEHOLD"], dtype="S10")
    arrays["c_mktsegment"] = segs[rng.integers(0, 5, size=n_cust)]
    return tables_from_arrays(arrays)


def q3_stats(tables) -> dict:
    """What \\analyze would record (exact min/max, row counts)."""
    c, o, l = tables["customer"], tables["orders"], tables["lineitem"]
    return dict(c_custkey_min=int(c.col("c_custkey").data.min()), c_custkey_max=int(c.col("c_custkey").data.max()),
                o_orderkey_min=int(o.col("o_orderkey").data.min()), o_orderkey_max=int(o.col("o_orderkey").data.max()),
                orders_rows=o.n_rows, lineitem_rows=l.n_rows, customer_rows=c.n_rows,
                t2_estimate=max(1024, o.n_rows // 2), groups_estimate=max(1024, o.n_rows // 4))
