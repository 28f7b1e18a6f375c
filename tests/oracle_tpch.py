"""TPC-H Q1/Q6/Q3 executed by the CPU oracle over the same plan descriptions
(quickstep_b200.tpch.*Plan) the device path runs.  Test infrastructure."""
from __future__ import annotations

import os
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (_ROOT, os.path.join(_ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import qs_oracle as O  # noqa: E402
from quickstep_b200 import capi as A  # noqa: E402
from quickstep_b200 import tpch as T  # noqa: E402
from quickstep_b200.table import Column, HostTable  # noqa: E402


def q6(lineitem: HostTable, plan=None):
    plan = plan or T.Q6Plan()
    r = O.aggregate(plan.es, plan.pred, plan.aggregates, [], lineitem)
    return float(r.values[0][0]), r.is_null[0]


def q1(lineitem: HostTable, plan=None):
    plan = plan or T.Q1Plan()
    r = O.aggregate(plan.es, plan.pred, plan.aggregates, plan.group_by, lineitem)
    flag = r.keys[:, 0:1].copy().view("S1").reshape(-1)
    status = r.keys[:, 1:2].copy().view("S1").reshape(-1)
    return T.q1_rows_from_states(flag, status, r.values[:5], r.values[5])


def q3(tables, stats, plan=None, info=None):
    plan = plan or T.Q3Plan()
    cust, orders, lineitem = tables["customer"], tables["orders"], tables["lineitem"]
    f_cust = O.Lip(A.QS_LIP_BITVECTOR_EXACT, stats["c_custkey_min"], stats["c_custkey_max"])
    f_ord = O.Lip(A.QS_LIP_BITVECTOR_EXACT, stats["o_orderkey_min"], stats["o_orderkey_max"])
    O.build_lip_filter(plan.e1, plan.p1, cust, None, [(f_cust, plan.c_custkey)])
    t2c = O.select(plan.e2, plan.p2, orders, [(f_cust, plan.o_custkey)], plan.proj2, plan.t2_schema)
    t2 = HostTable("t2", [Column(f"c{i}", t, t2c[i], w) for i, (t, w) in enumerate(plan.t2_schema)])
    # BuildHash also feeds the exact filter on o_orderkey
    O.build_lip_filter(None, -1, t2, None, [(f_ord, 0)])
    t0c = O.select(plan.e0, plan.p0, lineitem, [(f_ord, plan.l_orderkey)], plan.proj0, plan.t0_schema)
    t0 = HostTable("t0", [Column(f"c{i}", t, t0c[i], w) for i, (t, w) in enumerate(plan.t0_schema)])
    t4c = O.hash_join(plan.e4, t2, -1, 0, t0, -1, 0, None, A.QS_JOIN_INNER, -1, plan.proj4, plan.t4_schema,
                      max(1, t0.n_rows))
    t4 = HostTable("t4", [Column(f"c{i}", t, t4c[i], w) for i, (t, w) in enumerate(plan.t4_schema)])
    r = O.aggregate(plan.e6, -1, plan.aggregates, plan.group_by, t4)
    G = r.n_groups
    okey = r.keys[:, 0:4].copy().view("<i4").reshape(-1)
    odate = r.keys[:, 4:12].copy().view(np.dtype([("year", "<i4"), ("month", "u1"), ("day", "u1"), ("pad", "<u2")])).reshape(-1)
    prio = r.keys[:, 12:16].copy().view("<i4").reshape(-1)
    fin = HostTable("fin", [Column("l_orderkey", A.QS_INT, okey), Column("o_orderdate", A.QS_DATE, odate),
                            Column("o_shippriority", A.QS_INT, prio), Column("revenue", A.QS_DOUBLE, r.values[0])])
    if info is not None:
        info.update(t2_rows=t2.n_rows, t0_rows=t0.n_rows, t4_rows=t4.n_rows, groups=G, fin=fin, t4=t4, t2=t2, t0=t0,
                    f_cust=f_cust, f_ord=f_ord)
    ids = O.topk(fin, plan.sort_keys, plan.limit)
    return [(int(okey[i]), float(r.values[0][i]), (int(odate[i]["year"]), int(odate[i]["month"]), int(odate[i]["day"])),
             int(prio[i])) for i in ids]
