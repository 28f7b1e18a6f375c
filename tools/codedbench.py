"""Q1 / Q6 over a native lineitem vs the same relation resident as dictionary codes (one GPU).
Prints one JSON line: kernel-only and whole-query times, bytes per row, and that the answers agree."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=59_986_052)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--coded", default="l_quantity,l_discount,l_tax,l_shipdate")
    args = ap.parse_args()
    import numpy as np
    import torch
    from quickstep_b200 import capi as A
    from quickstep_b200 import engine as E
    from quickstep_b200 import synth as S
    from quickstep_b200 import tpch as T
    dev = torch.device("cuda", 0)
    E.init([0])
    cols = S.generate(args.rows, seed=1234, device=dev)
    cols.pop("_stats")
    nat = S.wrap_relations(E, cols, 0)["lineitem"]
    cod, info = S.wrap_lineitem_coded(E, cols, 0, tuple(args.coded.split(",")))
    torch.cuda.synchronize()
    q1p, q6p = T.Q1Plan(), T.Q6Plan()

    def kernel_ms(plan, rel):
        st = E.AggState(plan.strategy, plan.es, plan.pred, plan.aggregates, plan.group_by, estimated=8, dev=0)
        E.set_timing(True)
        xs = []
        try:
            for i in range(args.reps + 3):
                st.run(rel)
                if i >= 3:
                    xs.append(E.last_kernel_ms(A.QS_K_SCAN_AGG))
        finally:
            E.set_timing(False)
            st.destroy()
        return float(np.mean(xs)), float(np.min(xs))

    def query_ms(fn, rel):
        for _ in range(3):
            fn(rel)
        E.synchronize(0)
        E.timer_start(0)
        for _ in range(args.reps):
            out = fn(rel)
        return E.timer_stop(0) / args.reps, out

    width = {n: w for (n, _t, w) in T.LINEITEM}
    def bpr(names):
        return sum(info[n][0] if n in info else width[n] for n in names)
    q1_cols = ["l_shipdate", "l_returnflag", "l_linestatus", "l_quantity", "l_extendedprice", "l_discount", "l_tax"]
    q6_cols = ["l_shipdate", "l_discount", "l_quantity", "l_extendedprice"]
    out = {"rows": args.rows, "dictionaries": info, "bytes_per_row": {"q1": bpr(q1_cols), "q6": bpr(q6_cols), "q1_native": 42, "q6_native": 32}}
    for name, plan, fn in (("q1", q1p, lambda r: T.run_q1(r, q1p)), ("q6", q6p, lambda r: T.run_q6(r, q6p))):
        kn, kc = kernel_ms(plan, nat), kernel_ms(plan, cod)
        qn, rn = query_ms(fn, nat)
        qc, rc = query_ms(fn, cod)
        out[name] = {"kernel_ms_native": kn[0], "kernel_ms_coded": kc[0], "kernel_min_native": kn[1], "kernel_min_coded": kc[1],
                     "query_ms_native": qn, "query_ms_coded": qc}
        if name == "q1":
            assert [r["count_order"] for r in rn] == [r["count_order"] for r in rc]
            for a, b in zip(rn, rc):
                for k in ("sum_qty", "sum_base_price", "sum_disc_price", "sum_charge"):
                    assert abs(a[k] - b[k]) <= 1e-9 * abs(b[k]), (k, a[k], b[k])
        else:
            assert abs(rn[0] - rc[0]) <= 1e-9 * abs(rn[0]), (rn, rc)
        out[name]["answers_agree"] = True
    print(json.dumps(out))


if __name__ == "__main__":
    main()
