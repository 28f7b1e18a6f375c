"""Regenerates tests/golden/*.  Run in the build container (needs
/root/reference for dbgen): `python tests/golden/make_golden.py`.

  tpch_sf001.npz      dbgen -s 0.01 (the reference's vendored generator, compiled by
                      oracle/Makefile from /root/reference/benchmarks/tpch/dbgen),
                      the columns Q1/Q3/Q6 touch, in dbgen's row order.
  reference_answers.json
                      * outputs printed by the UNMODIFIED reference binary
                        (quickstep_cli_shell, Release) on dbgen SF1 data, recorded
                        by the survey run (SURVEY.md section 8c / BASELINE.md section 1);
                      * literal expected tables copied as data from the reference's
                        SQL golden tests (query_optimizer/tests/execution_generator/
                        LIP.test:42-151, Select.test:641-683, Partition.test:57-75).

  reference_engine_results.json
                      the COMPLETE result tables of TPC-H Q1 / Q3 / Q6 as printed by the unmodified reference engine
                      (oracle/_ref/quickstep_cli_shell, built by oracle/build_ref.sh; driven by oracle/ref_engine.py the
                      way benchmarks/tpch/run-benchmark.sh does) on dbgen -s 0.01 and -s 1 data, plus its query times
                      in this container.
  ref_blocks_sf0001.npz
                      the block files the engine wrote for customer / orders / lineitem at dbgen -s 0.001 (whole 4 MB
                      images, compressed) and the columns dbgen generated for them: the engine's own physical formats
                      as staging input (tests/test_reference_blocks.py).
"""
import json
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import tpch_data as D  # noqa: E402


def engine_goldens():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
    import ref_blocks as RB
    import ref_engine as R
    if not R.available():
        print("oracle/_ref/quickstep_cli_shell not built: reference_engine_results.json left as it is")
        return
    out = {"source": "oracle/_ref/quickstep_cli_shell (UNMODIFIED reference, Release, -march=x86-64-v3) via oracle/ref_engine.py; "
                     "values exactly as printed", "workers": os.cpu_count()}
    for sf in ("0.01", "1"):
        store = tempfile.mkdtemp(prefix="qs_store_")
        try:
            R.load(sf, store)
            res = {}
            for q in ("01", "03", "06"):
                rows, times = R.run_query(store, q)
                res["q" + q.lstrip("0")] = {"rows": rows, "first_run_ms": times}
            res["timing_ms"] = R.time_queries(store)
            out["sf" + sf] = res
        finally:
            shutil.rmtree(store, ignore_errors=True)
    with open(os.path.join(HERE, "reference_engine_results.json"), "w") as f:
        json.dump(out, f, indent=1)
    # the engine's own block files at SF0.001
    store, tbl = tempfile.mkdtemp(prefix="qs_store_"), tempfile.mkdtemp(prefix="qs_tbl_")
    try:
        R.load("0.001", store, tbl_dir=tbl, workers=1)
        import pandas as pd
        blocks = RB.load_blocks(store)
        arrays = {f"block{i}": np.frombuffer(mem, dtype=np.uint8) for i, (_p, mem) in enumerate(blocks)}
        li = pd.read_csv(os.path.join(tbl, "lineitem.tbl"), sep="|", header=None, usecols=[0, 4, 5, 6, 7, 8, 9, 10], dtype={8: str, 9: str, 10: str})
        od = pd.read_csv(os.path.join(tbl, "orders.tbl"), sep="|", header=None, usecols=[0, 1, 4, 7], dtype={4: str})
        cu = pd.read_csv(os.path.join(tbl, "customer.tbl"), sep="|", header=None, usecols=[0, 6], dtype={6: str})
        arrays.update(l_orderkey=li[0].to_numpy(np.int32), l_quantity=li[4].to_numpy(np.float64), l_extendedprice=li[5].to_numpy(np.float64),
                      l_discount=li[6].to_numpy(np.float64), l_tax=li[7].to_numpy(np.float64), l_returnflag=li[8].to_numpy(dtype="S1"),
                      l_linestatus=li[9].to_numpy(dtype="S1"), l_shipdate=li[10].to_numpy(dtype="datetime64[D]").astype(np.int64),
                      o_orderkey=od[0].to_numpy(np.int32), o_custkey=od[1].to_numpy(np.int32),
                      o_orderdate=od[4].to_numpy(dtype="datetime64[D]").astype(np.int64), o_shippriority=od[7].to_numpy(np.int32),
                      c_custkey=cu[0].to_numpy(np.int32), c_mktsegment=cu[6].to_numpy(dtype="S10"))
        np.savez_compressed(os.path.join(HERE, "ref_blocks_sf0001.npz"), **arrays)
    finally:
        shutil.rmtree(store, ignore_errors=True)
        shutil.rmtree(tbl, ignore_errors=True)


def main():
    os.environ.setdefault("QS_TPCH_CACHE", "/tmp/qs_tpch_cache")
    engine_goldens()
    D.dbgen_tables(0.01)
    z = np.load(os.path.join(D.CACHE, "sf0.01.npz"))
    np.savez_compressed(os.path.join(HERE, "tpch_sf001.npz"), **{k: z[k] for k in z.files})
    answers = {
        "tpch_sf1_reference_binary": {
            "source": "quickstep_cli_shell (unmodified reference, Release) on dbgen -s 1; SURVEY.md 8c, BASELINE.md 1",
            "q6_revenue_printed": "123141078.22829996",
            "q3_first_row": {"l_orderkey": 2456423, "revenue": 406181.0111, "o_orderdate": "1995-03-05",
                             "o_shippriority": 0},
            "q1_groups": ["AF", "NF", "NO", "RF"],
            "lineitem_rows": 6001215,
        },
        "lip_test": {
            "source": "query_optimizer/tests/execution_generator/LIP.test:19-151",
            "R": "x=y=i for i in range(0,100001,2)", "S": "z=i for i in range(0,100001,3)",
            "q1_x_mod_10000": [0, 30000, 60000, 90000],
            "q2_sum_union": 285685710,
        },
        "select_test_groupby": {
            "source": "query_optimizer/tests/execution_generator/Select.test:659-683 over TestDatabaseLoader.cpp:141-183",
            "rows_count_g1_g2": [[1, 3, 6], [1, 3, 7], [2, 4, 8], [1, 4, 9], [1, 5, 10], [1, 5, 11]],
        },
        "partition_test_join": {
            "source": "query_optimizer/tests/execution_generator/Partition.test:57-75",
            "ids": [4, 8, 12, 16, 24, 2, 6, 14, 18, 22],
        },
    }
    with open(os.path.join(HERE, "reference_answers.json"), "w") as f:
        json.dump(answers, f, indent=1)


if __name__ == "__main__":
    main()
