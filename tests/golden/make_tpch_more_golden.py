"""TEST INFRASTRUCTURE, run in the build container: python tests/golden/make_tpch_more_golden.py

The other TPC-H queries whose reference plans lower completely with the in-tree binding (profiles/r4_tpch_plan_coverage.md:
Q4, Q5, Q17, Q19, Q21 besides Q1 / Q3 / Q6), as the UNMODIFIED engine answers them:

  tpch_full_sf001.npz                  dbgen -s 0.01, all eight relations, every fixed-width attribute (VARCHAR attributes
                                       are left out: the device path does not stage them and no plan here reads one)
  reference_engine_results_more.json   the tables oracle/_ref/quickstep_cli_shell prints for benchmarks/tpch/queries/
                                       {04,05,17,19,21}.sql over that database (benchmarks/tpch/create.sql read where it
                                       lies, COPY, \\analyze: run-benchmark.sh's procedure)
  reference_plans_more.json            the same queries planned by the reference's optimizer over the full catalog and
                                       lowered by quickstep_b200/host/intree (tests/golden/make_plan_golden.cpp --plan)
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_engine as R  # noqa: E402

REF = os.environ.get("QS_REFERENCE", "/root/reference")
QUERIES = ["04", "05", "17", "19", "21"]
# benchmarks/tpch/create.sql: attribute names in order; "V" marks VARCHAR (skipped), ("C", n) CHAR(n), "I" INT, "D" DECIMAL = DOUBLE, "T" DATE
SCHEMA = {
    "region": [("r_regionkey", "I"), ("r_name", ("C", 25)), ("r_comment", "V")],
    "nation": [("n_nationkey", "I"), ("n_name", ("C", 25)), ("n_regionkey", "I"), ("n_comment", "V")],
    "supplier": [("s_suppkey", "I"), ("s_name", ("C", 25)), ("s_address", "V"), ("s_nationkey", "I"), ("s_phone", ("C", 15)), ("s_acctbal", "D"),
                 ("s_comment", "V")],
    "customer": [("c_custkey", "I"), ("c_name", "V"), ("c_address", "V"), ("c_nationkey", "I"), ("c_phone", ("C", 15)), ("c_acctbal", "D"),
                 ("c_mktsegment", ("C", 10)), ("c_comment", "V")],
    "part": [("p_partkey", "I"), ("p_name", "V"), ("p_mfgr", ("C", 25)), ("p_brand", ("C", 10)), ("p_type", "V"), ("p_size", "I"),
             ("p_container", ("C", 10)), ("p_retailprice", "D"), ("p_comment", "V")],
    "partsupp": [("ps_partkey", "I"), ("ps_suppkey", "I"), ("ps_availqty", "I"), ("ps_supplycost", "D"), ("ps_comment", "V")],
    "orders": [("o_orderkey", "I"), ("o_custkey", "I"), ("o_orderstatus", ("C", 1)), ("o_totalprice", "D"), ("o_orderdate", "T"),
               ("o_orderpriority", ("C", 15)), ("o_clerk", ("C", 15)), ("o_shippriority", "I"), ("o_comment", "V")],
    "lineitem": [("l_orderkey", "I"), ("l_partkey", "I"), ("l_suppkey", "I"), ("l_linenumber", "I"), ("l_quantity", "D"), ("l_extendedprice", "D"),
                 ("l_discount", "D"), ("l_tax", "D"), ("l_returnflag", ("C", 1)), ("l_linestatus", ("C", 1)), ("l_shipdate", "T"),
                 ("l_commitdate", "T"), ("l_receiptdate", "T"), ("l_shipinstruct", ("C", 25)), ("l_shipmode", ("C", 10)), ("l_comment", "V")],
}


def main():
    assert R.available(), "build oracle/_ref first (oracle/build_ref.sh)"
    tbl = tempfile.mkdtemp(prefix="qs_tbl_")
    store = tempfile.mkdtemp(prefix="qs_store_")
    try:
        subprocess.check_call([R.DBGEN, "-f", "-q", "-s", "0.01", "-b", R.DISTS], cwd=tbl, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        # ---- the columns
        arrays = {}
        for rel, attrs in SCHEMA.items():
            df = pd.read_csv(os.path.join(tbl, rel + ".tbl"), sep="|", header=None, usecols=range(len(attrs)), dtype=str, keep_default_na=False)
            for i, (name, kind) in enumerate(attrs):
                col = df[i]
                if kind == "V":
                    continue
                if kind == "I":
                    arrays[name] = col.astype(np.int32).to_numpy()
                elif kind == "D":
                    arrays[name] = col.astype(np.float64).to_numpy()
                elif kind == "T":
                    arrays[name] = (pd.to_datetime(col).to_numpy().astype("datetime64[D]").astype(np.int64)).astype(np.int32)      # days since 1970
                else:
                    arrays[name] = col.to_numpy(dtype=f"S{kind[1]}")
        np.savez_compressed(os.path.join(HERE, "tpch_full_sf001.npz"), **arrays)
        # ---- the engine: the reference's own create.sql, all eight COPYs, \analyze, the queries as they lie
        args = [R.CLI, f"-storage_path={store}/", "-num_workers=4", "-initialize_db=true"]
        r = subprocess.run(args, input=open(os.path.join(REF, "benchmarks/tpch/create.sql")).read(), capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        for rel in SCHEMA:
            R._cli(store, f"COPY {rel} FROM '{os.path.join(tbl, rel + '.tbl')}' WITH (DELIMITER '|');\n", 4)
        R._cli(store, "\\analyze\n", 4)
        out = {"source": "oracle/_ref/quickstep_cli_shell (UNMODIFIED reference) over dbgen -s 0.01, all eight relations of "
                         "benchmarks/tpch/create.sql; benchmarks/tpch/queries/NN.sql as they lie; values exactly as printed"}
        for q in QUERIES:
            sql = open(os.path.join(REF, "benchmarks/tpch/queries", q + ".sql")).read()
            tabs = R.parse_tables(R._cli(store, sql + "\n", 4))
            out["q" + q.lstrip("0")] = {"rows": tabs[0] if tabs else []}
            print(q, len(out["q" + q.lstrip("0")]["rows"]), "rows", out["q" + q.lstrip("0")]["rows"][:2])
        json.dump(out, open(os.path.join(HERE, "reference_engine_results_more.json"), "w"), indent=1)
    finally:
        shutil.rmtree(tbl, ignore_errors=True)
        shutil.rmtree(store, ignore_errors=True)


if __name__ == "__main__":
    main()
