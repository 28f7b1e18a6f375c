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
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import tpch_data as D  # noqa: E402


def main():
    os.environ.setdefault("QS_TPCH_CACHE", "/tmp/qs_tpch_cache")
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
