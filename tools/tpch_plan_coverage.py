#!/usr/bin/env python
"""Which of the reference's 22 TPC-H plans the device path could take UNCHANGED.

Test infrastructure / analysis, run in the build container: tests/golden/make_plan_golden.cpp (linked against the
unmodified reference; build it with QS_PLAN_BIN=/tmp/make_plan_golden tests/golden/make_plan_golden.sh) plans every query of
benchmarks/tpch/queries over the full TPC-H catalog with the statistics of --sf, and lowers the whole QueryContext with the
in-tree binding.  A query is "takes the plan as is" when every operator of its DAG has a GPU work order and every
predicate / scalar / aggregation state / filter / join table lowers; otherwise the first thing the binding refuses is
recorded (those operators keep their CPU work orders).  Writes a markdown table.

  python tools/tpch_plan_coverage.py --bin /tmp/make_plan_golden --sf 100 > profiles/r4_tpch_plan_coverage.md
"""
import argparse
import collections
import re
import subprocess

GPU_OPERATORS = {"SelectOperator", "BuildHashOperator", "HashJoinOperator", "HashJoinOperator(LeftSemi)", "HashJoinOperator(LeftAnti)",
                 "HashJoinOperator(LeftOuter)", "DestroyHashOperator", "AggregationOperator", "InitializeAggregationOperator",
                 "FinalizeAggregationOperator", "DestroyAggregationStateOperator", "BuildLIPFilterOperator",
                 "BuildAggregationExistenceMapOperator", "SortRunGenerationOperator", "SortMergeRunOperator"}
HOST_OPERATORS = {"DropTableOperator"}          # catalog bookkeeping: unchanged reference code


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bin", default="/tmp/make_plan_golden")
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--sf", type=float, default=100.0)
    args = ap.parse_args()
    rows, ok = [], 0
    for q in range(1, 23):
        r = subprocess.run([args.bin, "--coverage", args.ref, f"{q:02d}.sql", str(args.sf)], capture_output=True, text=True)
        ops = []
        for line in r.stdout.splitlines():
            if line.startswith("operators:"):
                ops = line.split()[1:]
        ops = [o for o in ops if o not in HOST_OPERATORS]
        count = collections.Counter(ops)
        foreign = sorted(o for o in count if o not in GPU_OPERATORS)
        fatal = re.findall(r"\[FATAL [^\]]*\] (.*)", r.stderr)
        if r.returncode != 0 and not fatal:
            fatal = [(r.stderr.strip().splitlines() or ["planning failed"])[-1][:160]]
        reason = "; ".join(([f"no GPU work order for {', '.join(foreign)}"] if foreign else []) + fatal[:1])
        status = "as is" if not reason else "CPU operators stay"
        ok += not reason
        summary = ", ".join(f"{n}× {o.replace('Operator', '')}" if n > 1 else o.replace("Operator", "") for o, n in count.items())
        rows.append((q, status, summary, reason))
    print(f"# The reference's own TPC-H plans and the device path (statistics of SF{args.sf:g})\n")
    print("Planned by the unmodified reference's parser / optimizer / ExecutionGenerator (`tests/golden/make_plan_golden.cpp --coverage`), "
          "every entry of the resulting `serialization::QueryContext` lowered by `quickstep_b200/host/intree/`. "
          f"**{ok} of 22** plans lower completely and consist only of operators that have a GPU work order; for the others the "
          "first construct the binding refuses is named (its operator keeps the reference's CPU work orders; nothing is silently approximated). "
          "Only Q1, Q3 and Q6 are *executed* on the device in this repository (bench.py, tests); this table is what comes next.\n")
    print("| query | plan | operators (DropTable omitted) | first refusal |")
    print("|---|---|---|---|")
    for q, status, summary, reason in rows:
        print(f"| Q{q} | {status} | {summary} | {reason} |")


if __name__ == "__main__":
    main()
