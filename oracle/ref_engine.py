#!/usr/bin/env python
"""TEST INFRASTRUCTURE (oracle side).  Drives the UNMODIFIED reference engine, oracle/_ref/quickstep_cli_shell (built by
oracle/build_ref.sh from /root/reference), the way benchmarks/tpch/run-benchmark.sh does:

    create.sql -> COPY <table> FROM '<dbgen .tbl>' WITH (DELIMITER '|') -> \\analyze -> queries/NN.sql

  load(sf, storage_dir)          dbgen -s sf (oracle/_ref/dbgen) + the three COPYs the hot path needs + \\analyze
  run_query(storage_dir, "06")   -> (result rows as printed, [Time: ... ms values])
  time_queries(storage_dir, ...) the benchmark procedure: 5 runs, mean of the middle 3 (benchmarks/tpch/process.py:33,39)

Used by tests/golden/make_golden.py (golden result tables), tests/test_reference_blocks.py (the engine's own 4 MB block
files as staging input) and tools/ref_engine_bench.py (the reference's CPU time beside the GPU numbers).  Never imported
by anything under quickstep_b200/."""
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("QS_REFERENCE", "/root/reference")
CLI = os.path.join(HERE, "_ref", "quickstep_cli_shell")
DBGEN = os.path.join(HERE, "_ref", "dbgen")
DISTS = os.path.join(HERE, "_ref", "dists.dss")

# benchmarks/tpch/create.sql:18-114 restricted to the three relations the hot path reads (same column types,
# same block layouts: lineitem / orders compressed column store sorted on the order key, customer split row store).
# Kept here as data (DDL text) because the GPU box has no /root/reference; verified against the reference's file by
# tests/test_oracle_golden.py when the tree is present.
CREATE_SQL = """
CREATE TABLE customer (
  c_custkey INT NOT NULL, c_name VARCHAR(25) NOT NULL, c_address VARCHAR(40) NOT NULL, c_nationkey INT NOT NULL,
  c_phone CHAR(15) NOT NULL, c_acctbal DECIMAL NOT NULL, c_mktsegment CHAR(10) NOT NULL, c_comment VARCHAR(117) NOT NULL
) WITH BLOCKPROPERTIES (TYPE split_rowstore, BLOCKSIZEMB 4);
CREATE TABLE orders (
  o_orderkey INT NOT NULL, o_custkey INT NOT NULL, o_orderstatus CHAR(1) NOT NULL, o_totalprice DECIMAL NOT NULL,
  o_orderdate DATE NOT NULL, o_orderpriority CHAR(15) NOT NULL, o_clerk CHAR(15) NOT NULL, o_shippriority INT NOT NULL,
  o_comment VARCHAR(79) NOT NULL
) WITH BLOCKPROPERTIES (TYPE compressed_columnstore, SORT o_orderkey, COMPRESS ALL, BLOCKSIZEMB 4);
CREATE TABLE lineitem (
  l_orderkey INT NOT NULL, l_partkey INT NOT NULL, l_suppkey INT NOT NULL, l_linenumber INT NOT NULL,
  l_quantity DECIMAL NOT NULL, l_extendedprice DECIMAL NOT NULL, l_discount DECIMAL NOT NULL, l_tax DECIMAL NOT NULL,
  l_returnflag CHAR(1) NOT NULL, l_linestatus CHAR(1) NOT NULL, l_shipdate DATE NOT NULL, l_commitdate DATE NOT NULL,
  l_receiptdate DATE NOT NULL, l_shipinstruct CHAR(25) NOT NULL, l_shipmode CHAR(10) NOT NULL, l_comment VARCHAR(44) NOT NULL
) WITH BLOCKPROPERTIES (TYPE compressed_columnstore, SORT l_orderkey, COMPRESS ALL, BLOCKSIZEMB 4);
"""

# benchmarks/tpch/queries/{01,03,06}.sql (TPC-H validation parameters)
QUERIES = {
    "01": """SELECT l_returnflag, l_linestatus, SUM(l_quantity) AS sum_qty, SUM(l_extendedprice) AS sum_base_price,
  SUM(l_extendedprice * (1 - l_discount)) AS sum_disc_price, SUM(l_extendedprice * (1 - l_discount) * (1 + l_tax)) AS sum_charge,
  AVG(l_quantity) AS avg_qty, AVG(l_extendedprice) AS avg_price, AVG(l_discount) AS avg_disc, COUNT(*) AS count_order
FROM lineitem WHERE l_shipdate <= DATE '1998-09-01'
GROUP BY l_returnflag, l_linestatus ORDER BY l_returnflag, l_linestatus;""",
    "03": """SELECT l_orderkey, SUM(l_extendedprice * (1 - l_discount)) AS revenue, o_orderdate, o_shippriority
FROM customer, orders, lineitem
WHERE c_mktsegment = 'BUILDING' AND c_custkey = o_custkey AND l_orderkey = o_orderkey
  AND o_orderdate < DATE '1995-03-15' AND l_shipdate > DATE '1995-03-15'
GROUP BY l_orderkey, o_orderdate, o_shippriority ORDER BY revenue DESC, o_orderdate LIMIT 10;""",
    "06": """SELECT SUM(l_extendedprice * l_discount) AS revenue FROM lineitem
WHERE l_shipdate >= DATE '1994-01-01' AND l_shipdate < DATE '1994-01-01' + INTERVAL '1' YEAR
  AND l_discount BETWEEN 0.05 AND 0.07 AND l_quantity < 24;""",
}


def available() -> bool:
    return os.path.exists(CLI) and os.path.exists(DBGEN)


def _cli(storage, sql, workers=None, printing=True, timeout=3600):
    args = [CLI, f"-storage_path={storage.rstrip('/')}/", f"-num_workers={workers or os.cpu_count() or 1}"]
    if not printing:
        args.append("-printing_enabled=false")
    r = subprocess.run(args, input=sql, capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError(f"quickstep_cli_shell failed ({r.returncode}): {r.stderr[-2000:]}\n{r.stdout[-2000:]}")
    return r.stdout


def load(sf, storage, tbl_dir=None, workers=None):
    """Fresh database of the three relations at scale factor `sf`.  -> directory holding the .tbl files."""
    own = tbl_dir is None
    tbl_dir = tbl_dir or tempfile.mkdtemp(prefix="qs_tbl_")
    for t in ("c", "O", "L"):          # customer, orders, lineitem
        subprocess.check_call([DBGEN, "-f", "-q", "-s", str(sf), "-T", t, "-b", DISTS], cwd=tbl_dir,
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    os.makedirs(storage, exist_ok=True)
    args = [CLI, f"-storage_path={storage.rstrip('/')}/", f"-num_workers={workers or os.cpu_count() or 1}", "-initialize_db=true"]
    r = subprocess.run(args, input=CREATE_SQL, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("create failed: " + r.stderr[-2000:])
    for rel in ("customer", "orders", "lineitem"):
        _cli(storage, f"COPY {rel} FROM '{os.path.join(tbl_dir, rel + '.tbl')}' WITH (DELIMITER '|');\n", workers)
    _cli(storage, "\\analyze\n", workers)
    if own:
        for rel in ("customer", "orders", "lineitem"):
            os.remove(os.path.join(tbl_dir, rel + ".tbl"))
    return tbl_dir


_ROW = re.compile(r"^\|(.*)\|\s*$")


def parse_tables(out):
    """The CLI's ASCII result tables -> list of tables, each a list of rows of stripped cell strings (header dropped)."""
    tables, cur, header_seen = [], None, False
    for line in out.splitlines():
        line = re.sub(r"^(quickstep>\s*|\s*\.\.\.>\s*)+", "", line)        # the prompt precedes the first table line
        if line.startswith("+"):
            if cur is None:
                cur, header_seen = [], False
            continue
        m = _ROW.match(line)
        if m and cur is not None:
            cells = [c.strip() for c in m.group(1).split("|")]
            if not header_seen:
                header_seen = True
            else:
                cur.append(cells)
            continue
        if cur is not None:
            tables.append(cur)
            cur = None
    if cur is not None:
        tables.append(cur)
    return tables


def run_query(storage, q, workers=None):
    out = _cli(storage, QUERIES[q] + "\n", workers)
    times = [float(x) for x in re.findall(r"Time: ([0-9.]+) ms", out)]
    tabs = parse_tables(out)
    return (tabs[0] if tabs else []), times


def time_queries(storage, qs=("01", "06", "03"), workers=None, runs=5):
    """run-benchmark.sh's procedure: each query `runs` times in one session, results not printed; the mean of the middle
    runs (drop min and max) is the number (benchmarks/tpch/process.py:33-39)."""
    res = {}
    for q in qs:
        out = _cli(storage, (QUERIES[q] + "\n") * runs, workers, printing=False)
        t = sorted(float(x) for x in re.findall(r"Time: ([0-9.]+) ms", out))
        mid = t[1:-1] if len(t) > 2 else t
        res[q] = {"runs_ms": t, "ms": sum(mid) / max(1, len(mid))}
    return res


if __name__ == "__main__":
    sf, storage = sys.argv[1], sys.argv[2]
    load(sf, storage)
    for q in ("06", "01", "03"):
        rows, times = run_query(storage, q)
        print(q, times, rows[:3])
