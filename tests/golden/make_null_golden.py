"""Regenerates tests/golden/ref_null_blocks.npz and tests/golden/ref_null_results.json.  Run in the build container
after oracle/build_ref.sh: `python tests/golden/make_null_golden.py`.

One relation with NULL-able attributes,  t (g INT NOT NULL, x DOUBLE NULL, y INT NULL, c CHAR(4) NULL),  is loaded
into the UNMODIFIED reference engine (oracle/_ref/quickstep_cli_shell) in each of its three fixed-width layouts:

    t_row   split row store              NULLs = a BitVector<true> at the head of every tuple slot
    t_col   column store, sorted on g    NULLs = one BitVector<false> per NULL-able attribute
    t_cmp   compressed column store      NULLs = the dictionary's null code / a bitmap for uncompressed attributes

through `COPY ... FROM` a text file with \\N markers.  The engine's block files (2 MB each) and the tables it prints
for the queries in QUERIES are the fixture: tests/test_reference_nulls.py stages the blocks with their NULL
representations (qs_stage_desc.null_kind), checks the decoded columns and masks against the source data, runs the
same queries through the C ABI and compares every cell with what the engine printed.
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_engine as R  # noqa: E402
import ref_blocks as RB  # noqa: E402

N_ROWS = 5000
TABLES = {
    "t_row": "WITH BLOCKPROPERTIES (TYPE split_rowstore, BLOCKSIZEMB 2)",
    "t_col": "WITH BLOCKPROPERTIES (TYPE columnstore, SORT g, BLOCKSIZEMB 2)",
    "t_cmp": "WITH BLOCKPROPERTIES (TYPE compressed_columnstore, SORT g, COMPRESS ALL, BLOCKSIZEMB 2)",
}
QUERIES = {
    "single": "SELECT COUNT(*), COUNT(x), SUM(x), AVG(x), MIN(y), MAX(y), SUM(x * y) FROM {t} WHERE y > -500 OR NOT x < 12;",
    "grouped": "SELECT g, COUNT(x), SUM(x), AVG(y), MIN(x), COUNT(c) FROM {t} GROUP BY g ORDER BY g;",
    "single_all_null": "SELECT AVG(x), MAX(x), SUM(x * y), COUNT(*), SUM(x), COUNT(x) FROM {t} WHERE g = 3;",
    "grouped_all_null": "SELECT g, AVG(x), MAX(x), SUM(x * y), COUNT(*), SUM(y) FROM {t} WHERE g = 3 OR g = 4 GROUP BY g ORDER BY g;",
    "not_lt": "SELECT COUNT(*) FROM {t} WHERE NOT x < 12;",
    "ge": "SELECT COUNT(*) FROM {t} WHERE x >= 12;",
    "lt": "SELECT COUNT(*) FROM {t} WHERE x < 12;",
    "char_eq": "SELECT COUNT(*) FROM {t} WHERE c = 'ab';",
    "char_not_eq": "SELECT COUNT(*) FROM {t} WHERE NOT c = 'ab';",
    "attr_vs_attr": "SELECT COUNT(*), SUM(y) FROM {t} WHERE x < y;",
    "group_by_nullable": "SELECT y, COUNT(*), SUM(x) FROM {t} WHERE g = 0 AND x < 2 GROUP BY y ORDER BY y;",
    "group_by_nullable_counts": "SELECT COUNT(*), COUNT(y) FROM {t} WHERE g = 0 AND x < 2;",
}


def source_data():
    rng = np.random.default_rng(1)
    n = N_ROWS
    g = rng.integers(0, 7, size=n).astype(np.int32)
    x = np.round(rng.normal(10, 5, size=n), 2)
    y = rng.integers(-1000, 1000, size=n).astype(np.int32)
    c = rng.choice(np.array([b"ab", b"cd", b"efgh", b"x"], dtype="S4"), size=n)
    xn, yn, cn = rng.random(n) < 0.3, rng.random(n) < 0.1, rng.random(n) < 0.5
    xn[g == 3] = True                                  # one group whose x is always NULL
    return g, x, y, c, xn, yn, cn


def cli(store, sql, timeout=120):
    args = [R.CLI, f"-storage_path={store.rstrip('/')}/", "-num_workers=2"]
    r = subprocess.run(args, input=sql, capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError(r.stderr[-2000:] + r.stdout[-2000:])
    return r.stdout


def main():
    if not R.available():
        sys.exit("oracle/_ref/quickstep_cli_shell not built (oracle/build_ref.sh)")
    g, x, y, c, xn, yn, cn = source_data()
    work = tempfile.mkdtemp(prefix="qs_null_")
    store = os.path.join(work, "store")
    os.makedirs(store)
    try:
        tbl = os.path.join(work, "t.tbl")
        with open(tbl, "w") as f:       # no trailing delimiter: the engine drops such rows when the last column is CHAR
            for i in range(N_ROWS):
                f.write("%d|%s|%s|%s\n" % (g[i], "\\N" if xn[i] else "%.2f" % x[i], "\\N" if yn[i] else str(y[i]),
                                           "\\N" if cn[i] else c[i].decode()))
        ddl = "".join(f"CREATE TABLE {t} (g INT NOT NULL, x DOUBLE NULL, y INT NULL, c CHAR(4) NULL) {props};\n" for t, props in TABLES.items())
        r = subprocess.run([R.CLI, f"-storage_path={store}/", "-num_workers=2", "-initialize_db=true"], input=ddl,
                           capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stderr
        results = {"source": "oracle/_ref/quickstep_cli_shell (UNMODIFIED reference, Release) via tests/golden/make_null_golden.py; cells exactly as printed",
                   "rows": N_ROWS, "queries": QUERIES, "tables": {}}
        arrays = dict(g=g, x=np.where(xn, 0.0, x), y=np.where(yn, 0, y).astype(np.int32), c=np.where(cn, b"", c).astype("S4"),
                      x_null=xn, y_null=yn, c_null=cn)
        for t in TABLES:
            before = {p for p, _m in RB.load_blocks(store)}
            cli(store, f"COPY {t} FROM '{tbl}' WITH (DELIMITER '|');\n")
            new = [(p, m) for p, m in RB.load_blocks(store) if p not in before]
            assert len(new) == 1, (t, len(new))
            arrays["block_" + t] = np.frombuffer(new[0][1], dtype=np.uint8)
            res = {}
            for name, q in QUERIES.items():
                tabs = R.parse_tables(cli(store, q.format(t=t) + "\n"))
                res[name] = tabs[0]
            assert res["lt"][0][0] != "0"
            results["tables"][t] = res
        np.savez_compressed(os.path.join(HERE, "ref_null_blocks.npz"), **arrays)
        with open(os.path.join(HERE, "ref_null_results.json"), "w") as f:
            json.dump(results, f, indent=1)
        print({t: results["tables"][t]["single"] for t in TABLES})
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
