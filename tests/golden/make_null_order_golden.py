"""TEST INFRASTRUCTURE, run in the build container: python tests/golden/make_null_order_golden.py

ORDER BY over NULL-able attributes as the UNMODIFIED engine answers it: the relation of make_null_golden.py (same rows,
same seed) loaded into oracle/_ref/quickstep_cli_shell, and the tables it prints for ORDER BY ... LIMIT queries with
the default NULL placement, explicit NULLS FIRST / LAST, descending keys and a second key
-> tests/golden/ref_null_order_results.json.  tests/test_null_oracle.py holds the ordering rule against it."""
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, HERE)
import make_null_golden as G  # noqa: E402
import ref_engine as R  # noqa: E402

QUERIES = {
    "asc_default": "SELECT y, g FROM t WHERE g < 2 ORDER BY y, g LIMIT 40;",
    "desc_default": "SELECT y, g FROM t WHERE g < 2 ORDER BY y DESC, g LIMIT 40;",
    "asc_nulls_first": "SELECT y, g FROM t WHERE g < 2 ORDER BY y NULLS FIRST, g LIMIT 40;",
    "desc_nulls_last": "SELECT y, g FROM t WHERE g < 2 ORDER BY y DESC NULLS LAST, g LIMIT 400;",
    "two_nullable_keys": "SELECT c, y, g FROM t WHERE g = 5 ORDER BY c DESC, y NULLS FIRST, g LIMIT 60;",
    "double_key_desc": "SELECT x, g FROM t WHERE g = 6 ORDER BY x DESC, g LIMIT 30;",
}


def main():
    assert R.available()
    g, x, y, c, xn, yn, cn = G.source_data()
    work = tempfile.mkdtemp(prefix="qs_nullorder_")
    store = os.path.join(work, "store")
    os.makedirs(store)
    try:
        tbl = os.path.join(work, "t.tbl")
        with open(tbl, "w") as f:
            for i in range(G.N_ROWS):
                f.write("%d|%s|%s|%s\n" % (g[i], "\\N" if xn[i] else "%.2f" % x[i], "\\N" if yn[i] else str(y[i]), "\\N" if cn[i] else c[i].decode()))
        ddl = "CREATE TABLE t (g INT NOT NULL, x DOUBLE NULL, y INT NULL, c CHAR(4) NULL) WITH BLOCKPROPERTIES (TYPE split_rowstore, BLOCKSIZEMB 2);\n"
        r = subprocess.run([R.CLI, f"-storage_path={store}/", "-num_workers=2", "-initialize_db=true"], input=ddl, capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stderr
        G.cli(store, f"COPY t FROM '{tbl}' WITH (DELIMITER '|');\n")
        out = {"source": "oracle/_ref/quickstep_cli_shell (UNMODIFIED reference) via tests/golden/make_null_order_golden.py; cells exactly as printed",
               "queries": QUERIES, "results": {}}
        for name, q in QUERIES.items():
            out["results"][name] = R.parse_tables(G.cli(store, q + "\n"))[0]
            print(name, len(out["results"][name]), out["results"][name][:3], "...", out["results"][name][-2:])
        json.dump(out, open(os.path.join(HERE, "ref_null_order_results.json"), "w"), indent=1)
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
