"""Device-free unit tests of the C++ host layer (quickstep_b200/host/tests/host_unittest.cpp): expression-set
merging and the scheduling contracts (blocking dependencies, streamed inputs fed block by block,
getAllWorkOrders called repeatedly, every work order executed exactly once on the Worker pool)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "quickstep_b200", "lib", "qshost_unittest")


def test_host_unit_binary():
    assert os.path.exists(BIN), "build it: make -C quickstep_b200/host"
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    for case in ("expr_set_append ok", "block_builder ok", "insert_tuples_blocks ok", "blocking_dependency ok", "pipelined_feed ok", "diamond ok", "all host unit tests passed"):
        assert case in r.stdout, r.stdout
