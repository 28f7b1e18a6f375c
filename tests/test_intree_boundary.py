"""The drop-in boundary against the REAL reference headers: quickstep_b200/host/intree/GpuWorkOrders.cpp declares GPU work
orders / operators as subclasses of the reference's own quickstep::WorkOrder and quickstep::RelationalOperator and lowers
the reference's own serialized Predicate / Scalar protos into the C ABI's node arrays.  `g++ -fsyntax-only` type-checks
it against /root/reference plus the generated headers (*.pb.h, *Config.h) of oracle/build_ref.sh's build tree.  Skipped
where the reference tree or that build tree is absent (the GPU box)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("QS_REFERENCE", "/root/reference")
BUILD = os.environ.get("QS_REF_BUILD", "/tmp/qs_ref_build")
SHIMS = os.path.join(ROOT, "oracle", "ref_shims")


@pytest.mark.timeout(900)
@pytest.mark.parametrize("source", ["GpuWorkOrders.cpp", "GpuJoinWorkOrders.cpp"])
def test_gpu_work_orders_compile_against_reference_headers(source):
    if not os.path.isdir(REF):
        pytest.skip("reference tree absent")
    if not os.path.exists(os.path.join(BUILD, "expressions", "Expressions.pb.h")):
        pytest.skip("generated headers absent: run oracle/build_ref.sh (builds the unmodified reference under /tmp)")
    inc = [REF, BUILD, os.path.join(REF, "third_party", "src", "protobuf", "src"), os.path.join(REF, "third_party", "src", "tmb", "include"),
           os.path.join(REF, "third_party", "src"),
           os.path.join(BUILD, "third_party"), os.path.join(SHIMS, "gflags", "include"), os.path.join(SHIMS, "glog", "src"),
           os.path.join(SHIMS, "googletest", "googletest", "include"), os.path.join(ROOT, "include"),
           os.path.join(ROOT, "quickstep_b200", "host", "intree")]
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-DNDEBUG", "-Wall", "-Wno-unused-parameter"] + [f"-I{i}" for i in inc] + \
          [os.path.join(ROOT, "quickstep_b200", "host", "intree", source)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-6000:]


def _link_against_reference(source, exe, with_qsgpu):
    """g++ <source> + the static libraries of the unmodified reference's build tree (+ libqsgpu.so) -> exe; skips when the
    scratch copy / build tree of oracle/build_ref.sh is absent."""
    src = os.environ.get("QS_REF_SCRATCH", "/tmp/qs_ref_src")
    lib = os.path.join(ROOT, "quickstep_b200", "lib")
    if not os.path.isdir(REF) or not os.path.isdir(src) or not os.path.exists(os.path.join(BUILD, "expressions", "Expressions.pb.h")):
        pytest.skip("reference tree / build tree absent: run oracle/build_ref.sh")
    if with_qsgpu and not os.path.exists(os.path.join(lib, "libqsgpu.so")):
        pytest.skip("libqsgpu.so not built")
    archives = subprocess.run(["find", BUILD, "-name", "*.a"], capture_output=True, text=True).stdout.split()
    archives = [a for a in archives if "gtest" not in a and "benchmark" not in a]
    inc = [src, BUILD, os.path.join(src, "third_party", "src"), os.path.join(BUILD, "third_party", "gflags", "include"),
           os.path.join(src, "third_party", "src", "glog", "src"), os.path.join(BUILD, "third_party", "glog"),
           os.path.join(src, "third_party", "src", "tmb", "include"), os.path.join(ROOT, "include"),
           os.path.join(ROOT, "quickstep_b200", "host", "intree")]
    sysinc = [os.path.join(src, "third_party", "src", "protobuf", "src"), os.path.join(src, "third_party", "src", "googletest", "googletest", "include")]
    # the reference's own compile definitions (CMakeFiles/quickstep_cli_shell.dir/flags.make): they select inline code paths
    defs = ["-DNDEBUG", "-DQUICKSTEP_ENABLE_COMPARISON_INLINE_EXPANSION", "-DQUICKSTEP_ENABLE_VECTOR_COPY_ELISION_SELECTION",
            "-DQUICKSTEP_ENABLE_VECTOR_PREDICATE_SHORT_CIRCUIT", "-D_ISOC11_SOURCE"]
    cmd = ["g++", "-std=c++17", "-O0", "-march=x86-64-v3", "-Wno-deprecated-declarations"] + defs + [f"-I{i}" for i in inc] + \
          [x for i in sysinc for x in ("-isystem", i)] + [source, "-o", exe, "-Wl,--start-group"] + archives + ["-Wl,--end-group"] + \
          ([f"-L{lib}", "-lqsgpu", f"-Wl,-rpath,{lib}", "-Wl,--allow-shlib-undefined"] if with_qsgpu else []) + ["-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-6000:]


@pytest.mark.timeout(900)
def test_in_tree_operators_link_and_keep_the_scheduling_contract(tmp_path):
    """tests/intree/contract_test.cpp: both in-tree files compiled for real, linked with the reference's static libraries
    (scratch copy + build tree of oracle/build_ref.sh) and with libqsgpu.so, then run: the GPU operators' getAllWorkOrders /
    feedInputBlock / doneFeedingInputBlocks against the reference's real WorkOrdersContainer and a real QueryContext holding
    the predicates the reference's optimizer produced for TPC-H Q3."""
    exe = str(tmp_path / "contract_test")
    _link_against_reference(os.path.join(ROOT, "tests", "intree", "contract_test.cpp"), exe, with_qsgpu=True)
    needed = subprocess.run(["nm", "-D", "--undefined-only", exe], capture_output=True, text=True).stdout
    for sym in ("qsgpu_select", "qsgpu_agg_run", "qsgpu_agg_create", "qsgpu_agg_finalize", "qsgpu_join_build", "qsgpu_join_probe",
                "qsgpu_build_lip_filter", "qsgpu_lip_create", "qsgpu_relation_read_rows"):
        assert sym in needed, sym            # the binding really calls the C ABI, and the library resolves it
    r = subprocess.run([exe, REF], capture_output=True, text=True)
    assert r.returncode == 0 and "in-tree contract ok" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])


@pytest.mark.timeout(900)
def test_host_layer_result_blocks_open_with_the_reference_storage_block(tmp_path):
    """Rows that leave the device (InsertDestination): quickstep_b200/lib/qshost_unittest writes the
    2 MB SplitRowStore block images StorageManager::insertTuples produces for 100,000 result tuples; tests/intree/
    read_blocks.cpp, linked with the reference, opens them with the REAL quickstep::StorageBlock (header parsed and
    validated, MalformedBlock otherwise) and reads every tuple back through the real sub-block's accessors."""
    unit = os.path.join(ROOT, "quickstep_b200", "lib", "qshost_unittest")
    if not os.path.exists(unit):
        pytest.skip("qshost_unittest not built: make -C quickstep_b200/host")
    exe = str(tmp_path / "read_blocks")
    _link_against_reference(os.path.join(ROOT, "tests", "intree", "read_blocks.cpp"), exe, with_qsgpu=False)
    blocks = tmp_path / "blocks"
    blocks.mkdir()
    r = subprocess.run([unit, str(blocks)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "insert_tuples_blocks ok" in r.stdout, r.stdout + r.stderr
    assert sorted(p.name for p in blocks.iterdir()) == ["block_0.bin", "block_1.bin"]
    r = subprocess.run([exe, str(blocks)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "reference StorageBlock read 100000 tuples" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
