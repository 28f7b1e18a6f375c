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
