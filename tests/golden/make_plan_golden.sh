#!/bin/bash
# TEST INFRASTRUCTURE: builds tests/golden/make_plan_golden.cpp against the UNMODIFIED reference (the scratch copy and
# build tree of oracle/build_ref.sh) and writes tests/golden/reference_plans.json.  Nothing from the reference is
# copied into the repository; the binary lives under /tmp.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
SRC=${QS_REF_SCRATCH:-/tmp/qs_ref_src}
BLD=${QS_REF_BUILD:-/tmp/qs_ref_build}
OUT=${1:-$HERE/reference_plans.json}
[ -f "$BLD/expressions/Expressions.pb.h" ] || { echo "run oracle/build_ref.sh first"; exit 1; }
BIN=${QS_PLAN_BIN:-$(mktemp -d)/make_plan_golden}
# the reference's own compile definitions and flags (CMakeFiles/quickstep_cli_shell.dir/flags.make)
g++ -std=c++17 -O1 -DNDEBUG -Wno-deprecated-declarations -march=x86-64-v3 \
  -DQUICKSTEP_ENABLE_COMPARISON_INLINE_EXPANSION -DQUICKSTEP_ENABLE_VECTOR_COPY_ELISION_SELECTION \
  -DQUICKSTEP_ENABLE_VECTOR_PREDICATE_SHORT_CIRCUIT -D_ISOC11_SOURCE \
  -I"$SRC" -I"$BLD" -I"$SRC/third_party/src" -I"$BLD/third_party/gflags/include" -I"$SRC/third_party/src/glog/src" \
  -I"$BLD/third_party/glog" -I"$SRC/third_party/src/tmb/include" -isystem "$SRC/third_party/src/protobuf/src" -isystem "$SRC/third_party/src/googletest/googletest/include" \
  -I"$ROOT/include" -I"$ROOT/quickstep_b200/host/intree" \
  "$HERE/make_plan_golden.cpp" -o "$BIN" \
  -Wl,--start-group $(find "$BLD" -name '*.a' | grep -v -e gtest -e benchmark) -Wl,--end-group -lpthread
"$BIN" "${QS_REFERENCE:-/root/reference}" "$OUT"
echo "wrote $OUT"
