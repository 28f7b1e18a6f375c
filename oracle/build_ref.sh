#!/bin/bash
# TEST INFRASTRUCTURE (oracle side): builds the UNMODIFIED reference engine, quickstep_cli_shell, from
# /root/reference into oracle/_ref/ (git-ignored). Recipe = SURVEY.md §8c / Appendix A. No reference
# source is copied into the repository: the work tree is a scratch copy under /tmp, and the only additions
# are the third-party stand-ins under oracle/ref_shims/ (flags, logging, FRIEND_TEST, regex — none of them
# carries hot-path arithmetic).
#
# -march=native is replaced by x86-64-v3 so the binary also runs on the GPU box's host CPU.
# -Werror is switched off through its cache variable (gcc 13 warns about the vendored protobuf 2.6.1).
# Cost: about 1 h on 8 vCPUs; the six types/operations/comparisons units need about 3.4 GB each.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF=${QS_REFERENCE:-/root/reference}
SRC=${QS_REF_SCRATCH:-/tmp/qs_ref_src}
BLD=${QS_REF_BUILD:-/tmp/qs_ref_build}
JOBS=${JOBS:-6}
OUT="$HERE/_ref"
[ -d "$REF" ] || { echo "no reference tree at $REF"; exit 1; }
mkdir -p "$OUT"
if [ ! -d "$SRC" ]; then
  cp -r "$REF" "$SRC"
  chmod -R u+w "$SRC"
  cp -r "$HERE"/ref_shims/* "$SRC/third_party/src/"
fi
mkdir -p "$BLD"
cd "$BLD"
[ -f CMakeCache.txt ] || cmake -DCMAKE_BUILD_TYPE=Release -DUSE_TCMALLOC=OFF -DUSE_LINENOISE=OFF \
  -DENABLE_HDFS=OFF -DCMAKE_POLICY_VERSION_MINIMUM=3.5 -DGCC_HAS_MARCH_NATIVE=OFF -DCOMPILER_HAS_WERROR=OFF \
  -DCMAKE_CXX_FLAGS="-march=x86-64-v3" "$SRC" > "$OUT/cmake.log" 2>&1
make -j"$JOBS" quickstep_cli_shell > "$OUT/make.log" 2>&1
cp "$BLD/quickstep_cli_shell" "$OUT/quickstep_cli_shell"
echo "built $OUT/quickstep_cli_shell"
