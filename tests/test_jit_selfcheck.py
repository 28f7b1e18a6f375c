"""The query compiler (quickstep_b200/csrc/qs_jit.cu) generates and NVRTC-compiles one
representative kernel per family WITHOUT a device: the CPU-side proof that every kernel
template still builds for sm_100a.  Parity of what the kernels compute is the -m gpu suite."""
import ctypes as C

import pytest

from quickstep_b200 import capi as A

CASES = ["q6_single_state", "q1_compact_key", "select_lip_probe", "build_lip_filter", "join_build",
         "join_probe_inner_residual", "join_probe_anti", "groupby_hash", "groupby_dense", "join_build_dense", "join_probe_dense", "join_probe_left_outer",
         "q6_on_dictionary_codes", "q1_on_dictionary_codes", "join_probe_coded_build_and_probe",
         "nullable_single_state", "nullable_compact_key", "nullable_select", "nullable_join_probe"]


@pytest.mark.parametrize("which", range(A.JIT_SELFCHECK_CASES), ids=CASES)
def test_kernel_family_compiles(which):
    lib = A.load()
    src = C.create_string_buffer(1 << 16)
    log = C.create_string_buffer(1 << 16)
    rc = lib.qsgpu_jit_selfcheck(which, src, len(src), log, len(log))
    assert rc == 0, (lib.qsgpu_last_error() or b"").decode()[:4000]
    text = src.value.decode()
    assert 'extern "C" __global__' in text and "struct Q" in text
    # the program is a compile-time table, not a kernel argument
    assert "QSC Instr code(int pc)" in text


def test_selfcheck_rejects_unknown_case():
    lib = A.load()
    assert lib.qsgpu_jit_selfcheck(A.JIT_SELFCHECK_CASES, None, 0, None, 0) == A.QSGPU_ERR_INVALID
