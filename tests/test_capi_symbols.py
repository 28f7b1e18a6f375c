"""The C-ABI library loads and exports every symbol include/qsgpu.h declares (no compute calls)."""
import ctypes
import os
import re

from quickstep_b200 import capi as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "qsgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qsgpu_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    lib = A.load()
    names = _declared()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"libqsgpu.so does not export {n}"
    assert set(names) == set(A.SIGNATURES), set(names) ^ set(A.SIGNATURES)


def test_struct_layouts_match_header():
    assert ctypes.sizeof(A.qs_node) == 24
    assert A.qs_node.lit.offset == 16
    assert ctypes.sizeof(A.qs_attr) == 4
    assert ctypes.sizeof(A.qs_stage_desc) == 64
    assert ctypes.sizeof(A.qs_lip_ref) == 16
    assert ctypes.sizeof(A.qs_aggregate) == 8


def test_no_cpu_fallback_without_device():
    """Without a CUDA device every compute entry point must fail, never fall back."""
    lib = A.load()
    n = ctypes.c_int(0)
    lib.qsgpu_device_count(ctypes.byref(n))
    if n.value > 0:
        return
    assert lib.qsgpu_init(0, None) == A.QSGPU_ERR_NO_DEVICE
    out = ctypes.c_void_p()
    attrs = (A.qs_attr * 1)()
    attrs[0].type, attrs[0].width = A.QS_INT, 4
    assert lib.qsgpu_relation_create(0, 1, attrs, 16, ctypes.byref(out)) == A.QSGPU_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.qsgpu_last_error() or b"not been called" in lib.qsgpu_last_error()
