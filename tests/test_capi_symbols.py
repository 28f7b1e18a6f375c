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


def test_struct_layouts_against_the_compiler(tmp_path):
    """gcc compiles include/qsgpu.h as C and prints sizeof / offsetof of every struct that crosses the ABI and the values
    of the enums; the ctypes mirrors in quickstep_b200/capi.py must agree field by field."""
    import subprocess
    structs = {
        "qs_node": ["kind", "op", "type", "width", "a", "b", "lit"],
        "qs_expr_set": ["nodes", "n_nodes", "str_pool", "str_pool_bytes"],
        "qs_attr": ["type", "width"],
        "qs_stage_desc": ["attr", "encoding", "host", "code_width", "stride", "dict", "dict_entries", "null_kind", "null_arg",
                          "null_stride", "null_width", "null_bitmap"],
        "qs_block_image": ["host", "bytes", "n_rows", "descs"],
        "qs_lip_ref": ["lip", "attr"],
        "qs_aggregate": ["function", "argument_root"],
        "qs_agg_spec": ["dev", "strategy", "exprs", "predicate_root", "n_aggregates", "aggregates", "n_group_by", "group_by_roots",
                        "estimated_num_entries", "collision_free_max_key", "nullable_arguments"],
        "qs_sort_key": ["attr", "descending"],
    }
    enums = ["QS_INT", "QS_LONG", "QS_FLOAT", "QS_DOUBLE", "QS_CHAR", "QS_VARCHAR", "QS_DATE", "QS_ENC_PLAIN", "QS_ENC_STRIDED", "QS_ENC_DICT",
             "QS_ENC_TRUNCATED", "QS_ENC_SKIP", "QS_NULL_NONE", "QS_NULL_CODE", "QS_NULL_BITMAP", "QS_NULL_SLOT_WORD", "QS_JOIN_INNER",
             "QS_JOIN_LEFT_SEMI", "QS_JOIN_LEFT_ANTI", "QS_JOIN_LEFT_OUTER", "QS_AGG_AVG", "QS_AGG_COUNT", "QS_AGG_MAX", "QS_AGG_MIN", "QS_AGG_SUM",
             "QS_AGG_SINGLE_STATE", "QS_AGG_COMPACT_KEY", "QS_AGG_SEPARATE_CHAINING", "QS_AGG_COLLISION_FREE", "QS_LIP_BITVECTOR_EXACT",
             "QS_LIP_SINGLE_IDENTITY_HASH", "QSGPU_ERR_NO_DEVICE", "QSGPU_ERR_UNSUPPORTED", "QSGPU_ERR_CAPACITY"]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "qsgpu.h"', 'int main(void) {']
    for s, fields in structs.items():
        lines.append(f'  printf("sizeof {s} %zu\\n", sizeof({s}));')
        for f in fields:
            lines.append(f'  printf("offsetof {s} {f} %zu\\n", offsetof({s}, {f}));')
    for e in enums:
        lines.append(f'  printf("enum {e} %d\\n", (int){e});')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    for line in out.splitlines():
        parts = line.split()
        if parts[0] == "sizeof":
            assert ctypes.sizeof(getattr(A, parts[1])) == int(parts[2]), line
        elif parts[0] == "offsetof":
            assert getattr(getattr(A, parts[1]), parts[2]).offset == int(parts[3]), line
        else:
            assert getattr(A, parts[1]) == int(parts[2]), line
