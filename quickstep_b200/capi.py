"""ctypes binding of libqsgpu.so (include/qsgpu.h).

This is harness plumbing for tests/ and bench.py: the product is the C-ABI
library itself (CUDA kernels + C++ host code).  There is no CPU fallback: if
the shared library is missing, or no B200 is visible, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libqsgpu.so")

# ---- ids (values are the reference's own enums, see qsgpu.h) ---------------
QS_INT, QS_LONG, QS_FLOAT, QS_DOUBLE, QS_CHAR, QS_VARCHAR, QS_DATE = range(7)
QS_EQ, QS_NE, QS_LT, QS_LE, QS_GT, QS_GE = range(6)
QS_ADD, QS_SUB, QS_MUL, QS_DIV, QS_MOD = range(5)
QS_NEGATE, QS_CAST = 0, 1
QS_AGG_AVG, QS_AGG_COUNT, QS_AGG_MAX, QS_AGG_MIN, QS_AGG_SUM = range(5)
QS_N_LITERAL, QS_N_ATTRIBUTE, QS_N_UNARY, QS_N_BINARY = 0, 1, 2, 3
QS_N_SHARED = 5
QS_N_TRUE, QS_N_FALSE, QS_N_COMPARISON, QS_N_NEGATION, QS_N_CONJUNCTION, QS_N_DISJUNCTION = 16, 17, 18, 19, 20, 21
QS_ENC_PLAIN, QS_ENC_STRIDED, QS_ENC_DICT, QS_ENC_TRUNCATED, QS_ENC_SKIP = range(5)
QS_LIP_BITVECTOR_EXACT, QS_LIP_SINGLE_IDENTITY_HASH = 0, 1
QS_AGG_SINGLE_STATE, QS_AGG_COMPACT_KEY, QS_AGG_SEPARATE_CHAINING, QS_AGG_COLLISION_FREE = range(4)
QS_NULL_NONE, QS_NULL_CODE, QS_NULL_BITMAP, QS_NULL_SLOT_WORD = range(4)
QS_JOIN_INNER, QS_JOIN_LEFT_SEMI, QS_JOIN_LEFT_ANTI, QS_JOIN_LEFT_OUTER = range(4)
(QS_K_SCAN_AGG, QS_K_SELECT, QS_K_LIP, QS_K_JOIN_BUILD, QS_K_JOIN_PROBE, QS_K_GROUPBY,
 QS_K_PARTITION, QS_K_TOPK, QS_K_STAGE) = range(9)

QSGPU_OK, QSGPU_ERR_NO_DEVICE, QSGPU_ERR_CUDA, QSGPU_ERR_INVALID, QSGPU_ERR_UNSUPPORTED, \
    QSGPU_ERR_CAPACITY, QSGPU_ERR_OOM = range(7)

UINT64_MAX = (1 << 64) - 1


class _Date(C.Structure):
    _fields_ = [("year", C.c_int32), ("month", C.c_uint8), ("day", C.c_uint8), ("pad", C.c_uint8 * 2)]


class _Lit(C.Union):
    _fields_ = [("i32", C.c_int32), ("i64", C.c_int64), ("f32", C.c_float), ("f64", C.c_double),
                ("date", _Date), ("pool_offset", C.c_uint64)]


class qs_node(C.Structure):
    _fields_ = [("kind", C.c_uint16), ("op", C.c_uint16), ("type", C.c_uint16), ("width", C.c_uint16),
                ("a", C.c_int32), ("b", C.c_int32), ("lit", _Lit)]


class qs_expr_set(C.Structure):
    _fields_ = [("nodes", C.POINTER(qs_node)), ("n_nodes", C.c_uint32),
                ("str_pool", C.c_char_p), ("str_pool_bytes", C.c_uint32)]


class qs_attr(C.Structure):
    _fields_ = [("type", C.c_uint16), ("width", C.c_uint16)]


class qs_stage_desc(C.Structure):
    _fields_ = [("attr", C.c_uint32), ("encoding", C.c_uint32), ("host", C.c_void_p),
                ("code_width", C.c_uint32), ("stride", C.c_uint32), ("dict", C.c_void_p),
                ("dict_entries", C.c_uint32), ("null_kind", C.c_uint32), ("null_arg", C.c_uint32),
                ("null_stride", C.c_uint32), ("null_width", C.c_uint32), ("reserved", C.c_uint32),
                ("null_bitmap", C.c_void_p)]


class qs_block_image(C.Structure):
    _fields_ = [("host", C.c_void_p), ("bytes", C.c_uint64), ("n_rows", C.c_uint64),
                ("descs", C.POINTER(qs_stage_desc))]


class qs_ipc_handle(C.Structure):
    _fields_ = [("bytes", C.c_ubyte * 64)]


class qs_comm_id(C.Structure):
    _fields_ = [("bytes", C.c_ubyte * 128)]


class qs_lip_ref(C.Structure):
    _fields_ = [("lip", C.c_void_p), ("attr", C.c_uint32), ("reserved", C.c_uint32)]


class qs_scan(C.Structure):
    _fields_ = [("input", C.c_void_p), ("row_begin", C.c_uint64), ("row_end", C.c_uint64),
                ("exprs", C.POINTER(qs_expr_set)), ("predicate_root", C.c_int32),
                ("n_lip_probe", C.c_uint32), ("lip_probe", C.POINTER(qs_lip_ref))]


class qs_aggregate(C.Structure):
    _fields_ = [("function", C.c_uint32), ("argument_root", C.c_int32)]


class qs_agg_spec(C.Structure):
    _fields_ = [("dev", C.c_int), ("strategy", C.c_uint32), ("exprs", C.POINTER(qs_expr_set)),
                ("predicate_root", C.c_int32), ("n_aggregates", C.c_uint32),
                ("aggregates", C.POINTER(qs_aggregate)), ("n_group_by", C.c_uint32),
                ("group_by_roots", C.POINTER(C.c_int32)), ("estimated_num_entries", C.c_uint64),
                ("collision_free_max_key", C.c_int64), ("nullable_arguments", C.c_uint64)]


class qs_sort_key(C.Structure):
    _fields_ = [("attr", C.c_uint32), ("descending", C.c_uint32)]


# Every symbol include/qsgpu.h declares, with its argument types.
_VP, _VPP = C.c_void_p, C.POINTER(C.c_void_p)
_U64P, _U32P = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)
SIGNATURES = {
    "qsgpu_last_error": (C.c_char_p, []),
    "qsgpu_init": (C.c_int, [C.c_int, C.POINTER(C.c_int)]),
    "qsgpu_shutdown": (C.c_int, []),
    "qsgpu_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "qsgpu_synchronize": (C.c_int, [C.c_int]),
    "qsgpu_stream": (C.c_int, [C.c_int, _VPP]),
    "qsgpu_launch_count": (C.c_int, [_U64P]),
    "qsgpu_malloc": (C.c_int, [C.c_int, C.c_size_t, _VPP]),
    "qsgpu_free": (C.c_int, [C.c_int, _VP]),
    "qsgpu_memcpy_h2d": (C.c_int, [C.c_int, _VP, _VP, C.c_size_t]),
    "qsgpu_memcpy_d2h": (C.c_int, [C.c_int, _VP, _VP, C.c_size_t]),
    "qsgpu_memcpy_d2d": (C.c_int, [C.c_int, _VP, _VP, C.c_size_t]),
    "qsgpu_memcpy_d2d_async": (C.c_int, [C.c_int, _VP, _VP, C.c_size_t]),
    "qsgpu_timer_start": (C.c_int, [C.c_int]),
    "qsgpu_timer_stop": (C.c_int, [C.c_int, C.POINTER(C.c_float)]),
    "qsgpu_host_alloc": (C.c_int, [C.c_size_t, _VPP]),
    "qsgpu_host_free": (C.c_int, [_VP]),
    "qsgpu_relation_create": (C.c_int, [C.c_int, C.c_uint32, C.POINTER(qs_attr), C.c_uint64, _VPP]),
    "qsgpu_relation_destroy": (C.c_int, [_VP]),
    "qsgpu_relation_num_rows": (C.c_int, [_VP, _U64P]),
    "qsgpu_relation_set_num_rows": (C.c_int, [_VP, C.c_uint64]),
    "qsgpu_relation_column": (C.c_int, [_VP, C.c_uint32, _VPP]),
    "qsgpu_relation_wrap": (C.c_int, [C.c_int, C.c_uint32, C.POINTER(qs_attr), _VPP, C.c_uint64, _VPP]),
    "qsgpu_relation_read": (C.c_int, [_VP, C.c_uint32, C.c_uint64, C.c_uint64, _VP]),
    "qsgpu_relation_set_dictionary": (C.c_int, [_VP, C.c_uint32, C.c_uint32, _VP, C.c_uint32]),
    "qsgpu_dictionary_code_range": (C.c_int, [C.c_uint16, C.c_uint16, _VP, C.c_uint32, C.c_uint32, C.POINTER(qs_node), C.c_char_p,
                                              C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_int)]),
    "qsgpu_relation_dictionary": (C.c_int, [_VP, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), _VP]),
    "qsgpu_relation_read_nulls": (C.c_int, [_VP, C.c_uint64, C.c_uint64, _U64P]),
    "qsgpu_relation_set_nullable": (C.c_int, [_VP, C.c_uint64]),
    "qsgpu_relation_nullable": (C.c_int, [_VP, _U64P]),
    "qsgpu_relation_write_nulls": (C.c_int, [_VP, C.c_uint64, C.c_uint64, _U64P]),
    "qsgpu_relation_read_all": (C.c_int, [_VP, C.c_uint64, C.c_uint64, _VPP]),
    "qsgpu_relation_read_rows": (C.c_int, [_VP, C.c_uint64, _VPP, _U64P, _U64P]),
    "qsgpu_stage_block": (C.c_int, [_VP, C.c_uint64, C.POINTER(qs_stage_desc), C.c_uint32]),
    "qsgpu_stage_blocks": (C.c_int, [_VP, C.c_uint32, C.POINTER(qs_block_image), C.c_uint32]),
    "qsgpu_stage_columns": (C.c_int, [_VP, C.c_uint64, C.c_uint32, C.POINTER(qs_block_image), C.c_uint32]),
    "qsgpu_lip_create": (C.c_int, [C.c_int, C.c_uint32, C.c_uint32, C.c_int64, C.c_int64, C.c_uint64, C.c_int, _VPP]),
    "qsgpu_lip_destroy": (C.c_int, [_VP]),
    "qsgpu_lip_num_words": (C.c_int, [_VP, _U64P]),
    "qsgpu_lip_read": (C.c_int, [_VP, _U64P]),
    "qsgpu_lip_device_words": (C.c_int, [_VP, _VPP]),
    "qsgpu_lip_probe_stats": (C.c_int, [_VP, _U64P, _U64P]),
    "qsgpu_build_lip_filter": (C.c_int, [C.POINTER(qs_scan), C.c_uint32, C.POINTER(qs_lip_ref)]),
    "qsgpu_select": (C.c_int, [C.POINTER(qs_scan), C.c_uint32, C.POINTER(C.c_int32), _VP]),
    "qsgpu_agg_create": (C.c_int, [C.POINTER(qs_agg_spec), _VPP]),
    "qsgpu_agg_run": (C.c_int, [_VP, _VP, C.c_uint64, C.c_uint64, C.c_uint32, C.POINTER(qs_lip_ref)]),
    "qsgpu_agg_num_groups": (C.c_int, [_VP, _U64P]),
    "qsgpu_agg_partial": (C.c_int, [_VP, _VPP, _VPP, _U64P, _U32P, _U32P]),
    "qsgpu_agg_partial_layout": (C.c_int, [_VP, _VPP, _VPP, _U64P, _U32P, _U32P]),
    "qsgpu_agg_merge_partial": (C.c_int, [_VP, _VP, _VP, C.c_uint64]),
    "qsgpu_agg_existence_map": (C.c_int, [_VP, _VPP]),
    "qsgpu_agg_finalize": (C.c_int, [_VP, _VPP, _U64P]),
    "qsgpu_agg_destroy": (C.c_int, [_VP]),
    "qsgpu_join_create": (C.c_int, [C.c_int, C.c_uint32, C.c_uint64, _VPP]),
    "qsgpu_join_create_dense": (C.c_int, [C.c_int, C.c_uint32, C.c_int64, C.c_int64, _VPP]),
    "qsgpu_join_build": (C.c_int, [_VP, C.POINTER(qs_scan), C.c_uint32, C.c_uint32, C.POINTER(qs_lip_ref)]),
    "qsgpu_join_num_entries": (C.c_int, [_VP, _U64P]),
    "qsgpu_join_probe": (C.c_int, [_VP, C.POINTER(qs_scan), C.c_uint32, C.c_uint32, C.c_int32, C.c_uint32,
                                   C.POINTER(C.c_int32), _VP]),
    "qsgpu_join_build_composite": (C.c_int, [_VP, C.POINTER(qs_scan), C.c_uint32, _U32P, C.c_uint32, C.POINTER(qs_lip_ref)]),
    "qsgpu_join_probe_composite": (C.c_int, [_VP, C.POINTER(qs_scan), C.c_uint32, _U32P, C.c_uint32, C.c_int32, C.c_uint32,
                                             C.POINTER(C.c_int32), _VP]),
    "qsgpu_join_destroy": (C.c_int, [_VP]),
    "qsgpu_topk": (C.c_int, [_VP, C.c_uint32, C.POINTER(qs_sort_key), C.c_uint64, _VPP]),
    "qsgpu_radix_partition": (C.c_int, [_VP, C.c_uint32, C.c_uint32, _VP, _U64P]),
    "qsgpu_hash_partition": (C.c_int, [_VP, C.c_uint32, C.c_uint32, _VP, _U64P]),
    "qsgpu_range_partition": (C.c_int, [_VP, C.c_uint32, C.c_int64, C.c_uint64, C.c_uint32, _VP, _U64P]),
    "qsgpu_ipc_alloc": (C.c_int, [C.c_int, C.c_size_t, _VPP, C.POINTER(qs_ipc_handle)]),
    "qsgpu_ipc_open": (C.c_int, [C.c_int, C.POINTER(qs_ipc_handle), _VPP]),
    "qsgpu_ipc_close": (C.c_int, [C.c_int, _VP]),
    "qsgpu_ipc_free": (C.c_int, [C.c_int, _VP]),
    "qsgpu_partition_count": (C.c_int, [_VP, C.c_uint32, C.c_uint32, _U64P]),
    "qsgpu_partition_scatter_peers": (C.c_int, [_VP, C.c_uint32, C.c_uint32, _VPP, _U64P]),
    "qsgpu_join_partition": (C.c_int, [_VP, _VP, C.c_uint32, C.c_uint32, _VP, _U64P]),
    "qsgpu_comm_unique_id": (C.c_int, [C.POINTER(qs_comm_id)]),
    "qsgpu_comm_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(qs_comm_id), _VPP]),
    "qsgpu_comm_destroy": (C.c_int, [_VP]),
    "qsgpu_comm_rank": (C.c_int, [_VP, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "qsgpu_comm_peer_memory": (C.c_int, [_VP, C.POINTER(C.c_int)]),
    "qsgpu_comm_barrier": (C.c_int, [_VP]),
    "qsgpu_comm_allreduce_i64": (C.c_int, [_VP, C.POINTER(C.c_int64), C.c_uint32, C.c_uint32]),
    "qsgpu_agg_merge_all": (C.c_int, [_VP, _VP]),
    "qsgpu_lip_allreduce": (C.c_int, [_VP, _VP]),
    "qsgpu_relation_allgather": (C.c_int, [_VP, _VP, _VPP]),
    "qsgpu_relation_allgather_small": (C.c_int, [_VP, _VP, C.c_uint64, _VPP]),
    "qsgpu_set_timing": (C.c_int, [C.c_int]),
    "qsgpu_last_kernel_ms": (C.c_int, [C.c_uint32, C.POINTER(C.c_float)]),
    "qsgpu_kernel_ms_stats": (C.c_int, [C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float), _U32P]),
    "qsgpu_jit_selfcheck": (C.c_int, [C.c_uint32, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]),
    "qsgpu_jit_stats": (C.c_int, [_U64P, _U64P, _U64P]),
}
JIT_SELFCHECK_CASES = 19


class QsGpuError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"qsgpu status {status}: {msg}")
        self.status = status


_lib = None


def load():
    """dlopen libqsgpu.so and bind every declared symbol.  Raises when the
    library is missing -- there is nothing to fall back to."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C quickstep_b200/csrc).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status):
    if status != 0:
        raise QsGpuError(status, (load().qsgpu_last_error() or b"").decode())
    return status
