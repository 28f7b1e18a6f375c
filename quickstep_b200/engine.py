"""Thin Python handles over the C-ABI objects of libqsgpu.so.

Harness plumbing only (tests/, bench.py, __graft_entry__): every method is one
C-ABI call; results are read back as numpy arrays.  All compute happens in the
CUDA kernels behind the ABI -- a missing library or GPU raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi as A
from .table import Column, HostTable, np_dtype

_inited = False


def init(devices=None):
    """qsgpu_init; devices=None -> device 0 only (one process per GPU)."""
    global _inited
    L = A.load()
    if _inited:
        return L
    devs = [0] if devices is None else list(devices)
    arr = (C.c_int * len(devs))(*devs)
    A.check(L.qsgpu_init(len(devs), arr))
    _inited = True
    return L


def shutdown():
    global _inited
    if _inited:
        A.load().qsgpu_shutdown()
        _inited = False


def stream_ptr(dev=0) -> int:
    """cudaStream_t of the library's stream on `dev` (for torch.cuda.ExternalStream)."""
    p = C.c_void_p()
    A.check(A.load().qsgpu_stream(dev, C.byref(p)))
    return p.value


def launch_count() -> int:
    n = C.c_uint64(0)
    A.check(A.load().qsgpu_launch_count(C.byref(n)))
    return n.value


def synchronize(dev=0):
    A.check(A.load().qsgpu_synchronize(dev))


def timer_start(dev=0):
    A.check(A.load().qsgpu_timer_start(dev))


def timer_stop(dev=0) -> float:
    ms = C.c_float(0)
    A.check(A.load().qsgpu_timer_stop(dev, C.byref(ms)))
    return ms.value


def memcpy_d2d(dst_ptr: int, src_ptr: int, nbytes: int, dev=0):
    A.check(A.load().qsgpu_memcpy_d2d(dev, dst_ptr, src_ptr, nbytes))


def memcpy_d2d_async(dst_ptr: int, src_ptr: int, nbytes: int, dev=0):
    A.check(A.load().qsgpu_memcpy_d2d_async(dev, dst_ptr, src_ptr, nbytes))


def set_timing(on: bool):
    A.check(A.load().qsgpu_set_timing(1 if on else 0))


def kernel_ms_stats(family: int) -> dict:
    last, mx, sm, n = C.c_float(0), C.c_float(0), C.c_float(0), C.c_uint32(0)
    A.check(A.load().qsgpu_kernel_ms_stats(family, C.byref(last), C.byref(mx), C.byref(sm), C.byref(n)))
    return dict(last=last.value, max=mx.value, sum=sm.value, count=n.value)


def last_kernel_ms(family: int) -> float:
    ms = C.c_float(0)
    A.check(A.load().qsgpu_last_kernel_ms(family, C.byref(ms)))
    return ms.value


def _attrs(schema):
    arr = (A.qs_attr * len(schema))()
    for i, (t, w) in enumerate(schema):
        arr[i].type, arr[i].width = t, w
    return arr


class Relation:
    """qsgpu_relation_t: device-resident columns of one relation."""

    def __init__(self, handle, schema, names=None, dev=0, owner=True, keep=None):
        self.h = C.c_void_p(handle) if not isinstance(handle, C.c_void_p) else handle
        self.schema = list(schema)         # [(type, width)]
        self.names = list(names) if names else [f"c{i}" for i in range(len(schema))]
        self.dev = dev
        self.owner = owner
        self._keep = keep

    # -- construction ---------------------------------------------------
    @classmethod
    def create(cls, schema, capacity, names=None, dev=0):
        L = init()
        out = C.c_void_p()
        A.check(L.qsgpu_relation_create(dev, len(schema), _attrs(schema), capacity, C.byref(out)))
        return cls(out, schema, names, dev)

    @classmethod
    def from_host(cls, table: HostTable, dev=0, block_rows=None):
        """Stage a host table block by block (QS_ENC_PLAIN stripes)."""
        schema = [(c.type, c.width) for c in table.columns]
        rel = cls.create(schema, max(table.n_rows, 1), [c.name for c in table.columns], dev)
        step = block_rows or max(table.n_rows, 1)
        for lo in range(0, table.n_rows, step):
            hi = min(table.n_rows, lo + step)
            rel.stage_plain([c.data[lo:hi] for c in table.columns])
        return rel

    @classmethod
    def wrap(cls, schema, device_ptrs, n_rows, names=None, dev=0, keep=None):
        L = init()
        out = C.c_void_p()
        ptrs = (C.c_void_p * len(device_ptrs))(*device_ptrs)
        A.check(L.qsgpu_relation_wrap(dev, len(schema), _attrs(schema), ptrs, n_rows, C.byref(out)))
        return cls(out, schema, names, dev, keep=keep)

    def set_nullable(self, attrs):
        """qsgpu_relation_set_nullable: `attrs` (indices) may hold NULLs."""
        mask = 0
        for a in attrs:
            mask |= 1 << a
        A.check(A.load().qsgpu_relation_set_nullable(self.h, mask))

    def write_nulls(self, masks: np.ndarray, lo=0):
        """qsgpu_relation_write_nulls: per-row NULL masks (bit a = attribute a is NULL) of rows [lo, lo+len)."""
        m = np.ascontiguousarray(masks, dtype=np.uint64)
        A.check(A.load().qsgpu_relation_write_nulls(self.h, lo, len(m), m.ctypes.data_as(C.POINTER(C.c_uint64))))

    def set_dictionary(self, attr: int, code_width: int, dict_values: np.ndarray):
        """qsgpu_relation_set_dictionary: attribute `attr` is resident as `code_width`-byte codes into the sorted
        relation-wide dictionary `dict_values` (native values)."""
        t, w = self.schema[attr]
        d = np.ascontiguousarray(dict_values)
        assert d.dtype.itemsize == np_dtype(t, w).itemsize
        A.check(A.load().qsgpu_relation_set_dictionary(self.h, attr, code_width, d.ctypes.data, len(d)))

    def dictionary(self, attr: int):
        """-> (code_width, dictionary values) ; code_width 0 = native attribute."""
        cw, n = C.c_uint32(0), C.c_uint32(0)
        A.check(A.load().qsgpu_relation_dictionary(self.h, attr, C.byref(cw), C.byref(n), None))
        t, w = self.schema[attr]
        out = np.zeros(max(n.value, 1), dtype=np_dtype(t, w))
        if cw.value:
            A.check(A.load().qsgpu_relation_dictionary(self.h, attr, None, None, out.ctypes.data))
        return cw.value, out[: n.value]

    @classmethod
    def from_host_coded(cls, table: HostTable, coded: dict, dev=0, block_rows=None):
        """Stage a host table as compressed-column-store style blocks of `block_rows` tuples: attributes in
        `coded` ({attr index: code width}) are declared dictionary-coded relation-wide and arrive as
        QS_ENC_DICT stripes with PER-BLOCK dictionaries (re-coded by qsgpu_stage_blocks); the others as plain
        stripes.  Test / bench harness helper: the host side only builds the block images."""
        schema = [(c.type, c.width) for c in table.columns]
        rel = cls.create(schema, max(table.n_rows, 1), [c.name for c in table.columns], dev)
        for a, cw in coded.items():
            rel.set_dictionary(a, cw, np.unique(table.columns[a].data))
        step = block_rows or max(table.n_rows, 1)
        images = []
        for lo in range(0, table.n_rows, step):
            hi = min(table.n_rows, lo + step)
            parts, descs, off = [], [], 0

            def put(arr):
                nonlocal off
                b = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
                pad = (-len(b)) % 16
                start = off
                parts.append(b)
                if pad:
                    parts.append(np.zeros(pad, dtype=np.uint8))
                off += len(b) + pad
                return start

            for a, c in enumerate(table.columns):
                data = c.data[lo:hi]
                if a in coded:
                    d, inv = np.unique(data, return_inverse=True)
                    bcw = 1 if len(d) <= 256 else 2 if len(d) <= 65536 else 4
                    doff = put(d)
                    coff = put(inv.astype({1: np.uint8, 2: np.uint16, 4: np.uint32}[bcw]))
                    descs.append(dict(attr=a, encoding=A.QS_ENC_DICT, offset=coff, code_width=bcw, dict_offset=doff,
                                      dict_entries=len(d)))
                else:
                    descs.append(dict(attr=a, encoding=A.QS_ENC_PLAIN, offset=put(data)))
            images.append((np.concatenate(parts) if parts else np.zeros(16, np.uint8), hi - lo, descs))
        if images:
            rel.stage_blocks(images)
        return rel

    def stage_plain(self, arrays):
        descs = (A.qs_stage_desc * len(arrays))()
        keep = []
        n = len(arrays[0])
        for i, a in enumerate(arrays):
            a = np.ascontiguousarray(a)
            keep.append(a)
            descs[i].attr, descs[i].encoding, descs[i].host = i, A.QS_ENC_PLAIN, a.ctypes.data
        A.check(A.load().qsgpu_stage_block(self.h, n, descs, len(arrays)))

    def stage(self, n_rows, descs_py):
        """descs_py: list of dicts(attr, encoding, host(np), code_width, stride, dict(np))."""
        descs = (A.qs_stage_desc * len(descs_py))()
        for i, d in enumerate(descs_py):
            descs[i].attr = d["attr"]
            descs[i].encoding = d["encoding"]
            descs[i].host = d["host"].ctypes.data
            descs[i].code_width = d.get("code_width", 0)
            descs[i].stride = d.get("stride", 0)
            if d.get("dict") is not None:
                descs[i].dict = d["dict"].ctypes.data
                descs[i].dict_entries = len(d["dict"])
        A.check(A.load().qsgpu_stage_block(self.h, n_rows, descs, len(descs_py)))

    def stage_blocks(self, images):
        """qsgpu_stage_blocks.  images: list of (memory: np.uint8 array, n_rows, descs) where descs is a
        list of dicts(attr, encoding, offset, code_width, stride, dict_offset, dict_entries): byte
        offsets of the stripe / dictionary inside `memory`."""
        n_desc = len(self.schema)
        imgs = (A.qs_block_image * len(images))()
        keep = []
        for b, (mem, n_rows, descs_py) in enumerate(images):
            assert len(descs_py) == n_desc
            descs = (A.qs_stage_desc * n_desc)()
            base = mem.ctypes.data
            for i, d in enumerate(descs_py):
                descs[i].attr, descs[i].encoding = d["attr"], d["encoding"]
                descs[i].host = base + d["offset"]
                descs[i].code_width = d.get("code_width", 0)
                descs[i].stride = d.get("stride", 0)
                if d.get("dict_offset") is not None:
                    descs[i].dict = base + d["dict_offset"]
                    descs[i].dict_entries = d["dict_entries"]
                if d.get("null_kind"):
                    descs[i].null_kind = d["null_kind"]
                    descs[i].null_arg = d.get("null_arg", 0)
                    descs[i].null_stride = d.get("null_stride", 0)
                    descs[i].null_width = d.get("null_width", 0)
                    if d.get("null_offset") is not None:
                        descs[i].null_bitmap = base + d["null_offset"]
            keep.append(descs)
            imgs[b].host, imgs[b].bytes, imgs[b].n_rows, imgs[b].descs = base, mem.nbytes, n_rows, descs
        A.check(A.load().qsgpu_stage_blocks(self.h, len(images), imgs, n_desc))

    # -- access -----------------------------------------------------------
    @property
    def n_rows(self) -> int:
        n = C.c_uint64(0)
        A.check(A.load().qsgpu_relation_num_rows(self.h, C.byref(n)))
        return n.value

    def column_ptr(self, attr: int) -> int:
        p = C.c_void_p()
        A.check(A.load().qsgpu_relation_column(self.h, attr, C.byref(p)))
        return p.value

    def read(self, attr: int, lo=0, n=None) -> np.ndarray:
        t, w = self.schema[attr]
        if n is None:
            n = self.n_rows - lo
        out = np.zeros(max(n, 1), dtype=np_dtype(t, w))
        if n:
            A.check(A.load().qsgpu_relation_read(self.h, attr, lo, n, out.ctypes.data))
        return out[:n]

    def read_nulls(self, lo=0, n=None) -> np.ndarray:
        """NULL mask per row (bit j = column j is NULL); all zeros unless a LEFT OUTER join wrote the relation."""
        if n is None:
            n = self.n_rows - lo
        out = np.zeros(max(n, 1), dtype=np.uint64)
        if n:
            A.check(A.load().qsgpu_relation_read_nulls(self.h, lo, n, out.ctypes.data_as(C.POINTER(C.c_uint64))))
        return out[:n]

    def read_all(self, lo=0, n=None):
        """Every column of rows [lo, lo+n) with one synchronisation (qsgpu_relation_read_all)."""
        if n is None:
            n = self.n_rows - lo
        outs = [np.zeros(max(n, 1), dtype=np_dtype(t, w)) for (t, w) in self.schema]
        if n:
            ptrs = (C.c_void_p * len(outs))(*[o.ctypes.data for o in outs])
            A.check(A.load().qsgpu_relation_read_all(self.h, lo, n, ptrs))
        return [o[:n] for o in outs]

    def to_host(self, name=None) -> HostTable:
        n = self.n_rows
        return HostTable(name or "rel", [Column(self.names[i], t, self.read(i, 0, n), w)
                                         for i, (t, w) in enumerate(self.schema)])

    def attr(self, es, i: int, side: int = 0) -> int:
        t, w = self.schema[i]
        return es.attr(i, t, w, side)

    def destroy(self):
        if self.h and self.owner:
            A.load().qsgpu_relation_destroy(self.h)
        self.h = None


class LipFilter:
    def __init__(self, kind, attr_type, min_value=0, max_value=0, cardinality=0, is_anti=False, dev=0):
        L = init()
        self.h = C.c_void_p()
        A.check(L.qsgpu_lip_create(dev, kind, attr_type, min_value, max_value, cardinality, 1 if is_anti else 0,
                                   C.byref(self.h)))

    def words(self) -> np.ndarray:
        n = C.c_uint64(0)
        A.check(A.load().qsgpu_lip_num_words(self.h, C.byref(n)))
        out = np.zeros(max(1, n.value), dtype=np.uint64)
        A.check(A.load().qsgpu_lip_read(self.h, out.ctypes.data_as(C.POINTER(C.c_uint64))))
        return out[: n.value]

    def probe_stats(self):
        """(rows that probed the filter, rows it rejected) over every scan so far."""
        p, m = C.c_uint64(0), C.c_uint64(0)
        A.check(A.load().qsgpu_lip_probe_stats(self.h, C.byref(p), C.byref(m)))
        return p.value, m.value

    def device_words(self):
        p = C.c_void_p()
        A.check(A.load().qsgpu_lip_device_words(self.h, C.byref(p)))
        n = C.c_uint64(0)
        A.check(A.load().qsgpu_lip_num_words(self.h, C.byref(n)))
        return p.value, n.value

    def destroy(self):
        if self.h:
            A.load().qsgpu_lip_destroy(self.h)
        self.h = None


def _lip_refs(refs):
    """refs: list of (LipFilter, attr)."""
    if not refs:
        return 0, None
    arr = (A.qs_lip_ref * len(refs))()
    for i, (f, attr) in enumerate(refs):
        arr[i].lip = f.h
        arr[i].attr = attr
    return len(refs), arr


def _i32(v):
    return (C.c_int32 * max(1, len(v)))(*v)


def _scan(rel: Relation, es, pred_root, lip_probe, row_begin=0, row_end=A.UINT64_MAX):
    s = A.qs_scan()
    s.input = rel.h
    s.row_begin, s.row_end = row_begin, row_end
    s.exprs = es.ptr() if es is not None else None
    s.predicate_root = pred_root
    n, arr = _lip_refs(lip_probe)
    s.n_lip_probe = n
    s.lip_probe = arr
    return s, arr


def build_lip_filter(rel, es, pred_root, lip_probe, lip_build, row_begin=0, row_end=A.UINT64_MAX):
    """BuildLIPFilterWorkOrder::execute."""
    s, _k = _scan(rel, es, pred_root, lip_probe, row_begin, row_end)
    n, arr = _lip_refs(lip_build)
    A.check(A.load().qsgpu_build_lip_filter(C.byref(s), n, arr))


def select(rel, es, pred_root, lip_probe, project_roots, output: Relation, row_begin=0, row_end=A.UINT64_MAX):
    """SelectWorkOrder::execute."""
    s, _k = _scan(rel, es, pred_root, lip_probe, row_begin, row_end)
    A.check(A.load().qsgpu_select(C.byref(s), len(project_roots), _i32(project_roots), output.h))


class AggState:
    """qsgpu_agg_state_t (AggregationOperationState)."""

    def __init__(self, strategy, es, pred_root, aggregates, group_by_roots, estimated=1024, max_key=-1, dev=0,
                 nullable_args=()):
        L = init()
        self.es = es
        self.aggs = (A.qs_aggregate * max(1, len(aggregates)))()
        for i, (f, r) in enumerate(aggregates):
            self.aggs[i].function, self.aggs[i].argument_root = f, r
        self.groups = _i32(group_by_roots)
        spec = A.qs_agg_spec()
        spec.dev, spec.strategy = dev, strategy
        spec.exprs = es.ptr()
        spec.predicate_root = pred_root
        spec.n_aggregates, spec.aggregates = len(aggregates), self.aggs
        spec.n_group_by, spec.group_by_roots = len(group_by_roots), self.groups
        spec.estimated_num_entries = estimated
        spec.collision_free_max_key = max_key
        for j in nullable_args:            # aggregates whose argument has a NULL-able type
            spec.nullable_arguments |= 1 << j
        self.h = C.c_void_p()
        self.n_aggregates = len(aggregates)
        self.n_group_by = len(group_by_roots)
        self.dev = dev
        A.check(L.qsgpu_agg_create(C.byref(spec), C.byref(self.h)))

    def run(self, rel: Relation, row_begin=0, row_end=A.UINT64_MAX, lip_probe=None):
        """AggregationWorkOrder::execute."""
        n, arr = _lip_refs(lip_probe)
        A.check(A.load().qsgpu_agg_run(self.h, rel.h, row_begin, row_end, n, arr))

    def num_groups(self) -> int:
        n = C.c_uint64(0)
        A.check(A.load().qsgpu_agg_num_groups(self.h, C.byref(n)))
        return n.value

    def partial(self):
        """-> (d_states, d_keys, n_groups, words_per_group, key_words)."""
        ds, dk = C.c_void_p(), C.c_void_p()
        n, w, kw = C.c_uint64(0), C.c_uint32(0), C.c_uint32(0)
        A.check(A.load().qsgpu_agg_partial(self.h, C.byref(ds), C.byref(dk), C.byref(n), C.byref(w), C.byref(kw)))
        return ds.value, dk.value, n.value, w.value, kw.value

    def existence_map(self) -> "LipFilter":
        """The COLLISION_FREE table's existence map as a (borrowed) exact LIP filter: build it with
        build_lip_filter(..., [(state.existence_map(), build_attr)]) = BuildAggregationExistenceMapWorkOrder."""
        h = C.c_void_p()
        A.check(A.load().qsgpu_agg_existence_map(self.h, C.byref(h)))
        f = LipFilter.__new__(LipFilter)
        f.h = h
        f.destroy = lambda: None          # owned by the state
        return f

    def partial_layout(self):
        """-> (d_states, d_keys, rows, words_per_group, key_words) without synchronising."""
        ds, dk = C.c_void_p(), C.c_void_p()
        n, w, kw = C.c_uint64(0), C.c_uint32(0), C.c_uint32(0)
        A.check(A.load().qsgpu_agg_partial_layout(self.h, C.byref(ds), C.byref(dk), C.byref(n), C.byref(w), C.byref(kw)))
        return ds.value, dk.value, n.value, w.value, kw.value

    def merge_partial(self, d_states, d_keys, n_groups):
        A.check(A.load().qsgpu_agg_merge_partial(self.h, d_states, d_keys, n_groups))

    def finalize(self, key_schema, key_names=None):
        """FinalizeAggregationWorkOrder::execute -> (Relation, null_mask).
        key_schema: [(type,width)] of the group-by attributes."""
        out = C.c_void_p()
        mask = C.c_uint64(0)
        A.check(A.load().qsgpu_agg_finalize(self.h, C.byref(out), C.byref(mask)))
        schema = list(key_schema)
        # aggregate output types are defined by the library; query them from the handle layout
        return out, mask.value

    def destroy(self):
        if self.h:
            A.load().qsgpu_agg_destroy(self.h)
        self.h = None


def finalize_relation(state: AggState, key_schema, agg_out_types, names=None) -> tuple[Relation, int]:
    """Wraps qsgpu_agg_finalize's output relation; agg_out_types: [(type,width)] per aggregate
    (SUM(int)->LONG, SUM(fp)->DOUBLE, AVG->DOUBLE, COUNT->LONG, MIN/MAX->argument type)."""
    h, mask = state.finalize(key_schema)
    schema = list(key_schema) + list(agg_out_types)
    return Relation(h, schema, names, state.dev), mask


class JoinTable:
    def __init__(self, key_type, estimated, dev=0, dense_range=None):
        """dense_range=(min_key, max_key): a dense (collision-free vector style) table over that key range."""
        L = init()
        self.h = C.c_void_p()
        if dense_range is not None:
            A.check(L.qsgpu_join_create_dense(dev, key_type, dense_range[0], dense_range[1], C.byref(self.h)))
        else:
            A.check(L.qsgpu_join_create(dev, key_type, estimated, C.byref(self.h)))

    def build(self, rel, es, pred_root, key_attr, lip_probe=None, lip_build=None, row_begin=0,
              row_end=A.UINT64_MAX):
        """BuildHashWorkOrder::execute."""
        s, _k = _scan(rel, es, pred_root, lip_probe, row_begin, row_end)
        n, arr = _lip_refs(lip_build)
        keys = list(key_attr) if isinstance(key_attr, (list, tuple)) else [key_attr]      # 2 INT attrs: composite key
        A.check(A.load().qsgpu_join_build_composite(self.h, C.byref(s), len(keys), (C.c_uint32 * len(keys))(*keys), n, arr))

    def num_entries(self) -> int:
        n = C.c_uint64(0)
        A.check(A.load().qsgpu_join_num_entries(self.h, C.byref(n)))
        return n.value

    def probe(self, rel, es, pred_root, key_attr, join_type, residual_root, project_roots, output: Relation,
              lip_probe=None, row_begin=0, row_end=A.UINT64_MAX):
        """Hash{Inner,Semi,Anti}JoinWorkOrder::execute."""
        s, _k = _scan(rel, es, pred_root, lip_probe, row_begin, row_end)
        keys = list(key_attr) if isinstance(key_attr, (list, tuple)) else [key_attr]
        A.check(A.load().qsgpu_join_probe_composite(self.h, C.byref(s), len(keys), (C.c_uint32 * len(keys))(*keys),
                                                    join_type, residual_root, len(project_roots), _i32(project_roots),
                                                    output.h))

    def partition(self, rel, key_attr, n_parts, output: "Relation") -> np.ndarray:
        """Group rows by the slice of this table their key lands in (qsgpu_join_partition) -> offsets."""
        offs = np.zeros(n_parts + 1, dtype=np.uint64)
        A.check(A.load().qsgpu_join_partition(self.h, rel.h, key_attr, n_parts, output.h,
                                              offs.ctypes.data_as(C.POINTER(C.c_uint64))))
        return offs

    def destroy(self):
        if self.h:
            A.load().qsgpu_join_destroy(self.h)
        self.h = None


def topk(rel: Relation, keys, limit) -> Relation:
    """keys: [(attr, descending)] or [(attr, descending, nulls_first)] (nulls_first None = the reference's default:
    NULLs first iff descending)."""
    ks = (A.qs_sort_key * max(1, len(keys)))()
    for i, key in enumerate(keys):
        a, d = key[0], key[1]
        nf = key[2] if len(key) > 2 else None
        ks[i].attr = a
        ks[i].descending = (1 if d else 0) | (0 if nf is None else 2 if nf else 4)
    out = C.c_void_p()
    A.check(A.load().qsgpu_topk(rel.h, len(keys), ks, limit, C.byref(out)))
    return Relation(out, rel.schema, rel.names, rel.dev)


def ipc_alloc(nbytes: int, dev=0):
    """Device memory other processes can map -> (device pointer, 64-byte handle as bytes)."""
    p, h = C.c_void_p(), A.qs_ipc_handle()
    A.check(A.load().qsgpu_ipc_alloc(dev, nbytes, C.byref(p), C.byref(h)))
    return p.value, bytes(h.bytes)


def ipc_open(handle: bytes, dev=0) -> int:
    h = A.qs_ipc_handle()
    C.memmove(C.byref(h), handle, 64)
    p = C.c_void_p()
    A.check(A.load().qsgpu_ipc_open(dev, C.byref(h), C.byref(p)))
    return p.value


def ipc_close(ptr: int, dev=0):
    A.check(A.load().qsgpu_ipc_close(dev, ptr))


def ipc_free(ptr: int, dev=0):
    A.check(A.load().qsgpu_ipc_free(dev, ptr))


def partition_count(rel: Relation, key_attr, n_parts) -> np.ndarray:
    counts = np.zeros(n_parts, dtype=np.uint64)
    A.check(A.load().qsgpu_partition_count(rel.h, key_attr, n_parts, counts.ctypes.data_as(C.POINTER(C.c_uint64))))
    return counts


def partition_scatter_peers(rel: Relation, key_attr, n_parts, peer_cols, first_rows):
    """peer_cols: [n_parts][n_attrs] device addresses (own or IPC-mapped); first_rows: [n_parts]."""
    flat = [p for row in peer_cols for p in row]
    ptrs = (C.c_void_p * len(flat))(*flat)
    fr = np.ascontiguousarray(first_rows, dtype=np.uint64)
    A.check(A.load().qsgpu_partition_scatter_peers(rel.h, key_attr, n_parts, ptrs, fr.ctypes.data_as(C.POINTER(C.c_uint64))))


def range_partition(rel: Relation, key_attr, min_key, part_width, n_parts, output: Relation) -> np.ndarray:
    offs = np.zeros(n_parts + 1, dtype=np.uint64)
    A.check(A.load().qsgpu_range_partition(rel.h, key_attr, min_key, part_width, n_parts, output.h,
                                           offs.ctypes.data_as(C.POINTER(C.c_uint64))))
    return offs


def hash_partition(rel: Relation, key_attr, n_parts, output: Relation) -> np.ndarray:
    """HashPartitionSchemeHeader's partition function (value & (n - 1) / value % n) -> partition start offsets."""
    offs = (C.c_uint64 * (n_parts + 1))()
    A.check(A.load().qsgpu_hash_partition(rel.h, key_attr, n_parts, output.h, offs))
    return np.array(list(offs), dtype=np.uint64)


def radix_partition(rel: Relation, key_attr, n_parts, output: Relation) -> np.ndarray:
    offs = np.zeros(n_parts + 1, dtype=np.uint64)
    A.check(A.load().qsgpu_radix_partition(rel.h, key_attr, n_parts, output.h,
                                           offs.ctypes.data_as(C.POINTER(C.c_uint64))))
    return offs


# ------------------------------------------------------------------------------------------ multi-GPU (NCCL in C)
class Comm:
    """qsgpu_comm_t: the communicator behind qsgpu_agg_merge_all / qsgpu_lip_allreduce / qsgpu_relation_allgather.
    The data-path collectives live in the C ABI; the harness only carries the 128-byte id from rank 0 to the others."""

    def __init__(self, dev: int, rank: int, world: int, id_bytes: bytes):
        cid = A.qs_comm_id()
        C.memmove(cid.bytes, id_bytes, 128)
        self.h = C.c_void_p()
        self.rank, self.world = rank, world
        A.check(A.load().qsgpu_comm_create(dev, rank, world, C.byref(cid), C.byref(self.h)))

    @staticmethod
    def unique_id() -> bytes:
        cid = A.qs_comm_id()
        A.check(A.load().qsgpu_comm_unique_id(C.byref(cid)))
        return bytes(cid.bytes)

    @classmethod
    def from_torch_distributed(cls, dev: int):
        """One communicator per rank of the initialised torch.distributed group (plumbing: broadcasts the id)."""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        on_gpu = dist.get_backend() == "nccl"
        t = torch.zeros(128, dtype=torch.uint8, device=torch.device("cuda", dev) if on_gpu else "cpu")
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(cls.unique_id()), dtype=torch.uint8))
        dist.broadcast(t, 0)
        return cls(dev, rank, world, bytes(t.cpu().numpy().tobytes()))

    def peer_memory(self) -> bool:
        """Have the ranks mapped one another's mailbox (the one-kernel merge over NVLink peer memory)?"""
        on = C.c_int(0)
        A.check(A.load().qsgpu_comm_peer_memory(self.h, C.byref(on)))
        return bool(on.value)

    def barrier(self):
        A.check(A.load().qsgpu_comm_barrier(self.h))

    def allreduce_i64(self, values, op: int = 0):
        arr = (C.c_int64 * len(values))(*values)
        A.check(A.load().qsgpu_comm_allreduce_i64(self.h, arr, len(values), op))
        return list(arr)

    def merge_all(self, state: "AggState"):
        A.check(A.load().qsgpu_agg_merge_all(state.h, self.h))

    def lip_allreduce(self, lip: "LipFilter"):
        A.check(A.load().qsgpu_lip_allreduce(lip.h, self.h))

    def allgather_small(self, rel: "Relation", max_rows_per_rank: int) -> "Relation":
        out = C.c_void_p()
        A.check(A.load().qsgpu_relation_allgather_small(rel.h, self.h, max_rows_per_rank, C.byref(out)))
        return Relation(out, rel.schema, rel.names, rel.dev)

    def allgather(self, rel: "Relation") -> "Relation":
        out = C.c_void_p()
        A.check(A.load().qsgpu_relation_allgather(rel.h, self.h, C.byref(out)))
        return Relation(out, rel.schema, rel.names, rel.dev)

    def destroy(self):
        if self.h:
            A.load().qsgpu_comm_destroy(self.h)
        self.h = None
