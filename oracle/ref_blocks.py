"""TEST INFRASTRUCTURE (oracle side).  Reads the block files the UNMODIFIED reference engine writes (qsblk_*.qsb, 4 MB
each for the TPC-H relations) and describes the stripes the hot path stages, so that tests can hand the engine's own
bytes to qsgpu_stage_blocks.  Layouts restated from the reference:

  block            [int header_len][StorageBlockHeader proto][tuple-store sub-block][index sub-blocks]
                   storage/StorageBlockLayout.proto:103-124, storage/StorageBlock.cpp:95-160
  compressed       [tuple_id(int) num_tuples][int info_len][CompressedBlockInfo proto][dictionaries...]
  column store     [null bitmaps of uncompressed nullable attributes][stripe a: max_tuples x attribute_size(a)]...
                   storage/CompressedTupleStorageSubBlock.cpp:281-342,
                   storage/CompressedColumnStoreTupleStorageSubBlock.cpp:755-798
  dictionary       [u32 num_codes][u32 null_code][num_codes values]            (fixed-length types)
                   compression/CompressionDictionaryLite.hpp:40-51
  split row store  [Header{int num_tuples, int max_tid, u32 var_bytes, bool compact}][occupancy bitmap][slots]
                   slot = [null bitmap][fixed-length attributes back to back][(offset, length) u32 pairs]
                   storage/SplitRowStoreTupleStorageSubBlock.cpp:103-179, .hpp:356-361

The protobuf messages are decoded with a ~30-line wire-format reader (varint / fixed64 / length-delimited); no
generated code.  Nothing under quickstep_b200/ imports this module."""
import glob
import os
import struct

import numpy as np

BASIC_COLUMN_STORE, COMPRESSED_COLUMN_STORE, SPLIT_ROW_STORE = 0, 2, 3


def _varint(b, i):
    v, s = 0, 0
    while True:
        c = b[i]
        i += 1
        v |= (c & 0x7F) << s
        if c < 0x80:
            return v, i
        s += 7


def pb_fields(b):
    """-> list of (field number, wire type, value): value = int for varint / fixed, bytes for length-delimited."""
    out, i = [], 0
    while i < len(b):
        key, i = _varint(b, i)
        f, wt = key >> 3, key & 7
        if wt == 0:
            v, i = _varint(b, i)
        elif wt == 1:
            v = struct.unpack_from("<Q", b, i)[0]
            i += 8
        elif wt == 2:
            n, i = _varint(b, i)
            v = bytes(b[i:i + n])
            i += n
        elif wt == 5:
            v = struct.unpack_from("<I", b, i)[0]
            i += 4
        else:
            raise ValueError(f"wire type {wt}")
        out.append((f, wt, v))
    return out


def _packed_fixed64(v):
    return list(struct.unpack(f"<{len(v) // 8}Q", v))


class BlockHeader:
    def __init__(self, mem):
        (n,) = struct.unpack_from("<i", mem, 0)
        self.header_bytes = 4 + n
        self.tuple_store_size, self.sub_block_type, self.index_sizes = 0, -1, []
        for f, _wt, v in pb_fields(mem[4:4 + n]):
            if f == 1:                                        # StorageBlockLayoutDescription
                for f2, _w2, v2 in pb_fields(v):
                    if f2 == 2:                               # TupleStorageSubBlockDescription
                        for f3, _w3, v3 in pb_fields(v2):
                            if f3 == 1:
                                self.sub_block_type = v3
            elif f == 2:
                self.tuple_store_size = v
            elif f == 3:
                self.index_sizes = _packed_fixed64(v)


def read_compressed_column_store(mem, attr_widths):
    """mem: one block image; attr_widths[a] = catalog byte width of attribute a, or None for a variable-length attribute.
    -> dict(n_rows, stripes=[dict(encoding, code_width, offset, dict_offset, dict_entries) per attribute])."""
    h = BlockHeader(mem)
    assert h.sub_block_type == COMPRESSED_COLUMN_STORE
    sb = h.header_bytes
    n_rows, info_len = struct.unpack_from("<ii", mem, sb)
    attr_size, dict_size, null_bits, has_nulls = [], [], 0, []
    for f, _wt, v in pb_fields(mem[sb + 8: sb + 8 + info_len]):
        if f == 1:
            attr_size = _packed_fixed64(v)
        elif f == 2:
            dict_size = _packed_fixed64(v)
        elif f == 3:
            null_bits = v
        elif f == 4:
            has_nulls = list(v)
    assert len(attr_size) == len(attr_widths) == len(dict_size), (len(attr_size), len(attr_widths))
    pos = sb + 8 + info_len
    dict_at = []
    for a in range(len(attr_size)):
        dict_at.append(pos if dict_size[a] else None)
        pos += dict_size[a]
    # one BitVector<false> of null_bitmap_bits bits per uncompressed attribute that holds NULLs
    # (CompressedColumnStoreTupleStorageSubBlock.cpp:755-775)
    null_at = [None] * len(attr_size)
    if null_bits:
        for a in range(len(has_nulls)):
            if has_nulls[a]:
                null_at[a] = pos
                pos += ((null_bits + 63) // 64) * 8
    tuple_len = sum(attr_size)
    max_tuples = (sb + h.tuple_store_size - pos) // tuple_len
    stripes = []
    for a in range(len(attr_size)):
        w = attr_widths[a]
        s = dict(offset=pos, code_width=int(attr_size[a]), dict_offset=None, dict_entries=0, encoding="skip",
                 null_offset=null_at[a])
        if w is not None:
            if dict_size[a]:
                num_codes, null_code = struct.unpack_from("<II", mem, dict_at[a])
                s.update(encoding="dict", dict_offset=dict_at[a] + 8, dict_entries=num_codes, null_code=null_code)
            elif attr_size[a] != w:
                s.update(encoding="truncated")
            else:
                s.update(encoding="plain")
        stripes.append(s)
        pos += max_tuples * attr_size[a]
    return dict(n_rows=n_rows, max_tuples=max_tuples, stripes=stripes, n_attrs=len(attr_size))


def read_basic_column_store(mem, attr_widths, nullable):
    """Uncompressed column store (storage/BasicColumnStoreTupleStorageSubBlock.cpp:100-183):
    [Header{int num_tuples, int nulls_in_sort_column}][one BitVector<false>(max_tuples) per NULL-able attribute]
    [stripe a: max_tuples x width(a)]...   -> dict(n_rows, stripes=[dict(offset, null_offset or None)])."""
    h = BlockHeader(mem)
    assert h.sub_block_type == BASIC_COLUMN_STORE
    sb = h.header_bytes
    n_rows, _nulls_in_sort = struct.unpack_from("<ii", mem, sb)
    size, fixed, n_null = h.tuple_store_size, sum(attr_widths), sum(1 for x in nullable if x)
    max_tuples = ((size - 8) << 3) // ((fixed << 3) + n_null)
    bitmap_bytes = ((max_tuples + 63) // 64) * 8
    max_tuples = (size - 8 - n_null * bitmap_bytes) // fixed
    bitmap_bytes = ((max_tuples + 63) // 64) * 8
    pos = sb + 8
    null_at = []
    for a in range(len(attr_widths)):
        null_at.append(pos if nullable[a] else None)
        if nullable[a]:
            pos += bitmap_bytes
    stripes = []
    for a, w in enumerate(attr_widths):
        stripes.append(dict(offset=pos, null_offset=null_at[a], encoding="plain"))
        pos += max_tuples * w
    return dict(n_rows=n_rows, max_tuples=max_tuples, stripes=stripes)


def read_split_row_store(mem, fixed_widths, n_varlen, min_varlen_bytes, n_nullable=0):
    """fixed_widths: byte widths of the fixed-length attributes in attribute order.
    -> dict(n_rows, first_slot offset, slot_bytes, attr_offsets (inside a slot), occupied row indices)."""
    h = BlockHeader(mem)
    assert h.sub_block_type == SPLIT_ROW_STORE
    sb = h.header_bytes
    n_rows, max_tid, _var_bytes = struct.unpack_from("<iiI", mem, sb)
    header_bytes = 16
    null_bytes = 0 if n_nullable == 0 else (1 if n_nullable < 9 else 2 if n_nullable < 17 else 4 if n_nullable < 33 else ((n_nullable + 63) // 64) * 8)
    slot = null_bytes + sum(fixed_widths) + n_varlen * 8
    max_tuples = (h.tuple_store_size - header_bytes) // (slot + min_varlen_bytes)
    occ_bytes = ((max_tuples + 63) // 64) * 8
    occ = np.frombuffer(mem, dtype="<u8", count=occ_bytes // 8, offset=sb + header_bytes)
    rows = [i for i in range(max_tid + 1) if (int(occ[i >> 6]) >> (63 - (i & 63))) & 1]      # MSB-first bits (BitVector.hpp:934)
    assert len(rows) == n_rows, (len(rows), n_rows)
    offs, o = [], null_bytes
    for w in fixed_widths:
        offs.append(o)
        o += w
    return dict(n_rows=n_rows, first_slot=sb + header_bytes + occ_bytes, slot_bytes=slot, attr_offsets=offs, rows=rows,
                contiguous=rows == list(range(n_rows)), null_bytes=null_bytes)


def load_blocks(storage_dir):
    """-> [(path, bytes)] of every block file, in block-id order."""
    out = []
    for p in sorted(glob.glob(os.path.join(storage_dir, "qsblk_*.qsb"))):
        with open(p, "rb") as f:
            out.append((p, f.read()))
    return out
