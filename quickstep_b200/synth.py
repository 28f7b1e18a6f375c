"""Synthetic TPC-H-shaped relations generated directly in HBM with torch (bench harness only;
there is no network for dbgen output at SF10+ and dbgen at SF10 takes minutes).

Value domains and correlations follow dbgen (SURVEY.md section 8d "Data"):
  orders:   o_orderkey sparse (8 used of every 32), o_orderdate uniform 1992-01-01..1998-08-02,
            o_custkey in [1, n_cust] with custkey % 3 != 0, o_shippriority 0
  lineitem: ~4 rows per order, sorted on l_orderkey; l_shipdate = o_orderdate + 1..121 days;
            l_quantity 1..50; l_discount 0.00..0.10; l_tax 0.00..0.08;
            l_extendedprice = quantity * retail price (900.00..2100.00), cents;
            l_returnflag R/A if receipt date <= 1995-06-17 else N; l_linestatus O if shipped after
            1995-06-17 else F
  customer: c_custkey 1..n_cust, c_mktsegment uniform over the five segments
DATE columns are Quickstep DateLit structs {int32 year; u8 month; u8 day; 2 pad} packed in int64.
"""
from __future__ import annotations

import torch

from . import capi as A
from . import tpch as T

_EPOCH_1992 = 8035        # days from 1970-01-01 to 1992-01-01
_CUTOFF = 9298            # 1995-06-17
SEGMENTS = [b"AUTOMOBILE", b"BUILDING", b"FURNITURE", b"MACHINERY", b"HOUSEHOLD"]


def datelit_from_days(days: torch.Tensor) -> torch.Tensor:
    """days since 1970-01-01 (int64) -> packed DateLit (int64).  Civil-from-days, proleptic Gregorian."""
    z = days + 719468
    era = torch.div(z, 146097, rounding_mode="floor")
    doe = z - era * 146097
    yoe = torch.div(doe - torch.div(doe, 1460, rounding_mode="floor") + torch.div(doe, 36524, rounding_mode="floor")
                    - torch.div(doe, 146096, rounding_mode="floor"), 365, rounding_mode="floor")
    y = yoe + era * 400
    doy = doe - (365 * yoe + torch.div(yoe, 4, rounding_mode="floor") - torch.div(yoe, 100, rounding_mode="floor"))
    mp = torch.div(5 * doy + 2, 153, rounding_mode="floor")
    d = doy - torch.div(153 * mp + 2, 5, rounding_mode="floor") + 1
    m = torch.where(mp < 10, mp + 3, mp - 9)
    y = y + (m <= 2).to(torch.int64)
    return (y & 0xFFFFFFFF) | (m << 32) | (d << 40)


def generate(n_lineitem: int, seed: int, device, key_base: int = 0):
    """-> dict of torch tensors (device) for customer / orders / lineitem columns."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    ri = lambda lo, hi, n: torch.randint(lo, hi, (n,), generator=g, device=device, dtype=torch.int64)
    n_orders = max(8, n_lineitem // 4)
    n_cust = max(3, n_orders // 10)
    out = {}
    # ---- orders
    dense = torch.arange(n_orders, device=device, dtype=torch.int64) + key_base
    out["o_orderkey"] = ((((dense >> 3) << 5) | (dense & 7)) + 1).to(torch.int32)
    ck = ri(0, (n_cust * 2) // 3, n_orders)
    out["o_custkey"] = (ck + torch.div(ck, 2, rounding_mode="floor") + 1).clamp_(max=n_cust).to(torch.int32)  # skips multiples of 3
    o_days = _EPOCH_1992 + ri(0, 2406, n_orders)
    out["o_orderdate"] = datelit_from_days(o_days)
    out["o_shippriority"] = torch.zeros(n_orders, device=device, dtype=torch.int32)
    # ---- lineitem (sorted on l_orderkey)
    oidx, _ = torch.sort(ri(0, n_orders, n_lineitem))
    out["l_orderkey"] = out["o_orderkey"][oidx].contiguous()
    ship = o_days[oidx] + ri(1, 122, n_lineitem)
    del oidx
    out["l_shipdate"] = datelit_from_days(ship)
    receipt = ship + ri(1, 31, n_lineitem)
    qty = ri(1, 51, n_lineitem)
    out["l_quantity"] = qty.to(torch.float64)
    cents = ri(90000, 210001, n_lineitem) * qty
    out["l_extendedprice"] = cents.to(torch.float64) / 100.0
    del cents, qty
    out["l_discount"] = ri(0, 11, n_lineitem).to(torch.float64) / 100.0
    out["l_tax"] = ri(0, 9, n_lineitem).to(torch.float64) / 100.0
    ra = torch.where(ri(0, 2, n_lineitem) == 0, ord("R"), ord("A"))
    out["l_returnflag"] = torch.where(receipt <= _CUTOFF, ra, torch.full_like(ra, ord("N"))).to(torch.uint8)
    out["l_linestatus"] = torch.where(ship > _CUTOFF, ord("O"), ord("F")).to(torch.uint8)
    del ra, receipt, ship
    # ---- customer
    out["c_custkey"] = torch.arange(1, n_cust + 1, device=device, dtype=torch.int32)
    seg = torch.tensor([list(s.ljust(10, b"\0")) for s in SEGMENTS], dtype=torch.uint8, device=device)
    out["c_mktsegment"] = seg[ri(0, 5, n_cust)].contiguous()       # [n_cust, 10] bytes
    out["_stats"] = dict(c_custkey_min=1, c_custkey_max=n_cust,
                         o_orderkey_min=int(out["o_orderkey"][0]), o_orderkey_max=int(out["o_orderkey"][-1]),
                         orders_rows=n_orders, lineitem_rows=n_lineitem, customer_rows=n_cust,
                         t2_estimate=n_orders // 4, groups_estimate=max(1024, n_orders // 8),
                         t4_capacity=max(1024, n_lineitem // 4))
    return out


def _padded(t: torch.Tensor) -> torch.Tensor:
    """Relations wrapped without a copy must be readable 16 bytes past the last row (qsgpu.h)."""
    flat = t.reshape(-1).view(torch.uint8) if t.dtype != torch.uint8 else t.reshape(-1)
    buf = torch.zeros(flat.numel() + 256, dtype=torch.uint8, device=t.device)
    buf[: flat.numel()] = flat
    return buf


def wrap_relations(E, cols: dict, dev: int):
    """Wrap the generated tensors as device relations (no copy besides tail padding)."""
    rels, keep = {}, []
    for name, schema in (("customer", T.CUSTOMER), ("orders", T.ORDERS), ("lineitem", T.LINEITEM)):
        bufs = [_padded(cols[n]) for (n, _t, _w) in schema]
        keep.append(bufs)
        n_rows = cols[schema[0][0]].shape[0]
        rels[name] = E.Relation.wrap([(t, w) for (_n, t, w) in schema], [b.data_ptr() for b in bufs], n_rows,
                                     [n for (n, _t, _w) in schema], dev, keep=bufs)
    return rels


# CHAR(1) flags stay native: a 1-byte code saves nothing
LINEITEM_CODED = ("l_quantity", "l_discount", "l_tax", "l_shipdate")


def wrap_lineitem_coded(E, cols: dict, dev: int, coded=LINEITEM_CODED):
    """lineitem with the low-cardinality attributes resident as dictionary codes (what the reference's
    compressed-column-store blocks of lineitem hold, benchmarks/tpch/create.sql): the sorted dictionary and the
    code column of each attribute are derived on the device with torch.unique (harness plumbing, like the
    generator itself) and handed over with qsgpu_relation_wrap + qsgpu_relation_set_dictionary.
    -> (relation, {name: (code width, dictionary entries)})"""
    import numpy as np
    schema = T.LINEITEM
    bufs, dicts, info = [], {}, {}
    for a, (name, t, w) in enumerate(schema):
        col = cols[name]
        if name not in coded:
            bufs.append(_padded(col))
            continue
        if t == A.QS_DATE:      # packed DateLit -> order-preserving key (year, month, day) for the sort
            key = ((col & 0xFFFFFFFF) << 16) | (((col >> 32) & 0xFF) << 8) | ((col >> 40) & 0xFF)
            uniq, inv = torch.unique(key, sorted=True, return_inverse=True)
            d = ((uniq >> 16) & 0xFFFFFFFF) | (((uniq >> 8) & 0xFF) << 32) | ((uniq & 0xFF) << 40)
            del key
        else:
            d, inv = torch.unique(col.reshape(col.shape[0], -1)[:, 0] if col.dim() > 1 else col, sorted=True, return_inverse=True)
        n = d.numel()
        cw = 1 if n <= 256 else 2 if n <= 65536 else 4
        codes = inv.to({1: torch.uint8, 2: torch.int16, 4: torch.int32}[cw])   # int16 holds the u16 bit pattern
        del inv
        bufs.append(_padded(codes))
        dicts[a] = (cw, d.cpu().numpy())
        info[name] = (cw, n)
    n_rows = cols[schema[0][0]].shape[0]
    rel = E.Relation.wrap([(t, w) for (_n, t, w) in schema], [b.data_ptr() for b in bufs], n_rows,
                          [n for (n, _t, _w) in schema], dev, keep=bufs)
    for a, (cw, d) in dicts.items():
        t, w = schema[a][1], schema[a][2]
        if t == A.QS_DATE:
            from .table import DATE_DTYPE
            d = d.view(DATE_DTYPE)
        elif t == A.QS_CHAR:
            d = np.ascontiguousarray(d).view(f"S{w}")
        rel.set_dictionary(a, cw, d)
    return rel, info


def host_table(cols: dict, schema, n: int | None = None):
    """Host (numpy) copy of one relation's columns, for the CPU baseline and the e2e legs."""
    from .table import Column, HostTable, DATE_DTYPE
    out = []
    for (name, t, w) in schema:
        a = cols[name][:n] if n is not None else cols[name]
        a = a.cpu().numpy()
        if t == A.QS_DATE:
            a = a.view(DATE_DTYPE)
        elif t == A.QS_CHAR:
            a = a.reshape(a.shape[0], -1).view(f"S{w}").reshape(-1)
        out.append(Column(name, t, a, w))
    return HostTable(schema[0][0].split("_")[0], out)


# ------------------------------------------------------------------------------------------------------------------
# ONE database, cut into chunks any rank can regenerate (bench.py at SF100, 1 / 2 / 4 / 8 GPUs).
#
# The database of `n_lineitem` rows is a fixed sequence of N chunks; chunk c holds orders [o_lo, o_hi), the lineitem
# rows of exactly those orders, and customers [c_lo, c_hi), all drawn from a generator seeded with (seed, c).  The
# content of the database therefore does not depend on how many ranks there are: rank r of `world` owns the chunks
# [r*N/world, (r+1)*N/world) -- lineitem block-partitioned on l_orderkey boundaries (it is sorted on it,
# benchmarks/tpch/create.sql:112), orders and customer in matching shares -- and rank 0 can regenerate every chunk
# for the CPU oracle.  Same value domains and correlations as generate() above.

def db_shape(n_lineitem: int, n_chunks: int | None = None):
    n_orders = max(8, n_lineitem // 4)
    n_cust = max(3, n_orders // 10)
    if n_chunks is None:
        n_chunks = 256 if n_orders >= 256 * 4096 else 8
    return dict(n_lineitem=n_lineitem, n_orders=n_orders, n_cust=n_cust, n_chunks=n_chunks)


def _cut(total: int, n: int, i: int) -> int:
    return (i * total) // n


def rank_chunks(shape, world: int, rank: int):
    n = shape["n_chunks"]
    return range(_cut(n, world, rank), _cut(n, world, rank + 1))


def chunk_rows(shape, c: int):
    """-> ((o_lo, o_hi), (l_lo, l_hi), (c_lo, c_hi)) global row ranges of chunk c."""
    n = shape["n_chunks"]
    return ((_cut(shape["n_orders"], n, c), _cut(shape["n_orders"], n, c + 1)),
            (_cut(shape["n_lineitem"], n, c), _cut(shape["n_lineitem"], n, c + 1)),
            (_cut(shape["n_cust"], n, c), _cut(shape["n_cust"], n, c + 1)))


def generate_chunk(shape, c: int, seed: int, device):
    """-> dict of device tensors: the orders, lineitem and customer rows of chunk c."""
    (o_lo, o_hi), (l_lo, l_hi), (c_lo, c_hi) = chunk_rows(shape, c)
    n_o, n_l, n_c, n_cust = o_hi - o_lo, l_hi - l_lo, c_hi - c_lo, shape["n_cust"]
    g = torch.Generator(device=device)
    g.manual_seed(seed * 1_000_003 + c)
    ri = lambda lo, hi, n: torch.randint(lo, hi, (n,), generator=g, device=device, dtype=torch.int64)
    out = {}
    dense = torch.arange(o_lo, o_hi, device=device, dtype=torch.int64)
    out["o_orderkey"] = ((((dense >> 3) << 5) | (dense & 7)) + 1).to(torch.int32)
    ck = ri(0, max(1, (n_cust * 2) // 3), n_o)
    out["o_custkey"] = (ck + torch.div(ck, 2, rounding_mode="floor") + 1).clamp_(max=n_cust).to(torch.int32)
    o_days = _EPOCH_1992 + ri(0, 2406, n_o)
    out["o_orderdate"] = datelit_from_days(o_days)
    out["o_shippriority"] = torch.zeros(n_o, device=device, dtype=torch.int32)
    oidx, _ = torch.sort(ri(0, max(1, n_o), n_l))
    out["l_orderkey"] = out["o_orderkey"][oidx].contiguous() if n_o else torch.zeros(0, device=device, dtype=torch.int32)
    ship = (o_days[oidx] if n_o else torch.zeros(0, device=device, dtype=torch.int64)) + ri(1, 122, n_l)
    out["l_shipdate"] = datelit_from_days(ship)
    receipt = ship + ri(1, 31, n_l)
    qty = ri(1, 51, n_l)
    out["l_quantity"] = qty.to(torch.float64)
    out["l_extendedprice"] = (ri(90000, 210001, n_l) * qty).to(torch.float64) / 100.0
    out["l_discount"] = ri(0, 11, n_l).to(torch.float64) / 100.0
    out["l_tax"] = ri(0, 9, n_l).to(torch.float64) / 100.0
    ra = torch.where(ri(0, 2, n_l) == 0, ord("R"), ord("A"))
    out["l_returnflag"] = torch.where(receipt <= _CUTOFF, ra, torch.full_like(ra, ord("N"))).to(torch.uint8)
    out["l_linestatus"] = torch.where(ship > _CUTOFF, ord("O"), ord("F")).to(torch.uint8)
    out["c_custkey"] = torch.arange(c_lo + 1, c_hi + 1, device=device, dtype=torch.int32)
    seg = torch.tensor([list(s.ljust(10, b"\0")) for s in SEGMENTS], dtype=torch.uint8, device=device)
    out["c_mktsegment"] = seg[ri(0, 5, n_c)].contiguous()
    return out


def generate_host(shape, chunks, seed: int, device):
    """The rows of `chunks` (a contiguous range) as HOST numpy columns in Quickstep's native layouts:
    -> {"customer": [arrays in schema order], "orders": [...], "lineitem": [...]}.  Generated chunk by chunk on the
    device (torch) and copied straight into preallocated host arrays."""
    import numpy as np
    from .table import DATE_DTYPE
    chunks = list(chunks)
    first, last = (chunks[0], chunks[-1]) if chunks else (0, -1)
    (o0, _), (l0, _), (c0, _) = chunk_rows(shape, first) if chunks else ((0, 0), (0, 0), (0, 0))
    (_, o1), (_, l1), (_, c1) = chunk_rows(shape, last) if chunks else ((0, 0), (0, 0), (0, 0))
    n = {"customer": c1 - c0, "orders": o1 - o0, "lineitem": l1 - l0}
    base = {"customer": c0, "orders": o0, "lineitem": l0}
    np_raw = {A.QS_INT: np.int32, A.QS_DOUBLE: np.float64, A.QS_DATE: np.int64}
    host = {}
    for rel, schema in (("customer", T.CUSTOMER), ("orders", T.ORDERS), ("lineitem", T.LINEITEM)):
        for (name, t, w) in schema:
            host[name] = np.empty((n[rel], w), dtype=np.uint8) if t == A.QS_CHAR else np.empty(n[rel], dtype=np_raw[t])
    for c in chunks:
        cols = generate_chunk(shape, c, seed, device)
        rows = dict(zip(("orders", "lineitem", "customer"), chunk_rows(shape, c)))
        for rel, schema in (("customer", T.CUSTOMER), ("orders", T.ORDERS), ("lineitem", T.LINEITEM)):
            lo, hi = rows[rel][0] - base[rel], rows[rel][1] - base[rel]
            for (name, _t, w) in schema:
                src = cols[name].reshape(hi - lo, -1) if host[name].ndim == 2 else cols[name]
                torch.from_numpy(host[name][lo:hi]).copy_(src)
        del cols
    out = {}
    for rel, schema in (("customer", T.CUSTOMER), ("orders", T.ORDERS), ("lineitem", T.LINEITEM)):
        arrs = []
        for (name, t, w) in schema:
            a = host[name]
            if t == A.QS_DATE:
                a = a.view(DATE_DTYPE)
            elif t == A.QS_CHAR:
                a = a.view(f"S{w}").reshape(-1)
            arrs.append(a)
        out[rel] = arrs
    return out


def host_tables(host):
    """generate_host() columns -> {"customer" | "orders" | "lineitem": HostTable} for the CPU oracle (no copy)."""
    from .table import Column, HostTable
    return {rel: HostTable(rel, [Column(nm, t, a, w) for (nm, t, w), a in zip(schema, host[rel])])
            for rel, schema in (("customer", T.CUSTOMER), ("orders", T.ORDERS), ("lineitem", T.LINEITEM))}
