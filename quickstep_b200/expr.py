"""Builder for flattened expression trees (qs_node arrays, include/qsgpu.h).

Mirrors the reference's serialization::Predicate / serialization::Scalar
protos (expressions/Expressions.proto:29-137): a node per ScalarLiteral,
ScalarAttribute, ScalarUnaryExpression, ScalarBinaryExpression,
ScalarSharedExpression, ComparisonPredicate, NegationPredicate,
ConjunctionPredicate, DisjunctionPredicate.  Children precede parents.
"""
from __future__ import annotations

import ctypes as C

from . import capi as A

_WIDTH = {A.QS_INT: 4, A.QS_LONG: 8, A.QS_FLOAT: 4, A.QS_DOUBLE: 8, A.QS_DATE: 8}


class ExprSet:
    def __init__(self):
        self.nodes: list[A.qs_node] = []
        self.pool = bytearray()
        self._c = None

    # ---- scalars ------------------------------------------------------
    def _add(self, **kw) -> int:
        n = A.qs_node()
        for k, v in kw.items():
            setattr(n, k, v)
        self.nodes.append(n)
        self._c = None
        return len(self.nodes) - 1

    def attr(self, attr_id: int, type_id: int, width: int = 0, side: int = 0) -> int:
        """ScalarAttribute; side 2 = join build side (JoinSide RIGHT_SIDE)."""
        return self._add(kind=A.QS_N_ATTRIBUTE, type=type_id, width=width or _WIDTH.get(type_id, 0),
                         a=attr_id, b=side)

    def lit_int(self, v: int) -> int:
        i = self._add(kind=A.QS_N_LITERAL, type=A.QS_INT, width=4)
        self.nodes[i].lit.i32 = v
        return i

    def lit_long(self, v: int) -> int:
        i = self._add(kind=A.QS_N_LITERAL, type=A.QS_LONG, width=8)
        self.nodes[i].lit.i64 = v
        return i

    def lit_float(self, v: float) -> int:
        i = self._add(kind=A.QS_N_LITERAL, type=A.QS_FLOAT, width=4)
        self.nodes[i].lit.f32 = v
        return i

    def lit_double(self, v: float) -> int:
        i = self._add(kind=A.QS_N_LITERAL, type=A.QS_DOUBLE, width=8)
        self.nodes[i].lit.f64 = v
        return i

    def lit_date(self, year: int, month: int, day: int) -> int:
        i = self._add(kind=A.QS_N_LITERAL, type=A.QS_DATE, width=8)
        self.nodes[i].lit.i64 = 0
        self.nodes[i].lit.date.year = year
        self.nodes[i].lit.date.month = month
        self.nodes[i].lit.date.day = day
        return i

    def lit_char(self, s: bytes) -> int:
        off = len(self.pool)
        self.pool += s + b"\0"
        i = self._add(kind=A.QS_N_LITERAL, type=A.QS_CHAR, width=len(s))
        self.nodes[i].lit.pool_offset = off
        return i

    def binary(self, op: int, a: int, b: int) -> int:
        return self._add(kind=A.QS_N_BINARY, op=op, a=a, b=b)

    def add(self, a, b): return self.binary(A.QS_ADD, a, b)
    def sub(self, a, b): return self.binary(A.QS_SUB, a, b)
    def mul(self, a, b): return self.binary(A.QS_MUL, a, b)
    def div(self, a, b): return self.binary(A.QS_DIV, a, b)
    def mod(self, a, b): return self.binary(A.QS_MOD, a, b)

    def neg(self, a: int) -> int:
        return self._add(kind=A.QS_N_UNARY, op=A.QS_NEGATE, a=a, b=-1)

    def cast(self, a: int, to_type: int) -> int:
        return self._add(kind=A.QS_N_UNARY, op=A.QS_CAST, type=to_type, a=a, b=-1)

    def shared(self, a: int, share_id: int) -> int:
        """ScalarSharedExpression (common sub-expression, evaluated once)."""
        return self._add(kind=A.QS_N_SHARED, a=a, b=share_id)

    # ---- predicates ---------------------------------------------------
    def cmp(self, op: int, a: int, b: int) -> int:
        return self._add(kind=A.QS_N_COMPARISON, op=op, a=a, b=b)

    def true_(self): return self._add(kind=A.QS_N_TRUE, a=-1, b=-1)
    def false_(self): return self._add(kind=A.QS_N_FALSE, a=-1, b=-1)
    def not_(self, a): return self._add(kind=A.QS_N_NEGATION, a=a, b=-1)

    def _fold(self, kind, ops):
        ops = list(ops)
        cur = ops[0]
        for o in ops[1:]:
            cur = self._add(kind=kind, a=cur, b=o)
        return cur

    def and_(self, *ops): return self._fold(A.QS_N_CONJUNCTION, ops)
    def or_(self, *ops): return self._fold(A.QS_N_DISJUNCTION, ops)

    # ---- C view ---------------------------------------------------------
    def c(self) -> A.qs_expr_set:
        if self._c is None:
            arr = (A.qs_node * max(1, len(self.nodes)))(*self.nodes)
            pool = bytes(self.pool) + b"\0"
            es = A.qs_expr_set()
            es.nodes = C.cast(arr, C.POINTER(A.qs_node))
            es.n_nodes = len(self.nodes)
            es.str_pool = pool
            es.str_pool_bytes = len(self.pool)
            self._c = (es, arr, pool)
        return self._c[0]

    def ptr(self):
        return C.pointer(self.c())
