// Owning builder for the flattened expression trees of the C ABI (qs_node arrays,
// include/qsgpu.h:44-83).  In an in-tree build this is what the lowering of
// serialization::Predicate / serialization::Scalar protos
// (expressions/Expressions.proto:29-137) produces: node kinds, comparison ids,
// operation ids and type ids are the reference's enum values, so the walk is 1:1.
#pragma once

#include <cstring>
#include <string>
#include <vector>

#include "QsTypes.hpp"

namespace quickstep {

class ExprSet {
 public:
  int attr(attribute_id a, qs_attr t, int side = 0) {
    qs_node x{}; x.kind = QS_N_ATTRIBUTE; x.type = t.type; x.width = t.width; x.a = a; x.b = side; return add(x);
  }
  int lit_int(std::int32_t v) { qs_node x{}; x.kind = QS_N_LITERAL; x.type = QS_INT; x.lit.i32 = v; return add(x); }
  int lit_long(std::int64_t v) { qs_node x{}; x.kind = QS_N_LITERAL; x.type = QS_LONG; x.lit.i64 = v; return add(x); }
  int lit_float(float v) { qs_node x{}; x.kind = QS_N_LITERAL; x.type = QS_FLOAT; x.lit.f32 = v; return add(x); }
  int lit_double(double v) { qs_node x{}; x.kind = QS_N_LITERAL; x.type = QS_DOUBLE; x.lit.f64 = v; return add(x); }
  int lit_date(int y, int m, int d) {
    qs_node x{}; x.kind = QS_N_LITERAL; x.type = QS_DATE;
    x.lit.date.year = y; x.lit.date.month = static_cast<std::uint8_t>(m); x.lit.date.day = static_cast<std::uint8_t>(d);
    return add(x);
  }
  int lit_char(const std::string &s) {
    qs_node x{}; x.kind = QS_N_LITERAL; x.type = QS_CHAR; x.width = static_cast<std::uint16_t>(s.size());
    x.lit.pool_offset = pool_.size(); pool_ += s; return add(x);
  }
  int binary(int op, int a, int b) { qs_node x{}; x.kind = QS_N_BINARY; x.op = static_cast<std::uint16_t>(op); x.a = a; x.b = b; return add(x); }
  int negate(int a) { qs_node x{}; x.kind = QS_N_UNARY; x.op = QS_NEGATE; x.a = a; return add(x); }
  int cast(int a, int to_type) { qs_node x{}; x.kind = QS_N_UNARY; x.op = QS_CAST; x.type = static_cast<std::uint16_t>(to_type); x.a = a; return add(x); }
  int shared(int a, int share_id) { qs_node x{}; x.kind = QS_N_SHARED; x.a = a; x.b = share_id; return add(x); }
  int cmp(int op, int a, int b) { qs_node x{}; x.kind = QS_N_COMPARISON; x.op = static_cast<std::uint16_t>(op); x.a = a; x.b = b; return add(x); }
  int negation(int a) { qs_node x{}; x.kind = QS_N_NEGATION; x.a = a; return add(x); }
  // n-ary conjunction / disjunction lists fold into left-deep binary chains (qsgpu.h:56-57)
  int conj(const std::vector<int> &ops) { return fold(QS_N_CONJUNCTION, ops); }
  int disj(const std::vector<int> &ops) { return fold(QS_N_DISJUNCTION, ops); }
  int true_() { qs_node x{}; x.kind = QS_N_TRUE; return add(x); }

  // Appends every node of `o` (children rebased) and returns the index offset:
  // a work order that evaluates a Predicate and a scalar group of the same
  // QueryContext hands them to the C ABI as ONE expression set.
  int append(const ExprSet &o) {
    const int off = static_cast<int>(nodes_.size());
    const std::uint64_t poff = pool_.size();
    for (qs_node x : o.nodes_) {
      switch (x.kind) {
        case QS_N_UNARY: case QS_N_SHARED: case QS_N_NEGATION: x.a += off; break;
        case QS_N_BINARY: case QS_N_COMPARISON: case QS_N_CONJUNCTION: case QS_N_DISJUNCTION: x.a += off; x.b += off; break;
        case QS_N_LITERAL: if (x.type == QS_CHAR) x.lit.pool_offset += poff; break;
        default: break;
      }
      nodes_.push_back(x);
    }
    pool_ += o.pool_;
    return off;
  }

  qs_expr_set view() const {
    qs_expr_set e{};
    e.nodes = nodes_.data(); e.n_nodes = static_cast<std::uint32_t>(nodes_.size());
    e.str_pool = pool_.data(); e.str_pool_bytes = static_cast<std::uint32_t>(pool_.size());
    return e;
  }
  std::size_t size() const { return nodes_.size(); }
  // Bit a set: some ATTRIBUTE node of join side `side` (0 = scanned / probe relation, 2 = build relation)
  // names attribute a.  What an operator asks the device cache to hold before its work orders run.
  std::uint64_t referencedAttributes(int side = 0) const {
    std::uint64_t m = 0;
    for (const qs_node &x : nodes_)
      if (x.kind == QS_N_ATTRIBUTE && x.b == side && x.a >= 0 && x.a < 64) m |= 1ull << x.a;
    return m;
  }

 private:
  int add(const qs_node &x) { nodes_.push_back(x); return static_cast<int>(nodes_.size()) - 1; }
  int fold(int kind, const std::vector<int> &ops) {
    QS_CHECK(!ops.empty());
    int acc = ops[0];
    for (std::size_t i = 1; i < ops.size(); ++i) {
      qs_node x{}; x.kind = static_cast<std::uint16_t>(kind); x.a = acc; x.b = ops[i]; acc = add(x);
    }
    return acc;
  }
  std::vector<qs_node> nodes_;
  std::string pool_;
};

}  // namespace quickstep
