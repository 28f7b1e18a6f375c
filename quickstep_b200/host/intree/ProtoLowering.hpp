// IN-TREE binding, part 1: the reference's own serialized expression trees -> the C ABI's qs_node arrays.
//
// Compiled against the REAL reference headers (generated Expressions.pb.h, types/*.pb.h), not against the stand-ins
// of quickstep_b200/host/QsTypes.hpp: tests/test_intree_boundary.py runs `g++ -fsyntax-only` on intree/GpuWorkOrders.cpp
// with /root/reference and the build tree of oracle/build_ref.sh on the include path.
//
//   serialization::Predicate / serialization::Scalar        expressions/Expressions.proto:29-137
//   Predicate::getProto() / Scalar::getProto()              expressions/predicate/Predicate.hpp, expressions/scalar/Scalar.hpp
//   serialization::TypedValue / Type                        types/TypedValue.proto, types/Type.proto
// The proto's Type.TypeID numbers DATE as 10; the C ABI uses types/TypeID.hpp's enum values (kDate = 6).
#pragma once

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "expressions/Expressions.pb.h"
#include "types/DatetimeLit.hpp"
#include "types/Type.hpp"
#include "types/Type.pb.h"
#include "types/TypeFactory.hpp"
#include "types/TypeID.hpp"
#include "types/TypedValue.hpp"
#include "types/TypedValue.pb.h"
#include "types/operations/Operation.pb.h"
#include "types/operations/binary_operations/BinaryOperation.hpp"
#include "types/operations/binary_operations/BinaryOperationFactory.hpp"
#include "types/operations/unary_operations/UnaryOperation.hpp"
#include "types/operations/unary_operations/UnaryOperationFactory.hpp"

#include "glog/logging.h"

#include "qsgpu.h"

namespace quickstep {
namespace gpu {

// A qs_expr_set under construction (children precede their parents, as qsgpu.h asks).
class ExprBuilder {
 public:
  int add(const qs_node &n) { nodes_.push_back(n); return static_cast<int>(nodes_.size()) - 1; }
  std::uint64_t addString(const std::string &s, std::uint16_t width) {
    const std::uint64_t off = pool_.size();
    pool_.append(s);
    pool_.append(width > s.size() ? width - s.size() : 0, '\0');
    return off;
  }
  qs_expr_set view() const {
    qs_expr_set es{};
    es.nodes = nodes_.data();
    es.n_nodes = static_cast<std::uint32_t>(nodes_.size());
    es.str_pool = pool_.data();
    es.str_pool_bytes = static_cast<std::uint32_t>(pool_.size());
    return es;
  }
  const qs_node &node(int i) const { return nodes_[static_cast<std::size_t>(i)]; }

 private:
  std::vector<qs_node> nodes_;
  std::string pool_;
};

inline std::uint16_t LowerTypeID(const serialization::Type &t, std::uint16_t *width) {
  *width = 0;
  switch (t.type_id()) {
    case serialization::Type::INT: return QS_INT;
    case serialization::Type::LONG: return QS_LONG;
    case serialization::Type::FLOAT: return QS_FLOAT;
    case serialization::Type::DOUBLE: return QS_DOUBLE;
    case serialization::Type::DATE: return QS_DATE;
    case serialization::Type::CHAR:
      *width = static_cast<std::uint16_t>(t.GetExtension(serialization::CharType::length));
      return QS_CHAR;
    default:
      LOG(FATAL) << "GPU path: type " << t.type_id() << " is not staged on the device (VARCHAR / DATETIME / intervals stay on the CPU path)";
  }
  return QS_INT;
}

// The attribute types of the relation(s) the expression reads: the proto of a ScalarAttribute carries only ids.
struct AttributeTypes {
  // relation_id -> per attribute (type, width); filled by the caller from CatalogRelationSchema
  std::vector<std::pair<int, std::vector<qs_attr>>> relations;
  // The expression is evaluated over ONE relation (a Select / Aggregation / BuildLIPFilter predicate, a BuildHash
  // predicate): join sides are dropped.  The optimizer leaves them on, e.g., a BuildHashOperator's build predicate that it
  // pushed down from a join (TPC-H Q19 / Q21: its attributes still say RIGHT_SIDE), and the reference's single-relation
  // evaluation (Predicate::getAllMatches over one accessor) never looks at them.
  bool single_relation = false;
  qs_attr lookup(int relation_id, int attribute_id) const {
    for (const auto &r : relations)
      if (r.first == relation_id) return r.second[static_cast<std::size_t>(attribute_id)];
    LOG(FATAL) << "GPU path: attribute of an unknown relation " << relation_id;
    return qs_attr{};
  }
};

// Constant sub-expressions.  The optimizer does not fold them: TPC-H Q6's `date '1994-01-01' + interval '1' year` reaches
// the QueryContext as ScalarBinaryExpression(ADD, literal DATE, literal YEAR-MONTH INTERVAL), and the reference evaluates
// it once when the expression is reconstructed (ScalarBinaryExpression::initHelper, static_value_).  The lowering does the
// same with the reference's own operations, so whatever they can apply to literals -- date arithmetic included -- reaches
// the device as ONE literal of the result type.
inline bool FoldStaticScalar(const serialization::Scalar &s, TypedValue *value, const Type **type) {
  switch (s.data_source()) {
    case serialization::Scalar::LITERAL:
      *value = TypedValue::ReconstructFromProto(s.GetExtension(serialization::ScalarLiteral::literal));
      *type = &TypeFactory::ReconstructFromProto(s.GetExtension(serialization::ScalarLiteral::literal_type));
      return true;
    case serialization::Scalar::UNARY_EXPRESSION: {
      TypedValue v;
      const Type *t = nullptr;
      if (!FoldStaticScalar(s.GetExtension(serialization::ScalarUnaryExpression::operand), &v, &t)) return false;
      const UnaryOperation &op = UnaryOperationFactory::ReconstructFromProto(s.GetExtension(serialization::ScalarUnaryExpression::operation));
      *type = op.resultTypeForArgumentType(*t);
      if (*type == nullptr) return false;
      *value = op.applyToChecked(v, *t);
      return true;
    }
    case serialization::Scalar::BINARY_EXPRESSION: {
      TypedValue l, r;
      const Type *lt = nullptr, *rt = nullptr;
      if (!FoldStaticScalar(s.GetExtension(serialization::ScalarBinaryExpression::left_operand), &l, &lt) ||
          !FoldStaticScalar(s.GetExtension(serialization::ScalarBinaryExpression::right_operand), &r, &rt))
        return false;
      const BinaryOperation &op = BinaryOperationFactory::ReconstructFromProto(s.GetExtension(serialization::ScalarBinaryExpression::operation));
      *type = op.resultTypeForArgumentTypes(*lt, *rt);
      if (*type == nullptr) return false;
      *value = op.applyToChecked(l, *lt, r, *rt);
      return true;
    }
    default:
      return false;
  }
}

// A literal node from a (folded) value of the reference's own TypedValue.
inline int LowerTypedValue(const TypedValue &v, const Type &type, ExprBuilder *b) {
  qs_node n{};
  n.kind = QS_N_LITERAL;
  CHECK(!v.isNull()) << "GPU path: NULL literals keep their CPU operators";
  switch (type.getTypeID()) {
    case kInt: n.type = QS_INT; n.lit.i32 = v.getLiteral<int>(); break;
    case kLong: n.type = QS_LONG; n.lit.i64 = v.getLiteral<std::int64_t>(); break;
    case kFloat: n.type = QS_FLOAT; n.lit.f32 = v.getLiteral<float>(); break;
    case kDouble: n.type = QS_DOUBLE; n.lit.f64 = v.getLiteral<double>(); break;
    case kDate: {
      const DateLit d = v.getLiteral<DateLit>();
      n.type = QS_DATE;
      n.lit.date.year = d.year;
      n.lit.date.month = d.month;
      n.lit.date.day = d.day;
      break;
    }
    case kChar:
    case kVarChar: {
      const std::string bytes(static_cast<const char *>(v.getOutOfLineData()), v.getAsciiStringLength());
      n.type = QS_CHAR;
      n.width = static_cast<std::uint16_t>(type.getTypeID() == kChar ? type.maximumByteLength() : bytes.size());
      n.lit.pool_offset = b->addString(bytes, n.width);
      break;
    }
    default:
      LOG(FATAL) << "GPU path: a constant of type " << type.getName() << " is not staged on the device";
  }
  return b->add(n);
}

inline int LowerScalar(const serialization::Scalar &s, const AttributeTypes &types, ExprBuilder *b) {
  qs_node n{};
  switch (s.data_source()) {
    case serialization::Scalar::LITERAL: {
      const serialization::TypedValue &v = s.GetExtension(serialization::ScalarLiteral::literal);
      std::uint16_t w = 0;
      n.kind = QS_N_LITERAL;
      const serialization::Type &lit_type = s.GetExtension(serialization::ScalarLiteral::literal_type);
      if (lit_type.type_id() == serialization::Type::VAR_CHAR) {
        // The parser types a quoted string VARCHAR(n) (TPC-H Q3's c_mktsegment = 'BUILDING' compares a CHAR(10) attribute
        // with a VARCHAR(8) literal); its bytes carry the terminating NUL.  On the device a string literal is its bytes:
        // the comparison pads or cuts against the attribute's width like the reference's mixed CHAR / VARCHAR comparators
        // (types/operations/comparisons/AsciiStringComparators.hpp:218-251).
        std::string bytes = v.out_of_line_data();
        while (!bytes.empty() && bytes.back() == '\0') bytes.pop_back();
        n.type = QS_CHAR;
        n.width = static_cast<std::uint16_t>(bytes.size());
        n.lit.pool_offset = b->addString(bytes, n.width);
        return b->add(n);
      }
      n.type = LowerTypeID(lit_type, &w);
      n.width = w;
      switch (n.type) {
        case QS_INT: n.lit.i32 = v.int_value(); break;
        case QS_LONG: n.lit.i64 = v.long_value(); break;
        case QS_FLOAT: n.lit.f32 = v.float_value(); break;
        case QS_DOUBLE: n.lit.f64 = v.double_value(); break;
        case QS_DATE:
          n.lit.date.year = v.date_value().year();
          n.lit.date.month = static_cast<std::uint8_t>(v.date_value().month());
          n.lit.date.day = static_cast<std::uint8_t>(v.date_value().day());
          break;
        default: {   // CHAR: bytes travel in the string pool, NUL padded to the type's width
          const std::string &bytes = v.out_of_line_data();
          if (w == 0) n.width = static_cast<std::uint16_t>(bytes.size());
          n.lit.pool_offset = b->addString(bytes, n.width);
        }
      }
      return b->add(n);
    }
    case serialization::Scalar::ATTRIBUTE: {
      const qs_attr a = types.lookup(s.GetExtension(serialization::ScalarAttribute::relation_id),
                                     s.GetExtension(serialization::ScalarAttribute::attribute_id));
      if (a.type == QS_VARCHAR) LOG(FATAL) << "GPU path: VARCHAR attributes are not staged on the device";
      n.kind = QS_N_ATTRIBUTE;
      n.type = a.type;
      n.width = a.width;
      n.a = s.GetExtension(serialization::ScalarAttribute::attribute_id);
      n.b = types.single_relation ? 0 : static_cast<std::int32_t>(s.GetExtension(serialization::ScalarAttribute::join_side));   // RIGHT_SIDE = 2 = build side
      return b->add(n);
    }
    case serialization::Scalar::UNARY_EXPRESSION: {
      {
        TypedValue folded;
        const Type *folded_type = nullptr;
        if (FoldStaticScalar(s, &folded, &folded_type)) return LowerTypedValue(folded, *folded_type, b);
      }
      const serialization::UnaryOperation &op = s.GetExtension(serialization::ScalarUnaryExpression::operation);
      const int operand = LowerScalar(s.GetExtension(serialization::ScalarUnaryExpression::operand), types, b);
      n.kind = QS_N_UNARY;
      n.a = operand;
      n.type = b->node(operand).type;
      if (op.operation_id() == serialization::UnaryOperation::NEGATE) {
        n.op = QS_NEGATE;
      } else if (op.operation_id() == serialization::UnaryOperation::CAST) {
        std::uint16_t w = 0;
        n.op = QS_CAST;
        n.type = LowerTypeID(op.GetExtension(serialization::CastOperation::target_type), &w);
      } else {
        LOG(FATAL) << "GPU path: DATE_EXTRACT / SUBSTRING are not lowered";
      }
      return b->add(n);
    }
    case serialization::Scalar::BINARY_EXPRESSION: {
      {
        TypedValue folded;
        const Type *folded_type = nullptr;
        if (FoldStaticScalar(s, &folded, &folded_type)) return LowerTypedValue(folded, *folded_type, b);
      }
      const int l = LowerScalar(s.GetExtension(serialization::ScalarBinaryExpression::left_operand), types, b);
      const int r = LowerScalar(s.GetExtension(serialization::ScalarBinaryExpression::right_operand), types, b);
      n.kind = QS_N_BINARY;
      // BinaryOperationID: ADD, SUBTRACT, MULTIPLY, DIVIDE, MODULO = 0..4 in both enums
      n.op = static_cast<std::uint16_t>(s.GetExtension(serialization::ScalarBinaryExpression::operation).operation_id());
      n.a = l;
      n.b = r;
      // result type: C++ arithmetic promotion of the operand types (ArithmeticBinaryOperators.hpp:74-78); the library
      // recomputes it, the field only has to name a numeric type
      n.type = std::max(b->node(l).type, b->node(r).type);
      return b->add(n);
    }
    case serialization::Scalar::SHARED_EXPRESSION: {
      const int operand = LowerScalar(s.GetExtension(serialization::ScalarSharedExpression::operand), types, b);
      n.kind = QS_N_SHARED;
      n.a = operand;
      n.b = s.GetExtension(serialization::ScalarSharedExpression::share_id);
      n.type = b->node(operand).type;
      return b->add(n);
    }
    default:
      LOG(FATAL) << "GPU path: CASE expressions are not lowered";
  }
  return -1;
}

inline int LowerPredicate(const serialization::Predicate &p, const AttributeTypes &types, ExprBuilder *b) {
  qs_node n{};
  switch (p.predicate_type()) {
    case serialization::Predicate::TRUE: n.kind = QS_N_TRUE; return b->add(n);
    case serialization::Predicate::FALSE: n.kind = QS_N_FALSE; return b->add(n);
    case serialization::Predicate::COMPARISON: {
      const auto id = p.GetExtension(serialization::ComparisonPredicate::comparison).comparison_id();
      if (id > serialization::Comparison::GREATER_OR_EQUAL) LOG(FATAL) << "GPU path: LIKE / REGEX comparisons stay on the CPU path";
      const int l = LowerScalar(p.GetExtension(serialization::ComparisonPredicate::left_operand), types, b);
      const int r = LowerScalar(p.GetExtension(serialization::ComparisonPredicate::right_operand), types, b);
      n.kind = QS_N_COMPARISON;
      n.op = static_cast<std::uint16_t>(id);      // EQUAL..GREATER_OR_EQUAL = 0..5 = QS_EQ..QS_GE (ComparisonID.hpp)
      n.a = l;
      n.b = r;
      return b->add(n);
    }
    case serialization::Predicate::NEGATION:
      n.kind = QS_N_NEGATION;
      n.a = LowerPredicate(p.GetExtension(serialization::NegationPredicate::operand), types, b);
      return b->add(n);
    case serialization::Predicate::CONJUNCTION:
    case serialization::Predicate::DISJUNCTION: {
      // n-ary operand lists fold into left-deep binary chains
      const int count = p.ExtensionSize(serialization::PredicateWithList::operands);
      CHECK_GT(count, 0);
      int acc = LowerPredicate(p.GetExtension(serialization::PredicateWithList::operands, 0), types, b);
      for (int i = 1; i < count; ++i) {
        qs_node c{};
        c.kind = p.predicate_type() == serialization::Predicate::CONJUNCTION ? QS_N_CONJUNCTION : QS_N_DISJUNCTION;
        c.a = acc;
        c.b = LowerPredicate(p.GetExtension(serialization::PredicateWithList::operands, i), types, b);
        acc = b->add(c);
      }
      return acc;
    }
  }
  return -1;
}

}  // namespace gpu
}  // namespace quickstep
