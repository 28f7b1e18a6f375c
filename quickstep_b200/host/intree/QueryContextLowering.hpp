// IN-TREE binding, host-only part: the reference's serialized QueryContext entries -> the C ABI's descriptions.
// Nothing here touches the device, so it is not only type-checked against the reference's headers
// (tests/test_intree_boundary.py) but also EXECUTED against the reference's real optimizer output:
// tests/golden/make_plan_golden.cpp parses the TPC-H queries with the reference's parser, plans them with its optimizer
// and ExecutionGenerator, and lowers the serialization::QueryContext they produce with the functions below.
#pragma once

#include <cstddef>
#include <cstdint>
#include <vector>

#include "catalog/CatalogAttribute.hpp"
#include "catalog/CatalogRelationSchema.hpp"
#include "catalog/CatalogTypedefs.hpp"
#include "expressions/aggregation/AggregateFunction.pb.h"
#include "storage/AggregationOperationState.pb.h"
#include "storage/HashTable.pb.h"
#include "utility/SortConfiguration.pb.h"
#include "types/Type.hpp"
#include "types/TypeID.hpp"

#include "glog/logging.h"

#include "ProtoLowering.hpp"
#include "qsgpu.h"

namespace quickstep {
namespace gpu {

inline std::vector<qs_attr> AttributesOf(const CatalogRelationSchema &relation) {
  std::vector<qs_attr> out;
  for (CatalogRelationSchema::const_iterator it = relation.begin(); it != relation.end(); ++it) {
    const Type &t = it->getType();
    qs_attr a{};
    a.type = static_cast<std::uint16_t>(t.getTypeID() == kDate ? QS_DATE : static_cast<int>(t.getTypeID()));
    a.width = static_cast<std::uint16_t>(t.isVariableLength() ? 0 : t.maximumByteLength());
    out.push_back(a);
  }
  return out;
}

// A qs_agg_spec with the storage its pointers refer to.
struct LoweredAggregationState {
  ExprBuilder exprs;
  int predicate_root = -1;
  std::vector<qs_aggregate> aggregates;
  std::vector<std::int32_t> group_by_roots;
  std::uint32_t strategy = QS_AGG_SINGLE_STATE;
  std::uint64_t estimated_num_entries = 0;
  std::int64_t collision_free_max_key = -1;
  std::uint64_t nullable_arguments = 0;

  // `es` must be exprs.view() and outlive the returned spec
  qs_agg_spec spec(const int device, const qs_expr_set *es) const {
    qs_agg_spec s{};
    s.dev = device;
    s.strategy = strategy;
    s.exprs = es;
    s.predicate_root = predicate_root;
    s.n_aggregates = static_cast<std::uint32_t>(aggregates.size());
    s.aggregates = aggregates.data();
    s.n_group_by = static_cast<std::uint32_t>(group_by_roots.size());
    s.group_by_roots = group_by_roots.data();
    s.estimated_num_entries = estimated_num_entries;
    s.collision_free_max_key = collision_free_max_key;
    s.nullable_arguments = nullable_arguments;
    return s;
  }
};

// Is the scalar's type NULL-able, from the catalog types of the attributes it reads (Scalar::getType().isNullable() of the
// reconstructed expression: arithmetic over a NULL-able operand is NULL-able, ArithmeticBinaryOperators.hpp:178-186).
inline bool ScalarProtoIsNullable(const serialization::Scalar &s, const CatalogRelationSchema &input) {
  switch (s.data_source()) {
    case serialization::Scalar::ATTRIBUTE:
      return input.getAttributeById(s.GetExtension(serialization::ScalarAttribute::attribute_id))->getType().isNullable();
    case serialization::Scalar::UNARY_EXPRESSION:
      return ScalarProtoIsNullable(s.GetExtension(serialization::ScalarUnaryExpression::operand), input);
    case serialization::Scalar::BINARY_EXPRESSION:
      return ScalarProtoIsNullable(s.GetExtension(serialization::ScalarBinaryExpression::left_operand), input) ||
             ScalarProtoIsNullable(s.GetExtension(serialization::ScalarBinaryExpression::right_operand), input);
    case serialization::Scalar::SHARED_EXPRESSION:
      return ScalarProtoIsNullable(s.GetExtension(serialization::ScalarSharedExpression::operand), input);
    default:
      return false;
  }
}

// AggregationOperationState::ReconstructFromProto (storage/AggregationOperationState.cpp:186-260) as a qs_agg_spec: the
// aggregates (AggregationID + argument Scalar), the GROUP BY scalars, the predicate, estimated_num_entries and the
// hash-table implementation the optimizer chose (query_optimizer/ExecutionGenerator.cpp:1924-1965) -> strategy.
inline void LowerAggregationState(const serialization::AggregationOperationState &proto, const CatalogRelationSchema &input,
                                  LoweredAggregationState *out) {
  AttributeTypes types;
  types.relations.emplace_back(proto.relation_id(), AttributesOf(input));
  types.single_relation = true;
  out->predicate_root = proto.has_predicate() ? LowerPredicate(proto.predicate(), types, &out->exprs) : -1;
  for (int j = 0; j < proto.aggregates_size(); ++j) {
    const serialization::Aggregate &a = proto.aggregates(j);
    if (a.is_distinct()) LOG(FATAL) << "GPU path: DISTINCT aggregates keep their CPU operators";
    CHECK_LE(a.argument_size(), 1);
    qs_aggregate q{};
    q.function = static_cast<std::uint32_t>(a.function().aggregation_id());
    q.argument_root = a.argument_size() ? LowerScalar(a.argument(0), types, &out->exprs) : -1;
    if (a.argument_size() && j < 64 && ScalarProtoIsNullable(a.argument(0), input)) out->nullable_arguments |= 1ull << j;
    out->aggregates.push_back(q);
  }
  for (int g = 0; g < proto.group_by_expressions_size(); ++g)
    out->group_by_roots.push_back(LowerScalar(proto.group_by_expressions(g), types, &out->exprs));
  out->estimated_num_entries = proto.estimated_num_entries();
  if (out->group_by_roots.empty()) {
    out->strategy = QS_AGG_SINGLE_STATE;
  } else {
    switch (proto.hash_table_impl_type()) {
      case serialization::HashTableImplType::THREAD_PRIVATE_COMPACT_KEY: out->strategy = QS_AGG_COMPACT_KEY; break;
      case serialization::HashTableImplType::COLLISION_FREE_VECTOR:
        out->strategy = QS_AGG_COLLISION_FREE;
        out->collision_free_max_key = static_cast<std::int64_t>(proto.estimated_num_entries()) - 1;
        break;
      default: out->strategy = QS_AGG_SEPARATE_CHAINING;
    }
  }
}

// SortConfiguration (utility/SortConfiguration.proto; reconstructed at query_execution/QueryContext.cpp:128-131) as the
// qs_sort_key list of qsgpu_topk: ORDER BY expressions are attributes of the sorted relation (the optimizer projects
// anything else first); `null_first` is always spelled out in the proto -- the parser has already applied the default
// (NULLs first iff descending, parser/ParseOrderBy.hpp:53-66) -- so it travels as an explicit NULLS FIRST / LAST.
inline std::vector<qs_sort_key> LowerSortConfiguration(const serialization::SortConfiguration &proto) {
  std::vector<qs_sort_key> keys;
  for (int k = 0; k < proto.order_by_list_size(); ++k) {
    const serialization::SortConfiguration::OrderBy &ob = proto.order_by_list(k);
    if (ob.expression().data_source() != serialization::Scalar::ATTRIBUTE)
      LOG(FATAL) << "GPU path: ORDER BY expressions other than attributes keep their CPU operators";
    qs_sort_key key{};
    key.attr = static_cast<std::uint32_t>(ob.expression().GetExtension(serialization::ScalarAttribute::attribute_id));
    key.descending = (ob.is_ascending() ? 0u : QS_SORT_DESCENDING) | (ob.null_first() ? QS_SORT_NULLS_FIRST : QS_SORT_NULLS_LAST);
    keys.push_back(key);
  }
  if (keys.size() > 4) LOG(FATAL) << "GPU path: more than four sort attributes";
  return keys;
}

}  // namespace gpu
}  // namespace quickstep
