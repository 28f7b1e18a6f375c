// IN-TREE binding, part 3: the remaining operators of the hot path as subclasses of the REFERENCE's own classes --
// BuildHash, HashJoin (inner / semi / anti / outer), BuildLIPFilter, FinalizeAggregation, the two Destroy work orders --
// plus the two translations they need: the reference's serialized QueryContext entries (AggregationOperationState,
// HashTable, LIPFilter, LIPFilterDeployment protos) into the C ABI's create calls, and a device result relation into the
// reference's own InsertDestination (ColumnVectorsValueAccessor + bulkInsertTuples).
//
// Like GpuWorkOrders.cpp it is type-checked by tests/test_intree_boundary.py (g++ -fsyntax-only) against
// /root/reference's headers and the generated *.pb.h of oracle/build_ref.sh's build tree; it is not linked into
// libqshost.so (outside the reference's build the same operators run over quickstep_b200/host/'s stand-ins).
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <memory>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "catalog/CatalogAttribute.hpp"
#include "catalog/CatalogRelation.hpp"
#include "catalog/CatalogRelationSchema.hpp"
#include "catalog/CatalogTypedefs.hpp"
#include "expressions/predicate/Predicate.hpp"
#include "expressions/scalar/Scalar.hpp"
#include "query_execution/QueryContext.hpp"
#include "query_execution/QueryContext.pb.h"
#include "query_execution/WorkOrderProtosContainer.hpp"
#include "query_execution/WorkOrdersContainer.hpp"
#include "relational_operators/HashJoinOperator.hpp"
#include "relational_operators/RelationalOperator.hpp"
#include "relational_operators/SortMergeRunOperator.hpp"
#include "relational_operators/WorkOrder.hpp"
#include "storage/AggregationOperationState.pb.h"
#include "storage/HashTable.pb.h"
#include "storage/InsertDestination.hpp"
#include "storage/StorageBlockInfo.hpp"
#include "storage/StorageManager.hpp"
#include "types/Type.hpp"
#include "types/TypeID.hpp"
#include "types/containers/ColumnVector.hpp"
#include "types/containers/ColumnVectorsValueAccessor.hpp"
#include "utility/lip_filter/LIPFilter.pb.h"

#include "glog/logging.h"
#include "tmb/id_typedefs.h"

#include "ProtoLowering.hpp"
#include "QueryContextLowering.hpp"
#include "qsgpu.h"

namespace tmb { class MessageBus; }

namespace quickstep {
namespace gpu {

#define QS_GPU_CHECK(call)                                                                         \
  do {                                                                                             \
    const int st__ = (call);                                                                       \
    if (st__ != 0) LOG(FATAL) << #call << " failed with status " << st__ << ": " << qsgpu_last_error(); \
  } while (0)

// expressions/aggregation/AggregateFunction.proto's AggregationID and relational_operators/HashJoinOperator.hpp's
// JoinType ARE the C ABI's ids.
static_assert(static_cast<int>(serialization::AggregateFunction::AVG) == QS_AGG_AVG &&
                  static_cast<int>(serialization::AggregateFunction::COUNT) == QS_AGG_COUNT &&
                  static_cast<int>(serialization::AggregateFunction::MAX) == QS_AGG_MAX &&
                  static_cast<int>(serialization::AggregateFunction::MIN) == QS_AGG_MIN &&
                  static_cast<int>(serialization::AggregateFunction::SUM) == QS_AGG_SUM,
              "qsgpu_types.h must keep AggregateFunction.proto's AggregationID values");
static_assert(static_cast<int>(HashJoinOperator::JoinType::kInnerJoin) == QS_JOIN_INNER &&
                  static_cast<int>(HashJoinOperator::JoinType::kLeftSemiJoin) == QS_JOIN_LEFT_SEMI &&
                  static_cast<int>(HashJoinOperator::JoinType::kLeftAntiJoin) == QS_JOIN_LEFT_ANTI &&
                  static_cast<int>(HashJoinOperator::JoinType::kLeftOuterJoin) == QS_JOIN_LEFT_OUTER,
              "qsgpu_types.h must keep HashJoinOperator::JoinType's values");

// ---------------------------------------------------------------------------------------------------------------------
// The device-side twins of what QueryContext owns, created from the SAME serialized entries and addressed by the SAME
// ids (query_execution/QueryContext.cpp:66-141 is where the reference reconstructs its own objects from them).
// ---------------------------------------------------------------------------------------------------------------------
class GpuQueryState {
 public:
  enum LipAction { kBuild, kProbe };

  GpuQueryState(const int device, const serialization::QueryContext &proto) : device_(device) {
    // LIPFilterFactory::ReconstructFromProto (utility/lip_filter/LIPFilterFactory.cpp:35-96)
    for (int i = 0; i < proto.lip_filters_size(); ++i) {
      const serialization::LIPFilter &f = proto.lip_filters(i);
      qsgpu_lip_t lip = nullptr;
      switch (f.lip_filter_type()) {
        case serialization::LIPFilterType::BIT_VECTOR_EXACT_FILTER: {
          const std::uint64_t size = f.GetExtension(serialization::BitVectorExactFilter::attribute_size);
          QS_GPU_CHECK(qsgpu_lip_create(device_, QS_LIP_BITVECTOR_EXACT, size == 4 ? QS_INT : QS_LONG,
                                        f.GetExtension(serialization::BitVectorExactFilter::min_value),
                                        f.GetExtension(serialization::BitVectorExactFilter::max_value), 0,
                                        f.GetExtension(serialization::BitVectorExactFilter::is_anti_filter) ? 1 : 0, &lip));
          break;
        }
        case serialization::LIPFilterType::SINGLE_IDENTITY_HASH_FILTER: {
          const std::uint64_t size = f.GetExtension(serialization::SingleIdentityHashFilter::attribute_size);
          QS_GPU_CHECK(qsgpu_lip_create(device_, QS_LIP_SINGLE_IDENTITY_HASH, size == 4 ? QS_INT : QS_LONG, 0, 0,
                                        f.GetExtension(serialization::SingleIdentityHashFilter::filter_cardinality), 0, &lip));
          break;
        }
        default:
          LOG(FATAL) << "GPU path: BLOOM_FILTER deployments keep their CPU operators";
      }
      lip_filters_.push_back(lip);
    }
    for (int i = 0; i < proto.lip_filter_deployments_size(); ++i) deployments_.push_back(proto.lip_filter_deployments(i));
    // JoinHashTable creation (query_execution/QueryContext.cpp:78-97): one table per (id, partition)
    for (int i = 0; i < proto.join_hash_tables_size(); ++i) {
      const serialization::QueryContext::HashTableContext &c = proto.join_hash_tables(i);
      const serialization::HashTable &h = c.join_hash_table();
      CHECK_GE(h.key_types_size(), 1);
      std::vector<qsgpu_join_table_t> parts;
      for (std::uint64_t p = 0; p < c.num_partitions(); ++p) {
        qsgpu_join_table_t t = nullptr;
        // a composite (two-INT) key is packed into one LONG by qsgpu_join_build_composite
        const std::uint32_t key_type =
            h.key_types_size() > 1 || h.key_types(0).type_id() == serialization::Type::LONG ? QS_LONG : QS_INT;
        QS_GPU_CHECK(qsgpu_join_create(device_, key_type, h.estimated_num_entries(), &t));
        parts.push_back(t);
      }
      join_tables_.push_back(std::move(parts));
    }
  }

  ~GpuQueryState() {
    for (auto &parts : join_tables_)
      for (qsgpu_join_table_t t : parts)
        if (t) qsgpu_join_destroy(t);
    for (auto &parts : agg_states_)
      for (qsgpu_agg_state_t s : parts.second)
        if (s) qsgpu_agg_destroy(s);
    for (qsgpu_lip_t l : lip_filters_)
      if (l) qsgpu_lip_destroy(l);
  }

  // AggregationOperationState::ReconstructFromProto (storage/AggregationOperationState.cpp:186-260) as a qs_agg_spec
  // (LowerAggregationState, QueryContextLowering.hpp).  `input` is the relation the proto's relation_id names.
  void addAggregationState(const QueryContext::aggregation_state_id id, const serialization::AggregationOperationState &proto,
                           const CatalogRelationSchema &input, const std::size_t num_partitions) {
    LoweredAggregationState lowered;
    LowerAggregationState(proto, input, &lowered);
    const qs_expr_set es = lowered.exprs.view();
    const qs_agg_spec spec = lowered.spec(device_, &es);
    std::vector<qsgpu_agg_state_t> parts;
    for (std::size_t p = 0; p < num_partitions; ++p) {
      qsgpu_agg_state_t s = nullptr;
      QS_GPU_CHECK(qsgpu_agg_create(&spec, &s));
      parts.push_back(s);
    }
    agg_states_[id] = std::move(parts);
  }

  qsgpu_agg_state_t aggregationState(const QueryContext::aggregation_state_id id, const partition_id part) const {
    return agg_states_.at(id).at(part);
  }
  void destroyAggregationState(const QueryContext::aggregation_state_id id, const partition_id part) {
    qsgpu_agg_state_t &s = agg_states_.at(id).at(part);
    if (s) QS_GPU_CHECK(qsgpu_agg_destroy(s));
    s = nullptr;
  }
  qsgpu_join_table_t joinHashTable(const QueryContext::join_hash_table_id id, const partition_id part) const {
    return join_tables_.at(id).at(part);
  }
  void destroyJoinHashTable(const QueryContext::join_hash_table_id id, const partition_id part) {
    qsgpu_join_table_t &t = join_tables_.at(id).at(part);
    if (t) QS_GPU_CHECK(qsgpu_join_destroy(t));
    t = nullptr;
  }

  // LIPFilterDeployment (utility/lip_filter/LIPFilterDeployment.cpp:39-87): the filters an operator builds / probes, each
  // bound to its attribute of the scanned relation.
  std::vector<qs_lip_ref> lipRefs(const QueryContext::lip_deployment_id id, const LipAction action) const {
    std::vector<qs_lip_ref> out;
    if (id == QueryContext::kInvalidLIPDeploymentId) return out;
    const serialization::LIPFilterDeployment &d = deployments_.at(id);
    const int n = action == kBuild ? d.build_entries_size() : d.probe_entries_size();
    for (int i = 0; i < n; ++i) {
      const serialization::LIPFilterDeployment::Entry &e = action == kBuild ? d.build_entries(i) : d.probe_entries(i);
      qs_lip_ref r{};
      r.lip = lip_filters_.at(e.lip_filter_id());
      r.attr = static_cast<std::uint32_t>(e.attribute_id());
      out.push_back(r);
    }
    return out;
  }

  int device() const { return device_; }

 private:
  const int device_;
  std::vector<qsgpu_lip_t> lip_filters_;
  std::vector<serialization::LIPFilterDeployment> deployments_;
  std::vector<std::vector<qsgpu_join_table_t>> join_tables_;
  std::unordered_map<QueryContext::aggregation_state_id, std::vector<qsgpu_agg_state_t>> agg_states_;
};

// A row range of a relation's device image (the GPU twin of a run of BlockReferences).
struct DeviceExtent {
  qsgpu_relation_t relation = nullptr;
  std::uint64_t row_begin = 0, row_end = UINT64_MAX;
};

// predicate (may be null) + scalars lowered into ONE node array, as a work order that evaluates both needs them
struct LoweredExprs {
  ExprBuilder builder;
  int predicate_root = -1;
  std::vector<std::int32_t> roots;
};

inline void LowerInto(LoweredExprs *out, const AttributeTypes &types, const Predicate *predicate,
                      const std::vector<std::unique_ptr<const Scalar>> *scalars) {
  if (predicate) out->predicate_root = LowerPredicate(predicate->getProto(), types, &out->builder);
  if (scalars)
    for (const std::unique_ptr<const Scalar> &s : *scalars) out->roots.push_back(LowerScalar(s->getProto(), types, &out->builder));
}

// ---------------------------------------------------------------------------------------------------------------------
// Rows that leave the device: a (small) device relation -> the reference's own InsertDestination, on the thread that runs
// the work order (it is the one registered in ClientIDMap, storage/InsertDestination.cpp:403-406).  ONE device-to-host copy
// (qsgpu_relation_read_rows), then NativeColumnVectors over the host columns, NULLs from the rows' masks.
// ---------------------------------------------------------------------------------------------------------------------
inline void EmitToInsertDestination(qsgpu_relation_t rows, const CatalogRelationSchema &schema, const std::uint64_t max_rows,
                                    InsertDestination *destination) {
  std::vector<const Type *> types;
  for (CatalogRelationSchema::const_iterator it = schema.begin(); it != schema.end(); ++it) types.push_back(&it->getType());
  std::vector<std::unique_ptr<char[]>> host;
  std::vector<void *> host_ptrs;
  for (const Type *t : types) {
    CHECK(!t->isVariableLength()) << "GPU path: variable-length result attributes keep their CPU operators";
    host.emplace_back(new char[t->maximumByteLength() * max_rows + 16]);
    host_ptrs.push_back(host.back().get());
  }
  std::vector<std::uint64_t> null_masks(max_rows);
  std::uint64_t n = 0;
  QS_GPU_CHECK(qsgpu_relation_read_rows(rows, max_rows, host_ptrs.data(), &n, null_masks.data()));
  CHECK_LE(n, max_rows) << "result larger than the bound its producer declared";
  ColumnVectorsValueAccessor accessor;                          // types/containers/ColumnVectorsValueAccessor.hpp:46-90
  for (std::size_t a = 0; a < types.size(); ++a) {
    NativeColumnVector *column = new NativeColumnVector(*types[a], n);
    const std::size_t width = types[a]->maximumByteLength();
    for (std::uint64_t r = 0; r < n; ++r) {
      if (types[a]->isNullable() && (null_masks[r] >> a & 1))
        column->appendNullValue();
      else
        column->appendUntypedValue(host[a].get() + r * width);
    }
    accessor.addColumn(column);                                 // takes ownership
  }
  destination->bulkInsertTuples(&accessor);                     // storage/InsertDestination.cpp:202-216
}

// ---------------------------------------------------------------------------------------------------------------------
// Work orders
// ---------------------------------------------------------------------------------------------------------------------

// BuildHashWorkOrder (relational_operators/BuildHashOperator.hpp:180-266, execute at BuildHashOperator.cpp:162-207):
// predicate -> LIP build -> put(key -> tuple reference).
class GpuBuildHashWorkOrder : public WorkOrder {
 public:
  GpuBuildHashWorkOrder(const std::size_t query_id, const CatalogRelationSchema &input_relation,
                        const std::vector<attribute_id> &join_key_attributes, const bool any_join_key_attributes_nullable,
                        const partition_id part_id, const DeviceExtent &input, const Predicate *predicate,
                        qsgpu_join_table_t hash_table, std::vector<qs_lip_ref> lip_build)
      : WorkOrder(query_id, part_id), input_relation_(input_relation), join_key_attributes_(join_key_attributes),
        any_join_key_attributes_nullable_(any_join_key_attributes_nullable), input_(input), predicate_(predicate),
        hash_table_(hash_table), lip_build_(std::move(lip_build)) {}
  ~GpuBuildHashWorkOrder() override {}

  void execute() override {
    AttributeTypes types;
    types.relations.emplace_back(input_relation_.getID(), AttributesOf(input_relation_));
    types.single_relation = true;
    LoweredExprs e;
    LowerInto(&e, types, predicate_, nullptr);
    const qs_expr_set es = e.builder.view();
    qs_scan scan{};
    scan.input = input_.relation;
    scan.row_begin = input_.row_begin;
    scan.row_end = input_.row_end;
    scan.exprs = &es;
    scan.predicate_root = e.predicate_root;
    // NULL keys never enter the table (storage/HashTable.hpp:1384): the library reads the relation's NULL-able set
    // (qsgpu_relation_set_nullable), any_join_key_attributes_nullable_ needs no flag of its own
    if (join_key_attributes_.size() == 1) {
      QS_GPU_CHECK(qsgpu_join_build(hash_table_, &scan, static_cast<std::uint32_t>(join_key_attributes_[0]),
                                    static_cast<std::uint32_t>(lip_build_.size()), lip_build_.data()));
    } else {
      std::vector<std::uint32_t> keys(join_key_attributes_.begin(), join_key_attributes_.end());
      QS_GPU_CHECK(qsgpu_join_build_composite(hash_table_, &scan, static_cast<std::uint32_t>(keys.size()), keys.data(),
                                              static_cast<std::uint32_t>(lip_build_.size()), lip_build_.data()));
    }
  }

 private:
  const CatalogRelationSchema &input_relation_;
  const std::vector<attribute_id> join_key_attributes_;
  const bool any_join_key_attributes_nullable_;
  const DeviceExtent input_;
  const Predicate *predicate_;
  qsgpu_join_table_t hash_table_;
  const std::vector<qs_lip_ref> lip_build_;
};

// HashInnerJoinWorkOrder / HashSemiJoinWorkOrder / HashAntiJoinWorkOrder / HashOuterJoinWorkOrder
// (relational_operators/HashJoinOperator.hpp:301-760, execute bodies at HashJoinOperator.cpp:450-1099): LIP probe -> hash
// probe -> residual predicate over both sides -> projection.  One class: the join type is an argument of the kernel.
class GpuHashJoinWorkOrder : public WorkOrder {
 public:
  GpuHashJoinWorkOrder(const std::size_t query_id, const CatalogRelationSchema &build_relation,
                       const CatalogRelationSchema &probe_relation, const std::vector<attribute_id> &join_key_attributes,
                       const bool any_join_key_attributes_nullable, const partition_id part_id, const DeviceExtent &probe,
                       const Predicate *residual_predicate, const std::vector<std::unique_ptr<const Scalar>> &selection,
                       const HashJoinOperator::JoinType join_type, qsgpu_join_table_t hash_table, qsgpu_relation_t output,
                       std::vector<qs_lip_ref> lip_probe)
      : WorkOrder(query_id, part_id), build_relation_(build_relation), probe_relation_(probe_relation),
        join_key_attributes_(join_key_attributes), any_join_key_attributes_nullable_(any_join_key_attributes_nullable),
        probe_(probe), residual_predicate_(residual_predicate), selection_(selection), join_type_(join_type),
        hash_table_(hash_table), output_(output), lip_probe_(std::move(lip_probe)) {}
  ~GpuHashJoinWorkOrder() override {}

  void execute() override {
    // ScalarAttribute protos carry (relation_id, attribute_id, join_side): both relations' types are needed, and
    // join_side == RIGHT_SIDE (2) becomes qs_node.b == 2 = "read through the matched build row"
    AttributeTypes types;
    types.relations.emplace_back(probe_relation_.getID(), AttributesOf(probe_relation_));
    types.relations.emplace_back(build_relation_.getID(), AttributesOf(build_relation_));
    LoweredExprs e;
    LowerInto(&e, types, residual_predicate_, &selection_);
    const qs_expr_set es = e.builder.view();
    qs_scan scan{};
    scan.input = probe_.relation;
    scan.row_begin = probe_.row_begin;
    scan.row_end = probe_.row_end;
    scan.exprs = &es;
    scan.predicate_root = -1;                 // a probe-side filter is its own SelectOperator upstream in the plans
    scan.n_lip_probe = static_cast<std::uint32_t>(lip_probe_.size());
    scan.lip_probe = lip_probe_.data();
    const std::uint32_t type = static_cast<std::uint32_t>(join_type_);
    if (join_key_attributes_.size() == 1) {
      QS_GPU_CHECK(qsgpu_join_probe(hash_table_, &scan, static_cast<std::uint32_t>(join_key_attributes_[0]), type,
                                    e.predicate_root, static_cast<std::uint32_t>(e.roots.size()), e.roots.data(), output_));
    } else {
      std::vector<std::uint32_t> keys(join_key_attributes_.begin(), join_key_attributes_.end());
      QS_GPU_CHECK(qsgpu_join_probe_composite(hash_table_, &scan, static_cast<std::uint32_t>(keys.size()), keys.data(), type,
                                              e.predicate_root, static_cast<std::uint32_t>(e.roots.size()), e.roots.data(),
                                              output_));
    }
  }

 private:
  const CatalogRelationSchema &build_relation_;
  const CatalogRelationSchema &probe_relation_;
  const std::vector<attribute_id> join_key_attributes_;
  const bool any_join_key_attributes_nullable_;
  const DeviceExtent probe_;
  const Predicate *residual_predicate_;
  const std::vector<std::unique_ptr<const Scalar>> &selection_;
  const HashJoinOperator::JoinType join_type_;
  qsgpu_join_table_t hash_table_;
  qsgpu_relation_t output_;
  const std::vector<qs_lip_ref> lip_probe_;
};

// BuildLIPFilterWorkOrder (relational_operators/BuildLIPFilterOperator.hpp:148-212, execute at
// BuildLIPFilterOperator.cpp:146-172): predicate -> probe the upstream filters -> insert the survivors.
class GpuBuildLIPFilterWorkOrder : public WorkOrder {
 public:
  GpuBuildLIPFilterWorkOrder(const std::size_t query_id, const CatalogRelationSchema &input_relation,
                             const partition_id part_id, const DeviceExtent &input, const Predicate *build_side_predicate,
                             std::vector<qs_lip_ref> lip_probe, std::vector<qs_lip_ref> lip_build)
      : WorkOrder(query_id, part_id), input_relation_(input_relation), input_(input), predicate_(build_side_predicate),
        lip_probe_(std::move(lip_probe)), lip_build_(std::move(lip_build)) {}
  ~GpuBuildLIPFilterWorkOrder() override {}

  void execute() override {
    AttributeTypes types;
    types.relations.emplace_back(input_relation_.getID(), AttributesOf(input_relation_));
    types.single_relation = true;
    LoweredExprs e;
    LowerInto(&e, types, predicate_, nullptr);
    const qs_expr_set es = e.builder.view();
    qs_scan scan{};
    scan.input = input_.relation;
    scan.row_begin = input_.row_begin;
    scan.row_end = input_.row_end;
    scan.exprs = &es;
    scan.predicate_root = e.predicate_root;
    scan.n_lip_probe = static_cast<std::uint32_t>(lip_probe_.size());
    scan.lip_probe = lip_probe_.data();
    QS_GPU_CHECK(qsgpu_build_lip_filter(&scan, static_cast<std::uint32_t>(lip_build_.size()), lip_build_.data()));
  }

 private:
  const CatalogRelationSchema &input_relation_;
  const DeviceExtent input_;
  const Predicate *predicate_;
  const std::vector<qs_lip_ref> lip_probe_, lip_build_;
};

// FinalizeAggregationWorkOrder (relational_operators/FinalizeAggregationOperator.hpp:124-166, execute at
// FinalizeAggregationOperator.cpp:99-101 -> AggregationOperationState::finalizeAggregate): states -> result tuples.
// `device_output` != nullptr: the consumer is another GPU operator and the rows stay in HBM; otherwise they are handed to
// the reference's InsertDestination.
class GpuFinalizeAggregationWorkOrder : public WorkOrder {
 public:
  GpuFinalizeAggregationWorkOrder(const std::size_t query_id, const partition_id part_id, qsgpu_agg_state_t state,
                                  const CatalogRelationSchema &output_relation, const std::uint64_t max_groups,
                                  InsertDestination *output_destination, qsgpu_relation_t *device_output)
      : WorkOrder(query_id, part_id), state_(state), output_relation_(output_relation), max_groups_(max_groups),
        output_destination_(output_destination), device_output_(device_output) {}
  ~GpuFinalizeAggregationWorkOrder() override {}

  void execute() override {
    qsgpu_relation_t rows = nullptr;
    QS_GPU_CHECK(qsgpu_agg_finalize(state_, &rows, nullptr));         // enqueue only: the group count stays on the device
    if (device_output_) {
      *device_output_ = rows;
      return;
    }
    EmitToInsertDestination(rows, output_relation_, max_groups_, output_destination_);
    QS_GPU_CHECK(qsgpu_relation_destroy(rows));
  }

 private:
  qsgpu_agg_state_t state_;
  const CatalogRelationSchema &output_relation_;
  const std::uint64_t max_groups_;
  InsertDestination *output_destination_;
  qsgpu_relation_t *device_output_;
};

// DestroyAggregationStateWorkOrder (relational_operators/DestroyAggregationStateOperator.cpp:69-71) and
// DestroyHashWorkOrder (relational_operators/DestroyHashOperator.cpp:70-72).
class GpuDestroyAggregationStateWorkOrder : public WorkOrder {
 public:
  GpuDestroyAggregationStateWorkOrder(const std::size_t query_id, const QueryContext::aggregation_state_id aggr_state_index,
                                      const partition_id part_id, GpuQueryState *state)
      : WorkOrder(query_id, part_id), aggr_state_index_(aggr_state_index), part_id_(part_id), state_(state) {}
  ~GpuDestroyAggregationStateWorkOrder() override {}
  void execute() override { state_->destroyAggregationState(aggr_state_index_, part_id_); }

 private:
  const QueryContext::aggregation_state_id aggr_state_index_;
  const partition_id part_id_;
  GpuQueryState *state_;
};

class GpuDestroyHashWorkOrder : public WorkOrder {
 public:
  GpuDestroyHashWorkOrder(const std::size_t query_id, const QueryContext::join_hash_table_id hash_table_index,
                          const partition_id part_id, GpuQueryState *state)
      : WorkOrder(query_id, part_id), hash_table_index_(hash_table_index), part_id_(part_id), state_(state) {}
  ~GpuDestroyHashWorkOrder() override {}
  void execute() override { state_->destroyJoinHashTable(hash_table_index_, part_id_); }

 private:
  const QueryContext::join_hash_table_id hash_table_index_;
  const partition_id part_id_;
  GpuQueryState *state_;
};

// ---------------------------------------------------------------------------------------------------------------------
// Operators.  Same constructor arguments as the reference's (what ExecutionGenerator passes), plus the device objects;
// one coarse work order per partition instead of one per 4 MB block.
// ---------------------------------------------------------------------------------------------------------------------

// BuildHashOperator (relational_operators/BuildHashOperator.hpp:66-178).
class GpuBuildHashOperator : public RelationalOperator {
 public:
  GpuBuildHashOperator(const std::size_t query_id, const CatalogRelation &input_relation, const bool input_relation_is_stored,
                       const std::vector<attribute_id> &join_key_attributes, const bool any_join_key_attributes_nullable,
                       const std::size_t num_partitions, const QueryContext::join_hash_table_id hash_table_index,
                       const QueryContext::predicate_id build_predicate_index, GpuQueryState *state,
                       std::vector<DeviceExtent> device_input /* one per partition */)
      : RelationalOperator(query_id, num_partitions), input_relation_(input_relation),
        input_relation_is_stored_(input_relation_is_stored), join_key_attributes_(join_key_attributes),
        any_join_key_attributes_nullable_(any_join_key_attributes_nullable), hash_table_index_(hash_table_index),
        build_predicate_index_(build_predicate_index), state_(state), device_input_(std::move(device_input)), started_(false) {}
  ~GpuBuildHashOperator() override {}

  OperatorType getOperatorType() const override { return kBuildHash; }
  std::string getName() const override { return "GpuBuildHashOperator"; }

  bool getAllWorkOrders(WorkOrdersContainer *container, QueryContext *query_context, StorageManager *storage_manager,
                        const tmb::client_id scheduler_client_id, tmb::MessageBus *bus) override {
    // a blocking consumer of its input in every plan of the hot path: the temporary relation it reads is complete (its
    // producer's kernels are queued on the same stream) once the input is stored or done feeding
    if (!input_relation_is_stored_ && !done_feeding_input_relation_) return false;
    if (!started_) {
      const Predicate *predicate = query_context->getPredicate(build_predicate_index_);
      for (partition_id part = 0; part < num_partitions_; ++part)
        container->addNormalWorkOrder(
            new GpuBuildHashWorkOrder(query_id_, input_relation_, join_key_attributes_, any_join_key_attributes_nullable_, part,
                                      device_input_.at(part), predicate, state_->joinHashTable(hash_table_index_, part),
                                      state_->lipRefs(lip_deployment_index_, GpuQueryState::kBuild)),
            op_index_);
      started_ = true;
    }
    return true;
  }

  bool getAllWorkOrderProtos(WorkOrderProtosContainer *container) override {
    LOG(FATAL) << "GPU work orders are single-node: -DENABLE_DISTRIBUTED builds keep the CPU operators";
    return true;
  }

  void feedInputBlock(const block_id input_block_id, const relation_id input_relation_id, const partition_id part_id) override {}

 private:
  const CatalogRelation &input_relation_;
  const bool input_relation_is_stored_;
  const std::vector<attribute_id> join_key_attributes_;
  const bool any_join_key_attributes_nullable_;
  const QueryContext::join_hash_table_id hash_table_index_;
  const QueryContext::predicate_id build_predicate_index_;
  GpuQueryState *state_;
  const std::vector<DeviceExtent> device_input_;
  bool started_;
};

// HashJoinOperator (relational_operators/HashJoinOperator.hpp:66-299), all four join types.
class GpuHashJoinOperator : public RelationalOperator {
 public:
  GpuHashJoinOperator(const std::size_t query_id, const CatalogRelation &build_relation, const CatalogRelation &probe_relation,
                      const bool probe_relation_is_stored, const std::vector<attribute_id> &join_key_attributes,
                      const bool any_join_key_attributes_nullable, const std::size_t num_partitions, const bool has_repartition,
                      const CatalogRelation &output_relation, const QueryContext::insert_destination_id output_destination_index,
                      const QueryContext::join_hash_table_id hash_table_index,
                      const QueryContext::predicate_id residual_predicate_index, const QueryContext::scalar_group_id selection_index,
                      const HashJoinOperator::JoinType join_type, GpuQueryState *state, std::vector<DeviceExtent> device_probe,
                      qsgpu_relation_t device_output)
      : RelationalOperator(query_id, num_partitions, has_repartition, output_relation.getNumPartitions()),
        build_relation_(build_relation), probe_relation_(probe_relation), probe_relation_is_stored_(probe_relation_is_stored),
        join_key_attributes_(join_key_attributes), any_join_key_attributes_nullable_(any_join_key_attributes_nullable),
        output_relation_(output_relation), output_destination_index_(output_destination_index),
        hash_table_index_(hash_table_index), residual_predicate_index_(residual_predicate_index),
        selection_index_(selection_index), join_type_(join_type), state_(state), device_probe_(std::move(device_probe)),
        device_output_(device_output), started_(false) {}
  ~GpuHashJoinOperator() override {}

  OperatorType getOperatorType() const override {
    switch (join_type_) {
      case HashJoinOperator::JoinType::kLeftSemiJoin: return kLeftSemiJoin;
      case HashJoinOperator::JoinType::kLeftAntiJoin: return kLeftAntiJoin;
      case HashJoinOperator::JoinType::kLeftOuterJoin: return kLeftOuterJoin;
      default: return kInnerJoin;
    }
  }
  std::string getName() const override { return "GpuHashJoinOperator"; }

  bool getAllWorkOrders(WorkOrdersContainer *container, QueryContext *query_context, StorageManager *storage_manager,
                        const tmb::client_id scheduler_client_id, tmb::MessageBus *bus) override {
    // the DAG's blocking edge from the BuildHashOperator holds this call back until the table is built
    // (query_execution/QueryManagerBase.cpp: blocking dependencies)
    if (!probe_relation_is_stored_ && !done_feeding_input_relation_) return false;
    if (!started_) {
      const Predicate *residual = query_context->getPredicate(residual_predicate_index_);
      const std::vector<std::unique_ptr<const Scalar>> &selection = query_context->getScalarGroup(selection_index_);
      for (partition_id part = 0; part < num_partitions_; ++part)
        container->addNormalWorkOrder(
            new GpuHashJoinWorkOrder(query_id_, build_relation_, probe_relation_, join_key_attributes_,
                                     any_join_key_attributes_nullable_, part, device_probe_.at(part), residual, selection, join_type_,
                                     state_->joinHashTable(hash_table_index_, part), device_output_,
                                     state_->lipRefs(lip_deployment_index_, GpuQueryState::kProbe)),
            op_index_);
      started_ = true;
    }
    return true;
  }

  bool getAllWorkOrderProtos(WorkOrderProtosContainer *container) override {
    LOG(FATAL) << "GPU work orders are single-node: -DENABLE_DISTRIBUTED builds keep the CPU operators";
    return true;
  }

  void feedInputBlock(const block_id input_block_id, const relation_id input_relation_id, const partition_id part_id) override {}

  void doneFeedingInputBlocks(const relation_id rel_id) override {
    // ignore the message that follows the completion of the BuildHashOperator (HashJoinOperator.hpp:240-247)
    if (probe_relation_.getID() == rel_id) done_feeding_input_relation_ = true;
  }

  QueryContext::insert_destination_id getInsertDestinationID() const override { return output_destination_index_; }
  const relation_id getOutputRelationID() const override { return output_relation_.getID(); }

 private:
  const CatalogRelation &build_relation_;
  const CatalogRelation &probe_relation_;
  const bool probe_relation_is_stored_;
  const std::vector<attribute_id> join_key_attributes_;
  const bool any_join_key_attributes_nullable_;
  const CatalogRelation &output_relation_;
  const QueryContext::insert_destination_id output_destination_index_;
  const QueryContext::join_hash_table_id hash_table_index_;
  const QueryContext::predicate_id residual_predicate_index_;
  const QueryContext::scalar_group_id selection_index_;
  const HashJoinOperator::JoinType join_type_;
  GpuQueryState *state_;
  const std::vector<DeviceExtent> device_probe_;
  qsgpu_relation_t device_output_;
  bool started_;
};

// BuildLIPFilterOperator (relational_operators/BuildLIPFilterOperator.hpp:62-146).
class GpuBuildLIPFilterOperator : public RelationalOperator {
 public:
  GpuBuildLIPFilterOperator(const std::size_t query_id, const CatalogRelation &input_relation,
                            const QueryContext::predicate_id build_side_predicate_index, const bool input_relation_is_stored,
                            GpuQueryState *state, std::vector<DeviceExtent> device_input)
      : RelationalOperator(query_id, input_relation.getNumPartitions()), input_relation_(input_relation),
        build_side_predicate_index_(build_side_predicate_index), input_relation_is_stored_(input_relation_is_stored),
        state_(state), device_input_(std::move(device_input)), started_(false) {}
  ~GpuBuildLIPFilterOperator() override {}

  OperatorType getOperatorType() const override { return kBuildLIPFilter; }
  std::string getName() const override { return "GpuBuildLIPFilterOperator"; }

  bool getAllWorkOrders(WorkOrdersContainer *container, QueryContext *query_context, StorageManager *storage_manager,
                        const tmb::client_id scheduler_client_id, tmb::MessageBus *bus) override {
    if (!input_relation_is_stored_ && !done_feeding_input_relation_) return false;
    if (!started_) {
      const Predicate *predicate = query_context->getPredicate(build_side_predicate_index_);
      for (partition_id part = 0; part < num_partitions_; ++part)
        container->addNormalWorkOrder(
            new GpuBuildLIPFilterWorkOrder(query_id_, input_relation_, part, device_input_.at(part), predicate,
                                           state_->lipRefs(lip_deployment_index_, GpuQueryState::kProbe),
                                           state_->lipRefs(lip_deployment_index_, GpuQueryState::kBuild)),
            op_index_);
      started_ = true;
    }
    return true;
  }

  bool getAllWorkOrderProtos(WorkOrderProtosContainer *container) override {
    LOG(FATAL) << "GPU work orders are single-node: -DENABLE_DISTRIBUTED builds keep the CPU operators";
    return true;
  }

  void feedInputBlock(const block_id input_block_id, const relation_id input_relation_id, const partition_id part_id) override {}

 private:
  const CatalogRelation &input_relation_;
  const QueryContext::predicate_id build_side_predicate_index_;
  const bool input_relation_is_stored_;
  GpuQueryState *state_;
  const std::vector<DeviceExtent> device_input_;
  bool started_;
};

// FinalizeAggregationOperator (relational_operators/FinalizeAggregationOperator.hpp:53-122): one work order per partition
// of the state, generated once (its blocking dependency, the AggregationOperator, has finished by then).
class GpuFinalizeAggregationOperator : public RelationalOperator {
 public:
  GpuFinalizeAggregationOperator(const std::size_t query_id, const QueryContext::aggregation_state_id aggr_state_index,
                                 const std::size_t num_partitions, const bool has_repartition,
                                 const std::size_t aggr_state_num_partitions, const CatalogRelation &output_relation,
                                 const QueryContext::insert_destination_id output_destination_index, GpuQueryState *state,
                                 const std::uint64_t max_groups)
      : RelationalOperator(query_id, num_partitions, has_repartition, output_relation.getNumPartitions()),
        aggr_state_index_(aggr_state_index), aggr_state_num_partitions_(aggr_state_num_partitions),
        output_relation_(output_relation), output_destination_index_(output_destination_index), state_(state),
        max_groups_(max_groups), started_(false) {}
  ~GpuFinalizeAggregationOperator() override {}

  OperatorType getOperatorType() const override { return kFinalizeAggregation; }
  std::string getName() const override { return "GpuFinalizeAggregationOperator"; }

  bool getAllWorkOrders(WorkOrdersContainer *container, QueryContext *query_context, StorageManager *storage_manager,
                        const tmb::client_id scheduler_client_id, tmb::MessageBus *bus) override {
    if (!started_) {
      InsertDestination *destination = query_context->getInsertDestination(output_destination_index_);
      for (partition_id part = 0; part < num_partitions_; ++part)
        container->addNormalWorkOrder(
            new GpuFinalizeAggregationWorkOrder(query_id_, part, state_->aggregationState(aggr_state_index_, part),
                                                output_relation_, max_groups_, destination, nullptr),
            op_index_);
      started_ = true;
    }
    return true;
  }

  bool getAllWorkOrderProtos(WorkOrderProtosContainer *container) override {
    LOG(FATAL) << "GPU work orders are single-node: -DENABLE_DISTRIBUTED builds keep the CPU operators";
    return true;
  }

  QueryContext::insert_destination_id getInsertDestinationID() const override { return output_destination_index_; }
  const relation_id getOutputRelationID() const override { return output_relation_.getID(); }

 private:
  const QueryContext::aggregation_state_id aggr_state_index_;
  const std::size_t aggr_state_num_partitions_;
  const CatalogRelation &output_relation_;
  const QueryContext::insert_destination_id output_destination_index_;
  GpuQueryState *state_;
  const std::uint64_t max_groups_;
  bool started_;
};

// ---------------------------------------------------------------------------------------------------------------------
// ORDER BY [LIMIT]: the reference plans a SortRunGenerationOperator (sort every block into a run) followed by a
// SortMergeRunOperator (merge the runs, keep top_k; relational_operators/SortMergeRunOperator.hpp:72-240).  On the device
// the whole input is one relation: the GPU run generation has nothing to do, and the merge operator sorts it with ONE
// qsgpu_topk (limit = top_k, or every row when there is no LIMIT).
// ---------------------------------------------------------------------------------------------------------------------
class GpuTopKWorkOrder : public WorkOrder {
 public:
  GpuTopKWorkOrder(const std::size_t query_id, qsgpu_relation_t input, std::vector<qs_sort_key> keys, const std::uint64_t limit,
                   const CatalogRelationSchema &output_relation, InsertDestination *output_destination, qsgpu_relation_t *device_output)
      : WorkOrder(query_id), input_(input), keys_(std::move(keys)), limit_(limit), output_relation_(output_relation),
        output_destination_(output_destination), device_output_(device_output) {}
  ~GpuTopKWorkOrder() override {}

  void execute() override {
    qsgpu_relation_t sorted = nullptr;
    QS_GPU_CHECK(qsgpu_topk(input_, static_cast<std::uint32_t>(keys_.size()), keys_.data(), limit_, &sorted));
    if (device_output_) {
      *device_output_ = sorted;
      return;
    }
    EmitToInsertDestination(sorted, output_relation_, limit_, output_destination_);      // the query's result relation
    QS_GPU_CHECK(qsgpu_relation_destroy(sorted));
  }

 private:
  qsgpu_relation_t input_;
  const std::vector<qs_sort_key> keys_;
  const std::uint64_t limit_;
  const CatalogRelationSchema &output_relation_;
  InsertDestination *output_destination_;
  qsgpu_relation_t *device_output_;
};

// SortMergeRunOperator's constructor arguments (what ExecutionGenerator::convertSort passes, ExecutionGenerator.cpp:
// 2227-2351) plus the lowered sort keys and the device image of the input.
class GpuSortMergeRunOperator : public RelationalOperator {
 public:
  GpuSortMergeRunOperator(const std::size_t query_id, const CatalogRelation &input_relation, const CatalogRelation &output_relation,
                          const QueryContext::insert_destination_id output_destination_index, const CatalogRelation &run_relation,
                          const QueryContext::insert_destination_id run_block_destination_index,
                          const QueryContext::sort_config_id sort_config_index, const std::size_t merge_factor, const std::size_t top_k,
                          const bool input_relation_is_stored, std::vector<qs_sort_key> keys, qsgpu_relation_t device_input,
                          const std::uint64_t max_rows)
      : RelationalOperator(query_id, 1u, false, 1u), input_relation_(input_relation), output_relation_(output_relation),
        output_destination_index_(output_destination_index), sort_config_index_(sort_config_index), top_k_(top_k),
        input_relation_is_stored_(input_relation_is_stored), keys_(std::move(keys)), device_input_(device_input),
        max_rows_(max_rows), started_(false) {}
  ~GpuSortMergeRunOperator() override {}

  OperatorType getOperatorType() const override { return kSortMergeRun; }
  std::string getName() const override { return "GpuSortMergeRunOperator"; }

  bool getAllWorkOrders(WorkOrdersContainer *container, QueryContext *query_context, StorageManager *storage_manager,
                        const tmb::client_id scheduler_client_id, tmb::MessageBus *bus) override {
    // a sort consumes its whole input: nothing before the producer has finished
    if (!input_relation_is_stored_ && !done_feeding_input_relation_) return false;
    if (!started_) {
      container->addNormalWorkOrder(
          new GpuTopKWorkOrder(query_id_, device_input_, keys_, top_k_ ? top_k_ : max_rows_, output_relation_,
                               query_context->getInsertDestination(output_destination_index_), nullptr),
          op_index_);
      started_ = true;
    }
    return true;
  }

  bool getAllWorkOrderProtos(WorkOrderProtosContainer *container) override {
    LOG(FATAL) << "GPU work orders are single-node: -DENABLE_DISTRIBUTED builds keep the CPU operators";
    return true;
  }

  void feedInputBlock(const block_id input_block_id, const relation_id input_relation_id, const partition_id part_id) override {}

  QueryContext::insert_destination_id getInsertDestinationID() const override { return output_destination_index_; }
  const relation_id getOutputRelationID() const override { return output_relation_.getID(); }

 private:
  const CatalogRelation &input_relation_;
  const CatalogRelation &output_relation_;
  const QueryContext::insert_destination_id output_destination_index_;
  const QueryContext::sort_config_id sort_config_index_;
  const std::size_t top_k_;
  const bool input_relation_is_stored_;
  const std::vector<qs_sort_key> keys_;
  qsgpu_relation_t device_input_;
  const std::uint64_t max_rows_;
  bool started_;
};

}  // namespace gpu
}  // namespace quickstep
