// IN-TREE binding, part 2: GPU work orders and operators as subclasses of the REFERENCE's own
// quickstep::WorkOrder / quickstep::RelationalOperator -- the file a maintainer drops into relational_operators/.
//
// tests/test_intree_boundary.py type-checks it (g++ -fsyntax-only) against /root/reference's headers plus the
// generated headers of oracle/build_ref.sh's build tree: every `override` below is checked against the real virtual
// (relational_operators/WorkOrder.hpp:251, relational_operators/RelationalOperator.hpp:101-196), every QueryContext /
// CatalogRelation / Predicate call against the real class.  It is not linked into libqshost.so: outside the
// reference's build the same operators run over the stand-in types of quickstep_b200/host/.
#include <cstddef>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "catalog/CatalogAttribute.hpp"
#include "catalog/CatalogRelation.hpp"
#include "catalog/CatalogTypedefs.hpp"
#include "expressions/predicate/Predicate.hpp"
#include "expressions/scalar/Scalar.hpp"
#include "query_execution/QueryContext.hpp"
#include "query_execution/WorkOrderProtosContainer.hpp"
#include "query_execution/WorkOrdersContainer.hpp"
#include "relational_operators/RelationalOperator.hpp"
#include "relational_operators/WorkOrder.hpp"
#include "storage/StorageBlockInfo.hpp"
#include "storage/StorageManager.hpp"
#include "types/Type.hpp"
#include "types/TypeID.hpp"

#include "glog/logging.h"
#include "tmb/id_typedefs.h"

#include "ProtoLowering.hpp"
#include "qsgpu.h"

namespace tmb { class MessageBus; }

namespace quickstep {
namespace gpu {

// The reference's error convention on this path: LOG(FATAL) (relational_operators/BuildHashOperator.cpp:205-206).
#define QS_GPU_CHECK(call)                                                                         \
  do {                                                                                             \
    const int st__ = (call);                                                                       \
    if (st__ != 0) LOG(FATAL) << #call << " failed with status " << st__ << ": " << qsgpu_last_error(); \
  } while (0)

// types/TypeID.hpp's enum values ARE the C ABI's type ids.
static_assert(static_cast<int>(kInt) == QS_INT && static_cast<int>(kLong) == QS_LONG && static_cast<int>(kFloat) == QS_FLOAT &&
                  static_cast<int>(kDouble) == QS_DOUBLE && static_cast<int>(kChar) == QS_CHAR && static_cast<int>(kVarChar) == QS_VARCHAR,
              "qsgpu_types.h must keep types/TypeID.hpp's values");

inline std::vector<qs_attr> SchemaOf(const CatalogRelationSchema &relation) {
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

// NULL-able attributes of a relation, as qsgpu_relation_set_nullable takes them (bit a = attribute a): the reference
// types every attribute as nullable or not (types/Type.hpp:129) and picks its NULL-checking code paths from that.
inline std::uint64_t NullableMaskOf(const CatalogRelationSchema &relation) {
  std::uint64_t mask = 0;
  for (CatalogRelationSchema::const_iterator it = relation.begin(); it != relation.end(); ++it)
    if (it->getType().isNullable() && it->getID() < 64) mask |= 1ull << it->getID();
  return mask;
}

// qs_agg_spec.nullable_arguments: bit j = aggregate j's argument has a NULL-able type.  `arguments` is what
// AggregationOperationState's constructor receives (storage/AggregationOperationState.hpp:109-121); the reference hands
// the same types to AggregateFunction::createHandle, which is where its handles learn whether to check for NULL.
inline std::uint64_t NullableArgumentsOf(const std::vector<std::vector<std::unique_ptr<const Scalar>>> &arguments) {
  std::uint64_t mask = 0;
  for (std::size_t j = 0; j < arguments.size() && j < 64; ++j)
    if (!arguments[j].empty() && arguments[j].front()->getType().isNullable()) mask |= 1ull << j;
  return mask;
}

// The device image of the input relation's blocks this work order scans (the GPU twin of a BlockReference).
struct DeviceRows {
  qsgpu_relation_t relation = nullptr;
  std::uint64_t row_begin = 0, row_end = UINT64_MAX;
};

// AggregationWorkOrder (relational_operators/AggregationOperator.hpp:151-204) on the device.
class GpuAggregationWorkOrder : public WorkOrder {
 public:
  GpuAggregationWorkOrder(const std::size_t query_id, const partition_id part_id, const DeviceRows &input,
                          qsgpu_agg_state_t state)
      : WorkOrder(query_id, part_id), input_(input), state_(state) {}
  ~GpuAggregationWorkOrder() override {}

  void execute() override {      // runs on a Worker thread (query_execution/Worker.cpp:127-139)
    QS_GPU_CHECK(qsgpu_agg_run(state_, input_.relation, input_.row_begin, input_.row_end, 0, nullptr));
  }

 private:
  const DeviceRows input_;
  qsgpu_agg_state_t state_;
};

// SelectWorkOrder (relational_operators/SelectOperator.hpp:262-383) on the device: the predicate and the projected
// scalars are the reference's own objects, lowered through their protos.
class GpuSelectWorkOrder : public WorkOrder {
 public:
  GpuSelectWorkOrder(const std::size_t query_id, const CatalogRelationSchema &input_relation, const DeviceRows &input,
                     const Predicate *predicate, const std::vector<std::unique_ptr<const Scalar>> *selection,
                     qsgpu_relation_t output)
      : WorkOrder(query_id), input_relation_(input_relation), input_(input), predicate_(predicate), selection_(selection),
        output_(output) {}
  // simple projection (SelectOperator.hpp:149-187): attribute ids instead of a scalar group
  GpuSelectWorkOrder(const std::size_t query_id, const CatalogRelationSchema &input_relation, const DeviceRows &input,
                     const Predicate *predicate, const std::vector<attribute_id> &simple_selection, qsgpu_relation_t output)
      : WorkOrder(query_id), input_relation_(input_relation), input_(input), predicate_(predicate), selection_(nullptr),
        simple_selection_(simple_selection), output_(output) {}
  ~GpuSelectWorkOrder() override {}

  void execute() override {
    AttributeTypes types;
    types.relations.emplace_back(input_relation_.getID(), SchemaOf(input_relation_));
    types.single_relation = true;
    ExprBuilder b;
    const int pred = predicate_ ? LowerPredicate(predicate_->getProto(), types, &b) : -1;
    std::vector<std::int32_t> roots;
    if (selection_) {
      for (const std::unique_ptr<const Scalar> &s : *selection_) roots.push_back(LowerScalar(s->getProto(), types, &b));
    } else {
      const std::vector<qs_attr> schema = SchemaOf(input_relation_);
      for (const attribute_id a : simple_selection_) {      // a bare attribute node is the selectSimple path (StorageBlock.cpp:390-398)
        qs_node n{};
        n.kind = QS_N_ATTRIBUTE;
        n.type = schema[static_cast<std::size_t>(a)].type;
        n.width = schema[static_cast<std::size_t>(a)].width;
        n.a = a;
        roots.push_back(b.add(n));
      }
    }
    const qs_expr_set es = b.view();
    qs_scan scan{};
    scan.input = input_.relation;
    scan.row_begin = input_.row_begin;
    scan.row_end = input_.row_end;
    scan.exprs = &es;
    scan.predicate_root = pred;
    QS_GPU_CHECK(qsgpu_select(&scan, static_cast<std::uint32_t>(roots.size()), roots.data(), output_));
  }

 private:
  const CatalogRelationSchema &input_relation_;
  const DeviceRows input_;
  const Predicate *predicate_;
  const std::vector<std::unique_ptr<const Scalar>> *selection_;
  const std::vector<attribute_id> simple_selection_;
  qsgpu_relation_t output_;
};

// AggregationOperator (relational_operators/AggregationOperator.hpp:60-149) emitting ONE coarse GPU work order per
// run of blocks instead of one per 4 MB block; same constructor arguments, same streaming contract.
class GpuAggregationOperator : public RelationalOperator {
 public:
  GpuAggregationOperator(const std::size_t query_id, const CatalogRelation &input_relation, bool input_relation_is_stored,
                         const QueryContext::aggregation_state_id aggr_state_index, const std::size_t num_partitions,
                         qsgpu_agg_state_t device_state, qsgpu_relation_t device_input)
      : RelationalOperator(query_id, num_partitions), input_relation_(input_relation),
        input_relation_is_stored_(input_relation_is_stored), aggr_state_index_(aggr_state_index), device_state_(device_state),
        device_input_(device_input), started_(false) {}
  ~GpuAggregationOperator() override {}

  OperatorType getOperatorType() const override { return kAggregation; }
  std::string getName() const override { return "GpuAggregationOperator"; }

  bool getAllWorkOrders(WorkOrdersContainer *container, QueryContext *query_context, StorageManager *storage_manager,
                        const tmb::client_id scheduler_client_id, tmb::MessageBus *bus) override {
    // Foreman thread only (query_execution/QueryManagerSingleNode.cpp:100-104); may be called many times
    if (input_relation_is_stored_) {
      if (!started_) {
        DeviceRows all;
        all.relation = device_input_;
        container->addNormalWorkOrder(new GpuAggregationWorkOrder(query_id_, 0, all, device_state_), op_index_);
        started_ = true;
      }
      return true;
    }
    // streamed input: one work order per batch of blocks fed so far (their rows are contiguous in the device image of the
    // temporary relation; noteDeviceRows() is told how many rows that image holds, and the last batch runs to its end)
    if (!fed_.empty()) {
      DeviceRows rows;
      rows.relation = device_input_;
      rows.row_begin = fed_row_begin_;
      rows.row_end = done_feeding_input_relation_ ? UINT64_MAX : fed_row_end_;
      container->addNormalWorkOrder(new GpuAggregationWorkOrder(query_id_, 0, rows, device_state_), op_index_);
      fed_.clear();
      fed_row_begin_ = fed_row_end_;
    }
    return done_feeding_input_relation_;
  }
  void noteDeviceRows(const std::uint64_t rows_so_far) { fed_row_end_ = rows_so_far; }

  bool getAllWorkOrderProtos(WorkOrderProtosContainer *container) override {
    LOG(FATAL) << "GPU work orders are single-node: -DENABLE_DISTRIBUTED builds keep the CPU operators";
    return true;
  }

  void feedInputBlock(const block_id input_block_id, const relation_id input_relation_id, const partition_id part_id) override {
    if (input_relation_id == input_relation_.getID()) fed_.push_back(input_block_id);
  }

 private:
  const CatalogRelation &input_relation_;
  const bool input_relation_is_stored_;
  const QueryContext::aggregation_state_id aggr_state_index_;
  qsgpu_agg_state_t device_state_;
  qsgpu_relation_t device_input_;
  bool started_;
  std::vector<block_id> fed_;
  std::uint64_t fed_row_begin_ = 0, fed_row_end_ = 0;
};

// SelectOperator (relational_operators/SelectOperator.hpp:66-260), both constructors: a scalar group or a simple
// projection.  Stored input: ONE coarse work order over the relation's device image; streamed input: one per batch of
// blocks fed since the last call (their rows are contiguous in the device image of the temporary relation).
class GpuSelectOperator : public RelationalOperator {
 public:
  GpuSelectOperator(const std::size_t query_id, const CatalogRelation &input_relation, const bool has_repartition,
                    const CatalogRelation &output_relation, const QueryContext::insert_destination_id output_destination_index,
                    const QueryContext::predicate_id predicate_index, const QueryContext::scalar_group_id selection_index,
                    const bool input_relation_is_stored, qsgpu_relation_t device_input, qsgpu_relation_t device_output)
      : RelationalOperator(query_id, input_relation.getNumPartitions(), has_repartition, output_relation.getNumPartitions()),
        input_relation_(input_relation), output_relation_(output_relation), output_destination_index_(output_destination_index),
        predicate_index_(predicate_index), selection_index_(selection_index), simple_projection_(false),
        input_relation_is_stored_(input_relation_is_stored), device_input_(device_input), device_output_(device_output),
        started_(false) {}
  GpuSelectOperator(const std::size_t query_id, const CatalogRelation &input_relation, const bool has_repartition,
                    const CatalogRelation &output_relation, const QueryContext::insert_destination_id output_destination_index,
                    const QueryContext::predicate_id predicate_index, std::vector<attribute_id> &&selection,
                    const bool input_relation_is_stored, qsgpu_relation_t device_input, qsgpu_relation_t device_output)
      : RelationalOperator(query_id, input_relation.getNumPartitions(), has_repartition, output_relation.getNumPartitions()),
        input_relation_(input_relation), output_relation_(output_relation), output_destination_index_(output_destination_index),
        predicate_index_(predicate_index), selection_index_(QueryContext::kInvalidScalarGroupId),
        simple_selection_(std::move(selection)), simple_projection_(true), input_relation_is_stored_(input_relation_is_stored),
        device_input_(device_input), device_output_(device_output), started_(false) {}
  ~GpuSelectOperator() override {}

  OperatorType getOperatorType() const override { return kSelect; }
  std::string getName() const override { return "GpuSelectOperator"; }

  bool getAllWorkOrders(WorkOrdersContainer *container, QueryContext *query_context, StorageManager *storage_manager,
                        const tmb::client_id scheduler_client_id, tmb::MessageBus *bus) override {
    const Predicate *predicate = query_context->getPredicate(predicate_index_);
    DeviceRows rows;
    rows.relation = device_input_;
    if (input_relation_is_stored_) {
      if (started_) return true;
      started_ = true;
    } else {
      if (fed_row_end_ == fed_row_begin_ && !fresh_blocks_) return done_feeding_input_relation_;
      rows.row_begin = fed_row_begin_;
      rows.row_end = done_feeding_input_relation_ ? UINT64_MAX : fed_row_end_;
      fed_row_begin_ = fed_row_end_;
      fresh_blocks_ = false;
    }
    container->addNormalWorkOrder(
        simple_projection_
            ? new GpuSelectWorkOrder(query_id_, input_relation_, rows, predicate, simple_selection_, device_output_)
            : new GpuSelectWorkOrder(query_id_, input_relation_, rows, predicate, &query_context->getScalarGroup(selection_index_),
                                     device_output_),
        op_index_);
    return input_relation_is_stored_ || done_feeding_input_relation_;
  }

  bool getAllWorkOrderProtos(WorkOrderProtosContainer *container) override {
    LOG(FATAL) << "GPU work orders are single-node: -DENABLE_DISTRIBUTED builds keep the CPU operators";
    return true;
  }

  // The producer's rows of this block are already in the temporary relation's device image; `rows_so_far` of that image
  // is what the binding's StorageManager hook reports (host/Operators.cpp InputFeed::take is the stand-in's version).
  void feedInputBlock(const block_id input_block_id, const relation_id input_relation_id, const partition_id part_id) override {
    if (input_relation_id != input_relation_.getID()) return;
    fresh_blocks_ = true;
  }
  void noteDeviceRows(const std::uint64_t rows_so_far) { fed_row_end_ = rows_so_far; }

  QueryContext::insert_destination_id getInsertDestinationID() const override { return output_destination_index_; }
  const relation_id getOutputRelationID() const override { return output_relation_.getID(); }

 private:
  const CatalogRelation &input_relation_;
  const CatalogRelation &output_relation_;
  const QueryContext::insert_destination_id output_destination_index_;
  const QueryContext::predicate_id predicate_index_;
  const QueryContext::scalar_group_id selection_index_;
  const std::vector<attribute_id> simple_selection_;
  const bool simple_projection_;
  const bool input_relation_is_stored_;
  qsgpu_relation_t device_input_, device_output_;
  bool started_;
  bool fresh_blocks_ = false;
  std::uint64_t fed_row_begin_ = 0, fed_row_end_ = 0;
};

}  // namespace gpu
}  // namespace quickstep
