// GPU twins of the six operators on the hot path (+ their state lifetime
// operators and the top-k sort), with the reference's class names, constructor
// argument order and meaning, and streaming contract:
//   SelectOperator               relational_operators/SelectOperator.hpp:69-260
//   BuildLIPFilterOperator       relational_operators/BuildLIPFilterOperator.hpp:62-158
//   BuildHashOperator            relational_operators/BuildHashOperator.hpp:66-182
//   HashJoinOperator             relational_operators/HashJoinOperator.hpp:66-299
//   AggregationOperator          relational_operators/AggregationOperator.hpp:60-149
//   InitializeAggregationOperator / FinalizeAggregationOperator /
//   DestroyAggregationStateOperator / DestroyHashOperator
//   SortMergeRunOperator (top_k) relational_operators/SortMergeRunOperator.hpp:72
//
// Granularity (SURVEY.md section 7 "hard parts"): the CPU operators emit one
// work order per 4 MB block; a B200 wants >= 100 MB per launch.  A GPU operator
// therefore emits ONE work order per run of blocks it currently knows about
// (all blocks of a stored relation; each fed block of a streamed one), never
// breaking the contracts: getAllWorkOrders may be called many times, blocks may
// arrive at any time through feedInputBlock, and the return value says whether
// more work orders can still come.  `gpu_rows_per_workorder` caps the rows of
// one work order (0 = no cap) so tests can force many work orders per operator.
#pragma once

#include <functional>
#include <string>
#include <vector>

#include "WorkOrder.hpp"

namespace quickstep {

extern std::uint64_t FLAGS_gpu_rows_per_workorder;

enum class JoinType { kInnerJoin = 0, kLeftSemiJoin, kLeftAntiJoin, kLeftOuterJoin };   // HashJoinOperator.hpp:82-87

// Turns "the blocks this operator knows and has not yet scheduled" into device extents.
class InputFeed {
 public:
  InputFeed(const CatalogRelation &rel, bool stored) : relation_(rel), stored_(stored) {}
  void feed(block_id b) { pending_.push_back(b); }
  // needed_attrs: the attributes of the input relation the operator's work orders will read
  std::vector<DeviceExtent> take(StorageManager *sm, std::uint64_t needed_attrs = ~0ull);
  // true once no further extent can appear
  bool exhausted(bool done_feeding) const { return stored_ ? started_ : (done_feeding && pending_.empty()); }
  const CatalogRelation &relation() const { return relation_; }

 private:
  const CatalogRelation &relation_;
  const bool stored_;
  bool started_ = false;
  std::vector<block_id> pending_;
};

// ------------------------------------------------------------------ Select
class SelectOperator : public RelationalOperator {
 public:
  SelectOperator(const std::size_t query_id, const CatalogRelation &input_relation, const bool has_repartition,
                 const CatalogRelation &output_relation,
                 const QueryContext::insert_destination_id output_destination_index,
                 const QueryContext::predicate_id predicate_index,
                 const QueryContext::scalar_group_id selection_index, const bool input_relation_is_stored)
      : RelationalOperator(query_id), feed_(input_relation, input_relation_is_stored), output_relation_(output_relation),
        output_destination_index_(output_destination_index), predicate_index_(predicate_index),
        selection_index_(selection_index), simple_projection_(false) {}
  // simple projection: a list of attribute ids (SelectOperator.hpp:120-160)
  SelectOperator(const std::size_t query_id, const CatalogRelation &input_relation, const bool has_repartition,
                 const CatalogRelation &output_relation,
                 const QueryContext::insert_destination_id output_destination_index,
                 const QueryContext::predicate_id predicate_index, std::vector<attribute_id> &&selection,
                 const bool input_relation_is_stored)
      : RelationalOperator(query_id), feed_(input_relation, input_relation_is_stored), output_relation_(output_relation),
        output_destination_index_(output_destination_index), predicate_index_(predicate_index),
        selection_index_(QueryContext::kInvalidScalarGroupId), simple_selection_(std::move(selection)),
        simple_projection_(true) {}
  OperatorType getOperatorType() const override { return kSelect; }
  std::string getName() const override { return "SelectOperator"; }
  bool getAllWorkOrders(WorkOrdersContainer *container, QueryContext *query_context, StorageManager *storage_manager,
                        const tmb::client_id scheduler_client_id, tmb::MessageBus *bus) override;
  void feedInputBlock(const block_id input_block_id, const relation_id, const partition_id) override { feed_.feed(input_block_id); }
  QueryContext::insert_destination_id getInsertDestinationID() const override { return output_destination_index_; }
  const relation_id getOutputRelationID() const override { return output_relation_.getID(); }

 private:
  InputFeed feed_;
  const CatalogRelation &output_relation_;
  const QueryContext::insert_destination_id output_destination_index_;
  const QueryContext::predicate_id predicate_index_;
  const QueryContext::scalar_group_id selection_index_;
  const std::vector<attribute_id> simple_selection_;
  const bool simple_projection_;
};

class SelectWorkOrder : public WorkOrder {
 public:
  SelectWorkOrder(const std::size_t query_id, const CatalogRelation &input_relation, const DeviceExtent &input,
                  const QueryContext::Predicate *predicate, const QueryContext::ScalarGroup *selection,
                  const std::vector<attribute_id> *simple_selection, InsertDestination *output_destination,
                  std::vector<qs_lip_ref> lip_probe)
      : WorkOrder(query_id), input_relation_(input_relation), input_(input), predicate_(predicate), selection_(selection),
        simple_selection_(simple_selection), output_destination_(output_destination), lip_probe_(std::move(lip_probe)) {}
  void execute() override;      // SelectWorkOrder::execute, SelectOperator.cpp:161-195

 private:
  const CatalogRelation &input_relation_;
  const DeviceExtent input_;
  const QueryContext::Predicate *predicate_;
  const QueryContext::ScalarGroup *selection_;
  const std::vector<attribute_id> *simple_selection_;
  InsertDestination *output_destination_;
  const std::vector<qs_lip_ref> lip_probe_;      // LIPFilterAdaptiveProber, owned by the work order
};

// ---------------------------------------------------------- BuildLIPFilter
class BuildLIPFilterOperator : public RelationalOperator {
 public:
  BuildLIPFilterOperator(const std::size_t query_id, const CatalogRelation &input_relation,
                         const QueryContext::predicate_id build_side_predicate_index,
                         const bool input_relation_is_stored)
      : RelationalOperator(query_id), feed_(input_relation, input_relation_is_stored),
        build_side_predicate_index_(build_side_predicate_index) {}
  OperatorType getOperatorType() const override { return kBuildLIPFilter; }
  std::string getName() const override { return "BuildLIPFilterOperator"; }
  bool getAllWorkOrders(WorkOrdersContainer *container, QueryContext *query_context, StorageManager *storage_manager,
                        const tmb::client_id scheduler_client_id, tmb::MessageBus *bus) override;
  void feedInputBlock(const block_id input_block_id, const relation_id, const partition_id) override { feed_.feed(input_block_id); }

 private:
  InputFeed feed_;
  const QueryContext::predicate_id build_side_predicate_index_;
};

class BuildLIPFilterWorkOrder : public WorkOrder {
 public:
  BuildLIPFilterWorkOrder(const std::size_t query_id, const DeviceExtent &input, const QueryContext::Predicate *build_side_predicate,
                          std::vector<qs_lip_ref> lip_probe, std::vector<qs_lip_ref> lip_build)
      : WorkOrder(query_id), input_(input), build_side_predicate_(build_side_predicate), lip_probe_(std::move(lip_probe)),
        lip_build_(std::move(lip_build)) {}
  void execute() override;      // BuildLIPFilterOperator.cpp:146-172

 private:
  const DeviceExtent input_;
  const QueryContext::Predicate *build_side_predicate_;
  const std::vector<qs_lip_ref> lip_probe_, lip_build_;
};

// --------------------------------------------------------------- BuildHash
class BuildHashOperator : public RelationalOperator {
 public:
  BuildHashOperator(const std::size_t query_id, const CatalogRelation &input_relation, const bool input_relation_is_stored,
                    const std::vector<attribute_id> &join_key_attributes, const bool any_join_key_attributes_nullable,
                    const std::size_t num_partitions, const QueryContext::join_hash_table_id hash_table_index,
                    const QueryContext::predicate_id build_predicate_index = QueryContext::kInvalidPredicateId)
      : RelationalOperator(query_id, num_partitions), feed_(input_relation, input_relation_is_stored),
        join_key_attributes_(join_key_attributes), hash_table_index_(hash_table_index),
        build_predicate_index_(build_predicate_index) {
    QS_CHECK(join_key_attributes.size() == 1u || join_key_attributes.size() == 2u);   // composite: two INT attributes
    QS_CHECK(!any_join_key_attributes_nullable);
  }
  OperatorType getOperatorType() const override { return kBuildHash; }
  std::string getName() const override { return "BuildHashOperator"; }
  bool getAllWorkOrders(WorkOrdersContainer *container, QueryContext *query_context, StorageManager *storage_manager,
                        const tmb::client_id scheduler_client_id, tmb::MessageBus *bus) override;
  void feedInputBlock(const block_id input_block_id, const relation_id, const partition_id) override { feed_.feed(input_block_id); }

 private:
  InputFeed feed_;
  const std::vector<attribute_id> join_key_attributes_;
  const QueryContext::join_hash_table_id hash_table_index_;
  const QueryContext::predicate_id build_predicate_index_;
};

class BuildHashWorkOrder : public WorkOrder {
 public:
  BuildHashWorkOrder(const std::size_t query_id, const DeviceExtent &input, const std::vector<attribute_id> &join_key_attributes,
                     const QueryContext::Predicate *predicate, qsgpu_join_table_t hash_table,
                     std::vector<qs_lip_ref> lip_probe, std::vector<qs_lip_ref> lip_build)
      : WorkOrder(query_id), input_(input), join_key_attributes_(join_key_attributes.begin(), join_key_attributes.end()), predicate_(predicate),
        hash_table_(hash_table), lip_probe_(std::move(lip_probe)), lip_build_(std::move(lip_build)) {}
  void execute() override;      // BuildHashOperator.cpp:162-207

 private:
  const DeviceExtent input_;
  const std::vector<std::uint32_t> join_key_attributes_;
  const QueryContext::Predicate *predicate_;
  qsgpu_join_table_t hash_table_;
  const std::vector<qs_lip_ref> lip_probe_, lip_build_;
};

// ---------------------------------------------------------------- HashJoin
class HashJoinOperator : public RelationalOperator {
 public:
  HashJoinOperator(const std::size_t query_id, const CatalogRelation &build_relation, const CatalogRelation &probe_relation,
                   const bool probe_relation_is_stored, const std::vector<attribute_id> &join_key_attributes,
                   const bool any_join_key_attributes_nullable, const std::size_t num_partitions, const bool has_repartition,
                   const CatalogRelation &output_relation, const QueryContext::insert_destination_id output_destination_index,
                   const QueryContext::join_hash_table_id hash_table_index,
                   const QueryContext::predicate_id residual_predicate_index,
                   const QueryContext::scalar_group_id selection_index,
                   const std::vector<bool> *is_selection_on_build = nullptr, const JoinType join_type = JoinType::kInnerJoin)
      : RelationalOperator(query_id, num_partitions), build_relation_(build_relation), feed_(probe_relation, probe_relation_is_stored),
        join_key_attributes_(join_key_attributes), output_relation_(output_relation),
        output_destination_index_(output_destination_index), hash_table_index_(hash_table_index),
        residual_predicate_index_(residual_predicate_index), selection_index_(selection_index), join_type_(join_type) {
    QS_CHECK(join_key_attributes.size() == 1u || join_key_attributes.size() == 2u);
    QS_CHECK(!any_join_key_attributes_nullable);
  }
  OperatorType getOperatorType() const override {
    switch (join_type_) {
      case JoinType::kInnerJoin: return kInnerJoin;
      case JoinType::kLeftSemiJoin: return kLeftSemiJoin;
      case JoinType::kLeftAntiJoin: return kLeftAntiJoin;
      default: return kLeftOuterJoin;
    }
  }
  std::string getName() const override { return "HashJoinOperator"; }
  bool getAllWorkOrders(WorkOrdersContainer *container, QueryContext *query_context, StorageManager *storage_manager,
                        const tmb::client_id scheduler_client_id, tmb::MessageBus *bus) override;
  void feedInputBlock(const block_id input_block_id, const relation_id input_relation_id, const partition_id) override {
    if (input_relation_id == feed_.relation().getID()) feed_.feed(input_block_id);
  }
  void doneFeedingInputBlocks(const relation_id rel_id) override {
    if (rel_id == feed_.relation().getID()) done_feeding_input_relation_ = true;
  }
  QueryContext::insert_destination_id getInsertDestinationID() const override { return output_destination_index_; }
  const relation_id getOutputRelationID() const override { return output_relation_.getID(); }

 private:
  const CatalogRelation &build_relation_;
  InputFeed feed_;
  const std::vector<attribute_id> join_key_attributes_;
  const CatalogRelation &output_relation_;
  const QueryContext::insert_destination_id output_destination_index_;
  const QueryContext::join_hash_table_id hash_table_index_;
  const QueryContext::predicate_id residual_predicate_index_;
  const QueryContext::scalar_group_id selection_index_;
  const JoinType join_type_;
};

// One class for Hash{Inner,Semi,Anti}JoinWorkOrder (HashJoinOperator.cpp:450-987): the join type is a
// compile-time property of the kernel the C ABI instantiates, not of the host object.
class HashJoinWorkOrder : public WorkOrder {
 public:
  HashJoinWorkOrder(const std::size_t query_id, const DeviceExtent &probe, const std::vector<attribute_id> &join_key_attributes,
                    const QueryContext::Predicate *residual_predicate, const QueryContext::ScalarGroup *selection,
                    qsgpu_join_table_t hash_table, InsertDestination *output_destination, JoinType join_type,
                    std::vector<qs_lip_ref> lip_probe)
      : WorkOrder(query_id), probe_(probe), join_key_attributes_(join_key_attributes.begin(), join_key_attributes.end()), residual_predicate_(residual_predicate),
        selection_(selection), hash_table_(hash_table), output_destination_(output_destination), join_type_(join_type),
        lip_probe_(std::move(lip_probe)) {}
  void execute() override;

 private:
  const DeviceExtent probe_;
  const std::vector<std::uint32_t> join_key_attributes_;
  const QueryContext::Predicate *residual_predicate_;
  const QueryContext::ScalarGroup *selection_;
  qsgpu_join_table_t hash_table_;
  InsertDestination *output_destination_;
  const JoinType join_type_;
  const std::vector<qs_lip_ref> lip_probe_;
};

// ------------------------------------------------------------- Aggregation
class AggregationOperator : public RelationalOperator {
 public:
  AggregationOperator(const std::size_t query_id, const CatalogRelation &input_relation, bool input_relation_is_stored,
                      const QueryContext::aggregation_state_id aggr_state_index, const std::size_t num_partitions)
      : RelationalOperator(query_id, num_partitions), feed_(input_relation, input_relation_is_stored),
        aggr_state_index_(aggr_state_index) {}
  OperatorType getOperatorType() const override { return kAggregation; }
  std::string getName() const override { return "AggregationOperator"; }
  bool getAllWorkOrders(WorkOrdersContainer *container, QueryContext *query_context, StorageManager *storage_manager,
                        const tmb::client_id scheduler_client_id, tmb::MessageBus *bus) override;
  void feedInputBlock(const block_id input_block_id, const relation_id, const partition_id) override { feed_.feed(input_block_id); }

 private:
  InputFeed feed_;
  const QueryContext::aggregation_state_id aggr_state_index_;
};

class AggregationWorkOrder : public WorkOrder {
 public:
  AggregationWorkOrder(const std::size_t query_id, const DeviceExtent &input, qsgpu_agg_state_t state,
                       std::vector<qs_lip_ref> lip_probe)
      : WorkOrder(query_id), input_(input), state_(state), lip_probe_(std::move(lip_probe)) {}
  void execute() override;      // AggregationOperator.cpp:124 -> aggregateBlock

 private:
  const DeviceExtent input_;
  qsgpu_agg_state_t state_;
  const std::vector<qs_lip_ref> lip_probe_;
};

// BuildAggregationExistenceMapOperator (relational_operators/BuildAggregationExistenceMapOperator.hpp:61-150):
// marks, in the CollisionFreeVectorTable of an aggregation state, the keys that exist on the build side of a
// fused aggregate-join, so that they are finalized even when no probe-side row reaches them.
class BuildAggregationExistenceMapOperator : public RelationalOperator {
 public:
  BuildAggregationExistenceMapOperator(const std::size_t query_id, const CatalogRelation &input_relation,
                                       const attribute_id build_attribute, const bool input_relation_is_stored,
                                       const QueryContext::aggregation_state_id aggr_state_index,
                                       const std::size_t num_partitions)
      : RelationalOperator(query_id, num_partitions), feed_(input_relation, input_relation_is_stored),
        build_attribute_(build_attribute), aggr_state_index_(aggr_state_index) {}
  OperatorType getOperatorType() const override { return kBuildAggregationExistenceMap; }
  std::string getName() const override { return "BuildAggregationExistenceMapOperator"; }
  bool getAllWorkOrders(WorkOrdersContainer *container, QueryContext *query_context, StorageManager *storage_manager,
                        const tmb::client_id scheduler_client_id, tmb::MessageBus *bus) override;
  void feedInputBlock(const block_id input_block_id, const relation_id, const partition_id) override { feed_.feed(input_block_id); }

 private:
  InputFeed feed_;
  const attribute_id build_attribute_;
  const QueryContext::aggregation_state_id aggr_state_index_;
};

class BuildAggregationExistenceMapWorkOrder : public WorkOrder {
 public:
  BuildAggregationExistenceMapWorkOrder(const std::size_t query_id, const DeviceExtent &input, attribute_id build_attribute,
                                        qsgpu_agg_state_t state)
      : WorkOrder(query_id), input_(input), build_attribute_(build_attribute), state_(state) {}
  void execute() override;      // BuildAggregationExistenceMapOperator.cpp:176-212

 private:
  const DeviceExtent input_;
  const attribute_id build_attribute_;
  qsgpu_agg_state_t state_;
};

// InitializeAggregationOperator.cpp:36-93 memsets CollisionFreeVectorTable segments in parallel; the
// device state is zeroed by one fill kernel inside qsgpu_agg_create, so the work order has nothing left to do.
class InitializeAggregationOperator : public RelationalOperator {
 public:
  InitializeAggregationOperator(const std::size_t query_id, const QueryContext::aggregation_state_id aggr_state_index,
                                const std::size_t num_partitions = 1)
      : RelationalOperator(query_id, num_partitions), aggr_state_index_(aggr_state_index) {}
  OperatorType getOperatorType() const override { return kInitializeAggregation; }
  std::string getName() const override { return "InitializeAggregationOperator"; }
  bool getAllWorkOrders(WorkOrdersContainer *, QueryContext *, StorageManager *, const tmb::client_id, tmb::MessageBus *) override { return true; }

 private:
  const QueryContext::aggregation_state_id aggr_state_index_;
};

class FinalizeAggregationOperator : public RelationalOperator {
 public:
  FinalizeAggregationOperator(const std::size_t query_id, const QueryContext::aggregation_state_id aggr_state_index,
                              const std::size_t num_partitions, const bool has_repartition,
                              const std::size_t aggr_state_num_partitions, const CatalogRelation &output_relation,
                              const QueryContext::insert_destination_id output_destination_index)
      : RelationalOperator(query_id, num_partitions), aggr_state_index_(aggr_state_index), output_relation_(output_relation),
        output_destination_index_(output_destination_index) {}
  OperatorType getOperatorType() const override { return kFinalizeAggregation; }
  std::string getName() const override { return "FinalizeAggregationOperator"; }
  bool getAllWorkOrders(WorkOrdersContainer *container, QueryContext *query_context, StorageManager *storage_manager,
                        const tmb::client_id scheduler_client_id, tmb::MessageBus *bus) override;
  QueryContext::insert_destination_id getInsertDestinationID() const override { return output_destination_index_; }
  const relation_id getOutputRelationID() const override { return output_relation_.getID(); }

 private:
  const QueryContext::aggregation_state_id aggr_state_index_;
  const CatalogRelation &output_relation_;
  const QueryContext::insert_destination_id output_destination_index_;
  bool started_ = false;
};

class FinalizeAggregationWorkOrder : public WorkOrder {
 public:
  FinalizeAggregationWorkOrder(const std::size_t query_id, qsgpu_agg_state_t state, InsertDestination *output_destination,
                               qsgpu_comm_t merge_comm = nullptr)
      : WorkOrder(query_id), state_(state), output_destination_(output_destination), merge_comm_(merge_comm) {}
  void execute() override;      // FinalizeAggregationOperator.cpp:99 -> finalizeAggregate

 private:
  qsgpu_agg_state_t state_;
  InsertDestination *output_destination_;
  qsgpu_comm_t merge_comm_;     // non-null: merge the per-device partial states first (qsgpu_agg_merge_all)
};

class DestroyAggregationStateOperator : public RelationalOperator {
 public:
  DestroyAggregationStateOperator(const std::size_t query_id, const QueryContext::aggregation_state_id aggr_state_index,
                                  const std::size_t num_partitions = 1)
      : RelationalOperator(query_id, num_partitions), aggr_state_index_(aggr_state_index) {}
  OperatorType getOperatorType() const override { return kDestroyAggregationState; }
  std::string getName() const override { return "DestroyAggregationStateOperator"; }
  bool getAllWorkOrders(WorkOrdersContainer *container, QueryContext *query_context, StorageManager *storage_manager,
                        const tmb::client_id scheduler_client_id, tmb::MessageBus *bus) override;

 private:
  const QueryContext::aggregation_state_id aggr_state_index_;
  bool started_ = false;
};

class DestroyHashOperator : public RelationalOperator {
 public:
  DestroyHashOperator(const std::size_t query_id, const std::size_t num_partitions,
                      const QueryContext::join_hash_table_id hash_table_index)
      : RelationalOperator(query_id, num_partitions), hash_table_index_(hash_table_index) {}
  OperatorType getOperatorType() const override { return kDestroyHash; }
  std::string getName() const override { return "DestroyHashOperator"; }
  bool getAllWorkOrders(WorkOrdersContainer *container, QueryContext *query_context, StorageManager *storage_manager,
                        const tmb::client_id scheduler_client_id, tmb::MessageBus *bus) override;

 private:
  const QueryContext::join_hash_table_id hash_table_index_;
  bool started_ = false;
};

// Work order whose whole job is one call on the QueryContext (the Destroy* work orders).
class ContextCallWorkOrder : public WorkOrder {
 public:
  template <class F> ContextCallWorkOrder(const std::size_t query_id, F f) : WorkOrder(query_id), fn_(std::move(f)) {}
  void execute() override { fn_(); }
 private:
  std::function<void()> fn_;
};

// ------------------------------------------------------- top-k
// SortRunGeneration + SortMergeRun with LIMIT collapse into one work order on the device
// (qsgpu_topk).  Blocking: runs once its input is complete.
class SortMergeRunOperator : public RelationalOperator {
 public:
  SortMergeRunOperator(const std::size_t query_id, const CatalogRelation &input_relation,
                       const CatalogRelation &output_relation,
                       const QueryContext::insert_destination_id output_destination_index,
                       const QueryContext::sort_config_id sort_config_index, const std::size_t top_k,
                       const bool input_relation_is_stored)
      : RelationalOperator(query_id), feed_(input_relation, input_relation_is_stored), output_relation_(output_relation),
        output_destination_index_(output_destination_index), sort_config_index_(sort_config_index), top_k_(top_k) {}
  OperatorType getOperatorType() const override { return kSortMergeRun; }
  std::string getName() const override { return "SortMergeRunOperator"; }
  bool getAllWorkOrders(WorkOrdersContainer *container, QueryContext *query_context, StorageManager *storage_manager,
                        const tmb::client_id scheduler_client_id, tmb::MessageBus *bus) override;
  void feedInputBlock(const block_id input_block_id, const relation_id, const partition_id) override { feed_.feed(input_block_id); }
  QueryContext::insert_destination_id getInsertDestinationID() const override { return output_destination_index_; }
  const relation_id getOutputRelationID() const override { return output_relation_.getID(); }

 private:
  InputFeed feed_;
  const CatalogRelation &output_relation_;
  const QueryContext::insert_destination_id output_destination_index_;
  const QueryContext::sort_config_id sort_config_index_;
  const std::size_t top_k_;
};

class TopKWorkOrder : public WorkOrder {
 public:
  TopKWorkOrder(const std::size_t query_id, const DeviceExtent &input, const QueryContext::SortConfig *config,
                std::size_t top_k, InsertDestination *output_destination, qsgpu_comm_t gather_comm = nullptr)
      : WorkOrder(query_id), input_(input), config_(config), top_k_(top_k), output_destination_(output_destination),
        gather_comm_(gather_comm) {}
  void execute() override;

 private:
  const DeviceExtent input_;
  const QueryContext::SortConfig *config_;
  const std::size_t top_k_;
  InsertDestination *output_destination_;
  qsgpu_comm_t gather_comm_;    // non-null: all-gather every device's candidates and select again
};

}  // namespace quickstep
