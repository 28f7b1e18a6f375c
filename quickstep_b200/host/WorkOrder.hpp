// The drop-in seam, with the reference's signatures:
//   WorkOrder::execute()                       relational_operators/WorkOrder.hpp:251
//   RelationalOperator::getAllWorkOrders(...)  relational_operators/RelationalOperator.hpp:132-136
//   feedInputBlock / doneFeedingInputBlocks    RelationalOperator.hpp:174,196
//   WorkOrdersContainer::addNormalWorkOrder    query_execution/WorkOrdersContainer.hpp:243
#pragma once

#include <deque>
#include <memory>
#include <string>
#include <unordered_set>
#include <vector>

#include "QueryContext.hpp"

namespace quickstep {

class WorkOrder {
 public:
  virtual ~WorkOrder() {}
  // Runs on a Worker thread.  Everything it needs was captured at construction
  // as non-owning pointers to objects of the QueryContext (WorkOrder.hpp:53-335).
  // A GPU work order lowers its expressions, enqueues its kernels on the
  // device's stream and returns; it throws nothing and returns nothing --
  // failure is fatal (QS_CHECK_GPU), the reference's convention.
  virtual void execute() = 0;
  std::size_t getQueryID() const { return query_id_; }
  partition_id getPartitionId() const { return partition_id_; }

 protected:
  explicit WorkOrder(std::size_t query_id, partition_id part_id = 0) : query_id_(query_id), partition_id_(part_id) {}
  const std::size_t query_id_;
  const partition_id partition_id_;
};

class WorkOrdersContainer {
 public:
  explicit WorkOrdersContainer(std::size_t num_operators) : normal_(num_operators) {}
  void addNormalWorkOrder(WorkOrder *workorder, std::size_t operator_index) { normal_[operator_index].emplace_back(workorder); }
  bool hasNormalWorkOrder(std::size_t operator_index) const { return !normal_[operator_index].empty(); }
  WorkOrder *getNormalWorkOrder(std::size_t operator_index) {
    if (normal_[operator_index].empty()) return nullptr;
    WorkOrder *w = normal_[operator_index].front().release();
    normal_[operator_index].pop_front();
    return w;
  }
  std::size_t getNumNormalWorkOrders(std::size_t operator_index) const { return normal_[operator_index].size(); }

 private:
  std::vector<std::deque<std::unique_ptr<WorkOrder>>> normal_;
};

class RelationalOperator {
 public:
  virtual ~RelationalOperator() {}
  enum OperatorType : std::uint8_t {      // values of the operators on this path (RelationalOperator.hpp:65-96)
    kAggregation = 0, kBuildAggregationExistenceMap = 1, kBuildHash = 2, kBuildLIPFilter = 3, kDestroyAggregationState = 7, kDestroyHash = 8,
    kFinalizeAggregation = 10, kInitializeAggregation = 11, kInnerJoin = 12, kLeftAntiJoin = 14,
    kLeftOuterJoin = 15, kLeftSemiJoin = 16, kSelect = 20, kSortMergeRun = 21
  };
  virtual OperatorType getOperatorType() const = 0;
  virtual std::string getName() const = 0;
  // Called only on the Foreman thread, possibly many times; returns true when
  // no more work orders will ever be generated.
  virtual bool getAllWorkOrders(WorkOrdersContainer *container, QueryContext *query_context,
                                StorageManager *storage_manager, const tmb::client_id scheduler_client_id,
                                tmb::MessageBus *bus) = 0;
  virtual void feedInputBlock(const block_id input_block_id, const relation_id input_relation_id,
                              const partition_id part_id) {}
  virtual void doneFeedingInputBlocks(const relation_id rel_id) { done_feeding_input_relation_ = true; }
  virtual QueryContext::insert_destination_id getInsertDestinationID() const { return QueryContext::kInvalidInsertDestinationId; }
  virtual const relation_id getOutputRelationID() const { return -1; }
  void setOperatorIndex(std::size_t operator_index) { op_index_ = operator_index; }
  std::size_t getOperatorIndex() const { return op_index_; }
  std::size_t getNumPartitions() const { return num_partitions_; }
  void deployLIPFilters(const QueryContext::lip_deployment_id lip_deployment_index,
                        const std::unordered_set<QueryContext::lip_filter_id> &lip_filter_indexes) {
    lip_deployment_index_ = lip_deployment_index;
    lip_filter_indexes_ = lip_filter_indexes;
  }

 protected:
  explicit RelationalOperator(std::size_t query_id, std::size_t num_partitions = 1u)
      : query_id_(query_id), num_partitions_(num_partitions), done_feeding_input_relation_(false), op_index_(0),
        lip_deployment_index_(QueryContext::kInvalidLIPDeploymentId) {}
  const std::size_t query_id_;
  const std::size_t num_partitions_;
  bool done_feeding_input_relation_;
  std::size_t op_index_;
  QueryContext::lip_deployment_id lip_deployment_index_;
  std::unordered_set<QueryContext::lip_filter_id> lip_filter_indexes_;
};

}  // namespace quickstep
