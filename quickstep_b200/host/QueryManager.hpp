// A small scheduler with the reference's division of labour, used to drive and test the GPU
// operators outside the reference's build:
//   QueryPlan      query_optimizer/QueryPlan.hpp (DAG<RelationalOperator, bool is_pipeline_breaker>)
//   QueryManager   query_execution/QueryManagerBase.cpp + QueryManagerSingleNode.cpp:
//                  fetchNormalWorkOrders / processWorkOrderCompleteMessage / markOperatorFinished
//   Foreman/Worker query_execution/ForemanSingleNode.cpp:102, Worker.cpp:54-139
// Contracts kept: getAllWorkOrders / feedInputBlock / doneFeedingInputBlocks are called on the
// Foreman thread only; execute() runs on Worker threads, the work order is destroyed right after;
// an operator starts once its blocking dependencies finished; when an operator finishes, the blocks
// of its output relation are fed to its consumers, then doneFeedingInputBlocks.
// In an in-tree build none of this file is used: the GPU operators sit in the reference's own DAG.
#pragma once

#include <atomic>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <queue>
#include <string>
#include <utility>
#include <thread>
#include <vector>

#include "WorkOrder.hpp"

namespace quickstep {

class QueryPlan {
 public:
  typedef std::size_t DAGNodeIndex;
  // takes ownership; returns the operator index
  DAGNodeIndex addRelationalOperator(RelationalOperator *op) {
    ops_.emplace_back(op);
    op->setOperatorIndex(ops_.size() - 1);
    edges_.emplace_back();
    return ops_.size() - 1;
  }
  void addDirectDependency(DAGNodeIndex consumer, DAGNodeIndex producer, bool is_pipeline_breaker) {
    edges_[producer].push_back({consumer, is_pipeline_breaker});
  }
  std::size_t size() const { return ops_.size(); }
  RelationalOperator *op(DAGNodeIndex i) const { return ops_[i].get(); }
  struct Edge { DAGNodeIndex consumer; bool is_pipeline_breaker; };
  const std::vector<Edge> &consumers(DAGNodeIndex producer) const { return edges_[producer]; }

 private:
  std::vector<std::unique_ptr<RelationalOperator>> ops_;
  std::vector<std::vector<Edge>> edges_;
};

// The Worker threads (cli/QuickstepCli.cpp:251-263 starts --num_workers of them once per process).
class WorkerPool {
 public:
  explicit WorkerPool(int num_workers);
  ~WorkerPool();
  int size() const { return static_cast<int>(threads_.size()); }
  void submit(WorkOrder *w, std::size_t op_index);      // kWorkOrderMessage
  // kWorkOrderCompleteMessage -> operator index (+ the wall time execute() took, for the query profile
  // the reference prints with --visualize_execution_dag / -profile_and_report_workorder_perf)
  std::size_t waitForCompletion(double *execute_ms = nullptr);

 private:
  struct Message { WorkOrder *work_order; std::size_t op_index; };
  void workerLoop();
  std::mutex mu_;
  std::condition_variable work_cv_, done_cv_;
  std::queue<Message> work_queue_;
  std::queue<std::pair<std::size_t, double>> done_queue_;
  bool shutdown_ = false;
  std::atomic<int> n_work_{0}, n_done_{0};       // queue lengths, polled without the lock (bounded spinning)
  std::atomic<bool> shutdown_flag_{false};
  std::vector<std::thread> threads_;
};

class QueryManager {
 public:
  QueryManager(QueryPlan *plan, QueryContext *context, StorageManager *storage_manager, WorkerPool *workers)
      : plan_(plan), context_(context), sm_(storage_manager), workers_(workers), container_(plan->size()) {}
  // Admit -> completion (the window the CLI prints as "Time:", cli/QuickstepCli.cpp:373-388).
  void run();
  std::size_t numWorkOrdersExecuted(std::size_t op_index) const { return executed_[op_index]; }
  std::size_t totalWorkOrdersExecuted() const { std::size_t n = 0; for (auto x : executed_) n += x; return n; }
  // one line per operator: index, name, work orders, ms inside execute(), ms inside getAllWorkOrders()
  std::string profile() const;

 private:
  void fetchNormalWorkOrders(std::size_t op);
  void markOperatorFinished(std::size_t op);

  QueryPlan *plan_;
  QueryContext *context_;
  StorageManager *sm_;
  WorkerPool *workers_;
  WorkOrdersContainer container_;
  std::vector<bool> done_gen_, finished_;
  std::vector<std::size_t> pending_, blocking_deps_, executed_;
  std::vector<double> execute_ms_, generate_ms_;
};

}  // namespace quickstep
