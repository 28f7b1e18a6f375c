#include "QueryManager.hpp"

#include <chrono>
#include <sstream>

namespace quickstep {

void QueryManager::fetchNormalWorkOrders(std::size_t op) {
  if (done_gen_[op] || blocking_deps_[op] != 0) return;
  const auto t0 = std::chrono::steady_clock::now();
  done_gen_[op] = plan_->op(op)->getAllWorkOrders(&container_, context_, sm_, /*scheduler_client_id=*/0, /*bus=*/nullptr);
  generate_ms_[op] += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

void QueryManager::markOperatorFinished(std::size_t op) {
  finished_[op] = true;
  RelationalOperator *producer = plan_->op(op);
  const relation_id out_rel = producer->getOutputRelationID();
  InsertDestination *dest = context_->getInsertDestination(producer->getInsertDestinationID());
  for (const QueryPlan::Edge &e : plan_->consumers(op)) {
    RelationalOperator *consumer = plan_->op(e.consumer);
    if (dest != nullptr && out_rel >= 0) {
      for (const std::pair<block_id, partition_id> &b : dest->getTouchedBlocks()) consumer->feedInputBlock(b.first, out_rel, b.second);
      consumer->doneFeedingInputBlocks(out_rel);
    }
    if (e.is_pipeline_breaker) --blocking_deps_[e.consumer];
    fetchNormalWorkOrders(e.consumer);
  }
}

WorkerPool::WorkerPool(int num_workers) {
  for (int i = 0; i < (num_workers < 1 ? 1 : num_workers); ++i) threads_.emplace_back([this] { workerLoop(); });
}

WorkerPool::~WorkerPool() {
  {
    std::lock_guard<std::mutex> lk(mu_);
    shutdown_ = true;
    shutdown_flag_.store(true, std::memory_order_release);
  }
  work_cv_.notify_all();
  for (auto &t : threads_) t.join();
}

// A GPU work order is a few tens of microseconds of host work (lower, look the kernel up, enqueue), so the
// latency of waking a sleeping thread (20-50 us each way) would dominate a query of five operators.  Both sides
// therefore poll an atomic counter for a bounded time before they fall back to the condition variable: Workers for
// kSpinUs after their last work order (they stay hot for the length of a query and sleep between queries), the
// Foreman while it waits for a completion.
namespace {
constexpr int kSpinUs = 300;
template <class Pred>
bool spinUntil(Pred &&ready) {
  const auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(kSpinUs);
  for (int i = 0;; ++i) {
    if (ready()) return true;
    if ((i & 63) == 63 && std::chrono::steady_clock::now() > deadline) return false;
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
}
}  // namespace

void WorkerPool::submit(WorkOrder *w, std::size_t op_index) {
  {
    std::lock_guard<std::mutex> lk(mu_);
    work_queue_.push({w, op_index});
    n_work_.fetch_add(1, std::memory_order_release);
  }
  work_cv_.notify_one();
}

std::size_t WorkerPool::waitForCompletion(double *execute_ms) {
  spinUntil([&] { return n_done_.load(std::memory_order_acquire) != 0; });
  std::unique_lock<std::mutex> lk(mu_);
  done_cv_.wait(lk, [&] { return !done_queue_.empty(); });
  const std::pair<std::size_t, double> m = done_queue_.front();
  done_queue_.pop();
  n_done_.fetch_sub(1, std::memory_order_release);
  if (execute_ms) *execute_ms = m.second;
  return m.first;
}

void WorkerPool::workerLoop() {
  for (;;) {
    Message m;
    spinUntil([&] { return n_work_.load(std::memory_order_acquire) != 0 || shutdown_flag_.load(std::memory_order_acquire); });
    {
      std::unique_lock<std::mutex> lk(mu_);
      work_cv_.wait(lk, [&] { return shutdown_ || !work_queue_.empty(); });
      if (work_queue_.empty()) return;
      m = work_queue_.front();
      work_queue_.pop();
      n_work_.fetch_sub(1, std::memory_order_release);
    }
    std::unique_ptr<WorkOrder> wo(m.work_order);     // Worker.cpp:127-139: executed, then destroyed
    const auto t0 = std::chrono::steady_clock::now();
    wo->execute();
    wo.reset();
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    {
      std::lock_guard<std::mutex> lk(mu_);
      done_queue_.push({m.op_index, ms});
      n_done_.fetch_add(1, std::memory_order_release);
    }
    done_cv_.notify_one();
  }
}

void QueryManager::run() {
  const std::size_t n = plan_->size();
  done_gen_.assign(n, false);
  finished_.assign(n, false);
  pending_.assign(n, 0);
  blocking_deps_.assign(n, 0);
  executed_.assign(n, 0);
  execute_ms_.assign(n, 0.0);
  generate_ms_.assign(n, 0.0);
  for (std::size_t p = 0; p < n; ++p)
    for (const QueryPlan::Edge &e : plan_->consumers(p))
      if (e.is_pipeline_breaker) ++blocking_deps_[e.consumer];

  for (std::size_t op = 0; op < n; ++op) fetchNormalWorkOrders(op);
  std::size_t n_finished = 0;
  while (n_finished < n) {
    // dispatch everything that is ready
    bool dispatched = false;
    for (std::size_t op = 0; op < n; ++op) {
      while (WorkOrder *w = container_.getNormalWorkOrder(op)) {
        ++pending_[op];
        ++executed_[op];
        workers_->submit(w, op);
        dispatched = true;
      }
    }
    // operators with nothing left
    bool progressed = false;
    for (std::size_t op = 0; op < n; ++op) {
      if (!finished_[op] && done_gen_[op] && pending_[op] == 0 && !container_.hasNormalWorkOrder(op)) {
        markOperatorFinished(op);
        ++n_finished;
        progressed = true;
      }
    }
    if (progressed || n_finished == n) continue;
    (void)dispatched;
    // wait for a completion
    double ms = 0;
    const std::size_t op = workers_->waitForCompletion(&ms);
    execute_ms_[op] += ms;
    --pending_[op];
    fetchNormalWorkOrders(op);
  }
}

std::string QueryManager::profile() const {
  std::ostringstream o;
  for (std::size_t op = 0; op < plan_->size(); ++op)
    o << op << " " << plan_->op(op)->getName() << " work_orders=" << executed_[op] << " execute_ms=" << execute_ms_[op]
      << " get_all_work_orders_ms=" << generate_ms_[op] << "\n";
  return o.str();
}

}  // namespace quickstep
