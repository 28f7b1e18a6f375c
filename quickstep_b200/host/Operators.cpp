#include "Operators.hpp"

namespace quickstep {

std::uint64_t FLAGS_gpu_rows_per_workorder = 0;

std::vector<DeviceExtent> InputFeed::take(StorageManager *sm, std::uint64_t needed_attrs) {
  std::vector<DeviceExtent> out;
  if (stored_) {
    if (started_) return out;
    started_ = true;
    // K0: stage whatever is not resident yet (one batch), then one work order per run of adjacent blocks
    return sm->stagedExtents(relation_, needed_attrs, FLAGS_gpu_rows_per_workorder);
  }
  for (block_id b : pending_) out.push_back(sm->blockExtent(b));
  pending_.clear();
  return out;
}

namespace {

struct Lowered {
  ExprSet es;
  int predicate_root = -1;
  std::vector<std::int32_t> roots;
};

void addPredicate(Lowered *L, const QueryContext::Predicate *p) {
  if (!p) return;
  const int off = L->es.append(p->exprs);
  L->predicate_root = p->root + off;
}

void addScalars(Lowered *L, const QueryContext::ScalarGroup *g) {
  if (!g) return;
  const int off = L->es.append(g->exprs);
  for (int r : g->roots) L->roots.push_back(r + off);
}

std::uint64_t attrsOf(const QueryContext::Predicate *p) { return p ? p->exprs.referencedAttributes(0) : 0; }
std::uint64_t attrsOf(const std::vector<qs_lip_ref> &refs) {
  std::uint64_t m = 0;
  for (const qs_lip_ref &r : refs) if (r.attr < 64) m |= 1ull << r.attr;
  return m;
}

// Probe filters of a work order that is about to run, most selective first by what the kernels have seen so far.
std::vector<qs_lip_ref> rankedNow(const std::vector<qs_lip_ref> &lip_probe) {
  std::vector<qs_lip_ref> refs = lip_probe;
  RankLIPFiltersByMissRate(&refs);
  return refs;
}

qs_scan makeScan(const DeviceExtent &in, const qs_expr_set *es, int predicate_root, const std::vector<qs_lip_ref> &lip_probe) {
  qs_scan s{};
  s.input = in.relation;
  s.row_begin = in.row_begin;
  s.row_end = in.row_end;
  s.exprs = es;
  s.predicate_root = predicate_root;
  s.n_lip_probe = static_cast<std::uint32_t>(lip_probe.size());
  s.lip_probe = lip_probe.empty() ? nullptr : lip_probe.data();
  return s;
}

}  // namespace

// ------------------------------------------------------------------ Select
bool SelectOperator::getAllWorkOrders(WorkOrdersContainer *container, QueryContext *query_context,
                                      StorageManager *storage_manager, const tmb::client_id, tmb::MessageBus *) {
  InsertDestination *dest = query_context->getInsertDestination(output_destination_index_);
  std::uint64_t needed = attrsOf(query_context->getPredicate(predicate_index_)) |
                         attrsOf(query_context->lipRefs(lip_deployment_index_, QueryContext::LIPAction::kProbe));
  if (simple_projection_) { for (attribute_id a : simple_selection_) needed |= 1ull << a; }
  else needed |= query_context->getScalarGroup(selection_index_).exprs.referencedAttributes(0);
  // several devices: a row-by-row operator keeps its input's distribution
  storage_manager->setPartitioned(output_relation_.getID(), storage_manager->isPartitioned(feed_.relation().getID()));
  for (const DeviceExtent &e : feed_.take(storage_manager, needed)) {
    container->addNormalWorkOrder(
        new SelectWorkOrder(query_id_, feed_.relation(), e, query_context->getPredicate(predicate_index_),
                            simple_projection_ ? nullptr : &query_context->getScalarGroup(selection_index_),
                            simple_projection_ ? &simple_selection_ : nullptr, dest,
                            query_context->lipRefs(lip_deployment_index_, QueryContext::LIPAction::kProbe)),
        op_index_);
  }
  return feed_.exhausted(done_feeding_input_relation_);
}

void SelectWorkOrder::execute() {
  Lowered L;
  addPredicate(&L, predicate_);
  if (simple_selection_) {
    for (attribute_id a : *simple_selection_) L.roots.push_back(L.es.attr(a, input_relation_.getAttributeById(a).type));
  } else {
    addScalars(&L, selection_);
  }
  const qs_expr_set es = L.es.view();
  const std::vector<qs_lip_ref> probe = rankedNow(lip_probe_);
  const qs_scan scan = makeScan(input_, &es, L.predicate_root, probe);
  QS_CHECK_GPU(qsgpu_select(&scan, static_cast<std::uint32_t>(L.roots.size()), L.roots.data(),
                            output_destination_->deviceRelation()));
}

// ---------------------------------------------------------- BuildLIPFilter
bool BuildLIPFilterOperator::getAllWorkOrders(WorkOrdersContainer *container, QueryContext *query_context,
                                              StorageManager *storage_manager, const tmb::client_id, tmb::MessageBus *) {
  const std::uint64_t needed = attrsOf(query_context->getPredicate(build_side_predicate_index_)) |
                               attrsOf(query_context->lipRefs(lip_deployment_index_, QueryContext::LIPAction::kProbe)) |
                               attrsOf(query_context->lipRefs(lip_deployment_index_, QueryContext::LIPAction::kBuild));
  query_context->noteLIPFiltersBuiltFrom(lip_deployment_index_, storage_manager->isPartitioned(feed_.relation().getID()));
  for (const DeviceExtent &e : feed_.take(storage_manager, needed)) {
    container->addNormalWorkOrder(
        new BuildLIPFilterWorkOrder(query_id_, e, query_context->getPredicate(build_side_predicate_index_),
                                    query_context->lipRefs(lip_deployment_index_, QueryContext::LIPAction::kProbe),
                                    query_context->lipRefs(lip_deployment_index_, QueryContext::LIPAction::kBuild)),
        op_index_);
  }
  return feed_.exhausted(done_feeding_input_relation_);
}

void BuildLIPFilterWorkOrder::execute() {
  Lowered L;
  addPredicate(&L, build_side_predicate_);
  const qs_expr_set es = L.es.view();
  const std::vector<qs_lip_ref> probe = rankedNow(lip_probe_);
  const qs_scan scan = makeScan(input_, &es, L.predicate_root, probe);
  QS_CHECK_GPU(qsgpu_build_lip_filter(&scan, static_cast<std::uint32_t>(lip_build_.size()), lip_build_.data()));
}

// --------------------------------------------------------------- BuildHash
bool BuildHashOperator::getAllWorkOrders(WorkOrdersContainer *container, QueryContext *query_context,
                                         StorageManager *storage_manager, const tmb::client_id, tmb::MessageBus *) {
  std::vector<DeviceExtent> extents = feed_.take(storage_manager);
  const bool partitioned_input = storage_manager->multiDevice() && storage_manager->isPartitioned(feed_.relation().getID());
  if (!extents.empty() && partitioned_input && num_partitions_ > 1) {
    // Partition-wise join: build and probe side are partitioned on the join key by the same scheme, so every
    // device builds the table of ITS partition and probes it with its partition of the probe side -- the
    // reference's own mechanism for partitioned relations (num_partitions hash tables, one per partition,
    // query_execution/QueryContext.cpp:78-97; BuildHashOperator.hpp `num_partitions`).  No rows cross devices, and
    // a LIP filter filled here is complete for every probe row this device will ever see.
    query_context->noteLIPFiltersBuiltFrom(lip_deployment_index_, false);
  } else if (!extents.empty() && partitioned_input) {
    // Broadcast join (SURVEY.md section 8e "Join, small build"): ONE logical table (num_partitions == 1) whose
    // build side was filtered in shares; its rows are all-gathered so that every device builds its own copy of
    // the table and probes its partition of the probe side locally, with no shuffle.  Needs the complete input:
    // the operator runs once its producer has finished.
    QS_CHECK(feed_.relation().isTemporary() && done_feeding_input_relation_);
    DeviceExtent all;
    all.relation = storage_manager->replicated(feed_.relation());
    extents.assign(1, all);
    query_context->noteLIPFiltersBuiltFrom(lip_deployment_index_, false);     // built from every device's rows
  } else if (!extents.empty()) {
    query_context->noteLIPFiltersBuiltFrom(lip_deployment_index_, storage_manager->isPartitioned(feed_.relation().getID()));
  }
  for (const DeviceExtent &e : extents) {
    container->addNormalWorkOrder(
        new BuildHashWorkOrder(query_id_, e, join_key_attributes_, query_context->getPredicate(build_predicate_index_),
                               query_context->getJoinHashTable(hash_table_index_),
                               query_context->lipRefs(lip_deployment_index_, QueryContext::LIPAction::kProbe),
                               query_context->lipRefs(lip_deployment_index_, QueryContext::LIPAction::kBuild)),
        op_index_);
  }
  return feed_.exhausted(done_feeding_input_relation_);
}

void BuildHashWorkOrder::execute() {
  Lowered L;
  addPredicate(&L, predicate_);
  const qs_expr_set es = L.es.view();
  const std::vector<qs_lip_ref> probe = rankedNow(lip_probe_);
  const qs_scan scan = makeScan(input_, &es, L.predicate_root, probe);
  QS_CHECK_GPU(qsgpu_join_build_composite(hash_table_, &scan, static_cast<std::uint32_t>(join_key_attributes_.size()),
                                          join_key_attributes_.data(), static_cast<std::uint32_t>(lip_build_.size()),
                                          lip_build_.empty() ? nullptr : lip_build_.data()));
}

// ---------------------------------------------------------------- HashJoin
bool HashJoinOperator::getAllWorkOrders(WorkOrdersContainer *container, QueryContext *query_context,
                                        StorageManager *storage_manager, const tmb::client_id, tmb::MessageBus *) {
  InsertDestination *dest = query_context->getInsertDestination(output_destination_index_);
  storage_manager->setPartitioned(output_relation_.getID(), storage_manager->isPartitioned(feed_.relation().getID()));
  for (const DeviceExtent &e : feed_.take(storage_manager)) {
    container->addNormalWorkOrder(
        new HashJoinWorkOrder(query_id_, e, join_key_attributes_, query_context->getPredicate(residual_predicate_index_),
                              &query_context->getScalarGroup(selection_index_),
                              query_context->getJoinHashTable(hash_table_index_), dest, join_type_,
                              query_context->lipRefs(lip_deployment_index_, QueryContext::LIPAction::kProbe)),
        op_index_);
  }
  return feed_.exhausted(done_feeding_input_relation_);
}

void HashJoinWorkOrder::execute() {
  Lowered L;
  int residual_root = -1;
  if (residual_predicate_) {
    const int off = L.es.append(residual_predicate_->exprs);
    residual_root = residual_predicate_->root + off;
  }
  addScalars(&L, selection_);
  const qs_expr_set es = L.es.view();
  const std::vector<qs_lip_ref> probe = rankedNow(lip_probe_);
  const qs_scan scan = makeScan(probe_, &es, -1, probe);
  std::uint32_t jt = QS_JOIN_INNER;
  switch (join_type_) {
    case JoinType::kInnerJoin: jt = QS_JOIN_INNER; break;
    case JoinType::kLeftSemiJoin: jt = QS_JOIN_LEFT_SEMI; break;
    case JoinType::kLeftAntiJoin: jt = QS_JOIN_LEFT_ANTI; break;
    case JoinType::kLeftOuterJoin: jt = QS_JOIN_LEFT_OUTER; break;
  }
  QS_CHECK_GPU(qsgpu_join_probe_composite(hash_table_, &scan, static_cast<std::uint32_t>(join_key_attributes_.size()),
                                          join_key_attributes_.data(), jt, residual_root,
                                          static_cast<std::uint32_t>(L.roots.size()), L.roots.data(),
                                          output_destination_->deviceRelation()));
}

// ------------------------------------------------------------- Aggregation
bool AggregationOperator::getAllWorkOrders(WorkOrdersContainer *container, QueryContext *query_context,
                                           StorageManager *storage_manager, const tmb::client_id, tmb::MessageBus *) {
  const std::uint64_t needed = query_context->getAggregationSpec(aggr_state_index_).exprs.referencedAttributes(0) |
                               attrsOf(query_context->lipRefs(lip_deployment_index_, QueryContext::LIPAction::kProbe));
  if (storage_manager->isPartitioned(feed_.relation().getID())) query_context->noteAggregationInput(aggr_state_index_, true);
  for (const DeviceExtent &e : feed_.take(storage_manager, needed)) {
    container->addNormalWorkOrder(
        new AggregationWorkOrder(query_id_, e, query_context->getAggregationState(aggr_state_index_),
                                 query_context->lipRefs(lip_deployment_index_, QueryContext::LIPAction::kProbe)),
        op_index_);
  }
  return feed_.exhausted(done_feeding_input_relation_);
}

void AggregationWorkOrder::execute() {
  const std::vector<qs_lip_ref> probe = rankedNow(lip_probe_);
  QS_CHECK_GPU(qsgpu_agg_run(state_, input_.relation, input_.row_begin, input_.row_end,
                             static_cast<std::uint32_t>(probe.size()), probe.empty() ? nullptr : probe.data()));
}

bool BuildAggregationExistenceMapOperator::getAllWorkOrders(WorkOrdersContainer *container, QueryContext *query_context,
                                                            StorageManager *storage_manager, const tmb::client_id,
                                                            tmb::MessageBus *) {
  for (const DeviceExtent &e : feed_.take(storage_manager, 1ull << build_attribute_)) {
    container->addNormalWorkOrder(
        new BuildAggregationExistenceMapWorkOrder(query_id_, e, build_attribute_,
                                                  query_context->getAggregationState(aggr_state_index_)),
        op_index_);
  }
  return feed_.exhausted(done_feeding_input_relation_);
}

void BuildAggregationExistenceMapWorkOrder::execute() {
  qs_lip_ref target{};
  QS_CHECK_GPU(qsgpu_agg_existence_map(state_, &target.lip));
  target.attr = static_cast<std::uint32_t>(build_attribute_);
  const std::vector<qs_lip_ref> no_probe;
  const qs_scan scan = makeScan(input_, nullptr, -1, no_probe);
  QS_CHECK_GPU(qsgpu_build_lip_filter(&scan, 1, &target));
}

bool FinalizeAggregationOperator::getAllWorkOrders(WorkOrdersContainer *container, QueryContext *query_context,
                                                   StorageManager *storage_manager, const tmb::client_id, tmb::MessageBus *) {
  if (!started_) {
    started_ = true;
    // several devices: per-device partial states are merged first (AggregationHandle::mergeStates across
    // GPUs) unless the input is partitioned on the group-by key, in which case every device finalizes its own
    // groups and the result stays partitioned
    const bool partial = query_context->aggregationIsPartial(aggr_state_index_);
    const bool merge = partial && !query_context->getAggregationSpec(aggr_state_index_).partitioned_on_group_by;
    storage_manager->setPartitioned(output_relation_.getID(), partial && !merge);
    container->addNormalWorkOrder(
        new FinalizeAggregationWorkOrder(query_id_, query_context->getAggregationState(aggr_state_index_),
                                         query_context->getInsertDestination(output_destination_index_),
                                         merge ? query_context->comm() : nullptr),
        op_index_);
  }
  return true;
}

void FinalizeAggregationWorkOrder::execute() {
  qsgpu_relation_t out = nullptr;
  if (merge_comm_) QS_CHECK_GPU(qsgpu_agg_merge_all(state_, merge_comm_));
  // enqueue only: NULLs of aggregates over zero rows travel in the output relation's per-row NULL mask
  QS_CHECK_GPU(qsgpu_agg_finalize(state_, &out, nullptr));
  output_destination_->adopt(out);
}

bool DestroyAggregationStateOperator::getAllWorkOrders(WorkOrdersContainer *container, QueryContext *query_context,
                                                       StorageManager *, const tmb::client_id, tmb::MessageBus *) {
  if (!started_) {
    started_ = true;
    const QueryContext::aggregation_state_id id = aggr_state_index_;
    container->addNormalWorkOrder(new ContextCallWorkOrder(query_id_, [query_context, id] { query_context->destroyAggregationState(id); }),
                                  op_index_);
  }
  return true;
}

bool DestroyHashOperator::getAllWorkOrders(WorkOrdersContainer *container, QueryContext *query_context, StorageManager *,
                                           const tmb::client_id, tmb::MessageBus *) {
  if (!started_) {
    started_ = true;
    const QueryContext::join_hash_table_id id = hash_table_index_;
    container->addNormalWorkOrder(new ContextCallWorkOrder(query_id_, [query_context, id] { query_context->destroyJoinHashTable(id); }),
                                  op_index_);
  }
  return true;
}

// ------------------------------------------------------------------- top-k
bool SortMergeRunOperator::getAllWorkOrders(WorkOrdersContainer *container, QueryContext *query_context,
                                            StorageManager *storage_manager, const tmb::client_id, tmb::MessageBus *) {
  // several devices, partitioned input: every device selects its own top k, the candidates are all-gathered and
  // the final k are selected from them on every device (SortMergeRunOperator's last merge pass); the result is replicated
  const bool gather = storage_manager->multiDevice() && storage_manager->isPartitioned(feed_.relation().getID());
  storage_manager->setPartitioned(output_relation_.getID(), false);
  for (const DeviceExtent &e : feed_.take(storage_manager)) {
    container->addNormalWorkOrder(new TopKWorkOrder(query_id_, e, &query_context->getSortConfig(sort_config_index_), top_k_,
                                                    query_context->getInsertDestination(output_destination_index_),
                                                    gather ? query_context->comm() : nullptr),
                                  op_index_);
  }
  return feed_.exhausted(done_feeding_input_relation_);
}

void TopKWorkOrder::execute() {
  qsgpu_relation_t out = nullptr;
  QS_CHECK_GPU(qsgpu_topk(input_.relation, static_cast<std::uint32_t>(config_->keys.size()), config_->keys.data(), top_k_, &out));
  if (gather_comm_) {
    qsgpu_relation_t all = nullptr, top = nullptr;
    // every rank contributes at most top_k_ rows: the one-kernel gather over peer memory when the ranks have it
    QS_CHECK_GPU(qsgpu_relation_allgather_small(out, gather_comm_, top_k_, &all));
    QS_CHECK_GPU(qsgpu_topk(all, static_cast<std::uint32_t>(config_->keys.size()), config_->keys.data(), top_k_, &top));
    QS_CHECK_GPU(qsgpu_relation_destroy(all));
    QS_CHECK_GPU(qsgpu_relation_destroy(out));
    out = top;
  }
  output_destination_->adopt(out);
}

}  // namespace quickstep
