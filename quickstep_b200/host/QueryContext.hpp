// Per-query execution state shared by the work orders of a query: the GPU twin
// of query_execution/QueryContext.hpp:57-620 (built from serialization::QueryContext,
// query_execution/QueryContext.cpp:57-171).  Same id spaces, same getters; what
// differs is what an id resolves to:
//   aggregation_state_id -> qsgpu_agg_state_t       (AggregationOperationState)
//   join_hash_table_id   -> qsgpu_join_table_t      (JoinHashTable)
//   lip_filter_id        -> qsgpu_lip_t             (LIPFilter)
//   predicate_id / scalar_group_id -> expression sets (Predicate, vector<Scalar>)
//   insert_destination_id -> InsertDestination over a device-resident temporary relation
// Objects are owned here and released by the Destroy* work orders or with the
// context, as in the reference.
#pragma once

#include <algorithm>
#include <memory>
#include <mutex>
#include <vector>

#include "ExprSet.hpp"
#include "QsTypes.hpp"
#include "StorageManager.hpp"

namespace quickstep {

// storage/InsertDestination.hpp: the sink of Select / HashJoin / FinalizeAggregation.
// The device relation is created on first use with the capacity the plan estimated.
class InsertDestination {
 public:
  InsertDestination(const CatalogRelation *relation, std::uint64_t capacity_rows, StorageManager *sm)
      : relation_(relation), capacity_(capacity_rows), sm_(sm) {}
  virtual ~InsertDestination() {}
  const CatalogRelation &getRelation() const { return *relation_; }
  qsgpu_relation_t deviceRelation() { sm_->createTemporary(*relation_, capacity_); return sm_->temporary(*relation_); }
  // InsertDestination::bulkInsertTuples (storage/InsertDestination.cpp:202-216): rows that leave the device, given as
  // host column vectors, go into SplitRowStore blocks of the output relation (the reference's layout for temporaries);
  // called on the thread that runs the work order / reads the result, like the reference requires
  // (storage/InsertDestination.cpp:403-406).  The blocks are what getTouchedBlocks() hands to a CPU consumer.
  std::vector<block_id> bulkInsertTuples(const std::vector<const void *> &columns, std::uint64_t n_rows) {
    return sm_->insertTuples(*relation_, columns, n_rows);
  }
  // FinalizeAggregation / top-k create their output relation themselves
  void adopt(qsgpu_relation_t handle) { sm_->adoptTemporary(*relation_, handle); }
  // QueryManagerBase::markOperatorFinished -> getPartiallyFilledBlocks: the blocks to feed downstream
  // (block, partition id) pairs: what feedInputBlock(block_id, relation_id, partition_id) receives downstream
  virtual std::vector<std::pair<block_id, partition_id>> getTouchedBlocks() {
    return {{sm_->createTemporary(*relation_, capacity_), 0}};
  }

 protected:
  const CatalogRelation *relation_;
  std::uint64_t capacity_;
  StorageManager *sm_;
};

// catalog/PartitionSchemeHeader.hpp:170-218: hash partitioning on one attribute.
struct HashPartitionSchemeHeader {
  std::size_t num_partitions = 1;
  attribute_id partition_attribute = 0;
};

// storage/InsertDestination.cpp:471-722.  The reference routes every tuple to the blocks of its partition as it is
// inserted; on the device the operator writes its output rows once, and when it has finished K8 regroups them by
// HashPartitionSchemeHeader's partition function (qsgpu_hash_partition: one histogram pass + one scatter pass).
// Downstream operators are fed one block per partition together with its partition id, exactly what the reference's
// scheduler does with the blocks of a partitioned relation (QueryManagerBase.cpp: feedInputBlock(block, rel, part_id)).
class PartitionAwareInsertDestination : public InsertDestination {
 public:
  PartitionAwareInsertDestination(const HashPartitionSchemeHeader &header, const CatalogRelation *relation,
                                  std::uint64_t capacity_rows, StorageManager *sm)
      : InsertDestination(relation, capacity_rows, sm), header_(header) {}
  const HashPartitionSchemeHeader &getPartitionSchemeHeader() const { return header_; }
  std::vector<std::pair<block_id, partition_id>> getTouchedBlocks() override {
    sm_->createTemporary(*relation_, capacity_);
    if (blocks_.empty()) blocks_ = sm_->repartitionTemporary(*relation_, header_.partition_attribute, header_.num_partitions);
    std::vector<std::pair<block_id, partition_id>> out;
    for (std::size_t p = 0; p < blocks_.size(); ++p) out.emplace_back(blocks_[p], p);
    return out;
  }

 private:
  HashPartitionSchemeHeader header_;
  std::vector<block_id> blocks_;
};

// The adaptive part of LIPFilterAdaptiveProber: most selective filter first, from the probe / miss counts the scan
// kernels have accumulated in each filter so far.  Called when work orders are generated and again when a work order
// with several probe filters starts executing (so the order adapts between the work orders of one operator).
inline void RankLIPFiltersByMissRate(std::vector<qs_lip_ref> *refs) {
  if (refs->size() < 2) return;
  std::vector<std::pair<double, qs_lip_ref>> ranked;
  for (const qs_lip_ref &r : *refs) {
    std::uint64_t probes = 0, misses = 0;
    QS_CHECK_GPU(qsgpu_lip_probe_stats(r.lip, &probes, &misses));
    ranked.emplace_back(probes ? static_cast<double>(misses) / static_cast<double>(probes) : 0.0, r);
  }
  std::stable_sort(ranked.begin(), ranked.end(),
                   [](const std::pair<double, qs_lip_ref> &a, const std::pair<double, qs_lip_ref> &b) { return a.first > b.first; });
  for (std::size_t i = 0; i < refs->size(); ++i) (*refs)[i] = ranked[i].second;
}

class QueryContext {
 public:
  typedef std::uint32_t aggregation_state_id;
  typedef std::int32_t insert_destination_id;
  typedef std::uint32_t join_hash_table_id;
  typedef std::int32_t lip_deployment_id;
  typedef std::uint32_t lip_filter_id;
  typedef std::int32_t predicate_id;
  typedef std::int32_t scalar_group_id;
  typedef std::uint32_t sort_config_id;
  static constexpr insert_destination_id kInvalidInsertDestinationId = -1;
  static constexpr lip_deployment_id kInvalidLIPDeploymentId = -1;
  static constexpr predicate_id kInvalidPredicateId = -1;
  static constexpr scalar_group_id kInvalidScalarGroupId = -1;

  struct Predicate { ExprSet exprs; int root = -1; };
  struct ScalarGroup { ExprSet exprs; std::vector<int> roots; };
  // serialization::AggregationOperationState: aggregates, group-by, predicate, estimate, strategy
  struct AggregationSpec {
    ExprSet exprs;
    int predicate_root = -1;
    std::vector<qs_aggregate> aggregates;
    std::vector<int> group_by_roots;
    std::uint32_t strategy = QS_AGG_SINGLE_STATE;
    std::uint64_t estimated_num_entries = 1024;
    std::int64_t collision_free_max_key = -1;
    // bit j: aggregate j's argument has a NULL-able type (the reference creates the handle from the argument types,
    // AggregateFunctionSum::createHandle; NULL arguments are skipped, AggregationHandleSum.hpp:117-127)
    std::uint64_t nullable_arguments = 0;
    // Several devices: the input is partitioned on (a prefix of) the group-by attributes, so no group spans two
    // devices and every device finalizes its own groups -- what the reference does per partition when
    // `is_partitioned_on_group_by` holds (query_optimizer/ExecutionGenerator.cpp, aggregation state per partition).
    // false: the per-device partial states are merged (qsgpu_agg_merge_all) before FinalizeAggregation.
    bool partitioned_on_group_by = false;
  };
  // utility/lip_filter/LIPFilterDeployment.hpp: which filters an operator builds or probes, on which attribute
  // (LIPFilter.proto:50-62: a deployment carries build entries and probe entries)
  enum class LIPAction { kBuild, kProbe };
  struct LIPEntry { lip_filter_id filter; attribute_id attr; };
  struct LIPDeployment { std::vector<LIPEntry> build_entries, probe_entries; };
  // SortConfiguration (+ LIMIT) of SortMergeRunOperator's top-k path
  struct SortConfig { std::vector<qs_sort_key> keys; };

  QueryContext(StorageManager *sm, int device) : sm_(sm), device_(device) {}
  qsgpu_comm_t comm() const { return sm_->communicator(); }
  // ---- several devices: which per-device objects still hold only this device's share
  // A LIP filter filled from a partitioned input is all-reduced (bitwise OR over the devices) the first time
  // an operator asks for it as a probe filter; one filled from a replicated input is complete everywhere.
  void noteLIPFiltersBuiltFrom(lip_deployment_id id, bool input_partitioned) {
    const LIPDeployment *d = getLIPDeployment(id);
    if (!d) return;
    for (const LIPEntry &e : d->build_entries) lip_partial_[e.filter] = input_partitioned && sm_->multiDevice();
  }
  void noteAggregationInput(aggregation_state_id id, bool input_partitioned) { agg_partial_[id] = input_partitioned && sm_->multiDevice(); }
  bool aggregationIsPartial(aggregation_state_id id) const { return agg_partial_[id]; }
  ~QueryContext() {
    for (auto h : agg_states_) if (h) qsgpu_agg_destroy(h);
    for (auto h : join_tables_) if (h) qsgpu_join_destroy(h);
    for (auto h : lip_filters_) if (h) qsgpu_lip_destroy(h);
  }
  int device() const { return device_; }

  // ---- construction (what QueryContext::QueryContext does from the proto)
  predicate_id addPredicate(Predicate p) { predicates_.push_back(std::move(p)); return static_cast<predicate_id>(predicates_.size()) - 1; }
  scalar_group_id addScalarGroup(ScalarGroup g) { scalar_groups_.push_back(std::move(g)); return static_cast<scalar_group_id>(scalar_groups_.size()) - 1; }
  aggregation_state_id addAggregationState(AggregationSpec spec) {
    agg_specs_.push_back(std::move(spec));
    agg_states_.push_back(nullptr);
    agg_partial_.push_back(false);
    const AggregationSpec &s = agg_specs_.back();
    const qs_expr_set es = s.exprs.view();
    qs_agg_spec c{};
    c.dev = device_; c.strategy = s.strategy; c.exprs = &es; c.predicate_root = s.predicate_root;
    c.n_aggregates = static_cast<std::uint32_t>(s.aggregates.size()); c.aggregates = s.aggregates.data();
    c.n_group_by = static_cast<std::uint32_t>(s.group_by_roots.size()); c.group_by_roots = s.group_by_roots.data();
    c.estimated_num_entries = s.estimated_num_entries; c.collision_free_max_key = s.collision_free_max_key;
    c.nullable_arguments = s.nullable_arguments;
    QS_CHECK_GPU(qsgpu_agg_create(&c, &agg_states_.back()));
    return static_cast<aggregation_state_id>(agg_states_.size()) - 1;
  }
  join_hash_table_id addJoinHashTable(std::uint32_t key_type, std::uint64_t estimated_num_entries) {
    qsgpu_join_table_t t = nullptr;
    QS_CHECK_GPU(qsgpu_join_create(device_, key_type, estimated_num_entries, &t));
    join_tables_.push_back(t);
    return static_cast<join_hash_table_id>(join_tables_.size()) - 1;
  }
  lip_filter_id addLIPFilter(std::uint32_t kind, std::uint32_t attr_type, std::int64_t min_value, std::int64_t max_value,
                             std::uint64_t cardinality, bool is_anti) {
    qsgpu_lip_t f = nullptr;
    QS_CHECK_GPU(qsgpu_lip_create(device_, kind, attr_type, min_value, max_value, cardinality, is_anti ? 1 : 0, &f));
    lip_filters_.push_back(f);
    lip_partial_.push_back(false);
    return static_cast<lip_filter_id>(lip_filters_.size()) - 1;
  }
  lip_deployment_id addLIPDeployment(LIPDeployment d) { lip_deployments_.push_back(std::move(d)); return static_cast<lip_deployment_id>(lip_deployments_.size()) - 1; }
  insert_destination_id addInsertDestination(const CatalogRelation *rel, std::uint64_t capacity_rows) {
    destinations_.emplace_back(new InsertDestination(rel, capacity_rows, sm_));
    return static_cast<insert_destination_id>(destinations_.size()) - 1;
  }
  // serialization::InsertDestination with a partition scheme (query_execution/QueryContext.cpp:99-123 builds a
  // PartitionAwareInsertDestination for it)
  insert_destination_id addPartitionAwareInsertDestination(const HashPartitionSchemeHeader &header, const CatalogRelation *rel,
                                                           std::uint64_t capacity_rows) {
    destinations_.emplace_back(new PartitionAwareInsertDestination(header, rel, capacity_rows, sm_));
    return static_cast<insert_destination_id>(destinations_.size()) - 1;
  }
  sort_config_id addSortConfig(SortConfig c) { sort_configs_.push_back(std::move(c)); return static_cast<sort_config_id>(sort_configs_.size()) - 1; }

  // ---- getters (same names as the reference)
  const Predicate *getPredicate(predicate_id id) const { return id < 0 ? nullptr : &predicates_[static_cast<std::size_t>(id)]; }
  const ScalarGroup &getScalarGroup(scalar_group_id id) const { return scalar_groups_[static_cast<std::size_t>(id)]; }
  qsgpu_agg_state_t getAggregationState(aggregation_state_id id, partition_id = 0) const { return agg_states_[id]; }
  const AggregationSpec &getAggregationSpec(aggregation_state_id id) const { return agg_specs_[id]; }
  void destroyAggregationState(aggregation_state_id id, partition_id = 0) {
    if (agg_states_[id]) QS_CHECK_GPU(qsgpu_agg_destroy(agg_states_[id]));
    agg_states_[id] = nullptr;
  }
  qsgpu_join_table_t getJoinHashTable(join_hash_table_id id, partition_id = 0) const { return join_tables_[id]; }
  void destroyJoinHashTable(join_hash_table_id id, partition_id = 0) {
    if (join_tables_[id]) QS_CHECK_GPU(qsgpu_join_destroy(join_tables_[id]));
    join_tables_[id] = nullptr;
  }
  qsgpu_lip_t getLIPFilter(lip_filter_id id) const { return lip_filters_[id]; }
  const LIPDeployment *getLIPDeployment(lip_deployment_id id) const { return id < 0 ? nullptr : &lip_deployments_[static_cast<std::size_t>(id)]; }
  InsertDestination *getInsertDestination(insert_destination_id id) { return id < 0 ? nullptr : destinations_[static_cast<std::size_t>(id)].get(); }
  const SortConfig &getSortConfig(sort_config_id id) const { return sort_configs_[id]; }

  // LIPFilterUtil: the C-ABI references of a deployment (filter handle + attribute)
  // (Foreman thread only: called from getAllWorkOrders.)
  std::vector<qs_lip_ref> lipRefs(lip_deployment_id id, LIPAction action) {
    std::vector<qs_lip_ref> refs;
    const LIPDeployment *d = getLIPDeployment(id);
    if (d)
      for (const LIPEntry &e : (action == LIPAction::kBuild ? d->build_entries : d->probe_entries)) {
        if (action == LIPAction::kProbe && lip_partial_[e.filter]) {
          QS_CHECK_GPU(qsgpu_lip_allreduce(lip_filters_[e.filter], sm_->communicator()));
          lip_partial_[e.filter] = false;
        }
        qs_lip_ref r{}; r.lip = lip_filters_[e.filter]; r.attr = static_cast<std::uint32_t>(e.attr); refs.push_back(r);
      }
    // LIPFilterAdaptiveProber: with several probe filters, the ones that rejected the larger share of what they have
    // seen so far go first (the kernels skip rows an earlier filter rejected).  The statistics come from the work
    // orders executed so far -- of this or of any earlier operator probing the same filter -- so the order adapts
    // between the work orders of a streamed / row-capped operator.  Result-neutral.
    if (action == LIPAction::kProbe) RankLIPFiltersByMissRate(&refs);
    return refs;
  }

 private:
  StorageManager *sm_;
  int device_;
  std::vector<Predicate> predicates_;
  std::vector<ScalarGroup> scalar_groups_;
  std::vector<AggregationSpec> agg_specs_;
  std::vector<qsgpu_agg_state_t> agg_states_;
  std::vector<qsgpu_join_table_t> join_tables_;
  std::vector<qsgpu_lip_t> lip_filters_;
  std::vector<bool> lip_partial_, agg_partial_;
  std::vector<LIPDeployment> lip_deployments_;
  std::vector<std::unique_ptr<InsertDestination>> destinations_;
  std::vector<SortConfig> sort_configs_;
};

}  // namespace quickstep
