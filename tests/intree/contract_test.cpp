// TEST INFRASTRUCTURE: the in-tree GPU operators' scheduling contract, run against the reference's REAL classes.
//
// tests/test_intree_boundary.py compiles this file together with quickstep_b200/host/intree/GpuWorkOrders.cpp and
// GpuJoinWorkOrders.cpp (full code generation this time, not -fsyntax-only), links it with the static libraries of the
// unmodified reference's build tree and with quickstep_b200/lib/libqsgpu.so -- every qsgpu_* symbol the binding calls
// resolves -- and runs it.  No work order is executed (there is no device here); what runs is what the Foreman calls
// (query_execution/QueryManagerSingleNode.cpp:100-135, QueryManagerBase.cpp): getAllWorkOrders / feedInputBlock /
// doneFeedingInputBlocks against a real WorkOrdersContainer and a real QueryContext whose predicates and scalar groups are
// the ones the reference's optimizer produced for TPC-H Q3.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "../../quickstep_b200/host/intree/GpuWorkOrders.cpp"
#include "../../quickstep_b200/host/intree/GpuJoinWorkOrders.cpp"

#include "catalog/CatalogDatabase.hpp"
#include "parser/ParseStatement.hpp"
#include "parser/SqlParserWrapper.hpp"
#include "query_optimizer/Optimizer.hpp"
#include "query_optimizer/OptimizerContext.hpp"
#include "query_optimizer/QueryHandle.hpp"
#include "types/TypeFactory.hpp"

using namespace quickstep;  // NOLINT

#define EXPECT(cond)                                                              \
  do {                                                                            \
    if (!(cond)) {                                                                \
      std::fprintf(stderr, "%s:%d: EXPECT(%s) failed\n", __FILE__, __LINE__, #cond); \
      std::exit(1);                                                               \
    }                                                                             \
  } while (0)

namespace {

CatalogRelation *AddRelation(CatalogDatabase *db, const char *name, const std::vector<std::pair<const char *, const Type *>> &attrs) {
  CatalogRelation *rel = new CatalogRelation(db, name);
  for (const auto &a : attrs) rel->addAttribute(new CatalogAttribute(rel, a.first, *a.second));
  db->addRelation(rel);
  for (int b = 0; b < 3; ++b) rel->addBlock(BlockIdUtil::GetBlockId(1, 10 * (1 + rel->getID()) + b));
  rel->getStatisticsMutable()->setNumTuples(1000);
  return rel;
}

}  // namespace

int main(int argc, char **argv) {
  EXPECT(argc >= 2);
  std::ifstream in(std::string(argv[1]) + "/benchmarks/tpch/queries/03.sql");
  EXPECT(in.good());
  std::stringstream sql;
  sql << in.rdbuf();

  const Type &I = TypeFactory::GetType(kInt, false), &D = TypeFactory::GetType(kDouble, false), &DT = TypeFactory::GetType(kDate, false);
  CatalogDatabase db(nullptr, "default");
  CatalogRelation *lineitem = AddRelation(&db, "lineitem", {{"l_orderkey", &I}, {"l_quantity", &D}, {"l_extendedprice", &D}, {"l_discount", &D},
                                                             {"l_tax", &D}, {"l_returnflag", &TypeFactory::GetType(kChar, 1, false)},
                                                             {"l_linestatus", &TypeFactory::GetType(kChar, 1, false)}, {"l_shipdate", &DT}});
  AddRelation(&db, "orders", {{"o_orderkey", &I}, {"o_custkey", &I}, {"o_orderdate", &DT}, {"o_shippriority", &I}});
  AddRelation(&db, "customer", {{"c_custkey", &I}, {"c_mktsegment", &TypeFactory::GetType(kChar, 10, false)}});

  // the reference plans Q3; its predicates and scalar groups become a REAL QueryContext (the entries that need a storage
  // manager -- states, tables, destinations -- are left out: nothing below touches them)
  SqlParserWrapper parser;
  parser.feedNextBuffer(new std::string(sql.str()));
  ParseResult parsed = parser.getNextStatement();
  EXPECT(parsed.condition == ParseResult::kSuccess);
  QueryHandle handle(1, 0);
  optimizer::OptimizerContext optimizer_context;
  optimizer::Optimizer optimizer;
  optimizer.generateQueryHandle(*parsed.parsed_statement, &db, &optimizer_context, &handle);
  const serialization::QueryContext &planned = handle.getQueryContextProto();
  serialization::QueryContext trimmed;
  trimmed.set_query_id(1);
  for (int i = 0; i < planned.predicates_size(); ++i) *trimmed.add_predicates() = planned.predicates(i);
  for (int i = 0; i < planned.scalar_groups_size(); ++i) *trimmed.add_scalar_groups() = planned.scalar_groups(i);
  EXPECT(trimmed.predicates_size() == 3 && trimmed.scalar_groups_size() >= 3);
  QueryContext query_context(trimmed, db, nullptr, 0, nullptr);
  EXPECT(query_context.getPredicate(0) != nullptr);

  CatalogRelation *t0 = new CatalogRelation(&db, "t0", -1, true);
  t0->addAttribute(new CatalogAttribute(t0, "l_orderkey", I));
  t0->addAttribute(new CatalogAttribute(t0, "l_extendedprice", D));
  t0->addAttribute(new CatalogAttribute(t0, "l_discount", D));
  db.addRelation(t0);

  const std::size_t kOps = 4;
  WorkOrdersContainer container(kOps, 1);

  // ---- Select over a STORED relation, simple projection (the optimizer's first operator of Q3): one coarse work order,
  // generated once, whatever the number of blocks
  gpu::GpuSelectOperator select_stored(1, *lineitem, false, *t0, 0, 0 /* l_shipdate > ... */, std::vector<attribute_id>{0, 2, 3}, true,
                                       nullptr, nullptr);
  select_stored.setOperatorIndex(0);
  EXPECT(select_stored.getOperatorType() == RelationalOperator::kSelect);
  EXPECT(select_stored.getOutputRelationID() == t0->getID() && select_stored.getInsertDestinationID() == 0);
  EXPECT(select_stored.getAllWorkOrders(&container, &query_context, nullptr, 0, nullptr) == true);
  EXPECT(container.getNumNormalWorkOrders(0) == 1);
  EXPECT(select_stored.getAllWorkOrders(&container, &query_context, nullptr, 0, nullptr) == true);
  EXPECT(container.getNumNormalWorkOrders(0) == 1);
  std::unique_ptr<WorkOrder> wo(container.getNormalWorkOrder(0));
  EXPECT(wo != nullptr && wo->getQueryID() == 1 && dynamic_cast<gpu::GpuSelectWorkOrder *>(wo.get()) != nullptr);

  // ---- Select over a STREAMED relation, scalar group: nothing before the first block, one work order per batch of
  // blocks fed, finished only after doneFeedingInputBlocks
  gpu::GpuSelectOperator select_streamed(1, *t0, false, *t0, 0, QueryContext::kInvalidPredicateId, 0, false, nullptr, nullptr);
  select_streamed.setOperatorIndex(1);
  EXPECT(select_streamed.getAllWorkOrders(&container, &query_context, nullptr, 0, nullptr) == false);
  EXPECT(container.getNumNormalWorkOrders(1) == 0);
  select_streamed.feedInputBlock(BlockIdUtil::GetBlockId(1, 900), t0->getID(), 0);
  select_streamed.feedInputBlock(BlockIdUtil::GetBlockId(1, 901), t0->getID(), 0);
  select_streamed.feedInputBlock(BlockIdUtil::GetBlockId(1, 902), lineitem->getID(), 0);      // not its input: ignored
  select_streamed.noteDeviceRows(2000);
  EXPECT(select_streamed.getAllWorkOrders(&container, &query_context, nullptr, 0, nullptr) == false);
  EXPECT(container.getNumNormalWorkOrders(1) == 1);                                           // two blocks, ONE work order
  EXPECT(select_streamed.getAllWorkOrders(&container, &query_context, nullptr, 0, nullptr) == false);
  EXPECT(container.getNumNormalWorkOrders(1) == 1);
  select_streamed.feedInputBlock(BlockIdUtil::GetBlockId(1, 903), t0->getID(), 0);
  select_streamed.noteDeviceRows(2500);
  select_streamed.doneFeedingInputBlocks(t0->getID());
  EXPECT(select_streamed.getAllWorkOrders(&container, &query_context, nullptr, 0, nullptr) == true);
  EXPECT(container.getNumNormalWorkOrders(1) == 2);
  EXPECT(select_streamed.getAllWorkOrders(&container, &query_context, nullptr, 0, nullptr) == true);
  EXPECT(container.getNumNormalWorkOrders(1) == 2);

  // ---- Aggregation, stored and streamed
  gpu::GpuAggregationOperator agg_stored(1, *lineitem, true, 0, 1, nullptr, nullptr);
  agg_stored.setOperatorIndex(2);
  EXPECT(agg_stored.getOperatorType() == RelationalOperator::kAggregation);
  EXPECT(agg_stored.getAllWorkOrders(&container, &query_context, nullptr, 0, nullptr) == true);
  EXPECT(agg_stored.getAllWorkOrders(&container, &query_context, nullptr, 0, nullptr) == true);
  EXPECT(container.getNumNormalWorkOrders(2) == 1);
  gpu::GpuAggregationOperator agg_streamed(1, *t0, false, 0, 1, nullptr, nullptr);
  agg_streamed.setOperatorIndex(3);
  EXPECT(agg_streamed.getAllWorkOrders(&container, &query_context, nullptr, 0, nullptr) == false);
  EXPECT(container.getNumNormalWorkOrders(3) == 0);
  agg_streamed.feedInputBlock(BlockIdUtil::GetBlockId(1, 910), t0->getID(), 0);
  agg_streamed.noteDeviceRows(700);
  EXPECT(agg_streamed.getAllWorkOrders(&container, &query_context, nullptr, 0, nullptr) == false);
  EXPECT(container.getNumNormalWorkOrders(3) == 1);
  agg_streamed.doneFeedingInputBlocks(t0->getID());
  EXPECT(agg_streamed.getAllWorkOrders(&container, &query_context, nullptr, 0, nullptr) == true);
  EXPECT(container.getNumNormalWorkOrders(3) == 1);

  // drain what is left (a WorkOrdersContainer complains about pending work orders when destroyed)
  for (std::size_t op = 0; op < kOps; ++op)
    while (container.hasNormalWorkOrder(op)) delete container.getNormalWorkOrder(op);

  // ---- the lowering the work orders run inside execute(), on the reference's reconstructed objects
  {
    gpu::AttributeTypes types;
    types.relations.emplace_back(lineitem->getID(), gpu::SchemaOf(*lineitem));
    gpu::ExprBuilder b;
    const int root = gpu::LowerPredicate(query_context.getPredicate(0)->getProto(), types, &b);
    const qs_expr_set es = b.view();
    EXPECT(root == 2 && es.n_nodes == 3 && es.nodes[2].kind == QS_N_COMPARISON && es.nodes[2].op == QS_GT);     // l_shipdate > DATE '1995-03-15'
    EXPECT(es.nodes[0].kind == QS_N_ATTRIBUTE && es.nodes[0].a == 7 && es.nodes[0].type == QS_DATE);
    EXPECT(es.nodes[1].kind == QS_N_LITERAL && es.nodes[1].lit.date.year == 1995 && es.nodes[1].lit.date.month == 3 && es.nodes[1].lit.date.day == 15);
  }
  // ---- every other in-tree class is instantiated too, so that its code -- and with it every qsgpu_* call of the binding --
  // is generated and has to link.  Never executed: these need a device.
  if (argc > 100) {
    gpu::GpuQueryState state(0, planned);
    state.addAggregationState(0, planned.aggregation_states(0).aggregation_state(), *t0, 1);
    std::vector<attribute_id> key{0};
    const gpu::DeviceExtent extent;
    WorkOrder *never[] = {
        new gpu::GpuBuildHashWorkOrder(1, *t0, key, false, 0, extent, nullptr, state.joinHashTable(0, 0), state.lipRefs(0, gpu::GpuQueryState::kBuild)),
        new gpu::GpuHashJoinWorkOrder(1, *t0, *lineitem, key, false, 0, extent, nullptr, query_context.getScalarGroup(0),
                                      HashJoinOperator::JoinType::kInnerJoin, state.joinHashTable(0, 0), nullptr, {}),
        new gpu::GpuBuildLIPFilterWorkOrder(1, *t0, 0, extent, nullptr, {}, {}),
        new gpu::GpuFinalizeAggregationWorkOrder(1, 0, state.aggregationState(0, 0), *t0, 256, nullptr, nullptr),
        new gpu::GpuDestroyAggregationStateWorkOrder(1, 0, 0, &state), new gpu::GpuDestroyHashWorkOrder(1, 0, 0, &state)};
    for (WorkOrder *w : never) w->execute();
    RelationalOperator *never_ops[] = {
        new gpu::GpuBuildHashOperator(1, *t0, false, key, false, 1, 0, QueryContext::kInvalidPredicateId, &state, {extent}),
        new gpu::GpuHashJoinOperator(1, *t0, *lineitem, true, key, false, 1, false, *t0, 0, 0, QueryContext::kInvalidPredicateId, 0,
                                     HashJoinOperator::JoinType::kLeftSemiJoin, &state, {extent}, nullptr),
        new gpu::GpuBuildLIPFilterOperator(1, *t0, QueryContext::kInvalidPredicateId, false, &state, {extent}),
        new gpu::GpuFinalizeAggregationOperator(1, 0, 1, false, 1, *t0, 0, &state, 256),
        new gpu::GpuSortMergeRunOperator(1, *t0, *t0, 0, *t0, 0, 0, 128, 10, false, gpu::LowerSortConfiguration(planned.sort_configs(0)), nullptr, 1000)};
    for (RelationalOperator *o : never_ops) o->getAllWorkOrders(&container, &query_context, nullptr, 0, nullptr);
  }
  std::printf("in-tree contract ok\n");
  return 0;
}
