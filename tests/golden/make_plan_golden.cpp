// TEST INFRASTRUCTURE: generates tests/golden/reference_plans.json with the UNMODIFIED reference's parser, optimizer and
// ExecutionGenerator (linked from oracle/build_ref.sh's build tree; recipe: tests/golden/make_plan_golden.sh).
//
// north_star: "GPU work orders drop into the existing scheduling and the plans produced by ExecutionGenerator need no
// changes".  This program takes that literally: it parses benchmarks/tpch/queries/{01,03,06}.sql (read where they lie),
// lets the reference plan them over a catalog with the TPC-H attributes of the hot path and the statistics `\analyze`
// records at SF0.01, and then lowers what the ExecutionGenerator serialized -- the serialization::QueryContext: aggregation
// states, predicates, scalar groups, LIP filters and their deployments, join hash tables, sort configurations -- with the
// in-tree binding (quickstep_b200/host/intree/ProtoLowering.hpp, QueryContextLowering.hpp) into the C ABI's descriptions.
// The operator DAG is written next to it.  tests/test_reference_plans.py executes the lowered plans with the oracle over
// dbgen's SF0.01 relations and must get the answers the engine itself printed for the same SQL.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "catalog/CatalogAttribute.hpp"
#include "catalog/CatalogDatabase.hpp"
#include "catalog/CatalogRelation.hpp"
#include "catalog/CatalogRelationStatistics.hpp"
#include "parser/ParseStatement.hpp"
#include "parser/SqlParserWrapper.hpp"
#include "query_execution/QueryContext.pb.h"
#include "query_execution/WorkOrderProtosContainer.hpp"
#include "relational_operators/WorkOrder.pb.h"
#include "storage/InsertDestination.pb.h"
#include "storage/StorageBlockInfo.hpp"
#include "query_optimizer/Optimizer.hpp"
#include "query_optimizer/OptimizerContext.hpp"
#include "query_optimizer/QueryHandle.hpp"
#include "query_optimizer/QueryPlan.hpp"
#include "relational_operators/RelationalOperator.hpp"
#include "types/Type.hpp"
#include "types/TypeFactory.hpp"
#include "types/TypeID.hpp"
#include "types/TypedValue.hpp"
#include "utility/DAG.hpp"
#include "utility/SortConfiguration.pb.h"
#include "utility/lip_filter/LIPFilter.pb.h"

#include "ProtoLowering.hpp"
#include "QueryContextLowering.hpp"

using namespace quickstep;  // NOLINT

namespace {

std::string Hex(const void *p, std::size_t n) {
  static const char *d = "0123456789abcdef";
  std::string s;
  const unsigned char *b = static_cast<const unsigned char *>(p);
  for (std::size_t i = 0; i < n; ++i) {
    s.push_back(d[b[i] >> 4]);
    s.push_back(d[b[i] & 15]);
  }
  return s;
}

struct AttrDef {
  const char *name;
  TypeID type;
  std::size_t length;           // CHAR only
  std::int64_t distinct, min, max;   // what \analyze records (integers / row counts here; 0 = not recorded)
};

// The attributes of the hot path, in the order tests/tpch_data.py and quickstep_b200/tpch.py hold them (DECIMAL is DOUBLE,
// parser/SqlParser.ypp:791); row counts and key ranges of dbgen -s 0.01.
CatalogRelation *AddRelation(CatalogDatabase *db, const char *name, std::size_t num_tuples, const std::vector<AttrDef> &attrs) {
  CatalogRelation *rel = new CatalogRelation(db, name);
  for (const AttrDef &a : attrs) {
    const Type &t = a.type == kChar ? TypeFactory::GetType(kChar, a.length, false) : TypeFactory::GetType(a.type, false);
    rel->addAttribute(new CatalogAttribute(rel, a.name, t));
  }
  db->addRelation(rel);
  // one block id per stored relation (never dereferenced: no storage manager exists here), so that operators over stored
  // relations describe one work order
  rel->addBlock(BlockIdUtil::GetBlockId(1 /* domain */, 1 + rel->getID()));
  CatalogRelationStatistics *stats = rel->getStatisticsMutable();
  stats->setExactness(true);
  stats->setNumTuples(num_tuples);
  for (std::size_t i = 0; i < attrs.size(); ++i) {
    const AttrDef &a = attrs[i];
    if (a.distinct) stats->setNumDistinctValues(static_cast<attribute_id>(i), a.distinct);
    if (a.type == kInt && a.max) {
      stats->setMinValue(static_cast<attribute_id>(i), TypedValue(static_cast<int>(a.min)));
      stats->setMaxValue(static_cast<attribute_id>(i), TypedValue(static_cast<int>(a.max)));
    }
  }
  return rel;
}

void EmitNodes(FILE *out, const gpu::ExprBuilder &b) {
  const qs_expr_set es = b.view();
  std::fprintf(out, "\"pool\": \"%s\", \"nodes\": [", Hex(es.str_pool, es.str_pool_bytes).c_str());
  for (std::uint32_t i = 0; i < es.n_nodes; ++i) {
    const qs_node &n = es.nodes[i];
    std::fprintf(out, "%s[%u, %u, %u, %u, %d, %d, \"%s\"]", i ? ", " : "", n.kind, n.op, n.type, n.width, n.a, n.b,
                 Hex(&n.lit, sizeof(n.lit)).c_str());
  }
  std::fprintf(out, "]");
}

// every relation of the database (the temporaries the ExecutionGenerator created included): ScalarAttribute protos name
// (relation_id, attribute_id)
gpu::AttributeTypes AllTypes(const CatalogDatabase &db) {
  gpu::AttributeTypes types;
  for (CatalogDatabase::const_iterator it = db.begin(); it != db.end(); ++it)
    types.relations.emplace_back(it->getID(), gpu::AttributesOf(*it));
  return types;
}

void EmitRelations(FILE *out, const CatalogDatabase &db) {
  std::fprintf(out, "\"relations\": [");
  bool first = true;
  for (CatalogDatabase::const_iterator it = db.begin(); it != db.end(); ++it) {
    std::fprintf(out, "%s\n    {\"id\": %d, \"name\": \"%s\", \"temporary\": %s, \"attributes\": [", first ? "" : ",", it->getID(),
                 it->getName().c_str(), it->isTemporary() ? "true" : "false");
    const std::vector<qs_attr> attrs = gpu::AttributesOf(*it);
    std::size_t k = 0;
    for (CatalogRelationSchema::const_iterator a = it->begin(); a != it->end(); ++a, ++k)
      std::fprintf(out, "%s[\"%s\", %u, %u]", k ? ", " : "", a->getName().c_str(), attrs[k].type, attrs[k].width);
    std::fprintf(out, "]}");
    first = false;
  }
  std::fprintf(out, "]");
}

struct FullAttr {
  const char *name;
  TypeID type;
  std::size_t length;
  int key_of;        // 0: not an integer key; otherwise index into kRowsSf1 below of the relation whose key range it has
};

// benchmarks/tpch/create.sql's eight relations, all attributes, with the statistics `\\analyze` records at scale factor sf
void BuildFullCatalog(CatalogDatabase *db, double sf) {
  // rows at SF1: region, nation, supplier, customer, part, partsupp, orders, lineitem
  const double kRowsSf1[9] = {0, 5, 25, 10000, 150000, 200000, 800000, 1500000, 6000000};
  auto rows = [&](int rel) { return rel <= 2 ? static_cast<std::int64_t>(kRowsSf1[rel]) : static_cast<std::int64_t>(kRowsSf1[rel] * sf); };
  struct Rel { const char *name; int rows_of; std::vector<FullAttr> attrs; };
  const std::vector<Rel> schema = {
      {"region", 1, {{"r_regionkey", kInt, 0, 1}, {"r_name", kChar, 25, 0}, {"r_comment", kVarChar, 152, 0}}},
      {"nation", 2, {{"n_nationkey", kInt, 0, 2}, {"n_name", kChar, 25, 0}, {"n_regionkey", kInt, 0, 1}, {"n_comment", kVarChar, 152, 0}}},
      {"supplier", 3, {{"s_suppkey", kInt, 0, 3}, {"s_name", kChar, 25, 0}, {"s_address", kVarChar, 40, 0}, {"s_nationkey", kInt, 0, 2},
                       {"s_phone", kChar, 15, 0}, {"s_acctbal", kDouble, 0, 0}, {"s_comment", kVarChar, 101, 0}}},
      {"customer", 4, {{"c_custkey", kInt, 0, 4}, {"c_name", kVarChar, 25, 0}, {"c_address", kVarChar, 40, 0}, {"c_nationkey", kInt, 0, 2},
                       {"c_phone", kChar, 15, 0}, {"c_acctbal", kDouble, 0, 0}, {"c_mktsegment", kChar, 10, 0}, {"c_comment", kVarChar, 117, 0}}},
      {"part", 5, {{"p_partkey", kInt, 0, 5}, {"p_name", kVarChar, 55, 0}, {"p_mfgr", kChar, 25, 0}, {"p_brand", kChar, 10, 0}, {"p_type", kVarChar, 25, 0},
                   {"p_size", kInt, 0, 0}, {"p_container", kChar, 10, 0}, {"p_retailprice", kDouble, 0, 0}, {"p_comment", kVarChar, 23, 0}}},
      {"partsupp", 6, {{"ps_partkey", kInt, 0, 5}, {"ps_suppkey", kInt, 0, 3}, {"ps_availqty", kInt, 0, 0}, {"ps_supplycost", kDouble, 0, 0},
                       {"ps_comment", kVarChar, 199, 0}}},
      {"orders", 7, {{"o_orderkey", kInt, 0, 7}, {"o_custkey", kInt, 0, 4}, {"o_orderstatus", kChar, 1, 0}, {"o_totalprice", kDouble, 0, 0},
                     {"o_orderdate", kDate, 0, 0}, {"o_orderpriority", kChar, 15, 0}, {"o_clerk", kChar, 15, 0}, {"o_shippriority", kInt, 0, 0},
                     {"o_comment", kVarChar, 79, 0}}},
      {"lineitem", 8, {{"l_orderkey", kInt, 0, 7}, {"l_partkey", kInt, 0, 5}, {"l_suppkey", kInt, 0, 3}, {"l_linenumber", kInt, 0, 0},
                       {"l_quantity", kDouble, 0, 0}, {"l_extendedprice", kDouble, 0, 0}, {"l_discount", kDouble, 0, 0}, {"l_tax", kDouble, 0, 0},
                       {"l_returnflag", kChar, 1, 0}, {"l_linestatus", kChar, 1, 0}, {"l_shipdate", kDate, 0, 0}, {"l_commitdate", kDate, 0, 0},
                       {"l_receiptdate", kDate, 0, 0}, {"l_shipinstruct", kChar, 25, 0}, {"l_shipmode", kChar, 10, 0}, {"l_comment", kVarChar, 44, 0}}}};
  for (const Rel &r : schema) {
    CatalogRelation *rel = new CatalogRelation(db, r.name);
    for (const FullAttr &a : r.attrs) {
      const Type &t = a.length ? TypeFactory::GetType(a.type, a.length, false) : TypeFactory::GetType(a.type, false);
      rel->addAttribute(new CatalogAttribute(rel, a.name, t));
    }
    db->addRelation(rel);
    rel->addBlock(BlockIdUtil::GetBlockId(1, 1 + rel->getID()));
    CatalogRelationStatistics *stats = rel->getStatisticsMutable();
    stats->setExactness(true);
    stats->setNumTuples(rows(r.rows_of));
    for (std::size_t i = 0; i < r.attrs.size(); ++i) {
      if (!r.attrs[i].key_of) continue;
      const std::int64_t n = rows(r.attrs[i].key_of);
      const std::int64_t max_key = r.attrs[i].key_of == 7 ? n * 4 : n;            // dbgen's order keys are sparse
      stats->setNumDistinctValues(static_cast<attribute_id>(i), std::min<std::int64_t>(n, rows(r.rows_of)));
      stats->setMinValue(static_cast<attribute_id>(i), TypedValue(static_cast<int>(r.attrs[i].key_of <= 2 ? 0 : 1)));
      stats->setMaxValue(static_cast<attribute_id>(i), TypedValue(static_cast<int>(r.attrs[i].key_of <= 2 ? n - 1 : max_key)));
    }
  }
}

// `sf`: the scale factor whose `\\analyze` statistics the catalog carries (row counts, distinct counts and key ranges grow
// with it; dbgen's order keys are sparse: 8 of every 32).
void PlanQuery(FILE *out, const char *name, const std::string &sql, bool first_query, double sf = 0.01, bool full_catalog = false) {
  // a fresh catalog per query: relation ids of the temporaries start from the same point every time
  CatalogDatabase db(nullptr, "default");
  if (full_catalog) BuildFullCatalog(&db, sf);
  const std::int64_t n_orders = static_cast<std::int64_t>(1500000 * sf), n_cust = static_cast<std::int64_t>(150000 * sf);
  const std::int64_t n_lineitem = sf == 0.01 ? 60175 : static_cast<std::int64_t>(6000000 * sf), max_okey = n_orders * 4;
  if (!full_catalog) {
  AddRelation(&db, "lineitem", n_lineitem,
              {{"l_orderkey", kInt, 0, n_orders, 1, max_okey}, {"l_quantity", kDouble, 0, 50, 0, 0}, {"l_extendedprice", kDouble, 0, 35921, 0, 0},
               {"l_discount", kDouble, 0, 11, 0, 0}, {"l_tax", kDouble, 0, 9, 0, 0}, {"l_returnflag", kChar, 1, 3, 0, 0},
               {"l_linestatus", kChar, 1, 2, 0, 0}, {"l_shipdate", kDate, 0, 2518, 0, 0}});
  AddRelation(&db, "orders", n_orders,
              {{"o_orderkey", kInt, 0, n_orders, 1, max_okey}, {"o_custkey", kInt, 0, n_cust * 2 / 3, 1, n_cust - 1}, {"o_orderdate", kDate, 0, 2401, 0, 0},
               {"o_shippriority", kInt, 0, 1, 0, 0}});
  AddRelation(&db, "customer", n_cust, {{"c_custkey", kInt, 0, n_cust, 1, n_cust}, {"c_mktsegment", kChar, 10, 5, 0, 0}});
  }

  SqlParserWrapper parser;
  parser.feedNextBuffer(new std::string(sql));
  ParseResult result = parser.getNextStatement();
  CHECK(result.condition == ParseResult::kSuccess) << result.error_message;
  QueryHandle handle(1 /* query_id */, 0 /* cli_id */);
  optimizer::OptimizerContext context;
  optimizer::Optimizer optimizer;
  optimizer.generateQueryHandle(*result.parsed_statement, &db, &context, &handle);

  const serialization::QueryContext &qc = handle.getQueryContextProto();
  const gpu::AttributeTypes types = AllTypes(db);

  std::fprintf(out, "%s\n {\"query\": \"%s\", \"statistics_of_sf\": %g,\n  ", first_query ? "" : ",", name, sf);
  EmitRelations(out, db);

  // ---- the operator DAG, as the Foreman will see it (query_execution/QueryManagerBase.cpp)
  const DAG<RelationalOperator, bool> &dag = handle.getQueryPlanMutable()->getQueryPlanDAG();
  std::fprintf(out, ",\n  \"operators\": [");
  for (std::size_t i = 0; i < dag.size(); ++i) {
    const RelationalOperator &op = dag.getNodePayload(i);
    std::fprintf(out, "%s\n    {\"index\": %zu, \"name\": \"%s\", \"num_partitions\": %zu, \"output_relation\": %d, \"dependents\": [",
                 i ? "," : "", i, op.getName().c_str(), op.getNumPartitions(), static_cast<int>(op.getOutputRelationID()));
    bool f = true;
    for (const auto &d : dag.getDependents(i)) {
      std::fprintf(out, "%s[%zu, %s]", f ? "" : ", ", static_cast<std::size_t>(d.first), d.second ? "true" : "false");   // (consumer, pipeline-breaking)
      f = false;
    }
    std::fprintf(out, "]");
    // What the operator tells its work orders: the serialized form the reference's own (distributed) path ships to a
    // Shiftboss -- relation ids and the QueryContext indices of predicate / scalar group / state / table / deployment /
    // destination.  One block is "fed" for every temporary relation (operators keep only their own input's).
    RelationalOperator *mop = handle.getQueryPlanMutable()->getQueryPlanDAGMutable()->getNodePayloadMutable(i);
    mop->setOperatorIndex(i);
    std::string text;
    if (op.getName().find("DropTable") == std::string::npos && op.getName().find("SortMergeRun") == std::string::npos) {
      for (CatalogDatabase::const_iterator it = db.begin(); it != db.end(); ++it) {
        if (!it->isTemporary()) continue;
        mop->feedInputBlock(BlockIdUtil::GetBlockId(1, 100 + it->getID()), it->getID(), 0);
        mop->doneFeedingInputBlocks(it->getID());
      }
      WorkOrderProtosContainer protos(dag.size());
      mop->getAllWorkOrderProtos(&protos);
      if (protos.hasWorkOrderProto(i)) {
        std::unique_ptr<serialization::WorkOrder> wo(protos.getWorkOrderProto(i));
        text = wo->DebugString();
      }
      while (protos.hasWorkOrderProto(i)) delete protos.getWorkOrderProto(i);
    }
    std::string escaped;
    for (char c : text) {
      if (c == '\n') escaped += "\\n";
      else if (c == '"') escaped += "\\\"";
      else if (c == '\\') escaped += "\\\\";
      else escaped.push_back(c);
    }
    std::fprintf(out, ", \"work_order\": \"%s\"}", escaped.c_str());
  }
  std::fprintf(out, "]");

  // ---- insert destinations: which relation each one fills
  std::fprintf(out, ",\n  \"insert_destinations\": [");
  for (int i = 0; i < qc.insert_destinations_size(); ++i)
    std::fprintf(out, "%s{\"relation_id\": %d, \"relational_op_index\": %llu}", i ? ", " : "", qc.insert_destinations(i).relation_id(),
                 static_cast<unsigned long long>(qc.insert_destinations(i).relational_op_index()));
  std::fprintf(out, "]");

  // ---- aggregation states -> qs_agg_spec
  std::fprintf(out, ",\n  \"aggregation_states\": [");
  for (int i = 0; i < qc.aggregation_states_size(); ++i) {
    const serialization::AggregationOperationState &proto = qc.aggregation_states(i).aggregation_state();
    gpu::LoweredAggregationState lowered;
    gpu::LowerAggregationState(proto, db.getRelationSchemaById(proto.relation_id()), &lowered);
    std::fprintf(out, "%s\n    {\"relation_id\": %d, \"num_partitions\": %llu, \"hash_table_impl_type\": %d, \"strategy\": %u, "
                 "\"estimated_num_entries\": %llu, \"collision_free_max_key\": %lld, \"nullable_arguments\": %llu, \"predicate_root\": %d, ",
                 i ? "," : "", proto.relation_id(), static_cast<unsigned long long>(qc.aggregation_states(i).num_partitions()),
                 proto.has_hash_table_impl_type() ? static_cast<int>(proto.hash_table_impl_type()) : -1, lowered.strategy,
                 static_cast<unsigned long long>(lowered.estimated_num_entries), static_cast<long long>(lowered.collision_free_max_key),
                 static_cast<unsigned long long>(lowered.nullable_arguments), lowered.predicate_root);
    std::fprintf(out, "\"aggregates\": [");
    for (std::size_t j = 0; j < lowered.aggregates.size(); ++j)
      std::fprintf(out, "%s[%u, %d]", j ? ", " : "", lowered.aggregates[j].function, lowered.aggregates[j].argument_root);
    std::fprintf(out, "], \"group_by_roots\": [");
    for (std::size_t g = 0; g < lowered.group_by_roots.size(); ++g) std::fprintf(out, "%s%d", g ? ", " : "", lowered.group_by_roots[g]);
    std::fprintf(out, "], ");
    EmitNodes(out, lowered.exprs);
    std::fprintf(out, "}");
  }
  std::fprintf(out, "]");

  // ---- predicates and scalar groups -> node arrays
  std::fprintf(out, ",\n  \"predicates\": [");
  for (int i = 0; i < qc.predicates_size(); ++i) {
    gpu::ExprBuilder b;
    const int root = gpu::LowerPredicate(qc.predicates(i), types, &b);
    std::fprintf(out, "%s\n    {\"root\": %d, ", i ? "," : "", root);
    EmitNodes(out, b);
    std::fprintf(out, "}");
  }
  std::fprintf(out, "],\n  \"scalar_groups\": [");
  for (int i = 0; i < qc.scalar_groups_size(); ++i) {
    gpu::ExprBuilder b;
    std::fprintf(out, "%s\n    {\"roots\": [", i ? "," : "");
    std::vector<int> roots;
    for (int j = 0; j < qc.scalar_groups(i).scalars_size(); ++j) roots.push_back(gpu::LowerScalar(qc.scalar_groups(i).scalars(j), types, &b));
    for (std::size_t j = 0; j < roots.size(); ++j) std::fprintf(out, "%s%d", j ? ", " : "", roots[j]);
    std::fprintf(out, "], ");
    EmitNodes(out, b);
    std::fprintf(out, "}");
  }

  // ---- LIP filters and deployments -> qsgpu_lip_create arguments / qs_lip_ref lists
  std::fprintf(out, "],\n  \"lip_filters\": [");
  for (int i = 0; i < qc.lip_filters_size(); ++i) {
    const serialization::LIPFilter &f = qc.lip_filters(i);
    if (f.lip_filter_type() == serialization::LIPFilterType::BIT_VECTOR_EXACT_FILTER)
      std::fprintf(out, "%s\n    {\"kind\": %d, \"min_value\": %lld, \"max_value\": %lld, \"attribute_size\": %llu, \"is_anti\": %s}", i ? "," : "",
                   QS_LIP_BITVECTOR_EXACT, static_cast<long long>(f.GetExtension(serialization::BitVectorExactFilter::min_value)),
                   static_cast<long long>(f.GetExtension(serialization::BitVectorExactFilter::max_value)),
                   static_cast<unsigned long long>(f.GetExtension(serialization::BitVectorExactFilter::attribute_size)),
                   f.GetExtension(serialization::BitVectorExactFilter::is_anti_filter) ? "true" : "false");
    else if (f.lip_filter_type() == serialization::LIPFilterType::SINGLE_IDENTITY_HASH_FILTER)
      std::fprintf(out, "%s\n    {\"kind\": %d, \"cardinality\": %llu, \"attribute_size\": %llu}", i ? "," : "", QS_LIP_SINGLE_IDENTITY_HASH,
                   static_cast<unsigned long long>(f.GetExtension(serialization::SingleIdentityHashFilter::filter_cardinality)),
                   static_cast<unsigned long long>(f.GetExtension(serialization::SingleIdentityHashFilter::attribute_size)));
    else
      std::fprintf(out, "%s\n    {\"kind\": -1}", i ? "," : "");
  }
  std::fprintf(out, "],\n  \"lip_filter_deployments\": [");
  for (int i = 0; i < qc.lip_filter_deployments_size(); ++i) {
    const serialization::LIPFilterDeployment &d = qc.lip_filter_deployments(i);
    std::fprintf(out, "%s\n    {\"build\": [", i ? "," : "");
    for (int k = 0; k < d.build_entries_size(); ++k)
      std::fprintf(out, "%s[%u, %d]", k ? ", " : "", d.build_entries(k).lip_filter_id(), d.build_entries(k).attribute_id());
    std::fprintf(out, "], \"probe\": [");
    for (int k = 0; k < d.probe_entries_size(); ++k)
      std::fprintf(out, "%s[%u, %d]", k ? ", " : "", d.probe_entries(k).lip_filter_id(), d.probe_entries(k).attribute_id());
    std::fprintf(out, "]}");
  }

  // ---- join hash tables -> qsgpu_join_create arguments
  std::fprintf(out, "],\n  \"join_hash_tables\": [");
  for (int i = 0; i < qc.join_hash_tables_size(); ++i) {
    const serialization::HashTable &h = qc.join_hash_tables(i).join_hash_table();
    std::fprintf(out, "%s\n    {\"impl_type\": %d, \"estimated_num_entries\": %llu, \"num_partitions\": %llu, \"key_types\": [", i ? "," : "",
                 static_cast<int>(h.hash_table_impl_type()), static_cast<unsigned long long>(h.estimated_num_entries()),
                 static_cast<unsigned long long>(qc.join_hash_tables(i).num_partitions()));
    for (int k = 0; k < h.key_types_size(); ++k) {
      std::uint16_t w = 0;
      std::fprintf(out, "%s%u", k ? ", " : "", gpu::LowerTypeID(h.key_types(k), &w));
    }
    std::fprintf(out, "]}");
  }

  // ---- sort configurations -> qs_sort_key lists
  std::fprintf(out, "],\n  \"sort_configs\": [");
  for (int i = 0; i < qc.sort_configs_size(); ++i) {
    const serialization::SortConfiguration &sc = qc.sort_configs(i);
    std::fprintf(out, "%s\n    {\"keys\": [", i ? "," : "");
    for (int k = 0; k < sc.order_by_list_size(); ++k) {
      const serialization::SortConfiguration::OrderBy &ob = sc.order_by_list(k);
      CHECK(ob.expression().data_source() == serialization::Scalar::ATTRIBUTE) << "ORDER BY expressions are attributes of the sorted relation";
      std::fprintf(out, "%s{\"relation_id\": %d, \"attribute_id\": %d, \"ascending\": %s, \"null_first\": %s}", k ? ", " : "",
                   ob.expression().GetExtension(serialization::ScalarAttribute::relation_id),
                   ob.expression().GetExtension(serialization::ScalarAttribute::attribute_id), ob.is_ascending() ? "true" : "false",
                   ob.null_first() ? "true" : "false");
    }
    std::fprintf(out, "], \"qs_sort_keys\": [");
    const std::vector<qs_sort_key> lowered = gpu::LowerSortConfiguration(sc);
    for (std::size_t k = 0; k < lowered.size(); ++k) std::fprintf(out, "%s[%u, %u]", k ? ", " : "", lowered[k].attr, lowered[k].descending);
    std::fprintf(out, "]}");
  }
  std::fprintf(out, "]}");
}

std::string ReadFile(const std::string &path) {
  std::ifstream in(path);
  CHECK(in.good()) << path;
  std::stringstream ss;
  ss << in.rdbuf();
  return ss.str();
}


// ---------------------------------------------------------------------------------------------------------------------
// Coverage mode (tools/tpch_plan_coverage.py): ONE query of benchmarks/tpch/queries planned over the full TPC-H catalog
// (benchmarks/tpch/create.sql's eight relations, all attributes) with the statistics of scale factor `sf`; prints the
// operator list, then lowers every entry of the QueryContext.  What the device path does not take (LIKE, CASE, SUBSTRING,
// EXTRACT, VARCHAR attributes, DISTINCT aggregates, ...) ends in the binding's LOG(FATAL) naming the reason -- such an
// operator keeps its CPU work orders -- and the driver script records it.
// ---------------------------------------------------------------------------------------------------------------------
int Coverage(const std::string &ref, const std::string &query_file, double sf) {
  CatalogDatabase db(nullptr, "default");
  BuildFullCatalog(&db, sf);
  SqlParserWrapper parser;
  parser.feedNextBuffer(new std::string(ReadFile(ref + "/benchmarks/tpch/queries/" + query_file)));
  ParseResult result = parser.getNextStatement();
  CHECK(result.condition == ParseResult::kSuccess) << result.error_message;
  QueryHandle handle(1, 0);
  optimizer::OptimizerContext context;
  optimizer::Optimizer optimizer;
  optimizer.generateQueryHandle(*result.parsed_statement, &db, &context, &handle);
  const serialization::QueryContext &qc = handle.getQueryContextProto();
  const DAG<RelationalOperator, bool> &dag = handle.getQueryPlanMutable()->getQueryPlanDAG();
  std::printf("operators:");
  for (std::size_t i = 0; i < dag.size(); ++i) std::printf(" %s", dag.getNodePayload(i).getName().c_str());
  std::printf("\ncounts: aggregation_states=%d predicates=%d scalar_groups=%d lip_filters=%d join_hash_tables=%d sort_configs=%d\n",
              qc.aggregation_states_size(), qc.predicates_size(), qc.scalar_groups_size(), qc.lip_filters_size(), qc.join_hash_tables_size(),
              qc.sort_configs_size());
  std::fflush(stdout);
  const gpu::AttributeTypes types = AllTypes(db);
  for (int i = 0; i < qc.predicates_size(); ++i) {
    gpu::ExprBuilder b;
    gpu::LowerPredicate(qc.predicates(i), types, &b);
  }
  std::printf("lowered: predicates\n");
  std::fflush(stdout);
  for (int i = 0; i < qc.scalar_groups_size(); ++i) {
    gpu::ExprBuilder b;
    for (int j = 0; j < qc.scalar_groups(i).scalars_size(); ++j) gpu::LowerScalar(qc.scalar_groups(i).scalars(j), types, &b);
  }
  std::printf("lowered: scalar_groups\n");
  std::fflush(stdout);
  for (int i = 0; i < qc.aggregation_states_size(); ++i) {
    const serialization::AggregationOperationState &proto = qc.aggregation_states(i).aggregation_state();
    gpu::LoweredAggregationState lowered;
    gpu::LowerAggregationState(proto, db.getRelationSchemaById(proto.relation_id()), &lowered);
    for (const std::int32_t g : lowered.group_by_roots)
      CHECK(lowered.exprs.node(g).kind == QS_N_ATTRIBUTE) << "GPU path: GROUP BY expressions other than attributes";
  }
  std::printf("lowered: aggregation_states\n");
  for (int i = 0; i < qc.lip_filters_size(); ++i)
    CHECK(qc.lip_filters(i).lip_filter_type() != serialization::LIPFilterType::BLOOM_FILTER) << "GPU path: BLOOM_FILTER";
  for (int i = 0; i < qc.join_hash_tables_size(); ++i) {
    const serialization::HashTable &h = qc.join_hash_tables(i).join_hash_table();
    CHECK_LE(h.key_types_size(), 2) << "GPU path: join keys of more than two attributes";
    for (int k = 0; k < h.key_types_size(); ++k)
      CHECK(h.key_types(k).type_id() == serialization::Type::INT || (h.key_types_size() == 1 && h.key_types(k).type_id() == serialization::Type::LONG))
          << "GPU path: join key that is not INT / LONG";
  }
  std::printf("lowered: lip_filters join_hash_tables\nok\n");
  return 0;
}

}  // namespace

int main(int argc, char **argv) {
  if (argc >= 5 && std::string(argv[1]) == "--coverage") return Coverage(argv[2], argv[3], std::atof(argv[4]));
  if (argc >= 6 && std::string(argv[1]) == "--plan") {      // --plan <reference root> <out.json> <sf> NN ...: full catalog
    FILE *out = std::fopen(argv[3], "w");
    std::fprintf(out, "{\"generator\": \"tests/golden/make_plan_golden.cpp --plan: full TPC-H catalog, statistics of SF%s\",\n \"plans\": [", argv[4]);
    for (int i = 5; i < argc; ++i) {
      const std::string q = argv[i];
      PlanQuery(out, ("q" + std::to_string(std::atoi(q.c_str()))).c_str(), ReadFile(std::string(argv[2]) + "/benchmarks/tpch/queries/" + q + ".sql"), i == 5,
                std::atof(argv[4]), true);
    }
    std::fprintf(out, "\n ]}\n");
    std::fclose(out);
    return 0;
  }
  CHECK_GE(argc, 3) << "usage: make_plan_golden <reference root> <out.json>";
  const std::string ref = argv[1];
  FILE *out = std::fopen(argv[2], "w");
  std::fprintf(out, "{\"generator\": \"tests/golden/make_plan_golden.cpp: parser + optimizer + ExecutionGenerator of the unmodified reference (fee4c630), "
                    "QueryContext lowered by quickstep_b200/host/intree\",\n \"plans\": [");
  PlanQuery(out, "q1", ReadFile(ref + "/benchmarks/tpch/queries/01.sql"), true);
  PlanQuery(out, "q6", ReadFile(ref + "/benchmarks/tpch/queries/06.sql"), false);
  PlanQuery(out, "q3", ReadFile(ref + "/benchmarks/tpch/queries/03.sql"), false);
  // the same SQL planned over the statistics of the benchmarked scale factors: what the hand-built operator DAGs of
  // quickstep_b200/host/TpchPlans.cpp (bench.py's queries) have to look like
  PlanQuery(out, "q3_sf10", ReadFile(ref + "/benchmarks/tpch/queries/03.sql"), false, 10.0);
  PlanQuery(out, "q3_sf100", ReadFile(ref + "/benchmarks/tpch/queries/03.sql"), false, 100.0);
  PlanQuery(out, "q1_sf100", ReadFile(ref + "/benchmarks/tpch/queries/01.sql"), false, 100.0);
  PlanQuery(out, "q6_sf100", ReadFile(ref + "/benchmarks/tpch/queries/06.sql"), false, 100.0);
  std::fprintf(out, "\n ]}\n");
  std::fclose(out);
  return 0;
}
