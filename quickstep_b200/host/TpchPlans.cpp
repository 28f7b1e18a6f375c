// TPC-H Q1 / Q6 / Q3 as the operator DAGs the reference's optimizer produces for them
// (benchmarks/tpch/queries/{01,06,03}.sql), built the way
// ExecutionGenerator does: QueryContext entries first (predicates, scalar groups, aggregation
// states, hash tables, LIP filters + deployments, insert destinations), then operators, then
// dependency edges.  Also the C entry points of libqshost.so (include/qshost.h).
#include <algorithm>
#include <array>
#include <cstring>
#include <memory>

#include <execinfo.h>
#include <signal.h>
#include <unistd.h>

#include "Operators.hpp"
#include "QueryManager.hpp"
#include "qshost.h"

using namespace quickstep;

namespace {

const qs_attr kInt{QS_INT, 4}, kLong{QS_LONG, 8}, kDouble{QS_DOUBLE, 8}, kDate{QS_DATE, 8}, kChar1{QS_CHAR, 1},
    kChar10{QS_CHAR, 10};

std::vector<CatalogAttribute> customerSchema() { return {{"c_custkey", kInt}, {"c_mktsegment", kChar10}}; }
std::vector<CatalogAttribute> ordersSchema() {
  return {{"o_orderkey", kInt}, {"o_custkey", kInt}, {"o_orderdate", kDate}, {"o_shippriority", kInt}};
}
std::vector<CatalogAttribute> lineitemSchema() {
  return {{"l_orderkey", kInt}, {"l_quantity", kDouble}, {"l_extendedprice", kDouble}, {"l_discount", kDouble},
          {"l_tax", kDouble}, {"l_returnflag", kChar1}, {"l_linestatus", kChar1}, {"l_shipdate", kDate}};
}
enum { L_ORDERKEY = 0, L_QUANTITY, L_EXTENDEDPRICE, L_DISCOUNT, L_TAX, L_RETURNFLAG, L_LINESTATUS, L_SHIPDATE };
enum { O_ORDERKEY = 0, O_CUSTKEY, O_ORDERDATE, O_SHIPPRIORITY };
enum { C_CUSTKEY = 0, C_MKTSEGMENT };

std::vector<CatalogAttribute> anon(std::initializer_list<qs_attr> types) {
  std::vector<CatalogAttribute> v;
  int i = 0;
  for (qs_attr t : types) v.push_back({"a" + std::to_string(i++), t});
  return v;
}

}  // namespace

struct qshost_db {
  int dev = 0;
  std::unique_ptr<StorageManager> sm;
  std::unique_ptr<WorkerPool> workers;
  std::unique_ptr<CatalogRelation> rel[3];
  std::uint64_t rows[3] = {0, 0, 0};          // this rank's rows
  std::uint64_t global_rows[3] = {0, 0, 0};   // rows of the whole relation (== rows[] on one device)
  qsgpu_comm_t comm = nullptr;
  // Several devices: orders and lineitem are partitioned on the order key by the same scheme (every rank holds the
  // lineitem rows of exactly the orders it holds), so Q3's join can run partition-wise.  0 = use that (default),
  // 1 = ignore it and broadcast the build side (what a build side partitioned any other way needs).
  int join_mode = 0;
  // what \analyze records and AttachLIPFilters / InjectJoinFilters read (exact min/max statistics)
  std::int64_t c_custkey_min = 0, c_custkey_max = 0, o_orderkey_min = 0, o_orderkey_max = 0;
  std::string last_profile;
  relation_id next_relation_id = 100;
  std::size_t next_query_id = 1;

  // temporary relations of the running query
  std::vector<std::unique_ptr<CatalogRelation>> temps;
  CatalogRelation *temp(std::vector<CatalogAttribute> attrs) {
    temps.emplace_back(new CatalogRelation(next_relation_id++, "tmp", std::move(attrs), true));
    return temps.back().get();
  }
  // the result relation of the last query: its host blocks (InsertDestination::bulkInsertTuples) stay readable
  // (qshost_result_block) until the next query starts
  std::unique_ptr<CatalogRelation> last_result;
  void dropTemps(const CatalogRelation *keep = nullptr) {
    if (last_result) { sm->dropTemporary(*last_result); last_result.reset(); }
    for (auto &t : temps) {
      if (t.get() == keep) last_result = std::move(t);
      else sm->dropTemporary(*t);
    }
    temps.clear();
  }
};

extern "C" {

static void backtraceOnSegv(int sig) {
  void *frames[64];
  const int n = backtrace(frames, 64);
  backtrace_symbols_fd(frames, n, 2);
  signal(sig, SIG_DFL);
  raise(sig);
}

int qshost_db_create(int dev, int num_workers, qshost_db_t *out) {
  if (std::getenv("QSHOST_BACKTRACE")) signal(SIGSEGV, backtraceOnSegv);
  const int st = qsgpu_init(1, &dev);
  if (st != 0) return st;
  std::unique_ptr<qshost_db> db(new qshost_db);
  db->dev = dev;
  db->sm.reset(new StorageManager(dev));
  db->workers.reset(new WorkerPool(num_workers));
  *out = db.release();
  return 0;
}

int qshost_db_destroy(qshost_db_t db) {
  if (!db) return 0;
  db->dropTemps();
  db->dropTemps();     // ... and the result relation kept from the last query
  delete db;
  return 0;
}

int qshost_db_set_comm(qshost_db_t db, void *comm) {
  db->comm = static_cast<qsgpu_comm_t>(comm);
  db->sm->setCommunicator(db->comm);
  return 0;
}

int qshost_db_set_join_mode(qshost_db_t db, int mode) {
  if (mode < 0 || mode > 1) return QSGPU_ERR_INVALID;
  db->join_mode = mode;
  return 0;
}

int qshost_result_block(qshost_db_t db, uint32_t index, const void **memory, uint64_t *bytes, uint64_t *n_tuples) {
  if (!db->last_result) return QSGPU_ERR_INVALID;
  const std::vector<block_id> blocks = db->sm->hostBlocksOf(*db->last_result);
  if (index >= blocks.size()) return QSGPU_ERR_INVALID;
  const StorageBlock &B = db->sm->getBlock(blocks[index]);
  if (memory) *memory = B.memory;
  if (bytes) *bytes = B.size;
  if (n_tuples) *n_tuples = static_cast<uint64_t>(B.num_tuples);
  return 0;
}

int qshost_set_rows_per_workorder(uint64_t rows) { FLAGS_gpu_rows_per_workorder = rows; return 0; }

int qshost_db_load(qshost_db_t db, int which, const void *const *columns, uint64_t n_rows, uint64_t rows_per_block,
                   int layout) {
  if (which < 0 || which > 2 || layout < 0 || layout > 2) return QSGPU_ERR_INVALID;
  static const char *names[3] = {"customer", "orders", "lineitem"};
  if (db->rel[which]) db->sm->evict(*db->rel[which]);      // replaced: the old blocks stay owned by the manager
  std::vector<CatalogAttribute> schema = which == QSHOST_CUSTOMER ? customerSchema() : which == QSHOST_ORDERS ? ordersSchema() : lineitemSchema();
  db->rel[which].reset(new CatalogRelation(db->next_relation_id++, names[which], schema));
  std::vector<const void *> cols(columns, columns + schema.size());
  db->sm->loadRelation(db->rel[which].get(), cols, n_rows, rows_per_block, static_cast<TupleStoreLayout>(layout));
  db->rows[which] = n_rows;
  auto minmax = [&](const void *col, std::int64_t *mn, std::int64_t *mx) {
    const std::int32_t *v = static_cast<const std::int32_t *>(col);
    *mn = n_rows ? v[0] : 0; *mx = n_rows ? v[0] : 0;
    for (std::uint64_t i = 1; i < n_rows; ++i) { *mn = std::min<std::int64_t>(*mn, v[i]); *mx = std::max<std::int64_t>(*mx, v[i]); }
  };
  if (which == QSHOST_CUSTOMER) minmax(columns[C_CUSTKEY], &db->c_custkey_min, &db->c_custkey_max);
  if (which == QSHOST_ORDERS) minmax(columns[O_ORDERKEY], &db->o_orderkey_min, &db->o_orderkey_max);
  db->global_rows[which] = n_rows;
  if (db->sm->multiDevice()) {
    // this rank loaded its partition: \analyze's statistics are those of the whole relation
    db->sm->setPartitioned(db->rel[which]->getID(), true);
    std::int64_t total = static_cast<std::int64_t>(n_rows);
    QS_CHECK_GPU(qsgpu_comm_allreduce_i64(db->comm, &total, 1, 0));
    db->global_rows[which] = static_cast<std::uint64_t>(total);
    if (which != QSHOST_LINEITEM) {
      std::int64_t *mn = which == QSHOST_CUSTOMER ? &db->c_custkey_min : &db->o_orderkey_min;
      std::int64_t *mx = which == QSHOST_CUSTOMER ? &db->c_custkey_max : &db->o_orderkey_max;
      if (n_rows == 0) { *mn = INT64_MAX; *mx = INT64_MIN; }
      QS_CHECK_GPU(qsgpu_comm_allreduce_i64(db->comm, mn, 1, 1));
      QS_CHECK_GPU(qsgpu_comm_allreduce_i64(db->comm, mx, 1, 2));
    }
  }
  return 0;
}

int qshost_last_profile(qshost_db_t db, char *buf, uint64_t buf_bytes) {
  if (!buf || buf_bytes == 0) return QSGPU_ERR_INVALID;
  std::snprintf(buf, buf_bytes, "%s", db->last_profile.c_str());
  return 0;
}

int qshost_db_evict(qshost_db_t db, int which) {
  if (which < 0 || which > 2 || !db->rel[which]) return QSGPU_ERR_INVALID;
  db->sm->evict(*db->rel[which]);
  return 0;
}

int qshost_db_set_code_resident(qshost_db_t db, int on) {
  db->sm->setCodeResident(on != 0);
  for (int which = 0; which < 3; ++which) if (db->rel[which]) db->sm->evict(*db->rel[which]);
  return 0;
}

int qshost_db_resident_coding(qshost_db_t db, int which, uint32_t attr, uint32_t *code_width, uint32_t *n_entries) {
  if (which < 0 || which > 2 || !db->rel[which] || attr >= db->rel[which]->schema().size()) return QSGPU_ERR_INVALID;
  const auto c = db->sm->residentCoding(*db->rel[which], attr);
  if (code_width) *code_width = c.first;
  if (n_entries) *n_entries = c.second;
  return 0;
}

int qshost_db_stats(qshost_db_t db, int which, uint64_t *host_bytes, uint64_t *n_blocks, uint64_t *n_rows) {
  if (which < 0 || which > 2 || !db->rel[which]) return QSGPU_ERR_INVALID;
  if (host_bytes) *host_bytes = db->sm->hostBytes(*db->rel[which]);
  if (n_blocks) *n_blocks = db->rel[which]->getBlocksSnapshot().size();
  if (n_rows) *n_rows = db->rows[which];
  return 0;
}

// ------------------------------------------------------------------------ Q6
// lineitem -> Aggregation(single state) -> FinalizeAggregation -> DestroyAggregationState
int qshost_q6(qshost_db_t db, double *revenue, int *is_null, uint64_t *work_orders) {
  if (!db->rel[QSHOST_LINEITEM]) return QSGPU_ERR_INVALID;
  const CatalogRelation &lineitem = *db->rel[QSHOST_LINEITEM];
  const std::size_t query_id = db->next_query_id++;
  QueryContext ctx(db->sm.get(), db->dev);
  QueryContext::AggregationSpec spec;
  {
    ExprSet &e = spec.exprs;
    auto a = [&](int id) { return e.attr(id, lineitem.getAttributeById(id).type); };
    // l_shipdate >= 1994-01-01 AND l_shipdate < 1995-01-01 (interval arithmetic constant-folded)
    // AND l_discount BETWEEN 0.05 AND 0.07 AND l_quantity < 24
    spec.predicate_root = e.conj({e.cmp(QS_GE, a(L_SHIPDATE), e.lit_date(1994, 1, 1)),
                                  e.cmp(QS_LT, a(L_SHIPDATE), e.lit_date(1995, 1, 1)),
                                  e.cmp(QS_GE, a(L_DISCOUNT), e.lit_double(0.05)),
                                  e.cmp(QS_LE, a(L_DISCOUNT), e.lit_double(0.07)),
                                  e.cmp(QS_LT, a(L_QUANTITY), e.lit_int(24))});
    spec.aggregates.push_back({QS_AGG_SUM, e.binary(QS_MUL, a(L_EXTENDEDPRICE), a(L_DISCOUNT))});
    spec.strategy = QS_AGG_SINGLE_STATE;
  }
  const auto state = ctx.addAggregationState(std::move(spec));
  CatalogRelation *result = db->temp(anon({kDouble}));
  const auto dest = ctx.addInsertDestination(result, 1);

  QueryPlan plan;
  const auto agg = plan.addRelationalOperator(new AggregationOperator(query_id, lineitem, true, state, 1));
  const auto fin = plan.addRelationalOperator(new FinalizeAggregationOperator(query_id, state, 1, false, 1, *result, dest));
  const auto destroy = plan.addRelationalOperator(new DestroyAggregationStateOperator(query_id, state));
  plan.addDirectDependency(fin, agg, true);
  plan.addDirectDependency(destroy, fin, true);
  QueryManager qm(&plan, &ctx, db->sm.get(), db->workers.get());
  qm.run();

  // the query's single host wait: value, row count and NULL mask in one transfer
  qsgpu_relation_t out = db->sm->temporary(*result);
  double v = 0.0;
  void *cols[1] = {&v};
  std::uint64_t n = 0, nulls = 0;
  QS_CHECK_GPU(qsgpu_relation_read_rows(out, 1, cols, &n, &nulls));
  // the result relation's rows leave the device through its InsertDestination: host blocks in the reference's layout
  const std::vector<const void *> ccols = {&v};
  const std::vector<block_id> blocks = ctx.getInsertDestination(dest)->bulkInsertTuples(ccols, n);
  const SplitRowStoreReader rd(db->sm->getBlock(blocks[0]).memory, result->schema());
  *revenue = 0.0;
  if (rd.numTuples()) std::memcpy(revenue, rd.value(0, 0), 8);
  if (is_null) *is_null = (n == 0 || (nulls & 1)) ? 1 : 0;
  if (work_orders) *work_orders = qm.totalWorkOrdersExecuted();
  db->last_profile = qm.profile();
  db->dropTemps(result);
  return 0;
}

// ------------------------------------------------------------------------ Q1
// lineitem -> Aggregation(compact key) -> Finalize -> Selection(AVG = SUM / COUNT) [-> Sort]
int qshost_q1(qshost_db_t db, qshost_q1_row *rows, uint32_t *n_rows, uint64_t *work_orders) {
  if (!db->rel[QSHOST_LINEITEM]) return QSGPU_ERR_INVALID;
  const CatalogRelation &lineitem = *db->rel[QSHOST_LINEITEM];
  const std::size_t query_id = db->next_query_id++;
  QueryContext ctx(db->sm.get(), db->dev);
  QueryContext::AggregationSpec spec;
  {
    // After ReuseAggregateExpressions: SUM(qty), SUM(price), SUM(disc_price), SUM(charge), SUM(discount),
    // COUNT(*); disc_price is a shared subexpression (rules/ReuseAggregateExpressions.cpp:59-67,229-246).
    ExprSet &e = spec.exprs;
    auto a = [&](int id) { return e.attr(id, lineitem.getAttributeById(id).type); };
    spec.predicate_root = e.cmp(QS_LE, a(L_SHIPDATE), e.lit_date(1998, 9, 1));
    const int disc_price = e.shared(e.binary(QS_MUL, a(L_EXTENDEDPRICE), e.binary(QS_SUB, e.lit_int(1), a(L_DISCOUNT))), 0);
    const int charge = e.binary(QS_MUL, disc_price, e.binary(QS_ADD, e.lit_int(1), a(L_TAX)));
    spec.aggregates = {{QS_AGG_SUM, a(L_QUANTITY)}, {QS_AGG_SUM, a(L_EXTENDEDPRICE)}, {QS_AGG_SUM, disc_price},
                       {QS_AGG_SUM, charge}, {QS_AGG_SUM, a(L_DISCOUNT)}, {QS_AGG_COUNT, -1}};
    spec.group_by_roots = {a(L_RETURNFLAG), a(L_LINESTATUS)};
    spec.strategy = QS_AGG_COMPACT_KEY;
    spec.estimated_num_entries = 8;
  }
  const auto state = ctx.addAggregationState(std::move(spec));
  // finalize output: flag, status, 5 sums, count
  CatalogRelation *t_fin = db->temp(anon({kChar1, kChar1, kDouble, kDouble, kDouble, kDouble, kDouble, kLong}));
  const auto d_fin = ctx.addInsertDestination(t_fin, 256);
  // wrapping Selection: flag, status, sum_qty, sum_base_price, sum_disc_price, sum_charge, avg_qty, avg_price, avg_disc,
  // count (+ the raw SUM(l_discount), see qshost_q1_row)
  CatalogRelation *t_out = db->temp(anon({kChar1, kChar1, kDouble, kDouble, kDouble, kDouble, kDouble, kDouble, kDouble, kLong, kDouble}));
  const auto d_out = ctx.addInsertDestination(t_out, 256);
  QueryContext::ScalarGroup sel;
  {
    ExprSet &e = sel.exprs;
    auto a = [&](int id) { return e.attr(id, t_fin->getAttributeById(id).type); };
    sel.roots = {a(0), a(1), a(2), a(3), a(4), a(5), e.binary(QS_DIV, a(2), a(7)), e.binary(QS_DIV, a(3), a(7)),
                 e.binary(QS_DIV, a(6), a(7)), a(7), a(6)};
  }
  const auto sel_id = ctx.addScalarGroup(std::move(sel));
  // ORDER BY l_returnflag, l_linestatus: the sort operators collapse into one top-k work order (<= 6 groups)
  CatalogRelation *t_sorted = db->temp(anon({kChar1, kChar1, kDouble, kDouble, kDouble, kDouble, kDouble, kDouble, kDouble, kLong, kDouble}));
  const auto d_sorted = ctx.addInsertDestination(t_sorted, 64);
  QueryContext::SortConfig sc; sc.keys = {{0, 0}, {1, 0}};
  const auto sort_id = ctx.addSortConfig(sc);

  QueryPlan plan;
  const auto agg = plan.addRelationalOperator(new AggregationOperator(query_id, lineitem, true, state, 1));
  const auto fin = plan.addRelationalOperator(new FinalizeAggregationOperator(query_id, state, 1, false, 1, *t_fin, d_fin));
  const auto select = plan.addRelationalOperator(new SelectOperator(query_id, *t_fin, false, *t_out, d_out, QueryContext::kInvalidPredicateId, sel_id, false));
  const auto destroy = plan.addRelationalOperator(new DestroyAggregationStateOperator(query_id, state));
  const auto sort = plan.addRelationalOperator(new SortMergeRunOperator(query_id, *t_out, *t_sorted, d_sorted, sort_id, 64, false));
  plan.addDirectDependency(fin, agg, true);
  plan.addDirectDependency(select, fin, false);
  plan.addDirectDependency(destroy, fin, true);
  plan.addDirectDependency(sort, select, true);
  QueryManager qm(&plan, &ctx, db->sm.get(), db->workers.get());
  qm.run();

  qsgpu_relation_t out = db->sm->temporary(*t_sorted);
  constexpr std::uint64_t kMaxRows = 64;         // the sort's LIMIT
  std::vector<char> flag(kMaxRows), status(kMaxRows);
  std::array<std::vector<double>, 7> d;
  for (auto &v : d) v.resize(kMaxRows);
  std::vector<std::int64_t> count(kMaxRows);
  std::vector<double> sum_disc(kMaxRows);
  std::uint64_t n = 0;
  {
    // the query's single host wait: rows and row count in one transfer
    void *cols[11] = {flag.data(), status.data(), d[0].data(), d[1].data(), d[2].data(), d[3].data(), d[4].data(),
                      d[5].data(), d[6].data(), count.data(), sum_disc.data()};
    QS_CHECK_GPU(qsgpu_relation_read_rows(out, kMaxRows, cols, &n, nullptr));
    n = std::min(n, kMaxRows);
  }
  // InsertDestination::bulkInsertTuples: the result rows become SplitRowStore blocks of the result relation on the host
  // (what the reference's PrintToScreen reads); the rows handed back are read from those blocks
  const std::vector<const void *> ccols = {flag.data(), status.data(), d[0].data(), d[1].data(), d[2].data(), d[3].data(), d[4].data(),
                                           d[5].data(), d[6].data(), count.data(), sum_disc.data()};
  const std::vector<block_id> blocks = ctx.getInsertDestination(d_sorted)->bulkInsertTuples(ccols, n);
  const SplitRowStoreReader rd(db->sm->getBlock(blocks[0]).memory, t_sorted->schema());
  std::vector<qshost_q1_row> res(n);
  for (std::uint64_t i = 0; i < n; ++i) {
    qshost_q1_row &r = res[i];
    std::memset(&r, 0, sizeof(r));
    r.l_returnflag = *rd.value(i, 0); r.l_linestatus = *rd.value(i, 1);
    double *dst[7] = {&r.sum_qty, &r.sum_base_price, &r.sum_disc_price, &r.sum_charge, &r.avg_qty, &r.avg_price, &r.avg_disc};
    for (int k = 0; k < 7; ++k) std::memcpy(dst[k], rd.value(i, 2 + k), 8);
    std::memcpy(&r.count_order, rd.value(i, 9), 8);
    std::memcpy(&r.sum_disc, rd.value(i, 10), 8);
  }
  const std::uint32_t cap = *n_rows;
  *n_rows = static_cast<std::uint32_t>(n);
  for (std::uint32_t i = 0; i < std::min<std::uint64_t>(cap, n); ++i) rows[i] = res[i];
  if (work_orders) *work_orders = qm.totalWorkOrdersExecuted();
  db->last_profile = qm.profile();
  db->dropTemps(t_sorted);
  return n > cap ? QSGPU_ERR_CAPACITY : 0;
}

// ------------------------------------------------------------------------ Q3
// [1] BuildLIPFilter(customer: c_mktsegment = 'BUILDING' -> exact filter on c_custkey)   (InjectJoinFilters)
// [2] Select(orders: o_orderdate < 1995-03-15, probe c_custkey filter with o_custkey) -> T2
// [3] BuildHash(T2 on o_orderkey) + build exact filter on o_orderkey                       (AttachLIPFilters)
// [0] Select(lineitem: l_shipdate > 1995-03-15, probe o_orderkey filter with l_orderkey) -> T0
// [4] HashJoin(probe T0, build T2) -> T4   [5] DestroyHash
// [6] Aggregation(T4 GROUP BY l_orderkey, o_orderdate, o_shippriority; SUM(price * (1 - discount)))
// [7] FinalizeAggregation -> T7   [8] DestroyAggregationState
// [9] SortMergeRun(revenue DESC, o_orderdate; top 10) -> T9
int qshost_q3(qshost_db_t db, qshost_q3_row *rows, uint32_t *n_rows, uint64_t *work_orders) {
  if (!db->rel[0] || !db->rel[1] || !db->rel[2]) return QSGPU_ERR_INVALID;
  const CatalogRelation &customer = *db->rel[QSHOST_CUSTOMER], &orders = *db->rel[QSHOST_ORDERS], &lineitem = *db->rel[QSHOST_LINEITEM];
  const std::size_t query_id = db->next_query_id++;
  const std::uint64_t n_orders = db->rows[QSHOST_ORDERS], n_lineitem = db->rows[QSHOST_LINEITEM];
  QueryContext ctx(db->sm.get(), db->dev);

  const auto f_cust = ctx.addLIPFilter(QS_LIP_BITVECTOR_EXACT, QS_INT, db->c_custkey_min, db->c_custkey_max, 0, false);
  const auto f_ord = ctx.addLIPFilter(QS_LIP_BITVECTOR_EXACT, QS_INT, db->o_orderkey_min, db->o_orderkey_max, 0, false);
  QueryContext::LIPDeployment dep1; dep1.build_entries = {{f_cust, C_CUSTKEY}};
  QueryContext::LIPDeployment dep2; dep2.probe_entries = {{f_cust, O_CUSTKEY}};
  QueryContext::LIPDeployment dep3; dep3.build_entries = {{f_ord, 0}};
  QueryContext::LIPDeployment dep0; dep0.probe_entries = {{f_ord, L_ORDERKEY}};
  const auto d1 = ctx.addLIPDeployment(dep1), d2 = ctx.addLIPDeployment(dep2), d3 = ctx.addLIPDeployment(dep3), d0 = ctx.addLIPDeployment(dep0);

  QueryContext::Predicate p1, p2, p0;
  p1.root = p1.exprs.cmp(QS_EQ, p1.exprs.attr(C_MKTSEGMENT, kChar10), p1.exprs.lit_char("BUILDING"));
  p2.root = p2.exprs.cmp(QS_LT, p2.exprs.attr(O_ORDERDATE, kDate), p2.exprs.lit_date(1995, 3, 15));
  p0.root = p0.exprs.cmp(QS_GT, p0.exprs.attr(L_SHIPDATE, kDate), p0.exprs.lit_date(1995, 3, 15));
  const auto pid1 = ctx.addPredicate(std::move(p1)), pid2 = ctx.addPredicate(std::move(p2)), pid0 = ctx.addPredicate(std::move(p0));

  CatalogRelation *t2 = db->temp(anon({kInt, kDate, kInt}));
  CatalogRelation *t0 = db->temp(anon({kInt, kDouble, kDouble}));
  CatalogRelation *t4 = db->temp(anon({kInt, kDate, kInt, kDouble, kDouble}));
  CatalogRelation *t7 = db->temp(anon({kInt, kDate, kInt, kDouble}));
  CatalogRelation *t9 = db->temp(anon({kInt, kDate, kInt, kDouble}));
  const auto dst2 = ctx.addInsertDestination(t2, n_orders), dst0 = ctx.addInsertDestination(t0, n_lineitem),
             dst4 = ctx.addInsertDestination(t4, n_lineitem), dst7 = ctx.addInsertDestination(t7, 1),
             dst9 = ctx.addInsertDestination(t9, 10);
  // Several devices: partition-wise join (one hash table per partition = per device, the reference's num_partitions)
  // when orders and lineitem are co-partitioned on the order key, else ONE logical table whose build side is
  // all-gathered to every device (broadcast join) and therefore sized for the whole relation.
  int n_ranks = 1;
  QS_CHECK_GPU(qsgpu_comm_rank(db->comm, nullptr, &n_ranks));
  const std::size_t join_partitions = (n_ranks > 1 && db->join_mode == 0) ? static_cast<std::size_t>(n_ranks) : 1u;
  const std::uint64_t build_rows = join_partitions > 1 ? n_orders : db->global_rows[QSHOST_ORDERS];
  const auto ht = ctx.addJoinHashTable(QS_INT, std::max<std::uint64_t>(1024, build_rows / 4));

  QueryContext::ScalarGroup s4;     // l_orderkey, o_orderdate, o_shippriority, l_extendedprice, l_discount
  s4.roots = {s4.exprs.attr(0, kInt), s4.exprs.attr(1, kDate, 2), s4.exprs.attr(2, kInt, 2), s4.exprs.attr(1, kDouble),
              s4.exprs.attr(2, kDouble)};
  const auto sel4 = ctx.addScalarGroup(std::move(s4));

  QueryContext::AggregationSpec spec;
  {
    ExprSet &e = spec.exprs;
    spec.aggregates.push_back({QS_AGG_SUM, e.binary(QS_MUL, e.attr(3, kDouble), e.binary(QS_SUB, e.lit_int(1), e.attr(4, kDouble)))});
    spec.group_by_roots = {e.attr(0, kInt), e.attr(1, kDate), e.attr(2, kInt)};
    spec.strategy = QS_AGG_SEPARATE_CHAINING;
    // lineitem is partitioned on l_orderkey boundaries (it is sorted on it, benchmarks/tpch/create.sql:112):
    // no group spans two devices
    spec.partitioned_on_group_by = true;
    // the optimizer's group estimate (StarSchemaSimpleCostModel::estimateNumGroupsForAggregate): ~ #orders / 8
    spec.estimated_num_entries = std::max<std::uint64_t>(1u << 16, n_orders / 8);
  }
  const auto state = ctx.addAggregationState(std::move(spec));
  QueryContext::SortConfig sc; sc.keys = {{3, 1}, {1, 0}};
  const auto sort_id = ctx.addSortConfig(sc);

  QueryPlan plan;
  auto *op1 = new BuildLIPFilterOperator(query_id, customer, pid1, true);
  op1->deployLIPFilters(d1, {f_cust});
  auto *op2 = new SelectOperator(query_id, orders, false, *t2, dst2, pid2, std::vector<attribute_id>{O_ORDERKEY, O_ORDERDATE, O_SHIPPRIORITY}, true);
  op2->deployLIPFilters(d2, {f_cust});
  auto *op3 = new BuildHashOperator(query_id, *t2, false, {0}, false, join_partitions, ht);
  op3->deployLIPFilters(d3, {f_ord});
  auto *op0 = new SelectOperator(query_id, lineitem, false, *t0, dst0, pid0, std::vector<attribute_id>{L_ORDERKEY, L_EXTENDEDPRICE, L_DISCOUNT}, true);
  op0->deployLIPFilters(d0, {f_ord});
  auto *op4 = new HashJoinOperator(query_id, *t2, *t0, false, {0}, false, join_partitions, false, *t4, dst4, ht, QueryContext::kInvalidPredicateId, sel4);
  const auto i1 = plan.addRelationalOperator(op1), i2 = plan.addRelationalOperator(op2), i3 = plan.addRelationalOperator(op3),
             i0 = plan.addRelationalOperator(op0), i4 = plan.addRelationalOperator(op4);
  const auto i5 = plan.addRelationalOperator(new DestroyHashOperator(query_id, 1, ht));
  const auto i6 = plan.addRelationalOperator(new AggregationOperator(query_id, *t4, false, state, 1));
  const auto i7 = plan.addRelationalOperator(new FinalizeAggregationOperator(query_id, state, 1, false, 1, *t7, dst7));
  const auto i8 = plan.addRelationalOperator(new DestroyAggregationStateOperator(query_id, state));
  const auto i9 = plan.addRelationalOperator(new SortMergeRunOperator(query_id, *t7, *t9, dst9, sort_id, 10, false));
  plan.addDirectDependency(i2, i1, true);     // LIP filter must be complete before it is probed
  plan.addDirectDependency(i3, i2, false);    // pipelined
  plan.addDirectDependency(i0, i3, true);     // LIP filter on o_orderkey
  plan.addDirectDependency(i4, i3, true);     // hash table complete before probing
  plan.addDirectDependency(i4, i0, false);
  plan.addDirectDependency(i5, i4, true);
  plan.addDirectDependency(i6, i4, false);
  plan.addDirectDependency(i7, i6, true);
  plan.addDirectDependency(i8, i7, true);
  plan.addDirectDependency(i9, i7, true);
  QueryManager qm(&plan, &ctx, db->sm.get(), db->workers.get());
  qm.run();

  qsgpu_relation_t out = db->sm->temporary(*t9);
  constexpr std::uint64_t kMaxRows = 10;         // LIMIT 10
  std::vector<std::int32_t> ok(kMaxRows), sp(kMaxRows);
  std::vector<std::uint64_t> od(kMaxRows);
  std::vector<double> rev(kMaxRows);
  std::uint64_t n = 0;
  {
    void *cols[4] = {ok.data(), od.data(), sp.data(), rev.data()};
    QS_CHECK_GPU(qsgpu_relation_read_rows(out, kMaxRows, cols, &n, nullptr));
    n = std::min(n, kMaxRows);
  }
  const std::vector<const void *> ccols = {ok.data(), od.data(), sp.data(), rev.data()};
  const std::vector<block_id> blocks = ctx.getInsertDestination(dst9)->bulkInsertTuples(ccols, n);
  const SplitRowStoreReader rd(db->sm->getBlock(blocks[0]).memory, t9->schema());
  const std::uint32_t cap = *n_rows;
  *n_rows = static_cast<std::uint32_t>(n);
  for (std::uint32_t i = 0; i < std::min<std::uint64_t>(cap, n); ++i) {
    qshost_q3_row &r = rows[i];
    std::memset(&r, 0, sizeof(r));
    std::uint64_t date = 0;
    std::memcpy(&r.l_orderkey, rd.value(i, 0), 4);
    std::memcpy(&date, rd.value(i, 1), 8);
    std::memcpy(&r.o_shippriority, rd.value(i, 2), 4);
    std::memcpy(&r.revenue, rd.value(i, 3), 8);
    r.year = static_cast<std::int32_t>(date & 0xffffffffu);
    r.month = static_cast<std::uint8_t>((date >> 32) & 0xff);
    r.day = static_cast<std::uint8_t>((date >> 40) & 0xff);
  }
  if (work_orders) *work_orders = qm.totalWorkOrdersExecuted();
  db->last_profile = qm.profile();
  db->dropTemps(t9);
  return n > cap ? QSGPU_ERR_CAPACITY : 0;
}

}  // extern "C"
