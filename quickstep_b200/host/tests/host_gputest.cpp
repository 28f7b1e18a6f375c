// GPU tests of the operator classes that the TPC-H plans do not reach: HashJoinOperator as a LEFT OUTER join on a
// composite (two INT) key, and BuildAggregationExistenceMapOperator feeding a collision-free aggregation.  Plans
// are built the way ExecutionGenerator builds them and run by the QueryManager with 4 workers over relations
// stored as several blocks; expected results are computed right here with plain loops.  Needs a B200.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <tuple>
#include <vector>

#include "Operators.hpp"
#include "QueryManager.hpp"

using namespace quickstep;

static int g_failed = 0;
#define EXPECT(cond)                                                                  \
  do {                                                                                \
    if (!(cond)) { std::printf("  FAILED %s (%s:%d)\n", #cond, __FILE__, __LINE__); ++g_failed; } \
  } while (0)

static const qs_attr kInt{QS_INT, 4}, kLong{QS_LONG, 8}, kDouble{QS_DOUBLE, 8};

template <class T>
static std::vector<T> readColumn(qsgpu_relation_t rel, std::uint32_t attr, std::uint64_t n) {
  std::vector<T> v(std::max<std::uint64_t>(n, 1));
  if (n) QS_CHECK_GPU(qsgpu_relation_read(rel, attr, 0, n, v.data()));
  v.resize(n);
  return v;
}

static std::uint64_t g_rng = 0x9e3779b97f4a7c15ull;
static std::uint64_t rnd() { g_rng ^= g_rng << 13; g_rng ^= g_rng >> 7; g_rng ^= g_rng << 17; return g_rng; }

static void testOuterJoinCompositeKey(StorageManager *sm, WorkerPool *pool) {
  const std::uint64_t nb = 1500, np = 6000;
  std::vector<std::int32_t> b0(nb), b1(nb), p0(np), p1(np);
  std::vector<std::int64_t> bp(nb);
  std::vector<double> pv(np);
  for (std::uint64_t i = 0; i < nb; ++i) { b0[i] = static_cast<std::int32_t>(rnd() % 30) - 10; b1[i] = static_cast<std::int32_t>(rnd() % 30); bp[i] = static_cast<std::int64_t>(i) * 3 + 1; }
  for (std::uint64_t i = 0; i < np; ++i) { p0[i] = static_cast<std::int32_t>(rnd() % 36) - 12; p1[i] = static_cast<std::int32_t>(rnd() % 36); pv[i] = static_cast<double>(i) * 0.5; }
  CatalogRelation build(1, "b", {{"k0", kInt}, {"k1", kInt}, {"p", kLong}});
  CatalogRelation probe(2, "a", {{"k0", kInt}, {"k1", kInt}, {"v", kDouble}});
  CatalogRelation out(3, "t", {{"k0", kInt}, {"k1", kInt}, {"p", kLong}, {"v", kDouble}}, true);
  sm->loadRelation(&build, {b0.data(), b1.data(), bp.data()}, nb, 400, TupleStoreLayout::kCompressedColumnStore);
  sm->loadRelation(&probe, {p0.data(), p1.data(), pv.data()}, np, 1000, TupleStoreLayout::kSplitRowStore);

  QueryContext ctx(sm, sm->device());
  const auto ht = ctx.addJoinHashTable(QS_LONG, nb);
  QueryContext::ScalarGroup sel;
  sel.roots = {sel.exprs.attr(0, kInt), sel.exprs.attr(1, kInt), sel.exprs.attr(2, kLong, 2), sel.exprs.attr(2, kDouble)};
  const auto sel_id = ctx.addScalarGroup(std::move(sel));
  const auto dst = ctx.addInsertDestination(&out, 200000);
  QueryPlan plan;
  const std::vector<bool> on_build = {false, false, true, false};
  const auto i_build = plan.addRelationalOperator(new BuildHashOperator(1, build, true, {0, 1}, false, 1, ht));
  const auto i_join = plan.addRelationalOperator(new HashJoinOperator(1, build, probe, true, {0, 1}, false, 1, false, out, dst, ht,
                                                                      QueryContext::kInvalidPredicateId, sel_id, &on_build,
                                                                      JoinType::kLeftOuterJoin));
  const auto i_destroy = plan.addRelationalOperator(new DestroyHashOperator(1, 1, ht));
  plan.addDirectDependency(i_join, i_build, true);
  plan.addDirectDependency(i_destroy, i_join, true);
  FLAGS_gpu_rows_per_workorder = 2000;            // several probe work orders appending to one output
  QueryManager qm(&plan, &ctx, sm, pool);
  qm.run();
  FLAGS_gpu_rows_per_workorder = 0;
  EXPECT(qm.numWorkOrdersExecuted(i_join) == 3);

  qsgpu_relation_t rel = sm->temporary(out);
  std::uint64_t n = 0;
  QS_CHECK_GPU(qsgpu_relation_num_rows(rel, &n));
  const auto k0 = readColumn<std::int32_t>(rel, 0, n), k1 = readColumn<std::int32_t>(rel, 1, n);
  const auto p = readColumn<std::int64_t>(rel, 2, n);
  const auto v = readColumn<double>(rel, 3, n);
  std::vector<std::uint64_t> nulls(std::max<std::uint64_t>(n, 1));
  QS_CHECK_GPU(qsgpu_relation_read_nulls(rel, 0, n, nulls.data()));
  typedef std::tuple<std::int32_t, std::int32_t, std::int64_t, double, int> Row;
  std::vector<Row> got, want;
  for (std::uint64_t i = 0; i < n; ++i) got.emplace_back(k0[i], k1[i], p[i], v[i], static_cast<int>(nulls[i]));
  std::multimap<std::pair<std::int32_t, std::int32_t>, std::int64_t> table;
  for (std::uint64_t i = 0; i < nb; ++i) table.insert({{b0[i], b1[i]}, bp[i]});
  std::uint64_t unmatched = 0;
  for (std::uint64_t i = 0; i < np; ++i) {
    auto range = table.equal_range({p0[i], p1[i]});
    if (range.first == range.second) { want.emplace_back(p0[i], p1[i], 0, pv[i], 0b0100); ++unmatched; }
    for (auto it = range.first; it != range.second; ++it) want.emplace_back(p0[i], p1[i], it->second, pv[i], 0);
  }
  std::sort(got.begin(), got.end());
  std::sort(want.begin(), want.end());
  EXPECT(unmatched > 100 && got.size() == want.size());
  EXPECT(got == want);
  sm->dropTemporary(out);
  std::printf("outer_join_composite_key %s (%zu rows, %llu without a match)\n", got == want ? "ok" : "MISMATCH", got.size(),
              static_cast<unsigned long long>(unmatched));
}

static void testExistenceMapAggregation(StorageManager *sm, WorkerPool *pool) {
  const std::uint64_t nl = 1000, nr = 30000;
  std::vector<std::int32_t> lk(nl), rk(nr);
  std::vector<std::int64_t> rv(nr);
  for (std::uint64_t i = 0; i < nl; ++i) lk[i] = static_cast<std::int32_t>(i * 3);
  for (std::uint64_t i = 0; i < nr; ++i) { rk[i] = lk[(rnd() % (nl / 2)) * 2]; rv[i] = static_cast<std::int64_t>(rnd() % 100) + 1; }
  CatalogRelation left(11, "l", {{"k", kInt}});
  CatalogRelation right(12, "r", {{"k", kInt}, {"v", kLong}});
  CatalogRelation out(13, "o", {{"k", kInt}, {"c", kLong}, {"s", kLong}}, true);
  sm->loadRelation(&left, {lk.data()}, nl, 300, TupleStoreLayout::kBasicColumnStore);
  sm->loadRelation(&right, {rk.data(), rv.data()}, nr, 7000, TupleStoreLayout::kCompressedColumnStore);
  QueryContext ctx(sm, sm->device());
  QueryContext::AggregationSpec spec;
  spec.aggregates = {{QS_AGG_COUNT, -1}, {QS_AGG_SUM, spec.exprs.attr(1, kLong)}};
  spec.group_by_roots = {spec.exprs.attr(0, kInt)};
  spec.strategy = QS_AGG_COLLISION_FREE;
  spec.collision_free_max_key = 2999;
  const auto state = ctx.addAggregationState(std::move(spec));
  const auto dst = ctx.addInsertDestination(&out, 1);
  QueryPlan plan;
  const auto i_init = plan.addRelationalOperator(new InitializeAggregationOperator(1, state));
  const auto i_exist = plan.addRelationalOperator(new BuildAggregationExistenceMapOperator(1, left, 0, true, state, 1));
  const auto i_agg = plan.addRelationalOperator(new AggregationOperator(1, right, true, state, 1));
  const auto i_fin = plan.addRelationalOperator(new FinalizeAggregationOperator(1, state, 1, false, 1, out, dst));
  const auto i_destroy = plan.addRelationalOperator(new DestroyAggregationStateOperator(1, state));
  plan.addDirectDependency(i_exist, i_init, true);       // ExecutionGenerator.cpp:2173-2180
  plan.addDirectDependency(i_agg, i_exist, true);
  plan.addDirectDependency(i_fin, i_agg, true);
  plan.addDirectDependency(i_destroy, i_fin, true);
  QueryManager qm(&plan, &ctx, sm, pool);
  qm.run();
  qsgpu_relation_t rel = sm->temporary(out);
  std::uint64_t n = 0;
  QS_CHECK_GPU(qsgpu_relation_num_rows(rel, &n));
  const auto k = readColumn<std::int32_t>(rel, 0, n);
  const auto c = readColumn<std::int64_t>(rel, 1, n), s = readColumn<std::int64_t>(rel, 2, n);
  std::map<std::int32_t, std::pair<std::int64_t, std::int64_t>> want, got;
  for (std::uint64_t i = 0; i < nl; ++i) want[lk[i]] = {0, 0};
  for (std::uint64_t i = 0; i < nr; ++i) { want[rk[i]].first += 1; want[rk[i]].second += rv[i]; }
  for (std::uint64_t i = 0; i < n; ++i) got[k[i]] = {c[i], s[i]};
  EXPECT(n == nl && got == want);
  sm->dropTemporary(out);
  std::printf("existence_map_aggregation %s (%llu groups)\n", got == want ? "ok" : "MISMATCH", static_cast<unsigned long long>(n));
}

// SelectOperator writing through a PartitionAwareInsertDestination (hash partition scheme on the key, 4 and 3
// partitions), an AggregationOperator consuming the partitions: one work order per partition, every row in the
// partition HashPartitionSchemeHeader::getPartitionId assigns it to, and the same aggregate as without repartitioning.
static void testPartitionAwareInsertDestination(StorageManager *sm, WorkerPool *pool, std::size_t num_partitions) {
  const std::uint64_t n = 50000;
  std::vector<std::int32_t> k(n);
  std::vector<std::int64_t> v(n);
  for (std::uint64_t i = 0; i < n; ++i) { k[i] = static_cast<std::int32_t>(rnd() % 5000) - 100; v[i] = static_cast<std::int64_t>(rnd() % 1000); }
  const relation_id base_id = 20 + 10 * static_cast<relation_id>(num_partitions);    // every call loads its own relation
  CatalogRelation in(base_id + 1, "in", {{"k", kInt}, {"v", kLong}});
  CatalogRelation mid(base_id + 2, "mid", {{"k", kInt}, {"v", kLong}}, true);
  CatalogRelation out(base_id + 3, "out", {{"c", kLong}, {"s", kLong}}, true);
  sm->loadRelation(&in, {k.data(), v.data()}, n, 9000, TupleStoreLayout::kCompressedColumnStore);
  QueryContext ctx(sm, sm->device());
  QueryContext::Predicate pred;
  pred.root = pred.exprs.cmp(QS_LT, pred.exprs.attr(1, kLong), pred.exprs.lit_int(900));
  const auto pid = ctx.addPredicate(std::move(pred));
  HashPartitionSchemeHeader header;
  header.num_partitions = num_partitions;
  header.partition_attribute = 0;
  const auto d_mid = ctx.addPartitionAwareInsertDestination(header, &mid, n);
  QueryContext::AggregationSpec spec;
  spec.aggregates = {{QS_AGG_COUNT, -1}, {QS_AGG_SUM, spec.exprs.attr(1, kLong)}};
  spec.strategy = QS_AGG_SINGLE_STATE;
  const auto state = ctx.addAggregationState(std::move(spec));
  const auto d_out = ctx.addInsertDestination(&out, 1);
  QueryPlan plan;
  const auto i_sel = plan.addRelationalOperator(new SelectOperator(1, in, /*has_repartition=*/true, mid, d_mid, pid, std::vector<attribute_id>{0, 1}, true));
  const auto i_agg = plan.addRelationalOperator(new AggregationOperator(1, mid, false, state, num_partitions));
  const auto i_fin = plan.addRelationalOperator(new FinalizeAggregationOperator(1, state, 1, false, 1, out, d_out));
  plan.addDirectDependency(i_agg, i_sel, true);          // the repartitioned relation must be complete
  plan.addDirectDependency(i_fin, i_agg, true);
  QueryManager qm(&plan, &ctx, sm, pool);
  qm.run();
  EXPECT(qm.numWorkOrdersExecuted(i_agg) == num_partitions);
  // every row sits in the partition the reference's scheme assigns it to
  std::uint64_t rows_seen = 0;
  bool placed = true;
  for (const std::pair<block_id, partition_id> &b : ctx.getInsertDestination(d_mid)->getTouchedBlocks()) {
    const DeviceExtent e = sm->blockExtent(b.first);
    std::vector<std::int32_t> keys(std::max<std::uint64_t>(e.row_end - e.row_begin, 1));
    if (e.row_end > e.row_begin) QS_CHECK_GPU(qsgpu_relation_read(e.relation, 0, e.row_begin, e.row_end - e.row_begin, keys.data()));
    for (std::uint64_t i = 0; i < e.row_end - e.row_begin; ++i) {
      const std::uint64_t h = static_cast<std::uint32_t>(keys[i]);        // identity hash of the INT's bit pattern
      const std::uint64_t p = (num_partitions & (num_partitions - 1)) == 0 ? (h & (num_partitions - 1)) : (h % num_partitions);
      placed = placed && p == b.second;
    }
    rows_seen += e.row_end - e.row_begin;
  }
  std::int64_t want_c = 0, want_s = 0;
  for (std::uint64_t i = 0; i < n; ++i) if (v[i] < 900) { ++want_c; want_s += v[i]; }
  qsgpu_relation_t rel = sm->temporary(out);
  const auto c = readColumn<std::int64_t>(rel, 0, 1), s2 = readColumn<std::int64_t>(rel, 1, 1);
  EXPECT(placed);
  EXPECT(rows_seen == static_cast<std::uint64_t>(want_c));
  EXPECT(c[0] == want_c && s2[0] == want_s);
  sm->dropTemporary(mid);
  sm->dropTemporary(out);
  std::printf("partition_aware_insert_destination(%zu) %s\n", num_partitions, (placed && c[0] == want_c && s2[0] == want_s) ? "ok" : "MISMATCH");
}

// A relation with NULL-able attributes in each of the three block layouts (per-tuple bitmap, per-column bitmap,
// dictionary null code), aggregated with GROUP BY through the operator DAG: NULL arguments are skipped
// (AggregationHandleSum.hpp:117-127), comparisons with a NULL are false, a group without a non-NULL argument has
// COUNT(x) = 0, SUM(x) = 0 and MIN(y) NULL -- as the reference engine prints (tests/golden/ref_null_results.json).
static void testNullableAggregation(StorageManager *sm, WorkerPool *pool, TupleStoreLayout layout, relation_id base_id) {
  const std::uint64_t n = 30000;
  std::vector<std::int32_t> g(n), y(n);
  std::vector<double> x(n);
  std::vector<std::uint8_t> xn(n), yn(n);
  for (std::uint64_t i = 0; i < n; ++i) {
    g[i] = static_cast<std::int32_t>(rnd() % 6);
    xn[i] = (rnd() % 10) < 3 || g[i] == 2;                 // group 2: x is always NULL
    yn[i] = (rnd() % 10) < 1 || g[i] == 2;
    x[i] = xn[i] ? 0.0 : static_cast<double>(rnd() % 100000) / 100.0;
    y[i] = yn[i] ? 0 : static_cast<std::int32_t>(rnd() % 2000) - 1000;
  }
  CatalogRelation in(base_id, "tn", {{"g", kInt}, {"x", kDouble, true}, {"y", kInt, true}});
  CatalogRelation out(base_id + 1, "tn_out", {{"g", kInt}, {"sx", kDouble}, {"cx", kLong}, {"my", kInt}, {"c", kLong}}, true);
  sm->loadRelation(&in, {g.data(), x.data(), y.data()}, n, 7000, layout, {nullptr, xn.data(), yn.data()});
  QueryContext ctx(sm, sm->device());
  QueryContext::AggregationSpec spec;
  spec.predicate_root = spec.exprs.disj({spec.exprs.cmp(QS_GT, spec.exprs.attr(2, kInt), spec.exprs.lit_int(-500)),
                                         spec.exprs.cmp(QS_EQ, spec.exprs.attr(0, kInt), spec.exprs.lit_int(2))});
  const int ax = spec.exprs.attr(1, kDouble), ay = spec.exprs.attr(2, kInt);
  spec.aggregates = {{QS_AGG_SUM, ax}, {QS_AGG_COUNT, ax}, {QS_AGG_MIN, ay}, {QS_AGG_COUNT, -1}};
  spec.nullable_arguments = 0b0111;
  spec.group_by_roots = {spec.exprs.attr(0, kInt)};
  spec.strategy = QS_AGG_COMPACT_KEY;
  spec.estimated_num_entries = 8;
  const auto state = ctx.addAggregationState(std::move(spec));
  const auto dst = ctx.addInsertDestination(&out, 16);
  QueryPlan plan;
  const auto i_agg = plan.addRelationalOperator(new AggregationOperator(1, in, true, state, 1));
  const auto i_fin = plan.addRelationalOperator(new FinalizeAggregationOperator(1, state, 1, false, 1, out, dst));
  plan.addDirectDependency(i_fin, i_agg, true);
  QueryManager qm(&plan, &ctx, sm, pool);
  qm.run();
  qsgpu_relation_t rel = sm->temporary(out);
  std::uint64_t rows = 0;
  QS_CHECK_GPU(qsgpu_relation_num_rows(rel, &rows));
  const auto k = readColumn<std::int32_t>(rel, 0, rows);
  const auto sx = readColumn<double>(rel, 1, rows);
  const auto cx = readColumn<std::int64_t>(rel, 2, rows), c = readColumn<std::int64_t>(rel, 4, rows);
  const auto my = readColumn<std::int32_t>(rel, 3, rows);
  std::vector<std::uint64_t> nulls(std::max<std::uint64_t>(rows, 1));
  QS_CHECK_GPU(qsgpu_relation_read_nulls(rel, 0, rows, nulls.data()));
  struct Want { double sx = 0; std::int64_t cx = 0, c = 0; std::int32_t my = 0; bool any_y = false; };
  std::map<std::int32_t, Want> want;
  for (std::uint64_t i = 0; i < n; ++i) {
    const bool pass = (!yn[i] && y[i] > -500) || g[i] == 2;          // a comparison with a NULL is false
    if (!pass) continue;
    Want &w = want[g[i]];
    ++w.c;
    if (!xn[i]) { w.sx += x[i]; ++w.cx; }
    if (!yn[i]) { w.my = w.any_y ? std::min(w.my, y[i]) : y[i]; w.any_y = true; }
  }
  bool ok = rows == want.size();
  for (std::uint64_t i = 0; ok && i < rows; ++i) {
    const Want &w = want[k[i]];
    const bool y_null = (nulls[i] >> 3) & 1, x_null = (nulls[i] >> 1) & 1;
    ok = ok && cx[i] == w.cx && c[i] == w.c && std::fabs(sx[i] - w.sx) <= 1e-9 * std::fabs(w.sx) && !x_null &&
         y_null == !w.any_y && (y_null || my[i] == w.my);
  }
  EXPECT(ok && want.count(2) && want[2].cx == 0 && !want[2].any_y);
  sm->dropTemporary(out);
  std::printf("nullable_aggregation(layout %d) %s (%llu groups)\n", static_cast<int>(layout), ok ? "ok" : "MISMATCH", static_cast<unsigned long long>(rows));
}

// Code-resident mode with FLOAT / DOUBLE values that have no place in a numerically ordered dictionary: -0.0 (equal to
// 0.0 but a different bit pattern) and NaN (equal to nothing).  The block builder still dictionary-compresses them
// (one code per bit pattern); the storage manager keeps such an attribute at native width, and what comes back from
// the device is bit-identical to what was loaded.
static void testSpecialFloatsStayNative(StorageManager *sm) {
  const std::uint64_t n = 20000;
  std::vector<double> v(n), plain(n);
  std::vector<std::int32_t> k(n);
  const double specials[4] = {0.0, -0.0, std::nan(""), 1.5};
  for (std::uint64_t i = 0; i < n; ++i) { v[i] = specials[rnd() % 4]; plain[i] = static_cast<double>(rnd() % 7); k[i] = static_cast<std::int32_t>(i); }
  CatalogRelation rel(130, "sf", {{"k", kInt}, {"v", kDouble}, {"p", kDouble}});
  sm->loadRelation(&rel, {k.data(), v.data(), plain.data()}, n, 6000, TupleStoreLayout::kCompressedColumnStore);
  sm->setCodeResident(true);
  qsgpu_relation_t dev = sm->deviceRelation(rel);
  sm->setCodeResident(false);
  const auto cv = sm->residentCoding(rel, 1), cp = sm->residentCoding(rel, 2);
  const auto got = readColumn<double>(dev, 1, n), gotp = readColumn<double>(dev, 2, n);
  const bool same = std::memcmp(got.data(), v.data(), n * 8) == 0 && std::memcmp(gotp.data(), plain.data(), n * 8) == 0;
  EXPECT(cv.first == 0);            // -0.0 / NaN: native
  EXPECT(cp.first == 1);            // seven ordinary values: 1-byte codes
  EXPECT(same);
  sm->evict(rel);
  std::printf("special_floats_stay_native %s\n", (cv.first == 0 && cp.first == 1 && same) ? "ok" : "MISMATCH");
}

int main() {
  int dev = 0;
  if (qsgpu_init(1, &dev) != 0) { std::printf("no CUDA device: %s\n", qsgpu_last_error()); return 2; }
  {
    StorageManager sm(0);
    WorkerPool pool(4);
    testOuterJoinCompositeKey(&sm, &pool);
    testExistenceMapAggregation(&sm, &pool);
    testPartitionAwareInsertDestination(&sm, &pool, 4);
    testPartitionAwareInsertDestination(&sm, &pool, 3);
    testNullableAggregation(&sm, &pool, TupleStoreLayout::kSplitRowStore, 100);
    testNullableAggregation(&sm, &pool, TupleStoreLayout::kBasicColumnStore, 110);
    testNullableAggregation(&sm, &pool, TupleStoreLayout::kCompressedColumnStore, 120);
    testSpecialFloatsStayNative(&sm);
  }
  if (g_failed) { std::printf("%d check(s) failed\n", g_failed); return 1; }
  std::printf("all host GPU tests passed\n");
  return 0;
}
