// Device-free unit tests of the host layer: expression-set plumbing and the scheduling contracts the GPU
// operators rely on (the reference checks the same contracts with a MockOperator in
// query_execution/tests/QueryManagerSingleNode_unittest.cpp).  Exit code 0 = all passed; prints one line per case.
#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#include "ExprSet.hpp"
#include "QueryManager.hpp"
#include "StorageManager.hpp"
#include "WorkOrder.hpp"

using namespace quickstep;

static int g_failed = 0;
#define EXPECT(cond)                                                                  \
  do {                                                                                \
    if (!(cond)) { std::printf("  FAILED %s (%s:%d)\n", #cond, __FILE__, __LINE__); ++g_failed; } \
  } while (0)

// ------------------------------------------------------------------ ExprSet
static void testExprSetAppend() {
  const qs_attr kInt{QS_INT, 4}, kChar{QS_CHAR, 10}, kDouble{QS_DOUBLE, 8};
  ExprSet pred;       // c_mktsegment = 'BUILDING' AND c_custkey > 5
  const int p = pred.conj({pred.cmp(QS_EQ, pred.attr(1, kChar), pred.lit_char("BUILDING")),
                           pred.cmp(QS_GT, pred.attr(0, kInt), pred.lit_int(5))});
  ExprSet sel;        // a3 * (1 - b.a4), 'X'
  const int s0 = sel.binary(QS_MUL, sel.attr(3, kDouble), sel.binary(QS_SUB, sel.lit_int(1), sel.attr(4, kDouble, 2)));
  const int s1 = sel.lit_char("X");
  ExprSet all;
  const int off_p = all.append(pred), off_s = all.append(sel);
  EXPECT(off_p == 0 && off_s == static_cast<int>(pred.size()));
  const qs_expr_set v = all.view();
  EXPECT(v.n_nodes == pred.size() + sel.size());
  // children of appended nodes are rebased, attribute ids are not
  const qs_node &mul = v.nodes[s0 + off_s];
  EXPECT(mul.kind == QS_N_BINARY && mul.op == QS_MUL);
  EXPECT(v.nodes[mul.a].kind == QS_N_ATTRIBUTE && v.nodes[mul.a].a == 3);
  const qs_node &sub = v.nodes[mul.b];
  EXPECT(sub.kind == QS_N_BINARY && sub.op == QS_SUB && v.nodes[sub.a].kind == QS_N_LITERAL && v.nodes[sub.b].a == 4 && v.nodes[sub.b].b == 2);
  // CHAR literals keep pointing at their own bytes in the merged pool
  const qs_node &x = v.nodes[s1 + off_s];
  EXPECT(x.kind == QS_N_LITERAL && x.type == QS_CHAR && x.width == 1 && v.str_pool[x.lit.pool_offset] == 'X');
  const qs_node &conj = v.nodes[p + off_p];
  EXPECT(conj.kind == QS_N_CONJUNCTION);
  const qs_node &eq = v.nodes[conj.a];
  EXPECT(eq.kind == QS_N_COMPARISON && std::strncmp(v.str_pool + v.nodes[eq.b].lit.pool_offset, "BUILDING", 8) == 0);
  // attribute masks per join side
  EXPECT(pred.referencedAttributes(0) == 0b11 && pred.referencedAttributes(2) == 0);
  EXPECT(sel.referencedAttributes(0) == (1ull << 3) && sel.referencedAttributes(2) == (1ull << 4));
  EXPECT(all.referencedAttributes(0) == (0b11 | (1ull << 3)));
  std::printf("expr_set_append ok\n");
}

// ------------------------------------------------------------- scheduling
struct Trace {
  std::mutex mu;
  std::vector<std::string> events;            // "<op>:<what>"
  void add(const std::string &e) { std::lock_guard<std::mutex> lk(mu); events.push_back(e); }
  int index(const std::string &e) {
    std::lock_guard<std::mutex> lk(mu);
    for (std::size_t i = 0; i < events.size(); ++i) if (events[i] == e) return static_cast<int>(i);
    return -1;
  }
  int count(const std::string &prefix) {
    std::lock_guard<std::mutex> lk(mu);
    int n = 0;
    for (const auto &e : events) n += e.rfind(prefix, 0) == 0;
    return n;
  }
};

class MockWorkOrder : public WorkOrder {
 public:
  MockWorkOrder(Trace *t, std::string tag, std::atomic<int> *running, std::atomic<int> *max_running)
      : WorkOrder(1), trace_(t), tag_(std::move(tag)), running_(running), max_running_(max_running) {}
  void execute() override {
    const int now = ++*running_;
    int seen = max_running_->load();
    while (now > seen && !max_running_->compare_exchange_weak(seen, now)) {}
    trace_->add(tag_);
    --*running_;
  }
 private:
  Trace *trace_;
  std::string tag_;
  std::atomic<int> *running_, *max_running_;
};

// Emits `own_work_orders` work orders on its first call (a stored input) plus one per block it is fed (a
// streamed input); with a streamed input it is done only after doneFeedingInputBlocks.
class MockOperator : public RelationalOperator {
 public:
  MockOperator(std::string name, Trace *t, int own_work_orders, bool has_streamed_input, relation_id output_rel,
               std::atomic<int> *running, std::atomic<int> *max_running)
      : RelationalOperator(1), name_(std::move(name)), trace_(t), own_(own_work_orders), streamed_(has_streamed_input),
        output_rel_(output_rel), running_(running), max_running_(max_running) {}
  OperatorType getOperatorType() const override { return kSelect; }
  std::string getName() const override { return name_; }
  bool getAllWorkOrders(WorkOrdersContainer *container, QueryContext *, StorageManager *, const tmb::client_id, tmb::MessageBus *) override {
    ++calls_;
    trace_->add(name_ + ":generate");
    if (!started_) {
      started_ = true;
      for (int i = 0; i < own_; ++i)
        container->addNormalWorkOrder(new MockWorkOrder(trace_, name_ + ":wo", running_, max_running_), getOperatorIndex());
    }
    for (block_id b : fed_) {
      (void)b;
      container->addNormalWorkOrder(new MockWorkOrder(trace_, name_ + ":wo_fed", running_, max_running_), getOperatorIndex());
    }
    fed_.clear();
    return streamed_ ? done_feeding_input_relation_ : true;
  }
  void feedInputBlock(const block_id b, const relation_id rel, const partition_id) override {
    trace_->add(name_ + ":fed");
    fed_rel_ = rel;
    fed_.push_back(b);
  }
  void doneFeedingInputBlocks(const relation_id rel) override {
    trace_->add(name_ + ":done_feeding");
    RelationalOperator::doneFeedingInputBlocks(rel);
  }
  QueryContext::insert_destination_id getInsertDestinationID() const override { return dest_; }
  const relation_id getOutputRelationID() const override { return output_rel_; }
  void setDestination(QueryContext::insert_destination_id d) { dest_ = d; }
  int calls() const { return calls_; }
  relation_id fedRelation() const { return fed_rel_; }
 private:
  std::string name_;
  Trace *trace_;
  int own_;
  bool streamed_, started_ = false;
  relation_id output_rel_, fed_rel_ = -1;
  QueryContext::insert_destination_id dest_ = QueryContext::kInvalidInsertDestinationId;
  std::vector<block_id> fed_;
  int calls_ = 0;
  std::atomic<int> *running_, *max_running_;
};

static void testBlockingDependency() {
  Trace t;
  std::atomic<int> running{0}, max_running{0};
  QueryContext ctx(nullptr, 0);
  QueryPlan plan;
  const auto a = plan.addRelationalOperator(new MockOperator("A", &t, 40, false, -1, &running, &max_running));
  const auto b = plan.addRelationalOperator(new MockOperator("B", &t, 3, false, -1, &running, &max_running));
  plan.addDirectDependency(b, a, /*is_pipeline_breaker=*/true);
  WorkerPool pool(4);
  QueryManager qm(&plan, &ctx, nullptr, &pool);
  qm.run();
  EXPECT(qm.numWorkOrdersExecuted(a) == 40 && qm.numWorkOrdersExecuted(b) == 3);
  EXPECT(t.count("A:wo") == 40 && t.count("B:wo") == 3);
  // B generates (and therefore runs) only after every work order of A completed
  int last_a = -1, first_b = 1 << 30;
  {
    std::lock_guard<std::mutex> lk(t.mu);
    for (std::size_t i = 0; i < t.events.size(); ++i) {
      if (t.events[i] == "A:wo") last_a = static_cast<int>(i);
      if (t.events[i] == "B:generate" || t.events[i] == "B:wo") first_b = std::min(first_b, static_cast<int>(i));
    }
  }
  EXPECT(last_a >= 0 && last_a < first_b);
  EXPECT(max_running.load() >= 1 && max_running.load() <= 4);
  std::printf("blocking_dependency ok (max concurrent work orders %d)\n", max_running.load());
}

static void testPipelinedFeed() {
  // Without a GPU there is no real InsertDestination; a destination over a temporary relation needs the storage
  // manager only when blocks are asked for, so the producer here has no destination and the consumer is fed by
  // hand through the same calls QueryManager::markOperatorFinished makes.
  Trace t;
  std::atomic<int> running{0}, max_running{0};
  QueryContext ctx(nullptr, 0);
  QueryPlan plan;
  auto *prod = new MockOperator("P", &t, 5, false, 77, &running, &max_running);
  auto *cons = new MockOperator("C", &t, 0, true, -1, &running, &max_running);
  const auto p = plan.addRelationalOperator(prod), c = plan.addRelationalOperator(cons);
  plan.addDirectDependency(c, p, /*is_pipeline_breaker=*/false);
  // consumer asked before anything was fed: no work orders, not done
  WorkOrdersContainer probe_container(plan.size());
  EXPECT(cons->getAllWorkOrders(&probe_container, &ctx, nullptr, 0, nullptr) == false);
  EXPECT(!probe_container.hasNormalWorkOrder(c));
  cons->feedInputBlock(1001, 77, 0);
  cons->feedInputBlock(1002, 77, 0);
  EXPECT(cons->getAllWorkOrders(&probe_container, &ctx, nullptr, 0, nullptr) == false);      // more may come
  EXPECT(probe_container.getNumNormalWorkOrders(c) == 2 && cons->fedRelation() == 77);
  cons->doneFeedingInputBlocks(77);
  EXPECT(cons->getAllWorkOrders(&probe_container, &ctx, nullptr, 0, nullptr) == true);
  EXPECT(probe_container.getNumNormalWorkOrders(c) == 2);                                      // nothing new
  while (WorkOrder *w = probe_container.getNormalWorkOrder(c)) { w->execute(); delete w; }
  EXPECT(t.count("C:wo_fed") == 2);
  std::printf("pipelined_feed ok (%d getAllWorkOrders calls on the consumer)\n", cons->calls());
}

static void testDiamondAndRepeatedCalls() {
  Trace t;
  std::atomic<int> running{0}, max_running{0};
  QueryContext ctx(nullptr, 0);
  QueryPlan plan;
  // S -> (L, R) -> J : J blocks on both L and R; L and R block on S
  const auto s = plan.addRelationalOperator(new MockOperator("S", &t, 8, false, -1, &running, &max_running));
  const auto l = plan.addRelationalOperator(new MockOperator("L", &t, 6, false, -1, &running, &max_running));
  const auto r = plan.addRelationalOperator(new MockOperator("R", &t, 7, false, -1, &running, &max_running));
  const auto j = plan.addRelationalOperator(new MockOperator("J", &t, 2, false, -1, &running, &max_running));
  plan.addDirectDependency(l, s, true);
  plan.addDirectDependency(r, s, true);
  plan.addDirectDependency(j, l, true);
  plan.addDirectDependency(j, r, true);
  WorkerPool pool(3);
  QueryManager qm(&plan, &ctx, nullptr, &pool);
  qm.run();
  EXPECT(qm.totalWorkOrdersExecuted() == 23);
  EXPECT(t.index("S:generate") < t.index("L:generate") && t.index("S:generate") < t.index("R:generate"));
  int last_lr = -1;
  {
    std::lock_guard<std::mutex> lk(t.mu);
    for (std::size_t i = 0; i < t.events.size(); ++i)
      if (t.events[i] == "L:wo" || t.events[i] == "R:wo") last_lr = static_cast<int>(i);
  }
  EXPECT(last_lr < t.index("J:wo"));
  const std::string prof = qm.profile();
  EXPECT(prof.find("work_orders=8") != std::string::npos && prof.find("work_orders=2") != std::string::npos);
  std::printf("diamond ok\n");
}

// ------------------------------------------------------------ block builder
// Decodes one stripe of a block image back to native values (independent of the CUDA decoders).
template <class T>
static std::vector<T> decodeStripe(const qs_stage_desc &s, std::uint64_t n) {
  std::vector<T> out(n);
  const unsigned char *h = static_cast<const unsigned char *>(s.host);
  for (std::uint64_t i = 0; i < n; ++i) {
    std::uint32_t code = 0;
    switch (s.encoding) {
      case QS_ENC_PLAIN: std::memcpy(&out[i], h + i * sizeof(T), sizeof(T)); break;
      case QS_ENC_STRIDED: std::memcpy(&out[i], h + i * s.stride, sizeof(T)); break;
      case QS_ENC_DICT:
        std::memcpy(&code, h + i * s.code_width, s.code_width);
        std::memcpy(&out[i], static_cast<const char *>(s.dict) + static_cast<std::size_t>(code) * sizeof(T), sizeof(T));
        break;
      default: {   // QS_ENC_TRUNCATED
        std::memcpy(&code, h + i * s.code_width, s.code_width);
        const T v = static_cast<T>(code);
        out[i] = v;
      }
    }
  }
  return out;
}

static void testBlockBuilder() {
  // 10,000 tuples: few distinct doubles (dictionary, 1-byte codes), small non-negative ints (truncation to one
  // byte), ~2,500 distinct dates-as-longs (dictionary, 2-byte codes), all-distinct doubles (kept native).
  const std::uint64_t n = 10000, rows_per_block = 4096;
  std::vector<double> few(n), distinct(n);
  std::vector<std::int32_t> small(n);
  std::vector<std::int64_t> mid(n);
  std::uint64_t x = 88172645463325252ull;
  auto rnd = [&] { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
  for (std::uint64_t i = 0; i < n; ++i) {
    few[i] = static_cast<double>(rnd() % 11) / 100.0;
    small[i] = static_cast<std::int32_t>(rnd() % 200);
    mid[i] = static_cast<std::int64_t>(rnd() % 2500) * 7 - 10000;        // some values negative: truncation is out
    distinct[i] = static_cast<double>(i) * 1.25 + 0.5;
  }
  const qs_attr kInt{QS_INT, 4}, kLong{QS_LONG, 8}, kDouble{QS_DOUBLE, 8};
  for (int layout = 0; layout < 3; ++layout) {
    CatalogRelation rel(1, "t", {{"few", kDouble}, {"small", kInt}, {"mid", kLong}, {"distinct", kDouble}});
    StorageManager sm(0, /*pinned_blocks=*/false);
    sm.loadRelation(&rel, {few.data(), small.data(), mid.data(), distinct.data()}, n, rows_per_block,
                    static_cast<TupleStoreLayout>(layout));
    const std::vector<block_id> ids = rel.getBlocksSnapshot();
    EXPECT(ids.size() == 3);
    std::uint64_t row = 0;
    for (block_id id : ids) {
      const StorageBlock &B = sm.getBlock(id);
      const std::uint64_t m = static_cast<std::uint64_t>(B.num_tuples);
      EXPECT(m == std::min<std::uint64_t>(rows_per_block, n - row) && B.stripes.size() == 4 && B.size % 16 == 0);
      if (layout == static_cast<int>(TupleStoreLayout::kCompressedColumnStore)) {
        // CompressedBlockBuilder's rule: the smaller of truncation / dictionary / native wins
        EXPECT(B.stripes[0].encoding == QS_ENC_DICT && B.stripes[0].code_width == 1 && B.stripes[0].dict_entries <= 11);
        EXPECT(B.stripes[1].encoding == QS_ENC_TRUNCATED && B.stripes[1].code_width == 1);
        EXPECT(B.stripes[2].encoding == QS_ENC_DICT && B.stripes[2].code_width == 2);
        EXPECT(B.stripes[3].encoding == QS_ENC_PLAIN);
        EXPECT(B.size < m * 24 * 3 / 4);             // 24 native bytes per tuple -> 12 of codes + the dictionaries
      } else if (layout == static_cast<int>(TupleStoreLayout::kSplitRowStore)) {
        for (const qs_stage_desc &s : B.stripes) EXPECT(s.encoding == QS_ENC_STRIDED && s.stride == 28);
      } else {
        for (const qs_stage_desc &s : B.stripes) EXPECT(s.encoding == QS_ENC_PLAIN);
      }
      for (const qs_stage_desc &s : B.stripes) {     // every stripe / dictionary lies inside the block image
        const char *p = static_cast<const char *>(s.host);
        EXPECT(p >= B.memory && p < B.memory + B.size);
        if (s.encoding == QS_ENC_DICT) EXPECT(static_cast<const char *>(s.dict) >= B.memory && static_cast<const char *>(s.dict) < B.memory + B.size);
      }
      const auto d0 = decodeStripe<double>(B.stripes[0], m), d3 = decodeStripe<double>(B.stripes[3], m);
      const auto d1 = decodeStripe<std::int32_t>(B.stripes[1], m);
      const auto d2 = decodeStripe<std::int64_t>(B.stripes[2], m);
      bool same = true;
      for (std::uint64_t i = 0; i < m; ++i)
        same = same && std::memcmp(&d0[i], &few[row + i], 8) == 0 && d1[i] == small[row + i] && d2[i] == mid[row + i] &&
               std::memcmp(&d3[i], &distinct[row + i], 8) == 0;
      EXPECT(same);
      row += m;
    }
    EXPECT(row == n);
  }
  std::printf("block_builder ok\n");
}


// ------------------------------------------------------------------ insertTuples: rows that leave the device
// StorageManager::insertTuples writes result rows as SplitRowStore blocks in the reference's on-disk layout.  Checked here
// with the layer's own reader; with a directory as argv[1] the block images are also written out, and
// tests/test_intree_boundary.py opens them with the REFERENCE's real StorageBlock class (tests/intree/read_blocks.cpp).
static void testInsertTuples(const char *dump_dir) {
  const std::uint64_t n = 100000;          // 38-byte tuples: ~55,000 per 2 MB block -> two blocks
  std::vector<std::int32_t> a(n);
  std::vector<double> b(n);
  std::vector<char> c(n * 10, '\0');
  struct Date { std::int32_t year; std::uint8_t month, day; std::uint8_t pad[2]; };
  std::vector<Date> d(n);
  std::vector<std::int64_t> e(n);
  for (std::uint64_t i = 0; i < n; ++i) {
    a[i] = static_cast<std::int32_t>(i) - 50000;
    b[i] = static_cast<double>(i) * 0.25 - 7.5;
    std::snprintf(&c[i * 10], 10, "row%llu", static_cast<unsigned long long>(i % 1000));
    d[i] = Date{1992 + static_cast<std::int32_t>(i % 7), static_cast<std::uint8_t>(1 + i % 12), static_cast<std::uint8_t>(1 + i % 28), {0, 0}};
    e[i] = static_cast<std::int64_t>(i) * 1000003ll - (1ll << 40);
  }
  CatalogRelation rel(9, "result", {{"a", qs_attr{QS_INT, 4}}, {"b", qs_attr{QS_DOUBLE, 8}}, {"c", qs_attr{QS_CHAR, 10}},
                                    {"d", qs_attr{QS_DATE, 8}}, {"e", qs_attr{QS_LONG, 8}}});
  StorageManager sm(0, /*pinned_blocks=*/false);
  const std::vector<block_id> ids = sm.insertTuples(rel, {a.data(), b.data(), c.data(), d.data(), e.data()}, n);
  EXPECT(ids.size() == 2 && sm.hostBlocksOf(rel) == ids);
  std::uint64_t row = 0;
  int file_no = 0;
  for (block_id id : ids) {
    const StorageBlock &B = sm.getBlock(id);
    SplitRowStoreReader reader(B.memory, rel.schema());
    EXPECT(reader.numTuples() == static_cast<std::uint64_t>(B.num_tuples));
    bool same = true;
    for (std::uint64_t i = 0; i < reader.numTuples(); ++i)
      same = same && std::memcmp(reader.value(i, 0), &a[row + i], 4) == 0 && std::memcmp(reader.value(i, 1), &b[row + i], 8) == 0 &&
             std::memcmp(reader.value(i, 2), &c[(row + i) * 10], 10) == 0 && std::memcmp(reader.value(i, 3), &d[row + i], 6) == 0 &&
             std::memcmp(reader.value(i, 4), &e[row + i], 8) == 0;
    EXPECT(same);
    if (dump_dir) {
      const std::string path = std::string(dump_dir) + "/block_" + std::to_string(file_no++) + ".bin";
      std::FILE *f = std::fopen(path.c_str(), "wb");
      EXPECT(f != nullptr);
      if (f) { EXPECT(std::fwrite(B.memory, 1, B.size, f) == B.size); std::fclose(f); }
    }
    row += reader.numTuples();
  }
  EXPECT(row == n);
  std::printf("insert_tuples_blocks ok\n");
}

int main(int argc, char **argv) {
  testExprSetAppend();
  testBlockBuilder();
  testInsertTuples(argc > 1 ? argv[1] : nullptr);
  testBlockingDependency();
  testPipelinedFeed();
  testDiamondAndRepeatedCalls();
  if (g_failed) { std::printf("%d check(s) failed\n", g_failed); return 1; }
  std::printf("all host unit tests passed\n");
  return 0;
}
