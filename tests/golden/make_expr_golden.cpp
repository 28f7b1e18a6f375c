// TEST INFRASTRUCTURE: generates tests/golden/reference_expressions.json with the UNMODIFIED reference's own classes.
//
// Linked against the static libraries of oracle/build_ref.sh's build tree (tests/golden/make_expr_golden.sh is the
// recipe).  For a fixed relation of random tuples it
//   * builds predicates and scalars out of the reference's real expression classes (ScalarAttribute, ScalarLiteral,
//     ScalarUnaryExpression, ScalarBinaryExpression, ScalarSharedExpression, ComparisonPredicate, NegationPredicate,
//     ConjunctionPredicate, DisjunctionPredicate),
//   * evaluates them with the reference's own vectorised code paths -- Predicate::getAllMatches
//     (expressions/predicate/Predicate.hpp:167) and Scalar::getAllValues (expressions/scalar/Scalar.hpp:215) -- over a
//     ColumnVectorsValueAccessor holding the tuples,
//   * lowers each through its getProto() with the in-tree binding's LowerPredicate / LowerScalar
//     (quickstep_b200/host/intree/ProtoLowering.hpp) into the C ABI's qs_node array,
// and writes tuples, node arrays and the reference's answers.  tests/test_reference_expressions.py evaluates the SAME
// node arrays with the oracle (CPU) and with the CUDA path (-m gpu) and demands the reference's bits.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "catalog/CatalogAttribute.hpp"
#include "catalog/CatalogRelation.hpp"
#include "catalog/PartitionSchemeHeader.hpp"
#include "compression/CompressionDictionary.hpp"
#include "compression/CompressionDictionaryBuilder.hpp"
#include "expressions/aggregation/AggregateFunction.hpp"
#include "expressions/aggregation/AggregateFunctionFactory.hpp"
#include "expressions/aggregation/AggregationHandle.hpp"
#include "expressions/aggregation/AggregationID.hpp"
#include "expressions/predicate/ComparisonPredicate.hpp"
#include "expressions/predicate/ConjunctionPredicate.hpp"
#include "expressions/predicate/DisjunctionPredicate.hpp"
#include "expressions/predicate/NegationPredicate.hpp"
#include "expressions/predicate/Predicate.hpp"
#include "expressions/scalar/Scalar.hpp"
#include "expressions/scalar/ScalarAttribute.hpp"
#include "expressions/scalar/ScalarBinaryExpression.hpp"
#include "expressions/scalar/ScalarLiteral.hpp"
#include "expressions/scalar/ScalarSharedExpression.hpp"
#include "expressions/scalar/ScalarUnaryExpression.hpp"
#include "storage/TupleIdSequence.hpp"
#include "storage/ValueAccessor.hpp"
#include "storage/ValueAccessorMultiplexer.hpp"
#include "types/DatetimeLit.hpp"
#include "types/Type.hpp"
#include "types/TypeFactory.hpp"
#include "types/TypeID.hpp"
#include "types/TypedValue.hpp"
#include "types/containers/ColumnVector.hpp"
#include "types/containers/ColumnVectorsValueAccessor.hpp"
#include "types/operations/binary_operations/BinaryOperationFactory.hpp"
#include "types/operations/binary_operations/BinaryOperationID.hpp"
#include "types/operations/comparisons/ComparisonFactory.hpp"
#include "types/operations/comparisons/ComparisonID.hpp"
#include "types/operations/unary_operations/NumericCastOperation.hpp"
#include "types/operations/unary_operations/UnaryOperationFactory.hpp"
#include "types/operations/unary_operations/UnaryOperationID.hpp"
#include "utility/ColumnVectorCache.hpp"
#include "utility/lip_filter/LIPFilter.hpp"
#include "utility/lip_filter/LIPFilter.pb.h"
#include "utility/lip_filter/LIPFilterFactory.hpp"

#include "ProtoLowering.hpp"

using namespace quickstep;  // NOLINT

namespace {

constexpr int kRows = 512;
constexpr int kRelationId = 7;
enum Attr { A_I, A_I2, A_L, A_F, A_D, A_DISC, A_TAX, A_DT, A_C1, A_C4, A_C10, A_L2, kNumAttrs };

std::uint64_t g_state = 0x9E3779B97F4A7C15ull;
std::uint64_t Next() {      // splitmix64
  std::uint64_t z = (g_state += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

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

struct Column {
  const Type *type;
  std::string name;
  std::vector<char> bytes;     // kRows * width, canonical (zero padding)
};

std::vector<Column> g_cols;
CatalogRelation *g_rel = nullptr;

// The same tuples as a second relation `tn` whose attributes i, l, d, dt, c4 are NULL-able: every fifth value or so is
// NULL (g_nulls: bit a of row r = attribute a is NULL).  g_nullable_mode selects which of the two the constructors and
// the accessor below refer to.
constexpr int kNullableRelationId = 8;
const int kNullableAttrs[] = {A_I, A_L, A_D, A_DT, A_C4};
CatalogRelation *g_rel_n = nullptr;
std::vector<const Type *> g_types_n;
std::vector<std::uint64_t> g_nulls;
bool g_nullable_mode = false;

void PutChar(std::vector<char> *out, std::size_t width, const char *s) {
  std::vector<char> v(width, '\0');
  std::memcpy(v.data(), s, std::min(width, std::strlen(s)));
  out->insert(out->end(), v.begin(), v.end());
}

template <typename T>
void Put(std::vector<char> *out, T v) {
  const char *p = reinterpret_cast<const char *>(&v);
  out->insert(out->end(), p, p + sizeof(T));
}

void MakeRelation() {
  g_rel = new CatalogRelation(nullptr, "t", kRelationId);
  const char *names[kNumAttrs] = {"i", "i2", "l", "f", "d", "disc", "tax", "dt", "c1", "c4", "c10", "l2"};
  const Type *types[kNumAttrs] = {
      &TypeFactory::GetType(kInt, false),    &TypeFactory::GetType(kInt, false),    &TypeFactory::GetType(kLong, false),
      &TypeFactory::GetType(kFloat, false),  &TypeFactory::GetType(kDouble, false), &TypeFactory::GetType(kDouble, false),
      &TypeFactory::GetType(kDouble, false), &TypeFactory::GetType(kDate, false),   &TypeFactory::GetType(kChar, 1, false),
      &TypeFactory::GetType(kChar, 4, false), &TypeFactory::GetType(kChar, 10, false), &TypeFactory::GetType(kLong, false)};
  g_cols.resize(kNumAttrs);
  for (int a = 0; a < kNumAttrs; ++a) {
    g_rel->addAttribute(new CatalogAttribute(g_rel, names[a], *types[a]));
    g_cols[a].type = types[a];
    g_cols[a].name = names[a];
  }
  const char *flags[3] = {"R", "A", "N"};
  const char *modes[7] = {"AIR", "MAIL", "SHIP", "RAIL", "FOB", "TRUC", "REG"};
  const char *segs[5] = {"BUILDING", "AUTOMOBILE", "MACHINERY", "HOUSEHOLD", "FURNITURE"};
  const double special[8] = {0.0, -0.0, std::numeric_limits<double>::quiet_NaN(), std::numeric_limits<double>::infinity(),
                             -std::numeric_limits<double>::infinity(), 1e-310 /* subnormal */, 1e308, -1.5};
  for (int r = 0; r < kRows; ++r) {
    Put<std::int32_t>(&g_cols[A_I].bytes, static_cast<std::int32_t>(Next() % 2001) - 1000);
    std::int32_t i2 = static_cast<std::int32_t>(Next() % 97) - 48;
    Put<std::int32_t>(&g_cols[A_I2].bytes, i2 == 0 ? 7 : i2);                            // divisors: never zero
    Put<std::int64_t>(&g_cols[A_L].bytes, static_cast<std::int64_t>(Next() % 20000000001ull) - 10000000000ll);
    Put<float>(&g_cols[A_F].bytes, static_cast<float>(static_cast<std::int64_t>(Next() % 200001) - 100000) / 64.0f);
    double d = 900.0 + static_cast<double>(Next() % 10409500) / 100.0;                   // l_extendedprice-like
    if (r % 16 == 5) d = special[(r / 16) % 8];
    Put<double>(&g_cols[A_D].bytes, d);
    Put<double>(&g_cols[A_DISC].bytes, static_cast<double>(Next() % 11) / 100.0);
    Put<double>(&g_cols[A_TAX].bytes, static_cast<double>(Next() % 9) / 100.0);
    DateLit dt;
    std::memset(&dt, 0, sizeof(dt));
    dt.year = 1992 + static_cast<std::int32_t>(Next() % 7);
    dt.month = static_cast<std::uint8_t>(1 + Next() % 12);
    dt.day = static_cast<std::uint8_t>(1 + Next() % 28);
    Put<DateLit>(&g_cols[A_DT].bytes, dt);
    PutChar(&g_cols[A_C1].bytes, 1, flags[Next() % 3]);
    PutChar(&g_cols[A_C4].bytes, 4, modes[Next() % 7]);
    PutChar(&g_cols[A_C10].bytes, 10, segs[Next() % 5]);
    Put<std::int64_t>(&g_cols[A_L2].bytes, static_cast<std::int64_t>(Next() % 601) - 300);
  }
  g_rel_n = new CatalogRelation(nullptr, "tn", kNullableRelationId);
  for (int a = 0; a < kNumAttrs; ++a) {
    bool nullable = false;
    for (int na : kNullableAttrs) nullable |= na == a;
    const Type &t = *types[a];
    const Type &tn = !nullable ? t
                     : t.getTypeID() == kChar ? TypeFactory::GetType(kChar, t.maximumByteLength(), true)
                                              : TypeFactory::GetType(t.getTypeID(), true);
    g_types_n.push_back(&tn);
    g_rel_n->addAttribute(new CatalogAttribute(g_rel_n, names[a], tn));
  }
  g_nulls.assign(kRows, 0);
  for (int r = 0; r < kRows; ++r)
    for (int na : kNullableAttrs)
      if (Next() % 5 == 0) g_nulls[r] |= 1ull << na;
}

ColumnVectorsValueAccessor *MakeAccessor() {
  ColumnVectorsValueAccessor *acc = new ColumnVectorsValueAccessor();
  for (int a = 0; a < kNumAttrs; ++a) {
    const Column &c = g_cols[a];
    const Type &type = g_nullable_mode ? *g_types_n[a] : *c.type;
    NativeColumnVector *cv = new NativeColumnVector(type, kRows);
    const std::size_t w = c.type->maximumByteLength();
    for (int r = 0; r < kRows; ++r) {
      if (g_nullable_mode && (g_nulls[r] >> a & 1))
        cv->appendNullValue();
      else
        cv->appendUntypedValue(c.bytes.data() + r * w);
    }
    acc->addColumn(cv);
  }
  return acc;
}

// ---- small constructors over the reference's classes
Scalar *Attr(int a) { return new ScalarAttribute(*(g_nullable_mode ? g_rel_n : g_rel)->getAttributeById(a)); }
Scalar *LitI(int v) { return new ScalarLiteral(TypedValue(v), TypeFactory::GetType(kInt, false)); }
Scalar *LitL(std::int64_t v) { return new ScalarLiteral(TypedValue(v), TypeFactory::GetType(kLong, false)); }
Scalar *LitF(float v) { return new ScalarLiteral(TypedValue(v), TypeFactory::GetType(kFloat, false)); }
Scalar *LitD(double v) { return new ScalarLiteral(TypedValue(v), TypeFactory::GetType(kDouble, false)); }
Scalar *LitDate(int y, int m, int d) {
  DateLit v;
  std::memset(&v, 0, sizeof(v));
  v.year = y;
  v.month = static_cast<std::uint8_t>(m);
  v.day = static_cast<std::uint8_t>(d);
  return new ScalarLiteral(TypedValue(v), TypeFactory::GetType(kDate, false));
}
Scalar *LitC(const char *s) {      // a CHAR(strlen) literal, as the parser types string literals
  const std::size_t n = std::strlen(s);
  TypedValue v(kChar, s, n);
  return new ScalarLiteral(v, TypeFactory::GetType(kChar, n, false));
}
Scalar *Bin(BinaryOperationID op, Scalar *l, Scalar *r) {
  return new ScalarBinaryExpression(BinaryOperationFactory::GetBinaryOperation(op), l, r);
}
Scalar *Neg(Scalar *x) { return new ScalarUnaryExpression(UnaryOperationFactory::GetUnaryOperation(UnaryOperationID::kNegate), x); }
Scalar *Cast(TypeID to, Scalar *x) {
  // the target type is NULL-able exactly when the operand is (NumericCastOperation::canApplyToType)
  const bool nullable = x->getType().isNullable();
  return new ScalarUnaryExpression(NumericCastOperation::Instance(TypeFactory::GetType(to, nullable)), x);
}
Scalar *Shared(int id, Scalar *x) { return new ScalarSharedExpression(id, x); }
Predicate *Cmp(ComparisonID op, Scalar *l, Scalar *r) {
  return new ComparisonPredicate(ComparisonFactory::GetComparison(op), l, r);
}
Predicate *And(std::vector<Predicate *> ps) {
  ConjunctionPredicate *c = new ConjunctionPredicate();
  for (Predicate *p : ps) c->addPredicate(p);
  return c;
}
Predicate *Or(std::vector<Predicate *> ps) {
  DisjunctionPredicate *c = new DisjunctionPredicate();
  for (Predicate *p : ps) c->addPredicate(p);
  return c;
}
Predicate *Not(Predicate *p) { return new NegationPredicate(p); }

constexpr BinaryOperationID ADD = BinaryOperationID::kAdd, SUB = BinaryOperationID::kSubtract, MUL = BinaryOperationID::kMultiply,
                            DIV = BinaryOperationID::kDivide, MOD = BinaryOperationID::kModulo;
constexpr ComparisonID EQ = ComparisonID::kEqual, NE = ComparisonID::kNotEqual, LT = ComparisonID::kLess,
                       LE = ComparisonID::kLessOrEqual, GT = ComparisonID::kGreater, GE = ComparisonID::kGreaterOrEqual;

bool g_first_case = true;

void EmitNodes(FILE *out, const gpu::ExprBuilder &b, int root) {
  const qs_expr_set es = b.view();
  std::fprintf(out, "\"root\": %d, \"pool\": \"%s\", \"nodes\": [", root, Hex(es.str_pool, es.str_pool_bytes).c_str());
  for (std::uint32_t i = 0; i < es.n_nodes; ++i) {
    const qs_node &n = es.nodes[i];
    std::fprintf(out, "%s[%u, %u, %u, %u, %d, %d, \"%s\"]", i ? ", " : "", n.kind, n.op, n.type, n.width, n.a, n.b,
                 Hex(&n.lit, sizeof(n.lit)).c_str());
  }
  std::fprintf(out, "]");
}

gpu::AttributeTypes Types() {
  gpu::AttributeTypes t;
  std::vector<qs_attr> attrs;
  for (const Column &c : g_cols) {
    qs_attr a{};
    a.type = static_cast<std::uint16_t>(c.type->getTypeID() == kDate ? QS_DATE : static_cast<int>(c.type->getTypeID()));
    a.width = static_cast<std::uint16_t>(c.type->maximumByteLength());
    attrs.push_back(a);
  }
  t.relations.emplace_back(g_nullable_mode ? kNullableRelationId : kRelationId, attrs);
  return t;
}

void PredicateCase(FILE *out, const char *name, const char *sql, Predicate *p) {
  std::unique_ptr<Predicate> owner(p);
  std::unique_ptr<ColumnVectorsValueAccessor> acc(MakeAccessor());
  std::unique_ptr<TupleIdSequence> matches(p->getAllMatches(acc.get(), nullptr, nullptr, nullptr));
  std::string bits;
  for (int r = 0; r < kRows; ++r) bits.push_back(matches->get(r) ? '1' : '0');
  gpu::ExprBuilder b;
  const int root = gpu::LowerPredicate(p->getProto(), Types(), &b);
  std::fprintf(out, "%s\n  {\"name\": \"%s%s\", \"sql\": \"%s\", \"kind\": \"predicate\", \"nullable\": %s, ", g_first_case ? "" : ",",
               g_nullable_mode ? "nullable_" : "", name, sql, g_nullable_mode ? "true" : "false");
  EmitNodes(out, b, root);
  std::fprintf(out, ", \"matches\": \"%s\", \"n_matches\": %zu}", bits.c_str(), static_cast<std::size_t>(matches->numTuples()));
  g_first_case = false;
}

void ScalarCase(FILE *out, const char *name, const char *sql, Scalar *s) {
  std::unique_ptr<Scalar> owner(s);
  std::unique_ptr<ColumnVectorsValueAccessor> acc(MakeAccessor());
  ColumnVectorCache cache;
  ColumnVectorPtr values = s->getAllValues(acc.get(), nullptr, &cache);
  const Type &t = s->getType();
  const std::size_t w = t.maximumByteLength();
  std::string bytes, value_nulls;
  CHECK(values->isNative());
  const NativeColumnVector &ncv = static_cast<const NativeColumnVector &>(*values);
  CHECK_EQ(static_cast<std::size_t>(kRows), ncv.size());
  const std::vector<char> zeros(w, '\0');
  for (int r = 0; r < kRows; ++r) {
    const void *v = ncv.getUntypedValue(r);           // nullptr: the value is NULL
    bytes += Hex(v ? v : zeros.data(), w);
    value_nulls.push_back(v ? '0' : '1');
  }
  gpu::ExprBuilder b;
  const int root = gpu::LowerScalar(s->getProto(), Types(), &b);
  std::fprintf(out, "%s\n  {\"name\": \"%s%s\", \"sql\": \"%s\", \"kind\": \"scalar\", \"nullable\": %s, ", g_first_case ? "" : ",",
               g_nullable_mode ? "nullable_" : "", name, sql, g_nullable_mode ? "true" : "false");
  EmitNodes(out, b, root);
  std::fprintf(out, ", \"result_type\": %d, \"result_width\": %zu, \"result_nullable\": %s, \"values\": \"%s\", \"value_nulls\": \"%s\"}",
               t.getTypeID() == kDate ? QS_DATE : static_cast<int>(t.getTypeID()), w, t.isNullable() ? "true" : "false", bytes.c_str(),
               value_nulls.c_str());
  g_first_case = false;
}


// ---------------------------------------------------------------------------------------------------------------------
// Aggregation handles (A6): AggregateFunction::createHandle -> accumulateValueAccessor over two halves of the relation ->
// mergeStates -> finalize (expressions/aggregation/AggregationHandle.hpp:140-209), the path of an aggregate without
// GROUP BY (storage/AggregationOperationState.cpp:476-520, 641-680).
// ---------------------------------------------------------------------------------------------------------------------
bool g_first_agg = true;

void AggregateCase(FILE *out, const char *name, const char *sql, AggregationID fn, Scalar *argument /* null: COUNT(*) */) {
  std::unique_ptr<Scalar> owner(argument);
  const int half = kRows / 2;
  std::vector<const Type *> arg_types;
  if (argument) arg_types.push_back(&argument->getType());
  std::unique_ptr<AggregationHandle> handle(AggregateFunctionFactory::Get(fn).createHandle(arg_types));
  std::unique_ptr<AggregationState> total(handle->createInitialState());
  ColumnVectorPtr values;
  std::unique_ptr<ColumnVectorsValueAccessor> acc(MakeAccessor());
  ColumnVectorCache cache;
  if (argument) values = argument->getAllValues(acc.get(), nullptr, &cache);
  for (int part = 0; part < 2; ++part) {                 // two "work orders": rows [0, 256) and [256, 512)
    std::unique_ptr<AggregationState> state;
    if (!argument) {
      state.reset(handle->accumulateNullary(half));
    } else {
      const NativeColumnVector &all = static_cast<const NativeColumnVector &>(*values);
      NativeColumnVector *piece = new NativeColumnVector(argument->getType(), half);
      for (int r = part * half; r < (part + 1) * half; ++r) {
        const void *v = all.getUntypedValue(r);
        if (v) piece->appendUntypedValue(v); else piece->appendNullValue();
      }
      ColumnVectorsValueAccessor piece_acc;
      piece_acc.addColumn(piece);
      state.reset(handle->accumulateValueAccessor({MultiSourceAttributeId(ValueAccessorSource::kBase, 0)},
                                                  ValueAccessorMultiplexer(&piece_acc)));
    }
    handle->mergeStates(*state, total.get());
  }
  const TypedValue result = handle->finalize(*total);
  const Type &rt = *handle->getResultType();
  char raw[8] = {0};
  if (!result.isNull()) result.copyInto(raw);
  gpu::ExprBuilder b;
  const int root = argument ? gpu::LowerScalar(argument->getProto(), Types(), &b) : -1;
  std::fprintf(out, "%s\n  {\"name\": \"%s%s\", \"sql\": \"%s\", \"nullable\": %s, \"function\": %d, ", g_first_agg ? "" : ",",
               g_nullable_mode ? "nullable_" : "", name, sql, g_nullable_mode ? "true" : "false", static_cast<int>(fn));
  EmitNodes(out, b, root);
  std::fprintf(out, ", \"result_type\": %d, \"result_null\": %s, \"result\": \"%s\"}", static_cast<int>(rt.getTypeID()),
               result.isNull() ? "true" : "false", Hex(raw, rt.maximumByteLength()).c_str());
  g_first_agg = false;
}

// ---------------------------------------------------------------------------------------------------------------------
// LIP filters (L1, L2, L4): LIPFilterFactory::ReconstructFromProto from the proto ExecutionGenerator writes, built with
// insertValueAccessor over the tuples a build-side predicate keeps, probed with filterBatch over all tuples
// (utility/lip_filter/BitVectorExactFilter.hpp:76-176, SingleIdentityHashFilter.hpp:62-171).
// ---------------------------------------------------------------------------------------------------------------------
bool g_first_lip = true;

void LipCase(FILE *out, const char *name, bool exact, std::int64_t min_value, std::int64_t max_value, std::uint64_t cardinality,
             bool is_anti, Predicate *build_predicate, int build_attr, int probe_attr) {
  std::unique_ptr<Predicate> owner(build_predicate);
  const std::vector<const Type *> &types_n = g_types_n;
  const Type &build_type = g_nullable_mode ? *types_n[build_attr] : *g_cols[build_attr].type;
  const Type &probe_type = g_nullable_mode ? *types_n[probe_attr] : *g_cols[probe_attr].type;
  serialization::LIPFilter proto;
  if (exact) {
    proto.set_lip_filter_type(serialization::LIPFilterType::BIT_VECTOR_EXACT_FILTER);
    proto.SetExtension(serialization::BitVectorExactFilter::min_value, min_value);
    proto.SetExtension(serialization::BitVectorExactFilter::max_value, max_value);
    proto.SetExtension(serialization::BitVectorExactFilter::attribute_size, build_type.maximumByteLength());
    proto.SetExtension(serialization::BitVectorExactFilter::is_anti_filter, is_anti);
  } else {
    proto.set_lip_filter_type(serialization::LIPFilterType::SINGLE_IDENTITY_HASH_FILTER);
    proto.SetExtension(serialization::SingleIdentityHashFilter::filter_cardinality, cardinality);
    proto.SetExtension(serialization::SingleIdentityHashFilter::attribute_size, build_type.maximumByteLength());
  }
  CHECK(LIPFilterFactory::ProtoIsValid(proto));
  std::unique_ptr<LIPFilter> filter(LIPFilterFactory::ReconstructFromProto(proto));
  std::unique_ptr<ColumnVectorsValueAccessor> acc(MakeAccessor());
  std::unique_ptr<TupleIdSequence> keep(build_predicate->getAllMatches(acc.get(), nullptr, nullptr, nullptr));
  std::unique_ptr<ValueAccessor> build_rows(acc->createSharedTupleIdSequenceAdapterVirtual(*keep));
  filter->insertValueAccessor(build_rows.get(), build_attr, &build_type);
  std::vector<tuple_id> batch(kRows);
  for (int r = 0; r < kRows; ++r) batch[r] = r;
  const std::size_t kept = filter->filterBatch(acc.get(), probe_attr, probe_type.isNullable(), &batch, kRows);
  std::string bits(kRows, '0');
  for (std::size_t k = 0; k < kept; ++k) bits[batch[k]] = '1';
  gpu::ExprBuilder b;
  const int root = gpu::LowerPredicate(build_predicate->getProto(), Types(), &b);
  std::fprintf(out, "%s\n  {\"name\": \"%s%s\", \"nullable\": %s, \"exact\": %s, \"min_value\": %lld, \"max_value\": %lld, "
               "\"cardinality\": %llu, \"is_anti\": %s, \"attribute_size\": %zu, \"build_attr\": %d, \"probe_attr\": %d, ",
               g_first_lip ? "" : ",", g_nullable_mode ? "nullable_" : "", name, g_nullable_mode ? "true" : "false",
               exact ? "true" : "false", static_cast<long long>(min_value), static_cast<long long>(max_value),
               static_cast<unsigned long long>(cardinality), is_anti ? "true" : "false", build_type.maximumByteLength(), build_attr,
               probe_attr);
  EmitNodes(out, b, root);
  std::fprintf(out, ", \"n_built_from\": %zu, \"passes\": \"%s\", \"n_passes\": %zu}", static_cast<std::size_t>(keep->numTuples()),
               bits.c_str(), kept);
  g_first_lip = false;
}


// ---------------------------------------------------------------------------------------------------------------------
// Dictionary limit codes (f2): a block's sorted CompressionDictionary built by the reference's own builder from the
// column's values, and the code range CompressedTupleStorageSubBlock::getMatchesForPredicate evaluates `attribute <cmp>
// literal` with (compression/CompressionDictionary.hpp:209-236, CompressionDictionary.cpp:201-330) -- literals of the
// attribute's own type and of other comparable types.
// ---------------------------------------------------------------------------------------------------------------------
bool g_first_dict = true;

void DictionaryCase(FILE *out, int attr, const std::vector<std::pair<const char *, Scalar *>> &literals) {
  const Type &type = *g_cols[attr].type;
  const std::size_t w = type.maximumByteLength();
  CompressionDictionaryBuilder builder(type);
  for (int r = 0; r < kRows; ++r) builder.insertEntry(type.makeValue(g_cols[attr].bytes.data() + r * w, w));
  std::vector<char> memory(builder.dictionarySizeBytes());
  builder.buildDictionary(memory.data());
  CompressionDictionary dict(type, memory.data(), memory.size());
  std::string entries;
  for (std::uint32_t c = 0; c < dict.numberOfCodes(); ++c) {
    std::vector<char> v(w, '\0');
    const TypedValue tv = dict.getTypedValueForCode(c);
    if (type.getTypeID() == kChar) std::memcpy(v.data(), tv.getOutOfLineData(), std::min<std::size_t>(w, tv.getAsciiStringLength()));
    else tv.copyInto(v.data());
    entries += Hex(v.data(), w);
  }
  std::fprintf(out, "%s\n  {\"attr\": %d, \"name\": \"%s\", \"type\": %d, \"width\": %zu, \"n_codes\": %u, \"entries\": \"%s\", \"comparisons\": [",
               g_first_dict ? "" : ",", attr, g_cols[attr].name.c_str(),
               type.getTypeID() == kDate ? QS_DATE : static_cast<int>(type.getTypeID()), w, dict.numberOfCodes(), entries.c_str());
  g_first_dict = false;
  const ComparisonID cmps[5] = {EQ, LT, LE, GT, GE};          // kNotEqual is the caller's complement of kEqual (:217-219)
  bool first = true;
  for (const auto &lit : literals) {
    std::unique_ptr<Scalar> owner(lit.second);
    CHECK(lit.second->hasStaticValue());
    const TypedValue &value = lit.second->getStaticValue();
    gpu::ExprBuilder b;
    const int root = gpu::LowerScalar(lit.second->getProto(), Types(), &b);
    for (const ComparisonID cmp : cmps) {
      const std::pair<std::uint32_t, std::uint32_t> limits = dict.getLimitCodesForComparisonTyped(cmp, value, lit.second->getType());
      std::fprintf(out, "%s\n    {\"literal\": \"%s\", \"cmp\": %d, \"first\": %u, \"second\": %u, ", first ? "" : ",", lit.first, static_cast<int>(cmp),
                   limits.first, limits.second);
      EmitNodes(out, b, root);
      std::fprintf(out, "}");
      first = false;
    }
  }
  std::fprintf(out, "]}");
}

Predicate *IsR() { return Cmp(EQ, Attr(A_C1), LitC("R")); }
Predicate *InRange(int a, int lo, int hi) { return And({Cmp(GE, Attr(a), LitI(lo)), Cmp(LE, Attr(a), LitI(hi))}); }

void AllAggregates(FILE *out) {
  struct Fn { AggregationID id; const char *name; } fns[5] = {{AggregationID::kSum, "sum"}, {AggregationID::kAvg, "avg"},
      {AggregationID::kMin, "min"}, {AggregationID::kMax, "max"}, {AggregationID::kCount, "count"}};
  struct Arg { int attr; const char *name; } args[5] = {{A_I, "i"}, {A_L, "l"}, {A_F, "f"}, {A_DISC, "disc"}, {A_L2, "l2"}};
  for (const Fn &fn : fns)
    for (const Arg &arg : args) {
      const std::string name = std::string(fn.name) + "_" + arg.name, sql = std::string(fn.name) + "(" + arg.name + ")";
      AggregateCase(out, name.c_str(), sql.c_str(), fn.id, Attr(arg.attr));
    }
  AggregateCase(out, "count_star", "count(*)", AggregationID::kCount, nullptr);
  AggregateCase(out, "sum_disc_price", "sum(d2 * (1 - disc))", AggregationID::kSum,
                Bin(MUL, Attr(A_TAX), Bin(SUB, LitI(1), Attr(A_DISC))));
  AggregateCase(out, "avg_expr_mixed", "avg(i * disc)", AggregationID::kAvg, Bin(MUL, Attr(A_I), Attr(A_DISC)));
  AggregateCase(out, "sum_int_expr", "sum(i + i2)", AggregationID::kSum, Bin(ADD, Attr(A_I), Attr(A_I2)));
  AggregateCase(out, "min_negated", "min(-l)", AggregationID::kMin, Neg(Attr(A_L)));
  AggregateCase(out, "max_float_expr", "max(f * 2.5)", AggregationID::kMax, Bin(MUL, Attr(A_F), LitF(2.5f)));
  AggregateCase(out, "sum_with_nan_and_inf", "sum(d)", AggregationID::kSum, Attr(A_D));
  AggregateCase(out, "count_char", "count(c4)", AggregationID::kCount, Attr(A_C4));
}

}  // namespace

int main(int argc, char **argv) {
  FILE *out = argc > 1 ? std::fopen(argv[1], "w") : stdout;
  MakeRelation();
  std::fprintf(out, "{\"generator\": \"tests/golden/make_expr_golden.cpp against the unmodified reference (fee4c630)\",\n");
  std::fprintf(out, " \"relation_id\": %d, \"n_rows\": %d,\n \"columns\": [", kRelationId, kRows);
  for (int a = 0; a < kNumAttrs; ++a) {
    const Column &c = g_cols[a];
    std::fprintf(out, "%s\n  {\"name\": \"%s\", \"type\": %d, \"width\": %zu, \"data\": \"%s\"}", a ? "," : "", c.name.c_str(),
                 c.type->getTypeID() == kDate ? QS_DATE : static_cast<int>(c.type->getTypeID()), c.type->maximumByteLength(),
                 Hex(c.bytes.data(), c.bytes.size()).c_str());
  }
  std::fprintf(out, "],\n \"nullable_relation_id\": %d, \"nullable_attributes\": [", kNullableRelationId);
  for (std::size_t k = 0; k < sizeof(kNullableAttrs) / sizeof(kNullableAttrs[0]); ++k) std::fprintf(out, "%s%d", k ? ", " : "", kNullableAttrs[k]);
  std::fprintf(out, "],\n \"nulls\": \"%s\",\n \"cases\": [", Hex(g_nulls.data(), g_nulls.size() * 8).c_str());

  // ------------------------------------------------------------------ scalars (E1: Scalar::getAllValues)
  ScalarCase(out, "int_plus_literal", "i + 1", Bin(ADD, Attr(A_I), LitI(1)));
  ScalarCase(out, "int_times_long", "i * l", Bin(MUL, Attr(A_I), Attr(A_L)));
  ScalarCase(out, "long_minus_int", "l - i", Bin(SUB, Attr(A_L), Attr(A_I)));
  ScalarCase(out, "float_times_literal", "f * 2.5", Bin(MUL, Attr(A_F), LitF(2.5f)));
  ScalarCase(out, "float_plus_double", "f + d", Bin(ADD, Attr(A_F), Attr(A_D)));
  ScalarCase(out, "int_plus_float", "i + f", Bin(ADD, Attr(A_I), Attr(A_F)));
  ScalarCase(out, "long_times_double", "l * disc", Bin(MUL, Attr(A_L), Attr(A_DISC)));
  ScalarCase(out, "q1_disc_price", "d * (1 - disc)", Bin(MUL, Attr(A_D), Bin(SUB, LitI(1), Attr(A_DISC))));
  ScalarCase(out, "q1_charge", "d * (1 - disc) * (1 + tax)",
             Bin(MUL, Bin(MUL, Attr(A_D), Bin(SUB, LitI(1), Attr(A_DISC))), Bin(ADD, LitI(1), Attr(A_TAX))));
  ScalarCase(out, "q1_charge_shared", "SHARED#3(d * (1 - disc)) * (1 + tax)",
             Bin(MUL, Shared(3, Bin(MUL, Attr(A_D), Bin(SUB, LitI(1), Attr(A_DISC)))), Bin(ADD, LitI(1), Attr(A_TAX))));
  ScalarCase(out, "shared_nested_right", "(i + 1) * (SHARED#7(i2 + 2) + 3)",
             Bin(MUL, Bin(ADD, Attr(A_I), LitI(1)), Bin(ADD, Shared(7, Bin(ADD, Attr(A_I2), LitI(2))), LitI(3))));
  ScalarCase(out, "q6_revenue", "d * disc", Bin(MUL, Attr(A_D), Attr(A_DISC)));
  ScalarCase(out, "negate_double", "-d", Neg(Attr(A_D)));
  ScalarCase(out, "negate_int", "-i", Neg(Attr(A_I)));
  ScalarCase(out, "negate_long_expr", "-(l + i)", Neg(Bin(ADD, Attr(A_L), Attr(A_I))));
  ScalarCase(out, "int_divide", "i / i2", Bin(DIV, Attr(A_I), Attr(A_I2)));
  ScalarCase(out, "long_divide_int", "l / i2", Bin(DIV, Attr(A_L), Attr(A_I2)));
  ScalarCase(out, "double_divide", "d / 7.0", Bin(DIV, Attr(A_D), LitD(7.0)));
  ScalarCase(out, "double_divide_by_column", "d / disc", Bin(DIV, Attr(A_D), Attr(A_DISC)));       // x / 0.0 -> inf / nan
  ScalarCase(out, "int_modulo", "i % i2", Bin(MOD, Attr(A_I), Attr(A_I2)));
  ScalarCase(out, "long_modulo", "l % 1000", Bin(MOD, Attr(A_L), LitL(1000)));
  ScalarCase(out, "cast_int_to_double", "CAST(i AS DOUBLE) * disc", Bin(MUL, Cast(kDouble, Attr(A_I)), Attr(A_DISC)));
  ScalarCase(out, "cast_long_to_float", "CAST(l AS FLOAT)", Cast(kFloat, Attr(A_L)));
  ScalarCase(out, "cast_float_to_int", "CAST(f AS INT)", Cast(kInt, Attr(A_F)));
  ScalarCase(out, "cast_int_to_long_times_long", "CAST(i AS LONG) * l", Bin(MUL, Cast(kLong, Attr(A_I)), Attr(A_L)));
  ScalarCase(out, "nested_mixed", "i + l * 2 - 3", Bin(SUB, Bin(ADD, Attr(A_I), Bin(MUL, Attr(A_L), LitI(2))), LitI(3)));
  ScalarCase(out, "literal_only_arith", "i + (2 + 3)", Bin(ADD, Attr(A_I), Bin(ADD, LitI(2), LitI(3))));
  ScalarCase(out, "double_sum_three", "d + disc + tax", Bin(ADD, Bin(ADD, Attr(A_D), Attr(A_DISC)), Attr(A_TAX)));
  ScalarCase(out, "bare_attribute_char", "c4", Attr(A_C4));
  ScalarCase(out, "bare_attribute_date", "dt", Attr(A_DT));

  // ------------------------------------------------------------------ predicates (P1 / P2: getAllMatches)
  PredicateCase(out, "int_lt", "i < 100", Cmp(LT, Attr(A_I), LitI(100)));
  PredicateCase(out, "int_ge_negative", "i >= -250", Cmp(GE, Attr(A_I), LitI(-250)));
  PredicateCase(out, "long_eq_never", "l = 12345", Cmp(EQ, Attr(A_L), LitL(12345)));
  PredicateCase(out, "long_gt_int_literal", "l > 0", Cmp(GT, Attr(A_L), LitI(0)));
  PredicateCase(out, "double_le", "disc <= 0.07", Cmp(LE, Attr(A_DISC), LitD(0.07)));
  PredicateCase(out, "double_eq_exact_step", "disc = 0.05", Cmp(EQ, Attr(A_DISC), LitD(0.05)));
  PredicateCase(out, "double_ne_special", "d <> 0.0", Cmp(NE, Attr(A_D), LitD(0.0)));             // NaN <> 0 true, -0.0 <> 0 false
  PredicateCase(out, "double_lt_special", "d < 1000.0", Cmp(LT, Attr(A_D), LitD(1000.0)));          // NaN false, -inf true
  PredicateCase(out, "double_ge_special", "d >= 0.0", Cmp(GE, Attr(A_D), LitD(0.0)));               // -0.0 true, NaN false
  PredicateCase(out, "float_gt", "f > 1.5", Cmp(GT, Attr(A_F), LitF(1.5f)));
  PredicateCase(out, "float_vs_double_literal", "f <= 0.1", Cmp(LE, Attr(A_F), LitD(0.1)));
  PredicateCase(out, "int_vs_double_literal", "i < 0.5", Cmp(LT, Attr(A_I), LitD(0.5)));
  PredicateCase(out, "date_le_q1", "dt <= DATE '1998-09-02'", Cmp(LE, Attr(A_DT), LitDate(1998, 9, 2)));
  PredicateCase(out, "date_range_q6", "dt >= DATE '1994-01-01' AND dt < DATE '1995-01-01'",
                And({Cmp(GE, Attr(A_DT), LitDate(1994, 1, 1)), Cmp(LT, Attr(A_DT), LitDate(1995, 1, 1))}));
  PredicateCase(out, "date_eq_month_boundary", "dt = DATE '1995-03-15'", Cmp(EQ, Attr(A_DT), LitDate(1995, 3, 15)));
  PredicateCase(out, "date_gt_q3", "dt > DATE '1995-03-15'", Cmp(GT, Attr(A_DT), LitDate(1995, 3, 15)));
  PredicateCase(out, "char1_eq", "c1 = 'R'", Cmp(EQ, Attr(A_C1), LitC("R")));
  PredicateCase(out, "char1_ne", "c1 <> 'N'", Cmp(NE, Attr(A_C1), LitC("N")));
  PredicateCase(out, "char4_lt_same_width", "c4 < 'MAIL'", Cmp(LT, Attr(A_C4), LitC("MAIL")));
  PredicateCase(out, "char4_eq_shorter_literal", "c4 = 'AIR'", Cmp(EQ, Attr(A_C4), LitC("AIR")));
  PredicateCase(out, "char4_ge_shorter_literal", "c4 >= 'RA'", Cmp(GE, Attr(A_C4), LitC("RA")));
  PredicateCase(out, "char10_eq_q3", "c10 = 'BUILDING'", Cmp(EQ, Attr(A_C10), LitC("BUILDING")));
  PredicateCase(out, "char10_eq_full_width", "c10 = 'AUTOMOBILE'", Cmp(EQ, Attr(A_C10), LitC("AUTOMOBILE")));
  PredicateCase(out, "char10_gt", "c10 > 'HOUSEHOLD'", Cmp(GT, Attr(A_C10), LitC("HOUSEHOLD")));
  PredicateCase(out, "literal_on_the_left", "5 < i", Cmp(LT, LitI(5), Attr(A_I)));
  PredicateCase(out, "attr_vs_attr_int_long", "i < l", Cmp(LT, Attr(A_I), Attr(A_L)));
  PredicateCase(out, "attr_vs_attr_double_float", "d > f", Cmp(GT, Attr(A_D), Attr(A_F)));
  PredicateCase(out, "attr_vs_attr_int_int", "i = i2", Cmp(EQ, Attr(A_I), Attr(A_I2)));
  PredicateCase(out, "attr_vs_attr_double", "disc <= tax", Cmp(LE, Attr(A_DISC), Attr(A_TAX)));
  PredicateCase(out, "expr_vs_expr", "i + 1 < l * 2", Cmp(LT, Bin(ADD, Attr(A_I), LitI(1)), Bin(MUL, Attr(A_L), LitI(2))));
  PredicateCase(out, "expr_vs_literal", "d * (1 - disc) > 50000.0",
                Cmp(GT, Bin(MUL, Attr(A_D), Bin(SUB, LitI(1), Attr(A_DISC))), LitD(50000.0)));
  PredicateCase(out, "q6_where", "dt >= '1994-01-01' AND dt < '1995-01-01' AND disc >= 0.05 AND disc <= 0.07 AND i2 < 24",
                And({Cmp(GE, Attr(A_DT), LitDate(1994, 1, 1)), Cmp(LT, Attr(A_DT), LitDate(1995, 1, 1)),
                     Cmp(GE, Attr(A_DISC), LitD(0.05)), Cmp(LE, Attr(A_DISC), LitD(0.07)), Cmp(LT, Attr(A_I2), LitI(24))}));
  PredicateCase(out, "or_of_three", "i < -900 OR c1 = 'A' OR disc = 0.1",
                Or({Cmp(LT, Attr(A_I), LitI(-900)), Cmp(EQ, Attr(A_C1), LitC("A")), Cmp(EQ, Attr(A_DISC), LitD(0.1))}));
  PredicateCase(out, "not_or_and", "NOT (i < 10 OR d > 50000.0) AND c1 <> 'N'",
                And({Not(Or({Cmp(LT, Attr(A_I), LitI(10)), Cmp(GT, Attr(A_D), LitD(50000.0))})), Cmp(NE, Attr(A_C1), LitC("N"))}));
  PredicateCase(out, "not_of_nan_comparison", "NOT (d < 1000.0)", Not(Cmp(LT, Attr(A_D), LitD(1000.0))));   // NaN rows: true
  PredicateCase(out, "and_inside_or", "(c4 = 'MAIL' AND tax < 0.03) OR (c4 = 'SHIP' AND tax > 0.05)",
                Or({And({Cmp(EQ, Attr(A_C4), LitC("MAIL")), Cmp(LT, Attr(A_TAX), LitD(0.03))}),
                    And({Cmp(EQ, Attr(A_C4), LitC("SHIP")), Cmp(GT, Attr(A_TAX), LitD(0.05))})}));
  PredicateCase(out, "nested_not_not", "NOT (NOT (i2 > 0))", Not(Not(Cmp(GT, Attr(A_I2), LitI(0)))));
  PredicateCase(out, "always_false_conjunction", "i < 0 AND i > 0", And({Cmp(LT, Attr(A_I), LitI(0)), Cmp(GT, Attr(A_I), LitI(0))}));
  PredicateCase(out, "cast_in_comparison", "CAST(i AS DOUBLE) / 3.0 >= disc * 100.0",
                Cmp(GE, Bin(DIV, Cast(kDouble, Attr(A_I)), LitD(3.0)), Bin(MUL, Attr(A_DISC), LitD(100.0))));

  // ------------------------------------------------------------------ the same over NULL-able attributes (relation tn)
  // comparisons with a NULL operand are false and NOT complements them (LiteralComparators-inl.hpp:168-223,
  // NegationPredicate.cpp:75-94); arithmetic over a NULL is NULL (ArithmeticBinaryOperators.hpp:178-186)
  g_nullable_mode = true;
  ScalarCase(out, "int_plus_literal", "i + 1", Bin(ADD, Attr(A_I), LitI(1)));
  ScalarCase(out, "int_times_long", "i * l", Bin(MUL, Attr(A_I), Attr(A_L)));
  ScalarCase(out, "q1_disc_price", "d * (1 - disc)", Bin(MUL, Attr(A_D), Bin(SUB, LitI(1), Attr(A_DISC))));
  ScalarCase(out, "q1_charge", "d * (1 - disc) * (1 + tax)",
             Bin(MUL, Bin(MUL, Attr(A_D), Bin(SUB, LitI(1), Attr(A_DISC))), Bin(ADD, LitI(1), Attr(A_TAX))));
  ScalarCase(out, "negate_double", "-d", Neg(Attr(A_D)));
  ScalarCase(out, "long_divide_int", "l / i2", Bin(DIV, Attr(A_L), Attr(A_I2)));
  ScalarCase(out, "cast_int_to_double", "CAST(i AS DOUBLE) * disc", Bin(MUL, Cast(kDouble, Attr(A_I)), Attr(A_DISC)));
  ScalarCase(out, "nested_mixed", "i + l * 2 - 3", Bin(SUB, Bin(ADD, Attr(A_I), Bin(MUL, Attr(A_L), LitI(2))), LitI(3)));
  ScalarCase(out, "shared_nullable", "SHARED#2(i + l) * (SHARED#2(i + l) - 1)",
             Bin(MUL, Shared(2, Bin(ADD, Attr(A_I), Attr(A_L))), Bin(SUB, Shared(2, Bin(ADD, Attr(A_I), Attr(A_L))), LitI(1))));
  ScalarCase(out, "bare_attribute_int", "i", Attr(A_I));
  ScalarCase(out, "bare_attribute_char", "c4", Attr(A_C4));
  ScalarCase(out, "bare_attribute_date", "dt", Attr(A_DT));
  ScalarCase(out, "not_null_operands_only", "disc + tax", Bin(ADD, Attr(A_DISC), Attr(A_TAX)));
  PredicateCase(out, "int_lt", "i < 100", Cmp(LT, Attr(A_I), LitI(100)));
  PredicateCase(out, "not_int_lt", "NOT (i < 100)", Not(Cmp(LT, Attr(A_I), LitI(100))));
  PredicateCase(out, "double_ge_special", "d >= 0.0", Cmp(GE, Attr(A_D), LitD(0.0)));
  PredicateCase(out, "double_ne_special", "d <> 0.0", Cmp(NE, Attr(A_D), LitD(0.0)));
  PredicateCase(out, "char4_eq", "c4 = 'MAIL'", Cmp(EQ, Attr(A_C4), LitC("MAIL")));
  PredicateCase(out, "char4_ne", "c4 <> 'MAIL'", Cmp(NE, Attr(A_C4), LitC("MAIL")));
  PredicateCase(out, "not_char4_eq", "NOT (c4 = 'MAIL')", Not(Cmp(EQ, Attr(A_C4), LitC("MAIL"))));
  PredicateCase(out, "char4_lt_shorter_literal", "c4 < 'RA'", Cmp(LT, Attr(A_C4), LitC("RA")));
  PredicateCase(out, "date_le", "dt <= DATE '1995-06-17'", Cmp(LE, Attr(A_DT), LitDate(1995, 6, 17)));
  PredicateCase(out, "attr_vs_attr_both_nullable", "i < l", Cmp(LT, Attr(A_I), Attr(A_L)));
  PredicateCase(out, "attr_vs_attr_one_nullable", "i = i2", Cmp(EQ, Attr(A_I), Attr(A_I2)));
  PredicateCase(out, "literal_on_the_left", "5 < i", Cmp(LT, LitI(5), Attr(A_I)));
  PredicateCase(out, "expr_vs_expr", "i + 1 < l * 2", Cmp(LT, Bin(ADD, Attr(A_I), LitI(1)), Bin(MUL, Attr(A_L), LitI(2))));
  PredicateCase(out, "or_of_nullables", "i < 0 OR d > 1000.0", Or({Cmp(LT, Attr(A_I), LitI(0)), Cmp(GT, Attr(A_D), LitD(1000.0))}));
  PredicateCase(out, "not_and_of_nullables", "NOT (i < 0 AND l > 0)", Not(And({Cmp(LT, Attr(A_I), LitI(0)), Cmp(GT, Attr(A_L), LitI(0))})));
  PredicateCase(out, "not_or_of_nullables", "NOT (i < 0 OR l > 0)", Not(Or({Cmp(LT, Attr(A_I), LitI(0)), Cmp(GT, Attr(A_L), LitI(0))})));
  PredicateCase(out, "mixed_nullable_and_not", "disc <= 0.05 AND i > 0", And({Cmp(LE, Attr(A_DISC), LitD(0.05)), Cmp(GT, Attr(A_I), LitI(0))}));
  PredicateCase(out, "not_not", "NOT (NOT (d < 50000.0))", Not(Not(Cmp(LT, Attr(A_D), LitD(50000.0)))));
  PredicateCase(out, "q6_where", "dt >= '1994-01-01' AND dt < '1995-01-01' AND disc >= 0.05 AND disc <= 0.07 AND i < 240",
                And({Cmp(GE, Attr(A_DT), LitDate(1994, 1, 1)), Cmp(LT, Attr(A_DT), LitDate(1995, 1, 1)),
                     Cmp(GE, Attr(A_DISC), LitD(0.05)), Cmp(LE, Attr(A_DISC), LitD(0.07)), Cmp(LT, Attr(A_I), LitI(240))}));
  g_nullable_mode = false;

  // ------------------------------------------------------------------ more type mixes (checked against the oracle on the
  // CPU only: added after the device runs of profiles/r4a-r4d; names start with "extra_")
  ScalarCase(out, "extra_float_times_long", "f * l", Bin(MUL, Attr(A_F), Attr(A_L)));
  ScalarCase(out, "extra_float_plus_long_literal", "f + 100 (LONG)", Bin(ADD, Attr(A_F), LitL(100)));
  ScalarCase(out, "extra_long_minus_float", "l2 - f", Bin(SUB, Attr(A_L2), Attr(A_F)));
  ScalarCase(out, "extra_int_div_double", "i / disc", Bin(DIV, Attr(A_I), Attr(A_DISC)));
  ScalarCase(out, "extra_double_mod_free_chain", "(d - tax) * (disc + 1) / (tax + 1)",
             Bin(DIV, Bin(MUL, Bin(SUB, Attr(A_D), Attr(A_TAX)), Bin(ADD, Attr(A_DISC), LitI(1))), Bin(ADD, Attr(A_TAX), LitI(1))));
  ScalarCase(out, "extra_cast_double_to_float", "CAST(disc AS FLOAT) * f", Bin(MUL, Cast(kFloat, Attr(A_DISC)), Attr(A_F)));
  ScalarCase(out, "extra_cast_long_to_double", "CAST(l AS DOUBLE) / 3.0", Bin(DIV, Cast(kDouble, Attr(A_L)), LitD(3.0)));
  ScalarCase(out, "extra_negate_float", "-f", Neg(Attr(A_F)));
  PredicateCase(out, "extra_float_vs_long_literal", "f < 3 (LONG)", Cmp(LT, Attr(A_F), LitL(3)));
  PredicateCase(out, "extra_long_vs_double_literal", "l < 1e9 + 0.5", Cmp(LT, Attr(A_L), LitD(1e9 + 0.5)));
  PredicateCase(out, "extra_int_vs_long_literal_out_of_int_range", "i < 10000000000", Cmp(LT, Attr(A_I), LitL(10000000000ll)));
  PredicateCase(out, "extra_int_vs_float_literal", "i >= 2.5 (FLOAT)", Cmp(GE, Attr(A_I), LitF(2.5f)));
  PredicateCase(out, "extra_long_vs_float_attr", "l2 > f", Cmp(GT, Attr(A_L2), Attr(A_F)));
  PredicateCase(out, "extra_double_vs_int_attr", "d <= i2", Cmp(LE, Attr(A_D), Attr(A_I2)));
  PredicateCase(out, "extra_date_vs_date_attr_expr", "dt < DATE '1995-01-01' OR dt >= DATE '1998-01-01'",
                Or({Cmp(LT, Attr(A_DT), LitDate(1995, 1, 1)), Cmp(GE, Attr(A_DT), LitDate(1998, 1, 1))}));
  PredicateCase(out, "extra_char1_range", "c1 >= 'N'", Cmp(GE, Attr(A_C1), LitC("N")));
  PredicateCase(out, "extra_char10_le_prefix", "c10 <= 'FURN'", Cmp(LE, Attr(A_C10), LitC("FURN")));

  std::fprintf(out, "\n ],\n \"aggregates\": [");
  AllAggregates(out);
  g_nullable_mode = true;
  AllAggregates(out);
  g_nullable_mode = false;

  std::fprintf(out, "\n ],\n \"lip_filters\": [");
  for (int pass = 0; pass < 2; ++pass) {
    g_nullable_mode = pass == 1;
    LipCase(out, "exact_int_probe_same_attr", true, -1000, 1000, 0, false, IsR(), A_I, A_I);
    LipCase(out, "exact_int_probe_other_attr", true, -1000, 1000, 0, false, IsR(), A_I, A_I2);
    LipCase(out, "exact_int_anti", true, -1000, 1000, 0, true, IsR(), A_I, A_I);
    LipCase(out, "exact_int_narrow_range", true, -500, 500, 0, false, InRange(A_I, -500, 500), A_I, A_I);
    LipCase(out, "exact_int_narrow_range_anti", true, -500, 500, 0, true, InRange(A_I, -500, 500), A_I, A_I);
    LipCase(out, "exact_long", true, -300, 300, 0, false, IsR(), A_L2, A_L2);
    LipCase(out, "exact_long_probe_out_of_range", true, -300, 300, 0, false, IsR(), A_L2, A_L);
    LipCase(out, "exact_long_probe_out_of_range_anti", true, -300, 300, 0, true, IsR(), A_L2, A_L);
    LipCase(out, "identity_hash_int_negative_values", false, 0, 0, 257, false, IsR(), A_I, A_I);
    LipCase(out, "identity_hash_int_small_cardinality", false, 0, 0, 64, false, IsR(), A_I2, A_I);
    LipCase(out, "identity_hash_long", false, 0, 0, 1000, false, IsR(), A_L, A_L);
    LipCase(out, "identity_hash_long_probe_other", false, 0, 0, 1021, false, InRange(A_I, -100, 900), A_L2, A_L);
  }
  g_nullable_mode = false;

  // ------------------------------------------------------------------ HashPartitionSchemeHeader::getPartitionId
  // (catalog/PartitionSchemeHeader.hpp:200-214) for the tuples' i (INT, negative values included), l and l2 (LONG): the
  // partition PartitionAwareInsertDestination sends a tuple to (storage/InsertDestination.cpp:598-640).
  std::fprintf(out, "\n ],\n \"hash_partitions\": [");
  bool first_part = true;
  for (const int attr : {static_cast<int>(A_I), static_cast<int>(A_L), static_cast<int>(A_L2)})
    for (const std::size_t n_parts : {2u, 4u, 7u, 8u, 13u}) {
      HashPartitionSchemeHeader header(n_parts, PartitionSchemeHeader::PartitionAttributeIds(1, attr));
      const std::size_t w = g_cols[attr].type->maximumByteLength();
      std::fprintf(out, "%s\n  {\"attr\": %d, \"n_parts\": %zu, \"partition_of_row\": [", first_part ? "" : ",", attr, n_parts);
      for (int r = 0; r < kRows; ++r) {
        const TypedValue v = g_cols[attr].type->makeValue(g_cols[attr].bytes.data() + r * w, w);
        std::fprintf(out, "%s%zu", r ? "," : "", static_cast<std::size_t>(header.getPartitionId({v})));
      }
      std::fprintf(out, "]}");
      first_part = false;
    }

  // ------------------------------------------------------------------ CompressionDictionary limit codes
  std::fprintf(out, "\n ],\n \"dictionaries\": [");
  DictionaryCase(out, A_I2, {{"-49", LitI(-49)}, {"-48", LitI(-48)}, {"0", LitI(0)}, {"7", LitI(7)}, {"48", LitI(48)}, {"49", LitI(49)},
                             {"6.5", LitD(6.5)}, {"7.0", LitD(7.0)}, {"-100 (LONG)", LitL(-100)}, {"7 (LONG)", LitL(7)}});
  DictionaryCase(out, A_L2, {{"-301", LitL(-301)}, {"-300", LitL(-300)}, {"12", LitL(12)}, {"300", LitL(300)}, {"301", LitL(301)}, {"12 (INT)", LitI(12)},
                             {"12.5", LitD(12.5)}});
  DictionaryCase(out, A_DISC, {{"-0.01", LitD(-0.01)}, {"0.0", LitD(0.0)}, {"0.05", LitD(0.05)}, {"0.055", LitD(0.055)}, {"0.1", LitD(0.1)}, {"0.11", LitD(0.11)},
                               {"0 (INT)", LitI(0)}, {"1 (INT)", LitI(1)}});
  DictionaryCase(out, A_DT, {{"1991-12-31", LitDate(1991, 12, 31)}, {"1995-06-17", LitDate(1995, 6, 17)}, {"1998-12-28", LitDate(1998, 12, 28)},
                             {"1999-01-01", LitDate(1999, 1, 1)}});
  DictionaryCase(out, A_C4, {{"AIR", LitC("AIR")}, {"MAIL", LitC("MAIL")}, {"MA", LitC("MA")}, {"TRUC", LitC("TRUC")}, {"TRUCKS", LitC("TRUCKS")},
                             {"ZZZZ", LitC("ZZZZ")}, {"A", LitC("A")}});
  DictionaryCase(out, A_C10, {{"BUILDING", LitC("BUILDING")}, {"AUTOMOBILE", LitC("AUTOMOBILE")}, {"AUTOMOBILES", LitC("AUTOMOBILES")}, {"B", LitC("B")},
                              {"HOUSEHOLD", LitC("HOUSEHOLD")}});

  std::fprintf(out, "\n ]}\n");
  if (out != stdout) std::fclose(out);
  return 0;
}
