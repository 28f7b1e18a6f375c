// Expression-tree -> VM program lowering (host code; see qs_lower.h).
#include <cstring>

#include "qs_host.h"
#include "qs_lower.h"

namespace qs {

uint8_t vtype_of(uint16_t t) {
  switch (t) {
    case QS_INT: return V_I32;
    case QS_LONG: return V_I64;
    case QS_FLOAT: return V_F32;
    case QS_DOUBLE: return V_F64;
    case QS_DATE: return V_DATE;
    default: return 0xff;
  }
}

// TypeFactory::GetUnifyingType for numeric pairs (types/TypeFactory.cpp:159-180).
uint8_t unify(uint8_t a, uint8_t b) {
  if (a == V_DATE) a = V_I64;
  if (b == V_DATE) b = V_I64;
  if (a == b) return a;
  if (a == V_F64 || b == V_F64) return V_F64;
  if ((a == V_I64 && b == V_F32) || (a == V_F32 && b == V_I64)) return V_F64;
  if (a == V_F32 || b == V_F32) return V_F32;   // INT with FLOAT
  return V_I64;                                  // INT with LONG
}

static uint64_t host_cvt(uint64_t raw, uint8_t from, uint8_t to) {
  if (from == V_DATE) from = V_I64;
  if (from == to) return raw;
  double d = 0;
  int64_t i = 0;
  bool is_fp = false;
  switch (from) {
    case V_I32: i = static_cast<int32_t>(raw); break;
    case V_I64: i = static_cast<int64_t>(raw); break;
    case V_F32: { uint32_t u = static_cast<uint32_t>(raw); float f; std::memcpy(&f, &u, 4); d = f; is_fp = true; break; }
    default: { std::memcpy(&d, &raw, 8); is_fp = true; break; }
  }
  switch (to) {
    case V_I32: return static_cast<uint64_t>(static_cast<int64_t>(is_fp ? static_cast<int32_t>(d) : static_cast<int32_t>(i)));
    case V_I64: return static_cast<uint64_t>(is_fp ? static_cast<int64_t>(d) : i);
    case V_F32: { float f = is_fp ? static_cast<float>(d) : static_cast<float>(i); uint32_t u; std::memcpy(&u, &f, 4); return u; }
    default: { double x = is_fp ? d : static_cast<double>(i); uint64_t u; std::memcpy(&u, &x, 8); return u; }
  }
}

Lowering::Lowering(const qs_expr_set *e, const qsgpu_relation *r, const qsgpu_relation *b)
    : ex(e), rel(r), build_rel(b) {
  std::memset(&P, 0, sizeof(P));
  if (r) slot_of_attr.assign(r->attrs.size(), -1);
  if (b) bslot_of_attr.assign(b->attrs.size(), -1);
}

const qs_node *Lowering::node(int i) {
  if (!ex || i < 0 || static_cast<uint32_t>(i) >= ex->n_nodes) {
    fail(QSGPU_ERR_INVALID, "expression node index out of range");
    return nullptr;
  }
  return &ex->nodes[i];
}

int Lowering::stage_attr(uint32_t attr, uint8_t use) {
  if (!rel || attr >= rel->attrs.size()) { fail(QSGPU_ERR_INVALID, "attribute id out of range"); return 0; }
  if (slot_of_attr[attr] >= 0) { staged_use[slot_of_attr[attr]] |= use; return slot_of_attr[attr]; }
  if (staged_attrs.size() >= static_cast<size_t>(kMaxCols)) {
    fail(QSGPU_ERR_UNSUPPORTED, "more than kMaxCols attributes referenced by one scan");
    return 0;
  }
  slot_of_attr[attr] = static_cast<int>(staged_attrs.size());
  staged_attrs.push_back(attr);
  staged_use.push_back(use);
  return slot_of_attr[attr];
}

int Lowering::stage_null_mask() {
  if (null_slot >= 0) return null_slot;
  if (!rel || !rel->d_nulls) { fail(QSGPU_ERR_INVALID, "relation without a NULL mask"); return 0; }
  if (staged_attrs.size() >= static_cast<size_t>(kMaxCols)) {
    fail(QSGPU_ERR_UNSUPPORTED, "more than kMaxCols attributes referenced by one scan");
    return 0;
  }
  null_slot = static_cast<int>(staged_attrs.size());
  staged_attrs.push_back(kNullMaskAttr);
  staged_use.push_back(USE_RAW);
  return null_slot;
}

// A scalar is NULL when any attribute it reads is (every operation of the path propagates NULL:
// ArithmeticBinaryOperators.hpp:178-186 applyToTypedValues returns a NULL of the result type).
uint64_t Lowering::null_bits(int i) {
  const qs_node *n = node(i);
  if (!n || !rel) return 0;
  switch (n->kind) {
    case QS_N_ATTRIBUTE:
      if (n->b == 2) return 0;            // build side: build_null_bits
      return static_cast<uint32_t>(n->a) < 64 ? (rel->nullable_mask & (1ull << n->a)) : 0ull;
    case QS_N_UNARY: case QS_N_SHARED: return null_bits(n->a);
    case QS_N_BINARY: return null_bits(n->a) | null_bits(n->b);
    default: return 0;
  }
}

uint64_t Lowering::build_null_bits(int i) {
  const qs_node *n = node(i);
  if (!n || !build_rel) return 0;
  switch (n->kind) {
    case QS_N_ATTRIBUTE:
      return (n->b == 2 && static_cast<uint32_t>(n->a) < 64) ? (build_rel->nullable_mask & (1ull << n->a)) : 0ull;
    case QS_N_UNARY: case QS_N_SHARED: return build_null_bits(n->a);
    case QS_N_BINARY: return build_null_bits(n->a) | build_null_bits(n->b);
    default: return 0;
  }
}

void Lowering::push_notnull_build(uint64_t bits, bool and_it) {
  Instr in{};
  in.op = OP_NOTNULL_BUILD;
  in.aux = static_cast<uint8_t>(add_lit(bits));
  push(in);
  if (and_it) { Instr a{}; a.op = OP_AND; push(a); }
}

void Lowering::lower_emit_null_build(uint32_t out_col, uint64_t bits) {
  Instr in{};
  in.op = OP_EMIT_NULL_BUILD;
  in.arg = static_cast<uint16_t>(out_col);
  in.aux = static_cast<uint8_t>(add_lit(bits));
  push(in);
}

void Lowering::push_notnull(uint64_t bits, bool and_it) {
  Instr in{};
  in.op = OP_NOTNULL;
  in.arg = static_cast<uint16_t>(stage_null_mask());
  in.aux = static_cast<uint8_t>(add_lit(bits));
  push(in);
  if (and_it) { Instr a{}; a.op = OP_AND; push(a); }
}

void Lowering::lower_null_select(uint64_t bits, uint64_t identity) {
  Instr in{};
  in.op = OP_NULLSEL;
  in.arg = static_cast<uint16_t>(stage_null_mask());
  in.aux = static_cast<uint8_t>(add_lit(bits));
  add_lit(identity);
  push(in);
}

void Lowering::lower_emit_null(uint32_t out_col, uint64_t bits) {
  Instr in{};
  in.op = OP_EMIT_NULL;
  in.arg = static_cast<uint16_t>(out_col);
  in.flags = static_cast<uint8_t>(stage_null_mask());
  in.aux = static_cast<uint8_t>(add_lit(bits));
  push(in);
}

int Lowering::build_attr(uint32_t attr) {
  if (!build_rel || attr >= build_rel->attrs.size()) {
    fail(QSGPU_ERR_INVALID, "build-side attribute without a build relation");
    return 0;
  }
  if (bslot_of_attr[attr] >= 0) return bslot_of_attr[attr];
  if (build_attrs.size() >= static_cast<size_t>(kMaxCols)) {
    fail(QSGPU_ERR_UNSUPPORTED, "too many build-side attributes");
    return 0;
  }
  bslot_of_attr[attr] = static_cast<int>(build_attrs.size());
  build_attrs.push_back(attr);
  return bslot_of_attr[attr];
}

void Lowering::push(Instr in) {
  if (n_code >= static_cast<uint32_t>(kMaxInstr)) { fail(QSGPU_ERR_UNSUPPORTED, "expression program too long"); return; }
  P.code[n_code++] = in;
}

int Lowering::add_lit(uint64_t v) {
  // no de-duplication by value: the instruction stream (= the compiled kernel's identity)
  // must not depend on literal values
  if (n_lits >= static_cast<uint32_t>(kMaxLits)) { fail(QSGPU_ERR_UNSUPPORTED, "too many literals"); return 0; }
  P.L.lits[n_lits] = v;
  return static_cast<int>(n_lits++);
}

uint8_t Lowering::scalar_vtype(int i) {
  const qs_node *n = node(i);
  if (!n) return V_I64;
  switch (n->kind) {
    case QS_N_LITERAL:
    case QS_N_ATTRIBUTE: {
      const uint8_t v = vtype_of(n->type);
      if (v == 0xff) { fail(QSGPU_ERR_UNSUPPORTED, "CHAR/VARCHAR value in arithmetic context"); return V_I64; }
      return v == V_DATE ? V_I64 : v;
    }
    case QS_N_UNARY:
      if (n->op == QS_CAST) { const uint8_t v = vtype_of(n->type); return v == V_DATE ? V_I64 : v; }
      return scalar_vtype(n->a);
    case QS_N_BINARY: return unify(scalar_vtype(n->a), scalar_vtype(n->b));
    case QS_N_SHARED: return scalar_vtype(n->a);
    default: fail(QSGPU_ERR_INVALID, "predicate node used as scalar"); return V_I64;
  }
}

bool Lowering::is_leaf(int i) {
  const qs_node *n = node(i);
  if (!n) return false;
  if (n->kind == QS_N_LITERAL || n->kind == QS_N_ATTRIBUTE) return true;
  if (n->kind == QS_N_SHARED) return shared.count(n->b) > 0;
  return false;
}

static uint64_t literal_raw(const qs_node *n) {
  switch (n->type) {
    case QS_INT: return static_cast<uint64_t>(static_cast<int64_t>(n->lit.i32));
    case QS_LONG: return static_cast<uint64_t>(n->lit.i64);
    case QS_FLOAT: { uint32_t u; std::memcpy(&u, &n->lit.f32, 4); return u; }
    case QS_DOUBLE: { uint64_t u; std::memcpy(&u, &n->lit.f64, 8); return u; }
    case QS_DATE:
      // same key as date_key() on the device: year in the high word, month<<8 | day in the low word
      return (static_cast<uint64_t>(static_cast<uint32_t>(n->lit.date.year)) << 32) |
             (static_cast<uint64_t>(n->lit.date.month) << 8) | n->lit.date.day;
    default: return 0;
  }
}

bool Lowering::leaf_ref(int i, uint8_t want, Instr *in) {
  const qs_node *n = node(i);
  if (!n) return false;
  if (n->kind == QS_N_LITERAL) {
    const uint8_t own = vtype_of(n->type);
    if (own == 0xff) return fail(QSGPU_ERR_UNSUPPORTED, "CHAR literal in arithmetic context");
    in->leaf = LEAF_LIT;
    in->ltype = want;
    in->arg = static_cast<uint16_t>(add_lit(host_cvt(literal_raw(n), own, want)));
    return true;
  }
  if (n->kind == QS_N_ATTRIBUTE) {
    const uint8_t own = vtype_of(n->type);
    if (own == 0xff) return fail(QSGPU_ERR_UNSUPPORTED, "CHAR attribute in arithmetic context");
    in->ltype = own;
    if (n->b == 2) { in->leaf = LEAF_BUILD; in->arg = static_cast<uint16_t>(build_attr(static_cast<uint32_t>(n->a))); }
    else { in->leaf = LEAF_COL; in->arg = static_cast<uint16_t>(stage_attr(static_cast<uint32_t>(n->a), USE_VALUE)); }
    return true;
  }
  if (n->kind == QS_N_SHARED) {
    auto it = shared.find(n->b);
    if (it == shared.end()) return fail(QSGPU_ERR_INVALID, "shared expression not materialised");
    in->leaf = LEAF_TMP;
    in->ltype = it->second.type;
    in->arg = static_cast<uint16_t>(it->second.tmp);
    return true;
  }
  return fail(QSGPU_ERR_INVALID, "not a leaf");
}

void Lowering::lower_cast_acc(uint8_t from, uint8_t to) {
  if (from == V_DATE) from = V_I64;
  if (from == to) return;
  Instr in{};
  in.op = OP_CVT; in.type = from; in.aux = to;
  push(in);
}

uint8_t Lowering::lower_scalar(int i) {
  const qs_node *n = node(i);
  if (!n || !ok()) return V_I64;
  switch (n->kind) {
    case QS_N_LITERAL:
    case QS_N_ATTRIBUTE: {
      const uint8_t t = scalar_vtype(i);
      Instr in{};
      in.op = OP_LOAD; in.type = t;
      leaf_ref(i, t, &in);
      push(in);
      return t;
    }
    case QS_N_SHARED: {
      auto it = shared.find(n->b);
      if (it != shared.end()) {
        Instr in{};
        in.op = OP_LOAD; in.type = it->second.type;
        leaf_ref(i, it->second.type, &in);
        push(in);
        return it->second.type;
      }
      const uint8_t t = lower_scalar(n->a);
      for (int k = 0; k < kMaxTmp; ++k) {
        if (!tmp_busy[k]) {
          tmp_busy[k] = true;            // held for the rest of the program
          Instr st{};
          st.op = OP_ST_TMP; st.type = t; st.arg = static_cast<uint16_t>(k);
          push(st);
          shared[n->b] = Shared{k, t};
          break;
        }
      }
      return t;                           // no free temp: recomputed at the next use
    }
    case QS_N_UNARY: {
      const uint8_t t = lower_scalar(n->a);
      if (n->op == QS_NEGATE) {
        Instr in{};
        in.op = OP_NEG; in.type = t;
        push(in);
        return t;
      }
      if (n->op == QS_CAST) {
        uint8_t to = vtype_of(n->type);
        if (to == 0xff) { fail(QSGPU_ERR_UNSUPPORTED, "CAST to a non-numeric type"); return t; }
        if (to == V_DATE) to = V_I64;
        lower_cast_acc(t, to);
        return to;
      }
      fail(QSGPU_ERR_UNSUPPORTED, "unary operation not lowered (DateExtract/Substring)");
      return t;
    }
    case QS_N_BINARY: {
      const uint8_t ta = scalar_vtype(n->a), tb = scalar_vtype(n->b);
      const uint8_t T = unify(ta, tb);
      if (n->op > QS_MOD) { fail(QSGPU_ERR_INVALID, "bad binary operation id"); return T; }
      Instr in{};
      in.op = static_cast<uint8_t>(OP_ADD + n->op);
      in.type = T;
      if (is_leaf(n->b)) {
        const uint8_t t = lower_scalar(n->a);
        lower_cast_acc(t, T);
        leaf_ref(n->b, T, &in);
      } else if (is_leaf(n->a)) {
        const uint8_t t = lower_scalar(n->b);
        lower_cast_acc(t, T);
        leaf_ref(n->a, T, &in);
        in.flags = 1;
      } else {
        // The temporary is chosen AFTER the right operand is lowered: a ScalarSharedExpression that is first used
        // inside it takes a temporary for the rest of the program, and must not end up sharing this one.
        const uint8_t t_b = lower_scalar(n->b);
        int k = -1;
        for (int q = 0; q < kMaxTmp; ++q) if (!tmp_busy[q]) { k = q; break; }
        if (k < 0) { fail(QSGPU_ERR_UNSUPPORTED, "expression too deep for the VM temporaries"); return T; }
        tmp_busy[k] = true;
        Instr st{};
        st.op = OP_ST_TMP; st.type = t_b; st.arg = static_cast<uint16_t>(k);
        push(st);
        const uint8_t t_a = lower_scalar(n->a);
        lower_cast_acc(t_a, T);
        in.leaf = LEAF_TMP; in.ltype = t_b; in.arg = static_cast<uint16_t>(k);
        tmp_busy[k] = false;
      }
      push(in);
      return T;
    }
    default:
      fail(QSGPU_ERR_INVALID, "predicate node used as scalar");
      return V_I64;
  }
}

static uint8_t flip_cmp(uint8_t c) {
  switch (c) {
    case QS_LT: return QS_GT;
    case QS_LE: return QS_GE;
    case QS_GT: return QS_LT;
    case QS_GE: return QS_LE;
    default: return c;
  }
}

// Three-way comparison of a dictionary entry with a literal, both in the comparison's compute type
// (LiteralComparators.hpp:36-72); 2 = unordered (NaN).
static int host_cmp3(uint8_t T, uint64_t a, uint64_t b) {
  switch (T) {
    case V_F64: { double x, y; std::memcpy(&x, &a, 8); std::memcpy(&y, &b, 8); return x < y ? -1 : x > y ? 1 : x == y ? 0 : 2; }
    case V_F32: { float x, y; uint32_t ua = static_cast<uint32_t>(a), ub = static_cast<uint32_t>(b); std::memcpy(&x, &ua, 4); std::memcpy(&y, &ub, 4); return x < y ? -1 : x > y ? 1 : x == y ? 0 : 2; }
    case V_I32: { const int32_t x = static_cast<int32_t>(a), y = static_cast<int32_t>(b); return x < y ? -1 : x > y ? 1 : 0; }
    default: { const int64_t x = static_cast<int64_t>(a), y = static_cast<int64_t>(b); return x < y ? -1 : x > y ? 1 : 0; }
  }
}

static uint64_t dict_raw(const char *p, uint16_t type) {
  switch (type) {
    case QS_INT: { int32_t v; std::memcpy(&v, p, 4); return static_cast<uint64_t>(static_cast<int64_t>(v)); }
    case QS_FLOAT: { uint32_t v; std::memcpy(&v, p, 4); return v; }
    case QS_DATE: {   // DateLit bytes -> the order-preserving key of date_key()
      int32_t year; std::memcpy(&year, p, 4);
      const uint8_t month = static_cast<uint8_t>(p[4]), day = static_cast<uint8_t>(p[5]);
      return (static_cast<uint64_t>(static_cast<uint32_t>(year)) << 32) | (static_cast<uint64_t>(month) << 8) | day;
    }
    default: { uint64_t v; std::memcpy(&v, p, 8); return v; }
  }
}

// Order of two dictionary entries of an attribute (the type's less-than, as the reference's
// CompressionDictionaryBuilder sorts them, compression/CompressionDictionaryBuilder.cpp:120-160).
int dict_compare(uint16_t type, uint32_t width, const char *a, const char *b) {
  if (type == QS_CHAR) { const int r = std::strncmp(a, b, width); return r < 0 ? -1 : r > 0 ? 1 : 0; }
  const uint8_t T = vtype_of(type) == V_DATE ? V_I64 : vtype_of(type);
  return host_cmp3(T, dict_raw(a, type), dict_raw(b, type));
}

/*
 * attribute <cmp> literal where the attribute is held as codes of a sorted relation-wide dictionary: count the
 * dictionary entries below / equal to / above the literal and compare codes against those bounds.  This is
 * what CompressedTupleStorageSubBlock::getMatchesForPredicate does per block with
 * CompressionDictionary::getLimitCodesForComparisonTyped (storage/CompressedTupleStorageSubBlock.cpp:160-251,
 * compression/CompressionDictionary.hpp:241-318); here once per work order.  The satisfying codes always form
 * one range [lo, lo+span) or its complement, so the kernel runs one unsigned range test per row; the bounds
 * travel as literals (the kernel does not depend on their values).
 */
int dict_code_range(uint16_t attr_type, uint32_t w, const char *dict, uint32_t n_entries, uint8_t cmp,
                    const qs_node *lit, const char *str_pool, uint32_t str_pool_bytes, uint64_t *lo_out,
                    uint64_t *span_out, bool *negate_out, std::string *err) {
  uint64_t n_lt = 0, n_eq = 0, n_gt = 0;
  if ((attr_type == QS_CHAR) != (lit->type == QS_CHAR)) { *err = "comparison of a coded attribute with a literal of another kind"; return QSGPU_ERR_UNSUPPORTED; }
  if (attr_type == QS_CHAR) {
    // strncmp over the attribute width, literal NUL-padded / cut like the native CHAR path
    if (lit->lit.pool_offset + lit->width > str_pool_bytes) { *err = "CHAR literal outside pool"; return QSGPU_ERR_INVALID; }
    std::string l(w, '\0');
    for (uint32_t b = 0; b < w && b < lit->width; ++b) l[b] = str_pool[lit->lit.pool_offset + b];
    // literal longer than the attribute: the reference compares the full strings ('abc' < 'abcdef'), so an entry that
    // fills all w bytes and equals the literal's first w bytes is BELOW the literal, never equal to it
    const bool lit_longer = lit->width > w && str_pool[lit->lit.pool_offset + w] != 0;
    for (uint32_t e = 0; e < n_entries; ++e) {
      const char *v = dict + static_cast<size_t>(e) * w;
      int r = std::strncmp(v, l.data(), w);
      if (r == 0 && lit_longer && std::memchr(v, 0, w) == nullptr) r = -1;
      (r < 0 ? n_lt : r > 0 ? n_gt : n_eq)++;
    }
  } else {
    const uint8_t own = vtype_of(attr_type), lown = vtype_of(lit->type);
    if (own == 0xff || lown == 0xff) { *err = "comparison of a coded attribute with a non-numeric literal"; return QSGPU_ERR_UNSUPPORTED; }
    const uint8_t T = unify(own, lown);
    const uint64_t lv = host_cvt(literal_raw(lit), lown, T);
    for (uint32_t e = 0; e < n_entries; ++e) {
      const uint64_t dv = host_cvt(dict_raw(dict + static_cast<size_t>(e) * w, attr_type), own, T);
      const int r = host_cmp3(T, dv, lv);
      if (r < 0) ++n_lt; else if (r == 0) ++n_eq; else if (r == 1) ++n_gt;
    }
  }
  // codes [0, n_lt) are below the literal, [n_lt, n_lt + n_eq) equal, the last n_gt above
  const uint64_t n = n_entries;
  uint64_t lo = 0, span = 0;
  bool negate = false;
  switch (cmp) {
    case QS_EQ: lo = n_lt; span = n_eq; break;
    case QS_NE: lo = n_lt; span = n_eq; negate = true; break;   // NaN literal: span 0, every row differs
    case QS_LT: lo = 0; span = n_lt; break;
    case QS_LE: lo = 0; span = n_lt + n_eq; break;
    case QS_GT: lo = n - n_gt; span = n_gt; break;
    case QS_GE: lo = n - n_gt - n_eq; span = n_gt + n_eq; break;
    default: *err = "not a comparison id"; return QSGPU_ERR_INVALID;
  }
  *lo_out = lo; *span_out = span; *negate_out = negate;
  return QSGPU_OK;
}

bool Lowering::lower_code_compare(const qs_node *attr, const qs_node *lit, uint8_t cmp) {
  const uint32_t a = static_cast<uint32_t>(attr->a);
  const qs_coded_attr &C = rel->coded[a];
  uint64_t lo = 0, span = 0;
  bool negate = false;
  std::string why;
  const int st = dict_code_range(attr->type, rel->attrs[a].width, C.h_dict.data(), C.n_entries, cmp, lit, ex->str_pool,
                                 ex->str_pool_bytes, &lo, &span, &negate, &why);
  if (st != QSGPU_OK) return fail(st, why);
  Instr in{};
  in.op = OP_CMP_CODE;
  in.arg = static_cast<uint16_t>(stage_attr(a, USE_CODE));
  in.flags = negate ? 1 : 0;
  in.aux = static_cast<uint8_t>(add_lit(lo));
  add_lit(span);
  push(in);
  return ok();
}

void Lowering::lower_pred(int i) {
  const qs_node *n = node(i);
  if (!n || !ok()) return;
  Instr in{};
  switch (n->kind) {
    case QS_N_TRUE: in.op = OP_PUSH_TRUE; push(in); return;
    case QS_N_FALSE: in.op = OP_PUSH_FALSE; push(in); return;
    case QS_N_NEGATION: lower_pred(n->a); in.op = OP_NOT; push(in); return;
    case QS_N_CONJUNCTION:
    case QS_N_DISJUNCTION:
      lower_pred(n->a);
      lower_pred(n->b);
      in.op = n->kind == QS_N_CONJUNCTION ? OP_AND : OP_OR;
      push(in);
      return;
    case QS_N_COMPARISON: {
      const qs_node *l = node(n->a), *r = node(n->b);
      if (!l || !r) return;
      if (n->op > QS_GE) { fail(QSGPU_ERR_UNSUPPORTED, "LIKE / regex comparisons are not lowered"); return; }
      // a comparison with a NULL operand is false (LiteralComparators-inl.hpp:168-223: `!(cv_nullable && value ==
      // nullptr) && compare(...)`); NOT then complements it like NegationPredicate::getAllMatches does (no three-
      // valued logic in the reference either).  The stored bytes of a NULL value are zero, so the comparison
      // itself is well defined and its answer is AND-ed with "no operand is NULL".
      lower_comparison(n, l, r);
      if (const uint64_t nb = null_bits(n->a) | null_bits(n->b)) push_notnull(nb, true);
      if (const uint64_t nbb = build_null_bits(n->a) | build_null_bits(n->b)) push_notnull_build(nbb, true);
      return;
    }
    default:
      fail(QSGPU_ERR_INVALID, "scalar node used as predicate");
  }
}

void Lowering::lower_comparison(const qs_node *n, const qs_node *l, const qs_node *r) {
  Instr in{};
  {
    // coded attribute vs literal (either order): evaluated on the codes
    const qs_node *attr = l->kind == QS_N_ATTRIBUTE ? l : r;
    const qs_node *lit = l->kind == QS_N_ATTRIBUTE ? r : l;
    if (attr->kind == QS_N_ATTRIBUTE && lit->kind == QS_N_LITERAL && attr->b != 2 && rel &&
        static_cast<uint32_t>(attr->a) < rel->attrs.size() && rel->code_width(static_cast<uint32_t>(attr->a)) != 0 &&
        ((attr->type == QS_CHAR) == (lit->type == QS_CHAR)) &&
        (attr->type != QS_CHAR || lit->lit.pool_offset + lit->width <= ex->str_pool_bytes)) {
      lower_code_compare(attr, lit, attr == l ? static_cast<uint8_t>(n->op) : flip_cmp(static_cast<uint8_t>(n->op)));
      return;
    }
  }
  if (l->type == QS_CHAR || r->type == QS_CHAR) {
    // attribute vs literal only; the literal is NUL-padded to the attribute width
    const qs_node *attr = l->kind == QS_N_ATTRIBUTE ? l : r;
    const qs_node *lit = l->kind == QS_N_ATTRIBUTE ? r : l;
    if (attr->kind != QS_N_ATTRIBUTE || lit->kind != QS_N_LITERAL || attr->type != QS_CHAR ||
        lit->type != QS_CHAR || attr->b == 2) {
      fail(QSGPU_ERR_UNSUPPORTED, "CHAR comparison other than attribute-vs-literal");
      return;
    }
    const uint32_t w = attr->width;
    if (lit->lit.pool_offset + lit->width > ex->str_pool_bytes) { fail(QSGPU_ERR_INVALID, "CHAR literal outside pool"); return; }
    // A literal longer than the attribute (col CHAR(3) vs 'abcdef'): the reference compares the full strings, so
    // a value equal to the literal's first w bytes is LESS than the literal; flag 4 tells the kernel
    const bool lit_longer = lit->width > w && ex->str_pool[lit->lit.pool_offset + w] != 0;
    if (n_str + w > static_cast<uint32_t>(kStrPool)) { fail(QSGPU_ERR_UNSUPPORTED, "string pool full"); return; }
    const uint32_t off = n_str;
    for (uint32_t b = 0; b < w; ++b)
      P.L.str_pool[off + b] = b < lit->width ? ex->str_pool[lit->lit.pool_offset + b] : 0;
    n_str += w;
    in.op = OP_CMP_CHAR;
    in.arg = static_cast<uint16_t>(stage_attr(static_cast<uint32_t>(attr->a)));
    in.ltype = static_cast<uint8_t>(off);
    in.aux = (attr == l) ? static_cast<uint8_t>(n->op) : flip_cmp(static_cast<uint8_t>(n->op));
    in.flags = lit_longer ? 4 : 0;
    push(in);
    return;
  }
  const uint8_t T = unify(scalar_vtype(n->a), scalar_vtype(n->b));
  in.op = OP_CMP; in.type = T; in.aux = static_cast<uint8_t>(n->op);
  if (is_leaf(n->b)) {
    lower_cast_acc(lower_scalar(n->a), T);
    leaf_ref(n->b, T, &in);
  } else if (is_leaf(n->a)) {
    lower_cast_acc(lower_scalar(n->b), T);
    leaf_ref(n->a, T, &in);
    in.flags = 1;
  } else {
    const uint8_t t_b = lower_scalar(n->b);      // first (see lower_scalar): nested shared expressions claim theirs
    int k = -1;
    for (int q = 0; q < kMaxTmp; ++q) if (!tmp_busy[q]) { k = q; break; }
    if (k < 0) { fail(QSGPU_ERR_UNSUPPORTED, "comparison too deep for the VM temporaries"); return; }
    tmp_busy[k] = true;
    Instr st{};
    st.op = OP_ST_TMP; st.type = t_b; st.arg = static_cast<uint16_t>(k);
    push(st);
    lower_cast_acc(lower_scalar(n->a), T);
    in.leaf = LEAF_TMP; in.ltype = t_b; in.arg = static_cast<uint16_t>(k);
    tmp_busy[k] = false;
  }
  push(in);
  return;
}

// LIPFilterAdaptiveProber: probe attribute `attr` against filter `lip_index`
// and AND the answer into the running predicate.
void Lowering::lower_lip_probe(uint32_t lip_index, uint32_t attr, bool have_pred) {
  if (!rel || attr >= rel->attrs.size()) { fail(QSGPU_ERR_INVALID, "LIP probe attribute out of range"); return; }
  const uint8_t lt = vtype_of(rel->attrs[attr].type);
  if (lt != V_I32 && lt != V_I64) { fail(QSGPU_ERR_UNSUPPORTED, "LIP filters take INT/LONG attributes"); return; }
  // a NULL key joins with nothing (HashTable::getAllFromValueAccessor skips NULL keys of a nullable key attribute,
  // storage/HashTable.hpp:1903), so the row is dropped here like a filter miss
  if (attr < 64 && ((rel->nullable_mask >> attr) & 1ull)) {
    push_notnull(1ull << attr, have_pred);
    have_pred = true;
  }
  Instr ld{};
  ld.op = OP_LOAD; ld.type = V_I64; ld.leaf = LEAF_COL; ld.ltype = lt;
  ld.arg = static_cast<uint16_t>(stage_attr(attr));
  push(ld);
  Instr pr{};
  pr.op = OP_LIP; pr.type = V_I64; pr.arg = static_cast<uint16_t>(lip_index);
  pr.flags = have_pred ? 2 : 0;
  push(pr);
  if (have_pred) { Instr a{}; a.op = OP_AND; push(a); }
}

}  // namespace qs
