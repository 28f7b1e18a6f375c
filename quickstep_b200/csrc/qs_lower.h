// Host-side lowering of qs_node expression trees (the flattened
// serialization::Predicate / serialization::Scalar of
// expressions/Expressions.proto:29-137) into VM programs.
//
// Typing rules restate the reference:
//   * binary arithmetic result = TypeFactory::GetUnifyingType
//     (types/TypeFactory.cpp:159-180): INT<LONG, INT<FLOAT<DOUBLE, LONG+FLOAT->DOUBLE
//   * operands are converted with static_cast, then the C++ operator is applied
//     (ArithmeticBinaryOperators.hpp:51-159), literals are converted once here
//   * comparisons use the same promotion (LiteralComparators.hpp:36-72),
//     DATE compares lexicographically (types/DatetimeLit.hpp:65-93)
#pragma once

#include <map>
#include <string>
#include <vector>

#include "qs_common.cuh"
#include "qs_ops.cuh"

struct qsgpu_relation;

namespace qs {

struct Lowering {
  const qs_expr_set *ex = nullptr;
  const qsgpu_relation *rel = nullptr;        // scanned (probe) relation
  const qsgpu_relation *build_rel = nullptr;  // join build relation (attribute nodes with b == 2)
  Program P;
  std::vector<int> slot_of_attr;              // scanned attr -> staged column slot
  std::vector<uint32_t> staged_attrs;         // slot -> attr
  // How a staged attribute is used (matters for dictionary-coded attributes only): USE_CODE comparisons on
  // codes, USE_VALUE scalar leaves (dictionary lookup in registers), USE_RAW native bytes needed in the tile.
  enum : uint8_t { USE_CODE = 1, USE_VALUE = 2, USE_RAW = 4 };
  std::vector<uint8_t> staged_use;            // slot -> OR of the uses
  std::vector<int> bslot_of_attr;             // build attr -> JoinDesc::build_cols slot
  std::vector<uint32_t> build_attrs;
  uint32_t n_code = 0, n_lits = 0, n_str = 0;
  bool tmp_busy[kMaxTmp] = {false, false};
  struct Shared { int tmp; uint8_t type; };
  std::map<int, Shared> shared;
  std::string err;
  int status = QSGPU_OK;

  Lowering(const qs_expr_set *e, const qsgpu_relation *r, const qsgpu_relation *b = nullptr);

  bool fail(int st, const std::string &m) { if (status == QSGPU_OK) { status = st; err = m; } return false; }
  bool ok() const { return status == QSGPU_OK; }

  // NULL-able inputs: the scanned relation's per-row NULL mask travels as one more staged column (pseudo attribute
  // kNullMaskAttr, 8 bytes per row), staged only by programs that read a NULL-able attribute.
  static constexpr uint32_t kNullMaskAttr = 0xffffffffu;
  int null_slot = -1;
  int stage_null_mask();
  uint64_t null_bits(int i);                  // NULL-mask bits of the NULL-able scanned attributes scalar i reads
  uint64_t build_null_bits(int i);            // ... of the NULL-able BUILD-side attributes it reads (joins)
  void push_notnull_build(uint64_t bits, bool and_it);
  void lower_emit_null_build(uint32_t out_col, uint64_t bits);
  void push_notnull(uint64_t bits, bool and_it);          // push(no attribute of `bits` is NULL) [and AND it in]
  void lower_null_select(uint64_t bits, uint64_t identity);   // acc = NULL ? identity : acc
  void lower_emit_null(uint32_t out_col, uint64_t bits);

  int stage_attr(uint32_t attr, uint8_t use = USE_RAW);   // staged slot of a scanned attribute
  bool lower_code_compare(const qs_node *attr, const qs_node *lit, uint8_t cmp);   // coded attribute <cmp> literal
  int build_attr(uint32_t attr);
  void push(Instr in);
  int add_lit(uint64_t v);

  const qs_node *node(int i);
  uint8_t scalar_vtype(int i);                // compute type of a scalar node (DATE -> V_I64)
  bool is_leaf(int i);
  bool leaf_ref(int i, uint8_t want, Instr *in);   // fill leaf/ltype/arg of `in`
  uint8_t lower_scalar(int i);                // value left in acc; returns its VType
  void lower_cast_acc(uint8_t from, uint8_t to);
  void lower_pred(int i);
  void lower_comparison(const qs_node *n, const qs_node *l, const qs_node *r);   // one comparison, operands not NULL
  void lower_lip_probe(uint32_t lip_index, uint32_t attr, bool have_pred);
  void mark_pred_end() { P.n_pred = n_code; P.n_mid = n_code; }
  void mark_mid_end() { P.n_mid = n_code; }
  void finish() { P.n_total = n_code; }
};

uint8_t vtype_of(uint16_t qs_type);           // native VType (QS_DATE -> V_DATE)
uint8_t unify(uint8_t a, uint8_t b);
int dict_compare(uint16_t qs_type, uint32_t width, const char *a, const char *b);   // -1 / 0 / 1, 2 = unordered
// Codes of a sorted dictionary that satisfy `attribute <cmp> literal`: [lo, lo + span), or the complement (negate).
int dict_code_range(uint16_t attr_type, uint32_t width, const char *dict, uint32_t n_entries, uint8_t cmp,
                    const qs_node *lit, const char *str_pool, uint32_t str_pool_bytes, uint64_t *lo, uint64_t *span,
                    bool *negate, std::string *err);

}  // namespace qs
