// The predicate / scalar VM and the TMA tile pipeline shared by every scan
// kernel (K1-K7).
//
// Replaces, on the device, the reference's column-at-a-time evaluation:
//   Predicate::getAllMatches        expressions/predicate/ComparisonPredicate.cpp:115
//   ConjunctionPredicate            expressions/predicate/ConjunctionPredicate.cpp:109
//   Scalar::getAllValues            storage/StorageBlock.cpp:363-388
//   arithmetic functors             types/operations/binary_operations/ArithmeticBinaryOperators.hpp:51-159
//   comparison functors             types/operations/comparisons/LiteralComparators.hpp:36-72
//
// Design: an accumulator machine whose instruction stream is a compile-time
// constant.  The host lowers the expression trees of a work order into a
// linear program (lower.cu); the query compiler (qs_jit.cu) prints that
// program as the constexpr tables of a struct Q and instantiates the kernel
// template for it with NVRTC.  vm_run<Q, PC, END> below is the interpreter
// loop unrolled by template recursion: every `if constexpr` on Q::code(PC)
// folds away, so what reaches SASS is the straight-line typed arithmetic a
// hand-written kernel for that query would contain (the reference gets the
// same effect on the CPU from its template-instantiated functors).  Each
// thread owns kRows rows of the current tile and applies each instruction to
// all of them (kRows independent chains of ILP).  Column operands come from
// the shared-memory tile the TMA unit filled.  The reference materialises one
// ColumnVector per expression node; here no intermediate ever leaves registers.
//
// Arithmetic is IEEE, evaluated in the reference's operand order, compiled
// with --fmad=false so a*b+c is never contracted: per-row values are
// bit-identical to the CPU's.
#pragma once

#include "qs_common.cuh"

namespace qs {

__device__ __forceinline__ uint64_t d2u(double d) { return static_cast<uint64_t>(__double_as_longlong(d)); }
__device__ __forceinline__ double u2d(uint64_t u) { return __longlong_as_double(static_cast<long long>(u)); }
__device__ __forceinline__ uint64_t f2u(float f) { return static_cast<uint64_t>(__float_as_uint(f)); }
__device__ __forceinline__ float u2f(uint64_t u) { return __uint_as_float(static_cast<uint32_t>(u)); }

// DateLit {int32 year; u8 month; u8 day; pad} -> order-preserving int64 key: year in the high word (signed),
// month<<8 | day in the low word.  One PRMT: the year is already its own 32-bit register, and a 64-bit signed
// compare against a literal key is an ISETP pair (the previous year*65536+month*256+day form cost 9
// instructions per row in the Q1/Q6 predicates, r01c SASS).  lower.cu builds literal keys the same way.
__device__ __forceinline__ uint64_t date_key(uint64_t raw) {
  const uint32_t year = static_cast<uint32_t>(raw);
  const uint32_t md = __byte_perm(static_cast<uint32_t>(raw >> 32), 0u, 0x4401);   // byte0 = day, byte1 = month
  return (static_cast<uint64_t>(year) << 32) | md;
}

// Load one native value (sign-extending ints) from a staged / global column.
__device__ __forceinline__ uint64_t load_native(const char *p, uint8_t ltype) {
  switch (ltype) {
    case V_I32: return static_cast<uint64_t>(static_cast<int64_t>(*reinterpret_cast<const int32_t *>(p)));
    case V_F32: return static_cast<uint64_t>(*reinterpret_cast<const uint32_t *>(p));
    case V_DATE: return date_key(*reinterpret_cast<const uint64_t *>(p));
    default: return *reinterpret_cast<const uint64_t *>(p);   // I64, F64
  }
}
__device__ __forceinline__ uint32_t native_width(uint8_t ltype) {
  return (ltype == V_I32 || ltype == V_F32) ? 4u : 8u;
}

// Dictionary code of a row (CompressedTupleStorageSubBlock codes are 1, 2 or 4 bytes,
// storage/CompressedTupleStorageSubBlock.hpp:93-120).
template <uint32_t CW>
__device__ __forceinline__ uint32_t load_code_w(const char *p) {
  if constexpr (CW == 1) return *reinterpret_cast<const uint8_t *>(p);
  else if constexpr (CW == 2) return *reinterpret_cast<const uint16_t *>(p);
  else return *reinterpret_cast<const uint32_t *>(p);
}
// Row `row` of a coded column tile -> byte offset of its value inside the dictionary.  1/2-byte codes index a
// dictionary buffer that is readable for every code value; 4-byte codes are clamped (rows past the end of a
// ragged tile hold whatever the shared memory held before).
template <uint32_t CW, uint32_t W>
__device__ __forceinline__ uint32_t dict_offset(const char *codes, uint32_t row, uint32_t dict_entries) {
  uint32_t c = load_code_w<CW>(codes + row * CW);
  if constexpr (CW == 4) c = c < dict_entries ? c : dict_entries - 1;
  return c * W;
}

// The kernel's dynamic shared memory (every extern __shared__ array starts at the same address): dictionaries
// of 1-byte-coded attributes are copied to Q::col_doff(c) by scan_tiles before the first tile.
extern __shared__ __align__(128) char qs_dyn_smem[];
template <class Q, int C>
__device__ __forceinline__ const char *dict_of(const ScanDesc &S) {
  if constexpr (Q::col_dsmem(C) != 0) return qs_dyn_smem + Q::col_doff(C);
  else return S.cols[C].dict;
}

// Value conversion with C++ static_cast semantics (types/*Type.cpp coerceValue).
__device__ __forceinline__ uint64_t vcvt(uint64_t raw, uint8_t from, uint8_t to) {
  if (from == V_DATE) from = V_I64;
  if (from == to) return raw;
  switch (from * 4 + to) {
    case V_I32 * 4 + V_I64: return static_cast<uint64_t>(static_cast<int64_t>(static_cast<int32_t>(raw)));
    case V_I32 * 4 + V_F32: return f2u(static_cast<float>(static_cast<int32_t>(raw)));
    case V_I32 * 4 + V_F64: return d2u(static_cast<double>(static_cast<int32_t>(raw)));
    case V_I64 * 4 + V_I32: return static_cast<uint64_t>(static_cast<int64_t>(static_cast<int32_t>(raw)));
    case V_I64 * 4 + V_F32: return f2u(static_cast<float>(static_cast<int64_t>(raw)));
    case V_I64 * 4 + V_F64: return d2u(static_cast<double>(static_cast<int64_t>(raw)));
    case V_F32 * 4 + V_I32: return static_cast<uint64_t>(static_cast<int64_t>(static_cast<int32_t>(u2f(raw))));
    case V_F32 * 4 + V_I64: return static_cast<uint64_t>(static_cast<int64_t>(u2f(raw)));
    case V_F32 * 4 + V_F64: return d2u(static_cast<double>(u2f(raw)));
    case V_F64 * 4 + V_I32: return static_cast<uint64_t>(static_cast<int64_t>(static_cast<int32_t>(u2d(raw))));
    case V_F64 * 4 + V_I64: return static_cast<uint64_t>(static_cast<int64_t>(u2d(raw)));
    case V_F64 * 4 + V_F32: return f2u(static_cast<float>(u2d(raw)));
  }
  return raw;
}

template <typename T>
__device__ __forceinline__ T arith(uint8_t op, T a, T b) {
  switch (op) {
    case OP_ADD: return a + b;
    case OP_SUB: return a - b;
    case OP_MUL: return a * b;
    case OP_DIV: return a / b;
  }
  return a;
}
__device__ __forceinline__ int64_t imod(int64_t a, int64_t b) { return b == 0 ? 0 : a % b; }

__device__ __forceinline__ uint64_t valu(uint8_t op, uint8_t type, uint64_t a, uint64_t b) {
  switch (type) {
    case V_F64:
      return d2u(op == OP_MOD ? fmod(u2d(a), u2d(b)) : arith<double>(op, u2d(a), u2d(b)));
    case V_F32:
      return f2u(op == OP_MOD ? fmodf(u2f(a), u2f(b)) : arith<float>(op, u2f(a), u2f(b)));
    case V_I64: {
      const int64_t x = static_cast<int64_t>(a), y = static_cast<int64_t>(b);
      if (op == OP_DIV) return static_cast<uint64_t>(y == 0 ? 0 : x / y);
      if (op == OP_MOD) return static_cast<uint64_t>(imod(x, y));
      return static_cast<uint64_t>(arith<int64_t>(op, x, y));
    }
    default: {
      const int32_t x = static_cast<int32_t>(a), y = static_cast<int32_t>(b);
      int32_t r;
      if (op == OP_DIV) r = (y == 0 ? 0 : x / y);
      else if (op == OP_MOD) r = static_cast<int32_t>(imod(x, y));
      else r = static_cast<int32_t>(arith<uint32_t>(op, static_cast<uint32_t>(x), static_cast<uint32_t>(y)));
      return static_cast<uint64_t>(static_cast<int64_t>(r));
    }
  }
}

template <typename T>
__device__ __forceinline__ bool cmp_t(uint8_t c, T a, T b) {
  switch (c) {
    case QS_EQ: return a == b;
    case QS_NE: return a != b;
    case QS_LT: return a < b;
    case QS_LE: return a <= b;
    case QS_GT: return a > b;
    default: return a >= b;
  }
}
__device__ __forceinline__ bool vcmp(uint8_t c, uint8_t type, uint64_t a, uint64_t b) {
  switch (type) {
    case V_F64: return cmp_t<double>(c, u2d(a), u2d(b));
    case V_F32: return cmp_t<float>(c, u2f(a), u2f(b));
    case V_I32: return cmp_t<int32_t>(c, static_cast<int32_t>(a), static_cast<int32_t>(b));
    default: return cmp_t<int64_t>(c, static_cast<int64_t>(a), static_cast<int64_t>(b));
  }
}

// strncmp(col, lit, n) <cmp> 0 for fixed-width NUL-padded CHAR(n)
// (types/operations/comparisons/AsciiStringComparators.hpp:218-251).
// lit_longer: the literal continues past the attribute's N bytes, so a value that fills all N bytes and equals the
// literal's first N is less than the literal (the reference compares the full strings: 'abc' < 'abcdef').
template <uint32_t N>
__device__ __forceinline__ bool char_cmp(uint8_t c, const char *v, const char *lit, bool lit_longer = false) {
  int res = lit_longer ? -1 : 0;
#pragma unroll
  for (uint32_t i = 0; i < N; ++i) {
    const unsigned char a = static_cast<unsigned char>(v[i]);
    const unsigned char b = static_cast<unsigned char>(lit[i]);
    if (a != b) { res = a < b ? -1 : 1; break; }
    if (a == 0) { res = 0; break; }
  }
  return cmp_t<int>(c, res, 0);
}

// ------------------------------------------------------ compile-time loops
template <int I> struct IC { static constexpr int value = I; };
template <int B, int E, class F>
__device__ __forceinline__ void static_for(F &&f) {
  if constexpr (B < E) {
    f(IC<B>{});
    static_for<B + 1, E>(f);
  }
}
#define QS_IDX(x) (decltype(x)::value)

// Default (no-op) sink hooks; concrete sinks override what they consume.
struct SinkBase {
  template <int J, int TYPE> __device__ __forceinline__ void emit(const uint64_t (&)[kRows]) {}
  template <int J, int COL, int W> __device__ __forceinline__ void emit_raw(const char *) {}
  template <int J, int COL, int W> __device__ __forceinline__ void emit_raw_build() {}
  template <int COL, int LTYPE, int W> __device__ __forceinline__ uint64_t build_leaf(int) { return 0; }
  template <int J> __device__ __forceinline__ void emit_null(const bool (&)[kRows]) {}
  __device__ __forceinline__ uint64_t build_null_mask(int) { return 0; }
};

// Per-thread VM state that survives between the predicate and emit sections.
struct VmRegs {
  uint64_t tmp[kMaxTmp][kRows];
  uint64_t acc[kRows];
  // per-thread LIP probe statistics over all tiles (dead registers in kernels that probe no filter)
  uint32_t lip_cnt[kMaxLip] = {0, 0, 0, 0}, lip_miss[kMaxLip] = {0, 0, 0, 0};
};

// Row i of the tile that thread `tid` owns in its r-th lane: r*kBlock + tid.
__device__ __forceinline__ uint32_t tile_row(int r, int tid) { return r * kBlock + tid; }

/*
 * Run Q::code[PC,END) for this thread's kRows rows.
 *   bits[r]  bit 0 = the predicate value on top of the stack when the section starts / ends
 *   Sink     provides emit<J,TYPE>(acc), emit_raw<J,COL,W>(col tile),
 *            emit_raw_build<J,COL,W>() and build_leaf<COL,LTYPE,W>(r).
 *
 * The predicate stack lives in `pst[depth][row]` with a COMPILE-TIME stack pointer SP (the program is a
 * compile-time constant, so the depth before every instruction is too): a comparison writes one predicate
 * register, AND / OR / NOT are predicate-logic instructions.  (The first version kept the stack as a shifted
 * bit word per row; ncu showed SEL + SHF + LOP3 per comparison and per connective, 18 of Q6's 50 instructions
 * per row once the scan on dictionary codes had made that kernel issue-bound.)
 */
__host__ __device__ constexpr bool op_pushes(uint8_t op) {
  return op == OP_CMP || op == OP_CMP_CHAR || op == OP_CMP_CODE || op == OP_PUSH_TRUE || op == OP_PUSH_FALSE || op == OP_LIP ||
         op == OP_NOTNULL || op == OP_NOTNULL_BUILD;
}
template <class Q>
__host__ __device__ constexpr int pred_depth(int pc, int end) {
  int sp = 1, mx = 1;
  for (int i = pc; i < end; ++i) {
    const uint8_t op = Q::code(i).op;
    if (op_pushes(op)) ++sp;
    else if (op == OP_AND || op == OP_OR) --sp;
    if (sp > mx) mx = sp;
  }
  return mx;
}

// LIP probe statistics are kept only by scans that probe SEVERAL filters: the order of a single filter cannot be
// adapted, and the per-row counters cost the lineitem select of Q3 (one filter) 20 % when they were unconditional
// (70 -> 80+ registers: one resident CTA per SM fewer).
template <class Q>
__host__ __device__ constexpr int lip_ops() {
  int n = 0;
  for (int i = 0; i < Q::n_total; ++i) n += Q::code(i).op == OP_LIP ? 1 : 0;
  return n;
}

// Does the program record NULL-ness of projected columns (a scan of a relation with NULL-able attributes)?
template <class Q>
__host__ __device__ constexpr bool emits_null() {
  for (int i = 0; i < Q::n_total; ++i) if (Q::code(i).op == OP_EMIT_NULL || Q::code(i).op == OP_EMIT_NULL_BUILD) return true;
  return false;
}

template <class Q, int PC, int END, int SP, int D, class Sink>
__device__ __forceinline__ void vm_step(const Lits &L, const ScanDesc &S, const char *__restrict__ stage, int tid,
                                        VmRegs &regs, bool (&pst)[D][kRows], uint32_t (&bits)[kRows], Sink &sink) {
  if constexpr (PC >= END) {
#pragma unroll
    for (int r = 0; r < kRows; ++r) bits[r] = pst[SP - 1][r] ? 1u : 0u;
  } else {
    constexpr Instr in = Q::code(PC);
    constexpr int NSP = op_pushes(in.op) ? SP + 1 : (in.op == OP_AND || in.op == OP_OR) ? SP - 1 : SP;
    static_assert(SP >= 1 && NSP >= 1 && NSP <= D, "predicate stack out of bounds");
    uint64_t (&acc)[kRows] = regs.acc;
    uint64_t leaf[kRows];
    constexpr bool wants_leaf = (in.op <= OP_MOD) || in.op == OP_CMP;
    if constexpr (wants_leaf) {
      if constexpr (in.leaf == LEAF_COL) {
        constexpr uint32_t w = (in.ltype == V_I32 || in.ltype == V_F32) ? 4u : 8u;
        if constexpr (Q::col_cw(in.arg) != 0 && !Q::col_expand(in.arg)) {
          // dictionary-coded attribute: value = dict[code], the code tile is all that came from HBM
          const char *codes = stage + Q::col_coff(in.arg);
          const char *dict = dict_of<Q, in.arg>(S);
          const uint32_t n = S.cols[in.arg].dict_entries;
#pragma unroll
          for (int r = 0; r < kRows; ++r)
            leaf[r] = vcvt(load_native(dict + dict_offset<Q::col_cw(in.arg), w>(codes, tile_row(r, tid), n), in.ltype),
                           in.ltype, in.type);
        } else {
          const char *base = stage + Q::col_off(in.arg);
#pragma unroll
          for (int r = 0; r < kRows; ++r)
            leaf[r] = vcvt(load_native(base + tile_row(r, tid) * w, in.ltype), in.ltype, in.type);
        }
      } else if constexpr (in.leaf == LEAF_LIT) {
        const uint64_t v = L.lits[in.arg];
#pragma unroll
        for (int r = 0; r < kRows; ++r) leaf[r] = v;
      } else if constexpr (in.leaf == LEAF_TMP) {
#pragma unroll
        for (int r = 0; r < kRows; ++r) leaf[r] = vcvt(regs.tmp[in.arg == 0 ? 0 : kMaxTmp - 1][r], in.ltype, in.type);
      } else {  // LEAF_BUILD
#pragma unroll
        for (int r = 0; r < kRows; ++r)
          leaf[r] = vcvt(sink.template build_leaf<in.arg, in.ltype, Q::build_w(in.arg)>(r), in.ltype, in.type);
      }
    }
    if constexpr (in.op == OP_LOAD) {
#pragma unroll
      for (int r = 0; r < kRows; ++r) acc[r] = leaf[r];
    } else if constexpr (in.op >= OP_ADD && in.op <= OP_MOD) {
#pragma unroll
      for (int r = 0; r < kRows; ++r)
        acc[r] = (in.flags & 1) ? valu(in.op, in.type, leaf[r], acc[r]) : valu(in.op, in.type, acc[r], leaf[r]);
    } else if constexpr (in.op == OP_NEG) {
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        if constexpr (in.type == V_F64) acc[r] = d2u(-u2d(acc[r]));
        else if constexpr (in.type == V_F32) acc[r] = f2u(-u2f(acc[r]));
        else if constexpr (in.type == V_I64) acc[r] = static_cast<uint64_t>(-static_cast<int64_t>(acc[r]));
        else acc[r] = static_cast<uint64_t>(static_cast<int64_t>(-static_cast<int32_t>(acc[r])));
      }
    } else if constexpr (in.op == OP_CVT) {
#pragma unroll
      for (int r = 0; r < kRows; ++r) acc[r] = vcvt(acc[r], in.type, in.aux);
    } else if constexpr (in.op == OP_ST_TMP) {
#pragma unroll
      for (int r = 0; r < kRows; ++r) regs.tmp[in.arg == 0 ? 0 : kMaxTmp - 1][r] = acc[r];
    } else if constexpr (in.op == OP_CMP) {
#pragma unroll
      for (int r = 0; r < kRows; ++r)
        pst[SP][r] = (in.flags & 1) ? vcmp(in.aux, in.type, leaf[r], acc[r]) : vcmp(in.aux, in.type, acc[r], leaf[r]);
    } else if constexpr (in.op == OP_CMP_CHAR) {
      const char *base = stage + Q::col_off(in.arg);
      constexpr uint32_t w = Q::col_w(in.arg);
      const char *lit = L.str_pool + in.ltype;    // ltype doubles as pool offset
#pragma unroll
      for (int r = 0; r < kRows; ++r) pst[SP][r] = char_cmp<w>(in.aux, base + tile_row(r, tid) * w, lit, (in.flags & 4) != 0);
    } else if constexpr (in.op == OP_CMP_CODE) {
      // attribute <cmp> literal on a dictionary-coded attribute: the host turned the literal into the range
      // of codes that satisfy it (the dictionary is sorted), CompressedTupleStorageSubBlock::getMatchesForPredicate
      // (storage/CompressedTupleStorageSubBlock.cpp:160-251); one unsigned range test per row, no value load
      const char *codes = stage + Q::col_coff(in.arg);
      const uint32_t lo = static_cast<uint32_t>(L.lits[in.aux]);
      const uint32_t span = static_cast<uint32_t>(L.lits[in.aux + 1]);
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        const uint32_t c = load_code_w<Q::col_cw(in.arg)>(codes + tile_row(r, tid) * Q::col_cw(in.arg));
        pst[SP][r] = ((c - lo) < span) != ((in.flags & 1) != 0);
      }
    } else if constexpr (in.op == OP_AND) {
#pragma unroll
      for (int r = 0; r < kRows; ++r) pst[SP - 2][r] = pst[SP - 2][r] && pst[SP - 1][r];
    } else if constexpr (in.op == OP_OR) {
#pragma unroll
      for (int r = 0; r < kRows; ++r) pst[SP - 2][r] = pst[SP - 2][r] || pst[SP - 1][r];
    } else if constexpr (in.op == OP_NOT) {
#pragma unroll
      for (int r = 0; r < kRows; ++r) pst[SP - 1][r] = !pst[SP - 1][r];
    } else if constexpr (in.op == OP_PUSH_TRUE) {
#pragma unroll
      for (int r = 0; r < kRows; ++r) pst[SP][r] = true;
    } else if constexpr (in.op == OP_PUSH_FALSE) {
#pragma unroll
      for (int r = 0; r < kRows; ++r) pst[SP][r] = false;
    } else if constexpr (in.op == OP_LIP) {
      // LIP probes are always AND-ed right after (flags&2), so rows whose
      // current top value is false skip the (random) memory access.
      const LipDesc &f = S.lip[in.arg];
      // the filter's bit words: the CTA's shared-memory copy when plan_scan made one (small filters), else global
      const uint64_t *words;
      if constexpr (Q::lip_soff(in.arg) != 0) words = reinterpret_cast<const uint64_t *>(qs_dyn_smem + Q::lip_soff(in.arg));
      else words = f.words;
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        bool b = false;
        // pst[0] starts as the row's validity (rows of a ragged first / last tile outside the scanned range) and only
        // ever gets AND-ed: invalid rows are never probed
        if (pst[0][r] && ((in.flags & 2) == 0 || pst[SP - 1][r])) {
          b = lip_contains<Q::lip_kind(in.arg), Q::lip_anti(in.arg)>(f, words, static_cast<int64_t>(acc[r]));
          if constexpr (lip_ops<Q>() >= 2) {
            ++regs.lip_cnt[in.arg];
            regs.lip_miss[in.arg] += b ? 0u : 1u;
          }
        }
        pst[SP][r] = b;
      }
    } else if constexpr (in.op == OP_NOTNULL) {
      const char *base = stage + Q::col_off(in.arg);
      const uint64_t m = L.lits[in.aux];
#pragma unroll
      for (int r = 0; r < kRows; ++r)
        pst[SP][r] = (*reinterpret_cast<const uint64_t *>(base + tile_row(r, tid) * 8u) & m) == 0ull;
    } else if constexpr (in.op == OP_NULLSEL) {
      const char *base = stage + Q::col_off(in.arg);
      const uint64_t m = L.lits[in.aux], ident = L.lits[in.aux + 1];
#pragma unroll
      for (int r = 0; r < kRows; ++r)
        acc[r] = (*reinterpret_cast<const uint64_t *>(base + tile_row(r, tid) * 8u) & m) == 0ull ? acc[r] : ident;
    } else if constexpr (in.op == OP_EMIT_NULL) {
      const char *base = stage + Q::col_off(in.flags);
      const uint64_t m = L.lits[in.aux];
      bool isnull[kRows];
#pragma unroll
      for (int r = 0; r < kRows; ++r)
        isnull[r] = (*reinterpret_cast<const uint64_t *>(base + tile_row(r, tid) * 8u) & m) != 0ull;
      sink.template emit_null<in.arg>(isnull);
    } else if constexpr (in.op == OP_NOTNULL_BUILD) {
      const uint64_t m = L.lits[in.aux];
#pragma unroll
      for (int r = 0; r < kRows; ++r) pst[SP][r] = (sink.build_null_mask(r) & m) == 0ull;
    } else if constexpr (in.op == OP_EMIT_NULL_BUILD) {
      const uint64_t m = L.lits[in.aux];
      bool isnull[kRows];
#pragma unroll
      for (int r = 0; r < kRows; ++r) isnull[r] = (sink.build_null_mask(r) & m) != 0ull;
      sink.template emit_null<in.arg>(isnull);
    } else if constexpr (in.op == OP_EMIT) {
      sink.template emit<in.arg, in.type>(acc);
    } else if constexpr (in.op == OP_EMIT_RAW) {
      sink.template emit_raw<in.arg, in.flags, Q::col_w(in.flags)>(stage + Q::col_off(in.flags));
    } else if constexpr (in.op == OP_EMIT_RAW_BUILD) {
      sink.template emit_raw_build<in.arg, in.flags, Q::build_w(in.flags)>();
    }
    vm_step<Q, PC + 1, END, NSP, D>(L, S, stage, tid, regs, pst, bits, sink);
  }
}

template <class Q, int PC, int END, class Sink>
__device__ __forceinline__ void vm_run(const Lits &L, const ScanDesc &S, const char *__restrict__ stage, int tid,
                                       VmRegs &regs, uint32_t (&bits)[kRows], Sink &sink) {
  if constexpr (PC < END) {
    constexpr int D = pred_depth<Q>(PC, END);
    bool pst[D][kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) pst[0][r] = (bits[r] & 1u) != 0;
    vm_step<Q, PC, END, 1, D>(L, S, stage, tid, regs, pst, bits, sink);
  }
}

// --------------------------------------------------------- tile pipeline
// Shared memory layout: [kMaxStages mbarriers][pad to 128][stage 0][stage 1]...

// Run-time extent of a scan: the row count of a temporary relation may only be
// known on the device (it is the output counter of the previous operator), so
// operators chain without a host round trip.
struct ScanRt {
  uint64_t row_end;
  uint32_t n_tiles;
};

__device__ __forceinline__ ScanRt scan_extent(const ScanDesc &S) {
  ScanRt rt;
  rt.row_end = S.row_end;
  rt.n_tiles = S.n_tiles;
  if (S.d_row_end != nullptr) {
    const uint64_t re = *S.d_row_end;
    if (re < rt.row_end) rt.row_end = re;
    rt.n_tiles = rt.row_end > S.first_row
                     ? static_cast<uint32_t>((rt.row_end - S.first_row + kTileRows - 1) / kTileRows)
                     : 0u;
  }
  return rt;
}

template <class Q>
__device__ __forceinline__ void issue_tile(const ScanDesc &S, const ScanRt &rt, uint32_t tile, char *stage,
                                           uint64_t *bar) {
  const uint64_t row0 = S.first_row + static_cast<uint64_t>(tile) * kTileRows;
  uint64_t rows64 = rt.row_end - row0;
  const uint32_t rows = rows64 > kTileRows ? kTileRows : static_cast<uint32_t>(rows64);
  // a coded attribute travels as codes (col_cw bytes per row) into its code tile
  uint32_t total = 0;
  static_for<0, Q::n_cols>([&](auto c) {
    constexpr uint32_t w = Q::col_cw(QS_IDX(c)) ? Q::col_cw(QS_IDX(c)) : Q::col_w(QS_IDX(c));
    total += (rows * w + 15u) & ~15u;
  });
  mbar_expect_tx(bar, total);
  static_for<0, Q::n_cols>([&](auto c) {
    constexpr bool coded = Q::col_cw(QS_IDX(c)) != 0;
    constexpr uint32_t w = coded ? Q::col_cw(QS_IDX(c)) : Q::col_w(QS_IDX(c));
    constexpr uint32_t off = coded ? Q::col_coff(QS_IDX(c)) : Q::col_off(QS_IDX(c));
    bulk_g2s(stage + off, S.cols[QS_IDX(c)].ptr + row0 * w, (rows * w + 15u) & ~15u, bar);
  });
}

// Coded attributes whose native bytes are needed in place (group-by keys, pass-through projections, join and
// LIP keys, CHAR comparisons against non-literals): the CTA decodes the code tile into the attribute's native
// tile once per tile.  Attributes that are only compared with literals or read as scalar leaves never get here.
template <class Q>
__device__ __forceinline__ constexpr bool any_expand() {
  bool any = false;
  for (int c = 0; c < static_cast<int>(Q::n_cols); ++c) any = any || (Q::col_cw(c) != 0 && Q::col_expand(c));
  return any;
}
template <class Q>
__device__ __forceinline__ void expand_tile(const ScanDesc &S, char *stage, int tid) {
  static_for<0, Q::n_cols>([&](auto cc) {
    constexpr int c = QS_IDX(cc);
    if constexpr (Q::col_cw(c) != 0 && Q::col_expand(c)) {
      constexpr uint32_t w = Q::col_w(c);
      const char *codes = stage + Q::col_coff(c);
      const char *dict = dict_of<Q, c>(S);
      const uint32_t n = S.cols[c].dict_entries;
      char *out = stage + Q::col_off(c);
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        const uint32_t row = tile_row(r, tid);
        const char *src = dict + dict_offset<Q::col_cw(c), w>(codes, row, n);
        if constexpr (w == 8) *reinterpret_cast<uint64_t *>(out + row * 8u) = *reinterpret_cast<const uint64_t *>(src);
        else if constexpr (w == 4) *reinterpret_cast<uint32_t *>(out + row * 4u) = *reinterpret_cast<const uint32_t *>(src);
        else {
#pragma unroll
          for (uint32_t b = 0; b < w; ++b) out[row * w + b] = src[b];
        }
      }
    }
  });
}

/*
 * Persistent tile loop: CTA b owns tiles b, b+grid, b+2*grid, ...  (static
 * assignment => every CTA's partial result is reproducible run to run).
 * Thread 0 is the TMA producer; all threads consume.  body(tile, stage, rt)
 * may use CTA-wide barriers only if every thread reaches them.
 */
template <class Q, class Body>
__device__ __forceinline__ void scan_tiles(const ScanDesc &S, char *smem, Body &&body) {
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
  char *stages = smem + kBarBytes;
  const int tid = threadIdx.x;
  const ScanRt rt = scan_extent(S);
  if (tid == 0) {
    for (uint32_t s = 0; s < Q::n_stages; ++s) mbar_init(&bars[s], 1);
    fence_barrier_init();
  }
  // shared-memory copies of the small dictionaries (256 entries each; the device buffer is that long)
  static_for<0, Q::n_cols>([&](auto cc) {
    constexpr int c = QS_IDX(cc);
    if constexpr (Q::col_cw(c) != 0 && Q::col_dsmem(c) != 0) {
      constexpr uint32_t n16 = 256u * Q::col_w(c) / 16u;
      const uint4 *src = reinterpret_cast<const uint4 *>(S.cols[c].dict);
      uint4 *dst = reinterpret_cast<uint4 *>(smem + Q::col_doff(c));
      for (uint32_t i = tid; i < n16; i += kBlock) dst[i] = src[i];
    }
  });
  // ... and of the small LIP filters this scan probes (LIPFilterAdaptiveProber's filters live in the probing
  // thread's cache in the reference; here the CTA's shared memory).  The filter is complete: its build ran earlier
  // on the same stream.
  static_for<0, kMaxLip>([&](auto ff) {
    constexpr int f = QS_IDX(ff);
    if constexpr (f < static_cast<int>(Q::n_lip)) {
      if constexpr (Q::lip_soff(f) != 0) {
        const uint64_t n = S.lip[f].n_words;
        const uint64_t *src = S.lip[f].words;
        uint64_t *dst = reinterpret_cast<uint64_t *>(smem + Q::lip_soff(f));
        for (uint64_t i = tid; i < n; i += kBlock) dst[i] = src[i];
      }
    }
  });
  __syncthreads();
  if (tid == 0) {
    uint32_t tile = blockIdx.x;
    for (uint32_t s = 0; s < Q::n_stages && tile < rt.n_tiles; ++s, tile += gridDim.x)
      issue_tile<Q>(S, rt, tile, stages + s * Q::stage_bytes, &bars[s]);
  }
  uint32_t s = 0, parity = 0;
  const uint32_t ahead = Q::n_stages * gridDim.x;
  for (uint32_t tile = blockIdx.x; tile < rt.n_tiles; tile += gridDim.x) {
    char *stage = stages + s * Q::stage_bytes;
    mbar_wait(&bars[s], parity);
    if constexpr (any_expand<Q>()) {
      expand_tile<Q>(S, stage, tid);
      __syncthreads();
    }
    body(tile, stage, rt);
    __syncthreads();
    if (tid == 0) {
      const uint32_t next = tile + ahead;
      if (next < rt.n_tiles) issue_tile<Q>(S, rt, next, stage, &bars[s]);
    }
    if (++s == Q::n_stages) { s = 0; parity ^= 1; }
  }
}

// End of a scan kernel: this CTA's probe / miss counts go to the filters' statistics words (one atomic pair per warp
// and filter per kernel).  The host reads them to order the filters of later work orders by miss rate
// (LIPFilterAdaptiveProber, utility/lip_filter/LIPFilterAdaptiveProber.hpp:89-232); results never depend on them.
template <class Q>
__device__ __forceinline__ void flush_lip_stats(const ScanDesc &S, const VmRegs &regs) {
  if constexpr (lip_ops<Q>() >= 2)
  static_for<0, kMaxLip>([&](auto ff) {
    constexpr int f = QS_IDX(ff);
    bool probed = false;
    for (int pc = 0; pc < Q::n_total; ++pc) probed = probed || (Q::code(pc).op == OP_LIP && Q::code(pc).arg == f);
    if (!probed) return;
    uint32_t c = regs.lip_cnt[f], m = regs.lip_miss[f];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      c += __shfl_xor_sync(0xffffffffu, c, off);
      m += __shfl_xor_sync(0xffffffffu, m, off);
    }
    if ((threadIdx.x & 31) == 0 && c != 0 && S.lip[f].stats != nullptr) {
      atomicAdd(&S.lip[f].stats[0], static_cast<unsigned long long>(c));
      if (m) atomicAdd(&S.lip[f].stats[1], static_cast<unsigned long long>(m));
    }
  });
}

// Validity of this thread's rows in `tile` (row range may start/end mid-tile).  Only the first and the
// last tile of a scan can hold rows outside [row_begin, row_end): every other tile takes the CTA-uniform
// fast path and pays nothing per row.
__device__ __forceinline__ void tile_valid(const ScanDesc &S, const ScanRt &rt, uint32_t tile, int tid,
                                           bool (&valid)[kRows]) {
  const uint64_t row0 = S.first_row + static_cast<uint64_t>(tile) * kTileRows;
  uint32_t mask = (1u << kRows) - 1u;
  if (row0 < S.row_begin || row0 + kTileRows > rt.row_end) {
    mask = 0;
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      const uint64_t row = row0 + tile_row(r, tid);
      mask |= (row >= S.row_begin && row < rt.row_end) ? (1u << r) : 0u;
    }
  }
#pragma unroll
  for (int r = 0; r < kRows; ++r) valid[r] = (mask >> r) & 1u;
}

}  // namespace qs
