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
// Design: an accumulator machine.  Each thread owns kRows rows of the current
// tile; one decoded instruction is applied to all of them (so decode cost is
// amortised and the kRows independent chains give ILP).  Column operands come
// from the shared-memory tile that the TMA unit filled, so operand fetch is an
// LDS with a *dynamic* column index -- something registers cannot do.  The
// reference materialises one ColumnVector per expression node; here no
// intermediate ever leaves registers.
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

// DateLit {int32 year; u8 month; u8 day; pad} -> order-preserving int64 key.
__device__ __forceinline__ uint64_t date_key(uint64_t raw) {
  const int64_t year = static_cast<int32_t>(raw & 0xffffffffu);
  const uint64_t md = ((raw >> 32) & 0xff) << 8 | ((raw >> 40) & 0xff);
  return static_cast<uint64_t>(year * 65536 + static_cast<int64_t>(md));
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
__device__ __forceinline__ bool char_cmp(uint8_t c, const char *v, const char *lit, uint32_t n) {
  int res = 0;
  for (uint32_t i = 0; i < n; ++i) {
    const unsigned char a = static_cast<unsigned char>(v[i]);
    const unsigned char b = static_cast<unsigned char>(lit[i]);
    if (a != b) { res = a < b ? -1 : 1; break; }
    if (a == 0) break;
  }
  return cmp_t<int>(c, res, 0);
}

// Default (no-op) sink hooks; concrete sinks override what they consume.
struct SinkBase {
  __device__ __forceinline__ void emit(uint32_t, uint8_t, const uint64_t (&)[kRows]) {}
  __device__ __forceinline__ void emit_raw(uint32_t, const char *, uint32_t) {}
  __device__ __forceinline__ void emit_raw_build(uint32_t, uint32_t) {}
  __device__ __forceinline__ uint64_t build_leaf(uint32_t, uint8_t, int) { return 0; }
};

// Per-thread VM state that survives between the predicate and emit sections.
struct VmRegs {
  uint64_t tmp[kMaxTmp][kRows];
};

// Row i of the tile that thread `tid` owns in its r-th lane: r*kBlock + tid.
__device__ __forceinline__ uint32_t tile_row(int r, int tid) { return r * kBlock + tid; }

/*
 * Run code[pc,end) for this thread's kRows rows.
 *   bits[r]  predicate bit stack (bit 0 = top)
 *   Sink     provides emit(j, type, acc), emit_raw(j, src, width, r) and
 *            build_leaf(col, ltype, r) (join build-side operand).
 */
template <class Sink>
__device__ __forceinline__ void vm_run(const Program &P, uint32_t pc, uint32_t end,
                                       const ScanDesc &S, const char *stage, int tid,
                                       VmRegs &regs, uint32_t (&bits)[kRows], Sink &sink) {
  uint64_t acc[kRows];
#pragma unroll
  for (int r = 0; r < kRows; ++r) acc[r] = 0;

  for (; pc < end; ++pc) {
    const Instr in = P.code[pc];
    uint64_t leaf[kRows];
    const bool wants_leaf = (in.op <= OP_MOD) || in.op == OP_CMP;
    if (wants_leaf) {
      switch (in.leaf) {
        case LEAF_COL: {
          const char *base = stage + S.cols[in.arg].smem_off;
          const uint32_t w = native_width(in.ltype);
#pragma unroll
          for (int r = 0; r < kRows; ++r)
            leaf[r] = vcvt(load_native(base + tile_row(r, tid) * w, in.ltype), in.ltype, in.type);
          break;
        }
        case LEAF_LIT: {
          const uint64_t v = P.lits[in.arg];
#pragma unroll
          for (int r = 0; r < kRows; ++r) leaf[r] = v;
          break;
        }
        case LEAF_TMP: {
          if (in.arg == 0) {
#pragma unroll
            for (int r = 0; r < kRows; ++r) leaf[r] = vcvt(regs.tmp[0][r], in.ltype, in.type);
          } else {
#pragma unroll
            for (int r = 0; r < kRows; ++r) leaf[r] = vcvt(regs.tmp[kMaxTmp - 1][r], in.ltype, in.type);
          }
          break;
        }
        default: {  // LEAF_BUILD
#pragma unroll
          for (int r = 0; r < kRows; ++r)
            leaf[r] = vcvt(sink.build_leaf(in.arg, in.ltype, r), in.ltype, in.type);
          break;
        }
      }
    }
    switch (in.op) {
      case OP_LOAD:
#pragma unroll
        for (int r = 0; r < kRows; ++r) acc[r] = leaf[r];
        break;
      case OP_ADD: case OP_SUB: case OP_MUL: case OP_DIV: case OP_MOD:
        if (in.flags & 1) {
#pragma unroll
          for (int r = 0; r < kRows; ++r) acc[r] = valu(in.op, in.type, leaf[r], acc[r]);
        } else {
#pragma unroll
          for (int r = 0; r < kRows; ++r) acc[r] = valu(in.op, in.type, acc[r], leaf[r]);
        }
        break;
      case OP_NEG:
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
          switch (in.type) {
            case V_F64: acc[r] = d2u(-u2d(acc[r])); break;
            case V_F32: acc[r] = f2u(-u2f(acc[r])); break;
            case V_I64: acc[r] = static_cast<uint64_t>(-static_cast<int64_t>(acc[r])); break;
            default: acc[r] = static_cast<uint64_t>(static_cast<int64_t>(-static_cast<int32_t>(acc[r]))); break;
          }
        }
        break;
      case OP_CVT:
#pragma unroll
        for (int r = 0; r < kRows; ++r) acc[r] = vcvt(acc[r], in.type, in.aux);
        break;
      case OP_ST_TMP:
        if (in.arg == 0) {
#pragma unroll
          for (int r = 0; r < kRows; ++r) regs.tmp[0][r] = acc[r];
        } else {
#pragma unroll
          for (int r = 0; r < kRows; ++r) regs.tmp[kMaxTmp - 1][r] = acc[r];
        }
        break;
      case OP_CMP:
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
          const bool b = (in.flags & 1) ? vcmp(in.aux, in.type, leaf[r], acc[r])
                                        : vcmp(in.aux, in.type, acc[r], leaf[r]);
          bits[r] = (bits[r] << 1) | (b ? 1u : 0u);
        }
        break;
      case OP_CMP_CHAR: {
        const char *base = stage + S.cols[in.arg].smem_off;
        const uint32_t w = S.cols[in.arg].width;
        const char *lit = P.str_pool + in.ltype;    // ltype doubles as pool offset
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
          const bool b = char_cmp(in.aux, base + tile_row(r, tid) * w, lit, w);
          bits[r] = (bits[r] << 1) | (b ? 1u : 0u);
        }
        break;
      }
      case OP_AND:
#pragma unroll
        for (int r = 0; r < kRows; ++r) bits[r] = (bits[r] >> 1) & (bits[r] | ~1u);
        break;
      case OP_OR:
#pragma unroll
        for (int r = 0; r < kRows; ++r) bits[r] = (bits[r] >> 1) | (bits[r] & 1u);
        break;
      case OP_NOT:
#pragma unroll
        for (int r = 0; r < kRows; ++r) bits[r] ^= 1u;
        break;
      case OP_PUSH_TRUE:
#pragma unroll
        for (int r = 0; r < kRows; ++r) bits[r] = (bits[r] << 1) | 1u;
        break;
      case OP_PUSH_FALSE:
#pragma unroll
        for (int r = 0; r < kRows; ++r) bits[r] = bits[r] << 1;
        break;
      case OP_LIP: {
        // Probe only rows still alive below the new stack top would need the
        // conjunction structure; LIP probes are always AND-ed right after, so
        // rows whose current top bit is 0 skip the (random) memory access.
        const LipDesc &f = S.lip[in.arg];
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
          bool b = false;
          if ((in.flags & 2) == 0 || (bits[r] & 1u)) {
            const int64_t v = static_cast<int64_t>(acc[r]);
            b = lip_contains(f, v);
          }
          bits[r] = (bits[r] << 1) | (b ? 1u : 0u);
        }
        break;
      }
      case OP_EMIT:
        sink.emit(in.arg, in.type, acc);
        break;
      case OP_EMIT_RAW:
        sink.emit_raw(in.arg, stage + S.cols[in.flags].smem_off, S.cols[in.flags].width);
        break;
      case OP_EMIT_RAW_BUILD:
        sink.emit_raw_build(in.arg, in.flags);
        break;
    }
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

__device__ __forceinline__ void issue_tile(const ScanDesc &S, const ScanRt &rt, uint32_t tile, char *stage,
                                           uint64_t *bar) {
  const uint64_t row0 = S.first_row + static_cast<uint64_t>(tile) * kTileRows;
  uint64_t rows = rt.row_end - row0;
  if (rows > kTileRows) rows = kTileRows;
  uint32_t total = 0;
  for (uint32_t c = 0; c < S.n_cols; ++c)
    total += (static_cast<uint32_t>(rows) * S.cols[c].width + 15u) & ~15u;
  mbar_expect_tx(bar, total);
  for (uint32_t c = 0; c < S.n_cols; ++c) {
    const uint32_t w = S.cols[c].width;
    const uint32_t bytes = (static_cast<uint32_t>(rows) * w + 15u) & ~15u;
    bulk_g2s(stage + S.cols[c].smem_off, S.cols[c].ptr + row0 * w, bytes, bar);
  }
}

/*
 * Persistent tile loop: CTA b owns tiles b, b+grid, b+2*grid, ...  (static
 * assignment => every CTA's partial result is reproducible run to run).
 * Thread 0 is the TMA producer; all threads consume.  body(tile, stage, rt)
 * may use CTA-wide barriers only if every thread reaches them.
 */
template <class Body>
__device__ __forceinline__ void scan_tiles(const ScanDesc &S, char *smem, Body &&body) {
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
  char *stages = smem + kBarBytes;
  const int tid = threadIdx.x;
  const ScanRt rt = scan_extent(S);
  if (tid == 0) {
    for (uint32_t s = 0; s < S.n_stages; ++s) mbar_init(&bars[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0) {
    uint32_t tile = blockIdx.x;
    for (uint32_t s = 0; s < S.n_stages && tile < rt.n_tiles; ++s, tile += gridDim.x)
      issue_tile(S, rt, tile, stages + s * S.stage_bytes, &bars[s]);
  }
  uint32_t s = 0, parity = 0;
  const uint32_t ahead = S.n_stages * gridDim.x;
  for (uint32_t tile = blockIdx.x; tile < rt.n_tiles; tile += gridDim.x) {
    char *stage = stages + s * S.stage_bytes;
    mbar_wait(&bars[s], parity);
    body(tile, stage, rt);
    __syncthreads();
    if (tid == 0) {
      const uint32_t next = tile + ahead;
      if (next < rt.n_tiles) issue_tile(S, rt, next, stage, &bars[s]);
    }
    if (++s == S.n_stages) { s = 0; parity ^= 1; }
  }
}

// Validity of this thread's rows in `tile` (row range may start/end mid-tile).
__device__ __forceinline__ void tile_valid(const ScanDesc &S, const ScanRt &rt, uint32_t tile, int tid,
                                           bool (&valid)[kRows]) {
  const uint64_t row0 = S.first_row + static_cast<uint64_t>(tile) * kTileRows;
#pragma unroll
  for (int r = 0; r < kRows; ++r) {
    const uint64_t row = row0 + tile_row(r, tid);
    valid[r] = row >= S.row_begin && row < rt.row_end;
  }
}

}  // namespace qs
