// Shared definitions for the libqsgpu kernels (sm_100a only).
//
// Device-side program format, tile geometry and the PTX wrappers for the
// mbarrier + bulk-copy (TMA, SASS UBLKCP) pipeline every scan kernel uses.
#pragma once

// Under NVRTC (the query compiler, qs_jit.cu) the three includes below resolve
// to the minimal stand-ins the JIT embeds, so this header and everything that
// builds on it compile without a CUDA toolkit tree at run time.
#include <cuda_runtime.h>
#include <stdint.h>

#include "qsgpu_types.h"

namespace qs {

// ------------------------------------------------------------ tile geometry
constexpr int kBlock = 256;                      // threads per CTA
#ifndef QS_ROWS
#define QS_ROWS 4
#endif
constexpr int kRows = QS_ROWS;                   // rows per thread per tile
constexpr int kTileRows = kBlock * kRows;        // 1024 rows per tile
constexpr int kMaxCols = 12;                     // staged columns per scan
constexpr int kMaxStages = 8;                    // ring depth upper bound
constexpr int kMaxInstr = 112;
constexpr int kMaxLits = 32;
constexpr int kStrPool = 96;
constexpr int kMaxLip = 4;
constexpr int kMaxTmp = 2;
constexpr int kMaxAgg = 12;                      // state words per group besides the row count (values + non-NULL counts)
constexpr int kMaxKeyCols = 8;
constexpr int kMaxKeyWords = 4;                  // <= 32 byte composite keys
constexpr int kMaxOut = 12;                      // projected columns
constexpr int kCompactMaxGroups = 256;           // K2 per-CTA / global cap
constexpr int kCompactLocalSlots = 512;          // smem open-addressing slots
constexpr uint32_t kBarBytes = 128;              // mbarrier area at the head of dynamic smem
constexpr uint32_t kCompactSmemBytes = (32 + 4) * 4;      // per-(row, warp) counts (<= 32) + base / overflow words

// ------------------------------------------------------------ VM value types
// Compute types of the scalar VM.  Native column types map onto them at leaf
// load: INT->I32, LONG->I64, FLOAT->F32, DOUBLE->F64, DATE->DATEKEY (I64 key
// year<<16|month<<8|day, order-preserving for DateLit::operator<).
enum VType : uint8_t { V_I32 = 0, V_I64 = 1, V_F32 = 2, V_F64 = 3, V_DATE = 4 /*native only*/ };

enum Op : uint8_t {
  OP_LOAD = 0,    // acc = leaf
  OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_MOD,   // acc = acc op leaf  (flags&1: leaf op acc)
  OP_NEG,
  OP_CVT,         // acc: type -> aux
  OP_ST_TMP,      // tmp[arg] = acc
  OP_CMP,         // push(acc <aux> leaf)          (flags&1: leaf <aux> acc)
  OP_CMP_CHAR,    // push(strncmp(col[arg], pool+lit, width) <aux> 0); arg2 in lits
  OP_AND, OP_OR, OP_NOT, OP_PUSH_TRUE, OP_PUSH_FALSE,
  OP_LIP,         // push(lip[arg].contains(acc))  acc is I32/I64 per type
  OP_EMIT,        // sink.emit(arg, acc)
  OP_EMIT_RAW,    // sink.emit_raw(arg, column[flags])  (pass-through attribute)
  OP_EMIT_RAW_BUILD, // sink.emit_raw_build(arg, build column[flags])
  OP_CMP_CODE,    // push((code(col[arg]) - lits[aux]) < lits[aux+1], unsigned) ^ (flags&1): a comparison of a
                  // dictionary-coded attribute with a literal, translated by the host into a code range
  // NULL-able inputs.  col[arg] is the relation's per-row NULL mask (bit a = attribute a is NULL), staged like a
  // LONG column; lits[aux] is the set of attributes the expression at hand reads.
  OP_NOTNULL,     // push((col[arg] & lits[aux]) == 0): a comparison with a NULL operand is false
                  // (LiteralComparators-inl.hpp:168-223)
  OP_NULLSEL,     // acc = (col[arg] & lits[aux]) == 0 ? acc : lits[aux+1]: an aggregate skips NULL arguments
                  // (AggregationHandleSum.hpp:117-127); lits[aux+1] is the identity of the aggregate's combine
  OP_EMIT_NULL,   // sink.emit_null(arg, (col[flags] & lits[aux]) != 0): NULL-ness of projected column `arg`
  // the same two for attributes of a join's BUILD side: the mask is the matched build row's (sink.build_null_mask)
  OP_NOTNULL_BUILD,
  OP_EMIT_NULL_BUILD,
};

enum Leaf : uint8_t { LEAF_COL = 0, LEAF_LIT = 1, LEAF_TMP = 2, LEAF_BUILD = 3 /*join build side*/ };

struct Instr {          // 8 bytes
  uint8_t op;
  uint8_t type;         // compute type (VType)
  uint8_t leaf;         // Leaf kind
  uint8_t ltype;        // native VType of a column leaf
  uint16_t arg;         // column slot / literal index / tmp index / emit slot / lip index
  uint8_t flags;        // bit0: swap operand order
  uint8_t aux;          // comparison id / cvt target / char width low bits
};

// Host-side result of lowering.  The instruction stream is a COMPILE-TIME
// constant of the kernel the query compiler instantiates for it (qs_jit.cu);
// only the literal values travel as a kernel argument (Lits), so the same
// kernel serves every query of the same shape.
struct Lits {
  uint64_t lits[kMaxLits];
  char str_pool[kStrPool];
};

struct Program {
  uint32_t n_pred;      // code[0,n_pred): predicate section (bit-stack result)
  uint32_t n_mid;       // code[n_pred,n_mid): join residual predicate (else == n_pred)
  uint32_t n_total;     // code[n_mid,n_total): emit section
  uint32_t pad;
  Instr code[kMaxInstr];
  Lits L;
};

struct ColDesc {
  const char *ptr;      // device base pointer (row 0 of the relation): native values, or codes when cw != 0
  uint32_t width;       // bytes per native value
  uint32_t smem_off;    // byte offset of this column's native-value tile inside a stage
  // Dictionary-coded attribute (cw = 1, 2 or 4 bytes per row in HBM; 0 = native column).  The tile the TMA
  // unit brings in holds codes at code_off; scalar leaves look values up in `dict` (sorted, native width,
  // readable for every code value of a 1/2-byte code), comparisons with literals run on the codes, and
  // only uses that need the native bytes in place (group-by keys, pass-through projections, join / LIP keys)
  // make the CTA expand the tile into smem_off first (expand != 0).
  const char *dict;
  uint32_t dict_entries;
  uint32_t code_off;
  uint8_t cw, expand;
  // dict_smem: the CTA keeps a copy of the dictionary in shared memory at byte dict_soff of its dynamic shared
  // memory (1-byte codes only: 256 entries, so any code value stays inside the copy)
  uint8_t dict_smem, pad;
  uint32_t dict_soff;
};

struct LipDesc {
  uint64_t *words;      // 64-bit words, MSB-first bits
  int64_t min_value, max_value;
  uint64_t cardinality;
  uint32_t kind;        // QS_LIP_*
  uint32_t is_anti;
  // [0] = rows that probed this filter, [1] = rows it rejected, accumulated by every scan that probes it: what
  // LIPFilterAdaptiveProber keeps per filter (cnt / miss, utility/lip_filter/LIPFilterAdaptiveProber.hpp:103-127)
  unsigned long long *stats;
  // A probe filter of at most kLipSmemBytes is copied into the CTA's dynamic shared memory at byte smem_off before the
  // first tile (plan_scan decides; 0 = probed in global memory / L2): SURVEY.md section 8 north star, "LIP filters in
  // shared memory".  n_words = 64-bit words of the filter.
  uint64_t n_words;
  uint32_t smem_off, pad;
};
constexpr uint32_t kLipSmemBytes = 32u << 10;   // per filter; 48 KB for all filters of one scan

// What a scan kernel iterates over.
struct ScanDesc {
  uint64_t first_row;   // 16-row aligned start (<= row_begin)
  uint64_t row_begin;   // first valid row
  uint64_t row_end;     // one past the last valid row
  const unsigned long long *d_row_end;  // optional device-side row count (min with row_end)
  uint32_t n_tiles;
  uint32_t n_cols;
  uint32_t n_stages;
  uint32_t stage_bytes;
  ColDesc cols[kMaxCols];
  uint32_t n_lip;
  LipDesc lip[kMaxLip];
};

// ----------------------------------------------------------- PTX primitives
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// 1-D bulk copy global -> shared through the TMA unit; completion is signalled
// on `bar` with complete_tx::bytes.  dst/src 16-byte aligned, bytes % 16 == 0.
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ------------------------------------------------------------- bit helpers
// BitVector / BarrieredReadWriteConcurrentBitVector layout: bit i lives in
// word i>>6 at mask (1<<63) >> (i & 63)   (utility/BitVector.hpp:934).
__device__ __forceinline__ bool bv_get(const uint64_t *words, uint64_t bit) {
  return (words[bit >> 6] << (bit & 63)) >> 63;
}
__device__ __forceinline__ void bv_set(uint64_t *words, uint64_t bit) {
  atomicOr(reinterpret_cast<unsigned long long *>(words + (bit >> 6)),
           0x8000000000000000ull >> (bit & 63));
}

// KIND / ANTI are compile-time properties of the kernel (Q::lip_kind); bounds,
// cardinality and the bit words are run-time.
template <uint32_t KIND, uint32_t ANTI>
__device__ __forceinline__ bool lip_contains(const LipDesc &f, const uint64_t *words, int64_t v) {
  if constexpr (KIND == QS_LIP_BITVECTOR_EXACT) {
    if (v < f.min_value || v > f.max_value) return ANTI != 0;
    const bool set = bv_get(words, static_cast<uint64_t>(v - f.min_value));
    return ANTI ? !set : set;
  } else {
    return bv_get(words, static_cast<uint64_t>(v) % f.cardinality);
  }
}
template <uint32_t KIND>
__device__ __forceinline__ void lip_insert(const LipDesc &f, int64_t v) {
  if constexpr (KIND == QS_LIP_BITVECTOR_EXACT) {
    if (v < f.min_value || v > f.max_value) return;   // DCHECK in the reference
    bv_set(f.words, static_cast<uint64_t>(v - f.min_value));
  } else {
    bv_set(f.words, static_cast<uint64_t>(v) % f.cardinality);
  }
}

__device__ __forceinline__ uint64_t mix64(uint64_t x) {   // splitmix64 finalizer
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27; x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}

}  // namespace qs
