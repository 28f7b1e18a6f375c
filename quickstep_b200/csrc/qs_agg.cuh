// Aggregation state layout shared by the K1/K2/K7 kernels and the C-ABI.
#pragma once

#include "qs_common.cuh"

namespace qs {

// How a value word is combined (AggregationHandle*::mergeStates).
enum AggKind : uint8_t {
  AK_SUM_F64 = 0,   // AggregationHandleSum over FLOAT/DOUBLE: double accumulator (AggregationHandleSum.cpp:49-64)
  AK_SUM_I64 = 1,   // over INT/LONG: int64 accumulator
  AK_MIN_F64 = 2, AK_MAX_F64 = 3,   // AggregationHandleMin/Max.hpp, numeric arguments
  AK_MIN_I64 = 4, AK_MAX_I64 = 5,
};

__host__ __device__ inline uint64_t agg_identity(uint8_t kind) {
  switch (kind) {
    case AK_MIN_F64: return 0x7ff0000000000000ull;                 // +inf
    case AK_MAX_F64: return 0xfff0000000000000ull;                 // -inf
    case AK_MIN_I64: return 0x7fffffffffffffffull;
    case AK_MAX_I64: return 0x8000000000000000ull;
    default: return 0;
  }
}

__device__ __forceinline__ uint64_t agg_combine(uint8_t kind, uint64_t a, uint64_t b) {
  switch (kind) {
    case AK_SUM_F64: return static_cast<uint64_t>(__double_as_longlong(__longlong_as_double(a) + __longlong_as_double(b)));
    case AK_SUM_I64: return a + b;
    case AK_MIN_F64: { double x = __longlong_as_double(a), y = __longlong_as_double(b); return y < x ? b : a; }
    case AK_MAX_F64: { double x = __longlong_as_double(a), y = __longlong_as_double(b); return y > x ? b : a; }
    case AK_MIN_I64: return static_cast<int64_t>(b) < static_cast<int64_t>(a) ? b : a;
    default: return static_cast<int64_t>(b) > static_cast<int64_t>(a) ? b : a;
  }
}

// Device-side description of one AggregationOperationState.
struct AggDesc {
  uint32_t n_agg;                    // value words per group (row count is word 0)
  uint32_t words;                    // n_agg + 1
  uint8_t kind[kMaxAgg];
  // group key: packed like ThreadPrivateCompactKeyHashTable::ConstructKeyCode
  // (storage/ThreadPrivateCompactKeyHashTable.hpp:125-142): key i is memcpy'd
  // at byte offset sum(widths of keys < i) of a zeroed word array.
  uint32_t n_key_cols;
  uint32_t key_words;                // 1 for the compact strategy
  uint16_t key_col[kMaxKeyCols];     // staged column slot
  uint8_t key_width[kMaxKeyCols];
  uint8_t key_off[kMaxKeyCols];
  uint32_t strategy;                 // QS_AGG_*
  // --- compact / single: per-CTA partials + global key directory
  uint64_t *partials;                // [grid][partial_rows][words]
  uint32_t partial_rows;             // kCompactMaxGroups, or 1 without GROUP BY
  uint32_t pad0;
  uint64_t *dir_keys;                // [dir_cap]
  int *dir_gid;                      // [dir_cap]  -1 empty, -2 busy, >=0 group id
  uint32_t dir_cap;                  // power of two
  uint32_t *n_groups;                // dense group counter
  uint64_t *gid_keys;                // [kCompactMaxGroups] key of each dense id
  // --- separate chaining (open addressing on device) / collision free
  uint32_t *tags;                    // [cap] 0 empty, 1 busy, 2 ready
  uint64_t *keys;                    // [cap][key_words]
  uint64_t *states;                  // [cap][words]  (also the compact totals)
  uint64_t cap;                      // slots (pow2) / num_entries (collision free)
  uint32_t *error_flag;              // set to QSGPU_ERR_CAPACITY on overflow
};

}  // namespace qs
