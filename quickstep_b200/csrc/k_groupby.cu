// K7: fixed kernels of the hash / dense GROUP BY path (rehash, foreign merge, slot
// collection, finalize); the scan kernel is qs_kernels.cuh scan_groupby_body.
//
//   generic keys (<= 32 B)  PackedPayloadHashTable::upsertValueAccessorCompositeKeyInternal
//                           (storage/PackedPayloadHashTable.hpp:780-909)  -- the
//                           reference chains buckets under a SpinMutex; here
//                           open addressing, keys claimed by one 64 / 128-bit CAS
//                           (a per-slot tag for keys wider than 16 bytes) and
//                           native global RED atomics on the state words.   (TPC-H Q3)
//   dense INT/LONG key      CollisionFreeVectorTable::upsertValueAccessor*
//                           (storage/CollisionFreeVectorTable.hpp:530-645): slot = key,
//                           fetch_add / CAS -> RED.ADD; the row-count word doubles
//                           as the existence bit (…:543).
//   finalize                AggregationOperationState::finalizeAggregate
//                           (storage/AggregationOperationState.cpp:641-948),
//                           AggregationHandleAvg::finalize (AggregationHandleAvg.cpp:144-155)
#include "qs_kernels.cuh"

namespace qs {

__device__ __forceinline__ void global_update(uint8_t kind, uint64_t *p, uint64_t v) {
  switch (kind) {
    case AK_SUM_F64: atomic_update<AK_SUM_F64>(p, v); break;
    case AK_SUM_I64: atomic_update<AK_SUM_I64>(p, v); break;
    case AK_MIN_I64: atomic_update<AK_MIN_I64>(p, v); break;
    case AK_MAX_I64: atomic_update<AK_MAX_I64>(p, v); break;
    case AK_MIN_F64: atomic_update<AK_MIN_F64>(p, v); break;
    default: atomic_update<AK_MAX_F64>(p, v); break;
  }
}

__device__ __forceinline__ int64_t table_upsert_rt(const uint64_t *key, uint32_t kw, const AggDesc &A) {
  switch (kw) {
    case 1: return table_upsert<1>(key, A);
    case 2: return table_upsert<2>(key, A);
    case 3: return table_upsert<3>(key, A);
    default: return table_upsert<4>(key, A);
  }
}

// Rehash every ready slot of `from` into `to` (table growth between work orders).
__global__ void k_rehash(const __grid_constant__ AggDesc from, const __grid_constant__ AggDesc to) {
  // (a slot is occupied iff its row count is non-zero: holds between kernels for both key protocols; the table has
  // cap + 1 rows, the last one reserved for the all-ones key of the CAS-claimed protocol)
  for (uint64_t s = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; s <= from.cap;
       s += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    if (from.states[s * from.words] == 0) continue;
    uint64_t key[kMaxKeyWords];
    for (uint32_t i = 0; i < from.key_words; ++i) key[i] = from.keys[s * from.key_words + i];
    const int64_t d = table_upsert_rt(key, from.key_words, to);
    if (d < 0) continue;
    for (uint32_t w = 0; w < from.words; ++w) to.states[d * to.words + w] = from.states[s * from.words + w];
  }
}

// Merge a foreign partial table (dense rows: states + keys) into a hash / dense state.
__global__ void k_merge_foreign_table(const __grid_constant__ AggDesc A, const uint64_t *f_states,
                                      const uint64_t *f_keys, uint64_t f_groups) {
  for (uint64_t g = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; g < f_groups;
       g += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    if (f_states[g * A.words] == 0) continue;
    int64_t slot;
    if (A.strategy == QS_AGG_COLLISION_FREE) {
      slot = static_cast<int64_t>(f_keys[g]);
      if (slot < 0 || static_cast<uint64_t>(slot) >= A.cap) continue;
    } else {
      uint64_t key[kMaxKeyWords];
      for (uint32_t i = 0; i < A.key_words; ++i) key[i] = f_keys[g * A.key_words + i];
      slot = table_upsert_rt(key, A.key_words, A);
      if (slot < 0) continue;
    }
    for (uint32_t w = 0; w < A.words; ++w) {
      const uint8_t kind = w == 0 ? AK_SUM_I64 : A.kind[w - 1];
      global_update(kind, &A.states[slot * A.words + w], f_states[g * A.words + w]);
    }
  }
}

// List the occupied slots (row count > 0) of a hash / dense table.  One CTA covers 4096 slots and makes
// ONE reservation on the global counter (a ballot-per-warp version spent its time serialising 131k
// same-address atomics for a 4M-slot table: 228 us in the r01b launch list).
constexpr int kCollectPerThread = 16;
__global__ void __launch_bounds__(256) k_collect_slots(const uint64_t *states, uint32_t words, uint64_t cap,
                                                       uint64_t *out_idx, unsigned long long *counter,
                                                       const uint64_t *exist_words) {
  __shared__ uint32_t s_warp[8];
  __shared__ unsigned long long s_base;
  const uint64_t base_slot = static_cast<uint64_t>(blockIdx.x) * (256 * kCollectPerThread);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t occ = 0;                       // bit i: slot base + i*256 + tid is occupied
#pragma unroll
  for (int i = 0; i < kCollectPerThread; ++i) {
    const uint64_t s = base_slot + static_cast<uint64_t>(i) * 256 + threadIdx.x;
    // a dense table's group also exists when the existence map says so, even with no rows (COUNT = 0)
    if (s < cap && (states[s * words] != 0 || (exist_words && bv_get(exist_words, s)))) occ |= 1u << i;
  }
  const uint32_t mine = __popc(occ);
  uint32_t incl = mine;                   // inclusive scan inside the warp
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += y;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t total = 0;
    for (int w = 0; w < 8; ++w) { const uint32_t c = s_warp[w]; s_warp[w] = total; total += c; }
    s_base = total ? atomicAdd(counter, static_cast<unsigned long long>(total)) : 0ull;
  }
  __syncthreads();
  unsigned long long pos = s_base + s_warp[warp] + (incl - mine);
#pragma unroll
  for (int i = 0; i < kCollectPerThread; ++i)
    if (occ & (1u << i)) out_idx[pos++] = base_slot + static_cast<uint64_t>(i) * 256 + threadIdx.x;
}

// Dense copy of (states, keys) rows for qsgpu_agg_partial.
__global__ void k_gather_rows(const uint64_t *states, const uint64_t *keys, uint32_t words, uint32_t kw,
                              const uint64_t *idx, uint64_t n, uint64_t *o_states, uint64_t *o_keys,
                              int keys_are_slots) {
  for (uint64_t g = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; g < n;
       g += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint64_t s = idx[g];
    for (uint32_t w = 0; w < words; ++w) o_states[g * words + w] = states[s * words + w];
    if (keys_are_slots) o_keys[g] = s;
    else for (uint32_t i = 0; i < kw; ++i) o_keys[g * kw + i] = keys[s * kw + i];
  }
}


__global__ void k_finalize(const uint64_t *states, const uint64_t *keys, uint32_t words, const uint64_t *idx,
                           uint64_t n, const __grid_constant__ FinalizeDesc F) {
  if (F.d_n_groups) { const uint64_t live = *F.d_n_groups; if (live < n) n = live; }
  if (blockIdx.x == 0 && threadIdx.x == 0 && F.rows_out) *F.rows_out = n;
  for (uint64_t g = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; g < n;
       g += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint64_t s = idx ? idx[g] : g;
    for (uint32_t k = 0; k < F.n_key_cols; ++k) {
      const uint32_t w = F.key_width[k];
      char *dst = F.key_out[k] + g * w;
      if (F.keys_are_slots) {
        if (w == 4) *reinterpret_cast<int32_t *>(dst) = static_cast<int32_t>(s);
        else *reinterpret_cast<int64_t *>(dst) = static_cast<int64_t>(s);
      } else {
        for (uint32_t b = 0; b < w; ++b) {
          const uint32_t pos = F.key_off[k] + b;
          dst[b] = static_cast<char>(keys[s * F.key_words + (pos >> 3)] >> (8 * (pos & 7)));
        }
      }
    }
    uint64_t nulls = 0;
    for (uint32_t j = 0; j < F.n_out; ++j) {
      const uint64_t v = states[s * words + F.word[j]];
      // rows that count for this aggregate: the group's rows, or those whose (NULL-able) argument was not NULL
      const uint64_t count = states[s * words + F.nn_word[j]];
      if (count == 0) nulls |= 1ull << (F.n_key_cols + j);
      uint64_t o;
      uint8_t from;
      if (F.function[j] == QS_AGG_COUNT) { o = count; from = V_I64; }
      else if (F.function[j] == QS_AGG_AVG) {
        // sum / static_cast<double>(count)   (AggregationHandleAvg.cpp:144-155)
        const double sum = F.word_is_f64[j] ? u2d(v) : static_cast<double>(static_cast<int64_t>(v));
        // Without GROUP BY an AVG over no values is NULL (stored as 0).  With GROUP BY the reference divides whatever the
        // entry holds, AggregationHandleAvg.hpp:180-189: a group whose arguments were all NULL prints 0 / 0.0 = NaN
        // (verified on the unmodified engine, tests/golden/ref_null_results.json).
        o = d2u((count || F.n_key_cols) ? sum / static_cast<double>(static_cast<int64_t>(count)) : 0.0);
        from = V_F64;
      } else {
        o = count ? v : 0;
        from = F.word_is_f64[j] ? V_F64 : V_I64;
      }
      o = vcvt(o, from, F.out_vtype[j]);
      if (F.out_vtype[j] == V_I32 || F.out_vtype[j] == V_F32)
        *reinterpret_cast<uint32_t *>(F.out[j] + g * 4) = static_cast<uint32_t>(o);
      else
        *reinterpret_cast<uint64_t *>(F.out[j] + g * 8) = o;
    }
    if (F.null_out) F.null_out[g] = nulls & F.null_bits;
  }
}

// ------------------------------------------------------------------ launchers
static int grid_for(uint64_t n, int block) {
  uint64_t g = (n + block - 1) / block;
  if (g > 148ull * 16) g = 148ull * 16;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

cudaError_t launch_rehash(const AggDesc &from, const AggDesc &to, cudaStream_t st) {
  k_rehash<<<grid_for(from.cap, 256), 256, 0, st>>>(from, to);
  return cudaGetLastError();
}

cudaError_t launch_merge_foreign_table(const AggDesc &A, const uint64_t *f_states, const uint64_t *f_keys,
                                       uint64_t f_groups, cudaStream_t st) {
  if (f_groups == 0) return cudaSuccess;
  k_merge_foreign_table<<<grid_for(f_groups, 256), 256, 0, st>>>(A, f_states, f_keys, f_groups);
  return cudaGetLastError();
}

cudaError_t launch_collect_slots(const uint64_t *states, uint32_t words, uint64_t cap, uint64_t *out_idx,
                                 unsigned long long *counter, const uint64_t *exist_words, cudaStream_t st) {
  const uint64_t per_block = 256ull * kCollectPerThread;
  const uint64_t blocks = (cap + per_block - 1) / per_block;
  k_collect_slots<<<static_cast<unsigned>(blocks), 256, 0, st>>>(states, words, cap, out_idx, counter, exist_words);
  return cudaGetLastError();
}

cudaError_t launch_gather_rows(const uint64_t *states, const uint64_t *keys, uint32_t words, uint32_t kw,
                               const uint64_t *idx, uint64_t n, uint64_t *o_states, uint64_t *o_keys,
                               int keys_are_slots, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  k_gather_rows<<<grid_for(n, 256), 256, 0, st>>>(states, keys, words, kw, idx, n, o_states, o_keys,
                                                  keys_are_slots);
  return cudaGetLastError();
}

cudaError_t launch_finalize(const uint64_t *states, const uint64_t *keys, uint32_t words, const uint64_t *idx,
                            uint64_t n, const FinalizeDesc &F, cudaStream_t st) {
  k_finalize<<<grid_for(n ? n : 1, 256), 256, 0, st>>>(states, keys, words, idx, n, F);
  return cudaGetLastError();
}

}  // namespace qs
