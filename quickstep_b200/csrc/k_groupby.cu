// K7: hash and dense GROUP BY aggregation over a global table, plus finalize.
//
//   generic keys (<= 32 B)  PackedPayloadHashTable::upsertValueAccessorCompositeKeyInternal
//                           (storage/PackedPayloadHashTable.hpp:780-909)  -- the
//                           reference chains buckets under a SpinMutex; here
//                           open addressing with a per-slot tag (CAS claim) and
//                           native global RED atomics on the state words.   (TPC-H Q3)
//   dense INT/LONG key      CollisionFreeVectorTable::upsertValueAccessor*
//                           (storage/CollisionFreeVectorTable.hpp:530-645): slot = key,
//                           fetch_add / CAS -> RED.ADD; the row-count word doubles
//                           as the existence bit (…:543).
//   finalize                AggregationOperationState::finalizeAggregate
//                           (storage/AggregationOperationState.cpp:641-948),
//                           AggregationHandleAvg::finalize (AggregationHandleAvg.cpp:144-155)
#include "qs_ops.cuh"
#include "qs_vm.cuh"

namespace qs {

__device__ __forceinline__ void global_update(uint8_t kind, uint64_t *p, uint64_t v) {
  switch (kind) {
    case AK_SUM_F64: atomicAdd(reinterpret_cast<double *>(p), u2d(v)); break;
    case AK_SUM_I64: atomicAdd(reinterpret_cast<unsigned long long *>(p), static_cast<unsigned long long>(v)); break;
    case AK_MIN_I64: atomicMin(reinterpret_cast<long long *>(p), static_cast<long long>(v)); break;
    case AK_MAX_I64: atomicMax(reinterpret_cast<long long *>(p), static_cast<long long>(v)); break;
    default: {
      unsigned long long *q = reinterpret_cast<unsigned long long *>(p);
      unsigned long long old = *q;
      while (true) {
        const uint64_t want = agg_combine(kind, old, v);
        if (want == old) break;
        const unsigned long long seen = atomicCAS(q, old, static_cast<unsigned long long>(want));
        if (seen == old) break;
        old = seen;
      }
    }
  }
}

// Find-or-insert `key` (kw words); returns the slot or -1 when the table is full.
__device__ __forceinline__ int64_t table_upsert(const uint64_t *key, uint32_t kw, const AggDesc &A) {
  uint64_t h = 0x9e3779b97f4a7c15ull;
  for (uint32_t i = 0; i < kw; ++i) h = mix64(h ^ key[i]);
  const uint64_t mask = A.cap - 1;
  uint64_t slot = h & mask;
  volatile uint32_t *tags = A.tags;
  volatile uint64_t *keys = A.keys;
  for (uint64_t probes = 0; probes <= mask;) {
    const uint32_t t = tags[slot];
    if (t == 2u) {
      bool eq = true;
      for (uint32_t i = 0; i < kw; ++i) eq &= keys[slot * kw + i] == key[i];
      if (eq) return static_cast<int64_t>(slot);
      slot = (slot + 1) & mask;
      ++probes;
      continue;
    }
    if (t == 0u && atomicCAS(&A.tags[slot], 0u, 1u) == 0u) {
      for (uint32_t i = 0; i < kw; ++i) keys[slot * kw + i] = key[i];
      __threadfence();
      tags[slot] = 2u;
      atomicAdd(A.n_groups, 1u);
      return static_cast<int64_t>(slot);
    }
    // busy (or lost the race): look at the same slot again
  }
  atomicExch(A.error_flag, static_cast<uint32_t>(QSGPU_ERR_CAPACITY));
  return -1;
}

struct GlobalAggSink : SinkBase {
  int64_t slot[kRows];
  const AggDesc *A;
  __device__ __forceinline__ void emit(uint32_t j, uint8_t, const uint64_t (&acc)[kRows]) {
    const uint8_t kind = A->kind[j];
#pragma unroll
    for (int r = 0; r < kRows; ++r)
      if (slot[r] >= 0) global_update(kind, &A->states[slot[r] * A->words + 1 + j], acc[r]);
  }
};

__global__ void __launch_bounds__(kBlock, 2)
k_scan_groupby(const __grid_constant__ ScanDesc S, const __grid_constant__ Program P,
               const __grid_constant__ AggDesc A) {
  extern __shared__ __align__(128) char smem[];
  const int tid = threadIdx.x;
  GlobalAggSink sink;
  sink.A = &A;
  VmRegs regs;
  scan_tiles(S, smem, [&](uint32_t tile, const char *stage, const ScanRt &rt) {
    bool valid[kRows];
    tile_valid(S, rt, tile, tid, valid);
    uint32_t bits[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) bits[r] = 1u;
    SinkBase ns;
    vm_run(P, 0, P.n_pred, S, stage, tid, regs, bits, ns);
    bool any = false;
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      const bool pass = valid[r] && (bits[r] & 1u);
      sink.slot[r] = -1;
      if (pass) {
        if (A.strategy == QS_AGG_COLLISION_FREE) {
          const uint32_t w = A.key_width[0];
          const char *src = stage + S.cols[A.key_col[0]].smem_off + tile_row(r, tid) * w;
          const int64_t k = w == 4 ? static_cast<int64_t>(*reinterpret_cast<const int32_t *>(src))
                                   : *reinterpret_cast<const int64_t *>(src);
          if (k >= 0 && static_cast<uint64_t>(k) < A.cap) sink.slot[r] = k;
          else atomicExch(A.error_flag, static_cast<uint32_t>(QSGPU_ERR_CAPACITY));
        } else {
          uint64_t key[kMaxKeyWords] = {0, 0, 0, 0};
          for (uint32_t k = 0; k < A.n_key_cols; ++k) {
            const uint32_t w = A.key_width[k];
            const char *src = stage + S.cols[A.key_col[k]].smem_off + tile_row(r, tid) * w;
            for (uint32_t b = 0; b < w; ++b) {
              const uint32_t pos = A.key_off[k] + b;
              key[pos >> 3] |= static_cast<uint64_t>(static_cast<unsigned char>(src[b])) << (8 * (pos & 7));
            }
          }
          sink.slot[r] = table_upsert(key, A.key_words, A);
        }
        if (sink.slot[r] >= 0)
          atomicAdd(reinterpret_cast<unsigned long long *>(&A.states[sink.slot[r] * A.words]), 1ull);
      }
      any |= sink.slot[r] >= 0;
    }
    if (!__any_sync(0xffffffffu, any)) return;
    vm_run(P, P.n_mid, P.n_total, S, stage, tid, regs, bits, sink);
  });
}

// Rehash every ready slot of `from` into `to` (table growth between work orders).
__global__ void k_rehash(const __grid_constant__ AggDesc from, const __grid_constant__ AggDesc to) {
  for (uint64_t s = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; s < from.cap;
       s += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    if (from.tags[s] != 2u) continue;
    uint64_t key[kMaxKeyWords];
    for (uint32_t i = 0; i < from.key_words; ++i) key[i] = from.keys[s * from.key_words + i];
    const int64_t d = table_upsert(key, from.key_words, to);
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
      slot = table_upsert(key, A.key_words, A);
      if (slot < 0) continue;
    }
    for (uint32_t w = 0; w < A.words; ++w) {
      const uint8_t kind = w == 0 ? AK_SUM_I64 : A.kind[w - 1];
      global_update(kind, &A.states[slot * A.words + w], f_states[g * A.words + w]);
    }
  }
}

// List the occupied slots (row count > 0) of a hash / dense table.
__global__ void k_collect_slots(const uint64_t *states, uint32_t words, uint64_t cap, uint64_t *out_idx,
                                unsigned long long *counter) {
  const uint64_t s = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
  const bool occ = s < cap && states[s * words] != 0;
  const uint32_t ballot = __ballot_sync(0xffffffffu, occ);
  if (ballot == 0) return;
  const int lane = threadIdx.x & 31;
  unsigned long long base = 0;
  if (lane == 0) base = atomicAdd(counter, static_cast<unsigned long long>(__popc(ballot)));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (occ) out_idx[base + __popc(ballot & ((1u << lane) - 1))] = s;
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
    const uint64_t count = states[s * words];
    for (uint32_t j = 0; j < F.n_out; ++j) {
      const uint64_t v = states[s * words + F.word[j]];
      uint64_t o;
      uint8_t from;
      if (F.function[j] == QS_AGG_COUNT) { o = count; from = V_I64; }
      else if (F.function[j] == QS_AGG_AVG) {
        // sum / static_cast<double>(count)   (AggregationHandleAvg.cpp:144-155)
        const double sum = F.word_is_f64[j] ? u2d(v) : static_cast<double>(static_cast<int64_t>(v));
        o = d2u(count ? sum / static_cast<double>(static_cast<int64_t>(count)) : 0.0);
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
  }
}

// ------------------------------------------------------------------ launchers
cudaError_t launch_scan_groupby(const ScanDesc &S, const Program &P, const AggDesc &A, int grid, size_t smem,
                                cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(k_scan_groupby, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  k_scan_groupby<<<grid, kBlock, smem, st>>>(S, P, A);
  return cudaGetLastError();
}

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
                                 unsigned long long *counter, cudaStream_t st) {
  const uint64_t blocks = (cap + 255) / 256;
  k_collect_slots<<<static_cast<unsigned>(blocks), 256, 0, st>>>(states, words, cap, out_idx, counter);
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
  if (n == 0) return cudaSuccess;
  k_finalize<<<grid_for(n, 256), 256, 0, st>>>(states, keys, words, idx, n, F);
  return cudaGetLastError();
}

}  // namespace qs
