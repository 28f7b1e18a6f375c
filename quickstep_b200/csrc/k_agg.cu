// Fixed (query-independent) kernels of the K1 / K2 aggregation path; the scan
// kernel itself is qs_kernels.cuh scan_agg_body, instantiated per query by the
// query compiler (qs_jit.cu).
//
//   K1  no GROUP BY     AggregationOperationState::aggregateBlockSingleState
//                       (storage/AggregationOperationState.cpp:476-519) +
//                       AggregationHandleSum::accumulateValueAccessor ->
//                       ArithmeticBinaryOperators.hpp:714-744          (TPC-H Q6)
//   K2  compact key     ThreadPrivateCompactKeyHashTable::upsertValueAccessorCompositeKey
//                       + mergeFrom (storage/ThreadPrivateCompactKeyHashTable.cpp:203-363)
//                                                                      (TPC-H Q1)
//
// Each scan CTA writes ONE partial state per group; k_merge_partials folds them
// in a fixed order (lane-strided over CTAs, then a lane-ordered shuffle tree),
// so a run is reproducible bit for bit -- the reference's own result depends on
// work-order completion order.
#include <algorithm>

#include "qs_jit.h"
#include "qs_kernels.cuh"

namespace qs {

// Fill [n_rows][words] state rows with identities (row count = 0).
__global__ void k_fill_identity(uint64_t *states, uint64_t n_rows, const __grid_constant__ AggDesc A) {
  const uint64_t n = n_rows * A.words;
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint32_t w = static_cast<uint32_t>(i % A.words);
    states[i] = w == 0 ? 0 : agg_identity(A.kind[w - 1]);
  }
}

// totals[g][w] (+)= fold over CTAs.  One warp per (group, word): lane l folds
// CTAs l, l+32, ... in order, then the 32 lane partials are combined by a
// shuffle tree with lane-ordered operands (deterministic).
__global__ void __launch_bounds__(128) k_merge_partials(const __grid_constant__ AggDesc A, uint32_t n_ctas) {
  const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  // without GROUP BY there is exactly one state and the key directory is never touched
  const uint32_t n_groups = A.n_key_cols == 0 ? 1u : min(*A.n_groups, A.partial_rows);
  if (i >= n_groups * A.words) return;
  const uint32_t g = i / A.words, w = i % A.words;
  const uint8_t kind = w == 0 ? AK_SUM_I64 : A.kind[w - 1];
  const uint64_t ident = w == 0 ? 0 : agg_identity(kind);
  uint64_t x = ident;
  for (uint32_t c = lane; c < n_ctas; c += 32) {
    uint64_t *p = &A.partials[(static_cast<uint64_t>(c) * A.partial_rows + g) * A.words + w];
    x = agg_combine(kind, x, *p);
    *p = ident;                     // consumed: ready for the next work order
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const uint64_t y = __shfl_xor_sync(0xffffffffu, x, off);
    x = (lane & off) ? agg_combine(kind, y, x) : agg_combine(kind, x, y);
  }
  if (lane == 0) A.states[i] = agg_combine(kind, A.states[i], x);
}

// Merge a foreign partial (another GPU's totals) into this state, compact strategy.
__global__ void k_merge_foreign_compact(const __grid_constant__ AggDesc A, const uint64_t *f_states,
                                        const uint64_t *f_keys, uint32_t f_groups) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= f_groups) return;
  if (f_states[static_cast<uint64_t>(g) * A.words] == 0) return;   // empty group
  const int gid = dir_insert(f_keys[g], A);
  if (gid < 0) return;                      // group limit exceeded: QSGPU_ERR_CAPACITY was raised
  for (uint32_t w = 0; w < A.words; ++w) {
    const uint8_t kind = w == 0 ? AK_SUM_I64 : A.kind[w - 1];
    uint64_t *dst = &A.states[static_cast<uint64_t>(gid) * A.words + w];
    *dst = agg_combine(kind, *dst, f_states[static_cast<uint64_t>(g) * A.words + w]);
  }
}

// Initial contents of a fresh fixed-size state in ONE launch (qsgpu_agg_create is on the critical path of every
// query: the scan cannot start before its state exists).
__global__ void k_agg_init(const __grid_constant__ AggDesc A, uint64_t partial_sets, unsigned long long *ctl) {
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  const uint64_t t0 = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
  const uint64_t n_partial = partial_sets * A.partial_rows * A.words;
  for (uint64_t i = t0; i < n_partial; i += stride) {
    const uint32_t w = static_cast<uint32_t>(i % A.words);
    A.partials[i] = w == 0 ? 0 : agg_identity(A.kind[w - 1]);
  }
  const uint64_t n_states = static_cast<uint64_t>(A.partial_rows) * A.words;
  for (uint64_t i = t0; i < n_states; i += stride) {
    const uint32_t w = static_cast<uint32_t>(i % A.words);
    A.states[i] = w == 0 ? 0 : agg_identity(A.kind[w - 1]);
  }
  for (uint64_t i = t0; i < A.partial_rows; i += stride) A.gid_keys[i] = 0;
  for (uint64_t i = t0; i < A.dir_cap; i += stride) A.dir_gid[i] = -1;
  for (uint64_t i = t0; i < 96; i += stride) ctl[i] = 0;          // 3 x 256 bytes of counters
}

struct PackDesc { ColDesc cols[kMaxCols]; };
// dst: [u64 rows][u32 error, u32 0][nulls: max_rows x u64 when d_nulls][column 0: max_rows x w0, padded to 16]...
__global__ void k_pack_rows(char *dst, const __grid_constant__ PackDesc P, uint32_t n_cols, uint64_t max_rows,
                            const unsigned long long *d_rows, const unsigned long long *d_nulls, int with_nulls,
                            uint32_t *error_flag) {
  const uint64_t n = min(static_cast<uint64_t>(*d_rows), max_rows);
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    *reinterpret_cast<uint64_t *>(dst) = *d_rows;
    reinterpret_cast<uint32_t *>(dst)[2] = *error_flag;
    reinterpret_cast<uint32_t *>(dst)[3] = 0;
    *error_flag = 0;                                 // reported with this result
  }
  size_t off = 16;
  const uint64_t t0 = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  if (with_nulls) {        // a relation without a NULL mask holds no NULLs
    for (uint64_t i = t0; i < n; i += stride) reinterpret_cast<unsigned long long *>(dst + off)[i] = d_nulls ? d_nulls[i] : 0ull;
    off += (max_rows * 8 + 15) & ~static_cast<size_t>(15);
  }
  for (uint32_t c = 0; c < n_cols; ++c) {
    const uint64_t bytes = n * P.cols[c].width;
    for (uint64_t i = t0; i < bytes; i += stride) dst[off + i] = P.cols[c].ptr[i];
    off += (max_rows * P.cols[c].width + 15) & ~static_cast<size_t>(15);
  }
}

// ------------------------------------------------------------------ launchers
cudaError_t launch_agg_init(const AggDesc &A, uint64_t partial_sets, void *ctl, cudaStream_t st) {
  const uint64_t n = std::max<uint64_t>(partial_sets * A.partial_rows * A.words, A.dir_cap);
  int grid = static_cast<int>(std::min<uint64_t>((n + 255) / 256, 148 * 8));
  k_agg_init<<<std::max(grid, 1), 256, 0, st>>>(A, partial_sets, static_cast<unsigned long long *>(ctl));
  return cudaGetLastError();
}

cudaError_t launch_pack_rows(char *dst, const ColDesc *cols, uint32_t n_cols, uint64_t max_rows,
                             const unsigned long long *d_rows, const unsigned long long *d_nulls, bool with_nulls,
                             uint32_t *error_flag, cudaStream_t st) {
  PackDesc P{};
  uint64_t bytes = 0;
  for (uint32_t c = 0; c < n_cols; ++c) { P.cols[c] = cols[c]; bytes += max_rows * cols[c].width; }
  const int grid = static_cast<int>(std::min<uint64_t>((bytes + 255) / 256, 148));
  k_pack_rows<<<std::max(grid, 1), 256, 0, st>>>(dst, P, n_cols, max_rows, d_rows, d_nulls, with_nulls ? 1 : 0, error_flag);
  return cudaGetLastError();
}

size_t agg_smem_extra(int hot, int n_agg, bool grouped, uint32_t words, bool priv) {
  const size_t NA = n_agg > 0 ? n_agg : 1;
  const size_t LG = grouped ? kCompactMaxGroups : 1, LS = grouped ? kCompactLocalSlots : 0;
  // priv: hot groups x NA value words x one 8-byte slot per thread (row counts stay in registers)
  const size_t pv = priv ? static_cast<size_t>(hot) * NA * kBlock * 8 : 0;
  return static_cast<size_t>(8) * hot * (NA + 1) * 8 + LG * words * 8 + LG * 8 + LS * 8 + LS * 4 + LG * 4 + 16 + pv;
}

int agg_hot_groups(const AggDesc &A) { return A.n_key_cols > 0 ? 4 : 1; }

cudaError_t launch_fill_identity(uint64_t *states, uint64_t n_rows, const AggDesc &A, cudaStream_t st) {
  const uint64_t n = n_rows * A.words;
  int grid = static_cast<int>((n + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  if (grid < 1) grid = 1;
  k_fill_identity<<<grid, 256, 0, st>>>(states, n_rows, A);
  return cudaGetLastError();
}

cudaError_t launch_merge_partials(const AggDesc &A, uint32_t n_ctas, cudaStream_t st) {
  const uint32_t n = A.partial_rows * A.words;          // upper bound; the kernel reads the live group count
  const uint32_t warps_per_block = 4;
  uint32_t blocks = (n + warps_per_block - 1) / warps_per_block;
  if (A.partial_rows > 1 && blocks > 148u * 4) {
    // groups are dense ids 0..n_groups-1 and n_groups <= kCompactMaxGroups: cover them all
    blocks = (n + warps_per_block - 1) / warps_per_block;
  }
  k_merge_partials<<<blocks, 32 * warps_per_block, 0, st>>>(A, n_ctas);
  return cudaGetLastError();
}

cudaError_t launch_merge_foreign_compact(const AggDesc &A, const uint64_t *f_states, const uint64_t *f_keys,
                                         uint32_t f_groups, cudaStream_t st) {
  if (f_groups == 0) return cudaSuccess;
  k_merge_foreign_compact<<<(f_groups + 127) / 128, 128, 0, st>>>(A, f_states, f_keys, f_groups);
  return cudaGetLastError();
}

}  // namespace qs
