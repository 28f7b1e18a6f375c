// K1 / K2: fused predicate scan + expression evaluation + aggregation.
//
//   K1  no GROUP BY     AggregationOperationState::aggregateBlockSingleState
//                       (storage/AggregationOperationState.cpp:476-519) +
//                       AggregationHandleSum::accumulateValueAccessor ->
//                       ArithmeticBinaryOperators.hpp:714-744          (TPC-H Q6)
//   K2  compact key     ThreadPrivateCompactKeyHashTable::upsertValueAccessorCompositeKey
//                       + mergeFrom (storage/ThreadPrivateCompactKeyHashTable.cpp:203-363)
//                                                                      (TPC-H Q1)
//
// One persistent CTA pair per SM.  Column tiles arrive in shared memory through
// the TMA bulk-copy ring (qs_vm.cuh); the VM evaluates predicate and aggregate
// arguments in registers; the first HOT groups a CTA meets are accumulated in
// per-thread registers (no atomics at all for Q1's four groups), later groups
// in a per-CTA shared-memory table.  Each CTA writes ONE partial state per
// group; k_merge_partials folds them in CTA order, so a run is reproducible
// bit for bit -- the reference's own result depends on work-order completion
// order.
#include "qs_ops.cuh"
#include "qs_vm.cuh"

namespace qs {

struct AggSmem {
  uint64_t *red;         // [8 warps][HOT*(NAGG+1)]
  uint64_t *lstate;      // [LG][words]
  uint64_t *lkey_by_id;  // [LG]
  uint64_t *lkeys;       // [LS]
  int *lslot;            // [LS]
  uint32_t *nlocal;
};

__device__ __forceinline__ void cold_update(uint8_t kind, uint64_t *p, uint64_t v) {
  switch (kind) {
    case AK_SUM_F64: atomicAdd(reinterpret_cast<double *>(p), u2d(v)); break;
    case AK_SUM_I64: atomicAdd(reinterpret_cast<unsigned long long *>(p), static_cast<unsigned long long>(v)); break;
    case AK_MIN_I64: atomicMin(reinterpret_cast<long long *>(p), static_cast<long long>(v)); break;
    case AK_MAX_I64: atomicMax(reinterpret_cast<long long *>(p), static_cast<long long>(v)); break;
    default: {   // MIN/MAX over doubles: CAS loop
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

template <int HOT, int NAGG>
struct AggSink : SinkBase {
  uint64_t hv[HOT][NAGG];
  uint32_t hc[HOT];
  int slot[kRows];
  const AggDesc *A;
  uint64_t *lstate;

  __device__ __forceinline__ void emit(uint32_t j, uint8_t, const uint64_t (&acc)[kRows]) {
#pragma unroll
    for (int jj = 0; jj < NAGG; ++jj) {
      if (jj != static_cast<int>(j)) continue;
      const uint8_t kind = A->kind[jj];
      if (kind == AK_SUM_F64) {
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
#pragma unroll
          for (int g = 0; g < HOT; ++g)
            if (slot[r] == g) hv[g][jj] = d2u(u2d(hv[g][jj]) + u2d(acc[r]));
        }
      } else if (kind == AK_SUM_I64) {
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
#pragma unroll
          for (int g = 0; g < HOT; ++g)
            if (slot[r] == g) hv[g][jj] += acc[r];
        }
      } else {
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
#pragma unroll
          for (int g = 0; g < HOT; ++g)
            if (slot[r] == g) hv[g][jj] = agg_combine(kind, hv[g][jj], acc[r]);
        }
      }
#pragma unroll
      for (int r = 0; r < kRows; ++r)
        if (slot[r] >= HOT) cold_update(kind, &lstate[slot[r] * A->words + 1 + jj], acc[r]);
    }
  }
};

__device__ __forceinline__ int local_lookup(uint64_t key, const AggSmem &M, uint32_t LS, uint32_t LG,
                                            uint32_t *error_flag) {
  uint32_t h = static_cast<uint32_t>(mix64(key)) & (LS - 1);
  volatile int *lslot = M.lslot;
  volatile uint64_t *lkeys = M.lkeys;
  while (true) {
    const int s = lslot[h];
    if (s >= 0) {
      if (lkeys[h] == key) return s;
      h = (h + 1) & (LS - 1);
      continue;
    }
    if (s == -1 && atomicCAS(&M.lslot[h], -1, -2) == -1) {
      uint32_t id = atomicAdd(M.nlocal, 1u);
      if (id >= LG) {
        atomicExch(error_flag, static_cast<uint32_t>(QSGPU_ERR_CAPACITY));
        id = 0;
      } else {
        M.lkey_by_id[id] = key;
      }
      lkeys[h] = key;
      __threadfence_block();
      lslot[h] = static_cast<int>(id);
      return static_cast<int>(id);
    }
  }
}

// Global key -> dense group id directory (persists across work orders).
__device__ __forceinline__ int dir_insert(uint64_t key, const AggDesc &A) {
  uint32_t h = static_cast<uint32_t>(mix64(key)) & (A.dir_cap - 1);
  volatile int *gid = A.dir_gid;
  volatile uint64_t *keys = A.dir_keys;
  while (true) {
    const int g = gid[h];
    if (g >= 0) {
      if (keys[h] == key) return g;
      h = (h + 1) & (A.dir_cap - 1);
      continue;
    }
    if (g == -1 && atomicCAS(&A.dir_gid[h], -1, -2) == -1) {
      uint32_t id = atomicAdd(A.n_groups, 1u);
      if (id >= A.partial_rows) {
        atomicExch(A.error_flag, static_cast<uint32_t>(QSGPU_ERR_CAPACITY));
        id = A.partial_rows - 1;
      }
      keys[h] = key;
      A.gid_keys[id] = key;
      __threadfence();
      gid[h] = static_cast<int>(id);
      return static_cast<int>(id);
    }
  }
}

template <int HOT, int NAGG>
__global__ void __launch_bounds__(kBlock, 2)
k_scan_agg(const __grid_constant__ ScanDesc S, const __grid_constant__ Program P,
           const __grid_constant__ AggDesc A) {
  extern __shared__ __align__(128) char smem[];
  const int tid = threadIdx.x;
  const bool grouped = A.n_key_cols > 0;
  const uint32_t LG = grouped ? kCompactMaxGroups : 1;
  const uint32_t LS = grouped ? kCompactLocalSlots : 0;

  AggSmem M;
  {
    char *p = smem + kBarBytes + S.n_stages * S.stage_bytes;
    M.red = reinterpret_cast<uint64_t *>(p); p += 8 * HOT * (NAGG + 1) * 8;
    M.lstate = reinterpret_cast<uint64_t *>(p); p += LG * A.words * 8;
    M.lkey_by_id = reinterpret_cast<uint64_t *>(p); p += LG * 8;
    M.lkeys = reinterpret_cast<uint64_t *>(p); p += LS * 8;
    M.lslot = reinterpret_cast<int *>(p); p += LS * 4;
    M.nlocal = reinterpret_cast<uint32_t *>(p);
  }
  for (uint32_t i = tid; i < LG * A.words; i += kBlock) {
    const uint32_t w = i % A.words;
    M.lstate[i] = w == 0 ? 0 : agg_identity(A.kind[w - 1]);
  }
  for (uint32_t i = tid; i < LS; i += kBlock) M.lslot[i] = -1;
  if (tid == 0) {
    *M.nlocal = grouped ? 0u : 1u;
    if (!grouped) M.lkey_by_id[0] = 0;
  }
  // (scan_tiles starts with a __syncthreads)

  AggSink<HOT, NAGG> sink;
  sink.A = &A;
  sink.lstate = M.lstate;
#pragma unroll
  for (int g = 0; g < HOT; ++g) {
    sink.hc[g] = 0;
#pragma unroll
    for (int j = 0; j < NAGG; ++j) sink.hv[g][j] = agg_identity(j < static_cast<int>(A.n_agg) ? A.kind[j] : 0);
  }
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
      sink.slot[r] = pass ? 0 : -1;
      any |= pass;
    }
    if (!__any_sync(0xffffffffu, any)) return;
    if (grouped) {
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        if (sink.slot[r] < 0) continue;
        uint64_t key = 0;
        for (uint32_t k = 0; k < A.n_key_cols; ++k) {
          const uint32_t w = A.key_width[k];
          const char *src = stage + S.cols[A.key_col[k]].smem_off + tile_row(r, tid) * w;
          uint64_t v = 0;
          for (uint32_t b = 0; b < w; ++b) v |= static_cast<uint64_t>(static_cast<unsigned char>(src[b])) << (8 * b);
          key |= v << (8 * A.key_off[k]);
        }
        sink.slot[r] = local_lookup(key, M, LS, LG, A.error_flag);
      }
    }
    // row counts
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
#pragma unroll
      for (int g = 0; g < HOT; ++g) sink.hc[g] += (sink.slot[r] == g) ? 1u : 0u;
      if (sink.slot[r] >= HOT)
        atomicAdd(reinterpret_cast<unsigned long long *>(&M.lstate[sink.slot[r] * A.words]), 1ull);
    }
    vm_run(P, P.n_mid, P.n_total, S, stage, tid, regs, bits, sink);
  });

  // ---- CTA reduction of the register-resident (hot) groups, fixed tree.
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int g = 0; g < HOT; ++g) {
    uint64_t c = sink.hc[g];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
    if (lane == 0) M.red[(warp * HOT + g) * (NAGG + 1)] = c;
#pragma unroll
    for (int j = 0; j < NAGG; ++j) {
      if (j >= static_cast<int>(A.n_agg)) break;
      const uint8_t kind = A.kind[j];
      uint64_t x = sink.hv[g][j];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const uint64_t y = __shfl_xor_sync(0xffffffffu, x, off);
        // keep operand order lane-independent: lower lane first
        x = (lane & off) ? agg_combine(kind, y, x) : agg_combine(kind, x, y);
      }
      if (lane == 0) M.red[(warp * HOT + g) * (NAGG + 1) + 1 + j] = x;
    }
  }
  __syncthreads();
  const uint32_t nlocal = min(*M.nlocal, LG);
  if (tid < HOT * (NAGG + 1)) {
    const int g = tid / (NAGG + 1), w = tid % (NAGG + 1);
    if (static_cast<uint32_t>(g) < nlocal && static_cast<uint32_t>(w) < A.words) {
      uint64_t x = M.red[(0 * HOT + g) * (NAGG + 1) + w];
      for (int wp = 1; wp < kBlock / 32; ++wp) {
        const uint64_t y = M.red[(wp * HOT + g) * (NAGG + 1) + w];
        x = w == 0 ? x + y : agg_combine(A.kind[w - 1], x, y);
      }
      // lstate[g] holds identity (hot groups never touch it during the scan)
      M.lstate[g * A.words + w] = x;
    }
  }
  __syncthreads();
  // ---- publish this CTA's partial state, one row per group it met.
  for (uint32_t l = tid; l < nlocal; l += kBlock) {
    const int gid = dir_insert(M.lkey_by_id[l], A);
    uint64_t *dst = A.partials + (static_cast<uint64_t>(blockIdx.x) * A.partial_rows + gid) * A.words;
    for (uint32_t w = 0; w < A.words; ++w) dst[w] = M.lstate[l * A.words + w];
  }
}

// Fill [n_rows][words] state rows with identities (row count = 0).
__global__ void k_fill_identity(uint64_t *states, uint64_t n_rows, const __grid_constant__ AggDesc A) {
  const uint64_t n = n_rows * A.words;
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint32_t w = static_cast<uint32_t>(i % A.words);
    states[i] = w == 0 ? 0 : agg_identity(A.kind[w - 1]);
  }
}

// totals[g][w] (+)= fold over CTAs, in CTA order (deterministic).
__global__ void k_merge_partials(const __grid_constant__ AggDesc A, uint32_t n_ctas) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n_groups = min(*A.n_groups, A.partial_rows);
  if (i >= n_groups * A.words) return;
  const uint32_t g = i / A.words, w = i % A.words;
  uint64_t x = A.states[i];
  const uint8_t kind = w == 0 ? AK_SUM_I64 : A.kind[w - 1];
  const uint64_t ident = w == 0 ? 0 : agg_identity(kind);
  for (uint32_t c = 0; c < n_ctas; ++c) {
    uint64_t *p = &A.partials[(static_cast<uint64_t>(c) * A.partial_rows + g) * A.words + w];
    x = agg_combine(kind, x, *p);
    *p = ident;                     // consumed: ready for the next work order
  }
  A.states[i] = x;
}

// Merge a foreign partial (another GPU's totals) into this state, compact strategy.
__global__ void k_merge_foreign_compact(const __grid_constant__ AggDesc A, const uint64_t *f_states,
                                        const uint64_t *f_keys, uint32_t f_groups) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= f_groups) return;
  if (f_states[static_cast<uint64_t>(g) * A.words] == 0) return;   // empty group
  const int gid = dir_insert(f_keys[g], A);
  for (uint32_t w = 0; w < A.words; ++w) {
    const uint8_t kind = w == 0 ? AK_SUM_I64 : A.kind[w - 1];
    uint64_t *dst = &A.states[static_cast<uint64_t>(gid) * A.words + w];
    *dst = agg_combine(kind, *dst, f_states[static_cast<uint64_t>(g) * A.words + w]);
  }
}

// ------------------------------------------------------------------ launchers
template <int HOT, int NAGG>
static cudaError_t launch_one(const ScanDesc &S, const Program &P, const AggDesc &A, int grid, size_t smem,
                              cudaStream_t st) {
  auto kern = k_scan_agg<HOT, NAGG>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  kern<<<grid, kBlock, smem, st>>>(S, P, A);
  return cudaGetLastError();
}

size_t agg_smem_extra(int hot, int nagg_t, bool grouped, uint32_t words) {
  const size_t LG = grouped ? kCompactMaxGroups : 1, LS = grouped ? kCompactLocalSlots : 0;
  return static_cast<size_t>(8) * hot * (nagg_t + 1) * 8 + LG * words * 8 + LG * 8 + LS * 8 + LS * 4 + 16;
}

void agg_template_shape(const AggDesc &A, int *hot, int *nagg_t) {
  *hot = A.n_key_cols > 0 ? 4 : 1;
  *nagg_t = A.n_agg <= 1 ? 1 : A.n_agg <= 2 ? 2 : A.n_agg <= 4 ? 4 : A.n_agg <= 6 ? 6 : 8;
}

cudaError_t launch_scan_agg(const ScanDesc &S, const Program &P, const AggDesc &A, int grid, size_t smem,
                            cudaStream_t st) {
  int hot, nt;
  agg_template_shape(A, &hot, &nt);
#define QS_CASE(H, N) if (hot == H && nt == N) return launch_one<H, N>(S, P, A, grid, smem, st)
  QS_CASE(1, 1); QS_CASE(1, 2); QS_CASE(1, 4); QS_CASE(1, 6); QS_CASE(1, 8);
  QS_CASE(4, 1); QS_CASE(4, 2); QS_CASE(4, 4); QS_CASE(4, 6); QS_CASE(4, 8);
#undef QS_CASE
  return cudaErrorInvalidValue;
}

cudaError_t launch_fill_identity(uint64_t *states, uint64_t n_rows, const AggDesc &A, cudaStream_t st) {
  const uint64_t n = n_rows * A.words;
  int grid = static_cast<int>((n + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  if (grid < 1) grid = 1;
  k_fill_identity<<<grid, 256, 0, st>>>(states, n_rows, A);
  return cudaGetLastError();
}

cudaError_t launch_merge_partials(const AggDesc &A, uint32_t n_ctas, cudaStream_t st) {
  const uint32_t n = A.partial_rows * A.words;
  k_merge_partials<<<(n + 127) / 128, 128, 0, st>>>(A, n_ctas);
  return cudaGetLastError();
}

cudaError_t launch_merge_foreign_compact(const AggDesc &A, const uint64_t *f_states, const uint64_t *f_keys,
                                         uint32_t f_groups, cudaStream_t st) {
  if (f_groups == 0) return cudaSuccess;
  // one thread block, sequential enough to be deterministic per group
  k_merge_foreign_compact<<<(f_groups + 127) / 128, 128, 0, st>>>(A, f_states, f_keys, f_groups);
  return cudaGetLastError();
}

}  // namespace qs
