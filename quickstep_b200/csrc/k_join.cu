// K5 / K6: hash-join build and probe.
//
//   BuildHashWorkOrder::execute      relational_operators/BuildHashOperator.cpp:162-207
//     HashTable::putValueAccessor    storage/HashTable.hpp:1365-1461
//   HashInnerJoinWorkOrder::execute  relational_operators/HashJoinOperator.cpp:450-671
//     getAllFromValueAccessor        storage/HashTable.hpp:2155-2181
//     residual predicate per pair    HashJoinOperator.cpp:511-525
//     Scalar::getAllValuesForJoin    HashJoinOperator.cpp:527-536
//   HashSemiJoin / HashAntiJoin      HashJoinOperator.cpp:673-987
//
// The reference's JoinHashTable is a separate-chaining multi-map from key to
// TupleReference{block, tuple}; probes collect (probe_tid, build_tid) pairs per
// build block and re-open every build block.  On the device the table is one
// open-addressing array of 16-byte {key, build row} slots (linear probing,
// duplicates occupy their own slots), the probe keys arrive as TMA-staged
// tiles, and matched pairs are projected in the same kernel: build-side
// operands are gathered through the stored row id.  Rows that share a probe
// tile advance in lock-step "rounds" (one match per row per round) so that the
// warp-ballot compaction and the VM stay CTA-uniform even with duplicate keys.
#include "qs_compact.cuh"
#include "qs_ops.cuh"
#include "qs_vm.cuh"

namespace qs {

constexpr unsigned long long kEmptyRow = ~0ull;

__global__ void k_join_clear(JoinSlot *slots, uint64_t cap) {
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < cap;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    slots[i].key = 0;
    slots[i].row = kEmptyRow;
  }
}

__global__ void __launch_bounds__(kBlock, 2)
k_join_build(const __grid_constant__ ScanDesc S, const __grid_constant__ Program P,
             const __grid_constant__ SinkDesc K, const __grid_constant__ JoinDesc J) {
  extern __shared__ __align__(128) char smem[];
  const int tid = threadIdx.x;
  VmRegs regs;
  const uint64_t mask = J.cap - 1;
  scan_tiles(S, smem, [&](uint32_t tile, const char *stage, const ScanRt &rt) {
    bool valid[kRows];
    tile_valid(S, rt, tile, tid, valid);
    uint32_t bits[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) bits[r] = 1u;
    SinkBase ns;
    vm_run(P, 0, P.n_pred, S, stage, tid, regs, bits, ns);
    const uint64_t row0 = S.first_row + static_cast<uint64_t>(tile) * kTileRows;
    for (uint32_t f = 0; f < K.n_lip_build; ++f) {
      const char *base = stage + S.cols[K.lip_build_col[f]].smem_off;
      const uint8_t lt = K.lip_build_ltype[f];
      const uint32_t w = native_width(lt);
#pragma unroll
      for (int r = 0; r < kRows; ++r)
        if (valid[r] && (bits[r] & 1u))
          lip_insert(K.lip_build[f], static_cast<int64_t>(load_native(base + tile_row(r, tid) * w, lt)));
    }
    const char *kbase = stage + S.cols[J.key_col].smem_off;
    const uint32_t kw = native_width(J.key_ltype);
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      if (!(valid[r] && (bits[r] & 1u))) continue;
      const int64_t key = static_cast<int64_t>(load_native(kbase + tile_row(r, tid) * kw, J.key_ltype));
      const unsigned long long row = row0 + tile_row(r, tid);
      uint64_t h = mix64(static_cast<uint64_t>(key)) & mask;
      bool done = false;
      for (uint64_t probes = 0; probes <= mask; ++probes) {
        if (atomicCAS(&J.slots[h].row, kEmptyRow, row) == kEmptyRow) {
          J.slots[h].key = key;
          done = true;
          break;
        }
        h = (h + 1) & mask;
      }
      if (done) atomicAdd(J.n_entries, 1ull);
      else atomicExch(J.error_flag, static_cast<uint32_t>(QSGPU_ERR_CAPACITY));
    }
  });
}

struct JoinSink : SinkBase {
  const SinkDesc *K;
  const JoinDesc *J;
  uint64_t idx[kRows];
  unsigned long long brow[kRows];
  int tid;
  __device__ __forceinline__ void emit(uint32_t j, uint8_t, const uint64_t (&acc)[kRows]) {
    const uint32_t w = K->out_width[j];
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      if (idx[r] == ~0ull) continue;
      if (w == 4) *reinterpret_cast<uint32_t *>(K->out[j] + idx[r] * 4) = static_cast<uint32_t>(acc[r]);
      else *reinterpret_cast<uint64_t *>(K->out[j] + idx[r] * 8) = acc[r];
    }
  }
  __device__ __forceinline__ void emit_raw(uint32_t j, const char *col, uint32_t w) {
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      if (idx[r] == ~0ull) continue;
      const char *src = col + tile_row(r, tid) * w;
      char *dst = K->out[j] + idx[r] * w;
      if (w == 8) *reinterpret_cast<uint64_t *>(dst) = *reinterpret_cast<const uint64_t *>(src);
      else if (w == 4) *reinterpret_cast<uint32_t *>(dst) = *reinterpret_cast<const uint32_t *>(src);
      else for (uint32_t b = 0; b < w; ++b) dst[b] = src[b];
    }
  }
  __device__ __forceinline__ void emit_raw_build(uint32_t j, uint32_t col) {
    const uint32_t w = J->build_cols[col].width;
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      if (idx[r] == ~0ull || brow[r] == kEmptyRow) continue;
      const char *src = J->build_cols[col].ptr + brow[r] * w;
      char *dst = K->out[j] + idx[r] * w;
      if (w == 8) *reinterpret_cast<uint64_t *>(dst) = *reinterpret_cast<const uint64_t *>(src);
      else if (w == 4) *reinterpret_cast<uint32_t *>(dst) = *reinterpret_cast<const uint32_t *>(src);
      else for (uint32_t b = 0; b < w; ++b) dst[b] = src[b];
    }
  }
  __device__ __forceinline__ uint64_t build_leaf(uint32_t col, uint8_t ltype, int r) {
    if (brow[r] == kEmptyRow) return 0;
    return load_native(J->build_cols[col].ptr + brow[r] * J->build_cols[col].width, ltype);
  }
};

__global__ void __launch_bounds__(kBlock, 2)
k_join_probe(const __grid_constant__ ScanDesc S, const __grid_constant__ Program P,
             const __grid_constant__ SinkDesc K, const __grid_constant__ JoinDesc J) {
  extern __shared__ __align__(128) char smem[];
  const int tid = threadIdx.x;
  uint32_t *s_compact = reinterpret_cast<uint32_t *>(smem + kBarBytes + S.n_stages * S.stage_bytes);
  JoinSink sink;
  sink.K = &K;
  sink.J = &J;
  sink.tid = tid;
  VmRegs regs;
  const uint64_t mask = J.cap - 1;
  const bool has_residual = P.n_mid > P.n_pred;
  const bool inner = J.join_type == QS_JOIN_INNER;

  scan_tiles(S, smem, [&](uint32_t tile, const char *stage, const ScanRt &rt) {
    bool valid[kRows];
    tile_valid(S, rt, tile, tid, valid);
    uint32_t bits[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) bits[r] = 1u;
    SinkBase ns;
    vm_run(P, 0, P.n_pred, S, stage, tid, regs, bits, ns);

    bool pass[kRows], active[kRows], matched[kRows];
    int64_t key[kRows];
    uint64_t h[kRows];
    const char *kbase = stage + S.cols[J.key_col].smem_off;
    const uint32_t kw = native_width(J.key_ltype);
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      pass[r] = valid[r] && (bits[r] & 1u);
      active[r] = pass[r];
      matched[r] = false;
      key[r] = static_cast<int64_t>(load_native(kbase + tile_row(r, tid) * kw, J.key_ltype));
      h[r] = mix64(static_cast<uint64_t>(key[r])) & mask;
    }

    while (true) {
      bool found[kRows];
      bool any = false;
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        found[r] = false;
        sink.brow[r] = kEmptyRow;
        if (!active[r]) continue;
        for (uint64_t probes = 0; probes <= mask; ++probes) {
          const ulonglong2 s = *reinterpret_cast<const ulonglong2 *>(&J.slots[h[r]]);
          if (s.y == kEmptyRow) { active[r] = false; break; }
          h[r] = (h[r] + 1) & mask;
          if (static_cast<int64_t>(s.x) == key[r]) { found[r] = true; sink.brow[r] = s.y; break; }
        }
        if (!found[r]) active[r] = false;
        any |= found[r];
      }
      if (!__syncthreads_or(any)) break;
      bool ok[kRows];
      if (has_residual) {
        uint32_t rb[kRows];
#pragma unroll
        for (int r = 0; r < kRows; ++r) rb[r] = 1u;
        vm_run(P, P.n_pred, P.n_mid, S, stage, tid, regs, rb, sink);
#pragma unroll
        for (int r = 0; r < kRows; ++r) ok[r] = found[r] && (rb[r] & 1u);
      } else {
#pragma unroll
        for (int r = 0; r < kRows; ++r) ok[r] = found[r];
      }
      if (inner) {
        cta_compact(ok, s_compact, K.counter, K.capacity, K.error_flag, sink.idx);
        vm_run(P, P.n_mid, P.n_total, S, stage, tid, regs, bits, sink);
      } else {
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
          if (ok[r]) { matched[r] = true; active[r] = false; }   // existence is enough
        }
      }
    }
    if (!inner) {
      bool flag[kRows];
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        flag[r] = pass[r] && (J.join_type == QS_JOIN_LEFT_SEMI ? matched[r] : !matched[r]);
        sink.brow[r] = kEmptyRow;
      }
      cta_compact(flag, s_compact, K.counter, K.capacity, K.error_flag, sink.idx);
      vm_run(P, P.n_mid, P.n_total, S, stage, tid, regs, bits, sink);
    }
  });
}

// ------------------------------------------------------------------ launchers
cudaError_t launch_join_clear(const JoinDesc &J, cudaStream_t st) {
  uint64_t g = (J.cap + 255) / 256;
  if (g > 148ull * 16) g = 148ull * 16;
  k_join_clear<<<static_cast<unsigned>(g), 256, 0, st>>>(J.slots, J.cap);
  return cudaGetLastError();
}

cudaError_t launch_join_build(const ScanDesc &S, const Program &P, const SinkDesc &K, const JoinDesc &J,
                              int grid, size_t smem, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(k_join_build, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  k_join_build<<<grid, kBlock, smem, st>>>(S, P, K, J);
  return cudaGetLastError();
}

cudaError_t launch_join_probe(const ScanDesc &S, const Program &P, const SinkDesc &K, const JoinDesc &J,
                              int grid, size_t smem, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(k_join_probe, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  k_join_probe<<<grid, kBlock, smem, st>>>(S, P, K, J);
  return cudaGetLastError();
}

}  // namespace qs
