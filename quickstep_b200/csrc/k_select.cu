// K3 + K4-build: predicate scan, LIP probe/insert and order-preserving
// ballot compaction into an output relation.
//
//   SelectWorkOrder::execute         relational_operators/SelectOperator.cpp:161-195
//     getMatchesForPredicate         storage/StorageBlock.cpp:1053-1081
//     LIPFilterAdaptiveProber        utility/lip_filter/LIPFilterAdaptiveProber.hpp:89-232
//     select / selectSimple          storage/StorageBlock.cpp:363-398
//     bulkInsertTuples               storage/InsertDestination.cpp:202-216
//   BuildLIPFilterWorkOrder::execute relational_operators/BuildLIPFilterOperator.cpp:146-172
//
// The reference builds a TupleIdSequence bitmap, then one ColumnVector per
// projected expression, then copies tuples into the destination block.  Here
// the bitmap is a warp ballot: each warp counts its survivors, one atomicAdd
// per CTA tile reserves the output range, and survivors are written straight
// from the staged tile to their final position (rows of a tile keep their
// input order).
#include "qs_compact.cuh"
#include "qs_ops.cuh"
#include "qs_vm.cuh"

namespace qs {

struct SelectSink : SinkBase {
  const SinkDesc *K;
  uint64_t idx[kRows];     // output row, ~0 when the row does not survive
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
};

__global__ void __launch_bounds__(kBlock, 2)
k_scan_select(const __grid_constant__ ScanDesc S, const __grid_constant__ Program P,
              const __grid_constant__ SinkDesc K) {
  extern __shared__ __align__(128) char smem[];
  const int tid = threadIdx.x;
  uint32_t *s_compact = reinterpret_cast<uint32_t *>(smem + kBarBytes + S.n_stages * S.stage_bytes);
  SelectSink sink;
  sink.K = &K;
  sink.tid = tid;
  VmRegs regs;
  scan_tiles(S, smem, [&](uint32_t tile, const char *stage, const ScanRt &rt) {
    bool valid[kRows];
    tile_valid(S, rt, tile, tid, valid);
    uint32_t bits[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) bits[r] = 1u;
    SinkBase ns;
    vm_run(P, 0, P.n_pred, S, stage, tid, regs, bits, ns);
    bool pass[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) pass[r] = valid[r] && (bits[r] & 1u);

    // LIPFilterBuilder::insertValueAccessor on the survivors.
    for (uint32_t f = 0; f < K.n_lip_build; ++f) {
      const char *base = stage + S.cols[K.lip_build_col[f]].smem_off;
      const uint8_t lt = K.lip_build_ltype[f];
      const uint32_t w = native_width(lt);
#pragma unroll
      for (int r = 0; r < kRows; ++r)
        if (pass[r]) lip_insert(K.lip_build[f], static_cast<int64_t>(load_native(base + tile_row(r, tid) * w, lt)));
    }
    if (K.n_out == 0) return;      // BuildLIPFilter: nothing to materialise (uniform per launch)

    cta_compact(pass, s_compact, K.counter, K.capacity, K.error_flag, sink.idx);
    vm_run(P, P.n_mid, P.n_total, S, stage, tid, regs, bits, sink);
  });
}

cudaError_t launch_scan_select(const ScanDesc &S, const Program &P, const SinkDesc &K, int grid, size_t smem,
                               cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(k_scan_select, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  k_scan_select<<<grid, kBlock, smem, st>>>(S, P, K);
  return cudaGetLastError();
}

}  // namespace qs
