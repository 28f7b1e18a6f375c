// CTA-wide, order-preserving stream compaction: warp ballots -> one warp scan
// over the kRows*8 per-warp counts -> one atomicAdd per CTA per tile.
// The device analogue of building a TupleIdSequence and iterating its set bits
// (storage/TupleIdSequence.hpp:121, utility/BitVector.hpp:625-640).
#pragma once

#include "qs_common.cuh"

namespace qs {



// Must be called by all threads of the CTA (contains __syncthreads).
__device__ __forceinline__ void cta_compact(const bool (&flag)[kRows], uint32_t *s,
                                            unsigned long long *counter, uint64_t capacity,
                                            uint32_t *error_flag, uint64_t (&idx)[kRows]) {
  constexpr int kCounts = kRows * (kBlock / 32);
  static_assert(kCounts <= 32, "one warp scans the per-warp counts");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t ball[kRows];
  // No barrier before `s` is written: every caller has a CTA-wide barrier between two compactions anyway -- the
  // end-of-tile __syncthreads of scan_tiles, and the __syncthreads_or that opens every round of the join probe -- so the
  // threads that read `s` in the previous compaction are past it.  (It used to be the first of three barriers here;
  // ncu had `barrier` as the top stall of Q3's orders select.)
#pragma unroll
  for (int r = 0; r < kRows; ++r) {
    ball[r] = __ballot_sync(0xffffffffu, flag[r]);
    if (lane == 0) s[r * (kBlock / 32) + warp] = __popc(ball[r]);
  }
  __syncthreads();
  if (warp == 0) {
    const uint32_t v = lane < kCounts ? s[lane] : 0u;
    uint32_t inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc, off);
      if (lane >= off) inc += t;
    }
    if (lane < kCounts) s[lane] = inc - v;       // exclusive prefix in (r, warp) order
    if (lane == 31) {
      unsigned long long base = 0;
      uint32_t overflow = 0;
      if (inc) {
        base = atomicAdd(counter, static_cast<unsigned long long>(inc));
        if (base + inc > capacity) {
          overflow = 1;
          atomicExch(error_flag, static_cast<uint32_t>(QSGPU_ERR_CAPACITY));
        }
      }
      s[32] = static_cast<uint32_t>(base);
      s[33] = static_cast<uint32_t>(base >> 32);
      s[34] = overflow;
    }
  }
  __syncthreads();
  const unsigned long long base = static_cast<unsigned long long>(s[32]) | (static_cast<unsigned long long>(s[33]) << 32);
  const bool overflow = s[34] != 0;
#pragma unroll
  for (int r = 0; r < kRows; ++r) {
    idx[r] = (flag[r] && !overflow)
                 ? base + s[r * (kBlock / 32) + warp] + __popc(ball[r] & ((1u << lane) - 1u))
                 : ~0ull;
  }
}

// Warp-wide form (NOT the default: measured slower).  Every warp reserves the output range of its own survivors with one
// atomicAdd and places them in (lane-row, lane) order: no CTA barrier, no shared memory, a warp that waits for its
// reservation stalls alone.  The idea came from ncu of Q3's orders select (`barrier` the top stall, 5.7 cycles per issued
// instruction, 64 % of the HBM peak).  Measured through the operator layer at SF10 (profiles/README.md, r3a): Q3 0.806 ms
// with this form against 0.772 ms with cta_compact, Q1 / Q6 unchanged.  What it costs outweighs the barriers it removes:
// up to 8x the atomics on the relation's row counter, and output runs of a warp's few survivors (3 % of the lineitem rows
// pass: ~4 per warp) instead of a tile's ~30 -- more partially written sectors in the output relation.  Kept behind
// QSGPU_WARP_COMPACT=1 so the comparison can be repeated.
__device__ __forceinline__ void warp_compact(const bool (&flag)[kRows], unsigned long long *counter, uint64_t capacity,
                                             uint32_t *error_flag, uint64_t (&idx)[kRows]) {
  const int lane = threadIdx.x & 31;
  uint32_t ball[kRows], before[kRows], total = 0;
#pragma unroll
  for (int r = 0; r < kRows; ++r) {
    ball[r] = __ballot_sync(0xffffffffu, flag[r]);
    before[r] = total;
    total += __popc(ball[r]);
  }
  unsigned long long base = 0;
  int overflow = 0;
  if (total != 0) {                                  // warp-uniform
    if (lane == 0) {
      base = atomicAdd(counter, static_cast<unsigned long long>(total));
      if (base + total > capacity) {
        overflow = 1;
        atomicExch(error_flag, static_cast<uint32_t>(QSGPU_ERR_CAPACITY));
      }
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    overflow = __shfl_sync(0xffffffffu, overflow, 0);
  }
#pragma unroll
  for (int r = 0; r < kRows; ++r)
    idx[r] = (flag[r] && !overflow) ? base + before[r] + __popc(ball[r] & ((1u << lane) - 1u)) : ~0ull;
}

}  // namespace qs
