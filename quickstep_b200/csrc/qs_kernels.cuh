// The five scan-kernel families of libqsgpu, as templates over a compile-time
// query description Q (printed by the query compiler, qs_jit.cu):
//
//   scan_agg_body      K1 / K2  predicate scan + aggregation, no / compact GROUP BY
//                      AggregationOperationState::aggregateBlockSingleState
//                        (storage/AggregationOperationState.cpp:476-519)
//                      ThreadPrivateCompactKeyHashTable::upsertValueAccessorCompositeKey
//                        (storage/ThreadPrivateCompactKeyHashTable.cpp:203-363)
//   scan_groupby_body  K7       hash / dense GROUP BY over a global table
//                      PackedPayloadHashTable (storage/PackedPayloadHashTable.hpp:780-909)
//                      CollisionFreeVectorTable (storage/CollisionFreeVectorTable.hpp:530-645)
//   scan_select_body   K3 / K4  Select and BuildLIPFilter
//                      SelectWorkOrder::execute (relational_operators/SelectOperator.cpp:161-195)
//                      BuildLIPFilterWorkOrder::execute (…/BuildLIPFilterOperator.cpp:146-172)
//   join_build_body    K5       BuildHashWorkOrder::execute (…/BuildHashOperator.cpp:162-207)
//   join_probe_body    K6       HashInnerJoin / Semi / Anti (…/HashJoinOperator.cpp:450-987)
//
// Common skeleton: one persistent CTA pair per SM; column tiles arrive in
// shared memory through the TMA bulk-copy ring (qs_vm.cuh scan_tiles); the
// compile-time VM evaluates predicates and expressions in registers.
//
// Q provides (all static constexpr):
//   n_cols n_stages stage_bytes col_w(c) col_off(c)        staged columns / ring
//   n_pred n_mid n_total code(pc)                          the program
//   lip_kind(i) lip_anti(i)                                LIP filters probed
//   n_agg words hot grouped strategy n_key_cols key_words  aggregation
//   agg_kind(j) key_col(k) key_w(k) key_off(k)
//   n_out out_w(j) n_lip_build lb_col(i) lb_ltype(i) lb_kind(i)   output side
//   j_key_col j_key_ltype j_type build_w(c) build_cw(c)    join
#pragma once

#include "qs_compact.cuh"
#include "qs_ops.cuh"
#include "qs_vm.cuh"

namespace qs {

// ------------------------------------------------------------ small helpers
// 128-bit compare-and-swap (ATOMG.CAS.128, sm_90+) on a 16-byte aligned pair of words: (o0, o1) = old value; the
// pair becomes (v0, v1) iff it was (c0, c1).
__device__ __forceinline__ void cas128(void *addr, uint64_t c0, uint64_t c1, uint64_t v0, uint64_t v1, uint64_t &o0, uint64_t &o1) {
  asm volatile(
      "{\n.reg .b128 c, v, o;\nmov.b128 c, {%2, %3};\nmov.b128 v, {%4, %5};\n"
      "atom.global.cas.b128 o, [%6], c, v;\nmov.b128 {%0, %1}, o;\n}"
      : "=l"(o0), "=l"(o1)
      : "l"(c0), "l"(c1), "l"(v0), "l"(v1), "l"(addr)
      : "memory");
}

template <uint32_t W>
__device__ __forceinline__ uint64_t load_bytes(const char *p) {
  if constexpr (W == 1) return static_cast<uint64_t>(*reinterpret_cast<const unsigned char *>(p));
  else if constexpr (W == 2) return static_cast<uint64_t>(*reinterpret_cast<const uint16_t *>(p));
  else if constexpr (W == 4) return static_cast<uint64_t>(*reinterpret_cast<const uint32_t *>(p));
  else if constexpr (W == 8) return *reinterpret_cast<const uint64_t *>(p);
  else {
    static_assert(W < 8, "wide values are copied, not loaded");
    uint64_t v = 0;
#pragma unroll
    for (uint32_t b = 0; b < W; ++b) v |= static_cast<uint64_t>(static_cast<unsigned char>(p[b])) << (8 * b);
    return v;
  }
}

template <uint32_t W>
__device__ __forceinline__ void copy_value(char *dst, const char *src) {
  if constexpr (W == 8) *reinterpret_cast<uint64_t *>(dst) = *reinterpret_cast<const uint64_t *>(src);
  else if constexpr (W == 4) *reinterpret_cast<uint32_t *>(dst) = *reinterpret_cast<const uint32_t *>(src);
  else if constexpr (W == 2) *reinterpret_cast<uint16_t *>(dst) = *reinterpret_cast<const uint16_t *>(src);
  else {
#pragma unroll
    for (uint32_t b = 0; b < W; ++b) dst[b] = src[b];
  }
}

// Group key of one row, packed like ThreadPrivateCompactKeyHashTable::ConstructKeyCode
// (storage/ThreadPrivateCompactKeyHashTable.hpp:125-142): key i is memcpy'd at
// byte offset sum(widths of keys < i) of a zeroed word array.
template <class Q>
__device__ __forceinline__ void pack_key(const char *stage, uint32_t row, uint64_t (&key)[kMaxKeyWords]) {
#pragma unroll
  for (int i = 0; i < kMaxKeyWords; ++i) key[i] = 0;
  static_for<0, Q::n_key_cols>([&](auto kk) {
    constexpr int k = QS_IDX(kk);
    constexpr uint32_t w = Q::key_w(k), off = Q::key_off(k);
    const char *src = stage + Q::col_off(Q::key_col(k)) + row * w;
    if constexpr (w <= 8 && (w == 1 || w == 2 || w == 4 || w == 8 || (off & 7) + w <= 8)) {
      const uint64_t v = load_bytes<w>(src);
      constexpr uint32_t sh = 8 * (off & 7);
      key[off >> 3] |= v << sh;
      if constexpr ((off & 7) + w > 8) key[(off >> 3) + 1] |= v >> (64 - sh);
    } else {
#pragma unroll
      for (uint32_t b = 0; b < w; ++b) {
        constexpr uint32_t base = off;
        const uint32_t pos = base + b;
        key[pos >> 3] |= static_cast<uint64_t>(static_cast<unsigned char>(src[b])) << (8 * (pos & 7));
      }
    }
  });
}

template <uint8_t KIND>
__device__ __forceinline__ void atomic_update(uint64_t *p, uint64_t v) {
  if constexpr (KIND == AK_SUM_F64) atomicAdd(reinterpret_cast<double *>(p), u2d(v));
  else if constexpr (KIND == AK_SUM_I64) atomicAdd(reinterpret_cast<unsigned long long *>(p), static_cast<unsigned long long>(v));
  else if constexpr (KIND == AK_MIN_I64) atomicMin(reinterpret_cast<long long *>(p), static_cast<long long>(v));
  else if constexpr (KIND == AK_MAX_I64) atomicMax(reinterpret_cast<long long *>(p), static_cast<long long>(v));
  else {   // MIN/MAX over doubles: CAS loop
    unsigned long long *q = reinterpret_cast<unsigned long long *>(p);
    unsigned long long old = *q;
    while (true) {
      const uint64_t want = agg_combine(KIND, old, v);
      if (want == old) break;
      const unsigned long long seen = atomicCAS(q, old, static_cast<unsigned long long>(want));
      if (seen == old) break;
      old = seen;
    }
  }
}

// =========================================================== K1 / K2  scan_agg
struct AggSmem {
  uint64_t *red;         // [8 warps][HOT*(NA+1)]
  uint64_t *lstate;      // [LG][words]
  uint64_t *lkey_by_id;  // [LG]
  uint64_t *lkeys;       // [LS]
  int *lslot;            // [LS]
  uint32_t *lready;      // [LG]  1 once lkey_by_id[id] is published
  uint32_t *nlocal;
  uint64_t *priv;        // Q::priv: [HOT][NA][kBlock] per-thread value accumulators of the hot groups
};

// hv (+)= v when slot == G, as ONE predicated instruction (ptxas shares the
// setp between the aggregates of a row).  Written in PTX because the compiler
// turns the equivalent if-chain over the hot groups into a jump table
// (BRX + BSSY/BSYNC per row; r01c profile: 30 % of all issued instructions).
template <uint8_t KIND, int G>
__device__ __forceinline__ void hot_update(uint64_t &h, uint64_t v, int slot) {
  if constexpr (KIND == AK_SUM_F64) {
    double x = u2d(h);
    asm("{\n.reg .pred p;\nsetp.eq.s32 p, %1, %2;\n@p add.rn.f64 %0, %0, %3;\n}" : "+d"(x) : "r"(slot), "n"(G), "d"(u2d(v)));
    h = d2u(x);
  } else if constexpr (KIND == AK_SUM_I64) {
    asm("{\n.reg .pred p;\nsetp.eq.s32 p, %1, %2;\n@p add.s64 %0, %0, %3;\n}" : "+l"(h) : "r"(slot), "n"(G), "l"(v));
  } else {
    h = agg_combine(KIND, h, slot == G ? v : agg_identity(KIND));
  }
}
template <int G>
__device__ __forceinline__ void hot_count(uint32_t &c, int slot) {
  asm("{\n.reg .pred p;\nsetp.eq.s32 p, %1, %2;\n@p add.u32 %0, %0, 1;\n}" : "+r"(c) : "r"(slot), "n"(G));
}

template <class Q>
struct AggSink : SinkBase {
  static constexpr int HOT = Q::hot;
  static constexpr int NA = Q::n_agg > 0 ? Q::n_agg : 1;
  uint64_t hv[HOT][NA];
  uint32_t hc[HOT];
  int slot[kRows];
  // High word of the double 1.0 when row r belongs to hot group g, else 0 (low word is always 0): the
  // per-row group selection of a double SUM becomes ONE DFMA per (row, group), hv += v * {1.0 | 0.0},
  // instead of DADD + two FSEL + a predicate unpack (r01c SASS: 60 of 183 instructions per row).
  // v * 1.0 + h rounds once, exactly like h + v; v * 0.0 + h == h for every finite v.
  uint32_t mh[kRows][HOT];
  bool cold;            // warp-uniform: some row of this warp's tile slice is in a non-hot group
  uint64_t *lstate;
  // Q::priv form: the hot groups' value accumulators are per-thread slots in SHARED memory, addressed by the row's
  // group id -- LDS + op + STS per (row, aggregate) instead of one masked DFMA (plus a mask move) per (row,
  // aggregate, hot group), and ~55 registers fewer.  prow[r] = this thread's slot of row r's group (value word 0);
  // a row that fails the predicate or belongs to a cold group computes against the last hot group's slot and its
  // store is predicated off.  Slot (g, j, tid) sits at ((g * NA + j) * kBlock + tid) * 8: the 32 lanes of a warp
  // always touch 32 consecutive 8-byte words, whatever their groups.  Row counts stay in registers (hc).
  // The values of a tile's rows are first parked in registers (pv) and applied after the emit section, row by row
  // with the aggregates of one row side by side: the NA slots of a row are distinct addresses, so their loads issue
  // back to back and the dependent chain per tile is kRows read-modify-writes long instead of kRows x NA (rows of
  // one thread may share a group, i.e. a slot, so rows stay ordered; the order of additions per slot is unchanged).
  char *prow[kRows];
  bool pok[kRows];
  uint64_t pv[NA][kRows];
  __device__ __forceinline__ void flush_private() {
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      uint64_t x[NA];
      static_for<0, Q::n_agg>([&](auto jj) {
        constexpr int j = QS_IDX(jj);
        x[j] = *reinterpret_cast<const uint64_t *>(prow[r] + j * kBlock * 8);
      });
      static_for<0, Q::n_agg>([&](auto jj) {
        constexpr int j = QS_IDX(jj);
        x[j] = agg_combine(Q::agg_kind(j), x[j], pv[j][r]);
      });
      if (pok[r]) {
        static_for<0, Q::n_agg>([&](auto jj) {
          constexpr int j = QS_IDX(jj);
          *reinterpret_cast<uint64_t *>(prow[r] + j * kBlock * 8) = x[j];
        });
      }
    }
    if (cold) {
#pragma unroll
      for (int r = 0; r < kRows; ++r)
        if (slot[r] >= HOT)
          static_for<0, Q::n_agg>([&](auto jj) {
            constexpr int j = QS_IDX(jj);
            atomic_update<Q::agg_kind(j)>(&lstate[slot[r] * Q::words + 1 + j], pv[j][r]);
          });
    }
  }
  __device__ __forceinline__ void set_private_rows(uint64_t *priv, int tid) {
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      pok[r] = static_cast<uint32_t>(slot[r]) < static_cast<uint32_t>(HOT);
      const uint32_t g = pok[r] ? static_cast<uint32_t>(slot[r]) : static_cast<uint32_t>(HOT - 1);
      prow[r] = reinterpret_cast<char *>(priv + (g * NA) * kBlock + tid);
    }
  }

  __device__ __forceinline__ void set_masks() {
#pragma unroll
    for (int r = 0; r < kRows; ++r)
#pragma unroll
      for (int g = 0; g < HOT; ++g) mh[r][g] = slot[r] == g ? 0x3ff00000u : 0u;
  }

  template <int J, int TYPE>
  __device__ __forceinline__ void emit(const uint64_t (&acc)[kRows]) {
    constexpr uint8_t kind = Q::agg_kind(J);
    if constexpr (Q::priv != 0) {
#pragma unroll
      for (int r = 0; r < kRows; ++r) pv[J][r] = acc[r];
      return;
    }
    bool fast = false;
    if constexpr (kind == AK_SUM_F64 && (HOT > 1)) {
      // inf * 0 and NaN * 0 are NaN: a non-finite value must not reach the other groups' accumulators, so
      // a warp that sees one (never, in TPC-H data) takes the predicated-add form below for this aggregate.
      bool fin = true;
#pragma unroll
      for (int r = 0; r < kRows; ++r) fin &= fabs(u2d(acc[r])) < __longlong_as_double(0x7ff0000000000000ll);
      fast = __all_sync(0xffffffffu, fin);
      if (fast) {
#pragma unroll
        for (int r = 0; r < kRows; ++r)
#pragma unroll
          for (int g = 0; g < HOT; ++g)
            hv[g][J] = d2u(fma(u2d(acc[r]), __hiloint2double(static_cast<int>(mh[r][g]), 0), u2d(hv[g][J])));
      }
    }
    if (!fast) {
#pragma unroll
      for (int r = 0; r < kRows; ++r)
        static_for<0, HOT>([&](auto gg) { hot_update<kind, QS_IDX(gg)>(hv[QS_IDX(gg)][J], acc[r], slot[r]); });
    }
    if constexpr (Q::grouped) {
      if (cold) {
#pragma unroll
        for (int r = 0; r < kRows; ++r)
          if (slot[r] >= HOT) atomic_update<kind>(&lstate[slot[r] * Q::words + 1 + J], acc[r]);
      }
    }
  }
};

// Per-CTA key -> local group id table (shared memory, open addressing).  Ids
// are handed out in arrival order; ids < HOT live in registers.  At most LG (= LS / 2) keys are ever
// installed, so the probe always reaches an empty slot; a key that arrives when all LG ids are taken is NOT
// installed: the row is dropped (-1) and QSGPU_ERR_CAPACITY is raised (the call fails at its next sync).
__device__ __forceinline__ int local_lookup(uint64_t key, const AggSmem &M, uint32_t LS, uint32_t LG,
                                            uint32_t *error_flag) {
  uint32_t h = static_cast<uint32_t>(mix64(key)) & (LS - 1);
  volatile int *lslot = M.lslot;
  volatile uint64_t *lkeys = M.lkeys;
  for (uint32_t moved = 0; moved < LS;) {
    const int s = lslot[h];
    if (s >= 0) {
      __threadfence_block();                 // acquire side of the publication below: the id first, then the key it guards
      if (lkeys[h] == key) return s;
      h = (h + 1) & (LS - 1);
      ++moved;
      continue;
    }
    if (s == -1 && atomicCAS(&M.lslot[h], -1, -2) == -1) {
      const uint32_t id = atomicAdd(M.nlocal, 1u);
      if (id >= LG) {
        atomicExch(error_flag, static_cast<uint32_t>(QSGPU_ERR_CAPACITY));
        lslot[h] = -1;                       // release the slot: nothing was installed
        return -1;
      }
      M.lkey_by_id[id] = key;
      lkeys[h] = key;
      __threadfence_block();
      *reinterpret_cast<volatile uint32_t *>(&M.lready[id]) = 1u;
      lslot[h] = static_cast<int>(id);
      return static_cast<int>(id);
    }
    // busy (another thread is publishing this slot): look again
  }
  atomicExch(error_flag, static_cast<uint32_t>(QSGPU_ERR_CAPACITY));
  return -1;
}

// Global key -> dense group id directory (persists across work orders).  Same rule: at most partial_rows
// (<= dir_cap / 4) keys are installed; one group too many raises QSGPU_ERR_CAPACITY and returns -1.
__device__ __forceinline__ int dir_insert(uint64_t key, const AggDesc &A) {
  uint32_t h = static_cast<uint32_t>(mix64(key)) & (A.dir_cap - 1);
  volatile int *gid = A.dir_gid;
  volatile uint64_t *keys = A.dir_keys;
  for (uint32_t moved = 0; moved < A.dir_cap;) {
    const int g = gid[h];
    if (g >= 0) {
      if (keys[h] == key) return g;
      h = (h + 1) & (A.dir_cap - 1);
      ++moved;
      continue;
    }
    if (g == -1 && atomicCAS(&A.dir_gid[h], -1, -2) == -1) {
      const uint32_t id = atomicAdd(A.n_groups, 1u);
      if (id >= A.partial_rows) {
        atomicExch(A.error_flag, static_cast<uint32_t>(QSGPU_ERR_CAPACITY));
        gid[h] = -1;
        return -1;
      }
      keys[h] = key;
      A.gid_keys[id] = key;
      __threadfence();
      gid[h] = static_cast<int>(id);
      return static_cast<int>(id);
    }
  }
  atomicExch(A.error_flag, static_cast<uint32_t>(QSGPU_ERR_CAPACITY));
  return -1;
}

template <class Q>
__device__ __forceinline__ void scan_agg_body(char *smem, const ScanDesc &S, const Lits &L, const AggDesc &A) {
  constexpr int HOT = Q::hot;
  constexpr int NA = Q::n_agg > 0 ? Q::n_agg : 1;
  constexpr uint32_t LG = Q::grouped ? kCompactMaxGroups : 1;
  constexpr uint32_t LS = Q::grouped ? kCompactLocalSlots : 0;
  constexpr uint32_t W = Q::words;
  const int tid = threadIdx.x;

  AggSmem M;
  {
    char *p = smem + kBarBytes + Q::n_stages * Q::stage_bytes;
    M.red = reinterpret_cast<uint64_t *>(p); p += 8 * HOT * (NA + 1) * 8;
    M.lstate = reinterpret_cast<uint64_t *>(p); p += LG * W * 8;
    M.lkey_by_id = reinterpret_cast<uint64_t *>(p); p += LG * 8;
    M.lkeys = reinterpret_cast<uint64_t *>(p); p += LS * 8;
    M.lslot = reinterpret_cast<int *>(p); p += LS * 4;
    M.lready = reinterpret_cast<uint32_t *>(p); p += LG * 4;
    M.nlocal = reinterpret_cast<uint32_t *>(p); p += 16;
    M.priv = reinterpret_cast<uint64_t *>(p);
  }
  if constexpr (Q::priv != 0) {     // every thread initialises its own slots
#pragma unroll
    for (int g = 0; g < HOT; ++g)
      static_for<0, Q::n_agg>([&](auto j) { M.priv[(g * NA + QS_IDX(j)) * kBlock + tid] = agg_identity(Q::agg_kind(QS_IDX(j))); });
  }
  if constexpr (Q::grouped) {
    for (uint32_t i = tid; i < LG * W; i += kBlock) {
      const uint32_t w = i % W;
      M.lstate[i] = w == 0 ? 0 : agg_identity(A.kind[w - 1]);
    }
    for (uint32_t i = tid; i < LS; i += kBlock) M.lslot[i] = -1;
    for (uint32_t i = tid; i < LG; i += kBlock) M.lready[i] = 0;
  }
  if (tid == 0) {
    *M.nlocal = Q::grouped ? 0u : 1u;
    if (!Q::grouped) M.lkey_by_id[0] = 0;
  }
  // (scan_tiles starts with a __syncthreads)

  AggSink<Q> sink;
  sink.lstate = M.lstate;
#pragma unroll
  for (int g = 0; g < HOT; ++g) {
    sink.hc[g] = 0;
    if constexpr (Q::priv == 0)
      static_for<0, Q::n_agg>([&](auto j) { sink.hv[g][QS_IDX(j)] = agg_identity(Q::agg_kind(QS_IDX(j))); });
  }
  VmRegs regs;
  uint64_t hk[HOT];          // keys of the register-resident groups (ids 0..nhot-1)
  int nhot = 0;
#pragma unroll
  for (int g = 0; g < HOT; ++g) hk[g] = 0;

  scan_tiles<Q>(S, smem, [&](uint32_t tile, const char *__restrict__ stage, const ScanRt &rt) {
    bool valid[kRows];
    tile_valid(S, rt, tile, tid, valid);
    uint32_t bits[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) bits[r] = valid[r] ? 1u : 0u;
    SinkBase ns;
    vm_run<Q, 0, Q::n_pred>(L, S, stage, tid, regs, bits, ns);
    bool any = false;
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      const bool pass = valid[r] && (bits[r] & 1u);
      sink.slot[r] = pass ? 0 : -1;
      any |= pass;
    }
    if (!__any_sync(0xffffffffu, any)) return;
    if constexpr (Q::grouped) {
      // Fast path: compare against the (<= HOT) keys already known to live in
      // registers; only rows of other groups take the shared-memory table.
      if (nhot < HOT) {
        nhot = 0;
#pragma unroll
        for (int g = 0; g < HOT; ++g) {
          if (nhot == g && *reinterpret_cast<volatile uint32_t *>(&M.lready[g]) != 0u) {
            __threadfence_block();           // lready[g] was raised after lkey_by_id[g] was written (local_lookup)
            hk[g] = *reinterpret_cast<volatile uint64_t *>(&M.lkey_by_id[g]);
            nhot = g + 1;
          }
        }
      }
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        if (sink.slot[r] < 0) continue;
        uint64_t key[kMaxKeyWords];
        pack_key<Q>(stage, tile_row(r, tid), key);
        int s = -2;
#pragma unroll
        for (int g = 0; g < HOT; ++g) {
          bool eq;
          if constexpr (Q::key_off(Q::n_key_cols - 1) + Q::key_w(Q::n_key_cols - 1) <= 4)
            eq = static_cast<uint32_t>(key[0]) == static_cast<uint32_t>(hk[g]);   // packed key fits one word
          else
            eq = key[0] == hk[g];
          if (g < nhot && eq) s = g;
        }
        if (s == -2) s = local_lookup(key[0], M, LS, LG, A.error_flag);
        sink.slot[r] = s;
      }
    }
    if constexpr (Q::priv == 0) sink.set_masks();
    // row counts
    sink.cold = false;
    if constexpr (Q::grouped) {
      bool c = false;
#pragma unroll
      for (int r = 0; r < kRows; ++r) c |= sink.slot[r] >= HOT;
      sink.cold = __any_sync(0xffffffffu, c);
    }
    if constexpr (Q::priv != 0) sink.set_private_rows(M.priv, tid);
#pragma unroll
    for (int r = 0; r < kRows; ++r)
      static_for<0, HOT>([&](auto gg) { hot_count<QS_IDX(gg)>(sink.hc[QS_IDX(gg)], sink.slot[r]); });
    if constexpr (Q::grouped) {
      if (sink.cold) {
#pragma unroll
        for (int r = 0; r < kRows; ++r)
          if (sink.slot[r] >= HOT)
            atomicAdd(reinterpret_cast<unsigned long long *>(&M.lstate[sink.slot[r] * W]), 1ull);
      }
    }
    vm_run<Q, Q::n_mid, Q::n_total>(L, S, stage, tid, regs, bits, sink);
    if constexpr (Q::priv != 0) sink.flush_private();
  });

  flush_lip_stats<Q>(S, regs);
  // ---- CTA reduction of the register-resident (hot) groups, fixed tree.
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int g = 0; g < HOT; ++g) {
    uint64_t c = sink.hc[g];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
    if (lane == 0) M.red[(warp * HOT + g) * (NA + 1)] = c;
    static_for<0, Q::n_agg>([&](auto jj) {
      constexpr int j = QS_IDX(jj);
      constexpr uint8_t kind = Q::agg_kind(j);
      uint64_t x;
      if constexpr (Q::priv != 0) x = M.priv[(g * NA + j) * kBlock + tid];
      else x = sink.hv[g][j];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const uint64_t y = __shfl_xor_sync(0xffffffffu, x, off);
        // keep operand order lane-independent: lower lane first
        x = (lane & off) ? agg_combine(kind, y, x) : agg_combine(kind, x, y);
      }
      if (lane == 0) M.red[(warp * HOT + g) * (NA + 1) + 1 + j] = x;
    });
  }
  __syncthreads();
  const uint32_t nlocal = min(*M.nlocal, LG);
  if (tid < HOT * (NA + 1)) {
    const int g = tid / (NA + 1), w = tid % (NA + 1);
    if (static_cast<uint32_t>(g) < nlocal && static_cast<uint32_t>(w) < W) {
      uint64_t x = M.red[(0 * HOT + g) * (NA + 1) + w];
      for (int wp = 1; wp < kBlock / 32; ++wp) {
        const uint64_t y = M.red[(wp * HOT + g) * (NA + 1) + w];
        x = w == 0 ? x + y : agg_combine(A.kind[w - 1], x, y);
      }
      // lstate[g] holds identity (hot groups never touch it during the scan)
      M.lstate[g * W + w] = x;
    }
  }
  __syncthreads();
  // ---- publish this CTA's partial state, one row per group it met.
  for (uint32_t l = tid; l < nlocal; l += kBlock) {
    const int gid = Q::grouped ? dir_insert(M.lkey_by_id[l], A) : 0;
    if (gid < 0) continue;                  // more than partial_rows groups: error raised, nothing written
    uint64_t *dst = A.partials + (static_cast<uint64_t>(blockIdx.x) * A.partial_rows + gid) * W;
    for (uint32_t w = 0; w < W; ++w) dst[w] = M.lstate[l * W + w];
  }
}

// ============================================================ K7  scan_groupby
// The hash GROUP BY table is an open-addressing array of cap + 1 rows: keys[slot][KW], states[slot][words].
//
// Keys of one or two words (<= 16 bytes: Q3's (l_orderkey, o_orderdate, o_shippriority)) are claimed AND published by
// ONE compare-and-swap on the key itself -- 64-bit, or the 128-bit ATOMG.CAS.128 of sm_90+ -- from the all-ones
// "empty" pattern: the value that comes back says empty-now-mine / same key / other key, so a probe step is a single
// atomic with no tag word, no separate key read and no fence.  (The tag protocol it replaces -- CAS a tag, write the
// key words, __threadfence, publish the tag -- showed 43 membar and 76 long-scoreboard stall cycles per issued
// instruction in ncu, at 4 % issue utilisation.)  A key that IS all ones lives in the reserved row `cap`.
// Wider keys (3-4 words) keep the tag protocol.
constexpr uint64_t kEmptyKeyWord = ~0ull;

// One probe step of a CAS-claimed table: 0 = the key now sits in `slot` (newly inserted: *inserted = true), 1 = `slot`
// holds another key.
template <uint32_t KW>
__device__ __forceinline__ int key_claim(const uint64_t *key, uint64_t slot, const AggDesc &A, bool *inserted) {
  static_assert(KW == 1 || KW == 2, "CAS-claimed keys are one or two words");
  if constexpr (KW == 1) {
    const unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long *>(&A.keys[slot]),
                                             static_cast<unsigned long long>(kEmptyKeyWord), static_cast<unsigned long long>(key[0]));
    *inserted = old == kEmptyKeyWord;
    return (old == kEmptyKeyWord || old == key[0]) ? 0 : 1;
  } else {
    uint64_t o0, o1;
    cas128(&A.keys[slot * 2], kEmptyKeyWord, kEmptyKeyWord, key[0], key[1], o0, o1);
    *inserted = o0 == kEmptyKeyWord && o1 == kEmptyKeyWord;
    return (*inserted || (o0 == key[0] && o1 == key[1])) ? 0 : 1;
  }
}

template <uint32_t KW>
__device__ __forceinline__ bool key_is_empty_pattern(const uint64_t *key) {
  bool e = true;
#pragma unroll
  for (uint32_t i = 0; i < KW; ++i) e &= key[i] == kEmptyKeyWord;
  return e;
}

template <uint32_t KW>
__device__ __forceinline__ uint64_t key_home(const uint64_t *key, uint64_t mask) {
  uint64_t h = 0x9e3779b97f4a7c15ull;
#pragma unroll
  for (uint32_t i = 0; i < KW; ++i) h = mix64(h ^ key[i]);
  return h & mask;
}

// The reserved row of the all-ones key: counted as a group the first time anybody lands on it.
__device__ __forceinline__ int64_t reserved_slot(const AggDesc &A) {
  if (atomicCAS(&A.n_groups[1], 0u, 1u) == 0u) atomicAdd(A.n_groups, 1u);
  return static_cast<int64_t>(A.cap);
}

// Find-or-insert `key` (kw words); returns the slot or -1 when the table is full.
template <uint32_t KW>
__device__ __forceinline__ int64_t table_upsert(const uint64_t *key, const AggDesc &A) {
  const uint64_t mask = A.cap - 1;
  uint64_t slot = key_home<KW>(key, mask);
  if constexpr (KW <= 2) {
    if (key_is_empty_pattern<KW>(key)) return reserved_slot(A);
    for (uint64_t probes = 0; probes <= mask; ++probes) {
      bool inserted;
      if (key_claim<KW>(key, slot, A, &inserted) == 0) {
        if (inserted) atomicAdd(A.n_groups, 1u);
        return static_cast<int64_t>(slot);
      }
      slot = (slot + 1) & mask;
    }
  } else {
    volatile uint32_t *tags = A.tags;
    volatile uint64_t *keys = A.keys;
    for (uint64_t probes = 0; probes <= mask;) {
      const uint32_t t = tags[slot];
      if (t == 2u) {
        bool eq = true;
#pragma unroll
        for (uint32_t i = 0; i < KW; ++i) eq &= keys[slot * KW + i] == key[i];
        if (eq) return static_cast<int64_t>(slot);
        slot = (slot + 1) & mask;
        ++probes;
        continue;
      }
      if (t == 0u && atomicCAS(&A.tags[slot], 0u, 1u) == 0u) {
#pragma unroll
        for (uint32_t i = 0; i < KW; ++i) keys[slot * KW + i] = key[i];
        __threadfence();
        tags[slot] = 2u;
        atomicAdd(A.n_groups, 1u);
        return static_cast<int64_t>(slot);
      }
      // busy (or lost the race): look at the same slot again
    }
  }
  atomicExch(A.error_flag, static_cast<uint32_t>(QSGPU_ERR_CAPACITY));
  return -1;
}

template <class Q>
struct GlobalAggSink : SinkBase {
  int64_t slot[kRows];
  const AggDesc *A;
  template <int J, int TYPE>
  __device__ __forceinline__ void emit(const uint64_t (&acc)[kRows]) {
#pragma unroll
    for (int r = 0; r < kRows; ++r)
      if (slot[r] >= 0) atomic_update<Q::agg_kind(J)>(&A->states[slot[r] * Q::words + 1 + J], acc[r]);
  }
};

template <class Q>
__device__ __forceinline__ void scan_groupby_body(char *smem, const ScanDesc &S, const Lits &L, const AggDesc &A) {
  const int tid = threadIdx.x;
  GlobalAggSink<Q> sink;
  sink.A = &A;
  VmRegs regs;
  uint32_t new_groups = 0;       // groups this thread opened, over all its tiles: one counter update per warp per kernel
  scan_tiles<Q>(S, smem, [&](uint32_t tile, const char *stage, const ScanRt &rt) {
    bool valid[kRows];
    tile_valid(S, rt, tile, tid, valid);
    uint32_t bits[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) bits[r] = valid[r] ? 1u : 0u;
    SinkBase ns;
    vm_run<Q, 0, Q::n_pred>(L, S, stage, tid, regs, bits, ns);
    bool pass[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) { pass[r] = valid[r] && (bits[r] & 1u); sink.slot[r] = -1; }
    if constexpr (Q::strategy == QS_AGG_COLLISION_FREE) {
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        if (!pass[r]) continue;
        constexpr uint32_t w = Q::key_w(0);
        const char *src = stage + Q::col_off(Q::key_col(0)) + tile_row(r, tid) * w;
        const int64_t k = w == 4 ? static_cast<int64_t>(*reinterpret_cast<const int32_t *>(src))
                                 : *reinterpret_cast<const int64_t *>(src);
        if (k >= 0 && static_cast<uint64_t>(k) < A.cap) sink.slot[r] = k;
        else atomicExch(A.error_flag, static_cast<uint32_t>(QSGPU_ERR_CAPACITY));
      }
    } else if constexpr (Q::key_words <= 2) {
      // CAS-claimed keys, in ROUNDS: one compare-and-swap per still-searching row per round, issued back to back,
      // so up to kRows random atomics per thread are in flight instead of one
      constexpr uint32_t KW = Q::key_words;
      uint64_t key[kRows][KW];
      uint64_t h[kRows];
      uint32_t pend = 0;
      const uint64_t mask = A.cap - 1;
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        uint64_t full[kMaxKeyWords];
        pack_key<Q>(stage, tile_row(r, tid), full);
#pragma unroll
        for (uint32_t i = 0; i < KW; ++i) key[r][i] = full[i];
        h[r] = key_home<KW>(key[r], mask);
        if (pass[r]) {
          if (key_is_empty_pattern<KW>(key[r])) sink.slot[r] = reserved_slot(A);
          else pend |= 1u << r;
        }
      }
      for (uint64_t probes = 0; pend != 0; ++probes) {
        if (probes > mask) { atomicExch(A.error_flag, static_cast<uint32_t>(QSGPU_ERR_CAPACITY)); break; }
        int res[kRows];
        bool ins[kRows];
#pragma unroll
        for (int r = 0; r < kRows; ++r)
          if ((pend >> r) & 1u) res[r] = key_claim<KW>(key[r], h[r], A, &ins[r]);
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
          if (!((pend >> r) & 1u)) continue;
          if (res[r] == 0) {
            sink.slot[r] = static_cast<int64_t>(h[r]);
            pend &= ~(1u << r);
            if (ins[r]) ++new_groups;
          } else {
            h[r] = (h[r] + 1) & mask;
          }
        }
      }
    } else {
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        if (!pass[r]) continue;
        uint64_t key[kMaxKeyWords];
        pack_key<Q>(stage, tile_row(r, tid), key);
        sink.slot[r] = table_upsert<Q::key_words>(key, A);
      }
    }
    bool any = false;
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      if (sink.slot[r] >= 0)
        atomicAdd(reinterpret_cast<unsigned long long *>(&A.states[sink.slot[r] * Q::words]), 1ull);
      any |= sink.slot[r] >= 0;
    }
    if (!__any_sync(0xffffffffu, any)) return;
    vm_run<Q, Q::n_mid, Q::n_total>(L, S, stage, tid, regs, bits, sink);
  });
  flush_lip_stats<Q>(S, regs);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) new_groups += __shfl_xor_sync(0xffffffffu, new_groups, off);
  if ((tid & 31) == 0 && new_groups) atomicAdd(A.n_groups, new_groups);
}

// ======================================================= K3 / K4  scan_select
// The reference builds a TupleIdSequence bitmap, then one ColumnVector per
// projected expression, then copies tuples into the destination block.  Here
// the bitmap is a warp ballot: each warp counts its survivors, one atomicAdd
// per CTA tile reserves the output range, and survivors are written straight
// from the staged tile to their final position (rows of a tile keep their
// input order).
// Survivors of a tile -> rows of the output relation: CTA-wide and tile-ordered.  QS_WARP_COMPACT (a JIT define,
// QSGPU_WARP_COMPACT=1) selects the warp-wide form, which measured SLOWER (see qs_compact.cuh).
__device__ __forceinline__ void compact_rows(const bool (&flag)[kRows], uint32_t *s, unsigned long long *counter, uint64_t capacity,
                                             uint32_t *error_flag, uint64_t (&idx)[kRows]) {
#ifdef QS_WARP_COMPACT
  (void)s;
  warp_compact(flag, counter, capacity, error_flag, idx);
#else
  cta_compact(flag, s, counter, capacity, error_flag, idx);
#endif
}

template <class Q>
struct SelectSink : SinkBase {
  const SinkDesc *K;
  uint64_t idx[kRows];     // output row, ~0 when the row does not survive
  uint64_t nm[kRows];      // NULL mask of the output row (bit = projected column), scans of NULL-able relations only
  int tid;
  template <int J>
  __device__ __forceinline__ void emit_null(const bool (&isnull)[kRows]) {
#pragma unroll
    for (int r = 0; r < kRows; ++r) nm[r] |= isnull[r] ? (1ull << J) : 0ull;
  }
  template <int J, int TYPE>
  __device__ __forceinline__ void emit(const uint64_t (&acc)[kRows]) {
    constexpr uint32_t w = Q::out_w(J);
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      if (idx[r] == ~0ull) continue;
      if constexpr (w == 4) *reinterpret_cast<uint32_t *>(K->out[J] + idx[r] * 4) = static_cast<uint32_t>(acc[r]);
      else *reinterpret_cast<uint64_t *>(K->out[J] + idx[r] * 8) = acc[r];
    }
  }
  template <int J, int COL, int W>
  __device__ __forceinline__ void emit_raw(const char *col) {
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      if (idx[r] == ~0ull) continue;
      copy_value<W>(K->out[J] + idx[r] * W, col + tile_row(r, tid) * W);
    }
  }
};

template <class Q, class SinkT>
__device__ __forceinline__ void lip_build_rows(const SinkDesc &K, const char *stage, int tid,
                                               const bool (&pass)[kRows]) {
  static_for<0, Q::n_lip_build>([&](auto ff) {
    constexpr int f = QS_IDX(ff);
    constexpr uint8_t lt = Q::lb_ltype(f);
    constexpr uint32_t w = (lt == V_I32 || lt == V_F32) ? 4u : 8u;
    const char *base = stage + Q::col_off(Q::lb_col(f));
#pragma unroll
    for (int r = 0; r < kRows; ++r)
      if (pass[r])
        lip_insert<Q::lb_kind(f)>(K.lip_build[f], static_cast<int64_t>(load_native(base + tile_row(r, tid) * w, lt)));
  });
}

template <class Q>
__device__ __forceinline__ void scan_select_body(char *smem, const ScanDesc &S, const Lits &L, const SinkDesc &K) {
  const int tid = threadIdx.x;
  uint32_t *s_compact = reinterpret_cast<uint32_t *>(smem + kBarBytes + Q::n_stages * Q::stage_bytes);
  SelectSink<Q> sink;
  sink.K = &K;
  sink.tid = tid;
  VmRegs regs;
  scan_tiles<Q>(S, smem, [&](uint32_t tile, const char *stage, const ScanRt &rt) {
    bool valid[kRows];
    tile_valid(S, rt, tile, tid, valid);
    uint32_t bits[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) bits[r] = valid[r] ? 1u : 0u;
    SinkBase ns;
    vm_run<Q, 0, Q::n_pred>(L, S, stage, tid, regs, bits, ns);
    bool pass[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) pass[r] = valid[r] && (bits[r] & 1u);

    // LIPFilterBuilder::insertValueAccessor on the survivors.
    lip_build_rows<Q, SelectSink<Q>>(K, stage, tid, pass);
    if constexpr (Q::n_out > 0) {   // BuildLIPFilter has nothing to materialise
      compact_rows(pass, s_compact, K.counter, K.capacity, K.error_flag, sink.idx);
      if constexpr (emits_null<Q>()) {
#pragma unroll
        for (int r = 0; r < kRows; ++r) sink.nm[r] = 0ull;
      }
      vm_run<Q, Q::n_mid, Q::n_total>(L, S, stage, tid, regs, bits, sink);
      if constexpr (emits_null<Q>()) {
#pragma unroll
        for (int r = 0; r < kRows; ++r)
          if (sink.idx[r] != ~0ull) K.null_out[sink.idx[r]] = sink.nm[r];
      }
    }
  });
  flush_lip_stats<Q>(S, regs);
}

// Join key of row `row` of the staged tile: one INT/LONG attribute, or two INT attributes packed into 64 bits.
template <class Q>
__device__ __forceinline__ int64_t join_key(const char *stage, uint32_t row) {
  constexpr uint8_t klt = Q::j_key_ltype;
  constexpr uint32_t kw = (klt == V_I32) ? 4u : 8u;
  const int64_t k0 = static_cast<int64_t>(load_native(stage + Q::col_off(Q::j_key_col) + row * kw, klt));
  if constexpr (Q::j_key2) {
    const uint32_t k1 = *reinterpret_cast<const uint32_t *>(stage + Q::col_off(Q::j_key2_col) + row * 4u);
    return static_cast<int64_t>((static_cast<uint64_t>(k1) << 32) | static_cast<uint32_t>(k0));
  } else {
    return k0;
  }
}

// ============================================================== K5 join build
// The reference's JoinHashTable is a separate-chaining multi-map from key to
// TupleReference{block, tuple}; probes collect (probe_tid, build_tid) pairs per
// build block and re-open every build block.  On the device the table is one
// open-addressing array of 16-byte {key, build row} slots (linear probing,
// duplicates occupy their own slots), the probe keys arrive as TMA-staged
// tiles, and matched pairs are projected in the same kernel: build-side
// operands are gathered through the stored row id.  Rows that share a probe
// tile advance in lock-step "rounds" (one match per row per round) so that the
// warp-ballot compaction and the VM stay CTA-uniform even with duplicate keys.
constexpr unsigned long long kEmptyRow = ~0ull;
constexpr unsigned long long kChainBit = 1ull << 63;     // dense join heads: "more rows follow in next[]"

template <class Q>
__device__ __forceinline__ void join_build_body(char *smem, const ScanDesc &S, const Lits &L, const SinkDesc &K,
                                                const JoinDesc &J) {
  const int tid = threadIdx.x;
  VmRegs regs;
  const uint64_t mask = J.cap - 1;
  uint32_t inserted = 0;       // this thread's inserts over ALL its tiles: one counter update per warp per kernel
  bool dup = false;            // met an equal key on the way to a free slot
  scan_tiles<Q>(S, smem, [&](uint32_t tile, const char *stage, const ScanRt &rt) {
    bool valid[kRows];
    tile_valid(S, rt, tile, tid, valid);
    uint32_t bits[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) bits[r] = valid[r] ? 1u : 0u;
    SinkBase ns;
    vm_run<Q, 0, Q::n_pred>(L, S, stage, tid, regs, bits, ns);
    const uint64_t row0 = S.first_row + static_cast<uint64_t>(tile) * kTileRows;
    bool pass[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) pass[r] = valid[r] && (bits[r] & 1u);
    lip_build_rows<Q, SinkBase>(K, stage, tid, pass);
    if constexpr (Q::j_dense) {
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        if (!pass[r]) continue;
        const int64_t key = join_key<Q>(stage, tile_row(r, tid));
        const unsigned long long row = row0 + tile_row(r, tid);
        // push the row on the front of its key's chain: one exchange, one store
        const uint64_t k = static_cast<uint64_t>(key - J.min_key);
        if (key < J.min_key || k >= J.cap) { atomicExch(J.error_flag, static_cast<uint32_t>(QSGPU_ERR_INVALID)); continue; }
        // A head whose chain has more than one row carries kChainBit, so a probe of a unique key (the
        // common case: the build side is a primary key) never has to read next[] to learn the chain ended.
        // Every insert that displaces a non-empty head ORs the bit in afterwards; whichever order the
        // exchanges and ORs land in, the final head of a multi-row chain has it set.
        const unsigned long long old = atomicExch(&J.heads[k], row);
        J.next[row] = old == kEmptyRow ? kEmptyRow : (old & ~kChainBit);
        if (old != kEmptyRow) atomicOr(&J.heads[k], kChainBit);
        ++inserted;
      }
    } else {
      // Open addressing, in ROUNDS: every round issues one compare-and-swap for each of the thread's rows that is
      // still looking for a slot -- back to back, so up to kRows random atomics per thread are in flight -- and
      // only then looks at what came back.  (The row-after-row loop it replaces kept ONE atomic in flight per
      // thread: ncu showed 7 % issue utilisation with 72 long-scoreboard stall cycles per issued instruction.)
      int64_t key[kRows];
      uint64_t h[kRows];
      uint32_t pend = 0;
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        key[r] = join_key<Q>(stage, tile_row(r, tid));
        h[r] = mix64(static_cast<uint64_t>(key[r])) & mask;
        if (pass[r]) pend |= 1u << r;
      }
      for (uint64_t probes = 0; pend != 0; ++probes) {
        if (probes > mask) { atomicExch(J.error_flag, static_cast<uint32_t>(QSGPU_ERR_CAPACITY)); break; }
        // {key, row} goes in with ONE 128-bit compare-and-swap against the empty slot {0, ~0}: no separate key store,
        // and what comes back on a collision is the occupant's key -- an equal one means the table holds duplicates
        uint64_t ok[kRows], orow[kRows];
#pragma unroll
        for (int r = 0; r < kRows; ++r)
          if ((pend >> r) & 1u)
            cas128(&J.slots[h[r]], 0ull, kEmptyRow, static_cast<uint64_t>(key[r]), row0 + tile_row(r, tid), ok[r], orow[r]);
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
          if (!((pend >> r) & 1u)) continue;
          if (orow[r] == kEmptyRow) {
            pend &= ~(1u << r);
            ++inserted;
          } else {
            dup |= static_cast<int64_t>(ok[r]) == key[r];
            h[r] = (h[r] + 1) & mask;
          }
        }
      }
    }
  });
  flush_lip_stats<Q>(S, regs);
  // (a per-tile update put one same-address atomic per warp per 1024 rows on the table's entry counter)
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) inserted += __shfl_xor_sync(0xffffffffu, inserted, off);
  if ((tid & 31) == 0 && inserted) atomicAdd(J.n_entries, static_cast<unsigned long long>(inserted));
  // n_entries[1]: "some key occurs more than once".  While it is 0 a probe stops at its first match instead of walking
  // on to the end of the (non-existent) duplicate run -- half the random accesses of a primary-key join.
  if (__any_sync(0xffffffffu, dup) && (tid & 31) == 0) J.n_entries[1] = 1ull;
}

// ============================================================== K6 join probe
template <class Q>
struct JoinSink : SinkBase {
  const SinkDesc *K;
  const JoinDesc *J;
  uint64_t idx[kRows];
  unsigned long long brow[kRows];
  uint64_t nm[kRows];      // NULL-ness the probe side brings (probe relations with NULL-able attributes only)
  int tid;
  template <int JJ>
  __device__ __forceinline__ void emit_null(const bool (&isnull)[kRows]) {
#pragma unroll
    for (int r = 0; r < kRows; ++r) nm[r] |= isnull[r] ? (1ull << JJ) : 0ull;
  }
  template <int JJ, int TYPE>
  __device__ __forceinline__ void emit(const uint64_t (&acc)[kRows]) {
    constexpr uint32_t w = Q::out_w(JJ);
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      if (idx[r] == ~0ull) continue;
      uint64_t v = acc[r];
      if constexpr (Q::j_type == QS_JOIN_LEFT_OUTER) {
        // an expression over the build side of a probe row without a match is NULL: zero bytes
        if (brow[r] == kEmptyRow && ((K->null_bits >> JJ) & 1ull)) v = 0;
      }
      if constexpr (w == 4) *reinterpret_cast<uint32_t *>(K->out[JJ] + idx[r] * 4) = static_cast<uint32_t>(v);
      else *reinterpret_cast<uint64_t *>(K->out[JJ] + idx[r] * 8) = v;
    }
  }
  template <int JJ, int COL, int W>
  __device__ __forceinline__ void emit_raw(const char *col) {
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      if (idx[r] == ~0ull) continue;
      copy_value<W>(K->out[JJ] + idx[r] * W, col + tile_row(r, tid) * W);
    }
  }
  template <int JJ, int COL, int W>
  __device__ __forceinline__ void emit_raw_build() {
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      if (idx[r] == ~0ull) continue;
      if (brow[r] == kEmptyRow) {
        // probe row without a match (only a LEFT OUTER join emits such rows): the value is NULL, stored as zeros
        if constexpr (Q::j_type == QS_JOIN_LEFT_OUTER) {
#pragma unroll
          for (uint32_t b = 0; b < W; ++b) K->out[JJ][idx[r] * W + b] = 0;
        }
        continue;
      }
      copy_value<W>(K->out[JJ] + idx[r] * W, build_value<COL, W>(brow[r]));
    }
  }
  // Address of build row `row`'s value of build column COL: the native column, or -- for a dictionary-coded
  // build attribute -- the dictionary entry its code selects (the TupleReference gather goes through the code).
  template <int COL, int W>
  __device__ __forceinline__ const char *build_value(uint64_t row) const {
    const ColDesc &C = J->build_cols[COL];
    if constexpr (Q::build_cw(COL) != 0) {
      uint32_t c = load_code_w<Q::build_cw(COL)>(C.ptr + row * Q::build_cw(COL));
      if constexpr (Q::build_cw(COL) == 4) c = c < C.dict_entries ? c : C.dict_entries - 1;
      return C.dict + static_cast<uint64_t>(c) * W;
    } else {
      return C.ptr + row * W;
    }
  }
  template <int COL, int LTYPE, int W>
  __device__ __forceinline__ uint64_t build_leaf(int r) {
    if (brow[r] == kEmptyRow) return 0;
    return load_native(build_value<COL, W>(brow[r]), LTYPE);
  }
  // NULL mask of the build row matched by row r (only programs over a NULL-able build relation ask)
  __device__ __forceinline__ uint64_t build_null_mask(int r) {
    return brow[r] == kEmptyRow ? 0ull : J->build_nulls[brow[r]];
  }
};

template <class Q>
__device__ __forceinline__ void join_probe_body(char *smem, const ScanDesc &S, const Lits &L, const SinkDesc &K,
                                                const JoinDesc &J) {
  const int tid = threadIdx.x;
  uint32_t *s_compact = reinterpret_cast<uint32_t *>(smem + kBarBytes + Q::n_stages * Q::stage_bytes);
  JoinSink<Q> sink;
  sink.K = &K;
  sink.J = &J;
  sink.tid = tid;
  VmRegs regs;
  const uint64_t mask = J.cap - 1;
  constexpr bool has_residual = Q::n_mid > Q::n_pred;
  constexpr bool outer = Q::j_type == QS_JOIN_LEFT_OUTER;
  const bool unique_keys = !Q::j_dense && *reinterpret_cast<const volatile unsigned long long *>(&J.n_entries[1]) == 0ull;
  // a LEFT OUTER join emits its matches exactly like an inner join, then the probe rows that never matched
  constexpr bool inner = Q::j_type == QS_JOIN_INNER || outer;

  scan_tiles<Q>(S, smem, [&](uint32_t tile, const char *stage, const ScanRt &rt) {
    bool valid[kRows];
    tile_valid(S, rt, tile, tid, valid);
    uint32_t bits[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) bits[r] = valid[r] ? 1u : 0u;
    SinkBase ns;
    vm_run<Q, 0, Q::n_pred>(L, S, stage, tid, regs, bits, ns);

    bool pass[kRows], active[kRows], matched[kRows];
    int64_t key[kRows];
    uint64_t h[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      pass[r] = valid[r] && (bits[r] & 1u);
      active[r] = pass[r];
      if constexpr (Q::j_null_col != 0xffffu) {
        // a NULL key passes but does not search (see JoinDesc::null_col)
        const uint64_t m = *reinterpret_cast<const uint64_t *>(stage + Q::col_off(Q::j_null_col) + tile_row(r, tid) * 8u);
        active[r] = active[r] && (m & J.key_null_bits) == 0ull;
      }
      matched[r] = false;
      key[r] = join_key<Q>(stage, tile_row(r, tid));
      if constexpr (Q::j_dense) {
        // h[r] walks the chain of build rows: head of the key's chain, then next[]
        const uint64_t k = static_cast<uint64_t>(key[r] - J.min_key);
        h[r] = (active[r] && key[r] >= J.min_key && k < J.cap) ? J.heads[k] : kEmptyRow;
      } else {
        h[r] = mix64(static_cast<uint64_t>(key[r])) & mask;
      }
    }

    bool first_step = true;
    while (true) {
      bool found[kRows];
      bool any = false;
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        found[r] = false;
        sink.brow[r] = kEmptyRow;
        if (!active[r]) continue;
        if constexpr (Q::j_dense) {
          if (h[r] != kEmptyRow) {
            found[r] = true;
            sink.brow[r] = h[r] & ~kChainBit;
            // first step: the head says whether a chain follows; later steps walk next[] (rows there
            // never carry the bit, so the walk ends at the kEmptyRow stored by the first insert)
            h[r] = (first_step && !(h[r] & kChainBit)) ? kEmptyRow : J.next[h[r] & ~kChainBit];
          }
        }
        if constexpr (Q::j_dense) {
          if (!found[r]) active[r] = false;
          any |= found[r];
        }
      }
      if constexpr (!Q::j_dense) {
        // open addressing, step by step: every step loads the next slot of EACH row that is still searching (up to
        // kRows independent 16-byte loads in flight), then looks at what came back
        uint32_t search = 0;
#pragma unroll
        for (int r = 0; r < kRows; ++r) if (active[r]) search |= 1u << r;
        for (uint64_t probes = 0; search != 0 && probes <= mask; ++probes) {
          ulonglong2 s[kRows];
#pragma unroll
          for (int r = 0; r < kRows; ++r)
            if ((search >> r) & 1u) s[r] = *reinterpret_cast<const ulonglong2 *>(&J.slots[h[r]]);
#pragma unroll
          for (int r = 0; r < kRows; ++r) {
            if (!((search >> r) & 1u)) continue;
            if (s[r].y == kEmptyRow) { active[r] = false; search &= ~(1u << r); continue; }
            h[r] = (h[r] + 1) & mask;
            if (static_cast<int64_t>(s[r].x) == key[r]) { found[r] = true; sink.brow[r] = s[r].y; search &= ~(1u << r); }
          }
        }
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
          if (!found[r] || unique_keys) active[r] = false;     // unique keys: the first match is the only one
          any |= found[r];
        }
      }
      first_step = false;
      if (!__syncthreads_or(any)) break;
      bool ok[kRows];
      if constexpr (has_residual) {
        uint32_t rb[kRows];
#pragma unroll
        for (int r = 0; r < kRows; ++r) rb[r] = 1u;
        vm_run<Q, Q::n_pred, Q::n_mid>(L, S, stage, tid, regs, rb, sink);
#pragma unroll
        for (int r = 0; r < kRows; ++r) ok[r] = found[r] && (rb[r] & 1u);
      } else {
#pragma unroll
        for (int r = 0; r < kRows; ++r) ok[r] = found[r];
      }
      if constexpr (inner) {
        compact_rows(ok, s_compact, K.counter, K.capacity, K.error_flag, sink.idx);
#pragma unroll
        for (int r = 0; r < kRows; ++r) sink.nm[r] = 0ull;
        vm_run<Q, Q::n_mid, Q::n_total>(L, S, stage, tid, regs, bits, sink);
        if constexpr (outer || emits_null<Q>()) {
#pragma unroll
          for (int r = 0; r < kRows; ++r) {
            if constexpr (outer) matched[r] |= ok[r];
            if (sink.idx[r] != ~0ull) K.null_out[sink.idx[r]] = sink.nm[r];
          }
        }
      } else {
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
          if (ok[r]) { matched[r] = true; active[r] = false; }   // existence is enough
        }
      }
    }
    if constexpr (!inner || outer) {
      bool flag[kRows];
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        flag[r] = pass[r] && (Q::j_type == QS_JOIN_LEFT_SEMI ? matched[r] : !matched[r]);
        sink.brow[r] = kEmptyRow;
      }
      compact_rows(flag, s_compact, K.counter, K.capacity, K.error_flag, sink.idx);
#pragma unroll
      for (int r = 0; r < kRows; ++r) sink.nm[r] = 0ull;
      vm_run<Q, Q::n_mid, Q::n_total>(L, S, stage, tid, regs, bits, sink);
      if constexpr (outer || emits_null<Q>()) {
#pragma unroll
        for (int r = 0; r < kRows; ++r)
          if (sink.idx[r] != ~0ull) K.null_out[sink.idx[r]] = sink.nm[r] | (outer ? K.null_bits : 0ull);
      }
    }
  });
  flush_lip_stats<Q>(S, regs);
}

}  // namespace qs
