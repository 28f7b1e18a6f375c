// K8: radix (hash) partition of a device relation, K9: top-k.
//
//   K8  the reference repartitions through PartitionAwareInsertDestination
//       (storage/InsertDestination.cpp:471-722, partition id =
//       PartitionSchemeHeader::getPartitionId of the key); on the device the
//       partition id is mix64(key) % n_parts and rows are regrouped so that
//       each partition is one contiguous slice per column -- the send buffers
//       of the NVLink all-to-all that follows.
//   K9  SortRunGenerationOperator + SortMergeRunOperator with a LIMIT
//       (relational_operators/SortRunGenerationOperator.hpp:76,
//       SortMergeRunOperator.hpp:72; plan at ExecutionGenerator.cpp:2227-2351):
//       radix-select the k-th primary sort key, gather the <= k + ties
//       candidates, sort them with the full multi-attribute comparator in one CTA.
#include <algorithm>
#include <cstring>
#include <vector>

#include "qs_host.h"
#include "qs_lower.h"
#include "qs_vm.cuh"

namespace qs {

// ------------------------------------------------------------------- K8
struct PartDesc {
  uint32_t n_cols;
  ColDesc in[kMaxCols];
  char *out[kMaxCols];
  const char *key;
  uint8_t key_ltype;
  uint32_t n_parts;
  uint32_t range_mode;           // 0: mix64(key) % n_parts;  1: (key - min_key) / part_width (clamped);
                                 // 2: HashPartitionSchemeHeader::getPartitionId (identity hash, & or % n_parts)
  int64_t min_key;
  uint64_t part_width;
  uint32_t width_shift;          // log2(part_width) when it is a power of two, else 64 (slot mode: log2(cap / n_parts))
  uint64_t slot_mask;            // != 0: partition by the home slot of a join table with slot_mask + 1 slots
  uint64_t n_rows;
  unsigned long long *hist;      // [n_parts]
  unsigned long long *cursor;    // [n_parts] absolute write positions
  // Fused partition + all-to-all: when set, partition p's rows go to out_table[p * n_cols + c], which may be
  // the receive relation of ANOTHER GPU mapped over NVLink (CUDA IPC), at the rows reserved from cursor[p].
  char *const *out_table;
};

__device__ __forceinline__ uint32_t part_of(const PartDesc &D, uint64_t row) {
  const int64_t k = static_cast<int64_t>(load_native(D.key + row * native_width(D.key_ltype), D.key_ltype));
  if (D.range_mode == 2) {
    // The reference's own hash partitioning of a relation on one INT / LONG attribute: TypedValue::getHash() is the
    // bit pattern of an inline scalar (types/TypedValue.hpp:575-607), the partition is hash & (n - 1) for a power of
    // two and hash % n otherwise (catalog/PartitionSchemeHeader.hpp:207-214) -- so rows land in the partition the
    // reference's PartitionAwareInsertDestination would put them in
    const uint64_t h = D.key_ltype == V_I32 ? static_cast<uint64_t>(static_cast<uint32_t>(k)) : static_cast<uint64_t>(k);
    return static_cast<uint32_t>((D.n_parts & (D.n_parts - 1)) == 0 ? (h & (D.n_parts - 1)) : (h % D.n_parts));
  }
  if (D.range_mode) {
    // key-range partitions: partition p holds keys [min + p*width, min + (p+1)*width), so the slice of a
    // dense join table (and of the range-partitioned build relation) one partition touches is contiguous.
    // A power-of-two width is a shift; any other width pays a 64-bit division per row (~100 instructions,
    // which made the two partition passes of the 1 Gi-row microbench cost more than their memory traffic).
    if (k < D.min_key) return 0u;
    const uint64_t off = static_cast<uint64_t>(k - D.min_key);
    const uint64_t p = D.width_shift < 64 ? off >> D.width_shift : off / D.part_width;
    return p >= D.n_parts ? D.n_parts - 1 : static_cast<uint32_t>(p);
  }
  if (D.slot_mask) {
    // partition = leading bits of the row's home slot in an open-addressing join table of slot_mask+1 slots:
    // one partition's keys probe one contiguous slice of the table (plus the few slots linear probing spills
    // into the next slice), so a partition-at-a-time probe keeps its slice in L2
    return static_cast<uint32_t>((mix64(static_cast<uint64_t>(k)) & D.slot_mask) >> D.width_shift);
  }
  // multiply-shift range reduction of the hash's high word: uniform, and no 64-bit modulo per row
  return __umulhi(static_cast<uint32_t>(mix64(static_cast<uint64_t>(k)) >> 32), D.n_parts);
}

__global__ void __launch_bounds__(kBlock) k_part_hist(const __grid_constant__ PartDesc D) {
  extern __shared__ unsigned int s_hist[];
  for (uint32_t p = threadIdx.x; p < D.n_parts; p += blockDim.x) s_hist[p] = 0;
  __syncthreads();
  for (uint64_t row = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; row < D.n_rows;
       row += static_cast<uint64_t>(gridDim.x) * blockDim.x)
    atomicAdd(&s_hist[part_of(D, row)], 1u);
  __syncthreads();
  for (uint32_t p = threadIdx.x; p < D.n_parts; p += blockDim.x)
    if (s_hist[p]) atomicAdd(&D.hist[p], static_cast<unsigned long long>(s_hist[p]));
}

// Scatter: a CTA takes kPartRows rows per thread (2048 rows per step), counts them per partition in shared
// memory, reserves one contiguous output range per partition with ONE global atomic each, then moves the
// step's rows column by column THROUGH shared memory: values are first placed in partition order in a
// 32 KB buffer, and consecutive threads then copy consecutive buffer elements, i.e. consecutive output
// addresses of one partition.  Writing straight from registers made every warp store touch up to 32
// partitions (32 sectors per instruction): 19.6 ms per 1 Gi 8-byte rows for 32 partitions, growing with the
// partition count (r01o).
// 8 rows per thread and (partition, rank) packed in one register: the kernel is latency-bound (global loads,
// the cursor reservation), so it needs many resident CTAs; the first version held 16 rows' partition and rank
// in 128 registers, ran 2 CTAs per SM and spent 79 % of its issue slots with no eligible warp (r01p ncu).
constexpr int kPartRows = 8;
constexpr uint32_t kPartStep = kBlock * kPartRows;            // 2048 rows per CTA step
__global__ void __launch_bounds__(kBlock, 6) k_part_scatter(const __grid_constant__ PartDesc D) {
  extern __shared__ __align__(16) unsigned char s_part_raw[];
  uint64_t *s_buf = reinterpret_cast<uint64_t *>(s_part_raw);                               // [kPartStep] staged values
  unsigned long long *s_base = reinterpret_cast<unsigned long long *>(s_part_raw + kPartStep * 8);   // [n_parts] global start
  unsigned int *s_count = reinterpret_cast<unsigned int *>(s_base + D.n_parts);             // [n_parts]
  unsigned int *s_off = s_count + D.n_parts;                                                // [n_parts + 1] start inside the step
  unsigned short *s_pid = reinterpret_cast<unsigned short *>(s_off + D.n_parts + 1);        // [kPartStep] partition of each staged element
  const uint64_t n_steps = (D.n_rows + kPartStep - 1) / kPartStep;
  for (uint64_t step = blockIdx.x; step < n_steps; step += gridDim.x) {
    for (uint32_t p = threadIdx.x; p < D.n_parts; p += blockDim.x) s_count[p] = 0;
    __syncthreads();
    const uint64_t row0 = step * kPartStep;
    const uint32_t n_here = static_cast<uint32_t>(min(static_cast<uint64_t>(kPartStep), D.n_rows - row0));
    uint32_t pr[kPartRows];                 // partition << 16 | rank inside the step's share of it; ~0 = no row
#pragma unroll
    for (int r = 0; r < kPartRows; ++r) {
      const uint32_t i = static_cast<uint32_t>(r) * kBlock + threadIdx.x;
      pr[r] = i < n_here ? part_of(D, row0 + i) : 0xffffffffu;
    }
#pragma unroll
    for (int r = 0; r < kPartRows; ++r)
      if (pr[r] != 0xffffffffu) pr[r] = (pr[r] << 16) | atomicAdd(&s_count[pr[r]], 1u);
    __syncthreads();
    if (threadIdx.x == 0) {          // exclusive scan of the per-partition counts (n_parts <= 1024)
      unsigned int acc = 0;
      for (uint32_t p = 0; p < D.n_parts; ++p) { s_off[p] = acc; acc += s_count[p]; }
      s_off[D.n_parts] = acc;
    }
    for (uint32_t p = threadIdx.x; p < D.n_parts; p += blockDim.x)
      s_base[p] = s_count[p] ? atomicAdd(&D.cursor[p], static_cast<unsigned long long>(s_count[p])) : 0ull;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kPartRows; ++r)
      if (pr[r] != 0xffffffffu) s_pid[s_off[pr[r] >> 16] + (pr[r] & 0xffffu)] = static_cast<unsigned short>(pr[r] >> 16);
    for (uint32_t c = 0; c < D.n_cols; ++c) {
      const uint32_t w = D.in[c].width;
      if (w == 8 || w == 4) {
        // place in partition order ...
#pragma unroll
        for (int r = 0; r < kPartRows; ++r) {
          if (pr[r] == 0xffffffffu) continue;
          const uint64_t row = row0 + static_cast<uint32_t>(r) * kBlock + threadIdx.x;
          const uint32_t pos = s_off[pr[r] >> 16] + (pr[r] & 0xffffu);
          if (w == 8) s_buf[pos] = *reinterpret_cast<const uint64_t *>(D.in[c].ptr + row * 8);
          else reinterpret_cast<uint32_t *>(s_buf)[pos] = *reinterpret_cast<const uint32_t *>(D.in[c].ptr + row * 4);
        }
        __syncthreads();
        // ... and copy out: element i of the buffer belongs to the partition whose [s_off[p], s_off[p+1]) holds i
        for (uint32_t i = threadIdx.x; i < n_here; i += kBlock) {
          const uint32_t lo = s_pid[i];
          const uint64_t dst = s_base[lo] + (i - s_off[lo]);
          char *ob = D.out_table ? D.out_table[lo * D.n_cols + c] : D.out[c];
          if (w == 8) *reinterpret_cast<uint64_t *>(ob + dst * 8) = s_buf[i];
          else *reinterpret_cast<uint32_t *>(ob + dst * 4) = reinterpret_cast<const uint32_t *>(s_buf)[i];
        }
        __syncthreads();
      } else {
        // odd widths (CHAR(n)): straight from global to global
#pragma unroll
        for (int r = 0; r < kPartRows; ++r) {
          if (pr[r] == 0xffffffffu) continue;
          const uint64_t row = row0 + static_cast<uint32_t>(r) * kBlock + threadIdx.x;
          const uint64_t dst = s_base[pr[r] >> 16] + (pr[r] & 0xffffu);
          const char *src = D.in[c].ptr + row * w;
          char *o = (D.out_table ? D.out_table[(pr[r] >> 16) * D.n_cols + c] : D.out[c]) + dst * w;
          for (uint32_t b = 0; b < w; ++b) o[b] = src[b];
        }
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------- K9
// Order-preserving map of a native value to an unsigned 64-bit key.
constexpr uint8_t kSortChar = 0x40;      // key_ltype = kSortChar | width for CHAR(n), n <= 8

__device__ __forceinline__ uint64_t sort_key(const char *p, uint8_t ltype, bool desc) {
  uint64_t k;
  if (ltype & kSortChar) {
    // CHAR(n), n <= 8: strncmp order of NUL-padded strings == order of the bytes read big-endian
    // (types/operations/comparisons/AsciiStringComparators.hpp:218-251); bytes after the first NUL are ignored
    const uint32_t w = ltype & 0x0f;
    k = 0;
    bool ended = false;
    for (uint32_t b = 0; b < 8; ++b) {
      unsigned char c = (b < w && !ended) ? static_cast<unsigned char>(p[b]) : 0;
      if (c == 0) ended = true;
      k = (k << 8) | c;
    }
    return desc ? ~k : k;
  }
  switch (ltype) {
    // (-0.0 and +0.0 compare equal in the reference's `<`: one key for both, or later sort keys would never decide)
    case V_F32: {
      const double d = static_cast<double>(*reinterpret_cast<const float *>(p));
      uint64_t b = d2u(d);
      if ((b << 1) == 0ull) b = 0ull;
      k = (b >> 63) ? ~b : (b | 0x8000000000000000ull);
      break;
    }
    case V_F64: {
      uint64_t b = *reinterpret_cast<const uint64_t *>(p);
      if ((b << 1) == 0ull) b = 0ull;
      k = (b >> 63) ? ~b : (b | 0x8000000000000000ull);
      break;
    }
    default:
      k = load_native(p, ltype) ^ 0x8000000000000000ull;   // I32 / I64 / DATE key
  }
  return desc ? ~k : k;
}

struct TopkDesc {
  uint32_t n_keys;
  ColDesc key_col[4];
  uint8_t key_ltype[4];
  uint8_t desc[4];
  // NULL-able sort keys: the input's per-row NULL masks (nullptr = no sort key can be NULL), key q's bit in them
  // (0 = not NULL-able) and where a NULL sorts: rank 0 = before every value, 2 = after (values have rank 1).  The
  // reference's rule (parser/ParseOrderBy.hpp:53-66): NULLS FIRST / LAST as written, else first iff descending.
  const unsigned long long *nulls;
  uint64_t key_null_bit[4];
  uint8_t null_rank[4];
  uint64_t n_rows;
  const unsigned long long *d_n_rows;   // optional device-side row count of the input (min with n_rows)
  uint32_t n_cols;
  ColDesc in[kMaxCols];
  char *out[kMaxCols];
};

constexpr int kTopkMaxCand = 2048;

// One candidate of the final sort: per key a 2-bit rank (NULL before / value / NULL after; byte q of `ranks`) and the
// order-preserving key of the value, compared lexicographically as (rank 0, key 0, rank 1, key 1, ...), row id last.
struct SortElem { uint64_t k[4]; uint64_t row; uint32_t ranks; uint32_t pad; };

__device__ __forceinline__ bool elem_less(const SortElem &a, const SortElem &b) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t ra = (a.ranks >> (8 * i)) & 0xffu, rb = (b.ranks >> (8 * i)) & 0xffu;
    if (ra != rb) return ra < rb;
    if (a.k[i] != b.k[i]) return a.k[i] < b.k[i];
  }
  return a.row < b.row;
}

// ---- multi-block top-k whose selection state never leaves the device -------------------------------------------
// (the first version read every pass's histogram back to choose the bucket: up to eight host round trips in the
// middle of Q3's tail).  Primary keys -> up to 8 histogram passes, each closed by the LAST block to finish (ticket):
// it picks the bucket holding the k-th key, extends the prefix and raises `done` once the keys at or below the
// bucket fit the sorting CTA; later passes see `done` and return at once -> collect -> sort.  All launches are
// queued back to back; nothing waits.
struct TopkState {
  unsigned long long prefix, remaining, k, n_cand;
  unsigned int done, ticket;
  unsigned long long hist[256];
};

__device__ __forceinline__ uint64_t topk_rows(const TopkDesc &D) {
  return D.d_n_rows ? min(D.n_rows, static_cast<uint64_t>(*D.d_n_rows)) : D.n_rows;
}

// Is sort key q of `row` NULL?
__device__ __forceinline__ bool topk_key_is_null(const TopkDesc &D, uint32_t q, uint64_t row) {
  return D.nulls != nullptr && (D.nulls[row] & D.key_null_bit[q]) != 0ull;
}
// The key the radix select runs on.  For a NULL-able first sort key it is COARSE: the NULL rank in the two top bits over
// the value's key shifted down by two -- monotone in (rank, key), so every row at or below the selected threshold is
// collected (a superset of the answer when low bits tie) and the exact comparator of the final sort decides.
__device__ __forceinline__ uint64_t topk_primary_key(const TopkDesc &D, uint64_t row) {
  const uint64_t k = sort_key(D.key_col[0].ptr + row * D.key_col[0].width, D.key_ltype[0], D.desc[0] != 0);
  if (D.nulls == nullptr || D.key_null_bit[0] == 0ull) return k;
  const bool isnull = (D.nulls[row] & D.key_null_bit[0]) != 0ull;
  return (static_cast<uint64_t>(isnull ? D.null_rank[0] : 1u) << 62) | (isnull ? 0ull : (k >> 2));
}
__device__ __forceinline__ SortElem topk_elem(const TopkDesc &D, uint64_t row) {
  SortElem x;
  x.row = row;
  x.ranks = 0;
  x.pad = 0;
#pragma unroll
  for (uint32_t q = 0; q < 4; ++q) {
    if (q < D.n_keys) {
      const bool isnull = topk_key_is_null(D, q, row);
      x.k[q] = isnull ? 0ull : sort_key(D.key_col[q].ptr + row * D.key_col[q].width, D.key_ltype[q], D.desc[q] != 0);
      x.ranks |= static_cast<uint32_t>(isnull ? D.null_rank[q] : 1u) << (8 * q);
    } else {
      x.k[q] = 0;
    }
  }
  return x;
}
__device__ __forceinline__ SortElem topk_pad_elem() {       // sorts behind every row
  SortElem x;
  x.row = ~0ull;
  x.ranks = 0xffffffffu;
  x.pad = 0;
  for (int q = 0; q < 4; ++q) x.k[q] = ~0ull;
  return x;
}

__global__ void k_topk_primary_dev(const __grid_constant__ TopkDesc D, uint64_t *pk, TopkState *S, uint64_t limit) {
  const uint64_t n = topk_rows(D);
  if (blockIdx.x == 0) {
    if (threadIdx.x == 0) { S->prefix = 0; S->k = min(limit, n); S->remaining = S->k; S->n_cand = 0; S->done = 0; S->ticket = 0; }
    for (int i = threadIdx.x; i < 256; i += blockDim.x) S->hist[i] = 0;
  }
  for (uint64_t row = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; row < n;
       row += static_cast<uint64_t>(gridDim.x) * blockDim.x)
    pk[row] = topk_primary_key(D, row);
}

__global__ void __launch_bounds__(256) k_topk_hist_dev(const __grid_constant__ TopkDesc D, const uint64_t *pk, TopkState *S, int shift) {
  if (*reinterpret_cast<volatile unsigned int *>(&S->done)) return;
  __shared__ unsigned int s[256];
  __shared__ bool last;
  s[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t n = topk_rows(D);
  const uint64_t prefix = S->prefix;
  const uint64_t hi_mask = shift == 56 ? 0ull : ~0ull << (shift + 8);
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint64_t k = pk[i];
    if ((k & hi_mask) == (prefix & hi_mask)) atomicAdd(&s[(k >> shift) & 0xff], 1u);
  }
  __syncthreads();
  if (s[threadIdx.x]) atomicAdd(&S->hist[threadIdx.x], static_cast<unsigned long long>(s[threadIdx.x]));
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(&S->ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (threadIdx.x == 0) {
    volatile unsigned long long *h = S->hist;
    const uint64_t k = S->k;
    uint64_t acc = 0, remaining = S->remaining;
    int b = 0;
    for (; b < 256; ++b) {
      if (acc + h[b] >= remaining) break;
      acc += h[b];
    }
    if (b == 256) b = 255;
    const uint64_t below_or_in = (k - remaining) + acc + h[b];     // keys <= every key of bucket b
    uint64_t p = prefix | (static_cast<uint64_t>(b) << shift);
    S->remaining = remaining - acc;
    if (below_or_in <= static_cast<uint64_t>(kTopkMaxCand) || shift == 0) {
      if (shift > 0) p |= (1ull << shift) - 1;                      // take the whole bucket
      S->done = 1;
    }
    S->prefix = p;
    S->ticket = 0;
  }
  __syncthreads();
  S->hist[threadIdx.x] = 0;
}

__global__ void k_topk_collect_dev(const __grid_constant__ TopkDesc D, const uint64_t *pk, TopkState *S, uint64_t *cand) {
  const uint64_t n = topk_rows(D);
  const uint64_t threshold = S->prefix;
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    if (pk[i] <= threshold) {
      const unsigned long long pos = atomicAdd(&S->n_cand, 1ull);
      if (pos < static_cast<unsigned long long>(kTopkMaxCand)) cand[pos] = i;
    }
  }
}

// k_topk_sort with the candidate count read on the device; more than kTopkMaxCand rows tying on the primary key at
// the cut raise QSGPU_ERR_CAPACITY (reported at the next read of the result).
__global__ void __launch_bounds__(1024) k_topk_sort_dev(const __grid_constant__ TopkDesc D, const uint64_t *cand, TopkState *S,
                                                        uint32_t limit, unsigned long long *rows_out, uint32_t *error_flag) {
  extern __shared__ __align__(16) char s_raw[];
  SortElem *e = reinterpret_cast<SortElem *>(s_raw);
  uint32_t m = static_cast<uint32_t>(min(S->n_cand, static_cast<unsigned long long>(kTopkMaxCand) + 1));
  if (m > static_cast<uint32_t>(kTopkMaxCand)) {
    if (threadIdx.x == 0) atomicExch(error_flag, static_cast<uint32_t>(QSGPU_ERR_CAPACITY));
    m = kTopkMaxCand;
  }
  uint32_t N = 1;
  while (N < m) N <<= 1;
  for (uint32_t i = threadIdx.x; i < N; i += blockDim.x) {
    e[i] = i < m ? topk_elem(D, cand[i]) : topk_pad_elem();
  }
  __syncthreads();
  for (uint32_t size = 2; size <= N; size <<= 1) {
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      for (uint32_t i = threadIdx.x; i < N; i += blockDim.x) {
        const uint32_t j = i ^ stride;
        if (j > i) {
          const bool up = (i & size) == 0;
          const SortElem a = e[i], b = e[j];
          if (elem_less(b, a) == up) { e[i] = b; e[j] = a; }
        }
      }
      __syncthreads();
    }
  }
  const uint32_t n_out = min(limit, m);
  for (uint32_t i = threadIdx.x; i < n_out; i += blockDim.x) {
    const uint64_t row = e[i].row;
    for (uint32_t c = 0; c < D.n_cols; ++c) {
      const uint32_t w = D.in[c].width;
      const char *src = D.in[c].ptr + row * w;
      char *o = D.out[c] + static_cast<uint64_t>(i) * w;
      for (uint32_t b = 0; b < w; ++b) o[b] = src[b];
    }
  }
  if (threadIdx.x == 0) *rows_out = n_out;
}

// ---- the same selection in ONE launch for inputs of any size ------------------------------------------------------
// The multi-launch form above costs a launch per pass: 11 dependent launches, ~10 us each, 127 us for Q3's 1.4e5
// groups per GPU at SF100 / 8 GPUs -- all of it latency, the data is a megabyte.  Here every pass is a phase of one
// kernel, separated by a grid-wide barrier (~2 us): the grid is launched cooperatively with at most one CTA per SM,
// so all CTAs are resident and the barrier (a monotonic arrival counter in global memory) cannot deadlock.  Because a
// pass is cheap now, selection continues until at most kTopkFewCand candidates remain (or the key is exhausted), which
// keeps the final sort in block 0 short.  Candidates are appended in arrival order and then sorted with the full
// comparator (row id last), so the result is deterministic.
constexpr uint64_t kTopkFewCand = 256;
struct TopkCoopState {
  unsigned long long prefix, remaining, n_cand;
  unsigned int done, arrived;
  unsigned long long hist[256];
};

__device__ __forceinline__ void grid_barrier(unsigned int *arrived, unsigned int &generation) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int target = (generation + 1) * gridDim.x;
    atomicAdd(arrived, 1u);
    while (*reinterpret_cast<volatile unsigned int *>(arrived) < target) __nanosleep(32);
    __threadfence();
  }
  __syncthreads();
  ++generation;
}

__global__ void __launch_bounds__(256) k_topk_coop(const __grid_constant__ TopkDesc D, TopkCoopState *S, uint64_t *cand, uint32_t limit,
                                                   unsigned long long *rows_out, uint32_t *error_flag) {
  extern __shared__ __align__(16) char s_coop_raw[];
  __shared__ unsigned int s_hist[256];
  unsigned int generation = 0;
  const uint64_t n = topk_rows(D);
  const uint64_t k = min(static_cast<uint64_t>(limit), n);
  const uint64_t first = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
  const uint64_t step = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  // (S was zeroed by the host-side memset queued before the launch; remaining starts at k)
  if (blockIdx.x == 0 && threadIdx.x == 0) S->remaining = k;
  grid_barrier(&S->arrived, generation);
  for (int shift = 56; shift >= 0; shift -= 8) {
    s_hist[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t prefix = *reinterpret_cast<volatile unsigned long long *>(&S->prefix);
    const uint64_t hi_mask = shift == 56 ? 0ull : ~0ull << (shift + 8);
    for (uint64_t row = first; row < n; row += step) {
      const uint64_t key = topk_primary_key(D, row);
      if ((key & hi_mask) == (prefix & hi_mask)) atomicAdd(&s_hist[(key >> shift) & 0xff], 1u);
    }
    __syncthreads();
    if (s_hist[threadIdx.x]) atomicAdd(&S->hist[threadIdx.x], static_cast<unsigned long long>(s_hist[threadIdx.x]));
    grid_barrier(&S->arrived, generation);
    if (blockIdx.x == 0) {
      if (threadIdx.x == 0) {
        volatile unsigned long long *h = S->hist;
        uint64_t acc = 0, remaining = S->remaining;
        int b = 0;
        for (; b < 256; ++b) {
          if (acc + h[b] >= remaining) break;
          acc += h[b];
        }
        if (b == 256) b = 255;
        const uint64_t below_or_in = (k - remaining) + acc + h[b];     // keys <= every key of bucket b
        uint64_t p = prefix | (static_cast<uint64_t>(b) << shift);
        S->remaining = remaining - acc;
        if (below_or_in <= kTopkFewCand || shift == 0) {
          if (shift > 0) p |= (1ull << shift) - 1;                      // take the whole bucket
          S->done = 1;
        }
        S->prefix = p;
      }
      __syncthreads();
      S->hist[threadIdx.x] = 0;
    }
    grid_barrier(&S->arrived, generation);
    if (*reinterpret_cast<volatile unsigned int *>(&S->done)) break;
  }
  const uint64_t threshold = *reinterpret_cast<volatile unsigned long long *>(&S->prefix);
  for (uint64_t row = first; row < n; row += step) {
    if (topk_primary_key(D, row) <= threshold) {
      const unsigned long long pos = atomicAdd(&S->n_cand, 1ull);
      if (pos < static_cast<unsigned long long>(kTopkMaxCand)) cand[pos] = row;
    }
  }
  grid_barrier(&S->arrived, generation);
  if (blockIdx.x != 0) return;
  // block 0: full-comparator sort of the candidates, output
  SortElem *e = reinterpret_cast<SortElem *>(s_coop_raw);
  uint32_t m = static_cast<uint32_t>(min(*reinterpret_cast<volatile unsigned long long *>(&S->n_cand), static_cast<unsigned long long>(kTopkMaxCand) + 1));
  if (m > static_cast<uint32_t>(kTopkMaxCand)) {       // more than kTopkMaxCand rows tie on the primary key at the cut
    if (threadIdx.x == 0) atomicExch(error_flag, static_cast<uint32_t>(QSGPU_ERR_CAPACITY));
    m = kTopkMaxCand;
  }
  uint32_t N = 1;
  while (N < m) N <<= 1;
  for (uint32_t i = threadIdx.x; i < N; i += blockDim.x) {
    e[i] = i < m ? topk_elem(D, __ldcg(&cand[i])) : topk_pad_elem();
  }
  __syncthreads();
  for (uint32_t size = 2; size <= N; size <<= 1) {
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      for (uint32_t i = threadIdx.x; i < N; i += blockDim.x) {
        const uint32_t j = i ^ stride;
        if (j > i) {
          const bool up = (i & size) == 0;
          const SortElem a = e[i], b = e[j];
          if (elem_less(b, a) == up) { e[i] = b; e[j] = a; }
        }
      }
      __syncthreads();
    }
  }
  const uint32_t n_out = min(limit, m);
  for (uint32_t i = threadIdx.x; i < n_out; i += blockDim.x) {
    const uint64_t row = e[i].row;
    for (uint32_t c = 0; c < D.n_cols; ++c) {
      const uint32_t w = D.in[c].width;
      const char *src = D.in[c].ptr + row * w;
      char *o = D.out[c] + static_cast<uint64_t>(i) * w;
      for (uint32_t b = 0; b < w; ++b) o[b] = src[b];
    }
  }
  if (threadIdx.x == 0) *rows_out = n_out;
}

// Whole top-k in ONE launch for inputs of up to kTopkSingleCta rows (Q3's ~1e5 groups, Q1's 4): one CTA does
// the radix select of the limit-th primary key (byte passes, stopping as soon as the candidates fit), collects
// the candidates, sorts them with the full comparator and writes the result and its row count.  The multi-kernel
// path below needs a host round trip per pass; here the host never waits.
// beyond this many rows one CTA is slower than the multi-block passes (120 us for Q3's 1.5e5 groups at SF10)
constexpr uint64_t kTopkSingleCta = 1ull << 14;
__global__ void __launch_bounds__(1024) k_topk_single(const __grid_constant__ TopkDesc D, uint32_t limit,
                                                      unsigned long long *rows_out, uint32_t *error_flag) {
  extern __shared__ __align__(16) char s_topk_raw[];
  SortElem *e = reinterpret_cast<SortElem *>(s_topk_raw);
  __shared__ unsigned int s_hist[256];
  __shared__ unsigned long long s_prefix, s_remaining;
  __shared__ unsigned int s_m;
  __shared__ int s_done;
  const uint64_t n = D.d_n_rows ? min(D.n_rows, static_cast<uint64_t>(*D.d_n_rows)) : D.n_rows;
  const uint64_t k = min(static_cast<uint64_t>(limit), n);
  if (threadIdx.x == 0) { s_prefix = ~0ull; s_remaining = k; s_done = 0; s_m = 0; }
  __syncthreads();
  // every row is a candidate when they all fit the sorting buffer (Q1's 4 groups, a gathered top-10 per GPU):
  // no selection passes, the threshold stays at the largest key
  for (int shift = n <= static_cast<uint64_t>(kTopkMaxCand) ? -1 : 56; shift >= 0; shift -= 8) {
    if (shift == 56 && threadIdx.x == 0) s_prefix = 0;
    if (threadIdx.x < 256) s_hist[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t hi_mask = shift == 56 ? 0ull : ~0ull << (shift + 8);
    const uint64_t prefix = s_prefix;
    for (uint64_t row = threadIdx.x; row < n; row += blockDim.x) {
      const uint64_t key = topk_primary_key(D, row);
      if ((key & hi_mask) == (prefix & hi_mask)) atomicAdd(&s_hist[(key >> shift) & 0xff], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint64_t acc = 0, remaining = s_remaining;
      int b = 0;
      for (; b < 256; ++b) {
        if (acc + s_hist[b] >= remaining) break;
        acc += s_hist[b];
      }
      if (b == 256) b = 255;
      const uint64_t below_or_in = (k - remaining) + acc + s_hist[b];
      s_remaining = remaining - acc;
      uint64_t p = prefix | (static_cast<uint64_t>(b) << shift);
      if (below_or_in <= static_cast<uint64_t>(kTopkMaxCand) || shift == 0) {
        if (shift > 0) p |= (1ull << shift) - 1;       // take the whole bucket
        s_done = 1;
      }
      s_prefix = p;
    }
    __syncthreads();
    if (s_done) break;
  }
  const uint64_t threshold = s_prefix;
  for (uint64_t row = threadIdx.x; row < n; row += blockDim.x) {
    if (topk_primary_key(D, row) <= threshold) {
      const unsigned int pos = atomicAdd(&s_m, 1u);
      if (pos < static_cast<unsigned int>(kTopkMaxCand)) e[pos] = topk_elem(D, row);
    }
  }
  __syncthreads();
  uint32_t m = s_m;
  if (m > static_cast<uint32_t>(kTopkMaxCand)) {       // more than kTopkMaxCand rows tie on the primary key at the cut
    if (threadIdx.x == 0) atomicExch(error_flag, static_cast<uint32_t>(QSGPU_ERR_CAPACITY));
    m = kTopkMaxCand;
  }
  uint32_t N = 1;
  while (N < m) N <<= 1;
  for (uint32_t i = m + threadIdx.x; i < N; i += blockDim.x) e[i] = topk_pad_elem();
  __syncthreads();
  for (uint32_t size = 2; size <= N; size <<= 1) {
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      for (uint32_t i = threadIdx.x; i < N; i += blockDim.x) {
        const uint32_t j = i ^ stride;
        if (j > i) {
          const bool up = (i & size) == 0;
          const SortElem a = e[i], b = e[j];
          if (elem_less(b, a) == up) { e[i] = b; e[j] = a; }
        }
      }
      __syncthreads();
    }
  }
  const uint32_t n_out = min(limit, m);
  for (uint32_t i = threadIdx.x; i < n_out; i += blockDim.x) {
    const uint64_t row = e[i].row;
    for (uint32_t c = 0; c < D.n_cols; ++c) {
      const uint32_t w = D.in[c].width;
      const char *src = D.in[c].ptr + row * w;
      char *o = D.out[c] + static_cast<uint64_t>(i) * w;
      for (uint32_t b = 0; b < w; ++b) o[b] = src[b];
    }
  }
  if (threadIdx.x == 0) *rows_out = n_out;
}

static int grid_for(uint64_t n) {
  uint64_t g = (n + 255) / 256;
  if (g > 148ull * 8) g = 148ull * 8;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace qs

using namespace qs;

extern "C" {

int qsgpu_relation_create(int, uint32_t, const qs_attr *, uint64_t, qsgpu_relation_t *);
int qsgpu_relation_destroy(qsgpu_relation_t);
int qsgpu_relation_num_rows(qsgpu_relation_t, uint64_t *);
int qsgpu_relation_set_num_rows(qsgpu_relation_t, uint64_t);

static int partition_impl(qsgpu_relation_t input, uint32_t key_attr, uint32_t n_parts, uint32_t range_mode, int64_t min_key,
                          uint64_t part_width, qsgpu_relation_t output, uint64_t *host_offsets, uint64_t slot_mask = 0);

int qsgpu_join_partition(qsgpu_join_table_t table, qsgpu_relation_t input, uint32_t key_attr, uint32_t n_parts,
                         qsgpu_relation_t output, uint64_t *host_offsets) {
  if (!table || n_parts == 0 || (n_parts & (n_parts - 1)) != 0) { set_error(QSGPU_ERR_INVALID, "join partitioning needs a power-of-two partition count"); return QSGPU_ERR_INVALID; }
  if (table->J.dense) {
    // dense table: contiguous key ranges of equal width
    const uint64_t width = (table->J.cap + n_parts - 1) / n_parts;
    uint64_t w2 = 1;
    while (w2 < width) w2 <<= 1;
    return partition_impl(input, key_attr, n_parts, 1, table->J.min_key, w2, output, host_offsets);
  }
  if (table->J.cap < n_parts) { set_error(QSGPU_ERR_INVALID, "more partitions than table slots"); return QSGPU_ERR_INVALID; }
  table->cap_frozen = true;      // rows are grouped by slices of THIS mask: the first build must not pick a smaller one
  return partition_impl(input, key_attr, n_parts, 0, 0, table->J.cap / n_parts, output, host_offsets, table->J.cap - 1);
}

int qsgpu_radix_partition(qsgpu_relation_t input, uint32_t key_attr, uint32_t n_parts, qsgpu_relation_t output,
                          uint64_t *host_offsets) {
  return partition_impl(input, key_attr, n_parts, 0, 0, 1, output, host_offsets);
}

int qsgpu_hash_partition(qsgpu_relation_t input, uint32_t key_attr, uint32_t n_parts, qsgpu_relation_t output,
                         uint64_t *host_offsets) {
  return partition_impl(input, key_attr, n_parts, 2, 0, 1, output, host_offsets);
}

int qsgpu_range_partition(qsgpu_relation_t input, uint32_t key_attr, int64_t min_key, uint64_t part_width, uint32_t n_parts,
                          qsgpu_relation_t output, uint64_t *host_offsets) {
  if (part_width == 0) { set_error(QSGPU_ERR_INVALID, "range partition needs a positive partition width"); return QSGPU_ERR_INVALID; }
  return partition_impl(input, key_attr, n_parts, 1, min_key, part_width, output, host_offsets);
}

static int partition_impl(qsgpu_relation_t input, uint32_t key_attr, uint32_t n_parts, uint32_t range_mode, int64_t min_key,
                          uint64_t part_width, qsgpu_relation_t output, uint64_t *host_offsets, uint64_t slot_mask) {
  Device *d = device(input->dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  uint64_t n = 0;
  int st = qsgpu_relation_num_rows(input, &n);
  if (st) return st;
  if (input->has_codes()) { set_error(QSGPU_ERR_UNSUPPORTED, "this operator reads native columns; its input holds dictionary-coded attributes (select them into a temporary relation first)"); return QSGPU_ERR_UNSUPPORTED; }
  if (n_parts == 0 || n_parts > 1024 || key_attr >= input->attrs.size() || output->dev != input->dev ||
      output->attrs.size() != input->attrs.size() || output->capacity < n || input->attrs.size() > static_cast<size_t>(kMaxCols)) {
    set_error(QSGPU_ERR_INVALID, "bad radix partition arguments");
    return QSGPU_ERR_INVALID;
  }
  const uint8_t lt = vtype_of(input->attrs[key_attr].type);
  if (lt != V_I32 && lt != V_I64) { set_error(QSGPU_ERR_UNSUPPORTED, "partition key must be INT/LONG"); return QSGPU_ERR_UNSUPPORTED; }
  if (key_attr < 64 && ((input->nullable_mask >> key_attr) & 1ull)) { set_error(QSGPU_ERR_UNSUPPORTED, "partitioning on a NULL-able attribute"); return QSGPU_ERR_UNSUPPORTED; }
  PartDesc D{};
  D.n_cols = static_cast<uint32_t>(input->attrs.size());
  for (uint32_t c = 0; c < D.n_cols; ++c) {
    if (input->attrs[c].width != output->attrs[c].width) { set_error(QSGPU_ERR_INVALID, "partition output schema differs"); return QSGPU_ERR_INVALID; }
    D.in[c].ptr = input->cols[c];
    D.in[c].width = input->attrs[c].width;
    D.out[c] = output->cols[c];
  }
  if (input->nullable_mask) {      // the rows' NULL masks move with them, as one more 8-byte column
    if (D.n_cols >= static_cast<uint32_t>(kMaxCols)) { set_error(QSGPU_ERR_UNSUPPORTED, "too many attributes to partition along with the NULL mask"); return QSGPU_ERR_UNSUPPORTED; }
    st = ensure_null_mask(output, d);
    if (st) return st;
    output->nullable_mask |= input->nullable_mask;
    D.in[D.n_cols].ptr = reinterpret_cast<const char *>(input->d_nulls);
    D.in[D.n_cols].width = 8;
    D.out[D.n_cols] = reinterpret_cast<char *>(output->d_nulls);
    ++D.n_cols;
  }
  D.key = input->cols[key_attr];
  D.key_ltype = lt;
  D.n_parts = n_parts;
  D.range_mode = range_mode;
  D.min_key = min_key;
  D.part_width = part_width;
  D.width_shift = 64;
  if ((part_width & (part_width - 1)) == 0) { D.width_shift = 0; while ((1ull << D.width_shift) < part_width) ++D.width_shift; }
  D.slot_mask = slot_mask;
  D.n_rows = n;
  unsigned long long *d_buf = nullptr;
  QS_CUDA(dev_malloc(&d_buf, 2ull * n_parts * 8 + 64));
  QS_CUDA(cudaMemsetAsync(d_buf, 0, 2ull * n_parts * 8, d->stream));
  D.hist = d_buf;
  D.cursor = d_buf + n_parts;
  const int grid = std::min(grid_for(n), d->sm_count * 8);
  {
    cudaEvent_t e0 = d->ev0, e1 = d->ev1;
    const bool on = timing_enabled();
    if (on) cudaEventRecord(e0, d->stream);
    k_part_hist<<<grid, kBlock, n_parts * 4, d->stream>>>(D);
    std::vector<unsigned long long> hist(n_parts), cur(n_parts);
    cudaError_t e = cudaMemcpyAsync(hist.data(), D.hist, n_parts * 8, cudaMemcpyDeviceToHost, d->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(d->stream);
    if (e != cudaSuccess) { dev_free(d_buf); return cuda_fail(e, "partition histogram"); }
    uint64_t acc = 0;
    for (uint32_t p = 0; p < n_parts; ++p) { host_offsets[p] = acc; cur[p] = acc; acc += hist[p]; }
    host_offsets[n_parts] = acc;
    e = cudaMemcpyAsync(D.cursor, cur.data(), n_parts * 8, cudaMemcpyHostToDevice, d->stream);
    if (e != cudaSuccess) { dev_free(d_buf); return cuda_fail(e, "partition cursors"); }
    const size_t smem = static_cast<size_t>(kPartStep) * 8 + n_parts * 8 + (2 * n_parts + 2) * 4 + kPartStep * 2 + 16;
    k_part_scatter<<<grid, kBlock, smem, d->stream>>>(D);
    count_launch(2);
    if (on) {
      cudaEventRecord(e1, d->stream);
      cudaEventSynchronize(e1);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      record_ms(QS_K_PARTITION, ms);
    }
    e = cudaStreamSynchronize(d->stream);
    dev_free(d_buf);
    if (e != cudaSuccess) return cuda_fail(e, "partition scatter");
  }
  return qsgpu_relation_set_num_rows(output, n);
}

// ---- K8 fused with the exchange: count, then scatter straight into the peers' receive relations
static int fill_part_desc(qsgpu_relation_t input, uint32_t key_attr, uint32_t n_parts, PartDesc *D, uint64_t *n_rows) {
  int st = qsgpu_relation_num_rows(input, n_rows);
  if (st) return st;
  if (input->has_codes()) { set_error(QSGPU_ERR_UNSUPPORTED, "this operator reads native columns; its input holds dictionary-coded attributes (select them into a temporary relation first)"); return QSGPU_ERR_UNSUPPORTED; }
  if (n_parts == 0 || n_parts > 1024 || key_attr >= input->attrs.size() || input->attrs.size() > static_cast<size_t>(kMaxCols)) {
    set_error(QSGPU_ERR_INVALID, "bad partition arguments");
    return QSGPU_ERR_INVALID;
  }
  const uint8_t lt = vtype_of(input->attrs[key_attr].type);
  if (lt != V_I32 && lt != V_I64) { set_error(QSGPU_ERR_UNSUPPORTED, "partition key must be INT/LONG"); return QSGPU_ERR_UNSUPPORTED; }
  if (input->nullable_mask) { set_error(QSGPU_ERR_UNSUPPORTED, "peer scatter of a relation with NULL-able attributes"); return QSGPU_ERR_UNSUPPORTED; }
  D->n_cols = static_cast<uint32_t>(input->attrs.size());
  for (uint32_t c = 0; c < D->n_cols; ++c) { D->in[c].ptr = input->cols[c]; D->in[c].width = input->attrs[c].width; }
  D->key = input->cols[key_attr];
  D->key_ltype = lt;
  D->n_parts = n_parts;
  D->part_width = 1;
  D->width_shift = 0;
  D->n_rows = *n_rows;
  return QSGPU_OK;
}

int qsgpu_partition_count(qsgpu_relation_t input, uint32_t key_attr, uint32_t n_parts, uint64_t *host_counts) {
  Device *d = device(input->dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  PartDesc D{};
  uint64_t n = 0;
  int st = fill_part_desc(input, key_attr, n_parts, &D, &n);
  if (st) return st;
  unsigned long long *d_hist = nullptr;
  QS_CUDA(dev_malloc(&d_hist, n_parts * 8 + 64));
  QS_CUDA(cudaMemsetAsync(d_hist, 0, n_parts * 8, d->stream));
  D.hist = d_hist;
  k_part_hist<<<std::min(grid_for(n), d->sm_count * 8), kBlock, n_parts * 4, d->stream>>>(D);
  count_launch();
  std::vector<unsigned long long> h(n_parts);
  cudaError_t e = cudaMemcpyAsync(h.data(), d_hist, n_parts * 8, cudaMemcpyDeviceToHost, d->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(d->stream);
  dev_free(d_hist);
  if (e != cudaSuccess) return cuda_fail(e, "partition count");
  for (uint32_t p = 0; p < n_parts; ++p) host_counts[p] = h[p];
  return QSGPU_OK;
}

int qsgpu_partition_scatter_peers(qsgpu_relation_t input, uint32_t key_attr, uint32_t n_parts, void *const *peer_cols,
                                  const uint64_t *first_rows) {
  Device *d = device(input->dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  PartDesc D{};
  uint64_t n = 0;
  int st = fill_part_desc(input, key_attr, n_parts, &D, &n);
  if (st) return st;
  if (!peer_cols || !first_rows) { set_error(QSGPU_ERR_INVALID, "scatter to peers needs the destination columns and first rows"); return QSGPU_ERR_INVALID; }
  const size_t n_ptrs = static_cast<size_t>(n_parts) * D.n_cols;
  char *d_buf = nullptr;
  QS_CUDA(dev_malloc(&d_buf, n_parts * 8 + n_ptrs * 8 + 64));
  D.cursor = reinterpret_cast<unsigned long long *>(d_buf);
  D.out_table = reinterpret_cast<char *const *>(d_buf + n_parts * 8);
  cudaError_t e = cudaMemcpyAsync(d_buf, first_rows, n_parts * 8, cudaMemcpyHostToDevice, d->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_buf + n_parts * 8, peer_cols, n_ptrs * 8, cudaMemcpyHostToDevice, d->stream);
  if (e != cudaSuccess) { dev_free(d_buf); return cuda_fail(e, "scatter to peers: descriptors"); }
  const bool on = timing_enabled();
  if (on) cudaEventRecord(d->ev0, d->stream);
  const size_t smem = static_cast<size_t>(kPartStep) * 8 + n_parts * 8 + (2 * n_parts + 2) * 4 + kPartStep * 2 + 16;
  k_part_scatter<<<std::min(grid_for(n), d->sm_count * 8), kBlock, smem, d->stream>>>(D);
  count_launch();
  if (on) {
    cudaEventRecord(d->ev1, d->stream);
    cudaEventSynchronize(d->ev1);
    float ms = 0;
    cudaEventElapsedTime(&ms, d->ev0, d->ev1);
    record_ms(QS_K_PARTITION, ms);
  }
  e = cudaStreamSynchronize(d->stream);       // the caller's barrier across ranks follows
  dev_free(d_buf);
  if (e != cudaSuccess) return cuda_fail(e, "scatter to peers");
  return QSGPU_OK;
}

// ---- device memory other processes can map (CUDA IPC): the receive relations of the fused exchange
int qsgpu_ipc_alloc(int dev, size_t bytes, void **dptr, qs_ipc_handle *handle) {
  Device *d = device(dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  QS_CUDA(cudaMalloc(dptr, bytes ? bytes : 256));          // not from the pool: IPC handles need their own allocation
  static_assert(sizeof(cudaIpcMemHandle_t) <= sizeof(qs_ipc_handle), "IPC handle size");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, *dptr);
  if (e != cudaSuccess) { cudaFree(*dptr); *dptr = nullptr; return cuda_fail(e, "cudaIpcGetMemHandle"); }
  std::memset(handle, 0, sizeof(*handle));
  std::memcpy(handle, &h, sizeof(h));
  return QSGPU_OK;
}

int qsgpu_ipc_open(int dev, const qs_ipc_handle *handle, void **dptr) {
  Device *d = device(dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, sizeof(h));
  QS_CUDA(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
  return QSGPU_OK;
}

int qsgpu_ipc_close(int dev, void *dptr) {
  Device *d = device(dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  QS_CUDA(cudaIpcCloseMemHandle(dptr));
  return QSGPU_OK;
}

int qsgpu_ipc_free(int dev, void *dptr) {
  Device *d = device(dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  QS_CUDA(cudaStreamSynchronize(d->stream));
  QS_CUDA(cudaFree(dptr));
  return QSGPU_OK;
}

int qsgpu_topk(qsgpu_relation_t input, uint32_t n_keys, const qs_sort_key *keys, uint64_t limit,
               qsgpu_relation_t *out) {
  Device *d = device(input->dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  uint64_t n = 0;
  int st = QSGPU_OK;
  // A small input whose row count is still device-side (the output of the operator before): the one-launch kernel
  // reads the count itself, the host does not wait for the producer.
  const bool device_count = input->dirty && input->capacity > 0;
  if (device_count) n = input->capacity;
  else st = qsgpu_relation_num_rows(input, &n);
  if (st) return st;
  if (input->has_codes()) { set_error(QSGPU_ERR_UNSUPPORTED, "this operator reads native columns; its input holds dictionary-coded attributes (select them into a temporary relation first)"); return QSGPU_ERR_UNSUPPORTED; }
  if (n_keys == 0 || n_keys > 4 || limit == 0 || limit > 1024 || input->attrs.size() > static_cast<size_t>(kMaxCols)) {
    set_error(QSGPU_ERR_UNSUPPORTED, "top-k supports 1..4 sort attributes and LIMIT <= 1024");
    return QSGPU_ERR_UNSUPPORTED;
  }
  TopkDesc D{};
  D.n_keys = n_keys;
  D.n_rows = n;
  D.d_n_rows = device_count ? input->d_rows : nullptr;
  for (uint32_t q = 0; q < n_keys; ++q) {
    if (keys[q].attr >= input->attrs.size()) { set_error(QSGPU_ERR_INVALID, "sort attribute out of range"); return QSGPU_ERR_INVALID; }
    const uint8_t lt = vtype_of(input->attrs[keys[q].attr].type);
    uint8_t klt = lt;
    if (lt == 0xff) {
      const qs_attr &ka = input->attrs[keys[q].attr];
      if (ka.type != QS_CHAR || ka.width > 8) { set_error(QSGPU_ERR_UNSUPPORTED, "sort keys: numeric, DATE or CHAR(n <= 8) attributes"); return QSGPU_ERR_UNSUPPORTED; }
      klt = static_cast<uint8_t>(kSortChar | ka.width);
    }
    D.key_col[q].ptr = input->cols[keys[q].attr];
    D.key_col[q].width = input->attrs[keys[q].attr].width;
    D.key_ltype[q] = klt;
    D.desc[q] = (keys[q].descending & 1u) ? 1 : 0;
    // NULL ordering (qs_sort_key.descending bits 1-2): as written, else NULLs first iff descending (ParseOrderBy.hpp:53-66)
    const uint32_t null_order = (keys[q].descending >> 1) & 3u;
    if (null_order == 3u) { set_error(QSGPU_ERR_INVALID, "qs_sort_key: NULLS FIRST and NULLS LAST both set"); return QSGPU_ERR_INVALID; }
    const bool nulls_first = null_order == 1u ? true : null_order == 2u ? false : D.desc[q] != 0;
    D.null_rank[q] = nulls_first ? 0 : 2;
    if (keys[q].attr < 64 && ((input->nullable_mask >> keys[q].attr) & 1ull)) {
      D.key_null_bit[q] = 1ull << keys[q].attr;
      D.nulls = input->d_nulls;
    }
  }
  qsgpu_relation *rel = nullptr;
  st = qsgpu_relation_create(input->dev, static_cast<uint32_t>(input->attrs.size()), input->attrs.data(),
                             std::max<uint64_t>(limit, 1), &rel);
  if (st) return st;
  D.n_cols = static_cast<uint32_t>(input->attrs.size());
  for (uint32_t c = 0; c < D.n_cols; ++c) {
    D.in[c].ptr = input->cols[c];
    D.in[c].width = input->attrs[c].width;
    D.out[c] = rel->cols[c];
  }
  if (input->nullable_mask) {      // the rows' NULL masks move with them, as one more 8-byte column
    if (D.n_cols >= static_cast<uint32_t>(kMaxCols)) { qsgpu_relation_destroy(rel); set_error(QSGPU_ERR_UNSUPPORTED, "too many attributes to sort along with the NULL mask"); return QSGPU_ERR_UNSUPPORTED; }
    st = ensure_null_mask(rel, d);
    if (st) { qsgpu_relation_destroy(rel); return st; }
    rel->nullable_mask = input->nullable_mask;
    D.in[D.n_cols].ptr = reinterpret_cast<const char *>(input->d_nulls);
    D.in[D.n_cols].width = 8;
    D.out[D.n_cols] = reinterpret_cast<char *>(rel->d_nulls);
    ++D.n_cols;
  }
  if (n == 0) { *out = rel; return qsgpu_relation_set_num_rows(rel, 0); }
  if (n <= kTopkSingleCta) {
    // one launch, no host round trip: the result's row count stays on the device until somebody asks
    const size_t smem = static_cast<size_t>(kTopkMaxCand) * sizeof(SortElem);
    cudaFuncSetAttribute(k_topk_single, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));   // per device
    const bool on = timing_enabled();
    if (on) cudaEventRecord(d->ev0, d->stream);
    k_topk_single<<<1, 1024, smem, d->stream>>>(D, static_cast<uint32_t>(limit), rel->d_rows, d->d_error);
    count_launch();
    cudaError_t ke = cudaGetLastError();
    if (ke != cudaSuccess) { qsgpu_relation_destroy(rel); return cuda_fail(ke, "top-k"); }
    if (on) {
      cudaEventRecord(d->ev1, d->stream);
      cudaEventSynchronize(d->ev1);
      float ms = 0;
      cudaEventElapsedTime(&ms, d->ev0, d->ev1);
      record_ms(QS_K_TOPK, ms);
    }
    rel->dirty = true;
    *out = rel;
    return QSGPU_OK;
  }

  {
    // one cooperative launch: every pass is a phase of the same kernel (see k_topk_coop).  Falls through to the
    // multi-launch form when the device refuses the cooperative launch or QSGPU_TOPK_COOP=0.
    const char *coop_env = std::getenv("QSGPU_TOPK_COOP");          // read per call: tests switch between the two forms
    const bool coop_enabled = !(coop_env && coop_env[0] == '0');
    int coop_ok = 0;
    cudaDeviceGetAttribute(&coop_ok, cudaDevAttrCooperativeLaunch, d->id);
    if (coop_enabled && coop_ok) {
      uint64_t *cand2 = nullptr;
      TopkCoopState *st2 = nullptr;
      cudaError_t ce = dev_malloc(&cand2, kTopkMaxCand * 8);
      if (ce == cudaSuccess) ce = dev_malloc(&st2, sizeof(TopkCoopState));
      if (ce == cudaSuccess) ce = cudaMemsetAsync(st2, 0, sizeof(TopkCoopState), d->stream);
      const size_t smem = static_cast<size_t>(kTopkMaxCand) * sizeof(SortElem);
      if (ce == cudaSuccess) ce = cudaFuncSetAttribute(k_topk_coop, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      const bool on = timing_enabled();
      if (ce == cudaSuccess) {
        if (on) cudaEventRecord(d->ev0, d->stream);
        const int grid = static_cast<int>(std::min<uint64_t>(static_cast<uint64_t>(d->sm_count), (n + 1023) / 1024));
        uint32_t lim32 = static_cast<uint32_t>(limit);
        unsigned long long *rows_out = rel->d_rows;
        uint32_t *err = d->d_error;
        void *args[] = {&D, &st2, &cand2, &lim32, &rows_out, &err};
        ce = cudaLaunchCooperativeKernel(reinterpret_cast<const void *>(k_topk_coop), dim3(static_cast<unsigned>(std::max(grid, 1))), dim3(256), args, smem, d->stream);
      }
      if (ce == cudaSuccess) {
        count_launch();
        if (on) {
          cudaEventRecord(d->ev1, d->stream);
          cudaEventSynchronize(d->ev1);
          float ms = 0;
          cudaEventElapsedTime(&ms, d->ev0, d->ev1);
          record_ms(QS_K_TOPK, ms);
        }
        dev_free(cand2); dev_free(st2);
        rel->dirty = true;
        *out = rel;
        return QSGPU_OK;
      }
      cudaGetLastError();
      dev_free(cand2); dev_free(st2);
    }
  }
  uint64_t *pk = nullptr, *cand = nullptr;
  TopkState *state = nullptr;
  auto cleanup = [&]() { dev_free(pk); dev_free(cand); dev_free(state); };     // stream-ordered: after the launches below
  cudaError_t e = dev_malloc(&pk, n * 8 + 64);
  if (e == cudaSuccess) e = dev_malloc(&cand, kTopkMaxCand * 8);
  if (e == cudaSuccess) e = dev_malloc(&state, sizeof(TopkState));
  if (e != cudaSuccess) { cleanup(); qsgpu_relation_destroy(rel); return cuda_fail(e, "top-k scratch"); }
  const bool on = timing_enabled();
  if (on) cudaEventRecord(d->ev0, d->stream);
  const int grid = grid_for(n);
  k_topk_primary_dev<<<grid, 256, 0, d->stream>>>(D, pk, state, limit);
  // radix select of the k-th smallest primary key, one byte per pass; the passes after the one that raised `done`
  // (for Q3's 1e5..1e6 groups: the 2nd or 3rd) return immediately
  for (int shift = 56; shift >= 0; shift -= 8) k_topk_hist_dev<<<grid, 256, 0, d->stream>>>(D, pk, state, shift);
  k_topk_collect_dev<<<grid, 256, 0, d->stream>>>(D, pk, state, cand);
  const size_t smem = static_cast<size_t>(kTopkMaxCand) * sizeof(SortElem);
  cudaFuncSetAttribute(k_topk_sort_dev, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  k_topk_sort_dev<<<1, 1024, smem, d->stream>>>(D, cand, state, static_cast<uint32_t>(limit), rel->d_rows, d->d_error);
  count_launch(11);
  e = cudaGetLastError();
  if (on) {
    cudaEventRecord(d->ev1, d->stream);
    cudaEventSynchronize(d->ev1);
    float ms = 0;
    cudaEventElapsedTime(&ms, d->ev0, d->ev1);
    record_ms(QS_K_TOPK, ms);
  }
  cleanup();
  if (e != cudaSuccess) { qsgpu_relation_destroy(rel); return cuda_fail(e, "top-k"); }
  rel->dirty = true;                  // the row count stays on the device until somebody asks
  *out = rel;
  return QSGPU_OK;
}

}  // extern "C"
