// K0: decode one staged storage-block stripe into a native-width device column.
//
//   dictionary codes    CompressedColumnStoreTupleStorageSubBlock::getAttributeValue
//                       (storage/CompressedColumnStoreTupleStorageSubBlock.cpp:203-215)
//                       -> CompressionDictionaryLite::getUntypedValueForCode
//                       (compression/CompressionDictionaryLite.hpp:40-51, values sorted)
//   truncated ints      CompressedBlockBuilder truncation to 1/2/4 bytes of a
//                       non-negative INT/LONG (storage/CompressedBlockBuilder.cpp:434-506)
//   row-store slots     SplitRowStoreTupleStorageSubBlock fixed-width attribute
//                       at slot + offset (storage/SplitRowStoreTupleStorageSubBlock.cpp:103-179)
//
// All three are pure streaming byte shuffles: one thread per value, coalesced
// stores, the dictionary (<= a few hundred KB per 4 MB block) stays in L2.
#include "qs_ops.cuh"

namespace qs {

__device__ __forceinline__ uint32_t load_code(const unsigned char *codes, uint64_t i, uint32_t cw) {
  switch (cw) {
    case 1: return codes[i];
    case 2: return reinterpret_cast<const uint16_t *>(codes)[i];
    default: return reinterpret_cast<const uint32_t *>(codes)[i];
  }
}

__global__ void k_decode_dict(char *dst, const unsigned char *codes, const char *dict, uint64_t n,
                              uint32_t cw, uint32_t vw, uint32_t dict_entries) {
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    uint32_t c = load_code(codes, i, cw);
    if (c >= dict_entries) c = dict_entries - 1;   // never for a well-formed block
    const char *src = dict + static_cast<uint64_t>(c) * vw;
    char *d = dst + i * vw;
    if (vw == 8) *reinterpret_cast<uint64_t *>(d) = *reinterpret_cast<const uint64_t *>(src);
    else if (vw == 4) *reinterpret_cast<uint32_t *>(d) = *reinterpret_cast<const uint32_t *>(src);
    else for (uint32_t b = 0; b < vw; ++b) d[b] = src[b];
  }
}

__global__ void k_decode_truncated(char *dst, const unsigned char *codes, uint64_t n, uint32_t cw,
                                   uint32_t vw) {
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint32_t c = load_code(codes, i, cw);
    if (vw == 8) reinterpret_cast<int64_t *>(dst)[i] = static_cast<int64_t>(c);
    else reinterpret_cast<int32_t *>(dst)[i] = static_cast<int32_t>(c);
  }
}

__global__ void k_decode_strided(char *dst, const char *slots, uint64_t n, uint32_t stride, uint32_t vw) {
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const char *src = slots + i * stride;
    char *d = dst + i * vw;
    for (uint32_t b = 0; b < vw; ++b) d[b] = src[b];   // slots are not aligned to vw
  }
}

// Batched form: every (block, attribute) stripe of a staging batch is one
// StageSeg; the batch is cut into 4096-row tiles and a persistent grid walks
// the tiles (binary search tile -> segment), so a whole relation's blocks are
// decoded by ONE launch instead of one per stripe.
// Code stripes inside a block image start wherever the previous stripe ended
// (stripe i is max_tuples x attribute_size(i) bytes), so 2/4-byte codes may be misaligned.
__device__ __forceinline__ uint32_t load_code_any(const unsigned char *codes, uint64_t i, uint32_t cw, bool aligned) {
  if (aligned || cw == 1) return load_code(codes, i, cw);
  uint32_t v = 0;
  for (uint32_t b = 0; b < cw; ++b) v |= static_cast<uint32_t>(codes[i * cw + b]) << (8 * b);
  return v;
}

__device__ __forceinline__ void copy_vw(char *d, const char *src, uint32_t vw, bool aligned) {
  if (aligned && vw == 8) *reinterpret_cast<uint64_t *>(d) = *reinterpret_cast<const uint64_t *>(src);
  else if (aligned && vw == 4) *reinterpret_cast<uint32_t *>(d) = *reinterpret_cast<const uint32_t *>(src);
  else if (aligned && vw == 2) *reinterpret_cast<uint16_t *>(d) = *reinterpret_cast<const uint16_t *>(src);
  else for (uint32_t b = 0; b < vw; ++b) d[b] = src[b];
}

__global__ void __launch_bounds__(256) k_decode_segments(const StageSeg *segs, uint32_t n_segs, uint64_t n_tiles) {
  __shared__ StageSeg sg;
  for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t lo = 0, hi = n_segs - 1;            // last segment with tile_begin <= tile
      while (lo < hi) {
        const uint32_t mid = (lo + hi + 1) >> 1;
        if (segs[mid].tile_begin <= tile) lo = mid; else hi = mid - 1;
      }
      sg = segs[lo];
    }
    __syncthreads();
    const uint64_t r0 = (tile - sg.tile_begin) * kStageTileRows;
    const uint64_t r1 = min(sg.n_rows, r0 + kStageTileRows);
    const bool al = (sg.aligned & 1) != 0;
    // DATE values are DateLit structs {int32 year; u8 month; u8 day; 2 padding bytes}: the reference never initialises
    // the padding (its block files carry garbage there), and a group-by / join key is compared as 8 raw bytes on the
    // device -- so staging canonicalises every date to zero padding
    const bool date = (sg.aligned & 4) != 0;
    // CHAR(n) values end at their first NUL (the reference compares them with strncmp and leaves whatever was in memory
    // behind the terminator): bytes after it are zeroed, so that equal strings are equal bytes
    const bool chr = (sg.aligned & 8) != 0;
    const uint32_t vw = sg.vw;
    if (sg.encoding == QS_ENC_PLAIN && al && !date && !chr && sg.null_kind == QS_NULL_NONE && ((vw * r0) & 15) == 0 &&
        ((reinterpret_cast<uintptr_t>(sg.src) | reinterpret_cast<uintptr_t>(sg.dst)) & 15) == 0) {
      // 16-byte vector copy of the tile, byte tail
      const uint64_t b0 = r0 * vw, b1 = r1 * vw;
      const uint64_t n16 = (b1 - b0) >> 4;
      const uint4 *s4 = reinterpret_cast<const uint4 *>(sg.src + b0);
      uint4 *d4 = reinterpret_cast<uint4 *>(sg.dst + b0);
      for (uint64_t i = threadIdx.x; i < n16; i += blockDim.x) d4[i] = s4[i];
      for (uint64_t b = b0 + (n16 << 4) + threadIdx.x; b < b1; b += blockDim.x) sg.dst[b] = sg.src[b];
      continue;
    }
    for (uint64_t i = r0 + threadIdx.x; i < r1; i += blockDim.x) {
      char *d = sg.dst + i * vw;
      bool is_null = false;
      switch (sg.encoding) {
        case QS_ENC_PLAIN: copy_vw(d, sg.src + i * vw, vw, al); break;
        case QS_ENC_STRIDED: copy_vw(d, sg.src + i * sg.stride, vw, false); break;
        case QS_ENC_DICT: {
          uint32_t c = load_code_any(reinterpret_cast<const unsigned char *>(sg.src), i, sg.cw, (sg.aligned & 2) != 0);
          is_null = sg.null_kind == QS_NULL_CODE && c == sg.null_arg;
          if (c >= sg.dict_entries) c = sg.dict_entries - 1;
          copy_vw(d, sg.dict + static_cast<uint64_t>(c) * vw, vw, al);
          break;
        }
        default: {   // QS_ENC_TRUNCATED
          const uint32_t c = load_code_any(reinterpret_cast<const unsigned char *>(sg.src), i, sg.cw, (sg.aligned & 2) != 0);
          if (vw == 8) *reinterpret_cast<int64_t *>(d) = static_cast<int64_t>(c);
          else *reinterpret_cast<int32_t *>(d) = static_cast<int32_t>(c);
        }
      }
      if (date) { d[6] = 0; d[7] = 0; }
      if (chr) {
        bool ended = false;
        for (uint32_t b = 0; b < vw; ++b) { if (ended) d[b] = 0; else ended = d[b] == 0; }
      }
      if (sg.null_kind != QS_NULL_NONE) {
        // both bitmap forms are most-significant-bit-first bit strings inside little-endian words: one byte load
        if (sg.null_kind == QS_NULL_BITMAP) {
          const uint64_t g = sg.null_arg + i * sg.null_stride;
          is_null = (sg.null_src[(g >> 6) * 8 + 7 - ((g & 63) >> 3)] >> (7 - (g & 7))) & 1;
        } else if (sg.null_kind == QS_NULL_SLOT_WORD) {
          is_null = (sg.null_src[i * sg.null_stride + (sg.null_width - 1 - (sg.null_arg >> 3))] >> (7 - (sg.null_arg & 7))) & 1;
        }
        // rows of one relation row are decoded by different CTAs (one stripe each): atomics on the shared mask word
        if (is_null) {
          for (uint32_t b = 0; b < vw; ++b) d[b] = 0;
          atomicOr(&sg.null_dst[i], static_cast<unsigned long long>(sg.null_bit));
        } else {
          atomicAnd(&sg.null_dst[i], ~static_cast<unsigned long long>(sg.null_bit));
        }
      }
    }
  }
}

__global__ void k_zero_after_nul(char *col, uint64_t n, uint32_t w) {
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    char *d = col + i * w;
    bool ended = false;
    for (uint32_t b = 0; b < w; ++b) { if (ended) d[b] = 0; else ended = d[b] == 0; }
  }
}

cudaError_t launch_zero_after_nul(void *col, uint64_t n, uint32_t w, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  const int grid = static_cast<int>(std::min<uint64_t>((n + 255) / 256, 148 * 8));
  k_zero_after_nul<<<grid, 256, 0, st>>>(static_cast<char *>(col), n, w);
  return cudaGetLastError();
}

__global__ void k_zero_date_padding(char *col, uint64_t n) {
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x)
    *reinterpret_cast<uint16_t *>(col + i * 8 + 6) = 0;
}

cudaError_t launch_zero_date_padding(void *col, uint64_t n, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  const int grid = static_cast<int>(std::min<uint64_t>((n + 255) / 256, 148 * 8));
  k_zero_date_padding<<<grid, 256, 0, st>>>(static_cast<char *>(col), n);
  return cudaGetLastError();
}

// ---- re-coding tables of dictionary-coded attributes -------------------------------------------------------
// Block dictionaries sit wherever the block builder put them: assemble values byte by byte.
template <class T>
__device__ __forceinline__ T load_unaligned(const char *p) {
  T v;
  char *d = reinterpret_cast<char *>(&v);
  for (uint32_t b = 0; b < sizeof(T); ++b) d[b] = p[b];
  return v;
}
template <class T>
__device__ __forceinline__ int cmp3(T a, T b) { return a < b ? -1 : a > b ? 1 : a == b ? 0 : 2; }

// Order of two dictionary entries in the attribute's own order (same as dict_compare on the host, lower.cu).
__device__ int dev_dict_compare(uint32_t qtype, uint32_t w, const char *a, const char *b) {
  switch (qtype) {
    case QS_INT: return cmp3(load_unaligned<int32_t>(a), load_unaligned<int32_t>(b));
    case QS_LONG: return cmp3(load_unaligned<int64_t>(a), load_unaligned<int64_t>(b));
    case QS_FLOAT: return cmp3(load_unaligned<float>(a), load_unaligned<float>(b));
    case QS_DOUBLE: return cmp3(load_unaligned<double>(a), load_unaligned<double>(b));
    case QS_DATE: {   // DateLit {int32 year; u8 month; u8 day}: lexicographic (types/DatetimeLit.hpp:65-93)
      const int y = cmp3(load_unaligned<int32_t>(a), load_unaligned<int32_t>(b));
      if (y) return y;
      const int m = cmp3(static_cast<unsigned char>(a[4]), static_cast<unsigned char>(b[4]));
      return m ? m : cmp3(static_cast<unsigned char>(a[5]), static_cast<unsigned char>(b[5]));
    }
    default:          // CHAR(w): strncmp
      for (uint32_t i = 0; i < w; ++i) {
        const unsigned char x = static_cast<unsigned char>(a[i]), y = static_cast<unsigned char>(b[i]);
        if (x != y) return x < y ? -1 : 1;
        if (x == 0) break;
      }
      return 0;
  }
}

// One CTA per coded stripe: entry e of the block's dictionary -> its position in the relation-wide dictionary
// (binary search; both are sorted).  A value the relation's dictionary does not hold raises the sticky error.
__global__ void __launch_bounds__(128) k_build_recode(const StageSeg *segs, uint32_t n_segs, uint32_t *error_flag) {
  for (uint32_t s = blockIdx.x; s < n_segs; s += gridDim.x) {
    const StageSeg sg = segs[s];
    if (sg.gdict == nullptr) continue;
    char *table = const_cast<char *>(sg.dict);
    for (uint32_t e = threadIdx.x; e < sg.dict_entries; e += blockDim.x) {
      const char *v = sg.bdict + static_cast<uint64_t>(e) * sg.bw;
      uint32_t lo = 0, hi = sg.g_entries;
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (dev_dict_compare(sg.qtype, sg.bw, sg.gdict + static_cast<uint64_t>(mid) * sg.bw, v) == -1) lo = mid + 1; else hi = mid;
      }
      if (lo >= sg.g_entries || dev_dict_compare(sg.qtype, sg.bw, sg.gdict + static_cast<uint64_t>(lo) * sg.bw, v) != 0) {
        atomicExch(error_flag, static_cast<uint32_t>(QSGPU_ERR_INVALID));
        lo = 0;
      }
      for (uint32_t b = 0; b < sg.vw; ++b) table[static_cast<uint64_t>(e) * sg.vw + b] = static_cast<char>(lo >> (8 * b));
    }
  }
}

cudaError_t launch_build_recode(const StageSeg *d_segs, uint32_t n_segs, uint32_t *error_flag, cudaStream_t st) {
  if (n_segs == 0) return cudaSuccess;
  k_build_recode<<<n_segs < 148u * 16 ? n_segs : 148u * 16, 128, 0, st>>>(d_segs, n_segs, error_flag);
  return cudaGetLastError();
}

cudaError_t launch_decode_segments(const StageSeg *d_segs, uint32_t n_segs, uint64_t n_tiles, int sm_count,
                                   cudaStream_t st) {
  if (n_segs == 0 || n_tiles == 0) return cudaSuccess;
  const uint64_t grid = n_tiles < static_cast<uint64_t>(sm_count) * 8 ? n_tiles : static_cast<uint64_t>(sm_count) * 8;
  k_decode_segments<<<static_cast<unsigned>(grid), 256, 0, st>>>(d_segs, n_segs, n_tiles);
  return cudaGetLastError();
}

static int grid_for(uint64_t n) {
  uint64_t g = (n + 255) / 256;
  if (g > 148ull * 8) g = 148ull * 8;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

cudaError_t launch_decode_dict(void *dst, const void *codes, const void *dict, uint64_t n, uint32_t code_width,
                               uint32_t value_width, uint32_t dict_entries, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  k_decode_dict<<<grid_for(n), 256, 0, st>>>(static_cast<char *>(dst), static_cast<const unsigned char *>(codes),
                                             static_cast<const char *>(dict), n, code_width, value_width,
                                             dict_entries);
  return cudaGetLastError();
}

cudaError_t launch_decode_truncated(void *dst, const void *codes, uint64_t n, uint32_t code_width,
                                    uint32_t value_width, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  k_decode_truncated<<<grid_for(n), 256, 0, st>>>(static_cast<char *>(dst),
                                                  static_cast<const unsigned char *>(codes), n, code_width,
                                                  value_width);
  return cudaGetLastError();
}

cudaError_t launch_decode_strided(void *dst, const void *slots, uint64_t n, uint32_t stride,
                                  uint32_t value_width, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  k_decode_strided<<<grid_for(n), 256, 0, st>>>(static_cast<char *>(dst), static_cast<const char *>(slots), n,
                                                stride, value_width);
  return cudaGetLastError();
}

}  // namespace qs
