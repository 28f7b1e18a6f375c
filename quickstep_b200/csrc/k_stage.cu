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
