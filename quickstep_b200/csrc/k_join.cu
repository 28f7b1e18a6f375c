// K5 / K6 fixed kernel: clearing a join table.  Build and probe are
// qs_kernels.cuh join_build_body / join_probe_body, instantiated per query.
#include "qs_kernels.cuh"

namespace qs {

__global__ void k_join_clear(JoinSlot *slots, uint64_t cap) {
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < cap;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    slots[i].key = 0;
    slots[i].row = kEmptyRow;
  }
}

__global__ void k_fill_u64(unsigned long long *p, uint64_t n, unsigned long long v) {
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x)
    p[i] = v;
}

cudaError_t launch_fill_u64(unsigned long long *p, uint64_t n, unsigned long long v, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  uint64_t g = (n + 255) / 256;
  if (g > 148ull * 16) g = 148ull * 16;
  k_fill_u64<<<static_cast<unsigned>(g), 256, 0, st>>>(p, n, v);
  return cudaGetLastError();
}

// Re-insert every entry of an open-addressing table into a larger one (HashTable::resize,
// storage/HashTable.hpp: the reference resizes under an exclusive lock; here growth happens between work orders).
__global__ void k_join_rehash(const JoinSlot *from, uint64_t from_cap, JoinSlot *to, uint64_t to_cap) {
  const uint64_t mask = to_cap - 1;
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < from_cap;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const JoinSlot s = from[i];
    if (s.row == kEmptyRow) continue;
    uint64_t h = mix64(static_cast<uint64_t>(s.key)) & mask;
    for (;;) {
      uint64_t ok, orow;
      cas128(&to[h], 0ull, kEmptyRow, static_cast<uint64_t>(s.key), s.row, ok, orow);
      if (orow == kEmptyRow) break;
      h = (h + 1) & mask;
    }
  }
}

cudaError_t launch_join_rehash(const JoinSlot *from, uint64_t from_cap, JoinSlot *to, uint64_t to_cap, cudaStream_t st) {
  uint64_t g = (from_cap + 255) / 256;
  if (g > 148ull * 16) g = 148ull * 16;
  k_join_rehash<<<static_cast<unsigned>(g), 256, 0, st>>>(from, from_cap, to, to_cap);
  return cudaGetLastError();
}

cudaError_t launch_join_clear(const JoinDesc &J, cudaStream_t st) {
  if (J.dense) return launch_fill_u64(J.heads, J.cap, kEmptyRow, st);
  uint64_t g = (J.cap + 255) / 256;
  if (g > 148ull * 16) g = 148ull * 16;
  k_join_clear<<<static_cast<unsigned>(g), 256, 0, st>>>(J.slots, J.cap);
  return cudaGetLastError();
}

}  // namespace qs
