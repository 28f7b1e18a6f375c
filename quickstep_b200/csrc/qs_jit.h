// The query compiler: turns one lowered work order (staged-column layout +
// VM program + operator descriptors) into a compile-time description `Q`,
// instantiates the matching kernel template of qs_kernels.cuh for it with
// NVRTC (sm_100a cubin) and caches the result in memory and on disk.
//
// The reference specialises its inner loops at build time by template
// instantiation over (type, operation, accessor) triples
// (types/operations/comparisons/Comparison-inl.hpp:283-431,
// types/operations/binary_operations/ArithmeticBinaryOperators.hpp:714-744);
// here the specialisation happens per query shape at first use.  Literal
// values, pointers, row ranges and table sizes stay run-time arguments, so
// e.g. TPC-H Q6 with other dates reuses the kernel.
#pragma once

#include <string>

#include "qs_host.h"

namespace qs {

enum JitFamily { JF_AGG = 0, JF_GROUPBY = 1, JF_SELECT = 2, JF_JOIN_BUILD = 3, JF_JOIN_PROBE = 4 };

struct JitKernel {
  cudaLibrary_t lib = nullptr;
  cudaKernel_t fn = nullptr;
  size_t smem_set[16] = {0};    // per device: largest dynamic smem size already opted in
  size_t occ_smem = ~static_cast<size_t>(0);   // dynamic smem size `occ` was computed for
  int occ = 0;                  // resident CTAs per SM at that size (cudaOccupancyMaxActiveBlocksPerMultiprocessor)
  std::string name;             // kernel symbol: family + hash of the description
};

// What the kernel is specialised on, besides the program itself.
struct JitSpec {
  JitFamily family;
  const ScanDesc *S;
  const Program *P;
  const AggDesc *A = nullptr;
  const SinkDesc *K = nullptr;
  const JoinDesc *J = nullptr;
  int hot = 1;                  // register-resident groups (JF_AGG)
  int priv = 0;                 // JF_AGG: the hot groups' accumulators are per-thread shared-memory slots, not registers
  int ctas_per_sm = 2;
};

// Full CUDA source of the kernel for `spec` (also the cache key).
std::string jit_source(const JitSpec &spec, std::string *kernel_name = nullptr);

// Compiles (or fetches) the kernel; returns a qsgpu_status and sets the
// thread's last error on failure.
int jit_get(const JitSpec &spec, JitKernel **out);

// Compile only (no device needed): used by build() and the CPU-side tests.
int jit_compile_only(const std::string &source, std::string *cubin, std::string *log);

cudaError_t jit_launch(JitKernel *k, int grid, size_t smem, cudaStream_t st, void **args);
// Resident CTAs per SM of the compiled kernel with `smem` bytes of dynamic shared memory (0 if unknown).
int jit_occupancy(JitKernel *k, size_t smem);

// Number of NVRTC compilations / disk-cache hits since load (diagnostics).
void jit_stats(uint64_t *compiled, uint64_t *disk_hits, uint64_t *mem_hits);

}  // namespace qs
