// Host-side objects behind the opaque C-ABI handles.
#pragma once

#include <atomic>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "qs_ops.cuh"
#include "qsgpu.h"

// Relation-wide dictionary of an attribute held as codes (qsgpu_relation_set_dictionary).
struct qs_coded_attr {
  uint32_t cw = 0;                 // bytes per code in HBM: 1, 2 or 4; 0 = native column
  uint32_t n_entries = 0;
  char *d_dict = nullptr;          // device copy, readable for every code value of a 1/2-byte code
  std::vector<char> h_dict;        // host copy: literals are translated into code ranges at lowering time
};

struct qsgpu_relation {
  int dev = 0;
  std::vector<qs_attr> attrs;
  std::vector<char *> cols;
  std::vector<qs_coded_attr> coded;   // empty, or one per attribute
  uint32_t code_width(uint32_t a) const { return a < coded.size() ? coded[a].cw : 0u; }
  uint32_t stored_width(uint32_t a) const { const uint32_t c = code_width(a); return c ? c : attrs[a].width; }
  bool has_codes() const { for (const auto &c : coded) if (c.cw) return true; return false; }
  bool owns_memory = true;
  uint64_t capacity = 0;
  // Row count: authoritative copy lives on the device (kernels append with
  // atomics); host_rows is valid only while !dirty.
  unsigned long long *d_rows = nullptr;
  // Per-row NULL mask (bit j = attribute j is NULL), allocated only for relations that can hold NULLs
  // (the output of a LEFT OUTER join); absent = no NULLs.
  unsigned long long *d_nulls = nullptr;
  // attributes that may be NULL (bit j = attribute j; qsgpu_relation_set_nullable, or the columns a LEFT OUTER join /
  // an aggregate finalization can leave NULL).  Scans read the mask only for attributes in this set.
  uint64_t nullable_mask = 0;
  uint64_t host_rows = 0;
  bool dirty = false;
};

struct qsgpu_lip {
  int dev = 0;
  qs::LipDesc d{};
  uint32_t attr_type = QS_INT;
  uint64_t n_words = 0;
};

struct qsgpu_agg_state {
  // Work orders of one operator run concurrently on Worker threads and share the state (the reference
  // guards its states with SpinMutex / atomics, AggregationHandleSum.cpp:113).  Here the lock makes each
  // call's launches one unit in stream order: a scan's per-CTA partial rows must be folded by ITS merge
  // launch before the next scan overwrites them, and table growth must not race a scan.
  std::mutex mu;
  int dev = 0;
  uint32_t strategy = 0;
  // deep copy of the expression set
  std::vector<qs_node> nodes;
  std::string str_pool;
  qs_expr_set exprs{};
  int32_t predicate_root = -1;
  std::vector<qs_aggregate> aggregates;
  std::vector<int32_t> group_by_roots;
  std::vector<int> value_word;        // per aggregate: state word (0 = row count only)
  std::vector<int> nn_word;           // per aggregate: state word counting its non-NULL arguments (0 = the row count)
  std::vector<uint8_t> arg_vtype;     // per aggregate: VType of the argument (V_I32.. ; 0 for COUNT(*))
  std::vector<qs_attr> key_attrs;     // group-by attribute types
  std::vector<uint32_t> key_attr_ids;
  qs::AggDesc A{};
  char *ctl = nullptr;                // fixed-size strategies: the one block behind counters, directory, states, keys
  uint32_t max_ctas = 0;
  unsigned int *d_done = nullptr;     // last-CTA-merges ticket
  // dense export buffers (hash / collision-free partials, finalize index)
  uint64_t *d_idx = nullptr;
  uint64_t idx_cap = 0;
  unsigned long long *d_idx_count = nullptr;
  uint64_t *d_exp_states = nullptr, *d_exp_keys = nullptr;
  // COLLISION_FREE only: the table's existence map (CollisionFreeVectorTable::getExistenceMap), an exact
  // bit-vector filter over [0, cap) that BuildAggregationExistenceMap work orders fill; owned by the state
  qsgpu_lip *existence = nullptr;
  uint64_t exp_cap = 0;
  uint64_t estimated = 0;
  uint64_t rows_fed = 0;              // SEPARATE_CHAINING: rows handed to work orders + merged foreign groups (>= groups)
};

struct qsgpu_join_table {
  std::mutex mu;
  int dev = 0;
  uint32_t key_type = QS_INT;
  qs::JoinDesc J{};
  const qsgpu_relation *build_rel = nullptr;
  // Open addressing: slots are allocated from the optimizer's estimate, but the table is CLEARED and its mask
  // chosen at the first build work order, from the rows that work order really holds (2x, power of two): an
  // over-estimate must not spread 15 M entries over 2 GB of slots (every insert and probe is a random access,
  // and the smaller table is cleared 4x faster).  Later work orders that would push the load factor above 1/2
  // re-hash into a larger array (HashTable::resize); an under-estimate therefore grows instead of failing.
  uint64_t alloc_cap = 0;          // slots allocated
  uint64_t upper_entries = 0;      // rows handed to build work orders so far (>= entries)
  bool cleared = false;            // J.cap slots are initialised
  bool cap_frozen = false;         // qsgpu_join_partition already grouped rows by slices of J.cap
};

namespace qs {

constexpr size_t kReadScratchBytes = 256u << 10;

// Exact-size cache of large device blocks in front of the stream-ordered pool.  A query allocates the
// same temporaries (Select / join outputs, hash tables, block images) in the same sizes every time; the
// driver pool may split a bigger free block for a smaller request and then has to map fresh memory for the
// big one (15 ms for 1.9 GB, r01g profile).  Blocks >= 1 MB are therefore rounded to 2 MB multiples and,
// when freed, parked here for the next request of the same size.  Safe without events because every user
// of such a block runs on the device's single library stream.
struct BlockCache {
  std::mutex mu;
  std::unordered_map<void *, size_t> live;          // large blocks handed out -> rounded size
  std::multimap<size_t, void *> parked;             // freed, ready for reuse
  size_t parked_bytes = 0;
  static constexpr size_t kMinBytes = 1u << 20, kRound = 2u << 20;
  static constexpr size_t kMaxParkedBytes = 64ull << 30;
};

struct Device {
  int id = 0;
  std::shared_ptr<BlockCache> cache;
  // staging for qsgpu_relation_read_all of small results: device pack buffer + pinned landing buffer
  char *read_scratch = nullptr, *read_pinned = nullptr;
  std::shared_ptr<std::mutex> read_mu;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;   // host-to-device block images, overlapped with decoding (qsgpu_stage_blocks)
  int sm_count = 148;
  size_t smem_per_sm = 0, smem_per_block_optin = 0;
  uint32_t *d_error = nullptr;        // sticky device-side error word
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;       // per-kernel-family timing
  cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;   // qsgpu_timer_start/stop
  cudaMemPool_t pool = nullptr;       // stream-ordered arena for everything allocated per query
};

Device *device(int dev);               // nullptr (+ last error) when unavailable
// Stream-ordered allocation on the device the calling thread last resolved with device():
// ordered on that device's library stream, served from its retained pool (no device sync).
cudaError_t dev_malloc_bytes(void **p, size_t bytes);
// The relation's per-row NULL mask, allocated (all zeros, padded like a LONG column) on first need.
int ensure_null_mask(qsgpu_relation *rel, Device *d);
template <class T> inline cudaError_t dev_malloc(T **p, size_t bytes) {
  return dev_malloc_bytes(reinterpret_cast<void **>(p), bytes);
}
cudaError_t dev_free(void *p);
void set_error(int status, const std::string &msg);
int cuda_fail(cudaError_t e, const char *what);
void count_launch(int n = 1);
bool timing_enabled();
void record_ms(uint32_t family, float ms);

// Launch geometry for a scan over `n_cols` staged columns.
struct ScanPlan {
  int grid = 0;
  int ctas = 2;        // resident CTAs per SM the shared-memory budget allows
  size_t smem = 0;
};
int plan_scan(Device *d, ScanDesc *S, size_t extra_smem, ScanPlan *plan, int max_ctas = 4);

}  // namespace qs

#define QS_CUDA(call)                                                     \
  do {                                                                    \
    cudaError_t e__ = (call);                                             \
    if (e__ != cudaSuccess) return ::qs::cuda_fail(e__, #call);           \
  } while (0)
