// Multi-GPU collectives behind the C ABI (include/qsgpu.h, "multi-GPU"): one process per GPU, NCCL over
// NVLink / NVSwitch.  What crosses GPUs on this path:
//   * partial aggregation states            AggregationHandle*::mergeStates
//                                           (expressions/aggregation/AggregationHandleSum.cpp:109-117),
//                                           ThreadPrivateCompactKeyHashTable::mergeFrom
//                                           (storage/ThreadPrivateCompactKeyHashTable.cpp:306-363)
//   * LIP filter bit words (OR)             one filter per query in the reference (QueryContext.cpp:66-97);
//                                           here every GPU fills its copy from its share of the build side
//   * rows of a small relation (all-gather) the broadcast build side of a hash join, top-k candidates
// The reference instantiates one state / hash table per partition inside ONE process
// (query_execution/QueryContext.cpp:66-97); "partition id <-> device id" is the mapping used here.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2): a process that already loaded it (torch) shares that
// copy, and a single-GPU process never needs it.  Every collective is queued on the library stream of the
// device, behind the kernels that produce its input, and is followed by ONE kernel that consumes the gathered
// buffer: the host never waits between the scan and the finalize.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

#include "qs_host.h"
#include "qs_kernels.cuh"

struct qsgpu_comm {
  int dev = 0, rank = 0, n_ranks = 1;
  ncclComm_t comm = nullptr;
  std::mutex mu;                       // one collective at a time per communicator
  char *scratch = nullptr;             // receive buffer of the gather-type collectives
  size_t scratch_bytes = 0;
  unsigned long long *d_counts = nullptr;     // [n_ranks] row counts on the device
  unsigned long long *h_counts = nullptr;     // pinned landing buffer
  // Peer-memory mailbox (CUDA IPC over NVLink / NVSwitch), the latency path of the small collectives: every rank maps
  // every other rank's mailbox and WRITES its contribution straight into it from the merging kernel; no NCCL launch,
  // no proxy thread, one kernel per collective.  Layout (all ranks alike):
  //   [flags: 2 parities x kMaxMergeRanks x u64, padded to 1 KB][parity 0: n_ranks slots of kMailSlotBytes][parity 1: ...]
  // nullptr on every rank when any rank could not map its peers (NCCL carries everything then).
  char *mailbox = nullptr;
  char *peer_mailbox[16] = {nullptr};         // peer_mailbox[rank] = own mailbox; others are IPC mappings
  char **d_peer_mailbox = nullptr;            // the same table on the device
  unsigned long long epoch = 0;               // collectives done through the mailbox (same on all ranks)
};

namespace qs {

namespace {

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;
std::mutex g_nccl_mu;

int load_nccl() {
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  if (g_nccl.handle) return QSGPU_OK;
  const char *names[] = {std::getenv("QSGPU_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void *h = nullptr;
  for (const char *n : names) {
    if (!n || !*n) continue;
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) {
    set_error(QSGPU_ERR_UNSUPPORTED, "libnccl.so.2 not found (set QSGPU_NCCL_LIB): the multi-GPU entry points need NCCL");
    return QSGPU_ERR_UNSUPPORTED;
  }
  bool ok = true;
  auto sym = [&](const char *name) { void *p = dlsym(h, name); ok = ok && p != nullptr; return p; };
  NcclApi a;
  a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
  a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
  a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
  a.AllGather = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
  a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
  a.Broadcast = reinterpret_cast<decltype(a.Broadcast)>(sym("ncclBroadcast"));
  a.Send = reinterpret_cast<decltype(a.Send)>(sym("ncclSend"));
  a.Recv = reinterpret_cast<decltype(a.Recv)>(sym("ncclRecv"));
  a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
  a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
  a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
  if (!ok) { set_error(QSGPU_ERR_UNSUPPORTED, "libnccl.so.2 lacks a required symbol"); return QSGPU_ERR_UNSUPPORTED; }
  a.handle = h;
  g_nccl = a;
  return QSGPU_OK;
}

int nccl_fail(ncclResult_t r, const char *what) {
  set_error(QSGPU_ERR_CUDA, std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error"));
  return QSGPU_ERR_CUDA;
}
#define QS_NCCL(call)                                                  \
  do {                                                                 \
    ncclResult_t r__ = (call);                                         \
    if (r__ != ncclSuccess) return nccl_fail(r__, #call);              \
  } while (0)

int ensure_scratch(qsgpu_comm *c, size_t bytes) {
  if (bytes <= c->scratch_bytes) return QSGPU_OK;
  if (c->scratch) dev_free(c->scratch);
  c->scratch = nullptr;
  c->scratch_bytes = 0;
  const size_t want = std::max<size_t>(bytes, 1u << 20);
  QS_CUDA(dev_malloc(&c->scratch, want));
  c->scratch_bytes = want;
  return QSGPU_OK;
}

constexpr uint32_t kMaxMergeRanks = 16;
constexpr size_t kMailFlagBytes = 1024;                 // 2 x 16 x 8 = 256 used
constexpr size_t kMailSlotBytes = 64u << 10;            // one rank's contribution: 256 groups x (13 + 1) words = 28 KB at most
inline size_t mailbox_bytes(int n_ranks) { return kMailFlagBytes + 2 * static_cast<size_t>(n_ranks) * kMailSlotBytes; }

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// Wait until a peer has raised `flag` to `epoch`.  A peer that never arrives (its process died, or it took another
// path through the plan) must not wedge this GPU: after kPeerWaitNs the kernel gives up, raises QSGPU_ERR_CUDA in the
// device error word (reported at the next read of a result) and goes on with whatever is in the mailbox.
constexpr unsigned long long kPeerWaitNs = 20ull * 1000 * 1000 * 1000;
__device__ __forceinline__ void wait_peer_flag(const unsigned long long *flag, unsigned long long epoch, uint32_t *error_flag) {
  if (ld_acquire_sys(flag) == epoch) return;
  const unsigned long long t0 = global_timer_ns();
  uint32_t spins = 0;
  while (ld_acquire_sys(flag) != epoch) {
    __nanosleep(20);
    if ((++spins & 0xfffu) == 0 && global_timer_ns() - t0 > kPeerWaitNs) {
      atomicExch(error_flag, static_cast<uint32_t>(QSGPU_ERR_CUDA));
      return;
    }
  }
}

// ---- kernels ---------------------------------------------------------------------------------------------
// Fold the gathered [states | keys] blocks of all ranks, IN RANK ORDER, into this rank's state (the own block
// is part of `gathered`, so the state is rebuilt from scratch): every rank computes the same additions in the
// same order and ends with bit-identical totals.  One CTA; partial_rows <= 256 threads do the work.
__global__ void __launch_bounds__(256) k_merge_gathered_compact(const __grid_constant__ AggDesc A, const uint64_t *gathered,
                                                                uint32_t n_ranks);

// The fold of k_merge_gathered_compact, shared with the peer-memory form below.
__device__ __forceinline__ void fold_gathered_compact(const AggDesc &A, const uint64_t *gathered, uint64_t slot_words, uint32_t n_ranks,
                                                      short (*inv)[kCompactMaxGroups]) {
  const uint32_t t = threadIdx.x;
  const uint32_t rows = A.partial_rows, W = A.words;
  for (uint32_t i = t; i < kMaxMergeRanks * kCompactMaxGroups; i += blockDim.x) (&inv[0][0])[i] = -1;
  __syncthreads();
  for (uint32_t r = 0; r < n_ranks; ++r) {
    const uint64_t *st = gathered + r * slot_words;
    const uint64_t *keys = st + static_cast<uint64_t>(rows) * W;
    if (t < rows && __ldcg(&st[static_cast<uint64_t>(t) * W]) != 0) {
      const int gid = A.n_key_cols == 0 ? 0 : dir_insert(__ldcg(&keys[t]), A);
      if (gid >= 0) inv[r][gid] = static_cast<short>(t);
    }
    __syncthreads();
  }
  const uint32_t n_groups = A.n_key_cols == 0 ? 1u : min(*reinterpret_cast<volatile uint32_t *>(A.n_groups), rows);
  if (t < n_groups) {
    for (uint32_t w = 0; w < W; ++w) {
      const uint8_t kind = w == 0 ? AK_SUM_I64 : A.kind[w - 1];
      uint64_t x = w == 0 ? 0 : agg_identity(kind);
      for (uint32_t r = 0; r < n_ranks; ++r) {
        const int g = inv[r][t];
        if (g >= 0) x = agg_combine(kind, x, __ldcg(&gathered[r * slot_words + static_cast<uint64_t>(g) * W + w]));
      }
      A.states[static_cast<uint64_t>(t) * W + w] = x;
    }
  }
}

__global__ void __launch_bounds__(256) k_merge_gathered_compact(const __grid_constant__ AggDesc A, const uint64_t *gathered,
                                                                uint32_t n_ranks) {
  __shared__ short inv[kMaxMergeRanks][kCompactMaxGroups];     // inv[r][local group id] = row of rank r's block, -1 = none
  fold_gathered_compact(A, gathered, static_cast<uint64_t>(A.partial_rows) * (A.words + 1), n_ranks, inv);
}

/*
 * Merge of the fixed-size aggregation states over PEER MEMORY: all-gather and fold in ONE kernel, no NCCL call.
 *   1. this rank's [states | keys] block (<= 28 KB) is stored into slot `rank` of EVERY rank's mailbox -- plain
 *      stores to addresses that live in the peers' HBM, carried by NVLink (NVSwitch gives every pair full bandwidth);
 *   2. a system-scope release store raises this rank's flag in every mailbox to the collective's epoch;
 *   3. the kernel waits (system-scope acquire loads of its OWN flags) until every rank's block has landed here;
 *   4. the blocks are folded in rank order exactly like k_merge_gathered_compact: bit-identical totals on all ranks.
 * Two parities of slots and flags alternate: a rank can start collective e + 2 only after every peer has raised its
 * flag for e + 1, which a peer does after it has finished reading the blocks of e -- so a slot is never overwritten
 * while somebody still reads it.  Replaces ncclAllGather + k_merge_gathered_compact (two launches and the NCCL kernel's
 * own rendezvous, ~45 us at 8 GPUs) by one ~10 us kernel: what matters for a 0.5 ms query.
 */
__global__ void __launch_bounds__(256) k_merge_peer_compact(const __grid_constant__ AggDesc A, char *const *peers, uint32_t rank,
                                                            uint32_t n_ranks, unsigned long long epoch) {
  __shared__ short inv[kMaxMergeRanks][kCompactMaxGroups];
  const uint32_t t = threadIdx.x;
  const uint32_t parity = static_cast<uint32_t>(epoch & 1ull);
  const uint64_t block_words = static_cast<uint64_t>(A.partial_rows) * (A.words + 1);
  const size_t slot_off = kMailFlagBytes + (static_cast<size_t>(parity) * n_ranks + rank) * kMailSlotBytes;
  for (uint32_t r = 0; r < n_ranks; ++r) {
    uint64_t *dst = reinterpret_cast<uint64_t *>(peers[r] + slot_off);
    for (uint64_t i = t; i < block_words; i += blockDim.x) dst[i] = A.states[i];
  }
  __threadfence_system();
  __syncthreads();
  if (t < n_ranks)
    st_release_sys(reinterpret_cast<unsigned long long *>(peers[t]) + parity * kMaxMergeRanks + rank, epoch);
  if (t < n_ranks)
    wait_peer_flag(reinterpret_cast<const unsigned long long *>(peers[rank]) + parity * kMaxMergeRanks + t, epoch, A.error_flag);
  __syncthreads();
  const uint64_t *gathered = reinterpret_cast<const uint64_t *>(peers[rank] + kMailFlagBytes + static_cast<size_t>(parity) * n_ranks * kMailSlotBytes);
  fold_gathered_compact(A, gathered, kMailSlotBytes / 8, n_ranks, inv);
}

/*
 * All-gather of a SMALL relation (the top-k rows of every rank, a few hundred bytes) over peer memory, again one kernel
 * and no host wait: the rank's row count may still be device-only (the relation was just produced), so the kernel reads
 * it, stores [count | column 0 | column 1 ...] (fixed capacity per column) into its slot of every peer's mailbox, raises
 * its flag, waits for the others, and copies the ranks' rows -- in rank order -- into the output relation, whose row
 * count it sets.  The NCCL form needs the counts on the HOST first (one synchronisation) and one broadcast per
 * (rank, attribute).
 */
struct PeerGatherDesc {
  uint32_t n_cols, cap;                 // rows every rank may contribute
  uint32_t width[kMaxCols + 1];
  uint32_t col_off[kMaxCols + 1];       // byte offset of the column inside a slot (8-byte aligned), after the count word
  const char *in[kMaxCols + 1];
  char *out[kMaxCols + 1];
  const unsigned long long *in_rows;    // device row count of the local relation
  unsigned long long *out_rows;
  uint32_t *error_flag;
};

__global__ void __launch_bounds__(256) k_allgather_peer(const __grid_constant__ PeerGatherDesc G, char *const *peers, uint32_t rank,
                                                        uint32_t n_ranks, unsigned long long epoch) {
  __shared__ unsigned long long s_first[kMaxMergeRanks + 1];
  const uint32_t t = threadIdx.x;
  const uint32_t parity = static_cast<uint32_t>(epoch & 1ull);
  const size_t my_slot = kMailFlagBytes + (static_cast<size_t>(parity) * n_ranks + rank) * kMailSlotBytes;
  const unsigned long long have = *G.in_rows;
  if (have > G.cap && t == 0) atomicExch(G.error_flag, static_cast<uint32_t>(QSGPU_ERR_CAPACITY));   // more rows than the caller promised
  const unsigned long long mine = min(have, static_cast<unsigned long long>(G.cap));
  for (uint32_t r = 0; r < n_ranks; ++r) {
    char *dst = peers[r] + my_slot;
    if (t == 0) *reinterpret_cast<unsigned long long *>(dst) = mine;
    for (uint32_t c = 0; c < G.n_cols; ++c) {
      const uint64_t bytes = mine * G.width[c];
      for (uint64_t b = t; b < bytes; b += blockDim.x) dst[8 + G.col_off[c] + b] = G.in[c][b];
    }
  }
  __threadfence_system();
  __syncthreads();
  if (t < n_ranks)
    st_release_sys(reinterpret_cast<unsigned long long *>(peers[t]) + parity * kMaxMergeRanks + rank, epoch);
  if (t < n_ranks)
    wait_peer_flag(reinterpret_cast<const unsigned long long *>(peers[rank]) + parity * kMaxMergeRanks + t, epoch, G.error_flag);
  __syncthreads();
  const char *base = peers[rank] + kMailFlagBytes + static_cast<size_t>(parity) * n_ranks * kMailSlotBytes;
  if (t == 0) {
    unsigned long long acc = 0;
    for (uint32_t r = 0; r < n_ranks; ++r) {
      s_first[r] = acc;
      acc += __ldcg(reinterpret_cast<const unsigned long long *>(base + static_cast<size_t>(r) * kMailSlotBytes));
    }
    s_first[n_ranks] = acc;
    *G.out_rows = acc;
  }
  __syncthreads();
  for (uint32_t r = 0; r < n_ranks; ++r) {
    const char *src = base + static_cast<size_t>(r) * kMailSlotBytes + 8;
    const unsigned long long n = s_first[r + 1] - s_first[r];
    for (uint32_t c = 0; c < G.n_cols; ++c) {
      const uint64_t bytes = n * G.width[c];
      char *dst = G.out[c] + s_first[r] * G.width[c];
      for (uint64_t b = t; b < bytes; b += blockDim.x) dst[b] = __ldcg(src + G.col_off[c] + b);
    }
  }
}

// words[i] = OR over ranks of gathered[r][i]  (LIP filter bit words; BarrieredReadWriteConcurrentBitVector layout)
__global__ void k_or_gathered(uint64_t *words, const uint64_t *gathered, uint64_t n_words, uint32_t n_ranks) {
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n_words;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    uint64_t x = 0;
    for (uint32_t r = 0; r < n_ranks; ++r) x |= gathered[r * n_words + i];
    words[i] = x;
  }
}

// dst[i] |= src[r][i] for the received chunks of the reduce-scatter form (large filters)
__global__ void k_or_chunks(uint64_t *dst, const uint64_t *src, uint64_t n_words, uint32_t n_src) {
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n_words;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    uint64_t x = dst[i];
    for (uint32_t r = 0; r < n_src; ++r) x |= src[r * n_words + i];
    dst[i] = x;
  }
}

}  // namespace
}  // namespace qs

using namespace qs;

extern "C" {

int qsgpu_comm_unique_id(qs_comm_id *id) {
  static_assert(sizeof(qs_comm_id) == sizeof(ncclUniqueId), "qs_comm_id must hold an ncclUniqueId");
  int st = load_nccl();
  if (st) return st;
  ncclUniqueId u;
  QS_NCCL(g_nccl.GetUniqueId(&u));
  std::memcpy(id->bytes, &u, sizeof(u));
  return QSGPU_OK;
}

// Allocates this rank's mailbox, exchanges the CUDA IPC handles (an NCCL all-gather of 64 bytes per rank: the
// communicator exists already) and maps the peers'.  A rank that cannot map a peer (no P2P path, another node,
// QSGPU_PEER_MERGE=0) reports it, and the MINIMUM over ranks decides: either every rank uses the mailbox or none.
static int setup_mailbox(qsgpu_comm *c, Device *d) {
  const char *env = std::getenv("QSGPU_PEER_MERGE");
  int ok = (env && env[0] == '0') ? 0 : 1;
  const int R = c->n_ranks;
  cudaIpcMemHandle_t mine;
  std::memset(&mine, 0, sizeof(mine));
  if (ok) {
    if (cudaMalloc(&c->mailbox, mailbox_bytes(R)) != cudaSuccess) { cudaGetLastError(); c->mailbox = nullptr; ok = 0; }
    else if (cudaMemsetAsync(c->mailbox, 0, mailbox_bytes(R), d->stream) != cudaSuccess || cudaIpcGetMemHandle(&mine, c->mailbox) != cudaSuccess) { cudaGetLastError(); ok = 0; }
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  char *d_handles = nullptr;
  QS_CUDA(cudaMalloc(&d_handles, 64 * static_cast<size_t>(R + 1)));
  QS_CUDA(cudaMemcpyAsync(d_handles + 64 * static_cast<size_t>(R), &mine, 64, cudaMemcpyHostToDevice, d->stream));
  QS_NCCL(g_nccl.AllGather(d_handles + 64 * static_cast<size_t>(R), d_handles, 64, ncclChar, c->comm, d->stream));
  std::vector<cudaIpcMemHandle_t> all(static_cast<size_t>(R));
  QS_CUDA(cudaMemcpyAsync(all.data(), d_handles, 64 * static_cast<size_t>(R), cudaMemcpyDeviceToHost, d->stream));
  QS_CUDA(cudaStreamSynchronize(d->stream));
  cudaFree(d_handles);
  for (int r = 0; r < R && ok; ++r) {
    if (r == c->rank) { c->peer_mailbox[r] = c->mailbox; continue; }
    void *p = nullptr;
    if (cudaIpcOpenMemHandle(&p, all[static_cast<size_t>(r)], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
    c->peer_mailbox[r] = static_cast<char *>(p);
  }
  // agreement: min over ranks
  c->h_counts[0] = static_cast<unsigned long long>(ok);
  QS_CUDA(cudaMemcpyAsync(c->d_counts, c->h_counts, 8, cudaMemcpyHostToDevice, d->stream));
  QS_NCCL(g_nccl.AllReduce(c->d_counts, c->d_counts, 1, ncclUint64, ncclMin, c->comm, d->stream));
  QS_CUDA(cudaMemcpyAsync(c->h_counts, c->d_counts, 8, cudaMemcpyDeviceToHost, d->stream));
  QS_CUDA(cudaStreamSynchronize(d->stream));
  if (c->h_counts[0] == 0) {
    for (int r = 0; r < R; ++r) if (r != c->rank && c->peer_mailbox[r]) cudaIpcCloseMemHandle(c->peer_mailbox[r]);
    std::memset(c->peer_mailbox, 0, sizeof(c->peer_mailbox));
    if (c->mailbox) cudaFree(c->mailbox);
    c->mailbox = nullptr;
    return QSGPU_OK;
  }
  QS_CUDA(cudaMalloc(&c->d_peer_mailbox, sizeof(char *) * kMaxMergeRanks));
  QS_CUDA(cudaMemcpyAsync(c->d_peer_mailbox, c->peer_mailbox, sizeof(char *) * kMaxMergeRanks, cudaMemcpyHostToDevice, d->stream));
  QS_CUDA(cudaStreamSynchronize(d->stream));
  return QSGPU_OK;
}

int qsgpu_comm_create(int dev, int rank, int n_ranks, const qs_comm_id *id, qsgpu_comm_t *out) {
  Device *d = device(dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  if (n_ranks < 1 || rank < 0 || rank >= n_ranks || !id) { set_error(QSGPU_ERR_INVALID, "bad rank / n_ranks / id"); return QSGPU_ERR_INVALID; }
  int st = load_nccl();
  if (st) return st;
  std::unique_ptr<qsgpu_comm> c(new qsgpu_comm);
  c->dev = dev; c->rank = rank; c->n_ranks = n_ranks;
  ncclUniqueId u;
  std::memcpy(&u, id->bytes, sizeof(u));
  QS_NCCL(g_nccl.CommInitRank(&c->comm, n_ranks, u, rank));
  QS_CUDA(cudaMalloc(&c->d_counts, sizeof(unsigned long long) * static_cast<size_t>(n_ranks) * 2));
  QS_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&c->h_counts), sizeof(unsigned long long) * static_cast<size_t>(n_ranks) * 2, cudaHostAllocPortable));
  if (n_ranks > 1 && static_cast<uint32_t>(n_ranks) <= kMaxMergeRanks) {
    st = setup_mailbox(c.get(), d);
    if (st) return st;
  }
  *out = c.release();
  return QSGPU_OK;
}

int qsgpu_comm_destroy(qsgpu_comm_t c) {
  if (!c) return QSGPU_OK;
  Device *d = device(c->dev);
  if (d) cudaStreamSynchronize(d->stream);
  if (c->scratch) dev_free(c->scratch);
  if (c->mailbox) {
    // peers may still be inside their last mailbox collective: leave together
    if (g_nccl.AllReduce && c->comm && d) { g_nccl.AllReduce(c->d_counts, c->d_counts, 1, ncclUint64, ncclMax, c->comm, d->stream); cudaStreamSynchronize(d->stream); }
    for (int r = 0; r < c->n_ranks; ++r) if (r != c->rank && c->peer_mailbox[r]) cudaIpcCloseMemHandle(c->peer_mailbox[r]);
    cudaFree(c->mailbox);
    if (c->d_peer_mailbox) cudaFree(c->d_peer_mailbox);
  }
  if (c->d_counts) cudaFree(c->d_counts);
  if (c->h_counts) cudaFreeHost(c->h_counts);
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  delete c;
  return QSGPU_OK;
}

int qsgpu_comm_rank(qsgpu_comm_t c, int *rank, int *n_ranks) {
  if (!c) { if (rank) *rank = 0; if (n_ranks) *n_ranks = 1; return QSGPU_OK; }
  if (rank) *rank = c->rank;
  if (n_ranks) *n_ranks = c->n_ranks;
  return QSGPU_OK;
}

int qsgpu_comm_peer_memory(qsgpu_comm_t c, int *enabled) {
  *enabled = (c && c->mailbox) ? 1 : 0;
  return QSGPU_OK;
}

int qsgpu_comm_barrier(qsgpu_comm_t c) {
  if (!c || c->n_ranks == 1) return QSGPU_OK;
  std::lock_guard<std::mutex> lk(c->mu);
  Device *d = device(c->dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  QS_NCCL(g_nccl.AllReduce(c->d_counts, c->d_counts, 1, ncclUint64, ncclMax, c->comm, d->stream));
  QS_CUDA(cudaStreamSynchronize(d->stream));
  return QSGPU_OK;
}

int qsgpu_comm_allreduce_i64(qsgpu_comm_t c, int64_t *values, uint32_t n, uint32_t op) {
  if (!c || c->n_ranks == 1 || n == 0) return QSGPU_OK;
  if (op > 2) { set_error(QSGPU_ERR_INVALID, "op: 0 = sum, 1 = min, 2 = max"); return QSGPU_ERR_INVALID; }
  std::lock_guard<std::mutex> lk(c->mu);
  Device *d = device(c->dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  int64_t *dbuf = nullptr;
  QS_CUDA(dev_malloc(&dbuf, static_cast<size_t>(n) * 8));
  QS_CUDA(cudaMemcpyAsync(dbuf, values, static_cast<size_t>(n) * 8, cudaMemcpyHostToDevice, d->stream));
  QS_NCCL(g_nccl.AllReduce(dbuf, dbuf, n, ncclInt64, op == 0 ? ncclSum : op == 1 ? ncclMin : ncclMax, c->comm, d->stream));
  QS_CUDA(cudaMemcpyAsync(values, dbuf, static_cast<size_t>(n) * 8, cudaMemcpyDeviceToHost, d->stream));
  QS_CUDA(cudaStreamSynchronize(d->stream));
  dev_free(dbuf);
  return QSGPU_OK;
}

// rows of every rank (host), via an all-gather of one word per rank
static int gather_counts(qsgpu_comm *c, Device *d, unsigned long long mine, std::vector<uint64_t> *counts) {
  c->h_counts[c->n_ranks] = mine;
  QS_CUDA(cudaMemcpyAsync(c->d_counts + c->n_ranks, c->h_counts + c->n_ranks, 8, cudaMemcpyHostToDevice, d->stream));
  QS_NCCL(g_nccl.AllGather(c->d_counts + c->n_ranks, c->d_counts, 1, ncclUint64, c->comm, d->stream));
  QS_CUDA(cudaMemcpyAsync(c->h_counts, c->d_counts, 8 * static_cast<size_t>(c->n_ranks), cudaMemcpyDeviceToHost, d->stream));
  QS_CUDA(cudaStreamSynchronize(d->stream));
  counts->assign(c->h_counts, c->h_counts + c->n_ranks);
  return QSGPU_OK;
}

int qsgpu_agg_merge_all(qsgpu_agg_state_t state, qsgpu_comm_t c) {
  if (!c || c->n_ranks == 1) return QSGPU_OK;
  std::lock_guard<std::mutex> lk(c->mu);
  Device *d = device(state->dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  if (state->dev != c->dev) { set_error(QSGPU_ERR_INVALID, "state and communicator on different devices"); return QSGPU_ERR_INVALID; }
  const AggDesc &A = state->A;
  if (state->strategy == QS_AGG_SINGLE_STATE || state->strategy == QS_AGG_COMPACT_KEY) {
    std::lock_guard<std::mutex> state_lock(state->mu);
    if (static_cast<uint32_t>(c->n_ranks) > kMaxMergeRanks) { set_error(QSGPU_ERR_UNSUPPORTED, "more than 16 ranks"); return QSGPU_ERR_UNSUPPORTED; }
    // the state's [states | keys] block is contiguous (qsgpu_agg_create): gathered as it lies
    const size_t block_words = static_cast<size_t>(A.partial_rows) * (A.words + 1);
    if (c->mailbox && block_words * 8 <= kMailSlotBytes) {
      // peer-memory form: gather and fold in one kernel (see k_merge_peer_compact)
      ++c->epoch;
      k_merge_peer_compact<<<1, 256, 0, d->stream>>>(A, c->d_peer_mailbox, static_cast<uint32_t>(c->rank), static_cast<uint32_t>(c->n_ranks), c->epoch);
      QS_CUDA(cudaGetLastError());
      count_launch();
      return QSGPU_OK;
    }
    int st = ensure_scratch(c, block_words * 8 * c->n_ranks);
    if (st) return st;
    QS_NCCL(g_nccl.AllGather(A.states, c->scratch, block_words, ncclUint64, c->comm, d->stream));
    k_merge_gathered_compact<<<1, 256, 0, d->stream>>>(A, reinterpret_cast<const uint64_t *>(c->scratch), static_cast<uint32_t>(c->n_ranks));
    QS_CUDA(cudaGetLastError());
    count_launch();
    return QSGPU_OK;
  }
  // hash / dense tables: export the live groups, all-gather them padded to the largest rank, fold the foreign ones
  void *d_states = nullptr, *d_keys = nullptr;
  uint64_t n = 0;
  uint32_t words = 0, kw = 0;
  int st = qsgpu_agg_partial(state, &d_states, &d_keys, &n, &words, &kw);
  if (st) return st;
  std::vector<uint64_t> counts;
  st = gather_counts(c, d, n, &counts);
  if (st) return st;
  uint64_t mx = 0;
  for (uint64_t x : counts) mx = std::max(mx, x);
  if (mx == 0) return QSGPU_OK;
  const size_t row_words = static_cast<size_t>(words) + kw;
  st = ensure_scratch(c, (static_cast<size_t>(c->n_ranks) + 1) * mx * row_words * 8);
  if (st) return st;
  uint64_t *send = reinterpret_cast<uint64_t *>(c->scratch) + static_cast<size_t>(c->n_ranks) * mx * row_words;
  QS_CUDA(cudaMemcpyAsync(send, d_states, n * words * 8, cudaMemcpyDeviceToDevice, d->stream));
  QS_CUDA(cudaMemcpyAsync(send + mx * words, d_keys, n * kw * 8, cudaMemcpyDeviceToDevice, d->stream));
  QS_NCCL(g_nccl.AllGather(send, c->scratch, mx * row_words, ncclUint64, c->comm, d->stream));
  for (int r = 0; r < c->n_ranks; ++r) {
    if (r == c->rank || counts[r] == 0) continue;
    const uint64_t *blk = reinterpret_cast<const uint64_t *>(c->scratch) + static_cast<size_t>(r) * mx * row_words;
    st = qsgpu_agg_merge_partial(state, blk, blk + mx * words, counts[r]);
    if (st) return st;
  }
  return QSGPU_OK;
}

int qsgpu_lip_allreduce(qsgpu_lip_t lip, qsgpu_comm_t c) {
  if (!c || c->n_ranks == 1) return QSGPU_OK;
  std::lock_guard<std::mutex> lk(c->mu);
  Device *d = device(lip->dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  const uint64_t nw = lip->n_words;
  const uint32_t R = static_cast<uint32_t>(c->n_ranks);
  if (nw * 8 <= (8u << 20) || nw < R) {
    // small filter (Q3's customer filter: 188 KB at SF10): latency-bound, one all-gather + one OR kernel
    int st = ensure_scratch(c, nw * 8 * R);
    if (st) return st;
    QS_NCCL(g_nccl.AllGather(lip->d.words, c->scratch, nw, ncclUint64, c->comm, d->stream));
    const int grid = static_cast<int>(std::min<uint64_t>((nw + 255) / 256, 148 * 8));
    k_or_gathered<<<std::max(grid, 1), 256, 0, d->stream>>>(lip->d.words, reinterpret_cast<const uint64_t *>(c->scratch), nw, R);
    QS_CUDA(cudaGetLastError());
    count_launch();
    return QSGPU_OK;
  }
  // large filter: bandwidth-optimal form.  Rank j owns words [j*chunk, (j+1)*chunk): every rank sends chunk j of
  // its filter to rank j (one grouped exchange), ORs what it received into its own chunk, and the reduced chunks
  // are all-gathered in place.
  const uint64_t chunk = (nw + R - 1) / R;
  auto len = [&](uint32_t j) { const uint64_t b = j * chunk; return b >= nw ? 0ull : std::min<uint64_t>(chunk, nw - b); };
  int st = ensure_scratch(c, chunk * 8 * R);
  if (st) return st;
  uint64_t *recv = reinterpret_cast<uint64_t *>(c->scratch);
  const uint64_t my_len = len(c->rank);
  QS_NCCL(g_nccl.GroupStart());
  uint32_t slot = 0;
  for (uint32_t p = 0; p < R; ++p) {
    if (static_cast<int>(p) == c->rank) continue;
    if (len(p)) QS_NCCL(g_nccl.Send(lip->d.words + p * chunk, len(p), ncclUint64, static_cast<int>(p), c->comm, d->stream));
    if (my_len) QS_NCCL(g_nccl.Recv(recv + static_cast<uint64_t>(slot) * my_len, my_len, ncclUint64, static_cast<int>(p), c->comm, d->stream));
    ++slot;
  }
  QS_NCCL(g_nccl.GroupEnd());
  if (my_len) {
    const int grid = static_cast<int>(std::min<uint64_t>((my_len + 255) / 256, 148 * 8));
    k_or_chunks<<<std::max(grid, 1), 256, 0, d->stream>>>(lip->d.words + c->rank * chunk, recv, my_len, R - 1);
    QS_CUDA(cudaGetLastError());
    count_launch();
  }
  QS_NCCL(g_nccl.GroupStart());
  for (uint32_t p = 0; p < R; ++p)
    if (len(p)) QS_NCCL(g_nccl.Broadcast(lip->d.words + p * chunk, lip->d.words + p * chunk, len(p), ncclUint64, static_cast<int>(p), c->comm, d->stream));
  QS_NCCL(g_nccl.GroupEnd());
  return QSGPU_OK;
}

int qsgpu_relation_allgather(qsgpu_relation_t local, qsgpu_comm_t c, qsgpu_relation_t *out) {
  if (!local || !out) { set_error(QSGPU_ERR_INVALID, "null relation"); return QSGPU_ERR_INVALID; }
  Device *d = device(local->dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  if (local->has_codes()) { set_error(QSGPU_ERR_UNSUPPORTED, "all-gather of a dictionary-coded relation"); return QSGPU_ERR_UNSUPPORTED; }
  uint64_t mine = 0;
  int st = qsgpu_relation_num_rows(local, &mine);      // the row count of a temporary may still be device-only
  if (st) return st;
  const int R = c ? c->n_ranks : 1;
  std::vector<uint64_t> counts(1, mine);
  std::unique_lock<std::mutex> lk;
  if (R > 1) {
    lk = std::unique_lock<std::mutex>(c->mu);
    st = gather_counts(c, d, mine, &counts);
    if (st) return st;
  }
  uint64_t total = 0;
  for (uint64_t x : counts) total += x;
  qsgpu_relation *rel = nullptr;
  st = qsgpu_relation_create(local->dev, static_cast<uint32_t>(local->attrs.size()), local->attrs.data(), std::max<uint64_t>(total, 1), &rel);
  if (st) return st;
  // NULL-able attributes: the per-row masks travel like one more column.  Every rank must take the same branch, so
  // it hangs on the relation's declared set (the same on all ranks of a plan), not on the rows at hand.
  if (local->nullable_mask) {
    st = ensure_null_mask(rel, d);
    if (st) { qsgpu_relation_destroy(rel); return st; }
    rel->nullable_mask = local->nullable_mask;
  }
  if (R == 1) {
    for (size_t a = 0; a < local->attrs.size(); ++a)
      if (mine) QS_CUDA(cudaMemcpyAsync(rel->cols[a], local->cols[a], mine * local->attrs[a].width, cudaMemcpyDeviceToDevice, d->stream));
    if (mine && local->nullable_mask) QS_CUDA(cudaMemcpyAsync(rel->d_nulls, local->d_nulls, mine * 8, cudaMemcpyDeviceToDevice, d->stream));
  } else {
    // all-gather with per-rank counts: one broadcast per (rank, attribute), all in ONE NCCL group
    QS_NCCL(g_nccl.GroupStart());
    uint64_t first = 0;
    for (int r = 0; r < R; ++r) {
      if (counts[r]) {
        for (size_t a = 0; a < local->attrs.size(); ++a) {
          const size_t w = local->attrs[a].width;
          QS_NCCL(g_nccl.Broadcast(local->cols[a], rel->cols[a] + first * w, counts[r] * w, ncclChar, r, c->comm, d->stream));
        }
        if (local->nullable_mask)
          QS_NCCL(g_nccl.Broadcast(local->d_nulls, rel->d_nulls + first, counts[r] * 8, ncclChar, r, c->comm, d->stream));
      }
      first += counts[r];
    }
    QS_NCCL(g_nccl.GroupEnd());
  }
  st = qsgpu_relation_set_num_rows(rel, total);
  if (st) { qsgpu_relation_destroy(rel); return st; }
  *out = rel;
  return QSGPU_OK;
}

int qsgpu_relation_allgather_small(qsgpu_relation_t local, qsgpu_comm_t c, uint64_t max_rows_per_rank, qsgpu_relation_t *out) {
  if (!local || !out) { set_error(QSGPU_ERR_INVALID, "null relation"); return QSGPU_ERR_INVALID; }
  Device *d = device(local->dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  if (local->has_codes()) { set_error(QSGPU_ERR_UNSUPPORTED, "all-gather of a dictionary-coded relation"); return QSGPU_ERR_UNSUPPORTED; }
  // The choice between the two forms must come out the same on every rank: it depends only on the schema, on
  // max_rows_per_rank (the caller passes the same value everywhere, e.g. a LIMIT) and on the communicator.
  if (c && c->n_ranks > 1 && c->mailbox && local->attrs.size() <= static_cast<size_t>(kMaxCols)) {
    // every rank's share fits a mailbox slot: one kernel over peer memory, row counts never leave the device
    PeerGatherDesc G{};
    size_t off = 0;
    const uint64_t cap = max_rows_per_rank;
    G.n_cols = static_cast<uint32_t>(local->attrs.size()) + (local->nullable_mask ? 1u : 0u);
    for (uint32_t a = 0; a < G.n_cols; ++a) {
      const bool mask_col = a == local->attrs.size();
      G.width[a] = mask_col ? 8u : local->attrs[a].width;
      G.col_off[a] = static_cast<uint32_t>(off);
      off += (cap * G.width[a] + 7) & ~static_cast<size_t>(7);
    }
    if (8 + off <= kMailSlotBytes) {
      std::lock_guard<std::mutex> lk(c->mu);
      qsgpu_relation *rel = nullptr;
      int st = qsgpu_relation_create(local->dev, static_cast<uint32_t>(local->attrs.size()), local->attrs.data(), std::max<uint64_t>(cap * c->n_ranks, 1), &rel);
      if (st) return st;
      if (local->nullable_mask) {
        st = ensure_null_mask(rel, d);
        if (st) { qsgpu_relation_destroy(rel); return st; }
        rel->nullable_mask = local->nullable_mask;
      }
      for (uint32_t a = 0; a < G.n_cols; ++a) {
        const bool mask_col = a == local->attrs.size();
        G.in[a] = mask_col ? reinterpret_cast<const char *>(local->d_nulls) : local->cols[a];
        G.out[a] = mask_col ? reinterpret_cast<char *>(rel->d_nulls) : rel->cols[a];
      }
      G.cap = static_cast<uint32_t>(cap);
      G.in_rows = local->d_rows;
      G.out_rows = rel->d_rows;
      G.error_flag = d->d_error;
      ++c->epoch;
      k_allgather_peer<<<1, 256, 0, d->stream>>>(G, c->d_peer_mailbox, static_cast<uint32_t>(c->rank), static_cast<uint32_t>(c->n_ranks), c->epoch);
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) { qsgpu_relation_destroy(rel); return cuda_fail(e, "k_allgather_peer"); }
      count_launch();
      rel->dirty = true;                 // the row count stays on the device until somebody asks
      *out = rel;
      return QSGPU_OK;
    }
  }
  return qsgpu_relation_allgather(local, c, out);
}

}  // extern "C"
