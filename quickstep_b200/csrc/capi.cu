// C-ABI of libqsgpu.so (include/qsgpu.h): runtime, device relations, staging
// and the operator entry points.  Host code only; kernels live in k_*.cu.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>

#include "qs_host.h"
#include "qs_jit.h"
#include "qs_lower.h"

namespace qs {

static std::vector<Device> g_devices;
static std::mutex g_mutex;
static bool g_inited = false;
static std::atomic<uint64_t> g_launches{0};
static std::atomic<bool> g_timing{false};
static thread_local std::string t_error;

// qsgpu_jit_selfcheck: runs the real entry points on fake (host-only) handles
// up to the point where the query kernel would be fetched, compiles the
// generated source with NVRTC and stops.  No device is touched.
struct SelfCheck {
  Device dev;
  std::string source, log;
  bool reached = false, compiled = false;
  uint32_t which = 0;
};
static thread_local SelfCheck *t_sc = nullptr;
static constexpr int kSelfCheckStop = 1000;
// Per-family kernel times since qsgpu_set_timing(1).  Process-wide, not thread-local: the kernels of a query that
// runs through an operator layer are launched by Worker threads, the reader is the thread that submitted the query.
struct FamilyMs { float last = 0, max = 0, sum = 0; uint32_t count = 0; };
static FamilyMs g_ms[QS_K_FAMILIES];
static std::mutex g_ms_mu;

void set_error(int status, const std::string &msg) {
  char buf[32];
  std::snprintf(buf, sizeof(buf), "[qsgpu %d] ", status);
  t_error = std::string(buf) + msg;
}

int cuda_fail(cudaError_t e, const char *what) {
  set_error(QSGPU_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  cudaGetLastError();
  return e == cudaErrorMemoryAllocation ? QSGPU_ERR_OOM : QSGPU_ERR_CUDA;
}

void count_launch(int n) { g_launches.fetch_add(static_cast<uint64_t>(n)); }
bool timing_enabled() { return g_timing.load(); }
void record_ms(uint32_t family, float ms) {
  if (family >= QS_K_FAMILIES) return;
  std::lock_guard<std::mutex> lk(g_ms_mu);
  FamilyMs &f = g_ms[family];
  f.last = ms; f.max = std::max(f.max, ms); f.sum += ms; ++f.count;
}

// Device the calling thread last resolved through device(): where dev_malloc / dev_free go.
static thread_local Device *t_dev = nullptr;

cudaError_t dev_malloc_bytes(void **p, size_t bytes) {
  Device *d = t_dev;
  if (!d || !d->pool) return cudaErrorInvalidDevice;
  if (bytes < BlockCache::kMinBytes || !d->cache)
    return cudaMallocFromPoolAsync(p, bytes ? bytes : 256, d->pool, d->stream);
  const size_t rounded = (bytes + BlockCache::kRound - 1) / BlockCache::kRound * BlockCache::kRound;
  BlockCache &C = *d->cache;
  {
    std::lock_guard<std::mutex> lk(C.mu);
    auto it = C.parked.find(rounded);
    if (it != C.parked.end()) {
      *p = it->second;
      C.parked.erase(it);
      C.parked_bytes -= rounded;
      C.live[*p] = rounded;
      return cudaSuccess;
    }
  }
  cudaError_t e = cudaMallocFromPoolAsync(p, rounded, d->pool, d->stream);
  if (e == cudaErrorMemoryAllocation) {          // give the parked blocks back and retry once
    cudaGetLastError();
    std::lock_guard<std::mutex> lk(C.mu);
    for (auto &kv : C.parked) cudaFreeAsync(kv.second, d->stream);
    C.parked.clear();
    C.parked_bytes = 0;
    e = cudaMallocFromPoolAsync(p, rounded, d->pool, d->stream);
  }
  if (e == cudaSuccess) {
    std::lock_guard<std::mutex> lk(C.mu);
    C.live[*p] = rounded;
  }
  return e;
}

cudaError_t dev_free(void *p) {
  if (!p) return cudaSuccess;
  Device *d = t_dev;
  if (!d || !d->pool) return cudaFree(p);     // after shutdown: synchronous free is still legal
  if (d->cache) {
    BlockCache &C = *d->cache;
    std::lock_guard<std::mutex> lk(C.mu);
    auto it = C.live.find(p);
    if (it != C.live.end()) {
      const size_t sz = it->second;
      C.live.erase(it);
      if (C.parked_bytes + sz <= BlockCache::kMaxParkedBytes) {
        C.parked.emplace(sz, p);
        C.parked_bytes += sz;
        return cudaSuccess;
      }
    }
  }
  return cudaFreeAsync(p, d->stream);
}

Device *device(int dev) {
  if (t_sc) return &t_sc->dev;
  if (!g_inited) {
    set_error(QSGPU_ERR_NO_DEVICE, "qsgpu_init has not been called (or found no CUDA device); there is no CPU fallback");
    return nullptr;
  }
  for (auto &d : g_devices) if (d.id == dev) {
    if (cudaSetDevice(dev) != cudaSuccess) { set_error(QSGPU_ERR_CUDA, "cudaSetDevice failed"); return nullptr; }
    t_dev = &d;
    return &d;
  }
  set_error(QSGPU_ERR_INVALID, "device was not passed to qsgpu_init");
  return nullptr;
}

// Times one kernel family with CUDA events on the launching stream.
static std::mutex g_timer_mu;      // the two events are per device: one timed launch at a time
struct KernelTimer {
  Device *d; uint32_t family; bool on;
  std::unique_lock<std::mutex> lk;
  KernelTimer(Device *dev, uint32_t fam) : d(dev), family(fam), on(timing_enabled()) {
    if (on) { lk = std::unique_lock<std::mutex>(g_timer_mu); cudaEventRecord(d->ev0, d->stream); }
  }
  ~KernelTimer() {
    if (!on) return;
    cudaEventRecord(d->ev1, d->stream);
    cudaEventSynchronize(d->ev1);
    float ms = 0;
    cudaEventElapsedTime(&ms, d->ev0, d->ev1);
    record_ms(family, ms);
  }
};

// Picks ring depth, shared memory size and grid for a scan.
int plan_scan(Device *d, ScanDesc *S, size_t extra_smem, ScanPlan *plan, int max_ctas) {
  uint32_t stage = 0;
  for (uint32_t c = 0; c < S->n_cols; ++c) {
    ColDesc &C = S->cols[c];
    if (C.cw != 0) {                      // coded attribute: the code tile is what travels
      C.code_off = stage;
      stage += static_cast<uint32_t>(kTileRows) * C.cw;
      if (!C.expand) { C.smem_off = 0; continue; }
    }
    C.smem_off = stage;
    stage += static_cast<uint32_t>(kTileRows) * C.width;   // multiple of 16 by construction
  }
  stage = (stage + 127u) & ~127u;
  if (stage == 0) stage = 128;
  S->stage_bytes = stage;
  // shared-memory copies of the dictionaries of 1-byte-coded attributes whose values the kernel looks up
  size_t dict_bytes = 0;
  for (uint32_t c = 0; c < S->n_cols; ++c)
    if (S->cols[c].dict_smem) dict_bytes += 256u * S->cols[c].width;
  if (dict_bytes > 16384) {   // CHAR(n) dictionaries can be wide: leave those in global memory (L1-cached)
    for (uint32_t c = 0; c < S->n_cols; ++c) S->cols[c].dict_smem = 0;
    dict_bytes = 0;
  }
  // shared-memory copies of the small LIP filters the scan probes (north star: "LIP filters in shared memory"): a
  // filter of <= 32 KB (48 KB for all of one scan) is copied into every CTA before its first tile; larger ones are
  // probed where they are (L2 holds them; a 188 KB filter per CTA would leave one resident CTA per SM)
  size_t lip_bytes = 0;
  {
    static const bool lip_smem_on = [] { const char *e = std::getenv("QSGPU_LIP_SMEM"); return !(e && e[0] == '0'); }();
    for (uint32_t f = 0; f < S->n_lip; ++f) {
      S->lip[f].smem_off = 0;
      const size_t b = (S->lip[f].n_words * 8 + 15) & ~static_cast<size_t>(15);
      if (!lip_smem_on || S->lip[f].n_words == 0 || b > kLipSmemBytes || lip_bytes + b > (48u << 10)) continue;
      S->lip[f].smem_off = 1;             // marked; the offset is assigned below, once the ring is sized
      lip_bytes += b;
    }
  }
  const size_t fixed = kBarBytes + extra_smem + 256 + dict_bytes + lip_bytes;
  const size_t per_sm = d->smem_per_sm;                     // 228 KB on B200
  const size_t per_block_max = d->smem_per_block_optin;     // 227 KB
  // Resident CTAs per SM the shared-memory ring allows: narrow scans (join keys, LIP builds: a few KB per
  // tile) run 3-4 CTAs per SM so that more random table accesses are in flight; wide ones 2, very wide 1.
  // Registers may allow fewer: the grid is finally sized from the occupancy the driver reports for the
  // compiled kernel (launch_query_kernel).
  int ctas = 1;
  size_t budget = per_block_max;
  for (int c = std::min(8, std::max(1, max_ctas)); c >= 2; --c) {
    const size_t b = per_sm / c - 1024;                     // 1 KB per CTA is reserved by the driver
    if (fixed + 2ull * stage <= b) { ctas = c; budget = b; break; }     // double buffering is the minimum
  }
  if (fixed + 2ull * stage > budget) {
    set_error(QSGPU_ERR_UNSUPPORTED, "scan references too many bytes per row for the shared-memory tile ring");
    return QSGPU_ERR_UNSUPPORTED;
  }
  uint32_t n_stages = static_cast<uint32_t>((budget - fixed) / stage);
  if (n_stages > static_cast<uint32_t>(kMaxStages)) n_stages = kMaxStages;
  S->n_stages = n_stages;
  plan->smem = kBarBytes + static_cast<size_t>(n_stages) * stage + extra_smem + 64;
  plan->smem = (plan->smem + 15) & ~static_cast<size_t>(15);
  for (uint32_t c = 0; c < S->n_cols; ++c) {
    if (!S->cols[c].dict_smem) continue;
    S->cols[c].dict_soff = static_cast<uint32_t>(plan->smem);
    plan->smem += 256u * S->cols[c].width;
  }
  for (uint32_t f = 0; f < S->n_lip; ++f) {
    if (!S->lip[f].smem_off) continue;
    plan->smem = (plan->smem + 15) & ~static_cast<size_t>(15);
    S->lip[f].smem_off = static_cast<uint32_t>(plan->smem);
    plan->smem += (S->lip[f].n_words * 8 + 15) & ~static_cast<size_t>(15);
  }
  plan->ctas = ctas;
  int grid = d->sm_count * ctas;
  if (S->d_row_end == nullptr && S->n_tiles < static_cast<uint32_t>(grid)) grid = std::max<int>(1, S->n_tiles);
  plan->grid = grid;
  return QSGPU_OK;
}

static int check_device_error(Device *d) {
  uint32_t flag = 0;
  QS_CUDA(cudaMemcpyAsync(&flag, d->d_error, 4, cudaMemcpyDeviceToHost, d->stream));
  QS_CUDA(cudaStreamSynchronize(d->stream));
  if (flag != 0) {
    uint32_t zero = 0;
    cudaMemcpyAsync(d->d_error, &zero, 4, cudaMemcpyHostToDevice, d->stream);
    set_error(static_cast<int>(flag), "a kernel reported a capacity overflow (hash table, group limit or output relation too small)");
    return static_cast<int>(flag);
  }
  return QSGPU_OK;
}

// Fetches (compiling on first use) the kernel specialised for this work order.
static int query_kernel(JitFamily fam, const ScanDesc &S, const Program &P, const ScanPlan &plan, const AggDesc *A,
                        const SinkDesc *K, const JoinDesc *J, int hot, JitKernel **out, int priv = 0) {
  JitSpec sp;
  sp.family = fam; sp.S = &S; sp.P = &P; sp.A = A; sp.K = K; sp.J = J; sp.hot = hot; sp.priv = priv; sp.ctas_per_sm = plan.ctas;
  if (t_sc) {
    std::string cubin;
    t_sc->reached = true;
    t_sc->source = jit_source(sp);
    t_sc->compiled = jit_compile_only(t_sc->source, &cubin, &t_sc->log) == QSGPU_OK;
    if (const char *dir = std::getenv("QSGPU_JIT_DUMP")) {   // for cuobjdump / ptxas inspection
      const std::string stem = std::string(dir) + "/selfcheck_" + std::to_string(t_sc->which);
      if (FILE *f = std::fopen((stem + ".cu").c_str(), "wb")) { std::fwrite(t_sc->source.data(), 1, t_sc->source.size(), f); std::fclose(f); }
      if (t_sc->compiled) if (FILE *f = std::fopen((stem + ".cubin").c_str(), "wb")) { std::fwrite(cubin.data(), 1, cubin.size(), f); std::fclose(f); }
    }
    return kSelfCheckStop;
  }
  return jit_get(sp, out);
}

static cudaError_t launch_query_kernel(Device *d, JitKernel *k, const ScanDesc &S, const Program &P,
                                       const ScanPlan &plan, const AggDesc *A, const SinkDesc *K,
                                       const JoinDesc *J) {
  void *args[5];
  int n = 0;
  args[n++] = const_cast<ScanDesc *>(&S);
  args[n++] = const_cast<Lits *>(&P.L);
  if (A) args[n++] = const_cast<AggDesc *>(A);
  if (K) args[n++] = const_cast<SinkDesc *>(K);
  if (J) args[n++] = const_cast<JoinDesc *>(J);
  // persistent grid: no more CTAs than can be resident at once (registers may allow fewer than the ring does)
  int grid = plan.grid;
  const int occ = jit_occupancy(k, plan.smem);
  if (occ > 0) grid = std::min(grid, d->sm_count * std::min(occ, plan.ctas));
  return jit_launch(k, grid, plan.smem, d->stream, args);
}

static int sync_rows(qsgpu_relation *rel) {
  if (!rel->dirty) return QSGPU_OK;
  Device *d = device(rel->dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  unsigned long long n = 0;
  QS_CUDA(cudaMemcpyAsync(&n, rel->d_rows, 8, cudaMemcpyDeviceToHost, d->stream));
  QS_CUDA(cudaStreamSynchronize(d->stream));
  int st = check_device_error(d);
  if (st) return st;
  rel->host_rows = std::min<uint64_t>(n, rel->capacity);
  rel->dirty = false;
  return QSGPU_OK;
}

// Fills the row extent + staged columns of a ScanDesc from a lowering.
static int fill_scan(const qsgpu_relation *rel, uint64_t row_begin, uint64_t row_end, const Lowering &L,
                     ScanDesc *S) {
  std::memset(S, 0, sizeof(*S));
  uint64_t limit = rel->dirty ? rel->capacity : rel->host_rows;
  if (row_end > limit) row_end = limit;
  if (row_begin > row_end) row_begin = row_end;
  S->row_begin = row_begin;
  S->row_end = row_end;
  S->first_row = row_begin & ~15ull;
  S->d_row_end = rel->dirty ? rel->d_rows : nullptr;
  const uint64_t span = row_end - S->first_row;
  S->n_tiles = static_cast<uint32_t>((span + kTileRows - 1) / kTileRows);
  S->n_cols = static_cast<uint32_t>(L.staged_attrs.size());
  for (uint32_t c = 0; c < S->n_cols; ++c) {
    const uint32_t a = L.staged_attrs[c];
    if (a == Lowering::kNullMaskAttr) {       // the per-row NULL mask, staged like a LONG column
      S->cols[c].ptr = reinterpret_cast<const char *>(rel->d_nulls);
      S->cols[c].width = 8;
      continue;
    }
    S->cols[c].ptr = rel->cols[a];
    S->cols[c].width = rel->attrs[a].width;
    if (const uint32_t cw = rel->code_width(a)) {
      S->cols[c].cw = static_cast<uint8_t>(cw);
      S->cols[c].dict = rel->coded[a].d_dict;
      S->cols[c].dict_entries = rel->coded[a].n_entries;
      S->cols[c].expand = (L.staged_use[c] & Lowering::USE_RAW) ? 1 : 0;
      S->cols[c].dict_smem = (cw == 1 && rel->attrs[a].width <= 16 &&
                              (L.staged_use[c] & (Lowering::USE_RAW | Lowering::USE_VALUE))) ? 1 : 0;
    }
  }
  return QSGPU_OK;
}

static int fill_lips(uint32_t n, const qs_lip_ref *refs, const qsgpu_relation *rel, int dev, ScanDesc *S) {
  if (n > static_cast<uint32_t>(kMaxLip)) { set_error(QSGPU_ERR_UNSUPPORTED, "more than kMaxLip LIP filters on one scan"); return QSGPU_ERR_UNSUPPORTED; }
  S->n_lip = n;
  for (uint32_t i = 0; i < n; ++i) {
    if (!refs[i].lip || refs[i].lip->dev != dev) { set_error(QSGPU_ERR_INVALID, "LIP filter missing or on another device"); return QSGPU_ERR_INVALID; }
    S->lip[i] = refs[i].lip->d;
  }
  (void)rel;
  return QSGPU_OK;
}

// Native values of rows [row_begin, row_begin + n_rows) of a coded attribute in a fresh device buffer.
static int decode_coded(Device *d, const qsgpu_relation *rel, uint32_t attr, uint64_t row_begin, uint64_t n_rows,
                        char **out) {
  const qs_coded_attr &C = rel->coded[attr];
  const uint32_t w = rel->attrs[attr].width;
  QS_CUDA(dev_malloc(out, n_rows * w + 16));
  QS_CUDA(launch_decode_dict(*out, rel->cols[attr] + row_begin * C.cw, C.d_dict, n_rows, C.cw, w, C.n_entries, d->stream));
  if (n_rows) count_launch();
  return QSGPU_OK;
}

static uint16_t attr_width(uint16_t type, uint16_t width) {
  switch (type) {
    case QS_INT: case QS_FLOAT: return 4;
    case QS_LONG: case QS_DOUBLE: case QS_DATE: return 8;
    default: return width;
  }
}

static size_t padded_bytes(uint64_t rows, uint32_t width) {
  return ((rows * width + 255) & ~static_cast<uint64_t>(255)) + 256;   // 16+ readable bytes past the end
}

// The per-row NULL mask of a relation (bit a = attribute a is NULL), allocated on first need, all zeros.  Scans
// stage it like a LONG column, so it is padded like one.
int ensure_null_mask(qsgpu_relation *rel, Device *d) {
  if (rel->d_nulls) return QSGPU_OK;
  const size_t bytes = padded_bytes(std::max<uint64_t>(rel->capacity, 1), 8);
  QS_CUDA(dev_malloc(&rel->d_nulls, bytes));
  QS_CUDA(cudaMemsetAsync(rel->d_nulls, 0, bytes, d->stream));
  return QSGPU_OK;
}

}  // namespace qs

using namespace qs;

extern "C" {

const char *qsgpu_last_error(void) { return t_error.c_str(); }

int qsgpu_init(int n_dev, const int *dev_ids) {
  std::lock_guard<std::mutex> lk(g_mutex);
  if (g_inited) return QSGPU_OK;
  int visible = 0;
  cudaError_t e = cudaGetDeviceCount(&visible);
  if (e != cudaSuccess || visible == 0) {
    cudaGetLastError();
    set_error(QSGPU_ERR_NO_DEVICE, "no CUDA device visible; libqsgpu has no CPU fallback");
    return QSGPU_ERR_NO_DEVICE;
  }
  std::vector<int> ids;
  if (n_dev <= 0) for (int i = 0; i < visible; ++i) ids.push_back(i);
  else for (int i = 0; i < n_dev; ++i) ids.push_back(dev_ids ? dev_ids[i] : i);
  for (int id : ids) {
    if (id < 0 || id >= visible) { set_error(QSGPU_ERR_INVALID, "device id out of range"); return QSGPU_ERR_INVALID; }
    Device d;
    d.id = id;
    QS_CUDA(cudaSetDevice(id));
    cudaDeviceProp prop;
    QS_CUDA(cudaGetDeviceProperties(&prop, id));
    if (prop.major < 10) {
      set_error(QSGPU_ERR_NO_DEVICE, "libqsgpu is built for sm_100a (B200) only");
      return QSGPU_ERR_NO_DEVICE;
    }
    d.sm_count = prop.multiProcessorCount;
    d.smem_per_sm = prop.sharedMemPerMultiprocessor;
    d.smem_per_block_optin = prop.sharedMemPerBlockOptin;
    QS_CUDA(cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
    QS_CUDA(cudaStreamCreateWithFlags(&d.copy_stream, cudaStreamNonBlocking));
    {
      // Per-device stream-ordered arena: states, join tables and temporary relations are
      // allocated and freed once per query, so cudaMalloc/cudaFree (a device-wide
      // synchronisation each, ~0.7 ms apiece next to a large context; r01b call profile) must
      // stay off the query path.  Freed blocks are kept (release threshold = max) and reused.
      cudaMemPoolProps props{};
      props.allocType = cudaMemAllocationTypePinned;
      props.handleTypes = cudaMemHandleTypeNone;
      props.location.type = cudaMemLocationTypeDevice;
      props.location.id = id;
      QS_CUDA(cudaMemPoolCreate(&d.pool, &props));
      uint64_t keep = UINT64_MAX;
      QS_CUDA(cudaMemPoolSetAttribute(d.pool, cudaMemPoolAttrReleaseThreshold, &keep));
      d.cache = std::make_shared<BlockCache>();
      d.read_mu = std::make_shared<std::mutex>();
      QS_CUDA(cudaMalloc(&d.read_scratch, kReadScratchBytes));
      QS_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&d.read_pinned), kReadScratchBytes, cudaHostAllocPortable));
    }
    QS_CUDA(cudaMalloc(&d.d_error, 256));
    QS_CUDA(cudaMemset(d.d_error, 0, 256));
    QS_CUDA(cudaEventCreate(&d.ev0));
    QS_CUDA(cudaEventCreate(&d.ev1));
    QS_CUDA(cudaEventCreate(&d.ev_t0));
    QS_CUDA(cudaEventCreate(&d.ev_t1));
    g_devices.push_back(d);
  }
  g_inited = true;
  return QSGPU_OK;
}

int qsgpu_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_mutex);
  for (auto &d : g_devices) {
    cudaSetDevice(d.id);
    cudaStreamSynchronize(d.stream);
    cudaFree(d.d_error);
    cudaFree(d.read_scratch);
    cudaFreeHost(d.read_pinned);
    cudaEventDestroy(d.ev0);
    cudaEventDestroy(d.ev1);
    cudaEventDestroy(d.ev_t0);
    cudaEventDestroy(d.ev_t1);
    if (d.cache) for (auto &kv : d.cache->parked) cudaFreeAsync(kv.second, d.stream);
    cudaStreamSynchronize(d.stream);
    if (d.pool) cudaMemPoolDestroy(d.pool);
    cudaStreamDestroy(d.copy_stream);
    cudaStreamDestroy(d.stream);
  }
  t_dev = nullptr;
  g_devices.clear();
  g_inited = false;
  return QSGPU_OK;
}

int qsgpu_device_count(int *n_dev) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); n = 0; }
  *n_dev = n;
  return QSGPU_OK;
}

int qsgpu_synchronize(int dev) {
  Device *d = device(dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  QS_CUDA(cudaStreamSynchronize(d->stream));
  return check_device_error(d);
}

int qsgpu_stream(int dev, void **stream) {
  Device *d = device(dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  *stream = d->stream;
  return QSGPU_OK;
}

int qsgpu_launch_count(uint64_t *n) { *n = g_launches.load(); return QSGPU_OK; }

int qsgpu_set_timing(int enabled) {
  if (enabled) { std::lock_guard<std::mutex> lk(g_ms_mu); for (FamilyMs &f : g_ms) f = FamilyMs{}; }
  g_timing.store(enabled != 0);
  return QSGPU_OK;
}
int qsgpu_kernel_ms_stats(uint32_t family, float *last, float *max, float *sum, uint32_t *count) {
  if (family >= QS_K_FAMILIES) { set_error(QSGPU_ERR_INVALID, "unknown kernel family"); return QSGPU_ERR_INVALID; }
  std::lock_guard<std::mutex> lk(g_ms_mu);
  if (last) *last = g_ms[family].last;
  if (max) *max = g_ms[family].max;
  if (sum) *sum = g_ms[family].sum;
  if (count) *count = g_ms[family].count;
  return QSGPU_OK;
}
int qsgpu_last_kernel_ms(uint32_t family, float *ms) {
  if (family >= QS_K_FAMILIES) { set_error(QSGPU_ERR_INVALID, "bad kernel family"); return QSGPU_ERR_INVALID; }
  { std::lock_guard<std::mutex> lk(g_ms_mu); *ms = g_ms[family].last; }
  return QSGPU_OK;
}

int qsgpu_malloc(int dev, size_t bytes, void **dptr) {
  Device *d = device(dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  QS_CUDA(dev_malloc(dptr, bytes ? bytes : 256));
  QS_CUDA(cudaStreamSynchronize(d->stream));   // the caller may use the block on any stream
  return QSGPU_OK;
}
int qsgpu_free(int dev, void *dptr) {
  Device *d = device(dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  QS_CUDA(cudaStreamSynchronize(d->stream));
  QS_CUDA(dev_free(dptr));
  return QSGPU_OK;
}
int qsgpu_memcpy_h2d(int dev, void *dst, const void *src, size_t bytes) {
  Device *d = device(dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  QS_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, d->stream));
  QS_CUDA(cudaStreamSynchronize(d->stream));
  return QSGPU_OK;
}
int qsgpu_memcpy_d2h(int dev, void *dst, const void *src, size_t bytes) {
  Device *d = device(dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  QS_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, d->stream));
  QS_CUDA(cudaStreamSynchronize(d->stream));
  return QSGPU_OK;
}
int qsgpu_memcpy_d2d(int dev, void *dst, const void *src, size_t bytes) {
  Device *d = device(dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  QS_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, d->stream));
  QS_CUDA(cudaStreamSynchronize(d->stream));
  return QSGPU_OK;
}
int qsgpu_memcpy_d2d_async(int dev, void *dst, const void *src, size_t bytes) {
  Device *d = device(dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  QS_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, d->stream));
  return QSGPU_OK;
}
int qsgpu_timer_start(int dev) {
  Device *d = device(dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  QS_CUDA(cudaEventRecord(d->ev_t0, d->stream));
  return QSGPU_OK;
}
int qsgpu_timer_stop(int dev, float *ms) {
  Device *d = device(dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  QS_CUDA(cudaEventRecord(d->ev_t1, d->stream));
  QS_CUDA(cudaEventSynchronize(d->ev_t1));
  QS_CUDA(cudaEventElapsedTime(ms, d->ev_t0, d->ev_t1));
  return QSGPU_OK;
}
int qsgpu_host_alloc(size_t bytes, void **hptr) {
  if (!g_inited) { set_error(QSGPU_ERR_NO_DEVICE, "qsgpu_init has not been called"); return QSGPU_ERR_NO_DEVICE; }
  QS_CUDA(cudaHostAlloc(hptr, bytes ? bytes : 16, cudaHostAllocPortable));
  return QSGPU_OK;
}
int qsgpu_host_free(void *hptr) {
  QS_CUDA(cudaFreeHost(hptr));
  return QSGPU_OK;
}

/* ------------------------------------------------------------- relations */
static int relation_new(int dev, uint32_t n_attrs, const qs_attr *attrs, qsgpu_relation **out) {
  Device *d = device(dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  if (n_attrs == 0 || !attrs) { set_error(QSGPU_ERR_INVALID, "relation needs at least one attribute"); return QSGPU_ERR_INVALID; }
  std::unique_ptr<qsgpu_relation> r(new qsgpu_relation);
  r->dev = dev;
  for (uint32_t i = 0; i < n_attrs; ++i) {
    qs_attr a = attrs[i];
    if (a.type == QS_VARCHAR || a.type > QS_DATE) { set_error(QSGPU_ERR_UNSUPPORTED, "VARCHAR / non fixed-width attributes are not staged on the device"); return QSGPU_ERR_UNSUPPORTED; }
    a.width = attr_width(a.type, a.width);
    if (a.width == 0) { set_error(QSGPU_ERR_INVALID, "zero-width attribute"); return QSGPU_ERR_INVALID; }
    r->attrs.push_back(a);
  }
  QS_CUDA(dev_malloc(&r->d_rows, 256));
  QS_CUDA(cudaMemsetAsync(r->d_rows, 0, 256, d->stream));
  *out = r.release();
  return QSGPU_OK;
}

int qsgpu_relation_create(int dev, uint32_t n_attrs, const qs_attr *attrs, uint64_t capacity_rows,
                          qsgpu_relation_t *out) {
  qsgpu_relation *r = nullptr;
  int st = relation_new(dev, n_attrs, attrs, &r);
  if (st) return st;
  r->capacity = capacity_rows;
  r->owns_memory = true;
  for (auto &a : r->attrs) {
    char *p = nullptr;
    cudaError_t e = dev_malloc(&p, padded_bytes(capacity_rows, a.width));
    if (e != cudaSuccess) { qsgpu_relation_destroy(r); return cuda_fail(e, "dev_malloc(relation column)"); }
    r->cols.push_back(p);
  }
  *out = r;
  return QSGPU_OK;
}

int qsgpu_relation_wrap(int dev, uint32_t n_attrs, const qs_attr *attrs, void *const *dptrs, uint64_t n_rows,
                        qsgpu_relation_t *out) {
  qsgpu_relation *r = nullptr;
  int st = relation_new(dev, n_attrs, attrs, &r);
  if (st) return st;
  r->capacity = n_rows;
  r->host_rows = n_rows;
  r->owns_memory = false;
  for (uint32_t i = 0; i < n_attrs; ++i) {
    if ((reinterpret_cast<uintptr_t>(dptrs[i]) & 15) != 0) {
      qsgpu_relation_destroy(r);
      set_error(QSGPU_ERR_INVALID, "wrapped column is not 16-byte aligned");
      return QSGPU_ERR_INVALID;
    }
    r->cols.push_back(static_cast<char *>(dptrs[i]));
  }
  Device *d = device(dev);
  unsigned long long n = n_rows;
  QS_CUDA(cudaMemcpyAsync(r->d_rows, &n, 8, cudaMemcpyHostToDevice, d->stream));
  QS_CUDA(cudaStreamSynchronize(d->stream));
  *out = r;
  return QSGPU_OK;
}

int qsgpu_relation_destroy(qsgpu_relation_t rel) {
  if (!rel) return QSGPU_OK;
  Device *d = device(rel->dev);
  // owned columns are freed in stream order; wrapped ones belong to the caller, who may release
  // them as soon as this returns
  if (d && !rel->owns_memory) cudaStreamSynchronize(d->stream);
  if (rel->owns_memory) for (char *p : rel->cols) dev_free(p);
  for (auto &c : rel->coded) dev_free(c.d_dict);
  dev_free(rel->d_nulls);
  dev_free(rel->d_rows);
  delete rel;
  return QSGPU_OK;
}

int qsgpu_relation_num_rows(qsgpu_relation_t rel, uint64_t *n_rows) {
  int st = sync_rows(rel);
  if (st) return st;
  *n_rows = rel->host_rows;
  return QSGPU_OK;
}

int qsgpu_relation_set_num_rows(qsgpu_relation_t rel, uint64_t n_rows) {
  Device *d = device(rel->dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  if (n_rows > rel->capacity) { set_error(QSGPU_ERR_CAPACITY, "row count above relation capacity"); return QSGPU_ERR_CAPACITY; }
  unsigned long long n = n_rows;
  QS_CUDA(cudaMemcpyAsync(rel->d_rows, &n, 8, cudaMemcpyHostToDevice, d->stream));
  QS_CUDA(cudaStreamSynchronize(d->stream));
  rel->host_rows = n_rows;
  rel->dirty = false;
  return QSGPU_OK;
}

int qsgpu_relation_column(qsgpu_relation_t rel, uint32_t attr, void **dptr) {
  if (attr >= rel->cols.size()) { set_error(QSGPU_ERR_INVALID, "attribute id out of range"); return QSGPU_ERR_INVALID; }
  *dptr = rel->cols[attr];
  return QSGPU_OK;
}

int qsgpu_relation_read(qsgpu_relation_t rel, uint32_t attr, uint64_t row_begin, uint64_t n_rows,
                        void *host_out) {
  int st = sync_rows(rel);
  if (st) return st;
  if (attr >= rel->cols.size() || row_begin + n_rows > rel->host_rows) { set_error(QSGPU_ERR_INVALID, "read outside relation"); return QSGPU_ERR_INVALID; }
  Device *d = device(rel->dev);
  const uint32_t w = rel->attrs[attr].width;
  char *tmp = nullptr;
  const char *src = rel->cols[attr] + row_begin * w;
  if (rel->code_width(attr) != 0) {         // coded attribute: callers always see native values
    int rc = decode_coded(d, rel, attr, row_begin, n_rows, &tmp);
    if (rc) return rc;
    src = tmp;
  }
  cudaError_t e = cudaSuccess;
  if (n_rows) e = cudaMemcpyAsync(host_out, src, n_rows * w, cudaMemcpyDeviceToHost, d->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(d->stream);
  dev_free(tmp);
  if (e != cudaSuccess) return cuda_fail(e, "qsgpu_relation_read");
  return QSGPU_OK;
}

int qsgpu_relation_set_dictionary(qsgpu_relation_t rel, uint32_t attr, uint32_t code_width, const void *dict_values,
                                  uint32_t n_entries) {
  if (!rel) { set_error(QSGPU_ERR_INVALID, "null relation"); return QSGPU_ERR_INVALID; }
  Device *d = device(rel->dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  int st = sync_rows(rel);
  if (st) return st;
  if (attr >= rel->attrs.size()) { set_error(QSGPU_ERR_INVALID, "attribute id out of range"); return QSGPU_ERR_INVALID; }
  if (attr < 64 && ((rel->nullable_mask >> attr) & 1ull)) { set_error(QSGPU_ERR_UNSUPPORTED, "a NULL-able attribute is held at native width, not as codes of a relation-wide dictionary"); return QSGPU_ERR_UNSUPPORTED; }
  if (code_width != 1 && code_width != 2 && code_width != 4) { set_error(QSGPU_ERR_INVALID, "code width must be 1, 2 or 4"); return QSGPU_ERR_INVALID; }
  if (!dict_values || n_entries == 0 || (code_width < 4 && n_entries > (1u << (8 * code_width)))) {
    set_error(QSGPU_ERR_INVALID, "dictionary is empty or has more entries than the code width can address");
    return QSGPU_ERR_INVALID;
  }
  if (rel->code_width(attr) != 0) { set_error(QSGPU_ERR_INVALID, "attribute already has a dictionary"); return QSGPU_ERR_INVALID; }
  if (rel->owns_memory && rel->host_rows != 0) { set_error(QSGPU_ERR_INVALID, "a dictionary must be declared before rows are staged"); return QSGPU_ERR_INVALID; }
  const uint32_t w = rel->attrs[attr].width;
  const char *dv = static_cast<const char *>(dict_values);
  for (uint32_t e = 1; e < n_entries; ++e) {
    if (dict_compare(rel->attrs[attr].type, w, dv + static_cast<size_t>(e - 1) * w, dv + static_cast<size_t>(e) * w) != -1) {
      set_error(QSGPU_ERR_INVALID, "dictionary entries must be strictly increasing in the attribute's order");
      return QSGPU_ERR_INVALID;
    }
  }
  if (rel->coded.empty()) rel->coded.resize(rel->attrs.size());
  qs_coded_attr &C = rel->coded[attr];
  // readable for every value a 1/2-byte code can take: rows past the end of a ragged tile need no clamp
  const size_t slots = code_width < 4 ? (static_cast<size_t>(1) << (8 * code_width)) : n_entries;
  const size_t bytes = slots * w + 16;
  QS_CUDA(dev_malloc(&C.d_dict, bytes));
  QS_CUDA(cudaMemsetAsync(C.d_dict, 0, bytes, d->stream));
  C.h_dict.assign(dv, dv + static_cast<size_t>(n_entries) * w);
  if (rel->attrs[attr].type == QS_DATE)      // DateLit padding is not initialised by the reference: canonical zeros
    for (uint32_t e = 0; e < n_entries; ++e) { C.h_dict[static_cast<size_t>(e) * 8 + 6] = 0; C.h_dict[static_cast<size_t>(e) * 8 + 7] = 0; }
  if (rel->attrs[attr].type == QS_CHAR)      // ... nor the bytes behind a CHAR(n) value's terminator
    for (uint32_t e = 0; e < n_entries; ++e) {
      bool ended = false;
      for (uint32_t b = 0; b < w; ++b) { char &c = C.h_dict[static_cast<size_t>(e) * w + b]; if (ended) c = 0; else ended = c == 0; }
    }
  QS_CUDA(cudaMemcpyAsync(C.d_dict, C.h_dict.data(), C.h_dict.size(), cudaMemcpyHostToDevice, d->stream));
  QS_CUDA(cudaStreamSynchronize(d->stream));
  if (rel->owns_memory) {   // the column now holds codes: give the native-width buffer back
    char *p = nullptr;
    QS_CUDA(dev_malloc(&p, padded_bytes(rel->capacity, code_width)));
    dev_free(rel->cols[attr]);
    rel->cols[attr] = p;
  }
  C.cw = code_width;
  C.n_entries = n_entries;
  return QSGPU_OK;
}

int qsgpu_dictionary_code_range(uint16_t attr_type, uint16_t attr_width, const void *dict_values, uint32_t n_entries,
                                uint32_t cmp, const qs_node *literal, const char *str_pool, uint32_t str_pool_bytes,
                                uint32_t *first, uint32_t *count, int *negate) {
  if (!dict_values || !literal || !first || !count || !negate || literal->kind != QS_N_LITERAL) { set_error(QSGPU_ERR_INVALID, "qsgpu_dictionary_code_range: bad arguments"); return QSGPU_ERR_INVALID; }
  uint64_t lo = 0, span = 0;
  bool neg = false;
  std::string why;
  const int st = dict_code_range(attr_type, qs::attr_width(attr_type, attr_width), static_cast<const char *>(dict_values), n_entries,
                                 static_cast<uint8_t>(cmp), literal, str_pool ? str_pool : "", str_pool_bytes, &lo, &span, &neg, &why);
  if (st != QSGPU_OK) { set_error(st, why); return st; }
  *first = static_cast<uint32_t>(lo); *count = static_cast<uint32_t>(span); *negate = neg ? 1 : 0;
  return QSGPU_OK;
}

int qsgpu_relation_dictionary(qsgpu_relation_t rel, uint32_t attr, uint32_t *code_width, uint32_t *n_entries,
                              void *dict_out) {
  if (!rel || attr >= rel->attrs.size()) { set_error(QSGPU_ERR_INVALID, "attribute id out of range"); return QSGPU_ERR_INVALID; }
  const uint32_t cw = rel->code_width(attr);
  if (code_width) *code_width = cw;
  if (n_entries) *n_entries = cw ? rel->coded[attr].n_entries : 0;
  if (dict_out && cw) std::memcpy(dict_out, rel->coded[attr].h_dict.data(), rel->coded[attr].h_dict.size());
  return QSGPU_OK;
}

int qsgpu_relation_read_nulls(qsgpu_relation_t rel, uint64_t row_begin, uint64_t n_rows, uint64_t *host_out) {
  int st = sync_rows(rel);
  if (st) return st;
  if (row_begin + n_rows > rel->host_rows) { set_error(QSGPU_ERR_INVALID, "read outside relation"); return QSGPU_ERR_INVALID; }
  if (!rel->d_nulls) { std::memset(host_out, 0, n_rows * 8); return QSGPU_OK; }
  return qsgpu_memcpy_d2h(rel->dev, host_out, rel->d_nulls + row_begin, n_rows * 8);
}

int qsgpu_relation_set_nullable(qsgpu_relation_t rel, uint64_t attr_mask) {
  Device *d = device(rel->dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  if (rel->attrs.size() < 64 && (attr_mask >> rel->attrs.size()) != 0) { set_error(QSGPU_ERR_INVALID, "NULL-able attribute out of range"); return QSGPU_ERR_INVALID; }
  for (size_t a = 0; a < rel->attrs.size() && a < 64; ++a)
    if (((attr_mask >> a) & 1ull) && rel->code_width(static_cast<uint32_t>(a))) { set_error(QSGPU_ERR_UNSUPPORTED, "a NULL-able attribute is held at native width, not as codes of a relation-wide dictionary"); return QSGPU_ERR_UNSUPPORTED; }
  if (attr_mask) { const int st = ensure_null_mask(rel, d); if (st) return st; }
  rel->nullable_mask |= attr_mask;
  return QSGPU_OK;
}

int qsgpu_relation_nullable(qsgpu_relation_t rel, uint64_t *attr_mask) { *attr_mask = rel->nullable_mask; return QSGPU_OK; }

int qsgpu_relation_write_nulls(qsgpu_relation_t rel, uint64_t row_begin, uint64_t n_rows, const uint64_t *masks) {
  Device *d = device(rel->dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  if (row_begin + n_rows > rel->capacity) { set_error(QSGPU_ERR_INVALID, "write_nulls out of range"); return QSGPU_ERR_INVALID; }
  if (!rel->d_nulls) { set_error(QSGPU_ERR_INVALID, "write_nulls on a relation without NULL-able attributes (qsgpu_relation_set_nullable first)"); return QSGPU_ERR_INVALID; }
  for (uint64_t i = 0; i < n_rows; ++i)
    if (masks[i] & ~rel->nullable_mask) { set_error(QSGPU_ERR_INVALID, "NULL bit of an attribute that is not NULL-able"); return QSGPU_ERR_INVALID; }
  return qsgpu_memcpy_h2d(rel->dev, rel->d_nulls + row_begin, masks, n_rows * 8);
}

int qsgpu_relation_read_all(qsgpu_relation_t rel, uint64_t row_begin, uint64_t n_rows, void *const *host_out) {
  int st = sync_rows(rel);
  if (st) return st;
  if (row_begin + n_rows > rel->host_rows) { set_error(QSGPU_ERR_INVALID, "read outside relation"); return QSGPU_ERR_INVALID; }
  Device *d = device(rel->dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  if (n_rows == 0) return QSGPU_OK;
  if (rel->has_codes()) {   // base relations only; results of operators are always native
    for (uint32_t a = 0; a < rel->cols.size(); ++a) {
      st = qsgpu_relation_read(rel, a, row_begin, n_rows, host_out[a]);
      if (st) return st;
    }
    return QSGPU_OK;
  }
  size_t total = 0;
  for (size_t a = 0; a < rel->cols.size(); ++a) total += (n_rows * rel->attrs[a].width + 15) & ~static_cast<size_t>(15);
  if (total <= kReadScratchBytes && d->read_scratch) {
    // Small result: pack the columns with device-to-device copies (truly asynchronous), ONE transfer into
    // pinned memory, one wait.  A pageable destination would make every column copy a blocking call.
    std::lock_guard<std::mutex> lk(*d->read_mu);
    size_t off = 0;
    for (size_t a = 0; a < rel->cols.size(); ++a) {
      const size_t bytes = n_rows * rel->attrs[a].width;
      QS_CUDA(cudaMemcpyAsync(d->read_scratch + off, rel->cols[a] + row_begin * rel->attrs[a].width, bytes, cudaMemcpyDeviceToDevice, d->stream));
      off += (bytes + 15) & ~static_cast<size_t>(15);
    }
    QS_CUDA(cudaMemcpyAsync(d->read_pinned, d->read_scratch, total, cudaMemcpyDeviceToHost, d->stream));
    QS_CUDA(cudaStreamSynchronize(d->stream));
    off = 0;
    for (size_t a = 0; a < rel->cols.size(); ++a) {
      const size_t bytes = n_rows * rel->attrs[a].width;
      std::memcpy(host_out[a], d->read_pinned + off, bytes);
      off += (bytes + 15) & ~static_cast<size_t>(15);
    }
    return QSGPU_OK;
  }
  for (size_t a = 0; a < rel->cols.size(); ++a) {
    const uint32_t w = rel->attrs[a].width;
    QS_CUDA(cudaMemcpyAsync(host_out[a], rel->cols[a] + row_begin * w, n_rows * w, cudaMemcpyDeviceToHost, d->stream));
  }
  QS_CUDA(cudaStreamSynchronize(d->stream));
  return QSGPU_OK;
}

int qsgpu_relation_read_rows(qsgpu_relation_t rel, uint64_t max_rows, void *const *host_out, uint64_t *n_rows,
                             uint64_t *null_masks) {
  Device *d = device(rel->dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  if (rel->has_codes() || rel->cols.size() > static_cast<size_t>(kMaxCols)) { set_error(QSGPU_ERR_UNSUPPORTED, "read_rows: result relations only (native columns, <= 12 attributes)"); return QSGPU_ERR_UNSUPPORTED; }
  max_rows = std::min<uint64_t>(max_rows, rel->capacity);
  size_t total = 16 + (null_masks ? ((max_rows * 8 + 15) & ~static_cast<size_t>(15)) : 0);
  for (size_t a = 0; a < rel->cols.size(); ++a) total += (max_rows * rel->attrs[a].width + 15) & ~static_cast<size_t>(15);
  if (total > kReadScratchBytes || !d->read_scratch) { set_error(QSGPU_ERR_CAPACITY, "read_rows: result larger than the 256 KB landing buffer; use qsgpu_relation_read"); return QSGPU_ERR_CAPACITY; }
  ColDesc cols[kMaxCols];
  for (size_t a = 0; a < rel->cols.size(); ++a) { cols[a] = ColDesc{}; cols[a].ptr = rel->cols[a]; cols[a].width = rel->attrs[a].width; }
  std::lock_guard<std::mutex> lk(*d->read_mu);
  // one pack launch, one transfer into pinned memory, one wait: row count, error word, NULL masks and rows together
  QS_CUDA(launch_pack_rows(d->read_scratch, cols, static_cast<uint32_t>(rel->cols.size()), max_rows, rel->d_rows,
                           rel->d_nulls, null_masks != nullptr, d->d_error, d->stream));
  count_launch();
  QS_CUDA(cudaMemcpyAsync(d->read_pinned, d->read_scratch, total, cudaMemcpyDeviceToHost, d->stream));
  QS_CUDA(cudaStreamSynchronize(d->stream));
  uint64_t n = 0;
  uint32_t flag = 0;
  std::memcpy(&n, d->read_pinned, 8);
  std::memcpy(&flag, d->read_pinned + 8, 4);
  if (flag != 0) {
    set_error(static_cast<int>(flag), "a kernel reported a capacity overflow (hash table, group limit or output relation too small)");
    return static_cast<int>(flag);
  }
  rel->host_rows = std::min<uint64_t>(n, rel->capacity);
  rel->dirty = false;
  const uint64_t got = std::min<uint64_t>(rel->host_rows, max_rows);
  size_t off = 16;
  if (null_masks) {
    std::memcpy(null_masks, d->read_pinned + off, got * 8);
    off += (max_rows * 8 + 15) & ~static_cast<size_t>(15);
  }
  for (size_t a = 0; a < rel->cols.size(); ++a) {
    std::memcpy(host_out[a], d->read_pinned + off, got * rel->attrs[a].width);
    off += (max_rows * rel->attrs[a].width + 15) & ~static_cast<size_t>(15);
  }
  *n_rows = rel->host_rows;
  return QSGPU_OK;
}

int qsgpu_stage_block(qsgpu_relation_t rel, uint64_t n_rows, const qs_stage_desc *descs, uint32_t n_desc) {
  int st = sync_rows(rel);
  if (st) return st;
  Device *d = device(rel->dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  if (!rel->owns_memory) { set_error(QSGPU_ERR_INVALID, "cannot stage into a wrapped relation"); return QSGPU_ERR_INVALID; }
  if (rel->host_rows + n_rows > rel->capacity) { set_error(QSGPU_ERR_CAPACITY, "relation capacity exceeded while staging"); return QSGPU_ERR_CAPACITY; }
  if (n_desc != rel->attrs.size()) { set_error(QSGPU_ERR_INVALID, "qsgpu_stage_block must stage every attribute of the relation"); return QSGPU_ERR_INVALID; }
  if (rel->has_codes()) { set_error(QSGPU_ERR_UNSUPPORTED, "relations with dictionary-coded attributes are staged with qsgpu_stage_blocks"); return QSGPU_ERR_UNSUPPORTED; }
  KernelTimer timer(d, QS_K_STAGE);
  std::vector<void *> scratch;
  int rc = QSGPU_OK;
  for (uint32_t i = 0; i < n_desc && rc == QSGPU_OK; ++i) {
    const qs_stage_desc &s = descs[i];
    if (s.attr >= rel->attrs.size()) { set_error(QSGPU_ERR_INVALID, "stage: attribute out of range"); rc = QSGPU_ERR_INVALID; break; }
    if (s.null_kind != QS_NULL_NONE) { set_error(QSGPU_ERR_UNSUPPORTED, "stripes with NULLs are staged with qsgpu_stage_blocks"); rc = QSGPU_ERR_UNSUPPORTED; break; }
    const uint32_t w = rel->attrs[s.attr].width;
    char *dst = rel->cols[s.attr] + rel->host_rows * w;
    cudaError_t e = cudaSuccess;
    switch (s.encoding) {
      case QS_ENC_PLAIN:
        e = cudaMemcpyAsync(dst, s.host, n_rows * w, cudaMemcpyHostToDevice, d->stream);
        break;
      case QS_ENC_STRIDED: {
        void *tmp = nullptr;
        const size_t bytes = n_rows ? (n_rows - 1) * s.stride + w : 0;
        e = dev_malloc(&tmp, bytes + 16);
        if (e == cudaSuccess) { scratch.push_back(tmp); e = cudaMemcpyAsync(tmp, s.host, bytes, cudaMemcpyHostToDevice, d->stream); }
        if (e == cudaSuccess) { e = launch_decode_strided(dst, tmp, n_rows, s.stride, w, d->stream); count_launch(); }
        break;
      }
      case QS_ENC_DICT: {
        if (s.code_width != 1 && s.code_width != 2 && s.code_width != 4) { set_error(QSGPU_ERR_INVALID, "code width must be 1, 2 or 4"); rc = QSGPU_ERR_INVALID; break; }
        void *codes = nullptr, *dict = nullptr;
        e = dev_malloc(&codes, n_rows * s.code_width + 16);
        if (e == cudaSuccess) { scratch.push_back(codes); e = dev_malloc(&dict, static_cast<size_t>(s.dict_entries) * w + 16); }
        if (e == cudaSuccess) { scratch.push_back(dict); e = cudaMemcpyAsync(codes, s.host, n_rows * s.code_width, cudaMemcpyHostToDevice, d->stream); }
        if (e == cudaSuccess) e = cudaMemcpyAsync(dict, s.dict, static_cast<size_t>(s.dict_entries) * w, cudaMemcpyHostToDevice, d->stream);
        if (e == cudaSuccess) { e = launch_decode_dict(dst, codes, dict, n_rows, s.code_width, w, s.dict_entries, d->stream); count_launch(); }
        break;
      }
      case QS_ENC_TRUNCATED: {
        if ((w != 4 && w != 8) || (s.code_width != 1 && s.code_width != 2 && s.code_width != 4)) { set_error(QSGPU_ERR_INVALID, "truncation applies to INT/LONG with 1/2/4-byte codes"); rc = QSGPU_ERR_INVALID; break; }
        void *codes = nullptr;
        e = dev_malloc(&codes, n_rows * s.code_width + 16);
        if (e == cudaSuccess) { scratch.push_back(codes); e = cudaMemcpyAsync(codes, s.host, n_rows * s.code_width, cudaMemcpyHostToDevice, d->stream); }
        if (e == cudaSuccess) { e = launch_decode_truncated(dst, codes, n_rows, s.code_width, w, d->stream); count_launch(); }
        break;
      }
      default:
        set_error(QSGPU_ERR_INVALID, "unknown staging encoding");
        rc = QSGPU_ERR_INVALID;
    }
    // DateLit padding bytes are not initialised by the reference: canonicalise (see k_decode_segments)
    if (rc == QSGPU_OK && e == cudaSuccess && rel->attrs[s.attr].type == QS_DATE) { e = launch_zero_date_padding(dst, n_rows, d->stream); count_launch(); }
    // ... and CHAR(n) bytes behind the terminating NUL are not part of the value
    if (rc == QSGPU_OK && e == cudaSuccess && rel->attrs[s.attr].type == QS_CHAR && w > 1) { e = launch_zero_after_nul(dst, n_rows, w, d->stream); count_launch(); }
    if (rc == QSGPU_OK && e != cudaSuccess) rc = cuda_fail(e, "qsgpu_stage_block");
  }
  cudaStreamSynchronize(d->stream);     // host stripes may be unpinned / reused by the caller
  for (void *p : scratch) dev_free(p);
  if (rc != QSGPU_OK) return rc;
  return qsgpu_relation_set_num_rows(rel, rel->host_rows + n_rows);
}

// Shared body of qsgpu_stage_blocks (append) and qsgpu_stage_columns (fill attributes of rows that exist).
//
// Pipeline: the batch is cut into chunks of <= 64 MB of block images.  The images of chunk i travel on the
// device's COPY stream while chunk i-1 is decoded on the library stream (one decode launch per chunk, every
// stripe of the chunk in it).  Only the byte ranges of the stripes / dictionaries that are actually staged
// are copied (attributes marked QS_ENC_SKIP stay on the host: column pruning), adjacent ranges -- also
// across blocks that are contiguous in host memory, e.g. a buffer-pool slab -- merged into one copy.
static constexpr uint64_t kStageChunkBytes = 64ull << 20;
static constexpr uint64_t kStageMergeGap = 192ull << 10;

static int stage_impl(qsgpu_relation *rel, uint64_t first_row, bool append, uint32_t n_blocks,
                      const qs_block_image *blocks, uint32_t n_desc) {
  Device *d = device(rel->dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  if (!rel->owns_memory) { set_error(QSGPU_ERR_INVALID, "cannot stage into a wrapped relation"); return QSGPU_ERR_INVALID; }
  if (n_desc != rel->attrs.size()) { set_error(QSGPU_ERR_INVALID, "one stage descriptor per attribute of the relation is required (QS_ENC_SKIP leaves an attribute out)"); return QSGPU_ERR_INVALID; }
  if (n_blocks == 0) return QSGPU_OK;
  uint64_t total_rows = 0, image_bytes = 0;
  std::vector<uint64_t> img_off(n_blocks);
  for (uint32_t b = 0; b < n_blocks; ++b) {
    if (!blocks[b].host || !blocks[b].descs) { set_error(QSGPU_ERR_INVALID, "block image without memory / descriptors"); return QSGPU_ERR_INVALID; }
    img_off[b] = image_bytes;
    image_bytes += (blocks[b].bytes + 15) & ~15ull;
    total_rows += blocks[b].n_rows;
  }
  if (first_row + total_rows > rel->capacity) { set_error(QSGPU_ERR_CAPACITY, "relation capacity exceeded while staging"); return QSGPU_ERR_CAPACITY; }
  KernelTimer timer(d, QS_K_STAGE);
  char *d_img = nullptr;
  StageSeg *d_segs = nullptr;
  QS_CUDA(dev_malloc(&d_img, image_bytes + 64));

  struct Range { const char *host; uint64_t dev_off, bytes; };
  struct Chunk { uint32_t seg_begin, seg_end; size_t range_begin, range_end; uint64_t tiles; };
  std::vector<StageSeg> segs;
  std::vector<Range> ranges;
  std::vector<Chunk> chunks;
  segs.reserve(static_cast<size_t>(n_blocks) * n_desc);
  uint64_t row_base = first_row, chunk_bytes = 0;
  Chunk cur{0, 0, 0, 0, 0};
  int rc = QSGPU_OK;
  std::vector<std::pair<uint64_t, uint64_t>> need;          // (offset, bytes) inside one block image
  size_t remap_bytes = 0;                                   // re-coding tables of coded attributes, all blocks
  std::vector<std::pair<size_t, size_t>> remap_fix;         // (segment, offset of its table in `remap`)
  char *d_remap = nullptr;
  for (uint32_t b = 0; b < n_blocks && rc == QSGPU_OK; ++b) {
    const qs_block_image &B = blocks[b];
    const char *h0 = static_cast<const char *>(B.host);
    need.clear();
    for (uint32_t i = 0; i < n_desc; ++i) {
      const qs_stage_desc &s = B.descs[i];
      if (s.encoding == QS_ENC_SKIP) continue;
      if (s.attr >= rel->attrs.size()) { set_error(QSGPU_ERR_INVALID, "stage: attribute out of range"); rc = QSGPU_ERR_INVALID; break; }
      const uint32_t w = rel->attrs[s.attr].width;
      const char *hs = static_cast<const char *>(s.host);
      uint64_t bytes = 0;
      if (const uint32_t gcw = rel->code_width(s.attr)) {
        // Attribute held as codes of the relation-wide dictionary: the block's codes are re-coded through a
        // per-block table (block code -> relation code).  The table is built ON THE DEVICE (k_build_recode: one
        // binary search of the relation's dictionary per entry of the block's dictionary, which travels inside the
        // block image), so the host does no per-value work; the decode kernel then sees the table as a dictionary
        // whose "values" are gcw-byte codes.
        if (s.encoding != QS_ENC_DICT) { set_error(QSGPU_ERR_UNSUPPORTED, "a dictionary-coded attribute is staged from dictionary-compressed stripes only"); rc = QSGPU_ERR_UNSUPPORTED; break; }
        if (s.null_kind != QS_NULL_NONE) { set_error(QSGPU_ERR_UNSUPPORTED, "a NULL-able attribute is held at native width, not as codes of a relation-wide dictionary"); rc = QSGPU_ERR_UNSUPPORTED; break; }
        if (s.code_width != 1 && s.code_width != 2 && s.code_width != 4) { set_error(QSGPU_ERR_INVALID, "code width must be 1, 2 or 4"); rc = QSGPU_ERR_INVALID; break; }
        const char *hd = static_cast<const char *>(s.dict);
        bytes = B.n_rows * s.code_width;
        if (!hs || hs < h0 || hs + bytes > h0 + B.bytes || !hd || s.dict_entries == 0) { set_error(QSGPU_ERR_INVALID, "stage: stripe lies outside the block image"); rc = QSGPU_ERR_INVALID; break; }
        const qs_coded_attr &C = rel->coded[s.attr];
        const uint64_t dbytes = static_cast<uint64_t>(s.dict_entries) * w;
        if (hd < h0 || hd + dbytes > h0 + B.bytes) { set_error(QSGPU_ERR_INVALID, "stage: dictionary lies outside the block image"); rc = QSGPU_ERR_INVALID; break; }
        remap_bytes = (remap_bytes + 3) & ~static_cast<size_t>(3);
        const size_t roff = remap_bytes;
        remap_bytes += static_cast<size_t>(s.dict_entries) * gcw;
        StageSeg g{};
        g.dst = rel->cols[s.attr] + row_base * gcw;
        g.src = d_img + img_off[b] + (hs - h0);
        g.n_rows = B.n_rows;
        g.tile_begin = cur.tiles;
        g.encoding = QS_ENC_DICT;
        g.cw = s.code_width; g.vw = gcw; g.dict_entries = s.dict_entries;
        g.bdict = d_img + img_off[b] + (hd - h0);
        g.gdict = C.d_dict; g.g_entries = C.n_entries; g.qtype = rel->attrs[s.attr].type; g.bw = w;
        const bool code_al = s.code_width <= 1 || (reinterpret_cast<uintptr_t>(g.src) % s.code_width) == 0;
        g.aligned = ((reinterpret_cast<uintptr_t>(g.dst) % gcw) == 0 ? 1u : 0u) | (code_al ? 2u : 0u);
        if (B.n_rows) {
          remap_fix.emplace_back(segs.size(), roff);
          segs.push_back(g);
          cur.tiles += (B.n_rows + kStageTileRows - 1) / kStageTileRows;
          need.emplace_back(static_cast<uint64_t>(hs - h0), bytes);
          need.emplace_back(static_cast<uint64_t>(hd - h0), dbytes);
        }
        continue;
      }
      switch (s.encoding) {
        case QS_ENC_PLAIN: bytes = B.n_rows * w; break;
        case QS_ENC_STRIDED: bytes = B.n_rows ? (B.n_rows - 1) * s.stride + w : 0; break;
        case QS_ENC_DICT: case QS_ENC_TRUNCATED:
          if (s.code_width != 1 && s.code_width != 2 && s.code_width != 4) { set_error(QSGPU_ERR_INVALID, "code width must be 1, 2 or 4"); rc = QSGPU_ERR_INVALID; }
          if (s.encoding == QS_ENC_TRUNCATED && w != 4 && w != 8) { set_error(QSGPU_ERR_INVALID, "truncation applies to INT/LONG"); rc = QSGPU_ERR_INVALID; }
          bytes = B.n_rows * s.code_width;
          break;
        default: set_error(QSGPU_ERR_INVALID, "unknown staging encoding"); rc = QSGPU_ERR_INVALID;
      }
      if (rc != QSGPU_OK) break;
      if (!hs || hs < h0 || hs + bytes > h0 + B.bytes) { set_error(QSGPU_ERR_INVALID, "stage: stripe lies outside the block image"); rc = QSGPU_ERR_INVALID; break; }
      StageSeg g{};
      g.dst = rel->cols[s.attr] + row_base * w;
      g.src = d_img + img_off[b] + (hs - h0);
      g.n_rows = B.n_rows;
      g.tile_begin = cur.tiles;
      g.encoding = s.encoding;
      g.cw = s.code_width; g.vw = w; g.stride = s.stride; g.dict_entries = s.dict_entries;
      const bool pow2 = w == 4 || w == 8;
      bool val_al = pow2 && (reinterpret_cast<uintptr_t>(g.dst) % w) == 0;
      if (s.encoding == QS_ENC_PLAIN) val_al = val_al && (reinterpret_cast<uintptr_t>(g.src) % w) == 0;
      if (s.encoding == QS_ENC_DICT) {
        const char *hd = static_cast<const char *>(s.dict);
        const uint64_t dbytes = static_cast<uint64_t>(s.dict_entries) * w;
        if (!hd || s.dict_entries == 0 || hd < h0 || hd + dbytes > h0 + B.bytes) { set_error(QSGPU_ERR_INVALID, "stage: dictionary lies outside the block image"); rc = QSGPU_ERR_INVALID; break; }
        g.dict = d_img + img_off[b] + (hd - h0);
        val_al = val_al && (reinterpret_cast<uintptr_t>(g.dict) % w) == 0;
        need.emplace_back(static_cast<uint64_t>(hd - h0), dbytes);
      }
      const bool code_al = s.code_width <= 1 || (reinterpret_cast<uintptr_t>(g.src) % s.code_width) == 0;
      g.aligned = (val_al ? 1u : 0u) | (code_al ? 2u : 0u) | (rel->attrs[s.attr].type == QS_DATE ? 4u : 0u) |
                  ((rel->attrs[s.attr].type == QS_CHAR && w > 1) ? 8u : 0u);
      if (s.null_kind != QS_NULL_NONE && B.n_rows) {
        // the stripe's own NULL representation -> the relation's per-row mask (+ zero bytes for the value)
        if (s.attr >= 64 || !((rel->nullable_mask >> s.attr) & 1ull) || !rel->d_nulls) { set_error(QSGPU_ERR_INVALID, "stage: NULL information for an attribute not declared with qsgpu_relation_set_nullable"); rc = QSGPU_ERR_INVALID; break; }
        const char *hb = static_cast<const char *>(s.null_bitmap);
        uint64_t nbytes = 0;
        switch (s.null_kind) {
          case QS_NULL_CODE:
            if (s.encoding != QS_ENC_DICT) { set_error(QSGPU_ERR_INVALID, "stage: a NULL code belongs to a dictionary-compressed stripe"); rc = QSGPU_ERR_INVALID; }
            break;
          case QS_NULL_BITMAP:
            nbytes = (((s.null_arg + (B.n_rows - 1) * static_cast<uint64_t>(s.null_stride)) >> 6) + 1) * 8;
            break;
          case QS_NULL_SLOT_WORD:
            if ((s.null_width != 1 && s.null_width != 2 && s.null_width != 4 && s.null_width != 8) || s.null_arg >= 8 * s.null_width) { set_error(QSGPU_ERR_INVALID, "stage: bad NULL word width / bit"); rc = QSGPU_ERR_INVALID; }
            nbytes = (B.n_rows - 1) * static_cast<uint64_t>(s.null_stride) + s.null_width;
            break;
          default: set_error(QSGPU_ERR_INVALID, "unknown NULL representation"); rc = QSGPU_ERR_INVALID;
        }
        if (rc != QSGPU_OK) break;
        if (nbytes) {
          if (!hb || hb < h0 || hb + nbytes > h0 + B.bytes) { set_error(QSGPU_ERR_INVALID, "stage: NULL bitmap lies outside the block image"); rc = QSGPU_ERR_INVALID; break; }
          g.null_src = reinterpret_cast<const unsigned char *>(d_img + img_off[b] + (hb - h0));
          need.emplace_back(static_cast<uint64_t>(hb - h0), nbytes);
        }
        g.null_kind = s.null_kind; g.null_arg = s.null_arg; g.null_stride = s.null_stride; g.null_width = s.null_width;
        g.null_dst = rel->d_nulls + row_base;
        g.null_bit = 1ull << s.attr;
      }
      if (B.n_rows) {
        segs.push_back(g);
        cur.tiles += (B.n_rows + kStageTileRows - 1) / kStageTileRows;
        need.emplace_back(static_cast<uint64_t>(hs - h0), bytes);
      }
    }
    if (rc != QSGPU_OK) break;
    // byte ranges of this block that have to travel: sorted, neighbours closer than 4 KB merged
    std::sort(need.begin(), need.end());
    uint64_t block_copied = 0;
    for (size_t i = 0; i < need.size();) {
      uint64_t lo = need[i].first, hi = need[i].first + need[i].second;
      size_t j = i + 1;
      while (j < need.size() && need[j].first <= hi + kStageMergeGap) { hi = std::max(hi, need[j].first + need[j].second); ++j; }
      hi = std::min<uint64_t>(hi, B.bytes);
      const Range r{h0 + lo, img_off[b] + lo, hi - lo};
      // Same slab, same host-to-device displacement, and a hole smaller than kStageMergeGap: extend the
      // previous copy over the hole.  A DMA descriptor costs ~3 us, i.e. ~150 KB of PCIe time, so skipping
      // anything smaller than that (a narrow stripe, a dictionary) is slower than copying it.
      bool merged = false;
      if (ranges.size() > cur.range_begin) {
        Range &p = ranges.back();
        const char *p_end = p.host + p.bytes;
        if (r.host >= p_end && static_cast<uint64_t>(r.host - p_end) <= kStageMergeGap &&
            static_cast<uint64_t>(r.host - p.host) == r.dev_off - p.dev_off) {
          p.bytes = static_cast<uint64_t>(r.host - p.host) + r.bytes;
          merged = true;
        }
      }
      if (!merged) ranges.push_back(r);
      block_copied += hi - lo;
      i = j;
    }
    row_base += B.n_rows;
    chunk_bytes += block_copied;
    if (chunk_bytes >= kStageChunkBytes || b + 1 == n_blocks) {
      cur.seg_end = static_cast<uint32_t>(segs.size());
      cur.range_end = ranges.size();
      chunks.push_back(cur);
      cur = Chunk{cur.seg_end, cur.seg_end, ranges.size(), ranges.size(), 0};
      chunk_bytes = 0;
    }
  }
  cudaEvent_t ev = nullptr;
  if (rc == QSGPU_OK && !segs.empty()) {
    cudaError_t ce = cudaSuccess;
    if (remap_bytes) {
      ce = dev_malloc(&d_remap, remap_bytes + 16);
      for (const auto &f : remap_fix) segs[f.first].dict = d_remap + f.second;
    }
    if (ce == cudaSuccess) ce = dev_malloc(&d_segs, segs.size() * sizeof(StageSeg));
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(d_segs, segs.data(), segs.size() * sizeof(StageSeg), cudaMemcpyHostToDevice, d->stream);
    // the image buffer may be a recycled block that earlier work on the library stream still reads
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (ce == cudaSuccess) ce = cudaEventRecord(ev, d->stream);
    if (ce == cudaSuccess) ce = cudaStreamWaitEvent(d->copy_stream, ev, 0);
    std::vector<cudaEvent_t> done(chunks.size(), nullptr);
    for (size_t c = 0; c < chunks.size() && ce == cudaSuccess; ++c) {
      const Chunk &C = chunks[c];
      for (size_t r = C.range_begin; r < C.range_end && ce == cudaSuccess; ++r)
        ce = cudaMemcpyAsync(d_img + ranges[r].dev_off, ranges[r].host, ranges[r].bytes, cudaMemcpyHostToDevice, d->copy_stream);
      if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&done[c], cudaEventDisableTiming);
      if (ce == cudaSuccess) ce = cudaEventRecord(done[c], d->copy_stream);
      if (ce == cudaSuccess) ce = cudaStreamWaitEvent(d->stream, done[c], 0);
      if (ce == cudaSuccess && C.seg_end > C.seg_begin && remap_bytes) {   // block codes -> relation codes tables
        ce = launch_build_recode(d_segs + C.seg_begin, C.seg_end - C.seg_begin, d->d_error, d->stream);
        count_launch();
      }
      if (ce == cudaSuccess && C.seg_end > C.seg_begin) {
        ce = launch_decode_segments(d_segs + C.seg_begin, C.seg_end - C.seg_begin, C.tiles, d->sm_count, d->stream);
        count_launch();
      }
    }
    if (ce != cudaSuccess) rc = cuda_fail(ce, "qsgpu_stage_blocks");
    cudaStreamSynchronize(d->copy_stream);
    cudaStreamSynchronize(d->stream);     // host block memory and `segs` may be reused by the caller
    for (cudaEvent_t e : done) if (e) cudaEventDestroy(e);
    if (ev) cudaEventDestroy(ev);
  }
  dev_free(d_img);
  dev_free(d_segs);
  dev_free(d_remap);
  if (rc == QSGPU_OK && remap_bytes) {
    rc = check_device_error(d);
    if (rc == QSGPU_ERR_INVALID) set_error(rc, "stage: a block dictionary value is missing from the relation's dictionary");
  }
  if (rc != QSGPU_OK) return rc;
  const uint64_t rows_after = std::max<uint64_t>(rel->host_rows, first_row + total_rows);
  if (!append && rows_after == rel->host_rows) return QSGPU_OK;
  return qsgpu_relation_set_num_rows(rel, rows_after);
}

int qsgpu_stage_blocks(qsgpu_relation_t rel, uint32_t n_blocks, const qs_block_image *blocks, uint32_t n_desc) {
  int st = sync_rows(rel);
  if (st) return st;
  return stage_impl(rel, rel->host_rows, true, n_blocks, blocks, n_desc);
}

int qsgpu_stage_columns(qsgpu_relation_t rel, uint64_t first_row, uint32_t n_blocks, const qs_block_image *blocks,
                        uint32_t n_desc) {
  int st = sync_rows(rel);
  if (st) return st;
  if (first_row > rel->host_rows) { set_error(QSGPU_ERR_INVALID, "qsgpu_stage_columns: first_row beyond the relation's rows"); return QSGPU_ERR_INVALID; }
  return stage_impl(rel, first_row, false, n_blocks, blocks, n_desc);
}

/* ------------------------------------------------------------ LIP filters */
int qsgpu_lip_create(int dev, uint32_t kind, uint32_t attr_type, int64_t min_value, int64_t max_value,
                     uint64_t cardinality, int is_anti, qsgpu_lip_t *out) {
  Device *d = device(dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  if (attr_type != QS_INT && attr_type != QS_LONG) { set_error(QSGPU_ERR_UNSUPPORTED, "LIP filters are defined over INT/LONG"); return QSGPU_ERR_UNSUPPORTED; }
  uint64_t bits;
  if (kind == QS_LIP_BITVECTOR_EXACT) {
    if (max_value < min_value) { set_error(QSGPU_ERR_INVALID, "exact filter needs max >= min"); return QSGPU_ERR_INVALID; }
    bits = static_cast<uint64_t>(max_value - min_value) + 1;
  } else if (kind == QS_LIP_SINGLE_IDENTITY_HASH) {
    if (cardinality == 0) { set_error(QSGPU_ERR_INVALID, "hash filter needs cardinality >= 1"); return QSGPU_ERR_INVALID; }
    bits = cardinality;
  } else { set_error(QSGPU_ERR_INVALID, "unknown LIP filter kind"); return QSGPU_ERR_INVALID; }
  std::unique_ptr<qsgpu_lip> f(new qsgpu_lip);
  f->dev = dev;
  f->attr_type = attr_type;
  f->n_words = (bits + 63) / 64;
  f->d.kind = kind;
  f->d.min_value = min_value;
  f->d.max_value = max_value;
  f->d.cardinality = cardinality;
  f->d.is_anti = is_anti ? 1 : 0;
  QS_CUDA(dev_malloc(&f->d.words, f->n_words * 8 + 64));
  QS_CUDA(cudaMemsetAsync(f->d.words, 0, f->n_words * 8 + 64, d->stream));
  f->d.stats = reinterpret_cast<unsigned long long *>(f->d.words + f->n_words);      // the 64 bytes behind the bit words
  f->d.n_words = f->n_words;
  *out = f.release();
  return QSGPU_OK;
}
int qsgpu_lip_destroy(qsgpu_lip_t lip) {
  if (!lip) return QSGPU_OK;
  device(lip->dev);
  dev_free(lip->d.words);
  delete lip;
  return QSGPU_OK;
}
int qsgpu_lip_num_words(qsgpu_lip_t lip, uint64_t *n_words) { *n_words = lip->n_words; return QSGPU_OK; }
int qsgpu_lip_read(qsgpu_lip_t lip, uint64_t *host_words) {
  return qsgpu_memcpy_d2h(lip->dev, host_words, lip->d.words, lip->n_words * 8);
}
int qsgpu_lip_device_words(qsgpu_lip_t lip, void **dptr) { *dptr = lip->d.words; return QSGPU_OK; }
int qsgpu_lip_probe_stats(qsgpu_lip_t lip, uint64_t *probes, uint64_t *misses) {
  Device *d = device(lip->dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  unsigned long long v[2] = {0, 0};
  QS_CUDA(cudaMemcpyAsync(v, lip->d.stats, 16, cudaMemcpyDeviceToHost, d->stream));
  QS_CUDA(cudaStreamSynchronize(d->stream));
  if (probes) *probes = v[0];
  if (misses) *misses = v[1];
  return QSGPU_OK;
}

/* --------------------------------------------- shared scan-side lowering */
// key_attrs: attributes whose NULL rows take no part in the operator (join / LIP-build keys: HashTable::
// putValueAccessor and getAllFromValueAccessor skip NULL keys, storage/HashTable.hpp:1384,1903).
static int lower_scan_predicate(Lowering &L, const qs_scan *scan, uint64_t key_attrs = 0) {
  bool have = false;
  if (scan->predicate_root >= 0) { L.lower_pred(scan->predicate_root); have = true; }
  if (const uint64_t nb = key_attrs & scan->input->nullable_mask) { L.push_notnull(nb, have); have = true; }
  for (uint32_t i = 0; i < scan->n_lip_probe; ++i) {
    const uint32_t pa = scan->lip_probe[i].attr;
    // A NULL probe value passes NO filter, an anti filter included: filterBatchInternal<true> skips the tuple before it
    // looks at the filter (utility/lip_filter/BitVectorExactFilter.hpp:130-146; pinned by the reference-class golden,
    // tests/golden/reference_expressions.json "nullable_exact_int_anti").  lower_lip_probe ANDs "not NULL" in front.
    L.lower_lip_probe(i, pa, have);
    have = true;
  }
  L.mark_pred_end();
  if (!L.ok()) { set_error(L.status, L.err); return L.status; }
  return QSGPU_OK;
}

static int fill_lip_build(uint32_t n, const qs_lip_ref *refs, Lowering &L, const qsgpu_relation *rel, SinkDesc *K) {
  if (n > static_cast<uint32_t>(kMaxLip)) { set_error(QSGPU_ERR_UNSUPPORTED, "more than kMaxLip LIP filters built by one scan"); return QSGPU_ERR_UNSUPPORTED; }
  K->n_lip_build = n;
  for (uint32_t i = 0; i < n; ++i) {
    if (!refs[i].lip || refs[i].lip->dev != rel->dev || refs[i].attr >= rel->attrs.size()) { set_error(QSGPU_ERR_INVALID, "bad LIP build reference"); return QSGPU_ERR_INVALID; }
    const uint8_t lt = vtype_of(rel->attrs[refs[i].attr].type);
    if (lt != V_I32 && lt != V_I64) { set_error(QSGPU_ERR_UNSUPPORTED, "LIP filters take INT/LONG attributes"); return QSGPU_ERR_UNSUPPORTED; }
    // one target: the caller AND-ed "attribute is not NULL" into the scan predicate; several targets over different
    // NULL-able attributes would each need their own row set
    if (n > 1 && refs[i].attr < 64 && ((rel->nullable_mask >> refs[i].attr) & 1ull)) { set_error(QSGPU_ERR_UNSUPPORTED, "several LIP filters built from NULL-able attributes by one scan"); return QSGPU_ERR_UNSUPPORTED; }
    K->lip_build[i] = refs[i].lip->d;
    K->lip_build_col[i] = static_cast<uint16_t>(L.stage_attr(refs[i].attr));
    K->lip_build_ltype[i] = lt;
  }
  if (!L.ok()) { set_error(L.status, L.err); return L.status; }
  return QSGPU_OK;
}

// Lowers the projection list; an attribute root is a raw pass-through copy.
static int lower_projection(Lowering &L, uint32_t n_project, const int32_t *roots, qsgpu_relation *output, SinkDesc *K) {
  if (n_project > static_cast<uint32_t>(kMaxOut)) { set_error(QSGPU_ERR_UNSUPPORTED, "more than kMaxOut projected columns"); return QSGPU_ERR_UNSUPPORTED; }
  if (n_project != output->attrs.size()) { set_error(QSGPU_ERR_INVALID, "projection list does not match the output relation"); return QSGPU_ERR_INVALID; }
  K->n_out = n_project;
  uint64_t null_cols = 0;
  for (uint32_t j = 0; j < n_project; ++j) {
    const qs_node *n = L.node(roots[j]);
    if (!n) break;
    K->out[j] = output->cols[j];
    K->out_width[j] = static_cast<uint8_t>(output->attrs[j].width);
    // a pass-through projection of a dictionary-coded numeric attribute is emitted as a value (dictionary look-up in
    // registers) rather than copied out of a native tile, so the tile need not be decoded in shared memory
    const bool coded_value = n->kind == QS_N_ATTRIBUTE && n->b != 2 && L.rel && static_cast<uint32_t>(n->a) < L.rel->attrs.size() &&
                             L.rel->code_width(static_cast<uint32_t>(n->a)) != 0 && n->type <= QS_DOUBLE &&
                             n->type == L.rel->attrs[n->a].type && output->attrs[j].type == n->type;
    if (n->kind == QS_N_ATTRIBUTE && !coded_value) {
      const qsgpu_relation *src = n->b == 2 ? L.build_rel : L.rel;
      if (!src || static_cast<uint32_t>(n->a) >= src->attrs.size() ||
          src->attrs[n->a].width != output->attrs[j].width) { set_error(QSGPU_ERR_INVALID, "projected attribute does not match the output column width"); return QSGPU_ERR_INVALID; }
      Instr in{};
      in.arg = static_cast<uint16_t>(j);
      if (n->b == 2) { in.op = OP_EMIT_RAW_BUILD; in.flags = static_cast<uint8_t>(L.build_attr(static_cast<uint32_t>(n->a))); }
      else { in.op = OP_EMIT_RAW; in.flags = static_cast<uint8_t>(L.stage_attr(static_cast<uint32_t>(n->a))); }
      L.push(in);
    } else {
      const uint8_t t = L.lower_scalar(roots[j]);
      uint8_t to = vtype_of(output->attrs[j].type);
      if (to == 0xff || to == V_DATE) { set_error(QSGPU_ERR_INVALID, "expression projected into a CHAR/DATE column"); return QSGPU_ERR_INVALID; }
      L.lower_cast_acc(t, to);
      Instr in{};
      in.op = OP_EMIT; in.type = to; in.arg = static_cast<uint16_t>(j);
      L.push(in);
    }
    // NULL-ness of the projected column: a scalar over a NULL attribute is NULL (its stored bytes are whatever
    // the arithmetic on the zero bytes gave; the mask is what counts, as in the reference's ColumnVector)
    if (const uint64_t nb = L.null_bits(roots[j])) {
      L.lower_emit_null(j, nb);
      null_cols |= 1ull << j;
    }
    if (const uint64_t nbb = L.build_null_bits(roots[j])) {      // ... or the matched build row's
      L.lower_emit_null_build(j, nbb);
      null_cols |= 1ull << j;
    }
  }
  L.finish();
  if (!L.ok()) { set_error(L.status, L.err); return L.status; }
  if (null_cols) {
    if (!t_sc) { const int ns = ensure_null_mask(output, device(output->dev)); if (ns) return ns; }
    output->nullable_mask |= null_cols;
    K->null_out = output->d_nulls;
  }
  return QSGPU_OK;
}

int qsgpu_build_lip_filter(const qs_scan *scan, uint32_t n_build, const qs_lip_ref *build) {
  qsgpu_relation *rel = scan->input;
  Device *d = device(rel->dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  Lowering L(scan->exprs, rel);
  uint64_t key_mask = 0;
  for (uint32_t i = 0; i < n_build; ++i) if (build[i].attr < 64) key_mask |= 1ull << build[i].attr;
  int st = lower_scan_predicate(L, scan, key_mask);
  if (st) return st;
  SinkDesc K{};
  K.error_flag = d->d_error;
  st = fill_lip_build(n_build, build, L, rel, &K);
  if (st) return st;
  L.finish();
  ScanDesc S;
  fill_scan(rel, scan->row_begin, scan->row_end, L, &S);
  st = fill_lips(scan->n_lip_probe, scan->lip_probe, rel, rel->dev, &S);
  if (st) return st;
  ScanPlan plan;
  st = plan_scan(d, &S, kCompactSmemBytes, &plan);
  if (st) return st;
  JitKernel *kern = nullptr;
  st = query_kernel(JF_SELECT, S, L.P, plan, nullptr, &K, nullptr, 1, &kern);
  if (st) return st;
  KernelTimer timer(d, QS_K_LIP);
  QS_CUDA(launch_query_kernel(d, kern, S, L.P, plan, nullptr, &K, nullptr));
  count_launch();
  return QSGPU_OK;
}

int qsgpu_select(const qs_scan *scan, uint32_t n_project, const int32_t *project_roots, qsgpu_relation_t output) {
  qsgpu_relation *rel = scan->input;
  Device *d = device(rel->dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  if (output->dev != rel->dev) { set_error(QSGPU_ERR_INVALID, "output relation on another device"); return QSGPU_ERR_INVALID; }
  Lowering L(scan->exprs, rel);
  int st = lower_scan_predicate(L, scan);
  if (st) return st;
  SinkDesc K{};
  K.error_flag = d->d_error;
  K.capacity = output->capacity;
  K.counter = output->d_rows;
  st = lower_projection(L, n_project, project_roots, output, &K);
  if (st) return st;
  ScanDesc S;
  fill_scan(rel, scan->row_begin, scan->row_end, L, &S);
  st = fill_lips(scan->n_lip_probe, scan->lip_probe, rel, rel->dev, &S);
  if (st) return st;
  ScanPlan plan;
  st = plan_scan(d, &S, kCompactSmemBytes, &plan);
  if (st) return st;
  JitKernel *kern = nullptr;
  st = query_kernel(JF_SELECT, S, L.P, plan, nullptr, &K, nullptr, 1, &kern);
  if (st) return st;
  KernelTimer timer(d, QS_K_SELECT);
  QS_CUDA(launch_query_kernel(d, kern, S, L.P, plan, nullptr, &K, nullptr));
  count_launch();
  output->dirty = true;
  return QSGPU_OK;
}

/* ------------------------------------------------------------- aggregation */
static bool is_fp(uint8_t v) { return v == V_F32 || v == V_F64; }

int qsgpu_agg_create(const qs_agg_spec *spec, qsgpu_agg_state_t *out) {
  Device *d = device(spec->dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  std::unique_ptr<qsgpu_agg_state> s(new qsgpu_agg_state);
  s->dev = spec->dev;
  // ThreadPrivateCompactKeyHashTable is chosen by the reference for up to 10,000 estimated groups
  // (StarSchemaSimpleCostModel.cpp:722); the compact-key kernels keep at most kCompactMaxGroups (256) groups
  // per CTA and per state, so a larger estimate takes the hash-table strategy (any key <= 8 bytes fits it).
  const uint32_t strategy = (spec->strategy == QS_AGG_COMPACT_KEY && spec->estimated_num_entries > static_cast<uint64_t>(kCompactMaxGroups))
                                ? static_cast<uint32_t>(QS_AGG_SEPARATE_CHAINING) : spec->strategy;
  s->strategy = strategy;
  if (spec->exprs) {
    s->nodes.assign(spec->exprs->nodes, spec->exprs->nodes + spec->exprs->n_nodes);
    s->str_pool.assign(spec->exprs->str_pool ? spec->exprs->str_pool : "", spec->exprs->str_pool ? spec->exprs->str_pool_bytes : 0);
  }
  s->exprs.nodes = s->nodes.data();
  s->exprs.n_nodes = static_cast<uint32_t>(s->nodes.size());
  s->exprs.str_pool = s->str_pool.data();
  s->exprs.str_pool_bytes = static_cast<uint32_t>(s->str_pool.size());
  s->predicate_root = spec->predicate_root;
  s->aggregates.assign(spec->aggregates, spec->aggregates + spec->n_aggregates);
  s->group_by_roots.assign(spec->group_by_roots, spec->group_by_roots + spec->n_group_by);
  s->estimated = spec->estimated_num_entries;
  if (spec->n_aggregates > static_cast<uint32_t>(kMaxOut)) { set_error(QSGPU_ERR_UNSUPPORTED, "too many aggregates"); return QSGPU_ERR_UNSUPPORTED; }

  AggDesc &A = s->A;
  A.strategy = strategy;
  A.error_flag = d->d_error;
  // ---- value words: one per SUM/AVG/MIN/MAX, COUNT reads the row-count word
  Lowering typer(&s->exprs, nullptr);
  uint32_t n_agg = 0;
  s->nn_word.assign(s->aggregates.size(), 0);
  for (const qs_aggregate &a : s->aggregates) {
    if (a.function == QS_AGG_COUNT) { s->value_word.push_back(0); s->arg_vtype.push_back(V_I64); continue; }
    if (a.argument_root < 0) { set_error(QSGPU_ERR_INVALID, "aggregate without an argument"); return QSGPU_ERR_INVALID; }
    if (n_agg >= static_cast<uint32_t>(kMaxAgg)) { set_error(QSGPU_ERR_UNSUPPORTED, "more than kMaxAgg value aggregates in one state"); return QSGPU_ERR_UNSUPPORTED; }
    const qs_node &root = s->nodes[a.argument_root];
    uint8_t vt = typer.scalar_vtype(a.argument_root);
    if (!typer.ok()) { set_error(typer.status, typer.err); return typer.status; }
    if ((root.kind == QS_N_ATTRIBUTE || root.kind == QS_N_LITERAL) && root.type == QS_DATE) { set_error(QSGPU_ERR_UNSUPPORTED, "MIN/MAX over DATE is not lowered"); return QSGPU_ERR_UNSUPPORTED; }
    uint8_t own = vt;
    if (root.kind == QS_N_ATTRIBUTE || root.kind == QS_N_LITERAL) own = vtype_of(root.type);
    uint8_t kind;
    switch (a.function) {
      case QS_AGG_SUM: case QS_AGG_AVG: kind = is_fp(vt) ? AK_SUM_F64 : AK_SUM_I64; break;
      case QS_AGG_MIN: kind = is_fp(vt) ? AK_MIN_F64 : AK_MIN_I64; break;
      case QS_AGG_MAX: kind = is_fp(vt) ? AK_MAX_F64 : AK_MAX_I64; break;
      default: set_error(QSGPU_ERR_INVALID, "unknown aggregate function"); return QSGPU_ERR_INVALID;
    }
    A.kind[n_agg] = kind;
    s->value_word.push_back(static_cast<int>(1 + n_agg));
    s->arg_vtype.push_back(own);
    ++n_agg;
  }
  // NULL-able arguments: one SUM_I64 word per distinct argument counts the rows whose argument is not NULL
  // (aggregates over the same argument share it); COUNT(x) reads that word, AVG divides by it, and SUM / MIN / MAX
  // are NULL where it is zero.  It merges across CTAs, work orders and GPUs like any other integer sum.
  for (size_t j = 0; j < s->aggregates.size(); ++j) {
    if (j >= 64 || !((spec->nullable_arguments >> j) & 1ull) || s->aggregates[j].argument_root < 0) continue;
    for (size_t i = 0; i < j; ++i)
      if (s->nn_word[i] && s->aggregates[i].argument_root == s->aggregates[j].argument_root) { s->nn_word[j] = s->nn_word[i]; break; }
    if (!s->nn_word[j]) {
      if (n_agg >= static_cast<uint32_t>(kMaxAgg)) { set_error(QSGPU_ERR_UNSUPPORTED, "more than kMaxAgg state words (value aggregates + non-NULL counts) in one state"); return QSGPU_ERR_UNSUPPORTED; }
      A.kind[n_agg] = AK_SUM_I64;
      s->nn_word[j] = static_cast<int>(1 + n_agg);
      ++n_agg;
    }
    if (s->aggregates[j].function == QS_AGG_COUNT) s->value_word[j] = s->nn_word[j];
  }
  A.n_agg = n_agg;
  A.words = n_agg + 1;
  // ---- group-by keys (attribute nodes)
  uint32_t key_bytes = 0;
  A.n_key_cols = static_cast<uint32_t>(s->group_by_roots.size());
  if (A.n_key_cols > static_cast<uint32_t>(kMaxKeyCols)) { set_error(QSGPU_ERR_UNSUPPORTED, "too many group-by attributes"); return QSGPU_ERR_UNSUPPORTED; }
  for (uint32_t k = 0; k < A.n_key_cols; ++k) {
    const int32_t r = s->group_by_roots[k];
    if (r < 0 || static_cast<size_t>(r) >= s->nodes.size() || s->nodes[r].kind != QS_N_ATTRIBUTE) { set_error(QSGPU_ERR_UNSUPPORTED, "group-by expressions must be attributes"); return QSGPU_ERR_UNSUPPORTED; }
    qs_attr ka{s->nodes[r].type, attr_width(s->nodes[r].type, s->nodes[r].width)};
    s->key_attrs.push_back(ka);
    s->key_attr_ids.push_back(static_cast<uint32_t>(s->nodes[r].a));
    A.key_width[k] = static_cast<uint8_t>(ka.width);
    A.key_off[k] = static_cast<uint8_t>(key_bytes);
    key_bytes += ka.width;
  }
  A.key_words = std::max<uint32_t>(1, (key_bytes + 7) / 8);

  switch (strategy) {
    case QS_AGG_SINGLE_STATE:
      if (A.n_key_cols != 0) { set_error(QSGPU_ERR_INVALID, "SINGLE_STATE with GROUP BY"); return QSGPU_ERR_INVALID; }
      A.partial_rows = 1;
      break;
    case QS_AGG_COMPACT_KEY:
      // ThreadPrivateCompactKeyHashTable: keys fit one 64-bit code (…CompactKeyHashTable.hpp:113-116)
      if (A.n_key_cols == 0 || key_bytes > 8) { set_error(QSGPU_ERR_INVALID, "COMPACT_KEY needs 1..8 key bytes"); return QSGPU_ERR_INVALID; }
      A.partial_rows = kCompactMaxGroups;
      break;
    case QS_AGG_SEPARATE_CHAINING:
      if (A.n_key_cols == 0 || key_bytes > 8u * kMaxKeyWords) { set_error(QSGPU_ERR_UNSUPPORTED, "group-by key wider than 32 bytes"); return QSGPU_ERR_UNSUPPORTED; }
      break;
    case QS_AGG_COLLISION_FREE:
      if (A.n_key_cols != 1 || (s->key_attrs[0].type != QS_INT && s->key_attrs[0].type != QS_LONG) || spec->collision_free_max_key < 0) {
        set_error(QSGPU_ERR_INVALID, "COLLISION_FREE needs one INT/LONG key and a non-negative max key");
        return QSGPU_ERR_INVALID;
      }
      for (const qs_aggregate &a : s->aggregates)
        if (a.function != QS_AGG_SUM && a.function != QS_AGG_COUNT && a.function != QS_AGG_AVG) {
          // reference restricts this table to COUNT/SUM (StarSchemaSimpleCostModel.cpp:614-709)
          set_error(QSGPU_ERR_INVALID, "COLLISION_FREE supports COUNT/SUM/AVG only");
          return QSGPU_ERR_INVALID;
        }
      break;
    default: set_error(QSGPU_ERR_INVALID, "unknown aggregation strategy"); return QSGPU_ERR_INVALID;
  }

  if (t_sc) { *out = s.release(); return QSGPU_OK; }   // selfcheck: description only, no device memory
  if (strategy == QS_AGG_SINGLE_STATE || strategy == QS_AGG_COMPACT_KEY) {
    // one partial row set per CTA of the persistent grid: up to 4 resident CTAs per SM for states with a few
    // partial rows (Q6 on dictionary codes: 12 KB tiles, 32 registers), 2 for the 256-group compact-key states
    // (their kernels hold the hot groups' sums in registers and never fit more than 2)
    s->max_ctas = static_cast<uint32_t>(d->sm_count) * (A.partial_rows <= 8 ? 4u : 2u);
    const size_t prow = static_cast<size_t>(A.partial_rows) * A.words * 8;
    QS_CUDA(dev_malloc(&A.partials, prow * s->max_ctas));
    // Everything else lives in ONE control block, brought to its initial contents by ONE launch (this call is on
    // the critical path of the query: the scan cannot start before its state exists):
    //   [n_groups 256 B][done ticket 256 B][index count 256 B][dir_keys][dir_gid][states | packed keys]
    // The [states | packed keys] tail is contiguous on purpose: it is the send buffer of the cross-GPU merge
    // (qsgpu_agg_merge_all all-gathers it as it lies, no packing pass).
    A.dir_cap = 1024;
    const size_t ctl_bytes = 768 + A.dir_cap * 8 + A.dir_cap * 4 + prow + static_cast<size_t>(A.partial_rows) * 8;
    QS_CUDA(dev_malloc(&s->ctl, ctl_bytes));
    char *p = s->ctl;
    A.n_groups = reinterpret_cast<uint32_t *>(p); p += 256;
    s->d_done = reinterpret_cast<unsigned int *>(p); p += 256;
    s->d_idx_count = reinterpret_cast<unsigned long long *>(p); p += 256;
    A.dir_keys = reinterpret_cast<uint64_t *>(p); p += A.dir_cap * 8;
    A.dir_gid = reinterpret_cast<int *>(p); p += A.dir_cap * 4;
    A.states = reinterpret_cast<uint64_t *>(p);
    A.gid_keys = A.states + static_cast<size_t>(A.partial_rows) * A.words;
    A.cap = A.partial_rows;
    // every partial row starts as (and is reset to, by k_merge_partials) the identity, so rows of groups a CTA
    // never met contribute nothing to the fold
    QS_CUDA(launch_agg_init(A, s->max_ctas, s->ctl, d->stream));
    count_launch();
  } else {
    QS_CUDA(dev_malloc(&A.n_groups, 256));
    QS_CUDA(cudaMemsetAsync(A.n_groups, 0, 256, d->stream));
    QS_CUDA(dev_malloc(&s->d_done, 256));
    QS_CUDA(cudaMemsetAsync(s->d_done, 0, 256, d->stream));
    QS_CUDA(dev_malloc(&s->d_idx_count, 256));
    if (strategy == QS_AGG_COLLISION_FREE) {
      A.cap = static_cast<uint64_t>(spec->collision_free_max_key) + 1;
      QS_CUDA(dev_malloc(&A.states, A.cap * A.words * 8));
      QS_CUDA(launch_fill_identity(A.states, A.cap, A, d->stream));
      count_launch();
    }
    // SEPARATE_CHAINING: the table is allocated by the first work order (maybe_grow), sized from the rows that work
    // order really holds and the estimate together -- an optimizer estimate 10x too high must not make every query
    // initialise, probe and finally scan a table of gigabytes (Q3 at SF100: 64 M slots for 1.5 M groups)
  }
  *out = s.release();
  return QSGPU_OK;
}

// Grows a SEPARATE_CHAINING table so that `extra` more groups certainly fit
// (the reference resizes PackedPayloadHashTable under an exclusive lock,
// storage/PackedPayloadHashTable.hpp:856; here growth happens between work
// orders, never inside a kernel).
static int maybe_grow(qsgpu_agg_state *s, Device *d, uint64_t extra_rows) {
  AggDesc &A = s->A;
  uint32_t n = 0;
  if (A.cap != 0) {          // (a table that does not exist yet holds no groups: nothing to ask the device)
    QS_CUDA(cudaMemcpyAsync(&n, A.n_groups, 4, cudaMemcpyDeviceToHost, d->stream));
    QS_CUDA(cudaStreamSynchronize(d->stream));
  }
  // worst case every row opens a group, but never plan for more than 8x the optimizer's estimate at once
  // (a kernel that still overflows raises QSGPU_ERR_CAPACITY; nothing is silently dropped)
  uint64_t worst = n + std::min<uint64_t>(extra_rows, std::max<uint64_t>(8 * s->estimated, 1u << 20));
  if (A.cap != 0 && worst * 3 / 2 <= A.cap) return QSGPU_OK;
  uint64_t cap = std::max<uint64_t>(A.cap, 1024);
  while (cap < worst * 2) cap <<= 1;
  AggDesc B = A;
  B.cap = cap;
  // cap + 1 rows: the last one is the reserved row of the all-ones key (CAS-claimed keys, qs_kernels.cuh K7)
  B.tags = nullptr;
  if (A.key_words > 2) {
    QS_CUDA(dev_malloc(&B.tags, (cap + 1) * 4));
    QS_CUDA(cudaMemsetAsync(B.tags, 0, (cap + 1) * 4, d->stream));
  }
  QS_CUDA(dev_malloc(&B.keys, (cap + 1) * A.key_words * 8));
  QS_CUDA(cudaMemsetAsync(B.keys, 0xff, (cap + 1) * A.key_words * 8, d->stream));     // "empty" = all ones
  QS_CUDA(dev_malloc(&B.states, (cap + 1) * A.words * 8));
  QS_CUDA(launch_fill_identity(B.states, cap + 1, B, d->stream));
  count_launch();
  if (A.cap != 0) {
    QS_CUDA(cudaMemsetAsync(A.n_groups, 0, 8, d->stream));        // group counter + "reserved row in use" flag
    QS_CUDA(launch_rehash(A, B, d->stream));
    count_launch();
    QS_CUDA(cudaStreamSynchronize(d->stream));
    dev_free(A.tags); dev_free(A.keys); dev_free(A.states);
  }
  A = B;
  return QSGPU_OK;
}

int qsgpu_agg_run(qsgpu_agg_state_t state, qsgpu_relation_t input, uint64_t row_begin, uint64_t row_end,
                  uint32_t n_lip_probe, const qs_lip_ref *lip_probe) {
  std::lock_guard<std::mutex> state_lock(state->mu);
  Device *d = device(state->dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  if (input->dev != state->dev) { set_error(QSGPU_ERR_INVALID, "input relation on another device"); return QSGPU_ERR_INVALID; }
  qs_scan scan{};
  scan.input = input;
  scan.exprs = &state->exprs;
  scan.predicate_root = state->predicate_root;
  scan.n_lip_probe = n_lip_probe;
  scan.lip_probe = lip_probe;
  Lowering L(&state->exprs, input);
  // Rows whose group-by key is NULL belong to no group: PackedPayloadHashTable::upsertValueAccessorCompositeKey skips
  // them (storage/PackedPayloadHashTable.hpp:861-866, GetCompositeKeyFromValueAccessor<..., check_for_null_keys = true>;
  // the engine prints no NULL group, tests/golden/ref_null_results.json "group_by_nullable")
  uint64_t group_key_mask = 0;
  for (uint32_t k = 0; k < state->A.n_key_cols; ++k) if (state->key_attr_ids[k] < 64) group_key_mask |= 1ull << state->key_attr_ids[k];
  int st = lower_scan_predicate(L, &scan, group_key_mask);
  if (st) return st;
  AggDesc A = state->A;
  for (uint32_t k = 0; k < A.n_key_cols; ++k) {
    const uint32_t attr = state->key_attr_ids[k];
    if (attr >= input->attrs.size() || input->attrs[attr].width != A.key_width[k]) { set_error(QSGPU_ERR_INVALID, "group-by attribute does not match the input relation"); return QSGPU_ERR_INVALID; }
    A.key_col[k] = static_cast<uint16_t>(L.stage_attr(attr));
  }
  std::vector<bool> nn_done(A.n_agg + 1, false);
  for (size_t i = 0; i < state->aggregates.size(); ++i) {
    const int w = state->value_word[i];
    const int nw = state->nn_word[i];
    const int32_t root = state->aggregates[i].argument_root;
    const uint64_t nb = root >= 0 ? L.null_bits(root) : 0;
    if (nb && !nw) { set_error(QSGPU_ERR_INVALID, "aggregate over a NULL-able attribute not declared in qs_agg_spec.nullable_arguments"); return QSGPU_ERR_INVALID; }
    if (w != 0 && w != nw) {
      const uint8_t t = L.lower_scalar(root);
      const uint8_t kind = A.kind[w - 1];
      const uint8_t to = (kind == AK_SUM_F64 || kind == AK_MIN_F64 || kind == AK_MAX_F64) ? V_F64 : V_I64;
      L.lower_cast_acc(t, to);
      // a NULL argument leaves the state as it is (AggregationHandleSum.hpp:117-127 `if (value.isNull()) return;`):
      // the row contributes the identity of the combine (x + 0 = x bit for bit: the accumulators start at +0.0)
      if (nb) L.lower_null_select(nb, agg_identity(kind));
      Instr in{};
      in.op = OP_EMIT; in.type = to; in.arg = static_cast<uint16_t>(w - 1);
      L.push(in);
    }
    if (nw && !nn_done[nw]) {
      nn_done[nw] = true;
      Instr ld{};
      ld.op = OP_LOAD; ld.type = V_I64; ld.leaf = LEAF_LIT; ld.ltype = V_I64;
      ld.arg = static_cast<uint16_t>(L.add_lit(1));
      L.push(ld);
      if (nb) L.lower_null_select(nb, 0);
      Instr in{};
      in.op = OP_EMIT; in.type = V_I64; in.arg = static_cast<uint16_t>(nw - 1);
      L.push(in);
    }
  }
  L.finish();
  if (!L.ok()) { set_error(L.status, L.err); return L.status; }
  ScanDesc S;
  fill_scan(input, row_begin, row_end, L, &S);
  st = fill_lips(n_lip_probe, lip_probe, input, state->dev, &S);
  if (st) return st;
  ScanPlan plan;
  if (state->strategy == QS_AGG_SINGLE_STATE || state->strategy == QS_AGG_COMPACT_KEY) {
    const int hot = agg_hot_groups(A);
    // Grouped states over narrow tiles (scans of dictionary codes) keep the hot groups' value accumulators as
    // per-thread shared-memory slots instead of registers (scan_agg_body, Q::priv): taken when two CTAs per SM
    // still fit with a double-buffered ring.  Wide native tiles (Q1 on native columns: 43 KB per stage) do not.
    bool priv = false;
    if (A.n_key_cols > 0 && A.n_agg > 0 && A.n_agg <= 6) {   // a tile's values wait in registers: 8 per aggregate
      ScanDesc S2 = S;
      ScanPlan p2;
      priv = plan_scan(d, &S2, agg_smem_extra(hot, static_cast<int>(A.n_agg), true, A.words, true), &p2, 2) == QSGPU_OK && p2.ctas >= 2;
    }
    const size_t agg_extra = agg_smem_extra(hot, static_cast<int>(A.n_agg), A.n_key_cols > 0, A.words, priv);
    st = plan_scan(d, &S, agg_extra, &plan, priv ? 2 : 4);
    if (st) return st;
    // wide tiles (native Q6: 32 KB) measured best with 2 resident CTAs per SM (0.269 vs 0.279 ms with 3); narrow
    // code tiles take all 4 (Q6 on codes: 0.143 -> 0.128 ms)
    const uint32_t grid_cap = std::min<uint32_t>(state->max_ctas, static_cast<uint32_t>(d->sm_count) * (S.stage_bytes <= 16384 ? 4u : 2u));
    if (static_cast<uint32_t>(plan.grid) > grid_cap) plan.grid = static_cast<int>(grid_cap);
    JitKernel *kern = nullptr;
    st = query_kernel(JF_AGG, S, L.P, plan, &A, nullptr, nullptr, hot, &kern, priv ? 1 : 0);
    if (st) return st;
    // The ring was sized for the CTAs the shared memory allows; when the compiled kernel's registers allow fewer
    // (Q1: 2), give each resident CTA the deeper ring its share of the SM's shared memory affords.
    if (const int occ = jit_occupancy(kern, plan.smem); occ > 0 && occ < plan.ctas) {
      st = plan_scan(d, &S, agg_extra, &plan, occ);
      if (st) return st;
      if (static_cast<uint32_t>(plan.grid) > grid_cap) plan.grid = static_cast<int>(grid_cap);
      st = query_kernel(JF_AGG, S, L.P, plan, &A, nullptr, nullptr, hot, &kern, priv ? 1 : 0);
      if (st) return st;
    }
    {
      KernelTimer timer(d, QS_K_SCAN_AGG);      // the scan kernel alone (roofline numerator)
      QS_CUDA(launch_query_kernel(d, kern, S, L.P, plan, &A, nullptr, nullptr));
    }
    QS_CUDA(launch_merge_partials(A, static_cast<uint32_t>(plan.grid), d->stream));
    count_launch(2);
  } else {
    // everything that does not depend on the table's size first (ring plan, kernel look-up): the host wait below
    // drains the stream, and what follows it is on the query's critical path
    st = plan_scan(d, &S, 0, &plan);
    if (st) return st;
    JitKernel *kern = nullptr;
    st = query_kernel(JF_GROUPBY, S, L.P, plan, &A, nullptr, nullptr, 1, &kern);
    if (st) return st;
    if (state->strategy == QS_AGG_SEPARATE_CHAINING && !t_sc) {
      // Size the table for the rows this work order can really add.  The input's row count may still be
      // device-only (a temporary relation just produced); maybe_grow synchronises anyway, so read it
      // first instead of planning for the relation's capacity.
      st = sync_rows(input);
      if (st) return st;
      uint64_t rows = std::min<uint64_t>(row_end == UINT64_MAX ? input->host_rows : row_end, input->host_rows);
      rows = rows > row_begin ? rows - row_begin : 0;
      st = maybe_grow(state, d, rows);
      if (st) return st;
      state->rows_fed += rows;
      A.tags = state->A.tags; A.keys = state->A.keys; A.states = state->A.states; A.cap = state->A.cap;
    }
    KernelTimer timer(d, QS_K_GROUPBY);
    QS_CUDA(launch_query_kernel(d, kern, S, L.P, plan, &A, nullptr, nullptr));
    count_launch();
  }
  return QSGPU_OK;
}

// Enqueue-only form for finalization: the slot index list and its length (d_idx_count) stay on the device; *upper_out
// bounds the number of groups from what the host knows (slots of the table; for a hash table also the rows it was fed).
static int collect_groups_async(qsgpu_agg_state *s, Device *d, uint64_t *upper_out) {
  AggDesc &A = s->A;
  QS_CUDA(cudaMemsetAsync(s->d_idx_count, 0, 8, d->stream));
  if (A.cap == 0) { *upper_out = 0; return QSGPU_OK; }               // no work order ever ran: no groups
  if (s->idx_cap < A.cap + 1) {
    if (s->d_idx) dev_free(s->d_idx);
    QS_CUDA(dev_malloc(&s->d_idx, (A.cap + 1) * 8 + 64));
    s->idx_cap = A.cap + 1;
  }
  const uint64_t table_rows = A.cap + (s->strategy == QS_AGG_SEPARATE_CHAINING ? 1 : 0);
  QS_CUDA(launch_collect_slots(A.states, A.words, table_rows, s->d_idx, s->d_idx_count,
                               s->existence ? s->existence->d.words : nullptr, d->stream));
  count_launch();
  *upper_out = s->strategy == QS_AGG_SEPARATE_CHAINING ? std::min<uint64_t>(table_rows, s->rows_fed) : table_rows;
  return QSGPU_OK;
}

static int collect_groups(qsgpu_agg_state *s, Device *d, uint64_t *n_out) {
  AggDesc &A = s->A;
  if (A.cap == 0) { *n_out = 0; return check_device_error(d); }      // no work order ever ran: no groups
  if (s->idx_cap < A.cap + 1) {
    if (s->d_idx) dev_free(s->d_idx);
    QS_CUDA(dev_malloc(&s->d_idx, (A.cap + 1) * 8 + 64));
    s->idx_cap = A.cap + 1;
  }
  QS_CUDA(cudaMemsetAsync(s->d_idx_count, 0, 8, d->stream));
  const uint64_t table_rows = A.cap + (s->strategy == QS_AGG_SEPARATE_CHAINING ? 1 : 0);
  QS_CUDA(launch_collect_slots(A.states, A.words, table_rows, s->d_idx, s->d_idx_count,
                               s->existence ? s->existence->d.words : nullptr, d->stream));
  count_launch();
  unsigned long long n = 0;
  QS_CUDA(cudaMemcpyAsync(&n, s->d_idx_count, 8, cudaMemcpyDeviceToHost, d->stream));
  QS_CUDA(cudaStreamSynchronize(d->stream));
  *n_out = n;
  return check_device_error(d);
}

static int agg_num_groups_locked(qsgpu_agg_state_t state, uint64_t *n_groups) {
  Device *d = device(state->dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  if (state->strategy == QS_AGG_COLLISION_FREE) return collect_groups(state, d, n_groups);
  uint32_t n = 0;
  QS_CUDA(cudaMemcpyAsync(&n, state->A.n_groups, 4, cudaMemcpyDeviceToHost, d->stream));
  QS_CUDA(cudaStreamSynchronize(d->stream));
  int st = check_device_error(d);
  if (st) return st;
  *n_groups = state->strategy == QS_AGG_SINGLE_STATE ? 1 : n;
  return QSGPU_OK;
}

int qsgpu_agg_num_groups(qsgpu_agg_state_t state, uint64_t *n_groups) {
  std::lock_guard<std::mutex> state_lock(state->mu);
  return agg_num_groups_locked(state, n_groups);
}

int qsgpu_agg_partial(qsgpu_agg_state_t state, void **d_states, void **d_keys, uint64_t *n_groups,
                      uint32_t *words_per_group, uint32_t *key_words) {
  std::lock_guard<std::mutex> state_lock(state->mu);
  Device *d = device(state->dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  AggDesc &A = state->A;
  *words_per_group = A.words;
  *key_words = state->strategy == QS_AGG_COLLISION_FREE ? 1 : A.key_words;
  if (state->strategy == QS_AGG_SINGLE_STATE || state->strategy == QS_AGG_COMPACT_KEY) {
    int st = agg_num_groups_locked(state, n_groups);
    if (st) return st;
    *d_states = A.states;
    *d_keys = A.gid_keys;
    return QSGPU_OK;
  }
  uint64_t n = 0;
  int st = collect_groups(state, d, &n);
  if (st) return st;
  if (state->exp_cap < n || !state->d_exp_states) {
    if (state->d_exp_states) { dev_free(state->d_exp_states); dev_free(state->d_exp_keys); }
    const uint64_t cap = std::max<uint64_t>(n, 1024);
    QS_CUDA(dev_malloc(&state->d_exp_states, cap * A.words * 8));
    QS_CUDA(dev_malloc(&state->d_exp_keys, cap * (*key_words) * 8));
    state->exp_cap = cap;
  }
  QS_CUDA(launch_gather_rows(A.states, A.keys, A.words, A.key_words, state->d_idx, n, state->d_exp_states,
                             state->d_exp_keys, state->strategy == QS_AGG_COLLISION_FREE, d->stream));
  count_launch();
  QS_CUDA(cudaStreamSynchronize(d->stream));
  *d_states = state->d_exp_states;
  *d_keys = state->d_exp_keys;
  *n_groups = n;
  return QSGPU_OK;
}

int qsgpu_agg_partial_layout(qsgpu_agg_state_t state, void **d_states, void **d_keys, uint64_t *rows,
                             uint32_t *words_per_group, uint32_t *key_words) {
  if (state->strategy != QS_AGG_SINGLE_STATE && state->strategy != QS_AGG_COMPACT_KEY) {
    set_error(QSGPU_ERR_UNSUPPORTED, "only fixed-size states (single state, compact key) have a static partial layout");
    return QSGPU_ERR_UNSUPPORTED;
  }
  const AggDesc &A = state->A;
  *d_states = A.states;
  *d_keys = A.gid_keys;
  *rows = A.partial_rows;
  *words_per_group = A.words;
  *key_words = A.key_words;
  return QSGPU_OK;
}

int qsgpu_agg_merge_partial(qsgpu_agg_state_t state, const void *d_states, const void *d_keys, uint64_t n_groups) {
  std::lock_guard<std::mutex> state_lock(state->mu);
  Device *d = device(state->dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  const AggDesc &A = state->A;
  if (state->strategy == QS_AGG_SINGLE_STATE || state->strategy == QS_AGG_COMPACT_KEY) {
    QS_CUDA(launch_merge_foreign_compact(A, static_cast<const uint64_t *>(d_states), static_cast<const uint64_t *>(d_keys),
                                         static_cast<uint32_t>(n_groups), d->stream));
  } else {
    if (state->strategy == QS_AGG_SEPARATE_CHAINING) {
      int st = maybe_grow(state, d, n_groups);
      if (st) return st;
      state->rows_fed += n_groups;
    }
    QS_CUDA(launch_merge_foreign_table(state->A, static_cast<const uint64_t *>(d_states), static_cast<const uint64_t *>(d_keys),
                                       n_groups, d->stream));
  }
  count_launch();
  return QSGPU_OK;
}

int qsgpu_agg_finalize(qsgpu_agg_state_t state, qsgpu_relation_t *out, uint64_t *null_mask) {
  std::lock_guard<std::mutex> state_lock(state->mu);
  Device *d = device(state->dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  AggDesc &A = state->A;
  const bool dense = state->strategy == QS_AGG_SINGLE_STATE || state->strategy == QS_AGG_COMPACT_KEY;
  // Fixed-size states: the output relation is sized for the state's row limit and the kernel reads the live group
  // count on the device, so the call only ENQUEUES (no host wait; a capacity error of the scan surfaces at the
  // next read of the result).  Tables: the groups are collected first and their number sizes the output.
  // Tables: the occupied slots are collected on the device and their number stays there too (the output is sized for
  // an upper bound the host knows: the slots, and for a hash table the rows it was fed) -- no host wait here either.
  uint64_t n = dense ? A.partial_rows : 0;
  int st = dense ? QSGPU_OK : collect_groups_async(state, d, &n);
  if (st) return st;
  // output schema: group-by attributes, then one column per aggregate
  std::vector<qs_attr> attrs = state->key_attrs;
  FinalizeDesc F{};
  F.n_key_cols = A.n_key_cols;
  F.key_words = A.key_words;
  F.keys_are_slots = state->strategy == QS_AGG_COLLISION_FREE;
  for (uint32_t k = 0; k < A.n_key_cols; ++k) { F.key_width[k] = A.key_width[k]; F.key_off[k] = A.key_off[k]; }
  F.n_out = static_cast<uint32_t>(state->aggregates.size());
  bool any_nn = false;
  for (size_t j = 0; j < state->aggregates.size(); ++j) {
    const qs_aggregate &a = state->aggregates[j];
    const int w = state->value_word[j];
    F.function[j] = static_cast<uint8_t>(a.function);
    F.word[j] = static_cast<uint8_t>(w);
    F.nn_word[j] = static_cast<uint8_t>(state->nn_word[j]);
    any_nn = any_nn || state->nn_word[j] != 0;
    uint8_t kind = w ? A.kind[w - 1] : AK_SUM_I64;
    F.word_is_f64[j] = (kind == AK_SUM_F64 || kind == AK_MIN_F64 || kind == AK_MAX_F64) ? 1 : 0;
    qs_attr oa{};
    switch (a.function) {
      case QS_AGG_COUNT: oa.type = QS_LONG; F.out_vtype[j] = V_I64; break;
      case QS_AGG_AVG: oa.type = QS_DOUBLE; F.out_vtype[j] = V_F64; break;
      case QS_AGG_SUM: oa.type = F.word_is_f64[j] ? QS_DOUBLE : QS_LONG; F.out_vtype[j] = F.word_is_f64[j] ? V_F64 : V_I64; break;
      default: {   // MIN / MAX keep the argument type (AggregationHandleMin.hpp:101-207)
        const uint8_t v = state->arg_vtype[j];
        F.out_vtype[j] = v;
        oa.type = v == V_I32 ? QS_INT : v == V_I64 ? QS_LONG : v == V_F32 ? QS_FLOAT : QS_DOUBLE;
      }
    }
    oa.width = attr_width(oa.type, 0);
    attrs.push_back(oa);
  }
  qsgpu_relation *rel = nullptr;
  st = qsgpu_relation_create(state->dev, static_cast<uint32_t>(attrs.size()), attrs.data(), std::max<uint64_t>(n, 1), &rel);
  if (st) return st;
  for (uint32_t k = 0; k < A.n_key_cols; ++k) F.key_out[k] = rel->cols[k];
  for (uint32_t j = 0; j < F.n_out; ++j) F.out[j] = rel->cols[A.n_key_cols + j];
  const uint64_t *keys = dense ? A.gid_keys : A.keys;
  F.rows_out = rel->d_rows;
  // live group count on the device: the state's dense id counter (COMPACT_KEY), or the length of the collected slot
  // list (tables; a 64-bit counter whose low word is read)
  F.d_n_groups = state->strategy == QS_AGG_COMPACT_KEY ? A.n_groups
               : !dense ? reinterpret_cast<const uint32_t *>(state->d_idx_count) : nullptr;
  if (state->strategy == QS_AGG_SINGLE_STATE || any_nn) {
    // aggregates over zero rows -- or, for a NULL-able argument, over zero non-NULL values -- are SQL NULL:
    // recorded in the output relation's per-row NULL mask
    if (const int ns = ensure_null_mask(rel, d)) { qsgpu_relation_destroy(rel); return ns; }
    F.null_out = rel->d_nulls;
    // Which aggregates can be NULL is the reference's rule, odd as it is: without GROUP BY every one except COUNT
    // (AggregationHandleSum.cpp:134-143 and friends).  With GROUP BY the hash-table payload of SUM / AVG is the bare
    // running sum (AggregationHandleSum.hpp:176-178, AggregationHandleAvg.hpp:180-189), so a group whose arguments
    // were all NULL finalizes to SUM = 0 and AVG = NaN; only MIN / MAX, whose payload is a NULL-able TypedValue
    // (AggregationHandleMin.hpp:152-155), come out NULL.
    for (size_t j = 0; j < state->aggregates.size(); ++j) {
      const uint32_t fn = state->aggregates[j].function;
      const bool can_be_null = state->strategy == QS_AGG_SINGLE_STATE ? fn != QS_AGG_COUNT
                                                                     : (state->nn_word[j] && (fn == QS_AGG_MIN || fn == QS_AGG_MAX));
      if (can_be_null) F.null_bits |= 1ull << (A.n_key_cols + j);
    }
    rel->nullable_mask = F.null_bits;
  }
  KernelTimer timer(d, QS_K_GROUPBY);
  cudaError_t e = launch_finalize(A.states, keys, A.words, dense ? nullptr : state->d_idx, n, F, d->stream);
  count_launch();
  if (e != cudaSuccess) { qsgpu_relation_destroy(rel); return cuda_fail(e, "finalize"); }
  if (null_mask) {                    // asked for on the host: this (and only this) waits for the queued work
    *null_mask = 0;
    if (state->strategy == QS_AGG_SINGLE_STATE) {
      uint64_t row[kMaxAgg + 1] = {0};
      QS_CUDA(cudaMemcpyAsync(row, A.states, 8 * A.words, cudaMemcpyDeviceToHost, d->stream));
      QS_CUDA(cudaStreamSynchronize(d->stream));
      for (size_t j = 0; j < state->aggregates.size(); ++j)
        if (state->aggregates[j].function != QS_AGG_COUNT && row[state->nn_word[j]] == 0) *null_mask |= 1ull << j;
    }
  }
  if (state->strategy != QS_AGG_SINGLE_STATE) {
    rel->dirty = true;                // the kernel stored the live group count in the relation's device counter
  } else {
    rel->host_rows = n;
    rel->dirty = false;
  }
  *out = rel;
  return QSGPU_OK;
}

int qsgpu_agg_existence_map(qsgpu_agg_state_t state, qsgpu_lip_t *out) {
  std::lock_guard<std::mutex> state_lock(state->mu);
  if (state->strategy != QS_AGG_COLLISION_FREE) { set_error(QSGPU_ERR_INVALID, "only a COLLISION_FREE state has an existence map"); return QSGPU_ERR_INVALID; }
  if (!state->existence) {
    const uint32_t key_type = state->key_attrs[0].type;
    int st = qsgpu_lip_create(state->dev, QS_LIP_BITVECTOR_EXACT, key_type, 0, static_cast<int64_t>(state->A.cap) - 1, 0, 0, &state->existence);
    if (st) return st;
  }
  *out = state->existence;
  return QSGPU_OK;
}

int qsgpu_agg_destroy(qsgpu_agg_state_t s) {
  if (!s) return QSGPU_OK;
  if (s->existence) qsgpu_lip_destroy(s->existence);
  device(s->dev);
  AggDesc &A = s->A;
  if (s->ctl) {             // fixed-size strategies: counters, key directory, states and keys are one block
    dev_free(A.partials); dev_free(s->ctl);
  } else {
    dev_free(A.n_groups); dev_free(A.tags); dev_free(A.keys); dev_free(A.states);
    dev_free(s->d_done); dev_free(s->d_idx_count);
  }
  dev_free(s->d_idx); dev_free(s->d_exp_states); dev_free(s->d_exp_keys);
  delete s;
  return QSGPU_OK;
}

/* --------------------------------------------------------------- hash join */
int qsgpu_join_create(int dev, uint32_t key_type, uint64_t estimated_num_entries, qsgpu_join_table_t *out) {
  Device *d = device(dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  if (key_type != QS_INT && key_type != QS_LONG) { set_error(QSGPU_ERR_UNSUPPORTED, "join keys are single INT/LONG attributes"); return QSGPU_ERR_UNSUPPORTED; }
  std::unique_ptr<qsgpu_join_table> t(new qsgpu_join_table);
  t->dev = dev;
  t->key_type = key_type;
  uint64_t cap = 1024;
  while (cap < estimated_num_entries * 2) cap <<= 1;
  t->J.cap = cap;
  t->alloc_cap = cap;
  t->J.error_flag = d->d_error;
  t->J.key_ltype = key_type == QS_INT ? V_I32 : V_I64;
  QS_CUDA(dev_malloc(&t->J.slots, cap * sizeof(JoinSlot)));
  QS_CUDA(dev_malloc(&t->J.n_entries, 256));
  QS_CUDA(cudaMemsetAsync(t->J.n_entries, 0, 256, d->stream));
  // cleared (and its mask fixed) by the first build work order, see qsgpu_join_table
  *out = t.release();
  return QSGPU_OK;
}

// Open addressing: make room for `rows` more entries before a build work order is queued (table->mu held).
static int join_reserve(qsgpu_join_table *t, Device *d, uint64_t rows) {
  if (t->J.dense) return QSGPU_OK;
  uint64_t need = 1024;
  while (need < (t->upper_entries + rows) * 2) need <<= 1;
  if (!t->cleared) {
    if (t->cap_frozen) need = std::max(need, t->J.cap);
    if (need > t->alloc_cap) {
      dev_free(t->J.slots);
      t->J.slots = nullptr;
      QS_CUDA(dev_malloc(&t->J.slots, need * sizeof(JoinSlot)));
      t->alloc_cap = need;
    }
    t->J.cap = need;
    QS_CUDA(launch_join_clear(t->J, d->stream));
    count_launch();
    t->cleared = true;
  } else if (need > t->J.cap) {
    JoinDesc B = t->J;
    B.cap = need;
    B.slots = nullptr;
    QS_CUDA(dev_malloc(&B.slots, need * sizeof(JoinSlot)));
    QS_CUDA(launch_join_clear(B, d->stream));
    QS_CUDA(launch_join_rehash(t->J.slots, t->J.cap, B.slots, B.cap, d->stream));
    count_launch(2);
    dev_free(t->J.slots);            // stream-ordered: freed after the re-hash ran
    t->J = B;
    t->alloc_cap = need;
  }
  t->upper_entries += rows;
  return QSGPU_OK;
}

int qsgpu_join_create_dense(int dev, uint32_t key_type, int64_t min_key, int64_t max_key, qsgpu_join_table_t *out) {
  Device *d = device(dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  if (key_type != QS_INT && key_type != QS_LONG) { set_error(QSGPU_ERR_UNSUPPORTED, "join keys are single INT/LONG attributes"); return QSGPU_ERR_UNSUPPORTED; }
  if (max_key < min_key || static_cast<uint64_t>(max_key - min_key) >= (1ull << 34)) { set_error(QSGPU_ERR_INVALID, "dense join table needs min <= max and a key range below 2^34"); return QSGPU_ERR_INVALID; }
  std::unique_ptr<qsgpu_join_table> t(new qsgpu_join_table);
  t->dev = dev;
  t->key_type = key_type;
  t->J.dense = 1;
  t->J.min_key = min_key;
  t->J.cap = static_cast<uint64_t>(max_key - min_key) + 1;
  t->J.error_flag = d->d_error;
  t->J.key_ltype = key_type == QS_INT ? V_I32 : V_I64;
  QS_CUDA(dev_malloc(&t->J.heads, t->J.cap * 8));
  QS_CUDA(dev_malloc(&t->J.n_entries, 256));
  QS_CUDA(cudaMemsetAsync(t->J.n_entries, 0, 256, d->stream));
  QS_CUDA(launch_join_clear(t->J, d->stream));
  count_launch();
  *out = t.release();
  return QSGPU_OK;
}

// One key attribute of the table's key type, or two INT attributes packed into the LONG key of the table.
static int check_join_keys(const qsgpu_join_table *table, const qsgpu_relation *rel, uint32_t n_keys, const uint32_t *key_attrs) {
  if (n_keys == 0 || n_keys > 2 || !key_attrs) { set_error(QSGPU_ERR_UNSUPPORTED, "join keys: one INT/LONG attribute or two INT attributes"); return QSGPU_ERR_UNSUPPORTED; }
  for (uint32_t i = 0; i < n_keys; ++i)
    if (key_attrs[i] >= rel->attrs.size()) { set_error(QSGPU_ERR_INVALID, "join key attribute out of range"); return QSGPU_ERR_INVALID; }
  if (n_keys == 1) {
    const uint16_t t = rel->attrs[key_attrs[0]].type;
    if (t != QS_INT && t != QS_LONG) { set_error(QSGPU_ERR_UNSUPPORTED, "join key must be INT/LONG"); return QSGPU_ERR_UNSUPPORTED; }
    return QSGPU_OK;
  }
  if (rel->attrs[key_attrs[0]].type != QS_INT || rel->attrs[key_attrs[1]].type != QS_INT || table->key_type != QS_LONG || table->J.dense) {
    set_error(QSGPU_ERR_UNSUPPORTED, "a composite join key is two INT attributes on an open-addressing table created with QS_LONG keys");
    return QSGPU_ERR_UNSUPPORTED;
  }
  return QSGPU_OK;
}

int qsgpu_join_build(qsgpu_join_table_t table, const qs_scan *scan, uint32_t key_attr, uint32_t n_lip_build,
                     const qs_lip_ref *lip_build) {
  return qsgpu_join_build_composite(table, scan, 1, &key_attr, n_lip_build, lip_build);
}

int qsgpu_join_build_composite(qsgpu_join_table_t table, const qs_scan *scan, uint32_t n_keys, const uint32_t *key_attrs,
                               uint32_t n_lip_build, const qs_lip_ref *lip_build) {
  qsgpu_relation *rel = scan->input;
  Device *d = device(table->dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  if (rel->dev != table->dev) { set_error(QSGPU_ERR_INVALID, "build relation on another device"); return QSGPU_ERR_INVALID; }
  {
    const int ks = check_join_keys(table, rel, n_keys, key_attrs);
    if (ks) return ks;
  }
  const uint32_t key_attr = key_attrs[0];
  if (n_keys == 1 && rel->attrs[key_attr].type != table->key_type) { set_error(QSGPU_ERR_INVALID, "build key attribute does not match the table"); return QSGPU_ERR_INVALID; }
  // build work orders of one operator run concurrently: sizing (which may re-hash into a new slot array) and the
  // launch that uses the array are one unit in stream order
  std::lock_guard<std::mutex> lk(table->mu);
  {
    if (table->build_rel && table->build_rel != rel) { set_error(QSGPU_ERR_UNSUPPORTED, "one build relation per join table"); return QSGPU_ERR_UNSUPPORTED; }
    table->build_rel = rel;
    if (table->J.dense && !table->J.next && !t_sc) QS_CUDA(dev_malloc(&table->J.next, std::max<uint64_t>(rel->capacity, 1) * 8));
  }
  Lowering L(scan->exprs, rel);
  uint64_t key_mask = 0;
  for (uint32_t i = 0; i < n_keys; ++i) if (key_attrs[i] < 64) key_mask |= 1ull << key_attrs[i];
  if (n_lip_build == 1 && lip_build[0].attr < 64) {
    // the filter is built from the rows that enter the table: fine when its attribute is the key (the usual
    // deployment), otherwise its NULL rows would have to be skipped for the filter alone
    if (((rel->nullable_mask & ~key_mask) >> lip_build[0].attr) & 1ull) { set_error(QSGPU_ERR_UNSUPPORTED, "LIP filter built from a NULL-able non-key attribute by a join build"); return QSGPU_ERR_UNSUPPORTED; }
  }
  int st = lower_scan_predicate(L, scan, key_mask);
  if (st) return st;
  SinkDesc K{};
  K.error_flag = d->d_error;
  st = fill_lip_build(n_lip_build, lip_build, L, rel, &K);
  if (st) return st;
  JoinDesc J = table->J;
  J.null_col = 0xffff;               // (probe side only; NULL build keys are filtered by the scan predicate above)
  J.key_col = static_cast<uint16_t>(L.stage_attr(key_attr));
  J.key_ltype = vtype_of(rel->attrs[key_attr].type);
  if (n_keys == 2) { J.key2_present = 1; J.key2_col = static_cast<uint16_t>(L.stage_attr(key_attrs[1])); }
  L.finish();
  ScanDesc S;
  fill_scan(rel, scan->row_begin, scan->row_end, L, &S);
  st = fill_lips(scan->n_lip_probe, scan->lip_probe, rel, rel->dev, &S);
  if (st) return st;
  ScanPlan plan;
  st = plan_scan(d, &S, 0, &plan, 8);      // a key tile is a few KB and the kernel ~25 registers: all the warps an SM holds
  if (st) return st;
  JitKernel *kern = nullptr;
  st = query_kernel(JF_JOIN_BUILD, S, L.P, plan, nullptr, &K, &J, 1, &kern);
  if (st) return st;
  if (!table->J.dense && !t_sc) {
    // Sizing comes last: the rows this work order can insert may only be known on the device (a temporary relation
    // just produced), reading them drains the stream, and whatever the host does after that -- it used to be the
    // lowering and the kernel look-up above -- is on the query's critical path.
    int rs = sync_rows(rel);
    if (rs) return rs;
    const uint64_t hi = std::min<uint64_t>(scan->row_end, rel->host_rows);
    rs = join_reserve(table, d, hi > scan->row_begin ? hi - scan->row_begin : 0);
    if (rs) return rs;
    J.slots = table->J.slots;          // the reservation may have chosen the mask / moved the slots
    J.cap = table->J.cap;
  }
  KernelTimer timer(d, QS_K_JOIN_BUILD);
  QS_CUDA(launch_query_kernel(d, kern, S, L.P, plan, nullptr, &K, &J));
  count_launch();
  return QSGPU_OK;
}

int qsgpu_join_num_entries(qsgpu_join_table_t table, uint64_t *n) {
  Device *d = device(table->dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  unsigned long long v = 0;
  QS_CUDA(cudaMemcpyAsync(&v, table->J.n_entries, 8, cudaMemcpyDeviceToHost, d->stream));
  QS_CUDA(cudaStreamSynchronize(d->stream));
  *n = v;
  return check_device_error(d);
}

int qsgpu_join_probe(qsgpu_join_table_t table, const qs_scan *probe, uint32_t probe_key_attr, uint32_t join_type,
                     int32_t residual_root, uint32_t n_project, const int32_t *project_roots,
                     qsgpu_relation_t output) {
  return qsgpu_join_probe_composite(table, probe, 1, &probe_key_attr, join_type, residual_root, n_project, project_roots, output);
}

int qsgpu_join_probe_composite(qsgpu_join_table_t table, const qs_scan *probe, uint32_t n_keys, const uint32_t *probe_key_attrs,
                               uint32_t join_type, int32_t residual_root, uint32_t n_project,
                               const int32_t *project_roots, qsgpu_relation_t output) {
  qsgpu_relation *rel = probe->input;
  Device *d = device(table->dev);
  if (!d) return QSGPU_ERR_NO_DEVICE;
  {
    const int ks = check_join_keys(table, rel, n_keys, probe_key_attrs);
    if (ks) return ks;
  }
  const uint32_t probe_key_attr = probe_key_attrs[0];
  if (join_type > QS_JOIN_LEFT_OUTER) { set_error(QSGPU_ERR_INVALID, "unknown join type"); return QSGPU_ERR_INVALID; }
  if (join_type == QS_JOIN_LEFT_OUTER && residual_root >= 0) {
    // DCHECK in the reference too (relational_operators/HashJoinOperator.hpp:139-141)
    set_error(QSGPU_ERR_INVALID, "a LEFT OUTER join takes no residual predicate");
    return QSGPU_ERR_INVALID;
  }
  if (rel->dev != table->dev || output->dev != table->dev || probe_key_attr >= rel->attrs.size()) { set_error(QSGPU_ERR_INVALID, "bad probe arguments"); return QSGPU_ERR_INVALID; }
  const uint8_t klt = vtype_of(rel->attrs[probe_key_attr].type);
  if (klt != V_I32 && klt != V_I64) { set_error(QSGPU_ERR_UNSUPPORTED, "probe key must be INT/LONG"); return QSGPU_ERR_UNSUPPORTED; }
  Lowering L(probe->exprs, rel, table->build_rel);
  uint64_t key_mask = 0;
  for (uint32_t i = 0; i < n_keys; ++i) if (probe_key_attrs[i] < 64) key_mask |= 1ull << probe_key_attrs[i];
  // a probe row with a NULL key matches nothing: inner / semi joins drop it, an anti join emits it and an outer join
  // emits it NULL-padded (HashTable::runOverKeysFromValueAccessor, storage/HashTable.hpp:1999-2003) -- the probe
  // kernel lets such a row pass without searching (JoinDesc::null_col), so the predicate stays what the plan says
  int st = lower_scan_predicate(L, probe);
  if (st) return st;
  if (residual_root >= 0) { L.lower_pred(residual_root); L.mark_mid_end(); }
  SinkDesc K{};
  K.error_flag = d->d_error;
  K.capacity = output->capacity;
  K.counter = output->d_rows;
  const size_t builds_before = L.build_attrs.size();
  st = lower_projection(L, n_project, project_roots, output, &K);
  if (st) return st;
  if (join_type != QS_JOIN_INNER && join_type != QS_JOIN_LEFT_OUTER && L.build_attrs.size() != builds_before) { set_error(QSGPU_ERR_INVALID, "semi/anti joins cannot project build-side attributes"); return QSGPU_ERR_INVALID; }
  if (join_type == QS_JOIN_LEFT_OUTER) {
    // is_selection_on_build: the projected scalars that read the build side are NULL for unmatched probe rows
    std::function<bool(int32_t)> on_build = [&](int32_t i) -> bool {
      const qs_node *n = L.node(i);
      if (!n) return false;
      switch (n->kind) {
        case QS_N_ATTRIBUTE: return n->b == 2;
        case QS_N_UNARY: case QS_N_SHARED: return on_build(n->a);
        case QS_N_BINARY: return on_build(n->a) || on_build(n->b);
        default: return false;
      }
    };
    for (uint32_t j = 0; j < n_project; ++j) if (on_build(project_roots[j])) K.null_bits |= 1ull << j;
    if (!t_sc) { const int ns = ensure_null_mask(output, d); if (ns) return ns; }
    K.null_out = output->d_nulls;
    output->nullable_mask |= K.null_bits;
  }
  if (!t_sc) {     // probing a table no build work order ever touched (empty build side): give it its (empty) slots
    std::lock_guard<std::mutex> lk(table->mu);
    if (!table->J.dense && !table->cleared) { const int rs = join_reserve(table, d, 0); if (rs) return rs; }
  }
  JoinDesc J = table->J;
  J.join_type = static_cast<uint8_t>(join_type);
  J.null_col = 0xffff;
  J.build_nulls = table->build_rel ? table->build_rel->d_nulls : nullptr;
  J.key_null_bits = key_mask & rel->nullable_mask;
  if (J.key_null_bits) J.null_col = static_cast<uint16_t>(L.stage_null_mask());
  J.key_col = static_cast<uint16_t>(L.stage_attr(probe_key_attr));
  J.key_ltype = klt;
  if (n_keys == 2) { J.key2_present = 1; J.key2_col = static_cast<uint16_t>(L.stage_attr(probe_key_attrs[1])); }
  J.n_build_cols = static_cast<uint32_t>(L.build_attrs.size());
  for (uint32_t c = 0; c < J.n_build_cols; ++c) {
    const uint32_t a = L.build_attrs[c];
    if (const uint32_t bcw = table->build_rel->code_width(a)) {
      J.build_cols[c].cw = static_cast<uint8_t>(bcw);
      J.build_cols[c].dict = table->build_rel->coded[a].d_dict;
      J.build_cols[c].dict_entries = table->build_rel->coded[a].n_entries;
    }
    J.build_cols[c].ptr = table->build_rel->cols[a];
    J.build_cols[c].width = table->build_rel->attrs[a].width;
  }
  if (!L.ok()) { set_error(L.status, L.err); return L.status; }
  ScanDesc S;
  fill_scan(rel, probe->row_begin, probe->row_end, L, &S);
  st = fill_lips(probe->n_lip_probe, probe->lip_probe, rel, rel->dev, &S);
  if (st) return st;
  ScanPlan plan;
  st = plan_scan(d, &S, kCompactSmemBytes, &plan);
  if (st) return st;
  JitKernel *kern = nullptr;
  st = query_kernel(JF_JOIN_PROBE, S, L.P, plan, nullptr, &K, &J, 1, &kern);
  if (st) return st;
  KernelTimer timer(d, QS_K_JOIN_PROBE);
  QS_CUDA(launch_query_kernel(d, kern, S, L.P, plan, nullptr, &K, &J));
  count_launch();
  output->dirty = true;
  return QSGPU_OK;
}

int qsgpu_join_destroy(qsgpu_join_table_t t) {
  if (!t) return QSGPU_OK;
  device(t->dev);
  dev_free(t->J.slots);
  dev_free(t->J.heads);
  dev_free(t->J.next);
  dev_free(t->J.n_entries);
  delete t;
  return QSGPU_OK;
}

}  // extern "C"

#include "qs_selfcheck.inc"

extern "C" int qsgpu_jit_stats(uint64_t *compiled, uint64_t *disk_hits, uint64_t *mem_hits) {
  uint64_t a = 0, b = 0, c = 0;
  qs::jit_stats(&a, &b, &c);
  if (compiled) *compiled = a;
  if (disk_hits) *disk_hits = b;
  if (mem_hits) *mem_hits = c;
  return QSGPU_OK;
}
