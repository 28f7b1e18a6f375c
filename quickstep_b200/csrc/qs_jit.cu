// Query compiler of libqsgpu (see qs_jit.h).  Host code only.
#include <dlfcn.h>
#include <nvrtc.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <unordered_map>
#include <memory>
#include <sstream>
#include <vector>

#include "qs_jit.h"

namespace qs {

// Kernel headers embedded at build time (Makefile -> build/jit_headers.inc):
// {include name, text} pairs handed to nvrtcCreateProgram, so the library needs
// neither its own source tree nor a CUDA include directory at run time.
struct EmbeddedHeader { const char *name; const char *text; };
static const EmbeddedHeader kHeaders[] = {
#include "build/jit_headers.inc"
    // minimal stand-ins for the system headers the kernel headers name
    {"cuda_runtime.h", "#pragma once\n"},
    {"stddef.h", "#pragma once\ntypedef unsigned long size_t;\n"},
    {"stdint.h",
     "#pragma once\n"
     "typedef signed char int8_t; typedef unsigned char uint8_t;\n"
     "typedef short int16_t; typedef unsigned short uint16_t;\n"
     "typedef int int32_t; typedef unsigned int uint32_t;\n"
     "typedef long long int64_t; typedef unsigned long long uint64_t;\n"
     "typedef unsigned long uintptr_t;\n"
     "#define UINT64_MAX 0xffffffffffffffffull\n"},
};
static constexpr int kNumHeaders = sizeof(kHeaders) / sizeof(kHeaders[0]);

static const char *kOptions[] = {"--gpu-architecture=sm_100a", "--std=c++17", "--fmad=false", "-lineinfo", "-default-device",
                                 "-DQS_JIT=1"};
static constexpr int kNumOptions = sizeof(kOptions) / sizeof(kOptions[0]);

static std::mutex g_jit_mutex;
static std::map<std::string, std::unique_ptr<JitKernel>> g_kernels;   // keyed by full source
static uint64_t g_compiled = 0, g_disk_hits = 0, g_mem_hits = 0;

// ------------------------------------------------------------------ source
static uint64_t fnv1a(const void *p, size_t n, uint64_t h) {
  const unsigned char *b = static_cast<const unsigned char *>(p);
  for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 0x100000001b3ull; }
  return h;
}

template <class T, class F>
static void table(std::ostringstream &o, const char *type, const char *name, uint32_t n, const T *v, F get) {
  o << "  QSC " << type << " " << name << "(int i) { constexpr " << type << " t[] = {";
  if (n == 0) o << "0";
  for (uint32_t i = 0; i < n; ++i) o << (i ? "," : "") << static_cast<unsigned long long>(get(v[i])) << "u";
  o << "}; return t[i]; }\n";
}

// QSGPU_WARP_COMPACT=1: compile the scan kernels with the warp-wide compaction instead of the CTA-wide, tile-ordered one
// (qs_compact.cuh), for comparison.  Part of the generated source, hence of every cache key.
static bool warp_compact_requested() {
  static const bool on = [] { const char *e = std::getenv("QSGPU_WARP_COMPACT"); return e && e[0] == '1'; }();
  return on;
}

std::string jit_source(const JitSpec &sp, std::string *kernel_name) {
  const ScanDesc &S = *sp.S;
  const Program &P = *sp.P;
  std::ostringstream o;
  if (warp_compact_requested()) o << "#define QS_WARP_COMPACT 1\n";
  o << "#include \"qs_kernels.cuh\"\n#define QSC __host__ __device__ static constexpr\nnamespace qs {\nstruct Q {\n";
  o << "  static constexpr uint32_t n_cols = " << S.n_cols << ", n_stages = " << S.n_stages
    << ", stage_bytes = " << S.stage_bytes << ";\n";
  table(o, "uint32_t", "col_w", S.n_cols, S.cols, [](const ColDesc &c) { return c.width; });
  table(o, "uint32_t", "col_off", S.n_cols, S.cols, [](const ColDesc &c) { return c.smem_off; });
  table(o, "uint32_t", "col_cw", S.n_cols, S.cols, [](const ColDesc &c) { return c.cw; });
  table(o, "uint32_t", "col_coff", S.n_cols, S.cols, [](const ColDesc &c) { return c.code_off; });
  table(o, "uint32_t", "col_expand", S.n_cols, S.cols, [](const ColDesc &c) { return c.expand; });
  table(o, "uint32_t", "col_dsmem", S.n_cols, S.cols, [](const ColDesc &c) { return c.dict_smem; });
  table(o, "uint32_t", "col_doff", S.n_cols, S.cols, [](const ColDesc &c) { return c.dict_soff; });
  o << "  static constexpr int n_pred = " << P.n_pred << ", n_mid = " << P.n_mid << ", n_total = " << P.n_total
    << ";\n";
  o << "  QSC Instr code(int pc) { constexpr Instr t[] = {";
  if (P.n_total == 0) o << "{0,0,0,0,0,0,0}";
  for (uint32_t i = 0; i < P.n_total; ++i) {
    const Instr &in = P.code[i];
    o << (i ? "," : "") << "{" << int(in.op) << "," << int(in.type) << "," << int(in.leaf) << "," << int(in.ltype)
      << "," << int(in.arg) << "," << int(in.flags) << "," << int(in.aux) << "}";
  }
  o << "}; return t[pc]; }\n";
  o << "  static constexpr uint32_t n_lip = " << S.n_lip << ";\n";
  table(o, "uint32_t", "lip_kind", S.n_lip, S.lip, [](const LipDesc &l) { return l.kind; });
  table(o, "uint32_t", "lip_anti", S.n_lip, S.lip, [](const LipDesc &l) { return l.is_anti; });
  table(o, "uint32_t", "lip_soff", S.n_lip, S.lip, [](const LipDesc &l) { return l.smem_off; });
  // aggregation
  {
    AggDesc Z{};
    const AggDesc &A = sp.A ? *sp.A : Z;
    o << "  static constexpr int n_agg = " << A.n_agg << ", hot = " << sp.hot << ", priv = " << sp.priv << ";\n";
    o << "  static constexpr uint32_t words = " << (A.n_agg + 1) << ", strategy = " << A.strategy
      << ", n_key_cols = " << A.n_key_cols << ", key_words = " << (A.key_words ? A.key_words : 1) << ";\n";
    o << "  static constexpr bool grouped = " << (A.n_key_cols > 0 ? "true" : "false") << ";\n";
    table(o, "uint8_t", "agg_kind", A.n_agg, A.kind, [](uint8_t k) { return k; });
    table(o, "uint32_t", "key_col", A.n_key_cols, A.key_col, [](uint16_t k) { return k; });
    table(o, "uint32_t", "key_w", A.n_key_cols, A.key_width, [](uint8_t k) { return k; });
    table(o, "uint32_t", "key_off", A.n_key_cols, A.key_off, [](uint8_t k) { return k; });
  }
  // output side
  {
    SinkDesc Z{};
    const SinkDesc &K = sp.K ? *sp.K : Z;
    o << "  static constexpr int n_out = " << K.n_out << ", n_lip_build = " << K.n_lip_build << ";\n";
    table(o, "uint32_t", "out_w", K.n_out, K.out_width, [](uint8_t w) { return w; });
    table(o, "uint32_t", "lb_col", K.n_lip_build, K.lip_build_col, [](uint16_t c) { return c; });
    table(o, "uint8_t", "lb_ltype", K.n_lip_build, K.lip_build_ltype, [](uint8_t t) { return t; });
    table(o, "uint32_t", "lb_kind", K.n_lip_build, K.lip_build, [](const LipDesc &l) { return l.kind; });
  }
  // join
  {
    JoinDesc Z{};
    const JoinDesc &J = sp.J ? *sp.J : Z;
    o << "  static constexpr uint32_t j_key_col = " << J.key_col << ", j_type = " << int(J.join_type) << ";\n";
    o << "  static constexpr uint8_t j_key_ltype = " << int(J.key_ltype) << ";\n";
    o << "  static constexpr bool j_dense = " << (J.dense ? "true" : "false") << ";\n";
    o << "  static constexpr bool j_key2 = " << (J.key2_present ? "true" : "false") << ";\n";
    o << "  static constexpr uint32_t j_key2_col = " << J.key2_col << ";\n";
    o << "  static constexpr uint32_t j_null_col = " << (sp.J ? J.null_col : 0xffffu) << ";\n";
    table(o, "uint32_t", "build_w", J.n_build_cols, J.build_cols, [](const ColDesc &c) { return c.width; });
    table(o, "uint32_t", "build_cw", J.n_build_cols, J.build_cols, [](const ColDesc &c) { return c.cw; });
  }
  o << "};\n}  // namespace qs\n";
  // kernel name = family + hash of the description, so launch lists and ncu reports tell queries apart
  static const char *kFamilyName[] = {"qs_scan_agg", "qs_scan_groupby", "qs_scan_select", "qs_join_build", "qs_join_probe"};
  char name[64];
  {
    const std::string q = o.str();
    std::snprintf(name, sizeof(name), "%s_%08x", kFamilyName[sp.family],
                  static_cast<unsigned>(fnv1a(q.data(), q.size(), 0xcbf29ce484222325ull) >> 32));
  }
  if (kernel_name) *kernel_name = name;
  // min-blocks 2 at most: 3-4 resident CTAs are taken when the kernel's registers happen to allow it, never
  // forced by capping registers at 64 (that would spill the heavier probe / select kernels)
  o << "extern \"C\" __global__ void __launch_bounds__(qs::kBlock, " << (sp.ctas_per_sm > 2 ? 2 : sp.ctas_per_sm) << ")\n" << name << "(";
  o << "const __grid_constant__ qs::ScanDesc S, const __grid_constant__ qs::Lits L";
  const char *body = "";
  switch (sp.family) {
    case JF_AGG: o << ", const __grid_constant__ qs::AggDesc A"; body = "scan_agg_body<qs::Q>(smem, S, L, A)"; break;
    case JF_GROUPBY: o << ", const __grid_constant__ qs::AggDesc A"; body = "scan_groupby_body<qs::Q>(smem, S, L, A)"; break;
    case JF_SELECT: o << ", const __grid_constant__ qs::SinkDesc K"; body = "scan_select_body<qs::Q>(smem, S, L, K)"; break;
    case JF_JOIN_BUILD:
      o << ", const __grid_constant__ qs::SinkDesc K, const __grid_constant__ qs::JoinDesc J";
      body = "join_build_body<qs::Q>(smem, S, L, K, J)";
      break;
    case JF_JOIN_PROBE:
      o << ", const __grid_constant__ qs::SinkDesc K, const __grid_constant__ qs::JoinDesc J";
      body = "join_probe_body<qs::Q>(smem, S, L, K, J)";
      break;
  }
  o << ") {\n  extern __shared__ __align__(128) char smem[];\n  qs::" << body << ";\n}\n";
  return o.str();
}

// ----------------------------------------------------------------- compile
static std::string cache_name(const std::string &source) {
  uint64_t h = 0xcbf29ce484222325ull;
  h = fnv1a(source.data(), source.size(), h);
  for (int i = 0; i < kNumHeaders; ++i) h = fnv1a(kHeaders[i].text, std::strlen(kHeaders[i].text), h);
  for (int i = 0; i < kNumOptions; ++i) h = fnv1a(kOptions[i], std::strlen(kOptions[i]), h);
  int major = 0, minor = 0;
  nvrtcVersion(&major, &minor);
  h = fnv1a(&major, sizeof(major), h);
  h = fnv1a(&minor, sizeof(minor), h);
  char buf[32];
  std::snprintf(buf, sizeof(buf), "q%016llx", static_cast<unsigned long long>(h));
  return buf;
}

static std::string cache_dir() {
  if (const char *e = std::getenv("QSGPU_JIT_CACHE")) return e;
  Dl_info info;
  if (dladdr(reinterpret_cast<const void *>(&cache_dir), &info) && info.dli_fname) {
    std::string p = info.dli_fname;
    const size_t slash = p.rfind('/');
    p = slash == std::string::npos ? "." : p.substr(0, slash);
    return p + "/jitcache";
  }
  return "./jitcache";
}

static bool read_file(const std::string &path, std::string *out) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  std::ostringstream ss;
  ss << f.rdbuf();
  *out = ss.str();
  return !out->empty();
}

static void write_file(const std::string &path, const std::string &data) {
  const std::string tmp = path + ".tmp" + std::to_string(static_cast<long>(getpid()));
  {
    std::ofstream f(tmp, std::ios::binary);
    if (!f) return;
    f.write(data.data(), static_cast<std::streamsize>(data.size()));
  }
  std::rename(tmp.c_str(), path.c_str());
}

int jit_compile_only(const std::string &source, std::string *cubin, std::string *log) {
  std::vector<const char *> names, texts;
  for (int i = 0; i < kNumHeaders; ++i) { names.push_back(kHeaders[i].name); texts.push_back(kHeaders[i].text); }
  nvrtcProgram prog = nullptr;
  nvrtcResult r = nvrtcCreateProgram(&prog, source.c_str(), "qs_query.cu", kNumHeaders, texts.data(), names.data());
  if (r != NVRTC_SUCCESS) { if (log) *log = nvrtcGetErrorString(r); return QSGPU_ERR_CUDA; }
  r = nvrtcCompileProgram(prog, kNumOptions, kOptions);
  size_t log_size = 0;
  nvrtcGetProgramLogSize(prog, &log_size);
  if (log && log_size > 1) { log->resize(log_size); nvrtcGetProgramLog(prog, &(*log)[0]); }
  if (r != NVRTC_SUCCESS) {
    if (log && log->empty()) *log = nvrtcGetErrorString(r);
    nvrtcDestroyProgram(&prog);
    return QSGPU_ERR_CUDA;
  }
  size_t n = 0;
  r = nvrtcGetCUBINSize(prog, &n);
  if (r == NVRTC_SUCCESS && n > 0) { cubin->resize(n); r = nvrtcGetCUBIN(prog, &(*cubin)[0]); }
  nvrtcDestroyProgram(&prog);
  if (r != NVRTC_SUCCESS || n == 0) { if (log) *log = "NVRTC produced no cubin"; return QSGPU_ERR_CUDA; }
  return QSGPU_OK;
}

// Binary description of everything jit_source() prints, field for field: the key of the hot in-memory cache.  Printing
// the source (a few KB through an ostringstream) and looking a multi-KB string up cost ~20 us per launch, on the
// critical path of every query (the scan cannot start before its kernel is found); this is a few hundred bytes.
// QSGPU_JIT_VERIFY=1 re-prints the source on every hit and checks it is the one the kernel was compiled from (the
// test suite runs with it): the two descriptions cannot drift apart unnoticed.
static std::string jit_key(const JitSpec &sp) {
  std::string k;
  k.reserve(512);
  auto put = [&](uint64_t v, int bytes) { k.append(reinterpret_cast<const char *>(&v), static_cast<size_t>(bytes)); };
  const ScanDesc &S = *sp.S;
  const Program &P = *sp.P;
  put(static_cast<uint64_t>(sp.family), 1); put(static_cast<uint64_t>(sp.hot), 2); put(static_cast<uint64_t>(sp.priv), 1);
  put(static_cast<uint64_t>(sp.ctas_per_sm > 2 ? 2 : sp.ctas_per_sm), 1);
  put(S.n_cols, 4); put(S.n_stages, 4); put(S.stage_bytes, 4);
  for (uint32_t c = 0; c < S.n_cols; ++c) {
    const ColDesc &C = S.cols[c];
    put(C.width, 4); put(C.smem_off, 4); put(C.cw, 1); put(C.code_off, 4); put(C.expand, 1); put(C.dict_smem, 1); put(C.dict_soff, 4);
  }
  put(P.n_pred, 4); put(P.n_mid, 4); put(P.n_total, 4);
  for (uint32_t i = 0; i < P.n_total; ++i) {
    const Instr &in = P.code[i];
    put(in.op, 1); put(in.type, 1); put(in.leaf, 1); put(in.ltype, 1); put(in.arg, 2); put(in.flags, 1); put(in.aux, 1);
  }
  put(warp_compact_requested() ? 1 : 0, 1);
  put(S.n_lip, 4);
  for (uint32_t i = 0; i < S.n_lip; ++i) { put(S.lip[i].kind, 4); put(S.lip[i].is_anti, 4); put(S.lip[i].smem_off, 4); }
  if (sp.A) {
    const AggDesc &A = *sp.A;
    put(1, 1); put(A.n_agg, 4); put(A.strategy, 4); put(A.n_key_cols, 4); put(A.key_words, 4);
    for (uint32_t i = 0; i < A.n_agg; ++i) put(A.kind[i], 1);
    for (uint32_t i = 0; i < A.n_key_cols; ++i) { put(A.key_col[i], 2); put(A.key_width[i], 1); put(A.key_off[i], 1); }
  } else put(0, 1);
  if (sp.K) {
    const SinkDesc &K = *sp.K;
    put(1, 1); put(K.n_out, 4); put(K.n_lip_build, 4);
    for (uint32_t i = 0; i < K.n_out; ++i) put(K.out_width[i], 1);
    for (uint32_t i = 0; i < K.n_lip_build; ++i) { put(K.lip_build_col[i], 2); put(K.lip_build_ltype[i], 1); put(K.lip_build[i].kind, 4); }
  } else put(0, 1);
  if (sp.J) {
    const JoinDesc &J = *sp.J;
    put(1, 1); put(J.key_col, 2); put(J.join_type, 1); put(J.key_ltype, 1); put(J.dense, 4); put(J.key2_present, 1); put(J.key2_col, 2); put(J.null_col, 2);
    put(J.n_build_cols, 4);
    for (uint32_t i = 0; i < J.n_build_cols; ++i) { put(J.build_cols[i].width, 4); put(J.build_cols[i].cw, 1); }
  } else put(0, 1);
  return k;
}

static std::unordered_map<std::string, std::pair<JitKernel *, std::string>> g_by_key;   // key -> (kernel, its source)

int jit_get(const JitSpec &spec, JitKernel **out) {
  const std::string key = jit_key(spec);
  {
    std::lock_guard<std::mutex> lk(g_jit_mutex);
    auto hit = g_by_key.find(key);
    if (hit != g_by_key.end()) {
      static const bool verify = std::getenv("QSGPU_JIT_VERIFY") != nullptr;
      if (verify && jit_source(spec) != hit->second.second) {
        set_error(QSGPU_ERR_INVALID, "internal: jit_key() and jit_source() disagree (a field printed by one is missing in the other)");
        return QSGPU_ERR_INVALID;
      }
      ++g_mem_hits;
      *out = hit->second.first;
      return QSGPU_OK;
    }
  }
  std::string kernel_name;
  const std::string source = jit_source(spec, &kernel_name);
  std::lock_guard<std::mutex> lk(g_jit_mutex);
  auto it = g_kernels.find(source);
  if (it != g_kernels.end()) {
    ++g_mem_hits;
    *out = it->second.get();
    g_by_key.emplace(key, std::make_pair(it->second.get(), source));
    return QSGPU_OK;
  }

  const std::string name = cache_name(source);
  const std::string dir = cache_dir();
  const std::string path = dir + "/" + name + ".cubin";
  std::string cubin;
  if (read_file(path, &cubin)) {
    ++g_disk_hits;
  } else {
    std::string log;
    const int st = jit_compile_only(source, &cubin, &log);
    if (st != QSGPU_OK) {
      if (log.size() > 3000) log.resize(3000);
      set_error(st, "query compilation (NVRTC) failed: " + log);
      if (std::getenv("QSGPU_JIT_DUMP")) write_file(std::string(std::getenv("QSGPU_JIT_DUMP")) + "/failed_" + name + ".cu", source);
      return st;
    }
    ++g_compiled;
    mkdir(dir.c_str(), 0755);
    write_file(path, cubin);
    write_file(dir + "/" + name + ".cu", source);
  }
  std::unique_ptr<JitKernel> k(new JitKernel);
  k->name = kernel_name;
  cudaError_t e = cudaLibraryLoadData(&k->lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
  if (e != cudaSuccess) return cuda_fail(e, "cudaLibraryLoadData(query kernel)");
  e = cudaLibraryGetKernel(&k->fn, k->lib, kernel_name.c_str());
  if (e != cudaSuccess) return cuda_fail(e, "cudaLibraryGetKernel(query kernel)");
  *out = k.get();
  g_by_key.emplace(key, std::make_pair(k.get(), source));
  g_kernels.emplace(source, std::move(k));
  return QSGPU_OK;
}

cudaError_t jit_launch(JitKernel *k, int grid, size_t smem, cudaStream_t st, void **args) {
  int dev = 0;
  cudaGetDevice(&dev);
  size_t &set = k->smem_set[dev & 15];
  if (smem > set) {
    std::lock_guard<std::mutex> lk(g_jit_mutex);
    if (smem > set) {
      cudaError_t e = cudaFuncSetAttribute(reinterpret_cast<const void *>(k->fn),
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      if (e != cudaSuccess) return e;
      set = smem;
    }
  }
  return cudaLaunchKernel(reinterpret_cast<const void *>(k->fn), dim3(grid), dim3(kBlock), args, smem, st);
}

int jit_occupancy(JitKernel *k, size_t smem) {
  std::lock_guard<std::mutex> lk(g_jit_mutex);
  if (k->occ_smem == smem) return k->occ;
  int dev = 0;
  cudaGetDevice(&dev);
  size_t &set = k->smem_set[dev & 15];
  if (smem > set) {
    if (cudaFuncSetAttribute(reinterpret_cast<const void *>(k->fn), cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess) { cudaGetLastError(); return 0; }
    set = smem;
  }
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, reinterpret_cast<const void *>(k->fn), kBlock, smem) != cudaSuccess) { cudaGetLastError(); return 0; }
  k->occ_smem = smem;
  k->occ = occ;
  return occ;
}

void jit_stats(uint64_t *compiled, uint64_t *disk_hits, uint64_t *mem_hits) {
  std::lock_guard<std::mutex> lk(g_jit_mutex);
  *compiled = g_compiled;
  *disk_hits = g_disk_hits;
  *mem_hits = g_mem_hits;
}

}  // namespace qs
