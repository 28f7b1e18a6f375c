#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <functional>
#include <vector>
#include <map>
#include <string>
namespace gflags {
struct FlagReg { void *ptr; char kind; };  // b,i,l,u,d,s
inline std::map<std::string, FlagReg>& Registry() { static std::map<std::string, FlagReg> r; return r; }
struct Registrar { Registrar(const char *n, void *p, char k) { Registry()[n] = FlagReg{p, k}; } };
inline std::vector<std::function<bool()>>& Validators() { static std::vector<std::function<bool()>> v; return v; }
template <typename T, typename F> inline bool RegisterFlagValidator(const T *flag, F fn) { Validators().push_back([flag, fn]() { return fn("", *flag); }); return true; }
inline void SetUsageMessage(const std::string&) {}
inline void SetVersionString(const std::string&) {}
inline bool SetOne(const std::string &name, const char *val, bool has_val) {
  auto it = Registry().find(name);
  if (it == Registry().end()) {
    if (name.rfind("no", 0) == 0) { auto jt = Registry().find(name.substr(2)); if (jt != Registry().end() && jt->second.kind=='b') { *static_cast<bool*>(jt->second.ptr) = false; return true; } }
    return false;
  }
  FlagReg &r = it->second;
  switch (r.kind) {
    case 'b': *static_cast<bool*>(r.ptr) = !has_val || !strcmp(val,"true") || !strcmp(val,"1") || !strcmp(val,"t") || !strcmp(val,"yes"); break;
    case 'i': *static_cast<std::int32_t*>(r.ptr) = std::atoi(val); break;
    case 'l': *static_cast<std::int64_t*>(r.ptr) = std::atoll(val); break;
    case 'u': *static_cast<std::uint64_t*>(r.ptr) = std::strtoull(val, nullptr, 10); break;
    case 'd': *static_cast<double*>(r.ptr) = std::atof(val); break;
    case 's': *static_cast<std::string*>(r.ptr) = val; break;
  }
  return true;
}
inline std::uint32_t ParseCommandLineFlags(int *argc, char ***argv, bool) {
  for (int i = 1; i < *argc; ++i) {
    const char *a = (*argv)[i];
    if (a[0] != '-') continue;
    while (*a == '-') ++a;
    const char *eq = std::strchr(a, '=');
    std::string name = eq ? std::string(a, eq - a) : std::string(a);
    if (!SetOne(name, eq ? eq + 1 : "", eq != nullptr)) { std::fprintf(stderr, "unknown flag %s\n", (*argv)[i]); std::exit(1); }
  }
  for (auto &v : Validators()) { if (!v()) { std::fprintf(stderr, "flag validation failed\n"); std::exit(1); } }
  return *argc;
}
}  // namespace gflags
namespace google { using namespace gflags; }
#define QS_DEFINE_FLAG(type, kind, name, val) \
  type FLAGS_##name = val; \
  static ::gflags::Registrar qs_flag_registrar_##name(#name, &FLAGS_##name, kind)
#define DEFINE_bool(name, val, txt)   QS_DEFINE_FLAG(bool, 'b', name, val)
#define DEFINE_int32(name, val, txt)  QS_DEFINE_FLAG(std::int32_t, 'i', name, val)
#define DEFINE_int64(name, val, txt)  QS_DEFINE_FLAG(std::int64_t, 'l', name, val)
#define DEFINE_uint64(name, val, txt) QS_DEFINE_FLAG(std::uint64_t, 'u', name, val)
#define DEFINE_double(name, val, txt) QS_DEFINE_FLAG(double, 'd', name, val)
#define DEFINE_string(name, val, txt) QS_DEFINE_FLAG(std::string, 's', name, val)
#define DEFINE_validator(name, fn) static const bool qs_flag_validator_##name = true
#define DECLARE_bool(name)   extern bool FLAGS_##name
#define DECLARE_int32(name)  extern std::int32_t FLAGS_##name
#define DECLARE_int64(name)  extern std::int64_t FLAGS_##name
#define DECLARE_uint64(name) extern std::uint64_t FLAGS_##name
#define DECLARE_double(name) extern double FLAGS_##name
#define DECLARE_string(name) extern std::string FLAGS_##name
