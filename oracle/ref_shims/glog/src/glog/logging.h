#pragma once
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>
namespace google {
enum { INFO = 0, WARNING = 1, ERROR = 2, FATAL = 3 };
inline int& MinLevel() { static int l = WARNING; return l; }
inline void InitGoogleLogging(const char*) {}
class LogMessage {
 public:
  LogMessage(const char *f, int l, int sev) : sev_(sev) { s_ << "[" << f << ":" << l << "] "; }
  ~LogMessage() { if (sev_ >= MinLevel()) { s_ << "\n"; std::cerr << s_.str(); } }
  std::ostream& stream() { return s_; }
 private:
  std::ostringstream s_; int sev_;
};
class LogMessageFatal {
 public:
  LogMessageFatal(const char *f, int l) { s_ << "[FATAL " << f << ":" << l << "] "; }
  [[noreturn]] ~LogMessageFatal() { s_ << "\n"; std::cerr << s_.str(); std::abort(); }
  std::ostream& stream() { return s_; }
 private:
  std::ostringstream s_;
};
struct Voidify { void operator&(std::ostream&) {} };
template <typename T> T&& CheckNotNull(const char *f, int l, const char *n, T &&t) {
  if (t == nullptr) { LogMessageFatal(f, l).stream() << n; }
  return std::forward<T>(t);
}
}  // namespace google
#define QS_LOG_INFO    ::google::LogMessage(__FILE__, __LINE__, ::google::INFO).stream()
#define QS_LOG_WARNING ::google::LogMessage(__FILE__, __LINE__, ::google::WARNING).stream()
#define QS_LOG_ERROR   ::google::LogMessage(__FILE__, __LINE__, ::google::ERROR).stream()
#define QS_LOG_FATAL   ::google::LogMessageFatal(__FILE__, __LINE__).stream()
#define LOG(sev) QS_LOG_##sev
#define LOG_IF(sev, cond) !(cond) ? (void)0 : ::google::Voidify() & LOG(sev)
#define LOG_FIRST_N(sev, n) LOG(sev)
#define VLOG_IS_ON(n) false
#define VLOG(n) true ? (void)0 : ::google::Voidify() & LOG(INFO)
#define CHECK(cond) (cond) ? (void)0 : ::google::Voidify() & QS_LOG_FATAL << "Check failed: " #cond " "
#define QS_CHECK_OP(a, b, op) CHECK((a) op (b))
#define CHECK_EQ(a, b) QS_CHECK_OP(a, b, ==)
#define CHECK_NE(a, b) QS_CHECK_OP(a, b, !=)
#define CHECK_LT(a, b) QS_CHECK_OP(a, b, <)
#define CHECK_LE(a, b) QS_CHECK_OP(a, b, <=)
#define CHECK_GT(a, b) QS_CHECK_OP(a, b, >)
#define CHECK_GE(a, b) QS_CHECK_OP(a, b, >=)
#define CHECK_NOTNULL(v) ::google::CheckNotNull(__FILE__, __LINE__, "'" #v "' Must be non NULL", (v))
#ifdef NDEBUG
#define DLOG(sev) true ? (void)0 : ::google::Voidify() & LOG(sev)
#define DVLOG(n) VLOG(n)
#define DCHECK(cond) while (false) CHECK(cond)
#define DCHECK_EQ(a, b) while (false) CHECK_EQ(a, b)
#define DCHECK_NE(a, b) while (false) CHECK_NE(a, b)
#define DCHECK_LT(a, b) while (false) CHECK_LT(a, b)
#define DCHECK_LE(a, b) while (false) CHECK_LE(a, b)
#define DCHECK_GT(a, b) while (false) CHECK_GT(a, b)
#define DCHECK_GE(a, b) while (false) CHECK_GE(a, b)
#define DCHECK_NOTNULL(v) (v)
#else
#define DLOG(sev) LOG(sev)
#define DVLOG(n) VLOG(n)
#define DCHECK(cond) CHECK(cond)
#define DCHECK_EQ(a, b) CHECK_EQ(a, b)
#define DCHECK_NE(a, b) CHECK_NE(a, b)
#define DCHECK_LT(a, b) CHECK_LT(a, b)
#define DCHECK_LE(a, b) CHECK_LE(a, b)
#define DCHECK_GT(a, b) CHECK_GT(a, b)
#define DCHECK_GE(a, b) CHECK_GE(a, b)
#define DCHECK_NOTNULL(v) CHECK_NOTNULL(v)
#endif
