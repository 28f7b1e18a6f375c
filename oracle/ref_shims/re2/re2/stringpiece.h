#pragma once
#include <cstring>
#include <string>
namespace re2 {
class StringPiece {
 public:
  StringPiece() : p_(nullptr), n_(0) {}
  StringPiece(const char *p) : p_(p), n_(std::strlen(p)) {}
  StringPiece(const std::string &s) : p_(s.data()), n_(s.size()) {}
  StringPiece(const char *p, std::size_t n) : p_(p), n_(n) {}
  void set(const char *p, std::size_t n) { p_ = p; n_ = n; }
  const char* data() const { return p_; }
  std::size_t size() const { return n_; }
  std::string as_string() const { return std::string(p_, n_); }
 private:
  const char *p_; std::size_t n_;
};
}  // namespace re2
