#pragma once
#include <regex>
#include <string>
#include "re2/stringpiece.h"
namespace re2 {
class RE2 {
 public:
  RE2(const StringPiece &p) : re_(p.as_string(), std::regex::ECMAScript) {}
  RE2(const std::string &p) : re_(p, std::regex::ECMAScript) {}
  RE2(const char *p) : re_(p, std::regex::ECMAScript) {}
  static bool FullMatch(const StringPiece &text, const RE2 &re) {
    return std::regex_match(text.data(), text.data() + text.size(), re.re_);
  }
  static std::string QuoteMeta(const StringPiece &s) {
    std::string out;
    for (std::size_t i = 0; i < s.size(); ++i) { char c = s.data()[i]; if (!isalnum((unsigned char)c) && c != '_') out.push_back('\\'); out.push_back(c); }
    return out;
  }
 private:
  std::regex re_;
};
}  // namespace re2
