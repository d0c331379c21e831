// Minimal JSON reader (objects, arrays, strings, numbers, true/false/null) - enough for config.json and the
// safetensors header.  No external dependencies.
#pragma once
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace dsocr {

struct Json {
  enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
  bool b = false;
  double num = 0;
  std::string str;
  std::vector<Json> arr;
  std::vector<std::pair<std::string, Json>> obj;

  const Json* get(const std::string& key) const {
    if (kind != Obj) return nullptr;
    for (auto& kv : obj) if (kv.first == key) return &kv.second;
    return nullptr;
  }
  const Json& at(const std::string& key) const {
    const Json* j = get(key);
    if (!j) throw std::runtime_error("json: missing key '" + key + "'");
    return *j;
  }
  bool is_null() const { return kind == Null; }
  long long as_int(long long dflt) const { return kind == Num ? (long long)num : dflt; }
  double as_num(double dflt) const { return kind == Num ? num : dflt; }
  bool as_bool(bool dflt) const { return kind == Bool ? b : dflt; }
};

class JsonParser {
 public:
  JsonParser(const char* p, size_t n) : p_(p), end_(p + n) {}
  Json parse() {
    Json j = value();
    ws();
    return j;
  }

 private:
  const char* p_;
  const char* end_;
  void ws() { while (p_ < end_ && (*p_ == ' ' || *p_ == '\n' || *p_ == '\t' || *p_ == '\r')) ++p_; }
  [[noreturn]] void fail(const char* m) { throw std::runtime_error(std::string("json parse error: ") + m); }
  Json value() {
    ws();
    if (p_ >= end_) fail("unexpected end");
    Json j;
    switch (*p_) {
      case '{': {
        j.kind = Json::Obj; ++p_; ws();
        if (p_ < end_ && *p_ == '}') { ++p_; return j; }
        for (;;) {
          ws();
          std::string k = string();
          ws();
          if (p_ >= end_ || *p_ != ':') fail("expected ':'");
          ++p_;
          j.obj.emplace_back(std::move(k), value());
          ws();
          if (p_ < end_ && *p_ == ',') { ++p_; continue; }
          if (p_ < end_ && *p_ == '}') { ++p_; return j; }
          fail("expected ',' or '}'");
        }
      }
      case '[': {
        j.kind = Json::Arr; ++p_; ws();
        if (p_ < end_ && *p_ == ']') { ++p_; return j; }
        for (;;) {
          j.arr.push_back(value());
          ws();
          if (p_ < end_ && *p_ == ',') { ++p_; continue; }
          if (p_ < end_ && *p_ == ']') { ++p_; return j; }
          fail("expected ',' or ']'");
        }
      }
      case '"': j.kind = Json::Str; j.str = string(); return j;
      case 't': if (end_ - p_ >= 4) { p_ += 4; j.kind = Json::Bool; j.b = true; return j; } fail("bad literal");
      case 'f': if (end_ - p_ >= 5) { p_ += 5; j.kind = Json::Bool; j.b = false; return j; } fail("bad literal");
      case 'n': if (end_ - p_ >= 4) { p_ += 4; j.kind = Json::Null; return j; } fail("bad literal");
      default: {
        char* e = nullptr;
        j.num = std::strtod(p_, &e);
        if (e == p_) fail("bad number");
        p_ = e; j.kind = Json::Num; return j;
      }
    }
  }
  std::string string() {
    if (p_ >= end_ || *p_ != '"') fail("expected string");
    ++p_;
    std::string s;
    while (p_ < end_ && *p_ != '"') {
      if (*p_ == '\\' && p_ + 1 < end_) {
        ++p_;
        switch (*p_) {
          case 'n': s += '\n'; break;
          case 't': s += '\t'; break;
          case 'r': s += '\r'; break;
          case 'b': s += '\b'; break;
          case 'f': s += '\f'; break;
          case 'u': s += '?'; p_ += 4; break;  // non-ASCII escapes are not needed for tensor names / config keys
          default: s += *p_;
        }
        ++p_;
      } else {
        s += *p_++;
      }
    }
    if (p_ >= end_) fail("unterminated string");
    ++p_;
    return s;
  }
};

}  // namespace dsocr
