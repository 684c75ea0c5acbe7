// Minimal JSON reader (objects, arrays, numbers, strings, true/false/null) for the backbone config that the reference
// keeps in configs/*.json ("backbone.args", read by inference/sample.py:266-274).  No external dependency.
#pragma once
#include <cctype>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace ivid {

struct JsonValue {
  enum Kind { kNull, kBool, kNumber, kString, kArray, kObject } kind = kNull;
  bool b = false;
  double num = 0.0;
  std::string str;
  std::vector<JsonValue> arr;
  std::map<std::string, JsonValue> obj;

  bool has(const std::string& k) const { return kind == kObject && obj.count(k) && obj.at(k).kind != kNull; }
  const JsonValue& at(const std::string& k) const { return obj.at(k); }
};

class JsonParser {
 public:
  explicit JsonParser(const std::string& s) : s_(s) {}
  JsonValue parse() {
    JsonValue v = value();
    ws();
    if (i_ != s_.size()) fail("trailing characters");
    return v;
  }

 private:
  const std::string& s_;
  size_t i_ = 0;
  [[noreturn]] void fail(const char* m) { throw std::runtime_error(std::string("config json: ") + m); }
  void ws() { while (i_ < s_.size() && std::isspace(static_cast<unsigned char>(s_[i_]))) ++i_; }
  bool eat(const char* lit) {
    size_t n = std::char_traits<char>::length(lit);
    if (s_.compare(i_, n, lit) == 0) { i_ += n; return true; }
    return false;
  }
  JsonValue value() {
    ws();
    if (i_ >= s_.size()) fail("unexpected end");
    JsonValue v;
    const char c = s_[i_];
    if (c == '{') {
      v.kind = JsonValue::kObject;
      ++i_;
      ws();
      if (s_[i_] == '}') { ++i_; return v; }
      while (true) {
        ws();
        JsonValue k = string_value();
        ws();
        if (s_[i_] != ':') fail("expected ':'");
        ++i_;
        v.obj[k.str] = value();
        ws();
        if (s_[i_] == ',') { ++i_; continue; }
        if (s_[i_] == '}') { ++i_; break; }
        fail("expected ',' or '}'");
      }
    } else if (c == '[') {
      v.kind = JsonValue::kArray;
      ++i_;
      ws();
      if (s_[i_] == ']') { ++i_; return v; }
      while (true) {
        v.arr.push_back(value());
        ws();
        if (s_[i_] == ',') { ++i_; continue; }
        if (s_[i_] == ']') { ++i_; break; }
        fail("expected ',' or ']'");
      }
    } else if (c == '"') {
      v = string_value();
    } else if (eat("true")) {
      v.kind = JsonValue::kBool; v.b = true;
    } else if (eat("false")) {
      v.kind = JsonValue::kBool; v.b = false;
    } else if (eat("null")) {
      v.kind = JsonValue::kNull;
    } else {
      char* end = nullptr;
      v.num = std::strtod(s_.c_str() + i_, &end);
      if (end == s_.c_str() + i_) fail("bad number");
      i_ = static_cast<size_t>(end - s_.c_str());
      v.kind = JsonValue::kNumber;
    }
    return v;
  }
  JsonValue string_value() {
    if (s_[i_] != '"') fail("expected string");
    ++i_;
    JsonValue v;
    v.kind = JsonValue::kString;
    while (i_ < s_.size() && s_[i_] != '"') {
      if (s_[i_] == '\\' && i_ + 1 < s_.size()) { ++i_; }
      v.str.push_back(s_[i_++]);
    }
    if (i_ >= s_.size()) fail("unterminated string");
    ++i_;
    return v;
  }
};

}  // namespace ivid
