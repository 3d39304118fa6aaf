// json.h -- minimal JSON reader for the reference's "opts" parameter
// (bigseqkit/helper.go:47-66: the option struct encoded by encoding/json).
#pragma once
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace bsk {

struct JValue {
  enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
  bool b = false;
  double num = 0;
  std::string str;
  std::vector<JValue> arr;
  std::vector<std::pair<std::string, JValue>> obj;

  const JValue *get(const std::string &key) const {
    if (kind != Obj) return nullptr;
    for (auto &kv : obj)
      if (kv.first == key) return &kv.second;
    return nullptr;
  }
};

class JParser {
 public:
  explicit JParser(const std::string &s) : s_(s) {}
  bool parse(JValue &out, std::string &err) {
    ws();
    if (!value(out, err)) return false;
    ws();
    if (i_ != s_.size()) { err = "trailing characters in JSON options"; return false; }
    return true;
  }

 private:
  const std::string &s_;
  size_t i_ = 0;
  void ws() { while (i_ < s_.size() && (s_[i_] == ' ' || s_[i_] == '\n' || s_[i_] == '\t' || s_[i_] == '\r')) i_++; }
  bool lit(const char *w) {
    size_t l = strlen(w);
    if (s_.compare(i_, l, w) == 0) { i_ += l; return true; }
    return false;
  }
  static void utf8(std::string &o, unsigned cp) {
    if (cp < 0x80) o += (char)cp;
    else if (cp < 0x800) { o += (char)(0xC0 | (cp >> 6)); o += (char)(0x80 | (cp & 0x3F)); }
    else { o += (char)(0xE0 | (cp >> 12)); o += (char)(0x80 | ((cp >> 6) & 0x3F)); o += (char)(0x80 | (cp & 0x3F)); }
  }
  bool string(std::string &o, std::string &err) {
    i_++;  // opening quote
    while (i_ < s_.size() && s_[i_] != '"') {
      char c = s_[i_++];
      if (c != '\\') { o += c; continue; }
      if (i_ >= s_.size()) break;
      char e = s_[i_++];
      switch (e) {
        case 'n': o += '\n'; break;
        case 't': o += '\t'; break;
        case 'r': o += '\r'; break;
        case 'b': o += '\b'; break;
        case 'f': o += '\f'; break;
        case 'u': {
          if (i_ + 4 > s_.size()) { err = "bad \\u escape in JSON options"; return false; }
          unsigned cp = (unsigned)strtoul(s_.substr(i_, 4).c_str(), nullptr, 16);
          i_ += 4;
          utf8(o, cp);
          break;
        }
        default: o += e;
      }
    }
    if (i_ >= s_.size()) { err = "unterminated string in JSON options"; return false; }
    i_++;
    return true;
  }
  bool value(JValue &v, std::string &err) {
    ws();
    if (i_ >= s_.size()) { err = "unexpected end of JSON options"; return false; }
    char c = s_[i_];
    if (c == '{') {
      v.kind = JValue::Obj;
      i_++;
      ws();
      if (i_ < s_.size() && s_[i_] == '}') { i_++; return true; }
      for (;;) {
        ws();
        if (i_ >= s_.size() || s_[i_] != '"') { err = "expected key in JSON options"; return false; }
        std::string k;
        if (!string(k, err)) return false;
        ws();
        if (i_ >= s_.size() || s_[i_] != ':') { err = "expected ':' in JSON options"; return false; }
        i_++;
        JValue child;
        if (!value(child, err)) return false;
        v.obj.emplace_back(std::move(k), std::move(child));
        ws();
        if (i_ < s_.size() && s_[i_] == ',') { i_++; continue; }
        if (i_ < s_.size() && s_[i_] == '}') { i_++; return true; }
        err = "expected ',' or '}' in JSON options";
        return false;
      }
    }
    if (c == '[') {
      v.kind = JValue::Arr;
      i_++;
      ws();
      if (i_ < s_.size() && s_[i_] == ']') { i_++; return true; }
      for (;;) {
        JValue child;
        if (!value(child, err)) return false;
        v.arr.push_back(std::move(child));
        ws();
        if (i_ < s_.size() && s_[i_] == ',') { i_++; continue; }
        if (i_ < s_.size() && s_[i_] == ']') { i_++; return true; }
        err = "expected ',' or ']' in JSON options";
        return false;
      }
    }
    if (c == '"') { v.kind = JValue::Str; return string(v.str, err); }
    if (lit("true")) { v.kind = JValue::Bool; v.b = true; return true; }
    if (lit("false")) { v.kind = JValue::Bool; v.b = false; return true; }
    if (lit("null")) { v.kind = JValue::Null; return true; }
    char *end = nullptr;
    double d = strtod(s_.c_str() + i_, &end);
    if (end == s_.c_str() + i_) { err = "invalid value in JSON options"; return false; }
    i_ = (size_t)(end - s_.c_str());
    v.kind = JValue::Num;
    v.num = d;
    return true;
  }
};

}  // namespace bsk
