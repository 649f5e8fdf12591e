#include "json.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>

namespace gph {

namespace {
const Json kNull;

struct Parser {
  const std::string& t;
  size_t i = 0;
  int struct_depth = 0;
  explicit Parser(const std::string& text) : t(text) {}

  [[noreturn]] void err(const char* what) const {
    throw std::runtime_error(std::string("JSON: ") + what + " at offset " + std::to_string(i));
  }
  void ws() {
    while (i < t.size() && (t[i] == ' ' || t[i] == '\n' || t[i] == '\t' || t[i] == '\r')) ++i;
  }
  bool lit(const char* s) {
    size_t n = 0;
    while (s[n]) ++n;
    if (t.compare(i, n, s) == 0) {
      i += n;
      return true;
    }
    return false;
  }
  static void utf8(std::string& out, uint32_t cp) {
    if (cp < 0x80) out += (char)cp;
    else if (cp < 0x800) out += (char)(0xC0 | (cp >> 6)), out += (char)(0x80 | (cp & 0x3F));
    else if (cp < 0x10000)
      out += (char)(0xE0 | (cp >> 12)), out += (char)(0x80 | ((cp >> 6) & 0x3F)), out += (char)(0x80 | (cp & 0x3F));
    else
      out += (char)(0xF0 | (cp >> 18)), out += (char)(0x80 | ((cp >> 12) & 0x3F)),
          out += (char)(0x80 | ((cp >> 6) & 0x3F)), out += (char)(0x80 | (cp & 0x3F));
  }
  uint32_t hex4() {
    if (i + 4 > t.size()) err("short \\u escape");
    uint32_t v = 0;
    for (int k = 0; k < 4; ++k) {
      char c = t[i++];
      v <<= 4;
      if (c >= '0' && c <= '9') v |= c - '0';
      else if (c >= 'a' && c <= 'f') v |= c - 'a' + 10;
      else if (c >= 'A' && c <= 'F') v |= c - 'A' + 10;
      else err("bad \\u escape");
    }
    return v;
  }
  std::string str() {
    if (t[i] != '"') err("expected string");
    ++i;
    std::string out;
    while (true) {
      // the run of plain characters up to the next quote or escape is appended in one piece
      size_t j = i;
      while (j < t.size() && t[j] != '"' && t[j] != '\\') ++j;
      if (j > i) out.append(t, i, j - i), i = j;
      if (i >= t.size()) err("unterminated string");
      char c = t[i++];
      if (c == '"') break;
      if (i >= t.size()) err("unterminated escape");
      char e = t[i++];
      switch (e) {
        case '"': out += '"'; break;
        case '\\': out += '\\'; break;
        case '/': out += '/'; break;
        case 'b': out += '\b'; break;
        case 'f': out += '\f'; break;
        case 'n': out += '\n'; break;
        case 'r': out += '\r'; break;
        case 't': out += '\t'; break;
        case 'u': {
          uint32_t cp = hex4();
          if (cp >= 0xD800 && cp <= 0xDBFF && i + 1 < t.size() && t[i] == '\\' && t[i + 1] == 'u') {
            i += 2;
            uint32_t lo = hex4();
            cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
          }
          utf8(out, cp);
          break;
        }
        default: err("bad escape");
      }
    }
    return out;
  }
  Json value(int depth) {
    if (depth > 200) err("nesting too deep");
    ws();
    if (i >= t.size()) err("unexpected end");
    char c = t[i];
    if (c == '{') {
      ++i;
      Json o = Json::object();
      ws();
      if (i < t.size() && t[i] == '}') { ++i; return o; }
      while (true) {
        ws();
        std::string k = str();
        ws();
        if (i >= t.size() || t[i] != ':') err("expected ':'");
        ++i;
        if (depth < struct_depth && o.find(k)) err("duplicate member name");
        o.set(k, value(depth + 1));
        ws();
        if (i < t.size() && t[i] == ',') { ++i; continue; }
        if (i < t.size() && t[i] == '}') { ++i; return o; }
        err("expected ',' or '}'");
      }
    }
    if (c == '[') {
      ++i;
      Json a = Json::array();
      ws();
      if (i < t.size() && t[i] == ']') { ++i; return a; }
      while (true) {
        a.push(value(depth + 1));
        ws();
        if (i < t.size() && t[i] == ',') { ++i; continue; }
        if (i < t.size() && t[i] == ']') { ++i; return a; }
        err("expected ',' or ']'");
      }
    }
    if (c == '"') return Json(str());
    if (lit("true")) return Json(true);
    if (lit("false")) return Json(false);
    if (lit("null")) return Json();
    char* end = nullptr;
    double d = strtod(t.c_str() + i, &end);
    if (end == t.c_str() + i) err("unexpected character");
    i = (size_t)(end - t.c_str());
    return Json(d);
  }
};
}  // namespace

const Json* Json::find(const std::string& key) const {
  if (type_ != Type::Object) return nullptr;
  for (const Member& m : o_)
    if (m.first == key) return &m.second;
  return nullptr;
}
const Json& Json::operator[](const std::string& key) const {
  const Json* j = find(key);
  return j ? *j : kNull;
}
const Json& Json::operator[](size_t i) const { return type_ == Type::Array && i < a_.size() ? a_[i] : kNull; }

Json& Json::set(const std::string& key, Json v) {
  if (type_ != Type::Object) *this = Json::object();
  for (Member& m : o_)
    if (m.first == key) {
      m.second = std::move(v);
      return m.second;
    }
  if (o_.empty()) o_.reserve(16);  // label maps / small objects: no regrowth (a Member is ~200 bytes)
  o_.emplace_back(key, std::move(v));
  return o_.back().second;
}
Json& Json::push(Json v) {
  if (type_ != Type::Array) *this = Json::array();
  a_.push_back(std::move(v));
  return a_.back();
}

std::string json_escape(const std::string& s) {
  std::string out;
  for (unsigned char c : s) {
    switch (c) {
      case '"': out += "\\\""; break;
      case '\\': out += "\\\\"; break;
      case '\n': out += "\\n"; break;
      case '\r': out += "\\r"; break;
      case '\t': out += "\\t"; break;
      default:
        if (c < 0x20) {
          char b[8];
          snprintf(b, sizeof b, "\\u%04x", c);
          out += b;
        } else {
          out += (char)c;
        }
    }
  }
  return out;
}

void Json::dump_to(std::string& out) const {
  switch (type_) {
    case Type::Null: out += "null"; break;
    case Type::Bool: out += b_ ? "true" : "false"; break;
    case Type::Number: {
      if (std::isfinite(n_) && n_ == std::floor(n_) && std::fabs(n_) < 9e15) {
        out += std::to_string((long long)n_);
      } else {
        char b[40];
        snprintf(b, sizeof b, "%.17g", n_);
        out += b;
      }
      break;
    }
    case Type::String: out += '"', out += json_escape(s_), out += '"'; break;
    case Type::Array: {
      out += '[';
      for (size_t k = 0; k < a_.size(); ++k) {
        if (k) out += ',';
        a_[k].dump_to(out);
      }
      out += ']';
      break;
    }
    case Type::Object: {
      out += '{';
      for (size_t k = 0; k < o_.size(); ++k) {
        if (k) out += ',';
        out += '"', out += json_escape(o_[k].first), out += "\":";
        o_[k].second.dump_to(out);
      }
      out += '}';
      break;
    }
  }
}
std::string Json::dump() const {
  std::string s;
  dump_to(s);
  return s;
}

Json Json::parse(const std::string& text, int struct_depth) {
  Parser p(text);
  p.struct_depth = struct_depth;
  Json v = p.value(0);
  p.ws();
  if (p.i != text.size()) p.err("trailing characters");
  return v;
}

Json Json::parse_file(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("cannot open " + path);
  std::stringstream ss;
  ss << f.rdbuf();
  return parse(ss.str());
}

}  // namespace gph
