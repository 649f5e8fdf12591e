// json.hpp — small self-contained JSON reader/writer for the host side (Prometheus API
// responses, Kubernetes objects from fixtures, patch bodies).  No third-party dependency.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace gph {

class Json {
 public:
  enum class Type { Null, Bool, Number, String, Array, Object };
  using Array = std::vector<Json>;
  using Member = std::pair<std::string, Json>;
  using Object = std::vector<Member>;  // insertion order preserved

  Json() = default;
  Json(std::nullptr_t) {}
  Json(bool b) : type_(Type::Bool), b_(b) {}
  Json(double d) : type_(Type::Number), n_(d) {}
  Json(int64_t i) : type_(Type::Number), n_((double)i) {}
  Json(int i) : type_(Type::Number), n_(i) {}
  Json(const char* s) : type_(Type::String), s_(s) {}
  Json(std::string s) : type_(Type::String), s_(std::move(s)) {}
  static Json array() { Json j; j.type_ = Type::Array; return j; }
  static Json object() { Json j; j.type_ = Type::Object; return j; }

  Type type() const { return type_; }
  bool is_null() const { return type_ == Type::Null; }
  bool is_string() const { return type_ == Type::String; }
  bool is_object() const { return type_ == Type::Object; }
  bool is_array() const { return type_ == Type::Array; }
  bool is_number() const { return type_ == Type::Number; }

  bool as_bool(bool dflt = false) const { return type_ == Type::Bool ? b_ : dflt; }
  double as_number(double dflt = 0) const { return type_ == Type::Number ? n_ : dflt; }
  const std::string& as_string() const { static const std::string e; return type_ == Type::String ? s_ : e; }
  const Array& items() const { static const Array e; return type_ == Type::Array ? a_ : e; }
  const Object& members() const { static const Object e; return type_ == Type::Object ? o_ : e; }

  // object access; returns a shared Null for anything missing so lookups chain safely
  const Json& operator[](const std::string& key) const;
  const Json& operator[](size_t i) const;
  const Json* find(const std::string& key) const;
  size_t size() const { return type_ == Type::Array ? a_.size() : type_ == Type::Object ? o_.size() : 0; }

  Json& set(const std::string& key, Json v);   // object
  Json& push(Json v);                          // array

  std::string dump() const;                    // compact, deterministic
  // throws std::runtime_error with offset.  struct_depth: objects nested less deep than this are decoded like structs —
  // a repeated member name is an error (a response envelope {status, data: {resultType, result}} is one in the
  // reference's decoder: struct_depth = 2); deeper objects are maps, the last value wins.
  static Json parse(const std::string& text, int struct_depth = 0);
  static Json parse_file(const std::string& path);

 private:
  Type type_ = Type::Null;
  bool b_ = false;
  double n_ = 0;
  std::string s_;
  Array a_;
  Object o_;
  void dump_to(std::string& out) const;
};

std::string json_escape(const std::string& s);

}  // namespace gph
