// yaml-lite: the slice of the yaml-cpp API that OCPConfig / OptimalControlProblem
// read their configuration through (reference: src/OCP_config/OCPConfig.cpp:83-249,
// src/OptimalControlProblem.cpp:12-62).  yaml-cpp is not in this image, so this
// header-only stand-in keeps the names (`YAML::Node`, `YAML::Load`, `as<T>()`,
// `IsSequence()` ...) and parses the block/flow subset those configs use:
// nested maps, block and flow sequences, plain/quoted scalars, comments.
#pragma once

#include <cstdlib>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace YAML {

class Exception : public std::runtime_error {
 public:
  explicit Exception(const std::string& m) : std::runtime_error("yaml: " + m) {}
};
class BadConversion : public Exception {
 public:
  explicit BadConversion(const std::string& m) : Exception("bad conversion: " + m) {}
};
class InvalidNode : public Exception {
 public:
  explicit InvalidNode(const std::string& m) : Exception("invalid node: " + m) {}
};

namespace NodeType { enum value { Undefined, Null, Scalar, Sequence, Map }; }

class Node;
namespace detail {
struct NodeData {
  NodeType::value type = NodeType::Undefined;
  std::string scalar;
  std::vector<std::shared_ptr<NodeData>> seq;
  std::vector<std::pair<std::string, std::shared_ptr<NodeData>>> map;  // insertion ordered
};
struct iterator_value;
}  // namespace detail

class Node {
 public:
  Node() : d_(std::make_shared<detail::NodeData>()) { d_->type = NodeType::Null; }
  explicit Node(std::shared_ptr<detail::NodeData> d) : d_(std::move(d)) {}
  template <typename T, typename = typename std::enable_if<std::is_arithmetic<T>::value>::type>
  explicit Node(T v) : Node() { *this = v; }
  explicit Node(const std::string& s) : Node() { *this = s; }

  NodeType::value Type() const { return d_ ? d_->type : NodeType::Undefined; }
  bool IsDefined() const { return Type() != NodeType::Undefined; }
  bool IsNull() const { return Type() == NodeType::Null; }
  bool IsScalar() const { return Type() == NodeType::Scalar; }
  bool IsSequence() const { return Type() == NodeType::Sequence; }
  bool IsMap() const { return Type() == NodeType::Map; }
  explicit operator bool() const { return IsDefined(); }
  bool operator!() const { return !IsDefined(); }

  std::size_t size() const {
    if (IsSequence()) return d_->seq.size();
    if (IsMap()) return d_->map.size();
    return 0;
  }
  const std::string& Scalar() const { return d_->scalar; }

  // map access; a missing key yields an undefined node (which becomes a real
  // entry on assignment, as in yaml-cpp)
  Node operator[](const std::string& key) const { return get(key); }
  Node operator[](const char* key) const { return get(std::string(key)); }
  Node operator[](const std::string& key) { return get_or_create(key); }
  Node operator[](const char* key) { return get_or_create(std::string(key)); }
  template <typename I, typename = typename std::enable_if<std::is_integral<I>::value>::type>
  Node operator[](I idx) const {
    if (IsSequence()) {
      if (static_cast<std::size_t>(idx) >= d_->seq.size()) return undefined();
      return Node(d_->seq[static_cast<std::size_t>(idx)]);
    }
    if (IsMap()) return get(std::to_string(idx));
    return undefined();
  }

  template <typename T> T as() const;

  template <typename T> Node& operator=(const T& v) {
    std::ostringstream ss;
    ss.precision(17);
    ss << std::boolalpha << v;
    d_->type = NodeType::Scalar;
    d_->scalar = ss.str();
    d_->seq.clear(); d_->map.clear();
    return *this;
  }
  Node& operator=(const Node& o) {
    if (this != &o) {
      if (d_ && o.d_) *d_ = *o.d_; else d_ = o.d_;
    }
    return *this;
  }
  Node(const Node&) = default;
  void push_back(const Node& n) {
    if (!IsSequence()) { d_->type = NodeType::Sequence; d_->seq.clear(); d_->map.clear(); }
    d_->seq.push_back(n.d_);
  }

  class const_iterator;
  const_iterator begin() const;
  const_iterator end() const;

 private:
  static Node undefined() {
    auto d = std::make_shared<detail::NodeData>();
    d->type = NodeType::Undefined;
    return Node(d);
  }
  Node get(const std::string& key) const {
    if (!IsMap()) return undefined();
    for (const auto& kv : d_->map) if (kv.first == key) return Node(kv.second);
    return undefined();
  }
  Node get_or_create(const std::string& key) {
    if (!IsMap()) {
      if (IsDefined() && !IsNull()) throw InvalidNode("operator[] on a non-map node (key '" + key + "')");
      d_->type = NodeType::Map; d_->seq.clear();
    }
    for (const auto& kv : d_->map) if (kv.first == key) return Node(kv.second);
    auto d = std::make_shared<detail::NodeData>();
    d->type = NodeType::Undefined;
    d_->map.emplace_back(key, d);
    return Node(d);
  }
  std::shared_ptr<detail::NodeData> d_;
  friend class Parser;
};

namespace detail {
// like yaml-cpp: dereferencing an iterator gives something usable both as a
// Node (sequence element) and as a key/value pair (map entry)
struct iterator_value : public Node, public std::pair<Node, Node> {
  iterator_value() {}
  explicit iterator_value(const Node& n) : Node(n) {}
  iterator_value(const Node& k, const Node& v) : Node(v), std::pair<Node, Node>(k, v) {}
};
}  // namespace detail

class Node::const_iterator {
 public:
  const_iterator(const Node* n, std::size_t i) : n_(n), i_(i) {}
  detail::iterator_value operator*() const {
    if (n_->IsSequence()) return detail::iterator_value(Node(n_->d_->seq[i_]));
    Node k;
    k = n_->d_->map[i_].first;
    return detail::iterator_value(k, Node(n_->d_->map[i_].second));
  }
  struct arrow_proxy {
    detail::iterator_value v;
    detail::iterator_value* operator->() { return &v; }
  };
  arrow_proxy operator->() const { return arrow_proxy{**this}; }
  const_iterator& operator++() { ++i_; return *this; }
  bool operator!=(const const_iterator& o) const { return i_ != o.i_; }
  bool operator==(const const_iterator& o) const { return i_ == o.i_; }
 private:
  const Node* n_;
  std::size_t i_;
};
inline Node::const_iterator Node::begin() const { return const_iterator(this, 0); }
inline Node::const_iterator Node::end() const { return const_iterator(this, size()); }

// ---- conversions ----------------------------------------------------------
namespace detail {
inline std::string lower(std::string s) {
  for (char& c : s) c = static_cast<char>(std::tolower(static_cast<unsigned char>(c)));
  return s;
}
}  // namespace detail

template <> inline std::string Node::as<std::string>() const {
  if (IsNull()) return "null";
  if (!IsScalar()) throw BadConversion("node is not a scalar");
  return d_->scalar;
}
template <> inline double Node::as<double>() const {
  if (!IsScalar()) throw BadConversion("node is not a scalar");
  std::string s = detail::lower(d_->scalar);
  if (s == ".inf" || s == "+.inf") return std::numeric_limits<double>::infinity();
  if (s == "-.inf") return -std::numeric_limits<double>::infinity();
  if (s == ".nan") return std::numeric_limits<double>::quiet_NaN();
  char* end = nullptr;
  double v = std::strtod(d_->scalar.c_str(), &end);
  if (end == d_->scalar.c_str() || *end != '\0') throw BadConversion("'" + d_->scalar + "' is not a number");
  return v;
}
template <> inline float Node::as<float>() const { return static_cast<float>(as<double>()); }
template <> inline int Node::as<int>() const {
  if (!IsScalar()) throw BadConversion("node is not a scalar");
  char* end = nullptr;
  long v = std::strtol(d_->scalar.c_str(), &end, 0);
  if (end == d_->scalar.c_str() || *end != '\0') throw BadConversion("'" + d_->scalar + "' is not an integer");
  return static_cast<int>(v);
}
template <> inline long Node::as<long>() const { return as<int>(); }
template <> inline bool Node::as<bool>() const {
  if (!IsScalar()) throw BadConversion("node is not a scalar");
  std::string s = detail::lower(d_->scalar);
  if (s == "true" || s == "yes" || s == "on" || s == "y") return true;
  if (s == "false" || s == "no" || s == "off" || s == "n") return false;
  throw BadConversion("'" + d_->scalar + "' is not a bool");
}
template <> inline std::vector<double> Node::as<std::vector<double>>() const {
  if (!IsSequence()) throw BadConversion("node is not a sequence");
  std::vector<double> v;
  for (std::size_t i = 0; i < size(); ++i) v.push_back((*this)[i].as<double>());
  return v;
}

// ---- parser ---------------------------------------------------------------
class Parser {
 public:
  explicit Parser(const std::string& text) {
    std::istringstream in(text);
    std::string raw;
    int lineno = 0;
    while (std::getline(in, raw)) {
      ++lineno;
      std::string s = strip_comment(raw);
      std::size_t ind = 0;
      while (ind < s.size() && s[ind] == ' ') ++ind;
      if (ind < s.size() && s[ind] == '\t') throw Exception("tab indentation at line " + std::to_string(lineno));
      std::string body = rtrim(s.substr(ind));
      if (body.empty() || body == "---" || body == "...") continue;
      lines_.push_back({static_cast<int>(ind), body, lineno});
    }
  }
  Node parse() {
    if (lines_.empty()) return Node();
    std::size_t pos = 0;
    auto d = block(pos, lines_[0].indent);
    if (pos != lines_.size()) throw Exception("unexpected content at line " + std::to_string(lines_[pos].no));
    return Node(d);
  }

 private:
  struct Line { int indent; std::string body; int no; };
  typedef std::shared_ptr<detail::NodeData> P;
  std::vector<Line> lines_;

  static std::string rtrim(std::string s) {
    while (!s.empty() && (s.back() == ' ' || s.back() == '\r' || s.back() == '\t')) s.pop_back();
    return s;
  }
  static std::string trim(std::string s) {
    s = rtrim(s);
    std::size_t i = 0;
    while (i < s.size() && (s[i] == ' ' || s[i] == '\t')) ++i;
    return s.substr(i);
  }
  static std::string strip_comment(const std::string& s) {
    char q = 0;
    for (std::size_t i = 0; i < s.size(); ++i) {
      char c = s[i];
      if (q) { if (c == q) q = 0; continue; }
      if (c == '"' || c == '\'') { q = c; continue; }
      if (c == '#' && (i == 0 || s[i - 1] == ' ' || s[i - 1] == '\t')) return s.substr(0, i);
    }
    return s;
  }
  static P make(NodeType::value t) { auto d = std::make_shared<detail::NodeData>(); d->type = t; return d; }
  static P scalar(const std::string& raw) {
    std::string s = trim(raw);
    if (s.empty() || s == "~" || s == "null") return make(NodeType::Null);
    P d = make(NodeType::Scalar);
    if (s.size() >= 2 && ((s.front() == '"' && s.back() == '"') || (s.front() == '\'' && s.back() == '\'')))
      d->scalar = s.substr(1, s.size() - 2);
    else
      d->scalar = s;
    return d;
  }
  // position of the ':' that separates a key from its value, or npos
  static std::size_t key_colon(const std::string& s) {
    char q = 0;
    int depth = 0;
    for (std::size_t i = 0; i < s.size(); ++i) {
      char c = s[i];
      if (q) { if (c == q) q = 0; continue; }
      if (c == '"' || c == '\'') { q = c; continue; }
      if (c == '[' || c == '{') ++depth;
      if (c == ']' || c == '}') --depth;
      if (c == ':' && depth == 0 && (i + 1 == s.size() || s[i + 1] == ' ')) return i;
    }
    return std::string::npos;
  }
  static P flow(const std::string& s, std::size_t& i) {
    auto skip = [&]() { while (i < s.size() && s[i] == ' ') ++i; };
    skip();
    if (i < s.size() && s[i] == '[') {
      P d = make(NodeType::Sequence);
      ++i; skip();
      if (i < s.size() && s[i] == ']') { ++i; return d; }
      while (true) {
        d->seq.push_back(flow(s, i));
        skip();
        if (i >= s.size()) throw Exception("unterminated flow sequence: " + s);
        if (s[i] == ',') { ++i; continue; }
        if (s[i] == ']') { ++i; break; }
        throw Exception("bad flow sequence: " + s);
      }
      return d;
    }
    if (i < s.size() && s[i] == '{') {
      P d = make(NodeType::Map);
      ++i; skip();
      if (i < s.size() && s[i] == '}') { ++i; return d; }
      while (true) {
        skip();
        std::size_t k0 = i;
        while (i < s.size() && s[i] != ':') ++i;
        if (i >= s.size()) throw Exception("bad flow map: " + s);
        std::string key = scalar(s.substr(k0, i - k0))->scalar;
        ++i;
        d->map.emplace_back(key, flow(s, i));
        skip();
        if (i >= s.size()) throw Exception("unterminated flow map: " + s);
        if (s[i] == ',') { ++i; continue; }
        if (s[i] == '}') { ++i; break; }
        throw Exception("bad flow map: " + s);
      }
      return d;
    }
    // plain / quoted scalar up to , ] }
    std::size_t b = i;
    char q = 0;
    while (i < s.size()) {
      char c = s[i];
      if (q) { if (c == q) q = 0; ++i; continue; }
      if (c == '"' || c == '\'') { q = c; ++i; continue; }
      if (c == ',' || c == ']' || c == '}') break;
      ++i;
    }
    return scalar(s.substr(b, i - b));
  }
  static P inline_value(const std::string& s) {
    std::string t = trim(s);
    if (!t.empty() && (t[0] == '[' || t[0] == '{')) {
      std::size_t i = 0;
      P d = flow(t, i);
      return d;
    }
    return scalar(t);
  }

  P block(std::size_t& pos, int indent) {
    if (pos >= lines_.size()) return make(NodeType::Null);
    const Line& first = lines_[pos];
    if (first.body.rfind("- ", 0) == 0 || first.body == "-") return sequence(pos, indent);
    if (key_colon(first.body) != std::string::npos) return mapping(pos, indent);
    P d = inline_value(first.body);
    ++pos;
    return d;
  }

  P sequence(std::size_t& pos, int indent) {
    P d = make(NodeType::Sequence);
    while (pos < lines_.size() && lines_[pos].indent == indent &&
           (lines_[pos].body.rfind("- ", 0) == 0 || lines_[pos].body == "-")) {
      std::string rest = lines_[pos].body.size() > 1 ? lines_[pos].body.substr(2) : "";
      std::size_t extra = 0;
      while (extra < rest.size() && rest[extra] == ' ') ++extra;
      rest = rest.substr(extra);
      int child_indent = indent + 2 + static_cast<int>(extra);
      if (rest.empty()) {
        ++pos;
        if (pos < lines_.size() && lines_[pos].indent > indent) d->seq.push_back(block(pos, lines_[pos].indent));
        else d->seq.push_back(make(NodeType::Null));
      } else if (key_colon(rest) != std::string::npos && rest[0] != '[' && rest[0] != '{') {
        // "- key: value" opens a map whose further keys sit at child_indent
        lines_[pos].indent = child_indent;
        lines_[pos].body = rest;
        d->seq.push_back(mapping(pos, child_indent));
      } else {
        d->seq.push_back(inline_value(rest));
        ++pos;
      }
    }
    return d;
  }

  P mapping(std::size_t& pos, int indent) {
    P d = make(NodeType::Map);
    while (pos < lines_.size() && lines_[pos].indent == indent) {
      const std::string body = lines_[pos].body;
      if (body.rfind("- ", 0) == 0 || body == "-") break;
      std::size_t c = key_colon(body);
      if (c == std::string::npos) throw Exception("expected 'key: value' at line " + std::to_string(lines_[pos].no));
      std::string key = scalar(body.substr(0, c))->scalar;
      std::string rest = trim(body.substr(c + 1));
      ++pos;
      P val;
      if (!rest.empty()) {
        val = inline_value(rest);
      } else if (pos < lines_.size() && lines_[pos].indent > indent) {
        val = block(pos, lines_[pos].indent);
      } else if (pos < lines_.size() && lines_[pos].indent == indent &&
                 (lines_[pos].body.rfind("- ", 0) == 0 || lines_[pos].body == "-")) {
        val = sequence(pos, indent);  // sequence written at the key's own indentation
      } else {
        val = make(NodeType::Null);
      }
      bool replaced = false;
      for (auto& kv : d->map) if (kv.first == key) { kv.second = val; replaced = true; }
      if (!replaced) d->map.emplace_back(key, val);
    }
    if (pos < lines_.size() && lines_[pos].indent > indent)
      throw Exception("bad indentation at line " + std::to_string(lines_[pos].no));
    return d;
  }
};

inline Node Load(const std::string& text) { return Parser(text).parse(); }
inline Node Load(const char* text) { return Parser(std::string(text)).parse(); }
inline Node LoadFile(const std::string& path) {
  std::ifstream f(path);
  if (!f.good()) throw Exception("cannot open file '" + path + "'");
  std::stringstream ss;
  ss << f.rdbuf();
  return Load(ss.str());
}

}  // namespace YAML
