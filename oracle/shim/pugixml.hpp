// ORACLE TEST INFRASTRUCTURE -- not product code.
//
// Stand-in for the subset of pugixml 1.12.1 (pinned in the reference's
// CMakeLists.txt:96) that agtumulak/minimc calls, so the reference's unmodified
// translation units can be compiled here (no network, no pugixml). Written from
// the usage census in SURVEY.md section 8(c); behaviour mirrored:
//   * whitespace-only text is dropped, comments / PIs / DOCTYPE are skipped
//     (pugixml parse_default);
//   * child_value() = text of the first PCDATA child, "" when there is none;
//   * iteration visits element children in document order;
//   * path() = '/'-joined element names from the document node;
//   * null handles are safe to chain (child() of a null node is a null node).
#pragma once

#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

namespace pugi {

namespace detail {
struct Attr {
  std::string name, value;
};
struct Node {
  std::string name;            // empty for the document node
  std::string text;            // first PCDATA child
  bool has_text = false;
  std::vector<Attr> attrs;
  std::vector<std::unique_ptr<Node>> kids;
  Node* parent = nullptr;
};

class Parser {
public:
  explicit Parser(const std::string& s) : s(s) {}
  bool Parse(Node& doc, std::string& err) {
    try {
      SkipMisc();
      while (i < s.size()) {
        if (s[i] != '<') throw std::string("text outside of root element");
        doc.kids.push_back(ParseElement(&doc));
        SkipMisc();
      }
      if (doc.kids.empty()) throw std::string("No document element found");
    } catch (const std::string& e) {
      err = e;
      return false;
    }
    return true;
  }

private:
  const std::string& s;
  size_t i = 0;
  static bool IsSpace(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; }
  void SkipSpace() {
    while (i < s.size() && IsSpace(s[i])) i++;
  }
  bool StartsWith(const char* p) const { return s.compare(i, std::strlen(p), p) == 0; }
  void SkipUntil(const char* end) {
    const auto pos = s.find(end, i);
    if (pos == std::string::npos) throw std::string("unterminated construct");
    i = pos + std::strlen(end);
  }
  // whitespace, comments, processing instructions, doctype
  void SkipMisc() {
    while (true) {
      SkipSpace();
      if (StartsWith("<!--")) SkipUntil("-->");
      else if (StartsWith("<?")) SkipUntil("?>");
      else if (StartsWith("<!DOCTYPE")) SkipUntil(">");
      else return;
    }
  }
  std::string ParseName() {
    const size_t b = i;
    while (i < s.size() && !IsSpace(s[i]) && s[i] != '>' && s[i] != '/' && s[i] != '=') i++;
    if (i == b) throw std::string("expected name");
    return s.substr(b, i - b);
  }
  static std::string Unescape(const std::string& in) {
    std::string out;
    for (size_t k = 0; k < in.size(); k++) {
      if (in[k] != '&') { out += in[k]; continue; }
      const auto semi = in.find(';', k);
      const std::string ent = semi == std::string::npos ? "" : in.substr(k + 1, semi - k - 1);
      if (ent == "lt") out += '<';
      else if (ent == "gt") out += '>';
      else if (ent == "amp") out += '&';
      else if (ent == "quot") out += '"';
      else if (ent == "apos") out += '\'';
      else { out += in[k]; continue; }
      k = semi;
    }
    return out;
  }
  std::unique_ptr<Node> ParseElement(Node* parent) {
    auto n = std::make_unique<Node>();
    n->parent = parent;
    i++;  // '<'
    n->name = ParseName();
    while (true) {
      SkipSpace();
      if (i >= s.size()) throw std::string("unterminated start tag");
      if (s[i] == '/') {
        if (i + 1 >= s.size() || s[i + 1] != '>') throw std::string("bad empty tag");
        i += 2;
        return n;
      }
      if (s[i] == '>') { i++; break; }
      Attr a;
      a.name = ParseName();
      SkipSpace();
      if (i >= s.size() || s[i] != '=') throw std::string("expected '=' in attribute");
      i++;
      SkipSpace();
      if (i >= s.size() || (s[i] != '"' && s[i] != '\'')) throw std::string("expected quote");
      const char q = s[i++];
      const auto e = s.find(q, i);
      if (e == std::string::npos) throw std::string("unterminated attribute");
      a.value = Unescape(s.substr(i, e - i));
      i = e + 1;
      n->attrs.push_back(std::move(a));
    }
    // content
    while (true) {
      if (i >= s.size()) throw std::string("unterminated element " + n->name);
      if (s[i] == '<') {
        if (StartsWith("<!--")) { SkipUntil("-->"); continue; }
        if (StartsWith("<?")) { SkipUntil("?>"); continue; }
        if (StartsWith("<![CDATA[")) {
          const auto e = s.find("]]>", i);
          if (e == std::string::npos) throw std::string("unterminated CDATA");
          if (!n->has_text) { n->text = s.substr(i + 9, e - i - 9); n->has_text = true; }
          i = e + 3;
          continue;
        }
        if (StartsWith("</")) {
          i += 2;
          const auto name = ParseName();
          if (name != n->name) throw std::string("Start-end tags mismatch");
          SkipSpace();
          if (i >= s.size() || s[i] != '>') throw std::string("bad end tag");
          i++;
          return n;
        }
        n->kids.push_back(ParseElement(n.get()));
      } else {
        const auto e = s.find('<', i);
        if (e == std::string::npos) throw std::string("unterminated element " + n->name);
        const std::string raw = s.substr(i, e - i);
        i = e;
        bool all_space = true;
        for (char c : raw) all_space = all_space && IsSpace(c);
        if (!all_space && !n->has_text) { n->text = Unescape(raw); n->has_text = true; }
      }
    }
  }
};
}  // namespace detail

class xml_attribute {
public:
  xml_attribute() = default;
  explicit xml_attribute(const detail::Attr* a) : a(a) {}
  explicit operator bool() const { return a != nullptr; }
  bool operator!() const { return a == nullptr; }
  bool empty() const { return a == nullptr; }
  const char* as_string(const char* def = "") const { return a ? a->value.c_str() : def; }
  double as_double(double def = 0) const { return a ? std::strtod(a->value.c_str(), nullptr) : def; }
  unsigned int as_uint(unsigned int def = 0) const {
    return a ? static_cast<unsigned int>(std::strtoul(a->value.c_str(), nullptr, 10)) : def;
  }
  unsigned long long as_ullong(unsigned long long def = 0) const {
    return a ? std::strtoull(a->value.c_str(), nullptr, 10) : def;
  }
  const char* value() const { return as_string(); }

private:
  const detail::Attr* a = nullptr;
};

class xml_node;

class xpath_node {
public:
  xpath_node() = default;
  explicit xpath_node(const detail::Node* n) : n(n) {}
  xml_node node() const;

private:
  const detail::Node* n = nullptr;
};

class xml_node_iterator;

class xml_node {
public:
  xml_node() = default;
  explicit xml_node(const detail::Node* n) : n(n) {}
  explicit operator bool() const { return n != nullptr; }
  bool operator!() const { return n == nullptr; }
  bool empty() const { return n == nullptr; }
  bool operator==(const xml_node& o) const { return n == o.n; }
  bool operator!=(const xml_node& o) const { return n != o.n; }
  const char* name() const { return n ? n->name.c_str() : ""; }
  const char* child_value() const { return (n && n->has_text) ? n->text.c_str() : ""; }
  xml_node child(const char* name) const {
    if (n)
      for (const auto& k : n->kids)
        if (k->name == name) return xml_node{k.get()};
    return {};
  }
  xml_node first_child() const { return (n && !n->kids.empty()) ? xml_node{n->kids.front().get()} : xml_node{}; }
  xml_node parent() const { return n ? xml_node{n->parent} : xml_node{}; }
  xml_node root() const {
    const detail::Node* r = n;
    while (r && r->parent) r = r->parent;
    return xml_node{r};
  }
  xml_attribute attribute(const char* name) const {
    if (n)
      for (const auto& a : n->attrs)
        if (a.name == name) return xml_attribute{&a};
    return {};
  }
  xml_node find_child_by_attribute(const char* attr_name, const char* attr_value) const {
    if (n)
      for (const auto& k : n->kids)
        for (const auto& a : k->attrs)
          if (a.name == attr_name && a.value == attr_value) return xml_node{k.get()};
    return {};
  }
  std::string path(char delimiter = '/') const {
    if (!n) return "";
    std::string result;
    for (const detail::Node* c = n; c && c->parent; c = c->parent) {
      result = std::string(1, delimiter) + c->name + result;
    }
    return result;
  }
  // Only simple relative or absolute step paths ("a/b/c") are supported; the
  // reference uses exactly one: "minimc/general/particles" (Interaction.cpp:20).
  xpath_node select_node(const char* query) const {
    xml_node cur = *this;
    std::string q{query};
    if (!q.empty() && q[0] == '/') {
      cur = root();
      q = q.substr(1);
    }
    std::stringstream ss{q};
    std::string step;
    while (std::getline(ss, step, '/')) {
      if (step.empty()) continue;
      cur = cur.child(step.c_str());
    }
    return xpath_node{cur.n};
  }
  xml_node_iterator begin() const;
  xml_node_iterator end() const;

protected:
  const detail::Node* n = nullptr;
  friend class xml_node_iterator;
};

inline xml_node xpath_node::node() const { return xml_node{n}; }

class xml_node_iterator {
public:
  using value_type = xml_node;
  using difference_type = std::ptrdiff_t;
  using pointer = const xml_node*;
  using reference = const xml_node&;
  using iterator_category = std::forward_iterator_tag;
  xml_node_iterator() = default;
  xml_node_iterator(const detail::Node* parent, size_t idx) : parent(parent), idx(idx) { Load(); }
  reference operator*() const { return cur; }
  pointer operator->() const { return &cur; }
  xml_node_iterator& operator++() {
    idx++;
    Load();
    return *this;
  }
  xml_node_iterator operator++(int) {
    auto t = *this;
    ++*this;
    return t;
  }
  bool operator==(const xml_node_iterator& o) const { return parent == o.parent && idx == o.idx; }
  bool operator!=(const xml_node_iterator& o) const { return !(*this == o); }

private:
  void Load() { cur = (parent && idx < parent->kids.size()) ? xml_node{parent->kids[idx].get()} : xml_node{}; }
  const detail::Node* parent = nullptr;
  size_t idx = 0;
  xml_node cur;
};

inline xml_node_iterator xml_node::begin() const { return {n, 0}; }
inline xml_node_iterator xml_node::end() const { return {n, n ? n->kids.size() : 0}; }

struct xml_parse_result {
  bool ok = false;
  std::string message;
  explicit operator bool() const { return ok; }
  const char* description() const { return message.c_str(); }
};

class xml_document : public xml_node {
public:
  xml_document() : store(std::make_unique<detail::Node>()) { n = store.get(); }
  xml_document(xml_document&& o) noexcept : xml_node{o.store.get()}, store(std::move(o.store)) { o.n = nullptr; }
  xml_document& operator=(xml_document&& o) noexcept {
    store = std::move(o.store);
    n = store.get();
    o.n = nullptr;
    return *this;
  }
  xml_document(const xml_document&) = delete;
  xml_document& operator=(const xml_document&) = delete;
  xml_parse_result load_file(const char* path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return {false, "File was not found"};
    std::stringstream ss;
    ss << f.rdbuf();
    return load_string(ss.str().c_str());
  }
  xml_parse_result load_string(const char* contents) {
    store = std::make_unique<detail::Node>();
    n = store.get();
    const std::string s{contents};
    std::string err;
    detail::Parser p{s};
    if (!p.Parse(*store, err)) return {false, err};
    return {true, "No error"};
  }

private:
  std::unique_ptr<detail::Node> store;
};

}  // namespace pugi
