#include "xml.hpp"

#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace minimc::xml {

const Node* Node::child(const std::string& name) const {
  for (const auto& c : children_)
    if (c->name_ == name) return c.get();
  return nullptr;
}

const Node* Node::child_by_attribute(const std::string& key, const std::string& value) const {
  for (const auto& c : children_) {
    const std::string* v = c->find_attribute(key);
    if (v && *v == value) return c.get();
  }
  return nullptr;
}

const std::string* Node::find_attribute(const std::string& key) const {
  for (const auto& kv : attributes_)
    if (kv.first == key) return &kv.second;
  return nullptr;
}

std::string Node::attribute(const std::string& key, const std::string& fallback) const {
  const std::string* v = find_attribute(key);
  return v ? *v : fallback;
}

double Node::attribute_double(const std::string& key, double fallback) const {
  const std::string* v = find_attribute(key);
  return v ? std::strtod(v->c_str(), nullptr) : fallback;
}

unsigned long long Node::attribute_ull(const std::string& key, unsigned long long fallback) const {
  const std::string* v = find_attribute(key);
  return v ? std::strtoull(v->c_str(), nullptr, 10) : fallback;
}

std::string Node::path() const {
  std::string result;
  for (const Node* n = this; n; n = n->parent_) result = "/" + n->name_ + result;
  return result;
}

const Node& Node::root() const {
  const Node* n = this;
  while (n->parent_) n = n->parent_;
  return *n;
}

class Parser {
public:
  Parser(const std::string& text, const std::string& what) : s(text), what(what) {}

  std::unique_ptr<Node> ParseDocument() {
    SkipMisc();
    if (pos >= s.size() || s[pos] != '<') Fail("no document element");
    auto root = ParseElement(nullptr);
    SkipMisc();
    if (pos != s.size()) Fail("content after the document element");
    return root;
  }

private:
  [[noreturn]] void Fail(const std::string& description) const {
    size_t line = 1;
    for (size_t i = 0; i < pos && i < s.size(); i++) line += s[i] == '\n';
    throw std::runtime_error(what + ": XML parse error at line " + std::to_string(line) + ": " + description);
  }
  bool StartsWith(const char* lit) const { return s.compare(pos, std::strlen(lit), lit) == 0; }
  static bool IsSpace(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; }
  static bool IsNameChar(char c) {
    return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9') || c == '_' || c == '-' ||
           c == '.' || c == ':' || static_cast<unsigned char>(c) >= 0x80;
  }
  void SkipSpace() {
    while (pos < s.size() && IsSpace(s[pos])) pos++;
  }
  void SkipUntil(const char* terminator) {
    const size_t end = s.find(terminator, pos);
    if (end == std::string::npos) Fail(std::string("unterminated construct, expected ") + terminator);
    pos = end + std::strlen(terminator);
  }
  // whitespace, comments, processing instructions, DOCTYPE
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
    const size_t begin = pos;
    while (pos < s.size() && IsNameChar(s[pos])) pos++;
    if (pos == begin) Fail("expected a name");
    return s.substr(begin, pos - begin);
  }
  std::string Decode(const std::string& raw) const {
    if (raw.find('&') == std::string::npos) return raw;
    std::string out;
    for (size_t i = 0; i < raw.size(); i++) {
      if (raw[i] != '&') {
        out += raw[i];
        continue;
      }
      const size_t semi = raw.find(';', i);
      if (semi == std::string::npos) Fail("unterminated entity reference");
      const std::string ent = raw.substr(i + 1, semi - i - 1);
      if (ent == "lt") out += '<';
      else if (ent == "gt") out += '>';
      else if (ent == "amp") out += '&';
      else if (ent == "quot") out += '"';
      else if (ent == "apos") out += '\'';
      else if (!ent.empty() && ent[0] == '#') {
        const unsigned long code = ent.size() > 1 && (ent[1] == 'x' || ent[1] == 'X')
                                       ? std::strtoul(ent.c_str() + 2, nullptr, 16)
                                       : std::strtoul(ent.c_str() + 1, nullptr, 10);
        // UTF-8 encode
        if (code < 0x80) out += static_cast<char>(code);
        else if (code < 0x800) {
          out += static_cast<char>(0xC0 | (code >> 6));
          out += static_cast<char>(0x80 | (code & 0x3F));
        } else if (code < 0x10000) {
          out += static_cast<char>(0xE0 | (code >> 12));
          out += static_cast<char>(0x80 | ((code >> 6) & 0x3F));
          out += static_cast<char>(0x80 | (code & 0x3F));
        } else {
          out += static_cast<char>(0xF0 | (code >> 18));
          out += static_cast<char>(0x80 | ((code >> 12) & 0x3F));
          out += static_cast<char>(0x80 | ((code >> 6) & 0x3F));
          out += static_cast<char>(0x80 | (code & 0x3F));
        }
      } else {
        Fail("unknown entity &" + ent + ";");
      }
      i = semi;
    }
    return out;
  }
  static bool AllSpace(const std::string& t) {
    for (char c : t)
      if (!IsSpace(c)) return false;
    return true;
  }

  std::unique_ptr<Node> ParseElement(Node* parent) {
    pos++;  // '<'
    auto node = std::make_unique<Node>();
    node->parent_ = parent;
    node->name_ = ParseName();
    while (true) {
      SkipSpace();
      if (pos >= s.size()) Fail("unterminated start tag <" + node->name_);
      if (s[pos] == '/') {
        if (!StartsWith("/>")) Fail("expected />");
        pos += 2;
        return node;
      }
      if (s[pos] == '>') {
        pos++;
        break;
      }
      const std::string key = ParseName();
      SkipSpace();
      if (pos >= s.size() || s[pos] != '=') Fail("expected = after attribute " + key);
      pos++;
      SkipSpace();
      if (pos >= s.size() || (s[pos] != '"' && s[pos] != '\'')) Fail("expected a quoted value for attribute " + key);
      const char quote = s[pos++];
      const size_t end = s.find(quote, pos);
      if (end == std::string::npos) Fail("unterminated value of attribute " + key);
      if (node->find_attribute(key)) Fail("duplicate attribute " + key);
      node->attributes_.emplace_back(key, Decode(s.substr(pos, end - pos)));
      pos = end + 1;
    }
    // content
    while (true) {
      if (pos >= s.size()) Fail("unterminated element <" + node->name_ + ">");
      if (s[pos] != '<') {
        const size_t end = s.find('<', pos);
        const std::string raw = s.substr(pos, end == std::string::npos ? std::string::npos : end - pos);
        if (!AllSpace(raw) && !node->has_text_) {
          node->text_ = Decode(raw);
          node->has_text_ = true;
        }
        pos = end == std::string::npos ? s.size() : end;
      } else if (StartsWith("<!--")) {
        SkipUntil("-->");
      } else if (StartsWith("<![CDATA[")) {
        const size_t begin = pos + 9;
        SkipUntil("]]>");
        if (!node->has_text_) {
          node->text_ = s.substr(begin, pos - 3 - begin);
          node->has_text_ = true;
        }
      } else if (StartsWith("<?")) {
        SkipUntil("?>");
      } else if (StartsWith("</")) {
        pos += 2;
        const std::string closing = ParseName();
        if (closing != node->name_) Fail("mismatched end tag </" + closing + "> for <" + node->name_ + ">");
        SkipSpace();
        if (pos >= s.size() || s[pos] != '>') Fail("expected > in end tag");
        pos++;
        return node;
      } else {
        node->children_.push_back(ParseElement(node.get()));
      }
    }
  }

  const std::string& s;
  const std::string& what;
  size_t pos = 0;
};

Document Document::FromFile(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error(path + ": File was not found");
  std::stringstream buffer;
  buffer << f.rdbuf();
  return FromString(buffer.str(), path);
}

Document Document::FromString(const std::string& text, const std::string& what) {
  Document doc;
  doc.root_ = Parser(text, what).ParseDocument();
  return doc;
}

}  // namespace minimc::xml
