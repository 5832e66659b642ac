// Minimal XML DOM for minimc input decks (test/minimc.xsd of the reference).
//
// The reference validates with Xerces-C and then walks a pugixml DOM
// (XMLDocument.cpp:74-90).  Neither library exists in this image, and the host
// only needs a small read-only tree: elements, attributes, one text value per
// element.  Comments, processing instructions and DOCTYPE are skipped,
// whitespace-only text is dropped, the five predefined entities and numeric
// character references are decoded.
#pragma once

#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace minimc::xml {

class Node {
public:
  const std::string& name() const { return name_; }
  // Text of the first character-data child ("" if none), like pugi's child_value().
  const std::string& text() const { return text_; }
  const Node* parent() const { return parent_; }
  const std::vector<std::unique_ptr<Node>>& children() const { return children_; }

  // First child element with this name, or nullptr.
  const Node* child(const std::string& name) const;
  const Node* first_child() const { return children_.empty() ? nullptr : children_.front().get(); }
  // First child whose attribute `key` equals `value`, or nullptr.
  const Node* child_by_attribute(const std::string& key, const std::string& value) const;

  bool has_attribute(const std::string& key) const { return find_attribute(key) != nullptr; }
  // Attribute value, or `fallback` when absent.
  std::string attribute(const std::string& key, const std::string& fallback = "") const;
  // strtod / strtoull of the attribute (0 / fallback when absent), as pugixml's as_double / as_ullong do.
  double attribute_double(const std::string& key, double fallback = 0.0) const;
  unsigned long long attribute_ull(const std::string& key, unsigned long long fallback = 0) const;

  // "/minimc/nuclides/multigroup/nuclide": the location text the reference puts
  // in front of its construction error messages (pugi::xml_node::path()).
  std::string path() const;
  // The document element this node belongs to.
  const Node& root() const;

private:
  friend class Parser;
  const std::string* find_attribute(const std::string& key) const;
  std::string name_, text_;
  bool has_text_ = false;
  std::vector<std::pair<std::string, std::string>> attributes_;
  std::vector<std::unique_ptr<Node>> children_;
  Node* parent_ = nullptr;
};

// Owns the tree.  Throws std::runtime_error("<what>: <description>") on I/O and
// syntax errors.
class Document {
public:
  static Document FromFile(const std::string& path);
  static Document FromString(const std::string& text, const std::string& what = "<string>");
  const Node& root() const { return *root_; }

private:
  std::unique_ptr<Node> root_;
};

}  // namespace minimc::xml
