// ORACLE TEST INFRASTRUCTURE -- not product code.
//
// Xerces-free replacement for the reference's src/XMLDocument.cpp:74-90: the
// schema validation pass (xerces-c 3.2.3, not installed, not on the transport
// path) is skipped; the DOM is built by the pugixml stand-in in this directory.
#include "XMLDocument.hpp"

#include <stdexcept>

XMLDocument::XMLDocument(const std::filesystem::path& xml_filepath)
    : doc{ValidateXML(xml_filepath)} {}

pugi::xml_document
XMLDocument::ValidateXML(const std::filesystem::path& xml_filepath) {
  pugi::xml_document doc{};
  if (auto result = doc.load_file(xml_filepath.c_str()); !result) {
    throw std::runtime_error(result.description());
  }
  return doc;
}
