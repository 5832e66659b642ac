// KEigenvalue: power iteration on the GPU.  The reference's KEigenvalue::Solve
// is a stub (KEigenvalue.cpp:36-62); DESIGN.md "k-eigenvalue"
// defines the algorithm implemented here.
#include <cmath>
#include <stdexcept>

#include "minimc.hpp"

namespace minimc {

namespace {
const xml::Node& KNode(const xml::Node& root) {
  const xml::Node* problemtype = root.child("problemtype");
  const xml::Node* node = problemtype ? problemtype->child("keigenvalue") : nullptr;
  if (!node) throw std::runtime_error("/minimc/problemtype: \"keigenvalue\" node not found");
  return *node;
}
const xml::Node& InitialSource(const xml::Node& root) {
  const xml::Node* node = KNode(root).child("initialsource");
  if (!node) throw std::runtime_error("/minimc/problemtype/keigenvalue: \"initialsource\" node not found");
  return *node;
}
}  // namespace

KEigenvalue::KEigenvalue(const xml::Node& root)
    : Driver{root}, last_inactive{KNode(root).attribute_ull("inactive")},
      last_active{last_inactive + KNode(root).attribute_ull("active")}, source{InitialSource(root)} {}

EstimatorSet KEigenvalue::Solve() { throw std::runtime_error("KEigenvalue::Solve: not available in this build"); }

}  // namespace minimc
