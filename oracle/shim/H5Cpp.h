// ORACLE TEST INFRASTRUCTURE -- not product code.
//
// Stand-in for the subset of the HDF5 1.8.22 C++ API that the reference's
// HDF5DataSet.hpp:88-130 calls. libhdf5 is not installed here and every *.hdf5
// under /root/reference is a git-lfs pointer, so instead of pandas-HDF5 this
// shim reads the flat "MMCTAB1" table file that this repo defines (see
// DESIGN.md, "table file format"): the same D sorted axes + row-major values
// that the pandas "fixed" layout carries, so HDF5DataSet<D> stays unmodified.
//
//   char     magic[8]  = "MMCTAB1\0"
//   uint64   ndim
//   uint64   shape[ndim]
//   double   axis_i[shape[i]]   for i in 0..ndim-1
//   double   values[prod(shape)]  (row-major, last axis fastest)
#pragma once

#include <cstdint>
#include <cstring>
#include <fstream>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#define H5F_ACC_RDONLY 0u

namespace H5 {

struct PredType {
  int id;
  static const PredType NATIVE_ULLONG;
  static const PredType NATIVE_DOUBLE;
};
inline const PredType PredType::NATIVE_ULLONG{1};
inline const PredType PredType::NATIVE_DOUBLE{2};

namespace detail {
struct Table {
  std::vector<std::vector<double>> axes;
  std::vector<double> values;
};
inline std::shared_ptr<const Table> Load(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("H5 shim: cannot open " + path);
  char magic[8];
  f.read(magic, 8);
  if (!f || std::memcmp(magic, "MMCTAB1\0", 8) != 0)
    throw std::runtime_error("H5 shim: " + path + " is not an MMCTAB1 table (git-lfs pointer?)");
  std::uint64_t ndim = 0;
  f.read(reinterpret_cast<char*>(&ndim), 8);
  if (!f || ndim == 0 || ndim > 8) throw std::runtime_error("H5 shim: bad ndim in " + path);
  std::vector<std::uint64_t> shape(ndim);
  f.read(reinterpret_cast<char*>(shape.data()), 8 * ndim);
  auto t = std::make_shared<Table>();
  std::uint64_t total = 1;
  for (auto n : shape) {
    t->axes.emplace_back(n);
    f.read(reinterpret_cast<char*>(t->axes.back().data()), 8 * n);
    total *= n;
  }
  t->values.resize(total);
  f.read(reinterpret_cast<char*>(t->values.data()), 8 * total);
  if (!f) throw std::runtime_error("H5 shim: truncated table " + path);
  return t;
}
}  // namespace detail

class DataSpace {
public:
  explicit DataSpace(long long n) : n(n) {}
  long long getSimpleExtentNpoints() const { return n; }

private:
  long long n;
};

class DataSet {
public:
  explicit DataSet(const std::vector<double>* v) : v(v) {}
  DataSpace getSpace() const { return DataSpace{static_cast<long long>(v->size())}; }
  void read(void* buf, const PredType&) const { std::memcpy(buf, v->data(), 8 * v->size()); }

private:
  const std::vector<double>* v;
};

class Attribute {
public:
  explicit Attribute(unsigned long long value) : value(value) {}
  void read(const PredType&, void* buf) const { std::memcpy(buf, &value, sizeof(value)); }

private:
  unsigned long long value;
};

class Group {
public:
  explicit Group(std::shared_ptr<const detail::Table> t) : t(std::move(t)) {}
  Attribute openAttribute(const std::string& name) const {
    if (name != "axis1_nlevels") throw std::runtime_error("H5 shim: unknown attribute " + name);
    return Attribute{t->axes.size()};
  }
  DataSet openDataSet(const std::string& name) const {
    if (name == "block0_values") return DataSet{&t->values};
    const std::string prefix = "axis1_level";
    if (name.compare(0, prefix.size(), prefix) == 0) {
      const auto i = std::stoul(name.substr(prefix.size()));
      return DataSet{&t->axes.at(i)};
    }
    throw std::runtime_error("H5 shim: unknown dataset " + name);
  }

private:
  std::shared_ptr<const detail::Table> t;
};

class H5File {
public:
  H5File(const std::string& path, unsigned) : t(detail::Load(path)) {}
  Group openGroup(const std::string&) const { return Group{t}; }

private:
  std::shared_ptr<const detail::Table> t;
};

}  // namespace H5
