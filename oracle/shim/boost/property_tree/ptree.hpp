// TEST INFRASTRUCTURE — the two Boost.PropertyTree calls the reference's main() makes
// (FHNmodel_torus.cpp:158-174): read_ini(path, pt) and pt.get<T>("Section.key").  A missing key
// throws, like Boost's ptree_bad_path.  Not a product component (the product's own ini reader is
// crdmodel_b200/host/crd_ini.hpp).
#ifndef CRD_ORACLE_SHIM_BOOST_PTREE_HPP
#define CRD_ORACLE_SHIM_BOOST_PTREE_HPP
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>

namespace boost { namespace property_tree {

class ptree_bad_path : public std::runtime_error {
 public:
  explicit ptree_bad_path(const std::string &w) : std::runtime_error(w) {}
};

class ptree {
 public:
  std::map<std::string, std::string> kv;
  template <class T> T get(const std::string &path) const {
    auto it = kv.find(path);
    if (it == kv.end()) throw ptree_bad_path("No such node (" + path + ")");
    std::istringstream is(it->second);
    T v{};
    is >> v;
    if (is.fail()) throw std::runtime_error("conversion of data failed (" + path + ")");
    return v;
  }
};
template <> inline std::string ptree::get<std::string>(const std::string &path) const {
  auto it = kv.find(path);
  if (it == kv.end()) throw ptree_bad_path("No such node (" + path + ")");
  return it->second;
}

}}  // namespace boost::property_tree
#endif
