// TEST INFRASTRUCTURE — see ptree.hpp.  Boost's ini grammar as far as the shipped data/*.ini use it:
// [Section] headers, key = value, '#' or ';' comment lines, surrounding blanks trimmed.
#ifndef CRD_ORACLE_SHIM_BOOST_INI_PARSER_HPP
#define CRD_ORACLE_SHIM_BOOST_INI_PARSER_HPP
#include <fstream>
#include "ptree.hpp"

namespace boost { namespace property_tree { namespace ini_parser {

inline std::string trim_(const std::string &s) {
  size_t a = s.find_first_not_of(" \t\r\n");
  if (a == std::string::npos) return "";
  size_t b = s.find_last_not_of(" \t\r\n");
  return s.substr(a, b - a + 1);
}

inline void read_ini(const std::string &filename, ptree &pt) {
  std::ifstream in(filename.c_str());
  if (!in) throw std::runtime_error(filename + ": cannot open file");
  std::string line, section;
  while (std::getline(in, line)) {
    line = trim_(line);
    if (line.empty() || line[0] == '#' || line[0] == ';') continue;
    if (line[0] == '[') {
      size_t e = line.find(']');
      if (e == std::string::npos) throw std::runtime_error(filename + ": unmatched '['");
      section = trim_(line.substr(1, e - 1));
      continue;
    }
    size_t eq = line.find('=');
    if (eq == std::string::npos) throw std::runtime_error(filename + ": '=' character not found in line");
    std::string key = trim_(line.substr(0, eq)), val = trim_(line.substr(eq + 1));
    pt.kv[section.empty() ? key : section + "." + key] = val;
  }
}

}}}  // namespace boost::property_tree::ini_parser
#endif
