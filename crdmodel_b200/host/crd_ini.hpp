// crd_ini.hpp — the ini reader of the host drivers.  Same file format and the same lookup contract as the
// reference's Boost.PropertyTree use (src/FHNmodel_torus.cpp:158-174): [Section] headers, key = value,
// '#' / ';' comment lines, get<T>("Section.key") throws when the key is missing or when the WHOLE value does not convert
// (Boost's stream translator: `400abc` or `1e3` for an int is an error, not 400 / 1), a key defined twice in a section
// is an error (Boost: "duplicate key name").
#pragma once
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>

namespace crd {

class Ini {
 public:
  explicit Ini(const std::string &path) {
    std::ifstream in(path.c_str());
    if (!in) throw std::runtime_error(path + ": cannot open file");
    std::string line, section;
    while (std::getline(in, line)) {
      line = trim(line);
      if (line.empty() || line[0] == '#' || line[0] == ';') continue;
      if (line[0] == '[') {
        const size_t e = line.find(']');
        if (e == std::string::npos) throw std::runtime_error(path + ": unmatched '['");
        section = trim(line.substr(1, e - 1));
        continue;
      }
      const size_t eq = line.find('=');
      if (eq == std::string::npos) throw std::runtime_error(path + ": '=' character not found in line");
      const std::string key = section.empty() ? trim(line.substr(0, eq)) : section + "." + trim(line.substr(0, eq));
      if (kv_.count(key)) throw std::runtime_error(path + ": duplicate key name (" + key + ")");
      kv_[key] = trim(line.substr(eq + 1));
    }
  }

  bool has(const std::string &key) const { return kv_.count(key) != 0; }

  template <class T> T get(const std::string &key) const {
    auto it = kv_.find(key);
    if (it == kv_.end()) throw std::runtime_error("No such node (" + key + ")");
    std::istringstream is(it->second);
    T v{};
    is >> v;
    if (is.fail() || !(is >> std::ws).eof()) throw std::runtime_error("conversion of data to type failed (" + key + ")");
    return v;
  }
  template <class T> T get(const std::string &key, const T &fallback) const { return has(key) ? get<T>(key) : fallback; }
  std::string str(const std::string &key) const {
    auto it = kv_.find(key);
    if (it == kv_.end()) throw std::runtime_error("No such node (" + key + ")");
    return it->second;
  }

 private:
  static std::string trim(const std::string &s) {
    const size_t a = s.find_first_not_of(" \t\r\n");
    if (a == std::string::npos) return "";
    const size_t b = s.find_last_not_of(" \t\r\n");
    return s.substr(a, b - a + 1);
  }
  std::map<std::string, std::string> kv_;
};

}  // namespace crd
